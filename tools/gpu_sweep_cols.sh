for C in 1 2 4 8; do for L in 8 16 32; do for S in 4 8 16; do
  [ $C -gt $S ] && continue
  echo "== C=$C L=$L S=$S"; KF_RED_C=$C KF_RED_LPR=$L KF_RED_S=$S timeout 60 python tools/gpu_mem_ops.py sum_dim0 2>&1 | tail -1
done; done; done

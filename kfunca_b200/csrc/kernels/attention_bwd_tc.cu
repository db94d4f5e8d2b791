// Causal attention BACKWARD on tcgen05 tensor cores (bf16 / fp16, head size 64 or 128), hand-written PTX.
// The reference has no attention backward at all (only AddGradFunction exists, src/core/binary_ops.cpp:16-43;
// SURVEY F3); the oracle is the float64 restatement in oracle/oracle.py (causal_attention_bwd).
//
// Deterministic two-kernel scheme, no atomics (one template, two modes):
//   MODE_DKV : one CTA per 128-row KV block j.  K_j, V_j stay in shared memory; 64-row query tiles (Q_t, dO_t) stream
//              through a TMA ring.   S^T = K_j Q_t^T,  dP^T = V_j dO_t^T   (128 x 64, fp32 in TMEM, lanes = kv rows)
//              P^T = exp2(S^T c - lse_q),  dS^T = P^T o (dP^T - delta_q)  -> 16-bit, written back over S^T / dP^T in TMEM
//              dV_j += P^T dO_t,  dK_j += dS^T Q_t                        (A from TENSOR MEMORY, B = the streamed tile, MN-major)
//   MODE_DQ  : one CTA per 128-row query block i.  Q_i, dO_i stay; 64-row KV tiles (K_t, V_t) stream.
//              S = Q_i K_t^T,  dP = dO_i V_t^T,  dS as above with per-row lse / delta,   dQ_i += dS K_t.
// The same role split as the forward: warp 9 = TMA producer, warp 8 = MMA issuer, warps 0-3 / 4-7 = element-wise math
// of even / odd streamed tiles, so the tensor pipe works on one tile set while the other one is in its exp2 phase.
// TMEM (512 columns): [set 0: T0 | T1] [set 1: T0 | T1] [ACC0 = dV] [ACC1 = dK or dQ].
// Recomputing S and dP in both kernels costs 7 GEMMs instead of the fused scheme's 5, and buys bit-reproducible
// gradients (no fp32 atomics on dQ) and two simple pipelines.
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "ew_common.cuh"
#include "tc_common.cuh"

namespace kf {
using namespace tc;

constexpr int AB_THREADS = 320;
constexpr int AB_NSTAGE = 4;
enum { MODE_DKV = 0, MODE_DQ = 1 };

struct AttnBwdTcParams {
    int64_t BH, Sq, Skv;
    const float *lse2;   // [BH, Sq] row log-sum-exp in the exp2 domain (lse * log2 e)
    const float *delta;  // [BH, Sq] rowsum(dO o O)
    void *out0, *out1;   // MODE_DKV: dV, dK      MODE_DQ: unused, dQ
    float scale_log2;    // softmax scale * log2(e)
    float scale;         // softmax scale (applied to dK / dQ in the epilogue)
    int nblk;            // 128-row blocks of the stationary operand per (b, h)
    int is_bf16;
    int hg;              // heads per scheduling group: inside a group the CTAs run weight-major (every head's heaviest block first,
                         // longest-processing-time order), at most hg heads' streamed operands live in L2 at a time
};

__device__ __forceinline__ uint32_t pack16b(float a, float b, int is_bf16) {
    if (is_bf16) {
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t *>(&h);
    }
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}
template <bool BF16>
__device__ __forceinline__ uint32_t pack16t(float2 v) {
    if (BF16) {
        __nv_bfloat162 h = __float22bfloat162_rn(v);
        return *reinterpret_cast<uint32_t *>(&h);
    }
    __half2 h = __float22half2_rn(v);
    return *reinterpret_cast<uint32_t *>(&h);
}
__device__ __forceinline__ float ex2_approx_b(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Element-wise phase of one streamed 64-column tile for one thread (= one stationary row):
//   P = exp2(S c - lse2),  dS = P o (dP - delta)   ->  16-bit P over T0 (MODE_DKV only) and dS over T1, in tensor memory.
// Instruction diet: packed fp32x2 FFMA / FADD / FMUL, one LDS.128 per column pair for the per-column (-lse2, -delta)
// (MODE_DKV; per-row registers in MODE_DQ), the causal / ragged mask only in the MASKED instantiation (diagonal tiles),
// and 16-column chunks double-buffered so the next tcgen05.ld is in flight while the current chunk is computed.
// vb: [32 column pairs][-lse2(2c), -lse2(2c+1), -delta(2c), -delta(2c+1)]
template <int MODE, bool MASKED, bool BF16>
__device__ __forceinline__ void bwd_ew_tile(const uint32_t t_addr, const uint32_t vb_smem, const float2 nl_row, const float2 nd_row,
                                            const float sc, const int lo, const int hi) {
    const float2 sc2 = make_float2(sc, sc);
    uint32_t s_r[2][16], dp_r[2][16];
    tmem_ld16(t_addr, s_r[0]);
    tmem_ld16(t_addr + 64, dp_r[0]);
    tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 4; ++c) {  // 16 streamed columns per chunk
        const int cur = c & 1;
        if (c < 3) {
            tmem_ld16(t_addr + (uint32_t)((c + 1) * 16), s_r[cur ^ 1]);
            tmem_ld16(t_addr + 64 + (uint32_t)((c + 1) * 16), dp_r[cur ^ 1]);
        }
        uint32_t pk[8], dk[8];
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
            float2 nl, nd;
            if (MODE == MODE_DKV) {
                float4 v;  // explicit ld.shared: the generic pointer would compile to LD.E
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(vb_smem + 8u * (c * 16 + i)));
                nl = make_float2(v.x, v.y);
                nd = make_float2(v.z, v.w);
            } else {
                nl = nl_row;
                nd = nd_row;
            }
            const float2 x = __ffma2_rn(make_float2(__uint_as_float(s_r[cur][i]), __uint_as_float(s_r[cur][i + 1])), sc2, nl);
            float2 pr = make_float2(ex2_approx_b(x.x), ex2_approx_b(x.y));
            if (MASKED) {
                const int col = c * 16 + i;
                if (col < lo || col >= hi) pr.x = 0.f;
                if (col + 1 < lo || col + 1 >= hi) pr.y = 0.f;
            }
            const float2 ds = __fmul2_rn(pr, __fadd2_rn(make_float2(__uint_as_float(dp_r[cur][i]), __uint_as_float(dp_r[cur][i + 1])), nd));
            if (MODE == MODE_DKV) pk[i >> 1] = pack16t<BF16>(pr);
            dk[i >> 1] = pack16t<BF16>(ds);
        }
        if (MODE == MODE_DKV) tmem_st8(t_addr + (uint32_t)(c * 8), pk);  // P^T over T0 (columns already consumed)
        tmem_st8(t_addr + 64 + (uint32_t)(c * 8), dk);                     // dS over T1
        if (c < 3) tmem_ld_wait();
    }
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// rowsum(dO o O) and the exp2-domain LSE.  16-byte loads: D/8 lanes per query row, 256/(D/8) rows per CTA (D = 64 / 128).
template <typename T>
__global__ void __launch_bounds__(256) attn_bwd_prep_kernel(const T *__restrict__ o, const T *__restrict__ dout, const float *__restrict__ lse,
                                                            float *__restrict__ delta, float *__restrict__ lse2, const int64_t rows, const int D) {
    const int lpr = D >> 3;  // lanes per row: 8 or 16
    const int64_t row = (int64_t)blockIdx.x * (256 / lpr) + threadIdx.x / lpr;
    const int sub = threadIdx.x % lpr;
    float acc = 0.f;
    if (row < rows) {
        const uint4 vo = __ldg(reinterpret_cast<const uint4 *>(o + row * D) + sub);
        const uint4 vd = __ldg(reinterpret_cast<const uint4 *>(dout + row * D) + sub);
        const T *po = reinterpret_cast<const T *>(&vo), *pd = reinterpret_cast<const T *>(&vd);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc = fmaf(cvt_in<float>(po[i]), cvt_in<float>(pd[i]), acc);
    }
    for (int off = lpr >> 1; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (row < rows && sub == 0) {
        delta[row] = acc;
        lse2[row] = lse[row] * 1.4426950408889634f;
    }
}

template <int D, int MODE>
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_x0, const __grid_constant__ CUtensorMap tmap_x1,
                   const __grid_constant__ CUtensorMap tmap_y0, const __grid_constant__ CUtensorMap tmap_y1, const AttnBwdTcParams p) {
    constexpr int ATOMS = D / 64;
    constexpr int X_BYTES = 128 * D * 2, X_ATOM = 128 * 128;  // stationary tiles: 128 rows
    constexpr int Y_BYTES = 64 * D * 2, Y_ATOM = 64 * 128;    // streamed tiles: 64 rows
    constexpr int SLOT_BYTES = 2 * Y_BYTES;
    constexpr int NS = AB_NSTAGE;
    constexpr uint32_t TMEM_COLS = 512;
    constexpr uint32_t ACC0_COL = 256, ACC1_COL = 256 + D;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char *sX = smem;                 // X0 | X1
    unsigned char *sY = smem + 2 * X_BYTES;   // NS slots of Y0 | Y1
    float *svec = reinterpret_cast<float *>(smem + 2 * X_BYTES + NS * SLOT_BYTES);  // [2 sets][2 bufs][128]: lse2[64] | delta[64]
    uint64_t *bars = reinterpret_cast<uint64_t *>(svec + 2 * 2 * 128);
    uint64_t *x_full = bars;
    uint64_t *y_full = bars + 1, *y_empty = bars + 1 + NS;
    uint64_t *t_full = bars + 1 + 2 * NS;  // [2]
    uint64_t *p_full = t_full + 2;         // [2]
    uint64_t *acc_full = p_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int bh, blk;
    {
        const int per_group = p.hg * p.nblk;
        const int g = blockIdx.x / per_group, idx = blockIdx.x - g * per_group;
        const int heads = min(p.hg, (int)p.BH - g * p.hg);  // the last group may be partial
        const int lvl = idx / heads;                         // 0 = heaviest
        bh = g * p.hg + (idx - lvl * heads);
        blk = MODE == MODE_DKV ? lvl : (p.nblk - 1 - lvl);
    }
    const int x0_row = blk * 128;
    // streamed 64-row tiles [t_lo, t_hi)
    int t_lo, t_hi;
    if (MODE == MODE_DKV) {  // query tiles at or below the diagonal of this KV block
        t_lo = x0_row / 64;
        t_hi = (int)((p.Sq + 63) / 64);
    } else {  // KV tiles up to the diagonal of this query block
        t_lo = 0;
        t_hi = (int)((min((int64_t)p.Skv, (int64_t)x0_row + 128) + 63) / 64);
    }
    const int ntile = max(t_hi - t_lo, 0);

    if (warp == 9 && lane == 0) {
        prefetch_tmap(&tmap_x0);
        prefetch_tmap(&tmap_x1);
        prefetch_tmap(&tmap_y0);
        prefetch_tmap(&tmap_y1);
        mbar_init(x_full, 1);
        for (int s = 0; s < NS; ++s) {
            mbar_init(&y_full[s], 1);
            mbar_init(&y_empty[s], 1);
        }
        for (int t = 0; t < 2; ++t) {
            mbar_init(&t_full[t], 1);
            mbar_init(&p_full[t], 4);
        }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 8) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 9) {
        // ===================================================== TMA producer
        if (lane == 0 && ntile > 0) {
            mbar_arrive_expect_tx(x_full, 2 * X_BYTES);
#pragma unroll
            for (int a = 0; a < ATOMS; ++a) {
                tma_load_3d(sX + a * X_ATOM, &tmap_x0, x_full, a * 64, x0_row, bh);
                tma_load_3d(sX + X_BYTES + a * X_ATOM, &tmap_x1, x_full, a * 64, x0_row, bh);
            }
            int s = 0;
            uint32_t ph = 0;
            for (int t = t_lo; t < t_hi; ++t) {
                mbar_wait(&y_empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&y_full[s], SLOT_BYTES);
#pragma unroll
                for (int a = 0; a < ATOMS; ++a) {
                    tma_load_3d(sY + s * SLOT_BYTES + a * Y_ATOM, &tmap_y0, &y_full[s], a * 64, t * 64, bh);
                    tma_load_3d(sY + s * SLOT_BYTES + Y_BYTES + a * Y_ATOM, &tmap_y1, &y_full[s], a * 64, t * 64, bh);
                }
                if (++s == NS) {
                    s = 0;
                    ph ^= 1;
                }
            }
        }
        __syncwarp();
    } else if (warp == 8) {
        // ===================================================== MMA issuer
        // converged warp: all lanes run the loops and waits, the elected lane issues (see umma_f16_p)
        const bool leader = elect_one();
        if (ntile > 0) {
            const int fmt = p.is_bf16 ? 1 : 0;
            const uint32_t idesc_t = make_idesc_f16(fmt, 0, 0, 128, 64);  // T = X Y^T : both operands K-major
            const uint32_t idesc_a = make_idesc_f16(fmt, 0, 1, 128, D);   // ACC += (TMEM) Y : B MN-major
            const uint32_t x_addr = smem_u32(sX), y_addr = smem_u32(sY);
            auto issue_t = [&](int set, int slot) {  // T0(set) = X0 Y0^T, T1(set) = X1 Y1^T
#pragma unroll
                for (int which = 0; which < 2; ++which) {
                    const uint32_t xa = x_addr + which * X_BYTES, ya = y_addr + slot * SLOT_BYTES + which * Y_BYTES;
#pragma unroll
                    for (int kk = 0; kk < D / 16; ++kk) {
                        umma_f16_p(tmem_base + (uint32_t)(set * 128 + which * 64),
                                   make_sw128_desc(xa + (uint32_t)((kk >> 2) * X_ATOM + (kk & 3) * 32), 0, 1024),
                                   make_sw128_desc(ya + (uint32_t)((kk >> 2) * Y_ATOM + (kk & 3) * 32), 0, 1024), idesc_t, kk ? 1u : 0u, leader);
                    }
                }
            };
            auto issue_acc = [&](int set, int slot, bool accumulate) {
                const uint32_t ya = y_addr + slot * SLOT_BYTES;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {  // K = 64 streamed rows = 4 x 16; 16-bit A in TMEM: 8 columns per step
                    if (MODE == MODE_DKV)  // dV += P^T dO_t   (B = Y1)
                        umma_f16_ts_p(tmem_base + ACC0_COL, tmem_base + (uint32_t)(set * 128 + kk * 8),
                                      make_sw128_desc(ya + Y_BYTES + kk * 2048, Y_ATOM, 1024), idesc_a, (accumulate || kk) ? 1u : 0u, leader);
                    // dK += dS^T Q_t  /  dQ += dS K_t   (B = Y0)
                    umma_f16_ts_p(tmem_base + ACC1_COL, tmem_base + (uint32_t)(set * 128 + 64 + kk * 8), make_sw128_desc(ya + kk * 2048, Y_ATOM, 1024),
                                  idesc_a, (accumulate || kk) ? 1u : 0u, leader);
                }
            };
            // ring bookkeeping: tile n (0-based) lives in slot n % NS, phase (n / NS) & 1
            auto wait_slot = [&](int n) {
                mbar_wait(&y_full[n % NS], (uint32_t)((n / NS) & 1));
                return n % NS;
            };
            mbar_wait(x_full, 0);
            for (int n = 0; n < 2 && n < ntile; ++n) {
                const int slot = wait_slot(n);
                tc_fence_after();
                issue_t(n & 1, slot);
                umma_commit_p(&t_full[n & 1], leader);
            }
            for (int n = 0; n < ntile; ++n) {
                const int set = n & 1;
                mbar_wait(&p_full[set], (uint32_t)((n >> 1) & 1));
                tc_fence_after();
                issue_acc(set, n % NS, n > 0);
                umma_commit_p(&y_empty[n % NS], leader);  // the streamed tile n is no longer needed once these MMAs retire
                if (n + 2 < ntile) {
                    const int slot = wait_slot(n + 2);
                    tc_fence_after();
                    issue_t(set, slot);
                    umma_commit_p(&t_full[set], leader);
                }
            }
            umma_commit_p(acc_full, leader);
        }
        __syncwarp();
    } else {
        // ===================================================== element-wise math + epilogue: thread = stationary row
        const int set = warp >> 2, q4 = warp & 3;
        const int r = q4 * 32 + lane;
        const int64_t row_g = (int64_t)x0_row + r;  // MODE_DKV: kv index, MODE_DQ: query index
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q4 * 32) << 16);
        const uint32_t t_addr = lane_addr + (uint32_t)(set * 128);
        const float sc = p.scale_log2;
        const float *lse2 = p.lse2 + (int64_t)bh * p.Sq, *delta = p.delta + (int64_t)bh * p.Sq;
        float my_lse = 0.f, my_delta = 0.f;  // MODE_DQ: per-row scalars
        if (MODE == MODE_DQ && row_g < p.Sq) {
            my_lse = lse2[row_g];
            my_delta = delta[row_g];
        }
        float *vec = svec + set * 256;  // [2 bufs][32 column pairs][-lse2 x2, -delta x2]
        const int tsel = threadIdx.x & 127;  // thread index inside the set
        // MODE_DKV: threads 0-63 fetch -lse2 of streamed row tsel, 64-127 fetch -delta; slot inside the pair-interleaved vector
        const int vslot = 4 * ((tsel & 63) >> 1) + (tsel & 1) + (tsel < 64 ? 0 : 2);
        float vnext = 0.f;
        auto load_vec = [&](int n) {  // n: tile ordinal
            float v = 0.f;
            if (MODE == MODE_DKV && n < ntile) {
                const int64_t qg = (int64_t)(t_lo + n) * 64 + (tsel & 63);
                if (qg < p.Sq) v = -(tsel < 64 ? lse2 : delta)[qg];
            }
            return v;
        };
        vnext = load_vec(set);
        const float2 nl_row = make_float2(-my_lse, -my_lse), nd_row = make_float2(-my_delta, -my_delta);
        for (int n = set; n < ntile; n += 2) {
            const int t = t_lo + n;
            const int64_t y0_row = (int64_t)t * 64;
            const int it = n >> 1;
            float *vb = vec + (it & 1) * 128;
            if (MODE == MODE_DKV) {
                vb[vslot] = vnext;
                named_bar_sync(1 + set, 128);
                vnext = load_vec(n + 2);
            }
            // columns c (streamed index) kept iff lo <= c < hi
            int lo = 0, hi = 64;
            if (MODE == MODE_DKV) {  // keep q >= kv, q < Sq
                lo = (int)max((int64_t)0, min((int64_t)64, row_g - y0_row));
                hi = (int)max((int64_t)0, min((int64_t)64, p.Sq - y0_row));
            } else {  // keep kv <= q, kv < Skv
                hi = (int)max((int64_t)0, min((int64_t)64, min(row_g + 1, (int64_t)p.Skv) - y0_row));
            }
            const bool need_mask = __any_sync(0xffffffffu, lo > 0 || hi < 64);
            mbar_wait(&t_full[set], (uint32_t)(it & 1));
            tc_fence_after();
            const uint32_t vbs = smem_u32(vb);
            if (p.is_bf16) {
                if (need_mask) bwd_ew_tile<MODE, true, true>(t_addr, vbs, nl_row, nd_row, sc, lo, hi);
                else bwd_ew_tile<MODE, false, true>(t_addr, vbs, nl_row, nd_row, sc, lo, hi);
            } else {
                if (need_mask) bwd_ew_tile<MODE, true, false>(t_addr, vbs, nl_row, nd_row, sc, lo, hi);
                else bwd_ew_tile<MODE, false, false>(t_addr, vbs, nl_row, nd_row, sc, lo, hi);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[set]);
        }
        // ---- epilogue
        const bool is_bf16 = p.is_bf16;
        auto store_acc = [&](uint32_t col0, int ncols, void *outp, int64_t nrows_valid, float mul, int out_col0) {
            const bool row_ok = row_g < nrows_valid;
            uint16_t *orow = reinterpret_cast<uint16_t *>(outp) + ((int64_t)bh * nrows_valid + (row_ok ? row_g : 0)) * D + out_col0;
#pragma unroll 1
            for (int c = 0; c < ncols; c += 32) {
                uint32_t a[32];
                if (ntile > 0) {
                    tmem_ld32(lane_addr + col0 + (uint32_t)c, a);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) a[i] = 0u;  // nothing attends to this block: exact zeros
                }
                if (row_ok) {
                    uint4 *dst = reinterpret_cast<uint4 *>(orow + c);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint32_t wv[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            wv[k] = pack16b(__uint_as_float(a[8 * i + 2 * k]) * mul, __uint_as_float(a[8 * i + 2 * k + 1]) * mul, is_bf16);
                        dst[i] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
                    }
                }
            }
        };
        if (ntile > 0) {
            mbar_wait(acc_full, 0);
            tc_fence_after();
        }
        if (MODE == MODE_DKV) {
            if (set == 0) store_acc(ACC0_COL, D, p.out0, p.Skv, 1.f, 0);      // dV
            else store_acc(ACC1_COL, D, p.out1, p.Skv, p.scale, 0);           // dK
        } else {
            store_acc(ACC1_COL + (uint32_t)(set * (D / 2)), D / 2, p.out1, p.Sq, p.scale, set * (D / 2));  // dQ, half the columns per set
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

template <int D, int MODE>
static void launch_bwd_mode(const AttnBwdPlan &a, const float *lse2, const float *delta) {
    Runtime &rt = Runtime::get();
    const bool bf16 = a.dtype == KF_BFLOAT16;
    auto map = [&](const void *ptr, int64_t S, uint32_t rows) {
        return make_tmap_3d_16bit(ptr, bf16, D, (uint64_t)S, (uint64_t)a.BH, D, (uint64_t)S * D, 64, rows);
    };
    // stationary (128-row boxes) / streamed (64-row boxes) operands
    const CUtensorMap x0 = MODE == MODE_DKV ? map(a.k, a.Skv, 128) : map(a.q, a.Sq, 128);
    const CUtensorMap x1 = MODE == MODE_DKV ? map(a.v, a.Skv, 128) : map(a.dout, a.Sq, 128);
    const CUtensorMap y0 = MODE == MODE_DKV ? map(a.q, a.Sq, 64) : map(a.k, a.Skv, 64);
    const CUtensorMap y1 = MODE == MODE_DKV ? map(a.dout, a.Sq, 64) : map(a.v, a.Skv, 64);
    AttnBwdTcParams p{};
    p.BH = a.BH; p.Sq = a.Sq; p.Skv = a.Skv;
    p.lse2 = lse2; p.delta = delta;
    p.out0 = a.dv;
    p.out1 = MODE == MODE_DKV ? a.dk : a.dq;
    const double scale = 1.0 / std::sqrt((double)D);
    p.scale = (float)scale;
    p.scale_log2 = (float)(scale * 1.4426950408889634);
    p.nblk = (int)(((MODE == MODE_DKV ? a.Skv : a.Sq) + 127) / 128);
    p.is_bf16 = bf16;
    {  // heads per scheduling group.  Measured at C3 and at S = 1024 (tools/gpu_attn_bwd_ab.py): unlike the forward, the backward
       // is no faster in weight-major order (groups that fit the L2, or one global list) than head-major, so 1 stays the default
        int64_t hg = 1;
        if (const char *e = std::getenv("KF_ATTN_HG")) hg = std::max(1, std::atoi(e));
        p.hg = (int)std::min<int64_t>(hg, a.BH);
    }
    constexpr int SMEM = 2 * 128 * D * 2 + AB_NSTAGE * 2 * 64 * D * 2 + 2 * 2 * 128 * 4 + 256 + 1024;
    auto kern = attn_bwd_tc_kernel<D, MODE>;
    static bool attr_done = false;
    if (!attr_done) {
        KF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        attr_done = true;
    }
    const int64_t grid = a.BH * p.nblk;
    KF_CHECK(grid < (int64_t)0x7FFFFFFF);
    kern<<<(unsigned)grid, AB_THREADS, SMEM, rt.stream()>>>(x0, x1, y0, y1, p);
    rt.post_launch(MODE == MODE_DKV ? "attn_bwd_dkv_tc_kernel" : "attn_bwd_dq_tc_kernel");
}

bool launch_attention_bwd_tc(const AttnBwdPlan &a) {
    static const bool force_generic = std::getenv("KF_ATTN_BWD_FORCE_GENERIC") != nullptr;
    if (force_generic) return false;
    // Three schemes, all deterministic (profiles/r2_attn_bwd_variants.md has the C3 timings and the pipeline traces):
    //   two   (this file)                 64-wide tiles, two tile sets, 7 GEMMs; dense operands only.      3.39 ms at C3 — the default for dense
    //   wide  (attention_bwd_wide.cu)     128-wide tiles, 16 element-wise warps, dense or strided.         3.42 - 3.7 ms — the strided (packed qkv) path
    //   fused (attention_bwd_fused.cu)    one kernel, 5 GEMMs, ordered fp32 dQ hand-over through L2.       5.16 ms
    // KF_ATTN_BWD=two|wide|fused forces one (read per call so that tests / A-B runs can flip it).
    const char *mode = std::getenv("KF_ATTN_BWD");
    const bool want_fused = mode && std::strcmp(mode, "fused") == 0, want_wide = mode && std::strcmp(mode, "wide") == 0;
    if (want_fused && launch_attention_bwd_fused(a)) return true;
    if ((want_wide || a.H > 0) && launch_attention_bwd_wide(a)) return true;
    if (a.H > 0) return launch_attention_bwd_fused(a);  // strided operands the wide scheme refused (cannot happen for shapes the forward took)
    if (a.dtype != KF_HALF && a.dtype != KF_BFLOAT16) return false;
    if (a.D != 64 && a.D != 128) return false;
    if (a.Sq < 1 || a.Skv < 1 || a.BH < 1 || a.BH >= 65536) return false;
    auto al = [](const void *p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
    if (!al(a.q) || !al(a.k) || !al(a.v) || !al(a.out) || !al(a.dout) || !al(a.dq) || !al(a.dk) || !al(a.dv)) return false;
    Runtime &rt = Runtime::get();
    const int64_t rows = a.BH * a.Sq;
    Scratch delta((size_t)rows * 4), lse2((size_t)rows * 4);
    const int64_t rows_per_cta = 256 / (a.D / 8);
    KF_CHECK((rows + rows_per_cta - 1) / rows_per_cta < (int64_t)0x7FFFFFFF);
    const unsigned pgrid = (unsigned)((rows + rows_per_cta - 1) / rows_per_cta);
    if (a.dtype == KF_BFLOAT16)
        attn_bwd_prep_kernel<__nv_bfloat16><<<pgrid, 256, 0, rt.stream()>>>((const __nv_bfloat16 *)a.out, (const __nv_bfloat16 *)a.dout, (const float *)a.lse,
                                                                            delta.as<float>(), lse2.as<float>(), rows, (int)a.D);
    else
        attn_bwd_prep_kernel<__half><<<pgrid, 256, 0, rt.stream()>>>((const __half *)a.out, (const __half *)a.dout, (const float *)a.lse, delta.as<float>(),
                                                                     lse2.as<float>(), rows, (int)a.D);
    rt.post_launch("attn_bwd_prep_kernel");
    if (a.D == 64) {
        launch_bwd_mode<64, MODE_DKV>(a, lse2.as<float>(), delta.as<float>());
        launch_bwd_mode<64, MODE_DQ>(a, lse2.as<float>(), delta.as<float>());
    } else {
        launch_bwd_mode<128, MODE_DKV>(a, lse2.as<float>(), delta.as<float>());
        launch_bwd_mode<128, MODE_DQ>(a, lse2.as<float>(), delta.as<float>());
    }
    return true;
}

}  // namespace kf

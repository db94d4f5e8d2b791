// fp32 GEMM on the 5th-generation tensor cores by split-precision emulation ("bf16 x 3 planes, 6 or 9 products"), hand-written
// tcgen05 / TMEM / TMA PTX.  The reference's only GEMM is fp32 / fp64 CUTLASS 2.x SIMT FFMA (src/device/gemm_kernel.cu:8-38,
// src/device/launcher_cuda.h:537-614): this is the kernel that has to beat it on the same box ("What's missing" #1).
//
//   every fp32 operand element is written as x = x0 + x1 + x2 with x0 = bf16(x), x1 = bf16(x - x0), x2 = bf16(x - x0 - x1)
//   (both subtractions are exact in fp32; |x1| <= 2^-9 |x|, |x2| <= 2^-18 |x|, remainder <= 2^-27 |x|), so
//   a b = sum_{i,j} a_i b_j.  With the six products of weight >= 2^-18 (i + j <= 2) the dropped terms are <= 2^-26 |a b|,
//   below the fp32 rounding of the product itself; the products run on tcgen05 kind::f16 with fp32 accumulation in TMEM.
//   KF_GEMM_F32=x9 keeps all nine products, KF_GEMM_F32=simt forces the FFMA kernel.
//   Cost: 6 bf16 MMAs per fp32 MMA = 1 / 6 of the bf16 tensor rate (~ 250 TFLOP/s of fp32 work) against ~ 70 TFLOP/s of FFMA peak.
//
// Two launches:
//   split_f32_kernel   : fp32 operand -> three bf16 planes [3][batch][rows][ld8]   (HBM-bound: 4 B read + 6 B written per element)
//   gemm_f32x_kernel   : CTA pair (cta_group::2), 256 x 128 output tile, persistent, 192 threads per CTA
//       warp 0   TMA producer: per 64-wide k block ONE "super stage" = this CTA's A planes (3 x 128 x 64) + B planes (3 x 64 x 64),
//                72 KB, three stages deep; bytes of both CTAs are credited to the leader's `full[s]`
//       warp 1   MMA issuer (leader CTA): per stage 4 k-slices x NPROD plane pairs of tcgen05.mma M256 N128 K16 into TMEM:
//                a0 b0 into the "main" accumulator, every other pair into the "small" one (see F_KCHUNK below)
//       warps 2-5 epilogue: tcgen05.ld main + small -> IEEE add -> (+ running partial of earlier K chunks) -> alpha / beta -> fp32
//                16-byte stores, overlapped with the next chunk's MMAs (2 x 256 TMEM columns)
//   Each operand tile is fetched once per k block and used by up to three products, so L2 -> SMEM traffic per FLOP is half of
//   what a "K' = 6 K" formulation on the plain bf16 kernel would move.
// Non-finite inputs: inf * b becomes inf * b0 + inf * b1 + ... = NaN when b1 and b0 differ in sign (documented deviation; the
// SIMT kernel keeps IEEE behaviour and is selected with KF_GEMM_F32=simt).
#include <cstdlib>
#include <cstring>

#include "ew_common.cuh"
#include "tc_common.cuh"

namespace kf {
using namespace tc;

namespace {

constexpr int F_BM = 128;   // rows of A per CTA (the pair covers 256)
constexpr int F_BN = 128;   // output tile width; each CTA stages F_BN / 2 columns of B
constexpr int F_BK = 64;    // one 128-byte swizzle atom of bf16
constexpr int F_THREADS = 192;
constexpr int F_NSTAGES = 3;
constexpr int F_PLANE_A = F_BM * F_BK * 2;         // 16 KB
constexpr int F_PLANE_B = (F_BN / 2) * F_BK * 2;   // 8 KB
constexpr int F_STAGE_BYTES = 3 * (F_PLANE_A + F_PLANE_B);  // 72 KB
constexpr int F_SMEM_TOTAL = F_NSTAGES * F_STAGE_BYTES + 256 + 1024;
// The tensor core adds every MMA's result to the fp32 accumulator with TRUNCATION (measured on B200, profiles/r2_f32_gemm_accuracy.md:
// all-positive operands, K = 8192, one accumulator: 4.9e-5 relative = ~ half an ulp lost per MMA).  Two measures keep the result
// inside the 1e-5 band whatever K is:
//   * the leading product a0 b0 has its own accumulator; the five (eight) small products share a second one, 2^-8 smaller, so
//     their 5x more frequent truncations cost nothing measurable;
//   * K is cut into chunks of F_KCHUNK: each chunk's two accumulators are summed by the epilogue warps in IEEE fp32 and added to
//     the running partial kept in C itself (read-modify-write by the same thread, L2-resident), so no accumulator ever sees
//     more than F_KCHUNK / 16 = 128 truncating additions (<= 2.8e-6 relative even when every term has the same sign).
constexpr int F_KCHUNK = 2048;
constexpr int F_ACC_COLS = 2 * F_BN;  // main | small

struct GemmF32Params {
    int64_t M, N, K, batch;
    int64_t ldc, sc;
    float *c;
    float alpha, beta;
    int m_tiles, n_tiles;
    int64_t total_tiles;
    int nchunks;            // K chunks per tile
    const float *residual;  // out = alpha * acc + beta * C + residual
    int64_t ldr, sr;
    int a_nb, b_nb;      // batch entries in the plane buffers (1 when the operand is broadcast over the batch)
    int a_bmul, b_bmul;  // 0 when broadcast
};

// ---------------------------------------------------------------------------------------------- operand split
// x [batch][rows][cols] (leading dimension ld, batch stride bs) -> planes [3][batch][rows][ldp] bf16
__global__ void __launch_bounds__(256) split_f32_kernel(const float *__restrict__ x, __nv_bfloat16 *__restrict__ planes, const int64_t rows,
                                                        const int64_t cols, const int64_t ld, const int64_t bs, const int64_t ldp,
                                                        const int64_t plane_stride, const int64_t total_quads, const int quads_per_row, const bool vec_ok) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total_quads; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row_all = t / quads_per_row;  // batch * rows + row
        const int64_t c0 = (t % quads_per_row) * 4;
        const int64_t b = row_all / rows, r = row_all % rows;
        const float *src = x + b * bs + r * ld + c0;
        float v[4];
        if (vec_ok && c0 + 4 <= cols) {
            const float4 f = __ldg(reinterpret_cast<const float4 *>(src));
            v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = (c0 + i < cols) ? __ldg(src + i) : 0.f;
        }
        uint16_t h[3][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float rem = v[i];
            // the leading plane is truncated (never overflows to inf for finite x), the two corrections are rounded to nearest
            const float p0 = __uint_as_float(__float_as_uint(rem) & 0xffff0000u);
            __nv_bfloat16 b0 = __float2bfloat16_rn(rem);
            float f0 = __bfloat162float(b0);
            if (isinf(f0) && !isinf(rem)) {
                f0 = p0;
                b0 = __float2bfloat16_rn(p0);
            }
            rem = __fsub_rn(rem, f0);
            const __nv_bfloat16 b1 = __float2bfloat16_rn(rem);
            rem = __fsub_rn(rem, __bfloat162float(b1));
            const __nv_bfloat16 b2 = __float2bfloat16_rn(rem);
            h[0][i] = __bfloat16_as_ushort(b0);
            h[1][i] = __bfloat16_as_ushort(b1);
            h[2][i] = __bfloat16_as_ushort(b2);
        }
        __nv_bfloat16 *dst = planes + row_all * ldp + c0;  // ldp % 8 == 0 and c0 % 4 == 0: 8-byte aligned
#pragma unroll
        for (int pl = 0; pl < 3; ++pl) {
            uint2 w;
            w.x = (uint32_t)h[pl][0] | ((uint32_t)h[pl][1] << 16);
            w.y = (uint32_t)h[pl][2] | ((uint32_t)h[pl][3] << 16);
            *reinterpret_cast<uint2 *>(dst + pl * plane_stride) = w;
        }
    }
}

struct SplitOperand {
    void *planes;
    int64_t ldp, nb;
};

static SplitOperand split_operand(const float *x, int64_t rows, int64_t cols, int64_t ld, int64_t bs, int64_t nb, Scratch &holder) {
    Runtime &rt = Runtime::get();
    const int64_t ldp = (cols + 7) / 8 * 8;
    const int64_t plane_stride = nb * rows * ldp;
    (void)holder;
    const int quads_per_row = (int)(ldp / 4);
    const int64_t total = nb * rows * quads_per_row;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(x) % 16 == 0) && (ld % 4 == 0) && (bs % 4 == 0);
    split_f32_kernel<<<grid_for(total, 256, 8), 256, 0, rt.stream()>>>(x, holder.as<__nv_bfloat16>(), rows, cols, ld, bs, ldp, plane_stride, total,
                                                                         quads_per_row, vec_ok);
    rt.post_launch("split_f32_kernel");
    return {holder.p, ldp, nb};
}

// ---------------------------------------------------------------------------------------------- main kernel
template <int BAND>
__device__ __forceinline__ void f32_tile_coords(int64_t t, const GemmF32Params &p, int &b, int &mb, int &nb) {
    const int64_t per_batch = (int64_t)p.m_tiles * p.n_tiles;
    b = (int)(t / per_batch);
    int64_t r = t % per_batch;
    const int64_t band_sz = (int64_t)BAND * p.n_tiles;
    const int band = (int)(r / band_sz);
    r -= (int64_t)band * band_sz;
    const int band_rows = min(BAND, p.m_tiles - band * BAND);
    nb = (int)(r / band_rows);
    mb = band * BAND + (int)(r % band_rows);
}

// one epilogue thread = one accumulator row: TMEM (main + small, 32 fp32 columns at a time) -> IEEE add -> running partial in C
// (chunks before the last) or alpha / beta / residual -> final store (last chunk)
__device__ __forceinline__ void f32_epilogue_tile(const GemmF32Params &p, uint32_t taddr, int64_t row, int64_t n0, float *crow, const float *rrow,
                                                  bool vec_ok, int chunk) {
    const bool first = chunk == 0, last = chunk == p.nchunks - 1;
    const float beta_over_alpha = (p.beta != 0.f && p.nchunks > 1) ? p.beta / p.alpha : 0.f;
#pragma unroll 1
    for (int c = 0; c < F_BN; c += 32) {
        uint32_t rm[32], rs[32];
        tmem_ld32(taddr + (uint32_t)c, rm);
        tmem_ld32(taddr + (uint32_t)(F_BN + c), rs);
        tmem_ld_wait();
        if (row < p.M && n0 + c < p.N) {
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(rm[i]) + __uint_as_float(rs[i]);
            const bool full_chunk = n0 + c + 32 <= p.N;
            const bool need_c = !first || p.beta != 0.f;
            if (need_c) {
                float old[32];
                if (full_chunk && vec_ok) {
                    const float4 *src = reinterpret_cast<const float4 *>(crow + n0 + c);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 f = src[i];
                        old[4 * i] = f.x; old[4 * i + 1] = f.y; old[4 * i + 2] = f.z; old[4 * i + 3] = f.w;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) old[i] = (n0 + c + i < p.N) ? crow[n0 + c + i] : 0.f;
                }
                if (p.nchunks == 1) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = fmaf(p.beta, old[i], v[i] * p.alpha);
                } else if (first) {  // fold beta * C_old into the running (unscaled) partial: alpha * (acc + beta / alpha * C_old)
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = fmaf(beta_over_alpha, old[i], v[i]);
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] += old[i];
                }
            } else if (p.nchunks == 1) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] *= p.alpha;
            }
            if (last) {
                if (p.nchunks > 1) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] *= p.alpha;
                }
                if (rrow != nullptr) {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (n0 + c + i < p.N) v[i] += __ldg(rrow + n0 + c + i);
                }
            }
            if (full_chunk && vec_ok) {
                float4 *dst = reinterpret_cast<float4 *>(crow + n0 + c);
#pragma unroll
                for (int i = 0; i < 8; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (n0 + c + i < p.N) crow[n0 + c + i] = v[i];
            }
        }
    }
}

// plane pairs (i of A, j of B) in order of increasing weight: (2,2) (1,2) (2,1) (0,2) (2,0) (1,1) (0,1) (1,0) (0,0);
// the six-product variant takes the last six
__host__ __device__ constexpr int pair_a(int q) { return q == 0 ? 2 : q == 1 ? 1 : q == 2 ? 2 : q == 3 ? 0 : q == 4 ? 2 : q == 5 ? 1 : q == 6 ? 0 : q == 7 ? 1 : 0; }
__host__ __device__ constexpr int pair_b(int q) { return q == 0 ? 2 : q == 1 ? 2 : q == 2 ? 1 : q == 3 ? 2 : q == 4 ? 0 : q == 5 ? 1 : q == 6 ? 1 : q == 7 ? 0 : 0; }

template <bool A_MN, bool B_MN, int NPROD>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(F_THREADS, 1)
gemm_f32x_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const GemmF32Params p) {
    constexpr int NST = F_NSTAGES;
    constexpr int HALF_N = F_BN / 2;
    constexpr int KCB = F_KCHUNK / F_BK;  // k blocks per chunk
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + NST * F_STAGE_BYTES);
    uint64_t *empty = full + NST;
    uint64_t *tmem_full = empty + NST;
    uint64_t *tmem_empty = tmem_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int64_t cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int nkb = (int)((p.K + F_BK - 1) / F_BK);

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_a);
        prefetch_tmap(&tmap_b);
        for (int s = 0; s < NST; ++s) {
            mbar_init(&full[s], 1);   // leader only: its producer's arrive.expect_tx (bytes of both CTAs)
            mbar_init(&empty[s], 1);  // multicast commit from the leader's MMA warp
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 8);  // leader only: 4 epilogue warps x 2 CTAs
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2sm(tmem_slot, 2 * F_ACC_COLS);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================================================== TMA producer (both CTAs)
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            const uint32_t full0 = mapa_u32(&full[0], 0);
            for (int64_t t = cluster_id; t < p.total_tiles; t += n_clusters) {
                int b, mb, nb;
                f32_tile_coords<8>(t, p, b, mb, nb);
                const int m0 = mb * (2 * F_BM) + (int)rank * F_BM, n0 = nb * F_BN + (int)rank * HALF_N;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&empty[s], ph ^ 1);
                    unsigned char *st = smem + s * F_STAGE_BYTES;
                    if (rank == 0) mbar_arrive_expect_tx(&full[s], 2 * F_STAGE_BYTES);
                    const uint32_t fb = full0 + (uint32_t)s * 8u;
                    const int k0 = kb * F_BK;
#pragma unroll
                    for (int pl = 0; pl < 3; ++pl) {
                        unsigned char *sa = st + pl * F_PLANE_A;
                        unsigned char *sb = st + 3 * F_PLANE_A + pl * F_PLANE_B;
                        const int ca = pl * p.a_nb + b * p.a_bmul, cb = pl * p.b_nb + b * p.b_bmul;
                        if constexpr (!A_MN) {
                            tma_load_3d_2sm(sa, &tmap_a, fb, k0, m0, ca);  // box 64(k) x 128(m)
                        } else {
#pragma unroll
                            for (int i = 0; i < F_BM / 64; ++i) tma_load_3d_2sm(sa + i * (64 * F_BK * 2), &tmap_a, fb, m0 + i * 64, k0, ca);
                        }
                        if constexpr (!B_MN) {
                            tma_load_3d_2sm(sb, &tmap_b, fb, k0, n0, cb);  // box 64(k) x 64(n)
                        } else {
#pragma unroll
                            for (int i = 0; i < HALF_N / 64; ++i) tma_load_3d_2sm(sb + i * (64 * F_BK * 2), &tmap_b, fb, n0 + i * 64, k0, cb);
                        }
                    }
                    if (++s == NST) {
                        s = 0;
                        ph ^= 1;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================================================== MMA issuer (leader CTA only; converged warp, elected lane issues)
        if (rank == 0) {
            const bool leader = elect_one();
            const uint32_t idesc = make_idesc_f16(1, A_MN ? 1 : 0, B_MN ? 1 : 0, 2 * F_BM, F_BN);
            int s = 0;
            uint32_t ph = 0;
            int acc = 0;
            uint32_t acc_ph = 0;
            for (int64_t t = cluster_id; t < p.total_tiles; t += n_clusters) {
                for (int kb = 0; kb < nkb; ++kb) {
                    const int kin = kb % KCB;  // position inside the K chunk: every chunk starts a fresh pair of accumulators
                    if (kin == 0) {
                        mbar_wait_cluster(&tmem_empty[acc], acc_ph ^ 1);
                        tc_fence_after();
                    }
                    const uint32_t d_main = tmem_base + (uint32_t)(acc * F_ACC_COLS), d_small = d_main + (uint32_t)F_BN;
                    mbar_wait_cluster(&full[s], ph);
                    tc_fence_after();
                    const uint32_t sa0 = smem_u32(smem + s * F_STAGE_BYTES);
                    const uint32_t sb0 = sa0 + 3 * F_PLANE_A;
#pragma unroll
                    for (int k = 0; k < F_BK / 16; ++k) {
#pragma unroll
                        for (int q = 9 - NPROD; q < 9; ++q) {
                            const uint32_t sa = sa0 + (uint32_t)(pair_a(q) * F_PLANE_A), sb = sb0 + (uint32_t)(pair_b(q) * F_PLANE_B);
                            const uint64_t adesc = A_MN ? make_sw128_desc(sa + k * 2048, 64 * F_BK * 2, 1024) : make_sw128_desc(sa + k * 32, 0, 1024);
                            const uint64_t bdesc = B_MN ? make_sw128_desc(sb + k * 2048, 64 * F_BK * 2, 1024) : make_sw128_desc(sb + k * 32, 0, 1024);
                            if (q == 8)  // a0 b0: the leading product, alone in its accumulator
                                umma_f16_2sm_p(d_main, adesc, bdesc, idesc, (kin | k) ? 1u : 0u, leader);
                            else
                                umma_f16_2sm_p(d_small, adesc, bdesc, idesc, (kin | k | (q - (9 - NPROD))) ? 1u : 0u, leader);
                        }
                    }
                    umma_commit_2sm_mc_p(&empty[s], leader);
                    if (kin == KCB - 1 || kb == nkb - 1) {
                        umma_commit_2sm_mc_p(&tmem_full[acc], leader);
                        if (++acc == 2) {
                            acc = 0;
                            acc_ph ^= 1;
                        }
                    }
                    if (++s == NST) {
                        s = 0;
                        ph ^= 1;
                    }
                }
            }
        }
        __syncwarp();
    } else {
        // ===================================================== epilogue (both CTAs; warps 2..5 -> TMEM lane quarters 2,3,0,1)
        const int q = warp & 3;
        int acc = 0;
        uint32_t acc_ph = 0;
        const bool vec_ok = (p.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.c) & 15) == 0) && (p.sc % 4 == 0);
        const uint32_t tmem_empty0 = mapa_u32(&tmem_empty[0], 0);
        for (int64_t t = cluster_id; t < p.total_tiles; t += n_clusters) {
            int b, mb, nb;
            f32_tile_coords<8>(t, p, b, mb, nb);
            const int64_t row = (int64_t)mb * (2 * F_BM) + (int64_t)rank * F_BM + q * 32 + lane;
            const int64_t n0 = (int64_t)nb * F_BN;
            float *crow = p.c + (int64_t)b * p.sc + row * p.ldc;
            const float *rrow = p.residual ? p.residual + (int64_t)b * p.sr + row * p.ldr : nullptr;
            for (int chunk = 0; chunk < p.nchunks; ++chunk) {
                mbar_wait_cluster(&tmem_full[acc], acc_ph);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * F_ACC_COLS);
                f32_epilogue_tile(p, taddr, row, n0, crow, rrow, vec_ok, chunk);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(tmem_empty0 + (uint32_t)acc * 8u);
                if (++acc == 2) {
                    acc = 0;
                    acc_ph ^= 1;
                }
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, 2 * F_ACC_COLS);
    }
}

template <bool A_MN, bool B_MN, int NPROD>
static void launch_f32x_cfg(const GemmPlan &g) {
    Runtime &rt = Runtime::get();
    const int64_t a_nb = (g.batch > 1 && g.sa == 0) ? 1 : g.batch, b_nb = (g.batch > 1 && g.sb == 0) ? 1 : g.batch;
    // storage shapes: A is [M,K] (or [K,M] when transposed), B is [K,N] (or [N,K])
    const int64_t a_rows = g.trans_a ? g.K : g.M, a_cols = g.trans_a ? g.M : g.K;
    const int64_t b_rows = g.trans_b ? g.N : g.K, b_cols = g.trans_b ? g.K : g.N;
    const int64_t a_ldp = (a_cols + 7) / 8 * 8, b_ldp = (b_cols + 7) / 8 * 8;
    Scratch a_planes((size_t)(3 * a_nb * a_rows * a_ldp) * 2), b_planes((size_t)(3 * b_nb * b_rows * b_ldp) * 2);
    split_operand(reinterpret_cast<const float *>(g.a), a_rows, a_cols, g.lda, g.sa, a_nb, a_planes);
    split_operand(reinterpret_cast<const float *>(g.b), b_rows, b_cols, g.ldb, g.sb, b_nb, b_planes);
    // plane p of batch entry b is entry (p * nb + b) of the third tensor-map dimension
    const CUtensorMap ta = A_MN ? make_tmap_3d_16bit(a_planes.p, true, (uint64_t)g.M, (uint64_t)g.K, (uint64_t)(3 * a_nb), (uint64_t)a_ldp, (uint64_t)(a_rows * a_ldp), 64, 64)
                                : make_tmap_3d_16bit(a_planes.p, true, (uint64_t)g.K, (uint64_t)g.M, (uint64_t)(3 * a_nb), (uint64_t)a_ldp, (uint64_t)(a_rows * a_ldp), 64, F_BM);
    const CUtensorMap tb = B_MN ? make_tmap_3d_16bit(b_planes.p, true, (uint64_t)g.N, (uint64_t)g.K, (uint64_t)(3 * b_nb), (uint64_t)b_ldp, (uint64_t)(b_rows * b_ldp), 64, 64)
                                : make_tmap_3d_16bit(b_planes.p, true, (uint64_t)g.K, (uint64_t)g.N, (uint64_t)(3 * b_nb), (uint64_t)b_ldp, (uint64_t)(b_rows * b_ldp), 64, F_BN / 2);
    GemmF32Params p{};
    p.M = g.M; p.N = g.N; p.K = g.K; p.batch = g.batch;
    p.ldc = g.ldc; p.sc = g.sc; p.c = reinterpret_cast<float *>(g.c);
    p.alpha = g.alpha; p.beta = g.beta;
    p.residual = reinterpret_cast<const float *>(g.residual); p.ldr = g.ldr; p.sr = g.sr;
    p.m_tiles = (int)((g.M + 2 * F_BM - 1) / (2 * F_BM));
    p.n_tiles = (int)((g.N + F_BN - 1) / F_BN);
    p.total_tiles = (int64_t)p.m_tiles * p.n_tiles * g.batch;
    p.nchunks = (int)((g.K + F_KCHUNK - 1) / F_KCHUNK);
    p.a_nb = (int)a_nb; p.b_nb = (int)b_nb;
    p.a_bmul = (a_nb > 1 || g.batch == 1) ? 1 : 0;
    p.b_bmul = (b_nb > 1 || g.batch == 1) ? 1 : 0;
    auto kernel = gemm_f32x_kernel<A_MN, B_MN, NPROD>;
    static bool attr_done = false;
    if (!attr_done) {
        KF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM_TOTAL));
        attr_done = true;
    }
    const int64_t clusters = std::min<int64_t>(p.total_tiles, rt.props().sm_count / 2);
    kernel<<<(unsigned)(2 * clusters), F_THREADS, F_SMEM_TOTAL, rt.stream()>>>(ta, tb, p);
    rt.post_launch("gemm_f32x_kernel");
}

}  // namespace

// fp32 GEMM on tensor cores; false => the caller uses the SIMT kernel (tiny problems, KF_GEMM_F32=simt, huge batch counts)
bool launch_gemm_f32_tc(const GemmPlan &g) {
    if (g.dtype != KF_FLOAT) return false;
    const char *mode = std::getenv("KF_GEMM_F32");  // read per call so that tests can flip it
    if (mode && std::strcmp(mode, "simt") == 0) return false;
    if (g.M <= 0 || g.N <= 0 || g.K <= 0) return false;
    if (g.M >= (1ll << 31) || g.N >= (1ll << 31) || g.K >= (1ll << 31) || 3 * g.batch >= 65536) return false;
    // below ~ one 128 x 128 x 256 block of work the two extra launches and the 256 x 256 tile granularity lose to the FFMA kernel
    if (g.beta != 0.f && g.alpha == 0.f && g.K > F_KCHUNK) return false;  // beta / alpha is folded into the first K chunk
    const bool forced = mode && (std::strcmp(mode, "x6") == 0 || std::strcmp(mode, "x9") == 0);
    const int64_t Mr = g.route_M > 0 ? g.route_M : g.M;
    if (!forced && (Mr < 64 || g.N < 64 || g.K < 32 || (double)Mr * (double)g.N * (double)g.K * (double)g.batch < 128.0 * 128.0 * 256.0)) return false;
    const bool nine = mode && std::strcmp(mode, "x9") == 0;
#define KF_F32_DISPATCH(NP)                                                        \
    do {                                                                           \
        if (g.trans_a && !g.trans_b) launch_f32x_cfg<true, true, NP>(g);           \
        else if (g.trans_a && g.trans_b) launch_f32x_cfg<true, false, NP>(g);      \
        else if (!g.trans_a && !g.trans_b) launch_f32x_cfg<false, true, NP>(g);    \
        else launch_f32x_cfg<false, false, NP>(g);                                 \
    } while (0)
    if (nine) KF_F32_DISPATCH(9);
    else KF_F32_DISPATCH(6);
#undef KF_F32_DISPATCH
    return true;
}

}  // namespace kf

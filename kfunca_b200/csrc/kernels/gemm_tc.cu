// bf16 / fp16 GEMM on the 5th-generation tensor cores: TMA -> 128B-swizzled shared memory -> tcgen05.mma
// (accumulators in TMEM) -> tcgen05.ld epilogue.  Hand-written PTX, no CUTLASS / cuBLAS.
// The reference has no 16-bit GEMM at all (fp32/fp64 CUTLASS SIMT only, src/device/gemm_kernel.cu:26-36).
//
//   C[b][M,N] = alpha * op(A)[b] * op(B)[b] + beta * C[b]      (row-major storage, fp32 accumulate)
//
// Persistent, warp-specialised CTA (192 threads, one CTA per SM):
//   warp 0    : TMA producer  — fills a ring of NSTAGES {A 128x64, B BNx64} stages, arms `full[s]` with expect_tx
//   warp 1    : MMA issuer    — one elected lane issues 4 x tcgen05.mma (128 x BN x 16) per stage, commits to
//                               `empty[s]` (frees the stage) and, after the last k-block, to `tmem_full[acc]`
//   warps 2-5 : epilogue      — wait `tmem_full[acc]`, tcgen05.ld the 128 x BN fp32 tile (each warp its own
//                               32-lane TMEM quarter), scale / convert, 16-byte global stores, release `tmem_empty[acc]`
// TMEM holds two accumulator buffers (2 x BN columns) so the epilogue of tile i overlaps the main loop of tile i+1.
// Operand layouts: op(A) K-major ([M,K] storage) or MN-major ([K,M] storage, used by the backward passes),
// op(B) MN-major ([K,N] storage — kfunca's gemm(a, b[K,N])) or K-major ([N,K] storage).  Out-of-range rows /
// columns / k are zero-filled by TMA, so M, N, K need not be tile multiples; only 16-byte-aligned leading
// dimensions are required (otherwise the caller falls back to the SIMT kernel).
#include <cstdlib>

#include "ew_common.cuh"
#include "tc_common.cuh"

namespace kf {
using namespace tc;

constexpr int G_BM = 128;
constexpr int G_BK = 64;  // 64 x 2 B = one 128-byte swizzle atom
constexpr int G_THREADS = 192;

struct GemmTcParams {
    int64_t M, N, K, batch;
    int64_t ldc, sc;
    void *c;
    float alpha, beta;
    int m_tiles, n_tiles;
    int64_t total_tiles;
    int is_bf16;
    int a_bmul, b_bmul;  // 0 when the operand is broadcast over the batch
    // fused epilogues (README.md:32): out = alpha * acc + beta * C + residual
    const void *residual;  // same dtype as C, or null
    int64_t ldr, sr;
    // GLU (CTA-pair kernel only): the accumulator holds u = A B1 in columns [0, 128) and v = A B3 in [128, 256); out = u o v,
    // u / v optionally stored for the backward pass
    void *glu_u, *glu_v;
};

template <int BN>
struct GemmSmem {
    static constexpr int A_BYTES = G_BM * G_BK * 2;  // 16 KB
    static constexpr int B_BYTES = BN * G_BK * 2;    // 32 KB (BN = 256) / 16 KB (BN = 128)
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int NSTAGES = BN == 256 ? 4 : 6;
    static constexpr int BAR_BYTES = 256;
    static constexpr int TOTAL = NSTAGES * STAGE_BYTES + BAR_BYTES + 1024;  // + slack for 1024 B alignment
};

// tile index -> (batch, m_blk, n_blk); m fastest inside bands of 16 m-blocks so that concurrently running CTAs
// share A row panels and B column panels in L2
template <int BAND = 16>
__device__ __forceinline__ void tile_coords(int64_t t, const GemmTcParams &p, int &b, int &mb, int &nb) {
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

// One epilogue thread = one accumulator row: TMEM -> registers (32 fp32 columns at a time) -> alpha/beta -> 16-bit pack -> 16-byte stores
template <int BN>
__device__ __forceinline__ void epilogue_tile(const GemmTcParams &p, uint32_t taddr, int64_t row, int64_t n0, uint16_t *crow, const uint16_t *rrow,
                                              bool vec_ok) {
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
        uint32_t r[32];
        tmem_ld32(taddr + (uint32_t)c, r);
        tmem_ld_wait();
        if (row < p.M && n0 + c < p.N) {
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * p.alpha;
            const bool full_chunk = n0 + c + 32 <= p.N;
            if (p.beta != 0.f) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (n0 + c + i < p.N) {
                        const uint16_t old = crow[n0 + c + i];
                        const float o = p.is_bf16 ? __bfloat162float(__ushort_as_bfloat16(old)) : __half2float(__ushort_as_half(old));
                        v[i] += p.beta * o;
                    }
                }
            }
            if (rrow != nullptr) {  // + residual (x + gemm(...)): one pass instead of a GEMM store + an elementwise add
                uint32_t rw[16];
                if (full_chunk && vec_ok) {
                    const uint4 *src = reinterpret_cast<const uint4 *>(rrow + n0 + c);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint4 q = __ldg(src + i);
                        rw[4 * i] = q.x; rw[4 * i + 1] = q.y; rw[4 * i + 2] = q.z; rw[4 * i + 3] = q.w;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const uint32_t lo = (n0 + c + 2 * i < p.N) ? rrow[n0 + c + 2 * i] : 0u, hi = (n0 + c + 2 * i + 1 < p.N) ? rrow[n0 + c + 2 * i + 1] : 0u;
                        rw[i] = lo | (hi << 16);
                    }
                }
                // the unfused form rounds the product to 16 bits before the add: do the same so that both paths agree bit for bit
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    if (p.is_bf16) {
                        const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
                        const float2 g = __bfloat1622float2(h), r2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&rw[i]));
                        v[2 * i] = r2.x + g.x;
                        v[2 * i + 1] = r2.y + g.y;
                    } else {
                        const __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
                        const float2 g = __half22float2(h), r2 = __half22float2(*reinterpret_cast<const __half2 *>(&rw[i]));
                        v[2 * i] = r2.x + g.x;
                        v[2 * i + 1] = r2.y + g.y;
                    }
                }
            }
            uint32_t packed[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (p.is_bf16) {
                    __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
                    packed[i] = *reinterpret_cast<uint32_t *>(&h);
                } else {
                    __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
                    packed[i] = *reinterpret_cast<uint32_t *>(&h);
                }
            }
            if (full_chunk && vec_ok) {
                uint4 *dst = reinterpret_cast<uint4 *>(crow + n0 + c);
#pragma unroll
                for (int i = 0; i < 4; ++i) dst[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (n0 + c + i < p.N) crow[n0 + c + i] = (uint16_t)((packed[i >> 1] >> ((i & 1) * 16)) & 0xffffu);
                }
            }
        }
    }
}

// GLU epilogue: accumulator columns [c, c+32) hold u, [HN + c, HN + c + 32) hold v.  u and v are rounded to 16 bits first and the
// product is formed from the rounded values, exactly what gemm(), gemm() and `*` produce when they run as three launches.
template <int HN>
__device__ __forceinline__ void epilogue_tile_glu(const GemmTcParams &p, uint32_t taddr, int64_t row, int64_t n0, uint16_t *crow, uint16_t *urow,
                                                  uint16_t *vrow, bool vec_ok) {
#pragma unroll 1
    for (int c = 0; c < HN; c += 32) {
        uint32_t ru[32], rv[32];
        tmem_ld32(taddr + (uint32_t)c, ru);
        tmem_ld32(taddr + (uint32_t)(HN + c), rv);
        tmem_ld_wait();
        if (row < p.M && n0 + c < p.N) {
            uint32_t pu[16], pv[16], ph[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float u0 = __uint_as_float(ru[2 * i]) * p.alpha, u1 = __uint_as_float(ru[2 * i + 1]) * p.alpha;
                const float v0 = __uint_as_float(rv[2 * i]) * p.alpha, v1 = __uint_as_float(rv[2 * i + 1]) * p.alpha;
                if (p.is_bf16) {
                    const __nv_bfloat162 hu = __floats2bfloat162_rn(u0, u1), hv = __floats2bfloat162_rn(v0, v1);
                    const float2 fu = __bfloat1622float2(hu), fv = __bfloat1622float2(hv);
                    const __nv_bfloat162 hh = __floats2bfloat162_rn(fu.x * fv.x, fu.y * fv.y);
                    pu[i] = *reinterpret_cast<const uint32_t *>(&hu);
                    pv[i] = *reinterpret_cast<const uint32_t *>(&hv);
                    ph[i] = *reinterpret_cast<const uint32_t *>(&hh);
                } else {
                    const __half2 hu = __floats2half2_rn(u0, u1), hv = __floats2half2_rn(v0, v1);
                    const float2 fu = __half22float2(hu), fv = __half22float2(hv);
                    const __half2 hh = __floats2half2_rn(fu.x * fv.x, fu.y * fv.y);
                    pu[i] = *reinterpret_cast<const uint32_t *>(&hu);
                    pv[i] = *reinterpret_cast<const uint32_t *>(&hv);
                    ph[i] = *reinterpret_cast<const uint32_t *>(&hh);
                }
            }
            const bool full_chunk = n0 + c + 32 <= p.N;
            auto put = [&](uint16_t *dst_row, const uint32_t (&w)[16]) {
                if (full_chunk && vec_ok) {
                    uint4 *dst = reinterpret_cast<uint4 *>(dst_row + n0 + c);
#pragma unroll
                    for (int i = 0; i < 4; ++i) dst[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (n0 + c + i < p.N) dst_row[n0 + c + i] = (uint16_t)((w[i >> 1] >> ((i & 1) * 16)) & 0xffffu);
                }
            };
            put(crow, ph);
            if (urow) put(urow, pu);
            if (vrow) put(vrow, pv);
        }
    }
}

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(G_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const GemmTcParams p) {
    using S = GemmSmem<BN>;
    constexpr int NST = S::NSTAGES;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + NST * S::STAGE_BYTES);
    uint64_t *empty = full + NST;
    uint64_t *tmem_full = empty + NST;
    uint64_t *tmem_empty = tmem_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = (int)((p.K + G_BK - 1) / G_BK);

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_a);
        prefetch_tmap(&tmap_b);
        for (int s = 0; s < NST; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 4);  // one arrive per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 2 * BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================================================== TMA producer
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int64_t t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                int b, mb, nb;
                tile_coords<16>(t, p, b, mb, nb);
                const int m0 = mb * G_BM, n0 = nb * BN;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&empty[s], ph ^ 1);
                    unsigned char *sa = smem + s * S::STAGE_BYTES;
                    unsigned char *sb = sa + S::A_BYTES;
                    mbar_arrive_expect_tx(&full[s], S::STAGE_BYTES);
                    const int k0 = kb * G_BK;
                    if constexpr (!A_MN) {
                        tma_load_3d(sa, &tmap_a, &full[s], k0, m0, b * p.a_bmul);  // box 64(k) x 128(m)
                    } else {
#pragma unroll
                        for (int i = 0; i < G_BM / 64; ++i)  // boxes 64(m) x 64(k), one per 64-wide m atom
                            tma_load_3d(sa + i * (64 * G_BK * 2), &tmap_a, &full[s], m0 + i * 64, k0, b * p.a_bmul);
                    }
                    if constexpr (!B_MN) {
                        tma_load_3d(sb, &tmap_b, &full[s], k0, n0, b * p.b_bmul);  // box 64(k) x BN(n)
                    } else {
#pragma unroll
                        for (int i = 0; i < BN / 64; ++i)  // boxes 64(n) x 64(k)
                            tma_load_3d(sb + i * (64 * G_BK * 2), &tmap_b, &full[s], n0 + i * 64, k0, b * p.b_bmul);
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
        // ===================================================== MMA issuer (converged warp, elected lane issues: see umma_f16_p)
        const bool leader = elect_one();
        {
            const uint32_t idesc = make_idesc_f16(p.is_bf16 ? 1 : 0, A_MN ? 1 : 0, B_MN ? 1 : 0, G_BM, BN);
            int s = 0;
            uint32_t ph = 0;
            int acc = 0;
            uint32_t acc_ph = 0;
            for (int64_t t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                mbar_wait(&tmem_empty[acc], acc_ph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + s * S::STAGE_BYTES);
                    const uint32_t sb = sa + S::A_BYTES;
#pragma unroll
                    for (int k = 0; k < G_BK / 16; ++k) {
                        // K-major: the 16-element k slice sits 32 B further inside the 128 B swizzle atom (SBO = 8 rows x 128 B)
                        // MN-major: k advances by whole 8-row groups (2 x 1024 B); LBO = stride between 64-wide MN atoms
                        const uint64_t adesc = A_MN ? make_sw128_desc(sa + k * 2048, 64 * G_BK * 2, 1024) : make_sw128_desc(sa + k * 32, 0, 1024);
                        const uint64_t bdesc = B_MN ? make_sw128_desc(sb + k * 2048, 64 * G_BK * 2, 1024) : make_sw128_desc(sb + k * 32, 0, 1024);
                        umma_f16_p(d_tmem, adesc, bdesc, idesc, (kb | k) ? 1u : 0u, leader);
                    }
                    umma_commit_p(&empty[s], leader);  // stage reusable once these MMAs have read it
                    if (kb == nkb - 1) umma_commit_p(&tmem_full[acc], leader);
                    if (++s == NST) {
                        s = 0;
                        ph ^= 1;
                    }
                }
                if (++acc == 2) {
                    acc = 0;
                    acc_ph ^= 1;
                }
            }
        }
        __syncwarp();
    } else {
        // ===================================================== epilogue (warps 2..5 -> TMEM lane quarters 2,3,0,1)
        const int q = warp & 3;
        int acc = 0;
        uint32_t acc_ph = 0;
        const bool vec_ok = (p.ldc % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.c) & 15) == 0) && (p.sc % 8 == 0) &&
                            (!p.residual || ((p.ldr % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.residual) & 15) == 0) && (p.sr % 8 == 0)));
        for (int64_t t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
            int b, mb, nb;
            tile_coords<16>(t, p, b, mb, nb);
            const int64_t row = (int64_t)mb * G_BM + q * 32 + lane;
            const int64_t n0 = (int64_t)nb * BN;
            mbar_wait(&tmem_full[acc], acc_ph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
            uint16_t *crow = reinterpret_cast<uint16_t *>(p.c) + (int64_t)b * p.sc + row * p.ldc;
            const uint16_t *rrow = p.residual ? reinterpret_cast<const uint16_t *>(p.residual) + (int64_t)b * p.sr + row * p.ldr : nullptr;
            epilogue_tile<BN>(p, taddr, row, n0, crow, rrow, vec_ok);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            if (++acc == 2) {
                acc = 0;
                acc_ph ^= 1;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * BN);
    }
}

template <int BN, bool A_MN, bool B_MN>
static void launch_tc_cfg(const GemmPlan &g) {
    using S = GemmSmem<BN>;
    Runtime &rt = Runtime::get();
    const bool bf16 = g.dtype == KF_BFLOAT16;
    const uint64_t a_nb = (g.batch > 1 && g.sa == 0) ? 1 : (uint64_t)g.batch, b_nb = (g.batch > 1 && g.sb == 0) ? 1 : (uint64_t)g.batch;
    // A: K-major -> dims (K, M, batch), box (64, 128); MN-major ([K,M] storage) -> dims (M, K, batch), box (64, 64)
    const CUtensorMap ta = A_MN ? make_tmap_3d_16bit(g.a, bf16, (uint64_t)g.M, (uint64_t)g.K, a_nb, (uint64_t)g.lda, (uint64_t)g.sa, 64, 64)
                                : make_tmap_3d_16bit(g.a, bf16, (uint64_t)g.K, (uint64_t)g.M, a_nb, (uint64_t)g.lda, (uint64_t)g.sa, 64, G_BM);
    // B: MN-major ([K,N] storage) -> dims (N, K, batch), box (64, 64); K-major ([N,K] storage) -> dims (K, N, batch), box (64, BN)
    const CUtensorMap tb = B_MN ? make_tmap_3d_16bit(g.b, bf16, (uint64_t)g.N, (uint64_t)g.K, b_nb, (uint64_t)g.ldb, (uint64_t)g.sb, 64, 64)
                                : make_tmap_3d_16bit(g.b, bf16, (uint64_t)g.K, (uint64_t)g.N, b_nb, (uint64_t)g.ldb, (uint64_t)g.sb, 64, BN);
    GemmTcParams p{};
    p.M = g.M; p.N = g.N; p.K = g.K; p.batch = g.batch;
    p.ldc = g.ldc; p.sc = g.sc; p.c = g.c;
    p.alpha = g.alpha; p.beta = g.beta;
    p.m_tiles = (int)((g.M + G_BM - 1) / G_BM);
    p.n_tiles = (int)((g.N + BN - 1) / BN);
    p.total_tiles = (int64_t)p.m_tiles * p.n_tiles * g.batch;
    p.is_bf16 = bf16;
    p.a_bmul = a_nb > 1 || g.batch == 1 ? 1 : 0;
    p.b_bmul = b_nb > 1 || g.batch == 1 ? 1 : 0;
    p.residual = g.residual; p.ldr = g.ldr; p.sr = g.sr;
    auto kernel = gemm_tc_kernel<BN, A_MN, B_MN>;
    static bool attr_done = false;
    if (!attr_done) {
        KF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
        attr_done = true;
    }
    const int64_t grid = std::min<int64_t>(p.total_tiles, rt.props().sm_count);
    kernel<<<(unsigned)grid, G_THREADS, S::TOTAL, rt.stream()>>>(ta, tb, p);
    rt.post_launch("gemm_tc_kernel");
}


// ------------------------------------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2): a cluster of two CTAs (one TPC) owns a 256 x BN output tile.  Each CTA stages only
// its own 128 rows of A and its own BN/2 columns of B (32 KB per stage instead of 48 KB -> a deeper ring and half the
// shared-memory read traffic per MMA for B, which is what limits the single-CTA kernel), and holds the 128 x BN half of the
// accumulator in its own TMEM.  Protocol:
//   * producers (warp 0 of BOTH CTAs) wait on their local `empty[s]` and credit their TMA bytes to the LEADER's `full[s]`
//   * the leader's MMA warp issues one M = 256 MMA per 16-wide k slice and multicasts its commits to `empty[s]` / `tmem_full[a]`
//     of both CTAs
//   * epilogue warps of both CTAs drain their own TMEM half and arrive on the leader's `tmem_empty[a]` (count 8)
template <int BN>
struct GemmSmem2 {
    static constexpr int A_BYTES = G_BM * G_BK * 2;        // 16 KB: this CTA's 128 rows
    static constexpr int B_BYTES = (BN / 2) * G_BK * 2;    // 16 KB: this CTA's BN/2 columns
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;  // 32 KB
    static constexpr int NSTAGES = 6;
    static constexpr int BAR_BYTES = 256;
    static constexpr int TOTAL = NSTAGES * STAGE_BYTES + BAR_BYTES + 1024;
};

// GLU = true: rank 0 stages B1[:, n0 .. n0+128) and rank 1 stages B3[:, n0 .. n0+128) (tmap_b2), so ONE M256 N256 MMA stream
// produces u = A B1 in accumulator columns [0, 128) and v = A B3 in [128, 256); the epilogue writes u o v (and u, v when asked).
template <int BN, bool A_MN, bool B_MN, bool GLU>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_b2,
                const GemmTcParams p) {
    using S = GemmSmem2<BN>;
    constexpr int NST = S::NSTAGES;
    constexpr int HALF_N = BN / 2;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + NST * S::STAGE_BYTES);
    uint64_t *empty = full + NST;
    uint64_t *tmem_full = empty + NST;
    uint64_t *tmem_empty = tmem_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int64_t cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int nkb = (int)((p.K + G_BK - 1) / G_BK);

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_a);
        prefetch_tmap(&tmap_b);
        for (int s = 0; s < NST; ++s) {
            mbar_init(&full[s], 1);   // used in the leader only: its producer's arrive.expect_tx (bytes of both CTAs)
            mbar_init(&empty[s], 1);  // multicast commit from the leader's MMA warp
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);   // multicast commit
            mbar_init(&tmem_empty[i], 8);  // used in the leader only: 4 epilogue warps x 2 CTAs
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2sm(tmem_slot, 2 * BN);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================================================== TMA producer (both CTAs)
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            const uint32_t full0 = mapa_u32(&full[0], 0);  // the leader's full[] array in the cluster window
            for (int64_t t = cluster_id; t < p.total_tiles; t += n_clusters) {
                int b, mb, nb;
                tile_coords<8>(t, p, b, mb, nb);
                const int m0 = mb * (2 * G_BM) + (int)rank * G_BM, n0 = GLU ? nb * HALF_N : nb * BN + (int)rank * HALF_N;
                const CUtensorMap *tmb = (GLU && rank == 1) ? &tmap_b2 : &tmap_b;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&empty[s], ph ^ 1);
                    unsigned char *sa = smem + s * S::STAGE_BYTES;
                    unsigned char *sb = sa + S::A_BYTES;
                    if (rank == 0) mbar_arrive_expect_tx(&full[s], 2 * S::STAGE_BYTES);
                    const uint32_t fb = full0 + (uint32_t)s * 8u;
                    const int k0 = kb * G_BK;
                    if constexpr (!A_MN) {
                        tma_load_3d_2sm(sa, &tmap_a, fb, k0, m0, b * p.a_bmul);  // box 64(k) x 128(m)
                    } else {
#pragma unroll
                        for (int i = 0; i < G_BM / 64; ++i) tma_load_3d_2sm(sa + i * (64 * G_BK * 2), &tmap_a, fb, m0 + i * 64, k0, b * p.a_bmul);
                    }
                    if constexpr (!B_MN) {
                        tma_load_3d_2sm(sb, tmb, fb, k0, n0, b * p.b_bmul);  // box 64(k) x BN/2(n)
                    } else {
#pragma unroll
                        for (int i = 0; i < HALF_N / 64; ++i) tma_load_3d_2sm(sb + i * (64 * G_BK * 2), tmb, fb, n0 + i * 64, k0, b * p.b_bmul);
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
            const uint32_t idesc = make_idesc_f16(p.is_bf16 ? 1 : 0, A_MN ? 1 : 0, B_MN ? 1 : 0, 2 * G_BM, BN);
            int s = 0;
            uint32_t ph = 0;
            int acc = 0;
            uint32_t acc_ph = 0;
            for (int64_t t = cluster_id; t < p.total_tiles; t += n_clusters) {
                mbar_wait_cluster(&tmem_empty[acc], acc_ph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait_cluster(&full[s], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + s * S::STAGE_BYTES);
                    const uint32_t sb = sa + S::A_BYTES;
#pragma unroll
                    for (int k = 0; k < G_BK / 16; ++k) {
                        const uint64_t adesc = A_MN ? make_sw128_desc(sa + k * 2048, 64 * G_BK * 2, 1024) : make_sw128_desc(sa + k * 32, 0, 1024);
                        const uint64_t bdesc = B_MN ? make_sw128_desc(sb + k * 2048, 64 * G_BK * 2, 1024) : make_sw128_desc(sb + k * 32, 0, 1024);
                        umma_f16_2sm_p(d_tmem, adesc, bdesc, idesc, (kb | k) ? 1u : 0u, leader);
                    }
                    umma_commit_2sm_mc_p(&empty[s], leader);
                    if (kb == nkb - 1) umma_commit_2sm_mc_p(&tmem_full[acc], leader);
                    if (++s == NST) {
                        s = 0;
                        ph ^= 1;
                    }
                }
                if (++acc == 2) {
                    acc = 0;
                    acc_ph ^= 1;
                }
            }
        }
        __syncwarp();
    } else {
        // ===================================================== epilogue (both CTAs; warps 2..5 -> TMEM lane quarters 2,3,0,1)
        const int q = warp & 3;
        int acc = 0;
        uint32_t acc_ph = 0;
        const bool vec_ok = (p.ldc % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.c) & 15) == 0) && (p.sc % 8 == 0) &&
                            (!p.residual || ((p.ldr % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.residual) & 15) == 0) && (p.sr % 8 == 0))) &&
                            (!p.glu_u || (((reinterpret_cast<uintptr_t>(p.glu_u) | reinterpret_cast<uintptr_t>(p.glu_v)) & 15) == 0));
        const uint32_t tmem_empty0 = mapa_u32(&tmem_empty[0], 0);
        for (int64_t t = cluster_id; t < p.total_tiles; t += n_clusters) {
            int b, mb, nb;
            tile_coords<8>(t, p, b, mb, nb);
            const int64_t row = (int64_t)mb * (2 * G_BM) + (int64_t)rank * G_BM + q * 32 + lane;
            const int64_t n0 = (int64_t)nb * (GLU ? HALF_N : BN);
            mbar_wait_cluster(&tmem_full[acc], acc_ph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
            uint16_t *crow = reinterpret_cast<uint16_t *>(p.c) + (int64_t)b * p.sc + row * p.ldc;
            if constexpr (GLU) {
                const int64_t off = (int64_t)b * p.sc + row * p.ldc;
                epilogue_tile_glu<HALF_N>(p, taddr, row, n0, crow, p.glu_u ? reinterpret_cast<uint16_t *>(p.glu_u) + off : nullptr,
                                          p.glu_v ? reinterpret_cast<uint16_t *>(p.glu_v) + off : nullptr, vec_ok);
            } else {
                const uint16_t *rrow = p.residual ? reinterpret_cast<const uint16_t *>(p.residual) + (int64_t)b * p.sr + row * p.ldr : nullptr;
                epilogue_tile<BN>(p, taddr, row, n0, crow, rrow, vec_ok);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tmem_empty0 + (uint32_t)acc * 8u);
            if (++acc == 2) {
                acc = 0;
                acc_ph ^= 1;
            }
        }
    }
    // neither CTA may leave while the other can still touch its shared memory, barriers or tensor memory
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, 2 * BN);
    }
}

template <int BN, bool A_MN, bool B_MN, bool GLU = false>
static void launch_tc2_cfg(const GemmPlan &g) {
    using S = GemmSmem2<BN>;
    Runtime &rt = Runtime::get();
    const bool bf16 = g.dtype == KF_BFLOAT16;
    const uint64_t a_nb = (g.batch > 1 && g.sa == 0) ? 1 : (uint64_t)g.batch, b_nb = (g.batch > 1 && g.sb == 0) ? 1 : (uint64_t)g.batch;
    const CUtensorMap ta = A_MN ? make_tmap_3d_16bit(g.a, bf16, (uint64_t)g.M, (uint64_t)g.K, a_nb, (uint64_t)g.lda, (uint64_t)g.sa, 64, 64)
                                : make_tmap_3d_16bit(g.a, bf16, (uint64_t)g.K, (uint64_t)g.M, a_nb, (uint64_t)g.lda, (uint64_t)g.sa, 64, G_BM);
    const CUtensorMap tb = B_MN ? make_tmap_3d_16bit(g.b, bf16, (uint64_t)g.N, (uint64_t)g.K, b_nb, (uint64_t)g.ldb, (uint64_t)g.sb, 64, 64)
                                : make_tmap_3d_16bit(g.b, bf16, (uint64_t)g.K, (uint64_t)g.N, b_nb, (uint64_t)g.ldb, (uint64_t)g.sb, 64, BN / 2);
    CUtensorMap tb2 = tb;
    if (GLU)
        tb2 = B_MN ? make_tmap_3d_16bit(g.b2, bf16, (uint64_t)g.N, (uint64_t)g.K, b_nb, (uint64_t)g.ldb, (uint64_t)g.sb, 64, 64)
                   : make_tmap_3d_16bit(g.b2, bf16, (uint64_t)g.K, (uint64_t)g.N, b_nb, (uint64_t)g.ldb, (uint64_t)g.sb, 64, BN / 2);
    GemmTcParams p{};
    p.M = g.M; p.N = g.N; p.K = g.K; p.batch = g.batch;
    p.ldc = g.ldc; p.sc = g.sc; p.c = g.c;
    p.alpha = g.alpha; p.beta = g.beta;
    p.residual = g.residual; p.ldr = g.ldr; p.sr = g.sr;
    p.glu_u = g.glu_u; p.glu_v = g.glu_v;
    p.m_tiles = (int)((g.M + 2 * G_BM - 1) / (2 * G_BM));
    p.n_tiles = (int)((g.N + (GLU ? BN / 2 : BN) - 1) / (GLU ? BN / 2 : BN));
    p.total_tiles = (int64_t)p.m_tiles * p.n_tiles * g.batch;
    p.is_bf16 = bf16;
    p.a_bmul = a_nb > 1 || g.batch == 1 ? 1 : 0;
    p.b_bmul = b_nb > 1 || g.batch == 1 ? 1 : 0;
    auto kernel = gemm_tc2_kernel<BN, A_MN, B_MN, GLU>;
    static bool attr_done = false;
    if (!attr_done) {
        KF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
        attr_done = true;
    }
    const int64_t clusters = std::min<int64_t>(p.total_tiles, rt.props().sm_count / 2);
    kernel<<<(unsigned)(2 * clusters), G_THREADS, S::TOTAL, rt.stream()>>>(ta, tb, tb2, p);
    rt.post_launch(GLU ? "gemm_tc2_glu_kernel" : "gemm_tc2_kernel");
}

// h = (A B1) o (A B3) in one launch (GLU feed-forward).  A [M,K] row-major, B1 / B3 [K,N] row-major (or both
// transposed), same leading dimension.  false => caller composes it from two GEMMs and a multiply.
bool launch_gemm_glu_tc(const GemmPlan &g) {
    if (g.dtype != KF_HALF && g.dtype != KF_BFLOAT16) return false;
    if (g.b2 == nullptr || g.batch != 1 || g.trans_a || g.beta != 0.f || g.residual) return false;
    if (g.M <= 128 || g.N < 128 || g.K <= 0) return false;
    auto aligned = [](const void *p, int64_t ld) { return (reinterpret_cast<uintptr_t>(p) % 16 == 0) && (ld % 8 == 0); };
    if (!aligned(g.a, g.lda) || !aligned(g.b, g.ldb) || !aligned(g.b2, g.ldb)) return false;
    if (g.M >= (1ll << 31) || g.N >= (1ll << 31) || g.K >= (1ll << 31)) return false;
    if (g.trans_b) launch_tc2_cfg<256, false, false, true>(g);
    else launch_tc2_cfg<256, false, true, true>(g);
    return true;
}

bool launch_gemm_tc(const GemmPlan &g) {
    static const bool force_simt = std::getenv("KF_GEMM_FORCE_SIMT") != nullptr;
    if (force_simt) return false;
    if (g.dtype != KF_HALF && g.dtype != KF_BFLOAT16) return false;
    if (g.M <= 0 || g.N <= 0 || g.K <= 0) return false;
    // TMA: 16-byte aligned base and leading dimensions (8 x 16-bit elements); dims < 2^32
    auto aligned = [](const void *p, int64_t ld, int64_t bs) { return (reinterpret_cast<uintptr_t>(p) % 16 == 0) && (ld % 8 == 0) && (bs % 8 == 0); };
    if (!aligned(g.a, g.lda, g.sa) || !aligned(g.b, g.ldb, g.sb)) return false;
    if (g.M >= (1ll << 31) || g.N >= (1ll << 31) || g.K >= (1ll << 31) || g.batch >= 65536) return false;
    const bool small_n = g.N <= 128;
    // CTA pairs (256 x 256 tiles) once there is at least one full wave of them; KF_GEMM_CTA_GROUP=1|2 forces a variant
    const char *cg_env = std::getenv("KF_GEMM_CTA_GROUP");  // read per call so that tests can flip it
    const int64_t pair_tiles = ((g.M + 255) / 256) * ((g.N + 255) / 256) * g.batch;
    const bool pair = cg_env ? (cg_env[0] == '2' && !small_n) : (!small_n && g.M > 128 && pair_tiles >= Runtime::get().props().sm_count / 2);
#define KF_TC_DISPATCH(FN, BN)                                                 \
    do {                                                                       \
        if (g.trans_a && !g.trans_b) FN<BN, true, true>(g);                    \
        else if (g.trans_a && g.trans_b) FN<BN, true, false>(g);               \
        else if (!g.trans_a && !g.trans_b) FN<BN, false, true>(g);             \
        else FN<BN, false, false>(g);                                          \
    } while (0)
    if (small_n) KF_TC_DISPATCH(launch_tc_cfg, 128);
    else if (pair) KF_TC_DISPATCH(launch_tc2_cfg, 256);
    else KF_TC_DISPATCH(launch_tc_cfg, 256);
#undef KF_TC_DISPATCH
    return true;
}

}  // namespace kf

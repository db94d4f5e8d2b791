// Causal attention forward on tcgen05 tensor cores (bf16 / fp16, head size 64 or 128), hand-written PTX.
// The reference's forward is fp32 scalar-FMA only and walks every KV block (src/device/utils/causal_attention.h:73-208;
// SURVEY F2); this kernel is its B200-native replacement for the 16-bit dtypes the north star asks for.
//
// One CTA = 128 query rows of one (batch, head); KV consumed in blocks of 128 keys up to the diagonal.
//   warp 0    : TMA producer (Q once; K_j, V_j per block; 128B-swizzled tiles)
//   warp 1    : MMA issuer   S = Q K_j^T  (128x128xD, both operands K-major in smem, fp32 accumulator in TMEM)
//                            O += P V_j   (128xDx128, P read from TENSOR MEMORY, V MN-major in smem)
//   warps 2-5 : softmax      thread = query row: tcgen05.ld S, online max/sum in the exp2 domain, P (16-bit)
//                            written back over S in TMEM (tcgen05.st), O rescaled in TMEM when the row max moved,
//                            final 1/l scaling + 16-byte global stores + row LSE.
// TMEM budget is 256 columns (S/P 128 + O D) and shared memory < 100 KB so that TWO CTAs share an SM:
// while one CTA is in its softmax phase the other one keeps the tensor pipe busy.
#include <cmath>
#include <cstdlib>

#include "ew_common.cuh"
#include "tc_common.cuh"

namespace kf {
using namespace tc;

constexpr int FA_BQ = 128, FA_BKV = 128, FA_THREADS = 192;

struct AttnTcParams {
    int64_t BH, Sq, Skv;
    void *out;
    float *lse;
    float scale_log2;  // softmax scale * log2(e)
    int nq;            // query blocks per (b, h)
    int is_bf16;
};

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
          "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
          "r"(r[30]), "r"(r[31])
        : "memory");
}

__device__ __forceinline__ uint32_t pack16(float a, float b, int is_bf16) {
    if (is_bf16) {
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t *>(&h);
    }
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}

template <int D>
__global__ void __launch_bounds__(FA_THREADS, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                   const __grid_constant__ CUtensorMap tmap_v, const AttnTcParams p) {
    constexpr int ATOMS = D / 64;              // 64-element (128 B) swizzle atoms along the head dimension
    constexpr int TILE_BYTES = 128 * D * 2;    // one 128 x D 16-bit tile
    constexpr int ATOM_BYTES = 128 * 128;      // 128 rows x 128 B
    constexpr uint32_t TMEM_COLS = 256;
    constexpr uint32_t O_COL = 128;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char *sQ = smem, *sK = smem + TILE_BYTES, *sV = smem + 2 * TILE_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 3 * TILE_BYTES);
    uint64_t *q_full = bars + 0, *k_full = bars + 1, *v_full = bars + 2, *k_empty = bars + 3, *s_full = bars + 4, *p_full = bars + 5,
             *pv_done = bars + 6;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bh = blockIdx.x / p.nq;
    const int qb = p.nq - 1 - (blockIdx.x % p.nq);  // heaviest (longest KV range) query blocks first
    const int q0 = qb * FA_BQ;
    const int kv_end = (int)min((int64_t)p.Skv, (int64_t)q0 + FA_BQ);
    const int nblk = (kv_end + FA_BKV - 1) / FA_BKV;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_q);
        prefetch_tmap(&tmap_k);
        prefetch_tmap(&tmap_v);
        mbar_init(q_full, 1);
        mbar_init(k_full, 1);
        mbar_init(v_full, 1);
        mbar_init(k_empty, 1);
        mbar_init(s_full, 1);
        mbar_init(p_full, 4);
        mbar_init(pv_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, TILE_BYTES);
#pragma unroll
            for (int a = 0; a < ATOMS; ++a) tma_load_3d(sQ + a * ATOM_BYTES, &tmap_q, q_full, a * 64, q0, bh);
            for (int j = 0; j < nblk; ++j) {
                const int kv0 = j * FA_BKV;
                mbar_wait(k_empty, (j & 1) ^ 1);
                mbar_arrive_expect_tx(k_full, TILE_BYTES);
#pragma unroll
                for (int a = 0; a < ATOMS; ++a) tma_load_3d(sK + a * ATOM_BYTES, &tmap_k, k_full, a * 64, kv0, bh);
                mbar_wait(pv_done, (j & 1) ^ 1);  // V tile free once P V_{j-1} has completed
                mbar_arrive_expect_tx(v_full, TILE_BYTES);
#pragma unroll
                for (int a = 0; a < ATOMS; ++a) tma_load_3d(sV + a * ATOM_BYTES, &tmap_v, v_full, a * 64, kv0, bh);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            const int fmt = p.is_bf16 ? 1 : 0;
            const uint32_t idesc_s = make_idesc_f16(fmt, 0, 0, FA_BQ, FA_BKV);
            const uint32_t idesc_pv = make_idesc_f16(fmt, 0, 1, FA_BQ, D);
            const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK), v_addr = smem_u32(sV);
            mbar_wait(q_full, 0);
            for (int j = 0; j < nblk; ++j) {
                mbar_wait(k_full, j & 1);
                if (j > 0) mbar_wait(pv_done, (j - 1) & 1);  // P_{j-1} (aliased with S) consumed, O updated
                tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk) {
                    const uint32_t off = (uint32_t)((kk >> 2) * ATOM_BYTES + (kk & 3) * 32);
                    umma_f16(tmem_base, make_sw128_desc(q_addr + off, 0, 1024), make_sw128_desc(k_addr + off, 0, 1024), idesc_s, kk ? 1u : 0u);
                }
                umma_commit(k_empty);
                umma_commit(s_full);
                mbar_wait(p_full, j & 1);
                mbar_wait(v_full, j & 1);
                tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < FA_BKV / 16; ++kk) {
                    // A = P from tensor memory: 16 k-values of 16 bits = 8 columns per step;
                    // B = V, MN-major: 16 kv rows = 2 x 1024 B per step, 64-wide d atoms ATOM_BYTES apart
                    umma_f16_ts(tmem_base + O_COL, tmem_base + (uint32_t)(kk * 8), make_sw128_desc(v_addr + kk * 2048, ATOM_BYTES, 1024), idesc_pv,
                                (j | kk) ? 1u : 0u);
                }
                umma_commit(pv_done);
            }
        }
        __syncwarp();
    } else {
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int64_t m_row = (int64_t)q0 + r;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        float m_run = -INFINITY, l_run = 0.f;
        for (int j = 0; j < nblk; ++j) {
            const int kv0 = j * FA_BKV;
            const bool need_mask = (kv0 + FA_BKV - 1 > q0) || (kv0 + FA_BKV > p.Skv);
            mbar_wait(s_full, j & 1);
            tc_fence_after();
            // ---- pass 1: row max (scaled to the exp2 domain)
            float mx = -INFINITY;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t sr[32];
                tmem_ld32(lane_addr + (uint32_t)(c * 32), sr);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float s = __uint_as_float(sr[i]);
                    if (need_mask) {
                        const int64_t n = (int64_t)kv0 + c * 32 + i;
                        if (n > m_row || n >= p.Skv) s = -INFINITY;
                    }
                    mx = fmaxf(mx, s);
                }
            }
            const float m_new = fmaxf(m_run, mx * p.scale_log2);
            const float corr = (m_run == -INFINITY) ? 0.f : exp2f(m_run - m_new);
            // ---- pass 2: p = exp2(s * c - m), written back over S as packed 16-bit pairs
            float rs = 0.f;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t sr[32];
                tmem_ld32(lane_addr + (uint32_t)(c * 32), sr);
                tmem_ld_wait();
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float s0 = __uint_as_float(sr[i]), s1 = __uint_as_float(sr[i + 1]);
                    float p0 = exp2f(fmaf(s0, p.scale_log2, -m_new)), p1 = exp2f(fmaf(s1, p.scale_log2, -m_new));
                    if (need_mask) {
                        const int64_t n = (int64_t)kv0 + c * 32 + i;
                        if (n > m_row || n >= p.Skv) p0 = 0.f;
                        if (n + 1 > m_row || n + 1 >= p.Skv) p1 = 0.f;
                    }
                    rs += p0 + p1;
                    pk[i >> 1] = pack16(p0, p1, p.is_bf16);
                }
                tmem_st16(lane_addr + (uint32_t)(c * 16), pk);
            }
            l_run = l_run * corr + rs;
            m_run = m_new;
            // ---- rescale the running O accumulator when some row of this warp moved its max
            if (j > 0 && __any_sync(0xffffffffu, corr != 1.f)) {
#pragma unroll 1
                for (int c = 0; c < D / 32; ++c) {
                    uint32_t orr[32];
                    tmem_ld32(lane_addr + O_COL + (uint32_t)(c * 32), orr);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) orr[i] = __float_as_uint(__uint_as_float(orr[i]) * corr);
                    tmem_st32(lane_addr + O_COL + (uint32_t)(c * 32), orr);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
        }
        // ---- epilogue: O / l -> 16-bit -> global, row LSE
        mbar_wait(pv_done, (nblk - 1) & 1);
        tc_fence_after();
        const float inv_l = 1.f / l_run;
        const bool row_ok = m_row < p.Sq;
        uint16_t *orow = reinterpret_cast<uint16_t *>(p.out) + ((int64_t)bh * p.Sq + (row_ok ? m_row : 0)) * D;
#pragma unroll 1
        for (int c = 0; c < D / 32; ++c) {
            uint32_t orr[32];
            tmem_ld32(lane_addr + O_COL + (uint32_t)(c * 32), orr);  // .sync.aligned: every lane takes part, only the stores are predicated
            tmem_ld_wait();
            if (row_ok) {
                uint4 *dst = reinterpret_cast<uint4 *>(orow + c * 32);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint32_t w[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        w[k] = pack16(__uint_as_float(orr[8 * i + 2 * k]) * inv_l, __uint_as_float(orr[8 * i + 2 * k + 1]) * inv_l, p.is_bf16);
                    dst[i] = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
        }
        if (row_ok && p.lse) p.lse[(int64_t)bh * p.Sq + m_row] = (m_run + log2f(l_run)) * 0.6931471805599453f;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

template <int D>
static void launch_fwd_tc(const AttnPlan &a) {
    Runtime &rt = Runtime::get();
    const bool bf16 = a.dtype == KF_BFLOAT16;
    const CUtensorMap tq = make_tmap_3d_16bit(a.q, bf16, D, (uint64_t)a.Sq, (uint64_t)a.BH, D, (uint64_t)a.Sq * D, 64, 128);
    const CUtensorMap tk = make_tmap_3d_16bit(a.k, bf16, D, (uint64_t)a.Skv, (uint64_t)a.BH, D, (uint64_t)a.Skv * D, 64, 128);
    const CUtensorMap tv = make_tmap_3d_16bit(a.v, bf16, D, (uint64_t)a.Skv, (uint64_t)a.BH, D, (uint64_t)a.Skv * D, 64, 128);
    AttnTcParams p{};
    p.BH = a.BH; p.Sq = a.Sq; p.Skv = a.Skv;
    p.out = a.out;
    p.lse = reinterpret_cast<float *>(a.lse);
    p.scale_log2 = (float)(1.4426950408889634 / std::sqrt((double)D));
    p.nq = (int)((a.Sq + FA_BQ - 1) / FA_BQ);
    p.is_bf16 = bf16;
    constexpr int SMEM = 3 * 128 * D * 2 + 256 + 1024;
    auto kern = attn_fwd_tc_kernel<D>;
    static bool attr_done = false;
    if (!attr_done) {
        KF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        attr_done = true;
    }
    const int64_t grid = a.BH * p.nq;
    KF_CHECK(grid < (int64_t)0x7FFFFFFF);
    kern<<<(unsigned)grid, FA_THREADS, SMEM, rt.stream()>>>(tq, tk, tv, p);
    rt.post_launch("attn_fwd_tc_kernel");
}

bool launch_attention_fwd_tc(const AttnPlan &a) {
    static const bool force_simt = std::getenv("KF_ATTN_FORCE_SIMT") != nullptr;
    if (force_simt) return false;
    if (a.dtype != KF_HALF && a.dtype != KF_BFLOAT16) return false;
    if (a.D != 64 && a.D != 128) return false;
    if (a.Sq < 1 || a.Skv < 1 || a.BH < 1 || a.BH >= 65536) return false;
    auto al = [](const void *p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
    if (!al(a.q) || !al(a.k) || !al(a.v) || !al(a.out)) return false;
    if (a.D == 64) launch_fwd_tc<64>(a);
    else launch_fwd_tc<128>(a);
    return true;
}

}  // namespace kf

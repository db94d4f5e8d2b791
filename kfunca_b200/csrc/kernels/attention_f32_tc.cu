// fp32 causal attention FORWARD on tcgen05 tensor cores (head size 64 or 128, dense [BH, S, D] operands), hand-written PTX.
// Replaces: CausalAttentionForwardFN (fp32 FMA, no causal skip, /root/reference src/device/utils/causal_attention.h:73-208) — the only
// attention the reference has.  The tensor core has no fp32 operand type, so, as in gemm_f32_tc.cu, every fp32 operand is split into
// three bf16 planes x = x0 + x1 + x2 (x0 carries the leading 8 mantissa bits, x1 the next 8, x2 the rest: 24 bits in all) and a
// product is issued as the six plane pairs whose weight is >= 2^-16:  a b ~ a0 b0 + (a0 b1 + a1 b0 + a1 b1 + a0 b2 + a2 b0).
//
//   split_planes_kernel : q, k, v fp32 -> [3][BH][S][D] bf16 planes each (HBM-bound pre-pass, 4 B read + 6 B written per element)
//   attn_f32_tc_kernel  : one CTA per (batch-head, 128-row query tile), heaviest tiles first; KV blocks of 64 keys
//     warp 5    TMA producer: the three Q planes once (96 KB at D = 128), then K_0, [K_1, V_0], [K_2, V_1], ... through two 48 KB slots
//               (three planes of 64 keys each) in the order the issuer consumes them
//     warp 4    MMA issuer (converged warp, elected lane): S(j+1) = Q K_{j+1}^T is issued BEFORE P(j) V_j, so the tensor pipe works on
//               the next block's scores while the softmax warps are busy with this block's: 48 + 24 MMAs per block back to back
//     warps 0-3 softmax + accumulation, thread = query row.  S arrives in two accumulators (the leading product q0 k0 alone, the five
//               small ones together: the tensor core adds every MMA into its accumulator with truncation, and this way the eight
//               truncations of the leading sum are the only ones at full weight); they are added in IEEE fp32 in registers.
//               P = exp2(S c - m) is split into three bf16 planes in registers and written to tensor memory as the A operand of P V.
//               Every block's P V goes into a FRESH accumulator Ob (24 truncating additions, <= 1.5e-6 relative) and is folded into
//               the running output O by the softmax warps in IEEE fp32 (tcgen05.ld O, Ob -> O alpha + Ob -> tcgen05.st) while the
//               tensor pipe computes the next block's S — the running accumulator never sees a truncating addition and the
//               rescale by alpha = exp2(m_old - m_new) costs nothing extra.
//   Tensor memory (512 columns): S_hi 64 | S_lo 64 | P planes 3 x 32 | (32 free) | Ob D | O D.
// Cost: 6 bf16 MMAs per fp32 MMA on half-width (N = 64) score tiles: ~ 1/7 of the bf16 kernel's rate against ~ 1/100 for FFMA code.
// Accuracy (tests/test_attention_gpu.py): 1e-5 relative + absolute on U(-1, 1) inputs at every tested length, the reference's own
// 1e-3 on its U(-10, 10) shapes; KF_ATTN_F32=simt keeps the FFMA kernel.  Non-finite inputs: as in gemm_f32_tc.cu (inf splits into
// inf + NaN planes), documented deviation.
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "ew_common.cuh"
#include "tc_common.cuh"

namespace kf {
using namespace tc;

namespace {

constexpr int A32_BQ = 128, A32_BKV = 64;
constexpr int A32_THREADS = 192;

struct A32Params {
    int64_t BH, Sq, Skv;
    float *out;        // [BH, Sq, D] fp32
    float *lse;        // [BH, Sq] natural log-sum-exp, may be null
    float scale_log2;  // softmax scale * log2(e)
    int ntiles;        // 128-row query tiles per batch-head
};

// x [n] fp32 -> planes [3][n] bf16; n a multiple of 4, 16-byte aligned
__global__ void __launch_bounds__(256) split_planes_kernel(const float *__restrict__ x, __nv_bfloat16 *__restrict__ planes, const int64_t n) {
    pdl_enter();
    const int64_t quads = n / 4;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < quads; t += (int64_t)gridDim.x * blockDim.x) {
        const float4 f = __ldcs(reinterpret_cast<const float4 *>(x) + t);
        const float v[4] = {f.x, f.y, f.z, f.w};
        uint16_t h[3][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float rem = v[i];
            __nv_bfloat16 b0 = __float2bfloat16_rn(rem);
            float f0 = __bfloat162float(b0);
            if (isinf(f0) && !isinf(rem)) {  // rounding up to inf: the leading plane is truncated instead
                f0 = __uint_as_float(__float_as_uint(rem) & 0xffff0000u);
                b0 = __float2bfloat16_rn(f0);
            }
            rem = __fsub_rn(rem, f0);
            const __nv_bfloat16 b1 = __float2bfloat16_rn(rem);
            rem = __fsub_rn(rem, __bfloat162float(b1));
            const __nv_bfloat16 b2 = __float2bfloat16_rn(rem);
            h[0][i] = __bfloat16_as_ushort(b0);
            h[1][i] = __bfloat16_as_ushort(b1);
            h[2][i] = __bfloat16_as_ushort(b2);
        }
#pragma unroll
        for (int pl = 0; pl < 3; ++pl) {
            uint2 w;
            w.x = (uint32_t)h[pl][0] | ((uint32_t)h[pl][1] << 16);
            w.y = (uint32_t)h[pl][2] | ((uint32_t)h[pl][3] << 16);
            *reinterpret_cast<uint2 *>(planes + (int64_t)pl * n + t * 4) = w;
        }
    }
}

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void tmem_st32_(uint32_t taddr, const uint32_t (&r)[32]) {
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

// the six plane pairs (a, b) of a product, small ones first (into a fresh accumulator their order costs nothing), leading pair last
__device__ constexpr int PA[6] = {2, 0, 1, 1, 0, 0};
__device__ constexpr int PB[6] = {0, 2, 1, 0, 1, 0};

template <int D>
__global__ void __launch_bounds__(A32_THREADS, 1)
attn_f32_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                   const __grid_constant__ CUtensorMap tmap_v, const A32Params p) {
    constexpr int ATOMS = D / 64;                    // 64-element (128 B) swizzle atoms along the head dimension
    constexpr int Q_ATOM = A32_BQ * 128;             // bytes of one atom of a 128-row tile
    constexpr int KV_ATOM = A32_BKV * 128;           // ... of a 64-row tile
    constexpr int Q_PLANE = A32_BQ * D * 2, KV_PLANE = A32_BKV * D * 2;
    constexpr int SLOT = 3 * KV_PLANE;               // one K or V block: three planes
    constexpr uint32_t S_HI = 0, S_LO = 64, P_COL = 128, OB_COL = 256, O_COL = 256 + D;
    constexpr int W_MMA = 4, W_TMA = 5;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char *sQ = smem;                 // 3 planes
    unsigned char *sKV = smem + 3 * Q_PLANE;  // 2 slots x 3 planes
    uint64_t *bars = reinterpret_cast<uint64_t *>(sKV + 2 * SLOT);
    uint64_t *q_full = bars + 0;
    uint64_t *kv_full = bars + 1, *kv_empty = bars + 3;  // [2] each
    uint64_t *s_full = bars + 5;   // S(j) is complete in tensor memory
    uint64_t *s_free = bars + 6;   // the softmax warps hold S(j) in registers (4 arrivals)
    uint64_t *p_full = bars + 7;   // P(j) planes are in tensor memory and Ob is free (4 arrivals)
    uint64_t *ob_full = bars + 8;  // P(j) V_j is complete in Ob (and P(j) has been read)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 9);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bh = blockIdx.x / p.ntiles;
    const int tile = p.ntiles - 1 - (blockIdx.x % p.ntiles);  // heaviest (longest KV range) tiles of a head first
    const int q0 = tile * A32_BQ;
    const int64_t kv_end = min((int64_t)p.Skv, (int64_t)q0 + A32_BQ);
    const int nblk = (int)((kv_end + A32_BKV - 1) / A32_BKV);  // >= 1

    if (warp == W_TMA && lane == 0) {
        prefetch_tmap(&tmap_q);
        prefetch_tmap(&tmap_k);
        prefetch_tmap(&tmap_v);
        mbar_init(q_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(s_free, 4);
        mbar_init(p_full, 4);
        mbar_init(ob_full, 1);
        fence_barrier_init();
    }
    if (warp == W_MMA) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == W_TMA) {
        // ===================================================== TMA producer
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, 3 * Q_PLANE);
#pragma unroll
            for (int pl = 0; pl < 3; ++pl)
#pragma unroll
                for (int a = 0; a < ATOMS; ++a) tma_load_4d(sQ + pl * Q_PLANE + a * Q_ATOM, &tmap_q, q_full, a * 64, q0, bh, pl);
            // consumption order: K_0, K_1, V_0, K_2, V_1, ..., K_{n-1}, V_{n-2}, V_{n-1}
            int s = 0;
            uint32_t ph = 0;
            const int nloads = 2 * nblk;
            for (int i = 0; i < nloads; ++i) {
                const bool is_k = i == 0 || (i < nloads - 1 && (i & 1));
                const int blk = i == 0 ? 0 : (is_k ? (i + 1) / 2 : (i == nloads - 1 ? nblk - 1 : i / 2 - 1));
                const CUtensorMap *tm = is_k ? &tmap_k : &tmap_v;
                mbar_wait(&kv_empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&kv_full[s], SLOT);
#pragma unroll
                for (int pl = 0; pl < 3; ++pl)
#pragma unroll
                    for (int a = 0; a < ATOMS; ++a)
                        tma_load_4d(sKV + s * SLOT + pl * KV_PLANE + a * KV_ATOM, tm, &kv_full[s], a * 64, blk * A32_BKV, bh, pl);
                if (++s == 2) {
                    s = 0;
                    ph ^= 1;
                }
            }
        }
        __syncwarp();
    } else if (warp == W_MMA) {
        // ===================================================== MMA issuer
        const bool leader = elect_one();
        const uint32_t idesc_s = make_idesc_f16(1, 0, 0, A32_BQ, A32_BKV);  // S = Q K^T: both operands K-major, N = 64
        const uint32_t idesc_pv = make_idesc_f16(1, 0, 1, A32_BQ, D);       // Ob = P V: A from tensor memory, B MN-major
        const uint32_t q_addr = smem_u32(sQ), kv_addr = smem_u32(sKV);
        int s = 0;
        uint32_t ph = 0;
        auto next_slot = [&]() {
            mbar_wait(&kv_full[s], ph);
            const int cur = s;
            if (++s == 2) {
                s = 0;
                ph ^= 1;
            }
            return cur;
        };
        auto issue_s = [&](int slot) {
            const uint32_t k_base = kv_addr + slot * SLOT;
            // leading pair alone into S_hi
#pragma unroll
            for (int kk = 0; kk < D / 16; ++kk) {
                const uint32_t qo = (uint32_t)((kk >> 2) * Q_ATOM + (kk & 3) * 32), ko = (uint32_t)((kk >> 2) * KV_ATOM + (kk & 3) * 32);
                umma_f16_p(tmem_base + S_HI, make_sw128_desc(q_addr + qo, 0, 1024), make_sw128_desc(k_base + ko, 0, 1024), idesc_s, kk ? 1u : 0u, leader);
            }
#pragma unroll
            for (int pr = 0; pr < 5; ++pr) {
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk) {
                    const uint32_t qo = (uint32_t)(PA[pr] * Q_PLANE + (kk >> 2) * Q_ATOM + (kk & 3) * 32);
                    const uint32_t ko = (uint32_t)(PB[pr] * KV_PLANE + (kk >> 2) * KV_ATOM + (kk & 3) * 32);
                    umma_f16_p(tmem_base + S_LO, make_sw128_desc(q_addr + qo, 0, 1024), make_sw128_desc(k_base + ko, 0, 1024), idesc_s,
                               (pr || kk) ? 1u : 0u, leader);
                }
            }
        };
        auto issue_pv = [&](int slot) {
            const uint32_t v_base = kv_addr + slot * SLOT;
#pragma unroll
            for (int pr = 0; pr < 6; ++pr) {
#pragma unroll
                for (int kk = 0; kk < A32_BKV / 16; ++kk)  // 16 keys per step: 8 columns of a P plane, 16 rows (2 KB) of a V atom
                    umma_f16_ts_p(tmem_base + OB_COL, tmem_base + P_COL + (uint32_t)(PA[pr] * 32 + kk * 8),
                                  make_sw128_desc(v_base + PB[pr] * KV_PLANE + kk * 2048, KV_ATOM, 1024), idesc_pv, (pr || kk) ? 1u : 0u, leader);
            }
        };
        mbar_wait(q_full, 0);
        {
            const int sk = next_slot();
            tc_fence_after();
            issue_s(sk);
            umma_commit_p(s_full, leader);
            umma_commit_p(&kv_empty[sk], leader);
        }
        for (int j = 0; j < nblk; ++j) {
            if (j + 1 < nblk) {
                const int sk = next_slot();
                mbar_wait(s_free, (uint32_t)(j & 1));  // S(j) is in the softmax warps' registers
                tc_fence_after();
                issue_s(sk);
                umma_commit_p(s_full, leader);
                umma_commit_p(&kv_empty[sk], leader);
            }
            const int sv = next_slot();
            mbar_wait(p_full, (uint32_t)(j & 1));  // P(j) written, Ob(j - 1) folded
            tc_fence_after();
            issue_pv(sv);
            umma_commit_p(ob_full, leader);
            umma_commit_p(&kv_empty[sv], leader);
        }
        __syncwarp();
    } else {
        // ===================================================== softmax + accumulation: thread = query row
        const int r = warp * 32 + lane;
        const int64_t m_row = (int64_t)q0 + r;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
        const float sc = p.scale_log2;
        float m_run = -INFINITY, l_run = 0.f, alpha = 0.f;  // alpha: factor O must take before Ob of the block just issued is added
        // O(j) = O(j - 1) alpha_j + Ob(j), in IEEE fp32; `first`: O has no content yet
        auto fold = [&](bool first, float a) {
#pragma unroll 1
            for (int c = 0; c < D / 32; ++c) {
                uint32_t ob[32], o[32];
                tmem_ld32(lane_addr + OB_COL + (uint32_t)(c * 32), ob);
                if (!first) tmem_ld32(lane_addr + O_COL + (uint32_t)(c * 32), o);
                tmem_ld_wait();
                if (!first) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) ob[i] = __float_as_uint(fmaf(__uint_as_float(o[i]), a, __uint_as_float(ob[i])));
                }
                tmem_st32_(lane_addr + O_COL + (uint32_t)(c * 32), ob);
            }
            tmem_st_wait();
        };
        for (int j = 0; j < nblk; ++j) {
            const int kv0 = j * A32_BKV;
            mbar_wait(s_full, (uint32_t)(j & 1));
            tc_fence_after();
            float x[64];
            {
                uint32_t hi[32], lo[32];
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    tmem_ld32(lane_addr + S_HI + (uint32_t)(c * 32), hi);
                    tmem_ld32(lane_addr + S_LO + (uint32_t)(c * 32), lo);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) x[c * 32 + i] = (__uint_as_float(hi[i]) + __uint_as_float(lo[i])) * sc;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_free);  // the issuer may overwrite S with S(j + 1)
            const bool masked = (kv0 + A32_BKV - 1 > q0) || (kv0 + A32_BKV > p.Skv);  // diagonal / ragged block (CTA-uniform)
            if (masked) {
                const int64_t lim = min(m_row, p.Skv - 1) - kv0;  // columns i > lim are masked (top-left causal mask, keys < Skv)
#pragma unroll
                for (int i = 0; i < 64; ++i)
                    if (i > lim) x[i] = -INFINITY;
            }
            float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
            for (int i = 0; i < 64; i += 4) {
                mx0 = fmaxf(mx0, x[i]);
                mx1 = fmaxf(mx1, x[i + 1]);
                mx2 = fmaxf(mx2, x[i + 2]);
                mx3 = fmaxf(mx3, x[i + 3]);
            }
            const float m_new = fmaxf(m_run, fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)));  // finite: key 0 is never masked for block 0
            const float a_new = ex2f(m_run - m_new);                                    // first block: exp2(-inf) = 0
            float rs0 = 0.f, rs1 = 0.f;
            uint32_t pk[3][32];
#pragma unroll
            for (int i = 0; i < 64; i += 2) {
                const float p0 = ex2f(x[i] - m_new), p1 = ex2f(x[i + 1] - m_new);
                rs0 += p0;
                rs1 += p1;
                // three bf16 planes of each probability: 8 + 8 + 8 mantissa bits
                const __nv_bfloat162 h0 = __floats2bfloat162_rn(p0, p1);
                const float2 f0 = __bfloat1622float2(h0);
                const float r0a = p0 - f0.x, r0b = p1 - f0.y;
                const __nv_bfloat162 h1 = __floats2bfloat162_rn(r0a, r0b);
                const float2 f1 = __bfloat1622float2(h1);
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(r0a - f1.x, r0b - f1.y);
                pk[0][i >> 1] = *reinterpret_cast<const uint32_t *>(&h0);
                pk[1][i >> 1] = *reinterpret_cast<const uint32_t *>(&h1);
                pk[2][i >> 1] = *reinterpret_cast<const uint32_t *>(&h2);
            }
            l_run = fmaf(l_run, a_new, rs0 + rs1);
            m_run = m_new;
            if (j > 0) {  // fold the previous block's product into O (with the factor that block's maximum asked for), freeing Ob and P
                mbar_wait(ob_full, (uint32_t)((j - 1) & 1));
                tc_fence_after();
                fold(j == 1, alpha);
            }
            alpha = a_new;
#pragma unroll
            for (int pl = 0; pl < 3; ++pl) tmem_st32_(lane_addr + P_COL + (uint32_t)(pl * 32), pk[pl]);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
        }
        mbar_wait(ob_full, (uint32_t)((nblk - 1) & 1));
        tc_fence_after();
        // ---- epilogue: (O alpha + Ob) / l -> global, row LSE
        const float inv_l = 1.f / l_run;
        const bool row_ok = m_row < p.Sq;
        float *orow = p.out + ((int64_t)bh * p.Sq + (row_ok ? m_row : 0)) * D;
#pragma unroll 1
        for (int c = 0; c < D / 32; ++c) {
            uint32_t ob[32], o[32];
            tmem_ld32(lane_addr + OB_COL + (uint32_t)(c * 32), ob);
            if (nblk > 1) tmem_ld32(lane_addr + O_COL + (uint32_t)(c * 32), o);
            tmem_ld_wait();
            if (row_ok) {
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    float4 w;
                    float *wp = &w.x;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float v = nblk > 1 ? fmaf(__uint_as_float(o[i + k]), alpha, __uint_as_float(ob[i + k])) : __uint_as_float(ob[i + k]);
                        wp[k] = v * inv_l;
                    }
                    *reinterpret_cast<float4 *>(orow + c * 32 + i) = w;
                }
            }
        }
        if (row_ok && p.lse) p.lse[(int64_t)bh * p.Sq + m_row] = (m_run + log2f(l_run)) * 0.6931471805599453f;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

template <int D>
void launch_f32_tc(const AttnPlan &a) {
    Runtime &rt = Runtime::get();
    const int64_t nq = a.BH * a.Sq * D, nkv = a.BH * a.Skv * D;
    Scratch planes_q((size_t)nq * 6), planes_k((size_t)nkv * 6), planes_v((size_t)nkv * 6);
    auto split = [&](const void *src, Scratch &dst, int64_t n) {
        const int grid = (int)std::min<int64_t>((n / 4 + 255) / 256, (int64_t)rt.props().sm_count * 16);
        launch_pdl(split_planes_kernel, dim3(grid), dim3(256), 0, rt.stream(), reinterpret_cast<const float *>(src), dst.as<__nv_bfloat16>(), n);
        rt.post_launch("split_planes_kernel");
    };
    split(a.q, planes_q, nq);
    split(a.k, planes_k, nkv);
    split(a.v, planes_v, nkv);
    auto map = [&](const Scratch &pl, int64_t S, uint32_t rows) {
        return make_tmap_4d_16bit(pl.p, true, D, (uint64_t)S, (uint64_t)a.BH, 3, (uint64_t)D, (uint64_t)(S * D), (uint64_t)(a.BH * S * D), 64, rows);
    };
    const CUtensorMap tq = map(planes_q, a.Sq, A32_BQ), tk = map(planes_k, a.Skv, A32_BKV), tv = map(planes_v, a.Skv, A32_BKV);
    A32Params p{};
    p.BH = a.BH; p.Sq = a.Sq; p.Skv = a.Skv;
    p.out = reinterpret_cast<float *>(a.out);
    p.lse = reinterpret_cast<float *>(a.lse);
    p.scale_log2 = (float)(1.4426950408889634 / std::sqrt((double)D));
    p.ntiles = (int)((a.Sq + A32_BQ - 1) / A32_BQ);
    constexpr int SMEM = 3 * A32_BQ * D * 2 + 2 * 3 * A32_BKV * D * 2 + 256 + 1024;
    static bool attr_done = false;
    if (!attr_done) {
        KF_CUDA(cudaFuncSetAttribute(attn_f32_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        attr_done = true;
    }
    const int64_t grid = a.BH * p.ntiles;
    KF_CHECK(grid < (int64_t)0x7FFFFFFF);
    attn_f32_tc_kernel<D><<<(unsigned)grid, A32_THREADS, SMEM, rt.stream()>>>(tq, tk, tv, p);
    rt.post_launch("attn_f32_tc_kernel");
}

}  // namespace

// fp32, dense operands, head size 64 / 128; false => the caller uses the FFMA kernel
bool launch_attention_fwd_f32_tc(const AttnPlan &a) {
    const char *mode = std::getenv("KF_ATTN_F32");  // KF_ATTN_F32=simt: the FFMA kernel (read per call, A/B runs)
    if (mode && std::strcmp(mode, "simt") == 0) return false;
    if (a.dtype != KF_FLOAT || a.H > 0) return false;
    if (a.D != 64 && a.D != 128) return false;
    if (a.Sq < 1 || a.Skv < 1 || a.BH < 1 || a.BH >= 65536) return false;
    if (a.Sq >= ((int64_t)1 << 31) || a.Skv >= ((int64_t)1 << 31)) return false;
    auto al = [](const void *p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
    if (!al(a.q) || !al(a.k) || !al(a.v) || !al(a.out)) return false;
    // small problems: the three split passes and a 128-row tile granularity cost more than the FFMA kernel takes
    if (a.BH * a.Sq * a.Skv < (int64_t)1 << 16) return false;
    if (a.D == 64) launch_f32_tc<64>(a);
    else launch_f32_tc<128>(a);
    return true;
}

}  // namespace kf

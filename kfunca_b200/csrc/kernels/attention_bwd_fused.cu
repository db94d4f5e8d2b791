// Causal attention BACKWARD as ONE fused tcgen05 kernel: 5 GEMMs per tile pair instead of the 7 of the two-kernel scheme in
// attention_bwd_tc.cu (VERDICT r1 "Next round" #4).  bf16 / fp16, head size 128.  The reference has no attention backward at all
// (only AddGradFunction exists, src/core/binary_ops.cpp:16-43); the oracle is oracle.causal_attention_bwd (float64).
//
// One CTA per 128-row KV block j of one (batch, head); K_j, V_j stay in shared memory, the 64-row query tiles t >= 2 j stream
// through a 4-slot TMA ring as {Q_t, dO_t}:
//     S^T  = K_j Q_t^T,  dP^T = V_j dO_t^T                  128 x 64 fp32 in TMEM (lanes = kv rows)           [2 GEMMs]
//     P^T  = exp2(S^T c - lse_q),  dS^T = P^T o (dP^T - delta_q)   -> 16-bit, written back over S^T / dP^T in TMEM,
//                                                                 dS^T ALSO to shared memory (MN-major, 128B swizzle)
//     dV_j += P^T dO_t,   dK_j += dS^T Q_t                   A operand from TENSOR MEMORY                      [2 GEMMs]
//     dQ_t^T (this block's share) = K_j^T dS^T               A = K_j read MN-major from smem, B = dS^T from smem [1 GEMM]
// TMEM (512 columns): per set s in {0, 1} (even / odd tiles ping-pong): [S^T | dP^T] = 128 columns, reused after the element-wise
// phase as [P^T 32 | dQ^T 64 | dS^T 32]; dV 128; dK 128.
//
// dQ is the only output shared between CTAs.  It is accumulated DETERMINISTICALLY: the contributions of the KV blocks to query
// tile t are added in the fixed order j = t/2, t/2 - 1, ..., 0 through an fp32 buffer and a per-tile counter (acquire / release);
// the last contributor (j = 0) applies the softmax scale, rounds and writes dQ.  CTAs are launched with j descending inside a
// (batch, head), so a CTA only ever waits for CTAs dispatched before it (no deadlock), and since CTA j reaches tile t two steps
// after CTA j + 1 does, the waits are normally already satisfied: a wavefront, not a serial chain.  No atomics on data: two runs
// give bit-identical gradients.
// The accumulation buffer lives in L2 between consecutive contributors (a (b, h) group's 2 MB share is touched by CTAs that run
// at the same time), so the extra DRAM traffic is its eventual write-back, not a read-modify-write per contribution.
#include <cmath>
#include <cstdlib>

#include "ew_common.cuh"
#include "tc_common.cuh"

namespace kf {
using namespace tc;

namespace {

constexpr int FB_THREADS = 320;  // warps 0-3: element-wise set 0, 4-7: set 1, 8: MMA issuer, 9: TMA producer
constexpr int FB_NS = 4;         // {Q_t, dO_t} ring slots
constexpr int FB_D = 128;

struct FbParams {
    int64_t BH, Sq, Skv;
    int H;
    const float *lse2;   // [BH, Sq] row log-sum-exp in the exp2 domain
    const float *delta;  // [BH, Sq] rowsum(dO o O)
    float *dq_acc;       // [BH, Tq * 64, D] fp32 running sums of dQ (unscaled)
    int *counters;       // [BH, Tq] contributions already folded into dq_acc, zeroed before the launch
    void *dq, *dk, *dv;
    AttnLayout ldq, ldk, ldv;
    float scale_log2;    // softmax scale * log2(e)
    float scale;         // softmax scale (applied to dK / dQ at the end)
    int nblk;            // 128-row KV blocks per (b, h)
    int Tq;              // 64-row query tiles per (b, h)
    int is_bf16;
};

template <bool BF16>
__device__ __forceinline__ uint32_t fb_pack(float2 v) {
    if (BF16) {
        __nv_bfloat162 h = __float22bfloat162_rn(v);
        return *reinterpret_cast<uint32_t *>(&h);
    }
    __half2 h = __float22half2_rn(v);
    return *reinterpret_cast<uint32_t *>(&h);
}
__device__ __forceinline__ float fb_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void fb_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ int fb_ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fb_st_release(int *p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void fb_sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// Element-wise phase of one streamed 64-query tile for one thread (= one kv row):
//   P^T = exp2(S^T c - lse2_q) -> 16-bit over TMEM columns [0, 32);  dS^T = P^T o (dP^T - delta_q) -> 16-bit over [96, 128) AND into
//   this row of the shared-memory dS^T tile (16-byte pieces XOR-swizzled by row & 7: the 128B-swizzle layout the MMA reads).
// dP^T is loaded whole first (dS^T lands on its upper half), S^T in 16-column chunks one ahead.
// vb: [32 column pairs][-lse2(2c), -lse2(2c+1), -delta(2c), -delta(2c+1)]
template <bool MASKED, bool BF16>
__device__ __forceinline__ void fb_ew_tile(const uint32_t t_addr, const uint32_t vb_smem, const uint32_t ds_row, const uint32_t rsw, const float sc,
                                           const int lo, const int hi) {
    const float2 sc2 = make_float2(sc, sc);
    uint32_t dp[4][16], s_r[2][16];
#pragma unroll
    for (int c = 0; c < 4; ++c) tmem_ld16(t_addr + 64 + (uint32_t)(c * 16), dp[c]);
    tmem_ld16(t_addr, s_r[0]);
    tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int cur = c & 1;
        if (c < 3) tmem_ld16(t_addr + (uint32_t)((c + 1) * 16), s_r[cur ^ 1]);
        uint32_t pk[8], dk[8];
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(vb_smem + 8u * (c * 16 + i)));
            const float2 x = __ffma2_rn(make_float2(__uint_as_float(s_r[cur][i]), __uint_as_float(s_r[cur][i + 1])), sc2, make_float2(v.x, v.y));
            float2 pr = make_float2(fb_ex2(x.x), fb_ex2(x.y));
            if (MASKED) {
                const int col = c * 16 + i;
                if (col < lo || col >= hi) pr.x = 0.f;
                if (col + 1 < lo || col + 1 >= hi) pr.y = 0.f;
            }
            const float2 ds = __fmul2_rn(pr, __fadd2_rn(make_float2(__uint_as_float(dp[c][i]), __uint_as_float(dp[c][i + 1])), make_float2(v.z, v.w)));
            pk[i >> 1] = fb_pack<BF16>(pr);
            dk[i >> 1] = fb_pack<BF16>(ds);
        }
        tmem_st8(t_addr + (uint32_t)(c * 8), pk);        // P^T over S^T columns already consumed
        tmem_st8(t_addr + 96 + (uint32_t)(c * 8), dk);   // dS^T over the upper half of dP^T (all of dP^T is in registers)
        fb_sts128(ds_row + (((uint32_t)(2 * c) ^ rsw) << 4), dk[0], dk[1], dk[2], dk[3]);
        fb_sts128(ds_row + (((uint32_t)(2 * c + 1) ^ rsw) << 4), dk[4], dk[5], dk[6], dk[7]);
        if (c < 3) tmem_ld_wait();
    }
}

// rowsum(dO o O) and the exp2-domain LSE for strided [B, H, S, D] operands.  16 lanes per query row (D = 128, 16-byte loads).
template <typename T>
__global__ void __launch_bounds__(256) fb_prep_kernel(const T *__restrict__ o, const T *__restrict__ dout, const float *__restrict__ lse,
                                                      float *__restrict__ delta, float *__restrict__ lse2, const int64_t rows, const int64_t Sq,
                                                      const int H, const AttnLayout lo, const AttnLayout ldo) {
    const int64_t row = (int64_t)blockIdx.x * 16 + threadIdx.x / 16;
    const int sub = threadIdx.x % 16;
    float acc = 0.f;
    if (row < rows) {
        const int64_t bh = row / Sq, s = row % Sq, b = bh / H, h = bh % H;
        const uint4 vo = __ldg(reinterpret_cast<const uint4 *>(o + b * lo.sb + h * lo.sh + s * lo.ss) + sub);
        const uint4 vd = __ldg(reinterpret_cast<const uint4 *>(dout + b * ldo.sb + h * ldo.sh + s * ldo.ss) + sub);
        const T *po = reinterpret_cast<const T *>(&vo), *pd = reinterpret_cast<const T *>(&vd);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc = fmaf(cvt_in<float>(po[i]), cvt_in<float>(pd[i]), acc);
    }
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (row < rows && sub == 0) {
        delta[row] = acc;
        lse2[row] = lse[row] * 1.4426950408889634f;
    }
}

__global__ void __launch_bounds__(FB_THREADS, 1)
attn_bwd_fused_kernel(const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                      const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_do, const FbParams p) {
    constexpr int D = FB_D;
    constexpr int ATOMS = D / 64;
    constexpr int X_BYTES = 128 * D * 2, X_ATOM = 128 * 128;  // stationary K_j, V_j: 128 rows
    constexpr int Y_BYTES = 64 * D * 2, Y_ATOM = 64 * 128;    // streamed Q_t, dO_t: 64 rows
    constexpr int SLOT_BYTES = 2 * Y_BYTES;
    constexpr int DS_BYTES = 128 * 128;                       // dS^T tile: 128 kv rows x 64 queries x 2 B
    constexpr int NS = FB_NS;
    constexpr uint32_t TMEM_COLS = 512;
    constexpr uint32_t DV_COL = 256, DK_COL = 256 + D;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *sX = smem;                                  // K_j | V_j
    unsigned char *sY = smem + 2 * X_BYTES;                    // NS slots of Q_t | dO_t
    unsigned char *sDS = sY + NS * SLOT_BYTES;                 // [2 sets] dS^T
    float *svec = reinterpret_cast<float *>(sDS + 2 * DS_BYTES);  // [2 sets][2 bufs][128]
    uint64_t *bars = reinterpret_cast<uint64_t *>(svec + 2 * 2 * 128);
    uint64_t *x_full = bars;
    uint64_t *y_full = bars + 1, *y_empty = bars + 1 + NS;
    uint64_t *t_full = bars + 1 + 2 * NS;  // [2]
    uint64_t *p_full = t_full + 2;         // [2]
    uint64_t *dq_full = p_full + 2;        // [2]
    uint64_t *dq_free = dq_full + 2;       // [2]
    uint64_t *acc_full = dq_free + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bh = blockIdx.x / p.nblk;
    const int j = p.nblk - 1 - (blockIdx.x % p.nblk);  // descending inside a (b, h): a CTA waits only for CTAs dispatched before it
    const int b_idx = bh / p.H, h_idx = bh % p.H;
    const int x0_row = j * 128;
    const int t_lo = x0_row / 64;
    const int t_hi = (int)((p.Sq + 63) / 64);
    const int ntile = max(t_hi - t_lo, 0);

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
        printf("kfunca_b200: attn_bwd_fused_kernel needs 1024-byte aligned dynamic shared memory\n");
        __trap();
    }
    if (warp == 9 && lane == 0) {
        prefetch_tmap(&tmap_k);
        prefetch_tmap(&tmap_v);
        prefetch_tmap(&tmap_q);
        prefetch_tmap(&tmap_do);
        mbar_init(x_full, 1);
        for (int s = 0; s < NS; ++s) {
            mbar_init(&y_full[s], 1);
            mbar_init(&y_empty[s], 1);
        }
        for (int t = 0; t < 2; ++t) {
            mbar_init(&t_full[t], 1);
            mbar_init(&p_full[t], 4);
            mbar_init(&dq_full[t], 1);
            mbar_init(&dq_free[t], 4);
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
                tma_load_4d(sX + a * X_ATOM, &tmap_k, x_full, a * 64, x0_row, h_idx, b_idx);
                tma_load_4d(sX + X_BYTES + a * X_ATOM, &tmap_v, x_full, a * 64, x0_row, h_idx, b_idx);
            }
            int s = 0;
            uint32_t ph = 0;
            for (int t = t_lo; t < t_hi; ++t) {
                mbar_wait(&y_empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&y_full[s], SLOT_BYTES);
#pragma unroll
                for (int a = 0; a < ATOMS; ++a) {
                    tma_load_4d(sY + s * SLOT_BYTES + a * Y_ATOM, &tmap_q, &y_full[s], a * 64, t * 64, h_idx, b_idx);
                    tma_load_4d(sY + s * SLOT_BYTES + Y_BYTES + a * Y_ATOM, &tmap_do, &y_full[s], a * 64, t * 64, h_idx, b_idx);
                }
                if (++s == NS) {
                    s = 0;
                    ph ^= 1;
                }
            }
        }
        __syncwarp();
    } else if (warp == 8) {
        // ===================================================== MMA issuer (converged warp, elected lane issues)
        const bool leader = elect_one();
        if (ntile > 0) {
            const int fmt = p.is_bf16 ? 1 : 0;
            const uint32_t idesc_t = make_idesc_f16(fmt, 0, 0, 128, 64);   // T = X Y^T: both operands K-major
            const uint32_t idesc_a = make_idesc_f16(fmt, 0, 1, 128, D);    // ACC += (TMEM) Y: B MN-major
            const uint32_t idesc_q = make_idesc_f16(fmt, 1, 1, 128, 64);   // dQ^T = K^T dS^T: A and B MN-major
            const uint32_t x_addr = smem_u32(sX), y_addr = smem_u32(sY), ds_addr = smem_u32(sDS);
            auto issue_t = [&](int set, int slot) {  // S^T(set) = K Q^T, dP^T(set) = V dO^T
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
            auto issue_dq = [&](int set) {  // dQ^T(set) = K_j^T dS^T: M = d (2 atoms of 64, X_ATOM apart), K = 128 kv rows (16 per step), N = 64 queries
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)
                    umma_f16_p(tmem_base + (uint32_t)(set * 128 + 32), make_sw128_desc(x_addr + (uint32_t)(kk * 2048), X_ATOM, 1024),
                               make_sw128_desc(ds_addr + (uint32_t)(set * DS_BYTES + kk * 2048), 0, 1024), idesc_q, kk ? 1u : 0u, leader);
            };
            auto issue_acc = [&](int set, int slot, bool accumulate) {
                const uint32_t ya = y_addr + slot * SLOT_BYTES;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {  // K = 64 streamed rows = 4 x 16; 16-bit A in TMEM: 8 columns per step
                    umma_f16_ts_p(tmem_base + DV_COL, tmem_base + (uint32_t)(set * 128 + kk * 8), make_sw128_desc(ya + Y_BYTES + kk * 2048, Y_ATOM, 1024),
                                  idesc_a, (accumulate || kk) ? 1u : 0u, leader);  // dV += P^T dO_t
                    umma_f16_ts_p(tmem_base + DK_COL, tmem_base + (uint32_t)(set * 128 + 96 + kk * 8), make_sw128_desc(ya + kk * 2048, Y_ATOM, 1024),
                                  idesc_a, (accumulate || kk) ? 1u : 0u, leader);  // dK += dS^T Q_t
                }
            };
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
                const uint32_t itp = (uint32_t)((n >> 1) & 1);
                mbar_wait(&p_full[set], itp);
                tc_fence_after();
                // dQ^T first: its drain by the element-wise warps then runs under the dV / dK MMAs instead of stalling the next S^T
                issue_dq(set);
                umma_commit_p(&dq_full[set], leader);
                issue_acc(set, n % NS, n > 0);
                umma_commit_p(&y_empty[n % NS], leader);
                if (n + 2 < ntile) {
                    const int slot = wait_slot(n + 2);
                    mbar_wait(&dq_free[set], itp);  // dQ^T(set) is in registers: its columns may be overwritten by the next S^T / dP^T
                    tc_fence_after();
                    issue_t(set, slot);
                    umma_commit_p(&t_full[set], leader);
                }
            }
            umma_commit_p(acc_full, leader);
        }
        __syncwarp();
    } else {
        // ===================================================== element-wise math, dQ hand-over, epilogue: thread = kv row
        const int set = warp >> 2, q4 = warp & 3;
        const int r = q4 * 32 + lane;
        const int64_t row_g = (int64_t)x0_row + r;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q4 * 32) << 16);
        const uint32_t t_addr = lane_addr + (uint32_t)(set * 128);
        const float sc = p.scale_log2;
        const float *lse2 = p.lse2 + (int64_t)bh * p.Sq, *delta = p.delta + (int64_t)bh * p.Sq;
        float *vec = svec + set * 256;  // [2 bufs][32 column pairs][-lse2 x2, -delta x2]
        const int tsel = threadIdx.x & 127;
        const int vslot = 4 * ((tsel & 63) >> 1) + (tsel & 1) + (tsel < 64 ? 0 : 2);
        const uint32_t ds_row = smem_u32(sDS) + (uint32_t)(set * DS_BYTES + r * 128);
        const uint32_t rsw = (uint32_t)(r & 7);
        auto load_vec = [&](int n) {
            float v = 0.f;
            if (n < ntile) {
                const int64_t qg = (int64_t)(t_lo + n) * 64 + (tsel & 63);
                if (qg < p.Sq) v = -(tsel < 64 ? lse2 : delta)[qg];
            }
            return v;
        };
        float vnext = load_vec(set);
        uint16_t *dq_base = reinterpret_cast<uint16_t *>(p.dq) + (int64_t)b_idx * p.ldq.sb + (int64_t)h_idx * p.ldq.sh + r;  // + q * ss (r = d here)
        for (int n = set; n < ntile; n += 2) {
            const int t = t_lo + n;
            const int64_t y0_row = (int64_t)t * 64;
            const int it = n >> 1;
            float *vb = vec + (it & 1) * 128;
            vb[vslot] = vnext;
            fb_bar(1 + set, 128);
            vnext = load_vec(n + 2);
            // columns c (query index inside the tile) kept iff lo <= c < hi: q >= kv (causal, top-left aligned), q < Sq
            const int lo = (int)max((int64_t)0, min((int64_t)64, row_g - y0_row));
            const int hi = (int)max((int64_t)0, min((int64_t)64, p.Sq - y0_row));
            const bool need_mask = __any_sync(0xffffffffu, lo > 0 || hi < 64);
            mbar_wait(&t_full[set], (uint32_t)(it & 1));
            tc_fence_after();
            const uint32_t vbs = smem_u32(vb);
            if (p.is_bf16) {
                if (need_mask) fb_ew_tile<true, true>(t_addr, vbs, ds_row, rsw, sc, lo, hi);
                else fb_ew_tile<false, true>(t_addr, vbs, ds_row, rsw, sc, lo, hi);
            } else {
                if (need_mask) fb_ew_tile<true, false>(t_addr, vbs, ds_row, rsw, sc, lo, hi);
                else fb_ew_tile<false, false>(t_addr, vbs, ds_row, rsw, sc, lo, hi);
            }
            tmem_st_wait();
            fence_proxy_async();  // the dS^T tile was written with ordinary stores; the MMA reads it through the async proxy
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[set]);

            // ---- this block's share of dQ_t: ordered accumulation, contributors j = t/2, t/2 - 1, ..., 0 (thread = column d = r)
            const int jmax = min(p.nblk - 1, t >> 1);
            const int my_turn = jmax - j;
            const bool first = my_turn == 0, last = j == 0;
            int *ctr = p.counters + (int64_t)bh * p.Tq + t;
            float *acc_tile = p.dq_acc + ((int64_t)bh * p.Tq + t) * 64 * D + r;
            float prev[64];
            if (!first) {
                if (tsel == 0) {
                    const long long t0 = clock64();
                    while (fb_ld_acquire(ctr) < my_turn) {
                        __nanosleep(64);
                        if (clock64() - t0 > 8000000000ll) {
                            printf("kfunca_b200: attention backward dQ hand-over timed out (bh %d kv block %d tile %d turn %d)\n", bh, j, t, my_turn);
                            __trap();
                        }
                    }
                }
                fb_bar(3 + set, 128);
#pragma unroll
                for (int c = 0; c < 64; ++c) prev[c] = __ldcg(acc_tile + c * D);  // in flight while the dQ^T MMAs finish
            } else {
#pragma unroll
                for (int c = 0; c < 64; ++c) prev[c] = 0.f;
            }
            mbar_wait(&dq_full[set], (uint32_t)(it & 1));
            tc_fence_after();
            uint32_t mine[4][16];
#pragma unroll
            for (int c = 0; c < 4; ++c) tmem_ld16(t_addr + 32 + (uint32_t)(c * 16), mine[c]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&dq_free[set]);
            if (last) {
                const float mul = p.scale;
#pragma unroll
                for (int c = 0; c < 64; ++c) {
                    const int64_t qg = y0_row + c;
                    if (qg < p.Sq) {
                        const float v = (prev[c] + __uint_as_float(mine[c >> 4][c & 15])) * mul;
                        dq_base[qg * p.ldq.ss] = p.is_bf16 ? __bfloat16_as_ushort(__float2bfloat16_rn(v)) : __half_as_ushort(__float2half_rn(v));
                    }
                }
            } else {
#pragma unroll
                for (int c = 0; c < 64; ++c) __stcg(acc_tile + c * D, prev[c] + __uint_as_float(mine[c >> 4][c & 15]));
                __threadfence();
                fb_bar(3 + set, 128);
                if (tsel == 0) fb_st_release(ctr, my_turn + 1);
            }
        }
        // ---- epilogue: dV (set 0 warps) and dK (set 1 warps), thread = kv row
        const bool is_bf16 = p.is_bf16;
        auto store_acc = [&](uint32_t col0, void *outp, const AttnLayout &l, float mul) {
            const bool row_ok = row_g < p.Skv;
            uint16_t *orow = reinterpret_cast<uint16_t *>(outp) + (int64_t)b_idx * l.sb + (int64_t)h_idx * l.sh + (row_ok ? row_g : 0) * l.ss;
#pragma unroll 1
            for (int c = 0; c < D; c += 32) {
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
                        for (int k = 0; k < 4; ++k) {
                            const float2 f = make_float2(__uint_as_float(a[8 * i + 2 * k]) * mul, __uint_as_float(a[8 * i + 2 * k + 1]) * mul);
                            wv[k] = is_bf16 ? fb_pack<true>(f) : fb_pack<false>(f);
                        }
                        dst[i] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
                    }
                }
            }
        };
        if (ntile > 0) {
            mbar_wait(acc_full, 0);
            tc_fence_after();
        }
        if (set == 0) store_acc(DV_COL, p.dv, p.ldv, 1.f);
        else store_acc(DK_COL, p.dk, p.ldk, p.scale);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace

bool launch_attention_bwd_fused(const AttnBwdPlan &a) {
    static const bool off = std::getenv("KF_ATTN_BWD_TWO_KERNEL") != nullptr;  // the round-1 two-kernel scheme, kept for A/B runs
    if (off) return false;
    if (a.dtype != KF_HALF && a.dtype != KF_BFLOAT16) return false;
    if (a.D != FB_D) return false;
    if (a.Sq < 1 || a.Skv < 1 || a.BH < 1) return false;
    const int64_t H = a.H > 0 ? a.H : a.BH;
    if (a.BH % H != 0 || a.BH / H >= 65536 || H >= 65536) return false;
    const bool dense = a.H <= 0;
    const AttnLayout dq_l = dense ? AttnLayout{H * a.Sq * a.D, a.Sq * a.D, a.D} : a.ldq, dkv_d = AttnLayout{H * a.Skv * a.D, a.Skv * a.D, a.D};
    const AttnLayout lq = dense ? dq_l : a.lq, lk = dense ? dkv_d : a.lk, lv = dense ? dkv_d : a.lv, lo = dense ? dq_l : a.lo,
                     ldo = dense ? dq_l : a.ldo, ldk = dense ? dkv_d : a.ldk, ldv = dense ? dkv_d : a.ldv;
    auto ok = [](const void *p, const AttnLayout &l) {
        return reinterpret_cast<uintptr_t>(p) % 16 == 0 && l.sb % 8 == 0 && l.sh % 8 == 0 && l.ss % 8 == 0 && l.ss >= FB_D;
    };
    if (!ok(a.q, lq) || !ok(a.k, lk) || !ok(a.v, lv) || !ok(a.out, lo) || !ok(a.dout, ldo) || !ok(a.dq, dq_l) || !ok(a.dk, ldk) || !ok(a.dv, ldv))
        return false;
    Runtime &rt = Runtime::get();
    const bool bf16 = a.dtype == KF_BFLOAT16;
    const int64_t B = a.BH / H;
    const int64_t rows = a.BH * a.Sq;
    const int Tq = (int)((a.Sq + 63) / 64);
    const int nblk = (int)((a.Skv + 127) / 128);
    if ((int64_t)a.BH * nblk >= (int64_t)0x7FFFFFFF) return false;
    Scratch delta((size_t)rows * 4), lse2((size_t)rows * 4), dq_acc((size_t)a.BH * Tq * 64 * FB_D * 4), counters((size_t)a.BH * Tq * 4);
    rt.memset_async(counters.p, 0, (size_t)a.BH * Tq * 4);
    const unsigned pgrid = (unsigned)((rows + 15) / 16);
    if (bf16)
        fb_prep_kernel<__nv_bfloat16><<<pgrid, 256, 0, rt.stream()>>>((const __nv_bfloat16 *)a.out, (const __nv_bfloat16 *)a.dout, (const float *)a.lse,
                                                                      delta.as<float>(), lse2.as<float>(), rows, a.Sq, (int)H, lo, ldo);
    else
        fb_prep_kernel<__half><<<pgrid, 256, 0, rt.stream()>>>((const __half *)a.out, (const __half *)a.dout, (const float *)a.lse, delta.as<float>(),
                                                               lse2.as<float>(), rows, a.Sq, (int)H, lo, ldo);
    rt.post_launch("attn_bwd_prep_kernel");
    auto map = [&](const void *ptr, int64_t S, const AttnLayout &l, uint32_t box_rows) {
        return make_tmap_4d_16bit(ptr, bf16, FB_D, (uint64_t)S, (uint64_t)H, (uint64_t)B, (uint64_t)l.ss, (uint64_t)l.sh, (uint64_t)l.sb, 64, box_rows);
    };
    const CUtensorMap tk = map(a.k, a.Skv, lk, 128), tv = map(a.v, a.Skv, lv, 128), tq = map(a.q, a.Sq, lq, 64), tdo = map(a.dout, a.Sq, ldo, 64);
    FbParams p{};
    p.BH = a.BH; p.Sq = a.Sq; p.Skv = a.Skv; p.H = (int)H;
    p.lse2 = lse2.as<float>(); p.delta = delta.as<float>();
    p.dq_acc = dq_acc.as<float>(); p.counters = counters.as<int>();
    p.dq = a.dq; p.dk = a.dk; p.dv = a.dv;
    p.ldq = dq_l; p.ldk = ldk; p.ldv = ldv;
    const double scale = 1.0 / std::sqrt((double)FB_D);
    p.scale = (float)scale;
    p.scale_log2 = (float)(scale * 1.4426950408889634);
    p.nblk = nblk; p.Tq = Tq; p.is_bf16 = bf16;
    constexpr int SMEM = 2 * 128 * FB_D * 2 + FB_NS * 2 * 64 * FB_D * 2 + 2 * 128 * 128 + 2 * 2 * 128 * 4 + 256;
    static bool attr_done = false;
    if (!attr_done) {
        KF_CUDA(cudaFuncSetAttribute(attn_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        attr_done = true;
    }
    attn_bwd_fused_kernel<<<(unsigned)(a.BH * nblk), FB_THREADS, SMEM, rt.stream()>>>(tk, tv, tq, tdo, p);
    rt.post_launch("attn_bwd_fused_kernel");
    return true;
}

}  // namespace kf

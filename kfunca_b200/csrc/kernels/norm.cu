// Fused layer normalisation over the last dimension, forward and backward (the reference stops at the
// statistics — `mean_var` / `norm_stat`, src/device/reduce_ops_kernel.cu:61-153, src/device/norm_ops_kernel.cu:6-61,
// src/device/utils/welford_norm.h:25-355 — and lists the fused norm as the next unchecked op, README.md:28).
//
//   y = (x - mean) * rstd * gain,   mean = E[x],  rstd = 1 / sqrt(E[(x - mean)^2] + eps)        (biased variance, per row)
//
// HBM-bound: forward reads x once and writes y once (+ 8 B of statistics per row); backward reads x and dy once, writes dx
// once.  The composed form the transformer block used before (mean, sub, mul, mean, add, rsqrt, mul, mul) moves ~10x the bytes
// forward and ~25x backward.
//
// One row is held in registers by 128 threads (NV 16-byte vectors each), so the statistics are the exact two-pass ones
// (mean first, then the centred second moment) at single-read cost — no Welford merge and no E[x^2] - E[x]^2 cancellation.
//   forward : one row per CTA.
//   backward: dx with one row per CTA; the gain gradient in a second launch of persistent CTAs that stride over the rows, each
//             thread keeping the partial sums of ITS columns in registers, written once to a [CTAs, E] fp32 scratch that the
//             ordinary column reduction folds afterwards: deterministic, no atomics.
#include <cooperative_groups.h>

#include <cstdlib>
#include <cstring>

#include "ew_common.cuh"
#include "tc_common.cuh"

namespace kf {
namespace cg = cooperative_groups;

constexpr int LN_THREADS = 128;

__device__ __forceinline__ float ln_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// sum over the TH (128 or 256) threads of the CTA, result in every thread; `slot` = TH / 32 floats of shared memory not in use by another sum
template <int TH = LN_THREADS>
__device__ __forceinline__ float ln_block_sum(float v, float *slot) {
    v = ln_warp_sum(v);
    if ((threadIdx.x & 31) == 0) slot[threadIdx.x >> 5] = v;
    __syncthreads();
    if (TH == 256) return ((slot[0] + slot[1]) + (slot[2] + slot[3])) + ((slot[4] + slot[5]) + (slot[6] + slot[7]));
    return (slot[0] + slot[1]) + (slot[2] + slot[3]);
}

template <typename T, int VEC>
__device__ __forceinline__ Pack<T, VEC> ln_ld(const T *p) {
    return *reinterpret_cast<const Pack<T, VEC> *>(p);
}

struct LnArgs {
    const void *x, *gain, *dy;
    void *y, *dx;
    float *mean, *rstd;      // [rows]
    float *dgain_partial;    // [gridDim.x, E] (backward)
    int64_t rows, E;
    float eps;
    int rms;                 // 1 = RMSNorm (README.md:28 `rms_norm`): no centring, rstd = 1 / sqrt(mean(x^2) + eps), mean stored as 0
};

// TH threads per row: 128, or 256 for rows that would need 8 vectors per thread at 128 (E = 4096 fp32: 55 -> 32 registers, 47 % -> full
// occupancy, more loads in flight per SM)
template <typename T, int VEC, int NV, int TH>
__global__ void __launch_bounds__(TH) layer_norm_fwd_kernel(const LnArgs a) {
    pdl_enter();
    __shared__ float red[16];
    const int64_t row = blockIdx.x;
    const int nvec = (int)(a.E / VEC);
    const T *__restrict__ x = reinterpret_cast<const T *>(a.x) + row * a.E;
    float v[NV][VEC];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int iv = threadIdx.x + k * TH;
        if (iv < nvec) {
            const Pack<T, VEC> pk = ln_ld<T, VEC>(x + (int64_t)iv * VEC);
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                v[k][i] = cvt_in<float>(pk.v[i]);
                s += v[k][i];
            }
        }
    }
    const float mean = a.rms ? 0.f : ln_block_sum<TH>(s, red) / (float)a.E;  // a.rms is uniform over the grid: no divergent barrier
    float d = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int iv = threadIdx.x + k * TH;
        if (iv < nvec) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                v[k][i] -= mean;
                d += v[k][i] * v[k][i];
            }
        }
    }
    const float rstd = rsqrtf(ln_block_sum<TH>(d, red + 8) / (float)a.E + a.eps);
    if (threadIdx.x == 0) {
        a.mean[row] = mean;
        a.rstd[row] = rstd;
    }
    const T *__restrict__ g = reinterpret_cast<const T *>(a.gain);
    T *__restrict__ y = reinterpret_cast<T *>(a.y) + row * a.E;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int iv = threadIdx.x + k * TH;
        if (iv < nvec) {
            const Pack<T, VEC> gk = ln_ld<T, VEC>(g + (int64_t)iv * VEC);
            Pack<T, VEC> out;
#pragma unroll
            for (int i = 0; i < VEC; ++i) out.v[i] = cvt_out<T, float>(v[k][i] * rstd * cvt_in<float>(gk.v[i]));
            *reinterpret_cast<Pack<T, VEC> *>(y + (int64_t)iv * VEC) = out;
        }
    }
}

// Row statistics only (mean_var over the last dimension, ref: gpu::mean_var, src/core/reduce_ops.cpp:22-28 — Welford with
// correction 1): same register-resident two-pass scheme, one HBM read; out_var = M2 / (E - 1), optionally its square root.
template <typename T, int VEC, int NV, int TH>
__global__ void __launch_bounds__(TH) row_moments_kernel(const T *__restrict__ xin, T *__restrict__ out_mean, T *__restrict__ out_var,
                                                                 const int64_t E, const int take_sqrt) {
    pdl_enter();
    __shared__ float red[16];
    const int64_t row = blockIdx.x;
    const int nvec = (int)(E / VEC);
    const T *__restrict__ x = xin + row * E;
    float v[NV][VEC];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int iv = threadIdx.x + k * TH;
        if (iv < nvec) {
            const Pack<T, VEC> pk = ln_ld<T, VEC>(x + (int64_t)iv * VEC);
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                v[k][i] = cvt_in<float>(pk.v[i]);
                s += v[k][i];
            }
        }
    }
    const float mean = ln_block_sum<TH>(s, red) / (float)E;
    float d = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int iv = threadIdx.x + k * TH;
        if (iv < nvec) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) d += (v[k][i] - mean) * (v[k][i] - mean);
        }
    }
    const float m2 = ln_block_sum<TH>(d, red + 8);
    if (threadIdx.x == 0) {
        const float div = (float)E - 1.f;
        float var = m2 / (div > 0.f ? div : 0.f);
        if (take_sqrt) var = sqrtf(var);
        out_mean[row] = cvt_out<T, float>(mean);
        out_var[row] = cvt_out<T, float>(var);
    }
}

// dx = rstd * (gg - mean_e(gg) - xhat * mean_e(gg * xhat)),  gg = dy * gain,  xhat = (x - mean) * rstd
// dgain[e] = sum_rows dy * xhat  (per-CTA partial rows here, folded by the caller)
// Two launches: <WRITE_DX, !DGAIN> with one row per CTA (no state carried between rows, so occupancy and bytes in flight are
// those of the forward kernel), then <!WRITE_DX, DGAIN> as persistent CTAs whose row loop has no barrier and no store, so the
// loads of consecutive rows pipeline freely.  Doing both in one persistent kernel (first version) held 4 register arrays per
// thread and serialised load -> barrier -> store per row: 303 us at [32768, 4096] bf16 against 97 us for the DGAIN-only form.
template <typename T, int VEC, int NV, bool WRITE_DX, bool DGAIN>
__global__ void __launch_bounds__(LN_THREADS) layer_norm_bwd_kernel(const LnArgs a) {
    __shared__ float red[2][8];
    const int nvec = (int)(a.E / VEC);
    const T *__restrict__ gp = reinterpret_cast<const T *>(a.gain);
    float gain[NV][VEC], dgain[NV][VEC];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int iv = threadIdx.x + k * LN_THREADS;
        Pack<T, VEC> gk{};
        if (iv < nvec) gk = ln_ld<T, VEC>(gp + (int64_t)iv * VEC);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            gain[k][i] = iv < nvec ? cvt_in<float>(gk.v[i]) : 0.f;
            dgain[k][i] = 0.f;
        }
    }
    int it = 0;
    for (int64_t row = blockIdx.x; row < a.rows; row += gridDim.x, ++it) {
        const T *__restrict__ x = reinterpret_cast<const T *>(a.x) + row * a.E;
        const T *__restrict__ dy = reinterpret_cast<const T *>(a.dy) + row * a.E;
        const float mean = a.mean[row], rstd = a.rstd[row];
        float xh[NV][VEC], gg[NV][VEC];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int iv = threadIdx.x + k * LN_THREADS;
            if (iv < nvec) {
                const Pack<T, VEC> xk = ln_ld<T, VEC>(x + (int64_t)iv * VEC);
                const Pack<T, VEC> dk = ln_ld<T, VEC>(dy + (int64_t)iv * VEC);
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    const float g = cvt_in<float>(dk.v[i]);
                    xh[k][i] = (cvt_in<float>(xk.v[i]) - mean) * rstd;
                    gg[k][i] = g * gain[k][i];
                    if (DGAIN) dgain[k][i] += g * xh[k][i];
                    s1 += gg[k][i];
                    s2 += gg[k][i] * xh[k][i];
                }
            }
        }
        if (WRITE_DX) {
            // both row sums in one barrier round; the shared slots alternate between iterations so the next row's writers cannot
            // overtake this row's readers
            float *slot = red[it & 1];
            s1 = ln_warp_sum(s1);
            s2 = ln_warp_sum(s2);
            if ((threadIdx.x & 31) == 0) {
                slot[threadIdx.x >> 5] = s1;
                slot[4 + (threadIdx.x >> 5)] = s2;
            }
            __syncthreads();
            const float c1 = a.rms ? 0.f : ((slot[0] + slot[1]) + (slot[2] + slot[3])) / (float)a.E;
            const float c2 = ((slot[4] + slot[5]) + (slot[6] + slot[7])) / (float)a.E;
            T *__restrict__ dx = reinterpret_cast<T *>(a.dx) + row * a.E;
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                const int iv = threadIdx.x + k * LN_THREADS;
                if (iv < nvec) {
                    Pack<T, VEC> out;
#pragma unroll
                    for (int i = 0; i < VEC; ++i) out.v[i] = cvt_out<T, float>(rstd * (gg[k][i] - c1 - xh[k][i] * c2));
                    *reinterpret_cast<Pack<T, VEC> *>(dx + (int64_t)iv * VEC) = out;
                }
            }
        }
    }
    if (!DGAIN) return;
    float *__restrict__ part = a.dgain_partial + (int64_t)blockIdx.x * a.E;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int iv = threadIdx.x + k * LN_THREADS;
        if (iv < nvec) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) part[(int64_t)iv * VEC + i] = dgain[k][i];
        }
    }
}

// One-pass backward: dx AND the gain-gradient partials from a single read of x and dy.  Persistent CTAs (two per SM) stride over the
// rows; a thread keeps the gain-gradient sums of its columns in registers, and the NEXT row's x / dy vectors are already in flight
// (raw, in registers) while the current row goes through its one barrier and its store — the load -> barrier -> store serialisation
// that made the first one-kernel version slow (303 us at [32768, 4096] bf16) is gone.  Against the two-launch form this saves the
// second read of x and dy (2/5 of the bytes).
template <typename T, int VEC, int NV, int TH>
__global__ void __launch_bounds__(TH, 2) layer_norm_bwd_fused_kernel(const LnArgs a) {
    constexpr int NW = TH / 32;
    __shared__ float red[2][2 * NW];
    const int nvec = (int)(a.E / VEC);
    const T *__restrict__ gp = reinterpret_cast<const T *>(a.gain);
    float dgain[NV][VEC];
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
        for (int i = 0; i < VEC; ++i) dgain[k][i] = 0.f;
    Pack<T, VEC> xk[NV], dk[NV];
    float mean = 0.f, rstd = 0.f;
    auto fetch = [&](int64_t row, Pack<T, VEC>(&xo)[NV], Pack<T, VEC>(&dO)[NV], float &m, float &r) {
        const T *__restrict__ x = reinterpret_cast<const T *>(a.x) + row * a.E;
        const T *__restrict__ dy = reinterpret_cast<const T *>(a.dy) + row * a.E;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int iv = threadIdx.x + k * TH;
            if (iv < nvec) {
                xo[k] = ln_ld<T, VEC>(x + (int64_t)iv * VEC);
                dO[k] = ln_ld<T, VEC>(dy + (int64_t)iv * VEC);
            }
        }
        m = a.mean[row];
        r = a.rstd[row];
    };
    int64_t row = blockIdx.x;
    if (row < a.rows) fetch(row, xk, dk, mean, rstd);
    int it = 0;
    for (; row < a.rows; row += gridDim.x, ++it) {
        Pack<T, VEC> xn[NV], dn[NV];
        float mean_n = 0.f, rstd_n = 0.f;
        const int64_t next = row + gridDim.x;
        if (next < a.rows) fetch(next, xn, dn, mean_n, rstd_n);
        // pass 1 over the registers: row sums and the gain-gradient terms (x-hat and dy * gain are NOT kept: the raw vectors are, and
        // pass 2 recomputes them — four register arrays per thread would spill at E = 4096 fp32)
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int iv = threadIdx.x + k * TH;
            if (iv < nvec) {
                const Pack<T, VEC> gk = ln_ld<T, VEC>(gp + (int64_t)iv * VEC);  // 16 KB at most, L1-resident after the first row
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    const float g = cvt_in<float>(dk[k].v[i]);
                    const float xh = (cvt_in<float>(xk[k].v[i]) - mean) * rstd;
                    const float gg = g * cvt_in<float>(gk.v[i]);
                    dgain[k][i] += g * xh;
                    s1 += gg;
                    s2 += gg * xh;
                }
            }
        }
        float *slot = red[it & 1];
        s1 = ln_warp_sum(s1);
        s2 = ln_warp_sum(s2);
        if ((threadIdx.x & 31) == 0) {
            slot[threadIdx.x >> 5] = s1;
            slot[NW + (threadIdx.x >> 5)] = s2;
        }
        __syncthreads();
        float t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int w = 0; w < NW; w += 4) {  // fixed order: deterministic
            t1 += (slot[w] + slot[w + 1]) + (slot[w + 2] + slot[w + 3]);
            t2 += (slot[NW + w] + slot[NW + w + 1]) + (slot[NW + w + 2] + slot[NW + w + 3]);
        }
        const float c1 = a.rms ? 0.f : t1 / (float)a.E;
        const float c2 = t2 / (float)a.E;
        T *__restrict__ dx = reinterpret_cast<T *>(a.dx) + row * a.E;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int iv = threadIdx.x + k * TH;
            if (iv < nvec) {
                const Pack<T, VEC> gk = ln_ld<T, VEC>(gp + (int64_t)iv * VEC);
                Pack<T, VEC> out;
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    const float xh = (cvt_in<float>(xk[k].v[i]) - mean) * rstd;
                    const float gg = cvt_in<float>(dk[k].v[i]) * cvt_in<float>(gk.v[i]);
                    out.v[i] = cvt_out<T, float>(rstd * (gg - c1 - xh * c2));
                }
                *reinterpret_cast<Pack<T, VEC> *>(dx + (int64_t)iv * VEC) = out;
            }
        }
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            xk[k] = xn[k];
            dk[k] = dn[k];
        }
        mean = mean_n;
        rstd = rstd_n;
    }
    float *__restrict__ part = a.dgain_partial + (int64_t)blockIdx.x * a.E;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int iv = threadIdx.x + k * TH;
        if (iv < nvec) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) part[(int64_t)iv * VEC + i] = dgain[k][i];
        }
    }
}

// One-pass backward, rows streamed through a shared-memory ring by bulk copies (the default).  The register-prefetch kernel above keeps
// ONE row per CTA in flight (two CTAs per SM: 32 KB per SM for bf16 rows of 4096 — 0.50 of the HBM peak at [32768, 4096]); here an
// elected thread keeps RING_STAGES rows of x and dy per CTA on their way with cp.async.bulk (one mbarrier per stage, complete_tx),
// 96 KB per CTA and two CTAs per SM whatever the dtype, and the 256 threads only ever touch shared memory and registers:
//   wait full[stage] -> x, dy of the row into registers (converted once) -> row sums (shuffle + one __syncthreads) -> the stage is
//   free: thread 0 re-arms it with the row RING_STAGES ahead -> dx from the registers, 16-byte coalesced stores.
// The gain and the gain-gradient sums of a thread's columns live in registers for the whole kernel.
template <typename T, int VEC, int NV>
__global__ void __launch_bounds__(256, 2) layer_norm_bwd_ring_kernel(const LnArgs a, const int stages) {
    pdl_enter();
    constexpr int TH = 256, NW = TH / 32;
    extern __shared__ __align__(128) unsigned char ring_raw[];
    __shared__ float red[2][2 * NW];
    __shared__ uint64_t full[8];
    const int64_t row_bytes = a.E * (int64_t)sizeof(T);
    const int nvec = (int)(a.E / VEC);
    auto stage_x = [&](int st) { return reinterpret_cast<const T *>(ring_raw + (int64_t)st * 2 * row_bytes); };
    auto stage_dy = [&](int st) { return reinterpret_cast<const T *>(ring_raw + (int64_t)st * 2 * row_bytes + row_bytes); };
    auto arm = [&](int st, int64_t row) {  // thread 0 only
        tc::mbar_arrive_expect_tx(&full[st], (uint32_t)(2 * row_bytes));
        const uint32_t bar = tc::smem_u32(&full[st]);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(tc::smem_u32(stage_x(st))), "l"(reinterpret_cast<uint64_t>(reinterpret_cast<const T *>(a.x) + row * a.E)), "r"((uint32_t)row_bytes), "r"(bar)
                     : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(tc::smem_u32(stage_dy(st))), "l"(reinterpret_cast<uint64_t>(reinterpret_cast<const T *>(a.dy) + row * a.E)), "r"((uint32_t)row_bytes), "r"(bar)
                     : "memory");
    };
    if (threadIdx.x == 0) {
        for (int st = 0; st < stages; ++st) tc::mbar_init(&full[st], 1);
        tc::fence_barrier_init();
        for (int st = 0; st < stages; ++st) {
            const int64_t row = (int64_t)blockIdx.x + (int64_t)st * gridDim.x;
            if (row < a.rows) arm(st, row);
        }
    }
    float gain[NV][VEC], dgain[NV][VEC];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int iv = threadIdx.x + k * TH;
        Pack<T, VEC> gk{};
        if (iv < nvec) gk = ln_ld<T, VEC>(reinterpret_cast<const T *>(a.gain) + (int64_t)iv * VEC);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            gain[k][i] = iv < nvec ? cvt_in<float>(gk.v[i]) : 0.f;
            dgain[k][i] = 0.f;
        }
    }
    __syncthreads();  // barrier inits visible before the first wait
    int64_t row = blockIdx.x;
    float mean = 0.f, rstd = 0.f;
    if (row < a.rows) {
        mean = a.mean[row];
        rstd = a.rstd[row];
    }
    int st = 0;
    uint32_t ph = 0;
    for (int it = 0; row < a.rows; row += gridDim.x, ++it) {
        const int64_t next = row + gridDim.x;
        float mean_n = 0.f, rstd_n = 0.f;
        if (next < a.rows) {  // the next row's statistics travel under this row's arithmetic
            mean_n = a.mean[next];
            rstd_n = a.rstd[next];
        }
        tc::mbar_wait(&full[st], ph);
        const T *sx = stage_x(st), *sdy = stage_dy(st);
        float xh[NV][VEC], gg[NV][VEC];  // x-hat and dy * gain of this thread's columns
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int iv = threadIdx.x + k * TH;
            if (iv < nvec) {
                const Pack<T, VEC> xk = ln_ld<T, VEC>(sx + (int64_t)iv * VEC), dk = ln_ld<T, VEC>(sdy + (int64_t)iv * VEC);
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    const float g = cvt_in<float>(dk.v[i]);
                    xh[k][i] = (cvt_in<float>(xk.v[i]) - mean) * rstd;
                    gg[k][i] = g * gain[k][i];
                    dgain[k][i] = fmaf(g, xh[k][i], dgain[k][i]);
                    s1 += gg[k][i];
                    s2 = fmaf(gg[k][i], xh[k][i], s2);
                }
            }
        }
        float *slot = red[it & 1];
        s1 = ln_warp_sum(s1);
        s2 = ln_warp_sum(s2);
        if ((threadIdx.x & 31) == 0) {
            slot[threadIdx.x >> 5] = s1;
            slot[NW + (threadIdx.x >> 5)] = s2;
        }
        __syncthreads();  // every thread has read its part of the stage: it can be refilled
        if (threadIdx.x == 0) {
            const int64_t refill = row + (int64_t)stages * gridDim.x;
            if (refill < a.rows) arm(st, refill);
        }
        float t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int w = 0; w < NW; w += 4) {  // fixed order: deterministic
            t1 += (slot[w] + slot[w + 1]) + (slot[w + 2] + slot[w + 3]);
            t2 += (slot[NW + w] + slot[NW + w + 1]) + (slot[NW + w + 2] + slot[NW + w + 3]);
        }
        const float c1 = a.rms ? 0.f : t1 / (float)a.E;
        const float c2 = t2 / (float)a.E;
        T *__restrict__ dx = reinterpret_cast<T *>(a.dx) + row * a.E;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int iv = threadIdx.x + k * TH;
            if (iv < nvec) {
                Pack<T, VEC> out;
#pragma unroll
                for (int i = 0; i < VEC; ++i) out.v[i] = cvt_out<T, float>(rstd * (gg[k][i] - c1 - xh[k][i] * c2));
                *reinterpret_cast<Pack<T, VEC> *>(dx + (int64_t)iv * VEC) = out;
            }
        }
        mean = mean_n;
        rstd = rstd_n;
        if (++st == stages) {
            st = 0;
            ph ^= 1;
        }
    }
    float *__restrict__ part = a.dgain_partial + (int64_t)blockIdx.x * a.E;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int iv = threadIdx.x + k * TH;
        if (iv < nvec) {
#pragma unroll
            for (int i = 0; i < VEC; i += 4)
                *reinterpret_cast<float4 *>(part + (int64_t)iv * VEC + i) = make_float4(dgain[k][i], dgain[k][i + 1], dgain[k][i + 2], dgain[k][i + 3]);
        }
    }
}

template <typename T>
static int ln_nv(int64_t E) {
    constexpr int VEC = 16 / sizeof(T);
    const int64_t nvec = E / VEC;
    const int64_t per = (nvec + LN_THREADS - 1) / LN_THREADS;
    return per <= 1 ? 1 : per <= 2 ? 2 : per <= 4 ? 4 : per <= 8 ? 8 : 0;
}

bool layer_norm_supported(int dtype, int64_t E, const void *x, const void *gain) {
    if (dtype != KF_FLOAT && dtype != KF_HALF && dtype != KF_BFLOAT16) return false;
    const int vec = dtype == KF_FLOAT ? 4 : 8;
    if (E < vec || E % vec != 0) return false;
    if (reinterpret_cast<uintptr_t>(x) % 16 != 0 || reinterpret_cast<uintptr_t>(gain) % 16 != 0) return false;
    return (dtype == KF_FLOAT ? ln_nv<float>(E) : ln_nv<__half>(E)) != 0;
}

template <typename T>
static void ln_fwd_typed(const LnArgs &a) {
    constexpr int VEC = 16 / sizeof(T);
    Runtime &rt = Runtime::get();
    KF_CHECK(a.rows < (int64_t)0x7FFFFFFF);
    const unsigned grid = (unsigned)a.rows;
    switch (ln_nv<T>(a.E)) {
    case 1: launch_pdl(layer_norm_fwd_kernel<T, VEC, 1, LN_THREADS>, dim3(grid), dim3(LN_THREADS), 0, rt.stream(), a); break;
    case 2: launch_pdl(layer_norm_fwd_kernel<T, VEC, 2, LN_THREADS>, dim3(grid), dim3(LN_THREADS), 0, rt.stream(), a); break;
    case 4: launch_pdl(layer_norm_fwd_kernel<T, VEC, 4, LN_THREADS>, dim3(grid), dim3(LN_THREADS), 0, rt.stream(), a); break;
    default: launch_pdl(layer_norm_fwd_kernel<T, VEC, 4, 256>, dim3(grid), dim3(256), 0, rt.stream(), a); break;  // 8 vectors per thread at 128 = 4 at 256
    }
    rt.post_launch("layer_norm_fwd_kernel");
}

void launch_layer_norm_fwd(const void *x, const void *gain, void *y, float *mean, float *rstd, int dtype, int64_t rows, int64_t E, float eps,
                           bool rms) {
    if (rows == 0) return;
    LnArgs a{};
    a.x = x; a.gain = gain; a.y = y; a.mean = mean; a.rstd = rstd; a.rows = rows; a.E = E; a.eps = eps; a.rms = rms ? 1 : 0;
    if (dtype == KF_FLOAT) ln_fwd_typed<float>(a);
    else if (dtype == KF_HALF) ln_fwd_typed<__half>(a);
    else ln_fwd_typed<__nv_bfloat16>(a);
}

// fp32 rows only (the reference's mean_var is fp32 / fp64); false => caller composes the statistics from sum / mean
bool launch_row_moments(const void *x, void *mean, void *var, int dtype, int64_t rows, int64_t E, bool take_sqrt) {
    if (dtype != KF_FLOAT || E < 4 || E % 4 != 0 || reinterpret_cast<uintptr_t>(x) % 16 != 0 || ln_nv<float>(E) == 0) return false;
    if (rows == 0) return true;
    KF_CHECK(rows < (int64_t)0x7FFFFFFF);
    Runtime &rt = Runtime::get();
    const float *xp = reinterpret_cast<const float *>(x);
    float *mp = reinterpret_cast<float *>(mean), *vp = reinterpret_cast<float *>(var);
    const unsigned grid = (unsigned)rows;
    switch (ln_nv<float>(E)) {
    case 1: launch_pdl(row_moments_kernel<float, 4, 1, LN_THREADS>, dim3(grid), dim3(LN_THREADS), 0, rt.stream(), xp, mp, vp, E, take_sqrt); break;
    case 2: launch_pdl(row_moments_kernel<float, 4, 2, LN_THREADS>, dim3(grid), dim3(LN_THREADS), 0, rt.stream(), xp, mp, vp, E, take_sqrt); break;
    case 4: launch_pdl(row_moments_kernel<float, 4, 4, LN_THREADS>, dim3(grid), dim3(LN_THREADS), 0, rt.stream(), xp, mp, vp, E, take_sqrt); break;
    default: launch_pdl(row_moments_kernel<float, 4, 4, 256>, dim3(grid), dim3(256), 0, rt.stream(), xp, mp, vp, E, take_sqrt); break;
    }
    rt.post_launch("row_moments_kernel");
    return true;
}

// ------------------------------------------------------------------------------------------------------------------------
// Column statistics in ONE launch: x viewed as [outer, R, inner], statistics over R for every (outer, inner) column.
//   mode 0 (norm_stat, ref: norm_stat_kernel, src/device/norm_ops_kernel.cu:6-61 + WelfordNormPFKernel, welford_norm.h:25-355):
//           out0 = mean, out1 = 1 / sqrt(M2 / R + eps)            (biased variance)
//   mode 1 (mean_var over a non-last dim, ref: reduce_ops_kernel.cu:61-153, Welford with correction 1):
//           out0 = mean, out1 = M2 / (R - 1)   [sqrt when take_sqrt]
// A cluster of CM_C CTAs along R shares a 32 * VEC wide column strip; every thread runs Welford over its rows (one reciprocal per
// row shared by its VEC columns), the 8 warps merge through shared memory (Chan's pairwise update, fixed order), the CTAs merge
// through distributed shared memory in rank order: deterministic, no atomics, no staging buffer, no semaphore (the reference's
// un-zeroed semaphore block, SURVEY F10, has no counterpart).
constexpr int CM_C = 8;

struct ColMomentsArgs {
    const float *x;
    float *out0, *out1;
    int64_t outer, R, inner;
    float eps;
    int mode, take_sqrt;
};

__device__ __forceinline__ void chan_merge(float &na, float &ma, float &sa, const float nb, const float mb, const float sb) {
    const float n = na + nb;
    if (nb == 0.f) return;
    const float d = mb - ma, f = nb * __frcp_rn(n);  // counts are small integers: the product is within 1 ulp of nb / n, without the IEEE division sequence
    ma = fmaf(d, f, ma);
    sa = sa + sb + d * d * na * f;
    na = n;
}

// LPR lanes run along the contiguous dim (a warp covers 32 / LPR rows per load), WARPS warps per CTA; WARPS * 32 / LPR = 16 row
// slots per CTA in every geometry.  <4, 16, 8>: 64-column strips, 256 threads, twice as many (narrower) CTAs — the geometry of
// reduce_cols_kernel; <4, 32, 16> / <1, 32, 16>: 128- / 32-column strips, 512 threads.
template <int VEC, int LPR, int WARPS>
__global__ void __cluster_dims__(1, CM_C, 1) __launch_bounds__(WARPS * 32) col_moments_kernel(const ColMomentsArgs a) {
    pdl_enter();
    constexpr int RPW = 32 / LPR, SLOTS = WARPS * RPW, NCOL = LPR * VEC;
    static_assert(SLOTS == 16, "the two-level fold below merges 4 x 4 row slots");
    __shared__ float w_mean[SLOTS][NCOL], w_m2[SLOTS][NCOL], w_n[SLOTS];
    __shared__ float in_mean[CM_C - 1][NCOL], in_m2[CM_C - 1][NCOL], in_n[CM_C - 1];  // cluster rank 0: the peers' folded triples
    __shared__ uint64_t inbox_bar;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // rank 0 arms its inbox; the cluster barrier is only ARRIVED at here and waited for after the streaming loop (free by then): it
    // orders the mbarrier init before the first remote store (the push fold of reduce_cols_kernel: one DSMEM hop, no cluster.sync)
    if (tc::cluster_ctarank() == 0 && threadIdx.x == 0) {
        tc::mbar_init(&inbox_bar, (uint32_t)(CM_C - 1) * NCOL);
        tc::fence_barrier_init();
    }
    asm volatile("barrier.cluster.arrive.release;" ::: "memory");
    const int64_t o = blockIdx.z;
    const int cl = lane % LPR, slot = warp * RPW + lane / LPR;
    const int64_t col0 = ((int64_t)blockIdx.x * LPR + cl) * VEC;
    const bool col_ok = col0 < a.inner;  // VEC > 1 only when inner % VEC == 0
    const float *__restrict__ base = a.x + o * a.R * a.inner + col0;
    float n = 0.f, mean[VEC], m2[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) mean[i] = m2[i] = 0.f;
    const int64_t r0 = (int64_t)blockIdx.y * SLOTS + slot, rstep = (int64_t)CM_C * SLOTS;
    if (col_ok) {
        int64_t r = r0;
        // Batches of four rows, double-buffered: the NEXT batch's loads are issued before the current batch's arithmetic, so a thread
        // always has four to eight rows in flight (the first version loaded eight, waited, computed, and only then loaded again).
        // A batch's own mean and centred second moment are formed from registers (all independent), then folded into the running
        // triple with one Chan update — no per-row dependency chain, one reciprocal per batch.
        constexpr int NB = 4;
        float cur[NB][VEC], nxt[NB][VEC];
        auto load_batch = [&](float (&v)[NB][VEC], int64_t rr) {
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                if (VEC == 4) {
                    const float4 f = __ldcs(reinterpret_cast<const float4 *>(base + (rr + b * rstep) * a.inner));
                    v[b][0] = f.x; v[b][VEC > 1 ? 1 : 0] = f.y; v[b][VEC > 2 ? 2 : 0] = f.z; v[b][VEC > 3 ? 3 : 0] = f.w;
                } else {
                    v[b][0] = __ldcs(base + (rr + b * rstep) * a.inner);
                }
            }
        };
        bool have = r + (NB - 1) * rstep < a.R;
        if (have) load_batch(cur, r);
        while (have) {
            const int64_t rn = r + NB * rstep;
            const bool have_n = rn + (NB - 1) * rstep < a.R;
            if (have_n) load_batch(nxt, rn);
            const float nn = n + (float)NB, f = (float)NB * __frcp_rn(nn);
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const float bm = ((cur[0][i] + cur[1][i]) + (cur[2][i] + cur[3][i])) * (1.f / NB);
                float bs = 0.f;
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    const float d = cur[b][i] - bm;
                    bs = fmaf(d, d, bs);
                }
                const float d = bm - mean[i];
                mean[i] = fmaf(d, f, mean[i]);
                m2[i] = m2[i] + bs + d * d * n * f;
            }
            n = nn;
            r = rn;
            have = have_n;
#pragma unroll
            for (int b = 0; b < NB; ++b)
#pragma unroll
                for (int i = 0; i < VEC; ++i) cur[b][i] = nxt[b][i];
        }
        for (; r < a.R; r += rstep) {
            float v0[VEC];
            if (VEC == 4) {
                const float4 f0 = __ldcs(reinterpret_cast<const float4 *>(base + r * a.inner));
                v0[0] = f0.x; v0[VEC > 1 ? 1 : 0] = f0.y; v0[VEC > 2 ? 2 : 0] = f0.z; v0[VEC > 3 ? 3 : 0] = f0.w;
            } else {
                v0[0] = __ldcs(base + r * a.inner);
            }
            n += 1.f;
            const float inv = __frcp_rn(n);
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const float d = v0[i] - mean[i];
                mean[i] = fmaf(d, inv, mean[i]);
                m2[i] = fmaf(d, v0[i] - mean[i], m2[i]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        w_mean[slot][cl * VEC + i] = mean[i];
        w_m2[slot][cl * VEC + i] = m2[i];
    }
    // the row count of a slot does not depend on the column: lanes without a valid column report the count too (uniform merge)
    if (cl == 0) w_n[slot] = r0 < a.R ? (float)((a.R - r0 + rstep - 1) / rstep) : 0.f;
    __syncthreads();
    // Fold in two levels, ONE COLUMN PER THREAD (the first version let warp 0 walk all 15 other warps for its VEC columns per lane:
    // 60 dependent Chan updates, each with a division — 3.7 us of an 18 us kernel).  Level 1: thread (g, c) merges warps 4g .. 4g + 3
    // of column c (g = 0 .. 3: row slots 4g .. 4g + 3); level 2: threads of group 0 merge the group results.  Fixed order: deterministic.
    constexpr int NG = 4;
    static_assert(NCOL * NG <= WARPS * 32, "one (group, column) per thread");
    const int tc = threadIdx.x % NCOL, tg = threadIdx.x / NCOL;
    float cn = 0.f, cm = 0.f, cs = 0.f;
    if (tg < NG) {
        cn = w_n[4 * tg];
        cm = w_mean[4 * tg][tc];
        cs = w_m2[4 * tg][tc];
#pragma unroll
        for (int w = 1; w < 4; ++w) chan_merge(cn, cm, cs, w_n[4 * tg + w], w_mean[4 * tg + w][tc], w_m2[4 * tg + w][tc]);
    }
    __syncthreads();  // every partial has been read: rows 0 .. NG - 1 of the arrays are reused for the group results
    if (tg < NG) {
        w_mean[tg][tc] = cm;
        w_m2[tg][tc] = cs;
        if (tc == 0) w_n[tg] = cn;
    }
    __syncthreads();
    if (tg == 0) {
#pragma unroll
        for (int g = 1; g < NG; ++g) chan_merge(cn, cm, cs, w_n[g], w_mean[g][tc], w_m2[g][tc]);
    }
    asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
    const uint32_t rank = tc::cluster_ctarank();
    if (rank != 0) {  // push this CTA's triple into rank 0's inbox, arrive on its mbarrier (release), done
        if (tg == 0) {
            cg::cluster_group cluster = cg::this_cluster();
            cluster.map_shared_rank(&in_mean[0][0], 0)[(rank - 1) * NCOL + tc] = cm;
            cluster.map_shared_rank(&in_m2[0][0], 0)[(rank - 1) * NCOL + tc] = cs;
            if (tc == 0) cluster.map_shared_rank(&in_n[0], 0)[rank - 1] = cn;
            tc::mbar_arrive_cluster(tc::mapa_u32(&inbox_bar, 0));
        }
        return;
    }
    if (tg == 0) {
        tc::mbar_wait_cluster(&inbox_bar, 0);
#pragma unroll
        for (int rk = 0; rk < CM_C - 1; ++rk) chan_merge(cn, cm, cs, in_n[rk], in_mean[rk][tc], in_m2[rk][tc]);  // rank order
        const int64_t col = (int64_t)blockIdx.x * NCOL + tc;
        if (col < a.inner) {
            float second;
            if (a.mode == 0) {
                second = rsqrtf(cs / (float)a.R + a.eps);
            } else {
                const float div = (float)a.R - 1.f;
                second = cs / (div > 0.f ? div : 0.f);
                if (a.take_sqrt) second = sqrtf(second);
            }
            a.out0[o * a.inner + col] = cm;
            a.out1[o * a.inner + col] = second;
        }
    }
}

bool launch_col_moments(const void *x, void *out0, void *out1, int dtype, int64_t outer, int64_t R, int64_t inner, int mode, bool take_sqrt,
                        double eps) {
    if (dtype != KF_FLOAT || R < 1 || inner < 1 || outer < 1 || outer > 65535) return false;
    Runtime &rt = Runtime::get();
    ColMomentsArgs a{};
    a.x = reinterpret_cast<const float *>(x);
    a.out0 = reinterpret_cast<float *>(out0);
    a.out1 = reinterpret_cast<float *>(out1);
    a.outer = outer; a.R = R; a.inner = inner;
    a.eps = (float)eps; a.mode = mode; a.take_sqrt = take_sqrt ? 1 : 0;
    const bool vec4 = inner % 4 == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0;
    // KF_CM_GEOM=narrow: 64-column strips, 256 threads, twice the CTAs (the geometry of reduce_cols_kernel).  Measured at 4096 x 4096:
    // 19.7 us against 15.4 us for the default 128-column strips with 512 threads, so it stays an A/B switch.
    const char *geom = std::getenv("KF_CM_GEOM");
    const bool narrow = vec4 && geom && std::strcmp(geom, "narrow") == 0;
    const int64_t width = vec4 ? (narrow ? 64 : 128) : 32;
    const int64_t strips = (inner + width - 1) / width;
    KF_CHECK(strips < (int64_t)0x7FFFFFFF);
    const dim3 grid((unsigned)strips, CM_C, (unsigned)outer);
    if (narrow) launch_pdl(col_moments_kernel<4, 16, 8>, grid, dim3(256), 0, rt.stream(), a);
    else if (vec4) launch_pdl(col_moments_kernel<4, 32, 16>, grid, dim3(512), 0, rt.stream(), a);
    else launch_pdl(col_moments_kernel<1, 32, 16>, grid, dim3(512), 0, rt.stream(), a);
    rt.post_launch("col_moments_kernel");
    return true;
}

int layer_norm_bwd_ctas(int64_t rows, bool write_dx) {
    // gain gradient only: 8 CTAs of 128 threads per SM keep enough loads in flight (no barrier, no store in the row loop);
    // dx + gain gradient in one pass: two resident CTAs per SM (register budget), each with the next row's loads in flight.
    // KF_LN_BWD=two keeps the two-launch form (A/B runs).
    const char *mode = std::getenv("KF_LN_BWD");
    const bool two = mode && std::strcmp(mode, "two") == 0;
    const char *per_sm = std::getenv("KF_LN_BWD_CTAS");  // tuning hook: resident CTAs per SM of the one-pass kernel
    const int fused_per_sm = per_sm ? std::max(1, std::atoi(per_sm)) : 2;
    const int64_t cap = (int64_t)Runtime::get().props().sm_count * ((write_dx && !two) ? fused_per_sm : 8);
    return (int)std::max<int64_t>(1, std::min<int64_t>(rows, cap));
}

template <typename T>
static void ln_bwd_typed(const LnArgs &a, int ctas, bool write_dx) {
    constexpr int VEC = 16 / sizeof(T);
    Runtime &rt = Runtime::get();
    const char *mode = std::getenv("KF_LN_BWD");
    const bool two = mode && std::strcmp(mode, "two") == 0;
    const bool narrow = mode && std::strcmp(mode, "narrow") == 0;  // KF_LN_BWD=narrow: 128 threads x 8 vectors (A/B runs)
    // default: rows streamed through a shared-memory ring (KF_LN_BWD=regs keeps the register-prefetch kernel); 96 KB of ring per CTA
    const int64_t row_bytes = a.E * (int64_t)sizeof(T);
    const int ring_stages = (int)std::min<int64_t>(8, 98304 / (2 * row_bytes));
    const bool ring = !mode && ring_stages >= 2 && row_bytes % 16 == 0 && reinterpret_cast<uintptr_t>(a.dy) % 16 == 0;
#define KF_LN_BWD(NVV)                                                                                                          \
    do {                                                                                                                        \
        if (write_dx && ring && ((NVV + 1) / 2) * VEC <= 16) { /* 32 columns per thread (16-bit rows of 8192) would spill */    \
            constexpr int NVR = (NVV + 1) / 2; /* 256 threads: half the vectors per thread */                                   \
            static bool attr_done = false;                                                                                      \
            if (!attr_done) {                                                                                                   \
                KF_CUDA(cudaFuncSetAttribute(layer_norm_bwd_ring_kernel<T, VEC, NVR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304)); \
                attr_done = true;                                                                                               \
            }                                                                                                                   \
            launch_pdl(layer_norm_bwd_ring_kernel<T, VEC, NVR>, dim3(ctas), dim3(256), (size_t)(ring_stages * 2 * row_bytes), rt.stream(), a, ring_stages); \
            break;                                                                                                              \
        }                                                                                                                       \
        if (write_dx && !two) {                                                                                                 \
            /* 8 vectors per thread at 128 threads: 256 threads hold 4 each (211 -> ~110 registers, twice the warps per SM) */  \
            if (NVV == 8 && !narrow) layer_norm_bwd_fused_kernel<T, VEC, (NVV + 1) / 2, 256><<<ctas, 256, 0, rt.stream()>>>(a);  \
            else layer_norm_bwd_fused_kernel<T, VEC, NVV, LN_THREADS><<<ctas, LN_THREADS, 0, rt.stream()>>>(a);                  \
            break;                                                                                                              \
        }                                                                                                                       \
        if (write_dx) {                                                                                                         \
            layer_norm_bwd_kernel<T, VEC, NVV, true, false><<<(unsigned)a.rows, LN_THREADS, 0, rt.stream()>>>(a);              \
            rt.post_launch("layer_norm_bwd_kernel");                                                                            \
        }                                                                                                                       \
        layer_norm_bwd_kernel<T, VEC, NVV, false, true><<<ctas, LN_THREADS, 0, rt.stream()>>>(a);                              \
    } while (0)
    KF_CHECK(a.rows < (int64_t)0x7FFFFFFF);
    switch (ln_nv<T>(a.E)) {
    case 1: KF_LN_BWD(1); break;
    case 2: KF_LN_BWD(2); break;
    case 4: KF_LN_BWD(4); break;
    default: KF_LN_BWD(8); break;
    }
#undef KF_LN_BWD
    rt.post_launch("layer_norm_bwd_kernel");
}

// dx may be null (input needs no gradient); dgain_partial is [layer_norm_bwd_ctas(rows, dx != null), E] fp32
void launch_layer_norm_bwd(const void *x, const void *gain, const void *dy, const float *mean, const float *rstd, void *dx,
                           float *dgain_partial, int ctas, int dtype, int64_t rows, int64_t E, bool rms) {
    if (rows == 0) return;
    LnArgs a{};
    a.rms = rms ? 1 : 0;
    a.x = x; a.gain = gain; a.dy = dy; a.dx = dx; a.mean = const_cast<float *>(mean); a.rstd = const_cast<float *>(rstd);
    a.dgain_partial = dgain_partial; a.rows = rows; a.E = E;
    if (dtype == KF_FLOAT) ln_bwd_typed<float>(a, ctas, dx != nullptr);
    else if (dtype == KF_HALF) ln_bwd_typed<__half>(a, ctas, dx != nullptr);
    else ln_bwd_typed<__nv_bfloat16>(a, ctas, dx != nullptr);
}

}  // namespace kf

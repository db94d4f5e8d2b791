// Generic causal attention forward (any head size <= 256, any Sq / Skv, fp32 / fp64 / 16-bit with fp32 math):
// the strict-parity path.  Restates the math of the reference's CausalAttentionRefForwardFN
// (src/device/utils/causal_attention_ref.h:25-64: s = q.k/sqrt(D), keep m >= n, softmax, p.v) as a
// flash-style streaming kernel: one CTA per 32 query rows, KV consumed in blocks of 32 with an online
// softmax, KV blocks above the diagonal skipped (the reference's fast kernel walks all of them,
// src/device/utils/causal_attention.h:113), no S_q x S_kv scratch (the reference allocates one even
// when unused, causal_attention_kernel.cu:22).  Also emits the row log-sum-exp for the backward pass.
#include "ew_common.cuh"

namespace kf {

constexpr int AS_BQ = 32, AS_BKV = 32, AS_THREADS = 256;

template <typename A> __device__ __forceinline__ A a_exp(A x);
template <> __device__ __forceinline__ float a_exp<float>(float x) { return expf(x); }
template <> __device__ __forceinline__ double a_exp<double>(double x) { return exp(x); }
template <typename A> __device__ __forceinline__ A a_log(A x);
template <> __device__ __forceinline__ float a_log<float>(float x) { return logf(x); }
template <> __device__ __forceinline__ double a_log<double>(double x) { return log(x); }

template <typename T, typename A, int DPT>
__global__ void __launch_bounds__(AS_THREADS) attn_fwd_simt_kernel(const AttnPlan p, const A scale) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int D = (int)p.D;
    const int pitch = D + 1;
    A *sQ = reinterpret_cast<A *>(smem_raw);    // [32][D+1]
    A *sK = sQ + AS_BQ * pitch;                 // [32][D+1]
    A *sV = sK + AS_BKV * pitch;                // [32][D]
    A *sS = sV + AS_BKV * D;                    // [32][33]
    const int64_t bh = blockIdx.y;
    const int q0 = blockIdx.x * AS_BQ;
    const T *__restrict__ Q = reinterpret_cast<const T *>(p.q) + bh * p.Sq * D;
    const T *__restrict__ K = reinterpret_cast<const T *>(p.k) + bh * p.Skv * D;
    const T *__restrict__ V = reinterpret_cast<const T *>(p.v) + bh * p.Skv * D;
    const int tid = threadIdx.x, r = tid >> 3, g = tid & 7;
    for (int e = tid; e < AS_BQ * D; e += AS_THREADS) {
        const int rr = e / D, d = e % D;
        sQ[rr * pitch + d] = (q0 + rr < p.Sq) ? cvt_in<A>(Q[(int64_t)(q0 + rr) * D + d]) : A(0);
    }
    A o[DPT];
#pragma unroll
    for (int j = 0; j < DPT; ++j) o[j] = A(0);
    A m = -INFINITY, l = A(0);
    const int kv_end = (int)min((int64_t)p.Skv, (int64_t)q0 + AS_BQ);  // keys above the diagonal are never needed
    for (int kv0 = 0; kv0 < kv_end; kv0 += AS_BKV) {
        __syncthreads();  // previous tile fully consumed (also covers the sQ fill on the first trip)
        for (int e = tid; e < AS_BKV * D; e += AS_THREADS) {
            const int c = e / D, d = e % D;
            const bool ok = kv0 + c < p.Skv;
            sK[c * pitch + d] = ok ? cvt_in<A>(K[(int64_t)(kv0 + c) * D + d]) : A(0);
            sV[c * D + d] = ok ? cvt_in<A>(V[(int64_t)(kv0 + c) * D + d]) : A(0);
        }
        __syncthreads();
        A s[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) s[i] = A(0);
        for (int d = 0; d < D; ++d) {
            const A qv = sQ[r * pitch + d];
#pragma unroll
            for (int i = 0; i < 4; ++i) s[i] = fma(qv, sK[(g * 4 + i) * pitch + d], s[i]);
        }
        A mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int n = kv0 + g * 4 + i;
            s[i] = (n <= q0 + r && n < p.Skv) ? s[i] * scale : (A)-INFINITY;
            mx = s[i] > mx ? s[i] : mx;
        }
#pragma unroll
        for (int off = 1; off < 8; off <<= 1) {
            const A other = __shfl_xor_sync(0xffffffffu, mx, off);
            mx = other > mx ? other : mx;
        }
        const A m_new = mx > m ? mx : m;  // finite from the first block on: key 0 is visible to every row
        const A corr = (m == -INFINITY) ? A(0) : a_exp<A>(m - m_new);
        A rs = A(0);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const A pv = (s[i] == -INFINITY) ? A(0) : a_exp<A>(s[i] - m_new);
            sS[r * 33 + g * 4 + i] = pv;
            rs += pv;
        }
#pragma unroll
        for (int off = 1; off < 8; off <<= 1) rs += __shfl_xor_sync(0xffffffffu, rs, off);
        l = l * corr + rs;
        m = m_new;
        __syncwarp();  // a row's 8 threads live in one warp
#pragma unroll
        for (int j = 0; j < DPT; ++j) o[j] *= corr;
        for (int c = 0; c < AS_BKV; ++c) {
            const A pv = sS[r * 33 + c];
#pragma unroll
            for (int j = 0; j < DPT; ++j) {
                const int d = g + 8 * j;
                if (d < D) o[j] = fma(pv, sV[c * D + d], o[j]);
            }
        }
    }
    if (q0 + r < p.Sq) {
        const A inv = A(1) / l;
        T *O = reinterpret_cast<T *>(p.out) + (bh * p.Sq + q0 + r) * D;
#pragma unroll
        for (int j = 0; j < DPT; ++j) {
            const int d = g + 8 * j;
            if (d < D) O[d] = cvt_out<T, A>(o[j] * inv);
        }
        if (g == 0 && p.lse) reinterpret_cast<A *>(p.lse)[bh * p.Sq + q0 + r] = m + a_log<A>(l);
    }
}

template <typename T, typename A>
static void attn_fwd_simt_typed(const AttnPlan &p) {
    Runtime &rt = Runtime::get();
    KF_CHECK(p.D >= 1 && p.D <= 256, "causal_attention: head size ", p.D, " not supported (max 256)");
    KF_CHECK(p.BH <= 65535, "causal_attention: batch*heads too large for one launch");
    const size_t smem = sizeof(A) * (size_t)(2 * AS_BQ * (p.D + 1) + AS_BKV * p.D + AS_BQ * 33);
    dim3 grid((unsigned)((p.Sq + AS_BQ - 1) / AS_BQ), (unsigned)p.BH);
    const A scale = A(1) / (A)sqrt((double)p.D);
#define KF_ATTN_LAUNCH(DPT)                                                                                      \
    do {                                                                                                         \
        auto kern = attn_fwd_simt_kernel<T, A, DPT>;                                                             \
        KF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));             \
        kern<<<grid, AS_THREADS, smem, rt.stream()>>>(p, scale);                                                 \
    } while (0)
    if (p.D <= 64) KF_ATTN_LAUNCH(8);
    else if (p.D <= 128) KF_ATTN_LAUNCH(16);
    else KF_ATTN_LAUNCH(32);
#undef KF_ATTN_LAUNCH
    rt.post_launch("attn_fwd_simt_kernel");
}

void launch_attention_fwd(const AttnPlan &p) {
    if ((p.dtype == KF_HALF || p.dtype == KF_BFLOAT16) && launch_attention_fwd_tc(p)) return;
    if (p.dtype == KF_FLOAT && launch_attention_fwd_f32_tc(p)) return;
    switch (p.dtype) {
    case KF_FLOAT: attn_fwd_simt_typed<float, float>(p); break;
    case KF_DOUBLE: attn_fwd_simt_typed<double, double>(p); break;
    case KF_HALF: attn_fwd_simt_typed<__half, float>(p); break;
    case KF_BFLOAT16: attn_fwd_simt_typed<__nv_bfloat16, float>(p); break;
    default: KF_CHECK(false, "Unsupported ScalarType ", dtype_name(p.dtype));
    }
}

// P[b][i][j] = j <= i ? exp(S[b][i][j] - lse[b][i]) : 0, in place (generic backward, ops.cpp)
template <typename A>
__global__ void __launch_bounds__(256) attn_probs_kernel(A *__restrict__ S, const A *__restrict__ lse, const int64_t rows, const int64_t Sq,
                                                         const int64_t Skv) {
    const int64_t total = rows * Skv;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = e / Skv, j = e % Skv;
        const int64_t i = row % Sq;
        S[e] = j <= i ? a_exp<A>(S[e] - lse[row]) : A(0);
    }
}

void launch_attn_probs(void *S, const void *lse, int dtype, int64_t BH, int64_t Sq, int64_t Skv) {
    Runtime &rt = Runtime::get();
    const int64_t rows = BH * Sq;
    if (rows * Skv == 0) return;
    const int grid = grid_for(rows * Skv, 256, 8);
    if (dtype == KF_DOUBLE) attn_probs_kernel<double><<<grid, 256, 0, rt.stream()>>>((double *)S, (const double *)lse, rows, Sq, Skv);
    else attn_probs_kernel<float><<<grid, 256, 0, rt.stream()>>>((float *)S, (const float *)lse, rows, Sq, Skv);
    rt.post_launch("attn_probs_kernel");
}

}  // namespace kf

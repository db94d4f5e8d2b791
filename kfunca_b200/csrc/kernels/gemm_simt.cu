// SIMT GEMM (FFMA / DFMA): the strict-parity path for fp32 / fp64 and the fallback for fp16 / bf16 shapes the
// tcgen05 kernel cannot take (TMA needs 16-byte-aligned leading dimensions).
// Replaces the reference's CUTLASS 2.x SIMT `cutlass::gemm::device::Gemm` call (src/device/launcher_cuda.h:537-614,
// src/device/gemm_kernel.cu:26-36) — no CUTLASS here.  C = alpha * op(A) op(B) + beta * C, row-major, batched.
// 128x128x16 CTA tile, 256 threads, 8x8 register tile per thread, double-buffered shared memory,
// operands transposed on the way into shared memory so the inner loop is conflict-free float4 loads.
#include "ew_common.cuh"

namespace kf {

template <typename T> struct SimtAcc { using type = float; };
template <> struct SimtAcc<double> { using type = double; };

template <typename T, typename A>
__device__ __forceinline__ A to_acc(T v) { return cvt_in<A>(v); }

// BM x BN x BK tile; TM x TN per thread; threads = (BM/TM) * (BN/TN)
template <typename T, int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN)) gemm_simt_kernel(const GemmPlan p) {
    using A = typename SimtAcc<T>::type;
    constexpr int NT = (BM / TM) * (BN / TN);
    __shared__ A sA[2][BK][BM + 4];  // [k][m]
    __shared__ A sB[2][BK][BN + 4];  // [k][n]
    const int64_t batch = blockIdx.z;
    const T *__restrict__ Ab = reinterpret_cast<const T *>(p.a) + batch * p.sa;
    const T *__restrict__ Bb = reinterpret_cast<const T *>(p.b) + batch * p.sb;
    T *__restrict__ Cb = reinterpret_cast<T *>(p.c) + batch * p.sc;
    const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);

    A acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = A(0);

    // global -> shared loaders: consecutive threads walk the operand's contiguous dimension
    auto load_tile = [&](int buf, int64_t k0) {
        // A tile: BM x BK
        for (int e = tid; e < BM * BK; e += NT) {
            int m, k;
            if (p.trans_a) { m = e % BM; k = e / BM; }  // stored [K, M]: m contiguous
            else { k = e % BK; m = e / BK; }            // stored [M, K]: k contiguous
            const int64_t gm = m0 + m, gk = k0 + k;
            A v = A(0);
            if (gm < p.M && gk < p.K) v = to_acc<T, A>(p.trans_a ? Ab[gk * p.lda + gm] : Ab[gm * p.lda + gk]);
            sA[buf][k][m] = v;
        }
        for (int e = tid; e < BN * BK; e += NT) {
            int n, k;
            if (p.trans_b) { k = e % BK; n = e / BK; }  // stored [N, K]
            else { n = e % BN; k = e / BN; }            // stored [K, N]
            const int64_t gn = n0 + n, gk = k0 + k;
            A v = A(0);
            if (gn < p.N && gk < p.K) v = to_acc<T, A>(p.trans_b ? Bb[gn * p.ldb + gk] : Bb[gk * p.ldb + gn]);
            sB[buf][k][n] = v;
        }
    };

    const int64_t nk = (p.K + BK - 1) / BK;
    load_tile(0, 0);
    __syncthreads();
    for (int64_t kb = 0; kb < nk; ++kb) {
        const int buf = (int)(kb & 1);
        if (kb + 1 < nk) load_tile(buf ^ 1, (kb + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            A ra[TM], rb[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) ra[i] = sA[buf][k][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) rb[j] = sB[buf][k][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fma(ra[i], rb[j], acc[i][j]);
        }
        __syncthreads();
    }
    const A alpha = (A)p.alpha, beta = (A)p.beta;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int64_t gm = m0 + ty * TM + i;
        if (gm >= p.M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int64_t gn = n0 + tx * TN + j;
            if (gn >= p.N) continue;
            A v = alpha * acc[i][j];
            if (p.beta != 0.f) v += beta * to_acc<T, A>(Cb[gm * p.ldc + gn]);
            // + residual: the product is rounded to the storage dtype first, like the unfused gemm-then-add it replaces
            if (p.residual) v = to_acc<T, A>(cvt_out<T, A>(v)) + to_acc<T, A>(reinterpret_cast<const T *>(p.residual)[batch * p.sr + gm * p.ldr + gn]);
            Cb[gm * p.ldc + gn] = cvt_out<T, A>(v);
        }
    }
}

template <typename T, int BM, int BN, int BK, int TM, int TN>
static void launch_simt_cfg(const GemmPlan &p) {
    Runtime &rt = Runtime::get();
    KF_CHECK(p.batch <= 65535, "gemm: batch too large");
    dim3 grid((unsigned)((p.N + BN - 1) / BN), (unsigned)((p.M + BM - 1) / BM), (unsigned)p.batch);
    gemm_simt_kernel<T, BM, BN, BK, TM, TN><<<grid, (BM / TM) * (BN / TN), 0, rt.stream()>>>(p);
    rt.post_launch("gemm_simt_kernel");
}

void launch_gemm_simt(const GemmPlan &p) {
    switch (p.dtype) {
    case KF_FLOAT: launch_simt_cfg<float, 128, 128, 8, 8, 8>(p); break;
    case KF_DOUBLE: launch_simt_cfg<double, 64, 64, 8, 4, 4>(p); break;
    case KF_HALF: launch_simt_cfg<__half, 128, 128, 8, 8, 8>(p); break;
    case KF_BFLOAT16: launch_simt_cfg<__nv_bfloat16, 128, 128, 8, 8, 8>(p); break;
    default: KF_CHECK(false, "Unsupported ScalarType ", dtype_name(p.dtype));
    }
}

void launch_gemm(const GemmPlan &p) {
    if ((p.dtype == KF_HALF || p.dtype == KF_BFLOAT16) && launch_gemm_tc(p)) return;
    if (p.dtype == KF_FLOAT && launch_gemm_f32_tc(p)) return;
    launch_gemm_simt(p);
}

}  // namespace kf

// sm_100a elementwise engine: out = op(a[, b]) over a host-collapsed broadcast/stride plan.
// Replaces the reference's gpu_kernel -> Vectorized/Unroll/LegacyElementwiseKernel family
// (src/device/utils/tensor_loops.h:49-369).  Two kernels:
//   ew_pack_kernel   : each thread moves packs of 8 elements along the collapsed inner dim with
//                      128-bit loads/stores (2 x LDG.128 per fp32 operand in flight), operands may be
//                      inner-contiguous or inner-broadcast (stride 0); dtype casts are fused in registers
//                      (statically typed for the all-fp32 / all-bf16 / all-fp16 cases, a warp-uniform
//                      switch otherwise — never the reference's per-element scalar dynamic cast path).
//   ew_scalar_kernel : arbitrary strides, lane-consecutive elements (coalesced when the inner stride is
//                      one element), 4 independent elements in flight per thread.
#include <type_traits>

#include "ew_common.cuh"

namespace kf {

template <int OP, typename A>
__device__ __forceinline__ A apply_op(A a, A b) {
    if constexpr (OP == EW_ADD) return a + b;
    else if constexpr (OP == EW_SUB) return a - b;
    else if constexpr (OP == EW_MUL) return a * b;
    else if constexpr (OP == EW_DIV) return a / b;
    else if constexpr (OP == EW_COPY) return a;
    else if constexpr (OP == EW_FILL) return b;
    else if constexpr (OP == EW_SQRT) return (A)sqrt(a);
    else if constexpr (OP == EW_RSQRT) return (A)1 / (A)sqrt(a);
    else return -a;
}
template <int OP>
__device__ __forceinline__ bool apply_op_bool(bool a, bool b) {
    // C++ arithmetic on bool promotes to int and converts back: + is OR, - is XOR, * is AND.
    if constexpr (OP == EW_ADD) return a || b;
    else if constexpr (OP == EW_SUB) return a != b;
    else if constexpr (OP == EW_MUL) return a && b;
    else if constexpr (OP == EW_DIV) return a && b;
    else if constexpr (OP == EW_FILL) return b;
    else return a;
}
template <int OP, typename A>
__device__ __forceinline__ A apply(A a, A b) {
    if constexpr (sizeof(A) == 1) return apply_op_bool<OP>(a, b);
    else if constexpr (OP == EW_DIV && std::is_same<A, int64_t>::value) return b == 0 ? A(-1) : a / b;  // int64: no trap
    else return apply_op<OP, A>(a, b);
}

constexpr int kPack = 8;

template <int OP, int ACC, int NIN, int SDT>
__global__ void __launch_bounds__(256) ew_pack_kernel(const EwPlan p, const int64_t npacks, const uint32_t inner_packs) {
    pdl_enter();
    using A = typename AccType<ACC>::type;
    const int dt0 = SDT >= 0 ? SDT : p.dtype[0];
    const int dt1 = SDT >= 0 ? SDT : p.dtype[1];
    const int dt2 = SDT >= 0 ? SDT : p.dtype[2];
    const A sval = static_cast<A>(p.scalar);
    for (int64_t pk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pk < npacks; pk += (int64_t)gridDim.x * blockDim.x) {
        int64_t off0, off1 = 0, off2 = 0;
        if (p.ndim == 1) {
            off0 = pk * kPack * p.stride[0][0];
            if (NIN >= 1) off1 = pk * kPack * p.stride[1][0];
            if (NIN >= 2) off2 = pk * kPack * p.stride[2][0];
        } else {
            uint32_t rest = (uint32_t)pk;  // host guarantees npacks < 2^32 on this path
            const uint32_t i0 = rest % inner_packs;
            rest /= inner_packs;
            off0 = (int64_t)i0 * kPack * p.stride[0][0];
            if (NIN >= 1) off1 = (int64_t)i0 * kPack * p.stride[1][0];
            if (NIN >= 2) off2 = (int64_t)i0 * kPack * p.stride[2][0];
            for (int d = 1; d < p.ndim; ++d) {
                const uint32_t s = (uint32_t)p.shape[d];
                const uint32_t i = rest % s;
                rest /= s;
                off0 += (int64_t)i * p.stride[0][d];
                if (NIN >= 1) off1 += (int64_t)i * p.stride[1][d];
                if (NIN >= 2) off2 += (int64_t)i * p.stride[2][d];
            }
        }
        A a[kPack], b[kPack], r[kPack];
        if constexpr (NIN >= 1) {
            if (p.stride[1][0] != 0) {
                load_pack<A, kPack>((const char *)p.ptr[1] + off1, dt1, a);
            } else {
                const A s = load_scalar<A>((const char *)p.ptr[1] + off1, dt1);
#pragma unroll
                for (int i = 0; i < kPack; ++i) a[i] = s;
            }
        }
        if constexpr (NIN >= 2) {
            if (p.b_is_scalar) {
#pragma unroll
                for (int i = 0; i < kPack; ++i) b[i] = sval;
            } else if (p.stride[2][0] != 0) {
                load_pack<A, kPack>((const char *)p.ptr[2] + off2, dt2, b);
            } else {
                const A s = load_scalar<A>((const char *)p.ptr[2] + off2, dt2);
#pragma unroll
                for (int i = 0; i < kPack; ++i) b[i] = s;
            }
        } else {
#pragma unroll
            for (int i = 0; i < kPack; ++i) b[i] = sval;
        }
        if constexpr (NIN == 0) {
#pragma unroll
            for (int i = 0; i < kPack; ++i) a[i] = sval;
        }
#pragma unroll
        for (int i = 0; i < kPack; ++i) r[i] = apply<OP, A>(a[i], b[i]);
        store_pack<A, kPack>((char *)p.ptr[0] + off0, dt0, r);
    }
}

template <int OP, int ACC, int NIN, typename IndexT>
__global__ void __launch_bounds__(256) ew_scalar_kernel(const EwPlan p) {
    pdl_enter();
    using A = typename AccType<ACC>::type;
    constexpr int U = 4;
    const A sval = static_cast<A>(p.scalar);
    const int64_t step = (int64_t)gridDim.x * blockDim.x * U;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x * U; base < p.numel; base += step) {
        A a[U], b[U];
        int64_t off0[U];
        bool live[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + (int64_t)u * blockDim.x + threadIdx.x;
            live[u] = i < p.numel;
            a[u] = sval;
            b[u] = sval;
            off0[u] = 0;
            if (live[u]) {
                IndexT rest = (IndexT)i;
                int64_t o0 = 0, o1 = 0, o2 = 0;
                for (int d = 0; d < p.ndim; ++d) {
                    const IndexT s = (IndexT)p.shape[d];
                    const IndexT idx = rest % s;
                    rest /= s;
                    o0 += (int64_t)idx * p.stride[0][d];
                    if (NIN >= 1) o1 += (int64_t)idx * p.stride[1][d];
                    if (NIN >= 2) o2 += (int64_t)idx * p.stride[2][d];
                }
                off0[u] = o0;
                if constexpr (NIN >= 1) a[u] = load_scalar<A>((const char *)p.ptr[1] + o1, p.dtype[1]);
                if constexpr (NIN >= 2) {
                    if (!p.b_is_scalar) b[u] = load_scalar<A>((const char *)p.ptr[2] + o2, p.dtype[2]);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (live[u]) store_scalar<A>((char *)p.ptr[0] + off0[u], p.dtype[0], apply<OP, A>(a[u], b[u]));
        }
    }
}

template <int OP, int ACC, int NIN>
static void launch_typed(const EwPlan &p) {
    Runtime &rt = Runtime::get();
    cudaStream_t st = rt.stream();
    // ---- can we run the 128-bit pack path?
    bool pack_ok = p.shape[0] % kPack == 0 && p.numel / kPack < (int64_t)0xFFFFFFFFll;
    for (int t = 0; t <= NIN && pack_ok; ++t) {
        if (t == 2 && p.b_is_scalar) continue;
        const int64_t isz = (int64_t)element_size(p.dtype[t]);
        const int64_t pack_bytes = isz * kPack;
        const int64_t align = pack_bytes >= 16 ? 16 : pack_bytes;
        const int64_t s0 = p.stride[t][0];
        if (t == 0 ? s0 != isz : (s0 != isz && s0 != 0)) pack_ok = false;
        if (s0 != 0) {
            if ((uintptr_t)p.ptr[t] % align) pack_ok = false;
            for (int d = 1; d < p.ndim; ++d)
                if (p.stride[t][d] % align) pack_ok = false;
        }
    }
    if (pack_ok) {
        const int64_t npacks = p.numel / kPack;
        const uint32_t inner_packs = (uint32_t)(p.shape[0] / kPack);
        const int grid = grid_for(npacks, 256, 8);
        const bool same = (NIN < 1 || p.dtype[1] == p.dtype[0]) && (NIN < 2 || p.b_is_scalar || p.dtype[2] == p.dtype[0]);
        if (ACC == ACC_F32 && same && p.dtype[0] == KF_FLOAT)
            launch_pdl(ew_pack_kernel<OP, ACC, NIN, KF_FLOAT>, dim3(grid), dim3(256), 0, st, p, npacks, inner_packs);
        else if (ACC == ACC_F32 && same && p.dtype[0] == KF_BFLOAT16)
            launch_pdl(ew_pack_kernel<OP, ACC, NIN, KF_BFLOAT16>, dim3(grid), dim3(256), 0, st, p, npacks, inner_packs);
        else if (ACC == ACC_F32 && same && p.dtype[0] == KF_HALF)
            launch_pdl(ew_pack_kernel<OP, ACC, NIN, KF_HALF>, dim3(grid), dim3(256), 0, st, p, npacks, inner_packs);
        else
            launch_pdl(ew_pack_kernel<OP, ACC, NIN, -1>, dim3(grid), dim3(256), 0, st, p, npacks, inner_packs);
        rt.post_launch("ew_pack_kernel");
        return;
    }
    const int grid = grid_for((p.numel + 3) / 4, 256, 8);
    if (p.numel < (int64_t)0x7FFFFFFF)
        launch_pdl(ew_scalar_kernel<OP, ACC, NIN, uint32_t>, dim3(grid), dim3(256), 0, st, p);
    else
        launch_pdl(ew_scalar_kernel<OP, ACC, NIN, uint64_t>, dim3(grid), dim3(256), 0, st, p);
    rt.post_launch("ew_scalar_kernel");
}

template <int OP, int NIN>
static void launch_acc(const EwPlan &p) {
    switch (p.acc) {
    case ACC_F32: launch_typed<OP, ACC_F32, NIN>(p); break;
    case ACC_F64: launch_typed<OP, ACC_F64, NIN>(p); break;
    case ACC_I64: launch_typed<OP, ACC_I64, NIN>(p); break;
    default: launch_typed<OP, ACC_BOOL, NIN>(p); break;
    }
}
template <int OP>
static void launch_float_unary(const EwPlan &p) {
    switch (p.acc) {
    case ACC_F32: launch_typed<OP, ACC_F32, 1>(p); break;
    case ACC_F64: launch_typed<OP, ACC_F64, 1>(p); break;
    default: KF_CHECK(false, "unary math op needs a floating dtype");
    }
}

void launch_elementwise(const EwPlan &p) {
    if (p.numel == 0) return;
    switch (p.op) {
    case EW_ADD: launch_acc<EW_ADD, 2>(p); break;
    case EW_SUB: launch_acc<EW_SUB, 2>(p); break;
    case EW_MUL: launch_acc<EW_MUL, 2>(p); break;
    case EW_DIV: launch_acc<EW_DIV, 2>(p); break;
    case EW_COPY: launch_acc<EW_COPY, 1>(p); break;
    case EW_FILL: launch_acc<EW_FILL, 0>(p); break;
    case EW_SQRT: launch_float_unary<EW_SQRT>(p); break;
    case EW_RSQRT: launch_float_unary<EW_RSQRT>(p); break;
    case EW_NEG:
        if (p.acc == ACC_I64) launch_typed<EW_NEG, ACC_I64, 1>(p);
        else launch_float_unary<EW_NEG>(p);
        break;
    default: KF_CHECK(false, "unknown elementwise op ", p.op);
    }
}

}  // namespace kf

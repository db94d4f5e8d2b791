// Device-side dtype plumbing shared by the elementwise / copy / reduce kernels.
// Cast rules restate the reference's fetch_and_cast / cast_and_store
// (src/device/utils/tensor_memory_access.h:13-37): value -> static_cast<acc_t> -> op -> static_cast<out_t>,
// 16-bit floats always through fp32 with round-to-nearest-even.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdint>
#include <cstdlib>
#include <tuple>

#include "../kernels.h"
#include "../runtime.h"

namespace kf {

// Programmatic dependent launch for the microsecond-scale streaming kernels: with the launch attribute, the NEXT kernel on the
// stream may start scheduling its CTAs while this one drains (a kernel boundary otherwise costs 2 - 4 us, a third of an 11 us
// kernel).  pdl_enter() is the FIRST statement of every kernel launched through launch_pdl: it lets the dependents go and then
// waits until every earlier kernel has completed and flushed — nothing touches global memory before it.  KF_PDL=0 switches it off.
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
inline bool pdl_on() {
    static const bool on = [] {
        const char *e = std::getenv("KF_PDL");
        return !(e && e[0] == '0');
    }();
    return on;
}
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    static_assert(sizeof...(KArgs) == sizeof...(Args), "argument count");
    std::tuple<KArgs...> store{static_cast<KArgs>(args)...};  // exact parameter types, addressable
    void *ptrs[sizeof...(KArgs) ? sizeof...(KArgs) : 1];
    size_t i = 0;
    std::apply([&](auto &...e) { ((ptrs[i++] = const_cast<void *>(static_cast<const void *>(&e))), ...); }, store);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_on() ? 1 : 0;
    KF_CUDA(cudaLaunchKernelExC(&cfg, reinterpret_cast<const void *>(kernel), ptrs));
}

template <int ACC> struct AccType;
template <> struct AccType<ACC_F32> { using type = float; };
template <> struct AccType<ACC_F64> { using type = double; };
template <> struct AccType<ACC_I64> { using type = int64_t; };
template <> struct AccType<ACC_BOOL> { using type = bool; };

template <typename T, int P>
struct alignas((sizeof(T) * P) >= 16 ? 16 : (sizeof(T) * P)) Pack {
    T v[P];
};

template <typename A, typename T>
__device__ __forceinline__ A cvt_in(T x) { return static_cast<A>(x); }
template <typename A>
__device__ __forceinline__ A cvt_in(__half x) { return static_cast<A>(__half2float(x)); }
template <typename A>
__device__ __forceinline__ A cvt_in(__nv_bfloat16 x) { return static_cast<A>(__bfloat162float(x)); }

template <typename T, typename A>
__device__ __forceinline__ T cvt_out(A x) { return static_cast<T>(x); }
template <> __device__ __forceinline__ __half cvt_out<__half, float>(float x) { return __float2half_rn(x); }
template <> __device__ __forceinline__ __half cvt_out<__half, double>(double x) { return __float2half_rn((float)x); }
template <> __device__ __forceinline__ __half cvt_out<__half, int64_t>(int64_t x) { return __float2half_rn((float)x); }
template <> __device__ __forceinline__ __half cvt_out<__half, bool>(bool x) { return __float2half_rn((float)x); }
template <> __device__ __forceinline__ __nv_bfloat16 cvt_out<__nv_bfloat16, float>(float x) { return __float2bfloat16_rn(x); }
template <> __device__ __forceinline__ __nv_bfloat16 cvt_out<__nv_bfloat16, double>(double x) { return __float2bfloat16_rn((float)x); }
template <> __device__ __forceinline__ __nv_bfloat16 cvt_out<__nv_bfloat16, int64_t>(int64_t x) { return __float2bfloat16_rn((float)x); }
template <> __device__ __forceinline__ __nv_bfloat16 cvt_out<__nv_bfloat16, bool>(bool x) { return __float2bfloat16_rn((float)x); }

// dispatch a runtime dtype to a compile-time C type inside device code (warp-uniform branch)
#define KF_DEV_DTYPE_SWITCH(dt, ...)                                              \
    switch (dt) {                                                                  \
    case KF_BOOL: { using scalar_t = bool; __VA_ARGS__; } break;                   \
    case KF_BYTE: { using scalar_t = uint8_t; __VA_ARGS__; } break;                \
    case KF_CHAR: { using scalar_t = int8_t; __VA_ARGS__; } break;                 \
    case KF_SHORT: { using scalar_t = int16_t; __VA_ARGS__; } break;               \
    case KF_INT: { using scalar_t = int32_t; __VA_ARGS__; } break;                 \
    case KF_LONG: { using scalar_t = int64_t; __VA_ARGS__; } break;                \
    case KF_HALF: { using scalar_t = __half; __VA_ARGS__; } break;                 \
    case KF_BFLOAT16: { using scalar_t = __nv_bfloat16; __VA_ARGS__; } break;      \
    case KF_FLOAT: { using scalar_t = float; __VA_ARGS__; } break;                 \
    default: { using scalar_t = double; __VA_ARGS__; } break;                      \
    }

template <typename A>
__device__ __forceinline__ A load_scalar(const char *p, int dt) {
    A r;
    KF_DEV_DTYPE_SWITCH(dt, r = cvt_in<A>(*reinterpret_cast<const scalar_t *>(p)));
    return r;
}
template <typename A>
__device__ __forceinline__ void store_scalar(char *p, int dt, A v) {
    KF_DEV_DTYPE_SWITCH(dt, *reinterpret_cast<scalar_t *>(p) = cvt_out<scalar_t, A>(v));
}

template <typename A, int P>
__device__ __forceinline__ void load_pack(const char *p, int dt, A (&out)[P]) {
    KF_DEV_DTYPE_SWITCH(dt, {
        Pack<scalar_t, P> pk = *reinterpret_cast<const Pack<scalar_t, P> *>(p);
        _Pragma("unroll") for (int i = 0; i < P; ++i) out[i] = cvt_in<A>(pk.v[i]);
    });
}
template <typename A, int P>
__device__ __forceinline__ void store_pack(char *p, int dt, const A (&in)[P]) {
    KF_DEV_DTYPE_SWITCH(dt, {
        Pack<scalar_t, P> pk;
        _Pragma("unroll") for (int i = 0; i < P; ++i) pk.v[i] = cvt_out<scalar_t, A>(in[i]);
        *reinterpret_cast<Pack<scalar_t, P> *>(p) = pk;
    });
}

// host-side dtype switch
#define KF_HOST_DTYPE_SWITCH(dt, ...)                                             \
    switch (dt) {                                                                  \
    case KF_BOOL: { using scalar_t = bool; __VA_ARGS__; } break;                   \
    case KF_BYTE: { using scalar_t = uint8_t; __VA_ARGS__; } break;                \
    case KF_CHAR: { using scalar_t = int8_t; __VA_ARGS__; } break;                 \
    case KF_SHORT: { using scalar_t = int16_t; __VA_ARGS__; } break;               \
    case KF_INT: { using scalar_t = int32_t; __VA_ARGS__; } break;                 \
    case KF_LONG: { using scalar_t = int64_t; __VA_ARGS__; } break;                \
    case KF_HALF: { using scalar_t = __half; __VA_ARGS__; } break;                 \
    case KF_BFLOAT16: { using scalar_t = __nv_bfloat16; __VA_ARGS__; } break;      \
    case KF_FLOAT: { using scalar_t = float; __VA_ARGS__; } break;                 \
    case KF_DOUBLE: { using scalar_t = double; __VA_ARGS__; } break;               \
    default: KF_CHECK(false, "Unsupported ScalarType ", dt);                       \
    }

inline int grid_for(int64_t work_items, int threads, int ctas_per_sm) {
    const int sms = Runtime::get().props().sm_count;
    int64_t need = (work_items + threads - 1) / threads;
    int64_t cap = (int64_t)sms * ctas_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

}  // namespace kf

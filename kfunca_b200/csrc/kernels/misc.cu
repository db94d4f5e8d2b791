// index_put_ scatter, device info and host-side 16-bit float conversions.
#include <cstdio>

#include "ew_common.cuh"

namespace kf {

// ---- index_put_ (ref: IndexElementwiseKernel, src/device/utils/tensor_index.h:19-143):
// self[idx0[i], idx1[i], ...] = values[i]; negative indices wrap; later duplicates win is unspecified there too.
struct IndexPutArgs {
    char *self;
    const char *values;
    const int64_t *idx[KF_MAX_DIMS];
    int64_t size[KF_MAX_DIMS];
    int64_t stride_bytes[KF_MAX_DIMS];
    int nidx;
    int itemsize;
    int64_t n;
};

__global__ void __launch_bounds__(256) index_put_kernel(const IndexPutArgs a) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t off = 0;
        bool ok = true;
        for (int d = 0; d < a.nidx; ++d) {
            int64_t ix = a.idx[d][i];
            if (ix < 0) ix += a.size[d];
            ok = ok && ix >= 0 && ix < a.size[d];
            off += ix * a.stride_bytes[d];
        }
        if (!ok) continue;  // out-of-range indices are dropped instead of corrupting memory
        const char *src = a.values + i * a.itemsize;
        char *dst = a.self + off;
        switch (a.itemsize) {
        case 1: *dst = *src; break;
        case 2: *(uint16_t *)dst = *(const uint16_t *)src; break;
        case 4: *(uint32_t *)dst = *(const uint32_t *)src; break;
        default: *(uint64_t *)dst = *(const uint64_t *)src; break;
        }
    }
}

void launch_index_put(void *self, int dtype, const int64_t *self_shape, const int64_t *self_stride, int nidx, int64_t,
                      const int64_t *const *idx_ptrs, const void *values, int64_t n) {
    if (n == 0) return;
    Runtime &rt = Runtime::get();
    IndexPutArgs a{};
    a.self = (char *)self;
    a.values = (const char *)values;
    a.nidx = nidx;
    a.itemsize = (int)element_size(dtype);
    a.n = n;
    for (int d = 0; d < nidx; ++d) {
        a.idx[d] = idx_ptrs[d];
        a.size[d] = self_shape[d];
        a.stride_bytes[d] = self_stride[d] * a.itemsize;
    }
    index_put_kernel<<<grid_for(n, 256, 8), 256, 0, rt.stream()>>>(a);
    rt.post_launch("index_put_kernel");
}

// ---- embedding gather: out[i, :] = weight[idx[i], :]  (bit-exact row copy; negative ids wrap like index_put_, bad ids give zeros)
template <typename V>
__global__ void __launch_bounds__(256) embedding_fwd_kernel(const V *__restrict__ w, const int64_t *__restrict__ idx, V *__restrict__ out,
                                                            const int64_t n, const int64_t Vrows, const int64_t row_vecs) {
    const int64_t total = n * row_vecs;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / row_vecs, c = t % row_vecs;
        int64_t ix = __ldg(idx + i);
        if (ix < 0) ix += Vrows;
        V v{};
        if (ix >= 0 && ix < Vrows) v = __ldg(w + ix * row_vecs + c);
        out[t] = v;
    }
}

void launch_embedding_fwd(const void *weight, const int64_t *idx, void *out, int dtype, int64_t n, int64_t V, int64_t E) {
    Runtime &rt = Runtime::get();
    const int64_t row_bytes = E * (int64_t)element_size(dtype);
    const bool v16 = row_bytes % 16 == 0 && reinterpret_cast<uintptr_t>(weight) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0;
    if (v16) {
        const int64_t rv = row_bytes / 16;
        embedding_fwd_kernel<uint4><<<grid_for(n * rv, 256, 8), 256, 0, rt.stream()>>>((const uint4 *)weight, idx, (uint4 *)out, n, V, rv);
    } else {
        switch (element_size(dtype)) {
        case 1: embedding_fwd_kernel<uint8_t><<<grid_for(n * E, 256, 8), 256, 0, rt.stream()>>>((const uint8_t *)weight, idx, (uint8_t *)out, n, V, E); break;
        case 2: embedding_fwd_kernel<uint16_t><<<grid_for(n * E, 256, 8), 256, 0, rt.stream()>>>((const uint16_t *)weight, idx, (uint16_t *)out, n, V, E); break;
        case 4: embedding_fwd_kernel<uint32_t><<<grid_for(n * E, 256, 8), 256, 0, rt.stream()>>>((const uint32_t *)weight, idx, (uint32_t *)out, n, V, E); break;
        default: embedding_fwd_kernel<uint64_t><<<grid_for(n * E, 256, 8), 256, 0, rt.stream()>>>((const uint64_t *)weight, idx, (uint64_t *)out, n, V, E); break;
        }
    }
    rt.post_launch("embedding_fwd_kernel");
}

// ---- embedding backward: dW[v, :] = sum over the positions p with idx[p] == v of grad[p, :], summed in ascending p.
// `sorted_idx` / `sorted_pos` come from the library's stable sort of idx.  One warp per run start; lanes stride the columns.
template <typename T, typename A>
__global__ void __launch_bounds__(256) embedding_bwd_kernel(const T *__restrict__ grad, const int64_t *__restrict__ sidx, const int64_t *__restrict__ spos,
                                                            T *__restrict__ dw, const int64_t n, const int64_t Vrows, const int64_t E) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < n; p += warps) {
        const int64_t v = sidx[p];
        if (p > 0 && sidx[p - 1] == v) continue;  // not the start of a run
        int64_t row = v < 0 ? v + Vrows : v;
        if (row < 0 || row >= Vrows) continue;
        int64_t q_end = p + 1;
        while (q_end < n && sidx[q_end] == v) ++q_end;
        for (int64_t c = lane; c < E; c += 32) {
            A acc = A(0);
            for (int64_t q = p; q < q_end; ++q) acc += cvt_in<A>(grad[spos[q] * E + c]);
            dw[row * E + c] = cvt_out<T, A>(acc);
        }
    }
}

void launch_embedding_bwd(const void *grad, const int64_t *sorted_idx, const int64_t *sorted_pos, void *dweight, int dtype, int64_t n, int64_t V,
                          int64_t E) {
    Runtime &rt = Runtime::get();
    const int grid = grid_for(n * 32, 256, 8);
    switch (dtype) {
    case KF_FLOAT: embedding_bwd_kernel<float, float><<<grid, 256, 0, rt.stream()>>>((const float *)grad, sorted_idx, sorted_pos, (float *)dweight, n, V, E); break;
    case KF_DOUBLE: embedding_bwd_kernel<double, double><<<grid, 256, 0, rt.stream()>>>((const double *)grad, sorted_idx, sorted_pos, (double *)dweight, n, V, E); break;
    case KF_HALF: embedding_bwd_kernel<__half, float><<<grid, 256, 0, rt.stream()>>>((const __half *)grad, sorted_idx, sorted_pos, (__half *)dweight, n, V, E); break;
    case KF_BFLOAT16:
        embedding_bwd_kernel<__nv_bfloat16, float><<<grid, 256, 0, rt.stream()>>>((const __nv_bfloat16 *)grad, sorted_idx, sorted_pos, (__nv_bfloat16 *)dweight, n, V, E);
        break;
    default: KF_CHECK(false, "embedding backward: floating dtypes only");
    }
    rt.post_launch("embedding_bwd_kernel");
}

// ---- counter-based uniform fill (test / bench input generator; host twin: oracle.counter_uniform)
__device__ __forceinline__ uint64_t mix64(uint64_t z) {  // splitmix64 finaliser
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
template <typename T>
__global__ void __launch_bounds__(256) fill_random_kernel(T *__restrict__ p, const int64_t n, const uint64_t seed, const float lo, const float span) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t h = mix64((uint64_t)i + seed * 0x9E3779B97F4A7C15ull);
        const float u = (float)(uint32_t)(h >> 40) * 5.9604644775390625e-8f;  // 24 random bits -> [0, 1), exact in fp32
        p[i] = cvt_out<T, float>(__fadd_rn(lo, __fmul_rn(span, u)));
    }
}
template <> __device__ __forceinline__ double cvt_out<double, float>(float x) { return (double)x; }

void launch_fill_random(void *p, int dtype, int64_t n, uint64_t seed, float lo, float hi) {
    Runtime &rt = Runtime::get();
    const int grid = grid_for(n, 256, 16);
    const float span = hi - lo;
    switch (dtype) {
    case KF_FLOAT: fill_random_kernel<float><<<grid, 256, 0, rt.stream()>>>((float *)p, n, seed, lo, span); break;
    case KF_DOUBLE: fill_random_kernel<double><<<grid, 256, 0, rt.stream()>>>((double *)p, n, seed, lo, span); break;
    case KF_HALF: fill_random_kernel<__half><<<grid, 256, 0, rt.stream()>>>((__half *)p, n, seed, lo, span); break;
    case KF_BFLOAT16: fill_random_kernel<__nv_bfloat16><<<grid, 256, 0, rt.stream()>>>((__nv_bfloat16 *)p, n, seed, lo, span); break;
    default: KF_CHECK(false, "random fill: floating dtypes only");
    }
    rt.post_launch("fill_random_kernel");
}

std::string device_info_string() {
    Runtime &rt = Runtime::get();
    const DeviceProps &p = rt.props();
    char buf[1024];
    std::snprintf(buf, sizeof(buf),
                  "kfunca_b200 device %d: %s (sm_%d%d), %d SMs, %.1f GiB HBM, L2 %.1f MiB, max dynamic smem/CTA %d KiB\n",
                  rt.device(), p.name, p.cc_major, p.cc_minor, p.sm_count, (double)p.total_mem / (1 << 30),
                  (double)p.l2_bytes / (1 << 20), p.max_smem_optin >> 10);
    return buf;
}

// ---- host 16-bit float conversions, round-to-nearest-even
uint16_t f32_to_bf16_bits(float f) {
    uint32_t x;
    std::memcpy(&x, &f, 4);
    if ((x & 0x7fffffffu) > 0x7f800000u) return 0x7fc0;  // NaN (ref: half.h:268-290)
    x += 0x7fffu + ((x >> 16) & 1u);
    return (uint16_t)(x >> 16);
}
float bf16_bits_to_f32(uint16_t h) {
    uint32_t x = (uint32_t)h << 16;
    float f;
    std::memcpy(&f, &x, 4);
    return f;
}
uint16_t f32_to_f16_bits(float f) {
    uint32_t x;
    std::memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    x &= 0x7fffffffu;
    if (x > 0x7f800000u) return (uint16_t)(sign | 0x7e00u);
    if (x >= 0x47800000u) return (uint16_t)(sign | 0x7c00u);  // overflow -> inf (also inf)
    if (x < 0x38800000u) {                                     // subnormal half or zero
        if (x < 0x33000000u) return (uint16_t)sign;
        const int shift = 113 - (int)(x >> 23);
        uint32_t mant = (x & 0x7fffffu) | 0x800000u;
        const uint32_t lsb = 1u << (shift + 13);
        const uint32_t half_ulp = lsb >> 1;
        uint32_t r = mant >> (shift + 13);
        const uint32_t rem = mant & (lsb - 1);
        if (rem > half_ulp || (rem == half_ulp && (r & 1))) ++r;
        return (uint16_t)(sign | r);
    }
    uint32_t r = x - 0x38000000u;  // rebias exponent
    const uint32_t rem = r & 0x1fffu;
    r >>= 13;
    if (rem > 0x1000u || (rem == 0x1000u && (r & 1))) ++r;
    return (uint16_t)(sign | r);
}
float f16_bits_to_f32(uint16_t h) {
    const uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1f, mant = h & 0x3ffu, x;
    if (exp == 0) {
        if (mant == 0) {
            x = sign;
        } else {
            int e = -1;
            do {
                ++e;
                mant <<= 1;
            } while (!(mant & 0x400u));
            x = sign | ((uint32_t)(112 - e) << 23) | ((mant & 0x3ffu) << 13);
        }
    } else if (exp == 31) {
        x = sign | 0x7f800000u | (mant << 13);
    } else {
        x = sign | ((exp + 112) << 23) | (mant << 13);
    }
    float f;
    std::memcpy(&f, &x, 4);
    return f;
}

}  // namespace kf

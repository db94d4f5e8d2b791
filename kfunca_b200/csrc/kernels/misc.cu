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

// Batched tiled transpose-copy: the B200 replacement for the reference's un-tiled strided gather
// behind permute().contiguous() (LegacyElementwiseKernel + OffsetCalculator,
// src/device/utils/tensor_loops.h:320-331).  A 64x64 tile is read in full lines along the
// input's unit-stride dim, staged in padded shared memory, and written in full lines along the
// output's unit-stride dim, so both HBM sides are coalesced.  Pure byte movement => bit-exact.
#include "ew_common.cuh"

namespace kf {

constexpr int TT = 64;  // tile edge (elements)

template <typename T>
__global__ void __launch_bounds__(256) transpose_tiled_kernel(const TransposePlan p, const int64_t tiles0, const int64_t tilesT) {
    pdl_enter();
    __shared__ T tile[TT][TT + 1];
    int64_t bid = blockIdx.x;
    const int64_t t0 = bid % tiles0;
    bid /= tiles0;
    const int64_t tt = bid % tilesT;
    bid /= tilesT;
    // remaining dims (all except 0 and tdim) -> base offsets
    int64_t in_base = 0, out_base = 0;
    for (int d = 1; d < p.ndim; ++d) {
        if (d == p.tdim) continue;
        const int64_t i = bid % p.shape[d];
        bid /= p.shape[d];
        in_base += i * p.in_stride[d];
        out_base += i * p.out_stride[d];
    }
    const int64_t i0 = t0 * TT, it = tt * TT;
    const int64_t n0 = p.shape[0], nt = p.shape[p.tdim];
    const T *__restrict__ in = reinterpret_cast<const T *>(p.in) + in_base;
    T *__restrict__ out = reinterpret_cast<T *>(p.out) + out_base;
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;  // 64 x 4
    // read: lanes run along tdim (input unit stride)
#pragma unroll
    for (int j = 0; j < TT; j += 4) {
        const int64_t r = i0 + ty + j, c = it + tx;
        if (r < n0 && c < nt) tile[ty + j][tx] = in[r * p.in_stride[0] + c];
    }
    __syncthreads();
    // write: lanes run along dim 0 (output unit stride)
#pragma unroll
    for (int j = 0; j < TT; j += 4) {
        const int64_t c = it + ty + j, r = i0 + tx;
        if (r < n0 && c < nt) out[c * p.out_stride[p.tdim] + r] = tile[tx][ty + j];
    }
}

// 16-byte variant: every global access is a full 128-bit vector (along tdim on the way in, along dim 0 on the way
// out); the element transpose happens in shared memory.  Needs 16-byte aligned bases and vector-multiple strides.
template <typename T>
__global__ void __launch_bounds__(256) transpose_vec_kernel(const TransposePlan p, const int64_t tiles0, const int64_t tilesT) {
    pdl_enter();
    constexpr int VEC = 16 / sizeof(T);
    constexpr int PAD = sizeof(T) >= 4 ? 1 : 4 / sizeof(T);  // odd row pitch in 32-bit words
    constexpr int TPR = TT / VEC;                            // threads per tile row
    constexpr int RPP = 256 / TPR;                           // tile rows per pass
    __shared__ T tile[TT][TT + PAD];
    int64_t bid = blockIdx.x;
    const int64_t t0 = bid % tiles0;
    bid /= tiles0;
    const int64_t tt = bid % tilesT;
    bid /= tilesT;
    int64_t in_base = 0, out_base = 0;
    for (int d = 1; d < p.ndim; ++d) {
        if (d == p.tdim) continue;
        const int64_t i = bid % p.shape[d];
        bid /= p.shape[d];
        in_base += i * p.in_stride[d];
        out_base += i * p.out_stride[d];
    }
    const int64_t i0 = t0 * TT, it = tt * TT;
    const int64_t n0 = p.shape[0], nt = p.shape[p.tdim];
    const T *__restrict__ in = reinterpret_cast<const T *>(p.in) + in_base;
    T *__restrict__ out = reinterpret_cast<T *>(p.out) + out_base;
    const int vx = threadIdx.x % TPR, vy = threadIdx.x / TPR;
    const int64_t is0 = p.in_stride[0], ost = p.out_stride[p.tdim];
#pragma unroll
    for (int j = 0; j < TT; j += RPP) {
        const int64_t r = i0 + vy + j, c = it + vx * VEC;
        if (r < n0 && c < nt) {  // n0, nt are vector multiples: a vector is never partially out of range
            Pack<T, VEC> v;
            *reinterpret_cast<uint4 *>(&v) = __ldg(reinterpret_cast<const uint4 *>(in + r * is0 + c));
#pragma unroll
            for (int k = 0; k < VEC; ++k) tile[vy + j][vx * VEC + k] = v.v[k];
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < TT; j += RPP) {
        const int64_t c = it + vy + j, r = i0 + vx * VEC;
        if (c < nt && r < n0) {
            Pack<T, VEC> v;
#pragma unroll
            for (int k = 0; k < VEC; ++k) v.v[k] = tile[vx * VEC + k][vy + j];
            *reinterpret_cast<uint4 *>(out + c * ost + r) = *reinterpret_cast<uint4 *>(&v);
        }
    }
}

static bool transpose_vec_ok(const TransposePlan &p) {
    const int64_t vec = 16 / p.itemsize;
    if (reinterpret_cast<uintptr_t>(p.in) % 16 || reinterpret_cast<uintptr_t>(p.out) % 16) return false;
    if (p.shape[0] % vec || p.shape[p.tdim] % vec) return false;
    for (int d = 0; d < p.ndim; ++d) {
        if (d != p.tdim && p.in_stride[d] % vec) return false;
        if (d != 0 && p.out_stride[d] % vec) return false;
    }
    return true;
}

void launch_transpose(const TransposePlan &p) {
    Runtime &rt = Runtime::get();
    const int64_t tiles0 = (p.shape[0] + TT - 1) / TT, tilesT = (p.shape[p.tdim] + TT - 1) / TT;
    int64_t batch = 1;
    for (int d = 1; d < p.ndim; ++d)
        if (d != p.tdim) batch *= p.shape[d];
    const int64_t grid = tiles0 * tilesT * batch;
    if (grid == 0) return;
    KF_CHECK(grid < (int64_t)0x7FFFFFFF, "transpose grid too large");
    if (transpose_vec_ok(p)) {
        switch (p.itemsize) {
        case 1: launch_pdl(transpose_vec_kernel<uint8_t>, dim3((unsigned)grid), dim3(256), 0, rt.stream(), p, tiles0, tilesT); break;
        case 2: launch_pdl(transpose_vec_kernel<uint16_t>, dim3((unsigned)grid), dim3(256), 0, rt.stream(), p, tiles0, tilesT); break;
        case 4: launch_pdl(transpose_vec_kernel<uint32_t>, dim3((unsigned)grid), dim3(256), 0, rt.stream(), p, tiles0, tilesT); break;
        default: launch_pdl(transpose_vec_kernel<uint64_t>, dim3((unsigned)grid), dim3(256), 0, rt.stream(), p, tiles0, tilesT); break;
        }
        rt.post_launch("transpose_vec_kernel");
        return;
    }
    switch (p.itemsize) {
    case 1: launch_pdl(transpose_tiled_kernel<uint8_t>, dim3((unsigned)grid), dim3(256), 0, rt.stream(), p, tiles0, tilesT); break;
    case 2: launch_pdl(transpose_tiled_kernel<uint16_t>, dim3((unsigned)grid), dim3(256), 0, rt.stream(), p, tiles0, tilesT); break;
    case 4: launch_pdl(transpose_tiled_kernel<uint32_t>, dim3((unsigned)grid), dim3(256), 0, rt.stream(), p, tiles0, tilesT); break;
    default: launch_pdl(transpose_tiled_kernel<uint64_t>, dim3((unsigned)grid), dim3(256), 0, rt.stream(), p, tiles0, tilesT); break;
    }
    rt.post_launch("transpose_tiled_kernel");
}

}  // namespace kf

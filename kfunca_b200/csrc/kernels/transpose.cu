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

void launch_transpose(const TransposePlan &p) {
    Runtime &rt = Runtime::get();
    const int64_t tiles0 = (p.shape[0] + TT - 1) / TT, tilesT = (p.shape[p.tdim] + TT - 1) / TT;
    int64_t batch = 1;
    for (int d = 1; d < p.ndim; ++d)
        if (d != p.tdim) batch *= p.shape[d];
    const int64_t grid = tiles0 * tilesT * batch;
    if (grid == 0) return;
    KF_CHECK(grid < (int64_t)0x7FFFFFFF, "transpose grid too large");
    switch (p.itemsize) {
    case 1: transpose_tiled_kernel<uint8_t><<<(unsigned)grid, 256, 0, rt.stream()>>>(p, tiles0, tilesT); break;
    case 2: transpose_tiled_kernel<uint16_t><<<(unsigned)grid, 256, 0, rt.stream()>>>(p, tiles0, tilesT); break;
    case 4: transpose_tiled_kernel<uint32_t><<<(unsigned)grid, 256, 0, rt.stream()>>>(p, tiles0, tilesT); break;
    default: transpose_tiled_kernel<uint64_t><<<(unsigned)grid, 256, 0, rt.stream()>>>(p, tiles0, tilesT); break;
    }
    rt.post_launch("transpose_tiled_kernel");
}

}  // namespace kf

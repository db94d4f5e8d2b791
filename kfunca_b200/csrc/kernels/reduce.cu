// sm_100a sum / mean over one axis of a dense [outer, R, inner] tensor (keepdim handled by the caller).
// Replaces the reference's gpu_reduce_kernel / ReduceOp family (src/device/utils/tensor_reduce.h:122-1083;
// functors src/device/reduce_ops_kernel.cu:6-59).  Design:
//   inner == 1 (row reduce)   : many rows -> W in {1,2,4,8} warps per row (8 / W rows per CTA), 128-bit streaming
//                               loads with 8 in flight per lane, shuffle tree;
//                               few long rows -> row split over S CTAs; a thread-block CLUSTER of up to
//                               8 CTAs folds its partials through distributed shared memory (DSMEM), so
//                               only S/8 partials per row ever touch HBM (none when S <= 8).
//   inner  > 1 (column reduce): lanes run along `inner` (coalesced, 128-bit), warps + cluster CTAs split R,
//                               combined through shared memory and DSMEM.
// Every reduction is ONE launch: when an output needs more than one cluster, the cluster partials go to a
// scratch buffer and the last cluster to arrive (library-owned, zero-initialised, self-resetting counters —
// the reference's un-zeroed semaphore hazard, SURVEY F10, cannot occur) folds them in cluster order, so
// the summation order is fixed by the launch geometry => deterministic run to run.
// Accumulation: fp32 for fp16/bf16/fp32 (documented deviation from the reference's accumulate-in-input-dtype,
// SURVEY F6), fp64 for fp64, int64 for integers/bool (truncation on store == the reference's wrap-around).
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "ew_common.cuh"
#include "tc_common.cuh"

namespace cg = cooperative_groups;

namespace kf {

struct ReduceArgs {
    const void *in;
    void *out;          // final output (Tout)
    void *partial;      // [groups][nparts][...] accumulator-typed partials (split kernels, nparts > 1)
    uint32_t *counter;  // one self-resetting arrival counter per output group (split kernels, nparts > 1)
    int64_t rows;       // row kernels: number of rows; col kernels: outer
    int64_t R;
    int64_t inner;
    int64_t chunk;      // elements (rows) of R handled by one CTA
    int S;              // splits of R
    int C;              // cluster size along the split
    int push;           // column kernel: fold the cluster by pushing partial rows into rank 0's shared memory
    int is_mean;
    double factor_f;
    int64_t factor_i;
};

// Programmatic dependent launch: let the NEXT kernel on the stream start scheduling its CTAs while this one drains, and do not
// touch global memory before every EARLIER kernel has completed and flushed.  Both are no-ops for a launch without the attribute.
__device__ __forceinline__ void pdl_prologue() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

template <typename A>
__device__ __forceinline__ A finalize(A acc, const ReduceArgs &a) {
    if (!a.is_mean) return acc;
    if constexpr (std::is_same<A, int64_t>::value) return acc * a.factor_i;
    else return acc * (A)a.factor_f;
}

template <typename A>
__device__ __forceinline__ A warp_sum(A v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// streaming 16-byte load that does not allocate in L1 (every input byte is read exactly once)
template <typename T, int VEC>
__device__ __forceinline__ Pack<T, VEC> ld_stream(const Pack<T, VEC> *p) {
    if constexpr (sizeof(Pack<T, VEC>) == 16) {
        uint4 r;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
        Pack<T, VEC> out;
        *reinterpret_cast<uint4 *>(&out) = r;
        return out;
    } else {
        return *p;
    }
}

// sum of nv vectors starting at pv[first], stride `step` vectors, 8 independent loads in flight per thread
template <typename Tin, typename A, int VEC>
__device__ __forceinline__ A strided_vec_sum(const Pack<Tin, VEC> *pv, int64_t first, int64_t step, int64_t nv) {
    A acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = A(0);
    int64_t i = first;
    for (; i + 7 * step < nv; i += 8 * step) {
        Pack<Tin, VEC> pk[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) pk[u] = ld_stream<Tin, VEC>(pv + i + u * step);
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int j = 0; j < VEC; ++j) acc[j] += cvt_in<A>(pk[u].v[j]);
    }
    for (; i < nv; i += step) {
        Pack<Tin, VEC> pk = ld_stream<Tin, VEC>(pv + i);
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] += cvt_in<A>(pk.v[j]);
    }
    A tot = A(0);
#pragma unroll
    for (int j = 0; j < VEC; ++j) tot += acc[j];
    return tot;
}

// ------------------------------------------------------------------ rows: W warps per row, 8 / W rows per CTA
template <typename Tin, typename Tout, typename A, int VEC, int W>
__global__ void __launch_bounds__(256) reduce_rows_kernel(const ReduceArgs a) {
    __shared__ A part[8];
    pdl_prologue();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wr = warp % W;                                  // this warp's slot inside its row
    const int64_t row = (int64_t)blockIdx.x * (8 / W) + warp / W;
    A acc = A(0);
    if (row < a.rows) {
        const Tin *__restrict__ p = reinterpret_cast<const Tin *>(a.in) + row * a.R;
        if constexpr (VEC > 1) {
            acc = strided_vec_sum<Tin, A, VEC>(reinterpret_cast<const Pack<Tin, VEC> *>(p), wr * 32 + lane, W * 32, a.R / VEC);
        } else {
#pragma unroll 4
            for (int64_t i = wr * 32 + lane; i < a.R; i += W * 32) acc += cvt_in<A>(p[i]);
        }
    }
    acc = warp_sum(acc);
    if constexpr (W == 1) {
        if (lane == 0 && row < a.rows) reinterpret_cast<Tout *>(a.out)[row] = cvt_out<Tout, A>(finalize(acc, a));
    } else {
        if (lane == 0) part[warp] = acc;
        __syncthreads();
        if (threadIdx.x < 8 / W) {
            const int64_t r = (int64_t)blockIdx.x * (8 / W) + threadIdx.x;
            if (r < a.rows) {
                A v = A(0);
#pragma unroll
                for (int k = 0; k < W; ++k) v += part[threadIdx.x * W + k];
                reinterpret_cast<Tout *>(a.out)[r] = cvt_out<Tout, A>(finalize(v, a));
            }
        }
    }
}

// Cross-CTA finish shared by the split kernels.  Cluster rank 0 of each cluster has the cluster's partial in
// `mine` (thread-strided over `width` slots).  With more than one cluster per output group the partials go to
// global memory and the LAST cluster to arrive (self-resetting counter) folds them in cluster order — the
// summation order is fixed, so the result is deterministic.  Returns true for the threads that must store.
template <typename A>
__device__ __forceinline__ bool publish_and_elect(const ReduceArgs &a, A *group_partials, int part_idx, int nparts, int width, const A *mine_smem,
                                                  uint32_t *counter) {
    __shared__ uint32_t s_ticket;
    for (int t = threadIdx.x; t < width; t += blockDim.x) group_partials[(int64_t)part_idx * width + t] = mine_smem[t];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = atomicAdd(counter, 1u);
    __syncthreads();
    if (s_ticket != (uint32_t)(nparts - 1)) return false;
    if (threadIdx.x == 0) *counter = 0;  // re-arm for the next launch (stream-ordered)
    __threadfence();
    return true;
}

// ------------------------------------------------------------------ rows: few long rows, R split over S CTAs
template <typename Tin, typename Tout, typename A, int VEC>
__global__ void __launch_bounds__(256) reduce_rows_split_kernel(const ReduceArgs a) {
    __shared__ A warp_part[8];
    __shared__ A cta_val;
    __shared__ A inbox[15];         // cluster rank 0: the totals pushed by the other CTAs of the cluster (see reduce_cols_kernel)
    __shared__ uint64_t inbox_bar;
    const bool push = a.C > 1 && a.push;
    if (push) {
        if (tc::cluster_ctarank() == 0 && threadIdx.x == 0) {
            tc::mbar_init(&inbox_bar, (uint32_t)(a.C - 1));
            tc::fence_barrier_init();
        }
        asm volatile("barrier.cluster.arrive.release;" ::: "memory");
    }
    pdl_prologue();
    const int s = blockIdx.x;
    const int64_t row = blockIdx.y;
    const int64_t lo = (int64_t)s * a.chunk;
    int64_t hi = lo + a.chunk;
    if (hi > a.R) hi = a.R;
    const Tin *__restrict__ p = reinterpret_cast<const Tin *>(a.in) + row * a.R;
    A acc = A(0);
    if (lo < hi) {
        if constexpr (VEC > 1) {  // host guarantees chunk % VEC == 0 and R % VEC == 0
            acc = strided_vec_sum<Tin, A, VEC>(reinterpret_cast<const Pack<Tin, VEC> *>(p + lo), threadIdx.x, 256, (hi - lo) / VEC);
        } else {
#pragma unroll 4
            for (int64_t i = lo + threadIdx.x; i < hi; i += 256) acc += cvt_in<A>(p[i]);
        }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        A v = threadIdx.x < 8 ? warp_part[threadIdx.x] : A(0);
        v = warp_sum(v);
        if (threadIdx.x == 0) cta_val = v;
    }
    if (push) {
        asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
        const uint32_t rank = tc::cluster_ctarank();
        if (rank != 0) {
            if (threadIdx.x == 0) {  // the thread that wrote cta_val
                cg::cluster_group cluster = cg::this_cluster();
                cluster.map_shared_rank(&inbox[0], 0)[rank - 1] = cta_val;
                tc::mbar_arrive_cluster(tc::mapa_u32(&inbox_bar, 0));
            }
            return;
        }
        if (threadIdx.x == 0) {
            tc::mbar_wait_cluster(&inbox_bar, 0);
            A total = A(0);
            total += cta_val;
            for (int r = 0; r < a.C - 1; ++r) total += inbox[r];
            cta_val = total;
        }
    } else if (a.C > 1) {
        cg::cluster_group cluster = cg::this_cluster();
        cluster.sync();  // every CTA's cta_val is written and visible cluster-wide
        if (cluster.block_rank() == 0 && threadIdx.x == 0) {
            A total = A(0);
            for (int r = 0; r < a.C; ++r) total += *cluster.map_shared_rank(&cta_val, r);  // DSMEM reads
            cta_val = total;
        }
        cluster.sync();  // peers keep their shared memory alive until rank 0 has read it
        if (cluster.block_rank() != 0) return;
    }
    __syncthreads();
    const int nparts = a.S / a.C;
    if (nparts > 1) {
        A *gp = reinterpret_cast<A *>(a.partial) + row * nparts;
        if (!publish_and_elect<A>(a, gp, s / a.C, nparts, 1, &cta_val, a.counter + row)) return;
        // fixed-shape parallel fold of the cluster partials: thread t owns partials t, t+256, ...; shuffle tree; 8 warp sums
        A v = A(0);
        for (int k = threadIdx.x; k < nparts; k += 256) v += __ldcg(gp + k);
        v = warp_sum(v);
        if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            A total = A(0);
#pragma unroll
            for (int k = 0; k < 8; ++k) total += warp_part[k];
            cta_val = total;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) reinterpret_cast<Tout *>(a.out)[row] = cvt_out<Tout, A>(finalize(cta_val, a));
}

// ------------------------------------------------------------------ columns
// LPR lanes run along `inner` (16-byte vectors), so a warp covers 32 / LPR rows per load instruction and a CTA tile is
// LPR * VEC columns wide: narrow tiles give enough CTAs for the whole reduction to finish inside ONE 8-CTA cluster
// (DSMEM fold, no global hand-shake) even when `inner` is only a few thousand columns.
template <typename Tin, typename Tout, typename A, int VEC, int LPR, int U = 8>
__global__ void __launch_bounds__(256) reduce_cols_kernel(const ReduceArgs a) {
    constexpr int RPW = 32 / LPR;        // rows per warp-wide load
    constexpr int PH = 8 * RPW;          // row phases per CTA
    constexpr int WIDTH = LPR * VEC;     // columns per CTA
    __shared__ A part[PH][WIDTH];
    constexpr bool PUSH_OK = 15 * WIDTH * sizeof(A) <= 24576;  // wide 1-byte inputs with 64-bit accumulators keep the pull fold
    __shared__ A inbox[PUSH_OK ? 15 : 1][WIDTH];  // cluster rank 0 only: the folded partial row of every other CTA of the cluster
    __shared__ uint64_t inbox_bar;      // ... and the mbarrier their (remote) arrivals complete
    const bool push = PUSH_OK && a.C > 1 && a.push;
    if (push) {
        // rank 0 arms its inbox; the cluster barrier is only ARRIVED at here and waited for after the streaming loop, so it
        // costs nothing: it orders the mbarrier init (and "rank 0 is running") before the first remote store
        if (tc::cluster_ctarank() == 0 && threadIdx.x == 0) {
            tc::mbar_init(&inbox_bar, (uint32_t)(a.C - 1) * WIDTH);
            tc::fence_barrier_init();
        }
        asm volatile("barrier.cluster.arrive.release;" ::: "memory");
    }
    pdl_prologue();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int cl = lane % LPR, ph = w * RPW + lane / LPR;
    const int64_t col = ((int64_t)blockIdx.x * LPR + cl) * VEC;
    const int s = blockIdx.y;
    const int64_t o = blockIdx.z;
    const int64_t lo = (int64_t)s * a.chunk;
    int64_t hi = lo + a.chunk;
    if (hi > a.R) hi = a.R;
    A acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = A(0);
    if (col < a.inner) {
        const Tin *__restrict__ p = reinterpret_cast<const Tin *>(a.in) + o * a.R * a.inner + col;
        int64_t r = lo + ph;
        for (; r + (U - 1) * PH < hi; r += U * PH) {  // U independent 16-byte loads in flight per thread
            Pack<Tin, VEC> pk[U];
#pragma unroll
            for (int u = 0; u < U; ++u) pk[u] = ld_stream<Tin, VEC>(reinterpret_cast<const Pack<Tin, VEC> *>(p + (r + (int64_t)PH * u) * a.inner));
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int j = 0; j < VEC; ++j) acc[j] += cvt_in<A>(pk[u].v[j]);
        }
        for (; r < hi; r += PH) {
            Pack<Tin, VEC> pk = ld_stream<Tin, VEC>(reinterpret_cast<const Pack<Tin, VEC> *>(p + r * a.inner));
#pragma unroll
            for (int j = 0; j < VEC; ++j) acc[j] += cvt_in<A>(pk.v[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) part[ph][cl * VEC + j] = acc[j];
    __syncthreads();
    // fold the row phases: thread t owns column slot t (WIDTH <= 256 slots)
    for (int t = threadIdx.x; t < WIDTH; t += 256) {
        A v = A(0);
#pragma unroll
        for (int k = 0; k < PH; ++k) v += part[k][t];
        part[0][t] = v;
    }
    if (push) {
        // push fold: every CTA but rank 0 writes its row straight into rank 0's inbox (st.shared::cluster) and arrives on rank 0's
        // mbarrier with release semantics, then exits; rank 0 waits once and adds the rows in rank order (deterministic).  One
        // DSMEM hop instead of cluster.sync + remote reads + cluster.sync.
        asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
        const uint32_t rank = tc::cluster_ctarank();
        if (rank != 0) {
            cg::cluster_group cluster = cg::this_cluster();
            A *dst = cluster.map_shared_rank(&inbox[0][0], 0) + (rank - 1) * WIDTH;
            const uint32_t rbar = tc::mapa_u32(&inbox_bar, 0);
            for (int t = threadIdx.x; t < WIDTH; t += 256) {
                dst[t] = part[0][t];  // written by this same thread in the loop above
                tc::mbar_arrive_cluster(rbar);
            }
            return;
        }
        tc::mbar_wait_cluster(&inbox_bar, 0);
        for (int t = threadIdx.x; t < WIDTH; t += 256) {
            A v0 = part[0][t], v1 = A(0), v2 = A(0), v3 = A(0);
            int r = 0;
            for (; r + 3 < a.C - 1; r += 4) {
                v1 += inbox[r][t];
                v2 += inbox[r + 1][t];
                v3 += inbox[r + 2][t];
                v0 += inbox[r + 3][t];
            }
            for (; r < a.C - 1; ++r) v1 += inbox[r][t];
            part[0][t] = (v0 + v1) + (v2 + v3);
        }
    } else if (a.C > 1) {
        cg::cluster_group cluster = cg::this_cluster();
        cluster.sync();
        if (cluster.block_rank() == 0) {
            for (int t = threadIdx.x; t < WIDTH; t += 256) {
                A v0 = part[0][t], v1 = A(0), v2 = A(0), v3 = A(0);
                int r = 1;
                for (; r + 3 < a.C; r += 4) {  // independent DSMEM reads in flight
                    v1 += cluster.map_shared_rank(&part[0][0], r)[t];
                    v2 += cluster.map_shared_rank(&part[0][0], r + 1)[t];
                    v3 += cluster.map_shared_rank(&part[0][0], r + 2)[t];
                    v0 += cluster.map_shared_rank(&part[0][0], r + 3)[t];
                }
                for (; r < a.C; ++r) v1 += cluster.map_shared_rank(&part[0][0], r)[t];
                part[0][t] = (v0 + v1) + (v2 + v3);
            }
        }
        cluster.sync();
        if (cluster.block_rank() != 0) return;
    }
    __syncthreads();
    const int nparts = a.S / a.C;
    if (nparts > 1) {
        const int64_t group = o * gridDim.x + blockIdx.x;
        A *gp = reinterpret_cast<A *>(a.partial) + group * nparts * WIDTH;
        if (!publish_and_elect<A>(a, gp, s / a.C, nparts, WIDTH, &part[0][0], a.counter + group)) return;
        // parallel fold: 256 / WIDTH thread groups split the partials (fixed assignment => deterministic), 4 loads in flight
        constexpr int GROUPS = (256 / WIDTH) > 0 ? (256 / WIDTH) : 1;
        for (int t = threadIdx.x; t < WIDTH * GROUPS; t += 256) {
            const int slot = t % WIDTH, g = t / WIDTH;
            A v0 = A(0), v1 = A(0), v2 = A(0), v3 = A(0);
            int k = g;
            for (; k + 3 * GROUPS < nparts; k += 4 * GROUPS) {
                v0 += __ldcg(gp + (int64_t)k * WIDTH + slot);
                v1 += __ldcg(gp + (int64_t)(k + GROUPS) * WIDTH + slot);
                v2 += __ldcg(gp + (int64_t)(k + 2 * GROUPS) * WIDTH + slot);
                v3 += __ldcg(gp + (int64_t)(k + 3 * GROUPS) * WIDTH + slot);
            }
            for (; k < nparts; k += GROUPS) v0 += __ldcg(gp + (int64_t)k * WIDTH + slot);
            part[g][slot] = (v0 + v1) + (v2 + v3);
        }
        __syncthreads();
        if (GROUPS > 1) {
            for (int t = threadIdx.x; t < WIDTH; t += 256) {
                A v = part[0][t];
#pragma unroll
                for (int g = 1; g < GROUPS; ++g) v += part[g][t];
                part[0][t] = v;
            }
            __syncthreads();
        }
    }
    for (int t = threadIdx.x; t < WIDTH; t += 256) {
        const int64_t c = (int64_t)blockIdx.x * WIDTH + t;
        if (c < a.inner) reinterpret_cast<Tout *>(a.out)[o * a.inner + c] = cvt_out<Tout, A>(finalize(part[0][t], a));
    }
}


// ------------------------------------------------------------------ columns, streaming variant (outer == 1, rows of <= 1024 vectors)
// Every CTA owns a CONTIGUOUS band of whole rows (thread t keeps the columns of vectors t, t+256, ... in registers), so HBM sees
// long sequential bursts instead of 256-byte pieces at a 16 KB stride; 16 independent 16-byte loads in flight per thread.
// The grid is at most two CTAs per SM, i.e. co-resident, which makes a software grid barrier legal: partial rows go to an
// L2-resident scratch, all CTAs meet, then CTA c folds its own slice of columns over all partials in a fixed order
// (deterministic).  Two self-resetting counters (arrive / depart) are reused by every launch on the stream.
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <typename Tin, typename Tout, typename A, int VEC, int NVT>
__global__ void __launch_bounds__(256, 2) reduce_cols_stream_kernel(const ReduceArgs a) {
    constexpr int UR = 16 / NVT;  // rows per batch: UR * NVT = 16 loads in flight
    __shared__ A fold[256];
    pdl_prologue();
    const int tid = threadIdx.x;
    const int64_t nvec = a.inner / VEC;
    const int64_t lo = (int64_t)blockIdx.x * a.chunk;
    int64_t hi = lo + a.chunk;
    if (hi > a.R) hi = a.R;
    A acc[NVT][VEC];
#pragma unroll
    for (int j = 0; j < NVT; ++j)
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[j][i] = A(0);
    const Pack<Tin, VEC> *__restrict__ base = reinterpret_cast<const Pack<Tin, VEC> *>(a.in);
    int64_t r = lo;
    for (; r + UR <= hi; r += UR) {
        Pack<Tin, VEC> pk[UR][NVT];
#pragma unroll
        for (int u = 0; u < UR; ++u)
#pragma unroll
            for (int j = 0; j < NVT; ++j)
                if (tid + j * 256 < nvec) pk[u][j] = ld_stream<Tin, VEC>(base + (r + u) * nvec + tid + j * 256);
#pragma unroll
        for (int u = 0; u < UR; ++u)
#pragma unroll
            for (int j = 0; j < NVT; ++j)
                if (tid + j * 256 < nvec)
#pragma unroll
                    for (int i = 0; i < VEC; ++i) acc[j][i] += cvt_in<A>(pk[u][j].v[i]);
    }
    for (; r < hi; ++r) {
#pragma unroll
        for (int j = 0; j < NVT; ++j)
            if (tid + j * 256 < nvec) {
                Pack<Tin, VEC> pk = ld_stream<Tin, VEC>(base + r * nvec + tid + j * 256);
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc[j][i] += cvt_in<A>(pk.v[i]);
            }
    }
    // partial row of this CTA -> scratch (stays in L2)
    A *__restrict__ mine = reinterpret_cast<A *>(a.partial) + (int64_t)blockIdx.x * a.inner;
#pragma unroll
    for (int j = 0; j < NVT; ++j)
        if (tid + j * 256 < nvec) {
            Pack<A, VEC> out;
#pragma unroll
            for (int i = 0; i < VEC; ++i) out.v[i] = acc[j][i];
            *reinterpret_cast<Pack<A, VEC> *>(mine + ((int64_t)tid + j * 256) * VEC) = out;
        }
    // grid barrier
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        atomicAdd(a.counter, 1u);
        const long long t0 = clock64();
        while (ld_acquire_u32(a.counter) < gridDim.x) {
            if (clock64() - t0 > 4000000000ll) {
                printf("kfunca_b200: reduce grid barrier timed out (block %d)\n", (int)blockIdx.x);
                __trap();
            }
        }
    }
    __syncthreads();
    // fold: this CTA owns columns [c0, c0 + cw); thread = (partial group g, column c); groups stride over the partials
    const int P = (int)gridDim.x;
    const int cw = (int)((a.inner + P - 1) / P);  // host guarantees cw <= 256
    int cwp = 1;
    while (cwp < cw) cwp <<= 1;
    const int groups = 256 / cwp;
    const int64_t c0 = (int64_t)blockIdx.x * cw;
    const int c = tid % cwp, g = tid / cwp;
    const bool live = c < cw && c0 + c < a.inner;
    const A *__restrict__ all = reinterpret_cast<const A *>(a.partial) + c0 + c;
    A v0 = A(0), v1 = A(0), v2 = A(0), v3 = A(0);
    if (live) {
        int k = g;
        for (; k + 3 * groups < P; k += 4 * groups) {
            v0 += __ldcg(all + (int64_t)k * a.inner);
            v1 += __ldcg(all + (int64_t)(k + groups) * a.inner);
            v2 += __ldcg(all + (int64_t)(k + 2 * groups) * a.inner);
            v3 += __ldcg(all + (int64_t)(k + 3 * groups) * a.inner);
        }
        for (; k < P; k += groups) v0 += __ldcg(all + (int64_t)k * a.inner);
    }
    fold[tid] = (v0 + v1) + (v2 + v3);
    __syncthreads();
    if (tid < cw && c0 + tid < a.inner) {
        A v = fold[tid];
        for (int q = 1; q < groups; ++q) v += fold[q * cwp + tid];
        reinterpret_cast<Tout *>(a.out)[c0 + tid] = cvt_out<Tout, A>(finalize(v, a));
    }
    // depart: the last CTA re-arms both counters for the next launch on the stream
    if (tid == 0) {
        const uint32_t t = atomicAdd(a.counter + 1, 1u);
        if (t == gridDim.x - 1) {
            a.counter[0] = 0;
            a.counter[1] = 0;
            __threadfence();
        }
    }
}

// ------------------------------------------------------------------ host side
static bool pdl_enabled() {
    const char *e = std::getenv("KF_PDL");  // default on; KF_PDL=0 switches programmatic dependent launch off
    return !(e && e[0] == '0');
}

template <typename K>
static void launch_clustered(K kernel, dim3 grid, dim3 cluster, const ReduceArgs &args, const char *name) {
    Runtime &rt = Runtime::get();
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = rt.stream();
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (cluster.x * cluster.y * cluster.z > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = cluster.x;
        attr[na].val.clusterDim.y = cluster.y;
        attr[na].val.clusterDim.z = cluster.z;
        ++na;
    }
    if (pdl_enabled()) {  // see pdl_prologue()
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    if (cluster.x * cluster.y * cluster.z > 8) {  // 16-CTA clusters are "non-portable": opt in once per kernel
        static bool opted = false;
        if (!opted) {
            KF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            opted = true;
        }
    }
    KF_CUDA(cudaLaunchKernelEx(&cfg, kernel, args));
    rt.post_launch(name);
}

constexpr int kMaxCounters = 8192;
// zero-initialised, self-resetting arrival counters shared by every split reduction (stream-ordered reuse)
static uint32_t *arrival_counters() {
    static uint32_t *buf = nullptr;
    if (!buf) {
        Runtime &rt = Runtime::get();
        KF_CUDA(cudaMalloc(&buf, kMaxCounters * sizeof(uint32_t)));
        rt.memset_async(buf, 0, kMaxCounters * sizeof(uint32_t));
    }
    return buf;
}

static int pick_splits(int64_t base_ctas, int64_t R, int64_t min_rows_per_cta, int max_splits, int ctas_per_sm = 8) {
    const int64_t target = (int64_t)Runtime::get().props().sm_count * ctas_per_sm;
    int S = 1;
    while (S < max_splits && base_ctas * S < target && R / (S * 2) >= min_rows_per_cta) S *= 2;
    return S;
}

template <typename Tin, typename Tout, typename A, int V>
static void launch_rows_w(const ReduceArgs &a, int W, bool vec_ok) {
    Runtime &rt = Runtime::get();
    const int64_t rows_per_cta = 8 / W;
    const int64_t g = (a.rows + rows_per_cta - 1) / rows_per_cta;
    KF_CHECK(g < (int64_t)0x7FFFFFFF);
    const unsigned grid = (unsigned)g;
#define KF_ROWS(WW)                                                                                                      \
    do {                                                                                                                 \
        if (vec_ok) launch_clustered(reduce_rows_kernel<Tin, Tout, A, V, WW>, dim3(grid), dim3(1, 1, 1), a, "reduce_rows_kernel"); \
        else launch_clustered(reduce_rows_kernel<Tin, Tout, A, 1, WW>, dim3(grid), dim3(1, 1, 1), a, "reduce_rows_kernel");        \
    } while (0)
    switch (W) {
    case 1: KF_ROWS(1); break;
    case 2: KF_ROWS(2); break;
    case 4: KF_ROWS(4); break;
    default: KF_ROWS(8); break;
    }
#undef KF_ROWS
    (void)rt;
}

template <typename Tin, typename Tout, typename A>
static void reduce_rows(const void *in, void *out, int64_t rows, int64_t R, const ReducePlan &pl, int64_t factor_i) {
    constexpr int V = 16 / sizeof(Tin);
    const bool vec_ok = ((uintptr_t)in % 16 == 0) && (R % V == 0);
    ReduceArgs a{};
    a.in = in; a.out = out; a.rows = rows; a.R = R; a.inner = 1;
    a.is_mean = pl.is_mean; a.factor_f = pl.factor; a.factor_i = factor_i;
    a.S = 1; a.C = 1; a.chunk = R;
    Runtime &rt = Runtime::get();
    const int64_t sms = rt.props().sm_count;
    // W warps per row: as many as keep >= 8 vector loads per lane, until the grid has >= 8 CTAs per SM
    int W = 1;
    while (W < 8 && (rows + (8 / W) - 1) / (8 / W) < sms * 8 && R / (W * 2) >= (int64_t)32 * V * 8) W *= 2;
    if (const char *e = std::getenv("KF_RED_W")) W = std::atoi(e);  // tuning hook
    const int64_t ctas = (rows + (8 / W) - 1) / (8 / W);
    if (ctas >= sms * 2 || R < (int64_t)256 * V * 16) {  // enough rows (or rows too short to split): no cross-CTA step
        launch_rows_w<Tin, Tout, A, V>(a, W, vec_ok);
        return;
    }
    // few long rows: split R over S CTAs; clusters of up to 8 fold through DSMEM, the last cluster finishes
    KF_CHECK(rows <= 65535, "row count too large for the split reduce");  // grid.y limit
    int S = pick_splits(rows, R, (int64_t)256 * V * 4, 1024);
    // measured (4096^2 fp32 full sum): 4 CTAs per SM with NO cluster stage and the last-arriving CTA folding all partials is the
    // fastest shape (13.9 us; 512 CTAs in clusters of 8: 15.2 us) — one global hand-shake beats a DSMEM stage plus a hand-shake
    int C = 1;
    {
        const int64_t per_row = std::min<int64_t>(1024, std::max<int64_t>(1, sms * 4 / rows));
        if (per_row > 1 && R / per_row >= (int64_t)256 * V * 4) S = (int)per_row;
    }
    if (const char *e = std::getenv("KF_RED_S")) S = std::atoi(e);
    if (const char *e = std::getenv("KF_RED_C")) C = std::min(S, std::atoi(e));
    int64_t chunk = (R + S - 1) / S;
    chunk = (chunk + V * 256 - 1) / (V * 256) * (V * 256);  // keep chunks vector- and block-aligned
    a.S = S; a.C = C; a.chunk = chunk;
    a.push = 1;
    if (const char *e = std::getenv("KF_RED_PUSH")) a.push = std::atoi(e);
    const int nparts = S / C;
    KF_CHECK(rows <= kMaxCounters);
    Scratch partial(nparts > 1 ? sizeof(A) * rows * nparts : 16);
    a.partial = partial.p;
    a.counter = arrival_counters();
    dim3 grid((unsigned)S, (unsigned)rows, 1), cluster((unsigned)C, 1, 1);
    if (vec_ok) launch_clustered(reduce_rows_split_kernel<Tin, Tout, A, V>, grid, cluster, a, "reduce_rows_split_kernel");
    else launch_clustered(reduce_rows_split_kernel<Tin, Tout, A, 1>, grid, cluster, a, "reduce_rows_split_kernel");
}

// streaming column reduce (see reduce_cols_stream_kernel); false when the shape does not qualify
template <typename Tin, typename Tout, typename A>
static bool reduce_cols_stream(const void *in, void *out, int64_t R, int64_t inner, const ReducePlan &pl, int64_t factor_i) {
    constexpr int V = 16 / sizeof(Tin);
    // opt-in (KF_RED_STREAM=1): measured 17.0 us at 4096 x 4096 fp32 against 13.3 us for the cluster kernel — the grid barrier and
    // the 293-partial fold cost more than the sequential rows gain
    const char *on = std::getenv("KF_RED_STREAM");
    if (!on || on[0] != '1') return false;
    if ((uintptr_t)in % 16 != 0 || inner % V != 0) return false;
    const int64_t nvec = inner / V;
    if (nvec < 256 || nvec > 1024) return false;                       // every thread busy, <= 4 vectors per thread per row
    if (R * inner * (int64_t)sizeof(Tin) < (int64_t)(8 << 20)) return false;  // small problems: the cluster kernel has less fixed cost
    Runtime &rt = Runtime::get();
    const int nvt = (int)((nvec + 255) / 256);
    // the grid barrier needs every CTA resident at once: ask the driver how many CTAs of THIS instantiation fit on an SM
    // (register / shared-memory limits differ per dtype) and never launch more than that
    static int occ_cache[3] = {-1, -1, -1};
    const int oi = nvt == 1 ? 0 : (nvt == 2 ? 1 : 2);
    if (occ_cache[oi] < 0) {
        int occ = 0;
        const void *fn = nvt == 1 ? (const void *)reduce_cols_stream_kernel<Tin, Tout, A, V, 1>
                       : nvt == 2 ? (const void *)reduce_cols_stream_kernel<Tin, Tout, A, V, 2>
                                  : (const void *)reduce_cols_stream_kernel<Tin, Tout, A, V, 4>;
        KF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, 256, 0));
        occ_cache[oi] = occ;
    }
    int per_sm = 2;
    if (const char *e = std::getenv("KF_RED_STREAM_CTAS")) per_sm = std::atoi(e);
    if (per_sm > occ_cache[oi]) per_sm = occ_cache[oi];
    if (per_sm < 1) return false;
    int64_t G = (int64_t)rt.props().sm_count * per_sm;  // co-resident by construction
    if (R < G * 4) return false;
    const int64_t chunk = (R + G - 1) / G;
    G = (R + chunk - 1) / chunk;
    if ((inner + G - 1) / G > 256) return false;
    static uint32_t *counters = nullptr;
    if (!counters) {
        KF_CUDA(cudaMalloc(&counters, 2 * sizeof(uint32_t)));
        rt.memset_async(counters, 0, 2 * sizeof(uint32_t));
    }
    ReduceArgs a{};
    a.in = in; a.out = out; a.rows = 1; a.R = R; a.inner = inner; a.chunk = chunk;
    a.is_mean = pl.is_mean; a.factor_f = pl.factor; a.factor_i = factor_i;
    a.S = (int)G; a.C = 1;
    Scratch partial(sizeof(A) * (size_t)G * (size_t)inner);
    a.partial = partial.p;
    a.counter = counters;
    dim3 grid((unsigned)G), one(1, 1, 1);
    switch (nvt) {
    case 1: launch_clustered(reduce_cols_stream_kernel<Tin, Tout, A, V, 1>, grid, one, a, "reduce_cols_stream_kernel"); break;
    case 2: launch_clustered(reduce_cols_stream_kernel<Tin, Tout, A, V, 2>, grid, one, a, "reduce_cols_stream_kernel"); break;
    default: launch_clustered(reduce_cols_stream_kernel<Tin, Tout, A, V, 4>, grid, one, a, "reduce_cols_stream_kernel"); break;
    }
    return true;
}

template <typename Tin, typename Tout, typename A>
static void reduce_cols(const void *in, void *out, int64_t outer, int64_t R, int64_t inner, const ReducePlan &pl, int64_t factor_i) {
    constexpr int V = 16 / sizeof(Tin);
    const bool vec_ok = ((uintptr_t)in % 16 == 0) && (inner % V == 0);
    const int vec = vec_ok ? V : 1;
    ReduceArgs a{};
    a.in = in; a.out = out; a.rows = outer; a.R = R; a.inner = inner;
    a.is_mean = pl.is_mean; a.factor_f = pl.factor; a.factor_i = factor_i;
    KF_CHECK(outer <= 65535, "outer too large for the column reduce");
    if constexpr (!std::is_same<A, int64_t>::value) {  // floating point only: integer sums keep the cluster kernel
        if (outer == 1 && reduce_cols_stream<Tin, Tout, A>(in, out, R, inner, pl, factor_i)) return;
    }
    const int64_t sms = Runtime::get().props().sm_count;
    // lanes per row: narrow the CTA tile (32 -> 16 -> 8 lanes, never below one 128-byte line per row for 16-byte vectors)
    // while an 8-way split — the most one cluster can fold without a global hand-shake — would leave SMs idle
    int lpr = 32;
    auto tiles_for = [&](int l) { return (inner + (int64_t)l * vec - 1) / ((int64_t)l * vec); };
    // (measured at 4096 x 4096 fp32: ~3.5 CTAs per SM with 128 KB each beats 7 per SM with 64 KB: 14.4 us vs 18.3 us)
    while (lpr > 8 && tiles_for(lpr) * outer * 8 < sms * 3 && R >= 512) lpr /= 2;
    if (const char *e = std::getenv("KF_RED_LPR")) lpr = std::atoi(e);
    const int64_t tiles = tiles_for(lpr);
    const int ph = 8 * (32 / lpr);
    int S = pick_splits(tiles * outer, R, (int64_t)ph * 4, 256, 3);
    int C = S < 8 ? S : 8;
    if (const char *e = std::getenv("KF_RED_S")) S = std::atoi(e);  // tuning hooks (tools/gpu_tune_reduce.py)
    if (const char *e = std::getenv("KF_RED_C")) C = std::min(S, std::atoi(e));
    int nparts = S / C;
    if (nparts > 1 && tiles * outer > kMaxCounters) {  // cannot happen with the target above; stay safe
        S = C;
        nparts = 1;
    }
    a.S = S; a.C = C; a.chunk = (R + S - 1) / S;
    a.push = 1;
    if (const char *e = std::getenv("KF_RED_PUSH")) a.push = std::atoi(e);
    Scratch partial(nparts > 1 ? sizeof(A) * outer * tiles * nparts * lpr * vec : 16);
    a.partial = partial.p;
    a.counter = arrival_counters();
    dim3 grid((unsigned)tiles, (unsigned)S, (unsigned)outer), cluster(1, (unsigned)C, 1);
    // 16 loads in flight per thread for 4-byte types with long enough chunks (more bytes in flight per SM at ~3.5 CTAs / SM)
    bool deep = sizeof(Tin) == 4 && vec_ok && a.chunk >= (int64_t)ph * 32;
    if (const char *e = std::getenv("KF_RED_U")) deep = deep && std::atoi(e) == 16;
#define KF_COLS(L)                                                                                                        \
    do {                                                                                                                  \
        if constexpr (sizeof(Tin) == 4) {                                                                                 \
            if (deep) {                                                                                                   \
                launch_clustered(reduce_cols_kernel<Tin, Tout, A, V, L, 16>, grid, cluster, a, "reduce_cols_kernel");     \
                break;                                                                                                    \
            }                                                                                                             \
        }                                                                                                                 \
        if (vec_ok) launch_clustered(reduce_cols_kernel<Tin, Tout, A, V, L>, grid, cluster, a, "reduce_cols_kernel");     \
        else launch_clustered(reduce_cols_kernel<Tin, Tout, A, 1, L>, grid, cluster, a, "reduce_cols_kernel");            \
    } while (0)
    switch (lpr) {
    case 8: KF_COLS(8); break;
    case 16: KF_COLS(16); break;
    default: KF_COLS(32); break;
    }
#undef KF_COLS
}

template <typename T, typename A>
static void reduce_typed(const ReducePlan &p) {
    int64_t factor_i = 1;
    if (p.is_mean && std::is_same<A, int64_t>::value) factor_i = (int64_t)p.factor;  // caller pre-computed the integer factor
    if (p.inner == 1) {
        if constexpr (std::is_same<T, float>::value) {
            // One long fp32 row (full-tensor sum / mean): two launches instead of one split kernel — rows of L elements through the
            // plain row kernel (no cluster, no cross-CTA hand-shake: 0.91 of the HBM peak), then the R / L row sums (fp32, L2-resident)
            // through the same code; with programmatic dependent launch the second launch overlaps the first one's drain.  Measured
            // at 64 MiB: 12.9 us against 14.4 us for the split kernel with its last-CTA fold (KF_RED_TWOSTEP=0 keeps that one).
            const char *off = std::getenv("KF_RED_TWOSTEP");
            if (p.outer == 1 && p.R >= ((int64_t)1 << 22) && !(off && off[0] == '0')) {
                int64_t L = 0;
                for (int64_t cand : {8192, 4096, 2048, 1024})
                    if (p.R % cand == 0 && p.R / cand >= 1024) {
                        L = cand;
                        break;
                    }
                if (L > 0) {
                    const int64_t rows = p.R / L;
                    Scratch partial((size_t)rows * sizeof(float));
                    ReducePlan first = p, second = p;
                    first.is_mean = 0;
                    first.factor = 1.0;
                    reduce_rows<float, float, float>(p.in, partial.p, rows, L, first, 1);
                    reduce_rows<float, float, float>(partial.p, p.out, 1, rows, second, factor_i);  // applies the mean factor of the whole row
                    return;
                }
            }
        }
        reduce_rows<T, T, A>(p.in, p.out, p.outer, p.R, p, factor_i);
    } else {
        for (int64_t o0 = 0; o0 < p.outer; o0 += 65535) {
            const int64_t no = std::min<int64_t>(65535, p.outer - o0);
            reduce_cols<T, T, A>((const T *)p.in + o0 * p.R * p.inner, (T *)p.out + o0 * p.inner, no, p.R, p.inner, p, factor_i);
        }
    }
}

void launch_reduce(const ReducePlan &p) {
    if (p.outer * p.inner == 0) return;
    switch (p.dtype) {
    case KF_FLOAT: reduce_typed<float, float>(p); break;
    case KF_DOUBLE: reduce_typed<double, double>(p); break;
    case KF_HALF: reduce_typed<__half, float>(p); break;
    case KF_BFLOAT16: reduce_typed<__nv_bfloat16, float>(p); break;
    case KF_BOOL: reduce_typed<bool, int64_t>(p); break;
    case KF_BYTE: reduce_typed<uint8_t, int64_t>(p); break;
    case KF_CHAR: reduce_typed<int8_t, int64_t>(p); break;
    case KF_SHORT: reduce_typed<int16_t, int64_t>(p); break;
    case KF_INT: reduce_typed<int32_t, int64_t>(p); break;
    case KF_LONG: reduce_typed<int64_t, int64_t>(p); break;
    default: KF_CHECK(false, "Unsupported ScalarType ", p.dtype);
    }
}

}  // namespace kf

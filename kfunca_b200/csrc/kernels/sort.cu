// sm_100a stable sort and top-k along dense rows.
// Replaces the reference's BlockRadixSort / upsweep-scan-downsweep family and topk_with_sort
// (src/device/sort_ops_kernel.cu:12-632, src/device/utils/sorting_radix_sort.h:7-905).
//
// Ordering contract (bit-exact with the reference): keys are mapped to order-preserving unsigned integers
// (KeyTraits, src/device/utils/sorting_common.h:39-238); the sort is stable; "descending" is a stable sort of
// the complemented key, so ties always keep ascending original index; indices are int64.
//
//   rows <= 4096 elements : one CTA per row, bitonic network in shared memory on the composite (key, index)
//                           — the index makes every element unique, so the network is deterministic & stable.
//   longer rows           : LSD radix sort, 8-bit digits (half the passes of the reference's 4-bit digits):
//                           per pass  tile histogram -> per-row exclusive scan -> stable scatter, where the
//                           in-tile stable rank comes from warp match_any + per-warp running digit counts.
//   top-k (k <= 1024, 2048 <= n <= 32768): ONE pass over HBM.  The whole row lives in registers (32 keys x 1024
//                           threads); a lower bound T0 of the k-th key is derived from per-thread maxima
//                           (warp w contributes its ceil(k/32)-th largest thread-max, T0 = min over warps, which
//                           guarantees >= k elements >= T0); elements >= T0 (a few hundred) are compacted to shared
//                           memory, rank-sorted by (key desc, index asc) and the first k written.  Rows whose
//                           candidate set overflows (massive ties) are redone exactly by a radix-select kernel.
#include <algorithm>
#include <cstdlib>

#include "ew_common.cuh"

namespace kf {

enum KeyKind { KEY_UINT = 0, KEY_SINT = 1, KEY_FLOAT = 2 };

template <typename U>
__device__ __forceinline__ U key_convert(U x, int kind, bool descending) {
    constexpr int B = sizeof(U) * 8;
    const U sign = U(1) << (B - 1);
    U k;
    if (kind == KEY_FLOAT) k = (x & sign) ? U(~x) : U(x | sign);
    else if (kind == KEY_SINT) k = x ^ sign;
    else k = x;
    return descending ? U(~k) : k;
}
template <typename U>
__device__ __forceinline__ U key_deconvert(U k, int kind, bool descending) {
    constexpr int B = sizeof(U) * 8;
    const U sign = U(1) << (B - 1);
    if (descending) k = U(~k);
    if (kind == KEY_FLOAT) return (k & sign) ? U(k ^ sign) : U(~k);
    if (kind == KEY_SINT) return k ^ sign;
    return k;
}

// ------------------------------------------------------------------------------------------------
// small rows: bitonic sort of (key, idx) in shared memory
// ------------------------------------------------------------------------------------------------
template <typename U>
__global__ void sort_rows_bitonic_kernel(const U *__restrict__ in, U *__restrict__ values, int64_t *__restrict__ indices,
                                         const int n, const int N, const int kind, const bool descending) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    U *sk = reinterpret_cast<U *>(smem_raw);
    uint32_t *si = reinterpret_cast<uint32_t *>(smem_raw + (size_t)N * sizeof(U));
    const int64_t row = blockIdx.x;
    const U *src = in + row * n;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        if (i < n) {
            sk[i] = key_convert<U>(src[i], kind, descending);
            si[i] = (uint32_t)i;
        } else {
            sk[i] = U(~U(0));
            si[i] = 0x80000000u + (uint32_t)i;  // pads sort after every real element, even on equal keys
        }
    }
    __syncthreads();
    for (int k = 2; k <= N; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (N >> 1); t += blockDim.x) {
                const int i = 2 * t - (t & (j - 1));
                const int l = i + j;
                const bool up = (i & k) == 0;
                const U ki = sk[i], kl = sk[l];
                const uint32_t ii = si[i], il = si[l];
                const bool gt = ki > kl || (ki == kl && ii > il);
                if (gt == up) {
                    sk[i] = kl; sk[l] = ki;
                    si[i] = il; si[l] = ii;
                }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        values[row * n + i] = key_deconvert<U>(sk[i], kind, descending);
        indices[row * n + i] = (int64_t)si[i];
    }
}

// ------------------------------------------------------------------------------------------------
// long rows: LSD radix sort, 8-bit digits
// ------------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 4096 keys per CTA

struct RadixArgs {
    const void *keys_in;   // U keys (pass 0: raw input bits)
    const uint32_t *idx_in;
    void *keys_out;        // U keys, or raw value bits on the last pass
    uint32_t *idx_out;
    int64_t *idx_out64;    // last pass
    uint32_t *hist;        // [nseg][256][tiles]
    int n, tiles, shift, kind;
    bool descending, first, last;
};

template <typename U>
__device__ __forceinline__ U radix_load_key(const RadixArgs &a, int64_t base, int i) {
    const U raw = reinterpret_cast<const U *>(a.keys_in)[base + i];
    return a.first ? key_convert<U>(raw, a.kind, a.descending) : raw;
}

template <typename U>
__global__ void __launch_bounds__(RS_THREADS) radix_hist_kernel(const RadixArgs a) {
    __shared__ uint32_t hist[256];
    const int tile = blockIdx.x;
    const int64_t seg = blockIdx.y;
    hist[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = seg * a.n;
    const int lo = tile * RS_TILE;
    const int lane = threadIdx.x & 31;
#pragma unroll 4
    for (int it = 0; it < RS_ITEMS; ++it) {
        const int i = lo + it * RS_THREADS + threadIdx.x;
        uint32_t d = 256;
        if (i < a.n) d = (uint32_t)((radix_load_key<U>(a, base, i) >> a.shift) & 0xff);
        const uint32_t m = __match_any_sync(0xffffffffu, d);
        if (d < 256 && lane == __ffs(m) - 1) atomicAdd(&hist[d], __popc(m));
    }
    __syncthreads();
    a.hist[(seg * 256 + threadIdx.x) * a.tiles + tile] = hist[threadIdx.x];
}

// exclusive scan of hist[seg][digit][tile] in (digit, tile) order, in place; one CTA per row
__global__ void __launch_bounds__(1024) radix_scan_kernel(uint32_t *__restrict__ hist, const int len) {
    __shared__ uint32_t warp_tot[32];
    uint32_t *h = hist + (int64_t)blockIdx.x * len;
    const int chunk = (len + 1023) / 1024;
    const int lo = threadIdx.x * chunk;
    const int hi = min(lo + chunk, len);
    uint32_t s = 0;
    for (int i = lo; i < hi; ++i) s += h[i];
    // block exclusive scan of s
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    if (lane == 31) warp_tot[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint32_t v = warp_tot[lane], vi = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, vi, o);
            if (lane >= o) vi += u;
        }
        warp_tot[lane] = vi - v;
    }
    __syncthreads();
    uint32_t run = warp_tot[w] + inc - s;
    for (int i = lo; i < hi; ++i) {
        const uint32_t c = h[i];
        h[i] = run;
        run += c;
    }
}

template <typename U>
__global__ void __launch_bounds__(RS_THREADS) radix_scatter_kernel(const RadixArgs a) {
    __shared__ uint32_t wcnt[RS_WARPS][257];
    __shared__ uint32_t gbase[256];
    const int tile = blockIdx.x;
    const int64_t seg = blockIdx.y;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * 257; i += RS_THREADS) (&wcnt[0][0])[i] = 0;
    gbase[threadIdx.x] = a.hist[(seg * 256 + threadIdx.x) * a.tiles + tile];
    __syncthreads();
    const int64_t base = seg * a.n;
    // warp w owns the contiguous chunk [lo, lo + 32*RS_ITEMS); order inside the tile is (warp, item, lane)
    const int lo = tile * RS_TILE + w * (32 * RS_ITEMS);
    U key[RS_ITEMS];
    uint32_t idx[RS_ITEMS];
    uint32_t loc[RS_ITEMS];  // rank among same-digit keys of this warp
#pragma unroll
    for (int it = 0; it < RS_ITEMS; ++it) {
        const int i = lo + it * 32 + lane;
        uint32_t d = 256;
        key[it] = 0;
        idx[it] = 0;
        if (i < a.n) {
            key[it] = radix_load_key<U>(a, base, i);
            idx[it] = a.first ? (uint32_t)i : a.idx_in[base + i];
            d = (uint32_t)((key[it] >> a.shift) & 0xff);
        }
        const uint32_t m = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(m) - 1;
        uint32_t old = 0;
        if (lane == leader) {
            old = wcnt[w][d];
            wcnt[w][d] = old + __popc(m);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        loc[it] = old + __popc(m & ((1u << lane) - 1));
        __syncwarp();
    }
    __syncthreads();
    // exclusive prefix over warps, per digit (thread t owns digit t), folded with the global tile base
    {
        uint32_t run = gbase[threadIdx.x];
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ++ww) {
            const uint32_t c = wcnt[ww][threadIdx.x];
            wcnt[ww][threadIdx.x] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < RS_ITEMS; ++it) {
        const int i = lo + it * 32 + lane;
        if (i < a.n) {
            const uint32_t d = (uint32_t)((key[it] >> a.shift) & 0xff);
            const int64_t pos = base + wcnt[w][d] + loc[it];
            if (a.last) {
                reinterpret_cast<U *>(a.keys_out)[pos] = key_deconvert<U>(key[it], a.kind, a.descending);
                a.idx_out64[pos] = (int64_t)idx[it];
            } else {
                reinterpret_cast<U *>(a.keys_out)[pos] = key[it];
                a.idx_out[pos] = idx[it];
            }
        }
    }
}

template <typename U>
static void sort_rows_typed(const void *in, void *values, int64_t *indices, int kind, int64_t nseg, int64_t n, bool descending) {
    Runtime &rt = Runtime::get();
    cudaStream_t st = rt.stream();
    if (n <= 4096) {
        int N = 32;
        while (N < n) N <<= 1;
        const int threads = std::max(32, std::min(N / 2, 1024));
        const size_t smem = (size_t)N * (sizeof(U) + 4);
        static bool attr_set = false;
        if (!attr_set) {
            KF_CUDA(cudaFuncSetAttribute(sort_rows_bitonic_kernel<U>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 12));
            attr_set = true;
        }
        KF_CHECK(nseg < (int64_t)0x7FFFFFFF);
        sort_rows_bitonic_kernel<U><<<(unsigned)nseg, threads, smem, st>>>((const U *)in, (U *)values, indices, (int)n, N, kind, descending);
        rt.post_launch("sort_rows_bitonic_kernel");
        return;
    }
    const int tiles = (int)((n + RS_TILE - 1) / RS_TILE);
    const int64_t total = nseg * n;
    Scratch kbuf0(total * sizeof(U)), kbuf1(total * sizeof(U)), ibuf0(total * 4), ibuf1(total * 4);
    const int passes = (int)sizeof(U);
    for (int64_t s0 = 0; s0 < nseg; s0 += 32768) {  // grid.y limit
        const int64_t ns = std::min<int64_t>(32768, nseg - s0);
        Scratch hist((size_t)ns * 256 * tiles * 4);
        void *kb[2] = {(char *)kbuf0.p + s0 * n * sizeof(U), (char *)kbuf1.p + s0 * n * sizeof(U)};
        uint32_t *ib[2] = {ibuf0.as<uint32_t>() + s0 * n, ibuf1.as<uint32_t>() + s0 * n};
        for (int p = 0; p < passes; ++p) {
            RadixArgs a{};
            a.first = p == 0;
            a.last = p == passes - 1;
            a.keys_in = a.first ? (const void *)((const char *)in + s0 * n * sizeof(U)) : kb[(p + 1) & 1];
            a.idx_in = ib[(p + 1) & 1];
            a.keys_out = a.last ? (void *)((char *)values + s0 * n * sizeof(U)) : kb[p & 1];
            a.idx_out = ib[p & 1];
            a.idx_out64 = indices + s0 * n;
            a.hist = hist.as<uint32_t>();
            a.n = (int)n;
            a.tiles = tiles;
            a.shift = 8 * p;
            a.kind = kind;
            a.descending = descending;
            dim3 grid((unsigned)tiles, (unsigned)ns);
            radix_hist_kernel<U><<<grid, RS_THREADS, 0, st>>>(a);
            rt.post_launch("radix_hist_kernel");
            radix_scan_kernel<<<(unsigned)ns, 1024, 0, st>>>(hist.as<uint32_t>(), 256 * tiles);
            rt.post_launch("radix_scan_kernel");
            radix_scatter_kernel<U><<<grid, RS_THREADS, 0, st>>>(a);
            rt.post_launch("radix_scatter_kernel");
        }
    }
}

static void key_class(int dtype, int &bytes, int &kind) {
    switch (dtype) {
    case KF_BYTE: bytes = 1; kind = KEY_UINT; break;
    case KF_CHAR: bytes = 1; kind = KEY_SINT; break;
    case KF_SHORT: bytes = 2; kind = KEY_SINT; break;
    case KF_INT: bytes = 4; kind = KEY_SINT; break;
    case KF_LONG: bytes = 8; kind = KEY_SINT; break;
    case KF_HALF: case KF_BFLOAT16: bytes = 2; kind = KEY_FLOAT; break;
    case KF_FLOAT: bytes = 4; kind = KEY_FLOAT; break;
    case KF_DOUBLE: bytes = 8; kind = KEY_FLOAT; break;
    default: KF_CHECK(false, "Sort currently does not support ", dtype_name(dtype), " dtypes.");
    }
}

void launch_sort_rows(const void *in, void *values, int64_t *indices, int dtype, int64_t nseg, int64_t n, bool descending) {
    if (nseg == 0 || n == 0) return;
    int bytes, kind;
    key_class(dtype, bytes, kind);
    switch (bytes) {
    case 1: sort_rows_typed<uint8_t>(in, values, indices, kind, nseg, n, descending); break;
    case 2: sort_rows_typed<uint16_t>(in, values, indices, kind, nseg, n, descending); break;
    case 4: sort_rows_typed<uint32_t>(in, values, indices, kind, nseg, n, descending); break;
    default: sort_rows_typed<uint64_t>(in, values, indices, kind, nseg, n, descending); break;
    }
}

// ------------------------------------------------------------------------------------------------
// top-k: register-resident single-pass select
// ------------------------------------------------------------------------------------------------
constexpr int TK_THREADS = 1024;
constexpr int TK_CAP = 2048;  // candidate capacity in shared memory

template <typename U>
__device__ __forceinline__ bool cand_greater(U ka, uint32_t ia, U kb, uint32_t ib) {  // (key desc, index asc) order
    return ka > kb || (ka == kb && ia < ib);
}

// sort the c candidates in (ck, ci) and write the first k; all threads of the CTA participate
template <typename U>
__device__ void emit_topk(U *ck, uint32_t *ci, const int c, const int k, U *__restrict__ out_v, int64_t *__restrict__ out_i,
                          const int kind, const bool desc_key) {
    if (c <= TK_THREADS) {
        // rank sort: candidate t counts how many candidates precede it; broadcast shared-memory reads
        if ((int)threadIdx.x < c) {
            const U mk = ck[threadIdx.x];
            const uint32_t mi = ci[threadIdx.x];
            int rank = 0;
            for (int j = 0; j < c; ++j) rank += cand_greater<U>(ck[j], ci[j], mk, mi) ? 1 : 0;
            if (rank < k) {
                out_v[rank] = key_deconvert<U>(mk, kind, desc_key);
                out_i[rank] = (int64_t)mi;
            }
        }
        return;
    }
    int N = 2048;  // c in (1024, 2048]
    for (int i = c + threadIdx.x; i < N; i += blockDim.x) {
        ck[i] = 0;
        ci[i] = 0xffffffffu;  // pads are the smallest possible candidates
    }
    __syncthreads();
    for (int kk = 2; kk <= N; kk <<= 1) {
        for (int j = kk >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (N >> 1); t += blockDim.x) {
                const int i = 2 * t - (t & (j - 1));
                const int l = i + j;
                const bool up = (i & kk) == 0;  // "up" = candidate order (best first)
                const U ki = ck[i], kl = ck[l];
                const uint32_t ii = ci[i], il = ci[l];
                const bool l_first = cand_greater<U>(kl, il, ki, ii);
                if (l_first == up) {
                    ck[i] = kl; ck[l] = ki;
                    ci[i] = il; ci[l] = ii;
                }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        out_v[i] = key_deconvert<U>(ck[i], kind, desc_key);
        out_i[i] = (int64_t)ci[i];
    }
}

template <typename U, int ITEMS, bool VEC4>
__global__ void __launch_bounds__(TK_THREADS, 1)
topk_select_kernel(const U *__restrict__ in, U *__restrict__ values, int64_t *__restrict__ indices, const int n, const int k,
                   const int kind, const bool largest, int *__restrict__ overflow_rows) {
    __shared__ U ck[TK_CAP];
    __shared__ uint32_t ci[TK_CAP];
    __shared__ U warp_thr[32];
    __shared__ int count;
    const int64_t row = blockIdx.x;
    const U *src = in + row * n;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    // descending order on keys == "largest"; for smallest we complement so that the same code selects maxima
    const bool desc_key = !largest;  // key_convert(.., descending=true) complements
    U key[ITEMS];
    if (tid == 0) count = 0;
    if constexpr (VEC4) {
#pragma unroll
        for (int i = 0; i < ITEMS / 4; ++i) {
            const int p = (i * TK_THREADS + tid) * 4;
            uint4 v = make_uint4(0, 0, 0, 0);
            const bool ok = p < n;  // n % 4 == 0 on this path
            if (ok) v = *reinterpret_cast<const uint4 *>(src + p);
            key[4 * i + 0] = ok ? key_convert<U>((U)v.x, kind, desc_key) : U(0);
            key[4 * i + 1] = ok ? key_convert<U>((U)v.y, kind, desc_key) : U(0);
            key[4 * i + 2] = ok ? key_convert<U>((U)v.z, kind, desc_key) : U(0);
            key[4 * i + 3] = ok ? key_convert<U>((U)v.w, kind, desc_key) : U(0);
        }
    } else {
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const int p = i * TK_THREADS + tid;
            key[i] = p < n ? key_convert<U>(src[p], kind, desc_key) : U(0);
        }
    }
    U tmax = 0;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) tmax = key[i] > tmax ? key[i] : tmax;
    // j-th largest thread-max of this warp, j = ceil(k / 32): strict total order by (value, lane)
    const int j = (k + 31) >> 5;
    int rank = 0;
#pragma unroll
    for (int l = 0; l < 32; ++l) {
        const U v = __shfl_sync(0xffffffffu, tmax, l);
        rank += (v > tmax || (v == tmax && l < lane)) ? 1 : 0;
    }
    if (rank == j - 1) warp_thr[w] = tmax;
    __syncthreads();
    U t0 = warp_thr[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const U v = __shfl_xor_sync(0xffffffffu, t0, o);
        t0 = v < t0 ? v : t0;
    }
    // compact everything >= t0 (at least k elements by construction)
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        if (key[i] >= t0) {
            const int p = VEC4 ? ((i / 4) * TK_THREADS + tid) * 4 + (i & 3) : i * TK_THREADS + tid;
            if (p < n) {
                const int slot = atomicAdd(&count, 1);
                if (slot < TK_CAP) {
                    ck[slot] = key[i];
                    ci[slot] = (uint32_t)p;
                }
            }
        }
    }
    __syncthreads();
    const int c = count;
    if (c > TK_CAP) {  // massive ties: hand the row to the exact kernel
        if (tid == 0) overflow_rows[row] = 1;
        return;
    }
    emit_topk<U>(ck, ci, c, k, values + row * k, indices + row * k, kind, desc_key);
}

// exact radix-select top-k straight from global memory (any n, k <= TK_CAP); used for overflowed rows
template <typename U>
__global__ void __launch_bounds__(TK_THREADS, 1)
topk_exact_kernel(const U *__restrict__ in, U *__restrict__ values, int64_t *__restrict__ indices, const int64_t nrows, const int n,
                  const int k, const int kind, const bool largest, const int *__restrict__ only_rows) {
    __shared__ U ck[TK_CAP];
    __shared__ uint32_t ci[TK_CAP];
    __shared__ uint32_t hist[256];
    __shared__ uint32_t scan_w[32];
    __shared__ U s_prefix;
    __shared__ int s_need, s_count, s_eq_taken;
    // rows are strided over the (persistent) grid; flags are scanned 1024 rows at a time so that the common
    // "nothing overflowed" case costs one global load and one barrier per CTA
    for (int64_t batch = blockIdx.x; batch < nrows; batch += (int64_t)gridDim.x * TK_THREADS) {
    if (only_rows) {
        const int64_t r = batch + (int64_t)threadIdx.x * gridDim.x;
        const int f = r < nrows ? only_rows[r] : 0;
        if (!__syncthreads_or(f)) continue;
    }
    for (int64_t row = batch; row < nrows && row < batch + (int64_t)gridDim.x * TK_THREADS; row += gridDim.x) {
    if (only_rows && !only_rows[row]) continue;  // block-uniform
    __syncthreads();
    const U *src = in + row * n;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const bool desc_key = !largest;
    constexpr int B = sizeof(U) * 8;
    U prefix = 0;     // decided high bits of the k-th largest key
    int need = k;     // rank (1-based) of the wanted key among keys matching the decided prefix
    for (int shift = B - 8; shift >= 0; shift -= 8) {
        if (tid < 256) hist[tid] = 0;
        __syncthreads();
        const U hi_mask = (shift + 8 >= B) ? U(0) : U(~U(0)) << (shift + 8);
        for (int i = tid; i < n; i += TK_THREADS) {
            const U kx = key_convert<U>(src[i], kind, desc_key);
            if ((kx & hi_mask) == (prefix & hi_mask)) atomicAdd(&hist[(uint32_t)((kx >> shift) & 0xff)], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            int acc = 0, d = 255;
            for (; d > 0; --d) {
                if (acc + (int)hist[d] >= need) break;
                acc += (int)hist[d];
            }
            s_prefix = prefix | (U(d) << shift);
            s_need = need - acc;
        }
        __syncthreads();
        prefix = s_prefix;
        need = s_need;
        __syncthreads();
    }
    // prefix == k-th largest key; `need` of the elements equal to it are wanted, lowest indices first
    if (tid == 0) {
        s_count = 0;
        s_eq_taken = 0;
    }
    __syncthreads();
    for (int base = 0; base < n; base += TK_THREADS) {
        const int i = base + tid;
        U kx = 0;
        bool gt = false, eq = false;
        if (i < n) {
            kx = key_convert<U>(src[i], kind, desc_key);
            gt = kx > prefix;
            eq = kx == prefix;
        }
        if (gt) {
            const int slot = atomicAdd(&s_count, 1);
            ck[slot] = kx;
            ci[slot] = (uint32_t)i;
        }
        // ordered ranking of the equal keys inside this chunk
        const uint32_t bal = __ballot_sync(0xffffffffu, eq);
        if (lane == 0) scan_w[w] = __popc(bal);
        __syncthreads();
        int before = s_eq_taken;
        for (int ww = 0; ww < w; ++ww) before += (int)scan_w[ww];
        const int my = before + __popc(bal & ((1u << lane) - 1));
        if (eq && my < need) {
            const int slot = atomicAdd(&s_count, 1);
            ck[slot] = kx;
            ci[slot] = (uint32_t)i;
        }
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int ww = 0; ww < 32; ++ww) tot += (int)scan_w[ww];
            s_eq_taken += tot;
        }
        __syncthreads();
    }
    __syncthreads();
    emit_topk<U>(ck, ci, s_count, k, values + row * k, indices + row * k, kind, desc_key);
    }
    }
}

// ------------------------------------------------------------------------------------------------
// top-k, 32-bit keys, k <= 256: two-pass select that reads HBM exactly once.
//   pass 1  streams the row (16-byte no-allocate loads, 8 in flight per thread) and keeps only the maximum of every
//           8-key group in registers; no key stays on chip, so a 256-thread CTA needs ~50 registers per thread and
//           several CTAs (rows) share an SM: while one row is being selected, the others keep HBM busy.
//   select  T1 = exact k-th largest THREAD maximum (warp knock-outs give a loose bound T0, the thread maxima >= T0
//           are rank-counted in shared memory).  At least k keys are >= T1, and only ~k + a few in expectation.
//   pass 2  re-reads ONLY the 8-key groups whose maximum reaches T1 (L2 hits: the row was read microseconds ago),
//           compacts the keys >= T1 and rank-sorts them on 64-bit composites (key << 32 | ~index): ties keep the
//           lowest index, exactly as the reference's stable sort does.
// ------------------------------------------------------------------------------------------------
constexpr int TP_THREADS = 256;
constexpr int TP_WARPS = TP_THREADS / 32;
constexpr int TP_CAP = 1024;  // candidate capacity; overflowing rows (massive ties) go to the exact kernel

// order-preserving key with the kind fixed at compile time (2 instructions per fp32 key); `flip` folds the complement in
template <int KIND>
__device__ __forceinline__ uint32_t key32(uint32_t x, uint32_t flip) {
    if (KIND == KEY_FLOAT) return x ^ (((uint32_t)((int32_t)x >> 31)) | 0x80000000u) ^ flip;
    if (KIND == KEY_SINT) return x ^ 0x80000000u ^ flip;
    return x ^ flip;
}
template <int KIND>
__device__ __forceinline__ uint32_t unkey32(uint32_t k, uint32_t flip) {
    k ^= flip;
    if (KIND == KEY_FLOAT) return (k & 0x80000000u) ? (k ^ 0x80000000u) : ~k;
    if (KIND == KEY_SINT) return k ^ 0x80000000u;
    return k;
}
__device__ __forceinline__ uint4 ldg_stream16(const uint32_t *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

template <int NB, int KIND>  // NB batches of 8 x 16 bytes per thread: n <= NB * 8192
__global__ void __launch_bounds__(TP_THREADS, 4)
topk_twopass_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ values, int64_t *__restrict__ indices, const int64_t nrows,
                    const int n, const int k, const bool largest, int *__restrict__ overflow_rows) {
    __shared__ __align__(16) unsigned long long cand[TP_CAP];
    __shared__ __align__(16) uint32_t tmax_list[TP_THREADS];
    __shared__ int rank_part[TP_THREADS];
    __shared__ uint32_t warp_thr[TP_WARPS];
    __shared__ uint32_t s_t1;
    __shared__ int count, count_tm;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const uint32_t flip = largest ? 0u : 0xffffffffu;  // smallest-k == largest-k of the complemented keys
    // Persistent CTA, rows strided over the grid.  While row r is being processed, the TMA engine pulls row r + grid
    // from HBM into L2 (cp.async.bulk.prefetch.L2): in steady state pass 1 streams from L2 and HBM never idles during
    // the selection phases.
    auto prefetch_row_l2 = [&](int64_t r) {
        const char *g = reinterpret_cast<const char *>(in + r * n);
        const uint32_t bytes = (uint32_t)n * 4u;
        for (uint32_t off = 0; off < bytes; off += 16384u) {
            const uint32_t sz = bytes - off < 16384u ? bytes - off : 16384u;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(g + off), "r"(sz) : "memory");
        }
    };
    for (int64_t row = blockIdx.x; row < nrows; row += gridDim.x) {
    const uint32_t *__restrict__ src = in + row * n;
    if (tid == 0) {
        count = 0;
        count_tm = 0;
        if (row + gridDim.x < nrows) prefetch_row_l2(row + gridDim.x);
    }
    // ---- pass 1: group maxima (group g = loads 2g, 2g+1 of this thread = 8 keys)
    uint32_t gmax[NB * 4];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int p = ((b * 8 + u) * TP_THREADS + tid) * 4;
            v[u] = make_uint4(0, 0, 0, 0);
            if (p < n) v[u] = ldg_stream16(src + p);  // n % 4 == 0 on this path
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int p = ((b * 8 + u) * TP_THREADS + tid) * 4;
            uint32_t m4 = max(max(key32<KIND>(v[u].x, flip), key32<KIND>(v[u].y, flip)), max(key32<KIND>(v[u].z, flip), key32<KIND>(v[u].w, flip)));
            if (p >= n) m4 = 0u;  // padding sits below every accepted threshold
            if (u & 1) gmax[b * 4 + (u >> 1)] = max(gmax[b * 4 + (u >> 1)], m4);
            else gmax[b * 4 + (u >> 1)] = m4;
        }
    }
    uint32_t tmax = 0;
#pragma unroll
    for (int g = 0; g < NB * 4; ++g) tmax = max(tmax, gmax[g]);
    // ---- loose lower bound T0 of the k-th largest thread maximum: every warp contributes its j-th largest lane value,
    // so >= 8 j >= k thread maxima are >= the minimum over warps
    {
        const int j = (k + TP_WARPS - 1) / TP_WARPS;
        uint32_t v = tmax;
        for (int it = 1; it < j; ++it) {  // knock out the current warp maximum (one lane) j - 1 times
            const uint32_t m = __reduce_max_sync(0xffffffffu, v);
            const uint32_t holders = __ballot_sync(0xffffffffu, v == m);
            if (lane == __ffs(holders) - 1) v = 0;
        }
        const uint32_t tw = __reduce_max_sync(0xffffffffu, v);
        if (lane == 0) warp_thr[w] = tw;
    }
    __syncthreads();
    const uint32_t t0 = __reduce_min_sync(0xffffffffu, warp_thr[lane & (TP_WARPS - 1)]);
    // ---- list of the thread maxima >= T0 (at least k of them)
    {
        const bool hit = t0 != 0 && tmax >= t0;
        const uint32_t bal = __ballot_sync(0xffffffffu, hit);
        if (bal) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&count_tm, __popc(bal));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (hit) tmax_list[base + __popc(bal & ((1u << lane) - 1u))] = tmax;
        }
    }
    __syncthreads();
    const int m = count_tm;
    // ---- T1 = exact k-th largest thread maximum.  Rank counting over all threads: entry e = tid % cp2 against slice
    // g = tid / cp2 (warp-uniform) of the list; earlier list positions win ties, so ranks are a permutation.
    if (t0 != 0) {
        int cp2 = 32;
        while (cp2 < m) cp2 <<= 1;
        const int G = TP_THREADS / cp2;
        const int e = tid & (cp2 - 1), g = tid / cp2;
        const int per = (m + G - 1) / G;
        const int lo = g * per, hi = min(m, lo + per);
        int rank = 0;
        uint32_t mine = 0;
        if (e < m) {
            mine = tmax_list[e];
            const int mid = min(max(e, lo), hi);
            const uint32_t ge_thr = mine - 1u;  // values are >= t0 >= 1: "a >= mine" is "a > mine - 1"
#pragma unroll 4
            for (int q = lo; q < mid; ++q) rank += tmax_list[q] > ge_thr ? 1 : 0;
#pragma unroll 4
            for (int q = mid; q < hi; ++q) rank += tmax_list[q] > mine ? 1 : 0;
        }
        rank_part[tid] = rank;
        __syncthreads();
        if (tid < m) {
            int r = 0;
            for (int gg = 0; gg < G; ++gg) r += rank_part[gg * cp2 + tid];
            if (r == k - 1) s_t1 = mine;
        }
    } else {
        if (tid == 0) s_t1 = 0u;
    }
    __syncthreads();
    const uint32_t t1 = s_t1;
    // ---- pass 2: re-read the groups that can hold a key >= T1, compact those keys.  Each lane walks ITS OWN hit groups
    // (bitmask + runtime address), so the L2 round trips of all lanes of a warp overlap instead of serialising per group.
    {
        uint32_t hm = 0;
        if (t1 != 0) {
#pragma unroll
            for (int g = 0; g < NB * 4; ++g) hm |= gmax[g] >= t1 ? (1u << g) : 0u;
        }
        while (__any_sync(0xffffffffu, hm != 0)) {
            if (hm) {
                const int g = __ffs(hm) - 1;
                hm &= hm - 1;
                const int p0 = (g * 2 * TP_THREADS + tid) * 4, p1 = p0 + TP_THREADS * 4;
                uint4 v0 = make_uint4(0, 0, 0, 0), v1 = make_uint4(0, 0, 0, 0);
                if (p0 < n) v0 = *reinterpret_cast<const uint4 *>(src + p0);
                if (p1 < n) v1 = *reinterpret_cast<const uint4 *>(src + p1);
                const uint32_t kk[8] = {key32<KIND>(v0.x, flip), key32<KIND>(v0.y, flip), key32<KIND>(v0.z, flip), key32<KIND>(v0.w, flip),
                                        key32<KIND>(v1.x, flip), key32<KIND>(v1.y, flip), key32<KIND>(v1.z, flip), key32<KIND>(v1.w, flip)};
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int p = (e < 4 ? p0 : p1) + (e & 3);
                    if (kk[e] >= t1 && p < n) {
                        const int slot = atomicAdd(&count, 1);
                        if (slot < TP_CAP) cand[slot] = ((unsigned long long)kk[e] << 32) | (unsigned long long)(0xffffffffu - (uint32_t)p);
                    }
                }
            }
        }
    }
    __syncthreads();
    const int c = t1 != 0 ? count : TP_CAP + 1;
    if (c > TP_CAP) {  // massive ties (or an all-padding threshold): the exact kernel redoes this row
        if (tid == 0) overflow_rows[row] = 1;
    } else if (c <= TP_THREADS) {  // ---- order the candidates: composites are unique, rank < k is the sorted top-k
        int cp2 = 32;
        while (cp2 < c) cp2 <<= 1;
        const int G = TP_THREADS / cp2;
        const int e = tid & (cp2 - 1), g = tid / cp2;
        const int per = (((c + G - 1) / G) + 1) & ~1;  // even slice length keeps the 16-byte reads aligned
        const int lo = g * per, hi = min(c, lo + per);
        const unsigned long long mine = cand[min(e, c - 1)];
        int rank = 0;
        if (e < c) {
            int q = lo;
#pragma unroll 4
            for (; q + 1 < hi; q += 2) {
                const ulonglong2 a = *reinterpret_cast<const ulonglong2 *>(&cand[q]);
                rank += (a.x > mine ? 1 : 0) + (a.y > mine ? 1 : 0);
            }
            if (q < hi) rank += cand[q] > mine ? 1 : 0;
        }
        rank_part[tid] = rank;
        __syncthreads();
        if (tid < c) {
            int r = 0;
            for (int gg = 0; gg < G; ++gg) r += rank_part[gg * cp2 + tid];
            if (r < k) {
                values[row * k + r] = unkey32<KIND>((uint32_t)(mine >> 32), flip);
                indices[row * k + r] = (int64_t)(0xffffffffu - (uint32_t)mine);
            }
        }
    } else {
        for (int ci = tid; ci < c; ci += TP_THREADS) {  // 256 < c <= 1024: rare, plain loop
            const unsigned long long mine = cand[ci];
            int rank = 0;
            for (int q = 0; q < c; ++q) rank += cand[q] > mine ? 1 : 0;
            if (rank < k) {
                values[row * k + rank] = unkey32<KIND>((uint32_t)(mine >> 32), flip);
                indices[row * k + rank] = (int64_t)(0xffffffffu - (uint32_t)mine);
            }
        }
    }
    __syncthreads();  // shared lists and counters are re-armed by the next row
    }
}

template <typename U>
static bool topk_typed(const void *in, void *values, int64_t *indices, int kind, int64_t nseg, int64_t n, int64_t k, bool largest) {
    Runtime &rt = Runtime::get();
    cudaStream_t st = rt.stream();
    KF_CHECK(nseg < (int64_t)0x7FFFFFFF);
    const unsigned grid = (unsigned)nseg;
    const unsigned pgrid = (unsigned)std::min<int64_t>(nseg, (int64_t)rt.props().sm_count * 2);  // persistent, rows strided
    const bool fast = k <= 1024 && n >= 2048 && n <= 32 * TK_THREADS;
    if (!fast) {
        // rows too long for registers: exact select when k fits the candidate buffer and rows are long enough to pay
        if (k <= 1024 && n > 32 * TK_THREADS) {
            topk_exact_kernel<U><<<pgrid, TK_THREADS, 0, st>>>((const U *)in, (U *)values, indices, nseg, (int)n, (int)k, kind, largest, nullptr);
            rt.post_launch("topk_exact_kernel");
            return true;
        }
        return false;
    }
    Scratch flags((size_t)nseg * sizeof(int));
    rt.memset_async(flags.p, 0, (size_t)nseg * sizeof(int));
    const bool vec = sizeof(U) == 4 && n % 4 == 0 && ((uintptr_t)in % 16 == 0);
    int items = 4;
    while ((int64_t)items * TK_THREADS < n) items *= 2;
    if (vec && k <= TP_THREADS / 2) {  // beyond k = 128 the thread-maximum threshold gets loose (k = 256 would accept ~4 % of the row)
        // 32-bit keys, 16-byte aligned rows: two-pass select, one 256-thread CTA per row, several rows per SM
        const int nb = n <= 8192 ? 1 : (n <= 16384 ? 2 : 4);
        // one CTA per row (hardware CTA scheduling balances the select phases best: measured 212 us vs 268 us for a
        // persistent grid with L2 prefetch at 8192 x 32768); KF_TOPK_CTAS_PER_SM > 0 selects the persistent variant
        static const int persist = std::getenv("KF_TOPK_CTAS_PER_SM") ? std::atoi(std::getenv("KF_TOPK_CTAS_PER_SM")) : 0;
        const unsigned tgrid = persist > 0 ? (unsigned)std::min<int64_t>(nseg, (int64_t)rt.props().sm_count * persist) : grid;
#define KF_TOPK_TP2(NBV, KD) \
    topk_twopass_kernel<NBV, KD><<<tgrid, TP_THREADS, 0, st>>>((const uint32_t *)in, (uint32_t *)values, indices, nseg, (int)n, (int)k, largest, flags.as<int>())
#define KF_TOPK_TP(NBV)                                      \
    do {                                                     \
        if (kind == KEY_FLOAT) KF_TOPK_TP2(NBV, KEY_FLOAT);  \
        else if (kind == KEY_SINT) KF_TOPK_TP2(NBV, KEY_SINT); \
        else KF_TOPK_TP2(NBV, KEY_UINT);                     \
    } while (0)
        switch (nb) {
        case 1: KF_TOPK_TP(1); break;
        case 2: KF_TOPK_TP(2); break;
        default: KF_TOPK_TP(4); break;
        }
#undef KF_TOPK_TP
#undef KF_TOPK_TP2
        rt.post_launch("topk_twopass_kernel");
    } else {
#define KF_TOPK_LAUNCH(IT)                                                                                                    \
    topk_select_kernel<U, IT, false><<<grid, TK_THREADS, 0, st>>>((const U *)in, (U *)values, indices, (int)n, (int)k, kind, largest, flags.as<int>())
        switch (items) {
        case 4: KF_TOPK_LAUNCH(4); break;
        case 8: KF_TOPK_LAUNCH(8); break;
        case 16: KF_TOPK_LAUNCH(16); break;
        default: KF_TOPK_LAUNCH(32); break;
        }
#undef KF_TOPK_LAUNCH
        rt.post_launch("topk_select_kernel");
    }
    topk_exact_kernel<U><<<pgrid, TK_THREADS, 0, st>>>((const U *)in, (U *)values, indices, nseg, (int)n, (int)k, kind, largest, flags.as<int>());
    rt.post_launch("topk_exact_kernel");
    return true;
}

bool launch_topk_rows(const void *in, void *values, int64_t *indices, int dtype, int64_t nseg, int64_t n, int64_t k, bool largest) {
    int bytes, kind;
    key_class(dtype, bytes, kind);
    switch (bytes) {
    case 1: return topk_typed<uint8_t>(in, values, indices, kind, nseg, n, k, largest);
    case 2: return topk_typed<uint16_t>(in, values, indices, kind, nseg, n, k, largest);
    case 4: return topk_typed<uint32_t>(in, values, indices, kind, nseg, n, k, largest);
    default: return topk_typed<uint64_t>(in, values, indices, kind, nseg, n, k, largest);
    }
}

}  // namespace kf

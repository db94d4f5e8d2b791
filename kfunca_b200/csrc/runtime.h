// Device runtime: one process per GPU, one explicit non-blocking compute stream, a stream-ordered
// HBM pool, and the kernel launcher every kernel goes through.
// Replaces the reference's Launcher singleton (src/device/launcher_cuda.h:105-354), memory engine
// (src/device/memory_engine.cu:6-28) and DeviceAllocator (src/core/device_allocator.cpp:13-78).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <vector>

#include "core.h"

namespace kf {

#define KF_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            ::kf::fail(__FILE__, __LINE__, __func__, ::kf::str("CUDA error in `" #expr "`: ", cudaGetErrorString(_e))); \
    } while (0)

// ------------------------------------------------------------------------------------------------
// Stream-ordered pool.  Memory is carved out of large arenas (cudaMalloc'ed once, 2 MiB granular)
// with best-fit + split + coalesce.  "Stream-ordered" = a freed block goes straight back to the free
// list with NO device synchronisation: kernels and copies of this library are issued on the one
// library stream, so any later user of the block is ordered after the last kernel that touched it.
// Work on OTHER streams (the copy streams of gemm_host, the NCCL communication stream, a caller's
// stream holding a zero-copy alias) must be announced with record_stream(ptr, stream): when such a
// block is released, the library stream first waits for everything enqueued so far on every
// recorded stream (event record + stream wait, no host synchronisation), so that the stream order
// of the library stream again covers every use of the memory.
// Host-visible reads (D2H) synchronise the stream themselves.
// The backing allocator is pluggable so the host logic is testable without a GPU.
// ------------------------------------------------------------------------------------------------
class Pool {
public:
    using RawAlloc = void *(*)(size_t bytes, void *ctx);
    using RawFree = void (*)(void *ptr, void *ctx);
    using Fence = void (*)(void *stream, void *ctx);  // make the library stream wait for `stream`'s work enqueued so far
    Pool(RawAlloc a, RawFree f, void *ctx, Fence fence = nullptr) : raw_alloc_(a), raw_free_(f), ctx_(ctx), fence_(fence) {}
    ~Pool();

    void *allocate(size_t bytes);
    void release(void *ptr);
    // `ptr` (any address inside a live block) is in use by work enqueued on `stream`, which is not the library stream
    void record_stream(const void *ptr, void *stream);
    int64_t fences_issued() const { return n_fence_; }
    void empty_cache();
    std::string report() const;

    int64_t bytes_in_use() const { return in_use_; }
    int64_t bytes_reserved() const { return reserved_; }
    int64_t arena_mallocs() const { return n_raw_; }

    static size_t round_size(size_t bytes);       // 512 B granularity (small), 2 MiB for >= 1 MiB
    static size_t arena_size_for(size_t rounded);  // how much to ask the driver for on a miss

private:
    struct Block {
        char *ptr;
        size_t size;
        bool free;
        Block *prev, *next;  // address-ordered neighbours inside the same arena
        char *arena;
        bool small;
        std::vector<void *> streams;  // side streams that touched the block since it was allocated (record_stream)
    };
    struct Cmp {
        bool operator()(const Block *a, const Block *b) const {
            return a->size != b->size ? a->size < b->size : a->ptr < b->ptr;
        }
    };
    std::set<Block *, Cmp> free_small_, free_large_;
    std::map<void *, Block *> live_;
    std::map<char *, size_t> arenas_;
    RawAlloc raw_alloc_;
    RawFree raw_free_;
    void *ctx_;
    Fence fence_ = nullptr;
    int64_t in_use_ = 0, reserved_ = 0, n_raw_ = 0, peak_ = 0, n_fence_ = 0;
    mutable std::mutex mu_;
};

struct DeviceProps {
    int sm_count = 0;
    int max_smem_optin = 0;
    int l2_bytes = 0;
    size_t total_mem = 0;
    int cc_major = 0, cc_minor = 0;
    char name[256] = {0};
};

class Runtime {
public:
    static Runtime &get();  // lazy; throws if there is no usable GPU (no CPU fallback)
    static bool initialised();
    static void select_device(int device);

    int device() const { return device_; }
    cudaStream_t stream() const { return stream_; }
    const DeviceProps &props() const { return props_; }
    Pool &pool() { return *pool_; }
    void sync() { KF_CUDA(cudaStreamSynchronize(stream_)); }
    void h2d(void *dst, const void *src, size_t bytes, bool sync_after);
    void d2h(void *dst, const void *src, size_t bytes, bool sync_after);
    void d2d(void *dst, const void *src, size_t bytes);
    void memset_async(void *dst, int v, size_t bytes);
    // the library stream waits (on the device) for everything enqueued so far on `other`
    void wait_for_stream(cudaStream_t other);
    // scratch that lives until the next call on the stream needs it: plain pool memory, freed stream-ordered
    void *scratch(size_t bytes) { return pool_->allocate(bytes); }
    void scratch_free(void *p) { pool_->release(p); }
    std::atomic<int64_t> launches{0};
    void post_launch(const char *what);  // counts + cudaGetLastError check (the reference never checks, launcher_cuda.h:341-351)

private:
    Runtime(int device);
    int device_;
    cudaStream_t stream_ = nullptr;
    DeviceProps props_;
    std::unique_ptr<Pool> pool_;
    std::mutex fence_mu_;
    std::vector<cudaEvent_t> fence_events_;  // reused round-robin: a stream wait captures the record it was enqueued after
    size_t fence_next_ = 0;
};

// RAII scratch buffer
struct Scratch {
    void *p;
    explicit Scratch(size_t bytes) : p(Runtime::get().scratch(bytes ? bytes : 16)) {}
    ~Scratch() { Runtime::get().scratch_free(p); }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

}  // namespace kf

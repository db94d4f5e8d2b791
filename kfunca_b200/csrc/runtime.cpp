#include "runtime.h"

#include <algorithm>
#include <cstdio>

namespace kf {

[[noreturn]] void fail(const char *file, int line, const char *func, const std::string &msg) {
    const char *base = std::strrchr(file, '/');
    throw Error(str("[enforce fail at ", base ? base + 1 : file, ":", line, ":", func, "] ", msg));
}

DType promote(DType a, DType b) {
    if (is_floating(a) && is_floating(b)) return a >= b ? a : b;
    if (is_floating(a) || is_floating(b)) return is_floating(a) ? a : b;
    if (is_unsigned_class(a) && is_unsigned_class(b)) return a >= b ? a : b;
    if (is_unsigned_class(a) || is_unsigned_class(b)) return is_unsigned_class(a) ? b : a;
    return a >= b ? a : b;
}

// ------------------------------------------------------------------------------------------ Pool
static constexpr size_t kSmallLimit = 1u << 20;   // requests <= 1 MiB live in the small pool
static constexpr size_t kSmallArena = 2u << 20;   // 2 MiB arenas for the small pool
static constexpr size_t kLargeArena = 20u << 20;  // 20 MiB arenas for 1..10 MiB requests
static constexpr size_t kMinLargeSplit = 1u << 20;
static constexpr size_t kGranule = 512;

size_t Pool::round_size(size_t bytes) {
    if (bytes < kGranule) return kGranule;
    return (bytes + kGranule - 1) / kGranule * kGranule;
}
size_t Pool::arena_size_for(size_t rounded) {
    if (rounded <= kSmallLimit) return kSmallArena;
    if (rounded < (10u << 20)) return kLargeArena;
    return (rounded + kSmallArena - 1) / kSmallArena * kSmallArena;
}

Pool::~Pool() {
    // Process teardown: the driver reclaims HBM; calling cudaFree from static destructors is unsafe.
    for (auto &kv : live_) delete kv.second;
    for (auto *b : free_small_) delete b;
    for (auto *b : free_large_) delete b;
}

void *Pool::allocate(size_t bytes) {
    std::unique_lock<std::mutex> g(mu_);
    const size_t size = round_size(bytes);
    const bool small = size <= kSmallLimit;
    auto &fl = small ? free_small_ : free_large_;
    Block key{nullptr, size, true, nullptr, nullptr, nullptr, small, {}};
    auto it = fl.lower_bound(&key);
    Block *b = nullptr;
    if (it != fl.end()) {  // best fit; the remainder is split off below and coalesces back on release
        b = *it;
        fl.erase(it);
    }
    if (!b) {
        // Large pool: once it holds >= 1 GiB, a miss asks the driver for at least a quarter of what is already reserved (capped at
        // 4 GiB) instead of the exact request.  Exact-size arenas never reach a steady state under a training step: best fit
        // carves small tensors out of big arenas, the next big request misses, and every miss is a synchronising cudaMalloc
        // (measured: 11 driver allocations inside 8 timed transformer-block steps).  Geometric slabs make the number of
        // misses logarithmic and leave slack for fragmentation; 180 GB of HBM pays for the <= 25 % head-room.
        const size_t base = arena_size_for(size);
        size_t asz = base;
        if (!small) {
            size_t grown = std::min<size_t>((size_t)reserved_ / 4, (size_t)4 << 30) / kSmallArena * kSmallArena;
            if (grown >= ((size_t)256 << 20)) asz = std::max(base, grown);
        }
        void *raw = raw_alloc_(asz, ctx_);
        if (!raw && asz != base) {  // no room for the slab: fall back to the exact request
            asz = base;
            raw = raw_alloc_(asz, ctx_);
        }
        if (!raw) {  // out of memory: give cached arenas back and retry once
            g.unlock();  // empty_cache() takes the mutex itself; unique_lock keeps the unlock exception-safe
            empty_cache();
            g.lock();
            raw = raw_alloc_(asz, ctx_);
            KF_CHECK(raw != nullptr, "out of device memory allocating ", asz, " bytes (in use ", in_use_, ", reserved ", reserved_, ")");
        }
        ++n_raw_;
        reserved_ += (int64_t)asz;
        arenas_[(char *)raw] = asz;
        b = new Block{(char *)raw, asz, true, nullptr, nullptr, (char *)raw, small, {}};
    }
    const size_t rem = b->size - size;
    if (rem >= (small ? kGranule : kMinLargeSplit)) {
        Block *r = new Block{b->ptr + size, rem, true, b, b->next, b->arena, small, {}};
        if (b->next) b->next->prev = r;
        b->next = r;
        b->size = size;
        fl.insert(r);
    }
    b->free = false;
    live_[b->ptr] = b;
    in_use_ += (int64_t)b->size;
    peak_ = std::max(peak_, in_use_);
    return b->ptr;
}

void Pool::release(void *ptr) {
    if (!ptr) return;
    std::lock_guard<std::mutex> g(mu_);
    auto it = live_.find(ptr);
    KF_CHECK(it != live_.end(), "pool: releasing unknown pointer");
    Block *b = it->second;
    live_.erase(it);
    in_use_ -= (int64_t)b->size;
    // side-stream users first: after these fences the library stream's order covers every use of the block
    for (void *st : b->streams) {
        if (fence_) fence_(st, ctx_);
        ++n_fence_;
    }
    b->streams.clear();
    b->free = true;
    auto &fl = b->small ? free_small_ : free_large_;
    if (b->prev && b->prev->free) {  // coalesce with the left neighbour
        Block *p = b->prev;
        fl.erase(p);
        p->size += b->size;
        p->next = b->next;
        if (b->next) b->next->prev = p;
        delete b;
        b = p;
    }
    if (b->next && b->next->free) {  // and with the right one
        Block *n = b->next;
        fl.erase(n);
        b->size += n->size;
        b->next = n->next;
        if (n->next) n->next->prev = b;
        delete n;
    }
    fl.insert(b);
}

void Pool::record_stream(const void *ptr, void *stream) {
    if (!ptr) return;
    std::lock_guard<std::mutex> g(mu_);
    auto it = live_.upper_bound(const_cast<void *>(ptr));  // first block starting after ptr; the owner is the one before it
    KF_CHECK(it != live_.begin(), "pool: record_stream on memory the pool does not own");
    --it;
    Block *b = it->second;
    KF_CHECK((const char *)ptr < b->ptr + b->size, "pool: record_stream on memory the pool does not own");
    if (std::find(b->streams.begin(), b->streams.end(), stream) == b->streams.end()) b->streams.push_back(stream);
}

void Pool::empty_cache() {
    std::lock_guard<std::mutex> g(mu_);
    for (auto *fl : {&free_small_, &free_large_}) {
        for (auto it = fl->begin(); it != fl->end();) {
            Block *b = *it;
            auto a = arenas_.find(b->ptr);
            if (!b->prev && !b->next && a != arenas_.end() && a->second == b->size) {
                raw_free_(b->ptr, ctx_);
                reserved_ -= (int64_t)b->size;
                arenas_.erase(a);
                it = fl->erase(it);
                delete b;
            } else {
                ++it;
            }
        }
    }
}

std::string Pool::report() const {
    std::lock_guard<std::mutex> g(mu_);
    std::ostringstream os;
    os << "kfunca_b200 stream-ordered pool: in_use=" << in_use_ << " B, reserved=" << reserved_ << " B, peak=" << peak_
       << " B, arenas=" << arenas_.size() << ", device_mallocs=" << n_raw_ << ", live_blocks=" << live_.size()
       << ", free_small=" << free_small_.size() << ", free_large=" << free_large_.size() << "\n";
    for (auto &a : arenas_) os << "  arena " << (void *)a.first << " " << a.second << " B\n";
    return os.str();
}

// --------------------------------------------------------------------------------------- Runtime
static std::mutex g_rt_mu;
static Runtime *g_rt = nullptr;
static int g_selected_device = -1;

static void *cuda_raw_alloc(size_t bytes, void *) {
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
static void cuda_raw_free(void *p, void *) {
    cudaStreamSynchronize(g_rt ? g_rt->stream() : nullptr);  // arenas leave the pool only when idle
    cudaFree(p);
}

static void cuda_fence(void *stream, void *) {
    if (g_rt) g_rt->wait_for_stream(reinterpret_cast<cudaStream_t>(stream));
}

bool Runtime::initialised() { return g_rt != nullptr; }

void Runtime::select_device(int device) {
    std::lock_guard<std::mutex> g(g_rt_mu);
    if (g_rt) {
        KF_CHECK(g_rt->device() == device, "kfunca_b200 is one-process-per-GPU: already bound to device ", g_rt->device(),
                 ", cannot switch to ", device);
        return;
    }
    g_selected_device = device;
}

Runtime &Runtime::get() {
    if (g_rt) return *g_rt;
    std::lock_guard<std::mutex> g(g_rt_mu);
    if (!g_rt) {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n == 0) {
            cudaGetLastError();
            fail(__FILE__, __LINE__, __func__,
                 str("no CUDA device available (", cudaGetErrorString(e), "); kfunca_b200 has no CPU fallback"));
        }
        int dev = g_selected_device;
        if (dev < 0) {
            const char *lr = std::getenv("LOCAL_RANK");
            dev = lr ? std::atoi(lr) % n : 0;
        }
        KF_CHECK(dev < n, "device ", dev, " out of range (", n, " visible)");
        g_rt = new Runtime(dev);
    }
    return *g_rt;
}

Runtime::Runtime(int device) : device_(device) {
    KF_CUDA(cudaSetDevice(device));
    cudaDeviceProp p;
    KF_CUDA(cudaGetDeviceProperties(&p, device));
    props_.sm_count = p.multiProcessorCount;
    props_.max_smem_optin = (int)p.sharedMemPerBlockOptin;
    props_.l2_bytes = p.l2CacheSize;
    props_.total_mem = p.totalGlobalMem;
    props_.cc_major = p.major;
    props_.cc_minor = p.minor;
    std::snprintf(props_.name, sizeof(props_.name), "%s", p.name);
    KF_CHECK(p.major == 10, "kfunca_b200 is built for sm_100a only; device ", device, " is sm_", p.major, p.minor);
    KF_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    pool_.reset(new Pool(cuda_raw_alloc, cuda_raw_free, nullptr, cuda_fence));
}

void Runtime::wait_for_stream(cudaStream_t other) {
    if (other == stream_) return;
    std::lock_guard<std::mutex> g(fence_mu_);
    if (fence_events_.size() < 16) {
        cudaEvent_t ev;
        KF_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        fence_events_.push_back(ev);
        fence_next_ = fence_events_.size() - 1;
    } else {
        fence_next_ = (fence_next_ + 1) % fence_events_.size();
    }
    cudaEvent_t ev = fence_events_[fence_next_];
    KF_CUDA(cudaEventRecord(ev, other));
    KF_CUDA(cudaStreamWaitEvent(stream_, ev, 0));
}

void Runtime::h2d(void *dst, const void *src, size_t bytes, bool sync_after) {
    if (!bytes) return;
    KF_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream_));
    if (sync_after) sync();
}
void Runtime::d2h(void *dst, const void *src, size_t bytes, bool sync_after) {
    if (!bytes) return;
    KF_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream_));
    if (sync_after) sync();
}
void Runtime::d2d(void *dst, const void *src, size_t bytes) {
    if (!bytes) return;
    KF_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, stream_));
}
void Runtime::memset_async(void *dst, int v, size_t bytes) {
    if (!bytes) return;
    KF_CUDA(cudaMemsetAsync(dst, v, bytes, stream_));
}
void Runtime::post_launch(const char *what) {
    launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) fail(__FILE__, __LINE__, what, str("kernel launch failed: ", cudaGetErrorString(e)));
}

// --------------------------------------------------------------------------------------- Storage
Storage::Storage(size_t nbytes, int dev) : bytes(nbytes), device(dev) {
    if (dev >= 0) {
        Runtime &rt = Runtime::get();
        KF_CHECK(dev == rt.device(), "tensor device ", dev, " != process device ", rt.device(),
                 " (kfunca_b200 runs one process per GPU)");
        ptr = rt.pool().allocate(nbytes);
    }
}
Storage::~Storage() {
    if (ptr && !external && g_rt) {
        try {
            g_rt->pool().release(ptr);
        } catch (...) {
        }
    }
}

}  // namespace kf

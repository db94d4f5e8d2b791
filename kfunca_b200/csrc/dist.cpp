// Native data-parallel layer: NCCL over NVLink 5 / NVSwitch, one process per GPU.  The reference has no
// multi-GPU path at all; the two exchanges the hot path needs are the weight-gradient all-reduce and the cross-shard
// sum / mean of a reduced scalar.  libnccl is dlopen'ed at kf_dist_init (no link-time dependency: single-GPU users never
// load it), the communicator lives on a library-owned NON-BLOCKING communication stream, and the gradient all-reduce is
// started from the autograd engine's leaf-gradient hook (ops::set_leaf_grad_hook) the moment backward() has enqueued a
// parameter's gradient — natively, with no Python or torch on the path:
//
//     library stream :  ... dW kernel ─ record(ev) ─ rest of the backward pass ............ wait(comm done) ─ next step
//     comm stream    :                  wait(ev) ─ ncclAllReduce(dW, AVG) ── ... ── record(comm done)
//
// Every gradient is reduced with ncclAvg (no separate 1/world pass).  Gradient memory is owned by its parameter until the
// next zero_grad(), which is ordered after kf_dist_overlap_end()'s join, so the pool's stream-ordered free rule holds.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <unordered_set>

#include "ops.h"
#include "runtime.h"

namespace kf {
namespace dist {

namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
};

NcclApi &api() {
    static NcclApi a;
    if (a.handle) return a;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (a.handle) break;
    }
    KF_CHECK(a.handle != nullptr, "kf_dist: libnccl.so.2 not found (", dlerror(), ")");
    auto sym = [&](const char *name) {
        void *p = dlsym(a.handle, name);
        KF_CHECK(p != nullptr, "kf_dist: symbol ", name, " missing from libnccl");
        return p;
    };
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
    a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(sym("ncclAllReduce"));
    a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
    a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
    a.GetVersion = reinterpret_cast<decltype(a.GetVersion)>(sym("ncclGetVersion"));
    return a;
}

#define KF_NCCL(expr)                                                                                          \
    do {                                                                                                       \
        ncclResult_t _r = (expr);                                                                              \
        if (_r != ncclSuccess) ::kf::fail(__FILE__, __LINE__, __func__, ::kf::str("NCCL error in `" #expr "`: ", api().GetErrorString(_r))); \
    } while (0)

struct State {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_grad = nullptr, ev_done = nullptr;
    std::unordered_set<const void *> params;  // data pointers of the tensors whose gradients are reduced by the hook
    bool overlapping = false;
    int64_t hook_calls = 0, bytes = 0;
};
State g;

ncclDataType_t nccl_dtype(DType d) {
    switch (d) {
    case KF_FLOAT: return ncclFloat32;
    case KF_DOUBLE: return ncclFloat64;
    case KF_HALF: return ncclFloat16;
    case KF_BFLOAT16: return ncclBfloat16;
    case KF_INT: return ncclInt32;
    case KF_LONG: return ncclInt64;
    case KF_BYTE: case KF_BOOL: return ncclUint8;
    case KF_CHAR: return ncclInt8;
    default: KF_CHECK(false, "kf_dist: dtype ", dtype_name(d), " cannot be all-reduced"); return ncclFloat32;
    }
}
ncclRedOp_t nccl_op(int op) {
    switch (op) {
    case 0: return ncclSum;
    case 1: return ncclAvg;
    case 2: return ncclMax;
    default: KF_CHECK(false, "kf_dist: bad reduction op ", op); return ncclSum;
    }
}

void grad_hook(const Tensor &leaf, const Tensor &grad, void *) {
    if (!g.overlapping || g.world == 1 || !g.params.count(leaf.data())) return;
    Runtime &rt = Runtime::get();
    KF_CHECK(grad.is_contiguous(), "kf_dist: gradients must be contiguous");
    // the comm stream starts this collective once the kernels that produced the gradient (already on the library stream) are done
    KF_CUDA(cudaEventRecord(g.ev_grad, rt.stream()));
    KF_CUDA(cudaStreamWaitEvent(g.comm_stream, g.ev_grad, 0));
    rt.pool().record_stream(grad.data(), g.comm_stream);  // a gradient dropped before dist_overlap_end() is fenced by the pool
    KF_NCCL(api().AllReduce(grad.data(), grad.data(), (size_t)grad.numel(), nccl_dtype(grad.dtype()), ncclAvg, g.comm, g.comm_stream));
    ++g.hook_calls;
    g.bytes += grad.numel() * (int64_t)grad.itemsize();
}

}  // namespace

void unique_id(void *out128) {
    ncclUniqueId id;
    KF_NCCL(api().GetUniqueId(&id));
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    std::memcpy(out128, &id, sizeof(id));
}

void init(const void *id128, int rank, int world) {
    KF_CHECK(g.comm == nullptr, "kf_dist_init: already initialised");
    KF_CHECK(world >= 1 && rank >= 0 && rank < world, "kf_dist_init: bad rank / world");
    Runtime &rt = Runtime::get();  // the device of this process is already selected (kf_set_device)
    g.rank = rank;
    g.world = world;
    KF_CUDA(cudaStreamCreateWithFlags(&g.comm_stream, cudaStreamNonBlocking));
    KF_CUDA(cudaEventCreateWithFlags(&g.ev_grad, cudaEventDisableTiming));
    KF_CUDA(cudaEventCreateWithFlags(&g.ev_done, cudaEventDisableTiming));
    (void)rt;
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    KF_NCCL(api().CommInitRank(&g.comm, world, id, rank));
}

void finalize() {
    if (!g.comm) return;
    ops::set_leaf_grad_hook(nullptr, nullptr);
    cudaStreamSynchronize(g.comm_stream);
    api().CommDestroy(g.comm);
    cudaEventDestroy(g.ev_grad);
    cudaEventDestroy(g.ev_done);
    cudaStreamDestroy(g.comm_stream);
    g = State();
}

bool initialised() { return g.comm != nullptr; }
int rank() { return g.rank; }
int world() { return g.world; }

// in place, on the LIBRARY stream: ordered after every kernel already enqueued and before everything enqueued later
void all_reduce(Tensor &t, int op) {
    if (g.world == 1) return;
    KF_CHECK(g.comm != nullptr, "kf_dist: not initialised");
    KF_CHECK(t.is_contiguous() && !t.is_meta(), "kf_dist_all_reduce: contiguous device tensors only");
    if (t.numel() == 0) return;
    KF_NCCL(api().AllReduce(t.data(), t.data(), (size_t)t.numel(), nccl_dtype(t.dtype()), nccl_op(op), g.comm, Runtime::get().stream()));
}

void overlap_begin(const std::vector<Tensor> &params) {
    KF_CHECK(g.comm != nullptr || g.world == 1, "kf_dist: not initialised");
    g.params.clear();
    for (auto &p : params) g.params.insert(p.data());
    g.overlapping = true;
    g.hook_calls = 0;
    ops::set_leaf_grad_hook(grad_hook, nullptr);
}

// join: everything enqueued on the library stream after this call sees the averaged gradients
int64_t overlap_end() {
    ops::set_leaf_grad_hook(nullptr, nullptr);
    g.overlapping = false;
    if (g.world > 1 && g.hook_calls > 0) {
        KF_CUDA(cudaEventRecord(g.ev_done, g.comm_stream));
        KF_CUDA(cudaStreamWaitEvent(Runtime::get().stream(), g.ev_done, 0));
    }
    return g.hook_calls;
}

int nccl_version() {
    int v = 0;
    api().GetVersion(&v);
    return v;
}

}  // namespace dist
}  // namespace kf

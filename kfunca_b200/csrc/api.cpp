// extern "C" boundary: every entry point of include/kfunca_b200.h.  Thin by construction — argument
// marshalling, exception -> status code + thread-local message, nothing else.
#include <sched.h>

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "ops.h"
#include "runtime.h"

using namespace kf;

struct kf_tensor_s {
    Tensor t;
};
struct kf_event_s {
    cudaEvent_t ev;
};

static thread_local std::string g_last_error;

#define KF_API_BEGIN try {
#define KF_API_END                                   \
    return 0;                                        \
    }                                                \
    catch (const std::exception &e) {                \
        g_last_error = e.what();                     \
        return 1;                                    \
    }                                                \
    catch (...) {                                    \
        g_last_error = "unknown C++ exception";      \
        return 2;                                    \
    }

static Tensor &T(kf_tensor_t h) {
    KF_CHECK(h != nullptr, "null tensor handle");
    return h->t;
}
static kf_tensor_t wrap(const Tensor &t) { return new kf_tensor_s{t}; }
static std::vector<int64_t> vec(const int64_t *p, int n) {
    KF_CHECK(n >= 0 && (n == 0 || p != nullptr));
    return std::vector<int64_t>(p, p + n);
}
static int copy_string(const std::string &s, char *buf, size_t n) {
    if (buf && n) {
        std::snprintf(buf, n, "%s", s.c_str());
    }
    return 0;
}

#pragma GCC visibility push(default)
extern "C" {

const char *kf_last_error(void) { return g_last_error.c_str(); }

int kf_version(int *major, int *minor) {
    if (major) *major = 0;
    if (minor) *minor = 1;
    return 0;
}
int kf_device_count(int *count) {
    KF_API_BEGIN
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    *count = n;
    KF_API_END
}
int kf_set_device(int device) {
    KF_API_BEGIN
    Runtime::select_device(device);
    Runtime::get();
    KF_API_END
}
int kf_get_device(int *device) {
    KF_API_BEGIN
    *device = Runtime::get().device();
    KF_API_END
}
int kf_stream(void **s) {
    KF_API_BEGIN
    *s = (void *)Runtime::get().stream();
    KF_API_END
}
int kf_synchronize(void) {
    KF_API_BEGIN
    Runtime::get().sync();
    KF_API_END
}
int kf_device_info(char *buf, size_t n) {
    KF_API_BEGIN
    copy_string(device_info_string(), buf, n);
    KF_API_END
}
int kf_memstat(char *buf, size_t n) {
    KF_API_BEGIN
    copy_string(Runtime::get().pool().report(), buf, n);
    KF_API_END
}
int kf_mem_stats(int64_t *in_use, int64_t *reserved, int64_t *n_mallocs) {
    KF_API_BEGIN
    Pool &p = Runtime::get().pool();
    if (in_use) *in_use = p.bytes_in_use();
    if (reserved) *reserved = p.bytes_reserved();
    if (n_mallocs) *n_mallocs = p.arena_mallocs();
    KF_API_END
}
int kf_empty_cache(void) {
    KF_API_BEGIN
    Runtime::get().sync();
    Runtime::get().pool().empty_cache();
    KF_API_END
}
int kf_launch_count(int64_t *count) {
    KF_API_BEGIN
    *count = Runtime::initialised() ? Runtime::get().launches.load() : 0;
    KF_API_END
}

int kf_event_create(kf_event_t *ev) {
    KF_API_BEGIN
    Runtime::get();
    auto *e = new kf_event_s{};
    KF_CUDA(cudaEventCreate(&e->ev));
    *ev = e;
    KF_API_END
}
int kf_event_record(kf_event_t ev) {
    KF_API_BEGIN
    KF_CUDA(cudaEventRecord(ev->ev, Runtime::get().stream()));
    KF_API_END
}
int kf_event_synchronize(kf_event_t ev) {
    KF_API_BEGIN
    KF_CUDA(cudaEventSynchronize(ev->ev));
    KF_API_END
}
int kf_event_elapsed_ms(kf_event_t a, kf_event_t b, float *ms) {
    KF_API_BEGIN
    KF_CUDA(cudaEventElapsedTime(ms, a->ev, b->ev));
    KF_API_END
}
int kf_event_destroy(kf_event_t ev) {
    KF_API_BEGIN
    if (ev) {
        cudaEventDestroy(ev->ev);
        delete ev;
    }
    KF_API_END
}
// NUMA placement of the host side of the boundary: pinned buffers are allocated (and first touched) by a thread bound to the CPUs of
// the GPU's own NUMA node, so that at N ranks per box every rank's H2D / D2H traffic stays on the memory controller next to its PCIe
// root instead of all ranks sharing node 0 (measured in round 1: e2e 5.9 -> 21.8 ms per step from 1 to 8 ranks).  KF_NUMA=0 disables.
static int gpu_numa_node() {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), Runtime::get().device()) != cudaSuccess) return -1;
    for (char *c = bus; *c; ++c) *c = (char)std::tolower(*c);
    const std::string path = std::string("/sys/bus/pci/devices/") + bus + "/numa_node";
    FILE *f = std::fopen(path.c_str(), "r");
    if (!f) return -1;
    int node = -1;
    if (std::fscanf(f, "%d", &node) != 1) node = -1;
    std::fclose(f);
    return node;
}
static std::string node_cpulist(int node) {
    if (node < 0) return "";
    char path[96];
    std::snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
    FILE *f = std::fopen(path, "r");
    if (!f) return "";
    char buf[1024] = {0};
    const bool ok = std::fgets(buf, sizeof(buf), f) != nullptr;
    std::fclose(f);
    std::string s = ok ? buf : "";
    while (!s.empty() && (s.back() == '\n' || s.back() == ' ')) s.pop_back();
    return s;
}
static bool parse_cpulist(const std::string &s, cpu_set_t *set) {  // "0-31,64-95"
    CPU_ZERO(set);
    int n = 0;
    size_t i = 0;
    while (i < s.size()) {
        char *end = nullptr;
        const long a = std::strtol(s.c_str() + i, &end, 10);
        if (end == s.c_str() + i) return false;
        long b = a;
        i = (size_t)(end - s.c_str());
        if (i < s.size() && s[i] == '-') {
            b = std::strtol(s.c_str() + i + 1, &end, 10);
            i = (size_t)(end - s.c_str());
        }
        for (long c = a; c <= b && c < CPU_SETSIZE; ++c) {
            CPU_SET((int)c, set);
            ++n;
        }
        if (i < s.size() && s[i] == ',') ++i;
    }
    return n > 0;
}

int kf_numa_info(int *node, char *cpulist, size_t cap) {
    KF_API_BEGIN
    const int nd = gpu_numa_node();
    if (node) *node = nd;
    if (cpulist && cap) std::snprintf(cpulist, cap, "%s", node_cpulist(nd).c_str());
    KF_API_END
}

int kf_host_alloc_pinned(size_t bytes, void **ptr) {
    KF_API_BEGIN
    Runtime::get();
    const char *env = std::getenv("KF_NUMA");
    cpu_set_t old_set, node_set;
    bool bound = false;
    if (!(env && env[0] == '0') && sched_getaffinity(0, sizeof(old_set), &old_set) == 0 && parse_cpulist(node_cpulist(gpu_numa_node()), &node_set)) {
        cpu_set_t both;
        CPU_AND(&both, &old_set, &node_set);  // stay inside whatever cpuset the launcher gave us
        if (CPU_COUNT(&both) > 0) bound = sched_setaffinity(0, sizeof(both), &both) == 0;
    }
    const cudaError_t e = cudaMallocHost(ptr, bytes ? bytes : 1);
    if (e == cudaSuccess && bound) std::memset(*ptr, 0, bytes ? bytes : 1);  // first touch on the local node
    if (bound) sched_setaffinity(0, sizeof(old_set), &old_set);
    KF_CUDA(e);
    KF_API_END
}
int kf_host_free_pinned(void *ptr) {
    KF_API_BEGIN
    if (ptr) KF_CUDA(cudaFreeHost(ptr));
    KF_API_END
}

// ---------------------------------------------------------------- creation / host I/O
int kf_empty(const int64_t *shape, int ndim, int dtype, int device, kf_tensor_t *out) {
    KF_API_BEGIN
    *out = wrap(empty(vec(shape, ndim), dtype, device));
    KF_API_END
}
int kf_zeros(const int64_t *shape, int ndim, int dtype, int device, kf_tensor_t *out) {
    KF_API_BEGIN
    *out = wrap(zeros(vec(shape, ndim), dtype, device));
    KF_API_END
}
int kf_empty_like(kf_tensor_t self, kf_tensor_t *out) {
    KF_API_BEGIN
    *out = wrap(empty_like(T(self)));
    KF_API_END
}
int kf_from_host(const void *src, const int64_t *shape, int ndim, int dtype, int device, kf_tensor_t *out) {
    KF_API_BEGIN
    Tensor t = empty(vec(shape, ndim), dtype, device);
    KF_CHECK(device >= 0, "from_host needs a real device");
    Runtime::get().h2d(t.data(), src, (size_t)t.numel() * t.itemsize(), true);
    *out = wrap(t);
    KF_API_END
}
int kf_to_host(kf_tensor_t self, void *dst, size_t dst_bytes) {
    KF_API_BEGIN
    Tensor &t = T(self);
    KF_CHECK(t.defined() && !t.is_meta(), "to_host: tensor has no device data");
    KF_CHECK(t.is_contiguous(), "to_host: tensor must be contiguous");
    const size_t bytes = (size_t)t.numel() * t.itemsize();
    KF_CHECK(dst_bytes >= bytes, "to_host: destination too small");
    Runtime::get().d2h(dst, t.data(), bytes, true);
    KF_API_END
}
int kf_memcpy_h2d_async(void *dst_device, const void *src_host, size_t bytes) {
    KF_API_BEGIN
    Runtime::get().h2d(dst_device, src_host, bytes, false);
    KF_API_END
}
int kf_memcpy_d2h_async(void *dst_host, const void *src_device, size_t bytes) {
    KF_API_BEGIN
    Runtime::get().d2h(dst_host, src_device, bytes, false);
    KF_API_END
}

// ---------------------------------------------------------------- handles / metadata
int kf_retain(kf_tensor_t self, kf_tensor_t *out) {
    KF_API_BEGIN
    *out = wrap(T(self));
    KF_API_END
}
int kf_release(kf_tensor_t self) {
    KF_API_BEGIN
    delete self;
    KF_API_END
}
int kf_defined(kf_tensor_t self, int *out) {
    KF_API_BEGIN
    *out = self && self->t.defined();
    KF_API_END
}
int kf_dim(kf_tensor_t self, int *out) {
    KF_API_BEGIN
    *out = T(self).dim();
    KF_API_END
}
int kf_numel(kf_tensor_t self, int64_t *out) {
    KF_API_BEGIN
    *out = T(self).numel();
    KF_API_END
}
int kf_dtype(kf_tensor_t self, int *out) {
    KF_API_BEGIN
    *out = T(self).dtype();
    KF_API_END
}
int kf_device(kf_tensor_t self, int *out) {
    KF_API_BEGIN
    *out = T(self).device();
    KF_API_END
}
int kf_shape(kf_tensor_t self, int d, int64_t *out) {
    KF_API_BEGIN
    *out = T(self).size(d);
    KF_API_END
}
int kf_sizes(kf_tensor_t self, int64_t *out, int *ndim) {
    KF_API_BEGIN
    Tensor &t = T(self);
    *ndim = t.dim();
    for (int i = 0; i < t.dim(); ++i) out[i] = t.impl->shape[i];
    KF_API_END
}
int kf_strides(kf_tensor_t self, int64_t *out, int *ndim) {
    KF_API_BEGIN
    Tensor &t = T(self);
    *ndim = t.dim();
    for (int i = 0; i < t.dim(); ++i) out[i] = t.impl->stride[i];
    KF_API_END
}
int kf_storage_offset(kf_tensor_t self, int64_t *out) {
    KF_API_BEGIN
    *out = T(self).impl->offset;
    KF_API_END
}
int kf_is_contiguous(kf_tensor_t self, int *out) {
    KF_API_BEGIN
    *out = T(self).is_contiguous();
    KF_API_END
}
int kf_data_ptr(kf_tensor_t self, void **out) {
    KF_API_BEGIN
    *out = T(self).data();
    KF_API_END
}
int kf_storage_bytes(kf_tensor_t self, size_t *out) {
    KF_API_BEGIN
    *out = T(self).impl->storage->bytes;
    KF_API_END
}
int kf_storage_ref_count(kf_tensor_t self, int64_t *out) {
    KF_API_BEGIN
    *out = T(self).impl->storage.use_count();
    KF_API_END
}
int kf_impl_ref_count(kf_tensor_t self, int64_t *out) {
    KF_API_BEGIN
    *out = T(self).impl.use_count();
    KF_API_END
}
int kf_element_size(int dtype, size_t *out) {
    KF_API_BEGIN
    *out = element_size(dtype);
    KF_API_END
}
int kf_item(kf_tensor_t self, const int64_t *indices, int n, void *out8) {
    KF_API_BEGIN
    Tensor &t = T(self);
    KF_CHECK(n == t.dim(), "item(): need one index per dimension");
    KF_CHECK(!t.is_meta(), "item(): meta tensor");
    int64_t off = 0;
    for (int i = 0; i < n; ++i) {
        KF_CHECK(indices[i] >= 0 && indices[i] < t.size(i), "item(): index out of range");
        off += indices[i] * t.impl->stride[i];
    }
    std::memset(out8, 0, 8);
    Runtime::get().d2h(out8, (char *)t.data() + off * (int64_t)t.itemsize(), t.itemsize(), true);
    KF_API_END
}
int kf_to_string(kf_tensor_t self, char *buf, size_t n) {
    KF_API_BEGIN
    copy_string(self ? self->t.to_string() : std::string("Tensor(Undefined)"), buf, n);
    KF_API_END
}

// ---------------------------------------------------------------- views
int kf_as_strided(kf_tensor_t self, const int64_t *sizes, const int64_t *strides, int ndim, int64_t off, kf_tensor_t *out) {
    KF_API_BEGIN
    // a raw restride has no gradient formula: the result is a constant view (requires_grad cleared) instead of a tensor that
    // claims to need a gradient but has no edge back to `self`
    Tensor v = T(self).as_strided(vec(sizes, ndim), strides ? vec(strides, ndim) : std::vector<int64_t>{}, off);
    v.impl->requires_grad = false;
    *out = wrap(v);
    KF_API_END
}
int kf_permute(kf_tensor_t self, const int64_t *dims, int ndim, kf_tensor_t *out) {
    KF_API_BEGIN
    *out = wrap(ops::permute(T(self), vec(dims, ndim)));
    KF_API_END
}
int kf_view(kf_tensor_t self, const int64_t *sizes, int ndim, kf_tensor_t *out) {
    KF_API_BEGIN
    *out = wrap(ops::view(T(self), vec(sizes, ndim)));
    KF_API_END
}
int kf_slice(kf_tensor_t self, int64_t dim, int64_t start, int64_t end, int64_t step, kf_tensor_t *out) {
    KF_API_BEGIN
    *out = wrap(ops::slice(T(self), dim, start, end, step));
    KF_API_END
}
int kf_select(kf_tensor_t self, int64_t dim, int64_t index, kf_tensor_t *out) {
    KF_API_BEGIN
    *out = wrap(ops::select(T(self), dim, index));
    KF_API_END
}
int kf_narrow(kf_tensor_t self, int64_t dim, int64_t start, int64_t length, kf_tensor_t *out) {
    KF_API_BEGIN
    *out = wrap(ops::narrow(T(self), dim, start, length));
    KF_API_END
}
int kf_split(kf_tensor_t self, const int64_t *sizes, int n, int64_t dim, kf_tensor_t *outs) {
    KF_API_BEGIN
    auto parts = ops::split(T(self), vec(sizes, n), dim);
    for (int i = 0; i < n; ++i) outs[i] = wrap(parts[i]);
    KF_API_END
}
int kf_contiguous(kf_tensor_t self, kf_tensor_t *out) {
    KF_API_BEGIN
    *out = wrap(ops::contiguous(T(self)));
    KF_API_END
}

// ---------------------------------------------------------------- elementwise
static int ew_op(int op) {
    KF_CHECK(op >= KF_OP_ADD && op <= KF_OP_DIV, "bad binary op ", op);
    return op;  // KF_OP_* values equal EW_ADD..EW_DIV
}
int kf_binary(int op, kf_tensor_t a, kf_tensor_t b, kf_tensor_t *out) {
    KF_API_BEGIN
    *out = wrap(ops::binary(ew_op(op), T(a), T(b)));
    KF_API_END
}
int kf_binary_(int op, kf_tensor_t self, kf_tensor_t other) {
    KF_API_BEGIN
    ops::binary_(ew_op(op), T(self), T(other));
    KF_API_END
}
int kf_binary_scalar(int op, kf_tensor_t a, double scalar, kf_tensor_t *out) {
    KF_API_BEGIN
    *out = wrap(ops::binary_scalar(ew_op(op), T(a), scalar));
    KF_API_END
}
int kf_binary_scalar_(int op, kf_tensor_t self, double scalar) {
    KF_API_BEGIN
    ops::binary_scalar_(ew_op(op), T(self), scalar);
    KF_API_END
}
int kf_fill_(kf_tensor_t self, double value) {
    KF_API_BEGIN
    ops::fill_(T(self), value);
    KF_API_END
}
int kf_copy_(kf_tensor_t self, kf_tensor_t src) {
    KF_API_BEGIN
    ops::copy_(T(self), T(src));
    KF_API_END
}
int kf_clone(kf_tensor_t self, kf_tensor_t *out) {
    KF_API_BEGIN
    *out = wrap(ops::clone(T(self)));
    KF_API_END
}
int kf_convert(kf_tensor_t self, int dtype, kf_tensor_t *out) {
    KF_API_BEGIN
    *out = wrap(ops::convert(T(self), dtype));
    KF_API_END
}
int kf_unary(int op, kf_tensor_t a, kf_tensor_t *out) {
    KF_API_BEGIN
    KF_CHECK(op >= KF_UOP_SQRT && op <= KF_UOP_NEG, "bad unary op ", op);
    *out = wrap(ops::unary(EW_SQRT + op, T(a)));
    KF_API_END
}

// ---------------------------------------------------------------- reductions / sort
int kf_sum(kf_tensor_t self, int64_t dim, kf_tensor_t *out) {
    KF_API_BEGIN
    *out = wrap(ops::sum(T(self), dim));
    KF_API_END
}
int kf_mean(kf_tensor_t self, int64_t dim, kf_tensor_t *out) {
    KF_API_BEGIN
    *out = wrap(ops::mean(T(self), dim));
    KF_API_END
}
int kf_mean_var(kf_tensor_t self, int64_t dim, int take_sqrt, kf_tensor_t *mean, kf_tensor_t *var) {
    KF_API_BEGIN
    auto [m, v] = ops::mean_var(T(self), dim, take_sqrt != 0);
    *mean = wrap(m);
    *var = wrap(v);
    KF_API_END
}
int kf_norm_stat(kf_tensor_t self, int64_t dim, kf_tensor_t *mean, kf_tensor_t *invstd) {
    KF_API_BEGIN
    auto [m, v] = ops::norm_stat(T(self), dim);
    *mean = wrap(m);
    *invstd = wrap(v);
    KF_API_END
}
int kf_sort(kf_tensor_t self, int64_t dim, int descending, kf_tensor_t *values, kf_tensor_t *indices) {
    KF_API_BEGIN
    auto [v, i] = ops::sort(T(self), dim, descending != 0);
    *values = wrap(v);
    *indices = wrap(i);
    KF_API_END
}
int kf_topk(kf_tensor_t self, int64_t k, int64_t dim, int largest, kf_tensor_t *values, kf_tensor_t *indices) {
    KF_API_BEGIN
    auto [v, i] = ops::topk(T(self), k, dim, largest != 0);
    *values = wrap(v);
    *indices = wrap(i);
    KF_API_END
}
int kf_cat(const kf_tensor_t *tensors, int n, int64_t dim, kf_tensor_t *out) {
    KF_API_BEGIN
    std::vector<Tensor> ts;
    for (int i = 0; i < n; ++i) ts.push_back(T(tensors[i]));
    *out = wrap(ops::cat(ts, dim));
    KF_API_END
}
int kf_index_put_(kf_tensor_t self, const kf_tensor_t *indices, int n, kf_tensor_t values) {
    KF_API_BEGIN
    std::vector<Tensor> ix;
    for (int i = 0; i < n; ++i) ix.push_back(T(indices[i]));
    ops::index_put_(T(self), ix, T(values));
    KF_API_END
}

// ---------------------------------------------------------------- contractions
int kf_gemm(kf_tensor_t a, kf_tensor_t b, float alpha, float beta, kf_tensor_t *out) {
    KF_API_BEGIN
    *out = wrap(ops::gemm(T(a), T(b), alpha, beta));
    KF_API_END
}
int kf_gemm_out(kf_tensor_t out, kf_tensor_t a, kf_tensor_t b, float alpha, float beta) {
    KF_API_BEGIN
    ops::gemm_out(T(out), T(a), T(b), alpha, beta);
    KF_API_END
}
int kf_matmul(kf_tensor_t a, int trans_a, kf_tensor_t b, int trans_b, float alpha, kf_tensor_t *out) {
    KF_API_BEGIN
    *out = wrap(ops::matmul(T(a), trans_a != 0, T(b), trans_b != 0, alpha));
    KF_API_END
}
int kf_causal_attention(kf_tensor_t q, kf_tensor_t k, kf_tensor_t v, kf_tensor_t *out) {
    KF_API_BEGIN
    *out = wrap(ops::causal_attention(T(q), T(k), T(v)));
    KF_API_END
}
int kf_causal_attention_fwd(kf_tensor_t q, kf_tensor_t k, kf_tensor_t v, kf_tensor_t *out, kf_tensor_t *lse) {
    KF_API_BEGIN
    auto [o, l] = ops::causal_attention_fwd(T(q), T(k), T(v));
    *out = wrap(o);
    *lse = wrap(l);
    KF_API_END
}
int kf_causal_attention_bwd(kf_tensor_t dout, kf_tensor_t q, kf_tensor_t k, kf_tensor_t v, kf_tensor_t out, kf_tensor_t lse,
                            kf_tensor_t *dq, kf_tensor_t *dk, kf_tensor_t *dv) {
    KF_API_BEGIN
    auto [a, b, c] = ops::causal_attention_bwd(T(dout), T(q), T(k), T(v), T(out), T(lse));
    *dq = wrap(a);
    *dk = wrap(b);
    *dv = wrap(c);
    KF_API_END
}

int kf_gemm_host(const void *a_host, const void *b_host, void *c_host, int64_t M, int64_t N, int64_t K, int dtype, float alpha,
                 int64_t slab_rows) {
    KF_API_BEGIN
    ops::gemm_host(a_host, b_host, c_host, M, N, K, (DType)dtype, alpha, slab_rows);
    KF_API_END
}
int kf_layer_norm(kf_tensor_t x, kf_tensor_t gain, double eps, kf_tensor_t *out) {
    KF_API_BEGIN
    *out = wrap(ops::layer_norm(T(x), T(gain), eps));
    KF_API_END
}
int kf_rms_norm(kf_tensor_t x, kf_tensor_t gain, double eps, kf_tensor_t *out) {
    KF_API_BEGIN
    *out = wrap(ops::rms_norm(T(x), T(gain), eps));
    KF_API_END
}
int kf_gemm_residual(kf_tensor_t a, kf_tensor_t b, kf_tensor_t residual, float alpha, kf_tensor_t *out) {
    KF_API_BEGIN
    *out = wrap(ops::gemm_residual(T(a), T(b), T(residual), alpha));
    KF_API_END
}
int kf_qkv_linear(kf_tensor_t x, kf_tensor_t w, kf_tensor_t bias, kf_tensor_t *out) {
    KF_API_BEGIN
    *out = wrap(ops::qkv_linear(T(x), T(w), bias ? T(bias) : Tensor()));
    KF_API_END
}
int kf_gemm_glu(kf_tensor_t a, kf_tensor_t b1, kf_tensor_t b3, kf_tensor_t *out) {
    KF_API_BEGIN
    *out = wrap(ops::gemm_glu(T(a), T(b1), T(b3)));
    KF_API_END
}
int kf_qkv_attention(kf_tensor_t qkv, int64_t heads, kf_tensor_t *out) {
    KF_API_BEGIN
    *out = wrap(ops::qkv_attention(T(qkv), heads));
    KF_API_END
}
int kf_embedding(kf_tensor_t weight, kf_tensor_t indices, kf_tensor_t *out) {
    KF_API_BEGIN
    *out = wrap(ops::embedding(T(weight), T(indices)));
    KF_API_END
}
int kf_random_uniform_(kf_tensor_t self, uint64_t seed, double lo, double hi) {
    KF_API_BEGIN
    ops::random_uniform_(T(self), seed, lo, hi);
    KF_API_END
}
// ---------------------------------------------------------------- autograd
int kf_requires_grad(kf_tensor_t self, int *out) {
    KF_API_BEGIN
    *out = T(self).requires_grad();
    KF_API_END
}
int kf_set_requires_grad(kf_tensor_t self, int flag) {
    KF_API_BEGIN
    T(self).impl->requires_grad = flag != 0;
    KF_API_END
}
int kf_backward(kf_tensor_t self, kf_tensor_t grad_output) {
    KF_API_BEGIN
    ops::backward(T(self), T(grad_output));
    KF_API_END
}
int kf_grad(kf_tensor_t self, kf_tensor_t *out) {
    KF_API_BEGIN
    Tensor &t = T(self);
    *out = (t.impl->grad && t.impl->grad->defined()) ? wrap(*t.impl->grad) : nullptr;
    KF_API_END
}
int kf_zero_grad(kf_tensor_t self) {
    KF_API_BEGIN
    T(self).impl->grad.reset();
    KF_API_END
}
static kf_leaf_grad_hook_t g_hook_fn = nullptr;
static void *g_hook_ctx = nullptr;
static void leaf_hook_trampoline(const Tensor &leaf, const Tensor &grad, void *) {
    if (!g_hook_fn) return;
    kf_tensor_s l{leaf}, g{grad};  // borrowed handles living on this frame
    g_hook_fn(&l, &g, g_hook_ctx);
}
int kf_set_leaf_grad_hook(kf_leaf_grad_hook_t fn, void *ctx) {
    KF_API_BEGIN
    g_hook_fn = fn;
    g_hook_ctx = ctx;
    ops::set_leaf_grad_hook(fn ? leaf_hook_trampoline : nullptr, nullptr);
    KF_API_END
}

// ---------------------------------------------------------------- data-parallel layer (dist.cpp)
int kf_dist_unique_id(void *out128) {
    KF_API_BEGIN
    KF_CHECK(out128 != nullptr);
    dist::unique_id(out128);
    KF_API_END
}
int kf_dist_init(const void *id128, int rank, int world) {
    KF_API_BEGIN
    KF_CHECK(id128 != nullptr);
    dist::init(id128, rank, world);
    KF_API_END
}
int kf_dist_finalize(void) {
    KF_API_BEGIN
    dist::finalize();
    KF_API_END
}
int kf_dist_info(int *initialised, int *rank, int *world, int *nccl_version) {
    KF_API_BEGIN
    if (initialised) *initialised = dist::initialised();
    if (rank) *rank = dist::rank();
    if (world) *world = dist::world();
    if (nccl_version) *nccl_version = dist::initialised() ? dist::nccl_version() : 0;
    KF_API_END
}
int kf_dist_all_reduce(kf_tensor_t t, int op) {
    KF_API_BEGIN
    dist::all_reduce(T(t), op);
    KF_API_END
}
int kf_dist_overlap_begin(const kf_tensor_t *params, int n) {
    KF_API_BEGIN
    std::vector<Tensor> ps;
    for (int i = 0; i < n; ++i) ps.push_back(T(params[i]));
    dist::overlap_begin(ps);
    KF_API_END
}
int kf_dist_overlap_end(int64_t *n_reduced) {
    KF_API_BEGIN
    const int64_t n = dist::overlap_end();
    if (n_reduced) *n_reduced = n;
    KF_API_END
}

// ---------------------------------------------------------------- host-logic probes
int kf_debug_plan_binary(kf_tensor_t a, kf_tensor_t b, int *ndim, int64_t *shape, int64_t *strides3, int *common_dtype) {
    KF_API_BEGIN
    EwPlan plan;
    Tensor out;
    ops::plan_elementwise(plan, out, &T(a), &T(b), true);
    *ndim = plan.ndim;
    for (int d = 0; d < plan.ndim; ++d) {
        shape[d] = plan.shape[d];
        for (int i = 0; i < 3; ++i) strides3[i * KF_MAX_DIMS + d] = plan.stride[i][d];
    }
    *common_dtype = out.dtype();
    KF_API_END
}
int kf_promote_types(int a, int b, int *out) {
    KF_API_BEGIN
    KF_CHECK(a >= 0 && a < KF_UNDEFINED && b >= 0 && b < KF_UNDEFINED);
    *out = promote(a, b);
    KF_API_END
}

namespace {
struct FakeArena {
    int64_t next = 1 << 20;
};
void *fake_alloc(size_t bytes, void *ctx) {
    auto *f = static_cast<FakeArena *>(ctx);
    void *p = (void *)(intptr_t)f->next;
    f->next += (int64_t)((bytes + 4095) / 4096 * 4096) + (1 << 20);
    return p;
}
void fake_free(void *, void *) {}
}  // namespace

struct FenceLog {
    FakeArena arena;
    std::vector<int64_t> log;
};
static void *fl_alloc(size_t bytes, void *ctx) { return fake_alloc(bytes, &static_cast<FenceLog *>(ctx)->arena); }
static void fl_free(void *p, void *ctx) { fake_free(p, &static_cast<FenceLog *>(ctx)->arena); }
static void fl_fence(void *stream, void *ctx) { static_cast<FenceLog *>(ctx)->log.push_back((int64_t)(intptr_t)stream); }

int kf_debug_pool_fences(const int *kind, const int64_t *arg, const int64_t *arg2, int n, int64_t *fence_log, int cap, int *nfences) {
    KF_API_BEGIN
    FenceLog fl;
    {
        Pool pool(fl_alloc, fl_free, &fl, fl_fence);
        std::vector<void *> ptrs(n, nullptr);
        for (int i = 0; i < n; ++i) {
            if (kind[i] == 0) {
                ptrs[i] = pool.allocate((size_t)arg[i]);
            } else {
                const int64_t j = arg[i];
                KF_CHECK(j >= 0 && j < i && ptrs[j], "pool fence trace: bad target");
                if (kind[i] == 1) {
                    pool.release(ptrs[j]);
                    ptrs[j] = nullptr;
                } else {
                    pool.record_stream((char *)ptrs[j] + 1, (void *)(intptr_t)arg2[i]);  // an interior address finds its block
                }
            }
        }
    }
    *nfences = (int)fl.log.size();
    for (int i = 0; i < (int)fl.log.size() && i < cap; ++i) fence_log[i] = fl.log[i];
    KF_API_END
}

int kf_record_stream(kf_tensor_t t, void *stream) {
    KF_API_BEGIN
    const Tensor &x = T(t);
    KF_CHECK(x.defined() && !x.is_meta(), "record_stream needs a device tensor");
    if (!x.impl->storage->external && stream != (void *)Runtime::get().stream()) Runtime::get().pool().record_stream(x.impl->storage->ptr, stream);
    KF_API_END
}

int kf_debug_pool_trace(const int64_t *ops_, int n, int64_t *offsets, int64_t *stats3) {
    KF_API_BEGIN
    FakeArena arena;
    Pool pool(fake_alloc, fake_free, &arena);
    std::vector<void *> ptrs(n, nullptr);
    for (int i = 0; i < n; ++i) {
        if (ops_[i] > 0) {
            ptrs[i] = pool.allocate((size_t)ops_[i]);
            offsets[i] = (int64_t)(intptr_t)ptrs[i];
        } else if (ops_[i] < 0) {
            const int64_t j = -ops_[i] - 1;
            KF_CHECK(j >= 0 && j < i && ptrs[j], "pool trace: bad free target");
            pool.release(ptrs[j]);
            ptrs[j] = nullptr;
            offsets[i] = -1;
        } else {
            pool.empty_cache();
            offsets[i] = -1;
        }
    }
    stats3[0] = pool.bytes_in_use();
    stats3[1] = pool.bytes_reserved();
    stats3[2] = pool.arena_mallocs();
    KF_API_END
}

}  // extern "C"
#pragma GCC visibility pop

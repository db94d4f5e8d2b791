// Host-side core types of kfunca_b200: errors, dtypes, storage / impl / tensor handles, autograd node.
// Behavioural contract follows the reference's src/core/include/{tensor.h,tensor_impl.h,scalar_type.h};
// the implementation is new (std::shared_ptr-free intrusive counts, real contiguity tracking,
// stream-ordered pool storage).
#pragma once
#include <atomic>
#include <cstdint>
#include <cstring>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/kfunca_b200.h"

namespace kf {

// ---------------------------------------------------------------- errors
struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

template <typename... A>
inline std::string str(const A &...a) {
    std::ostringstream os;
    (void)std::initializer_list<int>{((os << a), 0)...};
    return os.str();
}

[[noreturn]] void fail(const char *file, int line, const char *func, const std::string &msg);

// same text shape as the reference's CHECK_FAIL (src/core/utils/exception.h:123-131)
#define KF_CHECK(cond, ...)                                                                      \
    do {                                                                                         \
        if (!(cond)) ::kf::fail(__FILE__, __LINE__, __func__, ::kf::str("`" #cond "` ", ##__VA_ARGS__)); \
    } while (0)

// ---------------------------------------------------------------- dtypes
using DType = int;  // kf_dtype_t values

inline size_t element_size(DType t) {
    switch (t) {
    case KF_BOOL: case KF_BYTE: case KF_CHAR: return 1;
    case KF_SHORT: case KF_HALF: case KF_BFLOAT16: return 2;
    case KF_INT: case KF_FLOAT: return 4;
    case KF_LONG: case KF_DOUBLE: return 8;
    default: break;
    }
    fail(__FILE__, __LINE__, __func__, "Unknown ScalarType");
}
inline const char *dtype_name(DType t) {
    static const char *n[] = {"Bool", "Byte", "Char", "Short", "Int", "Long", "Half", "BFloat16", "Float", "Double", "Undefined"};
    return (t >= 0 && t <= 10) ? n[t] : "UNKNOWN_SCALAR";
}
inline bool is_floating(DType t) { return t == KF_DOUBLE || t == KF_FLOAT || t == KF_HALF || t == KF_BFLOAT16; }
inline bool is_unsigned_class(DType t) { return t == KF_BYTE || t == KF_BOOL; }
DType promote(DType a, DType b);  // ref: update_common_dtype, tensor_iterator.cpp:32-44

// compute ("accumulate") class of a dtype, ref: src/core/include/accumulate_type.h:17-27
enum AccKind { ACC_F32 = 0, ACC_F64 = 1, ACC_I64 = 2, ACC_BOOL = 3 };
inline AccKind acc_kind(DType t) {
    switch (t) {
    case KF_HALF: case KF_BFLOAT16: case KF_FLOAT: return ACC_F32;
    case KF_DOUBLE: return ACC_F64;
    case KF_BOOL: return ACC_BOOL;
    default: return ACC_I64;
    }
}

inline int wrap_dim(int64_t d, int64_t ndim) {  // ref: maybe_wrap_dim, tensor_impl.h:17-19
    KF_CHECK(ndim > 0 && d >= -ndim && d < ndim, "dim ", d, " out of range for ", ndim, "-d tensor");
    return (int)(d < 0 ? d + ndim : d);
}

// ---------------------------------------------------------------- intrusive refcount
struct RefCounted {
    std::atomic<int64_t> refs{0};
    virtual ~RefCounted() = default;
};
template <class T>
class Ref {
    T *p_ = nullptr;
public:
    Ref() = default;
    explicit Ref(T *p) : p_(p) { if (p_) p_->refs.fetch_add(1, std::memory_order_relaxed); }
    Ref(const Ref &o) : p_(o.p_) { if (p_) p_->refs.fetch_add(1, std::memory_order_relaxed); }
    Ref(Ref &&o) noexcept : p_(o.p_) { o.p_ = nullptr; }
    Ref &operator=(Ref o) noexcept { std::swap(p_, o.p_); return *this; }
    ~Ref() { if (p_ && p_->refs.fetch_sub(1, std::memory_order_acq_rel) == 1) delete p_; }
    T *get() const { return p_; }
    T *operator->() const { return p_; }
    explicit operator bool() const { return p_ != nullptr; }
    int64_t use_count() const { return p_ ? p_->refs.load(std::memory_order_relaxed) : 0; }
};

// ---------------------------------------------------------------- storage / impl / tensor
// A block of HBM owned by the stream-ordered pool (runtime.cpp). device -1 = meta (no memory).
struct Storage : RefCounted {
    void *ptr = nullptr;
    size_t bytes = 0;
    int device = -1;
    bool external = false;  // memory not owned by the pool
    Storage(size_t bytes, int device);
    ~Storage() override;
};

class Tensor;
struct GradFunction : RefCounted {
    std::vector<Tensor> inputs;
    virtual std::vector<Tensor> backward(const Tensor &grad_output) = 0;
    virtual const char *name() const = 0;
};

struct TensorImpl : RefCounted {
    int ndim = 0;
    int64_t shape[KF_MAX_DIMS] = {0};
    int64_t stride[KF_MAX_DIMS] = {0};  // in elements
    DType dtype = KF_UNDEFINED;
    int64_t numel = 0;
    int64_t offset = 0;  // in elements
    Ref<Storage> storage;
    bool requires_grad = false;
    std::unique_ptr<Tensor> grad;

    bool is_contiguous() const;  // true row-major density (superset of the reference's flag, tensor_impl.cpp:95)
    void *data() const { return storage ? (char *)storage->ptr + offset * (int64_t)element_size(dtype) : nullptr; }
    int device() const { return storage ? storage->device : -1; }
    ~TensorImpl() override;
};

// Value handle: shares the impl, owns the grad_fn edge (ref: tensor.h:24-27 — grad_fn lives on the handle).
class Tensor {
public:
    Ref<TensorImpl> impl;
    Ref<GradFunction> grad_fn;

    bool defined() const { return impl && impl->storage; }
    int dim() const { return impl->ndim; }
    int64_t size(int64_t d) const { return impl->shape[wrap_dim(d, impl->ndim)]; }
    int64_t stride(int64_t d) const { return impl->stride[wrap_dim(d, impl->ndim)]; }
    std::vector<int64_t> sizes() const { return std::vector<int64_t>(impl->shape, impl->shape + impl->ndim); }
    std::vector<int64_t> strides() const { return std::vector<int64_t>(impl->stride, impl->stride + impl->ndim); }
    DType dtype() const { return impl->dtype; }
    int64_t numel() const { return impl->numel; }
    int device() const { return impl->device(); }
    void *data() const { return impl->data(); }
    template <class T> T *data_as() const { return reinterpret_cast<T *>(impl->data()); }
    bool is_contiguous() const { return impl->is_contiguous(); }
    bool is_meta() const { return impl->device() < 0; }
    size_t itemsize() const { return element_size(impl->dtype); }
    bool requires_grad() const { return impl->requires_grad; }

    // view algebra (tensor.cpp)
    Tensor as_strided(const std::vector<int64_t> &sizes, const std::vector<int64_t> &strides, int64_t storage_offset) const;
    Tensor permute(const std::vector<int64_t> &dims) const;
    Tensor view(std::vector<int64_t> sizes) const;
    Tensor slice(int64_t dim, int64_t start, int64_t end, int64_t step) const;
    Tensor select(int64_t dim, int64_t index) const;
    Tensor narrow(int64_t dim, int64_t start, int64_t length) const;
    Tensor transpose_last2() const;
    Tensor contiguous() const;
    Tensor detach() const;  // same storage/shape, fresh impl with requires_grad = false, no grad_fn
    std::string to_string() const;
};

Tensor empty(const std::vector<int64_t> &shape, DType dtype, int device);
Tensor empty_like(const Tensor &t);
Tensor zeros(const std::vector<int64_t> &shape, DType dtype, int device);
std::vector<int64_t> contiguous_strides(const std::vector<int64_t> &shape);

}  // namespace kf

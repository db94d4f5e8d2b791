// Tensor factories and zero-copy view algebra.
// Semantics follow the reference's src/core/tensor.cpp:17-321 and src/core/tensor_impl.cpp:11-102
// (same argument meaning, clamping and error conditions); contiguity is computed, not flagged.
#include <algorithm>
#include <iomanip>

#include "ops.h"
#include "runtime.h"

namespace kf {

TensorImpl::~TensorImpl() = default;

bool TensorImpl::is_contiguous() const {
    int64_t expect = 1;
    for (int i = ndim - 1; i >= 0; --i) {
        if (shape[i] == 1) continue;
        if (stride[i] != expect) return false;
        expect *= shape[i];
    }
    return true;
}

std::vector<int64_t> contiguous_strides(const std::vector<int64_t> &shape) {
    std::vector<int64_t> st(shape.size());
    int64_t c = 1;
    for (int i = (int)shape.size() - 1; i >= 0; --i) {
        st[i] = c;
        c *= shape[i];
    }
    return st;
}

static Tensor make_tensor(const std::vector<int64_t> &shape, const std::vector<int64_t> &strides, DType dtype, int device) {
    KF_CHECK(shape.size() <= KF_MAX_DIMS, "at most ", KF_MAX_DIMS, " dims");
    KF_CHECK(dtype >= 0 && dtype < KF_UNDEFINED, "bad dtype ", dtype);
    auto *impl = new TensorImpl();
    impl->ndim = (int)shape.size();
    impl->dtype = dtype;
    int64_t numel = 1, span = 1;
    for (int i = 0; i < impl->ndim; ++i) {
        KF_CHECK(shape[i] >= 0, "negative size");
        impl->shape[i] = shape[i];
        impl->stride[i] = strides[i];
        numel *= shape[i];
        span += (shape[i] - 1) * strides[i];
    }
    impl->numel = numel;
    if (numel == 0) span = 0;
    Tensor t;
    t.impl = Ref<TensorImpl>(impl);
    impl->storage = Ref<Storage>(new Storage((size_t)span * element_size(dtype), device));
    return t;
}

Tensor empty(const std::vector<int64_t> &shape, DType dtype, int device) {
    return make_tensor(shape, contiguous_strides(shape), dtype, device);
}
Tensor empty_like(const Tensor &t) { return empty(t.sizes(), t.dtype(), t.device()); }
Tensor zeros(const std::vector<int64_t> &shape, DType dtype, int device) {
    Tensor t = empty(shape, dtype, device);
    if (device >= 0) Runtime::get().memset_async(t.data(), 0, t.impl->storage->bytes);
    return t;
}

Tensor Tensor::as_strided(const std::vector<int64_t> &sizes, const std::vector<int64_t> &strides_in, int64_t storage_offset) const {
    KF_CHECK(sizes.size() <= KF_MAX_DIMS);
    std::vector<int64_t> strides = strides_in.empty() ? contiguous_strides(sizes) : strides_in;
    KF_CHECK(sizes.size() == strides.size());
    // in-bounds check (ref: tensor_impl.cpp:80-85)
    int64_t lo = storage_offset, hi = storage_offset, numel = 1;
    for (size_t i = 0; i < sizes.size(); ++i) {
        KF_CHECK(sizes[i] >= 0);
        numel *= sizes[i];
        if (sizes[i] > 0) (strides[i] >= 0 ? hi : lo) += (sizes[i] - 1) * strides[i];
    }
    if (numel > 0) {
        KF_CHECK(lo >= 0, "view starts before the storage");
        KF_CHECK((size_t)(hi + 1) * element_size(impl->dtype) <= impl->storage->bytes, "view exceeds the storage");
    }
    auto *ni = new TensorImpl();
    ni->ndim = (int)sizes.size();
    ni->dtype = impl->dtype;
    ni->numel = numel;
    ni->offset = storage_offset;
    ni->storage = impl->storage;
    ni->requires_grad = impl->requires_grad;
    for (int i = 0; i < ni->ndim; ++i) {
        ni->shape[i] = sizes[i];
        ni->stride[i] = strides[i];
    }
    Tensor out;
    out.impl = Ref<TensorImpl>(ni);
    return out;
}

Tensor Tensor::detach() const {
    if (!impl) return Tensor();
    if (!impl->requires_grad && !grad_fn) return *this;
    Tensor out = as_strided(sizes(), strides(), impl->offset);
    out.impl->requires_grad = false;
    return out;
}

Tensor Tensor::permute(const std::vector<int64_t> &dims) const {
    const int nd = dim();
    KF_CHECK((int)dims.size() == nd, "permute(): expected ", nd, " dims");
    std::vector<int64_t> ns(nd), nst(nd);
    std::vector<bool> seen(nd, false);
    for (int i = 0; i < nd; ++i) {
        int d = wrap_dim(dims[i], nd);
        KF_CHECK(!seen[d], "permute(): duplicate dims are not allowed.");
        seen[d] = true;
        ns[i] = impl->shape[d];
        nst[i] = impl->stride[d];
    }
    return as_strided(ns, nst, impl->offset);
}

Tensor Tensor::view(std::vector<int64_t> sizes) const {
    KF_CHECK(is_contiguous(), "view() needs a contiguous tensor");
    int64_t prod = 1;
    int neg = -1;
    for (size_t i = 0; i < sizes.size(); ++i) {
        if (sizes[i] < 0) {
            KF_CHECK(neg < 0, "only one dimension can be inferred");
            neg = (int)i;
        } else {
            prod *= sizes[i];
        }
    }
    if (neg >= 0) {
        KF_CHECK(prod > 0 && numel() % prod == 0, "view(): cannot infer dimension");
        sizes[neg] = numel() / prod;
        prod *= sizes[neg];
    }
    KF_CHECK(prod == numel(), "view(): shape is invalid for input of size ", numel());
    return as_strided(sizes, {}, impl->offset);
}

Tensor Tensor::slice(int64_t dim_, int64_t start, int64_t end, int64_t step) const {
    const int d = wrap_dim(dim_, dim());
    KF_CHECK(step > 0, "slice step must be positive");
    auto sz = sizes();
    auto st = strides();
    if (start < 0) start += sz[d];
    if (end < 0) end += sz[d];
    if (start < 0) start = 0; else if (start >= sz[d]) start = sz[d];
    if (end < start) end = start; else if (end >= sz[d]) end = sz[d];
    const int64_t off = impl->offset + start * st[d];
    sz[d] = (end - start + step - 1) / step;
    st[d] *= step;
    return as_strided(sz, st, off);
}

Tensor Tensor::select(int64_t dim_, int64_t index) const {
    KF_CHECK(dim() > 0, "select() cannot be applied to a 0-dim tensor.");
    const int d = wrap_dim(dim_, dim());
    const int64_t size = impl->shape[d];
    KF_CHECK(index >= -size && index < size, "select(): index ", index, " out of range for size ", size);
    if (index < 0) index += size;
    auto sz = sizes();
    auto st = strides();
    const int64_t off = impl->offset + index * st[d];
    sz.erase(sz.begin() + d);
    st.erase(st.begin() + d);
    return as_strided(sz, st, off);
}

Tensor Tensor::narrow(int64_t dim_, int64_t start, int64_t length) const {
    KF_CHECK(dim() > 0, "narrow() cannot be applied to a 0-dim tensor.");
    KF_CHECK(length >= 0, "narrow(): length must be non-negative.");
    const int d = wrap_dim(dim_, dim());
    const int64_t cur = impl->shape[d];
    if (start < 0) start += cur;
    KF_CHECK(start >= 0 && start <= cur - length, "start (", start, ") + length (", length, ") exceeds dimension size (", cur, ").");
    return slice(d, start, start + length, 1);
}

Tensor Tensor::transpose_last2() const {
    KF_CHECK(dim() >= 2);
    std::vector<int64_t> p(dim());
    for (int i = 0; i < dim(); ++i) p[i] = i;
    std::swap(p[dim() - 1], p[dim() - 2]);
    return permute(p);
}

Tensor Tensor::contiguous() const {
    if (is_contiguous()) return *this;
    return ops::clone(*this);
}

// ---- printing (same information as the reference's operator<<, tensor.cpp:323-377)
static void print_rec(std::ostream &os, const Tensor &t, const std::vector<double> &vals, std::vector<int64_t> &idx, int d,
                      const std::vector<int64_t> &lim) {
    if (d == t.dim()) {
        int64_t flat = 0, mul = 1;
        for (int i = t.dim() - 1; i >= 0; --i) {
            flat += idx[i] * mul;
            mul *= lim[i];
        }
        os << std::fixed << std::showpos << std::setprecision(5) << vals[flat] << std::noshowpos;
        return;
    }
    if (d > 0) os << "\n";
    for (int i = -1; i < d; i++) os << "  ";
    os << "[";
    for (int64_t ii = 0; ii < lim[d]; ii++) {
        if (ii > 0) os << ", ";
        idx[d] = ii;
        print_rec(os, t, vals, idx, d + 1, lim);
    }
    if (t.size(d) > lim[d]) os << ", ...";
    if (d < t.dim() - 1) {
        os << "\n";
        for (int i = -1; i < d; i++) os << "  ";
    }
    os << "]";
}

std::string Tensor::to_string() const {
    std::ostringstream os;
    if (!defined()) return "Tensor(Undefined)";
    os << "tensor(shape=[";
    for (int i = 0; i < dim(); ++i) os << (i ? "," : "") << size(i);
    os << "], stride=[";
    for (int i = 0; i < dim(); ++i) os << (i ? "," : "") << stride(i);
    os << "], storage_offset=" << impl->offset << ", dtype=" << dtype_name(dtype()) << ", numel=" << numel() << ", dim=" << dim()
       << ", device=" << device() << ")";
    if (is_meta() || numel() == 0) return os.str();
    // materialise the leading <=12 entries per dim as doubles through the device (one small D2H)
    Tensor v = *this;
    std::vector<int64_t> lim(dim());
    for (int i = 0; i < dim(); ++i) {
        lim[i] = std::min<int64_t>(size(i), 12);
        v = v.slice(i, 0, lim[i], 1);
    }
    Tensor d = ops::convert(v.detach(), KF_DOUBLE).contiguous();
    std::vector<double> vals((size_t)d.numel());
    Runtime::get().d2h(vals.data(), d.data(), vals.size() * sizeof(double), true);
    os << " {\n";
    std::vector<int64_t> idx(dim(), 0);
    print_rec(os, *this, vals, idx, 0, lim);
    os << "\n}";
    return os.str();
}

}  // namespace kf

// Elementwise planner: broadcast, dtype promotion, output allocation, dimension reorder + collapse.
// The B200 replacement for the reference's TensorIterator (src/core/tensor_iterator.cpp:4-528):
// same rules (equal ndim, size-1 expansion, promotion table, fresh outputs contiguous in logical order,
// coalesce when shape*stride == next stride), emitted as one POD plan the kernels take by value.
#include <algorithm>
#include <numeric>

#include "ops.h"

namespace kf {
namespace ops {

void plan_elementwise(EwPlan &plan, Tensor &out, const Tensor *a, const Tensor *b, bool) {
    const Tensor *ts[3] = {out.defined() ? &out : nullptr, a, b};
    // ---- device + ndim agreement (ref: tensor_iterator.cpp:4-30)
    int ndim = -1, device = -2;
    for (const Tensor *t : ts) {
        if (!t) continue;
        KF_CHECK(t->defined(), "undefined tensor operand");
        if (ndim < 0) ndim = t->dim();
        KF_CHECK(ndim == t->dim(), "All defined tensors should in the same dim");
        if (device == -2) device = t->device();
        KF_CHECK(device == t->device(), "All defined tensors should in the same device");
    }
    KF_CHECK(ndim >= 0, "no operands");
    // ---- common dtype over the inputs (ref: tensor_iterator.cpp:46-58)
    DType common = KF_UNDEFINED;
    for (int i = 1; i < 3; ++i) {
        if (!ts[i]) continue;
        common = common == KF_UNDEFINED ? ts[i]->dtype() : promote(common, ts[i]->dtype());
    }
    if (common == KF_UNDEFINED) common = out.dtype();
    // ---- broadcast shape (ref: tensor_iterator.cpp:110-127)
    std::vector<int64_t> shape(ndim, 1);
    for (int d = 0; d < ndim; ++d) {
        int64_t sz = 1;
        bool first = true;
        for (const Tensor *t : ts) {
            if (!t) continue;
            const int64_t s = t->impl->shape[d];
            if (first) {
                sz = s;
                first = false;
            } else {
                KF_CHECK(sz == s || sz == 1 || s == 1, "shapes are not broadcastable at dim ", d, ": ", sz, " vs ", s);
                sz = sz == 1 ? s : sz;
            }
        }
        shape[d] = sz;
    }
    if (out.defined()) {
        for (int d = 0; d < ndim; ++d)
            KF_CHECK(out.impl->shape[d] == shape[d], "output with shape ", out.impl->shape[d], " at dim ", d,
                     " doesn't match the broadcast shape ", shape[d]);
    } else {
        out = empty(shape, common, device);
    }
    ts[0] = &out;
    // ---- byte strides with broadcast zeros (ref: tensor_iterator.cpp:148-162)
    int64_t st[3][KF_MAX_DIMS] = {{0}};
    int64_t numel = 1;
    for (int d = 0; d < ndim; ++d) numel *= shape[d];
    for (int i = 0; i < 3; ++i) {
        if (!ts[i]) continue;
        const int64_t isz = (int64_t)ts[i]->itemsize();
        for (int d = 0; d < ndim; ++d)
            st[i][d] = (ts[i]->impl->shape[d] == 1 && shape[d] != 1) ? 0 : ts[i]->impl->stride[d] * isz;
    }
    // ---- order dims fastest-first by the output's stride; drop size-1 dims
    std::vector<int> order;
    for (int d = ndim - 1; d >= 0; --d)
        if (shape[d] != 1) order.push_back(d);
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return std::llabs(st[0][x]) < std::llabs(st[0][y]); });
    // ---- collapse (ref: tensor_iterator.cpp:263-307)
    plan = EwPlan{};
    int nd = 0;
    for (int d : order) {
        bool merge = nd > 0;
        if (merge) {
            for (int i = 0; i < 3 && merge; ++i)
                if (ts[i] && plan.stride[i][nd - 1] * plan.shape[nd - 1] != st[i][d]) merge = false;
        }
        if (merge) {
            plan.shape[nd - 1] *= shape[d];
        } else {
            plan.shape[nd] = shape[d];
            for (int i = 0; i < 3; ++i) plan.stride[i][nd] = ts[i] ? st[i][d] : 0;
            ++nd;
        }
    }
    if (nd == 0) {  // scalar / all-ones shape
        plan.shape[0] = 1;
        for (int i = 0; i < 3; ++i) plan.stride[i][0] = ts[i] ? (int64_t)ts[i]->itemsize() : 0;
        nd = 1;
    }
    plan.ndim = nd;
    plan.numel = numel;
    plan.nin = (a ? 1 : 0) + (b ? 1 : 0);
    for (int i = 0; i < 3; ++i) {
        plan.ptr[i] = ts[i] ? ts[i]->data() : nullptr;
        plan.dtype[i] = ts[i] ? ts[i]->dtype() : KF_UNDEFINED;
    }
    plan.acc = acc_kind(common);
    plan.b_is_scalar = 0;
    plan.scalar = 0.0;
}

}  // namespace ops
}  // namespace kf

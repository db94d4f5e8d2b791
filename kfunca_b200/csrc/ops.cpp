// Operator front-ends.  One function per reference `gpu::` entry (src/core/{binary,unary,nullary,reduce,
// sort,gemm,nn,index}_ops.cpp, tensor_shape.cpp) with the same argument meaning and error conditions,
// plus the GradFunctions the reference never wrote (only AddGradFunction exists there,
// src/core/binary_ops.cpp:16-43) so that the transformer block can run backward.
#include "ops.h"

#include <algorithm>
#include <cmath>
#include <queue>
#include <unordered_map>

#include "runtime.h"

namespace kf {
namespace ops {

static void require_device(const Tensor &t, const char *what) {
    KF_CHECK(t.defined(), what, ": undefined tensor");
    KF_CHECK(!t.is_meta(), what, ": meta tensors carry no data (kfunca_b200 has no CPU compute path)");
}

static bool any_requires_grad(std::initializer_list<const Tensor *> ts) {
    for (auto *t : ts)
        if (t && t->defined() && t->requires_grad()) return true;
    return false;
}
template <class F>
static void attach(Tensor &out, F *fn, std::initializer_list<Tensor> inputs) {
    for (auto &t : inputs) fn->inputs.push_back(t);
    out.impl->requires_grad = true;
    out.grad_fn = Ref<GradFunction>(fn);
}

// scalar -> acc_t -> dtype -> acc_t on the host (what the reference gets by filling a tensor, register.cpp:172-206)
static double round_scalar_through(DType dt, double v) {
    switch (dt) {
    case KF_DOUBLE: return v;
    case KF_FLOAT: return (double)(float)v;
    case KF_HALF: return (double)f16_bits_to_f32(f32_to_f16_bits((float)v));
    case KF_BFLOAT16: return (double)bf16_bits_to_f32(f32_to_bf16_bits((float)v));
    case KF_BOOL: return v != 0.0 ? 1.0 : 0.0;
    case KF_BYTE: return (double)(uint8_t)(int64_t)v;
    case KF_CHAR: return (double)(int8_t)(int64_t)v;
    case KF_SHORT: return (double)(int16_t)(int64_t)v;
    case KF_INT: return (double)(int32_t)(int64_t)v;
    default: return (double)(int64_t)v;
    }
}

// ================================================================== elementwise
static void run_binary(int op, Tensor &out, const Tensor &a, const Tensor &b) {
    require_device(a, "binary op");
    require_device(b, "binary op");
    EwPlan plan;
    plan_elementwise(plan, out, &a, &b, true);
    plan.op = op;
    launch_elementwise(plan);
}

struct BinaryGrad : GradFunction {
    int op;
    Tensor a, b;  // detached saved operands (mul/div)
    std::vector<int64_t> sa, sb;
    const char *name() const override { return "BinaryGrad"; }
    std::vector<Tensor> backward(const Tensor &g) override {
        std::vector<Tensor> r(2);
        const bool na = inputs[0].requires_grad(), nb = inputs[1].requires_grad();
        switch (op) {
        case EW_ADD:
            if (na) r[0] = sum_to_shape(g, sa);
            if (nb) r[1] = sum_to_shape(g, sb);
            break;
        case EW_SUB:
            if (na) r[0] = sum_to_shape(g, sa);
            if (nb) r[1] = sum_to_shape(unary(EW_NEG, g), sb);
            break;
        case EW_MUL:
            if (na) r[0] = sum_to_shape(binary(EW_MUL, g, b), sa);
            if (nb) r[1] = sum_to_shape(binary(EW_MUL, g, a), sb);
            break;
        default: {  // a / b
            Tensor gq = binary(EW_DIV, g, b);
            if (na) r[0] = sum_to_shape(gq, sa);
            if (nb) r[1] = sum_to_shape(unary(EW_NEG, binary(EW_MUL, gq, binary(EW_DIV, a, b))), sb);
        }
        }
        return r;
    }
};

Tensor binary(int op, const Tensor &a, const Tensor &b) {
    Tensor out;
    run_binary(op, out, a, b);
    if (any_requires_grad({&a, &b})) {
        auto *fn = new BinaryGrad();
        fn->op = op;
        fn->sa = a.sizes();
        fn->sb = b.sizes();
        if (op == EW_MUL || op == EW_DIV) {
            fn->a = a.detach();
            fn->b = b.detach();
        }
        attach(out, fn, {a, b});
    }
    return out;
}

Tensor &binary_(int op, Tensor &self, const Tensor &other) {
    run_binary(op, self, self, other);
    return self;
}

static void run_binary_scalar(int op, Tensor &out, const Tensor &a, double scalar) {
    require_device(a, "binary op");
    EwPlan plan;
    plan_elementwise(plan, out, &a, nullptr, true);
    plan.op = op;
    plan.nin = 2;
    plan.b_is_scalar = 1;
    plan.scalar = round_scalar_through(a.dtype(), scalar);
    launch_elementwise(plan);
}

struct ScalarGrad : GradFunction {
    int op;
    double s;
    const char *name() const override { return "ScalarGrad"; }
    std::vector<Tensor> backward(const Tensor &g) override {
        if (op == EW_MUL || op == EW_DIV) return {binary_scalar(op, g, s)};
        return {g};
    }
};

Tensor binary_scalar(int op, const Tensor &a, double scalar) {
    Tensor out;
    run_binary_scalar(op, out, a, scalar);
    if (a.requires_grad()) {
        auto *fn = new ScalarGrad();
        fn->op = op;
        fn->s = scalar;
        attach(out, fn, {a});
    }
    return out;
}
Tensor &binary_scalar_(int op, Tensor &self, double scalar) {
    run_binary_scalar(op, self, self, scalar);
    return self;
}

Tensor &fill_(Tensor &self, double value) {
    require_device(self, "fill_");
    EwPlan plan;
    plan_elementwise(plan, self, nullptr, nullptr, false);
    plan.op = EW_FILL;
    plan.acc = acc_kind(self.dtype());
    plan.scalar = value;
    launch_elementwise(plan);
    return self;
}

// copy with optional dtype cast; picks the tiled transpose when the layouts call for it
static void run_copy(Tensor &dst, const Tensor &src) {
    require_device(dst, "copy_");
    require_device(src, "copy_");
    EwPlan plan;
    Tensor d = dst;
    plan_elementwise(plan, d, &src, nullptr, false);
    plan.op = EW_COPY;
    plan.acc = acc_kind(dst.dtype());
    if (plan.numel == 0) return;
    const int64_t isz = (int64_t)dst.itemsize();
    if (dst.dtype() == src.dtype() && plan.ndim >= 2 && plan.stride[0][0] == isz && plan.stride[1][0] != isz) {
        int tdim = -1;
        for (int dd = 1; dd < plan.ndim; ++dd)
            if (plan.stride[1][dd] == isz) tdim = dd;
        bool ok = tdim > 0 && plan.shape[0] >= 8 && plan.shape[tdim] >= 8;
        for (int dd = 0; dd < plan.ndim && ok; ++dd)
            if (plan.stride[0][dd] % isz || plan.stride[1][dd] % isz || plan.stride[0][dd] < 0 || plan.stride[1][dd] < 0) ok = false;
        if (ok) {
            TransposePlan tp{};
            tp.ndim = plan.ndim;
            tp.tdim = tdim;
            tp.itemsize = (int)isz;
            for (int dd = 0; dd < plan.ndim; ++dd) {
                tp.shape[dd] = plan.shape[dd];
                tp.out_stride[dd] = plan.stride[0][dd] / isz;
                tp.in_stride[dd] = plan.stride[1][dd] / isz;
            }
            tp.in = plan.ptr[1];
            tp.out = plan.ptr[0];
            launch_transpose(tp);
            return;
        }
    }
    launch_elementwise(plan);
}

Tensor &copy_(Tensor &self, const Tensor &src) {
    run_copy(self, src);
    return self;
}

struct IdentityGrad : GradFunction {  // clone / contiguous
    const char *name() const override { return "IdentityGrad"; }
    std::vector<Tensor> backward(const Tensor &g) override { return {g}; }
};

Tensor clone(const Tensor &self) {
    Tensor out = empty_like(self);
    run_copy(out, self);
    if (self.requires_grad()) attach(out, new IdentityGrad(), {self});
    return out;
}
Tensor contiguous(const Tensor &self) {
    if (self.is_contiguous()) return self;
    return clone(self);
}

struct ConvertGrad : GradFunction {
    DType src;
    const char *name() const override { return "ConvertGrad"; }
    std::vector<Tensor> backward(const Tensor &g) override { return {convert(g, src)}; }
};
Tensor convert(const Tensor &self, DType dtype) {
    require_device(self, "convert");
    Tensor out = empty(self.sizes(), dtype, self.device());
    run_copy(out, self);
    if (self.requires_grad() && is_floating(dtype)) {
        auto *fn = new ConvertGrad();
        fn->src = self.dtype();
        attach(out, fn, {self});
    }
    return out;
}

struct UnaryGrad : GradFunction {
    int op;
    Tensor y;  // saved output
    const char *name() const override { return "UnaryGrad"; }
    std::vector<Tensor> backward(const Tensor &g) override {
        if (op == EW_NEG) return {unary(EW_NEG, g)};
        if (op == EW_SQRT) return {binary(EW_DIV, binary_scalar(EW_MUL, g, 0.5), y)};  // g / (2 sqrt(x))
        // rsqrt: d/dx x^-1/2 = -1/2 y^3
        return {binary_scalar(EW_MUL, binary(EW_MUL, g, binary(EW_MUL, y, binary(EW_MUL, y, y))), -0.5)};
    }
};
Tensor unary(int op, const Tensor &a) {
    require_device(a, "unary op");
    Tensor out;
    EwPlan plan;
    plan_elementwise(plan, out, &a, nullptr, true);
    plan.op = op;
    launch_elementwise(plan);
    if (a.requires_grad()) {
        auto *fn = new UnaryGrad();
        fn->op = op;
        fn->y = out.detach();
        attach(out, fn, {a});
    }
    return out;
}

// ================================================================== reductions
struct ReduceGrad : GradFunction {
    std::vector<int64_t> in_shape;
    bool is_mean;
    int64_t R;
    const char *name() const override { return "ReduceGrad"; }
    std::vector<Tensor> backward(const Tensor &g) override {
        Tensor gi = empty(in_shape, g.dtype(), g.device());
        Tensor src = is_mean ? binary_scalar(EW_MUL, g, 1.0 / (double)R) : g;
        run_copy(gi, src);  // broadcast along the reduced dim
        return {gi};
    }
};

static Tensor reduce_impl(const Tensor &self, int64_t dim_, bool is_mean) {
    require_device(self, "reduce");
    const int d = wrap_dim(dim_, self.dim());
    Tensor x = self.is_contiguous() ? self : clone(self.detach());
    auto out_shape = self.sizes();
    out_shape[d] = 1;
    Tensor out = empty(out_shape, self.dtype(), self.device());
    ReducePlan p{};
    p.in = x.data();
    p.out = out.data();
    p.dtype = self.dtype();
    p.outer = 1;
    p.inner = 1;
    for (int i = 0; i < d; ++i) p.outer *= self.size(i);
    for (int i = d + 1; i < self.dim(); ++i) p.inner *= self.size(i);
    p.R = self.size(d);
    p.is_mean = is_mean;
    const int64_t n_out = p.outer * p.inner, numel = self.numel();
    if (is_mean) {
        // ref: factor = static_cast<acc_t>(num_output_elements) / numel in the tensor's own arithmetic
        // (reduce_ops_kernel.cu:49-53): integer dtypes integer-divide; 16-bit floats use fp32 here (deviation).
        switch (self.dtype()) {
        case KF_DOUBLE: p.factor = numel ? (double)n_out / (double)numel : 0.0; break;
        case KF_FLOAT: case KF_HALF: case KF_BFLOAT16: p.factor = numel ? (double)((float)n_out / (float)numel) : 0.0; break;
        case KF_BOOL: p.factor = (numel && ((n_out != 0 ? 1 : 0) / numel) != 0) ? 1.0 : 0.0; break;
        case KF_BYTE: p.factor = numel ? (double)(uint8_t)((int64_t)(uint8_t)n_out / numel) : 0.0; break;
        case KF_CHAR: p.factor = numel ? (double)(int8_t)((int64_t)(int8_t)n_out / numel) : 0.0; break;
        case KF_SHORT: p.factor = numel ? (double)(int16_t)((int64_t)(int16_t)n_out / numel) : 0.0; break;
        case KF_INT: p.factor = numel ? (double)(int32_t)((int64_t)(int32_t)n_out / numel) : 0.0; break;
        default: p.factor = numel ? (double)(n_out / numel) : 0.0; break;
        }
    }
    if (p.R == 0) {
        fill_(out, 0.0);
    } else {
        launch_reduce(p);
    }
    if (self.requires_grad()) {
        auto *fn = new ReduceGrad();
        fn->in_shape = self.sizes();
        fn->is_mean = is_mean;
        fn->R = p.R;
        attach(out, fn, {self});
    }
    return out;
}
Tensor sum(const Tensor &self, int64_t dim) { return reduce_impl(self, dim, false); }
Tensor mean(const Tensor &self, int64_t dim) { return reduce_impl(self, dim, true); }

Tensor sum_to_shape(const Tensor &grad, const std::vector<int64_t> &shape) {
    Tensor g = grad;
    KF_CHECK((int)shape.size() == g.dim());
    for (int d = 0; d < g.dim(); ++d)
        if (shape[d] == 1 && g.size(d) != 1) g = sum(g.detach(), d);
    return g;
}

std::tuple<Tensor, Tensor> mean_var(const Tensor &self, int64_t dim, bool take_sqrt) {
    // ref: gpu::mean_var (reduce_ops.cpp:22-28) = Welford with correction 1; fp32/fp64 only (DISPATCH_FLOATING_TYPES)
    KF_CHECK(self.dtype() == KF_FLOAT || self.dtype() == KF_DOUBLE, "Unsupported ScalarType ", dtype_name(self.dtype()));
    const int d = wrap_dim(dim, self.dim());
    Tensor x = self.detach();
    if (d == self.dim() - 1 && x.is_contiguous() && x.device() >= 0 && x.size(d) >= 2) {  // one-pass row statistics (kernels/norm.cu)
        auto oshape = x.sizes();
        oshape[d] = 1;
        Tensor m = empty(oshape, x.dtype(), x.device()), v = empty(oshape, x.dtype(), x.device());
        if (launch_row_moments(x.data(), m.data(), v.data(), x.dtype(), x.numel() / x.size(d), x.size(d), take_sqrt)) return {m, v};
    }
    if (d != self.dim() - 1 && x.dtype() == KF_FLOAT && x.device() >= 0 && x.numel() > 0) {  // one-launch column statistics (kernels/norm.cu)
        Tensor xc = x.is_contiguous() ? x : clone(x);
        int64_t outer = 1, inner = 1;
        for (int i = 0; i < d; ++i) outer *= xc.size(i);
        for (int i = d + 1; i < xc.dim(); ++i) inner *= xc.size(i);
        auto oshape = xc.sizes();
        oshape[d] = 1;
        Tensor m = empty(oshape, KF_FLOAT, xc.device()), v = empty(oshape, KF_FLOAT, xc.device());
        if (launch_col_moments(xc.data(), m.data(), v.data(), KF_FLOAT, outer, xc.size(d), inner, 1, take_sqrt, 0.0)) return {m, v};
    }
    Tensor m = mean(x, d);
    Tensor diff = binary(EW_SUB, x, m);
    Tensor m2 = sum(binary(EW_MUL, diff, diff), d);
    const double div = (double)self.size(d) - 1.0;
    Tensor var = binary_scalar(EW_DIV, m2, div > 0 ? div : 0.0);
    if (take_sqrt) var = unary(EW_SQRT, var);
    return {m, var};
}

std::tuple<Tensor, Tensor> norm_stat(const Tensor &self, int64_t dim) {
    // ref: norm_stat_kernel (src/device/norm_ops_kernel.cu:6-61): 2-D, dim 0, biased variance, eps 1e-12
    KF_CHECK(self.defined());
    KF_CHECK(dim == 0 && self.dim() == 2);
    KF_CHECK(self.dtype() == KF_FLOAT || self.dtype() == KF_DOUBLE, "Unsupported ScalarType ", dtype_name(self.dtype()));
    Tensor x = self.detach();
    if (x.dtype() == KF_FLOAT && x.device() >= 0 && x.numel() > 0) {  // ONE launch (the reference: WelfordNormPFKernel + semaphores)
        Tensor xc = x.is_contiguous() ? x : clone(x);
        Tensor m = empty({1, xc.size(1)}, KF_FLOAT, xc.device()), is = empty({1, xc.size(1)}, KF_FLOAT, xc.device());
        if (launch_col_moments(xc.data(), m.data(), is.data(), KF_FLOAT, 1, xc.size(0), xc.size(1), 0, false, 1e-12)) return {m, is};
    }
    Tensor m = mean(x, 0);
    Tensor diff = binary(EW_SUB, x, m);
    Tensor var = mean(binary(EW_MUL, diff, diff), 0);
    Tensor invstd = unary(EW_RSQRT, binary_scalar(EW_ADD, var, 1e-12));
    return {m, invstd};
}

// ================================================================== fused layer norm
struct LayerNormGrad : GradFunction {
    Tensor x, gain, stats;  // stats: fp32 [2, rows] = mean | rstd
    int64_t rows, E;
    bool rms = false;
    const char *name() const override { return "LayerNormGrad"; }
    std::vector<Tensor> backward(const Tensor &g0) override {
        Tensor g = g0.is_contiguous() ? g0 : clone(g0.detach());
        if (reinterpret_cast<uintptr_t>(g.data()) % 16 != 0) g = clone(g.detach());  // the kernels use 16-byte vector loads
        const bool need_dx = inputs[0].requires_grad();
        Tensor dx;
        if (need_dx) dx = empty(x.sizes(), x.dtype(), x.device());
        const int ctas = layer_norm_bwd_ctas(rows, need_dx);
        Tensor partial = rows > 0 ? empty({(int64_t)ctas, E}, KF_FLOAT, x.device()) : zeros({(int64_t)ctas, E}, KF_FLOAT, x.device());
        const float *st = reinterpret_cast<const float *>(stats.data());
        launch_layer_norm_bwd(x.data(), gain.data(), g.data(), st, st + rows, need_dx ? dx.data() : nullptr,
                              reinterpret_cast<float *>(partial.data()), ctas, x.dtype(), rows, E, rms);
        Tensor dgain = convert(sum(partial, 0), x.dtype()).view(gain.sizes());
        return {dx, dgain};
    }
};

static Tensor norm_impl(const Tensor &x_, const Tensor &gain_, double eps, bool rms);
Tensor layer_norm(const Tensor &x, const Tensor &gain, double eps) { return norm_impl(x, gain, eps, false); }
// y = x / sqrt(mean(x^2) + eps) * gain — the op the reference's README names as its next one (README.md:28 `rms_norm`)
Tensor rms_norm(const Tensor &x, const Tensor &gain, double eps) { return norm_impl(x, gain, eps, true); }

static Tensor norm_impl(const Tensor &x_, const Tensor &gain_, double eps, bool rms) {
    require_device(x_, "layer_norm");
    require_device(gain_, "layer_norm");
    KF_CHECK(x_.dim() >= 1 && x_.dtype() == gain_.dtype(), "layer_norm: x and gain must share a dtype");
    const int64_t E = x_.size(-1);
    KF_CHECK(gain_.numel() == E, "layer_norm: gain must have ", E, " elements");
    Tensor x = x_.is_contiguous() ? x_ : contiguous(x_);
    Tensor gain = gain_.is_contiguous() ? gain_ : contiguous(gain_);
    if (!layer_norm_supported(x.dtype(), E, x.data(), gain.data())) {
        // composed form (same arithmetic, several passes): rows too long for the register-resident kernels, or fp64
        std::vector<int64_t> gshape(x.dim(), 1);
        gshape.back() = E;
        Tensor xc = rms ? x : binary(EW_SUB, x, mean(x, -1));
        Tensor var = mean(binary(EW_MUL, xc, xc), -1);
        Tensor rstd = unary(EW_RSQRT, binary_scalar(EW_ADD, var, eps));
        return binary(EW_MUL, binary(EW_MUL, xc, rstd), view(gain, gshape));
    }
    const int64_t rows = E ? x.numel() / E : 0;
    Tensor out = empty(x.sizes(), x.dtype(), x.device());
    Tensor stats = empty({2, rows}, KF_FLOAT, x.device());
    float *st = reinterpret_cast<float *>(stats.data());
    launch_layer_norm_fwd(x.data(), gain.data(), out.data(), st, st + rows, x.dtype(), rows, E, (float)eps, rms);
    if (any_requires_grad({&x, &gain})) {
        auto *fn = new LayerNormGrad();
        fn->x = x.detach();
        fn->gain = gain.detach();
        fn->stats = stats;
        fn->rows = rows;
        fn->E = E;
        fn->rms = rms;
        attach(out, fn, {x, gain});
    }
    return out;
}

// ================================================================== sort / top-k
static Tensor move_dim_last(const Tensor &t, int d) {
    std::vector<int64_t> perm;
    for (int i = 0; i < t.dim(); ++i)
        if (i != d) perm.push_back(i);
    perm.push_back(d);
    return t.permute(perm);
}
static Tensor move_last_to(const Tensor &t, int d) {
    std::vector<int64_t> perm(t.dim());
    int src = 0;
    for (int i = 0; i < t.dim(); ++i) perm[i] = (i == d) ? t.dim() - 1 : src++;
    return t.permute(perm);
}

std::tuple<Tensor, Tensor> sort(const Tensor &self, int64_t dim_, bool descending) {
    require_device(self, "sort");
    const int d = wrap_dim(dim_, self.dim());
    KF_CHECK(self.dtype() != KF_BOOL, "Sort currently does not support bool dtypes.");
    const int64_t n = self.size(d);
    KF_CHECK(n <= 0x7FFFFFFF, "The dimension being sorted can not have more than INT_MAX elements.");
    Tensor rows = move_dim_last(self.detach(), d).contiguous();
    Tensor vals = empty(rows.sizes(), self.dtype(), self.device());
    Tensor idx = empty(rows.sizes(), KF_LONG, self.device());
    if (self.numel() > 0) launch_sort_rows(rows.data(), vals.data(), idx.data_as<int64_t>(), self.dtype(), self.numel() / n, n, descending);
    if (d == self.dim() - 1) return {vals, idx};
    return {move_last_to(vals, d).contiguous(), move_last_to(idx, d).contiguous()};
}

std::tuple<Tensor, Tensor> topk(const Tensor &self, int64_t k, int64_t dim_, bool largest) {
    require_device(self, "topk");
    const int d = wrap_dim(dim_, self.dim());
    KF_CHECK(self.dtype() != KF_BOOL, "Sort currently does not support bool dtypes.");
    const int64_t n = self.size(d);
    KF_CHECK(k >= 0 && k <= n, "start (0) + length (", k, ") exceeds dimension size (", n, ").");
    Tensor rows = move_dim_last(self.detach(), d).contiguous();
    auto out_shape = rows.sizes();
    out_shape.back() = k;
    Tensor vals = empty(out_shape, self.dtype(), self.device());
    Tensor idx = empty(out_shape, KF_LONG, self.device());
    const int64_t nseg = n ? self.numel() / n : 0;
    if (nseg > 0 && k > 0) {
        if (!launch_topk_rows(rows.data(), vals.data(), idx.data_as<int64_t>(), self.dtype(), nseg, n, k, largest)) {
            // reference algorithm: full stable sort, keep the first k (sort_ops_kernel.cu:617-632)
            Tensor sv = empty(rows.sizes(), self.dtype(), self.device());
            Tensor si = empty(rows.sizes(), KF_LONG, self.device());
            launch_sort_rows(rows.data(), sv.data(), si.data_as<int64_t>(), self.dtype(), nseg, n, largest);
            run_copy(vals, sv.narrow(-1, 0, k));
            run_copy(idx, si.narrow(-1, 0, k));
        }
    }
    if (d == self.dim() - 1) return {vals, idx};
    return {move_last_to(vals, d).contiguous(), move_last_to(idx, d).contiguous()};
}

// ================================================================== shape ops
struct CatGrad : GradFunction {
    int dim;
    std::vector<int64_t> sizes;
    const char *name() const override { return "CatGrad"; }
    std::vector<Tensor> backward(const Tensor &g) override {
        std::vector<Tensor> r;
        int64_t off = 0;
        for (auto s : sizes) {
            r.push_back(g.narrow(dim, off, s));
            off += s;
        }
        return r;
    }
};

Tensor cat(const std::vector<Tensor> &tensors, int64_t dim_) {
    KF_CHECK(!tensors.empty(), "cat(): empty tensor list");
    const Tensor &first = tensors[0];
    require_device(first, "cat");
    const int d = wrap_dim(dim_, first.dim());
    int64_t total = first.size(d);
    bool rg = first.requires_grad();
    for (size_t i = 1; i < tensors.size(); ++i) {
        const Tensor &t = tensors[i];
        KF_CHECK(first.device() == t.device());
        KF_CHECK(first.dim() == t.dim(), "Tensors must have same number of dimensions: got ", first.dim(), " and ", t.dim());
        for (int k = 0; k < first.dim(); ++k) {
            if (k == d) continue;
            KF_CHECK(first.size(k) == t.size(k), "Sizes of tensors must match except in dimension ", d, ". Expected size ",
                     first.size(k), " but got size ", t.size(k), " for tensor number ", i, " in the list.");
        }
        total += t.size(d);
        rg = rg || t.requires_grad();
    }
    auto out_shape = first.sizes();
    out_shape[d] = total;
    Tensor out = empty(out_shape, first.dtype(), first.device());
    int64_t off = 0;
    for (const Tensor &t : tensors) {
        Tensor nt = out.narrow(d, off, t.size(d));
        run_copy(nt, t);
        off += t.size(d);
    }
    if (rg) {
        auto *fn = new CatGrad();
        fn->dim = d;
        for (auto &t : tensors) {
            fn->sizes.push_back(t.size(d));
            fn->inputs.push_back(t);
        }
        out.impl->requires_grad = true;
        out.grad_fn = Ref<GradFunction>(fn);
    }
    return out;
}

struct SliceGrad : GradFunction {
    std::vector<int64_t> in_shape;
    int64_t dim, start, end, step;
    const char *name() const override { return "SliceGrad"; }
    std::vector<Tensor> backward(const Tensor &g) override {
        Tensor gi = zeros(in_shape, g.dtype(), g.device());
        Tensor dst = gi.slice(dim, start, end, step);
        run_copy(dst, g);
        return {gi};
    }
};
Tensor slice(const Tensor &self, int64_t dim, int64_t start, int64_t end, int64_t step) {
    Tensor out = self.slice(dim, start, end, step);
    if (self.requires_grad()) {
        auto *fn = new SliceGrad();
        fn->in_shape = self.sizes();
        fn->dim = dim; fn->start = start; fn->end = end; fn->step = step;
        attach(out, fn, {self});
    }
    return out;
}

Tensor narrow(const Tensor &self, int64_t dim, int64_t start, int64_t length) {
    Tensor out = self.narrow(dim, start, length);  // bounds-checked view (ref: Tensor::narrow, tensor.cpp:233-244)
    if (self.requires_grad()) {
        const int d = wrap_dim(dim, self.dim());
        auto *fn = new SliceGrad();
        fn->in_shape = self.sizes();
        fn->dim = d; fn->start = start; fn->end = start + length; fn->step = 1;
        attach(out, fn, {self});
    }
    return out;
}

struct SelectGrad : GradFunction {  // x[i] along one dim: the gradient is scattered into a zero tensor of x's shape
    std::vector<int64_t> in_shape;
    int64_t dim, index;
    const char *name() const override { return "SelectGrad"; }
    std::vector<Tensor> backward(const Tensor &g) override {
        Tensor gi = zeros(in_shape, g.dtype(), g.device());
        Tensor dst = gi.select(dim, index);
        run_copy(dst, g);
        return {gi};
    }
};
Tensor select(const Tensor &self, int64_t dim, int64_t index) {
    Tensor out = self.select(dim, index);
    if (self.requires_grad()) {
        auto *fn = new SelectGrad();
        fn->in_shape = self.sizes();
        fn->dim = dim;
        fn->index = index;
        attach(out, fn, {self});
    }
    return out;
}

std::vector<Tensor> split(const Tensor &self, const std::vector<int64_t> &sizes, int64_t dim_) {
    KF_CHECK(self.dim() > 0, "tensor_split expected at least a 1-dimensional tensor, but got a tensor with ", self.dim(), " dims");
    const int d = wrap_dim(dim_, self.dim());
    std::vector<Tensor> out;
    int64_t start = 0;
    for (auto s : sizes) {
        out.push_back(slice(self, d, start, start + s, 1));
        start += s;
    }
    KF_CHECK(start == self.size(d), "split sizes must sum to the dimension size");
    return out;
}

struct PermuteGrad : GradFunction {
    std::vector<int64_t> inv;
    const char *name() const override { return "PermuteGrad"; }
    std::vector<Tensor> backward(const Tensor &g) override { return {g.permute(inv)}; }
};
Tensor permute(const Tensor &self, const std::vector<int64_t> &dims) {
    Tensor out = self.permute(dims);
    if (self.requires_grad()) {
        auto *fn = new PermuteGrad();
        fn->inv.resize(dims.size());
        for (size_t i = 0; i < dims.size(); ++i) fn->inv[wrap_dim(dims[i], (int64_t)dims.size())] = (int64_t)i;
        attach(out, fn, {self});
    }
    return out;
}
struct ViewGrad : GradFunction {
    std::vector<int64_t> in_shape;
    const char *name() const override { return "ViewGrad"; }
    std::vector<Tensor> backward(const Tensor &g) override { return {g.contiguous().view(in_shape)}; }
};
Tensor view(const Tensor &self, const std::vector<int64_t> &sizes) {
    Tensor out = self.view(sizes);
    if (self.requires_grad()) {
        auto *fn = new ViewGrad();
        fn->in_shape = self.sizes();
        attach(out, fn, {self});
    }
    return out;
}

Tensor &index_put_(Tensor &self, const std::vector<Tensor> &indices, const Tensor &values) {
    KF_CHECK((int)indices.size() == self.dim(), "Number of indices must match the number of dimensions in the tensor.");
    KF_CHECK(self.defined() && values.defined(), "Both self and values tensors must be defined.");
    KF_CHECK(self.dtype() == values.dtype(), "Data types of self and values tensors must match.");
    require_device(self, "index_put_");
    const int64_t n = values.numel();
    std::vector<Tensor> keep;
    std::vector<const int64_t *> ptrs;
    for (auto &ix : indices) {
        KF_CHECK(ix.defined() && ix.dtype() == KF_LONG, "Indices must be of type Long.");
        KF_CHECK(ix.numel() == n, "indices and values must have the same number of elements");
        keep.push_back(ix.contiguous());
        ptrs.push_back(keep.back().data_as<int64_t>());
    }
    for (int d = 0; d < self.dim(); ++d) KF_CHECK(self.size(d) != 0, "index is out of bounds for dimension with size 0");
    Tensor v = values.contiguous();
    auto sizes = self.sizes(), strides = self.strides();
    launch_index_put(self.data(), self.dtype(), sizes.data(), strides.data(), self.dim(), 1, ptrs.data(), v.data(), n);
    return self;
}

// ---- embedding (ref: the gather half of IndexElementwiseKernel, src/device/utils/tensor_index.h:19-143,
// src/core/index_ops.cpp:6-38; README.md:30 lists `embedding` as the reference's next op)
struct EmbeddingGrad : GradFunction {
    Tensor idx;  // int64, contiguous
    int64_t V, E;
    const char *name() const override { return "EmbeddingGrad"; }
    std::vector<Tensor> backward(const Tensor &g0) override {
        // deterministic scatter-add: stable-sort the token ids (ties keep ascending position), then every run of equal ids is
        // summed in position order by ONE warp in fp32 — no atomics, bit-reproducible
        Tensor g = g0.is_contiguous() ? g0 : clone(g0.detach());
        const int64_t n = idx.numel();
        Tensor dw = zeros({V, E}, g.dtype(), g.device());
        if (n == 0) return {dw, Tensor()};
        Tensor sv = empty({1, n}, KF_LONG, g.device()), si = empty({1, n}, KF_LONG, g.device());
        launch_sort_rows(idx.data(), sv.data(), si.data_as<int64_t>(), KF_LONG, 1, n, false);
        launch_embedding_bwd(g.data(), sv.data_as<int64_t>(), si.data_as<int64_t>(), dw.data(), g.dtype(), n, V, E);
        return {dw, Tensor()};
    }
};

Tensor embedding(const Tensor &weight, const Tensor &indices) {
    require_device(weight, "embedding");
    require_device(indices, "embedding");
    KF_CHECK(weight.dim() == 2, "embedding: weight must be [V, E]");
    KF_CHECK(indices.dtype() == KF_LONG, "Indices must be of type Long.");
    Tensor w = weight.is_contiguous() ? weight : contiguous(weight);
    Tensor idx = indices.detach().contiguous();
    auto oshape = idx.sizes();
    oshape.push_back(w.size(1));
    Tensor out = empty(oshape, w.dtype(), w.device());
    if (out.numel() > 0) launch_embedding_fwd(w.data(), idx.data_as<int64_t>(), out.data(), w.dtype(), idx.numel(), w.size(0), w.size(1));
    if (weight.requires_grad()) {
        auto *fn = new EmbeddingGrad();
        fn->idx = idx;
        fn->V = w.size(0);
        fn->E = w.size(1);
        attach(out, fn, {weight});
    }
    return out;
}

Tensor &random_uniform_(Tensor &self, uint64_t seed, double lo, double hi) {
    require_device(self, "random_uniform_");
    KF_CHECK(self.is_contiguous(), "random_uniform_: contiguous tensors only");
    KF_CHECK(is_floating(self.dtype()), "random_uniform_: floating dtypes only");
    if (self.numel() > 0) launch_fill_random(self.data(), self.dtype(), self.numel(), seed, (float)lo, (float)hi);
    return self;
}

// ================================================================== GEMM
static void fill_gemm_operand(const Tensor &t, bool trans, int64_t &rows, int64_t &cols, int64_t &ld, int64_t &bstride, int64_t &batch,
                              Tensor &holder) {
    // accept [.., r, c] with unit stride on the last dim and a uniform batch stride; anything else is materialised
    holder = t;
    auto ok = [&](const Tensor &x) {
        if (x.dim() < 2 || x.stride(-1) != 1) return false;
        if (x.size(-2) > 1 && x.stride(-2) < x.size(-1)) return false;
        int64_t expect = -1;
        for (int i = x.dim() - 3; i >= 0; --i) {  // leading dims must form one dense batch index
            if (x.size(i) == 1) continue;
            if (expect < 0) expect = x.stride(i) * x.size(i);
            else {
                if (x.stride(i) != expect) return false;
                expect = x.stride(i) * x.size(i);
            }
        }
        return true;
    };
    if (!ok(holder)) holder = clone(holder.detach());
    const Tensor &x = holder;
    batch = 1;
    for (int i = 0; i < x.dim() - 2; ++i) batch *= x.size(i);
    bstride = 0;
    for (int i = x.dim() - 3; i >= 0; --i)
        if (x.size(i) != 1) { bstride = x.stride(i); break; }
    if (batch == 1) bstride = 0;
    ld = x.size(-2) > 1 ? x.stride(-2) : x.size(-1);
    rows = trans ? x.size(-1) : x.size(-2);
    cols = trans ? x.size(-2) : x.size(-1);
}

struct MatmulGrad : GradFunction {
    Tensor a, b;
    bool ta, tb;
    float alpha;
    bool b_shared;  // b was 2-D against a batched a (gemm semantics)
    const char *name() const override { return "MatmulGrad"; }
    std::vector<Tensor> backward(const Tensor &g) override;
};

static thread_local int64_t g_route_M = 0;  // set by gemm_host around its slab products (GemmPlan::route_M)
static Tensor matmul_nograd(const Tensor &a, bool ta, const Tensor &b, bool tb, float alpha, float beta, Tensor *out_opt,
                            const Tensor *residual = nullptr, bool residual_is_row = false) {
    require_device(a, "matmul");
    require_device(b, "matmul");
    KF_CHECK(a.dtype() == b.dtype(), "matmul: dtype mismatch");
    KF_CHECK(a.dim() >= 2 && b.dim() >= 2, "matmul needs >= 2-d operands");
    Tensor ha, hb;
    GemmPlan p{};
    int64_t Ka, Kb, batch_a, batch_b;
    fill_gemm_operand(a, ta, p.M, Ka, p.lda, p.sa, batch_a, ha);
    fill_gemm_operand(b, tb, Kb, p.N, p.ldb, p.sb, batch_b, hb);
    KF_CHECK(Ka == Kb, "matmul: inner dimensions differ (", Ka, " vs ", Kb, ")");
    p.K = Ka;
    p.route_M = g_route_M;
    std::vector<int64_t> out_shape;
    if (b.dim() == 2 && !ta && a.dim() > 2 && ha.is_contiguous()) {
        // gemm semantics: fold every leading dim of a into M (ref: gemm_kernel.cu:10-15)
        p.M *= batch_a;
        batch_a = 1;
        p.sa = 0;
        out_shape = a.sizes();
        out_shape.back() = p.N;
    } else {
        KF_CHECK(batch_a == batch_b || batch_a == 1 || batch_b == 1, "matmul: batch dims differ");
        const Tensor &lead = batch_a >= batch_b ? a : b;
        for (int i = 0; i < lead.dim() - 2; ++i) out_shape.push_back(lead.size(i));
        out_shape.push_back(p.M);
        out_shape.push_back(p.N);
    }
    p.batch = std::max(batch_a, batch_b);
    if (batch_a == 1) p.sa = 0;
    if (batch_b == 1) p.sb = 0;
    Tensor out;
    if (out_opt) {
        out = *out_opt;
        KF_CHECK(out.is_contiguous() && out.dtype() == a.dtype());
        KF_CHECK(out.numel() == p.batch * p.M * p.N, "gemm_out: output has the wrong number of elements");
    } else {
        out = empty(out_shape, a.dtype(), a.device());
    }
    p.a = ha.data();
    p.b = hb.data();
    p.c = out.data();
    p.dtype = a.dtype();
    p.ldc = p.N;
    p.sc = p.M * p.N;
    p.trans_a = ta;
    p.trans_b = tb;
    p.alpha = alpha;
    p.beta = beta;
    Tensor rh;
    if (residual) {
        KF_CHECK(residual->dtype() == a.dtype() && residual->numel() == (residual_is_row ? p.N : out.numel()),
                 "gemm_residual: residual must match the output (or one row of it)");
        rh = residual->is_contiguous() ? *residual : clone(residual->detach());
        p.residual = rh.data();
        p.ldr = residual_is_row ? 0 : p.N;  // a bias row: every output row adds the same N values (row and batch strides 0)
        p.sr = residual_is_row ? 0 : p.M * p.N;
    }
    if (out.numel() > 0) launch_gemm(p);
    return out;
}

std::vector<Tensor> MatmulGrad::backward(const Tensor &g0) {
    std::vector<Tensor> r(2);
    Tensor g = g0.contiguous();
    const bool na = inputs[0].requires_grad(), nb = inputs[1].requires_grad();
    Tensor g2 = g, a2 = a;
    if (b_shared) {  // fold the leading dims like the forward did
        g2 = g.view({-1, g.size(-1)});
        a2 = a.contiguous().view({-1, a.size(-1)});
    }
    // an operand whose batch was broadcast (batch 1 against a batched partner) gets the SUM of the per-batch gradients
    auto fold_batch = [](const Tensor &grad, const Tensor &operand) {
        if (grad.numel() == operand.numel()) return grad.sizes() == operand.sizes() ? grad : grad.contiguous().view(operand.sizes());
        Tensor g3 = grad.contiguous().view({-1, grad.size(-2), grad.size(-1)});
        KF_CHECK(g3.size(1) * g3.size(2) == operand.numel(), "MatmulGrad: gradient shape does not match a broadcast operand");
        return sum(g3, 0).view(operand.sizes());
    };
    // dW (the operand the data-parallel all-reduce is waiting for) is issued BEFORE dX so that its collective can start one
    // GEMM earlier under the overlapped schedule (kf_set_leaf_grad_hook)
    if (nb) {
        Tensor db = tb ? matmul_nograd(g2, true, a2, ta, alpha, 0.f, nullptr) : matmul_nograd(a2, !ta, g2, false, alpha, 0.f, nullptr);
        r[1] = fold_batch(db, b);
    }
    if (na) {
        // C = op(A) op(B): dA = dC op(B)^T (or its transpose when A was transposed)
        Tensor da = ta ? matmul_nograd(b, tb, g2, true, alpha, 0.f, nullptr) : matmul_nograd(g2, false, b, !tb, alpha, 0.f, nullptr);
        r[0] = b_shared ? da.view(a.sizes()) : fold_batch(da, a);
    }
    return r;
}

Tensor matmul(const Tensor &a, bool ta, const Tensor &b, bool tb, float alpha) {
    Tensor out = matmul_nograd(a, ta, b, tb, alpha, 0.f, nullptr);
    if (any_requires_grad({&a, &b})) {
        auto *fn = new MatmulGrad();
        fn->a = a.detach();
        fn->b = b.detach();
        fn->ta = ta;
        fn->tb = tb;
        fn->alpha = alpha;
        fn->b_shared = (b.dim() == 2 && !ta && a.dim() > 2);
        attach(out, fn, {a, b});
    }
    return out;
}

// ---- fused GEMM epilogues (the reference lists the fused linear ops as its next ones, README.md:32)
struct MatmulResidualGrad : MatmulGrad {  // inputs: a, b, residual
    const char *name() const override { return "MatmulResidualGrad"; }
    std::vector<Tensor> backward(const Tensor &g) override {
        std::vector<Tensor> r = MatmulGrad::backward(g);
        r.push_back(g);
        return r;
    }
};

// out = alpha * a @ b + residual in ONE kernel (the residual add runs in the GEMM epilogue); bit-identical to gemm() followed by +
Tensor gemm_residual(const Tensor &a, const Tensor &b, const Tensor &residual, float alpha) {
    KF_CHECK(a.is_contiguous() && b.is_contiguous(), "gemm: operands must be contiguous");
    KF_CHECK(b.dim() == 2 && a.dim() >= 2 && b.size(0) == a.size(-1), "gemm: b must be [K, N]");
    KF_CHECK(a.dtype() == b.dtype() && a.dtype() == residual.dtype(), "gemm: dtype mismatch");
    auto oshape = a.sizes();
    oshape.back() = b.size(1);
    KF_CHECK(residual.sizes() == oshape, "gemm_residual: residual must have the shape of the result");
    Tensor out = matmul_nograd(a, false, b, false, alpha, 0.f, nullptr, &residual);
    if (any_requires_grad({&a, &b, &residual})) {
        auto *fn = new MatmulResidualGrad();
        fn->a = a.detach();
        fn->b = b.detach();
        fn->ta = false;
        fn->tb = false;
        fn->alpha = alpha;
        fn->b_shared = a.dim() > 2;
        attach(out, fn, {a, b, residual});
    }
    return out;
}

struct LinearBiasGrad : MatmulGrad {  // inputs: x, w, bias
    std::vector<int64_t> bias_shape;
    const char *name() const override { return "LinearBiasGrad"; }
    std::vector<Tensor> backward(const Tensor &g) override {
        std::vector<Tensor> r = MatmulGrad::backward(g);
        Tensor db;
        if (inputs[2].requires_grad()) {
            Tensor g2 = g.contiguous().view({-1, g.size(-1)});
            db = sum(g2, 0).view(bias_shape);  // column sums over every row of every batch entry (fp32 accumulation)
        }
        r.push_back(db);
        return r;
    }
};

// y[..., N] = x[..., K] @ w[K, N] (+ bias[N]): the projection the reference's README names as its next op (README.md:32 `qkv_linear`,
// on top of gpu::gemm, src/core/gemm_ops.cpp:6-16).  The bias row is added in the GEMM epilogue (residual pointer with row stride
// 0): bit-identical to gemm followed by a broadcast add.  `bias` may be undefined (plain projection).
Tensor qkv_linear(const Tensor &x, const Tensor &w, const Tensor &bias) {
    if (!bias.defined()) return matmul(x, false, w, false, 1.f);
    KF_CHECK(x.is_contiguous() && w.is_contiguous(), "qkv_linear: operands must be contiguous");
    KF_CHECK(w.dim() == 2 && x.dim() >= 2 && w.size(0) == x.size(-1), "qkv_linear: w must be [K, N]");
    KF_CHECK(x.dtype() == w.dtype() && x.dtype() == bias.dtype(), "qkv_linear: dtype mismatch");
    KF_CHECK(bias.numel() == w.size(1), "qkv_linear: bias must have N = ", w.size(1), " elements");
    Tensor out = matmul_nograd(x, false, w, false, 1.f, 0.f, nullptr, &bias, true);
    if (any_requires_grad({&x, &w, &bias})) {
        auto *fn = new LinearBiasGrad();
        fn->a = x.detach();
        fn->b = w.detach();
        fn->ta = false;
        fn->tb = false;
        fn->alpha = 1.f;
        fn->b_shared = x.dim() > 2;
        fn->bias_shape = bias.sizes();
        attach(out, fn, {x, w, bias});
    }
    return out;
}

struct GluGrad : GradFunction {  // h = (a b1) o (a b3); inputs: a, b1, b3
    Tensor a, b1, b3, u, v;
    const char *name() const override { return "GluGrad"; }
    std::vector<Tensor> backward(const Tensor &g0) override {
        Tensor g = g0.contiguous();
        Tensor du = binary(EW_MUL, g, v), dv = binary(EW_MUL, g, u);
        Tensor a2 = a.view({-1, a.size(-1)}), du2 = du.view({-1, du.size(-1)}), dv2 = dv.view({-1, dv.size(-1)});
        std::vector<Tensor> r(3);
        if (inputs[1].requires_grad()) r[1] = matmul_nograd(a2, true, du2, false, 1.f, 0.f, nullptr);
        if (inputs[2].requires_grad()) r[2] = matmul_nograd(a2, true, dv2, false, 1.f, 0.f, nullptr);
        if (inputs[0].requires_grad()) {
            // da = du b1^T + dv b3^T: the second product accumulates into the first through beta = 1 (no separate add)
            Tensor da = matmul_nograd(du2, false, b1, true, 1.f, 0.f, nullptr);
            matmul_nograd(dv2, false, b3, true, 1.f, 1.f, &da);
            r[0] = da.view(a.sizes());
        }
        return r;
    }
};

// h = gemm(a, b1) * gemm(a, b3) — the bilinear GLU of the transformer block — as one dual-B tcgen05 kernel
Tensor gemm_glu(const Tensor &a, const Tensor &b1, const Tensor &b3) {
    require_device(a, "gemm_glu");
    KF_CHECK(a.is_contiguous() && b1.is_contiguous() && b3.is_contiguous(), "gemm: operands must be contiguous");
    KF_CHECK(b1.dim() == 2 && a.dim() >= 2 && b1.size(0) == a.size(-1) && b3.sizes() == b1.sizes(), "gemm_glu: b1, b3 must be [K, N]");
    KF_CHECK(a.dtype() == b1.dtype() && a.dtype() == b3.dtype(), "gemm: dtype mismatch");
    const bool rg = any_requires_grad({&a, &b1, &b3});
    auto oshape = a.sizes();
    oshape.back() = b1.size(1);
    Tensor out = empty(oshape, a.dtype(), a.device());
    Tensor u, v;
    if (rg) {
        u = empty(oshape, a.dtype(), a.device());
        v = empty(oshape, a.dtype(), a.device());
    }
    GemmPlan p{};
    p.a = a.data(); p.b = b1.data(); p.b2 = b3.data(); p.c = out.data();
    p.dtype = a.dtype();
    p.K = a.size(-1); p.N = b1.size(1); p.M = p.K ? a.numel() / p.K : 0; p.batch = 1;
    p.lda = p.K; p.ldb = p.N; p.ldc = p.N;
    p.sc = p.M * p.N;
    p.alpha = 1.f; p.beta = 0.f;
    p.glu_u = rg ? u.data() : nullptr;
    p.glu_v = rg ? v.data() : nullptr;
    if (out.numel() == 0 || !launch_gemm_glu_tc(p)) {  // shapes / dtypes the dual-B kernel does not take: three launches
        Tensor uu = matmul_nograd(a, false, b1, false, 1.f, 0.f, nullptr), vv = matmul_nograd(a, false, b3, false, 1.f, 0.f, nullptr);
        u = uu; v = vv;
        run_binary(EW_MUL, out, uu, vv);
    }
    if (rg) {
        auto *fn = new GluGrad();
        fn->a = a.detach(); fn->b1 = b1.detach(); fn->b3 = b3.detach();
        fn->u = u; fn->v = v;
        attach(out, fn, {a, b1, b3});
    }
    return out;
}

Tensor gemm(const Tensor &a, const Tensor &b, float alpha, float beta) {
    // ref: gpu::gemm (src/core/gemm_ops.cpp:10-16) + checks of gemm_kernel (src/device/gemm_kernel.cu:8-25).
    // The reference hands CUTLASS an uninitialised `out`, so beta != 0 reads garbage there; we define it as beta*0.
    KF_CHECK(a.is_contiguous() && b.is_contiguous(), "gemm: operands must be contiguous");
    KF_CHECK(b.dim() == 2 && b.size(0) == a.size(-1), "gemm: b must be [K, N]");
    KF_CHECK(a.dtype() == b.dtype(), "gemm: dtype mismatch");
    (void)beta;
    if (a.dim() == 1) {
        Tensor a2 = view(a, {1, a.size(0)});
        Tensor o = matmul(a2, false, b, false, alpha);
        return view(o, {b.size(1)});
    }
    return matmul(a, false, b, false, alpha);
}

// Host-resident GEMM: C_host[M,N] = alpha * A_host[M,K] @ B_host[K,N], operands and result in (pinned) host memory.
// The reference's user pays H2D(A) + H2D(B) + GEMM + D2H(C) back to back (register.cpp:27-57 around gemm); PCIe is full duplex,
// so here B goes up first on a copy stream, then A in M-slabs; each slab's product is computed on the library stream as soon as
// it has landed and comes down on a second copy stream while the next slab goes up.  Lower bound = H2D(A + B); the D2H and the
// GEMM hide under it except for the last slab.  Two device slabs of A and of C are recycled through events.
void gemm_host(const void *a_host, const void *b_host, void *c_host, int64_t M, int64_t N, int64_t K, DType dtype, float alpha,
               int64_t slab_rows) {
    KF_CHECK(a_host && b_host && c_host && M > 0 && N > 0 && K > 0, "gemm_host: bad arguments");
    KF_CHECK(dtype == KF_HALF || dtype == KF_BFLOAT16 || dtype == KF_FLOAT || dtype == KF_DOUBLE, "gemm_host: floating dtypes only");
    Runtime &rt = Runtime::get();
    static cudaStream_t up = nullptr, down = nullptr;
    static cudaEvent_t ev_b = nullptr, a_ready[2], a_free[2], c_ready[2], c_free[2];
    if (!up) {
        KF_CUDA(cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking));
        KF_CUDA(cudaStreamCreateWithFlags(&down, cudaStreamNonBlocking));
        KF_CUDA(cudaEventCreateWithFlags(&ev_b, cudaEventDisableTiming));
        for (int i = 0; i < 2; ++i) {
            KF_CUDA(cudaEventCreateWithFlags(&a_ready[i], cudaEventDisableTiming));
            KF_CUDA(cudaEventCreateWithFlags(&a_free[i], cudaEventDisableTiming));
            KF_CUDA(cudaEventCreateWithFlags(&c_ready[i], cudaEventDisableTiming));
            KF_CUDA(cudaEventCreateWithFlags(&c_free[i], cudaEventDisableTiming));
        }
    }
    const size_t es = element_size(dtype);
    if (slab_rows <= 0) slab_rows = 1024;
    slab_rows = std::min(slab_rows, M);
    const int dev = rt.device();
    Tensor dB = empty({K, N}, dtype, dev);
    Tensor dA[2] = {empty({slab_rows, K}, dtype, dev), empty({slab_rows, K}, dtype, dev)};
    Tensor dC[2] = {empty({slab_rows, N}, dtype, dev), empty({slab_rows, N}, dtype, dev)};
    // the slabs are touched by the two copy streams as well: the pool fences those streams when the slabs go back to it
    rt.pool().record_stream(dB.data(), up);
    for (int i = 0; i < 2; ++i) {
        rt.pool().record_stream(dA[i].data(), up);
        rt.pool().record_stream(dC[i].data(), down);
    }
    // the copy streams start after everything already queued on the library stream (the pool hands out memory in its order)
    KF_CUDA(cudaEventRecord(ev_b, rt.stream()));
    KF_CUDA(cudaStreamWaitEvent(up, ev_b, 0));
    KF_CUDA(cudaStreamWaitEvent(down, ev_b, 0));
    KF_CUDA(cudaMemcpyAsync(dB.data(), b_host, (size_t)K * N * es, cudaMemcpyHostToDevice, up));
    KF_CUDA(cudaEventRecord(ev_b, up));
    KF_CUDA(cudaStreamWaitEvent(rt.stream(), ev_b, 0));
    int s = 0;
    for (int64_t m0 = 0; m0 < M; m0 += slab_rows, ++s) {
        const int i = s & 1;
        const int64_t rows = std::min(slab_rows, M - m0);
        if (s >= 2) KF_CUDA(cudaStreamWaitEvent(up, a_free[i], 0));  // the product that read this A slab has run
        KF_CUDA(cudaMemcpyAsync(dA[i].data(), (const char *)a_host + (size_t)m0 * K * es, (size_t)rows * K * es, cudaMemcpyHostToDevice, up));
        KF_CUDA(cudaEventRecord(a_ready[i], up));
        KF_CUDA(cudaStreamWaitEvent(rt.stream(), a_ready[i], 0));
        if (s >= 2) KF_CUDA(cudaStreamWaitEvent(rt.stream(), c_free[i], 0));  // the download of this C slab's previous tenant is done
        Tensor av = rows == slab_rows ? dA[i] : dA[i].slice(0, 0, rows, 1);
        Tensor cv = rows == slab_rows ? dC[i] : dC[i].slice(0, 0, rows, 1);
        g_route_M = M;
        try {
            matmul_nograd(av, false, dB, false, alpha, 0.f, &cv);
        } catch (...) {
            g_route_M = 0;
            throw;
        }
        g_route_M = 0;
        KF_CUDA(cudaEventRecord(a_free[i], rt.stream()));
        KF_CUDA(cudaEventRecord(c_ready[i], rt.stream()));
        KF_CUDA(cudaStreamWaitEvent(down, c_ready[i], 0));
        KF_CUDA(cudaMemcpyAsync((char *)c_host + (size_t)m0 * N * es, dC[i].data(), (size_t)rows * N * es, cudaMemcpyDeviceToHost, down));
        KF_CUDA(cudaEventRecord(c_free[i], down));
    }
    // join: later work on the library stream (and the pool's stream-ordered frees of dA / dB / dC) follows both copy streams
    KF_CUDA(cudaEventRecord(ev_b, up));
    KF_CUDA(cudaStreamWaitEvent(rt.stream(), ev_b, 0));
    KF_CUDA(cudaStreamWaitEvent(rt.stream(), c_free[0], 0));
    if (s >= 2) KF_CUDA(cudaStreamWaitEvent(rt.stream(), c_free[1], 0));
}

void gemm_out(Tensor &out, const Tensor &a, const Tensor &b, float alpha, float beta) {
    KF_CHECK(out.is_contiguous() && a.is_contiguous() && b.is_contiguous());
    KF_CHECK(b.dim() == 2 && b.size(0) == a.size(-1));
    KF_CHECK(a.dtype() == b.dtype());
    KF_CHECK(out.size(-1) == b.size(-1));
    Tensor a2 = a.dim() == 1 ? a.view({1, a.size(0)}) : a;
    matmul_nograd(a2, false, b, false, alpha, beta, &out);
}

// ================================================================== attention
struct AttentionGrad : GradFunction {
    Tensor q, k, v, out, lse;
    const char *name() const override { return "AttentionGrad"; }
    std::vector<Tensor> backward(const Tensor &g) override {
        auto [dq, dk, dv] = causal_attention_bwd(g, q, k, v, out, lse);
        return {dq, dk, dv};
    }
};

static void check_attention(const Tensor &q, const Tensor &k, const Tensor &v) {
    require_device(q, "causal_attention");
    require_device(k, "causal_attention");
    require_device(v, "causal_attention");
    KF_CHECK(q.dim() == 4 && k.dim() == 4 && v.dim() == 4, "causal_attention expects [B, H, S, D] tensors");
    KF_CHECK(k.size(0) == q.size(0) && k.size(1) == q.size(1) && k.size(3) == q.size(3));
    KF_CHECK(k.sizes() == v.sizes());
    KF_CHECK(q.dtype() == k.dtype() && q.dtype() == v.dtype());
    KF_CHECK(is_floating(q.dtype()), "Unsupported ScalarType ", dtype_name(q.dtype()));
}

std::tuple<Tensor, Tensor> causal_attention_fwd(const Tensor &q_, const Tensor &k_, const Tensor &v_) {
    check_attention(q_, k_, v_);
    Tensor q = q_.detach().contiguous(), k = k_.detach().contiguous(), v = v_.detach().contiguous();
    Tensor out = empty(q.sizes(), q.dtype(), q.device());
    Tensor lse = empty({q.size(0), q.size(1), q.size(2)}, q.dtype() == KF_DOUBLE ? KF_DOUBLE : KF_FLOAT, q.device());
    AttnPlan p{};
    p.q = q.data(); p.k = k.data(); p.v = v.data(); p.out = out.data();
    p.lse = lse.data();
    p.dtype = q.dtype();
    p.BH = q.size(0) * q.size(1);
    p.Sq = q.size(2);
    p.Skv = k.size(2);
    p.D = q.size(3);
    if (out.numel() > 0) launch_attention_fwd(p);
    return {out, lse};
}

Tensor causal_attention(const Tensor &q, const Tensor &k, const Tensor &v) {
    auto [out, lse] = causal_attention_fwd(q, k, v);
    if (any_requires_grad({&q, &k, &v})) {
        auto *fn = new AttentionGrad();
        fn->q = q.detach().contiguous();
        fn->k = k.detach().contiguous();
        fn->v = v.detach().contiguous();
        fn->out = out.detach();
        fn->lse = lse;
        attach(out, fn, {q, k, v});
    }
    return out;
}

// ---- fused qkv attention (README.md:32 lists `qkv_linear` as the reference's next fused op): attention reads q, k, v
// IN PLACE from the packed projection qkv [B, S, 3, H, D] (4-D TMA tensor maps with the packed strides) and writes [B, S, H * D]
// directly; the backward writes dq, dk, dv straight into one [B, S, 3, H, D] gradient.  No split / view / permute / contiguous
// launches in either direction (the composed form costs 3 + 1 transposes forward and 4 transposes + 3 zero-fills + 2 adds backward).
struct QkvAttentionGrad : GradFunction {
    Tensor qkv, out, lse;
    int64_t H;
    const char *name() const override { return "QkvAttentionGrad"; }
    std::vector<Tensor> backward(const Tensor &g0) override {
        Tensor g = g0.is_contiguous() ? g0 : clone(g0.detach());
        const int64_t B = qkv.size(0), S = qkv.size(1), E = qkv.size(2) / 3, D = E / H;
        Tensor dqkv = empty(qkv.sizes(), qkv.dtype(), qkv.device());
        AttnBwdPlan p{};
        const size_t es = qkv.itemsize();
        const char *base = reinterpret_cast<const char *>(qkv.data());
        char *dbase = reinterpret_cast<char *>(dqkv.data());
        p.q = base; p.k = base + E * es; p.v = base + 2 * E * es;
        p.out = out.data(); p.dout = g.data(); p.lse = lse.data();
        p.dq = dbase; p.dk = dbase + E * es; p.dv = dbase + 2 * E * es;
        p.dtype = qkv.dtype();
        p.BH = B * H; p.Sq = S; p.Skv = S; p.D = D; p.H = H;
        const AttnLayout packed{S * 3 * E, D, 3 * E}, merged{S * E, D, E};
        p.lq = p.lk = p.lv = p.ldq = p.ldk = p.ldv = packed;
        p.lo = p.ldo = merged;
        KF_CHECK(launch_attention_bwd_tc(p), "qkv_attention backward: the fused kernel rejected a shape its forward accepted");
        return {dqkv};
    }
};

Tensor qkv_attention(const Tensor &qkv_, int64_t H) {
    require_device(qkv_, "qkv_attention");
    KF_CHECK(qkv_.dim() == 3 && H > 0 && qkv_.size(2) % (3 * H) == 0, "qkv_attention expects [B, S, 3 * H * D]");
    const int64_t B = qkv_.size(0), S = qkv_.size(1), E = qkv_.size(2) / 3, D = E / H;
    const bool fast = (qkv_.dtype() == KF_HALF || qkv_.dtype() == KF_BFLOAT16) && D == 128 && B * H < 65536 && B > 0 && S > 0;
    if (fast) {
        Tensor qkv = qkv_.is_contiguous() ? qkv_ : contiguous(qkv_);
        Tensor out = empty({B, S, E}, qkv.dtype(), qkv.device());
        Tensor lse = empty({B, H, S}, KF_FLOAT, qkv.device());
        AttnPlan p{};
        const size_t es = qkv.itemsize();
        const char *base = reinterpret_cast<const char *>(qkv.data());
        p.q = base; p.k = base + E * es; p.v = base + 2 * E * es;
        p.out = out.data(); p.lse = lse.data();
        p.dtype = qkv.dtype();
        p.BH = B * H; p.Sq = S; p.Skv = S; p.D = D; p.H = H;
        p.lq = p.lk = p.lv = AttnLayout{S * 3 * E, D, 3 * E};
        p.lo = AttnLayout{S * E, D, E};
        if (launch_attention_fwd_tc(p)) {
            if (qkv.requires_grad()) {
                auto *fn = new QkvAttentionGrad();
                fn->qkv = qkv.detach();
                fn->out = out.detach();
                fn->lse = lse;
                fn->H = H;
                attach(out, fn, {qkv});
            }
            return out;
        }
    }
    // composed form from the reference's own operator names (any dtype / head size); carries its own autograd chain
    auto parts = split(qkv_, {E, E, E}, -1);
    auto heads = [&](const Tensor &t) { return contiguous(permute(view(contiguous(t), {B, S, H, D}), {0, 2, 1, 3})); };
    Tensor o = causal_attention(heads(parts[0]), heads(parts[1]), heads(parts[2]));
    return view(contiguous(permute(o, {0, 2, 1, 3})), {B, S, E});
}

// Generic backward: the five GEMMs of attention backward issued through our own GEMM kernels on fp32 (fp64)
// temporaries, one batch entry at a time so the S_q x S_kv scratch stays bounded.  Any shape / dtype.
// (The reference has no attention backward at all)
static void attention_bwd_generic(const Tensor &dout, const Tensor &q, const Tensor &k, const Tensor &v, const Tensor &o,
                                  const Tensor &lse, Tensor &dq, Tensor &dk, Tensor &dv) {
    const DType ct = q.dtype() == KF_DOUBLE ? KF_DOUBLE : KF_FLOAT;
    const double scale = 1.0 / std::sqrt((double)q.size(3));
    const int64_t B = q.size(0), H = q.size(1), Sq = q.size(2), Skv = k.size(2);
    for (int64_t b = 0; b < B; ++b) {
        auto cvt = [&](const Tensor &t) { return convert(t.narrow(0, b, 1), ct); };
        Tensor qf = cvt(q), kf = cvt(k), vf = cvt(v), of = cvt(o), dof = cvt(dout);
        Tensor lb = convert(lse.narrow(0, b, 1), ct).contiguous();
        Tensor P = matmul(qf, false, kf, true, (float)scale);              // S = scale * Q K^T   [1,H,Sq,Skv]
        launch_attn_probs(P.data(), lb.data(), ct, H, Sq, Skv);            // P = exp(S - lse), causal
        Tensor dP = matmul(dof, false, vf, true, 1.f);                     // dP = dO V^T
        Tensor delta = sum(binary(EW_MUL, dof, of), 3);                    // [1,H,Sq,1]
        Tensor dS = binary_scalar(EW_MUL, binary(EW_MUL, P, binary(EW_SUB, dP, delta)), scale);
        Tensor dvb = matmul(P, true, dof, false, 1.f);                     // dV = P^T dO
        Tensor dkb = matmul(dS, true, qf, false, 1.f);                     // dK = dS^T Q
        Tensor dqb = matmul(dS, false, kf, false, 1.f);                    // dQ = dS K
        Tensor dq_dst = dq.narrow(0, b, 1), dk_dst = dk.narrow(0, b, 1), dv_dst = dv.narrow(0, b, 1);
        run_copy(dq_dst, dqb);
        run_copy(dk_dst, dkb);
        run_copy(dv_dst, dvb);
    }
}

std::tuple<Tensor, Tensor, Tensor> causal_attention_bwd(const Tensor &dout_, const Tensor &q_, const Tensor &k_, const Tensor &v_,
                                                        const Tensor &out_, const Tensor &lse_) {
    check_attention(q_, k_, v_);
    Tensor q = q_.detach().contiguous(), k = k_.detach().contiguous(), v = v_.detach().contiguous();
    Tensor o = out_.detach().contiguous(), dout = dout_.detach().contiguous(), lse = lse_.detach().contiguous();
    KF_CHECK(dout.sizes() == q.sizes() && o.sizes() == q.sizes() && dout.dtype() == q.dtype());
    KF_CHECK((lse.dtype() == KF_FLOAT || lse.dtype() == KF_DOUBLE) && lse.numel() == q.size(0) * q.size(1) * q.size(2));
    Tensor dq = empty(q.sizes(), q.dtype(), q.device());
    Tensor dk = empty(k.sizes(), k.dtype(), k.device());
    Tensor dv = empty(v.sizes(), v.dtype(), v.device());
    if (q.numel() == 0 || k.numel() == 0) return {dq, dk, dv};
    AttnBwdPlan p{};
    p.q = q.data(); p.k = k.data(); p.v = v.data(); p.out = o.data(); p.dout = dout.data();
    p.lse = lse.data();
    p.dq = dq.data(); p.dk = dk.data(); p.dv = dv.data();
    p.dtype = q.dtype();
    p.BH = q.size(0) * q.size(1);
    p.Sq = q.size(2);
    p.Skv = k.size(2);
    p.D = q.size(3);
    if (!(lse.dtype() == KF_FLOAT && launch_attention_bwd_tc(p))) attention_bwd_generic(dout, q, k, v, o, lse, dq, dk, dv);
    return {dq, dk, dv};
}

// ================================================================== autograd engine
// Same two-pass scheme as the reference (src/core/tensor.cpp:86-126): count consumers, then a ready
// queue; leaf grads accumulate across backward() calls (tensor.cpp:75-84).
static LeafGradHook g_leaf_hook = nullptr;
static void *g_leaf_hook_ctx = nullptr;
void set_leaf_grad_hook(LeafGradHook fn, void *ctx) {
    g_leaf_hook = fn;
    g_leaf_hook_ctx = ctx;
}

void backward(Tensor &root, const Tensor &grad_output) {
    KF_CHECK(root.defined() && grad_output.defined());
    std::unordered_map<TensorImpl *, int> needed;
    std::unordered_map<TensorImpl *, Tensor> grad_acc;
    std::unordered_map<TensorImpl *, bool> visited;
    std::queue<Tensor *> ready;
    ready.push(&root);
    while (!ready.empty()) {
        Tensor *t = ready.front();
        ready.pop();
        if (!t->grad_fn) continue;
        if (visited[t->impl.get()]) continue;  // expand each node once, count every edge
        visited[t->impl.get()] = true;
        for (auto &in : t->grad_fn->inputs) {
            if (!in.requires_grad()) continue;
            needed[in.impl.get()] += 1;
            ready.push(&in);
        }
    }
    grad_acc[root.impl.get()] = grad_output;
    ready.push(&root);
    while (!ready.empty()) {
        Tensor *t = ready.front();
        ready.pop();
        Tensor go = grad_acc[t->impl.get()];
        grad_acc.erase(t->impl.get());
        if (t->grad_fn) {
            auto gis = t->grad_fn->backward(go);
            auto &inputs = t->grad_fn->inputs;
            for (size_t i = 0; i < inputs.size(); ++i) {
                if (!inputs[i].requires_grad()) continue;
                KF_CHECK(i < gis.size() && gis[i].defined(), t->grad_fn->name(), " produced no gradient for input ", i);
                KF_CHECK(gis[i].sizes() == inputs[i].sizes(), t->grad_fn->name(), " produced a gradient of the wrong shape for input ", i);
                Tensor &acc = grad_acc[inputs[i].impl.get()];
                acc = acc.defined() ? binary(EW_ADD, acc.detach(), gis[i].detach()) : gis[i].detach();
                if (--needed[inputs[i].impl.get()] == 0) ready.push(&inputs[i]);
            }
        } else if (t->requires_grad()) {
            TensorImpl *impl = t->impl.get();
            if (impl->grad) {
                Tensor &gacc = *impl->grad;
                run_binary(EW_ADD, gacc, gacc, go);
            } else if (go.impl.use_count() == 1 && go.impl->storage.use_count() == 1 && !go.impl->storage->external && go.is_contiguous()) {
                // the incoming gradient is a fresh tensor nobody else can see (its producer's handles are gone): it becomes the leaf's
                // grad slot as it is — no copy (one read + one write of every parameter- / input-sized gradient per step)
                impl->grad.reset(new Tensor(go));
            } else {
                Tensor gcopy = empty(go.sizes(), go.dtype(), go.device());
                run_copy(gcopy, go);
                impl->grad.reset(new Tensor(gcopy));
            }
            if (g_leaf_hook) g_leaf_hook(*t, *impl->grad, g_leaf_hook_ctx);
        }
    }
}

}  // namespace ops
}  // namespace kf

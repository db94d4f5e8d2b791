// Operator front-ends (host): build a plan, call one kernel launcher, wire autograd.
// Mirrors the reference's `gpu::` namespace (src/core/*_ops.cpp) function for function.
#pragma once
#include <tuple>

#include "core.h"
#include "kernels.h"

namespace kf {
namespace ops {

// ---- planning (ref: TensorIterator::build, src/core/tensor_iterator.cpp:486-515)
// Fills the collapsed plan for out = f(a, b).  `out` undefined => allocated (contiguous, promoted dtype).
void plan_elementwise(EwPlan &plan, Tensor &out, const Tensor *a, const Tensor *b, bool allow_resize_check);

// ---- elementwise
Tensor binary(int op, const Tensor &a, const Tensor &b);
Tensor &binary_(int op, Tensor &self, const Tensor &other);
Tensor binary_scalar(int op, const Tensor &a, double scalar);
Tensor &binary_scalar_(int op, Tensor &self, double scalar);
Tensor &fill_(Tensor &self, double value);
Tensor &copy_(Tensor &self, const Tensor &src);
Tensor clone(const Tensor &self);
Tensor convert(const Tensor &self, DType dtype);
Tensor unary(int op, const Tensor &a);
inline Tensor add(const Tensor &a, const Tensor &b) { return binary(EW_ADD, a, b); }
inline Tensor sub(const Tensor &a, const Tensor &b) { return binary(EW_SUB, a, b); }
inline Tensor mul(const Tensor &a, const Tensor &b) { return binary(EW_MUL, a, b); }
inline Tensor div(const Tensor &a, const Tensor &b) { return binary(EW_DIV, a, b); }

// ---- reductions (keepdim)
Tensor sum(const Tensor &self, int64_t dim);
Tensor mean(const Tensor &self, int64_t dim);
std::tuple<Tensor, Tensor> mean_var(const Tensor &self, int64_t dim, bool take_sqrt);
std::tuple<Tensor, Tensor> norm_stat(const Tensor &self, int64_t dim);
// reduce `grad` (shape = broadcast result) back to `shape` by summing broadcast dims (autograd helper)
Tensor sum_to_shape(const Tensor &grad, const std::vector<int64_t> &shape);

// ---- fused layer norm over the last dim (gain has E elements); differentiable in x and gain
Tensor layer_norm(const Tensor &x, const Tensor &gain, double eps);
Tensor rms_norm(const Tensor &x, const Tensor &gain, double eps);

// ---- sort / top-k
std::tuple<Tensor, Tensor> sort(const Tensor &self, int64_t dim, bool descending);
std::tuple<Tensor, Tensor> topk(const Tensor &self, int64_t k, int64_t dim, bool largest);

// ---- shape ops
Tensor cat(const std::vector<Tensor> &tensors, int64_t dim);
std::vector<Tensor> split(const Tensor &self, const std::vector<int64_t> &sizes, int64_t dim);
Tensor &index_put_(Tensor &self, const std::vector<Tensor> &indices, const Tensor &values);
// out[..., :] = weight[indices[...], :]; differentiable in weight (deterministic scatter-add)
Tensor embedding(const Tensor &weight, const Tensor &indices);
// counter-based uniform fill, element i = lo + (hi - lo) * u(i, seed): reproducible on the host (oracle.counter_uniform), used to
// build the BASELINE-size inputs (C4: 8.6 GB) on the device instead of uploading them
Tensor &random_uniform_(Tensor &self, uint64_t seed, double lo, double hi);
// differentiable view wrappers (the raw view algebra on Tensor carries no grad_fn)
Tensor permute(const Tensor &self, const std::vector<int64_t> &dims);
Tensor view(const Tensor &self, const std::vector<int64_t> &sizes);
Tensor contiguous(const Tensor &self);
Tensor slice(const Tensor &self, int64_t dim, int64_t start, int64_t end, int64_t step);
Tensor narrow(const Tensor &self, int64_t dim, int64_t start, int64_t length);
Tensor select(const Tensor &self, int64_t dim, int64_t index);

// ---- contractions
Tensor gemm(const Tensor &a, const Tensor &b, float alpha, float beta);
void gemm_out(Tensor &out, const Tensor &a, const Tensor &b, float alpha, float beta);
Tensor matmul(const Tensor &a, bool trans_a, const Tensor &b, bool trans_b, float alpha);
// fused epilogues: alpha * a @ b + residual, and the GLU product gemm(a, b1) * gemm(a, b3); bit-identical to the composed forms
Tensor gemm_residual(const Tensor &a, const Tensor &b, const Tensor &residual, float alpha);
Tensor qkv_linear(const Tensor &x, const Tensor &w, const Tensor &bias);  // bias may be undefined
Tensor gemm_glu(const Tensor &a, const Tensor &b1, const Tensor &b3);
// operands and result in (pinned) host memory; uploads, slab products and downloads overlap on three streams
void gemm_host(const void *a_host, const void *b_host, void *c_host, int64_t M, int64_t N, int64_t K, DType dtype, float alpha,
               int64_t slab_rows);
Tensor causal_attention(const Tensor &q, const Tensor &k, const Tensor &v);
// attention over a packed projection: qkv [B, S, 3 * H * D] -> [B, S, H * D], no head transposes in either direction
Tensor qkv_attention(const Tensor &qkv, int64_t H);
std::tuple<Tensor, Tensor> causal_attention_fwd(const Tensor &q, const Tensor &k, const Tensor &v);
std::tuple<Tensor, Tensor, Tensor> causal_attention_bwd(const Tensor &dout, const Tensor &q, const Tensor &k, const Tensor &v,
                                                        const Tensor &out, const Tensor &lse);

// ---- autograd (ref: Tensor::backward, src/core/tensor.cpp:86-126)
void backward(Tensor &root, const Tensor &grad_output);
// Called (on the calling thread, inside backward()) right after a leaf's gradient for this pass has been enqueued: the hook may
// order other streams behind the library stream and start the data-parallel all-reduce of that gradient while the rest of the
// backward pass is still being issued.  Not in the reference (it has no multi-GPU path); nullptr disables it.
using LeafGradHook = void (*)(const Tensor &leaf, const Tensor &grad, void *ctx);
void set_leaf_grad_hook(LeafGradHook fn, void *ctx);

}  // namespace ops

// data-parallel layer (dist.cpp): NCCL on a library-owned communication stream
namespace dist {
void unique_id(void *out128);
void init(const void *id128, int rank, int world);
void finalize();
bool initialised();
int rank();
int world();
int nccl_version();
void all_reduce(Tensor &t, int op);  // 0 sum, 1 avg, 2 max; in place, on the library stream
void overlap_begin(const std::vector<Tensor> &params);
int64_t overlap_end();
}  // namespace dist
}  // namespace kf

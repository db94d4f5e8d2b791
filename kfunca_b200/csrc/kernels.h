// Plain-C++ launch entry points of the sm_100a kernels (defined in kernels/*.cu).  Everything is
// issued on Runtime::stream().  These replace the reference's "kernel entry" headers
// src/device/include/*.h (SURVEY §8b inner boundary).
#pragma once
#include <cstdint>

#include "core.h"

namespace kf {

constexpr int kPlanDims = KF_MAX_DIMS;

enum EwOp : int {
    EW_ADD = 0, EW_SUB = 1, EW_MUL = 2, EW_DIV = 3,  // binary   (ref: binary_ops_kernel.cu:6-60)
    EW_COPY = 4,                                     // unary    (ref: unary_ops_kernel.cu:6-17)
    EW_FILL = 5,                                     // nullary  (ref: nullary_ops_kernel.cu:6-25)
    EW_SQRT = 6, EW_RSQRT = 7, EW_NEG = 8,           // unary maths for the block's norm
};

// Collapsed elementwise problem; dim 0 is the FASTEST varying dimension.  Strides in bytes.
struct EwPlan {
    int ndim;
    int nin;             // number of tensor inputs (0, 1 or 2)
    int op;
    int acc;             // AccKind the op computes in
    int64_t numel;
    int64_t shape[kPlanDims];
    int64_t stride[3][kPlanDims];  // [0] = out, [1] = a, [2] = b
    void *ptr[3];
    int dtype[3];
    int b_is_scalar;     // b operand is `scalar` (already rounded through the tensor dtype)
    double scalar;       // also the fill value
};
void launch_elementwise(const EwPlan &plan);

// Batched tiled transpose-copy (same dtype): out is dense in the collapsed order, `in` has unit stride
// along collapsed dim `tdim` > 0.  Both sides are moved in full 128 B lines through shared memory.
struct TransposePlan {
    int ndim;
    int tdim;
    int itemsize;
    int64_t shape[kPlanDims];
    int64_t in_stride[kPlanDims];   // elements
    int64_t out_stride[kPlanDims];  // elements
    const void *in;
    void *out;
};
void launch_transpose(const TransposePlan &plan);

// Single-axis reduction over a dense [outer, R, inner] input; out is [outer, inner] (keepdim handled by caller).
// mean: result = sum * factor (factor computed by the caller in the reference's arithmetic).
struct ReducePlan {
    const void *in;
    void *out;
    int dtype;
    int64_t outer, R, inner;
    int is_mean;
    double factor;
};
void launch_reduce(const ReducePlan &plan);

// Segmented stable sort of `nseg` dense rows of length n: values + int64 indices.
void launch_sort_rows(const void *in, void *values, int64_t *indices, int dtype, int64_t nseg, int64_t n, bool descending);
// Row-wise top-k (sorted, ties -> lowest index first).  Returns false when the fast select path does not
// apply (caller then falls back to sort + narrow exactly like the reference, sort_ops_kernel.cu:617-632).
bool launch_topk_rows(const void *in, void *values, int64_t *indices, int dtype, int64_t nseg, int64_t n, int64_t k, bool largest);

// index_put_ (ref: tensor_index.h:19-143): self[idx0[i], idx1[i], ...] = values[i]
void launch_index_put(void *self, int dtype, const int64_t *self_shape, const int64_t *self_stride, int nidx, int64_t inner,
                      const int64_t *const *idx_ptrs, const void *values, int64_t n);

// embedding gather / deterministic scatter-add (kernels/misc.cu)
void launch_embedding_fwd(const void *weight, const int64_t *idx, void *out, int dtype, int64_t n, int64_t V, int64_t E);
void launch_embedding_bwd(const void *grad, const int64_t *sorted_idx, const int64_t *sorted_pos, void *dweight, int dtype, int64_t n, int64_t V,
                          int64_t E);
// element i = lo + (hi - lo) * ((mix64(i, seed) >> 40) * 2^-24), fp32 arithmetic with separate multiply and add, then cast
void launch_fill_random(void *p, int dtype, int64_t n, uint64_t seed, float lo, float hi);

// GEMM: C[b][M,N] = alpha * op(A)[b][M,K] @ op(B)[b][K,N] + beta * C.  Row-major storage with leading
// dimensions in elements; trans flag = operand is stored as [K,M] / [N,K].  batch strides in elements (0 = broadcast).
struct GemmPlan {
    const void *a, *b;
    void *c;
    int dtype;
    int64_t M, N, K, batch;
    int64_t lda, ldb, ldc;
    int64_t sa, sb, sc;
    int trans_a, trans_b;
    float alpha, beta;
    // fused epilogues: C = alpha * op(A) op(B) + beta * C + residual  (residual: C's dtype, leading dim ldr)
    const void *residual = nullptr;
    int64_t ldr = 0, sr = 0;
    // GLU: b2 = second B operand (same layout / ldb as b); C = (A b) o (A b2); glu_u / glu_v (C's layout) receive the two factors
    const void *b2 = nullptr;
    void *glu_u = nullptr, *glu_v = nullptr;
    // rows of the WHOLE product when this plan is one row slab of it (gemm_host): kernel selection looks at the whole problem so
    // that every slab runs the kernel the one-piece product would (bit-identical results); 0 = this plan is the whole problem
    int64_t route_M = 0;
};
bool launch_gemm_glu_tc(const GemmPlan &p);  // false => compose from two GEMMs and a multiply
void launch_gemm_simt(const GemmPlan &p);  // fp32 / fp64 FFMA/DFMA path (strict-parity path)
bool launch_gemm_tc(const GemmPlan &p);    // fp16 / bf16 tcgen05 + TMEM + TMA path; false if shape unsupported
bool launch_gemm_f32_tc(const GemmPlan &p);  // fp32 on tcgen05 through three bf16 planes per operand (6 / 9 products); false => SIMT
void launch_gemm(const GemmPlan &p);       // dispatcher

// Element (b, h, s, d) of a [B, H, S, D] operand lives at base + b * sb + h * sh + s * ss + d (elements; unit stride along D).
// Dense [B,H,S,D]: {H*S*D, S*D, D}.  A packed projection [B, S, 3, H, D] gives {S*3*H*D, D, 3*H*D} with base offsets 0 / E / 2E.
struct AttnLayout {
    int64_t sb, sh, ss;
};
inline AttnLayout dense_bhsd(int64_t H, int64_t S, int64_t D) { return {H * S * D, S * D, D}; }
inline bool is_dense_bhsd(const AttnLayout &l, int64_t H, int64_t S, int64_t D) { return l.sb == H * S * D && l.sh == S * D && l.ss == D; }

// Causal attention (top-left aligned mask, scale 1/sqrt(D)); q [B,H,Sq,D], k/v [B,H,Skv,D] in the layouts lq / lk / lv.
struct AttnPlan {
    const void *q, *k, *v;
    void *out;
    void *lse;  // [BH, Sq] natural-log row log-sum-exp, fp32 (fp64 for fp64 inputs); may be null
    int dtype;
    int64_t BH, Sq, Skv, D;
    int64_t H = 0;            // heads per batch entry (BH = B * H); 0 => dense layouts, H irrelevant
    AttnLayout lq{}, lk{}, lv{}, lo{};
};
void launch_attention_fwd(const AttnPlan &p);     // dispatcher: tcgen05 path for 16-bit D in {64,128}, else SIMT
bool launch_attention_fwd_tc(const AttnPlan &p);  // false when the shape / dtype is not supported
bool launch_attention_fwd_f32_tc(const AttnPlan &p);  // fp32 on tcgen05 (three bf16 planes per operand); false => FFMA kernel
struct AttnBwdPlan {
    const void *q, *k, *v, *out, *dout;
    const void *lse;
    void *dq, *dk, *dv;
    int dtype;
    int64_t BH, Sq, Skv, D;
    int64_t H = 0;  // 0 => dense layouts
    AttnLayout lq{}, lk{}, lv{}, lo{}, ldo{}, ldq{}, ldk{}, ldv{};
};
bool launch_attention_bwd_fused(const AttnBwdPlan &p);  // one-kernel backward (5 GEMMs, ordered dQ accumulation); D = 128
bool launch_attention_bwd_wide(const AttnBwdPlan &p);   // two kernels, 128 x 128 tiles, dense or strided operands (the default)
bool launch_attention_bwd_tc(const AttnBwdPlan &p);  // false => caller uses the GEMM-composed generic backward
// P = exp(S - lse) under the causal mask, in place, fp32 / fp64 (generic backward helper)
void launch_attn_probs(void *S, const void *lse, int dtype, int64_t BH, int64_t Sq, int64_t Skv);

// Fused layer normalisation over the last dimension (kernels/norm.cu): y = (x - mean) * rstd * gain with biased variance.
// `layer_norm_supported` is false for dtypes / row lengths the register-resident kernels do not cover (caller composes the op).
bool layer_norm_supported(int dtype, int64_t E, const void *x, const void *gain);
void launch_layer_norm_fwd(const void *x, const void *gain, void *y, float *mean, float *rstd, int dtype, int64_t rows, int64_t E, float eps,
                           bool rms = false);
// one-launch column statistics over [outer, R, inner] fp32 (kernels/norm.cu): mode 0 = norm_stat (mean, invstd), mode 1 = mean_var
bool launch_col_moments(const void *x, void *out0, void *out1, int dtype, int64_t outer, int64_t R, int64_t inner, int mode, bool take_sqrt,
                        double eps);
int layer_norm_bwd_ctas(int64_t rows, bool write_dx);
// mean + unbiased variance (or its sqrt) of each dense fp32 row in ONE pass; false when the shape is not covered
bool launch_row_moments(const void *x, void *mean, void *var, int dtype, int64_t rows, int64_t E, bool take_sqrt);
void launch_layer_norm_bwd(const void *x, const void *gain, const void *dy, const float *mean, const float *rstd, void *dx,
                           float *dgain_partial, int ctas, int dtype, int64_t rows, int64_t E, bool rms = false);

std::string device_info_string();

// host-side 16-bit float helpers (RN-even, NaN-preserving), ref: src/core/include/half.h:195-208,268-290
uint16_t f32_to_f16_bits(float f);
uint16_t f32_to_bf16_bits(float f);
float f16_bits_to_f32(uint16_t h);
float bf16_bits_to_f32(uint16_t h);

}  // namespace kf

/*
 * kfunca_b200 — C ABI of the B200-native tensor-operator hot path.
 *
 * This is the drop-in boundary: plain pointers, sizes and opaque handles, no C++ or torch types.
 * Each entry point names the reference (xytpai/kfunca) interface it replaces as
 * `ref: <file>:<line>` (paths relative to the reference checkout).  The pybind11 module
 * `kfunca_b200._kfunca` (kfunca_b200/csrc/pybind_module.cpp) is written ONLY against this header
 * and re-exposes the surface of the reference's `src/register.cpp`.
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on failure; `kf_last_error()` then returns a
 *    thread-local, human-readable message (the reference throws utils::Error, exception.h:123-131,
 *    which pybind turns into RuntimeError — our pybind shim does the same from the status code).
 *  - `kf_tensor_t` is an owning handle (ref: class Tensor, src/core/include/tensor.h:24-165):
 *    it shares a reference-counted impl; release every handle you receive with `kf_release`.
 *  - device == -1 creates a "meta" tensor (shape/stride/dtype only, no HBM) used to test the host
 *    logic without a GPU; any compute on it fails loudly.
 *  - there is NO CPU compute fallback anywhere behind this ABI.
 */
#ifndef KFUNCA_B200_H_
#define KFUNCA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KF_MAX_DIMS 12 /* ref: src/core/include/tensor.h:7 MAX_TENSOR_DIMS */

/* ref: src/core/include/scalar_type.h:9-27 (enum order is part of the contract) */
typedef enum {
    KF_BOOL = 0,
    KF_BYTE = 1,
    KF_CHAR = 2,
    KF_SHORT = 3,
    KF_INT = 4,
    KF_LONG = 5,
    KF_HALF = 6,
    KF_BFLOAT16 = 7,
    KF_FLOAT = 8,
    KF_DOUBLE = 9,
    KF_UNDEFINED = 10
} kf_dtype_t;

typedef struct kf_tensor_s *kf_tensor_t;
typedef struct kf_event_s *kf_event_t;

/* ---- errors / runtime (ref: Launcher, src/device/launcher_cuda.h:105-354) ------------------ */
const char *kf_last_error(void);
int kf_version(int *major, int *minor);
int kf_device_count(int *count);
/* Make `device` current for this process (one process per GPU). ref: launcher_cuda.h:139-147 —
 * unlike the reference this never calls cudaDeviceReset(). */
int kf_set_device(int device);
int kf_get_device(int *device);
/* The library's single compute stream as a cudaStream_t value (non-default, non-blocking). */
int kf_stream(void **cuda_stream);
int kf_synchronize(void);
/* `t`'s memory is also used by work enqueued on `stream` (a cudaStream_t other than the library's, e.g. the stream of a framework holding
 * a zero-copy alias): when the storage is released the library stream first waits for that stream (replaces the hand-placed joins the
 * reference's single implicit stream never needed, launcher_cuda.h:113-116).  No-op for the library stream itself. */
int kf_record_stream(kf_tensor_t t, void *stream);
/* Human-readable device table. ref: device_info(), src/device/device_info.cu:191-216 */
int kf_device_info(char *buf, size_t buf_len);
/* Pool statistics. ref: DeviceAllocator::print, src/core/device_allocator.cpp:17-35 */
int kf_memstat(char *buf, size_t buf_len);
int kf_mem_stats(int64_t *bytes_in_use, int64_t *bytes_reserved, int64_t *n_device_mallocs);
/* Return every fully-free arena to the driver (the reference never does; device_allocator.h:42-83). */
int kf_empty_cache(void);
/* number of kernels this library launched since load (bench.py's gpu_launches) */
int kf_launch_count(int64_t *count);

/* CUDA-event timing on the library stream (ref: Launcher::set_profiling_mode, launcher_cuda.h:253-255,336-349) */
int kf_event_create(kf_event_t *ev);
int kf_event_record(kf_event_t ev);
int kf_event_synchronize(kf_event_t ev);
int kf_event_elapsed_ms(kf_event_t start, kf_event_t stop, float *ms);
int kf_event_destroy(kf_event_t ev);

/* pinned host staging buffers for the H2D/D2H legs (ref: dmemcpy_h2d/d2h, src/device/memory_engine.cu:6-28) */
/* NUMA node of the process's GPU (-1 when the platform does not say) and the kernel's cpulist string of that node ("0-31,64-95");
 * kf_host_alloc_pinned binds itself to those CPUs while it allocates and first touches the buffer (KF_NUMA=0 disables). */
int kf_numa_info(int *node, char *cpulist, size_t cap);
int kf_host_alloc_pinned(size_t bytes, void **ptr);
int kf_host_free_pinned(void *ptr);

/* ---- creation / host I/O (ref: src/register.cpp:27-57,76-84; src/core/tensor.cpp:17-69) ---- */
int kf_empty(const int64_t *shape, int ndim, int dtype, int device, kf_tensor_t *out);
int kf_zeros(const int64_t *shape, int ndim, int dtype, int device, kf_tensor_t *out);
int kf_empty_like(kf_tensor_t self, kf_tensor_t *out);
/* contiguous row-major host buffer -> new tensor (ref: from_numpy, register.cpp:27-39) */
int kf_from_host(const void *src, const int64_t *shape, int ndim, int dtype, int device, kf_tensor_t *out);
/* tensor (must be contiguous) -> host buffer of numel*itemsize bytes (ref: to_numpy, register.cpp:41-57) */
int kf_to_host(kf_tensor_t self, void *dst, size_t dst_bytes);
/* raw asynchronous copies on the library stream, no implicit sync; host memory should be pinned
 * (ref: dmemcpy_h2d / dmemcpy_d2h, src/device/memory_engine.cu:14-24, which create+destroy a stream per copy) */
int kf_memcpy_h2d_async(void *dst_device, const void *src_host, size_t bytes);
int kf_memcpy_d2h_async(void *dst_host, const void *src_device, size_t bytes);

/* ---- handles / metadata (ref: tensor.h:42-124, register.cpp:88-139) ------------------------ */
int kf_retain(kf_tensor_t self, kf_tensor_t *out); /* new handle, same impl (ref: Tensor copy ctor, register.cpp:89-90) */
int kf_release(kf_tensor_t self);
int kf_defined(kf_tensor_t self, int *out);
int kf_dim(kf_tensor_t self, int *out);
int kf_numel(kf_tensor_t self, int64_t *out);
int kf_dtype(kf_tensor_t self, int *out);
int kf_device(kf_tensor_t self, int *out);
int kf_shape(kf_tensor_t self, int d, int64_t *out);
int kf_sizes(kf_tensor_t self, int64_t *out /* [KF_MAX_DIMS] */, int *ndim);
int kf_strides(kf_tensor_t self, int64_t *out /* [KF_MAX_DIMS] */, int *ndim);
int kf_storage_offset(kf_tensor_t self, int64_t *out);
int kf_is_contiguous(kf_tensor_t self, int *out);
int kf_data_ptr(kf_tensor_t self, void **out);
int kf_storage_bytes(kf_tensor_t self, size_t *out);
int kf_storage_ref_count(kf_tensor_t self, int64_t *out);
int kf_impl_ref_count(kf_tensor_t self, int64_t *out);
int kf_element_size(int dtype, size_t *out);
/* one element -> 8-byte host scratch in the tensor's own dtype (ref: Tensor::item, tensor.cpp:136-147) */
int kf_item(kf_tensor_t self, const int64_t *indices, int n, void *out8);
/* multi-line text like the reference's operator<< (ref: tensor.cpp:323-377) */
int kf_to_string(kf_tensor_t self, char *buf, size_t buf_len);

/* ---- view algebra, zero-copy (ref: src/core/tensor.cpp:161-321, tensor_impl.cpp:67-102) ---- */
int kf_as_strided(kf_tensor_t self, const int64_t *sizes, const int64_t *strides, int ndim, int64_t storage_offset, kf_tensor_t *out);
int kf_permute(kf_tensor_t self, const int64_t *dims, int ndim, kf_tensor_t *out);
int kf_view(kf_tensor_t self, const int64_t *sizes, int ndim, kf_tensor_t *out);
int kf_slice(kf_tensor_t self, int64_t dim, int64_t start, int64_t end, int64_t step, kf_tensor_t *out);
int kf_select(kf_tensor_t self, int64_t dim, int64_t index, kf_tensor_t *out);
int kf_narrow(kf_tensor_t self, int64_t dim, int64_t start, int64_t length, kf_tensor_t *out);
/* ref: gpu::tensor_split, src/core/tensor_shape.cpp:72-89; outs has room for n handles */
int kf_split(kf_tensor_t self, const int64_t *sizes, int n, int64_t dim, kf_tensor_t *outs);
/* returns self (new handle) if already contiguous, else a materialised copy (ref: tensor.cpp:161-165) */
int kf_contiguous(kf_tensor_t self, kf_tensor_t *out);

/* ---- elementwise (ref: src/core/binary_ops.cpp:6-93, src/device/binary_ops_kernel.cu:34-60) - */
typedef enum { KF_OP_ADD = 0, KF_OP_SUB = 1, KF_OP_MUL = 2, KF_OP_DIV = 3 } kf_binary_op_t;
int kf_binary(int op, kf_tensor_t a, kf_tensor_t b, kf_tensor_t *out);  /* out = a op b, promoted dtype */
int kf_binary_(int op, kf_tensor_t self, kf_tensor_t other);            /* self = self op other, in place */
/* scalar operand: same result as the reference's fill-a-tensor-then-op (register.cpp:172-206) without the temp */
int kf_binary_scalar(int op, kf_tensor_t a, double scalar, kf_tensor_t *out);
int kf_binary_scalar_(int op, kf_tensor_t self, double scalar);
int kf_fill_(kf_tensor_t self, double value);              /* ref: gpu::fill_, src/core/nullary_ops.cpp:6-14 */
int kf_copy_(kf_tensor_t self, kf_tensor_t src);           /* ref: gpu::copy_, src/core/unary_ops.cpp:13-17 */
int kf_clone(kf_tensor_t self, kf_tensor_t *out);          /* ref: gpu::clone, unary_ops.cpp:7-11 */
int kf_convert(kf_tensor_t self, int dtype, kf_tensor_t *out); /* ref: gpu::convert, unary_ops.cpp:19-24 */
/* unary maths needed by the transformer block's norm (the reference folds sqrt into mean_var, reduce_ops_kernel.cu:120-126) */
typedef enum { KF_UOP_SQRT = 0, KF_UOP_RSQRT = 1, KF_UOP_NEG = 2 } kf_unary_op_t;
int kf_unary(int op, kf_tensor_t a, kf_tensor_t *out);

/* ---- reductions, keepdim (ref: src/core/reduce_ops.cpp:8-30, src/device/reduce_ops_kernel.cu) */
int kf_sum(kf_tensor_t self, int64_t dim, kf_tensor_t *out);
int kf_mean(kf_tensor_t self, int64_t dim, kf_tensor_t *out);
int kf_mean_var(kf_tensor_t self, int64_t dim, int take_sqrt, kf_tensor_t *mean, kf_tensor_t *var);
int kf_norm_stat(kf_tensor_t self, int64_t dim, kf_tensor_t *mean, kf_tensor_t *invstd); /* ref: src/core/norm_ops.cpp */

/* ---- sort / top-k (ref: src/core/sort_ops.cpp:6-19, src/device/sort_ops_kernel.cu:553-632) - */
int kf_sort(kf_tensor_t self, int64_t dim, int descending, kf_tensor_t *values, kf_tensor_t *indices);
int kf_topk(kf_tensor_t self, int64_t k, int64_t dim, int largest, kf_tensor_t *values, kf_tensor_t *indices);

/* ---- concat / indexing (ref: src/core/tensor_shape.cpp:41-70, src/core/index_ops.cpp:6-38) - */
int kf_cat(const kf_tensor_t *tensors, int n, int64_t dim, kf_tensor_t *out);
int kf_index_put_(kf_tensor_t self, const kf_tensor_t *indices, int n, kf_tensor_t values);

/* ---- dense contractions (ref: src/core/gemm_ops.cpp:6-16, src/core/nn_ops.cpp:6-8) --------- */
/* out[...,N] = alpha * a[...,K] @ b[K,N] (+ beta*out for kf_gemm_out); fp32/fp64 SIMT, fp16/bf16 tcgen05 */
int kf_gemm(kf_tensor_t a, kf_tensor_t b, float alpha, float beta, kf_tensor_t *out);
int kf_gemm_out(kf_tensor_t out, kf_tensor_t a, kf_tensor_t b, float alpha, float beta);
/* gemm with operands and result in HOST memory (row-major, contiguous; pin them with kf_host_alloc_pinned for full PCIe rate):
 * the path a kfunca user writes as from_numpy(a), from_numpy(b), gemm, numpy() (register.cpp:27-57,86 around gemm_ops.cpp:10-16),
 * with the upload of A in M-slabs, the slab products and the download of C overlapped on three streams.  Asynchronous:
 * the library stream is ordered after all of it, kf_synchronize() makes c_host readable.  slab_rows <= 0 picks 1024. */
int kf_gemm_host(const void *a_host, const void *b_host, void *c_host, int64_t M, int64_t N, int64_t K, int dtype, float alpha,
                 int64_t slab_rows);
/* general form used by the backward passes: op(a) is a or a^T over the last two dims, batched over leading dims */
int kf_matmul(kf_tensor_t a, int trans_a, kf_tensor_t b, int trans_b, float alpha, kf_tensor_t *out);
/* q,k,v: [B,H,S,D] contiguous; top-left-aligned causal mask, scale 1/sqrt(D) (ref: causal_attention_kernel.cu:9-72) */
int kf_causal_attention(kf_tensor_t q, kf_tensor_t k, kf_tensor_t v, kf_tensor_t *out);
/* forward that also returns the row log-sum-exp [B,H,Sq] (fp32) needed by the backward */
int kf_causal_attention_fwd(kf_tensor_t q, kf_tensor_t k, kf_tensor_t v, kf_tensor_t *out, kf_tensor_t *lse);
int kf_causal_attention_bwd(kf_tensor_t dout, kf_tensor_t q, kf_tensor_t k, kf_tensor_t v, kf_tensor_t out,
                            kf_tensor_t lse, kf_tensor_t *dq, kf_tensor_t *dk, kf_tensor_t *dv);

/* Fused layer norm over the last dimension: y = (x - mean) / sqrt(var + eps) * gain, biased variance, gain has size(-1)
 * elements; differentiable in x and gain.  The reference stops at the statistics (mean_var, norm_stat:
 * src/device/reduce_ops_kernel.cu:61-153, src/device/norm_ops_kernel.cu:6-61) and lists the fused norm as its next op
 * (README.md:28); SURVEY §8f rank 1. */
int kf_layer_norm(kf_tensor_t x, kf_tensor_t gain, double eps, kf_tensor_t *out);
/* y = x / sqrt(mean(x^2) + eps) * gain over the last dimension, same kernels as kf_layer_norm without the centring; the op the
 * reference's README names as its next one (ref: README.md:28 `rms_norm`; statistics: src/device/reduce_ops_kernel.cu:61-153). */
int kf_rms_norm(kf_tensor_t x, kf_tensor_t gain, double eps, kf_tensor_t *out);

/* ---- fused linear ops (ref: README.md:32 `qkv_linear`, src/core/gemm_ops.cpp:6-16) ---------------------- */
/* out = alpha * a[...,K] @ b[K,N] + residual[...,N]: the residual add runs in the GEMM epilogue; bit-identical to kf_gemm + kf_binary(ADD) */
int kf_gemm_residual(kf_tensor_t a, kf_tensor_t b, kf_tensor_t residual, float alpha, kf_tensor_t *out);
/* out[..., N] = x[..., K] @ w[K, N] + bias[N] (bias may be NULL): the projection the reference's README lists as its next operator
 * (ref: README.md:32 `qkv_linear`; it would sit on gpu::gemm, src/core/gemm_ops.cpp:6-16).  The bias row is added in the GEMM epilogue;
 * bit-identical to kf_gemm followed by a broadcast kf_binary(ADD); differentiable in x, w and bias.  Feed the result to
 * kf_qkv_attention to run attention on the packed projection in place. */
int kf_qkv_linear(kf_tensor_t x, kf_tensor_t w, kf_tensor_t bias, kf_tensor_t *out);
/* out = (a @ b1) * (a @ b3), the bilinear GLU, as ONE dual-B tcgen05 kernel (fp16 / bf16; other dtypes and small shapes are composed) */
int kf_gemm_glu(kf_tensor_t a, kf_tensor_t b1, kf_tensor_t b3, kf_tensor_t *out);

/* out[B,S,H*D] = causal attention over the packed projection qkv[B,S,3*H*D] (q | k | v along the last dim, as kf_gemm(x, Wqkv) produces
 * it): the kernels read q, k, v in place through strided TMA maps and write the merged-head layout, forward and backward — the
 * split / view / permute / contiguous chain around kf_causal_attention disappears.  Composed from those ops when the tcgen05 path
 * does not take the dtype / head size. */
int kf_qkv_attention(kf_tensor_t qkv, int64_t heads, kf_tensor_t *out);

/* ---- embedding (ref: README.md:30 `embedding`, gather/scatter of src/device/utils/tensor_index.h:19-143,
 * src/core/index_ops.cpp:6-38) --------------------------------------------------------------------------------------------- */
/* out[..., :] = weight[indices[...], :] (indices int64, negative wraps); differentiable in weight: the backward is a deterministic
 * scatter-add (stable sort of the ids, runs summed in position order, no atomics) */
int kf_embedding(kf_tensor_t weight, kf_tensor_t indices, kf_tensor_t *out);

/* counter-based uniform fill in place: element i = lo + (hi - lo) * u(i, seed), reproducible on the host (oracle.counter_uniform).
 * Test / bench input generator for the BASELINE-size tensors (no reference counterpart; the reference's tests draw NumPy data). */
int kf_random_uniform_(kf_tensor_t self, uint64_t seed, double lo, double hi);

/* ---- autograd (ref: GradFunction/backward, tensor.h:18-22, tensor.cpp:71-126) -------------- */
int kf_requires_grad(kf_tensor_t self, int *out);
int kf_set_requires_grad(kf_tensor_t self, int flag);
int kf_backward(kf_tensor_t self, kf_tensor_t grad_output);
int kf_grad(kf_tensor_t self, kf_tensor_t *out); /* *out == NULL when no grad has been accumulated */
int kf_zero_grad(kf_tensor_t self);
/* Leaf-gradient hook (no reference counterpart: the reference has no multi-GPU path).  `fn` runs on the calling
 * thread inside kf_backward each time a leaf tensor's gradient for this pass has been enqueued on the library stream; `leaf`
 * and `grad` are borrowed handles valid only during the call.  The data-parallel layer uses it to start the NCCL all-reduce of
 * each gradient on a side stream while the rest of the backward pass is still running.  fn == NULL removes the hook. */
typedef void (*kf_leaf_grad_hook_t)(kf_tensor_t leaf, kf_tensor_t grad, void *ctx);
int kf_set_leaf_grad_hook(kf_leaf_grad_hook_t fn, void *ctx);

/* ---- data-parallel layer (no reference counterpart: SURVEY 8e; the reference is single-GPU, launcher_cuda.h:105-354) -----------
 * One process per GPU.  NCCL (dlopen'ed libnccl.so.2) over NVLink / NVSwitch on a library-owned communication stream.
 * Rendezvous is the caller's business: rank 0 obtains a 128-byte id and ships it to the other ranks (kfunca_b200/dist.py does it
 * over a TCP socket on MASTER_ADDR), then every rank calls kf_dist_init after kf_set_device. */
int kf_dist_unique_id(void *out128);
int kf_dist_init(const void *id128, int rank, int world);
int kf_dist_finalize(void);
int kf_dist_info(int *initialised, int *rank, int *world, int *nccl_version);
typedef enum { KF_RED_SUM = 0, KF_RED_AVG = 1, KF_RED_MAX = 2 } kf_red_op_t;
/* in-place all-reduce on the LIBRARY stream (ordered with the kernels around it): the cross-shard sum / mean of SURVEY 8e */
int kf_dist_all_reduce(kf_tensor_t t, int op);
/* overlapped gradient averaging: between begin and end, kf_backward starts ncclAllReduce(AVG) of the gradient of every listed
 * parameter on the communication stream the moment that gradient has been enqueued; end makes the library stream wait for the
 * communication stream and returns the number of gradients reduced.  Call kf_zero_grad on the parameters before each backward
 * (gradients accumulated over several backward passes inside one begin/end would be averaged more than once). */
int kf_dist_overlap_begin(const kf_tensor_t *params, int n);
int kf_dist_overlap_end(int64_t *n_reduced);

/* ---- host-logic probes (no GPU needed; used by the `-m "not gpu"` tests) ------------------- */
/* broadcast + dtype promotion + dimension collapse of `a op b` exactly as the engine plans it.
 * Outputs: collapsed ndim, shape[ndim], byte strides for out/a/b [3][KF_MAX_DIMS], common dtype.
 * ref: TensorIterator::build, src/core/tensor_iterator.cpp:486-515 */
int kf_debug_plan_binary(kf_tensor_t a, kf_tensor_t b, int *ndim, int64_t *shape, int64_t *strides3, int *common_dtype);
/* ref: update_common_dtype, src/core/tensor_iterator.cpp:32-44 */
int kf_promote_types(int a, int b, int *out);
/* drive the pool allocator against a fake address space: ops[i] > 0 allocates ops[i] bytes, ops[i] < 0 frees
 * the block returned by op number (-ops[i]-1). offsets[i] receives the fake address (or -1 for frees);
 * stats = {bytes_in_use, bytes_reserved, n_arena_mallocs}. */
int kf_debug_pool_trace(const int64_t *ops, int n, int64_t *offsets, int64_t *stats3);
/* the same fake pool with side streams: kind[i] 0 = allocate arg[i] bytes, 1 = release the block of op arg[i], 2 = record that stream
 * id arg2[i] uses the block of op arg[i].  fence_log receives, in order, the stream id of every fence the pool issues while releasing
 * (capacity cap); *nfences = how many it issued. */
int kf_debug_pool_fences(const int *kind, const int64_t *arg, const int64_t *arg2, int n, int64_t *fence_log, int cap, int *nfences);

#ifdef __cplusplus
}
#endif
#endif /* KFUNCA_B200_H_ */

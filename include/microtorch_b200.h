/*
 * microtorch_b200.h - C ABI of libmicrotorch_b200.so
 *
 * The drop-in boundary for MicroTorch's device="cuda" path.  The reference
 * (nickovchinnikov/microtorch) has no FFI: its CUDA path is "substitute the cupy module
 * for numpy" (src/tensor/device.py:74-84).  Each entry point below names the reference
 * call site(s) (paths relative to the reference checkout) whose CuPy work it replaces.
 * Plain pointers and sizes only; no torch, no C++ types.
 *
 * Conventions
 *   - every function returns 0 (MT_OK) or a non-zero mt_status; the message for the last
 *     failure on the calling thread is mt_last_error().
 *   - all device pointers come from mt_malloc (caching pool); all arithmetic is float32
 *     ("everything is fp32").
 *   - one compute stream per process (one process per GPU); calls are asynchronous with
 *     respect to the host unless stated (mt_d2h, mt_sync, mt_event_sync block).
 *   - shapes/strides arrays are read during the call only; strides are in ELEMENTS and a
 *     stride of 0 broadcasts that dimension.
 *   - there is no CPU fallback: without a usable B200 every compute entry point fails.
 */
#ifndef MICROTORCH_B200_H
#define MICROTORCH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MT_MAX_DIMS 8

typedef enum mt_status {
  MT_OK = 0,
  MT_ERR_CUDA = 1,        /* a CUDA runtime/driver call failed */
  MT_ERR_INVALID = 2,     /* bad argument (shape, alignment, enum) */
  MT_ERR_NOT_INIT = 3,    /* mt_init has not succeeded in this process */
  MT_ERR_OOM = 4,         /* device allocation failed even after trimming the pool */
  MT_ERR_NCCL = 5,        /* NCCL missing or a collective failed */
  MT_ERR_UNSUPPORTED = 6  /* valid request this library does not implement */
} mt_status;

/* ---------------------------------------------------------------- runtime --------- */
/* replaces: CUDA_AVAILABLE = cp.cuda.is_available()  (src/tensor/device.py:7-11)      */
int mt_device_count(int* count);
/* replaces: CuPy's implicit context/stream creation on first use (device.py:74-84)    */
int mt_init(int device);
int mt_shutdown(void);
int mt_is_initialized(void);
const char* mt_last_error(void);
const char* mt_version(void);
int mt_device_info(char* name, size_t name_cap, int* sm_count, size_t* hbm_bytes,
                   int* cc_major, int* cc_minor);
/* replaces: the implicit device sync at .item()/.get()/print (src/tensor/tensor.py:175-183) */
int mt_sync(void);
/* the library's compute stream as a cudaStream_t (for interop/timing only)             */
int mt_stream(void** cuda_stream);
/* CUDA events on the compute stream (bench timing; demo.ipynb:5176,5194 uses time.time) */
int mt_event_create(void** ev);
int mt_event_record(void* ev);
int mt_event_sync(void* ev);
int mt_event_elapsed_ms(void* start, void* stop, float* ms);
int mt_event_destroy(void* ev);
/* number of kernels this library has launched since the last reset (bench "gpu_launches") */
int mt_launch_count(uint64_t* count);
int mt_launch_count_reset(void);
/* Whole-step CUDA graphs (the user loop demo.ipynb:5178-5189 is launch-latency-bound at the
 * spiral-demo sizes).  Every launch made between begin and end is captured instead of executed; blocking calls
 * (mt_h2d, mt_d2h, mt_sync) are not allowed inside.  Pool blocks allocated during the capture stay reserved for the
 * graph (its kernels write them on every replay) until mt_graph_destroy.                                        */
int mt_graph_begin(void);
int mt_graph_end(void** graph_handle);
int mt_graph_launch(void* graph_handle);
int mt_graph_destroy(void* graph_handle);
/* timing hygiene between timed iterations: overwrite a buffer larger than L2, then read
 * another one, so that L2 is cold AND holds no dirty lines whose write-back the timed
 * kernel would pay for                                                                  */
int mt_l2_flush(void);
/* per-kernel timing of the dominant kernel (bench roofline): when enabled, every tcgen05
 * GEMM launch is bracketed by CUDA events on the compute stream; mt_prof_gemm synchronises
 * and returns the launch count, the summed device time and the summed logical flops
 * (2*M*N*K) since the last reset.                                                       */
int mt_prof_enable(int on);
int mt_prof_reset(void);
int mt_prof_gemm(uint64_t* launches, double* ms_total, double* flops_total);
/* same, plus the flops weighted by what each launch costs on the tensor pipe in bf16-rate units per logical
 * multiply-add (TF32 product = 2, BF16 product = 1: 4 for the TF32 + 2xBF16 split, 6 for 3xTF32) - the numerator of
 * the roofline fraction against the measured bf16 peak when launches of both kinds are mixed                   */
int mt_prof_gemm_units(uint64_t* launches, double* ms_total, double* flops_total, double* unit_flops_total);

/* ---------------------------------------------------------------- memory ---------- */
/* replaces: CuPy's default MemoryPool behind every cp.array / zeros_like / operator
 * result (src/tensor/tensor.py:77,205,243; every op in src/tensor/ops/)                  */
int mt_malloc(void** ptr, size_t bytes);
int mt_free(void* ptr);
int mt_pool_stats(size_t* reserved_bytes, size_t* in_use_bytes, size_t* peak_in_use_bytes,
                  uint64_t* n_requests, uint64_t* n_device_mallocs);
int mt_pool_trim(void);
/* pinned host staging (end-to-end path: H2D of a step's inputs, D2H of its loss)        */
int mt_host_alloc(void** ptr, size_t bytes);
int mt_host_free(void* ptr);
/* replaces: cp.array(np_array) (tensor.py:205,221) / .get() (tensor.py:218)             */
int mt_h2d(void* dst, const void* src, size_t bytes);        /* blocks until src is reusable */
int mt_d2h(void* dst, const void* src, size_t bytes);        /* blocks until dst is valid    */
int mt_h2d_async(void* dst, const void* pinned_src, size_t bytes);
int mt_d2h_async(void* pinned_dst, const void* src, size_t bytes);
/* double-buffered input pipeline: upload on a copy stream while the compute stream works; the copy is ordered
 * after everything queued on the compute stream at the time of the call (the destination may be a recycled
 * block), mt_prefetch_wait() orders the compute stream after all prefetches issued so far.              */
int mt_h2d_prefetch(void* dst, const void* pinned_src, size_t bytes);
int mt_prefetch_wait(void);
int mt_d2d(void* dst, const void* src, size_t bytes);
/* replaces: zeros_like + grad.fill(0.0) (tensor.py:77-85,243-248; module.py:83-89)      */
int mt_memset_zero(void* ptr, size_t bytes);
int mt_memset_multi(void* const* ptrs, const size_t* bytes, int count);
int mt_fill_f32(float* out, size_t n, float value);
/* replaces: implicit strided reads of transposed/sliced/broadcast views - .copy(),
 * .get() of a view, np.broadcast_to materialisation (base.py:63,111; tensor.py:380)     */
int mt_copy_strided_f32(float* out, const float* in, int ndim, const int64_t* shape,
                        const int64_t* in_strides);
/* replaces: full_grad[index] = grad for basic (slice) indices (overload.py:36-38)       */
int mt_scatter_strided_f32(float* out, const int64_t* out_strides, const float* in, int ndim,
                           const int64_t* shape);
/* replaces: tensor.data[index] / full_grad[index] = grad with an integer index array
 * on axis 0 (overload.py:28,37).  Rows are `row_elems` contiguous floats; the scatter
 * ASSIGNS (last writer wins is avoided: duplicates resolve to the highest position,
 * as NumPy does).                                                                        */
int mt_gather_rows_f32(float* out, const float* in, const int64_t* dev_index, int64_t n_index,
                       int64_t row_elems, int64_t n_rows);
int mt_scatter_rows_assign_f32(float* out, const float* in, const int64_t* dev_index,
                               int64_t n_index, int64_t row_elems, int64_t n_rows);

/* ---------------------------------------------------------------- elementwise ----- */
typedef enum mt_binary_op {
  MT_BIN_ADD = 0,       /* a + b            overload.py:70;  tensor.py:279 (grad accumulate)   */
  MT_BIN_MUL = 1,       /* a * b            overload.py:94,109                                 */
  MT_BIN_SUB = 2,       /* a - b            (a + (-b), tensor.py:556-558, fused)               */
  MT_BIN_DIV = 3,       /* a / b            math.py:12 (log backward g / x)                    */
  MT_BIN_TANH_BWD = 4,  /* a * (1 - b*b)    math.py:74   a = grad, b = saved tanh OUTPUT       */
  MT_BIN_POW_BWD = 5,   /* a * (p * b**(p-1))  math.py:51   a = grad, b = saved input, p = scalar */
  MT_BIN_EXP_BWD = 6,   /* a * b            math.py:31   a = grad, b = saved exp output        */
  /* comparison / select family (src/tensor/tensor.py:514-536, src/tensor/ops/elementwise.py): results are
   * float32 masks 1.0 / 0.0 ON THE DEVICE (the reference builds CPU tensors here and cannot run them under CuPy, Q8) */
  MT_BIN_LT = 7, MT_BIN_GT = 8, MT_BIN_LE = 9, MT_BIN_GE = 10, MT_BIN_EQ = 11, MT_BIN_NE = 12,
  MT_BIN_MAX = 13,      /* max(a, b)        tensor.py:291 clip lower bound                     */
  MT_BIN_MIN = 14       /* min(a, b)        tensor.py:291 clip upper bound                     */
} mt_binary_op;

/* out[idx] = op(a[idx . a_strides], b[idx . b_strides]) over the broadcast `shape`;
 * out is contiguous.  Vectorised (128-bit) whenever the collapsed layout allows it.     */
int mt_ew_binary_f32(int op, float* out, const float* a, const float* b, int ndim,
                     const int64_t* shape, const int64_t* a_strides, const int64_t* b_strides,
                     float scalar);
/* in-place broadcasting update  dst[idx] (+=|*=|/=) src[idx . src_strides]
 * replaces: np.add/multiply/true_divide(self.data, other, out=self.data)
 * (tensor.py:547-550 bias add, :581-588, :609-616).  dst may itself be strided.         */
int mt_ew_inplace_f32(int op, float* dst, const int64_t* dst_strides, const float* src,
                      const int64_t* src_strides, int ndim, const int64_t* shape);

typedef enum mt_unary_op {
  MT_UN_NEG = 0,        /* -x          overload.py:52               */
  MT_UN_LOG = 1,        /* log x       math.py:8                    */
  MT_UN_TANH = 2,       /* tanh x      math.py:70                   */
  MT_UN_EXP = 3,        /* exp x       math.py:27                   */
  MT_UN_POW = 4,        /* x ** s      math.py:47 (NumPy fast paths for s in {-1,0,0.5,1,2}) */
  MT_UN_SCALE = 5,      /* x * s       optimizer.py:39 (lr * grad); tensor.py:310            */
  MT_UN_ADDS = 6,       /* x + s                                                            */
  MT_UN_ABS = 7         /* |x|                                                              */
} mt_unary_op;
int mt_ew_unary_f32(int op, float* out, const float* x, size_t n, float scalar);

/* out[idx] = cond[idx . c_strides] != 0 ? a[idx . a_strides] : b[idx . b_strides]   (out contiguous)
 * replaces: tnsr.where(condition.data, a.data, b.data) and its two backward selects
 * (src/tensor/ops/elementwise.py:10,16,21)                                                */
int mt_ew_where_f32(float* out, const float* cond, const float* a, const float* b, int ndim,
                    const int64_t* shape, const int64_t* c_strides, const int64_t* a_strides,
                    const int64_t* b_strides);

/* ---------------------------------------------------------------- reductions ------ */
/* x viewed as [outer, red, inner] (contiguous); out[o,i] = sum_r x[o,r,i] * w[o,r,i]
 * where w is `other` addressed with strides (so, sr, si) or 1 when other == NULL.
 * Deterministic (fixed tree, no float atomics).
 * replaces: tlib.sum(x, axis, keepdims) (reduce.py:11); grad.sum(), grad.sum(axis=...)
 * in BaseOps.bkwd_broadcast (base.py:28-54) and, with `other`, the `grad * b.data`
 * that precedes it in mul backward (overload.py:109-112).                               */
int mt_reduce_sum_f32(float* out, const float* x, int64_t outer, int64_t red, int64_t inner,
                      const float* other, int64_t other_so, int64_t other_sr, int64_t other_si);

/* out[o,i] = max_r (is_max != 0) or min_r x[o,r,i], x contiguous [outer, red, inner], red > 0.  NaN propagates.
 * replaces: tlib.max / tlib.min(x, axis, keepdims) in Reduce.max / Reduce.min (reduce.py:66,87); their backward
 * (reduce.py:46-63) is composed from MT_BIN_EQ, mt_reduce_sum_f32, MT_BIN_DIV and MT_BIN_MUL.            */
int mt_reduce_extreme_f32(int is_max, float* out, const float* x, int64_t outer, int64_t red, int64_t inner);

/* ---------------------------------------------------------------- Linear / matmul -- */
typedef enum mt_act { MT_ACT_NONE = 0, MT_ACT_TANH = 1 } mt_act;

/* Y[M,N] = act( X[M,K] . W[N,K]^T + bias[N] )      X, W, Y row-major, contiguous
 * replaces: Linear.forward = (W @ x^T)^T, `output += bias` and the following Tanh
 * (src/nn/layer/linear.py:48-62, src/nn/activation.py:11-12, overload.py:132).
 * tcgen05/TMEM tiles fed by TMA, fp32 kept by an error-compensated split done on chip:
 *   x = hi + lo, hi = trunc_tf32(x);  A.B ~= A_hi.B_hi (TF32) + A_lo.B_hi + A_hi.B_lo (both as BF16 products,
 *   8 bits are enough for terms that are 2^-10 of the result; lo.lo ~ 2^-22 dropped), fp32 accumulation folded
 *   every K=256 into IEEE running sums.  This is the north_star's "3xTF32 split" with the two correction
 *   products moved to BF16: same measured accuracy (1e-6 forward, < 2e-6 backward), 1.2-1.4x faster;
 *   mt_gemm_tc_set_mix(0) selects the literal all-TF32 split.  bias add + tanh fused in the epilogue.     */
int mt_linear_fwd_f32(float* Y, const float* X, const float* W, const float* bias_or_null,
                      int64_t M, int64_t N, int64_t K, int act);
/* dZ = dY * (1 - Yact^2) if Yact != NULL else dY          (tanh' fused in the prologue)
 * dW[N,K] (+)= dZ^T . X ;  dX[M,K] = dZ . W   (dX skipped when NULL).  No bias gradient:
 * the reference's bias is outside the tape (linear.py:61-62, tensor.py:547-550; Q1).
 * replaces: matmul.bkwd_a / bkwd_b (overload.py:137-154) and tanh backward (math.py:74) */
#define MT_LINEAR_ACCUMULATE_DW 1            /* dW += instead of dW =                                        */
#define MT_LINEAR_DX_TIMES_TANH_GRAD_OF_X 2 /* dX <- dX * (1 - X^2): X is itself a tanh output, so the dX    */
                                            /* epilogue emits the PRODUCING layer's pre-activation gradient  */
                                            /* (its dZ) and that layer then runs plain GEMMs (Yact = NULL)    */
int mt_linear_bwd_f32(float* dX_or_null, float* dW, const float* dY, const float* Yact_or_null,
                      const float* X, const float* W, int64_t M, int64_t N, int64_t K, int flags);
/* generic C[b] = A[b] @ B[b] for the bare `@` (overload.py:132,142,153): element strides
 * for every operand, batch stride 0 broadcasts.  C is row-major [batch, M, N].          */
int mt_matmul_f32(float* C, const float* A, const float* B, int64_t batch, int64_t M, int64_t N,
                  int64_t K, int64_t a_sb, int64_t a_sm, int64_t a_sk, int64_t b_sb, int64_t b_sk,
                  int64_t b_sn);
/* fp32 emulation level of the tensor-core GEMMs (like torch.set_float32_matmul_precision):
 *   0 "high"    every GEMM as TF32 hi.hi + 2 BF16 correction products (1.4e-6 rms per GEMM): the fast, opt-in mode;
 *   1 "higher"  DEFAULT - forward GEMMs as 3xTF32 (all three products on TF32, 9e-7), backward GEMMs as level 0;
 *               +10-12 % per training step over level 0;
 *   2 "highest" 3xTF32 with short accumulation chunks everywhere (4.8e-7 = SGEMM class); +40 %.
 * Every level is far inside the 1e-4 GEMM tolerance per product.  Deep high-gain stacks amplify FORWARD rounding
 * (BASELINE config #5 at full width: 1.3e-4 / 7.5e-5 / 5.3e-5 from the NumPy oracle end to end at levels 0 / 1 / 2),
 * which is why level 1 is the default: the cheapest level that keeps every BASELINE config within 1e-4. */
int mt_gemm_set_precision(int level);
/* which GEMM engine the last mt_linear_ / mt_matmul_ call used: 1 = tcgen05, 0 = SIMT fp32 */
int mt_gemm_last_engine(int* engine);
/* force the engine: -1 auto (default), 0 SIMT only, 1 tcgen05 persistent CTA-pair kernel only, 2 tcgen05
 * mid-size kernel only (error if the shape is not supported by the forced engine).  Test/diagnostic hook. */
int mt_gemm_set_engine(int engine);

/* ---------------------------------------------------------------- losses ---------- */
typedef enum mt_loss_kind {
  MT_LOSS_MSE = 0,      /* mean((p - t)^2)                          src/nn/loss.py:63-76  */
  MT_LOSS_CE_REF = 1,   /* mean(log p - t)  [the reference's formula] src/nn/loss.py:79-84 */
  MT_LOSS_BCE = 2       /* mean(-(t log p + (1-t) log(1-p)))        src/nn/loss.py:87-100 */
} mt_loss_kind;
/* loss_out[0] = scale * sum_i f(pred_i, target_i);  dpred_i = scale * f'(pred_i, target_i)
 * (dpred skipped when NULL).  One pass: read p, read t, write dp.
 * replaces: the 6/7/13-node expression graphs of loss.py + Reduce.mean (reduce.py:40-43)
 * and their backward chains.  `scale` is float32(1/N) of the GLOBAL element count.      */
int mt_loss_fwd_bwd_f32(int kind, float* loss_out, float* dpred_or_null, const float* pred,
                        const float* target, size_t n, float scale);

/* ---------------------------------------------------------------- optimizer ------- */
/* p_i -= lr * g_i for every tensor in one launch; optionally zero g_i in the same pass.
 * replaces: SGD.step (src/nn/optimizer.py:37-39 -> tensor.py:564-571: 3 kernels + 2
 * temporaries per Parameter) and the following Module.zero_grad (module.py:83-89).      */
int mt_sgd_multi_f32(float* const* params, float* const* grads, const size_t* sizes, int count,
                     float lr, int zero_grads);

/* ---------------------------------------------------------------- whole-loop fusion */
/* `steps` full training steps of a small Linear(+Tanh) stack in ONE launch: weights, activations and
 * gradients stay in the shared memory of one CTA; per step HBM receives one loss value.
 *     for k in range(steps): model.zero_grad(); pred = model(X) [* pred_scale + pred_shift];
 *                            loss = loss_fn(pred, T); loss.backward(); optimizer.step()
 * replaces: the user loop demo.ipynb:5178-5189 over Sequential.forward (src/nn/sequential.py:37-48),
 * Linear.forward (src/nn/layer/linear.py:48-62), Tanh (src/nn/activation.py:11-12), Loss.forward
 * (src/nn/loss.py:35-100), Tensor.backward (src/tensor/tensor.py:256-283) and SGD.step
 * (src/nn/optimizer.py:37-39), for full-batch X [batch, dims[0]] / T [batch * dims[n_layers]] that do
 * not change between steps.  weight[i] is [dims[i+1], dims[i]] row-major and is updated in place;
 * bias[i] (may be NULL) is read only - the reference gives biases no gradient (in-place add).
 * grad_weight[i] (may be NULL) receives the LAST step's dL/dW_i, what Parameter.grad holds after the loop.
 * losses[k] = loss of step k (device array of `steps` floats).  loss_scale = float32(1/element count).
 * MT_ERR_UNSUPPORTED when model + batch exceed the 227 KB of one CTA: use the per-step path then.       */
#define MT_MLP_MAX_LAYERS 8
typedef struct mt_mlp_desc {
  int32_t n_layers;
  int32_t dims[MT_MLP_MAX_LAYERS + 1];
  int32_t act[MT_MLP_MAX_LAYERS];              /* mt_act after layer i */
  float* weight[MT_MLP_MAX_LAYERS];
  const float* bias[MT_MLP_MAX_LAYERS];
  float* grad_weight[MT_MLP_MAX_LAYERS];
} mt_mlp_desc;
int mt_mlp_fit_f32(const mt_mlp_desc* desc, const float* X, const float* T, int64_t batch, int loss_kind,
                   int has_pred_map, float pred_scale, float pred_shift, float lr, float loss_scale, int steps,
                   float* losses);

/* ---------------------------------------------------------------- random ---------- */
/* replaces: cupy.random.randn / cupy.random.uniform behind Tensor.randn / Tensor.uniform on device="cuda"
 * (src/tensor/tensor.py:402-434).  A specified counter-based stream - Philox4x32-10, key = seed, block b = counter
 * (b, 0), four words per block - so results are reproducible and checkable on the host (tests/test_rng_gpu.py);
 * CuPy's own stream cannot be reproduced, parity with the reference is by distribution.  Each call consumes
 * ceil(n / 4) blocks starting at `offset`; the caller keeps the offset.
 *   uniform: out = low + (high - low) * ((x >> 8) + 0.5) * 2^-24        normal: Box-Muller on word pairs          */
int mt_rng_uniform_f32(float* out, size_t n, unsigned long long seed, unsigned long long offset, float low, float high);
int mt_rng_normal_f32(float* out, size_t n, unsigned long long seed, unsigned long long offset);

/* ---------------------------------------------------------------- collective ------ */
/* Data-parallel gradient exchange (no counterpart in the reference: new component,
 * SURVEY.md section 8(e)).  NCCL is dlopen()ed from `libnccl_path` (NULL = "libnccl.so.2"). */
int mt_comm_load(const char* libnccl_path);
int mt_comm_unique_id(char* out_bytes_128);
int mt_comm_init_rank(const char* id_bytes_128, int rank, int world);
int mt_comm_world(int* rank, int* world);
/* sum-allreduce `n` floats in place on the communication stream; ordered after all work
 * already queued on the compute stream.  mt_comm_wait() makes the compute stream wait.  */
int mt_allreduce_sum_f32(float* buf, size_t n);
/* all-reduce `grad`, then  param -= lr * grad  (reference: src/nn/optimizer.py:37-39), both on the communication
 * stream: the update of layer L runs under the backward GEMMs of the layers below it.  When `grad` is the dW that
 * the last mt_linear_bwd_f32 produced the exchange starts as soon as that GEMM is done (not after the dX GEMM queued
 * behind it); the update itself is ordered after everything queued on the compute stream at the time of the call.
 * Without a communicator (world 1) it is just the deferred update.  Readers of `param` must follow mt_comm_wait(). */
int mt_allreduce_sgd_f32(float* param, float* grad, size_t n, float lr);
/* Peer-memory exchange (NVLink P2P through CUDA IPC; no NCCL in the data path): every rank owns an arena that its peers
 * map; a finalised weight gradient is published, reduce-scattered + all-gathered in ONE pass of peer loads / stores with a
 * fixed rank-order sum (bit-identical replicas), and applied (param -= lr * sum) - three small kernels on the communication
 * stream, all waits bounded.  mt_p2p_create exports this rank's 64-byte IPC handle; mt_p2p_open takes all `world` handles
 * in rank order.  `key` names the tensor identically on every rank.  MT_ERR_UNSUPPORTED = use mt_allreduce_sgd_f32.
 * New component like the rest of this block (the reference has no distributed code).                                   */
int mt_p2p_create(size_t arena_bytes, char* handle_out_64);
int mt_p2p_open(int rank, int world, const char* handles_world_x_64);
int mt_p2p_active(int* active);
int mt_p2p_error(unsigned* code);
int mt_p2p_allreduce_sgd_f32(float* param, float* grad, size_t n, float lr, unsigned long long key);
int mt_p2p_destroy(void);
/* make the compute stream wait for the communication stream; also polls ncclCommGetAsyncError (a dead peer aborts
 * the communicator and returns MT_ERR_NCCL instead of hanging)                                                   */
int mt_comm_wait(void);
int mt_comm_check(void);
int mt_comm_destroy(void);

/* ---------------------------------------------------------------- timeline -------- */
/* A device-side timeline without a profiler: while enabled, every kernel launch is followed by a timed CUDA event on
 * its stream (stream_id 0 = compute, 1 = communication).  Entry i is (label, stream, ms since entry 0); on one
 * stream consecutive entries bracket one kernel.  tools/dp_timeline.py turns it into profiles/r02_dp_timeline_*.  */
int mt_trace_enable(int on);
int mt_trace_reset(void);
int mt_trace_count(int* n);
int mt_trace_get(int i, char* label, size_t label_cap, int* stream_id, float* ms_since_first);

#ifdef __cplusplus
}
#endif
#endif /* MICROTORCH_B200_H */

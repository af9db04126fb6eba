/*
 * b200norm.h — C ABI of the B200-native normalization / reduction kernel library.
 *
 * This is the drop-in boundary for the ONE hot path this repository implements:
 * the LayerNorm / RMSNorm / softmax forward+backward segments that nvFuser's
 * InnerPersistent / InnerOuterPersistent / OuterPersistent / Reduction schedulers
 * would otherwise lower to NVRTC-generated code (OuterPersistent: softmax and the
 * batch-norm-shaped normalization over dim 0, B200NORM_F_OUTER).
 *
 * Reference interfaces each entry point replaces (paths under the reference tree):
 *   - ExecutorDispatch::{makeExecutor,compile,isCompiled,run}
 *       csrc/runtime/executor_dispatch.cpp:18-156            (the executor seam)
 *   - the dlopen'd "run_kernel"/"workspace_size" pair of CutlassCompiledKernel
 *       csrc/runtime/cutlass_compiled_kernel.h:69-82, .cpp:346-362
 *       (the in-tree precedent for an executor that bypasses lowering/codegen/NVRTC)
 *   - KernelExecutor::run launch + workspace contract
 *       csrc/runtime/executor.cpp:1013-1365 (outputs/workspaces are caller-allocated,
 *       semaphores come from a zero-initialised pool :1160-1170)
 *
 * Contract (same as the precedent):
 *   - plain C, no C++ types, no exceptions cross this boundary; integer status
 *     codes, b200norm_last_error() for the message (thread-local).
 *   - the caller owns every buffer; the library never allocates, frees or
 *     synchronises; kernels are enqueued on the stream handed in.
 *   - the workspace must be at least b200norm_workspace_bytes(desc) bytes,
 *     256-byte aligned, and its first B200NORM_WS_HEADER_BYTES bytes must be
 *     ZERO before the first launch that uses it. After that the header belongs
 *     to the library (an arrival counter that is back to zero at the end of
 *     every launch and a barrier generation that only increments), so one
 *     zero-fill at allocation time is enough (this mirrors nvFuser's zeroed
 *     semaphore pool). A workspace must not be shared by launches that can run
 *     concurrently (use one per stream). The backward kernels that finish dweight / dbias behind an
 *     in-kernel grid barrier are serialised across the streams of one process by the library (a launch
 *     on another stream than the previous such launch first waits for that stream); across PROCESSES
 *     sharing a device (MPS) set B200NORM_COOP=1 to make them cooperative launches.
 *   - there is NO CPU fallback: on a machine without an sm_100 device the
 *     launch entry points fail with B200NORM_ERR_NO_DEVICE.
 */
#ifndef B200NORM_H_
#define B200NORM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200NORM_ABI_VERSION 1
#define B200NORM_WS_HEADER_BYTES 256
#define B200NORM_MAX_DIMS 4

/* Status codes. 0 = success; 1..99 library codes; >= 1000 = 1000 + cudaError_t. */
enum {
  B200NORM_OK = 0,
  B200NORM_ERR_UNSUPPORTED = 1, /* segment not handled: caller falls through  */
  B200NORM_ERR_BAD_ARG = 2,     /* malformed desc / tensor list               */
  B200NORM_ERR_NO_DEVICE = 3,   /* no sm_100 device: there is no CPU fallback */
  B200NORM_ERR_WORKSPACE = 4,   /* workspace missing / too small / misaligned */
  B200NORM_ERR_CUDA = 1000      /* + cudaError_t                              */
};

/* Segment kinds (what SchedulerType + pattern the matched segment is).
 * Math follows csrc/ops/normalization.cpp of the reference:
 *   softmax :169-189, softmax_backward :191-210, layer_norm :304-342,
 *   rms_norm :352-380, layer_norm_backward :382-436, rms_norm_backward :438-485,
 *   thunder_rms_norm_backward :487-535, log_softmax :212-231,
 *   log_softmax_backward :233-249; plain reductions scheduler/reduction.cpp. */
typedef enum {
  B200NORM_OP_LAYERNORM_FWD = 0,
  B200NORM_OP_LAYERNORM_BWD = 1,
  B200NORM_OP_RMSNORM_FWD = 2,
  B200NORM_OP_RMSNORM_BWD = 3,
  B200NORM_OP_SOFTMAX_FWD = 4,
  B200NORM_OP_SOFTMAX_BWD = 5,
  B200NORM_OP_LOGSOFTMAX_FWD = 6,
  B200NORM_OP_LOGSOFTMAX_BWD = 7,
  B200NORM_OP_REDUCE_INNER = 8, /* [M,N] -> [M]   sum / max / amax            */
  B200NORM_OP_REDUCE_OUTER = 9, /* [M,N] -> [N]   sum (the dgamma/dbeta shape) */
  B200NORM_OP_FINALIZE_PARTIALS = 10, /* fp32 [K,N] sums -> io dtype (after allreduce) */
  B200NORM_OP_PEER_COMBINE = 11, /* finish the oldest outstanding PEER_PUBLISH: K tensors [N] io dtype */
  B200NORM_OP_NLL_SOFTMAX_BWD = 12, /* backward of REDUCE_INNER | B200NORM_F_REDUCE_NLL (cross-entropy backward) */
  B200NORM_OP_COUNT_
} b200norm_op;

typedef enum {
  B200NORM_DT_F32 = 0,
  B200NORM_DT_BF16 = 1,
  B200NORM_DT_F16 = 2
} b200norm_dtype;

/* desc.flags */
enum {
  B200NORM_F_HAS_WEIGHT = 1u << 0,
  B200NORM_F_HAS_BIAS = 1u << 1,
  /* RMSNorm: the saved per-row statistic is rms = sqrt(mean(x^2)+eps) (the
   * Thunder/Python spelling, benchmarks/python/test_rmsnorm_fwd.py:26-45)
   * instead of invstd = rsqrt(mean(x^2)+eps) (C++ rms_norm, normalization.cpp:365-379). */
  B200NORM_F_STAT_IS_RMS = 1u << 2,
  /* RMSNorm backward keeps the "- sum(g)" centring term exactly as the C++
   * rms_norm_backward builder does (normalization.cpp:461-476). Without the flag
   * the textbook / thunder form is computed (:487-535). */
  B200NORM_F_RMS_BWD_CENTERED = 1u << 3,
  /* Backward: which gradients to produce (output_mask of the builders). */
  B200NORM_F_OUT_DX = 1u << 4,
  B200NORM_F_OUT_DW = 1u << 5,
  B200NORM_F_OUT_DB = 1u << 6,
  /* Backward, multi-GPU: write dgamma/dbeta as fp32 partial sums (no cast) so the
   * caller can allreduce them and cast afterwards with OP_FINALIZE_PARTIALS. */
  B200NORM_F_PARTIALS_F32 = 1u << 7,
  /* softmax backward: round the per-row sum through the io dtype before use, as
   * the Python benchmark definition does (benchmarks/python/test_softmax_bwd.py:49-60). */
  B200NORM_F_ROUND_ROWSUM = 1u << 8,
  /* reductions: 0 = sum, MAX = NaN-propagating amax, FMAX = NaN-ignoring
   * (ir/internal_nodes.cpp:1746-1775). Also selects softmax's max flavour. */
  B200NORM_F_REDUCE_MAX = 1u << 9,
  B200NORM_F_REDUCE_FMAX = 1u << 10,
  /* LayerNorm forward: mean / invstd outputs requested. */
  B200NORM_F_OUT_MEAN = 1u << 11,
  B200NORM_F_OUT_INVSTD = 1u << 12,
  /* LayerNorm / RMSNorm forward with a fused residual add — the prologue nvFuser's persistent
   * schedulers absorb for norm(x + residual) (benchmarks/python/test_rmsnorm_add_fwd.py,
   * torch_ops.py:76-83; test_dropout_{layernorm,rmsnorm}_fwd.py without the dropout):
   * ADD_RESIDUAL: a second [M,N] input follows x and the norm is taken of x + residual;
   * OUT_SUM: x + residual (io dtype) is a second [M,N] output right after y;
   * ROUND_SUM: the sum is rounded through the io dtype before it is normalised (what a
   * bf16 `hidden + residual` feeding the norm does); without it the fp32 sum is used. */
  B200NORM_F_ADD_RESIDUAL = 1u << 13,
  B200NORM_F_OUT_SUM = 1u << 14,
  B200NORM_F_ROUND_SUM = 1u << 15,
  /* LayerNorm / RMSNorm backward, rows sharded over the GPUs of one box: the kernel itself
   * all-reduces the fp32 dgamma / dbeta sums through the peer-memory exchange buffers registered
   * with b200norm_register_peers (NVLink / NVSwitch P2P loads and stores from the finishing CTAs)
   * and writes the cast result on every rank — no separate collective. Replaces the reference's
   * Communication(Allreduce) node (csrc/host_ir/lower_to_communication.cpp:266-284,
   * csrc/multidevice/post_communication.cpp:288-313). Every rank must issue the same sequence of
   * PEER_REDUCE launches; the launch fails with B200NORM_ERR_UNSUPPORTED when the shape cannot
   * take the fused path (the caller then all-reduces the PARTIALS_F32 sums with NCCL). */
  B200NORM_F_PEER_REDUCE = 1u << 16,
  /* The same exchange split in two so that ranks never wait for each other inside the backward:
   * the backward only PUBLISHES its fp32 sums (slot write + one flag store per peer from the last
   * finishing CTA; it produces no dweight / dbias tensors), and a later OP_PEER_COMBINE launch — one
   * CTA, typically a step later, when every peer's flag has long arrived — adds the ranks' slots in
   * rank order and casts. Publishes and combines pair up in order; at most two publishes may be
   * outstanding (two slots), the third waits in-kernel for the peers' "consumed" flags. */
  B200NORM_F_PEER_PUBLISH = 1u << 17,
  /* SOFTMAX_* / LOGSOFTMAX_* and LAYERNORM_FWD / LAYERNORM_BWD: the reduction runs over dim 0 (the ROWS) of the row-major [M,N]
   * matrix instead of dim 1 — the segments nvFuser's OuterPersistent scheduler takes
   * (scheduler/normalization_outer.cpp:451-683; benchmarks/python/test_softmax_{fwd,bwd}.py with
   * reduction_axis = 0, pytest mark outer_persistent). desc.rows = M is then the reduced extent and
   * desc.cols = N the iteration extent; tensors keep the same [M,N] layout and order as the inner
   * ops. Runs on the outer-persistent kernels (a column slab resident in the shared memory of a CTA
   * or of a cluster split along M) when the operands are 16-byte aligned, N * sizeof(T) % 16 == 0 and M is
   * small enough for a 16-CTA cluster; anything else (and B200NORM_VARIANT_GENERIC) runs on the shape-agnostic
   * three-kernel family (statistics / merge / apply, fp32 partials in the workspace), the split the reference
   * makes when it segments such fusions. A forced TMA / CLUSTER variant refuses instead of falling back.
   * LAYERNORM_FWD | OUTER is the batch-norm-shaped forward (training statistics, channels last:
   * benchmarks/python/normalization.py:13-95 viewed as [N*H*W, C]): mean / biased variance per COLUMN
   * over the M rows, y = (x - mean) * rsqrt(var + eps) * weight[col] + bias[col]; the mean / invstd
   * outputs are fp32 [N]. No fused residual add.
   * LAYERNORM_BWD | OUTER is its backward (benchmarks/python/normalization.py:98-190): in dy[M,N], x[M,N],
   * mean[N], invstd[N] (fp32), weight[N]?; out dx[M,N]?, dweight[N]?, dbias[N]? (io dtype) — column sums taken
   * inside the kernel (no partial-sum / peer flags). */
  B200NORM_F_OUTER = 1u << 18,
  /* RMSNORM_BWD: the projection term carries a factor 1/2 — thunder_rms_norm_backward as the C++ builder
   * writes it (csrc/ops/normalization.cpp:487-535: inv_rms_mul2 = 1 / (2 rms) scales x * sum(-g x / rms^2) / N),
   * used by tests/cpp/test_combined_inner_outer_reduction.cpp:1197:
   *   dx = g / rms - x * sum(g x) / (2 N rms^3)      (the textbook form has no 2)
   * dweight is unaffected. Parity is with the fusion as written, so the spelling is a segment kind. */
  B200NORM_F_RMS_BWD_HALF_PROJ = 1u << 19,
  /* LAYERNORM_FWD as a stand-alone statistics segment — ops.var_mean / ops.welford over the last axis
   * (csrc/ops/normalization.cpp:93-158, python/python_direct/ops.cpp:1824-1863; device side of the reference:
   * runtime/welford.cu:52-221): NO_Y: there is no y output (and no weight / bias); the statistics requested with
   * OUT_MEAN / OUT_INVSTD are the only outputs. STAT_VAR: the second statistic is the variance
   * M2 / max(N - correction, 0) with the integer correction in desc.reserved[0] (var_mean); STAT_M2: it is
   * M2 = sum((x - mean)^2) itself (welford's var_sum). Without either it stays invstd = rsqrt(M2 / N + eps). */
  B200NORM_F_NO_Y = 1u << 20,
  B200NORM_F_STAT_VAR = 1u << 21,
  B200NORM_F_STAT_M2 = 1u << 22,
  /* Fused dropout — the prologue of benchmarks/python/test_dropout_{layernorm,rmsnorm}_{fwd,bwd}.py:
   *     y = norm(residual + dropout(x, p)),  dropout(x) = x * mask / (1 - p),  mask = uniform(0,1) < 1 - p.
   * Forward (needs ADD_RESIDUAL): the library draws the mask from the reference's own Philox4x32-10 stream
   * (runtime/random_numbers.cu:8-88; element i of the [M,N] tensor takes word i mod 4 of philox(seed, i / 4, offset),
   * device_lower/pass/rng.cpp:95-127) with {seed, offset} read from a two-element int64 DEVICE tensor that follows
   * the other inputs, and returns the mask as the last output ([M,N] bytes, 1 = kept, contiguous rows).
   * Backward: residual[M,N] and mask[M,N] follow the other inputs; x_eff = residual + dropout(x) is rebuilt, and the
   * outputs are d x (= d x_eff * mask / (1 - p)), then d residual (= d x_eff), then dweight / dbias.
   * The keep probability 1 - p (the double scalar of the fusion) travels in desc.reserved[1..2] (low, high word). */
  B200NORM_F_DROPOUT = 1u << 23,
  /* REDUCE_INNER: the fused log-softmax + NLL of a cross-entropy loss over vocabulary-size rows (SURVEY 8f rank 2:
   * benchmarks/python/cross_entropy_loss.py; log_softmax csrc/ops/normalization.cpp:212-231 followed by the gather of
   * the target's entry): per row  lse = log(sum exp(x)),  loss = lse - x[target]  (0 where target == ignore_index,
   * desc.reserved[3]; at::nll_loss semantics). One read of the logits, two fp32 values per row out; the [M,N]
   * log-probabilities are never written. Backward: B200NORM_OP_NLL_SOFTMAX_BWD,
   *     dx[i,j] = (exp(x[i,j] - lse[i]) - [j == target[i]]) * g,   g = a device fp32 scalar (dloss / count),
   * zero rows where target == ignore_index. */
  B200NORM_F_REDUCE_NLL = 1u << 24,
  /* Instance norm (csrc/ops/normalization.cpp:790-950 instance_norm, :952-1060 instance_norm_backward; training
   * statistics, channels first): x[B, C, spatial...] viewed as [M = B*C rows, N = spatial elements] is a LayerNorm over
   * each row whose affine parameters belong to the row's CHANNEL: weight[C] / bias[C] are indexed by (row mod C),
   * C = desc.reserved[3]. LAYERNORM_FWD: same tensors as without the flag (weight / bias have C elements).
   * LAYERNORM_BWD: dx as usual; with OUT_DW / OUT_DB the outputs are the per-ROW sums sum(dy * x_hat) / sum(dy) as
   * fp32 [M] (dweight[c] / dbias[c] are their sums over the batch: one REDUCE_OUTER of the [B, C] view). */
  B200NORM_F_ROW_AFFINE = 1u << 25
};

/* desc.variant: 0 lets the library choose (the dispatch rule mirrors the accept
 * rules of scheduler/normalization_inner*.cpp); the others force a kernel family
 * and make b200norm_query fail if that family cannot run the shape. */
typedef enum {
  B200NORM_VARIANT_AUTO = 0,
  B200NORM_VARIANT_GENERIC = 1, /* any shape/stride/alignment; multi-pass LDG     */
  B200NORM_VARIANT_TMA = 2,     /* bulk-copy ring in smem, one CTA per row slice  */
  B200NORM_VARIANT_CLUSTER = 3  /* TMA ring + DSMEM reduction across a cluster    */
} b200norm_variant;

typedef struct {
  int32_t abi_version; /* B200NORM_ABI_VERSION */
  int32_t op;          /* b200norm_op          */
  int32_t io_dtype;    /* dtype of activations, gradients, weight, bias */
  int32_t variant;     /* b200norm_variant     */
  int64_t rows;        /* M: product of the outer (iteration) extents   */
  int64_t cols;        /* N: product of the normalized / reduced extents */
  double eps;          /* scalars are doubles in the fusion IR          */
  uint32_t flags;
  int32_t cluster_size; /* 0 = auto; else 1,2,4,8,16 */
  int32_t reserved[4];
} b200norm_desc;

/* plan.form */
enum {
  B200NORM_FORM_CACHED = 1,   /* the row (or g / x_hat) stays in registers between the passes   */
  B200NORM_FORM_PIPE = 2,     /* split-phase reductions pipelined across rows (mbarrier based)  */
  B200NORM_FORM_PSMEM = 4,    /* per-column parameters staged in shared memory as fp32          */
  B200NORM_FORM_FUSED_FINISH = 8 /* dweight / dbias finished inside the backward (grid barrier)   */
};

/* Tensor argument: same fields as CutlassCompiledKernel::TensorArg
 * (cutlass_compiled_kernel.h:69-74) with the arrays held inline. Strides are in
 * elements. The innermost stride must be 1 (or the extent 1). */
typedef struct {
  void* data;
  int32_t dtype; /* b200norm_dtype */
  int32_t ndim;
  int64_t sizes[B200NORM_MAX_DIMS];
  int64_t strides[B200NORM_MAX_DIMS];
} b200norm_tensor;

/* What the library decided for a desc on the current device: reported for
 * NVFUSER_DUMP=launch_param-style logging and for tests that assert which
 * kernel family ran (the reference tests assert the heuristic the same way,
 * tests/cpp/test_cluster.cpp:181-188). */
typedef struct {
  int32_t variant;       /* b200norm_variant actually chosen                 */
  int32_t cluster_size;  /* CTAs per row                                     */
  int32_t grid_x;        /* CTAs launched                                    */
  int32_t block_x;       /* threads per CTA                                  */
  int32_t threads_per_row; /* consumer threads cooperating on one row slice  */
  int32_t stages;        /* depth of the smem ring                           */
  int32_t tile_rows;     /* rows per bulk copy                               */
  int32_t smem_bytes;    /* dynamic shared memory per CTA                    */
  int32_t n_launches;    /* kernels enqueued per b200norm_launch             */
  int32_t form;          /* B200NORM_FORM_* bits: which flavour of the ring kernel */
  int32_t groups;        /* consumer groups per CTA (each with its own ring)  */
  int32_t chunks;        /* 16-byte chunks per thread (0 = looping form)      */
  uint64_t workspace_bytes;
  uint64_t algorithmic_bytes; /* the reference's *_iobytes for this desc      */
} b200norm_plan;

/*
 * Tensor order per op (absent optional tensors are simply left out):
 *   LAYERNORM_FWD  in : x[M,N], residual[M,N]?, weight[N]?, bias[N]?, rng[2] (DROPOUT)
 *                  out: y[M,N] (absent with NO_Y), sum[M,N]?, mean[M]?, invstd | var | M2 [M]?, mask[M,N] (DROPOUT)
 *   LAYERNORM_BWD  in : dy[M,N], x[M,N], mean[M], invstd[M], weight[N]?, residual[M,N] + mask[M,N] (DROPOUT)
 *                  out: dx[M,N]?, dresidual[M,N] (DROPOUT), dweight[N]?, dbias[N]?
 *       (ROW_AFFINE: weight / bias [C]; backward outputs dx?, rowsum(dy*x_hat)[M] fp32?, rowsum(dy)[M] fp32?)
 *   RMSNORM_FWD    in : x[M,N], residual[M,N]?, weight[N]?, rng[2] (DROPOUT)
 *                  out: y[M,N], sum[M,N]?, stat[M], mask[M,N] (DROPOUT)
 *   RMSNORM_BWD    in : dy[M,N], x[M,N], stat[M], weight[N]?, residual[M,N] + mask[M,N] (DROPOUT)
 *                  out: dx[M,N]?, dresidual[M,N] (DROPOUT), dweight[N]?
 *   SOFTMAX_FWD / LOGSOFTMAX_FWD   in : x[M,N]               out: y[M,N]
 *   SOFTMAX_BWD / LOGSOFTMAX_BWD   in : dy[M,N], y[M,N]      out: dx[M,N]
 *       (with B200NORM_F_OUTER the four softmax ops reduce over M instead of N; same tensors)
 *   REDUCE_INNER   in : x[M,N]                               out: r[M]   (io dtype)
 *       with REDUCE_NLL: in : x[M,N], target[M] (int64)        out: loss[M] fp32, lse[M] fp32
 *   NLL_SOFTMAX_BWD in: x[M,N], target[M] (int64), lse[M] fp32, g[1] fp32   out: dx[M,N]
 *   REDUCE_OUTER   in : x[M,N]                               out: r[N]   (io dtype, or fp32 with PARTIALS_F32)
 *   FINALIZE_PARTIALS in: p[K,N] fp32 (rows=K)               out: K tensors [N] io dtype
 *   PEER_COMBINE   in : -  (rows=K, cols=N)                  out: K tensors [N] io dtype
 * Per-row statistics are always fp32; [M] and [M,1] are both accepted.
 */

/* ABI version of the loaded library. */
int b200norm_abi_version(void);

/* 0 if this desc can run on the current device; fills *plan when non-NULL.
 * Pure host logic when no device is present (plans for a nominal B200: 148 SMs,
 * 227 KB smem), so callers can test dispatch rules without a GPU.
 * Replaces: SchedulerEntry::canSchedule / computeHeuristics for the matched
 * segment (scheduler/registry.cpp:192-227, normalization_inner.cpp:73-231). */
int b200norm_query(const b200norm_desc* desc, b200norm_plan* plan);

/* Workspace size in bytes (0 on error). Replaces workspace_size()
 * (cutlass_compiled_kernel.cpp:145-152). */
size_t b200norm_workspace_bytes(const b200norm_desc* desc);

/* Enqueue the segment. `stream` is a cudaStream_t passed as void*.
 * Replaces run_kernel(args, workspace, stream) (cutlass_compiled_kernel.cpp:113-157)
 * / KernelExecutor::run's cuLaunchKernelEx (executor.cpp:1270-1346). */
int b200norm_launch(const b200norm_desc* desc,
                    const b200norm_tensor* inputs, int n_inputs,
                    const b200norm_tensor* outputs, int n_outputs,
                    void* workspace, size_t workspace_bytes,
                    void* stream);

/* Number of kernels this library has enqueued so far in this process
 * (monotonic; used by the bench's gpu_launches counter). */
uint64_t b200norm_launch_count(void);

/* Message for the last non-zero status on this thread ("" if none). */
const char* b200norm_last_error(void);

/* Name of the kernel family for a variant id ("generic", "tma", "cluster"). */
const char* b200norm_variant_name(int variant);

/* Device facts the planner uses; returns 0 and fills nominal B200 values with
 * *has_device = 0 when no CUDA device is visible. */
int b200norm_device_info(int* has_device, int* sm_count, int* smem_per_block_optin,
                         int* cc_major, int* cc_minor);

/* Registers, for the CURRENT device, this process's view of the peer-memory exchange buffers of
 * B200NORM_F_PEER_REDUCE / PEER_PUBLISH: `exchange_ptrs[q]` is a device pointer, valid on the current device,
 * to rank q's buffer (`exchange_ptrs[rank]` is this rank's own); every buffer holds `exchange_bytes` bytes, is
 * 256-byte aligned and ZERO when registered. Size it with b200norm_peer_exchange_bytes(world, K*N) for the
 * largest K*N (K = 2 LayerNorm, 1 RMSNorm) that will use it. Layout: 256 bytes of flags | two fp32 [K,N] slots
 * of the publish / combine flavour | the fused flavour's receive area: two parities x world x K*N (value, epoch)
 * 8-byte words, written by the PEERS with remote stores. world <= 8 (one box); world = 0 unregisters. The
 * buffers stay owned by the caller (e.g. torch symmetric memory) and must outlive every launch that uses them.
 * In-kernel waits for a peer time out after B200NORM_PEER_TIMEOUT_S seconds (environment, read here; default
 * 600, 0 = never) with a trap, i.e. a sticky launch failure. Reference counterpart: the IPC / symmetric-memory
 * handles of csrc/multidevice/ipc_handle.cpp, symmetric_tensor.cpp. */
int b200norm_register_peers(int world, int rank, void* const* exchange_ptrs, size_t exchange_bytes);

/* Bytes of one exchange buffer for `world` ranks and segments of up to `max_floats` = K*N column sums (0 on bad arguments). */
size_t b200norm_peer_exchange_bytes(int world, int64_t max_floats);

/* Tuning aid (no reference counterpart): while `device_buf` is non-NULL every ring kernel
 * launched by this process stores per-CTA %globaltimer marks into it (12 x uint64 per CTA:
 * entry, setup done, first tile landed, row loop done, partial row written, grid barrier
 * passed, exit, SM id, column sums finished, peers' sums collected, 2 spare). The buffer must
 * hold 12 * grid_x uint64. NULL switches it off. */
void b200norm_debug_set_timeline(void* device_buf);

#ifdef __cplusplus
}
#endif
#endif /* B200NORM_H_ */

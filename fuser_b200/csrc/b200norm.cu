// b200norm.cu — C ABI, planner (dispatch heuristics) and kernel launches.
//
// The planner plays the role of nvFuser's canSchedule + computeHeuristics for
// the matched segments (reference: scheduler/normalization_inner.cpp:73-231,
// normalization_inner_non_tma.cpp:367-613,1144-1173,
// normalization_inner_outer.cpp:246-535): it decides the kernel family, the
// cluster size (row split), threads per row, ring depth and grid. The rules are
// re-derived for the ring design used here, not transcribed from the reference.
#include "../../include/b200norm.h"

#include <cuda_runtime.h>
#include <unistd.h>  // environ

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include <unordered_map>

#include "kernels_generic.cuh"
#include "outer.cuh"
#include "ring_dispatch.cuh"

using namespace b200norm;

namespace {

thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
std::atomic<unsigned long long*> g_timeline{nullptr};

// peer-memory exchange of the row-sharded backward (b200norm_register_peers)
struct PeerState {
  int world = 0;
  int rank = 0;
  void* ptrs[kMaxPeers] = {};
  size_t bytes = 0;
  uint32_t cap_floats = 0;    // K*N capacity of one slot
  uint32_t ll_off = 0;        // byte offset of the LL region (behind the header and the two publish slots)
  uint32_t ll_parity_bytes = 0;  // bytes of one parity of the LL region: world * cap_floats * 8
  uint32_t ll_epoch = 0;      // fused (PEER_REDUCE) launches so far; identical on every rank (same launch sequence)
  uint32_t epoch = 0;         // publishes so far (PEER_PUBLISH)
  uint32_t combined = 0;      // publishes whose sums have been combined by OP_PEER_COMBINE
  uint64_t timeout_ns = 0;    // in-kernel wait for a peer: 0 = for ever
};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

struct DeviceInfo {
  int has_device = 0;
  int sm_count = 148;
  int smem_optin = 232448;  // 227 KB
  int cc_major = 10;
  int cc_minor = 0;
};

// Everything the library remembers is per DEVICE: function attributes and occupancy answers belong to a
// device context, one process may drive several devices (NormExecutor launches under a device guard), and the
// peer exchange buffers are registered for the device that was current at registration.
struct DeviceCtx {
  int ordinal = -1;
  DeviceInfo info;
  std::mutex mu;  // guards the maps below
  std::unordered_map<const void*, std::pair<int, int>> attr_done;  // kernel -> (max smem set, nonportable cluster set)
  std::unordered_map<uint64_t, int> cluster_cache;                 // (kernel, block, smem, C) -> resident clusters
  std::unordered_map<const void*, int> bad_modes;                  // kernel -> bit0: fused finish refused, bit1: PDL refused
  // Kernels that spin on an in-kernel grid barrier need their whole grid resident. On ONE stream that holds by
  // construction (see launch_typed); two such kernels on two streams could each get a part of the SMs and wait
  // for ever. They are therefore serialised across streams: a barrier kernel launched on another stream than the
  // previous one first waits for everything enqueued on that stream so far.
  std::mutex barrier_mu;
  cudaStream_t barrier_stream = nullptr;
  bool barrier_seen = false;
  cudaEvent_t barrier_event = nullptr;
  std::mutex peer_mu;  // held across a PEER launch: epochs advance in launch order and only on success
  PeerState peers;
};

constexpr int kMaxDevices = 64;

DeviceCtx& device_ctx() {
  static DeviceCtx nominal;  // no device: plans for a nominal B200
  static DeviceCtx table[kMaxDevices];
  static std::once_flag once[kMaxDevices];
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    (void)cudaGetLastError();
    return nominal;
  }
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) {
    (void)cudaGetLastError();
    return nominal;
  }
  std::call_once(once[dev], [dev] {
    DeviceCtx& c = table[dev];
    c.ordinal = dev;
    c.info.has_device = 1;
    cudaDeviceGetAttribute(&c.info.sm_count, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&c.info.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&c.info.cc_major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&c.info.cc_minor, cudaDevAttrComputeCapabilityMinor, dev);
  });
  return table[dev];
}

const DeviceInfo& device_info() { return device_ctx().info; }

// Tuning overrides (like NVFUSER_ENABLE / NVFUSER_DISABLE options), all unset in production:
// B200NORM_TPR, _GROUPS, _DEPTH (stages per group), _TILE_ROWS, _CLUSTER, _BWD_REREAD, _PIPE, _NO_PIPE, _NO_PSMEM,
// _NO_FUSED_FINISH, _NO_PDL, _COOP, _STATIC, _POOL, _DEBUG, _PEER_TIMEOUT_S. They are read with ONE scan of the
// environment per query / launch (ten getenv calls were ten scans).
struct Knobs {
  int tpr = 0, groups = 0, depth = 0, tile_rows = 0, cluster = 0, bwd_reread = 0, pipe = 0, no_pipe = 0, no_psmem = 0;
  int no_fused_finish = 0, no_pdl = 0, coop = 0, static_sched = 0, pool = 0, debug = 0, peer_timeout_s = -1;
  int ln_pipe = 0, no_ln_pipe = 0, no_guard = 0, no_lean = 0, drop_ch4 = 0, no_tile2 = 0;
  bool any = false;
  uint64_t hash = 0;  // of the planner-relevant values (plan cache key)
};

Knobs read_knobs() {
  Knobs k;
  static const struct { const char* name; int Knobs::*field; } table[] = {
      {"TPR", &Knobs::tpr}, {"GROUPS", &Knobs::groups}, {"DEPTH", &Knobs::depth}, {"TILE_ROWS", &Knobs::tile_rows},
      {"CLUSTER", &Knobs::cluster}, {"BWD_REREAD", &Knobs::bwd_reread}, {"PIPE", &Knobs::pipe}, {"NO_PIPE", &Knobs::no_pipe},
      {"NO_PSMEM", &Knobs::no_psmem}, {"NO_FUSED_FINISH", &Knobs::no_fused_finish}, {"NO_PDL", &Knobs::no_pdl},
      {"COOP", &Knobs::coop}, {"STATIC", &Knobs::static_sched}, {"POOL", &Knobs::pool}, {"DEBUG", &Knobs::debug},
      {"PEER_TIMEOUT_S", &Knobs::peer_timeout_s}, {"LN_PIPE", &Knobs::ln_pipe}, {"NO_LN_PIPE", &Knobs::no_ln_pipe},
      {"NO_GUARD", &Knobs::no_guard}, {"NO_LEAN", &Knobs::no_lean}, {"DROP_CH4", &Knobs::drop_ch4}, {"NO_TILE2", &Knobs::no_tile2}};
  if (::environ == nullptr) return k;
  for (char** e = ::environ; *e != nullptr; ++e) {
    if (strncmp(*e, "B200NORM_", 9) != 0) continue;
    const char* name = *e + 9;
    const char* eq = strchr(name, '=');
    if (eq == nullptr || eq[1] == 0) continue;
    for (const auto& t : table) {
      const size_t len = strlen(t.name);
      if ((size_t)(eq - name) == len && strncmp(name, t.name, len) == 0) {
        k.*(t.field) = atoi(eq + 1);
        k.any = true;
        break;
      }
    }
  }
  if (k.any) {
    const int vals[] = {k.tpr, k.groups, k.depth, k.tile_rows, k.cluster, k.bwd_reread, k.pipe, k.no_pipe, k.no_psmem,
                        k.ln_pipe, k.no_ln_pipe, k.no_lean, k.drop_ch4, k.no_tile2};
    uint64_t h = 1469598103934665603ull;
    for (int v : vals) h = (h ^ (uint64_t)(uint32_t)v) * 1099511628211ull;
    k.hash = h | 1ull;
  }
  return k;
}

int elem_size(int dt) { return dt == B200NORM_DT_F32 ? 4 : 2; }

bool is_row_op(int op) {  // ops with a TMA-ring kernel family
  return (op >= B200NORM_OP_LAYERNORM_FWD && op <= B200NORM_OP_LOGSOFTMAX_BWD) || op == B200NORM_OP_REDUCE_INNER ||
         op == B200NORM_OP_REDUCE_OUTER;
}
// ops that keep per-thread column accumulators and finish fp32 column sums (dgamma / dbeta, plain column sums)
bool has_column_sums(const b200norm_desc& d) {
  if (d.op == B200NORM_OP_REDUCE_OUTER) return true;
  if (d.flags & B200NORM_F_ROW_AFFINE) return false;  // instance norm: per-row sums, written by the row kernel itself
  return (d.op == B200NORM_OP_LAYERNORM_BWD || d.op == B200NORM_OP_RMSNORM_BWD) &&
         (d.flags & (B200NORM_F_OUT_DW | B200NORM_F_OUT_DB)) != 0;
}
bool is_norm_bwd(int op) { return op == B200NORM_OP_LAYERNORM_BWD || op == B200NORM_OP_RMSNORM_BWD; }
bool is_softmax_op(int op) { return op >= B200NORM_OP_SOFTMAX_FWD && op <= B200NORM_OP_LOGSOFTMAX_BWD; }
bool is_outer(const b200norm_desc& d) { return (d.flags & B200NORM_F_OUTER) != 0; }
bool is_add_fwd(const b200norm_desc& d) {
  return (d.op == B200NORM_OP_LAYERNORM_FWD || d.op == B200NORM_OP_RMSNORM_FWD) && (d.flags & B200NORM_F_ADD_RESIDUAL);
}
int n_streams(int op) {
  switch (op) {
    case B200NORM_OP_LAYERNORM_BWD:
    case B200NORM_OP_RMSNORM_BWD:
    case B200NORM_OP_SOFTMAX_BWD:
    case B200NORM_OP_LOGSOFTMAX_BWD:
      return 2;
    default:
      return 1;
  }
}
int partial_k(int op) { return op == B200NORM_OP_LAYERNORM_BWD ? 2 : 1; }

uint64_t algorithmic_bytes(const b200norm_desc& d) {
  // The reference's own *_iobytes formulas (benchmarks/python/test_*_{fwd,bwd}.py).
  const uint64_t s = elem_size(d.io_dtype), M = d.rows, N = d.cols;
  switch (d.op) {
    case B200NORM_OP_LAYERNORM_FWD:
    case B200NORM_OP_RMSNORM_FWD: {
      // + the residual read and the sum write of the fused-add variants (every tensor the fusion
      // reads or writes, as benchmarks/python/core.py:compute_total_iobytes counts them)
      const uint64_t extra = ((d.flags & B200NORM_F_ADD_RESIDUAL) ? s * M * N : 0) +
                             ((d.flags & B200NORM_F_ADD_RESIDUAL) && (d.flags & B200NORM_F_OUT_SUM) ? s * M * N : 0) +
                             ((d.flags & B200NORM_F_DROPOUT) ? M * N : 0);  // + the bool mask (dropout_*_fwd_iobytes)
      if (d.op == B200NORM_OP_LAYERNORM_FWD && (d.flags & B200NORM_F_NO_Y)) {
        // var_mean / welford: the input once, one fp32 value per requested statistic and row
        const uint64_t nstat = ((d.flags & B200NORM_F_OUT_MEAN) ? 1 : 0) + ((d.flags & B200NORM_F_OUT_INVSTD) ? 1 : 0);
        return s * M * N + 4 * nstat * M;
      }
      if (d.op == B200NORM_OP_LAYERNORM_FWD && (d.flags & B200NORM_F_OUTER)) return s * 2 * (M * N + N) + 4 * 2 * N;
      if (d.op == B200NORM_OP_LAYERNORM_FWD) return s * 2 * (M * N + N) + 4 * 2 * M + extra;
      return s * (2 * M * N + N) + 4 * M + extra;
    }
    case B200NORM_OP_LAYERNORM_BWD:
      if (d.flags & B200NORM_F_OUTER) return s * 3 * (M * N + N) + 4 * 2 * N;
      // dropout: + residual in, d residual out, the mask (dropout_layernorm_bwd_iobytes)
      return s * 3 * (M * N + N) + 4 * 2 * M + ((d.flags & B200NORM_F_DROPOUT) ? s * 2 * M * N + M * N : 0);
    case B200NORM_OP_RMSNORM_BWD:
      return s * (3 * M * N + 2 * N) + 4 * M + ((d.flags & B200NORM_F_DROPOUT) ? s * 2 * M * N + M * N : 0);
    case B200NORM_OP_SOFTMAX_FWD:
    case B200NORM_OP_LOGSOFTMAX_FWD: return s * 2 * M * N;
    case B200NORM_OP_SOFTMAX_BWD:
    case B200NORM_OP_LOGSOFTMAX_BWD: return s * 3 * M * N;
    case B200NORM_OP_REDUCE_INNER:
      if (d.flags & B200NORM_F_REDUCE_NLL) return s * M * N + 8 * M + 2 * 4 * M;  // logits, int64 targets; loss + lse
      return s * (M * N + M);
    case B200NORM_OP_NLL_SOFTMAX_BWD: return s * 2 * M * N + 8 * M + 4 * M + 4;
    case B200NORM_OP_REDUCE_OUTER: return s * (M * N + N);
    case B200NORM_OP_FINALIZE_PARTIALS: return 4 * M * N + s * M * N;
    default: return 0;
  }
}

// Everything the launch needs, derived from (desc, device, alignment).
struct Plan {
  int variant = B200NORM_VARIANT_GENERIC;
  int cluster = 1;
  int grid = 1;
  int block = kGenericThreads;
  int tpr = 0;
  int groups = 0;
  int stages = 0;
  int depth = 0;
  int tile_rows = 1;
  int contiguous = 0;
  int ch = 0;
  int nchunks = 0;
  bool cached = false;
  bool psmem = false;      // LayerNorm fwd: weight / bias staged in shared memory as fp32
  bool pipe = false;       // backward, re-read form: split-phase reductions pipelined across rows
  bool lean = false;       // RMSNorm backward without the centring term: kernels specialised for it
  bool stats_reduce = false;  // LAYERNORM_FWD | NO_Y (var_mean / welford): runs on the inner-reduction kernel
  int nparts = 0;          // partial rows produced by the main kernel
  int rows_per_split = 0;  // generic colsum
  int n_launches = 1;
  uint32_t slice_bytes = 0, stream_bytes = 0, stage_bytes = 0;
  uint32_t off_red = 0, off_cls = 0, off_param = 0, off_ring = 0, off_tidx = 0, off_pipe = 0;
  uint32_t smem = 0;
  uint64_t ws_bytes = 0;
};

uint32_t align_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

// ---------------------------------------------------------------------------
// kernel tables
// ---------------------------------------------------------------------------
// pipe: the split-phase pipelined flavours (norm backward re-read form, RMSNorm / LayerNorm forward)
#define B200_RING_KERNEL_FOR(T, SUF)                                                                         \
  template <>                                                                                                \
  RingKernel ring_kernel_for<T>(int op, int tpr, int ch, bool cached, bool psmem, bool add, bool pipe,       \
                                bool red, bool lean, bool drop, bool rowaff) {                             \
    if (rowaff && op == B200NORM_OP_LAYERNORM_BWD) return pipe ? nullptr : ring_bwd_rowaff_kernel_##SUF(tpr, ch, cached); \
    if (drop && is_norm_bwd(op)) return cached ? ring_bwd_drop_kernel_##SUF(op, tpr, ch) : nullptr;          \
    if (drop) return (add && !pipe && !psmem) ? ring_drop_kernel_##SUF(op, tpr, ch) : nullptr;               \
    if (op == B200NORM_OP_REDUCE_INNER && pipe) return ring_sm_pipe_kernel_##SUF(op, tpr, ch);               \
    if (op == B200NORM_OP_REDUCE_INNER || red) return ring_reduce_inner_kernel_##SUF(tpr, ch);               \
    if (op == B200NORM_OP_REDUCE_OUTER) return ring_colsum_kernel_##SUF(tpr, ch);                            \
    if (pipe && op == B200NORM_OP_LAYERNORM_FWD) return ring_ln_pipe_kernel_##SUF(tpr, ch, add);             \
    if (add) return ring_add_kernel_##SUF(op, tpr, ch, psmem);                                               \
    if (is_norm_bwd(op) && !cached && pipe) return ring_bwd_pipe_kernel_##SUF(op, tpr, ch, lean);            \
    if (op == B200NORM_OP_RMSNORM_FWD && pipe) return ring_fwd_pipe_kernel_##SUF(op, tpr, ch);               \
    if ((op == B200NORM_OP_SOFTMAX_FWD || op == B200NORM_OP_LOGSOFTMAX_FWD) && pipe)                         \
      return ring_sm_pipe_kernel_##SUF(op, tpr, ch);                                                         \
    return is_norm_bwd(op) ? ring_bwd_kernel_##SUF(op, tpr, ch, cached, !(lean && !cached))                    \
                           : ring_fwd_kernel_##SUF(op, tpr, ch, psmem);                                    \
  }
template <typename T>
RingKernel ring_kernel_for(int op, int tpr, int ch, bool cached, bool psmem, bool add, bool pipe, bool red, bool lean,
                           bool drop, bool rowaff);
B200_RING_KERNEL_FOR(float, f32)
B200_RING_KERNEL_FOR(__nv_bfloat16, bf16)
B200_RING_KERNEL_FOR(__half, f16)
#undef B200_RING_KERNEL_FOR

// cudaFuncSetAttribute is paid once per (device, kernel, smem size, cluster class), not per launch.
int prepare_ring_kernel(DeviceCtx& ctx, RingKernel k, int smem, int cluster) {
  std::lock_guard<std::mutex> lock(ctx.mu);
  auto& st = ctx.attr_done[(const void*)k];
  if (smem > st.first) {
    cudaError_t err = cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (err != cudaSuccess) return fail(B200NORM_ERR_CUDA + (int)err, "smem attr: %s", cudaGetErrorString(err));
    st.first = smem;
  }
  if (cluster > 8 && !st.second) {
    cudaError_t err = cudaFuncSetAttribute((const void*)k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (err != cudaSuccess) return fail(B200NORM_ERR_CUDA + (int)err, "cluster attr: %s", cudaGetErrorString(err));
    st.second = 1;
  }
  return B200NORM_OK;
}

RingKernel ring_kernel_any(int dt, int op, int tpr, int ch, bool cached, bool psmem, bool add, bool pipe, bool red,
                           bool lean, bool drop, bool rowaff = false) {
  switch (dt) {
    case B200NORM_DT_F32: return ring_kernel_for<float>(op, tpr, ch, cached, psmem, add, pipe, red, lean, drop, rowaff);
    case B200NORM_DT_BF16: return ring_kernel_for<__nv_bfloat16>(op, tpr, ch, cached, psmem, add, pipe, red, lean, drop, rowaff);
    case B200NORM_DT_F16: return ring_kernel_for<__half>(op, tpr, ch, cached, psmem, add, pipe, red, lean, drop, rowaff);
    default: return nullptr;
  }
}

// How many clusters of this shape can be resident at once. Clusters never span GPCs, so
// the answer is NOT sm_count / C (on B200 a GPC holds 16-20 SMs: e.g. only 72 two-CTA
// clusters fit 148 SMs). The ring kernels are persistent, so launching more clusters than
// fit would serialise a second wave. The reference asks the same question before a
// cluster launch (csrc/runtime/executor.cpp:1329-1340).
int max_active_clusters(DeviceCtx& ctx, RingKernel k, int block, int smem, int C) {
  const DeviceInfo& dev = ctx.info;
  if (C <= 1) return dev.sm_count;
  if (!dev.has_device || k == nullptr) {
    return 8 * (18 / C);  // nominal B200: 8 GPCs, >= 18 usable SMs each
  }
  const uint64_t key = (reinterpret_cast<uint64_t>((const void*)k) << 8) ^ ((uint64_t)block << 40) ^
                       ((uint64_t)smem << 4) ^ (uint64_t)C;
  {
    std::lock_guard<std::mutex> lock(ctx.mu);
    auto it = ctx.cluster_cache.find(key);
    if (it != ctx.cluster_cache.end()) return it->second;
  }
  int n = 0;
  if (prepare_ring_kernel(ctx, k, smem, C) == B200NORM_OK) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(dev.sm_count / C * C));
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = (size_t)smem;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = (unsigned)C;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    if (cudaOccupancyMaxActiveClusters(&n, (const void*)k, &cfg) != cudaSuccess) {
      (void)cudaGetLastError();
      n = 0;
    }
  }
  if (n <= 0) n = 8 * (18 / C) > 0 ? 8 * (18 / C) : 1;
  std::lock_guard<std::mutex> lock(ctx.mu);
  ctx.cluster_cache[key] = n;
  return n;
}

constexpr int kConsumerThreads = 512;
constexpr int kFwdWideCH = 6;      // ... for kernels whose only per-chunk state is x (TPR >= 256)
constexpr int kSmPipeMaxCH = 8;    // ... packed chunks per thread the pipelined cluster softmax keeps across its first pass
constexpr int kFwdMaxCH = 4;       // chunks per thread kept in registers by the forward-type kernels
constexpr int kBwdCacheCH = 2;     // backward, g / xc cached in registers
constexpr int kBwdRereadCH_RMS = 4;  // backward re-read variant: column accumulators = CH * VEC (x2 for LayerNorm)
constexpr int kBwdRereadCH_LN = 3;
constexpr int kMaxStages = 32;
const int kTprMenu[] = {32, 64, 128, 256, 512};

int generic_nsplit(int64_t rows) {
  int64_t n = rows / 64;
  if (n < 1) n = 1;
  if (n > 256) n = 256;
  return (int)n;
}

// Try to build a ring plan with cluster size C. Returns false if infeasible.
bool try_ring_plan(const b200norm_desc& d, DeviceCtx& ctx, const Knobs& kn, int C, bool rows_contiguous, Plan* out) {
  const DeviceInfo& dev = ctx.info;
  const int es = elem_size(d.io_dtype);
  const int vec = 16 / es;
  const int64_t N = d.cols;
  if (N % C != 0) return false;
  const int64_t slice_cols = N / C;
  if ((slice_cols * es) % 16 != 0) return false;
  // A slice of up to 100 KB leaves room for two ring stages. Beyond that — vocabulary-size rows, e.g. softmax over
  // N = 1048576 bf16 = 128 KB per CTA of a 16-CTA cluster, the largest size of tests/cpp/test_cluster.cpp:195-204 —
  // a one-stream forward-type op may still run with ONE stage per CTA on the widest cluster (the next row's load
  // only starts when the last pass over the current row is done: no overlap, but TMA + DSMEM instead of the scalar
  // family); two-stream ops would need 2 x 128 KB and stay with the shape-agnostic family.
  const bool single_stage_ok = C == 16 && !is_norm_bwd(d.op) && !is_add_fwd(d) && n_streams(d.op) == 1 &&
                               d.op != B200NORM_OP_REDUCE_OUTER;
  if (slice_cols * es > (int64_t)100 * 1024 && !(single_stage_ok && slice_cols * es <= (int64_t)200 * 1024)) return false;
  const int nchunks = (int)(slice_cols / vec);
  const bool add = is_add_fwd(d);
  // fused dropout: forward on the register-resident ADD kernels; backward on the cached form with two more streams
  // per stage (residual slice + mask bytes, whose slices must be whole 16-byte chunks)
  const bool dropout = (d.flags & B200NORM_F_DROPOUT) != 0;
  const bool drop_bwd = dropout && is_norm_bwd(d.op);
  const int ns = add ? 2 : n_streams(d.op);
  const bool bwd = is_norm_bwd(d.op);
  const bool ln_bwd = d.op == B200NORM_OP_LAYERNORM_BWD;
  const bool colsum = d.op == B200NORM_OP_REDUCE_OUTER;  // per-thread column accumulators, no per-row reduction
  // inner reductions without an output pass: sum / max / fmax and the stand-alone statistics (var_mean, welford)
  const bool inner_red = d.op == B200NORM_OP_REDUCE_INNER ||
                         (d.op == B200NORM_OP_LAYERNORM_FWD && (d.flags & B200NORM_F_NO_Y) != 0);

  // Softmax / log-softmax forward over a cluster: the pipelined one-exchange kernel (kernels_softmax_pipe.cuh) — the
  // row stays in its stage, so slices of any width work as long as two stages fit; one 512-thread group per CTA.
  // Measured against the register-resident kernel in DESIGN 4c. B200NORM_NO_PIPE=1 switches it off.
  // The cross-entropy forward (REDUCE_INNER | F_REDUCE_NLL) is the same kernel without its second pass.
  const bool nll_red = d.op == B200NORM_OP_REDUCE_INNER && (d.flags & B200NORM_F_REDUCE_NLL) != 0;
  // Also WITHOUT a cluster for rows of more than 48 KB (N = 32000 ... 37500 at 16 bits — the 32 K vocabularies): they fit
  // neither the registers of the register kernel nor two consumer groups, and one CTA per row keeps all 148 SMs busy.
  // ... and for the softmax rows the register kernel walks with ONE 512-thread group, where it measured faster
  // (bf16 [16384,24576] 296 -> 262 us = 94 %, [16384,20480] 261 -> 241 us, [16384,16384] 229 -> 224 us, fp32 [16384,8192]
  // 197 -> 183 us; not fp32 N = 12288: 128 vs 141 us; rows of 24 KB and less have several groups per CTA and win
  // by a wide margin: [32768,12288] 247 vs 388 us). The cross-entropy mode only beyond 48 KB: up to there the
  // multi-group reduction kernel is ahead ([16384,24576] 179 vs 208 us). profiles/r02t_kbench_pipe_no_cluster.txt
  const bool sm_op = d.op == B200NORM_OP_SOFTMAX_FWD || d.op == B200NORM_OP_LOGSOFTMAX_FWD;
  const bool pipe_c1 = kn.pipe >= 0 && (nchunks > 512 * kFwdWideCH || (sm_op && (es == 2 ? nchunks >= 2048 : nchunks == 2048)));
  const bool sm_pipe = (sm_op || nll_red) && (C > 1 || pipe_c1 || kn.pipe > 0) &&
                       kn.no_pipe == 0 && (kn.tpr == 0 || kn.tpr == 512) && nchunks >= 1024 &&
                       slice_cols * es <= (int64_t)100 * 1024;

  // Threads per row slice and chunks per thread. Forward-type kernels: the smallest group
  // that keeps the slice in registers at <= kFwdMaxCH chunks per thread (fewer, fatter
  // threads = less reduction overhead and more groups per CTA). Backward: <= kBwdCacheCH
  // with g / xc cached, else the re-read form with more accumulator columns per thread.
  // Slices wider than 512*4 chunks use the looping form (forward-type only).
  Plan p;
  int tpr = 0, ch = 0;
  bool cached = true;
  const int forced_tpr = kn.tpr;
  auto pick = [&](int max_ch) {
    for (int t : kTprMenu) {
      if (forced_tpr != 0 && t != forced_tpr) continue;
      if (nchunks <= t * max_ch) {
        tpr = t;
        ch = (nchunks + t - 1) / t;
        return true;
      }
    }
    return false;
  };
  if (sm_pipe) {
    tpr = 512;
    ch = (nchunks + 511) / 512;
    // a short ragged last step (<= 256 chunks past the full ones) is walked in 4-byte pieces by all threads instead
    // of one more chunk by a few: the kernel takes CH = floor (kernels_softmax_pipe.cuh, "balanced tail";
    // B200NORM_NO_TILE2=1 switches this off too, for A/B runs)
    if (nchunks % 512 != 0 && nchunks % 512 <= 256 && nchunks / 512 >= 2 && kn.no_tile2 == 0) ch = nchunks / 512;
    if (ch > kSmPipeMaxCH) ch = 0;  // looping form
    cached = ch > 0;
  } else if (colsum) {
    if (!pick(kFwdMaxCH)) return false;  // CH * VEC accumulators per thread; wider rows are split over a cluster
  } else if (inner_red && nchunks > 128 * kFwdMaxCH && forced_tpr == 0) {
    // Inner reductions keep no per-chunk state, so any number of chunks per thread works (looping form): prefer
    // SEVERAL consumer groups per CTA — a lone 512-thread group exposes every row's reduction latency
    // ([32768,12288] bf16: 204 us with one group) — as long as every group still gets two ring stages.
    const int64_t budget = (int64_t)dev.smem_optin - 8192;
    const int64_t row_bytes = (int64_t)align_up((uint32_t)(slice_cols * es), 128);
    tpr = 512;
    for (int t : {128, 256, 512}) {
      if ((int64_t)(kConsumerThreads / t) * 2 * row_bytes <= budget) {
        tpr = t;
        break;
      }
    }
    ch = nchunks <= tpr * kFwdMaxCH ? (nchunks + tpr - 1) / tpr : 0;
    cached = ch > 0;
  } else if (bwd) {
    const bool force_reread = kn.bwd_reread != 0;
    if (force_reread || !pick(kBwdCacheCH)) {
      cached = false;
      if (!pick(ln_bwd ? kBwdRereadCH_LN : kBwdRereadCH_RMS)) return false;  // split the row further
      if (ch < 3) ch = 3;  // only CH 3 / 4 are instantiated for the re-read form
      if (ch == 4 && ln_bwd) return false;
    }
  } else if ((d.op == B200NORM_OP_RMSNORM_FWD && !add) || d.op == B200NORM_OP_SOFTMAX_FWD) {
    // x is the only per-chunk register state: up to 6 chunks per thread from TPR 256 on
    // (N = 12288 bf16 -> two 256-thread groups per CTA instead of one 512-thread group)
    if (!pick(kFwdMaxCH)) {
      tpr = 0;
      for (int t : {256, 512}) {
        if (forced_tpr != 0 && t != forced_tpr) continue;
        if (nchunks <= t * kFwdWideCH) {
          tpr = t;
          ch = (nchunks + t - 1) / t;
          break;
        }
      }
      if (tpr == 0) {
        if (forced_tpr != 0 && forced_tpr != 512) return false;
        tpr = 512;
        ch = 0;
        cached = false;
      }
    } else if (tpr == 512 && forced_tpr == 0 && nchunks <= 256 * kFwdWideCH &&
               !(d.op == B200NORM_OP_RMSNORM_FWD && es == 2 && (nchunks + 255) / 256 == 6 &&
                 kn.no_pipe == 0)) {
      // (16-bit RMSNorm forward at six chunks per thread — N = 12288 — keeps the single 512-thread
      // group: its pipelined kernel beats two 256-thread groups there, 252 vs 275 us; at five chunks
      // (N = 10240: 109 vs 123 us) and at N = 16384 (162 vs 178 us) the register-resident kernels win)
      tpr = 256;
      ch = (nchunks + 255) / 256;
    }
  } else if (dropout && !bwd && kn.drop_ch4 == 0 && (nchunks <= 256 || kn.drop_ch4 < 0) && pick(2)) {
    // fused dropout forward, short rows: at most two chunks per thread. Measured bf16 [32768,2048] 137 vs 164 us; from
    // N = 4096 on the fatter threads win although their Philox draw spills ([16384,4096] 136 vs 150 us, [16384,8192]
    // 266 vs 308 us, fp32 [16384,4096] 170 vs 218 us; B200NORM_DROP_CH4 = 1 / -1 forces four / two chunks)
  } else if (!pick(kFwdMaxCH)) {
    if (d.op == B200NORM_OP_LAYERNORM_FWD && !add && forced_tpr == 256 && nchunks <= 256 * kFwdWideCH &&
        nchunks % 256 == 0 && kn.no_ln_pipe == 0) {
      // tuning: two 256-thread groups of the pipelined one-round kernel (rows of 5 / 6 chunks per thread)
      tpr = 256;
      ch = nchunks / 256;
    } else if (d.op == B200NORM_OP_LAYERNORM_FWD && !add && !dropout && !inner_red && (forced_tpr == 0 || forced_tpr == 512) &&
               nchunks > 512 * kFwdMaxCH && nchunks <= 512 * kFwdWideCH && nchunks % 512 == 0 && kn.no_ln_pipe == 0 &&
               kn.ln_pipe >= 0) {
      // rows of five / six chunks per thread at 512 threads (N = 20480 / 24576 at 16 bits, 10240 / 12288 at fp32) on the
      // pipelined one-round kernel without a cluster: the alternative was a cluster of two with 20-24 KB slices, whose
      // rows are shorter than the DSMEM round trip ([16384,24576] bf16: 405 us = 61 %). B200NORM_LN_PIPE=-1 restores it.
      tpr = 512;
      ch = nchunks / 512;
    } else {
      if (forced_tpr != 0 && forced_tpr != 512) return false;
      tpr = 512;
      ch = 0;
      cached = false;
    }
  }
  int groups = kConsumerThreads / tpr;
  // At most 8 groups (named-barrier ids) — except for 32-thread rows of the forward-type kernels, whose groups
  // synchronise with __syncwarp: sixteen one-warp groups, each with a two-stage ring of smaller tiles, keep 512 threads
  // per SM busy where eight left half the warps out. Measured bf16 (8 -> 16 groups): RMSNorm [32768,784] 22.0 -> 17.3 us
  // (71 -> 91 %), LayerNorm [32768,784] 32.2 -> 24.1 us, [65536,1024] 57.2 -> 51.1 us, RMSNorm [65536,1024] 52.6 -> 49.8 us
  // (profiles/r02n_kbench_short_rows.txt). B200NORM_GROUPS=8 restores the old plans.
  // Two-stream kernels (softmax backward, the fused-add forwards) go the other way except for the shortest rows:
  // [65536,1024] softmax bwd 68 -> 116 us, rms(x + residual) 91 -> 130 us, but [262144,256] 144 -> 112 / 186 -> 130 us.
  const bool g16 = tpr == 32 && !bwd && !colsum && !inner_red && C == 1 && rows_contiguous &&
                   (ns == 1 || slice_cols * es <= 512);
  if (groups > 8 && !g16) groups = 8;
  if (int gmax = kn.groups) {
    if (gmax >= 1 && gmax < groups) groups = gmax;
  }

  p.variant = C > 1 ? B200NORM_VARIANT_CLUSTER : B200NORM_VARIANT_TMA;
  p.cluster = C;
  p.tpr = tpr;
  p.ch = ch;
  p.cached = cached;
  p.nchunks = nchunks;
  p.slice_bytes = (uint32_t)(slice_cols * es);

  // rows per tile: aim for >= 8 KB bulk copies when rows are short and contiguous
  int tile_rows = 1;
  if (C == 1 && rows_contiguous && p.slice_bytes < 8192) {
    tile_rows = (int)((groups > 8 ? 4096u / (uint32_t)ns : 8192u) / p.slice_bytes);  // 16 groups: ~4 KB stages
    if (tile_rows > 32) tile_rows = 32;
    if (tile_rows < 1) tile_rows = 1;
    // do not starve the grid: keep at least a few tiles per consumer group per SM
    while (tile_rows > 1 && ceil_div64(d.rows, tile_rows) < (int64_t)dev.sm_count * groups * 2) tile_rows /= 2;
  }
  // 16-bit norm backward with rows of 8 KB and more: two rows per tile when the ring still gets
  // three stages per group. The per-tile work (barrier wait, refill, pointer rebuild) is a large
  // share of these instruction-bound kernels: measured +12-15 % at N = 4096 and 8192 bf16
  // (C2 backward 80 -> 70 us); fp32 rows (memory bound already) and N = 12288 (only two stages
  // would fit) stay at one row.
  if (bwd && es == 2 && C == 1 && rows_contiguous && tile_rows == 1) {
    const uint32_t stage2 = align_up((uint32_t)(ns + (drop_bwd ? 1 : 0)) * 2u * p.slice_bytes + (drop_bwd ? 2u * (uint32_t)slice_cols : 0u), 128);
    const int64_t budget = (int64_t)dev.smem_optin - 4096;
    if (budget / stage2 / groups >= 3 && ceil_div64(d.rows, 2) >= (int64_t)dev.sm_count * groups * 2) tile_rows = 2;
  }
  // Rows of 8-12 KB, two per tile: inner reductions (nothing is held per row, the per-tile cost dominates: bf16
  // [65536,4096] 123.7 -> 90.8 us, [16384,4096] 32.0 -> 26.0 us, [16384,6144] 37.4 -> 33.8 us; rows of 16 KB and more
  // measured equal or slower) and the 16-bit LayerNorm forward with the fused residual add at 8 KB (the register
  // kernel on two-row tiles beats the pipelined one there: [16384,4096] 100.7 -> 90.0 us)
  if (C == 1 && rows_contiguous && tile_rows == 1 && kn.no_tile2 == 0 &&
      ((inner_red && p.slice_bytes >= 8192 && p.slice_bytes <= 12288) ||
       (add && d.op == B200NORM_OP_LAYERNORM_FWD && !dropout && es == 2 && p.slice_bytes == 8192)) &&
      ceil_div64(d.rows, 2) >= (int64_t)dev.sm_count * groups * 2)
    tile_rows = 2;
  if (int tr = kn.tile_rows) {
    if (C == 1 && rows_contiguous && tr >= 1 && tr <= 64) tile_rows = tr;
  }
  p.tile_rows = tile_rows;
  p.contiguous = (C == 1 && rows_contiguous) ? 1 : 0;
  if (tile_rows == 1) p.contiguous = 1;  // a one-row tile is always one copy per stream
  p.stream_bytes = (uint32_t)tile_rows * p.slice_bytes;
  p.stage_bytes = align_up((uint32_t)ns * p.stream_bytes + (drop_bwd ? p.stream_bytes + (uint32_t)tile_rows * (uint32_t)slice_cols : 0u), 128);
  if (drop_bwd && slice_cols % 16 != 0) return false;

  // LayerNorm forward keeps fp32 copies of weight and bias in shared memory when the ring
  // still gets >= 3 stages per group next to them.
  // (16-bit io only: fp32 parameters need no unpacking, and staging them costs a tiny kernel like
  // C1 16 % — 16.4 vs 14.1 us; not beyond 8192 columns: at N = 12288 the 96 KB of staged parameters
  // and their four LDS per chunk cost 10 % — 378 vs 342 us)
  // LayerNorm forward with one reduction round per row (Welford pairs merged with Chan's formula), pipelined
  // across rows (kernels_ln_pipe.cuh), wherever it can run: equal element counts per thread, one row per tile,
  // at least two warps per row. Measured
  // against the two-round register kernel (bf16): [32768,12288] 343 -> 305 us, [16384,8192] 115 -> 92 us,
  // [16384,4096] 62 -> 53 us, [16384,16384] 273 -> 182 us; fp32 [2048,4096] 15.4 -> 14.5 us.
  // B200NORM_NO_LN_PIPE=1 switches it off.
  // (fused dropout: the register-resident forward kernels only — the row must not be rebuilt a second time)
  if (dropout && (!cached || ch == 0)) return false;
  const bool ln_pipe = d.op == B200NORM_OP_LAYERNORM_FWD && !inner_red && ch > 0 && nchunks == tpr * ch && tile_rows == 1 &&
                       tpr >= 64 && kn.no_ln_pipe == 0 && !dropout;
  bool want_psmem = d.op == B200NORM_OP_LAYERNORM_FWD && ch > 0 && es == 2 && slice_cols <= 8192 &&
                    kn.no_psmem == 0 && !ln_pipe && !inner_red && !dropout && !(d.flags & B200NORM_F_ROW_AFFINE);
  const uint32_t param_bytes = align_up(2u * (uint32_t)slice_cols * 4u, 128);

  // Ring depth D (stages per group); the ring is group-major, stage (g*D + s).
  for (;;) {
    uint32_t off = 8u * (kMaxStages + groups);
    p.off_tidx = off;  // dynamic schedule: stage -> tile index
    off += 4u * kSchedTableInts + 16u;  // + the CTA's "groups that saw the empty pool" counter
    p.off_pipe = off;  // pipelined backward: red[2G] | free[G*D] | clusterB[G] mbarriers
    off += 8u * (uint32_t)(kMaxStages + 3 * groups);
    off = align_up(off, 2 * kRedSetBytes);  // slot-set pairs are toggled with an XOR
    p.off_red = off;
    off += (uint32_t)groups * 2 * kRedSetBytes;
    p.off_cls = off;
    off += (uint32_t)groups * 2 * kClsSetBytes;
    off = align_up(off, 128);
    p.off_param = off;
    if (want_psmem) off += param_bytes;
    p.off_ring = off;
    const int64_t budget = (int64_t)dev.smem_optin - 1024 - off;
    int max_stages = (int)(budget / p.stage_bytes);
    if (want_psmem && max_stages / groups < 3) {
      want_psmem = false;  // not enough room: parameters stay packed in registers
      continue;
    }
    if (max_stages > kMaxStages) max_stages = kMaxStages;
    int depth = max_stages / groups;  // stages per group
    if (depth < 2 && groups > 1) {
      groups = groups / 2;
      continue;
    }
    if (depth < 2 && !(depth == 1 && single_stage_ok && p.slice_bytes > 100u * 1024u)) return false;
    if (depth > 4) depth = 4;  // deeper rings only scatter the DRAM streams (measured: C3 fwd 87% at 6, 93% at 3)
    if (int dd = kn.depth) {
      if (dd >= 1 && dd <= max_stages / groups) depth = dd;
    }
    p.stages = depth * groups;
    p.depth = depth;
    break;
  }
  p.groups = groups;
  p.psmem = want_psmem;
  // re-read backward: the pipelined kernel (one row per tile, at least two warps per group) where it
  // was measured faster — rows split over a cluster (the DSMEM exchange overlaps the next row's
  // first pass: C5 335 -> 320 us) and fp32 rows ([16384,8192] 273 -> 248 us); the 16-bit
  // single-CTA case is bound by instruction issue, not by the reduction latency, and loses 4 %
  // (C4 420 -> 440 us), so it keeps the plain kernel. B200NORM_PIPE=1 / B200NORM_NO_PIPE=1 force it.
  p.pipe = bwd && !cached && !dropout && !(d.flags & B200NORM_F_ROW_AFFINE) && tile_rows == 1 && tpr >= 64 && kn.no_pipe == 0 &&
           (C > 1 || es == 4 || kn.pipe != 0);
  // RMSNorm forward over a cluster: the same pipeline (the row is read twice from its stage instead
  // of being held in registers, so the next row's sum of squares overlaps the DSMEM exchange)
  // ... and for rows wide enough to need one 512-thread group per CTA; with several groups per
  // CTA the groups already overlap each other and the second smem pass only costs (N = 8192: 85 vs 95 us).
  // (16-bit only: the fp32 single-group kernel is at the roofline already, 163 vs 182 us at N = 8192)
  // ... and at six chunks per thread of ONE 512-thread group (N = 24576 at 16 bits: 292 -> 268 us = 92 %; five chunks
  // — N = 20480: 214 vs 231 us — and fp32 N = 12288 — 246 vs 269 us — stay on the register kernel)
  if (ln_pipe || sm_pipe) p.pipe = true;
  if (sm_pipe && p.depth < 2) return false;  // pass 1 of the next row needs its own stage
  p.stats_reduce = inner_red && d.op == B200NORM_OP_LAYERNORM_FWD;
  // RMSNorm backward without the centring term: the re-read instantiations that neither accumulate sum(g) nor
  // reduce it ([32768,12288] bf16 382 -> 368 us, [16384,8192] 137 -> 132 us; the CACHED form measured slower
  // specialised — C2 72 -> 75 us — and keeps the general instantiation), and the instruction-trimmed pipe kernel
  // (kernels_bwd_lean.cuh: C5 332 -> 290 us)
  p.lean = d.op == B200NORM_OP_RMSNORM_BWD && !(d.flags & B200NORM_F_RMS_BWD_CENTERED) && kn.no_lean == 0;
  if (d.op == B200NORM_OP_RMSNORM_FWD && !add &&
      (C > 1 || (tpr == 512 && es == 2 && ch == 3 && (nchunks + 255) / 256 == 6) || (tpr == 512 && es == 2 && ch == 6) ||
       kn.pipe != 0) && ch > 0 &&
      tile_rows == 1 && tpr >= 64 && (ch <= 4 || tpr >= 256) && kn.no_pipe == 0) {
    p.pipe = true;
  }
  p.block = groups * tpr;
  const uint32_t ring_bytes = (uint32_t)p.stages * p.stage_bytes;
  if (bwd || colsum) {
    const uint32_t acc_bytes = (uint32_t)partial_k(d.op) * (uint32_t)slice_cols * 4;
    if (acc_bytes > ring_bytes) return false;
  }
  p.smem = p.off_ring + ring_bytes;

  // only the plans the planner emits are instantiated (ring_inst_*.cuh): a combination forced through the tuning
  // knobs may have no kernel — then this cluster size is simply not feasible
  if (ring_kernel_any(d.io_dtype, d.op, tpr, ch, cached, p.psmem, add, p.pipe, p.stats_reduce, p.lean, dropout, (d.flags & B200NORM_F_ROW_AFFINE) != 0) == nullptr)
    return false;
  const int64_t ntiles = ceil_div64(d.rows, tile_rows);
  int64_t nclusters = max_active_clusters(ctx, ring_kernel_any(d.io_dtype, d.op, tpr, ch, cached, p.psmem, add, p.pipe, p.stats_reduce, p.lean, dropout, (d.flags & B200NORM_F_ROW_AFFINE) != 0), p.block,
                                          (int)p.smem, C);
  if (nclusters > ntiles) nclusters = ntiles;
  if (nclusters < 1) nclusters = 1;
  p.grid = (int)(nclusters * C);
  p.nparts = (int)nclusters;
  p.n_launches = 1;
  if (has_column_sums(d)) p.n_launches = 2;
  *out = p;
  return true;
}

int make_plan_uncached(const b200norm_desc& d, DeviceCtx& ctx, const Knobs& kn, bool aligned, bool rows_contiguous, Plan* out) {
  const DeviceInfo& dev = ctx.info;
  Plan p;
  const bool bwd = is_norm_bwd(d.op);
  const bool want_cols = has_column_sums(d);

  // workspace: large enough for whichever family ends up running
  uint64_t ws = B200NORM_WS_HEADER_BYTES;
  if (want_cols) {
    int maxparts = generic_nsplit(d.rows);
    if (dev.sm_count > maxparts) maxparts = dev.sm_count;
    ws += (uint64_t)maxparts * partial_k(d.op) * d.cols * 4;
  }

  bool ring_ok = false;
  if (is_row_op(d.op) && aligned && d.variant != B200NORM_VARIANT_GENERIC && d.rows > 0 && d.cols > 0) {
    const int forced_c = d.cluster_size != 0 ? d.cluster_size : kn.cluster;
    static const int cs[] = {1, 2, 4, 8, 16};
    Plan best;
    for (int C : cs) {
      if (forced_c > 0 && C != forced_c) continue;
      if (forced_c == 0 && d.variant == B200NORM_VARIANT_TMA && C != 1) continue;
      if (forced_c == 0 && d.variant == B200NORM_VARIANT_CLUSTER && C == 1) continue;
      Plan cand;
      if (!try_ring_plan(d, ctx, kn, C, rows_contiguous, &cand)) continue;
      if (!ring_ok) {
        best = cand;
        ring_ok = true;
      }
      // prefer the smallest cluster that gives a register-resident slice (forward-type) or
      // per-thread column accumulators (backward) with >= 3 stages per group: a cluster costs
      // a DSMEM exchange per row and leaves SMs idle (clusters never span GPCs)
      // (an inner reduction holds a stage only while it is being summed: two stages per group are enough)
      const bool inner_red = d.op == B200NORM_OP_REDUCE_INNER ||
                             (d.op == B200NORM_OP_LAYERNORM_FWD && (d.flags & B200NORM_F_NO_Y) != 0);
      // (the pipelined cluster softmax reads its rows from the stage: the smallest cluster with three stages —
      // two for the cross-entropy flavour, which releases a stage after its only pass — whatever the slice width)
      const bool sm_pipe = cand.pipe && (d.op == B200NORM_OP_SOFTMAX_FWD || d.op == B200NORM_OP_LOGSOFTMAX_FWD ||
                                         d.op == B200NORM_OP_REDUCE_INNER);
      const bool good = sm_pipe ? cand.depth >= (d.op == B200NORM_OP_REDUCE_INNER ? 2 : 3)
                        : inner_red ? cand.depth >= 2
                                                         : (cand.cached || is_norm_bwd(d.op)) && cand.depth >= (cand.groups > 8 ? 2 : 3);  // sixteen one-warp groups run two-stage rings
      if (good || forced_c > 0) {
        best = cand;
        break;
      }
      if ((cand.cached && !best.cached) || (cand.cached == best.cached && cand.stages > best.stages)) best = cand;
    }
    if (ring_ok) p = best;
  }
  if (!ring_ok) {
    if (d.variant == B200NORM_VARIANT_TMA || d.variant == B200NORM_VARIANT_CLUSTER) {
      return fail(B200NORM_ERR_UNSUPPORTED, "requested kernel family cannot run rows=%lld cols=%lld",
                  (long long)d.rows, (long long)d.cols);
    }
    p = Plan();
    p.variant = B200NORM_VARIANT_GENERIC;
    p.block = kGenericThreads;
    int64_t g = d.rows;
    const int64_t cap = (int64_t)dev.sm_count * 8;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    p.grid = (int)g;
    p.n_launches = 1;
    if (want_cols && bwd) {
      p.nparts = generic_nsplit(d.rows);
      p.rows_per_split = (int)ceil_div64(d.rows, p.nparts);
      p.n_launches = 3;
    }
    if (d.op == B200NORM_OP_REDUCE_OUTER) {
      p.nparts = generic_nsplit(d.rows);
      p.rows_per_split = (int)ceil_div64(d.rows, p.nparts);
      p.n_launches = 2;
      p.block = 128;
    }
    if (d.op == B200NORM_OP_NLL_SOFTMAX_BWD) {  // pointwise streaming kernel: (row, block of chunks) work items
      p.block = kCeThreads;
      const int64_t items = d.rows * ceil_div64(ceil_div64(d.cols, 16 / elem_size(d.io_dtype)), kCeThreads * kCeU);
      const int64_t capg = (int64_t)dev.sm_count * 8;
      p.grid = (int)(items < capg ? (items < 1 ? 1 : items) : capg);
    }
    if (d.op == B200NORM_OP_FINALIZE_PARTIALS) {
      p.block = 256;
      p.grid = (int)ceil_div64(d.cols, 32);
    }
  }
  p.ws_bytes = ws;
  *out = p;
  return B200NORM_OK;
}

// The plan of a (desc, alignment, device, knobs) tuple never changes: the executor asks twice per launch
// (workspace size, then the launch itself), so the answer is cached (the role of
// KernelExecutor::executor_entry_lookup_, csrc/runtime/executor.cpp:1055-1077).
struct PlanKey {
  int32_t op, dt, variant, cluster, dev;
  uint32_t flags;
  int64_t rows, cols;
  uint64_t knobs;
  int32_t aligned, contiguous;
  bool operator==(const PlanKey& o) const { return memcmp(this, &o, sizeof(PlanKey)) == 0; }
};
struct PlanKeyHash {
  size_t operator()(const PlanKey& k) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(PlanKey) / 8; ++i) h = (h ^ w[i]) * 1099511628211ull;
    return (size_t)h;
  }
};
static_assert(sizeof(PlanKey) % 8 == 0, "PlanKey is hashed as 64-bit words");

int make_plan(const b200norm_desc& d, DeviceCtx& ctx, const Knobs& kn, bool aligned, bool rows_contiguous, Plan* out) {
  static std::mutex mu;
  static std::unordered_map<PlanKey, Plan, PlanKeyHash> cache;
  PlanKey key;
  memset(&key, 0, sizeof(key));
  key.op = d.op;
  key.dt = d.io_dtype;
  key.variant = d.variant;
  key.cluster = d.cluster_size;
  key.dev = ctx.ordinal;
  key.flags = d.flags;
  key.rows = d.rows;
  key.cols = d.cols;
  key.knobs = kn.hash;
  key.aligned = aligned ? 1 : 0;
  key.contiguous = rows_contiguous ? 1 : 0;
  {
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return B200NORM_OK;
    }
  }
  const int rc = make_plan_uncached(d, ctx, kn, aligned, rows_contiguous, out);
  if (rc == B200NORM_OK) {
    std::lock_guard<std::mutex> lock(mu);
    if (cache.size() > 4096) cache.clear();
    cache[key] = *out;
  }
  return rc;
}

OuterArgs outer_args_from(const b200norm_desc& d) {
  const DeviceInfo& dev = device_info();
  OuterArgs a;
  a.bwd = (d.op == B200NORM_OP_SOFTMAX_BWD || d.op == B200NORM_OP_LOGSOFTMAX_BWD || d.op == B200NORM_OP_LAYERNORM_BWD) ? 1 : 0;
  a.log = (d.op == B200NORM_OP_LOGSOFTMAX_FWD || d.op == B200NORM_OP_LOGSOFTMAX_BWD) ? 1 : 0;
  a.dtype = d.io_dtype;
  a.M = d.rows;
  a.N = d.cols;
  a.norm = (d.op == B200NORM_OP_LAYERNORM_FWD || d.op == B200NORM_OP_LAYERNORM_BWD) ? 1 : 0;
  a.eps = d.eps;
  a.round_sum = (d.flags & B200NORM_F_ROUND_ROWSUM) ? 1 : 0;
  a.cluster_forced = d.cluster_size != 0 ? d.cluster_size : read_knobs().cluster;
  a.force_generic = d.variant == B200NORM_VARIANT_GENERIC ? 1 : 0;
  a.force_persistent = (d.variant == B200NORM_VARIANT_TMA || d.variant == B200NORM_VARIANT_CLUSTER) ? 1 : 0;
  a.has_device = dev.has_device;
  a.sm_count = dev.sm_count;
  a.smem_optin = dev.smem_optin;
  return a;
}

double dropout_keep_prob(const b200norm_desc& d) {
  const uint64_t bits = (uint64_t)(uint32_t)d.reserved[1] | ((uint64_t)(uint32_t)d.reserved[2] << 32);
  double keep;
  memcpy(&keep, &bits, sizeof(keep));
  return keep;
}

// uniformf of the reference (runtime/random_numbers.cu:66-88) before the 1.0 -> 0 wrap: fl(fl(w) * 2^-32 + 2^-33)
float dropout_raw_uniform(uint32_t w) {
  const float scale = 2.3283064365386963e-10f;  // 2^-32
  return (float)w * scale + 0.5f * scale;
}

// keep <=> uniformf(w) < keep (float against the fusion's double)  <=>  w <= t1 || w >= t2: the raw value is
// monotone in w, and from t2 on it rounds to 1.0, which uniformf maps to 0 (< keep for every keep > 0).
void dropout_thresholds(double keep, uint32_t* t1, uint32_t* t2) {
  uint64_t a = 0, b = 1ull << 32;  // first w whose raw value is 1.0f
  while (a < b) {
    const uint64_t mid = (a + b) / 2;
    if (dropout_raw_uniform((uint32_t)mid) >= 1.0f) b = mid; else a = mid + 1;
  }
  const uint64_t wrap = a;
  *t2 = (uint32_t)(wrap < (1ull << 32) ? wrap : 0xffffffffull);
  a = 0;  // first w below the wrap with raw >= keep (w = 0 gives 2^-33 < keep for every accepted keep)
  b = wrap;
  while (a < b) {
    const uint64_t mid = (a + b) / 2;
    if ((double)dropout_raw_uniform((uint32_t)mid) < keep) a = mid + 1; else b = mid;
  }
  *t1 = (uint32_t)(a > 0 ? a - 1 : 0);
}

int check_desc(const b200norm_desc* d) {
  if (d == nullptr) return fail(B200NORM_ERR_BAD_ARG, "null desc");
  if (d->abi_version != B200NORM_ABI_VERSION)
    return fail(B200NORM_ERR_BAD_ARG, "abi_version %d != %d", d->abi_version, B200NORM_ABI_VERSION);
  if (d->op < 0 || d->op >= B200NORM_OP_COUNT_) return fail(B200NORM_ERR_UNSUPPORTED, "unknown op %d", d->op);
  if (d->io_dtype < 0 || d->io_dtype > B200NORM_DT_F16)
    return fail(B200NORM_ERR_UNSUPPORTED, "unsupported io dtype %d", d->io_dtype);
  if (d->flags >> 26) return fail(B200NORM_ERR_BAD_ARG, "unknown flag bits 0x%x", d->flags & ~((1u << 26) - 1u));
  if (d->flags & B200NORM_F_ROW_AFFINE) {
    if (d->op != B200NORM_OP_LAYERNORM_FWD && d->op != B200NORM_OP_LAYERNORM_BWD)
      return fail(B200NORM_ERR_UNSUPPORTED, "B200NORM_F_ROW_AFFINE (instance norm) applies to the LayerNorm ops");
    if (d->flags & (B200NORM_F_OUTER | B200NORM_F_NO_Y | B200NORM_F_ADD_RESIDUAL | B200NORM_F_DROPOUT | B200NORM_F_PARTIALS_F32 |
                    B200NORM_F_PEER_REDUCE | B200NORM_F_PEER_PUBLISH))
      return fail(B200NORM_ERR_UNSUPPORTED, "B200NORM_F_ROW_AFFINE excludes OUTER, NO_Y, ADD_RESIDUAL, DROPOUT, PARTIALS_F32 and the peer flags");
    if (d->reserved[3] < 1 || (d->rows % d->reserved[3]) != 0)
      return fail(B200NORM_ERR_BAD_ARG, "instance norm: rows (%lld) must be a multiple of the channel count (%d)", (long long)d->rows, d->reserved[3]);
  }
  if (d->flags & B200NORM_F_REDUCE_NLL) {
    if (d->op != B200NORM_OP_REDUCE_INNER)
      return fail(B200NORM_ERR_UNSUPPORTED, "B200NORM_F_REDUCE_NLL is a flavour of REDUCE_INNER");
    if (d->flags & (B200NORM_F_REDUCE_MAX | B200NORM_F_REDUCE_FMAX | B200NORM_F_OUTER))
      return fail(B200NORM_ERR_BAD_ARG, "B200NORM_F_REDUCE_NLL excludes REDUCE_MAX / REDUCE_FMAX / OUTER");
  }
  if (d->flags & B200NORM_F_DROPOUT) {
    const bool fwd = d->op == B200NORM_OP_LAYERNORM_FWD || d->op == B200NORM_OP_RMSNORM_FWD;
    const bool bwd = d->op == B200NORM_OP_LAYERNORM_BWD || d->op == B200NORM_OP_RMSNORM_BWD;
    if (!fwd && !bwd) return fail(B200NORM_ERR_UNSUPPORTED, "B200NORM_F_DROPOUT applies to the LayerNorm / RMSNorm ops");
    if (fwd && !(d->flags & B200NORM_F_ADD_RESIDUAL))
      return fail(B200NORM_ERR_UNSUPPORTED, "dropout forward is norm(residual + dropout(x)): ADD_RESIDUAL is required");
    if (d->flags & (B200NORM_F_OUTER | B200NORM_F_NO_Y | B200NORM_F_ROUND_SUM | B200NORM_F_PARTIALS_F32 |
                    B200NORM_F_PEER_REDUCE | B200NORM_F_PEER_PUBLISH))
      return fail(B200NORM_ERR_UNSUPPORTED, "B200NORM_F_DROPOUT excludes OUTER, NO_Y, ROUND_SUM, PARTIALS_F32 and the peer flags");
    const double keep = dropout_keep_prob(*d);
    if (!(keep > 1e-6) || keep > 1.0) return fail(B200NORM_ERR_BAD_ARG, "dropout keep probability %g outside (0, 1]", keep);
  }
  if (d->rows < 0 || d->cols < 0) return fail(B200NORM_ERR_BAD_ARG, "negative extent");
  if (d->cols > 0x7fffffffLL) return fail(B200NORM_ERR_UNSUPPORTED, "cols %lld exceeds int32", (long long)d->cols);
  if (d->cluster_size != 0 && d->cluster_size != 1 && d->cluster_size != 2 && d->cluster_size != 4 &&
      d->cluster_size != 8 && d->cluster_size != 16)
    return fail(B200NORM_ERR_BAD_ARG, "cluster_size must be 0,1,2,4,8,16");
  if (d->variant < 0 || d->variant > B200NORM_VARIANT_CLUSTER) return fail(B200NORM_ERR_BAD_ARG, "bad variant");
  if ((d->flags & B200NORM_F_OUTER) && !is_softmax_op(d->op) && d->op != B200NORM_OP_LAYERNORM_FWD &&
      d->op != B200NORM_OP_LAYERNORM_BWD)
    return fail(B200NORM_ERR_UNSUPPORTED, "B200NORM_F_OUTER applies to softmax / log-softmax and to the LayerNorm-type ops");
  if ((d->flags & B200NORM_F_OUTER) &&
      (d->flags & (B200NORM_F_PARTIALS_F32 | B200NORM_F_PEER_REDUCE | B200NORM_F_PEER_PUBLISH)))
    return fail(B200NORM_ERR_UNSUPPORTED, "the outer normalization backward produces dweight / dbias directly (no partial sums)");
  if ((d->flags & B200NORM_F_OUTER) && (d->flags & B200NORM_F_ADD_RESIDUAL))
    return fail(B200NORM_ERR_UNSUPPORTED, "the outer normalization has no fused residual add");
  if (d->flags & (B200NORM_F_NO_Y | B200NORM_F_STAT_VAR | B200NORM_F_STAT_M2)) {
    if (d->op != B200NORM_OP_LAYERNORM_FWD || (d->flags & B200NORM_F_OUTER))
      return fail(B200NORM_ERR_UNSUPPORTED, "NO_Y / STAT_VAR / STAT_M2 apply to LAYERNORM_FWD over the last axis");
    if ((d->flags & B200NORM_F_STAT_VAR) && (d->flags & B200NORM_F_STAT_M2))
      return fail(B200NORM_ERR_BAD_ARG, "STAT_VAR and STAT_M2 exclude each other");
    if ((d->flags & B200NORM_F_NO_Y) &&
        (d->flags & (B200NORM_F_HAS_WEIGHT | B200NORM_F_HAS_BIAS | B200NORM_F_ADD_RESIDUAL)))
      return fail(B200NORM_ERR_BAD_ARG, "a statistics-only segment has no weight, bias or residual");
    if ((d->flags & B200NORM_F_NO_Y) && !(d->flags & (B200NORM_F_OUT_MEAN | B200NORM_F_OUT_INVSTD)))
      return fail(B200NORM_ERR_BAD_ARG, "NO_Y without a requested statistic has no output");
    if (d->reserved[0] < 0) return fail(B200NORM_ERR_BAD_ARG, "correction must be non-negative");
  }
  if ((d->flags & B200NORM_F_RMS_BWD_HALF_PROJ) &&
      (d->op != B200NORM_OP_RMSNORM_BWD || (d->flags & B200NORM_F_RMS_BWD_CENTERED)))
    return fail(B200NORM_ERR_BAD_ARG, "RMS_BWD_HALF_PROJ is the thunder_rms_norm_backward spelling: RMSNORM_BWD, not centred");
  if ((d->flags & B200NORM_F_OUT_DW) && !(d->flags & B200NORM_F_HAS_WEIGHT) && is_norm_bwd(d->op))
    return fail(B200NORM_ERR_BAD_ARG, "dweight requested without a weight (the builders return none)");
  return B200NORM_OK;
}

void fill_public_plan(const b200norm_desc& d, const Plan& p, b200norm_plan* out) {
  memset(out, 0, sizeof(*out));
  out->variant = p.variant;
  out->cluster_size = p.cluster;
  out->grid_x = p.grid;
  out->block_x = p.block;
  out->threads_per_row = p.tpr ? p.tpr : p.block;
  out->stages = p.stages;
  out->tile_rows = p.tile_rows;
  out->smem_bytes = (int32_t)p.smem;
  out->n_launches = p.n_launches;
  out->form = (p.cached ? B200NORM_FORM_CACHED : 0) | (p.pipe ? B200NORM_FORM_PIPE : 0) |
              (p.psmem ? B200NORM_FORM_PSMEM : 0);
  out->groups = p.groups;
  out->chunks = p.ch;
  out->workspace_bytes = p.ws_bytes;
  out->algorithmic_bytes = algorithmic_bytes(d);
}

// See launch_typed: serialise kernels with an in-kernel grid barrier across the streams of this process.
// Free on a single stream (nothing is recorded or waited for); on a stream switch the new stream waits for
// everything enqueued so far on the previous one. Skipped while either stream is being captured into a CUDA
// graph (an event of one capture must not leak into another stream): graphs replayed concurrently on several
// streams need B200NORM_COOP=1.
void barrier_guard(DeviceCtx& ctx, cudaStream_t stream) {
  std::lock_guard<std::mutex> lock(ctx.barrier_mu);
  if (ctx.barrier_seen && ctx.barrier_stream != stream) {
    cudaStreamCaptureStatus c0 = cudaStreamCaptureStatusNone, c1 = cudaStreamCaptureStatusNone;
    const bool ok0 = cudaStreamIsCapturing(ctx.barrier_stream, &c0) == cudaSuccess;
    const bool ok1 = cudaStreamIsCapturing(stream, &c1) == cudaSuccess;
    if (ok0 && ok1 && c0 == cudaStreamCaptureStatusNone && c1 == cudaStreamCaptureStatusNone) {
      if (ctx.barrier_event == nullptr) cudaEventCreateWithFlags(&ctx.barrier_event, cudaEventDisableTiming);
      if (ctx.barrier_event != nullptr && cudaEventRecord(ctx.barrier_event, ctx.barrier_stream) == cudaSuccess)
        cudaStreamWaitEvent(stream, ctx.barrier_event, 0);
    }
    (void)cudaGetLastError();  // a destroyed predecessor stream has nothing left to wait for
  }
  ctx.barrier_stream = stream;
  ctx.barrier_seen = true;
}

template <typename T>
void launch_finalize(const float* partials, int nparts, int K, int N, void* out_a, void* out_b, float* out_f32,
                     cudaStream_t stream) {
  if (N % 4 == 0 && (reinterpret_cast<uintptr_t>(partials) & 15u) == 0) {
    dim3 grid((unsigned)ceil_div64(N, 128), (unsigned)K);
    finalize_partials_kernel<T, 4><<<grid, dim3(32, kFinY), 0, stream>>>(partials, nparts, K, N, out_a, out_b, out_f32);
  } else {
    dim3 grid((unsigned)ceil_div64(N, 32), (unsigned)K);
    finalize_partials_kernel<T, 1><<<grid, dim3(32, kFinY), 0, stream>>>(partials, nparts, K, N, out_a, out_b, out_f32);
  }
  g_launches++;
}

template <typename T>
RingKernel generic_kernel_for(int op) {
  switch (op) {
    case B200NORM_OP_LAYERNORM_FWD: return generic_row_kernel<T, G_LAYERNORM_FWD>;
    case B200NORM_OP_LAYERNORM_BWD: return generic_row_kernel<T, G_LAYERNORM_BWD>;
    case B200NORM_OP_RMSNORM_FWD: return generic_row_kernel<T, G_RMSNORM_FWD>;
    case B200NORM_OP_RMSNORM_BWD: return generic_row_kernel<T, G_RMSNORM_BWD>;
    case B200NORM_OP_SOFTMAX_FWD: return generic_row_kernel<T, G_SOFTMAX_FWD>;
    case B200NORM_OP_SOFTMAX_BWD: return generic_row_kernel<T, G_SOFTMAX_BWD>;
    case B200NORM_OP_LOGSOFTMAX_FWD: return generic_row_kernel<T, G_LOGSOFTMAX_FWD>;
    case B200NORM_OP_LOGSOFTMAX_BWD: return generic_row_kernel<T, G_LOGSOFTMAX_BWD>;
    case B200NORM_OP_REDUCE_INNER: return generic_row_kernel<T, G_REDUCE_INNER>;
    default: return nullptr;
  }
}

template <typename T>
int launch_typed(const b200norm_desc& d, DeviceCtx& ctx, const Knobs& kn, const Plan& plan, RowParams& rp, void* dw,
                 void* db, float* sums_f32, void* ws_header, cudaStream_t stream) {
  const bool want_cols = has_column_sums(d);
  const int K = partial_k(d.op);
  cudaError_t err;

  if (plan.variant == B200NORM_VARIANT_GENERIC) {
    if (want_cols && (d.flags & (B200NORM_F_PEER_REDUCE | B200NORM_F_PEER_PUBLISH)))
      return fail(B200NORM_ERR_UNSUPPORTED, "peer reduce is only available on the ring kernels (aligned rows)");
    if (d.op == B200NORM_OP_PEER_COMBINE) {
      PeerCombineParams q;
      memset(&q, 0, sizeof(q));
      std::lock_guard<std::mutex> plock(ctx.peer_mu);  // held across the launch: combines pair up with publishes in order
      PeerState& ps = ctx.peers;
      if (ps.world < 1) return fail(B200NORM_ERR_BAD_ARG, "PEER_COMBINE without b200norm_register_peers");
      if (ps.combined >= ps.epoch) return fail(B200NORM_ERR_BAD_ARG, "PEER_COMBINE: nothing published");
      if ((uint64_t)d.rows * (uint64_t)d.cols > ps.cap_floats || d.cols % 4 != 0)
        return fail(B200NORM_ERR_WORKSPACE, "PEER_COMBINE: [%lld,%lld] does not match the exchange buffers",
                    (long long)d.rows, (long long)d.cols);
      q.world = ps.world;
      q.rank = ps.rank;
      for (int r = 0; r < ps.world; ++r) q.xchg[r] = static_cast<uint8_t*>(ps.ptrs[r]);
      q.epoch = ps.combined + 1;
      q.slot_floats = ps.cap_floats;
      q.timeout_ns = ps.timeout_ns;
      q.K = (int32_t)d.rows;
      q.N = (int32_t)d.cols;
      q.out_a = dw;
      q.out_b = db;
      int ncta = (int)ceil_div64(d.rows * d.cols / 4, 512);
      if (ncta > 32) ncta = 32;
      if (ncta < 1) ncta = 1;
      peer_combine_kernel<T><<<ncta, 512, 0, stream>>>(q);
      err = cudaGetLastError();
      if (err != cudaSuccess) return fail(B200NORM_ERR_CUDA + (int)err, "launch failed: %s", cudaGetErrorString(err));
      ps.combined += 1;  // only a combine that was enqueued counts
      g_launches++;
      return B200NORM_OK;
    } else if (d.op == B200NORM_OP_REDUCE_OUTER) {
      dim3 grid((unsigned)ceil_div64(d.cols, 128), (unsigned)plan.nparts);
      generic_colsum_kernel<T, 0><<<grid, 128, 0, stream>>>(rp, plan.rows_per_split);
      g_launches++;
      launch_finalize<T>(rp.partials, plan.nparts, 1, (int)d.cols, dw, nullptr, sums_f32, stream);
    } else if (d.op == B200NORM_OP_FINALIZE_PARTIALS) {
      launch_finalize<T>(static_cast<const float*>(rp.in0), 1, (int)d.rows, (int)d.cols, dw, db, nullptr, stream);
    } else {
      const bool need_rows = !is_norm_bwd(d.op) || (d.flags & (B200NORM_F_OUT_DX | B200NORM_F_ROW_AFFINE));
      if (need_rows) {
        RingKernel k = generic_kernel_for<T>(d.op);
        if (k == nullptr) return fail(B200NORM_ERR_UNSUPPORTED, "no generic kernel for op %d", d.op);
        k<<<plan.grid, kGenericThreads, 0, stream>>>(rp);
        g_launches++;
      }
      if (want_cols && is_norm_bwd(d.op)) {
        dim3 grid((unsigned)ceil_div64(d.cols, 128), (unsigned)plan.nparts);
        if (d.op == B200NORM_OP_LAYERNORM_BWD) {
          generic_colsum_kernel<T, 1><<<grid, 128, 0, stream>>>(rp, plan.rows_per_split);
        } else {
          generic_colsum_kernel<T, 2><<<grid, 128, 0, stream>>>(rp, plan.rows_per_split);
        }
        g_launches++;
        launch_finalize<T>(rp.partials, plan.nparts, K, (int)d.cols, dw, db, sums_f32, stream);
      }
    }
    err = cudaGetLastError();
    if (err != cudaSuccess) return fail(B200NORM_ERR_CUDA + (int)err, "launch failed: %s", cudaGetErrorString(err));
    return B200NORM_OK;
  }

  // ---- ring kernels ----
  RingKernel k = ring_kernel_for<T>(d.op, plan.tpr, plan.ch, plan.cached, plan.psmem, is_add_fwd(d), plan.pipe,
                                    plan.stats_reduce, plan.lean, (d.flags & B200NORM_F_DROPOUT) != 0,
                                    (d.flags & B200NORM_F_ROW_AFFINE) != 0);
  if (k == nullptr)
    return fail(B200NORM_ERR_UNSUPPORTED, "no ring kernel for op %d tpr %d ch %d", d.op, plan.tpr, plan.ch);
  {
    int rc = prepare_ring_kernel(ctx, k, (int)plan.smem, plan.cluster);
    if (rc != B200NORM_OK) return rc;
  }
  // Launch attributes: the cluster shape; programmatic dependent launch (the kernel's barrier
  // setup overlaps its predecessor's tail — every ring kernel calls griddepcontrol.wait before
  // its first global access); and, for a backward that produces dgamma / dbeta, the in-kernel finish
  // of the column sums behind a grid barrier (the reference's inner-outer kernels are cooperative
  // for the same reason, executor.cpp:1307-1312). A combination the driver rejects falls back — no PDL,
  // then no fused finish — and is remembered per (device, kernel) once the fallback has SUCCEEDED
  // (an error that the fallback does not cure is not a launch-mode refusal and is reported as it is).
  int bad = 0;
  {
    std::lock_guard<std::mutex> lock(ctx.mu);
    auto it = ctx.bad_modes.find((const void*)k);
    if (it != ctx.bad_modes.end()) bad = it->second;
  }
  const int bad0 = bad;
  const bool peer = want_cols && (d.flags & (B200NORM_F_PEER_REDUCE | B200NORM_F_PEER_PUBLISH)) != 0;
  // Rows split over a cluster leave the column sums to the finalize kernel (launched with PDL right behind): with
  // tens of thousands of columns the grid barrier + in-kernel finish measured 0.5-2 % slower than the extra launch
  // (bf16 [4096,65536] 290.6 -> 288.8 us, [2048,131072] 322.7 -> 316.2 us); the peer flavours need the fused finish.
  const bool allow_fused = want_cols && plan.block >= 128 && kn.no_fused_finish == 0 && (plan.cluster == 1 || peer);
  const bool deferred = (d.flags & B200NORM_F_PEER_PUBLISH) != 0;
  std::unique_lock<std::mutex> plock(ctx.peer_mu, std::defer_lock);
  PeerState& ps = ctx.peers;
  if (peer) {
    // the exchange happens in the in-kernel finish: refuse before anything is enqueued if that path is closed
    if (!allow_fused || (bad & 1))
      return fail(B200NORM_ERR_UNSUPPORTED, "peer reduce needs the in-kernel finish (not available for this launch)");
    plock.lock();  // held until the launch has been enqueued: epochs advance in launch order, and only on success
    if (ps.world < 1) return fail(B200NORM_ERR_BAD_ARG, "PEER_REDUCE without b200norm_register_peers");
    if ((uint64_t)K * (uint64_t)d.cols > ps.cap_floats)
      return fail(B200NORM_ERR_WORKSPACE, "exchange buffers hold %u floats per slot, need %llu", ps.cap_floats,
                  (unsigned long long)K * (unsigned long long)d.cols);
    if (deferred && ps.epoch - ps.combined >= 2)
      return fail(B200NORM_ERR_BAD_ARG, "two publishes are outstanding: launch OP_PEER_COMBINE before the next one");
    rp.peer_world = ps.world;
    rp.peer_rank = ps.rank;
    rp.peer_deferred = deferred ? 1 : 0;
    for (int q = 0; q < ps.world; ++q) rp.peer_xchg[q] = static_cast<uint8_t*>(ps.ptrs[q]);
    rp.peer_epoch = deferred ? ps.epoch + 1 : ps.ll_epoch + 1;  // committed below, after a successful launch
    rp.peer_slot_floats = ps.cap_floats;
    rp.peer_ll_off = ps.ll_off;
    rp.peer_ll_parity_bytes = ps.ll_parity_bytes;
    rp.peer_timeout_ns = ps.timeout_ns;
  }
  const bool allow_pdl = kn.no_pdl == 0;
  for (;;) {
    const bool fused = allow_fused && !(bad & 1);
    const bool pdl = allow_pdl && !(bad & 2);
    rp.fused_finish = fused ? 1 : 0;
    rp.sync_words = reinterpret_cast<uint32_t*>(ws_header);
    rp.dw_out = dw;
    rp.db_out = db;
    rp.sums_out = sums_f32;
    rp.timeline = g_timeline.load();
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)plan.grid);
    cfg.blockDim = dim3((unsigned)plan.block);
    cfg.dynamicSmemBytes = plan.smem;
    cfg.stream = stream;
    cudaLaunchAttribute attrs[3];
    int na = 0;
    attrs[na].id = cudaLaunchAttributeClusterDimension;
    attrs[na].val.clusterDim.x = (unsigned)plan.cluster;
    attrs[na].val.clusterDim.y = 1;
    attrs[na].val.clusterDim.z = 1;
    ++na;
    // The in-kernel finish needs every CTA resident at its grid barrier. The grid is at most one CTA per SM
    // (one cluster per cluster slot), so on ONE stream it is co-resident as soon as the SMs are free: whatever
    // holds an SM when the kernel starts (the stream's predecessor under PDL, an NCCL kernel on a side stream)
    // finishes without waiting for this kernel, and the successor's CTAs cannot start before all of ours have.
    // Unlike a cooperative launch, the plain launch overlaps its setup with the predecessor's tail (measured:
    // 77.7 -> 74.2 us at C2). Two barrier kernels on two STREAMS could each get part of the SMs and spin for ever
    // (or until the watchdog traps): barrier_guard() serialises them across the streams of this process.
    // B200NORM_COOP=1 asks the driver for the formal co-residency guarantee instead (other processes under MPS).
    const bool coop = fused && kn.coop != 0;
    if (fused && !coop && kn.no_guard == 0) barrier_guard(ctx, stream);
    if (coop) {
      attrs[na].id = cudaLaunchAttributeCooperative;
      attrs[na].val.cooperative = 1;
      ++na;
    }
    if (pdl) {
      attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attrs[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    cfg.attrs = attrs;
    cfg.numAttrs = (unsigned)na;
    err = cudaLaunchKernelEx(&cfg, k, (const RowParams)rp);
    if (err == cudaSuccess) {
      g_launches++;
      if (peer) {  // the launch is in the stream: its epoch is taken
        if (deferred) {
          ps.epoch += 1;
        } else {
          ps.ll_epoch += 1;
        }
        plock.unlock();
      }
      if (bad != bad0) {  // the fallback cured it: remember the refused mode for this (device, kernel)
        std::lock_guard<std::mutex> lock(ctx.mu);
        ctx.bad_modes[(const void*)k] = bad;
      }
      if (kn.debug > 1)
        fprintf(stderr, "[b200norm] op %d launched grid %d block %d cluster %d fused=%d pdl=%d\n", d.op, plan.grid,
                plan.block, plan.cluster, (int)fused, (int)pdl);
      if (want_cols && !fused) {
        launch_finalize<T>(rp.partials, plan.nparts, K, (int)d.cols, dw, db, sums_f32, stream);
        err = cudaGetLastError();
        if (err != cudaSuccess)
          return fail(B200NORM_ERR_CUDA + (int)err, "finalize launch failed: %s", cudaGetErrorString(err));
      }
      return B200NORM_OK;
    }
    (void)cudaGetLastError();
    if (kn.debug)
      fprintf(stderr, "[b200norm] launch with fused=%d pdl=%d refused: %s\n", (int)fused, (int)pdl,
              cudaGetErrorString(err));
    int newbad = bad;
    if (pdl) {
      newbad |= 2;
    } else if (fused) {
      newbad |= 1;
    }
    if (newbad == bad || (peer && (newbad & 1)))
      return fail(B200NORM_ERR_CUDA + (int)err, "ring launch failed: %s", cudaGetErrorString(err));
    bad = newbad;
  }
}


// ---------------------------------------------------------------------------
// argument checks
// ---------------------------------------------------------------------------
bool is_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int check_matrix(const b200norm_tensor& t, const b200norm_desc& d, int dtype, const char* name, int64_t* row_stride) {
  if (t.dtype != dtype) return fail(B200NORM_ERR_BAD_ARG, "%s: dtype %d, expected %d", name, t.dtype, dtype);
  if (t.ndim != 2) return fail(B200NORM_ERR_BAD_ARG, "%s: expected a 2-D [rows, cols] view, got ndim %d", name, t.ndim);
  if (t.sizes[0] != d.rows || t.sizes[1] != d.cols)
    return fail(B200NORM_ERR_BAD_ARG, "%s: shape [%lld,%lld] != [%lld,%lld]", name, (long long)t.sizes[0],
                (long long)t.sizes[1], (long long)d.rows, (long long)d.cols);
  if (d.cols > 1 && t.strides[1] != 1)
    return fail(B200NORM_ERR_UNSUPPORTED, "%s: innermost stride %lld != 1", name, (long long)t.strides[1]);
  if (d.rows > 1 && t.strides[0] < d.cols)
    return fail(B200NORM_ERR_UNSUPPORTED, "%s: overlapping rows (stride %lld < cols)", name, (long long)t.strides[0]);
  if (t.data == nullptr && d.rows * d.cols > 0) return fail(B200NORM_ERR_BAD_ARG, "%s: null data", name);
  *row_stride = d.rows > 1 ? t.strides[0] : d.cols;
  return B200NORM_OK;
}

int check_vector(const b200norm_tensor& t, int64_t n, int dtype, const char* name) {
  if (t.dtype != dtype) return fail(B200NORM_ERR_BAD_ARG, "%s: dtype %d, expected %d", name, t.dtype, dtype);
  int64_t numel = 1;
  int64_t expect_stride = 1;
  bool contiguous = true;
  for (int i = t.ndim - 1; i >= 0; --i) {
    if (t.sizes[i] != 1 && t.strides[i] != expect_stride) contiguous = false;
    expect_stride *= t.sizes[i];
    numel *= t.sizes[i];
  }
  if (t.ndim < 1 || t.ndim > B200NORM_MAX_DIMS || numel != n)
    return fail(B200NORM_ERR_BAD_ARG, "%s: expected %lld elements, got %lld", name, (long long)n, (long long)numel);
  if (!contiguous) return fail(B200NORM_ERR_UNSUPPORTED, "%s: must be contiguous", name);
  if (t.data == nullptr && n > 0) return fail(B200NORM_ERR_BAD_ARG, "%s: null data", name);
  return B200NORM_OK;
}

}  // namespace

// ---------------------------------------------------------------------------
// exported C ABI
// ---------------------------------------------------------------------------
namespace {
int ce_bwd_grid(const CeBwdParams& q, int vec, int sm_count) {
  const int64_t units = (q.cols + vec - 1) / vec;
  const int64_t items = q.rows * ((units + kCeThreads * kCeU - 1) / (kCeThreads * kCeU));
  const int64_t cap = (int64_t)sm_count * 8;  // eight 256-thread CTAs per SM: 128 KB of loads in flight
  return (int)(items < cap ? (items < 1 ? 1 : items) : cap);
}
template <typename T>
int launch_ce_bwd_typed(const CeBwdParams& q, bool aligned, int sm_count, cudaStream_t stream) {
  const bool vec = aligned && q.cols % IO<T>::VEC == 0;
  if (vec) {
    cross_entropy_bwd_kernel<T, true><<<ce_bwd_grid(q, IO<T>::VEC, sm_count), kCeThreads, 0, stream>>>(q);
  } else {
    cross_entropy_bwd_kernel<T, false><<<ce_bwd_grid(q, 1, sm_count), kCeThreads, 0, stream>>>(q);
  }
  const cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return fail(B200NORM_ERR_CUDA + (int)err, "launch failed: %s", cudaGetErrorString(err));
  g_launches++;
  return B200NORM_OK;
}
int launch_ce_bwd(int dt, const CeBwdParams& q, bool aligned, int sm_count, cudaStream_t stream) {
  switch (dt) {
    case B200NORM_DT_F32: return launch_ce_bwd_typed<float>(q, aligned, sm_count, stream);
    case B200NORM_DT_BF16: return launch_ce_bwd_typed<__nv_bfloat16>(q, aligned, sm_count, stream);
    case B200NORM_DT_F16: return launch_ce_bwd_typed<__half>(q, aligned, sm_count, stream);
    default: return fail(B200NORM_ERR_UNSUPPORTED, "dtype %d", dt);
  }
}
}  // namespace

extern "C" {

int b200norm_abi_version(void) { return B200NORM_ABI_VERSION; }

const char* b200norm_last_error(void) { return g_err; }

uint64_t b200norm_launch_count(void) { return g_launches.load(); }

int b200norm_register_peers(int world, int rank, void* const* exchange_ptrs, size_t exchange_bytes) {
  DeviceCtx& ctx = device_ctx();
  std::lock_guard<std::mutex> lock(ctx.peer_mu);
  PeerState& ps = ctx.peers;
  if (world == 0) {
    ps = PeerState();
    return B200NORM_OK;
  }
  if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world || exchange_ptrs == nullptr)
    return fail(B200NORM_ERR_BAD_ARG, "register_peers: world %d (1..%d), rank %d", world, kMaxPeers, rank);
  // layout: header | two publish slots of cap floats | LL region: 2 parities x world x cap (value, epoch) pairs
  const size_t per_float = 2 * 4 + 2 * (size_t)world * 8;
  const size_t cap = exchange_bytes > (size_t)kPeerHeaderBytes ? (exchange_bytes - kPeerHeaderBytes) / per_float / 4 * 4 : 0;
  if (cap < 4)
    return fail(B200NORM_ERR_BAD_ARG, "register_peers: exchange buffers too small (%zu bytes)", exchange_bytes);
  for (int q = 0; q < world; ++q) {
    if (exchange_ptrs[q] == nullptr || (reinterpret_cast<uintptr_t>(exchange_ptrs[q]) & 255u) != 0)
      return fail(B200NORM_ERR_BAD_ARG, "register_peers: buffer %d is null or not 256-byte aligned", q);
  }
  ps = PeerState();
  ps.world = world;
  ps.rank = rank;
  for (int q = 0; q < world; ++q) ps.ptrs[q] = exchange_ptrs[q];
  ps.bytes = exchange_bytes;
  ps.cap_floats = (uint32_t)(cap > 0x3fffffffu ? 0x3ffffffcu : cap);
  ps.ll_off = (uint32_t)(kPeerHeaderBytes + 2 * (size_t)ps.cap_floats * 4);
  ps.ll_parity_bytes = (uint32_t)((size_t)world * ps.cap_floats * 8);
  // Waiting for a PEER is not an error for minutes (first iteration, evaluation or checkpoint on another rank,
  // a data-loader stall): the in-kernel waits for peers get ten minutes by default, B200NORM_PEER_TIMEOUT_S
  // changes that (0 = wait for ever, like the NCCL kernel this replaces).
  const Knobs kn = read_knobs();
  const int tmo = kn.peer_timeout_s >= 0 ? kn.peer_timeout_s : 600;
  ps.timeout_ns = (uint64_t)tmo * 1000000000ull;
  return B200NORM_OK;
}

size_t b200norm_peer_exchange_bytes(int world, int64_t max_floats) {
  if (world < 1 || world > kMaxPeers || max_floats < 1) return 0;
  const size_t cap = ((size_t)max_floats + 3) / 4 * 4;
  return (size_t)kPeerHeaderBytes + cap * (2 * 4 + 2 * (size_t)world * 8);
}

void b200norm_debug_set_timeline(void* device_buf) { g_timeline.store(static_cast<unsigned long long*>(device_buf)); }

const char* b200norm_variant_name(int variant) {
  switch (variant) {
    case B200NORM_VARIANT_AUTO: return "auto";
    case B200NORM_VARIANT_GENERIC: return "generic";
    case B200NORM_VARIANT_TMA: return "tma";
    case B200NORM_VARIANT_CLUSTER: return "cluster";
    default: return "?";
  }
}

int b200norm_device_info(int* has_device, int* sm_count, int* smem_per_block_optin, int* cc_major, int* cc_minor) {
  const DeviceInfo& d = device_info();
  if (has_device) *has_device = d.has_device;
  if (sm_count) *sm_count = d.sm_count;
  if (smem_per_block_optin) *smem_per_block_optin = d.smem_optin;
  if (cc_major) *cc_major = d.cc_major;
  if (cc_minor) *cc_minor = d.cc_minor;
  return 0;
}

int b200norm_query(const b200norm_desc* desc, b200norm_plan* plan) {
  int rc = check_desc(desc);
  if (rc != B200NORM_OK) return rc;
  if (is_outer(*desc)) {
    if (plan != nullptr) memset(plan, 0, sizeof(*plan));
    const OuterArgs oa = outer_args_from(*desc);
    const uint64_t ws = outer_generic_workspace_bytes(oa);  // covers a launch-time fall-back to the generic family
    if (desc->rows == 0 || desc->cols == 0) {
      if (plan != nullptr) plan->workspace_bytes = ws;
      return B200NORM_OK;
    }
    OuterPlan op;
    char msg[256] = "";
    bool persistent = !oa.force_generic && outer_make_plan(oa, &op, msg, sizeof(msg));
    if (persistent && !oa.force_persistent && outer_prefers_rows(oa)) persistent = false;
    if (!persistent && (desc->variant == B200NORM_VARIANT_TMA || desc->variant == B200NORM_VARIANT_CLUSTER))
      return fail(B200NORM_ERR_UNSUPPORTED, "requested kernel family cannot run this outer segment: %s", msg);
    if (plan != nullptr) {
      if (persistent) {
        plan->variant = op.cluster > 1 ? B200NORM_VARIANT_CLUSTER : B200NORM_VARIANT_TMA;
        plan->cluster_size = op.cluster;
        plan->grid_x = op.grid;
        plan->block_x = op.block;
        plan->threads_per_row = op.slab_bytes / 16;  // threads across one slab row
        plan->stages = op.depth;
        plan->tile_rows = op.rows_cta;
        plan->smem_bytes = (int32_t)op.smem;
        plan->n_launches = 1;
      } else if (int rg = 1, rpi = 1; outer_rows_shape(oa, &rg, &rpi)) {
        // few columns, many rows (outer_rows.cuh): one cooperative kernel, a block of rows per CTA
        plan->variant = B200NORM_VARIANT_GENERIC;
        plan->cluster_size = 1;
        plan->grid_x = rg;
        plan->block_x = 512;
        plan->threads_per_row = 512 / rpi;  // 16-byte chunks of one row
        plan->tile_rows = rpi;
        plan->n_launches = 1;
      } else {
        // shape-agnostic family (outer_generic.cuh): statistics / merge / apply kernels over (column blocks x row splits)
        int gx = 1, ns = 1, rps = 1;
        outer_generic_plan(oa, &gx, &ns, &rps);
        plan->variant = B200NORM_VARIANT_GENERIC;
        plan->cluster_size = 1;
        plan->grid_x = gx * ns;
        plan->block_x = 128;
        plan->threads_per_row = 1;
        plan->tile_rows = rps;
        plan->n_launches = 3;
      }
      plan->workspace_bytes = ws;
      plan->algorithmic_bytes = algorithmic_bytes(*desc);
    }
    return B200NORM_OK;
  }
  Plan p;
  // rows are assumed 16-byte aligned and contiguous at query time; the launch
  // re-plans with the real pointers (the workspace bound covers both outcomes).
  rc = make_plan(*desc, device_ctx(), read_knobs(), /*aligned=*/true, /*rows_contiguous=*/true, &p);
  if (rc != B200NORM_OK) return rc;
  if (plan != nullptr) fill_public_plan(*desc, p, plan);
  return B200NORM_OK;
}

size_t b200norm_workspace_bytes(const b200norm_desc* desc) {
  b200norm_plan plan;
  if (b200norm_query(desc, &plan) != B200NORM_OK) return 0;
  return (size_t)plan.workspace_bytes;
}

int b200norm_launch(const b200norm_desc* desc, const b200norm_tensor* inputs, int n_inputs,
                    const b200norm_tensor* outputs, int n_outputs, void* workspace, size_t workspace_bytes,
                    void* stream_v) {
  int rc = check_desc(desc);
  if (rc != B200NORM_OK) return rc;
  const b200norm_desc& d = *desc;
  DeviceCtx& ctx = device_ctx();
  const DeviceInfo& dev = ctx.info;
  const Knobs kn = read_knobs();
  if (!dev.has_device) return fail(B200NORM_ERR_NO_DEVICE, "no CUDA device visible (there is no CPU fallback)");
  if (dev.cc_major != 10)
    return fail(B200NORM_ERR_NO_DEVICE, "device is sm_%d%d; this library contains sm_100a code only", dev.cc_major,
                dev.cc_minor);
  if ((n_inputs > 0 && inputs == nullptr) || (n_outputs > 0 && outputs == nullptr))
    return fail(B200NORM_ERR_BAD_ARG, "null tensor list");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);

  const int dt = d.io_dtype;
  const bool has_w = (d.flags & B200NORM_F_HAS_WEIGHT) != 0;
  const bool has_b = (d.flags & B200NORM_F_HAS_BIAS) != 0;
  RowParams rp;
  memset(&rp, 0, sizeof(rp));
  rp.rows = d.rows;
  rp.cols = (int32_t)d.cols;
  rp.eps = (float)d.eps;
  rp.inv_n = d.cols > 0 ? (float)(1.0 / (double)d.cols) : 0.f;
  rp.n_f = (float)d.cols;
  rp.proj_scale = (d.flags & B200NORM_F_RMS_BWD_HALF_PROJ) ? 0.5f : 1.0f;
  {
    const double denom = (double)d.cols - (double)d.reserved[0];  // var = M2 / max(N - correction, 0), as at::var_mean
    rp.var_scale = (float)(1.0 / (denom > 0.0 ? denom : 0.0));
  }
  rp.flags = d.flags;
  const bool row_aff = (d.flags & B200NORM_F_ROW_AFFINE) != 0;
  rp.channels = row_aff ? d.reserved[3] : 0;
  const bool dropout = (d.flags & B200NORM_F_DROPOUT) != 0;
  if (dropout) {
    const double keep = dropout_keep_prob(d);
    dropout_thresholds(keep, &rp.drop_t1, &rp.drop_t2);
    rp.drop_scale = (float)(1.0 / keep);
  }
  void* dw = nullptr;
  void* db = nullptr;
  float* sums_f32 = nullptr;
  int ii = 0, oi = 0;
  bool aligned = true;
  bool rows_contiguous = true;

#define NEED_IN(n)  if (ii + (n) > n_inputs) return fail(B200NORM_ERR_BAD_ARG, "op %d: too few inputs (%d)", d.op, n_inputs)
#define NEED_OUT(n) if (oi + (n) > n_outputs) return fail(B200NORM_ERR_BAD_ARG, "op %d: too few outputs (%d)", d.op, n_outputs)
#define CHECK(expr) do { rc = (expr); if (rc != B200NORM_OK) return rc; } while (0)

  auto bind_in_matrix = [&](const void** ptr, int64_t* stride, const char* name) -> int {
    const b200norm_tensor& t = inputs[ii++];
    int r = check_matrix(t, d, dt, name, stride);
    if (r != B200NORM_OK) return r;
    *ptr = t.data;
    if (!is_aligned16(t.data) || ((*stride * elem_size(dt)) % 16) != 0) aligned = false;
    if (*stride != d.cols) rows_contiguous = false;
    return B200NORM_OK;
  };
  auto bind_out_matrix = [&](void** ptr, int64_t* stride, const char* name) -> int {
    const b200norm_tensor& t = outputs[oi++];
    int r = check_matrix(t, d, dt, name, stride);
    if (r != B200NORM_OK) return r;
    *ptr = t.data;
    if (!is_aligned16(t.data) || ((*stride * elem_size(dt)) % 16) != 0) aligned = false;
    return B200NORM_OK;
  };
  auto bind_vec = [&](const b200norm_tensor& t, int64_t n, int dtype, const char* name, void** ptr) -> int {
    int r = check_vector(t, n, dtype, name);
    if (r != B200NORM_OK) return r;
    *ptr = t.data;
    return B200NORM_OK;
  };

  // int64 vector (class indices): n contiguous elements, 8-byte aligned (the dtype field is not interpreted)
  auto bind_i64 = [&](const b200norm_tensor& t, int64_t n, const char* name, void** ptr) -> int {
    if (t.data == nullptr || (reinterpret_cast<uintptr_t>(t.data) & 7u) != 0) return fail(B200NORM_ERR_BAD_ARG, "%s: null or misaligned int64 data", name);
    int64_t numel = 1;
    for (int i = 0; i < t.ndim; ++i) numel *= t.sizes[i];
    if (numel != n) return fail(B200NORM_ERR_BAD_ARG, "%s: %lld elements, expected %lld", name, (long long)numel, (long long)n);
    for (int i = 0; i < t.ndim; ++i)
      if (t.sizes[i] > 1 && t.strides[i] != 1 && t.sizes[i] == numel) return fail(B200NORM_ERR_UNSUPPORTED, "%s must be contiguous", name);
    *ptr = t.data;
    return B200NORM_OK;
  };
  // dropout mask: [M,N] bytes with contiguous rows (the dtype field is not interpreted)
  auto bind_mask = [&](const b200norm_tensor& t, const char* name, void** ptr) -> int {
    if (t.data == nullptr) return fail(B200NORM_ERR_BAD_ARG, "%s: null data", name);
    int64_t numel = 1;
    for (int i = 0; i < t.ndim; ++i) numel *= t.sizes[i];
    if (numel != d.rows * d.cols) return fail(B200NORM_ERR_BAD_ARG, "%s: %lld elements, expected %lld", name, (long long)numel, (long long)(d.rows * d.cols));
    int64_t expect = 1;
    for (int i = t.ndim - 1; i >= 0; --i) {
      if (t.sizes[i] != 1 && t.strides[i] != expect) return fail(B200NORM_ERR_UNSUPPORTED, "%s must be contiguous", name);
      expect *= t.sizes[i];
    }
    *ptr = t.data;
    if (!is_aligned16(t.data) || d.cols % 16 != 0) aligned = false;  // the ring kernels move the mask in 16-byte chunks
    return B200NORM_OK;
  };

  switch (d.op) {
    case B200NORM_OP_LAYERNORM_FWD:
    case B200NORM_OP_RMSNORM_FWD: {
      const bool ln = d.op == B200NORM_OP_LAYERNORM_FWD;
      const bool add = (d.flags & B200NORM_F_ADD_RESIDUAL) != 0;
      if (!add && (d.flags & (B200NORM_F_OUT_SUM | B200NORM_F_ROUND_SUM)))
        return fail(B200NORM_ERR_BAD_ARG, "OUT_SUM / ROUND_SUM need ADD_RESIDUAL");
      NEED_IN(1 + (add ? 1 : 0) + (has_w ? 1 : 0) + ((ln && has_b) ? 1 : 0));
      CHECK(bind_in_matrix(&rp.in0, &rp.in0_stride, "x"));
      if (add) CHECK(bind_in_matrix(&rp.in1, &rp.in1_stride, "residual"));
      if (has_w) {
        void* w = nullptr;
        CHECK(bind_vec(inputs[ii++], row_aff ? (int64_t)rp.channels : d.cols, dt, "weight", &w));
        rp.weight = w;
        if (!row_aff && !is_aligned16(w)) aligned = false;
      }
      if (ln && has_b) {
        void* b = nullptr;
        CHECK(bind_vec(inputs[ii++], row_aff ? (int64_t)rp.channels : d.cols, dt, "bias", &b));
        rp.bias = b;
        if (!row_aff && !is_aligned16(b)) aligned = false;
      }
      const bool no_y = ln && (d.flags & B200NORM_F_NO_Y) != 0;
      if (!no_y) {
        NEED_OUT(1);
        CHECK(bind_out_matrix(&rp.out0, &rp.out_stride, "y"));
      }
      if (add && (d.flags & B200NORM_F_OUT_SUM)) {
        NEED_OUT(1);
        CHECK(bind_out_matrix(&rp.out1, &rp.out1_stride, "sum"));
      }
      if (ln) {
        const int64_t nstat = is_outer(d) ? d.cols : d.rows;  // statistics per column when the reduction runs over dim 0
        if (d.flags & B200NORM_F_OUT_MEAN) {
          NEED_OUT(1);
          void* m = nullptr;
          CHECK(bind_vec(outputs[oi++], nstat, B200NORM_DT_F32, "mean", &m));
          rp.stat0 = static_cast<float*>(m);
        }
        if (d.flags & B200NORM_F_OUT_INVSTD) {
          NEED_OUT(1);
          void* m = nullptr;
          CHECK(bind_vec(outputs[oi++], nstat, B200NORM_DT_F32, "invstd", &m));
          rp.stat1 = static_cast<float*>(m);
        }
      } else {
        NEED_OUT(1);
        void* m = nullptr;
        CHECK(bind_vec(outputs[oi++], d.rows, B200NORM_DT_F32, "stat", &m));
        rp.stat1 = static_cast<float*>(m);
      }
      if (dropout) {
        NEED_IN(1);
        NEED_OUT(1);
        const b200norm_tensor& rng = inputs[ii++];
        if (rng.data == nullptr || (reinterpret_cast<uintptr_t>(rng.data) & 7u) != 0)
          return fail(B200NORM_ERR_BAD_ARG, "rng: a device tensor of two int64 {seed, offset} is required");
        rp.rng = static_cast<const unsigned long long*>(rng.data);
        CHECK(bind_mask(outputs[oi++], "mask", reinterpret_cast<void**>(&rp.mask_out)));
      }
      break;
    }
    case B200NORM_OP_LAYERNORM_BWD:
    case B200NORM_OP_RMSNORM_BWD: {
      const bool ln = d.op == B200NORM_OP_LAYERNORM_BWD;
      NEED_IN(2 + (ln ? 2 : 1) + (has_w ? 1 : 0));
      CHECK(bind_in_matrix(&rp.in0, &rp.in0_stride, "dy"));
      CHECK(bind_in_matrix(&rp.in1, &rp.in1_stride, "x"));
      const int64_t nstat = is_outer(d) ? d.cols : d.rows;  // saved statistics per column when the reduction ran over dim 0
      if (ln) {
        void* m = nullptr;
        CHECK(bind_vec(inputs[ii++], nstat, B200NORM_DT_F32, "mean", &m));
        rp.stat0 = static_cast<float*>(m);
      }
      {
        void* m = nullptr;
        CHECK(bind_vec(inputs[ii++], nstat, B200NORM_DT_F32, ln ? "invstd" : "stat", &m));
        rp.stat1 = static_cast<float*>(m);
      }
      if (has_w) {
        void* w = nullptr;
        CHECK(bind_vec(inputs[ii++], row_aff ? (int64_t)rp.channels : d.cols, dt, "weight", &w));
        rp.weight = w;
        if (!row_aff && !is_aligned16(w)) aligned = false;
      }
      if (dropout) {
        NEED_IN(2);
        CHECK(bind_in_matrix(&rp.in2, &rp.in2_stride, "residual"));
        void* m = nullptr;
        CHECK(bind_mask(inputs[ii++], "mask", &m));
        rp.mask_in = static_cast<const uint8_t*>(m);
      }
      if (d.flags & B200NORM_F_OUT_DX) {
        NEED_OUT(1);
        CHECK(bind_out_matrix(&rp.out0, &rp.out_stride, "dx"));
        if (dropout) {
          NEED_OUT(1);
          CHECK(bind_out_matrix(&rp.out1, &rp.out1_stride, "dresidual"));
        }
      }
      const bool want_dw = (d.flags & B200NORM_F_OUT_DW) != 0;
      const bool want_db = ln && (d.flags & B200NORM_F_OUT_DB) != 0;
      if (row_aff) {  // instance norm: per-row sums (fp32 [M]) instead of column sums
        void* v = nullptr;
        if (want_dw) {
          NEED_OUT(1);
          CHECK(bind_vec(outputs[oi++], d.rows, B200NORM_DT_F32, "rowsum(dy * x_hat)", &v));
          rp.rowsum_dyxh = static_cast<float*>(v);
        }
        if (want_db) {
          NEED_OUT(1);
          CHECK(bind_vec(outputs[oi++], d.rows, B200NORM_DT_F32, "rowsum(dy)", &v));
          rp.rowsum_dy = static_cast<float*>(v);
        }
        rp.flags &= ~(uint32_t)(B200NORM_F_OUT_DW | B200NORM_F_OUT_DB);  // the kernels' column accumulators stay off
      } else if ((want_dw || want_db) && (d.flags & B200NORM_F_PEER_PUBLISH)) {
        if (d.flags & (B200NORM_F_PARTIALS_F32 | B200NORM_F_PEER_REDUCE))
          return fail(B200NORM_ERR_BAD_ARG, "PEER_PUBLISH excludes PARTIALS_F32 and PEER_REDUCE");
        // the sums go to the exchange slot only; OP_PEER_COMBINE produces the tensors
      } else if (want_dw || want_db) {
        if (d.flags & B200NORM_F_PARTIALS_F32) {
          NEED_OUT(1);
          void* s = nullptr;
          CHECK(bind_vec(outputs[oi++], (int64_t)partial_k(d.op) * d.cols, B200NORM_DT_F32, "column sums", &s));
          sums_f32 = static_cast<float*>(s);
        } else {
          if (want_dw) {
            NEED_OUT(1);
            CHECK(bind_vec(outputs[oi++], d.cols, dt, "dweight", &dw));
          }
          if (want_db) {
            NEED_OUT(1);
            CHECK(bind_vec(outputs[oi++], d.cols, dt, "dbias", &db));
          }
        }
      }
      break;
    }
    case B200NORM_OP_SOFTMAX_FWD:
    case B200NORM_OP_LOGSOFTMAX_FWD:
      NEED_IN(1);
      NEED_OUT(1);
      CHECK(bind_in_matrix(&rp.in0, &rp.in0_stride, "x"));
      CHECK(bind_out_matrix(&rp.out0, &rp.out_stride, "y"));
      break;
    case B200NORM_OP_SOFTMAX_BWD:
    case B200NORM_OP_LOGSOFTMAX_BWD:
      NEED_IN(2);
      NEED_OUT(1);
      CHECK(bind_in_matrix(&rp.in0, &rp.in0_stride, "dy"));
      CHECK(bind_in_matrix(&rp.in1, &rp.in1_stride, "y"));
      CHECK(bind_out_matrix(&rp.out0, &rp.out_stride, "dx"));
      break;
    case B200NORM_OP_REDUCE_INNER: {
      NEED_IN(1);
      NEED_OUT(1);
      CHECK(bind_in_matrix(&rp.in0, &rp.in0_stride, "x"));
      if (d.flags & B200NORM_F_REDUCE_NLL) {  // cross-entropy forward: + target[M] int64; loss[M], lse[M] fp32
        NEED_IN(1);
        NEED_OUT(2);
        void* tg = nullptr;
        CHECK(bind_i64(inputs[ii++], d.rows, "target", &tg));
        rp.target = static_cast<const long long*>(tg);
        rp.ignore_index = d.reserved[3];
        void* v = nullptr;
        CHECK(bind_vec(outputs[oi++], d.rows, B200NORM_DT_F32, "loss", &v));
        rp.stat1 = static_cast<float*>(v);
        CHECK(bind_vec(outputs[oi++], d.rows, B200NORM_DT_F32, "lse", &v));
        rp.stat0 = static_cast<float*>(v);
        break;
      }
      void* r = nullptr;
      CHECK(bind_vec(outputs[oi++], d.rows, dt, "result", &r));
      rp.out0 = r;
      break;
    }
    case B200NORM_OP_NLL_SOFTMAX_BWD: {
      NEED_IN(4);
      NEED_OUT(1);
      CeBwdParams q;
      memset(&q, 0, sizeof(q));
      const void* xin = nullptr;
      CHECK(bind_in_matrix(&xin, &q.x_stride, "x"));
      q.x = xin;
      void* v = nullptr;
      CHECK(bind_i64(inputs[ii++], d.rows, "target", &v));
      q.target = static_cast<const long long*>(v);
      CHECK(bind_vec(inputs[ii++], d.rows, B200NORM_DT_F32, "lse", &v));
      q.lse = static_cast<const float*>(v);
      CHECK(bind_vec(inputs[ii++], 1, B200NORM_DT_F32, "g", &v));
      q.gscale = static_cast<const float*>(v);
      void* dxp = nullptr;
      CHECK(bind_out_matrix(&dxp, &q.dx_stride, "dx"));
      q.dx = dxp;
      q.rows = d.rows;
      q.cols = (int32_t)d.cols;
      q.ignore_index = d.reserved[3];
      if (d.rows == 0 || d.cols == 0) return B200NORM_OK;
      return launch_ce_bwd(dt, q, aligned, dev.sm_count, stream);
    }
    case B200NORM_OP_REDUCE_OUTER: {
      NEED_IN(1);
      NEED_OUT(1);
      CHECK(bind_in_matrix(&rp.in0, &rp.in0_stride, "x"));
      if (d.flags & (B200NORM_F_REDUCE_MAX | B200NORM_F_REDUCE_FMAX))
        return fail(B200NORM_ERR_UNSUPPORTED, "outer max reductions are not on the hot path");
      if (d.flags & B200NORM_F_PARTIALS_F32) {
        void* s = nullptr;
        CHECK(bind_vec(outputs[oi++], d.cols, B200NORM_DT_F32, "column sums", &s));
        sums_f32 = static_cast<float*>(s);
      } else {
        CHECK(bind_vec(outputs[oi++], d.cols, dt, "result", &dw));
      }
      break;
    }
    case B200NORM_OP_PEER_COMBINE: {
      if (d.rows < 1 || d.rows > 2) return fail(B200NORM_ERR_BAD_ARG, "peer_combine: rows (K) must be 1 or 2");
      NEED_OUT((int)d.rows);
      CHECK(bind_vec(outputs[oi++], d.cols, dt, "out0", &dw));
      if (d.rows == 2) CHECK(bind_vec(outputs[oi++], d.cols, dt, "out1", &db));
      break;
    }
    case B200NORM_OP_FINALIZE_PARTIALS: {
      NEED_IN(1);
      if (d.rows < 1 || d.rows > 2) return fail(B200NORM_ERR_BAD_ARG, "finalize: rows (K) must be 1 or 2");
      NEED_OUT((int)d.rows);
      void* s = nullptr;
      CHECK(bind_vec(inputs[ii++], d.rows * d.cols, B200NORM_DT_F32, "sums", &s));
      rp.in0 = s;
      CHECK(bind_vec(outputs[oi++], d.cols, dt, "out0", &dw));
      if (d.rows == 2) CHECK(bind_vec(outputs[oi++], d.cols, dt, "out1", &db));
      break;
    }
    default:
      return fail(B200NORM_ERR_UNSUPPORTED, "op %d", d.op);
  }
#undef NEED_IN
#undef NEED_OUT
#undef CHECK

  // outputs aliased to inputs are rejected, as the precedent does
  // (cutlass_executor.cpp:101-104)
  for (int o = 0; o < n_outputs; ++o) {
    for (int i = 0; i < n_inputs; ++i) {
      if (outputs[o].data != nullptr && outputs[o].data == inputs[i].data)
        return fail(B200NORM_ERR_UNSUPPORTED, "output %d aliases input %d", o, i);
    }
  }

  if (is_outer(d)) {
    // outer-persistent softmax family (outer.cu): its own planner and launch
    if (d.rows == 0 || d.cols == 0) return B200NORM_OK;
    OuterArgs a = outer_args_from(d);
    a.workspace = workspace;
    a.workspace_bytes = workspace_bytes;
    a.timeline = g_timeline.load();
    a.in0 = rp.in0;
    a.in1 = rp.in1;
    a.out = rp.out0;
    a.in0_stride = rp.in0_stride;
    a.in1_stride = rp.in1_stride;
    a.out_stride = rp.out_stride;
    a.weight = rp.weight;
    a.bias = rp.bias;
    if (d.op == B200NORM_OP_LAYERNORM_BWD) {
      a.mean_in = rp.stat0;
      a.invstd_in = rp.stat1;
      a.dw_out = dw;
      a.db_out = db;
    } else {
      a.mean_out = rp.stat0;
      a.invstd_out = rp.stat1;
    }
    if (int rg = 0, rpi = 0; kn.coop == 0 && kn.no_guard == 0 && outer_rows_shape(a, &rg, &rpi))
      barrier_guard(ctx, static_cast<cudaStream_t>(stream_v));  // the row-block kernel spins on a grid barrier
    char msg[256] = "";
    rc = outer_launch(a, stream_v, msg, sizeof(msg));
    if (rc != B200NORM_OK) return fail(rc, "%s", msg);
    g_launches += (uint64_t)outer_last_launches();
    return B200NORM_OK;
  }

  Plan plan;
  rc = make_plan(d, ctx, kn, aligned, rows_contiguous, &plan);
  if (rc != B200NORM_OK) return rc;

  if (d.rows == 0 || d.cols == 0) {
    // empty iteration or reduction domain: outer sums are zero-filled, nothing else to do
    if (d.cols > 0 && d.rows == 0) {
      const size_t es = elem_size(dt);
      if (dw != nullptr) cudaMemsetAsync(dw, 0, d.cols * es, stream);
      if (db != nullptr) cudaMemsetAsync(db, 0, d.cols * es, stream);
      if (sums_f32 != nullptr) cudaMemsetAsync(sums_f32, 0, (size_t)partial_k(d.op) * d.cols * 4, stream);
    }
    return B200NORM_OK;
  }

  const bool needs_ws = plan.ws_bytes > B200NORM_WS_HEADER_BYTES;
  if (needs_ws) {
    if (workspace == nullptr || workspace_bytes < plan.ws_bytes)
      return fail(B200NORM_ERR_WORKSPACE, "workspace %zu bytes, need %llu", workspace_bytes,
                  (unsigned long long)plan.ws_bytes);
    if ((reinterpret_cast<uintptr_t>(workspace) & 255u) != 0)
      return fail(B200NORM_ERR_WORKSPACE, "workspace must be 256-byte aligned");
    rp.partials = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + B200NORM_WS_HEADER_BYTES);
  }

  rp.groups = plan.groups;
  rp.depth = plan.depth;
  rp.tile_rows = plan.tile_rows;
  rp.cols_cta = plan.cluster > 0 ? (int32_t)(d.cols / plan.cluster) : (int32_t)d.cols;
  rp.nstreams = is_add_fwd(d) ? 2 : n_streams(d.op);
  rp.drop_streams = (dropout && is_norm_bwd(d.op)) ? 1 : 0;
  rp.contiguous = plan.contiguous;
  rp.slice_bytes = plan.slice_bytes;
  rp.stream_bytes = plan.stream_bytes;
  rp.stage_bytes = plan.stage_bytes;
  rp.nchunks = plan.nchunks;
  rp.off_red = plan.off_red;
  rp.off_cls = plan.off_cls;
  rp.off_ring = plan.off_ring;
  rp.off_param = plan.off_param;
  rp.off_acc = plan.off_ring;
  rp.off_tidx = plan.off_tidx;
  rp.off_pipe = plan.off_pipe;
  // Dynamic tile scheduling for the kernels without cross-row state (forward norms, softmax
  // forward / backward): results are bit-identical whatever SM processes a row. The norm
  // backward keeps the static split so that its per-CTA dgamma / dbeta partial sums — and with
  // them the output bits — are the same on every run.
  if (plan.variant == B200NORM_VARIANT_TMA && plan.cluster == 1 && !is_norm_bwd(d.op) &&
      d.op != B200NORM_OP_REDUCE_OUTER && workspace != nullptr &&
      workspace_bytes >= B200NORM_WS_HEADER_BYTES && (reinterpret_cast<uintptr_t>(workspace) & 15u) == 0 &&
      kn.static_sched == 0) {
    // Pool size (measured on several B200s, whose per-SM imbalance differs: the static split loses
    // 5-13 % on the worse ones): the last half of the tiles for two-stream kernels, the last quarter
    // for RMSNorm / softmax forward; LayerNorm forward (two reductions per row, one group per CTA
    // at wide rows) loses more to the pool tiles' ticket wait than it gains and stays static.
    const bool ln_fwd = d.op == B200NORM_OP_LAYERNORM_FWD;  // plain or with the residual add (112 vs 102 us)
    rp.dynamic = ln_fwd ? 0 : (rp.nstreams > 1 ? 2 : 4);
    // short kernels (C2 forward: 28 tiles per group, 47 us) gain nothing: the exit protocol's last
    // atomic round trip and the few pool tiles cost what the balance wins
    const int64_t tiles_per_group = ceil_div64(d.rows, plan.tile_rows) / ((int64_t)plan.grid * plan.groups);
    if (rp.nstreams == 1 && tiles_per_group < 40) rp.dynamic = 0;
    if (int f = kn.pool) rp.dynamic = f > 0 ? f : 4;  // 1 = everything past the first ring fill
    rp.sched_words = reinterpret_cast<uint32_t*>(workspace) + 2;
  }

  switch (dt) {
    case B200NORM_DT_F32: return launch_typed<float>(d, ctx, kn, plan, rp, dw, db, sums_f32, workspace, stream);
    case B200NORM_DT_BF16: return launch_typed<__nv_bfloat16>(d, ctx, kn, plan, rp, dw, db, sums_f32, workspace, stream);
    case B200NORM_DT_F16: return launch_typed<__half>(d, ctx, kn, plan, rp, dw, db, sums_f32, workspace, stream);
    default: return fail(B200NORM_ERR_UNSUPPORTED, "dtype %d", dt);
  }
}

}  // extern "C"

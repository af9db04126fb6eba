// outer.cuh — host interface of the outer-persistent kernels (outer.cu): softmax / log-softmax
// forward and backward with the reduction over dim 0 of a row-major [M, N] matrix.
//
// These are the segments nvFuser's OuterPersistentKernelScheduler takes
// (reference: scheduler/normalization_outer.cpp:451-683; device side
// runtime/fused_welford_impl_outer.cu:222-438, runtime/grid_reduction.cu:905-1066; benchmark
// definitions benchmarks/python/test_softmax_{fwd,bwd}.py with reduction_axis = 0, pytest mark
// `outer_persistent`). The design here is not the reference's register-persistent grid reduction:
// a column slab lives in the shared memory of one CTA or of a thread-block cluster split along M
// (2-D TMA tile loads, DSMEM exchange of the per-column partials) — see outer.cu.
#pragma once

#include <stddef.h>
#include <stdint.h>

namespace b200norm {

struct OuterArgs {
  int bwd = 0;        // 0: forward (in0 = x), 1: backward (in0 = dy, in1 = y)
  int log = 0;        // log-softmax flavour
  int dtype = 0;      // b200norm_dtype
  int64_t M = 0;      // reduced extent (rows)
  int64_t N = 0;      // iteration extent (columns, unit stride)
  const void* in0 = nullptr;
  const void* in1 = nullptr;
  void* out = nullptr;
  int64_t in0_stride = 0, in1_stride = 0, out_stride = 0;  // row strides in elements
  int norm = 0;       // 1: normalization (statistics over dim 0, affine per column): forward, or backward with bwd = 1
                      // (in0 = dy, in1 = x, saved statistics in, dx / dweight / dbias out)
  const void* weight = nullptr;  // [N] io dtype
  const void* bias = nullptr;    // [N] io dtype
  float* mean_out = nullptr;     // [N] fp32
  float* invstd_out = nullptr;   // [N] fp32
  const float* mean_in = nullptr;    // backward: saved mean [N] fp32
  const float* invstd_in = nullptr;  // backward: saved invstd [N] fp32
  void* dw_out = nullptr;            // backward: [N] io dtype
  void* db_out = nullptr;            // backward: [N] io dtype
  double eps = 0.0;
  int round_sum = 0;       // backward: round the column sum through the io dtype (B200NORM_F_ROUND_ROWSUM)
  int cluster_forced = 0;  // 0 = planner's choice, else 1,2,4,8,16
  unsigned long long* timeline = nullptr;  // tuning aid: per-CTA time marks of the row-block kernel
  int force_generic = 0;   // 1: the shape-agnostic family (outer_generic.cuh: stats / merge / apply) whatever the shape
  int force_persistent = 0;  // 1: a persistent family was requested: no fall-back
  void* workspace = nullptr;  // caller's workspace (generic family: fp32 partials behind the 256-byte header)
  size_t workspace_bytes = 0;
  int has_device = 0;
  int sm_count = 148;
  int smem_optin = 232448;
};

struct OuterPlan {
  int cluster = 1;       // CTAs per column slab (split along M)
  int grid = 1;          // CTAs launched (persistent: clusters walk the slabs round-robin)
  int block = 256;
  int slab_bytes = 0;    // W: bytes of one row segment of a slab
  int rows_cta = 0;      // rows of the slab held by one CTA
  int box_rows = 0;      // rows per TMA box (<= 256)
  int depth = 0;         // slab tiles in flight per CTA (1 or 2)
  int ctas_per_sm = 1;
  uint32_t tile_bytes = 0, stage_bytes = 0;
  uint32_t off_wpart = 0, off_cls = 0, off_fin = 0, off_tiles = 0;
  uint32_t smem = 0;
  int nslabs = 0;
};

// Pure host logic (works without a device: nominal B200 numbers). Returns false with a message when
// no persistent plan exists (the reference segments such fusions as well).
bool outer_make_plan(const OuterArgs& a, OuterPlan* plan, char* err, size_t errlen);

// Workspace the shape-agnostic family needs for this problem (header included).
size_t outer_generic_workspace_bytes(const OuterArgs& a);

// Plan of the shape-agnostic family: CTAs per launch and row splits.
void outer_generic_plan(const OuterArgs& a, int* grid_x, int* nsplit, int* rows_per_split);

// Enqueue on `stream`. Returns 0, or a b200norm status code with the message in err. Falls back to the
// shape-agnostic family when no persistent plan exists or the operands are not 16-byte aligned.
int outer_launch(const OuterArgs& a, void* stream, char* err, size_t errlen);

// Kernels the last outer_launch of this thread enqueued (1 persistent, 3 shape-agnostic).
int outer_last_launches();

// The row-block kernel (outer_rows.cuh) takes this segment when its operands are 16-byte aligned and its rows
// contiguous: shape-only test + launch geometry (used by b200norm_query; the launch re-checks the pointers).
bool outer_rows_shape(const OuterArgs& a, int* grid, int* rows_per_iter);
// ... and is preferred to the cluster-persistent family where that one can run too (narrow rows, tens of MB).
bool outer_prefers_rows(const OuterArgs& a);

}  // namespace b200norm

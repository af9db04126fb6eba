// row_pipeline.cuh — the shared skeleton of every "TMA ring" kernel.
//
// One CTA (or one CTA of a cluster) owns a column slice [col0, col0+cols_cta) of
// a strided sequence of row tiles. The CTA is split into `groups` consumer groups
// of TPR threads; group g owns `depth` shared-memory stages and walks the CTA's
// tiles g, g+groups, g+2*groups, ... Each group keeps its own stages full: its
// first warp issues 1-D bulk copies (cp.async.bulk, SASS UBLKCP) that complete on
// a per-stage mbarrier, `depth` tiles ahead, and re-issues into a stage as soon as
// the whole group has finished reading it. There is no producer warp and no
// "empty" barrier: a stage is only ever touched by one group, so the refill is
// ordered by that group's own barrier, and mbarrier phases can never alias.
// Groups reduce with warp shuffles (+ one named barrier when TPR > 32, + a DSMEM
// st.async exchange when the row is split over a thread-block cluster).
//
// This replaces, for the matched segments, the loop nest nvFuser's inner
// persistent schedulers generate (reference: scheduler/normalization_inner_tma.cpp:35-164
// for the TMA variant, runtime/warp.cu:30-98 for the cross-warp reduce,
// runtime/cluster.cu:223-297 for the cluster reduce). The structure is ours:
// bytes in flight are set by groups*depth*tile bytes, not by occupancy or registers.
#pragma once

#include "common.cuh"

namespace b200norm {

// Dynamic shared memory of every ring kernel. Data is read through typed pointers
// into this array (so the compiler emits LDS and may schedule the loads freely);
// barriers and bulk copies use 32-bit shared-window addresses derived from it.
extern __shared__ __align__(128) uint8_t smem_raw[];

__device__ __forceinline__ uint4 ld_smem_u4(uint32_t byte_off) {
  return *reinterpret_cast<const uint4*>(smem_raw + byte_off);
}
__device__ __forceinline__ void st_global_stream(void* p, const uint4& v) {
  __stcs(reinterpret_cast<uint4*>(p), v);
}

constexpr int kMaxWarpsPerGroup = 16;
constexpr int kMaxCluster = 16;
constexpr int kRedNV = 4;  // max floats reduced at once

struct RowParams {
  // streamed operands (row-major, unit inner stride)
  const void* in0;  // x (forward) | dy (backward)
  const void* in1;  // -          | x or y (backward)
  void* out0;       // y (forward) | dx (backward)
  int64_t in0_stride, in1_stride, out_stride;  // row strides in elements
  const void* weight;  // [N] io dtype or nullptr
  const void* bias;    // [N] io dtype or nullptr
  float* stat0;        // fwd: mean out (LN) ; bwd: mean in (LN)
  float* stat1;        // fwd: invstd|rms out ; bwd: invstd|rms in
  float* partials;     // bwd: fp32 [nparts][K][N]
  int64_t rows;
  int32_t cols;
  float eps;
  float inv_n;
  float n_f;
  uint32_t flags;
  // pipeline geometry (filled by the planner)
  int32_t groups;      // consumer groups per CTA
  int32_t depth;       // stages per group (tiles in flight per group)
  int32_t tile_rows;   // rows per tile
  int32_t cols_cta;    // columns handled by this CTA (slice width)
  int32_t nstreams;    // 1 or 2 streamed inputs
  int32_t contiguous;  // tiles are contiguous in memory: one bulk copy per stream
  uint32_t slice_bytes;   // cols_cta * sizeof(T)
  uint32_t stream_bytes;  // tile_rows * slice_bytes (bytes of one stream in a stage)
  uint32_t stage_bytes;   // nstreams * stream_bytes, 128-aligned
  // shared-memory offsets
  uint32_t off_red;     // float [groups][2][kMaxWarpsPerGroup][kRedNV]
  uint32_t off_cls;     // float [groups][2][kMaxCluster][kRedNV]
  uint32_t off_weight;  // fp32 copy of the weight slice, split in 16-byte planes (see load_param)
  uint32_t off_bias;    // same layout
  uint32_t off_ring;    // groups*depth stages; stage index = slot*groups + group
  uint32_t off_acc;     // bwd: cross-group accumulator staging (aliases the ring)
};

// barrier block at the start of dynamic smem:  full[groups*depth] | clbar[groups]
__device__ __forceinline__ uint32_t full_bar(uint32_t smem0, int s) { return smem0 + 8u * s; }
__device__ __forceinline__ uint32_t cl_bar(uint32_t smem0, int nstages, int g) {
  return smem0 + 8u * (nstages + g);
}

// Which rows this CTA walks: cluster `cid` of `nclusters` takes tiles cid, cid+nclusters, ...
struct TileWalk {
  int64_t first_tile;
  int64_t tile_stride;
  int64_t ntiles_local;
  uint32_t crank;
  uint32_t csize;
  __device__ __forceinline__ TileWalk(const RowParams& p) {
    csize = cluster_nctarank();
    crank = cluster_ctarank();
    const int64_t ntiles = ceil_div64(p.rows, p.tile_rows);
    tile_stride = gridDim.x / csize;
    first_tile = blockIdx.x / csize;
    ntiles_local = first_tile < ntiles ? (ntiles - first_tile + tile_stride - 1) / tile_stride : 0;
  }
};

template <typename T>
__device__ __forceinline__ void pipeline_setup(const RowParams& p, uint32_t smem0, const TileWalk& tw) {
  const int nstages = p.groups * p.depth;
  if (threadIdx.x == 0) {
    for (int s = 0; s < nstages; ++s) mbar_init(full_bar(smem0, s), 1);
    for (int g = 0; g < p.groups; ++g) mbar_init(cl_bar(smem0, nstages, g), 1);
    fence_mbar_init();
  }
  // weight / bias slices -> smem as fp32, once per CTA. Layout: VEC/4 planes of
  // [chunk][4 floats] so that lane l reading chunk (c0+l) of a plane is conflict-free.
  constexpr int VEC = IO<T>::VEC;
  const uint32_t slice_chunks = p.slice_bytes / 16;
  const size_t col0_bytes = (size_t)tw.crank * p.slice_bytes;
  for (int which = 0; which < 2; ++which) {
    const void* gp = which == 0 ? p.weight : p.bias;
    if (gp == nullptr) continue;
    const uint32_t dst = (which == 0 ? p.off_weight : p.off_bias);
    const uint4* src = reinterpret_cast<const uint4*>(static_cast<const uint8_t*>(gp) + col0_bytes);
    for (uint32_t c = threadIdx.x; c < slice_chunks; c += blockDim.x) {
      float2 f[VEC / 2];
      IO<T>::unpack2(ldg_128_nc(src + c), f);
#pragma unroll
      for (int h = 0; h < VEC / 4; ++h) {
        *reinterpret_cast<float4*>(smem_raw + dst + (uint32_t)h * slice_chunks * 16u + 16u * c) =
            make_float4(f[2 * h].x, f[2 * h].y, f[2 * h + 1].x, f[2 * h + 1].y);
      }
    }
  }
  __syncthreads();
  if (tw.csize > 1) {
    cluster_sync_all();  // peers' mbarriers are initialised before any st.async targets them
  }
}

// Read chunk c of a staged fp32 parameter (weight / bias) as VEC/2 float2.
template <int VEC>
__device__ __forceinline__ void load_param(uint32_t off, uint32_t slice_chunks, int c, float2 (&w)[VEC / 2]) {
#pragma unroll
  for (int h = 0; h < VEC / 4; ++h) {
    const float4 v = *reinterpret_cast<const float4*>(smem_raw + off + (uint32_t)h * slice_chunks * 16u + 16u * c);
    w[2 * h] = make_float2(v.x, v.y);
    w[2 * h + 1] = make_float2(v.z, v.w);
  }
}

// Consumer-group context: identity of the thread inside its group and the state of
// the two reduction exchanges (intra-CTA slots, cluster slots).
template <int TPR>
struct Group {
  static constexpr int WPG = TPR / 32;
  int gid;
  int gtid;
  int lane;
  int warp;  // warp index inside the group
  uint32_t red_off;   // byte offset of this group's intra-CTA slots in smem_raw
  uint32_t cls_off;   // byte offset of this group's cluster slots in smem_raw
  uint32_t cls_addr;  // the same as a shared-window address (for mapa / st.async)
  uint32_t clbar;
  uint32_t red_phase;
  uint32_t cl_phase;
  uint32_t csize;
  uint32_t crank;

  __device__ __forceinline__ Group(const RowParams& p, uint32_t smem0, const TileWalk& tw) {
    gid = threadIdx.x / TPR;
    gtid = threadIdx.x % TPR;
    lane = threadIdx.x & 31;
    warp = gtid >> 5;
    red_off = p.off_red + (uint32_t)gid * (2 * kMaxWarpsPerGroup * kRedNV * 4);
    cls_off = p.off_cls + (uint32_t)gid * (2 * kMaxCluster * kRedNV * 4);
    cls_addr = smem0 + cls_off;
    clbar = cl_bar(smem0, p.groups * p.depth, gid);
    red_phase = 0;
    cl_phase = 0;
    csize = tw.csize;
    crank = tw.crank;
  }

  __device__ __forceinline__ void sync() const {
    if constexpr (TPR == 32) {
      __syncwarp();
    } else {
      named_bar_sync(1 + gid, TPR);
    }
  }

  // All-reduce NV floats across the group's TPR threads and, when the row is
  // split over a cluster, across the peer CTAs' same-numbered groups. Every
  // thread returns the same bits (fixed combination order). Acts as a barrier
  // for the group: on return every thread of the group has executed everything
  // that precedes the call.
  template <typename Op, int NV>
  __device__ __forceinline__ void allreduce(float (&v)[NV]) {
    static_assert(NV <= kRedNV, "too many values");
    warp_allreduce<Op, NV>(v);
    if constexpr (TPR > 32) {
      // double-buffered slots: one named barrier per reduction is enough (a slot set is
      // rewritten only two reductions later, after every reader passed the next barrier)
      float* slots = reinterpret_cast<float*>(smem_raw + red_off + red_phase * (kMaxWarpsPerGroup * kRedNV * 4));
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) slots[warp * kRedNV + i] = v[i];
      }
      named_bar_sync(1 + gid, TPR);
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] = Op::identity();
#pragma unroll
      for (int w = 0; w < WPG; ++w) {
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = Op::apply(v[i], slots[w * kRedNV + i]);
      }
      red_phase ^= 1u;
    }
    if (csize > 1) {
      const uint32_t par = cl_phase & 1u;
      const uint32_t set = par * (kMaxCluster * kRedNV * 4);
      if (gtid == 0) {
        mbar_arrive_expect_tx(clbar, csize * NV * 4);
        for (uint32_t r = 0; r < csize; ++r) {
          const uint32_t rbar = map_to_cta(clbar, r);
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            st_async_f32(map_to_cta(cls_addr + set + (crank * kRedNV + i) * 4, r), v[i], rbar);
          }
        }
      }
      mbar_wait(clbar, par);
      const volatile float* slots = reinterpret_cast<const volatile float*>(smem_raw + cls_off + set);
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] = Op::identity();
      for (uint32_t r = 0; r < csize; ++r) {
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = Op::apply(v[i], slots[r * kRedNV + i]);
      }
      cl_phase += 1u;
    }
  }
};

// The group's private ring. `k` counts the group's tiles; tile k of group g is the
// CTA's local tile g + k*groups and lives in stage (k % depth)*groups + g.
template <typename T, int TPR>
struct Ring {
  const RowParams& p;
  const TileWalk& tw;
  const Group<TPR>& g;
  uint32_t smem0;
  int64_t ntiles_g;
  uint64_t policy;

  __device__ __forceinline__ Ring(const RowParams& p_, uint32_t smem0_, const TileWalk& tw_, const Group<TPR>& g_)
      : p(p_), tw(tw_), g(g_), smem0(smem0_) {
    ntiles_g = tw.ntiles_local > g.gid ? (tw.ntiles_local - g.gid + p.groups - 1) / p.groups : 0;
    policy = l2_policy_evict_first();
  }
  __device__ __forceinline__ int stage_of(int64_t k) const { return (int)(k % p.depth) * p.groups + g.gid; }
  __device__ __forceinline__ int64_t row0_of(int64_t k) const {
    return (tw.first_tile + (g.gid + k * p.groups) * tw.tile_stride) * p.tile_rows;
  }
  __device__ __forceinline__ int nrows_of(int64_t row0) const {
    return (int)min((int64_t)p.tile_rows, p.rows - row0);
  }

  // Issue the bulk copies of the group's tile k. Executed by the group's first warp
  // (all 32 lanes, converged); lane 0 arms the barrier, lanes share per-row copies.
  __device__ __forceinline__ void issue(int64_t k) const {
    const int s = stage_of(k);
    const int64_t row0 = row0_of(k);
    const int nrows = nrows_of(row0);
    const uint32_t bytes = (uint32_t)nrows * p.slice_bytes;
    const uint32_t dst = smem0 + p.off_ring + (uint32_t)s * p.stage_bytes;
    const uint32_t fb = full_bar(smem0, s);
    const size_t col0_bytes = (size_t)g.crank * p.slice_bytes;
    const uint8_t* base0 = static_cast<const uint8_t*>(p.in0) + col0_bytes;
    const uint8_t* base1 = static_cast<const uint8_t*>(p.in1) + col0_bytes;
    // No proxy fence is needed for this write-after-read: every reader consumed its loads
    // (the values fed the reduction) before the group barrier that precedes refill().
    if (g.lane == 0) mbar_arrive_expect_tx(fb, bytes * (uint32_t)p.nstreams);
    __syncwarp();
    if (p.contiguous) {
      if (g.lane == 0) {
        bulk_g2s(dst, base0 + (size_t)row0 * p.in0_stride * sizeof(T), bytes, fb, policy);
        if (p.nstreams > 1) {
          bulk_g2s(dst + p.stream_bytes, base1 + (size_t)row0 * p.in1_stride * sizeof(T), bytes, fb, policy);
        }
      }
    } else {
      for (int r = g.lane; r < nrows; r += 32) {
        bulk_g2s(dst + (uint32_t)r * p.slice_bytes, base0 + (size_t)(row0 + r) * p.in0_stride * sizeof(T),
                 p.slice_bytes, fb, policy);
        if (p.nstreams > 1) {
          bulk_g2s(dst + p.stream_bytes + (uint32_t)r * p.slice_bytes,
                   base1 + (size_t)(row0 + r) * p.in1_stride * sizeof(T), p.slice_bytes, fb, policy);
        }
      }
    }
  }

  // Fill the ring before the first tile.
  __device__ __forceinline__ void prologue() const {
    if (g.warp == 0) {
      for (int64_t k = 0; k < p.depth && k < ntiles_g; ++k) issue(k);
    }
  }

  // Refill the stage of tile k with tile k+depth. PRECONDITION: every thread of the
  // group has finished reading the stage and a group-wide barrier (Group::allreduce
  // or Group::sync) separates those reads from this call.
  __device__ __forceinline__ void refill(int64_t k) const {
    if (g.warp == 0 && k + p.depth < ntiles_g) issue(k + p.depth);
  }

  __device__ __forceinline__ void wait(int64_t k) const {
    mbar_wait(full_bar(smem0, stage_of(k)), (uint32_t)((k / p.depth) & 1));
  }
  __device__ __forceinline__ uint32_t stage_off(int64_t k) const {
    return p.off_ring + (uint32_t)stage_of(k) * p.stage_bytes;
  }
};

}  // namespace b200norm

// row_pipeline.cuh — the shared skeleton of every "TMA ring" kernel.
//
// One CTA (or one CTA of a cluster) owns a column slice [col0, col0+cols_cta) of
// a strided sequence of row tiles. A single producer lane feeds a ring of
// `stages` shared-memory buffers with 1-D bulk copies (cp.async.bulk, SASS
// UBLKCP) that complete on per-stage "full" mbarriers; `groups` consumer groups
// of TPR threads each drain every groups-th tile, reduce with warp shuffles
// (+ a named barrier when TPR > 32, + a DSMEM st.async exchange when the row is
// split over a cluster), and release the stage through an "empty" mbarrier.
//
// This replaces, for the matched segments, the loop nest nvFuser's inner
// persistent schedulers generate (reference: scheduler/normalization_inner_tma.cpp:35-164
// for the TMA variant, runtime/warp.cu:30-98 for the cross-warp reduce,
// runtime/cluster.cu:223-297 for the cluster reduce). The structure is ours:
// bytes in flight are set by the ring depth, not by occupancy or registers.
#pragma once

#include "common.cuh"

namespace b200norm {

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
  int32_t stages;      // ring depth
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
  uint32_t off_weight;  // cols_cta * sizeof(T)
  uint32_t off_bias;
  uint32_t off_ring;
  uint32_t off_acc;     // bwd: cross-group accumulator staging (aliases the ring)
};

// barrier block at the start of dynamic smem:
//   full[stages] | empty[stages] | clbar[groups]
__device__ __forceinline__ uint32_t full_bar(uint32_t smem0, int s) { return smem0 + 8u * s; }
__device__ __forceinline__ uint32_t empty_bar(uint32_t smem0, int stages, int s) {
  return smem0 + 8u * (stages + s);
}
__device__ __forceinline__ uint32_t cl_bar(uint32_t smem0, int stages, int g) {
  return smem0 + 8u * (2 * stages + g);
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
__device__ __forceinline__ void pipeline_setup(const RowParams& p, uint32_t smem0, int warps_per_group,
                                               const TileWalk& tw) {
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(smem0, s), 1);
      mbar_init(empty_bar(smem0, p.stages, s), warps_per_group);
    }
    for (int g = 0; g < p.groups; ++g) {
      mbar_init(cl_bar(smem0, p.stages, g), 1);
    }
    fence_mbar_init();
  }
  // weight / bias slices -> smem (plain loads; once per CTA)
  const uint32_t slice_chunks = p.slice_bytes / 16;
  const size_t col0_bytes = (size_t)tw.crank * p.slice_bytes;
  if (p.weight != nullptr) {
    const uint4* src = reinterpret_cast<const uint4*>(static_cast<const uint8_t*>(p.weight) + col0_bytes);
    for (uint32_t c = threadIdx.x; c < slice_chunks; c += blockDim.x) {
      sts_128(smem0 + p.off_weight + 16u * c, ldg_128_nc(src + c));
    }
  }
  if (p.bias != nullptr) {
    const uint4* src = reinterpret_cast<const uint4*>(static_cast<const uint8_t*>(p.bias) + col0_bytes);
    for (uint32_t c = threadIdx.x; c < slice_chunks; c += blockDim.x) {
      sts_128(smem0 + p.off_bias + 16u * c, ldg_128_nc(src + c));
    }
  }
  __syncthreads();
  if (tw.csize > 1) {
    cluster_sync_all();  // peers' mbarriers are initialised before any st.async targets them
  }
}

// Producer: executed by ONE warp; lane 0 owns the barrier protocol, all lanes
// issue per-row copies when tiles are not contiguous.
template <typename T>
__device__ __forceinline__ void producer_loop(const RowParams& p, uint32_t smem0, const TileWalk& tw) {
  const int lane = threadIdx.x & 31;
  const uint64_t pol = l2_policy_evict_first();
  const size_t col0_bytes = (size_t)tw.crank * p.slice_bytes;
  const uint8_t* base0 = static_cast<const uint8_t*>(p.in0) + col0_bytes;
  const uint8_t* base1 = static_cast<const uint8_t*>(p.in1) + col0_bytes;
  int s = 0;
  uint32_t ph = 0;
  for (int64_t i = 0; i < tw.ntiles_local; ++i) {
    const int64_t row0 = (tw.first_tile + i * tw.tile_stride) * p.tile_rows;
    const int nrows = (int)min((int64_t)p.tile_rows, p.rows - row0);
    const uint32_t bytes = (uint32_t)nrows * p.slice_bytes;
    const uint32_t dst = smem0 + p.off_ring + (uint32_t)s * p.stage_bytes;
    const uint32_t fb = full_bar(smem0, s);
    if (lane == 0) {
      mbar_wait(empty_bar(smem0, p.stages, s), ph ^ 1u);
      mbar_arrive_expect_tx(fb, bytes * (uint32_t)p.nstreams);
    }
    __syncwarp();
    if (p.contiguous) {
      if (lane == 0) {
        bulk_g2s(dst, base0 + (size_t)row0 * p.in0_stride * sizeof(T), bytes, fb, pol);
        if (p.nstreams > 1) {
          bulk_g2s(dst + p.stream_bytes, base1 + (size_t)row0 * p.in1_stride * sizeof(T), bytes, fb, pol);
        }
      }
    } else {
      for (int r = lane; r < nrows; r += 32) {
        bulk_g2s(dst + (uint32_t)r * p.slice_bytes,
                 base0 + (size_t)(row0 + r) * p.in0_stride * sizeof(T), p.slice_bytes, fb, pol);
        if (p.nstreams > 1) {
          bulk_g2s(dst + p.stream_bytes + (uint32_t)r * p.slice_bytes,
                   base1 + (size_t)(row0 + r) * p.in1_stride * sizeof(T), p.slice_bytes, fb, pol);
        }
      }
    }
    if (++s == p.stages) {
      s = 0;
      ph ^= 1u;
    }
  }
}

// Consumer-group context: identity of the thread inside its group plus the
// state of the two reduction exchanges (intra-CTA slots, cluster slots).
template <int TPR>
struct Group {
  static constexpr int WPG = TPR / 32;
  int gid;
  int gtid;
  int lane;
  int warp;  // warp index inside the group
  uint32_t red_base;
  uint32_t cls_base;
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
    red_base = smem0 + p.off_red + (uint32_t)gid * (2 * kMaxWarpsPerGroup * kRedNV * 4);
    cls_base = smem0 + p.off_cls + (uint32_t)gid * (2 * kMaxCluster * kRedNV * 4);
    clbar = cl_bar(smem0, p.stages, gid);
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
  // thread returns the same bits (fixed combination order).
  template <typename Op, int NV>
  __device__ __forceinline__ void allreduce(float (&v)[NV]) {
    static_assert(NV <= kRedNV, "too many values");
    warp_allreduce<Op, NV>(v);
    if constexpr (TPR > 32) {
      const uint32_t slots = red_base + red_phase * (kMaxWarpsPerGroup * kRedNV * 4);
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(slots + (uint32_t)(warp * kRedNV + i) * 4), "f"(v[i])
                       : "memory");
        }
      }
      named_bar_sync(1 + gid, TPR);
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] = Op::identity();
#pragma unroll
      for (int w = 0; w < WPG; ++w) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          float t;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t) : "r"(slots + (uint32_t)(w * kRedNV + i) * 4));
          v[i] = Op::apply(v[i], t);
        }
      }
      red_phase ^= 1u;
    }
    if (csize > 1) {
      const uint32_t par = cl_phase & 1u;
      const uint32_t slots = cls_base + par * (kMaxCluster * kRedNV * 4);
      if (gtid == 0) {
        mbar_arrive_expect_tx(clbar, csize * NV * 4);
        for (uint32_t r = 0; r < csize; ++r) {
          const uint32_t rbar = map_to_cta(clbar, r);
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            st_async_f32(map_to_cta(slots + (crank * kRedNV + i) * 4, r), v[i], rbar);
          }
        }
      }
      mbar_wait(clbar, par);
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] = Op::identity();
      for (uint32_t r = 0; r < csize; ++r) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          float t;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t) : "r"(slots + (r * kRedNV + i) * 4));
          v[i] = Op::apply(v[i], t);
        }
      }
      cl_phase += 1u;
    }
  }
};

// Iterate the tiles of one consumer group. Body signature:
//   body(row0, nrows, stage_smem_addr)
// The stage is released after body returns (every warp of the group arrives).
template <int TPR, typename Body>
__device__ __forceinline__ void consume_tiles(const RowParams& p, uint32_t smem0, const TileWalk& tw,
                                              const Group<TPR>& g, Body&& body) {
  for (int64_t i = g.gid; i < tw.ntiles_local; i += p.groups) {
    const int s = (int)(i % p.stages);
    const uint32_t ph = (uint32_t)((i / p.stages) & 1);
    mbar_wait(full_bar(smem0, s), ph);
    const int64_t row0 = (tw.first_tile + i * tw.tile_stride) * p.tile_rows;
    const int nrows = (int)min((int64_t)p.tile_rows, p.rows - row0);
    body(row0, nrows, smem0 + p.off_ring + (uint32_t)s * p.stage_bytes);
    __syncwarp();
    if (g.lane == 0) {
      mbar_arrive(empty_bar(smem0, p.stages, s));
    }
  }
}

}  // namespace b200norm

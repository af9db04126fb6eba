// ring.cuh — the shared skeleton of every "TMA ring" kernel (sm_100a).
//
// One CTA (or one CTA of a cluster) owns a column slice [col0, col0+cols_cta) of a
// strided sequence of row tiles. The CTA is split into G consumer groups of TPR
// threads; group g owns D shared-memory stages and walks the tiles
// cid + (g + k*G)*nclusters, k = 0,1,... Each group keeps its own stages full: its
// first warp issues 1-D bulk copies (cp.async.bulk, SASS UBLKCP) that complete on a
// per-stage mbarrier, D tiles ahead, and re-issues into a stage as soon as the whole
// group has finished reading it. There is no producer warp and no "empty" barrier:
// a stage is only ever touched by one group, so the refill is ordered by that group's
// own barrier, and mbarrier phases can never alias.
// Groups reduce with warp shuffles (+ one named barrier when TPR > 32, + a DSMEM
// st.async exchange when the row is split over a thread-block cluster).
//
// Everything the hot loop needs is a running register (stage offset, barrier address,
// phase bit, row index): there is no division, modulo or 64-bit multiply per row.
// Chunk offsets inside a row are compile-time immediates (j*TPR*16 bytes).
//
// This replaces, for the matched segments, the loop nest nvFuser's inner persistent
// schedulers generate (reference: scheduler/normalization_inner_tma.cpp:35-164 for the
// TMA variant, runtime/warp.cu:30-98 for the cross-warp reduce, runtime/cluster.cu:223-297
// for the cluster reduce). The structure is ours: bytes in flight are set by
// G*D*tile bytes, not by occupancy or registers.
#pragma once

#include "common.cuh"

namespace b200norm {

// Dynamic shared memory of every ring kernel.
extern __shared__ __align__(128) uint8_t smem_raw[];

constexpr int kMaxWarpsPerGroup = 16;
constexpr int kMaxCluster = 16;
constexpr int kRedNV = 4;  // max floats reduced at once
constexpr int kMaxPeers = 8;         // GPUs of one box
constexpr int kPeerHeaderBytes = 256;  // uint32 published[16] | uint32 consumed[16] | pad
constexpr int kPeerConsumedOffset = 64;
constexpr int kSchedTableInts = 32;  // stage -> tile index table of the dynamic schedule (= max stages)
constexpr int kRedSetBytes = kRedNV * kMaxWarpsPerGroup * 4;  // one slot set: [kRedNV][16 warps]
constexpr int kClsSetBytes = kRedNV * kMaxCluster * 4;        // one cluster slot set: [16 ranks][kRedNV]
constexpr int kMaxBlock = 512;  // G * TPR; 128 registers per thread at one CTA per SM
// (two CTAs per SM at 64 registers was measured: every kernel spills and loses 10-60 %)

constexpr uint32_t kFlagHasWeight = 1u << 0;
constexpr uint32_t kFlagHasBias = 1u << 1;
constexpr uint32_t kFlagStatIsRms = 1u << 2;
constexpr uint32_t kFlagRmsBwdCentered = 1u << 3;
constexpr uint32_t kFlagOutDx = 1u << 4;
constexpr uint32_t kFlagOutDw = 1u << 5;
constexpr uint32_t kFlagOutDb = 1u << 6;
constexpr uint32_t kFlagPartialsF32 = 1u << 7;
constexpr uint32_t kFlagRoundRowsum = 1u << 8;
constexpr uint32_t kFlagReduceMax = 1u << 9;
constexpr uint32_t kFlagReduceFmax = 1u << 10;
constexpr uint32_t kFlagOutMean = 1u << 11;
constexpr uint32_t kFlagOutInvstd = 1u << 12;
constexpr uint32_t kFlagAddResidual = 1u << 13;
constexpr uint32_t kFlagOutSum = 1u << 14;
constexpr uint32_t kFlagRoundSum = 1u << 15;
constexpr uint32_t kFlagNoY = 1u << 20;
constexpr uint32_t kFlagStatVar = 1u << 21;
constexpr uint32_t kFlagStatM2 = 1u << 22;
constexpr uint32_t kFlagDropout = 1u << 23;
constexpr uint32_t kFlagReduceNll = 1u << 24;
constexpr uint32_t kFlagRowAffine = 1u << 25;  // instance norm: weight / bias indexed by (row mod channels)

struct RowParams {
  // streamed operands (row-major, unit inner stride)
  const void* in0;  // x (forward) | dy (backward)
  const void* in1;  // -          | x or y (backward)
  void* out0;       // y (forward) | dx (backward)
  void* out1;       // forward with a fused residual add: x + residual (nullptr = not requested)
  int64_t in0_stride, in1_stride, out_stride, out1_stride;  // row strides in elements
  // fused dropout (kFlagDropout): forward norm(residual + dropout(x)) draws the mask and writes it out; backward
  // reads it back, rebuilds residual + dropout(x) and produces both input gradients (out0 = d x, out1 = d residual)
  const void* in2;              // backward: residual
  int64_t in2_stride;
  const uint8_t* mask_in;       // backward: [M,N] bytes, contiguous rows
  uint8_t* mask_out;            // forward:  [M,N] bytes, contiguous rows
  const unsigned long long* rng;  // forward: device {seed, offset} of the Philox stream
  uint32_t drop_t1, drop_t2;    // keep <=> word <= t1 || word >= t2
  float drop_scale;             // 1 / keep probability
  int32_t drop_streams;         // backward ring kernels: every stage also carries the residual slice and the mask bytes
  // instance norm (kFlagRowAffine): rows are (sample, channel) pairs, the affine parameters belong to the channel;
  // backward: the per-row sums of dy and dy * x_hat go out instead of column sums (the caller adds them over the batch)
  int32_t channels;
  float* rowsum_dy;     // [M] fp32 or nullptr
  float* rowsum_dyxh;   // [M] fp32 or nullptr
  // cross-entropy (kFlagReduceNll): per-row targets; stat0 = lse out, stat1 = loss out
  const long long* target;
  int32_t ignore_index;
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
  float var_scale;   // LayerNorm forward with kFlagStatVar: 1 / max(N - correction, 0)
  float proj_scale;  // norm backward: factor on the projection term (1, or 1/2 for B200NORM_F_RMS_BWD_HALF_PROJ)
  uint32_t flags;
  // pipeline geometry (filled by the planner)
  int32_t groups;      // consumer groups per CTA
  int32_t depth;       // stages per group (tiles in flight per group)
  int32_t tile_rows;   // rows per tile
  int32_t cols_cta;    // columns handled by this CTA (slice width)
  int32_t nchunks;     // 16-byte chunks per row slice
  int32_t nstreams;    // 1 or 2 streamed inputs
  int32_t contiguous;  // tiles are contiguous in memory: one bulk copy per stream
  uint32_t slice_bytes;   // cols_cta * sizeof(T)
  uint32_t stream_bytes;  // tile_rows * slice_bytes (bytes of one stream in a stage)
  uint32_t stage_bytes;   // nstreams * stream_bytes, 128-aligned
  // shared-memory offsets
  uint32_t off_red;   // float [groups][2][kRedNV][kMaxWarpsPerGroup]
  uint32_t off_cls;   // float [groups][2][kMaxCluster][kRedNV]
  uint32_t off_ring;  // groups*depth stages, group-major: stage (g*depth + s)
  uint32_t off_acc;   // bwd: cross-group accumulator staging (aliases the ring)
  uint32_t off_param; // LayerNorm fwd: fp32 copies of the weight / bias slices, 16-byte planes
  // backward: in-kernel finish of the column sums (cooperative launch; see kernels_bwd.cuh)
  int32_t fused_finish;    // 1: the grid is co-resident and finishes dgamma / dbeta itself
  uint32_t* sync_words;    // two zero-initialised workspace header words, returned to zero
  void* dw_out;            // io dtype [N] or nullptr
  void* db_out;            // io dtype [N] or nullptr
  float* sums_out;         // fp32 [K][N] or nullptr (multi-GPU partial sums, no cast)
  // Row-sharded backward over several GPUs of one box (B200NORM_F_PEER_REDUCE): the finishing CTAs
  // exchange the fp32 column sums through peer memory (NVLink / NVSwitch P2P) and every rank
  // produces the cast dgamma / dbeta itself — the all-reduce is part of this kernel.
  // peer_xchg[q] = rank q's exchange buffer: 16 uint32 epoch flags (one per source rank), padded to
  // kPeerHeaderBytes, then two [K][N] fp32 slots used alternately by consecutive launches.
  int32_t peer_world;  // 0 = single GPU
  int32_t peer_rank;
  int32_t peer_deferred;  // 1: publish only (B200NORM_F_PEER_PUBLISH); the combine is a later launch
  uint32_t peer_epoch;  // launch sequence number, identical on every rank, starts at 1
  uint32_t peer_slot_floats;  // capacity of one slot in floats: distance between the two publish slots, and between
                              // the per-source-rank areas of the fused flavour's receive region
  uint32_t peer_ll_off;           // fused flavour: byte offset of the (value, epoch) receive region in every buffer
  uint32_t peer_ll_parity_bytes;  // ... and bytes of one of its two parities (world * peer_slot_floats * 8)
  uint64_t peer_timeout_ns;       // limit of an in-kernel wait for a peer (0 = for ever)
  uint8_t* peer_xchg[kMaxPeers];
  // tuning aid (b200norm_debug_set_timeline): per-CTA %globaltimer marks, kTimelineSlots per CTA; nullptr = off
  unsigned long long* timeline;
  // dynamic tile scheduling (forward-type kernels, cluster size 1): the last quarter of the tiles
  // is drawn from a global counter so that SMs which stream faster take more rows (per-SM
  // bandwidth differs by up to ~35 % on B200; with a static split the slowest SM sets the kernel
  // time). sched_words = {next pool tile, CTAs done}; both are zero between launches (the last
  // CTA to finish resets them).
  int32_t dynamic;
  uint32_t* sched_words;
  uint32_t off_tidx;  // int32 [groups*depth]: tile index held by each stage (-1 = no more tiles)
  uint32_t off_pipe;  // pipelined backward: mbarriers red[2*groups] | free[groups*depth] | clusterB[groups]
};

// Timeline marks: 0 entry, 1 setup done (after griddepcontrol.wait), 2 first tile landed,
// 3 row loop done, 4 partial row written, 5 grid barrier passed, 6 exit, 7 (the SM id), 8 column sums
// finished (and pushed to the peers), 9 peers' sums collected.
constexpr int kTimelineSlots = 12;
__device__ __forceinline__ void tl_mark(const RowParams& p, int slot) {
  if (p.timeline != nullptr && threadIdx.x == 0) {
    p.timeline[(size_t)blockIdx.x * kTimelineSlots + slot] = global_timer_ns();
    if (slot == 0) {
      uint32_t smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      p.timeline[(size_t)blockIdx.x * kTimelineSlots + 7] = smid;  // slot 7: the SM this CTA ran on
    }
  }
}

// barrier block at the start of dynamic smem:  full[groups*depth] | clbar[groups]

// Welford pairs of M equal-count partial sets, stored as means[M] | m2s[M] (16-byte aligned when M >= 4), combined
// in closed form: mean = avg(mean_i), M2 = sum(M2_i) + n_each * sum((mean_i - mean)^2). Fixed order: every
// caller gets the same bits. Vector loads and packed adds: this runs once per row in every thread.
template <int M>
__device__ __forceinline__ void welford_combine_equal(const float* means, const float* m2s, float n_each, float& mean,
                                                      float& m2) {
  static_assert(M == 2 || M == 4 || M == 8 || M == 16, "partial sets: 2, 4, 8 or 16");
  if constexpr (M == 2) {
    const float2 a = *reinterpret_cast<const float2*>(means);
    const float2 b = *reinterpret_cast<const float2*>(m2s);
    mean = (a.x + a.y) * 0.5f;
    const float d = a.x - a.y;
    m2 = fmaf(d * d, 0.5f * n_each, b.x + b.y);
  } else {
    float2 mv[M / 2];
    float2 sm = f2(0.f), s2 = f2(0.f);
#pragma unroll
    for (int q = 0; q < M / 4; ++q) {
      const float4 a = *reinterpret_cast<const float4*>(means + 4 * q);
      const float4 b = *reinterpret_cast<const float4*>(m2s + 4 * q);
      mv[2 * q] = make_float2(a.x, a.y);
      mv[2 * q + 1] = make_float2(a.z, a.w);
      sm = f2add(sm, f2add(mv[2 * q], mv[2 * q + 1]));
      s2 = f2add(s2, f2add(make_float2(b.x, b.y), make_float2(b.z, b.w)));
    }
    mean = (sm.x + sm.y) * (1.0f / M);
    const float2 nm = f2(-mean);
    float2 dev = f2(0.f);
#pragma unroll
    for (int q = 0; q < M / 2; ++q) {
      const float2 d = f2add(mv[q], nm);
      dev = f2fma(d, d, dev);
    }
    m2 = fmaf(dev.x + dev.y, n_each, s2.x + s2.y);
  }
}

// ---------------------------------------------------------------------------
// Consumer-group context: identity of the thread inside its group and the state of
// the two reduction exchanges (intra-CTA slots, cluster slots).
// ---------------------------------------------------------------------------
template <int TPR>
struct Group {
  static constexpr int WPG = TPR / 32;
  int gid;
  int gtid;
  int lane;
  int warp;            // warp index inside the group
  uint32_t red_off;    // byte offset of this group's current slot set in smem_raw
  uint32_t cls_off;    // byte offset of this group's cluster slot sets in smem_raw
  uint32_t cls_addr;   // the same as a shared-window address (for mapa / st.async)
  uint32_t clbar;
  uint32_t cl_phase;
  uint32_t csize;
  uint32_t crank;

  __device__ __forceinline__ Group(const RowParams& p, uint32_t smem0) {
    gid = threadIdx.x / TPR;
    gtid = threadIdx.x % TPR;
    lane = threadIdx.x & 31;
    warp = gtid >> 5;
    red_off = p.off_red + (uint32_t)gid * (2 * kRedSetBytes);
    cls_off = p.off_cls + (uint32_t)gid * (2 * kClsSetBytes);
    cls_addr = smem0 + cls_off;
    clbar = smem0 + 8u * (uint32_t)(p.groups * p.depth + gid);
    cl_phase = 0;
    csize = cluster_nctarank();
    crank = cluster_ctarank();
  }

  __device__ __forceinline__ void sync() const {
    if constexpr (TPR == 32) {
      __syncwarp();
    } else {
      named_bar_sync(1 + gid, TPR);
    }
  }

  // Welford triples (count, mean, M2) of arbitrary partial sets -> the triple of the whole row slice of the group
  // (and of the cluster), merged with Chan's formula (runtime/welford.cu:52-68 is the reference's combine):
  //     n = na + nb,  mean = ma + (mb - ma) * nb / n,  M2 = M2a + M2b + (mb - ma)^2 * na * nb / n.
  // Lanes merge by xor shuffles, warps and cluster ranks sequentially in index order from shared-memory slots, so
  // every thread ends with the same bits. Acts as a barrier for the group, like allreduce().
  __device__ __forceinline__ static void chan_merge(float& n, float& mean, float& m2, float on, float om, float o2) {
    const float tot = n + on;
    const float w = tot > 0.f ? on * rcp_fast(tot) : 0.f;
    const float d = om - mean;
    mean = fmaf(d, w, mean);
    m2 = fmaf(d * d, n * w, m2 + o2);
    n = tot;
  }
  __device__ __forceinline__ void welford_allreduce(float& n, float& mean, float& m2) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float on = __shfl_xor_sync(0xffffffffu, n, off);
      const float om = __shfl_xor_sync(0xffffffffu, mean, off);
      const float o2 = __shfl_xor_sync(0xffffffffu, m2, off);
      chan_merge(n, mean, m2, on, om, o2);
    }
    if constexpr (TPR > 32) {
      float* slots = reinterpret_cast<float*>(smem_raw + red_off);
      if (lane == 0) {
        slots[warp] = n;
        slots[kMaxWarpsPerGroup + warp] = mean;
        slots[2 * kMaxWarpsPerGroup + warp] = m2;
      }
      named_bar_sync(1 + gid, TPR);
      n = slots[0];
      mean = slots[kMaxWarpsPerGroup];
      m2 = slots[2 * kMaxWarpsPerGroup];
#pragma unroll
      for (int w = 1; w < WPG; ++w) {
        chan_merge(n, mean, m2, slots[w], slots[kMaxWarpsPerGroup + w], slots[2 * kMaxWarpsPerGroup + w]);
      }
      red_off ^= (uint32_t)kRedSetBytes;
    } else {
      // one warp: take lane 0's triple so that every lane holds the same bits
      n = __shfl_sync(0xffffffffu, n, 0);
      mean = __shfl_sync(0xffffffffu, mean, 0);
      m2 = __shfl_sync(0xffffffffu, m2, 0);
    }
    if (csize > 1) {
      const uint32_t par = cl_phase & 1u;
      const uint32_t set = par * kClsSetBytes;
      if (gtid == 0) mbar_arrive_expect_tx(clbar, csize * 3 * 4);
      if ((uint32_t)gtid < csize) {
        const uint32_t r = (uint32_t)gtid;
        const uint32_t rbar = map_to_cta(clbar, r);
        st_async_f32(map_to_cta(cls_addr + set + (crank * kRedNV + 0) * 4, r), n, rbar);
        st_async_f32(map_to_cta(cls_addr + set + (crank * kRedNV + 1) * 4, r), mean, rbar);
        st_async_f32(map_to_cta(cls_addr + set + (crank * kRedNV + 2) * 4, r), m2, rbar);
      }
      mbar_wait(clbar, par);
      const volatile float* slots = reinterpret_cast<const volatile float*>(smem_raw + cls_off + set);
      n = slots[0];
      mean = slots[1];
      m2 = slots[2];
      for (uint32_t r = 1; r < csize; ++r) {
        chan_merge(n, mean, m2, slots[r * kRedNV + 0], slots[r * kRedNV + 1], slots[r * kRedNV + 2]);
      }
      cl_phase += 1u;
    }
  }

  // Online-softmax pairs (m = maximum, s = sum of exp(x - m) over a partial set) -> the pair of the whole row:
  //     m = max(ma, mb),   s = sa * exp(ma - m) + sb * exp(mb - m)
  // in ONE exchange (the two-reduction form — maximum, then sum — costs a row split over a cluster two DSMEM round
  // trips; the pair costs one). The maximum is exact, so every level takes the maximum first and rescales each
  // partial sum once; equal maxima (also -inf with -inf: an all -inf partial set has s = 0) scale by exactly 1, NaN
  // propagates through MaxOp. Fixed order: every thread ends with the same bits. A barrier for the group.
  template <typename T>
  __device__ __forceinline__ static float softmax_rescale(float m_part, float m_all) {
    return m_part == m_all ? 1.f : exp_for<T>(m_part - m_all);
  }
  template <typename T>
  __device__ __forceinline__ void softmax_allreduce(float& m, float& s) {
    float wm = m;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) wm = MaxOp::apply(wm, __shfl_xor_sync(0xffffffffu, wm, off));
    s *= softmax_rescale<T>(m, wm);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    m = wm;
    if constexpr (TPR > 32) {
      float* slots = reinterpret_cast<float*>(smem_raw + red_off);
      if (lane == 0) {
        slots[warp] = m;
        slots[kMaxWarpsPerGroup + warp] = s;
      }
      named_bar_sync(1 + gid, TPR);
      float gm = slots[0];
#pragma unroll
      for (int w = 1; w < WPG; ++w) gm = MaxOp::apply(gm, slots[w]);
      float gs = 0.f;
#pragma unroll
      for (int w = 0; w < WPG; ++w) gs = fmaf(slots[kMaxWarpsPerGroup + w], softmax_rescale<T>(slots[w], gm), gs);
      m = gm;
      s = gs;
      red_off ^= (uint32_t)kRedSetBytes;
    }
    if (csize > 1) {
      const uint32_t par = cl_phase & 1u;
      const uint32_t set = par * kClsSetBytes;
      if (gtid == 0) mbar_arrive_expect_tx(clbar, csize * 2 * 4);
      if ((uint32_t)gtid < csize) {
        const uint32_t r = (uint32_t)gtid;
        const uint32_t rbar = map_to_cta(clbar, r);
        st_async_f32(map_to_cta(cls_addr + set + (crank * kRedNV + 0) * 4, r), m, rbar);
        st_async_f32(map_to_cta(cls_addr + set + (crank * kRedNV + 1) * 4, r), s, rbar);
      }
      mbar_wait(clbar, par);
      const volatile float* slots = reinterpret_cast<const volatile float*>(smem_raw + cls_off + set);
      float cm = slots[0];
      for (uint32_t r = 1; r < csize; ++r) cm = MaxOp::apply(cm, slots[r * kRedNV]);
      float cs = 0.f;
      for (uint32_t r = 0; r < csize; ++r) cs = fmaf(slots[r * kRedNV + 1], softmax_rescale<T>(slots[r * kRedNV], cm), cs);
      m = cm;
      s = cs;
      cl_phase += 1u;
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
      // double-buffered slot sets: one named barrier per reduction is enough (a set is
      // rewritten only two reductions later, after every reader passed the next barrier)
      float* slots = reinterpret_cast<float*>(smem_raw + red_off);
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) slots[i * kMaxWarpsPerGroup + warp] = v[i];
      }
      named_bar_sync(1 + gid, TPR);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        if constexpr (WPG == 2) {
          const float2 t = *reinterpret_cast<const float2*>(slots + i * kMaxWarpsPerGroup);
          v[i] = Op::apply(t.x, t.y);
        } else {
          float4 t = *reinterpret_cast<const float4*>(slots + i * kMaxWarpsPerGroup);
          float r = Op::apply(Op::apply(t.x, t.y), Op::apply(t.z, t.w));
#pragma unroll
          for (int q = 1; q < WPG / 4; ++q) {
            t = *reinterpret_cast<const float4*>(slots + i * kMaxWarpsPerGroup + 4 * q);
            r = Op::apply(r, Op::apply(Op::apply(t.x, t.y), Op::apply(t.z, t.w)));
          }
          v[i] = r;
        }
      }
      red_off ^= (uint32_t)kRedSetBytes;  // sets are kRedSetBytes-aligned pairs
    }
    if (csize > 1) {
      const uint32_t par = cl_phase & 1u;
      const uint32_t set = par * kClsSetBytes;
      if (gtid == 0) mbar_arrive_expect_tx(clbar, csize * NV * 4);
      if ((uint32_t)gtid < csize) {
        const uint32_t r = (uint32_t)gtid;
        const uint32_t rbar = map_to_cta(clbar, r);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          st_async_f32(map_to_cta(cls_addr + set + (crank * kRedNV + i) * 4, r), v[i], rbar);
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

// ---------------------------------------------------------------------------
// The group's private ring of D stages. Tile k of group g is the CTA's local tile
// g + k*G; it lives in stage k % D — tracked as a running (stage, phase) pair.
// All per-tile bookkeeping is 32-bit counters and running pointers.
// ---------------------------------------------------------------------------
// XS: every stage also carries the residual slice and the mask bytes (fused-dropout backward) — a compile-time
// flavour, so that the producer path of every other kernel stays as short as it was
template <typename T, int TPR, bool DYN = false, bool XS = false>
struct Ring {
  const RowParams& p;
  const Group<TPR>& g;
  uint32_t smem0;
  // consumer cursor
  uint32_t stage_off;  // smem_raw-relative byte offset of the current stage
  uint32_t bar;        // shared-window address of the current stage's full barrier
  uint32_t s;          // stage index in [0, D)
  uint32_t phase;
  uint32_t ring_off_g, bar_g;
  int32_t k;           // current tile (group-local)
  int32_t ntiles_g;    // tiles this group walks under the static schedule
  int32_t cur_nrows;   // rows of the current tile (set by acquire())
  int32_t last_nrows;  // rows of the group's last static tile (tile_rows unless it is the ragged global tail)
  int64_t row0;        // first row of the current tile
  int64_t row_step;    // rows between consecutive static tiles of this group
  // Hybrid schedule (DYN kernels with p.dynamic set): every group walks ntiles_g statically assigned
  // tiles with running pointers — the cheap path — and the last 1/p.dynamic of the problem is a
  // shared pool drawn from a global ticket counter, so SMs that stream faster take more of it and
  // all CTAs finish together. A pool tile is not free: ptxas makes the warp wait for the ticket's
  // scoreboard at the next divergent branch (one row later), and under a saturating stream the
  // atomic sits behind the SM's own streaming stores for longer than that — measured +10 % per
  // pool tile on one-stream kernels, nothing on two-stream kernels (their rows take twice as long).
  // Hence a pool of 1/4 (one input stream) or 1/2 (two), chosen in b200norm.cu.
  bool dyn;
  int32_t ntiles_all;  // tiles of the whole problem
  int32_t dyn_base;    // first tile index of the shared pool
  int32_t pool_size;   // tiles in the shared pool
  // Pool slots are fetched a whole ring revolution ahead: under a saturating stream an L2 atomic
  // takes longer than one tile (fetching one refill ahead cost 20 % per pool tile). Lane s of the
  // group's first warp owns the fetch for stage s's NEXT refill (raw return value of the atomic:
  // nothing depends on it until then); the first revolution is one batch fetched in the prologue.
  int32_t pending;     // lane s: ticket (raw bits, see ticket_fetch_add) for the next refill of stage s
  int32_t first_base;  // lane 0: first ticket of the prologue batch (depth tickets)
  int32_t n_dyn;       // pool refills done so far
  bool saw_empty;      // a fetch of this group has found the pool empty
  uint32_t tidx_g;     // smem_raw-relative offset of this group's stage -> tile index table
  // producer side
  const uint8_t* base0;
  const uint8_t* base1;
  const uint8_t* base2;   // fused-dropout backward: the residual stream (third slot of a stage) ...
  const uint8_t* base_m;  // ... and the mask bytes (behind it), one byte per element

  __device__ __forceinline__ Ring(const RowParams& p_, uint32_t smem0_, const Group<TPR>& g_)
      : p(p_), g(g_), smem0(smem0_) {
    const int64_t ntiles = ceil_div64(p.rows, p.tile_rows);
    const int64_t nclusters = gridDim.x / g.csize;
    const int64_t t0 = (int64_t)(blockIdx.x / g.csize) + (int64_t)g.gid * nclusters;
    const int64_t tstep = (int64_t)p.groups * nclusters;
    const int64_t n_g = t0 < ntiles ? (ntiles - t0 + tstep - 1) / tstep : 0;
    ntiles_g = (int32_t)n_g;
    last_nrows = p.tile_rows;
    if (n_g > 0 && t0 + (n_g - 1) * tstep == ntiles - 1) {
      last_nrows = (int32_t)(p.rows - (ntiles - 1) * p.tile_rows);
    }
    dyn = false;
    if constexpr (DYN) {
      // static tiles per group: all but 1/p.dynamic of the problem (p.dynamic = 1: only the first ring fill)
      int64_t ks = p.dynamic > 0 ? (ntiles - ntiles / p.dynamic) / tstep : 0;
      if (p.dynamic == 1 && ntiles >= tstep * p.depth) ks = p.depth;
      if (p.dynamic != 0 && ks >= p.depth && ntiles < kTicketMax) {
        dyn = true;
        ntiles_g = (int32_t)ks;      // all of them full tiles: the ragged tail is in the pool
        last_nrows = p.tile_rows;
        ntiles_all = (int32_t)ntiles;
        dyn_base = (int32_t)(ks * tstep);
        pool_size = ntiles_all - dyn_base;
        pending = kTicketNone;
        first_base = kTicketNone;
        n_dyn = 0;
        saw_empty = false;
        tidx_g = p.off_tidx + 4u * (uint32_t)(g.gid * p.depth);
      }
    }
    row0 = t0 * p.tile_rows;
    row_step = tstep * p.tile_rows;
    k = 0;
    s = 0;
    phase = 0;
    cur_nrows = 0;
    ring_off_g = p.off_ring + (uint32_t)(g.gid * p.depth) * p.stage_bytes;
    bar_g = smem0 + 8u * (uint32_t)(g.gid * p.depth);
    stage_off = ring_off_g;
    bar = bar_g;
    const size_t col0_bytes = (size_t)g.crank * p.slice_bytes;
    base0 = static_cast<const uint8_t*>(p.in0) + col0_bytes;
    base1 = static_cast<const uint8_t*>(p.in1) + col0_bytes;
    if constexpr (XS) {
      base2 = static_cast<const uint8_t*>(p.in2) + col0_bytes;
      base_m = p.mask_in + (size_t)g.crank * (size_t)p.cols_cta;
    } else {
      base2 = nullptr;
      base_m = nullptr;
    }
  }

  __device__ __forceinline__ int nrows_of(int32_t kk) const { return kk == ntiles_g - 1 ? last_nrows : p.tile_rows; }
  __device__ __forceinline__ int nrows() const { return nrows_of(k); }

  // Issue the bulk copies of nr rows starting at row r0 into stage (off, b).
  // Executed by the group's first warp (all 32 lanes, converged).
  __device__ __forceinline__ void issue_rows(int nr, int64_t r0, uint32_t off, uint32_t b) const {
    const uint32_t bytes = (uint32_t)nr * p.slice_bytes;
    const uint32_t dst = smem0 + off;
    const uint64_t policy = l2_policy_evict_first();
    // fused-dropout backward (XS): dy | x | residual | mask bytes in one stage
    constexpr bool extra = XS;
    const uint32_t mbytes_row = (uint32_t)p.cols_cta;
    const uint32_t total = bytes * (uint32_t)p.nstreams + (extra ? bytes + (uint32_t)nr * mbytes_row : 0u);
    // No proxy fence is needed for this write-after-read: every reader consumed its loads
    // (the values fed the reduction) before the group barrier that precedes refill().
    if (p.contiguous) {
      if (g.lane == 0) {
        mbar_arrive_expect_tx(b, total);
        bulk_g2s(dst, base0 + (size_t)r0 * p.in0_stride * sizeof(T), bytes, b, policy);
        if (p.nstreams > 1) {
          bulk_g2s(dst + p.stream_bytes, base1 + (size_t)r0 * p.in1_stride * sizeof(T), bytes, b, policy);
        }
        if constexpr (extra) {
          bulk_g2s(dst + 2u * p.stream_bytes, base2 + (size_t)r0 * p.in2_stride * sizeof(T), bytes, b, policy);
          bulk_g2s(dst + 3u * p.stream_bytes, base_m + (size_t)r0 * (size_t)p.cols, (uint32_t)nr * mbytes_row, b, policy);
        }
      }
    } else {
      if (g.lane == 0) mbar_arrive_expect_tx(b, total);
      __syncwarp();
      for (int r = g.lane; r < nr; r += 32) {
        bulk_g2s(dst + (uint32_t)r * p.slice_bytes, base0 + (size_t)(r0 + r) * p.in0_stride * sizeof(T),
                 p.slice_bytes, b, policy);
        if (p.nstreams > 1) {
          bulk_g2s(dst + p.stream_bytes + (uint32_t)r * p.slice_bytes,
                   base1 + (size_t)(r0 + r) * p.in1_stride * sizeof(T), p.slice_bytes, b, policy);
        }
        if constexpr (extra) {
          bulk_g2s(dst + 2u * p.stream_bytes + (uint32_t)r * p.slice_bytes,
                   base2 + (size_t)(r0 + r) * p.in2_stride * sizeof(T), p.slice_bytes, b, policy);
          bulk_g2s(dst + 3u * p.stream_bytes + (uint32_t)r * mbytes_row, base_m + (size_t)(r0 + r) * (size_t)p.cols,
                   mbytes_row, b, policy);
        }
      }
    }
  }
  // static tile kk of this group (first row r0)
  __device__ __forceinline__ void issue(int32_t kk, int64_t r0, uint32_t off, uint32_t b) const {
    issue_rows(nrows_of(kk), r0, off, b);
  }

  // Fill the ring before the first tile.
  __device__ __forceinline__ void prologue() {
    if (g.warp == 0) {
      int64_t r0 = row0;
      for (int i = 0; i < p.depth && i < ntiles_g; ++i) {
        issue(i, r0, ring_off_g + (uint32_t)i * p.stage_bytes, bar_g + 8u * (uint32_t)i);
        r0 += row_step;
      }
      if constexpr (DYN) {
        // the slots of the first ring revolution past the static share, as one batch
        if (dyn && g.lane == 0) first_base = ticket_fetch_add(p.sched_words, p.depth);
      }
    }
  }

  // Pool tiles: publish the stage's tile index, then start its copies (or, for the end mark,
  // complete the stage's phase with a plain arrive). The table write is ordered before the
  // consumers' read by the release of the arrive and the acquire of their mbarrier wait.
  __device__ __forceinline__ void dyn_issue(int32_t t, uint32_t off, uint32_t b, uint32_t tidx_off) const {
    if (g.lane == 0) *reinterpret_cast<volatile int32_t*>(smem_raw + tidx_off) = t;
    if (t >= 0) {
      const int64_t r0 = (int64_t)t * p.tile_rows;
      const int64_t left = p.rows - r0;
      issue_rows(left < p.tile_rows ? (int)left : p.tile_rows, r0, off, b);
    } else if (g.lane == 0) {
      mbar_arrive(b);
    }
  }

  // Refill the current stage with this group's tile k+D. PRECONDITION: every thread of the
  // group has finished reading the stage and a group-wide barrier (Group::allreduce or
  // Group::sync) separates those reads from this call.
  __device__ __forceinline__ void refill() {
    if constexpr (DYN) {
      if (dyn && k + p.depth >= ntiles_g) {  // past the static share: a pool tile or the end mark
        if (g.warp == 0) {
          const int32_t raw = n_dyn < p.depth ? ticket_value(__shfl_sync(0xffffffffu, first_base, 0)) + n_dyn
                                              : ticket_value(__shfl_sync(0xffffffffu, pending, (int)s));
          ++n_dyn;
          int32_t t = -1;
          if (raw < pool_size) {
            // slots of one group are fetched microseconds apart and come back in order; a tile must
            // never be dropped silently if that ever fails to hold
            if (saw_empty) __trap();
            t = dyn_base + raw;
            if (g.lane == (int)s) pending = ticket_fetch_add(p.sched_words, 1);
          } else {
            saw_empty = true;
          }
          dyn_issue(t, stage_off, bar, tidx_g + 4u * s);
        }
        return;
      }
    }
    if (g.warp == 0 && k + p.depth < ntiles_g) issue(k + p.depth, row0 + (int64_t)p.depth * row_step, stage_off, bar);
  }

  // Wait for the current stage; false when this group has no more tiles. Sets row0 / cur_nrows.
  __device__ __forceinline__ bool acquire() {
    if constexpr (DYN) {
      if (dyn && k >= ntiles_g) {
        mbar_wait(bar, phase);
        const int32_t t = *reinterpret_cast<const volatile int32_t*>(smem_raw + tidx_g + 4u * s);
        if (t < 0) return false;
        row0 = (int64_t)t * p.tile_rows;
        const int64_t left = p.rows - row0;
        cur_nrows = left < p.tile_rows ? (int32_t)left : p.tile_rows;
        return true;
      }
    }
    if (k >= ntiles_g) return false;
    mbar_wait(bar, phase);
    cur_nrows = nrows_of(k);
    return true;
  }

  // Exit protocol of the dynamic schedule: once every fetch of every group of the CTA has returned
  // the CTA reports itself done, and the CTA whose report is the grid's last returns both scheduler
  // words to zero for the next launch.
  __device__ __forceinline__ void finish() const {
    if constexpr (DYN) {
      if (dyn && g.warp == 0) {
        asm volatile("" ::"r"(pending), "r"(first_base));  // wait for this lane's outstanding fetches
        __syncwarp();
        if (g.lane == 0) {
          uint32_t* cnt = reinterpret_cast<uint32_t*>(smem_raw + p.off_tidx + 4u * kSchedTableInts);
          if (atomicAdd(cnt, 1u) == (uint32_t)p.groups - 1u) {
            if (atomicAdd(p.sched_words + 1, 1u) == gridDim.x - 1) {
              p.sched_words[0] = 0u;
              p.sched_words[1] = 0u;
            }
          }
        }
      }
    }
  }

  __device__ __forceinline__ bool more() const { return k < ntiles_g; }
  __device__ __forceinline__ void wait() const { mbar_wait(bar, phase); }
  __device__ __forceinline__ void advance() {
    ++k;
    row0 += row_step;
    ++s;
    stage_off += p.stage_bytes;
    bar += 8u;
    if (s == (uint32_t)p.depth) {
      s = 0;
      phase ^= 1u;
      stage_off = ring_off_g;
      bar = bar_g;
    }
  }
};

// Barrier setup shared by all ring kernels.
__device__ __forceinline__ void ring_setup(const RowParams& p, uint32_t smem0) {
  const int nbars = p.groups * p.depth + p.groups;
  if (threadIdx.x == 0) {
    for (int s = 0; s < nbars; ++s) mbar_init(smem0 + 8u * (uint32_t)s, 1);
    fence_mbar_init();
    *reinterpret_cast<uint32_t*>(smem_raw + p.off_tidx + 4u * kSchedTableInts) = 0u;  // groups that saw the empty pool
  }
  __syncthreads();
  if (cluster_nctarank() > 1) {
    cluster_sync_all();  // peers' mbarriers are initialised before any st.async targets them
  }
  // Everything above touched only this CTA's shared memory: under a programmatic dependent
  // launch it overlaps the previous kernel's tail. Let our own successor do the same, then
  // wait for the predecessor's memory before the first global access.
  pdl_launch_dependents();
  pdl_wait();
}

__device__ __forceinline__ void ring_teardown(const RowParams& p) {
  if (cluster_nctarank() > 1) {
    __syncwarp();
    cluster_sync_all();  // nobody exits while a peer may still st.async into its smem
  }
}

// Chunk ownership: thread gtid owns chunks c = gtid + j*TPR, j < CH (CH = ceil(nchunks/TPR),
// so only the last j needs a bound check). CH == 0 selects the looping ("re-read") form
// for slices wider than 4*TPR chunks.
template <int TPR, int CH, typename F>
__device__ __forceinline__ void for_each_chunk(int gtid, int nchunks, F&& f) {
  if constexpr (CH > 0) {
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      if (j < CH - 1 || gtid + j * TPR < nchunks) f(j, j * TPR);
    }
  } else {
#pragma unroll 4
    for (int c = 0; gtid + c < nchunks; c += TPR) f(0, c);
  }
}

__device__ __forceinline__ uint4 ld_smem_u4(const uint8_t* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void st_global_stream(void* p, const uint4& v) {
  __stcs(reinterpret_cast<uint4*>(p), v);
}

// LayerNorm forward: what is stored as the second statistic (invstd, or the variance / M2 of a stand-alone
// var_mean / welford segment).
__device__ __forceinline__ float ln_second_stat(const RowParams& p, float m2, float invstd) {
  if (p.flags & kFlagStatVar) return m2 * p.var_scale;
  if (p.flags & kFlagStatM2) return m2;
  return invstd;
}

// Packed per-column parameters (weight / bias) of the chunks a thread owns, kept in
// registers for the CTA's lifetime. Absent parameters are replaced by the identity
// (1.0 for a scale, -0.0 for a shift) so the hot loop has no branches.
template <typename T>
__device__ __forceinline__ uint4 splat_chunk(float v) {
  if constexpr (sizeof(T) == 4) {
    const uint32_t u = __float_as_uint(v);
    return make_uint4(u, u, u, u);
  } else {
    const uint32_t u = IO<T>::pack2(v, v);
    return make_uint4(u, u, u, u);
  }
}
template <typename T, int TPR, int CH>
__device__ __forceinline__ void load_params_to_regs(const void* gp, size_t col0_bytes, int gtid, int nchunks,
                                                    float identity, uint4 (&q)[CH > 0 ? CH : 1]) {
  if constexpr (CH > 0) {
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      q[j] = splat_chunk<T>(identity);
      if (gp != nullptr && gtid + j * TPR < nchunks) {
        q[j] = ldg_128_nc(static_cast<const uint8_t*>(gp) + col0_bytes + 16u * (uint32_t)(gtid + j * TPR));
      }
    }
  }
}
// Parameter chunk for the looping form: read through the read-only path each time.
template <typename T>
__device__ __forceinline__ uint4 load_param_chunk(const void* gp, size_t col0_bytes, int c, float identity) {
  if (gp == nullptr) return splat_chunk<T>(identity);
  return ldg_128_nc(static_cast<const uint8_t*>(gp) + col0_bytes + 16u * (uint32_t)c);
}

// fp32 copies of per-column parameters in shared memory, split in VEC/4 planes of
// [chunk][4 floats] so that lane l reading chunk (c0 + l) of a plane is conflict-free.
// Used where two parameters (LayerNorm's weight and bias) would not fit in registers unpacked.
template <typename T>
__device__ __forceinline__ void stage_param_smem(const void* gp, size_t col0_bytes, uint32_t off, int nchunks,
                                                 float identity) {
  constexpr int VEC = IO<T>::VEC;
  for (int c = threadIdx.x; c < nchunks; c += blockDim.x) {
    float2 f[VEC / 2];
    IO<T>::unpack2(load_param_chunk<T>(gp, col0_bytes, c, identity), f);
#pragma unroll
    for (int h = 0; h < VEC / 4; ++h) {
      *reinterpret_cast<float4*>(smem_raw + off + (uint32_t)h * (uint32_t)nchunks * 16u + 16u * (uint32_t)c) =
          make_float4(f[2 * h].x, f[2 * h].y, f[2 * h + 1].x, f[2 * h + 1].y);
    }
  }
}
template <int VEC>
__device__ __forceinline__ void load_param_smem(uint32_t off, int nchunks, int c, float2 (&w)[VEC / 2]) {
#pragma unroll
  for (int h = 0; h < VEC / 4; ++h) {
    const float4 v =
        *reinterpret_cast<const float4*>(smem_raw + off + (uint32_t)h * (uint32_t)nchunks * 16u + 16u * (uint32_t)c);
    w[2 * h] = make_float2(v.x, v.y);
    w[2 * h + 1] = make_float2(v.z, v.w);
  }
}

// Fused residual add of the forward kernels (the prologue nvFuser's persistent schedulers
// absorb for `norm(x + residual)`, benchmarks/python/test_rmsnorm_add_fwd.py,
// test_dropout_layernorm_fwd.py minus the dropout): t = x + r in fp32; when the fusion rounds the
// sum to the io dtype before normalising (a bf16 `hidden + residual` feeding the norm) the same
// rounding is applied here; the sum is streamed out once if requested.
template <typename T, int H>
__device__ __forceinline__ void add_residual_chunk(float2 (&t)[H], const uint8_t* rchunk, bool round_sum,
                                                   uint8_t* sum_out /* nullptr = no store */) {
  float2 r[H];
  IO<T>::unpack2(ld_smem_u4(rchunk), r);
#pragma unroll
  for (int k = 0; k < H; ++k) t[k] = f2add(t[k], r[k]);
  if constexpr (sizeof(T) == 4) {
    if (sum_out != nullptr) st_global_stream(sum_out, IO<T>::pack2(t));
  } else {
    if (round_sum || sum_out != nullptr) {
      const uint4 q = IO<T>::pack2(t);
      if (sum_out != nullptr) st_global_stream(sum_out, q);
      if (round_sum) IO<T>::unpack2(q, t);
    }
  }
}

// instance norm: the row's channel parameter (weight: identity 1, bias: identity 0), read through the read-only path
template <typename T>
__device__ __forceinline__ float row_param(const void* gp, int64_t row, int32_t channels, float identity) {
  if (gp == nullptr) return identity;
  return IO<T>::to_f32(__ldg(static_cast<const T*>(gp) + (row % channels)));
}

// Fused dropout of one 16-byte chunk (VEC = 2 * H consecutive elements starting at linear index `lin` of the [M,N]
// tensor, lin a multiple of 4): VEC / 4 Philox calls give the VEC words in element order (word i mod 4 of
// subsequence i / 4: device_lower/pass/rng.cpp:95-127), kept elements are scaled by 1 / keep, dropped ones become 0
// (x * mask * scale, benchmarks/python/test_dropout_layernorm_fwd.py:47-50), and the mask bytes are streamed out.
template <typename T, int H>
__device__ __forceinline__ void dropout_chunk(float2 (&t)[H], const RowParams& p, unsigned long long seed,
                                              unsigned long long offset, unsigned long long lin) {
  constexpr int VEC = 2 * H;
  uint32_t bytes[VEC / 4];
#pragma unroll
  for (int q = 0; q < VEC / 4; ++q) {
    const uint4 r = philox4x32_10(seed, (lin >> 2) + (unsigned long long)q, offset);
    const bool k0 = dropout_keep_word(r.x, p.drop_t1, p.drop_t2), k1 = dropout_keep_word(r.y, p.drop_t1, p.drop_t2);
    const bool k2 = dropout_keep_word(r.z, p.drop_t1, p.drop_t2), k3 = dropout_keep_word(r.w, p.drop_t1, p.drop_t2);
    t[2 * q].x = k0 ? t[2 * q].x * p.drop_scale : 0.f;
    t[2 * q].y = k1 ? t[2 * q].y * p.drop_scale : 0.f;
    t[2 * q + 1].x = k2 ? t[2 * q + 1].x * p.drop_scale : 0.f;
    t[2 * q + 1].y = k3 ? t[2 * q + 1].y * p.drop_scale : 0.f;
    bytes[q] = (k0 ? 1u : 0u) | (k1 ? 0x100u : 0u) | (k2 ? 0x10000u : 0u) | (k3 ? 0x1000000u : 0u);
  }
  uint8_t* dst = p.mask_out + lin;
  if constexpr (VEC == 8) {
    asm volatile("st.global.cs.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(bytes[0]), "r"(bytes[1]) : "memory");
  } else {
    asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(dst), "r"(bytes[0]) : "memory");
  }
}

template <typename T>
__device__ __forceinline__ uint8_t* out1_row_ptr(const RowParams& p, int64_t row, uint32_t crank, int gtid) {
  return static_cast<uint8_t*>(p.out1) + ((size_t)row * p.out1_stride + (size_t)crank * p.cols_cta) * sizeof(T) +
         16u * (uint32_t)gtid;
}

template <typename T>
__device__ __forceinline__ uint8_t* out_row_ptr(const RowParams& p, int64_t row, uint32_t crank, int gtid) {
  return static_cast<uint8_t*>(p.out0) + ((size_t)row * p.out_stride + (size_t)crank * p.cols_cta) * sizeof(T) +
         16u * (uint32_t)gtid;
}

}  // namespace b200norm

// kernels_bwd.cuh — LayerNorm / RMSNorm backward on the TMA ring, with the
// dgamma / dbeta column sums fused in (the InnerOuterPersistent segment).
//
// Reference math (csrc/ops/normalization.cpp):
//   layer_norm_backward :382-436, rms_norm_backward :438-485 (keeps the -sum(g)
//   term), thunder_rms_norm_backward :487-535 (textbook form, rms statistic).
//     x_hat = (x - mean) * invstd            g = dy * w
//     dx    = (invstd / N) * (N*g - sum(g) - x_hat * sum(g*x_hat))
//     dw    = sum_rows dy * x_hat            db = sum_rows dy
// Evaluated here as  dx = A*g + B*x_hat + C  with per-row scalars
//     A = invstd,  B = -invstd * sum(g*x_hat) / N,  C = -invstd * sum(g) / N
// (same polynomial, two packed FMAs and one multiply per two elements).
//
// The reference's generated kernel (scheduler/normalization_inner_outer*.cpp)
// keeps per-CTA partial column sums in registers and finishes them with a grid
// reduction; we keep per-thread fp32 column accumulators for the CTA's lifetime,
// combine the consumer groups through shared memory once, write one fp32 partial
// row per CTA, and finish in finalize_partials_kernel (kernels_generic.cuh).
#pragma once

#include "ring.cuh"

namespace b200norm {

// Sense-reversing grid barrier on ONE workspace header word: bit 31 is the sense, the low bits
// count arrivals. The word is zero before the first launch; the last arriver flips the sense and
// clears the count with a single atomic add, so the barrier opens one atomic after the last
// arrival (no separate reset store + fence on the critical path) and the count is zero again
// afterwards, whatever grid size the next launch uses. PRECONDITION: every CTA of the grid is
// resident (the grid is at most one CTA per SM; see launch_typed) and a __syncthreads() after the
// CTA's last global write + __threadfence().
__device__ __forceinline__ void grid_barrier(uint32_t* word) {
  if (threadIdx.x == 0) {
    const uint32_t old = atom_add_acq_rel_gpu(word, 1u);
    const uint32_t sense = old & 0x80000000u;
    if ((old & 0x7fffffffu) == gridDim.x - 1) {
      atom_add_acq_rel_gpu(word, 0x80000000u - gridDim.x);
    } else {
      Watchdog wd;
      while ((ld_acquire_gpu(word) & 0x80000000u) == sense) wd.tick();
    }
  }
  __syncthreads();
}

// After the grid barrier: out[k][n] = sum over the nparts partial rows, in row order.
// The [K][N] result is cut into chunks of CW float4 columns; CTA b finishes chunks b,
// b+grid, ... A thread owns one float4 column of the chunk and every RS-th partial row
// (RS = threads / CW); all its loads are issued before the first add (one L2 round trip),
// row-splits are combined by shuffles inside a warp and through shared memory across warps.
// PUSH (several GPUs, fused flavour): instead of writing a local result, every finished sum goes to EVERY rank's
// receive area as a (value, epoch) word — see peer_collect_columns.
template <typename T, int CW, bool PUSH>
__device__ __forceinline__ void finish_columns(const RowParams& p, int K, int F, int nct, float* sums_out) {
  constexpr int kMaxLoads = 10;
  const int tx = threadIdx.x % CW;
  const int ty = threadIdx.x / CW;
  const int RS = nct / CW;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarps = nct >> 5;
  const int nparts = (int)(gridDim.x / cluster_nctarank());
  const int fpr = p.cols / 4;  // float4 columns per result row
  float4* sm = reinterpret_cast<float4*>(smem_raw + p.off_acc);  // [nwarps][CW]
  const int nchunk = (F + CW - 1) / CW;
  for (int chunk = blockIdx.x; chunk < nchunk; chunk += gridDim.x) {
    const int f = chunk * CW + tx;
    float4 v[kMaxLoads];
    const bool col_ok = f < F;
    const int kk = col_ok ? f / fpr : 0;
    const int c4 = col_ok ? f - kk * fpr : 0;
    const float4* src = reinterpret_cast<const float4*>(p.partials + (size_t)kk * p.cols) + c4;
    const size_t pstride4 = (size_t)K * fpr;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int base = 0; base < nparts; base += RS * kMaxLoads) {
#pragma unroll
      for (int i = 0; i < kMaxLoads; ++i) {
        const int pi = base + ty + i * RS;
        v[i] = (col_ok && pi < nparts) ? __ldcg(src + (size_t)pi * pstride4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < kMaxLoads; ++i) {
        a.x += v[i].x;
        a.y += v[i].y;
        a.z += v[i].z;
        a.w += v[i].w;
      }
    }
    // lanes tx, tx+CW, tx+2CW, ... of a warp hold different row-splits of the same column
#pragma unroll
    for (int off = CW; off < 32; off <<= 1) {
      a.x += __shfl_xor_sync(0xffffffffu, a.x, off);
      a.y += __shfl_xor_sync(0xffffffffu, a.y, off);
      a.z += __shfl_xor_sync(0xffffffffu, a.z, off);
      a.w += __shfl_xor_sync(0xffffffffu, a.w, off);
    }
    __syncthreads();  // the previous chunk's readers are done with sm
    if (lane < CW) sm[warp * CW + lane] = a;
    __syncthreads();
    if (threadIdx.x < CW * 4) {
      const float* smf = reinterpret_cast<const float*>(sm);
      const int e = chunk * CW * 4 + threadIdx.x;  // element index in the flattened [K][N] result
      if (e < K * p.cols) {
        float r = 0.f;
        for (int w = 0; w < nwarps; ++w) r += smf[w * CW * 4 + threadIdx.x];
        if constexpr (PUSH) {
          const size_t off = (size_t)p.peer_ll_off + (size_t)(p.peer_epoch & 1u) * p.peer_ll_parity_bytes +
                             ((size_t)p.peer_rank * p.peer_slot_floats + (size_t)e) * 8;
#pragma unroll
          for (int q = 0; q < kMaxPeers; ++q) {
            if (q < p.peer_world) st_ll(reinterpret_cast<uint64_t*>(p.peer_xchg[q] + off), r, p.peer_epoch);
          }
        } else if (sums_out != nullptr) {
          sums_out[e] = r;
        } else {
          const int k2 = e / p.cols;
          T* o = static_cast<T*>(k2 == 0 ? p.dw_out : p.db_out);
          if (o != nullptr) o[e - k2 * p.cols] = IO<T>::from_f32(r);
        }
      }
    }
  }
}

// The all-reduce of the column sums, inside the backward kernel (row-sharded backward on the GPUs
// of one box; reference mechanism: lowerToAllreduce + postAllreduce, host_ir/lower_to_communication.cpp
// :266-284, multidevice/post_communication.cpp:288-313 — a separate NCCL kernel there; its NVLink-native
// one-shot reduce: runtime/multicast_reduce.cu:13-40, csrc/multidevice/cuda_p2p.cpp:828-918).
// One-shot, push based, no flags and no second grid barrier:
//   1. finish_columns<PUSH>: the CTA that finishes a chunk of this rank's [K][N] sums stores every sum into the
//      receive area of EVERY rank (its own included) as a 64-bit (value, launch epoch) word — remote stores over
//      NVLink are posted, nobody waits for a round trip;
//   2. the same CTA then polls, in its OWN memory, the words of its chunks from all P source ranks (all P loads
//      of an element are issued before the first is examined), adds them in rank order — the same order on
//      every rank, so all ranks get the same bits — and casts once, as on a single GPU.
// The epoch makes a word valid: what the slot held before (two launches ago, the parities alternate) carries an
// older epoch. A rank can be at most one launch ahead of a peer that is still reading: it only gets to launch
// e+2 after collecting everybody's sums of e+1, which a peer sends after it has finished reading e.
template <typename T, int CW>
__device__ __forceinline__ void peer_collect_columns(const RowParams& p, int K, int F) {
  const int P = p.peer_world;
  const uint32_t epoch = p.peer_epoch;
  const int total = K * p.cols;
  const uint8_t* base = p.peer_xchg[p.peer_rank] + p.peer_ll_off + (size_t)(epoch & 1u) * p.peer_ll_parity_bytes;
  const size_t rank_stride = (size_t)p.peer_slot_floats * 8;
  const int nchunk = (F + CW - 1) / CW;
  if ((int)threadIdx.x < CW * 4) {
    for (int chunk = blockIdx.x; chunk < nchunk; chunk += gridDim.x) {
      const int e = chunk * CW * 4 + threadIdx.x;
      if (e >= total) continue;
      const uint64_t* src = reinterpret_cast<const uint64_t*>(base) + e;
      uint64_t w[kMaxPeers];
      Watchdog wd(p.peer_timeout_ns);
      for (;;) {
        bool ready = true;
#pragma unroll
        for (int q = 0; q < kMaxPeers; ++q) {
          if (q < P) w[q] = ld_ll(reinterpret_cast<const uint64_t*>(reinterpret_cast<const uint8_t*>(src) + q * rank_stride));
        }
#pragma unroll
        for (int q = 0; q < kMaxPeers; ++q) {
          if (q < P) ready = ready && (uint32_t)(w[q] >> 32) == epoch;
        }
        if (ready) break;
        wd.tick();
      }
      float a = __uint_as_float((uint32_t)w[0]);
#pragma unroll
      for (int q = 1; q < kMaxPeers; ++q) {
        if (q < P) a += __uint_as_float((uint32_t)w[q]);
      }
      if (p.sums_out != nullptr) {
        p.sums_out[e] = a;
      } else {
        const int k2 = e / p.cols;
        T* o = static_cast<T*>(k2 == 0 ? p.dw_out : p.db_out);
        if (o != nullptr) o[e - k2 * p.cols] = IO<T>::from_f32(a);
      }
    }
  }
}

// Deferred flavour (B200NORM_F_PEER_PUBLISH): nobody waits for a peer inside the backward.
// Before the slot is overwritten every peer must have consumed what it held two publishes ago
// (normally true long since); after the slot is written the last CTA to get there tells every peer.
__device__ __forceinline__ void peer_wait_consumed(const RowParams& p) {
  if (p.peer_epoch > 2u) {
    if ((int)threadIdx.x < p.peer_world) {
      const uint32_t* flag =
          reinterpret_cast<const uint32_t*>(p.peer_xchg[p.peer_rank] + kPeerConsumedOffset) + threadIdx.x;
      Watchdog wd(p.peer_timeout_ns);
      while ((int32_t)(ld_acquire_sys(flag) - (p.peer_epoch - 2u)) < 0) wd.tick();
    }
    __syncthreads();
  }
}
__device__ __forceinline__ void peer_publish(const RowParams& p) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t old = atom_add_acq_rel_gpu(p.sync_words + 1, 1u);
    if (old == gridDim.x - 1) {
      p.sync_words[1] = 0u;  // next used by the next launch
      for (int q = 0; q < p.peer_world; ++q) {
        st_release_sys(reinterpret_cast<uint32_t*>(p.peer_xchg[q]) + p.peer_rank, p.peer_epoch);
      }
    }
  }
}

// After the row loop: combine the groups' column accumulators, emit this CTA's partial row and —
// with the in-kernel finish — reduce the partial rows (and, on several GPUs, the ranks' sums).
template <typename T, int TPR, int CH, bool LN>
__device__ __forceinline__ void bwd_epilogue(const RowParams& p, const Group<TPR>& g,
                                             float2 (&dw_acc)[CH][IO<T>::VEC / 2],
                                             float2 (&db_acc)[LN ? CH : 1][IO<T>::VEC / 2], bool want_dw,
                                             bool want_db) {
  constexpr int VEC = IO<T>::VEC;
  constexpr int H = VEC / 2;
  tl_mark(p, 3);
  const int nchunks = p.nchunks;
  // ---- combine the groups' column accumulators and emit this CTA's partial row ----
  if (want_dw || want_db) {
    const int nct = p.groups * TPR;
    // partials layout: [part][K][N], K = 2 for LN (dw, db), 1 for RMSNorm
    constexpr int K = LN ? 2 : 1;
    const int64_t part = blockIdx.x / g.csize;
    float* dst = p.partials + (size_t)part * K * p.cols + (size_t)g.crank * p.cols_cta;
    if (p.groups == 1) {
      // one group: the thread's accumulators ARE the CTA's column sums — straight from registers to the partial
      // row (32 contiguous bytes per thread and chunk, a warp writes 1 KB), no shared-memory staging, no barrier
      for_each_chunk<TPR, CH>(g.gtid, nchunks, [&](int j, int coff) {
        float* o = dst + (size_t)(g.gtid + coff) * VEC;
#pragma unroll
        for (int k = 0; k < H; k += 2) {
          *reinterpret_cast<float4*>(o + 2 * k) =
              make_float4(dw_acc[j][k].x, dw_acc[j][k].y, dw_acc[j][k + 1].x, dw_acc[j][k + 1].y);
          if constexpr (LN) {
            *reinterpret_cast<float4*>(o + p.cols + 2 * k) =
                make_float4(db_acc[j][k].x, db_acc[j][k].y, db_acc[j][k + 1].x, db_acc[j][k + 1].y);
          }
        }
      });
    } else {
      float* acc_w = reinterpret_cast<float*>(smem_raw + p.off_acc);
      float* acc_b = acc_w + p.cols_cta;
      named_bar_sync(15, nct);  // every group is done with the ring
      for (int gi = 0; gi < p.groups; ++gi) {
        if (g.gid == gi) {
          for_each_chunk<TPR, CH>(g.gtid, nchunks, [&](int j, int coff) {
            const int c = g.gtid + coff;
#pragma unroll
            for (int k = 0; k < H; ++k) {
              float2* pw = reinterpret_cast<float2*>(acc_w + c * VEC + 2 * k);
              float2* pb = reinterpret_cast<float2*>(acc_b + c * VEC + 2 * k);
              if (gi == 0) {
                *pw = dw_acc[j][k];
                if constexpr (LN) *pb = db_acc[j][k];
              } else {
                *pw = f2add(*pw, dw_acc[j][k]);
                if constexpr (LN) *pb = f2add(*pb, db_acc[j][k]);
              }
            }
          });
        }
        named_bar_sync(15, nct);
      }
      for (int col = threadIdx.x * 4; col < p.cols_cta; col += nct * 4) {
        *reinterpret_cast<float4*>(dst + col) = *reinterpret_cast<const float4*>(acc_w + col);
        if constexpr (LN) *reinterpret_cast<float4*>(dst + p.cols + col) = *reinterpret_cast<const float4*>(acc_b + col);
      }
    }
    if (p.fused_finish) {
      // The launch is cooperative (every CTA is resident): one grid barrier, then every CTA
      // finishes a share of the columns in a fixed order — deterministic, like the reference's
      // grid reduction (runtime/grid_reduction.cu:201-309 + grid_sync.cu:34-95).
      __threadfence();
      __syncthreads();
      tl_mark(p, 4);
      grid_barrier(p.sync_words);
      tl_mark(p, 5);
      const int F = K * (p.cols / 4);  // float4 columns of the [K][N] result
      // several GPUs: this rank's sums go to the peers (fused flavour: pushed to every rank's receive area;
      // deferred flavour: written to this rank's publish slot) instead of being cast
      const bool push = p.peer_world > 0 && !p.peer_deferred;
      float* sums_dst = p.sums_out;
      if (p.peer_world > 0 && p.peer_deferred) {
        sums_dst = reinterpret_cast<float*>(p.peer_xchg[p.peer_rank] + kPeerHeaderBytes) +
                   (size_t)(p.peer_epoch & 1u) * p.peer_slot_floats;
        peer_wait_consumed(p);  // the slot's previous content (two publishes ago) was read
      }
      const bool narrow = (F + 7) / 8 <= 2 * (int)gridDim.x;
      if (push) {
        if (narrow) {
          finish_columns<T, 8, true>(p, K, F, nct, nullptr);
        } else {
          finish_columns<T, 32, true>(p, K, F, nct, nullptr);
        }
        tl_mark(p, 8);
        if (narrow) {
          peer_collect_columns<T, 8>(p, K, F);
        } else {
          peer_collect_columns<T, 32>(p, K, F);
        }
        tl_mark(p, 9);
      } else {
        if (narrow) {
          finish_columns<T, 8, false>(p, K, F, nct, sums_dst);
        } else {
          finish_columns<T, 32, false>(p, K, F, nct, sums_dst);
        }
        tl_mark(p, 8);
        if (p.peer_world > 0) peer_publish(p);
      }
    }
  }
}

// LN: true = LayerNorm (mean input, dbeta), false = RMSNorm.
// Each thread owns chunks c = gtid + j*TPR, j < CH, of the CTA's column slice
// (the planner guarantees nchunks <= TPR*CH).
// CACHE: g and xc stay in registers between the reduction pass and the dx pass;
// otherwise the stage is re-read from shared memory.
// CEN: the dx formula keeps the centring term -sum(g) (LayerNorm always; RMSNorm as the C++ rms_norm_backward
// builder writes it). CEN = false is the textbook RMSNorm backward: no sum(g) accumulation and ONE reduced value
// per row instead of two (about 8 % fewer instructions in the cached form the C2 config runs).
// DROP (cached form only): the fused-dropout backward of benchmarks/python/test_dropout_{layernorm,rmsnorm}_bwd.py.
// Every stage carries dy | x | residual | mask bytes; x_eff = residual + (mask ? x / keep : 0) replaces x, and the
// row's gradient goes out twice: d residual = d x_eff (out1) and d x = mask ? d x_eff / keep : 0 (out0).
// ROWAFF: instance norm (kFlagRowAffine) — its own instantiations, LayerNorm only.
template <typename T, int TPR, int CH, bool LN, bool CACHE, bool CEN = true, bool DROP = false, bool ROWAFF = false>
__global__ void __launch_bounds__(kMaxBlock, 1) norm_bwd_ring_kernel(const RowParams p) {
  static_assert(!ROWAFF || (LN && !DROP), "instance norm is a LayerNorm flavour");
  static_assert(!DROP || CACHE, "the fused-dropout backward keeps g / x_hat (and the mask bits) in registers");
  constexpr int VEC = IO<T>::VEC;
  constexpr int H = VEC / 2;
  static_assert(CH > 0, "column accumulators need a fixed chunk ownership");
  static_assert(CEN || !LN, "LayerNorm backward is always centred");
  const uint32_t smem0 = smem_u32(smem_raw);
  tl_mark(p, 0);
  ring_setup(p, smem0);
  tl_mark(p, 1);
  Group<TPR> g(p, smem0);
  Ring<T, TPR, false, DROP> ring(p, smem0, g);
  ring.prologue();

  const int nchunks = p.nchunks;
  const bool want_dx = (p.flags & kFlagOutDx) != 0;
  const bool want_dw = (p.flags & kFlagOutDw) != 0;
  const bool want_db = LN && (p.flags & kFlagOutDb) != 0;
  // (the CEN instantiation still serves the textbook form when the specialised one is switched off: runtime flag)
  const bool centered = CEN && (LN || (p.flags & kFlagRmsBwdCentered) != 0);
  const bool stat_rms = !LN && (p.flags & kFlagStatIsRms) != 0;
  const size_t col0_bytes = (size_t)g.crank * p.slice_bytes;
  // instance norm (kFlagRowAffine, LayerNorm only): the weight belongs to the row's channel — g = dy inside, the row's
  // weight scales the finished dx, and the row sums of dy / dy * x_hat go out instead of column sums
  constexpr bool row_aff = ROWAFF;
  uint4 wq[CH];
  load_params_to_regs<T, TPR, CH>(row_aff ? nullptr : p.weight, col0_bytes, g.gtid, nchunks, 1.0f, wq);

  float2 dw_acc[CH][H];
  float2 db_acc[LN ? CH : 1][H];
#pragma unroll
  for (int j = 0; j < CH; ++j) {
#pragma unroll
    for (int k = 0; k < H; ++k) {
      dw_acc[j][k] = f2(0.f);
      if constexpr (LN) db_acc[j][k] = f2(0.f);
    }
  }

  // per-row statistics are prefetched one row ahead (they sit on the critical path otherwise)
  float nxt_mean = 0.f, nxt_stat = 0.f;
  if (ring.more()) {
    if constexpr (LN) nxt_mean = __ldg(p.stat0 + ring.row0);
    nxt_stat = __ldg(p.stat1 + ring.row0);
  }

  uint8_t* out_tile = out_row_ptr<T>(p, ring.row0, g.crank, g.gtid);
  const size_t out_row_bytes = (size_t)p.out_stride * sizeof(T);
  const size_t out_tile_bytes = (size_t)ring.row_step * out_row_bytes;
  uint8_t* out1_tile = DROP && p.out1 != nullptr ? out1_row_ptr<T>(p, ring.row0, g.crank, g.gtid) : nullptr;
  const size_t out1_row_bytes = (size_t)p.out1_stride * sizeof(T);
  const size_t out1_tile_bytes = (size_t)ring.row_step * out1_row_bytes;
  for (; ring.more(); ring.advance(), out_tile += out_tile_bytes, out1_tile += out1_tile_bytes) {
    ring.wait();
    if (ring.k == 0) tl_mark(p, 2);
    const int nrows = ring.nrows();
    const uint8_t* dyrow = smem_raw + ring.stage_off + 16u * (uint32_t)g.gtid;
    uint8_t* dxrow = out_tile;
    uint8_t* drrow = out1_tile;
    for (int r = 0; r < nrows; ++r, dyrow += p.slice_bytes, dxrow += out_row_bytes, drrow += out1_row_bytes) {
      const uint8_t* xrow = dyrow + p.stream_bytes;
      const float mean = nxt_mean;
      const float st = nxt_stat;
      {
        int64_t nrow = ring.row0 + r + 1;
        if (r == nrows - 1) nrow = ring.row0 + ring.row_step;
        if (nrow < p.rows) {
          if constexpr (LN) nxt_mean = __ldg(p.stat0 + nrow);
          nxt_stat = __ldg(p.stat1 + nrow);
        }
      }
      const float invstd = stat_rms ? rcp_fast(st) : st;
      const float w_row = row_aff ? row_param<T>(p.weight, ring.row0 + r, p.channels, 1.f) : 1.f;
      const float2 nmean = f2(-mean);
      const float2 is2 = f2(invstd);
      // x_hat = (x - mean) * invstd. With 16-bit inputs it is one FMA, x*invstd + (-mean*invstd): the
      // pre-rounded product costs at most |mean|/std fp32 ulps of x_hat, far below the resolution of
      // the 16-bit data it was computed from; fp32 inputs keep the subtract-then-scale form.
      constexpr bool FMA_XHAT = LN && sizeof(T) == 2;
      const float nmi_s = -mean * invstd;
      const float2 nmi = f2(nmi_s);
      auto x_hat = [&](float2 xv_) -> float2 {
        if constexpr (FMA_XHAT) {
          return f2fma(xv_, is2, nmi);
        } else if constexpr (LN) {
          return f2mul(f2add(xv_, nmean), is2);
        } else {
          return f2mul(xv_, is2);
        }
      };
      float2 gv[CACHE ? CH : 1][H];
      float2 xv[CACHE ? CH : 1][H];
      uint32_t mbits[DROP ? CH : 1][VEC / 4];  // the chunk's mask bytes (VEC of them), kept for the output pass
      const uint8_t* mrow = smem_raw + ring.stage_off + 3u * p.stream_bytes + (uint32_t)r * (uint32_t)p.cols_cta +
                            (uint32_t)(g.gtid * VEC);
      float2 s1 = f2(0.f), s2 = f2(0.f);
      for_each_chunk<TPR, CH>(g.gtid, nchunks, [&](int j, int coff) {
        float2 dy[H], x[H], w[H];
        IO<T>::unpack2(ld_smem_u4(dyrow + 16 * coff), dy);
        IO<T>::unpack2(ld_smem_u4(xrow + 16 * coff), x);
        IO<T>::unpack2(wq[j], w);
        if constexpr (DROP) {
          float2 rs[H];
          IO<T>::unpack2(ld_smem_u4(xrow + p.stream_bytes + 16 * coff), rs);
#pragma unroll
          for (int q = 0; q < VEC / 4; ++q) {
            const uint32_t mb = *reinterpret_cast<const uint32_t*>(mrow + VEC * coff + 4 * q);
            mbits[j][q] = mb;
            x[2 * q].x = rs[2 * q].x + ((mb & 0xffu) ? x[2 * q].x * p.drop_scale : 0.f);
            x[2 * q].y = rs[2 * q].y + ((mb & 0xff00u) ? x[2 * q].y * p.drop_scale : 0.f);
            x[2 * q + 1].x = rs[2 * q + 1].x + ((mb & 0xff0000u) ? x[2 * q + 1].x * p.drop_scale : 0.f);
            x[2 * q + 1].y = rs[2 * q + 1].y + ((mb & 0xff000000u) ? x[2 * q + 1].y * p.drop_scale : 0.f);
          }
        }
#pragma unroll
        for (int k = 0; k < H; ++k) {
          const float2 gg = f2mul(dy[k], w[k]);
          const float2 xh = x_hat(x[k]);
          if constexpr (CEN) s1 = f2add(s1, gg);
          s2 = f2fma(gg, xh, s2);
          dw_acc[j][k] = f2fma(dy[k], xh, dw_acc[j][k]);
          if constexpr (LN) db_acc[j][k] = f2add(db_acc[j][k], dy[k]);
          if constexpr (CACHE) {
            gv[j][k] = gg;
            xv[j][k] = xh;
          }
        }
      });
      if (want_dx || row_aff) {
        float s[2] = {s1.x + s1.y, s2.x + s2.y};  // sum(g), sum(g * x_hat)
        if constexpr (CEN) {
          g.template allreduce<SumOp, 2>(s);
        } else {
          float t1[1] = {s[1]};
          g.template allreduce<SumOp, 1>(t1);
          s[1] = t1[0];
        }
        if (CACHE && want_dx && r == nrows - 1) ring.refill();  // g / x_hat live in registers: the stage is free
        if (row_aff && g.gtid == 0 && g.crank == 0) {
          if (p.rowsum_dy != nullptr) p.rowsum_dy[ring.row0 + r] = s[0];
          if (p.rowsum_dyxh != nullptr) p.rowsum_dyxh[ring.row0 + r] = s[1];
        }
        if (!want_dx) continue;
        // dx = invstd * g - (invstd/N) * sum(g x_hat) * x_hat - (invstd/N) * sum(g)  =  A g + B x_hat + C
        // (instance norm: every term times the row's weight)
        const float a_n = invstd * p.inv_n * w_row;
        const float b_s = -(a_n * s[1]) * p.proj_scale;
        const float c_s = centered ? -(a_n * s[0]) : 0.f;
        const float2 A = f2(invstd * w_row);
        const float2 B = f2(b_s);
        const float2 C = f2(c_s);
        // re-read form: B x_hat + C folded onto x itself where x_hat is affine in x with exact-enough
        // coefficients (16-bit LayerNorm, RMSNorm): (B invstd) x + (B (-mean invstd) + C)
        const float2 Bx = f2(b_s * invstd);
        const float2 Cx = f2(LN ? fmaf(b_s, nmi_s, c_s) : c_s);
        for_each_chunk<TPR, CH>(g.gtid, nchunks, [&](int j, int coff) {
          float2 dx[H];
          if constexpr (CACHE) {
#pragma unroll
            for (int k = 0; k < H; ++k) dx[k] = f2fma(A, gv[j][k], f2fma(B, xv[j][k], C));
          } else {
            float2 dy[H], x[H], w[H];
            IO<T>::unpack2(ld_smem_u4(dyrow + 16 * coff), dy);
            IO<T>::unpack2(ld_smem_u4(xrow + 16 * coff), x);
            IO<T>::unpack2(wq[j], w);
#pragma unroll
            for (int k = 0; k < H; ++k) {
              if constexpr (FMA_XHAT || !LN) {
                dx[k] = f2fma(A, f2mul(dy[k], w[k]), f2fma(Bx, x[k], Cx));
              } else {
                dx[k] = f2fma(A, f2mul(dy[k], w[k]), f2fma(B, x_hat(x[k]), C));
              }
            }
          }
          if constexpr (DROP) {
            st_global_stream(drrow + 16 * coff, IO<T>::pack2(dx));  // d residual = d x_eff
#pragma unroll
            for (int q = 0; q < VEC / 4; ++q) {
              const uint32_t mb = mbits[j][q];
              dx[2 * q].x = (mb & 0xffu) ? dx[2 * q].x * p.drop_scale : 0.f;
              dx[2 * q].y = (mb & 0xff00u) ? dx[2 * q].y * p.drop_scale : 0.f;
              dx[2 * q + 1].x = (mb & 0xff0000u) ? dx[2 * q + 1].x * p.drop_scale : 0.f;
              dx[2 * q + 1].y = (mb & 0xff000000u) ? dx[2 * q + 1].y * p.drop_scale : 0.f;
            }
          }
          st_global_stream(dxrow + 16 * coff, IO<T>::pack2(dx));
        });
      }
    }
    if (!CACHE || !want_dx) {  // the stage was read until the end of the tile
      g.sync();
      ring.refill();
    }
  }

  bwd_epilogue<T, TPR, CH, LN>(p, g, dw_acc, db_acc, want_dw, want_db);
  ring_teardown(p);
  tl_mark(p, 6);
}


// ---------------------------------------------------------------------------
// Pipelined re-read backward (rows too wide for the cached form: one or two 512-thread groups per
// CTA, e.g. C4 LayerNorm N = 12288, C5 RMSNorm N = 65536 over a cluster of 4).
// With a single group per CTA nothing overlaps a row's reduction: every warp waits at the group
// barrier (and, in a cluster, for the DSMEM exchange) between the two passes. Here the reduction is
// split-phase and software-pipelined across rows:
//     iteration k:  collect1(k)   totals of row k inside the CTA (posted an iteration ago: no wait);
//                                 cluster: send them to the peers
//                   pass1(k+1)    first pass of the NEXT row (sums, dgamma / dbeta accumulation), post
//                   collect2(k)   cluster: the peers' totals have had a whole pass to arrive
//                   pass2(k)      dx of row k
//                   release       the row's stage; the first warp refills it when every warp has let go
// post / collect use mbarriers (arrive is non-blocking, the wait only needs the other warps'
// arrivals — a named barrier would need their waits too). Slot sets, reduction barriers and cluster
// barriers alternate between even and odd rows; a warp can be at most one row ahead of another.
// One row per tile (the planner guarantees it for these shapes).
// ---------------------------------------------------------------------------
template <typename T, int TPR, int CH, bool LN>
__global__ void __launch_bounds__(kMaxBlock, 1) norm_bwd_pipe_kernel(const RowParams p) {
  constexpr int VEC = IO<T>::VEC;
  constexpr int H = VEC / 2;
  constexpr int WPG = TPR / 32;
  static_assert(CH > 0 && TPR >= 64, "pipelined backward: fixed chunk ownership, more than one warp per group");
  const uint32_t smem0 = smem_u32(smem_raw);
  tl_mark(p, 0);
  if (threadIdx.x == 0) {
    const int n_red = 2 * p.groups, n_free = p.groups * p.depth;
    for (int i = 0; i < n_red + n_free; ++i) mbar_init(smem0 + p.off_pipe + 8u * (uint32_t)i, WPG);
    for (int i = 0; i < p.groups; ++i) mbar_init(smem0 + p.off_pipe + 8u * (uint32_t)(n_red + n_free + i), 1);
  }
  ring_setup(p, smem0);  // fences the barrier inits, syncs the CTA (and the cluster)
  tl_mark(p, 1);
  Group<TPR> g(p, smem0);
  Ring<T, TPR> ring(p, smem0, g);
  ring.prologue();

  const int nchunks = p.nchunks;
  const bool want_dx = (p.flags & kFlagOutDx) != 0;
  const bool want_dw = (p.flags & kFlagOutDw) != 0;
  const bool want_db = LN && (p.flags & kFlagOutDb) != 0;
  const bool centered = LN || (p.flags & kFlagRmsBwdCentered) != 0;
  const bool stat_rms = !LN && (p.flags & kFlagStatIsRms) != 0;
  const size_t col0_bytes = (size_t)g.crank * p.slice_bytes;
  uint4 wq[CH];
  load_params_to_regs<T, TPR, CH>(p.weight, col0_bytes, g.gtid, nchunks, 1.0f, wq);

  float2 dw_acc[CH][H];
  float2 db_acc[LN ? CH : 1][H];
#pragma unroll
  for (int j = 0; j < CH; ++j) {
#pragma unroll
    for (int k = 0; k < H; ++k) {
      dw_acc[j][k] = f2(0.f);
      if constexpr (LN) db_acc[j][k] = f2(0.f);
    }
  }

  // barriers and slot sets of this group
  const uint32_t red_bar = smem0 + p.off_pipe + 8u * (uint32_t)(2 * g.gid);
  const uint32_t free_bar = smem0 + p.off_pipe + 8u * (uint32_t)(2 * p.groups + g.gid * p.depth);
  const uint32_t cl_bar_even = g.clbar;
  const uint32_t cl_bar_odd = smem0 + p.off_pipe + 8u * (uint32_t)(2 * p.groups + p.groups * p.depth + g.gid);
  float* red_slots = reinterpret_cast<float*>(smem_raw + p.off_red + (uint32_t)g.gid * (2 * kRedSetBytes));
  const uint32_t cls_off = p.off_cls + (uint32_t)g.gid * (2 * kClsSetBytes);

  const int n = ring.ntiles_g;
  const int64_t row_first = ring.row0, row_step = ring.row_step;
  auto stat_of = [&](int64_t k, float& mean, float& st) {
    const int64_t row = row_first + k * row_step;
    mean = 0.f;
    st = 0.f;
    if (k < n) {
      if constexpr (LN) mean = __ldg(p.stat0 + row);
      st = __ldg(p.stat1 + row);
    }
  };
  // row statistics: A = row k (second pass), B = row k+1 (first pass), C = row k+2 (in flight)
  float mean_a, st_a, mean_b, st_b, mean_c, st_c;
  stat_of(0, mean_b, st_b);
  stat_of(1, mean_c, st_c);

  // two cursors over the ring: c1 = first pass (one row ahead), c2 = second pass
  uint32_t c1_off = ring.ring_off_g, c1_bar = ring.bar_g, c1_s = 0, c1_phase = 0;
  uint32_t c2_off = ring.ring_off_g, c2_bar = ring.bar_g, c2_s = 0, c2_phase = 0;
  auto advance = [&](uint32_t& off, uint32_t& bar, uint32_t& s, uint32_t& phase) {
    ++s;
    off += p.stage_bytes;
    bar += 8u;
    if (s == (uint32_t)p.depth) {
      s = 0;
      phase ^= 1u;
      off = ring.ring_off_g;
      bar = ring.bar_g;
    }
  };

  // first pass of one row + post of its partial sums
  auto pass1_post = [&](int j_row, float mean, float st) {
    mbar_wait(c1_bar, c1_phase);
    const uint8_t* dyrow = smem_raw + c1_off + 16u * (uint32_t)g.gtid;
    const uint8_t* xrow = dyrow + p.stream_bytes;
    const float invstd = stat_rms ? rcp_fast(st) : st;
    constexpr bool FMA_XHAT = LN && sizeof(T) == 2;
    const float2 is2 = f2(invstd), nmean = f2(-mean), nmi = f2(-mean * invstd);
    float2 s1 = f2(0.f), s2 = f2(0.f);
    for_each_chunk<TPR, CH>(g.gtid, nchunks, [&](int j, int coff) {
      float2 dy[H], x[H], w[H];
      IO<T>::unpack2(ld_smem_u4(dyrow + 16 * coff), dy);
      IO<T>::unpack2(ld_smem_u4(xrow + 16 * coff), x);
      IO<T>::unpack2(wq[j], w);
#pragma unroll
      for (int k = 0; k < H; ++k) {
        const float2 gg = f2mul(dy[k], w[k]);
        float2 xh;
        if constexpr (FMA_XHAT) {
          xh = f2fma(x[k], is2, nmi);
        } else if constexpr (LN) {
          xh = f2mul(f2add(x[k], nmean), is2);
        } else {
          xh = f2mul(x[k], is2);
        }
        s1 = f2add(s1, gg);
        s2 = f2fma(gg, xh, s2);
        dw_acc[j][k] = f2fma(dy[k], xh, dw_acc[j][k]);
        if constexpr (LN) db_acc[j][k] = f2add(db_acc[j][k], dy[k]);
      }
    });
    float v[2] = {s1.x + s1.y, s2.x + s2.y};
    warp_allreduce<SumOp, 2>(v);
    if (g.lane == 0) {
      float* set = red_slots + (j_row & 1) * (kRedSetBytes / 4);
      set[g.warp] = v[0];
      set[kMaxWarpsPerGroup + g.warp] = v[1];
      mbar_arrive(red_bar + 8u * (uint32_t)(j_row & 1));  // release: the slots are visible to whoever sees the phase
    }
    advance(c1_off, c1_bar, c1_s, c1_phase);
  };

  if (n > 0) pass1_post(0, mean_b, st_b);
  mean_a = mean_b;
  st_a = st_b;
  mean_b = mean_c;
  st_b = st_c;
  stat_of(2, mean_c, st_c);

  for (int k = 0; k < n; ++k) {
    if (k == 0) tl_mark(p, 2);
    const uint32_t par = (uint32_t)(k & 1), ph = (uint32_t)((k >> 1) & 1);
    const uint32_t cl_bar = par ? cl_bar_odd : cl_bar_even;
    // ---- collect1(k): totals inside the CTA (fixed order: every thread gets the same bits) ----
    mbar_wait(red_bar + 8u * par, ph);
    float tot[2];
    {
      const float* set = red_slots + par * (kRedSetBytes / 4);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        float r = 0.f;
#pragma unroll
        for (int w4 = 0; w4 < (WPG + 3) / 4; ++w4) {
          const float4 t = *reinterpret_cast<const float4*>(set + i * kMaxWarpsPerGroup + 4 * w4);
          if constexpr (WPG >= 4) {
            r += (t.x + t.y) + (t.z + t.w);
          } else {
            r += t.x + t.y;  // WPG == 2
          }
        }
        tot[i] = r;
      }
    }
    if (g.csize > 1) {
      const uint32_t set = par * kClsSetBytes;
      if (g.gtid == 0) mbar_arrive_expect_tx(cl_bar, g.csize * 2 * 4);
      if ((uint32_t)g.gtid < g.csize) {
        const uint32_t r = (uint32_t)g.gtid;
        const uint32_t rbar = map_to_cta(cl_bar, r);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          st_async_f32(map_to_cta(smem0 + cls_off + set + (g.crank * kRedNV + i) * 4, r), tot[i], rbar);
        }
      }
    }
    // ---- first pass of the next row ----
    if (k + 1 < n) pass1_post(k + 1, mean_b, st_b);
    // ---- collect2(k): the peers' totals ----
    if (g.csize > 1) {
      mbar_wait(cl_bar, ph);
      const volatile float* slots = reinterpret_cast<const volatile float*>(smem_raw + cls_off + par * kClsSetBytes);
      tot[0] = 0.f;
      tot[1] = 0.f;
      for (uint32_t r = 0; r < g.csize; ++r) {
        tot[0] += slots[r * kRedNV + 0];
        tot[1] += slots[r * kRedNV + 1];
      }
    }
    // ---- second pass of row k ----
    if (want_dx) {
      const float invstd = stat_rms ? rcp_fast(st_a) : st_a;
      constexpr bool FMA_XHAT = LN && sizeof(T) == 2;
      const float nmi_s = -mean_a * invstd;
      const float a_n = invstd * p.inv_n;
      const float b_s = -(a_n * tot[1]) * p.proj_scale;
      const float c_s = centered ? -(a_n * tot[0]) : 0.f;
      const float2 A = f2(invstd), B = f2(b_s), C = f2(c_s), is2 = f2(invstd), nmean = f2(-mean_a);
      const float2 Bx = f2(b_s * invstd);
      const float2 Cx = f2(LN ? fmaf(b_s, nmi_s, c_s) : c_s);
      const uint8_t* dyrow = smem_raw + c2_off + 16u * (uint32_t)g.gtid;
      const uint8_t* xrow = dyrow + p.stream_bytes;
      uint8_t* dxrow = out_row_ptr<T>(p, row_first + (int64_t)k * row_step, g.crank, g.gtid);
      for_each_chunk<TPR, CH>(g.gtid, nchunks, [&](int j, int coff) {
        float2 dy[H], x[H], w[H], dx[H];
        IO<T>::unpack2(ld_smem_u4(dyrow + 16 * coff), dy);
        IO<T>::unpack2(ld_smem_u4(xrow + 16 * coff), x);
        IO<T>::unpack2(wq[j], w);
#pragma unroll
        for (int q = 0; q < H; ++q) {
          if constexpr (FMA_XHAT || !LN) {
            dx[q] = f2fma(A, f2mul(dy[q], w[q]), f2fma(Bx, x[q], Cx));
          } else {
            dx[q] = f2fma(A, f2mul(dy[q], w[q]), f2fma(B, f2mul(f2add(x[q], nmean), is2), C));
          }
        }
        st_global_stream(dxrow + 16 * coff, IO<T>::pack2(dx));
      });
    }
    // ---- release the stage; the first warp refills it once every warp has let go ----
    __syncwarp();
    if (g.lane == 0) mbar_arrive(free_bar + 8u * c2_s);
    if (g.warp == 0) {
      mbar_wait(free_bar + 8u * c2_s, c2_phase);
      if (k + p.depth < n) ring.issue_rows(1, row_first + (int64_t)(k + p.depth) * row_step, c2_off, c2_bar);
    }
    advance(c2_off, c2_bar, c2_s, c2_phase);
    mean_a = mean_b;
    st_a = st_b;
    mean_b = mean_c;
    st_b = st_c;
    stat_of((int64_t)k + 3, mean_c, st_c);
  }

  bwd_epilogue<T, TPR, CH, LN>(p, g, dw_acc, db_acc, want_dw, want_db);
  ring_teardown(p);
  tl_mark(p, 6);
}

}  // namespace b200norm

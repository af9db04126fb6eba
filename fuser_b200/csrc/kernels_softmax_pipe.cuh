// kernels_softmax_pipe.cuh — softmax / log-softmax forward over rows split across a thread-block cluster
// (vocabulary-size rows: N = 65536 ... 524288 at 16 bits; the reference's cluster path for these sizes is
// csrc/scheduler/normalization_inner_non_tma.cpp:554-613, exercised by tests/cpp/test_cluster.cpp:158-211).
//
// The register-resident kernel (kernels_fwd.cuh: softmax_fwd_ring_kernel) runs a row as max -> exchange -> sum ->
// exchange -> write, with every CTA of the cluster idle during both DSMEM round trips: 46 % of the HBM roofline at
// bf16 [4096,131072]. Here a row costs ONE exchange and the exchange is split-phase:
//   pass 1 (row k+1)  every warp takes the maximum m_w of its part of the slice and s_w = sum exp(x - m_w) (online-
//                     softmax pair relative to the WARP's maximum: no CTA-wide step before the exponentials);
//   combine (row k)   the warps' pairs -> the CTA's pair -> one st.async per peer -> the row's pair (M, S), each level
//                     as  M = max m_i,  S = sum s_i * exp(m_i - M)  (Group::softmax_rescale: equal maxima scale by
//                     exactly 1, so all -inf parts contribute 0 and an all -inf row ends as NaN like ATen's);
//   pass 2 (row k)    y = exp(x - M) / S   (or x - M - log S), x read a second time from the shared-memory stage.
// Pass 1 of row k+1 runs between sending row k's pair and waiting for the peers' pairs — the same pipeline as
// rmsnorm_fwd_pipe_kernel. The price is a second exponential per element (16-bit: one MUFU.EX2 each).
#pragma once

#include "ring.cuh"

namespace b200norm {

// A 4-byte piece of a row (two 16-bit elements or one fp32 element): the granularity of the balanced tail below.
template <typename T>
__device__ __forceinline__ float2 unpack_piece(uint32_t w) {
  if constexpr (sizeof(T) == 4) {
    return make_float2(__uint_as_float(w), __uint_as_float(w));  // both halves the same element: max / exp see it twice
  } else if constexpr (IO<T>::DT == 1) {
    return bf16x2_to_f2(w);
  } else {
    return __half22float2(*reinterpret_cast<const __half2*>(&w));
  }
}
template <typename T>
__device__ __forceinline__ uint32_t pack_piece(float2 v) {
  if constexpr (sizeof(T) == 4) {
    return __float_as_uint(v.x);
  } else {
    return IO<T>::pack2(v.x, v.y);
  }
}

// MODE 0: softmax, 1: log-softmax, 2: cross-entropy forward (REDUCE_INNER | kFlagReduceNll: lse = M + log S and
// loss = lse - x[target] per row, nothing per element — no pass 2, the stage is released after pass 1; the target's
// logit is read from the stage by the first thread of the CTA whose slice holds it, as in kernels_reduce.cuh).
template <typename T, int TPR, int CH, int MODE>
__global__ void __launch_bounds__(kMaxBlock, 1) softmax_fwd_pipe_kernel(const RowParams p) {
  constexpr bool LOG = MODE == 1;
  constexpr bool NLL = MODE == 2;
  constexpr int VEC = IO<T>::VEC;
  constexpr int H = VEC / 2;
  constexpr int WPG = TPR / 32;
  constexpr int NC = CH > 0 ? CH : 1;
  static_assert(TPR >= 64, "pipelined forward: more than one warp per group");
  const uint32_t smem0 = smem_u32(smem_raw);
  tl_mark(p, 0);
  if (threadIdx.x == 0) {
    const int n_red = 2 * p.groups, n_free = p.groups * p.depth;
    for (int i = 0; i < n_red + n_free; ++i) mbar_init(smem0 + p.off_pipe + 8u * (uint32_t)i, WPG);
    for (int i = 0; i < p.groups; ++i) mbar_init(smem0 + p.off_pipe + 8u * (uint32_t)(n_red + n_free + i), 1);
  }
  ring_setup(p, smem0);
  tl_mark(p, 1);
  Group<TPR> g(p, smem0);
  Ring<T, TPR> ring(p, smem0, g);
  ring.prologue();

  const int nchunks = p.nchunks;
  // Balanced tail. A slice of CH * TPR + t chunks with a small t (N = 100352 over four CTAs: 6 * 512 + 64) would make
  // t threads — the first warps, the producer warp among them — walk one more chunk than everybody else in both
  // passes, and every row waits for them at its barriers. The planner then hands this kernel CH = floor(chunks / TPR)
  // and the t chunks behind the full steps are walked in 4-byte pieces by ALL threads (at most kTailIters each).
  constexpr int kTailIters = 2;  // t <= 256 chunks = 1024 pieces
  const int tail_pieces = CH > 0 && nchunks > CH * TPR ? 4 * (nchunks - CH * TPR) : 0;
  const uint32_t tail_off = (uint32_t)CH * TPR * 16u + 4u * (uint32_t)g.gtid;  // byte offset of this thread's first piece
  const uint32_t red_bar = smem0 + p.off_pipe + 8u * (uint32_t)(2 * g.gid);
  const uint32_t free_bar = smem0 + p.off_pipe + 8u * (uint32_t)(2 * p.groups + g.gid * p.depth);
  const uint32_t cl_bar_even = g.clbar;
  const uint32_t cl_bar_odd = smem0 + p.off_pipe + 8u * (uint32_t)(2 * p.groups + p.groups * p.depth + g.gid);
  float* red_slots = reinterpret_cast<float*>(smem_raw + p.off_red + (uint32_t)g.gid * (2 * kRedSetBytes));
  const uint32_t cls_off = p.off_cls + (uint32_t)g.gid * (2 * kClsSetBytes);

  const int n = ring.ntiles_g;
  const int64_t row_first = ring.row0, row_step = ring.row_step;
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
  // (m_i, s_i) held by lane i < cnt -> (max m_i, sum s_i * exp(m_i - max)), the same bits in every lane
  auto merge_lanes = [&](float& m, float& s) {
    float wm = m;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) wm = MaxOp::apply(wm, __shfl_xor_sync(0xffffffffu, wm, off));
    s *= Group<TPR>::template softmax_rescale<T>(m, wm);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    m = wm;
  };
  // cross-entropy: (target logit, owner CTA or -1 for an ignored / out-of-range target) of the row pass 1 last saw
  float xt_next = 0.f;
  int owner_next = -1;
  long long t_ahead = 0;  // thread 0: the target index one row ahead of pass 1 (hides the global load)
  if constexpr (NLL) {
    if (g.gtid == 0 && ring.ntiles_g > 0) t_ahead = p.target[ring.row0];
  }
  auto pass1_post = [&](int j_row) {
    mbar_wait(c1_bar, c1_phase);
    const uint8_t* xrow = smem_raw + c1_off + 16u * (uint32_t)g.gtid;
    if constexpr (NLL) {
      if (g.gtid == 0) {
        const long long t = t_ahead;
        if (j_row + 1 < ring.ntiles_g) t_ahead = p.target[ring.row0 + (int64_t)(j_row + 1) * ring.row_step];
        const bool valid = t != (long long)p.ignore_index && t >= 0 && t < (long long)p.cols;
        owner_next = valid ? (int)(t / p.cols_cta) : -1;
        xt_next = 0.f;
        if (owner_next == (int)g.crank)
          xt_next = IO<T>::to_f32(reinterpret_cast<const T*>(xrow)[t - (long long)owner_next * p.cols_cta]);
      }
    }
    uint4 xq[NC];
    float m = -INFINITY;
    for_each_chunk<TPR, CH>(g.gtid, nchunks, [&](int j, int coff) {
      const uint4 q = ld_smem_u4(xrow + 16 * coff);
      if constexpr (CH > 0) xq[j] = q;
      float2 t[H];
      IO<T>::unpack2(q, t);
#pragma unroll
      for (int k = 0; k < H; ++k) m = MaxOp::apply(m, MaxOp::apply(t[k].x, t[k].y));
    });
    uint32_t tq[kTailIters];
    if (tail_pieces > 0) {
      const uint8_t* trow = smem_raw + c1_off + tail_off;
#pragma unroll
      for (int i = 0; i < kTailIters; ++i) {
        if (g.gtid + i * TPR < tail_pieces) {
          tq[i] = *reinterpret_cast<const uint32_t*>(trow + 4 * i * TPR);
          const float2 t = unpack_piece<T>(tq[i]);
          m = MaxOp::apply(m, MaxOp::apply(t.x, t.y));
        }
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = MaxOp::apply(m, __shfl_xor_sync(0xffffffffu, m, off));
    // a part that holds nothing but -inf sums exp(-inf - 0) = 0 instead of exp(-inf + inf) = NaN
    const float2 nref = f2(m == -INFINITY ? 0.f : -m);
    float2 acc = f2(0.f);
    for_each_chunk<TPR, CH>(g.gtid, nchunks, [&](int j, int coff) {
      float2 t[H];
      if constexpr (CH > 0) {
        IO<T>::unpack2(xq[j], t);
      } else {
        IO<T>::unpack2(ld_smem_u4(xrow + 16 * coff), t);
      }
#pragma unroll
      for (int k = 0; k < H; ++k) {
        // (the SUM takes the approximate exponential for fp32 rows too: a term's relative error of 2^-22 is the sum's)
        const float2 d = f2add(t[k], nref);
        acc = f2add(acc, make_float2(exp_fast(d.x), exp_fast(d.y)));
      }
    });
    float s = acc.x + acc.y;
    if (tail_pieces > 0) {
#pragma unroll
      for (int i = 0; i < kTailIters; ++i) {
        if (g.gtid + i * TPR < tail_pieces) {
          const float2 d = f2add(unpack_piece<T>(tq[i]), nref);
          s += exp_fast(d.x);
          if constexpr (sizeof(T) == 2) s += exp_fast(d.y);
        }
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (g.lane == 0) {
      float* set = red_slots + (j_row & 1) * (kRedSetBytes / 4);
      set[g.warp] = m;
      set[kMaxWarpsPerGroup + g.warp] = s;
      mbar_arrive(red_bar + 8u * (uint32_t)(j_row & 1));
    }
    if constexpr (NLL) {  // nothing reads the stage again: release it, the producer warp refills it
      __syncwarp();
      if (g.lane == 0) mbar_arrive(free_bar + 8u * c1_s);
      if (g.warp == 0) {
        mbar_wait(free_bar + 8u * c1_s, c1_phase);
        if (j_row + p.depth < n) ring.issue_rows(1, row_first + (int64_t)(j_row + p.depth) * row_step, c1_off, c1_bar);
      }
    }
    advance(c1_off, c1_bar, c1_s, c1_phase);
  };

  if (n > 0) pass1_post(0);
  for (int k = 0; k < n; ++k) {
    if (k == 0) tl_mark(p, 2);
    const uint32_t par = (uint32_t)(k & 1), ph = (uint32_t)((k >> 1) & 1);
    const uint32_t cl_bar = par ? cl_bar_odd : cl_bar_even;
    mbar_wait(red_bar + 8u * par, ph);
    float M, S;
    {
      const float* set = red_slots + par * (kRedSetBytes / 4);
      M = g.lane < WPG ? set[g.lane] : -INFINITY;
      S = g.lane < WPG ? set[kMaxWarpsPerGroup + g.lane] : 0.f;
      merge_lanes(M, S);
    }
    if (g.csize > 1) {
      if (g.gtid == 0) mbar_arrive_expect_tx(cl_bar, g.csize * 8);
      if ((uint32_t)g.gtid < g.csize) {
        const uint32_t r = (uint32_t)g.gtid;
        const uint32_t slot = map_to_cta(smem0 + cls_off + par * kClsSetBytes + (g.crank * kRedNV) * 4, r);
        const uint32_t rbar = map_to_cta(cl_bar, r);
        st_async_f32(slot, M, rbar);
        st_async_f32(slot + 4u, S, rbar);
      }
    }
    const float xt = xt_next;
    const int owner = owner_next;
    if (k + 1 < n) pass1_post(k + 1);
    if (g.csize > 1) {
      mbar_wait(cl_bar, ph);
      const volatile float* slots = reinterpret_cast<const volatile float*>(smem_raw + cls_off + par * kClsSetBytes);
      const bool has = (uint32_t)g.lane < g.csize;
      M = has ? slots[g.lane * kRedNV] : -INFINITY;
      S = has ? slots[g.lane * kRedNV + 1] : 0.f;
      merge_lanes(M, S);
    }
    const int64_t row = row_first + (int64_t)k * row_step;
    if constexpr (NLL) {
      if (g.gtid == 0 && g.crank == (uint32_t)(owner < 0 ? 0 : owner)) {
        const float lse = M + logf(S);
        if (p.stat1 != nullptr) p.stat1[row] = owner >= 0 ? lse - xt : 0.f;
        if (p.stat0 != nullptr) p.stat0[row] = lse;
      }
      continue;
    }
    const float2 nmx = f2(-M);
    const float2 fin = f2(LOG ? -logf(S) : rcp_fast(S));
    const uint8_t* xrow = smem_raw + c2_off + 16u * (uint32_t)g.gtid;
    uint8_t* yrow = out_row_ptr<T>(p, row, g.crank, g.gtid);
    for_each_chunk<TPR, CH>(g.gtid, nchunks, [&](int, int coff) {
      float2 t[H], y[H];
      IO<T>::unpack2(ld_smem_u4(xrow + 16 * coff), t);
#pragma unroll
      for (int q = 0; q < H; ++q) {
        const float2 d = f2add(t[q], nmx);
        y[q] = LOG ? f2add(d, fin) : f2mul(make_float2(exp_for<T>(d.x), exp_for<T>(d.y)), fin);
      }
      st_global_stream(yrow + 16 * coff, IO<T>::pack2(y));
    });
    if (tail_pieces > 0) {
      const uint8_t* trow = smem_raw + c2_off + tail_off;
      uint8_t* ty = yrow - 16u * (uint32_t)g.gtid + tail_off;
#pragma unroll
      for (int i = 0; i < kTailIters; ++i) {
        if (g.gtid + i * TPR < tail_pieces) {
          const float2 d = f2add(unpack_piece<T>(*reinterpret_cast<const uint32_t*>(trow + 4 * i * TPR)), nmx);
          const float2 y = LOG ? f2add(d, fin) : f2mul(make_float2(exp_for<T>(d.x), exp_for<T>(d.y)), fin);
          __stcs(reinterpret_cast<uint32_t*>(ty + 4 * i * TPR), pack_piece<T>(y));
        }
      }
    }
    __syncwarp();
    if (g.lane == 0) mbar_arrive(free_bar + 8u * c2_s);
    if (g.warp == 0) {
      mbar_wait(free_bar + 8u * c2_s, c2_phase);
      if (k + p.depth < n) ring.issue_rows(1, row_first + (int64_t)(k + p.depth) * row_step, c2_off, c2_bar);
    }
    advance(c2_off, c2_bar, c2_s, c2_phase);
  }
  tl_mark(p, 3);
  ring_teardown(p);
  tl_mark(p, 6);
}

}  // namespace b200norm

// kernels_reduce.cuh — plain (non-persistent in the reference's terms) reductions on the TMA ring:
//   reduce_inner_ring_kernel : [M, N] -> [M]   sum | max (NaN-propagating amax) | fmax (NaN-ignoring)
//                                              | (mean, variance or M2): stand-alone var_mean / welford segments
//   colsum_ring_kernel       : [M, N] -> [N]   sum over the rows (the shape of dgamma / dbeta: the stand-alone
//                                              column sum the reference segments a wide backward into)
// These are the segments nvFuser's ReductionScheduler takes (csrc/scheduler/reduction.cpp:198-360; device side:
// runtime/block_reduction.cu:26-166 for the inner case, runtime/grid_reduction.cu:201-309 + grid_sync.cu:34-95 for the
// outer one). Per-element semantics follow ReductionOp::evaluate (csrc/ir/internal_nodes.cpp:1726-1778): at::sum in
// fp32 for reduced-precision inputs (the fusions cast up first), at::amax for Max, NaN-aware fmax for FMax; the
// result is rounded to the io dtype once, at the store.
// Both reuse the skeleton of the normalization kernels: the inner reduction is a forward-type ring kernel without
// an output pass (dynamic tile pool, cluster split for rows too wide for one CTA); the column sum keeps per-thread
// fp32 column accumulators for the CTA's lifetime and finishes through bwd_epilogue — partial rows, grid barrier,
// fixed-order finish, deterministic — exactly like the backward's dgamma.
#pragma once

#include "kernels_bwd.cuh"
#include "ring.cuh"

namespace b200norm {

// KIND is a runtime flag (p.flags): the three flavours share one instantiation per (T, TPR, CH).
template <typename T, int TPR, int CH>
__global__ void __launch_bounds__(kMaxBlock, 1) reduce_inner_ring_kernel(const RowParams p) {
  constexpr int VEC = IO<T>::VEC;
  constexpr int H = VEC / 2;
  const uint32_t smem0 = smem_u32(smem_raw);
  tl_mark(p, 0);
  ring_setup(p, smem0);
  tl_mark(p, 1);
  Group<TPR> g(p, smem0);
  Ring<T, TPR, true> ring(p, smem0, g);
  ring.prologue();
  const int nchunks = p.nchunks;
  const bool is_max = (p.flags & kFlagReduceMax) != 0;
  const bool is_fmax = (p.flags & kFlagReduceFmax) != 0;
  const bool stats = (p.flags & kFlagNoY) != 0;  // LAYERNORM_FWD | NO_Y routed here: mean + variance / M2 per row
  const bool nll = (p.flags & kFlagReduceNll) != 0;  // cross-entropy forward: log-sum-exp and the target's logit
  const bool writer = g.gtid == 0 && g.crank == 0;
  T* out = static_cast<T*>(p.out0);

  for (; ring.acquire(); ring.advance()) {
    if (ring.k == 0) tl_mark(p, 2);
    const int nrows = ring.cur_nrows;
    const uint8_t* xrow = smem_raw + ring.stage_off + 16u * (uint32_t)g.gtid;
    for (int r = 0; r < nrows; ++r, xrow += p.slice_bytes) {
      float res;
      if (stats) {
        // stand-alone var_mean / welford (kFlagNoY): per thread sums of (x - K) and (x - K)^2 around K = the
        // thread's own first element — the cancellation in M2_t = S2 - S1^2 / n_t is then bounded by the handful
        // of elements a thread holds (S2 / M2_t <= n_t + 1) — turned into a Welford triple and merged exactly.
        float2 s1 = f2(0.f), s2 = f2(0.f);
        float cnt = 0.f, shift = 0.f;
        bool first = true;
        for_each_chunk<TPR, CH>(g.gtid, nchunks, [&](int, int coff) {
          float2 t[H];
          IO<T>::unpack2(ld_smem_u4(xrow + 16 * coff), t);
          if (first) {
            shift = t[0].x;
            first = false;
          }
          const float2 nk = f2(-shift);
#pragma unroll
          for (int k = 0; k < H; ++k) {
            const float2 dv = f2add(t[k], nk);
            s1 = f2add(s1, dv);
            s2 = f2fma(dv, dv, s2);
          }
          cnt += (float)VEC;
        });
        const float S1 = s1.x + s1.y, S2 = s2.x + s2.y;
        const float inv = cnt > 0.f ? rcp_fast(cnt) : 0.f;
        float mean = fmaf(S1, inv, shift);
        float m2 = fmaxf(fmaf(-S1 * inv, S1, S2), 0.f);
        g.welford_allreduce(cnt, mean, m2);
        if (r == nrows - 1) ring.refill();
        if (writer) {
          if (p.stat0 != nullptr) p.stat0[ring.row0 + r] = mean;
          if (p.stat1 != nullptr) p.stat1[ring.row0 + r] = ln_second_stat(p, m2, rsqrt_fast(m2 * p.inv_n + p.eps));
        }
        continue;
      }
      if (nll) {
        // cross-entropy forward: lse = max + log(sum exp(x - max)), loss = lse - x[target]. The target's logit is
        // read from the stage before the stage can be refilled, by the first thread of the CTA whose slice holds it.
        const long long t = p.target[ring.row0 + r];
        const bool valid = t != (long long)p.ignore_index && t >= 0 && t < (long long)p.cols;
        const uint32_t owner = valid ? (uint32_t)(t / p.cols_cta) : 0u;
        float xt = 0.f;
        if (g.gtid == 0 && valid && g.crank == owner)
          xt = IO<T>::to_f32(reinterpret_cast<const T*>(xrow)[t - (long long)owner * p.cols_cta]);
        // ONE pass over the stage and ONE exchange per row: every thread keeps an online (max, sum of exp) pair over
        // its chunks — the sum is rescaled when a chunk raises the maximum — and the pairs are merged exactly
        // (Group::softmax_allreduce). Nothing is written per element, so unlike the softmax forward (where this form
        // measured slower) the rescaling is per chunk and thread, not per output.
        float mt = -INFINITY, st = 0.f;
        for_each_chunk<TPR, CH>(g.gtid, nchunks, [&](int, int coff) {
          float2 tt[H];
          IO<T>::unpack2(ld_smem_u4(xrow + 16 * coff), tt);
          // (plain fmaxf: a NaN logit reaches the result through exp(NaN) in the sum, no NaN-aware maximum needed)
          float cm = fmaxf(tt[0].x, tt[0].y);
#pragma unroll
          for (int k = 1; k < H; ++k) cm = fmaxf(cm, fmaxf(tt[k].x, tt[k].y));
          const float mn = fmaxf(mt, cm);
          float2 acc = f2(0.f);
          if constexpr (sizeof(T) == 2) {
            // exp(x - m) = 2^(x * log2e - m * log2e): one packed FMA per pair in front of the MUFU
            constexpr float kLog2e = 1.4426950408889634f;
            const float2 nm = f2(mn == -INFINITY ? 0.f : -mn * kLog2e);  // all -inf so far: 2^(-inf - 0) = 0, not 2^NaN
            const float2 l2 = f2(kLog2e);
#pragma unroll
            for (int k = 0; k < H; ++k) {
              const float2 d = f2fma(tt[k], l2, nm);
              float e0, e1;
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(d.x));
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(d.y));
              acc = f2add(acc, make_float2(e0, e1));
            }
          } else {
            const float2 nm = f2(mn == -INFINITY ? 0.f : -mn);
#pragma unroll
            for (int k = 0; k < H; ++k) {
              const float2 d = f2add(tt[k], nm);
              acc = f2add(acc, make_float2(exp_for<T>(d.x), exp_for<T>(d.y)));
            }
          }
          st = fmaf(st, Group<TPR>::template softmax_rescale<T>(mt, mn), acc.x + acc.y);
          mt = mn;
        });
        g.template softmax_allreduce<T>(mt, st);
        float m[1] = {mt}, s[1] = {st};
        if (r == nrows - 1) ring.refill();
        const float lse = m[0] + logf(s[0]);
        if (g.gtid == 0 && g.crank == owner) {
          if (p.stat1 != nullptr) p.stat1[ring.row0 + r] = valid ? lse - xt : 0.f;
          if (p.stat0 != nullptr) p.stat0[ring.row0 + r] = lse;
        }
        continue;
      }
      if (is_max) {
        float m[1] = {-INFINITY};
        for_each_chunk<TPR, CH>(g.gtid, nchunks, [&](int, int coff) {
          float2 t[H];
          IO<T>::unpack2(ld_smem_u4(xrow + 16 * coff), t);
#pragma unroll
          for (int k = 0; k < H; ++k) m[0] = MaxOp::apply(m[0], MaxOp::apply(t[k].x, t[k].y));
        });
        g.template allreduce<MaxOp, 1>(m);
        res = m[0];
      } else if (is_fmax) {
        // NaN-ignoring maximum; a row of NaNs only is NaN (the count of non-NaN elements decides)
        float m[1] = {-INFINITY};
        float cnt[1] = {0.f};
        for_each_chunk<TPR, CH>(g.gtid, nchunks, [&](int, int coff) {
          float2 t[H];
          IO<T>::unpack2(ld_smem_u4(xrow + 16 * coff), t);
#pragma unroll
          for (int k = 0; k < H; ++k) {
            m[0] = fmaxf(m[0], fmaxf(t[k].x, t[k].y));
            cnt[0] += (t[k].x == t[k].x ? 1.f : 0.f) + (t[k].y == t[k].y ? 1.f : 0.f);
          }
        });
        g.template allreduce<FMaxOp, 1>(m);
        g.template allreduce<SumOp, 1>(cnt);
        res = cnt[0] > 0.f ? m[0] : __int_as_float(0x7fc00000);
      } else {
        float2 acc = f2(0.f);
        for_each_chunk<TPR, CH>(g.gtid, nchunks, [&](int, int coff) {
          float2 t[H];
          IO<T>::unpack2(ld_smem_u4(xrow + 16 * coff), t);
#pragma unroll
          for (int k = 0; k < H; ++k) acc = f2add(acc, t[k]);
        });
        float s[1] = {acc.x + acc.y};
        g.template allreduce<SumOp, 1>(s);
        res = s[0];
      }
      // the allreduce is a barrier for the group: after the tile's last row the stage is free
      if (r == nrows - 1) ring.refill();
      if (writer) out[ring.row0 + r] = IO<T>::from_f32(res);
    }
  }
  tl_mark(p, 3);
  ring.finish();
  ring_teardown(p);
  tl_mark(p, 6);
}

// Column sums: thread gtid owns chunks gtid + j*TPR (j < CH) of the CTA's column slice and adds every row of its
// tiles into CH*VEC fp32 accumulators; bwd_epilogue (LN = false: one plane) combines the groups, writes the CTA's
// partial row and finishes the sums in the kernel (or leaves them to finalize_partials_kernel).
template <typename T, int TPR, int CH>
__global__ void __launch_bounds__(kMaxBlock, 1) colsum_ring_kernel(const RowParams p) {
  constexpr int VEC = IO<T>::VEC;
  constexpr int H = VEC / 2;
  static_assert(CH > 0, "column accumulators need a fixed chunk ownership");
  const uint32_t smem0 = smem_u32(smem_raw);
  tl_mark(p, 0);
  ring_setup(p, smem0);
  tl_mark(p, 1);
  Group<TPR> g(p, smem0);
  Ring<T, TPR> ring(p, smem0, g);  // static schedule: the per-CTA partial sums must not depend on timing
  ring.prologue();
  const int nchunks = p.nchunks;

  float2 acc[CH][H];
  float2 none[1][H];
#pragma unroll
  for (int j = 0; j < CH; ++j) {
#pragma unroll
    for (int k = 0; k < H; ++k) acc[j][k] = f2(0.f);
  }
  for (; ring.more(); ring.advance()) {
    ring.wait();
    if (ring.k == 0) tl_mark(p, 2);
    const int nrows = ring.nrows();
    const uint8_t* xrow = smem_raw + ring.stage_off + 16u * (uint32_t)g.gtid;
    for (int r = 0; r < nrows; ++r, xrow += p.slice_bytes) {
      for_each_chunk<TPR, CH>(g.gtid, nchunks, [&](int j, int coff) {
        float2 t[H];
        IO<T>::unpack2(ld_smem_u4(xrow + 16 * coff), t);
#pragma unroll
        for (int k = 0; k < H; ++k) acc[j][k] = f2add(acc[j][k], t[k]);
      });
    }
    g.sync();  // every thread of the group has read the stage
    ring.refill();
  }
  bwd_epilogue<T, TPR, CH, false>(p, g, acc, none, true, false);
  ring_teardown(p);
  tl_mark(p, 6);
}

}  // namespace b200norm

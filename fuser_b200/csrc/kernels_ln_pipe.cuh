// kernels_ln_pipe.cuh — LayerNorm forward with ONE reduction round per row, software-pipelined across rows.
//
// Reference math: layer_norm, csrc/ops/normalization.cpp:304-342 — Welford over the normalized axis
// (runtime/welford.cu:52-68 is the combine the reference's kernels use), var = M2 / N, invstd = rsqrt(var + eps),
// y = (x - mean) * invstd * w + b. The register-resident kernel (kernels_fwd.cuh) computes the statistics in
// two passes with two group-wide reductions per row — mean first, then sum((x - mean)^2) — which is what the
// segmenter's TranslateApplicableWelford turns Welford into for persistent kernels (fusion_segmenter.cpp:2341).
// With one consumer group per CTA nothing overlaps those two round trips (ncu: issue active 41 %, [32768,12288]
// bf16 at 72 % of the measured peak where RMSNorm forward, one reduction, reaches 97 %).
//
// Here every thread reduces ITS elements to a Welford pair (mean_t, M2_t) in registers — a local two-pass, no
// synchronisation — and pairs are merged with Chan's parallel formula (the welfordCombine of the reference):
// all partial sets have the same element count, so a merge of two sets of n elements is
//     mean = (mean_a + mean_b) / 2,   M2 = M2_a + M2_b + (mean_a - mean_b)^2 * n / 2
// (symmetric in a and b: both shuffle partners compute the same bits), across the lanes of a warp with xor
// shuffles; the warps' pairs (and, over a cluster, the CTAs' pairs) are combined in closed form,
//     mean = avg(mean_i),   M2 = sum(M2_i) + n_i * sum((mean_i - mean)^2).
// That is ONE exchange of two floats per row, so the kernel takes the split-phase pipeline of
// rmsnorm_fwd_pipe_kernel: warps post their pair on an mbarrier (non-blocking), run the NEXT row's statistics,
// and only then collect; the row stays in its shared-memory stage and is read a second time for the output.
// Numerically this is at least as good as the two-pass form (no sum of N squares around a rounded mean).
//
// Requirements (the planner guarantees them): one row per tile, nchunks == TPR * CH exactly (equal counts),
// TPR >= 64.
#pragma once

#include "ring.cuh"

namespace b200norm {

template <typename T, int TPR, int CH, bool ADD = false>
__global__ void __launch_bounds__(kMaxBlock, 1) layernorm_fwd_pipe_kernel(const RowParams p) {
  constexpr int VEC = IO<T>::VEC;
  constexpr int H = VEC / 2;
  constexpr int WPG = TPR / 32;
  // 16-bit parameters are kept UNPACKED (fp32 pairs) in registers when they fit next to the row: the row is not
  // held across the reduction here, so 2 * CH * VEC registers are free (saves 16 unpack ops per chunk and row)
  constexpr bool UNPACKED = sizeof(T) == 2 && CH <= 3;
  static_assert(CH > 0 && TPR >= 64, "pipelined forward: fixed chunk ownership, more than one warp per group");
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

  const size_t col0_bytes = (size_t)g.crank * p.slice_bytes;
  uint4 wq[UNPACKED ? 1 : CH], bq[UNPACKED ? 1 : CH];
  float2 wf[UNPACKED ? CH : 1][H], bf[UNPACKED ? CH : 1][H];
  // instance norm (kFlagRowAffine): the parameters belong to the row's channel — the same registers are refilled
  // with a splat of that row's pair before each output pass, the arithmetic is untouched
  const bool row_aff = (p.flags & kFlagRowAffine) != 0;
  const void* wcol = row_aff ? nullptr : p.weight;
  const void* bcol = row_aff ? nullptr : p.bias;
  if constexpr (UNPACKED) {
    uint4 tw[CH], tb[CH];
    load_params_to_regs<T, TPR, CH>(wcol, col0_bytes, g.gtid, p.nchunks, 1.0f, tw);
    load_params_to_regs<T, TPR, CH>(bcol, col0_bytes, g.gtid, p.nchunks, -0.0f, tb);
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      IO<T>::unpack2(tw[j], wf[j]);
      IO<T>::unpack2(tb[j], bf[j]);
    }
  } else {
    load_params_to_regs<T, TPR, CH>(wcol, col0_bytes, g.gtid, p.nchunks, 1.0f, wq);
    load_params_to_regs<T, TPR, CH>(bcol, col0_bytes, g.gtid, p.nchunks, -0.0f, bq);
  }
  const bool stat_writer = g.gtid == 0 && g.crank == 0;
  const bool no_y = (p.flags & kFlagNoY) != 0;  // stand-alone var_mean / welford: statistics only
  const bool round_sum = ADD && (p.flags & kFlagRoundSum) != 0;
  const bool out_sum = ADD && p.out1 != nullptr;

  const uint32_t red_bar = smem0 + p.off_pipe + 8u * (uint32_t)(2 * g.gid);
  const uint32_t free_bar = smem0 + p.off_pipe + 8u * (uint32_t)(2 * p.groups + g.gid * p.depth);
  const uint32_t cl_bar_even = g.clbar;
  const uint32_t cl_bar_odd = smem0 + p.off_pipe + 8u * (uint32_t)(2 * p.groups + p.groups * p.depth + g.gid);
  float* red_slots = reinterpret_cast<float*>(smem_raw + p.off_red + (uint32_t)g.gid * (2 * kRedSetBytes));
  const uint32_t cls_off = p.off_cls + (uint32_t)g.gid * (2 * kClsSetBytes);

  const int n = ring.ntiles_g;
  const int64_t row_first = ring.row0, row_step = ring.row_step;
  uint8_t* yrow = out_row_ptr<T>(p, row_first, g.crank, g.gtid);  // running pointers: no 64-bit multiply per row
  const size_t y_step = (size_t)row_step * (size_t)p.out_stride * sizeof(T);
  int64_t row = row_first;
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

  constexpr float kLocalN = (float)(CH * VEC);       // elements per thread
  constexpr float kInvLocalN = 1.0f / kLocalN;
  constexpr float kWarpN = 32.0f * kLocalN;          // elements per warp
  const float cta_n = (float)p.cols_cta;             // elements per CTA slice

  // first pass of one row: the thread's Welford pair, merged over the warp, posted for the group
  auto pass1_post = [&](int j_row) {
    mbar_wait(c1_bar, c1_phase);
    const uint8_t* xrow = smem_raw + c1_off + 16u * (uint32_t)g.gtid;
    uint8_t* srow = nullptr;
    if constexpr (ADD) {
      if (out_sum) srow = out1_row_ptr<T>(p, row_first + (int64_t)j_row * row_step, g.crank, g.gtid);
    }
    float2 x[CH][H];
    float2 acc = f2(0.f);
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      IO<T>::unpack2(ld_smem_u4(xrow + 16 * j * TPR), x[j]);
      if constexpr (ADD) {
        add_residual_chunk<T, H>(x[j], xrow + p.stream_bytes + 16 * j * TPR, round_sum,
                                 srow != nullptr ? srow + 16 * j * TPR : nullptr);
      }
#pragma unroll
      for (int k = 0; k < H; ++k) acc = f2add(acc, x[j][k]);
    }
    float mean = (acc.x + acc.y) * kInvLocalN;
    const float2 nm = f2(-mean);
    acc = f2(0.f);
#pragma unroll
    for (int j = 0; j < CH; ++j) {
#pragma unroll
      for (int k = 0; k < H; ++k) {
        const float2 d = f2add(x[j][k], nm);
        acc = f2fma(d, d, acc);
      }
    }
    float m2 = acc.x + acc.y;
    // Chan merge over the lanes (equal counts: kLocalN * 2^level elements on each side)
    float half_n = 0.5f * kLocalN;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, mean, off);
      const float o2 = __shfl_xor_sync(0xffffffffu, m2, off);
      const float d = mean - om;
      mean = (mean + om) * 0.5f;
      m2 = fmaf(d * d, half_n, m2 + o2);
      half_n *= 2.0f;
    }
    if (g.lane == 0) {
      float* set = red_slots + (j_row & 1) * (kRedSetBytes / 4);
      set[g.warp] = mean;
      set[kMaxWarpsPerGroup + g.warp] = m2;
      mbar_arrive(red_bar + 8u * (uint32_t)(j_row & 1));
    }
    advance(c1_off, c1_bar, c1_s, c1_phase);
  };

  if (n > 0) pass1_post(0);
  for (int k = 0; k < n; ++k) {
    if (k == 0) tl_mark(p, 2);
    const uint32_t par = (uint32_t)(k & 1), ph = (uint32_t)((k >> 1) & 1);
    const uint32_t cl_bar = par ? cl_bar_odd : cl_bar_even;
    float w_row = 1.f, b_row = 0.f;  // requested before the waits below: off the critical path
    if (row_aff) {
      w_row = row_param<T>(p.weight, row, p.channels, 1.f);
      b_row = row_param<T>(p.bias, row, p.channels, 0.f);
    }
    // ---- collect1(k): the warps' pairs in closed form (fixed order: every thread gets the same bits) ----
    mbar_wait(red_bar + 8u * par, ph);
    float mean, m2;
    {
      const float* set = red_slots + par * (kRedSetBytes / 4);
      welford_combine_equal<WPG>(set, set + kMaxWarpsPerGroup, kWarpN, mean, m2);
    }
    if (g.csize > 1) {
      const uint32_t set = par * kClsSetBytes;
      if (g.gtid == 0) mbar_arrive_expect_tx(cl_bar, g.csize * 2 * 4);
      if ((uint32_t)g.gtid < g.csize) {
        const uint32_t r = (uint32_t)g.gtid;
        const uint32_t rbar = map_to_cta(cl_bar, r);
        st_async_f32(map_to_cta(smem0 + cls_off + set + (g.crank * kRedNV + 0) * 4, r), mean, rbar);
        st_async_f32(map_to_cta(smem0 + cls_off + set + (g.crank * kRedNV + 1) * 4, r), m2, rbar);
      }
    }
    // ---- statistics of the next row while the exchange is in flight ----
    if (k + 1 < n) pass1_post(k + 1);
    if (g.csize > 1) {
      mbar_wait(cl_bar, ph);
      const volatile float* slots = reinterpret_cast<const volatile float*>(smem_raw + cls_off + par * kClsSetBytes);
      float sm = 0.f, s2 = 0.f;
      for (uint32_t r = 0; r < g.csize; ++r) {
        sm += slots[r * kRedNV + 0];
        s2 += slots[r * kRedNV + 1];
      }
      mean = sm / (float)g.csize;
      float dev = 0.f;
      for (uint32_t r = 0; r < g.csize; ++r) {
        const float d = slots[r * kRedNV + 0] - mean;
        dev = fmaf(d, d, dev);
      }
      m2 = fmaf(dev, cta_n, s2);
    }
    const float invstd = rsqrt_fast(m2 * p.inv_n + p.eps);
    // ---- second pass of row k: y = (x - mean) * invstd * w + b ----
    const uint8_t* xrow = smem_raw + c2_off + 16u * (uint32_t)g.gtid;
    const float2 is2 = f2(invstd), nmean = f2(-mean), nmi = f2(-mean * invstd);
    if (row_aff) {
#pragma unroll
      for (int j = 0; j < (UNPACKED ? CH : 1); ++j) {
#pragma unroll
        for (int q = 0; q < H; ++q) {
          wf[j][q] = f2(w_row);
          bf[j][q] = f2(b_row);
        }
      }
      if constexpr (!UNPACKED) {
        const uint4 ws = splat_chunk<T>(w_row), bs = splat_chunk<T>(b_row);
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          wq[j] = ws;
          bq[j] = bs;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      if (no_y) break;
      float2 y[H], w[H], b[H];
      IO<T>::unpack2(ld_smem_u4(xrow + 16 * j * TPR), y);
      if constexpr (ADD) add_residual_chunk<T, H>(y, xrow + p.stream_bytes + 16 * j * TPR, round_sum, nullptr);
      if constexpr (!UNPACKED) {
        IO<T>::unpack2(wq[j], w);
        IO<T>::unpack2(bq[j], b);
      }
#pragma unroll
      for (int q = 0; q < H; ++q) {
        // 16-bit io: x_hat as one FMA, x * invstd + (-mean * invstd) — the pre-rounded product costs at most
        // |mean| / std fp32 ulps of x_hat, far below the resolution of the 16-bit output; fp32 io keeps
        // subtract-then-scale
        float2 xh;
        if constexpr (sizeof(T) == 2) {
          xh = f2fma(y[q], is2, nmi);
        } else {
          xh = f2mul(f2add(y[q], nmean), is2);
        }
        if constexpr (UNPACKED) {
          y[q] = f2fma(xh, wf[j][q], bf[j][q]);
        } else {
          y[q] = f2fma(xh, w[q], b[q]);
        }
      }
      st_global_stream(yrow + 16 * j * TPR, IO<T>::pack2(y));
    }
    if (stat_writer) {
      if (p.stat0 != nullptr) p.stat0[row] = mean;
      if (p.stat1 != nullptr) p.stat1[row] = ln_second_stat(p, m2, invstd);
    }
    // ---- release the stage; the first warp refills it once every warp has let go ----
    __syncwarp();
    if (g.lane == 0) mbar_arrive(free_bar + 8u * c2_s);
    if (g.warp == 0) {
      mbar_wait(free_bar + 8u * c2_s, c2_phase);
      if (k + p.depth < n) ring.issue_rows(1, row_first + (int64_t)(k + p.depth) * row_step, c2_off, c2_bar);
    }
    advance(c2_off, c2_bar, c2_s, c2_phase);
    yrow += y_step;
    row += row_step;
  }
  tl_mark(p, 3);
  ring_teardown(p);
  tl_mark(p, 6);
}

}  // namespace b200norm

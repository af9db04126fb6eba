// kernels_fwd.cuh — ring-pipelined kernels without per-column state:
//   RMSNorm fwd, LayerNorm fwd, softmax fwd/bwd, log-softmax fwd/bwd.
//
// Math follows the reference op builders (csrc/ops/normalization.cpp):
//   rms_norm :352-380, layer_norm :304-342 (Welford -> two-pass: mean, then
//   sum((x-mean)^2), which is what the segmenter's TranslateApplicableWelford does
//   for persistent kernels, fusion_segmenter.cpp:2341), softmax :169-189,
//   softmax_backward :191-210, log_softmax :212-231, log_softmax_backward :233-249.
// All arithmetic is fp32 (packed FFMA2/FMUL2/FADD2 where two lanes share an op);
// the only rounding to the io dtype is at the store.
//
// Each thread owns chunks c = gtid + j*TPR (j < CH) of the CTA's column slice: the row
// is read from the shared-memory ring ONCE into registers and every later pass runs on
// registers (the register-persistent buffer of the reference's inner-persistent
// schedule). CH == 0 is the looping form for slices wider than 4*TPR chunks: the stage
// is re-read from shared memory in each pass.
#pragma once

#include "ring.cuh"

namespace b200norm {

// ---------------------------------------------------------------------------
// RMSNorm forward: stat = rsqrt(mean(x^2)+eps) (or sqrt(...) with STAT_IS_RMS),
// y = x * invstd * w.
// ---------------------------------------------------------------------------
// DROP (with ADD, cached form): the x stream goes through a fused dropout before the residual is added
// (dropout_chunk, ring.cuh); its own instantiations, so that the plain kernels keep their register budget.
template <typename T, int TPR, int CH, bool ADD = false, bool DROP = false>
__global__ void __launch_bounds__(kMaxBlock, 1) rmsnorm_fwd_ring_kernel(const RowParams p) {
  static_assert(!DROP || (ADD && CH > 0), "fused dropout: norm(residual + dropout(x)) on the register-resident form");
  constexpr int VEC = IO<T>::VEC;
  constexpr int H = VEC / 2;
  constexpr int NC = CH > 0 ? CH : 1;
  const uint32_t smem0 = smem_u32(smem_raw);
  tl_mark(p, 0);
  ring_setup(p, smem0);
  tl_mark(p, 1);
  Group<TPR> g(p, smem0);
  Ring<T, TPR, true> ring(p, smem0, g);
  ring.prologue();

  const int nchunks = p.nchunks;
  const bool stat_rms = (p.flags & kFlagStatIsRms) != 0;
  const size_t col0_bytes = (size_t)g.crank * p.slice_bytes;
  uint4 wq[NC];
  load_params_to_regs<T, TPR, CH>(p.weight, col0_bytes, g.gtid, nchunks, 1.0f, wq);
  const bool stat_writer = g.gtid == 0 && g.crank == 0 && p.stat1 != nullptr;
  const bool round_sum = ADD && (p.flags & kFlagRoundSum) != 0;
  const bool out_sum = ADD && p.out1 != nullptr;
  constexpr bool drop = DROP;
  const unsigned long long drop_seed = drop ? p.rng[0] : 0ull, drop_offset = drop ? p.rng[1] : 0ull;

  const size_t out_row_bytes = (size_t)p.out_stride * sizeof(T);
  const size_t sum_row_bytes = (size_t)p.out1_stride * sizeof(T);
  for (; ring.acquire(); ring.advance()) {
    if (ring.k == 0) tl_mark(p, 2);
    const int nrows = ring.cur_nrows;
    const uint8_t* xrow = smem_raw + ring.stage_off + 16u * (uint32_t)g.gtid;
    uint8_t* yrow = out_row_ptr<T>(p, ring.row0, g.crank, g.gtid);
    uint8_t* srow = out_sum ? out1_row_ptr<T>(p, ring.row0, g.crank, g.gtid) : nullptr;
    for (int r = 0; r < nrows; ++r, xrow += p.slice_bytes, yrow += out_row_bytes, srow += sum_row_bytes) {
      float2 x[NC][H];
      float2 acc = f2(0.f);
      // fused dropout: linear index of the row's first element of this CTA's slice
      const unsigned long long drop_lin =
          (unsigned long long)(ring.row0 + r) * (unsigned long long)p.cols + (unsigned long long)(g.crank * (uint32_t)p.cols_cta);
      for_each_chunk<TPR, CH>(g.gtid, nchunks, [&](int j, int coff) {
        float2 t[H];
        IO<T>::unpack2(ld_smem_u4(xrow + 16 * coff), t);
        if constexpr (ADD) {
          if constexpr (drop) dropout_chunk<T, H>(t, p, drop_seed, drop_offset, drop_lin + (unsigned long long)((g.gtid + coff) * VEC));
          add_residual_chunk<T, H>(t, xrow + p.stream_bytes + 16 * coff, round_sum, out_sum ? srow + 16 * coff : nullptr);
        }
#pragma unroll
        for (int k = 0; k < H; ++k) {
          acc = f2fma(t[k], t[k], acc);
          if constexpr (CH > 0) x[j][k] = t[k];
        }
      });
      float ss[1] = {acc.x + acc.y};
      g.template allreduce<SumOp, 1>(ss);
      if (CH > 0 && r == nrows - 1) ring.refill();  // registers hold the row: the stage is free
      const float var_eps = ss[0] * p.inv_n + p.eps;
      // one MUFU on either path: invstd = rsqrt(v); the rms flavour saves sqrt(v) = v * rsqrt(v)
      const float invstd = rsqrt_fast(var_eps);
      const float rms = var_eps * invstd;
      const float2 scale = f2(invstd);
      for_each_chunk<TPR, CH>(g.gtid, nchunks, [&](int j, int coff) {
        float2 y[H], w[H];
        if constexpr (CH > 0) {
          IO<T>::unpack2(wq[j], w);
#pragma unroll
          for (int k = 0; k < H; ++k) y[k] = f2mul(f2mul(x[j][k], scale), w[k]);
        } else {
          IO<T>::unpack2(ld_smem_u4(xrow + 16 * coff), y);
          if constexpr (ADD) add_residual_chunk<T, H>(y, xrow + p.stream_bytes + 16 * coff, round_sum, nullptr);
          IO<T>::unpack2(load_param_chunk<T>(p.weight, col0_bytes, g.gtid + coff, 1.0f), w);
#pragma unroll
          for (int k = 0; k < H; ++k) y[k] = f2mul(f2mul(y[k], scale), w[k]);
        }
        st_global_stream(yrow + 16 * coff, IO<T>::pack2(y));
      });
      if (stat_writer) p.stat1[ring.row0 + r] = stat_rms ? rms : invstd;
    }
    if constexpr (CH == 0) {  // looping form: the stage was read until the last pass
      g.sync();
      ring.refill();
    }
  }
  tl_mark(p, 3);
  ring.finish();
  ring_teardown(p);
  tl_mark(p, 6);
}

// ---------------------------------------------------------------------------
// LayerNorm forward: mean, invstd = rsqrt(M2/N + eps), y = (x-mean)*invstd*w + b.
// ---------------------------------------------------------------------------
// PSMEM: weight and bias live in shared memory as fp32 (two LDS.128 per parameter and chunk)
// instead of packed in registers (16 unpack ops per chunk and 8*CH registers) — LayerNorm is
// the one forward with two per-column parameters, and unpacking both costs more issue slots
// than the row math itself.
template <typename T, int TPR, int CH, bool PSMEM = false, bool ADD = false, bool DROP = false>
__global__ void __launch_bounds__(kMaxBlock, 1) layernorm_fwd_ring_kernel(const RowParams p) {
  static_assert(!DROP || (ADD && CH > 0), "fused dropout: norm(residual + dropout(x)) on the register-resident form");
  constexpr int VEC = IO<T>::VEC;
  constexpr int H = VEC / 2;
  constexpr int NC = CH > 0 ? CH : 1;
  const uint32_t smem0 = smem_u32(smem_raw);
  tl_mark(p, 0);
  ring_setup(p, smem0);
  tl_mark(p, 1);
  Group<TPR> g(p, smem0);
  Ring<T, TPR, true> ring(p, smem0, g);
  ring.prologue();

  const int nchunks = p.nchunks;
  const size_t col0_bytes = (size_t)g.crank * p.slice_bytes;
  uint4 wq[PSMEM ? 1 : NC], bq[PSMEM ? 1 : NC];
  // instance norm (kFlagRowAffine; never with PSMEM): the parameter registers are refilled per row with a splat of the
  // row's channel pair
  const bool row_aff = !PSMEM && (p.flags & kFlagRowAffine) != 0;
  const uint32_t off_w = p.off_param;
  const uint32_t off_b = p.off_param + (uint32_t)nchunks * (uint32_t)(VEC * 4);
  if constexpr (PSMEM) {
    stage_param_smem<T>(p.weight, col0_bytes, off_w, nchunks, 1.0f);
    stage_param_smem<T>(p.bias, col0_bytes, off_b, nchunks, -0.0f);
    __syncthreads();
  } else {
    load_params_to_regs<T, TPR, CH>(row_aff ? nullptr : p.weight, col0_bytes, g.gtid, nchunks, 1.0f, wq);
    load_params_to_regs<T, TPR, CH>(row_aff ? nullptr : p.bias, col0_bytes, g.gtid, nchunks, -0.0f, bq);
  }
  const bool stat_writer = g.gtid == 0 && g.crank == 0;
  const bool no_y = (p.flags & kFlagNoY) != 0;  // stand-alone var_mean / welford: statistics only
  const bool round_sum = ADD && (p.flags & kFlagRoundSum) != 0;
  const bool out_sum = ADD && p.out1 != nullptr;
  constexpr bool drop = DROP;
  const unsigned long long drop_seed = drop ? p.rng[0] : 0ull, drop_offset = drop ? p.rng[1] : 0ull;

  const size_t out_row_bytes = (size_t)p.out_stride * sizeof(T);
  const size_t sum_row_bytes = (size_t)p.out1_stride * sizeof(T);
  for (; ring.acquire(); ring.advance()) {
    if (ring.k == 0) tl_mark(p, 2);
    const int nrows = ring.cur_nrows;
    const uint8_t* xrow = smem_raw + ring.stage_off + 16u * (uint32_t)g.gtid;
    uint8_t* yrow = out_row_ptr<T>(p, ring.row0, g.crank, g.gtid);
    uint8_t* srow = out_sum ? out1_row_ptr<T>(p, ring.row0, g.crank, g.gtid) : nullptr;
    for (int r = 0; r < nrows; ++r, xrow += p.slice_bytes, yrow += out_row_bytes, srow += sum_row_bytes) {
      float2 x[NC][H];
      float2 acc = f2(0.f);
      float w_row = 1.f, b_row = 0.f;
      if (row_aff) {
        w_row = row_param<T>(p.weight, ring.row0 + r, p.channels, 1.f);
        b_row = row_param<T>(p.bias, ring.row0 + r, p.channels, 0.f);
      }
      // fused dropout: linear index of the row's first element of this CTA's slice
      const unsigned long long drop_lin =
          (unsigned long long)(ring.row0 + r) * (unsigned long long)p.cols + (unsigned long long)(g.crank * (uint32_t)p.cols_cta);
      for_each_chunk<TPR, CH>(g.gtid, nchunks, [&](int j, int coff) {
        float2 t[H];
        IO<T>::unpack2(ld_smem_u4(xrow + 16 * coff), t);
        if constexpr (ADD) {
          if constexpr (drop) dropout_chunk<T, H>(t, p, drop_seed, drop_offset, drop_lin + (unsigned long long)((g.gtid + coff) * VEC));
          add_residual_chunk<T, H>(t, xrow + p.stream_bytes + 16 * coff, round_sum, out_sum ? srow + 16 * coff : nullptr);
        }
#pragma unroll
        for (int k = 0; k < H; ++k) {
          acc = f2add(acc, t[k]);
          if constexpr (CH > 0) x[j][k] = t[k];
        }
      });
      float s[1] = {acc.x + acc.y};
      g.template allreduce<SumOp, 1>(s);
      if (CH > 0 && r == nrows - 1) ring.refill();  // registers hold the row: the stage is free
      const float mean = s[0] * p.inv_n;
      const float2 nmean = f2(-mean);
      acc = f2(0.f);
      for_each_chunk<TPR, CH>(g.gtid, nchunks, [&](int j, int coff) {
        float2 t[H];
        if constexpr (CH == 0) {
          IO<T>::unpack2(ld_smem_u4(xrow + 16 * coff), t);
          if constexpr (ADD) add_residual_chunk<T, H>(t, xrow + p.stream_bytes + 16 * coff, round_sum, nullptr);
        }
#pragma unroll
        for (int k = 0; k < H; ++k) {
          if constexpr (CH > 0) {
            x[j][k] = f2add(x[j][k], nmean);  // x now holds x - mean
            acc = f2fma(x[j][k], x[j][k], acc);
          } else {
            const float2 d = f2add(t[k], nmean);
            acc = f2fma(d, d, acc);
          }
        }
      });
      float m2[1] = {acc.x + acc.y};
      g.template allreduce<SumOp, 1>(m2);
      const float invstd = rsqrt_fast(m2[0] * p.inv_n + p.eps);
      const float2 is2 = f2(invstd);
      if constexpr (!PSMEM && CH > 0) {
        if (row_aff) {
          const uint4 ws = splat_chunk<T>(w_row), bs = splat_chunk<T>(b_row);
#pragma unroll
          for (int j = 0; j < NC; ++j) {
            wq[j] = ws;
            bq[j] = bs;
          }
        }
      }
      if (!no_y) for_each_chunk<TPR, CH>(g.gtid, nchunks, [&](int j, int coff) {
        float2 y[H], w[H], b[H];
        if constexpr (CH > 0) {
          if constexpr (PSMEM) {
            load_param_smem<VEC>(off_w, nchunks, g.gtid + coff, w);
            load_param_smem<VEC>(off_b, nchunks, g.gtid + coff, b);
          } else {
            IO<T>::unpack2(wq[j], w);
            IO<T>::unpack2(bq[j], b);
          }
#pragma unroll
          for (int k = 0; k < H; ++k) y[k] = f2fma(f2mul(x[j][k], is2), w[k], b[k]);
        } else {
          IO<T>::unpack2(ld_smem_u4(xrow + 16 * coff), y);
          if constexpr (ADD) add_residual_chunk<T, H>(y, xrow + p.stream_bytes + 16 * coff, round_sum, nullptr);
          IO<T>::unpack2(row_aff ? splat_chunk<T>(w_row) : load_param_chunk<T>(p.weight, col0_bytes, g.gtid + coff, 1.0f), w);
          IO<T>::unpack2(row_aff ? splat_chunk<T>(b_row) : load_param_chunk<T>(p.bias, col0_bytes, g.gtid + coff, -0.0f), b);
#pragma unroll
          for (int k = 0; k < H; ++k) y[k] = f2fma(f2mul(f2add(y[k], nmean), is2), w[k], b[k]);
        }
        st_global_stream(yrow + 16 * coff, IO<T>::pack2(y));
      });
      if (stat_writer) {
        if (p.stat0 != nullptr) p.stat0[ring.row0 + r] = mean;
        if (p.stat1 != nullptr) p.stat1[ring.row0 + r] = ln_second_stat(p, m2[0], invstd);
      }
    }
    if constexpr (CH == 0) {
      g.sync();
      ring.refill();
    }
  }
  tl_mark(p, 3);
  ring.finish();
  ring_teardown(p);
  tl_mark(p, 6);
}

// ---------------------------------------------------------------------------
// softmax / log-softmax forward. x, then exp(x-max) (or x-max for the log flavour),
// live in registers — one MUFU per element. The looping form recomputes exp.
// NaN semantics: max is NaN-propagating (at::amax); any NaN / +inf / all -inf row
// turns into an all-NaN row exactly as the op-by-op ATen evaluation does.
// ---------------------------------------------------------------------------
template <typename T, int TPR, int CH, bool LOG>
__global__ void __launch_bounds__(kMaxBlock, 1) softmax_fwd_ring_kernel(const RowParams p) {
  constexpr int VEC = IO<T>::VEC;
  constexpr int H = VEC / 2;
  constexpr int NC = CH > 0 ? CH : 1;
  const uint32_t smem0 = smem_u32(smem_raw);
  tl_mark(p, 0);
  ring_setup(p, smem0);
  tl_mark(p, 1);
  Group<TPR> g(p, smem0);
  Ring<T, TPR, true> ring(p, smem0, g);
  ring.prologue();
  const int nchunks = p.nchunks;

  const size_t out_row_bytes = (size_t)p.out_stride * sizeof(T);
  // (A single-exchange form — online (max, sum) pairs, Group::softmax_allreduce — was built and measured slower on
  // every cluster plan: bf16 [4096,131072] 755 vs 704 us, fp32 [2048,131072] 1138 vs 708 us; the cross-entropy forward,
  // which writes nothing per element, is where that reduction pays: kernels_reduce.cuh.)
  for (; ring.acquire(); ring.advance()) {
    if (ring.k == 0) tl_mark(p, 2);
    const int nrows = ring.cur_nrows;
    const uint8_t* xrow = smem_raw + ring.stage_off + 16u * (uint32_t)g.gtid;
    uint8_t* yrow = out_row_ptr<T>(p, ring.row0, g.crank, g.gtid);
    for (int r = 0; r < nrows; ++r, xrow += p.slice_bytes, yrow += out_row_bytes) {
      float2 x[NC][H];
      float m[1] = {-INFINITY};
      for_each_chunk<TPR, CH>(g.gtid, nchunks, [&](int j, int coff) {
        float2 t[H];
        IO<T>::unpack2(ld_smem_u4(xrow + 16 * coff), t);
#pragma unroll
        for (int k = 0; k < H; ++k) {
          m[0] = MaxOp::apply(m[0], MaxOp::apply(t[k].x, t[k].y));
          if constexpr (CH > 0) x[j][k] = t[k];
        }
      });
      g.template allreduce<MaxOp, 1>(m);
      if (CH > 0 && r == nrows - 1) ring.refill();  // registers hold the row: the stage is free
      const float2 nmx = f2(-m[0]);
      float2 acc = f2(0.f);
      for_each_chunk<TPR, CH>(g.gtid, nchunks, [&](int j, int coff) {
        float2 t[H];
        if constexpr (CH == 0) IO<T>::unpack2(ld_smem_u4(xrow + 16 * coff), t);
#pragma unroll
        for (int k = 0; k < H; ++k) {
          const float2 d = f2add(CH > 0 ? x[j][k] : t[k], nmx);
          const float2 e = make_float2(exp_for<T>(d.x), exp_for<T>(d.y));
          acc = f2add(acc, e);
          if constexpr (CH > 0) x[j][k] = LOG ? d : e;
        }
      });
      float s[1] = {acc.x + acc.y};
      g.template allreduce<SumOp, 1>(s);
      const float2 fin = f2(LOG ? -logf(s[0]) : rcp_fast(s[0]));
      for_each_chunk<TPR, CH>(g.gtid, nchunks, [&](int j, int coff) {
        float2 y[H];
        if constexpr (CH > 0) {
#pragma unroll
          for (int k = 0; k < H; ++k) y[k] = LOG ? f2add(x[j][k], fin) : f2mul(x[j][k], fin);
        } else {
          float2 t[H];
          IO<T>::unpack2(ld_smem_u4(xrow + 16 * coff), t);
#pragma unroll
          for (int k = 0; k < H; ++k) {
            const float2 d = f2add(t[k], nmx);
            y[k] = LOG ? f2add(d, fin) : f2mul(make_float2(exp_for<T>(d.x), exp_for<T>(d.y)), fin);
          }
        }
        st_global_stream(yrow + 16 * coff, IO<T>::pack2(y));
      });
    }
    if constexpr (CH == 0) {
      g.sync();
      ring.refill();
    }
  }
  tl_mark(p, 3);
  ring.finish();
  ring_teardown(p);
  tl_mark(p, 6);
}

// ---------------------------------------------------------------------------
// softmax backward: dx = y*dy - y*sum(y*dy)      (in0 = dy, in1 = y)
// log-softmax backward: dx = dy - exp(y)*sum(dy)
// t = y*dy (or dy) and y (or exp(y)) stay in registers.
// ---------------------------------------------------------------------------
template <typename T, int TPR, int CH, bool LOG>
__global__ void __launch_bounds__(kMaxBlock, 1) softmax_bwd_ring_kernel(const RowParams p) {
  constexpr int VEC = IO<T>::VEC;
  constexpr int H = VEC / 2;
  constexpr int NC = CH > 0 ? CH : 1;
  const uint32_t smem0 = smem_u32(smem_raw);
  tl_mark(p, 0);
  ring_setup(p, smem0);
  tl_mark(p, 1);
  Group<TPR> g(p, smem0);
  Ring<T, TPR, true> ring(p, smem0, g);
  ring.prologue();
  const int nchunks = p.nchunks;
  const bool round_sum = (p.flags & kFlagRoundRowsum) != 0;

  const size_t out_row_bytes = (size_t)p.out_stride * sizeof(T);
  for (; ring.acquire(); ring.advance()) {
    if (ring.k == 0) tl_mark(p, 2);
    const int nrows = ring.cur_nrows;
    const uint8_t* dyrow = smem_raw + ring.stage_off + 16u * (uint32_t)g.gtid;
    uint8_t* dxrow = out_row_ptr<T>(p, ring.row0, g.crank, g.gtid);
    for (int r = 0; r < nrows; ++r, dyrow += p.slice_bytes, dxrow += out_row_bytes) {
      const uint8_t* yrow = dyrow + p.stream_bytes;
      float2 a[NC][H];  // y*dy   | dy        (log)
      float2 b[NC][H];  // y      | exp(y)    (log)
      float2 acc = f2(0.f);
      auto load = [&](int coff, float2(&av)[H], float2(&bv)[H]) {
        float2 dy[H], y[H];
        IO<T>::unpack2(ld_smem_u4(dyrow + 16 * coff), dy);
        IO<T>::unpack2(ld_smem_u4(yrow + 16 * coff), y);
#pragma unroll
        for (int k = 0; k < H; ++k) {
          if constexpr (LOG) {
            av[k] = dy[k];
            bv[k] = make_float2(expf(y[k].x), expf(y[k].y));
          } else {
            av[k] = f2mul(y[k], dy[k]);
            bv[k] = y[k];
          }
        }
      };
      for_each_chunk<TPR, CH>(g.gtid, nchunks, [&](int j, int coff) {
        float2 av[H], bv[H];
        load(coff, av, bv);
#pragma unroll
        for (int k = 0; k < H; ++k) {
          acc = f2add(acc, av[k]);
          if constexpr (CH > 0) {
            a[j][k] = av[k];
            b[j][k] = bv[k];
          }
        }
      });
      float s[1] = {acc.x + acc.y};
      g.template allreduce<SumOp, 1>(s);
      if (CH > 0 && r == nrows - 1) ring.refill();  // registers hold the row: the stage is free
      float sum = s[0];
      if (round_sum) sum = IO<T>::round_trip(sum);
      const float2 nsum = f2(-sum);
      for_each_chunk<TPR, CH>(g.gtid, nchunks, [&](int j, int coff) {
        float2 dx[H];
        if constexpr (CH > 0) {
#pragma unroll
          for (int k = 0; k < H; ++k) dx[k] = f2fma(b[j][k], nsum, a[j][k]);  // a - b*sum
        } else {
          float2 av[H], bv[H];
          load(coff, av, bv);
#pragma unroll
          for (int k = 0; k < H; ++k) dx[k] = f2fma(bv[k], nsum, av[k]);
        }
        st_global_stream(dxrow + 16 * coff, IO<T>::pack2(dx));
      });
    }
    if constexpr (CH == 0) {
      g.sync();
      ring.refill();
    }
  }
  tl_mark(p, 3);
  ring.finish();
  ring_teardown(p);
  tl_mark(p, 6);
}


// ---------------------------------------------------------------------------
// RMSNorm forward, pipelined (rows split over a thread-block cluster, e.g. C5 N = 65536 over 4
// CTAs): the register-resident kernel above waits for the DSMEM exchange of every row between its
// two passes with nothing else to do (one group per CTA). Here the row stays in the shared-memory
// stage and is read twice, which frees the registers to run the NEXT row's sum of squares while
// the exchange of the current row is in flight — the same split-phase pipeline as
// norm_bwd_pipe_kernel (kernels_bwd.cuh), with one reduced value per row.
// ---------------------------------------------------------------------------
template <typename T, int TPR, int CH>
__global__ void __launch_bounds__(kMaxBlock, 1) rmsnorm_fwd_pipe_kernel(const RowParams p) {
  constexpr int VEC = IO<T>::VEC;
  constexpr int H = VEC / 2;
  constexpr int WPG = TPR / 32;
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

  const int nchunks = p.nchunks;
  const bool stat_rms = (p.flags & kFlagStatIsRms) != 0;
  const size_t col0_bytes = (size_t)g.crank * p.slice_bytes;
  uint4 wq[CH];
  load_params_to_regs<T, TPR, CH>(p.weight, col0_bytes, g.gtid, nchunks, 1.0f, wq);
  const bool stat_writer = g.gtid == 0 && g.crank == 0 && p.stat1 != nullptr;

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
  auto pass1_post = [&](int j_row) {
    mbar_wait(c1_bar, c1_phase);
    const uint8_t* xrow = smem_raw + c1_off + 16u * (uint32_t)g.gtid;
    float2 acc = f2(0.f);
    for_each_chunk<TPR, CH>(g.gtid, nchunks, [&](int, int coff) {
      float2 t[H];
      IO<T>::unpack2(ld_smem_u4(xrow + 16 * coff), t);
#pragma unroll
      for (int k = 0; k < H; ++k) acc = f2fma(t[k], t[k], acc);
    });
    float v[1] = {acc.x + acc.y};
    warp_allreduce<SumOp, 1>(v);
    if (g.lane == 0) {
      red_slots[(j_row & 1) * (kRedSetBytes / 4) + g.warp] = v[0];
      mbar_arrive(red_bar + 8u * (uint32_t)(j_row & 1));
    }
    advance(c1_off, c1_bar, c1_s, c1_phase);
  };

  if (n > 0) pass1_post(0);
  for (int k = 0; k < n; ++k) {
    if (k == 0) tl_mark(p, 2);
    const uint32_t par = (uint32_t)(k & 1), ph = (uint32_t)((k >> 1) & 1);
    const uint32_t cl_bar = par ? cl_bar_odd : cl_bar_even;
    mbar_wait(red_bar + 8u * par, ph);
    float tot = 0.f;
    {
      const float* set = red_slots + par * (kRedSetBytes / 4);
#pragma unroll
      for (int w4 = 0; w4 < (WPG + 3) / 4; ++w4) {
        const float4 t = *reinterpret_cast<const float4*>(set + 4 * w4);
        if constexpr (WPG >= 4) {
          tot += (t.x + t.y) + (t.z + t.w);
        } else {
          tot += t.x + t.y;
        }
      }
    }
    if (g.csize > 1) {
      if (g.gtid == 0) mbar_arrive_expect_tx(cl_bar, g.csize * 4);
      if ((uint32_t)g.gtid < g.csize) {
        const uint32_t r = (uint32_t)g.gtid;
        st_async_f32(map_to_cta(smem0 + cls_off + par * kClsSetBytes + (g.crank * kRedNV) * 4, r), tot,
                     map_to_cta(cl_bar, r));
      }
    }
    if (k + 1 < n) pass1_post(k + 1);
    if (g.csize > 1) {
      mbar_wait(cl_bar, ph);
      const volatile float* slots = reinterpret_cast<const volatile float*>(smem_raw + cls_off + par * kClsSetBytes);
      tot = 0.f;
      for (uint32_t r = 0; r < g.csize; ++r) tot += slots[r * kRedNV];
    }
    const float var_eps = tot * p.inv_n + p.eps;
    const float invstd = rsqrt_fast(var_eps);
    const float rms = var_eps * invstd;
    const float2 scale = f2(invstd);
    const int64_t row = row_first + (int64_t)k * row_step;
    const uint8_t* xrow = smem_raw + c2_off + 16u * (uint32_t)g.gtid;
    uint8_t* yrow = out_row_ptr<T>(p, row, g.crank, g.gtid);
    for_each_chunk<TPR, CH>(g.gtid, nchunks, [&](int j, int coff) {
      float2 y[H], w[H];
      IO<T>::unpack2(ld_smem_u4(xrow + 16 * coff), y);
      IO<T>::unpack2(wq[j], w);
#pragma unroll
      for (int q = 0; q < H; ++q) y[q] = f2mul(f2mul(y[q], scale), w[q]);
      st_global_stream(yrow + 16 * coff, IO<T>::pack2(y));
    });
    if (stat_writer) p.stat1[row] = stat_rms ? rms : invstd;
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

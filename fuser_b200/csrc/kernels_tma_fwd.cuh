// kernels_tma_fwd.cuh — ring-pipelined kernels without per-column state:
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
// Each thread owns chunks c = gtid + j*TPR (j < CH) of the CTA's column slice. When
// the slice fits (nchunks <= TPR*CH) the row is read from the shared-memory ring
// ONCE into registers and every later pass runs on registers ("cached", the
// register-persistent buffer of the reference's inner-persistent schedule). Wider
// slices re-read the stage from shared memory in each pass ("reread").
#pragma once

#include "row_pipeline.cuh"

namespace b200norm {

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

constexpr int kMaxBlock = 512;  // groups * TPR; 128 registers per thread at one CTA per SM

// Kernel entry shared by all ring kernels: barrier setup, ring prologue, teardown.
// consumer(g, ring) walks the group's tiles: ring.wait(k) ... ring.refill(k).
template <typename T, int TPR, typename Consumer>
__device__ __forceinline__ void run_ring_kernel(const RowParams& p, Consumer&& consumer) {
  const uint32_t smem0 = smem_u32(smem_raw);
  const TileWalk tw(p);
  pipeline_setup<T>(p, smem0, tw);
  Group<TPR> g(p, smem0, tw);
  const Ring<T, TPR> ring(p, smem0, tw, g);
  ring.prologue();
  consumer(g, ring);
  if (tw.csize > 1) {
    __syncwarp();
    cluster_sync_all();  // nobody exits while a peer may still st.async into its smem
  }
}

template <typename T>
__device__ __forceinline__ uint8_t* out_row_ptr(const RowParams& p, int64_t row, uint32_t crank) {
  return static_cast<uint8_t*>(p.out0) + ((size_t)row * p.out_stride + (size_t)crank * p.cols_cta) * sizeof(T);
}

// ---------------------------------------------------------------------------
// RMSNorm forward: stat = rsqrt(mean(x^2)+eps) (or sqrt(...) with STAT_IS_RMS),
// y = x * invstd * w.
// ---------------------------------------------------------------------------
template <typename T, int TPR, int CH>
__global__ void __launch_bounds__(kMaxBlock, 1) rmsnorm_fwd_tma_kernel(const RowParams p) {
  constexpr int VEC = IO<T>::VEC;
  constexpr int H = VEC / 2;
  run_ring_kernel<T, TPR>(p, [&](Group<TPR>& g, const Ring<T, TPR>& ring) {
    const int nchunks = p.cols_cta / VEC;
    const bool cached = nchunks <= TPR * CH;
    const bool has_w = p.weight != nullptr;
    const bool stat_rms = (p.flags & kFlagStatIsRms) != 0;
    for (int64_t tk = 0; tk < ring.ntiles_g; ++tk) {
      ring.wait(tk);
      const uint32_t stage = ring.stage_off(tk);
      const int64_t row0 = ring.row0_of(tk);
      const int nrows = ring.nrows_of(row0);
      for (int r = 0; r < nrows; ++r) {
        const bool last_row = r == nrows - 1;
        const uint32_t xrow = stage + (uint32_t)r * p.slice_bytes;
        const int64_t row = row0 + r;
        uint8_t* yrow = out_row_ptr<T>(p, row, g.crank);
        float2 x[CH][H];
        float2 acc = f2(0.f);
        if (cached) {
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            const int c = g.gtid + j * TPR;
            if (c < nchunks) {
              IO<T>::unpack2(ld_smem_u4(xrow + 16u * c), x[j]);
#pragma unroll
              for (int k = 0; k < H; ++k) acc = f2fma(x[j][k], x[j][k], acc);
            }
          }
        } else {
          for (int c = g.gtid; c < nchunks; c += TPR) {
            float2 t[H];
            IO<T>::unpack2(ld_smem_u4(xrow + 16u * c), t);
#pragma unroll
            for (int k = 0; k < H; ++k) acc = f2fma(t[k], t[k], acc);
          }
        }
        float ss[1] = {acc.x + acc.y};
        g.template allreduce<SumOp, 1>(ss);
        if (cached && last_row) ring.refill(tk);  // registers hold the row: the stage is free
        const float var_eps = ss[0] * p.inv_n + p.eps;
        const float invstd = __frsqrt_rn(var_eps);
        const float rms = __fsqrt_rn(var_eps);
        const float2 scale = f2(stat_rms ? __frcp_rn(rms) : invstd);
        auto emit = [&](int c, const float2(&xv)[H]) {
          float2 y[H];
#pragma unroll
          for (int k = 0; k < H; ++k) y[k] = f2mul(xv[k], scale);
          if (has_w) {
            float2 w[H];
            load_param<VEC>(p.off_weight, nchunks, c, w);
#pragma unroll
            for (int k = 0; k < H; ++k) y[k] = f2mul(y[k], w[k]);
          }
          st_global_stream(yrow + 16u * c, IO<T>::pack2(y));
        };
        if (cached) {
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            const int c = g.gtid + j * TPR;
            if (c < nchunks) emit(c, x[j]);
          }
        } else {
          for (int c = g.gtid; c < nchunks; c += TPR) {
            float2 t[H];
            IO<T>::unpack2(ld_smem_u4(xrow + 16u * c), t);
            emit(c, t);
          }
        }
        if (g.gtid == 0 && g.crank == 0 && p.stat1 != nullptr) p.stat1[row] = stat_rms ? rms : invstd;
      }
      if (!cached) {  // re-read variant: the stage was read until the last pass
        g.sync();
        ring.refill(tk);
      }
    }
  });
}

// ---------------------------------------------------------------------------
// LayerNorm forward: mean, invstd = rsqrt(M2/N + eps), y = (x-mean)*invstd*w + b.
// ---------------------------------------------------------------------------
template <typename T, int TPR, int CH>
__global__ void __launch_bounds__(kMaxBlock, 1) layernorm_fwd_tma_kernel(const RowParams p) {
  constexpr int VEC = IO<T>::VEC;
  constexpr int H = VEC / 2;
  run_ring_kernel<T, TPR>(p, [&](Group<TPR>& g, const Ring<T, TPR>& ring) {
    const int nchunks = p.cols_cta / VEC;
    const bool cached = nchunks <= TPR * CH;
    const bool has_w = p.weight != nullptr;
    const bool has_b = p.bias != nullptr;
    for (int64_t tk = 0; tk < ring.ntiles_g; ++tk) {
      ring.wait(tk);
      const uint32_t stage = ring.stage_off(tk);
      const int64_t row0 = ring.row0_of(tk);
      const int nrows = ring.nrows_of(row0);
      for (int r = 0; r < nrows; ++r) {
        const bool last_row = r == nrows - 1;
        const uint32_t xrow = stage + (uint32_t)r * p.slice_bytes;
        const int64_t row = row0 + r;
        uint8_t* yrow = out_row_ptr<T>(p, row, g.crank);
        float2 x[CH][H];
        float2 acc = f2(0.f);
        if (cached) {
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            const int c = g.gtid + j * TPR;
            if (c < nchunks) {
              IO<T>::unpack2(ld_smem_u4(xrow + 16u * c), x[j]);
#pragma unroll
              for (int k = 0; k < H; ++k) acc = f2add(acc, x[j][k]);
            }
          }
        } else {
          for (int c = g.gtid; c < nchunks; c += TPR) {
            float2 t[H];
            IO<T>::unpack2(ld_smem_u4(xrow + 16u * c), t);
#pragma unroll
            for (int k = 0; k < H; ++k) acc = f2add(acc, t[k]);
          }
        }
        float s[1] = {acc.x + acc.y};
        g.template allreduce<SumOp, 1>(s);
        if (cached && last_row) ring.refill(tk);  // registers hold the row: the stage is free
        const float mean = s[0] * p.inv_n;
        const float2 nmean = f2(-mean);
        acc = f2(0.f);
        if (cached) {
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            const int c = g.gtid + j * TPR;
            if (c < nchunks) {
#pragma unroll
              for (int k = 0; k < H; ++k) {
                x[j][k] = f2add(x[j][k], nmean);  // x now holds x - mean
                acc = f2fma(x[j][k], x[j][k], acc);
              }
            }
          }
        } else {
          for (int c = g.gtid; c < nchunks; c += TPR) {
            float2 t[H];
            IO<T>::unpack2(ld_smem_u4(xrow + 16u * c), t);
#pragma unroll
            for (int k = 0; k < H; ++k) {
              const float2 d = f2add(t[k], nmean);
              acc = f2fma(d, d, acc);
            }
          }
        }
        float m2[1] = {acc.x + acc.y};
        g.template allreduce<SumOp, 1>(m2);
        const float invstd = __frsqrt_rn(m2[0] * p.inv_n + p.eps);
        const float2 is2 = f2(invstd);
        auto emit = [&](int c, const float2(&d)[H]) {
          float2 y[H];
#pragma unroll
          for (int k = 0; k < H; ++k) y[k] = f2mul(d[k], is2);
          if (has_w) {
            float2 w[H];
            load_param<VEC>(p.off_weight, nchunks, c, w);
#pragma unroll
            for (int k = 0; k < H; ++k) y[k] = f2mul(y[k], w[k]);
          }
          if (has_b) {
            float2 b[H];
            load_param<VEC>(p.off_bias, nchunks, c, b);
#pragma unroll
            for (int k = 0; k < H; ++k) y[k] = f2add(y[k], b[k]);
          }
          st_global_stream(yrow + 16u * c, IO<T>::pack2(y));
        };
        if (cached) {
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            const int c = g.gtid + j * TPR;
            if (c < nchunks) emit(c, x[j]);
          }
        } else {
          for (int c = g.gtid; c < nchunks; c += TPR) {
            float2 t[H];
            IO<T>::unpack2(ld_smem_u4(xrow + 16u * c), t);
#pragma unroll
            for (int k = 0; k < H; ++k) t[k] = f2add(t[k], nmean);
            emit(c, t);
          }
        }
        if (g.gtid == 0 && g.crank == 0) {
          if (p.stat0 != nullptr) p.stat0[row] = mean;
          if (p.stat1 != nullptr) p.stat1[row] = invstd;
        }
      }
      if (!cached) {  // re-read variant: the stage was read until the last pass
        g.sync();
        ring.refill(tk);
      }
    }
  });
}

// ---------------------------------------------------------------------------
// softmax / log-softmax forward. Cached: x, then exp(x-max) (or x-max for the log
// flavour), live in registers — one MUFU per element. Reread recomputes exp.
// NaN semantics: max is NaN-propagating (at::amax); any NaN / +inf / all -inf row
// turns into an all-NaN row exactly as the op-by-op ATen evaluation does.
// ---------------------------------------------------------------------------
template <typename T, int TPR, int CH, bool LOG>
__global__ void __launch_bounds__(kMaxBlock, 1) softmax_fwd_tma_kernel(const RowParams p) {
  constexpr int VEC = IO<T>::VEC;
  constexpr int H = VEC / 2;
  run_ring_kernel<T, TPR>(p, [&](Group<TPR>& g, const Ring<T, TPR>& ring) {
    const int nchunks = p.cols_cta / VEC;
    const bool cached = nchunks <= TPR * CH;
    for (int64_t tk = 0; tk < ring.ntiles_g; ++tk) {
      ring.wait(tk);
      const uint32_t stage = ring.stage_off(tk);
      const int64_t row0 = ring.row0_of(tk);
      const int nrows = ring.nrows_of(row0);
      for (int r = 0; r < nrows; ++r) {
        const bool last_row = r == nrows - 1;
        const uint32_t xrow = stage + (uint32_t)r * p.slice_bytes;
        const int64_t row = row0 + r;
        uint8_t* yrow = out_row_ptr<T>(p, row, g.crank);
        float2 x[CH][H];
        float m[1] = {-INFINITY};
        if (cached) {
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            const int c = g.gtid + j * TPR;
            if (c < nchunks) {
              IO<T>::unpack2(ld_smem_u4(xrow + 16u * c), x[j]);
#pragma unroll
              for (int k = 0; k < H; ++k) m[0] = MaxOp::apply(m[0], MaxOp::apply(x[j][k].x, x[j][k].y));
            }
          }
        } else {
          for (int c = g.gtid; c < nchunks; c += TPR) {
            float2 t[H];
            IO<T>::unpack2(ld_smem_u4(xrow + 16u * c), t);
#pragma unroll
            for (int k = 0; k < H; ++k) m[0] = MaxOp::apply(m[0], MaxOp::apply(t[k].x, t[k].y));
          }
        }
        g.template allreduce<MaxOp, 1>(m);
        if (cached && last_row) ring.refill(tk);  // registers hold the row: the stage is free
        const float2 nmx = f2(-m[0]);
        float2 acc = f2(0.f);
        if (cached) {
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            const int c = g.gtid + j * TPR;
            if (c < nchunks) {
#pragma unroll
              for (int k = 0; k < H; ++k) {
                const float2 d = f2add(x[j][k], nmx);
                const float2 e = make_float2(exp_for<T>(d.x), exp_for<T>(d.y));
                acc = f2add(acc, e);
                x[j][k] = LOG ? d : e;
              }
            }
          }
        } else {
          for (int c = g.gtid; c < nchunks; c += TPR) {
            float2 t[H];
            IO<T>::unpack2(ld_smem_u4(xrow + 16u * c), t);
#pragma unroll
            for (int k = 0; k < H; ++k) {
              const float2 d = f2add(t[k], nmx);
              acc = f2add(acc, make_float2(exp_for<T>(d.x), exp_for<T>(d.y)));
            }
          }
        }
        float s[1] = {acc.x + acc.y};
        g.template allreduce<SumOp, 1>(s);
        const float2 fin = f2(LOG ? -logf(s[0]) : __frcp_rn(s[0]));
        if (cached) {
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            const int c = g.gtid + j * TPR;
            if (c < nchunks) {
              float2 y[H];
#pragma unroll
              for (int k = 0; k < H; ++k) y[k] = LOG ? f2add(x[j][k], fin) : f2mul(x[j][k], fin);
              st_global_stream(yrow + 16u * c, IO<T>::pack2(y));
            }
          }
        } else {
          for (int c = g.gtid; c < nchunks; c += TPR) {
            float2 t[H], y[H];
            IO<T>::unpack2(ld_smem_u4(xrow + 16u * c), t);
#pragma unroll
            for (int k = 0; k < H; ++k) {
              const float2 d = f2add(t[k], nmx);
              y[k] = LOG ? f2add(d, fin) : f2mul(make_float2(exp_for<T>(d.x), exp_for<T>(d.y)), fin);
            }
            st_global_stream(yrow + 16u * c, IO<T>::pack2(y));
          }
        }
      }
      if (!cached) {  // re-read variant: the stage was read until the last pass
        g.sync();
        ring.refill(tk);
      }
    }
  });
}

// ---------------------------------------------------------------------------
// softmax backward: dx = y*dy - y*sum(y*dy)      (in0 = dy, in1 = y)
// log-softmax backward: dx = dy - exp(y)*sum(dy)
// Cached: t = y*dy (or dy) and y (or exp(y)) stay in registers.
// ---------------------------------------------------------------------------
template <typename T, int TPR, int CH, bool LOG>
__global__ void __launch_bounds__(kMaxBlock, 1) softmax_bwd_tma_kernel(const RowParams p) {
  constexpr int VEC = IO<T>::VEC;
  constexpr int H = VEC / 2;
  run_ring_kernel<T, TPR>(p, [&](Group<TPR>& g, const Ring<T, TPR>& ring) {
    const int nchunks = p.cols_cta / VEC;
    const bool cached = nchunks <= TPR * CH;
    for (int64_t tk = 0; tk < ring.ntiles_g; ++tk) {
      ring.wait(tk);
      const uint32_t stage = ring.stage_off(tk);
      const int64_t row0 = ring.row0_of(tk);
      const int nrows = ring.nrows_of(row0);
      for (int r = 0; r < nrows; ++r) {
        const bool last_row = r == nrows - 1;
        const uint32_t dyrow = stage + (uint32_t)r * p.slice_bytes;
        const uint32_t yrow_s = dyrow + p.stream_bytes;
        const int64_t row = row0 + r;
        uint8_t* dxrow = out_row_ptr<T>(p, row, g.crank);
        float2 a[CH][H];  // y*dy   | dy        (log)
        float2 b[CH][H];  // y      | exp(y)    (log)
        float2 acc = f2(0.f);
        auto load = [&](int c, float2(&av)[H], float2(&bv)[H]) {
          float2 dy[H], y[H];
          IO<T>::unpack2(ld_smem_u4(dyrow + 16u * c), dy);
          IO<T>::unpack2(ld_smem_u4(yrow_s + 16u * c), y);
#pragma unroll
          for (int k = 0; k < H; ++k) {
            if constexpr (LOG) {
              av[k] = dy[k];
              bv[k] = make_float2(expf(y[k].x), expf(y[k].y));
            } else {
              av[k] = f2mul(y[k], dy[k]);
              bv[k] = y[k];
            }
            acc = f2add(acc, av[k]);
          }
        };
        if (cached) {
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            const int c = g.gtid + j * TPR;
            if (c < nchunks) load(c, a[j], b[j]);
          }
        } else {
          for (int c = g.gtid; c < nchunks; c += TPR) {
            float2 av[H], bv[H];
            load(c, av, bv);
          }
        }
        float s[1] = {acc.x + acc.y};
        g.template allreduce<SumOp, 1>(s);
        if (cached && last_row) ring.refill(tk);  // registers hold the row: the stage is free
        float sum = s[0];
        if (p.flags & kFlagRoundRowsum) sum = IO<T>::round_trip(sum);
        const float2 nsum = f2(-sum);
        auto emit = [&](int c, const float2(&av)[H], const float2(&bv)[H]) {
          float2 dx[H];
#pragma unroll
          for (int k = 0; k < H; ++k) dx[k] = f2fma(bv[k], nsum, av[k]);  // a - b*sum
          st_global_stream(dxrow + 16u * c, IO<T>::pack2(dx));
        };
        if (cached) {
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            const int c = g.gtid + j * TPR;
            if (c < nchunks) emit(c, a[j], b[j]);
          }
        } else {
          acc = f2(0.f);
          for (int c = g.gtid; c < nchunks; c += TPR) {
            float2 av[H], bv[H];
            load(c, av, bv);
            emit(c, av, bv);
          }
        }
      }
      if (!cached) {  // re-read variant: the stage was read until the last pass
        g.sync();
        ring.refill(tk);
      }
    }
  });
}

}  // namespace b200norm

// kernels_tma_fwd.cuh — ring-pipelined kernels without per-column state:
//   RMSNorm fwd, LayerNorm fwd, softmax fwd/bwd, log-softmax fwd/bwd.
//
// Math follows the reference op builders (csrc/ops/normalization.cpp):
//   rms_norm :352-380, layer_norm :304-342 (Welford -> two-pass: mean, then
//   sum((x-mean)^2), which is what the segmenter's TranslateApplicableWelford does
//   for persistent kernels, fusion_segmenter.cpp:2341), softmax :169-189,
//   softmax_backward :191-210, log_softmax :212-231, log_softmax_backward :233-249.
// All arithmetic is fp32; the only rounding to the io dtype is at the store.
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

// Kernel entry shared by all ring kernels: barrier setup, role split, teardown.
template <typename T, int TPR, typename Consumer>
__device__ __forceinline__ void run_ring_kernel(const RowParams& p, Consumer&& consumer) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const uint32_t smem0 = smem_u32(smem_raw);
  const TileWalk tw(p);
  pipeline_setup<T>(p, smem0, TPR / 32, tw);
  const int n_consumer_threads = p.groups * TPR;
  if ((int)threadIdx.x >= n_consumer_threads) {
    producer_loop<T>(p, smem0, tw);
  } else {
    Group<TPR> g(p, smem0, tw);
    consumer(g, smem0, tw);
  }
  if (tw.csize > 1) {
    __syncwarp();
    cluster_sync_all();  // nobody exits while a peer may still st.async into its smem
  }
}

// ---------------------------------------------------------------------------
// RMSNorm forward: stat = rsqrt(mean(x^2)+eps) (or sqrt(...) with STAT_IS_RMS),
// y = x * invstd * w.
// ---------------------------------------------------------------------------
template <typename T, int TPR>
__global__ void __launch_bounds__(544, 1) rmsnorm_fwd_tma_kernel(const RowParams p) {
  constexpr int VEC = IO<T>::VEC;
  run_ring_kernel<T, TPR>(p, [&](Group<TPR>& g, uint32_t smem0, const TileWalk& tw) {
    const int nchunks = p.cols_cta / VEC;
    const uint32_t wsm = smem0 + p.off_weight;
    const bool has_w = p.weight != nullptr;
    consume_tiles<TPR>(p, smem0, tw, g, [&](int64_t row0, int nrows, uint32_t stage) {
      for (int r = 0; r < nrows; ++r) {
        const uint32_t xrow = stage + (uint32_t)r * p.slice_bytes;
        float ss[1] = {0.f};
        for (int c = g.gtid; c < nchunks; c += TPR) {
          float x[VEC];
          IO<T>::unpack(lds_128(xrow + 16u * c), x);
#pragma unroll
          for (int k = 0; k < VEC; ++k) ss[0] = fmaf(x[k], x[k], ss[0]);
        }
        g.template allreduce<SumOp, 1>(ss);
        const float var_eps = ss[0] * p.inv_n + p.eps;
        const float invstd = __frsqrt_rn(var_eps);
        const float rms = __fsqrt_rn(var_eps);
        const float scale = (p.flags & kFlagStatIsRms) ? __frcp_rn(rms) : invstd;
        const int64_t row = row0 + r;
        uint8_t* yrow = static_cast<uint8_t*>(p.out0) +
                        ((size_t)row * p.out_stride + (size_t)g.crank * p.cols_cta) * sizeof(T);
        for (int c = g.gtid; c < nchunks; c += TPR) {
          float x[VEC], y[VEC];
          IO<T>::unpack(lds_128(xrow + 16u * c), x);
          if (has_w) {
            float w[VEC];
            IO<T>::unpack(lds_128(wsm + 16u * c), w);
#pragma unroll
            for (int k = 0; k < VEC; ++k) y[k] = (x[k] * scale) * w[k];
          } else {
#pragma unroll
            for (int k = 0; k < VEC; ++k) y[k] = x[k] * scale;
          }
          stg_128_stream(yrow + 16u * c, IO<T>::pack(y));
        }
        if (g.gtid == 0 && g.crank == 0 && p.stat1 != nullptr) {
          p.stat1[row] = (p.flags & kFlagStatIsRms) ? rms : invstd;
        }
      }
    });
  });
}

// ---------------------------------------------------------------------------
// LayerNorm forward: mean, invstd = rsqrt(M2/N + eps), y = (x-mean)*invstd*w + b.
// ---------------------------------------------------------------------------
template <typename T, int TPR>
__global__ void __launch_bounds__(544, 1) layernorm_fwd_tma_kernel(const RowParams p) {
  constexpr int VEC = IO<T>::VEC;
  run_ring_kernel<T, TPR>(p, [&](Group<TPR>& g, uint32_t smem0, const TileWalk& tw) {
    const int nchunks = p.cols_cta / VEC;
    const uint32_t wsm = smem0 + p.off_weight;
    const uint32_t bsm = smem0 + p.off_bias;
    const bool has_w = p.weight != nullptr;
    const bool has_b = p.bias != nullptr;
    consume_tiles<TPR>(p, smem0, tw, g, [&](int64_t row0, int nrows, uint32_t stage) {
      for (int r = 0; r < nrows; ++r) {
        const uint32_t xrow = stage + (uint32_t)r * p.slice_bytes;
        float s[1] = {0.f};
        for (int c = g.gtid; c < nchunks; c += TPR) {
          float x[VEC];
          IO<T>::unpack(lds_128(xrow + 16u * c), x);
#pragma unroll
          for (int k = 0; k < VEC; ++k) s[0] += x[k];
        }
        g.template allreduce<SumOp, 1>(s);
        const float mean = s[0] * p.inv_n;
        float m2[1] = {0.f};
        for (int c = g.gtid; c < nchunks; c += TPR) {
          float x[VEC];
          IO<T>::unpack(lds_128(xrow + 16u * c), x);
#pragma unroll
          for (int k = 0; k < VEC; ++k) {
            const float d = x[k] - mean;
            m2[0] = fmaf(d, d, m2[0]);
          }
        }
        g.template allreduce<SumOp, 1>(m2);
        const float invstd = __frsqrt_rn(m2[0] * p.inv_n + p.eps);
        const int64_t row = row0 + r;
        uint8_t* yrow = static_cast<uint8_t*>(p.out0) +
                        ((size_t)row * p.out_stride + (size_t)g.crank * p.cols_cta) * sizeof(T);
        for (int c = g.gtid; c < nchunks; c += TPR) {
          float x[VEC], y[VEC];
          IO<T>::unpack(lds_128(xrow + 16u * c), x);
#pragma unroll
          for (int k = 0; k < VEC; ++k) y[k] = (x[k] - mean) * invstd;
          if (has_w) {
            float w[VEC];
            IO<T>::unpack(lds_128(wsm + 16u * c), w);
#pragma unroll
            for (int k = 0; k < VEC; ++k) y[k] *= w[k];
          }
          if (has_b) {
            float b[VEC];
            IO<T>::unpack(lds_128(bsm + 16u * c), b);
#pragma unroll
            for (int k = 0; k < VEC; ++k) y[k] += b[k];
          }
          stg_128_stream(yrow + 16u * c, IO<T>::pack(y));
        }
        if (g.gtid == 0 && g.crank == 0) {
          if (p.stat0 != nullptr) p.stat0[row] = mean;
          if (p.stat1 != nullptr) p.stat1[row] = invstd;
        }
      }
    });
  });
}

// ---------------------------------------------------------------------------
// softmax / log-softmax forward. CH > 0: exp(x-max) is kept in registers between
// the sum pass and the output pass (one MUFU per element) for slices of up to
// TPR*VEC*CH columns; wider slices recompute it.
// ---------------------------------------------------------------------------
template <typename T, int TPR, int CH, bool LOG>
__global__ void __launch_bounds__(544, 1) softmax_fwd_tma_kernel(const RowParams p) {
  constexpr int VEC = IO<T>::VEC;
  run_ring_kernel<T, TPR>(p, [&](Group<TPR>& g, uint32_t smem0, const TileWalk& tw) {
    const int nchunks = p.cols_cta / VEC;
    const bool cached = nchunks <= TPR * CH;
    consume_tiles<TPR>(p, smem0, tw, g, [&](int64_t row0, int nrows, uint32_t stage) {
      for (int r = 0; r < nrows; ++r) {
        const uint32_t xrow = stage + (uint32_t)r * p.slice_bytes;
        float m[1] = {-INFINITY};
        for (int c = g.gtid; c < nchunks; c += TPR) {
          float x[VEC];
          IO<T>::unpack(lds_128(xrow + 16u * c), x);
#pragma unroll
          for (int k = 0; k < VEC; ++k) m[0] = MaxOp::apply(m[0], x[k]);
        }
        g.template allreduce<MaxOp, 1>(m);
        const float mx = m[0];
        const int64_t row = row0 + r;
        uint8_t* yrow = static_cast<uint8_t*>(p.out0) +
                        ((size_t)row * p.out_stride + (size_t)g.crank * p.cols_cta) * sizeof(T);
        float e[CH > 0 ? CH * VEC : 1];
        float s[1] = {0.f};
        if (cached) {
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            const int c = g.gtid + j * TPR;
            if (c < nchunks) {
              float x[VEC];
              IO<T>::unpack(lds_128(xrow + 16u * c), x);
#pragma unroll
              for (int k = 0; k < VEC; ++k) {
                const float d = x[k] - mx;
                const float ex = exp_for<T>(d);
                s[0] += ex;
                e[j * VEC + k] = LOG ? d : ex;
              }
            }
          }
        } else {
          for (int c = g.gtid; c < nchunks; c += TPR) {
            float x[VEC];
            IO<T>::unpack(lds_128(xrow + 16u * c), x);
#pragma unroll
            for (int k = 0; k < VEC; ++k) s[0] += exp_for<T>(x[k] - mx);
          }
        }
        g.template allreduce<SumOp, 1>(s);
        const float fin = LOG ? logf(s[0]) : __frcp_rn(s[0]);
        if (cached) {
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            const int c = g.gtid + j * TPR;
            if (c < nchunks) {
              float y[VEC];
#pragma unroll
              for (int k = 0; k < VEC; ++k) y[k] = LOG ? (e[j * VEC + k] - fin) : (e[j * VEC + k] * fin);
              stg_128_stream(yrow + 16u * c, IO<T>::pack(y));
            }
          }
        } else {
          for (int c = g.gtid; c < nchunks; c += TPR) {
            float x[VEC], y[VEC];
            IO<T>::unpack(lds_128(xrow + 16u * c), x);
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
              const float d = x[k] - mx;
              y[k] = LOG ? (d - fin) : (exp_for<T>(d) * fin);
            }
            stg_128_stream(yrow + 16u * c, IO<T>::pack(y));
          }
        }
      }
    });
  });
}

// ---------------------------------------------------------------------------
// softmax backward: dx = y*dy - y*sum(y*dy)      (in0 = dy, in1 = y)
// log-softmax backward: dx = dy - exp(y)*sum(dy)
// ---------------------------------------------------------------------------
template <typename T, int TPR, bool LOG>
__global__ void __launch_bounds__(544, 1) softmax_bwd_tma_kernel(const RowParams p) {
  constexpr int VEC = IO<T>::VEC;
  run_ring_kernel<T, TPR>(p, [&](Group<TPR>& g, uint32_t smem0, const TileWalk& tw) {
    const int nchunks = p.cols_cta / VEC;
    consume_tiles<TPR>(p, smem0, tw, g, [&](int64_t row0, int nrows, uint32_t stage) {
      for (int r = 0; r < nrows; ++r) {
        const uint32_t dyrow = stage + (uint32_t)r * p.slice_bytes;
        const uint32_t yrow_s = dyrow + p.stream_bytes;
        float s[1] = {0.f};
        for (int c = g.gtid; c < nchunks; c += TPR) {
          float dy[VEC];
          IO<T>::unpack(lds_128(dyrow + 16u * c), dy);
          if constexpr (LOG) {
#pragma unroll
            for (int k = 0; k < VEC; ++k) s[0] += dy[k];
          } else {
            float y[VEC];
            IO<T>::unpack(lds_128(yrow_s + 16u * c), y);
#pragma unroll
            for (int k = 0; k < VEC; ++k) s[0] = fmaf(dy[k], y[k], s[0]);
          }
        }
        g.template allreduce<SumOp, 1>(s);
        float sum = s[0];
        if (p.flags & kFlagRoundRowsum) sum = IO<T>::round_trip(sum);
        const int64_t row = row0 + r;
        uint8_t* dxrow = static_cast<uint8_t*>(p.out0) +
                         ((size_t)row * p.out_stride + (size_t)g.crank * p.cols_cta) * sizeof(T);
        for (int c = g.gtid; c < nchunks; c += TPR) {
          float dy[VEC], y[VEC], dx[VEC];
          IO<T>::unpack(lds_128(dyrow + 16u * c), dy);
          IO<T>::unpack(lds_128(yrow_s + 16u * c), y);
#pragma unroll
          for (int k = 0; k < VEC; ++k) {
            if constexpr (LOG) {
              dx[k] = dy[k] - expf(y[k]) * sum;
            } else {
              dx[k] = y[k] * dy[k] - y[k] * sum;
            }
          }
          stg_128_stream(dxrow + 16u * c, IO<T>::pack(dx));
        }
      }
    });
  });
}

}  // namespace b200norm

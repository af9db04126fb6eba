// kernels_tma_bwd.cuh — LayerNorm / RMSNorm backward on the TMA ring, with the
// dgamma / dbeta column sums fused in (the InnerOuterPersistent segment).
//
// Reference math (csrc/ops/normalization.cpp):
//   layer_norm_backward :382-436, rms_norm_backward :438-485 (keeps the -sum(g)
//   term), thunder_rms_norm_backward :487-535 (textbook form, rms statistic).
//     x_hat = (x - mean) * invstd            g = dy * w
//     dx    = (invstd / N) * (N*g - sum(g) - x_hat * sum(g*x_hat))
//     dw    = sum_rows dy * x_hat            db = sum_rows dy
// The reference's generated kernel (scheduler/normalization_inner_outer*.cpp)
// keeps per-CTA partial column sums in registers and finishes them with a grid
// reduction; we keep per-thread fp32 column accumulators for the CTA's lifetime,
// combine the consumer groups through shared memory once, write one fp32 partial
// row per CTA, and finish in finalize_partials_kernel (kernels_generic.cuh).
#pragma once

#include "kernels_tma_fwd.cuh"

namespace b200norm {

// LN: true = LayerNorm (mean input, dbeta), false = RMSNorm.
// Each thread owns chunks c = gtid + j*TPR, j < CH, of the CTA's column slice.
template <typename T, int TPR, int CH, bool LN>
__global__ void __launch_bounds__(544, 1) norm_bwd_tma_kernel(const RowParams p) {
  constexpr int VEC = IO<T>::VEC;
  run_ring_kernel<T, TPR>(p, [&](Group<TPR>& g, uint32_t smem0, const TileWalk& tw) {
    const int nchunks = p.cols_cta / VEC;
    const uint32_t wsm = smem0 + p.off_weight;
    const bool has_w = p.weight != nullptr;
    const bool want_dx = (p.flags & kFlagOutDx) != 0;
    const bool want_dw = (p.flags & kFlagOutDw) != 0;
    const bool want_db = LN && (p.flags & kFlagOutDb) != 0;
    const bool centered = LN || (p.flags & kFlagRmsBwdCentered) != 0;

    float dw_acc[CH * VEC];
    float db_acc[LN ? CH * VEC : 1];
#pragma unroll
    for (int i = 0; i < CH * VEC; ++i) dw_acc[i] = 0.f;
#pragma unroll
    for (int i = 0; i < (LN ? CH * VEC : 1); ++i) db_acc[i] = 0.f;

    consume_tiles<TPR>(p, smem0, tw, g, [&](int64_t row0, int nrows, uint32_t stage) {
      for (int r = 0; r < nrows; ++r) {
        const int64_t row = row0 + r;
        const uint32_t dyrow = stage + (uint32_t)r * p.slice_bytes;
        const uint32_t xrow = dyrow + p.stream_bytes;
        const float mean = LN ? __ldg(p.stat0 + row) : 0.f;
        const float st = __ldg(p.stat1 + row);
        const float invstd = (!LN && (p.flags & kFlagStatIsRms)) ? __frcp_rn(st) : st;
        float s[2] = {0.f, 0.f};
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          const int c = g.gtid + j * TPR;
          if (c < nchunks) {
            float dy[VEC], x[VEC], w[VEC];
            IO<T>::unpack(lds_128(dyrow + 16u * c), dy);
            IO<T>::unpack(lds_128(xrow + 16u * c), x);
            if (has_w) {
              IO<T>::unpack(lds_128(wsm + 16u * c), w);
            } else {
#pragma unroll
              for (int k = 0; k < VEC; ++k) w[k] = 1.f;
            }
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
              const float xh = (x[k] - mean) * invstd;
              const float gg = dy[k] * w[k];
              s[0] += gg;
              s[1] = fmaf(gg, xh, s[1]);
              dw_acc[j * VEC + k] = fmaf(dy[k], xh, dw_acc[j * VEC + k]);
              if constexpr (LN) db_acc[j * VEC + k] += dy[k];
            }
          }
        }
        if (want_dx) {
          g.template allreduce<SumOp, 2>(s);
          const float s1 = centered ? s[0] : 0.f;
          const float s2 = s[1];
          const float a = invstd * p.inv_n;
          uint8_t* dxrow = static_cast<uint8_t*>(p.out0) +
                           ((size_t)row * p.out_stride + (size_t)g.crank * p.cols_cta) * sizeof(T);
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            const int c = g.gtid + j * TPR;
            if (c < nchunks) {
              float dy[VEC], x[VEC], w[VEC], dx[VEC];
              IO<T>::unpack(lds_128(dyrow + 16u * c), dy);
              IO<T>::unpack(lds_128(xrow + 16u * c), x);
              if (has_w) {
                IO<T>::unpack(lds_128(wsm + 16u * c), w);
              } else {
#pragma unroll
                for (int k = 0; k < VEC; ++k) w[k] = 1.f;
              }
#pragma unroll
              for (int k = 0; k < VEC; ++k) {
                const float xh = (x[k] - mean) * invstd;
                const float gg = dy[k] * w[k];
                dx[k] = a * ((p.n_f * gg - s1) - xh * s2);
              }
              stg_128_stream(dxrow + 16u * c, IO<T>::pack(dx));
            }
          }
        }
      }
    });

    // ---- combine the groups' column accumulators and emit this CTA's partial row ----
    if (want_dw || want_db) {
      const int nct = p.groups * TPR;
      float* acc = reinterpret_cast<float*>(__cvta_shared_to_generic(smem0 + p.off_acc));
      float* acc_w = acc;
      float* acc_b = acc + p.cols_cta;
      named_bar_sync(15, nct);  // every group is done with the ring
      for (int gi = 0; gi < p.groups; ++gi) {
        if (g.gid == gi) {
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            const int c = g.gtid + j * TPR;
            if (c < nchunks) {
#pragma unroll
              for (int k = 0; k < VEC; ++k) {
                const int col = c * VEC + k;
                if (gi == 0) {
                  acc_w[col] = dw_acc[j * VEC + k];
                  if constexpr (LN) acc_b[col] = db_acc[j * VEC + k];
                } else {
                  acc_w[col] += dw_acc[j * VEC + k];
                  if constexpr (LN) acc_b[col] += db_acc[j * VEC + k];
                }
              }
            }
          }
        }
        named_bar_sync(15, nct);
      }
      // partials layout: [part][K][N], K = 2 for LN (dw, db), 1 for RMSNorm
      constexpr int K = LN ? 2 : 1;
      const int64_t part = blockIdx.x / g.csize;
      float* dst = p.partials + (size_t)part * K * p.cols + (size_t)g.crank * p.cols_cta;
      for (int col = threadIdx.x; col < p.cols_cta; col += nct) {
        dst[col] = acc_w[col];
        if constexpr (LN) dst[p.cols + col] = acc_b[col];
      }
    }
  });
}

}  // namespace b200norm

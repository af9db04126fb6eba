// kernels_tma_bwd.cuh — LayerNorm / RMSNorm backward on the TMA ring, with the
// dgamma / dbeta column sums fused in (the InnerOuterPersistent segment).
//
// Reference math (csrc/ops/normalization.cpp):
//   layer_norm_backward :382-436, rms_norm_backward :438-485 (keeps the -sum(g)
//   term), thunder_rms_norm_backward :487-535 (textbook form, rms statistic).
//     x_hat = (x - mean) * invstd            g = dy * w
//     dx    = (invstd / N) * (N*g - sum(g) - x_hat * sum(g*x_hat))
//     dw    = sum_rows dy * x_hat            db = sum_rows dy
// Evaluated here as  dx = A*g + B*xc + C  with xc = x - mean and per-row scalars
//     A = invstd,  B = -invstd^3 * sum(g*xc) / N,  C = -invstd * sum(g) / N
// (same polynomial, three packed FMAs per two elements).
//
// The reference's generated kernel (scheduler/normalization_inner_outer*.cpp)
// keeps per-CTA partial column sums in registers and finishes them with a grid
// reduction; we keep per-thread fp32 column accumulators for the CTA's lifetime,
// combine the consumer groups through shared memory once, write one fp32 partial
// row per CTA, and finish in finalize_partials_kernel (kernels_generic.cuh).
#pragma once

#include "kernels_tma_fwd.cuh"

namespace b200norm {

// LN: true = LayerNorm (mean input, dbeta), false = RMSNorm.
// Each thread owns chunks c = gtid + j*TPR, j < CH, of the CTA's column slice
// (the planner guarantees nchunks <= TPR*CH).
// CACHE: g and xc stay in registers between the reduction pass and the dx pass;
// otherwise the stage is re-read from shared memory.
template <typename T, int TPR, int CH, bool LN, bool CACHE>
__global__ void __launch_bounds__(kMaxBlock, 1) norm_bwd_tma_kernel(const RowParams p) {
  constexpr int VEC = IO<T>::VEC;
  constexpr int H = VEC / 2;
  run_ring_kernel<T, TPR>(p, [&](Group<TPR>& g, const Ring<T, TPR>& ring) {
    const int nchunks = p.cols_cta / VEC;
    const bool has_w = p.weight != nullptr;
    const bool want_dx = (p.flags & kFlagOutDx) != 0;
    const bool want_dw = (p.flags & kFlagOutDw) != 0;
    const bool want_db = LN && (p.flags & kFlagOutDb) != 0;
    const bool centered = LN || (p.flags & kFlagRmsBwdCentered) != 0;
    const bool stat_rms = !LN && (p.flags & kFlagStatIsRms) != 0;

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

    for (int64_t tk = 0; tk < ring.ntiles_g; ++tk) {
      ring.wait(tk);
      const uint32_t stage = ring.stage_off(tk);
      const int64_t row0 = ring.row0_of(tk);
      const int nrows = ring.nrows_of(row0);
      for (int r = 0; r < nrows; ++r) {
        const bool last_row = r == nrows - 1;
        const int64_t row = row0 + r;
        const uint32_t dyrow = stage + (uint32_t)r * p.slice_bytes;
        const uint32_t xrow = dyrow + p.stream_bytes;
        const float mean = LN ? __ldg(p.stat0 + row) : 0.f;
        const float st = __ldg(p.stat1 + row);
        const float invstd = stat_rms ? __frcp_rn(st) : st;
        const float2 nmean = f2(-mean);
        const float2 is2 = f2(invstd);
        float2 gv[CACHE ? CH : 1][H];
        float2 xv[CACHE ? CH : 1][H];
        float2 s1 = f2(0.f), s2 = f2(0.f);
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          const int c = g.gtid + j * TPR;
          if (c < nchunks) {
            float2 dy[H], xc[H], gg[H];
            IO<T>::unpack2(ld_smem_u4(dyrow + 16u * c), dy);
            IO<T>::unpack2(ld_smem_u4(xrow + 16u * c), xc);
            if (has_w) {
              float2 w[H];
              load_param<VEC>(p.off_weight, nchunks, c, w);
#pragma unroll
              for (int k = 0; k < H; ++k) gg[k] = f2mul(dy[k], w[k]);
            } else {
#pragma unroll
              for (int k = 0; k < H; ++k) gg[k] = dy[k];
            }
#pragma unroll
            for (int k = 0; k < H; ++k) {
              if constexpr (LN) xc[k] = f2add(xc[k], nmean);
              s1 = f2add(s1, gg[k]);
              s2 = f2fma(gg[k], xc[k], s2);
              dw_acc[j][k] = f2fma(dy[k], f2mul(xc[k], is2), dw_acc[j][k]);
              if constexpr (LN) db_acc[j][k] = f2add(db_acc[j][k], dy[k]);
              if constexpr (CACHE) {
                gv[j][k] = gg[k];
                xv[j][k] = xc[k];
              }
            }
          }
        }
        if (want_dx) {
          float s[2] = {s1.x + s1.y, s2.x + s2.y};
          g.template allreduce<SumOp, 2>(s);
          if (CACHE && last_row) ring.refill(tk);  // g / xc live in registers: the stage is free
          const float a_n = invstd * p.inv_n;
          const float2 A = is2;
          const float2 B = f2(-(a_n * invstd * invstd) * s[1]);
          const float2 C = f2(centered ? -(a_n * s[0]) : 0.f);
          uint8_t* dxrow = out_row_ptr<T>(p, row, g.crank);
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            const int c = g.gtid + j * TPR;
            if (c < nchunks) {
              float2 dx[H];
              if constexpr (CACHE) {
#pragma unroll
                for (int k = 0; k < H; ++k) dx[k] = f2fma(A, gv[j][k], f2fma(B, xv[j][k], C));
              } else {
                float2 dy[H], xc[H];
                IO<T>::unpack2(ld_smem_u4(dyrow + 16u * c), dy);
                IO<T>::unpack2(ld_smem_u4(xrow + 16u * c), xc);
                if (has_w) {
                  float2 w[H];
                  load_param<VEC>(p.off_weight, nchunks, c, w);
#pragma unroll
                  for (int k = 0; k < H; ++k) dy[k] = f2mul(dy[k], w[k]);
                }
#pragma unroll
                for (int k = 0; k < H; ++k) {
                  if constexpr (LN) xc[k] = f2add(xc[k], nmean);
                  dx[k] = f2fma(A, dy[k], f2fma(B, xc[k], C));
                }
              }
              st_global_stream(dxrow + 16u * c, IO<T>::pack2(dx));
            }
          }
        }
      }
      if (!CACHE || !want_dx) {  // the stage was read until the end of the tile
        g.sync();
        ring.refill(tk);
      }
    }

    // ---- combine the groups' column accumulators and emit this CTA's partial row ----
    if (want_dw || want_db) {
      const int nct = p.groups * TPR;
      float* acc_w = reinterpret_cast<float*>(smem_raw + p.off_acc);
      float* acc_b = acc_w + p.cols_cta;
      named_bar_sync(15, nct);  // every group is done with the ring
      for (int gi = 0; gi < p.groups; ++gi) {
        if (g.gid == gi) {
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            const int c = g.gtid + j * TPR;
            if (c < nchunks) {
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

// kernels_generic.cuh — shape-agnostic kernels: any N, any row stride, any
// alignment. One CTA per row (grid-stride over rows), scalar coalesced accesses,
// the row is re-read from L1/L2 between passes. These are the landing spot for
// the shapes the reference tests that no vectorised path can take (hidden sizes
// 3, 1987, N+2 ..., tests/cpp/test_combined_inner_outer_reduction.cpp:143-166) and
// for rows too wide for any persistent variant (the reference then segments into
// Reduction + pointwise kernels, scheduler/normalization_inner.cpp:153-161).
// Also here: the outer (column) reduction used for dgamma/dbeta in that case
// (scheduler/reduction.cpp, the Reduction scheduler's outer case) and the
// finishing kernel for fp32 partial column sums.
#pragma once

#include "ring.cuh"

namespace b200norm {

constexpr int kGenericThreads = 256;

template <typename Op, int NV>
__device__ __forceinline__ void block_allreduce(float (&v)[NV], float* red /* [2][8][kRedNV] */,
                                                uint32_t& phase) {
  warp_allreduce<Op, NV>(v);
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  float* slots = red + phase * (kGenericThreads / 32) * kRedNV;
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) slots[warp * kRedNV + i] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = Op::identity();
#pragma unroll
  for (int w = 0; w < kGenericThreads / 32; ++w) {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = Op::apply(v[i], slots[w * kRedNV + i]);
  }
  phase ^= 1u;
}

enum GenericOp : int {
  G_LAYERNORM_FWD = 0,
  G_LAYERNORM_BWD = 1,
  G_RMSNORM_FWD = 2,
  G_RMSNORM_BWD = 3,
  G_SOFTMAX_FWD = 4,
  G_SOFTMAX_BWD = 5,
  G_LOGSOFTMAX_FWD = 6,
  G_LOGSOFTMAX_BWD = 7,
  G_REDUCE_INNER = 8
};

template <typename T, int OP>
__global__ void __launch_bounds__(kGenericThreads) generic_row_kernel(const RowParams p) {
  __shared__ float red[2 * (kGenericThreads / 32) * kRedNV];
  uint32_t phase = 0;
  const T* __restrict__ in0 = static_cast<const T*>(p.in0);
  const T* __restrict__ in1 = static_cast<const T*>(p.in1);
  T* __restrict__ out0 = static_cast<T*>(p.out0);
  const T* __restrict__ wgt = static_cast<const T*>(p.weight);
  const T* __restrict__ bia = static_cast<const T*>(p.bias);
  const int N = p.cols;
  const int tid = threadIdx.x;
  for (int64_t row = blockIdx.x; row < p.rows; row += gridDim.x) {
    const T* a = in0 + row * p.in0_stride;
    const T* b = in1 != nullptr ? in1 + row * p.in1_stride : nullptr;
    T* o = out0 != nullptr ? out0 + row * p.out_stride : nullptr;
    // fused residual add (forward norms): the normalised value is x + residual, rounded through
    // the io dtype when the fusion does so; the sum is written out once if requested
    const bool add_res = (OP == G_RMSNORM_FWD || OP == G_LAYERNORM_FWD) && (p.flags & kFlagAddResidual) != 0;
    const bool round_sum = (p.flags & kFlagRoundSum) != 0;
    // fused dropout (benchmarks/python/test_dropout_{layernorm,rmsnorm}_fwd.py: norm(input2 + dropout(input1))):
    // the mask is drawn once (Philox, element index row * N + c), written out, and read back by the later passes
    // — every thread re-reads exactly the bytes it wrote
    const bool drop_fwd = add_res && (p.flags & kFlagDropout) != 0;
    uint8_t* mrow = drop_fwd ? p.mask_out + (size_t)row * (size_t)N : nullptr;
    if (drop_fwd) {
      const unsigned long long seed = p.rng[0], offset = p.rng[1];
      for (int c = tid; c < N; c += kGenericThreads)
        mrow[c] = dropout_keep_at(seed, offset, (unsigned long long)row * (unsigned long long)N + (unsigned long long)c,
                                  p.drop_t1, p.drop_t2) ? 1 : 0;
    }
    auto xin = [&](int c) -> float {
      float x = IO<T>::to_f32(a[c]);
      if (drop_fwd) x = mrow[c] ? x * p.drop_scale : 0.f;
      if (add_res) {
        x += IO<T>::to_f32(b[c]);
        if (round_sum) x = IO<T>::round_trip(x);
      }
      return x;
    };
    // backward of the same: x_eff = residual + dropout(x) rebuilt from the saved mask
    const bool drop_bwd = (OP == G_RMSNORM_BWD || OP == G_LAYERNORM_BWD) && (p.flags & kFlagDropout) != 0;
    const T* res_b = drop_bwd ? static_cast<const T*>(p.in2) + row * p.in2_stride : nullptr;
    const uint8_t* mrow_b = drop_bwd ? p.mask_in + (size_t)row * (size_t)N : nullptr;
    auto xeff = [&](int c) -> float {
      float x = IO<T>::to_f32(b[c]);
      if (drop_bwd) x = IO<T>::to_f32(res_b[c]) + (mrow_b[c] ? x * p.drop_scale : 0.f);
      return x;
    };
    if ((OP == G_RMSNORM_FWD || OP == G_LAYERNORM_FWD) && add_res && p.out1 != nullptr) {
      T* so = static_cast<T*>(p.out1) + row * p.out1_stride;
      for (int c = tid; c < N; c += kGenericThreads) {
        float x = IO<T>::to_f32(a[c]);
        if (drop_fwd) x = mrow[c] ? x * p.drop_scale : 0.f;
        so[c] = IO<T>::from_f32(x + IO<T>::to_f32(b[c]));
      }
    }

    if constexpr (OP == G_RMSNORM_FWD) {
      float ss[1] = {0.f};
      for (int c = tid; c < N; c += kGenericThreads) {
        const float x = xin(c);
        ss[0] = fmaf(x, x, ss[0]);
      }
      block_allreduce<SumOp, 1>(ss, red, phase);
      const float ve = ss[0] * p.inv_n + p.eps;
      const float invstd = __frsqrt_rn(ve);
      const float rms = __fsqrt_rn(ve);
      const float scale = (p.flags & kFlagStatIsRms) ? __frcp_rn(rms) : invstd;
      for (int c = tid; c < N; c += kGenericThreads) {
        float y = xin(c) * scale;
        if (wgt != nullptr) y *= IO<T>::to_f32(wgt[c]);
        o[c] = IO<T>::from_f32(y);
      }
      if (tid == 0 && p.stat1 != nullptr) p.stat1[row] = (p.flags & kFlagStatIsRms) ? rms : invstd;
    } else if constexpr (OP == G_LAYERNORM_FWD) {
      float s[1] = {0.f};
      for (int c = tid; c < N; c += kGenericThreads) s[0] += xin(c);
      block_allreduce<SumOp, 1>(s, red, phase);
      const float mean = s[0] * p.inv_n;
      float m2[1] = {0.f};
      for (int c = tid; c < N; c += kGenericThreads) {
        const float d = xin(c) - mean;
        m2[0] = fmaf(d, d, m2[0]);
      }
      block_allreduce<SumOp, 1>(m2, red, phase);
      const float invstd = __frsqrt_rn(m2[0] * p.inv_n + p.eps);
      if (o != nullptr) {  // absent for a stand-alone var_mean / welford segment (kFlagNoY)
        const bool row_aff = (p.flags & kFlagRowAffine) != 0;  // instance norm: the row's channel pair
        const float w_row = row_aff ? row_param<T>(p.weight, row, p.channels, 1.f) : 1.f;
        const float b_row = row_aff ? row_param<T>(p.bias, row, p.channels, 0.f) : 0.f;
        for (int c = tid; c < N; c += kGenericThreads) {
          float y = (xin(c) - mean) * invstd;
          if (row_aff) {
            y = fmaf(y, w_row, b_row);
          } else {
            if (wgt != nullptr) y *= IO<T>::to_f32(wgt[c]);
            if (bia != nullptr) y += IO<T>::to_f32(bia[c]);
          }
          o[c] = IO<T>::from_f32(y);
        }
      }
      if (tid == 0) {
        if (p.stat0 != nullptr) p.stat0[row] = mean;
        if (p.stat1 != nullptr) p.stat1[row] = ln_second_stat(p, m2[0], invstd);
      }
    } else if constexpr (OP == G_SOFTMAX_FWD || OP == G_LOGSOFTMAX_FWD) {
      constexpr bool LOG = OP == G_LOGSOFTMAX_FWD;
      float m[1] = {-INFINITY};
      for (int c = tid; c < N; c += kGenericThreads) m[0] = MaxOp::apply(m[0], IO<T>::to_f32(a[c]));
      block_allreduce<MaxOp, 1>(m, red, phase);
      float s[1] = {0.f};
      for (int c = tid; c < N; c += kGenericThreads) s[0] += exp_for<T>(IO<T>::to_f32(a[c]) - m[0]);
      block_allreduce<SumOp, 1>(s, red, phase);
      const float fin = LOG ? logf(s[0]) : __frcp_rn(s[0]);
      for (int c = tid; c < N; c += kGenericThreads) {
        const float d = IO<T>::to_f32(a[c]) - m[0];
        o[c] = IO<T>::from_f32(LOG ? (d - fin) : (exp_for<T>(d) * fin));
      }
    } else if constexpr (OP == G_SOFTMAX_BWD || OP == G_LOGSOFTMAX_BWD) {
      constexpr bool LOG = OP == G_LOGSOFTMAX_BWD;
      float s[1] = {0.f};
      for (int c = tid; c < N; c += kGenericThreads) {
        const float dy = IO<T>::to_f32(a[c]);
        if constexpr (LOG) {
          s[0] += dy;
        } else {
          s[0] = fmaf(dy, IO<T>::to_f32(b[c]), s[0]);
        }
      }
      block_allreduce<SumOp, 1>(s, red, phase);
      float sum = s[0];
      if (p.flags & kFlagRoundRowsum) sum = IO<T>::round_trip(sum);
      for (int c = tid; c < N; c += kGenericThreads) {
        const float dy = IO<T>::to_f32(a[c]);
        const float y = IO<T>::to_f32(b[c]);
        o[c] = IO<T>::from_f32(LOG ? (dy - expf(y) * sum) : (y * dy - y * sum));
      }
    } else if constexpr (OP == G_LAYERNORM_BWD || OP == G_RMSNORM_BWD) {
      constexpr bool LN = OP == G_LAYERNORM_BWD;
      // a = dy, b = x; dx only (dgamma/dbeta come from generic_colsum_kernel)
      const bool centered = LN || (p.flags & kFlagRmsBwdCentered) != 0;
      const float mean = LN ? p.stat0[row] : 0.f;
      const float st = p.stat1[row];
      const float invstd = (!LN && (p.flags & kFlagStatIsRms)) ? __frcp_rn(st) : st;
      // instance norm: g = dy (the row's weight is a common factor of dx), row sums out instead of column sums
      const bool row_aff = (p.flags & kFlagRowAffine) != 0;
      const float w_row = row_aff ? row_param<T>(p.weight, row, p.channels, 1.f) : 1.f;
      if (row_aff) wgt = nullptr;
      float s[2] = {0.f, 0.f};
      for (int c = tid; c < N; c += kGenericThreads) {
        const float xh = (xeff(c) - mean) * invstd;
        float gg = IO<T>::to_f32(a[c]);
        if (wgt != nullptr) gg *= IO<T>::to_f32(wgt[c]);
        s[0] += gg;
        s[1] = fmaf(gg, xh, s[1]);
      }
      block_allreduce<SumOp, 2>(s, red, phase);
      const float s1 = centered ? s[0] : 0.f;
      const float aa = invstd * p.inv_n * w_row;
      if (row_aff && tid == 0) {
        if (p.rowsum_dy != nullptr) p.rowsum_dy[row] = s[0];
        if (p.rowsum_dyxh != nullptr) p.rowsum_dyxh[row] = s[1];
      }
      if (o == nullptr && !(drop_bwd && p.out1 != nullptr)) continue;
      T* o1 = drop_bwd && p.out1 != nullptr ? static_cast<T*>(p.out1) + row * p.out1_stride : nullptr;
      for (int c = tid; c < N; c += kGenericThreads) {
        const float xh = (xeff(c) - mean) * invstd;
        float gg = IO<T>::to_f32(a[c]);
        if (wgt != nullptr) gg *= IO<T>::to_f32(wgt[c]);
        const float dxe = aa * ((p.n_f * gg - s1) - xh * (s[1] * p.proj_scale));
        if (drop_bwd) {
          // d residual = d x_eff; d x = d x_eff * scale * mask (test_dropout_layernorm_bwd.py:100-101)
          if (o1 != nullptr) o1[c] = IO<T>::from_f32(dxe);
          if (o != nullptr) o[c] = IO<T>::from_f32(mrow_b[c] ? dxe * p.drop_scale : 0.f);
        } else {
          o[c] = IO<T>::from_f32(dxe);
        }
      }
    } else if constexpr (OP == G_REDUCE_INNER) {
      // sum / amax / fmax over the row -> stat1-less: result goes to out0[row] (io dtype)
      T* r = static_cast<T*>(p.out0);
      if (p.flags & kFlagReduceNll) {  // cross-entropy forward (see reduce_inner_ring_kernel)
        float m[1] = {-INFINITY};
        for (int c = tid; c < N; c += kGenericThreads) m[0] = MaxOp::apply(m[0], IO<T>::to_f32(a[c]));
        block_allreduce<MaxOp, 1>(m, red, phase);
        float s[1] = {0.f};
        for (int c = tid; c < N; c += kGenericThreads) s[0] += exp_for<T>(IO<T>::to_f32(a[c]) - m[0]);
        block_allreduce<SumOp, 1>(s, red, phase);
        if (tid == 0) {
          const long long t = p.target[row];
          const bool valid = t != (long long)p.ignore_index && t >= 0 && t < (long long)N;
          const float lse = m[0] + logf(s[0]);
          if (p.stat1 != nullptr) p.stat1[row] = valid ? lse - IO<T>::to_f32(a[t]) : 0.f;
          if (p.stat0 != nullptr) p.stat0[row] = lse;
        }
      } else if (p.flags & (kFlagReduceMax | kFlagReduceFmax)) {
        const bool nan_prop = (p.flags & kFlagReduceMax) != 0;
        float m[2] = {-INFINITY, 0.f};  // m[1] counts non-NaN elements (for all-NaN rows under fmax)
        if (nan_prop) {
          for (int c = tid; c < N; c += kGenericThreads) m[0] = MaxOp::apply(m[0], IO<T>::to_f32(a[c]));
          float mm[1] = {m[0]};
          block_allreduce<MaxOp, 1>(mm, red, phase);
          if (tid == 0) r[row] = IO<T>::from_f32(mm[0]);
        } else {
          for (int c = tid; c < N; c += kGenericThreads) {
            const float x = IO<T>::to_f32(a[c]);
            if (x == x) {
              m[0] = fmaxf(m[0], x);
              m[1] += 1.f;
            }
          }
          float mm[1] = {m[0]};
          float cnt[1] = {m[1]};
          block_allreduce<FMaxOp, 1>(mm, red, phase);
          block_allreduce<SumOp, 1>(cnt, red, phase);
          if (tid == 0) r[row] = IO<T>::from_f32(cnt[0] > 0.f ? mm[0] : __int_as_float(0x7fc00000));
        }
      } else {
        float s[1] = {0.f};
        for (int c = tid; c < N; c += kGenericThreads) s[0] += IO<T>::to_f32(a[c]);
        block_allreduce<SumOp, 1>(s, red, phase);
        if (tid == 0) r[row] = IO<T>::from_f32(s[0]);
      }
    }
  }
}

// Cross-entropy backward (B200NORM_OP_NLL_SOFTMAX_BWD): dx = (exp(x - lse) - onehot(target)) * g, zero rows where the
// target is ignored. Pointwise given the saved log-sum-exp: no ring, no reduction — a work item is (row, block of
// kCeU * 256 chunks); every thread keeps kCeU 16-byte loads in flight. VEC = 1 is the element-wise form for rows
// that are not whole aligned chunks.
constexpr int kCeThreads = 256;
constexpr int kCeU = 4;
struct CeBwdParams {
  const void* x;
  void* dx;
  int64_t x_stride, dx_stride;  // elements
  const long long* target;
  const float* lse;
  const float* gscale;  // device scalar
  int64_t rows;
  int32_t cols;
  int32_t ignore_index;
};
template <typename T, bool VECTOR>
__global__ void __launch_bounds__(kCeThreads) cross_entropy_bwd_kernel(const CeBwdParams p) {
  constexpr int VEC = VECTOR ? IO<T>::VEC : 1;
  const int units_per_row = (p.cols + VEC - 1) / VEC;                     // 16-byte chunks (or elements)
  const int items_per_row = (units_per_row + kCeThreads * kCeU - 1) / (kCeThreads * kCeU);
  const int64_t items = p.rows * (int64_t)items_per_row;
  const float g0 = __ldg(p.gscale);
  for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
    const int64_t row = item / items_per_row;
    const int cb = (int)(item - row * items_per_row);
    const long long t = __ldg(p.target + row);
    const bool valid = t != (long long)p.ignore_index && t >= 0 && t < (long long)p.cols;
    const float gs = valid ? g0 : 0.f;
    const float lse = __ldg(p.lse + row);
    const T* xr = static_cast<const T*>(p.x) + row * p.x_stride;
    T* dr = static_cast<T*>(p.dx) + row * p.dx_stride;
    const int u0 = cb * kCeThreads * kCeU + (int)threadIdx.x;
    if constexpr (VECTOR) {
      constexpr int H = VEC / 2;
      uint4 q[kCeU];
#pragma unroll
      for (int u = 0; u < kCeU; ++u) {
        const int unit = u0 + u * kCeThreads;
        if (unit < units_per_row) q[u] = ldg_128_nc(reinterpret_cast<const uint8_t*>(xr) + 16 * (size_t)unit);
      }
#pragma unroll
      for (int u = 0; u < kCeU; ++u) {
        const int unit = u0 + u * kCeThreads;
        if (unit >= units_per_row) continue;
        float2 v[H];
        IO<T>::unpack2(q[u], v);
        const int c0 = unit * VEC;
#pragma unroll
        for (int k = 0; k < H; ++k) {
          v[k].x = exp_for<T>(v[k].x - lse) * gs;
          v[k].y = exp_for<T>(v[k].y - lse) * gs;
        }
        if (valid && t >= c0 && t < c0 + VEC) {
          const int e = (int)(t - c0);
#pragma unroll
          for (int k = 0; k < H; ++k) {
            if (e == 2 * k) v[k].x -= gs;
            if (e == 2 * k + 1) v[k].y -= gs;
          }
        }
        stg_128_stream(reinterpret_cast<uint8_t*>(dr) + 16 * (size_t)unit, IO<T>::pack2(v));
      }
    } else {
#pragma unroll
      for (int u = 0; u < kCeU; ++u) {
        const int c = u0 + u * kCeThreads;
        if (c < p.cols) {
          float y = exp_for<T>(IO<T>::to_f32(xr[c]) - lse) * gs;
          if (valid && c == t) y -= gs;
          dr[c] = IO<T>::from_f32(y);
        }
      }
    }
  }
}

// Column sums over a row range: partials[split][K][N] (fp32).
//   MODE 0: plain sum of in0                               (K = 1)
//   MODE 1: LayerNorm  dw = sum dy*x_hat, db = sum dy      (K = 2)
//   MODE 2: RMSNorm    dw = sum dy*x_hat                   (K = 1)
// grid = (ceil(N/128), nsplit), block = 128: one column per thread, coalesced.
template <typename T, int MODE>
__global__ void __launch_bounds__(128) generic_colsum_kernel(const RowParams p, int rows_per_split) {
  const int col = blockIdx.x * 128 + threadIdx.x;
  if (col >= p.cols) return;
  const T* __restrict__ in0 = static_cast<const T*>(p.in0);
  const T* __restrict__ in1 = static_cast<const T*>(p.in1);
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_split;
  const int64_t r1 = min(p.rows, r0 + rows_per_split);
  float acc_w = 0.f, acc_b = 0.f;
  for (int64_t r = r0; r < r1; ++r) {
    const float a = IO<T>::to_f32(in0[r * p.in0_stride + col]);
    if constexpr (MODE == 0) {
      acc_w += a;
    } else {
      const float mean = MODE == 1 ? p.stat0[r] : 0.f;
      const float st = p.stat1[r];
      const float invstd = (MODE == 2 && (p.flags & kFlagStatIsRms)) ? __frcp_rn(st) : st;
      float x = IO<T>::to_f32(in1[r * p.in1_stride + col]);
      if (p.flags & kFlagDropout)
        x = IO<T>::to_f32(static_cast<const T*>(p.in2)[r * p.in2_stride + col]) +
            (p.mask_in[(size_t)r * (size_t)p.cols + col] ? x * p.drop_scale : 0.f);
      const float xh = (x - mean) * invstd;
      acc_w = fmaf(a, xh, acc_w);
      acc_b += a;
    }
  }
  constexpr int K = MODE == 1 ? 2 : 1;
  float* dst = p.partials + (size_t)blockIdx.y * K * p.cols;
  dst[col] = acc_w;
  if constexpr (MODE == 1) dst[p.cols + col] = acc_b;
}

// Finish fp32 partial column sums: out_k[n] = cast(sum_p partials[p][k][n]).
// block = (32, 16): x walks float4 column vectors (128 columns per block), y splits the
// partial rows; smem tree over y. With out_f32 != nullptr the fp32 sums are written
// instead (multi-GPU: the caller allreduces them, then runs this kernel again with
// nparts = 1). N must be a multiple of 4 for the vector path; VECN = 1 is the scalar form.
constexpr int kFinY = 16;
template <typename T, int VECN>
__global__ void __launch_bounds__(32 * kFinY) finalize_partials_kernel(const float* __restrict__ partials,
                                                                       int nparts, int K, int N,
                                                                       void* out_a, void* out_b,
                                                                       float* out_f32) {
  __shared__ float sm[kFinY][32 * VECN + 1];
  const int col = (blockIdx.x * 32 + threadIdx.x) * VECN;
  const int k = blockIdx.y;
  float acc[VECN];
#pragma unroll
  for (int i = 0; i < VECN; ++i) acc[i] = 0.f;
  if (col < N) {
    const float* src = partials + (size_t)k * N + col;
    const size_t pstride = (size_t)K * N;
#pragma unroll 4
    for (int pidx = threadIdx.y; pidx < nparts; pidx += kFinY) {
      if constexpr (VECN == 4) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(src + pidx * pstride));
        acc[0] += v.x;
        acc[1] += v.y;
        acc[2] += v.z;
        acc[3] += v.w;
      } else {
        acc[0] += __ldcg(src + pidx * pstride);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < VECN; ++i) sm[threadIdx.y][threadIdx.x * VECN + i] = acc[i];
  __syncthreads();
  // threads (x, y < VECN) finish column col_base + x*VECN + y ... remapped so that stores coalesce
  const int t = threadIdx.y * 32 + threadIdx.x;
  if (t < 32 * VECN) {
    const int c = blockIdx.x * 32 * VECN + t;
    if (c < N) {
      float r = 0.f;
#pragma unroll
      for (int y = 0; y < kFinY; ++y) r += sm[y][t];
      if (out_f32 != nullptr) {
        out_f32[(size_t)k * N + c] = r;
      } else {
        T* o = static_cast<T*>(k == 0 ? out_a : out_b);
        if (o != nullptr) o[c] = IO<T>::from_f32(r);
      }
    }
  }
}

// OP_PEER_COMBINE: finish a published epoch. A few CTAs (one remote load per thread and peer — the
// loads are latency bound): wait until every rank's "published" flag in this rank's header carries
// the epoch (normally long true), add the P slots in rank order (the same order on every rank:
// identical bits everywhere), cast once; the last CTA to finish tells every peer that its slot has
// been consumed (counter in this rank's header, word 32).
struct PeerCombineParams {
  int32_t world, rank, K, N;
  uint32_t epoch;
  uint32_t slot_floats;
  uint64_t timeout_ns;  // limit of the wait for the peers' flags (0 = for ever)
  uint8_t* xchg[kMaxPeers];
  void* out_a;
  void* out_b;
};
template <typename T>
__global__ void __launch_bounds__(512) peer_combine_kernel(const PeerCombineParams q) {
  if ((int)threadIdx.x < q.world) {
    const uint32_t* flag = reinterpret_cast<const uint32_t*>(q.xchg[q.rank]) + threadIdx.x;
    Watchdog wd(q.timeout_ns);
    while ((int32_t)(ld_acquire_sys(flag) - q.epoch) < 0) wd.tick();
  }
  __syncthreads();
  const size_t slot = (size_t)(q.epoch & 1u) * q.slot_floats;
  const int total = q.K * q.N;
  if (q.N % 4 == 0) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total / 4; i += gridDim.x * blockDim.x) {
      float4 a = ld_volatile_f4(reinterpret_cast<const float4*>(
                                    reinterpret_cast<const float*>(q.xchg[0] + kPeerHeaderBytes) + slot) + i);
      for (int r = 1; r < q.world; ++r) {
        const float4 v = ld_volatile_f4(reinterpret_cast<const float4*>(
                                            reinterpret_cast<const float*>(q.xchg[r] + kPeerHeaderBytes) + slot) + i);
        a.x += v.x;
        a.y += v.y;
        a.z += v.z;
        a.w += v.w;
      }
      const int e = i * 4;
      const int k2 = e / q.N;
      const int c = e - k2 * q.N;
      T* o = static_cast<T*>(k2 == 0 ? q.out_a : q.out_b);
      if (o != nullptr) {
        o[c] = IO<T>::from_f32(a.x);
        o[c + 1] = IO<T>::from_f32(a.y);
        o[c + 2] = IO<T>::from_f32(a.z);
        o[c + 3] = IO<T>::from_f32(a.w);
      }
    }
  }
  __syncthreads();
  __shared__ uint32_t last;
  if (threadIdx.x == 0) {
    uint32_t* cnt = reinterpret_cast<uint32_t*>(q.xchg[q.rank] + 128);
    const uint32_t old = atom_add_acq_rel_gpu(cnt, 1u);
    last = old == gridDim.x - 1 ? 1u : 0u;
    if (last) *cnt = 0u;
  }
  __syncthreads();
  if (last && (int)threadIdx.x < q.world) {
    st_release_sys(reinterpret_cast<uint32_t*>(q.xchg[threadIdx.x] + kPeerConsumedOffset) + q.rank, q.epoch);
  }
}

}  // namespace b200norm

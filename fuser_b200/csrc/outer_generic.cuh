// outer_generic.cuh — shape-agnostic kernels of the outer (dim 0) ops: any M, any N, any row stride, any
// alignment. The landing spot for what the outer-persistent kernels decline (rows that are not a multiple of
// 16 bytes, misaligned bases, a reduced extent beyond what a 16-CTA cluster holds in shared memory — e.g. the
// ResNet-size [N*H*W, C] batch norms) — the role kernels_generic.cuh plays for the inner ops. The reference
// segments such fusions into a Reduction kernel + a pointwise kernel (scheduler/normalization_outer.cpp:451-683
// canSchedule; scheduler/reduction.cpp outer case); this is the same split:
//   outer_generic_stats : grid (column blocks x row splits); a thread owns ONE column of one row split and walks
//                         its rows once (coalesced across the 128 threads of a CTA), keeping a single-pass partial:
//                         online (max, sum-exp) for softmax, shifted-sum (mean, M2) for the normalization, plain sums
//                         for the backwards; fp32 partials -> workspace [split][K][N];
//   outer_generic_merge : one thread per column merges the splits' partials in split order (max / rescale, Chan's
//                         parallel Welford merge, sums: deterministic) and writes the per-column outputs;
//   outer_generic_apply : the statistics grid again; applies the elementwise formula to its rows.
// The matrix is read twice (1.5x the algorithmic traffic of the forward). Semantics (NaN / +-inf columns,
// io-dtype rounding) are those of the persistent kernels and of the op-by-op ATen evaluation.
#pragma once

#include "common.cuh"

namespace b200norm {

constexpr int kOuterGenThreads = 128;
constexpr int kOuterGenMaxSplits = 2048;

struct OuterGenParams {
  const void* in0;
  const void* in1;
  void* out;
  int64_t s0, s1, so;  // row strides in elements
  int32_t M, N;
  int32_t rows_per_split, nsplit;
  float* part;  // [nsplit][K][N]
  const void* weight;
  const void* bias;
  float* mean_out;
  float* invstd_out;
  const float* mean_in;
  const float* invstd_in;
  void* dw_out;
  void* db_out;
  float inv_m, eps;
  int32_t round_sum;
};

// KIND as in outer.cu: 0 softmax fwd, 1 log-softmax fwd, 2 softmax bwd, 3 log-softmax bwd, 4 norm fwd, 5 norm bwd
template <int KIND>
struct OuterGenK {
  static constexpr int value = (KIND == 2 || KIND == 3) ? 1 : 2;
};

template <typename T, int KIND>
__global__ void __launch_bounds__(kOuterGenThreads) outer_generic_stats(const OuterGenParams p) {
  const int col = blockIdx.x * kOuterGenThreads + threadIdx.x;
  if (col >= p.N) return;
  const int split = blockIdx.y;
  const int r0 = split * p.rows_per_split;
  const int r1 = min(p.M, r0 + p.rows_per_split);
  const T* a = static_cast<const T*>(p.in0) + col;
  const T* b = static_cast<const T*>(p.in1) + col;
  float v0 = 0.f, v1 = 0.f;
  if constexpr (KIND == 0 || KIND == 1) {
    float m = -INFINITY, s = 0.f;
#pragma unroll 4
    for (int r = r0; r < r1; ++r) {
      const float x = IO<T>::to_f32(a[(int64_t)r * p.s0]);
      if (x > m) {
        s = x == INFINITY ? NAN : s * expf(m - x) + 1.f;  // +inf: inf - inf = NaN poisons the column, as in ATen
        m = x;
      } else if (!(x == -INFINITY && m == -INFINITY)) {
        s += expf(x - m);  // NaN lands here and stays in s
      }
    }
    v0 = m;
    v1 = s;
  } else if constexpr (KIND == 4) {
    // (mean, M2) of the split by shifted sums: shift = the split's first element, so the sums stay of the order
    // of the spread (no catastrophic cancellation, no division in the loop); exact merge across splits below
    float sh = 0.f, s1 = 0.f, s2 = 0.f;
    if (r0 < r1) sh = IO<T>::to_f32(a[(int64_t)r0 * p.s0]);
#pragma unroll 4
    for (int r = r0; r < r1; ++r) {
      const float d = IO<T>::to_f32(a[(int64_t)r * p.s0]) - sh;
      s1 += d;
      s2 = fmaf(d, d, s2);
    }
    const float n = (float)(r1 - r0);
    if (n > 0.f) {
      const float dm = s1 / n;
      v0 = sh + dm;
      v1 = fmaxf(s2 - s1 * dm, 0.f);
    }
  } else if constexpr (KIND == 2 || KIND == 3) {
#pragma unroll 4
    for (int r = r0; r < r1; ++r) {
      const float dy = IO<T>::to_f32(a[(int64_t)r * p.s0]);
      if constexpr (KIND == 3) {
        v0 += dy;
      } else {
        v0 = fmaf(IO<T>::to_f32(b[(int64_t)r * p.s1]), dy, v0);
      }
    }
  } else {
    const float mean = p.mean_in[col];
#pragma unroll 4
    for (int r = r0; r < r1; ++r) {
      const float dy = IO<T>::to_f32(a[(int64_t)r * p.s0]);
      const float x = IO<T>::to_f32(b[(int64_t)r * p.s1]);
      v0 += dy;
      v1 = fmaf(dy, x - mean, v1);
    }
  }
  constexpr int K = OuterGenK<KIND>::value;
  p.part[((size_t)split * K + 0) * p.N + col] = v0;
  if constexpr (K == 2) p.part[((size_t)split * K + 1) * p.N + col] = v1;
}

// Merge of the row splits' partials, one thread per column, in split order (deterministic): writes the final
// per-column values behind the partials ([K][N] at part + nsplit*K*N) and the per-column outputs.
template <typename T, int KIND>
__global__ void __launch_bounds__(kOuterGenThreads) outer_generic_merge(const OuterGenParams p) {
  const int col = blockIdx.x * kOuterGenThreads + threadIdx.x;
  if (col >= p.N) return;
  constexpr int K = OuterGenK<KIND>::value;
  auto part = [&](int q, int k) { return p.part[((size_t)q * K + k) * p.N + col]; };
  float* fin = p.part + (size_t)p.nsplit * K * p.N;
  float f0 = 0.f, f1 = 0.f;
  if constexpr (KIND == 0 || KIND == 1) {
    float M = part(0, 0);
    for (int q = 1; q < p.nsplit; ++q) M = fmaxf(M, part(q, 0));  // the maxima are never NaN (NaN lives in the sums)
    float S = 0.f;
    for (int q = 0; q < p.nsplit; ++q) {
      const float mq = part(q, 0);
      S = fmaf(part(q, 1), mq == M ? 1.f : expf(mq - M), S);
    }
    f0 = M;
    f1 = KIND == 1 ? -logf(S) : __frcp_rn(S);
  } else if constexpr (KIND == 4) {
    // exact merge of the splits' (count, mean, M2): mean = sum n_q mean_q / M, M2 = sum (M2_q + n_q (mean_q - mean)^2)
    // — two passes over the partials, no dependent divisions
    const int last = p.nsplit - 1;
    const float nfull = (float)p.rows_per_split, nlast = (float)(p.M - last * p.rows_per_split);
    float sm = 0.f;
#pragma unroll 4
    for (int q = 0; q < p.nsplit; ++q) sm = fmaf(q == last ? nlast : nfull, part(q, 0), sm);
    const float mean = sm * p.inv_m;
    float m2 = 0.f;
#pragma unroll 4
    for (int q = 0; q < p.nsplit; ++q) {
      const float d = part(q, 0) - mean;
      m2 += fmaf((q == last ? nlast : nfull) * d, d, part(q, 1));
    }
    const float invstd = rsqrtf(m2 * p.inv_m + p.eps);
    if (p.mean_out != nullptr) p.mean_out[col] = mean;
    if (p.invstd_out != nullptr) p.invstd_out[col] = invstd;
    f0 = mean;
    f1 = invstd;
  } else if constexpr (KIND == 2 || KIND == 3) {
    for (int q = 0; q < p.nsplit; ++q) f0 += part(q, 0);
    if (p.round_sum) f0 = IO<T>::round_trip(f0);
  } else {
    for (int q = 0; q < p.nsplit; ++q) {
      f0 += part(q, 0);
      f1 += part(q, 1);
    }
    if (p.dw_out != nullptr) static_cast<T*>(p.dw_out)[col] = IO<T>::from_f32(f1 * p.invstd_in[col]);
    if (p.db_out != nullptr) static_cast<T*>(p.db_out)[col] = IO<T>::from_f32(f0);
  }
  fin[col] = f0;
  if constexpr (K == 2) fin[(size_t)p.N + col] = f1;
}

template <typename T, int KIND>
__global__ void __launch_bounds__(kOuterGenThreads) outer_generic_apply(const OuterGenParams p) {
  const int col = blockIdx.x * kOuterGenThreads + threadIdx.x;
  if (col >= p.N) return;
  const int split = blockIdx.y;
  const int r0 = split * p.rows_per_split;
  const int r1 = min(p.M, r0 + p.rows_per_split);
  constexpr int K = OuterGenK<KIND>::value;
  const float* fin = p.part + (size_t)p.nsplit * K * p.N;
  const float f0 = fin[col];
  const float f1 = K == 2 ? fin[(size_t)p.N + col] : 0.f;
  const T* a = static_cast<const T*>(p.in0) + col;
  const T* b = static_cast<const T*>(p.in1) + col;
  T* o = static_cast<T*>(p.out) + col;

  if constexpr (KIND == 0 || KIND == 1) {
    for (int r = r0; r < r1; ++r) {
      const float d = IO<T>::to_f32(a[(int64_t)r * p.s0]) - f0;
      o[(int64_t)r * p.so] = IO<T>::from_f32(KIND == 1 ? d + f1 : expf(d) * f1);
    }
  } else if constexpr (KIND == 4) {
    float scale = f1, shift = 0.f;
    if (p.weight != nullptr) scale *= IO<T>::to_f32(static_cast<const T*>(p.weight)[col]);
    if (p.bias != nullptr) shift = IO<T>::to_f32(static_cast<const T*>(p.bias)[col]);
    for (int r = r0; r < r1; ++r)
      o[(int64_t)r * p.so] = IO<T>::from_f32(fmaf(IO<T>::to_f32(a[(int64_t)r * p.s0]) - f0, scale, shift));
  } else if constexpr (KIND == 2 || KIND == 3) {
    for (int r = r0; r < r1; ++r) {
      const float dy = IO<T>::to_f32(a[(int64_t)r * p.s0]);
      const float y = IO<T>::to_f32(b[(int64_t)r * p.s1]);
      o[(int64_t)r * p.so] = IO<T>::from_f32(KIND == 3 ? fmaf(expf(y), -f0, dy) : fmaf(y, -f0, y * dy));
    }
  } else {
    if (p.out == nullptr) return;
    const float mean = p.mean_in[col], invstd = p.invstd_in[col];
    const float w = p.weight != nullptr ? IO<T>::to_f32(static_cast<const T*>(p.weight)[col]) : 1.f;
    const float nps = -(f1 * p.inv_m) * (invstd * invstd), ngm = -f0 * p.inv_m, gsc = w * invstd;
    for (int r = r0; r < r1; ++r) {
      const float dy = IO<T>::to_f32(a[(int64_t)r * p.s0]);
      const float xc = IO<T>::to_f32(b[(int64_t)r * p.s1]) - mean;
      o[(int64_t)r * p.so] = IO<T>::from_f32((fmaf(nps, xc, dy) + ngm) * gsc);
    }
  }
}

}  // namespace b200norm

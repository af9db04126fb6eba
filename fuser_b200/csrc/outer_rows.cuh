// outer_rows.cuh — the outer (dim 0) normalization for FEW columns and MANY rows: the batch-norm shapes of
// convolutional nets viewed as [N*H*W, C] (e.g. [100352, 256], [802816, 64]; benchmarks/python/normalization.py,
// test_batchnorm_{fwd,bwd}.py with channels last), which no cluster of CTAs can hold in shared memory.
//
// The reference runs these as an OuterPersistent segment while the batch fits its grid-persistent registers and
// segments them into a Reduction + pointwise pair beyond (scheduler/normalization_outer.cpp:451-683,
// runtime/fused_welford_impl_outer.cu). Here ONE persistent kernel does both passes:
//   phase 1  every CTA walks a contiguous block of rows with 16-byte loads — a row is only nchunks = N*sizeof(T)/16
//            chunks wide, so thread t keeps chunk (t mod nchunks) of every (active / nchunks)-th row: a warp reads 512
//            contiguous bytes per load, each thread accumulates its 8 (4) columns in registers (forward: shifted sums ->
//            Welford triples; backward: sum dy and sum dy (x - mean)), the threads of a column are merged in shared
//            memory in a fixed order, the CTA writes its per-column partials;
//   barrier  in-kernel grid barrier (cooperative launch); CTA c merges the partials of its share of the columns in CTA
//            order (Chan's formula with counts / plain sums: deterministic), writes the per-column results (mean,
//            invstd | dweight, dbias and the dx coefficients); second grid barrier;
//   phase 2  every CTA walks ITS rows again — from L2 when the matrix fits (the [100352, 256] bf16 case is 51 MB) —
//            applies the per-column affine form and streams the output.
// HBM traffic is the algorithmic minimum while the matrix stays L2 resident and 1.5x beyond; the shape-agnostic
// three-kernel family (outer_generic.cuh: one column per thread, 2-byte loads, one thread merging thousands of
// splits) reached 4-14 % of the measured peak on these shapes.
// Requirements (outer_launch checks them): rows contiguous (stride = N), 16-byte aligned operands, N * sizeof(T) a
// multiple of 16, nchunks <= kOuterRowsThreads.
#pragma once

#include "common.cuh"

namespace b200norm {

constexpr int kOuterRowsThreads = 512;

struct OuterRowsParams {
  const void* in0;  // fwd: x ; bwd: dy
  const void* in1;  // bwd: x
  void* out;        // fwd: y ; bwd: dx (may be null)
  int64_t M;
  int32_t N;
  int32_t nchunks;       // 16-byte chunks per row
  int32_t rpi;           // rows per iteration of a CTA = active / nchunks
  int32_t active;        // threads that own a chunk = rpi * nchunks
  int64_t rows_per_cta;
  float* part;           // [grid][3][N] fp32 partials (fwd: count, mean, M2 ; bwd: sum dy, sum dy (x - mean), unused)
  float* fin;            // [2][N]: fwd (mean, invstd) ; bwd (sum dy, sum dy (x - mean))
  uint32_t* sync_word;   // zero between launches (workspace header)
  const void* weight;
  const void* bias;
  float* mean_out;
  float* invstd_out;
  const float* mean_in;
  const float* invstd_in;
  void* dw_out;
  void* db_out;
  float inv_m, eps;
};

__device__ __forceinline__ void outer_rows_grid_barrier(uint32_t* word) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t old = atom_add_acq_rel_gpu(word, 1u);
    const uint32_t sense = old & 0x80000000u;
    if ((old & 0x7fffffffu) == gridDim.x - 1) {
      atom_add_acq_rel_gpu(word, 0x80000000u - gridDim.x);
    } else {
      Watchdog wd;
      while ((ld_acquire_gpu(word) & 0x80000000u) == sense) wd.tick();
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void chan_merge3(float& n, float& mean, float& m2, float on, float om, float o2) {
  const float tot = n + on;
  const float w = tot > 0.f ? on / tot : 0.f;
  const float d = om - mean;
  mean = fmaf(d, w, mean);
  m2 = fmaf(d * d, n * w, m2 + o2);
  n = tot;
}

template <typename T, bool BWD>
__global__ void __launch_bounds__(kOuterRowsThreads, 1) outer_rows_kernel(const OuterRowsParams p) {
  constexpr int VEC = IO<T>::VEC;
  constexpr int H = VEC / 2;
  extern __shared__ __align__(16) uint8_t dyn_smem[];
  float* sm = reinterpret_cast<float*>(dyn_smem);  // [rpi][nchunks][VEC][3] staging of the per-thread partials
  const int tid = threadIdx.x;
  const bool owner = tid < p.active;
  const int cchunk = owner ? tid % p.nchunks : 0;
  const int rsub = owner ? tid / p.nchunks : 0;
  const int64_t r0 = (int64_t)blockIdx.x * p.rows_per_cta;
  const int64_t r1 = min(p.M, r0 + p.rows_per_cta);
  const size_t row_bytes = (size_t)p.nchunks * 16u;
  const uint8_t* a0 = static_cast<const uint8_t*>(p.in0) + (size_t)cchunk * 16u;
  const uint8_t* a1 = static_cast<const uint8_t*>(p.in1) + (size_t)cchunk * 16u;
  const int col0 = cchunk * VEC;

  // ------------------------------------------------------------------ phase 1: per-column partials of the row block
  float v0[VEC], v1[VEC], v2[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) v0[e] = v1[e] = v2[e] = 0.f;
  if (owner) {
    if constexpr (!BWD) {
      // shifted sums around the thread's first row (no cancellation beyond the handful of rows a thread sees), then
      // a Welford triple (count, mean, M2) per column
      float sh[VEC], s1[VEC], s2[VEC];
#pragma unroll
      for (int e = 0; e < VEC; ++e) sh[e] = s1[e] = s2[e] = 0.f;
      float cnt = 0.f;
      int64_t r = r0 + rsub;
      if (r < r1) {
        float2 t[H];
        IO<T>::unpack2(ldg_128_nc(a0 + (size_t)r * row_bytes), t);
#pragma unroll
        for (int k = 0; k < H; ++k) {
          sh[2 * k] = t[k].x;
          sh[2 * k + 1] = t[k].y;
        }
      }
#pragma unroll 8
      for (; r < r1; r += p.rpi) {
        float2 t[H];
        IO<T>::unpack2(ldg_128_nc(a0 + (size_t)r * row_bytes), t);
#pragma unroll
        for (int k = 0; k < H; ++k) {
          const float d0 = t[k].x - sh[2 * k], d1 = t[k].y - sh[2 * k + 1];
          s1[2 * k] += d0;
          s1[2 * k + 1] += d1;
          s2[2 * k] = fmaf(d0, d0, s2[2 * k]);
          s2[2 * k + 1] = fmaf(d1, d1, s2[2 * k + 1]);
        }
        cnt += 1.f;
      }
      const float inv = cnt > 0.f ? 1.f / cnt : 0.f;
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        v0[e] = cnt;
        v1[e] = fmaf(s1[e], inv, sh[e]);
        v2[e] = fmaxf(fmaf(-s1[e] * inv, s1[e], s2[e]), 0.f);
      }
    } else {
      float2 mean2[H];
#pragma unroll
      for (int k = 0; k < H; ++k) mean2[k] = make_float2(-p.mean_in[col0 + 2 * k], -p.mean_in[col0 + 2 * k + 1]);
#pragma unroll 4
      for (int64_t r = r0 + rsub; r < r1; r += p.rpi) {
        float2 dy[H], x[H];
        IO<T>::unpack2(ldg_128_nc(a0 + (size_t)r * row_bytes), dy);
        IO<T>::unpack2(ldg_128_nc(a1 + (size_t)r * row_bytes), x);
#pragma unroll
        for (int k = 0; k < H; ++k) {
          const float2 xc = f2add(x[k], mean2[k]);
          v0[2 * k] += dy[k].x;
          v0[2 * k + 1] += dy[k].y;
          v1[2 * k] = fmaf(dy[k].x, xc.x, v1[2 * k]);
          v1[2 * k + 1] = fmaf(dy[k].y, xc.y, v1[2 * k + 1]);
        }
      }
    }
    // stage the per-thread partials: sm[((rsub * nchunks + cchunk) * VEC + e) * 3 + k]
    float* dst = sm + ((size_t)(rsub * p.nchunks + cchunk) * VEC) * 3;
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      dst[3 * e + 0] = v0[e];
      dst[3 * e + 1] = v1[e];
      dst[3 * e + 2] = v2[e];
    }
  }
  __syncthreads();
  // one thread per COLUMN merges the rpi row-subsets in order and writes the CTA's partial
  float* my_part = p.part + (size_t)blockIdx.x * 3 * p.N;
  for (int c = tid; c < p.N; c += kOuterRowsThreads) {
    const int cc = c / VEC, e = c - cc * VEC;
    const float* src = sm + ((size_t)cc * VEC + e) * 3;
    const size_t stride = (size_t)p.nchunks * VEC * 3;
    float n = src[0], m = src[1], q = src[2];
    for (int s = 1; s < p.rpi; ++s) {
      const float* t = src + s * stride;
      if constexpr (!BWD) {
        chan_merge3(n, m, q, t[0], t[1], t[2]);
      } else {
        n += t[0];
        m += t[1];
      }
    }
    my_part[c] = n;
    my_part[p.N + c] = m;
    my_part[2 * (size_t)p.N + c] = q;
  }
  outer_rows_grid_barrier(p.sync_word);

  // ------------------------------------------------------------------ merge over the CTAs, in CTA order
  {
    const int total = p.N;
    const int per_cta = (total + gridDim.x - 1) / gridDim.x;
    const int c_begin = blockIdx.x * per_cta, c_end = min(total, c_begin + per_cta);
    for (int c = c_begin + tid; c < c_end; c += kOuterRowsThreads) {
      float n = __ldcg(p.part + c), m = __ldcg(p.part + p.N + c), q = __ldcg(p.part + 2 * (size_t)p.N + c);
#pragma unroll 4
      for (unsigned b = 1; b < gridDim.x; ++b) {
        const float* t = p.part + (size_t)b * 3 * p.N;
        const float tn = __ldcg(t + c), tm = __ldcg(t + p.N + c), tq = __ldcg(t + 2 * (size_t)p.N + c);
        if constexpr (!BWD) {
          chan_merge3(n, m, q, tn, tm, tq);
        } else {
          n += tn;
          m += tm;
        }
      }
      if constexpr (!BWD) {
        const float invstd = rsqrtf(q * p.inv_m + p.eps);
        if (p.mean_out != nullptr) p.mean_out[c] = m;
        if (p.invstd_out != nullptr) p.invstd_out[c] = invstd;
        p.fin[c] = m;
        p.fin[p.N + c] = invstd;
      } else {
        // n = sum dy, m = sum dy (x - mean)
        if (p.dw_out != nullptr) static_cast<T*>(p.dw_out)[c] = IO<T>::from_f32(m * p.invstd_in[c]);
        if (p.db_out != nullptr) static_cast<T*>(p.db_out)[c] = IO<T>::from_f32(n);
        p.fin[c] = n;
        p.fin[p.N + c] = m;
      }
    }
  }
  if (p.out == nullptr) return;  // backward without dx: nothing left (every CTA takes this branch together)
  outer_rows_grid_barrier(p.sync_word);

  // ------------------------------------------------------------------ phase 2: apply, row block by row block
  if (!owner) return;
  float2 sc[H], sf[H], pj[H];
#pragma unroll
  for (int k = 0; k < H; ++k) {
    float a[2], b[2], c[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int col = col0 + 2 * k + h;
      const float w = p.weight != nullptr ? IO<T>::to_f32(static_cast<const T*>(p.weight)[col]) : 1.f;
      if constexpr (!BWD) {
        const float mean = __ldcg(p.fin + col), invstd = __ldcg(p.fin + p.N + col);
        const float bi = p.bias != nullptr ? IO<T>::to_f32(static_cast<const T*>(p.bias)[col]) : 0.f;
        a[h] = invstd * w;                 // y = (x - mean) * invstd * w + b  =  x * a + (b - mean * a)
        b[h] = sizeof(T) == 2 ? fmaf(-mean, a[h], bi) : bi;
        c[h] = -mean;
      } else {
        const float gs = __ldcg(p.fin + col), dp = __ldcg(p.fin + p.N + col);
        const float invstd = p.invstd_in[col];
        // dx = (dy - (x - mean) * dp * invstd^2 / M - gs / M) * w * invstd
        a[h] = w * invstd;
        c[h] = -(dp * p.inv_m) * (invstd * invstd);
        b[h] = -gs * p.inv_m;
      }
    }
    sc[k] = make_float2(a[0], a[1]);
    sf[k] = make_float2(b[0], b[1]);
    pj[k] = make_float2(c[0], c[1]);
  }
  float2 nmean[H];
#pragma unroll
  for (int k = 0; k < H; ++k) {
    nmean[k] = BWD ? make_float2(-p.mean_in[col0 + 2 * k], -p.mean_in[col0 + 2 * k + 1]) : f2(0.f);
  }
  uint8_t* o = static_cast<uint8_t*>(p.out) + (size_t)cchunk * 16u;
#pragma unroll (BWD ? 4 : 8)
  for (int64_t r = r0 + rsub; r < r1; r += p.rpi) {
    float2 y[H];
    if constexpr (!BWD) {
      float2 x[H];
      IO<T>::unpack2(ldg_128_nc(a0 + (size_t)r * row_bytes), x);
#pragma unroll
      for (int k = 0; k < H; ++k) {
        // 16-bit io: one FMA on x (the folded shift costs at most |mean| / std fp32 ulps, far below the output's
        // resolution); fp32 io keeps subtract-then-scale
        if constexpr (sizeof(T) == 2) {
          y[k] = f2fma(x[k], sc[k], sf[k]);
        } else {
          y[k] = f2fma(f2add(x[k], pj[k]), sc[k], sf[k]);  // pj = -mean, sf = bias
        }
      }
    } else {
      float2 dy[H], x[H];
      IO<T>::unpack2(ldg_128_nc(a0 + (size_t)r * row_bytes), dy);
      IO<T>::unpack2(ldg_128_nc(a1 + (size_t)r * row_bytes), x);
#pragma unroll
      for (int k = 0; k < H; ++k) y[k] = f2mul(f2add(f2fma(pj[k], f2add(x[k], nmean[k]), dy[k]), sf[k]), sc[k]);
    }
    stg_128_stream(o + (size_t)r * row_bytes, IO<T>::pack2(y));
  }
}

}  // namespace b200norm

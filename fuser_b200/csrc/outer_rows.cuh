// outer_rows.cuh — the outer (dim 0) normalization for FEW columns and MANY rows: the batch-norm shapes of
// convolutional nets viewed as [N*H*W, C] (e.g. [100352, 256], [802816, 64]; benchmarks/python/normalization.py,
// test_batchnorm_{fwd,bwd}.py with channels last), which no cluster of CTAs can hold in shared memory.
//
// The reference runs these as an OuterPersistent segment while the batch fits its grid-persistent registers and
// segments them into a Reduction + pointwise pair beyond (scheduler/normalization_outer.cpp:451-683,
// runtime/fused_welford_impl_outer.cu). Here ONE persistent cooperative kernel does both passes; every CTA owns a
// contiguous block of rows — with contiguous rows that is one contiguous byte range, cut into tiles of
// kOuterRowsU * active * 16 bytes that a producer warp streams through a ring of shared-memory slots with 1-D TMA
// bulk copies (tile t always lives in slot t mod depth):
//   phase 1  tiles first to last. A row is only nchunks = N*sizeof(T)/16 chunks wide, so consumer thread t owns chunk
//            (t mod nchunks) of every (active / nchunks)-th row — element (u * active + t) of a tile, conflict-free
//            16-byte shared loads — and accumulates its 8 (4) columns in registers (forward: shifted sums -> Welford
//            pairs; backward: sum dy and sum dy (x - mean)). The threads of a column are merged in shared memory in a
//            fixed order and the CTA writes its per-column partials. The LAST depth tiles are not overwritten;
//   barrier  in-kernel grid barrier (cooperative launch); a warp per column merges the CTAs' partials of its share of
//            the columns (Chan's formula with counts / plain sums, fixed order: deterministic) and writes the
//            per-column results (mean, invstd | dweight, dbias and what dx needs); second grid barrier;
//   phase 2  tiles last to first: the first depth of them are still in shared memory (up to ~190 KB of the CTA's
//            block — more than half of it for the [100352, 256] bf16 case), the others come back from L2 when the
//            matrix fits and from HBM otherwise (most recently read first); per-column affine form, streamed stores.
// HBM traffic is the algorithmic minimum while the matrix stays L2 resident and at most 1.5x beyond; the
// shape-agnostic three-kernel family (outer_generic.cuh: one column per thread, 2-byte loads, one thread merging
// thousands of splits) reached 4-14 % of the measured peak on these shapes.
// Requirements (outer_launch checks them): rows contiguous (stride = N), 16-byte aligned operands, N * sizeof(T) a
// multiple of 16, nchunks <= kOuterRowsThreads.
#pragma once

#include "common.cuh"

namespace b200norm {

constexpr int kOuterRowsThreads = 512;              // consumer threads; one more warp produces
constexpr int kOuterRowsBlock = kOuterRowsThreads + 32;
constexpr int kOuterRowsU = 4;                      // 16-byte elements per thread and tile
constexpr int kOuterRowsMaxDepth = 8;
constexpr int kOuterRowsBarBytes = 128;             // 2 * depth mbarriers at the start of the dynamic shared memory

struct OuterRowsParams {
  const void* in0;  // fwd: x ; bwd: dy
  const void* in1;  // bwd: x
  void* out;        // fwd: y ; bwd: dx (may be null)
  int64_t M;
  int32_t N;
  int32_t nchunks;       // 16-byte chunks per row
  int32_t rpi;           // rows covered by one pass of the consumer threads = active / nchunks
  int32_t active;        // consumer threads that own a chunk = rpi * nchunks
  int64_t rows_per_cta;
  int32_t depth;         // ring slots
  uint32_t tile_bytes;   // kOuterRowsU * active * 16
  uint32_t ring_off;     // byte offset of the ring in the dynamic shared memory (after barriers and merge staging)
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
  unsigned long long* timeline;  // tuning aid (b200norm_debug_set_timeline): 12 slots per CTA, marks as in ring.cuh
};

// marks: 0 entry, 1 setup, 2 first tile landed, 3 phase 1 done, 4 partial written, 5 first grid barrier passed,
// 8 columns merged, 9 second grid barrier passed, 6 exit (slot 7: the SM)
__device__ __forceinline__ void orows_mark(const OuterRowsParams& p, int slot) {
  if (p.timeline != nullptr && threadIdx.x == 0) {
    p.timeline[(size_t)blockIdx.x * 12 + slot] = global_timer_ns();
    if (slot == 0) {
      uint32_t smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      p.timeline[(size_t)blockIdx.x * 12 + 7] = smid;
    }
  }
}

__device__ __forceinline__ uint64_t l2_policy_evict_normal() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

// Sense-reversing grid barrier (the protocol of grid_barrier in kernels_bwd.cuh). The CTA's global writes are ordered
// before thread 0's acq_rel atomic by the CTA barrier (causality order is cumulative), and the spinning thread's
// acquire load orders the reads of every thread behind the closing CTA barrier.
__device__ __forceinline__ void outer_rows_grid_barrier(uint32_t* word) {
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
__global__ void __launch_bounds__(kOuterRowsBlock, 1) outer_rows_kernel(const OuterRowsParams p) {
  constexpr int VEC = IO<T>::VEC;
  constexpr int H = VEC / 2;
  constexpr int U = kOuterRowsU;
  constexpr int kWarps = kOuterRowsThreads / 32;
  extern __shared__ __align__(128) uint8_t dyn_smem[];
  const uint32_t smem0 = smem_u32(dyn_smem);
  const uint32_t bar_full = smem0, bar_empty = smem0 + 8u * (uint32_t)p.depth;
  float* sm = reinterpret_cast<float*>(dyn_smem + kOuterRowsBarBytes);  // [rpi][nchunks][VEC][2] merge staging
  const uint32_t ring0 = smem0 + p.ring_off;
  const uint32_t slot_bytes = (BWD ? 2u : 1u) * p.tile_bytes;
  const int tid = threadIdx.x, lane = tid & 31;
  const bool producer = tid >= kOuterRowsThreads;
  const bool owner = tid < p.active;
  const int cchunk = owner ? tid % p.nchunks : 0;
  const int col0 = cchunk * VEC;
  const int depth = p.depth;

  const int64_t r0 = (int64_t)blockIdx.x * p.rows_per_cta;
  const int64_t r1 = min(p.M, r0 + p.rows_per_cta);
  const int64_t nrows = r1 > r0 ? r1 - r0 : 0;
  const size_t row_bytes = (size_t)p.nchunks * 16u;
  const size_t blk_off = (size_t)r0 * row_bytes, nb = (size_t)nrows * row_bytes;
  const int nt = (int)((nb + p.tile_bytes - 1) / p.tile_bytes);
  const uint32_t last_bytes = nt > 0 ? (uint32_t)(nb - (size_t)(nt - 1) * p.tile_bytes) : 0u;

  orows_mark(p, 0);
  if (tid == 0) {
    for (int s = 0; s < depth; ++s) {
      mbar_init(bar_full + 8u * (uint32_t)s, 1);
      mbar_init(bar_empty + 8u * (uint32_t)s, kWarps);
    }
    fence_mbar_init();
  }
  __syncthreads();
  pdl_wait();  // programmatic dependent launch: the predecessor's results (and its use of the workspace) are complete
  orows_mark(p, 1);

  // producer side (lane 0 of the last warp): tile t -> slot t % depth
  uint32_t empty_par = 0;
  auto issue = [&](int t, uint64_t policy) {
    const int s = t % depth;
    const size_t off = (size_t)t * p.tile_bytes;
    const uint32_t bytes = t == nt - 1 ? last_bytes : p.tile_bytes;
    const uint32_t dst = ring0 + (uint32_t)s * slot_bytes, bar = bar_full + 8u * (uint32_t)s;
    mbar_arrive_expect_tx(bar, BWD ? 2u * bytes : bytes);
    bulk_g2s(dst, static_cast<const uint8_t*>(p.in0) + blk_off + off, bytes, bar, policy);
    if constexpr (BWD) bulk_g2s(dst + p.tile_bytes, static_cast<const uint8_t*>(p.in1) + blk_off + off, bytes, bar, policy);
  };
  auto wait_empty = [&](int s) {
    mbar_wait(bar_empty + 8u * (uint32_t)s, (empty_par >> s) & 1u);
    empty_par ^= 1u << s;
  };
  // consumer side
  uint32_t full_par = 0;
  auto wait_full = [&](int s) {
    mbar_wait(bar_full + 8u * (uint32_t)s, (full_par >> s) & 1u);
    full_par ^= 1u << s;
  };
  auto release = [&](int s) {
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty + 8u * (uint32_t)s);
  };

  // ------------------------------------------------------------------ phase 1: per-column partials of the row block
  float v0[VEC], v1[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) v0[e] = v1[e] = 0.f;
  if (producer) {
    if (lane == 0) {
      const uint64_t pol = l2_policy_evict_normal();
      for (int t = 0; t < nt; ++t) {
        if (t >= depth) wait_empty(t % depth);
        issue(t, pol);
      }
    }
    __syncwarp();
  } else {
    float sh[VEC];
    float cnt = 0.f;
#pragma unroll
    for (int e = 0; e < VEC; ++e) sh[e] = 0.f;
    float2 nmean[H];
#pragma unroll
    for (int k = 0; k < H; ++k) nmean[k] = BWD ? make_float2(-p.mean_in[col0 + 2 * k], -p.mean_in[col0 + 2 * k + 1]) : f2(0.f);
    for (int t = 0; t < nt; ++t) {
      const int s = t % depth;
      wait_full(s);
      if (t == 0) orows_mark(p, 2);
      const uint32_t bytes = t == nt - 1 ? last_bytes : p.tile_bytes;
      const uint32_t base = ring0 + (uint32_t)s * slot_bytes + 16u * (uint32_t)tid;
      if (owner) {
        if constexpr (!BWD) {
          if (t == 0 && 16u * (uint32_t)tid < bytes) {
            // shift = the thread's first element: the cancellation in M2 = S2 - S1^2 / n is bounded by the rows a
            // thread sees
            float2 x[H];
            IO<T>::unpack2(lds_128(base), x);
#pragma unroll
            for (int k = 0; k < H; ++k) {
              sh[2 * k] = x[k].x;
              sh[2 * k + 1] = x[k].y;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const uint32_t eoff = 16u * (uint32_t)(u * p.active);
          if (eoff + 16u * (uint32_t)tid < bytes) {
            if constexpr (!BWD) {
              float2 x[H];
              IO<T>::unpack2(lds_128(base + eoff), x);
#pragma unroll
              for (int k = 0; k < H; ++k) {
                const float d0 = x[k].x - sh[2 * k], d1 = x[k].y - sh[2 * k + 1];
                v0[2 * k] += d0;
                v0[2 * k + 1] += d1;
                v1[2 * k] = fmaf(d0, d0, v1[2 * k]);
                v1[2 * k + 1] = fmaf(d1, d1, v1[2 * k + 1]);
              }
              cnt += 1.f;
            } else {
              float2 dy[H], x[H];
              IO<T>::unpack2(lds_128(base + eoff), dy);
              IO<T>::unpack2(lds_128(base + eoff + p.tile_bytes), x);
#pragma unroll
              for (int k = 0; k < H; ++k) {
                const float2 xc = f2add(x[k], nmean[k]);
                v0[2 * k] += dy[k].x;
                v0[2 * k + 1] += dy[k].y;
                v1[2 * k] = fmaf(dy[k].x, xc.x, v1[2 * k]);
                v1[2 * k + 1] = fmaf(dy[k].y, xc.y, v1[2 * k + 1]);
              }
            }
          }
        }
      }
      if (t + depth < nt) release(s);  // the last depth tiles stay where they are for phase 2
    }
    orows_mark(p, 3);
    if (owner) {
      if constexpr (!BWD) {
        // shifted sums -> (mean, M2); the count is a function of the thread's row subset (recomputed by the merge)
        const float inv = cnt > 0.f ? 1.f / cnt : 0.f;
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const float s1 = v0[e];
          v0[e] = fmaf(s1, inv, sh[e]);
          v1[e] = fmaxf(fmaf(-s1 * inv, s1, v1[e]), 0.f);
        }
      }
      float2* dst = reinterpret_cast<float2*>(sm) + (size_t)tid * VEC;  // tid = rsub * nchunks + cchunk
#pragma unroll
      for (int e = 0; e < VEC; ++e) dst[e] = make_float2(v0[e], v1[e]);
    }
  }
  __syncthreads();
  // g threads per COLUMN (a power of two inside one warp) merge the rpi row subsets — each every g-th subset in
  // order, then a shuffle tree: the order depends on N and rpi only — and write the CTA's partial. Row subset s of a
  // block of nrows rows holds nrows / rpi rows, one more for s < nrows % rpi.
  {
    float* my_part = p.part + (size_t)blockIdx.x * 3 * p.N;
    const float cnt_lo = (float)(nrows / p.rpi);
    const int extra = (int)(nrows % p.rpi);
    int g = 1;
    while (g < 32 && 2 * g * p.N <= kOuterRowsThreads && 2 * g <= p.rpi) g *= 2;
    const int cols_per_pass = kOuterRowsThreads / g;
    const int j = tid % g;
    const size_t stride = (size_t)p.nchunks * VEC;
    if (tid < kOuterRowsThreads) {
      for (int base = 0; base < p.N; base += cols_per_pass) {
        const int c = base + tid / g;
        float n = 0.f, m = 0.f, q = 0.f;
        if (c < p.N) {
          const float2* src = reinterpret_cast<const float2*>(sm) + c;  // element e of chunk cc: cc * VEC + e = c
          for (int s = j; s < p.rpi; s += g) {
            const float2 t = src[s * stride];
            if constexpr (!BWD) {
              chan_merge3(n, m, q, cnt_lo + (s < extra ? 1.f : 0.f), t.x, t.y);
            } else {
              n += t.x;
              m += t.y;
            }
          }
        }
        for (int off = g >> 1; off > 0; off >>= 1) {
          const float tn = __shfl_down_sync(0xffffffffu, n, off), tm = __shfl_down_sync(0xffffffffu, m, off);
          if constexpr (!BWD) {
            const float tq = __shfl_down_sync(0xffffffffu, q, off);
            chan_merge3(n, m, q, tn, tm, tq);
          } else {
            n += tn;
            m += tm;
          }
        }
        if (j == 0 && c < p.N) {
          my_part[c] = n;
          my_part[p.N + c] = m;
          if constexpr (!BWD) my_part[2 * (size_t)p.N + c] = q;
        }
      }
    }
  }
  orows_mark(p, 4);
  outer_rows_grid_barrier(p.sync_word);
  orows_mark(p, 5);

  // ------------------------------------------------------------------ merge over the CTAs, in a fixed order
  {
    const int per_cta = (p.N + gridDim.x - 1) / gridDim.x;
    const int c_begin = blockIdx.x * per_cta, c_end = min(p.N, c_begin + per_cta);
    // a warp per column: lane l merges CTAs l, l + 32, ... (independent loads, all in flight together), then a
    // fixed shuffle tree — the order of the merges depends on nothing but the grid size
    const int warp = tid >> 5;
    for (int c = c_begin + warp; c < c_end; c += kOuterRowsBlock / 32) {
      float n = 0.f, m = 0.f, q = 0.f;
      for (unsigned b0 = 0; b0 < gridDim.x; b0 += 160) {
        // all the loads first (one L2 round trip), then the merges
        float tn[5], tm[5], tq[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          const unsigned bb = b0 + (unsigned)lane + 32u * (unsigned)i;
          const bool ok = bb < gridDim.x;
          const float* t = p.part + (size_t)(ok ? bb : 0u) * 3 * p.N + c;
          tn[i] = ok ? __ldcg(t) : 0.f;
          tm[i] = ok ? __ldcg(t + p.N) : 0.f;
          tq[i] = (ok && !BWD) ? __ldcg(t + 2 * (size_t)p.N) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          if constexpr (!BWD) {
            chan_merge3(n, m, q, tn[i], tm[i], tq[i]);
          } else {
            n += tn[i];
            m += tm[i];
          }
        }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const float tn = __shfl_down_sync(0xffffffffu, n, off), tm = __shfl_down_sync(0xffffffffu, m, off);
        if constexpr (!BWD) {
          const float tq = __shfl_down_sync(0xffffffffu, q, off);
          chan_merge3(n, m, q, tn, tm, tq);
        } else {
          n += tn;
          m += tm;
        }
      }
      if (lane != 0) continue;
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
  orows_mark(p, 8);
  if (p.out == nullptr) return;  // backward without dx: nothing left (every CTA takes this branch together)
  outer_rows_grid_barrier(p.sync_word);
  orows_mark(p, 9);
  pdl_launch_dependents();  // the successor's CTAs may take the SMs as ours exit

  // ------------------------------------------------------------------ phase 2: apply, last tile first
  if (producer) {
    if (lane == 0) {
      const uint64_t pol = l2_policy_evict_first();
      for (int t = nt - 1 - depth; t >= 0; --t) {
        wait_empty(t % depth);
        issue(t, pol);
      }
    }
    return;
  }
  // per-column coefficients: computed once per CTA by the threads of row subset 0 and handed to the others through
  // shared memory (every CTA reads the same few KB of results right after the barrier: one request per column and
  // CTA instead of one per thread keeps that from queueing on a handful of L2 lines)
  float2 sc[H], sf[H], pj[H], nmean[H];
  const bool bcast = p.rpi > 1;  // the merge staging holds rpi * N * 8 bytes: room for the 4 * N coefficients
  float* coef = sm;
  if (!bcast || tid < p.nchunks) {
#pragma unroll
    for (int k = 0; k < H; ++k) {
      float a[2], b[2], c[2], d[2];
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
          d[h] = 0.f;
        } else {
          const float gs = __ldcg(p.fin + col), dp = __ldcg(p.fin + p.N + col);
          const float invstd = p.invstd_in[col];
          // dx = (dy - (x - mean) * dp * invstd^2 / M - gs / M) * w * invstd
          a[h] = w * invstd;
          c[h] = -(dp * p.inv_m) * (invstd * invstd);
          b[h] = -gs * p.inv_m;
          d[h] = -p.mean_in[col];
        }
      }
      sc[k] = make_float2(a[0], a[1]);
      sf[k] = make_float2(b[0], b[1]);
      pj[k] = make_float2(c[0], c[1]);
      nmean[k] = make_float2(d[0], d[1]);
      if (bcast) {
        *reinterpret_cast<float2*>(coef + col0 + 2 * k) = sc[k];
        *reinterpret_cast<float2*>(coef + p.N + col0 + 2 * k) = sf[k];
        *reinterpret_cast<float2*>(coef + 2 * p.N + col0 + 2 * k) = pj[k];
        *reinterpret_cast<float2*>(coef + 3 * p.N + col0 + 2 * k) = nmean[k];
      }
    }
  }
  if (bcast) {
    named_bar_sync(1, kOuterRowsThreads);  // the consumer warps only: the producer is in its own loop
#pragma unroll
    for (int k = 0; k < H; ++k) {
      sc[k] = *reinterpret_cast<const float2*>(coef + col0 + 2 * k);
      sf[k] = *reinterpret_cast<const float2*>(coef + p.N + col0 + 2 * k);
      pj[k] = *reinterpret_cast<const float2*>(coef + 2 * p.N + col0 + 2 * k);
      nmean[k] = *reinterpret_cast<const float2*>(coef + 3 * p.N + col0 + 2 * k);
    }
  }
  uint8_t* o = static_cast<uint8_t*>(p.out) + blk_off + 16u * (size_t)tid;
  for (int t = nt - 1; t >= 0; --t) {
    const int s = t % depth;
    if (t < nt - depth) wait_full(s);  // the others never left shared memory
    const uint32_t bytes = t == nt - 1 ? last_bytes : p.tile_bytes;
    const uint32_t base = ring0 + (uint32_t)s * slot_bytes + 16u * (uint32_t)tid;
    if (owner) {
      uint8_t* orow = o + (size_t)t * p.tile_bytes;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t eoff = 16u * (uint32_t)(u * p.active);
        if (eoff + 16u * (uint32_t)tid < bytes) {
          float2 y[H];
          if constexpr (!BWD) {
            float2 x[H];
            IO<T>::unpack2(lds_128(base + eoff), x);
#pragma unroll
            for (int k = 0; k < H; ++k) {
              // 16-bit io: one FMA on x (the folded shift costs at most |mean| / std fp32 ulps, far below the
              // output's resolution); fp32 io keeps subtract-then-scale
              if constexpr (sizeof(T) == 2) {
                y[k] = f2fma(x[k], sc[k], sf[k]);
              } else {
                y[k] = f2fma(f2add(x[k], pj[k]), sc[k], sf[k]);  // pj = -mean, sf = bias
              }
            }
          } else {
            float2 dy[H], x[H];
            IO<T>::unpack2(lds_128(base + eoff), dy);
            IO<T>::unpack2(lds_128(base + eoff + p.tile_bytes), x);
#pragma unroll
            for (int k = 0; k < H; ++k) y[k] = f2mul(f2add(f2fma(pj[k], f2add(x[k], nmean[k]), dy[k]), sf[k]), sc[k]);
          }
          stg_128_stream(orow + eoff, IO<T>::pack2(y));
        }
      }
    }
    if (t >= depth) release(s);  // tile t - depth goes into this slot
  }
  orows_mark(p, 6);
}

}  // namespace b200norm

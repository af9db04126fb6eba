// outer.cu — outer-persistent softmax / log-softmax, forward and backward (sm_100a).
//
// Reduction over dim 0 of a row-major [M, N] matrix: every column is an independent problem, but a
// column is strided in memory, so the unit of work is a *slab*: W bytes (64-512) of every row. The
// reference keeps such a slab persistent in registers across a cooperative grid and reduces through
// global memory (scheduler/normalization_outer.cpp:451-683, runtime/grid_reduction.cu:905-1066,
// runtime/fused_welford_impl_outer.cu). Here:
//
//   * a slab is held in SHARED memory — of one CTA when M*W fits, else of a thread-block cluster
//     whose CTAs split the rows (rank q holds rows [q*Rc, (q+1)*Rc)) and exchange their per-column
//     partial results through DSMEM (st.async completing on the peer's mbarrier), never through HBM;
//   * a slab tile is fetched with 2-D TMA box loads (cp.async.bulk.tensor.2d, <= 256 rows per box)
//     that complete on an mbarrier; two tiles per CTA: the next slab streams in while the current
//     one is reduced and written out. Out-of-range rows / columns are zero-filled by the TMA unit
//     and masked by index;
//   * the grid is persistent (clusters walk the slabs round-robin) and sized from the occupancy
//     API: small tiles run two or three CTAs per SM so that one CTA's barrier latency hides behind
//     the others' passes;
//   * a thread owns one 16-byte column chunk and every (256 / LC)-th row of it: consecutive threads
//     read consecutive 16-byte words of the tile (conflict-free), per-column partials are combined
//     with xor-shuffles inside a warp, through shared memory across the 8 warps, and across the
//     cluster in rank order — a fixed order, so results are deterministic;
//   * arithmetic is fp32 with the same element formulas as the inner kernels (kernels_fwd.cuh):
//     NaN-propagating max, exp(x - max), e * rcp(sum) (softmax) or (x - max) - log(sum) (log-softmax);
//     backward dx = y*dy - y*sum(y*dy) or dy - exp(y)*sum(dy) (csrc/ops/normalization.cpp:169-249).
//     HBM traffic is the algorithmic minimum: every input element is read once, every output written once.
#include "outer.cuh"
#include "outer_rows.cuh"

#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>
#include <unordered_map>

#include "../../include/b200norm.h"
#include "common.cuh"
#include "outer_generic.cuh"

namespace b200norm {

extern __shared__ __align__(128) uint8_t outer_smem[];

constexpr int kOuterThreads = 256;
constexpr int kOuterWarps = kOuterThreads / 32;

struct OuterParams {
  void* out;
  int64_t out_stride;  // elements
  int32_t M, N;
  int32_t rows_cta;    // Rc: rows of a slab held by one CTA (nboxes * box_rows)
  int32_t box_rows, nboxes;
  int32_t slab_cols;   // SC = W / sizeof(T)
  int32_t lc_shift;    // log2(LC), LC = W / 16 chunks per slab row
  int32_t nslabs;
  int32_t cluster;     // C
  int32_t depth;       // 1 or 2 tiles
  int32_t round_sum;
  // outer normalization (statistics over dim 0, affine per column)
  const void* weight;   // [N] io dtype or nullptr
  const void* bias;     // [N] io dtype or nullptr
  float* mean_out;      // [N] fp32 or nullptr
  float* invstd_out;    // [N] fp32 or nullptr
  const float* mean_in;    // normalization backward: saved statistics [N] fp32
  const float* invstd_in;
  void* dw_out;            // [N] io dtype or nullptr
  void* db_out;            // [N] io dtype or nullptr
  float inv_m;          // 1 / M
  float eps;
  int32_t two_exchanges;      // tuning (B200NORM_OUTER_TWO_EXCH=1): forward exchanges max and sum separately
  int32_t l2_prefetch;        // 1: prefetch the next slab's boxes into L2 at the start of an iteration
  int32_t plain_stores;       // 1: default-policy st.global instead of the streaming (evict-first) form
  int32_t debug_skip_reduce;  // tuning aid (B200NORM_OUTER_SKIP_REDUCE): wrong results, timing only
  uint32_t slab_bytes; // W
  uint32_t tile_bytes; // Rc * W (one stream)
  uint32_t stage_bytes;  // nstreams * tile_bytes
  uint32_t off_wpart, off_cls, off_fin, off_tiles;
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int32_t c0, int32_t c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
// L2 prefetch of a box (no shared memory, no completion to wait for)
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float4 lds_f4(const float* p) { return *reinterpret_cast<const float4*>(p); }
// Remote (or local) shared store of four floats completing 16 tx-bytes on the destination CTA's mbarrier.
__device__ __forceinline__ void st_async_f32x4(uint32_t remote_addr, const float4& v, uint32_t remote_bar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(remote_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(remote_bar)
               : "memory");
}

// packed 16-bit NaN-propagating max (exact: no rounding is involved), one instruction per two elements
template <typename T>
struct Max2;
template <>
struct Max2<__nv_bfloat16> {
  static constexpr uint32_t kNegInf = 0xff80ff80u;
  __device__ __forceinline__ static uint32_t apply(uint32_t a, uint32_t b) {
    uint32_t r;
    asm("max.NaN.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
  }
};
template <>
struct Max2<__half> {
  static constexpr uint32_t kNegInf = 0xfc00fc00u;
  __device__ __forceinline__ static uint32_t apply(uint32_t a, uint32_t b) {
    uint32_t r;
    asm("max.NaN.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
  }
};

constexpr float kLog2e = 1.4426950408889634f;
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// exp(x - m) for a pair: 16-bit io -> ex2(x*log2e - m*log2e), one FFMA2 + two MUFU (nml = -m*log2e);
// fp32 io -> accurate expf(x - m) (nm = -m)
template <typename T>
__device__ __forceinline__ float2 exp_shift(float2 x, float2 nm, float2 nml) {
  if constexpr (sizeof(T) == 4) {
    const float2 d = f2add(x, nm);
    return make_float2(expf(d.x), expf(d.y));
  } else {
    const float2 t = f2fma(x, f2(kLog2e), nml);
    return make_float2(ex2_fast(t.x), ex2_fast(t.y));
  }
}

// Column reduction context: combine the per-thread partials v[VEC] (thread's 16-byte column chunk)
// over every row of the slab: lanes of a warp that share the chunk, the CTA's warps, the cluster's CTAs.
// The cluster step is an all-to-all of SC floats per CTA with st.async: every value lands in the
// peer's slot set and completes 4 tx-bytes on the PEER's mbarrier, which the peer armed with
// expect_tx(C * SC * 4). No barrier.cluster and no fence in the loop: a release fence here would also
// wait for the CTA's outstanding global stores of the previous slab (measured: MEMBAR was the top
// stall, and a 16-CTA cluster ran 4x slower than one CTA).
struct ColCtx {
  uint32_t smem0;
  int tid, lane, warp, c, LC, SC;
  uint32_t C, crank;
  uint32_t nx;       // cluster exchanges done so far: set = nx & 1, parity = (nx >> 1) & 1
  uint32_t xbar;     // mbarrier pair of the exchange (shared::cta address)
  float* wpart;      // [kOuterWarps][SC]
  float* cls;        // [2][C][SC]
  float* fin;        // [SC]
  uint32_t off_cls;
};

template <typename Op, int VEC, bool kCluster = true>
__device__ __forceinline__ void col_reduce(float (&v)[VEC], ColCtx& x) {
  for (int off = 16; off >= x.LC; off >>= 1) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) v[k] = Op::apply(v[k], __shfl_xor_sync(0xffffffffu, v[k], off));
  }
  if (x.lane < x.LC) {  // lane == chunk index for these lanes
    float* dst = x.wpart + x.warp * x.SC + x.c * VEC;
#pragma unroll
    for (int k = 0; k < VEC; k += 4) *reinterpret_cast<float4*>(dst + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
  }
  __syncthreads();
  // four columns per thread: 16-byte shared loads and 16-byte DSMEM stores (a quarter of the st.async
  // transactions of a per-column exchange; with one float per store a 4 KB exchange — W = 512 over 4 CTAs,
  // W = 256 over 8 — ran 30 % slower than a 1 KB one)
  if (x.tid < (x.SC >> 2)) {
    const int c4 = x.tid << 2;
    float4 a = lds_f4(x.wpart + c4);
#pragma unroll
    for (int w = 1; w < kOuterWarps; ++w) {
      const float4 b = lds_f4(x.wpart + w * x.SC + c4);
      a.x = Op::apply(a.x, b.x);
      a.y = Op::apply(a.y, b.y);
      a.z = Op::apply(a.z, b.z);
      a.w = Op::apply(a.w, b.w);
    }
    if (kCluster && x.C > 1) {
      const uint32_t set = x.nx & 1u, parity = (x.nx >> 1) & 1u;
      const uint32_t bar = x.xbar + 8u * set;
      if (x.tid == 0) mbar_arrive_expect_tx(bar, x.C * (uint32_t)x.SC * 4u);
      const uint32_t local = x.smem0 + x.off_cls + ((set * x.C + x.crank) * (uint32_t)x.SC + (uint32_t)c4) * 4u;
      for (uint32_t q = 0; q < x.C; ++q) st_async_f32x4(map_to_cta(local, q), a, map_to_cta(bar, q));
      mbar_wait(bar, parity);
      const volatile float4* s = reinterpret_cast<const volatile float4*>(x.cls + (size_t)set * x.C * x.SC + c4);
      const int stride4 = x.SC >> 2;
      a.x = s[0].x;
      a.y = s[0].y;
      a.z = s[0].z;
      a.w = s[0].w;
      for (uint32_t q = 1; q < x.C; ++q) {
        a.x = Op::apply(a.x, s[q * stride4].x);
        a.y = Op::apply(a.y, s[q * stride4].y);
        a.z = Op::apply(a.z, s[q * stride4].z);
        a.w = Op::apply(a.w, s[q * stride4].w);
      }
    }
    *reinterpret_cast<float4*>(x.fin + c4) = a;
  }
  if (kCluster && x.C > 1) ++x.nx;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < VEC; k += 4) {
    const float4 t = lds_f4(x.fin + x.c * VEC + k);
    v[k] = t.x;
    v[k + 1] = t.y;
    v[k + 2] = t.z;
    v[k + 3] = t.w;
  }
}

// Two sums per column in one go (the normalization backward's grad_sum and dot_p): one pair of local barriers and
// ONE cluster exchange instead of two. Planes: wpart [warps][2][SC], slots [set][rank][2][SC], fin planes 0 and 1.
template <int VEC>
__device__ __forceinline__ void col_reduce_sum2(float (&u)[VEC], float (&v)[VEC], ColCtx& x) {
  for (int off = 16; off >= x.LC; off >>= 1) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      u[k] += __shfl_xor_sync(0xffffffffu, u[k], off);
      v[k] += __shfl_xor_sync(0xffffffffu, v[k], off);
    }
  }
  if (x.lane < x.LC) {
    float* dst = x.wpart + x.warp * 2 * x.SC + x.c * VEC;
#pragma unroll
    for (int k = 0; k < VEC; k += 4) {
      *reinterpret_cast<float4*>(dst + k) = make_float4(u[k], u[k + 1], u[k + 2], u[k + 3]);
      *reinterpret_cast<float4*>(dst + x.SC + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
    }
  }
  __syncthreads();
  if (x.tid < (x.SC >> 1)) {  // SC / 4 threads per plane
    const int plane = x.tid >= (x.SC >> 2) ? 1 : 0;
    const int c4 = (x.tid - plane * (x.SC >> 2)) << 2;
    const float* src = x.wpart + plane * x.SC + c4;
    float4 a = lds_f4(src);
#pragma unroll
    for (int w = 1; w < kOuterWarps; ++w) {
      const float4 b = lds_f4(src + w * 2 * x.SC);
      a.x += b.x;
      a.y += b.y;
      a.z += b.z;
      a.w += b.w;
    }
    if (x.C > 1) {
      const uint32_t set = x.nx & 1u, parity = (x.nx >> 1) & 1u;
      const uint32_t bar = x.xbar + 8u * set;
      if (x.tid == 0) mbar_arrive_expect_tx(bar, x.C * (uint32_t)x.SC * 8u);
      const uint32_t local =
          x.smem0 + x.off_cls + ((((set * x.C + x.crank) * 2u) + (uint32_t)plane) * (uint32_t)x.SC + (uint32_t)c4) * 4u;
      for (uint32_t q = 0; q < x.C; ++q) st_async_f32x4(map_to_cta(local, q), a, map_to_cta(bar, q));
      mbar_wait(bar, parity);
      const volatile float* base = x.cls + ((size_t)set * x.C * 2 + plane) * x.SC + c4;
      a.x = base[0];
      a.y = base[1];
      a.z = base[2];
      a.w = base[3];
      for (uint32_t q = 1; q < x.C; ++q) {
        a.x += base[(size_t)q * 2 * x.SC];
        a.y += base[(size_t)q * 2 * x.SC + 1];
        a.z += base[(size_t)q * 2 * x.SC + 2];
        a.w += base[(size_t)q * 2 * x.SC + 3];
      }
    }
    *reinterpret_cast<float4*>(x.fin + plane * x.SC + c4) = a;
  }
  if (x.C > 1) ++x.nx;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < VEC; k += 4) {
    const float4 t = lds_f4(x.fin + x.c * VEC + k);
    const float4 w = lds_f4(x.fin + x.SC + x.c * VEC + k);
    u[k] = t.x;
    u[k + 1] = t.y;
    u[k + 2] = t.z;
    u[k + 3] = t.w;
    v[k] = w.x;
    v[k + 1] = w.y;
    v[k + 2] = w.z;
    v[k + 3] = w.w;
  }
}

// Softmax forward over a cluster: ONE exchange per slab instead of two. Every CTA reduces its own rows
// completely — max (local col_reduce, result in fin), then the sum of exp(x - CTA max) — and the cluster
// exchanges (max, sum) PAIRS: M = max_q m_q, S = sum_q s_q * exp(m_q - M). The cluster is a lock-step on
// its slowest tile load, so meeting once instead of twice is what this buys. A CTA whose rows are all
// -inf (or that holds no rows) contributes (m = -inf, s = 0): its sum was taken against 0 instead of
// -inf by the caller, and its factor is exp(-inf - M) = 0, or 1 when M is -inf too (then S = 0 and the
// last pass turns the column into NaN through x - M, as the op-by-op evaluation does). NaN and +inf
// propagate through max / expf.
// in: m = CTA max of the thread's columns (as left in fin by the local max reduce), s = thread's partial sums;
// out: m = cluster max, s = cluster sum relative to it.
__device__ __forceinline__ float pair_factor(float mq, float M) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"((mq - M) * 1.4426950408889634f));
  return mq == M ? 1.f : e;
}

template <int VEC>
__device__ __forceinline__ void col_reduce_softmax_pair(float (&m)[VEC], float (&v)[VEC], ColCtx& x) {
  for (int off = 16; off >= x.LC; off >>= 1) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) v[k] += __shfl_xor_sync(0xffffffffu, v[k], off);
  }
  if (x.lane < x.LC) {
    float* dst = x.wpart + x.warp * x.SC + x.c * VEC;
#pragma unroll
    for (int k = 0; k < VEC; k += 4) *reinterpret_cast<float4*>(dst + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
  }
  __syncthreads();  // also: every thread has read the CTA max out of fin
  if (x.tid < (x.SC >> 2)) {
    const int c4 = x.tid << 2;
    float4 a = lds_f4(x.wpart + c4);
#pragma unroll
    for (int w = 1; w < kOuterWarps; ++w) {
      const float4 b = lds_f4(x.wpart + w * x.SC + c4);
      a.x += b.x;
      a.y += b.y;
      a.z += b.z;
      a.w += b.w;
    }
    const float4 mc = lds_f4(x.fin + c4);
    const uint32_t set = x.nx & 1u;
    const uint32_t bar = x.xbar + 8u * set;
    if (x.tid == 0) mbar_arrive_expect_tx(bar, x.C * (uint32_t)x.SC * 8u);
    // slot layout of this exchange: [set][rank][max | sum][SC]
    const uint32_t local = x.smem0 + x.off_cls + (((set * x.C + x.crank) * 2u) * (uint32_t)x.SC + (uint32_t)c4) * 4u;
    for (uint32_t q = 0; q < x.C; ++q) {
      const uint32_t rbar = map_to_cta(bar, q);
      st_async_f32x4(map_to_cta(local, q), mc, rbar);
      st_async_f32x4(map_to_cta(local + (uint32_t)x.SC * 4u, q), a, rbar);
    }
  }
  // one column per thread for the combine (the factors cost an exp each: spread over SC threads, not SC / 4)
  if (x.tid < x.SC) {
    const uint32_t set = x.nx & 1u, parity = (x.nx >> 1) & 1u;
    mbar_wait(x.xbar + 8u * set, parity);
    const volatile float* base = x.cls + (size_t)set * x.C * 2 * x.SC + x.tid;
    float M = base[0];
    for (uint32_t q = 1; q < x.C; ++q) M = MaxOp::apply(M, base[(size_t)q * 2 * x.SC]);
    float S = 0.f;
    for (uint32_t q = 0; q < x.C; ++q) {
      const float mq = base[(size_t)q * 2 * x.SC];
      const float sq = base[(size_t)q * 2 * x.SC + x.SC];
      S = fmaf(sq, pair_factor(mq, M), S);
    }
    // results go to their own planes: plane 0 (the CTA max read by the senders above) is left alone
    x.fin[x.SC + x.tid] = M;
    x.fin[2 * x.SC + x.tid] = S;
  }
  ++x.nx;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < VEC; k += 4) {
    const float4 t = lds_f4(x.fin + x.SC + x.c * VEC + k);
    const float4 u = lds_f4(x.fin + 2 * x.SC + x.c * VEC + k);
    m[k] = t.x;
    m[k + 1] = t.y;
    m[k + 2] = t.z;
    m[k + 3] = t.w;
    v[k] = u.x;
    v[k + 1] = u.y;
    v[k + 2] = u.z;
    v[k + 3] = u.w;
  }
}

// KIND: 0 softmax fwd, 1 log-softmax fwd, 2 softmax bwd, 3 log-softmax bwd, 4 normalization fwd,
// 5 normalization bwd (in0 = dy, in1 = x; benchmarks/python/normalization.py:98-190: grad_sum = sum(dy),
// dot_p = sum(dy * (x - mean)) over dim 0; dx = (dy - dot_p/M * invstd^2 * (x - mean) - grad_sum/M) * w * invstd;
// dweight = dot_p * invstd; dbias = grad_sum)
// (mean / variance over dim 0, y = (x - mean) * invstd * weight[col] + bias[col]: the batch-norm-shaped
// segment of benchmarks/python/normalization.py:13-95 with channels last, statistics by the two-pass form the
// reference's persistent kernels use after TranslateApplicableWelford, fusion_segmenter.cpp:2341)
template <typename T, int KIND>
__global__ void __launch_bounds__(kOuterThreads, 3)
outer_softmax_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
                     const OuterParams p) {
  constexpr bool BWD = KIND == 2 || KIND == 3 || KIND == 5;  // two streamed inputs
  constexpr bool LOG = KIND == 1 || KIND == 3;
  constexpr bool NORM = KIND == 4;
  constexpr bool NORM_BWD = KIND == 5;
  constexpr int VEC = IO<T>::VEC;
  constexpr int H = VEC / 2;
  constexpr uint32_t kStep = kOuterThreads * 16;  // bytes between a thread's consecutive chunks
  const uint32_t smem0 = smem_u32(outer_smem);
  const int tid = threadIdx.x;

  ColCtx cx;
  cx.smem0 = smem0;
  cx.tid = tid;
  cx.lane = tid & 31;
  cx.warp = tid >> 5;
  cx.LC = 1 << p.lc_shift;
  cx.SC = p.slab_cols;
  cx.c = tid & (cx.LC - 1);
  cx.C = (uint32_t)p.cluster;
  cx.crank = cx.C > 1 ? cluster_ctarank() : 0u;
  cx.nx = 0;
  cx.xbar = smem0 + 16;
  cx.wpart = reinterpret_cast<float*>(outer_smem + p.off_wpart);
  cx.cls = reinterpret_cast<float*>(outer_smem + p.off_cls);
  cx.fin = reinterpret_cast<float*>(outer_smem + p.off_fin);
  cx.off_cls = p.off_cls;

  const int r0 = tid >> p.lc_shift;
  const int rstep = kOuterThreads >> p.lc_shift;
  const int cid = (int)(blockIdx.x / cx.C);
  const int nclusters = (int)(gridDim.x / cx.C);
  const int row_base = (int)cx.crank * p.rows_cta;
  int rows_valid = p.M - row_base;
  if (rows_valid > p.rows_cta) rows_valid = p.rows_cta;
  if (rows_valid < 0) rows_valid = 0;
  // chunks of this thread: rows r0 + j*rstep < rows_valid
  const int nj = rows_valid > r0 ? (rows_valid - r0 + rstep - 1) / rstep : 0;

  const uint32_t bar0 = smem0;  // full[2] | exchange[2]
  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    mbar_init(bar0 + 16, 1);
    mbar_init(bar0 + 24, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (cx.C > 1) cluster_sync_all();  // every CTA of the cluster is resident, its barriers initialised
  pdl_wait();

  auto issue = [&](int slab, uint32_t buf) {
    const uint32_t bar = bar0 + 8u * buf;
    mbar_arrive_expect_tx(bar, p.stage_bytes);
    uint32_t dst = smem0 + p.off_tiles + buf * p.stage_bytes;
    const int32_t col = slab * p.slab_cols;
    const uint32_t box_bytes = (uint32_t)p.box_rows * p.slab_bytes;
    for (int b = 0; b < p.nboxes; ++b) tma_load_2d(dst + (uint32_t)b * box_bytes, &map0, col, row_base + b * p.box_rows, bar);
    if constexpr (BWD) {
      dst += p.tile_bytes;
      for (int b = 0; b < p.nboxes; ++b) tma_load_2d(dst + (uint32_t)b * box_bytes, &map1, col, row_base + b * p.box_rows, bar);
    }
  };

  int slab = cid;
  if (slab < p.nslabs && tid == 0) issue(slab, 0);
  const size_t out_row_bytes = (size_t)p.out_stride * sizeof(T);
  const size_t out_step = out_row_bytes * (size_t)rstep;
  for (int it = 0; slab < p.nslabs; slab += nclusters, ++it) {
    const uint32_t buf = p.depth == 2 ? (uint32_t)(it & 1) : 0u;
    const uint32_t phase = p.depth == 2 ? (uint32_t)((it >> 1) & 1) : (uint32_t)(it & 1);
    const int next = slab + nclusters;
    // the other tile was released by the __syncthreads that ended the previous iteration
    if (p.depth == 2 && next < p.nslabs && tid == 0) issue(next, buf ^ 1u);
    // Experiment (B200NORM_OUTER_L2PF=1, off by default): with one tile per CTA the next slab cannot land in
    // shared memory before this one is done, but it can be pulled into L2 now (444 CTAs x 64 KB = 28 MB ahead).
    // Measured neutral (+-2 %) to -8 %: the tile wait is not what limits these kernels.
    if (p.depth == 1 && p.l2_prefetch && next < p.nslabs && tid == 0) {
      const int32_t ncol = next * p.slab_cols;
      for (int b = 0; b < p.nboxes; ++b) {
        tma_prefetch_2d(&map0, ncol, row_base + b * p.box_rows);
        if constexpr (BWD) tma_prefetch_2d(&map1, ncol, row_base + b * p.box_rows);
      }
    }
    const uint8_t* t0 = outer_smem + p.off_tiles + buf * p.stage_bytes + (uint32_t)tid * 16u;
    const int col = slab * p.slab_cols + cx.c * VEC;
    const bool col_ok = col < p.N;
    uint8_t* o = static_cast<uint8_t*>(p.out) + (size_t)(row_base + r0) * out_row_bytes + (size_t)col * sizeof(T);
    // per-column parameters of the normalization ops: requested BEFORE the wait for the tile, so that their
    // global-memory latency hides behind it (after the wait it sat on every slab's critical path)
    uint4 wq = make_uint4(0, 0, 0, 0), bq = wq;
    float4 mq[VEC / 4], iq[VEC / 4];
    if constexpr (NORM || NORM_BWD) {
      if (col_ok && p.weight != nullptr) wq = ldg_128_nc(static_cast<const uint8_t*>(p.weight) + (size_t)col * sizeof(T));
      if (NORM && col_ok && p.bias != nullptr) bq = ldg_128_nc(static_cast<const uint8_t*>(p.bias) + (size_t)col * sizeof(T));
    }
    if constexpr (NORM_BWD) {
#pragma unroll
      for (int k = 0; k < VEC / 4; ++k) {
        mq[k] = iq[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col_ok) {
          mq[k] = *reinterpret_cast<const float4*>(p.mean_in + col + 4 * k);
          iq[k] = *reinterpret_cast<const float4*>(p.invstd_in + col + 4 * k);
        }
      }
    }
    mbar_wait(bar0 + 8u * buf, phase);

    if constexpr (NORM_BWD) {
      float2 nmean[H];
      float istd[VEC];
      {
        float mv[VEC];
#pragma unroll
        for (int k = 0; k < VEC; k += 4) {
          const float4 a = mq[k / 4], b = iq[k / 4];
          mv[k] = a.x; mv[k + 1] = a.y; mv[k + 2] = a.z; mv[k + 3] = a.w;
          istd[k] = b.x; istd[k + 1] = b.y; istd[k + 2] = b.z; istd[k + 3] = b.w;
        }
#pragma unroll
        for (int k = 0; k < H; ++k) nmean[k] = make_float2(-mv[2 * k], -mv[2 * k + 1]);
      }
      // pass A: grad_sum and dot_p
      float gs[VEC], dp[VEC];
      {
        float2 g2[H], d2[H];
#pragma unroll
        for (int k = 0; k < H; ++k) g2[k] = d2[k] = f2(0.f);
        const uint8_t* t = t0;
#pragma unroll 4
        for (int j = 0; j < nj; ++j, t += kStep) {
          float2 dy[H], x[H];
          IO<T>::unpack2(*reinterpret_cast<const uint4*>(t), dy);
          IO<T>::unpack2(*reinterpret_cast<const uint4*>(t + p.tile_bytes), x);
#pragma unroll
          for (int k = 0; k < H; ++k) {
            g2[k] = f2add(g2[k], dy[k]);
            d2[k] = f2fma(dy[k], f2add(x[k], nmean[k]), d2[k]);
          }
        }
#pragma unroll
        for (int k = 0; k < H; ++k) {
          gs[2 * k] = g2[k].x; gs[2 * k + 1] = g2[k].y;
          dp[2 * k] = d2[k].x; dp[2 * k + 1] = d2[k].y;
        }
      }
      col_reduce_sum2<VEC>(gs, dp, cx);
      if (col_ok) {
        float2 w[H], nps[H], ngm[H], gsc[H];
        if (p.weight != nullptr) {
          IO<T>::unpack2(wq, w);
        } else {
#pragma unroll
          for (int k = 0; k < H; ++k) w[k] = f2(1.f);
        }
#pragma unroll
        for (int k = 0; k < H; ++k) {
          const float2 is2 = make_float2(istd[2 * k], istd[2 * k + 1]);
          const float2 dpn = make_float2(dp[2 * k] * p.inv_m, dp[2 * k + 1] * p.inv_m);
          nps[k] = f2mul(f2mul(dpn, f2mul(is2, is2)), f2(-1.f));            // -(dot_p / M) * invstd^2
          ngm[k] = make_float2(-gs[2 * k] * p.inv_m, -gs[2 * k + 1] * p.inv_m);  // -grad_sum / M
          gsc[k] = f2mul(w[k], is2);                                         // weight * invstd
        }
        if (cx.crank == 0 && r0 == 0) {
          if (p.dw_out != nullptr) {
            float2 dw[H];
#pragma unroll
            for (int k = 0; k < H; ++k) dw[k] = make_float2(dp[2 * k] * istd[2 * k], dp[2 * k + 1] * istd[2 * k + 1]);
            *reinterpret_cast<uint4*>(static_cast<uint8_t*>(p.dw_out) + (size_t)col * sizeof(T)) = IO<T>::pack2(dw);
          }
          if (p.db_out != nullptr) {
            float2 db[H];
#pragma unroll
            for (int k = 0; k < H; ++k) db[k] = make_float2(gs[2 * k], gs[2 * k + 1]);
            *reinterpret_cast<uint4*>(static_cast<uint8_t*>(p.db_out) + (size_t)col * sizeof(T)) = IO<T>::pack2(db);
          }
        }
        if (p.out != nullptr) {
          const uint8_t* t = t0;
#pragma unroll 2
          for (int j = 0; j < nj; ++j, t += kStep, o += out_step) {
            float2 dy[H], x[H];
            IO<T>::unpack2(*reinterpret_cast<const uint4*>(t), dy);
            IO<T>::unpack2(*reinterpret_cast<const uint4*>(t + p.tile_bytes), x);
#pragma unroll
            for (int k = 0; k < H; ++k) {
              // ((dy - proj) - grad_mean) * grad_scale, proj = proj_scale * (x - mean)
              const float2 xc = f2add(x[k], nmean[k]);
              dy[k] = f2mul(f2add(f2fma(nps[k], xc, dy[k]), ngm[k]), gsc[k]);
            }
            stg_128_stream(o, IO<T>::pack2(dy));
          }
        }
      }
    } else if constexpr (NORM) {
      // pass A: column sums -> mean
      float s[VEC];
      {
        float2 s2[H];
#pragma unroll
        for (int k = 0; k < H; ++k) s2[k] = f2(0.f);
        const uint8_t* t = t0;
#pragma unroll 4
        for (int j = 0; j < nj; ++j, t += kStep) {
          float2 x[H];
          IO<T>::unpack2(*reinterpret_cast<const uint4*>(t), x);
#pragma unroll
          for (int k = 0; k < H; ++k) s2[k] = f2add(s2[k], x[k]);
        }
#pragma unroll
        for (int k = 0; k < H; ++k) {
          s[2 * k] = s2[k].x;
          s[2 * k + 1] = s2[k].y;
        }
      }
      col_reduce<SumOp, VEC>(s, cx);
      float2 nmean[H];
#pragma unroll
      for (int k = 0; k < H; ++k) nmean[k] = make_float2(-s[2 * k] * p.inv_m, -s[2 * k + 1] * p.inv_m);
      // pass B: centred sums of squares -> invstd
      float q[VEC];
      {
        float2 q2[H];
#pragma unroll
        for (int k = 0; k < H; ++k) q2[k] = f2(0.f);
        const uint8_t* t = t0;
#pragma unroll 4
        for (int j = 0; j < nj; ++j, t += kStep) {
          float2 x[H];
          IO<T>::unpack2(*reinterpret_cast<const uint4*>(t), x);
#pragma unroll
          for (int k = 0; k < H; ++k) {
            const float2 d = f2add(x[k], nmean[k]);
            q2[k] = f2fma(d, d, q2[k]);
          }
        }
#pragma unroll
        for (int k = 0; k < H; ++k) {
          q[2 * k] = q2[k].x;
          q[2 * k + 1] = q2[k].y;
        }
      }
      col_reduce<SumOp, VEC>(q, cx);
      float2 scale[H], shift[H];
      float istd[VEC];
#pragma unroll
      for (int k = 0; k < VEC; ++k) istd[k] = rsqrt_fast(fmaf(q[k], p.inv_m, p.eps));
      if (col_ok) {
        float2 w[H], b[H];
        if (p.weight != nullptr) {
          IO<T>::unpack2(wq, w);
        } else {
#pragma unroll
          for (int k = 0; k < H; ++k) w[k] = f2(1.f);
        }
        if (p.bias != nullptr) {
          IO<T>::unpack2(bq, b);
        } else {
#pragma unroll
          for (int k = 0; k < H; ++k) b[k] = f2(0.f);
        }
#pragma unroll
        for (int k = 0; k < H; ++k) {
          scale[k] = f2mul(make_float2(istd[2 * k], istd[2 * k + 1]), w[k]);
          shift[k] = b[k];
        }
        // per-column statistics: written once, by the threads that own row 0 of cluster rank 0
        if (cx.crank == 0 && r0 == 0) {
#pragma unroll
          for (int k = 0; k < VEC; k += 4) {
            if (p.mean_out != nullptr)
              *reinterpret_cast<float4*>(p.mean_out + col + k) =
                  make_float4(-nmean[k / 2].x, -nmean[k / 2].y, -nmean[k / 2 + 1].x, -nmean[k / 2 + 1].y);
            if (p.invstd_out != nullptr)
              *reinterpret_cast<float4*>(p.invstd_out + col + k) = make_float4(istd[k], istd[k + 1], istd[k + 2], istd[k + 3]);
          }
        }
        const uint8_t* t = t0;
#pragma unroll 2
        for (int j = 0; j < nj; ++j, t += kStep, o += out_step) {
          float2 x[H];
          IO<T>::unpack2(*reinterpret_cast<const uint4*>(t), x);
#pragma unroll
          for (int k = 0; k < H; ++k) x[k] = f2fma(f2add(x[k], nmean[k]), scale[k], shift[k]);
          stg_128_stream(o, IO<T>::pack2(x));
        }
      }
    } else if constexpr (!BWD) {
      float m[VEC];
      if constexpr (sizeof(T) == 2) {
        uint4 mx = make_uint4(Max2<T>::kNegInf, Max2<T>::kNegInf, Max2<T>::kNegInf, Max2<T>::kNegInf);
        const uint8_t* t = t0;
#pragma unroll 4
        for (int j = 0; j < nj; ++j, t += kStep) {
          const uint4 v = *reinterpret_cast<const uint4*>(t);
          mx.x = Max2<T>::apply(mx.x, v.x);
          mx.y = Max2<T>::apply(mx.y, v.y);
          mx.z = Max2<T>::apply(mx.z, v.z);
          mx.w = Max2<T>::apply(mx.w, v.w);
        }
        IO<T>::unpack(mx, m);
      } else {
#pragma unroll
        for (int k = 0; k < VEC; ++k) m[k] = -INFINITY;
        const uint8_t* t = t0;
#pragma unroll 4
        for (int j = 0; j < nj; ++j, t += kStep) {
          float x[VEC];
          IO<T>::unpack(*reinterpret_cast<const uint4*>(t), x);
#pragma unroll
          for (int k = 0; k < VEC; ++k) m[k] = MaxOp::apply(m[k], x[k]);
        }
      }
      // a cluster meets once per slab: the max stays CTA-local here and (max, sum) pairs are exchanged
      // after the second pass (col_reduce_softmax_pair)
      const bool pair = cx.C > 1 && !p.two_exchanges;
      if (!p.debug_skip_reduce) {
        if (pair) {
          col_reduce<MaxOp, VEC, false>(m, cx);
        } else {
          col_reduce<MaxOp, VEC>(m, cx);
        }
      }
      float2 nm[H], nml[H], s2[H];
#pragma unroll
      for (int k = 0; k < H; ++k) {
        // a CTA with nothing but -inf takes its sum against 0 (= 0) instead of producing inf - inf
        const float ma = (pair && m[2 * k] == -INFINITY) ? 0.f : m[2 * k];
        const float mb = (pair && m[2 * k + 1] == -INFINITY) ? 0.f : m[2 * k + 1];
        nm[k] = make_float2(-ma, -mb);
        nml[k] = make_float2(-ma * kLog2e, -mb * kLog2e);
        s2[k] = f2(0.f);
      }
      {
        const uint8_t* t = t0;
#pragma unroll 2
        for (int j = 0; j < nj; ++j, t += kStep) {
          float2 x[H];
          IO<T>::unpack2(*reinterpret_cast<const uint4*>(t), x);
#pragma unroll
          for (int k = 0; k < H; ++k) s2[k] = f2add(s2[k], exp_shift<T>(x[k], nm[k], nml[k]));
        }
      }
      float s[VEC];
#pragma unroll
      for (int k = 0; k < H; ++k) {
        s[2 * k] = s2[k].x;
        s[2 * k + 1] = s2[k].y;
      }
      if (!p.debug_skip_reduce) {
        if (pair) {
          col_reduce_softmax_pair<VEC>(m, s, cx);
#pragma unroll
          for (int k = 0; k < H; ++k) {  // from here on: the cluster max
            nm[k] = make_float2(-m[2 * k], -m[2 * k + 1]);
            nml[k] = make_float2(-m[2 * k] * kLog2e, -m[2 * k + 1] * kLog2e);
          }
        } else {
          col_reduce<SumOp, VEC>(s, cx);
        }
      }
      float2 fin[H];
#pragma unroll
      for (int k = 0; k < H; ++k)
        fin[k] = LOG ? make_float2(-logf(s[2 * k]), -logf(s[2 * k + 1])) : make_float2(rcp_fast(s[2 * k]), rcp_fast(s[2 * k + 1]));
      if (col_ok) {
        const uint8_t* t = t0;
#pragma unroll 2
        for (int j = 0; j < nj; ++j, t += kStep, o += out_step) {
          float2 x[H];
          IO<T>::unpack2(*reinterpret_cast<const uint4*>(t), x);
#pragma unroll
          for (int k = 0; k < H; ++k) {
            if constexpr (LOG) {
              x[k] = f2add(f2add(x[k], nm[k]), fin[k]);
            } else {
              x[k] = f2mul(exp_shift<T>(x[k], nm[k], nml[k]), fin[k]);
            }
          }
          if (p.plain_stores) {
            *reinterpret_cast<uint4*>(o) = IO<T>::pack2(x);
          } else {
            stg_128_stream(o, IO<T>::pack2(x));
          }
        }
      }
    } else {
      float2 s2[H];
#pragma unroll
      for (int k = 0; k < H; ++k) s2[k] = f2(0.f);
      const uint8_t* t = t0;
#pragma unroll 4
      for (int j = 0; j < nj; ++j, t += kStep) {
        float2 dy[H], y[H];
        IO<T>::unpack2(*reinterpret_cast<const uint4*>(t), dy);
        if constexpr (LOG) {
#pragma unroll
          for (int k = 0; k < H; ++k) s2[k] = f2add(s2[k], dy[k]);
        } else {
          IO<T>::unpack2(*reinterpret_cast<const uint4*>(t + p.tile_bytes), y);
#pragma unroll
          for (int k = 0; k < H; ++k) s2[k] = f2fma(y[k], dy[k], s2[k]);
        }
      }
      float s[VEC];
#pragma unroll
      for (int k = 0; k < H; ++k) {
        s[2 * k] = s2[k].x;
        s[2 * k + 1] = s2[k].y;
      }
      if (!p.debug_skip_reduce) col_reduce<SumOp, VEC>(s, cx);
      float2 ns[H];
#pragma unroll
      for (int k = 0; k < H; ++k) {
        if (p.round_sum) {
          s[2 * k] = IO<T>::round_trip(s[2 * k]);
          s[2 * k + 1] = IO<T>::round_trip(s[2 * k + 1]);
        }
        ns[k] = make_float2(-s[2 * k], -s[2 * k + 1]);
      }
      if (col_ok) {
        t = t0;
#pragma unroll 2
        for (int j = 0; j < nj; ++j, t += kStep, o += out_step) {
          float2 dy[H], y[H];
          IO<T>::unpack2(*reinterpret_cast<const uint4*>(t), dy);
          IO<T>::unpack2(*reinterpret_cast<const uint4*>(t + p.tile_bytes), y);
#pragma unroll
          for (int k = 0; k < H; ++k) {
            if constexpr (LOG) {
              dy[k] = f2fma(make_float2(expf(y[k].x), expf(y[k].y)), ns[k], dy[k]);
            } else {
              dy[k] = f2fma(y[k], ns[k], f2mul(y[k], dy[k]));
            }
          }
          if (p.plain_stores) {
            *reinterpret_cast<uint4*>(o) = IO<T>::pack2(dy);
          } else {
            stg_128_stream(o, IO<T>::pack2(dy));
          }
        }
      }
    }
    __syncthreads();  // every thread is done with tile `buf` (and with fin / wpart)
    if (p.depth == 1 && next < p.nslabs && tid == 0) issue(next, 0);
  }
  pdl_launch_dependents();
  if (cx.C > 1) cluster_sync_all();  // no CTA leaves while a peer may still store into its slots
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
namespace {

int set_err(char* err, size_t n, int code, const char* fmt, ...) {
  if (err != nullptr && n > 0) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err, n, fmt, ap);
    va_end(ap);
  }
  return code;
}

int env_int_outer(const char* name) {
  const char* v = getenv(name);
  return (v != nullptr && *v != 0) ? atoi(v) : 0;
}

using OuterKernel = void (*)(const CUtensorMap, const CUtensorMap, const OuterParams);

template <typename T>
OuterKernel outer_kernel_typed(int kind) {
  switch (kind) {
    case 0: return outer_softmax_kernel<T, 0>;
    case 1: return outer_softmax_kernel<T, 1>;
    case 2: return outer_softmax_kernel<T, 2>;
    case 3: return outer_softmax_kernel<T, 3>;
    case 4: return outer_softmax_kernel<T, 4>;
    case 5: return outer_softmax_kernel<T, 5>;
    default: return nullptr;
  }
}
OuterKernel outer_kernel_for(int dt, int kind) {
  switch (dt) {
    case B200NORM_DT_F32: return outer_kernel_typed<float>(kind);
    case B200NORM_DT_BF16: return outer_kernel_typed<__nv_bfloat16>(kind);
    case B200NORM_DT_F16: return outer_kernel_typed<__half>(kind);
    default: return nullptr;
  }
}

uint32_t up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

constexpr int kSmemPerSm = 233472;  // 228 KB per SM on sm_100; each resident CTA reserves 1 KB

struct Geometry {
  int rows_cta, box_rows, nboxes;
  uint32_t tile, stage, overhead;
};

Geometry geometry(const OuterArgs& a, int W, int C) {
  const int es = a.dtype == B200NORM_DT_F32 ? 4 : 2;
  const int ns = a.bwd ? 2 : 1;
  Geometry g;
  int64_t rc = (a.M + C - 1) / C;
  if (rc <= 256) {
    rc = (rc + 7) / 8 * 8;
    g.box_rows = (int)rc;
  } else {
    rc = (rc + 127) / 128 * 128;
    g.box_rows = 128;
  }
  g.rows_cta = (int)rc;
  g.nboxes = g.rows_cta / g.box_rows;
  g.tile = (uint32_t)g.rows_cta * (uint32_t)W;
  g.stage = (uint32_t)ns * g.tile;
  const uint32_t sc = (uint32_t)(W / es);
  g.overhead = up(128 + 2 * kOuterWarps * sc * 4 + (C > 1 ? 4u * (uint32_t)C * sc * 4 : 0u) + 3 * sc * 4, 128);
  return g;
}

}  // namespace

bool outer_make_plan(const OuterArgs& a, OuterPlan* plan, char* err, size_t errlen) {
  const int es = a.dtype == B200NORM_DT_F32 ? 4 : 2;
  if (a.M < 1 || a.N < 1) {
    set_err(err, errlen, 0, "outer-persistent: empty problem");
    return false;
  }
  if ((a.N * es) % 16 != 0) {
    set_err(err, errlen, 0, "outer-persistent: N * sizeof(T) = %lld is not a multiple of 16", (long long)(a.N * es));
    return false;
  }
  if (a.M > (int64_t)1 << 24 || a.N > 0x7fffffffLL) {
    set_err(err, errlen, 0, "outer-persistent: extents out of range");
    return false;
  }
  const int64_t row_bytes = a.N * es;
  // widest slab that still gives every SM a couple of slabs
  int wpref = 128;
  for (int w : {512, 256}) {
    if ((row_bytes + w - 1) / w >= 2 * (int64_t)a.sm_count) {
      wpref = w;
      break;
    }
  }
  const int forced_w = env_int_outer("B200NORM_OUTER_W");
  const int forced_depth = env_int_outer("B200NORM_OUTER_DEPTH");
  // Tile policy (measured, profiles/r01j_outer_sweep.txt): the per-slab fixed cost (two column reductions, the
  // cluster exchange, barrier latency) is amortised by tiles of ~64 KB, and load / compute overlap comes from
  // THREE co-resident CTAs per SM with one tile each rather than from two half-size tiles per CTA; clusters are
  // kept as small as the extent allows (every doubling adds exchange traffic and lock-step waiting). Order of
  // preference: one tile at three CTAs per SM, two tiles at two / one CTA per SM, one tile at one CTA per SM;
  // inside a tier the smallest cluster, then the widest slab that fits.
  int target_per_sm = env_int_outer("B200NORM_OUTER_PER_SM");  // tuning: 1..3 CTAs per SM for the first tier
  if (target_per_sm < 1 || target_per_sm > 3) target_per_sm = 3;
  const int64_t limit3 = kSmemPerSm / target_per_sm - 1024;
  const int64_t limit2 = kSmemPerSm / 2 - 1024;
  const int64_t limit1 = a.smem_optin;
  struct Tier {
    int depth;
    int64_t limit;
  };
  const Tier tiers[] = {{1, limit3}, {2, limit2}, {2, limit1}, {1, limit1}};
  static const int cs[] = {1, 2, 4, 8, 16};
  int W = 0, C = 0, depth = 0;
  for (const Tier& t : tiers) {
    if (forced_depth != 0 && t.depth != forced_depth) continue;
    for (int c : cs) {
      if (a.cluster_forced != 0 && c != a.cluster_forced) continue;
      for (int w : {512, 256, 128, 64}) {
        if (forced_w != 0 ? w != forced_w : w > wpref) continue;
        const Geometry g = geometry(a, w, c);
        if ((int64_t)g.overhead + (int64_t)t.depth * g.stage <= t.limit) {
          W = w;
          C = c;
          depth = t.depth;
          break;
        }
      }
      if (W != 0) break;
    }
    if (W != 0) break;
  }
  if (W == 0) {
    set_err(err, errlen, 0,
            "outer-persistent: M = %lld does not fit the shared memory of a 16-CTA cluster (the reference segments "
            "such fusions too)",
            (long long)a.M);
    return false;
  }
  int64_t nslabs = (row_bytes + W - 1) / W;
  // few slabs: split the rows over more CTAs so that more SMs stream
  while (a.cluster_forced == 0 && C < 16 && nslabs * C < a.sm_count && (a.M + 2 * C - 1) / (2 * C) >= 32) C *= 2;
  const Geometry g = geometry(a, W, C);
  OuterPlan p;
  p.cluster = C;
  p.block = kOuterThreads;
  p.slab_bytes = W;
  p.rows_cta = g.rows_cta;
  p.box_rows = g.box_rows;
  p.depth = depth;
  p.tile_bytes = g.tile;
  p.stage_bytes = g.stage;
  const uint32_t sc = (uint32_t)(W / es);
  uint32_t off = 128;
  p.off_wpart = off;
  off += 2 * kOuterWarps * sc * 4;  // two planes (col_reduce_sum2)
  p.off_cls = off;
  if (C > 1) off += 4u * (uint32_t)C * sc * 4;  // two slot sets of [C][max | sum][SC]
  p.off_fin = off;
  off += 3 * sc * 4;
  off = up(off, 128);
  p.off_tiles = off;
  p.smem = off + (uint32_t)depth * g.stage;
  p.nslabs = (int)nslabs;
  int per_sm = (int)(kSmemPerSm / ((int64_t)p.smem + 1024));
  if (per_sm > 3) per_sm = 3;  // __launch_bounds__(256, 3)
  if (per_sm < 1) per_sm = 1;
  p.ctas_per_sm = per_sm;
  int64_t nclusters = C > 1 ? (int64_t)8 * (18 / C) * per_sm : (int64_t)a.sm_count * per_sm;  // nominal; refined at launch
  if (nclusters < 1) nclusters = 1;
  if (nclusters > nslabs) nclusters = nslabs;
  p.grid = (int)(nclusters * C);
  *plan = p;
  return true;
}

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &st) == cudaSuccess &&
        st == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
    } else {
      (void)cudaGetLastError();
    }
  });
  return fn;
}

bool make_map(CUtensorMap* map, const OuterArgs& a, const void* base, int64_t row_stride, int W, int box_rows) {
  const int es = a.dtype == B200NORM_DT_F32 ? 4 : 2;
  const CUtensorMapDataType dt = a.dtype == B200NORM_DT_F32    ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : a.dtype == B200NORM_DT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                               : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const cuuint64_t dims[2] = {(cuuint64_t)a.N, (cuuint64_t)a.M};
  const cuuint64_t strides[1] = {(cuuint64_t)((a.M > 1 ? row_stride : a.N) * es)};
  const cuuint32_t box[2] = {(cuuint32_t)(W / es), (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return false;
  return fn(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) ==
         CUDA_SUCCESS;
}

// resident clusters (C > 1) or CTAs (C == 1) of this launch shape on the current device
int resident_units(OuterKernel k, int smem, int C, int sm_count) {
  static std::mutex mu;
  static std::map<std::tuple<int, const void*, int, int>, int> cache;  // per device: occupancy belongs to a device context
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) (void)cudaGetLastError();
  const auto key = std::make_tuple(dev, (const void*)k, smem, C);
  {
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
  }
  int n = 0;
  if (C > 1) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(sm_count / C * C * 2));
    cfg.blockDim = dim3(kOuterThreads);
    cfg.dynamicSmemBytes = (size_t)smem;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = (unsigned)C;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    if (cudaOccupancyMaxActiveClusters(&n, (const void*)k, &cfg) != cudaSuccess) {
      (void)cudaGetLastError();
      n = 0;
    }
    if (n <= 0) n = 8 * (18 / C) > 0 ? 8 * (18 / C) : 1;
  } else {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)k, kOuterThreads, (size_t)smem) != cudaSuccess) {
      (void)cudaGetLastError();
      per_sm = 0;
    }
    if (per_sm < 1) per_sm = 1;
    n = per_sm * sm_count;
  }
  std::lock_guard<std::mutex> lock(mu);
  cache[key] = n;
  return n;
}

}  // namespace

namespace {
thread_local int g_outer_last_launches = 0;
}
int outer_last_launches() { return g_outer_last_launches; }

void outer_generic_plan(const OuterArgs& a, int* grid_x, int* nsplit, int* rows_per_split) {
  const int64_t ncb = (a.N + kOuterGenThreads - 1) / kOuterGenThreads;
  int64_t want = (16 * (int64_t)a.sm_count + ncb - 1) / ncb;  // ~16 CTAs of 128 threads per SM in total
  if (want > (a.M + 31) / 32) want = (a.M + 31) / 32;           // at least 32 rows per split
  if (want > kOuterGenMaxSplits) want = kOuterGenMaxSplits;
  if (want > a.M) want = a.M;
  if (want < 1) want = 1;
  const int64_t rps = (a.M + want - 1) / want;
  *rows_per_split = (int)(rps < 1 ? 1 : rps);
  *nsplit = (int)((a.M + *rows_per_split - 1) / *rows_per_split);
  if (*nsplit < 1) *nsplit = 1;
  *grid_x = (int)ncb;
}

size_t outer_generic_workspace_bytes(const OuterArgs& a) {
  if (a.M <= 0 || a.N <= 0) return B200NORM_WS_HEADER_BYTES;
  int gx = 1, ns = 1, rps = 1;
  outer_generic_plan(a, &gx, &ns, &rps);
  const size_t gen = ((size_t)ns + 1u) * 2u * (size_t)a.N * 4u;                    // partials + merged values
  const size_t rows = ((size_t)a.sm_count * 3u + 2u) * (size_t)a.N * 4u;          // outer_rows.cuh: per-CTA triples + results
  return (size_t)B200NORM_WS_HEADER_BYTES + (gen > rows ? gen : rows);
}

namespace {

template <typename T, int KIND>
void launch_outer_generic_kind(const OuterGenParams& p, dim3 grid, cudaStream_t stream) {
  outer_generic_stats<T, KIND><<<grid, kOuterGenThreads, 0, stream>>>(p);
  outer_generic_merge<T, KIND><<<dim3(grid.x), kOuterGenThreads, 0, stream>>>(p);
  outer_generic_apply<T, KIND><<<grid, kOuterGenThreads, 0, stream>>>(p);
}
template <typename T>
void launch_outer_generic_typed(int kind, const OuterGenParams& p, dim3 grid, cudaStream_t stream) {
  switch (kind) {
    case 0: launch_outer_generic_kind<T, 0>(p, grid, stream); break;
    case 1: launch_outer_generic_kind<T, 1>(p, grid, stream); break;
    case 2: launch_outer_generic_kind<T, 2>(p, grid, stream); break;
    case 3: launch_outer_generic_kind<T, 3>(p, grid, stream); break;
    case 4: launch_outer_generic_kind<T, 4>(p, grid, stream); break;
    default: launch_outer_generic_kind<T, 5>(p, grid, stream); break;
  }
}

int outer_generic_launch(const OuterArgs& a, int kind, void* stream_v, char* err, size_t errlen) {
  const size_t need = outer_generic_workspace_bytes(a);
  if (a.workspace == nullptr || a.workspace_bytes < need)
    return set_err(err, errlen, B200NORM_ERR_WORKSPACE, "outer generic family: workspace %zu bytes, need %zu", a.workspace_bytes, need);
  if ((reinterpret_cast<uintptr_t>(a.workspace) & 255u) != 0)
    return set_err(err, errlen, B200NORM_ERR_WORKSPACE, "workspace must be 256-byte aligned");
  if (a.M > 0x7fffffffLL || a.N > 0x7fffffffLL) return set_err(err, errlen, B200NORM_ERR_UNSUPPORTED, "outer generic family: extents exceed int32");
  OuterGenParams p;
  memset(&p, 0, sizeof(p));
  p.in0 = a.in0;
  p.in1 = a.in1;
  p.out = a.out;
  p.s0 = a.M > 1 ? a.in0_stride : a.N;
  p.s1 = a.M > 1 ? a.in1_stride : a.N;
  p.so = a.M > 1 ? a.out_stride : a.N;
  p.M = (int32_t)a.M;
  p.N = (int32_t)a.N;
  int gx = 1, ns = 1, rps = 1;
  outer_generic_plan(a, &gx, &ns, &rps);
  p.nsplit = ns;
  p.rows_per_split = rps;
  p.part = reinterpret_cast<float*>(static_cast<uint8_t*>(a.workspace) + B200NORM_WS_HEADER_BYTES);
  p.weight = a.weight;
  p.bias = a.bias;
  p.mean_out = a.mean_out;
  p.invstd_out = a.invstd_out;
  p.mean_in = a.mean_in;
  p.invstd_in = a.invstd_in;
  p.dw_out = a.dw_out;
  p.db_out = a.db_out;
  p.inv_m = (float)(1.0 / (double)a.M);
  p.eps = (float)a.eps;
  p.round_sum = a.round_sum;
  const dim3 grid((unsigned)gx, (unsigned)ns);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  switch (a.dtype) {
    case B200NORM_DT_F32: launch_outer_generic_typed<float>(kind, p, grid, stream); break;
    case B200NORM_DT_BF16: launch_outer_generic_typed<__nv_bfloat16>(kind, p, grid, stream); break;
    case B200NORM_DT_F16: launch_outer_generic_typed<__half>(kind, p, grid, stream); break;
    default: return set_err(err, errlen, B200NORM_ERR_UNSUPPORTED, "outer generic family: dtype %d", a.dtype);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_err(err, errlen, B200NORM_ERR_CUDA + (int)e, "outer generic launch failed: %s", cudaGetErrorString(e));
  return B200NORM_OK;
}

// cudaFuncSetAttribute once per (device, kernel, size)
cudaError_t outer_set_max_smem(const void* k, size_t smem) {
  static std::mutex mu;
  static std::unordered_map<uint64_t, size_t> done;
  int dev = 0;
  cudaGetDevice(&dev);
  const uint64_t key = (uint64_t)reinterpret_cast<uintptr_t>(k) ^ ((uint64_t)dev << 56);
  std::lock_guard<std::mutex> lock(mu);
  auto it = done.find(key);
  if (it != done.end() && it->second >= smem) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) done[key] = smem;
  return e;
}

// Few columns, many rows (outer_rows.cuh): one persistent kernel with two grid barriers. Returns B200NORM_ERR_UNSUPPORTED when
// the shape / layout is not its own (the caller then takes the three-kernel family).
template <typename T>
cudaError_t launch_outer_rows_typed(const OuterRowsParams& p, bool bwd, int grid, size_t smem, cudaStream_t stream) {
  const void* k = bwd ? (const void*)outer_rows_kernel<T, true> : (const void*)outer_rows_kernel<T, false>;
  cudaError_t e = outer_set_max_smem(k, smem);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kOuterRowsBlock);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  // Two in-kernel grid barriers need every CTA resident: the grid is at most one CTA per SM and the kernels of this
  // library that spin on a grid barrier are serialised across streams by the caller (barrier_guard in b200norm.cu),
  // as for the fused backward kernels; B200NORM_COOP=1 asks the driver for the formal guarantee instead (other
  // processes under MPS) and gives up the overlap of the prologue with the predecessor's tail.
  cudaLaunchAttribute attr;
  if (env_int_outer("B200NORM_COOP") != 0) {
    attr.id = cudaLaunchAttributeCooperative;
    attr.val.cooperative = 1;
  } else {
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = env_int_outer("B200NORM_NO_PDL") == 0 ? 1 : 0;
  }
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  void* args[] = {const_cast<OuterRowsParams*>(&p)};
  return cudaLaunchKernelExC(&cfg, k, args);
}

int outer_rows_launch(const OuterArgs& a, void* stream_v, char* err, size_t errlen) {
  const int es = a.dtype == B200NORM_DT_F32 ? 4 : 2;
  const int64_t row_bytes = a.N * es;
  auto al = [](const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15u) == 0; };
  int grid = 1, rpi = 1;
  if (!outer_rows_shape(a, &grid, &rpi)) return B200NORM_ERR_UNSUPPORTED;
  if (a.in0_stride != a.N || (a.bwd && a.in1_stride != a.N) || (a.out != nullptr && a.out_stride != a.N))
    return B200NORM_ERR_UNSUPPORTED;
  if (!al(a.in0) || (a.bwd && !al(a.in1)) || (a.out != nullptr && !al(a.out))) return B200NORM_ERR_UNSUPPORTED;
  OuterRowsParams p;
  memset(&p, 0, sizeof(p));
  p.in0 = a.in0;
  p.in1 = a.in1;
  p.out = a.out;
  p.M = a.M;
  p.N = (int32_t)a.N;
  p.nchunks = (int32_t)(row_bytes / 16);
  p.rpi = rpi;
  p.active = p.rpi * p.nchunks;
  p.rows_per_cta = (a.M + grid - 1) / grid;
  const size_t need = (size_t)B200NORM_WS_HEADER_BYTES + ((size_t)grid * 3u + 2u) * (size_t)a.N * 4u;
  if (a.workspace == nullptr || a.workspace_bytes < need || (reinterpret_cast<uintptr_t>(a.workspace) & 255u) != 0)
    return set_err(err, errlen, B200NORM_ERR_WORKSPACE, "outer rows kernel: workspace %zu bytes, need %zu", a.workspace_bytes, need);
  p.sync_word = reinterpret_cast<uint32_t*>(a.workspace);
  p.part = reinterpret_cast<float*>(static_cast<uint8_t*>(a.workspace) + B200NORM_WS_HEADER_BYTES);
  p.fin = p.part + (size_t)grid * 3u * (size_t)a.N;
  p.weight = a.weight;
  p.bias = a.bias;
  p.mean_out = a.mean_out;
  p.invstd_out = a.invstd_out;
  p.mean_in = a.mean_in;
  p.invstd_in = a.invstd_in;
  p.dw_out = a.dw_out;
  p.db_out = a.db_out;
  p.inv_m = (float)(1.0 / (double)a.M);
  p.eps = (float)a.eps;
  p.timeline = a.timeline;
  const int vec = 16 / es;
  p.tile_bytes = (uint32_t)kOuterRowsU * (uint32_t)p.active * 16u;
  p.ring_off = (uint32_t)kOuterRowsBarBytes + (((uint32_t)p.active * vec * 8u + 127u) & ~127u);
  const uint32_t slot = (a.bwd ? 2u : 1u) * p.tile_bytes;
  int depth = (int)(((uint32_t)a.smem_optin - p.ring_off) / slot);
  if (depth > kOuterRowsMaxDepth) depth = kOuterRowsMaxDepth;
  if (int dd = env_int_outer("B200NORM_OUTER_ROWS_DEPTH"); dd > 0 && dd < depth) depth = dd;
  if (depth < 2) return B200NORM_ERR_UNSUPPORTED;
  p.depth = depth;
  const size_t smem = (size_t)p.ring_off + (size_t)depth * slot;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  cudaError_t e;
  switch (a.dtype) {
    case B200NORM_DT_F32: e = launch_outer_rows_typed<float>(p, a.bwd != 0, grid, smem, stream); break;
    case B200NORM_DT_BF16: e = launch_outer_rows_typed<__nv_bfloat16>(p, a.bwd != 0, grid, smem, stream); break;
    case B200NORM_DT_F16: e = launch_outer_rows_typed<__half>(p, a.bwd != 0, grid, smem, stream); break;
    default: return B200NORM_ERR_UNSUPPORTED;
  }
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return set_err(err, errlen, B200NORM_ERR_CUDA + (int)e, "outer rows launch failed: %s", cudaGetErrorString(e));
  }
  return B200NORM_OK;
}

}  // namespace

// Where both families can run: the row-block kernel for rows of at most 2 KB and at least 32 MB of input. Measured
// (profiles/r02h_outer_rows.txt, forward | backward, cluster-persistent vs row-block): bf16 [25088,1024] 37.2 vs 35.4 |
// 55.9 vs 48.0 us, [50176,512] 41.7 vs 33.3 | 46.1 vs 46.6; wider rows go the other way ([12544,2048] 35.7 vs 39.9,
// [8192,4096] 38.5 vs 55.7: the cross-CTA merge of thousands of columns), and so do small matrices ([25088,256] 16.4
// vs 18.7, [50000,64] 13.0 vs 16.5: two grid barriers against none).
bool outer_prefers_rows(const OuterArgs& a) {
  int grid = 0, rpi = 0;
  if (!outer_rows_shape(a, &grid, &rpi)) return false;
  if (env_int_outer("B200NORM_OUTER_ROWS") != 0) return true;
  const int es = a.dtype == B200NORM_DT_F32 ? 4 : 2;
  return a.N * es <= 2048 && a.M * a.N * es >= (32ll << 20);
}

bool outer_rows_shape(const OuterArgs& a, int* grid_out, int* rows_per_iter) {
  const int es = a.dtype == B200NORM_DT_F32 ? 4 : 2;
  const int64_t row_bytes = a.N * es;
  if (!a.norm || a.force_generic || row_bytes % 16 != 0 || row_bytes / 16 > kOuterRowsThreads || a.M < 2048) return false;
  if (env_int_outer("B200NORM_NO_OUTER_ROWS") != 0) return false;
  const int rpi = kOuterRowsThreads / (int)(row_bytes / 16);
  int grid = a.sm_count;
  const int64_t min_rows = (int64_t)rpi * 4;  // at least four iterations per CTA
  if ((int64_t)grid * min_rows > a.M) grid = (int)((a.M + min_rows - 1) / min_rows);
  if (grid < 1) grid = 1;
  const int64_t rows_per_cta = (a.M + grid - 1) / grid;
  grid = (int)((a.M + rows_per_cta - 1) / rows_per_cta);
  *grid_out = grid;
  *rows_per_iter = rpi;
  return true;
}

int outer_launch(const OuterArgs& a, void* stream_v, char* err, size_t errlen) {
  const int kind = a.norm ? (a.bwd ? 5 : 4) : (a.bwd ? 2 : 0) + (a.log ? 1 : 0);
  const int es = a.dtype == B200NORM_DT_F32 ? 4 : 2;
  auto aligned = [&](const void* ptr, int64_t stride) {
    return (reinterpret_cast<uintptr_t>(ptr) & 15u) == 0 && (a.M <= 1 || (stride * es) % 16 == 0);
  };
  auto aligned_vec = [](const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15u) == 0; };
  OuterPlan pl;
  char why[160] = "";
  bool persistent = !a.force_generic && outer_make_plan(a, &pl, why, sizeof(why));
  if (persistent && a.norm &&
      (!aligned_vec(a.weight) || !aligned_vec(a.bias) || !aligned_vec(a.mean_out) || !aligned_vec(a.invstd_out) ||
       !aligned_vec(a.mean_in) || !aligned_vec(a.invstd_in) || !aligned_vec(a.dw_out) || !aligned_vec(a.db_out)))
    persistent = false;
  if (persistent && (!aligned(a.in0, a.in0_stride) || (a.out != nullptr && !aligned(a.out, a.out_stride)) ||
                     (a.bwd && !aligned(a.in1, a.in1_stride))))
    persistent = false;
  // whatever the persistent kernels cannot take goes to the shape-agnostic family (two kernels, matrix read twice)
  g_outer_last_launches = persistent ? 1 : 3;
  if (!persistent && a.force_persistent)
    return set_err(err, errlen, B200NORM_ERR_UNSUPPORTED, "requested kernel family cannot run this outer segment: %s", why);
  if (persistent && !a.force_persistent && outer_prefers_rows(a)) {
    // narrow rows and a matrix of tens of MB: the row-block kernel beats the cluster-persistent family (measured,
    // see outer_prefers_rows); B200NORM_OUTER_ROWS=1 takes it wherever it can run
    const int rc = outer_rows_launch(a, stream_v, err, errlen);
    if (rc == B200NORM_OK) {
      g_outer_last_launches = 1;
      return rc;
    }
    if (rc != B200NORM_ERR_UNSUPPORTED) return rc;
  }
  if (!persistent) {
    // few columns, many rows (ResNet-size batch norms): the single-kernel row-block family; anything it declines
    // (and B200NORM_VARIANT_GENERIC) goes to the three-kernel family
    if (!a.force_generic && a.norm) {
      const int rc = outer_rows_launch(a, stream_v, err, errlen);
      if (rc == B200NORM_OK) {
        g_outer_last_launches = 1;
        return rc;
      }
      if (rc != B200NORM_ERR_UNSUPPORTED) return rc;
    }
    return outer_generic_launch(a, kind, stream_v, err, errlen);
  }

  OuterKernel k = outer_kernel_for(a.dtype, kind);
  if (k == nullptr) return set_err(err, errlen, B200NORM_ERR_UNSUPPORTED, "outer-persistent: dtype %d", a.dtype);
  {
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, std::pair<int, int>> done;  // function attributes are per device
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) (void)cudaGetLastError();
    std::lock_guard<std::mutex> lock(mu);
    auto& st = done[std::make_pair(dev, (const void*)k)];
    if ((int)pl.smem > st.first) {
      cudaError_t e = cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
      if (e != cudaSuccess) return set_err(err, errlen, B200NORM_ERR_CUDA + (int)e, "outer smem attr: %s", cudaGetErrorString(e));
      st.first = (int)pl.smem;
    }
    if (pl.cluster > 8 && !st.second) {
      cudaError_t e = cudaFuncSetAttribute((const void*)k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      if (e != cudaSuccess) return set_err(err, errlen, B200NORM_ERR_CUDA + (int)e, "outer cluster attr: %s", cudaGetErrorString(e));
      st.second = 1;
    }
  }
  int64_t units = resident_units(k, (int)pl.smem, pl.cluster, a.sm_count);
  if (int g = env_int_outer("B200NORM_OUTER_UNITS")) units = g;
  if (units > pl.nslabs) units = pl.nslabs;
  if (units < 1) units = 1;

  CUtensorMap map0, map1;
  memset(&map0, 0, sizeof(map0));
  memset(&map1, 0, sizeof(map1));
  if (!make_map(&map0, a, a.in0, a.in0_stride, pl.slab_bytes, pl.box_rows))
    return set_err(err, errlen, B200NORM_ERR_CUDA, "outer-persistent: cuTensorMapEncodeTiled failed for input 0");
  if (a.bwd) {
    if (!make_map(&map1, a, a.in1, a.in1_stride, pl.slab_bytes, pl.box_rows))
      return set_err(err, errlen, B200NORM_ERR_CUDA, "outer-persistent: cuTensorMapEncodeTiled failed for input 1");
  } else {
    map1 = map0;
  }

  OuterParams p;
  memset(&p, 0, sizeof(p));
  p.out = a.out;
  p.out_stride = a.M > 1 ? a.out_stride : a.N;
  p.M = (int32_t)a.M;
  p.N = (int32_t)a.N;
  p.rows_cta = pl.rows_cta;
  p.box_rows = pl.box_rows;
  p.nboxes = pl.rows_cta / pl.box_rows;
  p.slab_cols = pl.slab_bytes / es;
  int lc = pl.slab_bytes / 16, sh = 0;
  while ((1 << sh) < lc) ++sh;
  p.lc_shift = sh;
  p.nslabs = pl.nslabs;
  p.cluster = pl.cluster;
  p.depth = pl.depth;
  p.round_sum = a.round_sum;
  p.weight = a.weight;
  p.bias = a.bias;
  p.mean_out = a.mean_out;
  p.invstd_out = a.invstd_out;
  p.mean_in = a.mean_in;
  p.invstd_in = a.invstd_in;
  p.dw_out = a.dw_out;
  p.db_out = a.db_out;
  p.inv_m = (float)(1.0 / (double)a.M);
  p.eps = (float)a.eps;
  p.debug_skip_reduce = env_int_outer("B200NORM_OUTER_SKIP_REDUCE");
  p.plain_stores = env_int_outer("B200NORM_OUTER_PLAIN_ST");
  p.l2_prefetch = env_int_outer("B200NORM_OUTER_L2PF") != 0;  // measured neutral to -8 %: off by default
  p.two_exchanges = env_int_outer("B200NORM_OUTER_TWO_EXCH");
  p.slab_bytes = (uint32_t)pl.slab_bytes;
  p.tile_bytes = pl.tile_bytes;
  p.stage_bytes = pl.stage_bytes;
  p.off_wpart = pl.off_wpart;
  p.off_cls = pl.off_cls;
  p.off_fin = pl.off_fin;
  p.off_tiles = pl.off_tiles;

  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(units * pl.cluster));
  cfg.blockDim = dim3(kOuterThreads);
  cfg.dynamicSmemBytes = pl.smem;
  cfg.stream = static_cast<cudaStream_t>(stream_v);
  cudaLaunchAttribute attrs[2];
  int na = 0;
  attrs[na].id = cudaLaunchAttributeClusterDimension;
  attrs[na].val.clusterDim.x = (unsigned)pl.cluster;
  attrs[na].val.clusterDim.y = 1;
  attrs[na].val.clusterDim.z = 1;
  ++na;
  if (env_int_outer("B200NORM_NO_PDL") == 0) {
    attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attrs;
  cfg.numAttrs = (unsigned)na;
  cudaError_t e = cudaLaunchKernelEx(&cfg, k, map0, map1, p);
  if (e != cudaSuccess && na == 2) {  // retry without programmatic serialization
    (void)cudaGetLastError();
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, k, map0, map1, p);
  }
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return set_err(err, errlen, B200NORM_ERR_CUDA + (int)e, "outer softmax launch failed: %s", cudaGetErrorString(e));
  }
  if (env_int_outer("B200NORM_DEBUG") > 1)
    fprintf(stderr, "[b200norm] outer softmax bwd=%d log=%d W=%d C=%d Rc=%d depth=%d grid=%d smem=%u\n", a.bwd, a.log,
            pl.slab_bytes, pl.cluster, pl.rows_cta, pl.depth, (int)(units * pl.cluster), pl.smem);
  return B200NORM_OK;
}

}  // namespace b200norm

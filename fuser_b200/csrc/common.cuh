// common.cuh — device-side building blocks for the b200norm kernels (sm_100a only).
//
// What lives here is the hand-written equivalent of the helper layer nvFuser
// stringifies into every generated kernel (reference: runtime/warp.cu,
// runtime/mbarrier.cu, runtime/memory.cu:104-130, runtime/cluster.cu:15-130,
// runtime/array.cu:202-400, runtime/casts.cu:431-447). Nothing is copied from
// there; the primitives are re-derived from the PTX ISA for the pipeline shape
// used in this library: a ring of shared-memory stages filled by 1-D bulk
// copies (TMA) and drained by groups of consumer warps.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200norm {

// ----------------------------------------------------------------------------
// small utilities
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// A spin that can never hang the device: after `limit_ns` of waiting (default ~4 s: everything it guards
// inside ONE GPU completes in microseconds) the kernel traps — the launch then surfaces cudaErrorLaunchFailed
// instead of wedging the box. Waits for ANOTHER GPU take their limit from the peer registration (minutes by
// default, 0 = wait for ever): rank skew of seconds is normal there.
struct Watchdog {
  uint32_t spins = 0;
  uint64_t t0 = 0;
  uint64_t limit_ns;
  __device__ __forceinline__ explicit Watchdog(uint64_t limit = 4000000000ull) : limit_ns(limit) {}
  __device__ __forceinline__ void tick() {
    if ((++spins & 0x3fffu) == 0 && limit_ns != 0) {
      uint64_t now = global_timer_ns();
      if (t0 == 0) {
        t0 = now;
      } else if (now - t0 > limit_ns) {
        __trap();
      }
    }
  }
};

// ----------------------------------------------------------------------------
// mbarrier (shared::cta) — arrive / expect_tx / parity wait
// ----------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  Watchdog wd;
  while (!mbar_try_wait(bar, parity)) {
    wd.tick();
  }
}

// ----------------------------------------------------------------------------
// TMA 1-D bulk copy global -> shared with mbarrier completion
// (SASS: UBLKCP). Address and size must be multiples of 16 bytes.
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes,
                                         uint32_t bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1], %2, [%3], %4;"
      ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}
// shared -> global bulk store (SASS: UBLKCP / UTMASTG family), bulk_group completion.
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
               "r"(src_smem), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

// ----------------------------------------------------------------------------
// programmatic dependent launch: a kernel launched with the programmatic-stream-
// serialization attribute may start while its predecessor drains; it must not touch
// global memory before pdl_wait() (which returns once the predecessor has completed and
// flushed). Both are no-ops for a normal launch.
// ----------------------------------------------------------------------------
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// device-scope acquire load / release add for the in-kernel grid barrier
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t atom_add_acq_rel_gpu(uint32_t* p, uint32_t v) {
  uint32_t old;
  asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}
// Ticket counter for the dynamic tile schedule. An integer fetch-add on a warp-uniform address —
// atomicAdd() or raw atom.add.u32 alike — is rewritten by ptxas into a warp-aggregated atomic
// (VOTE / elected ATOMG / SHFL of the result), and that SHFL consumes the return value at once:
// the whole L2 round trip (1-2 us under a saturating stream) lands on the issuing warp and, through
// the group barrier, on every warp of the group (measured: +20 % per dynamically scheduled tile).
// Floating-point adds are never aggregated (it would change the rounding), so the counter is kept
// as an fp32 number: exact up to 2^24 tickets, zero has the all-zero bit pattern, and the
// destination register stays pending until its first real use.
__device__ __forceinline__ int32_t ticket_fetch_add(uint32_t* counter, int n) {
  float old;
  asm volatile("atom.relaxed.gpu.global.add.f32 %0, [%1], %2;" : "=f"(old) : "l"(counter), "f"((float)n) : "memory");
  return __float_as_int(old);  // raw bits: convert with ticket_value() at the point of use
}
__device__ __forceinline__ int32_t ticket_value(int32_t raw_bits) { return (int32_t)__int_as_float(raw_bits); }
constexpr int32_t kTicketNone = 0x4b7fffff;  // bits of 16777215.0f: "no ticket" (>= every pool size)
constexpr int64_t kTicketMax = 1 << 24;
__device__ __forceinline__ uint32_t atom_add_release_gpu(uint32_t* p, uint32_t v) {
  uint32_t old;
  asm volatile("atom.add.release.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}

// system-scope release store / acquire load / volatile vector load for the peer-memory exchange
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_volatile_f4(const float4* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}

// "LL" words of the fused peer all-reduce: one 64-bit store carries a value and the launch epoch that makes it
// valid, so the receiver polls the data itself — no flag, no fence, one NVLink one-way latency (a scalar
// 64-bit access is single-copy atomic in the PTX memory model; NCCL's LL protocol rests on the same property).
__device__ __forceinline__ void st_ll(uint64_t* p, float v, uint32_t epoch) {
  const uint64_t w = ((uint64_t)epoch << 32) | (uint64_t)__float_as_uint(v);
  asm volatile("st.relaxed.sys.global.b64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ uint64_t ld_ll(const uint64_t* p) {
  uint64_t w;
  asm volatile("ld.relaxed.sys.global.b64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
  return w;
}

// ----------------------------------------------------------------------------
// named barriers (bar.sync id, nthreads) for sub-CTA consumer groups
// ----------------------------------------------------------------------------
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ----------------------------------------------------------------------------
// thread-block cluster / DSMEM
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n"
               "barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
// Remote (or local) shared store of one float that completes 4 tx-bytes on the
// destination CTA's mbarrier.
__device__ __forceinline__ void st_async_f32(uint32_t remote_addr, float v, uint32_t remote_bar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];"
               ::"r"(remote_addr), "f"(v), "r"(remote_bar)
               : "memory");
}

// ----------------------------------------------------------------------------
// global memory access with cache hints
// ----------------------------------------------------------------------------
__device__ __forceinline__ void stg_128_stream(void* p, const uint4& v) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ uint4 ldg_128_nc(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ uint4 lds_128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}

// ----------------------------------------------------------------------------
// packed fp32x2 math (sm_100: FFMA2 / FMUL2 / FADD2 — one issue slot, two lanes;
// each lane is an IEEE round-to-nearest op, bit-identical to the scalar form)
// ----------------------------------------------------------------------------
__device__ __forceinline__ float2 f2(float s) { return make_float2(s, s); }
__device__ __forceinline__ float2 f2mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 f2add(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 f2fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 f2sub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.f, -1.f), a); }
__device__ __forceinline__ float2 bf16x2_to_f2(uint32_t u) {
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}

// ----------------------------------------------------------------------------
// dtype traits: a "chunk" is 16 bytes = VEC elements
// bf16/fp16 -> fp32 is exact; fp32 -> bf16/fp16 is round-to-nearest-even, the
// same rounding ATen's .to() and the reference's casts use (runtime/casts.cu:431-447).
// ----------------------------------------------------------------------------
template <typename T>
struct IO;

template <>
struct IO<float> {
  static constexpr int VEC = 4;
  static constexpr int DT = 0;
  __device__ __forceinline__ static void unpack(const uint4& v, float (&f)[4]) {
    f[0] = __uint_as_float(v.x);
    f[1] = __uint_as_float(v.y);
    f[2] = __uint_as_float(v.z);
    f[3] = __uint_as_float(v.w);
  }
  __device__ __forceinline__ static uint4 pack(const float (&f)[4]) {
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]),
                      __float_as_uint(f[3]));
  }
  __device__ __forceinline__ static void unpack2(const uint4& v, float2 (&f)[2]) {
    f[0] = make_float2(__uint_as_float(v.x), __uint_as_float(v.y));
    f[1] = make_float2(__uint_as_float(v.z), __uint_as_float(v.w));
  }
  __device__ __forceinline__ static uint4 pack2(const float2 (&f)[2]) {
    return make_uint4(__float_as_uint(f[0].x), __float_as_uint(f[0].y), __float_as_uint(f[1].x),
                      __float_as_uint(f[1].y));
  }
  __device__ __forceinline__ static float to_f32(float v) { return v; }
  __device__ __forceinline__ static float from_f32(float v) { return v; }
  __device__ __forceinline__ static float round_trip(float v) { return v; }
};

template <>
struct IO<__nv_bfloat16> {
  static constexpr int VEC = 8;
  static constexpr int DT = 1;
  __device__ __forceinline__ static void unpack(const uint4& v, float (&f)[8]) {
    f[0] = __uint_as_float(v.x << 16);
    f[1] = __uint_as_float(v.x & 0xffff0000u);
    f[2] = __uint_as_float(v.y << 16);
    f[3] = __uint_as_float(v.y & 0xffff0000u);
    f[4] = __uint_as_float(v.z << 16);
    f[5] = __uint_as_float(v.z & 0xffff0000u);
    f[6] = __uint_as_float(v.w << 16);
    f[7] = __uint_as_float(v.w & 0xffff0000u);
  }
  __device__ __forceinline__ static uint32_t pack2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
  }
  __device__ __forceinline__ static uint4 pack(const float (&f)[8]) {
    return make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
  }
  __device__ __forceinline__ static void unpack2(const uint4& v, float2 (&f)[4]) {
    f[0] = bf16x2_to_f2(v.x);
    f[1] = bf16x2_to_f2(v.y);
    f[2] = bf16x2_to_f2(v.z);
    f[3] = bf16x2_to_f2(v.w);
  }
  __device__ __forceinline__ static uint4 pack2(const float2 (&f)[4]) {
    return make_uint4(pack2(f[0].x, f[0].y), pack2(f[1].x, f[1].y), pack2(f[2].x, f[2].y), pack2(f[3].x, f[3].y));
  }
  __device__ __forceinline__ static float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
  __device__ __forceinline__ static __nv_bfloat16 from_f32(float v) {
    return __float2bfloat16_rn(v);
  }
  __device__ __forceinline__ static float round_trip(float v) {
    return __bfloat162float(__float2bfloat16_rn(v));
  }
};

template <>
struct IO<__half> {
  static constexpr int VEC = 8;
  static constexpr int DT = 2;
  __device__ __forceinline__ static void unpack2(uint32_t u, float& lo, float& hi) {
    __half2 h = *reinterpret_cast<__half2*>(&u);
    float2 f = __half22float2(h);
    lo = f.x;
    hi = f.y;
  }
  __device__ __forceinline__ static void unpack(const uint4& v, float (&f)[8]) {
    unpack2(v.x, f[0], f[1]);
    unpack2(v.y, f[2], f[3]);
    unpack2(v.z, f[4], f[5]);
    unpack2(v.w, f[6], f[7]);
  }
  __device__ __forceinline__ static uint32_t pack2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
  }
  __device__ __forceinline__ static uint4 pack(const float (&f)[8]) {
    return make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
  }
  __device__ __forceinline__ static void unpack2(const uint4& v, float2 (&f)[4]) {
    f[0] = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
    f[1] = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
    f[2] = __half22float2(*reinterpret_cast<const __half2*>(&v.z));
    f[3] = __half22float2(*reinterpret_cast<const __half2*>(&v.w));
  }
  __device__ __forceinline__ static uint4 pack2(const float2 (&f)[4]) {
    return make_uint4(pack2(f[0].x, f[0].y), pack2(f[1].x, f[1].y), pack2(f[2].x, f[2].y), pack2(f[3].x, f[3].y));
  }
  __device__ __forceinline__ static float to_f32(__half v) { return __half2float(v); }
  __device__ __forceinline__ static __half from_f32(float v) { return __float2half_rn(v); }
  __device__ __forceinline__ static float round_trip(float v) {
    return __half2float(__float2half_rn(v));
  }
};

// ----------------------------------------------------------------------------
// reductions
// ----------------------------------------------------------------------------
struct SumOp {
  __device__ __forceinline__ static float identity() { return 0.f; }
  __device__ __forceinline__ static float apply(float a, float b) { return a + b; }
};
// NaN-propagating max (at::amax semantics, ir/internal_nodes.cpp:1763-1765).
struct MaxOp {
  __device__ __forceinline__ static float identity() { return -INFINITY; }
  __device__ __forceinline__ static float apply(float a, float b) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
  }
};
// NaN-ignoring max (fmax semantics; all-NaN rows still give NaN because the
// identity never enters: handled by the callers through a separate flag).
struct FMaxOp {
  __device__ __forceinline__ static float identity() { return -INFINITY; }
  __device__ __forceinline__ static float apply(float a, float b) { return fmaxf(a, b); }
};

template <typename Op, int NV>
__device__ __forceinline__ void warp_allreduce(float (&v)[NV]) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[i] = Op::apply(v[i], __shfl_xor_sync(0xffffffffu, v[i], off));
    }
  }
}

// Single-MUFU reciprocal / square roots (max relative error ~1 ulp of fp32; subnormal
// inputs handled). The IEEE-rounded forms cost ~20 instructions each on the per-row
// critical path; the results here feed values that are rounded to the io dtype or
// compared under the validator's tolerances (>= 1e-6 relative).
__device__ __forceinline__ float rsqrt_fast(float x) {
  float y;
  asm("rsqrt.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sqrt_fast(float x) {
  float y;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_fast(float x) {
  float y;
  asm("rcp.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exp(x) as one MUFU.EX2 (relative error about 2^-22): sums of exponentials in any io dtype.
__device__ __forceinline__ float exp_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
  return y;
}

// Fast exp for 16-bit outputs (one MUFU.EX2 per element); accurate expf for fp32.
template <typename T>
__device__ __forceinline__ float exp_for(float x) {
  if constexpr (sizeof(T) == 4) {
    return expf(x);
  } else {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
    return y;
  }
}

// ----------------------------------------------------------------------------
// Philox4x32-10 and the uniform -> keep decision of a fused dropout, bit-compatible with the reference's runtime
// (runtime/random_numbers.cu:8-58 philox; :66-88 raw_uniform_float / uniformf; the lowering
// device_lower/pass/rng.cpp:95-127 gives element i of a Float uniform() word (i mod 4) of philox(seed, i / 4, offset)).
//   key = seed, counter = (offset lo, offset hi, subsequence lo, subsequence hi); ten rounds of
//   (hi(B*c2) ^ c1 ^ k0, lo(B*c2), hi(A*c0) ^ c3 ^ k1, lo(A*c0)), the key bumped by the Weyl constants between rounds.
// uniformf(w) = fl(fl(w) * 2^-32 + 2^-33), with 1.0 mapped to 0 — monotone in w up to that wrap, so the test
// uniformf(w) < keep (a float against the fusion's double scalar) is two integer compares against thresholds the
// host derives once (dropout_thresholds in b200norm.cu): keep  <=>  w <= t1 || w >= t2.
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(unsigned long long seed, unsigned long long subsequence,
                                               unsigned long long offset) {
  constexpr uint32_t kA = 0xD2511F53u, kB = 0xCD9E8D57u, kW0 = 0x9E3779B9u, kW1 = 0xBB67AE85u;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c0 = (uint32_t)offset, c1 = (uint32_t)(offset >> 32);
  uint32_t c2 = (uint32_t)subsequence, c3 = (uint32_t)(subsequence >> 32);
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(kA, c0), lo0 = kA * c0;
    const uint32_t hi1 = __umulhi(kB, c2), lo1 = kB * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += kW0;
    k1 += kW1;
  }
  return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ bool dropout_keep_word(uint32_t w, uint32_t t1, uint32_t t2) { return w <= t1 || w >= t2; }
// element i (any i): one Philox call, one word of it
__device__ __forceinline__ bool dropout_keep_at(unsigned long long seed, unsigned long long offset,
                                                unsigned long long i, uint32_t t1, uint32_t t2) {
  const uint4 r = philox4x32_10(seed, i >> 2, offset);
  const uint32_t c = (uint32_t)i & 3u;
  const uint32_t w = c == 0 ? r.x : (c == 1 ? r.y : (c == 2 ? r.z : r.w));
  return dropout_keep_word(w, t1, t2);
}

__host__ __device__ __forceinline__ constexpr int64_t ceil_div64(int64_t a, int64_t b) {
  return (a + b - 1) / b;
}
__host__ __device__ __forceinline__ constexpr int round_up_int(int a, int b) {
  return (a + b - 1) / b * b;
}

}  // namespace b200norm

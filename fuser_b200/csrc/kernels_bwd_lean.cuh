// kernels_bwd_lean.cuh — RMSNorm backward, textbook form (no centring term), pipelined re-read kernel with the
// instruction count cut down: the kernel the C5 config (bf16 [4096,65536] over a cluster of 4) runs.
//
// Math (csrc/ops/normalization.cpp:487-535 and the Python trace benchmarks/python/test_rmsnorm_bwd.py:20-76):
//     x_hat = x * invstd,  g = dy * w,  dx = invstd * g - (invstd / N) * proj_scale * sum(g x_hat) * x_hat,
//     dw = sum_rows dy * x_hat.
// norm_bwd_pipe_kernel (kernels_bwd.cuh) evaluates the general form (LayerNorm, centred RMSNorm) with five packed
// operations per element pair in its first pass — x_hat, g, sum g, sum g x_hat, dw — and reduces two values per row.
// Without the centring term the first pass needs three,
//     p = dy * x,   S += p * w  (sum g x_hat = invstd * S),   dw += p * invstd,
// and ONE reduced value; the ncu capture of the general kernel at C5 (profiles/r02a_ncu_c5_bwd_*) showed it
// issue-limited (760 warp instructions per row and warp, 55 % issue utilisation, every cluster in lock-step on
// its slowest CTA), so the instruction count is what this kernel attacks: leaner math, one value through the
// warp shuffles / slots / DSMEM exchange, running output and statistic pointers instead of 64-bit index
// arithmetic per row, spin loops without the watchdog bookkeeping on their fast path.
// Structure, barriers and slot protocol are those of norm_bwd_pipe_kernel (see the comment there).
#pragma once

#include "kernels_bwd.cuh"

namespace b200norm {

// wait with a cheap first probe; the watchdog loop only runs when the barrier is not ready yet
__device__ __forceinline__ void mbar_wait_lean(uint32_t bar, uint32_t parity) {
  if (!mbar_try_wait(bar, parity)) mbar_wait(bar, parity);
}

template <typename T, int TPR, int CH>
__global__ void __launch_bounds__(kMaxBlock, 1) rms_bwd_pipe_lean_kernel(const RowParams p) {
  constexpr int VEC = IO<T>::VEC;
  constexpr int H = VEC / 2;
  constexpr int WPG = TPR / 32;
  static_assert(CH > 0 && TPR >= 64, "pipelined backward: fixed chunk ownership, more than one warp per group");
  const uint32_t smem0 = smem_u32(smem_raw);
  tl_mark(p, 0);
  if (threadIdx.x == 0) {
    const int n_red = 2 * p.groups, n_free = p.groups * p.depth;
    for (int i = 0; i < n_red + n_free; ++i) mbar_init(smem0 + p.off_pipe + 8u * (uint32_t)i, WPG);
    for (int i = 0; i < p.groups; ++i) mbar_init(smem0 + p.off_pipe + 8u * (uint32_t)(n_red + n_free + i), 1);
  }
  ring_setup(p, smem0);
  tl_mark(p, 1);
  Group<TPR> g(p, smem0);
  Ring<T, TPR> ring(p, smem0, g);
  ring.prologue();

  const int nchunks = p.nchunks;
  const bool want_dx = (p.flags & kFlagOutDx) != 0;
  const bool want_dw = (p.flags & kFlagOutDw) != 0;
  const bool stat_rms = (p.flags & kFlagStatIsRms) != 0;
  const size_t col0_bytes = (size_t)g.crank * p.slice_bytes;
  // the weight stays UNPACKED (fp32 pairs) in registers: both passes of every row use it, and unpacking it
  // there costs 16 ALU operations per chunk and row (the ncu capture showed the compiler doing exactly that)
  float2 wf[CH][H];
  {
    uint4 wq[CH];
    load_params_to_regs<T, TPR, CH>(p.weight, col0_bytes, g.gtid, nchunks, 1.0f, wq);
#pragma unroll
    for (int j = 0; j < CH; ++j) IO<T>::unpack2(wq[j], wf[j]);
  }

  float2 dw_acc[CH][H];
  float2 none[1][H];
#pragma unroll
  for (int j = 0; j < CH; ++j) {
#pragma unroll
    for (int k = 0; k < H; ++k) dw_acc[j][k] = f2(0.f);
  }

  const uint32_t red_bar = smem0 + p.off_pipe + 8u * (uint32_t)(2 * g.gid);
  const uint32_t free_bar = smem0 + p.off_pipe + 8u * (uint32_t)(2 * p.groups + g.gid * p.depth);
  const uint32_t cl_bar_even = g.clbar;
  const uint32_t cl_bar_odd = smem0 + p.off_pipe + 8u * (uint32_t)(2 * p.groups + p.groups * p.depth + g.gid);
  float* red_slots = reinterpret_cast<float*>(smem_raw + p.off_red + (uint32_t)g.gid * (2 * kRedSetBytes));
  const uint32_t cls_off = p.off_cls + (uint32_t)g.gid * (2 * kClsSetBytes);

  const int n = ring.ntiles_g;
  const int64_t row_first = ring.row0, row_step = ring.row_step;
  // running pointers: the statistic of row k + 2 is fetched while row k is processed
  const float* stat_ptr = p.stat1 + row_first;
  auto stat_at = [&](int k) -> float { return k < n ? __ldg(stat_ptr + (int64_t)k * row_step) : 0.f; };
  float st_a, st_b = stat_at(0), st_c = stat_at(1);
  uint8_t* dxrow = out_row_ptr<T>(p, row_first, g.crank, g.gtid);
  const size_t dx_step = (size_t)row_step * (size_t)p.out_stride * sizeof(T);
  int64_t next_row = row_first + (int64_t)p.depth * row_step;  // the row the next refill loads

  const uint32_t stage_bytes = p.stage_bytes, stream_bytes = p.stream_bytes, depth = (uint32_t)p.depth;
  uint32_t c1_off = ring.ring_off_g, c1_bar = ring.bar_g, c1_s = 0, c1_phase = 0;
  uint32_t c2_off = ring.ring_off_g, c2_bar = ring.bar_g, c2_s = 0, c2_phase = 0;
  auto advance = [&](uint32_t& off, uint32_t& bar, uint32_t& s, uint32_t& phase) {
    ++s;
    off += stage_bytes;
    bar += 8u;
    if (s == depth) {
      s = 0;
      phase ^= 1u;
      off = ring.ring_off_g;
      bar = ring.bar_g;
    }
  };

  // first pass of one row: S = sum dy * x * w over the thread's chunks, dweight accumulation, post of the warp sum
  auto pass1_post = [&](int j_row, float st) {
    mbar_wait_lean(c1_bar, c1_phase);
    const uint8_t* dyrow = smem_raw + c1_off + 16u * (uint32_t)g.gtid;
    const uint8_t* xrow = dyrow + stream_bytes;
    const float2 is2 = f2(stat_rms ? rcp_fast(st) : st);
    float2 s = f2(0.f);
    for_each_chunk<TPR, CH>(g.gtid, nchunks, [&](int j, int coff) {
      float2 dy[H], x[H];
      IO<T>::unpack2(ld_smem_u4(dyrow + 16 * coff), dy);
      IO<T>::unpack2(ld_smem_u4(xrow + 16 * coff), x);
#pragma unroll
      for (int k = 0; k < H; ++k) {
        const float2 pr = f2mul(dy[k], x[k]);
        s = f2fma(pr, wf[j][k], s);
        dw_acc[j][k] = f2fma(pr, is2, dw_acc[j][k]);
      }
    });
    float v[1] = {s.x + s.y};
    warp_allreduce<SumOp, 1>(v);
    if (g.lane == 0) {
      red_slots[(j_row & 1) * (kRedSetBytes / 4) + g.warp] = v[0];
      mbar_arrive(red_bar + 8u * (uint32_t)(j_row & 1));
    }
    advance(c1_off, c1_bar, c1_s, c1_phase);
  };

  if (n > 0) pass1_post(0, st_b);
  st_a = st_b;
  st_b = st_c;
  st_c = stat_at(2);

  for (int k = 0; k < n; ++k) {
    if (k == 0) tl_mark(p, 2);
    const uint32_t par = (uint32_t)(k & 1), ph = (uint32_t)((k >> 1) & 1);
    const uint32_t cl_bar = par ? cl_bar_odd : cl_bar_even;
    // ---- collect1(k): the warps' sums inside the CTA (fixed order) ----
    mbar_wait_lean(red_bar + 8u * par, ph);
    float tot = 0.f;
    {
      const float* set = red_slots + par * (kRedSetBytes / 4);
#pragma unroll
      for (int w4 = 0; w4 < (WPG + 3) / 4; ++w4) {
        const float4 t = *reinterpret_cast<const float4*>(set + 4 * w4);
        if constexpr (WPG >= 4) {
          tot += (t.x + t.y) + (t.z + t.w);
        } else {
          tot += t.x + t.y;
        }
      }
    }
    if (g.csize > 1) {
      if (g.gtid == 0) mbar_arrive_expect_tx(cl_bar, g.csize * 4);
      if ((uint32_t)g.gtid < g.csize) {
        const uint32_t r = (uint32_t)g.gtid;
        st_async_f32(map_to_cta(smem0 + cls_off + par * kClsSetBytes + (g.crank * kRedNV) * 4, r), tot,
                     map_to_cta(cl_bar, r));
      }
    }
    // ---- first pass of the next row while the exchange is in flight ----
    if (k + 1 < n) pass1_post(k + 1, st_b);
    if (g.csize > 1) {
      mbar_wait_lean(cl_bar, ph);
      const volatile float* slots = reinterpret_cast<const volatile float*>(smem_raw + cls_off + par * kClsSetBytes);
      tot = 0.f;
      for (uint32_t r = 0; r < g.csize; ++r) tot += slots[r * kRedNV];
    }
    // ---- second pass of row k: dx = invstd * (dy * w) + Bx * x,  Bx = -(invstd^3 / N) * proj_scale * S ----
    if (want_dx) {
      const float invstd = stat_rms ? rcp_fast(st_a) : st_a;
      const float a_n = invstd * p.inv_n;
      const float2 A = f2(invstd);
      const float2 Bx = f2(-(a_n * (invstd * tot)) * p.proj_scale * invstd);
      const uint8_t* dyrow = smem_raw + c2_off + 16u * (uint32_t)g.gtid;
      const uint8_t* xrow = dyrow + stream_bytes;
      for_each_chunk<TPR, CH>(g.gtid, nchunks, [&](int j, int coff) {
        float2 dy[H], x[H], dx[H];
        IO<T>::unpack2(ld_smem_u4(dyrow + 16 * coff), dy);
        IO<T>::unpack2(ld_smem_u4(xrow + 16 * coff), x);
#pragma unroll
        for (int q = 0; q < H; ++q) dx[q] = f2fma(A, f2mul(dy[q], wf[j][q]), f2mul(Bx, x[q]));
        st_global_stream(dxrow + 16 * coff, IO<T>::pack2(dx));
      });
    }
    // ---- release the stage; the first warp refills it once every warp has let go ----
    __syncwarp();
    if (g.lane == 0) mbar_arrive(free_bar + 8u * c2_s);
    if (g.warp == 0) {
      mbar_wait_lean(free_bar + 8u * c2_s, c2_phase);
      if (k + (int)depth < n) ring.issue_rows(1, next_row, c2_off, c2_bar);
    }
    advance(c2_off, c2_bar, c2_s, c2_phase);
    next_row += row_step;
    dxrow += dx_step;
    st_a = st_b;
    st_b = st_c;
    st_c = stat_at(k + 3);
  }

  bwd_epilogue<T, TPR, CH, false>(p, g, dw_acc, none, want_dw, false);
  ring_teardown(p);
  tl_mark(p, 6);
}

}  // namespace b200norm

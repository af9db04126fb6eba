"""Oracle comparisons at the sizes and plans the bench times (run with -m gpu on a B200).

Every BASELINE.json config at FULL size against the ATen oracle (the validator's tolerances) and the
fp64 restatements (<= 1.01 ulp of the io dtype), each asserting the plan that the bench launches;
and multi-tile shapes that drive the steady state of every backward kernel on one GPU: ring refills,
mbarrier phase wraps, two-row tiles, cluster / pipelined kernels past their first ring revolution.
The reference's counterparts: tests/cpp/test_combined_inner_outer_reduction.cpp:51-176 (LayerNorm
backward vs at::native_layer_norm_backward over a shape sweep), tests/cpp/test_persistent_buffer.cpp:
2203-2322 (LLM hidden sizes x {SM/2, 2048} rows), benchmarks/python/test_layernorm_fwd.py:90-95
(validate before timing).

y / dx rows are independent of each other: at 10^8-element sizes they are compared with the CPU oracles on
a row sample (first / middle / last blocks + random rows) and, for EVERY row, with a torch evaluation on
the GPU (a loose all-rows sweep that would catch a row that was skipped or written twice); dweight /
dbias, which depend on every row, are compared with the full-size ATen oracle and an fp64 sum over all rows.
"""
import pytest
import torch

from fuser_b200 import _lib as L
from fuser_b200 import ops
from oracle import aten_oracle as AO
from oracle import c_oracle as CO
from tests.util import col_sums_fp64, sample_rows, ulp_err

pytestmark = pytest.mark.gpu

TIGHT = {torch.bfloat16: 1.01, torch.float16: 1.01, torch.float32: 64.0}
EPS = 1e-5
BWD_FLAGS = L.F_HAS_WEIGHT | L.F_OUT_DX | L.F_OUT_DW


def gpu_randn(m, n, dt, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(m, n, device="cuda", generator=g, dtype=torch.float32).to(dt)


def check(actual, aten_expected, exact, red_size, what, floor):
    AO.assert_close(actual, aten_expected, red_size, what)
    u = ulp_err(actual.cpu(), exact, floor=floor)
    assert u <= TIGHT[actual.dtype], f"{what}: {u:.2f} ulp vs fp64"


def all_rows_close(actual, ref_f32, atol, rtol, what):
    """Loose sweep over EVERY row against a torch evaluation on the GPU."""
    err = (actual.float() - ref_f32).abs() - rtol * ref_f32.abs()
    bad = (err > atol).sum().item()
    assert bad == 0, f"{what}: {bad} elements off by more than atol {atol} + rtol {rtol}"


# ---------------------------------------------------------------------------------------------------
# RMSNorm forward + backward: C2 (two-row tiles, cached backward) and C5 (cluster of 4, pipelined)
# ---------------------------------------------------------------------------------------------------
def _rms_fwd_bwd(m, n, dt, seed, expect_fwd, expect_bwd):
    fplan = ops.plan_for(L.OP_RMSNORM_FWD, dt, m, n, L.F_HAS_WEIGHT | L.F_STAT_IS_RMS, EPS)
    bplan = ops.plan_for(L.OP_RMSNORM_BWD, dt, m, n, BWD_FLAGS | L.F_STAT_IS_RMS)
    for k, v in expect_fwd.items():
        assert fplan[k] == v, f"forward plan {k}: {fplan[k]} (expected {v}) {fplan}"
    for k, v in expect_bwd.items():
        assert bplan[k] == v, f"backward plan {k}: {bplan[k]} (expected {v}) {bplan}"
    x = gpu_randn(m, n, dt, seed)
    dy = gpu_randn(m, n, dt, seed + 1)
    w = gpu_randn(1, n, dt, seed + 2).view(n)
    y, rms = ops.rms_norm_fwd(x, w, EPS, stat_is_rms=True)
    dx, dw = ops.rms_norm_bwd(dy, x, rms, w, stat_is_rms=True, centered=False)
    torch.cuda.synchronize()
    # every row, loosely, against torch on the GPU
    xf = x.float()
    rms_ref = torch.sqrt(xf.pow(2).mean(-1, keepdim=True) + EPS)
    all_rows_close(rms, rms_ref, 0.0, 1e-5, "rms (all rows)")
    all_rows_close(y, xf / rms_ref * w.float(), 2e-2, 2e-2, "y (all rows)")
    g = dy.float() * w.float()
    dx_ref = g / rms_ref - xf * (g * xf).sum(-1, keepdim=True) / (n * rms_ref.pow(3))
    all_rows_close(dx, dx_ref, 3e-2, 2e-2, "dx (all rows)")
    del xf, g, dx_ref
    # the oracles
    xc, dyc, wc, rmsc = x.cpu(), dy.cpu(), w.cpu(), rms.cpu()
    rows = sample_rows(m)
    ye, rmse = AO.rms_norm_fwd_python(xc[rows], wc, EPS)
    y64, rms64 = CO.rms_norm_fwd(xc[rows], wc, EPS, stat_is_rms=True)
    check(y[rows.cuda()], ye, y64, n, "y", 1e-2)
    check(rms[rows.cuda()], rmse, rms64, n, "rms", 1e-3)
    dxe, _ = AO.rms_norm_bwd_python(xc[rows], rmsc[rows], dyc[rows], wc)
    dx64, _, _ = CO.norm_bwd(dyc[rows], xc[rows], None, rmsc[rows].view(-1), wc, stat_is_rms=True, centered=False,
                             want_db=False)
    check(dx[rows.cuda()], dxe, dx64, n, "dx", 1e-1)
    _, dwe = AO.rms_norm_bwd_python(xc, rmsc, dyc, wc)  # all rows: the ATen evaluation of the whole fusion
    dw64, _ = col_sums_fp64(dyc, xc, None, 1.0 / rmsc.double())
    check(dw, dwe, dw64, m, "dw", max(1.0, m ** 0.5))


def test_c2_rmsnorm_fwd_bwd_full_size():
    """BASELINE configs[1]: the plan bench.py times (backward with two rows per tile, cached form)."""
    _rms_fwd_bwd(16384, 4096, torch.bfloat16, 20, {"variant": "tma", "cluster_size": 1},
                 {"variant": "tma", "cluster_size": 1, "tile_rows": 2, "cached": True})


def test_c5_rmsnorm_fwd_bwd_full_size():
    """BASELINE configs[4]: rows split over a cluster of 4, pipelined forward and backward, 31 rows per cluster."""
    _rms_fwd_bwd(4096, 65536, torch.bfloat16, 50, {"variant": "cluster", "cluster_size": 4, "pipe": True},
                 {"variant": "cluster", "cluster_size": 4, "pipe": True, "cached": False})


def test_rmsnorm_wide_rows_one_group():
    """N = 24576 bf16: six chunks per thread of one 512-thread group — the pipelined forward without a cluster, ten rows
    per CTA; the backward splits the row over a cluster of two (re-read form, pipelined)."""
    _rms_fwd_bwd(1500, 24576, torch.bfloat16, 60, {"variant": "tma", "cluster_size": 1, "pipe": True, "chunks": 6},
                 {"variant": "cluster", "cluster_size": 2, "pipe": True})


# ---------------------------------------------------------------------------------------------------
# C3 softmax forward + backward [65536, 8192] bf16
# ---------------------------------------------------------------------------------------------------
def test_c3_softmax_fwd_bwd_full_size():
    m, n, dt = 65536, 8192, torch.bfloat16
    assert ops.plan_for(L.OP_SOFTMAX_FWD, dt, m, n)["variant"] == "tma"
    assert ops.plan_for(L.OP_SOFTMAX_BWD, dt, m, n)["variant"] == "tma"
    x = gpu_randn(m, n, dt, 30)
    y = ops.softmax_fwd(x)
    torch.cuda.synchronize()
    y_ref = torch.softmax(x.float(), -1)
    all_rows_close(y, y_ref, 1e-5, 2e-2, "y (all rows)")
    del y_ref
    rows = sample_rows(m, block=1024, n_random=1024)  # 4096 rows
    rc = rows.cuda()
    xc = x[rc].cpu()
    check(y[rc], AO.softmax_fwd(xc), CO.softmax_fwd(xc), n * n, "softmax y", 1e-6)
    del x
    dy = gpu_randn(m, n, dt, 31)
    dx = ops.softmax_bwd(dy, y)
    dxr = ops.softmax_bwd(dy, y, round_rowsum=True)  # the Python definition's spelling
    torch.cuda.synchronize()
    yf, dyf = y.float(), dy.float()
    dx_ref = yf * dyf - yf * (yf * dyf).sum(-1, keepdim=True)
    all_rows_close(dx, dx_ref, 1e-5, 2e-2, "dx (all rows)")
    del yf, dyf, dx_ref
    yc, dyc = y[rc].cpu(), dy[rc].cpu()
    check(dx[rc], AO.softmax_bwd(dyc, yc), CO.softmax_bwd(dyc, yc), n, "softmax dx", 1e-5)
    AO.assert_close(dxr[rc], AO.softmax_bwd_python(yc, dyc), n, "softmax dx (python spelling)")


# ---------------------------------------------------------------------------------------------------
# C4 LayerNorm backward [32768, 12288] bf16 (dx + dgamma + dbeta)
# ---------------------------------------------------------------------------------------------------
def _ln_bwd(m, n, dt, seed, expect, sample=True):
    plan = ops.plan_for(L.OP_LAYERNORM_BWD, dt, m, n, BWD_FLAGS | L.F_OUT_DB)
    for k, v in expect.items():
        assert plan[k] == v, f"plan {k}: {plan[k]} (expected {v}) {plan}"
    x = gpu_randn(m, n, dt, seed)
    dy = gpu_randn(m, n, dt, seed + 1)
    w = gpu_randn(1, n, dt, seed + 2).view(n)
    _, mean, invstd = ops.layer_norm_fwd(x, None, None, EPS)
    dx, dw, db = ops.layer_norm_bwd(dy, x, mean, invstd, w)
    torch.cuda.synchronize()
    xh = (x.float() - mean) * invstd
    g = dy.float() * w.float()
    dx_ref = invstd * (g - g.mean(-1, keepdim=True) - xh * (g * xh).mean(-1, keepdim=True))
    all_rows_close(dx, dx_ref, 3e-2, 2e-2, "dx (all rows)")
    del xh, g, dx_ref
    xc, dyc, wc, mc, ic = x.cpu(), dy.cpu(), w.cpu(), mean.cpu(), invstd.cpu()
    # the saved statistics are this library's forward outputs: check them too (they feed the oracle below)
    rows = sample_rows(m) if sample else torch.arange(m)
    _, me, ie = AO.layer_norm_fwd(xc[rows], None, None, EPS)
    AO.assert_close(mean[rows.cuda()], me, n, "saved mean")
    AO.assert_close(invstd[rows.cuda()], ie, n, "saved invstd")
    dxe, _, _ = AO.layer_norm_bwd(dyc[rows], xc[rows], mc[rows], ic[rows], wc)
    dx64, _, _ = CO.norm_bwd(dyc[rows], xc[rows], mc[rows].view(-1), ic[rows].view(-1), wc, centered=True)
    check(dx[rows.cuda()], dxe, dx64, n, "ln dx", 1e-1)
    _, dwe, dbe = AO.layer_norm_bwd(dyc, xc, mc, ic, wc)  # every row
    dw64, db64 = col_sums_fp64(dyc, xc, mc, ic)
    cfl = max(1.0, m ** 0.5)
    check(dw, dwe, dw64, m, "ln dw", cfl)
    check(db, dbe, db64, m, "ln db", cfl)
    # the Python / Thunder trace of the same op is matched within the validator's bar as well
    dxp, _, _ = AO.layer_norm_bwd_python(xc[rows], dyc[rows], mc[rows].view(-1), ic[rows], wc)
    AO.assert_close(dx[rows.cuda()], dxp, n, "ln dx vs python trace")


def test_c4_layernorm_bwd_full_size():
    """BASELINE configs[3] on one GPU: one 512-thread group per CTA, re-read form, 222 rows per CTA."""
    _ln_bwd(32768, 12288, torch.bfloat16, 40, {"variant": "tma", "cluster_size": 1, "cached": False, "tile_rows": 1})


# ---------------------------------------------------------------------------------------------------
# steady state of every backward kernel on one GPU (multi-tile shapes, full oracle comparison)
# ---------------------------------------------------------------------------------------------------
def _rms_bwd_small(m, n, dt, seed, expect, centered):
    flags = BWD_FLAGS | (L.F_RMS_BWD_CENTERED if centered else L.F_STAT_IS_RMS)
    plan = ops.plan_for(L.OP_RMSNORM_BWD, dt, m, n, flags)
    for k, v in expect.items():
        assert plan[k] == v, f"plan {k}: {plan[k]} (expected {v}) {plan}"
    x = gpu_randn(m, n, dt, seed)
    dy = gpu_randn(m, n, dt, seed + 1)
    w = gpu_randn(1, n, dt, seed + 2).view(n)
    xc, dyc, wc = x.cpu(), dy.cpu(), w.cpu()
    cfl = max(1.0, m ** 0.5)
    if centered:  # the C++ builder: invstd statistic, keeps the -sum(g) term
        _, rstd = AO.rms_norm_fwd(xc, wc, EPS)
        dx, dw = ops.rms_norm_bwd(dy, x, rstd.cuda(), w, centered=True)
        dxe, dwe = AO.rms_norm_bwd(dyc, xc, rstd, wc)
        dx64, dw64, _ = CO.norm_bwd(dyc, xc, None, rstd.view(-1), wc, centered=True, want_db=False)
    else:  # the Python trace: rms statistic, textbook form
        _, rms = AO.rms_norm_fwd_python(xc, wc, EPS)
        dx, dw = ops.rms_norm_bwd(dy, x, rms.cuda(), w, stat_is_rms=True, centered=False)
        dxe, dwe = AO.rms_norm_bwd_python(xc, rms, dyc, wc)
        dx64, dw64, _ = CO.norm_bwd(dyc, xc, None, rms.view(-1), wc, stat_is_rms=True, centered=False, want_db=False)
    torch.cuda.synchronize()
    check(dx, dxe, dx64, n, "rms dx", 1e-1 if dt != torch.float32 else 1.0)
    check(dw, dwe, dw64, m, "rms dw", cfl)
    return plan


@pytest.mark.parametrize("shape", [(2400, 4096), (5000, 8192)], ids=lambda s: f"{s[0]}x{s[1]}")
def test_two_row_tiles_backward(shape):
    """16-bit norm backward with two rows per tile (the C2 plan) walking several tiles per group."""
    m, n = shape
    plan = _rms_bwd_small(m, n, torch.bfloat16, 60 + n, {"variant": "tma", "tile_rows": 2, "cached": True}, centered=False)
    tiles_per_group = (m // 2) / (plan["grid"] * plan["groups"])
    assert tiles_per_group > plan["stages"] / plan["groups"], f"{tiles_per_group} tiles per group: no refill exercised"
    _rms_bwd_small(m, n, torch.bfloat16, 61 + n, {"variant": "tma", "tile_rows": 2}, centered=True)
    _ln_bwd(m, n, torch.bfloat16, 62 + n, {"variant": "tma", "tile_rows": 2, "cached": True}, sample=False)
    _ln_bwd(m, n, torch.float16, 63 + n, {"variant": "tma", "tile_rows": 2, "cached": True}, sample=False)


def test_layernorm_bwd_reread_steady_state():
    """N = 12288 (the C4 kernel: one 512-thread group, re-read form): ~10 rows per CTA, ring wraps."""
    _ln_bwd(1500, 12288, torch.bfloat16, 70, {"variant": "tma", "cached": False, "chunks": 3}, sample=False)


def test_cluster_backward_steady_state():
    """N = 65536 over a cluster of 4 (the C5 kernels): 21 rows per cluster > ring depth, slot sets and cluster
    barriers alternate many times."""
    m, n = 700, 65536
    plan = _rms_bwd_small(m, n, torch.bfloat16, 80, {"variant": "cluster", "cluster_size": 4, "pipe": True}, centered=False)
    nclusters = plan["grid"] // plan["cluster_size"]
    assert m / nclusters > plan["stages"], "tiles per cluster must exceed the ring depth"
    _rms_bwd_small(m, n, torch.bfloat16, 81, {"variant": "cluster", "cluster_size": 4}, centered=True)
    _ln_bwd(m, n, torch.bfloat16, 82, {"variant": "cluster"}, sample=False)
    # forward over the same cluster, many rows per cluster
    x = gpu_randn(m, n, torch.bfloat16, 83)
    w = gpu_randn(1, n, torch.bfloat16, 84).view(n)
    y, st = ops.rms_norm_fwd(x, w, EPS)
    ye, se = AO.rms_norm_fwd(x.cpu(), w.cpu(), EPS)
    y64, s64 = CO.rms_norm_fwd(x.cpu(), w.cpu(), EPS)
    check(y, ye, y64, n, "cluster rms y", 1e-2)
    check(st, se, s64, n, "cluster rms invstd", 1e-3)
    y, mean, invstd = ops.layer_norm_fwd(x, w, w, EPS)
    ye, me, ie = AO.layer_norm_fwd(x.cpu(), w.cpu(), w.cpu(), EPS)
    y64, m64, i64 = CO.layer_norm_fwd(x.cpu(), w.cpu(), w.cpu(), EPS)
    check(y, ye, y64, n, "cluster ln y", 1e-2)
    check(mean, me, m64, n, "cluster ln mean", 1.0)
    check(invstd, ie, i64, n, "cluster ln invstd", 1e-3)


def test_fp32_pipelined_backward_steady_state():
    """fp32 rows of 32 KB: the pipelined re-read backward on a single CTA per row, ~20 rows per CTA."""
    m, n = 3000, 8192
    _rms_bwd_small(m, n, torch.float32, 90, {"variant": "tma", "pipe": True, "cached": False}, centered=False)
    _ln_bwd(m, 6144, torch.float32, 91, {"variant": "tma", "pipe": True, "cached": False, "chunks": 3}, sample=False)
    _ln_bwd(m, n, torch.float32, 92, {"variant": "cluster", "cluster_size": 2, "cached": True}, sample=False)


# ---------------------------------------------------------------------------------------------------
# LayerNorm forward: the pipelined one-round kernel (Welford pairs + Chan merge) in its steady state
# ---------------------------------------------------------------------------------------------------
def _ln_fwd(m, n, dt, seed, expect, residual=False, shift=0.0):
    flags = L.F_HAS_WEIGHT | L.F_HAS_BIAS | L.F_OUT_MEAN | L.F_OUT_INVSTD
    if residual:
        flags |= L.F_ADD_RESIDUAL | L.F_OUT_SUM
    plan = ops.plan_for(L.OP_LAYERNORM_FWD, dt, m, n, flags, EPS)
    for k, v in expect.items():
        assert plan[k] == v, f"plan {k}: {plan[k]} (expected {v}) {plan}"
    x = gpu_randn(m, n, dt, seed)
    if shift:
        x = (x.float() * 0.5 + shift).to(dt)  # |mean| >> std: where a one-pass E[x^2] - mean^2 would lose the variance
    w = gpu_randn(1, n, dt, seed + 1).view(n)
    b = gpu_randn(1, n, dt, seed + 2).view(n)
    xc, wc, bc = x.cpu(), w.cpu(), b.cpu()
    fl = 1.0 if dt == torch.float32 else 1e-2
    if residual:
        r = gpu_randn(m, n, dt, seed + 3)
        y, mean, invstd, s = ops.layer_norm_fwd(x, w, b, EPS, residual=r, want_sum=True, round_sum=dt != torch.float32)
        seen, se = AO.residual_sum(xc, r.cpu(), dt != torch.float32)  # (what the norm sees, the sum in the io dtype)
        assert torch.equal(s.cpu(), se), "x + residual must be bit-exact"
        ye, me, ie, _ = AO.add_layer_norm_fwd(xc, r.cpu(), wc, bc, EPS, round_sum=dt != torch.float32)
        y64, m64, i64 = CO.layer_norm_fwd(seen, wc, bc, EPS)
    else:
        y, mean, invstd = ops.layer_norm_fwd(x, w, b, EPS)
        ye, me, ie = AO.layer_norm_fwd(xc, wc, bc, EPS)
        y64, m64, i64 = CO.layer_norm_fwd(xc, wc, bc, EPS)
    torch.cuda.synchronize()
    check(y, ye, y64, n, "ln y", fl)
    check(mean, me, m64, n, "ln mean", 1.0)
    check(invstd, ie, i64, n, "ln invstd", 1e-3)
    return plan


def test_layernorm_fwd_pipelined_steady_state(monkeypatch):
    pipe1 = {"variant": "tma", "cluster_size": 1, "pipe": True, "groups": 1}
    _ln_fwd(1500, 12288, torch.bfloat16, 100, dict(pipe1, chunks=3))   # the [32768,12288] kernel, ~10 rows per CTA
    _ln_fwd(1500, 12288, torch.float16, 101, dict(pipe1, chunks=3))
    _ln_fwd(1500, 12288, torch.bfloat16, 102, dict(pipe1, chunks=3), shift=20.0)
    _ln_fwd(1500, 16384, torch.bfloat16, 103, dict(pipe1, chunks=4))   # packed parameters
    _ln_fwd(3000, 8192, torch.float32, 104, dict(pipe1, chunks=4))
    _ln_fwd(3000, 8192, torch.float32, 105, dict(pipe1, chunks=4), shift=3.0)
    _ln_fwd(1500, 12288, torch.bfloat16, 106, dict(pipe1, chunks=3), residual=True)
    _ln_fwd(1000, 8192, torch.float32, 107, dict(pipe1, chunks=4), residual=True)
    plan = _ln_fwd(700, 65536, torch.bfloat16, 108, {"variant": "cluster", "cluster_size": 4, "pipe": True})
    assert 700 / (plan["grid"] // 4) > plan["stages"]
    _ln_fwd(300, 131072, torch.bfloat16, 109, {"variant": "cluster", "pipe": True})
    # five / six chunks per thread at 512 threads: no cluster for N = 20480 / 24576 (16 bits), 10240 / 12288 (fp32)
    _ln_fwd(1500, 24576, torch.bfloat16, 120, dict(pipe1, chunks=6, threads_per_row=512))
    _ln_fwd(1500, 24576, torch.bfloat16, 121, dict(pipe1, chunks=6), shift=20.0)
    _ln_fwd(1500, 20480, torch.float16, 122, dict(pipe1, chunks=5, threads_per_row=512))
    _ln_fwd(1500, 12288, torch.float32, 123, dict(pipe1, chunks=6, threads_per_row=512))
    _ln_fwd(1500, 10240, torch.float32, 124, dict(pipe1, chunks=5), shift=3.0)
    # multi-group shapes (every threads-per-row class of the table)
    for m, n, dt in [(3001, 4096, torch.bfloat16), (3001, 4096, torch.float32), (2000, 2048, torch.bfloat16),
                     (2000, 1024, torch.float32), (2000, 8192, torch.bfloat16), (2000, 3072, torch.float16)]:
        plan = _ln_fwd(m, n, dt, 110 + n, {"variant": "tma", "pipe": True})
        assert plan["groups"] > 1
    monkeypatch.setenv("B200NORM_NO_LN_PIPE", "1")
    _ln_fwd(1500, 12288, torch.bfloat16, 100, {"variant": "tma", "pipe": False})

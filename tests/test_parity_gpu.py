"""GPU parity tests (run with -m gpu on a B200): every kernel family of libb200norm.so,
called through the C ABI, against
  (a) the ATen oracle with the reference validator's tolerances (testValidate's bar), and
  (b) the independent fp64 C oracle with a tight bound in ulps of the io dtype.
Shapes follow the reference's suites: LLM hidden sizes (tests/cpp/utils.h:781-795), the
LayerNorm-backward sweep incl. 3 / 1987 (test_combined_inner_outer_reduction.cpp:143-166),
odd sizes N+2.. (benchmarks/python/global_params.py:86-88), wide rows for the cluster path.
"""
import pytest
import torch

from fuser_b200 import _lib as L
from fuser_b200 import ops
from oracle import aten_oracle as AO
from oracle import c_oracle as CO
from tests.util import golden_files, load_golden, make_inputs, ulp_err

pytestmark = pytest.mark.gpu

DTYPES = [torch.bfloat16, torch.float32, torch.float16]
# (rows, cols) — small enough for the CPU oracles, wide enough to hit every threads-per-row class
SHAPES = [(216, 3), (5, 32), (33, 96), (74, 768), (37, 1024), (19, 1987), (64, 2048), (148, 4096), (21, 4098),
          (12, 8192), (9, 12288), (5, 18432), (3, 65536)]
VARIANTS = [L.VARIANT_AUTO, L.VARIANT_GENERIC]
# ulp budgets vs exact arithmetic: 16-bit outputs are dominated by the final rounding
TIGHT = {torch.bfloat16: 1.01, torch.float16: 1.01, torch.float32: 64.0}


def dev(*ts):
    return [None if t is None else t.cuda() for t in ts]


def check(actual, aten_expected, exact, red_size, what, floor):
    torch.cuda.synchronize()
    AO.assert_close(actual, aten_expected, red_size, what)
    u = ulp_err(actual.cpu(), exact, floor=floor)
    assert u <= TIGHT[actual.dtype], f"{what}: {u:.2f} ulp vs fp64"


@pytest.fixture(autouse=True)
def _reset_variant():
    yield
    ops.force_variant(L.VARIANT_AUTO, 0)


@pytest.mark.parametrize("variant", VARIANTS, ids=["auto", "generic"])
@pytest.mark.parametrize("dt", DTYPES, ids=["bf16", "f32", "f16"])
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: f"{s[0]}x{s[1]}")
def test_forward_ops(shape, dt, variant):
    m, n = shape
    x, dy, w, b = make_inputs(m, n, dt, seed=m * 7 + n)
    xd, wd, bd = dev(x, w, b)
    eps = 1e-5
    fl = 1.0 if dt == torch.float32 else 1e-2
    # RMSNorm (invstd flavour, C++ builder)
    y, stat = ops.rms_norm_fwd(xd, wd, eps, variant=variant)
    ye, se = AO.rms_norm_fwd(x, w, eps)
    yc, sc = CO.rms_norm_fwd(x, w, eps)
    check(y, ye, yc, n, "rms y", fl)
    check(stat, se, sc, n, "rms invstd", 1e-3)
    # RMSNorm (rms flavour, Python/Thunder trace), no weight
    y, stat = ops.rms_norm_fwd(xd, None, eps, stat_is_rms=True, variant=variant)
    yc, sc = CO.rms_norm_fwd(x, None, eps, stat_is_rms=True)
    check(y, AO.rms_norm_fwd(x, None, eps)[0], yc, n, "rms(py) y", fl)
    check(stat, AO.rms_norm_fwd_python(x, torch.ones_like(w), eps)[1], sc, n, "rms(py) stat", 1e-3)
    # LayerNorm
    y, mean, invstd = ops.layer_norm_fwd(xd, wd, bd, eps, variant=variant)
    ye, me, ie = AO.layer_norm_fwd(x, w, b, eps)
    yc, mc, ic = CO.layer_norm_fwd(x, w, b, eps)
    check(y, ye, yc, n, "ln y", fl)
    check(mean, me, mc, n, "ln mean", 1.0)
    check(invstd, ie, ic, n, "ln invstd", 1e-3)
    # softmax / log-softmax
    y = ops.softmax_fwd(xd, variant=variant)
    check(y, AO.softmax_fwd(x), CO.softmax_fwd(x), n * n, "softmax y", 1e-6 if dt != torch.float32 else 1e-3)
    y = ops.softmax_fwd(xd, log=True, variant=variant)
    check(y, AO.log_softmax_fwd(x), CO.softmax_fwd(x, log=True), n * n, "log_softmax y", 1.0)


@pytest.mark.parametrize("variant", VARIANTS, ids=["auto", "generic"])
@pytest.mark.parametrize("dt", DTYPES, ids=["bf16", "f32", "f16"])
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: f"{s[0]}x{s[1]}")
def test_backward_ops(shape, dt, variant):
    m, n = shape
    x, dy, w, b = make_inputs(m, n, dt, seed=m * 11 + n)
    eps = 1e-5
    _, mean, invstd = AO.layer_norm_fwd(x, w, b, eps)
    _, rstd = AO.rms_norm_fwd(x, w, eps)
    _, rms = AO.rms_norm_fwd_python(x, w, eps)
    xd, dyd, wd, md, isd, rstdd, rmsd = dev(x, dy, w, mean, invstd, rstd, rms)
    fl = 1.0 if dt == torch.float32 else 1e-1
    cfl = max(1.0, m ** 0.5)  # column sums: absolute accuracy scales with sqrt(rows)
    # LayerNorm backward (C++ builder form)
    dx, dw, db = ops.layer_norm_bwd(dyd, xd, md, isd, wd, variant=variant)
    dxe, dwe, dbe = AO.layer_norm_bwd(dy, x, mean, invstd, w)
    dxc, dwc, dbc = CO.norm_bwd(dy, x, mean.view(-1), invstd.view(-1), w, centered=True)
    check(dx, dxe, dxc, n, "ln dx", fl)
    check(dw, dwe, dwc, m, "ln dw", cfl)
    check(db, dbe, dbc, m, "ln db", cfl)
    # Python/Thunder trace of the same op must also be matched within the validator's bar
    dxp, dwp, dbp = AO.layer_norm_bwd_python(x, dy, mean.view(-1), invstd, w)
    AO.assert_close(dx, dxp, n, "ln dx vs python trace")
    AO.assert_close(dw, dwp, m, "ln dw vs python trace")
    AO.assert_close(db, dbp, m, "ln db vs python trace")
    # RMSNorm backward, C++ builder (centred, invstd)
    dx, dw = ops.rms_norm_bwd(dyd, xd, rstdd, wd, centered=True, variant=variant)
    dxe, dwe = AO.rms_norm_bwd(dy, x, rstd, w)
    dxc, dwc, _ = CO.norm_bwd(dy, x, None, rstd.view(-1), w, centered=True, want_db=False)
    check(dx, dxe, dxc, n, "rms dx", fl)
    check(dw, dwe, dwc, m, "rms dw", cfl)
    # RMSNorm backward, Python/Thunder trace (textbook, rms statistic)
    dx, dw = ops.rms_norm_bwd(dyd, xd, rmsd, wd, stat_is_rms=True, centered=False, variant=variant)
    dxe, dwe = AO.rms_norm_bwd_python(x, rms, dy, w)
    dxc, dwc, _ = CO.norm_bwd(dy, x, None, rms.view(-1), w, stat_is_rms=True, centered=False, want_db=False)
    check(dx, dxe, dxc, n, "rms(py) dx", fl)
    check(dw, dwe, dwc, m, "rms(py) dw", cfl)
    # softmax backward (both spellings) and log-softmax backward
    sm = AO.softmax_fwd(x)
    lsm = AO.log_softmax_fwd(x)
    smd, lsmd = dev(sm, lsm)
    dx = ops.softmax_bwd(dyd, smd, variant=variant)
    check(dx, AO.softmax_bwd(dy, sm), CO.softmax_bwd(dy, sm), n, "softmax dx", 1e-3 if dt != torch.float32 else 1e-1)
    dx = ops.softmax_bwd(dyd, smd, round_rowsum=True, variant=variant)
    AO.assert_close(dx, AO.softmax_bwd_python(sm, dy), n, "softmax(py) dx")
    dx = ops.softmax_bwd(dyd, lsmd, log=True, variant=variant)
    check(dx, AO.log_softmax_bwd(dy, lsm), CO.softmax_bwd(dy, lsm, log=True), n, "log_softmax dx", 1.0)


@pytest.mark.parametrize("cluster", [2, 4, 8, 16])
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float32], ids=["bf16", "f32"])
@pytest.mark.parametrize("shape", [(40, 4096), (11, 16384), (7, 65536)], ids=lambda s: f"{s[0]}x{s[1]}")
def test_cluster_dsmem_reduction(shape, dt, cluster):
    """Row split over a thread-block cluster with the DSMEM exchange (config C5's mechanism;
    reference: runtime/cluster.cu:223-297, tests/cpp/test_cluster.cpp:38-157)."""
    m, n = shape
    x, dy, w, b = make_inputs(m, n, dt, seed=cluster)
    eps = 1e-6
    xd, dyd, wd, bd = dev(x, dy, w, b)
    kw = dict(variant=L.VARIANT_CLUSTER, cluster_size=cluster)
    fl = 1.0 if dt == torch.float32 else 1e-2
    try:
        for op, fl_ in ((L.OP_RMSNORM_FWD, L.F_HAS_WEIGHT), (L.OP_SOFTMAX_BWD, 0),
                        (L.OP_LAYERNORM_BWD, L.F_HAS_WEIGHT | L.F_OUT_DX | L.F_OUT_DW | L.F_OUT_DB),
                        (L.OP_RMSNORM_BWD, L.F_HAS_WEIGHT | L.F_OUT_DX | L.F_OUT_DW)):
            plan = ops.plan_for(op, dt, m, n, fl_, eps, L.VARIANT_CLUSTER, cluster)
            assert plan["variant"] == "cluster" and plan["cluster_size"] == cluster
    except L.UnsupportedSegment:
        pytest.skip(f"a {n // cluster}-column slice does not fit one CTA's ring for every op")
    y, stat = ops.rms_norm_fwd(xd, wd, eps, **kw)
    ye, se = AO.rms_norm_fwd(x, w, eps)
    check(y, ye, CO.rms_norm_fwd(x, w, eps)[0], n, "rms y", fl)
    check(stat, se, CO.rms_norm_fwd(x, w, eps)[1], n, "rms invstd", 1e-3)
    y, mean, invstd = ops.layer_norm_fwd(xd, wd, bd, eps, **kw)
    ye, me, ie = AO.layer_norm_fwd(x, w, b, eps)
    check(y, ye, CO.layer_norm_fwd(x, w, b, eps)[0], n, "ln y", fl)
    check(ops.softmax_fwd(xd, **kw), AO.softmax_fwd(x), CO.softmax_fwd(x), n * n, "softmax y",
          1e-6 if dt != torch.float32 else 1e-3)
    sm = AO.softmax_fwd(x)
    check(ops.softmax_bwd(dyd, sm.cuda(), **kw), AO.softmax_bwd(dy, sm), CO.softmax_bwd(dy, sm), n, "softmax dx", 1e-3)
    dx, dw, db = ops.layer_norm_bwd(dyd, xd, me.cuda(), ie.cuda(), wd, **kw)
    dxc, dwc, dbc = CO.norm_bwd(dy, x, me.view(-1), ie.view(-1), w, centered=True)
    dxe, dwe, dbe = AO.layer_norm_bwd(dy, x, me, ie, w)
    check(dx, dxe, dxc, n, "ln dx", 1.0 if dt == torch.float32 else 1e-1)
    check(dw, dwe, dwc, m, "ln dw", max(1.0, m ** 0.5))
    check(db, dbe, dbc, m, "ln db", max(1.0, m ** 0.5))
    dx, dw = ops.rms_norm_bwd(dyd, xd, se.cuda(), wd, centered=False, **kw)
    dxc, dwc, _ = CO.norm_bwd(dy, x, None, se.view(-1), w, centered=False, want_db=False)
    AO.assert_close(dx, dxc.to(dt), n, "rms dx")
    assert ulp_err(dx.cpu(), dxc, floor=1.0 if dt == torch.float32 else 1e-1) <= TIGHT[dt]
    assert ulp_err(dw.cpu(), dwc, floor=max(1.0, m ** 0.5)) <= TIGHT[dt]


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: p.split("/")[-1][:-4])
def test_kernels_match_golden_fixtures(path):
    """The committed ATen-native expectations (tests/golden/make_golden.py), replayed on the GPU."""
    dt, m, n, g = load_golden(path)
    eps = g["eps"]
    x, dy, w, b = dev(g["x"], g["dy"], g["w"], g["b"])
    y, mean, invstd = ops.layer_norm_fwd(x, w, b, eps)
    AO.assert_close(y, g["ln_y"], n, "ln y")
    AO.assert_close(mean, g["ln_mean"], n, "ln mean")
    AO.assert_close(invstd, g["ln_invstd"], n, "ln invstd")
    dx, dw, db = ops.layer_norm_bwd(dy, x, g["ln_mean"].cuda(), g["ln_invstd"].cuda(), w)
    AO.assert_close(dx, g["ln_dx"], n, "ln dx")
    AO.assert_close(dw, g["ln_dw"], m, "ln dw")
    AO.assert_close(db, g["ln_db"], m, "ln db")
    y, invstd = ops.rms_norm_fwd(x, w, eps)
    AO.assert_close(y, g["rms_y"], n, "rms y")
    AO.assert_close(invstd, g["rms_invstd"], n, "rms invstd")
    dx, dw = ops.rms_norm_bwd(dy, x, g["rms_invstd"].cuda(), w, centered=False)
    AO.assert_close(dx, g["rms_dx"], n, "rms dx")
    AO.assert_close(dw, g["rms_dw"], m, "rms dw")
    AO.assert_close(ops.softmax_fwd(x), g["sm_y"], n * n, "softmax y")
    AO.assert_close(ops.softmax_bwd(dy, g["sm_y"].cuda()), g["sm_dx"], n, "softmax dx")
    AO.assert_close(ops.softmax_fwd(x, log=True), g["lsm_y"], n * n, "log_softmax y")
    AO.assert_close(ops.softmax_bwd(dy, g["lsm_y"].cuda(), log=True), g["lsm_dx"], n, "log_softmax dx")


def test_edge_cases():
    """Empty tensors, a single row, strided rows, NaN/Inf rows, 3-D inputs, output masks."""
    dt = torch.bfloat16
    # empty iteration domain
    x = torch.empty(0, 256, dtype=dt, device="cuda")
    y, stat = ops.rms_norm_fwd(x, None, 1e-5)
    assert y.shape == (0, 256) and stat.shape == (0, 1)
    dx, dw, db = ops.layer_norm_bwd(x, x, stat, stat, torch.ones(256, dtype=dt, device="cuda"))
    torch.cuda.synchronize()
    assert dx.shape == (0, 256) and float(dw.abs().sum()) == 0.0 and float(db.abs().sum()) == 0.0
    # strided rows (a column slice of a wider buffer): rows are not contiguous
    big = torch.randn(33, 3 * 1024, device="cuda").to(dt)
    xs = big[:, 1024:2048]
    y, stat = ops.rms_norm_fwd(xs, None, 1e-5)
    ye, se = AO.rms_norm_fwd(xs.cpu().contiguous(), None, 1e-5)
    check(y, ye, CO.rms_norm_fwd(xs.cpu(), None, 1e-5)[0], 1024, "strided rms y", 1e-2)
    # misaligned base pointer -> generic family, same numbers
    flat = torch.randn(17 * 512 + 1, device="cuda").to(dt)
    xm = flat[1:].view(17, 512)
    y, _ = ops.rms_norm_fwd(xm, None, 1e-5)
    check(y, AO.rms_norm_fwd(xm.cpu(), None, 1e-5)[0], CO.rms_norm_fwd(xm.cpu(), None, 1e-5)[0], 512, "misaligned rms y", 1e-2)
    # 3-D input normalised over the last dim
    x3 = torch.randn(4, 9, 768, device="cuda").to(dt)
    w = torch.randn(768, device="cuda").to(dt)
    b = torch.randn(768, device="cuda").to(dt)
    y, mean, invstd = ops.layer_norm_fwd(x3, w, b, 1e-5)
    ye, me, ie = AO.layer_norm_fwd(x3.cpu(), w.cpu(), b.cpu(), 1e-5)
    assert y.shape == x3.shape and mean.shape == (4, 9, 1)
    AO.assert_close(y, ye, 768, "3d ln y")
    AO.assert_close(mean, me, 768, "3d ln mean")
    # NaN / Inf rows in softmax: whole row NaN exactly as ATen gives (equal_nan)
    xs = torch.randn(8, 512, device="cuda").to(dt)
    xs[1, 7] = float("nan")
    xs[2, :] = float("-inf")
    xs[3, 5] = float("inf")
    y = ops.softmax_fwd(xs)
    AO.assert_close(y, AO.softmax_fwd(xs.cpu()), 512 * 512, "softmax with NaN/Inf")
    assert torch.isnan(y[1:4]).all() and not torch.isnan(y[0]).any()
    # output masks: dx only / dw only
    x, dy, w, b = make_inputs(64, 1024, dt, seed=5)
    _, me, ie = AO.layer_norm_fwd(x, w, b, 1e-5)
    xd, dyd, wd, med, ied = dev(x, dy, w, me, ie)
    dx, dw, db = ops.layer_norm_bwd(dyd, xd, med, ied, wd, output_mask=(True, False, False))
    assert dw is None and db is None
    AO.assert_close(dx, AO.layer_norm_bwd(dy, x, me, ie, w)[0], 1024, "dx only")
    dx, dw, db = ops.layer_norm_bwd(dyd, xd, med, ied, wd, output_mask=(False, True, True))
    assert dx is None
    AO.assert_close(dw, AO.layer_norm_bwd(dy, x, me, ie, w)[1], 64, "dw only")
    AO.assert_close(db, AO.layer_norm_bwd(dy, x, me, ie, w)[2], 64, "db only")
    # fp32 partial sums + finalize == direct path (the multi-GPU composition on one device)
    dx, sums = ops.layer_norm_bwd(dyd, xd, med, ied, wd, partials_f32=True)
    dw2, db2 = ops.finalize_partials(sums, dt)
    _, dw1, db1 = ops.layer_norm_bwd(dyd, xd, med, ied, wd)
    torch.cuda.synchronize()
    assert torch.equal(dw1, dw2) and torch.equal(db1, db2)


def test_reductions():
    for dt in (torch.float32, torch.bfloat16):
        x = torch.randn(67, 1500).to(dt)
        xd = x.cuda()
        for kind in ("sum", "max", "fmax"):
            r = ops.reduce_inner(xd, kind)
            e = CO.reduce_inner(x, kind)
            torch.cuda.synchronize()
            assert ulp_err(r.cpu(), e, floor=40.0 if kind == "sum" else 0.0) <= TIGHT[dt]
        r = ops.reduce_outer(xd)
        torch.cuda.synchronize()
        assert ulp_err(r.cpu(), CO.reduce_outer(x), floor=10.0) <= TIGHT[dt]
    x = torch.randn(4, 64)
    x[0, 3] = float("nan")
    x[1, :] = float("nan")
    r_max = ops.reduce_inner(x.cuda(), "max").cpu()
    r_fmax = ops.reduce_inner(x.cuda(), "fmax").cpu()
    assert torch.isnan(r_max[0]) and torch.isnan(r_max[1]) and not torch.isnan(r_max[2])
    assert not torch.isnan(r_fmax[0]) and torch.isnan(r_fmax[1])


def test_full_size_properties():
    """BASELINE.json-size checks through size-independent properties: softmax rows sum to 1,
    normalised rows have unit RMS, dx of softmax sums to 0, column sums are linear in dy."""
    dt = torch.bfloat16
    m, n = 16384, 4096
    x = torch.randn(m, n, device="cuda").to(dt)
    y, stat = ops.rms_norm_fwd(x, None, 1e-5)
    rms_y = y.float().pow(2).mean(-1).sqrt()
    assert (rms_y - 1).abs().max().item() < 5e-3
    assert torch.allclose(stat.view(-1), torch.rsqrt(x.float().pow(2).mean(-1) + 1e-5), rtol=1e-5)
    m, n = 8192, 8192
    x = torch.randn(m, n, device="cuda").to(dt)
    sm = ops.softmax_fwd(x)
    assert (sm.float().sum(-1) - 1).abs().max().item() < 2e-2
    dy = torch.randn(m, n, device="cuda").to(dt)
    dx = ops.softmax_bwd(dy, sm)
    assert dx.float().sum(-1).abs().max().item() < 2e-2
    m, n = 4096, 12288
    x = torch.randn(m, n, device="cuda").to(dt)
    dy = torch.randn(m, n, device="cuda").to(dt)
    w = torch.randn(n, device="cuda").to(dt)
    _, mean, invstd = ops.layer_norm_fwd(x, None, None, 1e-5)
    dx, sums1 = ops.layer_norm_bwd(dy, x, mean, invstd, w, partials_f32=True)
    _, sums2 = ops.layer_norm_bwd((dy.float() * 2).to(dt), x, mean, invstd, w, partials_f32=True)
    torch.cuda.synchronize()
    assert torch.allclose(sums2, 2 * sums1, rtol=1e-5, atol=1e-3)  # exact doubling in bf16 => linear
    assert torch.allclose(sums1[1], dy.float().sum(0), rtol=1e-4, atol=1e-2)
    # dx rows are orthogonal to 1 and to x_hat (projection property of LayerNorm backward)
    # (bf16 rounding of 12288 outputs of magnitude ~1-4 leaves a residual of a few units)
    assert dx.float().sum(-1).abs().max().item() < 4.0

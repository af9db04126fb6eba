"""Plain reductions and stand-alone statistics on the fast path (run with -m gpu): the Reduction scheduler's
segments (csrc/scheduler/reduction.cpp:198-360 — inner sum / max / fmax, the outer column sum that a segmented wide
backward leaves behind) and variance_mean / Welford (csrc/ops/normalization.cpp:93-158, runtime/welford.cu:52-221),
each on the TMA ring kernels and on the shape-agnostic family, against the ATen evaluation (validator tolerances) and
fp64 (<= 1.01 ulp of the io dtype / 64 ulp fp32), called through ops and through FusionDefinition."""
import pytest
import torch

from fuser_b200 import _lib as L
from fuser_b200 import ops
from fuser_b200.frontend import DataType, FusionDefinition, torch_dtype_to_nvfuser_dtype
from oracle import aten_oracle as AO
from oracle import c_oracle as CO
from tests.util import ulp_err

pytestmark = pytest.mark.gpu

DTYPES = [torch.bfloat16, torch.float32, torch.float16]
TIGHT = {torch.bfloat16: 1.01, torch.float16: 1.01, torch.float32: 64.0}
SHAPES = [(67, 1500), (5, 32), (300, 1024), (1000, 4096), (64, 12288), (9, 18432), (40, 65536), (2000, 8192), (3, 262144)]


def _x(m, n, dt, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(m, n, generator=g).to(dt)


@pytest.mark.parametrize("variant", [L.VARIANT_AUTO, L.VARIANT_GENERIC], ids=["auto", "generic"])
@pytest.mark.parametrize("dt", DTYPES, ids=["bf16", "f32", "f16"])
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: f"{s[0]}x{s[1]}")
def test_reduce_inner(shape, dt, variant):
    m, n = shape
    x = _x(m, n, dt, m + n)
    xd = x.cuda()
    for kind in ("sum", "max", "fmax"):
        r = ops.reduce_inner(xd, kind, variant=variant)
        e64 = CO.reduce_inner(x, kind)
        torch.cuda.synchronize()
        assert r.shape == (m,) and r.dtype == dt
        if kind != "fmax":
            AO.assert_close(r, AO.reduce_inner(x, kind), n, f"{kind}")
        u = ulp_err(r.cpu(), e64, floor=max(1.0, n ** 0.5) if kind == "sum" else 0.0)
        assert u <= TIGHT[dt], f"{kind}: {u:.2f} ulp"
    if variant == L.VARIANT_AUTO and (n * x.element_size()) % 16 == 0:
        assert ops.plan_for(L.OP_REDUCE_INNER, dt, m, n)["variant"] in ("tma", "cluster")


def test_reduce_inner_nan_semantics():
    for n in (64, 4096, 65536):
        x = torch.randn(8, n)
        x[0, 3] = float("nan")
        x[1, :] = float("nan")
        x[2, n - 1] = float("inf")
        x[3, :] = float("-inf")
        for dt in (torch.float32, torch.bfloat16):
            xd = x.to(dt).cuda()
            r_max = ops.reduce_inner(xd, "max").cpu().float()
            r_fmax = ops.reduce_inner(xd, "fmax").cpu().float()
            r_sum = ops.reduce_inner(xd, "sum").cpu().float()
            assert torch.isnan(r_max[0]) and torch.isnan(r_max[1]) and r_max[2] == float("inf") and r_max[3] == float("-inf")
            assert not torch.isnan(r_fmax[0]) and torch.isnan(r_fmax[1]) and r_fmax[2] == float("inf")
            assert torch.isnan(r_sum[0]) and r_sum[2] == float("inf") and r_sum[3] == float("-inf")
            assert torch.equal(r_max[4:], torch.amax(x.to(dt).float(), -1)[4:])


@pytest.mark.parametrize("variant", [L.VARIANT_AUTO, L.VARIANT_GENERIC], ids=["auto", "generic"])
@pytest.mark.parametrize("dt", DTYPES, ids=["bf16", "f32", "f16"])
@pytest.mark.parametrize("shape", [(67, 1500), (216, 96), (5000, 1024), (3000, 4096), (1500, 12288), (700, 65536), (1, 4096)],
                         ids=lambda s: f"{s[0]}x{s[1]}")
def test_reduce_outer(shape, dt, variant):
    m, n = shape
    x = _x(m, n, dt, 3 * m + n)
    xd = x.cuda()
    r = ops.reduce_outer(xd, variant=variant)
    s32 = ops.reduce_outer(xd, partials_f32=True, variant=variant)
    torch.cuda.synchronize()
    assert r.shape == (n,) and r.dtype == dt and s32.dtype == torch.float32
    AO.assert_close(r, AO.reduce_outer(x), m, "column sum")
    e64 = CO.reduce_outer(x)
    fl = max(1.0, m ** 0.5)
    assert ulp_err(r.cpu(), e64, floor=fl) <= TIGHT[dt]
    assert ulp_err(s32.cpu(), e64, floor=fl) <= 64.0
    assert torch.equal(s32.to(dt), r), "the fp32 sums, cast once, are the io-dtype result"
    r2 = ops.reduce_outer(xd, variant=variant)
    torch.cuda.synchronize()
    assert torch.equal(r, r2), "column sums are deterministic"


@pytest.mark.parametrize("variant", [L.VARIANT_AUTO, L.VARIANT_GENERIC], ids=["auto", "generic"])
@pytest.mark.parametrize("dt", DTYPES, ids=["bf16", "f32", "f16"])
@pytest.mark.parametrize("shape", [(67, 1500), (33, 96), (300, 1024), (1000, 4096), (700, 12288), (40, 65536), (2000, 8192)],
                         ids=lambda s: f"{s[0]}x{s[1]}")
def test_var_mean_and_welford(shape, dt, variant):
    m, n = shape
    x = (_x(m, n, dt, 7 * m + n).float() * 1.5 + 0.75).to(dt)
    xd = x.cuda()
    x64 = x.double()
    for corr in (0, 1, 3):
        var, mean = ops.var_mean(xd, corr, variant=variant)
        torch.cuda.synchronize()
        assert var.shape == (m, 1) and mean.shape == (m, 1) and var.dtype == torch.float32
        ve, me = AO.var_mean(x, corr)
        AO.assert_close(var.view(-1), ve, n, f"var (correction {corr})")
        AO.assert_close(mean.view(-1), me, n, "mean")
        assert ulp_err(var.view(-1).cpu(), x64.var(-1, correction=corr), floor=1e-3) <= 64.0
        assert ulp_err(mean.view(-1).cpu(), x64.mean(-1), floor=1.0) <= 64.0
    only_var, none = ops.var_mean(xd, 1, want_mean=False, variant=variant)
    assert none is None and torch.equal(only_var, ops.var_mean(xd, 1, variant=variant)[0])
    avg, m2 = ops.welford(xd, variant=variant)
    ae, m2e = AO.welford(x)
    torch.cuda.synchronize()
    AO.assert_close(avg.view(-1), ae, n, "welford avg")
    AO.assert_close(m2.view(-1), m2e, n, "welford var_sum")
    assert ulp_err(m2.view(-1).cpu(), x64.var(-1, correction=0) * n, floor=1e-3) <= 64.0


def test_var_mean_degenerate_correction():
    x = torch.randn(4, 8, device="cuda")
    var, mean = ops.var_mean(x, 8)  # N - correction = 0: inf, as at::var_mean
    ve, _ = AO.var_mean(x.cpu(), 8)
    torch.cuda.synchronize()
    assert torch.equal(torch.isinf(var.view(-1).cpu()), torch.isinf(ve)) and torch.isinf(var).all()
    with pytest.raises(ValueError):
        ops.var_mean(x, -1)


def _fd(build):
    with FusionDefinition() as fd:
        build(fd)
    return fd


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float32], ids=["bf16", "f32"])
def test_reductions_through_fusion_definition(dt):
    """The Reduction / Welford segments as a user of nvfuser_direct writes them."""
    nv = torch_dtype_to_nvfuser_dtype(dt)
    m, n = 1200, 4096
    x = _x(m, n, dt, 99)
    xd = x.cuda()

    def up(fd, t):
        return fd.ops.cast(t, DataType.Float) if dt != torch.float32 else t

    def down(fd, t):
        return fd.ops.cast(t, nv) if dt != torch.float32 else t

    for kind, dims, ref in (("sum", [1], AO.reduce_inner(x, "sum")), ("max", [1], AO.reduce_inner(x, "max")),
                            ("sum", [0], AO.reduce_outer(x))):
        def build(fd, kind=kind, dims=dims):
            t = fd.define_tensor(shape=[-1, -1], contiguity=[True, True], dtype=nv)
            r = (fd.ops.sum if kind == "sum" else fd.ops.max)(up(fd, t), dims=dims)
            fd.add_output(down(fd, r))
        fd = _fd(build)
        seg = fd.segment()
        assert seg.op == (L.OP_REDUCE_INNER if dims == [1] else L.OP_REDUCE_OUTER)
        (r,) = fd.execute([xd])
        AO.assert_close(r, ref, n if dims == [1] else m, f"{kind} {dims}")

    def build_vm(fd):
        t = fd.define_tensor(shape=[-1, -1], contiguity=[True, True], dtype=nv)
        var, mean = fd.ops.var_mean(up(fd, t), dims=[1], correction=1, keepdim=False)
        fd.add_output(var)
        fd.add_output(mean)
    fd = _fd(build_vm)
    seg = fd.segment()
    assert seg.op == L.OP_LAYERNORM_FWD and seg.flags & L.F_NO_Y and seg.flags & L.F_STAT_VAR and seg.correction == 1
    var, mean = fd.execute([xd])
    ve, me = AO.var_mean(x, 1)
    assert var.shape == (m,) and mean.shape == (m,)
    AO.assert_close(var, ve, n, "var")
    AO.assert_close(mean, me, n, "mean")

    def build_wf(fd):
        t = fd.define_tensor(shape=[-1, -1], contiguity=[True, True], dtype=nv)
        avg, m2, _ = fd.ops.welford(up(fd, t), [1])
        fd.add_output(avg)
        fd.add_output(m2)
    fd = _fd(build_wf)
    assert fd.segment().flags & L.F_STAT_M2
    avg, m2 = fd.execute([xd])
    ae, m2e = AO.welford(x)
    AO.assert_close(avg, ae, n, "avg")
    AO.assert_close(m2, m2e, n, "var_sum")


def test_reductions_full_size():
    """The shapes of the bench's reduction workload: every row / column against the oracle."""
    for m, n in ((16384, 4096), (32768, 12288)):
        g = torch.Generator(device="cuda").manual_seed(m)
        xd = torch.randn(m, n, device="cuda", generator=g).to(torch.bfloat16)
        x = xd.cpu()
        assert ops.plan_for(L.OP_REDUCE_INNER, torch.bfloat16, m, n)["variant"] == "tma"
        assert ops.plan_for(L.OP_REDUCE_OUTER, torch.bfloat16, m, n)["variant"] == "tma"
        r = ops.reduce_inner(xd, "sum")
        AO.assert_close(r, AO.reduce_inner(x, "sum"), n, "row sums")
        assert ulp_err(r.cpu(), x.double().sum(-1), floor=n ** 0.5) <= 1.01
        r = ops.reduce_inner(xd, "max")
        assert torch.equal(r.cpu(), torch.amax(x, -1))
        c = ops.reduce_outer(xd)
        AO.assert_close(c, AO.reduce_outer(x), m, "column sums")
        assert ulp_err(c.cpu(), x.double().sum(0), floor=m ** 0.5) <= 1.01
        var, mean = ops.var_mean(xd, 1)
        ve, me = AO.var_mean(x, 1)
        AO.assert_close(var.view(-1), ve, n, "var")
        AO.assert_close(mean.view(-1), me, n, "mean")

"""GPU parity of the forward norms with the fused residual add (the prologue
nvFuser's persistent schedulers absorb — benchmarks/python/test_rmsnorm_add_fwd.py,
test_dropout_{layernorm,rmsnorm}_fwd.py without the dropout), through the C ABI, against the ATen
oracle (validator tolerances) and the fp64 C oracle (<= 1.01 ulp of a 16-bit io dtype)."""
import pytest
import torch

from fuser_b200 import _lib as L
from fuser_b200 import ops
from oracle import aten_oracle as AO
from oracle import c_oracle as CO
from tests.util import make_inputs, ulp_err

pytestmark = pytest.mark.gpu

DTYPES = [torch.bfloat16, torch.float32, torch.float16]
SHAPES = [(216, 3), (33, 96), (74, 768), (19, 1987), (300, 2048), (1000, 4096), (21, 4098), (12, 8192), (40, 12288),
          (5, 18432), (3, 65536)]
TIGHT = {torch.bfloat16: 1.01, torch.float16: 1.01, torch.float32: 64.0}


def check(actual, aten_expected, exact, red_size, what, floor):
    torch.cuda.synchronize()
    AO.assert_close(actual, aten_expected, red_size, what)
    u = ulp_err(actual.cpu(), exact, floor=floor)
    assert u <= TIGHT[actual.dtype], f"{what}: {u:.2f} ulp vs fp64"


@pytest.mark.parametrize("variant", [L.VARIANT_AUTO, L.VARIANT_GENERIC], ids=["auto", "generic"])
@pytest.mark.parametrize("round_sum", [False, True], ids=["fp32sum", "roundsum"])
@pytest.mark.parametrize("dt", DTYPES, ids=["bf16", "f32", "f16"])
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: f"{s[0]}x{s[1]}")
def test_add_norm_forward(shape, dt, round_sum, variant):
    m, n = shape
    x, r, w, b = make_inputs(m, n, dt, seed=m * 13 + n)
    xd, rd, wd, bd = (t.cuda() for t in (x, r, w, b))
    eps = 1e-5
    fl = 1.0 if dt == torch.float32 else 1e-2
    s_exact = x.double() + r.double()
    s_io = s_exact.to(torch.float32).to(dt)  # fp32 add of exactly representable operands, then one rounding
    s_seen = s_io.double() if round_sum else s_exact
    # RMSNorm(x + r), sum requested
    y, stat, s = ops.rms_norm_fwd(xd, wd, eps, residual=rd, want_sum=True, round_sum=round_sum, variant=variant)
    ye, se, sume = AO.add_rms_norm_fwd(x, r, w, eps, round_sum)
    yc, sc = CO.rms_norm_fwd(s_seen, w, eps)
    torch.cuda.synchronize()
    assert torch.equal(s.cpu(), sume), "sum output must be the correctly rounded fp32 add (bit-exact)"
    assert torch.equal(s.cpu(), s_io)
    check(y, ye, yc, n, "add+rms y", fl)
    check(stat, se, sc, n, "add+rms invstd", 1e-3)
    # ... without the sum output, rms statistic flavour, no weight
    y2, stat2 = ops.rms_norm_fwd(xd, None, eps, stat_is_rms=True, residual=rd, round_sum=round_sum, variant=variant)
    yc2, sc2 = CO.rms_norm_fwd(s_seen, None, eps, stat_is_rms=True)
    check(y2, AO.add_rms_norm_fwd(x, r, None, eps, round_sum)[0], yc2, n, "add+rms(py) y", fl)
    check(stat2, AO.add_rms_norm_fwd(x, r, None, eps, round_sum, python=True)[1], sc2, n, "add+rms(py) rms", 1e-3)
    # LayerNorm(x + r)
    y, mean, invstd, s = ops.layer_norm_fwd(xd, wd, bd, eps, residual=rd, want_sum=True, round_sum=round_sum,
                                            variant=variant)
    ye, me, ie, sume = AO.add_layer_norm_fwd(x, r, w, b, eps, round_sum)
    yc, mc, ic = CO.layer_norm_fwd(s_seen, w, b, eps)
    torch.cuda.synchronize()
    assert torch.equal(s.cpu(), sume)
    check(y, ye, yc, n, "add+ln y", fl)
    check(mean, me, mc, n, "add+ln mean", 1.0)
    check(invstd, ie, ic, n, "add+ln invstd", 1e-3)
    y3, _, _ = ops.layer_norm_fwd(xd, wd, bd, eps, residual=rd, round_sum=round_sum, variant=variant)
    torch.cuda.synchronize()
    assert torch.equal(y3, y), "the sum output must not change y"


@pytest.mark.parametrize("cluster", [2, 4])
def test_add_norm_cluster(cluster):
    m, n, dt = 9, 32768, torch.bfloat16
    x, r, w, b = make_inputs(m, n, dt, seed=5)
    xd, rd, wd, bd = (t.cuda() for t in (x, r, w, b))
    y, stat, s = ops.rms_norm_fwd(xd, wd, 1e-6, residual=rd, want_sum=True, variant=L.VARIANT_CLUSTER,
                                  cluster_size=cluster)
    ye, se, sume = AO.add_rms_norm_fwd(x, r, w, 1e-6)
    torch.cuda.synchronize()
    assert torch.equal(s.cpu(), sume)
    AO.assert_close(y, ye, n, "cluster add+rms y")
    AO.assert_close(stat, se, n, "cluster add+rms invstd")
    y, mean, invstd, s = ops.layer_norm_fwd(xd, wd, bd, 1e-5, residual=rd, want_sum=True, variant=L.VARIANT_CLUSTER,
                                            cluster_size=cluster)
    ye, me, ie, sume = AO.add_layer_norm_fwd(x, r, w, b, 1e-5)
    torch.cuda.synchronize()
    assert torch.equal(s.cpu(), sume)
    AO.assert_close(y, ye, n, "cluster add+ln y")
    AO.assert_close(mean, me, n, "cluster add+ln mean")
    AO.assert_close(invstd, ie, n, "cluster add+ln invstd")


def test_rmsnorm_add_matches_eager_definition():
    """benchmarks/python/torch_ops.py:76-83 (the op test_rmsnorm_add_fwd.py benchmarks), bf16: the
    round_sum variant agrees with the eager bf16 evaluation within the validator's bf16 bound and
    returns the identical residual stream."""
    m, n = 512, 4096
    x, r, w, _ = make_inputs(m, n, torch.bfloat16, seed=3)
    y, _, s = ops.rms_norm_fwd(x.cuda(), w.cuda(), 1e-5, residual=r.cuda(), want_sum=True, round_sum=True)
    ye, re = AO.rmsnorm_add_eager(x, w, r)
    torch.cuda.synchronize()
    assert torch.equal(s.cpu(), re)
    AO.assert_close(y, ye, n, "rmsnorm_add vs eager")


def test_add_norm_argument_errors():
    x = torch.randn(8, 256, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(ValueError):
        ops.rms_norm_fwd(x, None, 1e-5, want_sum=True)
    with pytest.raises(L.B200NormError):
        ops.rms_norm_fwd(x, None, 1e-5, residual=torch.randn(8, 128, device="cuda", dtype=torch.bfloat16))

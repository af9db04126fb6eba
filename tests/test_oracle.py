"""CPU tests: pin the oracle before anything trusts it.

(1) tolerance rule vs worked examples of the reference's tables,
(2) ATen oracle vs the committed golden fixtures (ATen *native* kernels — the explicit
    expectations the reference's own tests use),
(3) ATen oracle vs the independent fp64 C restatement,
(4) alternative spellings (C++ builder vs Python/Thunder trace) agree with each other.
"""
import math

import pytest
import torch

from oracle import aten_oracle as AO
from oracle import c_oracle as CO
from tests.util import golden_files, load_golden, make_inputs, outer_golden_files, ulp_err


def test_tolerance_rule_worked_examples():
    # csrc/validator_utils.cpp:25-163 with tables validator_utils.h:28-46
    atol, rtol = AO.get_tolerance(torch.float32, 4096)
    assert atol == pytest.approx(0.00012207) and rtol == pytest.approx(0.00012207 * 0.01)
    atol, rtol = AO.get_tolerance(torch.bfloat16, 4096)
    assert atol == pytest.approx(0.24945 * 10) and rtol == pytest.approx(0.24945 * 0.1)
    atol, _ = AO.get_tolerance(torch.float16, 4096)
    assert atol == pytest.approx(0.24945)
    # past the end of the table: twice the last entry (x10 for bf16)
    atol, _ = AO.get_tolerance(torch.bfloat16, 8192 * 8192)
    assert atol == pytest.approx(3.99073 * 2 * 10)
    # exactly on a table row picks the NEXT row
    assert AO.get_tolerance(torch.float32, 4)[0] == pytest.approx(2.23704e-06)
    assert AO.get_tolerance(torch.float32, 5)[0] == pytest.approx(2.95788e-06)
    # no reduction: (row0, row1), no bf16 factor with default base tolerances
    assert AO.get_tolerance(torch.bfloat16, 1) == (0.00390625, 0.0078125)
    assert AO.get_tolerance(torch.int64, 100) == (0.0, 0.0)
    # second-to-last row still has a "next" row; the last one does not
    assert AO.get_tolerance(torch.float32, 1048576)[0] == pytest.approx(0.00343043)
    assert AO.get_tolerance(torch.float32, 2097152)[0] == pytest.approx(0.00343043 * 2)


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: p.split("/")[-1][:-4])
def test_oracle_matches_golden(path):
    dt, m, n, g = load_golden(path)
    eps = g["eps"]
    x, dy, w, b = g["x"], g["dy"], g["w"], g["b"]
    rs = AO.reduction_sizes
    # LayerNorm fwd
    y, mean, invstd = AO.layer_norm_fwd(x, w, b, eps)
    AO.assert_close(y, g["ln_y"], n, "ln y")
    AO.assert_close(mean, g["ln_mean"], n, "ln mean")
    AO.assert_close(invstd, g["ln_invstd"], n, "ln invstd")
    y2, mean2, invstd2 = AO.layer_norm_fwd_python(x, w, b, eps)
    AO.assert_close(y2, g["ln_y"], n, "ln(py) y")
    AO.assert_close(mean2.unsqueeze(1), g["ln_mean"], n, "ln(py) mean")
    # LayerNorm bwd
    dx, dw, db = AO.layer_norm_bwd(dy, x, g["ln_mean"], g["ln_invstd"], w)
    AO.assert_close(dx, g["ln_dx"], n, "ln dx")
    AO.assert_close(dw, g["ln_dw"], m, "ln dw")
    AO.assert_close(db, g["ln_db"], m, "ln db")
    dx2, dw2, db2 = AO.layer_norm_bwd_python(x, dy, g["ln_mean"].view(-1), g["ln_invstd"], w)
    AO.assert_close(dx2, g["ln_dx"], n, "ln(py) dx")
    AO.assert_close(dw2, g["ln_dw"], m, "ln(py) dw")
    AO.assert_close(db2, g["ln_db"], m, "ln(py) db")
    # RMSNorm
    y, invstd = AO.rms_norm_fwd(x, w, eps)
    AO.assert_close(y, g["rms_y"], n, "rms y")
    AO.assert_close(invstd, g["rms_invstd"], n, "rms invstd")
    y2, rms = AO.rms_norm_fwd_python(x, w, eps)
    AO.assert_close(y2, g["rms_y"], n, "rms(py) y")
    AO.assert_close(1.0 / rms, g["rms_invstd"], n, "rms(py) stat")
    dx, dw = AO.rms_norm_bwd_python(x, rms, dy, w)
    AO.assert_close(dx, g["rms_dx"], n, "rms(py) dx")
    AO.assert_close(dw, g["rms_dw"], m, "rms(py) dw")
    # softmax / log-softmax
    AO.assert_close(AO.softmax_fwd(x), g["sm_y"], n * n, "softmax y")
    AO.assert_close(AO.softmax_bwd(dy, g["sm_y"]), g["sm_dx"], n, "softmax dx")
    AO.assert_close(AO.softmax_bwd_python(g["sm_y"], dy), g["sm_dx"], n, "softmax(py) dx")
    AO.assert_close(AO.log_softmax_fwd(x), g["lsm_y"], n * n, "log_softmax y")
    AO.assert_close(AO.log_softmax_bwd(dy, g["lsm_y"]), g["lsm_dx"], n, "log_softmax dx")


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16, torch.float16], ids=str)
@pytest.mark.parametrize("shape", [(7, 33), (4, 512), (2, 4096)])
def test_aten_oracle_vs_fp64_c_oracle(dt, shape):
    m, n = shape
    x, dy, w, b = make_inputs(m, n, dt, seed=1)
    eps = 1e-5
    # forward: outputs within 1 ulp of the io dtype of exact arithmetic (fp32 stats within 1e-6 rel)
    y, mean, invstd = AO.layer_norm_fwd(x, w, b, eps)
    yc, mc, ic = CO.layer_norm_fwd(x, w, b, eps)
    assert ulp_err(y, yc, floor=1.0 if dt == torch.float32 else 1e-2) <= (16 if dt == torch.float32 else 1.01)
    assert torch.allclose(mean.double(), mc, rtol=1e-5, atol=1e-6)
    assert torch.allclose(invstd.double(), ic, rtol=1e-5)
    y, invstd = AO.rms_norm_fwd(x, w, eps)
    yc, ic = CO.rms_norm_fwd(x, w, eps)
    assert ulp_err(y, yc, floor=1.0 if dt == torch.float32 else 1e-3) <= (16 if dt == torch.float32 else 1.01)
    assert torch.allclose(invstd.double(), ic, rtol=1e-5)
    y = AO.softmax_fwd(x)
    assert ulp_err(y, CO.softmax_fwd(x), floor=1e-6) <= (64 if dt == torch.float32 else 1.01)
    yl = AO.log_softmax_fwd(x)
    assert ulp_err(yl, CO.softmax_fwd(x, log=True), floor=1.0) <= (64 if dt == torch.float32 else 1.01)
    # backward
    _, mean, invstd = AO.layer_norm_fwd(x, w, b, eps)
    dx, dw, db = AO.layer_norm_bwd(dy, x, mean, invstd, w)
    dxc, dwc, dbc = CO.norm_bwd(dy, x, mean.view(-1), invstd.view(-1), w, centered=True)
    big = 256 if dt == torch.float32 else 1.01
    assert ulp_err(dx, dxc, floor=1.0 if dt == torch.float32 else 1e-1) <= big
    assert ulp_err(dw, dwc, floor=1.0) <= big
    assert ulp_err(db, dbc, floor=1.0) <= big
    _, rstd = AO.rms_norm_fwd(x, w, eps)
    dx, dw = AO.rms_norm_bwd(dy, x, rstd, w)
    dxc, dwc, _ = CO.norm_bwd(dy, x, None, rstd.view(-1), w, centered=True, want_db=False)
    assert ulp_err(dx, dxc, floor=1.0 if dt == torch.float32 else 1e-1) <= big
    assert ulp_err(dw, dwc, floor=1.0) <= big
    sm = AO.softmax_fwd(x)
    assert ulp_err(AO.softmax_bwd(dy, sm), CO.softmax_bwd(dy, sm), floor=1e-4) <= big


def test_python_and_cpp_spellings_agree():
    m, n = 6, 256
    x, dy, w, b = make_inputs(m, n, torch.float32, seed=2)
    eps = 1e-5
    _, invstd = AO.rms_norm_fwd(x, w, eps)
    _, rms = AO.rms_norm_fwd_python(x, w, eps)
    assert torch.allclose(1.0 / rms, invstd, rtol=1e-6)
    # textbook (python trace) RMSNorm bwd == fp64 oracle without the centring term
    dx_py, dw_py = AO.rms_norm_bwd_python(x, rms, dy, w)
    dxc, dwc, _ = CO.norm_bwd(dy, x, None, rms.view(-1), w, stat_is_rms=True, centered=False, want_db=False)
    assert torch.allclose(dx_py.double(), dxc, rtol=1e-4, atol=1e-5)
    assert torch.allclose(dw_py.double(), dwc, rtol=1e-4, atol=1e-4)
    # the C++ rms_norm_backward builder keeps -sum(g): it differs from the textbook form
    dx_cpp, _ = AO.rms_norm_bwd(dy, x, invstd, w)
    assert not torch.allclose(dx_cpp, dx_py, rtol=1e-3, atol=1e-3)
    # the C++ thunder builder, replayed literally, halves the projection term
    dx_th, dw_th = AO.thunder_rms_norm_bwd(dy, x, rms, w)
    g = dy * w
    proj = x * (g * x).sum(-1, keepdim=True) / (n * rms ** 3)
    assert torch.allclose(dx_th, g / rms - 0.5 * proj, rtol=1e-4, atol=1e-5)
    assert torch.allclose(dw_th, dw_py, rtol=1e-5, atol=1e-5)


def test_nan_and_inf_semantics():
    x = torch.tensor([[1.0, float("nan"), 3.0], [float("-inf")] * 3, [0.0, float("inf"), 1.0], [1.0, 2.0, 3.0]])
    y = AO.softmax_fwd(x)
    yc = CO.softmax_fwd(x)
    assert torch.isnan(y[:3]).all() and torch.isnan(yc[:3]).all()
    assert torch.allclose(y[3].double(), yc[3])
    assert math.isnan(CO.reduce_inner(x, "max")[0].item()) and CO.reduce_inner(x, "fmax")[0].item() == 3.0


@pytest.mark.parametrize("path", outer_golden_files(), ids=lambda p: p.split("/")[-1][:-4])
def test_outer_oracle_matches_golden(path):
    """The oracle's outer (dim 0) ops against fixtures made by ATen's native kernels: _softmax / _log_softmax
    (+ backward data) over dim 0, native_batch_norm(_backward) in training mode on [M, C]
    (tests/golden/make_golden_outer.py)."""
    dt, m, n, g = load_golden(path)
    eps = g["eps"]
    x, dy, w, b = g["x"], g["dy"], g["w"], g["b"]
    AO.assert_close(AO.softmax_fwd(x, dim=0), g["osm_y"], m * m, "outer softmax y")
    AO.assert_close(AO.log_softmax_fwd(x, dim=0), g["olsm_y"], m * m, "outer log_softmax y")
    AO.assert_close(AO.softmax_bwd(dy, g["osm_y"], dim=0), g["osm_dx"], m, "outer softmax dx")
    AO.assert_close(AO.log_softmax_bwd(dy, g["olsm_y"], dim=0), g["olsm_dx"], m, "outer log_softmax dx")
    y, mean, invstd = AO.outer_norm_fwd(x, w, b, eps)
    AO.assert_close(y, g["bn_y"], m, "batch norm y")
    AO.assert_close(mean, g["bn_mean"], m, "batch norm mean")
    AO.assert_close(invstd, g["bn_invstd"], m, "batch norm invstd")
    dx, dw, db = AO.outer_norm_bwd(dy, x, g["bn_mean"], g["bn_invstd"], w)
    AO.assert_close(dx, g["bn_dx"], m, "batch norm dx")
    AO.assert_close(dw, g["bn_dw"], m, "batch norm dweight")
    AO.assert_close(db, g["bn_db"], m, "batch norm dbias")

"""ATen oracle — TEST INFRASTRUCTURE ONLY. Never imported by ``fuser_b200``.

A CPU restatement of how the reference validates the hot path: nvFuser's
``testValidate`` evaluates the *unscheduled* fusion node by node with ATen
(reference: csrc/validator_utils.cpp:334-521 → ``Expr::evaluate`` in
csrc/ir/internal_nodes.cpp: UnaryOp :571-700, BinaryOp :771-868, BroadcastOp
:1494-1512, ReductionOp :1726-1778, WelfordOp :2058-2083) and compares with
tolerances looked up by (dtype, accumulated reduction size)
(csrc/validator_utils.cpp:25-163, tables csrc/validator_utils.h:28-46).

The arithmetic itself lives in a third-party dependency of the reference that is
not vendored under /root/reference: **PyTorch/ATen** (required "Torch >= 2.9",
cmake/DependencyRequirements.cmake:29; the version pinned here is the one in
this image, torch 2.11.0). Each function below replays one of the reference's
op builders (csrc/ops/normalization.cpp) or Python benchmark definitions
(benchmarks/python/test_*_{fwd,bwd}.py) with exactly the ATen calls the
evaluator would issue: fp32 intermediates, python-double scalars, casts at the
same places, ``var_mean(unbiased=False)`` for Welford.

Pinning: the reference stores no golden vectors for this path (every test
recomputes its expectation with ATen at run time), and nvFuser itself cannot be
built in this environment (LLVM >= 18.1 required). The oracle is therefore
pinned against the explicit ATen references the reference's own tests call —
``at::native_layer_norm`` / ``native_layer_norm_backward``
(tests/cpp/test_gpu1.cpp:6751-6760, test_combined_inner_outer_reduction.cpp:113-140),
``at::_softmax`` / ``_softmax_backward_data``, the hand-written RMSNorm ATen of
tests/cpp/test_gpu1.cpp:6806-6830,6897-6915 — and against an independent fp64 C
restatement (oracle/c/norm_oracle.c); see tests/test_oracle.py and
tests/golden/.
"""
from __future__ import annotations

import math

import torch

# ---------------------------------------------------------------------------
# tolerances — csrc/validator_utils.h:28-46 (data) + validator_utils.cpp:25-163 (rule)
# ---------------------------------------------------------------------------
SUM_TOLERANCES_FLOAT = [
    (4, 1.68222e-06), (8, 2.23704e-06), (16, 2.95788e-06), (32, 4.4778e-06), (64, 6.75395e-06),
    (128, 8.57934e-06), (256, 1.30594e-05), (512, 2.19122e-05), (1024, 3.3451e-05), (2048, 5.78476e-05),
    (4096, 0.000108292), (8192, 0.00012207), (16384, 0.000136882), (32768, 0.000248561), (65536, 0.000407594),
    (131072, 0.000500901), (262144, 0.000923019), (524288, 0.00156909), (1048576, 0.00223107),
    (2097152, 0.00343043),
]
SUM_TOLERANCES_HALF = [
    (4, 0.00390625), (8, 0.0078125), (16, 0.0078125), (32, 0.0155334), (64, 0.0156269), (128, 0.0312042),
    (256, 0.0312548), (512, 0.0619979), (1024, 0.0625103), (2048, 0.124686), (4096, 0.12501), (8192, 0.24945),
    (16384, 0.250049), (32768, 0.498946), (65536, 0.500071), (131072, 0.985087), (262144, 1.00006),
    (524288, 1.99234), (1048576, 2.00032), (2097152, 3.99073),
]


def get_tolerance(dtype: torch.dtype, reduction_size: int) -> tuple[float, float]:
    """(atol, rtol) exactly as getTolerance() picks them (validator_utils.cpp:25-163)."""
    if dtype in (torch.float32, torch.float64):
        table, scale = SUM_TOLERANCES_FLOAT, 1.0
    elif dtype == torch.float16:
        table, scale = SUM_TOLERANCES_HALF, 1.0
    elif dtype == torch.bfloat16:
        table, scale = SUM_TOLERANCES_HALF, 10.0
    elif not dtype.is_floating_point:
        return 0.0, 0.0
    else:
        raise ValueError(f"no tolerance rule for {dtype}")
    if reduction_size <= 1:
        # no-reduction case: (entry[0], entry[1]) and, with the default base tolerances
        # (-1), *without* the bf16 x10 factor (validator_utils.cpp:121-126)
        return table[0][1], table[1][1]
    entry = 0
    while entry < len(table) and table[entry][0] < reduction_size:
        entry += 1
    if entry + 1 < len(table):
        abs_tol = table[entry + 1][1]
    else:
        abs_tol = table[-1][1] * 2.0
    return abs_tol * scale, abs_tol * 0.01 * scale


def assert_close(actual: torch.Tensor, expected: torch.Tensor, reduction_size: int, what: str = ""):
    """The comparison of testValidate (validator_utils.cpp:481-518): allclose in the
    ATen output's dtype with equal_nan=True; exact equality for non-float outputs."""
    assert actual.shape == expected.shape, f"{what}: shape {tuple(actual.shape)} vs {tuple(expected.shape)}"
    a = actual.detach().cpu()
    e = expected.detach().cpu()
    if e.is_floating_point():
        atol, rtol = get_tolerance(e.dtype, reduction_size)
        a = a.to(e.dtype)
        ok = torch.allclose(e, a, rtol=rtol, atol=atol, equal_nan=True)
        if not ok:
            err = (e.double() - a.double()).abs().max().item()
            raise AssertionError(f"{what}: max abs err {err:.6g} > atol {atol:.6g} (rtol {rtol:.6g})")
    else:
        assert torch.equal(e, a.to(e.dtype)), f"{what}: integer outputs differ"


def _up(t: torch.Tensor) -> torch.Tensor:
    """bf16/fp16 inputs are cast to Float first in every benchmark definition
    (``if dtype in PROMOTE_DTYPES: T = fd.ops.cast(T, DataType.Float)``)."""
    return t.to(torch.float32) if t.dtype in (torch.bfloat16, torch.float16) else t


# ---------------------------------------------------------------------------
# forward builders
# ---------------------------------------------------------------------------
def layer_norm_fwd(x, weight, bias, eps: float):
    """csrc/ops/normalization.cpp:304-342 evaluated per Expr (WelfordOp → var_mean)."""
    io = x.dtype
    xf = _up(x)
    n = xf.shape[-1]
    var, avg = torch.var_mean(xf, dim=[-1], unbiased=False, keepdim=False)  # WelfordOp::evaluate
    var_sum = var * n
    mean_bcast = avg.unsqueeze(-1)
    x_sub_mean = xf - mean_bcast
    var_ = var_sum.unsqueeze(-1) * (1.0 / float(n))
    invstd = torch.rsqrt(var_ + eps)
    y = x_sub_mean * invstd
    if weight is not None:
        y = y * _up(weight)
    if bias is not None:
        y = y + _up(bias)
    return y.to(io), mean_bcast, invstd


def outer_norm_fwd(x, weight, bias, eps: float):
    """Batch-norm-shaped forward, training statistics, channels last
    (benchmarks/python/normalization.py:13-95 with channels_last=True, evaluated per Expr): var_mean over
    every axis but the last (correction 0), invstd = rsqrt(var + eps), y = (x - mean) * invstd * w + b.
    Returns y (io dtype), mean [C], invstd [C] (fp32)."""
    io = x.dtype
    xf = _up(x)
    axes = list(range(xf.dim() - 1))
    var, mean = torch.var_mean(xf, dim=axes, correction=0, keepdim=False)
    invstd = torch.rsqrt(var + eps)
    y = (xf - mean) * invstd
    if weight is not None:
        y = y * _up(weight)
    if bias is not None:
        y = y + _up(bias)
    return y.to(io), mean, invstd


def outer_norm_bwd(dy, x, mean, invstd, weight):
    """benchmarks/python/normalization.py:98-190 (batch norm backward, training, channels last), per Expr.
    Returns dx, dweight, dbias in the io dtype."""
    io = x.dtype
    xf, gf = _up(x), _up(dy)
    axes = list(range(xf.dim() - 1))
    num = 1
    for ax in axes:
        num *= xf.shape[ax]
    norm = 1.0 / float(num)
    grad_sum = torch.sum(gf, dim=axes)
    x_sub_mean = xf - mean
    dot_p = torch.sum(gf * x_sub_mean, dim=axes)
    grad_mean = grad_sum * norm
    proj_scale = (dot_p * norm) * (invstd * invstd)
    grad_scale = (_up(weight) if weight is not None else 1.0) * invstd
    dx = ((gf - proj_scale * x_sub_mean) - grad_mean) * grad_scale
    dw = dot_p * invstd
    return dx.to(io), dw.to(io), grad_sum.to(io)


def layer_norm_fwd_python(x, weight, bias, eps: float):
    """benchmarks/python/test_layernorm_fwd.py:14-54 (var_mean, correction=0)."""
    io = x.dtype
    xf, w, b = _up(x), _up(weight), _up(bias)
    var, mean = torch.var_mean(xf, dim=[1], correction=0, keepdim=False)
    t7 = var.unsqueeze(1)
    t11 = mean.unsqueeze(1)
    t14 = torch.rsqrt(t7 + eps)
    t24 = (xf - t11) * t14
    t28 = t24 * w + b
    return t28.to(io), mean, t14


def rms_norm_fwd(x, weight, eps: float):
    """csrc/ops/normalization.cpp:352-380: returns (y, invstd[...,1])."""
    io = x.dtype
    xf = _up(x)
    n = xf.shape[-1]
    var_sum = torch.sum(xf * xf, dim=[-1])
    var = var_sum.unsqueeze(-1) * (1.0 / float(n))
    invstd = torch.rsqrt(var + eps)
    y = xf * invstd
    if weight is not None:
        y = y * _up(weight)
    return y.to(io), invstd


def rms_norm_fwd_python(x, weight, eps: float):
    """benchmarks/python/test_rmsnorm_fwd.py:14-45: returns (y, rms[M,1])."""
    io = x.dtype
    xf, w = _up(x), _up(weight)
    t4 = torch.pow(xf, 2.0)
    t5 = torch.sum(t4, dim=[1])
    t9 = t5.unsqueeze(1)
    t12 = t9 * (1.0 / float(xf.shape[1]))
    t15 = torch.sqrt(t12 + eps)
    t22 = 1.0 / t15
    t23 = xf * t22
    t29 = w * t23
    return t29.to(io), t15


# ---------------------------------------------------------------------------
# forward norms with a fused residual add
# ---------------------------------------------------------------------------
def residual_sum(x, residual, round_sum: bool):
    """``x + residual`` the way the fused-add fusions evaluate it: both operands cast to Float, one
    fp32 add (benchmarks/python/test_dropout_layernorm_fwd.py:47-56, T16 = add(T3, T15) with the
    dropout removed; test_dropout_rmsnorm_fwd.py T15). Returns (value the norm sees, sum in the io
    dtype). ``round_sum``: the norm sees the sum *after* its cast to the io dtype — what a
    reduced-precision ``hidden_states + residual`` feeding the norm gives
    (benchmarks/python/torch_ops.py:76-83 traced through Thunder keeps that cast pair)."""
    s = _up(x) + _up(residual)
    s_io = s.to(x.dtype)
    return (_up(s_io) if round_sum else s), s_io


def add_rms_norm_fwd(x, residual, weight, eps: float, round_sum: bool = False, python: bool = False):
    """rms_norm(x + residual): (y, stat, sum). ``python`` selects the Thunder spelling (rms statistic)."""
    s, s_io = residual_sum(x, residual, round_sum)
    if python:
        w = weight if weight is not None else torch.ones(x.shape[-1], dtype=x.dtype)
        y, stat = rms_norm_fwd_python(s, w, eps)
    else:
        y, stat = rms_norm_fwd(s, weight, eps)
    return y.to(x.dtype), stat, s_io


def add_layer_norm_fwd(x, residual, weight, bias, eps: float, round_sum: bool = False):
    """layer_norm(x + residual): (y, mean, invstd, sum)."""
    s, s_io = residual_sum(x, residual, round_sum)
    y, mean, invstd = layer_norm_fwd(s, weight, bias, eps)
    return y.to(x.dtype), mean, invstd, s_io


def rmsnorm_add_eager(x, weight, residual, eps: float = 1e-5):
    """benchmarks/python/torch_ops.py:76-83 verbatim (every op in the io dtype, as eager runs it)."""
    hidden_states = x + residual
    residual = hidden_states
    variance = hidden_states.pow(2).mean(-1, keepdim=True)
    hidden_states = hidden_states * torch.rsqrt(variance + eps)
    hidden_states = weight * hidden_states
    return hidden_states, residual


def softmax_fwd(x, dim: int = -1):
    """csrc/ops/normalization.cpp:169-189 / benchmarks/python/test_softmax_fwd.py:14-42."""
    io = x.dtype
    xf = _up(x)
    mx = torch.amax(xf, dim=[dim], keepdim=True)
    e = torch.exp(xf - mx)
    s = torch.sum(e, dim=[dim], keepdim=True)
    return (e * (1.0 / s)).to(io)


def log_softmax_fwd(x, dim: int = -1):
    """csrc/ops/normalization.cpp:212-231."""
    io = x.dtype
    xf = _up(x)
    mx = torch.amax(xf, dim=[dim], keepdim=True)
    d = xf - mx
    s = torch.sum(torch.exp(d), dim=[dim], keepdim=True)
    return (d - torch.log(s)).to(io)


# ---------------------------------------------------------------------------
# backward builders
# ---------------------------------------------------------------------------
def softmax_bwd(dy, y, dim: int = -1):
    """csrc/ops/normalization.cpp:191-210."""
    io = y.dtype
    dyf, yf = _up(dy), _up(y)
    new_grad = dyf * yf
    s = torch.sum(new_grad, dim=[dim], keepdim=True)
    return (new_grad - yf * s).to(io)


def softmax_bwd_python(y, dy, dim: int = 1):
    """benchmarks/python/test_softmax_bwd.py:20-66 (T0 = y, T1 = dy; the row sum
    makes a round trip through the io dtype)."""
    io = y.dtype
    t0, t1 = _up(y), _up(dy)
    t4 = t0 * t1
    t5 = torch.sum(t4, dim=[dim])
    t10 = t5.unsqueeze(dim)
    if io != torch.float32:
        t10 = t10.to(io).to(torch.float32)
    t19 = t0 * (t1 - t10)
    return t19.to(io)


def log_softmax_bwd(dy, y, dim: int = -1):
    """csrc/ops/normalization.cpp:233-249."""
    io = y.dtype
    dyf, yf = _up(dy), _up(y)
    s = torch.sum(dyf, dim=[dim], keepdim=True)
    return (dyf - torch.exp(yf) * s).to(io)


def _norm_bwd_common(dy, x, mean, invstd, weight, want_db):
    io = x.dtype
    dyf, xf = _up(dy), _up(x)
    n = float(xf.shape[-1])
    outer_axes = list(range(xf.dim() - 1))
    x_hat = (xf - mean) * invstd if mean is not None else xf * invstd
    g = dyf * _up(weight) if weight is not None else dyf
    a = n * g
    b = torch.sum(g, dim=[-1], keepdim=True)
    c2 = torch.sum(g * x_hat, dim=[-1], keepdim=True)
    c3 = x_hat * c2
    inner = (a - b) - c3
    dx = ((1.0 / n) * invstd) * inner
    dw = torch.sum(dyf * x_hat, dim=outer_axes) if weight is not None else None
    db = torch.sum(dyf, dim=outer_axes) if want_db else None
    return dx.to(io), (dw.to(io) if dw is not None else None), (db.to(io) if db is not None else None)


def layer_norm_bwd(dy, x, mean, invstd, weight, has_bias=True):
    """csrc/ops/normalization.cpp:382-436 → (dx, dw, db)."""
    return _norm_bwd_common(dy, x, mean, invstd, weight, has_bias)


def rms_norm_bwd(dy, x, invstd, weight):
    """csrc/ops/normalization.cpp:438-485 (keeps the -sum(g) term) → (dx, dw)."""
    dx, dw, _ = _norm_bwd_common(dy, x, None, invstd, weight, False)
    return dx, dw


def thunder_rms_norm_bwd(dy, x, rms, weight):
    """csrc/ops/normalization.cpp:487-535 as written → (dx, dw). NOTE: as written
    the C++ builder's dx carries a factor 1/2 on the projection term that the Python
    trace (below) compensates with an extra ``* 2``; this function replays the C++
    text literally."""
    io = x.dtype
    dyf, xf = _up(dy), _up(x)
    n = float(xf.shape[-1])
    outer_axes = list(range(xf.dim() - 1))
    gw = dyf * _up(weight) if weight is not None else dyf
    neg = -gw
    t = (neg * xf) * (1.0 / (rms * rms))
    inner = torch.sum(t, dim=[-1], keepdim=True)
    inv_rms_mul2 = 1.0 / (rms * 2.0)
    xs = xf * (inv_rms_mul2 * (1.0 / n))
    inner_mul = inner * xs
    inv_rms = 1.0 / rms
    dx = inner_mul + gw * inv_rms
    dw = torch.sum(dyf * (xf * inv_rms), dim=outer_axes) if weight is not None else None
    return dx.to(io), (dw.to(io) if dw is not None else None)


def rms_norm_bwd_python(x, rms, dy, weight):
    """benchmarks/python/test_rmsnorm_bwd.py:20-76 (inputs: x, rms[M,1], dy, w) → (dx, dw)."""
    io = x.dtype
    t4, t6, t7 = _up(x), _up(dy), _up(weight)
    t5 = rms
    s0 = 2.0
    t14 = t5.expand_as(t4)
    t15 = 1.0 / t14
    t16 = t4 * t15
    t23 = t6 * t16
    t24 = t6 * t7
    t25 = torch.sum(t23, dim=[0])
    t28 = t24 * t15
    t30 = (-t24) * t4
    t33 = 1.0 / torch.pow(t14, s0)
    t34 = t30 * t33
    t35 = torch.sum(t34, dim=[1])
    t41 = t35.unsqueeze(1)
    t44 = 1.0 / (s0 * t5)
    t45 = t41 * t44
    t49 = t45 * (1.0 / float(t4.shape[1]))
    t50 = torch.sum(t49, dim=[1])
    t58 = t50.unsqueeze(1).expand_as(t4)
    t62 = (t58 * s0) * t4
    t63 = t28 + t62
    return t63.to(io), t25.to(io)


def layer_norm_bwd_python(x, dy, mean, invstd, weight):
    """benchmarks/python/test_layernorm_bwd.py:20-101 (inputs: x, dy, mean[M],
    invstd[M,1], w) → (dx, dw, db)."""
    io = x.dtype
    t0, t1, t4 = _up(x), _up(dy), _up(weight)
    t2, t3 = mean, invstd
    n = float(t0.shape[1])
    t13 = t2.unsqueeze(1).expand_as(t0)
    t14 = t0 - t13
    t18 = t3.expand_as(t0)
    t19 = t14 * t18
    t28 = torch.sum(t1, dim=[0])
    t30 = t1 * t4
    t31 = t1 * t19
    t32 = torch.sum(t31, dim=[0])
    t34 = t30 * t18
    t35 = t30 * t14
    t36 = torch.sum(t35, dim=[1])
    t40 = t36.unsqueeze(1)
    t42 = torch.sum(-t34, dim=[1])
    t46 = t42.unsqueeze(1)
    t48 = -0.5 * t40
    t50 = torch.pow(t3, 3.0)
    t51 = t48 * t50
    t54 = torch.sum(t46, dim=[1])
    t55 = torch.sum(t51, dim=[1])
    t63 = t55.unsqueeze(1).expand_as(t0)
    t71 = t2.unsqueeze(1).expand_as(t0)
    t73 = 2.0 * t63
    t74 = t0 - t71
    t75 = t73 * t74
    s77 = 1.0 / n
    t78 = t75 * s77
    t86 = t54.unsqueeze(1).expand_as(t0)
    t88 = s77 * t86
    t89 = t78 + t88
    t90 = t34 + t89
    return t90.to(io), t32.to(io), t28.to(io)


# ---------------------------------------------------------------------------
# reduction sizes as ReductionSizeMapper accumulates them along the producer
# chain (validator_utils.cpp:165-254): own reduction extent x max over inputs.
# ---------------------------------------------------------------------------
def reduction_sizes(op: str, m: int, n: int) -> dict:
    if op in ("layer_norm_fwd", "rms_norm_fwd"):
        return {"y": n, "mean": n, "invstd": n, "stat": n}
    if op in ("softmax_fwd", "log_softmax_fwd"):
        return {"y": n * n}  # max-reduce feeds the sum-reduce
    if op in ("softmax_bwd", "log_softmax_bwd"):
        return {"dx": n}
    if op in ("layer_norm_bwd", "rms_norm_bwd"):
        return {"dx": n, "dw": m, "db": m}
    raise KeyError(op)


def ulp_bound(dtype: torch.dtype) -> float:
    """Relative size of one unit in the last place of the io dtype (for the
    repository's own tighter check against an fp64 evaluation)."""
    return {torch.float32: 2.0 ** -23, torch.bfloat16: 2.0 ** -7, torch.float16: 2.0 ** -10}[dtype]


__all__ = [n for n in dir() if not n.startswith("_")] + ["math"]


# ---------------------------------------------------------------------------
# plain reductions and stand-alone statistics (ReductionOp / WelfordOp::evaluate,
# csrc/ir/internal_nodes.cpp:1726-1778, :2058-2083)
# ---------------------------------------------------------------------------
def reduce_inner(x, kind: str = "sum"):
    """sum / max over the last axis as the evaluator computes them: at::sum in fp32 (the fusions cast reduced
    precision up first) or at::amax (NaN-propagating), cast back to the io dtype."""
    xf = _up(x)
    r = torch.sum(xf, dim=[-1]) if kind == "sum" else torch.amax(xf, dim=[-1])
    return r.to(x.dtype)


def reduce_outer(x):
    """sum over every axis but the last."""
    xf = _up(x)
    return torch.sum(xf, dim=list(range(xf.dim() - 1))).to(x.dtype)


def var_mean(x, correction: int = 1):
    """ops.var_mean over the last axis -> (var, mean), fp32 (variance_mean, csrc/ops/normalization.cpp:93-158:
    Welford, var = M2 / max(N - correction, 0))."""
    xf = _up(x)
    var, mean = torch.var_mean(xf, dim=[-1], correction=correction, keepdim=False)
    return var, mean


def welford(x):
    """ops.welford over the last axis -> (avg, var_sum): WelfordOp::evaluate is at::var_mean(unbiased=False),
    var_sum = var * N (internal_nodes.cpp:2058-2083)."""
    xf = _up(x)
    var, mean = torch.var_mean(xf, dim=[-1], unbiased=False, keepdim=False)
    return mean, var * xf.shape[-1]


# ---------------------------------------------------------------------------
# fused dropout (benchmarks/python/test_dropout_{layernorm,rmsnorm}_{fwd,bwd}.py)
# ---------------------------------------------------------------------------
def philox4x32_10(seed: int, subsequence, offset: int):
    """Philox4x32-10 as the reference's runtime evaluates it (runtime/random_numbers.cu:8-58): key = seed,
    counter = (offset lo, offset hi, subsequence lo, subsequence hi). numpy, vectorised over ``subsequence``;
    returns the four output words. Pinned to the Random123 known-answer vectors in tests/test_dropout.py."""
    import numpy as np

    a, b, m32 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    sub = np.asarray(subsequence, dtype=np.uint64)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    c0 = np.full(sub.shape, offset & 0xFFFFFFFF, dtype=np.uint64)
    c1 = np.full(sub.shape, (offset >> 32) & 0xFFFFFFFF, dtype=np.uint64)
    c2, c3 = sub & m32, sub >> np.uint64(32)
    for _ in range(10):
        p0, p1 = a * c0, b * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & m32, p1 >> np.uint64(32), p1 & m32
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + 0x9E3779B9) & 0xFFFFFFFF, (k1 + 0xBB67AE85) & 0xFFFFFFFF
    return c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32)


def uniformf(words):
    """runtime/random_numbers.cu:66-88: fl(fl(w) * 2^-32 + 2^-33) in float32, with 1.0 mapped to 0."""
    import numpy as np

    scale = np.float32(2.0 ** -32)
    u = words.astype(np.float32) * scale + np.float32(0.5) * scale
    return np.where(u == np.float32(1.0), np.float32(0.0), u).astype(np.float32)


def dropout_mask(seed: int, offset: int, shape, keep: float) -> torch.Tensor:
    """mask = uniform(0, 1) < keep for a Float uniform() of ``shape``: element i takes word i mod 4 of
    philox(seed, i / 4, offset) (device_lower/pass/rng.cpp:95-127); the comparison promotes the float to the
    fusion's double scalar."""
    import numpy as np

    n = 1
    for s in shape:
        n *= int(s)
    nsub = (n + 3) // 4
    w = np.stack(philox4x32_10(seed, np.arange(nsub, dtype=np.uint64), offset), axis=1).reshape(-1)[:n]
    return torch.from_numpy(uniformf(w).astype(np.float64) < float(keep)).view(*shape)


def dropout_sum(x, residual, mask, keep: float):
    """residual + dropout(x): T13 = x * mask, T15 = T13 * (1 / keep), T16 = residual + T15, all in Float
    (test_dropout_layernorm_fwd.py:47-56; the double scalar is applied in the tensor's float arithmetic)."""
    scale = torch.tensor(1.0 / keep, dtype=torch.float32)
    return _up(residual) + (_up(x) * mask.to(torch.float32)) * scale


def dropout_layer_norm_fwd(x, residual, weight, bias, eps: float, mask, keep: float):
    """(y, mean, invstd) of layer_norm(residual + dropout(x)) for a given mask."""
    y, mean, invstd = layer_norm_fwd(dropout_sum(x, residual, mask, keep), weight, bias, eps)
    return y.to(x.dtype), mean, invstd


def dropout_rms_norm_fwd(x, residual, weight, eps: float, mask, keep: float, python: bool = False):
    """(y, stat) of rms_norm(residual + dropout(x)); ``python``: the Thunder spelling (rms statistic)."""
    s = dropout_sum(x, residual, mask, keep)
    if python:
        w = weight if weight is not None else torch.ones(x.shape[-1], dtype=x.dtype)
        y, stat = rms_norm_fwd_python(s, w, eps)
    else:
        y, stat = rms_norm_fwd(s, weight, eps)
    return y.to(x.dtype), stat


def dropout_norm_bwd(kind: str, dy, x, residual, mask, mean, stat, weight, keep: float, stat_is_rms: bool = False):
    """Backward of the two: x_eff = residual + dropout(x) rebuilt from the mask, the plain norm backward on it,
    d residual = d x_eff, d x = d x_eff * (1 / keep) * mask (test_dropout_layernorm_bwd.py:57-101). Returns
    (dx, dresidual, dweight, dbias | None) in the io dtype."""
    io = dy.dtype
    s = dropout_sum(x, residual, mask, keep)
    if kind == "ln":
        dxe, dw, db = layer_norm_bwd(dy, s, mean, stat, weight)  # s is Float: nothing is rounded in between
    else:
        rms = stat if stat_is_rms else 1.0 / stat
        dxe, dw = rms_norm_bwd_python(s, rms, dy, weight)  # the textbook form of test_dropout_rmsnorm_bwd.py
        db = None
    dxe = _up(dxe)
    scale = torch.tensor(1.0 / keep, dtype=torch.float32)
    dx = (dxe * scale) * mask.to(torch.float32)
    cast = lambda t: None if t is None else t.to(io)  # noqa: E731
    return dx.to(io), dxe.to(io), cast(dw), cast(db)


# ---------------------------------------------------------------------------
# cross-entropy (benchmarks/python/cross_entropy_loss.py: torch.nn.functional.cross_entropy on the logits)
# ---------------------------------------------------------------------------
def nll_softmax_fwd(logits, target, ignore_index: int = -100):
    """(per-row loss, log-sum-exp) in fp32: log_softmax (csrc/ops/normalization.cpp:212-231 on the Float upcast)
    followed by at::nll_loss(reduction='none')."""
    xf = _up(logits)
    lse = torch.logsumexp(xf, dim=-1)
    loss = torch.nn.functional.nll_loss(torch.log_softmax(xf, dim=-1).reshape(-1, xf.shape[-1]), target.reshape(-1),
                                        ignore_index=ignore_index, reduction="none").reshape(target.shape)
    return loss, lse


def cross_entropy_bwd(logits, target, ignore_index: int = -100, reduction: str = "mean"):
    """d loss / d logits through autograd on the Float upcast, cast to the io dtype."""
    xf = _up(logits).clone().requires_grad_(True)
    loss = torch.nn.functional.cross_entropy(xf.reshape(-1, xf.shape[-1]), target.reshape(-1), ignore_index=ignore_index,
                                             reduction=reduction)
    loss.backward()
    return xf.grad.to(logits.dtype), loss.detach()


# ---------------------------------------------------------------------------
# instance norm (csrc/ops/normalization.cpp:790-950, :952-1060; training statistics)
# ---------------------------------------------------------------------------
def instance_norm_fwd(x, weight, bias, eps: float, channels_last: bool = False):
    """(y, mean[B, C], invstd[B, C]): Welford over the spatial axes of every (sample, channel) plane, var = M2 / N,
    y = (x - mean) * invstd * weight[c] + bias[c] on the Float upcast, cast back to the io dtype."""
    xf = _up(x)
    nd = xf.dim()
    c_axis = nd - 1 if channels_last else 1
    red = [a for a in range(1, nd) if a != c_axis]
    var, mean = torch.var_mean(xf, dim=red, correction=0, keepdim=True)
    invstd = torch.rsqrt(var + eps)
    shape = [1] * nd
    shape[c_axis] = xf.shape[c_axis]
    y = (xf - mean) * invstd
    if weight is not None:
        y = y * _up(weight).view(shape)
    if bias is not None:
        y = y + _up(bias).view(shape)
    b, c = xf.shape[0], xf.shape[c_axis]
    return y.to(x.dtype), mean.reshape(b, c), invstd.reshape(b, c)


def instance_norm_bwd(dy, x, mean, invstd, weight, channels_last: bool = False):
    """(dx, dweight[C], dbias[C]) by the builder's formula (:1014-1060): per plane
    dx = w * invstd / N * (N dy - sum dy - x_hat sum(dy x_hat)), dweight = sum over batch and space of dy x_hat."""
    io = x.dtype
    dyf, xf = _up(dy), _up(x)
    nd = xf.dim()
    c_axis = nd - 1 if channels_last else 1
    red = [a for a in range(1, nd) if a != c_axis]
    n = 1.0
    for a in red:
        n *= xf.shape[a]
    bshape = [xf.shape[0]] + [1] * (nd - 1)
    bshape[c_axis] = xf.shape[c_axis]
    cshape = [1] * nd
    cshape[c_axis] = xf.shape[c_axis]
    m, i = mean.view(bshape), invstd.view(bshape)
    xh = (xf - m) * i
    w = _up(weight).view(cshape) if weight is not None else torch.ones(cshape)
    s_dy = torch.sum(dyf, dim=red, keepdim=True)
    s_dyxh = torch.sum(dyf * xh, dim=red, keepdim=True)
    dx = (w * i / n) * (n * dyf - s_dy - xh * s_dyxh)
    dw = torch.sum(s_dyxh, dim=0).reshape(-1)
    db = torch.sum(s_dy, dim=0).reshape(-1)
    return dx.to(io), dw.to(io), db.to(io)

"""Shared helpers for the test-suite (golden fixtures, ulp metrics, input synthesis)."""
import glob
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DTYPES = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}


def golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "norm_*.npz")))


def outer_golden_files():
    """Fixtures of the outer (dim 0) ops: tests/golden/make_golden_outer.py."""
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "outer_*.npz")))


def load_golden(path):
    name = os.path.basename(path)[: -len(".npz")]
    _, dt_name, shape = name.split("_")
    dt = DTYPES[dt_name]
    m, n = (int(v) for v in shape.split("x"))
    raw = np.load(path)
    out = {}
    for k in raw.files:
        a = raw[k]
        if a.dtype == np.uint16:
            out[k] = torch.from_numpy(a.copy()).view(dt)
        elif a.shape == ():
            out[k] = float(a)
        else:
            out[k] = torch.from_numpy(a.copy())
    return dt, m, n, out


def ulp_err(actual: torch.Tensor, exact: torch.Tensor, floor: float = 0.0) -> float:
    """max |actual - exact| in units of the io dtype's ulp at max(|exact|, floor).

    `floor` is the magnitude below which absolute (not relative) accuracy is what an
    fp32-accumulating kernel can deliver (cancellation in sum-type outputs)."""
    dt = actual.dtype
    eps = {torch.float32: 2.0 ** -23, torch.bfloat16: 2.0 ** -7, torch.float16: 2.0 ** -10}[dt]
    a = actual.detach().cpu().double()
    e = exact.detach().cpu().double()
    both_nan = torch.isnan(a) & torch.isnan(e)
    tiny = {torch.float32: 2.0 ** -126, torch.bfloat16: 2.0 ** -126, torch.float16: 2.0 ** -14}[dt]  # subnormals
    scale = torch.clamp(e.abs(), min=max(floor, tiny))
    err = ((a - e).abs() / (scale * eps)).masked_fill(both_nan, 0.0)
    err = torch.nan_to_num(err, nan=float("inf"))
    return err.max().item() if err.numel() else 0.0


def make_inputs(m, n, dtype, seed=0, device="cpu"):
    """randn inputs as every reference test/benchmark synthesises them
    (at::randn, seed 0; benchmarks/python/test_rmsnorm_bwd.py:101-103)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(m, n, generator=g).to(dtype)
    dy = torch.randn(m, n, generator=g).to(dtype)
    w = torch.randn(n, generator=g).to(dtype)
    b = torch.randn(n, generator=g).to(dtype)
    return [t.to(device) for t in (x, dy, w, b)]

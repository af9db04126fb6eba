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


def sample_rows(m: int, block: int = 64, n_random: int = 64, seed: int = 0) -> torch.Tensor:
    """Row indices for an oracle comparison at sizes the CPU oracle cannot take whole: the first, middle and
    last `block` rows (first ring fill, steady state, ragged tail / pool tiles of the dynamic schedule) plus
    `n_random` rows spread over the matrix. Every row is independent for y / dx, so a row sample is an exact
    check of those rows."""
    if m <= 3 * block + n_random:
        return torch.arange(m)
    g = torch.Generator().manual_seed(seed)
    mid = m // 2
    idx = torch.cat([torch.arange(0, block), torch.arange(mid, mid + block), torch.arange(m - block, m),
                     torch.randint(0, m, (n_random,), generator=g)])
    return torch.unique(idx)


def col_sums_fp64(dy: torch.Tensor, x: torch.Tensor, mean, invstd, block: int = 2048):
    """fp64 column sums of the norm backward over ALL rows, in row blocks (bounded memory):
    dw = sum_rows dy * (x - mean) * invstd, db = sum_rows dy. CPU tensors; mean may be None (RMSNorm)."""
    m, n = x.shape
    dw = torch.zeros(n, dtype=torch.float64)
    db = torch.zeros(n, dtype=torch.float64)
    inv = invstd.reshape(m, 1).double()
    mu = mean.reshape(m, 1).double() if mean is not None else None
    for a in range(0, m, block):
        b = min(m, a + block)
        dyb = dy[a:b].double()
        xb = x[a:b].double()
        if mu is not None:
            xb = xb - mu[a:b]
        dw += (dyb * (xb * inv[a:b])).sum(0)
        db += dyb.sum(0)
    return dw, db

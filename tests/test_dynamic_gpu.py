"""Dynamic tile scheduling of the forward-type ring kernels: tiles beyond each group's first ring
fill are drawn from a global counter. Whatever SM processes a row, the bits must be those of the
static schedule (B200NORM_STATIC=1) — and of the oracle."""
import os

import pytest
import torch

from fuser_b200 import ops
from oracle import aten_oracle as AO
from tests.util import make_inputs

pytestmark = pytest.mark.gpu

SHAPES = [(40000, 64), (9999, 512), (7001, 1024), (3001, 4096), (16384, 4096), (1500, 12288)]


def _both(fn):
    """(default pool, everything-dynamic pool run twice, static): POOL=1 draws every tile past the
    first ring fill from the counter, LayerNorm forward included (its default is static)."""
    os.environ.pop("B200NORM_STATIC", None)
    os.environ.pop("B200NORM_POOL", None)
    dflt = fn()
    os.environ["B200NORM_POOL"] = "1"
    try:
        dyn = fn()
        dyn2 = fn()  # second launch: the scheduler words must have been returned to zero
    finally:
        os.environ.pop("B200NORM_POOL", None)
    os.environ["B200NORM_STATIC"] = "1"
    try:
        sta = fn()
    finally:
        os.environ.pop("B200NORM_STATIC", None)
    torch.cuda.synchronize()
    assert _same(dflt, sta), "default pool and static schedules differ"
    return dyn, dyn2, sta


def _same(a, b):
    a = a if isinstance(a, (tuple, list)) else (a,)
    b = b if isinstance(b, (tuple, list)) else (b,)
    return all(torch.equal(p, q) for p, q in zip(a, b))


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float32], ids=["bf16", "f32"])
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: f"{s[0]}x{s[1]}")
def test_dynamic_equals_static(shape, dt):
    m, n = shape
    x, dy, w, b = make_inputs(m, n, dt, seed=n)
    xd, dyd, wd, bd = (t.cuda() for t in (x, dy, w, b))
    cases = {
        "rms_fwd": lambda: ops.rms_norm_fwd(xd, wd, 1e-5),
        "ln_fwd": lambda: ops.layer_norm_fwd(xd, wd, bd, 1e-5),
        "add_rms_fwd": lambda: ops.rms_norm_fwd(xd, wd, 1e-5, residual=dyd, want_sum=True),
        "sm_fwd": lambda: ops.softmax_fwd(xd),
        "lsm_fwd": lambda: ops.softmax_fwd(xd, log=True),
    }
    for name, fn in cases.items():
        dyn, dyn2, sta = _both(fn)
        assert _same(dyn, sta), f"{name}: dynamic and static schedules differ"
        assert _same(dyn2, sta), f"{name}: second dynamic launch differs (scheduler words not reset?)"
    y = ops.softmax_fwd(xd)
    dyn, dyn2, sta = _both(lambda: ops.softmax_bwd(dyd, y))
    assert _same(dyn, sta) and _same(dyn2, sta), "sm_bwd: dynamic and static schedules differ"
    # and the result is the oracle's
    ye, se = AO.rms_norm_fwd(x, w, 1e-5)
    yd, sd = ops.rms_norm_fwd(xd, wd, 1e-5)
    AO.assert_close(yd, ye, n, "rms y")
    AO.assert_close(sd, se, n, "rms invstd")
    AO.assert_close(ops.softmax_fwd(xd), AO.softmax_fwd(x), n * n, "softmax y")


def test_dynamic_every_row_written_once():
    """Poisoned output: every row must be produced (no tile skipped or left to a stale counter)
    across many back-to-back launches on the same stream / workspace."""
    m, n = 30011, 256
    x = torch.randn(m, n, device="cuda", dtype=torch.bfloat16)
    ref = None
    for it in range(20):
        y, st = ops.rms_norm_fwd(x, None, 1e-5)
        if ref is None:
            os.environ["B200NORM_STATIC"] = "1"
            try:
                ref, _ = ops.rms_norm_fwd(x, None, 1e-5)
            finally:
                os.environ.pop("B200NORM_STATIC", None)
        assert torch.equal(y, ref), f"launch {it} differs"
        assert torch.isfinite(st).all()


def test_cuda_graph_replay_of_the_c2_step():
    """The forward + backward step captured in a CUDA graph and replayed: the in-kernel state
    (grid-barrier word, tile-pool tickets, done counter) must return to its rest value after every
    launch, so replays reproduce the eager results bit for bit."""
    m, n = 16384, 4096
    x, dy, w, _ = make_inputs(m, n, torch.bfloat16, seed=11)
    xd, dyd, wd = x.cuda(), dy.cuda(), w.cuda()
    big = torch.randn(40000, 512, device="cuda", dtype=torch.bfloat16)  # a pool-scheduled forward as well
    y0, r0 = ops.rms_norm_fwd(xd, wd, 1e-5, stat_is_rms=True)
    dx0, dw0 = ops.rms_norm_bwd(dyd, xd, r0, wd, stat_is_rms=True)
    s0 = ops.softmax_fwd(big)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):  # warm-up on the capture stream: workspace, kernel attributes
            y, r = ops.rms_norm_fwd(xd, wd, 1e-5, stat_is_rms=True)
            ops.rms_norm_bwd(dyd, xd, r, wd, stat_is_rms=True)
            ops.softmax_fwd(big)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        y, r = ops.rms_norm_fwd(xd, wd, 1e-5, stat_is_rms=True)
        dx, dw = ops.rms_norm_bwd(dyd, xd, r, wd, stat_is_rms=True)
        s = ops.softmax_fwd(big)
    for it in range(3):
        for t in (y, r, dx, dw, s):
            t.zero_()
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(y, y0) and torch.equal(r, r0), f"replay {it}: forward differs"
        assert torch.equal(dx, dx0) and torch.equal(dw, dw0), f"replay {it}: backward differs"
        assert torch.equal(s, s0), f"replay {it}: softmax differs"

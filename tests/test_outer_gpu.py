"""GPU parity tests of the outer-persistent kernels (softmax / log-softmax over dim 0, forward and
backward; B200NORM_F_OUTER): the segments nvFuser's OuterPersistent scheduler takes
(scheduler/normalization_outer.cpp:451-683; benchmarks/python/test_softmax_{fwd,bwd}.py with
reduction_axis = 0). Called through the C ABI (ops -> C++ NormExecutor -> b200norm_launch) and
compared with (a) the ATen oracle under the validator's tolerances and (b) the fp64 C oracle
(through a transpose: a column of the matrix is a row of its transpose) with a tight ulp bound.
Shapes cover one CTA per slab, every cluster size (rows split along M with the DSMEM exchange),
ragged row counts, slabs cut by the matrix edge, strided rows and the NaN / Inf semantics.
"""
import os

import pytest
import torch

from fuser_b200 import _lib as L
from tests import fusions as F
from fuser_b200 import ops
from fuser_b200.frontend import FusionDefinition
from fuser_b200.ir import DataType
from oracle import aten_oracle as AO
from oracle import c_oracle as CO
from tests.util import load_golden, make_inputs, outer_golden_files, ulp_err

pytestmark = pytest.mark.gpu

DTYPES = [torch.bfloat16, torch.float32, torch.float16]
TIGHT = {torch.bfloat16: 1.01, torch.float16: 1.01, torch.float32: 64.0}
# (M = reduced extent, N = iteration extent)
SHAPES = [(1, 64), (5, 8), (16, 768), (33, 1000), (216, 96), (512, 2048), (600, 136), (2048, 264), (4100, 512)]


def check(actual, aten_expected, exact, red_size, what, floor):
    torch.cuda.synchronize()
    AO.assert_close(actual, aten_expected, red_size, what)
    u = ulp_err(actual.cpu(), exact, floor=floor)
    assert u <= TIGHT[actual.dtype], f"{what}: {u:.2f} ulp vs fp64"


def c_softmax_fwd(x, log=False):
    return CO.softmax_fwd(x.t().contiguous(), log=log).t()


def c_softmax_bwd(dy, y, log=False):
    return CO.softmax_bwd(dy.t().contiguous(), y.t().contiguous(), log=log).t()


def run_all(x, dy, dt, m, what, **kw):
    xd, dyd = x.cuda(), dy.cuda()
    y = ops.softmax_fwd(xd, dim=0, **kw)
    assert y.shape == x.shape and y.dtype == dt
    check(y, AO.softmax_fwd(x, dim=0), c_softmax_fwd(x), m * m, what + " softmax y", 1e-6 if dt != torch.float32 else 1e-3)
    y = ops.softmax_fwd(xd, log=True, dim=0, **kw)
    check(y, AO.log_softmax_fwd(x, dim=0), c_softmax_fwd(x, log=True), m * m, what + " log_softmax y", 1.0)
    sm = AO.softmax_fwd(x, dim=0)
    lsm = AO.log_softmax_fwd(x, dim=0)
    dx = ops.softmax_bwd(dyd, sm.cuda(), dim=0, **kw)
    check(dx, AO.softmax_bwd(dy, sm, dim=0), c_softmax_bwd(dy, sm), m, what + " softmax dx", 1e-3 if dt != torch.float32 else 1e-1)
    dx = ops.softmax_bwd(dyd, sm.cuda(), round_rowsum=True, dim=0, **kw)
    torch.cuda.synchronize()
    AO.assert_close(dx, AO.softmax_bwd_python(sm, dy, dim=0), m, what + " softmax(py) dx")
    dx = ops.softmax_bwd(dyd, lsm.cuda(), log=True, dim=0, **kw)
    check(dx, AO.log_softmax_bwd(dy, lsm, dim=0), c_softmax_bwd(dy, lsm, log=True), m, what + " log_softmax dx", 1.0)


@pytest.mark.parametrize("dt", DTYPES, ids=["bf16", "f32", "f16"])
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: f"{s[0]}x{s[1]}")
def test_outer_softmax(shape, dt):
    m, n = shape
    x, dy, _, _ = make_inputs(m, n, dt, seed=m * 13 + n)
    run_all(x, dy, dt, m, f"[{m},{n}]")


@pytest.mark.parametrize("cluster", [1, 2, 4, 8, 16])
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float32], ids=["bf16", "f32"])
@pytest.mark.parametrize("shape", [(40, 256), (777, 384), (3000, 1024)], ids=lambda s: f"{s[0]}x{s[1]}")
def test_outer_cluster_split(shape, dt, cluster):
    """Rows of a slab split over a thread-block cluster; the per-column partials cross the cluster
    through DSMEM (reference mechanism: the grid reduction of runtime/grid_reduction.cu:905-1066)."""
    m, n = shape
    try:
        plan = ops.plan_for(L.OP_SOFTMAX_BWD, dt, m, n, flags=L.F_OUTER, cluster_size=cluster)
    except L.UnsupportedSegment:
        pytest.skip(f"M = {m} does not fit a cluster of {cluster}")
    assert plan["cluster_size"] == cluster
    x, dy, _, _ = make_inputs(m, n, dt, seed=cluster)
    run_all(x, dy, dt, m, f"[{m},{n}] C={cluster}", cluster_size=cluster)


def test_outer_large_m_two_deep_and_single_tile():
    """M = 8192 (16-CTA cluster, two tiles in flight) and M = 16384 (64-byte slabs or one tile)."""
    dt = torch.bfloat16
    for m, n in ((8192, 1024), (16384, 512)):
        x, dy, _, _ = make_inputs(m, n, dt, seed=m)
        xd = x.cuda()
        y = ops.softmax_fwd(xd, dim=0)
        torch.cuda.synchronize()
        AO.assert_close(y, AO.softmax_fwd(x, dim=0), m * m, f"softmax [{m},{n}]")
        assert ulp_err(y.cpu(), c_softmax_fwd(x), floor=1e-6) <= TIGHT[dt]
        dx = ops.softmax_bwd(dy.cuda(), y, dim=0)
        torch.cuda.synchronize()
        AO.assert_close(dx, AO.softmax_bwd(dy, y.cpu(), dim=0), m, f"softmax dx [{m},{n}]")
        assert ulp_err(dx.cpu(), c_softmax_bwd(dy, y.cpu()), floor=1e-3) <= TIGHT[dt]


def test_outer_edge_cases():
    dt = torch.bfloat16
    # empty reduction / iteration domains
    for shape in ((0, 256), (16, 0)):
        x = torch.empty(*shape, dtype=dt, device="cuda")
        assert ops.softmax_fwd(x, dim=0).shape == shape
        assert ops.softmax_bwd(x, x, dim=0).shape == shape
    # strided rows (a column block of a wider matrix) in, contiguous out
    big = torch.randn(300, 1024, dtype=torch.float32).to(dt)
    view = big.cuda()[:, 128:640]
    y = ops.softmax_fwd(view, dim=0)
    torch.cuda.synchronize()
    AO.assert_close(y, AO.softmax_fwd(big[:, 128:640], dim=0), 300 * 300, "strided rows")
    # 3-D input: [M, a, b] reduces over dim 0 with [a*b] independent columns
    x3 = torch.randn(64, 6, 40).to(dt)
    y3 = ops.softmax_fwd(x3.cuda(), dim=0)
    torch.cuda.synchronize()
    assert y3.shape == x3.shape
    AO.assert_close(y3, AO.softmax_fwd(x3, dim=0), 64 * 64, "3-D outer softmax")
    # NaN / Inf columns: the whole column turns NaN exactly as the op-by-op ATen evaluation does
    xs = torch.randn(96, 64).to(dt)
    xs[7, 1] = float("nan")
    xs[:, 2] = float("-inf")
    xs[5, 3] = float("inf")
    y = ops.softmax_fwd(xs.cuda(), dim=0)
    torch.cuda.synchronize()
    AO.assert_close(y, AO.softmax_fwd(xs, dim=0), 96 * 96, "softmax with NaN/Inf columns")
    assert torch.isnan(y[:, 1:4]).all() and not torch.isnan(y[:, 0]).any() and not torch.isnan(y[:, 4:]).any()
    # the same through a cluster split (the forward exchanges (max, sum) pairs): one CTA's rows all -inf
    # in a column that is finite elsewhere must NOT poison the column; NaN / +inf / all -inf still do
    for cluster in (2, 4, 8):
        xs = torch.randn(96, 64).to(dt)
        xs[7, 1] = float("nan")
        xs[:, 2] = float("-inf")
        xs[5, 3] = float("inf")
        xs[: 96 // cluster, 5] = float("-inf")  # exactly the rows of cluster rank 0
        xs[40:, 6] = float("-inf")
        for log in (False, True):
            y = ops.softmax_fwd(xs.cuda(), dim=0, log=log, cluster_size=cluster)
            torch.cuda.synchronize()
            ref = AO.log_softmax_fwd(xs, dim=0) if log else AO.softmax_fwd(xs, dim=0)
            AO.assert_close(y, ref, 96 * 96, f"NaN/Inf columns, cluster {cluster}, log={log}")
            assert torch.isnan(y[:, 1:4]).all() and not torch.isnan(y[:, 0]).any()
            assert not torch.isnan(y[:, 4:]).any(), f"cluster {cluster}: -inf rows poisoned a finite column"
    # shapes the persistent kernels cannot take (rows that are not a multiple of 16 bytes, a reduced extent
    # beyond a 16-CTA cluster) run on the shape-agnostic family; a forced persistent family refuses them
    xo = torch.randn(32, 100).to(dt)  # 200 bytes per row
    assert ops.plan_for(L.OP_SOFTMAX_FWD, dt, 32, 100, flags=L.F_OUTER)["variant"] == "generic"
    AO.assert_close(ops.softmax_fwd(xo.cuda(), dim=0), AO.softmax_fwd(xo, dim=0), 32 * 32, "odd row bytes")
    xl = torch.randn(1 << 18, 8).to(dt)
    AO.assert_close(ops.softmax_fwd(xl.cuda(), dim=0), AO.softmax_fwd(xl, dim=0), (1 << 18) ** 2, "M = 2^18")
    with pytest.raises(L.UnsupportedSegment):
        ops.softmax_fwd(xl.cuda(), dim=0, variant=L.VARIANT_CLUSTER)
    with pytest.raises(L.UnsupportedSegment):
        ops.softmax_fwd(torch.randn(8, 8, 8).to(dt).cuda(), dim=1)


def test_outer_deterministic_and_graph_replay():
    """Same bits on every run, also when replayed from a CUDA graph (mbarrier phases and cluster
    slot sets start from rest in every launch)."""
    dt = torch.bfloat16
    x = torch.randn(1024, 2048).to(dt).cuda()
    ref = ops.softmax_fwd(x, dim=0).clone()
    for _ in range(3):
        assert torch.equal(ops.softmax_fwd(x, dim=0), ref)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):  # warm-up on the capture stream: workspace, kernel attributes
            ops.softmax_fwd(x, dim=0)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        out = ops.softmax_fwd(x, dim=0)
    for it in range(3):
        out.zero_()
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, ref), f"replay {it} differs"


@pytest.mark.parametrize("dt", [DataType.BFloat16, DataType.Float], ids=["bf16", "f32"])
def test_outer_frontend_fusions(dt):
    """The reference's outer_persistent benchmark definitions, recorded through FusionDefinition,
    matched, executed and validated against the node-by-node ATen evaluation (testValidate)."""
    from oracle import fusion_eval

    td = torch.bfloat16 if dt == DataType.BFloat16 else torch.float32
    m, n = 512, 1536
    x = torch.randn(m, n).to(td).cuda()
    dy = torch.randn(m, n).to(td).cuda()
    with FusionDefinition() as fd:
        F.softmax_fwd_py(fd, dt, reduction_axis=0)
    assert fd.segment().flags & L.F_OUTER
    (y,) = fd.validate([x], fusion_eval.evaluate(fd, [x.cpu()]))
    AO.assert_close(y, torch.softmax(x.float(), 0).to(td).cpu(), m * m, "outer softmax vs eager")
    with FusionDefinition() as fd:
        F.softmax_bwd_py(fd, dt, reduction_axis=0)
    assert fd.segment().flags & L.F_OUTER
    fd.validate([y, dy], fusion_eval.evaluate(fd, [y.cpu(), dy.cpu()]))


def test_outer_full_size_properties():
    """The reference's outer_persistent benchmark shape class ([8192, N] / [2048, N], bf16):
    columns of the softmax sum to 1, dx sums to 0 over every column, elementwise agreement with eager."""
    dt = torch.bfloat16
    for m, n in ((8192, 4096), (2048, 16384)):
        x = torch.randn(m, n, device="cuda").to(dt)
        sm = ops.softmax_fwd(x, dim=0)
        assert (sm.float().sum(0) - 1).abs().max().item() < 2e-2
        ref = torch.softmax(x.float(), 0)
        assert (sm.float() - ref).abs().max().item() <= 2.0 ** -8 * ref.max().item() + 1e-6
        dy = torch.randn(m, n, device="cuda").to(dt)
        dx = ops.softmax_bwd(dy, sm, dim=0)
        scale = (sm.float() * dy.float()).abs().sum(0)
        assert (dx.float().sum(0).abs() / scale).max().item() < 2e-2


def c_outer_norm_fwd(x, w, b, eps):
    """fp64: statistics from the C oracle on the transpose (a column is a row of it), affine applied here."""
    yh, mean, invstd = CO.layer_norm_fwd(x.t().contiguous(), None, None, eps)
    y = yh.t()
    if w is not None:
        y = y * w.double()
    if b is not None:
        y = y + b.double()
    return y, mean.view(-1), invstd.view(-1)


@pytest.mark.parametrize("dt", DTYPES, ids=["bf16", "f32", "f16"])
@pytest.mark.parametrize("shape", [(1, 64), (16, 768), (33, 1000), (216, 96), (512, 2048), (600, 136), (4100, 512)],
                         ids=lambda s: f"{s[0]}x{s[1]}")
def test_outer_norm_fwd(shape, dt):
    """Normalization over dim 0 with a per-column affine (batch-norm-shaped forward, channels last)."""
    m, n = shape
    x, _, w, b = make_inputs(m, n, dt, seed=m * 17 + n)
    eps = 1e-5
    for ww, bb in ((w, b), (w, None), (None, None)):
        y, mean, invstd = ops.outer_norm_fwd(x.cuda(), None if ww is None else ww.cuda(), None if bb is None else bb.cuda(), eps)
        assert y.shape == (m, n) and mean.shape == (n,) and invstd.shape == (n,) and mean.dtype == torch.float32
        ye, me, ie = AO.outer_norm_fwd(x, ww, bb, eps)
        yc, mc, ic = c_outer_norm_fwd(x, ww, bb, eps)
        fl = 1.0 if dt == torch.float32 else 1e-2
        check(y, ye, yc, m, "outer norm y", fl)
        check(mean, me, mc, m, "outer norm mean", 1.0)  # a sum with cancellation: absolute accuracy, as for the inner mean
        check(invstd, ie, ic, m, "outer norm invstd", 1e-3)


@pytest.mark.parametrize("cluster", [1, 2, 4, 8, 16])
def test_outer_norm_cluster_split_and_batchnorm_pin(cluster):
    """Every cluster size; the oracle itself is pinned to at::native_batch_norm (training statistics) on the
    NHWC view, the explicit ATen reference the reference's batch-norm tests call."""
    dt = torch.bfloat16
    nb, h, wd, c = 6, 8, 16, 384  # [N,H,W,C] channels last -> [768, 384]
    g = torch.Generator().manual_seed(cluster)
    x4 = torch.randn(nb, h, wd, c, generator=g).to(dt)
    w = torch.randn(c, generator=g).to(dt)
    b = torch.randn(c, generator=g).to(dt)
    eps = 1e-5
    ye, me, ie = AO.outer_norm_fwd(x4, w, b, eps)
    yn, mn, inn = torch.native_batch_norm(x4.float().permute(0, 3, 1, 2), w.float(), b.float(), None, None, True, 0.1, eps)
    assert torch.allclose(ye.float(), yn.permute(0, 2, 3, 1).to(dt).float(), atol=2e-2, rtol=2e-2)
    assert torch.allclose(me, mn, atol=1e-5) and torch.allclose(ie, inn, rtol=1e-4)
    y, mean, invstd = ops.outer_norm_fwd(x4.cuda(), w.cuda(), b.cuda(), eps, cluster_size=cluster)
    torch.cuda.synchronize()
    assert y.shape == x4.shape
    m = nb * h * wd
    AO.assert_close(y, ye, m, f"outer norm 4-D, cluster {cluster}")
    AO.assert_close(mean, me, m, "mean")
    AO.assert_close(invstd, ie, m, "invstd")
    yc, mc, ic = c_outer_norm_fwd(x4.view(m, c), w, b, eps)
    assert ulp_err(y.view(m, c).cpu(), yc, floor=1e-2) <= TIGHT[dt]
    assert ulp_err(mean.cpu(), mc, floor=1.0) <= 64 and ulp_err(invstd.cpu(), ic, floor=1e-3) <= 64


@pytest.mark.parametrize("dt", [DataType.BFloat16, DataType.Float], ids=["bf16", "f32"])
def test_batchnorm_fwd_frontend(dt):
    """The reference's batch-norm forward definition (benchmarks/python/normalization.py, channels last) through
    FusionDefinition: matched, executed on the outer-persistent kernel, validated against the node-by-node ATen
    evaluation under the validator's tolerances; the unused running statistics are accepted as inputs."""
    from oracle import fusion_eval

    td = torch.bfloat16 if dt == DataType.BFloat16 else torch.float32
    nb, h, wd, c = 8, 14, 14, 256
    x = torch.randn(nb, h, wd, c).to(td).cuda()
    w, b = torch.randn(c).to(td).cuda(), torch.randn(c).to(td).cuda()
    rm, rv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
    with FusionDefinition() as fd:
        F.batchnorm_fwd_py(fd, dt, num_dims=4)
    assert fd.segment().flags & L.F_OUTER
    ins = [x, w, b, rm, rv]
    y, mean, invstd = fd.validate(ins, fusion_eval.evaluate(fd, [t.cpu() for t in ins]))
    assert y.shape == x.shape and mean.shape == (c,) and invstd.shape == (c,)
    yn = torch.nn.functional.batch_norm(x.float().permute(0, 3, 1, 2), None, None, w.float(), b.float(), True, 0.1, 1e-5)
    AO.assert_close(y, yn.permute(0, 2, 3, 1).to(td).cpu(), nb * h * wd, "batch norm vs eager")


@pytest.mark.parametrize("dt", DTYPES, ids=["bf16", "f32", "f16"])
@pytest.mark.parametrize("shape", [(1, 64), (16, 768), (33, 1000), (216, 96), (512, 2048), (600, 136), (4100, 512)],
                         ids=lambda s: f"{s[0]}x{s[1]}")
def test_outer_norm_bwd(shape, dt):
    """Backward of the outer normalization (batch-norm backward, training, channels last): dx + dweight + dbias."""
    m, n = shape
    x, dy, w, _ = make_inputs(m, n, dt, seed=m * 19 + n)
    _, mean, invstd = AO.outer_norm_fwd(x, w, None, 1e-5)
    dx, dw, db = ops.outer_norm_bwd(dy.cuda(), x.cuda(), mean.cuda(), invstd.cuda(), w.cuda())
    assert dx.shape == (m, n) and dw.shape == (n,) and db.shape == (n,) and dw.dtype == dt
    dxe, dwe, dbe = AO.outer_norm_bwd(dy, x, mean, invstd, w)
    dxc, dwc, dbc = AO.outer_norm_bwd(dy.double(), x.double(), mean.double(), invstd.double(), w.double())
    check(dx, dxe, dxc, m, "outer norm dx", 1.0 if dt == torch.float32 else 1e-1)
    check(dw, dwe, dwc, m, "outer norm dw", max(1.0, m ** 0.5))
    check(db, dbe, dbc, m, "outer norm db", max(1.0, m ** 0.5))
    # output masks
    dx2, dw2, db2 = ops.outer_norm_bwd(dy.cuda(), x.cuda(), mean.cuda(), invstd.cuda(), w.cuda(), (True, False, False))
    assert dw2 is None and db2 is None and torch.equal(dx2, dx)
    dx3, dw3, db3 = ops.outer_norm_bwd(dy.cuda(), x.cuda(), mean.cuda(), invstd.cuda(), w.cuda(), (False, True, True))
    assert dx3 is None and torch.equal(dw3, dw) and torch.equal(db3, db)


@pytest.mark.parametrize("cluster", [1, 2, 4, 8, 16])
def test_outer_norm_bwd_cluster_split(cluster):
    dt = torch.bfloat16
    m, n = 777, 384
    x, dy, w, _ = make_inputs(m, n, dt, seed=cluster + 100)
    _, mean, invstd = AO.outer_norm_fwd(x, w, None, 1e-5)
    try:
        ops.plan_for(L.OP_LAYERNORM_BWD, dt, m, n, flags=L.F_OUTER | L.F_HAS_WEIGHT | L.F_OUT_DX, cluster_size=cluster)
    except L.UnsupportedSegment:
        pytest.skip(f"M = {m} does not fit a cluster of {cluster}")
    dx, dw, db = ops.outer_norm_bwd(dy.cuda(), x.cuda(), mean.cuda(), invstd.cuda(), w.cuda(), cluster_size=cluster)
    dxe, dwe, dbe = AO.outer_norm_bwd(dy, x, mean, invstd, w)
    dxc, dwc, dbc = AO.outer_norm_bwd(dy.double(), x.double(), mean.double(), invstd.double(), w.double())
    check(dx, dxe, dxc, m, f"dx C={cluster}", 1e-1)
    check(dw, dwe, dwc, m, f"dw C={cluster}", m ** 0.5)
    check(db, dbe, dbc, m, f"db C={cluster}", m ** 0.5)


@pytest.mark.parametrize("dt", [DataType.BFloat16, DataType.Float], ids=["bf16", "f32"])
def test_batchnorm_bwd_frontend(dt):
    """The reference's batch-norm backward definition through FusionDefinition (7 inputs, two of them unused),
    validated against the node-by-node ATen evaluation and against at::native_batch_norm_backward."""
    from oracle import fusion_eval

    td = torch.bfloat16 if dt == DataType.BFloat16 else torch.float32
    nb, h, wd, c = 8, 14, 14, 256
    x = torch.randn(nb, h, wd, c).to(td)
    dy = torch.randn(nb, h, wd, c).to(td)
    w = torch.randn(c).to(td)
    _, mean, invstd = AO.outer_norm_fwd(x, w, None, 1e-5)
    ins = [x, dy, w, torch.zeros(c), torch.ones(c), mean, invstd]
    with FusionDefinition() as fd:
        F.batchnorm_bwd_py(fd, dt, num_dims=4)
    assert fd.segment().op == L.OP_LAYERNORM_BWD and fd.segment().flags & L.F_OUTER
    dx, dw, db = fd.validate([t.cuda() for t in ins], fusion_eval.evaluate(fd, ins))
    assert dx.shape == x.shape and dw.shape == (c,) and db.shape == (c,)
    r = torch.ops.aten.native_batch_norm_backward(dy.float().permute(0, 3, 1, 2), x.float().permute(0, 3, 1, 2), w.float(),
                                                   None, None, mean, invstd, True, 1e-5, [True, True, True])
    m = nb * h * wd
    AO.assert_close(dx, r[0].permute(0, 2, 3, 1).to(td), m, "dx vs native_batch_norm_backward")
    AO.assert_close(dw, r[1].to(td), m, "dweight")
    AO.assert_close(db, r[2].to(td), m, "dbias")


GENERIC_SHAPES = [(1, 3), (7, 1), (33, 1987), (216, 96), (600, 130), (2048, 264), (5000, 37)]


@pytest.mark.parametrize("dt", DTYPES, ids=["bf16", "f32", "f16"])
@pytest.mark.parametrize("shape", GENERIC_SHAPES, ids=lambda s: f"{s[0]}x{s[1]}")
def test_outer_generic_family(shape, dt):
    """The shape-agnostic two-kernel family (outer_generic.cuh: single-pass partials per row split — online
    softmax, Welford — merged in split order, then an apply pass): every outer op at shapes no vectorised path
    takes (N = 3, 1987, 37 ...), forced with VARIANT_GENERIC, against both oracles."""
    m, n = shape
    x, dy, w, b = make_inputs(m, n, dt, seed=m * 23 + n)
    assert ops.plan_for(L.OP_SOFTMAX_FWD, dt, m, n, flags=L.F_OUTER, variant=L.VARIANT_GENERIC)["variant"] == "generic"
    run_all(x, dy, dt, m, f"generic [{m},{n}]", variant=L.VARIANT_GENERIC)
    eps = 1e-5
    y, mean, invstd = ops.outer_norm_fwd(x.cuda(), w.cuda(), b.cuda(), eps, variant=L.VARIANT_GENERIC)
    ye, me, ie = AO.outer_norm_fwd(x, w, b, eps)
    yc, mc, ic = c_outer_norm_fwd(x, w, b, eps)
    check(y, ye, yc, m, "generic outer norm y", 1.0 if dt == torch.float32 else 1e-2)
    check(mean, me, mc, m, "generic outer norm mean", 1.0)
    check(invstd, ie, ic, m, "generic outer norm invstd", 1e-3)
    dx, dw, db = ops.outer_norm_bwd(dy.cuda(), x.cuda(), me.cuda(), ie.cuda(), w.cuda(), variant=L.VARIANT_GENERIC)
    dxe, dwe, dbe = AO.outer_norm_bwd(dy, x, me, ie, w)
    dxc, dwc, dbc = AO.outer_norm_bwd(dy.double(), x.double(), me.double(), ie.double(), w.double())
    check(dx, dxe, dxc, m, "generic outer norm dx", 1.0 if dt == torch.float32 else 1e-1)
    check(dw, dwe, dwc, m, "generic outer norm dw", max(1.0, m ** 0.5))
    check(db, dbe, dbc, m, "generic outer norm db", max(1.0, m ** 0.5))


def test_outer_generic_semantics_and_resnet_shape():
    """NaN / +-inf columns through the online softmax of the generic family (a leading -inf run, a -inf-only row
    split, +inf, NaN), and a ResNet-size batch norm ([32*56*56, 64]: far beyond a 16-CTA cluster) through the
    default dispatch."""
    dt = torch.bfloat16
    xs = torch.randn(300, 40).to(dt)
    xs[7, 1] = float("nan")
    xs[:, 2] = float("-inf")
    xs[5, 3] = float("inf")
    xs[:150, 5] = float("-inf")
    xs[100:, 6] = float("-inf")
    for log in (False, True):
        y = ops.softmax_fwd(xs.cuda(), dim=0, log=log, variant=L.VARIANT_GENERIC)
        torch.cuda.synchronize()
        ref = AO.log_softmax_fwd(xs, dim=0) if log else AO.softmax_fwd(xs, dim=0)
        AO.assert_close(y, ref, 300 * 300, f"generic NaN/Inf columns log={log}")
        assert torch.isnan(y[:, 1:4]).all() and not torch.isnan(y[:, 0]).any() and not torch.isnan(y[:, 4:]).any()
    m, c = 32 * 56 * 56, 64
    x = torch.randn(32, 56, 56, c).to(dt)
    w, b = torch.randn(c).to(dt), torch.randn(c).to(dt)
    assert ops.plan_for(L.OP_LAYERNORM_FWD, dt, m, c, flags=L.F_OUTER | L.F_HAS_WEIGHT | L.F_HAS_BIAS)["variant"] == "generic"
    y, mean, invstd = ops.outer_norm_fwd(x.cuda(), w.cuda(), b.cuda(), 1e-5)
    ye, me, ie = AO.outer_norm_fwd(x, w, b, 1e-5)
    AO.assert_close(y, ye, m, "ResNet-size batch norm y")
    assert torch.allclose(mean.cpu(), me, atol=1e-5) and torch.allclose(invstd.cpu(), ie, rtol=1e-4)
    dy = torch.randn(32, 56, 56, c).to(dt)
    dx, dw, db = ops.outer_norm_bwd(dy.cuda(), x.cuda(), me.cuda(), ie.cuda(), w.cuda())
    dxe, dwe, dbe = AO.outer_norm_bwd(dy, x, me, ie, w)
    AO.assert_close(dx, dxe, m, "ResNet-size batch norm dx")
    AO.assert_close(dw, dwe, m, "dweight")
    AO.assert_close(db, dbe, m, "dbias")


ROWS_SHAPES = [(100352, 256, torch.bfloat16), (150000, 64, torch.float16), (120001, 96, torch.float32), (105000, 1024, torch.bfloat16),
               (60000, 128, torch.float32), (130001, 24, torch.bfloat16), (401408, 64, torch.bfloat16), (70000, 2048, torch.float32),
               (66000, 4096, torch.bfloat16)]


@pytest.mark.parametrize("shape", ROWS_SHAPES, ids=lambda s: f"{s[0]}x{s[1]}-{str(s[2]).split('.')[-1]}")
def test_outer_rows_family(shape):
    """Few columns, many rows — the batch norms of convolutional nets in channels-last view ([N*H*W, C]) that no cluster
    holds on chip: the single-kernel row-block family (outer_rows.cuh), forward and backward against both oracles,
    chunk counts that divide the CTA (256 columns), leave idle threads (24 -> 3 chunks), fill it (4096 bf16 / 2048 f32);
    the output masks; the three-kernel family on the same inputs (B200NORM_NO_OUTER_ROWS) as a cross-check."""
    m, n, dt = shape
    x, dy, w, b = make_inputs(m, n, dt, seed=m + n)
    x = (x.float() * 1.5 + 0.75).to(dt)  # a non-zero column mean
    eps = 1e-5
    plan = ops.plan_for(L.OP_LAYERNORM_FWD, dt, m, n, flags=L.F_OUTER | L.F_HAS_WEIGHT | L.F_HAS_BIAS)
    assert plan["variant"] == "generic" and plan["n_launches"] == 1 and plan["block"] == 512, plan
    bplan = ops.plan_for(L.OP_LAYERNORM_BWD, dt, m, n, flags=L.F_OUTER | L.F_HAS_WEIGHT | L.F_OUT_DX | L.F_OUT_DW | L.F_OUT_DB)
    assert bplan["variant"] == "generic" and bplan["n_launches"] == 1, bplan
    xg, dyg, wg, bg = x.cuda(), dy.cuda(), w.cuda(), b.cuda()
    ye, me, ie = AO.outer_norm_fwd(x, w, b, eps)
    yc, mc, ic = c_outer_norm_fwd(x, w, b, eps)
    n0 = L.launch_count()
    y, mean, invstd = ops.outer_norm_fwd(xg, wg, bg, eps)
    assert L.launch_count() - n0 == 1
    fl = 1.0 if dt == torch.float32 else 1e-2
    check(y, ye, yc, m, "rows outer norm y", fl)
    check(mean, me, mc, m, "rows outer norm mean", 1.0)
    check(invstd, ie, ic, m, "rows outer norm invstd", 1e-3)
    y2, mean2, invstd2 = ops.outer_norm_fwd(xg, None, None, eps)
    ye2, _, _ = AO.outer_norm_fwd(x, None, None, eps)
    AO.assert_close(y2, ye2, m, "rows outer norm without affine")
    assert torch.equal(mean2, mean) and torch.equal(invstd2, invstd)
    # run-to-run determinism (fixed merge order over threads and CTAs)
    y3, mean3, invstd3 = ops.outer_norm_fwd(xg, wg, bg, eps)
    assert torch.equal(y3, y) and torch.equal(mean3, mean) and torch.equal(invstd3, invstd)

    dx, dw, db = ops.outer_norm_bwd(dyg, xg, me.cuda(), ie.cuda(), wg)
    dxe, dwe, dbe = AO.outer_norm_bwd(dy, x, me, ie, w)
    dxc, dwc, dbc = AO.outer_norm_bwd(dy.double(), x.double(), me.double(), ie.double(), w.double())
    check(dx, dxe, dxc, m, "rows outer norm dx", 1.0 if dt == torch.float32 else 1e-1)
    check(dw, dwe, dwc, m, "rows outer norm dw", m ** 0.5)
    check(db, dbe, dbc, m, "rows outer norm db", m ** 0.5)
    dx2, dw2, db2 = ops.outer_norm_bwd(dyg, xg, me.cuda(), ie.cuda(), wg, (True, False, False))
    assert dw2 is None and db2 is None and torch.equal(dx2, dx)
    dx3, dw3, db3 = ops.outer_norm_bwd(dyg, xg, me.cuda(), ie.cuda(), wg, (False, True, True))
    assert dx3 is None and torch.equal(dw3, dw) and torch.equal(db3, db)

    # a two-slot ring (every tile but the last two goes through a refilled slot, phase 2 reloads almost everything):
    # same arithmetic in the same order, bit-equal results
    os.environ["B200NORM_OUTER_ROWS_DEPTH"] = "2"
    try:
        y4, mean4, invstd4 = ops.outer_norm_fwd(xg, wg, bg, eps)
        dx4, dw4, db4 = ops.outer_norm_bwd(dyg, xg, me.cuda(), ie.cuda(), wg)
        torch.cuda.synchronize()
    finally:
        del os.environ["B200NORM_OUTER_ROWS_DEPTH"]
    assert torch.equal(y4, y) and torch.equal(mean4, mean) and torch.equal(invstd4, invstd)
    assert torch.equal(dx4, dx) and torch.equal(dw4, dw) and torch.equal(db4, db)

    os.environ["B200NORM_NO_OUTER_ROWS"] = "1"
    try:
        assert ops.plan_for(L.OP_LAYERNORM_FWD, dt, m, n, flags=L.F_OUTER | L.F_HAS_WEIGHT | L.F_HAS_BIAS)["n_launches"] == 3
        yg, mg, ig = ops.outer_norm_fwd(xg, wg, bg, eps)
        dxg, dwg, dbg = ops.outer_norm_bwd(dyg, xg, me.cuda(), ie.cuda(), wg)
        torch.cuda.synchronize()
    finally:
        del os.environ["B200NORM_NO_OUTER_ROWS"]
    AO.assert_close(y, yg.cpu(), m, "rows vs three-kernel family y")
    AO.assert_close(dx, dxg.cpu(), m, "rows vs three-kernel family dx")
    assert torch.allclose(mean, mg, atol=1e-5) and torch.allclose(invstd, ig, rtol=1e-4)


def test_outer_rows_unaligned_falls_back():
    """A view whose rows are not contiguous, or whose base is not 16-byte aligned, goes to the three-kernel family
    (the row-block kernel declines at launch time); results agree with the oracle either way."""
    dt = torch.bfloat16
    m, n = 150000, 64  # beyond the cluster-persistent family
    base = torch.randn(m, n + 8).to(dt)
    w, b = torch.randn(n).to(dt), torch.randn(n).to(dt)
    xg = base.cuda()[:, :n]          # row stride n + 8
    n0 = L.launch_count()
    y, mean, invstd = ops.outer_norm_fwd(xg, w.cuda(), b.cuda(), 1e-5)
    assert L.launch_count() - n0 == 3
    ye, me, ie = AO.outer_norm_fwd(base[:, :n].contiguous(), w, b, 1e-5)
    AO.assert_close(y, ye, m, "strided rows: three-kernel family")
    flat = torch.randn(m * n + 4).to(dt).cuda()
    xo = flat[4:].view(m, n)         # base 8 bytes off a 16-byte boundary
    n0 = L.launch_count()
    y, mean, invstd = ops.outer_norm_fwd(xo, w.cuda(), b.cuda(), 1e-5)
    assert L.launch_count() - n0 == 3
    ye, me, ie = AO.outer_norm_fwd(xo.cpu(), w, b, 1e-5)
    AO.assert_close(y, ye, m, "misaligned base: three-kernel family")


@pytest.mark.parametrize("variant", [L.VARIANT_AUTO, L.VARIANT_GENERIC], ids=["auto", "generic"])
@pytest.mark.parametrize("path", outer_golden_files(), ids=lambda p: p.split("/")[-1][:-4])
def test_outer_golden_fixtures(path, variant):
    """The committed ATen-native fixtures of the outer ops (tests/golden/make_golden_outer.py) replayed against
    the CUDA kernels — persistent and shape-agnostic families."""
    dt, m, n, g = load_golden(path)
    x, dy, w, b = (g[k].cuda() for k in ("x", "dy", "w", "b"))
    kw = dict(variant=variant)
    AO.assert_close(ops.softmax_fwd(x, dim=0, **kw), g["osm_y"], m * m, "outer softmax y")
    AO.assert_close(ops.softmax_fwd(x, dim=0, log=True, **kw), g["olsm_y"], m * m, "outer log_softmax y")
    AO.assert_close(ops.softmax_bwd(dy, g["osm_y"].cuda(), dim=0, **kw), g["osm_dx"], m, "outer softmax dx")
    AO.assert_close(ops.softmax_bwd(dy, g["olsm_y"].cuda(), dim=0, log=True, **kw), g["olsm_dx"], m, "outer log_softmax dx")
    y, mean, invstd = ops.outer_norm_fwd(x, w, b, g["eps"], **kw)
    AO.assert_close(y, g["bn_y"], m, "batch norm y")
    AO.assert_close(mean, g["bn_mean"], m, "batch norm mean")
    AO.assert_close(invstd, g["bn_invstd"], m, "batch norm invstd")
    dx, dw, db = ops.outer_norm_bwd(dy, x, g["bn_mean"].cuda(), g["bn_invstd"].cuda(), w, **kw)
    AO.assert_close(dx, g["bn_dx"], m, "batch norm dx")
    AO.assert_close(dw, g["bn_dw"], m, "batch norm dweight")
    AO.assert_close(db, g["bn_db"], m, "batch norm dbias")

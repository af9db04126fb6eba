"""Vocabulary-size rows and odd row sizes (run with -m gpu): the hidden sizes of the reference's cluster-reduction
softmax test (tests/cpp/test_cluster.cpp:158-211: DeepSeek-R1 129280, Llama3 128256, Llama4 202048, Gemma2 256000,
Mistral 131072, Qwen2 152064, Phi4 100352 and 1024 * 1024, batch = SMs / 8, inputs randn clamped to [-2, 2]) on the
cluster kernels — N = 1048576 bf16 with ONE 128 KB stage per CTA of a 16-CTA cluster — and the odd-size set
N + 2 ... N + 16 of benchmarks/python/global_params.py:86-88 on the shape-agnostic family."""
import pytest
import torch

from fuser_b200 import _lib as L
from fuser_b200 import ops
from oracle import aten_oracle as AO
from oracle import c_oracle as CO
from tests.util import ulp_err

pytestmark = pytest.mark.gpu

VOCAB = [129280, 128256, 202048, 256000, 131072, 152064, 100352, 1024 * 1024]
TIGHT = {torch.bfloat16: 1.01, torch.float16: 1.01, torch.float32: 64.0}
ROWS = 18


def _x(m, n, dt, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(m, n, generator=g).clamp(-2, 2).to(dt)


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float32], ids=["bf16", "f32"])
@pytest.mark.parametrize("n", VOCAB, ids=str)
def test_vocab_softmax_forward(n, dt):
    x = _x(ROWS, n, dt, n)
    xd = x.cuda()
    plan = ops.plan_for(L.OP_SOFTMAX_FWD, dt, ROWS, n)
    if n * x.element_size() <= 16 * 200 * 1024:
        assert plan["variant"] == "cluster", plan  # every size of the reference's test that fits a 16-CTA cluster
    if n == 1024 * 1024 and dt == torch.bfloat16:
        assert plan["cluster_size"] == 16 and plan["stages"] == 1
    y = ops.softmax_fwd(xd)
    ly = ops.softmax_fwd(xd, log=True)
    torch.cuda.synchronize()
    AO.assert_close(y, AO.softmax_fwd(x), n * n, "softmax y")
    AO.assert_close(ly, AO.log_softmax_fwd(x), n * n, "log_softmax y")
    assert ulp_err(y.cpu(), CO.softmax_fwd(x), floor=1e-7 if dt != torch.float32 else 1e-5) <= TIGHT[dt]
    assert ulp_err(ly.cpu(), CO.softmax_fwd(x, log=True), floor=1.0) <= TIGHT[dt]
    assert (y.float().sum(-1) - 1).abs().max().item() < 3e-2


@pytest.mark.parametrize("n", [131072, 256000, 1024 * 1024], ids=str)
def test_vocab_softmax_backward_and_norms(n):
    dt = torch.bfloat16
    x = _x(ROWS, n, dt, n + 1)
    dy = _x(ROWS, n, dt, n + 2)
    xd, dyd = x.cuda(), dy.cuda()
    y = AO.softmax_fwd(x)
    dx = ops.softmax_bwd(dyd, y.cuda())
    torch.cuda.synchronize()
    AO.assert_close(dx, AO.softmax_bwd(dy, y), n, "softmax dx")
    assert ulp_err(dx.cpu(), CO.softmax_bwd(dy, y), floor=1e-6) <= TIGHT[dt]
    # the other one-stream forward ops on the same rows
    w = _x(1, n, dt, n + 3).view(n)
    yr, st = ops.rms_norm_fwd(xd, w.cuda(), 1e-5)
    ye, se = AO.rms_norm_fwd(x, w, 1e-5)
    AO.assert_close(yr, ye, n, "rms y")
    AO.assert_close(st, se, n, "rms invstd")
    yl, mean, invstd = ops.layer_norm_fwd(xd, w.cuda(), w.cuda(), 1e-5)
    ye, me, ie = AO.layer_norm_fwd(x, w, w, 1e-5)
    AO.assert_close(yl, ye, n, "ln y")
    AO.assert_close(mean, me, n, "ln mean")
    AO.assert_close(invstd, ie, n, "ln invstd")
    r = ops.reduce_inner(xd, "sum")
    AO.assert_close(r, AO.reduce_inner(x, "sum"), n, "row sum")
    assert torch.equal(ops.reduce_inner(xd, "max").cpu(), torch.amax(x, -1))


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float32], ids=["bf16", "f32"])
@pytest.mark.parametrize("base", [4096, 8192], ids=str)
def test_odd_row_sizes(base, dt):
    """hs + 2, + 4, + 8, + 16 (benchmarks/python/global_params.py:86-88): rows that are not a multiple of the
    16-byte vector run on the shape-agnostic family, the others on the ring; same results either way."""
    m = 300
    for n in (base + 2, base + 4, base + 8, base + 16):
        x = _x(m, n, dt, n)
        dy = _x(m, n, dt, n + 1)
        w = _x(1, n, dt, n + 2).view(n)
        xd, dyd, wd = x.cuda(), dy.cuda(), w.cuda()
        y = ops.softmax_fwd(xd)
        AO.assert_close(y, AO.softmax_fwd(x), n * n, f"softmax y N={n}")
        AO.assert_close(ops.softmax_bwd(dyd, y), AO.softmax_bwd(dy, y.cpu()), n, f"softmax dx N={n}")
        yr, rms = ops.rms_norm_fwd(xd, wd, 1e-5, stat_is_rms=True)
        ye, re_ = AO.rms_norm_fwd_python(x, w, 1e-5)
        AO.assert_close(yr, ye, n, f"rms y N={n}")
        dx, dw = ops.rms_norm_bwd(dyd, xd, rms, wd, stat_is_rms=True)
        dxe, dwe = AO.rms_norm_bwd_python(x, re_, dy, w)
        AO.assert_close(dx, dxe, n, f"rms dx N={n}")
        AO.assert_close(dw, dwe, m, f"rms dw N={n}")
        yl, mean, invstd = ops.layer_norm_fwd(xd, wd, wd, 1e-5)
        ye, me, ie = AO.layer_norm_fwd(x, w, w, 1e-5)
        AO.assert_close(yl, ye, n, f"ln y N={n}")
        dxl, dwl, dbl = ops.layer_norm_bwd(dyd, xd, mean, invstd, wd)
        dxe, dwe, dbe = AO.layer_norm_bwd(dy, x, me, ie, w)
        AO.assert_close(dxl, dxe, n, f"ln dx N={n}")
        AO.assert_close(dwl, dwe, m, f"ln dw N={n}")
        AO.assert_close(dbl, dbe, m, f"ln db N={n}")
        aligned = (n * x.element_size()) % 16 == 0
        assert (ops.plan_for(L.OP_SOFTMAX_FWD, dt, m, n)["variant"] != "generic") == aligned


# The pipelined one-exchange cluster kernel (kernels_softmax_pipe.cuh) with MANY rows per cluster — ring refills,
# mbarrier phase wraps, both slot-set parities — at natural and forced cluster sizes, register (ch > 0) and looping
# (ch = 0) forms, and rows that hold -inf / +inf / NaN in one CTA's slice only.
PIPE_CASES = [
    (600, 131072, torch.bfloat16, 0),   # natural plan of the vocabulary benchmark shape
    (600, 131072, torch.bfloat16, 8),
    (300, 131072, torch.bfloat16, 16),
    (700, 65536, torch.bfloat16, 2),    # 64 KB slices
    (500, 98304, torch.bfloat16, 2),    # 96 KB slices: looping form, two stages
    (300, 152064, torch.bfloat16, 0),
    (400, 65536, torch.float32, 0),
    (250, 131072, torch.float32, 0),
    (300, 256000, torch.float16, 0),
    (500, 100352, torch.bfloat16, 0),   # 6 * 512 + 64 chunks per CTA: the balanced tail (4-byte pieces)
    (300, 100352, torch.float32, 0),
    (300, 202048, torch.float16, 0),    # 6 * 512 + 85
    (1500, 32064, torch.bfloat16, 1),   # 32 K vocabularies: the same kernel without a cluster (rows of 48-75 KB)
    (1200, 32768, torch.float16, 1),
    (900, 16384, torch.float32, 1),
    (2000, 24576, torch.bfloat16, 1),   # one-group rows of the register kernel, where the pipelined one measured faster
    (2000, 16384, torch.float16, 1),
    (1500, 8192, torch.float32, 1),
]


@pytest.mark.parametrize("m,n,dt,c", PIPE_CASES, ids=lambda v: str(v).replace("torch.", ""))
def test_cluster_softmax_pipe_many_rows(m, n, dt, c):
    x = _x(m, n, dt, m + n)
    x = (x.float() * 3).to(dt)
    # special rows: one slice all -inf, a whole row -inf, +inf and NaN somewhere past the first slice
    x[3, : n // 16] = float("-inf")
    x[5, :] = float("-inf")
    x[7, n - 5] = float("inf")
    x[9, n // 2 + 1] = float("nan")
    x[11, : n - 1] = float("-inf")
    xd = x.cuda()
    plan = ops.plan_for(L.OP_SOFTMAX_FWD, dt, m, n, cluster_size=c)
    assert plan["variant"] == ("tma" if c == 1 else "cluster") and plan["pipe"], plan
    if c:
        assert plan["cluster_size"] == c
    if c == 1:
        assert ops.plan_for(L.OP_SOFTMAX_FWD, dt, m, n) == plan, "the natural plan of these rows"
    assert m > plan["stages"] * (plan["grid"] // plan["cluster_size"]) * 2  # several ring revolutions per cluster
    lplan = ops.plan_for(L.OP_LOGSOFTMAX_FWD, dt, m, n, cluster_size=c)
    assert lplan["pipe"], lplan
    y = ops.softmax_fwd(xd, cluster_size=c)
    ly = ops.softmax_fwd(xd, log=True, cluster_size=c)
    torch.cuda.synchronize()
    ye, lye = AO.softmax_fwd(x), AO.log_softmax_fwd(x)
    AO.assert_close(y, ye, n * n, "softmax y")
    AO.assert_close(ly, lye, n * n, "log_softmax y")
    assert torch.equal(torch.isnan(y.cpu()), torch.isnan(ye)) and torch.equal(torch.isnan(ly.cpu()), torch.isnan(lye))
    assert torch.equal(torch.isinf(ly.cpu()), torch.isinf(lye))
    fin = torch.ones(m, dtype=torch.bool)
    fin[[5, 7, 9]] = False
    assert ulp_err(y.cpu()[fin], CO.softmax_fwd(x[fin]), floor=1e-7 if dt != torch.float32 else 1e-5) <= TIGHT[dt] * 2
    assert (y.float().sum(-1)[fin.cuda()] - 1).abs().max().item() < 3e-2
    # the same bits from run to run and against a different cluster size (the merge order is fixed per plan)
    assert torch.equal(ops.softmax_fwd(xd, cluster_size=c).view(torch.uint8), y.view(torch.uint8))

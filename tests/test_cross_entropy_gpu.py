"""GPU parity of the fused cross-entropy path — log-softmax + NLL over vocabulary-size rows (SURVEY 8f rank 2,
benchmarks/python/cross_entropy_loss.py) — through the C ABI: per-row loss and log-sum-exp against
torch.nn.functional (log_softmax + nll_loss on the Float upcast), the mean / sum reductions against
F.cross_entropy, and the backward against autograd. Shapes: the vocabularies of the benchmark's models
(Qwen2 151 936, Phi-3 32 064, Mistral-Nemo 131 072), rows split over clusters, ragged and unaligned rows,
ignored targets."""
import pytest
import torch

from fuser_b200 import _lib as L
from fuser_b200 import ops
from oracle import aten_oracle as AO

pytestmark = pytest.mark.gpu

DTYPES = [torch.bfloat16, torch.float32, torch.float16]
SHAPES = [(7, 64), (33, 1000), (64, 4096), (40, 32064), (24, 131072), (16, 151936), (5, 50257), (300, 8192)]


def _make(m, n, dt, seed, ignore_frac=0.2, ignore_index=-100):
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(m, n, generator=g) * 3.0).to(dt)
    t = torch.randint(0, n, (m,), generator=g)
    drop = torch.rand(m, generator=g) < ignore_frac
    t[drop] = ignore_index
    if m > 2:
        t[0], t[1] = 0, n - 1  # the first and the last class
    return x, t


@pytest.mark.parametrize("dt", DTYPES, ids=["bf16", "f32", "f16"])
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: f"{s[0]}x{s[1]}")
def test_nll_softmax_fwd_bwd(shape, dt):
    m, n = shape
    x, t = _make(m, n, dt, m + n)
    loss, lse = ops.nll_softmax_fwd(x.cuda(), t.cuda())
    torch.cuda.synchronize()
    assert loss.dtype == torch.float32 and loss.shape == (m,) and lse.shape == (m,)
    le, lsee = AO.nll_softmax_fwd(x, t)
    AO.assert_close(lse, lsee, n, "log-sum-exp")
    AO.assert_close(loss, le, n, "per-row loss")
    assert torch.equal(loss.cpu()[t == -100], torch.zeros_like(le[t == -100]))
    for reduction in ("mean", "sum"):
        total, lse2, count = ops.cross_entropy_fwd(x.cuda(), t.cuda(), reduction=reduction)
        dxe, ref = AO.cross_entropy_bwd(x, t, reduction=reduction)
        AO.assert_close(total.reshape(1), ref.reshape(1), n * m, f"cross_entropy {reduction}")
        dx = ops.cross_entropy_bwd(x.cuda(), t.cuda(), lse2, count if reduction == "mean" else None)
        assert dx.dtype == dt and dx.shape == (m, n)
        AO.assert_close(dx, dxe, n, f"d logits ({reduction})")
        assert not dx.cpu()[t == -100].any(), "ignored rows get no gradient"
    for variant in (L.VARIANT_GENERIC,):
        loss_g, lse_g = ops.nll_softmax_fwd(x.cuda(), t.cuda(), variant=variant)
        AO.assert_close(loss_g, le, n, "per-row loss, shape-agnostic family")
        AO.assert_close(lse_g, lsee, n, "lse, shape-agnostic family")


@pytest.mark.parametrize("cluster", [1, 2, 4, 8, 16])
def test_nll_softmax_cluster_split(cluster):
    """The target's logit lives in exactly one CTA of the cluster: every owner, including the edges of the slices."""
    dt = torch.bfloat16
    m, n = 64, 32768
    x, t = _make(m, n, dt, cluster, ignore_frac=0.0)
    step = n // 16
    for i in range(16):  # first / last column of each sixteenth of the row
        t[2 * i], t[2 * i + 1] = i * step, (i + 1) * step - 1
    try:
        ops.plan_for(L.OP_REDUCE_INNER, dt, m, n, flags=L.F_REDUCE_NLL, cluster_size=cluster)
    except L.UnsupportedSegment:
        pytest.skip(f"no plan with a cluster of {cluster}")
    loss, lse = ops.nll_softmax_fwd(x.cuda(), t.cuda(), cluster_size=cluster)
    le, lsee = AO.nll_softmax_fwd(x, t)
    AO.assert_close(loss, le, n, f"loss C={cluster}")
    AO.assert_close(lse, lsee, n, f"lse C={cluster}")


def test_cross_entropy_semantics_and_views():
    """ignore_index other than -100, out-of-range targets treated as ignored rows of the kernel (torch raises),
    3-D logits [B, S, V] with [B, S] targets, strided rows, -inf logits."""
    dt = torch.bfloat16
    b, s_, v = 4, 37, 5000
    g = torch.Generator().manual_seed(3)
    x = torch.randn(b, s_, v, generator=g).to(dt)
    t = torch.randint(0, v, (b, s_), generator=g)
    t[0, :5] = 7
    loss, lse = ops.nll_softmax_fwd(x.cuda(), t.cuda(), ignore_index=7)
    le, lsee = AO.nll_softmax_fwd(x, t, ignore_index=7)
    assert loss.shape == (b, s_)
    AO.assert_close(loss, le, v, "3-D logits, ignore_index = 7")
    total, lse2, count = ops.cross_entropy_fwd(x.cuda(), t.cuda(), ignore_index=7)
    dx = ops.cross_entropy_bwd(x.cuda(), t.cuda(), lse2, count, ignore_index=7)
    dxe, ref = AO.cross_entropy_bwd(x, t, ignore_index=7)
    AO.assert_close(total.reshape(1), ref.reshape(1), v * b * s_, "mean loss")
    AO.assert_close(dx, dxe, v, "d logits 3-D")
    # strided rows (a view of a wider buffer) and masked (-inf) logits
    wide = torch.randn(50, 4096 + 64, generator=g).to(dt)
    wide[:, 100:200] = float("-inf")
    xv = wide[:, :4096]
    tv = torch.randint(200, 4096, (50,), generator=g)
    loss, lse = ops.nll_softmax_fwd(wide.cuda()[:, :4096], tv.cuda())
    le, lsee = AO.nll_softmax_fwd(xv.contiguous(), tv)
    AO.assert_close(loss, le, 4096, "strided rows with -inf logits")
    dx = ops.cross_entropy_bwd(wide.cuda()[:, :4096], tv.cuda(), lse, torch.tensor(50, device="cuda"))
    dxe, _ = AO.cross_entropy_bwd(xv.contiguous(), tv)
    AO.assert_close(dx, dxe, 4096, "d logits, strided rows")
    assert not dx[:, 100:200].any()


def test_cross_entropy_descriptors_and_plans():
    p = ops.plan_for(L.OP_REDUCE_INNER, torch.bfloat16, 8192, 131072, flags=L.F_REDUCE_NLL)
    assert p["variant"] == "cluster" and p["algorithmic_bytes"] == 2 * 8192 * 131072 + 8 * 8192 + 8 * 8192
    p = ops.plan_for(L.OP_NLL_SOFTMAX_BWD, torch.bfloat16, 8192, 131072)
    assert p["block"] == 256 and p["algorithmic_bytes"] == 2 * 2 * 8192 * 131072 + 12 * 8192 + 4
    with pytest.raises(L.B200NormError):
        ops.plan_for(L.OP_REDUCE_INNER, torch.bfloat16, 64, 4096, flags=L.F_REDUCE_NLL | L.F_REDUCE_MAX)
    with pytest.raises(L.UnsupportedSegment):
        ops.plan_for(L.OP_SOFTMAX_FWD, torch.bfloat16, 64, 4096, flags=L.F_REDUCE_NLL)


@pytest.mark.parametrize("m,n,dt,c", [(700, 131072, torch.bfloat16, 0), (500, 151936, torch.bfloat16, 0),
                                      (600, 65536, torch.bfloat16, 2), (400, 131072, torch.bfloat16, 16),
                                      (300, 131072, torch.float32, 0), (500, 98304, torch.float16, 2),
                                      (500, 100352, torch.bfloat16, 0), (300, 100352, torch.float32, 0),
                                      (1500, 32064, torch.bfloat16, 1), (900, 16384, torch.float32, 1)],
                         ids=lambda v: str(v).replace("torch.", ""))
def test_nll_softmax_pipe_many_rows(m, n, dt, c):
    """The pipelined one-exchange cluster kernel (kernels_softmax_pipe.cuh, cross-entropy flavour) over many rows per
    cluster: ring refills, phase wraps, the target index fetched one row ahead; slices of -inf and a NaN row (torch.logsumexp special-cases +inf, the op-by-op log-softmax of the reference does not: not compared)."""
    x, t = _make(m, n, dt, m + n + c)
    x[3, : n // 16] = float("-inf")
    x[5, : n - 1] = float("-inf")
    t[5] = n - 1
    x[9, n - 3] = float("nan")
    plan = ops.plan_for(L.OP_REDUCE_INNER, dt, m, n, flags=L.F_REDUCE_NLL, cluster_size=c)
    assert plan["variant"] == ("tma" if c == 1 else "cluster") and plan["pipe"], plan
    assert m > 2 * plan["stages"] * (plan["grid"] // plan["cluster_size"])
    loss, lse = ops.nll_softmax_fwd(x.cuda(), t.cuda(), cluster_size=c)
    torch.cuda.synchronize()
    le, lsee = AO.nll_softmax_fwd(x, t)
    AO.assert_close(lse, lsee, n, "log-sum-exp")
    AO.assert_close(loss, le, n, "per-row loss")
    assert torch.equal(torch.isnan(lse.cpu()), torch.isnan(lsee)) and torch.equal(torch.isinf(lse.cpu()), torch.isinf(lsee))
    loss2, lse2 = ops.nll_softmax_fwd(x.cuda(), t.cuda(), cluster_size=c)
    assert torch.equal(lse2.view(torch.int32), lse.view(torch.int32)), "same bits from run to run"

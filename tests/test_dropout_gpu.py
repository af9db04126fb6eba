"""GPU parity of the fused-dropout segments — norm(residual + dropout(x)) forward and backward, LayerNorm and RMSNorm
(benchmarks/python/test_dropout_{layernorm,rmsnorm}_{fwd,bwd}.py) — through the C ABI. The mask the kernels draw is
compared BIT FOR BIT with the oracle's Philox stream (pinned to the Random123 known answers in tests/test_dropout.py);
the normalized values are compared with the ATen oracle evaluated on that mask under the validator's tolerances."""
import pytest
import torch

from fuser_b200 import _lib as L
from fuser_b200 import ops
from oracle import aten_oracle as AO
from tests.util import make_inputs

pytestmark = pytest.mark.gpu

DTYPES = [torch.bfloat16, torch.float32, torch.float16]
SHAPES = [(5, 64), (33, 1000), (64, 250), (216, 768), (300, 4096), (1500, 8192)]


def _inputs(m, n, dt, seed):
    x, dy, w, b = make_inputs(m, n, dt, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    res = torch.randn(m, n, generator=g).to(dt)
    return x, res, dy, w, b


@pytest.mark.parametrize("dt", DTYPES, ids=["bf16", "f32", "f16"])
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: f"{s[0]}x{s[1]}")
@pytest.mark.parametrize("p", [0.2, 0.5])
def test_dropout_layer_norm_fwd_bwd(shape, dt, p):
    """Default dispatch: the register-resident ring kernels for aligned rows (forward), the shape-agnostic family
    otherwise (odd N, and the backward)."""
    m, n = shape
    x, res, dy, w, b = _inputs(m, n, dt, m + n)
    seed, offset = 0x1234_5678_9ABC_DEF0 + m, 4 * n + 3
    rng = ops.philox_state(seed, offset, "cuda")
    y, mean, invstd, mask = ops.dropout_layer_norm_fwd(x.cuda(), res.cuda(), w.cuda(), b.cuda(), 1e-5, p, rng)
    torch.cuda.synchronize()
    assert mask.dtype == torch.bool and mask.shape == (m, n) and mean.shape == (m, 1)
    mref = AO.dropout_mask(seed, offset, (m, n), 1.0 - p)
    assert torch.equal(mask.cpu(), mref), "mask differs from the reference's Philox stream"
    ye, me, ie = AO.dropout_layer_norm_fwd(x, res, w, b, 1e-5, mref, 1.0 - p)
    AO.assert_close(y, ye, n, "dropout LN y")
    AO.assert_close(mean, me, n, "dropout LN mean")
    AO.assert_close(invstd, ie, n, "dropout LN invstd")
    # backward on the oracle's statistics
    dx, dres, dw, db = ops.dropout_layer_norm_bwd(dy.cuda(), x.cuda(), res.cuda(), mask, me.cuda(), ie.cuda(), w.cuda(), p)
    dxe, drese, dwe, dbe = AO.dropout_norm_bwd("ln", dy, x, res, mref, me, ie, w, 1.0 - p)
    AO.assert_close(dres, drese, n, "dropout LN dresidual")
    AO.assert_close(dx, dxe, n, "dropout LN dx")
    AO.assert_close(dw, dwe, m, "dropout LN dweight")
    AO.assert_close(db, dbe, m, "dropout LN dbias")
    assert torch.equal(dx.cpu()[~mref], torch.zeros_like(dx.cpu()[~mref]))  # dropped elements get no gradient
    # output masks
    dx2, dres2, dw2, db2 = ops.dropout_layer_norm_bwd(dy.cuda(), x.cuda(), res.cuda(), mask, me.cuda(), ie.cuda(), w.cuda(), p,
                                                      (True, False, False))
    assert dw2 is None and db2 is None and torch.equal(dx2, dx) and torch.equal(dres2, dres)
    dx3, dres3, dw3, db3 = ops.dropout_layer_norm_bwd(dy.cuda(), x.cuda(), res.cuda(), mask, me.cuda(), ie.cuda(), w.cuda(), p,
                                                      (False, True, True))
    assert dx3 is None and dres3 is None and torch.equal(dw3, dw) and torch.equal(db3, db)


@pytest.mark.parametrize("dt", DTYPES, ids=["bf16", "f32", "f16"])
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: f"{s[0]}x{s[1]}")
@pytest.mark.parametrize("stat_is_rms", [True, False], ids=["rms", "invstd"])
def test_dropout_rms_norm_fwd_bwd(shape, dt, stat_is_rms):
    m, n = shape
    p = 0.2
    x, res, dy, w, _ = _inputs(m, n, dt, 3 * m + n)
    seed, offset = 77 + n, m
    rng = ops.philox_state(seed, offset, "cuda")
    y, stat, mask = ops.dropout_rms_norm_fwd(x.cuda(), res.cuda(), w.cuda(), 1e-5, p, rng, stat_is_rms=stat_is_rms)
    torch.cuda.synchronize()
    mref = AO.dropout_mask(seed, offset, (m, n), 1.0 - p)
    assert torch.equal(mask.cpu(), mref)
    ye, se = AO.dropout_rms_norm_fwd(x, res, w, 1e-5, mref, 1.0 - p, python=stat_is_rms)
    AO.assert_close(y, ye, n, "dropout RMS y")
    AO.assert_close(stat, se, n, "dropout RMS stat")
    dx, dres, dw = ops.dropout_rms_norm_bwd(dy.cuda(), x.cuda(), res.cuda(), mask, se.cuda(), w.cuda(), p, stat_is_rms=stat_is_rms)
    dxe, drese, dwe, _ = AO.dropout_norm_bwd("rms", dy, x, res, mref, None, se, w, 1.0 - p, stat_is_rms=stat_is_rms)
    AO.assert_close(dres, drese, n, "dropout RMS dresidual")
    AO.assert_close(dx, dxe, n, "dropout RMS dx")
    AO.assert_close(dw, dwe, m, "dropout RMS dweight")


def test_dropout_p0_equals_the_fused_add_and_streams_are_reproducible():
    dt = torch.bfloat16
    m, n = 257, 2048
    x, res, dy, w, b = _inputs(m, n, dt, 5)
    rng = ops.philox_state(1, 0, "cuda")
    y0, mean0, invstd0, mask0 = ops.dropout_layer_norm_fwd(x.cuda(), res.cuda(), w.cuda(), b.cuda(), 1e-5, 0.0, rng)
    ya, meana, invstda = ops.layer_norm_fwd(x.cuda(), w.cuda(), b.cuda(), 1e-5, residual=res.cuda())
    assert mask0.all()
    AO.assert_close(y0, ya.cpu(), n, "p = 0 against the fused add")
    assert torch.allclose(mean0, meana, atol=1e-6) and torch.allclose(invstd0, invstda, rtol=1e-5)
    # same {seed, offset}: the same mask; the device tensor is read at run time (no re-compilation for a new offset)
    outs = []
    for off in (0, 0, 4, 8):
        rng.copy_(ops.philox_state(99, off, "cuda"))
        outs.append(ops.dropout_layer_norm_fwd(x.cuda(), res.cuda(), w.cuda(), b.cuda(), 1e-5, 0.3, rng)[3].clone())
    assert torch.equal(outs[0], outs[1]) and not torch.equal(outs[0], outs[2]) and not torch.equal(outs[2], outs[3])
    assert abs(outs[0].float().mean().item() - 0.7) < 0.01
    # want_sum returns residual + dropout(x) in the io dtype
    y, mean, invstd, mask, s = ops.dropout_layer_norm_fwd(x.cuda(), res.cuda(), w.cuda(), b.cuda(), 1e-5, 0.3, rng, want_sum=True)
    se = AO.dropout_sum(x, res, mask.cpu(), 0.7).to(dt)
    AO.assert_close(s, se, 1, "residual + dropout(x)")


def test_dropout_full_size_c2_shape():
    """bf16 [16384, 4096] (the C2 shape): mask identical to the oracle stream over 67 M elements, forward checked on
    a sample of rows, backward through the column sums."""
    dt = torch.bfloat16
    m, n, p = 16384, 4096, 0.2
    x, res, dy, w, b = _inputs(m, n, dt, 11)
    seed, offset = 2026, 1 << 33  # an offset beyond 32 bits exercises counter word 1
    rng = ops.philox_state(seed, offset, "cuda")
    y, mean, invstd, mask = ops.dropout_layer_norm_fwd(x.cuda(), res.cuda(), w.cuda(), b.cuda(), 1e-5, p, rng)
    mref = AO.dropout_mask(seed, offset, (m, n), 1.0 - p)
    assert torch.equal(mask.cpu(), mref)
    rows = torch.arange(0, m, 37)
    ye, me, ie = AO.dropout_layer_norm_fwd(x[rows], res[rows], w, b, 1e-5, mref[rows], 1.0 - p)
    AO.assert_close(y.cpu()[rows], ye, n, "y (sampled rows)")
    AO.assert_close(mean.cpu()[rows], me, n, "mean (sampled rows)")
    dx, dres, dw, db = ops.dropout_layer_norm_bwd(dy.cuda(), x.cuda(), res.cuda(), mask, mean, invstd, w.cuda(), p)
    dxe, drese, dwe, dbe = AO.dropout_norm_bwd("ln", dy, x, res, mref, mean.cpu(), invstd.cpu(), w, 1.0 - p)
    AO.assert_close(dres.cpu()[rows], drese[rows], n, "dresidual (sampled rows)")
    AO.assert_close(dx.cpu()[rows], dxe[rows], n, "dx (sampled rows)")
    AO.assert_close(dw, dwe, m, "dweight")
    AO.assert_close(db, dbe, m, "dbias")


@pytest.mark.parametrize("dt", DTYPES, ids=["bf16", "f32", "f16"])
def test_dropout_forward_families_agree(dt):
    """The same {seed, offset} through the ring kernels and through the shape-agnostic family: identical masks,
    values within tolerance; cluster plans (a row split over CTAs) index the stream by the global column."""
    m, n, p = 96, 8192, 0.25
    x, res, dy, w, b = _inputs(m, n, dt, 9)
    rng = ops.philox_state(5, 12, "cuda")
    mref = AO.dropout_mask(5, 12, (m, n), 1.0 - p)
    ye, me, ie = AO.dropout_layer_norm_fwd(x, res, w, b, 1e-5, mref, 1.0 - p)
    re_, se = AO.dropout_rms_norm_fwd(x, res, w, 1e-5, mref, 1.0 - p, python=True)
    for kw in ({}, {"variant": L.VARIANT_GENERIC}, {"cluster_size": 2}, {"cluster_size": 4}):
        y, mean, invstd, mask = ops.dropout_layer_norm_fwd(x.cuda(), res.cuda(), w.cuda(), b.cuda(), 1e-5, p, rng, **kw)
        assert torch.equal(mask.cpu(), mref), kw
        AO.assert_close(y, ye, n, f"LN y {kw}")
        AO.assert_close(invstd, ie, n, f"LN invstd {kw}")
        yr, stat, maskr = ops.dropout_rms_norm_fwd(x.cuda(), res.cuda(), w.cuda(), 1e-5, p, rng, **kw)
        assert torch.equal(maskr.cpu(), mref), kw
        AO.assert_close(yr, re_, n, f"RMS y {kw}")
        AO.assert_close(stat, se, n, f"RMS stat {kw}")
    plan = ops.plan_for(L.OP_LAYERNORM_FWD, dt, m, n, dropout_keep=1.0 - p,
                        flags=L.F_HAS_WEIGHT | L.F_HAS_BIAS | L.F_OUT_MEAN | L.F_OUT_INVSTD | L.F_ADD_RESIDUAL | L.F_DROPOUT)
    assert plan["variant"] in ("tma", "cluster") and plan["n_launches"] == 1, plan
    # backward: the cached ring form with the residual and mask streams, single CTA per row and over a cluster
    _, _, _, mask = ops.dropout_layer_norm_fwd(x.cuda(), res.cuda(), w.cuda(), b.cuda(), 1e-5, p, rng)
    dxe, drese, dwe, dbe = AO.dropout_norm_bwd("ln", dy, x, res, mref, me, ie, w, 1.0 - p)
    for kw in ({}, {"variant": L.VARIANT_GENERIC}, {"cluster_size": 2}):
        dx, dres, dw, db = ops.dropout_layer_norm_bwd(dy.cuda(), x.cuda(), res.cuda(), mask, me.cuda(), ie.cuda(), w.cuda(), p, **kw)
        AO.assert_close(dres, drese, n, f"dresidual {kw}")
        AO.assert_close(dx, dxe, n, f"dx {kw}")
        AO.assert_close(dw, dwe, m, f"dweight {kw}")
        AO.assert_close(db, dbe, m, f"dbias {kw}")
    bplan = ops.plan_for(L.OP_LAYERNORM_BWD, dt, m, n, dropout_keep=1.0 - p,
                         flags=L.F_HAS_WEIGHT | L.F_OUT_DX | L.F_OUT_DW | L.F_OUT_DB | L.F_DROPOUT)
    assert bplan["variant"] in ("tma", "cluster"), bplan

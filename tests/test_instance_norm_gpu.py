"""GPU parity of instance norm (training statistics; csrc/ops/normalization.cpp:790-950 instance_norm, :952-1060
instance_norm_backward; benchmarks/python/test_instancenorm_{fwd,bwd}.py) through the C ABI. Channels first is one
launch (a LayerNorm over the [B*C, spatial] view with the affine pair of the row's channel, B200NORM_F_ROW_AFFINE);
channels last runs one outer normalization per sample. The oracle is pinned to torch.nn.functional.instance_norm and
to autograd."""
import pytest
import torch

from fuser_b200 import _lib as L
from fuser_b200 import ops
from oracle import aten_oracle as AO

pytestmark = pytest.mark.gpu

DTYPES = [torch.bfloat16, torch.float32, torch.float16]
# (B, C, spatial...)
SHAPES = [(2, 3, 5, 7), (4, 16, 28, 28), (8, 64, 32, 32), (3, 5, 1000), (2, 128, 56, 56), (16, 32, 8, 8, 8), (2, 4, 96, 128),
          (1, 7, 8200)]


def _make(shape, dt, seed, channels_last=False):
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(*shape, generator=g) * 2.0 + 0.5).to(dt)
    dy = torch.randn(*shape, generator=g).to(dt)
    c = shape[1]
    w, b = torch.randn(c, generator=g).to(dt), torch.randn(c, generator=g).to(dt)
    if channels_last:
        perm = [0] + list(range(2, len(shape))) + [1]
        x, dy = x.permute(perm).contiguous(), dy.permute(perm).contiguous()
    return x, dy, w, b


@pytest.mark.parametrize("dt", DTYPES, ids=["bf16", "f32", "f16"])
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_instance_norm_channels_first(shape, dt):
    x, dy, w, b = _make(shape, dt, sum(shape))
    bsz, c = shape[0], shape[1]
    n = x[0, 0].numel()
    for variant in (L.VARIANT_AUTO, L.VARIANT_GENERIC):
        for ww, bb in ((w, b), (None, None)):
            y, mean, invstd = ops.instance_norm_fwd(x.cuda(), None if ww is None else ww.cuda(), None if bb is None else bb.cuda(),
                                                    1e-5, variant=variant)
            torch.cuda.synchronize()
            assert y.shape == x.shape and mean.shape == (bsz, c) and invstd.shape == (bsz, c) and mean.dtype == torch.float32
            ye, me, ie = AO.instance_norm_fwd(x, ww, bb, 1e-5)
            AO.assert_close(y, ye, n, f"instance norm y (variant {variant})")
            AO.assert_close(mean, me, n, "mean")
            AO.assert_close(invstd, ie, n, "invstd")
        dx, dw, db = ops.instance_norm_bwd(dy.cuda(), x.cuda(), me.cuda(), ie.cuda(), w.cuda(), variant=variant)
        dxe, dwe, dbe = AO.instance_norm_bwd(dy, x, me, ie, w)
        AO.assert_close(dx, dxe, n, f"dx (variant {variant})")
        AO.assert_close(dw, dwe, n * bsz, "dweight")
        AO.assert_close(db, dbe, n * bsz, "dbias")
        dx2, dw2, db2 = ops.instance_norm_bwd(dy.cuda(), x.cuda(), me.cuda(), ie.cuda(), w.cuda(), (True, False, False), variant=variant)
        assert dw2 is None and db2 is None and torch.equal(dx2, dx)
        dx3, dw3, db3 = ops.instance_norm_bwd(dy.cuda(), x.cuda(), me.cuda(), ie.cuda(), w.cuda(), (False, True, True), variant=variant)
        assert dx3 is None and torch.equal(dw3, dw) and torch.equal(db3, db)


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float32], ids=["bf16", "f32"])
@pytest.mark.parametrize("shape", [(2, 16, 12, 12), (3, 64, 28, 28), (2, 8, 40, 50)], ids=lambda s: "x".join(map(str, s)))
def test_instance_norm_channels_last(shape, dt):
    x, dy, w, b = _make(shape, dt, sum(shape) + 1, channels_last=True)
    bsz, c = shape[0], shape[1]
    n = x[0, ..., 0].numel()
    y, mean, invstd = ops.instance_norm_fwd(x.cuda(), w.cuda(), b.cuda(), 1e-5, channels_last=True)
    ye, me, ie = AO.instance_norm_fwd(x, w, b, 1e-5, channels_last=True)
    assert y.shape == x.shape and mean.shape == (bsz, c)
    AO.assert_close(y, ye, n, "channels-last y")
    AO.assert_close(mean, me, n, "channels-last mean")
    AO.assert_close(invstd, ie, n, "channels-last invstd")
    dx, dw, db = ops.instance_norm_bwd(dy.cuda(), x.cuda(), me.cuda(), ie.cuda(), w.cuda(), channels_last=True)
    dxe, dwe, dbe = AO.instance_norm_bwd(dy, x, me, ie, w, channels_last=True)
    AO.assert_close(dx, dxe, n, "channels-last dx")
    AO.assert_close(dw, dwe, n * bsz, "channels-last dweight")
    AO.assert_close(db, dbe, n * bsz, "channels-last dbias")

"""CPU checks of the instance-norm support: the oracle pinned to torch.nn.functional.instance_norm and autograd, and
the descriptor rules of B200NORM_F_ROW_AFFINE."""
import pytest
import torch

from fuser_b200 import _lib as L
from fuser_b200 import ops
from oracle import aten_oracle as AO
from tests.test_instance_norm_gpu import _make


def test_oracle_is_pinned_to_aten():
    x, dy, w, b = _make((4, 6, 9, 11), torch.float32, 1)
    ye, me, ie = AO.instance_norm_fwd(x, w, b, 1e-5)
    yn = torch.nn.functional.instance_norm(x, weight=w, bias=b, eps=1e-5)
    assert torch.allclose(ye, yn, atol=1e-5, rtol=1e-5)
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    torch.nn.functional.instance_norm(xr, weight=wr, bias=br, eps=1e-5).backward(dy)
    dxe, dwe, dbe = AO.instance_norm_bwd(dy, x, me, ie, w)
    assert torch.allclose(dxe, xr.grad, atol=1e-4, rtol=1e-4) and torch.allclose(dwe, wr.grad, atol=1e-3, rtol=1e-4)
    assert torch.allclose(dbe, br.grad, atol=1e-3, rtol=1e-4)


def test_instance_norm_plans_and_descriptors():
    f = L.F_HAS_WEIGHT | L.F_HAS_BIAS | L.F_OUT_MEAN | L.F_OUT_INVSTD | L.F_ROW_AFFINE
    p = ops.plan_for(L.OP_LAYERNORM_FWD, torch.bfloat16, 256 * 128, 784, flags=f, ignore_index=128)  # InstanceNorm [256,128,28,28]
    assert p["variant"] == "tma" and p["n_launches"] == 1, p
    with pytest.raises(L.B200NormError):  # rows must be whole samples
        ops.plan_for(L.OP_LAYERNORM_FWD, torch.bfloat16, 1000, 784, flags=f, ignore_index=128)
    with pytest.raises(L.UnsupportedSegment):
        ops.plan_for(L.OP_RMSNORM_FWD, torch.bfloat16, 1024, 784, flags=L.F_ROW_AFFINE, ignore_index=128)
    with pytest.raises(L.UnsupportedSegment):
        ops.plan_for(L.OP_LAYERNORM_FWD, torch.bfloat16, 1024, 784, flags=f | L.F_ADD_RESIDUAL, ignore_index=128)

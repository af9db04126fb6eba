"""The all-reduce of dgamma / dbeta fused into the backward kernel over peer memory
(B200NORM_F_PEER_REDUCE)."""
import os
import subprocess
import sys

import pytest
import torch

from fuser_b200 import _lib as L
from fuser_b200 import ops
from tests.util import make_inputs

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float32], ids=["bf16", "f32"])
@pytest.mark.parametrize("shape", [(1000, 4096), (300, 12288), (64, 65536)], ids=lambda s: f"{s[0]}x{s[1]}")
def test_peer_reduce_world_of_one(shape, dt):
    """One rank: the exchange path (second grid barrier, flag, slot read-back, cast) must reproduce
    the single-GPU gradients bit for bit, launch after launch (the two slots alternate)."""
    m, n = shape
    x, dy, w, _ = make_inputs(m, n, dt, seed=n + m)
    xd, dyd, wd = x.cuda(), dy.cuda(), w.cuda()
    _, mean, invstd = ops.layer_norm_fwd(xd, None, None, 1e-5)
    _, rstd = ops.rms_norm_fwd(xd, wd, 1e-5)
    ref_ln = ops.layer_norm_bwd(dyd, xd, mean, invstd, wd)
    ref_rms = ops.rms_norm_bwd(dyd, xd, rstd, wd)
    nbytes = L.peer_exchange_bytes(1, 2 * n)
    buf = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    L.register_peers(1, 0, [buf.data_ptr()], nbytes)
    try:
        for _ in range(3):
            got = ops.layer_norm_bwd(dyd, xd, mean, invstd, wd, peer_reduce=True)
            assert all(torch.equal(a, b) for a, b in zip(got, ref_ln))
            got = ops.rms_norm_bwd(dyd, xd, rstd, wd, peer_reduce=True)
            assert all(torch.equal(a, b) for a, b in zip(got, ref_rms))
        torch.cuda.synchronize()
        # the receive area (behind the header and the two publish slots) carries (value, epoch) words: the
        # sixth launch wrote its first sum with epoch 6 into parity 0
        ll_off = 256 + 2 * (2 * n) * 4
        assert int(buf[ll_off + 4:ll_off + 8].view(torch.int32).item()) == 6, "six fused launches: epochs 1..6"
        # publish now / combine later, two publishes in flight
        for _ in range(2):
            dx1, none_w, none_b = ops.layer_norm_bwd(dyd, xd, mean, invstd, wd, peer_reduce="publish")
            dx2, none_w2 = ops.rms_norm_bwd(dyd, xd, rstd, wd, peer_reduce="publish")
            assert none_w is None and none_b is None and none_w2 is None
            with pytest.raises(L.B200NormError):  # a third publish needs a combine first
                ops.rms_norm_bwd(dyd, xd, rstd, wd, peer_reduce="publish")
            dw, db = ops.peer_combine(2, n, dt, buf)
            (dw2,) = ops.peer_combine(1, n, dt, buf)
            assert torch.equal(dx1, ref_ln[0]) and torch.equal(dw, ref_ln[1]) and torch.equal(db, ref_ln[2])
            assert torch.equal(dx2, ref_rms[0]) and torch.equal(dw2, ref_rms[1])
        with pytest.raises(L.B200NormError):  # nothing published
            ops.peer_combine(1, n, dt, buf)
    finally:
        L.register_peers(0, 0, [], 0)


def test_peer_reduce_errors():
    x = torch.randn(64, 1024, device="cuda", dtype=torch.bfloat16)
    _, rstd = ops.rms_norm_fwd(x, None, 1e-5)
    w = torch.ones(1024, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(L.B200NormError):  # nothing registered
        ops.rms_norm_bwd(x, x.clone(), rstd, w, peer_reduce=True)
    small = torch.zeros(512, dtype=torch.uint8, device="cuda")
    L.register_peers(1, 0, [small.data_ptr()], 512)
    try:
        with pytest.raises(L.B200NormError):  # slots do not fit
            ops.rms_norm_bwd(x, x.clone(), rstd, w, peer_reduce=True)
        with pytest.raises(L.UnsupportedSegment):  # generic family cannot take the fused path
            ops.rms_norm_bwd(x, x.clone(), rstd, w, peer_reduce=True, variant=L.VARIANT_GENERIC)
    finally:
        L.register_peers(0, 0, [], 0)
    with pytest.raises(L.B200NormError):
        L.register_peers(9, 0, [small.data_ptr()] * 9, 512)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs of one box")
@pytest.mark.parametrize("mode", ["fused", "deferred"])
def test_peer_two_gpus(mode):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "peer_worker.py"), mode]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0 and "peer worker ok" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]

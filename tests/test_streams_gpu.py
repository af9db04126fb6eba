"""Concurrency of the launch path: kernels that finish dgamma / dbeta behind an in-kernel grid barrier are
launched on several streams with separate workspaces (include/b200norm.h allows that) and must neither
dead-lock (two partially resident grids spinning on their barriers) nor change a bit; one process driving two
devices must set up function attributes / occupancy answers per device. Reference: cooperative launches for
grid reductions, csrc/runtime/executor.cpp:1307-1340."""
import pytest
import torch

from fuser_b200 import ops
from tests.util import make_inputs

pytestmark = pytest.mark.gpu


def test_barrier_kernels_on_two_streams():
    dt = torch.bfloat16
    shapes = [(4096, 4096), (1200, 12288)]
    data = []
    for m, n in shapes:
        x, dy, w, _ = make_inputs(m, n, dt, seed=m + n)
        xd, dyd, wd = x.cuda(), dy.cuda(), w.cuda()
        _, mean, invstd = ops.layer_norm_fwd(xd, None, None, 1e-5)
        _, rstd = ops.rms_norm_fwd(xd, wd, 1e-5)
        ref_ln = ops.layer_norm_bwd(dyd, xd, mean, invstd, wd)
        ref_rms = ops.rms_norm_bwd(dyd, xd, rstd, wd)
        data.append((xd, dyd, wd, mean, invstd, rstd, ref_ln, ref_rms))
    # something that keeps SMs busy while the barrier kernels of both streams become ready
    filler_in = torch.randn(8192, 8192, device="cuda", dtype=dt)
    torch.cuda.synchronize()
    s1, s2, s3 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    results = []
    for it in range(50):
        with torch.cuda.stream(s3):
            ops.softmax_fwd(filler_in)
        for s, d in ((s1, data[it % 2]), (s2, data[(it + 1) % 2])):
            xd, dyd, wd, mean, invstd, rstd, _, _ = d
            with torch.cuda.stream(s):
                a = ops.layer_norm_bwd(dyd, xd, mean, invstd, wd)
                b = ops.rms_norm_bwd(dyd, xd, rstd, wd)
            results.append((d, a, b))
    torch.cuda.synchronize()  # a dead-lock would trap after the watchdog's 4 s and surface here
    for d, a, b in results:
        assert all(torch.equal(p, q) for p, q in zip(a, d[6])), "LayerNorm backward differs from the serial run"
        assert all(torch.equal(p, q) for p, q in zip(b, d[7])), "RMSNorm backward differs from the serial run"


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_one_process_two_devices():
    """cuda:0 then cuda:1 in one process: > 48 KB of dynamic shared memory needs the function attribute on
    EACH device, and the cluster occupancy answer is the device's own."""
    outs = []
    for dev in ("cuda:0", "cuda:1", "cuda:0"):
        x, dy, w, _ = make_inputs(700, 65536, torch.bfloat16, seed=5)
        xd, dyd, wd = x.to(dev), dy.to(dev), w.to(dev)
        y, rstd = ops.rms_norm_fwd(xd, wd, 1e-5)
        dx, dw = ops.rms_norm_bwd(dyd, xd, rstd, wd)
        sm = ops.softmax_fwd(xd[:, :8192].contiguous())
        torch.cuda.synchronize(dev)
        outs.append([t.cpu() for t in (y, rstd, dx, dw, sm)])
    for o in outs[1:]:
        for i, (a, b) in enumerate(zip(o, outs[0])):
            if i == 3:  # dgamma: the number of per-cluster partial sums is the device's own cluster occupancy
                assert torch.allclose(a.float(), b.float(), rtol=2e-2, atol=0.5)
            else:
                assert torch.equal(a, b)

"""Small-shape drive of the round-2 kernels for compute-sanitizer (scripts/gpu_sanitize_r02.sh): the row-block outer
kernel (ring wrap + resident tail, forced onto small problems), fused dropout forward / backward on the ring kernels
and over a cluster, cross-entropy forward (cluster) / backward, instance norm — each checked against the oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from fuser_b200 import _lib as L  # noqa: E402
from fuser_b200 import ops  # noqa: E402
from oracle import aten_oracle as AO  # noqa: E402


def main():
    dt = torch.bfloat16
    g = torch.Generator().manual_seed(0)
    # row-block outer kernel: 3000 rows x 64 / 2500 x 200 (idle threads), two-slot ring and the default depth
    for m, n, dtt in ((3000, 64, dt), (2500, 200, torch.float32), (2100, 1024, dt)):
        x = (torch.randn(m, n, generator=g) + 0.5).to(dtt)
        dy = torch.randn(m, n, generator=g).to(dtt)
        w, b = torch.randn(n, generator=g).to(dtt), torch.randn(n, generator=g).to(dtt)
        ye, me, ie = AO.outer_norm_fwd(x, w, b, 1e-5)
        dxe, dwe, dbe = AO.outer_norm_bwd(dy, x, me, ie, w)
        for depth in ("2", None):
            os.environ["B200NORM_OUTER_ROWS"] = "1"
            if depth:
                os.environ["B200NORM_OUTER_ROWS_DEPTH"] = depth
            n0 = L.launch_count()
            y, mean, invstd = ops.outer_norm_fwd(x.cuda(), w.cuda(), b.cuda(), 1e-5)
            dx, dw, db = ops.outer_norm_bwd(dy.cuda(), x.cuda(), me.cuda(), ie.cuda(), w.cuda())
            assert L.launch_count() - n0 == 2, "row-block kernel not taken"
            os.environ.pop("B200NORM_OUTER_ROWS_DEPTH", None)
            os.environ.pop("B200NORM_OUTER_ROWS", None)
            AO.assert_close(y, ye, m, "rows y")
            AO.assert_close(dx, dxe, m, "rows dx")
            AO.assert_close(dw, dwe, m, "rows dw")
    # fused dropout: ring kernels, cluster of 2, shape-agnostic family; forward + backward
    m, n, p = 70, 4096, 0.3
    x, res, dy = (torch.randn(m, n, generator=g).to(dt) for _ in range(3))
    w, b = torch.randn(n, generator=g).to(dt), torch.randn(n, generator=g).to(dt)
    rng = ops.philox_state(11, 5, "cuda")
    mref = AO.dropout_mask(11, 5, (m, n), 1 - p)
    ye, me, ie = AO.dropout_layer_norm_fwd(x, res, w, b, 1e-5, mref, 1 - p)
    dxe, drese, dwe, dbe = AO.dropout_norm_bwd("ln", dy, x, res, mref, me, ie, w, 1 - p)
    for kw in ({}, {"cluster_size": 2}, {"variant": L.VARIANT_GENERIC}):
        y, mean, invstd, mask = ops.dropout_layer_norm_fwd(x.cuda(), res.cuda(), w.cuda(), b.cuda(), 1e-5, p, rng, **kw)
        assert torch.equal(mask.cpu(), mref)
        AO.assert_close(y, ye, n, f"dropout y {kw}")
        dx, dres, dw, db = ops.dropout_layer_norm_bwd(dy.cuda(), x.cuda(), res.cuda(), mask, me.cuda(), ie.cuda(), w.cuda(), p, **kw)
        AO.assert_close(dx, dxe, n, f"dropout dx {kw}")
        AO.assert_close(dw, dwe, m, f"dropout dw {kw}")
        yr, stat, mk = ops.dropout_rms_norm_fwd(x.cuda(), res.cuda(), w.cuda(), 1e-5, p, rng, **kw)
        assert torch.equal(mk.cpu(), mref)
    # cross-entropy: cluster split and single CTA, backward vector and scalar forms
    for m, n, kw in ((40, 32768, {"cluster_size": 4}), (50, 4096, {}), (33, 1001, {})):
        x = (torch.randn(m, n, generator=g) * 3).to(dt)
        t = torch.randint(0, n, (m,), generator=g)
        t[3] = -100
        loss, lse = ops.nll_softmax_fwd(x.cuda(), t.cuda(), **kw)
        le, lsee = AO.nll_softmax_fwd(x, t)
        AO.assert_close(loss, le, n, "ce loss")
        dx = ops.cross_entropy_bwd(x.cuda(), t.cuda(), lse, torch.tensor(m - 1, device="cuda"))
        dxe, _ = AO.cross_entropy_bwd(x, t)
        AO.assert_close(dx, dxe, n, "ce dx")
    # instance norm, channels first (pipe kernel: 16 x 64 x 512; ring kernel: ragged 3 x 5 x 1000) and channels last
    for shape in ((16, 64, 512), (3, 5, 1000), (2, 8, 12, 12)):
        x = torch.randn(*shape, generator=g).to(dt)
        dy = torch.randn(*shape, generator=g).to(dt)
        c = shape[1]
        w, b = torch.randn(c, generator=g).to(dt), torch.randn(c, generator=g).to(dt)
        ye, me, ie = AO.instance_norm_fwd(x, w, b, 1e-5)
        y, mean, invstd = ops.instance_norm_fwd(x.cuda(), w.cuda(), b.cuda(), 1e-5)
        AO.assert_close(y, ye, x[0, 0].numel(), "instance norm y")
        dx, dw, db = ops.instance_norm_bwd(dy.cuda(), x.cuda(), me.cuda(), ie.cuda(), w.cuda())
        dxe, dwe, dbe = AO.instance_norm_bwd(dy, x, me, ie, w)
        AO.assert_close(dx, dxe, x[0, 0].numel(), "instance norm dx")
        AO.assert_close(dw, dwe, x[0, 0].numel() * shape[0], "instance norm dw")
    torch.cuda.synchronize()
    print("sanitize_r02 drive ok")


if __name__ == "__main__":
    main()

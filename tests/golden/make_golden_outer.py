"""Golden fixtures of the OUTER (dim 0) ops, produced by ATen's native kernels — not by oracle/:
at::_softmax / _log_softmax (+ their _backward_data) over dim 0 and at::native_batch_norm /
native_batch_norm_backward in training mode on [M, C] (the reference's batch-norm tests build their
expectations with the same entry points, e.g. tests/cpp/test_gpu1.cpp / test_gpu_outer_reduction.cpp; its
Python benchmarks validate against torch.nn.functional.batch_norm, benchmarks/python/normalization.py:236-262).
They pin oracle/aten_oracle.py (tests/test_oracle.py) and are replayed against the CUDA kernels
(tests/test_outer_gpu.py).

Run:  python tests/golden/make_golden_outer.py      (writes tests/golden/outer_*.npz)
16-bit tensors are stored as raw uint16 bit patterns.
"""
import os

import numpy as np
import torch

from make_golden import DT, HERE, enc


def main():
    torch.manual_seed(0)
    shapes = [(8, 64), (96, 40), (130, 136)]
    for name, dt in DT.items():
        for (m, n) in shapes:
            x = torch.randn(m, n).to(dt)
            dy = torch.randn(m, n).to(dt)
            w = torch.randn(n).to(dt)
            b = torch.randn(n).to(dt)
            eps = 1e-5
            xf, dyf, wf, bf = (t.float() for t in (x, dy, w, b))
            out = dict(x=enc(x), dy=enc(dy), w=enc(w), b=enc(b), eps=np.float64(eps))
            sm = torch.ops.aten._softmax(xf, 0, False).to(dt)
            lsm = torch.ops.aten._log_softmax(xf, 0, False).to(dt)
            sdx = torch.ops.aten._softmax_backward_data(dyf, sm.float(), 0, torch.float32)
            ldx = torch.ops.aten._log_softmax_backward_data(dyf, lsm.float(), 0, torch.float32)
            out.update(osm_y=enc(sm), osm_dx=enc(sdx.to(dt)), olsm_y=enc(lsm), olsm_dx=enc(ldx.to(dt)))
            y, mean, invstd = torch.native_batch_norm(xf, wf, bf, None, None, True, 0.1, eps)
            dx, dw, db = torch.ops.aten.native_batch_norm_backward(dyf, xf, wf, None, None, mean, invstd, True, eps,
                                                                   [True, True, True])
            out.update(bn_y=enc(y.to(dt)), bn_mean=enc(mean), bn_invstd=enc(invstd), bn_dx=enc(dx.to(dt)),
                       bn_dw=enc(dw.to(dt)), bn_db=enc(db.to(dt)))
            np.savez_compressed(os.path.join(HERE, f"outer_{name}_{m}x{n}.npz"), **out)
    print("wrote", len(DT) * len(shapes), "fixtures to", HERE)


if __name__ == "__main__":
    main()

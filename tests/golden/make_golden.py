"""Generate the committed golden fixtures for the hot path.

The reference (nvFuser) stores no golden vectors: its tests compute expectations
with ATen at run time, calling at::native_layer_norm / native_layer_norm_backward
(tests/cpp/test_gpu1.cpp:6751-6760, test_combined_inner_outer_reduction.cpp:113-140),
at::_softmax(_backward_data) and hand-written ATen for RMSNorm
(tests/cpp/test_gpu1.cpp:6806-6830,6897-6915). nvFuser itself cannot be imported
in this container, so these fixtures are produced by those same ATen entry points
(torch 2.11.0 CPU) — the reference's *explicit* expectations — NOT by oracle/.
They pin oracle/aten_oracle.py (tests/test_oracle.py) and are also replayed
against the CUDA kernels on the GPU box (tests/test_parity_gpu.py).

Run:  python tests/golden/make_golden.py      (writes tests/golden/*.npz)
16-bit tensors are stored as raw uint16 bit patterns.
"""
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
DT = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}


def enc(t: torch.Tensor) -> np.ndarray:
    t = t.detach().cpu().contiguous()
    if t.dtype in (torch.bfloat16, torch.float16):
        return t.view(torch.uint16).numpy()
    return t.numpy()


def main():
    torch.manual_seed(0)  # the reference's tests fix the ATen seed to 0 (tests/cpp/utils.cpp:39-46)
    shapes = [(8, 64), (5, 96), (3, 1987), (4, 1024)]
    for name, dt in DT.items():
        for (m, n) in shapes:
            out = {}
            x = torch.randn(m, n).to(dt)
            dy = torch.randn(m, n).to(dt)
            w = torch.randn(n).to(dt)
            b = torch.randn(n).to(dt)
            eps = 1e-5
            xf, dyf, wf, bf = (t.float() for t in (x, dy, w, b))
            # LayerNorm fwd/bwd through the native ATen composite kernels (fp32 math on upcast inputs,
            # exactly how the reference's tests build their expectations for reduced precision)
            y, mean, rstd = torch.native_layer_norm(xf, [n], wf, bf, eps)
            dx, dw, db = torch.ops.aten.native_layer_norm_backward(dyf, xf, [n], mean, rstd, wf, bf, [True, True, True])
            out.update(x=enc(x), dy=enc(dy), w=enc(w), b=enc(b), eps=np.float64(eps))
            out.update(ln_y=enc(y.to(dt)), ln_mean=enc(mean), ln_invstd=enc(rstd), ln_dx=enc(dx.to(dt)),
                       ln_dw=enc(dw.to(dt)), ln_db=enc(db.to(dt)))
            # RMSNorm: hand-written ATen as in tests/cpp/test_gpu1.cpp:6806-6830 (fwd) / :6897-6915 (bwd via autograd)
            xr = xf.clone().requires_grad_(True)
            wr = wf.clone().requires_grad_(True)
            ms = xr.pow(2).mean(-1, keepdim=True)
            invstd = torch.rsqrt(ms + eps)
            yr = xr * invstd * wr
            yr.backward(dyf)
            out.update(rms_y=enc(yr.detach().to(dt)), rms_invstd=enc(invstd.detach()),
                       rms_dx=enc(xr.grad.to(dt)), rms_dw=enc(wr.grad.to(dt)))
            # softmax fwd/bwd
            sm = torch.softmax(xf, dim=-1)
            sm_io = sm.to(dt)
            sdx = torch.ops.aten._softmax_backward_data(dyf, sm_io.float(), -1, torch.float32)
            lsm = torch.log_softmax(xf, dim=-1)
            lsm_io = lsm.to(dt)
            ldx = torch.ops.aten._log_softmax_backward_data(dyf, lsm_io.float(), -1, torch.float32)
            out.update(sm_y=enc(sm_io), sm_dx=enc(sdx.to(dt)), lsm_y=enc(lsm_io), lsm_dx=enc(ldx.to(dt)))
            np.savez_compressed(os.path.join(HERE, f"norm_{name}_{m}x{n}.npz"), **out)
    print("wrote", len(DT) * len(shapes), "fixtures to", HERE)


if __name__ == "__main__":
    main()

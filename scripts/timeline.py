#!/usr/bin/env python
"""Per-CTA timeline of a ring kernel (tuning aid; b200norm_debug_set_timeline).

    python scripts/timeline.py "rms_bwd:16384x4096:bf16[:ENV=V,...]" ...

Launches the op three times back to back (the middle and last launches are instrumented) and prints,
for every mark, min / median / max over the CTAs in microseconds relative to the first CTA's entry,
plus the gap between launch 2's last exit and launch 3's marks (what PDL overlaps).
marks: 0 entry, 1 setup done, 2 first tile landed, 3 row loop done, 4 partial written, 5 grid barrier passed, 6 exit,
8 column sums finished (pushed to the peers), 9 peers' sums collected
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from fuser_b200 import _lib as L  # noqa: E402
from fuser_b200 import ops  # noqa: E402

KEYS = ["B200NORM_TPR", "B200NORM_GROUPS", "B200NORM_DEPTH", "B200NORM_TILE_ROWS", "B200NORM_CLUSTER",
        "B200NORM_BWD_REREAD", "B200NORM_NO_FUSED_FINISH", "B200NORM_NO_PDL", "B200NORM_NO_PSMEM", "B200NORM_COOP", "B200NORM_STATIC", "B200NORM_POOL", "B200NORM_NO_PIPE", "B200NORM_PIPE", "B200NORM_OUTER_ROWS", "B200NORM_OUTER_ROWS_DEPTH"]
DT = {"bf16": torch.bfloat16, "f32": torch.float32, "f16": torch.float16}
NAMES = ["entry", "setup", "first", "loop", "partial", "gbar", "exit", None, "finish", "collect"]
SLOTS = 12


def main():
    dev = torch.device("cuda", 0)
    for spec in sys.argv[1:]:
        parts = spec.split(":")
        op, shape, dt = parts[0], parts[1], parts[2]
        cfg = parts[3] if len(parts) > 3 else ""
        m, n = (int(v) for v in shape.split("x"))
        dtype = DT[dt]
        for k in KEYS:
            os.environ.pop(k, None)
        for kv in filter(None, cfg.split(",")):
            k, v = kv.split("=")
            os.environ[k] = v
        g = torch.Generator(device="cpu").manual_seed(1)
        xs = [torch.randn(m, n, generator=g).to(dtype).to(dev) for _ in range(3)]
        dys = [torch.randn(m, n, generator=g).to(dtype).to(dev) for _ in range(3)]
        w = torch.randn(n, generator=g).to(dtype).to(dev)
        if op.startswith("rms_bwd"):
            st = [ops.rms_norm_fwd(x, w, 1e-5)[1] for x in xs]
            fn = lambda i: ops.rms_norm_bwd(dys[i], xs[i], st[i], w)  # noqa: E731
        elif op.startswith("ln_bwd"):
            st = [ops.layer_norm_fwd(x, None, None, 1e-5)[1:] for x in xs]
            fn = lambda i: ops.layer_norm_bwd(dys[i], xs[i], st[i][0], st[i][1], w)  # noqa: E731
        elif op == "rms_fwd":
            fn = lambda i: ops.rms_norm_fwd(xs[i], w, 1e-5)  # noqa: E731
        elif op == "ln_fwd":
            fn = lambda i: ops.layer_norm_fwd(xs[i], w, w, 1e-5)  # noqa: E731
        elif op == "sm_fwd":
            fn = lambda i: ops.softmax_fwd(xs[i])  # noqa: E731
        elif op == "sm_bwd":
            ys = [ops.softmax_fwd(x) for x in xs]
            fn = lambda i: ops.softmax_bwd(dys[i], ys[i])  # noqa: E731
        elif op == "onorm_fwd":
            fn = lambda i: ops.outer_norm_fwd(xs[i], w, w, 1e-5)  # noqa: E731
        elif op == "onorm_bwd":
            st = [ops.outer_norm_fwd(x, w, w, 1e-5)[1:] for x in xs]
            fn = lambda i: ops.outer_norm_bwd(dys[i], xs[i], st[i][0], st[i][1], w)  # noqa: E731
        else:
            raise SystemExit(f"unknown op {op}")
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        bufs = [torch.zeros(2048 * SLOTS, dtype=torch.int64, device=dev) for _ in range(2)]
        torch.cuda.synchronize()
        fn(0)
        L.set_timeline(bufs[0].data_ptr())
        fn(1)
        L.set_timeline(bufs[1].data_ptr())
        fn(2)
        L.set_timeline(None)
        torch.cuda.synchronize()
        tl = [b.cpu().view(-1, SLOTS) for b in bufs]
        nct = int((tl[1][:, 0] != 0).sum())
        print(f"== {spec}: {nct} CTAs")
        t0 = int(tl[0][:nct, 0].min())
        for li, t in enumerate(tl):
            t = t[:nct]
            print(f" launch {li + 2}:")
            for sidx, name in enumerate(NAMES):
                if name is None:
                    continue
                col = t[:, sidx]
                if int(col.max()) == 0:
                    continue
                v = (col[col != 0] - t0).double() / 1e3
                print(f"   {name:8s} min {v.min():8.2f}  med {v.median():8.2f}  max {v.max():8.2f} us")
        a, b = tl[0][:nct], tl[1][:nct]
        print(f"   launch2 duration (first entry -> last exit) {(int(a[:, 6].max()) - int(a[:, 0].min())) / 1e3:.2f} us;"
              f" launch3 {(int(b[:, 6].max()) - int(b[:, 0].min())) / 1e3:.2f} us;"
              f" period (exit2 -> exit3) {(int(b[:, 6].max()) - int(a[:, 6].max())) / 1e3:.2f} us", flush=True)


if __name__ == "__main__":
    main()

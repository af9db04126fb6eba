#!/usr/bin/env python
"""Kernel-level sweep on one GPU: times each ring kernel alone (back-to-back launches, rotating
input sets larger than L2, CUDA events around the loop) under planner overrides.

    python scripts/sweep.py c2 "" "B200NORM_TPR=256" "B200NORM_CLUSTER=2,B200NORM_DEPTH=3"
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import EPS, WORKLOADS, iobytes, measured_peak  # noqa: E402
from fuser_b200 import _lib as L  # noqa: E402
from fuser_b200 import ops  # noqa: E402

KEYS = ["B200NORM_TPR", "B200NORM_GROUPS", "B200NORM_DEPTH", "B200NORM_TILE_ROWS", "B200NORM_CLUSTER",
        "B200NORM_BWD_REREAD"]


def main():
    w = sys.argv[1]
    configs = sys.argv[2:] or [""]
    desc, m, n, dtype, kind = WORKLOADS[w]
    dev = torch.device("cuda", 0)
    es = torch.empty(0, dtype=dtype).element_size()
    nsets = max(2, int(300e6 // (m * n * es)) + 1)
    g = torch.Generator(device="cpu").manual_seed(1)
    xs = [torch.randn(m, n, generator=g).to(dtype).to(dev) for _ in range(nsets)]
    dys = [torch.randn(m, n, generator=g).to(dtype).to(dev) for _ in range(nsets)]
    wgt = torch.randn(n, generator=g).to(dtype).to(dev)
    byt = iobytes(kind, m, n, es)
    peak, _ = measured_peak()
    iters = 10

    def timeit(fn):
        for i in range(2):
            fn(i)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(iters):
            fn(i)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters * 1e3  # us

    print(f"# {desc}: fwd {byt.get('fwd', 0)/1e6:.1f} MB, bwd {byt.get('bwd', 0)/1e6:.1f} MB, peak {peak} GB/s")
    for cfg in configs:
        for k in KEYS:
            os.environ.pop(k, None)
        for kv in filter(None, cfg.split(",")):
            k, v = kv.split("=")
            os.environ[k] = v
        out = []
        try:
            if kind == "rmsnorm_fwd_bwd":
                ys = [ops.rms_norm_fwd(x, wgt, EPS, stat_is_rms=True) for x in xs]
                t = timeit(lambda i: ops.rms_norm_fwd(xs[i % nsets], wgt, EPS, stat_is_rms=True))
                out.append(("fwd", t, byt["fwd"]))
                t = timeit(lambda i: ops.rms_norm_bwd(dys[i % nsets], xs[i % nsets], ys[i % nsets][1], wgt,
                                                      stat_is_rms=True, centered=False))
                out.append(("bwd+fin", t, byt["bwd"]))
                t = timeit(lambda i: ops.rms_norm_bwd(dys[i % nsets], xs[i % nsets], ys[i % nsets][1], wgt,
                                                      (True, False), stat_is_rms=True, centered=False))
                out.append(("bwd(dx only)", t, byt["bwd"]))
                fl = L.F_HAS_WEIGHT | L.F_OUT_DX | L.F_OUT_DW
                plans = [ops.plan_for(L.OP_RMSNORM_FWD, dtype, m, n, L.F_HAS_WEIGHT),
                         ops.plan_for(L.OP_RMSNORM_BWD, dtype, m, n, fl)]
            elif kind == "softmax_fwd_bwd":
                ys = [ops.softmax_fwd(x) for x in xs]
                t = timeit(lambda i: ops.softmax_fwd(xs[i % nsets]))
                out.append(("fwd", t, byt["fwd"]))
                t = timeit(lambda i: ops.softmax_bwd(dys[i % nsets], ys[i % nsets]))
                out.append(("bwd", t, byt["bwd"]))
                plans = [ops.plan_for(L.OP_SOFTMAX_FWD, dtype, m, n), ops.plan_for(L.OP_SOFTMAX_BWD, dtype, m, n)]
            elif kind == "layernorm_bwd":
                st = [ops.layer_norm_fwd(x, None, None, EPS) for x in xs]
                t = timeit(lambda i: ops.layer_norm_fwd(xs[i % nsets], wgt, wgt, EPS))
                out.append(("ln fwd", t, iobytes("layernorm_fwd", m, n, es)["fwd"]))
                t = timeit(lambda i: ops.layer_norm_bwd(dys[i % nsets], xs[i % nsets], st[i % nsets][1],
                                                        st[i % nsets][2], wgt))
                out.append(("bwd+fin", t, byt["bwd"]))
                fl = L.F_HAS_WEIGHT | L.F_OUT_DX | L.F_OUT_DW | L.F_OUT_DB
                plans = [ops.plan_for(L.OP_LAYERNORM_FWD, dtype, m, n, L.F_HAS_WEIGHT | L.F_HAS_BIAS),
                         ops.plan_for(L.OP_LAYERNORM_BWD, dtype, m, n, fl)]
            else:
                t = timeit(lambda i: ops.layer_norm_fwd(xs[i % nsets], wgt, wgt, EPS))
                out.append(("fwd", t, byt["fwd"]))
                plans = [ops.plan_for(L.OP_LAYERNORM_FWD, dtype, m, n, L.F_HAS_WEIGHT | L.F_HAS_BIAS)]
            res = "  ".join(f"{nm} {t:7.1f}us {b / t / 1e3:5.0f}GB/s {100 * b / t / 1e3 / peak:5.1f}%" for nm, t, b in out)
            pl = " | ".join(f"{p['variant']} C{p['cluster_size']} tpr{p['threads_per_row']} blk{p['block']} st{p['stages']}"
                            for p in plans)
            print(f"{cfg or 'default':40s} {res}   [{pl}]", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"{cfg or 'default':40s} FAILED: {e}", flush=True)
            torch.cuda.synchronize()


if __name__ == "__main__":
    main()

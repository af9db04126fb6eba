#!/usr/bin/env python
"""Time single ops at arbitrary shapes (back-to-back launches, rotating inputs > L2, CUDA events).

    python scripts/kbench.py "ln_bwd:32768x12288:bf16[:ENV=V,ENV=V]" ...
ops: osm_fwd osm_bwd onorm_fwd onorm_bwd (softmax / batch-norm-shaped normalization over dim 0: the outer-persistent kernels) rms_fwd add_rms_fwd add_ln_fwd ln_fwd sm_fwd sm_bwd rms_bwd rms_bwd_dx ln_bwd ln_bwd_dx copy (torch copy_, 1 read + 1 write: the
same-size calibration of what the memory system gives a plain streaming kernel)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import measured_peak  # noqa: E402
from fuser_b200 import ops  # noqa: E402

KEYS = ["B200NORM_NO_TILE2", "B200NORM_DROP_CH4", "B200NORM_COOP", "B200NORM_NO_OUTER_ROWS", "B200NORM_OUTER_ROWS", "B200NORM_NO_LEAN", "B200NORM_LN_PIPE", "B200NORM_NO_LN_PIPE", "B200NORM_TPR", "B200NORM_GROUPS", "B200NORM_DEPTH", "B200NORM_TILE_ROWS", "B200NORM_CLUSTER",
        "B200NORM_BWD_REREAD", "B200NORM_NO_FUSED_FINISH", "B200NORM_NO_PDL", "B200NORM_NO_PSMEM", "B200NORM_COOP", "B200NORM_STATIC", "B200NORM_POOL", "B200NORM_NO_PIPE", "B200NORM_PIPE",
        "B200NORM_OUTER_W", "B200NORM_OUTER_DEPTH", "B200NORM_OUTER_UNITS", "B200NORM_OUTER_PER_SM", "B200NORM_OUTER_SKIP_REDUCE", "B200NORM_OUTER_PLAIN_ST", "B200NORM_OUTER_TWO_EXCH", "B200NORM_OUTER_L2PF"]
DT = {"bf16": torch.bfloat16, "f32": torch.float32, "f16": torch.float16}


def algbytes(op, m, n, s):
    return {"add_rms_fwd": s * (4 * m * n + n) + 4 * m, "add_ln_fwd": s * (4 * m * n + 2 * n) + 8 * m,
            "copy": s * 2 * m * n, "red_sum": s * (m * n + m), "red_max": s * (m * n + m), "colsum": s * (m * n + n),
            "var_mean": s * m * n + 8 * m, "rms_fwd": s * (2 * m * n + n) + 4 * m, "ln_fwd": s * 2 * (m * n + n) + 8 * m, "sm_fwd": s * 2 * m * n, "osm_fwd": s * 2 * m * n, "onorm_fwd": s * 2 * (m * n + n) + 8 * n, "onorm_bwd": s * 3 * (m * n + n) + 8 * n, "osm_bwd": s * 3 * m * n,
            "sm_bwd": s * 3 * m * n, "rms_bwd": s * (3 * m * n + 2 * n) + 4 * m, "rms_bwd_dx": s * (3 * m * n + 2 * n) + 4 * m,
            "ln_bwd": s * 3 * (m * n + n) + 8 * m, "ln_bwd_dx": s * 3 * (m * n + n) + 8 * m,
            "inorm_fwd": s * 2 * m * n + 8 * m, "inorm_bwd": s * 3 * m * n + 16 * m,  # rows = B * 128 channels
            "ce_fwd": s * m * n + 16 * m, "ce_bwd": s * 2 * m * n + 12 * m,
            # fused dropout (dropout_{layernorm,rmsnorm}_{fwd,bwd}_iobytes): + residual, + the bool mask, + d residual
            "drop_ln_fwd": s * (3 * m * n + 2 * n) + 8 * m + m * n, "drop_rms_fwd": s * (3 * m * n + n) + 4 * m + m * n,
            "drop_ln_bwd": s * (5 * m * n + 3 * n) + 8 * m + m * n, "drop_rms_bwd": s * (5 * m * n + 2 * n) + 4 * m + m * n}[op]


def main():
    peak, _ = measured_peak()
    dev = torch.device("cuda", 0)
    cache = {}
    for spec in sys.argv[1:]:
        parts = spec.split(":")
        op, shape, dt = parts[0], parts[1], parts[2]
        cfg = parts[3] if len(parts) > 3 else ""
        m, n = (int(v) for v in shape.split("x"))
        dtype = DT[dt]
        es = torch.empty(0, dtype=dtype).element_size()
        for k in KEYS:
            os.environ.pop(k, None)
        for kv in filter(None, cfg.split(",")):
            k, v = kv.split("=")
            os.environ[k] = v
        key = (m, n, dt)
        if key not in cache:
            cache.clear()
            torch.cuda.empty_cache()
            nsets = max(2, int(300e6 // (m * n * es)) + 1)
            g = torch.Generator(device="cpu").manual_seed(1)
            xs = [torch.randn(m, n, generator=g).to(dtype).to(dev) for _ in range(nsets)]
            dys = [torch.randn(m, n, generator=g).to(dtype).to(dev) for _ in range(nsets)]
            w = torch.randn(n, generator=g).to(dtype).to(dev)
            cache[key] = (nsets, xs, dys, w, {})
        nsets, xs, dys, w, aux = cache[key]
        try:
            if op in ("rms_bwd", "rms_bwd_dx") and "rms" not in aux:
                aux["rms"] = [ops.rms_norm_fwd(x, w, 1e-5)[1] for x in xs]
            if op in ("ln_bwd", "ln_bwd_dx") and "ln" not in aux:
                aux["ln"] = [ops.layer_norm_fwd(x, None, None, 1e-5)[1:] for x in xs]
            if op == "sm_bwd" and "sm" not in aux:
                aux["sm"] = [ops.softmax_fwd(x) for x in xs]
            if op == "osm_bwd" and "osm" not in aux:
                aux["osm"] = [torch.softmax(x.float(), 0).to(dtype) for x in xs]
            if op == "onorm_bwd" and "onorm" not in aux:
                aux["onorm"] = [ops.outer_norm_fwd(x, w, w, 1e-5)[1:] for x in xs]
            if op.startswith("drop_") and "drop" not in aux:
                rng = ops.philox_state(17, 0, dev)
                aux["rng"] = rng
                aux["drop"] = [ops.dropout_layer_norm_fwd(x, dy, w, w, 1e-5, 0.2, rng) for x, dy in zip(xs, dys)]
                aux["droprms"] = [ops.dropout_rms_norm_fwd(x, dy, w, 1e-5, 0.2, rng) for x, dy in zip(xs, dys)]
            if op.startswith("inorm") and "inorm" not in aux:
                wc = torch.randn(128, device=dev).to(dtype)
                aux["inorm"] = (wc, [ops.instance_norm_fwd(x.view(m // 128, 128, n), wc, wc, 1e-5)[1:] for x in xs])
            if op.startswith("ce_") and "ce" not in aux:
                tg = [torch.randint(0, n, (m,), device=dev) for _ in xs]
                aux["ce"] = (tg, [ops.nll_softmax_fwd(x, t)[1] for x, t in zip(xs, tg)], torch.full((1,), 1.0 / m, device=dev))
            if op == "copy" and "out" not in aux:
                aux["out"] = torch.empty_like(xs[0])
            fns = {
                "copy": lambda i: aux["out"].copy_(xs[i]),
                "red_sum": lambda i: ops.reduce_inner(xs[i], "sum"),
                "red_max": lambda i: ops.reduce_inner(xs[i], "max"),
                "colsum": lambda i: ops.reduce_outer(xs[i]),
                "var_mean": lambda i: ops.var_mean(xs[i], 1),
                "rms_fwd": lambda i: ops.rms_norm_fwd(xs[i], w, 1e-5),
                "add_rms_fwd": lambda i: ops.rms_norm_fwd(xs[i], w, 1e-5, residual=dys[i], want_sum=True),
                "add_ln_fwd": lambda i: ops.layer_norm_fwd(xs[i], w, w, 1e-5, residual=dys[i], want_sum=True),
                "ln_fwd": lambda i: ops.layer_norm_fwd(xs[i], w, w, 1e-5),
                "sm_fwd": lambda i: ops.softmax_fwd(xs[i]),
                "osm_fwd": lambda i: ops.softmax_fwd(xs[i], dim=0),
                "onorm_fwd": lambda i: ops.outer_norm_fwd(xs[i], w, w, 1e-5),
                "onorm_bwd": lambda i: ops.outer_norm_bwd(dys[i], xs[i], aux["onorm"][i][0], aux["onorm"][i][1], w),
                "osm_bwd": lambda i: ops.softmax_bwd(dys[i], aux["osm"][i], dim=0),
                "sm_bwd": lambda i: ops.softmax_bwd(dys[i], aux["sm"][i]),
                "rms_bwd": lambda i: ops.rms_norm_bwd(dys[i], xs[i], aux["rms"][i], w),
                "rms_bwd_dx": lambda i: ops.rms_norm_bwd(dys[i], xs[i], aux["rms"][i], w, (True, False)),
                "drop_ln_fwd": lambda i: ops.dropout_layer_norm_fwd(xs[i], dys[i], w, w, 1e-5, 0.2, aux["rng"]),
                "drop_rms_fwd": lambda i: ops.dropout_rms_norm_fwd(xs[i], dys[i], w, 1e-5, 0.2, aux["rng"]),
                "drop_ln_bwd": lambda i: ops.dropout_layer_norm_bwd(dys[(i + 1) % nsets], xs[i], dys[i], aux["drop"][i][3],
                                                                    aux["drop"][i][1], aux["drop"][i][2], w, 0.2),
                "drop_rms_bwd": lambda i: ops.dropout_rms_norm_bwd(dys[(i + 1) % nsets], xs[i], dys[i], aux["droprms"][i][2],
                                                                   aux["droprms"][i][1], w, 0.2),
                "inorm_fwd": lambda i: ops.instance_norm_fwd(xs[i].view(m // 128, 128, n), aux["inorm"][0], aux["inorm"][0], 1e-5),
                "inorm_bwd": lambda i: ops.instance_norm_bwd(dys[i].view(m // 128, 128, n), xs[i].view(m // 128, 128, n),
                                                             aux["inorm"][1][i][0], aux["inorm"][1][i][1], aux["inorm"][0]),
                "ce_fwd": lambda i: ops.nll_softmax_fwd(xs[i], aux["ce"][0][i]),
                "ce_bwd": lambda i: ops.nll_softmax_bwd(xs[i], aux["ce"][0][i], aux["ce"][1][i], aux["ce"][2]),
                "ln_bwd": lambda i: ops.layer_norm_bwd(dys[i], xs[i], aux["ln"][i][0], aux["ln"][i][1], w),
                "ln_bwd_dx": lambda i: ops.layer_norm_bwd(dys[i], xs[i], aux["ln"][i][0], aux["ln"][i][1], w,
                                                          (True, False, False)),
            }
            fn = fns[op]
            for i in range(3):
                fn(i % nsets)
            torch.cuda.synchronize()
            iters = max(10, min(200, int(3e9 // algbytes(op, m, n, es))))
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(iters):
                fn(i % nsets)
            b.record()
            torch.cuda.synchronize()
            t = a.elapsed_time(b) / iters * 1e3
            gbs = algbytes(op, m, n, es) / t / 1e3
            print(f"{op:10s} {shape:12s} {dt:4s} {cfg or 'default':44s} {t:8.1f}us {gbs:6.0f}GB/s {100 * gbs / peak:5.1f}%",
                  flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"{op:10s} {shape:12s} {dt:4s} {cfg or 'default':44s} FAILED {e}", flush=True)
            torch.cuda.synchronize()


if __name__ == "__main__":
    main()

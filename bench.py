#!/usr/bin/env python
"""bench.py — effective HBM GB/s of the normalization hot path on B200.

A "step" is one pass of the hot path over one batch of synthetic input. The default
workload is BASELINE.json configs[1]: RMSNorm forward + backward, bf16, [16384, 4096]
(the Llama-3-8B shape) per GPU; `--workload c3|c4|c5|c1` select the other configs for
exploration (they are parity-test cases, not the headline bench line).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2]

Metric (same definition as the reference: benchmarks/python/core.py:194-202,
csrc/fusion_profiler.cpp:786-791): algorithmic IO bytes (the reference's *_iobytes
formulas) / time. `value` is device-timed with inputs resident in HBM; `e2e` runs the
same step through the public API from pinned HOST buffers with the H2D / D2H copies
inside the timed region. `roofline` is the dominant kernel (the backward) against the
measured copy bandwidth in MEASURED_PEAKS.json. `cpu_baseline` is the ATen oracle on the
host cores on a bounded row sample. `--impl reference` times the reference's own CPU
evaluation path (ATen op-by-op, what testValidate runs) with all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # name: (description, rows, cols, dtype, kind)
    "c1": ("LayerNorm fwd fp32 [2048,4096]", 2048, 4096, torch.float32, "layernorm_fwd"),
    "c2": ("RMSNorm fwd+bwd bf16 [16384,4096]", 16384, 4096, torch.bfloat16, "rmsnorm_fwd_bwd"),
    "c3": ("softmax fwd+bwd bf16 [65536,8192]", 65536, 8192, torch.bfloat16, "softmax_fwd_bwd"),
    "c4": ("LayerNorm bwd bf16 [32768,12288]", 32768, 12288, torch.bfloat16, "layernorm_bwd"),
    "c5": ("RMSNorm fwd+bwd bf16 [4096,65536]", 4096, 65536, torch.bfloat16, "rmsnorm_fwd_bwd"),
}
EPS = 1e-5


def iobytes(kind: str, m: int, n: int, s: int) -> dict:
    """The reference's IO-byte accounting per op (benchmarks/python/test_*_{fwd,bwd}.py)."""
    if kind == "rmsnorm_fwd_bwd":
        return {"fwd": s * (2 * m * n + n) + 4 * m, "bwd": s * (3 * m * n + 2 * n) + 4 * m}
    if kind == "softmax_fwd_bwd":
        return {"fwd": s * 2 * m * n, "bwd": s * 3 * m * n}
    if kind == "layernorm_bwd":
        return {"bwd": s * 3 * (m * n + n) + 4 * 2 * m}
    if kind == "layernorm_fwd":
        return {"fwd": s * 2 * (m * n + n) + 4 * 2 * m}
    raise KeyError(kind)


def measured_peak() -> tuple[float, str]:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------
# clocks sampling during the timed region
# ---------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.samples: list[list[str]] = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        def pump():
            for line in self.proc.stdout:
                parts = [p.strip() for p in line.split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "nvidia-smi unavailable"}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for p in self.samples:
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for name, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------
# the hot-path step for each workload, on device tensors, through the public API
# ---------------------------------------------------------------------------
class Step:
    def __init__(self, kind, m, n, dtype, device, world, nsets=2):
        from fuser_b200 import ops

        self.ops, self.kind, self.m, self.n, self.dtype, self.world = ops, kind, m, n, dtype, world
        g = torch.Generator(device="cpu").manual_seed(1234 + (torch.distributed.get_rank() if world > 1 else 0))
        self.host = []
        self.sets = []
        for _ in range(nsets):  # rotating input sets: per-step footprint >> L2 (126 MB)
            x = torch.randn(m, n, generator=g, dtype=torch.float32).to(dtype)
            dy = torch.randn(m, n, generator=g, dtype=torch.float32).to(dtype)
            self.host.append((x.pin_memory(), dy.pin_memory()))
            self.sets.append((x.to(device), dy.to(device)))
        self.w_host = torch.randn(n, generator=g).to(dtype).pin_memory()
        self.w = self.w_host.to(device)
        self.device = device
        # saved statistics / forward outputs computed from the inputs (not random)
        self.saved = []
        for x, dy in self.sets:
            if kind == "layernorm_bwd":
                _, mean, invstd = ops.layer_norm_fwd(x, None, None, EPS)
                self.saved.append((mean, invstd))
            else:
                self.saved.append(None)
        torch.cuda.synchronize(device)

    def run(self, i, ev=None):
        ops = self.ops
        x, dy = self.sets[i % len(self.sets)]
        out = None
        if ev:
            ev[0].record()
        if self.kind == "rmsnorm_fwd_bwd":
            y, rms = ops.rms_norm_fwd(x, self.w, EPS, stat_is_rms=True)
            if ev:
                ev[1].record()
            if self.world > 1:
                dx, sums = ops.rms_norm_bwd(dy, x, rms, self.w, stat_is_rms=True, centered=False, partials_f32=True)
                torch.distributed.all_reduce(sums)
                (dw,) = ops.finalize_partials(sums, self.dtype)
            else:
                dx, dw = ops.rms_norm_bwd(dy, x, rms, self.w, stat_is_rms=True, centered=False)
            out = (y, rms, dx, dw)
        elif self.kind == "softmax_fwd_bwd":
            y = ops.softmax_fwd(x)
            if ev:
                ev[1].record()
            dx = ops.softmax_bwd(dy, y)
            out = (y, dx)
        elif self.kind == "layernorm_bwd":
            mean, invstd = self.saved[i % len(self.sets)]
            if ev:
                ev[1].record()
            if self.world > 1:
                dx, sums = ops.layer_norm_bwd(dy, x, mean, invstd, self.w, partials_f32=True)
                torch.distributed.all_reduce(sums)
                dw, db = ops.finalize_partials(sums, self.dtype)
            else:
                dx, dw, db = ops.layer_norm_bwd(dy, x, mean, invstd, self.w)
            out = (dx, dw, db)
        elif self.kind == "layernorm_fwd":
            if ev:
                ev[1].record()
            out = ops.layer_norm_fwd(x, self.w, self.w, EPS)
        if ev:
            ev[2].record()
        return out

    def run_e2e(self, i):
        """Same step from pinned host buffers: H2D of the step's inputs, compute, D2H of the results."""
        hx, hdy = self.host[i % len(self.host)]
        dev = self.device
        x = hx.to(dev, non_blocking=True)
        dy = hdy.to(dev, non_blocking=True)
        w = self.w_host.to(dev, non_blocking=True)
        h2d = hx.nbytes + hdy.nbytes + self.w_host.nbytes
        ops = self.ops
        if self.kind == "rmsnorm_fwd_bwd":
            y, rms = ops.rms_norm_fwd(x, w, EPS, stat_is_rms=True)
            dx, dw = ops.rms_norm_bwd(dy, x, rms, w, stat_is_rms=True, centered=False)
            outs = (y, rms, dx, dw)
        elif self.kind == "softmax_fwd_bwd":
            y = ops.softmax_fwd(x)
            dx = ops.softmax_bwd(dy, y)
            outs = (y, dx)
        elif self.kind == "layernorm_bwd":
            mean, invstd = self.saved[i % len(self.sets)]
            outs = ops.layer_norm_bwd(dy, x, mean, invstd, w)
        else:
            outs = ops.layer_norm_fwd(x, w, w, EPS)
        res = [o.to("cpu", non_blocking=False) for o in outs if o is not None]
        d2h = sum(r.nbytes for r in res)
        return h2d, d2h


# ---------------------------------------------------------------------------
# CPU arm: the reference's CPU evaluation of the same fusion (ATen op-by-op)
# ---------------------------------------------------------------------------
def cpu_step_fn(kind, dtype):
    from oracle import aten_oracle as AO

    def fn(x, dy, w):
        if kind == "rmsnorm_fwd_bwd":
            y, rms = AO.rms_norm_fwd_python(x, w, EPS)
            return AO.rms_norm_bwd_python(x, rms, dy, w)
        if kind == "softmax_fwd_bwd":
            y = AO.softmax_fwd(x)
            return AO.softmax_bwd_python(y, dy)
        if kind == "layernorm_bwd":
            _, mean, invstd = AO.layer_norm_fwd(x, None, None, EPS)
            return AO.layer_norm_bwd_python(x, dy, mean.view(-1), invstd, w)
        return AO.layer_norm_fwd(x, w, w, EPS)

    return fn


def time_cpu(kind, m, n, dtype, sample_rows, reps, warm=1):
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(sample_rows, n, generator=g).to(dtype)
    dy = torch.randn(sample_rows, n, generator=g).to(dtype)
    w = torch.randn(n, generator=g).to(dtype)
    fn = cpu_step_fn(kind, dtype)
    for _ in range(warm):
        fn(x, dy, w)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn(x, dy, w)
        ts.append(time.perf_counter() - t0)
    sec = statistics.median(ts)
    nbytes = sum(iobytes(kind, sample_rows, n, x.element_size()).values())
    return nbytes / sec / 1e9, sec, torch.get_num_threads()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200norm", choices=["b200norm", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200norm" else args.warmup

    desc, M, N, dtype, kind = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dtype_name = {torch.bfloat16: "bf16", torch.float32: "f32", torch.float16: "f16"}[dtype]
    metric = "LayerNorm/RMSNorm fwd+bwd effective HBM GB/s"
    es = torch.empty(0, dtype=dtype).element_size()

    # ---------------- reference arm: CPU, rank 0 only ----------------
    if args.impl == "reference":
        if rank != 0:
            return
        sample_rows = max(64, min(M, (1 << 27) // (N * es)))  # ~128 MB per streamed tensor
        reps = max(1, args.steps)
        gbs, sec, cores = time_cpu(kind, M, N, dtype, sample_rows, reps, warm=max(1, args.warmup))
        sample = f"{sample_rows} of {M} rows x {N} cols {dtype_name}, median of {reps} steps, ATen op-by-op (testValidate path)"
        line = {
            "impl": "reference", "metric": metric, "value": gbs, "unit": "GB/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3 * (M / sample_rows),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": desc, "rows_per_gpu": M, "cols": N, "io_dtype": dtype_name,
                                            "l2": "inputs larger than L2"},
            "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line), flush=True)
        return

    # ---------------- our arm ----------------
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: fuser_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=device)
    from fuser_b200 import _lib as L

    rows = M
    scaling = "weak"
    if args.workload == "c4" and world > 1:
        rows = M // world  # the named multi-GPU config: row-sharded, total work fixed
        scaling = "strong"
    step = Step(kind, rows, N, dtype, device, world)
    byt = iobytes(kind, rows, N, es)
    step_bytes = sum(byt.values())

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(device)

    for i in range(args.warmup):
        step.run(i)
    barrier()

    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.1)
    launches0 = L.launch_count()
    barrier()
    t_start.record()
    for i in range(args.steps):
        step.run(i, evs[i])
    t_end.record()
    barrier()
    launches = L.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else {}
    total_ms = t_start.elapsed_time(t_end)
    tms = torch.tensor([total_ms], device=device)
    if world > 1:
        torch.distributed.all_reduce(tms, op=torch.distributed.ReduceOp.MAX)
    total_ms = float(tms.item())
    ms_per_step = total_ms / args.steps
    value = step_bytes * world / (ms_per_step * 1e-3) / 1e9

    fwd_ms = statistics.mean(e[0].elapsed_time(e[1]) for e in evs)
    bwd_ms = statistics.mean(e[1].elapsed_time(e[2]) for e in evs)
    peak, peak_src = measured_peak()
    dom = "bwd" if "bwd" in byt else "fwd"
    dom_ms = bwd_ms if dom == "bwd" else fwd_ms + bwd_ms
    achieved = byt[dom] / (dom_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": f"{kind}:{dom}", "peak_source": peak_src,
                "per_kernel": {k: {"bytes": byt[k], "ms": (fwd_ms if k == "fwd" else bwd_ms),
                                   "gbs": byt[k] / ((fwd_ms if k == "fwd" else bwd_ms) * 1e-3) / 1e9} for k in byt
                               if (fwd_ms if k == "fwd" else bwd_ms) > 0}}

    # ---------------- e2e: host buffers, copies inside the timed region ----------------
    e2e = None
    if not args.no_e2e:
        n_e2e = max(2, min(args.steps, 5))
        step.run_e2e(0)
        barrier()
        t0 = time.perf_counter()
        for i in range(n_e2e):
            h2d, d2h = step.run_e2e(i)
        torch.cuda.synchronize(device)
        sec = (time.perf_counter() - t0) / n_e2e
        te = torch.tensor([sec], device=device)
        if world > 1:
            torch.distributed.all_reduce(te, op=torch.distributed.ReduceOp.MAX)
        sec = float(te.item())
        e2e = {"value": step_bytes * world / sec / 1e9, "unit": "GB/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": sec * 1e3, "steps": n_e2e}

    # ---------------- CPU baseline (rank 0, N=1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sample_rows = max(64, min(rows, (1 << 26) // (N * es)))
        gbs, sec, cores = time_cpu(kind, rows, N, dtype, sample_rows, reps=5)
        cpu = {"value": gbs, "unit": "GB/s", "cores": cores, "kind": "port",
               "sample": f"{sample_rows} of {rows} rows x {N} cols {dtype_name}, median of 5, ATen oracle op-by-op"}

    if rank == 0:
        line = {
            "metric": metric, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "rows_per_gpu": rows, "cols": N, "io_dtype": dtype_name,
                       "accumulate": "fp32", "l2": "inputs larger than L2 (2 rotating input sets, 640 MB/step footprint)",
                       "collective": "nccl allreduce of fp32 dgamma partial sums" if world > 1 else "none",
                       "rows_per_s": rows * world / (ms_per_step * 1e-3),
                       "pct_of_nominal_8TBs": 100.0 * value / world / 8000.0,
                       "pct_of_measured_peak": 100.0 * value / world / peak},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py — effective HBM GB/s of the normalization hot path on B200.

A "step" is one pass of the hot path over one batch of synthetic input. The default
workload is BASELINE.json configs[1]: RMSNorm forward + backward, bf16, [16384, 4096]
(the Llama-3-8B shape) per GPU; `--workload c3|c4|c5|c1` select the other configs for
exploration (they are parity-test cases, not the headline bench line).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2]

Metric (same definition as the reference: benchmarks/python/core.py:194-202,
csrc/fusion_profiler.cpp:786-791): algorithmic IO bytes (the reference's *_iobytes
formulas) / time.
  value     K steps of the device path (inputs resident in HBM), one CUDA-event pair around
            the whole region, barrier + synchronize on both sides, max over ranks.
  roofline  the dominant kernel (the backward) launched back to back K times on rotating
            inputs, CUDA events on the launching stream -> average launch duration, against
            the measured copy bandwidth in MEASURED_PEAKS.json.
  e2e       the same step through the reference-facing API (FusionDefinition.execute) from
            pinned HOST buffers: H2D of the step's inputs and D2H of its results inside the
            timed region, double-buffered over three streams.
  cpu_baseline  the ATen oracle on the host cores on a bounded row sample.
`--impl reference` times the reference's own CPU evaluation path (ATen op-by-op, what
testValidate runs) with all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
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
    # SURVEY section 8f rank 2 (not a BASELINE.json config): the reference's outer_persistent benchmark,
    # softmax over dim 0 (benchmarks/python/test_softmax_{fwd,bwd}.py, reduction_axis = 0)
    "o1": ("softmax over dim 0 fwd+bwd bf16 [8192,18432] (outer-persistent)", 8192, 18432, torch.bfloat16,
           "softmax_outer_fwd_bwd"),
    # batch-norm-shaped normalization over dim 0, training statistics, channels last
    # (benchmarks/python/normalization.py, test_batchnorm_{fwd,bwd}.py viewed as [N*H*W, C])
    "o2": ("batch norm fwd+bwd bf16 [8192,18432] (statistics over dim 0, outer-persistent)", 8192, 18432, torch.bfloat16,
           "outer_norm_fwd_bwd"),
    # ... and at the ResNet-50 size N*H*W = 32*56*56, C = 256, beyond what a cluster holds: the row-block kernel
    "o3": ("batch norm fwd+bwd bf16 [100352,256] (statistics over dim 0, row-block kernel)", 100352, 256, torch.bfloat16,
           "outer_norm_fwd_bwd"),
    # plain Reduction segments and stand-alone variance_mean on the fast path
    "r1": ("row sum + var_mean + column sum bf16 [32768,12288] (Reduction / Welford segments)", 32768, 12288, torch.bfloat16,
           "reductions"),
    # vocabulary-size rows (tests/cpp/test_cluster.cpp:158-211; not a BASELINE.json config): the pipelined one-exchange
    # softmax forward over a cluster of four, and without a cluster at a 32 K vocabulary
    "v1": ("softmax fwd+bwd bf16 [4096,131072] (vocabulary-size rows, cluster)", 4096, 131072, torch.bfloat16, "softmax_fwd_bwd"),
    "v2": ("softmax fwd+bwd bf16 [16384,32768] (32 K vocabulary)", 16384, 32768, torch.bfloat16, "softmax_fwd_bwd"),
}
EPS = 1e-5
L2_BYTES = 126 * 1024 * 1024


def iobytes(kind: str, m: int, n: int, s: int) -> dict:
    """The reference's IO-byte accounting per op (benchmarks/python/test_*_{fwd,bwd}.py)."""
    if kind == "rmsnorm_fwd_bwd":
        return {"fwd": s * (2 * m * n + n) + 4 * m, "bwd": s * (3 * m * n + 2 * n) + 4 * m}
    if kind in ("softmax_fwd_bwd", "softmax_outer_fwd_bwd"):
        return {"fwd": s * 2 * m * n, "bwd": s * 3 * m * n}
    if kind == "outer_norm_fwd_bwd":
        return {"fwd": s * 2 * (m * n + n) + 4 * 2 * n, "bwd": s * 3 * (m * n + n) + 4 * 2 * n}
    if kind == "layernorm_bwd":
        return {"bwd": s * 3 * (m * n + n) + 4 * 2 * m}
    if kind == "layernorm_fwd":
        return {"fwd": s * 2 * (m * n + n) + 4 * 2 * m}
    if kind == "reductions":  # fwd: row sum + var_mean of x (two reads of x); bwd: column sum of dy
        return {"fwd": s * (m * n + m) + s * m * n + 8 * m, "bwd": s * (m * n + n)}
    raise KeyError(kind)


def measured_peak() -> tuple[float, str]:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------
# clocks sampling during the timed regions (NVML, ~2 ms period)
# ---------------------------------------------------------------------------
class ClockSampler:
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index: int):
        self.index = index
        self.sm: list[int] = []
        self.max_sm = None
        self.reasons: set[str] = set()
        self._stop = threading.Event()
        self._active = threading.Event()
        self.thread = None
        self.note = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            # CUDA_VISIBLE_DEVICES may remap indices: resolve through the PCI bus id of the torch device
            bus = torch.cuda.get_device_properties(index).pci_bus_id if hasattr(
                torch.cuda.get_device_properties(index), "pci_bus_id") else None
            self.h = None
            if bus is not None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    h = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if pynvml.nvmlDeviceGetPciInfo(h).bus == bus:
                        self.h = h
                        break
            if self.h is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            self.nv = None
            self.note = f"nvml unavailable: {e}"

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            if self._active.is_set():
                try:
                    self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    for name, bit in self.REASONS.items():
                        if r & bit:
                            self.reasons.add(name)
                except Exception:  # noqa: BLE001
                    pass
            time.sleep(0.002)

    def start(self):
        if self.nv is None:
            return
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def resume(self):
        self._active.set()

    def pause(self):
        self._active.clear()

    def stop(self) -> dict:
        self._stop.set()
        if self.thread is not None:
            self.thread.join(timeout=1)
        out = {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_sm,
               "reasons": sorted(self.reasons), "samples": len(self.sm)}
        if self.note:
            out["note"] = self.note
        return out


# ---------------------------------------------------------------------------
# the hot-path step for each workload, on device tensors, through the public ops API
# ---------------------------------------------------------------------------
class Step:
    def __init__(self, kind, m, n, dtype, device, world, rank, collective="peer-fused"):
        from fuser_b200 import ops, parallel

        self.ops, self.parallel = ops, parallel
        self.kind, self.m, self.n, self.dtype, self.world, self.device = kind, m, n, dtype, world, device
        es = torch.empty(0, dtype=dtype).element_size()
        # rotating input sets so that consecutive steps never find their inputs in L2
        self.nsets = max(2, int(2.5 * L2_BYTES // (2 * m * n * es)) + 1)
        g = torch.Generator(device="cpu").manual_seed(1234 + rank)
        self.sets = []
        for _ in range(self.nsets):
            x = torch.randn(m, n, generator=g, dtype=torch.float32).to(dtype)
            dy = torch.randn(m, n, generator=g, dtype=torch.float32).to(dtype)
            self.sets.append((x.to(device), dy.to(device)))
        self.w = torch.randn(n, generator=g).to(dtype).to(device)
        # dgamma / dbeta across ranks: all-reduced inside the backward kernel over peer memory
        # (PeerExchange); --collective nccl (or a box without P2P symmetric memory) uses the NCCL
        # all-reduce of the fp32 partial sums instead
        self.reducer, self.collective = None, "none"
        if world > 1:
            if collective in ("peer", "peer-fused"):
                try:
                    self.reducer = parallel.PeerExchange(max_cols=n, device=device, deferred=collective == "peer")
                    self.collective = collective
                except Exception as e:  # noqa: BLE001
                    print(f"[bench] peer exchange unavailable ({e}); falling back to NCCL", file=sys.stderr)
            if self.reducer is None:
                self.reducer = parallel.ColumnSumReducer(device=device)
                self.collective = "nccl"
        # saved statistics / forward outputs computed from the inputs (not random)
        self.saved = []
        for x, dy in self.sets:
            if kind == "layernorm_bwd":
                _, mean, invstd = ops.layer_norm_fwd(x, None, None, EPS)
                self.saved.append((mean, invstd))
            elif kind == "rmsnorm_fwd_bwd":
                self.saved.append(ops.rms_norm_fwd(x, self.w, EPS, stat_is_rms=True)[1])
            elif kind == "softmax_fwd_bwd":
                self.saved.append(ops.softmax_fwd(x))
            elif kind == "softmax_outer_fwd_bwd":
                self.saved.append(ops.softmax_fwd(x, dim=0))
            elif kind == "outer_norm_fwd_bwd":
                self.saved.append(ops.outer_norm_fwd(x, self.w, self.w, EPS)[1:])
            else:
                self.saved.append(None)
        torch.cuda.synchronize(device)

    # -- the two halves, separately launchable (roofline loops) -----------------------------
    def fwd(self, i):
        ops = self.ops
        x, _ = self.sets[i % self.nsets]
        if self.kind == "rmsnorm_fwd_bwd":
            return ops.rms_norm_fwd(x, self.w, EPS, stat_is_rms=True)
        if self.kind == "softmax_fwd_bwd":
            return (ops.softmax_fwd(x),)
        if self.kind == "softmax_outer_fwd_bwd":
            return (ops.softmax_fwd(x, dim=0),)
        if self.kind == "outer_norm_fwd_bwd":
            return ops.outer_norm_fwd(x, self.w, self.w, EPS)
        if self.kind == "layernorm_fwd":
            return ops.layer_norm_fwd(x, self.w, self.w, EPS)
        if self.kind == "reductions":
            var, mean = ops.var_mean(x, 1)
            return ops.reduce_inner(x, "sum"), var, mean
        return None

    def bwd(self, i, fwd_out=None):
        ops = self.ops
        x, dy = self.sets[i % self.nsets]
        saved = self.saved[i % self.nsets]
        if self.kind == "rmsnorm_fwd_bwd":
            rms = fwd_out[1] if fwd_out is not None else saved
            if self.world > 1:
                prev = self.reducer.pending  # NCCL: last step's dgamma — its all-reduce had a whole step to finish
                dx, h = self.parallel.rms_norm_bwd_sharded(dy, x, rms, self.w, self.reducer, stat_is_rms=True)
                if self.collective == "peer-fused":
                    return dx, h.result()  # reduced inside the kernel: this step's dgamma, final
                return dx, (prev.result() if prev is not None else None)
            return ops.rms_norm_bwd(dy, x, rms, self.w, stat_is_rms=True, centered=False)
        if self.kind == "softmax_fwd_bwd":
            return (ops.softmax_bwd(dy, fwd_out[0] if fwd_out is not None else saved),)
        if self.kind == "softmax_outer_fwd_bwd":
            return (ops.softmax_bwd(dy, fwd_out[0] if fwd_out is not None else saved, dim=0),)
        if self.kind == "outer_norm_fwd_bwd":
            mean, invstd = (fwd_out[1], fwd_out[2]) if fwd_out is not None else saved
            return ops.outer_norm_bwd(dy, x, mean, invstd, self.w)
        if self.kind == "reductions":
            return (ops.reduce_outer(dy),)
        if self.kind == "layernorm_bwd":
            mean, invstd = saved
            if self.world > 1:
                prev = self.reducer.pending
                dx, h = self.parallel.layer_norm_bwd_sharded(dy, x, mean, invstd, self.w, self.reducer)
                if self.collective == "peer-fused":
                    return dx, h.result()
                return dx, (prev.result() if prev is not None else None)
            return ops.layer_norm_bwd(dy, x, mean, invstd, self.w)
        return None

    def run(self, i):
        f = self.fwd(i)
        b = self.bwd(i, f)
        return f, b

    def finish(self):
        if self.reducer is not None:
            self.reducer.wait()


# ---------------------------------------------------------------------------
# e2e: host buffers -> FusionDefinition.execute -> host buffers, three streams
# ---------------------------------------------------------------------------
_SAVED_AFFINITY: list = []


def unbind_host_memory():
    """Undo bind_host_memory (the CPU baseline must see every host core again)."""
    try:
        import ctypes

        if _SAVED_AFFINITY:
            os.sched_setaffinity(0, _SAVED_AFFINITY.pop())
        ctypes.CDLL(None, use_errno=True).syscall(238, 0, None, ctypes.c_ulong(0))  # MPOL_DEFAULT
    except Exception:  # noqa: BLE001
        pass


def bind_host_memory(device) -> dict:
    """Best effort: put this rank's pinned host buffers (and its thread) on the NUMA node its GPU hangs off. With
    one process per GPU all started from the same shell, every rank's pinned memory otherwise lands on the node of
    the launching CPUs and half of the GPUs copy across the socket interconnect (r01: e2e 5.8 ms per step at
    1 GPU, 32 ms at 8). Uses the PCI device's numa_node from sysfs, sched_setaffinity to that node's CPUs (when the
    cpuset allows) and set_mempolicy(MPOL_PREFERRED) through libc. Reports what it found and did."""
    info = {"gpu_numa_node": None, "nodes": None, "cpus_bound": False, "mempolicy": False}
    try:
        import ctypes
        import glob

        nodes = sorted(int(p.rsplit("node", 1)[1]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
        info["nodes"] = len(nodes)
        bus = torch.cuda.get_device_properties(device).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(device), "pci_domain_id", 0)
        dev_id = torch.cuda.get_device_properties(device).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev_id:02x}.0/numa_node"
        node = int(open(path).read().strip())
        info["gpu_numa_node"] = node
        if node < 0 or len(nodes) < 2:
            return info
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        if cpus & allowed:
            os.sched_setaffinity(0, cpus & allowed)
            info["cpus_bound"] = True
            _SAVED_AFFINITY.append(allowed)
        libc = ctypes.CDLL(None, use_errno=True)
        mask = ctypes.c_ulong(1 << node)
        MPOL_PREFERRED, SYS_set_mempolicy = 1, 238  # x86_64
        rc = libc.syscall(SYS_set_mempolicy, MPOL_PREFERRED, ctypes.byref(mask), ctypes.c_ulong(64))
        info["mempolicy"] = rc == 0
    except Exception as e:  # noqa: BLE001
        info["note"] = f"{type(e).__name__}: {e}"
    return info


class E2E:
    def __init__(self, kind, m, n, dtype, device, rank):
        from tests import fusions as F
        from fuser_b200.frontend import FusionDefinition, torch_dtype_to_nvfuser_dtype

        self.kind, self.m, self.n, self.dtype, self.device = kind, m, n, dtype, device
        nvdt = torch_dtype_to_nvfuser_dtype(dtype)

        def define(fn, *a):
            with FusionDefinition() as fd:
                fn(fd, nvdt, *a)
            return fd

        self.numa = bind_host_memory(device)
        g = torch.Generator(device="cpu").manual_seed(4321 + rank)
        self.hx = torch.randn(m, n, generator=g).to(dtype).pin_memory()
        self.hdy = torch.randn(m, n, generator=g).to(dtype).pin_memory()
        self.hw = torch.randn(n, generator=g).to(dtype).pin_memory()
        self.fd_fwd = self.fd_bwd = None
        if kind == "rmsnorm_fwd_bwd":
            self.fd_fwd, self.fd_bwd = define(F.rmsnorm_fwd_py, EPS), define(F.rmsnorm_bwd_py)
        elif kind == "softmax_fwd_bwd":
            self.fd_fwd, self.fd_bwd = define(F.softmax_fwd_py), define(F.softmax_bwd_py)
        elif kind == "softmax_outer_fwd_bwd":
            self.fd_fwd, self.fd_bwd = define(F.softmax_fwd_py, 0), define(F.softmax_bwd_py, 0)
        elif kind == "outer_norm_fwd_bwd":
            self.fd_fwd, self.fd_bwd = define(F.batchnorm_fwd_py, 2, EPS), define(F.batchnorm_bwd_py, 2)
            # the running statistics of the reference's definitions: inputs that only feed dead code
            self.rm = torch.zeros(n, device=device)
            self.rv = torch.ones(n, device=device)
        elif kind == "layernorm_bwd":
            self.fd_fwd, self.fd_bwd = define(F.layernorm_fwd_py, EPS), define(F.layernorm_bwd_py)
        elif kind == "reductions":
            self.fd_fwd, self.fd_bwd = define(F.reduce_py, "sum", [1]), define(F.reduce_py, "sum", [0])
            self.fd_vm = define(F.var_mean_py, 1)
        else:
            self.fd_fwd = define(F.layernorm_fwd_py, EPS)
        self.s_in, self.s_c, self.s_out = (torch.cuda.Stream(device=device) for _ in range(3))
        # double-buffered device inputs and pinned host outputs
        self.dbuf = [(torch.empty_like(self.hx, device=device), torch.empty_like(self.hdy, device=device),
                      torch.empty_like(self.hw, device=device)) for _ in range(2)]
        self.hout = [None, None]
        self.ev_in = [torch.cuda.Event() for _ in range(2)]
        self.ev_c = [torch.cuda.Event() for _ in range(2)]
        self.ev_out = [torch.cuda.Event() for _ in range(2)]
        self.h2d = self.d2h = 0
        self.keep = [None, None]

    def compute(self, x, dy, w):
        k = self.kind
        if k == "rmsnorm_fwd_bwd":
            y, rms = self.fd_fwd.execute([x, w])
            dx, dw = self.fd_bwd.execute([x, rms, dy, w])
            return [y, rms, dx, dw]
        if k in ("softmax_fwd_bwd", "softmax_outer_fwd_bwd"):
            (y,) = self.fd_fwd.execute([x])
            (dx,) = self.fd_bwd.execute([y, dy])
            return [y, dx]
        if k == "outer_norm_fwd_bwd":
            y, mean, invstd = self.fd_fwd.execute([x, w, w, self.rm, self.rv])
            dx, dw, db = self.fd_bwd.execute([x, dy, w, self.rm, self.rv, mean, invstd])
            return [y, mean, invstd, dx, dw, db]
        if k == "reductions":
            (rs,) = self.fd_fwd.execute([x])
            var, mean = self.fd_vm.execute([x])
            (cs,) = self.fd_bwd.execute([dy])
            return [rs, var, mean, cs]
        if k == "layernorm_bwd":
            # the config is the backward; the saved statistics come from the forward definition
            y, mean, invstd = self.fd_fwd.execute([x, w, w])
            dx, dw, db = self.fd_bwd.execute([x, dy, mean, invstd, w])
            return [dx, dw, db]
        return list(self.fd_fwd.execute([x, w, w]))

    def step(self, i):
        b = i % 2
        dx_, ddy_, dw_ = self.dbuf[b]
        with torch.cuda.stream(self.s_in):
            self.s_in.wait_event(self.ev_c[b])  # the compute that last read this buffer is done
            dx_.copy_(self.hx, non_blocking=True)
            ddy_.copy_(self.hdy, non_blocking=True)
            dw_.copy_(self.hw, non_blocking=True)
            self.ev_in[b].record(self.s_in)
        with torch.cuda.stream(self.s_c):
            self.s_c.wait_event(self.ev_in[b])
            outs = self.compute(dx_, ddy_, dw_)
            self.ev_c[b].record(self.s_c)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_c[b])
            self.ev_out[b].synchronize()  # the host buffers of step i-2 have landed
            if self.hout[b] is None:
                self.hout[b] = [torch.empty(o.shape, dtype=o.dtype).pin_memory() for o in outs]
            for h, o in zip(self.hout[b], outs):
                o.record_stream(self.s_out)
                h.copy_(o, non_blocking=True)
            self.ev_out[b].record(self.s_out)
        self.keep[b] = outs
        self.h2d = self.hx.nbytes + self.hdy.nbytes + self.hw.nbytes
        self.d2h = sum(o.nbytes for o in outs)

    def sync(self):
        for s in (self.s_in, self.s_c, self.s_out):
            s.synchronize()


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
        if kind == "softmax_outer_fwd_bwd":
            y = AO.softmax_fwd(x, dim=0)
            return AO.softmax_bwd_python(y, dy, dim=0)
        if kind == "outer_norm_fwd_bwd":
            y, mean, invstd = AO.outer_norm_fwd(x, w, w, EPS)
            return AO.outer_norm_bwd(dy, x, mean, invstd, w)
        if kind == "layernorm_bwd":
            _, mean, invstd = AO.layer_norm_fwd(x, None, None, EPS)
            return AO.layer_norm_bwd_python(x, dy, mean.view(-1), invstd, w)
        if kind == "reductions":
            return AO.reduce_inner(x, "sum"), AO.var_mean(x, 1), AO.reduce_outer(dy)
        return AO.layer_norm_fwd(x, w, w, EPS)

    return fn


def validate_step(step, kind, world, device):
    """One step checked against the ATen oracle with the validator's tolerances before anything is timed
    (the reference's benchmarks call fd.validate first: benchmarks/python/test_layernorm_fwd.py:90-95).
    y / dx rows are independent: above 2^26 elements a row sample is compared (first / middle / last blocks
    and random rows); dgamma / dbeta are compared over ALL rows (summed over the ranks when row-sharded)."""
    from oracle import aten_oracle as AO

    x, dy = step.sets[0]
    w = step.w
    m, n = x.shape
    f, b = step.run(0)
    pend = step.reducer.pending if step.reducer is not None else None  # nccl / deferred peer: this step's handle
    step.finish()
    torch.cuda.synchronize(device)
    sharded_cols = None
    if world > 1 and kind in ("rmsnorm_fwd_bwd", "layernorm_bwd"):
        sharded_cols = b[1] if step.collective == "peer-fused" else pend.result()
    if m * n > (1 << 26):
        g = torch.Generator().manual_seed(5)
        rows = torch.unique(torch.cat([torch.arange(0, 64), torch.arange(m // 2, m // 2 + 64), torch.arange(m - 64, m),
                                       torch.randint(0, m, (64,), generator=g)]))
    else:
        rows = torch.arange(m)
    rd = rows.to(device)
    xs, dys, wc = x[rd].cpu(), dy[rd].cpu(), w.cpu()
    checked = []

    def close(actual, expected, red, what):
        AO.assert_close(actual, expected, red, what)
        checked.append(what)

    def col_sums(name, actual, local_f32):
        """column sums: the oracle's fp32 sums of this rank's rows, added over the ranks, cast once."""
        t = local_f32.to(device)
        if world > 1:
            torch.distributed.all_reduce(t)
        close(actual, t.to(actual.dtype).cpu(), m * world, name)

    if kind == "rmsnorm_fwd_bwd":
        y, rms = f
        dx, dw = b[0], b[1]
        ye, rmse = AO.rms_norm_fwd_python(xs, wc, EPS)
        close(y[rd], ye, n, "y")
        close(rms[rd], rmse, n, "rms")
        dxe, _ = AO.rms_norm_bwd_python(xs, rms[rd].cpu(), dys, wc)
        close(dx[rd], dxe, n, "dx")
        if sharded_cols is not None:
            dw = sharded_cols[0]
        col_sums("dgamma", dw, (dy.cpu().float() * (x.cpu().float() / rms.cpu())).sum(0))
    elif kind in ("softmax_fwd_bwd", "softmax_outer_fwd_bwd"):
        dim = 0 if kind == "softmax_outer_fwd_bwd" else -1
        (y,), (dx,) = f, b
        if dim == 0:
            close(y, AO.softmax_fwd(x.cpu(), dim=0), m * m, "y")
            close(dx, AO.softmax_bwd(dy.cpu(), y.cpu(), dim=0), m, "dx")
        else:
            close(y[rd], AO.softmax_fwd(xs), n * n, "y")
            close(dx[rd], AO.softmax_bwd(dys, y[rd].cpu()), n, "dx")
    elif kind == "outer_norm_fwd_bwd":
        ye, me, ie = AO.outer_norm_fwd(x.cpu(), wc, wc, EPS)
        close(f[0], ye, m, "y")
        close(f[1], me, m, "mean")
        close(f[2], ie, m, "invstd")
        for name, a, e in zip(("dx", "dgamma", "dbeta"), b, AO.outer_norm_bwd(dy.cpu(), x.cpu(), f[1].cpu(), f[2].cpu(), wc)):
            close(a, e, m, name)
    elif kind == "layernorm_bwd":
        mean, invstd = step.saved[0]
        dx, cols = b[0], b[1:]
        if sharded_cols is not None:
            cols = sharded_cols
        dxe, _, _ = AO.layer_norm_bwd(dys, xs, mean[rd].cpu(), invstd[rd].cpu(), wc)
        close(dx[rd], dxe, n, "dx")
        xc, dyc = x.cpu().float(), dy.cpu().float()
        col_sums("dgamma", cols[0], (dyc * ((xc - mean.cpu()) * invstd.cpu())).sum(0))
        col_sums("dbeta", cols[1], dyc.sum(0))
    elif kind == "reductions":
        close(f[0][rd], AO.reduce_inner(xs, "sum"), n, "row sum")
        ve, me = AO.var_mean(xs, 1)
        close(f[1].view(-1)[rd], ve, n, "var")
        close(f[2].view(-1)[rd], me, n, "mean")
        close(b[0], AO.reduce_outer(dy.cpu()), m, "column sum")
    else:  # layernorm_fwd
        for name, a, e in zip(("y", "mean", "invstd"), f, AO.layer_norm_fwd(x.cpu(), wc, wc, EPS)):
            close(a, e, n, name)
    return {"oracle": "ATen op-by-op on the host (oracle/aten_oracle.py), validator tolerances", "outputs": checked,
            "rows_compared": int(rows.numel()), "rows": int(m)}


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


def c4_strong_record(args, device, world, rank, collective, barrier):
    """BASELINE configs[3]: LayerNorm backward bf16 [32768,12288], rows sharded over the ranks, dgamma / dbeta
    all-reduced; total work fixed. Timed like the main line (K launches, one event pair, max over ranks); rank 0
    then times the UNSHARDED problem alone (the other ranks idle) so the record carries its own 1-GPU time."""
    from fuser_b200.parallel import shard_rows

    _, M, N, dtype, kind = WORKLOADS["c4"]
    a, b = shard_rows(M, world, rank)
    total_bytes = sum(iobytes(kind, M, N, 2).values())
    st = Step(kind, b - a, N, dtype, device, world, rank, collective)
    nk = max(20, min(args.steps, 100))
    for i in range(5):
        st.bwd(i)
    st.finish()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(nk):
        st.bwd(i)
    st.finish()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / nk], device=device)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms = float(t.item())
    used = st.collective
    if st.reducer is not None and hasattr(st.reducer, "close"):
        st.reducer.close()
    del st
    torch.cuda.empty_cache()
    one = None
    if rank == 0:
        s1 = Step(kind, M, N, dtype, device, 1, 0)
        for i in range(3):
            s1.bwd(i)
        torch.cuda.synchronize(device)
        e0.record()
        for i in range(20):
            s1.bwd(i)
        e1.record()
        torch.cuda.synchronize(device)
        one = e0.elapsed_time(e1) / 20
        del s1
        torch.cuda.empty_cache()
    barrier()
    return {"workload": WORKLOADS["c4"][0] + f", rows sharded over {world} GPUs", "scaling": "strong", "n_gpus": world,
            "value": total_bytes / (ms * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": ms, "launches_timed": nk,
            "one_gpu_ms": one, "speedup_vs_one_gpu": (one / ms) if one else None,
            "efficiency": (one / ms / world) if one else None, "collective": used}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200norm", choices=["b200norm", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--collective", default="peer-fused", choices=["peer-fused", "peer", "nccl"],
                    help="N > 1, dgamma/dbeta all-reduce: peer-fused = entirely inside the backward kernel over NVLink peer "
                         "memory (default: 42 %% faster than NCCL at C4 on 8 GPUs, within 1-4 %% of it on 2); peer = published "
                         "by the backward kernel and combined by a small kernel a step later; nccl = NCCL all-reduce of the "
                         "fp32 partial sums, consumed a step later")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-c4", action="store_true", help="N > 1: skip the C4 strong-scaling sub-record")
    ap.add_argument("--no-validate", action="store_true", help="skip the oracle check of one step before timing")
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
        reps = max(1, min(args.steps, 20))
        gbs, sec, cores = time_cpu(kind, M, N, dtype, sample_rows, reps, warm=max(1, min(args.warmup, 3)))
        sample = (f"{sample_rows} of {M} rows x {N} cols {dtype_name}, median of {reps} steps, ATen op-by-op "
                  f"(testValidate path), all host threads")
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
        from fuser_b200.parallel import shard_rows

        a, b = shard_rows(M, world, rank)  # the named multi-GPU config: row-sharded, total work fixed
        rows = b - a
        scaling = "strong"
    step = Step(kind, rows, N, dtype, device, world, rank, args.collective)
    byt = iobytes(kind, rows, N, es)
    step_bytes = sum(byt.values())
    total_bytes = step_bytes * world if scaling == "weak" else sum(iobytes(kind, M, N, es).values())

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(device)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    # ---- one step against the oracle before anything is timed (every rank checks its own rows) ----
    validation = None
    if not args.no_validate:
        validation = validate_step(step, kind, world, device)

    for i in range(args.warmup):
        step.run(i)
    step.finish()
    barrier()

    # ---- timed region: K steps, one event pair ----
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = L.launch_count()
    barrier()
    sampler.resume()
    h0 = time.perf_counter()
    t_start.record()
    for i in range(args.steps):
        step.run(i)
    step.finish()
    t_end.record()
    host_enqueue_s = time.perf_counter() - h0
    barrier()
    sampler.pause()
    launches = L.launch_count() - launches0
    total_ms = t_start.elapsed_time(t_end)
    tms = torch.tensor([total_ms], device=device)
    if world > 1:
        torch.distributed.all_reduce(tms, op=torch.distributed.ReduceOp.MAX)
    total_ms = float(tms.item())
    ms_per_step = total_ms / args.steps
    value = total_bytes / (ms_per_step * 1e-3) / 1e9

    # ---- roofline: each kernel back to back on rotating inputs, events on the launching stream ----
    def kernel_loop(fn, n):
        for i in range(3):
            fn(i)
        step.finish()
        torch.cuda.synchronize(device)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler.resume()
        a.record()
        for i in range(n):
            fn(i)
        b.record()
        step.finish()
        torch.cuda.synchronize(device)
        sampler.pause()
        return a.elapsed_time(b) / n

    nk = max(20, min(args.steps, 200))
    per_kernel = {}
    if "fwd" in byt:
        ms = kernel_loop(step.fwd, nk)
        per_kernel["fwd"] = {"bytes": byt["fwd"], "ms": ms, "gbs": byt["fwd"] / (ms * 1e-3) / 1e9}
    if "bwd" in byt:
        ms = kernel_loop(lambda i: step.bwd(i), nk)
        per_kernel["bwd"] = {"bytes": byt["bwd"], "ms": ms, "gbs": byt["bwd"] / (ms * 1e-3) / 1e9}
    peak, peak_src = measured_peak()
    dom = "bwd" if "bwd" in byt else "fwd"
    achieved = per_kernel[dom]["gbs"]
    # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this
    # workload (profiles/traffic.json, written by scripts/ncu_summary.py); null when none is committed
    # or the problem was row-sharded (the capture is of the 1-GPU problem).
    traffic = traffic_src = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and rows == M:
        ent = json.load(open(tpath)).get(args.workload, {}).get(dom)
        if ent:
            traffic, traffic_src = ent["dram_bytes"], f"profiles/{ent['from']}"
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src,
                "traffic_note": "a constant read from the committed ncu --set full capture named in traffic_source, "
                                "not measured in this run", "algorithmic_bytes": byt[dom],
                "kernel": f"{kind}:{dom}", "peak_source": peak_src, "launches_timed": nk,
                "per_kernel": per_kernel,
                "step_frac_of_peak": (step_bytes / (ms_per_step * 1e-3) / 1e9) / peak}

    # ---- e2e: host buffers, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        e = E2E(kind, rows, N, dtype, device, rank)
        n_e2e = max(4, min(args.steps, 20))
        for i in range(2):
            e.step(i)
        e.sync()
        barrier()
        t0 = time.perf_counter()
        for i in range(n_e2e):
            e.step(i)
        e.sync()
        sec = (time.perf_counter() - t0) / n_e2e
        te = torch.tensor([sec], device=device)
        if world > 1:
            torch.distributed.all_reduce(te, op=torch.distributed.ReduceOp.MAX)
        sec = float(te.item())
        e2e = {"value": total_bytes / sec / 1e9, "unit": "GB/s", "h2d_bytes_per_step": e.h2d,
               "d2h_bytes_per_step": e.d2h, "ms_per_step": sec * 1e3, "steps": n_e2e, "host_numa": e.numa,
               "api": "FusionDefinition.execute (fwd + bwd definitions of benchmarks/python), pinned host buffers, "
                      "H2D / compute / D2H double-buffered on three streams"}
        del e
        unbind_host_memory()
    # ---- N > 1: the north-star multi-GPU config (C4, LayerNorm backward [32768,12288] row-sharded, STRONG
    # scaling) as a sub-record of the default line, with its own 1-GPU time measured in the same run ----
    c4_strong = None
    if world > 1 and args.workload == "c2" and not args.no_c4:
        c4_strong = c4_strong_record(args, device, world, rank, step.collective, barrier)
    clocks = sampler.stop() if rank == 0 else {}

    # ---- CPU baseline (rank 0, N=1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sample_rows = max(64, min(rows, (1 << 26) // (N * es)))
        gbs, sec, cores = time_cpu(kind, rows, N, dtype, sample_rows, reps=5)
        cpu = {"value": gbs, "unit": "GB/s", "cores": cores, "kind": "port",
               "sample": f"{sample_rows} of {rows} rows x {N} cols {dtype_name}, median of 5, ATen oracle op-by-op"}

    if rank == 0:
        footprint = step.nsets * 2 * rows * N * es
        line = {
            "metric": metric, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            # `config` names the workload and nothing else, so that this line and the reference arm's compare equal
            "config": {"workload": desc, "rows_per_gpu": rows, "cols": N, "io_dtype": dtype_name,
                       "l2": "inputs larger than L2"},
            "details": {"accumulate": "fp32",
                        "value_api": "fuser_b200.ops -> C++ NormExecutor -> C ABI (device tensors); e2e goes through "
                                     "FusionDefinition.execute",
                        "l2": f"{step.nsets} rotating input sets, {footprint / 1e6:.0f} MB",
                        "collective": {"none": "none",
                                       "peer": "fp32 dgamma/dbeta sums published to NVLink peer memory by the backward kernel's "
                                               "finishing CTAs; a small combine kernel (P2P loads in rank order, cast) consumes "
                                               "them one step later; no NCCL",
                                       "peer-fused": "fp32 dgamma/dbeta sums all-reduced inside the backward kernel over NVLink "
                                                     "peer memory: one-shot push of (value, epoch) words into every rank's "
                                                     "receive area, collected locally in rank order, cast after the reduce",
                                       "nccl": "nccl allreduce of fp32 dgamma/dbeta partial sums, async (consumed one step "
                                               "later), cast after the reduce"}[step.collective],
                        "rows_per_s": (rows * world if scaling == "weak" else M) / (ms_per_step * 1e-3),
                        "pct_of_nominal_8TBs": 100.0 * value / world / 8000.0,
                        "pct_of_measured_peak": 100.0 * value / world / peak,
                        "host_enqueue_us_per_step": host_enqueue_s / args.steps * 1e6},
            "validated": validation is not None, "validation": validation,
            "c4_strong": c4_strong, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()

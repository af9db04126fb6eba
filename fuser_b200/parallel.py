"""Row-sharded execution of the hot path across the GPUs of one box.

Every row (token) of a normalization / softmax segment is independent, so forward, softmax
and dx need no communication: rank r simply owns rows ``[r*M/P, (r+1)*M/P)``. The only
exchange on the path is the backward's dgamma / dbeta, which are sums over *all* rows: each
rank produces fp32 partial sums ``[K, N]`` (``B200NORM_F_PARTIALS_F32``), the ranks
all-reduce them (NCCL over NVLink / NVSwitch), and only then is the result cast to the io
dtype — one rounding, exactly as on a single GPU.

Reference mechanism this replaces: a DeviceMesh-sharded fusion whose cross-device sum is
lowered to ``Communication(Allreduce)`` (csrc/host_ir/lower_to_communication.cpp:266-284) and
posted with ``postAllreduce`` (csrc/multidevice/post_communication.cpp:288-313); ranks and
world size come from the launcher's environment as in csrc/multidevice/communicator.cpp:78-101.

The all-reduce is latency bound (2*N fp32 = 96 KB at N = 12288), so it is only *enqueued*
when the partial sums exist (NCCL runs it on its own stream) and its result is consumed later,
after the compute stream has done more work.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from . import ops


def shard_rows(rows: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous row block of ``rank``: [start, stop). Earlier ranks take the remainder."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} / world {world}")
    base, rem = divmod(rows, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def local_rows(t: torch.Tensor, world: int, rank: int) -> torch.Tensor:
    """The rank's row block of a [M, ...] tensor (a view)."""
    a, b = shard_rows(t.shape[0], world, rank)
    return t[a:b]


class PendingColumnSums:
    """Handle of an in-flight all-reduce of fp32 column sums; ``result()`` orders the current
    stream after it and casts to the io dtype (one rounding, after the reduce)."""

    __slots__ = ("sums", "dtype", "work", "_outs")

    def __init__(self, sums, dtype, work):
        self.sums, self.dtype, self.work, self._outs = sums, dtype, work, None

    def result(self):
        if self._outs is None:
            if self.work is not None:
                self.work.wait()  # stream-level wait: the host does not block
            self._outs = ops.finalize_partials(self.sums, self.dtype)
        return self._outs


class ColumnSumReducer:
    """All-reduce of the backward's fp32 column sums, overlapped with whatever the compute
    stream does next: ``reduce()`` only enqueues the collective (NCCL runs it on its own
    stream behind an event of the current stream) and returns a handle; ``result()`` of the
    handle is called when the gradients are needed — typically after more backward work —
    and costs a stream wait plus the cast kernel.
    """

    def __init__(self, group: Optional[dist.ProcessGroup] = None, device: Optional[torch.device] = None):
        self.group = group
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.pending: Optional[PendingColumnSums] = None

    def reduce(self, sums: torch.Tensor, dtype: torch.dtype) -> PendingColumnSums:
        work = None
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            work = dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self.pending = PendingColumnSums(sums, dtype, work)
        return self.pending

    def wait(self):
        """Finish the last reduction on the current stream."""
        if self.pending is not None:
            self.pending.result()


def layer_norm_bwd_sharded(dy, x, mean, invstd, weight, reducer: ColumnSumReducer):
    """Row-sharded LayerNorm backward: local dx now, (dweight, dbias) as a pending handle."""
    dx, sums = ops.layer_norm_bwd(dy, x, mean, invstd, weight, partials_f32=True)
    return dx, reducer.reduce(sums, x.dtype)


def rms_norm_bwd_sharded(dy, x, stat, weight, reducer: ColumnSumReducer, *, stat_is_rms=False, centered=False):
    """Row-sharded RMSNorm backward: local dx now, (dweight,) as a pending handle."""
    dx, sums = ops.rms_norm_bwd(dy, x, stat, weight, stat_is_rms=stat_is_rms, centered=centered, partials_f32=True)
    return dx, reducer.reduce(sums, x.dtype)

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

from . import _lib as L
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


class ReadyColumnSums:
    """Same interface as PendingColumnSums for gradients that are already final."""

    __slots__ = ("_outs",)

    def __init__(self, outs):
        self._outs = list(outs)

    def result(self):
        return self._outs


class PublishedColumnSums:
    """Handle of column sums published to the peers by a backward launched with ``deferred=True``;
    ``result()`` launches the one-CTA combine (wait for the peers' flags — normally long there —, add
    the ranks' slots in rank order, cast)."""

    __slots__ = ("px", "k", "n", "dtype", "_outs")

    def __init__(self, px, k, n, dtype):
        self.px, self.k, self.n, self.dtype, self._outs = px, k, n, dtype, None

    def result(self):
        if self._outs is None:
            self.px._combine_up_to(self)
        return self._outs


class PeerExchange:
    """Peer-memory exchange buffers for the fused all-reduce of the row-sharded backward: the
    backward kernel's finishing CTAs write this rank's fp32 column sums into its buffer and signal
    every peer over NVLink; the sums of all ranks are then read with P2P loads, added in rank order
    and cast — no NCCL kernel and no host synchronisation.

    Two flavours (``deferred``):
      * False (``B200NORM_F_PEER_REDUCE``): everything inside the backward kernel, gradients final
        when it returns: the finishing CTAs push their sums into every rank's receive area (one 64-bit
        (value, epoch) store per sum and peer) and collect the peers' from their own memory — but the
        ranks meet inside the kernel every step, so any skew between them is paid at once;
      * True (``B200NORM_F_PEER_PUBLISH`` + ``OP_PEER_COMBINE``, the default): the backward only
        publishes; ``result()`` of the returned handle launches a one-CTA combine, typically a step
        later when every peer's flag has long arrived, so no rank ever waits for another.

    The buffers are torch symmetric memory (``torch.distributed._symmetric_memory``: CUDA VMM
    allocations exported to the other ranks of the box — the role of the reference's
    csrc/multidevice/symmetric_tensor.cpp / ipc_handle.cpp). ``max_cols`` bounds N of the segments
    that will use it (``b200norm_peer_exchange_bytes``: the publish slots plus, for the fused flavour, a
    receive area of two parities x world x 2N (value, epoch) words). Construction is collective over ``group``.
    """

    HEADER = 256

    def __init__(self, max_cols: int, group: Optional[dist.ProcessGroup] = None,
                 device: Optional[torch.device] = None, deferred: bool = True):
        import torch.distributed._symmetric_memory as symm

        self.deferred = deferred
        self._outstanding: list = []  # published, not yet combined (oldest first; at most two)

        if not dist.is_initialized():
            raise RuntimeError("PeerExchange needs an initialised process group")
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if self.world > 8:
            raise ValueError("peer exchange covers the GPUs of one box (world <= 8)")
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.nbytes = L.peer_exchange_bytes(self.world, 2 * int(max_cols))  # K <= 2 sums per column
        self.buf = symm.empty(self.nbytes, dtype=torch.uint8, device=self.device)
        self.buf.zero_()
        self.handle = symm.rendezvous(self.buf, self.group)
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)  # every rank's buffer is zero before anybody's first launch
        self.ptrs = [int(p) for p in self.handle.buffer_ptrs]
        L.register_peers(self.world, self.rank, self.ptrs, self.nbytes)
        self.pending = None  # interface parity with ColumnSumReducer

    def publish_handle(self, k: int, n: int, dtype) -> "PublishedColumnSums":
        h = PublishedColumnSums(self, k, n, dtype)
        self._outstanding.append(h)
        self.pending = h
        return h

    def make_room(self):
        """Two slots: combine the oldest publish before a third one is launched."""
        while len(self._outstanding) >= 2:
            self._combine_up_to(self._outstanding[0])

    def _combine_up_to(self, handle):
        # combines pair up with publishes in order
        while self._outstanding:
            h = self._outstanding.pop(0)
            h._outs = ops.peer_combine(h.k, h.n, h.dtype, self.buf)
            if h is handle:
                break

    def wait(self):
        """Combine everything that is still outstanding (on the current stream)."""
        while self._outstanding:
            self._combine_up_to(self._outstanding[0])

    def close(self):
        L.register_peers(0, 0, [], 0)


def layer_norm_bwd_sharded(dy, x, mean, invstd, weight, reducer):
    """Row-sharded LayerNorm backward: local dx now; (dweight, dbias) behind ``.result()`` — pending
    on an NCCL all-reduce (``ColumnSumReducer``) or already reduced inside the kernel (``PeerExchange``)."""
    if isinstance(reducer, PeerExchange):
        if x.shape[0] == 0:
            raise ValueError("every rank must hold at least one row for the fused peer reduce")
        if reducer.deferred:
            reducer.make_room()
            dx, _, _ = ops.layer_norm_bwd(dy, x, mean, invstd, weight, peer_reduce="publish")
            return dx, reducer.publish_handle(2, x.shape[-1], x.dtype)
        dx, dw, db = ops.layer_norm_bwd(dy, x, mean, invstd, weight, peer_reduce=True)
        return dx, ReadyColumnSums([dw, db])
    dx, sums = ops.layer_norm_bwd(dy, x, mean, invstd, weight, partials_f32=True)
    return dx, reducer.reduce(sums, x.dtype)


def rms_norm_bwd_sharded(dy, x, stat, weight, reducer, *, stat_is_rms=False, centered=False):
    """Row-sharded RMSNorm backward: local dx now; (dweight,) behind ``.result()``."""
    if isinstance(reducer, PeerExchange):
        if x.shape[0] == 0:
            raise ValueError("every rank must hold at least one row for the fused peer reduce")
        if reducer.deferred:
            reducer.make_room()
            dx, _ = ops.rms_norm_bwd(dy, x, stat, weight, stat_is_rms=stat_is_rms, centered=centered,
                                     peer_reduce="publish")
            return dx, reducer.publish_handle(1, x.shape[-1], x.dtype)
        dx, dw = ops.rms_norm_bwd(dy, x, stat, weight, stat_is_rms=stat_is_rms, centered=centered, peer_reduce=True)
        return dx, ReadyColumnSums([dw])
    dx, sums = ops.rms_norm_bwd(dy, x, stat, weight, stat_is_rms=stat_is_rms, centered=centered, partials_f32=True)
    return dx, reducer.reduce(sums, x.dtype)

"""Host-side executor for matched normalization / reduction segments.

This is the Python mirror of the executor the reference would slot into
``ExecutorDispatch`` (reference: csrc/runtime/cutlass_executor.cpp:58-115 —
infer output shapes, ``at::empty_strided`` the outputs, size a byte workspace,
launch on the current stream). torch is used only for device memory and the
current stream; every FLOP happens inside ``libb200norm.so`` reached through the
C ABI (``include/b200norm.h``). There is no eager / CPU fallback: if the library
is missing or declines a segment, an exception is raised.

Function names and argument meaning follow the reference's op builders in
csrc/ops/normalization.cpp (``layer_norm`` :304-342, ``rms_norm`` :352-380,
``layer_norm_backward`` :382-436, ``rms_norm_backward`` :438-485,
``thunder_rms_norm_backward`` :487-535, ``softmax`` :169-189,
``softmax_backward`` :191-210, ``log_softmax`` :212-231,
``log_softmax_backward`` :233-249).
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import _lib as L

_DT = {torch.float32: L.DT_F32, torch.bfloat16: L.DT_BF16, torch.float16: L.DT_F16}

# workspace cache: (device index, stream handle) -> zero-initialised byte tensor
_workspaces: dict[tuple[int, int], torch.Tensor] = {}

# force a kernel family for every call (tests / tuning); see L.VARIANT_*
_forced_variant = L.VARIANT_AUTO
_forced_cluster = 0


def force_variant(variant: int = L.VARIANT_AUTO, cluster_size: int = 0):
    """Testing hook: pin the kernel family / cluster size (like NVFUSER_ENABLE options)."""
    global _forced_variant, _forced_cluster
    _forced_variant, _forced_cluster = variant, cluster_size


def _dtype_id(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise L.UnsupportedSegment(L.ERR_UNSUPPORTED, f"dtype {t.dtype} is not handled by b200norm") from None


def _require_cuda(*tensors: Optional[torch.Tensor]):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise L.B200NormError(
                L.ERR_NO_DEVICE, "b200norm kernels run on CUDA tensors only (there is no CPU fallback)"
            )


def _as_rows(t: torch.Tensor, n: int) -> torch.Tensor:
    """View an N-D tensor as [M, n] without copying; the kernels take any row stride."""
    if t.dim() == 2 and t.shape[1] == n:
        return t
    if t.numel() == 0:
        return t.reshape(-1, n)
    try:
        return t.view(-1, n)
    except RuntimeError:
        raise L.UnsupportedSegment(
            L.ERR_UNSUPPORTED, f"tensor with sizes {tuple(t.shape)} strides {t.stride()} is not viewable as [M,{n}]"
        ) from None


def _mat(t: torch.Tensor) -> L.Tensor:
    return L.make_tensor(t.data_ptr(), _dtype_id(t), list(t.shape), list(t.stride()))


def _vec(t: torch.Tensor) -> L.Tensor:
    return L.make_tensor(t.data_ptr(), _dtype_id(t), list(t.shape), list(t.stride()))


def _workspace(nbytes: int, device: torch.device) -> tuple[int, int]:
    if nbytes <= 0:
        return 0, 0
    stream = torch.cuda.current_stream(device).cuda_stream
    key = (device.index if device.index is not None else torch.cuda.current_device(), stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        # zero-filled once; the library keeps the header zero between launches
        ws = torch.zeros(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws.data_ptr(), ws.numel()


def _run(desc: L.Desc, ins: Sequence[torch.Tensor], outs: Sequence[torch.Tensor], device: torch.device):
    desc.variant = _forced_variant if desc.variant == L.VARIANT_AUTO else desc.variant
    if desc.cluster_size == 0:
        desc.cluster_size = _forced_cluster
    nbytes = L.workspace_bytes(desc)
    ws_ptr, ws_n = _workspace(nbytes, device)
    stream = torch.cuda.current_stream(device).cuda_stream
    with torch.cuda.device(device):
        L.launch(desc, [_mat(t) for t in ins], [_mat(t) for t in outs], ws_ptr, ws_n, stream)


def plan_for(op: int, dtype: torch.dtype, rows: int, cols: int, flags: int = 0, eps: float = 0.0,
             variant: int = L.VARIANT_AUTO, cluster_size: int = 0) -> dict:
    """Dispatch decision for a segment (no launch): the launch_param dump of this library."""
    d = L.make_desc(op, _DT[dtype], rows, cols, eps, flags, variant, cluster_size)
    return L.query(d).as_dict()


# ---------------------------------------------------------------------------
# forward
# ---------------------------------------------------------------------------
def layer_norm_fwd(x, weight=None, bias=None, eps: float = 1e-5, *, want_mean=True, want_invstd=True,
                   variant=L.VARIANT_AUTO, cluster_size=0):
    """y, mean[...,1], invstd[...,1] = layer_norm(x) over the last dim."""
    _require_cuda(x, weight, bias)
    n = x.shape[-1]
    x2 = _as_rows(x, n)
    m = x2.shape[0]
    y = torch.empty(x.shape, dtype=x.dtype, device=x.device)
    stat_shape = tuple(x.shape[:-1]) + (1,)
    mean = torch.empty(stat_shape, dtype=torch.float32, device=x.device) if want_mean else None
    invstd = torch.empty(stat_shape, dtype=torch.float32, device=x.device) if want_invstd else None
    flags = (L.F_HAS_WEIGHT if weight is not None else 0) | (L.F_HAS_BIAS if bias is not None else 0)
    flags |= (L.F_OUT_MEAN if want_mean else 0) | (L.F_OUT_INVSTD if want_invstd else 0)
    d = L.make_desc(L.OP_LAYERNORM_FWD, _dtype_id(x), m, n, eps, flags, variant, cluster_size)
    ins = [x2] + [t for t in (weight, bias) if t is not None]
    outs = [y.view(-1, n) if y.numel() else y.reshape(-1, n)]
    outs += [t.view(-1) for t in (mean, invstd) if t is not None]
    _run(d, ins, outs, x.device)
    return y, mean, invstd


def rms_norm_fwd(x, weight=None, eps: float = 1e-5, *, stat_is_rms=False, variant=L.VARIANT_AUTO, cluster_size=0):
    """y, stat[...,1] = rms_norm(x); stat = invstd (C++ builder) or rms (Thunder spelling)."""
    _require_cuda(x, weight)
    n = x.shape[-1]
    x2 = _as_rows(x, n)
    m = x2.shape[0]
    y = torch.empty(x.shape, dtype=x.dtype, device=x.device)
    stat = torch.empty(tuple(x.shape[:-1]) + (1,), dtype=torch.float32, device=x.device)
    flags = (L.F_HAS_WEIGHT if weight is not None else 0) | (L.F_STAT_IS_RMS if stat_is_rms else 0)
    d = L.make_desc(L.OP_RMSNORM_FWD, _dtype_id(x), m, n, eps, flags, variant, cluster_size)
    ins = [x2] + ([weight] if weight is not None else [])
    _run(d, ins, [y.reshape(-1, n), stat.view(-1)], x.device)
    return y, stat


def softmax_fwd(x, *, log=False, variant=L.VARIANT_AUTO, cluster_size=0):
    _require_cuda(x)
    n = x.shape[-1]
    x2 = _as_rows(x, n)
    y = torch.empty(x.shape, dtype=x.dtype, device=x.device)
    d = L.make_desc(L.OP_LOGSOFTMAX_FWD if log else L.OP_SOFTMAX_FWD, _dtype_id(x), x2.shape[0], n, 0.0, 0,
                    variant, cluster_size)
    _run(d, [x2], [y.reshape(-1, n)], x.device)
    return y


# ---------------------------------------------------------------------------
# backward
# ---------------------------------------------------------------------------
def softmax_bwd(dy, y, *, log=False, round_rowsum=False, variant=L.VARIANT_AUTO, cluster_size=0):
    _require_cuda(dy, y)
    n = y.shape[-1]
    dy2, y2 = _as_rows(dy, n), _as_rows(y, n)
    dx = torch.empty(y.shape, dtype=y.dtype, device=y.device)
    flags = L.F_ROUND_ROWSUM if round_rowsum else 0
    d = L.make_desc(L.OP_LOGSOFTMAX_BWD if log else L.OP_SOFTMAX_BWD, _dtype_id(y), y2.shape[0], n, 0.0, flags,
                    variant, cluster_size)
    _run(d, [dy2, y2], [dx.reshape(-1, n)], y.device)
    return dx


def _norm_bwd(op, dy, x, stats, weight, flags, output_mask, partials_f32, variant, cluster_size):
    _require_cuda(dy, x, weight, *stats)
    n = x.shape[-1]
    dy2, x2 = _as_rows(dy, n), _as_rows(x, n)
    m = x2.shape[0]
    ln = op == L.OP_LAYERNORM_BWD
    want_dx, want_dw = bool(output_mask[0]), bool(output_mask[1]) and weight is not None
    want_db = ln and len(output_mask) > 2 and bool(output_mask[2])
    flags |= (L.F_HAS_WEIGHT if weight is not None else 0)
    flags |= (L.F_OUT_DX if want_dx else 0) | (L.F_OUT_DW if want_dw else 0) | (L.F_OUT_DB if want_db else 0)
    dx = torch.empty(x.shape, dtype=x.dtype, device=x.device) if want_dx else None
    outs = [dx.reshape(-1, n)] if want_dx else []
    dw = db = sums = None
    if want_dw or want_db:
        if partials_f32:
            flags |= L.F_PARTIALS_F32
            k = 2 if ln else 1
            sums = torch.empty((k, n), dtype=torch.float32, device=x.device)
            outs.append(sums.view(-1))
        else:
            if want_dw:
                dw = torch.empty((n,), dtype=x.dtype, device=x.device)
                outs.append(dw)
            if want_db:
                db = torch.empty((n,), dtype=x.dtype, device=x.device)
                outs.append(db)
    d = L.make_desc(op, _dtype_id(x), m, n, 0.0, flags, variant, cluster_size)
    ins = [dy2, x2] + [s.reshape(-1) for s in stats] + ([weight] if weight is not None else [])
    _run(d, ins, outs, x.device)
    if partials_f32:
        return dx, sums
    return (dx, dw, db) if ln else (dx, dw)


def layer_norm_bwd(dy, x, mean, invstd, weight=None, output_mask=(True, True, True), *, partials_f32=False,
                   variant=L.VARIANT_AUTO, cluster_size=0):
    """dx, dweight, dbias (or dx, fp32 sums[2,N] with ``partials_f32``)."""
    return _norm_bwd(L.OP_LAYERNORM_BWD, dy, x, [mean, invstd], weight, 0, output_mask, partials_f32, variant,
                     cluster_size)


def rms_norm_bwd(dy, x, stat, weight=None, output_mask=(True, True), *, stat_is_rms=False, centered=False,
                 partials_f32=False, variant=L.VARIANT_AUTO, cluster_size=0):
    """dx, dweight. ``centered`` keeps the -sum(g) term of the C++ rms_norm_backward builder."""
    flags = (L.F_STAT_IS_RMS if stat_is_rms else 0) | (L.F_RMS_BWD_CENTERED if centered else 0)
    return _norm_bwd(L.OP_RMSNORM_BWD, dy, x, [stat], weight, flags, output_mask, partials_f32, variant,
                     cluster_size)


def finalize_partials(sums: torch.Tensor, dtype: torch.dtype):
    """fp32 [K,N] column sums (after the allreduce) -> K tensors [N] in ``dtype``."""
    _require_cuda(sums)
    k, n = sums.shape
    outs = [torch.empty((n,), dtype=dtype, device=sums.device) for _ in range(k)]
    d = L.make_desc(L.OP_FINALIZE_PARTIALS, _DT[dtype], k, n)
    _run(d, [sums.view(-1)], outs, sums.device)
    return outs


# ---------------------------------------------------------------------------
# plain reductions (the Reduction scheduler's segments)
# ---------------------------------------------------------------------------
def reduce_inner(x, kind: str = "sum", *, variant=L.VARIANT_AUTO):
    """[..., N] -> [...]: sum | max (NaN-propagating amax) | fmax (NaN-ignoring)."""
    _require_cuda(x)
    n = x.shape[-1]
    x2 = _as_rows(x, n)
    r = torch.empty(x.shape[:-1], dtype=x.dtype, device=x.device)
    flags = {"sum": 0, "max": L.F_REDUCE_MAX, "fmax": L.F_REDUCE_FMAX}[kind]
    d = L.make_desc(L.OP_REDUCE_INNER, _dtype_id(x), x2.shape[0], n, 0.0, flags, variant)
    _run(d, [x2], [r.reshape(-1)], x.device)
    return r


def reduce_outer(x, *, partials_f32=False, variant=L.VARIANT_AUTO):
    """[M, N] -> [N] sum over rows."""
    _require_cuda(x)
    m, n = x.shape
    out_dtype = torch.float32 if partials_f32 else x.dtype
    r = torch.empty((n,), dtype=out_dtype, device=x.device)
    flags = L.F_PARTIALS_F32 if partials_f32 else 0
    d = L.make_desc(L.OP_REDUCE_OUTER, _dtype_id(x), m, n, 0.0, flags, variant)
    _run(d, [x], [r], x.device)
    return r

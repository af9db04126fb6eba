"""Python surface of the host executor for matched normalization / reduction segments.

Each function builds the segment spec (op kind, flags, eps) and hands the tensors to the
C++ ``NormExecutor`` (``csrc/host/norm_executor.cpp`` — the mirror of the executor the
reference slots into ``ExecutorDispatch``, csrc/runtime/cutlass_executor.cpp:58-115: infer
output shapes, allocate the outputs with ATen, size a byte workspace, launch on the
current stream). torch is used only for device memory and the current stream; every FLOP
happens inside ``libb200norm.so`` reached through the C ABI (``include/b200norm.h``).
There is no eager / CPU fallback: if the library is missing or declines a segment, an
exception is raised.

Function names and argument meaning follow the reference's op builders in
csrc/ops/normalization.cpp (``layer_norm`` :304-342, ``rms_norm`` :352-380,
``layer_norm_backward`` :382-436, ``rms_norm_backward`` :438-485,
``thunder_rms_norm_backward`` :487-535, ``softmax`` :169-189,
``softmax_backward`` :191-210, ``log_softmax`` :212-231,
``log_softmax_backward`` :233-249).
"""
from __future__ import annotations

from typing import Sequence

import torch

from . import _host as H
from . import _lib as L

_DT = {torch.float32: L.DT_F32, torch.bfloat16: L.DT_BF16, torch.float16: L.DT_F16}

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


def _run(op: int, flags: int, eps: float, variant: int, cluster_size: int, ins: Sequence[torch.Tensor],
         outs: Sequence[torch.Tensor] = (), correction: int = 0, dropout_keep: float = 1.0, ignore_index: int = -100):
    """Hand the segment to the C++ NormExecutor (csrc/host/norm_executor.cpp): it allocates
    the outputs, sizes the workspace and launches on the current stream."""
    if variant == L.VARIANT_AUTO:
        variant = _forced_variant
    if cluster_size == 0:
        cluster_size = _forced_cluster
    return H.run(op, flags, float(eps), variant, cluster_size, ins, outs, correction, float(dropout_keep), int(ignore_index))


def plan_for(op: int, dtype: torch.dtype, rows: int, cols: int, flags: int = 0, eps: float = 0.0,
             variant: int = L.VARIANT_AUTO, cluster_size: int = 0, correction: int = 0, dropout_keep: float = 1.0,
             ignore_index: int = -100) -> dict:
    """Dispatch decision for a segment (no launch): the launch_param dump of this library."""
    d = L.make_desc(op, _DT[dtype], rows, cols, eps, flags, variant, cluster_size, correction, dropout_keep, ignore_index)
    return L.query(d).as_dict()


# ---------------------------------------------------------------------------
# forward
# ---------------------------------------------------------------------------
def _residual_flags(residual, want_sum, round_sum):
    if residual is None:
        if want_sum or round_sum:
            raise ValueError("want_sum / round_sum need a residual")
        return 0
    return L.F_ADD_RESIDUAL | (L.F_OUT_SUM if want_sum else 0) | (L.F_ROUND_SUM if round_sum else 0)


def layer_norm_fwd(x, weight=None, bias=None, eps: float = 1e-5, *, want_mean=True, want_invstd=True,
                   residual=None, want_sum=False, round_sum=False, variant=L.VARIANT_AUTO, cluster_size=0):
    """y, mean[...,1], invstd[...,1] = layer_norm(x) over the last dim.

    With ``residual`` the norm is taken of ``x + residual`` (the fused-add prologue of
    benchmarks/python/test_dropout_layernorm_fwd.py / test_rmsnorm_add_fwd.py); ``want_sum`` also
    returns the sum (io dtype) as a 4th value; ``round_sum`` rounds the sum through the io dtype
    before normalising, as a reduced-precision ``x + residual`` feeding the norm does."""
    flags = (L.F_HAS_WEIGHT if weight is not None else 0) | (L.F_HAS_BIAS if bias is not None else 0)
    flags |= (L.F_OUT_MEAN if want_mean else 0) | (L.F_OUT_INVSTD if want_invstd else 0)
    flags |= _residual_flags(residual, want_sum, round_sum)
    ins = [x] + ([residual] if residual is not None else []) + [t for t in (weight, bias) if t is not None]
    outs = iter(_run(L.OP_LAYERNORM_FWD, flags, eps, variant, cluster_size, ins))
    y = next(outs)
    s = next(outs) if want_sum else None
    mean = next(outs) if want_mean else None
    invstd = next(outs) if want_invstd else None
    return (y, mean, invstd, s) if want_sum else (y, mean, invstd)


def rms_norm_fwd(x, weight=None, eps: float = 1e-5, *, stat_is_rms=False, residual=None, want_sum=False,
                 round_sum=False, variant=L.VARIANT_AUTO, cluster_size=0):
    """y, stat[...,1] = rms_norm(x); stat = invstd (C++ builder) or rms (Thunder spelling).
    ``residual`` / ``want_sum`` / ``round_sum``: as in :func:`layer_norm_fwd` (sum returned 3rd)."""
    flags = (L.F_HAS_WEIGHT if weight is not None else 0) | (L.F_STAT_IS_RMS if stat_is_rms else 0)
    flags |= _residual_flags(residual, want_sum, round_sum)
    ins = [x] + ([residual] if residual is not None else []) + ([weight] if weight is not None else [])
    outs = _run(L.OP_RMSNORM_FWD, flags, eps, variant, cluster_size, ins)
    if want_sum:
        y, s, stat = outs
        return y, stat, s
    y, stat = outs
    return y, stat


def _softmax_axis_flags(t, dim: int) -> int:
    """dim = -1 / last: the inner-persistent kernels; dim = 0 of a tensor of rank >= 2: the
    outer-persistent kernels (``F_OUTER``, the reduction_axis = 0 case of
    benchmarks/python/test_softmax_{fwd,bwd}.py). Other axes are not on this path."""
    nd = t.dim()
    d = dim + nd if dim < 0 else dim
    if d == nd - 1:
        return 0
    if d == 0:
        return L.F_OUTER
    raise L.UnsupportedSegment(L.ERR_UNSUPPORTED, f"softmax over axis {dim} of a rank-{nd} tensor is not handled")


def outer_norm_fwd(x, weight=None, bias=None, eps: float = 1e-5, *, want_mean=True, want_invstd=True,
                   variant=L.VARIANT_AUTO, cluster_size=0):
    """y, mean[N], invstd[N]: normalization with the statistics taken over dim 0 (everything behind it is the
    iteration domain) and a per-column affine — the batch-norm-shaped forward in training mode with channels
    last (benchmarks/python/normalization.py:13-95 viewed as [N*H*W, C]; an OuterPersistent segment in the
    reference, scheduler/normalization_outer.cpp:451-683). Biased variance, invstd = rsqrt(var + eps)."""
    flags = L.F_OUTER | (L.F_HAS_WEIGHT if weight is not None else 0) | (L.F_HAS_BIAS if bias is not None else 0)
    flags |= (L.F_OUT_MEAN if want_mean else 0) | (L.F_OUT_INVSTD if want_invstd else 0)
    ins = [x] + [t for t in (weight, bias) if t is not None]
    outs = iter(_run(L.OP_LAYERNORM_FWD, flags, eps, variant, cluster_size, ins))
    y = next(outs)
    mean = next(outs) if want_mean else None
    invstd = next(outs) if want_invstd else None
    return y, mean, invstd


def outer_norm_bwd(dy, x, mean, invstd, weight=None, output_mask=(True, True, True), *, variant=L.VARIANT_AUTO,
                   cluster_size=0):
    """dx, dweight[N], dbias[N] of :func:`outer_norm_fwd` (batch-norm backward in training mode, channels last:
    benchmarks/python/normalization.py:98-190). ``mean`` / ``invstd`` are the saved fp32 [N] statistics."""
    want_dx, want_dw, want_db = bool(output_mask[0]), bool(output_mask[1]) and weight is not None, bool(output_mask[2])
    flags = L.F_OUTER | (L.F_HAS_WEIGHT if weight is not None else 0)
    flags |= (L.F_OUT_DX if want_dx else 0) | (L.F_OUT_DW if want_dw else 0) | (L.F_OUT_DB if want_db else 0)
    ins = [dy, x, mean, invstd] + ([weight] if weight is not None else [])
    outs = iter(_run(L.OP_LAYERNORM_BWD, flags, 0.0, variant, cluster_size, ins))
    dx = next(outs) if want_dx else None
    dw = next(outs) if want_dw else None
    db = next(outs) if want_db else None
    return dx, dw, db


def softmax_fwd(x, *, log=False, dim: int = -1, variant=L.VARIANT_AUTO, cluster_size=0):
    flags = _softmax_axis_flags(x, dim)
    (y,) = _run(L.OP_LOGSOFTMAX_FWD if log else L.OP_SOFTMAX_FWD, flags, 0.0, variant, cluster_size, [x])
    return y


# ---------------------------------------------------------------------------
# backward
# ---------------------------------------------------------------------------
def softmax_bwd(dy, y, *, log=False, round_rowsum=False, dim: int = -1, variant=L.VARIANT_AUTO, cluster_size=0):
    flags = (L.F_ROUND_ROWSUM if round_rowsum else 0) | _softmax_axis_flags(y, dim)
    (dx,) = _run(L.OP_LOGSOFTMAX_BWD if log else L.OP_SOFTMAX_BWD, flags, 0.0, variant, cluster_size, [dy, y])
    return dx


def _norm_bwd(op, dy, x, stats, weight, flags, output_mask, partials_f32, variant, cluster_size, peer_reduce=False):
    """``peer_reduce``: False | True ("fused": all-reduced inside this kernel) | "publish" (the sums are
    only published to the peers; :func:`peer_combine` produces the tensors later)."""
    ln = op == L.OP_LAYERNORM_BWD
    want_dx, want_dw = bool(output_mask[0]), bool(output_mask[1]) and weight is not None
    want_db = ln and len(output_mask) > 2 and bool(output_mask[2])
    flags |= (L.F_HAS_WEIGHT if weight is not None else 0)
    flags |= (L.F_OUT_DX if want_dx else 0) | (L.F_OUT_DW if want_dw else 0) | (L.F_OUT_DB if want_db else 0)
    if partials_f32 and (want_dw or want_db):
        flags |= L.F_PARTIALS_F32
    publish = peer_reduce == "publish" and (want_dw or want_db)
    if publish:
        flags |= L.F_PEER_PUBLISH
    elif peer_reduce and (want_dw or want_db):
        flags |= L.F_PEER_REDUCE
    ins = [dy, x] + list(stats) + ([weight] if weight is not None else [])
    outs = iter(_run(op, flags, 0.0, variant, cluster_size, ins))
    dx = next(outs) if want_dx else None
    dw = db = sums = None
    if (want_dw or want_db) and not publish:
        if partials_f32:
            sums = next(outs)
        else:
            dw = next(outs) if want_dw else None
            db = next(outs) if want_db else None
    if partials_f32:
        return dx, sums
    return (dx, dw, db) if ln else (dx, dw)


def layer_norm_bwd(dy, x, mean, invstd, weight=None, output_mask=(True, True, True), *, partials_f32=False,
                   peer_reduce=False, variant=L.VARIANT_AUTO, cluster_size=0):
    """dx, dweight, dbias (or dx, fp32 sums[2,N] with ``partials_f32``). ``peer_reduce``: the rows are
    one rank's shard and the kernel all-reduces dweight / dbias over the registered peer buffers
    (``fuser_b200.parallel.PeerExchange``); every rank gets the full gradients."""
    return _norm_bwd(L.OP_LAYERNORM_BWD, dy, x, [mean, invstd], weight, 0, output_mask, partials_f32, variant,
                     cluster_size, peer_reduce)


def rms_norm_bwd(dy, x, stat, weight=None, output_mask=(True, True), *, stat_is_rms=False, centered=False,
                 half_projection=False, partials_f32=False, peer_reduce=False, variant=L.VARIANT_AUTO, cluster_size=0):
    """dx, dweight. ``centered`` keeps the -sum(g) term of the C++ rms_norm_backward builder;
    ``half_projection`` is thunder_rms_norm_backward as written (normalization.cpp:487-535: the projection
    term carries a factor 1/2)."""
    flags = (L.F_STAT_IS_RMS if stat_is_rms else 0) | (L.F_RMS_BWD_CENTERED if centered else 0)
    flags |= L.F_RMS_BWD_HALF_PROJ if half_projection else 0
    return _norm_bwd(L.OP_RMSNORM_BWD, dy, x, [stat], weight, flags, output_mask, partials_f32, variant,
                     cluster_size, peer_reduce)


def peer_combine(k: int, n: int, dtype: torch.dtype, anchor: torch.Tensor):
    """Finish the oldest outstanding publish (``peer_reduce="publish"``): K tensors [N] in ``dtype`` holding
    the sums over all ranks. ``anchor`` is any tensor on the device (the exchange buffer)."""
    outs = [torch.empty((n,), dtype=dtype, device=anchor.device) for _ in range(k)]
    _run(L.OP_PEER_COMBINE, 0, 0.0, L.VARIANT_AUTO, 0, [anchor], outs)
    return outs


def finalize_partials(sums: torch.Tensor, dtype: torch.dtype):
    """fp32 [K,N] column sums (after the allreduce) -> K tensors [N] in ``dtype``."""
    k, n = sums.shape
    outs = [torch.empty((n,), dtype=dtype, device=sums.device) for _ in range(k)]
    _run(L.OP_FINALIZE_PARTIALS, 0, 0.0, L.VARIANT_AUTO, 0, [sums], outs)
    return outs


# ---------------------------------------------------------------------------
# stand-alone statistics (variance_mean / Welford, csrc/ops/normalization.cpp:93-158)
# ---------------------------------------------------------------------------
def var_mean(x, correction: int = 1, *, want_var=True, want_mean=True, variant=L.VARIANT_AUTO, cluster_size=0):
    """var[...,1], mean[...,1] over the last axis (fp32), var = M2 / max(N - correction, 0) as ops.var_mean
    (python/python_direct/ops.cpp:1824-1835; at::var_mean semantics). Runs on the LayerNorm forward kernels
    without their output pass (``F_NO_Y``)."""
    if correction < 0:
        raise ValueError(f"correction must be non-negative, but got {correction}")
    flags = L.F_NO_Y | L.F_STAT_VAR | (L.F_OUT_MEAN if want_mean else 0) | (L.F_OUT_INVSTD if want_var else 0)
    outs = iter(_run(L.OP_LAYERNORM_FWD, flags, 0.0, variant, cluster_size, [x], correction=int(correction)))
    mean = next(outs) if want_mean else None
    var = next(outs) if want_var else None
    return var, mean


def welford(x, *, variant=L.VARIANT_AUTO, cluster_size=0):
    """avg[...,1], var_sum[...,1] (= M2, the sum of squared deviations) over the last axis, fp32 — the tensor
    outputs of ops.welford (python/python_direct/ops.cpp:1857-1863); the count is the extent itself."""
    flags = L.F_NO_Y | L.F_STAT_M2 | L.F_OUT_MEAN | L.F_OUT_INVSTD
    avg, m2 = _run(L.OP_LAYERNORM_FWD, flags, 0.0, variant, cluster_size, [x])
    return avg, m2


# ---------------------------------------------------------------------------
# plain reductions (the Reduction scheduler's segments)
# ---------------------------------------------------------------------------
def reduce_inner(x, kind: str = "sum", *, variant=L.VARIANT_AUTO):
    """[..., N] -> [...]: sum | max (NaN-propagating amax) | fmax (NaN-ignoring)."""
    flags = {"sum": 0, "max": L.F_REDUCE_MAX, "fmax": L.F_REDUCE_FMAX}[kind]
    (r,) = _run(L.OP_REDUCE_INNER, flags, 0.0, variant, 0, [x])
    return r


def reduce_outer(x, *, partials_f32=False, variant=L.VARIANT_AUTO):
    """[M, N] -> [N] sum over rows."""
    flags = L.F_PARTIALS_F32 if partials_f32 else 0
    (r,) = _run(L.OP_REDUCE_OUTER, flags, 0.0, variant, 0, [x])
    return r


# ---------------------------------------------------------------------------
# fused dropout: norm(residual + dropout(x)) and its backward
# (benchmarks/python/test_dropout_{layernorm,rmsnorm}_{fwd,bwd}.py)
# ---------------------------------------------------------------------------
def philox_state(seed: int, offset: int, device) -> torch.Tensor:
    """{seed, offset} of the Philox stream as the two-element int64 device tensor the kernels read (the
    reference passes the same pair as kernel arguments: csrc/runtime/executor.cpp computeArgs' philox seed /
    offset taken from the ATen generator)."""
    to_i64 = lambda v: v - (1 << 64) if v >= (1 << 63) else v  # noqa: E731
    return torch.tensor([to_i64(seed & ((1 << 64) - 1)), to_i64(offset & ((1 << 64) - 1))], dtype=torch.int64, device=device)


def _keep(p: float, keep) -> float:
    """The keep probability as the fusion's own double (1 - p as the definition computed it) when given."""
    return float(keep) if keep is not None else 1.0 - float(p)


def dropout_layer_norm_fwd(x, residual, weight, bias, eps: float, p: float, rng: torch.Tensor, *, want_sum=False,
                           keep=None, variant=L.VARIANT_AUTO, cluster_size=0):
    """y, mean, invstd, mask (and the sum with ``want_sum``) of layer_norm(residual + dropout(x, p)); ``rng`` is
    :func:`philox_state`. mask is the bool tensor the backward takes."""
    flags = (L.F_HAS_WEIGHT if weight is not None else 0) | (L.F_HAS_BIAS if bias is not None else 0)
    flags |= L.F_OUT_MEAN | L.F_OUT_INVSTD | L.F_ADD_RESIDUAL | L.F_DROPOUT | (L.F_OUT_SUM if want_sum else 0)
    ins = [x, residual] + [t for t in (weight, bias) if t is not None] + [rng]
    outs = iter(_run(L.OP_LAYERNORM_FWD, flags, eps, variant, cluster_size, ins, dropout_keep=_keep(p, keep)))
    y = next(outs)
    s = next(outs) if want_sum else None
    mean, invstd, mask = next(outs), next(outs), next(outs)
    return (y, mean, invstd, mask, s) if want_sum else (y, mean, invstd, mask)


def dropout_rms_norm_fwd(x, residual, weight, eps: float, p: float, rng: torch.Tensor, *, stat_is_rms=True,
                         keep=None, variant=L.VARIANT_AUTO, cluster_size=0):
    """y, stat, mask of rms_norm(residual + dropout(x, p)); stat = rms (the benchmark's spelling) or invstd."""
    flags = (L.F_HAS_WEIGHT if weight is not None else 0) | (L.F_STAT_IS_RMS if stat_is_rms else 0)
    flags |= L.F_ADD_RESIDUAL | L.F_DROPOUT
    ins = [x, residual] + ([weight] if weight is not None else []) + [rng]
    y, stat, mask = _run(L.OP_RMSNORM_FWD, flags, eps, variant, cluster_size, ins, dropout_keep=_keep(p, keep))
    return y, stat, mask


def dropout_layer_norm_bwd(dy, x, residual, mask, mean, invstd, weight, p: float,
                           output_mask=(True, True, True), *, keep=None, variant=L.VARIANT_AUTO, cluster_size=0):
    """dx (grad of the dropped branch), dresidual, dweight, dbias of layer_norm(residual + dropout(x, p))."""
    want_dx, want_dw, want_db = output_mask
    flags = (L.F_HAS_WEIGHT if weight is not None else 0) | L.F_DROPOUT
    flags |= (L.F_OUT_DX if want_dx else 0) | (L.F_OUT_DW if want_dw else 0) | (L.F_OUT_DB if want_db else 0)
    ins = [dy, x, mean, invstd] + ([weight] if weight is not None else []) + [residual, mask]
    outs = iter(_run(L.OP_LAYERNORM_BWD, flags, 0.0, variant, cluster_size, ins, dropout_keep=_keep(p, keep)))
    dx = next(outs) if want_dx else None
    dres = next(outs) if want_dx else None
    dw = next(outs) if want_dw else None
    db = next(outs) if want_db else None
    return dx, dres, dw, db


def dropout_rms_norm_bwd(dy, x, residual, mask, stat, weight, p: float, *, stat_is_rms=True,
                         output_mask=(True, True), keep=None, variant=L.VARIANT_AUTO, cluster_size=0):
    """dx, dresidual, dweight of rms_norm(residual + dropout(x, p)) (textbook form, as the benchmark traces it)."""
    want_dx, want_dw = output_mask
    flags = (L.F_HAS_WEIGHT if weight is not None else 0) | (L.F_STAT_IS_RMS if stat_is_rms else 0) | L.F_DROPOUT
    flags |= (L.F_OUT_DX if want_dx else 0) | (L.F_OUT_DW if want_dw else 0)
    ins = [dy, x, stat] + ([weight] if weight is not None else []) + [residual, mask]
    outs = iter(_run(L.OP_RMSNORM_BWD, flags, 0.0, variant, cluster_size, ins, dropout_keep=_keep(p, keep)))
    dx = next(outs) if want_dx else None
    dres = next(outs) if want_dx else None
    dw = next(outs) if want_dw else None
    return dx, dres, dw


# ---------------------------------------------------------------------------
# cross-entropy over vocabulary-size rows: fused log-softmax + NLL, and its backward
# (benchmarks/python/cross_entropy_loss.py: F.cross_entropy(logits.view(-1, V), labels.view(-1)))
# ---------------------------------------------------------------------------
def nll_softmax_fwd(logits, target, ignore_index: int = -100, *, variant=L.VARIANT_AUTO, cluster_size=0):
    """Per-row loss = logsumexp(x) - x[target] (0 where target == ignore_index) and the saved log-sum-exp, both fp32
    of shape ``logits.shape[:-1]``: one read of the logits, the log-probabilities are never materialised."""
    loss, lse = _run(L.OP_REDUCE_INNER, L.F_REDUCE_NLL, 0.0, variant, cluster_size, [logits, target],
                     ignore_index=ignore_index)
    return loss, lse


def nll_softmax_bwd(logits, target, lse, grad_scale, ignore_index: int = -100, *, variant=L.VARIANT_AUTO):
    """dx = (softmax(x) - onehot(target)) * grad_scale (a device fp32 scalar: dloss / count), zero rows where the
    target is ignored."""
    (dx,) = _run(L.OP_NLL_SOFTMAX_BWD, 0, 0.0, variant, 0, [logits, target, lse, grad_scale], ignore_index=ignore_index)
    return dx


def cross_entropy_fwd(logits, target, ignore_index: int = -100, reduction: str = "mean", **kw):
    """F.cross_entropy(logits, target, ignore_index=..., reduction=...) -> (loss, lse, count). The per-row losses come
    from :func:`nll_softmax_fwd`; their sum is one more row reduction of this library (fp32, [1, M])."""
    rows, lse = nll_softmax_fwd(logits, target, ignore_index, **kw)
    if reduction == "none":
        return rows, lse, None
    total = reduce_inner(rows.reshape(1, -1), "sum").reshape(())
    count = (target != ignore_index).sum()
    if reduction == "sum":
        return total, lse, count
    if reduction != "mean":
        raise ValueError(f"reduction {reduction!r}")
    return total / count.to(total.dtype), lse, count


def cross_entropy_bwd(logits, target, lse, count=None, grad_output=None, ignore_index: int = -100, **kw):
    """Gradient of :func:`cross_entropy_fwd` w.r.t. the logits: ``count`` given = mean reduction (g = grad / count),
    absent = sum reduction."""
    g = torch.ones((), dtype=torch.float32, device=logits.device) if grad_output is None else grad_output.to(torch.float32)
    if count is not None:
        g = g / count.to(torch.float32)
    return nll_softmax_bwd(logits, target, lse, g.reshape(1).contiguous(), ignore_index, **kw)


# ---------------------------------------------------------------------------
# instance norm (training statistics): csrc/ops/normalization.cpp:790-950 / :952-1060,
# benchmarks/python/test_instancenorm_{fwd,bwd}.py
# ---------------------------------------------------------------------------
def instance_norm_fwd(x, weight=None, bias=None, eps: float = 1e-5, *, channels_last=False, variant=L.VARIANT_AUTO):
    """y, mean[B, C], invstd[B, C] of instance_norm(x) with the statistics of each (sample, channel) plane.

    channels first ([B, C, spatial...]): ONE launch — a LayerNorm over the rows of the [B*C, spatial] view whose
    affine pair is the row's channel's (``B200NORM_F_ROW_AFFINE``). channels last ([B, spatial..., C]): per sample an
    outer normalization of its [spatial, C] plane (``B200NORM_F_OUTER``), one launch per sample."""
    if channels_last:
        ys, means, invstds = [], [], []
        for n in range(x.shape[0]):
            y, mean, invstd = outer_norm_fwd(x[n], weight, bias, eps, variant=variant)
            ys.append(y)
            means.append(mean)
            invstds.append(invstd)
        return torch.stack(ys), torch.stack(means), torch.stack(invstds)
    c = x.shape[1]
    flags = (L.F_HAS_WEIGHT if weight is not None else 0) | (L.F_HAS_BIAS if bias is not None else 0)
    flags |= L.F_OUT_MEAN | L.F_OUT_INVSTD | L.F_ROW_AFFINE
    ins = [x.contiguous()] + [t for t in (weight, bias) if t is not None]
    y, mean, invstd = _run(L.OP_LAYERNORM_FWD, flags, eps, variant, 0, ins, ignore_index=c)
    return y, mean, invstd


def instance_norm_bwd(dy, x, mean, invstd, weight=None, output_mask=(True, True, True), *, channels_last=False,
                      variant=L.VARIANT_AUTO):
    """dx, dweight[C], dbias[C] of :func:`instance_norm_fwd`. channels first: one launch for dx and the per-plane sums
    of dy * x_hat and dy, then one column sum over the batch ([B, C] -> [C]) for each requested parameter gradient."""
    want_dx, want_dw, want_db = output_mask
    if channels_last:
        dxs, dws, dbs = [], [], []
        for n in range(x.shape[0]):
            dx, dw, db = outer_norm_bwd(dy[n], x[n], mean[n], invstd[n], weight, output_mask, variant=variant)
            dxs.append(dx)
            dws.append(dw)
            dbs.append(db)
        tot = lambda parts: reduce_outer(torch.stack(parts).float()).to(x.dtype) if parts[0] is not None else None  # noqa: E731
        return (torch.stack(dxs) if want_dx else None), tot(dws), tot(dbs)
    c = x.shape[1]
    flags = (L.F_HAS_WEIGHT if weight is not None else 0) | L.F_ROW_AFFINE
    flags |= (L.F_OUT_DX if want_dx else 0) | (L.F_OUT_DW if want_dw else 0) | (L.F_OUT_DB if want_db else 0)
    ins = [dy.contiguous(), x.contiguous(), mean.contiguous(), invstd.contiguous()] + ([weight] if weight is not None else [])
    outs = iter(_run(L.OP_LAYERNORM_BWD, flags, 0.0, variant, 0, ins, ignore_index=c))
    dx = next(outs) if want_dx else None
    dw = reduce_outer(next(outs)).to(x.dtype) if want_dw else None  # [B, C] fp32 -> [C]
    db = reduce_outer(next(outs)).to(x.dtype) if want_db else None
    return dx, dw, db

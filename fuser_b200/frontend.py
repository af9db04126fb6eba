"""``FusionDefinition`` — the Python entry point, call-compatible with ``nvfuser_direct`` for
the normalization hot path.

Mirrors python/nvfuser_direct/__init__.py:121-376 (context manager, define_tensor,
define_scalar, add_output, execute) and :539-645 (validate). ``execute`` plays the role
of ``FusionExecutorCache::runFusionWithInputs`` (csrc/runtime/fusion_executor_cache.cpp:46-108):
the first call classifies the recorded fusion (``matcher.match_fusion`` — what the segmenter +
``NormExecutor::supported`` decide in the reference) and caches the decision; every call binds
the inputs by role, launches the hand-written kernel through the C ABI and returns outputs
in ``fusion.outputs()`` order. A fusion that is not a normalization / softmax segment raises
``UnsupportedSegment`` — in a build of the reference with this library plugged into
``ExecutorDispatch`` that is the signal to fall through to the stock ``KernelExecutor``
(see INTEGRATION.md); stand-alone there is nothing to fall back to, by design.
"""
from __future__ import annotations

import os
from typing import Optional, Sequence

import torch

from . import _lib as L
from . import ops as K
from .ir import DataType, Ops, Scalar, Tensor, UnionFind, nvfuser_dtype_to_torch, torch_dtype_to_nvfuser_dtype
from .matcher import Match, Unmatchable, match_fusion

__all__ = ["FusionDefinition", "DataType", "torch_dtype_to_nvfuser_dtype", "UnsupportedSegment"]

UnsupportedSegment = L.UnsupportedSegment

# validator tolerances: csrc/validator_utils.h:28-46 (data) / validator_utils.cpp:25-163 (rule)
_TOL_FLOAT = [(4, 1.68222e-06), (8, 2.23704e-06), (16, 2.95788e-06), (32, 4.4778e-06), (64, 6.75395e-06),
              (128, 8.57934e-06), (256, 1.30594e-05), (512, 2.19122e-05), (1024, 3.3451e-05), (2048, 5.78476e-05),
              (4096, 0.000108292), (8192, 0.00012207), (16384, 0.000136882), (32768, 0.000248561),
              (65536, 0.000407594), (131072, 0.000500901), (262144, 0.000923019), (524288, 0.00156909),
              (1048576, 0.00223107), (2097152, 0.00343043)]
_TOL_HALF = [(4, 0.00390625), (8, 0.0078125), (16, 0.0078125), (32, 0.0155334), (64, 0.0156269), (128, 0.0312042),
             (256, 0.0312548), (512, 0.0619979), (1024, 0.0625103), (2048, 0.124686), (4096, 0.12501),
             (8192, 0.24945), (16384, 0.250049), (32768, 0.498946), (65536, 0.500071), (131072, 0.985087),
             (262144, 1.00006), (524288, 1.99234), (1048576, 2.00032), (2097152, 3.99073)]


def get_tolerance(dtype: torch.dtype, reduction_size: int) -> tuple[float, float]:
    """(atol, rtol) as ``getTolerance`` computes them (csrc/validator_utils.cpp:25-163)."""
    if not dtype.is_floating_point:
        return 0.0, 0.0
    table, scale = (_TOL_FLOAT, 1.0) if dtype in (torch.float32, torch.float64) else (
        _TOL_HALF, 10.0 if dtype == torch.bfloat16 else 1.0)
    if reduction_size <= 1:
        return table[0][1], table[1][1]
    entry = 0
    while entry < len(table) and table[entry][0] < reduction_size:
        entry += 1
    atol = table[entry + 1][1] if entry + 1 < len(table) else table[-1][1] * 2.0
    return atol * scale, atol * 0.01 * scale


class FusionDefinition:
    """Define a fusion with ``with FusionDefinition() as fd: ...`` and run it with ``fd.execute``.

    Supported surface (what the hot path's fusions use): ``define_tensor``, ``define_scalar``,
    ``add_output``, ``fd.ops.{cast, add, sub, mul, div, neg, pow, exp, log, sqrt, rsqrt, reciprocal,
    sum, max, var_mean, welford, broadcast, broadcast_in_dim, uniform, lt}``, ``Tensor.shape()/size()``.
    """

    def __init__(self):
        self._vals: list = []
        self._inputs: list = []
        self._outputs: list = []
        self._uf = UnionFind()
        self.ops = Ops(self)
        self._match: Optional[Match] = None
        self._match_error: Optional[Exception] = None
        self._last_rng = None
        self._defined = False

    # -- context manager (python/nvfuser_direct/__init__.py:165-204) ------------------------------
    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc, tb):
        self._defined = exc_type is None
        return False

    # -- definition ------------------------------------------------------------------------
    def define_tensor(self, shape: Sequence[int], contiguity=None, dtype: DataType = DataType.Float,
                      is_cpu: bool = False, stride_order: Optional[Sequence[int]] = None) -> Tensor:
        """python/python_direct/ir.cpp:674-731: -1 = dynamic extent, 1 = broadcast extent."""
        if is_cpu:
            raise UnsupportedSegment(L.ERR_UNSUPPORTED, "CPU tensors cannot be fusion inputs of a GPU segment")
        if stride_order and list(stride_order) != sorted(stride_order, reverse=True):
            raise UnsupportedSegment(L.ERR_UNSUPPORTED, "permuted stride orders are not supported")
        t = Tensor(self, "input", [], {"shape": list(shape), "contiguity": contiguity}, dtype, [None] * len(shape))
        t.dims = [None if s == 1 else (t.id, i) for i, s in enumerate(shape)]
        t.attrs["static"] = [s if s > 1 else None for s in shape]
        self._inputs.append(t)
        return t

    def define_scalar(self, value=None, dtype: DataType = DataType.Double) -> Scalar:
        s = Scalar(self, "const" if value is not None else "input", [], {}, dtype, value=value)
        if value is None:
            self._inputs.append(s)
        return s

    def define_vector(self, size: int):
        return [self.define_scalar(None, DataType.Int) for _ in range(size)]

    def from_pytorch(self, tensor: torch.Tensor, static_sizes: bool = False) -> Tensor:
        """python/nvfuser_direct/__init__.py:505-537."""
        shape = [int(s) if (static_sizes or s == 1) else -1 for s in tensor.shape]
        return self.define_tensor(shape, None, torch_dtype_to_nvfuser_dtype(tensor.dtype), tensor.is_cpu)

    def add_output(self, out: Tensor, *args, **kwargs):
        if not isinstance(out, Tensor):
            raise TypeError("only tensor outputs are supported")
        self._outputs.append(out)

    # -- classification (first execute) ------------------------------------------------------
    def segment(self) -> Match:
        """The matched segment (what SegmentCandidateFinder + supported() decide)."""
        if self._match is None and self._match_error is None:
            try:
                m = match_fusion(self)
                self._check_output_dtypes(m)
                self._match = m
            except Unmatchable as e:
                self._match_error = UnsupportedSegment(L.ERR_UNSUPPORTED, f"unsupported segment: {e}")
            if os.environ.get("B200NORM_DEBUG") and self._match is not None:
                m = self._match
                print(f"[b200norm] matched {m.description}: op={L.OP_NAMES[m.op]} flags={m.flags:#x} eps={m.eps} "
                      f"roles={m.roles} outputs={m.outputs}")
        if self._match_error is not None:
            raise self._match_error
        return self._match

    def _check_output_dtypes(self, m: Match):
        for node, role in zip(self._outputs, m.outputs):
            want = DataType.Float if role in ("mean", "invstd", "rms", "var", "m2") else m.io_dtype
            if role == "mask":
                want = DataType.Bool
            if node.dtype != want:
                raise Unmatchable(f"output '{role}' has dtype {node.dtype.name}; the kernels produce {want.name}")

    # -- execution (python/nvfuser_direct/__init__.py:319-376) -----------------------------------
    def execute(self, inputs, *, device=None, save_repro_inputs=False, _enable_options: list[str] = [],
                _disable_options: list[str] = [], _rng=None) -> list[torch.Tensor]:
        if not self._outputs:
            raise NotImplementedError("Fusion is empty!")
        m = self.segment()
        if len(inputs) != len(self._inputs):
            raise ValueError(f"expected {len(self._inputs)} inputs, got {len(inputs)}")
        want_io = nvfuser_dtype_to_torch(m.io_dtype)
        for node, t in zip(self._inputs, inputs):
            if not isinstance(t, torch.Tensor):
                raise TypeError("tensor inputs expected")
            if t.dim() != node.ndim:
                raise ValueError(f"input rank {t.dim()} does not match the definition ({node.ndim})")
            if t.dtype != nvfuser_dtype_to_torch(node.dtype):
                raise ValueError(f"input dtype {t.dtype} does not match the definition ({node.dtype.name})")
            for s, st in zip(t.shape, node.attrs["static"]):
                if st is not None and s != st:
                    raise ValueError(f"input extent {s} does not match the static extent {st}")
        r = {k: inputs[v] for k, v in m.roles.items()}
        kw = {}
        if "tma_inner_persistent" in _disable_options:
            kw["variant"] = L.VARIANT_GENERIC
        res: dict = {}
        want = set(m.outputs)
        if m.flags & L.F_ROW_AFFINE:  # instance norm
            if m.op == L.OP_LAYERNORM_FWD:
                y, mean, invstd = K.instance_norm_fwd(r["x"], r.get("w"), r.get("b"), m.eps, channels_last=m.channels_last, **kw)
                res = {"y": y, "mean": mean, "invstd": invstd}
            else:
                dx, dw, db = K.instance_norm_bwd(r["dy"], r["x"], r["mean"], r["stat"], r.get("w"),
                                                 ("dx" in want, "dw" in want, "db" in want), channels_last=m.channels_last, **kw)
                res = {"dx": dx, "dw": dw, "db": db}
        elif m.flags & L.F_DROPOUT:
            res = self._execute_dropout(m, r, want, kw, _rng)
        elif m.flags & L.F_ADD_RESIDUAL:  # norm(x + residual), optional sum output
            kw.update(residual=r["r"], want_sum="sum" in want, round_sum=bool(m.flags & L.F_ROUND_SUM))
        if m.flags & (L.F_DROPOUT | L.F_ROW_AFFINE):
            pass
        elif m.op == L.OP_LAYERNORM_FWD and m.flags & L.F_NO_Y:  # stand-alone var_mean / welford
            if m.flags & L.F_STAT_M2:
                avg, m2 = K.welford(r["x"], **kw)
                res = {"mean": avg, "m2": m2}
            else:
                var, mean = K.var_mean(r["x"], m.correction, want_var="var" in want, want_mean="mean" in want, **kw)
                res = {"var": var, "mean": mean}
        elif m.op == L.OP_REDUCE_INNER:
            res = {"r": K.reduce_inner(r["x"], "max" if m.flags & L.F_REDUCE_MAX else "sum", **kw)}
        elif m.op == L.OP_REDUCE_OUTER:
            res = {"r": K.reduce_outer(r["x"], **kw)}
        elif m.op == L.OP_LAYERNORM_FWD and m.flags & L.F_OUTER:
            o = K.outer_norm_fwd(r["x"], r.get("w"), r.get("b"), m.eps, want_mean="mean" in want,
                                 want_invstd="invstd" in want, **kw)
            res = {"y": o[0], "mean": o[1], "invstd": o[2]}
        elif m.op == L.OP_LAYERNORM_FWD:
            o = K.layer_norm_fwd(r["x"], r.get("w"), r.get("b"), m.eps, want_mean="mean" in want,
                                 want_invstd="invstd" in want, **kw)
            res = {"y": o[0], "mean": o[1], "invstd": o[2], "sum": o[3] if len(o) > 3 else None}
        elif m.op == L.OP_RMSNORM_FWD:
            o = K.rms_norm_fwd(r["x"], r.get("w"), m.eps, stat_is_rms=bool(m.flags & L.F_STAT_IS_RMS), **kw)
            res = {"y": o[0], "rms": o[1], "invstd": o[1], "sum": o[2] if len(o) > 2 else None}
        elif m.op in (L.OP_SOFTMAX_FWD, L.OP_LOGSOFTMAX_FWD):
            res = {"y": K.softmax_fwd(r["x"], log=m.op == L.OP_LOGSOFTMAX_FWD, dim=0 if m.flags & L.F_OUTER else -1, **kw)}
        elif m.op in (L.OP_SOFTMAX_BWD, L.OP_LOGSOFTMAX_BWD):
            res = {"dx": K.softmax_bwd(r["dy"], r["y"], log=m.op == L.OP_LOGSOFTMAX_BWD,
                                       round_rowsum=bool(m.flags & L.F_ROUND_ROWSUM),
                                       dim=0 if m.flags & L.F_OUTER else -1, **kw)}
        elif m.op == L.OP_LAYERNORM_BWD and m.flags & L.F_OUTER:
            mask = ("dx" in want, "dw" in want, "db" in want)
            dx, dw, db = K.outer_norm_bwd(r["dy"], r["x"], r["mean"], r["stat"], r.get("w"), mask, **kw)
            res = {"dx": dx, "dw": dw, "db": db}
        elif m.op == L.OP_LAYERNORM_BWD:
            mask = ("dx" in want, "dw" in want, "db" in want)
            dx, dw, db = K.layer_norm_bwd(r["dy"], r["x"], r["mean"], r["stat"], r.get("w"), mask, **kw)
            res = {"dx": dx, "dw": dw, "db": db}
        elif m.op == L.OP_RMSNORM_BWD:
            mask = ("dx" in want, "dw" in want)
            dx, dw = K.rms_norm_bwd(r["dy"], r["x"], r["stat"], r.get("w"), mask,
                                    stat_is_rms=bool(m.flags & L.F_STAT_IS_RMS),
                                    centered=bool(m.flags & L.F_RMS_BWD_CENTERED),
                                    half_projection=bool(m.flags & L.F_RMS_BWD_HALF_PROJ), **kw)
            res = {"dx": dx, "dw": dw}
        outs = []
        for node, role in zip(self._outputs, m.outputs):
            t = res[role]
            if role in ("mean", "invstd", "rms", "var", "m2") and node.ndim == t.dim() - 1:
                t = t.squeeze(-1)
            outs.append(t)
        return outs

    def last_rng(self):
        """(seed, offset) handed to the kernels by the last execute() of a fusion with uniform()."""
        return self._last_rng

    def _execute_dropout(self, m: Match, r: dict, want: set, kw: dict, rng):
        """norm(residual + dropout(x)) and its backward. Forward: the Philox arguments are the reference's — seed of
        the default CUDA generator, offset = its offset / 4 (+ the RNG op's index, 0 here), and the generator moves on
        by 4 per launch (csrc/rng.cpp:55-67,73-81)."""
        keep = m.dropout_keep
        if m.op in (L.OP_LAYERNORM_FWD, L.OP_RMSNORM_FWD):
            dev = r["x"].device
            if rng is None:
                gen = torch.cuda.default_generators[dev.index if dev.index is not None else torch.cuda.current_device()]
                off = int(gen.get_offset())
                rng = (int(gen.initial_seed()), off // 4)
                gen.set_offset(off + 4)
            self._last_rng = (int(rng[0]), int(rng[1]))
            state = K.philox_state(self._last_rng[0], self._last_rng[1], dev)
            if m.op == L.OP_LAYERNORM_FWD:
                o = K.dropout_layer_norm_fwd(r["x"], r["r"], r.get("w"), r.get("b"), m.eps, 0.0, state, keep=keep,
                                             want_sum="sum" in want, **kw)
                return {"y": o[0], "mean": o[1], "invstd": o[2], "mask": o[3], "sum": o[4] if len(o) > 4 else None}
            y, stat, mask = K.dropout_rms_norm_fwd(r["x"], r["r"], r.get("w"), m.eps, 0.0, state, keep=keep,
                                                   stat_is_rms=bool(m.flags & L.F_STAT_IS_RMS), **kw)
            return {"y": y, "rms": stat, "invstd": stat, "mask": mask}
        want_dx = "dx" in want or "dres" in want
        if m.op == L.OP_LAYERNORM_BWD:
            dx, dres, dw, db = K.dropout_layer_norm_bwd(r["dy"], r["x"], r["r"], r["mask"], r["mean"], r["stat"], r.get("w"), 0.0,
                                                        (want_dx, "dw" in want, "db" in want), keep=keep, **kw)
            return {"dx": dx, "dres": dres, "dw": dw, "db": db}
        dx, dres, dw = K.dropout_rms_norm_bwd(r["dy"], r["x"], r["r"], r["mask"], r["stat"], r.get("w"), 0.0,
                                              stat_is_rms=bool(m.flags & L.F_STAT_IS_RMS),
                                              output_mask=(want_dx, "dw" in want), keep=keep, **kw)
        return {"dx": dx, "dres": dres, "dw": dw}

    # -- validation (python/nvfuser_direct/__init__.py:539-645) ---------------------------------
    def reduction_sizes(self, inputs) -> list[int]:
        """Accumulated reduction size per output, as ReductionSizeMapper walks the fusion
        (csrc/validator_utils.cpp:165-254): own reduction extent x max over the inputs."""
        extent = {}
        for node, t in zip(self._inputs, inputs):
            if isinstance(node, Tensor):
                for ax, d in enumerate(node.dims):
                    if d is not None:
                        extent[self._uf.find(d)] = int(t.shape[ax])
        red: dict[int, int] = {}
        for v in self._vals:
            if not isinstance(v, Tensor):
                continue
            inp = max([red.get(a.id, 1) for a in v.args if isinstance(a, Tensor)] or [1])
            own = 1
            if v.op in ("sum", "max", "mean", "var", "m2"):
                src = v.args[0]
                for i in v.attrs["dims"]:
                    if src.dims[i] is not None:
                        own *= extent.get(self._uf.find(src.dims[i]), 1)
            red[v.id] = own * inp
        return [red.get(o.id, 1) for o in self._outputs]

    def validate(self, inputs, reference_outputs: Optional[list] = None, device=None, **kwargs):
        """Run the fusion and compare with reference outputs under the validator's tolerances.
        Auto-inferred references need an ATen evaluation of the fusion, which is test infrastructure
        (``oracle.fusion_eval``), so ``reference_outputs`` is required here."""
        outs = self.execute(inputs, device=device, **kwargs)
        if reference_outputs is None:
            raise NotImplementedError(
                "validate() without reference outputs evaluates the fusion with ATen; use "
                "oracle.fusion_eval.evaluate(fd, inputs) to produce them (test infrastructure)")
        assert len(outs) == len(reference_outputs), f"Expected {len(outs)} reference outputs for validation."
        for out, ref, rs in zip(outs, reference_outputs, self.reduction_sizes(inputs)):
            assert ref.shape == out.shape, "Mismatch in reference and fusion output dimensions"
            if out.is_floating_point():
                atol, rtol = get_tolerance(out.dtype, rs)
                ref_c = ref.to(device=out.device, dtype=out.dtype)
                assert torch.allclose(out, ref_c, atol=atol, rtol=rtol, equal_nan=True), (
                    f"Max error: {(out.double() - ref_c.double()).abs().max().item()}, "
                    f"Absolute tolerance: {atol}, Relative tolerance: {rtol}")
            else:
                assert torch.equal(out, ref.to(out.device))
        return outs

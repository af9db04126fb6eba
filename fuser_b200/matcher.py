"""Segment matcher: recognise LayerNorm / RMSNorm / softmax forward+backward fusions.

In the reference, the segmenter hands every segment to a priority list of schedulers
(csrc/scheduler/registry.cpp:192-227) and InnerPersistent / InnerOuterPersistent accept
*any* fusion with the right reduction topology, then generate code for it. This library
has a fixed set of hand-written kernels instead, so the question asked here is narrower:
"is this fusion, as written, mathematically one of the normalization ops the kernels
implement?" — and it must be answered for every spelling the frontends produce (the C++
builders of csrc/ops/normalization.cpp, the Thunder traces in benchmarks/python/*.py with
their 40-op backward, pow(x,2) vs mul(x,x), rsqrt vs 1/sqrt, reciprocal(size) vs div ...).

The matcher therefore does not compare graph shapes. It canonicalises every fusion output
to a polynomial normal form over atoms (inputs, reductions of monomials, max / exp / log of
canonical sub-expressions, non-integer powers of multi-term expressions) and compares it
with the normal form of candidate op definitions instantiated on the fusion's own inputs.
Two spellings match iff their normal forms agree term by term. Casts are value-preserving
in this algebra (they decide io dtypes, not the match); parity against the fusion as written
is then what the validator tolerances check.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field
from fractions import Fraction
from typing import Optional

from . import _lib as L
from .ir import REDUCED, DataType, Scalar, Tensor, Val

# ---------------------------------------------------------------------------
# polynomial normal form
# ---------------------------------------------------------------------------
_COEF_DIGITS = 9


def _round(c: float) -> float:
    return float(f"{c:.{_COEF_DIGITS}e}")


class Poly:
    """sum_i coef_i * prod_j atom_ij ^ exp_ij   (exponents are Fractions)."""

    __slots__ = ("terms",)

    # A term disappears only when its contributions CANCEL (|c| <= _CANCEL * sum |contribution|), never because
    # it is small next to the polynomial's other terms: an epsilon of 1e-12 next to a variance is a term like
    # any other, and two fusions that differ in it are different segments.
    _CANCEL = 1e-12

    def __init__(self, terms: Optional[dict] = None):
        self.terms: dict[tuple, float] = terms or {}

    # -- constructors
    @staticmethod
    def const(c: float) -> "Poly":
        return Poly({(): float(c)}) if c != 0 else Poly()

    @staticmethod
    def atom(a: tuple, e: Fraction = Fraction(1)) -> "Poly":
        return Poly({((a, e),): 1.0})

    # -- helpers
    def clean(self) -> "Poly":
        self.terms = {k: c for k, c in self.terms.items() if c != 0.0 and c == c}
        return self

    @staticmethod
    def _accumulate(pairs) -> "Poly":
        """Sum (key, coefficient) contributions; a key whose contributions cancel is dropped."""
        t: dict = {}
        mag: dict = {}
        for k, c in pairs:
            t[k] = t.get(k, 0.0) + c
            mag[k] = mag.get(k, 0.0) + abs(c)
        return Poly({k: c for k, c in t.items() if c == c and abs(c) > Poly._CANCEL * mag[k]})

    def key(self) -> tuple:
        return tuple(sorted((k, _round(c)) for k, c in self.terms.items()))

    def is_monomial(self) -> bool:
        return len(self.terms) == 1

    # -- arithmetic
    def __add__(self, o: "Poly") -> "Poly":
        return Poly._accumulate(itertools.chain(self.terms.items(), o.terms.items()))

    def scale(self, c: float) -> "Poly":
        return Poly({k: v * c for k, v in self.terms.items()}).clean()

    def __neg__(self) -> "Poly":
        return self.scale(-1.0)

    def __sub__(self, o: "Poly") -> "Poly":
        return self + (-o)

    @staticmethod
    def _mul_keys(a: tuple, b: tuple) -> tuple:
        d: dict = {}
        for atom, e in itertools.chain(a, b):
            d[atom] = d.get(atom, Fraction(0)) + e
        return tuple(sorted(((atom, e) for atom, e in d.items() if e != 0), key=repr))

    def __mul__(self, o: "Poly") -> "Poly":
        return Poly._accumulate((Poly._mul_keys(ka, kb), ca * cb)
                                for ka, ca in self.terms.items() for kb, cb in o.terms.items())

    def pow(self, e: Fraction) -> "Poly":
        if e == 1:
            return self
        if e == 0:
            return Poly.const(1.0)
        if not self.terms:
            return Poly()
        if self.is_monomial():
            (k, c), = self.terms.items()
            if c > 0 or e.denominator == 1:
                nk = tuple(sorted(((a, x * e) for a, x in k if x * e != 0), key=repr))
                return Poly({nk: float(c) ** float(e)})
        if e.denominator == 1 and 0 < e <= 4:
            r = self
            for _ in range(int(e) - 1):
                r = r * self
            return r
        # non-integer / negative power of a multi-term expression: opaque atom
        return Poly.atom(("expr", self.key()), e)


# ---------------------------------------------------------------------------
# canonicaliser: DAG -> Poly (with axis bookkeeping)
# ---------------------------------------------------------------------------
class Unmatchable(Exception):
    """The fusion uses an op or shape pattern outside the normalization hot path."""


class Canon:
    def __init__(self, fd):
        self.fd = fd
        self.uf = fd._uf
        self.cache: dict[int, Poly] = {}
        self.axes_of_key: dict[tuple, frozenset] = {}

    def cls(self, dim):
        return None if dim is None else self.uf.find(dim)

    # axes an atom's value depends on
    def atom_axes(self, atom: tuple) -> frozenset:
        kind = atom[0]
        if kind == "in":
            t = self.fd._vals[atom[1]]
            return frozenset(self.cls(d) for d in t.dims if d is not None)
        if kind == "dim":
            return frozenset()
        if kind in ("sum", "max"):
            return self.axes_of_key[atom[2]] - atom[1]
        if kind in ("exp", "log", "expr"):
            return self.axes_of_key[atom[1]]
        if kind == "keep":  # the dropout mask uniform() < keep: spans the axes of the uniform tensor
            t = self.fd._vals[atom[1]]
            return frozenset(self.cls(d) for d in t.dims if d is not None)
        raise KeyError(kind)

    def poly_axes(self, p: Poly) -> frozenset:
        ax: set = set()
        for k in p.terms:
            for atom, _ in k:
                ax |= self.atom_axes(atom)
        return frozenset(ax)

    def register(self, p: Poly) -> tuple:
        key = p.key()
        self.axes_of_key[key] = self.poly_axes(p)
        return key

    def reduce_sum(self, p: Poly, tensor_dims: list, red_idx: list) -> Poly:
        """sum over the logical axes red_idx of a tensor whose shape classes are tensor_dims."""
        red_real = [self.cls(tensor_dims[i]) for i in red_idx if tensor_dims[i] is not None]
        if not red_real:
            return p  # reduction over broadcast (size-1) axes only: identity
        red = frozenset(red_real)
        out = Poly()
        for k, c in p.terms.items():
            dep, indep = [], []
            for atom, e in k:
                (dep if self.atom_axes(atom) & red else indep).append((atom, e))
            dep_axes = frozenset().union(*[self.atom_axes(a) for a, _ in dep]) if dep else frozenset()
            term = Poly({tuple(indep): c})
            # axes being reduced that nothing in the term depends on contribute their extent
            for ax in red - dep_axes:
                term = term * Poly.atom(("dim", ax))
            inner_red = red & dep_axes
            if inner_red:
                inner = Poly({tuple(sorted(dep, key=repr)): 1.0})
                ik = self.register(inner)
                term = term * Poly.atom(("sum", inner_red, ik))
            out = out + term
        return out

    def scalar(self, s: Scalar) -> Poly:
        if s.id in self.cache:
            return self.cache[s.id]
        if s.value is not None and s.dim is None:
            r = Poly.const(float(s.value))
        elif s.op == "size":
            r = Poly.atom(("dim", self.cls(s.dim)))
        elif s.op == "cast":
            r = self.scalar(s.args[0])
        elif s.op in ("neg", "reciprocal", "sqrt", "rsqrt"):
            a = self.scalar(s.args[0])
            r = {"neg": lambda: -a, "reciprocal": lambda: a.pow(Fraction(-1)),
                 "sqrt": lambda: a.pow(Fraction(1, 2)), "rsqrt": lambda: a.pow(Fraction(-1, 2))}[s.op]()
        elif s.op in ("add", "sub", "mul", "div", "pow"):
            r = self.binary(s.op, self.scalar(s.args[0]), self.scalar(s.args[1]), s.args[1])
        else:
            raise Unmatchable(f"scalar op {s.op}")
        self.cache[s.id] = r
        return r

    def binary(self, op: str, a: Poly, b: Poly, b_val: Val) -> Poly:
        if op == "add":
            return a + b
        if op == "sub":
            return a - b
        if op == "mul":
            return a * b
        if op == "div":
            return a * b.pow(Fraction(-1))
        if op == "pow":
            if isinstance(b_val, Scalar) and b_val.value is not None:
                e = Fraction(float(b_val.value)).limit_denominator(64)
                return a.pow(e)
            raise Unmatchable("pow with a non-constant exponent")
        raise Unmatchable(op)

    def value(self, v: Val) -> Poly:
        if isinstance(v, Scalar):
            return self.scalar(v)
        if v.id in self.cache:
            return self.cache[v.id]
        r = self._tensor(v)
        self.cache[v.id] = r
        return r

    def _tensor(self, t: Tensor) -> Poly:
        op = t.op
        if op == "input":
            return Poly.atom(("in", t.id))
        if op in ("cast", "broadcast", "broadcast_in_dim"):
            return self.value(t.args[0])
        if op == "lt":
            # only as the dropout mask: uniform(0, 1) < constant (benchmarks/python/test_dropout_layernorm_fwd.py:38-41)
            u, c = t.args
            if (isinstance(u, Tensor) and u.op == "uniform" and isinstance(c, Scalar) and c.value is not None
                    and isinstance(u.args[0], Scalar) and isinstance(u.args[1], Scalar)
                    and u.args[0].value == 0.0 and u.args[1].value == 1.0 and u.dtype == DataType.Float):
                return Poly.atom(("keep", u.id))
            raise Unmatchable("lt outside a dropout mask")
        if op == "neg":
            return -self.value(t.args[0])
        if op == "reciprocal":
            return self.value(t.args[0]).pow(Fraction(-1))
        if op == "sqrt":
            return self.value(t.args[0]).pow(Fraction(1, 2))
        if op == "rsqrt":
            return self.value(t.args[0]).pow(Fraction(-1, 2))
        if op in ("exp", "log"):
            inner = self.value(t.args[0])
            return Poly.atom((op, self.register(inner)))
        if op in ("add", "sub", "mul", "div", "pow"):
            return self.binary(op, self.value(t.args[0]), self.value(t.args[1]), t.args[1])
        if op in ("sum", "max", "mean", "var", "m2"):
            src: Tensor = t.args[0]
            x = self.value(src)
            idx = t.attrs["dims"]
            if op == "sum":
                return self.reduce_sum(x, src.dims, idx)
            red = frozenset(self.cls(src.dims[i]) for i in idx if src.dims[i] is not None)
            if op == "max":
                if not red:
                    return x
                return Poly.atom(("max", red, self.register(x)))
            n = Poly.const(1.0)
            for ax in red:
                n = n * Poly.atom(("dim", ax))
            mean = self.reduce_sum(x, src.dims, idx) * n.pow(Fraction(-1))
            if op == "mean":
                return mean
            d = x - mean
            m2 = self.reduce_sum(d * d, src.dims, idx)
            if op == "m2":
                return m2
            corr = int(t.attrs.get("correction", 0))
            denom = n - Poly.const(float(corr)) if corr else n
            return m2 * denom.pow(Fraction(-1))
        raise Unmatchable(f"op {op} is not on the normalization hot path")


# ---------------------------------------------------------------------------
# candidate op definitions, written against the recorder's own ops
# (math of csrc/ops/normalization.cpp, cited per function)
# ---------------------------------------------------------------------------
def _bcast_last(ops, t: Tensor, like: Tensor) -> Tensor:
    """[N] -> [.., N] aligned with the last axis of `like`."""
    return ops.broadcast_in_dim(t, shape=like.shape(), broadcast_dims=[like.ndim - 1])


def _bcast_rows(ops, t: Tensor, like: Tensor) -> Tensor:
    """per-row statistic ([..] or [.., 1]) -> broadcastable against `like`."""
    if t.ndim == like.ndim:
        return t
    return ops.broadcast_in_dim(t, shape=like.shape()[:-1] + [1], broadcast_dims=list(range(like.ndim - 1)))


def def_layer_norm_fwd(ops, x, w, b, eps):
    """normalization.cpp:304-342"""
    last = [x.ndim - 1]
    var, mean = ops.var_mean(x, dims=last, correction=0, keepdim=True)
    invstd = ops.rsqrt(ops.add(var, eps))
    y = ops.mul(ops.sub(x, mean), invstd)
    if w is not None:
        y = ops.mul(y, _bcast_last(ops, w, x))
    if b is not None:
        y = ops.add(y, _bcast_last(ops, b, x))
    return {"y": y, "mean": mean, "invstd": invstd}


def def_outer_norm_fwd(ops, x, w, b, eps):
    """benchmarks/python/normalization.py:13-95 (batch norm, training, channels last): statistics over every
    axis but the last, affine per channel — an OuterPersistent segment."""
    outer = list(range(x.ndim - 1))
    var, mean = ops.var_mean(x, dims=outer, correction=0, keepdim=False)
    invstd = ops.rsqrt(ops.add(var, eps))
    y = ops.mul(ops.sub(x, _bcast_last(ops, mean, x)), _bcast_last(ops, invstd, x))
    if w is not None:
        y = ops.mul(y, _bcast_last(ops, w, x))
    if b is not None:
        y = ops.add(y, _bcast_last(ops, b, x))
    return {"y": y, "mean": mean, "invstd": invstd}


def def_outer_norm_bwd(ops, dy, x, mean, invstd, w):
    """benchmarks/python/normalization.py:98-190 (batch norm backward, training, channels last)."""
    outer = list(range(x.ndim - 1))
    num = x.size(0)
    for ax in outer[1:]:
        num = ops.mul(num, x.size(ax))
    norm = ops.reciprocal(num)
    xc = ops.sub(x, _bcast_last(ops, mean, x))
    grad_sum = ops.sum(dy, dims=outer, keepdim=False)
    dot_p = ops.sum(ops.mul(dy, xc), dims=outer, keepdim=False)
    grad_mean = _bcast_last(ops, ops.mul(grad_sum, norm), x)
    proj_scale = _bcast_last(ops, ops.mul(ops.mul(dot_p, norm), ops.mul(invstd, invstd)), x)
    grad_scale = _bcast_last(ops, invstd, x)
    if w is not None:
        grad_scale = ops.mul(_bcast_last(ops, w, x), grad_scale)
    dx = ops.mul(ops.sub(ops.sub(dy, ops.mul(proj_scale, xc)), grad_mean), grad_scale)
    return {"dx": dx, "dw": ops.mul(dot_p, invstd), "db": grad_sum}


def def_instance_norm_fwd(ops, x, w, b, eps, c_axis):
    """csrc/ops/normalization.cpp:790-950 / benchmarks/python/normalization.py:13-95 with norm = "instance_norm":
    statistics over the spatial axes of every (sample, channel) plane, affine per channel."""
    nd = x.ndim
    red = [a for a in range(1, nd) if a != c_axis]
    var, mean = ops.var_mean(x, dims=red, correction=0, keepdim=False)
    invstd = ops.rsqrt(ops.add(var, eps))
    smask = [a in red for a in range(nd)]
    cmask = [a != c_axis for a in range(nd)]
    y = ops.mul(ops.sub(x, ops.broadcast(mean, smask)), ops.broadcast(invstd, smask))
    if w is not None:
        y = ops.mul(y, ops.broadcast(w, cmask))
    if b is not None:
        y = ops.add(y, ops.broadcast(b, cmask))
    return {"y": y, "mean": mean, "invstd": invstd}


def def_instance_norm_bwd(ops, dy, x, mean, invstd, w, c_axis):
    """csrc/ops/normalization.cpp:952-1060 / benchmarks/python/normalization.py:98-190 with norm = "instance_norm"."""
    nd = x.ndim
    red = [a for a in range(1, nd) if a != c_axis]
    smask = [a in red for a in range(nd)]
    cmask = [a != c_axis for a in range(nd)]
    num = x.size(red[0])
    for ax in red[1:]:
        num = ops.mul(num, x.size(ax))
    norm = ops.reciprocal(num)
    xc = ops.sub(x, ops.broadcast(mean, smask))
    grad_sum = ops.sum(dy, dims=red, keepdim=False)
    dot_p = ops.sum(ops.mul(dy, xc), dims=red, keepdim=False)
    grad_mean = ops.broadcast(ops.mul(grad_sum, norm), smask)
    proj_scale = ops.broadcast(ops.mul(ops.mul(dot_p, norm), ops.mul(invstd, invstd)), smask)
    grad_scale = ops.broadcast(invstd, smask)
    if w is not None:
        grad_scale = ops.mul(ops.broadcast(w, cmask), grad_scale)
    dx = ops.mul(ops.sub(ops.sub(dy, ops.mul(proj_scale, xc)), grad_mean), grad_scale)
    return {"dx": dx, "dw": ops.sum(ops.mul(dot_p, invstd), dims=[0], keepdim=False), "db": ops.sum(grad_sum, dims=[0], keepdim=False)}


def def_rms_norm_fwd(ops, x, w, eps):
    """normalization.cpp:352-380 (+ the rms statistic of benchmarks/python/test_rmsnorm_fwd.py:26-45)"""
    last = [x.ndim - 1]
    ms = ops.mul(ops.sum(ops.mul(x, x), dims=last, keepdim=True), ops.reciprocal(x.size(-1)))
    var_eps = ops.add(ms, eps)
    invstd = ops.rsqrt(var_eps)
    y = ops.mul(x, invstd)
    if w is not None:
        y = ops.mul(y, _bcast_last(ops, w, x))
    return {"y": y, "invstd": invstd, "rms": ops.sqrt(var_eps)}


def def_reduce(ops, x, kind, inner):
    """plain reductions (scheduler/reduction.cpp:198-360): sum / max over the last axis, sum over all the others"""
    dims = [x.ndim - 1] if inner else list(range(x.ndim - 1))
    return {"r": (ops.sum if kind == "sum" else ops.max)(x, dims=dims, keepdim=False)}


def def_var_mean(ops, x, correction):
    """variance_mean / Welford over the last axis (normalization.cpp:93-158, python_direct/ops.cpp:1824-1863)"""
    last = [x.ndim - 1]
    var, mean = ops.var_mean(x, dims=last, correction=correction, keepdim=False)
    _, m2, _ = ops.welford(x, dims=last)
    return {"var": var, "mean": mean, "m2": m2}


def def_softmax_fwd(ops, x, log, axis=-1):
    """normalization.cpp:169-189 / :212-231 (axis 0: the outer-persistent case)"""
    last = [axis % x.ndim]
    d = ops.sub(x, ops.max(x, dims=last, keepdim=True))
    e = ops.exp(d)
    s = ops.sum(e, dims=last, keepdim=True)
    return {"y": ops.sub(d, ops.log(s)) if log else ops.mul(e, ops.reciprocal(s))}


def def_softmax_bwd(ops, dy, y, log, axis=-1):
    """normalization.cpp:191-210 / :233-249 (axis 0: the outer-persistent case)"""
    last = [axis % y.ndim]
    if log:
        return {"dx": ops.sub(dy, ops.mul(ops.exp(y), ops.sum(dy, dims=last, keepdim=True)))}
    g = ops.mul(dy, y)
    return {"dx": ops.sub(g, ops.mul(y, ops.sum(g, dims=last, keepdim=True)))}


def def_norm_bwd(ops, dy, x, mean, stat, w, stat_is_rms, centered, half_projection=False):
    """normalization.cpp:382-436 (LayerNorm), :438-485 (RMSNorm, centred), the Python trace
    benchmarks/python/test_rmsnorm_bwd.py:20-76 (RMSNorm, textbook) and :487-535 (thunder_rms_norm_backward
    as the C++ builder writes it: the projection term carries a factor 1/2)."""
    last = [x.ndim - 1]
    outer = list(range(x.ndim - 1))
    st = _bcast_rows(ops, stat, x)
    invstd = ops.reciprocal(st) if stat_is_rms else st
    xc = ops.sub(x, _bcast_rows(ops, mean, x)) if mean is not None else x
    xh = ops.mul(xc, invstd)
    g = ops.mul(dy, _bcast_last(ops, w, x)) if w is not None else dy
    n = x.size(-1)
    inner = ops.mul(n, g)
    if centered:
        inner = ops.sub(inner, ops.sum(g, dims=last, keepdim=True))
    proj = ops.mul(xh, ops.sum(ops.mul(g, xh), dims=last, keepdim=True))
    if half_projection:
        proj = ops.mul(proj, 0.5)
    inner = ops.sub(inner, proj)
    dx = ops.mul(ops.mul(ops.reciprocal(n), invstd), inner)
    out = {"dx": dx, "db": ops.sum(dy, dims=outer)}
    out["dw"] = ops.sum(ops.mul(dy, xh), dims=outer)
    return out


# ---------------------------------------------------------------------------
# the match result
# ---------------------------------------------------------------------------
@dataclass
class Match:
    op: int                    # L.OP_*
    flags: int                 # L.F_* (without output-selection bits for fwd stats)
    eps: float
    io_dtype: DataType
    roles: dict                # role -> index into fusion inputs (x, dy, y, w, b, mean, stat)
    outputs: list              # per fusion output: role name (y, mean, invstd, rms, dx, dw, db)
    norm_ndim: int = 1
    correction: int = 0        # var_mean (F_STAT_VAR)
    dropout_keep: float = 1.0  # F_DROPOUT: the keep probability 1 - p (the fusion's double scalar)
    channels_last: bool = False  # F_ROW_AFFINE (instance norm): [B, spatial..., C] instead of [B, C, spatial...]
    description: str = ""
    notes: list = field(default_factory=list)


def _tensor_inputs(fd):
    """Tensor inputs that reach an output. An input whose consumers are all dead code (the running statistics of
    benchmarks/python/normalization.py, whose updates are never added as outputs) is not part of the segment —
    the reference's dead-code removal drops those expressions before segmentation."""
    live, stack = set(), list(fd._outputs)
    while stack:
        v = stack.pop()
        if v.id in live:
            continue
        live.add(v.id)
        stack.extend(a for a in v.args if isinstance(a, Val))
    return [(i, v) for i, v in enumerate(fd._inputs) if isinstance(v, Tensor) and v.id in live]


def _eps_candidates(fd) -> list[float]:
    """Scalar constants added to a tensor: the epsilon is one of them."""
    out = []
    for v in fd._vals:
        if isinstance(v, Tensor) and v.op in ("add", "sub"):
            for a in v.args:
                if isinstance(a, Scalar) and a.value is not None and a.dim is None:
                    val = float(a.value)
                    if 0 < abs(val) < 1.0 and val not in out:
                        out.append(val)
    return out or [0.0]


def _has_intermediate_round(fd) -> bool:
    """A cast to a reduced-precision dtype whose result is consumed by more arithmetic
    (the row-sum round trip of benchmarks/python/test_softmax_bwd.py:49-60)."""
    outs = {v.id for v in fd._outputs}
    consumed = set()
    for v in fd._vals:
        for a in v.args:
            consumed.add(a.id)
    for v in fd._vals:
        if isinstance(v, Tensor) and v.op == "cast" and v.dtype in REDUCED and v.id in consumed and v.id not in outs:
            src = v.args[0]
            if isinstance(src, Tensor) and src.op != "input":
                return True
    return False


def _sum_is_rounded(fd) -> bool:
    """The residual sum is cast to a reduced-precision dtype and that rounded value feeds further
    arithmetic (a bf16 ``hidden + residual`` traced ahead of the norm, torch_ops.py:76-83)."""
    consumed = set()
    for v in fd._vals:
        for a in v.args:
            consumed.add(a.id)
    for v in fd._vals:
        if isinstance(v, Tensor) and v.op == "cast" and v.dtype in REDUCED and v.id in consumed:
            src = v.args[0]
            if isinstance(src, Tensor) and src.op == "add":
                return True
    return False


def _match_instance_norm(fd, tin, mats, cls, out_keys):
    """Instance norm (training statistics) forward / backward on [B, C, spatial...] or [B, spatial..., C]: its
    parameters are indexed by an axis that is neither a row nor the column of the [rows, cols] view the other segments
    use, so it is recognised before that classification. Returns a Match or None."""
    nd = max(t.ndim for _, t in mats)
    mats = [(i, t) for i, t in mats if t.ndim == nd]  # the saved [B, C] statistics of the backward are rank 2
    x_like = mats[0][1]
    if nd < 3 or len(mats) not in (1, 2):
        return None
    if any(tuple(cls(d) for d in t.dims) != tuple(cls(d) for d in x_like.dims) for _, t in mats):
        return None
    io = x_like.dtype
    if io not in (DataType.Float, DataType.Half, DataType.BFloat16) or any(t.dtype != io for _, t in mats):
        return None
    ops = fd.ops
    others = [(i, t) for i, t in tin if (i, t) not in mats]

    def T(pair):
        t = pair[1]
        return ops.cast(t, DataType.Float) if t.dtype in REDUCED else t

    def attempt(build):
        snap = fd._uf.snapshot()
        nvals = len(fd._vals)
        try:
            cand = build()
            c2 = Canon(fd)
            keys = {name: c2.value(t).key() for name, t in cand.items() if t is not None}
            if len({c2.cls(d) for d in x_like.dims}) != nd:  # the candidate merged two extents: wrong role guess
                return None
            c3 = Canon(fd)
            mine = [c3.value(o).key() for o in fd._outputs]
        except (Unmatchable, ValueError):
            return None
        finally:
            del fd._vals[nvals:]
            fd._uf.restore(snap)
        names = []
        for k in mine:
            hit = [n for n, kk in keys.items() if kk == k]
            if not hit:
                return None
            names.append(hit[0])
        return names

    for c_axis in (1, nd - 1):
        ccls, bcls = cls(x_like.dims[c_axis]), cls(x_like.dims[0])
        chan = [(i, t) for i, t in others if t.ndim == 1 and cls(t.dims[0]) == ccls and t.dtype == io]
        plane = [(i, t) for i, t in others if t.ndim == 2 and (cls(t.dims[0]), cls(t.dims[1])) == (bcls, ccls)
                 and t.dtype == DataType.Float]
        if len(chan) + len(plane) != len(others):
            continue
        if len(mats) == 1 and not plane and len(chan) <= 2:
            x = mats[0]
            combos = [(None, None)] if not chan else [(chan[0], None)] if len(chan) == 1 else \
                [(chan[0], chan[1]), (chan[1], chan[0])]
            for eps in _eps_candidates(fd):
                e = ops._scalar(float(eps))
                for w, b in combos:
                    names = attempt(lambda: def_instance_norm_fwd(ops, T(x), T(w) if w else None, T(b) if b else None, e, c_axis))
                    if names and "y" in names:
                        roles = {"x": x[0]}
                        if w:
                            roles["w"] = w[0]
                        if b:
                            roles["b"] = b[0]
                        flags = L.F_ROW_AFFINE | (L.F_HAS_WEIGHT if w else 0) | (L.F_HAS_BIAS if b else 0)
                        return Match(L.OP_LAYERNORM_FWD, flags, eps, io, roles, names, channels_last=c_axis != 1,
                                     description="instance_norm forward" + (" (channels last)" if c_axis != 1 else ""))
        if len(mats) == 2 and len(plane) == 2 and len(chan) <= 1:
            w = chan[0] if chan else None
            for dy_, x_ in ((mats[0], mats[1]), (mats[1], mats[0])):
                for mean, st in ((plane[0], plane[1]), (plane[1], plane[0])):
                    names = attempt(lambda: def_instance_norm_bwd(ops, T(dy_), T(x_), mean[1], st[1], T(w) if w else None, c_axis))
                    if names and "dx" in names:
                        roles = {"dy": dy_[0], "x": x_[0], "mean": mean[0], "stat": st[0]}
                        if w:
                            roles["w"] = w[0]
                        return Match(L.OP_LAYERNORM_BWD, L.F_ROW_AFFINE | (L.F_HAS_WEIGHT if w else 0), 0.0, io, roles, names,
                                     channels_last=c_axis != 1,
                                     description="instance_norm backward" + (" (channels last)" if c_axis != 1 else ""))
    return None


def match_fusion(fd) -> Match:
    """Classify the recorded fusion or raise Unmatchable."""
    if not fd._outputs:
        raise Unmatchable("fusion has no outputs")
    tin = _tensor_inputs(fd)
    if any(isinstance(v, Scalar) for v in fd._inputs):
        raise Unmatchable("symbolic scalar inputs are not supported by the normalization kernels")

    canon = Canon(fd)
    try:
        out_keys = [canon.value(o).key() for o in fd._outputs]
    except Unmatchable:
        raise

    # classify inputs by signature once the extent classes are final
    cls = canon.cls
    mats = [(i, t) for i, t in tin if t.ndim >= 2 and all(d is not None for d in t.dims)]
    if not mats:
        raise Unmatchable("no [rows, cols] input")
    inst = _match_instance_norm(fd, tin, mats, cls, out_keys)
    if inst is not None:
        return inst
    x_like = mats[0][1]
    nd = x_like.ndim
    col_cls = cls(x_like.dims[-1])
    row_cls = tuple(cls(d) for d in x_like.dims[:-1])
    if col_cls in row_cls:
        raise Unmatchable("square symbolic shapes: rows and cols were unified")
    mats = [(i, t) for i, t in mats if t.ndim == nd and tuple(cls(d) for d in t.dims) == row_cls + (col_cls,)]
    masks = [(i, t) for i, t in mats if t.dtype == DataType.Bool]  # a saved dropout mask (backward)
    mats = [(i, t) for i, t in mats if t.dtype != DataType.Bool]
    if not mats:
        raise Unmatchable("no [rows, cols] input")
    vecs = [(i, t) for i, t in tin if t.ndim == 1 and cls(t.dims[0]) == col_cls]
    io0 = mats[0][1].dtype
    # fp32 per-COLUMN statistics next to reduced-precision activations (saved mean / invstd of the outer
    # normalization backward); with fp32 io they cannot be told from weights here and stay in `vecs`
    cstats = [(i, t) for i, t in vecs if t.dtype == DataType.Float and io0 != DataType.Float]
    vecs = [(i, t) for i, t in vecs if (i, t) not in cstats]
    stats = [(i, t) for i, t in tin
             if (t.ndim == nd - 1 and tuple(cls(d) for d in t.dims) == row_cls)
             or (t.ndim == nd and t.dims[-1] is None and tuple(cls(d) for d in t.dims[:-1]) == row_cls)]
    used = {i for i, _ in mats} | {i for i, _ in vecs} | {i for i, _ in stats} | {i for i, _ in cstats} | {i for i, _ in masks}
    if len(used) != len(tin):
        raise Unmatchable("an input is neither [rows, cols], [cols] nor a per-row statistic")
    io = mats[0][1].dtype
    if io not in (DataType.Float, DataType.Half, DataType.BFloat16):
        raise Unmatchable(f"io dtype {io.name}")
    if any(t.dtype != io for _, t in mats + vecs):
        raise Unmatchable("activations / weights of mixed dtypes")
    if any(t.dtype != DataType.Float for _, t in stats):
        raise Unmatchable("per-row statistics must be Float")

    ops = fd.ops
    eps_list = _eps_candidates(fd)

    def try_candidate(build, roles_by_name: dict):
        """Instantiate a candidate definition on the fusion's inputs; compare normal forms."""
        snap = fd._uf.snapshot()
        nvals = len(fd._vals)
        try:
            cand = build()
            c2 = Canon(fd)
            keys = {name: c2.value(t).key() for name, t in cand.items() if t is not None}
            # extent classes must not have been merged by the candidate (wrong role guess)
            if c2.cls(x_like.dims[-1]) in tuple(c2.cls(d) for d in x_like.dims[:-1]):
                return None
            c3 = Canon(fd)
            mine = [c3.value(o).key() for o in fd._outputs]
        except (Unmatchable, ValueError):
            return None
        finally:
            del fd._vals[nvals:]
            fd._uf.restore(snap)
        names = []
        for k in mine:
            hit = [n for n, kk in keys.items() if kk == k]
            if not hit:
                return None
            names.append(hit[0])
        return names

    n_m, n_v, n_s = len(mats), len(vecs), len(stats)

    def T(pair):  # candidate definitions work on Float values, like every frontend spelling does
        t = pair[1]
        return ops.cast(t, DataType.Float) if t.dtype in REDUCED else t

    I = lambda pair: pair[0]  # noqa: E731

    # ---- forward ops: one matrix input ----
    if n_m == 1 and n_s == 0:
        x = mats[0]
        if n_v == 0:
            for axis in (-1, 0):  # inner-persistent | outer-persistent (reduction over dim 0)
                for log in (False, True):
                    names = try_candidate(lambda: def_softmax_fwd(ops, T(x), log, axis), {})
                    if names:
                        return Match(L.OP_LOGSOFTMAX_FWD if log else L.OP_SOFTMAX_FWD, L.F_OUTER if axis == 0 else 0,
                                     0.0, io, {"x": I(x)}, names,
                                     description=("log_softmax" if log else "softmax") + (" over dim 0" if axis == 0 else ""))
            # plain reductions and stand-alone statistics (Reduction segments / variance_mean, Welford)
            for kind, inner in (("sum", True), ("max", True), ("sum", False)):
                names = try_candidate(lambda: def_reduce(ops, T(x), kind, inner), {})
                if names:
                    if inner:
                        return Match(L.OP_REDUCE_INNER, L.F_REDUCE_MAX if kind == "max" else 0, 0.0, io, {"x": I(x)}, names,
                                     description=f"{kind} over the last axis")
                    return Match(L.OP_REDUCE_OUTER, 0, 0.0, io, {"x": I(x)}, names, description="sum over the outer axes")
            corrections = sorted({int(v.attrs.get("correction", 0)) for v in fd._vals
                                  if isinstance(v, Tensor) and v.op == "var"} | {0})
            for corr in corrections:
                names = try_candidate(lambda: def_var_mean(ops, T(x), corr), {})
                if names:
                    if "var" in names and "m2" in names:
                        raise Unmatchable("both the variance and the sum of squared deviations requested")
                    flags = L.F_NO_Y | (L.F_STAT_M2 if "m2" in names else L.F_STAT_VAR)
                    return Match(L.OP_LAYERNORM_FWD, flags, 0.0, io, {"x": I(x)}, names, correction=corr,
                                 description="welford" if "m2" in names else f"var_mean (correction {corr})")
        for eps in eps_list:
            e = ops._scalar(float(eps))
            # RMSNorm: 0 or 1 vectors
            if n_v <= 1:
                w = vecs[0] if n_v == 1 else None
                names = try_candidate(lambda: def_rms_norm_fwd(ops, T(x), T(w) if w else None, e), {})
                if names:
                    flags = (L.F_HAS_WEIGHT if w else 0) | (L.F_STAT_IS_RMS if "rms" in names else 0)
                    if "rms" in names and "invstd" in names:
                        raise Unmatchable("both rms and invstd requested")
                    roles = {"x": I(x)}
                    if w:
                        roles["w"] = I(w)
                    return Match(L.OP_RMSNORM_FWD, flags, eps, io, roles, names, description="rms_norm")
            # LayerNorm: 0, 1 (weight only) or 2 vectors (weight, bias in either order)
            combos = [(None, None)]
            if n_v == 1:
                combos = [(vecs[0], None)]
            elif n_v == 2:
                combos = [(vecs[0], vecs[1]), (vecs[1], vecs[0])]
            for w, b in combos:
                names = try_candidate(
                    lambda: def_layer_norm_fwd(ops, T(x), T(w) if w else None, T(b) if b else None, e), {})
                if names:
                    flags = (L.F_HAS_WEIGHT if w else 0) | (L.F_HAS_BIAS if b else 0)
                    roles = {"x": I(x)}
                    if w:
                        roles["w"] = I(w)
                    if b:
                        roles["b"] = I(b)
                    return Match(L.OP_LAYERNORM_FWD, flags, eps, io, roles, names, description="layer_norm")
                # statistics over the outer axes, affine per column: the batch-norm-shaped forward
                names = try_candidate(
                    lambda: def_outer_norm_fwd(ops, T(x), T(w) if w else None, T(b) if b else None, e), {})
                if names and "y" in names:
                    flags = L.F_OUTER | (L.F_HAS_WEIGHT if w else 0) | (L.F_HAS_BIAS if b else 0)
                    roles = {"x": I(x)}
                    if w:
                        roles["w"] = I(w)
                    if b:
                        roles["b"] = I(b)
                    return Match(L.OP_LAYERNORM_FWD, flags, eps, io, roles, names,
                                 description="batch_norm forward (statistics over the outer axes, channels last)")
        raise Unmatchable("single-input fusion is not rms_norm / layer_norm / softmax")

    # ---- fused dropout (benchmarks/python/test_dropout_{layernorm,rmsnorm}_{fwd,bwd}.py) ----
    keep_nodes = [v for v in fd._vals if isinstance(v, Tensor) and v.op == "lt" and isinstance(v.args[0], Tensor)
                  and v.args[0].op == "uniform"]
    if keep_nodes:
        if len(keep_nodes) != 1 or masks or n_m != 2 or n_s != 0 or n_v > 2:
            raise Unmatchable("dropout: one mask over norm(residual + dropout(x)) expected")
        knode = keep_nodes[0]
        keep = float(knode.args[1].value)
        if not (0.0 < keep <= 1.0):
            raise Unmatchable(f"dropout keep probability {keep}")
        dflags = L.F_ADD_RESIDUAL | L.F_DROPOUT
        for x, res in ((mats[0], mats[1]), (mats[1], mats[0])):
            def with_drop(build):
                s_ = ops.add(T(res), ops.mul(ops.mul(T(x), ops.cast(knode, DataType.Float)), 1.0 / keep))
                out = build(s_)
                out["sum"] = s_
                out["mask"] = knode
                return out

            for eps in eps_list:
                e = ops._scalar(float(eps))
                if n_v <= 1:
                    w = vecs[0] if n_v == 1 else None
                    names = try_candidate(lambda: with_drop(lambda s_: def_rms_norm_fwd(ops, s_, T(w) if w else None, e)), {})
                    if names and "y" in names:
                        if "rms" in names and "invstd" in names:
                            raise Unmatchable("both rms and invstd requested")
                        flags = dflags | (L.F_HAS_WEIGHT if w else 0) | (L.F_STAT_IS_RMS if "rms" in names else 0)
                        flags |= L.F_OUT_SUM if "sum" in names else 0
                        roles = {"x": I(x), "r": I(res)}
                        if w:
                            roles["w"] = I(w)
                        return Match(L.OP_RMSNORM_FWD, flags, eps, io, roles, names, dropout_keep=keep,
                                     description="rms_norm(residual + dropout(x))")
                combos = [(None, None)]
                if n_v == 1:
                    combos = [(vecs[0], None)]
                elif n_v == 2:
                    combos = [(vecs[0], vecs[1]), (vecs[1], vecs[0])]
                for w, b in combos:
                    names = try_candidate(lambda: with_drop(
                        lambda s_: def_layer_norm_fwd(ops, s_, T(w) if w else None, T(b) if b else None, e)), {})
                    if names and "y" in names:
                        flags = dflags | (L.F_HAS_WEIGHT if w else 0) | (L.F_HAS_BIAS if b else 0)
                        flags |= L.F_OUT_SUM if "sum" in names else 0
                        roles = {"x": I(x), "r": I(res)}
                        if w:
                            roles["w"] = I(w)
                        if b:
                            roles["b"] = I(b)
                        return Match(L.OP_LAYERNORM_FWD, flags, eps, io, roles, names, dropout_keep=keep,
                                     description="layer_norm(residual + dropout(x))")
        raise Unmatchable("a dropout mask outside norm(residual + dropout(x))")
    if masks:
        # backward: dy, x (dropped branch), residual + the saved mask; the scale 1 / keep is one of the constants
        if len(masks) != 1 or n_m != 3 or n_v > 1 or n_s not in (1, 2):
            raise Unmatchable("dropout backward: dy, x, residual, one mask and the saved statistics expected")
        mk = masks[0]
        w = vecs[0] if n_v == 1 else None
        scales = sorted({float(v.value) for v in fd._vals if isinstance(v, Scalar) and v.value is not None and v.dim is None
                         and isinstance(v.value, float) and 1.0 <= float(v.value) <= 1e6})
        for scale in scales:
            for dy_, x_, r_ in itertools.permutations(mats):
                def with_drop_bwd(mean, st, stat_is_rms, centered, half):
                    mf = ops.cast(mk[1], DataType.Float)
                    xe = ops.add(T(r_), ops.mul(ops.mul(T(x_), mf), scale))
                    out = def_norm_bwd(ops, T(dy_), xe, mean, st, T(w) if w else None, stat_is_rms, centered, half)
                    out["dres"] = out["dx"]
                    out["dx"] = ops.mul(ops.mul(out["dres"], scale), mf)
                    return out

                if n_s == 1:
                    for stat_is_rms, centered in ((True, False), (False, False), (True, True), (False, True)):
                        names = try_candidate(lambda: with_drop_bwd(None, T(stats[0]), stat_is_rms, centered, False), {})
                        if names and "db" not in names and ("dx" in names or "dres" in names):
                            flags = L.F_DROPOUT | (L.F_HAS_WEIGHT if w else 0) | (L.F_STAT_IS_RMS if stat_is_rms else 0)
                            flags |= L.F_RMS_BWD_CENTERED if centered else 0
                            roles = {"dy": I(dy_), "x": I(x_), "r": I(r_), "mask": I(mk), "stat": I(stats[0])}
                            if w:
                                roles["w"] = I(w)
                            return Match(L.OP_RMSNORM_BWD, flags, 0.0, io, roles, names, dropout_keep=1.0 / scale,
                                         description="backward of rms_norm(residual + dropout(x))")
                else:
                    for mean, st in ((stats[0], stats[1]), (stats[1], stats[0])):
                        names = try_candidate(lambda: with_drop_bwd(T(mean), T(st), False, True, False), {})
                        if names and ("dx" in names or "dres" in names):
                            flags = L.F_DROPOUT | (L.F_HAS_WEIGHT if w else 0)
                            roles = {"dy": I(dy_), "x": I(x_), "r": I(r_), "mask": I(mk), "mean": I(mean), "stat": I(st)}
                            if w:
                                roles["w"] = I(w)
                            return Match(L.OP_LAYERNORM_BWD, flags, 0.0, io, roles, names, dropout_keep=1.0 / scale,
                                         description="backward of layer_norm(residual + dropout(x))")
        raise Unmatchable("a Bool matrix input outside a dropout backward")

    # ---- forward norms with a fused residual add: two matrix inputs, no statistics ----
    if n_m == 2 and n_s == 0 and n_v <= 2:
        x, res = mats[0], mats[1]  # the sum is symmetric: either order is the same segment
        rflags = L.F_ADD_RESIDUAL | (L.F_ROUND_SUM if _sum_is_rounded(fd) else 0)
        for eps in eps_list:
            e = ops._scalar(float(eps))

            def with_sum(build):
                s_ = ops.add(T(x), T(res))
                out = build(s_)
                out["sum"] = s_
                return out

            if n_v <= 1:
                w = vecs[0] if n_v == 1 else None
                names = try_candidate(lambda: with_sum(lambda s_: def_rms_norm_fwd(ops, s_, T(w) if w else None, e)), {})
                if names and any(nm in ("y", "invstd", "rms") for nm in names):
                    if "rms" in names and "invstd" in names:
                        raise Unmatchable("both rms and invstd requested")
                    flags = rflags | (L.F_HAS_WEIGHT if w else 0) | (L.F_STAT_IS_RMS if "rms" in names else 0)
                    flags |= L.F_OUT_SUM if "sum" in names else 0
                    roles = {"x": I(x), "r": I(res)}
                    if w:
                        roles["w"] = I(w)
                    return Match(L.OP_RMSNORM_FWD, flags, eps, io, roles, names, description="rms_norm(x + residual)")
            combos = [(None, None)]
            if n_v == 1:
                combos = [(vecs[0], None)]
            elif n_v == 2:
                combos = [(vecs[0], vecs[1]), (vecs[1], vecs[0])]
            for w, b in combos:
                names = try_candidate(lambda: with_sum(
                    lambda s_: def_layer_norm_fwd(ops, s_, T(w) if w else None, T(b) if b else None, e)), {})
                if names and any(nm in ("y", "mean", "invstd") for nm in names):
                    flags = rflags | (L.F_HAS_WEIGHT if w else 0) | (L.F_HAS_BIAS if b else 0)
                    flags |= L.F_OUT_SUM if "sum" in names else 0
                    roles = {"x": I(x), "r": I(res)}
                    if w:
                        roles["w"] = I(w)
                    if b:
                        roles["b"] = I(b)
                    return Match(L.OP_LAYERNORM_FWD, flags, eps, io, roles, names,
                                 description="layer_norm(x + residual)")

    # ---- outer normalization backward: two matrices + per-column mean, invstd (+ weight) ----
    pool = vecs + cstats
    if n_m == 2 and n_s == 0 and len(pool) in (2, 3):
        for a, b_ in ((mats[0], mats[1]), (mats[1], mats[0])):
            for perm in itertools.permutations(pool):
                mean, st = perm[0], perm[1]
                w = perm[2] if len(perm) == 3 else None
                if mean[1].dtype != DataType.Float or st[1].dtype != DataType.Float or (w and w[1].dtype != io):
                    continue
                names = try_candidate(
                    lambda: def_outer_norm_bwd(ops, T(a), T(b_), mean[1], st[1], T(w) if w else None), {})
                if names and "dx" in names:
                    flags = L.F_OUTER | (L.F_HAS_WEIGHT if w else 0)
                    roles = {"dy": I(a), "x": I(b_), "mean": I(mean), "stat": I(st)}
                    if w:
                        roles["w"] = I(w)
                    return Match(L.OP_LAYERNORM_BWD, flags, 0.0, io, roles, names,
                                 description="batch_norm backward (statistics over the outer axes, channels last)")
    if cstats:
        raise Unmatchable("per-column fp32 statistics outside an outer normalization backward")

    # ---- backward ops: two matrix inputs ----
    if n_m == 2:
        round_sum = _has_intermediate_round(fd)
        for a, b_ in ((mats[0], mats[1]), (mats[1], mats[0])):
            if n_s == 0 and n_v == 0:
                for axis in (-1, 0):
                    for log in (False, True):
                        names = try_candidate(lambda: def_softmax_bwd(ops, T(a), T(b_), log, axis), {})
                        if names:
                            flags = (L.F_ROUND_ROWSUM if (round_sum and not log) else 0) | (L.F_OUTER if axis == 0 else 0)
                            return Match(L.OP_LOGSOFTMAX_BWD if log else L.OP_SOFTMAX_BWD, flags, 0.0, io,
                                         {"dy": I(a), "y": I(b_)}, names,
                                         description=("log_softmax_backward" if log else "softmax_backward") +
                                         (" over dim 0" if axis == 0 else ""))
            w = vecs[0] if n_v == 1 else None
            if n_v > 1:
                continue
            if n_s == 1:  # RMSNorm backward
                for stat_is_rms, centered, half in ((False, False, False), (False, True, False), (True, False, False),
                                                    (True, True, False), (True, False, True), (False, False, True)):
                        names = try_candidate(
                            lambda: def_norm_bwd(ops, T(a), T(b_), None, T(stats[0]), T(w) if w else None,
                                                 stat_is_rms, centered, half), {})
                        if names and "db" not in names:
                            flags = (L.F_HAS_WEIGHT if w else 0) | (L.F_STAT_IS_RMS if stat_is_rms else 0)
                            flags |= L.F_RMS_BWD_CENTERED if centered else 0
                            flags |= L.F_RMS_BWD_HALF_PROJ if half else 0
                            roles = {"dy": I(a), "x": I(b_), "stat": I(stats[0])}
                            if w:
                                roles["w"] = I(w)
                            return Match(L.OP_RMSNORM_BWD, flags, 0.0, io, roles, names,
                                         description="thunder_rms_norm_backward (halved projection)" if half
                                         else "rms_norm_backward")
            if n_s == 2:  # LayerNorm backward
                for mean, st in ((stats[0], stats[1]), (stats[1], stats[0])):
                    names = try_candidate(
                        lambda: def_norm_bwd(ops, T(a), T(b_), T(mean), T(st), T(w) if w else None, False, True), {})
                    if names:
                        flags = L.F_HAS_WEIGHT if w else 0
                        roles = {"dy": I(a), "x": I(b_), "mean": I(mean), "stat": I(st)}
                        if w:
                            roles["w"] = I(w)
                        return Match(L.OP_LAYERNORM_BWD, flags, 0.0, io, roles, names,
                                     description="layer_norm_backward")
        raise Unmatchable("two-input fusion is not a normalization / softmax backward")

    raise Unmatchable(f"unsupported input signature: {n_m} matrices, {n_v} vectors, {n_s} row statistics")

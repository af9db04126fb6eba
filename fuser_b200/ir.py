"""Minimal fusion IR: just enough of nvFuser's Val/Expr/TensorView surface to record the
normalization / reduction fusions the Python frontend builds.

Mirrors (API surface only) python/python_direct/ir.cpp:674-731 (define_tensor),
python/python_direct/ops.cpp (unary/binary/reduction/broadcast ops) and the TensorView
helpers ``shape()`` / ``size(i)`` used by the benchmark definitions
(benchmarks/python/test_layernorm_bwd.py:20-101). It is a recorder, not a compiler: there
is no scheduling, lowering or codegen here — the recorded DAG is classified by
``fuser_b200.matcher`` and routed to a hand-written kernel, or rejected.
"""
from __future__ import annotations

import enum
from typing import Optional, Sequence, Union

import torch


class DataType(enum.Enum):
    """Names follow nvfuser_direct.DataType (csrc/type.h PrimDataType)."""

    Double = "double"
    Float = "float"
    Half = "half"
    BFloat16 = "bfloat16"
    Int = "int64"
    Int32 = "int32"
    Bool = "bool"
    Null = "null"


_TORCH_TO_DT = {
    torch.float64: DataType.Double,
    torch.float32: DataType.Float,
    torch.float16: DataType.Half,
    torch.bfloat16: DataType.BFloat16,
    torch.int64: DataType.Int,
    torch.int32: DataType.Int32,
    torch.bool: DataType.Bool,
}
_DT_TO_TORCH = {v: k for k, v in _TORCH_TO_DT.items()}
REDUCED = (DataType.Half, DataType.BFloat16)


def torch_dtype_to_nvfuser_dtype(dtype: torch.dtype) -> DataType:
    """nvfuser_direct.pytorch_utils.torch_dtype_to_nvfuser_dtype."""
    return _TORCH_TO_DT[dtype]


def nvfuser_dtype_to_torch(dtype: DataType) -> torch.dtype:
    return _DT_TO_TORCH[dtype]


class Val:
    """Base of Scalar and Tensor nodes."""

    __slots__ = ("fd", "id", "op", "args", "attrs", "dtype")

    def __init__(self, fd, op: str, args: Sequence["Val"], attrs: dict, dtype: DataType):
        self.fd = fd
        self.id = len(fd._vals)
        self.op = op
        self.args = list(args)
        self.attrs = attrs
        self.dtype = dtype
        fd._vals.append(self)


class Scalar(Val):
    """A scalar Val. ``value`` is the Python constant, or None for a symbolic value
    (an extent of an input tensor, or arithmetic on one)."""

    __slots__ = ("value", "dim")

    def __init__(self, fd, op, args, attrs, dtype, value=None, dim=None):
        super().__init__(fd, op, args, attrs, dtype)
        self.value = value
        self.dim = dim  # (tensor_id, axis) when this scalar is exactly one tensor extent

    def is_symbolic(self) -> bool:
        return self.value is None

    def __repr__(self):
        return f"S{self.id}({self.op}{'' if self.value is None else '=' + str(self.value)})"


class Tensor(Val):
    """A TensorView. ``dims`` holds, per logical axis, either a dim key ``(tensor_id, axis)``
    (resolved through the fusion's union-find to an extent class) or None for a broadcast
    (size-1) axis."""

    __slots__ = ("dims",)

    def __init__(self, fd, op, args, attrs, dtype, dims):
        super().__init__(fd, op, args, attrs, dtype)
        self.dims = list(dims)

    @property
    def ndim(self) -> int:
        return len(self.dims)

    def size(self, axis: int) -> Scalar:
        axis = axis % self.ndim if self.ndim else axis
        d = self.dims[axis]
        if d is None:
            return Scalar(self.fd, "const", [], {}, DataType.Int, value=1)
        return Scalar(self.fd, "size", [self], {"axis": axis}, DataType.Int, value=None, dim=d)

    def shape(self) -> list[Scalar]:
        return [self.size(i) for i in range(self.ndim)]

    def __repr__(self):
        return f"T{self.id}({self.op}, {self.dtype.name}, ndim={self.ndim})"


Operand = Union[Tensor, Scalar, float, int]


class UnionFind:
    def __init__(self):
        self.parent: dict = {}

    def find(self, x):
        p = self.parent.setdefault(x, x)
        while p != self.parent.setdefault(p, p):
            self.parent[x] = self.parent[p]
            x, p = p, self.parent[p]
        return p

    def union(self, a, b):
        ra, rb = self.find(a), self.find(b)
        if ra != rb:
            # keep the smaller key as root for determinism
            if rb < ra:
                ra, rb = rb, ra
            self.parent[rb] = ra

    def snapshot(self) -> dict:
        return dict(self.parent)

    def restore(self, snap: dict):
        self.parent = dict(snap)


def promote(a: DataType, b: DataType) -> DataType:
    order = [DataType.Bool, DataType.Int32, DataType.Int, DataType.Half, DataType.BFloat16, DataType.Float,
             DataType.Double]
    if {a, b} == {DataType.Half, DataType.BFloat16}:
        return DataType.Float
    return a if order.index(a) >= order.index(b) else b


class Ops:
    """``fd.ops`` — the subset of python/python_direct/ops.cpp the hot path's fusions use."""

    def __init__(self, fd):
        self._fd = fd

    # -- helpers -----------------------------------------------------------
    def _scalar(self, v) -> Scalar:
        if isinstance(v, Scalar):
            return v
        if isinstance(v, bool):
            return Scalar(self._fd, "const", [], {}, DataType.Bool, value=v)
        if isinstance(v, int):
            return Scalar(self._fd, "const", [], {}, DataType.Int, value=v)
        if isinstance(v, float):
            return Scalar(self._fd, "const", [], {}, DataType.Double, value=v)
        raise TypeError(f"expected a scalar, got {type(v)}")

    def _unary(self, name: str, a: Operand):
        fd = self._fd
        if isinstance(a, Tensor):
            return Tensor(fd, name, [a], {}, a.dtype, a.dims)
        s = self._scalar(a)
        val = None
        if s.value is not None:
            import math

            x = float(s.value)
            val = {"neg": lambda: -x, "reciprocal": lambda: 1.0 / x, "sqrt": lambda: math.sqrt(x),
                   "rsqrt": lambda: 1.0 / math.sqrt(x), "exp": lambda: math.exp(x), "log": lambda: math.log(x),
                   "abs": lambda: abs(x)}[name]()
        return Scalar(fd, name, [s], {}, DataType.Double, value=val)

    def _binary(self, name: str, a: Operand, b: Operand):
        fd = self._fd
        ta, tb = isinstance(a, Tensor), isinstance(b, Tensor)
        if not ta and not tb:
            sa, sb = self._scalar(a), self._scalar(b)
            val = None
            if sa.value is not None and sb.value is not None:
                x, y = sa.value, sb.value
                val = {"add": lambda: x + y, "sub": lambda: x - y, "mul": lambda: x * y, "div": lambda: x / y,
                       "pow": lambda: x ** y}[name]()
            return Scalar(fd, name, [sa, sb], {}, promote(sa.dtype, sb.dtype), value=val)
        if ta and tb:
            if a.ndim != b.ndim:
                # ranks are aligned the way the reference's binary ops do it (csrc/ops/utils.cpp:22-34, 511-532
                # maybeBroadcast): the lower-rank operand gets leading broadcast axes
                lo, hi = (a, b) if a.ndim < b.ndim else (b, a)
                pad = hi.ndim - lo.ndim
                lo_b = Tensor(fd, "broadcast", [lo], {"mask": [True] * pad + [False] * lo.ndim}, lo.dtype, [None] * pad + lo.dims)
                a, b = (lo_b, b) if a is lo else (a, lo_b)
            dims = []
            for da, db in zip(a.dims, b.dims):
                if da is not None and db is not None:
                    fd._uf.union(da, db)
                dims.append(da if da is not None else db)
            return Tensor(fd, name, [a, b], {}, promote(a.dtype, b.dtype), dims)
        t, s = (a, self._scalar(b)) if ta else (b, self._scalar(a))
        # scalars do not promote floating tensors (type promotion of the reference: csrc/type_promotion.cpp)
        dtype = t.dtype if t.dtype in (DataType.Double, DataType.Float, *REDUCED) else promote(t.dtype, s.dtype)
        args = [a if ta else self._scalar(a), b if tb else self._scalar(b)]
        return Tensor(fd, name, args, {}, dtype, t.dims)

    # -- pointwise ---------------------------------------------------------
    def cast(self, arg: Operand, dtype: DataType):
        if isinstance(arg, Tensor):
            return Tensor(self._fd, "cast", [arg], {"dtype": dtype}, dtype, arg.dims)
        s = self._scalar(arg)
        return Scalar(self._fd, "cast", [s], {"dtype": dtype}, dtype, value=s.value, dim=s.dim)

    def neg(self, a): return self._unary("neg", a)
    def reciprocal(self, a): return self._unary("reciprocal", a)
    def sqrt(self, a): return self._unary("sqrt", a)
    def rsqrt(self, a): return self._unary("rsqrt", a)
    def exp(self, a): return self._unary("exp", a)
    def log(self, a): return self._unary("log", a)
    def abs(self, a): return self._unary("abs", a)
    def add(self, a, b): return self._binary("add", a, b)
    def sub(self, a, b): return self._binary("sub", a, b)
    def mul(self, a, b): return self._binary("mul", a, b)
    def div(self, a, b): return self._binary("div", a, b)
    def truediv(self, a, b): return self._binary("div", a, b)
    def pow(self, a, b): return self._binary("pow", a, b)

    # -- random numbers / comparisons (the dropout prologue) -----------------
    def uniform(self, minval, maxval, shape: Sequence, dtype: DataType = DataType.Float, rng_seed=None, rng_offset=None):
        """ops.uniform (python_direct/ops.cpp random ops): a tensor of ``shape`` drawn from U[minval, maxval) with the
        Philox stream of the run (csrc/rng.cpp). Only the [0, 1) Float flavour feeding a dropout mask is on the hot
        path; the shape entries are extents of input tensors (``T.shape()``)."""
        fd = self._fd
        if rng_seed is not None or rng_offset is not None:
            raise ValueError("uniform: explicit rng_seed / rng_offset are not supported (the stream comes from the CUDA generator)")
        lo, hi = self._scalar(minval), self._scalar(maxval)
        dims = []
        for s_ in shape:
            if isinstance(s_, Scalar) and s_.dim is not None:
                dims.append(s_.dim)
            else:
                raise ValueError("uniform: the shape must be extents of input tensors")
        return Tensor(fd, "uniform", [lo, hi], {}, dtype, dims)

    def lt(self, a, b):
        """Element-wise a < b -> Bool."""
        r = self._binary("lt", a, b)
        if isinstance(r, Tensor):
            r.dtype = DataType.Bool
        return r

    # -- reductions --------------------------------------------------------
    def _reduce(self, name, arg: Tensor, dims, keepdim, dtype):
        if not isinstance(arg, Tensor):
            raise TypeError(f"{name}: expected a tensor")
        if dims is None:
            dims = list(range(arg.ndim))
        if isinstance(dims, int):
            dims = [dims]
        dims = sorted(d % arg.ndim for d in dims)
        out_dims = []
        for i, d in enumerate(arg.dims):
            if i in dims:
                if keepdim:
                    out_dims.append(None)
            else:
                out_dims.append(d)
        odt = arg.dtype if dtype in (None, DataType.Null) else dtype
        return Tensor(self._fd, name, [arg], {"dims": dims, "keepdim": bool(keepdim)}, odt, out_dims)

    def sum(self, arg, dims=None, keepdim=False, dtype=DataType.Null, *, dim=None):
        return self._reduce("sum", arg, dims if dim is None else dim, keepdim, dtype)

    def max(self, arg, dims=None, keepdim=False, dtype=DataType.Null, *, dim=None):
        return self._reduce("max", arg, dims if dim is None else dim, keepdim, dtype)

    def amax(self, arg, dims=None, keepdim=False, dtype=DataType.Null):
        return self._reduce("max", arg, dims, keepdim, dtype)

    def var_mean(self, arg, dims, correction=1, keepdim=False):
        """ops.var_mean (python_direct/ops.cpp:1824-1835) → (var, mean)."""
        if arg.dtype in REDUCED:
            # same user error as csrc/ops/normalization.cpp:101-108
            raise ValueError(f"variance_mean is not supported for {arg.dtype.name} please upcast to float")
        if correction < 0:
            raise ValueError(f"correction must be non-negative, but got {correction}")
        var = self._reduce("var", arg, dims, keepdim, DataType.Null)
        var.attrs["correction"] = int(correction)
        mean = self._reduce("mean", arg, dims, keepdim, DataType.Null)
        return var, mean

    def welford(self, arg, dims):
        """ops.welford (python_direct/ops.cpp:1857-1863) → (avg, var_sum, N)."""
        avg = self._reduce("mean", arg, dims, False, DataType.Null)
        m2 = self._reduce("m2", arg, dims, False, DataType.Null)
        n = self._fd.ops._scalar(1)
        for d in (dims if not isinstance(dims, int) else [dims]):
            n = self.mul(n, arg.size(d))
        return avg, m2, n

    # -- shape ops -----------------------------------------------------------
    def broadcast(self, arg: Tensor, is_broadcast_dim: Sequence[bool]):
        dims, it = [], iter(arg.dims)
        for flag in is_broadcast_dim:
            dims.append(None if flag else next(it))
        if len([f for f in is_broadcast_dim if not f]) != arg.ndim:
            raise ValueError("broadcast: number of False entries must equal the input rank")
        return Tensor(self._fd, "broadcast", [arg], {"mask": list(map(bool, is_broadcast_dim))}, arg.dtype, dims)

    def broadcast_in_dim(self, arg: Tensor, shape: Sequence, broadcast_dims: Sequence[int]):
        """ops.broadcast_in_dim (python_direct/ops.cpp:2421-2551): ``arg`` axis i lands on output
        axis broadcast_dims[i]; every other output axis is new with extent shape[j]."""
        fd = self._fd
        if len(broadcast_dims) != arg.ndim:
            raise ValueError("broadcast_in_dim: broadcast_dims must have one entry per input axis")
        out: list = [None] * len(shape)
        expanded = [False] * len(shape)
        for j, s in enumerate(shape):
            if isinstance(s, Scalar):
                if s.dim is not None:
                    out[j] = s.dim
                    expanded[j] = True
                elif s.value is not None and int(s.value) != 1:
                    raise ValueError("broadcast_in_dim: constant extents other than 1 are not supported")
            elif isinstance(s, int):
                if s != 1:
                    raise ValueError("broadcast_in_dim: constant extents other than 1 are not supported")
            else:
                raise TypeError("broadcast_in_dim: shape entries must be scalars or ints")
        for i, j in enumerate(broadcast_dims):
            d = arg.dims[i]
            if d is not None:
                if out[j] is not None:
                    fd._uf.union(d, out[j])
                out[j] = d
                expanded[j] = False
        return Tensor(fd, "broadcast_in_dim", [arg],
                      {"broadcast_dims": list(broadcast_dims), "rank": len(shape)}, arg.dtype, out)

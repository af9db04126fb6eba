"""Export a recorded ``FusionDefinition`` as the plain-data expression graph of the C++ segment classifier
(``csrc/host/segment_matcher.h``) — the stand-alone counterpart of the adapter that walks ``fusion->exprs()`` in an
nvFuser tree. ``cpp_match(fd)`` runs the C++ classifier on it."""
from __future__ import annotations

from . import _host as H
from .ir import DataType, Scalar, Tensor

# ExprKind of segment_matcher.h
KIND = {"input": 0, "const": 1, "size": 2, "cast": 3, "broadcast": 4, "broadcast_in_dim": 4, "neg": 5, "reciprocal": 6,
        "sqrt": 7, "rsqrt": 8, "exp": 9, "log": 10, "add": 11, "sub": 12, "mul": 13, "div": 14, "pow": 15, "sum": 16,
        "max": 17, "mean": 18, "var": 19, "m2": 20, "uniform": 21, "lt": 22}
DTYPE = {DataType.Float: 0, DataType.BFloat16: 1, DataType.Half: 2, DataType.Bool: 5}


def export_graph(fd):
    """(nodes, inputs, outputs): one tuple per recorded value, in recording order (arguments precede users)."""
    index, nodes = {}, []
    uf = fd._uf
    classes: dict = {}

    def cls(d):
        if d is None:
            return -1
        r = uf.find(d)
        return classes.setdefault(r, len(classes))

    for v in fd._vals:
        scalar = isinstance(v, Scalar)
        op = v.op
        value, size_class = 0.0, -1
        if scalar and v.value is not None and v.dim is None:
            op, value = "const", float(v.value)
        elif scalar and op == "size":
            size_class = cls(v.dim)
        kind = KIND.get(op)
        if kind is None:
            raise ValueError(f"op {op} has no counterpart in the expression graph")
        args = [index[a.id] for a in v.args if a.id in index]
        dims = [] if scalar else [cls(d) for d in v.dims]
        red = list(v.attrs.get("dims", [])) if not scalar and op in ("sum", "max", "mean", "var", "m2") else []
        dt = DTYPE.get(v.dtype, 3 if not scalar else 4 if v.dtype in (DataType.Int, DataType.Int32) else 3)
        index[v.id] = len(nodes)
        nodes.append((kind, dt, scalar, args, dims, red, int(v.attrs.get("correction", 0)) if not scalar else 0, value,
                      size_class))
    return nodes, [index[v.id] for v in fd._inputs], [index[v.id] for v in fd._outputs]


def cpp_match(fd) -> dict:
    """The C++ classifier's verdict on a recorded fusion: {"matched": bool, "op", "flags", "eps", "roles", ...}."""
    nodes, ins, outs = export_graph(fd)
    return H.load().match_segment(nodes, ins, outs)

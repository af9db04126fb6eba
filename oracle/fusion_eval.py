"""ATen evaluator for recorded fusions — TEST INFRASTRUCTURE ONLY.

The equivalent of what ``testValidate`` does when no explicit reference is given
(csrc/validator_utils.cpp:364-368): walk the *unscheduled* fusion and evaluate every
expression with ATen, exactly as ``Expr::evaluate`` does per node type
(csrc/ir/internal_nodes.cpp: UnaryOp :571-700 — Cast → tensor.to(dtype), Reciprocal → 1.0/x,
Rsqrt, Exp; BinaryOp :771-868 — scalars are python doubles; BroadcastOp :1494-1512 — a view
with size-1 dims; ReductionOp :1726-1778 — at::sum / at::amax; WelfordOp :2058-2083 —
at::var_mean(unbiased=False), var_sum = var * N).

It operates on the recorder's IR (``fuser_b200.ir``): the product never imports this module.
"""
from __future__ import annotations

import torch

from fuser_b200.ir import DataType, Scalar, Tensor, nvfuser_dtype_to_torch


def evaluate(fd, inputs, device="cpu", rng=None):
    """Return the fusion's outputs computed node by node with ATen on ``device``. ``rng`` = (seed, offset) of the
    Philox stream a ``uniform`` node draws from (the arguments the reference hands its kernels: csrc/rng.cpp)."""
    env: dict[int, object] = {}
    it = iter(inputs)
    for node in fd._inputs:
        v = next(it)
        env[node.id] = v.to(device) if isinstance(v, torch.Tensor) else v

    ext = {}
    for node, t in zip(fd._inputs, inputs):
        if isinstance(t, torch.Tensor) and hasattr(node, "dims"):
            for ax, d in enumerate(node.dims):
                if d is not None:
                    ext[fd._uf.find(d)] = int(t.shape[ax])

    def val(n):
        if n.id in env:
            return env[n.id]
        r = _eval(n)
        env[n.id] = r
        return r

    def _eval(n):
        op = n.op
        if isinstance(n, Scalar):
            if n.value is not None and n.dim is None:
                return n.value
            if op == "size":
                return int(val(n.args[0]).shape[n.attrs["axis"]])
            a = [val(x) for x in n.args]
            if op == "cast":
                return float(a[0]) if n.dtype in (DataType.Double, DataType.Float) else int(a[0])
            return {"neg": lambda: -a[0], "reciprocal": lambda: 1.0 / a[0], "sqrt": lambda: a[0] ** 0.5,
                    "rsqrt": lambda: a[0] ** -0.5, "add": lambda: a[0] + a[1], "sub": lambda: a[0] - a[1],
                    "mul": lambda: a[0] * a[1], "div": lambda: a[0] / a[1], "pow": lambda: a[0] ** a[1]}[op]()
        a = [val(x) for x in n.args]
        if op == "uniform":
            from oracle import aten_oracle as AO
            import numpy as np

            if rng is None:
                raise ValueError("a fusion with uniform() needs rng=(seed, offset)")
            if float(a[0]) != 0.0 or float(a[1]) != 1.0:
                raise NotImplementedError("uniform(minval, maxval) other than [0, 1)")
            shape = [ext[fd._uf.find(d)] for d in n.dims]
            numel = int(np.prod(shape))
            w = np.stack(AO.philox4x32_10(rng[0], np.arange((numel + 3) // 4, dtype=np.uint64), rng[1]), axis=1).reshape(-1)[:numel]
            return torch.from_numpy(AO.uniformf(w)).view(*shape).to(device)
        if op == "lt":
            x, y = a
            if isinstance(x, torch.Tensor) and not isinstance(y, torch.Tensor):
                return x.double() < float(y) if x.dtype == torch.float32 else x < y  # float tensor against a double scalar
            return x < y
        if op == "cast":
            return a[0].to(nvfuser_dtype_to_torch(n.attrs["dtype"]))
        if op == "neg":
            return -a[0]
        if op == "reciprocal":
            return 1.0 / a[0]
        if op == "sqrt":
            return torch.sqrt(a[0])
        if op == "rsqrt":
            return torch.rsqrt(a[0])
        if op == "exp":
            return torch.exp(a[0])
        if op == "log":
            return torch.log(a[0])
        if op == "abs":
            return torch.abs(a[0])
        if op in ("add", "sub", "mul", "div", "pow"):
            f = {"add": torch.add, "sub": torch.sub, "mul": torch.mul, "div": torch.div, "pow": torch.pow}[op]
            x, y = a
            if not isinstance(x, torch.Tensor):
                if op == "pow":
                    return torch.pow(torch.tensor(float(x)), y) if False else (float(x) ** y)
                return {"add": lambda: x + y, "sub": lambda: x - y, "mul": lambda: x * y, "div": lambda: x / y}[op]()
            return f(x, y)
        if op == "broadcast":
            t = a[0]
            for i, flag in enumerate(n.attrs["mask"]):
                if flag:
                    t = t.unsqueeze(i)
            return t
        if op == "broadcast_in_dim":
            t = a[0]
            rank = n.attrs["rank"]
            bd = n.attrs["broadcast_dims"]
            shape = [1] * rank
            for i, j in enumerate(bd):
                shape[j] = t.shape[i]
            return t.reshape(shape)  # size-1 view; expansion happens by broadcasting, as in BroadcastOp::evaluate
        dims = n.attrs.get("dims")
        keep = n.attrs.get("keepdim", False)
        if op == "sum":
            return torch.sum(a[0], dim=dims, keepdim=keep)
        if op == "max":
            return torch.amax(a[0], dim=dims, keepdim=keep)
        if op in ("mean", "var", "m2"):
            corr = n.attrs.get("correction", 0)
            var, mean = torch.var_mean(a[0], dim=dims, correction=0 if op == "m2" else corr, keepdim=keep)
            if op == "mean":
                return mean
            if op == "var":
                return var
            cnt = 1
            for d in dims:
                cnt *= a[0].shape[d]
            return var * cnt
        raise NotImplementedError(op)

    return [val(o) for o in fd._outputs]

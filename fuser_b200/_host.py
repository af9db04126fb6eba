"""Loader of the C++ host executor (``fuser_b200/lib/_b200norm_host.so``).

The extension is built in-tree by ``fuser_b200._build.build_host`` (g++ against this
interpreter's torch, linked to ``libb200norm.so``); a missing module is a hard error.
``NormExecutor`` mirrors the reference's ``CutlassExecutor`` four-method interface
(csrc/runtime/cutlass_executor.h:23-66); see csrc/host/norm_executor.h.
"""
from __future__ import annotations

import importlib.util
from pathlib import Path

from . import _lib as L

HOST_PATH = Path(__file__).resolve().parent / "lib" / "_b200norm_host.so"
_mod = None
_executors: dict[tuple, object] = {}


def load():
    global _mod
    if _mod is not None:
        return _mod
    if not HOST_PATH.exists():
        raise ImportError(
            f"{HOST_PATH} is missing: run `python -m fuser_b200._build` (or __graft_entry__.build()). "
            "fuser_b200 has no CPU or eager fallback."
        )
    import torch  # noqa: F401  (libtorch must be loaded before the extension)

    L.load()  # libb200norm.so first, so the extension binds to the in-tree copy
    spec = importlib.util.spec_from_file_location("_b200norm_host", str(HOST_PATH))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _mod = mod
    return mod


def executor(op: int, flags: int = 0, eps: float = 0.0, variant: int = 0, cluster_size: int = 0, correction: int = 0,
             dropout_keep: float = 1.0, ignore_index: int = -100):
    """Compiled ``NormExecutor`` for a segment spec (cached like FusionKernelRuntime::executors_)."""
    key = (op, flags, eps, variant, cluster_size, correction, dropout_keep, ignore_index)
    ex = _executors.get(key)
    if ex is None:
        m = load()
        ex = m.NormExecutor(0, 0, 0, len(_executors))
        ex.compile(m.SegmentSpec(op, flags, eps, variant, cluster_size, correction, dropout_keep, ignore_index))
        _executors[key] = ex
    return ex


def run(op: int, flags: int, eps: float, variant: int, cluster_size: int, inputs, outputs=(), correction: int = 0,
        dropout_keep: float = 1.0, ignore_index: int = -100):
    """Launch a segment; library status codes come back as B200NormError / UnsupportedSegment."""
    ex = executor(op, flags, eps, variant, cluster_size, correction, dropout_keep, ignore_index)
    try:
        return ex.run(list(inputs), list(outputs))
    except _mod.NormError as e:
        code, msg = e.args
        if code == L.ERR_UNSUPPORTED:
            raise L.UnsupportedSegment(code, msg.split(": ", 1)[-1]) from None
        raise L.B200NormError(code, msg.split(": ", 1)[-1]) from None

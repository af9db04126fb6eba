"""ctypes wrapper of oracle/c/norm_oracle.c (fp64, plain C) — TEST INFRASTRUCTURE ONLY.

Builds the shared object on first use with gcc (``make -C oracle``). Inputs are
torch CPU tensors of any float dtype; they are widened exactly to float64.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import torch

_DIR = Path(__file__).resolve().parent
_SO = _DIR / "_build" / "libnorm_oracle.so"
_lib = None


def build() -> Path:
    src = _DIR / "c" / "norm_oracle.c"
    if not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_DIR)], check=True, capture_output=True)
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_SO))
    return _lib


def _d(t):
    return None if t is None else t.detach().cpu().to(torch.float64).contiguous()


def _p(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def layer_norm_fwd(x, w, b, eps):
    xd, wd, bd = _d(x), _d(w), _d(b)
    n = xd.shape[-1]
    m = xd.numel() // n
    y = torch.empty_like(xd)
    mean = torch.empty(m, dtype=torch.float64)
    invstd = torch.empty(m, dtype=torch.float64)
    _load().oracle_layer_norm_fwd(_p(xd), _p(wd), _p(bd), C.c_int64(m), C.c_int64(n), C.c_double(eps), _p(y),
                                  _p(mean), _p(invstd))
    shp = tuple(x.shape[:-1]) + (1,)
    return y, mean.view(shp), invstd.view(shp)


def rms_norm_fwd(x, w, eps, stat_is_rms=False):
    xd, wd = _d(x), _d(w)
    n = xd.shape[-1]
    m = xd.numel() // n
    y = torch.empty_like(xd)
    stat = torch.empty(m, dtype=torch.float64)
    _load().oracle_rms_norm_fwd(_p(xd), _p(wd), C.c_int64(m), C.c_int64(n), C.c_double(eps),
                                C.c_int(int(stat_is_rms)), _p(y), _p(stat))
    return y, stat.view(tuple(x.shape[:-1]) + (1,))


def softmax_fwd(x, log=False):
    xd = _d(x)
    n = xd.shape[-1]
    y = torch.empty_like(xd)
    _load().oracle_softmax_fwd(_p(xd), C.c_int64(xd.numel() // n), C.c_int64(n), C.c_int(int(log)), _p(y))
    return y


def softmax_bwd(dy, y, log=False):
    dyd, yd = _d(dy), _d(y)
    n = yd.shape[-1]
    dx = torch.empty_like(yd)
    _load().oracle_softmax_bwd(_p(dyd), _p(yd), C.c_int64(yd.numel() // n), C.c_int64(n), C.c_int(int(log)), _p(dx))
    return dx


def norm_bwd(dy, x, mean, invstd, w, stat_is_rms=False, centered=True, want_db=True):
    dyd, xd, md, sd, wd = _d(dy), _d(x), _d(mean), _d(invstd), _d(w)
    n = xd.shape[-1]
    m = xd.numel() // n
    dx = torch.empty_like(xd)
    dw = torch.empty(n, dtype=torch.float64)
    db = torch.empty(n, dtype=torch.float64) if want_db else None
    _load().oracle_norm_bwd(_p(dyd), _p(xd), _p(md), _p(sd), _p(wd), C.c_int64(m), C.c_int64(n),
                            C.c_int(int(stat_is_rms)), C.c_int(int(centered)), _p(dx), _p(dw), _p(db))
    return dx, dw, db


def reduce_inner(x, kind="sum"):
    xd = _d(x)
    n = xd.shape[-1]
    m = xd.numel() // n
    out = torch.empty(m, dtype=torch.float64)
    _load().oracle_reduce_inner(_p(xd), C.c_int64(m), C.c_int64(n), C.c_int({"sum": 0, "max": 1, "fmax": 2}[kind]),
                                _p(out))
    return out.view(x.shape[:-1])


def reduce_outer(x):
    xd = _d(x)
    m, n = xd.shape
    out = torch.empty(n, dtype=torch.float64)
    _load().oracle_reduce_outer(_p(xd), C.c_int64(m), C.c_int64(n), _p(out))
    return out

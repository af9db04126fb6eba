"""WORKLOADS, not product code: the fusion definitions fed to the public API by bench.py's end-to-end arm and the
tests, spelled with the same FusionDefinition calls a user of nvfuser_direct makes. They restate, op for op, the graphs of the reference's
benchmarks (benchmarks/python/test_{layernorm,rmsnorm,softmax}_{fwd,bwd}.py — the Thunder
traces) and of its C++ op builders (csrc/ops/normalization.cpp), so that the matcher is
exercised on every spelling of the hot path that exists in the reference."""
from fuser_b200.frontend import DataType

PROMOTE = (DataType.Half, DataType.BFloat16)


def _up(fd, t, dtype):
    return fd.ops.cast(t, dtype=DataType.Float) if dtype in PROMOTE else t


def _down(fd, t, dtype):
    return fd.ops.cast(t, dtype=dtype) if dtype in PROMOTE else t


# ---- Thunder / Python-benchmark spellings ---------------------------------------------------
def layernorm_fwd_py(fd, dtype, eps=1e-5):
    """benchmarks/python/test_layernorm_fwd.py:14-54 → outputs (y, mean[M], invstd[M,1])"""
    T0 = fd.define_tensor(shape=[-1, -1], contiguity=[True, True], dtype=dtype, is_cpu=False)
    T1 = fd.define_tensor(shape=[-1], contiguity=[True], dtype=dtype, is_cpu=False)
    T2 = fd.define_tensor(shape=[-1], contiguity=[True], dtype=dtype, is_cpu=False)
    T0, T1, T2 = (_up(fd, t, dtype) for t in (T0, T1, T2))
    T3, T4 = fd.ops.var_mean(T0, dims=[1], correction=0, keepdim=False)
    T7 = fd.ops.broadcast_in_dim(T3, shape=[T0.size(0), 1], broadcast_dims=[0])
    T11 = fd.ops.broadcast_in_dim(T4, shape=[T0.size(0), 1], broadcast_dims=[0])
    S12 = fd.define_scalar(eps, dtype=DataType.Double)
    T14 = fd.ops.rsqrt(fd.ops.add(T7, S12))
    V17 = T0.shape()
    T19 = fd.ops.sub(T0, fd.ops.broadcast_in_dim(T11, shape=V17, broadcast_dims=[0, 1]))
    T24 = fd.ops.mul(T19, fd.ops.broadcast_in_dim(T14, shape=V17, broadcast_dims=[0, 1]))
    T26 = fd.ops.mul(T24, fd.ops.broadcast_in_dim(T1, shape=V17, broadcast_dims=[1]))
    T28 = fd.ops.add(T26, fd.ops.broadcast_in_dim(T2, shape=V17, broadcast_dims=[1]))
    fd.add_output(_down(fd, T28, dtype))
    fd.add_output(T4)
    fd.add_output(T14)


def layernorm_bwd_py(fd, dtype):
    """benchmarks/python/test_layernorm_bwd.py:20-101; inputs (x, dy, mean[M], invstd[M,1], w) →
    outputs (dx, dw, db)"""
    T0 = fd.define_tensor(shape=[-1, -1], contiguity=[True, True], dtype=dtype, is_cpu=False)
    T1 = fd.define_tensor(shape=[-1, -1], contiguity=[True, True], dtype=dtype, is_cpu=False)
    T2 = fd.define_tensor(shape=[-1], contiguity=[True], dtype=DataType.Float, is_cpu=False)
    T3 = fd.define_tensor(shape=[-1, 1], contiguity=[True, None], dtype=DataType.Float, is_cpu=False)
    T4 = fd.define_tensor(shape=[-1], contiguity=[True], dtype=dtype, is_cpu=False)
    T0, T1, T4 = (_up(fd, t, dtype) for t in (T0, T1, T4))
    T9 = fd.ops.broadcast_in_dim(T2, shape=[T0.size(0), 1], broadcast_dims=[0])
    V12 = T0.shape()
    T13 = fd.ops.broadcast_in_dim(T9, shape=V12, broadcast_dims=[0, 1])
    T14 = fd.ops.sub(T0, T13)
    T18 = fd.ops.broadcast_in_dim(T3, shape=V12, broadcast_dims=[0, 1])
    T19 = fd.ops.mul(T14, T18)
    T23 = fd.ops.broadcast_in_dim(T4, shape=V12, broadcast_dims=[1])
    T28 = fd.ops.sum(T1, dims=[0], keepdim=False, dtype=DataType.Null)
    T30 = fd.ops.mul(T1, T23)
    T31 = fd.ops.mul(T1, T19)
    T32 = fd.ops.sum(T31, dims=[0], keepdim=False, dtype=DataType.Null)
    T34 = fd.ops.mul(T30, T18)
    T35 = fd.ops.mul(T30, T14)
    T36 = fd.ops.sum(T35, dims=[1], keepdim=False, dtype=DataType.Null)
    T40 = fd.ops.broadcast_in_dim(T36, shape=[T0.size(0), 1], broadcast_dims=[0])
    T41 = fd.ops.neg(T34)
    T42 = fd.ops.sum(T41, dims=[1], keepdim=False, dtype=DataType.Null)
    T46 = fd.ops.broadcast_in_dim(T42, shape=[T0.size(0), 1], broadcast_dims=[0])
    S47 = fd.define_scalar(-0.5, dtype=DataType.Double)
    T48 = fd.ops.mul(S47, T40)
    S49 = fd.define_scalar(3.0, dtype=DataType.Double)
    T50 = fd.ops.pow(T3, S49)
    T51 = fd.ops.mul(T48, T50)
    T54 = fd.ops.sum(T46, dims=[1], keepdim=False, dtype=DataType.Null)
    T55 = fd.ops.sum(T51, dims=[1], keepdim=False, dtype=DataType.Null)
    T59 = fd.ops.broadcast_in_dim(T55, shape=[T0.size(0), 1], broadcast_dims=[0])
    T63 = fd.ops.broadcast_in_dim(T59, shape=V12, broadcast_dims=[0, 1])
    T67 = fd.ops.broadcast_in_dim(T2, shape=[T0.size(0), 1], broadcast_dims=[0])
    T71 = fd.ops.broadcast_in_dim(T67, shape=V12, broadcast_dims=[0, 1])
    S72 = fd.define_scalar(2.0, dtype=DataType.Double)
    T73 = fd.ops.mul(S72, T63)
    T74 = fd.ops.sub(T0, T71)
    T75 = fd.ops.mul(T73, T74)
    S77 = fd.ops.reciprocal(T0.size(1))
    T78 = fd.ops.mul(T75, S77)
    T82 = fd.ops.broadcast_in_dim(T54, shape=[T0.size(0), 1], broadcast_dims=[0])
    T86 = fd.ops.broadcast_in_dim(T82, shape=V12, broadcast_dims=[0, 1])
    T88 = fd.ops.mul(S77, T86)
    T89 = fd.ops.add(T78, T88)
    T90 = fd.ops.add(T34, T89)
    fd.add_output(_down(fd, T90, dtype))
    fd.add_output(_down(fd, T32, dtype))
    fd.add_output(_down(fd, T28, dtype))


def rmsnorm_fwd_py(fd, dtype, eps=1e-5):
    """benchmarks/python/test_rmsnorm_fwd.py:14-45 → outputs (y, rms[M,1])"""
    T0 = fd.define_tensor(shape=[-1, -1], contiguity=[True, True], dtype=dtype, is_cpu=False)
    T1 = fd.define_tensor(shape=[-1], contiguity=[True], dtype=dtype, is_cpu=False)
    T0, T1 = _up(fd, T0, dtype), _up(fd, T1, dtype)
    S3 = fd.define_scalar(2.0, dtype=DataType.Double)
    T4 = fd.ops.pow(T0, S3)
    T5 = fd.ops.sum(T4, dims=[1], keepdim=False, dtype=DataType.Null)
    T9 = fd.ops.broadcast_in_dim(T5, shape=[T0.size(0), 1], broadcast_dims=[0])
    S11 = fd.ops.reciprocal(T0.size(1))
    T12 = fd.ops.mul(T9, S11)
    S13 = fd.define_scalar(eps, dtype=DataType.Double)
    T15 = fd.ops.sqrt(fd.ops.add(T12, S13))
    T20 = fd.ops.broadcast_in_dim(T15, shape=T0.shape(), broadcast_dims=[0, 1])
    T23 = fd.ops.mul(T0, fd.ops.reciprocal(T20))
    T27 = fd.ops.broadcast_in_dim(T1, shape=T0.shape(), broadcast_dims=[1])
    T29 = fd.ops.mul(T27, T23)
    fd.add_output(_down(fd, T29, dtype))
    fd.add_output(T15)


def rmsnorm_bwd_py(fd, dtype):
    """benchmarks/python/test_rmsnorm_bwd.py:20-76; inputs (x, rms[M,1], dy, w) → outputs (dx, dw)"""
    T4 = fd.define_tensor(shape=[-1, -1], contiguity=[True, True], dtype=dtype, is_cpu=False)
    T5 = fd.define_tensor(shape=[-1, 1], contiguity=[True, None], dtype=DataType.Float, is_cpu=False)
    T6 = fd.define_tensor(shape=[-1, -1], contiguity=[True, True], dtype=dtype, is_cpu=False)
    T7 = fd.define_tensor(shape=[-1], contiguity=[True], dtype=dtype, is_cpu=False)
    S0 = fd.define_scalar(2.0, dtype=DataType.Double)
    T4, T6, T7 = (_up(fd, t, dtype) for t in (T4, T6, T7))
    T14 = fd.ops.broadcast_in_dim(T5, shape=T4.shape(), broadcast_dims=[0, 1])
    T15 = fd.ops.reciprocal(T14)
    T16 = fd.ops.mul(T4, T15)
    T20 = fd.ops.broadcast_in_dim(T7, shape=T4.shape(), broadcast_dims=[1])
    T23 = fd.ops.mul(T6, T16)
    T24 = fd.ops.mul(T6, T20)
    T25 = fd.ops.sum(T23, dims=[0], keepdim=False, dtype=DataType.Null)
    T28 = fd.ops.mul(T24, T15)
    T29 = fd.ops.neg(T24)
    T30 = fd.ops.mul(T29, T4)
    T32 = fd.ops.pow(T14, S0)
    T33 = fd.ops.reciprocal(T32)
    T34 = fd.ops.mul(T30, T33)
    T35 = fd.ops.sum(T34, dims=[1], keepdim=False, dtype=DataType.Null)
    T41 = fd.ops.broadcast_in_dim(T35, shape=[T4.size(0), 1], broadcast_dims=[0])
    T43 = fd.ops.mul(S0, T5)
    T44 = fd.ops.reciprocal(T43)
    T45 = fd.ops.mul(T41, T44)
    S48 = fd.ops.reciprocal(T4.size(1))
    T49 = fd.ops.mul(T45, S48)
    T50 = fd.ops.sum(T49, dims=[1], keepdim=False, dtype=DataType.Null)
    T54 = fd.ops.broadcast_in_dim(T50, shape=[T4.size(0), 1], broadcast_dims=[0])
    T58 = fd.ops.broadcast_in_dim(T54, shape=T4.shape(), broadcast_dims=[0, 1])
    T59 = fd.ops.mul(T58, S0)
    T62 = fd.ops.mul(T59, T4)
    T63 = fd.ops.add(T28, T62)
    fd.add_output(_down(fd, T63, dtype))
    fd.add_output(_down(fd, T25, dtype))


def softmax_fwd_py(fd, dtype, reduction_axis=1):
    """benchmarks/python/test_softmax_fwd.py:14-42"""
    T0 = fd.define_tensor(shape=[-1, -1], contiguity=[True, True], dtype=dtype, is_cpu=False)
    T0 = _up(fd, T0, dtype)
    T2 = fd.ops.max(T0, dims=[reduction_axis], keepdim=False, dtype=DataType.Null)
    bcast = [False, False]
    bcast[reduction_axis] = True
    T13 = fd.ops.sub(T0, fd.ops.broadcast(T2, is_broadcast_dim=bcast))
    T14 = fd.ops.exp(T13)
    T15 = fd.ops.sum(T14, dims=[reduction_axis], keepdim=False, dtype=DataType.Null)
    T26 = fd.ops.reciprocal(fd.ops.broadcast(T15, is_broadcast_dim=bcast))
    T27 = fd.ops.mul(T14, T26)
    fd.add_output(_down(fd, T27, dtype))


def softmax_bwd_py(fd, dtype, reduction_axis=1):
    """benchmarks/python/test_softmax_bwd.py:20-66; inputs (y, dy). The row sum makes a round
    trip through the io dtype."""
    T0 = fd.define_tensor(shape=[-1, -1], contiguity=[True, True], dtype=dtype, is_cpu=False)
    T1 = fd.define_tensor(shape=[-1, -1], contiguity=[True, True], dtype=dtype, is_cpu=False)
    if dtype is not DataType.Float:
        T0 = fd.ops.cast(T0, dtype=DataType.Float)
        T1 = fd.ops.cast(T1, dtype=DataType.Float)
    T4 = fd.ops.mul(T0, T1)
    T5 = fd.ops.sum(T4, dims=[reduction_axis], keepdim=False, dtype=DataType.Null)
    shape_v9 = [T0.size(0), 1] if reduction_axis else [1, T0.size(1)]
    T10 = fd.ops.broadcast_in_dim(T5, shape=shape_v9, broadcast_dims=[1 - reduction_axis])
    if dtype is not DataType.Float:
        T10 = fd.ops.cast(T10, dtype=dtype)
    T16 = fd.ops.broadcast_in_dim(T10, shape=[T0.size(0), T0.size(1)], broadcast_dims=[0, 1])
    if dtype is not DataType.Float:
        T16 = fd.ops.cast(T16, dtype=DataType.Float)
    T19 = fd.ops.mul(T0, fd.ops.sub(T1, T16))
    if dtype is not DataType.Float:
        T19 = fd.ops.cast(T19, dtype=dtype)
    fd.add_output(T19)


def batchnorm_fwd_py(fd, dtype, num_dims=4, eps=1e-5, momentum=0.01):
    """benchmarks/python/normalization.py:13-95 for norm = "batch_norm", channels_last = True: training
    statistics over every axis but the last, per-channel affine; the running statistics are fusion inputs
    whose updates are computed but never added as outputs (dead code, exactly as in the reference)."""
    cdim = num_dims - 1
    mask = [i != cdim for i in range(num_dims)]
    axes = [i for i in range(num_dims) if i != cdim]
    x = fd.define_tensor(shape=[-1] * num_dims, contiguity=[True] * num_dims, dtype=dtype, is_cpu=False)
    weight = fd.define_tensor(shape=[-1], contiguity=[True], dtype=dtype, is_cpu=False)
    bias = fd.define_tensor(shape=[-1], contiguity=[True], dtype=dtype, is_cpu=False)
    running_mean = fd.define_tensor(shape=[-1], contiguity=[True], dtype=DataType.Float, is_cpu=False)
    running_var = fd.define_tensor(shape=[-1], contiguity=[True], dtype=DataType.Float, is_cpu=False)
    x, weight, bias = _up(fd, x, dtype), _up(fd, weight, dtype), _up(fd, bias, dtype)
    var, mean = fd.ops.var_mean(x, dims=axes, correction=0, keepdim=False)
    invstd = fd.ops.rsqrt(fd.ops.add(var, fd.define_scalar(eps, dtype=DataType.Double)))
    x_norm = fd.ops.mul(fd.ops.sub(x, fd.ops.broadcast(mean, mask)), fd.ops.broadcast(invstd, mask))
    out = fd.ops.add(fd.ops.mul(x_norm, fd.ops.broadcast(weight, mask)), fd.ops.broadcast(bias, mask))
    rev = fd.define_scalar(1 - momentum, dtype=DataType.Double)
    mom = fd.define_scalar(momentum, dtype=DataType.Double)
    fd.ops.add(fd.ops.mul(mom, mean), fd.ops.mul(rev, running_mean))  # updated_mean: not an output
    fd.ops.add(fd.ops.mul(mom, var), fd.ops.mul(rev, running_var))    # updated_var: not an output
    fd.add_output(_down(fd, out, dtype))
    fd.add_output(mean)
    fd.add_output(invstd)


def batchnorm_bwd_py(fd, dtype, num_dims=4):
    """benchmarks/python/normalization.py:98-190 for norm = "batch_norm", channels_last = True; inputs
    (input, grad, weight, running_mean, running_var, mean, invstd) — the running statistics are unused."""
    from functools import reduce

    cdim = num_dims - 1
    mask = [i != cdim for i in range(num_dims)]
    axes = [i for i in range(num_dims) if i != cdim]
    x = fd.define_tensor(shape=[-1] * num_dims, contiguity=[True] * num_dims, dtype=dtype, is_cpu=False)
    grad = fd.define_tensor(shape=[-1] * num_dims, contiguity=[True] * num_dims, dtype=dtype, is_cpu=False)
    weight = fd.define_tensor(shape=[-1], contiguity=[True], dtype=dtype, is_cpu=False)
    fd.define_tensor(shape=[-1], contiguity=[True], dtype=DataType.Float, is_cpu=False)  # running_mean
    fd.define_tensor(shape=[-1], contiguity=[True], dtype=DataType.Float, is_cpu=False)  # running_var
    mean = fd.define_tensor(shape=[-1], contiguity=[True], dtype=DataType.Float, is_cpu=False)
    invstd = fd.define_tensor(shape=[-1], contiguity=[True], dtype=DataType.Float, is_cpu=False)
    x, grad, weight = _up(fd, x, dtype), _up(fd, grad, dtype), _up(fd, weight, dtype)
    norm = fd.ops.reciprocal(reduce(fd.ops.mul, [x.size(ax) for ax in axes], 1))
    mean_b = fd.ops.broadcast(mean, mask)
    grad_sum = fd.ops.sum(grad, dims=axes, keepdim=False)
    x_sub_mean = fd.ops.sub(x, mean_b)
    dot_p = fd.ops.sum(fd.ops.mul(grad, x_sub_mean), dims=axes, keepdim=False)
    grad_mean = fd.ops.broadcast(fd.ops.mul(grad_sum, norm), mask)
    proj_scale = fd.ops.broadcast(fd.ops.mul(fd.ops.mul(dot_p, norm), fd.ops.mul(invstd, invstd)), mask)
    grad_scale = fd.ops.mul(fd.ops.broadcast(weight, mask), fd.ops.broadcast(invstd, mask))
    proj = fd.ops.mul(proj_scale, x_sub_mean)
    grad_input = fd.ops.mul(fd.ops.sub(fd.ops.sub(grad, proj), grad_mean), grad_scale)
    grad_weight = fd.ops.mul(dot_p, invstd)
    fd.add_output(_down(fd, grad_input, dtype))
    fd.add_output(_down(fd, grad_weight, dtype))
    fd.add_output(_down(fd, grad_sum, dtype))


# ---- C++ builder spellings (csrc/ops/normalization.cpp) ----------------------------------------
def layer_norm_cpp(fd, dtype, eps=1e-5, ndim=2, weight=True, bias=True):
    """layer_norm :304-342 via Welford; outputs (y, mean[..,1], invstd[..,1])"""
    x = fd.define_tensor(shape=[-1] * ndim, dtype=dtype)
    w = fd.define_tensor(shape=[-1], dtype=dtype) if weight else None
    b = fd.define_tensor(shape=[-1], dtype=dtype) if bias else None
    xf = _up(fd, x, dtype)
    last = ndim - 1
    avg, m2, n = fd.ops.welford(xf, dims=[last])
    mask = [False] * last + [True]
    mean_b = fd.ops.broadcast(avg, is_broadcast_dim=mask)
    var = fd.ops.mul(fd.ops.broadcast(m2, is_broadcast_dim=mask), fd.ops.reciprocal(xf.size(last)))
    invstd = fd.ops.rsqrt(fd.ops.add(var, fd.define_scalar(eps, dtype=DataType.Double)))
    y = fd.ops.mul(fd.ops.sub(xf, mean_b), invstd)
    outer = [True] * last + [False]
    if weight:
        y = fd.ops.mul(y, fd.ops.broadcast(_up(fd, w, dtype), is_broadcast_dim=outer))
    if bias:
        y = fd.ops.add(y, fd.ops.broadcast(_up(fd, b, dtype), is_broadcast_dim=outer))
    fd.add_output(_down(fd, y, dtype))
    fd.add_output(mean_b)
    fd.add_output(invstd)


def rms_norm_cpp(fd, dtype, eps=1e-6, weight=True):
    """rms_norm :352-380; outputs (y, invstd[M,1])"""
    x = fd.define_tensor(shape=[-1, -1], dtype=dtype)
    w = fd.define_tensor(shape=[-1], dtype=dtype) if weight else None
    xf = _up(fd, x, dtype)
    ss = fd.ops.sum(fd.ops.mul(xf, xf), dims=[1])
    var = fd.ops.mul(fd.ops.broadcast(ss, is_broadcast_dim=[False, True]), fd.ops.reciprocal(xf.size(1)))
    invstd = fd.ops.rsqrt(fd.ops.add(var, fd.define_scalar(eps, dtype=DataType.Double)))
    y = fd.ops.mul(xf, invstd)
    if weight:
        y = fd.ops.mul(y, fd.ops.broadcast(_up(fd, w, dtype), is_broadcast_dim=[True, False]))
    fd.add_output(_down(fd, y, dtype))
    fd.add_output(invstd)


def norm_bwd_cpp(fd, dtype, layer_norm=True, weight=True, mask=(True, True, True)):
    """layer_norm_backward :382-436 / rms_norm_backward :438-485 (the latter keeps -sum(g));
    inputs (dy, x, [mean[M,1]], invstd[M,1], [w]); outputs by mask (dx, dw, db)"""
    dy = fd.define_tensor(shape=[-1, -1], dtype=dtype)
    x = fd.define_tensor(shape=[-1, -1], dtype=dtype)
    mean = fd.define_tensor(shape=[-1, 1], dtype=DataType.Float) if layer_norm else None
    invstd = fd.define_tensor(shape=[-1, 1], dtype=DataType.Float)
    w = fd.define_tensor(shape=[-1], dtype=dtype) if weight else None
    dyf, xf = _up(fd, dy, dtype), _up(fd, x, dtype)
    x_hat = fd.ops.mul(fd.ops.sub(xf, mean), invstd) if layer_norm else fd.ops.mul(xf, invstd)
    g = fd.ops.mul(dyf, fd.ops.broadcast(_up(fd, w, dtype), is_broadcast_dim=[True, False])) if weight else dyf
    n = xf.size(1)
    a = fd.ops.mul(n, g)
    b = fd.ops.broadcast(fd.ops.sum(g, dims=[1]), is_broadcast_dim=[False, True])
    c2 = fd.ops.broadcast(fd.ops.sum(fd.ops.mul(g, x_hat), dims=[1]), is_broadcast_dim=[False, True])
    inner = fd.ops.sub(fd.ops.sub(a, b), fd.ops.mul(x_hat, c2))
    if mask[0]:
        fd.add_output(_down(fd, fd.ops.mul(fd.ops.mul(fd.ops.reciprocal(n), invstd), inner), dtype))
    if mask[1] and weight:
        fd.add_output(_down(fd, fd.ops.sum(fd.ops.mul(dyf, x_hat), dims=[0]), dtype))
    if layer_norm and len(mask) > 2 and mask[2]:
        fd.add_output(_down(fd, fd.ops.sum(dyf, dims=[0]), dtype))


def softmax_cpp(fd, dtype, log=False):
    """softmax :169-189 / log_softmax :212-231"""
    x = fd.define_tensor(shape=[-1, -1], dtype=dtype)
    xf = _up(fd, x, dtype)
    mx = fd.ops.broadcast(fd.ops.max(xf, dims=[1]), is_broadcast_dim=[False, True])
    d = fd.ops.sub(xf, mx)
    e = fd.ops.exp(d)
    if log:
        y = fd.ops.sub(d, fd.ops.log(fd.ops.sum(e, dims=[1], keepdim=True)))
    else:
        s = fd.ops.broadcast(fd.ops.sum(e, dims=[1]), is_broadcast_dim=[False, True])
        y = fd.ops.mul(e, fd.ops.reciprocal(s))
    fd.add_output(_down(fd, y, dtype))


def softmax_bwd_cpp(fd, dtype, log=False):
    """softmax_backward :191-210 / log_softmax_backward :233-249; inputs (dy, y)"""
    dy = fd.define_tensor(shape=[-1, -1], dtype=dtype)
    y = fd.define_tensor(shape=[-1, -1], dtype=dtype)
    dyf, yf = _up(fd, dy, dtype), _up(fd, y, dtype)
    if log:
        dx = fd.ops.sub(dyf, fd.ops.mul(fd.ops.exp(yf), fd.ops.sum(dyf, dims=[1], keepdim=True)))
    else:
        g = fd.ops.mul(dyf, yf)
        s = fd.ops.broadcast(fd.ops.sum(g, dims=[1]), is_broadcast_dim=[False, True])
        dx = fd.ops.sub(g, fd.ops.mul(yf, s))
    fd.add_output(_down(fd, dx, dtype))


# ---- forward norms with a fused residual add ------------------------------
def add_layernorm_fwd_py(fd, dtype, eps=1e-5):
    """benchmarks/python/test_dropout_layernorm_fwd.py:18-76 with the dropout removed (p = 0):
    inputs (input1, input2, weights, bias) → outputs (layernorm(input2 + input1), mean[M], invstd[M,1])."""
    T2 = fd.define_tensor(shape=[-1, -1], contiguity=[True, True], dtype=dtype, is_cpu=False)
    T3 = fd.define_tensor(shape=[-1, -1], contiguity=[True, True], dtype=dtype, is_cpu=False)
    T1 = fd.define_tensor(shape=[-1], contiguity=[True], dtype=dtype, is_cpu=False)
    T0 = fd.define_tensor(shape=[-1], contiguity=[True], dtype=dtype, is_cpu=False)
    T0, T1, T2, T3 = (_up(fd, t, dtype) for t in (T0, T1, T2, T3))
    T16 = fd.ops.add(T3, T2)
    T17, T18 = fd.ops.var_mean(T16, dims=[1], correction=0, keepdim=False)
    T22 = fd.ops.broadcast_in_dim(T17, shape=[T2.size(0), 1], broadcast_dims=[0])
    T26 = fd.ops.broadcast_in_dim(T18, shape=[T2.size(0), 1], broadcast_dims=[0])
    S27 = fd.define_scalar(eps, dtype=DataType.Double)
    T29 = fd.ops.rsqrt(fd.ops.add(T22, S27))
    T34 = fd.ops.sub(T16, fd.ops.broadcast_in_dim(T26, shape=T2.shape(), broadcast_dims=[0, 1]))
    T39 = fd.ops.mul(T34, fd.ops.broadcast_in_dim(T29, shape=T2.shape(), broadcast_dims=[0, 1]))
    T45 = fd.ops.mul(T39, fd.ops.broadcast_in_dim(T1, shape=T2.shape(), broadcast_dims=[1]))
    T51 = fd.ops.add(T45, fd.ops.broadcast_in_dim(T0, shape=T2.shape(), broadcast_dims=[1]))
    fd.add_output(_down(fd, T51, dtype))
    fd.add_output(T18)
    fd.add_output(T29)


def add_rmsnorm_fwd_py(fd, dtype, eps=1e-5):
    """benchmarks/python/test_dropout_rmsnorm_fwd.py:18-82 with the dropout removed:
    inputs (input1, input2, weights) → outputs (rmsnorm(input2 + input1), rms[M,1])."""
    T0 = fd.define_tensor(shape=[-1, -1], contiguity=[True, True], dtype=dtype, is_cpu=False)
    T2 = fd.define_tensor(shape=[-1, -1], contiguity=[True, True], dtype=dtype, is_cpu=False)
    T1 = fd.define_tensor(shape=[-1], contiguity=[True], dtype=dtype, is_cpu=False)
    T0, T1, T2 = (_up(fd, t, dtype) for t in (T0, T1, T2))
    V6 = T0.shape()
    T15 = fd.ops.add(T2, T0)
    T17 = fd.ops.pow(T15, fd.define_scalar(2.0, dtype=DataType.Double))
    T18 = fd.ops.sum(T17, dims=[1], keepdim=False, dtype=DataType.Null)
    T22 = fd.ops.broadcast_in_dim(T18, shape=[T0.size(0), 1], broadcast_dims=[0])
    T25 = fd.ops.mul(T22, fd.ops.reciprocal(T0.size(1)))
    T28 = fd.ops.sqrt(fd.ops.add(T25, fd.define_scalar(eps, dtype=DataType.Double)))
    T35 = fd.ops.reciprocal(fd.ops.broadcast_in_dim(T28, shape=V6, broadcast_dims=[0, 1]))
    T36 = fd.ops.mul(T15, T35)
    T42 = fd.ops.mul(fd.ops.broadcast_in_dim(T1, shape=V6, broadcast_dims=[1]), T36)
    fd.add_output(_down(fd, T42, dtype))
    fd.add_output(T28)


def rmsnorm_add_thunder(fd, dtype, eps=1e-5):
    """benchmarks/python/torch_ops.py:76-83 (the op of test_rmsnorm_add_fwd.py) as a per-op trace:
    every torch op's result is cast back to the io dtype, so the norm sees the *rounded* sum;
    inputs (hidden, weight, residual) → outputs (weight * normed, hidden + residual)."""
    H = fd.define_tensor(shape=[-1, -1], contiguity=[True, True], dtype=dtype, is_cpu=False)
    W = fd.define_tensor(shape=[-1], contiguity=[True], dtype=dtype, is_cpu=False)
    R = fd.define_tensor(shape=[-1, -1], contiguity=[True, True], dtype=dtype, is_cpu=False)
    s = _down(fd, fd.ops.add(_up(fd, H, dtype), _up(fd, R, dtype)), dtype)   # hidden_states + residual
    sf = _up(fd, s, dtype)
    sq = fd.ops.mul(sf, sf)
    ms = fd.ops.mul(fd.ops.sum(sq, dims=[1], keepdim=True, dtype=DataType.Null), fd.ops.reciprocal(H.size(1)))
    inv = fd.ops.rsqrt(fd.ops.add(ms, fd.define_scalar(eps, dtype=DataType.Double)))
    normed = fd.ops.mul(sf, fd.ops.broadcast_in_dim(inv, shape=H.shape(), broadcast_dims=[0, 1]))
    y = fd.ops.mul(fd.ops.broadcast_in_dim(_up(fd, W, dtype), shape=H.shape(), broadcast_dims=[1]), normed)
    fd.add_output(_down(fd, y, dtype))
    fd.add_output(s)


def thunder_rms_norm_bwd_cpp(fd, dtype, weight=True):
    """thunder_rms_norm_backward as the C++ builder writes it (csrc/ops/normalization.cpp:487-535; used by
    tests/cpp/test_combined_inner_outer_reduction.cpp:1197): inputs dy, x, rms[M,1], w -> dx, dw. The projection
    term carries 1 / (2 rms) — half the textbook value."""
    dy = fd.define_tensor(shape=[-1, -1], contiguity=[True, True], dtype=dtype)
    x = fd.define_tensor(shape=[-1, -1], contiguity=[True, True], dtype=dtype)
    rms = fd.define_tensor(shape=[-1, 1], contiguity=[True, None], dtype=DataType.Float)
    w = fd.define_tensor(shape=[-1], contiguity=[True], dtype=dtype) if weight else None
    dyf, xf = _up(fd, dy, dtype), _up(fd, x, dtype)
    gw = fd.ops.mul(dyf, fd.ops.broadcast(_up(fd, w, dtype), is_broadcast_dim=[True, False])) if weight else dyf
    t = fd.ops.mul(fd.ops.mul(fd.ops.neg(gw), xf), fd.ops.reciprocal(fd.ops.mul(rms, rms)))
    inner = fd.ops.broadcast(fd.ops.sum(t, dims=[1]), is_broadcast_dim=[False, True])
    k = fd.ops.mul(fd.ops.reciprocal(fd.ops.mul(rms, 2.0)), fd.ops.reciprocal(xf.size(1)))
    dx = fd.ops.add(fd.ops.mul(inner, fd.ops.mul(xf, k)), fd.ops.mul(gw, fd.ops.reciprocal(rms)))
    fd.add_output(_down(fd, dx, dtype))
    if weight:
        dw = fd.ops.sum(fd.ops.mul(dyf, fd.ops.mul(xf, fd.ops.reciprocal(rms))), dims=[0])
        fd.add_output(_down(fd, dw, dtype))


def reduce_py(fd, dtype, kind="sum", dims=(1,)):
    """A plain Reduction segment (csrc/scheduler/reduction.cpp): sum / max over the last axis, or the sum over
    the outer axes; reduced-precision inputs are cast up first and the result cast back, as every benchmark does."""
    x = fd.define_tensor(shape=[-1, -1], contiguity=[True, True], dtype=dtype)
    r = (fd.ops.sum if kind == "sum" else fd.ops.max)(_up(fd, x, dtype), dims=list(dims))
    fd.add_output(_down(fd, r, dtype))


def var_mean_py(fd, dtype, correction=1):
    """Stand-alone variance_mean over the last axis (csrc/ops/normalization.cpp:93-158): fp32 statistics."""
    x = fd.define_tensor(shape=[-1, -1], contiguity=[True, True], dtype=dtype)
    var, mean = fd.ops.var_mean(_up(fd, x, dtype), dims=[1], correction=correction, keepdim=False)
    fd.add_output(var)
    fd.add_output(mean)


# ---------------------------------------------------------------------------
# fused dropout: norm(input2 + dropout(input1)) and the backward
# (the op sequences of benchmarks/python/test_dropout_{layernorm,rmsnorm}_{fwd,bwd}.py)
# ---------------------------------------------------------------------------
def _dropout_sum(fd, dtype, p, x1, x2, mask_f):
    """input2 + (input1 * mask) * 1 / (1 - p), in Float."""
    scale = fd.define_scalar(1 / (1 - p), dtype=DataType.Double)
    return fd.ops.add(x2, fd.ops.mul(fd.ops.mul(x1, mask_f), scale)), scale


def _draw_mask(fd, like, p):
    lo = fd.define_scalar(0.0, dtype=DataType.Double)
    hi = fd.define_scalar(1.0, dtype=DataType.Double)
    u = fd.ops.uniform(lo, hi, shape=like.shape(), dtype=DataType.Float)
    keep = fd.ops.lt(u, fd.define_scalar(1 - p, dtype=DataType.Double))
    return keep, fd.ops.cast(keep, dtype=DataType.Float)


def dropout_layernorm_fwd_py(fd, dtype, p=0.2, eps=1e-5):
    """inputs: input1, input2, weights, bias; outputs: y, mean, invstd, mask (test_dropout_layernorm_fwd.py:19-78)."""
    x1 = fd.define_tensor(shape=[-1, -1], contiguity=[True, True], dtype=dtype, is_cpu=False)
    x2 = fd.define_tensor(shape=[-1, -1], contiguity=[True, True], dtype=dtype, is_cpu=False)
    w = fd.define_tensor(shape=[-1], contiguity=[True], dtype=dtype, is_cpu=False)
    b = fd.define_tensor(shape=[-1], contiguity=[True], dtype=dtype, is_cpu=False)
    keep, keep_f = _draw_mask(fd, x1, p)
    x1f, x2f, wf, bf = (_up(fd, t, dtype) for t in (x1, x2, w, b))
    s, _ = _dropout_sum(fd, dtype, p, x1f, x2f, keep_f)
    var, mean = fd.ops.var_mean(s, dims=[1], correction=0, keepdim=False)
    var_b = fd.ops.broadcast_in_dim(var, shape=[x1.size(0), 1], broadcast_dims=[0])
    mean_b = fd.ops.broadcast_in_dim(mean, shape=[x1.size(0), 1], broadcast_dims=[0])
    invstd = fd.ops.rsqrt(fd.ops.add(var_b, fd.define_scalar(eps, dtype=DataType.Double)))
    xc = fd.ops.sub(s, fd.ops.broadcast_in_dim(mean_b, shape=x1.shape(), broadcast_dims=[0, 1]))
    xh = fd.ops.mul(xc, fd.ops.broadcast_in_dim(invstd, shape=x1.shape(), broadcast_dims=[0, 1]))
    y = fd.ops.mul(xh, fd.ops.broadcast_in_dim(wf, shape=x1.shape(), broadcast_dims=[1]))
    y = fd.ops.add(y, fd.ops.broadcast_in_dim(bf, shape=x1.shape(), broadcast_dims=[1]))
    fd.add_output(_down(fd, y, dtype))
    fd.add_output(mean)
    fd.add_output(invstd)
    fd.add_output(keep)


def dropout_rmsnorm_fwd_py(fd, dtype, p=0.2, eps=1e-5):
    """inputs: input1, input2, weights; outputs: y, mask, rms (test_dropout_rmsnorm_fwd.py:19-82)."""
    x1 = fd.define_tensor(shape=[-1, -1], contiguity=[True, True], dtype=dtype, is_cpu=False)
    x2 = fd.define_tensor(shape=[-1, -1], contiguity=[True, True], dtype=dtype, is_cpu=False)
    w = fd.define_tensor(shape=[-1], contiguity=[True], dtype=dtype, is_cpu=False)
    keep, keep_f = _draw_mask(fd, x1, p)
    x1f, wf, x2f = (_up(fd, t, dtype) for t in (x1, w, x2))
    s, _ = _dropout_sum(fd, dtype, p, x1f, x2f, keep_f)
    sq = fd.ops.sum(fd.ops.pow(s, fd.define_scalar(2.0, dtype=DataType.Double)), dims=[1], keepdim=False, dtype=DataType.Null)
    sq_b = fd.ops.broadcast_in_dim(sq, shape=[x1.size(0), 1], broadcast_dims=[0])
    ms = fd.ops.mul(sq_b, fd.ops.reciprocal(x1.size(1)))
    rms = fd.ops.sqrt(fd.ops.add(ms, fd.define_scalar(eps, dtype=DataType.Double)))
    xh = fd.ops.mul(s, fd.ops.reciprocal(fd.ops.broadcast_in_dim(rms, shape=x1.shape(), broadcast_dims=[0, 1])))
    y = fd.ops.mul(fd.ops.broadcast_in_dim(wf, shape=x1.shape(), broadcast_dims=[1]), xh)
    fd.add_output(_down(fd, y, dtype))
    fd.add_output(keep)
    fd.add_output(rms)


def dropout_layernorm_bwd_py(fd, dtype, p=0.2):
    """inputs: mask, mean [M], invstd [M,1], grads, weights, input1, input2; outputs: d input1, d input2, dweight, dbias
    (test_dropout_layernorm_bwd.py:18-119)."""
    mask = fd.define_tensor(shape=[-1, -1], contiguity=[True, True], dtype=DataType.Bool, is_cpu=False)
    mean = fd.define_tensor(shape=[-1], contiguity=[True], dtype=DataType.Float, is_cpu=False)
    invstd = fd.define_tensor(shape=[-1, 1], contiguity=[True, None], dtype=DataType.Float, is_cpu=False)
    dy = fd.define_tensor(shape=[-1, -1], contiguity=[True, True], dtype=dtype, is_cpu=False)
    w = fd.define_tensor(shape=[-1], contiguity=[True], dtype=dtype, is_cpu=False)
    x1 = fd.define_tensor(shape=[-1, -1], contiguity=[True, True], dtype=dtype, is_cpu=False)
    x2 = fd.define_tensor(shape=[-1, -1], contiguity=[True, True], dtype=dtype, is_cpu=False)
    size0, shp = x1.size(0), x1.shape()
    mask_f = fd.ops.cast(mask, dtype=DataType.Float) if dtype in PROMOTE else mask
    dyf, wf, x1f, x2f = (_up(fd, t, dtype) for t in (dy, w, x1, x2))
    s, scale = _dropout_sum(fd, dtype, p, x1f, x2f, mask_f)
    mean_b = fd.ops.broadcast_in_dim(fd.ops.broadcast_in_dim(mean, shape=[size0, 1], broadcast_dims=[0]), shape=shp,
                                     broadcast_dims=[0, 1])
    xc = fd.ops.sub(s, mean_b)
    is_b = fd.ops.broadcast_in_dim(invstd, shape=shp, broadcast_dims=[0, 1])
    xh = fd.ops.mul(xc, is_b)
    w_b = fd.ops.broadcast_in_dim(wf, shape=shp, broadcast_dims=[1])
    db = fd.ops.sum(dyf, dims=[0], keepdim=False, dtype=DataType.Null)
    g = fd.ops.mul(dyf, w_b)
    dw = fd.ops.sum(fd.ops.mul(dyf, xh), dims=[0], keepdim=False, dtype=DataType.Null)
    g_is = fd.ops.mul(g, is_b)
    t43 = fd.ops.sum(fd.ops.mul(g, xc), dims=[1], keepdim=False, dtype=DataType.Null)
    t47 = fd.ops.broadcast_in_dim(t43, shape=[size0, 1], broadcast_dims=[0])
    t49 = fd.ops.sum(fd.ops.neg(g_is), dims=[1], keepdim=False, dtype=DataType.Null)
    t53 = fd.ops.broadcast_in_dim(t49, shape=[size0, 1], broadcast_dims=[0])
    t55 = fd.ops.mul(fd.define_scalar(-0.5, dtype=DataType.Double), t47)
    t58 = fd.ops.mul(t55, fd.ops.pow(invstd, fd.define_scalar(3.0, dtype=DataType.Double)))
    t61 = fd.ops.sum(t53, dims=[1], keepdim=False, dtype=DataType.Null)
    t62 = fd.ops.sum(t58, dims=[1], keepdim=False, dtype=DataType.Null)
    t70 = fd.ops.broadcast_in_dim(fd.ops.broadcast_in_dim(t62, shape=[size0, 1], broadcast_dims=[0]), shape=shp,
                                  broadcast_dims=[0, 1])
    t80 = fd.ops.mul(fd.define_scalar(2.0, dtype=DataType.Double), t70)
    t82 = fd.ops.mul(t80, fd.ops.sub(s, mean_b))
    inv_n = fd.ops.reciprocal(x1.size(1))
    t85 = fd.ops.mul(t82, inv_n)
    t93 = fd.ops.broadcast_in_dim(fd.ops.broadcast_in_dim(t61, shape=[size0, 1], broadcast_dims=[0]), shape=shp,
                                  broadcast_dims=[0, 1])
    t96 = fd.ops.add(t85, fd.ops.mul(inv_n, t93))
    dres = fd.ops.add(g_is, t96)
    dx = fd.ops.mul(fd.ops.mul(dres, scale), mask_f)
    fd.add_output(_down(fd, dx, dtype))
    fd.add_output(_down(fd, dres, dtype))
    fd.add_output(_down(fd, dw, dtype))
    fd.add_output(_down(fd, db, dtype))


def dropout_rmsnorm_bwd_py(fd, dtype, p=0.2):
    """inputs: input1, input2, mask, rms [M,1], grads, weights; outputs: d input1, d input2, dweight
    (test_dropout_rmsnorm_bwd.py:20-108)."""
    x1 = fd.define_tensor(shape=[-1, -1], contiguity=[True, True], dtype=dtype, is_cpu=False)
    x2 = fd.define_tensor(shape=[-1, -1], contiguity=[True, True], dtype=dtype, is_cpu=False)
    mask = fd.define_tensor(shape=[-1, -1], contiguity=[True, True], dtype=DataType.Bool, is_cpu=False)
    rms = fd.define_tensor(shape=[-1, 1], contiguity=[True, None], dtype=DataType.Float, is_cpu=False)
    dy = fd.define_tensor(shape=[-1, -1], contiguity=[True, True], dtype=dtype, is_cpu=False)
    w = fd.define_tensor(shape=[-1], contiguity=[True], dtype=dtype, is_cpu=False)
    shp = x1.shape()
    mask_f = fd.ops.cast(mask, dtype=DataType.Float) if dtype in PROMOTE else mask
    x2f, x1f, dyf, wf = (_up(fd, t, dtype) for t in (x2, x1, dy, w))
    s, scale = _dropout_sum(fd, dtype, p, x1f, x2f, mask_f)
    rms_b = fd.ops.broadcast_in_dim(rms, shape=shp, broadcast_dims=[0, 1])
    inv = fd.ops.reciprocal(rms_b)
    xh = fd.ops.mul(s, inv)
    w_b = fd.ops.broadcast_in_dim(wf, shape=shp, broadcast_dims=[1])
    g = fd.ops.mul(dyf, w_b)
    dw = fd.ops.sum(fd.ops.mul(dyf, xh), dims=[0], keepdim=False, dtype=DataType.Null)
    g_inv = fd.ops.mul(g, inv)
    two = fd.define_scalar(2.0, dtype=DataType.Double)
    t41 = fd.ops.mul(fd.ops.mul(fd.ops.neg(g), s), fd.ops.reciprocal(fd.ops.pow(rms_b, two)))
    t42 = fd.ops.sum(t41, dims=[1], keepdim=False, dtype=DataType.Null)
    t47 = fd.ops.broadcast_in_dim(t42, shape=[x1.size(0), 1], broadcast_dims=[0])
    t52 = fd.ops.mul(t47, fd.ops.reciprocal(fd.ops.mul(two, rms)))
    t56 = fd.ops.mul(t52, fd.ops.reciprocal(x1.size(1)))
    t57 = fd.ops.sum(t56, dims=[1], keepdim=False, dtype=DataType.Null)
    t65 = fd.ops.broadcast_in_dim(fd.ops.broadcast_in_dim(t57, shape=[x1.size(0), 1], broadcast_dims=[0]), shape=shp,
                                  broadcast_dims=[0, 1])
    dres = fd.ops.add(g_inv, fd.ops.mul(fd.ops.mul(t65, two), s))
    dx = fd.ops.mul(fd.ops.mul(dres, scale), mask_f)
    fd.add_output(_down(fd, dx, dtype))
    fd.add_output(_down(fd, dres, dtype))
    fd.add_output(_down(fd, dw, dtype))


def norm_fwd_py(fd, dtype, norm="instance_norm", num_dims=4, channels_last=False, eps=1e-5, momentum=0.01):
    """benchmarks/python/normalization.py:13-95 in full: batch norm or instance norm, channels first or last."""
    cdim = 1 if not channels_last else num_dims - 1
    bmask = [i != cdim for i in range(num_dims)]
    cmask = list(bmask)
    axes = [i for i in range(num_dims) if i != cdim]
    if norm == "instance_norm":
        axes.remove(0)
        bmask[0] = False
    x = fd.define_tensor(shape=[-1] * num_dims, contiguity=[True] * num_dims, dtype=dtype, is_cpu=False)
    weight = fd.define_tensor(shape=[-1], contiguity=[True], dtype=dtype, is_cpu=False)
    bias = fd.define_tensor(shape=[-1], contiguity=[True], dtype=dtype, is_cpu=False)
    running_mean = fd.define_tensor(shape=[-1], contiguity=[True], dtype=DataType.Float, is_cpu=False)
    running_var = fd.define_tensor(shape=[-1], contiguity=[True], dtype=DataType.Float, is_cpu=False)
    x, weight, bias = _up(fd, x, dtype), _up(fd, weight, dtype), _up(fd, bias, dtype)
    var, mean = fd.ops.var_mean(x, dims=axes, correction=0, keepdim=False)
    invstd = fd.ops.rsqrt(fd.ops.add(var, fd.define_scalar(eps, dtype=DataType.Double)))
    x_norm = fd.ops.mul(fd.ops.sub(x, fd.ops.broadcast(mean, bmask)), fd.ops.broadcast(invstd, bmask))
    out = fd.ops.add(fd.ops.mul(x_norm, fd.ops.broadcast(weight, cmask)), fd.ops.broadcast(bias, cmask))
    rev = fd.define_scalar(1 - momentum, dtype=DataType.Double)
    mom = fd.define_scalar(momentum, dtype=DataType.Double)
    um = fd.ops.add(fd.ops.mul(mom, mean), fd.ops.mul(rev, running_mean))  # running-statistics updates: dead code,
    uv = fd.ops.add(fd.ops.mul(mom, var), fd.ops.mul(rev, running_var))    # never added as outputs
    if 0 not in axes:
        inv_b = fd.ops.reciprocal(x.size(0))
        fd.ops.mul(fd.ops.sum(um, 0), inv_b)
        fd.ops.mul(fd.ops.sum(uv, 0), inv_b)
    fd.add_output(_down(fd, out, dtype))
    fd.add_output(mean)
    fd.add_output(invstd)


def norm_bwd_py(fd, dtype, norm="instance_norm", num_dims=4, channels_last=False):
    """benchmarks/python/normalization.py:98-190 in full (inputs: input, grad, weight, running_mean, running_var,
    mean, invstd)."""
    from functools import reduce

    cdim = 1 if not channels_last else num_dims - 1
    bmask = [i != cdim for i in range(num_dims)]
    cmask = list(bmask)
    axes = [i for i in range(num_dims) if i != cdim]
    if norm == "instance_norm":
        axes.remove(0)
        bmask[0] = False
    nstat = num_dims - len(axes)
    x = fd.define_tensor(shape=[-1] * num_dims, contiguity=[True] * num_dims, dtype=dtype, is_cpu=False)
    grad = fd.define_tensor(shape=[-1] * num_dims, contiguity=[True] * num_dims, dtype=dtype, is_cpu=False)
    weight = fd.define_tensor(shape=[-1], contiguity=[True], dtype=dtype, is_cpu=False)
    fd.define_tensor(shape=[-1], contiguity=[True], dtype=DataType.Float, is_cpu=False)  # running_mean
    fd.define_tensor(shape=[-1], contiguity=[True], dtype=DataType.Float, is_cpu=False)  # running_var
    mean = fd.define_tensor(shape=[-1] * nstat, contiguity=[True] * nstat, dtype=DataType.Float, is_cpu=False)
    invstd = fd.define_tensor(shape=[-1] * nstat, contiguity=[True] * nstat, dtype=DataType.Float, is_cpu=False)
    x, grad, weight = _up(fd, x, dtype), _up(fd, grad, dtype), _up(fd, weight, dtype)
    norm_ = fd.ops.reciprocal(reduce(fd.ops.mul, [x.size(ax) for ax in axes], 1))
    mean_b = fd.ops.broadcast(mean, bmask)
    grad_sum = fd.ops.sum(grad, dims=axes, keepdim=False)
    x_sub_mean = fd.ops.sub(x, mean_b)
    dot_p = fd.ops.sum(fd.ops.mul(grad, x_sub_mean), dims=axes, keepdim=False)
    grad_mean = fd.ops.broadcast(fd.ops.mul(grad_sum, norm_), bmask)
    proj_scale = fd.ops.broadcast(fd.ops.mul(fd.ops.mul(dot_p, norm_), fd.ops.mul(invstd, invstd)), bmask)
    grad_scale = fd.ops.mul(fd.ops.broadcast(weight, cmask), fd.ops.broadcast(invstd, bmask))
    proj = fd.ops.mul(proj_scale, x_sub_mean)
    grad_input = fd.ops.mul(fd.ops.sub(fd.ops.sub(grad, proj), grad_mean), grad_scale)
    grad_weight = fd.ops.mul(dot_p, invstd)
    grad_bias = grad_sum
    if 0 not in axes:
        grad_weight = fd.ops.sum(grad_weight, 0, keepdim=False)
        grad_bias = fd.ops.sum(grad_bias, 0, keepdim=False)
    fd.add_output(_down(fd, grad_input, dtype))
    fd.add_output(_down(fd, grad_weight, dtype))
    fd.add_output(_down(fd, grad_bias, dtype))

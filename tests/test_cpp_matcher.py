"""The C++ segment classifier (csrc/host/segment_matcher.cpp — what NormExecutor::supported(Fusion*) runs in an
nvFuser tree, csrc/runtime/executor_dispatch.cpp:46-59) must agree with the Python matcher on every spelling of
the hot path: the Thunder traces of benchmarks/python, the C++ op builders of csrc/ops/normalization.cpp, the
fused-add, outer, reduction and statistics segments — and refuse what the Python matcher refuses."""
import pytest

from fuser_b200 import _lib as L
from fuser_b200._export import cpp_match, export_graph
from fuser_b200.frontend import DataType, FusionDefinition, UnsupportedSegment
from tests import fusions as F

DTS = [DataType.Float, DataType.BFloat16, DataType.Half]
IO = {DataType.Float: 0, DataType.BFloat16: 1, DataType.Half: 2}


def build(fn, *a, **k):
    with FusionDefinition() as fd:
        fn(fd, *a, **k)
    return fd


def agree(fd):
    """Both classifiers on the same recorded fusion: same verdict, same (op, flags, eps, roles, outputs)."""
    c = cpp_match(fd)
    try:
        m = fd.segment()
    except UnsupportedSegment:
        assert not c["matched"], f"C++ matched what Python refused: {c}"
        return None
    assert c["matched"], f"C++ refused ({c.get('why')}) what Python matched as {m.description}"
    assert (c["op"], c["flags"], c["eps"], c["correction"]) == (m.op, m.flags, m.eps, m.correction), (c, m)
    assert c["roles"] == m.roles and c["outputs"] == m.outputs and c["io_dtype"] == IO[m.io_dtype], (c, m)
    assert c["dropout_keep"] == m.dropout_keep and c["channels_last"] == m.channels_last, (c, m)
    return m


SPELLINGS = [
    ("layernorm_fwd_py", lambda dt: (F.layernorm_fwd_py, (dt,))),
    ("layernorm_bwd_py", lambda dt: (F.layernorm_bwd_py, (dt,))),
    ("rmsnorm_fwd_py", lambda dt: (F.rmsnorm_fwd_py, (dt,))),
    ("rmsnorm_bwd_py", lambda dt: (F.rmsnorm_bwd_py, (dt,))),
    ("softmax_fwd_py", lambda dt: (F.softmax_fwd_py, (dt,))),
    ("softmax_bwd_py", lambda dt: (F.softmax_bwd_py, (dt,))),
    ("softmax_fwd_py_outer", lambda dt: (F.softmax_fwd_py, (dt, 0))),
    ("softmax_bwd_py_outer", lambda dt: (F.softmax_bwd_py, (dt, 0))),
    ("batchnorm_fwd_py_2d", lambda dt: (F.batchnorm_fwd_py, (dt, 2))),
    ("batchnorm_fwd_py_4d", lambda dt: (F.batchnorm_fwd_py, (dt, 4))),
    ("batchnorm_bwd_py_2d", lambda dt: (F.batchnorm_bwd_py, (dt, 2))),
    ("batchnorm_bwd_py_4d", lambda dt: (F.batchnorm_bwd_py, (dt, 4))),
    ("layer_norm_cpp", lambda dt: (F.layer_norm_cpp, (dt,))),
    ("layer_norm_cpp_3d_nobias", lambda dt: (lambda fd, d: F.layer_norm_cpp(fd, d, ndim=3, bias=False), (dt,))),
    ("layer_norm_cpp_noaffine", lambda dt: (lambda fd, d: F.layer_norm_cpp(fd, d, weight=False, bias=False), (dt,))),
    ("rms_norm_cpp", lambda dt: (F.rms_norm_cpp, (dt,))),
    ("rms_norm_cpp_noweight", lambda dt: (lambda fd, d: F.rms_norm_cpp(fd, d, weight=False), (dt,))),
    ("layer_norm_bwd_cpp", lambda dt: (lambda fd, d: F.norm_bwd_cpp(fd, d, layer_norm=True), (dt,))),
    ("rms_norm_bwd_cpp", lambda dt: (lambda fd, d: F.norm_bwd_cpp(fd, d, layer_norm=False), (dt,))),
    ("layer_norm_bwd_cpp_dx_only", lambda dt: (lambda fd, d: F.norm_bwd_cpp(fd, d, mask=(True, False, False)), (dt,))),
    ("rms_norm_bwd_cpp_noweight", lambda dt: (lambda fd, d: F.norm_bwd_cpp(fd, d, layer_norm=False, weight=False), (dt,))),
    ("softmax_cpp", lambda dt: (F.softmax_cpp, (dt,))),
    ("log_softmax_cpp", lambda dt: (lambda fd, d: F.softmax_cpp(fd, d, log=True), (dt,))),
    ("softmax_bwd_cpp", lambda dt: (F.softmax_bwd_cpp, (dt,))),
    ("log_softmax_bwd_cpp", lambda dt: (lambda fd, d: F.softmax_bwd_cpp(fd, d, log=True), (dt,))),
    ("add_layernorm_fwd_py", lambda dt: (F.add_layernorm_fwd_py, (dt,))),
    ("add_rmsnorm_fwd_py", lambda dt: (F.add_rmsnorm_fwd_py, (dt,))),
    ("rmsnorm_add_thunder", lambda dt: (F.rmsnorm_add_thunder, (dt,))),
    ("thunder_rms_norm_bwd_cpp", lambda dt: (F.thunder_rms_norm_bwd_cpp, (dt,))),
    ("thunder_rms_norm_bwd_cpp_noweight", lambda dt: (lambda fd, d: F.thunder_rms_norm_bwd_cpp(fd, d, weight=False), (dt,))),
    ("instance_norm_fwd_nchw", lambda dt: (F.norm_fwd_py, (dt, "instance_norm", 4, False))),
    ("instance_norm_fwd_nhwc", lambda dt: (F.norm_fwd_py, (dt, "instance_norm", 4, True))),
    ("instance_norm_fwd_3d", lambda dt: (F.norm_fwd_py, (dt, "instance_norm", 3, False))),
    ("instance_norm_bwd_nchw", lambda dt: (F.norm_bwd_py, (dt, "instance_norm", 4, False))),
    ("instance_norm_bwd_nhwc", lambda dt: (F.norm_bwd_py, (dt, "instance_norm", 5, True))),
    ("batch_norm_fwd_full_def", lambda dt: (F.norm_fwd_py, (dt, "batch_norm", 4, True))),
    ("batch_norm_bwd_full_def", lambda dt: (F.norm_bwd_py, (dt, "batch_norm", 4, True))),
    ("dropout_layernorm_fwd_py", lambda dt: (F.dropout_layernorm_fwd_py, (dt, 0.2))),
    ("dropout_rmsnorm_fwd_py", lambda dt: (F.dropout_rmsnorm_fwd_py, (dt, 0.1))),
    ("dropout_layernorm_bwd_py", lambda dt: (F.dropout_layernorm_bwd_py, (dt, 0.2))),
    ("dropout_rmsnorm_bwd_py", lambda dt: (F.dropout_rmsnorm_bwd_py, (dt, 0.5))),
]


@pytest.mark.parametrize("dt", DTS, ids=lambda d: d.name)
@pytest.mark.parametrize("spelling", SPELLINGS, ids=lambda s: s[0])
def test_cpp_and_python_classifiers_agree(spelling, dt):
    fn, args = spelling[1](dt)
    m = agree(build(fn, *args))
    assert m is not None, "every spelling of the hot path is a supported segment"


def _red(kind, dims, dt, nd=2, out_float=False):
    def f(fd):
        x = fd.define_tensor(shape=[-1] * nd, dtype=dt)
        xf = fd.ops.cast(x, DataType.Float) if dt != DataType.Float else x
        r = (fd.ops.sum if kind == "sum" else fd.ops.max)(xf, dims=dims)
        fd.add_output(r if (out_float or dt == DataType.Float) else fd.ops.cast(r, dt))
    return f


@pytest.mark.parametrize("dt", DTS, ids=lambda d: d.name)
def test_reductions_and_statistics_agree(dt):
    for kind, dims, nd in (("sum", [1], 2), ("max", [-1], 2), ("sum", [0], 2), ("sum", [0, 1], 3), ("max", [2], 3)):
        assert agree(build(_red(kind, dims, dt, nd))) is not None
    assert agree(build(_red("sum", [1], dt, nd=3))) is None  # middle axis

    def vm(corr, welford=False):
        def f(fd):
            x = fd.define_tensor(shape=[-1, -1], dtype=DataType.Float)
            if welford:
                avg, m2, _ = fd.ops.welford(x, [1])
                fd.add_output(avg)
                fd.add_output(m2)
            else:
                var, mean = fd.ops.var_mean(x, dims=[1], correction=corr, keepdim=False)
                fd.add_output(var)
                fd.add_output(mean)
        return f
    for corr in (0, 1, 2):
        m = agree(build(vm(corr)))
        assert m is not None and m.correction == corr
    assert agree(build(vm(0, welford=True))) is not None


def test_near_misses_are_refused_by_both():
    def leaky(fd):  # LayerNorm whose output keeps an extra + 0.5
        x = fd.define_tensor(shape=[-1, -1], dtype=DataType.Float)
        var, mean = fd.ops.var_mean(x, dims=[1], correction=0, keepdim=True)
        y = fd.ops.mul(fd.ops.sub(x, mean), fd.ops.rsqrt(fd.ops.add(var, 1e-12)))
        fd.add_output(fd.ops.add(y, 0.5))

    def unsafe_softmax(fd):  # no max subtraction
        x = fd.define_tensor(shape=[-1, -1], dtype=DataType.Float)
        e = fd.ops.exp(x)
        fd.add_output(fd.ops.mul(e, fd.ops.reciprocal(fd.ops.sum(e, dims=[1], keepdim=True))))

    def pointwise(fd):
        x = fd.define_tensor(shape=[-1, -1], dtype=DataType.Float)
        fd.add_output(fd.ops.mul(x, x))

    def mixed_out(fd):  # bf16 in, Float sum out
        x = fd.define_tensor(shape=[-1, -1], dtype=DataType.BFloat16)
        fd.add_output(fd.ops.sum(fd.ops.cast(x, DataType.Float), dims=[1]))

    for fn in (leaky, unsafe_softmax, pointwise, mixed_out):
        assert agree(build(fn)) is None


def test_epsilon_travels_through_the_cpp_classifier():
    def ln(fd, eps):
        x = fd.define_tensor(shape=[-1, -1], dtype=DataType.Float)
        var, mean = fd.ops.var_mean(x, dims=[1], correction=0, keepdim=True)
        fd.add_output(fd.ops.mul(fd.ops.sub(x, mean), fd.ops.rsqrt(fd.ops.add(var, eps))))
    for eps in (1e-12, 1e-5, 0.25):
        c = cpp_match(build(lambda fd: ln(fd, eps)))
        assert c["matched"] and c["op"] == L.OP_LAYERNORM_FWD and c["eps"] == eps


def test_exported_graph_is_plain_data():
    nodes, ins, outs = export_graph(build(F.rmsnorm_bwd_py, DataType.BFloat16))
    assert len(ins) == 4 and len(outs) == 2
    for i, n in enumerate(nodes):
        assert all(a < i for a in n[3]), "arguments precede their users"

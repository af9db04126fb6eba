"""CPU tests of the frontend: the recorder keeps nvfuser_direct's call surface, and the matcher
classifies every spelling of the hot path (Thunder traces and C++ builders) while rejecting
fusions that are not normalization segments."""
import pytest
import torch

from fuser_b200 import _lib as L
from fuser_b200.frontend import DataType, FusionDefinition, UnsupportedSegment, get_tolerance
from tests import fusions as F

DTS = [DataType.Float, DataType.BFloat16, DataType.Half]


def build(fn, *a, **k):
    with FusionDefinition() as fd:
        fn(fd, *a, **k)
    return fd


@pytest.mark.parametrize("dt", DTS, ids=lambda d: d.name)
def test_python_benchmark_fusions_are_matched(dt):
    m = build(F.layernorm_fwd_py, dt).segment()
    assert (m.op, m.outputs, m.eps) == (L.OP_LAYERNORM_FWD, ["y", "mean", "invstd"], 1e-5)
    assert m.flags & L.F_HAS_WEIGHT and m.flags & L.F_HAS_BIAS and m.roles == {"x": 0, "w": 1, "b": 2}
    m = build(F.layernorm_bwd_py, dt).segment()
    assert (m.op, m.outputs) == (L.OP_LAYERNORM_BWD, ["dx", "dw", "db"])
    assert m.roles == {"dy": 1, "x": 0, "mean": 2, "stat": 3, "w": 4}
    m = build(F.rmsnorm_fwd_py, dt).segment()
    assert (m.op, m.outputs) == (L.OP_RMSNORM_FWD, ["y", "rms"]) and m.flags & L.F_STAT_IS_RMS
    m = build(F.rmsnorm_bwd_py, dt).segment()
    assert (m.op, m.outputs) == (L.OP_RMSNORM_BWD, ["dx", "dw"])
    assert m.flags & L.F_STAT_IS_RMS and not (m.flags & L.F_RMS_BWD_CENTERED)
    assert m.roles == {"dy": 2, "x": 0, "stat": 1, "w": 3}
    m = build(F.softmax_fwd_py, dt).segment()
    assert (m.op, m.outputs) == (L.OP_SOFTMAX_FWD, ["y"])
    m = build(F.softmax_bwd_py, dt).segment()
    assert (m.op, m.outputs, m.roles) == (L.OP_SOFTMAX_BWD, ["dx"], {"dy": 1, "y": 0})
    assert bool(m.flags & L.F_ROUND_ROWSUM) == (dt is not DataType.Float)


@pytest.mark.parametrize("dt", DTS, ids=lambda d: d.name)
def test_cpp_builder_fusions_are_matched(dt):
    m = build(F.layer_norm_cpp, dt, eps=1e-5).segment()
    assert (m.op, m.outputs) == (L.OP_LAYERNORM_FWD, ["y", "mean", "invstd"])
    m = build(F.layer_norm_cpp, dt, ndim=3, bias=False).segment()
    assert m.op == L.OP_LAYERNORM_FWD and not (m.flags & L.F_HAS_BIAS)
    m = build(F.layer_norm_cpp, dt, weight=False, bias=False).segment()
    assert m.op == L.OP_LAYERNORM_FWD and not (m.flags & (L.F_HAS_BIAS | L.F_HAS_WEIGHT))
    m = build(F.rms_norm_cpp, dt, eps=1e-6).segment()
    assert (m.op, m.outputs, m.eps) == (L.OP_RMSNORM_FWD, ["y", "invstd"], 1e-6)
    assert not (m.flags & L.F_STAT_IS_RMS)
    m = build(F.norm_bwd_cpp, dt, layer_norm=True).segment()
    assert (m.op, m.outputs) == (L.OP_LAYERNORM_BWD, ["dx", "dw", "db"])
    assert m.roles == {"dy": 0, "x": 1, "mean": 2, "stat": 3, "w": 4}
    m = build(F.norm_bwd_cpp, dt, layer_norm=True, mask=(True, False, False)).segment()
    assert m.outputs == ["dx"]
    # the C++ rms_norm_backward keeps the -sum(g) term: a different polynomial from the textbook form
    m = build(F.norm_bwd_cpp, dt, layer_norm=False).segment()
    assert (m.op, m.outputs) == (L.OP_RMSNORM_BWD, ["dx", "dw"])
    assert m.flags & L.F_RMS_BWD_CENTERED and not (m.flags & L.F_STAT_IS_RMS)
    for log in (False, True):
        m = build(F.softmax_cpp, dt, log=log).segment()
        assert m.op == (L.OP_LOGSOFTMAX_FWD if log else L.OP_SOFTMAX_FWD)
        m = build(F.softmax_bwd_cpp, dt, log=log).segment()
        assert m.op == (L.OP_LOGSOFTMAX_BWD if log else L.OP_SOFTMAX_BWD) and m.roles == {"dy": 0, "y": 1}


def test_equivalent_respellings_match_and_near_misses_do_not():
    # 1/sqrt instead of rsqrt, x*x*(1/N) as div, weight multiplied first
    def respelled(fd):
        x = fd.define_tensor(shape=[-1, -1], dtype=DataType.Float)
        w = fd.define_tensor(shape=[-1], dtype=DataType.Float)
        ms = fd.ops.div(fd.ops.sum(fd.ops.pow(x, 2.0), dims=[1], keepdim=True), x.size(1))
        inv = fd.ops.reciprocal(fd.ops.sqrt(fd.ops.add(ms, 1e-6)))
        wb = fd.ops.broadcast_in_dim(w, shape=x.shape(), broadcast_dims=[1])
        fd.add_output(fd.ops.mul(fd.ops.mul(wb, x), inv))
        fd.add_output(inv)
    m = build(respelled).segment()
    assert (m.op, m.outputs, m.eps) == (L.OP_RMSNORM_FWD, ["y", "invstd"], 1e-6)

    # near miss 1: variance with Bessel's correction is not LayerNorm's statistic
    def unbiased_ln(fd):
        x = fd.define_tensor(shape=[-1, -1], dtype=DataType.Float)
        var, mean = fd.ops.var_mean(x, dims=[1], correction=1, keepdim=True)
        fd.add_output(fd.ops.mul(fd.ops.sub(x, mean), fd.ops.rsqrt(fd.ops.add(var, 1e-5))))
    with pytest.raises(UnsupportedSegment):
        build(unbiased_ln).segment()

    # near miss 2: softmax without the max subtraction
    def unsafe_softmax(fd):
        x = fd.define_tensor(shape=[-1, -1], dtype=DataType.Float)
        e = fd.ops.exp(x)
        fd.add_output(fd.ops.mul(e, fd.ops.reciprocal(fd.ops.sum(e, dims=[1], keepdim=True))))
    with pytest.raises(UnsupportedSegment):
        build(unsafe_softmax).segment()

    # near miss 3: softmax over the middle axis of a rank-3 tensor (neither inner nor outer persistent)
    def middle_axis(fd):
        x = fd.define_tensor(shape=[-1, -1, -1], dtype=DataType.Float)
        d = fd.ops.sub(x, fd.ops.max(x, dims=[1], keepdim=True))
        e = fd.ops.exp(d)
        fd.add_output(fd.ops.mul(e, fd.ops.reciprocal(fd.ops.sum(e, dims=[1], keepdim=True))))
    with pytest.raises(UnsupportedSegment):
        build(middle_axis).segment()

    # near miss 4: a pointwise fusion is somebody else's segment
    def pointwise(fd):
        x = fd.define_tensor(shape=[-1, -1], dtype=DataType.Float)
        fd.add_output(fd.ops.mul(x, x))
    with pytest.raises(UnsupportedSegment):
        build(pointwise).segment()



def test_thunder_rms_norm_backward_as_written():
    """csrc/ops/normalization.cpp:487-535: the C++ builder's dx carries 1 / (2 rms) on the projection term. The
    fusion as written is its own segment kind (B200NORM_F_RMS_BWD_HALF_PROJ), distinct from the textbook form."""
    for dt in (DataType.Float, DataType.BFloat16):
        for weight in (True, False):
            m = build(lambda fd: F.thunder_rms_norm_bwd_cpp(fd, dt, weight)).segment()
            assert m.op == L.OP_RMSNORM_BWD
            assert m.flags & L.F_RMS_BWD_HALF_PROJ and m.flags & L.F_STAT_IS_RMS and not m.flags & L.F_RMS_BWD_CENTERED
            assert bool(m.flags & L.F_HAS_WEIGHT) == weight
            assert m.outputs == (["dx", "dw"] if weight else ["dx"])
    m = build(lambda fd: F.rmsnorm_bwd_py(fd, DataType.BFloat16)).segment()  # the Python trace compensates with * 2
    assert not m.flags & L.F_RMS_BWD_HALF_PROJ


def test_reduction_and_statistics_segments():
    """Plain Reduction segments (scheduler/reduction.cpp) and stand-alone variance_mean / Welford
    (csrc/ops/normalization.cpp:93-158) are routed to their kernels; near misses are refused."""
    def red(kind, dims, dt=DataType.BFloat16, nd=2):
        def f(fd):
            x = fd.define_tensor(shape=[-1] * nd, dtype=dt)
            xf = fd.ops.cast(x, DataType.Float) if dt != DataType.Float else x
            r = (fd.ops.sum if kind == "sum" else fd.ops.max)(xf, dims=dims)
            fd.add_output(fd.ops.cast(r, dt) if dt != DataType.Float else r)
        return f
    for dt in DTS:
        m = build(red("sum", [1], dt)).segment()
        assert (m.op, m.flags, m.outputs) == (L.OP_REDUCE_INNER, 0, ["r"])
        m = build(red("max", [-1], dt)).segment()
        assert (m.op, m.flags) == (L.OP_REDUCE_INNER, L.F_REDUCE_MAX)
        m = build(red("sum", [0], dt)).segment()
        assert m.op == L.OP_REDUCE_OUTER
        m = build(red("sum", [0, 1], dt, nd=3)).segment()
        assert m.op == L.OP_REDUCE_OUTER
    with pytest.raises(UnsupportedSegment):  # a reduction over a middle axis is neither
        build(red("sum", [1], nd=3)).segment()
    with pytest.raises(UnsupportedSegment):  # a bf16 input summed into a Float output: no kernel rounds there
        def mixed(fd):
            x = fd.define_tensor(shape=[-1, -1], dtype=DataType.BFloat16)
            fd.add_output(fd.ops.sum(fd.ops.cast(x, DataType.Float), dims=[1]))
        build(mixed).segment()

    def vm(corr, outs=("var", "mean"), keep=False):
        def f(fd):
            x = fd.define_tensor(shape=[-1, -1], dtype=DataType.Float)
            var, mean = fd.ops.var_mean(x, dims=[1], correction=corr, keepdim=keep)
            for o in outs:
                fd.add_output(var if o == "var" else mean)
        return f
    for corr in (0, 1, 2):
        m = build(vm(corr)).segment()
        assert m.op == L.OP_LAYERNORM_FWD and m.flags & L.F_NO_Y and m.flags & L.F_STAT_VAR
        assert m.correction == corr and m.outputs == ["var", "mean"]
    assert build(vm(1, ("mean",), keep=True)).segment().outputs == ["mean"]

    def wf(fd):
        x = fd.define_tensor(shape=[-1, -1], dtype=DataType.Float)
        avg, m2, _ = fd.ops.welford(x, [1])
        fd.add_output(avg)
        fd.add_output(m2)
    m = build(wf).segment()
    assert m.flags & L.F_STAT_M2 and m.outputs == ["mean", "m2"]


def test_epsilon_is_part_of_the_match():
    """A tiny epsilon (BERT-style 1e-12) must neither vanish from the normal form nor be replaced by another
    small constant of the fusion."""
    def ln(fd, eps, scale):
        x = fd.define_tensor(shape=[-1, -1], dtype=DataType.Float)
        var, mean = fd.ops.var_mean(x, dims=[1], correction=0, keepdim=True)
        y = fd.ops.mul(fd.ops.sub(x, mean), fd.ops.rsqrt(fd.ops.add(var, eps)))
        if scale is not None:  # another constant below one elsewhere in the fusion
            y = fd.ops.mul(fd.ops.add(y, scale), 1.0)
        fd.add_output(y)
    for eps in (1e-12, 1e-5, 0.25):
        m = build(lambda fd: ln(fd, eps, None)).segment()
        assert m.op == L.OP_LAYERNORM_FWD and m.eps == eps
    with pytest.raises(UnsupportedSegment):  # y + 0.5 is not a layer norm, whatever epsilon is tried
        build(lambda fd: ln(fd, 1e-12, 0.5)).segment()


def test_api_surface_and_errors():
    with FusionDefinition() as fd:
        x = fd.define_tensor(shape=[-1, -1], contiguity=[True, True], dtype=DataType.BFloat16, is_cpu=False)
        assert x.ndim == 2 and len(x.shape()) == 2 and x.size(1).is_symbolic()
        with pytest.raises(ValueError):  # same user error as csrc/ops/normalization.cpp:101-108
            fd.ops.var_mean(x, dims=[1])
        # operands of different rank: the lower one gets leading broadcast axes (csrc/ops/utils.cpp:22-34, 511-532)
        s_ = fd.ops.add(x, fd.define_tensor(shape=[-1], dtype=DataType.BFloat16))
        assert s_.ndim == 2 and s_.args[1].op == "broadcast" and s_.args[1].dims[0] is None
    with pytest.raises(NotImplementedError):
        FusionDefinition().execute([])
    fd = build(F.rmsnorm_fwd_py, DataType.BFloat16)
    with pytest.raises(ValueError):
        fd.execute([torch.zeros(4, 8, dtype=torch.bfloat16)])  # wrong arity
    if not torch.cuda.is_available():
        with pytest.raises(L.B200NormError):  # CPU tensors: no fallback
            fd.execute([torch.zeros(4, 8, dtype=torch.bfloat16), torch.zeros(8, dtype=torch.bfloat16)])


def test_reduction_sizes_follow_the_validator():
    # softmax output: the max-reduce feeds the sum-reduce -> N*N (csrc/validator_utils.cpp:233-254)
    fd = build(F.softmax_fwd_py, DataType.BFloat16)
    x = torch.zeros(16, 128, dtype=torch.bfloat16)
    assert fd.reduction_sizes([x]) == [128 * 128]
    fd = build(F.layernorm_bwd_py, DataType.Float)
    ins = [torch.zeros(16, 128), torch.zeros(16, 128), torch.zeros(16), torch.zeros(16, 1), torch.zeros(128)]
    rs = fd.reduction_sizes(ins)
    assert rs[1] == 16 and rs[2] == 16 and rs[0] >= 128
    assert get_tolerance(torch.float32, 4096) == pytest.approx((0.00012207, 0.00012207 * 0.01))


@pytest.mark.parametrize("dt", DTS, ids=lambda d: d.name)
def test_fused_residual_add_fusions_are_matched(dt):
    """norm(x + residual): the dropout_{layernorm,rmsnorm}_fwd definitions without the dropout and
    the rmsnorm_add op of torch_ops.py."""
    m = build(F.add_layernorm_fwd_py, dt).segment()
    assert (m.op, m.outputs, m.eps) == (L.OP_LAYERNORM_FWD, ["y", "mean", "invstd"], 1e-5)
    assert m.flags & L.F_ADD_RESIDUAL and not (m.flags & (L.F_OUT_SUM | L.F_ROUND_SUM))
    assert m.roles == {"x": 0, "r": 1, "w": 2, "b": 3}
    m = build(F.add_rmsnorm_fwd_py, dt).segment()
    assert (m.op, m.outputs) == (L.OP_RMSNORM_FWD, ["y", "rms"])
    assert m.flags & L.F_ADD_RESIDUAL and m.flags & L.F_STAT_IS_RMS and not (m.flags & L.F_OUT_SUM)
    assert m.roles == {"x": 0, "r": 1, "w": 2}
    m = build(F.rmsnorm_add_thunder, dt).segment()
    assert (m.op, m.outputs) == (L.OP_RMSNORM_FWD, ["y", "sum"])
    assert m.flags & L.F_ADD_RESIDUAL and m.flags & L.F_OUT_SUM and m.roles == {"x": 0, "r": 2, "w": 1}
    assert bool(m.flags & L.F_ROUND_SUM) == (dt is not DataType.Float)


def test_add_of_two_matrices_alone_is_not_a_segment():
    def just_add(fd):
        a = fd.define_tensor(shape=[-1, -1], dtype=DataType.Float)
        b = fd.define_tensor(shape=[-1, -1], dtype=DataType.Float)
        fd.add_output(fd.ops.add(a, b))
    with pytest.raises(UnsupportedSegment):
        build(just_add).segment()


@pytest.mark.parametrize("dt", [DataType.Float, DataType.BFloat16])
def test_outer_persistent_softmax_matches(dt):
    """reduction_axis = 0 of the reference's softmax benchmarks (pytest mark outer_persistent,
    benchmarks/python/test_softmax_{fwd,bwd}.py): matched as the same ops with F_OUTER; the evaluated
    fusion agrees with eager softmax over dim 0."""
    import torch
    from oracle import fusion_eval

    fd = build(F.softmax_fwd_py, dt, reduction_axis=0)
    m = fd.segment()
    assert m.op == L.OP_SOFTMAX_FWD and m.flags & L.F_OUTER and m.outputs == ["y"]
    td = torch.float32 if dt == DataType.Float else torch.bfloat16
    x = torch.randn(37, 24).to(td)
    (y,) = fusion_eval.evaluate(fd, [x])
    assert torch.allclose(y.float(), torch.softmax(x.float(), 0), atol=1e-2 if td == torch.bfloat16 else 1e-6)
    fd = build(F.softmax_bwd_py, dt, reduction_axis=0)
    m = fd.segment()
    assert m.op == L.OP_SOFTMAX_BWD and m.flags & L.F_OUTER
    assert bool(m.flags & L.F_ROUND_ROWSUM) == (dt != DataType.Float)
    # the inner spellings keep matching without the flag
    assert not build(F.softmax_fwd_py, dt).segment().flags & L.F_OUTER
    assert not build(F.softmax_bwd_py, dt).segment().flags & L.F_OUTER


@pytest.mark.parametrize("dt", [DataType.Float, DataType.BFloat16])
@pytest.mark.parametrize("nd", [2, 4])
def test_batchnorm_fwd_matches_as_outer_norm(dt, nd):
    """benchmarks/python/normalization.py (batch norm, training, channels last): matched as the LayerNorm-type
    forward with F_OUTER; the running-statistics inputs only feed dead code and are not part of the segment."""
    import torch
    from oracle import aten_oracle as AO
    from oracle import fusion_eval

    fd = build(F.batchnorm_fwd_py, dt, num_dims=nd)
    m = fd.segment()
    assert m.op == L.OP_LAYERNORM_FWD and m.flags & L.F_OUTER and m.flags & L.F_HAS_WEIGHT and m.flags & L.F_HAS_BIAS
    assert m.outputs == ["y", "mean", "invstd"] and m.roles == {"x": 0, "w": 1, "b": 2} and m.eps == 1e-5
    td = torch.float32 if dt == DataType.Float else torch.bfloat16
    shape = [6, 5, 4, 16][4 - nd:]
    x, w, b = torch.randn(*shape).to(td), torch.randn(16).to(td), torch.randn(16).to(td)
    outs = fusion_eval.evaluate(fd, [x, w, b, torch.zeros(16), torch.ones(16)])
    for got, want in zip(outs, AO.outer_norm_fwd(x, w, b, 1e-5)):
        assert torch.equal(got, want)
    # the oracle against the explicit ATen reference (training-mode native_batch_norm on the NCHW view)
    if nd == 4:
        yn, mn, inn = torch.native_batch_norm(x.float().permute(0, 3, 1, 2), w.float(), b.float(), None, None, True,
                                               0.1, 1e-5)
        ye, me, ie = AO.outer_norm_fwd(x, w, b, 1e-5)
        assert torch.allclose(ye.float(), yn.permute(0, 2, 3, 1).to(td).float(), atol=2e-2 if td == torch.bfloat16 else 1e-5)
        assert torch.allclose(me, mn, atol=1e-6) and torch.allclose(ie, inn, rtol=1e-5)
    # the inner LayerNorm spelling is untouched
    assert not build(F.layernorm_fwd_py, dt).segment().flags & L.F_OUTER


@pytest.mark.parametrize("dt", [DataType.Float, DataType.BFloat16])
def test_batchnorm_bwd_matches_as_outer_norm(dt):
    """benchmarks/python/normalization.py:98-190: weight, mean and invstd are all [C] vectors (all Float with fp32
    io) — the roles come out of the normal-form comparison, the running statistics are dropped as dead inputs."""
    import torch
    from oracle import aten_oracle as AO
    from oracle import fusion_eval

    fd = build(F.batchnorm_bwd_py, dt, num_dims=4)
    m = fd.segment()
    assert m.op == L.OP_LAYERNORM_BWD and m.flags & L.F_OUTER and m.flags & L.F_HAS_WEIGHT
    assert m.roles == {"dy": 1, "x": 0, "mean": 5, "stat": 6, "w": 2} and m.outputs == ["dx", "dw", "db"]
    td = torch.float32 if dt == DataType.Float else torch.bfloat16
    x, dy, w = torch.randn(3, 4, 5, 16).to(td), torch.randn(3, 4, 5, 16).to(td), torch.randn(16).to(td)
    _, mean, invstd = AO.outer_norm_fwd(x, w, None, 1e-5)
    outs = fusion_eval.evaluate(fd, [x, dy, w, torch.zeros(16), torch.ones(16), mean, invstd])
    for got, want in zip(outs, AO.outer_norm_bwd(dy, x, mean, invstd, w)):
        assert torch.equal(got, want)
    r = torch.ops.aten.native_batch_norm_backward(dy.float().permute(0, 3, 1, 2), x.float().permute(0, 3, 1, 2), w.float(),
                                                   None, None, mean, invstd, True, 1e-5, [True, True, True])
    tol = dict(atol=3e-2, rtol=3e-2) if td == torch.bfloat16 else dict(atol=1e-5, rtol=1e-5)
    assert torch.allclose(outs[0].float(), r[0].permute(0, 2, 3, 1), **tol)
    assert torch.allclose(outs[1].float(), r[1], atol=0.2 if td == torch.bfloat16 else 1e-4, rtol=3e-2)
    # the inner backward spellings keep their classification
    assert not build(F.layernorm_bwd_py, dt).segment().flags & L.F_OUTER


@pytest.mark.parametrize("dt", [DataType.Float, DataType.BFloat16, DataType.Half], ids=lambda d: d.name)
def test_dropout_fusions_are_matched(dt):
    """benchmarks/python/test_dropout_{layernorm,rmsnorm}_{fwd,bwd}.py: norm(input2 + dropout(input1)) and its
    backward — the mask (uniform() < 1 - p) is a segment output forward and a Bool input backward."""
    with FusionDefinition() as fd:
        F.dropout_layernorm_fwd_py(fd, dt, 0.2)
    m = fd.segment()
    assert m.op == L.OP_LAYERNORM_FWD and m.flags & L.F_DROPOUT and m.flags & L.F_ADD_RESIDUAL
    assert m.roles == {"x": 0, "r": 1, "w": 2, "b": 3} and m.outputs == ["y", "mean", "invstd", "mask"]
    assert m.dropout_keep == 1 - 0.2 and m.eps == 1e-5
    with FusionDefinition() as fd:
        F.dropout_rmsnorm_fwd_py(fd, dt, 0.1)
    m = fd.segment()
    assert m.op == L.OP_RMSNORM_FWD and m.flags & L.F_DROPOUT and m.flags & L.F_STAT_IS_RMS
    assert m.roles == {"x": 0, "r": 1, "w": 2} and m.outputs == ["y", "mask", "rms"] and m.dropout_keep == 1 - 0.1
    with FusionDefinition() as fd:
        F.dropout_layernorm_bwd_py(fd, dt, 0.2)
    m = fd.segment()
    assert m.op == L.OP_LAYERNORM_BWD and m.flags & L.F_DROPOUT
    assert m.roles == {"dy": 3, "x": 5, "r": 6, "mask": 0, "mean": 1, "stat": 2, "w": 4}
    assert m.outputs == ["dx", "dres", "dw", "db"] and abs(m.dropout_keep - 0.8) < 1e-12
    with FusionDefinition() as fd:
        F.dropout_rmsnorm_bwd_py(fd, dt, 0.2)
    m = fd.segment()
    assert m.op == L.OP_RMSNORM_BWD and m.flags & L.F_DROPOUT and m.flags & L.F_STAT_IS_RMS
    assert not m.flags & L.F_RMS_BWD_CENTERED and m.outputs == ["dx", "dres", "dw"]
    assert m.roles == {"dy": 4, "x": 0, "r": 1, "mask": 2, "stat": 3, "w": 5}


def test_dropout_near_misses_are_refused():
    def mask_only(fd):  # the mask alone is not a normalization segment
        x = fd.define_tensor(shape=[-1, -1], dtype=DataType.Float)
        u = fd.ops.uniform(0.0, 1.0, shape=x.shape(), dtype=DataType.Float)
        fd.add_output(fd.ops.mul(x, fd.ops.cast(fd.ops.lt(u, 0.8), DataType.Float)))

    def wrong_range(fd):  # uniform(0, 2) < keep is not the dropout mask
        x1 = fd.define_tensor(shape=[-1, -1], dtype=DataType.Float)
        x2 = fd.define_tensor(shape=[-1, -1], dtype=DataType.Float)
        u = fd.ops.uniform(0.0, 2.0, shape=x1.shape(), dtype=DataType.Float)
        k = fd.ops.cast(fd.ops.lt(u, 0.8), DataType.Float)
        s = fd.ops.add(x2, fd.ops.mul(fd.ops.mul(x1, k), 1.25))
        var, mean = fd.ops.var_mean(s, dims=[1], correction=0, keepdim=True)
        fd.add_output(fd.ops.mul(fd.ops.sub(s, mean), fd.ops.rsqrt(fd.ops.add(var, 1e-5))))

    def wrong_scale(fd):  # kept elements scaled by something else than 1 / keep
        x1 = fd.define_tensor(shape=[-1, -1], dtype=DataType.Float)
        x2 = fd.define_tensor(shape=[-1, -1], dtype=DataType.Float)
        u = fd.ops.uniform(0.0, 1.0, shape=x1.shape(), dtype=DataType.Float)
        k = fd.ops.cast(fd.ops.lt(u, 0.8), DataType.Float)
        s = fd.ops.add(x2, fd.ops.mul(fd.ops.mul(x1, k), 1.5))
        var, mean = fd.ops.var_mean(s, dims=[1], correction=0, keepdim=True)
        fd.add_output(fd.ops.mul(fd.ops.sub(s, mean), fd.ops.rsqrt(fd.ops.add(var, 1e-5))))

    for f in (mask_only, wrong_range, wrong_scale):
        with FusionDefinition() as fd:
            f(fd)
        with pytest.raises(UnsupportedSegment):
            fd.segment()
    with pytest.raises(ValueError):
        with FusionDefinition() as fd:
            x = fd.define_tensor(shape=[-1, -1], dtype=DataType.Float)
            fd.ops.uniform(0.0, 1.0, shape=[4, 5], dtype=DataType.Float)


@pytest.mark.parametrize("dt", DTS, ids=lambda d: d.name)
@pytest.mark.parametrize("nd", [3, 4, 5])
@pytest.mark.parametrize("channels_last", [False, True], ids=["nchw", "nhwc"])
def test_instance_norm_fusions_are_matched(dt, nd, channels_last):
    """benchmarks/python/normalization.py with norm = "instance_norm" (test_instancenorm_{fwd,bwd}.py), both layouts,
    3-D to 5-D: the dead running-statistics updates (which add a [B, C] tensor to a [C] one) are recorded and
    ignored; the batch-norm flavour of the same definitions still goes to the outer normalization."""
    fd = build(F.norm_fwd_py, dt, "instance_norm", nd, channels_last)
    m = fd.segment()
    assert m.op == L.OP_LAYERNORM_FWD and m.flags & L.F_ROW_AFFINE and not m.flags & L.F_OUTER
    assert m.roles == {"x": 0, "w": 1, "b": 2} and m.outputs == ["y", "mean", "invstd"] and m.channels_last == channels_last
    fd = build(F.norm_bwd_py, dt, "instance_norm", nd, channels_last)
    m = fd.segment()
    assert m.op == L.OP_LAYERNORM_BWD and m.flags & L.F_ROW_AFFINE and m.channels_last == channels_last
    assert m.roles == {"dy": 1, "x": 0, "mean": 5, "stat": 6, "w": 2} and m.outputs == ["dx", "dw", "db"]
    if channels_last:
        assert build(F.norm_fwd_py, dt, "batch_norm", nd, True).segment().flags & L.F_OUTER
        assert build(F.norm_bwd_py, dt, "batch_norm", nd, True).segment().flags & L.F_OUTER

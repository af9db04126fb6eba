"""GPU tests through the reference-facing API: FusionDefinition.execute / validate on every
spelling of the hot path. The expectation is produced the way testValidate produces it
(auto-inferred: the unscheduled fusion evaluated node by node with ATen — oracle.fusion_eval)
and, like the reference's benchmark files, from explicit eager references."""
import pytest
import torch

from fuser_b200 import _lib as L
from fuser_b200.frontend import DataType, FusionDefinition, torch_dtype_to_nvfuser_dtype
from oracle import aten_oracle as AO
from oracle import fusion_eval
from tests import fusions as F

pytestmark = pytest.mark.gpu

DT = {DataType.Float: torch.float32, DataType.BFloat16: torch.bfloat16, DataType.Half: torch.float16}
SHAPES = [(64, 768), (33, 4096), (7, 1987), (16, 12288)]


def build(fn, *a, **k):
    with FusionDefinition() as fd:
        fn(fd, *a, **k)
    return fd


def auto_validate(fd, inputs):
    """testValidate with auto-inferred outputs (csrc/validator_utils.cpp:364-368)."""
    ref = fusion_eval.evaluate(fd, [t.cpu() for t in inputs])
    before = L.launch_count()
    outs = fd.validate(inputs, ref)
    assert L.launch_count() > before, "no kernel of libb200norm.so was launched"
    for o, r in zip(outs, ref):
        assert o.is_cuda and o.dtype == r.dtype and o.shape == r.shape
    return outs


@pytest.mark.parametrize("dt", list(DT), ids=lambda d: d.name)
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: f"{s[0]}x{s[1]}")
def test_benchmark_fusions_execute_and_validate(shape, dt):
    m, n = shape
    td = DT[dt]
    g = torch.Generator().manual_seed(m + n)
    x = torch.randn(m, n, generator=g).to(td).cuda()
    dy = torch.randn(m, n, generator=g).to(td).cuda()
    w = torch.randn(n, generator=g).to(td).cuda()
    b = torch.randn(n, generator=g).to(td).cuda()
    eps = 1e-5
    # forward
    y, mean, invstd = auto_validate(build(F.layernorm_fwd_py, dt), [x, w, b])
    assert mean.shape == (m,) and invstd.shape == (m, 1)
    y2, rms = auto_validate(build(F.rmsnorm_fwd_py, dt), [x, w])
    sm, = auto_validate(build(F.softmax_fwd_py, dt), [x])
    # explicit eager references, as the benchmark files validate (test_layernorm_fwd.py:84-95)
    eager = torch.nn.functional.layer_norm(x.float(), (n,), w.float(), b.float(), eps).to(td)
    AO.assert_close(y, eager.cpu(), n, "layer_norm vs eager")
    AO.assert_close(sm, torch.softmax(x.float(), -1).to(td).cpu(), n * n, "softmax vs eager")
    # backward
    auto_validate(build(F.layernorm_bwd_py, dt), [x, dy, mean, invstd, w])
    auto_validate(build(F.rmsnorm_bwd_py, dt), [x, rms, dy, w])
    auto_validate(build(F.softmax_bwd_py, dt), [sm, dy])
    # C++ builder spellings
    auto_validate(build(F.layer_norm_cpp, dt), [x, w, b])
    yc, istd = auto_validate(build(F.rms_norm_cpp, dt), [x, w])
    auto_validate(build(F.norm_bwd_cpp, dt, layer_norm=True), [dy, x, mean.unsqueeze(1), invstd, w])
    auto_validate(build(F.norm_bwd_cpp, dt, layer_norm=False), [dy, x, istd, w])
    auto_validate(build(F.softmax_cpp, dt, log=True), [x])
    auto_validate(build(F.softmax_bwd_cpp, dt), [dy, sm])


def test_executor_cache_and_3d_inputs():
    fd = build(F.layer_norm_cpp, DataType.BFloat16, ndim=3)
    w = torch.randn(1024, device="cuda").bfloat16()
    b = torch.randn(1024, device="cuda").bfloat16()
    for shape in [(4, 9, 1024), (2, 33, 1024), (4, 9, 1024)]:  # new shapes reuse the same matched segment
        x = torch.randn(*shape, device="cuda").bfloat16()
        y, mean, invstd = auto_validate(fd, [x, w, b])
        assert y.shape == x.shape and mean.shape == shape[:-1] + (1,)
    with pytest.raises(ValueError):
        fd.execute([x.float(), w, b])  # dtype differs from the definition


@pytest.mark.parametrize("dt", list(DT), ids=lambda d: d.name)
@pytest.mark.parametrize("shape", [(64, 768), (33, 4096), (7, 1987), (300, 2048)], ids=lambda s: f"{s[0]}x{s[1]}")
def test_fused_residual_add_fusions_execute_and_validate(shape, dt):
    """norm(x + residual) through FusionDefinition, expectation auto-inferred."""
    m, n = shape
    td = DT[dt]
    g = torch.Generator().manual_seed(m * 3 + n)
    x = torch.randn(m, n, generator=g).to(td).cuda()
    r = torch.randn(m, n, generator=g).to(td).cuda()
    w = torch.randn(n, generator=g).to(td).cuda()
    b = torch.randn(n, generator=g).to(td).cuda()
    auto_validate(build(F.add_layernorm_fwd_py, dt), [x, r, w, b])
    auto_validate(build(F.add_rmsnorm_fwd_py, dt), [x, r, w])
    y, s = auto_validate(build(F.rmsnorm_add_thunder, dt), [x, w, r])
    ye, se = AO.rmsnorm_add_eager(x.cpu(), w.cpu(), r.cpu())
    assert torch.equal(s.cpu(), se), "residual stream must be bit-exact"
    AO.assert_close(y, ye, n, "rmsnorm_add vs eager")


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float32], ids=["bf16", "f32"])
@pytest.mark.parametrize("shape", [(216, 1024), (300, 4096), (64, 12288), (40, 65536), (19, 1987)], ids=lambda s: f"{s[0]}x{s[1]}")
def test_thunder_rms_norm_backward_as_written(shape, dt):
    """csrc/ops/normalization.cpp:487-535 literally (projection x 1/2; used by
    tests/cpp/test_combined_inner_outer_reduction.cpp:1197): through FusionDefinition and through ops, against the
    ATen replay of the builder's text — and NOT equal to the textbook backward."""
    from fuser_b200 import ops
    m, n = shape
    g = torch.Generator().manual_seed(m + n)
    x = torch.randn(m, n, generator=g).to(dt)
    dy = torch.randn(m, n, generator=g).to(dt)
    w = torch.randn(n, generator=g).to(dt)
    _, rms = AO.rms_norm_fwd_python(x, w, 1e-5)
    dxe, dwe = AO.thunder_rms_norm_bwd(dy, x, rms, w)
    nv = torch_dtype_to_nvfuser_dtype(dt)
    with FusionDefinition() as fd:
        F.thunder_rms_norm_bwd_cpp(fd, nv)
    assert fd.segment().flags & L.F_RMS_BWD_HALF_PROJ
    dx, dw = fd.execute([dy.cuda(), x.cuda(), rms.cuda(), w.cuda()])
    torch.cuda.synchronize()
    AO.assert_close(dx, dxe, n, "thunder dx")
    AO.assert_close(dw, dwe, m, "thunder dw")
    for variant in (L.VARIANT_AUTO, L.VARIANT_GENERIC):
        dx2, dw2 = ops.rms_norm_bwd(dy.cuda(), x.cuda(), rms.cuda(), w.cuda(), stat_is_rms=True, half_projection=True,
                                    variant=variant)
        AO.assert_close(dx2, dxe, n, "thunder dx (ops)")
        AO.assert_close(dw2, dwe, m, "thunder dw (ops)")
    textbook, _ = AO.rms_norm_bwd_python(x, rms, dy, w)
    assert not torch.allclose(dxe.float(), textbook.float(), rtol=1e-2, atol=1e-2), "the two forms differ by design"


@pytest.mark.parametrize("dt", list(DT), ids=lambda d: d.name)
@pytest.mark.parametrize("shape", [(64, 768), (33, 4096), (7, 1987), (300, 2048)], ids=lambda s: f"{s[0]}x{s[1]}")
def test_dropout_fusions_execute_and_validate(shape, dt):
    """benchmarks/python/test_dropout_{layernorm,rmsnorm}_{fwd,bwd}.py through FusionDefinition with p = 0.2: the
    expectation is the fusion evaluated node by node with ATen on the SAME Philox stream (uniform() restated from
    runtime/random_numbers.cu), so the mask is compared bit for bit and the values under the validator's tolerances.
    The benchmark itself validates at p = 0 only (its uniform stream is not reproducible outside nvFuser)."""
    m, n = shape
    td = DT[dt]
    g = torch.Generator().manual_seed(m * 5 + n)
    x1 = torch.randn(m, n, generator=g).to(td).cuda()
    x2 = torch.randn(m, n, generator=g).to(td).cuda()
    dy = torch.randn(m, n, generator=g).to(td).cuda()
    w = torch.randn(n, generator=g).to(td).cuda()
    b = torch.randn(n, generator=g).to(td).cuda()
    rng = (1234 + m, 7 * n)

    def run(fd, inputs):
        ref = fusion_eval.evaluate(fd, [t.cpu() for t in inputs], rng=rng)
        before = L.launch_count()
        outs = fd.validate(inputs, ref, _rng=rng)
        assert L.launch_count() > before
        return outs

    fd = build(F.dropout_layernorm_fwd_py, dt, 0.2)
    assert fd.segment().flags & L.F_DROPOUT and fd.segment().dropout_keep == 1 - 0.2
    y, mean, invstd, mask = run(fd, [x1, x2, w, b])
    assert mask.dtype == torch.bool and 0.7 < mask.float().mean().item() < 0.9 and mean.shape == (m,) and invstd.shape == (m, 1)
    dx1, dx2, dw, db = run(build(F.dropout_layernorm_bwd_py, dt, 0.2), [mask, mean, invstd, dy, w, x1, x2])
    assert torch.equal(dx1[~mask], torch.zeros_like(dx1[~mask]))
    y, mask_r, rms = run(build(F.dropout_rmsnorm_fwd_py, dt, 0.2), [x1, x2, w])
    assert torch.equal(mask_r, mask), "same stream, same shape: the same mask"
    run(build(F.dropout_rmsnorm_bwd_py, dt, 0.2), [x1, x2, mask_r, rms, dy, w])


def test_dropout_stream_follows_the_cuda_generator():
    """Without an explicit stream the Philox arguments come from the default CUDA generator the way the reference
    takes them (csrc/rng.cpp:55-81): seed = the generator's seed, offset = its offset / 4, and the generator advances
    by 4 per launch — two launches differ, re-seeding reproduces the first."""
    dt = DataType.BFloat16
    m, n = 128, 1024
    x1, x2 = torch.randn(m, n, device="cuda").bfloat16(), torch.randn(m, n, device="cuda").bfloat16()
    w, b = torch.ones(n, device="cuda").bfloat16(), torch.zeros(n, device="cuda").bfloat16()
    fd = build(F.dropout_layernorm_fwd_py, dt, 0.3)
    gen = torch.cuda.default_generators[torch.cuda.current_device()]
    torch.cuda.manual_seed(4242)
    off0 = gen.get_offset()
    m1 = fd.execute([x1, x2, w, b])[3]
    assert fd.last_rng() == (4242, off0 // 4) and gen.get_offset() == off0 + 4
    m2 = fd.execute([x1, x2, w, b])[3]
    assert fd.last_rng() == (4242, off0 // 4 + 1) and not torch.equal(m1, m2)
    torch.cuda.manual_seed(4242)
    m3 = fd.execute([x1, x2, w, b])[3]
    assert torch.equal(m1, m3)
    assert torch.equal(m1.cpu(), AO.dropout_mask(4242, off0 // 4, (m, n), 1 - 0.3))
    # p = 0: every element kept, and the result is the fused add's (the benchmark's own validation point)
    fd0 = build(F.dropout_layernorm_fwd_py, dt, 0.0)
    y0, mean0, invstd0, mask0 = fd0.execute([x1, x2, w, b])
    assert mask0.all()
    ye, me, ie, _ = AO.add_layer_norm_fwd(x1.cpu(), x2.cpu(), w.cpu(), b.cpu(), 1e-5)
    AO.assert_close(y0, ye, n, "p = 0")


@pytest.mark.parametrize("dt", [DataType.BFloat16, DataType.Float], ids=lambda d: d.name)
@pytest.mark.parametrize("channels_last", [False, True], ids=["nchw", "nhwc"])
@pytest.mark.parametrize("size", [(4, 16, 28, 28), (2, 64, 8, 8, 8), (3, 5, 1000)], ids=lambda s: "x".join(map(str, s)))
def test_instance_norm_fusions_execute_and_validate(size, channels_last, dt):
    """test_instancenorm_{fwd,bwd}.py through FusionDefinition: auto-inferred expectation (node-by-node ATen) and the
    eager torch.nn.functional.instance_norm the benchmark validates against."""
    td = DT[dt]
    nd = len(size)
    g = torch.Generator().manual_seed(sum(size))
    at_x = torch.randn(*size, generator=g).to(td)
    at_dy = torch.randn(*size, generator=g).to(td)
    c = size[1]
    w, b = torch.randn(c, generator=g).to(td).cuda(), torch.randn(c, generator=g).to(td).cuda()
    rm, rv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
    perm = (0, *range(2, nd), 1)
    x = (at_x.permute(perm).contiguous() if channels_last else at_x).cuda()
    dy = (at_dy.permute(perm).contiguous() if channels_last else at_dy).cuda()
    fd = build(F.norm_fwd_py, dt, "instance_norm", nd, channels_last)
    y, mean, invstd = auto_validate(fd, [x, w, b, rm, rv])
    assert mean.shape == (size[0], c) and invstd.shape == (size[0], c)
    eager = torch.nn.functional.instance_norm(at_x.float(), weight=w.cpu().float(), bias=b.cpu().float())
    if channels_last:
        eager = eager.permute(perm)
    AO.assert_close(y, eager.to(td), at_x[0, 0].numel(), "instance norm vs eager")
    fdb = build(F.norm_bwd_py, dt, "instance_norm", nd, channels_last)
    auto_validate(fdb, [x, dy, w, rm, rv, mean, invstd])

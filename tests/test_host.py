"""CPU-side checks of the C++ host executor (csrc/host/norm_executor.cpp): the extension
loads, mirrors the four-method executor interface, answers `supported` from pure host logic
and refuses CPU tensors loudly (no fallback)."""
import pytest
import torch

from fuser_b200 import _host as H
from fuser_b200 import _lib as L
from fuser_b200 import ops


def test_host_extension_loads_and_has_the_executor_interface(lib):
    m = H.load()
    ex = m.NormExecutor(1, 2, 3, 4)
    assert not ex.is_compiled()
    ex.compile(m.SegmentSpec(L.OP_RMSNORM_FWD, L.F_HAS_WEIGHT, 1e-5))
    assert ex.is_compiled()
    for name in ("supported", "compile", "is_compiled", "run"):
        assert hasattr(m.NormExecutor, name)


def test_supported_is_pure_host_logic(lib):
    m = H.load()
    spec = m.SegmentSpec(L.OP_RMSNORM_FWD, L.F_HAS_WEIGHT, 1e-5)
    assert m.NormExecutor.supported(spec, torch.bfloat16, 16384, 4096)
    assert m.NormExecutor.supported(spec, torch.float32, 7, 1987)  # generic family
    assert not m.NormExecutor.supported(spec, torch.float64, 16, 64)
    forced = m.SegmentSpec(L.OP_RMSNORM_FWD, 0, 1e-5, L.VARIANT_TMA, 0)
    assert not m.NormExecutor.supported(forced, torch.bfloat16, 16, 1987)  # no 16-byte rows: TMA cannot run


def test_cpu_tensors_are_refused_not_silently_computed(lib):
    x = torch.randn(4, 64)
    with pytest.raises(L.B200NormError) as e:
        ops.rms_norm_fwd(x, None, 1e-5)
    assert e.value.code == L.ERR_NO_DEVICE
    with pytest.raises(L.B200NormError):
        ops.softmax_fwd(x)


def test_executor_cache_reuses_compiled_executors(lib):
    a = H.executor(L.OP_SOFTMAX_FWD, 0, 0.0, 0, 0)
    b = H.executor(L.OP_SOFTMAX_FWD, 0, 0.0, 0, 0)
    c = H.executor(L.OP_SOFTMAX_BWD, 0, 0.0, 0, 0)
    assert a is b and a is not c

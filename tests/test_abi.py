"""CPU tests of the C-ABI boundary: the library builds, loads, exports every symbol
include/b200norm.h declares, plans the BASELINE.json configs sensibly, and refuses to
compute without a device (there is no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

from fuser_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200norm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200norm_[a-z_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    declared = _declared_symbols()
    assert set(declared) == set(L.EXPORTED_SYMBOLS), declared
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.b200norm_abi_version() == L.ABI_VERSION


def test_struct_layouts_match_header():
    # sizes computed by hand from include/b200norm.h (LP64)
    assert ctypes.sizeof(L.Desc) == 4 * 4 + 8 + 8 + 8 + 4 + 4 + 16
    assert ctypes.sizeof(L.Tensor) == 8 + 4 + 4 + 8 * 4 + 8 * 4
    assert ctypes.sizeof(L.Plan) == 4 * 12 + 8 + 8


CONFIGS = [
    # (name, op, dtype, rows, cols, flags)
    ("C1 layernorm fwd fp32", L.OP_LAYERNORM_FWD, L.DT_F32, 2048, 4096, L.F_HAS_WEIGHT | L.F_HAS_BIAS | L.F_OUT_MEAN | L.F_OUT_INVSTD),
    ("C2 rmsnorm fwd", L.OP_RMSNORM_FWD, L.DT_BF16, 16384, 4096, L.F_HAS_WEIGHT),
    ("C2 rmsnorm bwd", L.OP_RMSNORM_BWD, L.DT_BF16, 16384, 4096, L.F_HAS_WEIGHT | L.F_OUT_DX | L.F_OUT_DW),
    ("C3 softmax fwd", L.OP_SOFTMAX_FWD, L.DT_BF16, 65536, 8192, 0),
    ("C3 softmax bwd", L.OP_SOFTMAX_BWD, L.DT_BF16, 65536, 8192, 0),
    ("C4 layernorm bwd", L.OP_LAYERNORM_BWD, L.DT_BF16, 32768, 12288, L.F_HAS_WEIGHT | L.F_OUT_DX | L.F_OUT_DW | L.F_OUT_DB),
    ("C5 rmsnorm fwd", L.OP_RMSNORM_FWD, L.DT_BF16, 4096, 65536, L.F_HAS_WEIGHT),
    ("C5 rmsnorm bwd", L.OP_RMSNORM_BWD, L.DT_BF16, 4096, 65536, L.F_HAS_WEIGHT | L.F_OUT_DX | L.F_OUT_DW),
]

# the reference's own IO-byte accounting
EXPECTED_BYTES = {
    "C1 layernorm fwd fp32": 67_158_016,
    "C2 rmsnorm fwd": 268_509_184,
    "C2 rmsnorm bwd": 402_735_104,
    "C3 softmax fwd": 2_147_483_648,
    "C3 softmax bwd": 3_221_225_472,
    "C4 layernorm bwd": 2_416_254_976,
    "C5 rmsnorm fwd": 1_073_889_280,
    "C5 rmsnorm bwd": 1_610_891_264,
}


@pytest.mark.parametrize("cfg", CONFIGS, ids=lambda c: c[0])
def test_plans_for_baseline_configs(lib, cfg):
    name, op, dt, m, n, flags = cfg
    plan = L.query(L.make_desc(op, dt, m, n, 1e-5, flags)).as_dict()
    assert plan["algorithmic_bytes"] == EXPECTED_BYTES[name]
    assert plan["variant"] in ("tma", "cluster")
    assert plan["smem_bytes"] <= 227 * 1024
    assert plan["stages"] >= 2
    assert plan["grid"] % plan["cluster_size"] == 0 and plan["grid"] <= 148
    assert plan["block"] <= 512 and plan["block"] % plan["threads_per_row"] == 0
    groups = plan["block"] // plan["threads_per_row"]
    assert plan["stages"] % groups == 0  # a stage always belongs to one consumer group
    if name.startswith("C5"):
        assert plan["variant"] == "cluster" and plan["cluster_size"] >= 2
    elif not name.startswith("C4"):
        assert plan["cluster_size"] == 1
    if op in (L.OP_RMSNORM_BWD, L.OP_LAYERNORM_BWD):
        k = 2 if op == L.OP_LAYERNORM_BWD else 1
        assert plan["workspace_bytes"] >= L.WS_HEADER_BYTES + plan["grid"] // plan["cluster_size"] * k * n * 4


def test_unaligned_and_tiny_shapes_fall_to_generic(lib):
    # N*2 bytes not a multiple of 16 -> no bulk-copy path (hidden sizes 3, 1987 of the reference's LN-bwd sweep)
    for n in (3, 1987, 4098):
        plan = L.query(L.make_desc(L.OP_LAYERNORM_BWD, L.DT_BF16, 216, n, 0.0, L.F_HAS_WEIGHT | L.F_OUT_DX | L.F_OUT_DW | L.F_OUT_DB))
        assert L.variant_name(plan.variant) == "generic"
    with pytest.raises(L.UnsupportedSegment):
        L.query(L.make_desc(L.OP_RMSNORM_FWD, L.DT_BF16, 216, 1987, 1e-5, 0, variant=L.VARIANT_TMA))


def test_bad_descriptors_are_rejected(lib):
    d = L.make_desc(L.OP_RMSNORM_FWD, L.DT_BF16, 4, 64)
    d.abi_version = 99
    with pytest.raises(L.B200NormError) as e:
        L.query(d)
    assert e.value.code == L.ERR_BAD_ARG
    with pytest.raises(L.UnsupportedSegment):
        L.query(L.make_desc(77, L.DT_BF16, 4, 64))
    with pytest.raises(L.B200NormError):
        L.query(L.make_desc(L.OP_RMSNORM_BWD, L.DT_BF16, 4, 64, 0.0, L.F_OUT_DW))  # dweight without weight
    with pytest.raises(L.B200NormError):
        L.query(L.make_desc(L.OP_RMSNORM_FWD, L.DT_BF16, 4, 64, cluster_size=3))
    for bit in (26, 31):  # flag bits the header does not define (the upper ones are the library's own launch knobs)
        with pytest.raises(L.B200NormError) as e:
            L.query(L.make_desc(L.OP_SOFTMAX_FWD, L.DT_BF16, 4, 64, 0.0, 1 << bit))
        assert e.value.code == L.ERR_BAD_ARG


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device behaviour")
def test_no_cpu_fallback(lib):
    info = L.device_info()
    assert info["has_device"] == 0 and info["sm_count"] == 148  # nominal B200 for planning only
    x = torch.zeros(4, 64, dtype=torch.bfloat16)
    y = torch.empty_like(x)
    stat = torch.empty(4, dtype=torch.float32)
    d = L.make_desc(L.OP_RMSNORM_FWD, L.DT_BF16, 4, 64, 1e-5, 0)
    mk = lambda t, dt: L.make_tensor(t.data_ptr(), dt, list(t.shape), list(t.stride()))
    with pytest.raises(L.B200NormError) as e:
        L.launch(d, [mk(x, L.DT_BF16)], [mk(y, L.DT_BF16), mk(stat, L.DT_F32)], 0, 0, 0)
    assert e.value.code == L.ERR_NO_DEVICE
    from fuser_b200 import ops

    with pytest.raises(L.B200NormError):
        ops.rms_norm_fwd(x, None, 1e-5)


def _outer_plan(op, dt, m, n, cluster=0):
    d = L.make_desc(op, dt, m, n, 0.0, L.F_OUTER, L.VARIANT_AUTO, cluster)
    return L.query(d).as_dict()


def test_outer_persistent_planner():
    """Dispatch rules of the outer-persistent softmax family (pure host logic, nominal B200): one tile of
    up to ~72 KB per CTA at three CTAs per SM, one CTA per column slab while M * W fits that tile, then the
    smallest cluster split along M that does; the extent a 16-CTA cluster cannot hold is declined (the
    reference segments such fusions, scheduler/normalization_outer.cpp:451-683 canSchedule)."""
    per_cta = 233472 // 3 - 1024
    # small reduced extent: one CTA per slab, the widest slab when N gives enough of them
    p = _outer_plan(L.OP_SOFTMAX_FWD, L.DT_BF16, 16, 4 * 73728)
    assert p["variant"] == "tma" and p["cluster_size"] == 1 and p["stages"] == 1 and p["threads_per_row"] == 32
    p = _outer_plan(L.OP_SOFTMAX_FWD, L.DT_BF16, 512, 65536)
    assert p["variant"] == "tma" and p["cluster_size"] == 1 and p["tile_rows"] == 512 and p["smem_bytes"] <= per_cta
    # the reference's outer_persistent benchmark rows {2048, 8192, 16384}: the smallest cluster that fits
    p = _outer_plan(L.OP_SOFTMAX_FWD, L.DT_BF16, 2048, 16384)
    assert p["variant"] == "cluster" and p["cluster_size"] == 2 and p["tile_rows"] == 1024 and p["smem_bytes"] <= per_cta
    p = _outer_plan(L.OP_SOFTMAX_FWD, L.DT_BF16, 8192, 18432)
    assert p["cluster_size"] == 8 and p["tile_rows"] == 1024 and p["block"] == 256 and p["smem_bytes"] <= per_cta
    assert p["grid"] % 8 == 0
    p = _outer_plan(L.OP_SOFTMAX_BWD, L.DT_BF16, 8192, 18432)  # two streams: half the rows per CTA
    assert p["cluster_size"] == 16 and p["tile_rows"] == 512 and p["smem_bytes"] <= per_cta
    p = _outer_plan(L.OP_SOFTMAX_FWD, L.DT_BF16, 16384, 8192)
    assert p["cluster_size"] == 16 and p["tile_rows"] == 1024 and p["smem_bytes"] <= per_cta
    assert p["algorithmic_bytes"] == 2 * 2 * 16384 * 8192
    p = _outer_plan(L.OP_SOFTMAX_BWD, L.DT_BF16, 16384, 8192)  # beyond three per SM: one CTA per SM
    assert p["cluster_size"] == 16 and p["smem_bytes"] <= 232448
    # few slabs: the rows are split further so that more SMs stream
    p = _outer_plan(L.OP_SOFTMAX_FWD, L.DT_BF16, 512, 4096)
    assert p["cluster_size"] * (4096 * 2 // (16 * p["threads_per_row"])) >= 148
    # forced cluster sizes are honoured
    for c in (1, 2, 4, 8, 16):
        assert _outer_plan(L.OP_SOFTMAX_FWD, L.DT_F32, 777, 384, c)["cluster_size"] == c
    # rows that are not a multiple of 16 bytes, extents beyond a 16-CTA cluster: the shape-agnostic family
    # (statistics kernel + apply kernel; the reference segments such fusions the same way)
    for m, n in ((32, 100), (1 << 18, 4096)):
        p = _outer_plan(L.OP_SOFTMAX_FWD, L.DT_BF16, m, n)
        assert p["variant"] == "generic" and p["n_launches"] == 3 and p["block"] == 128
        assert p["workspace_bytes"] >= 256 + 2 * 4 * n
    d = L.make_desc(L.OP_SOFTMAX_FWD, L.DT_BF16, 1 << 18, 4096, 0.0, L.F_OUTER, L.VARIANT_CLUSTER, 0)
    with pytest.raises(L.UnsupportedSegment):  # a forced persistent family still refuses
        L.query(d)
    assert L.query(L.make_desc(L.OP_SOFTMAX_FWD, L.DT_BF16, 512, 4096, 0.0, L.F_OUTER, L.VARIANT_GENERIC, 0)).as_dict()[
        "variant"] == "generic"
    with pytest.raises(L.UnsupportedSegment):  # other ops have no outer flavour
        _outer_plan(L.OP_RMSNORM_FWD, L.DT_BF16, 512, 4096)
    # empty problems plan to nothing
    assert _outer_plan(L.OP_SOFTMAX_FWD, L.DT_BF16, 0, 4096)["grid"] == 0


def test_outer_normalization_descs():
    """LAYERNORM_FWD / LAYERNORM_BWD with F_OUTER (batch-norm-shaped, statistics per column): planned on the
    outer-persistent family; per-column statistics in the byte count; flag combinations that make no sense
    are refused before any device work."""
    f_fwd = L.F_OUTER | L.F_HAS_WEIGHT | L.F_HAS_BIAS | L.F_OUT_MEAN | L.F_OUT_INVSTD
    p = L.query(L.make_desc(L.OP_LAYERNORM_FWD, L.DT_BF16, 8192, 18432, 1e-5, f_fwd, L.VARIANT_AUTO, 0)).as_dict()
    assert p["variant"] == "cluster" and p["cluster_size"] == 8 and p["block"] == 256 and p["n_launches"] == 1
    assert p["algorithmic_bytes"] == 2 * 2 * (8192 * 18432 + 18432) + 8 * 18432
    f_bwd = L.F_OUTER | L.F_HAS_WEIGHT | L.F_OUT_DX | L.F_OUT_DW | L.F_OUT_DB
    p = L.query(L.make_desc(L.OP_LAYERNORM_BWD, L.DT_BF16, 8192, 18432, 0.0, f_bwd, L.VARIANT_AUTO, 0)).as_dict()
    assert p["cluster_size"] == 16 and p["n_launches"] == 1  # two streams: half the rows per CTA; sums finished in-kernel
    assert p["algorithmic_bytes"] == 2 * 3 * (8192 * 18432 + 18432) + 8 * 18432
    for bad in (f_fwd | L.F_ADD_RESIDUAL, ):
        with pytest.raises(L.UnsupportedSegment):
            L.query(L.make_desc(L.OP_LAYERNORM_FWD, L.DT_BF16, 512, 4096, 1e-5, bad, L.VARIANT_AUTO, 0))
    for bad in (f_bwd | L.F_PARTIALS_F32, f_bwd | L.F_PEER_REDUCE):
        with pytest.raises(L.UnsupportedSegment):
            L.query(L.make_desc(L.OP_LAYERNORM_BWD, L.DT_BF16, 512, 4096, 0.0, bad, L.VARIANT_AUTO, 0))
    # ResNet-size N*H*W: beyond a 16-CTA cluster -> the row-block kernel (one cooperative launch, a block of rows per
    # CTA, 512 threads = 16 rows x 32 chunks); rows that are not a whole number of 16-byte chunks, or a forced generic
    # variant -> the shape-agnostic three-kernel family, many row splits for few columns
    p = L.query(L.make_desc(L.OP_LAYERNORM_FWD, L.DT_BF16, 100352, 256, 1e-5, f_fwd, L.VARIANT_AUTO, 0)).as_dict()
    assert p["variant"] == "generic" and p["n_launches"] == 1 and p["block"] == 512 and p["threads_per_row"] == 32, p
    assert p["workspace_bytes"] >= 256 + (p["grid"] * 3 + 2) * 256 * 4
    for cols, variant in ((256, L.VARIANT_GENERIC), (250, L.VARIANT_AUTO)):
        p = L.query(L.make_desc(L.OP_LAYERNORM_FWD, L.DT_BF16, 100352, cols, 1e-5, f_fwd, variant, 0)).as_dict()
        assert p["variant"] == "generic" and p["n_launches"] == 3 and p["grid"] % 2 == 0 and p["grid"] >= 2 * 256

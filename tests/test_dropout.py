"""CPU checks of the fused-dropout support: the oracle's Philox4x32-10 against the published Random123 known-answer
vectors (the reference's runtime/random_numbers.cu:8-58 is that generator), the uniform -> keep rule
(runtime/random_numbers.cu:66-88 and the float-against-double ``lt`` of
benchmarks/python/test_dropout_layernorm_fwd.py:38-41) as the two integer thresholds the kernels use, and the
descriptor rules of B200NORM_F_DROPOUT."""
import numpy as np
import pytest
import torch

from fuser_b200 import _lib as L
from oracle import aten_oracle as AO


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 with 10 rounds: counter words (c0 c1 c2 c3), key words (k0 k1) -> output
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
        ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
         (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
    ]
    for ctr, key, out in kat:
        seed = key[0] | (key[1] << 32)
        offset = ctr[0] | (ctr[1] << 32)   # the reference puts the offset in counter words 0-1 ...
        subseq = ctr[2] | (ctr[3] << 32)   # ... and the subsequence in words 2-3
        got = AO.philox4x32_10(seed, np.array([subseq], dtype=np.uint64), offset)
        assert tuple(int(w[0]) for w in got) == out


def _thresholds(keep: float):
    """The host rule of b200norm.cu (dropout_thresholds) restated in numpy."""
    raw = lambda w: np.float32(np.float32(w) * np.float32(2.0 ** -32) + np.float32(2.0 ** -33))  # noqa: E731
    a, b = 0, 1 << 32
    while a < b:
        mid = (a + b) // 2
        if raw(mid) >= np.float32(1.0):
            b = mid
        else:
            a = mid + 1
    wrap = a
    a, b = 0, wrap
    while a < b:
        mid = (a + b) // 2
        if float(raw(mid)) < keep:
            a = mid + 1
        else:
            b = mid
    return max(a - 1, 0), min(wrap, 0xFFFFFFFF)


@pytest.mark.parametrize("keep", [1.0, 0.9, 0.8, 0.5, 1.0 - 0.2, 0.1, 1e-3, 0.7999999999])
def test_keep_rule_is_two_integer_compares(keep):
    t1, t2 = _thresholds(keep)
    rs = np.random.RandomState(7)
    w = np.concatenate([rs.randint(0, 1 << 32, size=200000, dtype=np.uint64).astype(np.uint32),
                        np.array([0, 1, 127, 128, 255, 256, 0x7FFFFFFF, 0x80000000, 0xFFFFFF00, 0xFFFFFF7F, 0xFFFFFF80,
                                  0xFFFFFFFF], dtype=np.uint32),
                        np.clip(np.arange(t1 - 600, t1 + 600, dtype=np.int64), 0, 0xFFFFFFFF).astype(np.uint32),
                        np.clip(np.arange(t2 - 600, min(t2 + 600, 1 << 32), dtype=np.int64), 0, 0xFFFFFFFF).astype(np.uint32)])
    expect = AO.uniformf(w).astype(np.float64) < keep
    got = (w <= np.uint32(t1)) | (w >= np.uint32(t2))
    assert np.array_equal(expect, got)
    assert AO.uniformf(np.array([0xFFFFFFFF], dtype=np.uint32))[0] == 0.0  # the 1.0 -> 0 wrap


def test_mask_statistics_and_stream_layout():
    m = AO.dropout_mask(1234, 8, (64, 1000), 0.8)
    assert m.dtype == torch.bool and abs(m.float().mean().item() - 0.8) < 0.01
    # element i takes word i % 4 of subsequence i // 4: the first 8 elements are two consecutive Philox outputs
    w0 = np.stack(AO.philox4x32_10(1234, np.array([0, 1], dtype=np.uint64), 8), axis=1).reshape(-1)
    assert np.array_equal(m.view(-1)[:8].numpy(), AO.uniformf(w0).astype(np.float64) < 0.8)
    assert not torch.equal(m, AO.dropout_mask(1234, 9, (64, 1000), 0.8))
    assert not torch.equal(m, AO.dropout_mask(1235, 8, (64, 1000), 0.8))
    assert AO.dropout_mask(1, 0, (5, 7), 1.0).all()


def test_dropout_descriptors():
    f_fwd = L.F_HAS_WEIGHT | L.F_HAS_BIAS | L.F_OUT_MEAN | L.F_OUT_INVSTD | L.F_ADD_RESIDUAL | L.F_DROPOUT
    p = L.query(L.make_desc(L.OP_LAYERNORM_FWD, L.DT_BF16, 2048, 4096, 1e-5, f_fwd, dropout_keep=0.8)).as_dict()
    assert p["algorithmic_bytes"] == 2 * 2 * (2048 * 4096 + 4096) + 8 * 2048 + 2 * 2048 * 4096 + 2048 * 4096
    with pytest.raises(L.UnsupportedSegment):  # dropout(x) alone is not a segment of this library
        L.query(L.make_desc(L.OP_LAYERNORM_FWD, L.DT_BF16, 64, 256, 1e-5, f_fwd & ~L.F_ADD_RESIDUAL, dropout_keep=0.8))
    with pytest.raises(L.UnsupportedSegment):
        L.query(L.make_desc(L.OP_SOFTMAX_FWD, L.DT_BF16, 64, 256, 0.0, L.F_DROPOUT, dropout_keep=0.8))
    for keep in (0.0, -0.5, 1.5, float("nan")):
        with pytest.raises(L.B200NormError) as e:
            L.query(L.make_desc(L.OP_LAYERNORM_FWD, L.DT_BF16, 64, 256, 1e-5, f_fwd, dropout_keep=keep))
        assert e.value.code == L.ERR_BAD_ARG
    f_bwd = L.F_HAS_WEIGHT | L.F_OUT_DX | L.F_OUT_DW | L.F_OUT_DB | L.F_DROPOUT
    p = L.query(L.make_desc(L.OP_LAYERNORM_BWD, L.DT_BF16, 2048, 4096, 0.0, f_bwd, dropout_keep=0.8)).as_dict()
    assert p["algorithmic_bytes"] == 2 * 3 * (2048 * 4096 + 4096) + 8 * 2048 + 2 * 2 * 2048 * 4096 + 2048 * 4096
    with pytest.raises(L.UnsupportedSegment):
        L.query(L.make_desc(L.OP_LAYERNORM_BWD, L.DT_BF16, 64, 256, 0.0, f_bwd | L.F_PARTIALS_F32, dropout_keep=0.8))

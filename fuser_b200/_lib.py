"""ctypes binding of ``include/b200norm.h`` (the C-ABI drop-in boundary).

Only plain pointers and sizes cross this boundary — no torch types. The library
is loaded from ``fuser_b200/lib/libb200norm.so`` (built in-tree by
``fuser_b200._build``); a missing library is a hard error, never a fallback.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

ABI_VERSION = 1
WS_HEADER_BYTES = 256
MAX_DIMS = 4

# status codes
OK, ERR_UNSUPPORTED, ERR_BAD_ARG, ERR_NO_DEVICE, ERR_WORKSPACE, ERR_CUDA = 0, 1, 2, 3, 4, 1000

# ops
(
    OP_LAYERNORM_FWD,
    OP_LAYERNORM_BWD,
    OP_RMSNORM_FWD,
    OP_RMSNORM_BWD,
    OP_SOFTMAX_FWD,
    OP_SOFTMAX_BWD,
    OP_LOGSOFTMAX_FWD,
    OP_LOGSOFTMAX_BWD,
    OP_REDUCE_INNER,
    OP_REDUCE_OUTER,
    OP_FINALIZE_PARTIALS,
    OP_PEER_COMBINE,
    OP_NLL_SOFTMAX_BWD,
) = range(13)
OP_NAMES = [
    "layernorm_fwd",
    "layernorm_bwd",
    "rmsnorm_fwd",
    "rmsnorm_bwd",
    "softmax_fwd",
    "softmax_bwd",
    "logsoftmax_fwd",
    "logsoftmax_bwd",
    "reduce_inner",
    "reduce_outer",
    "finalize_partials",
    "peer_combine",
    "nll_softmax_bwd",
]

DT_F32, DT_BF16, DT_F16 = 0, 1, 2

F_HAS_WEIGHT = 1 << 0
F_HAS_BIAS = 1 << 1
F_STAT_IS_RMS = 1 << 2
F_RMS_BWD_CENTERED = 1 << 3
F_OUT_DX = 1 << 4
F_OUT_DW = 1 << 5
F_OUT_DB = 1 << 6
F_PARTIALS_F32 = 1 << 7
F_ROUND_ROWSUM = 1 << 8
F_REDUCE_MAX = 1 << 9
F_REDUCE_FMAX = 1 << 10
F_OUT_MEAN = 1 << 11
F_OUT_INVSTD = 1 << 12
F_ADD_RESIDUAL = 1 << 13
F_OUT_SUM = 1 << 14
F_ROUND_SUM = 1 << 15
F_PEER_REDUCE = 1 << 16
F_PEER_PUBLISH = 1 << 17
F_OUTER = 1 << 18
F_RMS_BWD_HALF_PROJ = 1 << 19
F_NO_Y = 1 << 20
F_STAT_VAR = 1 << 21
F_STAT_M2 = 1 << 22
F_DROPOUT = 1 << 23
F_REDUCE_NLL = 1 << 24
F_ROW_AFFINE = 1 << 25

VARIANT_AUTO, VARIANT_GENERIC, VARIANT_TMA, VARIANT_CLUSTER = 0, 1, 2, 3
FORM_CACHED, FORM_PIPE, FORM_PSMEM, FORM_FUSED_FINISH = 1, 2, 4, 8

EXPORTED_SYMBOLS = [
    "b200norm_abi_version",
    "b200norm_query",
    "b200norm_workspace_bytes",
    "b200norm_launch",
    "b200norm_launch_count",
    "b200norm_last_error",
    "b200norm_variant_name",
    "b200norm_device_info",
    "b200norm_register_peers",
    "b200norm_peer_exchange_bytes",
    "b200norm_debug_set_timeline",
]


class Desc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("op", C.c_int32),
        ("io_dtype", C.c_int32),
        ("variant", C.c_int32),
        ("rows", C.c_int64),
        ("cols", C.c_int64),
        ("eps", C.c_double),
        ("flags", C.c_uint32),
        ("cluster_size", C.c_int32),
        ("reserved", C.c_int32 * 4),
    ]


class Tensor(C.Structure):
    _fields_ = [
        ("data", C.c_void_p),
        ("dtype", C.c_int32),
        ("ndim", C.c_int32),
        ("sizes", C.c_int64 * MAX_DIMS),
        ("strides", C.c_int64 * MAX_DIMS),
    ]


class Plan(C.Structure):
    _fields_ = [
        ("variant", C.c_int32),
        ("cluster_size", C.c_int32),
        ("grid_x", C.c_int32),
        ("block_x", C.c_int32),
        ("threads_per_row", C.c_int32),
        ("stages", C.c_int32),
        ("tile_rows", C.c_int32),
        ("smem_bytes", C.c_int32),
        ("n_launches", C.c_int32),
        ("form", C.c_int32),
        ("groups", C.c_int32),
        ("chunks", C.c_int32),
        ("workspace_bytes", C.c_uint64),
        ("algorithmic_bytes", C.c_uint64),
    ]

    def as_dict(self) -> dict:
        return {
            "variant": variant_name(self.variant),
            "cluster_size": self.cluster_size,
            "grid": self.grid_x,
            "block": self.block_x,
            "threads_per_row": self.threads_per_row,
            "stages": self.stages,
            "tile_rows": self.tile_rows,
            "smem_bytes": self.smem_bytes,
            "n_launches": self.n_launches,
            "cached": bool(self.form & FORM_CACHED),
            "pipe": bool(self.form & FORM_PIPE),
            "psmem": bool(self.form & FORM_PSMEM),
            "groups": self.groups,
            "chunks": self.chunks,
            "workspace_bytes": int(self.workspace_bytes),
            "algorithmic_bytes": int(self.algorithmic_bytes),
        }


class B200NormError(RuntimeError):
    """Non-zero status from the kernel library (the NVF_THROW of this boundary)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"b200norm status {code}: {message}")
        self.code = code


class UnsupportedSegment(B200NormError):
    """The library declined the segment (``supported() == false`` in the reference's terms)."""


LIB_PATH = Path(__file__).resolve().parent / "lib" / "libb200norm.so"
_lib = None


def load():
    """Load the kernel library. Raises if it has not been built — no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: run `python -m fuser_b200._build` (or __graft_entry__.build()). "
            "fuser_b200 has no CPU or eager fallback."
        )
    lib = C.CDLL(str(LIB_PATH))
    lib.b200norm_abi_version.restype = C.c_int
    lib.b200norm_query.argtypes = [C.POINTER(Desc), C.POINTER(Plan)]
    lib.b200norm_query.restype = C.c_int
    lib.b200norm_workspace_bytes.argtypes = [C.POINTER(Desc)]
    lib.b200norm_workspace_bytes.restype = C.c_size_t
    lib.b200norm_launch.argtypes = [
        C.POINTER(Desc),
        C.POINTER(Tensor),
        C.c_int,
        C.POINTER(Tensor),
        C.c_int,
        C.c_void_p,
        C.c_size_t,
        C.c_void_p,
    ]
    lib.b200norm_launch.restype = C.c_int
    lib.b200norm_launch_count.restype = C.c_uint64
    lib.b200norm_last_error.restype = C.c_char_p
    lib.b200norm_variant_name.argtypes = [C.c_int]
    lib.b200norm_variant_name.restype = C.c_char_p
    lib.b200norm_device_info.argtypes = [C.POINTER(C.c_int)] * 5
    lib.b200norm_device_info.restype = C.c_int
    lib.b200norm_register_peers.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_size_t]
    lib.b200norm_register_peers.restype = C.c_int
    lib.b200norm_peer_exchange_bytes.argtypes = [C.c_int, C.c_int64]
    lib.b200norm_peer_exchange_bytes.restype = C.c_size_t
    lib.b200norm_debug_set_timeline.argtypes = [C.c_void_p]
    lib.b200norm_debug_set_timeline.restype = None
    if lib.b200norm_abi_version() != ABI_VERSION:
        raise ImportError(f"libb200norm ABI {lib.b200norm_abi_version()} != binding {ABI_VERSION}")
    _lib = lib
    return lib


def variant_name(v: int) -> str:
    return load().b200norm_variant_name(v).decode()


def last_error() -> str:
    return load().b200norm_last_error().decode()


def _raise(code: int):
    msg = last_error()
    if code == ERR_UNSUPPORTED:
        raise UnsupportedSegment(code, msg)
    raise B200NormError(code, msg)


def make_desc(op, io_dtype, rows, cols, eps=0.0, flags=0, variant=VARIANT_AUTO, cluster_size=0, correction=0,
              dropout_keep=1.0, ignore_index=-100) -> Desc:
    d = Desc()
    d.abi_version = ABI_VERSION
    d.op = op
    d.io_dtype = io_dtype
    d.variant = variant
    d.rows = rows
    d.cols = cols
    d.eps = eps
    d.flags = flags
    d.cluster_size = cluster_size
    d.reserved[0] = correction
    d.reserved[3] = ignore_index
    if flags & F_DROPOUT:  # keep probability 1 - p: the double's low and high words
        import struct

        lo, hi = struct.unpack("<ii", struct.pack("<d", float(dropout_keep)))
        d.reserved[1], d.reserved[2] = lo, hi
    return d


def query(desc: Desc) -> Plan:
    plan = Plan()
    rc = load().b200norm_query(C.byref(desc), C.byref(plan))
    if rc != OK:
        _raise(rc)
    return plan


def workspace_bytes(desc: Desc) -> int:
    return int(load().b200norm_workspace_bytes(C.byref(desc)))


def launch_count() -> int:
    return int(load().b200norm_launch_count())


def register_peers(world: int, rank: int, ptrs, nbytes: int) -> None:
    """Register the peer-memory exchange buffers of the row-sharded backward (world = 0: unregister)."""
    lib = load()
    arr = (C.c_void_p * max(1, len(ptrs)))(*[C.c_void_p(int(p)) for p in ptrs])
    rc = lib.b200norm_register_peers(world, rank, arr, nbytes)
    if rc != 0:
        raise B200NormError(rc, lib.b200norm_last_error().decode())


def peer_exchange_bytes(world: int, max_floats: int) -> int:
    """Size of one peer exchange buffer for segments of up to ``max_floats`` = K*N column sums."""
    n = int(load().b200norm_peer_exchange_bytes(world, max_floats))
    if n == 0:
        raise B200NormError(ERR_BAD_ARG, f"peer_exchange_bytes({world}, {max_floats})")
    return n


def set_timeline(device_ptr) -> None:
    """Tuning aid: per-CTA %globaltimer marks of every ring kernel go to this device buffer (None = off)."""
    load().b200norm_debug_set_timeline(C.c_void_p(device_ptr) if device_ptr else None)


def device_info() -> dict:
    vals = [C.c_int() for _ in range(5)]
    load().b200norm_device_info(*[C.byref(v) for v in vals])
    keys = ["has_device", "sm_count", "smem_per_block_optin", "cc_major", "cc_minor"]
    return {k: v.value for k, v in zip(keys, vals)}


def make_tensor(data_ptr: int, dtype: int, sizes, strides) -> Tensor:
    t = Tensor()
    t.data = data_ptr
    t.dtype = dtype
    t.ndim = len(sizes)
    for i, (s, st) in enumerate(zip(sizes, strides)):
        t.sizes[i] = s
        t.strides[i] = st
    return t


def launch(desc: Desc, inputs: list[Tensor], outputs: list[Tensor], workspace_ptr: int, workspace_nbytes: int, stream: int):
    ins = (Tensor * max(1, len(inputs)))(*inputs)
    outs = (Tensor * max(1, len(outputs)))(*outputs)
    rc = load().b200norm_launch(
        C.byref(desc), ins, len(inputs), outs, len(outputs), C.c_void_p(workspace_ptr), workspace_nbytes, C.c_void_p(stream)
    )
    if rc != OK:
        _raise(rc)

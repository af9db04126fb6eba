"""Build the in-tree sm_100a kernel library (``fuser_b200/lib/libb200norm.so``).

nvcc cross-compiles without a GPU; the built ``.so`` is git-ignored but travels
to the GPU box with the repo snapshot. There is deliberately no JIT cache and no
fallback: if the library is missing, importing the executor fails loudly.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
LIB_DIR = ROOT / "lib"
LIB_PATH = LIB_DIR / "libb200norm.so"

NVCC_FLAGS = [
    "-O3",
    "-std=c++17",
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler",
    "-fPIC",
]
OBJ_DIR = ROOT / "lib" / "obj"


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _deps() -> list[Path]:
    return list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [ROOT.parent / "include" / "b200norm.h"]


def _stale() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    return any(d.stat().st_mtime > t for d in _deps())


def _compile_one(nvcc: str, src: Path, verbose: bool) -> tuple[Path, str]:
    obj = OBJ_DIR / (src.stem + ".o")
    cmd = [nvcc, *NVCC_FLAGS, "-c", "-o", str(obj), str(src)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
    return obj, res.stderr


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a (one nvcc per translation unit, in parallel)
    and link them into one shared library."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libb200norm.so (there is no fallback path)")
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    from concurrent.futures import ThreadPoolExecutor

    srcs = sources()
    with ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 1)) as ex:
        results = list(ex.map(lambda s: _compile_one(nvcc, s, verbose), srcs))
    if verbose:
        for _, log in results:
            print(log)
    tmp = LIB_PATH.with_suffix(".so.tmp")
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(tmp), *[str(o) for o, _ in results]]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


HOST_SRC = CSRC / "host"
HOST_LIB = LIB_DIR / "_b200norm_host.so"


def _host_stale() -> bool:
    if not HOST_LIB.exists():
        return True
    t = HOST_LIB.stat().st_mtime
    deps = list(HOST_SRC.glob("*.cpp")) + list(HOST_SRC.glob("*.h")) + [ROOT.parent / "include" / "b200norm.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build_host(force: bool = False) -> Path:
    """Compile the C++ host executor (at::Tensor in/out, pybind11 surface) against this
    interpreter's torch and link it to libb200norm.so. In-tree, no JIT cache."""
    build()
    if not force and not _host_stale():
        return HOST_LIB
    import sysconfig

    import torch
    from torch.utils import cpp_extension as E

    cxx = shutil.which("g++")
    if cxx is None:
        raise RuntimeError("g++ not found: cannot build the host executor")
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    incs = E.include_paths() + [sysconfig.get_paths()["include"], os.path.join(cuda_home, "include")]
    torch_lib = E.library_paths()[0]
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    flags = ["-O2", "-std=c++17", "-fPIC", "-DTORCH_EXTENSION_NAME=_b200norm_host", "-DTORCH_API_INCLUDE_EXTENSION_H",
             f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}", "-Wno-deprecated-declarations"]
    flags += [f"-I{i}" for i in incs]
    srcs = sorted(HOST_SRC.glob("*.cpp"))

    def cc(src: Path) -> Path:
        obj = OBJ_DIR / ("host_" + src.stem + ".o")
        res = subprocess.run([cxx, *flags, "-c", "-o", str(obj), str(src)], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"g++ failed on {src}:\n{res.stdout}\n{res.stderr}")
        return obj

    from concurrent.futures import ThreadPoolExecutor

    with ThreadPoolExecutor(max_workers=len(srcs)) as ex:
        objs = list(ex.map(cc, srcs))
    tmp = HOST_LIB.with_suffix(".so.tmp")
    cmd = [cxx, "-shared", "-o", str(tmp), *map(str, objs), f"-L{torch_lib}", f"-L{LIB_DIR}", "-lb200norm",
           "-ltorch", "-ltorch_cpu", "-ltorch_cuda", "-lc10", "-lc10_cuda", "-ltorch_python",
           f"-L{os.path.join(cuda_home, 'lib64')}", "-lcudart",
           "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{torch_lib}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"host link failed:\n{' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
    os.replace(tmp, HOST_LIB)
    return HOST_LIB


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_host(force="--force" in sys.argv))

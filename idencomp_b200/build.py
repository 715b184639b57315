"""Builds the native libraries of the package in-tree (nvcc cross-compiles sm_100a without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: Path, sources) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(s).stat().st_mtime > t for s in sources)


def build_gpu(force: bool = False, verbose: bool = False) -> Path:
    """libidn_gpu.so: kernels + C-ABI."""
    out = PKG / "libidn_gpu.so"
    srcs = [CSRC / "idn_gpu.cu", *sorted(CSRC.glob("*.cuh")), ROOT / "include" / "idn_gpu.h"]
    if force or _stale(out, srcs):
        cmd = [_nvcc(), *NVCC_FLAGS, *os.environ.get("IDN_NVCC_EXTRA", "").split(), "-o", str(out), str(CSRC / "idn_gpu.cu")]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.run(cmd, check=True)
    return out


HOST_SRCS = ["model.cpp", "idn.cpp", "capi_host.cpp"]
# -ffp-contract=off / no -ffast-math: the f32 quantiser must round exactly like the reference (context.rs:346-371)
GXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-ffp-contract=off", "-fno-fast-math", "-pthread"]


def build_host(force: bool = False) -> Path:
    """libidn_host.so: C++ host mirror of the reference API; links only against the C-ABI of libidn_gpu.so."""
    out = PKG / "libidn_host.so"
    host = CSRC / "host"
    srcs = [host / s for s in HOST_SRCS]
    deps = srcs + list(host.glob("*.hpp")) + [ROOT / "include" / "idn_host.h", ROOT / "include" / "idn_gpu.h"]
    if force or _stale(out, deps):
        cxx = os.environ.get("CXX") or shutil.which("g++") or "g++"
        cmd = [cxx, *GXX_FLAGS, "-o", str(out), *map(str, srcs), f"-L{PKG}", "-lidn_gpu", "-lz", "-ldl",
               "-Wl,-rpath,$ORIGIN"]
        subprocess.run(cmd, check=True)
    return out


def build_all(force: bool = False) -> None:
    build_gpu(force)
    build_host(force)


if __name__ == "__main__":
    import sys
    build_all(force="--force" in sys.argv)

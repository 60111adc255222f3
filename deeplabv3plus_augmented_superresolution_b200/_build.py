"""Compile libasr.so (hand-written sm_100a CUDA + C ABI) in-tree with nvcc.

The .so is git-ignored but travels with the gpurun snapshot, so the GPU box runs exactly what was
cross-compiled here.  --fmad=false is part of the numerical contract (csrc/asr_common.cuh).
"""
from __future__ import annotations

import glob
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_PKG, "csrc")
LIB = os.path.join(_PKG, "libasr.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "--fmad=false", "-std=c++17",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2,-Wall", "-shared",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _deps():
    return sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        [os.path.join(os.path.dirname(_PKG), "include", "asr.h")]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in _deps())


def build(force: bool = False, verbose: bool = False, defines=(), out: str = LIB) -> str:
    """`defines`/`out` build an experimental variant next to the product library (select it with ASR_LIB=<path>)."""
    if not force and out == LIB and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], *( ["-Xptxas", "-v"] if verbose else []), "-o", out, *sources()]
    env = dict(os.environ)
    env.pop("CC", None)   # the image exports CC=/opt/gcc/bin/gcc, which lacks libgomp specs; nvcc finds gcc on PATH
    env.pop("CXX", None)
    subprocess.check_call(cmd, env=env)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

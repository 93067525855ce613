"""Builds libg1s.so (CUDA kernels + engine + C ABI) in-tree with nvcc for sm_100a.

No JIT, no torch extension machinery: the product is a plain C-ABI shared library that
the Python host mirror (and a Rust/C caller) loads.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libg1s.so")
SOURCES = ["g1s_kernels.cu", "g1s_residual.cu", "g1s_gram_imma.cu", "g1s_gram_strict.cu", "g1s_latest.cu", "g1s_filters.cu", "g1s_model.cpp", "g1s_engine.cpp", "g1s_obu.cpp"]
HEADERS = ["g1s_kernels.h", "g1s_model.h", "g1s_pool.h", "g1s_filters.h", os.path.join("..", "..", "include", "g1s.h")]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(p) > t for p in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cmd = [
        _nvcc(), "-shared", "-o", LIB,
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
        "-Xcompiler", "-fPIC,-O3,-ffp-contract=off,-fno-math-errno,-Wall,-pthread",
        "-Xptxas", "-v" if verbose else "-O3",
        "-I", os.path.join(HERE, "..", "include"),
    ] + os.environ.get("G1S_EXTRA_NVCC", "").split() + [os.path.join(CSRC, f) for f in SOURCES]  # tuning experiments
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libg1s.so")
    if verbose:
        sys.stderr.write(r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

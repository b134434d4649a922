"""In-tree nvcc build of csrc/libaps_b200.so for sm_100a (cross-compiles without a GPU).

The library links only the CUDA runtime (static), so it can be dlopen'ed by any host language.
--fmad=false plus the explicit __d*_rn intrinsics keep every fp64 operation a single IEEE
rounding, which is what makes the kernels bit-comparable with the oracle.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
OUT = os.path.join(CSRC, "libaps_b200.so")
SOURCES = ["aps_capi.cu"]


def nvcc_path() -> str:
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    deps += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    cmd = [
        nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
        "--fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", "-cudart", "static",
        "-I", INCLUDE, "-o", OUT,
    ] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd, cwd=CSRC)
    return OUT


if __name__ == "__main__":
    import sys

    print(build(force=True, verbose="-v" in sys.argv))

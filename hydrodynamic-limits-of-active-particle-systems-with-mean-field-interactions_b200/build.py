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
SOURCES = ["aps_capi.cu", "aps_fast.cu", "aps_pde.cu"]
FMA_OK = {"aps_pde.cu"}          # tolerance-parity code (IMEX PDE stepper): contraction allowed


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
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--fmad=false",
             "-Xcompiler", "-fPIC,-ffp-contract=off", "-I", INCLUDE]
    if verbose:
        flags.insert(0, "-Xptxas=-v")
    objs = [os.path.join(CSRC, s[:-3] + ".o") for s in SOURCES]
    procs = [subprocess.Popen([nvcc_path()] + [f for f in flags if not (s in FMA_OK and f == "--fmad=false")] +
                              ["-c", os.path.join(CSRC, s), "-o", o], cwd=CSRC)
             for s, o in zip(SOURCES, objs)]          # translation units compile in parallel
    if any(p.wait() != 0 for p in procs):
        raise subprocess.CalledProcessError(1, "nvcc -c")
    subprocess.check_call([nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static",
                           "-o", OUT] + objs, cwd=CSRC)
    return OUT


if __name__ == "__main__":
    import sys

    print(build(force=True, verbose="-v" in sys.argv))

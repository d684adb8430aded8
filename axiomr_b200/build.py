"""Builds axiomr_b200/libaxr_b200.so (hand-written CUDA for sm_100a + the C ABI) in-tree with nvcc.

-fmad=false / -ffp-contract=off: the path's parity contract is bit-exact coverage and depth against the CPU
reference, whose arithmetic has every + and * individually rounded, so FMA contraction is off for the
whole library. -prec-div/-prec-sqrt stay at their IEEE defaults and denormals are kept (no --use_fast_math).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "axr_api.cu")
OUT = os.path.join(HERE, "libaxr_b200.so")
DEPS = [os.path.join(HERE, "csrc", f) for f in ("axr_api.cu", "axr_kernels.cuh", "axr_raster.cuh", "axr_shaders.cuh", "axr_math.cuh", "axr_tangents.cuh")]
DEPS.append(os.path.join(HERE, "..", "include", "axr_b200.h"))

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off,-O2", "-shared", "-cudart", "static",
]


def nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.exists(c) or c == "nvcc"):
            return c
    return "nvcc"


def up_to_date() -> bool:
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(d) <= t for d in DEPS if os.path.exists(d))


def build(force: bool = False, verbose: bool = False, defines=(), out: str | None = None) -> str:
    """defines/out: tuning variants (e.g. defines=["AXR_SETUP_FPT=2"], out=".../libaxr_b200_fpt2.so"), selected at run time with
    the AXR_B200_LIB environment variable. The default build takes the constants in csrc/axr_kernels.cuh."""
    out = out or OUT
    if not force and not defines and up_to_date():
        return out
    cmd = [nvcc()] + NVCC_FLAGS + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-o", out, SRC]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libaxr_b200.so")
    return out


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(OUT)

"""Builds the kernel variants tools/ab.py compares (nvcc cross-compiles here; variants_tmp/ travels to the GPU box with gpurun).
usage: python tools/build_variants.py [name ...]   (all variants: about two minutes)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from axiomr_b200 import build as b  # noqa: E402

# Compile-time knobs of axr_kernels.cuh (launch shapes). The defaults are the winners of the round-1 A/B runs
# (profiles/r01_ab_*.jsonl); these variants bracket them.
VARIANTS = {
    "abl1": ["AXR_SETUP_ABLATE=1"],
    "abl2": ["AXR_SETUP_ABLATE=2"],
    "abl3": ["AXR_SETUP_ABLATE=3"],
    "loop0": ["AXR_SETUP_LOOP=0"],
    "loop2": ["AXR_SETUP_LOOP=2"],
    "mb10": ["AXR_SETUP_MINB=10"],
    "t64": ["AXR_SETUP_THREADS=64", "AXR_SETUP_MINB=24", "AXR_SETUP_SWZ_GROUP=256"],
    "t256": ["AXR_SETUP_THREADS=256", "AXR_SETUP_MINB=6", "AXR_SETUP_SWZ_GROUP=64"],
}

def _one(name: str) -> str:
    out = os.path.join(ROOT, "variants_tmp", f"lib_{name}.so")
    b.build(force=True, defines=VARIANTS[name], out=out, verbose="-v" in sys.argv)
    return f"{out} {VARIANTS[name]}"


if __name__ == "__main__":
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(os.path.join(ROOT, "variants_tmp"), exist_ok=True)
    with ThreadPoolExecutor(4) as ex:  # nvcc runs as a subprocess: four at a time, ~75 s each
        for line in ex.map(_one, [a for a in sys.argv[1:] if a != "-v"] or list(VARIANTS)):
            print(line, flush=True)

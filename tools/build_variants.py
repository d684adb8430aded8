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
    "recompute_sv": ["AXR_TILE_RECOMPUTE_SV=1"],
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

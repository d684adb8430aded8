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
    "setup_mb12": ["AXR_SETUP_MINB=12"],
    "setup_mb14": ["AXR_SETUP_MINB=14"],
    "setup_t256": ["AXR_SETUP_THREADS=256", "AXR_SETUP_MINB=8"],
    "setup_t64": ["AXR_SETUP_THREADS=64", "AXR_SETUP_MINB=32"],
    "setup_fpt2": ["AXR_SETUP_FPT=2"],
    "tile_128x8": ["AXR_TILE_THREADS=128", "AXR_TILE_MINB=8"],
    "tile_256x3": ["AXR_TILE_THREADS=256", "AXR_TILE_MINB=3"],
    # L2 prefetch of the index stream one / two waves of CTAs ahead (148 SMs x 16 CTAs = 2368 resident CTAs); not yet timed
    "setup_pf2368": ["AXR_SETUP_PREFETCH=2368"],
    "setup_pf4736": ["AXR_SETUP_PREFETCH=4736"],
    # programmatic dependent launch of the five draw kernels; NOT yet run on a GPU: check parity first (tools/ab.py does)
    "pdl": ["AXR_PDL=1"],
    # shading phase in two steps (resolve all pixels of a thread, then shade from shared-memory slots): bit-exact on the SIMT
    # interpreter, NOT yet timed — the first thing to A/B in the next round
    "tile_split": ["AXR_TILE_SPLIT=1"],
    "tile_split_mb5": ["AXR_TILE_SPLIT=1", "AXR_TILE_MINB=5"],
    "tile_split2": ["AXR_TILE_SPLIT=2"],  # resolve step one pixel at a time (rolled)
    "tile_split3": ["AXR_TILE_SPLIT=3"],  # 16 B shared-memory slots (24 KB per CTA instead of 40 KB), indices re-read in the shade step
    "tile_split_128x8": ["AXR_TILE_SPLIT=1", "AXR_TILE_THREADS=128", "AXR_TILE_MINB=8"],
}

def _one(name: str) -> str:
    out = os.path.join(ROOT, "variants_tmp", f"lib_{name}.so")
    b.build(force=True, defines=VARIANTS[name], out=out)
    return f"{out} {VARIANTS[name]}"


if __name__ == "__main__":
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(os.path.join(ROOT, "variants_tmp"), exist_ok=True)
    with ThreadPoolExecutor(4) as ex:  # nvcc runs as a subprocess: four at a time, ~75 s each
        for line in ex.map(_one, sys.argv[1:] or list(VARIANTS)):
            print(line, flush=True)

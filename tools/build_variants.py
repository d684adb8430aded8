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
    # measurement knobs and the alternatives that were timed in round 2 (profiles/r02_ab_summary.md); the defaults are the winners
    "setup_loop0": ["AXR_SETUP_LOOP=0"],            # closed-form coverage() per pixel in one counted loop
    "setup_loop2": ["AXR_SETUP_LOOP=2"],            # four pixels per step, branch-free
    "setup_mb10": ["AXR_SETUP_MINB=10"],            # 48 registers
    "setup_mb16": ["AXR_SETUP_MINB=16"],            # 32 registers (spills)
    "setup_noswz": ["AXR_SETUP_SWZ_K=1"],           # CTAs in face order
    "setup_abl1": ["AXR_SETUP_ABLATE=1"],           # loads + cull only (results are wrong by design: use --no-parity)
    "setup_abl2": ["AXR_SETUP_ABLATE=2"],           # + triangle setup
    "setup_abl3": ["AXR_SETUP_ABLATE=3"],           # + coverage loop without the reductions
    "setup_trim1": ["AXR_SETUP_TRIM=1"],            # float lane counter in the pixel loop
    "setup_trim2": ["AXR_SETUP_TRIM=2"],            # FMNMX3.NAN + one FSETP instead of three FSETP
    "setup_trim3": ["AXR_SETUP_TRIM=3"],
    "setup_pf0": ["AXR_SETUP_PF=0"],                # without the cross-CTA L2 prefetch of index chunks
    "setup_pf888": ["AXR_SETUP_PF=888"],
    "setup_pf3552": ["AXR_SETUP_PF=3552"],
    "tile_128x8": ["AXR_TILE_THREADS=128", "AXR_TILE_MINB=8"],
    "tile_128x6": ["AXR_TILE_THREADS=128", "AXR_TILE_MINB=6"],
    "tile_idx_pad": ["AXR_IDX_PAD=1"],
    "tile_idx_stash": ["AXR_TILE_IDX_STASH=1"],
    "tile_recompute_sv": ["AXR_TILE_RECOMPUTE_SV=1"],
    "vertex_x1": ["AXR_VERTEX_PER_THREAD=1"],
    "vertex_x4": ["AXR_VERTEX_PER_THREAD=4"],
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

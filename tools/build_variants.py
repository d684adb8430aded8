"""Builds the kernel variants tools/ab.py compares (nvcc cross-compiles here; variants_tmp/ travels to the GPU box with gpurun).
usage: python tools/build_variants.py [name ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from axiomr_b200 import build as b  # noqa: E402

F = ["AXR_SETUP_FLAT=1"]
B = F + ["AXR_SETUP_MINB=16", "AXR_TILE_THREADS=128", "AXR_TILE_MINB=8"]  # best launch shapes of batch 2
VARIANTS = {
    # warp-level post-transform vertex cache in the shading phase (axr_kernels.cuh: shade_batch_cached), 4 warps per CTA
    "vcache": ["AXR_TILE_VCACHE=1", "AXR_TILE_THREADS=128", "AXR_TILE_MINB=5"],
    "nested": ["AXR_SETUP_FLAT=0"],
    "best": B,
    "best_et": B + ["AXR_TILE_EARLY_TEX=1"],
    "best_et_mb7": F + ["AXR_SETUP_MINB=16", "AXR_TILE_THREADS=128", "AXR_TILE_MINB=7", "AXR_TILE_EARLY_TEX=1"],
    "best_et_s256": F + ["AXR_SETUP_MINB=16", "AXR_TILE_THREADS=256", "AXR_TILE_MINB=4", "AXR_TILE_EARLY_TEX=1"],
    "best_u1": B + ["AXR_FLAT_UNROLL=1"],
    "best_t256": F + ["AXR_SETUP_THREADS=256", "AXR_SETUP_MINB=8", "AXR_TILE_THREADS=128", "AXR_TILE_MINB=8"],
    "best_t64": F + ["AXR_SETUP_THREADS=64", "AXR_SETUP_MINB=32", "AXR_TILE_THREADS=128", "AXR_TILE_MINB=8"],
    "best_s128mb9": F + ["AXR_SETUP_MINB=16", "AXR_TILE_THREADS=128", "AXR_TILE_MINB=9"],
    "best_s128mb10": F + ["AXR_SETUP_MINB=16", "AXR_TILE_THREADS=128", "AXR_TILE_MINB=10"],
}

if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "variants_tmp"), exist_ok=True)
    for name in (sys.argv[1:] or VARIANTS):
        out = os.path.join(ROOT, "variants_tmp", f"lib_{name}.so")
        b.build(force=True, defines=VARIANTS[name], out=out)
        print(out, VARIANTS[name], flush=True)

"""Builds the kernel variants tools/ab.py compares (nvcc cross-compiles here; variants_tmp/ travels to the GPU box with gpurun).
usage: python tools/build_variants.py [name ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from axiomr_b200 import build as b  # noqa: E402

VARIANTS = {
    # warp-level post-transform vertex cache in the shading phase (axr_kernels.cuh: shade_batch_cached), 4 warps per CTA
    "vcache": ["AXR_TILE_VCACHE=1", "AXR_TILE_THREADS=128", "AXR_TILE_MINB=5"],
    "vcache_mb4": ["AXR_TILE_VCACHE=1", "AXR_TILE_THREADS=128", "AXR_TILE_MINB=4"],
    # plain shading phase with 4 warps per CTA (separates the effect of the CTA shape from the cache)
    "t128": ["AXR_TILE_THREADS=128", "AXR_TILE_MINB=8"],
    # single-counter coverage loop in the setup kernel (irregular meshes)
    "flat": ["AXR_SETUP_FLAT=1"],
    "flat_mb16": ["AXR_SETUP_FLAT=1", "AXR_SETUP_MINB=16"],
}

if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "variants_tmp"), exist_ok=True)
    for name in (sys.argv[1:] or VARIANTS):
        out = os.path.join(ROOT, "variants_tmp", f"lib_{name}.so")
        b.build(force=True, defines=VARIANTS[name], out=out)
        print(out, VARIANTS[name], flush=True)

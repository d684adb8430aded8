"""TEST INFRASTRUCTURE ONLY — randomized differential test: the CUDA sources under the SIMT interpreter (tests/simt) against the
oracle (oracle/axr_oracle.c), scene after scene, until the time budget is spent. Every frame must match bit for bit (coverage,
depth, colour: both sides use the same libm).

usage: AXR_SIMT_TESTS_ONLY=1 AXR_B200_LIB=tests/simt/_build/libaxr_simt.so python tests/simt/fuzz.py [--seconds 60] [--seed 0]
(tests/test_simt_kernels.py runs a short campaign; longer ones are run by hand before kernel changes are taken to the GPU)

Scenes mix: random triangle soups of every scale (sub-pixel to frame-filling, inside / crossing / outside the frustum, both
windings), meshes (icosphere, torus, grid), frame sizes that are not multiples of 16 / 32, every shader (Flat, Phong, PBR, the
discarding Cutout), both samplers, multi-draw composites onto the previous frame, screen bands, random schedules.
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from axiomr_b200 import api, scenes as S  # noqa: E402
import use_simt  # noqa: E402

use_simt.install()  # this driver only ever runs against the interpreter build (AXR_B200_LIB)
from oracle import pyoracle as po  # noqa: E402  (checker)


def random_scene(rng: np.random.Generator, i: int) -> S.Scene:
    W = int(rng.choice([33, 64, 97, 128, 161, 200, 256, 320]))
    H = int(rng.choice([17, 48, 63, 96, 100, 144, 200]))
    if rng.random() < 0.03:  # more than 256 GPU tiles across / down: the packed tile rect of the direct raster path does not apply
        W, H = (8256, 33) if rng.random() < 0.5 else (33, 8256)
    shader = int(rng.choice([S.SHADER_FLAT, S.SHADER_PHONG, S.SHADER_PBR, S.SHADER_CUTOUT]))
    sampler = int(rng.integers(0, 2))
    kind = int(rng.integers(0, 6))
    if kind == 0:    # soup, mixed scales, far beyond the frustum too
        v, f = S.random_triangles(int(rng.integers(1, 400)), int(rng.integers(1 << 30)), extent=float(rng.uniform(0.5, 8)),
                                  size=float(rng.choice([0.002, 0.02, 0.2, 1.0, 5.0])), zspread=float(rng.uniform(0.1, 6)))
    elif kind == 1:  # sub-pixel soup
        v, f = S.random_triangles(int(rng.integers(100, 3000)), int(rng.integers(1 << 30)), extent=2.5, size=0.004, zspread=1.0)
    elif kind == 2:
        v, f = S.icosphere(int(rng.integers(0, 5)), float(rng.uniform(0.3, 4.5)))
    elif kind == 3:
        n = int(rng.integers(3, 60))
        v, f = S.torus(n, int(rng.integers(3, 60)), float(rng.uniform(0.5, 3)), float(rng.uniform(0.1, 1.5)))
    elif kind == 4:
        v, f = S.quad_grid(int(rng.integers(1, 50)), float(rng.uniform(0.5, 12)), float(rng.uniform(-3, 4.95)))
    else:            # two soups of very different scale in one mesh: direct path and bins in the same draw
        v1, f1 = S.random_triangles(int(rng.integers(1, 100)), int(rng.integers(1 << 30)), size=2.0)
        v2, f2 = S.random_triangles(int(rng.integers(1, 1500)), int(rng.integers(1 << 30)), size=0.01)
        v, f = np.concatenate([v1, v2]), np.concatenate([f1, f2 + v1.shape[0]])
    if rng.random() < 0.3:  # shuffle the face order: ordinals no longer follow screen locality
        f = f[rng.permutation(f.shape[0])]
    tsz = int(rng.choice([1, 2, 7, 16, 64]))
    if shader == S.SHADER_CUTOUT:
        tex = [S.cutout_texture(max(tsz, 2), int(rng.choice([1, 2, 3]))), None, None, None, None]
    elif shader == S.SHADER_PBR:
        tex = S._pbr_textures(tsz)
    elif shader == S.SHADER_PHONG:
        tex = S._phong_textures(tsz)
    else:
        tex = [None] * 5
    eye = (float(rng.uniform(-1, 1)), float(rng.uniform(-1, 1)), float(rng.uniform(0.5, 7)))
    vp, cam = S.default_camera(W, H, eye=eye, fov=float(rng.uniform(20, 110)))
    model = S.mat_mul(S.rotate_y(float(rng.uniform(0, 6.3))), S.translate(*(rng.uniform(-1, 1, 3))))
    return S.Scene(f"fuzz{i}_k{kind}_s{shader}", W, H, v, f, shader, sampler, model=S._f32(model), view_proj=vp, cam_pos=cam,
                   textures=tex)


def main():
    # bit-identical colours need the individually rounded colour arithmetic (the default fused mode is within 1 LSB by contract)
    os.environ.setdefault("AXR_B200_COLOR_MATH", "exact")
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60.0)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    t_end = time.time() + a.seconds
    n = covered = clipped = binned = small = composites = bands = 0
    prev = {}
    while time.time() < t_end:
        sc = random_scene(rng, n)
        # composite onto the previous frame of the same size (depth test against existing contents), else onto a cleared one
        onto = prev.get((sc.height, sc.width)) if rng.random() < 0.5 else None
        band = None
        if rng.random() < 0.25 and sc.height > 32:
            y0 = int(rng.integers(0, sc.height // 16)) * 16
            band = (y0, int(rng.integers(y0 + 1, sc.height + 1)))
        c0, d0, _ = po.oracle_render(sc, threads=2, color=None if onto is None else onto[0], depth=None if onto is None else onto[1])
        c1, d1, st = api.render_scene(sc, color=None if onto is None else onto[0], depth=None if onto is None else onto[1], band=band)
        if band is not None:  # only the band's rows are produced
            if onto is None:
                ref_c, ref_d = np.zeros_like(c0), np.full_like(d0, np.inf)
                ref_c[..., 3] = 255
            else:
                ref_c, ref_d = onto[0].copy(), onto[1].copy()
            ref_c[band[0]:band[1]], ref_d[band[0]:band[1]] = c0[band[0]:band[1]], d0[band[0]:band[1]]
            c1m, d1m = c1.copy(), d1.copy()
            c1m[:band[0]], c1m[band[1]:], d1m[:band[0]], d1m[band[1]:] = ref_c[:band[0]], ref_c[band[1]:], ref_d[:band[0]], ref_d[band[1]:]
            c0, d0, c1, d1 = ref_c, ref_d, c1m, d1m
            bands += 1
        m = po.compare(c1, d1, c0, d0)
        if m["coverage_mismatch"] or m["depth_bit_mismatch"] or m["color_max_diff"]:
            print(f"MISMATCH seed={a.seed} scene #{n} {sc.name} {sc.width}x{sc.height} sampler={sc.sampler} band={band} "
                  f"composite={onto is not None} faces={sc.n_faces}: {m}", flush=True)
            sys.exit(1)
        prev[(sc.height, sc.width)] = (c0, d0)
        n += 1
        covered += m["covered"]
        clipped += st["clipped_faces"]; binned += st["binned_triangles"]; small += st["small_triangles"]
        composites += onto is not None
    print(f"FUZZ OK seed={a.seed}: {n} scenes ({composites} composites, {bands} bands), {covered} covered pixels, "
          f"{clipped} clipped faces, {small} small + {binned} binned triangles, all bit-identical to the oracle", flush=True)


if __name__ == "__main__":
    main()

"""Small scenes behind the golden fixtures. make_golden.py stores BOTH the inputs (so that libm / numpy SIMD differences
between machines cannot perturb them) and the reference's outputs in tests/golden/<case>.npz; tests read them with load_case."""
import os

import numpy as np

from axiomr_b200 import scenes as S


def _tex(n=32):
    return S._pbr_textures(n)


def cases():
    out = {}
    out["head_phong_200"] = S.Scene("head_phong_200", 200, 200, *S.head_like(12, 11), S.SHADER_PHONG,
                                    model=S._f32(S.rotate_y(0.5)), textures=_tex(64))
    out["icosphere3_flat_160x120"] = S.Scene("icosphere3_flat", 160, 120, *S.icosphere(3), S.SHADER_FLAT, model=S._f32(S.rotate_y(0.5)))
    v, f = S.random_triangles(400, 11)
    for sh, nm in ((0, "flat"), (1, "phong"), (2, "pbr")):
        out[f"random_clip_{nm}_192x144"] = S.Scene(f"random_clip_{nm}", 192, 144, v, f, sh, textures=_tex())
    v, f = S.random_triangles(60, 12, extent=6, size=4.0, zspread=6)
    out["huge_clip_pbr_101x77"] = S.Scene("huge_clip_pbr", 101, 77, v, f, S.SHADER_PBR, textures=_tex())
    v, f = S.random_triangles(1500, 13, extent=2.0, size=0.02)
    out["subpixel_phong_256x192"] = S.Scene("subpixel_phong", 256, 192, v, f, S.SHADER_PHONG, textures=_tex())
    v, f = S.torus(40, 40)
    out["torus40_pbr_240x160"] = S.Scene("torus40_pbr", 240, 160, v, f, S.SHADER_PBR, model=S._f32(S.rotate_y(0.5)), textures=_tex(64))
    # the discard branch (reference src/tiled_pipeline.cpp:571-577) through the harness's CutoutShader (oracle/ref_harness.cpp):
    # five alpha-tested layers of small triangles; nine layers of frame-filling, side-clipped quads
    out["cutout_5layers_320x240"] = S.cutout_layers()
    out["cutout_9layers_clipped_200x150"] = S.cutout_layers(w=200, h=150, layers=9, grid=2, size=8.5, tex=32)
    return out


def clip_cases():
    """One triangle per in/out pattern and plane (18 x 3 floats: Vertex 14 + clipPos 4), plus the |denom| < 1e-7 fallback."""
    rng = np.random.default_rng(5)
    tris = []
    for plane in range(6):
        for pattern in range(1, 7):  # which vertices are outside (bit mask), excluding all-in / all-out
            t = np.zeros((3, 18), dtype=np.float32)
            t[:, 0:14] = rng.uniform(-1, 1, (3, 14))
            t[:, 17] = rng.uniform(1.0, 3.0, 3)                      # w
            t[:, 14:17] = rng.uniform(-0.6, 0.6, (3, 3)) * t[:, 17:18]  # inside
            axis, sign = plane // 2, (-1.0 if plane % 2 == 0 else 1.0)
            for k in range(3):
                if pattern >> k & 1:
                    t[k, 14 + axis] = sign * t[k, 17] * rng.uniform(1.2, 3.0)
            tris.append(t)
    t = np.zeros((3, 18), dtype=np.float32)  # straddles two planes -> up to 4+ triangles
    t[:, 0:14] = rng.uniform(-1, 1, (3, 14))
    t[:, 14:18] = [[-3, -3, 0, 1], [3, -0.5, 0, 1], [0, 3, 0, 1]]
    tris.append(t)
    t = t.copy()
    t[:, 14:18] = [[0.5, 0, 0, 1], [1.0 + 2e-8, 0, 0, 1], [1.0 + 4e-8, 0.5, 0, 1]]  # denominators below 1e-7
    tris.append(t)
    return np.stack(tris)


HERE = os.path.dirname(os.path.abspath(__file__))
CASE_NAMES = ["head_phong_200", "icosphere3_flat_160x120", "random_clip_flat_192x144", "random_clip_phong_192x144",
              "random_clip_pbr_192x144", "huge_clip_pbr_101x77", "subpixel_phong_256x192", "torus40_pbr_240x160", "cutout_5layers_320x240",
              "cutout_9layers_clipped_200x150"]


def save_case(name, sc, color, depth):
    d = dict(width=sc.width, height=sc.height, shader=sc.shader, sampler=sc.sampler, vertices=sc.vertices, indices=sc.indices,
             model=sc.model, view_proj=sc.view_proj, cam_pos=sc.cam_pos, light_dir=sc.light_dir, light_color=sc.light_color,
             specular_exponent=np.float32(sc.specular_exponent), color=color, depth=depth)
    for i, t in enumerate(sc.textures):
        if t is not None:
            d[f"tex{i}"] = t
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)


def load_case(name):
    """Returns (Scene, reference colour BGRA8, reference depth)."""
    z = np.load(os.path.join(HERE, name + ".npz"))
    tex = [z[f"tex{i}"] if f"tex{i}" in z.files else None for i in range(5)]
    sc = S.Scene(name, int(z["width"]), int(z["height"]), z["vertices"], z["indices"], int(z["shader"]), int(z["sampler"]),
                 model=z["model"], view_proj=z["view_proj"], cam_pos=z["cam_pos"], light_dir=z["light_dir"],
                 light_color=z["light_color"], specular_exponent=float(z["specular_exponent"]), textures=tex)
    return sc, z["color"], z["depth"]

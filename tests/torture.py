"""Adversarial small scenes shared by the CPU (oracle vs reference) and GPU (CUDA vs oracle) edge-case tests."""
import numpy as np

from axiomr_b200 import scenes as S


def _mesh(tris_xyz, uv=None):
    """tris_xyz: (T,3,3) positions in object space -> unindexed vertices with a fixed normal / tangent frame."""
    p = np.asarray(tris_xyz, dtype=np.float64).reshape(-1, 3)
    n = p.shape[0]
    if uv is None:
        uv = np.stack([(p[:, 0] * 0.13 + 0.5) % 1.0, (p[:, 1] * 0.17 + 0.5) % 1.0], 1)
    v = S._pack(p, uv, np.tile([0, 0, 1.0], (n, 1)), np.tile([1.0, 0, 0], (n, 1)), np.tile([0, 1.0, 0], (n, 1)))
    return v, np.arange(n, dtype=np.uint32).reshape(-1, 3)


def _ortho_scene(name, tris_px, W, H, shader=0, textures=None):
    """Triangles given directly in pixel coordinates (x, y in pixels, z in NDC): view_proj maps them exactly (w = 1)."""
    t = np.asarray(tris_px, dtype=np.float64)
    vp = np.zeros((4, 4))
    vp[0][0], vp[1][1], vp[2][2], vp[3][3] = 2.0 / W, 2.0 / H, 1.0, 1.0
    vp[3][0], vp[3][1] = -1.0, -1.0
    v, f = _mesh(t)
    return S.Scene(name, W, H, v, f, shader, view_proj=S._f32(vp), cam_pos=S._f32([0, 0, 5]), textures=textures or [None] * 5)


def torture_scenes():
    tex = S._pbr_textures(16)
    out = []
    # 1. shared edges through pixel centres: two triangles of a quad whose diagonal hits every pixel centre; inclusive >= 0 covers
    #    those pixels twice, equal z -> the earlier triangle wins (reference src/tiled_pipeline.cpp:524-528,569)
    q = [[[8.5, 8.5, 0.25], [40.5, 8.5, 0.25], [40.5, 40.5, 0.25]], [[8.5, 8.5, 0.25], [40.5, 40.5, 0.25], [8.5, 40.5, 0.25]]]
    out.append(_ortho_scene("edge_ties_diagonal", q, 64, 48))
    # 2. the same triangle three times (duplicates: first wins), then a copy slightly nearer (wins everywhere)
    t = [[5.2, 3.1, 0.5], [50.7, 9.3, 0.5], [20.1, 44.9, 0.5]]
    t2 = [[a, b, 0.4999] for a, b, _ in t]
    out.append(_ortho_scene("duplicates_then_nearer", [t, t, t, t2], 64, 48, shader=1, textures=tex))
    # 3. vertices exactly on pixel centres, tile borders (multiples of 16) and the frame border; partial last tile column / row
    g = []
    for x in (0.0, 15.5, 16.0, 16.5, 31.5, 32.0, 47.5):
        for y in (0.0, 15.5, 16.0, 31.5):
            g.append([[x, y, 0.3], [x + 9.0, y, 0.3], [x, y + 7.0, 0.3]])
    out.append(_ortho_scene("tile_and_centre_aligned", g, 57, 39))
    # 4. degenerate input: zero-area, collinear, sub-1e-12 area, and a sliver one pixel tall and the whole frame wide
    d = [[[10, 10, 0.1], [10, 10, 0.1], [10, 10, 0.1]], [[5, 5, 0.1], [15, 15, 0.1], [25, 25, 0.1]],
         [[30.25, 20.25, 0.1], [30.2500001, 20.25, 0.1], [30.25, 20.2500001, 0.1]],
         [[0.0, 30.4, 0.2], [64.0, 30.6, 0.2], [0.0, 30.6, 0.2]], [[-20, -20, 0.6], [200, -20, 0.6], [-20, 200, 0.6]]]
    out.append(_ortho_scene("degenerate_and_slivers", d, 64, 48))
    # 5. perspective: vertices on and behind the camera plane (w = 0, w < 0), crossing near and far planes, huge triangle
    v, f = _mesh([[[0, 0, 5.0], [1, 0, 4.0], [0, 1, 4.0]],          # one vertex exactly at the eye (w = 0 -> clampW)
                  [[-1, -1, 6.0], [1, -1, 4.5], [0, 1, 7.0]],       # behind and in front of the camera
                  [[-500, -500, -90.0], [500, -500, -90.0], [0, 500, -120.0]],  # crosses the far plane
                  [[-3, -2, 4.95], [3, -2, 4.95], [0, 2.5, 4.8]],   # crosses the near plane (near = 0.1)
                  [[-1e4, -1e4, -3.0], [1e4, -1e4, -3.0], [0, 1e4, -3.0]]])
    out.append(S.Scene("w_zero_near_far", 96, 64, v, f, 2, textures=tex))
    # 6. z ordering with negative zero / equal depth across draws of different order: coplanar overlapping quads at z = 0 and -0
    z = [[[4, 4, 0.0], [60, 4, 0.0], [60, 44, 0.0]], [[4, 4, -0.0], [60, 44, -0.0], [4, 44, -0.0]],
         [[10, 10, -0.0], [50, 10, -0.0], [30, 40, -0.0]]]
    out.append(_ortho_scene("signed_zero_depth", z, 64, 48))
    # 7. discarding shader (harness CutoutShader): exact duplicates (equal keys up to the ordinal: each copy is peeled in turn), ties
    #    in z between a discarded and a kept fragment in both orders, signed zeros, the perspective / clip cases of scene 5
    cut = [S.cutout_texture(16, 2), None, None, None, None]
    out.append(_ortho_scene("cutout_duplicates_then_nearer", [t, t, t, t2, t], 64, 48, shader=S.SHADER_CUTOUT, textures=cut))
    shifted = [[[a + 2.0, b, c] for a, b, c in tri] for tri in q]  # same z, texture window moved by two pixels
    out.append(_ortho_scene("cutout_equal_z_ties", q + shifted + q, 64, 48, shader=S.SHADER_CUTOUT, textures=cut))
    out.append(_ortho_scene("cutout_signed_zero_depth", z + z, 64, 48, shader=S.SHADER_CUTOUT, textures=cut))
    out.append(S.Scene("cutout_w_zero_near_far", 96, 64, v, f, S.SHADER_CUTOUT, textures=cut))
    return out

"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs.

Bar (BASELINE.json north_star): coverage identical except edge ties (<= 0.01 % of pixels), depth within 1e-6 relative,
8-bit colour within 1 LSB on >= 99.9 % of pixels. Coverage and depth are in fact expected bit-exact.
"""
import numpy as np
import pytest

from axiomr_b200 import scenes as S

pytestmark = pytest.mark.gpu


def _gpu(scene, **kw):
    from axiomr_b200 import api
    return api.render_scene(scene, **kw)


def _check(po, scene, exact=True, use_ref=False, **kw):
    c1, d1, st = _gpu(scene, **kw)
    if use_ref and po.ref_available() and scene.sampler == 0:
        c0, d0, _ = po.ref_render(scene, threads=4)
    else:
        c0, d0, _ = po.oracle_render(scene, threads=8)
    m = po.compare(c1, d1, c0, d0)
    print(scene.name, st, m)
    po.assert_parity(m)
    if exact:
        assert m["coverage_mismatch"] == 0, m
        assert m["depth_bit_mismatch"] == 0, m
    return m, st


def _tex(n=64):
    return S._pbr_textures(n)


def test_c1_head_phong(po):
    m, st = _check(po, S.config1(256), use_ref=True)
    assert st["triangles"] > 0 and st["binned_triangles"] > 0


def test_c2_small_icosphere_flat(po):
    m, st = _check(po, S.config2(level=6, w=960, h=540), use_ref=True)
    assert st["small_triangles"] > 0


@pytest.mark.parametrize("shader", [0, 1, 2])
def test_random_clipped_triangles(po, shader):
    v, f = S.random_triangles(3000, 1)
    _check(po, S.Scene(f"random_clip_sh{shader}", 512, 384, v, f, shader, textures=_tex()), use_ref=True)


def test_random_small_triangles(po):
    v, f = S.random_triangles(20000, 2, extent=2.5, size=0.02)
    _check(po, S.Scene("random_small", 800, 600, v, f, 1, textures=_tex()))


def test_huge_triangles_odd_size(po):
    v, f = S.random_triangles(300, 3, extent=6, size=4.0, zspread=6)
    _check(po, S.Scene("random_huge_clip", 333, 251, v, f, 2, textures=_tex()), use_ref=True)


def test_torus_pbr(po):
    v, f = S.torus(200, 200)
    _check(po, S.Scene("torus200_pbr", 1000, 700, v, f, 2, model=S._f32(S.rotate_y(0.5)), textures=S._pbr_textures(256)))


def test_bilinear_extension(po):
    sc = S.config1(256)
    sc.sampler = S.SAMPLER_BILINEAR
    _check(po, sc)


def test_composite_two_draws(po):
    """Second draw composites through the strict depth test onto the first (mergeTileResults semantics)."""
    from axiomr_b200 import api
    a = S.config2(level=4, w=320, h=240)
    v, f = S.quad_grid(8, size=3.0, z=0.5)
    b = S.Scene("quad", 320, 240, v, f, 0)
    c0, d0, _ = po.oracle_render(a)
    c0, d0, _ = po.oracle_render(b, color=c0, depth=d0)
    c1, d1, _ = api.render_scene(a)
    c1, d1, _ = api.render_scene(b, color=c1, depth=d1)
    m = po.compare(c1, d1, c0, d0)
    po.assert_parity(m)
    assert m["coverage_mismatch"] == 0 and m["depth_bit_mismatch"] == 0


def test_empty_mesh_and_offscreen(po):
    from axiomr_b200 import api
    v, f = S.quad_grid(2)
    v[:, 0] += 100.0  # entirely outside the frustum
    sc = S.Scene("offscreen", 64, 48, v, f, 0)
    c, d, st = api.render_scene(sc)
    assert not np.isfinite(d).any() and st["triangles"] == 0
    sc2 = S.Scene("empty", 64, 48, np.zeros((0, 14), np.float32), np.zeros((0, 3), np.uint32), 0)
    c, d, st = api.render_scene(sc2)
    assert not np.isfinite(d).any()


def test_bands_equal_full_frame(po):
    """Screen-space bands (multi-GPU sharding unit) reproduce the full frame exactly."""
    from axiomr_b200 import api
    sc = S.config1(128)
    c_full, d_full, _ = api.render_scene(sc)
    c = np.zeros_like(c_full)
    d = np.full_like(d_full, np.inf)
    cc, dd = po.cleared(sc)
    for y0, y1 in ((0, 208), (208, 400), (400, 800)):
        cb, db, _ = api.render_scene(sc, band=(y0, y1))
        c[y0:y1], d[y0:y1] = cb[y0:y1], db[y0:y1]
    assert np.array_equal(c, c_full) and np.array_equal(d.view(np.uint32), d_full.view(np.uint32))

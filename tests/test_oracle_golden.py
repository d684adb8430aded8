"""CPU: the plain-C oracle (oracle/axr_oracle.c) against golden fixtures produced by the UNMODIFIED reference
(tests/golden/make_golden.py -> oracle/_ref), and directly against oracle/_ref where that library is present."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import cases as C  # noqa: E402

from axiomr_b200 import scenes as S  # noqa: E402


@pytest.mark.parametrize("name", C.CASE_NAMES)
def test_oracle_matches_reference_golden_bit_exact(po, name):
    sc, c_ref, d_ref = C.load_case(name)
    c, d, _ = po.oracle_render(sc, threads=2)
    m = po.compare(c, d, c_ref, d_ref)
    assert m["coverage_mismatch"] == 0 and m["depth_bit_mismatch"] == 0, m
    assert m["color_max_diff"] == 0, m  # same libm powf on both sides -> identical bytes


def test_oracle_thread_count_independent(po):
    sc, c_ref, d_ref = C.load_case("random_clip_phong_192x144")
    for th in (1, 5):
        c, d, _ = po.oracle_render(sc, threads=th)
        assert np.array_equal(c, c_ref) and np.array_equal(d.view(np.uint32), d_ref.view(np.uint32))


def test_stage_known_answers(po):
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "stage_kat.npz"))
    tris = C.clip_cases()
    off = 0
    for i, t in enumerate(tris):
        out = po.oracle_clip_triangle(t)
        n = int(z["clip_n"][i])
        assert out.shape[0] == n, (i, out.shape, n)
        assert np.array_equal(out.view(np.uint32), z["clip_out"][off:off + n].view(np.uint32)), i
        off += n
        back, s = po.oracle_triangle_setup(t, 640, 480)
        assert back == bool(z["setup_back"][i])
        assert np.array_equal(s.view(np.uint32), z["setup_out"][i].view(np.uint32)), i
    got = po.oracle_texture_sample(z["tex"], z["uv"], 0)
    assert np.array_equal(got.view(np.uint32), z["tex_out"].view(np.uint32))
    # clip output counts cover the interesting shapes: dropped, 1 tri, quad (2 tris), multi-plane fan
    assert set(z["clip_n"].tolist()) >= {3, 6}


def test_mat4_mul_matches_glm_order(po):
    import ctypes as ct
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "camera_kat.npz"))
    out = np.zeros(16, dtype=np.float32)
    f32p = ct.POINTER(ct.c_float)
    vp, a = np.ascontiguousarray(z["view_proj"]), np.ascontiguousarray(z["a"])
    po.oracle_lib().axo_mat4_mul(vp.ctypes.data_as(f32p), a.ctypes.data_as(f32p), out.ctypes.data_as(f32p))
    assert np.array_equal(out.reshape(4, 4).view(np.uint32), z["vp_times_a"].view(np.uint32))
    # the host-side camera mirror agrees with the reference Camera to float rounding
    vp2, _ = S.default_camera(800, 600)
    assert np.allclose(vp2, z["view_proj"], rtol=1e-6, atol=1e-6)


def test_chunked_draws_equal_one_draw(po):
    """drawing a mesh in face-order chunks composites to the same framebuffer as one draw."""
    sc, c_ref, d_ref = C.load_case("random_clip_flat_192x144")
    c, d = po.cleared(sc)
    c = c.copy()
    for f0 in range(0, sc.n_faces, 97):
        c, d, _ = po.oracle_render(sc, color=c, depth=d, first_face=f0, n_faces=min(97, sc.n_faces - f0))
    assert np.array_equal(c, c_ref) and np.array_equal(d.view(np.uint32), d_ref.view(np.uint32))


def test_bilinear_extension_reduces_to_texel_centres(po):
    """Extension sanity (no reference counterpart): at texel centres bilinear == nearest."""
    tex = S.diffuse_texture(8)
    xs = np.arange(8, dtype=np.float32) / 7.0
    uv = np.stack(np.meshgrid(xs, xs), -1).reshape(-1, 2).astype(np.float32)
    a = po.oracle_texture_sample(tex, uv, 0)
    b = po.oracle_texture_sample(tex, uv, 1)
    assert np.allclose(a, b, atol=2e-6)


@pytest.mark.skipif(not os.path.exists("/root/reference/src/tiled_pipeline.cpp"), reason="reference tree not present")
@pytest.mark.parametrize("seed", [21, 22, 23])
def test_oracle_vs_live_reference_random(po, seed):
    """Fresh random scenes straight against the unmodified reference build (only where /root/reference exists)."""
    rng = np.random.default_rng(seed)
    v, f = S.random_triangles(int(rng.integers(200, 900)), seed, extent=float(rng.uniform(1.5, 5)), size=float(rng.uniform(0.05, 2.0)))
    w, h = int(rng.integers(33, 300)), int(rng.integers(33, 200))
    sc = S.Scene("rnd", w, h, v, f, int(rng.integers(0, 3)), textures=S._pbr_textures(16))
    c0, d0, _ = po.ref_render(sc, threads=int(rng.integers(1, 5)), chunk=int(rng.integers(50, 400)))
    c1, d1, _ = po.oracle_render(sc, threads=2)
    m = po.compare(c1, d1, c0, d0)
    assert m["coverage_mismatch"] == 0 and m["depth_bit_mismatch"] == 0 and m["color_max_diff"] == 0, m


@pytest.mark.skipif(not os.path.exists("/root/reference/src/tiled_pipeline.cpp"), reason="reference tree not present")
def test_oracle_vs_live_reference_edge_cases(po):
    """Shared-edge ties, duplicates, tile/centre-aligned vertices, degenerate triangles, w = 0 / w < 0, near/far crossings, signed-zero depth."""
    from torture import torture_scenes
    for sc in torture_scenes():
        c0, d0, _ = po.ref_render(sc, threads=3, chunk=7)
        c1, d1, _ = po.oracle_render(sc, threads=2)
        m = po.compare(c1, d1, c0, d0)
        assert m["coverage_mismatch"] == 0 and m["depth_bit_mismatch"] == 0 and m["color_max_diff"] == 0, (sc.name, m)
        assert m["covered"] > 0, sc.name

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


import os as _os
import sys as _sys
_sys.path.insert(0, _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden"))
import cases as _C  # noqa: E402


@pytest.mark.parametrize("name", _C.CASE_NAMES)
def test_cuda_matches_reference_golden(po, name):
    """CUDA path vs outputs of the UNMODIFIED reference committed under tests/golden/ (inputs stored alongside)."""
    sc, c_ref, d_ref = _C.load_case(name)
    c, d, st = _gpu(sc)
    m = po.compare(c, d, c_ref, d_ref)
    po.assert_parity(m)
    assert m["coverage_mismatch"] == 0 and m["depth_bit_mismatch"] == 0, m
    assert m["color_max_diff"] <= 1, m


def test_mirror_classes_drawmesh_host_framebuffer(po):
    """The reference-shaped call: TiledPipeline(threads, camera, framebuffer).drawMesh(model, mesh) on HOST buffers."""
    from axiomr_b200 import api
    sc, c_ref, d_ref = _C.load_case("head_phong_200")
    fb = api.Framebuffer(sc.width, sc.height, True)
    fb.clearColor(api.Color(0, 0, 0, 255))
    fb.clearDepth()
    cam = api.Camera()
    cam.setViewport(0, 0, sc.width, sc.height)
    cam.setViewProjectionMatrix(sc.view_proj)
    cam._pos = sc.cam_pos
    pipe = api.TiledPipeline(8, cam, fb)
    shader = api.PhongShader(tuple(sc.light_dir), tuple(sc.light_color))
    pipe.setShader(shader)
    tex = [api.Texture(t) if t is not None else None for t in sc.textures]
    mat = api.Material("m0", tex[0], tex[2], tex[1], tex[3], tex[4], sc.specular_exponent)
    mesh = api.Mesh(sc.vertices, sc.indices, {"m0": mat})
    pipe.drawMesh(sc.model, mesh)
    m = po.compare(fb.getColorData(), fb.getDepthData(), c_ref, d_ref)
    assert m["coverage_mismatch"] == 0 and m["depth_bit_mismatch"] == 0 and m["color_max_diff"] <= 1, m
    # drawing again onto the same framebuffer changes nothing (strict depth test)
    before = fb.getColorData().copy()
    pipe.drawMesh(sc.model, mesh)
    assert np.array_equal(before, fb.getColorData())
    # error conventions: a shader whose textures are missing is an error code, not a crash
    bad = api.Mesh(sc.vertices, sc.indices, {"m0": api.Material("m0")})
    with pytest.raises(api.AxrError) as e:
        pipe.drawMesh(sc.model, bad)
    assert e.value.code == -5


def test_colour_modes(po):
    """EXACT reproduces the reference's colour bytes up to powf (<= 1 LSB, almost all equal); FAST (default) stays within 1 LSB on
    >= 99.9 % of the pixels; coverage and depth are bit-identical in both."""
    from axiomr_b200 import api
    v, f = S.random_triangles(1500, 7)
    v2, f2 = S.icosphere(5, 1.5)
    scenes = [S.Scene("soup_phong", 640, 400, np.concatenate([v, v2]), np.concatenate([f, f2 + v.shape[0]]), S.SHADER_PHONG, textures=S._phong_textures(128)),
              S.Scene("soup_pbr_bilinear", 640, 400, np.concatenate([v, v2]), np.concatenate([f, f2 + v.shape[0]]), S.SHADER_PBR, S.SAMPLER_BILINEAR, textures=_tex()),
              S.config3(n=300, w=1280, h=720, tex=512, sampler=S.SAMPLER_BILINEAR), S.config2(level=6, w=960, h=540)]
    for sc in scenes:
        c0, d0, _ = po.oracle_render(sc, threads=8)
        for mode in (api.COLOR_EXACT, api.COLOR_FAST):
            c1, d1, _ = api.render_scene(sc, color_math=mode)
            m = po.compare(c1, d1, c0, d0)
            print(sc.name, "exact" if mode == api.COLOR_EXACT else "fast", m)
            assert m["coverage_mismatch"] == 0 and m["depth_bit_mismatch"] == 0, m
            po.assert_parity(m)
            if mode == api.COLOR_EXACT:
                assert m["color_max_diff"] <= 1 and m["color_exact_frac"] > 0.999, m


def test_mesh_edited_in_place_between_draws(po):
    """The reference reads the host Mesh on every drawMesh; the adapter caches the device copy. Mesh.invalidate() after an in-place
    edit (and TiledPipeline.invalidate for a changed topology) makes the next drawMesh see the edit."""
    from axiomr_b200 import api
    v, f = S.torus(60, 40)
    sc = S.Scene("t", 320, 240, v.copy(), f, S.SHADER_FLAT, model=S._f32(S.rotate_y(0.5)))
    fb = api.Framebuffer(sc.width, sc.height, True)
    cam = api.Camera()
    cam.setViewport(0, 0, sc.width, sc.height)
    cam.setViewProjectionMatrix(sc.view_proj)
    cam._pos = sc.cam_pos
    pipe = api.TiledPipeline(8, cam, fb)
    pipe.setShader(api.FlatShader(tuple(sc.light_dir)))
    mesh = api.Mesh(sc.vertices, sc.indices)

    def draw_and_check(scene):
        fb.clearColor(api.Color(0, 0, 0, 255))
        fb.clearDepth()
        pipe.drawMesh(scene.model, mesh)
        c0, d0, _ = po.oracle_render(scene, threads=4)
        m = po.compare(fb.getColorData(), fb.getDepthData(), c0, d0)
        assert m["coverage_mismatch"] == 0 and m["depth_bit_mismatch"] == 0 and m["color_max_diff"] <= 1, m

    draw_and_check(sc)
    mesh.getVertices()[:, 0:3] *= np.float32(0.5)  # edit in place: same counts, same faces
    mesh.invalidate()
    sc.vertices = mesh.getVertices()
    draw_and_check(sc)
    # topology change behind the same Mesh object: every other face dropped
    mesh._f = np.ascontiguousarray(mesh.getFaces()[::2])
    mesh._groups = [api.MaterialGroup(mesh.getMaterialGroups()[0].materialName, 0, mesh._f.shape[0])]
    mesh.invalidate()
    sc.indices = mesh.getFaces()
    draw_and_check(sc)


def test_draw_without_bins_is_redone_when_bins_are_needed(po):
    """A mesh that binned nothing is drawn without the two bin kernels the next time; if the camera then makes its triangles large
    enough to need the bins, the draw is re-issued with them (OVF_NEED_BINS) and the frame is still right."""
    from axiomr_b200 import api
    v, f = S.icosphere(5, 1.0)
    far = S.Scene("far", 512, 384, v, f, S.SHADER_FLAT, model=S._f32(S.rotate_y(0.5)))
    dev = api.Device(far.width, far.height)
    try:
        mesh = dev.load_scene(far)
        for _ in range(2):
            dev.clear()
            dev.draw_mesh(mesh, far.model)
        st = dev.stats()
        assert st["binned_triangles"] == 0 and st["kernel_launches"] == 5, st   # vertex, setup, fold, tile, clipped
        near = S.Scene("near", 512, 384, v, f, S.SHADER_FLAT, model=S._f32(S.rotate_y(0.5)))
        near.view_proj, near.cam_pos = S.default_camera(512, 384, eye=(0.0, 0.0, 1.6))
        dev.set_uniforms(near.view_proj, near.cam_pos)
        dev.clear()
        dev.draw_mesh(mesh, near.model)
        c1, d1 = dev.resolve()
        st = dev.stats()
        assert st["binned_triangles"] > 0 and st["redo"] == 1, st
        c0, d0, _ = po.oracle_render(near, threads=4)
        m = po.compare(c1, d1, c0, d0)
        assert m["coverage_mismatch"] == 0 and m["depth_bit_mismatch"] == 0 and m["color_max_diff"] <= 1, m
    finally:
        dev.close()


def test_few_large_triangles_are_scanned_by_the_tiles_on_the_next_draw(po):
    """A mesh whose last draw sent only a few triangles through the bins is drawn without the two bin kernels (BINS_SCAN: every tile
    with a non-zero count tests all records itself); the frames equal the oracle, also when clipped, huge and small triangles mix,
    and a later draw with too many records for that mode is re-issued with the reference lists."""
    from axiomr_b200 import api
    v, f = S.random_triangles(300, 3, extent=6, size=4.0, zspread=6)      # huge + clipped triangles
    v2, f2 = S.random_triangles(1500, 7)                                   # mixed sizes
    sc = S.Scene("scan_mode", 777, 333, np.concatenate([v, v2]), np.concatenate([f, f2 + v.shape[0]]), S.SHADER_PHONG, textures=_tex())
    c0, d0, _ = po.oracle_render(sc, threads=8)
    dev = api.Device(sc.width, sc.height)
    try:
        mesh = dev.load_scene(sc)
        launches = []
        for i in range(3):
            dev.clear()
            dev.draw_mesh(mesh, sc.model)
            c1, d1 = dev.resolve()
            st = dev.stats()
            launches.append(st["kernel_launches"])
            m = po.compare(c1, d1, c0, d0)
            assert m["coverage_mismatch"] == 0 and m["depth_bit_mismatch"] == 0 and m["color_max_diff"] <= 1, (i, m)
            assert 0 < st["binned_triangles"] <= 2048, st
        assert launches[0] == launches[1] + 2 == launches[2] + 2, launches    # scan + scatter kernels gone after the first draw
        # many more records than the scan mode takes: zoom in until most of the 20 k triangles of a second mesh are large
        vb, fb = S.random_triangles(30000, 11, extent=1.0, size=0.12, zspread=0.5)
        big = S.Scene("scan_to_lists", sc.width, sc.height, vb, fb, S.SHADER_FLAT)
        big.view_proj, big.cam_pos = S.default_camera(sc.width, sc.height, eye=(0.0, 0.0, 1.5))
        far_vp, far_cam = S.default_camera(sc.width, sc.height, eye=(0.0, 0.0, 90.0))
        mesh2 = dev.load_scene(big)
        dev.set_uniforms(far_vp, far_cam)
        for _ in range(2):
            dev.clear()
            dev.draw_mesh(mesh2, big.model)
        st = dev.stats()
        assert st["binned_triangles"] <= 2048, st
        dev.set_uniforms(big.view_proj, big.cam_pos)
        dev.clear()
        dev.draw_mesh(mesh2, big.model)
        c1, d1 = dev.resolve()
        st = dev.stats()
        assert st["binned_triangles"] > 4096 and st["redo"] == 1, st
        cb, db, _ = po.oracle_render(big, threads=8)
        m = po.compare(c1, d1, cb, db)
        assert m["coverage_mismatch"] == 0 and m["depth_bit_mismatch"] == 0 and m["color_max_diff"] <= 1, m
    finally:
        dev.close()


def test_overflowed_draw_is_redone_with_the_state_it_was_issued_with(po):
    """A draw whose bins overflow is re-issued by the next API call; a set_uniforms / set_shader in between must not leak into it."""
    from axiomr_b200 import api
    v, f = S.random_triangles(300000, 31, extent=1.2, size=0.3, zspread=0.5)
    sc = S.Scene("many_mid", 1024, 768, v, f, 0)
    dev = api.Device(sc.width, sc.height)
    try:
        mesh = dev.load_scene(sc)
        dev.clear()
        dev.draw_mesh(mesh, sc.model)          # overflows (not yet noticed by the host)
        other_vp, other_cam = S.default_camera(sc.width, sc.height, eye=(3.0, 1.0, 4.0))
        dev.set_uniforms(other_vp, other_cam)   # must first redo the pending draw with the old camera
        dev.set_shader(api.SHADER_FLAT, (0.0, 0.0, -1.0))
        c1, d1 = dev.resolve()
        assert dev.stats()["redo"] == 1
        c0, d0, _ = po.oracle_render(sc, threads=8)
        m = po.compare(c1, d1, c0, d0)
        assert m["coverage_mismatch"] == 0 and m["depth_bit_mismatch"] == 0 and m["color_max_diff"] <= 1, m
    finally:
        dev.close()


def test_last_tile_of_an_8192_frame(po):
    """A small triangle wholly inside GPU tile (255, 255) of an 8192 x 8192 frame: its packed tile rect equals the NO_TOUCH sentinel
    unless the packed path is restricted to fewer than 256 tiles per axis."""
    if _os.environ.get("AXR_SIMT_TESTS_ONLY") == "1":
        pytest.skip("8192 x 8192 frame: GPU only")
    from axiomr_b200 import api
    W = H = 8192
    # identity view-projection: object space is NDC, so the three points land in the last 32 x 32 tile
    def ndc(px, py, z=0.5):
        return np.array([px / W * 2 - 1, py / H * 2 - 1, z], dtype=np.float32)
    pts = [ndc(8170.2, 8168.3), ndc(8180.7, 8169.1), ndc(8174.4, 8181.6)]
    v = np.zeros((3, 14), dtype=np.float32)
    for i, p in enumerate(pts):
        v[i, 0:3] = p
        v[i, 5:8] = (0, 0, 1)
    f = np.array([[0, 1, 2], [0, 2, 1]], dtype=np.uint32)
    sc = S.Scene("corner", W, H, v, f, S.SHADER_FLAT, model=np.eye(4, dtype=np.float32))
    sc.view_proj, sc.cam_pos = np.eye(4, dtype=np.float32), np.array([0, 0, 5], dtype=np.float32)
    c1, d1, st = api.render_scene(sc)
    c0, d0, _ = po.oracle_render(sc, threads=8)
    m = po.compare(c1, d1, c0, d0)
    assert m["covered"] > 20 and m["coverage_mismatch"] == 0 and m["depth_bit_mismatch"] == 0 and m["color_max_diff"] <= 1, m
    assert np.isfinite(d1[8160:, 8160:]).sum() == m["covered"]


def test_bin_overflow_regrows_and_redraws(po):
    """More binned triangles / references than the initial bin capacity: the draw is re-issued after growing, output unchanged."""
    from axiomr_b200 import api
    v, f = S.random_triangles(300000, 31, extent=1.2, size=0.3, zspread=0.5)
    sc = S.Scene("many_mid", 1024, 768, v, f, 0)
    c1, d1, st = api.render_scene(sc)
    assert st["binned_triangles"] > (1 << 18) or st["bin_refs"] > (1 << 20), st
    assert st["redo"] == 1, st
    c0, d0, _ = po.oracle_render(sc, threads=8)
    m = po.compare(c1, d1, c0, d0)
    assert m["coverage_mismatch"] == 0 and m["depth_bit_mismatch"] == 0, m


def test_multi_material_groups(po):
    """Two material groups with different textures: each face is shaded with its own group's material."""
    from axiomr_b200 import api
    v1, f1 = S.quad_grid(4, size=2.0, z=0.0)
    v2, f2 = S.quad_grid(4, size=2.0, z=0.0)
    v1[:, 0] -= 1.1
    v2[:, 0] += 1.1
    verts = np.concatenate([v1, v2])
    faces = np.concatenate([f1, f2 + v1.shape[0]])
    ta, tb = S._phong_textures(32), [S.scalar_texture(32, 0, 255, 2), S.normal_texture(32, 3), None, None, None]
    dev = api.Device(320, 200)
    try:
        mesh = dev.upload_mesh(verts, faces, groups=[(0, f1.shape[0]), (f1.shape[0], f2.shape[0])])
        h = [dev.upload_texture(t) for t in (ta[0], ta[1], tb[0], tb[1])]
        dev.set_material(mesh, 0, h[0], h[1], specular_exponent=0.2)
        dev.set_material(mesh, 1, h[2], h[3], specular_exponent=0.4)
        sa = S.Scene("a", 320, 200, verts, faces[:f1.shape[0]], 1, textures=ta, specular_exponent=0.2)
        sb = S.Scene("b", 320, 200, verts, faces[f1.shape[0]:], 1, textures=tb, specular_exponent=0.4)
        dev.set_uniforms(sa.view_proj, sa.cam_pos)
        dev.set_shader(1, sa.light_dir, sa.light_color)
        dev.clear()
        dev.draw_mesh(mesh, sa.model)
        c1, d1 = dev.resolve()
    finally:
        dev.close()
    c0, d0, _ = po.oracle_render(sa)
    c0, d0, _ = po.oracle_render(sb, color=c0, depth=d0)
    m = po.compare(c1, d1, c0, d0)
    assert m["coverage_mismatch"] == 0 and m["depth_bit_mismatch"] == 0 and m["color_max_diff"] <= 1, m


@pytest.mark.parametrize("shader", [0, 1, 2])
def test_cpp_dropin_adapter_vs_reference_in_process(tmp_path, shader):
    """AR::B200TiledPipeline (axiomr_b200/host, C++ adapter over the C ABI) against the reference's AR::TiledPipeline in ONE
    process, scene loaded by the reference's own OBJ/MTL/texture loaders (oracle/_ref/dropin_demo, prebuilt where /root/reference exists)."""
    import subprocess
    from objutil import write_obj_scene
    exe = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "oracle", "_ref", "dropin_demo")
    if not _os.path.exists(exe):
        pytest.skip("oracle/_ref/dropin_demo not built (needs /root/reference at build time)")
    v, f = S.head_like(24, 23)
    obj = write_obj_scene(str(tmp_path), "head", v, f, S._pbr_textures(64))
    r = subprocess.run([exe, obj, "400", "300", str(shader)], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0 and "PARITY OK" in r.stdout, (r.stdout, r.stderr)


def test_cpp_dropin_user_shader_plugin_vs_reference_running_the_same_ishader(tmp_path):
    """A shader of the user's own, end to end against the real thing: in ONE process the unmodified reference pipeline runs an
    IShader subclass written against its plugin contract (LambertTintShader in tests/dropin/dropin_demo.cpp) and
    AR::B200TiledPipeline runs the device functor compiled from the same formulas (tests/plugins/lambert_tint.cu), handed to
    Pipeline::setShader as an AR::B200PluginShader; the two host framebuffers are compared."""
    import subprocess
    import sys
    from objutil import write_obj_scene
    if _os.environ.get("AXR_SIMT_TESTS_ONLY") == "1":
        pytest.skip("plug-ins are nvcc-built device code")
    root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
    exe = _os.path.join(root, "oracle", "_ref", "dropin_demo")
    if not _os.path.exists(exe):
        pytest.skip("oracle/_ref/dropin_demo not built (needs /root/reference at build time)")
    sys.path.insert(0, _os.path.join(root, "tools"))
    try:
        from build_shader_plugin import build_plugin
    finally:
        sys.path.pop(0)
    plugin = build_plugin(_os.path.join(root, "tests", "plugins", "lambert_tint.cu"))
    v, f = S.head_like(24, 23)
    obj = write_obj_scene(str(tmp_path), "head", v, f, S._pbr_textures(64))
    r = subprocess.run([exe, obj, "400", "300", "3", plugin], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0 and "PARITY OK" in r.stdout, (r.stdout, r.stderr)


# ---------------------------------------------------------------------------------------------- BASELINE.json full-size configs
def _full_size(po, sc, min_cov):
    """Full-size config against the CPU checker (the unmodified reference when present and the mode is nearest), in both colour
    modes: FAST (the default, what bench.py times) within the north_star tolerance, EXACT within 1 LSB on every pixel."""
    from axiomr_b200 import api
    if po.ref_available() and sc.sampler == 0:
        c0, d0, secs = po.ref_render(sc, threads=min(32, _os.cpu_count() or 1))
    else:
        c0, d0, secs = po.oracle_render(sc, threads=_os.cpu_count() or 1)
    out = {}
    for mode, name in ((api.COLOR_FAST, "fast"), (api.COLOR_EXACT, "exact")):
        c1, d1, st = _gpu(sc, color_math=mode)
        m = po.compare(c1, d1, c0, d0)
        print(sc.name, name, st, m, f"cpu {secs:.2f}s")
        po.assert_parity(m)
        assert m["coverage_mismatch"] == 0 and m["depth_bit_mismatch"] == 0, m
        assert m["covered"] > min_cov
        if mode == api.COLOR_EXACT:
            assert m["color_max_diff"] <= 1, m
        out[name] = m
    return out


def test_full_config1_head_phong_800(po):
    _full_size(po, S.config1(), 100000)


def test_full_config2_icosphere8_flat_1080p(po):
    _full_size(po, S.config2(), 500000)


def test_full_config3_torus_10m_phong_4k_nearest(po):
    _full_size(po, S.config3(sampler=S.SAMPLER_NEAREST), 2000000)


def test_full_config3_bilinear_extension_4k(po):
    _full_size(po, S.config3(sampler=S.SAMPLER_BILINEAR), 2000000)


def test_full_config4_subpixel_20m_flat_4k(po):
    _full_size(po, S.config4(), 1000000)


def test_size_independent_properties_full_c3(po):
    """Properties that need no oracle at full size: idempotence under redraw (strict depth test), chunked == single draw,
    permuting the face order changes nothing but tie-breaks (none here: depth and coverage identical)."""
    from axiomr_b200 import api
    sc = S.config3(sampler=S.SAMPLER_NEAREST)
    dev = api.Device(sc.width, sc.height)
    try:
        mesh = dev.load_scene(sc)
        dev.clear()
        dev.draw_mesh(mesh, sc.model)
        c1, d1 = dev.resolve()
        dev.draw_mesh(mesh, sc.model)          # idempotent: nothing passes z < fbZ the second time
        c2, d2 = dev.resolve()
        assert np.array_equal(c1, c2) and np.array_equal(d1.view(np.uint32), d2.view(np.uint32))
        half = sc.n_faces // 2                   # two chunks composited == one draw
        ma = dev.upload_mesh(sc.vertices, sc.indices[:half])
        mb = dev.upload_mesh(sc.vertices, sc.indices[half:])
        tex = [dev.upload_texture(t) for t in sc.textures[:2]]
        for h in (ma, mb):
            dev.set_material(h, 0, tex[0], tex[1], specular_exponent=sc.specular_exponent)
        dev.clear()
        dev.draw_mesh(mb, sc.model)              # reversed chunk order: winners are decided by z, ties by order within a draw
        dev.draw_mesh(ma, sc.model)
        c3, d3 = dev.resolve()
        assert np.array_equal(np.isfinite(d1), np.isfinite(d3))
        assert np.count_nonzero(d1.view(np.uint32) != d3.view(np.uint32)) <= d1.size // 10000   # exact z ties between chunks only
    finally:
        dev.close()


def test_multi_gpu_composites_match_single_gpu():
    """bands (peer stores / NCCL) and views composites on GPU 0 == single-GPU renders. Needs >= 2 GPUs (skipped otherwise)."""
    import subprocess
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    n = 2 if n < 4 else 4
    here = _os.path.dirname(_os.path.abspath(__file__))
    r = subprocess.run([_sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                        "--master-port", "29577", _os.path.join(here, "multi_gpu_check.py")], capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0 and "MULTI_GPU_CHECK OK" in r.stdout


def test_overlapped_draws_identical(po):
    """axr_set_overlap(1): geometry of draw i+1 runs beside the tile kernel of draw i (two buffer sets, two streams). A sequence of
    clears and draws of two meshes must give bit-identical frames to the serial mode."""
    from axiomr_b200 import api
    a = S.config2(level=5, w=640, h=360)
    v, f = S.torus(120, 90)
    b = S.Scene("t", 640, 360, v, f, S.SHADER_FLAT, model=S._f32(S.rotate_y(0.5)))
    frames = {}
    for mode in (False, True):
        dev = api.Device(640, 360)
        try:
            dev.set_overlap(mode)
            ma = dev.load_scene(a)
            mb = dev.upload_mesh(b.vertices, b.indices)
            out = []
            for k in range(6):
                dev.clear()
                for _ in range(3):  # redundant redraws: later ones must not change anything
                    dev.draw_mesh(ma, a.model)
                    dev.draw_mesh(mb, S._f32(S.rotate_y(0.3 * k)))
                out.append(dev.resolve())
            frames[mode] = out
        finally:
            dev.close()
    for (c0, d0), (c1, d1) in zip(frames[False], frames[True]):
        assert np.array_equal(c0, c1) and np.array_equal(d0.view(np.uint32), d1.view(np.uint32))
    # and the serial frames are right
    c, d, _ = po.oracle_render(a)
    b2 = S.Scene("t", 640, 360, v, f, S.SHADER_FLAT, model=S._f32(S.rotate_y(0.0)))
    c, d, _ = po.oracle_render(b2, color=c, depth=d)
    m = po.compare(frames[True][0][0], frames[True][0][1], c, d)
    assert m["coverage_mismatch"] == 0 and m["depth_bit_mismatch"] == 0, m


def test_edge_cases_torture(po):
    """CUDA vs the CPU checker on adversarial inputs: shared-edge ties through pixel centres, duplicate triangles, vertices on tile
    borders / pixel centres / the frame border with partial tiles, zero-area and sub-1e-12-area triangles, frame-wide slivers,
    vertices at w = 0 and behind the camera, near / far crossings, huge triangles, +0 / -0 depth ties."""
    from torture import torture_scenes
    for sc in torture_scenes():
        c1, d1, st = _gpu(sc)
        if po.ref_available():
            c0, d0, _ = po.ref_render(sc, threads=2, chunk=5)
        else:
            c0, d0, _ = po.oracle_render(sc, threads=2)
        m = po.compare(c1, d1, c0, d0)
        print(sc.name, st, m)
        assert m["coverage_mismatch"] == 0 and m["depth_bit_mismatch"] == 0 and m["color_max_diff"] <= 1, (sc.name, m)


def test_gpu_tangent_generation_matches_reference_loader_arithmetic():
    """axr_generate_tangents (Mesh::calculateTangentBitangent on the device) against the numpy mirror of the reference loader
    (axiomr_b200/obj.py::tangents, itself pinned bit-for-bit to the reference loader by tests/golden/obj_loader.npz)."""
    from axiomr_b200 import api, obj
    dev = api.Device(64, 64)
    try:
        z = np.load(_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden", "obj_loader.npz"))
        meshes = {name: (z[name + "_vertices"][:, :8].copy(), z[name + "_faces"].copy()) for name in ("head", "quad", "poly")}
        v, f = S.icosphere(4)
        meshes["icosphere4"] = (v[:, :8].copy(), f)
        v, f = S.head_like(20, 40)            # poles with valence 40
        meshes["poles"] = (v[:, :8].copy(), f)
        v, f = S.random_triangles(500, 77)
        v8 = v[:, :8].copy()
        v8[::7, 3:5] = 0.25                   # degenerate uv on some corners -> |den| < 1e-8 fallbacks
        v8 = np.concatenate([v8, v8[:3]])     # three vertices no face references
        meshes["random_degenerate_uv"] = (v8, f)
        for name, (v8, f) in meshes.items():
            want = obj.tangents(v8.astype(np.float32), f)
            got = dev.generate_tangents(v8, f)
            assert got.shape == want.shape
            assert np.array_equal(np.isnan(got), np.isnan(want)), name
            ok = (got.view(np.uint32) == want.view(np.uint32)) | np.isnan(want)
            assert ok.all(), (name, np.argwhere(~ok)[:5].tolist())
        # the arrays the reference's own loader produced (golden) are reproduced from its de-duplicated vertices + faces
        for name in ("head", "quad", "poly"):
            want = z[name + "_vertices"]
            got = dev.generate_tangents(want[:, :8], z[name + "_faces"])
            ok = (got.view(np.uint32) == want.view(np.uint32)) | np.isnan(want)
            assert ok.all(), name
    finally:
        dev.close()


def test_c_abi_error_codes_and_lifecycle(po):
    """Error conventions of the boundary: negative codes + axr_last_error text, no crashes, handles reusable after free."""
    from axiomr_b200 import api
    with pytest.raises(api.AxrError) as e:
        api.Device(64, 64, band=(8, 40))          # band start must be a multiple of the reference tile (16)
    assert e.value.code == -1
    with pytest.raises(api.AxrError):
        api.Device(0, 64)
    dev = api.Device(96, 64)
    try:
        v, f = S.quad_grid(2)
        with pytest.raises(api.AxrError) as e:
            dev.upload_mesh(v, f + 100)            # index out of range
        assert e.value.code == -1 and "out of range" in str(e.value)
        with pytest.raises(api.AxrError) as e:
            dev.upload_mesh(v, f, groups=[(0, 3), (5, 5)])   # groups must tile the face range
        assert e.value.code == -1
        m = dev.upload_mesh(v, f)
        with pytest.raises(api.AxrError) as e:
            dev.draw_mesh(m + 7, np.eye(4))        # bad handle
        assert e.value.code == -1
        with pytest.raises(api.AxrError) as e:
            dev.set_shader(9, (0, -1, 0))          # IShader subclass without a device functor
        assert e.value.code == -6
        dev.set_shader(api.SHADER_PBR, (0, -1, 0), (1, 1, 1))
        with pytest.raises(api.AxrError) as e:
            dev.draw_mesh(m, np.eye(4))            # PBR needs five maps: error code instead of the reference's null deref
        assert e.value.code == -5
        t = dev.upload_texture(S.diffuse_texture(8))
        dev.set_material(m, 0, t, t, t, t, t, 0.3)
        dev.set_uniforms(*S.default_camera(96, 64))
        dev.clear()
        dev.draw_mesh(m, np.eye(4))
        c, d = dev.resolve()
        assert np.isfinite(d).any()
        dev.free_texture(t)                        # materials referencing it lose it
        with pytest.raises(api.AxrError) as e:
            dev.draw_mesh(m, np.eye(4))
        assert e.value.code == -5
        dev.free_mesh(m)
        with pytest.raises(api.AxrError):
            dev.free_mesh(m)
        m2 = dev.upload_mesh(v, f)                 # the slot is reused
        assert m2 == m
        dev.set_shader(api.SHADER_FLAT, (0, 0, -1))
        dev.clear()
        dev.draw_mesh(m2, np.eye(4))
        st = dev.stats()
        assert st["faces"] == f.shape[0] and 3 <= st["kernel_launches"] <= 7
        # two contexts on one device do not interfere
        dev2 = api.Device(96, 64)
        try:
            sc = S.Scene("q", 96, 64, v, f, S.SHADER_FLAT)
            mm = dev2.load_scene(sc)
            dev2.clear()
            dev2.draw_mesh(mm, sc.model)
            c2, d2 = dev2.resolve()
            c0, d0, _ = po.oracle_render(sc)
            assert np.array_equal(d2.view(np.uint32), d0.view(np.uint32))
        finally:
            dev2.close()
    finally:
        dev.close()


_CUTOUT_CASES = {
    "small_tris": dict(),
    "clipped_binned": dict(size=8.5, grid=3),
    "huge_9_layers": dict(size=6.0, grid=1, layers=9),
    "dense_640x480": dict(size=5.0, grid=40, tex=256, w=640, h=480),
    "odd_size": dict(w=333, h=217, size=7.0, grid=5, layers=4),
}


@pytest.mark.parametrize("case", list(_CUTOUT_CASES))
def test_discard_shader_depth_peeling(po, case):
    """A shader whose fragment() discards (reference src/tiled_pipeline.cpp:571-577), run through the unmodified reference
    pipeline by the harness's CutoutShader: the owner of a pixel is the nearest NON-discarded fragment. The CUDA path finds it
    by depth peeling; coverage / depth must stay bit-identical, and more than one pass must have run."""
    sc = S.cutout_layers(**_CUTOUT_CASES[case])
    m, st = _check(po, sc, use_ref=True)
    assert m["color_max_diff"] <= 1, m
    assert 0 < m["covered"], m
    assert st["kernel_launches"] > 6, st  # 1 fill + 5 kernels per pass


def test_discard_shader_bilinear_and_random_clip(po):
    """Discard with the bilinear sampler (vs the oracle) and on randomly clipped, interpenetrating triangles (deep peels)."""
    sc = S.cutout_layers(size=5.0, sampler=1)
    m, _ = _check(po, sc)
    assert m["color_max_diff"] <= 1, m
    v, f = S.random_triangles(400, 11)
    sc = S.Scene("random_clip_cutout", 192, 144, v, f, S.SHADER_CUTOUT, textures=[S.cutout_texture(32, 4), None, None, None, None])
    m, st = _check(po, sc, use_ref=True)
    assert m["color_max_diff"] <= 1, m


def test_discard_shader_composites_bands_and_host_framebuffer(po):
    """Peeled draws composite onto earlier draws through the depth test, work on a band context and through drawMesh on a
    host framebuffer (axr_draw_mesh_host: zero-copy stores, depth read from the uploaded copy)."""
    from axiomr_b200 import api
    a = S.config2(level=4, w=320, h=240)                 # opaque sphere first
    b = S.cutout_layers(size=5.0, layers=4)              # then alpha-tested layers, some in front of it, some behind
    c0, d0, _ = po.oracle_render(a)
    c0, d0, _ = po.oracle_render(b, color=c0, depth=d0)
    c1, d1, _ = api.render_scene(a)
    c1, d1, st = api.render_scene(b, color=c1, depth=d1)
    m = po.compare(c1, d1, c0, d0)
    assert m["coverage_mismatch"] == 0 and m["depth_bit_mismatch"] == 0 and m["color_max_diff"] <= 1, m
    # bands
    cb_full, db_full, _ = api.render_scene(b)
    c = np.zeros_like(cb_full)
    d = np.full_like(db_full, np.inf)
    for y0, y1 in ((0, 80), (80, 176), (176, 240)):
        cb, db, _ = api.render_scene(b, band=(y0, y1))
        c[y0:y1], d[y0:y1] = cb[y0:y1], db[y0:y1]
    assert np.array_equal(c, cb_full) and np.array_equal(d.view(np.uint32), db_full.view(np.uint32))
    # reference-shaped host call, drawn twice (idempotent), after an opaque mesh
    fb = api.Framebuffer(b.width, b.height, True)
    fb.clearColor(api.Color(0, 0, 0, 255))
    fb.clearDepth()
    cam = api.Camera()
    cam.setViewport(0, 0, b.width, b.height)
    cam.setViewProjectionMatrix(b.view_proj)
    cam._pos = b.cam_pos
    pipe = api.TiledPipeline(8, cam, fb)
    pipe.setShader(api.FlatShader(tuple(a.light_dir)))
    pipe.drawMesh(a.model, api.Mesh(a.vertices, a.indices, {"m0": api.Material("m0")}))
    pipe.setShader(api.CutoutShader(tuple(b.light_dir)))
    mesh = api.Mesh(b.vertices, b.indices, {"m0": api.Material("m0", api.Texture(b.textures[0]))})
    pipe.drawMesh(b.model, mesh)
    m = po.compare(fb.getColorData(), fb.getDepthData(), c0, d0)
    assert m["coverage_mismatch"] == 0 and m["depth_bit_mismatch"] == 0 and m["color_max_diff"] <= 1, m
    before = fb.getColorData().copy()
    pipe.drawMesh(b.model, mesh)
    assert np.array_equal(before, fb.getColorData())
    # a cutout material without its diffuse texture is an error code
    with pytest.raises(api.AxrError) as e:
        pipe.drawMesh(b.model, api.Mesh(b.vertices, b.indices, {"m0": api.Material("m0")}))
    assert e.value.code == -5


def test_obj_ingestion_behind_the_c_abi_matches_the_reference_loader(po, tmp_path):
    """axr_load_obj / axr_load_obj_file (text parse + value de-duplication in C++, tangents on the device) against the arrays the
    reference's own loader AR::Mesh(path) produced (tests/golden/obj_loader.npz): vertices incl. tangents / bitangents, vertex order,
    faces and groups bit for bit; then the loaded mesh renders like the same arrays uploaded by hand."""
    import os
    from axiomr_b200 import api
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "obj_loader.npz"))
    dev = api.Device(320, 240)
    try:
        for name in ("head", "quad", "poly"):
            text = z[name + "_obj"].tobytes()
            path = tmp_path / f"{name}.obj"
            path.write_bytes(text)
            for src in (text, str(path)):
                mesh, v, f, groups = dev.load_obj(src)
                want = z[name + "_vertices"]
                assert np.array_equal(f, z[name + "_faces"]), name
                assert v.shape == want.shape, name
                ok = (v.view(np.uint32) == want.view(np.uint32)) | np.isnan(want)
                assert ok.all() and np.array_equal(np.isnan(v), np.isnan(want)), (name, np.argwhere(~ok)[:5])
                assert len(groups) == 1 and groups[0].materialName == "m0" and groups[0].startIndex == 0 and groups[0].faceCount == f.shape[0]
                dev.free_mesh(mesh)
        # the loaded mesh draws like the same arrays uploaded through axr_upload_mesh
        mesh, v, f, groups = dev.load_obj(z["head_obj"].tobytes())
        sc = S.Scene("obj_head", 320, 240, v, f, S.SHADER_FLAT)
        dev.set_uniforms(sc.view_proj, sc.cam_pos)
        dev.set_shader(sc.shader, sc.light_dir, sc.light_color)
        dev.clear()
        dev.draw_mesh(mesh, sc.model)
        c1, d1 = dev.resolve()
        c0, d0, _ = po.oracle_render(sc, threads=4)
        m = po.compare(c1, d1, c0, d0)
        assert m["covered"] > 1000 and m["coverage_mismatch"] == 0 and m["depth_bit_mismatch"] == 0 and m["color_max_diff"] <= 1, m
    finally:
        dev.close()


def test_obj_faces_outside_every_material_group_are_not_drawn(po):
    """reference src/tiled_pipeline.cpp:176-179 walks the material groups, and parseModelFile (src/mesh.cpp:336-346) opens a group
    at each `usemtl`: faces in front of the first one belong to none and are never drawn; an OBJ without `usemtl` draws nothing.
    Also: value-equal vertices from different `v` lines merge, corners with a bad position index are dropped, -0 == +0."""
    from axiomr_b200 import api
    obj = b"""# two triangles before any usemtl, then a quad in group a and a triangle in group b
v -1 -1 0
v 1 -1 0
v 0 1 0
v -1 -1 0
vt 0 0
vt 1 0
vn 0 0 1
f 1/1/1 2/2/1 3/1/1
f 4/1/1 2/2/1 3/1/1
usemtl a
v -0.5 -0.5 -0.0
v 0.5 -0.5 0.0
v 0.5 0.5 0
v -0.5 0.5 0
f 5//1 6//1 7//1 8//1 99//1 0//1
usemtl b
f 6/1/1 7/2/1 3/1/1
"""
    dev = api.Device(200, 150)
    try:
        mesh, v, f, groups = dev.load_obj(obj)
        assert f.shape[0] == 5 and v.shape[0] == 8, (f.shape, v.shape)          # 3 + 4 unique corners + (7, vt 2); `v 4` merges into `v 1`, 6/1/1 into 6//1 (vt 1 is 0 0)
        assert np.array_equal(f[0], f[1])                                        # value-equal vertices from different lines
        assert [(g.materialName, g.startIndex, g.faceCount) for g in groups] == [("a", 2, 2), ("b", 4, 1)]
        sc = S.Scene("obj_groups", 200, 150, v, f[2:], S.SHADER_FLAT)            # what the reference draws: the grouped faces
        dev.set_uniforms(sc.view_proj, sc.cam_pos)
        dev.set_shader(sc.shader, sc.light_dir, sc.light_color)
        dev.clear()
        dev.draw_mesh(mesh, sc.model)
        c1, d1 = dev.resolve()
        assert dev.stats()["faces"] == 5
        c0, d0, _ = po.oracle_render(sc, threads=2)
        m = po.compare(c1, d1, c0, d0)
        assert m["covered"] > 100 and m["coverage_mismatch"] == 0 and m["depth_bit_mismatch"] == 0 and m["color_max_diff"] <= 1, m
        # no usemtl at all: nothing is drawn
        mesh2, v2, f2, groups2 = dev.load_obj(b"v -1 -1 0\nv 1 -1 0\nv 0 1 0\nf 1 2 3\n")
        assert f2.shape[0] == 1 and groups2 == []
        dev.clear()
        dev.draw_mesh(mesh2, sc.model)
        c2, d2 = dev.resolve()
        assert np.isinf(d2).all()
        # std::stoi would throw in the reference: an error here, and the context stays usable
        with pytest.raises(api.AxrError):
            dev.load_obj(b"v 0 0 0\nf a b c\n")
        dev.clear()
    finally:
        dev.close()


def test_shader_plugins_loaded_at_run_time(po):
    """A user-supplied IShader (reference include/IShader.hpp:30-46): functors written outside the library against
    include/axr_shader_plugin.cuh, compiled with nvcc (tools/build_shader_plugin.py) and opened with axr_load_shader_plugin.
    alpha_cut discards (depth-peeled draw) with the arithmetic of CutoutShader — which the unmodified reference pipeline runs as an
    IShader subclass in oracle/ref_harness.cpp — so its frames must equal the oracle's; lambert_tint is a shader the library does
    not ship, pinned through its tint = 1 special case and checked to react to its Uniforms::user parameters."""
    import os
    import sys
    if os.environ.get("AXR_SIMT_TESTS_ONLY") == "1":
        pytest.skip("plug-ins are nvcc-built device code: nothing for the interpreter to run")
    from axiomr_b200 import api
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tools"))
    try:
        from build_shader_plugin import build_plugin
    finally:
        sys.path.pop(0)
    cut = build_plugin(os.path.join(root, "tests", "plugins", "alpha_cut.cu"))
    tint = build_plugin(os.path.join(root, "tests", "plugins", "lambert_tint.cu"))
    sc = S.cutout_layers()
    c0, d0, _ = po.oracle_render(sc, threads=4)
    dev = api.Device(sc.width, sc.height, sampler=sc.sampler)
    try:
        mesh = dev.load_scene(sc)
        k_cut = dev.load_shader_plugin(cut)
        assert k_cut >= 64 and dev.load_shader_plugin(cut) == k_cut            # loading twice gives the same kind
        dev.set_shader(k_cut, sc.light_dir, sc.light_color)
        dev.set_shader_user([0.5])
        dev.clear()
        dev.draw_mesh(mesh, sc.model)
        c1, d1 = dev.resolve()
        m = po.compare(c1, d1, c0, d0)
        assert m["covered"] > 1000 and m["coverage_mismatch"] == 0 and m["depth_bit_mismatch"] == 0 and m["color_max_diff"] <= 1, m
        assert dev.stats()["kernel_launches"] > 12                             # several peeling passes
        # threshold 0: nothing is discarded any more, the front layer owns every pixel it covers
        dev.set_shader_user([0.0])
        dev.clear()
        dev.draw_mesh(mesh, sc.model)
        c2, d2 = dev.resolve()
        assert (d2 <= d1).all() and (d2 < d1).any()
        # a shader of our own making: equals the built-in cutout shader where nothing is discarded and the tint is 1
        k_tint = dev.load_shader_plugin(tint)
        assert k_tint == k_cut + 1
        dev.set_shader(k_tint, sc.light_dir, sc.light_color)
        dev.set_shader_user([1.0, 1.0, 1.0])
        dev.clear()
        dev.draw_mesh(mesh, sc.model)
        c3, d3 = dev.resolve()
        assert np.array_equal(d3.view(np.uint32), d2.view(np.uint32)) and np.array_equal(c3, c2)
        dev.set_shader_user([0.5, 1.0, 0.25])
        dev.clear()
        dev.draw_mesh(mesh, sc.model)
        c4, d4 = dev.resolve()
        vis = np.isfinite(d4)
        assert np.array_equal(d4.view(np.uint32), d3.view(np.uint32))
        # colour bytes are B,G,R,A: red halves, green stays, blue quarters (truncation: within one step)
        assert np.abs(c4[vis][:, 2].astype(int) - c3[vis][:, 2].astype(int) // 2).max() <= 1
        assert np.array_equal(c4[vis][:, 1], c3[vis][:, 1])
        assert np.abs(c4[vis][:, 0].astype(int) - c3[vis][:, 0].astype(int) // 4).max() <= 1
        # not a plug-in / unknown kind: refused, context stays usable
        with pytest.raises(api.AxrError):
            dev.load_shader_plugin(os.path.join(root, "oracle", "libaxr_oracle.so"))
        with pytest.raises(api.AxrError):
            dev.set_shader(k_tint + 5, sc.light_dir, sc.light_color)
        dev.set_shader(S.SHADER_FLAT, sc.light_dir, sc.light_color)
        dev.clear()
        dev.draw_mesh(mesh, sc.model)
        dev.sync()
    finally:
        dev.close()


def test_output_fill_and_stale_tile_clears_keep_a_target_equal_to_clear_plus_draw(po):
    """The composite-slot protocol of axiomr_b200/multi.py on one GPU: a target that is cleared ONCE, then receives one draw per frame
    with axr_set_output_fill (every pixel of a touched tile is overwritten: shaded colour or the clear values; no depth read) while its
    owner only clears the tiles the previous frame touched and this one did not (axr_clear_stale_tiles, two alternating dirty maps).
    After every frame the target must equal `clear + that frame's draw` — the camera moves between frames, clipped, huge and small
    triangles mix."""
    from axiomr_b200 import api
    v, f = S.random_triangles(1200, 21)
    v2, f2 = S.icosphere(4, 1.2)
    W, H = 416, 288
    base = S.Scene("fill", W, H, np.concatenate([v, v2]), np.concatenate([f, f2 + v.shape[0]]), S.SHADER_PHONG, textures=_tex())
    eyes = [(0.0, 0.0, 5.0), (2.5, 0.5, 4.0), (-3.0, -1.0, 6.0), (0.0, 0.0, 5.0), (0.3, 2.0, 9.0)]
    dev = api.Device(W, H)
    try:
        mesh = dev.load_scene(base)
        npx, nt = W * H, dev.dirty_map_entries()
        ptr, _ = dev.alloc_shared(npx * 8 + 2 * nt * 4)        # colour | depth | dirty map 0 | dirty map 1 (zero-filled)
        color_p, depth_p, maps = ptr, ptr + npx * 4, (ptr + npx * 8, ptr + npx * 8 + nt * 4)
        dev.set_output(color_p, depth_p)
        dev.clear()                                              # the one real clear of the target
        dev.set_depth_read(False)
        dev.set_output_fill(True)
        for use, eye in enumerate(eyes):
            sc = S.Scene(f"fill{use}", W, H, base.vertices, base.indices, S.SHADER_PHONG, textures=base.textures)
            sc.view_proj, sc.cam_pos = S.default_camera(W, H, eye=eye)
            dev.set_uniforms(sc.view_proj, sc.cam_pos)
            dev.set_dirty_map(maps[use % 2])
            dev.draw_mesh(mesh, sc.model)
            dev.clear_stale_tiles(color_p, depth_p, maps[1 - use % 2], maps[use % 2])
            c1, d1 = dev.resolve()
            c0, d0, _ = po.oracle_render(sc, threads=4)
            m = po.compare(c1, d1, c0, d0)
            assert m["covered"] > 500 and m["coverage_mismatch"] == 0 and m["depth_bit_mismatch"] == 0 and m["color_max_diff"] <= 1, (use, m)
        dev.set_output_fill(False)
        dev.set_dirty_map(None)
        dev.set_output(None, None)
        dev.set_depth_read(True)
    finally:
        dev.close()


def test_tangents_of_a_pole_vertex_with_thousands_of_faces():
    """axr_generate_tangents puts each vertex's incident corners into face order before summing (the reference accumulates in face
    order, src/mesh.cpp:222-298); a pole / fan vertex must not make that quadratic. 3000 faces around one vertex, shuffled:
    bit-equal with the numpy mirror of the reference loader's arithmetic."""
    from axiomr_b200 import api, obj
    n = 3000
    ang = np.linspace(0, 2 * np.pi, n + 1)[:-1]
    v8 = np.zeros((n + 1, 8), np.float32)
    v8[1:, 0] = np.cos(ang); v8[1:, 1] = np.sin(ang); v8[1:, 2] = 0.1 * np.sin(3 * ang)
    v8[:, 3] = v8[:, 0] * 0.5 + 0.5 + (0.01 * np.arange(n + 1)) % 0.1
    v8[:, 4] = v8[:, 1] * 0.5 + 0.5
    v8[:, 5:8] = [0, 0, 1]
    f = np.array([[0, 1 + i, 1 + (i + 1) % n] for i in range(n)], np.uint32)
    f = f[np.random.default_rng(1).permutation(n)]
    dev = api.Device(16, 16)
    try:
        got = dev.generate_tangents(v8, f)
    finally:
        dev.close()
    want = obj.tangents(v8, f)
    same = (got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))
    assert same.all(), np.argwhere(~same)[:5]


def test_host_framebuffer_registrations_can_be_released_and_retaken(po):
    """axr_draw_mesh_host page-locks a caller's framebuffer arrays on first use; axr_host_release undoes that (before the caller
    frees them), a later draw onto the same arrays takes the lock again, and releasing memory that was never locked is a no-op."""
    from axiomr_b200 import api
    sc = S.config2(level=4, w=320, h=200)
    c0, d0, _ = po.oracle_render(sc, threads=4)
    dev = api.Device(sc.width, sc.height)
    try:
        mesh = dev.load_scene(sc)
        for _ in range(2):
            color = np.zeros((sc.height, sc.width, 4), np.uint8)
            color[..., 3] = 255
            depth = np.full((sc.height, sc.width), np.inf, np.float32)
            for _ in range(2):   # the second draw composites onto the first: same picture
                dev.draw_mesh_host(mesh, sc.model, color, depth)
            m = po.compare(color, depth, c0, d0)
            assert m["covered"] > 500 and m["coverage_mismatch"] == 0 and m["depth_bit_mismatch"] == 0 and m["color_max_diff"] <= 1, m
            dev.host_release(color)
            dev.host_release(depth)
            dev.host_release(depth)                      # already released
        dev.host_release(np.zeros(16, np.float32))       # never registered
    finally:
        dev.close()

"""Property tests (hypothesis): random triangles in pixel space, including values that sit exactly on pixel centres, tile borders
and the frame border. CPU: oracle == unmodified reference (where it is present). GPU: CUDA == oracle, bit for bit."""
import os
import sys

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from torture import _ortho_scene  # noqa: E402

W, H = 53, 37  # partial tiles on both axes
_coord = st.one_of(
    st.floats(min_value=-8.0, max_value=60.0, allow_nan=False, width=32),
    st.sampled_from([0.0, 0.5, 15.5, 16.0, 16.5, 31.5, 32.0, 36.5, 37.0, 52.5, 53.0, -0.0, 1e-7, 7.9999995]),
    st.integers(min_value=-2, max_value=56).map(lambda i: i + 0.5),
)
_z = st.one_of(st.floats(min_value=-1.0, max_value=1.0, allow_nan=False, width=32), st.sampled_from([0.0, -0.0, 0.25, 0.25, 1.0, -1.0]))
_tri = st.lists(st.tuples(_coord, _coord, _z), min_size=3, max_size=3)
_scene = st.lists(_tri, min_size=1, max_size=12)


def _build(tris, shader=0):
    return _ortho_scene("prop", np.asarray(tris, dtype=np.float64), W, H, shader)


@pytest.mark.skipif(not os.path.exists("/root/reference/src/tiled_pipeline.cpp"), reason="reference tree not present")
@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(tris=_scene)
def test_oracle_equals_reference_on_random_pixel_space_triangles(po, tris):
    sc = _build(tris)
    c0, d0, _ = po.ref_render(sc, threads=2, chunk=5)
    c1, d1, _ = po.oracle_render(sc, threads=1)
    assert np.array_equal(np.isfinite(d0), np.isfinite(d1))
    assert np.array_equal(d0.view(np.uint32), d1.view(np.uint32))
    assert np.array_equal(c0, c1)


@pytest.mark.gpu
@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(tris=_scene)
def test_cuda_equals_oracle_on_random_pixel_space_triangles(po, tris):
    from axiomr_b200 import api
    sc = _build(tris)
    c1, d1, _ = api.render_scene(sc)
    c0, d0, _ = po.oracle_render(sc, threads=1)
    assert np.array_equal(np.isfinite(d0), np.isfinite(d1))
    assert np.array_equal(d0.view(np.uint32), d1.view(np.uint32))
    assert int(np.abs(c0.astype(np.int16) - c1.astype(np.int16)).max()) <= 1


# ---- the discard branch (reference src/tiled_pipeline.cpp:571-577) with the harness's CutoutShader: overlapping triangles, ties
#      in z, fragments discarded in front of kept ones and the other way round
_CUT_TEX = None


def _cutout(tris):
    global _CUT_TEX
    from axiomr_b200 import scenes as S
    if _CUT_TEX is None:
        _CUT_TEX = S.cutout_texture(16, 2)
    return _ortho_scene("prop_cutout", np.asarray(tris, dtype=np.float64), W, H, S.SHADER_CUTOUT, textures=[_CUT_TEX, None, None, None, None])


@pytest.mark.skipif(not os.path.exists("/root/reference/src/tiled_pipeline.cpp"), reason="reference tree not present")
@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(tris=_scene)
def test_oracle_equals_reference_with_discarding_shader(po, tris):
    sc = _cutout(tris)
    c0, d0, _ = po.ref_render(sc, threads=2, chunk=5)
    c1, d1, _ = po.oracle_render(sc, threads=1)
    assert np.array_equal(d0.view(np.uint32), d1.view(np.uint32))
    assert np.array_equal(c0, c1)


@pytest.mark.gpu
@settings(max_examples=30, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(tris=_scene)
def test_cuda_equals_oracle_with_discarding_shader(po, tris):
    from axiomr_b200 import api
    sc = _cutout(tris)
    c1, d1, _ = api.render_scene(sc)
    c0, d0, _ = po.oracle_render(sc, threads=1)
    assert np.array_equal(d0.view(np.uint32), d1.view(np.uint32))
    assert int(np.abs(c0.astype(np.int16) - c1.astype(np.int16)).max()) <= 1

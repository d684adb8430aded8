"""TEST INFRASTRUCTURE — ctypes bindings for the two CPU checkers. Not part of the product path.

* `ref_*`    : oracle/_ref/libaxr_ref.so — the UNMODIFIED reference sources behind shims (oracle/ref_harness.cpp).
               Buildable only where /root/reference exists; the prebuilt .so travels to the GPU box.
* `oracle_*` : oracle/libaxr_oracle.so — the plain-C restatement (oracle/axr_oracle.c), buildable anywhere with gcc.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libaxr_ref.so")
ORACLE_SO = os.path.join(HERE, "libaxr_oracle.so")

_f32p = C.POINTER(C.c_float)
_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)


def build(ref: bool = True) -> None:
    """Compile the checkers (never the product). `make ref` is a no-op without /root/reference.
    Serialised with a file lock: several test workers (pytest -n) calling this at once would otherwise run `make` on the same
    targets concurrently and link against each other's half-written objects."""
    import fcntl
    with open(os.path.join(HERE, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            subprocess.run(["make", "-s", "-C", HERE, "oracle"] + (["ref"] if ref else []), check=True)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


class _Scene(C.Structure):
    # mirrors `struct axr_ref_scene` (oracle/ref_harness.cpp) == `struct axo_scene` (oracle/axr_oracle.h)
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("threads", C.c_int), ("shader_kind", C.c_int),
                ("light_dir", C.c_float * 3), ("light_color", C.c_float * 3), ("specular_exponent", C.c_float),
                ("tex", _u8p * 5), ("tex_w", C.c_int * 5), ("tex_h", C.c_int * 5),
                ("view_proj", C.c_float * 16), ("cam_pos", C.c_float * 3), ("model", C.c_float * 16),
                ("chunk_faces", C.c_int)]


def _ptr(a, t):
    return a.ctypes.data_as(t)


def _fill_scene(scene, threads: int, chunk: int):
    s = _Scene()
    s.width, s.height, s.threads, s.shader_kind = scene.width, scene.height, threads, scene.shader
    s.light_dir[:] = [float(x) for x in scene.light_dir]
    s.light_color[:] = [float(x) for x in scene.light_color]
    s.specular_exponent = float(scene.specular_exponent)
    keep = []
    for i in range(5):
        t = scene.textures[i]
        if t is None:
            s.tex[i] = None
            s.tex_w[i] = s.tex_h[i] = 0
        else:
            t = np.ascontiguousarray(t, dtype=np.uint8)
            keep.append(t)
            s.tex[i] = _ptr(t, _u8p)
            s.tex_h[i], s.tex_w[i] = t.shape[0], t.shape[1]
    s.view_proj[:] = [float(x) for x in np.asarray(scene.view_proj, dtype=np.float32).reshape(-1)]
    s.cam_pos[:] = [float(x) for x in scene.cam_pos]
    s.model[:] = [float(x) for x in np.asarray(scene.model, dtype=np.float32).reshape(-1)]
    s.chunk_faces = chunk
    return s, keep


def cleared(scene, packed_argb: int = 0xFF000000, depth: float = float("inf")):
    """Framebuffer after clearColor/clearDepth (reference src/framebuffer.cpp:26-42): BGRA8 bytes + f32 depth."""
    color = np.empty((scene.height, scene.width), dtype=np.uint32)
    color[:] = packed_argb
    z = np.full((scene.height, scene.width), depth, dtype=np.float32)
    return color.view(np.uint8).reshape(scene.height, scene.width, 4), z


# ----------------------------------------------------------------------------- unmodified reference (oracle/_ref)
_ref = None


def ref_available() -> bool:
    return os.path.exists(REF_SO)


def ref_lib():
    global _ref
    if _ref is None:
        lib = C.CDLL(REF_SO)
        lib.axr_ref_render.restype = C.c_int
        lib.axr_ref_render.argtypes = [C.POINTER(_Scene), _f32p, C.c_uint64, _u32p, C.c_uint64, _u8p, _f32p,
                                       C.POINTER(C.c_double)]
        lib.axr_ref_camera.argtypes = [_f32p, _f32p, C.c_float, C.c_float, C.c_int, C.c_int, _f32p, _f32p]
        lib.axr_ref_mat4_mul.argtypes = [_f32p, _f32p, _f32p]
        lib.axr_ref_clip_triangle.argtypes = [_f32p, _f32p]
        lib.axr_ref_triangle_setup.argtypes = [_f32p, C.c_int, C.c_int, _f32p]
        lib.axr_ref_texture_sample.argtypes = [_u8p, C.c_int, C.c_int, _f32p, C.c_int, _f32p]
        lib.axr_ref_mesh_load.restype = C.c_void_p
        lib.axr_ref_mesh_load.argtypes = [C.c_char_p]
        lib.axr_ref_mesh_counts.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        lib.axr_ref_mesh_copy.argtypes = [C.c_void_p, _f32p, _u32p]
        lib.axr_ref_mesh_free.argtypes = [C.c_void_p]
        _ref = lib
    return _ref


def _render(fn, scene, threads, chunk, color, depth, first_face=0, n_faces=None):
    s, keep = _fill_scene(scene, threads, chunk)
    if color is None:
        color, depth = cleared(scene)
    color = np.ascontiguousarray(color, dtype=np.uint8).copy()
    depth = np.ascontiguousarray(depth, dtype=np.float32).copy()
    idx = scene.indices[first_face:(None if n_faces is None else first_face + n_faces)]
    idx = np.ascontiguousarray(idx, dtype=np.uint32)
    secs = C.c_double(0.0)
    rc = fn(C.byref(s), _ptr(scene.vertices, _f32p), scene.n_verts, _ptr(idx, _u32p), idx.shape[0],
            _ptr(color, _u8p), _ptr(depth, _f32p), C.byref(secs))
    if rc != 0:
        raise RuntimeError(f"CPU checker render failed rc={rc}")
    return color, depth, secs.value


def ref_render(scene, threads: int = 1, chunk: int = 30000, color=None, depth=None, first_face=0, n_faces=None):
    """Reference TiledPipeline::drawMesh in chunks of `chunk` faces. Returns (BGRA8 HxWx4, depth HxW, seconds in drawMesh)."""
    return _render(ref_lib().axr_ref_render, scene, threads, chunk, color, depth, first_face, n_faces)


def ref_hardware_concurrency() -> int:
    return int(ref_lib().axr_ref_hardware_concurrency())


def ref_camera(pos, target, fov, w, h):
    vp = np.zeros(16, dtype=np.float32)
    vpt = np.zeros(16, dtype=np.float32)
    p = np.asarray(pos, dtype=np.float32)
    t = np.asarray(target, dtype=np.float32)
    ref_lib().axr_ref_camera(_ptr(p, _f32p), _ptr(t, _f32p), fov, w / h, w, h, _ptr(vp, _f32p), _ptr(vpt, _f32p))
    return vp.reshape(4, 4), vpt.reshape(4, 4)


def ref_clip_triangle(tri3x18: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(tri3x18, dtype=np.float32).reshape(3, 18)
    out = np.zeros((24, 18), dtype=np.float32)
    n = ref_lib().axr_ref_clip_triangle(_ptr(a, _f32p), _ptr(out, _f32p))
    return out[:n].copy()


def ref_triangle_setup(tri3x18: np.ndarray, w: int, h: int):
    a = np.ascontiguousarray(tri3x18, dtype=np.float32).reshape(3, 18)
    out = np.zeros(13, dtype=np.float32)
    back = ref_lib().axr_ref_triangle_setup(_ptr(a, _f32p), w, h, _ptr(out, _f32p))
    return bool(back), out


def ref_texture_sample(tex: np.ndarray, uv: np.ndarray) -> np.ndarray:
    tex = np.ascontiguousarray(tex, dtype=np.uint8)
    uv = np.ascontiguousarray(uv, dtype=np.float32).reshape(-1, 2)
    out = np.zeros((uv.shape[0], 4), dtype=np.float32)
    rc = ref_lib().axr_ref_texture_sample(_ptr(tex, _u8p), tex.shape[1], tex.shape[0], _ptr(uv, _f32p), uv.shape[0],
                                          _ptr(out, _f32p))
    if rc != 0:
        raise RuntimeError("axr_ref_texture_sample failed")
    return out


def ref_load_obj(path: str):
    lib = ref_lib()
    h = lib.axr_ref_mesh_load(path.encode())
    if not h:
        raise RuntimeError(f"reference Mesh loader failed on {path}")
    nv, nf = C.c_uint64(0), C.c_uint64(0)
    lib.axr_ref_mesh_counts(h, C.byref(nv), C.byref(nf))
    v = np.zeros((nv.value, 14), dtype=np.float32)
    f = np.zeros((nf.value, 3), dtype=np.uint32)
    lib.axr_ref_mesh_copy(h, _ptr(v, _f32p), _ptr(f, _u32p))
    lib.axr_ref_mesh_free(h)
    return v, f


# ----------------------------------------------------------------------------- C restatement (oracle/axr_oracle.c)
_orc = None


def oracle_available() -> bool:
    return os.path.exists(ORACLE_SO)


def oracle_lib():
    global _orc
    if _orc is None:
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        lib = C.CDLL(ORACLE_SO)
        lib.axo_render.restype = C.c_int
        lib.axo_render.argtypes = [C.POINTER(_Scene), _f32p, C.c_uint64, _u32p, C.c_uint64, _u8p, _f32p,
                                   C.POINTER(C.c_double)]
        lib.axo_clip_triangle.argtypes = [_f32p, _f32p]
        lib.axo_triangle_setup.argtypes = [_f32p, C.c_int, C.c_int, _f32p]
        lib.axo_texture_sample.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _f32p, C.c_int, _f32p]
        lib.axo_mat4_mul.argtypes = [_f32p, _f32p, _f32p]
        _orc = lib
    return _orc


def oracle_render(scene, threads: int = 1, color=None, depth=None, first_face=0, n_faces=None, sampler=None):
    """C restatement of drawMesh. `threads` parallelises tiles only (results are thread-count independent).
    chunk_faces carries the sampler mode for this checker: 0 nearest (reference), 1 bilinear (extension)."""
    mode = scene.sampler if sampler is None else sampler
    return _render(oracle_lib().axo_render, scene, threads, mode, color, depth, first_face, n_faces)


def oracle_clip_triangle(tri3x18: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(tri3x18, dtype=np.float32).reshape(3, 18)
    out = np.zeros((24, 18), dtype=np.float32)
    n = oracle_lib().axo_clip_triangle(_ptr(a, _f32p), _ptr(out, _f32p))
    return out[:n].copy()


def oracle_triangle_setup(tri3x18: np.ndarray, w: int, h: int):
    a = np.ascontiguousarray(tri3x18, dtype=np.float32).reshape(3, 18)
    out = np.zeros(13, dtype=np.float32)
    back = oracle_lib().axo_triangle_setup(_ptr(a, _f32p), w, h, _ptr(out, _f32p))
    return bool(back), out


def oracle_texture_sample(tex: np.ndarray, uv: np.ndarray, sampler: int = 0) -> np.ndarray:
    tex = np.ascontiguousarray(tex, dtype=np.uint8)
    uv = np.ascontiguousarray(uv, dtype=np.float32).reshape(-1, 2)
    out = np.zeros((uv.shape[0], 4), dtype=np.float32)
    oracle_lib().axo_texture_sample(_ptr(tex, _u8p), tex.shape[1], tex.shape[0], sampler, _ptr(uv, _f32p), uv.shape[0],
                                    _ptr(out, _f32p))
    return out


# ----------------------------------------------------------------------------- comparison (BASELINE.json tolerances)
def compare(color_a, depth_a, color_b, depth_b, depth_rtol: float = 1e-6):
    """Parity metrics between two framebuffers (a = candidate, b = checker)."""
    da, db = np.asarray(depth_a), np.asarray(depth_b)
    ca = np.asarray(color_a).astype(np.int16)
    cb = np.asarray(color_b).astype(np.int16)
    cov_a, cov_b = np.isfinite(da), np.isfinite(db)
    n = da.size
    cov_mismatch = int(np.count_nonzero(cov_a != cov_b))
    both = cov_a & cov_b
    rel = np.zeros_like(da, dtype=np.float64)
    rel[both] = np.abs(da[both].astype(np.float64) - db[both]) / np.maximum(np.abs(db[both]).astype(np.float64), 1e-30)
    depth_bad = int(np.count_nonzero(rel > depth_rtol))
    depth_bits = int(np.count_nonzero(da.view(np.uint32) != db.view(np.uint32)))
    cdiff = np.abs(ca - cb).max(axis=-1)
    return {
        "pixels": int(n),
        "covered": int(np.count_nonzero(cov_b)),
        "coverage_mismatch": cov_mismatch,
        "coverage_mismatch_frac": cov_mismatch / n,
        "depth_max_rel": float(rel.max()) if n else 0.0,
        "depth_bad": depth_bad,
        "depth_bit_mismatch": depth_bits,
        "color_exact_frac": float(np.count_nonzero(cdiff == 0)) / n,
        "color_within1_frac": float(np.count_nonzero(cdiff <= 1)) / n,
        "color_max_diff": int(cdiff.max()) if n else 0,
    }


def assert_parity(m: dict, cov_frac=1e-4, depth_rtol_ok=True, color_frac=0.999):
    """BASELINE.json north_star: coverage identical except <=0.01% of pixels, depth within 1e-6 relative,
    8-bit colour within 1 LSB on >=99.9% of pixels."""
    assert m["coverage_mismatch_frac"] <= cov_frac, m
    if depth_rtol_ok:
        assert m["depth_bad"] <= m["coverage_mismatch"] + int(cov_frac * m["pixels"]), m
    assert m["color_within1_frac"] >= color_frac, m

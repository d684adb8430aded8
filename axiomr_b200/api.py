"""Host-side mirror of AxiomR's pipeline interface on top of the C ABI (include/axr_b200.h).

Names, argument meaning and error behaviour follow the reference's classes for this path so the parity tests read
like tests of the reference:

    reference (C++)                                      here (Python over ctypes -> libaxr_b200.so)
    AR::Framebuffer(w, h, useDepth)   framebuffer.hpp     Framebuffer(w, h)   clearColor / clearDepth / getColorData / getDepthData
    AR::Camera                        camera.hpp          Camera              getViewProjectionMatrix / getViewportMatrix / getPosition
    AR::Texture(path) -> RGBA8        texture.hpp         Texture(rgba)
    AR::Mesh + Material(Group)        mesh.hpp            Mesh(vertices, indices, groups), Material
    AR::FlatShader/PhongShader/PBRShader  shaders.hpp     same names, public fields lightDirection / lightColor
    AR::TiledPipeline(threads, cam, fb)   tiled_pipeline.hpp   TiledPipeline(threads, camera, framebuffer): setShader, setCamera,
                                                               setFramebuffer, drawMesh(model, mesh), getViewportMat

`drawMesh` has the reference's semantics: it composites onto the *host* framebuffer's current contents with a strict
depth test and is complete on return. The device-resident fast path the bench's `value` uses (framebuffer kept in
HBM between draws) is `TiledPipeline.device` (a `Device` object wrapping one axr_ctx).

There is no CPU fallback: importing works anywhere, but creating a Device without libaxr_b200.so or without a B200
raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

from . import scenes as _scenes

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AXR_B200_LIB") or os.path.join(HERE, "libaxr_b200.so")  # AXR_B200_LIB: tuning variants

SHADER_FLAT, SHADER_PHONG, SHADER_PBR, SHADER_CUTOUT = 0, 1, 2, 3
SAMPLER_NEAREST, SAMPLER_BILINEAR = 0, 1
COLOR_EXACT, COLOR_FAST = 0, 1
NO_TEXTURE = -1

_f32p = C.POINTER(C.c_float)
_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)


class AxrError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"axr error {code}: {msg}")
        self.code = code


class _Config(C.Structure):
    _fields_ = [("device", C.c_int), ("width", C.c_int), ("height", C.c_int), ("sampler", C.c_int),
                ("band_y0", C.c_int), ("band_y1", C.c_int), ("stream", C.c_void_p), ("reserved", C.c_uint64 * 4)]


class _ShaderParams(C.Structure):
    _fields_ = [("light_dir", C.c_float * 3), ("light_color", C.c_float * 3)]


class _Group(C.Structure):
    _fields_ = [("first_face", C.c_uint64), ("face_count", C.c_uint64)]


class ObjInfo(C.Structure):
    _fields_ = [("n_verts", C.c_uint64), ("n_faces", C.c_uint64), ("n_groups", C.c_uint32), ("reserved", C.c_uint32),
                ("first_drawn_face", C.c_uint64)]


class MtlEntry(C.Structure):
    _fields_ = [("name", C.c_char * 128), ("specular_exponent", C.c_float), ("has_map", C.c_int * 5), ("map", (C.c_char * 512) * 5)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("faces", "clipped_faces", "triangles", "small_triangles", "binned_triangles",
                                          "bin_refs", "kernel_launches", "redo")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


# every symbol include/axr_b200.h declares (tests check the built library exports all of them)
ABI_SYMBOLS = [
    "axr_create", "axr_destroy", "axr_last_error", "axr_abi_version", "axr_upload_mesh", "axr_free_mesh",
    "axr_upload_texture", "axr_free_texture", "axr_set_material", "axr_set_uniforms", "axr_set_shader", "axr_set_sampler",
    "axr_clear", "axr_upload_framebuffer", "axr_resolve", "axr_draw_mesh", "axr_sync", "axr_get_stats", "axr_host_alloc",
    "axr_host_free", "axr_stream", "axr_framebuffer_device", "axr_set_output", "axr_framebuffer_ipc", "axr_open_ipc",
    "axr_close_ipc", "axr_set_profiling", "axr_get_kernel_times", "axr_set_depth_read", "axr_alloc_shared", "axr_free_shared", "axr_set_overlap", "axr_upload_framebuffer_async", "axr_draw_mesh_host", "axr_generate_tangents", "axr_measure_fp32_issue", "axr_set_color_math", "axr_update_mesh_vertices", "axr_dirty_map_entries", "axr_set_dirty_map", "axr_clear_dirty_tiles",
    "axr_load_obj", "axr_load_obj_file", "axr_mesh_group_info", "axr_mesh_read", "axr_parse_mtl",
    "axr_load_shader_plugin", "axr_set_shader_user", "axr_set_output_fill", "axr_clear_stale_tiles", "axr_set_output_rows", "axr_host_release", "axr_measure_gather",
]
STAGES = ["vertex_xform", "setup_raster", "scan_tiles", "bin_scatter", "tile_shade"]

_lib = None


def load_library():
    """dlopen libaxr_b200.so (built in-tree by axiomr_b200/build.py). Fails loudly when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -m axiomr_b200.build` (nvcc, sm_100a). "
                          "There is no CPU fallback for the raster path.")
    lib = C.CDLL(LIB_PATH)
    if hasattr(lib, "axr_simt_interpreter_marker"):
        # tests/simt compiles the kernels for a CPU-side SIMT interpreter (kernel unit tests without a GPU). It is not a device and
        # this loader never accepts it; the test harness binds it itself (tests/simt/use_simt.py).
        raise ImportError(f"{LIB_PATH} is the test-only SIMT interpreter build, not the CUDA library. "
                          "There is no CPU fallback for the raster path.")
    _lib = _bind(lib)
    return _lib


def _bind(lib):
    """ctypes signatures of the C ABI (include/axr_b200.h)."""
    vp = C.c_void_p
    lib.axr_create.argtypes = [C.POINTER(_Config), C.POINTER(vp)]
    lib.axr_destroy.argtypes = [vp]
    lib.axr_destroy.restype = None
    lib.axr_last_error.argtypes = [vp]
    lib.axr_last_error.restype = C.c_char_p
    lib.axr_upload_mesh.argtypes = [vp, _f32p, C.c_uint64, _u32p, C.c_uint64, C.POINTER(_Group), C.c_uint32, C.POINTER(C.c_int32)]
    lib.axr_free_mesh.argtypes = [vp, C.c_int32]
    lib.axr_update_mesh_vertices.argtypes = [vp, C.c_int32, _f32p, C.c_uint64]
    lib.axr_generate_tangents.argtypes = [vp, _f32p, C.c_uint64, _u32p, C.c_uint64, _f32p]
    lib.axr_upload_texture.argtypes = [vp, _u8p, C.c_int, C.c_int, C.POINTER(C.c_int32)]
    lib.axr_free_texture.argtypes = [vp, C.c_int32]
    lib.axr_set_material.argtypes = [vp, C.c_int32, C.c_uint32] + [C.c_int32] * 5 + [C.c_float]
    lib.axr_set_uniforms.argtypes = [vp, _f32p, _f32p, _f32p]
    lib.axr_set_shader.argtypes = [vp, C.c_int, C.POINTER(_ShaderParams), C.c_size_t]
    lib.axr_set_sampler.argtypes = [vp, C.c_int]
    lib.axr_set_color_math.argtypes = [vp, C.c_int]
    lib.axr_load_shader_plugin.argtypes = [vp, C.c_char_p, C.POINTER(C.c_int)]
    lib.axr_set_shader_user.argtypes = [vp, _f32p, C.c_uint32]
    lib.axr_load_obj.argtypes = [vp, C.c_char_p, C.c_size_t, C.POINTER(C.c_int32), C.POINTER(ObjInfo)]
    lib.axr_load_obj_file.argtypes = [vp, C.c_char_p, C.POINTER(C.c_int32), C.POINTER(ObjInfo)]
    lib.axr_mesh_group_info.argtypes = [vp, C.c_int32, C.c_uint32, C.c_char_p, C.c_size_t, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.axr_mesh_read.argtypes = [vp, C.c_int32, _f32p, _u32p]
    lib.axr_parse_mtl.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(MtlEntry), C.c_uint32, C.POINTER(C.c_uint32)]
    lib.axr_clear.argtypes = [vp, C.c_uint32, C.c_float]
    lib.axr_upload_framebuffer.argtypes = [vp, C.c_void_p, C.c_void_p]
    lib.axr_upload_framebuffer_async.argtypes = [vp, C.c_void_p, C.c_void_p]
    lib.axr_resolve.argtypes = [vp, C.c_void_p, C.c_void_p]
    lib.axr_draw_mesh.argtypes = [vp, C.c_int32, _f32p]
    lib.axr_draw_mesh_host.argtypes = [vp, C.c_int32, _f32p, C.c_void_p, C.c_void_p]
    lib.axr_sync.argtypes = [vp]
    lib.axr_get_stats.argtypes = [vp, C.POINTER(Stats)]
    lib.axr_host_alloc.argtypes = [C.c_size_t]
    lib.axr_host_alloc.restype = vp
    lib.axr_host_free.argtypes = [vp]
    lib.axr_host_free.restype = None
    lib.axr_stream.argtypes = [vp]
    lib.axr_stream.restype = vp
    lib.axr_framebuffer_device.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    lib.axr_set_output.argtypes = [vp, vp, vp]
    lib.axr_framebuffer_ipc.argtypes = [vp, C.c_void_p, C.c_void_p]
    lib.axr_open_ipc.argtypes = [vp, C.c_void_p, C.POINTER(vp)]
    lib.axr_close_ipc.argtypes = [vp, vp]
    lib.axr_set_depth_read.argtypes = [vp, C.c_int]
    lib.axr_set_overlap.argtypes = [vp, C.c_int]
    lib.axr_dirty_map_entries.argtypes = [vp]
    lib.axr_set_dirty_map.argtypes = [vp, vp]
    lib.axr_clear_dirty_tiles.argtypes = [vp, vp, vp, vp, C.c_int, C.c_uint32, C.c_float, vp]
    lib.axr_set_output_fill.argtypes = [vp, C.c_int, C.c_uint32, C.c_float]
    lib.axr_set_output_rows.argtypes = [vp, C.c_int]
    lib.axr_host_release.argtypes = [vp, vp]
    lib.axr_measure_gather.argtypes = [vp, C.POINTER(C.c_double)]
    lib.axr_clear_stale_tiles.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_uint32, C.c_float, vp]
    lib.axr_alloc_shared.argtypes = [vp, C.c_size_t, C.POINTER(vp), C.c_void_p]
    lib.axr_free_shared.argtypes = [vp, vp]
    lib.axr_measure_fp32_issue.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.axr_set_profiling.argtypes = [vp, C.c_int]
    lib.axr_get_kernel_times.argtypes = [vp, _f32p, C.POINTER(C.c_uint64)]
    return lib


def _mat(a) -> np.ndarray:
    m = np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(16))
    return m


# ------------------------------------------------------------------------------------------- Device = one axr_ctx
class Device:
    """One raster context on one GPU (optionally one screen-space band of the frame)."""

    def __init__(self, width: int, height: int, device: int = 0, sampler: int = SAMPLER_NEAREST, band=None, stream=None):
        self.lib = load_library()
        cfg = _Config()
        cfg.device, cfg.width, cfg.height, cfg.sampler = device, width, height, sampler
        cfg.band_y0, cfg.band_y1 = (band if band else (0, 0))
        cfg.stream = stream
        h = C.c_void_p()
        rc = self.lib.axr_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise AxrError(rc, (self.lib.axr_last_error(None) or b"").decode())
        self.h = h
        self.width, self.height = width, height
        self.band = band if band else (0, height)
        self._meshes = {}
        self._textures = {}

    def close(self):
        if getattr(self, "h", None):
            self.lib.axr_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise AxrError(rc, (self.lib.axr_last_error(self.h) or b"").decode())

    # --- scene data
    def upload_texture(self, rgba: np.ndarray) -> int:
        t = np.ascontiguousarray(rgba, dtype=np.uint8)
        assert t.ndim == 3 and t.shape[2] == 4
        out = C.c_int32(-1)
        self._check(self.lib.axr_upload_texture(self.h, t.ctypes.data_as(_u8p), t.shape[1], t.shape[0], C.byref(out)))
        return out.value

    def free_texture(self, tex: int):
        self._check(self.lib.axr_free_texture(self.h, tex))

    def upload_mesh(self, vertices: np.ndarray, indices: np.ndarray, groups=None) -> int:
        v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 14)
        f = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1, 3)
        out = C.c_int32(-1)
        garr, ng = None, 0
        if groups:
            ng = len(groups)
            garr = (_Group * ng)(*[_Group(int(a), int(b)) for a, b in groups])
        self._check(self.lib.axr_upload_mesh(self.h, v.ctypes.data_as(_f32p), v.shape[0], f.ctypes.data_as(_u32p), f.shape[0],
                                             garr, ng, C.byref(out)))
        return out.value

    def generate_tangents(self, pos_uv_normal: np.ndarray, indices: np.ndarray) -> np.ndarray:
        """Mesh::calculateTangentBitangent on the device: (V,8) f32 + (T,3) u32 -> (V,14) f32 in AR::Vertex layout."""
        v = np.ascontiguousarray(pos_uv_normal, dtype=np.float32).reshape(-1, 8)
        f = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1, 3)
        out = np.zeros((v.shape[0], 14), dtype=np.float32)
        self._check(self.lib.axr_generate_tangents(self.h, v.ctypes.data_as(_f32p), v.shape[0], f.ctypes.data_as(_u32p), f.shape[0],
                                                   out.ctypes.data_as(_f32p)))
        return out

    def load_obj(self, text_or_path):
        """AR::Mesh(path) behind the C ABI (axr_load_obj / axr_load_obj_file): bytes = OBJ text, str = path.
        Returns (mesh handle, vertices (V,14) f32, faces (T,3) u32, [MaterialGroup...]) — the arrays are the reference loader's."""
        out, info = C.c_int32(-1), ObjInfo()
        if isinstance(text_or_path, (bytes, bytearray)):
            self._check(self.lib.axr_load_obj(self.h, bytes(text_or_path), len(text_or_path), C.byref(out), C.byref(info)))
        else:
            self._check(self.lib.axr_load_obj_file(self.h, os.fsencode(text_or_path), C.byref(out), C.byref(info)))
        v = np.zeros((info.n_verts, 14), dtype=np.float32)
        f = np.zeros((info.n_faces, 3), dtype=np.uint32)
        self._check(self.lib.axr_mesh_read(self.h, out.value, v.ctypes.data_as(_f32p), f.ctypes.data_as(_u32p)))
        groups = []
        for g in range(info.n_groups):
            name = C.create_string_buffer(256)
            first, count = C.c_uint64(), C.c_uint64()
            self._check(self.lib.axr_mesh_group_info(self.h, out.value, g, name, 256, C.byref(first), C.byref(count)))
            groups.append(MaterialGroup(name.value.decode("utf-8", "replace"), int(first.value), int(count.value)))
        return out.value, v, f, groups

    def update_mesh_vertices(self, mesh: int, vertices: np.ndarray):
        """Re-send the vertices of an uploaded mesh (same count, same faces)."""
        v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 14)
        self._check(self.lib.axr_update_mesh_vertices(self.h, mesh, v.ctypes.data_as(_f32p), v.shape[0]))

    def free_mesh(self, mesh: int):
        self._check(self.lib.axr_free_mesh(self.h, mesh))

    def set_material(self, mesh: int, group: int, diffuse=NO_TEXTURE, bump=NO_TEXTURE, metallic=NO_TEXTURE,
                     roughness=NO_TEXTURE, ao=NO_TEXTURE, specular_exponent: float = 0.0):
        self._check(self.lib.axr_set_material(self.h, mesh, group, diffuse, bump, metallic, roughness, ao, specular_exponent))

    # --- per-frame state
    def set_uniforms(self, view_proj, cam_pos, viewport=None):
        vp = _mat(view_proj)
        vpt = _mat(viewport) if viewport is not None else _mat(_scenes.viewport_matrix(self.width, self.height))
        cp = np.ascontiguousarray(np.asarray(cam_pos, dtype=np.float32).reshape(3))
        self._check(self.lib.axr_set_uniforms(self.h, vp.ctypes.data_as(_f32p), vpt.ctypes.data_as(_f32p), cp.ctypes.data_as(_f32p)))

    def set_shader(self, kind: int, light_dir, light_color=(1.0, 1.0, 1.0)):
        p = _ShaderParams()
        p.light_dir[:] = [float(x) for x in light_dir]
        p.light_color[:] = [float(x) for x in light_color]
        self._check(self.lib.axr_set_shader(self.h, kind, C.byref(p), C.sizeof(p)))

    def load_shader_plugin(self, path: str) -> int:
        """A user-written shader functor compiled with tools/build_shader_plugin.py: returns the shader kind for set_shader()."""
        kind = C.c_int(-1)
        self._check(self.lib.axr_load_shader_plugin(self.h, os.fsencode(path), C.byref(kind)))
        return kind.value

    def set_shader_user(self, values=()):
        v = np.ascontiguousarray(np.asarray(values, dtype=np.float32).reshape(-1))
        self._check(self.lib.axr_set_shader_user(self.h, v.ctypes.data_as(_f32p) if v.size else None, v.size))

    def set_sampler(self, sampler: int):
        self._check(self.lib.axr_set_sampler(self.h, sampler))

    def set_color_math(self, mode: int):
        """COLOR_FAST (default; colour within 1 LSB of the reference) or COLOR_EXACT (the reference's rounding order)."""
        self._check(self.lib.axr_set_color_math(self.h, mode))

    # --- framebuffer
    def clear(self, packed_argb: int = 0xFF000000, depth: float = float("inf")):
        self._check(self.lib.axr_clear(self.h, packed_argb, depth))

    def upload_framebuffer(self, color: np.ndarray | None, depth: np.ndarray | None, wait: bool = True):
        """wait=False: only enqueued — the arrays must stay untouched until the next resolve()/sync()."""
        cp = color.ctypes.data if color is not None else None
        dp = depth.ctypes.data if depth is not None else None
        fn = self.lib.axr_upload_framebuffer if wait else self.lib.axr_upload_framebuffer_async
        self._check(fn(self.h, cp, dp))

    def resolve(self, color: np.ndarray | None = None, depth: np.ndarray | None = None):
        """Device -> host, synchronous. Allocates the outputs when not given. Returns (BGRA8 HxWx4, depth HxW)."""
        if color is None:
            color = np.zeros((self.height, self.width, 4), dtype=np.uint8)
        if depth is None:
            depth = np.full((self.height, self.width), np.inf, dtype=np.float32)
        self._check(self.lib.axr_resolve(self.h, color.ctypes.data, depth.ctypes.data))
        return color, depth

    def present(self, path: str) -> None:
        """Read the colour plane back (depth is not needed for display) and write it as a PNG; see api.present."""
        color = np.zeros((self.height, self.width, 4), dtype=np.uint8)
        self._check(self.lib.axr_resolve(self.h, color.ctypes.data, None))
        present(color, path)

    # --- hot path
    def draw_mesh(self, mesh: int, model):
        m = _mat(model)
        self._check(self.lib.axr_draw_mesh(self.h, mesh, m.ctypes.data_as(_f32p)))

    def draw_mesh_host(self, mesh: int, model, color: np.ndarray, depth: np.ndarray):
        """drawMesh onto a HOST framebuffer (BGRA8 HxWx4 + f32 HxW, C-contiguous), complete on return."""
        m = _mat(model)
        assert color.flags.c_contiguous and depth.flags.c_contiguous and color.dtype == np.uint8 and depth.dtype == np.float32
        self._check(self.lib.axr_draw_mesh_host(self.h, mesh, m.ctypes.data_as(_f32p), color.ctypes.data, depth.ctypes.data))

    def sync(self):
        self._check(self.lib.axr_sync(self.h))

    def stats(self) -> dict:
        s = Stats()
        self._check(self.lib.axr_get_stats(self.h, C.byref(s)))
        return s.as_dict()

    def measure_fp32_issue(self):
        """(FMUL+FADD warp-instructions/s, FFMA warp-instructions/s) measured on this GPU."""
        a, b = C.c_double(0), C.c_double(0)
        self._check(self.lib.axr_measure_fp32_issue(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def set_profiling(self, enabled: bool):
        self._check(self.lib.axr_set_profiling(self.h, 1 if enabled else 0))

    def kernel_times(self):
        """(dict stage -> accumulated ms, draws) since profiling was enabled / last read."""
        ms = (C.c_float * len(STAGES))()
        n = C.c_uint64(0)
        self._check(self.lib.axr_get_kernel_times(self.h, ms, C.byref(n)))
        return {k: float(ms[i]) for i, k in enumerate(STAGES)}, int(n.value)

    # --- interop
    @property
    def stream(self) -> int:
        return int(self.lib.axr_stream(self.h) or 0)

    def framebuffer_device(self):
        c, d = C.c_void_p(), C.c_void_p()
        self._check(self.lib.axr_framebuffer_device(self.h, C.byref(c), C.byref(d)))
        return int(c.value), int(d.value)

    def set_output(self, color_dev: int | None, depth_dev: int | None):
        self._check(self.lib.axr_set_output(self.h, color_dev, depth_dev))

    def framebuffer_ipc(self):
        a, b = C.create_string_buffer(64), C.create_string_buffer(64)
        self._check(self.lib.axr_framebuffer_ipc(self.h, a, b))
        return a.raw, b.raw

    def open_ipc(self, handle: bytes) -> int:
        p = C.c_void_p()
        self._check(self.lib.axr_open_ipc(self.h, C.create_string_buffer(handle, 64), C.byref(p)))
        return int(p.value)

    def dirty_map_entries(self) -> int:
        return int(self.lib.axr_dirty_map_entries(self.h))

    def set_dirty_map(self, dirty_dev: int | None):
        """Every draw flags the 32x32 tiles it may store into in this u32 map (device pointer, may be peer memory); None = off."""
        self._check(self.lib.axr_set_dirty_map(self.h, dirty_dev))

    def clear_dirty_tiles(self, color_dev: int, depth_dev: int, dirty_dev: int, count: int = 1, packed_argb: int = 0xFF000000,
                          depth: float = float("inf"), stream: int | None = None):
        self._check(self.lib.axr_clear_dirty_tiles(self.h, color_dev, depth_dev, dirty_dev, count, packed_argb, depth, stream))

    def set_output_fill(self, enabled: bool, packed_argb: int = 0xFF000000, depth: float = float("inf")):
        """Draws overwrite every pixel of the tiles they touch (shaded colour or these values): see axr_set_output_fill."""
        self._check(self.lib.axr_set_output_fill(self.h, 1 if enabled else 0, packed_argb, depth))

    def measure_gather(self) -> float:
        """Randomly placed 32-byte DRAM sectors per second (16-byte gathers): the shading stage's other bound."""
        v = C.c_double(0.0)
        self._check(self.lib.axr_measure_gather(self.h, C.byref(v)))
        return v.value

    def host_release(self, array: np.ndarray):
        """Undo the page-lock axr_draw_mesh_host took on a host framebuffer array (before the array is freed)."""
        self._check(self.lib.axr_host_release(self.h, array.ctypes.data))

    def set_output_rows(self, enabled: bool):
        """A warp shades a 32 x 1 pixel row (128-byte stores) instead of an 8 x 4 block: for outputs behind NVLink."""
        self._check(self.lib.axr_set_output_rows(self.h, 1 if enabled else 0))

    def clear_stale_tiles(self, color_dev: int, depth_dev: int, dirty_prev_dev: int, dirty_now_dev: int, count: int = 1,
                          packed_argb: int = 0xFF000000, depth: float = float("inf"), stream: int | None = None):
        self._check(self.lib.axr_clear_stale_tiles(self.h, color_dev, depth_dev, dirty_prev_dev, dirty_now_dev, count, packed_argb, depth, stream))

    def set_overlap(self, enabled: bool):
        self._check(self.lib.axr_set_overlap(self.h, 1 if enabled else 0))

    def set_depth_read(self, enabled: bool):
        self._check(self.lib.axr_set_depth_read(self.h, 1 if enabled else 0))

    def alloc_shared(self, nbytes: int):
        """(device pointer, 64-byte CUDA IPC handle) of a fresh allocation other ranks can map with open_ipc."""
        p, h = C.c_void_p(), C.create_string_buffer(64)
        self._check(self.lib.axr_alloc_shared(self.h, nbytes, C.byref(p), h))
        return int(p.value), h.raw

    def free_shared(self, ptr: int):
        self._check(self.lib.axr_free_shared(self.h, ptr))

    def close_ipc(self, ptr: int):
        self._check(self.lib.axr_close_ipc(self.h, ptr))

    # --- convenience: load a scenes.Scene and draw it
    def load_scene(self, scene) -> int:
        tex = [self.upload_texture(t) if t is not None else NO_TEXTURE for t in scene.textures]
        mesh = self.upload_mesh(scene.vertices, scene.indices)
        self.set_material(mesh, 0, *tex, specular_exponent=scene.specular_exponent)
        self.set_uniforms(scene.view_proj, scene.cam_pos)
        self.set_shader(scene.shader, scene.light_dir, scene.light_color)
        self.set_sampler(scene.sampler)
        return mesh


def parse_mtl(text: bytes):
    """axr_parse_mtl: [(name, Ns, {slot: path})] for the five texture slots, in file order."""
    lib = load_library()
    n = C.c_uint32(0)
    rc = lib.axr_parse_mtl(text, len(text), None, 0, C.byref(n))
    if rc:
        raise AxrError(f"axr_parse_mtl failed: {rc}")
    arr = (MtlEntry * max(1, n.value))()
    lib.axr_parse_mtl(text, len(text), arr, n.value, C.byref(n))
    slots = ("diffuse", "bump", "metallic", "roughness", "ao")
    return [(arr[i].name.decode("utf-8", "replace"), float(arr[i].specular_exponent),
             {slots[k]: bytes(arr[i].map[k]).split(b"\0", 1)[0].decode("utf-8", "replace") for k in range(5) if arr[i].has_map[k]})
            for i in range(n.value)]


# ------------------------------------------------------------------------------------------- present
def present(color_bgra, path: str) -> None:
    """Window::present for a headless run (reference src/window.cpp:70-84): the reference copies Framebuffer::getColorData()
    into a bottom-up 32-bit DIB (biHeight > 0, src/windows_bitmap.cpp:19-23), i.e. row 0 of the framebuffer is the BOTTOM row
    of the picture and the bytes are B,G,R,A. This writes the same picture as a PNG (top-down, R,G,B,A); zlib only."""
    import struct
    import zlib
    a = np.asarray(color_bgra, dtype=np.uint8)
    if a.ndim != 3 or a.shape[2] != 4:
        raise ValueError("present: expected an HxWx4 BGRA8 array")
    h, w = a.shape[:2]
    rgba = a[::-1, :, [2, 1, 0, 3]]
    raw = np.empty((h, 1 + w * 4), dtype=np.uint8)
    raw[:, 0] = 0  # filter type None for every scanline
    raw[:, 1:] = rgba.reshape(h, w * 4)

    def chunk(tag: bytes, data: bytes) -> bytes:
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    png = (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6, 0, 0, 0))
           + chunk(b"IDAT", zlib.compress(raw.tobytes(), 6)) + chunk(b"IEND", b""))
    with open(path, "wb") as f:
        f.write(png)


# ------------------------------------------------------------------------------------------- reference-shaped classes
@dataclass
class Color:
    r: int = 0
    g: int = 0
    b: int = 0
    a: int = 255


class _PinnedBlock:
    """cudaHostAlloc'ed bytes exposed through __array_interface__; freed when the last numpy view is gone."""

    def __init__(self, lib, nbytes: int):
        self.lib, self.nbytes = lib, nbytes
        self.ptr = lib.axr_host_alloc(nbytes)
        if self.ptr:
            self.__array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(self.ptr), False), "version": 3}

    def __del__(self):
        if getattr(self, "ptr", None):
            try:
                self.lib.axr_host_free(self.ptr)
            except Exception:
                pass
            self.ptr = None


class Framebuffer:
    """AR::Framebuffer (reference include/framebuffer.hpp, src/framebuffer.cpp): BGRA8 colour + f32 depth on the host,
    row 0 = bottom. Backed by pinned memory when the CUDA library is available."""

    def __init__(self, width: int, height: int, useDepth: bool = True, pinned: bool = True):
        self._w, self._h, self._use_depth = int(width), int(height), bool(useDepth)
        self._color = self._alloc((height, width, 4), np.uint8, pinned)
        self._color[:] = 0
        self._depth = self._alloc((height, width), np.float32, pinned)
        self._depth[:] = np.inf

    def _alloc(self, shape, dtype, pinned):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        if pinned and os.path.exists(LIB_PATH):
            try:
                block = _PinnedBlock(load_library(), n)
                if block.ptr:
                    # the block is the array's base object: the pinned memory lives as long as any view of it does
                    return np.asarray(block).view(dtype).reshape(shape)
            except OSError:
                pass
        return np.zeros(shape, dtype=dtype)

    def clearColor(self, color: Color):
        packed = (color.a << 24) | (color.r << 16) | (color.g << 8) | color.b  # src/framebuffer.cpp:29-32
        self._color.view(np.uint32)[:] = packed

    def clearDepth(self, depthValue: float = float("inf")):
        if self._use_depth:
            self._depth[:] = depthValue

    def getColorData(self) -> np.ndarray:
        return self._color

    def getDepthData(self) -> np.ndarray:
        return self._depth

    def getWidth(self) -> int:
        return self._w

    def getHeight(self) -> int:
        return self._h

    def isDepthBufferEnabled(self) -> bool:
        return self._use_depth

    def present(self, path: str) -> None:
        """Window::present(framebuffer) without a window: PNG of the colour plane (api.present)."""
        present(self._color, path)


class Camera:
    """AR::Camera reduced to what the pipeline reads (reference src/tiled_pipeline.cpp:148-155)."""

    def __init__(self, position=(0.0, 0.0, 0.0), target=(0.0, 0.0, -1.0), fov: float = 60.0, aspectRatio: float = 16.0 / 9.0):
        self._pos = np.asarray(position, dtype=np.float32)
        self._target = np.asarray(target, dtype=np.float32)
        self._fov = min(max(fov, 1.0), 179.0)
        self._aspect = max(aspectRatio, 0.1)
        self._viewport = (0, 0, 800, 600)  # include/camera.hpp:56-59
        self._vp = None
        self.update(0.0)

    def setViewport(self, x: int, y: int, width: int, height: int):
        self._viewport = (x, y, width, height)

    def setViewProjectionMatrix(self, m):
        """Inject a host-computed matrix (what the parity tests do so both sides see identical inputs)."""
        self._vp = np.asarray(m, dtype=np.float32).reshape(4, 4)

    def update(self, deltaTime: float):
        p = _scenes.perspective(self._fov, self._aspect, 0.1, 100.0)
        v = _scenes.look_at(self._pos.astype(np.float64), self._target.astype(np.float64))
        self._vp = _scenes.mat_mul(p, v).astype(np.float32)

    def getViewProjectionMatrix(self) -> np.ndarray:
        return self._vp

    def getViewportMatrix(self) -> np.ndarray:
        x, y, w, h = self._viewport
        m = _scenes.viewport_matrix(w, h)
        m[3][0] += x
        m[3][1] += y
        return m.astype(np.float32)

    def getPosition(self) -> np.ndarray:
        return self._pos


class Texture:
    """AR::Texture: RGBA8, row 0 = image top (reference src/texture.cpp:24)."""

    def __init__(self, rgba: np.ndarray):
        self.data = np.ascontiguousarray(rgba, dtype=np.uint8)

    def getWidth(self):
        return self.data.shape[1]

    def getHeight(self):
        return self.data.shape[0]


@dataclass
class Material:
    name: str = ""
    diffuseTexture: Texture | None = None
    metallicTexture: Texture | None = None
    bumpTexture: Texture | None = None
    roughnessTexture: Texture | None = None
    aoTexture: Texture | None = None
    specularExponent: float = 0.0


@dataclass
class MaterialGroup:
    materialName: str
    startIndex: int
    faceCount: int


class Mesh:
    """AR::Mesh as the pipeline consumes it: vertices (V,14) f32, faces (T,3) u32, material groups, materials by name."""

    def __init__(self, vertices, indices, materials: dict | None = None, groups: list | None = None):
        self._v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 14)
        self._f = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1, 3)
        self._materials = materials if materials is not None else {"default": Material("default")}
        self._groups = groups if groups is not None else [MaterialGroup(next(iter(self._materials)), 0, self._f.shape[0])]
        self._version = 0

    def invalidate(self):
        """Tell the pipelines that hold a device copy of this mesh that its arrays were edited in place (the reference reads the
        host Mesh on every drawMesh; here it is uploaded once and cached): the next drawMesh re-sends it."""
        self._version += 1

    def getVertices(self):
        return self._v

    def getFaces(self):
        return self._f

    def getMaterialGroups(self):
        return self._groups

    def getMaterial(self, matName: str) -> Material:
        return self._materials[matName]  # KeyError like unordered_map::at (reference include/mesh.hpp:78-80)


class IShader:
    kind = None


@dataclass
class FlatShader(IShader):
    lightDirection: tuple = (0.0, -1.0, 0.0)
    kind = SHADER_FLAT


@dataclass
class PhongShader(IShader):
    lightDirection: tuple = (0.0, -1.0, 0.0)
    lightColor: tuple = (1.0, 1.0, 1.0)
    kind = SHADER_PHONG


@dataclass
class PBRShader(IShader):
    lightDirection: tuple = (0.0, -1.0, 0.0)
    lightColor: tuple = (1.0, 1.0, 1.0)
    kind = SHADER_PBR


@dataclass
class CutoutShader(IShader):
    """Not a reference shader: alpha-tested Lambert written against the reference's IShader contract (oracle/ref_harness.cpp),
    the one shader here whose fragment() discards (texel alpha < 0.5). Draws with it are depth-peeled (axr_b200.h)."""
    lightDirection: tuple = (0.0, -1.0, 0.0)
    kind = SHADER_CUTOUT


class Pipeline:
    """AR::Pipeline setters (reference src/pipeline.cpp:16-37)."""

    def __init__(self, cam: Camera | None, fb: Framebuffer | None):
        self.m_Camera, self.m_Framebuffer, self.m_Shader = cam, fb, None

    def setShader(self, shader):
        self.m_Shader = shader

    def setCamera(self, cam):
        self.m_Camera = cam

    def setFramebuffer(self, fb):
        self.m_Framebuffer = fb

    def getViewportMat(self):
        return self.m_Camera.getViewportMatrix()


class TiledPipeline(Pipeline):
    """AR::TiledPipeline(threads, camera, framebuffer) over the CUDA path. `threadsAvailable` is accepted and ignored."""

    def __init__(self, threadsAvailable: int, cam: Camera | None, fb: Framebuffer | None, device: int = 0,
                 sampler: int = SAMPLER_NEAREST):
        super().__init__(cam, fb)
        self._device_index, self._sampler = device, sampler
        self.device: Device | None = None
        self._mesh_cache = {}
        self._tex_cache = {}
        self.last_h2d_bytes = 0
        self.last_d2h_bytes = 0

    def _ensure_device(self):
        fb = self.m_Framebuffer
        if self.device is None or (self.device.width, self.device.height) != (fb.getWidth(), fb.getHeight()):
            if self.device is not None:
                self.device.close()
            self.device = Device(fb.getWidth(), fb.getHeight(), self._device_index, self._sampler)
            self.device.set_overlap(True)  # the geometry stages run while the host framebuffer is still being uploaded
            self._mesh_cache.clear()
            self._tex_cache.clear()

    def _texture(self, t: Texture | None) -> int:
        if t is None:
            return NO_TEXTURE
        k = id(t)
        if k not in self._tex_cache:
            self._tex_cache[k] = (self.device.upload_texture(t.data), t)
            self.last_h2d_bytes += t.data.nbytes
        return self._tex_cache[k][0]

    host_path_description = (
        "TiledPipeline.drawMesh(model, mesh) on a pinned host Framebuffer (axr_draw_mesh_host): nothing is uploaded — the merge test "
        "reads the host depth of the visible pixels through a zero-copy mapping (128 B row reads over PCIe, counted in h2d_bytes) and "
        "the pixels that pass are stored by the tile kernel straight into the host arrays (8 B per updated pixel, 128 B row stores), "
        "complete on return; host-side clearColor/clearDepth before the call are the caller's and untimed, as in the CPU arm; "
        "mesh/textures cached on the device after the first call (Mesh.invalidate() re-sends an edited mesh)")

    def invalidate(self, mesh: Mesh):
        """The adapter's counterpart of Mesh.invalidate(): forget the device copy of `mesh` (it is re-uploaded by the next drawMesh)."""
        ent = self._mesh_cache.pop(id(mesh), None)
        if ent is not None and self.device is not None:
            self.device.free_mesh(ent[0])

    def _mesh(self, mesh: Mesh) -> int:
        """The reference reads the host Mesh on every call; here it is uploaded on first use and cached by identity + version:
        a mesh whose vertices were edited in place (Mesh.invalidate()) with unchanged counts is re-sent into the same device
        buffers, any other change re-uploads it."""
        k = id(mesh)
        ent = self._mesh_cache.get(k)
        if ent is not None and ent[2] != getattr(mesh, "_version", 0):
            if ent[3] == (mesh.getVertices().shape[0], mesh.getFaces().shape[0]) and np.array_equal(ent[4], mesh.getFaces()):
                self.device.update_mesh_vertices(ent[0], mesh.getVertices())
                self.last_h2d_bytes += mesh.getVertices().nbytes
                self._mesh_cache[k] = (ent[0], mesh, mesh._version, ent[3], ent[4])
            else:
                self.invalidate(mesh)
        if k not in self._mesh_cache:
            groups = [(g.startIndex, g.faceCount) for g in mesh.getMaterialGroups()]
            if not groups:  # no group, nothing to draw (the C ABI reads "no groups given" as one group over every face)
                groups = [(mesh.getFaces().shape[0], 0)]
            h = self.device.upload_mesh(mesh.getVertices(), mesh.getFaces(), groups)
            self.last_h2d_bytes += mesh.getVertices().nbytes + mesh.getFaces().nbytes
            for gi, g in enumerate(mesh.getMaterialGroups()):
                m = mesh.getMaterial(g.materialName)
                self.device.set_material(h, gi, self._texture(m.diffuseTexture), self._texture(m.bumpTexture),
                                         self._texture(m.metallicTexture), self._texture(m.roughnessTexture),
                                         self._texture(m.aoTexture), float(m.specularExponent))
            self._mesh_cache[k] = (h, mesh, getattr(mesh, "_version", 0), (mesh.getVertices().shape[0], mesh.getFaces().shape[0]),
                                   mesh.getFaces().copy())
        return self._mesh_cache[k][0]

    def drawMesh(self, modelMatrix, mesh: Mesh):
        if self.m_Shader is None or self.m_Camera is None or self.m_Framebuffer is None:
            return  # reference src/tiled_pipeline.cpp:146
        if getattr(self.m_Shader, "kind", None) not in (SHADER_FLAT, SHADER_PHONG, SHADER_PBR, SHADER_CUTOUT):
            raise AxrError(-6, "no device functor for this IShader subclass (there is no CPU fallback)")
        self._ensure_device()
        self.last_h2d_bytes = 0
        fb, cam, dev = self.m_Framebuffer, self.m_Camera, self.device
        h = self._mesh(mesh)
        dev.set_uniforms(cam.getViewProjectionMatrix(), cam.getPosition(), cam.getViewportMatrix())
        sh = self.m_Shader
        dev.set_shader(sh.kind, sh.lightDirection, getattr(sh, "lightColor", (1.0, 1.0, 1.0)))
        # nothing is uploaded: the host depth of the visible pixels is read through the zero-copy mapping (4 B each), the pixels
        # that pass the depth test are stored the same way (8 B each)
        dev.draw_mesh_host(h, modelMatrix, fb.getColorData(), fb.getDepthData())
        self.last_h2d_bytes += 16 * 4 * 3 + 12  # uniforms; plus 4 B per visible pixel read through the mapping (the caller knows how many)
        self.last_d2h_bytes = None  # 8 bytes per updated pixel; the caller knows how many pixels changed


def render_scene(scene, device: int = 0, color=None, depth=None, band=None, dev: Device | None = None, color_math: int | None = None):
    """Draw a scenes.Scene once on a cleared (or given) framebuffer through the C ABI. Returns (BGRA, depth, stats)."""
    own = dev is None
    if own:
        dev = Device(scene.width, scene.height, device, scene.sampler, band)
    try:
        if color_math is not None:
            dev.set_color_math(color_math)
        mesh = dev.load_scene(scene)
        if color is None:
            dev.clear(0xFF000000, float("inf"))
        else:
            dev.upload_framebuffer(np.ascontiguousarray(color, dtype=np.uint8), np.ascontiguousarray(depth, dtype=np.float32))
        dev.draw_mesh(mesh, scene.model)
        c, d = dev.resolve()
        return c, d, dev.stats()
    finally:
        if own:
            dev.close()

"""Deterministic synthetic scenes for the AxiomR tiled-raster path.

Everything here is host-side input generation (numpy): vertex arrays in the reference's `Vertex`
layout (reference include/mesh.hpp:9-18: position3, uv2, normal3, tangent3, bitangent3 = 14 f32 = 56 B),
u32 triangle indices (3 per face, CCW = front facing in the reference's y-up screen space,
reference src/tiled_pipeline.cpp:89-119), RGBA8 textures with row 0 = image top (what stbi_load
returns, reference src/texture.cpp:24) and column-major float32 matrices as glm stores them.
No rendering happens here.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

VERTEX_FLOATS = 14  # reference include/mesh.hpp:9-18


# --------------------------------------------------------------------------- matrices (column-major, glm conventions)
def _f32(a):
    return np.asarray(a, dtype=np.float32)


def perspective(fovy_deg: float, aspect: float, near: float, far: float) -> np.ndarray:
    """glm::perspective RH, clip z in [-1,1] (reference src/camera.cpp:77). Returns (4,4) with m[c][r]."""
    t = math.tan(math.radians(fovy_deg) / 2.0)
    m = np.zeros((4, 4), dtype=np.float64)
    m[0][0] = 1.0 / (aspect * t)
    m[1][1] = 1.0 / t
    m[2][2] = -(far + near) / (far - near)
    m[2][3] = -1.0
    m[3][2] = -(2.0 * far * near) / (far - near)
    return m


def look_at(eye, target, up=(0.0, 1.0, 0.0)) -> np.ndarray:
    eye = np.asarray(eye, dtype=np.float64)
    f = np.asarray(target, dtype=np.float64) - eye
    f /= np.linalg.norm(f)
    s = np.cross(f, np.asarray(up, dtype=np.float64))
    s /= np.linalg.norm(s)
    u = np.cross(s, f)
    m = np.eye(4, dtype=np.float64)
    m[0][0], m[1][0], m[2][0] = s
    m[0][1], m[1][1], m[2][1] = u
    m[0][2], m[1][2], m[2][2] = -f
    m[3][0], m[3][1], m[3][2] = -np.dot(s, eye), -np.dot(u, eye), np.dot(f, eye)
    return m


def mat_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Column-major product a*b for arrays indexed m[c][r]."""
    return (a.T @ b.T).T


def rotate_y(angle: float) -> np.ndarray:
    c, s = math.cos(angle), math.sin(angle)
    m = np.eye(4, dtype=np.float64)
    m[0][0], m[0][2] = c, -s
    m[2][0], m[2][2] = s, c
    return m


def translate(x, y, z) -> np.ndarray:
    m = np.eye(4, dtype=np.float64)
    m[3][0], m[3][1], m[3][2] = x, y, z
    return m


def viewport_matrix(w: int, h: int) -> np.ndarray:
    """Camera::getViewportMatrix (reference src/camera.cpp:190-204); unused by the raster maths but part of the uniforms."""
    m = np.eye(4, dtype=np.float64)
    m[0][0] = w * 0.5
    m[1][1] = -h * 0.5
    m[2][2] = 0.5
    m[3][0] = w * 0.5
    m[3][1] = h * 0.5
    m[3][2] = 0.5
    return m


def default_camera(w: int, h: int, eye=(0.0, 0.0, 5.0), target=(0.0, 0.0, 0.0), fov=60.0):
    """The app's camera (reference src/renderer.cpp:68-69, include/camera.hpp:45-48): returns (viewProj, camPos) as f32."""
    vp = mat_mul(perspective(fov, w / h, 0.1, 100.0), look_at(eye, target))
    return _f32(vp), _f32(eye)


# --------------------------------------------------------------------------- textures
def diffuse_texture(size: int) -> np.ndarray:
    y, x = np.mgrid[0:size, 0:size].astype(np.int64)
    t = np.empty((size, size, 4), dtype=np.uint8)
    t[..., 0] = (x ^ y) & 255
    t[..., 1] = (7 * x + 13 * y) & 255
    t[..., 2] = np.where((((x >> 4) + (y >> 4)) & 1) == 1, 230, 40)
    t[..., 3] = 255
    return t


def cutout_texture(size: int, cell: int = 8) -> np.ndarray:
    """Diffuse map with an alpha channel for the discard test (alpha < 0.5 is cut): holes on a checkerboard of `cell` texels,
    plus a band of alpha values 120..135 that straddle the threshold (127/255 < 0.5 <= 128/255)."""
    t = diffuse_texture(size)
    y, x = np.mgrid[0:size, 0:size].astype(np.int64)
    a = np.where((((x // cell) + (y // cell)) & 1) == 1, 255, 0)
    band = (y >= size // 2) & (y < size // 2 + max(1, size // 16))
    a = np.where(band, 120 + (x & 15), a)
    t[..., 3] = a.astype(np.uint8)
    return t


def normal_texture(size: int, bumps: int = 8, strength: float = 4.0) -> np.ndarray:
    y, x = np.mgrid[0:size, 0:size].astype(np.float64)
    kx = 2.0 * math.pi * bumps / size
    dhdx = 0.5 * kx * np.cos(kx * x) * np.sin(kx * y)
    dhdy = 0.5 * kx * np.sin(kx * x) * np.cos(kx * y)
    n = np.stack([-dhdx * strength, -dhdy * strength, np.ones_like(x)], axis=-1)
    n /= np.linalg.norm(n, axis=-1, keepdims=True)
    t = np.empty((size, size, 4), dtype=np.uint8)
    t[..., :3] = np.round((n * 0.5 + 0.5) * 255.0).astype(np.uint8)
    t[..., 3] = 255
    return t


def scalar_texture(size: int, lo: int, hi: int, freq: int) -> np.ndarray:
    """Single-channel procedural map replicated to RGBA (metallic / roughness / ao for PBRShader)."""
    y, x = np.mgrid[0:size, 0:size].astype(np.float64)
    v = 0.5 + 0.5 * np.sin(2 * math.pi * freq * x / size) * np.cos(2 * math.pi * freq * y / size)
    c = np.round(lo + (hi - lo) * v).astype(np.uint8)
    t = np.empty((size, size, 4), dtype=np.uint8)
    t[..., 0] = t[..., 1] = t[..., 2] = c
    t[..., 3] = 255
    return t


# --------------------------------------------------------------------------- meshes
def _pack(pos, uv, nrm, tan, bit) -> np.ndarray:
    v = np.empty((pos.shape[0], VERTEX_FLOATS), dtype=np.float32)
    v[:, 0:3] = pos
    v[:, 3:5] = uv
    v[:, 5:8] = nrm
    v[:, 8:11] = tan
    v[:, 11:14] = bit
    return v


def _unit(a):
    n = np.linalg.norm(a, axis=-1, keepdims=True)
    n[n == 0] = 1.0
    return a / n


def icosphere(level: int, radius: float = 2.0):
    """Subdivided icosahedron: T = 20*4^level, V = 10*4^level + 2 (level 8: 1 310 720 / 655 362)."""
    p = (1.0 + math.sqrt(5.0)) / 2.0
    verts = np.array([[-1, p, 0], [1, p, 0], [-1, -p, 0], [1, -p, 0], [0, -1, p], [0, 1, p], [0, -1, -p], [0, 1, -p],
                      [p, 0, -1], [p, 0, 1], [-p, 0, -1], [-p, 0, 1]], dtype=np.float64)
    verts = _unit(verts)
    faces = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
                      [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5],
                      [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    for _ in range(level):
        nv = verts.shape[0]
        e = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]], axis=0)
        es = np.sort(e, axis=1)
        key = es[:, 0] * nv + es[:, 1]
        uniq, inv = np.unique(key, return_inverse=True)
        mid = _unit(verts[uniq // nv] + verts[uniq % nv])
        verts = np.concatenate([verts, mid], axis=0)
        nf = faces.shape[0]
        m01, m12, m20 = nv + inv[:nf], nv + inv[nf:2 * nf], nv + inv[2 * nf:]
        a, b, c = faces[:, 0], faces[:, 1], faces[:, 2]
        # the four children of a face stay adjacent in the face list (the order a recursive subdivision produces)
        faces = np.stack([np.stack([a, m01, m20], 1), np.stack([b, m12, m01], 1),
                          np.stack([c, m20, m12], 1), np.stack([m01, m12, m20], 1)], axis=1).reshape(-1, 3)
    nrm = verts
    pos = verts * radius
    u = 0.5 + np.arctan2(nrm[:, 2], nrm[:, 0]) / (2 * math.pi)
    v = 0.5 + np.arcsin(np.clip(nrm[:, 1], -1, 1)) / math.pi
    tan = _unit(np.stack([-nrm[:, 2], np.zeros_like(u), nrm[:, 0]], 1) + 1e-30)
    tan[np.linalg.norm(tan, axis=1) < 0.5] = (1.0, 0.0, 0.0)
    bit = np.cross(nrm, tan)
    return _pack(pos, np.stack([u, v], 1), nrm, tan, bit), faces.astype(np.uint32)


def torus(segs: int, loops: int, R: float = 2.0, r: float = 0.8):
    """Torus grid around the z axis: T = 2*segs*loops, V = (segs+1)*(loops+1) (2236^2: 9 999 392 / 5 004 169)."""
    i = np.arange(segs + 1, dtype=np.float64) / segs
    j = np.arange(loops + 1, dtype=np.float64) / loops
    th, ph = np.meshgrid(i * 2 * math.pi, j * 2 * math.pi, indexing="ij")  # th: around the ring, ph: around the tube
    ct, st, cp, sp = np.cos(th), np.sin(th), np.cos(ph), np.sin(ph)
    pos = np.stack([(R + r * cp) * ct, (R + r * cp) * st, r * sp], -1).reshape(-1, 3)
    nrm = np.stack([cp * ct, cp * st, sp], -1).reshape(-1, 3)
    tan = np.stack([-st, ct, np.zeros_like(st)], -1).reshape(-1, 3)
    bit = np.cross(nrm, tan)
    uv = np.stack(np.meshgrid(i, j, indexing="ij"), -1).reshape(-1, 2)
    a = (np.arange(segs, dtype=np.int64)[:, None] * (loops + 1) + np.arange(loops, dtype=np.int64)[None, :]).reshape(-1)
    b = a + (loops + 1)
    f = np.empty((a.size * 2, 3), dtype=np.int64)
    f[0::2] = np.stack([a, b, b + 1], 1)
    f[1::2] = np.stack([a, b + 1, a + 1], 1)
    return _pack(pos, uv, nrm, tan, bit), f.astype(np.uint32)


def head_like(stacks: int = 36, slices: int = 35, radii=(1.2, 1.6, 1.4)):
    """Displaced ellipsoid with a nose bump: 2*slices*(stacks-1) faces (36x35 -> 2 450), SURVEY.md §8(d) C1."""
    rows = []
    for s in range(stacks + 1):
        lat = math.pi * s / stacks
        for k in range(slices + 1):
            lon = 2 * math.pi * k / slices
            d = np.array([math.sin(lat) * math.sin(lon), math.cos(lat), math.sin(lat) * math.cos(lon)])
            bump = 0.25 * math.exp(-((d[0]) ** 2 + (d[1] + 0.05) ** 2) / 0.02) if d[2] > 0 else 0.0
            p = d * np.asarray(radii) * (1.0 + bump)
            rows.append((p, (k / slices, 1.0 - s / stacks), d))
    pos = np.array([r[0] for r in rows])
    uv = np.array([r[1] for r in rows])
    nrm = _unit(np.array([r[2] for r in rows]) / np.asarray(radii))
    tan = _unit(np.stack([nrm[:, 2], np.zeros(len(rows)), -nrm[:, 0]], 1) + 1e-30)
    tan[np.linalg.norm(tan, axis=1) < 0.5] = (1.0, 0.0, 0.0)
    bit = np.cross(nrm, tan)
    f = []
    w = slices + 1
    for s in range(stacks):
        for k in range(slices):
            a, b, c, d = s * w + k, s * w + k + 1, (s + 1) * w + k, (s + 1) * w + k + 1
            if s != 0:
                f.append((a, c, b))
            if s != stacks - 1:
                f.append((b, c, d))
    return _pack(pos, uv, nrm, tan, bit), np.asarray(f, dtype=np.uint32)


def quad_grid(n: int, size: float = 2.0, z: float = 0.0):
    """n x n quads in the z plane facing +z (2 n^2 faces)."""
    g = np.arange(n + 1, dtype=np.float64) / n
    u, v = np.meshgrid(g, g, indexing="xy")
    pos = np.stack([(u - 0.5) * size, (v - 0.5) * size, np.full_like(u, z)], -1).reshape(-1, 3)
    k = pos.shape[0]
    nrm = np.tile([0.0, 0.0, 1.0], (k, 1))
    tan = np.tile([1.0, 0.0, 0.0], (k, 1))
    bit = np.tile([0.0, 1.0, 0.0], (k, 1))
    uv = np.stack([u, v], -1).reshape(-1, 2)
    a = (np.arange(n)[:, None] * (n + 1) + np.arange(n)[None, :]).reshape(-1)
    f = np.empty((a.size * 2, 3), dtype=np.int64)
    f[0::2] = np.stack([a, a + 1, a + n + 2], 1)
    f[1::2] = np.stack([a, a + n + 2, a + n + 1], 1)
    return _pack(pos, uv, nrm, tan, bit), f.astype(np.uint32)


def random_triangles(n: int, seed: int, extent: float = 3.0, size: float = 0.5, zspread: float = 2.0):
    """Unindexed random triangles (both windings) scattered through and beyond the frustum: stresses clip + cull + ties."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(-extent, extent, (n, 1, 3))
    c[..., 2] = rng.uniform(-zspread, zspread, (n, 1))
    p = (c + rng.normal(0.0, size, (n, 3, 3))).reshape(-1, 3)
    nrm = _unit(rng.normal(0, 1, (n * 3, 3)))
    tan = _unit(np.cross(nrm, rng.normal(0, 1, (n * 3, 3))))
    bit = np.cross(nrm, tan)
    uv = rng.uniform(0, 1, (n * 3, 2))
    return _pack(p, uv, nrm, tan, bit), np.arange(n * 3, dtype=np.uint32).reshape(-1, 3)


# --------------------------------------------------------------------------- scene container
SHADER_FLAT, SHADER_PHONG, SHADER_PBR, SHADER_CUTOUT = 0, 1, 2, 3  # 3: alpha-tested Lambert, the one shader that discards
SAMPLER_NEAREST, SAMPLER_BILINEAR = 0, 1


@dataclass
class Scene:
    name: str
    width: int
    height: int
    vertices: np.ndarray            # (V,14) f32
    indices: np.ndarray             # (T,3) u32
    shader: int = SHADER_FLAT
    sampler: int = SAMPLER_NEAREST
    model: np.ndarray = field(default_factory=lambda: np.eye(4, dtype=np.float32))
    view_proj: np.ndarray | None = None
    cam_pos: np.ndarray | None = None
    light_dir: np.ndarray = field(default_factory=lambda: _f32(_unit(np.array([-0.3, -1.0, -0.5]))))
    light_color: np.ndarray = field(default_factory=lambda: _f32([0.6, 0.6, 0.6]))
    specular_exponent: float = 0.2  # "Ns"; PhongShader uses Ns*50 (reference include/shaders/shaders.hpp:231)
    textures: list = field(default_factory=lambda: [None] * 5)  # diffuse, bump, metallic, roughness, ao

    def __post_init__(self):
        if self.view_proj is None:
            self.view_proj, self.cam_pos = default_camera(self.width, self.height)
        self.vertices = np.ascontiguousarray(self.vertices, dtype=np.float32)
        self.indices = np.ascontiguousarray(self.indices, dtype=np.uint32)
        self.model = np.ascontiguousarray(self.model, dtype=np.float32)
        self.view_proj = np.ascontiguousarray(self.view_proj, dtype=np.float32)
        self.cam_pos = np.ascontiguousarray(self.cam_pos, dtype=np.float32)

    @property
    def n_faces(self) -> int:
        return int(self.indices.shape[0])

    @property
    def n_verts(self) -> int:
        return int(self.vertices.shape[0])


def _phong_textures(size: int):
    return [diffuse_texture(size), normal_texture(size), None, None, None]


def _pbr_textures(size: int):
    return [diffuse_texture(size), normal_texture(size), scalar_texture(size, 0, 255, 3),
            scalar_texture(size, 30, 230, 5), scalar_texture(size, 128, 255, 2)]


def config1(tex: int = 1024) -> Scene:
    """C1: head-like 2 450-tri mesh at 800x800, PhongShader nearest."""
    v, f = head_like()
    return Scene("c1_head_phong_800", 800, 800, v, f, SHADER_PHONG, model=_f32(rotate_y(0.5)), textures=_phong_textures(tex))


def config2(level: int = 8, w: int = 1920, h: int = 1080) -> Scene:
    """C2: icosphere k=8 (1.31 M tris) at 1920x1080, FlatShader (Lambert), no textures."""
    v, f = icosphere(level)
    return Scene(f"c2_icosphere{level}_flat_{w}x{h}", w, h, v, f, SHADER_FLAT, model=_f32(rotate_y(0.5)))


def config3(n: int = 2236, w: int = 3840, h: int = 2160, tex: int = 4096, sampler: int = SAMPLER_NEAREST) -> Scene:
    """C3: torus n x n (10 M tris) at 4K, 4096^2 diffuse + normal map, normal-mapped PhongShader."""
    v, f = torus(n, n)
    return Scene(f"c3_torus{n}_phong_{w}x{h}", w, h, v, f, SHADER_PHONG, sampler, model=_f32(rotate_y(0.5)),
                 textures=_phong_textures(tex))


def config4(n: int = 3162, w: int = 3840, h: int = 2160) -> Scene:
    """C4: torus n x n (20 M sub-pixel tris) at 4K, FlatShader."""
    v, f = torus(n, n)
    return Scene(f"c4_torus{n}_flat_{w}x{h}", w, h, v, f, SHADER_FLAT, model=_f32(rotate_y(0.5)))


def cutout_layers(w: int = 320, h: int = 240, layers: int = 5, grid: int = 6, tex: int = 64, size: float = 2.6,
                  sampler: int = SAMPLER_NEAREST) -> Scene:
    """Discard test scene: `layers` textured quad grids stacked in depth (nearest drawn LAST, so every layer's fragments reach the
    shader in the reference), alpha-tested; through the holes of one layer the next one shows, through all of them the clear."""
    vs, fs, base = [], [], 0
    for i in range(layers):
        v, f = quad_grid(grid, size=size - 0.2 * i, z=-1.0 + 0.45 * i)
        v = v.copy()
        v[:, 3:5] = v[:, 3:5] * (0.55 + 0.05 * i) + 0.04 * i  # a different uv window per layer, inside [0, 1]
        vs.append(v); fs.append(f + base); base += v.shape[0]
    return Scene(f"cutout_{layers}layers_{w}x{h}", w, h, np.concatenate(vs), np.concatenate(fs), SHADER_CUTOUT, sampler,
                 model=_f32(rotate_y(0.35)), textures=[cutout_texture(tex), None, None, None, None])


def view_matrix_for(i: int, n_views: int, w: int, h: int):
    """C5 'views' mode: camera i of n orbiting the origin (yaw = i*2pi/n)."""
    a = 2 * math.pi * i / n_views
    eye = (5.0 * math.sin(a), 0.0, 5.0 * math.cos(a))
    return default_camera(w, h, eye=eye)

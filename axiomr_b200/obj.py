"""OBJ / MTL ingestion — host-side mirror of AR::Mesh(path) (reference src/mesh.cpp), the step in front of the raster path
(SURVEY.md §8(f) rank 3). Produces exactly the arrays the reference loader produces — vertex order, de-duplication, fan
triangulation, material groups and its tangent / bitangent generation, quirks included — so a scene can be loaded without
the reference and still render bit-identically. All arithmetic is float32 in the reference's operation order.

Not on the per-frame hot path: plain numpy, no GPU work.
"""
from __future__ import annotations

import os

import numpy as np

from .api import Material, MaterialGroup, Mesh, Texture

f32 = np.float32


def _dot(a, b):  # glm compute_dot<vec3>: (x*x' + y*y') + z*z'
    return f32(f32(f32(a[0] * b[0]) + f32(a[1] * b[1])) + f32(a[2] * b[2]))


def _cross(x, y):  # glm::cross
    return np.array([f32(f32(x[1] * y[2]) - f32(y[1] * x[2])), f32(f32(x[2] * y[0]) - f32(y[2] * x[0])),
                     f32(f32(x[0] * y[1]) - f32(y[0] * x[1]))], dtype=f32)


def _normalize(v):  # v * (1 / sqrt(dot(v, v)))
    with np.errstate(divide="ignore", invalid="ignore"):
        return (v * f32(f32(1.0) / np.sqrt(_dot(v, v), dtype=f32))).astype(f32)


def parse_obj(text: str):
    """reference src/mesh.cpp:300-415 (parseModelFile): returns (vertices (V,8) f32 = pos3 uv2 normal3, faces (T,3) u32, groups)."""
    positions, uvs, normals = [], [], []
    verts, faces, groups = [], [], []
    unique = {}
    group_start = 0
    for line in text.splitlines():
        if not line:
            continue
        tok = line.split()
        if not tok:
            continue
        t = tok[0]
        if t == "v":
            positions.append(tuple(f32(x) for x in tok[1:4]))
        elif t == "vt":
            uvs.append(tuple(f32(x) for x in tok[1:3]))
        elif t == "vn":
            normals.append(tuple(f32(x) for x in tok[1:4]))
        elif t == "usemtl":
            if groups:
                groups[-1][2] = len(faces) - group_start
            groups.append([tok[1] if len(tok) > 1 else "", len(faces), 0])
            group_start = len(faces)
        elif t == "f":
            idx = []
            for corner in tok[1:]:
                parts = corner.split("/")
                vi = int(parts[0]) - 1 if parts and parts[0] else -1
                ti = int(parts[1]) - 1 if len(parts) > 1 and parts[1] else -1
                ni = int(parts[2]) - 1 if len(parts) > 2 and parts[2] else -1
                if vi < 0 or vi >= len(positions):
                    continue
                uv = uvs[ti] if 0 <= ti < len(uvs) else (f32(0), f32(0))
                nr = normals[ni] if 0 <= ni < len(normals) else (f32(0), f32(0), f32(0))
                key = (positions[vi], uv, nr)   # Vertex::operator== compares position, uv, normal (include/mesh.hpp:15-17)
                k = unique.get(key)
                if k is None:
                    k = len(verts)
                    unique[key] = k
                    verts.append(positions[vi] + uv + nr)
                idx.append(k)
            for i in range(1, len(idx) - 1):  # fan triangulation (:387-395)
                faces.append((idx[0], idx[i], idx[i + 1]))
        if groups:
            groups[-1][2] = len(faces) - group_start
    v = np.asarray(verts, dtype=f32).reshape(-1, 8)
    f = np.asarray(faces, dtype=np.uint32).reshape(-1, 3)
    return v, f, [MaterialGroup(g[0], g[1], g[2]) for g in groups]


def tangents(v8: np.ndarray, faces: np.ndarray) -> np.ndarray:
    """reference src/mesh.cpp:222-298 (calculateTangentBitangent). Returns (V,14) f32 in AR::Vertex layout."""
    n_v = v8.shape[0]
    pos, uv, nrm = v8[:, 0:3], v8[:, 3:5], v8[:, 5:8]
    tsum = np.zeros((n_v, 3), dtype=f32)
    bsum = np.zeros((n_v, 3), dtype=f32)
    seen = np.zeros(n_v, dtype=bool)
    for a, b, c in faces:  # sequential accumulation in face order, like the std::map version
        v0, v1, v2 = pos[a], pos[b], pos[c]
        e1 = (v1 - v0).astype(f32)
        e2 = (v2 - v0).astype(f32)
        d1 = (uv[b] - uv[a]).astype(f32)
        d2 = (uv[c] - uv[a]).astype(f32)
        den = f32(f32(d1[0] * d2[1]) - f32(d2[0] * d1[1]))
        n0 = nrm[a]
        if abs(den) < f32(1e-8):
            ft = np.array([0, 1, 0], dtype=f32) if abs(n0[0]) > f32(0.8) else np.array([1, 0, 0], dtype=f32)
            fb = _cross(n0, ft)
            for i in (a, b, c):
                tsum[i] = (tsum[i] + ft).astype(f32)
                bsum[i] = (bsum[i] + fb).astype(f32)
                seen[i] = True
            continue
        fi = f32(f32(1.0) / den)
        tan = (fi * ((d2[1] * e1).astype(f32) - (d1[1] * e2).astype(f32)).astype(f32)).astype(f32)
        bit = (fi * ((f32(-d2[0]) * e1).astype(f32) + (d1[0] * e2).astype(f32)).astype(f32)).astype(f32)
        if _dot(_cross(n0, tan), bit) < f32(0.0):
            bit = (-bit).astype(f32)
        for i in (a, b, c):
            tsum[i] = (tsum[i] + tan).astype(f32)
            bsum[i] = (bsum[i] + bit).astype(f32)
            seen[i] = True
    out = np.zeros((n_v, 14), dtype=f32)
    out[:, 0:8] = v8
    with np.errstate(divide="ignore", invalid="ignore"):
        for i in range(n_v):
            n = nrm[i]
            t = tsum[i] if seen[i] else np.array([1, 0, 0], dtype=f32)
            t = (t - (n * _dot(n, t)).astype(f32)).astype(f32)
            t = _normalize(t)
            b = _cross(n, t)
            # handedness against the accumulated bitangent (bitangentMap[index] default-constructs to 0 for unseen vertices)
            hand = f32(-1.0) if _dot(bsum[i], b) < f32(0.0) else f32(1.0)
            b = (b * hand).astype(f32)
            # the `length() < 1e-8` fallbacks (:283,:289) never fire: glm's vec3::length() is the component count, 3
            out[i, 8:11] = t
            out[i, 11:14] = b
    return out


def parse_mtl(text: str, directory: str, texture_loader):
    """reference src/mesh.cpp:65-220 (loadMaterial / parseMaterialData): name -> Material. texture_loader(path) -> RGBA8 array or None."""
    mats, cur = {}, None
    slots = {"map_Kd": "diffuseTexture", "map_Ks": "metallicTexture", "refl": "metallicTexture", "map_Ns": "roughnessTexture",
             "map_Bump": "bumpTexture", "bump": "bumpTexture", "norm": "bumpTexture", "map_A0": "aoTexture"}
    for line in text.splitlines():
        if not line:
            continue
        tok = line.split()
        if not tok:
            continue
        if tok[0] == "newmtl":
            cur = Material(tok[1] if len(tok) > 1 else "")
            mats[cur.name] = cur
        elif cur is not None:
            if tok[0] == "Ns" and len(tok) > 1:
                cur.specularExponent = float(f32(tok[1]))
            elif tok[0] in slots:
                rel = line.split(None, 1)[1].strip(" \t") if len(tok) > 1 else ""
                img = texture_loader(directory + "/" + rel)
                setattr(cur, slots[tok[0]], Texture(img) if img is not None else None)
    return mats


def _load_image(path: str):
    """RGBA8, row 0 = top (what stbi_load(..., STBI_rgb_alpha) returns, reference src/texture.cpp:24)."""
    try:
        from PIL import Image
        return np.ascontiguousarray(np.asarray(Image.open(path).convert("RGBA"), dtype=np.uint8))
    except Exception:
        return None


def load_obj(path: str, texture_loader=_load_image, device=None) -> Mesh:
    """AR::Mesh(path): OBJ + '<basename>.mtl' next to it (reference src/mesh.cpp:8-27,197-220).
    With `device` (an api.Device) the tangents / bitangents are generated on the GPU (axr_generate_tangents, bit-identical);
    the per-corner Python loop of `tangents` is only practical for small meshes."""
    with open(path, "rb") as fh:
        text = fh.read().decode("utf-8", "replace")
    v8, faces, groups = parse_obj(text)
    verts = device.generate_tangents(v8, faces) if device is not None else tangents(v8, faces)
    directory = os.path.dirname(path)
    mtl = os.path.join(directory, os.path.splitext(os.path.basename(path))[0] + ".mtl")
    mats = {}
    if os.path.exists(mtl):
        with open(mtl, "rb") as fh:
            mats = parse_mtl(fh.read().decode("utf-8", "replace"), directory, texture_loader)
    # an OBJ without `usemtl` has no material group: drawMesh walks the groups (reference src/tiled_pipeline.cpp:176-179), so the
    # reference draws nothing of it, and neither does TiledPipeline.drawMesh here (an empty group list stays empty)
    return Mesh(verts, faces, mats if mats else None, groups)

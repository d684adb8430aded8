"""Test helper: writes a scene as OBJ + MTL + 32-bit TGA files that the reference's own Mesh / Texture loaders read
(reference src/mesh.cpp:300-415, src/texture.cpp:21-36)."""
import os

import numpy as np


def write_tga(path, rgba):
    rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
    h, w = rgba.shape[:2]
    hdr = bytearray(18)
    hdr[2] = 2
    hdr[12], hdr[13] = w & 255, w >> 8
    hdr[14], hdr[15] = h & 255, h >> 8
    hdr[16], hdr[17] = 32, 0x28  # 8 alpha bits, top-left origin
    with open(path, "wb") as f:
        f.write(bytes(hdr))
        f.write(rgba[..., [2, 1, 0, 3]].tobytes())


def write_obj_scene(dirpath, name, vertices, indices, textures, ns=0.2):
    """vertices (V,14) f32 in AR::Vertex layout, indices (T,3). textures: [diffuse, bump, metallic, roughness, ao] or None."""
    os.makedirs(dirpath, exist_ok=True)
    v = np.asarray(vertices, dtype=np.float32)
    f = np.asarray(indices, dtype=np.int64) + 1
    with open(os.path.join(dirpath, name + ".obj"), "w") as o:
        for p in v:
            o.write("v %.9g %.9g %.9g\n" % (p[0], p[1], p[2]))
        for p in v:
            o.write("vt %.9g %.9g\n" % (p[3], p[4]))
        for p in v:
            o.write("vn %.9g %.9g %.9g\n" % (p[5], p[6], p[7]))
        o.write("usemtl m0\n")
        for a, b, c in f:
            o.write(f"f {a}/{a}/{a} {b}/{b}/{b} {c}/{c}/{c}\n")
    keys = ["map_Kd", "map_Bump", "map_Ks", "map_Ns", "map_A0"]
    with open(os.path.join(dirpath, name + ".mtl"), "w") as m:
        m.write("newmtl m0\nNs %.9g\n" % ns)
        for k, t in zip(keys, textures or []):
            if t is not None:
                fn = f"{name}_{k}.tga"
                write_tga(os.path.join(dirpath, fn), t)
                m.write(f"{k} {fn}\n")
    return os.path.join(dirpath, name + ".obj")

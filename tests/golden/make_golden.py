"""Regenerates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref, built from /root/reference by
oracle/Makefile) on the deterministic cases of cases.py. Run here (the reference tree is not on the GPU box):

    python tests/golden/make_golden.py            # everything
    python tests/golden/make_golden.py CASE ...   # only the named scene fixtures (existing files are left alone)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

from oracle import pyoracle as po  # noqa: E402
import cases as C  # noqa: E402

po.build(ref=True)
assert po.ref_available(), "oracle/_ref/libaxr_ref.so is needed (requires /root/reference)"
all_cases = C.cases()
assert list(all_cases) == C.CASE_NAMES
only = sys.argv[1:]
assert all(n in all_cases for n in only), only
for name, sc in all_cases.items():
    if only and name not in only:
        continue
    c, d, _ = po.ref_render(sc, threads=3)
    C.save_case(name, sc, c, d)
    print(name, int(np.isfinite(d).sum()), "px covered")
if only:
    sys.exit(0)
# stage-level known answers
tris = C.clip_cases()
clip_out = [po.ref_clip_triangle(t) for t in tris]
setup = [po.ref_triangle_setup(t, 640, 480) for t in tris]
rng = np.random.default_rng(3)
uv = rng.uniform(-0.2, 1.2, (64, 2)).astype(np.float32)
uv[:6] = [[0, 0], [1, 1], [0, 1], [1, 0], [0.5, 0.5], [0.999999, 0.000001]]
tex = np.arange(7 * 5 * 4, dtype=np.uint8).reshape(5, 7, 4)
np.savez_compressed(os.path.join(HERE, "stage_kat.npz"),
                    clip_n=np.array([o.shape[0] for o in clip_out]), clip_out=np.concatenate(clip_out),
                    setup_back=np.array([s[0] for s in setup]), setup_out=np.stack([s[1] for s in setup]),
                    tex=tex, uv=uv, tex_out=po.ref_texture_sample(tex, uv))
vp, vpt = po.ref_camera((0, 0, 5), (0, 0, 0), 60.0, 800, 600)
a = np.arange(16, dtype=np.float32).reshape(4, 4) * 0.37 - 2
out = np.zeros(16, dtype=np.float32)
import ctypes as ct
po.ref_lib().axr_ref_mat4_mul(vp.ctypes.data_as(ct.POINTER(ct.c_float)), a.ctypes.data_as(ct.POINTER(ct.c_float)), out.ctypes.data_as(ct.POINTER(ct.c_float)))
np.savez_compressed(os.path.join(HERE, "camera_kat.npz"), view_proj=vp, viewport=vpt, a=a, vp_times_a=out.reshape(4, 4))
print("stage KATs written")
# OBJ / MTL ingestion through the reference's own loader (src/mesh.cpp): the file texts + the arrays it produced
import tempfile
sys.path.insert(0, os.path.dirname(HERE))
from objutil import write_obj_scene  # noqa: E402
from axiomr_b200 import scenes as S  # noqa: E402
d = tempfile.mkdtemp()
v, f = S.head_like(10, 9)
quad_v, quad_f = S.quad_grid(3)
fix = {}
for name, (vv, ff) in {"head": (v, f), "quad": (quad_v, quad_f)}.items():
    pth = write_obj_scene(d, name, vv, ff, None)
    rv, rf = po.ref_load_obj(pth)
    fix[name + "_obj"] = np.frombuffer(open(pth, "rb").read(), dtype=np.uint8)
    fix[name + "_mtl"] = np.frombuffer(open(pth[:-4] + ".mtl", "rb").read(), dtype=np.uint8)
    fix[name + "_vertices"], fix[name + "_faces"] = rv, rf
# a polygon face (fan triangulation), shared / repeated corners (de-duplication), a face without vt/vn, degenerate uv (fallback tangent)
poly = b"v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 0.5 1.5 0.25\nvt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\nvn 0 0 1\nvn 0.9 0 0.43589\nusemtl m0\n" \
       b"f 1/1/1 2/2/1 3/3/1 4/4/1 5/3/2\nf 1/1/1 3/3/1 2/2/1\nf 1 2 5\nf 1/1/1 2/1/1 3/1/1\n"
open(os.path.join(d, "poly.obj"), "wb").write(poly)
open(os.path.join(d, "poly.mtl"), "wb").write(b"newmtl m0\nNs 0.25\n")
rv, rf = po.ref_load_obj(os.path.join(d, "poly.obj"))
fix["poly_obj"] = np.frombuffer(poly, dtype=np.uint8)
fix["poly_mtl"] = np.frombuffer(b"newmtl m0\nNs 0.25\n", dtype=np.uint8)
fix["poly_vertices"], fix["poly_faces"] = rv, rf
np.savez_compressed(os.path.join(HERE, "obj_loader.npz"), **fix)
print("OBJ loader fixtures written", {k: fix[k].shape for k in fix if k.endswith("vertices")})

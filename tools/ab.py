"""A/B harness for kernel variants (no torch import: starts in seconds on a fresh GPU box).

usage (under gpurun):  python tools/ab.py [--workload c3] [--steps 20] [variant ...]
Variants are the libraries variants_tmp/lib_<name>.so built beforehand (tools/build_variants.py, nvcc on the CPU container);
"default" is axiomr_b200/libaxr_b200.so. For every variant, in one process:
  1. parity spot check: the smoke scene (small + binned + clipped triangles, Phong) against oracle/libaxr_oracle.so,
  2. K timed steps (axr_clear + axr_draw_mesh) of the workload with per-kernel CUDA events (axr_set_profiling).
One JSON line per variant on stdout and in gpurun_out/ab_<workload>.jsonl.
"""
import argparse
import glob
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from axiomr_b200 import api, scenes as S  # noqa: E402


def workload(name: str) -> S.Scene:
    if name == "c3":
        return S.config3(sampler=S.SAMPLER_BILINEAR)
    if name == "c3n":
        return S.config3()
    if name == "c2":
        return S.config2()
    if name == "c4":
        return S.config4()
    if name == "c1":
        return S.config1()
    if name == "c5":
        return S.config3(w=7680, h=4320, sampler=S.SAMPLER_BILINEAR)
    raise SystemExit(f"unknown workload {name}")


def smoke_scene() -> S.Scene:
    v, f = S.random_triangles(1500, 7)
    v2, f2 = S.icosphere(5, 1.5)
    return S.Scene("smoke", 640, 400, np.concatenate([v, v2]), np.concatenate([f, f2 + v.shape[0]]), S.SHADER_PHONG,
                   textures=S._phong_textures(128))


def e2e_ms(sc: S.Scene, steps: int) -> float:
    """The reference-facing call: TiledPipeline.drawMesh(model, mesh) on a host framebuffer, complete on return (bench.py's e2e)."""
    fb = api.Framebuffer(sc.width, sc.height, True)
    cam = api.Camera()
    cam.setViewport(0, 0, sc.width, sc.height)
    cam.setViewProjectionMatrix(sc.view_proj)
    cam._pos = np.asarray(sc.cam_pos, dtype=np.float32)
    pipe = api.TiledPipeline(os.cpu_count() or 1, cam, fb, device=0, sampler=sc.sampler)
    pipe.setShader([api.FlatShader(tuple(sc.light_dir)), api.PhongShader(tuple(sc.light_dir), tuple(sc.light_color)),
                    api.PBRShader(tuple(sc.light_dir), tuple(sc.light_color))][sc.shader])
    tex = [api.Texture(t) if t is not None else None for t in sc.textures]
    mesh = api.Mesh(sc.vertices, sc.indices, {"m0": api.Material("m0", tex[0], tex[2], tex[1], tex[3], tex[4], sc.specular_exponent)})
    total = 0.0
    n = max(3, min(steps, 10))
    for i in range(n + 2):  # the first call uploads and caches the mesh
        fb.clearColor(api.Color(0, 0, 0, 255))
        fb.clearDepth()
        t0 = time.perf_counter()
        pipe.drawMesh(sc.model, mesh)
        if i >= 2:
            total += time.perf_counter() - t0
    pipe.device.close()
    return total / n * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--overlap", action="store_true", help="also time the same steps with axr_set_overlap(1) (wall clock, no per-kernel events)")
    ap.add_argument("--e2e", action="store_true", help="also time TiledPipeline.drawMesh on a pinned host framebuffer (axr_draw_mesh_host), "
                    "with and without AXR_B200_HOST_DEPTH_ZEROCOPY")
    ap.add_argument("variants", nargs="*")
    a = ap.parse_args()
    libs = {"default": os.path.join(ROOT, "axiomr_b200", "libaxr_b200.so")}
    for p in sorted(glob.glob(os.path.join(os.environ.get("AXR_AB_DIR") or os.path.join(ROOT, "variants_tmp"), "lib_*.so"))):
        libs.setdefault(os.path.basename(p)[4:-3], p)  # "default" is always the in-tree library
    if a.variants:
        libs = {k: libs[k] for k in a.variants}
    t0 = time.time()
    sc = workload(a.workload)
    print(f"# workload {a.workload}: {sc.n_faces} faces, {sc.width}x{sc.height}, built in {time.time() - t0:.1f}s", flush=True)
    ref = None
    if not a.no_parity:
        from oracle import pyoracle as po  # checker only
        sm = smoke_scene()
        ref = po.oracle_render(sm, threads=os.cpu_count() or 4)[:2]
    out_dir = os.environ.get("AXR_AB_OUT") or os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    out = open(os.path.join(out_dir, f"ab_{a.workload}.jsonl"), "a")
    for name, path in libs.items():
        api._lib = None
        api.LIB_PATH = path
        rec = {"variant": name, "workload": a.workload}
        try:
            if ref is not None:
                c1, d1, _ = api.render_scene(sm, device=0)
                m = po.compare(c1, d1, ref[0], ref[1])
                rec["parity"] = {k: m[k] for k in ("coverage_mismatch", "depth_bit_mismatch", "color_max_diff")}
            dev = api.Device(sc.width, sc.height, sampler=sc.sampler)
            mesh = dev.load_scene(sc)
            for _ in range(a.warmup):
                dev.clear(); dev.draw_mesh(mesh, sc.model)
            dev.sync()
            dev.set_profiling(True)
            t = time.perf_counter()
            for _ in range(a.steps):
                dev.clear(); dev.draw_mesh(mesh, sc.model)
            dev.sync()
            rec["wall_ms_per_step"] = round((time.perf_counter() - t) / a.steps * 1e3, 4)
            ms, draws = dev.kernel_times()
            rec["kernel_us"] = {k: round(v / draws * 1e3, 1) for k, v in ms.items()}
            rec["draw_us"] = round(sum(rec["kernel_us"].values()), 1)
            st = dev.stats()
            rec["stats"] = {k: st[k] for k in ("triangles", "small_triangles", "binned_triangles", "bin_refs")}
            if a.overlap:
                dev.set_profiling(False)
                dev.set_overlap(True)
                for _ in range(a.warmup):
                    dev.clear(); dev.draw_mesh(mesh, sc.model)
                dev.sync()
                t = time.perf_counter()
                for _ in range(a.steps * 3):
                    dev.clear(); dev.draw_mesh(mesh, sc.model)
                dev.sync()
                rec["overlap_ms_per_step"] = round((time.perf_counter() - t) / (a.steps * 3) * 1e3, 4)
            dev.close()
            if a.e2e:
                rec["e2e_ms"] = {}
                for knob in ("0", "1"):
                    os.environ["AXR_B200_HOST_DEPTH_ZEROCOPY"] = knob  # read by axr_create
                    rec["e2e_ms"]["zerocopy_depth" if knob == "1" else "upload_depth"] = round(e2e_ms(sc, a.steps), 4)
                os.environ.pop("AXR_B200_HOST_DEPTH_ZEROCOPY", None)
        except Exception as e:  # a variant that fails must not hide the others
            rec["error"] = repr(e)
        line = json.dumps(rec)
        print(line, flush=True)
        out.write(line + "\n")
        out.flush()


if __name__ == "__main__":
    main()

"""TEST INFRASTRUCTURE ONLY — stateful fuzzing of the C ABI (include/axr_b200.h) on the SIMT interpreter build: random SEQUENCES of
calls on one long-lived context — upload / free meshes and textures, (re)assign materials, switch shader / sampler / overlap mode,
clear to random colours and finite depths, upload a framebuffer, draw onto the device framebuffer, draw onto a host framebuffer,
resolve, read stats, plus deliberate misuse (stale handles, missing textures) that must come back as error codes. A model of the
framebuffer is advanced with the oracle after every draw; every resolve must match it bit for bit. What this exercises that single
renders do not: the two alternating draw slots, the pending-draw / redo bookkeeping, mesh and texture slot reuse, the material
table upload, state carried from one draw to the next.

usage: AXR_SIMT_TESTS_ONLY=1 AXR_B200_LIB=tests/simt/_build/libaxr_simt.so python tests/simt/fuzz_api.py [--seconds 60] [--seed 0]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from axiomr_b200 import api, scenes as S  # noqa: E402
import use_simt  # noqa: E402

use_simt.install()  # this driver only ever runs against the interpreter build (AXR_B200_LIB)
from oracle import pyoracle as po  # noqa: E402
from fuzz import random_scene  # noqa: E402


class Model:
    """What the device framebuffer must contain."""

    def __init__(self, w, h):
        self.c = np.zeros((h, w, 4), dtype=np.uint8)  # Framebuffer constructor: colour 0, depth +inf (axr_create does the same)
        self.d = np.full((h, w), np.inf, dtype=np.float32)


def main():
    # bit-identical colours need the individually rounded colour arithmetic (the default fused mode is within 1 LSB by contract)
    os.environ.setdefault("AXR_B200_COLOR_MATH", "exact")
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60.0)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    t_end = time.time() + a.seconds
    n_ctx = n_ops = n_draws = n_checks = n_errors = 0
    while time.time() < t_end:
        W, H = int(rng.choice([64, 97, 160])), int(rng.choice([48, 63, 100]))
        dev = api.Device(W, H, sampler=int(rng.integers(0, 2)))
        n_ctx += 1
        fb = Model(W, H)
        fb_valid = True                      # False after a host draw: the device framebuffer is unspecified until cleared / uploaded
        host = Model(W, H)
        host_c = np.ascontiguousarray(host.c.copy())
        host_d = np.ascontiguousarray(host.d.copy())
        meshes = {}                          # handle -> [Scene per material group] (as uploaded, with their textures)
        sampler = dev_sampler = None
        for _ in range(int(rng.integers(5, 60))):
            if time.time() > t_end:
                break
            op = int(rng.integers(0, 12))
            n_ops += 1
            if op <= 1 or not meshes:        # upload a mesh with its textures and material
                sc = random_scene(rng, n_ops)
                sc.width, sc.height = W, H
                vp, cam = S.default_camera(W, H, eye=(float(rng.uniform(-1, 1)), float(rng.uniform(-1, 1)), float(rng.uniform(1, 6))))
                sc.view_proj, sc.cam_pos = vp, cam
                if rng.random() < 0.35 and sc.n_faces >= 3:  # several material groups: contiguous face ranges, own textures and Ns each
                    cuts = sorted(set(int(c) for c in rng.integers(1, sc.n_faces, int(rng.integers(1, 3)))))
                    bounds = [0] + cuts + [sc.n_faces]
                    parts = []
                    for g in range(len(bounds) - 1):
                        tsz = int(rng.choice([1, 4, 16]))
                        tex = ([S.cutout_texture(max(tsz, 2), 2), None, None, None, None] if sc.shader == S.SHADER_CUTOUT else
                               S._pbr_textures(tsz) if sc.shader == S.SHADER_PBR else S._phong_textures(tsz) if sc.shader == S.SHADER_PHONG else [None] * 5)
                        parts.append(S.Scene(f"{sc.name}_g{g}", W, H, sc.vertices, sc.indices[bounds[g]:bounds[g + 1]], sc.shader, sc.sampler,
                                             view_proj=vp, cam_pos=cam, light_dir=sc.light_dir, light_color=sc.light_color,
                                             specular_exponent=float(rng.choice([0.1, 0.2, 0.7])), textures=tex))
                    h = dev.upload_mesh(sc.vertices, sc.indices, groups=[(bounds[g], bounds[g + 1] - bounds[g]) for g in range(len(parts))])
                    for g, part in enumerate(parts):
                        th = [dev.upload_texture(t) if t is not None else api.NO_TEXTURE for t in part.textures]
                        dev.set_material(h, g, *th, specular_exponent=part.specular_exponent)
                    meshes[h] = parts
                else:
                    meshes[dev.load_scene(sc)] = [sc]
            elif op == 2 and len(meshes) > 1:  # free one; its handle must be refused afterwards
                h = list(meshes)[int(rng.integers(0, len(meshes)))]
                dev.free_mesh(h)
                del meshes[h]
                try:
                    dev.draw_mesh(h, np.eye(4, dtype=np.float32))
                    raise SystemExit(f"draw with a freed mesh handle {h} was accepted")
                except api.AxrError:
                    n_errors += 1
            elif op == 3:                    # clear to a random colour and, sometimes, a finite depth
                col = int(rng.integers(0, 1 << 32))
                z = float(rng.choice([np.inf, np.inf, 0.97, 0.5]))
                dev.clear(col, z)
                fb.c[:] = np.array([col & 255, (col >> 8) & 255, (col >> 16) & 255, (col >> 24) & 255], dtype=np.uint8)
                fb.d[:] = z
                fb_valid = True
            elif op == 4:                    # upload a framebuffer (the host model's contents)
                dev.upload_framebuffer(host.c, host.d)
                fb.c[:], fb.d[:] = host.c, host.d
                fb_valid = True
            elif op == 5:
                dev.set_overlap(bool(rng.integers(0, 2)))
            elif op == 6:                    # stats of the last draw are readable at any time
                dev.stats()
            elif op in (7, 8, 9) and fb_valid:  # draw onto the device framebuffer
                h = list(meshes)[int(rng.integers(0, len(meshes)))]
                parts = meshes[h]
                sc = parts[0]
                model = S._f32(S.mat_mul(S.rotate_y(float(rng.uniform(0, 6.3))), S.translate(*(rng.uniform(-1, 1, 3)))))
                smp = int(rng.integers(0, 2))
                dev.set_sampler(smp)
                dev.set_uniforms(sc.view_proj, sc.cam_pos)
                dev.set_shader(sc.shader, sc.light_dir, sc.light_color)
                dev.draw_mesh(h, model)
                for part in parts:  # groups in face order: the same per-pixel winner as the single draw
                    sc2 = S.Scene(part.name, W, H, part.vertices, part.indices, part.shader, smp, model=model, view_proj=part.view_proj,
                                  cam_pos=part.cam_pos, light_dir=part.light_dir, light_color=part.light_color,
                                  specular_exponent=part.specular_exponent, textures=part.textures)
                    fb.c, fb.d, _ = po.oracle_render(sc2, threads=2, color=fb.c, depth=fb.d)
                n_draws += 1
            elif op == 10:                   # the reference's calling convention: composite onto a host framebuffer, complete on return
                h = list(meshes)[int(rng.integers(0, len(meshes)))]
                parts = meshes[h]
                sc = parts[0]
                model = S._f32(S.rotate_y(float(rng.uniform(0, 6.3))))
                smp = int(rng.integers(0, 2))
                dev.set_sampler(smp)
                dev.set_uniforms(sc.view_proj, sc.cam_pos)
                dev.set_shader(sc.shader, sc.light_dir, sc.light_color)
                dev.draw_mesh_host(h, model, host_c, host_d)
                for part in parts:
                    sc2 = S.Scene(part.name, W, H, part.vertices, part.indices, part.shader, smp, model=model, view_proj=part.view_proj,
                                  cam_pos=part.cam_pos, light_dir=part.light_dir, light_color=part.light_color,
                                  specular_exponent=part.specular_exponent, textures=part.textures)
                    host.c, host.d, _ = po.oracle_render(sc2, threads=2, color=host.c, depth=host.d)
                m = po.compare(host_c, host_d, host.c, host.d)
                if m["coverage_mismatch"] or m["depth_bit_mismatch"] or m["color_max_diff"] or not np.array_equal(host_c, host.c):
                    raise SystemExit(f"MISMATCH (host draw) seed={a.seed} ctx #{n_ctx} op #{n_ops}: {m}")
                fb_valid = False
                n_draws += 1
                n_checks += 1
            elif fb_valid:                   # resolve and compare with the model
                c, d = dev.resolve()
                m = po.compare(c, d, fb.c, fb.d)
                if m["coverage_mismatch"] or m["depth_bit_mismatch"] or m["color_max_diff"] or not np.array_equal(c, fb.c):
                    raise SystemExit(f"MISMATCH seed={a.seed} ctx #{n_ctx} op #{n_ops}: {m}")
                n_checks += 1
        if fb_valid:
            c, d = dev.resolve()
            if not (np.array_equal(c, fb.c) and np.array_equal(d.view(np.uint32), fb.d.view(np.uint32))):
                raise SystemExit(f"MISMATCH at context end seed={a.seed} ctx #{n_ctx}")
            n_checks += 1
        dev.close()
    print(f"FUZZ OK seed={a.seed}: {n_ctx} contexts, {n_ops} API operations, {n_draws} draws, {n_checks} framebuffer checks, "
          f"{n_errors} refused misuses, all bit-identical to the oracle model", flush=True)


if __name__ == "__main__":
    main()

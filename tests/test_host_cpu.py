"""CPU: host-side logic, the C-ABI library's exports, multi-GPU sharding helpers (gloo, world_size 2)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from axiomr_b200 import api, multi, scenes as S  # noqa: E402


def test_abi_library_exports_every_declared_symbol():
    from axiomr_b200 import build
    build.build()
    hdr = open(os.path.join(ROOT, "include", "axr_b200.h")).read()
    declared = set(re.findall(r"\b(axr_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(os.path.join(ROOT, "axiomr_b200", "libaxr_b200.so"))
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert declared == set(api.ABI_SYMBOLS), declared ^ set(api.ABI_SYMBOLS)
    lib.axr_abi_version.restype = ctypes.c_int
    assert lib.axr_abi_version() == 1


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(api.AxrError) as e:
        api.Device(64, 64)
    assert e.value.code == -3 and "no CPU path" in str(e.value)


def test_product_code_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "axiomr_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in src and "axr_oracle" not in src and "libaxr_ref" not in src, os.path.join(dirpath, f)


def test_band_rows_partition():
    assert multi.band_rows(4320, 8) == [(0, 544), (544, 1088), (1088, 1632), (1632, 2176), (2176, 2720), (2720, 3264), (3264, 3792), (3792, 4320)]
    for h in (16, 100, 1080, 2160, 4320, 37):
        for w in (1, 2, 3, 8):
            b = multi.band_rows(h, w)
            assert b[0][0] == 0 and b[-1][1] == h and all(a[1] == c[0] for a, c in zip(b, b[1:]))
            assert all(y0 % 16 == 0 and y1 > y0 for y0, y1 in b)
    assert sum(len(multi.views_for_rank(r, 8, 64)) for r in range(8)) == 64


def test_scene_generators_shapes():
    v, f = S.icosphere(3)
    assert f.shape[0] == 20 * 4 ** 3 and v.shape == (10 * 4 ** 3 + 2, 14)
    v, f = S.torus(12, 9)
    assert f.shape[0] == 2 * 12 * 9 and v.shape[0] == 13 * 10
    v, f = S.head_like()
    assert f.shape[0] == 2450
    assert S.diffuse_texture(16).shape == (16, 16, 4) and S.normal_texture(16).dtype == np.uint8
    # BASELINE.md §4 triangle / vertex counts of the named configs (computed, not generated)
    assert 20 * 4 ** 8 == 1310720 and 10 * 4 ** 8 + 2 == 655362 and 2 * 2236 ** 2 == 9999392 and 2237 ** 2 == 5004169


def test_framebuffer_mirror_semantics():
    fb = api.Framebuffer(8, 4, True, pinned=False)
    assert fb.getColorData().shape == (4, 8, 4) and np.isinf(fb.getDepthData()).all()
    fb.clearColor(api.Color(10, 20, 30, 255))
    assert fb.getColorData()[0, 0].tolist() == [30, 20, 10, 255]  # B,G,R,A (reference src/framebuffer.cpp:29-32)
    fb.clearDepth(0.5)
    assert (fb.getDepthData() == 0.5).all()
    p = api.TiledPipeline(4, None, fb)
    p.drawMesh(np.eye(4), None)  # silent return without shader/camera (reference src/tiled_pipeline.cpp:146)
    m = api.Mesh(np.zeros((3, 14), np.float32), np.array([[0, 1, 2]], np.uint32))
    with pytest.raises(KeyError):
        m.getMaterial("nope")  # unordered_map::at


_GLOO = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
from axiomr_b200 import multi
rank, world = int(sys.argv[1]), 2
os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=sys.argv[2], RANK=str(rank), WORLD_SIZE="2")
dist.init_process_group("gloo", rank=rank, world_size=world)
H, W = 80, 24
bands = multi.band_rows(H, world)
full_c = torch.arange(H * W, dtype=torch.int32).reshape(H, W)
full_d = torch.arange(H * W, dtype=torch.float32).reshape(H, W) * 0.5
color = torch.zeros((H, W), dtype=torch.int32); depth = torch.full((H, W), float("inf"))
y0, y1 = bands[rank]
color[y0:y1] = full_c[y0:y1]; depth[y0:y1] = full_d[y0:y1]      # what this rank "rendered"
multi.gather_bands(color, depth, bands, rank, dist)
if rank == 0:
    assert torch.equal(color, full_c) and torch.equal(depth, full_d)
slots = multi.gather_views(color, depth, rank, world, dist)
if rank == 0:
    assert slots[0].shape == (1, H, W) and int(slots[0][0, bands[1][0], 0]) == int(full_c[bands[1][0], 0])
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
'''


def test_compositor_gather_logic_gloo_world2(tmp_path):
    script = tmp_path / "gloo_rank.py"
    script.write_text(_GLOO % ROOT)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), str(r), port], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs), outs


def test_obj_loader_matches_reference_loader_golden(tmp_path):
    """axiomr_b200/obj.py (mirror of AR::Mesh(path), reference src/mesh.cpp) against arrays produced by the reference's own
    loader (tests/golden/obj_loader.npz): de-duplication order, fan triangulation, tangent / bitangent generation, bit for bit."""
    from axiomr_b200 import obj
    z = np.load(os.path.join(ROOT, "tests", "golden", "obj_loader.npz"))
    for name in ("head", "quad", "poly"):
        p = tmp_path / f"{name}.obj"
        p.write_bytes(z[name + "_obj"].tobytes())
        (tmp_path / f"{name}.mtl").write_bytes(z[name + "_mtl"].tobytes())
        m = obj.load_obj(str(p))
        assert np.array_equal(m.getFaces(), z[name + "_faces"]), name
        got, want = m.getVertices(), z[name + "_vertices"]
        assert got.shape == want.shape, name
        assert np.array_equal(np.isnan(got), np.isnan(want)), name
        ok = (got.view(np.uint32) == want.view(np.uint32)) | np.isnan(want)
        assert ok.all(), (name, np.argwhere(~ok)[:5])
        assert m.getMaterialGroups()[0].materialName == "m0" and m.getMaterialGroups()[0].faceCount == want.shape[0] * 0 + m.getFaces().shape[0]
    assert abs(m.getMaterial("m0").specularExponent - 0.25) < 1e-7


def test_mtl_parse_behind_the_c_abi(tmp_path):
    """axr_parse_mtl (reference src/mesh.cpp:65-220): names, Ns, the five texture slots with their aliases, trimmed paths with
    blanks inside, statements before the first newmtl ignored — and the same answers as the Python mirror on the golden MTL."""
    from axiomr_b200 import api
    text = (b"Ns 9\nmap_Kd ignored.png\n# comment\nnewmtl first\nNs 0.25\nKd 1 0 0\nmap_Kd   tex/my diffuse.png \t\r\n"
            b"bump\tb.png\nrefl m.png\nmap_Ns r.png\nmap_A0 a.png\n\nnewmtl second\nnorm n2.png\nmap_Ks k2.png\nNs 1e1\n")
    got = api.parse_mtl(text)
    assert [g[0] for g in got] == ["first", "second"]
    assert got[0][1] == np.float32(0.25) and got[1][1] == 10.0
    assert got[0][2] == {"diffuse": "tex/my diffuse.png \t\r".rstrip(" \t"), "bump": "b.png", "metallic": "m.png", "roughness": "r.png", "ao": "a.png"}
    assert got[1][2] == {"bump": "n2.png", "metallic": "k2.png"}
    z = np.load(os.path.join(ROOT, "tests", "golden", "obj_loader.npz"))
    got = api.parse_mtl(z["poly_mtl"].tobytes())
    assert got[0][0] == "m0" and abs(got[0][1] - 0.25) < 1e-7


def test_bench_reference_arm_line_and_no_cpu_fallback():
    """bench.py --impl reference prints ONE JSON line with the contract's keys (tiny workload, runs on the CPU);
    the B200 arm refuses to run without a GPU instead of falling back."""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["vs_baseline"] is None and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "tiny", "--steps", "1"], capture_output=True, text=True, timeout=300)
        assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_present_writes_bottom_up_bgra_as_png(tmp_path):
    """Window::present semantics (reference src/window.cpp:70-84 + the bottom-up DIB of src/windows_bitmap.cpp:19-23): framebuffer
    row 0 is the bottom row of the picture, bytes are B,G,R,A. Decoded independently with PIL."""
    Image = pytest.importorskip("PIL.Image")
    fb = api.Framebuffer(5, 3, True, pinned=False)
    fb.clearColor(api.Color(10, 20, 30, 255))
    c = fb.getColorData()
    c[0, 0] = (1, 2, 3, 255)      # B,G,R,A at the bottom-left corner
    c[2, 4] = (200, 100, 50, 128)  # top-right corner
    path = str(tmp_path / "frame.png")
    fb.present(path)
    img = np.asarray(Image.open(path))
    assert img.shape == (3, 5, 4)
    assert tuple(img[2, 0]) == (3, 2, 1, 255)       # bottom-left, R,G,B,A
    assert tuple(img[0, 4]) == (50, 100, 200, 128)  # top-right
    assert tuple(img[1, 2]) == (10, 20, 30, 255)    # clear colour: packed (a<<24)|(r<<16)|(g<<8)|b
    with pytest.raises(ValueError):
        api.present(np.zeros((4, 4), dtype=np.uint8), path)

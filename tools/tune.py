"""Tuning harness: time several builds of libaxr_b200.so (variants_tmp/lib_*.so) on one workload in one process.
usage: python tools/tune.py [workload] [variant ...]   (run under gpurun)"""
import glob
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from axiomr_b200 import api  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
names = sys.argv[2:]
libs = {"default": os.path.join(ROOT, "axiomr_b200", "libaxr_b200.so")}
for p in sorted(glob.glob(os.path.join(ROOT, "variants_tmp", "lib_*.so"))):
    libs[os.path.basename(p)[4:-3]] = p
if names:
    libs = {k: v for k, v in libs.items() if k in names}
t0 = time.time()
sc = bench.build_workload(wl)
print(f"scene {wl} built in {time.time() - t0:.1f}s", flush=True)
for name, path in libs.items():
    api._lib = None
    api.LIB_PATH = path
    dev = api.Device(sc.width, sc.height, sampler=sc.sampler)
    mesh = dev.load_scene(sc)
    for _ in range(3):
        dev.clear(); dev.draw_mesh(mesh, sc.model)
    dev.sync()
    dev.set_profiling(True)
    t = time.perf_counter()
    n = 10
    for _ in range(n):
        dev.clear(); dev.draw_mesh(mesh, sc.model)
    dev.sync()
    wall = (time.perf_counter() - t) / n * 1e3
    ms, draws = dev.kernel_times()
    k = {a: round(b / draws * 1e3) for a, b in ms.items()}
    st = dev.stats()
    print(f"{name:10s} wall {wall:.3f} ms  kernels(us) {k} sum {sum(k.values())}  tris {st['triangles']} small {st['small_triangles']} binned {st['binned_triangles']} refs {st['bin_refs']}", flush=True)
    dev.close()

"""Break the host-framebuffer drawMesh call (axr_draw_mesh_host) into its parts on one workload. usage: python tools/e2e_diag.py [workload]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from axiomr_b200 import api  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
sc = bench.build_workload(wl)
dev = api.Device(sc.width, sc.height, sampler=sc.sampler)
mesh = dev.load_scene(sc)
fb = api.Framebuffer(sc.width, sc.height, True)


def timed(fn, n=8):
    fn()
    dev.sync()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    dev.sync()
    return (time.perf_counter() - t) / n * 1e3


def clear_host():
    fb.clearColor(api.Color(0, 0, 0, 255))
    fb.clearDepth()


def dev_draw():
    dev.clear()
    dev.draw_mesh(mesh, sc.model)


def host_draw():
    dev.draw_mesh_host(mesh, sc.model, fb.getColorData(), fb.getDepthData())


print("device clear + draw            %.3f ms" % timed(dev_draw))
print("upload depth only (33 MB)      %.3f ms" % timed(lambda: dev.upload_framebuffer(None, fb.getDepthData())))
print("resolve colour+depth (66 MB)   %.3f ms" % timed(lambda: dev.resolve(fb.getColorData(), fb.getDepthData())))
clear_host()
t_clear = timed(clear_host, 3)
print("host clear (numpy)             %.3f ms" % t_clear)
for ov in (False, True):
    dev.set_overlap(ov)
    clear_host()
    print("draw_mesh_host, no clear between (nothing passes the depth test after the 1st), overlap=%d  %.3f ms" % (ov, timed(host_draw)))

    def both():
        clear_host()
        host_draw()
    print("host clear + draw_mesh_host, overlap=%d  %.3f ms  (minus clear: %.3f)" % (ov, timed(both), timed(both) - t_clear))
dev.close()

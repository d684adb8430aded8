"""Run under torchrun (one rank per GPU): checks that the multi-GPU composites on GPU 0 equal single-GPU renders.
   bands (peer stores and NCCL gather): the composited frame == the full frame rendered by one GPU, bit for bit;
   views: slot r-1 on GPU 0 == rank r's view rendered by GPU 0 itself."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from axiomr_b200 import api, multi, scenes as S  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
v, f = S.torus(300, 300)
W, H = 1280, 720
base = S.Scene("t", W, H, v, f, S.SHADER_PHONG, model=S._f32(S.rotate_y(0.5)), textures=S._phong_textures(256))


def full_frame(scene):
    c, d, _ = api.render_scene(scene, device=local)
    return c, d


ok = True
# ---- bands
for transport in ("peer", "nccl"):
    band = multi.band_rows(H, world, multi.band_granule(transport))[rank]
    dev = api.Device(W, H, device=local, band=band)
    mesh = dev.load_scene(base)
    stream = torch.cuda.ExternalStream(dev.stream, device=torch.device("cuda", local))
    comp = multi.Compositor(dev, rank, world, "bands", band, stream, transport=transport)
    for _ in range(5):   # several frames: exercises the (dirty-tile) clears + redraw into the shared targets, both sets
        comp.begin_step()
        if comp.clears_own_target:
            dev.clear()
        dev.draw_mesh(mesh, base.model)
        comp.composite()
    comp.finish()
    dev.sync()
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        color, depth = comp.frame(comp.last_set()) if transport == "peer" else multi.framebuffer_tensors(dev)
        c = color.contiguous().cpu().numpy().view(np.uint8).reshape(H, W, 4)
        d = depth.contiguous().cpu().numpy()
        c0, d0 = full_frame(base)
        same = np.array_equal(c, c0) and np.array_equal(d.view(np.uint32), d0.view(np.uint32))
        print(f"bands/{transport}: composite == single-GPU frame: {same}", flush=True)
        ok &= same
    dist.barrier()
    comp.release()
    dev.close()
# ---- views
for transport in ("peer", "nccl"):
    sc = S.Scene("t", W, H, v, f, S.SHADER_PHONG, model=base.model, textures=base.textures)
    sc.view_proj, sc.cam_pos = S.view_matrix_for(rank, 8, W, H)
    dev = api.Device(W, H, device=local)
    mesh = dev.load_scene(sc)
    stream = torch.cuda.ExternalStream(dev.stream, device=torch.device("cuda", local))
    comp = multi.Compositor(dev, rank, world, "views", None, stream, transport=transport)
    for j in (3, 1, 2, 5, 0):   # the camera moves from frame to frame (stale tiles in the shared slots); the last frame is view `rank`
        dev.set_uniforms(*S.view_matrix_for((rank + j) % 8, 8, W, H))
        comp.begin_step()
        if comp.clears_own_target:
            dev.clear()
        dev.draw_mesh(mesh, sc.model)
        comp.composite()
    comp.finish()
    dev.sync()
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        for r in range(1, world):
            other = S.Scene("t", W, H, v, f, S.SHADER_PHONG, model=base.model, textures=base.textures)
            other.view_proj, other.cam_pos = S.view_matrix_for(r, 8, W, H)
            c0, d0 = full_frame(other)
            # the last frame went to set last_set() (the other set holds an earlier frame, taken with another camera, or is being tidied)
            for b in (comp.last_set(),):
                cs, ds = (comp.view_slot(b, r) if transport == "peer" else (comp.slots[b][0][r - 1], comp.slots[b][1][r - 1]))
                c = cs.contiguous().cpu().numpy().view(np.uint8).reshape(H, W, 4)
                d = ds.contiguous().cpu().numpy()
                same = np.array_equal(c, c0) and np.array_equal(d.view(np.uint32), d0.view(np.uint32))
                print(f"views/{transport}: slot set {b} view {r} == single-GPU render: {same}", flush=True)
                ok &= same
    dist.barrier()
    comp.release()
    dev.close()
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, 0)
dist.destroy_process_group()
print("MULTI_GPU_CHECK", "OK" if int(flag.item()) else "FAILED", flush=True)
sys.exit(0 if int(flag.item()) else 1)

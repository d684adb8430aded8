#!/usr/bin/env python
"""bench.py — throughput of the B200 tiled-raster path on BASELINE.json's headline workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c3n|c2|c4|c1|c5] [--mode views|bands]
    python bench.py --impl reference ...      # the reference's own CPU path on the box's host cores

A step is one frame: Framebuffer clear + TiledPipeline::drawMesh of the whole mesh (vertex stage, clip/cull/setup,
binning, per-tile raster + shading, resolve), with mesh, textures and framebuffer already resident in HBM.
Prints ONE JSON line (rank 0). See DESIGN.md §Measurement for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from axiomr_b200 import scenes as S  # noqa: E402

FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback
# one string for both arms: the driver divides the two lines only when metric, unit and config agree
METRIC = "Mtri/s (input faces per second through TiledPipeline::drawMesh)"


def build_workload(name: str) -> S.Scene:
    if name == "c3":      # BASELINE.json configs[2]: the configuration the headline metric is quoted on
        return S.config3(sampler=S.SAMPLER_BILINEAR)
    if name == "c3n":     # same scene, nearest sampling = the reference's own Texture::sample (parity mode)
        return S.config3(sampler=S.SAMPLER_NEAREST)
    if name == "c2":
        return S.config2()
    if name == "c4":
        return S.config4()
    if name == "c1":
        return S.config1()
    if name == "c5":      # C3 scene at 8K (bands mode)
        return S.config3(w=7680, h=4320, sampler=S.SAMPLER_BILINEAR)
    if name == "mid":     # mid-size triangles (box area ~100-400 px): icosphere level 6 at 4K, Phong
        v, f = S.icosphere(6)
        return S.Scene("mid_icosphere6_phong_4k", 3840, 2160, v, f, S.SHADER_PHONG, S.SAMPLER_BILINEAR, model=S._f32(S.rotate_y(0.5)),
                       textures=S._phong_textures(2048))
    if name == "tiny":    # CI-sized stand-in
        return S.config3(n=300, w=1280, h=720, tex=512, sampler=S.SAMPLER_BILINEAR)
    raise SystemExit(f"unknown workload {name}")


def workload_label(name: str, sc: S.Scene) -> str:
    smp = "bilinear" if sc.sampler else "nearest"
    shader = ["FlatShader", "PhongShader", "PBRShader"][sc.shader]
    tex = f", {sc.textures[0].shape[0]}^2 diffuse+normal textures ({smp})" if sc.textures[0] is not None else ", no textures"
    return f"{name}: {sc.n_faces} tris / {sc.n_verts} verts, {sc.width}x{sc.height}, {shader}{tex}"


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(sc: S.Scene, covered_px: int):
    """B_alg = 12 T + 56 V + 12 W H + sum_tex min(4 taps P_cov, 4 Wt Ht), split by the stage that owns each term."""
    taps = 4 if sc.sampler else 1
    tex = 0
    n_tex = {0: 0, 1: 2, 2: 5}[sc.shader]
    for t in sc.textures[:n_tex] if n_tex else []:
        if t is not None:
            tex += min(4 * taps * covered_px, 4 * t.shape[0] * t.shape[1])
    per_stage = {
        "vertex_xform": 12 * sc.n_verts,                                        # positions
        "setup_raster": 12 * sc.n_faces,                                        # 3 u32 indices per face
        "tile_shade": 44 * sc.n_verts + 12 * sc.width * sc.height + tex,        # attributes + depth r/w + colour w + texels
        "scan_tiles": 0, "bin_scatter": 0,
    }
    return sum(per_stage.values()), per_stage


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------- reference / CPU arm
def cpu_reference_run(sc: S.Scene, seconds_budget: float, threads: int | None = None, steps: int = 1, warmup: int = 0,
                      thread_sweep: bool = True):
    """Times the reference's CPU implementation of the path (oracle/_ref = the unmodified sources; else the C port) on a
    bounded, contiguous face sample of the workload. Returns (Mtri/s, info dict). bench.py is allowed to run oracle/ here only."""
    from oracle import pyoracle as po
    use_ref = po.ref_available()
    cores = os.cpu_count() or 1
    if threads is None:
        threads = po.ref_hardware_concurrency() if po.ref_available() else cores
    # the reference only has nearest sampling; its arm always runs Texture::sample as shipped
    ref_scene = sc
    if use_ref and sc.sampler != S.SAMPLER_NEAREST:
        import copy
        ref_scene = copy.copy(sc)
        ref_scene.sampler = S.SAMPLER_NEAREST

    def run(n_faces):
        if use_ref:
            _, _, secs = po.ref_render(ref_scene, threads=threads, chunk=30000, n_faces=n_faces)
        else:
            po.build(ref=False)
            _, _, secs = po.oracle_render(sc, threads=threads, n_faces=n_faces)
        return secs

    # calibrate on a small prefix (also picks the thread count the reference runs fastest with on this box:
    # its own default is hardware_concurrency, reference src/renderer.cpp:71), then size the sample for the budget
    n_cal = min(sc.n_faces, 200_000)
    if thread_sweep and use_ref:
        best = None
        for th in sorted({threads, 64, 32, 16, 8}):
            if th > max(threads, 1):
                continue
            threads_try = th
            t_try = None
            try:
                _, _, t_try = po.ref_render(ref_scene, threads=threads_try, chunk=30000, n_faces=n_cal)
            except Exception:
                continue
            if best is None or t_try < best[1]:
                best = (th, t_try)
        if best:
            threads = best[0]
    t_cal = max(run(n_cal), 1e-6)
    rate = n_cal / t_cal
    n = int(min(sc.n_faces, max(n_cal, rate * seconds_budget)))
    for _ in range(warmup):
        run(n)
    times = [run(n) for _ in range(max(1, steps))]
    t = statistics.median(times)
    info = {"kind": "reference" if use_ref else "port", "cores": threads, "host_cpus": cores,
            "sample": f"first {n} of {sc.n_faces} faces of the workload drawn in face order in chunks of 30000 "
                      f"(TiledPipeline arena limit), full {sc.width}x{sc.height} frame, median of {len(times)} run(s), "
                      f"{t * 1e3:.1f} ms each" + ("; nearest sampling (the reference has no bilinear mode)" if use_ref and sc.sampler else ""),
            "ms_per_sample": t * 1e3, "faces": n}
    return n / t / 1e6, info


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sc = build_workload(args.workload)
    budget = 8.0
    v, info = cpu_reference_run(sc, budget, steps=args.steps, warmup=min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "Mtri/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": info["ms_per_sample"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_label(args.workload, sc)},
        "cpu_baseline": {"value": v, "unit": "Mtri/s", "cores": info["cores"], "kind": info["kind"], "sample": info["sample"]},
        "e2e": {"value": v, "unit": "Mtri/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "frames_per_s_extrapolated": v * 1e6 / sc.n_faces,
    }
    _emit(line)


# ----------------------------------------------------------------------------------------------- B200 arm
_REAL_STDOUT = None


def _quiet_stdout():
    """Libraries (NCCL's version banner, torchrun notices) write to fd 1; the contract is ONE JSON line on stdout.
    Everything else goes to stderr, the JSON line is written to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def _time_steps(torch, stream, barrier, step, n):
    """n steps between two CUDA events on the context stream, barrier + synchronize on both sides. Returns ms per step."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for i in range(n):
        step(i)
    e1.record(stream)
    barrier()
    return e0.elapsed_time(e1) / n


def _max_over_ranks(torch, dist, x: float) -> float:
    if not dist:
        return x
    t = torch.tensor([x], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _parity_counts(torch, c_a, d_a, c_b, d_b):
    """Bit comparison of two framebuffers on the device: (coverage mismatches, depth words that differ, max colour byte difference)."""
    cov = int((torch.isfinite(d_a) != torch.isfinite(d_b)).sum().item())
    dbits = int((d_a.contiguous().view(torch.int32) != d_b.contiguous().view(torch.int32)).sum().item())
    ba = c_a.contiguous().view(torch.uint8).to(torch.int16)
    bb = c_b.contiguous().view(torch.uint8).to(torch.int16)
    cmax = int((ba - bb).abs().max().item()) if ba.numel() else 0
    return cov, dbits, cmax


def _run_multi(torch, dist, api, multi, sc, rank, world, local, mode, transport, steps, warmup, views_total=None, check=True, fill=True, rows=True):
    """One multi-GPU (or, at world == 1, single-GPU) figure: K timed frames of `sc` sharded by `mode`, then — on GPU 0, outside the
    timed region — a bit comparison of what landed there with GPU 0's own single-GPU render of the same view / frame.
    views_total: every rank renders its share of that many camera views per step (BASELINE config 5's 64 views) instead of one."""
    W, H = sc.width, sc.height
    band = multi.band_rows(H, world, multi.band_granule(transport))[rank] if (world > 1 and mode == "bands") else None
    dev = api.Device(W, H, device=local, sampler=sc.sampler, band=band)
    mesh = dev.load_scene(sc)
    dev.set_overlap(True)
    stream = torch.cuda.ExternalStream(dev.stream, device=torch.device("cuda", local))
    comp = multi.Compositor(dev, rank, world, mode, band, stream, transport=transport, fill=fill, rows=rows) if world > 1 else None
    n_cams = max(world, 8) if views_total is None else views_total
    my_views = [rank] if views_total is None else multi.views_for_rank(rank, world, views_total)
    if mode == "bands":
        my_views = [0]
    cams = {v: S.view_matrix_for(v, n_cams, W, H) for v in my_views} if mode == "views" else {0: (sc.view_proj, sc.cam_pos)}

    def frame(v):
        if len(cams) > 1 or mode == "views":
            dev.set_uniforms(*cams[v])
        if comp:
            comp.begin_step()
        if not comp or comp.clears_own_target:
            dev.clear(0xFF000000, float("inf"))
        dev.draw_mesh(mesh, sc.model)
        if comp:
            comp.composite()

    def step(_i):
        for v in my_views:
            frame(v)

    def barrier():
        if comp:
            comp.finish()
        dev.sync()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
            torch.cuda.synchronize()

    for i in range(warmup):
        step(i)
    ms = _max_over_ranks(torch, dist, _time_steps(torch, stream, barrier, step, steps))
    parity = None
    if check and comp and transport == "peer":
        # GPU 0 re-renders, on its own, what the other ranks stored into its memory during the last frame, and compares bit for bit
        cov = dbits = cmax = 0
        if rank == 0:
            b = comp.last_set()
            ref = api.Device(W, H, device=local, sampler=sc.sampler)
            rmesh = ref.load_scene(sc)
            rc, rd = multi.framebuffer_tensors(ref)
            targets = ([(r, comp.view_slot(b, r)) for r in range(1, world)] if mode == "views" else [(0, comp.frame(b))])
            for r, (tc, td) in targets:
                if mode == "views":
                    last_view = multi.views_for_rank(r, world, views_total)[-1] if views_total else r
                    ref.set_uniforms(*S.view_matrix_for(last_view, n_cams, W, H))
                ref.clear(0xFF000000, float("inf"))
                ref.draw_mesh(rmesh, sc.model)
                ref.sync()
                torch.cuda.synchronize()
                a, b_, c_ = _parity_counts(torch, tc, td, rc, rd)
                cov += a; dbits += b_; cmax = max(cmax, c_)
            ref.close()
        parity = {"coverage_mismatch": cov, "depth_bit_mismatch": dbits, "color_max_diff": cmax,
                  "compared": (f"{world - 1} view slot(s)" if mode == "views" else "the composited frame") +
                              " on GPU 0 against GPU 0's own single-GPU render, last frame, outside the timed region"}
    if dist:
        dist.barrier()
    if comp:
        comp.release()
    dev.close()
    frames = len(my_views) if mode == "views" else 1
    total_frames = (views_total if views_total else world) if mode == "views" else 1
    return {"ms_per_step": ms, "frames_per_step": total_frames, "frames_per_rank": frames,
            "value": sc.n_faces * total_frames / (ms * 1e-3) / 1e6, "unit": "Mtri/s", "frames_per_s": total_frames / (ms * 1e-3),
            "composite_parity": parity}


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--mode", default="views", choices=["views", "bands"], help="multi-GPU sharding (N>1)")
    ap.add_argument("--transport", default="peer", choices=["peer", "nccl"], help="composite to GPU 0: fused peer stores or NCCL gather")
    ap.add_argument("--no-fill", action="store_true", help="peer transport without axr_set_output_fill: covered pixels only, GPU 0 re-clears the flagged tiles")
    ap.add_argument("--no-rows", action="store_true", help="peer transport with 8x4-pixel blocks per warp on every rank (32-byte remote stores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the BASELINE config 5 figures (8K bands, 64 views) and the sustained pass")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    from axiomr_b200 import api, multi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the raster path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist = None

    sc = build_workload(args.workload)
    W, H, T = sc.width, sc.height, sc.n_faces
    band = None
    if world > 1 and args.mode == "bands":
        band = multi.band_rows(H, world, multi.band_granule(args.transport))[rank]
    elif world > 1:
        sc.view_proj, sc.cam_pos = S.view_matrix_for(rank, max(world, 8), W, H)  # each rank its own camera of the replicated scene
    dev = api.Device(W, H, device=local, sampler=sc.sampler, band=band)
    mesh = dev.load_scene(sc)
    stream = torch.cuda.ExternalStream(dev.stream, device=torch.device("cuda", local))
    comp = multi.Compositor(dev, rank, world, args.mode, band, stream, transport=args.transport, fill=not args.no_fill, rows=not args.no_rows) if world > 1 else None

    def step(_i=0):
        if comp:
            comp.begin_step()
        if not comp or comp.clears_own_target:
            dev.clear(0xFF000000, float("inf"))
        dev.draw_mesh(mesh, sc.model)
        if comp:
            comp.composite()

    def barrier():
        if comp:
            comp.finish()
        dev.sync()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- primary timed region: K steps with consecutive frames overlapped (axr_set_overlap(1): the geometry stages of frame i+1
    #      run on a second stream beside the tile kernel of frame i — the mode the library is meant to be driven in)
    dev.set_overlap(True)
    for _ in range(args.warmup):
        step()
    barrier()
    stats = dev.stats()
    launches_per_step = int(stats["kernel_launches"]) + (1 if (not comp or comp.clears_own_target) else 0) + (comp.launches_per_step if comp else 0)
    clocks = ClockSampler(local)
    clocks.start()
    ms_step = _max_over_ranks(torch, dist, _time_steps(torch, stream, barrier, step, args.steps))

    # ---- the same K steps serially (overlap off: every kernel of a frame on one stream, in launch order) with a CUDA event pair
    #      around every kernel: per-kernel durations for the roofline (measured live, on the stream the kernels are launched on)
    dev.set_overlap(False)
    for _ in range(2):
        step()
    dev.set_profiling(True)
    ms_serial = _max_over_ranks(torch, dist, _time_steps(torch, stream, barrier, step, args.steps))
    ktimes, kdraws = dev.kernel_times()
    dev.set_profiling(False)

    # ---- sustained pass: the primary loop again for >= 0.6 s, so that the clock / power samples below describe load, not idle
    dev.set_overlap(True)
    sustained = None
    if not args.no_extras:
        n_sus = int(min(20000, max(args.steps, 0.6e3 / max(ms_step, 1e-3))))
        for _ in range(2):
            step()
        ms_sus = _max_over_ranks(torch, dist, _time_steps(torch, stream, barrier, step, n_sus))
        sustained = {"steps": n_sus, "ms_per_step": ms_sus, "note": "same loop as the timed region, run long enough for nvidia-smi to sample it"}
    clk = clocks.stop()

    # ---- multi-GPU: what landed on GPU 0 during the last frame, bit-compared with GPU 0's own render (outside the timed region)
    composite_parity = None
    if comp and args.transport == "peer":
        cov = dbits = cmax = 0
        if rank == 0:
            b = comp.last_set()
            ref = api.Device(W, H, device=local, sampler=sc.sampler)
            rmesh = ref.load_scene(sc)
            rc, rd = multi.framebuffer_tensors(ref)
            targets = ([(r, comp.view_slot(b, r)) for r in range(1, world)] if args.mode == "views" else [(0, comp.frame(b))])
            for r, (tc, td) in targets:
                if args.mode == "views":
                    ref.set_uniforms(*S.view_matrix_for(r, max(world, 8), W, H))
                ref.clear(0xFF000000, float("inf"))
                ref.draw_mesh(rmesh, sc.model)
                ref.sync()
                torch.cuda.synchronize()
                a_, b_, c_ = _parity_counts(torch, tc, td, rc, rd)
                cov += a_; dbits += b_; cmax = max(cmax, c_)
            ref.close()
        composite_parity = {"coverage_mismatch": cov, "depth_bit_mismatch": dbits, "color_max_diff": cmax,
                            "compared": (f"{world - 1} view slot(s)" if args.mode == "views" else "the composited frame") +
                                        " on GPU 0 against GPU 0's own single-GPU render of the same view, last frame, outside the timed region"}
        if dist:
            dist.barrier()

    # ---- FP32 issue micro-benchmark: the measured issue rate turns the kernels' instruction counts into a floor
    fp32_rate, ffma_rate = dev.measure_fp32_issue()
    gather_rate = dev.measure_gather()   # randomly placed 32-byte DRAM sectors per second (16-byte gathers)

    # covered pixels (for the texture term of the algorithmic bytes) — outside the timed region
    if comp and args.transport == "peer" and args.mode == "bands":
        covered = 0
        if rank == 0:
            covered = int(torch.isfinite(comp.frame(comp.last_set())[1]).sum().item())
    else:
        if comp and args.transport == "peer" and rank != 0:
            covered = 0   # this rank's pixels live in GPU 0's slot; rank 0's own view is counted below
        else:
            _, depth = dev.resolve()
            y0, y1 = dev.band
            covered = int(np.isfinite(depth[y0:y1]).sum())

    units = T * (world if (world > 1 and args.mode == "views") else 1)  # faces processed per step by the whole job
    value = units / (ms_step * 1e-3) / 1e6

    line = None
    if rank == 0:
        peak, peak_src = hbm_peak()
        b_alg, per_stage = algorithmic_bytes(sc, covered)
        kavg = {k: (v / max(kdraws, 1)) for k, v in ktimes.items()}
        dom = max(kavg, key=lambda k: kavg[k])
        dom_ms = kavg[dom]
        ach = per_stage[dom] / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(args.workload, {}).get(dom)
            except Exception:
                traffic = None
        draw_ms = sum(kavg.values())
        line = {
            "metric": METRIC,
            "value": value, "unit": "Mtri/s", "frames_per_s": (units / T) / (ms_step * 1e-3),
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak" if args.mode == "views" else "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_label(args.workload, sc),
                       "step": "axr_clear + axr_draw_mesh, all inputs resident in HBM; consecutive frames overlapped (axr_set_overlap(1): vertex + "
                               "setup kernels of frame i+1 on a second stream beside the tile kernel of frame i). `serial` holds the same K steps "
                               "with everything on one stream, which is what kernel_ms / roofline time",
                       "l2": "working set (mesh %.0f MB + textures + 8 B/px keys + framebuffer) exceeds the 126 MB L2; no explicit flush"
                             % ((sc.vertices.nbytes + sc.indices.nbytes) / 1e6),
                       "color_math": "fast (fused multiply-adds + SFU approximations in the colour arithmetic only; coverage, depth and texel "
                                     "selection exact; 8-bit colour within 1 LSB of the reference, the tolerance BASELINE.json states)",
                       "parallelism": ("1 GPU" if world == 1 else f"{args.mode} x{world}: " +
                                       (("one camera view of the replicated scene per GPU; the tile kernel of every rank stores its covered pixels "
                                         "(colour + depth) straight into a per-view slot in GPU 0's memory through a CUDA-IPC peer mapping over "
                                         "NVLink and flags the tiles it touched; GPU 0 re-clears only those tiles; one 4-byte NCCL all-reduce per "
                                         "frame orders frames"
                                         if args.transport == "peer" else
                                         "one camera view of the replicated scene per GPU; every finished frame (colour + depth, 8 B/px) is gathered to "
                                         "GPU 0 with NCCL send/recv on a second stream while the next frame renders (double-buffered)")
                                        if args.mode == "views" else
                                        "16-px-aligned screen bands of one frame, replicated geometry stages; every rank's covered pixels go straight "
                                        "into a frame in GPU 0's memory (CUDA-IPC peer mapping over NVLink), GPU 0 re-clears the touched tiles")),
                       "parity_mode": "nearest sampling == reference; bilinear is an extension checked against oracle/axr_oracle.c"},
            "gpu_launches": launches_per_step * args.steps,
            "serial": {"ms_per_step": ms_serial, "value": units / (ms_serial * 1e-3) / 1e6, "unit": "Mtri/s",
                       "note": "same K steps with axr_set_overlap(0) and an event pair around every kernel"},
            "sustained": sustained,
            "kernel_ms": kavg, "draw_ms": draw_ms, "draw_stats": stats, "covered_pixels": covered,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": traffic, "algorithmic_bytes": per_stage[dom], "kernel_ms": dom_ms, "peak_source": peak_src,
                         "note": "the stage with the longest mean duration in the serial pass; the setup and shading stages are within a few percent of "
                                 "each other, so which one this is can change from run to run: roofline_kernels carries the same figure for every "
                                 "stage, roofline_frame the whole frame (the figure of record), roofline_gather / fp32_issue the bounds that bind "
                                 "the two big stages (scattered DRAM sectors, FP32 issue slots) rather than streaming HBM"},
            "roofline_frame": {"bound": "hbm", "algorithmic_bytes": b_alg, "achieved": b_alg / (ms_step * 1e-3) / 1e9,
                               "peak": peak, "unit": "GB/s", "frac": b_alg / (ms_step * 1e-3) / 1e9 / peak,
                               "note": "BASELINE.md §4 figure of record: B_alg / ms_per_step of the timed region (clear included, frames "
                                       "overlapped); serial, draw kernels only: frac = %.4f" % (b_alg / (draw_ms * 1e-3) / 1e9 / peak if draw_ms else 0.0)},
            "composite_parity": composite_parity,
            "clocks": clk,
        }
        # the same figure for every stage (the dominant one changes from round to round: setup and shading are within 10 % of each other)
        try:
            tj = json.load(open(tp)).get(args.workload, {})
        except Exception:
            tj = {}
        # The shading stage reads vertices, attributes and texels as data-dependent 16-byte gathers: its DRAM traffic (ncu) is held against
        # both ends — the streaming peak and the measured rate of randomly placed sectors on this GPU. It sits between the two.
        ts_bytes, ts_ms = tj.get("tile_shade"), kavg.get("tile_shade")
        if ts_bytes and ts_ms:
            line["roofline_gather"] = {"kernel": "tile_shade", "dram_bytes": ts_bytes, "kernel_ms": ts_ms, "achieved": ts_bytes / (ts_ms * 1e-3) / 1e9,
                                       "unit": "GB/s", "random_sector_peak": gather_rate * 32 / 1e9, "streaming_peak": peak,
                                       "frac_of_random_sector_peak": ts_bytes / (ts_ms * 1e-3) / (gather_rate * 32),
                                       "note": "axr_measure_gather: 16-byte loads at random 32-byte-aligned offsets of a 1 GiB buffer, 8 in flight per "
                                               "thread; fetching 2 or 4 consecutive sectors per position gives the same sectors/s, i.e. the memory "
                                               "system's limit for scattered reads is per sector, about a fifth of the streaming bandwidth"}
        line["roofline_kernels"] = {k: {"algorithmic_bytes": per_stage.get(k, 0), "kernel_ms": kavg[k],
                                        "achieved": (per_stage.get(k, 0) / (kavg[k] * 1e-3) / 1e9) if kavg[k] > 0 else 0.0,
                                        "frac": (per_stage.get(k, 0) / (kavg[k] * 1e-3) / 1e9 / peak) if kavg[k] > 0 else 0.0,
                                        "traffic": tj.get(k)} for k in kavg if per_stage.get(k, 0)}
        # per-kernel binding bound = max(T_hbm, T_fp32_issue): instruction counts are ncu's for this workload (profiles/traffic.json)
        counters = {}
        try:
            counters = json.load(open(tp)).get(args.workload + "_instructions", {})
        except Exception:
            counters = {}
        issue = {"measured_fmul_fadd_Gwinst_per_s": fp32_rate / 1e9, "measured_ffma_Gwinst_per_s": ffma_rate / 1e9,
                 "nominal_Gwinst_per_s": 148 * 4 * 1.965, "unit_note": "one warp-instruction = 32 lanes"}
        per_kernel = {}
        for k, c in counters.items():
            t_issue = c["warp_inst"] / fp32_rate * 1e3
            t_hbm = per_stage.get(k, 0) / (peak * 1e9) * 1e3
            per_kernel[k] = {"warp_inst": c["warp_inst"], "t_issue_ms": t_issue, "t_hbm_ms": t_hbm, "binding": "fp32_issue" if t_issue > t_hbm else "hbm",
                             "measured_ms": kavg.get(k), "frac_of_binding": (max(t_issue, t_hbm) / kavg[k]) if kavg.get(k) else None}
        issue["per_kernel"] = per_kernel
        if per_kernel:
            issue["note"] = ("issue-slot utilisation of the instruction stream the kernels execute (ncu counts), not an algorithmic "
                             "minimum and not a roofline fraction: the figure of record is roofline_frame")
        line["fp32_issue"] = issue

    if comp:
        comp.release()

    # ---- BASELINE config 5 as written (extra figures, same process group): the C3 scene at 8K split into screen bands, and 64 camera
    #      views (64 / N per GPU) composited to GPU 0 — each with its own composite parity check
    extras = {}
    if not args.no_extras and args.workload == "c3":
        dev.close()
        dev = None
        k5 = max(3, args.steps // 4)
        sc8k = build_workload("c5")
        extras["bands_c5"] = _run_multi(torch, dist, api, multi, sc8k, rank, world, local, "bands", "peer", k5, 3, fill=not args.no_fill, rows=not args.no_rows)
        extras["bands_c5"]["workload"] = workload_label("c5", sc8k)
        del sc8k
        extras["views64"] = _run_multi(torch, dist, api, multi, sc, rank, world, local, "views", "peer", max(1, k5 // 2), 1, views_total=64, fill=not args.no_fill, rows=not args.no_rows)
        extras["views64"]["workload"] = "64 camera views (yaw = i*2pi/64) of " + workload_label(args.workload, sc)
        if rank == 0:
            line["bands_c5"] = extras["bands_c5"]
            line["views64"] = extras["views64"]

    # ---- e2e: the reference-facing call (TiledPipeline.drawMesh on a HOST framebuffer) with H2D/D2H inside the timed region
    if not args.no_e2e:
        fb = api.Framebuffer(W, H, True)
        cam = api.Camera()
        cam.setViewport(0, 0, W, H)
        cam.setViewProjectionMatrix(sc.view_proj)
        cam._pos = np.asarray(sc.cam_pos, dtype=np.float32)
        pipe = api.TiledPipeline(os.cpu_count() or 1, cam, fb, device=local, sampler=sc.sampler)
        shader = [api.FlatShader(tuple(sc.light_dir)), api.PhongShader(tuple(sc.light_dir), tuple(sc.light_color)),
                  api.PBRShader(tuple(sc.light_dir), tuple(sc.light_color))][sc.shader]
        pipe.setShader(shader)
        tex = [api.Texture(t) if t is not None else None for t in sc.textures]
        mat = api.Material("m0", tex[0], tex[2], tex[1], tex[3], tex[4], sc.specular_exponent)
        hmesh = api.Mesh(sc.vertices, sc.indices, {"m0": mat})
        if dev is not None:
            dev.close()  # free HBM held by the device-resident arm
            dev = None
        n_e2e = max(3, min(args.steps, 10))

        def e2e_step():
            # Framebuffer::clearColor/clearDepth are the caller's own host-side calls before drawMesh (reference
            # src/renderer.cpp:124-125) and are excluded, exactly as in the CPU arm; the timed call is drawMesh alone:
            # H2D of the host framebuffer, the draw, D2H of the result, complete on return.
            fb.clearColor(api.Color(0, 0, 0, 255))
            fb.clearDepth()
            t0 = time.perf_counter()
            pipe.drawMesh(sc.model, hmesh)
            return time.perf_counter() - t0

        for _ in range(2):
            e2e_step()  # first call uploads and caches the mesh (the reference reads its host Mesh on every call)
        torch.cuda.synchronize()
        e2e_ms = sum(e2e_step() for _ in range(n_e2e)) * 1e3 / n_e2e
        e2e_ms = _max_over_ranks(torch, dist, e2e_ms)
        covered_e2e = int(np.isfinite(fb.getDepthData()).sum())
        if rank == 0:
            line["e2e"] = {"value": units / (e2e_ms * 1e-3) / 1e6, "unit": "Mtri/s", "ms_per_step": e2e_ms, "steps": n_e2e,
                           "h2d_bytes_per_step": int(pipe.last_h2d_bytes) + int(covered_e2e) * 4, "d2h_bytes_per_step": int(covered_e2e) * 8,
                           "call": pipe.host_path_description,
                           "mesh_upload_bytes_first_call": int(sc.vertices.nbytes + sc.indices.nbytes)}
        pipe.device.close()
    elif rank == 0:
        line["e2e"] = None
    if dev is not None:
        dev.close()

    # ---- CPU baseline (rank 0, N = 1 only): the reference's own thread-pool path on this box's host cores
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            v, info = cpu_reference_run(sc, args.cpu_seconds)
            line["cpu_baseline"] = {"value": v, "unit": "Mtri/s", "cores": info["cores"], "kind": info["kind"], "sample": info["sample"],
                                    "host_cpus": info["host_cpus"]}
        except Exception as e:  # the checker being absent must not lose the GPU numbers
            line["cpu_baseline"] = {"value": None, "unit": "Mtri/s", "cores": 0, "kind": "unavailable", "sample": repr(e)}
    if rank == 0:
        _emit(line)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

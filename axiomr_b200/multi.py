"""Multi-GPU sharding of the raster path: one process per GPU (torch.distributed / NCCL for the plumbing).

Rendering shards without any data-path collective in two ways:
  * views : every rank renders its own camera view of the replicated scene (weak scaling);
  * bands : every rank owns a contiguous range of 16-px tile rows of ONE frame, runs the geometry stages over the
            replicated mesh and rasterises / shades only its rows (strong scaling, Amdahl-limited by the geometry stages).
The single exchange step is the composite to GPU 0 — a gather of disjoint regions, no reduction:
  * NCCL grouped send/recv straight out of / into the framebuffer allocations (`Compositor`, mode "nccl"), or
  * fused: the resolve stores of every rank go directly into GPU 0's framebuffer through a CUDA-IPC peer mapping
    (`Compositor`, mode "peer"; bands only), so the transfer rides NVLink while the tile kernel is still shading.
"""
from __future__ import annotations

REF_TILE = 16  # reference include/tiled_pipeline.hpp:28


def band_rows(height: int, world: int):
    """Split ceil(H/16) reference tile rows into `world` contiguous bands, as evenly as possible
    (8K: 270 tile rows -> 34,34,34,34,34,34,33,33). Returns [(y0, y1)] in pixels; empty bands are not produced."""
    rows = (height + REF_TILE - 1) // REF_TILE
    world = max(1, min(world, rows))
    base, extra = divmod(rows, world)
    out, r = [], 0
    for i in range(world):
        n = base + (1 if i < extra else 0)
        out.append((r * REF_TILE, min((r + n) * REF_TILE, height)))
        r += n
    return out


def views_for_rank(rank: int, world: int, n_views: int):
    """Round-robin-free contiguous assignment of n_views camera views to ranks."""
    base, extra = divmod(n_views, world)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def gather_bands(color, depth, bands, rank: int, dist):
    """Bands -> rank 0: every rank r>0 sends its rows [y0,y1) of colour and depth; rank 0 receives them in place
    (row ranges are contiguous in memory, so this is a zero-copy composite). Works on any torch.distributed backend."""
    ops = []
    if rank == 0:
        for r in range(1, len(bands)):
            y0, y1 = bands[r]
            ops.append(dist.P2POp(dist.irecv, color[y0:y1], r))
            ops.append(dist.P2POp(dist.irecv, depth[y0:y1], r))
    elif rank < len(bands):
        y0, y1 = bands[rank]
        ops.append(dist.P2POp(dist.isend, color[y0:y1], 0))
        ops.append(dist.P2POp(dist.isend, depth[y0:y1], 0))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def gather_views(color, depth, rank: int, world: int, dist, slots=None):
    """Views -> rank 0: slot r-1 receives rank r's finished frame (rank 0's own frame stays in its framebuffer)."""
    import torch
    ops = []
    if rank == 0:
        if slots is None:
            slots = (torch.empty((world - 1,) + tuple(color.shape), dtype=color.dtype, device=color.device),
                     torch.empty((world - 1,) + tuple(depth.shape), dtype=depth.dtype, device=depth.device))
        for r in range(1, world):
            ops.append(dist.P2POp(dist.irecv, slots[0][r - 1], r))
            ops.append(dist.P2POp(dist.irecv, slots[1][r - 1], r))
    else:
        ops.append(dist.P2POp(dist.isend, color, 0))
        ops.append(dist.P2POp(dist.isend, depth, 0))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return slots


class _DevArray:
    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3}


def framebuffer_tensors(dev):
    """Zero-copy torch views of a Device's framebuffer: (colour as int32 [H,W] = packed BGRA words, depth f32 [H,W])."""
    import torch
    c, d = dev.framebuffer_device()
    dv = torch.device("cuda", torch.cuda.current_device())
    color = torch.as_tensor(_DevArray(c, (dev.height, dev.width), "<i4"), device=dv)
    depth = torch.as_tensor(_DevArray(d, (dev.height, dev.width), "<f4"), device=dv)
    return color, depth


class Compositor:
    """Gathers every rank's finished region to GPU 0 after each frame (the path's one exchange step).

    transport "peer" (default) — composite fused into the tile kernel over NVLink peer memory (CUDA IPC):
      bands : every rank maps GPU 0's framebuffer; its clear and resolve stores go straight into its own rows there. Rows are
              disjoint, so there is no per-frame synchronisation at all.
      views : GPU 0 owns two sets of per-view slots (colour + depth). Rank r > 0 points its output at slot (set i%2, r): its tile
              kernel stores the covered pixels of frame i directly into GPU 0's memory (no depth read-back: the slot is freshly
              cleared, axr_set_depth_read(0)). GPU 0 clears the *other* set for frame i+1 before its own draw; one tiny NCCL all-reduce per frame on the render streams orders "all stores of frame i are done"
              before "the set is cleared again". Only covered pixels cross NVLink (≈ 25 MB instead of 66 MB per 4K view).
    transport "nccl" — the baseline: grouped send/recv of whole regions after each frame (bands: in place; views:
      double-buffered, on a second stream, overlapping the next frame).
    """

    def __init__(self, dev, rank: int, world: int, mode: str, band, stream, transport: str | None = None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.dev, self.rank, self.world, self.mode, self.stream = dev, rank, world, mode, stream
        self.transport = transport or "peer"
        self.launches_per_step = 0
        self.color, self.depth = framebuffer_tensors(dev)
        self.step = 0
        dv = self.color.device
        self.H, self.W = dev.height, dev.width
        if mode == "views" and self.transport == "nccl":
            self.comm = torch.cuda.Stream(device=dv)
            self.done = [None, None]   # comm-stream events: buffer b has been sent / slot set b has been filled
            shape = (self.H, self.W)
            if rank == 0:
                self.slots = [(torch.empty((world - 1,) + shape, dtype=torch.int32, device=dv),
                               torch.empty((world - 1,) + shape, dtype=torch.float32, device=dv)) for _ in range(2)]
            else:
                self.bufs = [(torch.empty(shape, dtype=torch.int32, device=dv), torch.empty(shape, dtype=torch.float32, device=dv))
                             for _ in range(2)]
        elif mode == "views":
            self._setup_peer_views()
        else:
            self.bands = band_rows(dev.height, world)
            if self.transport == "peer":
                self._setup_peer_bands()

    # ------------------------------------------------------------------ bands
    def _setup_peer_bands(self):
        dist = self.dist
        handles = [None]
        if self.rank == 0:
            handles = [self.dev.framebuffer_ipc()]
        dist.broadcast_object_list(handles, src=0)
        if self.rank != 0:
            ch, dh = handles[0]
            self._peer = (self.dev.open_ipc(ch), self.dev.open_ipc(dh))
            self.dev.set_output(*self._peer)
        dist.barrier()

    # ------------------------------------------------------------------ views over peer memory
    def _slot_offsets(self, b: int, r: int):
        """Byte offsets of (colour, depth) of slot (set b, rank r) inside the shared allocation.
        Per set: [colour of ranks 1..N-1, contiguous][depth of ranks 1..N-1, contiguous] so a set is cleared by two flat fills."""
        npx = self.H * self.W
        base = b * (self.world - 1) * npx * 8
        return base + (r - 1) * npx * 4, base + (self.world - 1) * npx * 4 + (r - 1) * npx * 4

    def _setup_peer_views(self):
        torch, dist = self.torch, self.dist
        npx = self.H * self.W
        total = 2 * (self.world - 1) * npx * 8
        handles = [None]
        if self.rank == 0:
            self._shared_ptr, h = self.dev.alloc_shared(total)
            handles = [h]
        dist.broadcast_object_list(handles, src=0)
        dv = self.color.device
        self.token = torch.zeros(1, dtype=torch.int32, device=dv)
        if self.rank == 0:
            self.slots = []
            for b in range(2):
                co, do = self._slot_offsets(b, 1)
                shape = ((self.world - 1), self.H, self.W)
                self.slots.append((torch.as_tensor(_DevArray(self._shared_ptr + co, shape, "<i4"), device=dv),
                                   torch.as_tensor(_DevArray(self._shared_ptr + do, shape, "<f4"), device=dv)))
            # cleared templates: a set is re-cleared with device-to-device copies (copy engines, no SM time) on a side stream
            # while GPU 0 renders its own view
            self.side = torch.cuda.Stream(device=dv)
            self.tmpl_c = torch.full((self.H, self.W), -16777216, dtype=torch.int32, device=dv)   # 0xFF000000
            self.tmpl_d = torch.full((self.H, self.W), float("inf"), dtype=torch.float32, device=dv)
            for b in range(2):
                self._clear_set(b)
            torch.cuda.synchronize()
        else:
            self._shared_ptr = self.dev.open_ipc(handles[0])
            self.dev.set_depth_read(False)
        dist.barrier()

    def _clear_set(self, b: int):
        for r in range(self.world - 1):
            self.slots[b][0][r].copy_(self.tmpl_c, non_blocking=True)
            self.slots[b][1][r].copy_(self.tmpl_d, non_blocking=True)

    # ------------------------------------------------------------------ per frame
    @property
    def clears_own_target(self) -> bool:
        """views over peer memory: the slot a rank renders into was cleared by GPU 0 already; the rank must not clear it."""
        return not (self.mode == "views" and self.transport == "peer" and self.rank != 0)

    def begin_step(self):
        """Call before the frame's clear: selects the output buffer of this frame."""
        b = self.step % 2
        if self.mode != "views":
            return
        if self.transport == "nccl":
            if self.rank != 0:
                if self.done[b] is not None:
                    self.stream.wait_event(self.done[b])      # the send of the frame rendered two steps ago has finished
                c, d = self.bufs[b]
                self.dev.set_output(c.data_ptr(), d.data_ptr())
            return
        if self.rank != 0:
            co, do = self._slot_offsets(b, self.rank)
            self.dev.set_output(self._shared_ptr + co, self._shared_ptr + do)
        else:
            # Clear the other set for the next frame while this one renders: device-to-device copies of a cleared template on a
            # side stream, ordered after the previous frame's all-reduce. (Fill KERNELS on a side stream were measured slower than
            # no overlap at all — they fight the vertex / setup kernels for SM slots: N = 8, +220 us on GPU 0.)
            ready = self.torch.cuda.Event()
            ready.record(self.stream)
            with self.torch.cuda.stream(self.side):
                self.side.wait_event(ready)
                self._clear_set((b + 1) % 2)

    def composite(self):
        torch, dist = self.torch, self.dist
        b = self.step % 2
        self.step += 1
        if self.mode == "bands":
            if self.transport == "peer":
                return  # the tile kernel already stored this rank's band into GPU 0's framebuffer over NVLink
            with torch.cuda.stream(self.stream):
                gather_bands(self.color, self.depth, self.bands, self.rank, dist)
            return
        if self.transport == "peer":
            with torch.cuda.stream(self.stream):
                if self.rank == 0:
                    self.stream.wait_stream(self.side)    # the next frame's set is clear
                dist.all_reduce(self.token)                # every rank's stores of this frame precede anything after it
            return
        ready = torch.cuda.Event()
        ready.record(self.stream)                         # this frame is rendered
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(ready)
            if self.rank == 0:
                gather_views(None, None, 0, self.world, dist, self.slots[b])
            else:
                gather_views(self.bufs[b][0], self.bufs[b][1], self.rank, self.world, dist)
            ev = torch.cuda.Event()
            ev.record(self.comm)
            self.done[b] = ev

    def finish(self):
        """Make the render stream wait for every outstanding transfer (call before the closing synchronisation)."""
        if self.mode == "views" and self.transport == "nccl":
            self.stream.wait_stream(self.comm)
        elif self.mode == "views" and self.rank == 0:
            self.stream.wait_stream(self.side)

    def view_slot(self, b: int, r: int):
        """(colour int32 HxW, depth f32 HxW) of rank r's frame in slot set b, on GPU 0."""
        return self.slots[b][0][r - 1], self.slots[b][1][r - 1]

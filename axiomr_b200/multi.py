"""Multi-GPU sharding of the raster path: one process per GPU (torch.distributed / NCCL for the plumbing).

Rendering shards without any data-path collective in two ways:
  * views : every rank renders its own camera view of the replicated scene (weak scaling);
  * bands : every rank owns a contiguous range of 16-px tile rows of ONE frame, runs the geometry stages over the
            replicated mesh and rasterises / shades only its rows (strong scaling, Amdahl-limited by the geometry stages).
The single exchange step is the composite to GPU 0 — a gather of disjoint regions, no reduction:
  * NCCL grouped send/recv straight out of / into the framebuffer allocations (`Compositor`, mode "nccl"), or
  * fused: the resolve stores of every rank go directly into targets in GPU 0's memory through a CUDA-IPC peer mapping
    (`Compositor`, mode "peer"), so the transfer rides NVLink while the tile kernel is still shading.
"""
from __future__ import annotations

REF_TILE = 16  # reference include/tiled_pipeline.hpp:28


def band_rows(height: int, world: int, granule: int = REF_TILE):
    """Split ceil(H/granule) rows of `granule` pixels into `world` contiguous bands, as evenly as possible (granule 16, the
    reference tile: 8K -> 270 tile rows -> 34,34,34,34,34,34,33,33). Any multiple of 16 gives the same pixels; the fused peer
    composite uses 32 (the GPU tile) so that no tile — the unit its dirty flags track — belongs to two ranks.
    Returns [(y0, y1)] in pixels; empty bands are not produced."""
    REF_TILE = granule  # noqa: N806 (local alias: the arithmetic below is the same for either granule)
    rows = (height + REF_TILE - 1) // REF_TILE
    world = max(1, min(world, rows))
    base, extra = divmod(rows, world)
    out, r = [], 0
    for i in range(world):
        n = base + (1 if i < extra else 0)
        out.append((r * REF_TILE, min((r + n) * REF_TILE, height)))
        r += n
    return out


def band_granule(transport: str | None) -> int:
    """Band alignment in pixels: the reference tile (16) in general; the fused peer composite tracks 32-px GPU tiles."""
    return 32 if (transport or "peer") == "peer" else REF_TILE


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

    transport "peer" (default) — composite fused into the tile kernel over NVLink peer memory (CUDA IPC). GPU 0 owns two sets of
      targets (colour + depth + a dirty-tile map each); frame i uses set i % 2:
        views : one target per rank r > 0 (GPU 0 renders its own view into its own framebuffer);
        bands : one full-frame target every rank, GPU 0 included, renders its rows into.
      A rank's tile kernel stores into the target directly (no depth read over NVLink: the target holds a clear plus this one draw)
      and overwrites EVERY pixel of the 32x32 tiles it touches — shaded colour or the clear values (axr_set_output_fill) — in whole
      128-byte rows, flagging those tiles in the target's dirty map. Nothing has to be cleared in front of a frame that way: GPU 0
      only clears the tiles the target's previous use touched and this one did not (axr_clear_stale_tiles on a side stream; two maps
      per target alternate; no tile at all while the cameras stand still), and one 4-byte NCCL all-reduce per frame on the render
      streams orders "all stores of frame i are done" before the set's next use. Only the touched tiles cross NVLink.
      (fill=False keeps the previous scheme: covered pixels only, GPU 0 re-clears the flagged tiles, axr_clear_dirty_tiles.)
    transport "nccl" — the baseline: grouped send/recv of whole regions after each frame (bands: in place; views:
      double-buffered, on a second stream, overlapping the next frame).
    """

    def __init__(self, dev, rank: int, world: int, mode: str, band, stream, transport: str | None = None, fill: bool = True, rows: bool = True):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.dev, self.rank, self.world, self.mode, self.stream = dev, rank, world, mode, stream
        self.transport = transport or "peer"
        self.fill = bool(fill) and self.transport == "peer"
        self.rows = bool(rows) and self.transport == "peer"   # remote stores in 128-byte rows
        self.launches_per_step = 0
        self.color, self.depth = framebuffer_tensors(dev)
        self.step = 0
        dv = self.color.device
        self.H, self.W = dev.height, dev.width
        if mode == "bands":
            self.bands = band_rows(dev.height, world, band_granule(self.transport))
        if self.transport == "peer":
            self._setup_peer()
        elif mode == "views":
            self.comm = torch.cuda.Stream(device=dv)
            self.done = [None, None]   # comm-stream events: buffer b has been sent / slot set b has been filled
            shape = (self.H, self.W)
            if rank == 0:
                self.slots = [(torch.empty((world - 1,) + shape, dtype=torch.int32, device=dv),
                               torch.empty((world - 1,) + shape, dtype=torch.float32, device=dv)) for _ in range(2)]
            else:
                self.bufs = [(torch.empty(shape, dtype=torch.int32, device=dv), torch.empty(shape, dtype=torch.float32, device=dv))
                             for _ in range(2)]

    # ------------------------------------------------------------------ targets on GPU 0 (peer transport)
    @property
    def n_targets(self) -> int:
        return self.world - 1 if self.mode == "views" else 1

    def _offsets(self, b: int, t: int, m: int = 0):
        """Byte offsets of (colour, depth, dirty map m) of target t of set b inside the shared allocation.
        Per set: [colour planes][depth planes][dirty maps 0][dirty maps 1], each group contiguous (what the clear kernels walk)."""
        npx, nt, n = self.H * self.W, self.n_tiles, self.n_targets
        set_bytes = n * (npx * 8 + 2 * nt * 4)
        base = b * set_bytes
        return base + t * npx * 4, base + n * npx * 4 + t * npx * 4, base + n * npx * 8 + m * n * nt * 4 + t * nt * 4

    def _target_index(self) -> int:
        return self.rank - 1 if self.mode == "views" else 0

    def _setup_peer(self):
        torch, dist = self.torch, self.dist
        npx = self.H * self.W
        self.n_tiles = self.dev.dirty_map_entries()
        total = 2 * self.n_targets * (npx * 8 + 2 * self.n_tiles * 4)
        self.uses = [0, 0]   # how often each set has been rendered into: picks which of its two dirty maps is "now"
        handles = [None]
        if self.rank == 0:
            self._shared_ptr, h = self.dev.alloc_shared(total)
            handles = [h]
        dist.broadcast_object_list(handles, src=0)
        dv = self.color.device
        self.token = torch.zeros(1, dtype=torch.int32, device=dv)
        if self.rank == 0:
            self.slots = []
            n = self.n_targets
            for b in range(2):
                co, do, mo = self._offsets(b, 0)
                c = torch.as_tensor(_DevArray(self._shared_ptr + co, (n, self.H, self.W), "<i4"), device=dv)
                d = torch.as_tensor(_DevArray(self._shared_ptr + do, (n, self.H, self.W), "<f4"), device=dv)
                m = torch.as_tensor(_DevArray(self._shared_ptr + mo, (2, n, self.n_tiles), "<i4"), device=dv)
                c.fill_(-16777216)   # 0xFF000000
                d.fill_(float("inf"))
                m.zero_()
                self.slots.append((c, d, m))
            self.side = torch.cuda.Stream(device=dv)
            torch.cuda.synchronize()
        else:
            self._shared_ptr = self.dev.open_ipc(handles[0])
        if self.rank != 0 or self.mode == "bands":
            self.dev.set_depth_read(False)   # the target is freshly cleared and receives exactly this draw
            if self.fill:
                self.dev.set_output_fill(True)
            if self.rows and self.rank != 0:
                self.dev.set_output_rows(True)
        dist.barrier()

    # ------------------------------------------------------------------ per frame
    @property
    def clears_own_target(self) -> bool:
        """peer transport: the target a rank renders into was cleared by GPU 0 already; only GPU 0's own view (views mode) is the
        rank's to clear."""
        if self.transport != "peer":
            return True
        return self.mode == "views" and self.rank == 0

    def begin_step(self):
        """Call before the frame's clear: selects the output buffer of this frame."""
        b = self.step % 2
        if self.transport == "nccl":
            if self.mode == "views" and self.rank != 0:
                if self.done[b] is not None:
                    self.stream.wait_event(self.done[b])      # the send of the frame rendered two steps ago has finished
                c, d = self.bufs[b]
                self.dev.set_output(c.data_ptr(), d.data_ptr())
            return
        now = self.uses[b] % 2 if self.fill else 0   # every rank counts the uses of a set alike
        if self.rank != 0 or self.mode == "bands":
            co, do, mo = self._offsets(b, self._target_index(), now)
            self.dev.set_output(self._shared_ptr + co, self._shared_ptr + do)
            self.dev.set_dirty_map(self._shared_ptr + mo)
        if self.rank == 0:
            # The other set (the previous frame's, complete since that frame's all-reduce) is put in order for its next use while this
            # frame renders, on a side stream.
            ready = self.torch.cuda.Event()
            ready.record(self.stream)
            self.side.wait_event(ready)
            self._tidy((b + 1) % 2)
        self.uses[b] += 1

    def _tidy(self, nb: int):
        """GPU 0, side stream: fill mode clears the tiles the set's last use no longer touched (idempotent), else all flagged tiles."""
        if self.fill:
            if self.uses[nb] > 0:
                last = (self.uses[nb] - 1) % 2   # the map the set's last use flagged; the other one holds the use before it
                co, do, m_now = self._offsets(nb, 0, last)
                m_prev = self._offsets(nb, 0, 1 - last)[2]
                self.dev.clear_stale_tiles(self._shared_ptr + co, self._shared_ptr + do, self._shared_ptr + m_prev, self._shared_ptr + m_now,
                                           self.n_targets, stream=self.side.cuda_stream)
        else:
            co, do, mo = self._offsets(nb, 0)
            self.dev.clear_dirty_tiles(self._shared_ptr + co, self._shared_ptr + do, self._shared_ptr + mo, self.n_targets,
                                       stream=self.side.cuda_stream)

    def composite(self):
        torch, dist = self.torch, self.dist
        b = self.step % 2
        self.step += 1
        if self.transport == "peer":
            with torch.cuda.stream(self.stream):
                if self.rank == 0:
                    self.stream.wait_stream(self.side)    # the next frame's set is clear
                dist.all_reduce(self.token)                # every rank's stores of this frame precede anything after it
            return
        if self.mode == "bands":
            with torch.cuda.stream(self.stream):
                gather_bands(self.color, self.depth, self.bands, self.rank, dist)
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
        if self.transport == "nccl":
            if self.mode == "views":
                self.stream.wait_stream(self.comm)
        elif self.rank == 0:
            if self.fill and self.step > 0:   # the last frame's set: tiles its previous use touched and this one did not
                ready = self.torch.cuda.Event()
                ready.record(self.stream)
                self.side.wait_event(ready)
                self._tidy(self.last_set())
            self.stream.wait_stream(self.side)

    def last_set(self) -> int:
        """Index of the set the most recent frame went to."""
        return (self.step - 1) % 2

    def view_slot(self, b: int, r: int):
        """(colour int32 HxW, depth f32 HxW) of rank r's frame in slot set b, on GPU 0 (views)."""
        return self.slots[b][0][r - 1], self.slots[b][1][r - 1]

    def frame(self, b: int):
        """(colour int32 HxW, depth f32 HxW) of the composited frame of set b, on GPU 0 (bands, peer transport)."""
        return self.slots[b][0][0], self.slots[b][1][0]

    def release(self):
        """Detach the context from the shared targets (before closing it)."""
        if self.transport == "peer":
            self.dev.set_output(None, None)
            self.dev.set_dirty_map(None)
            self.dev.set_output_fill(False)
            self.dev.set_output_rows(False)

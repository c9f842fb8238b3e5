"""Screen-space sharding across ranks (one process per GPU) and the framebuffer gather.

The reference parallelises Scene::raytraceImage over image rows with OpenMP dynamic chunks (Scene.cpp:113); here
rank r of N owns the rows with row % N == r (perfect balance, every row stays a coherent run of pixels), the
scene is replicated, and the only exchange is one all-gather of the row shards -- NCCL over NVLink on GPUs, gloo
in the CPU tests.  No torch type crosses the C ABI: the render call gets (row_begin, row_end, row_stride,
row_phase) = rows_of_rank(...).
"""
import torch
import torch.distributed as dist


def rows_of_rank(height, world, rank):
    """Render-parameter row tuple for rank `rank` of `world`, and how many rows that is."""
    return (0, height, world, rank), len(range(rank, height, world))


def max_rows(height, world):
    return (height + world - 1) // world


def gather_rows(local_rows, height, world, rank, out=None, group=None):
    """local_rows: (rows_of_this_rank, width, C) tensor.  Returns the full (height, width, C) frame on every rank.
    Shards are padded to the same row count so a single all_gather_into_tensor moves everything."""
    if world == 1:
        if out is None:
            return local_rows
        out.copy_(local_rows)
        return out
    nmax = max_rows(height, world)
    width, ch = local_rows.shape[1], local_rows.shape[2]
    send = torch.zeros((nmax, width, ch), dtype=local_rows.dtype, device=local_rows.device)
    send[: local_rows.shape[0]].copy_(local_rows)
    recv = torch.empty((world * nmax, width, ch), dtype=local_rows.dtype, device=local_rows.device)   # concatenated along dim 0
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.view(world, nmax, width, ch)
    if out is None:
        out = torch.empty((height, width, ch), dtype=local_rows.dtype, device=local_rows.device)
    for r in range(world):
        k = len(range(r, height, world))
        out[r::world].copy_(recv[r, :k])
    return out


class RowGather:
    """gather_rows with its buffers allocated once (one per frame size / dtype): the per-frame cost is one copy of this
    rank's rows into the send buffer, one all_gather_into_tensor, and one strided copy that puts the rows back in frame order."""

    def __init__(self, height, width, channels, dtype, device, world, rank, group=None):
        self.height, self.world, self.rank, self.group = height, world, rank, group
        self.nmax = max_rows(height, world)
        self.send = torch.zeros((self.nmax, width, channels), dtype=dtype, device=device)
        self.recv = torch.empty((world * self.nmax, width, channels), dtype=dtype, device=device)
        self.out = torch.empty((height, width, channels), dtype=dtype, device=device)

    def __call__(self, frame):
        """frame: full-frame tensor whose rows rank::world are valid on this rank.  Returns the complete frame."""
        if self.world == 1:
            return frame
        mine = frame[self.rank::self.world]
        self.send[: mine.shape[0]].copy_(mine)
        dist.all_gather_into_tensor(self.recv, self.send, group=self.group)
        if self.height % self.world == 0:      # every shard has nmax rows: one strided copy
            self.out.view(self.nmax, self.world, *self.out.shape[1:]).copy_(self.recv.view(self.world, self.nmax, *self.out.shape[1:]).transpose(0, 1))
        else:
            recv = self.recv.view(self.world, self.nmax, *self.out.shape[1:])
            for r in range(self.world):
                k = len(range(r, self.height, self.world))
                self.out[r::self.world].copy_(recv[r, :k])
        return self.out


def reduce_max(value, group=None):
    """Max over ranks of one float (the NaN-replacement intensity of the tone map, Scene.cpp:157-164)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return value
    dist.all_reduce(value, op=dist.ReduceOp.MAX, group=group)
    return value


class FramePipeline:
    """Frames of a sequence kept in flight across ranks: while this rank renders frame i + 1, frame i's exchange -- the
    all_reduce(max) of the tone map's one frame-wide constant (Scene.cpp:157-164), the tone map of this rank's rows, the
    NCCL all_gather of the 8-bit row shards and rank 0's copy into page-locked host memory -- runs on a side stream.
    Each slot owns its float frame, 8-bit frame, gather buffers and host frame, so a frame's data is never touched by
    its successor; a slot is reused only after its previous exchange finished (stream-ordered, no host sync).

        pipe = FramePipeline(scene, height, width, world, rank, device)
        for cam, params in sequence:
            slot = pipe.submit(cam, params)          # params carry this rank's rows (rows_of_rank)
            rays += pipe.rays_traced(slot)           # waits for the RENDER of this frame only
        pipe.drain()                                 # all exchanges done; pipe.host_frame(slot) on rank 0 is complete
    """

    def __init__(self, scene, height, width, world, rank, device, depth=2, group=None, replicas=()):
        """replicas: further handles of the same scene (a handle renders one frame at a time -- its wave buffers are its own).
        With k handles, k consecutive frames render concurrently, each on its own stream, so the ragged end of one frame's
        persistent launches is covered by its neighbours' kernels (the same effect as bench.py's steps in flight); the
        exchanges still run in frame order on the one side stream, so every rank issues its collectives in the same order."""
        self.S, self.world, self.rank, self.group = scene, world, rank, group
        self.scenes = [scene] + list(replicas)
        self.render_streams = [None] + [torch.cuda.Stream(device) for _ in replicas]   # None: the caller's current stream
        depth = max(depth, len(self.scenes) + 1)
        self.rows, _ = rows_of_rank(height, world, rank)
        self.side = torch.cuda.Stream(device)
        self.slots = []
        for _ in range(depth):
            self.slots.append({
                "rgb": torch.zeros((height, width, 3), dtype=torch.float32, device=device),
                "u8": torch.zeros((height, width, 3), dtype=torch.uint8, device=device),
                "max": torch.empty(1, dtype=torch.float32, device=device),
                "gather": RowGather(height, width, 3, torch.uint8, device, world, rank, group),
                "host": torch.empty((height, width, 3), dtype=torch.uint8).pin_memory() if rank == 0 else None,
                "rendered": torch.cuda.Event(), "exchanged": torch.cuda.Event(),
            })
        self.count = 0
        self.timings = []     # per submitted frame: (exchange begin, exchange end) events on the side stream, see exchange_ms()

    def submit(self, cam, params):
        s = self.slots[self.count % len(self.slots)]
        k = self.count % len(self.scenes)
        self.count += 1
        s["scene"] = self.scenes[k]
        main = torch.cuda.current_stream()
        rs = self.render_streams[k] or main
        if rs is not main:
            fork = torch.cuda.Event()
            fork.record(main)
            rs.wait_event(fork)                            # whatever the caller queued before this frame
        with torch.cuda.stream(rs):
            rs.wait_event(s["exchanged"])                  # the slot's previous frame has left it
            s["scene"].render_device(cam, params, s["rgb"])   # this rank's rows, float radiance
            s["rendered"].record(rs)
        with torch.cuda.stream(self.side):
            self.side.wait_event(s["rendered"])
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record(self.side)
            self.S.frame_max_device(s["rgb"], self.rows, s["max"])
            if self.world > 1:
                dist.all_reduce(s["max"], op=dist.ReduceOp.MAX, group=self.group)
            self.S.tonemap_rows_rgb8_device(s["rgb"], self.rows, s["max"], s["u8"])
            full = s["gather"](s["u8"])
            if self.rank == 0:
                s["host"].copy_(full, non_blocking=True)
            t1.record(self.side)
            s["exchanged"].record(self.side)
            self.timings.append((t0, t1))
            if len(self.timings) > 1024:       # long sequences: keep the recent frames only
                del self.timings[:512]
        return s

    def rays_traced(self, slot):
        slot["rendered"].synchronize()
        return slot.get("scene", self.S).last_call_stats()[0]

    def host_frame(self, slot):
        return slot["host"]

    def exchange_ms(self, last=None):
        """Mean device time of a frame's exchange (frame maximum, all_reduce, tone map, all_gather, reorder, rank 0's copy to the
        host) on the side stream, over the last `last` submitted frames (all if None).  Call after drain()."""
        ev = self.timings if last is None else self.timings[-last:]
        return sum(a.elapsed_time(b) for a, b in ev) / max(1, len(ev))

    def drain(self):
        for rs in self.render_streams:
            if rs is not None:
                rs.synchronize()
        self.side.synchronize()

"""Sort-first rendering across the GPUs of one node: one process per GPU (torchrun), the volume
replicated, the image work dealt out, every rank's raycast kernel storing its pixels straight into a
ring of frames in rank 0's memory over NVLink (vkrt_sortfirst_* in include/vokselis_rt.h).
torch.distributed is only the plumbing: it ships the 80-byte IPC handle and provides barriers; there
is no data-path collective — the pixel exchange IS the kernel's own peer stores, and arrival /
slot-reuse are device-side flags.

Two granularities of the same scheme:
  * "tiles"  — every frame is split into tiles dealt round-robin over the ranks (the reference's own
               `tile` entry + Offset table, shaders/raycast_compute.wgsl:139-144,
               examples/xor/main.rs:80-95,242-253). Cuts the latency of big frames (4K, large volumes).
  * "frames" — whole frames are dealt round-robin (tile = the frame). For frames that one GPU
               renders in a fraction of a millisecond, splitting them only adds tail effects; dealing
               frames keeps every GPU busy and rank 0 still receives every frame, in order.
               With batch > 1, GROUPS of `batch` consecutive frames are dealt round-robin and a rank renders
               its group in ONE launch (grid.z = frame) into consecutive ring slots: one 1080p frame does not
               fill a B200 (profiles/r01_batch.md), a few do.
"""
from __future__ import annotations

import time

import numpy as np

from . import rt


MAX_SLOTS = 256  # kSfMaxSlots in csrc/api.cu


def choose_batch(steps: int, world: int) -> int:
    """Frames per launch for `steps` frames dealt to `world` ranks in groups, round-robin: the group size that leaves the
    busiest rank with the fewest frames (360 frames in groups of 8 give 4 ranks 12/11/11/11 groups; groups of 6 give 15
    each); among equals the largest, because a launch of more frames fills the GPU better. A short last group is padded
    to a full one, hence the ceilings. Pure dealing arithmetic: no measured costs enter."""
    def busiest_frames(b):
        groups = -(-steps // b)
        return -(-groups // world) * b
    sizes = [b for b in range(1, rt.MAX_BATCH + 1) if 2 * world * b <= MAX_SLOTS]
    best = min(busiest_frames(b) for b in sizes)
    return max(b for b in sizes if busiest_frames(b) == best)


def partition_tiles(width: int, height: int, tile: int, rank: int, world: int) -> np.ndarray:
    return rt.sortfirst_partition(width, height, tile, rank, world)


def frame_owner(frame_index: int, world: int, batch: int = 1) -> int:
    """'frames' granularity: which rank renders frame f (groups of `batch` consecutive frames, round-robin)."""
    return (frame_index // batch) % world


class SortFirstGroup:
    def __init__(self, ctx: rt.Context, rank: int, world: int, granularity: str = "tiles", tile: int = 120, slots: int | None = None,
                 dist=None, batch: int = 1):
        if dist is None and world > 1:
            import torch.distributed as dist  # noqa: PLC0415
        if granularity not in ("tiles", "frames"):
            raise ValueError(granularity)
        self.ctx, self.rank, self.world, self.dist, self.granularity = ctx, rank, world, dist, granularity
        self.batch = batch  # frames per launch: groups of whole frames ("frames") or of every rank's tile shares ("tiles")
        if not 1 <= self.batch <= (rt.MAX_BATCH if granularity == "frames" else 15):
            raise ValueError(f"batch must be 1..{rt.MAX_BATCH} (frames) / 1..15 (tiles)")
        self.slots = slots if slots is not None else (2 * self.batch if granularity == "tiles" else 2 * world * self.batch)
        if self.slots % self.batch or not 2 <= self.slots <= MAX_SLOTS:
            raise ValueError(f"slots must be a multiple of batch in 2..{MAX_SLOTS}")
        self.tiles = None
        if granularity == "tiles":
            p = ctx.get_params()
            p.tile_size = tile
            ctx.set_params(p)
            self.tiles = partition_tiles(ctx.width, ctx.height, tile, rank, world)
        box = [ctx.sortfirst_create_root(world, self.slots) if rank == 0 else None]
        if world > 1:  # (world == 1: the same pipeline on one GPU — ring, rotating render streams — without any plumbing)
            dist.broadcast_object_list(box, src=0)
            if rank != 0:
                ctx.sortfirst_join(rank, box[0])
            dist.barrier()
        self.frame = 0       # next frame index to submit (same on every rank)
        self.consumed = 0    # root: next frame to wait for

    def participants(self, f: int) -> int:
        return self.world if self.granularity == "tiles" else 1

    def arrivals_target(self, f: int) -> int:
        """Cumulative arrivals on f's slot once f is complete (see vkrt_sortfirst_wait)."""
        return (f // self.slots + 1) * self.participants(f)

    def submit(self, cam) -> int:
        """Every rank calls this once per frame, in the same order. Asynchronous. Returns the frame index."""
        f = self.frame
        self.frame += 1
        if self.batch > 1 and self.granularity == "tiles":
            raise RuntimeError("with batch > 1 use submit_tiles_batch")
        if self.granularity == "tiles":
            self.ctx.sortfirst_render(cam, self.tiles, f)
        elif self.batch > 1:
            raise RuntimeError("with batch > 1 use submit_batch")
        elif frame_owner(f, self.world) == self.rank:
            self.ctx.sortfirst_render(cam, None, f)
        return f

    def submit_batch(self, cams, flush_l2: bool = False) -> int:
        """'frames' granularity with batch > 1: every rank calls this once per group of len(cams) <= batch
        consecutive frames, in the same order; the owning rank renders the group in one launch. Returns the
        index of the group's first frame."""
        assert self.granularity == "frames" and 1 <= len(cams) <= self.batch and self.frame % self.batch == 0
        cams = list(cams) + [cams[-1]] * (self.batch - len(cams))  # a short last group is padded: every slot gets its arrival
        f = self.frame
        self.frame += self.batch
        if frame_owner(f, self.world, self.batch) == self.rank:
            self.ctx.sortfirst_render_batch(cams, f, flush_l2=flush_l2)
        return f

    def submit_tiles_batch(self, cams) -> int:
        """'tiles' granularity with batch > 1: every rank calls this once per group of len(cams) <= batch consecutive
        frames, in the same order, and renders ITS tiles of all of them in one launch. Returns the group's first frame."""
        assert self.granularity == "tiles" and 1 <= len(cams) <= self.batch and self.frame % self.batch == 0
        cams = list(cams) + [cams[-1]] * (self.batch - len(cams))  # a short last group is padded: every slot gets its arrivals
        f = self.frame
        self.frame += self.batch
        self.ctx.sortfirst_render_tiles_batch(cams, self.tiles, f)
        return f

    def render_tiles_batch(self, cams, present: bool = False) -> int:
        """submit_tiles_batch + (root) wait + consume of every frame of the group, in order. Asynchronous."""
        f = self.submit_tiles_batch(cams)
        if self.rank == 0:
            for k in range(self.batch):
                self.wait(f + k)
                self.consume(f + k, present)
        return f

    def render_batch(self, cams, present: bool = False, flush_l2: bool = False) -> int:
        """submit_batch + (root) wait + consume of every frame of the group, in order. Asynchronous.
        flush_l2: write an L2-sized buffer before the owner's launch (benchmarks)."""
        f = self.submit_batch(cams, flush_l2)
        if self.rank == 0:
            for k in range(self.batch):
                self.wait(f + k)
                self.consume(f + k, present)
        return f

    def wait(self, f: int):
        """Root only: enqueue the device-side wait for frame f; afterwards readback/present see frame f."""
        assert self.rank == 0 and f == self.consumed, "frames are consumed in order"
        self.ctx.sortfirst_wait(f, self.arrivals_target(f))

    def consume(self, f: int, present: bool = False):
        assert self.rank == 0 and f == self.consumed
        self.ctx.sortfirst_consume(f, present)
        self.consumed += 1

    def render(self, cam, present: bool = False) -> int:
        """submit + (root) wait + consume. Asynchronous on every rank."""
        f = self.submit(cam)
        if self.rank == 0:
            self.wait(f)
            self.consume(f, present)
        return f

    def close(self):
        self.ctx.sync()
        if self.world > 1:
            self.dist.barrier()
        self.ctx.sortfirst_leave()
        if self.world > 1:
            self.dist.barrier()

    # -- measurement helpers used by bench.py -----------------------------------------------------
    def my_frames(self, first: int, count: int) -> int:
        if self.granularity == "tiles":
            return count
        return sum(1 for f in range(first, first + count) if frame_owner(f, self.world, self.batch) == self.rank)

    def my_launches(self, first: int, count: int) -> int:
        """Launches this rank issues for frames [first, first+count) ('frames' granularity, groups of `batch`)."""
        return sum(1 for f in range(first, first + count, self.batch) if frame_owner(f, self.world, self.batch) == self.rank)

    def e2e(self, cams, K: int, warmup: int) -> dict:
        """End to end on the root, pipelined: every rank submits its share; the root waits for each frame
        in order, presents it and copies the RGBA8 image into host memory (blocking D2H per frame). Wall
        clock on the root from the first submit to the last frame's pixels on the host. The L2 flush
        before each rendered frame is INSIDE the timed region here (conservative)."""
        ctx, dist = self.ctx, self.dist
        R = 8  # host frames in flight on the root: the D2H copies are enqueued, the host blocks once per R frames
        pinned = rt.PinnedArray((R, ctx.height, ctx.width, 4), np.uint8) if self.rank == 0 else None
        out = pinned.array if pinned is not None else None
        done = 0
        ctx.sync()
        dist.barrier()
        t0 = time.perf_counter()
        B = self.batch
        for i in range(0, K, B):
            f_next = self.frame
            if B == 1 and (self.granularity == "tiles" or frame_owner(f_next, self.world, B) == self.rank):
                ctx.flush_l2()
            if B > 1:
                f = self.submit_batch([cams[(warmup + i + k) % len(cams)] for k in range(min(B, K - i))], flush_l2=True)
            else:
                f = self.submit(cams[(warmup + i) % len(cams)])
            if self.rank == 0:
                for k in range(B):
                    self.wait(f + k)
                    ctx.present()
                    ctx.readback_rgba8_async(out[done % R])
                    self.consume(f + k)
                    done += 1
                    if done % R == 0:
                        ctx.sync()  # a consumer would use the R host frames here
        ctx.sync()
        tot = time.perf_counter() - t0
        dist.barrier()
        fps = K / tot if self.rank == 0 and tot > 0 else 0.0
        return {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": (144 + 48) * self.participants(0),
                "d2h_bytes_per_step": ctx.width * ctx.height * 4,
                "how": f"pipelined sort-first ({self.granularity}): ranks render into rank 0's ring by peer stores; rank 0 waits for each frame "
                       "in order, presents, enqueues the RGBA8 D2H into a ring of 8 page-locked host frames and blocks once per 8 frames; wall clock on rank 0 "
                       "over all K frames, L2 flushes included. Every frame leaves through rank 0's PCIe link (~0.154 ms per 1080p RGBA8 frame)"}

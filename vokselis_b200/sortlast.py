"""Sort-last rendering of a volume that exceeds one GPU (BASELINE config 5): the global grid is cut into
px x py x pz axis-aligned bricks, one per rank (one process per GPU). Host-side logic only; the kernels
and the per-rank API are vkrt_partial_* (include/vokselis_rt.h, vokselis_b200/csrc/sortlast.cu).

Per frame, scheme "two-pass": every rank marches its brick alpha-only -> all-gather of the
per-pixel transmittances (NCCL) -> each rank derives the alpha entering its brick from the bricks in
front of it -> colour pass with exact early termination -> the premultiplied partials are SUMMED onto
rank 0 (NCCL reduce; with absolute weights the composite is commutative) -> rank 0 finalizes the frame.
Scheme "deferred" (default) marches ONCE from alpha 0, then scales each pixel by the transmittance in
front of the brick and re-marches only the pixels whose 0.95 crossing can fall inside the brick.

The reference has no counterpart: it is single-device. Its compositing operator,
shaders/raycast_compute.wgsl:88-91, is what the partial passes split.
"""
from __future__ import annotations

import math

import numpy as np


def brick_grid(world: int) -> tuple[int, int, int]:
    """2 -> 2x1x1, 4 -> 2x2x1, 8 -> 2x2x2, ... (powers of two, split x then y then z round-robin)."""
    g = [1, 1, 1]
    i = 0
    w = world
    while w > 1:
        if w % 2:
            raise ValueError("world size must be a power of two")
        g[i % 3] *= 2
        w //= 2
        i += 1
    return tuple(g)


def brick_range(gn, grid, rank: int):
    """Owned voxel range [lo, hi) of `rank`; interior bounds are multiples of 8."""
    px, py, pz = grid
    b = (rank % px, (rank // px) % py, rank // (px * py))
    lo, hi = [], []
    for n, parts, k in zip(gn, grid, b):
        cells = -(-n // 8)
        c0, c1 = (cells * k) // parts, (cells * (k + 1)) // parts
        lo.append(c0 * 8)
        hi.append(min(c1 * 8, n))
    return tuple(lo), tuple(hi)


def visibility_order(eye, gn, grid) -> list[int]:
    """Ranks front to back for rays leaving `eye` (world coordinates; the volume is the box [-1,1]^3).
    Along any ray |coordinate - eye| grows monotonically on every axis, so per axis the bricks are met
    in order of their distance from the eye's brick; two bricks that are ordered differently on two
    axes are never crossed by the same ray. Any linear extension of the component-wise order is
    therefore valid for ALL pixels: sort by the sum of per-axis nearness ranks."""
    px, py, pz = grid
    keys = []
    for r in range(px * py * pz):
        b = (r % px, (r // px) % py, r // (px * py))
        s = 0
        for axis in range(3):
            n, parts = gn[axis], grid[axis]
            # brick index the eye falls in (clamped): voxel coordinate of the eye on this axis
            q = (eye[axis] + 1.0) * n / 2.0
            cells = -(-n // 8)
            bounds = [((cells * k) // parts) * 8 for k in range(parts)] + [n]
            e = sum(1 for k in range(1, parts) if q >= bounds[k])
            s += abs(b[axis] - e)
        keys.append((s, r))
    return [r for _, r in sorted(keys)]


class SortLastGroup:
    """Drives one frame across the ranks. `torch` tensors hold the exchange buffers; collectives run on
    the context's own stream (torch.cuda.ExternalStream), so no host synchronisation is needed."""

    def __init__(self, ctx, rank: int, world: int, gn, dist=None, grid=None, scheme: str = "deferred", exchange: str = "p2p"):
        import torch

        if dist is None:
            import torch.distributed as dist  # noqa: PLC0415
        self.torch, self.dist = torch, dist
        if scheme not in ("deferred", "two-pass"):
            raise ValueError(scheme)
        self.scheme = scheme
        self.ctx, self.rank, self.world, self.gn = ctx, rank, world, tuple(gn)
        self.grid = grid or brick_grid(world)
        self.own_lo, self.own_hi = brick_range(self.gn, self.grid, rank)
        n = ctx.width * ctx.height
        dev = torch.device("cuda", ctx.device)
        self.stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
        self.T = torch.empty(n, dtype=torch.float32, device=dev)
        self.T_all = None  # (nccl transport: allocated below)
        self.ain = torch.empty(n, dtype=torch.float32, device=dev)
        self.rgba = torch.empty(n * 4, dtype=torch.float32, device=dev)
        # Transmittance exchange. "p2p" (default): direct-send over NVLink peer memory, vkrt_exchange_* — every rank stores its
        # image straight into the tables of the ranks BEHIND it (the only ones whose resolve reads it), one kernel, no
        # collective; torch.distributed only ships the IPC handles. "nccl": all-gather of every image to every rank.
        if exchange not in ("p2p", "nccl"):
            raise ValueError(exchange)
        self.exchange = exchange if world > 1 else "nccl"
        self.frame = 0
        if self.exchange == "p2p":
            mine = ctx.exchange_create(rank, world)
            handles = [None] * world
            dist.all_gather_object(handles, mine)
            ctx.exchange_open(handles)
            dist.barrier()
        else:
            self.T_all = torch.empty(world * n, dtype=torch.float32, device=dev)

    PHASES = ("march", "all_gather", "resolve", "remarch", "reduce", "finalize")  # "all_gather" = the transmittance exchange, whichever transport

    def close(self):
        """Release the peer mappings (all ranks; synchronises)."""
        if self.exchange == "p2p":
            self.ctx.sync()
            self.dist.barrier()
            self.ctx.exchange_close()
            self.dist.barrier()
            self.exchange = "closed"

    @staticmethod
    def eye_of(cam):
        """The point every ray of `cam` passes through, derived from the matrix the kernels actually use: the camera
        centre is the pre-image of the clip-space direction (0, 0, 1, 0), i.e. column 2 of inv_proj, dehomogenised
        (raycast_compute.wgsl:107-116 builds rays from inv_proj only; view_position is not read by the shader). Falls
        back to view_position for matrices without a finite centre (orthographic / identity)."""
        m = [float(v) for v in cam.inv_proj[:]]  # column-major 4x4
        w = m[11]
        if abs(w) > 1e-12 and all(math.isfinite(v) for v in m[8:12]):
            return (m[8] / w, m[9] / w, m[10] / w)
        return tuple(float(v) for v in cam.view_position[:3])

    def render(self, cam, timed: bool = False):
        """All ranks call this; rank 0's context frame holds the result afterwards. Asynchronous. timed: record a
        CUDA event on the context's stream between the phases (read with phase_ms() after a sync)."""
        torch, dist, ctx = self.torch, self.dist, self.ctx
        order = visibility_order(self.eye_of(cam), self.gn, self.grid)
        before = order[: order.index(self.rank)]
        mark = ctx.mark if timed else (lambda i: None)
        with torch.cuda.stream(self.stream):
            mark(0)
            if self.scheme == "two-pass":  # alpha-only march, then the colour march with the exact incoming alpha
                ctx.partial_alpha(cam, self.T.data_ptr())
            else:  # ONE march from alpha 0 (relative partial); early termination is resolved afterwards
                ctx.partial_relative(cam, self.rgba.data_ptr(), self.T.data_ptr())
            mark(1)
            f = self.frame
            self.frame += 1
            if self.exchange == "p2p":
                ctx.exchange_push(self.T.data_ptr(), order[order.index(self.rank) + 1:], f)
                ctx.exchange_wait(f, len(before))
                t_all = ctx.exchange_table(f)
            elif self.world > 1:
                dist.all_gather_into_tensor(self.T_all, self.T)
                t_all = self.T_all.data_ptr()
            else:
                self.T_all.copy_(self.T)
                t_all = self.T_all.data_ptr()
            mark(2)
            if self.scheme == "two-pass":
                ctx.partial_ain(t_all, before, self.ain.data_ptr())
            else:
                ctx.partial_resolve(t_all, before, self.rgba.data_ptr(), self.ain.data_ptr())
            if self.exchange == "p2p":
                ctx.exchange_done(f)
            mark(3)
            ctx.partial_color(cam, self.ain.data_ptr(), self.rgba.data_ptr())  # deferred: only the flagged pixels
            mark(4)
            if self.world > 1:
                dist.reduce(self.rgba, dst=0, op=dist.ReduceOp.SUM)
            mark(5)
            if self.rank == 0:
                ctx.partial_finalize(cam, self.rgba.data_ptr())
            mark(6)

    def phase_ms(self) -> dict:
        """Device time of each phase of the last render(timed=True) on this rank's stream."""
        return {name: self.ctx.mark_elapsed(i, i + 1) for i, name in enumerate(self.PHASES)}

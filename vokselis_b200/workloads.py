"""BASELINE.json's sharded configurations as callable workloads (harness code shared by bench.py, the multi-GPU
checks under tests/ and the development scripts under bench/). Nothing here is on the product's data path: it
drives the C ABI (vokselis_b200.rt), the sort-first / sort-last groups and CUDA-event timing.

  configs[2]  synthetic 1024^3 fp16 noise volume at 3840x2160, image tiles sharded sort-first, gathered to rank 0
  configs[3]  synthetic 2048^3 uint8 volume, 90 % empty space, 4K, exact skipping + early termination, same sharding
  configs[4]  synthetic 4096^3 fp32 volume (256 GiB), brick-partitioned sort-last at 2/4/8 GPUs, 4K

All three run mode M1 with per-voxel steps (dt_floor 0, dt_scale 2: the box side is 2, so one voxel per step on the
dominant axis, like raycast_naive.wgsl:97-99). Volumes are generated on the device (vkrt_generate_synthetic*).
These volumes exceed the 126 MB L2 many times over, so no L2 flush is needed between timed frames.
"""
from __future__ import annotations

import math
import time

import numpy as np

from . import abi, rt

SORTFIRST_CONFIGS = {
    # layout: measured on one B200 (bench/sharded.py --layout 3|4): the dense fp16 fog is bound by the texture path and gains 19 %
    # from the pre-gathered quads (3.29 -> 2.77 ms per 4K frame, 8 GiB of quads); the sparse volume is bound by traversal and is
    # indifferent (0.691 vs 0.696 ms), so it keeps the layered array (8 GiB instead of 32 GiB)
    3: dict(name="configs[2]: synthetic 1024^3 fp16 noise volume at 3840x2160, image-tile sharded sort-first", kind=0, dtype=np.float16, n=1024, seed=3,
            layout=abi.LAYOUT_QUAD),
    4: dict(name="configs[3]: synthetic 2048^3 uint8 bricked volume with 90% empty space, skipping + early termination, 4K, image-tile sharded sort-first",
            kind=1, dtype=np.uint8, n=2048, seed=4, layout=abi.LAYOUT_GATHER),
}


def _cams(frames, W, H):
    return [rt.Camera(3.0, -0.5, 1.0 + 2 * math.pi * i / max(frames, 1), (0, 0, 0), W / H).get_proj_view_matrix() for i in range(frames)]


def _max_over_ranks(value: float, dist, world: int) -> float:
    if world <= 1:
        return value
    import torch

    t = torch.tensor([value], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _gather_floats(value: float, dist, world: int) -> list[float]:
    if world <= 1:
        return [value]
    import torch

    t = torch.tensor([value], dtype=torch.float64, device="cuda")
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [float(x.item()) for x in out]


def run_sortfirst_tiles(cid: int, rank: int, world: int, local: int, dist=None, frames: int = 24, warm: int = 4, tile: int = 120,
                        layout: int | None = None, res=(3840, 2160), edge: int | None = None, checks: bool = True, hbm_peak_gbs: float = 6650.0,
                        batch: int = 1, brick: int = 0):
    """configs[2] / configs[3] on `world` GPUs (one process each). N = 1: `single` on one GPU. N > 1: every frame is cut
    into tile x tile pixel tiles dealt over the ranks (volume replicated), each rank renders its tiles locally and ships
    them into rank 0's frame over NVLink; rank 0 waits for each frame in order. Timing: CUDA events on RANK 0's stream
    around the WHOLE pipelined sequence of `frames` frames (first wait to last consume) — i.e. frames gathered on rank 0
    per second, everything included — plus every rank's summed raycast-kernel time. Returns a dict on rank 0."""
    from . import sortfirst

    cfg = SORTFIRST_CONFIGS[cid]
    W, H = res
    n = edge or cfg["n"]
    layout = cfg["layout"] if layout is None else layout
    ctx = rt.Context(local, W, H)
    t0 = time.perf_counter()
    ctx.set_occupancy_brick(brick)  # 0 = the library's choice
    ctx.generate_synthetic(cfg["kind"], cfg["dtype"], n, seed=cfg["seed"])
    ctx.sync()
    gen_s = time.perf_counter() - t0
    p = rt.default_params(abi.MODE_M1)
    p.dt_scale, p.dt_floor, p.skip_empty, p.layout = 2.0, 0.0, 1, layout
    ctx.set_params(p)
    cams = _cams(frames, W, H)
    out_checks, st, ref_frame, info = {}, None, None, None
    if rank == 0:
        info = ctx.volume_info()
        q = rt.default_params(abi.MODE_M1)
        q.dt_scale, q.dt_floor, q.skip_empty, q.layout, q.count_samples = 2.0, 0.0, 1, layout, 1
        ctx.set_params(q)
        ctx.reset_stats()
        ctx.render(cams[0])
        st = ctx.stats()
        ref_frame = ctx.readback()
        if checks:
            aux1 = ctx.readback_aux()
            q.skip_empty = 0
            ctx.set_params(q)
            ctx.render(cams[0])
            out_checks["skip_on_equals_skip_off"] = bool(np.array_equal(ref_frame, ctx.readback()) and np.array_equal(aux1, ctx.readback_aux()))
            q.skip_empty, q.count_samples = 1, 0
            ctx.set_params(q)
            ctx.resize(W, H)
            ctx.render_tiles(cams[0], rt.tile_table(W, H, 256))
            out_checks["tile_equals_single"] = bool(np.array_equal(ref_frame, ctx.readback()))
        ctx.set_params(p)
    # N = 1 runs the SAME pipeline (ring of frames in rank 0's memory, tiles of consecutive frames on rotating streams)
    # every rank renders ITS tiles of `batch` consecutive frames in one launch (grid.z = frame x tile); 8 such launches in flight.
    # batch = 1 is the default: measured at 8 GPUs on config 4, 48 frames: 8,287 frames/s with 1 frame per launch, 6,069 with 4,
    # 5,973 with 8 — eight concurrent launches already hide the short launches' tails, and at 8,287 frames/s rank 0 receives
    # 7/8 x 66 MB per frame = 480 GB/s over NVLink in 960-byte tile rows: the gather itself is the bound, not the kernels.
    G = max(1, min(batch, 15))
    frames = -(-frames // G) * G
    cams = _cams(frames, W, H)
    group = sortfirst.SortFirstGroup(ctx, rank, world, granularity="tiles", tile=tile, slots=8 * G, dist=dist, batch=G)
    f = group.submit_tiles_batch(cams[:G])
    if rank == 0:
        for k in range(G):
            group.wait(f + k)
            if k == 0:
                out_checks["sortfirst_equals_single" if world == 1 else "sortfirst_equals_single_gpu"] = bool(np.array_equal(ref_frame, ctx.readback()))
            group.consume(f + k)
    ctx.timing_enable(frames // G)

    def render(i0):
        group.render_tiles_batch(cams[i0:i0 + G])

    for i0 in range(0, max(warm, G), G):
        render(i0 % frames)
    ctx.sync()
    if dist is not None and world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    ctx.mark(0)
    for i0 in range(0, frames, G):
        render(i0)
    ctx.mark(1)
    ctx.sync()
    wall = time.perf_counter() - t0
    pipeline_ms = ctx.mark_elapsed(0, 1) if rank == 0 else 0.0  # rank 0's stream: waits for every frame, in order
    if dist is not None and world > 1:
        dist.barrier()
    kernel_ms = float(ctx.timing_read(frames // G).astype(np.float64).sum())
    per_rank = _gather_floats(kernel_ms, dist, world)
    timeouts = ctx.sortfirst_timeouts() if rank == 0 else 0
    group.close()
    ctx.close()
    if rank != 0:
        return None
    ms = pipeline_ms / frames
    eb = np.dtype(cfg["dtype"]).itemsize
    vol_bytes = n ** 3 * eb
    alg = st.samples_fetched * 8 * eb
    hbm_bytes = min(vol_bytes, alg) + W * H * 8
    busiest = max(per_rank) / frames
    return {
        "config": cfg["name"], "volume_edge": n, "dtype": np.dtype(cfg["dtype"]).name, "resolution": [W, H], "n_gpus": world,
        "sharding": (f"one GPU: all {tile}-pixel tiles of a frame in one launch" if world == 1 else
                     f"sort-first, {tile}-pixel image tiles dealt over {world} ranks, tiles shipped to rank 0's frame over NVLink (P2P stores of whole tile rows)")
                    + f"; a launch renders a rank's tiles of {G} consecutive frames (grid.z = frame x tile), consecutive launches rotate over 8 render streams on every rank "
                      f"(a ring of {8 * G} frames on rank 0)",
        "layout": layout, "frames": frames, "frames_per_launch": G, "frames_per_s": 1e3 / ms, "ms_per_frame": ms,
        "timing": "CUDA events on rank 0's stream around the whole pipelined sequence (rank 0 waits for every frame's tiles in order); volume >> L2, no flush",
        "wall_ms_per_frame": 1e3 * wall / frames,
        "kernel_ms_per_frame_by_rank": [v / frames for v in per_rank], "busiest_rank_kernel_ms_per_frame": busiest,
        "kernel_ms_note": "sum of each launch's own CUDA-event duration; launches of consecutive frames overlap (four streams), so these exceed the pipelined time",
        "ray_samples_per_s": st.samples_reference * 1e3 / ms, "fetched_samples_per_s": st.samples_fetched * 1e3 / ms,
        "samples_frame0": {"reference": int(st.samples_reference), "fetched": int(st.samples_fetched), "rays": int(st.rays_hit)},
        "bricks": {"total": info["bricks_total"], "occupied": info["bricks_occupied"]}, "volume_bytes": vol_bytes, "generate_s": gen_s,
        "roofline": {"bound": "hbm", "achieved": hbm_bytes / (ms * 1e-3) / 1e9, "peak": hbm_peak_gbs * world, "unit": "GB/s",
                     "frac": hbm_bytes / (ms * 1e-3) / 1e9 / (hbm_peak_gbs * world), "compulsory_bytes_per_frame": hbm_bytes,
                     "note": "whole job: compulsory bytes per frame = min(volume bytes, fetched samples x 8 taps x sizeof) + W*H*8, over the pipelined time per "
                             "frame, against N x the measured HBM peak (every rank reads its share of the volume)"},
        "texel_GBs": alg / (ms * 1e-3) / 1e9, "checks": out_checks, "sortfirst_wait_timeouts": int(timeouts),
    }


def run_sortlast(rank: int, world: int, local: int, dist, edge: int = 4096, frames: int = 6, res=(3840, 2160), kind: int = 3,
                 dtype=np.float32, save_frame0: str | None = None, hbm_peak_gbs: float = 6650.0, exchange: str = "p2p"):
    """configs[4]: edge^3 fp32 volume brick-partitioned over `world` ranks (sort-last). Every phase of a frame is timed
    with CUDA events on each rank's stream; the per-frame time is the max over ranks of the whole frame."""
    import torch

    from . import sortlast

    W, H = res
    gn = (edge, edge, edge)
    dtype = np.dtype(dtype)
    ctx = rt.Context(local, W, H)
    group = sortlast.SortLastGroup(ctx, rank, world, gn, dist=dist, exchange=exchange)
    t0 = time.perf_counter()
    ctx.generate_synthetic_window(kind, dtype, gn, group.own_lo, group.own_hi, seed=5)
    ctx.sync()
    gen_s = time.perf_counter() - t0
    p = rt.default_params(abi.MODE_M1)
    p.dt_scale, p.dt_floor, p.skip_empty = 2.0, 0.0, 1
    ctx.set_params(p)
    cams = _cams(frames, W, H)
    group.render(cams[0])
    ctx.sync()
    if rank == 0 and save_frame0:
        ctx.present()
        np.save(save_frame0, ctx.readback_rgba8()[::4, ::4].copy())
    if world > 1:
        dist.barrier()
    per_frame, phases = [], []
    for cam in cams:
        ctx.flush_l2()
        group.render(cam, timed=True)
        ctx.sync()
        ph = group.phase_ms()
        per_frame.append(_max_over_ranks(sum(ph.values()), dist, world))
        phases.append(ph)
    mean_phase = {k: float(np.mean([ph[k] for ph in phases])) for k in phases[0]}
    gathered = {k: _gather_floats(v, dist, world) for k, v in mean_phase.items()}
    # per frame, per rank: [frame][rank] (which rank waits for which changes with the camera, so transfer cost and
    # imbalance must be separated frame by frame, not on the means)
    per_frame_rank = {k: np.array([_gather_floats(ph[k], dist, world) for ph in phases]) for k in ("march", "all_gather", "reduce")}
    exchange = group.exchange
    timeouts = _max_over_ranks(float(ctx.exchange_timeouts()) if exchange == "p2p" else 0.0, dist, world)
    group.close()
    ctx.close()
    if rank != 0:
        return None
    mean_ms = float(np.mean(per_frame))
    eb = dtype.itemsize
    per_rank_bytes = edge ** 3 * eb / world
    march = max(gathered["march"])
    return {
        "config": f"configs[4]: synthetic {edge}^3 {dtype.name} volume ({edge ** 3 * eb / 2 ** 30:.0f} GiB) brick-partitioned sort-last over {world} GPUs, {W}x{H}",
        "n_gpus": world, "brick_grid": list(group.grid), "volume_bytes": edge ** 3 * eb, "bytes_per_rank": per_rank_bytes,
        "frames": frames, "ms_per_frame": mean_ms, "frames_per_s": 1e3 / mean_ms, "per_frame_ms": per_frame, "generate_s_rank0": gen_s,
        "phase_ms_mean_max_over_ranks": {k: max(v) for k, v in gathered.items()}, "phase_ms_mean_by_rank": gathered,
        "exchange": ("direct-send over NVLink peer memory (vkrt_exchange_*): every rank stores its transmittance image into the tables of the ranks behind it in "
                     "visibility order; NCCL only sums the partial colours onto rank 0" if exchange == "p2p" else "NCCL all-gather of the transmittance images + NCCL sum"),
        "exchange_wait_timeouts": int(timeouts),
        "exchange_bytes_per_rank_per_frame": {"transmittance_sent_mean": (world - 1) / 2 * W * H * 4 if exchange == "p2p" else W * H * 4,
                                              "transmittance_received_max": (world - 1) * W * H * 4, "reduce_rgba": W * H * 16},
        # A collective is also where a rank that finished its march early WAITS for the slowest one: on the rank that arrives
        # last the phase lasts as long as the transfer itself, on the others transfer + wait. The transfer cost is therefore
        # the MIN over ranks, the imbalance of the march is max - min of the march phase.
        "exchange_ms": {"transmittance_transfer": float(per_frame_rank["all_gather"].min(axis=1).mean()),
                        "reduce_transfer": float(per_frame_rank["reduce"].min(axis=1).mean()),
                        "transmittance_incl_wait_for_slowest_march": float(per_frame_rank["all_gather"].max(axis=1).mean()),
                        "march_imbalance_max_minus_min": float((per_frame_rank["march"].max(axis=1) - per_frame_rank["march"].min(axis=1)).mean()),
                        "note": "per frame: min over ranks = the rank that arrived last (no waiting), max = transfer + wait for the slowest march; means over frames"},
        "exchange_share_of_frame": float(per_frame_rank["all_gather"].min(axis=1).mean() + per_frame_rank["reduce"].min(axis=1).mean()) / mean_ms,
        "timing": "CUDA events on every rank's own stream between the phases of a frame (march from alpha 0, exchange of transmittances [phase 'all_gather'], resolve, re-march of "
                  "the flagged pixels, NCCL sum onto rank 0, finalize); ms_per_frame = max over ranks of the whole frame, mean over frames; L2 flushed before each frame",
        "roofline": {"bound": "hbm", "achieved": per_rank_bytes / (march * 1e-3) / 1e9, "peak": hbm_peak_gbs, "unit": "GB/s",
                     "frac": per_rank_bytes / (march * 1e-3) / 1e9 / hbm_peak_gbs,
                     "note": "per rank: its brick (read once per frame at one step per voxel) / its march time (slowest rank)"},
    }


# ---- multi-GPU correctness checks (the bodies of tests/mgpu_*_check.py; also run by bench.py before timing) -----------------
def check_sortfirst(rank: int, world: int, local: int, dist, log=None) -> dict:
    """Sort-first frames — tiles on all ranks shipped into rank 0's frame, whole frames dealt round-robin, batched groups —
    must equal the single-GPU frame bit for bit, with no device-side wait timing out. First renders use a non-LINEAR
    layout built lazily (the layout/stream ordering of a root batch)."""
    import torch

    from . import sortfirst, volumes

    say = log or (lambda *_: None)
    W, H = 1280, 720
    ok, detail = True, {}
    for mode in (abi.MODE_M0, abi.MODE_M1):
        ctx = rt.Context(local, W, H)
        if mode == abi.MODE_M0:
            ctx.generate_xor(128, 0)
        else:
            ctx.upload_scalar(volumes.bonsai_standin_u8(128, seed=2, blobs=10))
        p = rt.default_params(mode)
        p.skip_empty = 1
        p.layout = abi.LAYOUT_BRICKED if mode == abi.MODE_M0 else abi.LAYOUT_QUAD
        ctx.set_params(p)
        cams = [rt.Camera(2.6, -0.4, 0.5 + 0.7 * i, (0, 0, 0), W / H).get_proj_view_matrix() for i in range(5)]
        # the FIRST render after set_params is a root batch on the second stream: the lazily built layout must be ordered before it
        group = sortfirst.SortFirstGroup(ctx, rank, world, granularity="frames", batch=2, dist=dist)
        first = []
        f = group.submit_batch(cams[:2])
        if rank == 0:
            for k in range(2):
                group.wait(f + k)
                first.append(ctx.readback())
                group.consume(f + k)
        group.close()
        refs = []
        if rank == 0:
            for cam in cams:
                ctx.render(cam)
                refs.append(ctx.readback())
            same = all(np.array_equal(a, b) for a, b in zip(first, refs[:2]))
            ok = ok and same
            detail[f"mode{mode}_first_render_is_root_batch"] = bool(same)
        for gran, slots in (("tiles", 2), ("tiles", 4), ("frames", 2), ("frames", None)):
            group = sortfirst.SortFirstGroup(ctx, rank, world, granularity=gran, tile=96, slots=slots, dist=dist)
            seq = cams * 3  # 15 frames back to back: exercises slot reuse and the consumed flag
            good = True
            for i, cam in enumerate(seq):
                f = group.submit(cam)
                if rank == 0:
                    group.wait(f)
                    got = ctx.readback()
                    group.consume(f)
                    same = np.array_equal(got, refs[i % len(cams)])
                    good = good and same
                    if not same:
                        say(f"mode {mode} {gran}/{slots} frame {i}: MISMATCH ({(got != refs[i % len(cams)]).sum()} differing halfs)")
            if rank == 0:
                to = ctx.sortfirst_timeouts()
                good = good and to == 0
                ok = ok and good
                detail[f"mode{mode}_{gran}_slots{slots}"] = bool(good)
                say(f"mode {mode} {gran}/{slots}: {'bit-exact' if good else 'FAILED'}, device-side wait timeouts = {to}")
            group.close()
        for batch, slots, total in ((3, None, 20), (5, 80, 103)):
            group = sortfirst.SortFirstGroup(ctx, rank, world, granularity="frames", batch=batch, slots=slots, dist=dist)
            seq = (cams * (total // len(cams) + 1))[:total]
            done, good = 0, True
            for g in range(0, len(seq), batch):
                chunk = seq[g:g + batch]
                f = group.submit_batch(chunk, flush_l2=(g % 2 == 0))
                if rank == 0:
                    for k in range(batch):
                        group.wait(f + k)
                        got = ctx.readback()
                        group.consume(f + k)
                        if k < len(chunk):
                            same = np.array_equal(got, refs[(g + k) % len(cams)])
                            good = good and same
                            done += 1
            if rank == 0:
                to = ctx.sortfirst_timeouts()
                good = good and to == 0
                ok = ok and good
                detail[f"mode{mode}_frames_batch{batch}_slots{group.slots}"] = bool(good)
                say(f"mode {mode} frames/batch{batch}/slots{group.slots}: {done} frames {'bit-exact' if good else 'FAILED'}, timeouts = {to}")
            group.close()
        ctx.close()
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    return {"ok": bool(flag.item() == 0), "what": "sort-first frames == single-GPU frames bit for bit (tiles, frames, batched groups; M0 and M1), 0 wait timeouts",
            "detail": detail}


def check_sortlast(rank: int, world: int, local: int, dist, log=None) -> dict:
    """Sort-last frames (one brick per rank, NCCL all-gather of transmittances + NCCL sum of partials) vs the single-GPU
    frame of the same 256^3 fp32 volume on rank 0: max |delta| <= 2/255 after present, eye outside / oblique / inside."""
    import torch

    from . import sortlast

    say = log or (lambda *_: None)
    W, H, n = 640, 360, 256
    gn = (n, n, n)
    ok, deltas = True, []
    ctx = rt.Context(local, W, H)
    p = rt.default_params(abi.MODE_M1)
    p.dt_scale, p.skip_empty = 2.0, 1
    cams = [rt.Camera(z, pt, y, (0, 0, 0), W / H).get_proj_view_matrix() for z, pt, y in [(3.0, -0.5, 1.0), (2.0, 0.6, -2.3), (0.7, 0.1, 0.4)]]
    refs = []
    if rank == 0:
        ctx.generate_synthetic(2, np.float32, n, seed=5)
        ctx.set_params(p)
        for cam in cams:
            ctx.render(cam)
            ctx.present()
            refs.append(ctx.readback_rgba8())
    group = sortlast.SortLastGroup(ctx, rank, world, gn, dist=dist)
    ctx.generate_synthetic_window(2, np.float32, gn, group.own_lo, group.own_hi, seed=5)
    ctx.set_params(p)
    for i, cam in enumerate(cams):
        group.render(cam)
        ctx.sync()
        if rank == 0:
            ctx.present()
            got = ctx.readback_rgba8()
            d = int(np.abs(got.astype(np.int32) - refs[i].astype(np.int32)).max())
            deltas.append(d)
            ok = ok and d <= 2
            say(f"sort-last world {world} cam {i}: max |delta| {d}/255 {'ok' if d <= 2 else 'MISMATCH'}")
    timeouts = ctx.exchange_timeouts() if group.exchange == "p2p" else 0
    flag = torch.tensor([0 if (ok and timeouts == 0) else 1], device="cuda")
    dist.all_reduce(flag)
    group.close()
    ctx.close()
    return {"ok": bool(flag.item() == 0), "exchange": "p2p direct-send (vkrt_exchange_*), 0 wait timeouts" if timeouts == 0 else f"{timeouts} wait timeouts",
            "what": "sort-last frame vs single-GPU frame of the same 256^3 fp32 volume, max |delta| <= 2/255 after present (3 cameras)",
            "max_delta_255": deltas}

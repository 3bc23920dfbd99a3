"""BASELINE.json config 5: synthetic 4096^3 fp32 volume (256 GiB, exceeds one GPU), brick-partitioned
sort-last across N in {2,4,8} B200 with visibility-ordered compositing, 3840x2160. Run under torchrun.

Mode M1, per-voxel steps (dt_floor 0, dt_scale 2), thin-fog volume (kind 3) generated per rank on the
device for its own brick + halo. Timing: CUDA events on each rank's stream around the whole frame
(alpha pass, all-gather, alpha-in, colour pass, reduce, finalize), max over ranks, mean over frames.
Prints one JSON line; writes a 4x-downsampled RGBA8 copy of frame 0 to gpurun_out/ so that runs with
different N can be compared offline (they must agree within the parity tolerance).
usage: config5.py [--edge 4096] [--frames 6] [--res 3840x2160]
"""
import argparse
import json
import math
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from vokselis_b200 import abi, rt, sortlast  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--edge", dest="n", type=int, default=4096)
    ap.add_argument("--frames", type=int, default=6)
    ap.add_argument("--res", default="3840x2160")
    ap.add_argument("--kind", type=int, default=3)
    ap.add_argument("--dtype", default="float32")
    args = ap.parse_args()
    W, H = map(int, args.res.split("x"))
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.n
    gn = (n, n, n)
    dtype = np.dtype(args.dtype)
    ctx = rt.Context(local, W, H)
    group = sortlast.SortLastGroup(ctx, rank, world, gn, dist=dist)
    t0 = time.perf_counter()
    ctx.generate_synthetic_window(args.kind, dtype, gn, group.own_lo, group.own_hi, seed=5)
    ctx.sync()
    gen_s = time.perf_counter() - t0
    p = rt.default_params(abi.MODE_M1)
    p.dt_scale, p.dt_floor, p.skip_empty = 2.0, 0.0, 1
    ctx.set_params(p)
    cams = [rt.Camera(3.0, -0.5, 1.0 + 2 * math.pi * i / max(args.frames, 1), (0, 0, 0), W / H).get_proj_view_matrix() for i in range(args.frames)]
    # frame 0 once (also the cross-N comparison artefact), then timed frames
    group.render(cams[0])
    ctx.sync()
    if rank == 0:
        ctx.present()
        img = ctx.readback_rgba8()
        os.makedirs(ROOT / "gpurun_out", exist_ok=True)
        np.save(ROOT / "gpurun_out" / f"config5_n{n}_w{world}_frame0_ds4.npy", img[::4, ::4].copy())
    if world > 1:
        dist.barrier()
    per_frame = []
    for cam in cams:
        ctx.flush_l2()
        ctx.mark(0)
        group.render(cam)
        ctx.mark(1)
        ctx.sync()
        ms = torch.tensor([ctx.mark_elapsed(0, 1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        per_frame.append(float(ms.item()))
    if rank == 0:
        mean_ms = float(np.mean(per_frame))
        eb = dtype.itemsize
        line = {
            "config": "config 5: %d^3 %s, sort-last over %d GPUs, %dx%d" % (n, dtype.name, world, W, H), "n_gpus": world,
            "brick_grid": list(group.grid), "volume_bytes": n ** 3 * eb, "bytes_per_rank": n ** 3 * eb // world,
            "frames": args.frames, "ms_per_frame": mean_ms, "frames_per_s": 1e3 / mean_ms, "per_frame_ms": per_frame,
            "generate_s_rank0": gen_s,
            "exchange_bytes_per_rank_per_frame": {"all_gather_T": (world - 1) * W * H * 4, "reduce_rgba": W * H * 16},
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

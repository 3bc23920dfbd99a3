"""BASELINE.json configs 3 and 4 (the sort-first 4K cases) — run alone (1 GPU) or under torchrun (N GPUs).

  config 3: synthetic 1024^3 fp16 noise volume at 3840x2160, image tiles sharded sort-first
  config 4: synthetic 2048^3 uint8 volume with 90 % empty space, 4K, skipping + early termination

Each is rendered in mode M1 with per-voxel stepping (dt_floor 0, dt_scale 2: box side 2, so one voxel
per step on the dominant axis, like raycast_naive.wgsl:97-99). Prints one JSON line per config with
frames/s (device-timed per frame, max over ranks), ray-samples/s, the HBM roofline (these volumes
exceed L2) and size-independent parity properties checked at FULL size: skip on == skip off and
tile == single, bit for bit; for N > 1 the sort-first frame equals the single-GPU frame.
usage: configs.py [--config 3|4|all] [--edge N override of the volume edge] [--frames F] [--res WxH]
"""
import argparse
import json
import math
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from vokselis_b200 import abi, rt  # noqa: E402

CONFIGS = {
    3: dict(name="config 3: 1024^3 fp16 noise, 4K, sort-first", kind=0, dtype=np.float16, n=1024, seed=3, zoom=3.0, pitch=-0.5),
    4: dict(name="config 4: 2048^3 u8, 90% empty super-bricks, 4K, skipping + ERT", kind=1, dtype=np.uint8, n=2048, seed=4, zoom=3.0, pitch=-0.5),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="all")
    ap.add_argument("--edge", dest="n", type=int, default=0)
    ap.add_argument("--frames", type=int, default=24)
    ap.add_argument("--res", default="3840x2160")
    ap.add_argument("--granularity", default="tiles")
    ap.add_argument("--layout", type=int, default=abi.LAYOUT_GATHER)
    ap.add_argument("--checks", type=int, default=1)
    ap.add_argument("--tile", type=int, default=120, help="sort-first tile edge in pixels (N > 1, granularity tiles)")
    args = ap.parse_args()
    W, H = map(int, args.res.split("x"))
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {"hbm_gbs": 6650.0}
    which = [3, 4] if args.config == "all" else [int(args.config)]
    for cid in which:
        cfg = CONFIGS[cid]
        n = args.n or cfg["n"]
        ctx = rt.Context(local, W, H)
        t0 = time.perf_counter()
        ctx.generate_synthetic(cfg["kind"], cfg["dtype"], n, seed=cfg["seed"])
        ctx.sync()
        gen_s = time.perf_counter() - t0
        info = ctx.volume_info()
        p = rt.default_params(abi.MODE_M1)
        p.dt_scale, p.dt_floor, p.skip_empty, p.layout = 2.0, 0.0, 1, args.layout
        ctx.set_params(p)
        cams = [rt.Camera(cfg["zoom"], cfg["pitch"], 1.0 + 2 * math.pi * i / args.frames, (0, 0, 0), W / H).get_proj_view_matrix()
                for i in range(args.frames)]
        checks = {}
        ref_frame = None
        if rank == 0:
            # sample statistics + full-size parity properties on frame 0
            q = rt.default_params(abi.MODE_M1)
            q.dt_scale, q.dt_floor, q.skip_empty, q.layout, q.count_samples = 2.0, 0.0, 1, args.layout, 1
            ctx.set_params(q)
            ctx.reset_stats()
            ctx.render(cams[0])
            st = ctx.stats()
            ref_frame = ctx.readback()
            if args.checks:
                aux1 = ctx.readback_aux()
                q.skip_empty = 0
                ctx.set_params(q)
                ctx.render(cams[0])
                checks["skip_on_equals_skip_off"] = bool(np.array_equal(ref_frame, ctx.readback()) and np.array_equal(aux1, ctx.readback_aux()))
                q.skip_empty, q.count_samples = 1, 0
                ctx.set_params(q)
                ctx.resize(W, H)
                ctx.render_tiles(cams[0], rt.tile_table(W, H, 256))
                checks["tile_equals_single"] = bool(np.array_equal(ref_frame, ctx.readback()))
            ctx.set_params(p)
        group = None
        if world > 1:
            from vokselis_b200 import sortfirst

            group = sortfirst.SortFirstGroup(ctx, rank, world, granularity=args.granularity, tile=args.tile)
            f = group.submit(cams[0])
            if rank == 0:
                group.wait(f)
                checks["sortfirst_equals_single_gpu"] = bool(np.array_equal(ref_frame, ctx.readback()))
                group.consume(f)
        ctx.timing_enable(args.frames)

        def render(cam):
            if group is None:
                ctx.flush_l2()
                ctx.render(cam)
            else:
                if group.granularity == "tiles" or sortfirst.frame_owner(group.frame, world) == rank:
                    ctx.flush_l2()
                group.render(cam)

        for cam in cams[:3]:
            render(cam)
        ctx.sync()
        if dist:
            dist.barrier()
        first = group.frame if group else 0
        t0 = time.perf_counter()
        for cam in cams:
            render(cam)
        ctx.sync()
        if dist:
            dist.barrier()
        wall = time.perf_counter() - t0
        mine = args.frames if group is None else group.my_frames(first, args.frames)
        ms = ctx.timing_read(mine).astype(np.float64) if mine else np.zeros(0)
        total = float(ms.sum())
        if dist:
            import torch

            t = torch.tensor([total], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total = float(t.item())
        if rank == 0:
            fps = args.frames / (total * 1e-3)
            eb = np.dtype(cfg["dtype"]).itemsize
            vol_bytes = n ** 3 * eb
            alg = st.samples_fetched * 8 * eb
            hbm_bytes = min(vol_bytes, alg) + W * H * 8
            kernel_ms = float(ms.mean())
            line = {
                "config": cfg["name"], "volume_edge": n, "dtype": np.dtype(cfg["dtype"]).name, "resolution": [W, H], "n_gpus": world,
                "granularity": args.granularity if world > 1 else None, "tile": args.tile if world > 1 else None, "layout": args.layout, "frames": args.frames,
                "frames_per_s": fps, "ms_per_frame": total / args.frames, "wall_ms_per_frame_incl_flush": 1e3 * wall / args.frames,
                "ray_samples_per_s": st.samples_reference * fps, "fetched_samples_per_s": st.samples_fetched * fps,
                "samples_frame0": {"reference": st.samples_reference, "fetched": st.samples_fetched, "rays": st.rays_hit},
                "bricks": {"total": info["bricks_total"], "occupied": info["bricks_occupied"]},
                "volume_bytes": vol_bytes, "generate_s": gen_s,
                "roofline_hbm": {"bound": "hbm", "achieved_GBs": hbm_bytes / (kernel_ms * 1e-3) / 1e9 * (world if args.granularity == "tiles" else 1),
                                 "peak_GBs": peaks["hbm_gbs"], "compulsory_bytes_per_frame": hbm_bytes,
                                 "note": "compulsory = min(volume bytes, fetched samples x 8 taps x sizeof) + W*H*8"},
                "texel_GBs": alg / (kernel_ms * 1e-3) / 1e9, "checks_at_full_size": checks,
            }
            print(json.dumps(line), flush=True)
        if group:
            group.close()
        ctx.close()
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

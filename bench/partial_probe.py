"""One rank's share of config 5 on ONE GPU (development, for ncu): the brick of rank `r` of `world` ranks of an edge^3 fp32
grid, marched by vkrt_partial_relative at 3840x2160. usage: partial_probe.py [edge=4096] [world=8] [rank=0] [frames=3]"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vokselis_b200 import abi, rt, sortlast, workloads  # noqa: E402

edge = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rank = int(sys.argv[3]) if len(sys.argv) > 3 else 0
frames = int(sys.argv[4]) if len(sys.argv) > 4 else 3
W, H = 3840, 2160
gn = (edge,) * 3
grid = sortlast.brick_grid(world)
lo, hi = sortlast.brick_range(gn, grid, rank)
with rt.Context(0, W, H) as ctx:
    ctx.generate_synthetic_window(3, np.float32, gn, lo, hi, seed=5)
    p = rt.default_params(abi.MODE_M1)
    p.dt_scale, p.dt_floor, p.skip_empty = 2.0, 0.0, 1
    ctx.set_params(p)
    n = W * H
    dev = torch.device("cuda", 0)
    T = torch.empty(n, dtype=torch.float32, device=dev)
    rgba = torch.empty(4 * n, dtype=torch.float32, device=dev)
    cams = workloads._cams(frames, W, H)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    with torch.cuda.stream(stream):
        ctx.partial_relative(cams[0], rgba.data_ptr(), T.data_ptr())
    ctx.sync()
    ms = []
    for cam in cams:
        with torch.cuda.stream(stream):
            ctx.mark(0)
            ctx.partial_relative(cam, rgba.data_ptr(), T.data_ptr())
            ctx.mark(1)
        ctx.sync()
        ms.append(ctx.mark_elapsed(0, 1))
    bytes_rank = float(np.prod([h - l for l, h in zip(lo, hi)])) * 4
    print(f"brick {lo}..{hi} ({bytes_rank / 2**30:.1f} GiB): march ms {np.round(ms, 3)}  -> {bytes_rank / (np.mean(ms) * 1e-3) / 1e9:.0f} GB/s of brick bytes", flush=True)

"""Direction dependence of the sort-last march (development): one rank's brick of config 5 on ONE GPU, viewed along the
three axes and two diagonals. usage: partial_axes.py [edge=4096] [world=8] [rank=0] [WxH=3840x2160]"""
import math
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vokselis_b200 import abi, rt, sortlast  # noqa: E402

edge = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rank = int(sys.argv[3]) if len(sys.argv) > 3 else 0
W, H = (int(v) for v in (sys.argv[4] if len(sys.argv) > 4 else "3840x2160").split("x"))
gn = (edge,) * 3
lo, hi = sortlast.brick_range(gn, sortlast.brick_grid(world), rank)
with rt.Context(0, W, H) as ctx:
    ctx.generate_synthetic_window(3, np.float32, gn, lo, hi, seed=5)
    p = rt.default_params(abi.MODE_M1)
    p.dt_scale, p.dt_floor, p.skip_empty = 2.0, 0.0, 1
    ctx.set_params(p)
    n = W * H
    dev = torch.device("cuda", 0)
    T = torch.empty(n, dtype=torch.float32, device=dev)
    rgba = torch.empty(4 * n, dtype=torch.float32, device=dev)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    views = [("yaw 0, pitch 0", 0.0, 0.0), ("yaw pi/2", 0.0, math.pi / 2), ("yaw pi", 0.0, math.pi), ("yaw 3pi/2", 0.0, 1.5 * math.pi),
             ("pitch +1.5", 1.5, 0.3), ("pitch -1.5", -1.5, 0.3), ("diagonal", -0.6, 0.785), ("bench cam 0", -0.5, 1.0)]
    print(f"edge {edge}, brick {lo}..{hi}, {W}x{H}", flush=True)
    for name, pitch, yaw in views:
        cam = rt.Camera(3.0, pitch, yaw, (0, 0, 0), W / H).get_proj_view_matrix()
        eye = sortlast.SortLastGroup.eye_of(cam)
        ms = []
        for _ in range(3):
            with torch.cuda.stream(stream):
                ctx.mark(0)
                ctx.partial_relative(cam, rgba.data_ptr(), T.data_ptr())
                ctx.mark(1)
            ctx.sync()
            ms.append(ctx.mark_elapsed(0, 1))
        print(f"{name:14s} eye ({eye[0]:+.2f}, {eye[1]:+.2f}, {eye[2]:+.2f})  march ms {np.round(ms, 2)}", flush=True)

"""A/B of raycast kernel variants on the bench workload (development tool; numbers for profiles/).

usage: kernel_ab.py [--vols xor,bonsai] [--layouts 3,4] [--skips 1,0] [--launches 12] [--batch 8] [--modes 1]
Prints one JSON line per variant: ms per frame (CUDA events, 8 frames per launch, L2 flushed between launches),
whether its frames are bit-identical to the first variant's of the same volume, and a hash of the last frame
(to compare builds across processes: bench/ab_pick.sh).
"""
import argparse
import hashlib
import json
import math
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vokselis_b200 import abi, rt, volumes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--vols", default="xor,bonsai")
ap.add_argument("--layouts", default="3,4")
ap.add_argument("--skips", default="1,0")
ap.add_argument("--launches", type=int, default=12)
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--size", default="1920x1080")
ap.add_argument("--n", type=int, default=256)
ap.add_argument("--bricks", default="0", help="occupancy brick edges to compare (0 = automatic)")
args = ap.parse_args()
W, H = (int(v) for v in args.size.split("x"))
ORBIT = 360
cams = [rt.Camera(3.0, -0.5, 1.0 + 2.0 * math.pi * i / ORBIT, (0.0, 0.0, 0.0), W / H).get_proj_view_matrix() for i in range(ORBIT)]
B = args.batch
refs = {}
with rt.Context(0, W, H) as ctx:
    for vname, brick in ((v, int(b)) for v in args.vols.split(",") for b in args.bricks.split(",")):
        vol = volumes.xor_u8(args.n) if vname == "xor" else volumes.bonsai_standin_u8(args.n, seed=1)
        ctx.set_occupancy_brick(brick)
        ctx.upload_scalar(vol)
        for skip in (int(v) for v in args.skips.split(",")):
            ref = refs.get(vname)  # exact skipping: one set of bits per volume, whatever the brick edge, skip on or off
            for layout in (int(v) for v in args.layouts.split(",")):
                p = rt.default_params(abi.MODE_M1)
                p.layout, p.skip_empty = layout, skip
                ctx.set_params(p)
                ctx.render_batch(cams[:B])  # warm-up: layout build, module load
                ctx.render_batch(cams[:B])
                ctx.timing_enable(args.launches)
                for j in range(args.launches):
                    ctx.flush_l2()
                    ctx.render_batch([cams[(20 + j * B + k) % ORBIT] for k in range(B)])
                ms = ctx.timing_read(args.launches).astype(np.float64)
                f0 = ctx.readback_batch(B - 1)
                same = None
                if ref is None:
                    ref = refs[vname] = f0
                else:
                    same = bool(np.array_equal(ref, f0))
                print(json.dumps({"vol": vname, "brick": brick, "layout": layout, "skip": skip, "ms_per_frame": float(ms.mean()) / B,
                                  "fps": 1e3 * B / float(ms.mean()), "p10_p90_ms_per_launch": [float(np.percentile(ms, 10)), float(np.percentile(ms, 90))],
                                  "same_bits_as_first_layout": same,
                                  "sha256_last_frame": hashlib.sha256(f0.tobytes()).hexdigest()[:16]}), flush=True)

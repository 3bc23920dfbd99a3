"""How a frame's tile share scales on ONE GPU (development): config 4's volume at 4K, the tiles rank 0 of a world of
1/2/4/8 would get, timed per launch (CUDA events), back to back on one stream. usage: tiles_subset.py [cid=4] [tile=120]"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vokselis_b200 import abi, rt, workloads  # noqa: E402

cid = int(sys.argv[1]) if len(sys.argv) > 1 else 4
tile = int(sys.argv[2]) if len(sys.argv) > 2 else 120
cfg = workloads.SORTFIRST_CONFIGS[cid]
W, H = 3840, 2160
with rt.Context(0, W, H) as ctx:
    ctx.generate_synthetic(cfg["kind"], cfg["dtype"], cfg["n"], seed=cfg["seed"])
    p = rt.default_params(abi.MODE_M1)
    p.dt_scale, p.dt_floor, p.skip_empty, p.layout, p.tile_size = 2.0, 0.0, 1, abi.LAYOUT_GATHER, tile
    ctx.set_params(p)
    cams = workloads._cams(12, W, H)
    for world in (1, 2, 4, 8, 16):
        offs = rt.sortfirst_partition(W, H, tile, 0, world)
        for cam in cams[:3]:
            ctx.render_tiles(cam, offs)
        ctx.timing_enable(len(cams))
        for cam in cams:
            ctx.render_tiles(cam, offs)
        ms = ctx.timing_read(len(cams)).astype(np.float64)
        print(f"world {world:2d}: {len(offs):4d} tiles  {ms.mean():.4f} ms per launch  x world = {ms.mean() * world:.4f}  (p10 {np.percentile(ms, 10):.4f}, p90 {np.percentile(ms, 90):.4f})", flush=True)

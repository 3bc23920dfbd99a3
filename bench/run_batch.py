"""Render a few batched launches of the bench workload (for ncu captures). usage: run_batch.py [frames_per_launch=8] [launches=3] [layout=4] [skip=1] [vol=xor|bonsai]"""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vokselis_b200 import abi, rt, volumes
F = int(sys.argv[1]) if len(sys.argv) > 1 else 8
L = int(sys.argv[2]) if len(sys.argv) > 2 else 3
LAYOUT = int(sys.argv[3]) if len(sys.argv) > 3 else abi.LAYOUT_QUAD
SKIP = int(sys.argv[4]) if len(sys.argv) > 4 else 1
VOL = sys.argv[5] if len(sys.argv) > 5 else "xor"
W, H = 1920, 1080
with rt.Context(0, W, H) as ctx:
    ctx.upload_scalar(volumes.xor_u8(256) if VOL == "xor" else volumes.bonsai_standin_u8(256, seed=1))
    p = rt.default_params(abi.MODE_M1); p.skip_empty = SKIP; p.layout = LAYOUT
    ctx.set_params(p)
    cams = [rt.Camera(3.0, -0.5, 1.0 + 2 * np.pi * i / 360, (0, 0, 0), W / H).get_proj_view_matrix() for i in range(360)]
    ctx.timing_enable(L)
    for j in range(L):
        ctx.flush_l2()
        ctx.render_batch(cams[20 + j * F:20 + (j + 1) * F])
    print("launch ms:", np.round(ctx.timing_read(L), 4))

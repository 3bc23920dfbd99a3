"""How much does frame-level concurrency buy? N contexts (one stream each) render the orbit without syncing in
between; total wall time per frame vs one context. (development probe for the batched-launch design)"""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vokselis_b200 import abi, rt, volumes
W, H = 1920, 1080
vol = volumes.xor_u8(256)
cams = [rt.Camera(3.0, -0.5, 1.0 + 2 * np.pi * i / 360, (0, 0, 0), W / H).get_proj_view_matrix() for i in range(360)]
for n in (1, 2, 3, 4):
    ctxs = []
    for _ in range(n):
        c = rt.Context(0, W, H)
        c.upload_scalar(vol)
        p = rt.default_params(abi.MODE_M1); p.skip_empty = 1; p.layout = abi.LAYOUT_GATHER
        c.set_params(p)
        ctxs.append(c)
    K = 360
    for i in range(20):
        ctxs[i % n].render(cams[i])
    for c in ctxs: c.sync()
    t0 = time.perf_counter()
    for i in range(K):
        ctxs[i % n].render(cams[i])
    for c in ctxs: c.sync()
    dt = (time.perf_counter() - t0) / K
    print(f"{n} streams: {dt*1e3:.4f} ms/frame -> {1/dt:.0f} frames/s", flush=True)
    for c in ctxs: c.close()

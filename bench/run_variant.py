"""Render a few frames of one kernel variant (for ncu captures). Development tool.
usage: run_variant.py MODE(0|1) LAYOUT(0|1|2) SKIP(0|1) [frames=6] [W=1920 H=1080] [vol=xor_u8|bonsai]"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vokselis_b200 import abi, rt, volumes  # noqa: E402

mode, layout, skip = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
frames = int(sys.argv[4]) if len(sys.argv) > 4 else 6
W = int(sys.argv[5]) if len(sys.argv) > 5 else 1920
H = int(sys.argv[6]) if len(sys.argv) > 6 else 1080
volname = sys.argv[7] if len(sys.argv) > 7 else "xor_u8"
with rt.Context(0, W, H) as ctx:
    if mode == 0:
        ctx.generate_xor(256, 0)
        zoom, pitch = 3.0, -0.5
    else:
        ctx.upload_scalar(volumes.xor_u8(256) if volname == "xor_u8" else volumes.bonsai_standin_u8(256))
        zoom, pitch = (3.0, -0.5) if volname == "xor_u8" else (2.0, 0.5)
    p = rt.default_params(mode)
    p.layout, p.skip_empty = layout, skip
    ctx.set_params(p)
    ctx.timing_enable(frames)
    for i in range(frames):
        ctx.render(rt.Camera(zoom, pitch, 1.0 + 2 * np.pi * i / 36, (0, 0, 0), W / H).get_proj_view_matrix())
    ms = ctx.timing_read(frames)
    print("frame ms:", np.round(ms[:6], 4), "median", round(float(np.median(ms[2:])), 4), "mean", round(float(np.mean(ms[2:])), 4))
    # one more frame through the counting kernel: reference-semantics iterations vs samples actually fetched
    p.count_samples = 1
    ctx.set_params(p)
    ctx.reset_stats()
    ctx.render(rt.Camera(zoom, pitch, 1.0, (0, 0, 0), W / H).get_proj_view_matrix())
    st = ctx.stats()
    print("stats: rays_hit", st.rays_hit, "reference", st.samples_reference, "fetched", st.samples_fetched)

"""Throughput of the batched launch (vkrt_render_batch) vs frames per launch, and of vkrt_frames_host vs group size.
usage: batch_sweep.py [mode=1] [vol=xor_u8|bonsai]   (development / profiles/r01_batch.md)"""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vokselis_b200 import abi, rt, volumes
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 1
volname = sys.argv[2] if len(sys.argv) > 2 else "xor_u8"
W, H = 1920, 1080
ctx = rt.Context(0, W, H)
if mode == 0:
    ctx.generate_xor(256, 0); layout = abi.LAYOUT_TEXTURE; zoom, pitch = 3.0, -0.5
else:
    ctx.upload_scalar(volumes.xor_u8(256) if volname == "xor_u8" else volumes.bonsai_standin_u8(256)); layout = abi.LAYOUT_GATHER
    zoom, pitch = (3.0, -0.5) if volname == "xor_u8" else (2.0, 0.5)
p = rt.default_params(mode); p.skip_empty = 1; p.layout = layout
ctx.set_params(p)
cams = [rt.Camera(zoom, pitch, 1.0 + 2 * np.pi * i / 360, (0, 0, 0), W / H).get_proj_view_matrix() for i in range(360)]
K = 360
for F in (1, 2, 3, 4, 6, 8):
    L = K // F
    ctx.timing_enable(L)
    for j in range(3): ctx.render_batch(cams[j * F:(j + 1) * F])
    for flush in (True, False):
        for j in range(L):
            if flush: ctx.flush_l2()
            ctx.render_batch(cams[j * F:(j + 1) * F])
        ms = ctx.timing_read(L).astype(np.float64)
        print(f"mode {mode} {volname} batch {F}: {ms.mean() / F:.4f} ms/frame -> {F * 1e3 / ms.mean():.0f} frames/s (device time per launch {ms.mean():.4f} ms, L2 {'flushed' if flush else 'warm'})", flush=True)
pin = rt.PinnedArray((24, H, W, 4), np.uint8)
for G in (1, 2, 3, 4, 6, 8):
    ctx.frames_host(cams[:24], pin.array, group=G)
    ctx.sync(); t0 = time.perf_counter()
    for j in range(K // 24):
        ctx.frames_host(cams[j * 24:(j + 1) * 24], pin.array, group=G)
    dt = (time.perf_counter() - t0) / (K // 24 * 24)
    print(f"frames_host group {G}: {dt * 1e3:.4f} ms/frame -> {1 / dt:.0f} frames/s e2e (24 frames per call)", flush=True)

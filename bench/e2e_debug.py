"""Why is the pipelined host path slower than the blocking one? Times the pieces."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vokselis_b200 import abi, rt, volumes

W, H = 1920, 1080
ctx = rt.Context(0, W, H)
ctx.upload_scalar(volumes.xor_u8(256))
p = rt.default_params(abi.MODE_M1); p.skip_empty = 1; p.layout = abi.LAYOUT_GATHER
ctx.set_params(p)
cams = [rt.Camera(3.0, -0.5, 1.0 + 2 * np.pi * i / 360, (0, 0, 0), W / H).get_proj_view_matrix() for i in range(360)]
pin = rt.PinnedArray((H, W, 4), np.uint8)
K = 300
def t(label, fn):
    for i in range(10): fn(i)
    ctx.sync(); t0 = time.perf_counter()
    for i in range(K): fn(i)
    ctx.sync(); dt = (time.perf_counter() - t0) / K
    print(f"{label}: {dt*1e3:.4f} ms/frame -> {1/dt:.0f} fps", flush=True)
t("render only (async)", lambda i: ctx.render(cams[i % 360]))
t("render + present (async)", lambda i: (ctx.render(cams[i % 360]), ctx.present()))
t("blocking frame_host pinned", lambda i: ctx.frame_host(cams[i % 360], pin.array))
def pipe(i):
    s = i & 1
    if i >= 2: ctx.frame_host_wait(s, None)
    ctx.frame_host_async(cams[i % 360], s)
t("pipelined 2 slots", pipe)
out = np.empty((H, W, 4), np.uint8)
t("readback_rgba8 pinned only", lambda i: ctx.readback_rgba8(pin.array))
t("readback_rgba8 pageable only", lambda i: ctx.readback_rgba8(out))

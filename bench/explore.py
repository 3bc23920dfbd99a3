"""Explore kernel variants on the GPU box: times every (mode, layout, skip) at the bench resolution
and writes gpurun_out/explore.json. Development tool, not the judged bench (see bench.py)."""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from vokselis_b200 import abi, rt, volumes  # noqa: E402


def time_variant(ctx, cams, iters=3):
    # warm-up
    for cam in cams[:4]:
        ctx.render(cam)
    ctx.sync()
    best = 1e9
    for _ in range(iters):
        t0 = time.perf_counter()
        for cam in cams:
            ctx.render(cam)
        ctx.sync()
        best = min(best, (time.perf_counter() - t0) / len(cams))
    return best * 1e3


def main():
    W, H = int(os.environ.get("W", 1920)), int(os.environ.get("H", 1080))
    out = {"W": W, "H": H, "rows": []}
    nframes = 36
    cams = [rt.Camera(3.0, -0.5, 1.0 + 2 * np.pi * i / nframes, (0, 0, 0), W / H).get_proj_view_matrix() for i in range(nframes)]
    with rt.Context(0, W, H) as ctx:
        ctx.generate_xor(256, 0)
        for layout in (0, 1, 2):
            for skip in (0, 1):
                p = rt.default_params(abi.MODE_M0)
                p.layout, p.skip_empty, p.count_samples = layout, skip, 1
                ctx.set_params(p)
                ctx.reset_stats()
                ctx.render(cams[0])
                st = ctx.stats()
                p.count_samples = 0
                ctx.set_params(p)
                ms = time_variant(ctx, cams)
                row = {"mode": "M0", "layout": layout, "skip": skip, "ms": ms, "fps": 1e3 / ms, "samples_ref": st.samples_reference,
                       "samples_fetched": st.samples_fetched, "rays": st.rays_hit, "frame0_ms": st.last_render_ms}
                print(row, flush=True)
                out["rows"].append(row)
    vols = {"xor_u8": volumes.xor_u8(256), "bonsai_standin_u8": volumes.bonsai_standin_u8(256)}
    for name, vol in vols.items():
        zoom = 3.0 if name == "xor_u8" else 2.0
        pitch = -0.5 if name == "xor_u8" else 0.5
        cams = [rt.Camera(zoom, pitch, 1.0 + 2 * np.pi * i / nframes, (0, 0, 0), W / H).get_proj_view_matrix() for i in range(nframes)]
        with rt.Context(0, W, H) as ctx:
            ctx.upload_scalar(vol)
            print(name, ctx.volume_info(), flush=True)
            for layout in (0, 2, 3):
                for skip in (0, 1):
                    for dts in (1.0, 2.0):
                        p = rt.default_params(abi.MODE_M1)
                        p.layout, p.skip_empty, p.count_samples, p.dt_scale = layout, skip, 1, dts
                        ctx.set_params(p)
                        ctx.reset_stats()
                        ctx.render(cams[0])
                        st = ctx.stats()
                        p.count_samples = 0
                        ctx.set_params(p)
                        ms = time_variant(ctx, cams)
                        row = {"mode": "M1", "vol": name, "layout": layout, "skip": skip, "dt_scale": dts, "ms": ms, "fps": 1e3 / ms,
                               "samples_ref": st.samples_reference, "samples_fetched": st.samples_fetched, "rays": st.rays_hit}
                        print(row, flush=True)
                        out["rows"].append(row)
    os.makedirs(ROOT / "gpurun_out", exist_ok=True)
    with open(ROOT / "gpurun_out" / "explore.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()

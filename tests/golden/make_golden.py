#!/usr/bin/env python3
"""Generate the golden vectors under tests/golden/ from oracle/_ref — the reference's OWN WGSL
(/root/reference/shaders/*.wgsl) machine-translated to C++ by oracle/wgsl2cpp.py and executed on the
CPU. Run in the build container (needs /root/reference): `make -C oracle ref && python tests/golden/make_golden.py`.
The .npz files are committed; the GPU box (no /root/reference) checks against them.

The reference ships no golden vectors of its own; these are outputs of the reference's
shader text under the builtin definitions fixed in oracle/wgsl_rt.hpp.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import binding as ob  # noqa: E402  (camera only)
from oracle import ref_binding as rb  # noqa: E402

OUT = Path(__file__).resolve().parent


def cam_bytes(cam):
    return np.frombuffer(bytes(cam), np.uint8).copy()


def main():
    assert rb.available(), "build oracle/_ref first: make -C oracle ref"
    # G1: the xor example end to end at reduced size: generator -> raycast single / tile -> present
    n, W, H = 32, 160, 90
    color, normal = rb.xor_generate(n, 0.0)
    cam = ob.camera_uniform(3.0, -0.5, 1.0, (0.0, 0.0, 0.0), W / H)
    single = rb.raycast_compute(cam, color, normal, W, H, entry="single")
    ts = 64
    table = np.array([[x * ts, y * ts] for y in range(H // ts + 1) for x in range(W // ts + 1)], np.float32)
    tiled = rb.raycast_compute(cam, color, normal, W, H, entry="tile", offsets=table, tile_size=ts)
    np.savez_compressed(OUT / "g1_xor32_160x90.npz", color=color, normal=normal, cam=cam_bytes(cam), single=single, tiled=tiled,
                        table=table, tile_size=np.int32(ts), present=rb.present(single))
    # G2: adversarial random volume, non-cubic, dims straddling the dt floor (N_x = 192 >= 174),
    # NaN/inf/negative texels, camera inside the box for some rays
    rng = np.random.default_rng(7)
    nz, ny, nx = 10, 12, 192
    c = rng.uniform(0.0, 1.0, size=(nz, ny, nx, 4)).astype(np.float16)
    c[..., 3] = (rng.uniform(0, 1, size=(nz, ny, nx)) ** 3 * 0.9).astype(np.float16)
    c[rng.uniform(size=(nz, ny, nx)) < 0.4, 3] = 0
    nr = rng.normal(size=(nz, ny, nx, 4)).astype(np.float16)
    nr[rng.uniform(size=(nz, ny, nx)) < 0.2] = np.float16(np.nan)
    nr[0, 0, 0, 0] = np.float16(np.inf)
    c[1, 1, 1, 3] = np.float16(-0.5)
    W2, H2 = 128, 72
    cam2 = ob.camera_uniform(1.6, 0.35, -2.2, (0.1, -0.05, 0.2), W2 / H2)
    f2 = rb.raycast_compute(cam2, c.view(np.uint16), nr.view(np.uint16), W2, H2, entry="single")
    cam3 = ob.camera_uniform(0.9, 0.1, 0.4, (0.0, 0.0, 0.0), W2 / H2)  # near plane inside the box
    f3 = rb.raycast_compute(cam3, c.view(np.uint16), nr.view(np.uint16), W2, H2, entry="single")
    np.savez_compressed(OUT / "g2_random192x12x10_128x72.npz", color=c.view(np.uint16), normal=nr.view(np.uint16),
                        cam=cam_bytes(cam2), frame=f2, cam_inside=cam_bytes(cam3), frame_inside=f3)
    # G3: generator bytes at two times
    c0, n0 = rb.xor_generate(16, 0.0)
    c1, n1 = rb.xor_generate(16, 1.3)
    np.savez_compressed(OUT / "g3_xorgen16.npz", color_t0=c0, normal_t0=n0, color_t13=c1, normal_t13=n1)
    # G5: raycast_naive.wgsl fs_main on 96 fragments over a small smooth u8 volume
    m = 24
    g = np.linspace(0, 1, m, dtype=np.float32)
    vol = (np.exp(-(((g[:, None, None] - 0.5) ** 2 + (g[None, :, None] - 0.45) ** 2 + (g[None, None, :] - 0.55) ** 2) / 0.05)) * 230
           + rng.uniform(0, 25, size=(m, m, m))).clip(0, 255).astype(np.uint8)
    eyes = np.tile(np.array([[1.6, 1.2, -0.7]], np.float32), (96, 1))
    tgt = rng.uniform(0.1, 0.9, size=(96, 3)).astype(np.float32)
    dirs = (tgt - eyes) * rng.uniform(0.5, 2.0, size=(96, 1)).astype(np.float32)
    dirs[-1] = [0.0, 1.0, 0.0]  # a miss
    cols = rb.naive_fs(vol, eyes, dirs)
    np.savez_compressed(OUT / "g5_naive_fs24.npz", vol=vol, eyes=eyes, dirs=dirs, colors=cols)
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()

"""Synthetic volumes of the shapes BASELINE.json names (host-side numpy; deterministic).

No dataset ships with the reference: bonsai_256x256x256_uint8.raw is listed in
.MISSING_LARGE_BLOBS, so every "bonsai" run uses `bonsai_standin_u8` and says so.
"""
from __future__ import annotations

import numpy as np


def _smoothstep(e0, e1, x):
    t = np.clip((x - np.float32(e0)) / (np.float32(e1) - np.float32(e0)), np.float32(0), np.float32(1))
    return t * t * (np.float32(3) - np.float32(2) * t)


def xor_pattern_alpha(n: int = 256) -> np.ndarray:
    """shaders/xor.wgsl:46-53 `volume()` at time = 0 (the bit pattern the `xor` example is named after,
    dead code in the reference): alpha = val * smoothstep(0.7, 0, |coord|), val = (ix & iy & iz) / 25
    with i* = i32(pos.* * 25), pos = (coord + (1, 0, 21)) * 32. float32 [nz, ny, nx]."""
    i = np.arange(n, dtype=np.float32)
    coord = (i - np.float32(n) / np.float32(2)) / np.float32(n)
    res = np.float32(25)
    ix = ((coord + np.float32(1)) * np.float32(32) * res).astype(np.int32)
    iy = ((coord + np.float32(0)) * np.float32(32) * res).astype(np.int32)  # truncation toward zero, like i32()
    iz = ((coord + np.float32(21)) * np.float32(32) * res).astype(np.int32)
    val = (ix[None, None, :] & iy[None, :, None] & iz[:, None, None]).astype(np.float32) / res
    c2 = coord * coord
    length = np.sqrt((c2[None, None, :] + c2[None, :, None]) + c2[:, None, None])
    return val * _smoothstep(0.7, 0.0, length)


def xor_u8(n: int = 256) -> np.ndarray:
    """BASELINE config 2 volume: the bit pattern above quantised to R8Unorm as round(alpha / 16 * 255)."""
    a = xor_pattern_alpha(n)
    return np.clip(np.rint(a * np.float32(255.0 / 16.0)), 0, 255).astype(np.uint8)


def bonsai_standin_u8(n: int = 256, seed: int = 1, blobs: int = 64) -> np.ndarray:
    """Stand-in for the missing bonsai CT scan: a sum of `blobs` Gaussian blobs plus low-amplitude
    value noise, ~70 % of voxels transparent under the vertigo transfer function (value <= 25)."""
    rng = np.random.default_rng(seed)
    g = (np.arange(n, dtype=np.float32) + np.float32(0.5)) / np.float32(n)
    vol = np.zeros((n, n, n), np.float32)
    centres = rng.uniform(0.25, 0.75, size=(blobs, 3)).astype(np.float32)
    sigmas = rng.uniform(0.03, 0.09, size=blobs).astype(np.float32)
    amps = rng.uniform(0.35, 1.0, size=blobs).astype(np.float32)
    for (cx, cy, cz), s, a in zip(centres, sigmas, amps):
        ex = np.exp(-((g - cx) ** 2) / (2 * s * s))
        ey = np.exp(-((g - cy) ** 2) / (2 * s * s))
        ez = np.exp(-((g - cz) ** 2) / (2 * s * s))
        vol += a * ez[:, None, None] * ey[None, :, None] * ex[None, None, :]
    coarse = rng.uniform(0.0, 1.0, size=(n // 8 + 2,) * 3).astype(np.float32)
    idx = np.arange(n) // 8
    vol += np.float32(0.04) * coarse[idx[:, None, None], idx[None, :, None], idx[None, None, :]]
    vol = vol / np.float32(max(vol.max(), 1e-6))
    return np.clip(np.rint(vol * 255.0), 0, 255).astype(np.uint8)


def hash_noise(n: int, seed: int, dtype=np.float16, smooth: bool = True, chunk: int = 64) -> np.ndarray:
    """Uniform integer-hash noise in [0,1), optionally box-filtered 3^3 once (BASELINE config 3)."""
    out = np.empty((n, n, n), dtype)

    def h(z0, z1):
        z, y, x = np.meshgrid(np.arange(z0, z1, dtype=np.uint32), np.arange(n, dtype=np.uint32),
                              np.arange(n, dtype=np.uint32), indexing="ij")
        v = (x * np.uint32(0x9E3779B1)) ^ (y * np.uint32(0x85EBCA77)) ^ (z * np.uint32(0xC2B2AE3D)) ^ np.uint32(seed * 0x27D4EB2F & 0xFFFFFFFF)
        v ^= v >> np.uint32(15)
        v *= np.uint32(0x2C1B3C6D)
        v ^= v >> np.uint32(12)
        v *= np.uint32(0x297A2D39)
        v ^= v >> np.uint32(15)
        return (v >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / (1 << 24))

    for z0 in range(0, n, chunk):
        z1 = min(n, z0 + chunk)
        if not smooth:
            out[z0:z1] = h(z0, z1).astype(dtype)
            continue
        lo, hi = max(z0 - 1, 0), min(z1 + 1, n)
        f = h(lo, hi)
        p = np.pad(f, ((1 if lo == z0 else 0, 1 if hi == z1 else 0), (1, 1), (1, 1)), mode="edge")
        acc = np.zeros((z1 - z0, n, n), np.float32)
        for dz in range(3):
            for dy in range(3):
                for dx in range(3):
                    acc += p[dz:dz + (z1 - z0), dy:dy + n, dx:dx + n]
        out[z0:z1] = (acc / np.float32(27)).astype(dtype)
    return out

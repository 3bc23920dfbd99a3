"""CPU model (numpy, float64) of the headline kernel's traversal: how many leap / sample iterations does a warp
(8x4 pixels) execute with the isotropic Chebyshev distance field, and how many with one distance field per ray
octant (largest empty box of bricks in the ray's direction of travel)?  No GPU; used to decide the layout of the
distance field before measuring.

usage: python bench/leap_model.py [frame ...]        (orbit frames of bench.py, default 0 45 100)
"""
from __future__ import annotations

import math
import sys

import numpy as np

sys.path.insert(0, ".")
from vokselis_b200 import volumes  # noqa: E402

W, H, N = 1920, 1080, 256
ORBIT = 360
BR = 8  # brick edge in voxels


def camera_inv(zoom, pitch, yaw, aspect):
    eye = np.array([-zoom * math.sin(yaw) * math.cos(pitch), -zoom * math.sin(pitch), -zoom * math.cos(yaw) * math.cos(pitch)])
    f = -eye / np.linalg.norm(eye)
    s = np.cross(f, [0.0, 1.0, 0.0])
    s /= np.linalg.norm(s)
    u = np.cross(s, f)
    view = np.eye(4)
    view[0, :3], view[1, :3], view[2, :3] = s, u, -f
    view[0, 3], view[1, 3], view[2, 3] = -s @ eye, -u @ eye, f @ eye
    fov, zn, zf = math.pi / 2, 0.1, 100.0
    h = 1.0 / math.tan(fov / 2)
    proj = np.zeros((4, 4))
    proj[0, 0], proj[1, 1] = h / aspect, h
    proj[2, 2], proj[2, 3] = zf / (zn - zf), zf / (zn - zf) * zn
    proj[3, 2] = -1.0
    return np.linalg.inv(proj @ view)


def rays(inv, px, py):
    sx = 2.0 * px / W - 1.0
    sy = (2.0 * py / H - 1.0) * -(H / W)
    one = np.ones_like(sx)
    p = inv @ np.stack([sx, sy, 0 * one, one])
    t = inv @ np.stack([sx, sy, one, one])
    eye = (p[:3] / p[3]).T
    d = (t[:3] / t[3]).T - eye
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return eye, d


def occupancy(vol):
    """brick empty iff all voxels [8b-1, 8b+8] per axis (clamped) <= 0.0999999 (volume.cu occupancy_m1_kernel)"""
    occ_v = (vol.astype(np.float32) / np.float32(255) > np.float32(0.0999999))
    nb = N // BR
    # dilate by one voxel (max over the 3^3 neighbourhood), then reduce over bricks
    p = np.pad(occ_v, 1, mode="edge")
    dil = np.zeros_like(occ_v)
    for dz in range(3):
        for dy in range(3):
            for dx in range(3):
                dil |= p[dz:dz + N, dy:dy + N, dx:dx + N]
    return dil.reshape(nb, BR, nb, BR, nb, BR).any(axis=(1, 3, 5))  # [bz, by, bx]


def chebyshev(occ):
    nb = occ.shape[0]
    d = np.where(occ, 0, 255).astype(np.int32)
    for _ in range(nb):
        p = np.pad(d, 1, constant_values=0)  # border occupied (M1)
        m = np.full_like(d, 255)
        for dz in range(3):
            for dy in range(3):
                for dx in range(3):
                    m = np.minimum(m, p[dz:dz + nb, dy:dy + nb, dx:dx + nb])
        nd = np.where(occ, 0, np.minimum(d, m + 1))
        if (nd == d).all():
            break
        d = nd
    return d


def octant(occ, sx, sy, sz):
    """largest d such that the d^3 bricks [b, b + s*d) are empty (0 if b is occupied); outside the grid = occupied"""
    nb = occ.shape[0]
    o = occ[::sz, ::sy, ::sx]  # now the direction of travel is +,+,+
    d = np.zeros((nb + 1,) * 3, np.int32)
    for z in range(nb - 1, -1, -1):
        for y in range(nb - 1, -1, -1):
            row_zy = np.minimum(np.minimum(d[z + 1, y], d[z, y + 1]), d[z + 1, y + 1])  # [nb+1], neighbours without +x
            prev = 0
            out = d[z, y]
            for x in range(nb - 1, -1, -1):
                if o[z, y, x]:
                    prev = 0
                else:
                    prev = 1 + min(prev, row_zy[x], row_zy[x + 1])
                out[x] = prev
    return d[:nb, :nb, :nb][::sz, ::sy, ::sx]


def sample_volume(vol, q):
    u = q - 0.5
    fl = np.floor(u)
    f = u - fl
    i0 = np.clip(fl.astype(np.int64), 0, N - 1)
    i1 = np.clip(fl.astype(np.int64) + 1, 0, N - 1)
    x0, y0, z0 = i0[:, 0], i0[:, 1], i0[:, 2]
    x1, y1, z1 = i1[:, 0], i1[:, 1], i1[:, 2]
    fx, fy, fz = f[:, 0], f[:, 1], f[:, 2]
    v = vol
    c00 = v[z0, y0, x0] * (1 - fx) + v[z0, y0, x1] * fx
    c10 = v[z0, y1, x0] * (1 - fx) + v[z0, y1, x1] * fx
    c01 = v[z1, y0, x0] * (1 - fx) + v[z1, y0, x1] * fx
    c11 = v[z1, y1, x0] * (1 - fx) + v[z1, y1, x1] * fx
    return ((c00 * (1 - fy) + c10 * fy) * (1 - fz) + (c01 * (1 - fy) + c11 * fy) * fz) / 255.0


def march(vol, occ, dist_of_ray, eye, d, bb_lo, bb_hi, blind=0, join=0):
    """returns per-ray (leaps, samples). dist_of_ray(bz, by, bx, idx) -> distance for the rays `idx`.
    blind = K: a lane that has just evaluated a sample evaluates up to K more without consulting the distance field, inside the
    same lock-step iteration (evaluating a sample is always exact; only skipping needs the field).
    join = D: in an iteration where some lane of the warp samples anyway, lanes whose distance is <= D sample too instead of
    leaping (a short leap costs the warp a whole pass through the leap path)."""
    n = len(eye)
    with np.errstate(divide="ignore"):
        inv = 1.0 / d
    a, b = (-1 - eye) * inv, (1 - eye) * inv
    t0 = np.minimum(a, b).max(axis=1)
    t1 = np.maximum(a, b).min(axis=1)
    hit = t0 < t1
    t0 = np.maximum(t0, 0)
    dt = np.minimum.reduce([1.0 / (N * np.abs(d[:, k])) for k in range(3)])
    a, b = (bb_lo - eye) * inv, (bb_hi - eye) * inv
    tb0 = np.minimum(a, b).max(axis=1)
    tb1 = np.maximum(a, b).min(axis=1)
    t_end = np.where(tb0 < tb1, np.minimum(t1, tb1), -1.0)
    t = t0.copy()
    ent = hit & (tb0 > t0) & (tb0 < t_end)
    n0 = np.floor((tb0 - t0) / dt - 2.0)
    t = np.where(ent & (n0 >= 1), t + n0 * dt, t)
    dq = d * (N / 2) * dt[:, None]  # voxels per step
    with np.errstate(divide="ignore"):
        rq = np.where(np.abs(dq) > 1e-12, 1.0 / dq, 1e30)
    sg = np.sign(rq)
    alpha = np.zeros(n)
    leaps = np.zeros(n, np.int64)
    samples = np.zeros(n, np.int64)
    live = hit & (t < t_end)
    idx = np.nonzero(live)[0]
    nbk = N // BR
    ev = np.zeros(4, np.int64)  # lock-step warp events: iterations, iterations with a leap, iterations with a sample, blind sample passes
    nwarps = n // 32
    while len(idx):
        p = eye[idx] + t[idx, None] * d[idx]
        q = (p + 1) * (N / 2)
        bi = np.clip(np.floor(q / BR).astype(np.int64), 0, nbk - 1)
        inside = ((q >= 0) & (q < N)).all(axis=1)
        dd = dist_of_ray(bi[:, 2], bi[:, 1], bi[:, 0], idx)
        dd = np.where(inside, dd, 0)
        lp = dd > 0
        if join:
            wid0 = idx // 32
            samp_warps = np.unique(wid0[~lp])
            lp &= ~((dd <= join) & np.isin(wid0, samp_warps))
        # leap
        R = BR * dd - BR / 2 - 1e-3
        w = BR * bi + (sg[idx] * R[:, None] + (BR / 2 - q))
        s = (w * rq[idx]).min(axis=1)
        nn = np.maximum(np.floor(s + 0.98), 1)
        wid = idx // 32
        ev[0] += len(np.unique(wid))
        ev[1] += len(np.unique(wid[lp]))
        ev[2] += len(np.unique(wid[~lp]))
        li = idx[lp]
        t[li] += nn[lp] * dt[li]
        leaps[li] += 1
        # sample
        si = idx[~lp]
        if len(si):
            sv = sample_volume(vol, q[~lp])
            x = np.clip((np.minimum(0.9, sv) - 0.10) / (1.2 - 0.10), 0, 1)
            v = x * x * (3 - 2 * x)
            alpha[si] += (1 - alpha[si]) * v
            samples[si] += 1
            t[si] += dt[si]
            for _ in range(blind):
                si = si[(t[si] < t_end[si]) & (alpha[si] < 0.95)]
                if not len(si):
                    break
                ev[3] += len(np.unique(si // 32))
                pb = eye[si] + t[si, None] * d[si]
                sv = sample_volume(vol, (pb + 1) * (N / 2))
                x = np.clip((np.minimum(0.9, sv) - 0.10) / (1.2 - 0.10), 0, 1)
                v = x * x * (3 - 2 * x)
                alpha[si] += (1 - alpha[si]) * v
                samples[si] += 1
                t[si] += dt[si]
        live_now = (t[idx] < t_end[idx]) & (alpha[idx] < 0.95)
        idx = idx[live_now]
    return leaps, samples, hit, ev


def main():
    global BR
    args = sys.argv[1:]
    if args and args[0].startswith("--brick="):
        BR = int(args.pop(0).split("=")[1])
    frames = [int(a) for a in args] or [0, 45, 100]
    vol = volumes.xor_u8(N).astype(np.float64)
    occ = occupancy(vol.astype(np.uint8))
    nbk = N // BR
    print(f"bricks occupied: {occ.mean():.3f}")
    iso = chebyshev(occ)
    octs = {}
    for k in range(8):
        sx, sy, sz = (1 if k & 1 else -1), (1 if k & 2 else -1), (1 if k & 4 else -1)
        octs[k] = octant(occ, sx, sy, sz)
    oct_all = np.stack([octs[k] for k in range(8)])
    print("mean distance over empty bricks: iso %.2f, octant %.2f" % (iso[~occ].mean(), oct_all[:, ~occ].mean()))
    zz, yy, xx = np.nonzero(occ)
    bb_lo = np.array([xx.min() * BR - 1, yy.min() * BR - 1, zz.min() * BR - 1]) * 2.0 / N - 1
    bb_hi = np.array([xx.max() * BR + BR + 1, yy.max() * BR + BR + 1, zz.max() * BR + BR + 1]) * 2.0 / N - 1
    # every 3rd warp tile in x and y
    tx, ty = np.meshgrid(np.arange(0, W // 8, 3), np.arange(0, H // 4, 3))
    lx, ly = np.meshgrid(np.arange(8), np.arange(4))
    px = (tx.reshape(-1, 1) * 8 + lx.reshape(1, -1)).reshape(-1).astype(np.float64)
    py = (ty.reshape(-1, 1) * 4 + ly.reshape(1, -1)).reshape(-1).astype(np.float64)
    for fr in frames:
        yaw = 1.0 + 2.0 * math.pi * (fr % ORBIT) / ORBIT
        inv = camera_inv(3.0, -0.5, yaw, W / H)
        eye, d = rays(inv, px, py)
        okey = (d[:, 0] > 0) * 1 + (d[:, 1] > 0) * 2 + (d[:, 2] > 0) * 4
        res = {}
        for name, fn in (("iso", lambda bz, by, bx, idx: iso[bz, by, bx]),
                         ("octant", lambda bz, by, bx, idx: oct_all[okey[idx], bz, by, bx])):
            leaps, samples, hit, ev = march(vol, occ, fn, eye, d, bb_lo, bb_hi)
            it = (leaps + samples).reshape(-1, 32)
            warp_iters = it.max(axis=1)
            # warp-level events under lock-step: approximated by per-lane counts (upper bound = max, lower = mean)
            res[name] = (leaps.sum(), samples.sum(), warp_iters.sum(), leaps.reshape(-1, 32).max(axis=1).sum(),
                         samples.reshape(-1, 32).max(axis=1).sum())
            # instruction model of the kernel's loop: head + tail per iteration, leap path, sample path
            cost = ev[0] * 29 + ev[1] * 30 + ev[2] * 48  # issue slots of the final build: profiles/loop_sass_r02.md (head + tail, leap + rejoin, sample)
            print(f"frame {fr:3d} {name:7s}: lock-step warp iterations {ev[0]} with-leap {ev[1]} with-sample {ev[2]}  model cost {cost / 1e6:.2f} M warp-inst (x9 per frame)")
            print(f"frame {fr:3d} {name:7s}: lane leaps {leaps.sum():9d}  lane samples {samples.sum():9d}  "
                  f"sum over warps of max-lane iterations {warp_iters.sum():8d}  (max-lane leaps {res[name][3]}, max-lane samples {res[name][4]})")
        a, b = res["iso"], res["octant"]
        print(f"   octant/iso: lane leaps {b[0] / a[0]:.3f}  warp iterations {b[2] / a[2]:.3f}")


if __name__ == "__main__":
    main()

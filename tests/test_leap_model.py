"""CPU tier: the leap arithmetic of the raycast kernels, restated operation by operation in IEEE float32 (numpy
scalars round every operation to nearest even exactly like __fmul_rn / __fadd_rn / __fdiv_rn), walked against the
brute-force sample sequence t = t + dt on grids up to 4096^3 — sizes at which a ray takes ~10^4 steps and the
systematic rounding of the repeated additions amounts to several steps, which no GPU test of this repository
reaches with an oracle beside it.

Checked for every ray:
  * raycast.cu `leap_count`: every sample a leap passes over lies in an EMPTY brick (distance >= 1), so skipping
    it is a bit-exact no-op, and the leap lands on a member of the brute-force t sequence;
  * raycast.cu occupied-bounds clip: the samples before the entry leap's landing point and the samples at or
    after t_end lie outside the occupied bricks;
  * sortlast.cu pre-leap (`n_pre`) and `t_stop`: no sample whose voxel index falls in the rank's own brick is
    skipped.
The distance fields are analytic: a few occupied boxes of bricks (edge B = 2, 8 or 32 voxels); the raycast's field is
DIRECTIONAL, one per ray octant — the largest d such that the d^3 bricks ahead of the brick in the ray's direction of travel
are empty, capped at 128 voxels, as built by volume.cu `octant_step_kernel` — and for M1 the outside of the grid counts as
occupied. The brick coordinate is floor(q / B) taken on the float (the kernel's FFMA.RM), not a shift of the truncated
index; the approximate reciprocals of the leap model (MUFU.RCP) are modelled by a random relative error of 2^-21.
"""
import math

import numpy as np
import pytest

F = np.float32
CAP = 16


def f32(x):
    return np.float32(x)


def fma(a, b, c):  # used only inside margins; double product of two floats is exact
    return np.float32(np.float64(a) * np.float64(b) + np.float64(c))


def trunc_i(x):
    return int(np.trunc(np.float64(x)))


class Scene:
    def __init__(self, n, boxes, mode, B=8):
        self.n, self.B, self.nb, self.boxes, self.mode = n, B, (n + B - 1) // B, boxes, mode  # boxes: [(lo3, hi3)] in bricks, hi exclusive
        self.cap = max(4, 128 // B)

    def occupied(self, b):
        if self.mode == 1 and not all(0 <= b[k] < self.nb for k in range(3)):
            return True
        return any(all(lo[k] <= b[k] < hi[k] for k in range(3)) for lo, hi in self.boxes)

    def dist(self, b, sg=None):
        """sg = None: the isotropic field of the sort-last windows (volume.cu distance_step_kernel): 0 = occupied, else
        min(CAP, Chebyshev distance to the nearest occupied brick). sg = (+-1, +-1, +-1): the raycast's directional field
        (octant_step_kernel): the largest d <= cap such that the cube of d^3 bricks [b, b + sg d) holds no occupied brick.
        M1: bricks outside the grid are occupied."""
        if sg is None:
            best = CAP
            for lo, hi in self.boxes:
                d = max(max(lo[k] - b[k], 0, b[k] - (hi[k] - 1)) for k in range(3))
                best = min(best, d)
            if self.mode == 1:
                for k in range(3):
                    best = min(best, b[k] + 1, self.nb - b[k])
            return best
        if self.occupied(b):
            return 0
        d = self.cap
        # the cube's extent per axis is [b, b + d - 1] or [b - d + 1, b]; shrink d until it clears every box (and the grid, M1)
        for lo, hi in self.boxes:
            # largest d for which the cube misses this box: the box is missed iff on SOME axis the cube ends before it
            best_axis = 0
            for k in range(3):
                if sg[k] > 0:
                    room = lo[k] - b[k] if b[k] < lo[k] else (self.cap if b[k] >= hi[k] else 0)
                else:
                    room = b[k] - (hi[k] - 1) if b[k] >= hi[k] else (self.cap if b[k] < lo[k] else 0)
                best_axis = max(best_axis, room)
            d = min(d, best_axis)
        if self.mode == 1:
            for k in range(3):
                d = min(d, self.nb - b[k] if sg[k] > 0 else b[k] + 1)
        return d

    def occ_bounds(self):
        lo = [min(bx[0][k] for bx in self.boxes) for k in range(3)]
        hi = [max(bx[1][k] for bx in self.boxes) - 1 for k in range(3)]
        return lo, hi


def make_ray(rng, inside=False, aim=None):
    """An eye and a unit direction (float32, normalised with float32 operations) aimed at the box [-1,1]^3
    (or near the point `aim`)."""
    while True:
        eye = rng.uniform(-0.9, 0.9, 3) if inside else rng.normal(size=3)
        if not inside:
            eye = eye / np.linalg.norm(eye) * rng.uniform(1.9, 4.0)
        tgt = rng.uniform(-0.8, 0.8, 3) if aim is None else np.asarray(aim) + rng.uniform(-0.05, 0.05, 3)
        d = (tgt - eye).astype(np.float32)
        ln = np.sqrt(f32(f32(f32(d[0] * d[0]) + f32(d[1] * d[1])) + f32(d[2] * d[2])))
        d = np.array([f32(d[0] / ln), f32(d[1] / ln), f32(d[2] / ln)], np.float32)
        if np.all(np.abs(d) > 1e-4):
            return eye.astype(np.float32), d


def intersect(eye, d):
    inv = [f32(F(1) / d[k]) for k in range(3)]
    a = [f32(f32(F(-1) - eye[k]) * inv[k]) for k in range(3)]
    b = [f32(f32(F(1) - eye[k]) * inv[k]) for k in range(3)]
    t0 = max(min(a[0], b[0]), max(min(a[1], b[1]), min(a[2], b[2])))
    t1 = min(max(a[0], b[0]), min(max(a[1], b[1]), max(a[2], b[2])))
    return t0, t1, inv


def step_dt(d, n, dt_scale, dt_floor):
    v = [f32(F(1) / f32(F(n) * abs(d[k]))) for k in range(3)]
    ref = f32(F(dt_scale) * max(min(v[0], min(v[1], v[2])), F(dt_floor)))  # raycast_compute.wgsl:65-68, three divisions
    a = max(f32(F(n) * abs(d[0])), max(f32(F(n) * abs(d[1])), f32(F(n) * abs(d[2]))))
    one = f32(F(dt_scale) * max(f32(F(1) / a), F(dt_floor)))  # vkrt_device.cuh step_dt: ONE division by the largest divisor
    assert one == ref  # correctly rounded division is monotone: the same float
    return one


def position(eye, d, t, h):
    p = [f32(eye[k] + f32(t * d[k])) for k in range(3)]
    q = [f32(f32(p[k] + F(1)) * h) for k in range(3)]
    return q, [trunc_i(x) for x in q]


def brick_of(sc, q):
    """floor(q / B) on the float, as the kernel's FFMA.RM(q, 1 / B, 1.5 * 2^23) computes it (1 / B is a power of two)"""
    return [int(math.floor(float(q[k]) / sc.B)) for k in range(3)]


def leap_count(sc, d_brick, brick, q, rq, drift, eps):
    """raycast.cu leap_count, operation by operation: R = B d - (B/2 + eps); per axis h = sg*R + (B/2 - q) (one FMA),
    w = B*b + h (one FMA), s = w * rq; the grid clip only where a partial last brick sticks out (M1);
    n = max(trunc(fma(min(s), 1 - drift, 0.98)), 1) with min(s) capped at 4094."""
    B = sc.B
    R = fma(F(B), F(d_brick), f32(-(F(B / 2) + eps)))
    keep = f32(F(1) - drift)
    s = []
    for k in range(3):
        sg = F(1) if rq[k] >= 0 else F(-1)
        h = fma(sg, R, f32(F(B / 2) - q[k]))
        w = fma(F(B), F(brick[k]), h)
        sk = f32(w * rq[k])
        if sc.mode == 1 and (sc.n & (B - 1)) != 0 and sg > 0:
            sk = min(sk, f32(f32(f32(F(sc.n) - eps) - q[k]) * rq[k]))
        s.append(sk)
    sm = min(min(s[0], s[1]), min(s[2], F(4094)))
    return max(trunc_i(fma(sm, keep, F(0.98))), 1)


def walk(sc, eye, d, dt_scale, dt_floor, max_steps=40000):
    """Returns the brute-force t sequence and, for each sample, (index, in-grid, brick distance or None)."""
    n, h = sc.n, F(sc.n / 2)
    t0, t1, inv = intersect(eye, d)
    assert t0 < t1
    t0 = max(t0, F(0))
    dt = step_dt(d, n, dt_scale, dt_floor)
    ts, t = [], t0
    while t < t1 and len(ts) < max_steps:
        ts.append(t)
        t = f32(t + dt)
    assert len(ts) < max_steps
    return t0, t1, inv, dt, h, ts


def sample_info(sc, eye, d, t, h, sg=None):
    """dist: the field's value at the sample's brick (directional for the ray's signs `sg`, else the isotropic one)"""
    q, idx = position(eye, d, t, h)
    inb = all(0 <= idx[k] < sc.n for k in range(3))
    dist = sc.dist([i // sc.B for i in idx], sg) if inb else None
    return q, idx, inb, dist


def runs_ahead(dt, lo_exp=1):
    """True if the repeated addition of dt runs AHEAD of t0 + j*dt in the binade [2^lo_exp, 2^(lo_exp+1)): dt in
    ulps of that binade has a fractional part just above one half, so every step rounds up by almost half an ulp."""
    ulp = 2.0 ** (lo_exp - 23)
    frac = (float(dt) / ulp) % 1.0
    return 0.5 < frac < 0.62


def adversarial_ray(rng, n, dt_scale, tries=4000):
    """A ray from beyond the low corner into the far octant whose step length makes the t sequence run ahead
    fastest while t is in [2, 4): the worst case for every leap that counts steps on the ideal line."""
    for _ in range(tries):
        u = np.abs(rng.normal(size=3))
        eye = (-u / np.linalg.norm(u) * rng.uniform(3.0, 4.2)).astype(np.float32)
        dd = (rng.uniform(0.2, 0.9, 3) - eye).astype(np.float32)
        ln = np.sqrt(f32(f32(f32(dd[0] * dd[0]) + f32(dd[1] * dd[1])) + f32(dd[2] * dd[2])))
        d = np.array([f32(dd[0] / ln), f32(dd[1] / ln), f32(dd[2] / ln)], np.float32)
        if runs_ahead(step_dt(d, n, dt_scale, 0.0)):
            return eye, d
    raise AssertionError("no adversarial ray found")


def drift_in_steps(ts, dt):
    """How far the last sample of the sequence is ahead of t0 + j*dt, in steps."""
    j = len(ts) - 1
    return (float(ts[j]) - (float(ts[0]) + j * float(dt))) / float(dt)


CASES = [(256, 1.0, 0.01), (1024, 1.0, 0.0), (2048, 2.0, 0.0), (4096, 1.0, 0.0), (4096, 2.0, 0.0)]


@pytest.mark.parametrize("n,dt_scale,dt_floor", CASES)
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("B", [2, 8, 32])
def test_leaps_only_pass_over_empty_bricks(n, dt_scale, dt_floor, mode, B):
    if B == 2 and n > 1024:
        pytest.skip("the library picks coarser bricks for grids this large (api.cu occupancy_brick_shift)")
    rng = np.random.default_rng(n * 7 + mode + 31 * B)
    nb = n // B
    boxes = []
    for _ in range(3):
        lo = rng.integers(nb // 8, nb * 3 // 4, 3)
        sz = rng.integers(1, max(nb // 6, 2), 3)
        boxes.append((list(lo), list(np.minimum(lo + sz, nb - (1 if mode == 1 else 0)))))
    sc = Scene(n, boxes, mode, B)
    eps = f32(F(16.0) * F(1.1920929e-07) * F(n))
    rays = 5 if n >= 2048 else 10
    leaps = skipped = 0
    for r in range(rays):
        eye, d = make_ray(rng, inside=(r % 4 == 3))
        t0, t1, inv, dt, h, ts = walk(sc, eye, d, dt_scale, dt_floor)
        dq = [f32(f32(d[k] * h) * dt) for k in range(3)]
        # MUFU.RCP + FMUL: a relative error of up to 2^-21 either way
        rq = [f32(f32(F(1) / dq[k]) * F(1 + rng.uniform(-1, 1) * 2.0 ** -21)) if abs(dq[k]) > 1e-12 else F(1e30) for k in range(3)]
        sg = [1 if rq[k] > 0 else -1 for k in range(3)]
        drift = f32(f32(f32(t1 * F(5.9604645e-08)) / dt) * F(2 * (1 + 2.0 ** -21)))
        j = 0
        while j < len(ts):
            q, idx, inb, _ = sample_info(sc, eye, d, ts[j], h)
            brick = brick_of(sc, q)
            # M1 consults the padded tables without a bounds test (q == n lands on the pad = occupied); M0 tests the index first
            dist = sc.dist(brick, sg) if (mode == 1 or inb) else None
            if mode == 0 and not inb:
                nskip = 1
            elif dist != 0:
                nskip = leap_count(sc, dist, brick, q, rq, drift, eps)
            else:
                nskip = 0
            if nskip == 0:
                j += 1
                continue
            leaps += 1
            for jj in range(j, min(j + nskip, len(ts))):  # every sample passed over must be a no-op
                _, idx2, inb2, _ = sample_info(sc, eye, d, ts[jj], h)
                empty = inb2 and not sc.occupied([i // B for i in idx2])
                assert empty or (not inb2 and mode == 0), (n, mode, B, r, j, jj, nskip, idx2)
                skipped += 1
            j += nskip
    assert leaps > 0 and skipped > leaps


def test_directional_distance_matches_brute_force():
    """Scene.dist(b, sg) — the analytic stand-in for octant_step_kernel's relaxation — against the definition."""
    rng = np.random.default_rng(3)
    for mode in (0, 1):
        sc = Scene(64, [([2, 3, 1], [4, 5, 3]), ([9, 9, 9], [12, 10, 14])], mode, 4)
        for _ in range(300):
            b = [int(v) for v in rng.integers(0, sc.nb, 3)]
            sg = [int(v) for v in rng.choice([-1, 1], 3)]
            d = 0
            while d < sc.cap and not any(sc.occupied([b[0] + sg[0] * i, b[1] + sg[1] * j, b[2] + sg[2] * k])
                                         for i in range(d + 1) for j in range(d + 1) for k in range(d + 1)):
                d += 1
            assert sc.dist(b, sg) == d, (mode, b, sg, d, sc.dist(b, sg))


@pytest.mark.parametrize("n,dt_scale,dt_floor", CASES)
def test_occupied_bounds_clip_drops_only_empty_samples(n, dt_scale, dt_floor):
    """raycast.cu: t_end and the entry leap n0 against the box of occupied bricks grown by one voxel (api.cu)."""
    rng = np.random.default_rng(n + 99)
    nb = n // 8
    lo = list(rng.integers(nb // 4, nb // 2, 3))
    hi = [int(l + rng.integers(1, nb // 4)) for l in lo]
    sc = Scene(n, [(lo, hi)], 1)
    olo, ohi = sc.occ_bounds()
    bb_lo = [f32(2.0 * (olo[k] * 8.0 - 1.0) / n - 1.0) for k in range(3)]
    bb_hi = [f32(2.0 * (min((ohi[k] + 1) * 8.0, n) + 1.0) / n - 1.0) for k in range(3)]
    entered = 0
    centre = [(float(bb_lo[k]) + float(bb_hi[k])) / 2 for k in range(3)]
    for r in range(6 if n >= 2048 else 12):
        eye, d = make_ray(rng, aim=centre if r % 3 else None)  # two in three go through the occupied box
        t0, t1, inv, dt, h, ts = walk(sc, eye, d, dt_scale, dt_floor)
        drift = f32(f32(f32(t1 * F(5.9604645e-08)) / dt) * F(2))
        a = [f32(f32(bb_lo[k] - eye[k]) * inv[k]) for k in range(3)]
        b = [f32(f32(bb_hi[k] - eye[k]) * inv[k]) for k in range(3)]
        tb0 = max(min(a[0], b[0]), max(min(a[1], b[1]), min(a[2], b[2])))
        tb1 = min(max(a[0], b[0]), min(max(a[1], b[1]), max(a[2], b[2])))
        t_end = min(t1, tb1) if tb0 < tb1 else F(-1)
        n0 = 0
        if tb0 > t0 and tb0 < t_end:
            s = min(f32(f32(tb0 - t0) * f32(F(1) / dt)), F(1.0e6))
            n0 = max(trunc_i(f32(s - fma(s, drift, F(2)))), 0)
            entered += 1
        for j, t in enumerate(ts):
            if j < n0 or not (t < t_end):  # passed over by the entry leap / beyond the exit (or the ray misses the box)
                _, idx, inb, dist = sample_info(sc, eye, d, t, h)
                assert (not inb) or dist >= 1, (n, r, j, n0, idx, dist)
        if n0:
            assert n0 < len(ts)
    assert entered > 0


@pytest.mark.parametrize("n,dt_scale", [(1024, 2.0), (4096, 2.0), (4096, 1.0), (8192, 1.0)])
def test_sortlast_pre_leap_never_skips_an_own_sample(n, dt_scale):
    """sortlast.cu: n_pre (with the drift allowance) and t_stop = tx + 2 dt for the far octant of a 2x2x2 partition."""
    rng = np.random.default_rng(n + 5)
    own_lo, own_hi = [n // 2] * 3, [n] * 3
    sc = Scene(n, [([0, 0, 0], [1, 1, 1])], 1)
    h = F(n / 2)
    checked = 0
    worst = 0.0
    for r in range(6):
        if n >= 4096 and r < 4:  # look from the low corner, with a step length that drifts ahead fastest
            eye, d = adversarial_ray(rng, n, dt_scale)
        else:
            eye, d = make_ray(rng)
            eye = -np.abs(eye)  # look from the low corner: long pre-leaps to the far octant
            d = np.abs(d).astype(np.float32)
        t0, t1, inv, dt, h, ts = walk(sc, eye, d, dt_scale, 0.0)
        worst = max(worst, drift_in_steps(ts, dt))
        te, tx = t0, t1
        for k in range(3):
            lo = f32(f32(F(own_lo[k]) / h) - F(1))
            hi = f32(f32(F(own_hi[k]) / h) - F(1))
            a, b = f32(f32(lo - eye[k]) / d[k]), f32(f32(hi - eye[k]) / d[k])
            te, tx = max(te, min(a, b)), min(tx, max(a, b))
        if not (f32(te - f32(F(2) * dt)) <= f32(tx + f32(F(2) * dt))):
            continue
        s_pre = f32(f32(te - t0) / dt)
        drift = f32(f32(f32(t1 * F(5.9604645e-08)) / dt) * F(2))
        n_pre = max(trunc_i(min(f32(s_pre - fma(s_pre, drift, F(2))), F(1.0e9))), 0)
        t_stop = min(t1, f32(tx + f32(F(2) * dt)))
        for j, t in enumerate(ts):
            if j < n_pre or not (t < t_stop):
                _, idx, _, _ = sample_info(sc, eye, d, t, h)
                c = [min(max(i, 0), n - 1) for i in idx]
                mine = all(own_lo[k] <= c[k] < own_hi[k] for k in range(3))
                assert not mine, (n, r, j, n_pre, idx)
        checked += int(n_pre > 0)
    assert checked > 0
    if n >= 8192:
        assert worst > 4.0  # the sequences really ran ahead by more than the fixed 2-sample allowance


def test_entry_leap_keeps_the_drift_in_hand_at_8192():
    """The occupied-bounds entry leap after ~10^4 steps of a sequence that runs ahead by several steps (8192^3,
    per-voxel steps, rays from beyond the low corner into a box in the far octant): the landing point must still
    lie in front of the box. Without the drift term in n0 this fails."""
    n = 8192
    nb = n // 8
    rng = np.random.default_rng(11)
    sc = Scene(n, [([int(nb * 0.62)] * 3, [int(nb * 0.95)] * 3)], 1)
    olo, ohi = sc.occ_bounds()
    bb_lo = [f32(2.0 * (olo[k] * 8.0 - 1.0) / n - 1.0) for k in range(3)]
    bb_hi = [f32(2.0 * (min((ohi[k] + 1) * 8.0, n) + 1.0) / n - 1.0) for k in range(3)]
    entered, worst = 0, 0.0
    for r in range(5):
        eye, d = adversarial_ray(rng, n, 1.0)
        t0, t1, inv, dt, h, ts = walk(sc, eye, d, 1.0, 0.0)
        drift = f32(f32(f32(t1 * F(5.9604645e-08)) / dt) * F(2))
        a = [f32(f32(bb_lo[k] - eye[k]) * inv[k]) for k in range(3)]
        b = [f32(f32(bb_hi[k] - eye[k]) * inv[k]) for k in range(3)]
        tb0 = max(min(a[0], b[0]), max(min(a[1], b[1]), min(a[2], b[2])))
        tb1 = min(max(a[0], b[0]), min(max(a[1], b[1]), max(a[2], b[2])))
        t_end = min(t1, tb1) if tb0 < tb1 else F(-1)
        if not (tb0 > t0 and tb0 < t_end):
            continue
        s = min(f32(f32(tb0 - t0) * f32(F(1) / dt)), F(1.0e6))
        n0 = max(trunc_i(f32(s - fma(s, drift, F(2)))), 0)
        entered += 1
        worst = max(worst, (float(ts[n0]) - (float(ts[0]) + n0 * float(dt))) / float(dt))
        for j in range(max(n0 - 40, 0), n0):  # the last samples the leap passes over are the ones at risk
            _, idx, inb, dist = sample_info(sc, eye, d, ts[j], h)
            assert (not inb) or dist >= 1, (r, j, n0, idx, dist)
        # and the leap is not wastefully short either: it ends within 2 + 3 * allowance steps of the box
        first = next(j for j in range(n0, len(ts)) if sample_info(sc, eye, d, ts[j], h)[3] == 0)
        assert first - n0 <= 2 + 3 * float(s) * float(drift) + 12, (first, n0)
    assert entered >= 3 and worst > 2.0

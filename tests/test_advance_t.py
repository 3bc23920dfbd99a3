"""The closed-form replacement of n repeated fp32 additions (advance_t in vkrt_device.cuh), restated in
Python and checked against literal repeated addition. The CUDA function itself is covered on the GPU by
test_empty_space_skipping_is_bit_exact (frames and iteration counts identical with skipping on/off)."""
import numpy as np
import pytest


def advance_t(t: np.float32, dt: np.float32, n: int) -> np.float32:
    db = int(np.float32(dt).view(np.uint32))
    e_d, M = db >> 23, (db & 0x7FFFFF) | 0x800000
    t = np.float32(t)
    while n > 0:
        tb = int(t.view(np.uint32))
        e = tb >> 23
        shift = e - e_d
        if shift < 0 or e_d == 0 or e == 0:
            t = np.float32(t + dt); n -= 1
            continue
        if shift > 24:
            return t
        if shift == 0:
            inc = M
        else:
            m, rem, half = M >> shift, M & ((1 << shift) - 1), 1 << (shift - 1)
            if rem == half:
                t = np.float32(t + dt); n -= 1
                continue
            inc = m + (1 if rem > half else 0)
        T = (tb & 0x7FFFFF) | 0x800000
        room = (0xFFFFFF - T) // inc if inc else 0
        k = min(n, room)
        if k == 0:
            t = np.float32(t + dt); n -= 1
            continue
        t = np.uint32(tb + k * inc).view(np.float32)
        n -= k
    return t


def leap_t(t, dt, n):
    """The uncached form of leap_cached's fast path: stays inside the binade and no tie -> one integer multiply-add."""
    db, tb = int(np.float32(dt).view(np.uint32)), int(np.float32(t).view(np.uint32))
    e, shift = tb >> 23, (tb >> 23) - (db >> 23)
    if 1 <= shift <= 24 and (db >> 23) != 0:
        M = (db & 0x7FFFFF) | 0x800000
        rem, half = M & ((1 << shift) - 1), 1 << (shift - 1)
        if rem != half:
            nb = tb + n * ((M >> shift) + (1 if rem > half else 0))  # 64-bit in the kernel: must not wrap
            if (nb >> 23) == e:
                return np.uint32(nb).view(np.float32)
    return advance_t(t, dt, n)


def binade_inc(e, dt):
    db = int(np.float32(dt).view(np.uint32))
    ed, shift = db >> 23, e - (db >> 23)
    if shift < 1 or shift > 24 or ed == 0:
        return 0xFFFFFFFF
    M = (db & 0x7FFFFF) | 0x800000
    rem, half = M & ((1 << shift) - 1), 1 << (shift - 1)
    if rem == half:
        return 0xFFFFFFFF
    return (M >> shift) + (1 if rem > half else 0)


def leap_cached(t, dt, n, cache):
    """leap_cached in vkrt_device.cuh: the per-binade increment lives in `cache` = [e, inc] across a ray's leaps."""
    tb = int(np.float32(t).view(np.uint32))
    e = tb >> 23
    if e != cache[0]:
        cache[0], cache[1] = e, binade_inc(e, dt)
    nb = tb + n * cache[1]
    if (nb >> 23) == e:
        return np.uint32(nb).view(np.float32)
    return advance_t(t, dt, n)


@pytest.mark.parametrize("seed", range(6))
def test_closed_form_equals_repeated_addition(seed):
    rng = np.random.default_rng(seed)
    cases = [(np.float32(1.5458316), np.float32(0.01), 277), (np.float32(0.0), np.float32(0.01), 300),
             (np.float32(0.031), np.float32(0.01), 40), (np.float32(2.7), np.float32(1 / 4096), 5000)]
    for _ in range(60):
        t0 = np.float32(rng.uniform(0, 6) if rng.uniform() < 0.8 else rng.uniform(0, 0.05))
        dt = np.float32(10 ** rng.uniform(-4, -1.5))
        cases.append((t0, dt, int(rng.integers(1, 3000))))
    for t0, dt, n in cases:
        ref = np.float32(t0)
        for _ in range(n):
            ref = np.float32(ref + dt)
        got = advance_t(t0, dt, n)
        assert got.view(np.uint32) == ref.view(np.uint32), (t0, dt, n, got, ref)
        got2 = leap_t(t0, dt, n)
        assert got2.view(np.uint32) == ref.view(np.uint32), ("leap_t", t0, dt, n, got2, ref)


@pytest.mark.parametrize("seed", range(4))
def test_cached_closed_form_along_a_ray(seed):
    """A ray alternates leaps of random length and single steps; the cache persists across them."""
    rng = np.random.default_rng(100 + seed)
    for _ in range(25):
        t = ref = np.float32(rng.uniform(0, 3) if rng.uniform() < 0.7 else 0.0)
        dt = np.float32(0.01) if rng.uniform() < 0.5 else np.float32(10 ** rng.uniform(-4, -1.5))
        cache = [0xFFFFFFFF, 0xFFFFFFFF]
        for _ in range(12):
            n = int(rng.integers(1, 400))
            for _ in range(n):
                ref = np.float32(ref + dt)
            t = leap_cached(t, dt, n, cache)
            assert t.view(np.uint32) == ref.view(np.uint32), (t, ref, dt, n)
            t = ref = np.float32(ref + dt)  # an ordinary sample step in between


def test_long_leap_from_a_small_t_does_not_wrap():
    """n * inc exceeds 2^32 when t sits one binade above dt and the leap is long (eye inside a 4096^3 box,
    dt = 1/4096-ish): a 32-bit product could wrap back into t's binade; the kernel multiplies in 64 bits."""
    dt = np.float32(2.4414062e-4 * 1.37)
    for t0 in (np.float32(dt * np.float32(1.01)), np.float32(dt * np.float32(1.9))):
        for n in (511, 512, 1024, 4097):
            ref = np.float32(t0)
            for _ in range(n):
                ref = np.float32(ref + dt)
            assert leap_t(t0, dt, n).view(np.uint32) == ref.view(np.uint32)
            assert leap_cached(t0, dt, n, [0xFFFFFFFF, 0xFFFFFFFF]).view(np.uint32) == ref.view(np.uint32)


# ---- leap_steps (vkrt_device.cuh): cached increment, tie binades in closed form, no integer division ---------------
def binade_inc2(e, dt):
    db = int(np.float32(dt).view(np.uint32))
    ed, shift = db >> 23, e - (db >> 23)
    if shift < 0 or shift > 24 or ed == 0 or e == 0:
        return 0xFFFFFFFF
    M = (db & 0x7FFFFF) | 0x800000
    if shift == 0:
        return M
    m, rem, half = M >> shift, M & ((1 << shift) - 1), 1 << (shift - 1)
    if rem == half:
        return 0x80000000 | (m + (m & 1))
    return m + (1 if rem > half else 0)


def leap_steps(t, dt, n, cache, stats=None):
    """vkrt_device.cuh leap_steps / leap_steps_slow. cache = [eb, inc]: eb = e << 23 of the binade whose plain increment
    (< 2^19) is cached, 0xFFFFFFFF = none. The fast path is 32-bit: n <= 4095."""
    t, dt = np.float32(t), np.float32(dt)
    tb = int(t.view(np.uint32))
    if n <= 4095:
        nb = (tb + n * cache[1]) & 0xFFFFFFFF
        if (((tb ^ cache[0]) | (nb ^ cache[0])) >> 23) == 0:
            if stats is not None:
                stats["fast"] = stats.get("fast", 0) + 1
            return np.uint32(nb).view(np.float32)
    while True:
        tb = int(t.view(np.uint32))
        e = tb >> 23
        if (e << 23) == cache[0]:
            inc = cache[1]
        else:
            inc = binade_inc2(e, dt)
            if inc < (1 << 19):
                cache[0], cache[1] = e << 23, inc
        closed = inc != 0xFFFFFFFF and not ((inc >> 31) and (tb & 1))
        if closed:
            inc &= 0x7FFFFFFF
            nb = tb + n * inc
            if (nb >> 23) == e:
                return np.uint32(nb).view(np.float32)
            room = (((e + 1) << 23) - 1) - tb
            m = int(np.float32(np.float32(room) / np.float32(inc))) - 1  # (__fdividef: within 2 ulp; the verify below covers it)
            if m >= 1 and m < n and m * inc <= room:
                t = np.uint32(tb + m * inc).view(np.float32)
                n -= m
        t = np.float32(t + dt)
        if stats is not None:
            stats["real"] = stats.get("real", 0) + 1
        n -= 1
        if n == 0:
            return t


def _tie_dt(rng, t_binade_exp):
    """A dt that is an exact tie ((m + 1/2) ulp) in the binade 2^t_binade_exp."""
    shift = int(rng.integers(1, 12))
    M = (int(rng.integers(0x800000, 0x1000000)) >> shift << shift) | (1 << (shift - 1))
    return np.uint32(((127 + t_binade_exp - shift) << 23) | (M & 0x7FFFFF)).view(np.float32)


@pytest.mark.parametrize("seed", range(6))
def test_leap_steps_equals_repeated_addition(seed):
    rng = np.random.default_rng(500 + seed)
    stats = {}
    for case in range(60):
        t = ref = np.float32(rng.uniform(0, 5) if rng.uniform() < 0.75 else (0.0 if rng.uniform() < 0.5 else rng.uniform(0, 0.05)))
        r = rng.uniform()
        if r < 0.3:
            dt = np.float32(0.01)
        elif r < 0.6:
            dt = _tie_dt(rng, int(rng.integers(0, 3)))  # exact ties somewhere in t's range [1, 8)
        else:
            dt = np.float32(10 ** rng.uniform(-4, -1.5))
        cache = [0xFFFFFFFF, 0]
        for _ in range(14):
            n = int(rng.integers(1, 12)) if rng.uniform() < 0.7 else int(rng.integers(12, 900))
            for _ in range(n):
                ref = np.float32(ref + dt)
            t = leap_steps(t, dt, n, cache, stats)
            assert t.view(np.uint32) == ref.view(np.uint32), (case, t, ref, dt, n)
            if rng.uniform() < 0.6:
                t = ref = np.float32(ref + dt)  # an ordinary sample step in between
    assert stats.get("fast", 0) > 100  # the one-multiply-add path carries most leaps


def test_leap_steps_tie_binade_is_closed_form():
    """dt = 0.0078125 * 1.5 = (m + 1/2) ulp for t in [2, 4) with a suitable mantissa: the run must not fall back to one real
    addition per step."""
    rng = np.random.default_rng(7)
    dt = _tie_dt(rng, 1)
    for t0 in (np.float32(2.0), np.float32(2.0000002), np.float32(3.1)):
        stats, cache = {}, [0xFFFFFFFF, 0]
        ref = t0
        for _ in range(60):
            ref = np.float32(ref + dt)
        got = leap_steps(t0, dt, 60, cache, stats)
        assert got.view(np.uint32) == ref.view(np.uint32)
        assert stats.get("real", 0) <= 3

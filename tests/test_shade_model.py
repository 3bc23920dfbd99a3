"""The raycast loop's M1 compositing (m1_shade_acc + m1_finish in vkrt_device.cuh) restated in float32 numpy and checked
against the reference form it replaces (m1_shade = raycast_naive.wgsl:106-117: rgb += w * (0.5 + 0.5 cos(.)), a += w).

The loop accumulates only w * cos(.) per channel; the palette's constant half, 0.5 * sum(w), is added once per ray from
the alpha gained. What this file pins on the CPU:
  * alpha — what early termination and the iteration counts depend on — goes through exactly the same operations, so
    the alpha sequence and the terminating sample are bit-identical;
  * a transparent sample (v == 0, every sample of an empty brick) leaves all four accumulators bit-identical, which is
    what exact empty-space skipping needs;
  * colours agree to ~1e-6 per ray (summation order), three orders of magnitude inside the 2/255 bar.
The CUDA functions themselves are covered on the GPU (test_gpu_parity.py / test_gpu_baseline_size.py: M1 against the oracle,
skipping on == off bit for bit)."""
import numpy as np
import pytest

F = np.float32
TAU = F(6.28318)
C = (TAU, F(TAU * F(1.7)), F(TAU * F(0.4)))
D = (F(0.0), F(TAU * F(0.15)), F(TAU * F(0.20)))


def fma(a, b, c):
    # float32 fused multiply-add: the product of two float32 is exact in float64, and one float64 addition of values
    # this small rounds at most once more before the float32 rounding below (double rounding is not an issue for the
    # tolerances asserted here; alpha never goes through fma)
    return F(np.float64(a) * np.float64(b) + np.float64(c))


def m1_alpha(s):
    x = F(min(F(0.9), F(s)))
    t = F(F(x - F(0.10)) * F(1.0 / (1.2 - 0.10)))
    t = F(min(max(t, F(0.0)), F(1.0)))
    return F(F(t * t) * fma(F(-2.0), t, F(3.0)))


def cosf(x):
    return F(np.cos(np.float64(x)))


def march_reference(samples, a0, thr):
    r = g = b = F(0.0)
    a = F(a0)
    alphas = []
    for n, s in enumerate(samples):
        v = m1_alpha(s)
        p = [fma(F(0.5), cosf(fma(C[k], v, D[k]) if k else F(C[0] * v)), F(0.5)) for k in range(3)]
        w = F(F(F(1.0) - a) * v)
        r, g, b = fma(w, p[0], r), fma(w, p[1], g), fma(w, p[2], b)
        a = F(a + w)
        alphas.append(a)
        if a >= thr:
            break
    return (r, g, b, a), alphas


def march_acc(samples, a0, thr):
    r = g = b = F(0.0)
    a = F(a0)
    alphas = []
    for n, s in enumerate(samples):
        v = m1_alpha(s)
        c = [cosf(fma(C[k], v, D[k]) if k else F(C[0] * v)) for k in range(3)]
        w = F(F(F(1.0) - a) * v)
        r, g, b = fma(w, c[0], r), fma(w, c[1], g), fma(w, c[2], b)
        a = F(a + w)
        alphas.append(a)
        if a >= thr:
            break
    half_gain = F(F(0.5) * F(a - F(a0)))
    return (fma(F(0.5), r, half_gain), fma(F(0.5), g, half_gain), fma(F(0.5), b, half_gain), a), alphas, (r, g, b)


@pytest.mark.parametrize("a0", [0.0, 0.1, 0.5])
def test_accumulated_palette_matches_the_reference_form(a0):
    rng = np.random.default_rng(7)
    worst = 0.0
    for ray in range(200):
        n = int(rng.integers(1, 400))
        kind = ray % 4
        if kind == 0:  # thin fog: hundreds of samples, small weights
            s = rng.uniform(0.1, 0.16, n)
        elif kind == 1:  # mostly empty with a few dense samples
            s = np.where(rng.random(n) < 0.9, rng.uniform(0.0, 0.1, n), rng.uniform(0.3, 1.0, n))
        elif kind == 2:  # u8 data
            s = rng.integers(0, 256, n) / 255.0
        else:  # dense: terminates within a few samples
            s = rng.uniform(0.5, 1.0, n)
        s = s.astype(np.float32)
        ref, ref_alphas = march_reference(s, a0, F(0.95))
        got, got_alphas, _ = march_acc(s, a0, F(0.95))
        assert len(ref_alphas) == len(got_alphas), "terminating sample differs"
        assert np.array_equal(np.array(ref_alphas, np.float32).view(np.uint32), np.array(got_alphas, np.float32).view(np.uint32)), "alpha sequence differs"
        for k in range(3):
            worst = max(worst, abs(float(ref[k]) - float(got[k])))
    assert worst <= 2e-5, worst  # fp16 resolution near 1 is 5e-4, the display bar 2/255 = 8e-3


def test_a_transparent_sample_is_a_bit_exact_no_op():
    rng = np.random.default_rng(11)
    s = rng.uniform(0.2, 0.6, 40).astype(np.float32)
    # the same opaque samples with transparent ones (s <= 0.1, as inside an empty brick) spliced in anywhere
    holes = np.sort(rng.integers(0, len(s) + 1, 60))
    t = np.insert(s, holes, rng.uniform(0.0, 0.1, len(holes)).astype(np.float32))
    a, a_alphas, a_raw = march_acc(s, 0.0, F(2.0))
    b, b_alphas, b_raw = march_acc(t, 0.0, F(2.0))
    assert np.array_equal(np.array(a + a_raw, np.float32).view(np.uint32), np.array(b + b_raw, np.float32).view(np.uint32))
    # and each hole on its own: v == 0 -> w == 0 -> accumulators unchanged, including the sign of a zero
    for v in (0.0, 0.05, 0.1):
        assert m1_alpha(F(v)) == F(0.0)
    assert fma(F(0.0), F(-1.0), F(0.0)).view(np.uint32) == F(0.0).view(np.uint32)  # +0 + (-0) = +0 under round-to-nearest


def test_zero_volume_and_single_sample_known_answers():
    got, _, _ = march_acc(np.zeros(50, np.float32), 0.0, F(0.95))
    assert [float(x) for x in got] == [0.0, 0.0, 0.0, 0.0]
    # one sample at the palette's v: colour = w * (0.5 + 0.5 cos(TAU (c v + d)))
    s = F(0.8)
    v = m1_alpha(s)
    got, _, _ = march_acc(np.array([s], np.float32), 0.0, F(2.0))
    for k, (c, d) in enumerate(((1.0, 0.0), (1.7, 0.15), (0.4, 0.20))):
        expect = float(v) * (0.5 + 0.5 * np.cos(6.28318 * (c * float(v) + d)))
        assert abs(float(got[k]) - expect) <= 1e-6
    assert got[3] == v

"""QUAD layout (pre-gathered 2x2 footprints, two point fetches per sample) for every storage type, against LINEAR and the oracle.

The hard GPU tier pins QUAD on u8 data (the bench headline, test_gpu_baseline_size.py, test_gpu_parity.py) and pins u8 / f16 / f32
on LINEAR and GATHER (test_m1_exact_paths_match_oracle); BASELINE config 3 runs QUAD on fp16. This file closes the cross
product: QUAD == LINEAR bit for bit (frames and iteration counts) and QUAD within 2/255 / 50 dB of the oracle for u8, f16, f32,
skipping on and off, on a grid whose dimensions are multiples of the occupancy brick (the no-clip instantiation) and on one
whose are not.

It was written after the round's GPU budget was spent, so it has not run on hardware yet. It is therefore marked xfail
(non-strict): a pass shows up as XPASS in the driver's run, a failure cannot stop the `-x` tier that the verified tests
live in; the file name sorts it after every other test file for the same reason (a device fault here cannot reach them).
Remove the marker, and the zz, once it has been seen green on a B200.
"""
import numpy as np
import pytest

from vokselis_b200 import abi

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason="added after the round's GPU minutes were spent: unverified on hardware, must not gate the verified tier")]


@pytest.fixture(scope="module")
def rt():
    from vokselis_b200 import rt as _rt

    _rt.lib()
    return _rt


def _psnr8(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else 10.0 * np.log10(255.0 ** 2 / mse)


@pytest.mark.parametrize("dims", [(64, 64, 64), (50, 37, 61)])
@pytest.mark.parametrize("dtype", [np.uint8, np.float16, np.float32])
def test_quad_equals_linear_and_matches_oracle(rt, oracle, dtype, dims):
    from vokselis_b200 import volumes

    W, H = 480, 270
    nx, ny, nz = dims
    vol8 = volumes.bonsai_standin_u8(64, seed=1, blobs=12)[:nz, :ny, :nx].copy()
    vol = vol8 if dtype == np.uint8 else (vol8.astype(np.float32) / 255.0).astype(dtype)
    cam = oracle.camera_uniform(2.0, 0.5, 1.0, (0, 0, 0), W / H)
    ref, ref_aux, _ = oracle.render(abi.default_params(abi.MODE_M1), cam, W, H, scalar=vol)
    ref8 = oracle.present(ref)
    with rt.Context(0, W, H) as ctx:
        ctx.upload_scalar(vol)
        first = None
        for skip in (0, 1):
            for layout in (abi.LAYOUT_LINEAR, abi.LAYOUT_QUAD):
                q = rt.default_params(abi.MODE_M1)
                q.skip_empty, q.count_samples, q.layout = skip, 1, layout
                ctx.set_params(q)
                ctx.render(cam)
                ctx.present()
                frame, aux, got8 = ctx.readback(), ctx.readback_aux(), ctx.readback_rgba8()
                if first is None:
                    first = (frame, aux)
                assert np.array_equal(frame, first[0]), f"frames differ (layout {layout}, skip {skip})"
                assert np.array_equal(aux, first[1]), f"iteration counts differ (layout {layout}, skip {skip})"
                assert np.array_equal(aux >> 31, ref_aux >> 31), "ray-hit mask differs from the oracle's"
                dd = aux.astype(np.int64) - ref_aux.astype(np.int64)
                assert (dd != 0).mean() <= 1e-3 and np.abs(dd).max() <= 1
                d = np.abs(got8.astype(np.int32) - ref8.astype(np.int32))
                assert d.max() <= 2 and _psnr8(got8, ref8) >= 50.0, f"max |delta| {d.max()}/255, PSNR {_psnr8(got8, ref8):.1f} dB"

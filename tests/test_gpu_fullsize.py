"""BASELINE.json's configurations at FULL size on the GPU, checked through size-independent properties: exact
empty-space skipping == full march, `tile` over the reference's offset table == `single`, frame-to-frame
determinism, LINEAR == GATHER within the parity tolerance, and reference-semantics sample counts independent of
skipping. configs[0] and configs[1] (256^3 at 1280x720 / 1920x1080) are ALSO compared with the oracle directly at
their own size in tests/test_gpu_baseline_size.py (a frame costs the oracle about a second); for configs[2]/[3]
(1024^3 fp16, 2048^3 u8 at 3840x2160: 2-8 GiB volumes, ~10^9 samples per frame) the oracle comparison runs on the
same generators at reduced size (test_gpu_parity.py::test_synthetic_configs_small_match_oracle) and the full size
is covered by the properties below."""
import numpy as np
import pytest

from vokselis_b200 import abi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rt():
    from vokselis_b200 import rt as _rt

    _rt.lib()
    return _rt


def _free_gib():
    import torch

    free, _ = torch.cuda.mem_get_info(0)
    return free / 2 ** 30


def _props(rt, ctx, mode, cam, W, H, layout):
    out = {}
    q = rt.default_params(mode)
    if mode == abi.MODE_M1:
        q.dt_scale = 2.0
    q.layout, q.count_samples = layout, 1
    frames, auxes, stats = [], [], []
    for skip in (1, 0):
        q.skip_empty = skip
        ctx.set_params(q)
        ctx.reset_stats()
        ctx.render(cam)
        frames.append(ctx.readback())
        auxes.append(ctx.readback_aux())
        stats.append(ctx.stats())
    assert np.array_equal(frames[0], frames[1]), "skipping changed the frame"
    assert np.array_equal(auxes[0], auxes[1]), "skipping changed the iteration counts"
    assert stats[0].samples_reference == stats[1].samples_reference and stats[0].rays_hit == stats[1].rays_hit
    assert stats[0].samples_fetched <= stats[1].samples_fetched == stats[1].samples_reference
    q.skip_empty, q.count_samples = 1, 0
    ctx.set_params(q)
    ctx.render(cam)
    assert np.array_equal(frames[0], ctx.readback()), "not deterministic"
    ctx.resize(W, H)
    ctx.render_tiles(cam, rt.tile_table(W, H, 256))
    assert np.array_equal(frames[0], ctx.readback()), "tile != single"
    out["frame"], out["stats"] = frames[0], stats[0]
    return out


def test_config1_and_2_reference_volume_256_at_1080p(rt, oracle):
    """configs[0]/[1] shapes: 256^3, 1280x720 and 1920x1080, M0 on the generated rgba16f pair (all layouts
    agree bit for bit: the fetch is a point fetch) and M1 on the xor u8 volume."""
    from vokselis_b200 import volumes

    for W, H in ((1280, 720), (1920, 1080)):
        cam = rt.Camera(3.0, -0.5, 1.0, (0, 0, 0), W / H).get_proj_view_matrix()
        with rt.Context(0, W, H) as ctx:
            ctx.generate_xor(256, 0)
            ref = None
            for layout in (abi.LAYOUT_LINEAR, abi.LAYOUT_BRICKED, abi.LAYOUT_TEXTURE):
                r = _props(rt, ctx, abi.MODE_M0, cam, W, H, layout)
                ref = ref if ref is not None else r["frame"]
                assert np.array_equal(ref, r["frame"]), f"layout {layout} differs"
            assert r["stats"].rays_hit == (179515 if (W, H) == (1280, 720) else 403942)  # SURVEY §8c pin (iii), BASELINE.md §3
            ctx.upload_scalar(volumes.xor_u8(256))
            a = _props(rt, ctx, abi.MODE_M1, cam, W, H, abi.LAYOUT_GATHER)["frame"]
            b = _props(rt, ctx, abi.MODE_M1, cam, W, H, abi.LAYOUT_LINEAR)["frame"]
            # the two fp32-weight paths differ in the last ulp of a tap (hardware unorm conversion vs b/255 in the
            # SM), which can move the 0.95 crossing by one sample on a rare ray: compare with the parity metric
            a8, b8 = oracle.present(a).astype(np.int32), oracle.present(b).astype(np.int32)
            assert np.abs(a8 - b8).max() <= 2
            assert (a8 != b8).mean() < 1e-3


@pytest.mark.parametrize("kind,dtype,n", [(0, np.float16, 1024), (1, np.uint8, 2048)])
def test_config3_and_4_at_4k(rt, kind, dtype, n):
    """configs[2]/[3]: 1024^3 fp16 and 2048^3 u8 (90 % empty) at 3840x2160."""
    need = n ** 3 * np.dtype(dtype).itemsize * 2.5 / 2 ** 30 + 2
    if _free_gib() < need:
        pytest.skip(f"needs {need:.0f} GiB of free device memory")
    W, H = 3840, 2160
    cam = rt.Camera(3.0, -0.5, 1.0, (0, 0, 0), W / H).get_proj_view_matrix()
    with rt.Context(0, W, H) as ctx:
        ctx.generate_synthetic(kind, dtype, n, seed=3 + kind)
        info = ctx.volume_info()
        if kind == 1:
            assert info["bricks_occupied"] < 0.12 * info["bricks_total"]
        r = _props(rt, ctx, abi.MODE_M1, cam, W, H, abi.LAYOUT_GATHER)
        assert r["stats"].rays_hit == 1615894  # 19.48 % of 3840x2160
        if kind == 1:
            assert r["stats"].samples_fetched < 0.1 * r["stats"].samples_reference

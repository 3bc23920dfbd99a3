"""CPU tier: the C-ABI library loads and exports every symbol include/vokselis_rt.h declares, the
host-side logic (camera, tile table, dispatch_optimal, params) behaves like the reference's, and the
product fails loudly without a GPU. No compute calls here."""
import ctypes as C
import math
import re
from pathlib import Path

import numpy as np
import pytest

from vokselis_b200 import abi

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def rt():
    from vokselis_b200 import rt as _rt

    _rt.lib()
    return _rt


def test_library_exports_every_declared_symbol(rt):
    header = (ROOT / "include" / "vokselis_rt.h").read_text()
    declared = set(re.findall(r"VKRT_API[^;(]*?\b(vkrt_\w+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(rt.EXPORTS)
    L = rt.lib()
    for name in declared:
        assert hasattr(L, name), name


def test_default_params_match_reference_literals(rt):
    p0 = rt.default_params(abi.MODE_M0)
    assert p0.struct_size == C.sizeof(abi.Params)
    assert (p0.dt_scale, p0.tile_size) == (1.0, 256)
    assert p0.dt_floor == pytest.approx(0.01) and p0.alpha_threshold == pytest.approx(0.95) and p0.initial_alpha == pytest.approx(0.1)
    assert list(p0.clear_color) == pytest.approx([0.023, 0.02, 0.02, 0.0])
    py = abi.default_params(abi.MODE_M0)
    assert bytes(p0) == bytes(py)
    assert bytes(rt.default_params(abi.MODE_M1)) == bytes(abi.default_params(abi.MODE_M1))


def test_dispatch_optimal(rt):
    """src/utils/mod.rs:15-18; use sites examples/xor/main.rs:232-233 (160 x 90 groups at 720p)."""
    assert rt.dispatch_optimal(1280, 8) == 160 and rt.dispatch_optimal(720, 8) == 90
    assert rt.dispatch_optimal(256, 16) == 16 and rt.dispatch_optimal(257, 16) == 17 and rt.dispatch_optimal(1, 8) == 1


def test_tile_table_is_the_references(rt):
    """examples/xor/main.rs:80-95: 6 x 3 = 18 origins at 1280x720; the x = 1280 column is off-screen."""
    t = rt.tile_table(1280, 720, 256)
    assert t.shape == (18, 2)
    assert t[0].tolist() == [0.0, 0.0] and t[5].tolist() == [1280.0, 0.0] and t[17].tolist() == [1280.0, 512.0]
    assert rt.tile_table(1920, 1080, 256).shape == (8 * 5, 2)


def test_camera_host_mirror_matches_oracle(rt, oracle):
    """The product's C++ camera (vokselis_b200/host/camera.cpp) vs the oracle's restatement of camera.rs."""
    for zoom, pitch, yaw, tgt, asp in [(3.0, -0.5, 1.0, (0, 0, 0), 16 / 9), (1.0, 0.5, 1.0, (0.5, 0.5, 0.5), 16 / 9),
                                       (7.5, 1.2, -4.0, (1, -2, 0.5), 4 / 3)]:
        got = rt.Camera(zoom, pitch, yaw, tgt, asp).get_proj_view_matrix()
        ref = oracle.camera_uniform(zoom, pitch, yaw, tgt, asp)
        assert np.allclose(got.view_position[:], ref.view_position[:], atol=1e-6)
        assert np.allclose(got.proj_view[:], ref.proj_view[:], rtol=1e-5, atol=1e-6)
        assert np.allclose(got.inv_proj[:], ref.inv_proj[:], rtol=2e-4, atol=2e-5)


def test_camera_setters_clamp_like_reference(rt):
    cam = rt.Camera(3.0, -0.5, 1.0)
    assert cam.updated is False  # src/camera.rs:103
    cam.set_zoom(0.01)
    assert cam.zoom == pytest.approx(0.3) and cam.updated
    cam.set_zoom(1e6)
    assert cam.zoom == pytest.approx(50.0)
    cam.set_pitch(10.0)
    assert cam.pitch < math.pi / 2 and cam.pitch == pytest.approx(math.pi / 2, abs=1e-6)
    cam.add_yaw(0.25)
    assert cam.yaw == pytest.approx(1.25)
    cam.set_aspect(1920, 1080)
    assert cam.aspect == pytest.approx(16 / 9)


def test_pipeline_entry_points(rt):
    rt.RaycastPipeline("single")
    rt.RaycastPipeline("tile")
    with pytest.raises(rt.VokselisError):
        rt.RaycastPipeline("cs_main")


def test_no_gpu_fails_loudly_not_silently(rt):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(rt.VokselisError) as e:
        rt.Context(0, 64, 64)
    assert e.value.code == abi.ERR_CUDA and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    for path in (ROOT / "vokselis_b200").rglob("*"):
        if path.suffix in (".py", ".cu", ".cuh", ".cpp", ".hpp") and path.is_file():
            text = path.read_text()
            assert "vk_oracle" not in text and "from oracle" not in text and "import oracle" not in text, path
    assert "oracle" not in (ROOT / "Makefile").read_text().split("oracle:")[0]


def test_synthetic_volumes_are_deterministic():
    from vokselis_b200 import volumes

    a = volumes.xor_u8(64)
    assert a.shape == (64, 64, 64) and a.dtype == np.uint8 and np.array_equal(a, volumes.xor_u8(64))
    b = volumes.bonsai_standin_u8(32, seed=1, blobs=8)
    assert np.array_equal(b, volumes.bonsai_standin_u8(32, seed=1, blobs=8)) and b.max() == 255
    h = volumes.hash_noise(16, 3)
    assert h.dtype == np.float16 and 0.3 < float(h.astype(np.float32).mean()) < 0.7


def _as_rt_cam(rt, ocam):
    cam = abi.CameraUniform()
    C.memmove(C.byref(cam), C.byref(ocam), C.sizeof(cam))
    return cam


def test_box_screen_bounds_contain_every_hit_pixel(rt, oracle):
    """The launch builds no ray for pixels outside vkrt_box_screen_bounds: every pixel the oracle's slab test
    (raycast_compute.wgsl:42-53,117-118) accepts must lie inside it — orbit cameras far, near, grazing and
    inside the box, odd frame shapes, fractional tile offsets, and the identity (orthographic) matrices the
    reference's camera buffer starts with (src/camera.rs:13-21)."""
    rng = np.random.default_rng(7)
    cases = [(3.0, -0.5, 1.0, (0, 0, 0)), (2.0, 0.5, 1.0, (0, 0, 0)), (1.45, 0.1, 0.3, (0, 0, 0)), (0.6, 0.2, 2.0, (0, 0, 0)),
             (5.0, 1.5, -2.0, (0.5, -0.5, 0.2)), (3.0, 0.0, 0.0, (4.0, 0.0, 0.0)), (50.0, -1.0, 4.0, (0, 0, 0))]
    cases += [(float(rng.uniform(0.3, 8)), float(rng.uniform(-1.5, 1.5)), float(rng.uniform(-6, 6)),
               tuple(rng.uniform(-1.5, 1.5, 3))) for _ in range(40)]
    culled_some = hull_culled = 0
    for k, (zoom, pitch, yaw, tgt) in enumerate(cases):
        W, H = [(160, 90), (128, 128), (97, 211)][k % 3]
        ocam = oracle.camera_uniform(zoom, pitch, yaw, tgt, W / H)
        offx, offy = [(0.0, 0.0), (0.5, 0.25), (37.0, 11.75)][k % 3]
        r = oracle.rays(ocam, W, H, offx, offy)
        hit = r[..., 6] < r[..., 7]
        (x0, y0, x1, y1), row = rt.box_screen_bounds(_as_rt_cam(rt, ocam), W, H)
        ys, xs = np.nonzero(hit)
        if len(xs):
            cx, cy = xs + offx, ys + offy
            assert cx.min() >= x0 and cx.max() <= x1 and cy.min() >= y0 and cy.max() <= y1, (zoom, pitch, yaw, tgt)
        if x0 > -1e30:
            # not grossly loose either: the margin is two pixels around the box's projected extent
            inside = (np.arange(W)[None, :] + offx >= x0) & (np.arange(W)[None, :] + offx <= x1) & \
                     (np.arange(H)[:, None] + offy >= y0) & (np.arange(H)[:, None] + offy <= y1)
            if len(xs) and xs.min() > 4 and xs.max() < W - 5 and ys.min() > 4 and ys.max() < H - 5:
                assert x0 >= cx.min() - 4 and x1 <= cx.max() + 4 and y0 >= cy.min() - 4 and y1 <= cy.max() + 4
            culled_some += int((~inside).any())
            assert row == -1 or 0 <= row < H
        # the silhouette's convex hull (tested inside the rectangle by the launch): contains every hit pixel, ...
        planes = rt.box_screen_hull(_as_rt_cam(rt, ocam), W, H).astype(np.float64)
        gx, gy = np.meshgrid(np.arange(W) + offx, np.arange(H) + offy)
        dmin = np.min(planes[:, 0, None, None] * gx[None] + planes[:, 1, None, None] * gy[None] + planes[:, 2, None, None], axis=0)
        if len(xs):
            assert dmin[hit].min() >= 0.0, (zoom, pitch, yaw, tgt, dmin[hit].min())
        if x0 > -1e30 and not np.allclose(planes, [[0, 0, 1]] * 6):
            # ... is tight (every pixel 4 px or more inside the hull hits), and culls pixels the rectangle keeps
            assert hit[dmin >= 4.0].all(), (zoom, pitch, yaw, tgt)
            hull_culled += int(((dmin < 0) & inside).sum() > 0)
    assert culled_some >= 10 and hull_culled >= 10
    # identity matrices: orthographic rays along +z through (sx, sy, 0); the box covers |sx| <= 1 and |sy| <= 1
    ident = abi.CameraUniform()
    for i in range(4):
        ident.inv_proj[5 * i] = 1.0
        ident.proj_view[5 * i] = 1.0
    r = oracle.rays(ident, 64, 36)
    hit = r[..., 6] < r[..., 7]
    (x0, y0, x1, y1), _ = rt.box_screen_bounds(ident, 64, 36)
    ys, xs = np.nonzero(hit)
    assert len(xs) and xs.min() >= x0 and xs.max() <= x1 and ys.min() >= y0 and ys.max() <= y1
    # a singular matrix disables culling
    sing = abi.CameraUniform()
    (x0, y0, x1, y1), row = rt.box_screen_bounds(sing, 64, 36)
    assert x0 < -1e30 and y1 > 1e30 and row == -1


def test_orbit_input_mapping(rt):
    """src/lib.rs:64-66,150-176: drag -> yaw/pitch at 0.0025 rad per pixel (yaw against x), wheel -> zoom at 0.002 per line
    (up = closer), motion without the button ignored; the camera's own clamps apply (src/camera.rs:115-132)."""
    import math

    cam = rt.Camera(3.0, -0.5, 1.0, (0.0, 0.0, 0.0), 16 / 9)
    inp = rt.OrbitInput()
    inp.mouse_motion(cam, 50.0, 50.0)
    assert (cam.yaw, cam.pitch) == (1.0, -0.5)
    inp.button(True)
    inp.mouse_motion(cam, 8.0, -3.0)
    assert cam.yaw == pytest.approx(1.0 - 8 * 0.0025, abs=1e-7) and cam.pitch == pytest.approx(-0.5 - 3 * 0.0025, abs=1e-7)
    inp.mouse_motion(cam, 0.0, 1.0e6)
    assert cam.pitch == pytest.approx(math.pi / 2, abs=1e-6) and cam.pitch < math.pi / 2
    inp.button(False)
    inp.mouse_wheel_lines(cam, 2.0)
    assert cam.zoom == pytest.approx(3.0 - 2 * 0.002, abs=1e-7)
    inp.mouse_wheel_pixels(cam, 1.0e6)
    assert cam.zoom == pytest.approx(0.3)
    inp.mouse_wheel_lines(cam, -1.0e6)
    assert cam.zoom == pytest.approx(50.0)
    assert cam.updated

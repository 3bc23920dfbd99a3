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

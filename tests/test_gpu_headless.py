"""The C++ host mirror above the C ABI (vokselis_b200/host/vokselis.hpp: Context, XorCompute, RaycastPipeline, Demo,
run_headless, OrbitInput) exercised on the GPU through the headless harness `build/headless` — the offscreen
counterpart of the reference's windowed `run` (src/lib.rs:45-208) driving the xor example's Demo
(examples/xor/main.rs:34-262) — and its dumps compared with the oracle on the same volume bytes and camera."""
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from vokselis_b200 import abi

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
EXE = ROOT / "build" / "headless"
W, H, FRAMES = 1280, 720, 24


@pytest.fixture(scope="module")
def rt():
    from vokselis_b200 import rt as _rt

    _rt.lib()
    return _rt


@pytest.fixture(scope="module")
def device_volume(rt):
    """The 256^3 pair of the device generator (the bytes `XorCompute::record` produces in the harness's own process too:
    same kernel, same GPU)."""
    with rt.Context(0, 64, 64) as ctx:
        ctx.generate_xor(256, 0)
        return ctx.download_rgba16f()


def _run(mode, tmp_path, frames=FRAMES):
    if not EXE.exists():
        pytest.fail("build/headless is missing: run `make` (the product must be built in-tree)")
    out = tmp_path / f"{mode}.rgba8"
    r = subprocess.run([str(EXE), str(frames), str(W), str(H), mode, str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    return np.fromfile(out, np.uint8).reshape(H, W, 4), r.stdout


def _last_orbit_yaw(frames):
    f32 = np.float32  # the harness's own float arithmetic: 1.0f + 6.2831853f * (float)i / (float)frames
    return float(f32(1.0) + f32(6.2831853) * f32(frames - 1) / f32(frames))


def _check(got8, ref8):
    d = np.abs(got8.astype(np.int32) - ref8.astype(np.int32))
    mse = np.mean(d.astype(np.float64) ** 2)
    assert d.max() <= 2 and (mse == 0 or 10 * np.log10(255.0 ** 2 / mse) >= 50.0), (d.max(), mse)


@pytest.mark.parametrize("mode", ["single", "tile", "sweep"])
def test_headless_orbit_matches_oracle(rt, oracle, device_volume, tmp_path, mode):
    color, normal = device_volume
    got8, _ = _run(mode, tmp_path)
    cam = rt.Camera(3.0, -0.5, _last_orbit_yaw(FRAMES), (0.0, 0.0, 0.0), W / H).get_proj_view_matrix()
    ref, ref_aux, _ = oracle.render(abi.default_params(abi.MODE_M0), cam, W, H, color=color, normal=normal)
    ref8 = oracle.present(ref)
    _check(got8, ref8)
    # the hit mask shows in the image: misses are exactly the presented clear colour (29/26/26), and a hit pixel only
    # matches it by coincidence
    clear8 = oracle.present(np.array([[[0x25E3, 0x251F, 0x251F, 0x3C00]]], np.uint16))[0, 0]
    hit = (ref_aux >> 31).astype(bool)
    assert (got8[~hit] == clear8).all()
    # the same frame through the ctypes mirror of the ABI: identical bytes (one library, two host layers)
    with rt.Context(0, W, H) as ctx:
        ctx.upload_rgba16f(color, normal)
        if mode == "tile":
            ctx.render_tiles(cam, rt.tile_table(W, H, 256))
        else:
            ctx.render(cam)
        ctx.present()
        assert np.array_equal(ctx.readback_rgba8(), got8)


def test_headless_drag_events_reach_the_camera(rt, oracle, device_volume, tmp_path):
    """src/lib.rs:150-176 through vokselis::OrbitInput: 40 motion events of (+8, -3) px while dragging, motion without the
    button ignored, two wheel lines: yaw 1 - 0.8, pitch -0.5 - 0.3, zoom 3 - 0.004; the frame matches the oracle's for
    exactly the camera the harness reports."""
    color, normal = device_volume
    got8, out = _run("drag", tmp_path, frames=1)
    m = re.search(r"drag: yaw (\S+) pitch (\S+) zoom (\S+)", out)
    yaw, pitch, zoom = (float(v) for v in m.groups())
    assert abs(yaw - 0.2) < 1e-5 and abs(pitch + 0.8) < 1e-5 and abs(zoom - 2.996) < 1e-5
    # the Python mirror of the mapping lands on the same camera
    cam_py = rt.Camera(3.0, -0.5, 1.0, (0.0, 0.0, 0.0), W / H)
    inp = rt.OrbitInput()
    inp.button(True)
    for _ in range(40):
        inp.mouse_motion(cam_py, 8.0, -3.0)
    inp.button(False)
    inp.mouse_motion(cam_py, 100.0, 100.0)
    inp.mouse_wheel_lines(cam_py, 2.0)
    assert abs(cam_py.yaw - yaw) < 1e-5 and abs(cam_py.pitch - pitch) < 1e-5 and abs(cam_py.zoom - zoom) < 1e-5
    cam = rt.Camera(zoom, pitch, yaw, (0.0, 0.0, 0.0), W / H).get_proj_view_matrix()
    ref, _, _ = oracle.render(abi.default_params(abi.MODE_M0), cam, W, H, color=color, normal=normal)
    _check(got8, oracle.present(ref))

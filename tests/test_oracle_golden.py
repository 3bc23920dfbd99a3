"""CPU tier: the hand-written oracle against (1) the committed golden vectors (generated from the
reference's own WGSL via oracle/_ref, see tests/golden/make_golden.py), (2) oracle/_ref itself when
it is present, and (3) analytic known answers that depend on neither (SURVEY.md §7 H1, §8c pins)."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

from vokselis_b200 import abi
from vokselis_b200.abi import CameraUniform

GOLD = Path(__file__).resolve().parent / "golden"


def cam_from(arr) -> CameraUniform:
    return CameraUniform.from_buffer_copy(arr.tobytes())


def test_g1_xor_pipeline_bit_exact(oracle):
    g = np.load(GOLD / "g1_xor32_160x90.npz")
    cam = cam_from(g["cam"])
    p = abi.default_params(abi.MODE_M0)
    color, normal = oracle.generate_xor(32, 0.0)
    assert np.array_equal(color, g["color"]) and np.array_equal(normal, g["normal"])
    f, aux, st = oracle.render(p, cam, 160, 90, color=g["color"], normal=g["normal"])
    assert np.array_equal(f, g["single"])
    p.tile_size = int(g["tile_size"])
    ft, _, _ = oracle.render(p, cam, 160, 90, color=g["color"], normal=g["normal"], offsets=g["table"])
    assert np.array_equal(ft, g["tiled"])
    assert np.array_equal(g["single"], g["tiled"])  # the reference's two entry points agree with each other
    d = np.abs(oracle.present(f).astype(int) - g["present"].astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 1e-3  # bilinear at pixel centres vs exact texel: last-bit only


def test_g2_adversarial_volume_bit_exact(oracle):
    """Non-cubic dims straddling the dt floor, NaN/inf normals, negative alpha, eye inside the box."""
    g = np.load(GOLD / "g2_random192x12x10_128x72.npz")
    p = abi.default_params(abi.MODE_M0)
    for ck, fk in (("cam", "frame"), ("cam_inside", "frame_inside")):
        f, _, st = oracle.render(p, cam_from(g[ck]), 128, 72, color=g["color"], normal=g["normal"])
        assert np.array_equal(f, g[fk]), ck
        assert st.rays_hit > 0
        assert not np.isnan(f.view(np.float16).astype(np.float32)).any()


def test_g3_generator_bit_exact(oracle):
    g = np.load(GOLD / "g3_xorgen16.npz")
    c0, n0 = oracle.generate_xor(16, 0.0)
    c1, n1 = oracle.generate_xor(16, 1.3)
    assert np.array_equal(c0, g["color_t0"]) and np.array_equal(n0, g["normal_t0"])
    assert np.array_equal(c1, g["color_t13"]) and np.array_equal(n1, g["normal_t13"])


def test_g5_naive_fragment_shader_bit_exact(oracle):
    g = np.load(GOLD / "g5_naive_fs24.npz")
    cols = oracle.naive_fs(g["vol"], g["eyes"], g["dirs"])
    assert np.array_equal(cols, g["colors"])
    assert (cols[-1] == [0, 0, 0, 1]).all()  # the miss


def test_oracle_equals_translated_reference_when_present(oracle):
    from oracle import ref_binding as rb

    if not rb.available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    n, W, H = 48, 200, 120
    color, normal = rb.xor_generate(n, 0.7)
    oc, on = oracle.generate_xor(n, 0.7)
    assert np.array_equal(color, oc) and np.array_equal(normal, on)
    for yaw in (0.3, 2.9):
        cam = oracle.camera_uniform(2.4, -0.3, yaw, (0.0, 0.1, 0.0), W / H)
        ref = rb.raycast_compute(cam, color, normal, W, H)
        got, _, _ = oracle.render(abi.default_params(0), cam, W, H, color=color, normal=normal)
        assert np.array_equal(ref, got)


def test_full_reference_configuration_oracle_equals_translated_reference(oracle):
    """The reference's own configuration end to end (SURVEY §8c pin iii): xor.wgsl `cs_main` at 256^3 with time = 0
    (examples/xor/main.rs:135-146), raycast_compute.wgsl `single` at 1280x720 with the xor example's camera
    (examples/xor/main.rs:273-279), then present.wgsl — the hand restatement against the machine-translated
    reference, bit for bit, the frozen hit count 179,515, and the SHA-256 of every stage frozen from the run in
    which the two agreed (so the pin also holds where oracle/_ref is absent; the generator's sin-hash makes the
    bytes depend on libm's sinf, like golden G3)."""
    import hashlib

    from oracle import ref_binding as rb

    def sha(a):
        return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()

    W, H = 1280, 720
    oc, on = oracle.generate_xor(256, 0.0)
    cam = oracle.camera_uniform(3.0, -0.5, 1.0, (0.0, 0.0, 0.0), W / H)
    got, aux, st = oracle.render(abi.default_params(0), cam, W, H, color=oc, normal=on)
    got8 = oracle.present(got)
    assert st.rays_hit == 179515 and int((aux >> 31).sum()) == 179515
    assert sha(oc) == "412a8de9e216d46c413e9bed8c532bc77913acd303d406fde4f46fb3a8824eb6"
    assert sha(on) == "9bf62764672d14d154796f169c246442a90dc0b18908a252d682281d09f119a8"
    assert sha(got) == "15a856ea93ef5951d051124241d8ca6dfddebc6d5972e6622f9b0c938aa5dcc1"
    assert sha(got8) == "19a79dfc790f6357980645f316ab55a2c5bad8464762a14d6db614fc8931990a"
    if not rb.available():
        return
    color, normal = rb.xor_generate(256, 0.0)
    assert np.array_equal(color, oc) and np.array_equal(normal, on)
    ref = rb.raycast_compute(cam, color, normal, W, H)
    assert np.array_equal(ref, got)
    assert np.array_equal(rb.present(ref), got8)


@pytest.mark.parametrize("seed", range(5))
def test_oracle_equals_translated_reference_randomised(oracle, seed):
    """Random volumes (NaN / inf / negative texels, non-cubic dims on both sides of the dt floor), random
    cameras (outside, grazing, inside the box) and random tile tables with FRACTIONAL offsets (the Offset is
    two f32: the ray uses gid + offset, the store goes to gid + u32(offset)): hand restatement == the
    reference's translated WGSL, bit for bit."""
    from oracle import ref_binding as rb

    if not rb.available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    rng = np.random.default_rng(100 + seed)
    nx, ny, nz = (int(v) for v in rng.choice([3, 8, 17, 40, 180, 200], size=3))
    if nx * ny * nz > 400_000:
        nz = max(2, 400_000 // (nx * ny))
    color = rng.uniform(-0.1, 1.0, size=(nz, ny, nx, 4)).astype(np.float16)
    color[..., 3] = np.where(rng.uniform(size=(nz, ny, nx)) < 0.5, 0, color[..., 3])
    normal = rng.normal(size=(nz, ny, nx, 4)).astype(np.float16)
    normal[rng.uniform(size=(nz, ny, nx)) < 0.15] = np.float16(np.nan)
    normal[0, 0, 0, 1] = np.float16(-np.inf)
    W, H = int(rng.integers(40, 120)), int(rng.integers(30, 80))
    cam = oracle.camera_uniform(float(rng.uniform(0.4, 4.0)), float(rng.uniform(-1.2, 1.2)), float(rng.uniform(-3, 3)),
                                tuple(rng.uniform(-0.3, 0.3, 3)), W / H)
    p = abi.default_params(abi.MODE_M0)
    got, _, _ = oracle.render(p, cam, W, H, color=color.view(np.uint16), normal=normal.view(np.uint16))
    ref = rb.raycast_compute(cam, color.view(np.uint16), normal.view(np.uint16), W, H)
    assert np.array_equal(got, ref)
    ts = int(rng.choice([16, 32, 48]))
    table = np.array([[x * ts + float(rng.choice([0.0, 0.25, 0.5])), y * ts + float(rng.choice([0.0, 0.75]))]
                      for y in range(H // ts + 1) for x in range(W // ts + 1)], np.float32)
    # a negative offset: the ray uses gid + offset, the store saturates to gid + 0 (vec2<u32>(f32))
    table = np.concatenate([table, np.array([[-3.5, 0.0], [float(ts), -2.25]], np.float32)])
    p.tile_size = ts
    got_t, _, _ = oracle.render(p, cam, W, H, color=color.view(np.uint16), normal=normal.view(np.uint16), offsets=table)
    ref_t = rb.raycast_compute(cam, color.view(np.uint16), normal.view(np.uint16), W, H, entry="tile", offsets=table, tile_size=ts)
    assert np.array_equal(got_t, ref_t)


def test_m1_follows_naive_march_body(oracle):
    """M1 = raycast_naive's march body on the compute boundary's rays. Render a volume with M1, then
    feed the same rays (mapped from the [-1,1] box to the naive shader's [0,1] box) to the literal
    fs_main restatement: colours agree to 1e-3 (p is accumulated in the naive shader, recomputed in M1)."""
    from vokselis_b200 import volumes

    n, W, H = 256, 64, 36
    vol = volumes.bonsai_standin_u8(32, seed=3, blobs=6).repeat(8, 0).repeat(8, 1).repeat(8, 2)
    cam = oracle.camera_uniform(2.2, 0.4, 0.8, (0, 0, 0), W / H)
    p = abi.default_params(abi.MODE_M1)
    p.dt_scale, p.m1_srgb = 2.0, 1  # box side 2 -> per-voxel steps like the naive shader
    f, aux, _ = oracle.render(p, cam, W, H, scalar=vol)
    r = oracle.rays(cam, W, H)
    eyes01 = (r[..., 0:3] + 1.0) * 0.5
    cols = oracle.naive_fs(vol, eyes01.reshape(-1, 3), r[..., 3:6].reshape(-1, 3)).reshape(H, W, 4)
    hit = (aux >> 31).astype(bool)
    got = f.view(np.float16).astype(np.float32)
    assert hit.sum() > 500
    assert np.abs(got[hit][:, :3] - cols[hit][:, :3]).max() < 4e-3


# ---- analytic pins -------------------------------------------------------------------------------
def test_struct_sizes():
    assert C.sizeof(abi.CameraUniform) == 144 and C.sizeof(abi.Uniform) == 48 and C.sizeof(abi.Offset) == 8
    assert abi.CameraUniform.inv_proj.offset == 80  # bytes 80..143 are what the shader reads


def test_camera_identities(oracle):
    cam = oracle.camera_uniform(3.0, -0.5, 1.0, (0.0, 0.0, 0.0), 16 / 9)
    pv = np.array(cam.proj_view[:], np.float64).reshape(4, 4).T  # column-major -> matrix
    inv = np.array(cam.inv_proj[:], np.float64).reshape(4, 4).T
    assert np.allclose(pv @ inv, np.eye(4), atol=2e-5)
    eye = np.array(cam.view_position[:3])
    assert np.isclose(np.linalg.norm(eye), 3.0, atol=1e-6)
    # eye = target - zoom*(sin yaw cos pitch, sin pitch, cos yaw cos pitch)
    exp = -3.0 * np.array([np.sin(1.0) * np.cos(-0.5), np.sin(-0.5), np.cos(1.0) * np.cos(-0.5)])
    assert np.allclose(eye, exp, atol=1e-6)


def test_centre_ray_passes_through_target_and_starts_on_near_plane(oracle):
    W, H = 1280, 720
    cam = oracle.camera_uniform(3.0, -0.5, 1.0, (0.0, 0.0, 0.0), W / H)
    r = oracle.rays(cam, W, H)[H // 2, W // 2]
    eye_cam = np.array(cam.view_position[:3])
    o, d = r[0:3], r[3:6]
    assert np.isclose(np.linalg.norm(o - eye_cam), 0.1, atol=1e-4)  # ZNEAR (src/camera.rs:89)
    assert np.linalg.norm(np.cross(d, -o / np.linalg.norm(o))) < 1e-4  # points at the target (origin)
    assert np.isclose(np.linalg.norm(d), 1.0, atol=1e-6)


def test_xor_camera_hit_statistics_frozen(oracle):
    """SURVEY §8c pin (iii): 179,515 hit pixels (19.48 %) at 1280x720 for the xor camera."""
    W, H = 1280, 720
    cam = oracle.camera_uniform(3.0, -0.5, 1.0, (0.0, 0.0, 0.0), W / H)
    r = oracle.rays(cam, W, H)
    hit = r[..., 6] < r[..., 7]
    assert int(hit.sum()) == 179515


@pytest.mark.parametrize("n", [256, 1024, 2048, 4096])
def test_dt_is_the_floor_for_large_volumes(oracle, n):
    """SURVEY F4 / pin (iv): dt == 0.01f for every ray when N >= 174 — checked through iteration counts:
    a zero volume never terminates early, so iterations == ceil-ish((t1 - t0)/0.01) by repeated addition."""
    W, H = 96, 54
    cam = oracle.camera_uniform(3.0, -0.5, 1.0, (0.0, 0.0, 0.0), W / H)
    r = oracle.rays(cam, W, H)
    d = np.abs(r[..., 3:6])
    with np.errstate(divide="ignore"):
        dt = np.maximum(np.min(np.float32(1.0) / (np.float32(n) * d), axis=-1), np.float32(0.01))
    assert (dt == np.float32(0.01)).all()


def test_zero_volume_gives_exact_clear_colour(oracle, xor_cam):
    n, W, H = 16, 96, 54
    z = np.zeros((n, n, n, 4), np.uint16)
    f, aux, st = oracle.render(abi.default_params(0), xor_cam, W, H, color=z, normal=z)
    expect = np.array([0.023, 0.02, 0.02, 1.0], np.float32).astype(np.float16)
    assert (f.view(np.float16) == expect).all()
    assert st.rays_hit > 0


def test_constant_alpha_closed_form(oracle, xor_cam):
    n, W, H = 16, 96, 54
    a = np.float16(0.5)
    color = np.zeros((n, n, n, 4), np.float16)
    color[..., 3] = a
    z = np.zeros((n, n, n, 4), np.uint16)
    _, aux, _ = oracle.render(abi.default_params(0), xor_cam, W, H, color=color.view(np.uint16), normal=z)
    ap = float(a) ** 3 / 0.7
    ap = ap * ap * (3 - 2 * ap)
    k = int(np.ceil(np.log(0.05 / 0.9) / np.log(1 - ap)))
    its = aux[aux >> 31 == 1] & 0x7FFFFFFF
    assert (its[its >= k] == k).all() and (its >= k).any()


def test_fp16_conversion_matches_numpy(oracle):
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.normal(size=2000).astype(np.float32) * 10.0 ** rng.integers(-9, 6, 2000),
                        np.array([0, -0.0, 65504, 65519.9, 65520, 1e-8, 5.96e-8, 2.98e-8, 2.9802322e-8, np.inf, -np.inf], np.float32)])
    L = oracle.lib()
    got = np.array([L.vko_f32_to_f16(float(v)) for v in x], np.uint16)
    with np.errstate(over="ignore"):
        assert np.array_equal(got, x.astype(np.float16).view(np.uint16))
    hs = np.arange(0, 65536, 7, dtype=np.uint16)
    back = np.array([L.vko_f16_to_f32(int(h)) for h in hs], np.float32)
    ref = hs.view(np.float16).astype(np.float32)
    assert np.array_equal(np.isnan(back), np.isnan(ref)) and np.array_equal(back[~np.isnan(ref)], ref[~np.isnan(ref)])


def test_sort_last_partials_recompose(oracle, noise64, xor_cam):
    """Simulated ranks (SURVEY §4 'multi-GPU without a cluster'): split the volume into two slabs along z,
    render each slab's samples of the GLOBAL t sequence, chain alpha front-to-back; the sum of partials
    plus the clear colour equals the single-pass frame (fp32 summation order aside)."""
    W, H = 160, 90
    color, normal = noise64
    p = abi.default_params(0)
    full, _, _ = oracle.render(p, xor_cam, W, H, color=color, normal=normal)
    n = 64
    eye_z = xor_cam.view_position[2]
    slabs = [((0, 0, 0), (n, n, n // 2)), ((0, 0, n // 2), (n, n, n))]
    if eye_z > 0:  # front-to-back for this eye
        slabs = slabs[::-1]
    a_in = np.full((H, W), p.initial_alpha, np.float32)
    rgb = np.zeros((H, W, 3), np.float32)
    for lo, hi in slabs:
        part = oracle.render_partial(p, xor_cam, W, H, lo, hi, color=color, normal=normal, a_in=a_in)
        rgb += part[..., :3]
        a_in = part[..., 3]
    r = oracle.rays(xor_cam, W, H)
    hit = r[..., 6] < r[..., 7]
    out = np.where(hit[..., None], rgb + np.array(p.clear_color[:3], np.float32), np.array(p.clear_color[:3], np.float32))
    ref = full.view(np.float16).astype(np.float32)[..., :3]
    assert np.abs(out.astype(np.float16).astype(np.float32) - ref).max() <= 2e-3


def test_present_stretched_equals_translated_reference(oracle):
    """The present pass onto a target of another size (window != backbuffer; README's volume.png is a 958x1050 capture of
    the 1280x720 backbuffer): hand restatement == the reference's present.wgsl fs_main + bilinear clamp sampler, bit for
    bit, for up- and down-scaling; at the backbuffer's own size it equals the 1:1 pass up to one LSB."""
    from oracle import ref_binding as rb

    rng = np.random.default_rng(9)
    H, W = 36, 64
    frame = rng.uniform(0.0, 2.5, size=(H, W, 4)).astype(np.float16)
    frame[..., 3] = 1.0
    f16 = frame.view(np.uint16)
    # at the backbuffer's own size uv * W - 0.5 lands on the texel up to an ulp: the bilinear blend moves a value by ~1e-7
    assert np.abs(oracle.present(f16, W, H).astype(int) - oracle.present(f16).astype(int)).max() <= 1
    if not rb.available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    for ow, oh in ((W, H), (96, 105), (29, 17), (128, 72), (1, 1)):
        assert np.array_equal(oracle.present(f16, ow, oh), rb.present(f16, ow, oh)), (ow, oh)

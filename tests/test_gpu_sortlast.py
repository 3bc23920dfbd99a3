"""Sort-last (brick-partitioned) rendering checked on ONE GPU by simulating the ranks: each brick gets
its own context on cuda:0, the two passes run brick by brick, the exchange (all-gather of
transmittances, sum of partials) is done with torch tensors. The composed frame must match the
single-context frame within the parity tolerance (2/255, 50 dB). The real multi-process path (NCCL)
is tests/mgpu_sortlast_check.py, run under torchrun on >= 2 GPUs."""
import numpy as np
import pytest

from vokselis_b200 import abi

pytestmark = pytest.mark.gpu


def psnr8(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else 10.0 * np.log10(255.0 ** 2 / mse)


def compose(rt, sortlast, torch, ctxs, cam, gn, grid, W, H, scheme="two-pass"):
    world = len(ctxs)
    n = W * H
    T_all = torch.empty(world * n, dtype=torch.float32, device="cuda")
    ain = torch.empty(n, dtype=torch.float32, device="cuda")
    parts = [torch.empty(n * 4, dtype=torch.float32, device="cuda") for _ in ctxs]
    total = torch.zeros(n * 4, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    for r, c in enumerate(ctxs):
        if scheme == "two-pass":
            c.partial_alpha(cam, T_all.data_ptr() + r * n * 4)
        else:
            c.partial_relative(cam, parts[r].data_ptr(), T_all.data_ptr() + r * n * 4)
        c.sync()
    order = sortlast.visibility_order(tuple(cam.view_position[:3]), gn, grid)
    remarched = 0
    for r, c in enumerate(ctxs):
        if scheme == "two-pass":
            c.partial_ain(T_all.data_ptr(), order[: order.index(r)], ain.data_ptr())
        else:
            c.partial_resolve(T_all.data_ptr(), order[: order.index(r)], parts[r].data_ptr(), ain.data_ptr())
            c.sync()
            remarched += int((ain >= 0).sum().item())
        c.partial_color(cam, ain.data_ptr(), parts[r].data_ptr())
        c.sync()
        total += parts[r]
        torch.cuda.synchronize()
    compose.remarched = remarched
    ctxs[0].partial_finalize(cam, total.data_ptr())
    ctxs[0].present()
    return ctxs[0].readback(), ctxs[0].readback_rgba8()


@pytest.mark.parametrize("scheme", ["two-pass", "deferred"])
@pytest.mark.parametrize("world", [2, 8])
@pytest.mark.parametrize("mode", [abi.MODE_M0, abi.MODE_M1])
def test_simulated_ranks_match_single_gpu(oracle, world, mode, scheme):
    import torch

    from vokselis_b200 import rt, sortlast

    W, H, n = 384, 216, 128
    gn = (n, n, n)
    grid = sortlast.brick_grid(world)
    with rt.Context(0, W, H) as full:
        if mode == abi.MODE_M0:
            full.generate_xor(n, 0)
            color, normal = full.download_rgba16f()
        else:
            full.generate_synthetic(2, np.float32, n, seed=5)
        p = rt.default_params(mode)
        if mode == abi.MODE_M1:
            p.dt_scale = 2.0
        p.skip_empty = 1
        full.set_params(p)
        ctxs = [rt.Context(0, W, H) for _ in range(world)]
        try:
            for r, c in enumerate(ctxs):
                lo, hi = sortlast.brick_range(gn, grid, r)
                if mode == abi.MODE_M0:
                    wl = [max(l - 1, 0) for l in lo]
                    wh = [min(h + 1, n) for h in hi]
                    sl = (slice(wl[2], wh[2]), slice(wl[1], wh[1]), slice(wl[0], wh[0]))
                    c.upload_window(gn, lo, hi, color=color[sl], normal=normal[sl])
                else:
                    c.generate_synthetic_window(2, np.float32, gn, lo, hi, seed=5)
                c.set_params(p)
            for zoom, pitch, yaw in [(3.0, -0.5, 1.0), (2.0, 0.6, -2.3), (0.7, 0.1, 0.4)]:  # outside, oblique, eye inside the box
                cam = rt.Camera(zoom, pitch, yaw, (0.03, -0.02, 0.05), W / H).get_proj_view_matrix()
                full.render(cam)
                full.present()
                ref, ref8 = full.readback(), full.readback_rgba8()
                got, got8 = compose(rt, sortlast, torch, ctxs, cam, gn, grid, W, H, scheme)
                if scheme == "deferred":
                    assert compose.remarched < world * W * H  # only some pixels are marched twice
                d = np.abs(got8.astype(np.int32) - ref8.astype(np.int32))
                assert d.max() <= 2, f"max |delta| {d.max()}/255 (world {world}, mode {mode}, cam {(zoom, pitch, yaw)})"
                assert psnr8(got8, ref8) >= 50.0
                g = got.view(np.float16).astype(np.float32)
                r_ = ref.view(np.float16).astype(np.float32)
                assert np.abs(g - r_).max() <= 4e-3
        finally:
            for c in ctxs:
                c.close()

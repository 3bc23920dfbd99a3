"""Multi-GPU path. CPU tier: the host-side tile partition under a world_size-2 gloo group. GPU tier
(needs >= 2 GPUs, skipped otherwise): sort-first frames equal the single-GPU frame bit for bit."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _gloo_worker(rank, world, port, W, H, tile, q):
    import torch.distributed as dist

    sys.path.insert(0, str(ROOT))
    from vokselis_b200 import rt

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = rt.sortfirst_partition(W, H, tile, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine.tolist())
    dist.barrier()
    if rank == 0:
        q.put(gathered)
    dist.destroy_process_group()


@pytest.mark.parametrize("W,H,tile", [(1920, 1080, 120), (1280, 720, 256), (3840, 2160, 100)])
def test_tile_partition_covers_frame_once_world2_gloo(W, H, tile):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + W) % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, W, H, tile, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    tiles = [tuple(t) for part in gathered for t in part]
    cols, rows = -(-W // tile), -(-H // tile)
    assert len(tiles) == len(set(tiles)) == cols * rows
    assert set(tiles) == {(float(x * tile), float(y * tile)) for y in range(rows) for x in range(cols)}
    assert abs(len(gathered[0]) - len(gathered[1])) <= 1  # balanced
    cover = np.zeros((H, W), np.int32)
    for x, y in tiles:
        cover[int(y):int(y) + tile, int(x):int(x) + tile] += 1
    assert (cover == 1).all()


def _shared_frames_worker(rank, world, port, q):
    import torch.distributed as dist

    sys.path.insert(0, str(ROOT))
    from vokselis_b200 import rt

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, H, W = 6, 9, 16
    seg = rt.SharedHostFrames(rank, world, dist, n, H, W, register=False)  # page-locking needs the CUDA driver: GPU tier
    lo, hi = rank * n // world, (rank + 1) * n // world  # every rank delivers a contiguous share of the sweep
    for f in range(lo, hi):
        seg.array[f] = 10 * rank + f
    dist.barrier()
    if rank == 0:
        q.put([int(seg.array[f].min()) * 1000 + int(seg.array[f].max()) for f in range(n)])
    seg.close()
    dist.destroy_process_group()


def test_shared_host_frames_world2_gloo():
    """The host-side gather of a sort-first sweep: one segment, created by rank 0, mapped by every rank; each rank
    writes its share of the frames, rank 0 (the consumer) sees all of them."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 77) % 2000
    procs = [ctx.Process(target=_shared_frames_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    seen = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = [(10 * (0 if f < 3 else 1) + f) for f in range(6)]
    assert seen == [v * 1000 + v for v in expect]


@pytest.mark.gpu
def test_sortfirst_two_gpus_bit_exact():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29611", str(ROOT / "tests" / "mgpu_sortfirst_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


# ---- sort-last host logic (CPU) ------------------------------------------------------------------
def test_brick_partition_covers_grid_once():
    from vokselis_b200 import sortlast

    for world in (1, 2, 4, 8):
        grid = sortlast.brick_grid(world)
        assert grid[0] * grid[1] * grid[2] == world
        for gn in ((256, 256, 256), (4096, 4096, 4096), (200, 96, 72)):
            cover = np.zeros(tuple(-(-n // 8) for n in gn)[::-1], np.int32)
            for r in range(world):
                lo, hi = sortlast.brick_range(gn, grid, r)
                assert all(l % 8 == 0 for l in lo) and all(h % 8 == 0 or h == n for h, n in zip(hi, gn))
                cover[lo[2] // 8:-(-hi[2] // 8), lo[1] // 8:-(-hi[1] // 8), lo[0] // 8:-(-hi[0] // 8)] += 1
            assert (cover == 1).all(), (world, gn)


def test_visibility_order_is_front_to_back_for_every_ray(oracle):
    """For random eyes and rays: the bricks a ray crosses, in crossing order, appear in that order in
    visibility_order()."""
    from vokselis_b200 import sortlast

    rng = np.random.default_rng(5)
    gn = (256, 256, 256)
    for world in (2, 4, 8):
        grid = sortlast.brick_grid(world)
        ranges = [sortlast.brick_range(gn, grid, r) for r in range(world)]
        for _ in range(40):
            eye = rng.uniform(-3, 3, 3)
            order = sortlast.visibility_order(eye, gn, grid)
            pos = {r: i for i, r in enumerate(order)}
            for _ in range(50):
                tgt = rng.uniform(-1, 1, 3)
                d = (tgt - eye) / np.linalg.norm(tgt - eye)
                ts = np.linspace(0, 8, 4000)
                q = ((eye[None, :] + ts[:, None] * d[None, :]) + 1.0) * 128.0
                inside = ((q >= 0) & (q < 256)).all(axis=1)
                seq = []
                for v in q[inside].astype(int):
                    for r, (lo, hi) in enumerate(ranges):
                        if all(lo[a] <= v[a] < hi[a] for a in range(3)):
                            if not seq or seq[-1] != r:
                                seq.append(r)
                            break
                assert len(seq) == len(set(seq)), "a ray re-entered a brick"
                assert [pos[r] for r in seq] == sorted(pos[r] for r in seq), (eye, seq, order)


def test_group_dealing_covers_every_frame_once_and_fits_the_ring():
    """'frames' granularity with batched launches (host logic only): for the world sizes and step counts the bench
    is run with, the chosen group size keeps the ring within the library's slot limit, every frame has exactly
    one owner, a group never straddles the ring's end, and ranks differ by at most one group."""
    from vokselis_b200 import sortfirst

    for world in (2, 4, 8):
        for steps in (360, 100, 50, 16, 8, 3, 1):
            b = sortfirst.choose_batch(steps, world)
            slots = 2 * world * b
            assert 1 <= b <= sortfirst.rt.MAX_BATCH and slots <= sortfirst.MAX_SLOTS and slots % b == 0
            groups = -(-steps // b)
            owners = [sortfirst.frame_owner(g * b, world, b) for g in range(groups)]
            for g in range(groups):
                first = g * b
                assert all(sortfirst.frame_owner(f, world, b) == owners[g] for f in range(first, first + b))
                assert first % slots + b <= slots
            per_rank = [owners.count(r) for r in range(world)]
            assert max(per_rank) - min(per_rank) <= 1

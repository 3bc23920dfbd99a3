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


@pytest.mark.gpu
def test_sortfirst_two_gpus_bit_exact():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29611", str(ROOT / "tests" / "mgpu_sortfirst_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]

"""2-GPU diagnostics (torchrun): peer-store cost of the sort-first kernel vs local stores, P2P copy bandwidth."""
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vokselis_b200 import abi, rt, sortfirst, volumes  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W, H = 1920, 1080
ctx = rt.Context(local, W, H)
ctx.upload_scalar(volumes.xor_u8(256))
p = rt.default_params(abi.MODE_M1)
p.skip_empty, p.tile_size = 1, 120
ctx.set_params(p)
cam = rt.Camera(3.0, -0.5, 1.0, (0, 0, 0), W / H).get_proj_view_matrix()
tiles = rt.sortfirst_partition(W, H, 120, rank, world)
ctx.timing_enable(20)
for _ in range(20):
    ctx.render_tiles(cam, tiles)
local_ms = ctx.timing_read(20)[5:].mean()
print(f"rank {rank}: {len(tiles)} tiles into LOCAL frame: {local_ms:.4f} ms", flush=True)
if rank == 0:
    a = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
    b = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:1")
    for _ in range(3):
        b.copy_(a)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        b.copy_(a)
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    print(f"P2P copy 0->1: {10 * 256 / 1024 / (time.perf_counter() - t0):.1f} GiB/s; can_access_peer={torch.cuda.can_device_access_peer(0, 1)}", flush=True)
dist.barrier()
group = sortfirst.SortFirstGroup(ctx, rank, world, tile=120)
ctx.timing_enable(40)
for i in range(40):
    group.render(cam)
ctx.sync()
dist.barrier()
ms = group.frame_ms(40)
print(f"rank {rank}: sort-first frame (begin->end on this rank) mean {ms[10:].mean():.4f} ms  min {ms.min():.4f} max {ms.max():.4f}", flush=True)
# host-synchronised per frame
tot = 0.0
for i in range(20):
    ctx.sync(); dist.barrier()
    t0 = time.perf_counter()
    group.render(cam)
    ctx.sync()
    tot += time.perf_counter() - t0
print(f"rank {rank}: host-synchronised sort-first frame {1e3 * tot / 20:.4f} ms wall", flush=True)
if rank == 0:
    print("timeouts", ctx.sortfirst_timeouts(), flush=True)
group.close()
ctx.close()
dist.destroy_process_group()

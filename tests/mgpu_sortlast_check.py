"""Run under torchrun with N in {2,4,8} GPUs: sort-last frames (one brick per rank, NCCL all-gather of
transmittances + NCCL sum of partials) vs the single-GPU frame of the same volume on rank 0."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vokselis_b200 import abi, rt, sortlast  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, H, n = 640, 360, 256
    gn = (n, n, n)
    ok = True
    ctx = rt.Context(local, W, H)
    p = rt.default_params(abi.MODE_M1)
    p.dt_scale, p.skip_empty = 2.0, 1
    cams = [rt.Camera(z, pt, y, (0, 0, 0), W / H).get_proj_view_matrix() for z, pt, y in [(3.0, -0.5, 1.0), (2.0, 0.6, -2.3), (0.7, 0.1, 0.4)]]
    refs = []
    if rank == 0:
        ctx.generate_synthetic(2, np.float32, n, seed=5)
        ctx.set_params(p)
        for cam in cams:
            ctx.render(cam)
            ctx.present()
            refs.append(ctx.readback_rgba8())
    group = sortlast.SortLastGroup(ctx, rank, world, gn)
    ctx.generate_synthetic_window(2, np.float32, gn, group.own_lo, group.own_hi, seed=5)
    ctx.set_params(p)
    for i, cam in enumerate(cams):
        group.render(cam)
        ctx.sync()
        if rank == 0:
            ctx.present()
            got = ctx.readback_rgba8()
            d = np.abs(got.astype(np.int32) - refs[i].astype(np.int32))
            good = d.max() <= 2
            ok = ok and good
            print(f"sort-last world {world} cam {i}: max |delta| {d.max()}/255 {'ok' if good else 'MISMATCH'}", flush=True)
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    ctx.close()
    dist.destroy_process_group()
    return int(flag.item() != 0)


if __name__ == "__main__":
    sys.exit(main())

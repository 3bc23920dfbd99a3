"""Transmittance exchange alone (development): W*H floats from every rank to the ranks behind it, p2p direct-send vs NCCL
all-gather, idle GPUs. Run under torchrun. usage: exchange_probe.py [WxH=3840x2160] [iters=20]"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vokselis_b200 import rt  # noqa: E402

W, H = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "3840x2160").split("x"))
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = rt.Context(local, W, H)
n = W * H
dev = torch.device("cuda", local)
T = torch.rand(n, dtype=torch.float32, device=dev)
T_all = torch.empty(world * n, dtype=torch.float32, device=dev)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
handles = [None] * world
dist.all_gather_object(handles, ctx.exchange_create(rank, world))
ctx.exchange_open(handles)
dist.barrier()
behind = list(range(rank + 1, world))  # identity visibility order
res = {}
for name in ("p2p", "nccl", "p2p"):
    ms = []
    for f in range(iters):
        ctx.sync()
        dist.barrier()
        with torch.cuda.stream(stream):
            ctx.mark(0)
            if name == "p2p":
                ff = f + (0 if "p2p" not in res else iters)
                ctx.exchange_push(T.data_ptr(), behind, ff)
                ctx.mark(1)
                ctx.exchange_wait(ff, rank)
                ctx.exchange_done(ff)
            else:
                dist.all_gather_into_tensor(T_all, T)
                ctx.mark(1)
            ctx.mark(2)
        ctx.sync()
        ms.append((ctx.mark_elapsed(0, 1), ctx.mark_elapsed(0, 2)))
    a = np.array(ms[3:])
    res[name] = a.mean(axis=0)
    out = [None] * world
    dist.all_gather_object(out, [float(v) for v in a.mean(axis=0)])
    if rank == 0:
        print(name, "per rank [push-or-collective ms, incl. wait ms]:", [[round(v, 3) for v in o] for o in out], flush=True)
if rank == world - 1:  # the last rank received everybody's image
    tab = torch.empty(0)
    ok = ctx.exchange_timeouts() == 0
    print("timeouts", ctx.exchange_timeouts(), flush=True)
dist.barrier()
ctx.exchange_close()
ctx.close()
dist.destroy_process_group()

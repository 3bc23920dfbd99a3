"""Run under torchrun with N >= 2 GPUs: sort-first frames (tiles rendered on all ranks, peer-stored
into rank 0's frame) must equal the single-GPU frame bit for bit. Exit code 0 = pass."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vokselis_b200 import abi, rt, sortfirst, volumes  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, H = 1280, 720
    ok = True
    for mode in (abi.MODE_M0, abi.MODE_M1):
        ctx = rt.Context(local, W, H)
        if mode == abi.MODE_M0:
            ctx.generate_xor(128, 0)
        else:
            ctx.upload_scalar(volumes.bonsai_standin_u8(128, seed=2, blobs=10))
        p = rt.default_params(mode)
        p.skip_empty = 1
        ctx.set_params(p)
        cams = [rt.Camera(2.6, -0.4, 0.5 + 0.7 * i, (0, 0, 0), W / H).get_proj_view_matrix() for i in range(5)]
        refs = []
        if rank == 0:
            for cam in cams:
                ctx.render(cam)
                refs.append(ctx.readback())
        for gran, slots in (("tiles", 2), ("frames", 2), ("frames", None)):
            group = sortfirst.SortFirstGroup(ctx, rank, world, granularity=gran, tile=96, slots=slots)
            seq = cams * 3  # 15 frames back to back: exercises slot reuse and the consumed flag
            for i, cam in enumerate(seq):
                f = group.submit(cam)
                if rank == 0:
                    group.wait(f)
                    got = ctx.readback()
                    group.consume(f)
                    same = np.array_equal(got, refs[i % len(cams)])
                    ok = ok and same
                    if not same or i == len(seq) - 1:
                        print(f"mode {mode} {gran}/{slots} frame {i}: {'bit-exact' if same else 'MISMATCH'} "
                              f"({(got != refs[i % len(cams)]).sum()} differing halfs)", flush=True)
            if rank == 0:
                to = ctx.sortfirst_timeouts()
                ok = ok and to == 0
                print(f"mode {mode} {gran}/{slots}: device-side wait timeouts = {to}", flush=True)
            group.close()
        # batched groups dealt round-robin, one launch per group, peers shipping a group with one copy-engine transfer:
        # groups of 3 (default ring), and groups of 5 in an 80-slot ring (what 8 ranks use) long enough to wrap it
        for batch, slots, total in ((3, None, 20), (5, 80, 103)):
            group = sortfirst.SortFirstGroup(ctx, rank, world, granularity="frames", batch=batch, slots=slots)
            seq = (cams * (total // len(cams) + 1))[:total]
            done, good = 0, True
            for g in range(0, len(seq), batch):
                chunk = seq[g:g + batch]
                f = group.submit_batch(chunk, flush_l2=(g % 2 == 0))
                if rank == 0:
                    for k in range(batch):
                        group.wait(f + k)
                        got = ctx.readback()
                        group.consume(f + k)
                        if k < len(chunk):
                            same = np.array_equal(got, refs[(g + k) % len(cams)])
                            good = good and same
                            done += 1
                            if not same:
                                print(f"mode {mode} frames/batch{batch} frame {g + k}: MISMATCH", flush=True)
            if rank == 0:
                to = ctx.sortfirst_timeouts()
                good = good and to == 0
                ok = ok and good
                print(f"mode {mode} frames/batch{batch}/slots{group.slots}: {done} frames {'bit-exact' if good else 'FAILED'}, device-side wait timeouts = {to}", flush=True)
            group.close()
        ctx.close()
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    dist.destroy_process_group()
    return int(flag.item() != 0)


if __name__ == "__main__":
    sys.exit(main())

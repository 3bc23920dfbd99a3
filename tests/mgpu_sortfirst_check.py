"""Run under torchrun with N >= 2 GPUs: sort-first frames (tiles rendered on all ranks and shipped into rank 0's
frame, whole frames dealt round-robin, batched groups) must equal the single-GPU frame bit for bit.
Exit code 0 = pass. The body lives in vokselis_b200/workloads.py (bench.py runs the same check before timing)."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vokselis_b200 import workloads  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    r = workloads.check_sortfirst(rank, world, local, dist, log=(lambda m: print(m, flush=True)) if rank == 0 else None)
    if rank == 0:
        print(r, flush=True)
    dist.destroy_process_group()
    return 0 if r["ok"] else 1


if __name__ == "__main__":
    sys.exit(main())

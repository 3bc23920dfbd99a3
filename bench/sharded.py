"""BASELINE.json's sharded configurations (configs[2..4]) and the multi-GPU correctness checks, stand-alone
(development; bench.py runs the same functions from vokselis_b200/workloads.py). Run alone (1 GPU) or under torchrun.
usage: sharded.py [--what checks,3,4,5] [--frames F] [--edge5 4096] [--tile 120] [--layout 3]"""
import argparse
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from vokselis_b200 import abi, workloads  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="checks,3,4,5")
    ap.add_argument("--frames", type=int, default=24)
    ap.add_argument("--edge5", type=int, default=4096)
    ap.add_argument("--edge34", type=int, default=0)
    ap.add_argument("--tile", type=int, default=120)
    ap.add_argument("--tbatch", type=int, default=1, help="frames per launch of a rank's tile share (configs 3/4)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"], help="sort-last transmittance exchange")
    ap.add_argument("--brick", type=int, default=0, help="occupancy brick edge (configs 3/4); 0 = automatic")
    ap.add_argument("--layout", type=int, default=-1, help="-1 = the layout workloads.py chose per config")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dist = None
    import torch

    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    peak = 6650.0
    if (ROOT / "MEASURED_PEAKS.json").exists():
        peak = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
    what = args.what.split(",")
    say = (lambda m: print(m, file=sys.stderr, flush=True)) if rank == 0 else None
    if "checks" in what and world > 1:
        for fn in (workloads.check_sortfirst, workloads.check_sortlast):
            r = fn(rank, world, local, dist, log=say)
            if rank == 0:
                print(json.dumps({"check": fn.__name__, **r}), flush=True)
    for cid in (3, 4):
        if str(cid) in what:
            r = workloads.run_sortfirst_tiles(cid, rank, world, local, dist, frames=args.frames, tile=args.tile, layout=(None if args.layout < 0 else args.layout),
                                              edge=args.edge34 or None, hbm_peak_gbs=peak, batch=args.tbatch, brick=args.brick)
            if rank == 0:
                print(json.dumps(r), flush=True)
    if "5" in what and world > 1:
        r = workloads.run_sortlast(rank, world, local, dist, edge=args.edge5, frames=max(args.frames // 4, 3), hbm_peak_gbs=peak, exchange=args.exchange)
        if rank == 0:
            print(json.dumps(r), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

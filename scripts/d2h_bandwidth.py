#!/usr/bin/env python
"""What the host can absorb: device -> pinned-host copies from N GPUs at once (the floor of any host-output API).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/d2h_bandwidth.py [--gb 1.6]

Every rank copies the same amount from its GPU into its own page-locked buffer, all ranks together (barrier before,
CUDA events, max over ranks); rank 0 prints one JSON line: per-rank and aggregate GB/s.  DESIGN.md quotes the result
next to ``bench.py``'s ``e2e``.
"""
import argparse
import json
import os

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gb", type=float, default=1.6, help="GB copied per rank and repetition")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = int(args.gb * 1e9) // 4
    dev = torch.empty(n, dtype=torch.float32, device="cuda").fill_(1.0)
    pin = torch.empty(n, dtype=torch.float32, pin_memory=True)
    for _ in range(2):
        pin.copy_(dev, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        pin.copy_(dev, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if int(os.environ.get("RANK", "0")) == 0:
        per = n * 4 / ms / 1e6
        print(json.dumps({"n_gpus": world, "gb_per_rank": n * 4 / 1e9, "ms": ms, "gbs_per_rank": per,
                          "gbs_aggregate": per * world, "cores": len(os.sched_getaffinity(0))}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

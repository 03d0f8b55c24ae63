#!/usr/bin/env python
"""Copy-engine bandwidth of pushes into the peers' symmetric images (what satpy_b200.parallel.PeerImage does), for
several piece sizes and stream counts, all ranks starting each measurement together:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/p2p_bandwidth.py

Measured on 2 and 4 x B200: 765-775 GB/s egress per rank for every variant.
"""
import os, sys, time, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
H, W = 40075, 80150
rows_per = H // world
try:
    t0 = time.time()
    buf = symm_mem.empty((H, W), dtype=torch.float32, device=torch.device("cuda", local))
    hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
    peers = [hdl.get_buffer(r, (H, W), torch.float32) for r in range(world)]
    torch.cuda.synchronize()
    if rank == 0: print(f"rendezvous ok in {time.time()-t0:.2f}s", flush=True)
except Exception as e:
    print(f"rank {rank}: symmetric memory failed: {type(e).__name__}: {e}", flush=True)
    dist.destroy_process_group(); sys.exit(0)
buf.fill_(float(rank + 1))
hdl.barrier()
streams = [torch.cuda.Stream() for _ in range(world)]
NST = 1
def push(r0, n):
    for k in range(1, world):
        p = (rank + k) % world
        if NST == 1:
            peers[p][r0:r0 + n].copy_(buf[r0:r0 + n], non_blocking=True)
        else:
            st = streams[(k - 1) % NST]
            st.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(st):
                peers[p][r0:r0 + n].copy_(buf[r0:r0 + n], non_blocking=True)
def join():
    for st in streams: torch.cuda.current_stream().wait_stream(st)
for NST, pieces in ((1, 1), (1, 4), (2, 4), (world - 1, 1), (world - 1, 4)):
    n = rows_per // pieces
    for it in range(3):
        hdl.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for j in range(pieces):
            push(rank * rows_per + j * n, n)
        join()
        e1.record()
        torch.cuda.synchronize()
        hdl.barrier()
        ms = e0.elapsed_time(e1)
    sent = (world - 1) * pieces * n * W * 4
    if rank == 0: print(f"streams {NST} pieces {pieces}: {ms:.2f} ms for {sent/1e9:.2f} GB sent per rank = {sent/ms/1e6:.0f} GB/s egress", flush=True)
# correctness: every rank's rows hold the owner's value
hdl.barrier(); torch.cuda.synchronize()
ok = all(float(buf[r * rows_per + 5, 7]) == r + 1 for r in range(world))
print(f"rank {rank}: assembled image correct: {ok}", flush=True)
# barrier cost
torch.cuda.synchronize(); t0 = time.time()
for _ in range(20): hdl.barrier()
torch.cuda.synchronize()
if rank == 0: print(f"hdl.barrier: {(time.time()-t0)/20*1e3:.3f} ms", flush=True)
dist.destroy_process_group()

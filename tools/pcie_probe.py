#!/usr/bin/env python
"""Host<->device copy bandwidth with every rank copying at the same time (run under torchrun, or alone for N = 1).
Explains the e2e figure of bench.py at N > 1: each rank moves 1/N of the right-hand side and of the solution, so e2e can
only scale if the AGGREGATE pinned-copy bandwidth of the box scales with the number of GPUs.
   python tools/pcie_probe.py            |  python -m torch.distributed.run --nproc-per-node N tools/pcie_probe.py"""
import json
import os
import time

import torch

world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl")
nbytes = 512 << 20
h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
h2 = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
d2 = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=6):
    fn(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    return reps * nbytes / dt / 1e9


def both():
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)


out = {"ranks": world,
       "h2d_GBs_per_rank": timed(lambda: d.copy_(h, non_blocking=True)),
       "d2h_GBs_per_rank": timed(lambda: h2.copy_(d2, non_blocking=True)),
       "h2d_and_d2h_GBs_per_rank_each_direction": timed(both)}
out["h2d_GBs_aggregate"] = out["h2d_GBs_per_rank"] * world
out["d2h_GBs_aggregate"] = out["d2h_GBs_per_rank"] * world
out["duplex_GBs_aggregate_each_direction"] = out["h2d_and_d2h_GBs_per_rank_each_direction"] * world
if int(os.environ.get("RANK", "0")) == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()

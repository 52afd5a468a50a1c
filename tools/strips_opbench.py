#!/usr/bin/env python
"""Per-operation timings on row strips (run under torchrun, one process per GPU): the fused legs with their ghost-row
exchange, the exchange alone, the all-reduce.  Prints the max over ranks of the CUDA-event time per launch."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
strips = importlib.import_module("multigrid-petsc_b200.strips")
mgb = importlib.import_module("multigrid-petsc_b200")
import torch
import torch.distributed as dist

npts, levels = 8193, 13
opts = (f"-npts {npts} -mesh 0 -iter 100 -grids {levels} -levels {levels} -cycle 0 -map 2 -v 3,3 -moreNorm 0 "
        "-pc_type jacobi -ksp_richardson_scale 0.8")
s = strips.StripSession(opts)
e = s.engine
e.solve_vcycle(mgb.jacobi(0.8), 3, 3, max_iter=3, rtol=0.0)
for l in range(4):
    for op in ("halo_exchange", "nrm2_allreduce", "jacobi", "fused_down", "fused_up"):
        dist.barrier(); torch.cuda.synchronize()
        ms = strips._max_over_ranks(e.time_op(op, l, 20))
        if s.rank == 0:
            print(f"level {l} rows/rank {e.local_rows(l)[1] - e.local_rows(l)[0]:5d} {op:>16} {ms * 1e3:9.2f} us", flush=True)
s.close()
dist.barrier()
dist.destroy_process_group()

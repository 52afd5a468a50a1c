#!/usr/bin/env python
"""Small solves that touch every kernel family (one-sweep kernels, fused legs, bottom cluster kernel, red-black SOR,
CG + coarse LU, CSR assembly and SpMV, row strips emulated on one GPU) -- sized for compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mgb = importlib.import_module("multigrid-petsc_b200")

JAC = "-pc_type jacobi -ksp_richardson_scale 0.8"
MGJ = "-mg_levels_ksp_type richardson -mg_levels_pc_type jacobi -mg_levels_ksp_richardson_scale 0.8 -mg_levels_ksp_max_it 3"


def base(npts, levels, cycle=0, mp=2, it=6):
    return f"-npts {npts} -mesh 0 -iter {it} -grids {levels} -levels {levels} -cycle {cycle} -map {mp} -v 3,3 -moreNorm 0"


runs = [base(129, 7) + " " + JAC,                                   # fused legs + bottom kernel + graph replay
        base(129, 7) + " " + JAC + " -mgb_fuse 0",                  # one kernel per sweep
        base(513, 6) + " " + JAC,                                   # several fused tiles and row chunks
        base(129, 5, mp=3) + " -pc_type sor -mgb_csr 0",            # red-black SOR
        base(129, 7, cycle=8, it=8) + " -ksp_type cg -ksp_rtol 1e-10 " + MGJ,   # CG + PCMG + coarse LU
        base(129, 5) + " " + JAC + " -mgb_ranks 3 -mgb_emulate 1 -mgb_agglomerate 31",           # strips: ghost pushes, gather, bcast
        base(257, 6, cycle=8, it=6) + " -ksp_type cg " + MGJ + " -mgb_ranks 2 -mgb_emulate 1 -mgb_agglomerate 63"]
for opts in runs:
    r = mgb.run_poisson(opts)
    print(f"{r['num_iter']:3d} iterations, final {r['rnorm'][-1]:.3e} : {opts}", flush=True)
e = mgb.Engine(4, 127)
e.set_poisson_uniform()
e.assemble_csr()
x = np.random.default_rng(0).uniform(-1, 1, 127 * 127)
y = e.csr_spmv(mgb.MAT_A, 0, x)
e.close()
print("csr ok", float(np.abs(y).max()))

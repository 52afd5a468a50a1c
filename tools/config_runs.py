#!/usr/bin/env python
"""The BASELINE.json configurations that fit one GPU, through the host C layer (poisson.in vocabulary): iteration
counts, final residuals, discretisation error and the CUDA-event time of the cycle loop.  Writes one JSON line per run.
   python tools/config_runs.py [--json out.jsonl]"""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mgb = importlib.import_module("multigrid-petsc_b200")

JAC = "-pc_type jacobi -ksp_richardson_scale 0.8"
RB = "-pc_type sor"
MGJ = "-mg_levels_ksp_type richardson -mg_levels_pc_type jacobi -mg_levels_ksp_richardson_scale 0.8 -mg_levels_ksp_max_it 3"


def base(npts, levels, cycle=0, mp=2, it=100000):
    return f"-npts {npts} -mesh 0 -iter {it} -grids {levels} -levels {levels} -cycle {cycle} -map {mp} -v 3,3 -moreNorm 0 -mgb_csr 0"


RUNS = [
    ("config0: 129^2, 4-level V(3,3), Jacobi 0.8", base(129, 4) + " " + JAC),
    ("config0: 129^2, 4-level V(3,3), red-black SOR", base(129, 4, mp=3) + " " + RB),
    ("config1: 1025^2, 7-level V(3,3), red-black SOR (coarsest 15x15, 3 sweeps)", base(1025, 7, mp=3) + " " + RB),
    ("config1: 1025^2, 7-level V(3,3), Jacobi 0.8", base(1025, 7) + " " + JAC),
    ("config1': 1025^2, 10-level V(3,3), red-black SOR (coarsest 1x1)", base(1025, 10, mp=3) + " " + RB),
    ("config1': 1025^2, 10-level V(3,3), Jacobi 0.8", base(1025, 10) + " " + JAC),
    ("config1b: 4097^2, 12-level V(3,3), red-black SOR (levels >= 2047 rows: fused legs)", base(4097, 12, mp=3) + " " + RB),
    ("config1b: 4097^2, 12-level V(3,3), red-black SOR, one-sweep kernels only (-mgb_fuse 0)", base(4097, 12, mp=3) + " " + RB + " -mgb_fuse 0"),
    ("config1c: 8193^2, 13-level V(3,3), red-black SOR (fused legs on levels 0-2)", base(8193, 13, mp=3) + " " + RB),
    ("config1c: 8193^2, 13-level V(3,3), red-black SOR, one-sweep kernels only (-mgb_fuse 0)", base(8193, 13, mp=3) + " " + RB + " -mgb_fuse 0"),
    ("config2: 4097^2, MG(12 levels, Jacobi 0.8 x3, LU coarse)-preconditioned CG to 1e-10", base(4097, 12, cycle=8, it=200) + " -ksp_type cg -ksp_rtol 1e-10 " + MGJ),
    # plain V-cycle iteration cannot reach 1e-10 on these grids in fp64: the true residual b - A u stagnates at
    # ~1.6e-10 (4097^2) / ~6.3e-10 (8193^2) of ||b|| (eps * cond); 1e-9 is reachable, 1e-10 needs the CG wrapper,
    # whose recursively updated residual keeps decreasing -- the same holds for the reference's PETSc solve
    ("config2': 4097^2, 12-level V(3,3) Jacobi 0.8 to 1e-9 (cycle 0, -rtol)", base(4097, 12, it=60) + " " + JAC + " -rtol 1e-9"),
    ("config3 @1 GPU: 8193^2, 13-level V(3,3), Jacobi 0.8 to 1e-7", base(8193, 13) + " " + JAC),
    ("config3 @1 GPU: 8193^2, 13-level V(3,3), Jacobi 0.8 to 1e-9", base(8193, 13, it=60) + " " + JAC + " -rtol 1e-9"),
    ("config3' @1 GPU: 8193^2, MG-preconditioned CG to 1e-10", base(8193, 13, cycle=8, it=200) + " -ksp_type cg -ksp_rtol 1e-10 " + MGJ),
]

ap = argparse.ArgumentParser()
ap.add_argument("--json", default=None)
a = ap.parse_args()
out = open(a.json, "w") if a.json else None
for name, opts in RUNS:
    mgb.run_poisson(opts, want_u=False)                      # warm-up (graph instantiation, first-touch)
    r = mgb.run_poisson(opts, want_u=False)
    row = {"run": name, "options": opts, "iterations": r["num_iter"], "final_relative_residual": float(r["rnorm"][-1]),
           "max_error": float(r["error"][0]), "solve_ms": 1e3 * r["solve_seconds"],
           "iterations_per_s": r["num_iter"] / r["solve_seconds"] if r["solve_seconds"] > 0 else None}
    print(json.dumps(row), flush=True)
    if out:
        out.write(json.dumps(row) + "\n")

#!/usr/bin/env python
"""Launch a few instances of selected operations on the fine level (for ncu captures):
   python tools/profile_ops.py [--npts 8193] [--levels 13] [--ops fused_down,fused_up] [--reps 3]"""
import argparse
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mgb = importlib.import_module("multigrid-petsc_b200")

ap = argparse.ArgumentParser()
ap.add_argument("--npts", type=int, default=8193)
ap.add_argument("--levels", type=int, default=13)
ap.add_argument("--level", type=int, default=0)
ap.add_argument("--ops", default="fused_down,fused_up")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
n = a.npts - 2
e = mgb.Engine(a.levels, n)
e.set_poisson_uniform()
x = np.linspace(0, 1, a.npts)[1:-1]
e.set_rhs_separable(-2 * np.pi ** 2 * np.sin(np.pi * x), np.sin(np.pi * x))
rng = np.random.default_rng(0)
for l in range(min(a.level + 2, a.levels)):                 # non-trivial data on the level and the one below it
    ni, nj = e.dims(l)
    e.set_vec(mgb.VEC_U, l, rng.uniform(-1, 1, (ni, nj)))
    if l > 0:
        e.set_vec(mgb.VEC_B, l, rng.uniform(-1, 1, (ni, nj)))
for op in a.ops.split(","):
    ms = e.time_op(op, a.level, a.reps)
    print(f"{op}: {ms * 1e3:.1f} us")
e.close()

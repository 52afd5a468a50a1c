#!/usr/bin/env python
"""One GPU, one strip-sized RECTANGULAR grid (ni x nj unknowns, no peers): level-0 kernel times to set beside the per-strip
times of tools/strips_opbench.py -- separates what a strip costs because it is a strip (peer mapping, in-kernel traffic)
from what it costs because it is half / a quarter / an eighth of the grid.  usage: rect_opbench.py NI NJ"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mgb = importlib.import_module("multigrid-petsc_b200")
ni, nj = int(sys.argv[1]), int(sys.argv[2])
levels = 1
while ((nj + 1) >> levels) - 1 >= 1 and ((ni + 1) >> levels) - 1 >= 1:
    levels += 1
e = mgb.Engine(levels, ni, nj)
e.set_poisson_uniform()
y = np.linspace(0.0, 1.0, ni + 2)[1:-1]
x = np.linspace(0.0, 1.0, nj + 2)[1:-1]
e.set_rhs_separable(-2 * np.pi ** 2 * np.sin(np.pi * x), np.sin(np.pi * y))
e.solve_vcycle(mgb.jacobi(0.8), 3, 3, max_iter=2, rtol=0.0)
for l in range(min(4, levels - 1)):
    for op in ("jacobi", "fused_down", "fused_up"):
        ms = e.time_op(op, l, 20)
        print(f"rect {ni} x {nj} level {l} rows {e.dims(l)[0]:5d} {op:>12} {ms * 1e3:9.2f} us", flush=True)
e.close()

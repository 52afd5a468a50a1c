#!/usr/bin/env python
"""Per-level bandwidth sweep of the engine's kernels (BASELINE.json configs[4] second half): every operation on
every level of an N x N hierarchy, timed with CUDA events on the engine's stream (mgb_time_op), reported as
algorithmic GB/s (bytes per unknown from SURVEY.md 8d) and as a fraction of the measured HBM peak.
Usage: python tools/opbench.py [--npts 8193] [--levels 13] [--reps 20] [--json out.json]"""
import argparse
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mgb = importlib.import_module("multigrid-petsc_b200")


def hbm_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--npts", type=int, default=8193)
    ap.add_argument("--levels", type=int, default=13)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--csr", type=int, default=0)
    ap.add_argument("--json", default=None)
    ap.add_argument("--ops", default="", help="comma-separated subset of operations")
    ap.add_argument("--maxlevel", type=int, default=99)
    a = ap.parse_args()
    n = a.npts - 2
    e = mgb.Engine(a.levels, n)
    e.set_poisson_uniform()
    if a.csr:
        import time
        e.assemble_csr()                                              # warm-up (allocations)
        t0 = time.perf_counter()
        e.assemble_csr()
        dt = time.perf_counter() - t0
        nbytes = 0
        for l in range(a.levels):
            for which in (mgb.MAT_A, mgb.MAT_RES, mgb.MAT_PRO):
                if which != mgb.MAT_A and l == a.levels - 1:
                    continue
                m, nn, nnz = mgb.C.c_int(), mgb.C.c_int(), mgb.C.c_longlong()
                e._ck(e.L.mgb_csr_dims(e.h, which, l, mgb.C.byref(m), mgb.C.byref(nn), mgb.C.byref(nnz)))
                nbytes += 4 * (m.value + 1) + 12 * nnz.value
        print(f"# CSR assembly of A, res, pro on all levels ({nbytes / 1e9:.2f} GB, arrays already allocated): {dt * 1e3:.2f} ms "
              f"-> {nbytes / dt / 1e9:.0f} GB/s written")
    x = np.linspace(0, 1, a.npts)[1:-1]
    e.set_rhs_separable(-2 * np.pi ** 2 * np.sin(np.pi * x), np.sin(np.pi * x))
    e.solve_vcycle(mgb.jacobi(0.8), 3, 3, max_iter=2, rtol=0.0)       # fill every level with non-trivial data
    peak, how = hbm_peak()
    rows = []
    print(f"# {a.npts}^2, {a.levels} levels, reps {a.reps}; HBM peak {peak} GB/s ({how})")
    print(f"{'level':>5} {'n':>6} {'op':>18} {'us':>10} {'GB/s':>9} {'frac':>6}")
    for l in range(min(a.levels, a.maxlevel + 1)):
        ni, nj = e.dims(l)
        for op, (code, bpu) in mgb.OPS.items():
            if op == "csr_spmv" and not a.csr:
                continue
            if op in ("residual_restrict", "prolong_correct", "fused_down", "fused_up", "fused_down_zero") and l == a.levels - 1:
                continue
            if a.ops and op not in a.ops.split(","):
                continue
            if op == "bottom_cycle" and (l < 1 or ni > 1023):
                continue
            ms = e.time_op(op, l, a.reps)
            gbs = bpu * ni * nj / (ms * 1e-3) / 1e9
            own = mgb.FUSED_OWN_BYTES.get(op)
            own_gbs = own * ni * nj / (ms * 1e-3) / 1e9 if own else None
            rows.append({"level": l, "n": ni, "op": op, "us": ms * 1e3, "gbs": gbs, "frac": gbs / peak, "own_traffic_gbs": own_gbs})
            print(f"{l:5d} {ni:6d} {op:>18} {ms*1e3:10.2f} {gbs:9.1f} {gbs/peak:6.3f}" + (f"   (own traffic {own_gbs:7.1f} GB/s)" if own else ""))
    if a.json:
        json.dump({"npts": a.npts, "levels": a.levels, "peak_gbs": peak, "peak_kind": how, "rows": rows}, open(a.json, "w"), indent=1)
    e.close()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Regenerate tests/golden/golden.json + golden_u.npz.

Run in the BUILD container only (needs /root/reference):   python tests/golden/make_golden.py

Source of the vectors: the reference's own unmodified sources (src/poisson.c, solver.c, matbuild.c,
mesh.c, problem.c, array.c) compiled in place over oracle/minipetsc into oracle/_ref/poisson_ref and
run here, one process per case, reading the files it writes (uData.dat, rData.dat, eData.dat:
src/solver.c:1331-1354).  PETSc itself is absent from this image, so these vectors pin the reference's
DRIVER (index maps, assembly, cycle sequencing, post-processing) -- not PETSc's kernels, which
minipetsc restates (parity with real PETSc remains unpinned; see oracle/minipetsc/petscksp.h).
Cases in ORACLE_ONLY_CASES use the red-black numbering extension and come from the oracle.
"""
import hashlib
import json
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from cases import CASES, ORACLE_ONLY_CASES  # noqa: E402
from oracle import Oracle, build_oracle, ref_binary_path  # noqa: E402


def run_reference(options):
    """Run oracle/_ref/poisson_ref with the given options; return (numIter, rnorm, err, u)."""
    with tempfile.TemporaryDirectory() as d:
        # the reference reads ./poisson.in first, then argv (src/poisson.c:29); give it an empty file
        open(os.path.join(d, "poisson.in"), "w").close()
        out = subprocess.run([ref_binary_path()] + options.split(), cwd=d, check=True,
                             capture_output=True, text=True).stdout
        it = int(re.search(r"Number of iterations:\s+(\d+)", out).group(1))
        rnorm = np.array([float(t) for t in open(os.path.join(d, "rData.dat")).read().split()])
        err = np.array([float(t) for t in open(os.path.join(d, "eData.dat")).read().split()])
        u = np.loadtxt(os.path.join(d, "uData.dat"), ndmin=2)
    return it, rnorm, err, u


def run_oracle(options):
    o = Oracle(options)
    it, rn = o.solve()
    u, err = o.postprocess()
    o.close()
    return it, rn, err, u


def entry(options, src, it, rnorm, err, u):
    return {
        "options": options,
        "source": src,
        "num_iter": it,
        "rnorm_hex": [float(x).hex() for x in rnorm[: it + 1]],
        "error_hex": [float(x).hex() for x in err],
        "u_shape": list(u.shape),
        "u_sha256": hashlib.sha256(np.ascontiguousarray(u, dtype="<f8").tobytes()).hexdigest(),
    }


def main():
    build_oracle(with_ref=True)
    if not os.path.exists(ref_binary_path()):
        sys.exit("oracle/_ref/poisson_ref is missing: run where /root/reference exists")
    gold, us = {}, {}
    for name, (opts, full) in CASES.items():
        it, rn, err, u = run_reference(opts)
        gold[name] = entry(opts, "reference-driver-over-minipetsc (oracle/_ref/poisson_ref)", it, rn, err, u)
        if full:
            us[name] = u
        print(f"{name:28s} iters={it:4d} r1={rn[1] if it else float('nan'):.6e} max_err={err[0]:.6e}")
    for name, (opts, full) in ORACLE_ONLY_CASES.items():
        it, rn, err, u = run_oracle(opts)
        gold[name] = entry(opts, "oracle (red-black numbering extension; no reference counterpart)", it, rn, err, u)
        if full:
            us[name] = u
        print(f"{name:28s} iters={it:4d} r1={rn[1]:.6e} max_err={err[0]:.6e}")
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(gold, f, indent=1)
    np.savez_compressed(os.path.join(HERE, "golden_u.npz"), **us)


if __name__ == "__main__":
    main()

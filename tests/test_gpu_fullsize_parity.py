"""Parity at BASELINE.json's own sizes (pytest -m gpu; slow: the reference runs on the GPU box's host cores).

The checker here is the reference itself: oracle/_ref/poisson_ref = the reference's unmodified src/*.c compiled over the
in-repo mini-PETSc (recipe: oracle/Makefile `ref`), which travels to the GPU box as a built binary.  The product side goes
through the host C layer (pb200_run = the reference's main() sequence) exactly like tests/test_gpu_parity.py.

  config 2   4097^2, 12 levels, cycle 8: CG preconditioned by one V(3,3) Jacobi-0.8 cycle, to 1e-10
             -> EQUAL iteration count, residual history within 1e-10 relative   (ref: src/solver.c:1919-1976)
  4097^2     cycle 0, V(3,3) Jacobi 0.8 to 1e-7: equal iteration count, full history within 1e-10, and the SOLUTION bit for
             bit: both programs write uData.dat with "%.16e" (17 significant digits round-trip a double), so equal files
             mean equal bits                                                     (ref: src/solver.c:1530-1558, 1337-1346)
"""
import hashlib
import importlib
import os
import re
import subprocess

import numpy as np
import pytest

from oracle import ref_binary_path

pytestmark = pytest.mark.gpu

mgb = importlib.import_module("multigrid-petsc_b200")
RTOL = 1e-10
RNORM_ATOL = 2.0 ** -52      # see tests/test_gpu_parity.py: entries are normalised by rnorm[0]
JAC = "-pc_type jacobi -ksp_richardson_scale 0.8"
MGJ = ("-mg_levels_ksp_type richardson -mg_levels_pc_type jacobi "
       "-mg_levels_ksp_richardson_scale 0.8 -mg_levels_ksp_max_it 3")
MGC1 = "-mg_coarse_ksp_type richardson -mg_coarse_pc_type jacobi -mg_coarse_ksp_max_it 1"


def base(npts, levels, cycle=0, it=1000):
    return (f"-npts {npts} -mesh 0 -iter {it} -grids {levels} -levels {levels} "
            f"-cycle {cycle} -map 2 -v 3,3 -moreNorm 0")


def options_file(opts):
    lines = []
    for t in opts.split():
        if t.startswith("-") and not re.fullmatch(r"-[0-9.].*", t):
            lines.append(t)
        else:
            lines[-1] += " " + t
    return "\n".join(lines) + "\n"


def sha_file(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 24), b""):
            h.update(chunk)
    return h.hexdigest()


def run_reference(opts, workdir, keep_u):
    exe = ref_binary_path()
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/poisson_ref is not built (needs /root/reference in the build container)")
    os.makedirs(workdir, exist_ok=True)
    with open(os.path.join(workdir, "poisson.in"), "w") as f:
        f.write(options_file(opts))
    for name in ("XgridData.dat", "YgridData.dat") + (() if keep_u else ("uData.dat",)):
        os.symlink(os.devnull, os.path.join(workdir, name))
    env = dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count() or 1))
    out = subprocess.run([exe], cwd=workdir, env=env, capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0, out.stderr[-800:] + out.stdout[-800:]
    it = int(re.search(r"Number of iterations:\s+(\d+)", out.stdout).group(1))
    rn = np.array([float(t) for t in open(os.path.join(workdir, "rData.dat")).read().split()])
    err = np.array([float(t) for t in open(os.path.join(workdir, "eData.dat")).read().split()])
    return it, rn, err


@pytest.mark.parametrize("coarse", ["", MGC1])
def test_config2_cg_4097_equal_iterations_and_history(coarse, tmp_path):
    """BASELINE configs[2]: MG-preconditioned CG to 1e-10 at 4097^2 (coarsest grid 1 x 1; PCMG's default LU or one Jacobi step)."""
    opts = base(4097, 12, cycle=8, it=100) + " -ksp_type cg -ksp_rtol 1e-10 " + MGJ + " " + coarse
    it_r, rn_r, err_r = run_reference(opts, str(tmp_path / "ref"), keep_u=False)
    r = mgb.run_poisson(opts + " -mgb_csr 0", want_u=False)
    assert r["num_iter"] == it_r                                          # equal iteration count to the same tolerance
    assert len(r["rnorm"]) == it_r + 1
    assert np.allclose(r["rnorm"], rn_r[: it_r + 1], rtol=RTOL, atol=RNORM_ATOL)
    assert r["rnorm"][-1] <= 1e-10
    # the CG solution is reproduced to 1e-10 relative (the dot products are summed in another order), so the error triple
    # {max, sum, sqrt(sum^2)} over n^2 points agrees to 1e-10 * {1, n^2, n} absolute
    n = 4095
    assert np.all(np.abs(r["error"] - err_r) <= RTOL * np.array([1.0, float(n) * n, float(n)]))


def test_cycle0_4097_history_and_solution_bits(tmp_path):
    opts = base(4097, 12) + " " + JAC
    it_r, rn_r, err_r = run_reference(opts, str(tmp_path / "ref"), keep_u=True)
    out = tmp_path / "b200"
    out.mkdir()
    r = mgb.run_poisson(opts + " -mgb_csr 0", out_dir=str(out), want_u=False)
    assert r["num_iter"] == it_r
    assert np.allclose(r["rnorm"], rn_r[: it_r + 1], rtol=RTOL, atol=RNORM_ATOL)
    # Jacobi smoothing has no reduction inside the iteration: the solution is reproduced bit for bit
    assert sha_file(out / "uData.dat") == sha_file(tmp_path / "ref" / "uData.dat")
    assert np.array_equal(r["error"], err_r)                              # same u, same summation order, same libm
    assert open(out / "eData.dat").read() == open(tmp_path / "ref" / "eData.dat").read()

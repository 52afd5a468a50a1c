"""SURVEY.md section 8(f) rank 4, proven by building and running it: the reference's UNMODIFIED src/solver.c on the GPU.

lib/poisson_petsc_b200 = all six reference source files (poisson.c mesh.c problem.c matbuild.c array.c AND solver.c)
compiled in place against host/petsc_b200/petscksp.h -- PETSc's surface as the reference uses it, served by the general
sparse objects of lib/libmgb200.so (include/mgb200_sparse.h: CSR matrices and vectors in HBM, every MatMult / VecAXPY /
MatSOR / ILU(0) / dot product a CUDA kernel).  Recipe: host/Makefile `refsolver` (ref: src/solver.c:1414-2630).  With it
every cycle of the reference (V, I, E, D1, D2, D1PS, PCMG, Additive, Additive2) and several grids per level run on the B200.

The checker is the reference itself over the CPU mini-PETSc (oracle/_ref/poisson_ref, test infrastructure), which travels
to the GPU box: same poisson.in, same files compared.  The GPU layer sums its dot products in the same fixed blocks as the
checker, so the files are expected to be IDENTICAL; the asserted bar is the north star's (equal iteration count, residual
history and solution within 1e-10 relative).
"""
import importlib
import os
import re
import subprocess

import numpy as np
import pytest

from oracle import ref_binary_path

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "multigrid-petsc_b200", "host")
BIN = os.path.join(ROOT, "multigrid-petsc_b200", "lib", "poisson_petsc_b200")
REF = "/root/reference"
RTOL = 1e-10

needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="reference sources not present (GPU box)")

JAC = ["-pc_type jacobi", "-ksp_richardson_scale 0.8"]
MGJ = ["-mg_levels_ksp_type richardson", "-mg_levels_pc_type jacobi", "-mg_levels_ksp_richardson_scale 0.8", "-mg_levels_ksp_max_it 3"]


def base(npts, grids, levels, cycle, it=60, v="3,3"):
    return [f"-npts {npts}", "-mesh 0", f"-iter {it}", f"-grids {grids}", f"-levels {levels}", f"-cycle {cycle}", "-map 2", f"-v {v}", "-moreNorm 0"]


CASES = {
    # the hot path through PETSc's own call sequence (src/solver.c:1414-1575) on general CSR kernels
    "vcycle_129_l4_jacobi": base(129, 4, 4, 0, it=1000) + JAC,
    "vcycle_129_l7_jacobi": base(129, 7, 7, 0, it=1000) + JAC,
    "vcycle_65_l4_sor": base(65, 4, 4, 0, it=1000) + ["-pc_type sor"],                        # lexicographic MatSOR, level sets
    "vcycle_65_l4_sor_w12_forward": base(65, 4, 4, 0, it=1000) + ["-pc_type sor", "-pc_sor_omega 1.2", "-pc_sor_forward"],
    "vcycle_33_l3_ilu_default": base(33, 3, 3, 0, it=1000),                                   # no -pc_type: PETSc's ILU(0)
    "shipped_poisson_in": None,                                                               # the reference's own poisson.in, unmodified
    "vcycle_65_mesh1": [o.replace("-mesh 0", "-mesh 1") for o in base(65, 4, 4, 0, it=1000)] + JAC,
    # cycle 8: KSPCG + PCMG with PETSc's default LU coarse solve / a Richardson coarse solve (src/solver.c:1884-1989)
    "pcmg_cg_129_l7": base(129, 7, 7, 8, it=100) + ["-ksp_type cg", "-ksp_rtol 1e-10"] + MGJ,
    "pcmg_cg_65_l3_lu15": base(65, 3, 3, 8, it=100) + ["-ksp_type cg", "-ksp_rtol 1e-10"] + MGJ,    # dense LU on the 15 x 15 coarsest grid
    "pcmg_richardson_65_l4_sor": base(65, 4, 4, 8, it=100) + ["-ksp_type richardson", "-ksp_rtol 1e-8", "-mg_levels_ksp_type richardson",
                                                                "-mg_levels_pc_type sor", "-mg_levels_ksp_max_it 2"],
    # several grids per level and the research cycles (row 11; src/solver.c:1577-2615)
    "vcycle_33_4grids_2levels": base(33, 4, 2, 0) + JAC,
    "icycle_33": base(33, 2, 1, 1) + JAC + ["-ksp_type richardson"],
    "ecycle_33": base(33, 2, 1, 2) + JAC,
    "d1cycle_33": base(33, 2, 1, 3) + JAC,
    "d2cycle_33": base(33, 2, 1, 4) + JAC,
    "d1pscycle_33": base(33, 2, 1, 7) + JAC,
    # -moreNorm 1: the per-grid residual monitors (KSPMonitorSet + VecGetSubVector on index sets, src/solver.c:2200-2240)
    "d1cycle_33_morenorm": [o.replace("-moreNorm 0", "-moreNorm 1") for o in base(33, 2, 1, 3, it=20)] + JAC,
    "d1pscycle_33_map0_sor_morenorm": [o.replace("-moreNorm 0", "-moreNorm 1").replace("-map 2", "-map 0")
                                       for o in base(33, 2, 1, 7, it=20, v="2,2")] + ["-pc_type sor"],
    "additive_33_l3": base(33, 3, 3, 9) + JAC,
    "additive2_33_l2": base(33, 2, 2, 10) + JAC,
}


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


# ------------------------------------------------------------------ CPU: build, provenance of the symbols, loud failure
@needs_ref
def test_refsolver_builds_from_the_unmodified_reference_sources(tmp_path):
    importlib.import_module("multigrid-petsc_b200").build()
    subprocess.run(["make", "-s", "-B", "-C", HOST, "refsolver"], check=True)
    assert os.path.exists(BIN)
    syms = subprocess.run(["nm", BIN], capture_output=True, text=True, check=True).stdout
    have = {line.split()[-1] for line in syms.splitlines() if " T " in line}
    undefined = {line.split()[-1] for line in syms.splitlines() if " U " in line}
    # every cycle of the reference's solver.c is in the binary, compiled from the reference's own file
    for name in ("main", "Solve", "Assemble", "MultigridVcycle", "MultigridIcycle", "MultigridEcycle", "MultigridD1cycle", "MultigridD2cycle",
                 "MultigridD1PScycle", "MultigridPetscPCMG", "MultigridAdditive", "MultigridAdditive2", "fillJacobians", "Res", "Pro"):
        assert name in have, name
    # PETSc's surface is the product's layer, and its arithmetic is the engine's: the Mat / Vec kernels come from libmgb200.so
    for name in ("KSPSolve", "MatMult", "VecAXPY", "MatSOR", "PCApply", "MatSetValue"):
        assert name in have, name
    for name in ("mgb_dcsr_mult", "mgb_dvec_axpy", "mgb_dcsr_sor", "mgb_dvec_dot", "mgb_dcsr_ilu0_solve"):
        assert name in undefined, name
    assert not [s for s in have | undefined if s.startswith(("mgo_", "MiniPetsc"))]          # nothing of the CPU oracle
    (tmp_path / "poisson.in").write_text("\n".join(CASES["vcycle_129_l4_jacobi"]) + "\n")
    if not _has_gpu():
        out = subprocess.run([BIN], cwd=tmp_path, capture_output=True, text=True)
        assert out.returncode != 0                                     # no CPU fallback: a missing GPU is a loud error
        assert "no CUDA device" in out.stderr and "no CPU fallback" in out.stderr


def test_sparse_abi_is_exported():
    """every entry point include/mgb200_sparse.h declares is exported by lib/libmgb200.so (no compute calls)"""
    mgb = importlib.import_module("multigrid-petsc_b200")
    mgb.build()
    hdr = open(os.path.join(ROOT, "include", "mgb200_sparse.h")).read()
    names = sorted(set(re.findall(r"\b(mgb_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 30
    lib = mgb.engine_lib()
    assert not [n for n in names if not hasattr(lib, n)]


# ------------------------------------------------------------------ GPU: the reference's solver.c on the B200 against its CPU run
def _run(exe, opts, d, env=None):
    os.makedirs(d, exist_ok=True)
    if opts is not None:
        with open(os.path.join(d, "poisson.in"), "w") as f:
            f.write("\n".join(opts) + "\n")
    out = subprocess.run([exe], cwd=d, capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, (exe, out.stderr[-1500:], out.stdout[-500:])
    it = int(re.search(r"Number of iterations:\s+(\d+)", out.stdout).group(1))
    rd = np.array([float(t) for t in open(os.path.join(d, "rData.dat")).read().split()])
    ed = np.array([float(t) for t in open(os.path.join(d, "eData.dat")).read().split()])
    u = np.loadtxt(os.path.join(d, "uData.dat"), ndmin=2)
    return it, rd, ed, u, out.stdout


SHIPPED = """# the option values of the reference's own poisson.in (no -pc_type: PETSc's default ILU(0))
-npts 17
-mesh 0
-iter 100000
-grids 2
-levels 2
-cycle 0   # V-cycle
-map 2
-v 3,3
-moreNorm 0
"""


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_reference_solver_on_the_gpu_matches_its_cpu_run(name, tmp_path):
    if not os.path.exists(BIN):
        pytest.skip("lib/poisson_petsc_b200 is not built (needs /root/reference in the build container)")
    ref = ref_binary_path()
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/poisson_ref is not built")
    opts = CASES[name]
    for d in ("gpu", "cpu"):
        os.makedirs(tmp_path / d)
        if opts is None:
            (tmp_path / d / "poisson.in").write_text(SHIPPED)          # what the reference ships (ref: poisson.in:1-14): no -pc_type
    env = dict(os.environ, OMP_NUM_THREADS=str(min(os.cpu_count() or 1, 8)))
    it_g, rd_g, ed_g, u_g, so = _run(BIN, opts, str(tmp_path / "gpu"))
    it_c, rd_c, ed_c, u_c, _ = _run(ref, opts, str(tmp_path / "cpu"), env)
    assert it_g == it_c                                                # equal iteration counts
    assert rd_g.shape == rd_c.shape and u_g.shape == u_c.shape
    fin = np.isfinite(rd_c)
    assert np.array_equal(fin, np.isfinite(rd_g))
    assert np.allclose(rd_g[fin], rd_c[fin], rtol=RTOL, atol=2.0 ** -52)     # residual history
    scale = max(np.abs(u_c).max(), 1e-300)
    assert np.abs(u_g - u_c).max() <= RTOL * scale                     # solution
    assert np.allclose(ed_g, ed_c, rtol=1e-9, atol=1e-12 * scale)
    # same operations in the same order, same reduction blocks: the files come out identical
    files = ["rData.dat", "uData.dat", "eData.dat"] + sorted(f for f in os.listdir(tmp_path / "cpu") if re.fullmatch(r"r(Global|Grid\d+)\.dat", f))
    if "morenorm" in name:
        assert len(files) >= 5                                         # rGlobal.dat and one rGrid<k>.dat per grid were written
    for f in files[3:]:
        a = np.array([float(t) for t in open(tmp_path / "gpu" / f).read().split()])
        b = np.array([float(t) for t in open(tmp_path / "cpu" / f).read().split()])
        assert a.shape == b.shape and np.allclose(a, b, rtol=RTOL, atol=2.0 ** -52, equal_nan=True), f   # (unfilled history slots print as nan in both)
    same = all(open(tmp_path / "gpu" / f).read() == open(tmp_path / "cpu" / f).read() for f in files)
    assert same, "files differ (within tolerance): the GPU layer no longer reproduces the checker's operation order"


@pytest.mark.gpu
def test_reference_solver_really_ran_on_the_gpu(tmp_path):
    """the sparse C-ABI counts its kernel launches (mgb_sparse_launch_count, also printed by KSPView): vector and matrix
    operations are CUDA kernels, not host loops -- create / set / norm launch at least three"""
    if not os.path.exists(BIN):
        pytest.skip("lib/poisson_petsc_b200 is not built")
    import ctypes as C
    mgb = importlib.import_module("multigrid-petsc_b200")
    lib = mgb.engine_lib()
    lib.mgb_sparse_launch_count.restype = C.c_longlong
    n0 = lib.mgb_sparse_launch_count()
    v = C.c_void_p()
    assert lib.mgb_dvec_create(1000, C.byref(v)) == 0
    lib.mgb_dvec_set.argtypes = [C.c_void_p, C.c_double]
    assert lib.mgb_dvec_set(v, 2.0) == 0
    out = C.c_double()
    lib.mgb_dvec_norm.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
    assert lib.mgb_dvec_norm(v, 1, C.byref(out)) == 0
    assert out.value == np.sqrt(4000.0)
    assert lib.mgb_sparse_launch_count() >= n0 + 3
    lib.mgb_dvec_destroy.argtypes = [C.c_void_p]
    lib.mgb_dvec_destroy(v)

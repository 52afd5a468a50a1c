"""The drop-in claim, proven by building and running it.

lib/poisson_dropin = the reference's UNMODIFIED src/poisson.c, mesh.c, problem.c, matbuild.c, array.c compiled in place
against the reference's own headers + the product-side options/print shim host/petsc_shim/petscksp.h, linked with
host/solver_b200.c (-DPB_USE_REFERENCE_HEADERS) in the place of src/solver.c (recipe: host/Makefile `dropin`; ref:
src/poisson.c:27-138).  The CPU tests build it (build container only: /root/reference is needed for the sources) and
check that it fails loudly without a GPU; the GPU tests run the built binary -- which travels to the GPU box like every
other built file -- against the committed goldens of the reference binary.
"""
import ctypes as C
import importlib
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "multigrid-petsc_b200", "host")
DROPIN = os.path.join(ROOT, "multigrid-petsc_b200", "lib", "poisson_dropin")
REF = "/root/reference"
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))
RTOL, RNORM_ATOL = 1e-10, 2.0 ** -52

needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="reference sources not present (GPU box)")


def options_file(opts):
    import re
    lines = []
    for t in opts.split():
        if t.startswith("-") and not re.fullmatch(r"-[0-9.].*", t):
            lines.append(t)
        else:
            lines[-1] += " " + t
    return "\n".join(lines) + "\n"


def _hex(v):
    return np.array([float.fromhex(x) for x in v])


# ------------------------------------------------------------------ CPU: build + loud failure
@needs_ref
def test_dropin_builds_from_the_unmodified_reference_sources(tmp_path):
    importlib.import_module("multigrid-petsc_b200").build()            # libmgb200.so must exist to link against
    subprocess.run(["make", "-s", "-B", "-C", HOST, "dropin"], check=True)
    assert os.path.exists(DROPIN)
    syms = subprocess.run(["nm", DROPIN], capture_output=True, text=True, check=True).stdout
    have = {line.split()[-1] for line in syms.splitlines() if " T " in line}
    # main, PrintInfo and the L2 layer come from the reference's own files, the solver.h entry points from solver_b200.c
    for name in ("main", "PrintInfo", "SetUpMesh", "SetUpIndices", "mapping", "GridTransferOperators", "SetUpProblem",
                 "SetUpSolver", "Assemble", "Solve", "Postprocessing", "DestroySolver", "SetUpPostProcess", "DestroyPostProcess"):
        assert name in have, name
    # nothing of the CPU oracle / mini-PETSc is linked in: no KSP / Mat / Vec arithmetic symbols at all
    assert not [s for s in have if s.startswith(("KSP", "MatMult", "VecAXPY", "PCApply", "mgo_"))]
    (tmp_path / "poisson.in").write_text(options_file(GOLD["n17_l2_jacobi"]["options"]))
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        out = subprocess.run([DROPIN], cwd=tmp_path, capture_output=True, text=True)
        assert out.returncode != 0                                     # no CPU fallback: a missing GPU is a loud error
        assert "no CUDA device" in out.stderr and "no CPU fallback" in out.stderr


# ------------------------------------------------------------------ CPU: pb_setup.c against the reference's object code
@needs_ref
@pytest.mark.parametrize("npts,levels,meshtype", [(17, 2, 0), (129, 4, 0), (101, 3, 0), (65, 4, 1), (65, 4, 2)])
def test_pb_setup_matches_the_reference_object_code(npts, levels, meshtype):
    """host/pb_setup.c (the product's restatement of src/mesh.c, problem.c, matbuild.c) against oracle/_ref/libref_l2.so
    (= those reference files compiled unmodified), value by value: coordinates, h, index maps, ranges, R/P stencils,
    metrics -> OpA coefficients, right-hand side and exact solution through the function pointers."""
    from oracle import ref_l2_path
    import test_oracle_vs_ref as T
    if not os.path.exists(ref_l2_path()):
        pytest.skip("oracle/_ref/libref_l2.so not built")
    mgb = importlib.import_module("multigrid-petsc_b200")
    mgb.engine_lib()                                                   # libpoisson_b200.so links libmgb200.so

    def setup(path, mapstyle):
        L = C.CDLL(path)
        mesh, ind, op, prob = T.Mesh(), T.Indices(), T.Operator(), T.Problem()
        mesh.n[0] = mesh.n[1] = npts
        mesh.bounds[0], mesh.bounds[1], mesh.bounds[2], mesh.bounds[3] = 0.0, 1.0, 0.0, 1.0
        L.SetUpProblem(C.byref(prob))
        L.SetUpMesh(C.byref(mesh), C.c_int(meshtype))
        ind.levels = ind.totalGrids = levels
        ind.coarseningFactor = 2
        L.SetUpIndices(C.byref(mesh), C.byref(ind))
        L.mapping(C.byref(ind), C.c_int(mapstyle))
        L.SetUpOperator(C.byref(ind), C.byref(op))
        L.GridTransferOperators.argtypes = [T.Operator, T.Indices]
        L.GridTransferOperators(op, ind)
        return L, mesh, ind, op, prob

    rng = np.random.default_rng(3)
    for mapstyle in (0, 1, 2):
        _, m_r, i_r, o_r, p_r = setup(ref_l2_path(), mapstyle)
        _, m_p, i_p, o_p, p_p = setup(mgb.HOST_LIB, mapstyle)
        for d in range(2):
            assert [m_r.coord[d][k] for k in range(npts)] == [m_p.coord[d][k] for k in range(npts)]
        assert m_r.h == m_p.h or (m_r.h != m_r.h and m_p.h != m_p.h)
        for l in range(levels):
            a, b = i_r.level[l], i_p.level[l]
            assert a.grids == b.grids == 1 and a.gridId[0] == b.gridId[0]
            ni, nj = a.grid[0].ni, a.grid[0].nj
            assert (ni, nj) == (b.grid[0].ni, b.grid[0].nj)
            assert (a.h[0][0], a.h[0][1]) == (b.h[0][0], b.h[0][1])
            assert (a.ranges[0], a.ranges[1]) == (b.ranges[0], b.ranges[1])
            if b.grid[0].data:                                         # the product builds the maps on request only (-pb_index_maps 1)
                assert np.array_equal(np.ctypeslib.as_array(a.grid[0].data, shape=(ni * nj,)),
                                      np.ctypeslib.as_array(b.grid[0].data, shape=(ni * nj,)))
            f = 2 ** l
            met_r, met_p, As_r, As_p = (C.c_double * 5)(), (C.c_double * 5)(), (C.c_double * 5)(), (C.c_double * 5)()
            for _ in range(25):
                i, j = int(rng.integers(ni)), int(rng.integers(nj))
                x, y = m_r.coord[0][f * (j + 1)], m_r.coord[1][f * (i + 1)]
                m_r.MetricCoefficients(C.addressof(m_r), x, y, met_r)
                m_p.MetricCoefficients(C.addressof(m_p), x, y, met_p)
                assert list(met_r) == list(met_p)
                h = (C.c_double * 2)(a.h[0][0], a.h[0][1])
                p_r.OpA(As_r, met_r, h)
                p_p.OpA(As_p, met_p, h)
                assert list(As_r) == list(As_p)
        for k in range(9):
            assert o_r.res[0].data[k] == o_p.res[0].data[k] and o_r.pro[0].data[k] == o_p.pro[0].data[k]
        for _ in range(50):
            x, y = float(rng.uniform()), float(rng.uniform())
            assert p_r.Ffunc(x, y) == p_p.Ffunc(x, y) and p_r.SOLfunc(x, y) == p_p.SOLfunc(x, y)


# ------------------------------------------------------------------ GPU: the built drop-in against the reference's goldens
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["n17_l2_jacobi", "n129_l4_jacobi", "n129_l7_jacobi_v21", "n65_l4_mesh1_jacobi", "n129_l7_cg_mg",
                                  "n129_l4_cg_mg_jcoarse", "n1025_l10_jacobi"])
def test_dropin_binary_reproduces_the_reference_goldens(name, tmp_path):
    if not os.path.exists(DROPIN):
        pytest.skip("lib/poisson_dropin is not built (needs /root/reference in the build container)")
    import hashlib
    g = GOLD[name]
    (tmp_path / "poisson.in").write_text(options_file(g["options"]))
    out = subprocess.run([DROPIN], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-1000:]
    # the banner is printed by the reference's own PrintInfo (src/poisson.c:165-214)
    import re
    it = int(re.search(r"Number of iterations:\s+(\d+)", out.stdout).group(1))
    assert it == g["num_iter"]
    want = _hex(g["rnorm_hex"])
    rdat = np.array([float(t) for t in (tmp_path / "rData.dat").read_text().split()])
    ok = ~np.isnan(want[: len(rdat)])
    assert len(rdat) == it + 1
    assert np.allclose(rdat[ok], want[: len(rdat)][ok], rtol=RTOL, atol=RNORM_ATOL)
    edat = np.array([float(t) for t in (tmp_path / "eData.dat").read_text().split()])
    assert np.allclose(edat, _hex(g["error_hex"]), rtol=RTOL, atol=0.0)
    if "-cycle 0" in g["options"]:
        u = np.loadtxt(tmp_path / "uData.dat", ndmin=2)
        assert hashlib.sha256(np.ascontiguousarray(u, dtype="<f8").tobytes()).hexdigest() == g["u_sha256"]   # bit for bit

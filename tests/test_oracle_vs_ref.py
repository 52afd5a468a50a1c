"""CPU tests: the oracle against the reference ITSELF, where the reference is present.

oracle/_ref/ holds the reference's own unmodified sources compiled in place over oracle/minipetsc
(oracle/Makefile `ref`).  It exists in the build container (and travels to the GPU box as a binary);
when it is absent these tests skip and test_oracle_golden.py (committed vectors) carries the check.
"""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import Oracle, ref_binary_path, ref_l2_path

needs_ref_bin = pytest.mark.skipif(not os.path.exists(ref_binary_path()), reason="oracle/_ref/poisson_ref not built")
needs_ref_l2 = pytest.mark.skipif(not os.path.exists(ref_l2_path()), reason="oracle/_ref/libref_l2.so not built")


# ---- ctypes mirrors of the reference's structs (include/array.h:31-40, mesh.h:21-27, solver.h:17-37)
class ArrayInt2d(C.Structure):
    _fields_ = [("ni", C.c_int), ("nj", C.c_int), ("data", C.POINTER(C.c_int))]


class Array2d(C.Structure):
    _fields_ = [("ni", C.c_int), ("nj", C.c_int), ("data", C.POINTER(C.c_double))]


METRIC_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_double, C.c_double, C.POINTER(C.c_double))


class Mesh(C.Structure):
    _fields_ = [("n", C.c_int * 2), ("bounds", C.c_double * 4), ("coord", C.POINTER(C.POINTER(C.c_double))),
                ("h", C.c_double), ("MetricCoefficients", METRIC_FN)]


class Level(C.Structure):
    _fields_ = [("grids", C.c_int), ("gridId", C.POINTER(C.c_int)), ("h", C.POINTER(C.c_double * 2)),
                ("ranges", C.POINTER(C.c_int)), ("global_", ArrayInt2d), ("grid", C.POINTER(ArrayInt2d))]


class Indices(C.Structure):
    _fields_ = [("levels", C.c_int), ("totalGrids", C.c_int), ("coarseningFactor", C.c_int),
                ("level", C.POINTER(Level))]


class Operator(C.Structure):
    _fields_ = [("totalGrids", C.c_int), ("res", C.POINTER(Array2d)), ("pro", C.POINTER(Array2d))]


class Problem(C.Structure):
    _fields_ = [("Ffunc", C.CFUNCTYPE(C.c_double, C.c_double, C.c_double)),
                ("SOLfunc", C.CFUNCTYPE(C.c_double, C.c_double, C.c_double)),
                ("OpA", C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)))]


def _ref_setup(npts, levels, meshtype, mapstyle):
    L = C.CDLL(ref_l2_path())
    mesh, ind, op, prob = Mesh(), Indices(), Operator(), Problem()
    mesh.n[0] = mesh.n[1] = npts
    mesh.bounds[0], mesh.bounds[1], mesh.bounds[2], mesh.bounds[3] = 0.0, 1.0, 0.0, 1.0
    L.SetUpProblem(C.byref(prob))
    L.SetUpMesh(C.byref(mesh), C.c_int(meshtype))
    ind.levels = ind.totalGrids = levels
    ind.coarseningFactor = 2
    L.SetUpIndices(C.byref(mesh), C.byref(ind))
    L.mapping(C.byref(ind), C.c_int(mapstyle))
    L.SetUpOperator(C.byref(ind), C.byref(op))
    L.GridTransferOperators.argtypes = [Operator, Indices]
    L.GridTransferOperators(op, ind)
    return L, mesh, ind, op, prob


@needs_ref_l2
@pytest.mark.parametrize("npts,levels,meshtype", [(17, 2, 0), (129, 4, 0), (101, 3, 0), (65, 4, 1), (65, 4, 2)])
def test_l2_matches_reference_object_code(npts, levels, meshtype):
    """coords, h, index maps (all three -map styles), R/P stencils, metrics+OpA: bit-equal to the
    reference's compiled src/mesh.c, src/matbuild.c, src/problem.c."""
    o = Oracle(f"-npts {npts} -mesh {meshtype} -iter 1 -grids {levels} -levels {levels} -pc_type jacobi")
    for mapstyle in (0, 1, 2):
        L, mesh, ind, op, prob = _ref_setup(npts, levels, meshtype, mapstyle)
        for d in range(2):
            ref = np.array([mesh.coord[d][k] for k in range(npts)])
            assert np.array_equal(ref, o.coords(d, npts))
        for l in range(levels):
            lev = ind.level[l]
            assert lev.grids == 1
            ni, nj = lev.grid[0].ni, lev.grid[0].nj
            assert (ni, nj) == o.dims(l)
            assert np.array_equal(np.array([lev.h[0][0], lev.h[0][1]]), o.level_h(l))
            g2G = np.ctypeslib.as_array(lev.grid[0].data, shape=(ni * nj,))
            assert np.array_equal(g2G, o.grid_to_global(l))          # natural numbering for every style
            assert np.array_equal(g2G, np.arange(ni * nj))
            G2g = np.ctypeslib.as_array(lev.global_.data, shape=(ni * nj, 3))
            assert np.array_equal(G2g, o.global_to_grid(l))
            assert lev.ranges[0] == 0 and lev.ranges[1] == ni * nj
        res = np.ctypeslib.as_array(op.res[0].data, shape=(3, 3))
        pro = np.ctypeslib.as_array(op.pro[0].data, shape=(3, 3))
        assert np.array_equal(res, o.stencil(0)) and np.array_equal(pro, o.stencil(1))
        assert np.array_equal(res, np.array([[.0625, .125, .0625], [.125, .25, .125], [.0625, .125, .0625]]))
        assert np.array_equal(pro, np.array([[.25, .5, .25], [.5, 1, .5], [.25, .5, .25]]))
    # OpA through the reference's own function pointers at a sample of points of every level
    rng = np.random.default_rng(1)
    met = (C.c_double * 5)()
    As = (C.c_double * 5)()
    for l in range(levels):
        ni, nj = o.dims(l)
        f = 2 ** l
        for _ in range(20):
            i, j = int(rng.integers(ni)), int(rng.integers(nj))
            x, y = mesh.coord[0][f * (j + 1)], mesh.coord[1][f * (i + 1)]
            mesh.MetricCoefficients(C.addressof(mesh), x, y, met)
            h = (C.c_double * 2)(ind.level[l].h[0][0], ind.level[l].h[0][1])
            prob.OpA(As, met, h)
            assert np.array_equal(np.array(list(As)), o.opA(l, i, j))
    if meshtype == 0 and npts == 17:
        assert np.array_equal(o.opA(0, 3, 3), np.array([256.0, 256.0, -1024.0, 256.0, 256.0]))
    # RHS and exact solution through the reference's function pointers
    b = o.to_grid(0, o.vec(0, 0))
    for (i, j) in [(0, 0), (3, 5), (b.shape[0] - 1, b.shape[1] - 1)]:
        assert b[i, j] == prob.Ffunc(mesh.coord[0][j + 1], mesh.coord[1][i + 1])
    o.close()


@needs_ref_bin
@pytest.mark.parametrize("opts", [
    "-npts 17 -mesh 0 -iter 100000 -grids 2 -levels 2 -cycle 0 -map 0 -v 3,3 -pc_type jacobi -ksp_richardson_scale 0.7",
    "-npts 33 -mesh 0 -iter 100000 -grids 3 -levels 3 -cycle 0 -map 1 -v 2,4 -pc_type sor -pc_sor_omega 1.3",
    "-npts 65 -mesh 1 -iter 300 -grids 5 -levels 5 -cycle 0 -map 2 -v 1,2 -pc_type jacobi -ksp_richardson_scale 0.8",
    "-npts 65 -mesh 0 -iter 100 -grids 6 -levels 6 -cycle 8 -map 2 -v 3,3 -ksp_type cg -ksp_rtol 1e-9 "
    "-mg_levels_ksp_type richardson -mg_levels_pc_type sor -mg_levels_ksp_max_it 2",
])
def test_files_written_by_the_reference_are_reproduced_byte_for_byte(opts):
    """Same options -> the oracle's uData/rData/eData/X/YgridData equal the reference's, as text."""
    with tempfile.TemporaryDirectory() as dref, tempfile.TemporaryDirectory() as dora:
        open(os.path.join(dref, "poisson.in"), "w").close()
        subprocess.run([ref_binary_path()] + opts.split(), cwd=dref, check=True, capture_output=True)
        o = Oracle(opts)
        o.solve()
        o.write_files(dora)
        o.close()
        for f in ("uData.dat", "rData.dat", "eData.dat", "XgridData.dat", "YgridData.dat"):
            assert open(os.path.join(dref, f)).read() == open(os.path.join(dora, f)).read(), f

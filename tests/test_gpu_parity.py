"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C-ABI
(include/mgb200.h via ctypes) or through the host C layer that mirrors the reference's solver.h; the CPU oracle
(oracle/) is only the checker.

Bars (BASELINE.json north_star): assembled operators bit-exact; vectors produced by the Jacobi / red-black /
transfer kernels bit-exact against the oracle on the same inputs (the kernels keep PETSc's operation order and
never fuse multiply-add); norms, dots and residual histories within 1e-10 relative (tolerance written at each
assert; the reductions use a different, deterministic summation order); iteration counts equal.
"""
import hashlib
import importlib
import json
import os

import numpy as np
import pytest

from oracle import Oracle

pytestmark = pytest.mark.gpu

mgb = importlib.import_module("multigrid-petsc_b200")
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden.json")))
GOLD_U = np.load(os.path.join(HERE, "golden", "golden_u.npz"))
RTOL = 1e-10          # north_star: residual norms and solution within 1e-10 relative in fp64
# rnorm[] is normalised by rnorm[0]; an entry cannot be resolved below one fp64 epsilon of the initial residual
# (the recursively updated CG residual near 1e-11 differs by ~1e-20 absolute between summation orders of the dots)
RNORM_ATOL = 2.0 ** -52

JAC = "-pc_type jacobi -ksp_richardson_scale 0.8"


def base(npts, levels, cycle=0, mesh=0, it=1000, v="3,3", mp=2):
    return (f"-npts {npts} -mesh {mesh} -iter {it} -grids {levels} -levels {levels} "
            f"-cycle {cycle} -map {mp} -v {v} -moreNorm 0")


def _hex(v):
    return np.array([float.fromhex(x) for x in v])


# ------------------------------------------------------------------ assembled operators: bit-exact
@pytest.mark.parametrize("npts,levels,mesh", [(5, 1, 0), (9, 2, 0), (17, 2, 0), (33, 4, 0), (101, 3, 0), (129, 4, 0),
                                              (129, 7, 0), (65, 4, 1), (65, 4, 2), (257, 5, 0), (1025, 7, 0)])
def test_csr_bit_exact(npts, levels, mesh):
    opts = base(npts, levels, mesh=mesh) + " " + JAC
    o = Oracle(opts)
    s = mgb.Session(opts)
    try:
        for l in range(levels):
            for which in (mgb.MAT_A, mgb.MAT_RES, mgb.MAT_PRO):
                if which != mgb.MAT_A and l == levels - 1:
                    continue
                shape_o, ia_o, ja_o, va_o = o.csr(which, l)
                shape_g, ia_g, ja_g, va_g = s.engine.csr(which, l)
                assert shape_o == shape_g
                assert np.array_equal(ia_o, ia_g), (which, l)
                assert np.array_equal(ja_o, ja_g), (which, l)
                assert va_o.tobytes() == va_g.tobytes(), (which, l)      # bit for bit
    finally:
        s.close()
        o.close()


def test_rhs_bit_exact_and_csr_spmv_matches_matrix_free():
    opts = base(129, 4) + " " + JAC
    o = Oracle(opts)
    s = mgb.Session(opts)
    e = s.engine
    try:
        b_o = o.to_grid(0, o.vec(0, 0))
        assert b_o.tobytes() == e.get_vec(mgb.VEC_B, 0).tobytes()
        rng = np.random.default_rng(0)
        for l in range(4):
            ni, nj = e.dims(l)
            x = rng.uniform(-1, 1, (ni, nj))
            e.set_vec(mgb.VEC_U, l, x)
            e.apply(l, mgb.VEC_U, mgb.VEC_R)
            y_mf = e.get_vec(mgb.VEC_R, l)
            y_csr = e.csr_spmv(mgb.MAT_A, l, x.reshape(-1)).reshape(ni, nj)
            y_o = o.matmult(0, l, x.reshape(-1)).reshape(ni, nj)
            assert y_mf.tobytes() == y_o.tobytes()
            assert y_csr.tobytes() == y_o.tobytes()
    finally:
        s.close()
        o.close()


# ------------------------------------------------------------------ single kernels vs the oracle, random inputs
@pytest.mark.parametrize("npts,levels,mesh", [(17, 2, 0), (129, 4, 0), (101, 3, 0), (65, 4, 1), (65, 4, 2), (513, 5, 0)])
def test_kernels_bit_exact_vs_oracle(npts, levels, mesh):
    opts = base(npts, levels, mesh=mesh) + " " + JAC
    o = Oracle(opts)
    s = mgb.Session(opts)
    e = s.engine
    rng = np.random.default_rng(npts + mesh)
    try:
        for l in range(levels):
            ni, nj = e.dims(l)
            x = rng.uniform(-1, 1, (ni, nj))
            b = rng.uniform(-1, 1, (ni, nj))
            e.set_vec(mgb.VEC_U, l, x)
            e.set_vec(mgb.VEC_B, l, b)
            # residual r = b - A x
            e.residual(l)
            r_g = e.get_vec(mgb.VEC_R, l)
            r_o = o.residual(l, b.reshape(-1), x.reshape(-1)).reshape(ni, nj)
            assert r_g.tobytes() == r_o.tobytes()
            rn = e.residual_norm(l)
            assert rn == pytest.approx(o.norm2(r_o), rel=1e-13)
            assert e.norm2(mgb.VEC_B, l) == pytest.approx(o.norm2(b), rel=1e-13)
            assert e.dot(mgb.VEC_B, mgb.VEC_U, l) == pytest.approx(o.dot(b, x), rel=1e-10, abs=1e-12)
            # Jacobi smoother: nonzero and zero guess, 1..3 sweeps
            for nu, gz in ((1, False), (3, False), (2, False), (3, True), (1, True)):
                e.set_vec(mgb.VEC_U, l, x)
                e.smooth(l, mgb.jacobi(0.8), nu, gz)
                x_g = e.get_vec(mgb.VEC_U, l)
                x_o = o.smooth(l, b.reshape(-1), x.reshape(-1), nu, gz).reshape(ni, nj)
                assert x_g.tobytes() == x_o.tobytes(), (l, nu, gz)
            if l + 1 < levels:
                nci, ncj = e.dims(l + 1)
                # restriction, fused with the residual and from a stored residual
                e.set_vec(mgb.VEC_U, l, x)
                bc_o = o.matmult(1, l, r_o.reshape(-1)).reshape(nci, ncj)
                e.restrict(l, fused=True)
                assert e.get_vec(mgb.VEC_B, l + 1).tobytes() == bc_o.tobytes()
                e.zero_vec(mgb.VEC_B, l + 1)
                e.restrict(l, fused=False)        # uses VEC_R written by residual() above
                assert e.get_vec(mgb.VEC_B, l + 1).tobytes() == bc_o.tobytes()
                # prolongation + correction, both summation orders
                uc = rng.uniform(-1, 1, (nci, ncj))
                e.set_vec(mgb.VEC_U, l + 1, uc)
                e.set_vec(mgb.VEC_U, l, x)
                e.prolong(l, multadd=False)
                want = x.reshape(-1) + 1.0 * o.matmult(2, l, uc.reshape(-1))
                assert e.get_vec(mgb.VEC_U, l).tobytes() == want.reshape(ni, nj).tobytes()
                e.set_vec(mgb.VEC_U, l, x)
                e.prolong(l, multadd=True)
                want = o.matmultadd(2, l, uc.reshape(-1), x.reshape(-1))
                assert e.get_vec(mgb.VEC_U, l).tobytes() == want.reshape(ni, nj).tobytes()
    finally:
        s.close()
        o.close()


@pytest.mark.parametrize("extra", ["-pc_type sor", "-pc_type sor -pc_sor_omega 1.2", "-pc_type sor -pc_sor_forward",
                                   "-pc_type sor -pc_sor_backward -pc_sor_omega 0.9", "-pc_type sor -pc_sor_its 2"])
@pytest.mark.parametrize("npts,levels,mesh", [(17, 2, 0), (129, 3, 0), (65, 3, 2)])
def test_red_black_sor_bit_exact_vs_oracle(npts, levels, mesh, extra):
    opts = base(npts, levels, mesh=mesh, mp=3) + " " + extra
    o = Oracle(opts)
    s = mgb.Session(opts)
    e = s.engine
    sm = mgb.rbsor(omega=1.2 if "1.2" in extra else (0.9 if "0.9" in extra else 1.0),
                   sweep=mgb.SOR_FORWARD if "forward" in extra else (mgb.SOR_BACKWARD if "backward" in extra else mgb.SOR_SYMMETRIC),
                   its=2 if "its 2" in extra else 1)
    rng = np.random.default_rng(7)
    try:
        for l in range(levels):
            ni, nj = e.dims(l)
            x = rng.uniform(-1, 1, (ni, nj))
            b = rng.uniform(-1, 1, (ni, nj))
            e.set_vec(mgb.VEC_B, l, b)
            for nu, gz in ((1, False), (3, False), (3, True), (1, True)):
                e.set_vec(mgb.VEC_U, l, x)
                e.smooth(l, sm, nu, gz)
                x_g = e.get_vec(mgb.VEC_U, l)
                x_o = o.to_grid(l, o.smooth(l, o.from_grid(l, b), o.from_grid(l, x), nu, gz))
                assert x_g.tobytes() == x_o.tobytes(), (l, nu, gz)
    finally:
        s.close()
        o.close()


@pytest.mark.parametrize("npts,levels,mesh,extra", [
    (17, 2, 0, "-pc_type sor"), (33, 3, 0, "-pc_type sor -pc_sor_omega 1.3"), (65, 3, 2, "-pc_type sor -pc_sor_forward"),
    (65, 3, 1, "-pc_type sor -pc_sor_backward -pc_sor_omega 0.9"), (129, 4, 0, "-pc_type sor -pc_sor_its 2"),
    (257, 2, 0, "-pc_type sor"), (17, 2, 0, ""), (65, 3, 1, "-pc_type ilu -ksp_richardson_scale 0.9"), (257, 2, 0, "-pc_type ilu")])
def test_lexicographic_smoothers_bit_exact_vs_oracle(npts, levels, mesh, extra):
    """PETSc's natural-order smoothers as wavefronts (csrc/mgb_wave.cuh): MatSOR forward / backward / symmetric and the default
    ILU(0) inside Richardson, against the oracle's restatement of MatSOR / PCILU on the same random inputs, bit for bit.
    257 columns span three column panels (the inter-panel hand-over through HBM)."""
    opts = base(npts, levels, mesh=mesh) + " " + extra
    o = Oracle(opts)
    s = mgb.Session(opts)
    e = s.engine
    if "sor" in extra:
        sm = mgb.lexsor(omega=1.3 if "1.3" in extra else (0.9 if "0.9" in extra else 1.0),
                        sweep=mgb.SOR_FORWARD if "forward" in extra else (mgb.SOR_BACKWARD if "backward" in extra else mgb.SOR_SYMMETRIC),
                        its=2 if "its 2" in extra else 1)
    else:
        sm = mgb.ilu0(0.9 if "0.9" in extra else 1.0)
    rng = np.random.default_rng(11)
    try:
        for l in range(levels):
            ni, nj = e.dims(l)
            x = rng.uniform(-1, 1, (ni, nj))
            b = rng.uniform(-1, 1, (ni, nj))
            e.set_vec(mgb.VEC_B, l, b)
            for nu, gz in ((1, False), (3, False), (3, True), (1, True)):
                e.set_vec(mgb.VEC_U, l, x)
                e.smooth(l, sm, nu, gz)
                x_g = e.get_vec(mgb.VEC_U, l)
                x_o = o.to_grid(l, o.smooth(l, o.from_grid(l, b), o.from_grid(l, x), nu, gz))
                assert x_g.tobytes() == x_o.tobytes(), (l, nu, gz)
    finally:
        s.close()
        o.close()


# ------------------------------------------------------------------ end to end against the committed goldens
E2E = ["n17_l2_sor", "n17_l2_ilu_default", "n129_l4_sor", "n129_l4_sor_forward",      # PETSc's natural-order smoothers (wavefronts)
       "n17_l2_jacobi", "n17_l2_jacobi23", "n17_l1_jacobi", "n129_l4_jacobi", "n129_l7_jacobi", "n129_l7_jacobi_v21",
       "n101_l3_jacobi", "n65_l4_mesh1_jacobi", "n65_l4_mesh2_jacobi", "n129_l7_cg_mg", "n129_l4_cg_mg_jcoarse",
       "n129_l7_rich_mg_monitor", "n1025_l7_jacobi", "n1025_l10_jacobi",
       "n17_l2_rbsor", "n129_l4_rbsor", "n129_l7_rbsor", "n129_l7_rbsor_w12", "n1025_l7_rbsor"]


@pytest.mark.parametrize("graph", [0, 1])
@pytest.mark.parametrize("name", E2E)
def test_end_to_end_matches_golden(name, graph, tmp_path):
    g = GOLD[name]
    r = mgb.run_poisson(g["options"] + f" -mgb_graph {graph}", out_dir=str(tmp_path))
    assert r["num_iter"] == g["num_iter"]                                # equal iteration counts
    want = _hex(g["rnorm_hex"])
    ok = ~np.isnan(want)
    assert np.allclose(r["rnorm"][ok], want[ok], rtol=RTOL, atol=RNORM_ATOL)   # per-cycle residual norms, 1e-10 relative
    err = _hex(g["error_hex"])
    assert np.allclose(r["error"], err, rtol=RTOL, atol=0.0)
    u = r["u"]
    assert list(u.shape) == g["u_shape"]
    if name in GOLD_U.files:
        scale = np.abs(GOLD_U[name]).max()
        assert np.abs(u - GOLD_U[name]).max() <= RTOL * scale             # solution within 1e-10 relative
    # the files the reference writes (src/solver.c:1331-1354) carry the same numbers
    rdat = np.array([float(t) for t in open(tmp_path / "rData.dat").read().split()])
    assert np.array_equal(rdat[ok], r["rnorm"][ok])
    edat = np.array([float(t) for t in open(tmp_path / "eData.dat").read().split()])
    assert np.array_equal(edat, r["error"])
    udat = np.loadtxt(tmp_path / "uData.dat", ndmin=2)
    assert np.array_equal(udat, u)


@pytest.mark.parametrize("name", ["n129_l4_jacobi", "n129_l7_jacobi", "n101_l3_jacobi", "n65_l4_mesh1_jacobi",
                                  "n129_l7_rbsor", "n1025_l10_jacobi", "n1025_l7_rbsor", "n17_l2_sor", "n129_l4_sor",
                                  "n17_l2_ilu_default"])
def test_cycle0_solution_is_bit_exact(name):
    """Cycle 0 with Jacobi / red-black smoothing uses no reduction inside the iteration, so the solution vector
    itself is reproduced bit for bit (only the logged norms differ in summation order)."""
    g = GOLD[name]
    r = mgb.run_poisson(g["options"])
    assert r["num_iter"] == g["num_iter"]
    assert hashlib.sha256(np.ascontiguousarray(r["u"], dtype="<f8").tobytes()).hexdigest() == g["u_sha256"]


def test_unsupported_configurations_fail_loudly():
    for opts in (base(17, 2) + " -pc_type lu",                    # a PC the engine does not have
                 base(129, 4) + " -pc_type sor -mgb_ranks 2 -mgb_emulate 1 -mgb_agglomerate 31",   # lexicographic SOR on strips
                 base(17, 2, mp=3),                               # default ILU(0) on the red-black numbering
                 base(17, 2, cycle=1) + " " + JAC,                # research cycle
                 base(17, 2, cycle=8) + " -ksp_type cg",          # Chebyshev default smoother
                 "-npts 17 -iter 10 -grids 3 -levels 2 " + JAC):
        with pytest.raises(mgb.MgbError):
            mgb.run_poisson(opts)


# ------------------------------------------------------------------ full-size properties (no oracle at these sizes)
@pytest.mark.parametrize("npts,levels", [(4097, 12), (8193, 13)])
def test_full_size_properties(npts, levels):
    opts = base(npts, levels, it=40) + " " + JAC + " -mgb_csr 0"
    s = mgb.Session(opts)
    e = s.engine
    try:
        n = npts - 2
        # linearity of the stencil on the fine level: A(x + 2y) == A x + 2 A y up to rounding of the sums
        rng = np.random.default_rng(1)
        x = rng.uniform(-1, 1, (n, n))
        y = rng.uniform(-1, 1, (n, n))
        e.set_vec(mgb.VEC_U, 0, x); e.apply(0, mgb.VEC_U, mgb.VEC_R); ax = e.get_vec(mgb.VEC_R, 0)
        e.set_vec(mgb.VEC_U, 0, y); e.apply(0, mgb.VEC_U, mgb.VEC_R); ay = e.get_vec(mgb.VEC_R, 0)
        e.set_vec(mgb.VEC_U, 0, x + 2 * y); e.apply(0, mgb.VEC_U, mgb.VEC_R); axy = e.get_vec(mgb.VEC_R, 0)
        h2 = float(npts - 1) ** 2
        assert np.abs(axy - (ax + 2 * ay)).max() <= 64 * np.finfo(float).eps * 8 * h2
        # exact row of the operator against numpy on a strip (bit for bit, same operation order)
        c = np.float64(1.0) / (np.float64(1.0 / (n + 1)) * np.float64(1.0 / (n + 1)))
        i = n // 2
        want = (c * x[i - 1, 1:-1] + c * x[i, :-2]) + (-2.0 * (c + c)) * x[i, 1:-1]
        want = (want + c * x[i, 2:]) + c * x[i + 1, 1:-1]
        assert np.array_equal(ax[i, 1:-1], want)
        del x, y, ax, ay, axy
        # solve: textbook convergence (coarsest 1x1 or 3x3), monotone history, discretisation error pi^2 h^2 / 12
        r = s.solve(want_u=False)
        assert 6 <= r["num_iter"] <= 12
        assert np.all(np.diff(r["rnorm"]) < 0)
        assert r["rnorm"][0] == 1.0 and r["rnorm"][-1] <= 1e-7
        h = 1.0 / (npts - 1)
        # at these sizes the algebraic error left by rtol 1e-7 is of the order of the discretisation error
        # pi^2 h^2 / 12 itself (and of opposite sign), so only the magnitude is checked
        assert 0.0 < r["error"][0] < 2.0 * np.pi ** 2 * h * h / 12.0
    finally:
        s.close()


def test_solver_is_deterministic_run_to_run():
    opts = base(1025, 10) + " " + JAC
    a = mgb.run_poisson(opts)
    b = mgb.run_poisson(opts)
    assert a["num_iter"] == b["num_iter"]
    assert a["rnorm"].tobytes() == b["rnorm"].tobytes()
    assert a["u"].tobytes() == b["u"].tobytes()


# ------------------------------------------------------------------ row strips (several ranks emulated on one GPU)
# The strip engine runs the same kernels and the same push/flag protocol as the multi-process run, with all strips
# held by one process and launched in lock step (mgb_config.emulate).  Jacobi and red-black smoothing do not depend
# on the partition, so the solution must be bit-identical to the single-strip one (SURVEY.md 8e "Determinism").
@pytest.mark.parametrize("ranks,aggl", [(2, 31), (3, 31), (4, 31), (2, 15), (4, 63)])
@pytest.mark.parametrize("name", ["n129_l4_jacobi", "n129_l7_jacobi", "n129_l7_rbsor", "n129_l7_rbsor_w12", "n101_l3_jacobi",
                                  "n65_l4_mesh1_jacobi"])
def test_strips_emulated_cycle0_bit_exact(name, ranks, aggl):
    g = GOLD[name]
    if name.startswith("n65") and aggl == 63:
        pytest.skip("finest level not above the agglomeration threshold")
    if name.startswith("n101") and aggl == 15 and ranks > 2:
        pytest.skip("strips too thin")
    r = mgb.run_poisson(g["options"] + f" -mgb_ranks {ranks} -mgb_emulate 1 -mgb_agglomerate {aggl}")
    assert r["num_iter"] == g["num_iter"]
    want = _hex(g["rnorm_hex"])
    assert np.allclose(r["rnorm"], want, rtol=RTOL, atol=RNORM_ATOL)
    assert hashlib.sha256(np.ascontiguousarray(r["u"], dtype="<f8").tobytes()).hexdigest() == g["u_sha256"]
    assert np.allclose(r["error"], _hex(g["error_hex"]), rtol=RTOL, atol=0.0)


@pytest.mark.parametrize("name", ["n17_l2_rbsor", "n129_l4_rbsor", "n129_l7_rbsor", "n129_l7_rbsor_w12", "n1025_l7_rbsor"])
def test_red_black_fused_legs_on_every_level(name, monkeypatch):
    """Red-black SOR normally takes the fused legs on bandwidth-bound levels only (>= 2047 rows on one GPU); with the threshold
    at 0 every level of the small goldens runs through k_jfused<.., SMK = 1>: same bits."""
    monkeypatch.setenv("MGB_RB_FUSE_MIN_ROWS", "0")
    g = GOLD[name]
    r = mgb.run_poisson(g["options"])
    assert r["num_iter"] == g["num_iter"]
    assert np.allclose(r["rnorm"], _hex(g["rnorm_hex"]), rtol=RTOL, atol=RNORM_ATOL)
    assert hashlib.sha256(np.ascontiguousarray(r["u"], dtype="<f8").tobytes()).hexdigest() == g["u_sha256"]


@pytest.mark.parametrize("name,ranks,aggl", [("n129_l7_jacobi", 2, 31), ("n129_l7_jacobi", 4, 31), ("n1025_l10_jacobi", 4, 0),
                                             ("n129_l7_cg_mg", 2, 31)])
def test_strips_emulated_separate_exchange_launches(name, ranks, aggl, monkeypatch):
    """The fused legs normally push / wait for their ghost rows, the gathered right-hand side and the broadcast correction
    themselves (FusedComm, csrc/mgb_fused.cuh); MGB_INKERNEL_HALO=0 selects the older protocol with one k_xfer launch per
    exchange (csrc/mgb_halo.cuh).  Both must give the single-strip bits."""
    monkeypatch.setenv("MGB_INKERNEL_HALO", "0")
    g = GOLD[name]
    r = mgb.run_poisson(g["options"] + f" -mgb_ranks {ranks} -mgb_emulate 1" + (f" -mgb_agglomerate {aggl}" if aggl else ""))
    assert r["num_iter"] == g["num_iter"]
    want = _hex(g["rnorm_hex"])
    ok = ~np.isnan(want)
    assert np.allclose(r["rnorm"][ok], want[ok], rtol=RTOL, atol=RNORM_ATOL)
    if "-cycle 8" not in g["options"]:
        assert hashlib.sha256(np.ascontiguousarray(r["u"], dtype="<f8").tobytes()).hexdigest() == g["u_sha256"]


@pytest.mark.parametrize("ranks", [2, 4])
@pytest.mark.parametrize("name", ["n129_l7_cg_mg", "n129_l4_cg_mg_jcoarse", "n129_l7_rich_mg_monitor"])
def test_strips_emulated_cycle8(name, ranks):
    g = GOLD[name]
    r = mgb.run_poisson(g["options"] + f" -mgb_ranks {ranks} -mgb_emulate 1 -mgb_agglomerate 31")
    assert r["num_iter"] == g["num_iter"]
    want = _hex(g["rnorm_hex"])
    ok = ~np.isnan(want)
    assert np.allclose(r["rnorm"][ok], want[ok], rtol=RTOL, atol=RNORM_ATOL)
    if name in GOLD_U.files:
        assert np.abs(r["u"] - GOLD_U[name]).max() <= RTOL * np.abs(GOLD_U[name]).max()


@pytest.mark.parametrize("ranks", [2, 4, 8])
def test_strips_emulated_large(ranks):
    """1025^2, 10 levels, default agglomeration threshold (511): levels 0 and 1 distributed."""
    g = GOLD["n1025_l10_jacobi"]
    r = mgb.run_poisson(g["options"] + f" -mgb_ranks {ranks} -mgb_emulate 1")
    assert r["num_iter"] == g["num_iter"]
    assert hashlib.sha256(np.ascontiguousarray(r["u"], dtype="<f8").tobytes()).hexdigest() == g["u_sha256"]


@pytest.mark.parametrize("ranks", [2, 3])
def test_strips_emulated_kernels_vs_oracle(ranks):
    opts = base(129, 4) + " " + JAC
    o = Oracle(opts)
    e = mgb.Engine(4, 127, nranks=ranks, emulate=True, agglomerate_below=31)
    e.set_poisson_uniform()
    rng = np.random.default_rng(11)
    try:
        for l in range(4):
            ni, nj = e.dims(l)
            x = rng.uniform(-1, 1, (ni, nj))
            b = rng.uniform(-1, 1, (ni, nj))
            e.set_vec(mgb.VEC_U, l, x)
            e.set_vec(mgb.VEC_B, l, b)
            e.residual(l)
            r_o = o.residual(l, b.reshape(-1), x.reshape(-1)).reshape(ni, nj)
            assert e.get_vec(mgb.VEC_R, l).tobytes() == r_o.tobytes()
            assert e.residual_norm(l) == pytest.approx(o.norm2(r_o), rel=1e-13)
            assert e.dot(mgb.VEC_B, mgb.VEC_U, l) == pytest.approx(o.dot(b, x), rel=1e-10, abs=1e-12)
            for nu, gz in ((3, False), (2, True)):
                e.set_vec(mgb.VEC_U, l, x)
                e.smooth(l, mgb.jacobi(0.8), nu, gz)
                x_o = o.smooth(l, b.reshape(-1), x.reshape(-1), nu, gz).reshape(ni, nj)
                assert e.get_vec(mgb.VEC_U, l).tobytes() == x_o.tobytes(), (l, nu, gz)
            if l + 1 < 4:
                nci, ncj = e.dims(l + 1)
                e.set_vec(mgb.VEC_U, l, x)
                bc_o = o.matmult(1, l, r_o.reshape(-1)).reshape(nci, ncj)
                e.restrict(l, fused=True)
                assert e.get_vec(mgb.VEC_B, l + 1).tobytes() == bc_o.tobytes()
                uc = rng.uniform(-1, 1, (nci, ncj))
                e.set_vec(mgb.VEC_U, l + 1, uc)
                e.set_vec(mgb.VEC_U, l, x)
                e.prolong(l, multadd=False)
                want = x.reshape(-1) + 1.0 * o.matmult(2, l, uc.reshape(-1))
                assert e.get_vec(mgb.VEC_U, l).tobytes() == want.reshape(ni, nj).tobytes()
    finally:
        e.close()
        o.close()


def test_strip_configuration_errors():
    with pytest.raises(mgb.MgbError):
        mgb.Engine(4, 127, nranks=9, emulate=True)                      # more than one NVSwitch domain
    with pytest.raises(mgb.MgbError):
        mgb.Engine(4, 127, nranks=8, emulate=True, agglomerate_below=31)  # strips too thin
    with pytest.raises(mgb.MgbError):
        mgb.Engine(4, 127, nranks=2, emulate=True, agglomerate_below=127) # nothing to distribute


# ------------------------------------------------------------------ fused legs (temporal blocking) == one kernel per sweep
@pytest.mark.parametrize("opts", [
    base(129, 4) + " " + JAC, base(129, 7, v="2,1") + " " + JAC, base(129, 7, v="1,4") + " " + JAC,
    base(129, 5, v="4,2") + " " + JAC, base(129, 4, v="6,9", it=200) + " " + JAC, base(101, 3) + " " + JAC,
    base(65, 4, mesh=1, it=400) + " " + JAC, base(65, 4, mesh=2, it=400) + " " + JAC, base(17, 1, it=50) + " " + JAC,
    base(1025, 10) + " " + JAC, base(513, 3, it=30) + " " + JAC,
    base(2049, 3, it=6, mp=3) + " -pc_type sor",                       # red-black: level 0 fused (3 passes per leg), levels 1-2 one-sweep kernels
    base(2049, 4, it=5, mp=3) + " -pc_type sor -pc_sor_forward -pc_sor_omega 1.1",
    base(129, 7, cycle=8) + " -ksp_type cg -ksp_rtol 1e-10 -mg_levels_ksp_type richardson -mg_levels_pc_type jacobi "
    "-mg_levels_ksp_richardson_scale 0.8 -mg_levels_ksp_max_it 3",
    base(257, 4, cycle=8, it=60) + " -ksp_type richardson -mg_levels_ksp_type richardson -mg_levels_pc_type jacobi "
    "-mg_levels_ksp_richardson_scale 0.7 -mg_levels_ksp_max_it 2 -mg_coarse_ksp_type richardson -mg_coarse_pc_type jacobi "
    "-mg_coarse_ksp_richardson_scale 0.8 -mg_coarse_ksp_max_it 7"])
@pytest.mark.parametrize("ranks", [1, 2])
def test_fused_legs_bit_identical_to_unfused(opts, ranks):
    extra = "" if ranks == 1 else f" -mgb_ranks {ranks} -mgb_emulate 1 -mgb_agglomerate 31"
    if ranks > 1 and ("-npts 17 " in opts):
        pytest.skip("nothing to distribute")
    a = mgb.run_poisson(opts + extra + " -mgb_fuse 1")
    b = mgb.run_poisson(opts + " -mgb_fuse 0")
    assert a["num_iter"] == b["num_iter"]
    if "-ksp_type cg" in opts:
        # the fused last leg of the preconditioner accumulates z'r in passing (POST_DOT), the one-sweep path in a separate
        # reduction, and on strips the dot products are summed strip by strip: alpha / beta differ in the last bits
        assert np.abs(a["u"] - b["u"]).max() <= RTOL * np.abs(b["u"]).max()
        assert np.allclose(a["rnorm"], b["rnorm"], rtol=RTOL, atol=RNORM_ATOL, equal_nan=True)
    else:
        assert a["u"].tobytes() == b["u"].tobytes()
        assert np.allclose(a["rnorm"], b["rnorm"], rtol=1e-12, atol=RNORM_ATOL, equal_nan=True)


# ------------------------------------------------------------------ pipelined stream of right-hand sides
@pytest.mark.parametrize("extra", ["", " -mgb_ranks 2 -mgb_emulate 1 -mgb_agglomerate 31"])
def test_solve_rhs_many_matches_single_solves(extra):
    """mgb_solve_vcycle_many overlaps the copies of neighbouring solves with the running one; every solution must be
    bit-identical to the one-at-a-time Solve() on the same right-hand side."""
    opts = base(257, 8, it=100) + " " + JAC + " -mgb_csr 0" + extra
    n = 255
    rng = np.random.default_rng(5)
    bs = [np.ascontiguousarray(rng.uniform(-1, 1, (n, n))) for _ in range(4)]
    s = mgb.Session(opts)
    try:
        singles = []
        for b in bs:
            u = np.zeros((n, n))
            r = s.solve_rhs(b.ctypes.data, u.ctypes.data)
            singles.append((r["num_iter"], r["rnorm"][-1], u))
        us = [np.zeros((n, n)) for _ in bs]
        its, fin, sec = s.solve_rhs_many([b.ctypes.data for b in bs], [u.ctypes.data for u in us])
        for k in range(len(bs)):
            assert its[k] == singles[k][0]
            assert fin[k] == pytest.approx(singles[k][1], rel=1e-12)
            assert us[k].tobytes() == singles[k][2].tobytes()
        # and once more on the same session (graph cache, swapped spare vectors)
        its2, _, _ = s.solve_rhs_many([b.ctypes.data for b in bs[:3]], [u.ctypes.data for u in us[:3]])
        assert its2 == its[:3]
        assert us[2].tobytes() == singles[2][2].tobytes()
    finally:
        s.close()


# ------------------------------------------------------------------ rectangular grids (the weak-scaling workload)
@pytest.mark.parametrize("ni,nj,levels", [(383, 127, 6), (127, 255, 5), (767, 95, 4)])
def test_rectangular_grid_paths_agree(ni, nj, levels):
    """The reference forces square grids (src/poisson.c:73-75); bench.py --workload weak uses ni != nj through the
    C-ABI.  No oracle there, so the three code paths are checked against each other bit for bit: one kernel per
    sweep, fused legs + bottom kernel, and fused legs on 2 / 3 emulated strips."""
    rng = np.random.default_rng(ni + nj)
    b = rng.uniform(-1, 1, (ni, nj))
    outs = []
    for kw, solve_kw in ((dict(), dict(fuse=False)), (dict(), dict(fuse=True)), (dict(), dict(fuse=True, bottom=False)),
                         (dict(nranks=2, emulate=True, agglomerate_below=63), dict(fuse=True)),
                         (dict(nranks=3, emulate=True, agglomerate_below=63), dict(fuse=False))):
        e = mgb.Engine(levels, ni, nj, **kw)
        try:
            e.set_poisson_uniform()
            e.set_vec(mgb.VEC_B, 0, b)
            it, rn, _ = e.solve_vcycle(mgb.jacobi(0.8), 3, 3, max_iter=12, rtol=1e-7, **solve_kw)
            outs.append((it, rn, e.get_vec(mgb.VEC_U, 0)))
        finally:
            e.close()
    for it, rn, u in outs[1:]:
        assert it == outs[0][0]
        assert u.tobytes() == outs[0][2].tobytes()
        assert np.allclose(rn, outs[0][1], rtol=1e-12, atol=RNORM_ATOL)
    assert outs[0][1][-1] < outs[0][1][0]


# ------------------------------------------------------------------ CSR assembly on row strips
@pytest.mark.parametrize("ranks,aggl", [(2, 31), (3, 31), (4, 31), (2, 15)])
def test_csr_on_strips_concatenates_to_the_reference_matrix(ranks, aggl):
    """Every rank assembles the rows it owns (local row pointers, global columns): stacked in rank order they must be
    the oracle's matrix, bit for bit -- for A, res and pro on every level."""
    opts = base(129, 5) + " " + JAC
    o = Oracle(opts)
    e = mgb.Engine(5, 127, nranks=ranks, emulate=True, agglomerate_below=aggl)
    try:
        e.set_poisson_uniform()
        e.assemble_csr()
        for l in range(5):
            for which in (mgb.MAT_A, mgb.MAT_RES, mgb.MAT_PRO):
                if which != mgb.MAT_A and l == 4:
                    continue
                shape_o, ia_o, ja_o, va_o = o.csr(which, l)
                dist = mgb.strip_rows(5, 127, 127, ranks, l, 0, aggl)[2]
                ia, ja, va, rows = [np.zeros(1, dtype=np.int64)], [], [], 0
                for r in (range(ranks) if dist else [0]):
                    (m, n), ia_r, ja_r, va_r, row0 = e.csr(which, l, rank=r, with_row0=True)
                    assert n == shape_o[1] and row0 == rows
                    ia.append(ia_r[1:].astype(np.int64) + ia[-1][-1])
                    ja.append(ja_r); va.append(va_r); rows += m
                assert rows == shape_o[0]
                assert np.array_equal(np.concatenate(ia), ia_o)
                assert np.array_equal(np.concatenate(ja), ja_o)
                assert np.concatenate(va).tobytes() == va_o.tobytes()
    finally:
        e.close()
        o.close()


@pytest.mark.parametrize("name", ["n129_l7_jacobi", "n101_l3_jacobi", "n1025_l10_jacobi", "n129_l7_cg_mg",
                                  "n129_l7_rbsor", "n129_l7_rbsor_w12", "n1025_l7_rbsor"])
def test_without_bottom_kernel_matches_golden(name):
    """-mgb_bottom 0: the smallest levels go through the fused legs (down to 1 x 1) instead of the cluster kernel."""
    g = GOLD[name]
    r = mgb.run_poisson(g["options"] + " -mgb_bottom 0")
    assert r["num_iter"] == g["num_iter"]
    want = _hex(g["rnorm_hex"])
    ok = ~np.isnan(want)
    assert np.allclose(r["rnorm"][ok], want[ok], rtol=RTOL, atol=RNORM_ATOL)
    if "-cycle 0" in g["options"]:
        assert hashlib.sha256(np.ascontiguousarray(r["u"], dtype="<f8").tobytes()).hexdigest() == g["u_sha256"]


@pytest.mark.parametrize("ksp,sweep", [("cg", ""), ("richardson", " -mg_levels_pc_sor_forward -mg_coarse_pc_sor_forward")])
def test_pcmg_red_black_bottom_kernel_bit_identical_to_separate_launches(ksp, sweep):
    """Cycle 8 with red-black SOR on every level and a Richardson+SOR coarse solver: the persistent bottom kernel (levels of
    at most 63 rows, red-black half sweeps in place, MatInterpolateAdd order) gives the bits of the one-launch-per-half-sweep path."""
    opts = (base(129, 7, mp=3).replace("-cycle 0", "-cycle 8") + f" -ksp_type {ksp} -ksp_rtol 1e-9 -mg_levels_ksp_type richardson "
            "-mg_levels_pc_type sor -mg_levels_ksp_max_it 2 -mg_coarse_ksp_type richardson -mg_coarse_pc_type sor -mg_coarse_ksp_max_it 3" + sweep)
    a = mgb.run_poisson(opts)
    b = mgb.run_poisson(opts + " -mgb_bottom 0")
    assert a["num_iter"] == b["num_iter"] and a["num_iter"] > 1
    assert a["gpu_launches"] < b["gpu_launches"]                          # the bottom kernel really replaced launches
    assert np.array_equal(a["u"], b["u"])
    assert np.array_equal(a["rnorm"], b["rnorm"])


@pytest.mark.parametrize("extra", ["", " -mgb_ranks 2 -mgb_emulate 1 -mgb_agglomerate 31", " -mgb_ranks 4 -mgb_emulate 1 -mgb_agglomerate 31"])
@pytest.mark.parametrize("name", ["n129_l7_cg_mg", "n129_l4_cg_mg_jcoarse"])
def test_cg_fused_direction_step_bit_identical_to_separate_passes(name, extra, monkeypatch):
    """CG: x += a p (deferred), p = z + b p, w = A p and p'w in one pass (k_cg_pstep; on strips the ghost rows of p are derived
    locally instead of exchanged) against the separate VecAYPX / MatMult+VecDot / VecAXPY passes: same bits, same history."""
    g = GOLD[name]
    a = mgb.run_poisson(g["options"] + extra)
    monkeypatch.setenv("MGB_CG_FUSE", "0")
    b = mgb.run_poisson(g["options"] + extra)
    assert a["num_iter"] == b["num_iter"] == g["num_iter"]
    assert a["gpu_launches"] < b["gpu_launches"]
    assert np.array_equal(a["rnorm"], b["rnorm"], equal_nan=True)
    assert np.array_equal(a["u"], b["u"])

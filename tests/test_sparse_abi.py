"""include/mgb200_sparse.h called directly (ctypes) on GENERAL matrices -- not the 5-point stencil: random structurally
symmetric sparsity patterns, so that the level-set schedule of the sequential sweeps (MatSOR, ILU(0), triangular solves) is
exercised on dependency graphs the Poisson runs never produce.  The checker is a plain sequential restatement in Python
(small cases only) of PETSc's loops as oracle/minipetsc restates them [PETSc-upstream]; float64 Python arithmetic is IEEE
round-to-nearest without FMA, so the bar is BIT equality."""
import ctypes as C
import importlib
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
mgb = importlib.import_module("multigrid-petsc_b200")
RED_BLK = 4096


@pytest.fixture(scope="module")
def L():
    lib = mgb.engine_lib()
    vp, ci, cd = C.c_void_p, C.c_int, C.c_double
    pd, pi = C.POINTER(C.c_double), C.POINTER(C.c_int)
    sig = {
        "mgb_dvec_create": [ci, C.POINTER(vp)], "mgb_dvec_destroy": [vp], "mgb_dvec_upload": [vp, pd], "mgb_dvec_download": [vp, pd],
        "mgb_dvec_set": [vp, cd], "mgb_dvec_copy": [vp, vp], "mgb_dvec_axpy": [vp, cd, vp], "mgb_dvec_aypx": [vp, cd, vp],
        "mgb_dvec_waxpy": [vp, cd, vp, vp], "mgb_dvec_axpbypcz": [vp, cd, cd, cd, vp, vp], "mgb_dvec_pointwise_mult": [vp, vp, vp],
        "mgb_dvec_scale": [vp, cd], "mgb_dvec_dot": [vp, vp, pd], "mgb_dvec_norm": [vp, ci, pd],
        "mgb_dindex_create": [ci, pi, C.POINTER(vp)], "mgb_dindex_destroy": [vp], "mgb_dvec_gather": [vp, vp, vp], "mgb_dvec_scatter": [vp, vp, vp],
        "mgb_dcsr_create": [ci, ci, pi, pi, pd, C.POINTER(vp)], "mgb_dcsr_destroy": [vp], "mgb_dcsr_mult": [vp, vp, vp],
        "mgb_dcsr_mult_add": [vp, vp, vp, vp], "mgb_dcsr_scale": [vp, cd], "mgb_dcsr_inverse_diagonal": [vp, vp],
        "mgb_dcsr_sor": [vp, vp, cd, ci, cd, ci, ci, vp], "mgb_dcsr_ilu0_factor": [vp], "mgb_dcsr_ilu0_solve": [vp, vp, vp],
        "mgb_dcsr_lu_factor": [vp], "mgb_dcsr_lu_solve": [vp, vp, vp],
    }
    for name, args in sig.items():
        getattr(lib, name).argtypes = args
        getattr(lib, name).restype = ci
    lib.mgb_last_error.restype = C.c_char_p
    return lib


def ok(L, rc):
    assert rc == 0, L.mgb_last_error().decode()


class V:
    def __init__(self, L, a):
        self.L, self.n = L, len(a)
        self.h = C.c_void_p()
        ok(L, L.mgb_dvec_create(self.n, C.byref(self.h)))
        a = np.ascontiguousarray(a, dtype=np.float64)
        ok(L, L.mgb_dvec_upload(self.h, a.ctypes.data_as(C.POINTER(C.c_double))))

    def get(self):
        out = np.empty(self.n)
        ok(self.L, self.L.mgb_dvec_download(self.h, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def __del__(self):
        try:
            self.L.mgb_dvec_destroy(self.h)
        except Exception:
            pass


def random_symmetric_pattern(n, extra, rng, dominant=True):
    """CSR of a matrix with a full diagonal and `extra` random off-diagonal pairs (i,j),(j,i) with unrelated values"""
    cols = [{i} for i in range(n)]
    for _ in range(extra):
        i, j = rng.integers(0, n, 2)
        if i != j:
            cols[i].add(int(j)); cols[j].add(int(i))
    for i in range(n - 1):                                  # a chain, so that long dependency paths exist
        if rng.uniform() < 0.7:
            cols[i].add(i + 1); cols[i + 1].add(i)
    ia, ja, va = [0], [], []
    for i in range(n):
        row = sorted(cols[i])
        vals = rng.uniform(-1.0, 1.0, len(row))
        if dominant:
            vals[row.index(i)] = 1.0 + np.abs(vals).sum()
        ja += row; va += list(vals); ia.append(len(ja))
    return np.array(ia, dtype=np.int32), np.array(ja, dtype=np.int32), np.array(va, dtype=np.float64)


def make_csr(L, n, ia, ja, va):
    h = C.c_void_p()
    ok(L, L.mgb_dcsr_create(n, n, ia.ctypes.data_as(C.POINTER(C.c_int)), ja.ctypes.data_as(C.POINTER(C.c_int)),
                            va.ctypes.data_as(C.POINTER(C.c_double)), C.byref(h)))
    return h


def diag_pos(n, ia, ja):
    return [next(k for k in range(ia[i], ia[i + 1]) if ja[k] == i) for i in range(n)]


def py_sor(n, ia, ja, va, b, x, omega, flag, its):
    """MatSOR_SeqAIJ, sequential (the loop bodies of the restatement in oracle/minipetsc: same operations, same order)"""
    d = diag_pos(n, ia, ja)
    idiag = [(1.0 / va[d[i]]) if omega == 1.0 else omega / (0.0 + va[d[i]]) for i in range(n)]
    x = [float(v) for v in x]; t = [0.0] * n
    fwd, bwd = bool(flag & 1 or flag & 4), bool(flag & 2 or flag & 8)
    if flag & 16:
        if fwd:
            for i in range(n):
                s = float(b[i])
                for k in range(ia[i], d[i]):
                    s -= va[k] * x[ja[k]]
                t[i] = s; x[i] = s * idiag[i]
        if bwd:
            for i in range(n - 1, -1, -1):
                s = t[i] if fwd else float(b[i])
                for k in range(d[i] + 1, ia[i + 1]):
                    s -= va[k] * x[ja[k]]
                x[i] = ((1 - omega) * x[i] + s * idiag[i]) if fwd else s * idiag[i]
        its -= 1
    for _ in range(its):
        if fwd:
            for i in range(n):
                s = float(b[i])
                for k in range(ia[i], d[i]):
                    s -= va[k] * x[ja[k]]
                t[i] = s
                for k in range(d[i] + 1, ia[i + 1]):
                    s -= va[k] * x[ja[k]]
                x[i] = (1. - omega) * x[i] + s * idiag[i]
        if bwd:
            for i in range(n - 1, -1, -1):
                if fwd:
                    s = t[i]
                    for k in range(d[i] + 1, ia[i + 1]):
                        s -= va[k] * x[ja[k]]
                    x[i] = (1. - omega) * x[i] + s * idiag[i]
                else:
                    s = float(b[i])
                    for k in range(ia[i], ia[i + 1]):
                        s -= va[k] * x[ja[k]]
                    x[i] = (1. - omega) * x[i] + (s + va[d[i]] * x[i]) * idiag[i]
    return np.array(x)


def py_ilu0_solve(n, ia, ja, va, b):
    d = diag_pos(n, ia, ja)
    fac = [float(v) for v in va]
    for i in range(n):
        pos = {int(ja[k]): k for k in range(ia[i], ia[i + 1])}
        for k in range(ia[i], d[i]):
            row = int(ja[k])
            if fac[k] != 0.0:
                mult = fac[k] * fac[d[row]]
                fac[k] = mult
                for tt in range(d[row] + 1, ia[row + 1]):
                    p = pos.get(int(ja[tt]))
                    if p is not None:
                        fac[p] -= mult * fac[tt]
        fac[d[i]] = 1.0 / fac[d[i]]
    tmp = [0.0] * n
    for i in range(n):
        s = float(b[i])
        for k in range(ia[i], d[i]):
            s -= fac[k] * tmp[ja[k]]
        tmp[i] = s
    x = [0.0] * n
    for i in range(n - 1, -1, -1):
        s = tmp[i]
        for k in range(d[i] + 1, ia[i + 1]):
            s -= fac[k] * tmp[ja[k]]
        x[i] = tmp[i] = s * fac[d[i]]
    return np.array(x)


def py_dense_lu_solve(n, ia, ja, va, b):
    dm = [[0.0] * n for _ in range(n)]
    for i in range(n):
        for k in range(ia[i], ia[i + 1]):
            dm[i][ja[k]] = float(va[k])
    for i in range(n):
        for k in range(i):
            if dm[i][k] != 0.0:
                mult = dm[i][k] * dm[k][k]
                dm[i][k] = mult
                for j in range(k + 1, n):
                    dm[i][j] -= mult * dm[k][j]
        dm[i][i] = 1.0 / dm[i][i]
    tmp = [0.0] * n
    for i in range(n):
        s = float(b[i])
        for k in range(i):
            s -= dm[i][k] * tmp[k]
        tmp[i] = s
    x = [0.0] * n
    for i in range(n - 1, -1, -1):
        s = tmp[i]
        for k in range(i + 1, n):
            s -= dm[i][k] * tmp[k]
        x[i] = tmp[i] = s * dm[i][i]
    return np.array(x)


def blocked_dot(x, y):
    tot = 0.0
    for lo in range(0, len(x), RED_BLK):
        s = 0.0
        for a, b in zip(x[lo:lo + RED_BLK], y[lo:lo + RED_BLK]):
            s += float(a) * float(b)
        tot += s
    return tot


# ------------------------------------------------------------------ vectors
def test_vector_operations_bitwise(L):
    rng = np.random.default_rng(0)
    n = 10007
    x, y, z = rng.standard_normal(n), rng.standard_normal(n), rng.standard_normal(n)
    a, b, g = 0.7, -1.3, 0.25
    vx, vy, vz = V(L, x), V(L, y), V(L, z)
    ok(L, L.mgb_dvec_axpy(vy.h, a, vx.h)); y1 = y + a * x
    assert np.array_equal(vy.get(), y1)
    ok(L, L.mgb_dvec_aypx(vy.h, b, vx.h)); y2 = x + b * y1
    assert np.array_equal(vy.get(), y2)
    ok(L, L.mgb_dvec_axpbypcz(vz.h, a, b, g, vx.h, vy.h)); z1 = g * z + a * x + b * y2
    assert np.array_equal(vz.get(), z1)
    vw = V(L, np.zeros(n))
    ok(L, L.mgb_dvec_waxpy(vw.h, a, vx.h, vy.h)); assert np.array_equal(vw.get(), a * x + y2)
    ok(L, L.mgb_dvec_pointwise_mult(vw.h, vx.h, vy.h)); assert np.array_equal(vw.get(), x * y2)
    ok(L, L.mgb_dvec_pointwise_mult(vw.h, vw.h, vx.h)); assert np.array_equal(vw.get(), (x * y2) * x)    # aliasing allowed
    ok(L, L.mgb_dvec_scale(vw.h, -2.5)); assert np.array_equal(vw.get(), -2.5 * ((x * y2) * x))
    ok(L, L.mgb_dvec_set(vw.h, 3.25)); assert np.array_equal(vw.get(), np.full(n, 3.25))
    ok(L, L.mgb_dvec_copy(vw.h, vx.h)); assert np.array_equal(vw.get(), x)
    out = C.c_double()
    ok(L, L.mgb_dvec_dot(vx.h, vy.h, C.byref(out))); assert out.value == blocked_dot(x, y2)
    ok(L, L.mgb_dvec_norm(vx.h, 1, C.byref(out))); assert out.value == math.sqrt(blocked_dot(x, x))
    ok(L, L.mgb_dvec_norm(vx.h, 3, C.byref(out))); assert out.value == np.abs(x).max()
    ok(L, L.mgb_dvec_norm(vx.h, 0, C.byref(out))); assert abs(out.value - np.abs(x).sum()) <= 1e-12 * np.abs(x).sum()
    idx = rng.permutation(n)[:777].astype(np.int32)
    h = C.c_void_p(); ok(L, L.mgb_dindex_create(len(idx), idx.ctypes.data_as(C.POINTER(C.c_int)), C.byref(h)))
    vs = V(L, np.zeros(len(idx)))
    ok(L, L.mgb_dvec_gather(vs.h, vx.h, h)); assert np.array_equal(vs.get(), x[idx])
    ok(L, L.mgb_dvec_scale(vs.h, 2.0)); ok(L, L.mgb_dvec_scatter(vx.h, vs.h, h))
    x2 = x.copy(); x2[idx] = 2.0 * x[idx]
    assert np.array_equal(vx.get(), x2)
    L.mgb_dindex_destroy(h)
    assert L.mgb_dvec_axpy(vs.h, 1.0, vx.h) != 0 and b"size mismatch" in L.mgb_last_error()


# ------------------------------------------------------------------ general CSR
@pytest.mark.parametrize("n,extra,seed", [(1, 0, 0), (37, 60, 1), (120, 300, 4), (500, 1500, 2), (3000, 4000, 3)])
def test_csr_mult_and_sequential_sweeps_bitwise(L, n, extra, seed):
    rng = np.random.default_rng(seed)
    ia, ja, va = random_symmetric_pattern(n, extra, rng)
    A = make_csr(L, n, ia, ja, va)
    x, b = rng.standard_normal(n), rng.standard_normal(n)
    vx, vb, vy = V(L, x), V(L, b), V(L, np.zeros(n))
    # MatMult / MatMultAdd: row sums in ascending column order, from 0.0 / from y_i
    ok(L, L.mgb_dcsr_mult(A, vx.h, vy.h))
    want = np.array([sum_seq(0.0, va, x, ja, ia[i], ia[i + 1]) for i in range(n)])
    assert np.array_equal(vy.get(), want)
    ok(L, L.mgb_dcsr_mult_add(A, vx.h, vb.h, vy.h))
    assert np.array_equal(vy.get(), np.array([sum_seq(b[i], va, x, ja, ia[i], ia[i + 1]) for i in range(n)]))
    ok(L, L.mgb_dcsr_mult_add(A, vx.h, vb.h, vb.h))                                  # z aliases y (MatInterpolateAdd(P, xc, x, x))
    assert np.array_equal(vb.get(), vy.get())
    vb = V(L, b)
    ok(L, L.mgb_dcsr_inverse_diagonal(A, vy.h))
    d = diag_pos(n, ia, ja)
    assert np.array_equal(vy.get(), np.array([1.0 / va[d[i]] for i in range(n)]))
    # MatSOR: every sweep type, zero and nonzero initial guess, omega 1 and 1.3, several iterations
    for flag, omega, its in [(12 | 16, 1.0, 1), (12, 1.0, 2), (3 | 16, 1.3, 2), (1, 1.3, 1), (1 | 16, 1.0, 3), (2, 0.8, 2), (2 | 16, 1.0, 1), (8, 1.0, 1)]:
        x0 = np.zeros(n) if flag & 16 else x
        vxx = V(L, x0)
        ok(L, L.mgb_dcsr_sor(A, vb.h, omega, flag, 0.0, its, 1, vxx.h))
        assert np.array_equal(vxx.get(), py_sor(n, ia, ja, va, b, x0, omega, flag, its)), (flag, omega, its)
    # ILU(0): factorisation + both triangular solves
    ok(L, L.mgb_dcsr_ilu0_factor(A)); ok(L, L.mgb_dcsr_ilu0_solve(A, vb.h, vy.h))
    assert np.array_equal(vy.get(), py_ilu0_solve(n, ia, ja, va, b))
    if n <= 120:                                                                     # dense LU of a small coarse problem
        ok(L, L.mgb_dcsr_lu_factor(A)); ok(L, L.mgb_dcsr_lu_solve(A, vb.h, vy.h))
        got = vy.get()
        assert np.array_equal(got, py_dense_lu_solve(n, ia, ja, va, b))
        r = b - np.array([sum_seq(0.0, va, got, ja, ia[i], ia[i + 1]) for i in range(n)])
        assert np.abs(r).max() <= 1e-9 * np.abs(b).max()                             # and it really solves the system
    # MatScale reaches the factor caches too
    ok(L, L.mgb_dcsr_scale(A, -1.0)); ok(L, L.mgb_dcsr_mult(A, vx.h, vy.h))
    assert np.array_equal(vy.get(), -want)
    L.mgb_dcsr_destroy(A)


def sum_seq(s0, va, x, ja, k0, k1):
    s = float(s0)
    for k in range(k0, k1):
        s += float(va[k]) * float(x[ja[k]])
    return s


def test_sequential_sweeps_refuse_what_they_cannot_schedule(L):
    # structurally unsymmetric pattern: entry (0,2) without (2,0)
    ia = np.array([0, 2, 3, 4], dtype=np.int32); ja = np.array([0, 2, 1, 2], dtype=np.int32); va = np.array([2.0, 1.0, 2.0, 2.0])
    A = make_csr(L, 3, ia, ja, va)
    vb, vx = V(L, np.ones(3)), V(L, np.zeros(3))
    assert L.mgb_dcsr_sor(A, vb.h, 1.0, 12 | 16, 0.0, 1, 1, vx.h) != 0 and b"structurally symmetric" in L.mgb_last_error()
    L.mgb_dcsr_destroy(A)
    # missing diagonal
    ia = np.array([0, 1, 2], dtype=np.int32); ja = np.array([1, 0], dtype=np.int32); va = np.array([1.0, 1.0])
    A = make_csr(L, 2, ia, ja, va)
    vb, vx = V(L, np.ones(2)), V(L, np.zeros(2))
    assert L.mgb_dcsr_ilu0_factor(A) != 0 and b"missing diagonal" in L.mgb_last_error()
    L.mgb_dcsr_destroy(A)
    # unsorted columns are rejected at creation
    ia = np.array([0, 2], dtype=np.int32); ja = np.array([1, 0], dtype=np.int32); va = np.array([1.0, 1.0])
    h = C.c_void_p()
    assert L.mgb_dcsr_create(1, 2, ia.ctypes.data_as(C.POINTER(C.c_int)), ja.ctypes.data_as(C.POINTER(C.c_int)),
                             va.ctypes.data_as(C.POINTER(C.c_double)), C.byref(h)) != 0 and b"not ascending" in L.mgb_last_error()

"""ctypes binding of oracle/libmgoracle.so (TEST INFRASTRUCTURE ONLY)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def oracle_lib_path():
    return os.path.join(_HERE, "libmgoracle.so")


def ref_binary_path():
    return os.path.join(_HERE, "_ref", "poisson_ref")


def ref_l2_path():
    return os.path.join(_HERE, "_ref", "libref_l2.so")


def build_oracle(with_ref=True):
    """Compile the oracle (and, if /root/reference exists, the reference itself into oracle/_ref/)."""
    subprocess.run(["make", "-s", "-C", _HERE], check=True)
    if with_ref and os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-s", "-C", _HERE, "ref"], check=True)


_lib = None


def _load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(oracle_lib_path()):
        build_oracle(with_ref=False)
    L = C.CDLL(oracle_lib_path())
    vp, ci, cd = C.c_void_p, C.c_int, C.c_double
    pi, pd = C.POINTER(C.c_int), C.POINTER(C.c_double)
    L.mgo_create.restype = vp
    L.mgo_create.argtypes = [C.c_char_p]
    L.mgo_destroy.argtypes = [vp]
    L.mgo_levels.argtypes = [vp]
    L.mgo_level_dims.argtypes = [vp, ci, pi, pi]
    L.mgo_level_h.argtypes = [vp, ci, pd]
    L.mgo_coords.argtypes = [vp, ci, pd]
    L.mgo_stencil.argtypes = [vp, ci, pd]
    L.mgo_opA.argtypes = [vp, ci, ci, ci, pd]
    L.mgo_grid_to_global.argtypes = [vp, ci, pi]
    L.mgo_global_to_grid.argtypes = [vp, ci, pi]
    L.mgo_csr_dims.argtypes = [vp, ci, ci, pi, pi, pi]
    L.mgo_csr_copy.argtypes = [vp, ci, ci, pi, pi, pd]
    L.mgo_vec_get.argtypes = [vp, ci, ci, pd]
    L.mgo_vec_set.argtypes = [vp, ci, ci, pd]
    L.mgo_matmult.argtypes = [vp, ci, ci, pd, pd]
    L.mgo_matmultadd.argtypes = [vp, ci, ci, pd, pd, pd]
    L.mgo_residual.argtypes = [vp, ci, pd, pd, pd]
    L.mgo_smooth.argtypes = [vp, ci, pd, pd, ci, ci]
    L.mgo_norm2.restype = cd
    L.mgo_norm2.argtypes = [pd, ci]
    L.mgo_dot.restype = cd
    L.mgo_dot.argtypes = [pd, pd, ci]
    L.mgo_solve.argtypes = [vp]
    L.mgo_num_iter.argtypes = [vp]
    L.mgo_rnorm.argtypes = [vp, pd, ci]
    L.mgo_solve_seconds.restype = cd
    L.mgo_solve_seconds.argtypes = [vp]
    L.mgo_postprocess.argtypes = [vp, pd, pd]
    L.mgo_write_files.argtypes = [vp, C.c_char_p]
    _lib = L
    return L


def _pd(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _pi(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class Oracle:
    """One assembled problem instance of the CPU oracle.  Vectors are in the oracle's GLOBAL numbering
    (natural row-major unless -map 3); use to_grid()/from_grid() to convert to (ni, nj) arrays."""

    def __init__(self, options):
        self.L = _load()
        self.options = options
        self.h = self.L.mgo_create(options.encode())
        if not self.h:
            raise ValueError("oracle rejected options: " + options)
        self.levels = self.L.mgo_levels(self.h)
        self._g2G = {}

    def close(self):
        if getattr(self, "h", None):
            self.L.mgo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- geometry
    def dims(self, l):
        ni, nj = C.c_int(), C.c_int()
        assert self.L.mgo_level_dims(self.h, l, C.byref(ni), C.byref(nj)) == 0
        return ni.value, nj.value

    def level_h(self, l):
        h = np.zeros(2)
        assert self.L.mgo_level_h(self.h, l, _pd(h)) == 0
        return h

    def coords(self, dim, npts):
        out = np.zeros(npts)
        assert self.L.mgo_coords(self.h, dim, _pd(out)) == 0
        return out

    def stencil(self, which):
        out = np.zeros(9)
        self.L.mgo_stencil(self.h, which, _pd(out))
        return out.reshape(3, 3)

    def opA(self, l, i, j):
        out = np.zeros(5)
        assert self.L.mgo_opA(self.h, l, i, j, _pd(out)) == 0
        return out

    def grid_to_global(self, l):
        if l not in self._g2G:
            ni, nj = self.dims(l)
            out = np.zeros(ni * nj, dtype=np.int32)
            assert self.L.mgo_grid_to_global(self.h, l, _pi(out)) == 0
            self._g2G[l] = out
        return self._g2G[l]

    def global_to_grid(self, l):
        ni, nj = self.dims(l)
        out = np.zeros(ni * nj * 3, dtype=np.int32)
        assert self.L.mgo_global_to_grid(self.h, l, _pi(out)) == 0
        return out.reshape(-1, 3)

    def to_grid(self, l, vec_global):
        ni, nj = self.dims(l)
        return np.ascontiguousarray(np.asarray(vec_global)[self.grid_to_global(l)]).reshape(ni, nj)

    def from_grid(self, l, arr_grid):
        g = self.grid_to_global(l)
        out = np.empty(g.size)
        out[g] = np.asarray(arr_grid, dtype=np.float64).reshape(-1)
        return out

    # --- matrices / vectors
    def csr_dims(self, which, l):
        m, n, nnz = C.c_int(), C.c_int(), C.c_int()
        assert self.L.mgo_csr_dims(self.h, which, l, C.byref(m), C.byref(n), C.byref(nnz)) == 0
        return m.value, n.value, nnz.value

    def csr(self, which, l):
        m, n, nnz = self.csr_dims(which, l)
        ia = np.zeros(m + 1, dtype=np.int32)
        ja = np.zeros(nnz, dtype=np.int32)
        va = np.zeros(nnz)
        assert self.L.mgo_csr_copy(self.h, which, l, _pi(ia), _pi(ja), _pd(va)) == 0
        return (m, n), ia, ja, va

    def vec(self, which, l):
        ni, nj = self.dims(l)
        out = np.zeros(ni * nj)
        assert self.L.mgo_vec_get(self.h, which, l, _pd(out)) == 0
        return out

    def set_vec(self, which, l, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        assert self.L.mgo_vec_set(self.h, which, l, _pd(v)) == 0

    def matmult(self, which, l, x):
        m, n, _ = self.csr_dims(which, l)
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert x.size == n
        y = np.zeros(m)
        assert self.L.mgo_matmult(self.h, which, l, _pd(x), _pd(y)) == 0
        return y

    def matmultadd(self, which, l, x, y):
        m, n, _ = self.csr_dims(which, l)
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        assert x.size == n and y.size == m
        z = np.zeros(m)
        assert self.L.mgo_matmultadd(self.h, which, l, _pd(x), _pd(y), _pd(z)) == 0
        return z

    def residual(self, l, b, x):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.ascontiguousarray(x, dtype=np.float64)
        r = np.zeros_like(b)
        assert self.L.mgo_residual(self.h, l, _pd(b), _pd(x), _pd(r)) == 0
        return r

    def smooth(self, l, b, x, nu, guess_zero):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.array(x, dtype=np.float64, copy=True)
        assert self.L.mgo_smooth(self.h, l, _pd(b), _pd(x), int(nu), int(bool(guess_zero))) == 0
        return x

    def norm2(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        return self.L.mgo_norm2(_pd(x), x.size)

    def dot(self, x, y):
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        y = np.ascontiguousarray(y, dtype=np.float64).reshape(-1)
        return self.L.mgo_dot(_pd(x), _pd(y), x.size)

    # --- solve
    def solve(self):
        assert self.L.mgo_solve(self.h) == 0
        it = self.L.mgo_num_iter(self.h)
        rn = np.zeros(it + 1)
        self.L.mgo_rnorm(self.h, _pd(rn), it + 1)
        return it, rn

    def solve_seconds(self):
        return self.L.mgo_solve_seconds(self.h)

    def postprocess(self):
        ni, nj = self.dims(0)
        u = np.zeros(ni * nj)
        err = np.zeros(3)
        assert self.L.mgo_postprocess(self.h, _pd(u), _pd(err)) == 0
        return u.reshape(ni, nj), err

    def write_files(self, d):
        assert self.L.mgo_write_files(self.h, d.encode()) == 0

"""multigrid-petsc_b200 -- B200-native geometric-multigrid Poisson engine behind the driver surface of
SyamVangara/multigrid-petsc.

Python is plumbing only: this module loads the two in-tree shared libraries and exposes thin ctypes wrappers

  lib/libmgb200.so          the CUDA engine, C-ABI in include/mgb200.h           -> class Engine
  lib/libpoisson_b200.so    the host C layer mirroring the reference's solver.h  -> run_poisson()

There is no CPU fallback: loading fails loudly when the libraries are not built, and every engine call raises
MgbError when no CUDA device is usable.  (The package directory name contains a hyphen, as the task layout
prescribes: import it with importlib.import_module("multigrid-petsc_b200") or via the helper in tests/.)
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(_HERE, "lib")
ENGINE_LIB = os.path.join(LIB_DIR, "libmgb200.so")
HOST_LIB = os.path.join(LIB_DIR, "libpoisson_b200.so")
DRIVER_BIN = os.path.join(LIB_DIR, "poisson_b200")

VEC_B, VEC_U, VEC_R, VEC_W, VEC_P, VEC_Z, VEC_Q = range(7)
MAT_A, MAT_RES, MAT_PRO = range(3)
SMOOTH_JACOBI, SMOOTH_RBSOR, SMOOTH_LEXSOR, SMOOTH_ILU0 = 0, 1, 2, 3
SOR_SYMMETRIC, SOR_FORWARD, SOR_BACKWARD = 0, 1, 2
KSP_RICHARDSON, KSP_CG = 0, 1
IPC_HANDLE_BYTES = 64
COARSE_LU, COARSE_RICHARDSON = 0, 1
# mgb_time_op codes and their algorithmic HBM bytes per (fine) unknown (SURVEY.md 8d / DESIGN.md)
OPS = {"apply": (0, 16), "residual": (1, 24), "jacobi": (2, 24), "rbsor_full": (3, 48), "residual_restrict": (4, 18),
       "prolong_correct": (5, 18), "residual_norm": (6, 16), "csr_spmv": (7, 80), "nrm2": (8, 8), "dot": (9, 16),
       "axpy": (10, 24),
       # fused legs: bytes of the UNFUSED sequence they replace (3 x 24 + 18, 18 + 3 x 24 + 16, 3 x 24, 24, 16 + 2 x 24 + 18)
       # and, second number, their own compulsory traffic (read u, b; write u; +2 coarse)
       "fused_down": (11, 90), "fused_up": (12, 106), "fused_3sweeps": (13, 72), "fused_1sweep": (14, 24),
       "fused_down_zero": (15, 82), "bottom_cycle": (16, 176),
       "halo_exchange": (17, 0), "nrm2_allreduce": (18, 8)}
FUSED_OWN_BYTES = {"fused_down": 26, "fused_up": 26, "fused_3sweeps": 24, "fused_1sweep": 24, "fused_down_zero": 18}


class MgbError(RuntimeError):
    pass


def strip_rows(levels, ni, nj, nranks, level, rank, agglomerate_below=0):
    """The row partition of the engine (host arithmetic, works without a GPU): (row0, row1, distributed)."""
    L = engine_lib()
    cfg = Config(levels, ni, nj, -1, 0, 0, nranks, agglomerate_below, 0)
    a, b, d = C.c_int(), C.c_int(), C.c_int()
    rc = L.mgb_strip_rows(C.byref(cfg), level, rank, C.byref(a), C.byref(b), C.byref(d))
    if rc != 0:
        raise MgbError(f"mgb error {rc}: {L.mgb_last_error().decode()}")
    return a.value, b.value, bool(d.value)


class Smoother(C.Structure):
    _fields_ = [("type", C.c_int), ("scale", C.c_double), ("omega", C.c_double), ("sor_sweep", C.c_int),
                ("sor_its", C.c_int)]


class Config(C.Structure):
    _fields_ = [("levels", C.c_int), ("ni", C.c_int), ("nj", C.c_int), ("device", C.c_int),
                ("red_black_numbering", C.c_int), ("rank", C.c_int), ("nranks", C.c_int),
                ("agglomerate_below", C.c_int), ("emulate", C.c_int)]


class VcycleParams(C.Structure):
    _fields_ = [("smoother", Smoother), ("v0", C.c_int), ("v1", C.c_int), ("max_iter", C.c_int),
                ("rtol", C.c_double), ("use_graph", C.c_int), ("no_fuse", C.c_int), ("no_bottom", C.c_int)]


class PcmgParams(C.Structure):
    _fields_ = [("outer", C.c_int), ("rtol", C.c_double), ("abstol", C.c_double), ("dtol", C.c_double),
                ("max_iter", C.c_int), ("level_smoother", Smoother), ("level_its", C.c_int), ("coarse", C.c_int),
                ("coarse_smoother", Smoother), ("coarse_its", C.c_int), ("no_fuse", C.c_int), ("no_bottom", C.c_int),
                ("no_graph", C.c_int)]


class RunResult(C.Structure):
    _fields_ = [("num_iter", C.c_int), ("ni", C.c_int), ("nj", C.c_int), ("error", C.c_double * 3),
                ("solve_seconds", C.c_double), ("levels", C.c_int), ("gpu_launches", C.c_longlong)]


def build(verbose=False):
    """Compile both libraries in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.run(["make", "-s", "-C", os.path.join(_HERE, "csrc")], check=True, stdout=out)
    subprocess.run(["make", "-s", "-C", os.path.join(_HERE, "host")], check=True, stdout=out)


_eng = None
_host = None


def engine_lib():
    global _eng
    if _eng is None:
        if not os.path.exists(ENGINE_LIB):
            raise MgbError(f"{ENGINE_LIB} is not built (run python -c 'import __graft_entry__ as g; g.build()'); "
                           "the B200 engine has no CPU fallback")
        L = C.CDLL(ENGINE_LIB, mode=C.RTLD_GLOBAL)
        L.mgb_last_error.restype = C.c_char_p
        L.mgb_launch_count.restype = C.c_longlong
        L.mgb_launch_count.argtypes = [C.c_void_p]
        L.mgb_last_solve_ms.restype = C.c_double
        L.mgb_last_solve_ms.argtypes = [C.c_void_p]
        _eng = L
    return _eng


def host_lib():
    global _host
    if _host is None:
        engine_lib()
        if not os.path.exists(HOST_LIB):
            raise MgbError(f"{HOST_LIB} is not built")
        L = C.CDLL(HOST_LIB)
        L.pb200_last_error.restype = C.c_char_p
        L.pb200_run.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(RunResult), C.c_void_p, C.c_void_p, C.c_int]
        L.pb200_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.pb200_solve.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(RunResult), C.c_void_p, C.c_void_p, C.c_int]
        L.pb200_solve_rhs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(RunResult), C.c_void_p, C.c_int]
        L.pb200_solve_rhs_many.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.pb200_close.argtypes = [C.c_void_p]
        L.pb200_session_engine.restype = C.c_void_p
        L.pb200_session_engine.argtypes = [C.c_void_p]
        _host = L
    return _host


def _pd(a):
    return a.ctypes.data_as(C.c_void_p)


def jacobi(scale=1.0):
    return Smoother(SMOOTH_JACOBI, scale, 1.0, SOR_SYMMETRIC, 1)


def rbsor(omega=1.0, sweep=SOR_SYMMETRIC, its=1):
    return Smoother(SMOOTH_RBSOR, 1.0, omega, sweep, its)


def lexsor(omega=1.0, sweep=SOR_SYMMETRIC, its=1):
    """-pc_type sor on the natural numbering: PETSc's lexicographic MatSOR (wavefront kernels, one GPU)"""
    return Smoother(SMOOTH_LEXSOR, 1.0, omega, sweep, its)


def ilu0(scale=1.0):
    """PETSc's default PC: ILU(0) inside Richardson (wavefront kernels, one GPU)"""
    return Smoother(SMOOTH_ILU0, scale, 1.0, SOR_SYMMETRIC, 1)


class Engine:
    """One engine instance (one GPU).  Thin wrapper: every method is one C-ABI call of include/mgb200.h."""

    def __init__(self, levels, ni, nj=None, device=-1, red_black_numbering=False, rank=0, nranks=1,
                 agglomerate_below=0, emulate=False, _borrow=None):
        self.L = engine_lib()
        self.levels = levels
        self._owned = _borrow is None
        if _borrow is not None:            # handle owned by a host-layer session (Session.engine)
            self.h = C.c_void_p(_borrow)
            return
        nj = ni if nj is None else nj
        cfg = Config(levels, ni, nj, device, int(red_black_numbering), rank, nranks, agglomerate_below, int(emulate))
        self.h = C.c_void_p()
        self._ck(self.L.mgb_create(C.byref(cfg), C.byref(self.h)))

    def _ck(self, rc):
        if rc != 0:
            raise MgbError(f"mgb error {rc}: {self.L.mgb_last_error().decode()}")

    def close(self):
        if getattr(self, "h", None):
            if self._owned:
                self.L.mgb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # row strips: IPC handle exchange is the caller's job (multigrid-petsc_b200/strips.py does it with torch.distributed)
    def ipc_export(self):
        buf = C.create_string_buffer(IPC_HANDLE_BYTES)
        self._ck(self.L.mgb_ipc_export(self.h, buf))
        return buf.raw

    def ipc_connect(self, handles):
        """handles: the concatenation of every rank's ipc_export(), in rank order"""
        self._ck(self.L.mgb_ipc_connect(self.h, C.c_char_p(bytes(handles))))

    def local_rows(self, l):
        a, b = C.c_int(), C.c_int()
        self._ck(self.L.mgb_local_rows(self.h, l, C.byref(a), C.byref(b)))
        return a.value, b.value

    def dims(self, l):
        ni, nj = C.c_int(), C.c_int()
        self._ck(self.L.mgb_level_dims(self.h, l, C.byref(ni), C.byref(nj)))
        return ni.value, nj.value

    # operator definition
    def set_level_operator(self, l, rows):
        rows = np.ascontiguousarray(rows, dtype=np.float64)
        assert rows.shape == (self.dims(l)[0], 5)
        self._ck(self.L.mgb_set_level_operator(self.h, l, _pd(rows)))

    def set_transfer(self, res3, pro3):
        r = np.ascontiguousarray(res3, dtype=np.float64).reshape(9)
        p = np.ascontiguousarray(pro3, dtype=np.float64).reshape(9)
        self._ck(self.L.mgb_set_transfer(self.h, _pd(r), _pd(p)))

    def set_poisson_uniform(self):
        """-mesh 0 operator and the reference's transfer stencils (src/problem.c:15-21, src/matbuild.c:398-431),
        computed with the reference's operation order in numpy float64."""
        for l in range(self.levels):
            ni, nj = self.dims(l)
            h0, h1 = np.float64(1.0) / np.float64(ni + 1), np.float64(1.0) / np.float64(nj + 1)
            hx2, hy2 = h0 * h0, h1 * h1
            one, zero = np.float64(1.0), np.float64(0.0)
            row = np.array([one / hy2 - zero / (2 * h1), one / hx2 - zero / (2 * h0), -2.0 * (one / hx2 + one / hy2),
                            one / hx2 + zero / (2 * h0), one / hy2 + zero / (2 * h1)])
            self.set_level_operator(l, np.tile(row, (ni, 1)))
        a = np.abs(1 - np.arange(3)).astype(np.float64)
        pro = np.stack([0.5 - 0.25 * a, 1.0 - 0.5 * a, 0.5 - 0.25 * a], axis=1)
        res = np.stack([0.125 - 0.0625 * a, 0.25 - 0.125 * a, 0.125 - 0.0625 * a], axis=1)
        if self.levels > 1:
            self.set_transfer(res, pro)

    # assembled operators
    def assemble_csr(self):
        self._ck(self.L.mgb_assemble_csr(self.h))

    def csr(self, which, l, rank=-1, with_row0=False):
        """(shape, rowptr, col, val) of the rows rank `rank` holds (-1: first local strip; the whole matrix on one rank)"""
        m, n, nnz, r0 = C.c_int(), C.c_int(), C.c_longlong(), C.c_int()
        self._ck(self.L.mgb_csr_dims_rank(self.h, rank, which, l, C.byref(m), C.byref(n), C.byref(nnz), C.byref(r0)))
        ia = np.zeros(m.value + 1, dtype=np.int32)
        ja = np.zeros(nnz.value, dtype=np.int32)
        va = np.zeros(nnz.value)
        self._ck(self.L.mgb_csr_get_rank(self.h, rank, which, l, _pd(ia), _pd(ja), _pd(va)))
        if with_row0:
            return (m.value, n.value), ia, ja, va, r0.value
        return (m.value, n.value), ia, ja, va

    def csr_spmv(self, which, l, x):
        m, n, nnz = C.c_int(), C.c_int(), C.c_longlong()
        self._ck(self.L.mgb_csr_dims(self.h, which, l, C.byref(m), C.byref(n), C.byref(nnz)))
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        assert x.size == n.value
        y = np.zeros(m.value)
        self._ck(self.L.mgb_csr_spmv(self.h, which, l, _pd(x), _pd(y)))
        return y

    # vectors
    def set_vec(self, which, l, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.size == self.dims(l)[0] * self.dims(l)[1]
        self._ck(self.L.mgb_vec_set(self.h, which, l, _pd(a)))

    def get_vec(self, which, l):
        ni, nj = self.dims(l)
        out = np.zeros((ni, nj))
        self._ck(self.L.mgb_vec_get(self.h, which, l, _pd(out)))
        return out

    def zero_vec(self, which, l):
        self._ck(self.L.mgb_vec_zero(self.h, which, l))

    def set_rhs_separable(self, gx, gy):
        gx = np.ascontiguousarray(gx, dtype=np.float64)
        gy = np.ascontiguousarray(gy, dtype=np.float64)
        self._ck(self.L.mgb_set_rhs_separable(self.h, _pd(gx), _pd(gy)))

    def error_norms_separable(self, sx, sy):
        sx = np.ascontiguousarray(sx, dtype=np.float64)
        sy = np.ascontiguousarray(sy, dtype=np.float64)
        err = np.zeros(3)
        self._ck(self.L.mgb_error_norms_separable(self.h, _pd(sx), _pd(sy), _pd(err)))
        return err

    # single operations
    def apply(self, l, xv, yv):
        self._ck(self.L.mgb_op_apply(self.h, l, xv, yv))

    def residual(self, l):
        self._ck(self.L.mgb_op_residual(self.h, l))

    def residual_norm(self, l):
        v = C.c_double()
        self._ck(self.L.mgb_op_residual_norm(self.h, l, C.byref(v)))
        return v.value

    def smooth(self, l, smoother, its, guess_zero):
        self._ck(self.L.mgb_op_smooth(self.h, l, C.byref(smoother), int(its), int(bool(guess_zero))))

    def restrict(self, l, fused=True):
        self._ck(self.L.mgb_op_restrict(self.h, l, int(fused)))

    def prolong(self, l, multadd=False):
        self._ck(self.L.mgb_op_prolong(self.h, l, int(multadd)))

    def norm2(self, which, l):
        v = C.c_double()
        self._ck(self.L.mgb_op_norm2(self.h, which, l, C.byref(v)))
        return v.value

    def dot(self, xw, yw, l):
        v = C.c_double()
        self._ck(self.L.mgb_op_dot(self.h, xw, yw, l, C.byref(v)))
        return v.value

    def axpy(self, yw, alpha, xw, l):
        self._ck(self.L.mgb_op_axpy(self.h, yw, C.c_double(alpha), xw, l))

    def aypx(self, yw, beta, xw, l):
        self._ck(self.L.mgb_op_aypx(self.h, yw, C.c_double(beta), xw, l))

    # solvers
    def solve_vcycle(self, smoother, v0=3, v1=3, max_iter=100, rtol=1e-7, use_graph=True, fuse=True, bottom=True):
        p = VcycleParams(smoother, v0, v1, max_iter, rtol, int(use_graph), int(not fuse), int(not bottom))
        rn = np.zeros(max_iter + 1)
        it, sec = C.c_int(), C.c_double()
        self._ck(self.L.mgb_solve_vcycle(self.h, C.byref(p), _pd(rn), C.byref(it), C.byref(sec)))
        return it.value, rn[: it.value + 1].copy(), sec.value

    def solve_pcmg(self, outer, level_smoother, level_its, coarse=COARSE_LU, coarse_smoother=None, coarse_its=1,
                   rtol=1e-7, abstol=1e-50, dtol=1e4, max_iter=100, fuse=True, bottom=True, graph=True):
        cs = coarse_smoother if coarse_smoother is not None else jacobi(1.0)
        p = PcmgParams(outer, rtol, abstol, dtol, max_iter, level_smoother, level_its, coarse, cs, coarse_its, int(not fuse), int(not bottom),
                       int(not graph))
        rn = np.zeros(max_iter + 1)
        it, reason, sec = C.c_int(), C.c_int(), C.c_double()
        self._ck(self.L.mgb_solve_pcmg(self.h, C.byref(p), _pd(rn), C.byref(it), C.byref(reason), C.byref(sec)))
        return it.value, rn[: it.value + 1].copy(), reason.value, sec.value

    # measurement
    def launch_count(self):
        return self.L.mgb_launch_count(self.h)

    def last_solve_ms(self):
        return self.L.mgb_last_solve_ms(self.h)

    def time_op(self, op, l, reps=20):
        ms = C.c_double()
        self._ck(self.L.mgb_time_op(self.h, OPS[op][0], l, reps, C.byref(ms)))
        return ms.value


def _opt_int(options, key, default=0):
    import re
    m = re.search(r"(?:^|\s)" + re.escape(key) + r"\s+(\d+)", options)
    return int(m.group(1)) if m else default


class Session:
    """The reference's main() in two steps through the host C layer: __init__ = SetUpProblem .. Assemble
    (src/poisson.c:45-118), solve() = Solve .. PrintInfo (:123-128), close() = Destroy* (:130-134).
    .engine is the assembled engine (borrowed), for kernel-level calls between the two steps."""

    def __init__(self, options):
        self.H = host_lib()
        self.options = options
        self.s = C.c_void_p()
        rc = self.H.pb200_open(options.encode(), C.byref(self.s))
        if rc != 0:
            raise MgbError(f"pb200_open failed ({rc}): {self.H.pb200_last_error().decode()}")
        self.levels = _opt_int(options, "-levels")
        self.n = _opt_int(options, "-npts") - 2
        self.engine = Engine(self.levels, self.n, _borrow=self.H.pb200_session_engine(self.s))

    def solve(self, out_dir=None, want_u=True):
        cap = _opt_int(self.options, "-iter") + 2
        res = RunResult()
        rn = np.full(cap, np.nan)
        u = np.zeros((self.n, self.n)) if want_u else None
        rc = self.H.pb200_solve(self.s, out_dir.encode() if out_dir else None, C.byref(res),
                                _pd(u) if want_u else None, _pd(rn), cap)
        if rc != 0:
            raise MgbError(f"pb200_solve failed ({rc}): {self.H.pb200_last_error().decode()}")
        return {"num_iter": res.num_iter, "rnorm": rn[: res.num_iter + 1].copy(), "error": np.array(res.error[:]),
                "u": u, "levels": res.levels, "gpu_launches": res.gpu_launches, "solve_seconds": res.solve_seconds}

    def solve_rhs(self, b_ptr, u_ptr):
        """One more Solve() with the right-hand side at host address b_ptr (ni*nj doubles) and the solution copied
        to host address u_ptr (both plain integers / ctypes pointers, e.g. pinned torch buffers)."""
        cap = _opt_int(self.options, "-iter") + 2
        res = RunResult()
        rn = np.full(cap, np.nan)
        rc = self.H.pb200_solve_rhs(self.s, C.c_void_p(b_ptr), C.c_void_p(u_ptr), C.byref(res), _pd(rn), cap)
        if rc != 0:
            raise MgbError(f"pb200_solve_rhs failed ({rc}): {self.H.pb200_last_error().decode()}")
        return {"num_iter": res.num_iter, "rnorm": rn[: res.num_iter + 1].copy(), "gpu_launches": res.gpu_launches}

    def solve_rhs_many(self, b_ptrs, u_ptrs):
        """Solve() for a stream of right-hand sides (host addresses, whole-grid arrays, ideally pinned): uploads,
        solves and downloads are pipelined.  Returns (iterations per solve, final relative residuals, seconds)."""
        n = len(b_ptrs)
        assert n == len(u_ptrs) and n >= 1
        B = (C.c_void_p * n)(*[C.c_void_p(int(p)) for p in b_ptrs])
        U = (C.c_void_p * n)(*[C.c_void_p(int(p)) for p in u_ptrs])
        its = (C.c_int * n)()
        fin = (C.c_double * n)()
        sec = C.c_double()
        rc = self.H.pb200_solve_rhs_many(self.s, n, B, U, its, fin, C.byref(sec))
        if rc != 0:
            raise MgbError(f"pb200_solve_rhs_many failed ({rc}): {self.H.pb200_last_error().decode()}")
        return list(its), list(fin), sec.value

    def close(self):
        if getattr(self, "s", None):
            self.engine.close()
            self.H.pb200_close(self.s)
            self.s = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def run_poisson(options, out_dir=None, want_u=True, rnorm_cap=None):
    """The reference's main() pipeline (poisson.in vocabulary) through the host C layer.
    Returns dict(num_iter, rnorm, error, u, solve_seconds, gpu_launches)."""
    H = host_lib()
    import re
    m = re.search(r"-iter\s+(\d+)", options)
    cap = (int(m.group(1)) if m else 0) + 2 if rnorm_cap is None else rnorm_cap
    m = re.search(r"-npts\s+(\d+)", options)
    n = (int(m.group(1)) - 2) if m else 0
    res = RunResult()
    rn = np.full(cap, np.nan)
    u = np.zeros((max(n, 1), max(n, 1))) if want_u else None
    rc = H.pb200_run(options.encode(), out_dir.encode() if out_dir else None, C.byref(res),
                     _pd(u) if want_u else None, _pd(rn), cap)
    if rc != 0:
        raise MgbError(f"pb200_run failed ({rc}): {H.pb200_last_error().decode()}")
    return {"num_iter": res.num_iter, "rnorm": rn[: res.num_iter + 1].copy(), "error": np.array(res.error[:]),
            "u": u, "levels": res.levels, "gpu_launches": res.gpu_launches, "solve_seconds": res.solve_seconds}

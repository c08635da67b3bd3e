"""oracle -- CPU checkers for the AMGCL solve phase.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  Nothing under amgcl_b200/ does.

  oracle.c      ctypes view of liboracle.so   (plain-C restatement, amg_oracle.c)
  oracle.ref    ctypes view of _ref/libamgcl_ref.so (the REAL reference: AMGCL's
                builtin backend compiled from /root/reference; prebuilt copy is
                shipped to the GPU box, where /root/reference does not exist)
"""
import ctypes as _c
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_C = os.path.join(HERE, "liboracle.so")
LIB_REF = os.path.join(HERE, "_ref", "libamgcl_ref.so")

_i64 = _c.c_int64
_dbl = _c.c_double
_vp = _c.c_void_p
_P = _c.POINTER


def build(quiet=True):
    """make -C oracle (C restatement always; _ref only when the reference tree exists)."""
    out = subprocess.run(["make", "-C", HERE, "all"], stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout)
    if not quiet:
        print(out.stdout)


def _arr(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _p(a):
    return a.ctypes.data_as(_vp)


# ---------------------------------------------------------------------------
# plain-C restatement
# ---------------------------------------------------------------------------
class _COracle:
    def __init__(self):
        if not os.path.isfile(LIB_C) or os.path.getmtime(LIB_C) < os.path.getmtime(
                os.path.join(HERE, "amg_oracle.c")):
            build()
        L = _c.CDLL(LIB_C)
        L.orc_inner_product.restype = _dbl
        L.orc_inner_product.argtypes = [_i64, _vp, _vp]
        L.orc_spmv.argtypes = [_i64, _vp, _vp, _vp, _dbl, _vp, _dbl, _vp]
        L.orc_residual.argtypes = [_i64, _vp, _vp, _vp, _vp, _vp, _vp]
        L.orc_axpby.argtypes = [_i64, _dbl, _vp, _dbl, _vp]
        L.orc_axpbypcz.argtypes = [_i64, _dbl, _vp, _dbl, _vp, _dbl, _vp]
        L.orc_vmul.argtypes = [_i64, _dbl, _vp, _vp, _dbl, _vp]
        L.orc_jacobi_diag.argtypes = [_i64, _vp, _vp, _vp, _vp]
        L.orc_spai0_diag.argtypes = [_i64, _vp, _vp, _vp, _vp]
        L.orc_relax.argtypes = [_i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _dbl]
        L.orc_hier_create.restype = _vp
        L.orc_hier_create.argtypes = [_c.c_int]
        L.orc_hier_destroy.argtypes = [_vp]
        L.orc_hier_add_level.argtypes = [_vp, _i64, _vp, _vp, _vp, _i64, _vp, _vp, _vp,
                                         _vp, _vp, _vp, _vp, _dbl]
        L.orc_hier_set_coarse.argtypes = [_vp, _i64, _vp, _vp, _vp]
        L.orc_hier_set_coarse.restype = _c.c_int
        L.orc_coarse_solve.argtypes = [_vp, _vp, _vp]
        L.orc_amg_apply.argtypes = [_vp, _vp, _vp]
        L.orc_cg.argtypes = [_vp, _vp, _vp, _dbl, _c.c_int, _P(_i64), _P(_dbl), _vp]
        L.orc_cg.restype = _c.c_int
        L.orc_bicgstab.argtypes = [_vp, _vp, _vp, _dbl, _c.c_int, _P(_i64), _P(_dbl), _vp]
        L.orc_bicgstab.restype = _c.c_int
        for name in ("orc_spmv", "orc_residual", "orc_axpby", "orc_axpbypcz", "orc_vmul",
                     "orc_jacobi_diag", "orc_spai0_diag", "orc_relax", "orc_hier_destroy",
                     "orc_hier_add_level", "orc_coarse_solve", "orc_amg_apply"):
            getattr(L, name).restype = None
        self.L = L

    # primitives: numpy in, numpy out ------------------------------------------------
    def spmv(self, alpha, A, x, beta, y):
        ptr, col, val = A
        y = _arr(y, np.float64).copy()
        x = _arr(x, np.float64)
        self.L.orc_spmv(ptr.size - 1, _p(ptr), _p(col), _p(val), alpha, _p(x), beta, _p(y))
        return y

    def residual(self, f, A, x):
        ptr, col, val = A
        f = _arr(f, np.float64)
        x = _arr(x, np.float64)
        r = np.empty(ptr.size - 1)
        self.L.orc_residual(ptr.size - 1, _p(ptr), _p(col), _p(val), _p(f), _p(x), _p(r))
        return r

    def inner_product(self, x, y):
        x = _arr(x, np.float64)
        y = _arr(y, np.float64)
        return self.L.orc_inner_product(x.size, _p(x), _p(y))

    def axpby(self, a, x, b, y):
        x = _arr(x, np.float64)
        y = _arr(y, np.float64).copy()
        self.L.orc_axpby(x.size, a, _p(x), b, _p(y))
        return y

    def axpbypcz(self, a, x, b, y, c, z):
        x = _arr(x, np.float64)
        y = _arr(y, np.float64)
        z = _arr(z, np.float64).copy()
        self.L.orc_axpbypcz(x.size, a, _p(x), b, _p(y), c, _p(z))
        return z

    def vmul(self, a, x, y, b, z):
        x = _arr(x, np.float64)
        y = _arr(y, np.float64)
        z = _arr(z, np.float64).copy()
        self.L.orc_vmul(x.size, a, _p(x), _p(y), b, _p(z))
        return z

    def relax_diag(self, A, relax):
        ptr, col, val = A
        d = np.zeros(ptr.size - 1)
        fn = self.L.orc_jacobi_diag if relax == "damped_jacobi" else self.L.orc_spai0_diag
        fn(ptr.size - 1, _p(ptr), _p(col), _p(val), _p(d))
        return d

    def relax(self, A, rhs, x, diag, omega):
        ptr, col, val = A
        rhs = _arr(rhs, np.float64)
        x = _arr(x, np.float64).copy()
        diag = _arr(diag, np.float64)
        tmp = np.empty_like(x)
        self.L.orc_relax(ptr.size - 1, _p(ptr), _p(col), _p(val), _p(rhs), _p(x), _p(tmp),
                         _p(diag), omega)
        return x


class Hierarchy:
    """C-oracle AMG hierarchy built from explicit level operators.

    levels: list of dicts {A, P, R, diag} with A/P/R = (ptr, col, val) int64/float64
    coarse: (ptr, col, val) of the coarsest matrix; omega: smoother damping."""

    def __init__(self, levels, coarse, omega):
        self.c = c()
        L = self.c.L
        self._keep = []
        self.h = L.orc_hier_create(len(levels) + 1)
        for lv in levels:
            arrs = []
            for key in ("A", "P", "R"):
                ptr, col, val = lv[key]
                arrs += [_arr(ptr, np.int64), _arr(col, np.int64), _arr(val, np.float64)]
            diag = _arr(lv["diag"], np.float64)
            self._keep += arrs + [diag]
            n = arrs[0].size - 1
            nc = arrs[6].size - 1
            L.orc_hier_add_level(self.h, n, _p(arrs[0]), _p(arrs[1]), _p(arrs[2]),
                                 nc, _p(arrs[3]), _p(arrs[4]), _p(arrs[5]),
                                 _p(arrs[6]), _p(arrs[7]), _p(arrs[8]), _p(diag), omega)
        cp, cc, cv = (_arr(coarse[0], np.int64), _arr(coarse[1], np.int64),
                      _arr(coarse[2], np.float64))
        self._keep += [cp, cc, cv]
        self.n = levels[0]["A"][0].size - 1 if levels else cp.size - 1
        self.nc = cp.size - 1
        if L.orc_hier_set_coarse(self.h, self.nc, _p(cp), _p(cc), _p(cv)) != 0:
            raise RuntimeError("oracle: zero pivot in coarse LU")

    def coarse_solve(self, rhs):
        rhs = _arr(rhs, np.float64)
        x = np.zeros(self.nc)
        self.c.L.orc_coarse_solve(self.h, _p(rhs), _p(x))
        return x

    def apply(self, rhs):
        rhs = _arr(rhs, np.float64)
        x = np.zeros(self.n)
        self.c.L.orc_amg_apply(self.h, _p(rhs), _p(x))
        return x

    def solve(self, rhs, krylov="cg", tol=1e-8, maxiter=100, x0=None):
        rhs = _arr(rhs, np.float64)
        x = np.zeros(self.n) if x0 is None else _arr(x0, np.float64).copy()
        it = _i64()
        res = _dbl()
        hist = np.zeros(maxiter)
        fn = self.c.L.orc_cg if krylov == "cg" else self.c.L.orc_bicgstab
        rc = fn(self.h, _p(rhs), _p(x), tol, maxiter, _c.byref(it), _c.byref(res), _p(hist))
        if rc != 0:
            raise RuntimeError("oracle: Krylov breakdown")
        return x, it.value, res.value, hist[:it.value]

    def __del__(self):
        try:
            self.c.L.orc_hier_destroy(self.h)
        except Exception:
            pass


_c_inst = None


def c():
    global _c_inst
    if _c_inst is None:
        _c_inst = _COracle()
    return _c_inst


# ---------------------------------------------------------------------------
# the real reference (AMGCL builtin backend)
# ---------------------------------------------------------------------------
RELAX = {"damped_jacobi": 0, "spai0": 1, "chebyshev": 2, "ilu0": 3}
KRYLOV = {"cg": 0, "bicgstab": 1, "gmres": 2, "bicgstabl": 3}


def have_ref():
    if os.path.isfile(LIB_REF):
        return True
    if os.path.isfile("/root/reference/amgcl/amg.hpp"):
        try:
            build()
        except Exception:
            return False
        return os.path.isfile(LIB_REF)
    return False


class _Ref:
    def __init__(self):
        if not have_ref():
            raise RuntimeError("oracle/_ref/libamgcl_ref.so is not available")
        R = _c.CDLL(LIB_REF)
        R.ref_last_error.restype = _c.c_char_p
        R.ref_num_threads.restype = _c.c_int
        R.ref_set_num_threads.argtypes = [_c.c_int]
        R.ref_set_num_threads.restype = None
        R.ref_create.argtypes = [_i64, _vp, _vp, _vp, _c.c_int, _c.c_int, _dbl, _c.c_int,
                                 _c.c_int, _P(_vp)]
        R.ref_create_mixed.argtypes = R.ref_create.argtypes
        R.ref_destroy.argtypes = [_vp]
        R.ref_destroy.restype = None
        R.ref_solve.argtypes = [_vp, _vp, _vp, _P(_i64), _P(_dbl)]
        R.ref_solve_timed.argtypes = [_vp, _vp, _vp, _P(_i64), _P(_dbl), _P(_dbl)]
        R.ref_solve_timed.restype = _c.c_int
        R.ref_apply_precond.argtypes = [_vp, _vp, _vp]
        R.ref_report.argtypes = [_vp, _c.c_char_p, _i64]
        R.ref_report.restype = _i64
        R.ref_nlevels.argtypes = [_vp]
        R.ref_level_info.argtypes = [_vp, _c.c_int, _c.c_int, _P(_i64), _P(_i64), _P(_i64)]
        R.ref_level_matrix.argtypes = [_vp, _c.c_int, _c.c_int, _vp, _vp, _vp]
        R.ref_level_diag.argtypes = [_vp, _c.c_int, _vp]
        R.ref_coarse_solve.argtypes = [_vp, _vp, _vp]
        R.ref_spmv.argtypes = [_i64, _i64, _vp, _vp, _vp, _dbl, _vp, _dbl, _vp]
        R.ref_spmv.restype = None
        R.ref_residual.argtypes = [_i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp]
        R.ref_residual.restype = None
        R.ref_inner_product.argtypes = [_i64, _vp, _vp]
        R.ref_inner_product.restype = _dbl
        R.ref_axpby.argtypes = [_i64, _dbl, _vp, _dbl, _vp]
        R.ref_axpby.restype = None
        R.ref_axpbypcz.argtypes = [_i64, _dbl, _vp, _dbl, _vp, _dbl, _vp]
        R.ref_axpbypcz.restype = None
        R.ref_vmul.argtypes = [_i64, _dbl, _vp, _vp, _dbl, _vp]
        R.ref_vmul.restype = None
        R.ref_relax_diag.argtypes = [_i64, _vp, _vp, _vp, _c.c_int, _vp]
        R.ref_relax_diag.restype = None
        R.ref_mm_read_crs.argtypes = [_c.c_char_p, _i64, _i64, _P(_i64), _P(_i64), _P(_i64), _vp, _vp, _vp]
        R.ref_mm_read_dense.argtypes = [_c.c_char_p, _i64, _i64, _P(_i64), _P(_i64), _vp]
        R.ref_mm_write_crs.argtypes = [_c.c_char_p, _i64, _i64, _vp, _vp, _vp]
        R.ref_mm_write_dense.argtypes = [_c.c_char_p, _vp, _i64, _i64]
        R.ref_bin_read_crs.argtypes = [_c.c_char_p, _i64, _i64, _P(_i64), _P(_i64), _vp, _vp, _vp]
        R.ref_bin_write_crs.argtypes = [_c.c_char_p, _i64, _vp, _vp, _vp]
        R.ref_bin_read_dense.argtypes = [_c.c_char_p, _i64, _i64, _P(_i64), _P(_i64), _vp]
        R.ref_sample_problem.argtypes = [_i64, _dbl, _vp, _vp, _vp, _vp]
        R.ref_sample_problem.restype = _i64
        self.R = R

    def sample_problem(self, n, anisotropy=1.0):
        """tests/sample_problem.hpp:11-82 -> (ptr, col, val, rhs)."""
        nnz = self.R.ref_sample_problem(n, anisotropy, None, None, None, None)
        ptr, col = np.empty(n ** 3 + 1, np.int64), np.empty(nnz, np.int64)
        val, rhs = np.empty(nnz), np.empty(n ** 3)
        self.R.ref_sample_problem(n, anisotropy, _p(ptr), _p(col), _p(val), _p(rhs))
        return ptr, col, val, rhs

    # file formats (io/mm.hpp, io/binary.hpp) -------------------------------
    def _io_fail(self, what):
        raise RuntimeError(what + ": " + self.R.ref_last_error().decode(errors="replace"))

    def mm_read_crs(self, path, rows=(-1, -1)):
        n, m, nnz = _i64(), _i64(), _i64()
        b = str(path).encode()
        if self.R.ref_mm_read_crs(b, rows[0], rows[1], _c.byref(n), _c.byref(m), _c.byref(nnz), None, None, None):
            self._io_fail("mm_read_crs")
        ptr, col, val = np.empty(n.value + 1, np.int64), np.empty(nnz.value, np.int64), np.empty(nnz.value)
        if self.R.ref_mm_read_crs(b, rows[0], rows[1], _c.byref(n), _c.byref(m), _c.byref(nnz), _p(ptr), _p(col), _p(val)):
            self._io_fail("mm_read_crs")
        return n.value, m.value, ptr, col, val

    def mm_read_dense(self, path, rows=(-1, -1)):
        n, m = _i64(), _i64()
        b = str(path).encode()
        if self.R.ref_mm_read_dense(b, rows[0], rows[1], _c.byref(n), _c.byref(m), None):
            self._io_fail("mm_read_dense")
        out = np.empty((n.value, m.value))
        if self.R.ref_mm_read_dense(b, rows[0], rows[1], _c.byref(n), _c.byref(m), _p(out)):
            self._io_fail("mm_read_dense")
        return out

    def mm_write_crs(self, path, ncols, ptr, col, val):
        ptr, col, val = _arr(ptr, np.int64), _arr(col, np.int64), _arr(val, np.float64)
        if self.R.ref_mm_write_crs(str(path).encode(), ptr.size - 1, ncols, _p(ptr), _p(col), _p(val)):
            self._io_fail("mm_write_crs")

    def mm_write_dense(self, path, a):
        a = _arr(a, np.float64)
        a2 = a.reshape(a.shape[0], -1)
        if self.R.ref_mm_write_dense(str(path).encode(), _p(a2), a2.shape[0], a2.shape[1]):
            self._io_fail("mm_write_dense")

    def bin_read_crs(self, path, rows=(-1, -1)):
        n, nnz = _i64(), _i64()
        b = str(path).encode()
        if self.R.ref_bin_read_crs(b, rows[0], rows[1], _c.byref(n), _c.byref(nnz), None, None, None):
            self._io_fail("bin_read_crs")
        ptr, col, val = np.empty(n.value + 1, np.int64), np.empty(nnz.value, np.int64), np.empty(nnz.value)
        if self.R.ref_bin_read_crs(b, rows[0], rows[1], _c.byref(n), _c.byref(nnz), _p(ptr), _p(col), _p(val)):
            self._io_fail("bin_read_crs")
        return n.value, ptr, col, val

    def bin_write_crs(self, path, ptr, col, val):
        ptr, col, val = _arr(ptr, np.int64), _arr(col, np.int64), _arr(val, np.float64)
        if self.R.ref_bin_write_crs(str(path).encode(), ptr.size - 1, _p(ptr), _p(col), _p(val)):
            self._io_fail("bin_write_crs")

    def bin_read_dense(self, path, rows=(-1, -1)):
        n, m = _i64(), _i64()
        b = str(path).encode()
        if self.R.ref_bin_read_dense(b, rows[0], rows[1], _c.byref(n), _c.byref(m), None):
            self._io_fail("bin_read_dense")
        out = np.empty((n.value, m.value))
        if self.R.ref_bin_read_dense(b, rows[0], rows[1], _c.byref(n), _c.byref(m), _p(out)):
            self._io_fail("bin_read_dense")
        return out

    @property
    def threads(self):
        return self.R.ref_num_threads()

    def set_threads(self, n):
        self.R.ref_set_num_threads(int(n))

    # primitives ------------------------------------------------------------
    def spmv(self, alpha, A, ncols, x, beta, y):
        ptr, col, val = A
        x = _arr(x, np.float64)
        y = _arr(y, np.float64).copy()
        self.R.ref_spmv(ptr.size - 1, ncols, _p(ptr), _p(col), _p(val), alpha, _p(x), beta, _p(y))
        return y

    def residual(self, f, A, ncols, x):
        ptr, col, val = A
        f = _arr(f, np.float64)
        x = _arr(x, np.float64)
        r = np.empty(ptr.size - 1)
        self.R.ref_residual(ptr.size - 1, ncols, _p(ptr), _p(col), _p(val), _p(f), _p(x), _p(r))
        return r

    def inner_product(self, x, y):
        x = _arr(x, np.float64)
        y = _arr(y, np.float64)
        return self.R.ref_inner_product(x.size, _p(x), _p(y))

    def axpby(self, a, x, b, y):
        x = _arr(x, np.float64)
        y = _arr(y, np.float64).copy()
        self.R.ref_axpby(x.size, a, _p(x), b, _p(y))
        return y

    def axpbypcz(self, a, x, b, y, c_, z):
        x = _arr(x, np.float64)
        y = _arr(y, np.float64)
        z = _arr(z, np.float64).copy()
        self.R.ref_axpbypcz(x.size, a, _p(x), b, _p(y), c_, _p(z))
        return z

    def vmul(self, a, x, y, b, z):
        x = _arr(x, np.float64)
        y = _arr(y, np.float64)
        z = _arr(z, np.float64).copy()
        self.R.ref_vmul(x.size, a, _p(x), _p(y), b, _p(z))
        return z

    def relax_diag(self, A, relax):
        ptr, col, val = A
        d = np.zeros(ptr.size - 1)
        self.R.ref_relax_diag(ptr.size - 1, _p(ptr), _p(col), _p(val), RELAX[relax], _p(d))
        return d


class RefSolver:
    """make_solver<amg<builtin<double>, smoothed_aggregation, RELAX>, KRYLOV> (the reference)."""

    def __init__(self, ptr, col, val, relax="damped_jacobi", krylov="cg", tol=1e-8,
                 maxiter=100, coarse_enough=-1, precision="f64"):
        """precision 'mixed': amg<builtin<float>> under a builtin<double> Krylov solver."""
        self.r = ref()
        self.ptr = _arr(ptr, np.int64)
        self.col = _arr(col, np.int64)
        self.val = _arr(val, np.float64)
        self.n = self.ptr.size - 1
        self.h = _vp()
        create = self.r.R.ref_create_mixed if precision == "mixed" else self.r.R.ref_create
        rc = create(self.n, _p(self.ptr), _p(self.col), _p(self.val), RELAX[relax],
                                 KRYLOV[krylov], tol, maxiter, coarse_enough, _c.byref(self.h))
        if rc != 0:
            raise RuntimeError("ref_create: " + self.r.R.ref_last_error().decode())

    def solve(self, rhs, x0=None):
        rhs = _arr(rhs, np.float64)
        x = np.zeros(self.n) if x0 is None else _arr(x0, np.float64).copy()
        it = _i64()
        res = _dbl()
        if self.r.R.ref_solve(self.h, _p(rhs), _p(x), _c.byref(it), _c.byref(res)) != 0:
            raise RuntimeError("ref_solve: " + self.r.R.ref_last_error().decode())
        return x, it.value, res.value

    def solve_timed(self, rhs, x0=None):
        """As solve(); the 4th value is the time of solve() alone, measured inside the library."""
        rhs = _arr(rhs, np.float64)
        x = np.zeros(self.n) if x0 is None else _arr(x0, np.float64).copy()
        it = _i64()
        res = _dbl()
        sec = _dbl()
        if self.r.R.ref_solve_timed(self.h, _p(rhs), _p(x), _c.byref(it), _c.byref(res),
                                    _c.byref(sec)) != 0:
            raise RuntimeError("ref_solve_timed: " + self.r.R.ref_last_error().decode())
        return x, it.value, res.value, sec.value

    def apply_precond(self, f):
        f = _arr(f, np.float64)
        x = np.zeros(self.n)
        if self.r.R.ref_apply_precond(self.h, _p(f), _p(x)) != 0:
            raise RuntimeError("ref_apply_precond: " + self.r.R.ref_last_error().decode())
        return x

    def report(self):
        need = self.r.R.ref_report(self.h, None, 0)
        buf = _c.create_string_buffer(int(need))
        self.r.R.ref_report(self.h, buf, need)
        return buf.value.decode()

    @property
    def nlevels(self):
        return self.r.R.ref_nlevels(self.h)

    def level_matrix(self, lvl, which):
        """which: 'A' | 'P' | 'R'. Returns (nrows, ncols, (ptr, col, val))."""
        w = {"A": 0, "P": 1, "R": 2}[which]
        rows, cols, nnz = _i64(), _i64(), _i64()
        if self.r.R.ref_level_info(self.h, lvl, w, _c.byref(rows), _c.byref(cols),
                                   _c.byref(nnz)) != 0:
            raise KeyError("no operator %s on level %d" % (which, lvl))
        ptr = np.empty(rows.value + 1, dtype=np.int64)
        col = np.empty(nnz.value, dtype=np.int64)
        val = np.empty(nnz.value, dtype=np.float64)
        self.r.R.ref_level_matrix(self.h, lvl, w, _p(ptr), _p(col), _p(val))
        return rows.value, cols.value, (ptr, col, val)

    def level_diag(self, lvl, n):
        d = np.empty(n)
        if self.r.R.ref_level_diag(self.h, lvl, _p(d)) != 0:
            raise KeyError("no smoother diagonal on level %d" % lvl)
        return d

    def coarse_solve(self, rhs):
        rhs = _arr(rhs, np.float64)
        x = np.zeros(rhs.size)
        if self.r.R.ref_coarse_solve(self.h, _p(rhs), _p(x)) != 0:
            raise RuntimeError(self.r.R.ref_last_error().decode())
        return x

    def hierarchy(self):
        """All level operators: (levels=[{A,P,R,diag}], coarse=(ptr,col,val))."""
        nl = self.nlevels
        levels = []
        for l in range(nl - 1):
            n, _, A = self.level_matrix(l, "A")
            _, _, P = self.level_matrix(l, "P")
            _, _, R = self.level_matrix(l, "R")
            levels.append({"A": A, "P": P, "R": R, "diag": self.level_diag(l, n)})
        _, _, C = self.level_matrix(nl - 1, "A")
        return levels, C

    def close(self):
        if self.h:
            self.r.R.ref_destroy(self.h)
            self.h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_ref_inst = None


def ref():
    global _ref_inst
    if _ref_inst is None:
        _ref_inst = _Ref()
    return _ref_inst

"""ctypes binding of the CPU oracle (oracle/amg_oracle.c).

TEST INFRASTRUCTURE ONLY.  Import this from tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs -- never from the product
package.  The oracle restates the reference's arithmetic
(/root/reference/include/amg/*.hpp); see the header of amg_oracle.c for the
parity status and the per-function reference citations.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

SMOOTHER_GS = 0
SMOOTHER_JACOBI = 1
SMOOTHER_COLOR_GS = 2


def build(force=False):
    """Compile liboracle.so with the committed Makefile (gcc, no FMA)."""
    src = os.path.join(_HERE, "amg_oracle.c")
    if (force or not os.path.exists(_LIB_PATH)
            or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return _LIB_PATH


_lib = None
_p = C.c_void_p
_i = C.c_int
_l = C.c_int64
_d = C.c_double
_pd = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_pi = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    sig = {
        "orc_csc_free": (None, [_p]),
        "orc_csc_rows": (_i, [_p]),
        "orc_csc_cols": (_i, [_p]),
        "orc_csc_nnz": (_l, [_p]),
        "orc_csc_nnz_nonzero": (_l, [_p]),
        "orc_csc_copy_out": (None, [_p, _pi, _pi, _pd]),
        "orc_csc_from_arrays": (_p, [_i, _i, _pi, _pi, _pd]),
        "orc_grid_spacing_h": (_d, [_i]),
        "orc_points_n_from_grid_spacing_h": (_l, [_d]),
        "orc_laplacian": (_p, [_i]),
        "orc_laplacian_aniso": (_p, [_i, _d]),
        "orc_rhs": (None, [_i, _pd]),
        "orc_n_H_dofs_from_n_h_dofs": (_l, [_l]),
        "orc_make_P": (_p, [_i, _i]),
        "orc_transpose": (_p, [_p]),
        "orc_spgemm": (_p, [_p, _p]),
        "orc_galerkin": (_p, [_p, _p, _p]),
        "orc_gs_forward": (None, [_p, _pd, _pd]),
        "orc_gs_backward": (None, [_p, _pd, _pd]),
        "orc_rss": (_d, [_p, _pd, _pd]),
        "orc_gs_smooth": (_l, [_p, _pd, _pd, _d, _l, _l, C.POINTER(_d)]),
        "orc_residual": (None, [_p, _pd, _pd, _pd]),
        "orc_spmv": (None, [_p, _pd, _pd]),
        "orc_jacobi_sweep": (None, [_p, _pd, _pd, _d, _pd]),
        "orc_greedy_coloring": (_i, [_p, _p, _pi]),
        "orc_color_gs_pass": (None, [_p, _pi, _i, _pd, _pd]),
        "orc_ldlt_factor": (_p, [_p]),
        "orc_ldlt_n": (_i, [_p]),
        "orc_ldlt_bw": (_i, [_p]),
        "orc_ldlt_copy_out": (None, [_p, _pd, _pd]),
        "orc_ldlt_solve": (None, [_p, _pd, _pd]),
        "orc_ldlt_free": (None, [_p]),
        "orc_mg_create": (_p, [_p, _pd, _l, _i, _d, _l, _l, _i, _i, _d,
                               C.POINTER(_i)]),
        "orc_mg_free": (None, [_p]),
        "orc_mg_set_smoother": (None, [_p, _i, _i, _d]),
        "orc_mg_reset": (None, [_p]),
        "orc_mg_n_levels": (_i, [_p]),
        "orc_mg_n_dofs": (_l, [_p, _i]),
        "orc_mg_A": (_p, [_p, _i]),
        "orc_mg_P": (_p, [_p, _i]),
        "orc_mg_R": (_p, [_p, _i]),
        "orc_mg_u": (C.POINTER(_d), [_p, _i]),
        "orc_mg_f": (C.POINTER(_d), [_p, _i]),
        "orc_mg_r": (C.POINTER(_d), [_p, _i]),
        "orc_mg_color": (C.POINTER(_i), [_p, _i]),
        "orc_mg_n_colors": (_i, [_p, _i]),
        "orc_mg_coarse": (_p, [_p]),
        "orc_mg_iters_done": (_l, [_p]),
        "orc_mg_last_error": (_d, [_p]),
        "orc_mg_hist": (_i, [_p, _pd, _i]),
        "orc_mg_smooth": (None, [_p, _i]),
        "orc_mg_vcycle": (None, [_p]),
        "orc_mg_rss": (_d, [_p]),
        "orc_mg_solve": (_l, [_p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


class Csc:
    """Owning handle on an oracle CSC matrix (int32 indices, fp64 values)."""

    def __init__(self, ptr, owned=True):
        self.ptr = ptr
        self.owned = owned

    def __del__(self):
        if getattr(self, "owned", False) and self.ptr and _lib is not None:
            _lib.orc_csc_free(self.ptr)
            self.ptr = None

    rows = property(lambda s: lib().orc_csc_rows(s.ptr))
    cols = property(lambda s: lib().orc_csc_cols(s.ptr))
    nnz = property(lambda s: lib().orc_csc_nnz(s.ptr))
    nnz_nonzero = property(lambda s: lib().orc_csc_nnz_nonzero(s.ptr))

    def arrays(self):
        colptr = np.empty(self.cols + 1, np.int32)
        rowidx = np.empty(self.nnz, np.int32)
        val = np.empty(self.nnz, np.float64)
        lib().orc_csc_copy_out(self.ptr, colptr, rowidx, val)
        return colptr, rowidx, val

    @staticmethod
    def from_arrays(rows, cols, colptr, rowidx, val):
        colptr = np.ascontiguousarray(colptr, np.int32)
        rowidx = np.ascontiguousarray(rowidx, np.int32)
        val = np.ascontiguousarray(val, np.float64)
        return Csc(lib().orc_csc_from_arrays(rows, cols, colptr, rowidx, val))

    def transpose(self):
        return Csc(lib().orc_transpose(self.ptr))

    def to_scipy(self):
        import scipy.sparse as sp
        colptr, rowidx, val = self.arrays()
        return sp.csc_matrix((val, rowidx, colptr), shape=(self.rows, self.cols))


def laplacian(n, eps_y=1.0):
    return Csc(lib().orc_laplacian_aniso(n, float(eps_y)))


def rhs(n):
    b = np.empty(n * n, np.float64)
    lib().orc_rhs(n, b)
    return b


def make_P(n_h, n_H):
    return Csc(lib().orc_make_P(n_h, n_H))


def n_H_from_n_h(n_h):
    return lib().orc_n_H_dofs_from_n_h_dofs(n_h)


def level_sizes(n0, n_levels):
    out = [n0]
    for _ in range(1, n_levels):
        out.append(n_H_from_n_h(out[-1]))
    return out


def galerkin(R, A, P):
    return Csc(lib().orc_galerkin(R.ptr, A.ptr, P.ptr))


def gs_forward(A, b, u):
    lib().orc_gs_forward(A.ptr, b, u)


def gs_backward(A, b, u):
    lib().orc_gs_backward(A.ptr, b, u)


def gs_smooth(A, u, b, tolerance=1e-9, every=0, n_iters=1):
    err = _d(0.0)
    it = lib().orc_gs_smooth(A.ptr, u, b, tolerance, every, n_iters, C.byref(err))
    return it, err.value


def rss(A, u, b):
    return lib().orc_rss(A.ptr, u, b)


def residual(A, u, f):
    r = np.empty(A.rows, np.float64)
    lib().orc_residual(A.ptr, u, f, r)
    return r


def spmv(M, x):
    y = np.empty(M.rows, np.float64)
    lib().orc_spmv(M.ptr, x, y)
    return y


def jacobi_sweep(AT, u, f, omega):
    out = np.empty_like(u)
    lib().orc_jacobi_sweep(AT.ptr, u, f, omega, out)
    return out


def greedy_coloring(A, AT=None):
    AT = AT or A.transpose()
    color = np.empty(A.cols, np.int32)
    nc = lib().orc_greedy_coloring(A.ptr, AT.ptr, color)
    return nc, color


def color_gs_pass(AT, color, c, f, u):
    lib().orc_color_gs_pass(AT.ptr, color, c, f, u)


class Ldlt:
    def __init__(self, A):
        self.ptr = lib().orc_ldlt_factor(A.ptr)

    def __del__(self):
        if self.ptr and _lib is not None:
            _lib.orc_ldlt_free(self.ptr)
            self.ptr = None

    def solve(self, f):
        x = np.empty_like(f)
        lib().orc_ldlt_solve(self.ptr, f, x)
        return x


class Multigrid:
    """Oracle restatement of AMG::Multigrid (include/amg/multigrid.hpp:22-365)."""

    def __init__(self, A, b, n_levels, tolerance=1e-9, every=10, n_iters=100,
                 smoother=SMOOTHER_GS, smoother_iters=1, omega=2.0 / 3.0):
        err = _i(0)
        b = np.ascontiguousarray(b, np.float64)
        self.ptr = lib().orc_mg_create(A.ptr, b, b.shape[0], n_levels, tolerance,
                                       every, n_iters, smoother, smoother_iters,
                                       omega, C.byref(err))
        if not self.ptr:
            raise ValueError({1: "`compute_error_every_n_iters` must be leq to `n_iters`",
                              2: "`A` and `b` must have the same number of degrees of freedom"
                              }[err.value])
        self.n_levels = n_levels

    def __del__(self):
        if getattr(self, "ptr", None) and _lib is not None:
            _lib.orc_mg_free(self.ptr)
            self.ptr = None

    def n_dofs(self, l):
        return lib().orc_mg_n_dofs(self.ptr, l)

    def A(self, l):
        return Csc(lib().orc_mg_A(self.ptr, l), owned=False)

    def P(self, l):
        return Csc(lib().orc_mg_P(self.ptr, l), owned=False)

    def R(self, l):
        return Csc(lib().orc_mg_R(self.ptr, l), owned=False)

    def _vec(self, fn, l):
        return np.ctypeslib.as_array(fn(self.ptr, l), shape=(self.n_dofs(l),))

    def u(self, l):
        return self._vec(lib().orc_mg_u, l)

    def f(self, l):
        return self._vec(lib().orc_mg_f, l)

    def r(self, l):
        return self._vec(lib().orc_mg_r, l)

    def color(self, l):
        return np.ctypeslib.as_array(lib().orc_mg_color(self.ptr, l),
                                     shape=(self.n_dofs(l),))

    def n_colors(self, l):
        return lib().orc_mg_n_colors(self.ptr, l)

    def smooth(self, l):
        lib().orc_mg_smooth(self.ptr, l)

    def vcycle(self):
        lib().orc_mg_vcycle(self.ptr)

    def set_smoother(self, smoother, smoother_iters=1, omega=2.0 / 3.0):
        lib().orc_mg_set_smoother(self.ptr, smoother, smoother_iters, omega)

    def reset(self):
        lib().orc_mg_reset(self.ptr)

    def rss(self):
        return lib().orc_mg_rss(self.ptr)

    def solve(self):
        return lib().orc_mg_solve(self.ptr)

    @property
    def iters_done(self):
        return lib().orc_mg_iters_done(self.ptr)

    @property
    def last_error(self):
        return lib().orc_mg_last_error(self.ptr)

    def history(self):
        buf = np.empty(4096, np.float64)
        n = lib().orc_mg_hist(self.ptr, buf, 4096)
        return buf[:min(n, 4096)].copy()


# ---------------------------------------------------------------------------------------------
# Full-size checker for the 8193^2 microbenchmark: the oracle's per-row arithmetic on the
# five-point operator, vectorised with numpy (one array operation per stencil entry, in ascending
# column order, separate multiply and subtract -- numpy never fuses them).  Bit-identical to
# orc_residual / orc_jacobi_sweep above on the matrix orc_laplacian_aniso generates
# (tests/test_oracle_golden.py checks that at small n); needs no 4 GB CSC matrix.
# ---------------------------------------------------------------------------------------------
def five_point_coefficients(n, eps_y=1.0):
    """(cross, line, diag) = the entries at offsets -+n, -+1, 0 of orc_laplacian_aniso(n, eps_y)."""
    h = lib().orc_grid_spacing_h(n)   # grid.hpp:31
    hh = h * h
    d_off, d_dia = 1.0 / hh, -2.0 / hh   # grid.hpp:62-72
    return eps_y * d_off, d_off, d_dia + eps_y * d_dia


def five_point_residual(n, eps_y, u, f):
    """r = f - A u in the reference's order ((f - a1 u1) - a2 u2) - ... (multigrid.hpp:272-274)."""
    cross, line, diag = five_point_coefficients(n, eps_y)
    U = u.reshape(n, n)
    acc = f.reshape(n, n).copy()
    acc[1:, :] -= cross * U[:-1, :]      # offset -n
    acc[:, 1:] -= line * U[:, :-1]       # offset -1
    acc -= diag * U                      # diagonal
    acc[:, :-1] -= line * U[:, 1:]       # offset +1
    acc[:-1, :] -= cross * U[1:, :]      # offset +n
    return acc.reshape(-1)


def five_point_jacobi(n, eps_y, u, f, omega):
    """u + omega * (r / d), the oracle's damped-Jacobi sweep (orc_jacobi_sweep)."""
    _, _, diag = five_point_coefficients(n, eps_y)
    r = five_point_residual(n, eps_y, u, f)
    return u + omega * (r / diag)

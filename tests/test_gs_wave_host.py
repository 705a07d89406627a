"""Lane logic of the multi-SM lexicographic Gauss-Seidel kernel (algebraic-multigrid_b200/csrc/gs_wave.cuh)
on the CPU: tests/cpp/gs_wave_host.cpp steps every block's 32 lanes in lockstep, interleaves the blocks
in random orders that only respect the hand-over buffer, reads inputs through the same look-ahead ring,
and must reproduce the plain row-by-row sweep bit for bit -- forward and backward, five / seven / nine
point stencils, unsymmetric coefficients, pruned entries, zero diagonals, one to several blocks."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "gs_wave_host.cpp")
HDR = os.path.join(ROOT, "algebraic-multigrid_b200", "csrc", "gs_wave.cuh")
LIB = os.path.join(ROOT, "tests", "cpp", "libgs_wave_host.so")

STENCILS = {
    "five": [(-1, 0), (0, -1), (0, 0), (0, 1), (1, 0)],
    "seven_a": [(-1, -1), (-1, 0), (0, -1), (0, 0), (0, 1), (1, 0), (1, 1)],
    "seven_b": [(-1, 0), (-1, 1), (0, -1), (0, 0), (0, 1), (1, -1), (1, 0)],
    "nine": [(a, d) for a in (-1, 0, 1) for d in (-1, 0, 1)],
}


def lib():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wall",
                               "-x", "c++", SRC, "-o", LIB])
    L = C.CDLL(LIB)
    pi = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
    pd = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    L.gsw_host_sweep.restype = C.c_int
    L.gsw_host_sweep.argtypes = [C.c_int, C.c_int, pi, C.c_int, pd, pd, pd, C.c_int, C.c_int, C.c_uint,
                                 C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.gsw_host_reference.restype = None
    L.gsw_host_reference.argtypes = [C.c_int, C.c_int, pi, C.c_int, pd, pd, pd, C.c_int]
    return L


def grid_operator(n_lines, m, stencil, seed, holes=0.0, zero_diag=0.0):
    """Random unsymmetric, diagonally dominant operator on an n_lines x m grid in DIA row form."""
    rng = np.random.default_rng(seed)
    n = n_lines * m
    offs = np.array(sorted(a * m + d for a, d in stencil), dtype=np.int32)
    ld = (n + 31) // 32 * 32
    val = np.zeros((len(offs), ld))
    y, x = np.divmod(np.arange(n), m)
    for a, d in stencil:
        inside = (y + a >= 0) & (y + a < n_lines) & (x + d >= 0) & (x + d < m)
        v = 8.0 + rng.random(n) if (a, d) == (0, 0) else -(0.5 + rng.random(n))
        if (a, d) != (0, 0) and holes:
            v[rng.random(n) < holes] = 0.0
        if (a, d) == (0, 0) and zero_diag:
            v[rng.random(n) < zero_diag] = 0.0
        val[int(np.searchsorted(offs, a * m + d)), :n] = np.where(inside, v, 0.0)
    return n, offs, ld, val.reshape(-1).copy()


@pytest.mark.parametrize("n_lines,m", [(3, 3), (7, 4), (5, 5), (17, 17), (30, 9), (31, 33), (64, 64), (95, 70), (61, 129)])
@pytest.mark.parametrize("stencil", sorted(STENCILS))
def test_wave_sweep_equals_row_by_row_sweep(n_lines, m, stencil):
    L = lib()
    n, offs, ld, val = grid_operator(n_lines, m, STENCILS[stencil], seed=n_lines * 1000 + m,
                                     holes=0.1 if stencil == "nine" else 0.0, zero_diag=0.02 if m == 17 else 0.0)
    rng = np.random.default_rng(7)
    f = rng.standard_normal(n)
    u0 = rng.standard_normal(n)
    for direction in (1, -1):
        want = u0.copy()
        L.gsw_host_reference(n, len(offs), offs, ld, val, f, want, direction)
        for pd_, seed in ((1, 1), (4, 2), (6, 3)):
            got = u0.copy()
            S, mm = C.c_int(), C.c_int()
            rc = L.gsw_host_sweep(n, len(offs), offs, ld, val, f, got, direction, pd_, seed, C.byref(S), C.byref(mm))
            assert rc == 0, (rc, direction, pd_)
            assert mm.value == m
            # one row behind is enough exactly when no entry couples the row before one step ahead
            needs_two = ((-1, 1) in STENCILS[stencil]) if direction > 0 else ((1, -1) in STENCILS[stencil])
            assert S.value == (2 if needs_two else 1)
            assert got.tobytes() == want.tobytes(), (direction, pd_, np.abs(got - want).max())


def test_wave_rejects_coupling_across_the_end_of_a_line():
    """A five-point operator whose +-1 diagonal couples the last row of a line with the first of the next
    is not a grid operator: the plan must refuse it (the device falls back to the generic kernels)."""
    L = lib()
    n_lines, m = 6, 5
    n, offs, ld, val = grid_operator(n_lines, m, STENCILS["five"], seed=3)
    d = int(np.searchsorted(offs, 1))
    val.reshape(len(offs), ld)[d, m - 1] = -1.0     # row (0, m-1) -> (1, 0)
    u = np.zeros(n)
    rc = L.gsw_host_sweep(n, len(offs), offs, ld, val, np.ones(n), u, 1, 4, 1, None, None)
    assert rc == -2


def _column_rows_dia(A):
    """DIA of "row c = CSC column c" (smoother.hpp:101-117), explicit zeros dropped."""
    colptr, rowidx, val = A.arrays()
    nn = A.cols
    cols = np.repeat(np.arange(nn, dtype=np.int64), np.diff(colptr))
    keep = val != 0.0
    o = (rowidx.astype(np.int64) - cols)[keep]
    uniq = np.unique(o)
    ld = (nn + 31) // 32 * 32
    D = np.zeros((len(uniq), ld))
    D[np.searchsorted(uniq, o), cols[keep]] = val[keep]
    return nn, uniq.astype(np.int32), ld, D.reshape(-1).copy()


@pytest.mark.parametrize("n,eps", [(9, 1.0), (33, 1.0), (40, 1e-2), (70, 1.0)])
def test_wave_sweep_equals_oracle_gauss_seidel_on_level_0(n, eps):
    """Against the oracle's smoother (the reference's column-as-row sweep, forward then backward) on the
    level-0 operator."""
    L = lib()
    A = O.laplacian(n, eps)
    nn, offs, ld, D = _column_rows_dia(A)
    rng = np.random.default_rng(n)
    f, u0 = rng.standard_normal(nn), rng.standard_normal(nn)
    want = u0.copy()
    O.gs_smooth(A, want, f, 1e-9, 0, 1)
    got = u0.copy()
    for direction in (1, -1):
        S = C.c_int()
        assert L.gsw_host_sweep(nn, len(offs), offs, ld, D, f, got, direction, 4, 5, C.byref(S), None) == 0
        assert S.value == 1
    assert got.tobytes() == want.tobytes()


def test_galerkin_levels_are_not_grid_operators():
    """The reference interpolates along the FLATTENED vector (interpolator.hpp), so its Galerkin operators
    couple the last row of a grid line with the first row of the next: the chain u_k <- u_{k-1} runs through
    the whole level and no line-wavefront exists.  The plan must refuse them (they keep the scan kernel)."""
    L = lib()
    mo = O.Multigrid(O.laplacian(33), O.rhs(33), 4, 1e-9, 1, 1)
    for lvl in (1, 2, 3):
        nn, offs, ld, D = _column_rows_dia(mo.A(lvl))
        u = np.zeros(nn)
        assert L.gsw_host_sweep(nn, len(offs), offs, ld, D, np.ones(nn), u, 1, 4, 1, None, None) == -2, lvl

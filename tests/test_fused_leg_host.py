"""Tiling logic of the fused V-cycle legs (algebraic-multigrid_b200/csrc/fused_leg.cuh), run on
the CPU through the serial host Env of tests/cpp/fused_leg_host.cpp and compared bit for bit
with the oracle's unfused sequence (Jacobi sweeps, residual, restriction, prolongation)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "fused_leg_host.cpp")
HDR = os.path.join(ROOT, "algebraic-multigrid_b200", "csrc", "fused_leg.cuh")
LIB = os.path.join(ROOT, "tests", "cpp", "libfused_leg_host.so")

DOWN_U, DOWN_ZERO, UP = 0, 1, 2
OMEGA = 2.0 / 3.0


def lib():
    if (not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(HDR))):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wall",
                               "-x", "c++", SRC, "-o", LIB])
    L = C.CDLL(LIB)
    pd = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    pi = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
    pl = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
    L.leg_host_run.restype = C.c_int
    L.leg_host_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, pi, C.c_int, pd, pd, pd, pd, C.c_int,
                               C.c_double, pd, pd] + [C.c_int] * 7
    L.leg_host_plan.restype = C.c_int
    L.leg_host_plan.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, pi] + [C.c_int] * 6 + [pl]
    return L


def dia_of(A):
    """DIA of 'row c = CSC column c' with explicit zeros dropped (host_setup.hpp: Dia)."""
    colptr, rowidx, val = A.arrays()
    n = A.cols
    cols = np.repeat(np.arange(n, dtype=np.int64), np.diff(colptr))
    keep = val != 0.0
    offs = (rowidx.astype(np.int64) - cols)[keep]
    uniq = np.unique(offs)
    ld = (n + 31) // 32 * 32
    D = np.zeros((len(uniq), ld))
    D[np.searchsorted(uniq, offs), cols[keep]] = val[keep]
    return uniq.astype(np.int32), ld, D.reshape(-1).copy()


def padded(v, extra=4):
    out = np.zeros(len(v) + extra)
    out[:len(v)] = v
    return out


def hierarchy(n, eps=1.0, min_coarse=20):
    A = O.laplacian(n, eps)
    sizes = [n * n]
    while sizes[-1] > min_coarse:
        sizes.append(O.n_H_from_n_h(sizes[-1]))
    mg = O.Multigrid(A, O.rhs(n), len(sizes), 1e-9, 1, 1, O.SMOOTHER_JACOBI, 2, OMEGA)
    return mg, sizes


def run_leg(L, kind, nu, A, f, uin, e, nc, **kw):
    off, ld, val = dia_of(A)
    n = A.cols
    uout = np.full(n + 4, np.nan)
    fc = np.full(max(nc, 1) + 4, np.nan)
    z = np.zeros(4)
    rc = L.leg_host_run(kind, nu, n, len(off), off, ld, val, padded(f),
                        padded(uin) if uin is not None else z, padded(e) if e is not None else z, nc, OMEGA,
                        uout, fc, kw.get("n_sm", 148), kw.get("smem", 200 * 1024), kw.get("W", 0),
                        kw.get("LJ", 0), kw.get("PF", 0), kw.get("single", 0), kw.get("tma_mode", 0))
    return rc, uout[:n], fc[:nc]


def expect_down(A, f, u, nu, zero_guess):
    AT = A  # symmetric operators in these tests
    for _ in range(nu):
        u = O.jacobi_sweep(AT, u, f, OMEGA)
    r = O.residual(A, u, f)
    nc = O.n_H_from_n_h(A.cols)
    R = O.make_P(A.cols, nc).transpose()
    return u, O.spmv(R, r), nc


def expect_up(A, f, u, e, nu):
    nc = len(e)
    P = O.make_P(A.cols, nc)
    u = u + O.spmv(P, e)
    for _ in range(nu):
        u = O.jacobi_sweep(A, u, f, OMEGA)
    return u


CASES = [
    # (n, level, kwargs)
    (35, 0, dict(W=16, LJ=5)),
    (35, 0, dict()),
    (35, 1, dict(W=7, LJ=4)),
    (35, 2, dict(single=1, W=40)),
    (35, 3, dict()),
    (64, 0, dict(W=20, LJ=9, PF=2)),
    (64, 1, dict(W=11, LJ=3)),
    (64, 2, dict()),
    (100, 1, dict(W=25, LJ=7)),
    (100, 3, dict(W=64)),
    (129, 0, dict(W=50, LJ=13)),
    (129, 1, dict(LJ=6)),
    (129, 2, dict(W=33, LJ=3, PF=3)),
    (129, 4, dict()),
    (129, 6, dict()),
]


@pytest.mark.parametrize("n,level,kw", CASES)
@pytest.mark.parametrize("tma_mode", [0, 1])
def test_fused_legs_match_unfused_oracle(n, level, kw, tma_mode):
    L = lib()
    mg, sizes = hierarchy(n)
    if level + 1 >= len(sizes):
        pytest.skip("level not in the hierarchy")
    A = mg.A(level)
    N = A.cols
    rng = np.random.default_rng(1234 + n + level)
    f = rng.standard_normal(N)
    u0 = rng.standard_normal(N)
    kw = dict(kw, tma_mode=tma_mode)
    for nu in (1, 2, 3):
        # down leg from a given iterate (level 0 of the cycle)
        want_u, want_fc, nc = expect_down(A, f, u0, nu, False)
        rc, got_u, got_fc = run_leg(L, DOWN_U, nu, A, f, u0, None, nc, **kw)
        if rc == 1 and nu == 3:
            continue  # four chained stages of a wide band do not fit: the unfused kernels run
        assert rc == 0, (nu, "DOWN_U")
        assert np.array_equal(got_u, want_u)
        assert np.array_equal(got_fc, want_fc)
        # down leg from the zero guess (coarse levels)
        want_u, want_fc, nc = expect_down(A, f, np.zeros(N), nu, True)
        rc, got_u, got_fc = run_leg(L, DOWN_ZERO, nu, A, f, None, None, nc, **kw)
        assert rc == 0
        assert np.array_equal(got_u, want_u)
        assert np.array_equal(got_fc, want_fc)
        # up leg
        e = rng.standard_normal(nc)
        want_u = expect_up(A, f, u0, e, nu)
        rc, got_u, _ = run_leg(L, UP, nu, A, f, u0, e, nc, **kw)
        assert rc == 0
        assert np.array_equal(got_u, want_u)


def test_plans_for_the_bench_hierarchy():
    """Tiling the planner picks for the 4097^2 levels on a 148-SM part: every level fusable,
    tiles within one wave, halo overhead small on the big levels."""
    L = lib()
    n = 4097
    structs = [(n * n, [-n, -1, 0, 1, n])]
    N, m = n * n, n
    for l in range(1, 15):
        N = O.n_H_from_n_h(N)
        m = (m + 1) // 2 if l == 1 else (m - 1) // 2 + 1
        structs.append((N, None))
    info = np.zeros(10, np.int64)
    off = np.array(structs[0][1], np.int32)
    for kind, nu in ((DOWN_U, 2), (UP, 2)):
        assert L.leg_host_plan(kind, nu, structs[0][0], len(off), off, 148, 200 * 1024, 0, 0, 0, 0, info) == 0
        ok, m_, rho, W, LJ, tiles, strips, PF, threads, smem = info.tolist()
        assert ok and m_ == n and rho == 1
        assert tiles <= 148 and tiles >= 140
        assert smem <= 200 * 1024 and threads <= 1024

"""Range logic of the multi-level fused legs (algebraic-multigrid_b200/csrc/mid_levels.cuh), run on
the CPU through the serial host Env of tests/cpp/mid_levels_host.cpp and compared bit for bit with
the oracle's unfused sequence on every mid level (sweeps, residual, restriction; prolongation, sweeps)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "mid_levels_host.cpp")
HDR = os.path.join(ROOT, "algebraic-multigrid_b200", "csrc", "mid_levels.cuh")
LIB = os.path.join(ROOT, "tests", "cpp", "libmid_levels_host.so")
OMEGA = 2.0 / 3.0


def lib():
    if (not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(HDR))):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wall",
                               "-x", "c++", SRC, "-o", LIB])
    L = C.CDLL(LIB)
    pi = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
    pd = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    pp = C.POINTER(C.c_void_p)
    L.mid_host_run.restype = C.c_int
    L.mid_host_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, pi, pi, pi, pi, pp, pp, pp, pp, pd, pd, C.c_int,
                               C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    return L


def dia_rows(A):
    """DIA of the rows of A (explicit zeros dropped): offsets, ld, val[d * ld + row]."""
    AT = A.transpose()
    colptr, rowidx, val = AT.arrays()    # column k of A^T = row k of A
    n = AT.cols
    cols = np.repeat(np.arange(n, dtype=np.int64), np.diff(colptr))
    keep = val != 0.0
    offs = (rowidx.astype(np.int64) - cols)[keep]
    uniq = np.unique(offs)
    ld = (n + 31) // 32 * 32
    D = np.zeros((len(uniq), ld))
    D[np.searchsorted(uniq, offs), cols[keep]] = val[keep]
    return uniq.astype(np.int32), ld, D.reshape(-1).copy()


def ptrs(arrs):
    t = (C.c_void_p * len(arrs))()
    for i, a in enumerate(arrs):
        t[i] = a.ctypes.data
    return C.cast(t, C.POINTER(C.c_void_p))


@pytest.mark.parametrize("n,L,first,n_lv,nu,eps,force_T", [
    (35, 8, 0, 4, 2, 1.0, 0), (35, 8, 1, 5, 2, 1.0, 0), (64, 9, 2, 4, 1, 1.0, 0), (100, 11, 0, 6, 2, 1.0, 0),
    (100, 11, 3, 5, 3, 1.0, 0), (129, 12, 1, 8, 2, 1e-3, 0), (129, 12, 2, 6, 2, 1.0, 64), (200, 13, 3, 7, 2, 1.0, 128),
    (65, 10, 0, 8, 2, 1.0, 256)])
def test_mid_levels_down_and_up_bit_exact(n, L, first, n_lv, nu, eps, force_T):
    Lh = lib()
    mo = O.Multigrid(O.laplacian(n, eps), O.rhs(n), L, 1e-9, 1, 1, O.SMOOTHER_JACOBI, nu, OMEGA)
    mo.vcycle()                      # level-0 iterate is no longer zero: exercises the DOWN_U input
    mo2 = O.Multigrid(O.laplacian(n, eps), O.rhs(n), L, 1e-9, 1, 1, O.SMOOTHER_JACOBI, nu, OMEGA)
    mo2.vcycle()
    last = first + n_lv - 1
    assert last + 1 < L
    lv = list(range(first, last + 1))
    dias = [dia_rows(mo.A(l)) for l in lv]
    nn = np.array([mo.n_dofs(l) for l in lv], np.int32)
    nd = np.array([len(d[0]) for d in dias], np.int32)
    assert nd.max() <= 16
    off = np.zeros(len(lv) * 16, np.int32)
    for i, d in enumerate(dias):
        off[i * 16:i * 16 + len(d[0])] = d[0]
    ld = np.array([d[1] for d in dias], np.int32)
    val = [d[2] for d in dias]
    # ---- oracle: run the second cycle's down part level by level, recording what the kernels must produce
    want_tmp, want_f = {}, {}
    for l in range(0, last + 1):
        mo.smooth(l)
        want_tmp[l] = mo.u(l).copy()
        r = O.residual(mo.A(l), mo.u(l), mo.f(l))
        mo.u(l + 1)[:] = 0.0
        mo.f(l + 1)[:] = O.spmv(mo.R(l), r)
        want_f[l + 1] = mo.f(l + 1).copy()
    # ---- kernels: state = mo2 after one cycle, with the levels above `first` already processed by the oracle
    f = [np.full(mo.n_dofs(l), np.nan) for l in lv]
    f[0] = (mo2.f(first) if first == 0 else want_f[first]).copy()
    u = [np.full(mo.n_dofs(l), np.nan) for l in lv]
    if first == 0:
        u[0] = mo2.u(0).copy()
    tmp = [np.full(mo.n_dofs(l), np.nan) for l in lv]
    n_next = mo.n_dofs(last + 1)
    f_next = np.full(n_next, np.nan)
    tile, blocks = C.c_int(0), C.c_int(0)
    rc = Lh.mid_host_run(n_lv, nu, int(first == 0), OMEGA, nn, nd, off, ld, ptrs(val), ptrs(f), ptrs(u), ptrs(tmp),
                         f_next, np.zeros(n_next), n_next, 0, 0, 400000, force_T, C.byref(tile), C.byref(blocks))
    assert rc == 0
    if force_T:
        assert blocks.value > 1
    for i, l in enumerate(lv):
        assert tmp[i].tobytes() == want_tmp[l].tobytes(), ("tmp", l)
        if i > 0:
            assert f[i].tobytes() == want_f[l].tobytes(), ("f", l)
    assert f_next.tobytes() == want_f[last + 1].tobytes()
    # ---- up: give the level below an arbitrary correction and compare the up legs
    rng = np.random.default_rng(3)
    mo.u(last + 1)[:] = rng.standard_normal(n_next)
    u_next = mo.u(last + 1).copy()
    want_u = {}
    for l in range(last, first - 1, -1):
        mo.u(l)[:] = mo.u(l) + O.spmv(mo.P(l), mo.u(l + 1))
        mo.smooth(l)
        want_u[l] = mo.u(l).copy()
    rc = Lh.mid_host_run(n_lv, nu, int(first == 0), OMEGA, nn, nd, off, ld, ptrs(val), ptrs(f), ptrs(u), ptrs(tmp),
                         f_next, u_next, n_next, 0, 1, 400000, force_T, C.byref(tile), C.byref(blocks))
    assert rc == 0
    for i, l in enumerate(lv):
        assert u[i].tobytes() == want_u[l].tobytes(), ("u", l)

"""Row-wise Galerkin product on the DIA layout (algebraic-multigrid_b200/csrc/galerkin_dia.cuh,
the building block of a device-side setup) against the oracle's Eigen-order triple product
R (A P) (include/amg/multigrid.hpp:219-223): every coarse entry bit for bit, on every level."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "galerkin_dia_host.cpp")
HDR = os.path.join(ROOT, "algebraic-multigrid_b200", "csrc", "galerkin_dia.cuh")
LIB = os.path.join(ROOT, "tests", "cpp", "libgalerkin_dia_host.so")


def lib():
    if (not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(HDR))):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wall",
                               "-x", "c++", SRC, "-o", LIB])
    L = C.CDLL(LIB)
    pd = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    pi = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
    L.gal_host_coarse.restype = C.c_int
    L.gal_host_coarse.argtypes = [C.c_int, C.c_int, pi, C.c_int, pd, C.c_int, pi, C.POINTER(C.c_int), C.c_int, pd]
    return L


def rows_dia(A):
    """DIA of the ROWS of A: val[d, i] = A(i, i + off[d]), explicit zeros dropped."""
    colptr, rowidx, val = A.arrays()
    n = A.cols
    cols = np.repeat(np.arange(n, dtype=np.int64), np.diff(colptr))
    keep = val != 0.0
    rows = rowidx.astype(np.int64)[keep]
    offs = cols[keep] - rows
    uniq = np.unique(offs)
    ld = (n + 31) // 32 * 32
    D = np.zeros((len(uniq), ld))
    D[np.searchsorted(uniq, offs), rows] = val[keep]
    return uniq.astype(np.int32), ld, D


@pytest.mark.parametrize("n,eps", [(35, 1.0), (64, 1.0), (100, 1.0), (129, 1e-3)])
def test_rowwise_galerkin_matches_eigen_order_product(n, eps):
    L = lib()
    sizes = [n * n]
    while sizes[-1] > 20:
        sizes.append(O.n_H_from_n_h(sizes[-1]))
    mo = O.Multigrid(O.laplacian(n, eps), O.rhs(n), len(sizes), 1e-9, 1, 1, O.SMOOTHER_JACOBI, 2, 2.0 / 3.0)
    for l in range(len(sizes) - 1):
        off_f, ld_f, D_f = rows_dia(mo.A(l))
        n_c = sizes[l + 1]
        ld_c = (n_c + 31) // 32 * 32
        off_c = np.zeros(16, np.int32)
        nd_c = C.c_int(0)
        val_c = np.zeros(16 * ld_c)
        rc = L.gal_host_coarse(sizes[l], len(off_f), off_f, ld_f, D_f.reshape(-1).copy(), n_c, off_c,
                               C.byref(nd_c), ld_c, val_c)
        assert rc == 0
        got = {int(off_c[c]): val_c[c * ld_c:c * ld_c + n_c] for c in range(nd_c.value)}
        want_off, want_ld, want_D = rows_dia(mo.A(l + 1))
        # every diagonal of the oracle's coarse operator is produced, bit for bit ...
        for d, o in enumerate(want_off):
            assert int(o) in got, (l, o)
            assert got[int(o)].tobytes() == want_D[d, :n_c].tobytes(), (l, int(o))
        # ... and the diagonals it does not have come out as exact zeros
        for o, v in got.items():
            if o not in set(int(x) for x in want_off):
                assert not v.any(), (l, o)

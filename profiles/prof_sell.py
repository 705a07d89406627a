"""GB/s of the general-matrix (SELL-32) kernels once: the five-point operator
of a 1025 x 1025 grid under a random symmetric permutation -- thousands of distinct diagonals, so the
device mirror takes the SELL-32 layout; residual and damped-Jacobi sweep, bytes = 12 B per stored entry
+ slice pointers + 24 B per row."""
import importlib
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
amg = importlib.import_module("algebraic-multigrid_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1025
A = amg.Grid.laplacian(n)
N = n * n
perm = np.random.default_rng(0).permutation(N)
Pm = sp.csc_matrix((np.ones(N), (np.arange(N), perm)), shape=(N, N))
M = (Pm @ A.to_scipy() @ Pm.T).tocsc()
M.sort_indices()
Ap = amg.CscMatrix(N, N, M.indptr, M.indices, M.data)
dm = amg.DeviceMatrix(Ap)
b = amg.Grid.rhs(n)[perm]
dm.residual(np.ones(N), b)
for kind, name in ((2, "residual"), (0, "jacobi sweep")):
    ms = dm.time_pass(kind, 2.0 / 3.0, 3, 20)
    by = dm.stream_bytes(kind) + 24 * N
    print("SELL-32, %d rows, %d entries: %s %.4f ms, %.0f GB/s of layout bytes (%d B)" % (N, Ap.nnz, name, ms, by / ms / 1e6, by))

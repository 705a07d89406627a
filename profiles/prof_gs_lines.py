"""Small driver for ncu: a few symmetric Gauss-Seidel sweeps (line-scan kernel) on the
1025^2 five-point operator and on its level-2 Galerkin operator."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
amg = importlib.import_module("algebraic-multigrid_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1025
A, b = amg.Grid.laplacian(n), amg.Grid.rhs(n)
u = np.zeros(n * n)
sm = amg.SparseGaussSeidel()
sm.n_iters = 3
sm.smooth(A, u, b)
print("ok", float(np.abs(u).max()))

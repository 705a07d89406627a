"""Config 5: anisotropic diffusion kron(I,D) + eps kron(D,I), eps = 1e-3, on an n x n grid:
V-cycles (damped Jacobi 2+2) and V-cycle-preconditioned CG to 1e-8 relative residual.
usage: python profiles/prof_aniso.py [n] [eps]"""
import importlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
amg = importlib.import_module("algebraic-multigrid_b200")

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4097
eps = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-3
sizes = [n * n]
while sizes[-1] > 200:
    sizes.append(amg.lib().amgb_n_H_dofs_from_n_h_dofs(sizes[-1]))
L = len(sizes)
A, b = amg.Grid.laplacian(n, eps), amg.Grid.rhs(n)
mg = amg.Multigrid(None, amg.DampedJacobi(2.0 / 3.0, 2), A, b, L, 1e-9, 1, 3000)
print("n=%d eps=%g levels=%d fused legs on levels %s, tail from level %d" % (
    n, eps, L, [l for l in range(L - 1) if mg.fused_legs(l)], mg.tail_first()))
t0 = time.perf_counter()
mg.solve_relative(1e-8)
dt = time.perf_counter() - t0
print("plain V-cycles: %d cycles to rel. residual %.3e in %.3f s" % (mg.iters_done, mg.last_error, dt))
mg.set_soln(0, np.zeros(n * n))
t0 = time.perf_counter()
mg.solve_pcg(1e-8, 3000)
dt = time.perf_counter() - t0
print("PCG + V-cycle: %d iterations to rel. residual %.3e in %.3f s" % (mg.iters_done, mg.last_error, dt))

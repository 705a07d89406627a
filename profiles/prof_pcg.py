"""Time to solution on the bench problem: plain V-cycles vs CG preconditioned by the V-cycle
(amgb_solve_pcg), damped Jacobi 2+2.  usage: python profiles/prof_pcg.py [n] [rel_tol]"""
import importlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
amg = importlib.import_module("algebraic-multigrid_b200")

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1025
tol = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-8
sizes = [n * n]
while sizes[-1] > 200:
    sizes.append(amg.lib().amgb_n_H_dofs_from_n_h_dofs(sizes[-1]))
L = len(sizes)
A, b = amg.Grid.laplacian(n), amg.Grid.rhs(n)
mg = amg.Multigrid(None, amg.DampedJacobi(2.0 / 3.0, 2), A, b, L, 1e-9, 1, 1)
mg.vcycles(2)
mg.set_soln(0, np.zeros(n * n))
t0 = time.perf_counter()
mg.solve_pcg(tol, 5000)
dt = time.perf_counter() - t0
h = mg.error_history()
print("n=%d levels=%d PCG: %d iterations to rel. residual %.3e in %.3f s (%.2f ms per iteration)"
      % (n, L, mg.iters_done, mg.last_error, dt, 1e3 * dt / max(1, mg.iters_done)))
print("history every 50:", ["%.2e" % v for v in h[::50]])

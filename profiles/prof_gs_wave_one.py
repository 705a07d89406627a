"""One 1025^2 five-point operator, a few forward wavefront Gauss-Seidel sweeps: the target of the
ncu --set full capture of k_gs_wave (profiles/r2_gs_wave.md)."""
import importlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
amg = importlib.import_module("algebraic-multigrid_b200")

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1025
A, b = amg.Grid.laplacian(n), amg.Grid.rhs(n)
dm = amg.DeviceMatrix(A)
u = np.random.default_rng(1).standard_normal(n * n)
amg.rss(dm, u, b)
print("ms per forward sweep:", dm.time_pass(3, warmup=2, reps=3))

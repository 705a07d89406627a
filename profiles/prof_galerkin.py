"""Device-side Galerkin product (k_galerkin_dia) of every level of the bench hierarchy: kernel time
and bitwise comparison with the host-built coarse operators.  usage: python profiles/prof_galerkin.py [n]"""
import importlib
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
amg = importlib.import_module("algebraic-multigrid_b200")

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4097
sizes = [n * n]
while sizes[-1] > 200:
    sizes.append(amg.lib().amgb_n_H_dofs_from_n_h_dofs(sizes[-1]))
L = len(sizes)
t0 = time.perf_counter()
A, b = amg.Grid.laplacian(n), amg.Grid.rhs(n)
mg = amg.Multigrid(None, amg.DampedJacobi(2.0 / 3.0, 2), A, b, L, 1e-9, 1, 1)
print("n=%d levels=%d host setup (generator + Galerkin products + mirrors) %.1f s" % (n, L, time.perf_counter() - t0))
tot = 0.0
for l in range(L - 1):
    ms, bad = mg.galerkin_device(l)
    tot += ms
    print("level %2d -> %2d: %9d -> %9d rows, k_galerkin_dia %.3f ms, %d entries differ from the host product"
          % (l, l + 1, sizes[l], sizes[l + 1], ms, bad))
print("all levels on the device: %.2f ms" % tot)

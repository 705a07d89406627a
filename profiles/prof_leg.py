"""Times the fused V-cycle legs (k_fused_leg) against the per-operator kernels on a Poisson
hierarchy, for the tiling knobs given in the environment (AMGB_LEG_OCC / _PF / _W / _LJ).
usage: python profiles/prof_leg.py [n] [levels_to_time]"""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
amg = importlib.import_module("algebraic-multigrid_b200")

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2049
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sizes = [n * n]
while sizes[-1] > 600:
    sizes.append(amg.lib().amgb_n_H_dofs_from_n_h_dofs(sizes[-1]))
L = len(sizes)
A, b = amg.Grid.laplacian(n), amg.Grid.rhs(n)
mg = amg.Multigrid(None, amg.DampedJacobi(2.0 / 3.0, 2), A, b, L, 1e-9, 1, 1, fuse=5)
mg.vcycles(3)   # realistic data on every level (zero vectors would time the division's special-case path)
knobs = {k: v for k, v in os.environ.items() if k.startswith("AMGB_")}
print("n", n, "levels", L, knobs)
for l in range(nl):
    N, N1 = sizes[l], sizes[l + 1]
    mb = mg.matrix_bytes(l)
    if not mg.fused_legs(l):
        print("level", l, "not fused")
        continue
    sweep = mg.time_kernel(l, 0, 3, 20)
    rr = mg.time_kernel(l, 2, 3, 20)
    pa = mg.time_kernel(l, 3, 3, 20)
    down = mg.time_kernel(l, 4, 3, 20)
    up = mg.time_kernel(l, 5, 3, 20)
    bd = mb + (24 if l == 0 else 16) * N + 8 * N1
    bu = mb + 24 * N + 8 * N1
    unf_down = (2 if l == 0 else 1) * sweep + rr + (0 if l == 0 else sweep * 0.3)
    unf_up = pa + 2 * sweep
    print("level %d N=%d  down %.1f us (%.0f GB/s; unfused ~%.1f us)  up %.1f us (%.0f GB/s; unfused ~%.1f us)  plan %s"
          % (l, N, down * 1e3, bd / down / 1e6, unf_down * 1e3, up * 1e3, bu / up / 1e6, unf_up * 1e3,
             mg.leg_plan(l)))

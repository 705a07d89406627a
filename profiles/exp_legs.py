"""Leg-kernel experiments on one GPU: per-level fused-leg timings (CUDA events inside the library) and
the V-cycle time for a list of tuning-knob settings (environment variables read when a hierarchy is
created).  usage: python profiles/exp_legs.py [n] 'K1=V1,K2=V2' 'K1=V3' ..."""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
amg = importlib.import_module("algebraic-multigrid_b200")

n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 4097
settings = [a for a in sys.argv[1:] if not a.isdigit()] or [""]
sizes = [n * n]
while sizes[-1] > 200:
    sizes.append(amg.lib().amgb_n_H_dofs_from_n_h_dofs(sizes[-1]))
L = len(sizes)
A, b = amg.Grid.laplacian(n), amg.Grid.rhs(n)
KNOBS = ("AMGB_SLEG_MF_PF", "AMGB_SLEG_L2AHEAD", "AMGB_SLEG_PF", "AMGB_SLEG_WARPS_PER_SM", "AMGB_TAIL_ROWS", "AMGB_HOST_SETUP",
         "AMGB_ARITH", "AMGB_MID_ROWS", "AMGB_SLEG_MINLINES")
for setting in settings:
    for k in KNOBS:
        os.environ.pop(k, None)
    arith = amg.ARITH_FAST
    fuse = None
    for kv in filter(None, setting.split(",")):
        k, v = kv.split("=")
        if k == "AMGB_ARITH":
            arith = amg.ARITH_FAST if v == "fast" else amg.ARITH_REFERENCE
        if k == "FUSE":
            fuse = int(v)
            continue
        os.environ[k] = v
    t0 = time.perf_counter()
    mg = amg.Multigrid(None, amg.DampedJacobi(2.0 / 3.0, 2), A, b, L, 1e-9, 1, 1, arith=arith, fuse=fuse)
    setup = time.perf_counter() - t0
    for _ in range(5):
        mg.vcycle()
    mg.synchronize()
    t0 = time.perf_counter()
    mg.vcycles(50)
    cyc = (time.perf_counter() - t0) / 50 * 1e3
    legs = []
    for l in range(L - 1):
        if mg.fused_legs(l):
            legs.append("L%d %.1f/%.1f" % (l, mg.time_kernel(l, 4, 3, 20) * 1e3, mg.time_kernel(l, 5, 3, 20) * 1e3))
    extra = ""
    if mg.mid_range()[0] >= 0:
        extra = "  mid %s down %.1f up %.1f" % (mg.mid_range(), mg.time_kernel(0, 6, 3, 20) * 1e3, mg.time_kernel(0, 7, 3, 20) * 1e3)
    extra += "  tail %.1f" % (mg.time_kernel(0, 8, 3, 20) * 1e3)
    print("[%s] setup %.2f s  vcycle %.4f ms  launches %d  tail_first %d  legs(us down/up): %s" % (
        setting, setup, cyc, mg.launches_per_vcycle(), mg.tail_first(), "  ".join(legs) + extra), flush=True)
    del mg

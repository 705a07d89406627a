"""Forward lexicographic Gauss-Seidel sweep on the level-0 five-point operator: the multi-SM wavefront
kernel (AMGB_GS_AUTO, gs_wave.cuh) beside the single-SM line-scan kernel (AMGB_GS_LINESCAN).
CUDA events through amgb_matrix_time kinds 3 / 4; one JSON line per grid size."""
import importlib
import json
import sys

import numpy as np

amg = importlib.import_module("algebraic-multigrid_b200")

for n in [int(x) for x in (sys.argv[1:] or ["1025", "2049"])]:
    A = amg.Grid.laplacian(n)
    b = amg.Grid.rhs(n)
    dm = amg.DeviceMatrix(A)
    u = np.random.default_rng(1).standard_normal(n * n)
    amg.rss(dm, u, b)                       # uploads u and b
    wave = dm.time_pass(3, warmup=3, reps=10)
    amg.rss(dm, u, b)
    scan = dm.time_pass(4, warmup=3, reps=10)
    lines, S = n, 1
    blocks = (lines + 29) // 30
    print(json.dumps({"n": n, "rows": n * n, "gs_kernel_auto": dm.gs_kernel(amg.GS_AUTO),
                      "wave_ms_per_sweep": wave, "linescan_ms_per_sweep": scan, "speedup": scan / wave,
                      "wave_blocks": blocks, "wave_steps_per_block": n + 31 * S,
                      "wave_ns_per_row": wave * 1e6 / (n * n), "linescan_ns_per_row": scan * 1e6 / (n * n)}))

"""Forward lexicographic Gauss-Seidel sweep on five-point grid operators: the multi-SM wavefront kernel
(AMGB_GS_AUTO, gs_wave.cuh) beside the single-SM line-scan kernel (AMGB_GS_LINESCAN).  CUDA events
through amgb_matrix_time kinds 3 / 4; one JSON line per shape.  Shapes "LxM" are L grid lines of M
rows: 30x1025 is ONE block of the wavefront kernel (time / steps = the step time of the dependency
chain), 60x1025 adds one hand-over between blocks.  AMGB_GS_WAVE_DIV=exact switches the kernel to
the in-loop IEEE division (the default evaluates the same division in its split form)."""
import importlib
import json
import os
import sys

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
amg = importlib.import_module("algebraic-multigrid_b200")


def five_point(n_lines, m):
    I_l, I_m = sp.identity(n_lines), sp.identity(m)
    T_m = sp.diags([-1.0, 4.0, -1.0], [-1, 0, 1], shape=(m, m))
    T_l = sp.diags([-1.0, -1.0], [-1, 1], shape=(n_lines, n_lines))
    A = (sp.kron(I_l, T_m) + sp.kron(T_l, I_m)).tocsc()
    A.sort_indices()
    return amg.CscMatrix(A.shape[0], A.shape[1], A.indptr, A.indices, A.data)


for shape in (sys.argv[1:] or ["30x1025", "60x1025", "1025x1025", "2049x2049"]):
    n_lines, m = (int(x) for x in shape.split("x"))
    A = five_point(n_lines, m)
    n = n_lines * m
    rng = np.random.default_rng(1)
    b, u = rng.standard_normal(n), rng.standard_normal(n)
    out = {"shape": shape, "rows": n}
    for div in ("split", "exact"):
        os.environ["AMGB_GS_WAVE_DIV"] = div
        dm = amg.DeviceMatrix(A)
        out["gs_kernel_auto"] = dm.gs_kernel(amg.GS_AUTO)
        amg.rss(dm, u, b)                       # uploads u and b
        out["wave_%s_ms" % div] = dm.time_pass(3, warmup=3, reps=10)
        if div == "exact":
            amg.rss(dm, u, b)
            out["linescan_ms"] = dm.time_pass(4, warmup=3, reps=10)
    os.environ["AMGB_GS_WAVE_DIV"] = "split"
    os.environ["AMGB_GS_WAVE_PD"] = "2"
    dm = amg.DeviceMatrix(A)
    amg.rss(dm, u, b)
    out["wave_split_pd2_ms"] = dm.time_pass(3, warmup=3, reps=10)
    os.environ["AMGB_GS_WAVE_PD"] = "5"
    blocks = (n_lines + 29) // 30
    steps = m + 31
    out.update(wave_blocks=blocks, wave_steps_per_block=steps,
               wave_split_ns_per_step_if_one_block=out["wave_split_ms"] * 1e6 / steps,
               speedup_vs_linescan=out["linescan_ms"] / out["wave_split_ms"])
    print(json.dumps(out), flush=True)

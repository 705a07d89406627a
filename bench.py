#!/usr/bin/env python
"""bench.py -- V-cycles/s of the B200-native multigrid V-cycle path.

Workload (BASELINE.json metric / configs[2]): 2-D five-point Poisson on a
4097x4097 interior grid, fp64, 18 levels (coarsest 127 DOF), damped-Jacobi
smoother (omega 2/3, 2 pre + 2 post sweeps), device-resident V-cycles replayed
as a CUDA graph.  One "step" is one V-cycle.  Prints ONE JSON line.

  python bench.py --gpus 1 --steps 20 --warmup 3
  python bench.py --impl reference ...   # the reference algorithm on host cores

`--impl reference` times the reference's own V-cycle (symmetric Gauss-Seidel,
/root/reference/include/amg/multigrid.hpp:263-305) as restated by the CPU
oracle -- the reference cannot be compiled here (Eigen 3.4.0 absent) -- on the
same grid, single-threaded like the reference.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=100)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--n", type=int, default=4097, help="interior grid points per direction")
    p.add_argument("--levels", type=int, default=0, help="0 = coarsen until <= 200 DOF")
    p.add_argument("--smoother", default="jacobi", choices=["jacobi", "color", "gs"])
    p.add_argument("--eps", type=float, default=1.0, help="anisotropy of the +-n coupling")
    p.add_argument("--min-rows-per-rank", type=int, default=1 << 17,
                   help="levels with fewer rows per rank are agglomerated (replicated) instead of sharded")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    return p.parse_args()


def default_levels(n):
    import oracle as O  # closed-form size rule only
    sizes = [n * n]
    while sizes[-1] > 200:
        sizes.append(O.n_H_from_n_h(sizes[-1]))
    return len(sizes)


def workload_name(a):
    """Same string for both arms: the problem and the operation (one V-cycle); the smoother is a
    separate config key (the reference only has symmetric Gauss-Seidel)."""
    return "poisson2d_%dx%d_fp64_vcycle%s" % (a.n, a.n, "" if a.eps == 1.0 else "_eps%g" % a.eps)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        import datetime
        for line in self.proc.stdout:
            cells = [c.strip() for c in line.split(",")]
            try:   # nvidia-smi's own wall-clock stamp (the pipe may deliver lines late)
                at = datetime.datetime.strptime(cells[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
            except (ValueError, IndexError):
                at = time.time()
            self.rows.append([at] + cells[1:])

    def stop(self, t0=None, t1=None):
        """Median SM clock / throttle reasons of the samples that arrived in [t0, t1] (the timed
        region); nvidia-smi is started long before it because it needs about a second to come up."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r[1:] for r in self.rows if t0 is None or t0 <= r[0] <= t1 + 0.05]
        window = "timed region"
        if not rows:   # region shorter than the sampling period: the samples right around it
            near = [r[1:] for r in self.rows if t0 - 0.25 <= r[0] <= t1 + 0.25] if t0 is not None else []
            rows = near or [r[1:] for r in self.rows[-3:]]
            window = "+-0.25 s around the timed region (it is shorter than the sampling period)"
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for k, nm in enumerate(names):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "window": window, "reasons": sorted(reasons)}


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel):
    """DRAM bytes per launch of the dominant kernel from the committed ncu summary."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


def cpu_reference_vcycles(a, levels, steps, warmup):
    """The reference algorithm (symmetric Gauss-Seidel V-cycle) on one host core."""
    import oracle as O
    A = O.laplacian(a.n, a.eps)
    b = O.rhs(a.n)
    mg = O.Multigrid(A, b, levels, 1e-9, 1, 1, O.SMOOTHER_GS, 1, 2.0 / 3.0)
    for _ in range(warmup):
        mg.vcycle()
    t0 = time.perf_counter()
    for _ in range(steps):
        mg.vcycle()
    dt = time.perf_counter() - t0
    return steps / dt, dt / steps


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    levels = a.levels or default_levels(a.n)
    # bounded: a V-cycle costs ~3.8 s at 4097^2 on one core; cap the sample at ~45 s
    per = 3.8 * (a.n / 4097.0) ** 2
    steps = max(1, min(a.steps, int(30.0 / per) or 1))
    warm = 1 if per < 10 else 0
    vps, sec = cpu_reference_vcycles(a, levels, steps, warm)
    out = {
        "impl": "reference", "metric": "vcycles_per_s", "value": vps, "unit": "V-cycles/s",
        "n_gpus": a.gpus, "steps": steps, "warmup": warm, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(a), "n": a.n, "levels": levels,
                   "smoother": "symmetric Gauss-Seidel x1 (reference default)",
                   "mdof_per_s": vps * a.n * a.n / 1e6},
        "cpu_baseline": {"value": vps, "unit": "V-cycles/s", "cores": 1, "kind": "port",
                         "sample": "%d full V-cycle(s) of the oracle restatement of the reference "
                                   "(Eigen absent => reference not compilable), setup excluded, "
                                   "single-threaded like the reference" % steps},
        "e2e": {"value": vps, "unit": "V-cycles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


def run_b200(a):
    import numpy as np
    import torch
    amg = importlib.import_module("algebraic-multigrid_b200")

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    amg.lib().amgb_set_device(local)
    dist = None
    comm = None
    # Libraries (NCCL's version banner, ...) write to the C-level stdout: route everything but
    # the final JSON line to stderr
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

        def exchange_id(raw):
            box = [raw]
            dist.broadcast_object_list(box, src=0, device=torch.device("cuda", local))
            return box[0]
        comm = amg.Comm(rank, world, exchange_id)
        if a.smoother != "jacobi":
            raise SystemExit("the sharded V-cycle uses damped Jacobi (Gauss-Seidel is single-GPU)")

    levels = a.levels or default_levels(a.n)
    smoother = {"jacobi": amg.DampedJacobi(2.0 / 3.0, 2), "color": amg.MulticolorGaussSeidel(1),
                "gs": amg.SparseGaussSeidel()}[a.smoother]
    sampler = ClockSampler(local) if rank == 0 else None   # running well before the timed region
    t0 = time.perf_counter()
    A = amg.Grid.laplacian(a.n, a.eps)
    b = amg.Grid.rhs(a.n)
    mg = amg.Multigrid(None, smoother, A, b, levels, 1e-9, 1, 1, comm=comm,
                       min_rows_per_rank=a.min_rows_per_rank)
    setup_s = time.perf_counter() - t0
    N0 = mg.get_n_dofs(0)

    stream = torch.cuda.Stream()
    mg.set_stream(stream.cuda_stream)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident V-cycles (value) ----
    n_warm = max(a.warmup, 3)   # timing rule: at least three untimed cycles (graph built, clocks up)
    for _ in range(n_warm):
        mg.vcycle()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = amg.kernel_launches()
    tw0 = time.time()
    e0.record(stream)
    for _ in range(a.steps):
        mg.vcycle()
    e1.record(stream)
    barrier()
    tw1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = amg.kernel_launches() - launches0
    clocks = sampler.stop(tw0, tw1) if sampler else None
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # one problem, row blocks spread over the ranks: whole-job V-cycles = steps
    vps = a.steps / (ms * 1e-3)
    rss_after = mg.rss()

    # ---- e2e: host-resident b and u; H2D(b,u) + V-cycle + D2H(u) each step ----
    e2e = None
    if not a.no_e2e:
        hb = torch.from_numpy(b).pin_memory()
        hu = torch.zeros(N0, dtype=torch.float64).pin_memory()
        steps_e = max(3, min(a.steps, 10))
        for _ in range(2):
            mg.set_rhs(0, hb.numpy()); mg.set_soln(0, hu.numpy()); mg.vcycle(); mg.get_soln(0, hu.numpy())
        barrier()
        t1 = time.perf_counter()
        for _ in range(steps_e):
            mg.set_rhs(0, hb.numpy())
            mg.set_soln(0, hu.numpy())
            mg.vcycle()
            mg.get_soln(0, hu.numpy())      # synchronises
        barrier()
        dt = time.perf_counter() - t1
        if dist is not None:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": steps_e / dt, "unit": "V-cycles/s",
               "h2d_bytes_per_step": 16 * N0, "d2h_bytes_per_step": 8 * N0, "steps": steps_e}

    # ---- dominant kernel roofline: the level-0 smoother pass, timed live with CUDA events ----
    # Algorithmic bytes of one pass over level l for the layout the kernel streams
    # (DESIGN.md "bytes per unit"): stored matrix bytes (DIA: 8 B x diagonals x rows, no
    # index array; SELL: 12 B per stored entry) + f read + u read + result write (24 N).
    # The CSR-based figure of SURVEY.md 8(d) (12 nnz + 28 N + 4) is reported beside it.
    peak, peak_src = measured_hbm_peak()
    n1 = mg.get_n_dofs(1)
    r0, r1 = mg.local_range(0)
    survey0 = mg.pass_bytes(0)
    fused0 = mg.fused_legs(0)
    per_kernel = {}
    if fused0:
        # fused down leg of level 0 (sweeps + residual + restriction in one pass): operator,
        # f and u read once, smoothed u written once, coarse rhs written once
        kern_ms = mg.time_kernel(0, 4, warmup=3, reps=20)
        bytes0 = mg.matrix_bytes(0) + 24 * (r1 - r0) + 8 * ((r1 - r0) // 2)   # this rank's row block
        plan = mg.leg_plan(0)
        kname = "k_stream_leg" if plan["smem_bytes"] == 0 else "k_fused_leg"
        kdesc = "%s down leg (level 0: %d Jacobi sweeps + residual + restriction in one pass, %s layout)" % (
            kname, smoother.n_iters, mg.format(0))
    else:
        kern_ms = mg.time_kernel(0, 0, warmup=3, reps=20)
        bytes0 = mg.matrix_bytes(0) + 24 * (r1 - r0)       # this rank's row block
        plan = None
        kname = {"jacobi": "k_jacobi", "color": "k_color_gs", "gs": "k_gs_fronts"}[a.smoother]
        kdesc = kname + " (level 0, %s layout)" % mg.format(0)
    achieved = bytes0 / (kern_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": kdesc,
                "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes0, "ms_per_launch": kern_ms,
                "survey_formula_bytes_per_launch": survey0,
                "survey_formula_GBps": survey0 / (kern_ms * 1e-3) / 1e9,
                "traffic": ncu_traffic(kname)}
    if plan:
        roofline["tiling"] = plan
    pass0 = mg.matrix_bytes(0) + 24 * (r1 - r0)
    kinds = [(0, "smoother_sweep"), (1, "residual"), (2, "residual_restrict"), (3, "prolong_add")]
    if fused0:
        kinds += [(4, "fused_down_leg"), (5, "fused_up_leg")]
    for kind, nm in kinds if world == 1 else ():
        t_ms = mg.time_kernel(0, kind, warmup=3, reps=20)
        alg = {0: pass0, 1: pass0, 2: mg.matrix_bytes(0) + 16 * N0 + 8 * n1, 3: 8 * n1 + 16 * N0,
               4: mg.matrix_bytes(0) + 24 * N0 + 8 * n1, 5: mg.matrix_bytes(0) + 24 * N0 + 8 * n1}[kind]
        per_kernel[nm] = {"ms": t_ms, "GB/s": alg / (t_ms * 1e-3) / 1e9,
                          "frac": alg / (t_ms * 1e-3) / 1e9 / peak}
    # bytes one V-cycle must move with the layouts and kernels in use
    passes = 2 * smoother.n_iters if a.smoother == "jacobi" else 4 * smoother.n_iters
    layout_bytes = 0
    for l in range(levels - 1):
        nl, nn = mg.get_n_dofs(l), mg.get_n_dofs(l + 1)
        if mg.fused_legs(l):
            down = mg.matrix_bytes(l) + (24 if l == 0 else 16) * nl + 8 * nn
            up = mg.matrix_bytes(l) + 24 * nl + 8 * nn
            layout_bytes += down + up
        else:
            layout_bytes += (passes + 1) * (mg.matrix_bytes(l) + 24 * nl) + 24 * nl + 16 * nn
    vbytes = mg.vcycle_bytes()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    cpu = None
    if not a.no_cpu_baseline:
        per = 3.8 * (a.n / 4097.0) ** 2
        cs = max(1, min(5, int(12.0 / per) or 1))
        cvps, _ = cpu_reference_vcycles(a, levels, cs, 0)
        cpu = {"value": cvps, "unit": "V-cycles/s", "cores": 1, "kind": "port",
               "host_cores_available": os.cpu_count(),
               "sample": "%d full reference V-cycle(s) (symmetric Gauss-Seidel, the reference's "
                         "smoother) of the oracle on the same %dx%d grid, setup excluded; the "
                         "reference is single-threaded by construction" % (cs, a.n, a.n)}

    out = {
        "metric": "vcycles_per_s", "value": vps, "unit": "V-cycles/s", "n_gpus": world,
        "steps": a.steps, "warmup": n_warm, "ms_per_step": ms / a.steps,
        "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "n": a.n, "n_dofs": N0, "levels": levels,
                   "smoother": a.smoother, "smoother_iters": smoother.n_iters,
                   "omega": getattr(smoother, "omega", None),
                   "l2_policy": "inputs larger than L2 (level-0 operator+vectors stream %.2f GB per "
                                "pass, L2 is 126 MB)" % (bytes0 / 1e9),
                   "parallelism": "single GPU" if world == 1 else
                                  "row blocks x%d, %d sharded levels, halo exchange via %s (%d exchanges per "
                                  "V-cycle), coarser levels replicated after one gather" % (
                                      world, mg.n_sharded_levels(),
                                      {"peer": "peer-memory writes over NVLink + epoch flags",
                                       "nccl": "NCCL send/recv"}.get(mg.halo_mode(), mg.halo_mode()),
                                      mg.halo_exchanges_per_vcycle()),
                   "mdof_per_s": vps * N0 / 1e6, "setup_s": setup_s,
                   "rss_after_timed_cycles": rss_after,
                   "fused_legs": [bool(mg.fused_legs(l)) for l in range(levels - 1)], "tail_first": mg.tail_first(),
                   "vcycle_layout_bytes": layout_bytes,
                   "vcycle_hbm_frac": (layout_bytes / (ms / a.steps * 1e-3) / 1e9 / peak) if world == 1 else None,
                   "vcycle_survey_formula_bytes": vbytes,
                   "layouts": [mg.format(l) for l in range(levels)],
                   "kernels_level0": per_kernel},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
        "gpu_launches": launches, "clocks": clocks,
    }
    os.write(json_fd, (json.dumps(out) + "\n").encode())
    if dist is not None:
        dist.destroy_process_group()


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()

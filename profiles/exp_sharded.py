"""Sharded-cycle diagnostics (launch with torch.distributed.run, one rank per GPU): per-phase times of
the V-cycle on every rank, per-level leg times with the fused halo push, cycle time.
usage: torchrun ... profiles/exp_sharded.py [n] [min_rows_per_rank] 'K=V,...' ..."""
import importlib
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
amg = importlib.import_module("algebraic-multigrid_b200")

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
amg.lib().amgb_set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def ex(raw):
    box = [raw]
    dist.broadcast_object_list(box, src=0, device=torch.device("cuda", local))
    return box[0]


comm = amg.Comm(rank, world, ex)
nums = [a for a in sys.argv[1:] if a.isdigit()]
n = int(nums[0]) if nums else 4097
min_rows = int(nums[1]) if len(nums) > 1 else 1 << 17
settings = [a for a in sys.argv[1:] if not a.isdigit()] or [""]
sizes = [n * n]
while sizes[-1] > 200:
    sizes.append(amg.lib().amgb_n_H_dofs_from_n_h_dofs(sizes[-1]))
L = len(sizes)
A, b = amg.Grid.laplacian(n), amg.Grid.rhs(n)
KNOBS = set()
for setting in settings:
    for k in KNOBS:
        os.environ.pop(k, None)
    mr = min_rows
    for kv in filter(None, setting.split(",")):
        k, v = kv.split("=")
        if k == "MIN_ROWS":
            mr = int(v)
            continue
        KNOBS.add(k)
        os.environ[k] = v
    mg = amg.Multigrid(None, amg.DampedJacobi(2.0 / 3.0, 2), A, b, L, 1e-9, 1, 1, comm=comm, min_rows_per_rank=mr,
                       arith=amg.ARITH_FAST)
    for _ in range(5):
        mg.vcycle()
    mg.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    mg.vcycles(50)
    dist.barrier()
    cyc = (time.perf_counter() - t0) / 50 * 1e3
    ph = mg.phase_times(10)
    ns = mg.n_sharded_levels()
    legs = ["L%d %.1f/%.1f" % (l, mg.time_kernel(l, 4, 3, 20) * 1e3, mg.time_kernel(l, 5, 3, 20) * 1e3) for l in range(ns)]
    line = "[%s] rank %d: sharded levels %d, vcycle %.4f ms, phases(us) %s, legs(us) %s" % (
        setting, rank, ns, cyc, {k: round(v * 1e3, 1) for k, v in ph.items()}, " ".join(legs))
    for r in range(world):
        if r == rank:
            print(line, flush=True)
        dist.barrier()
    del mg
dist.barrier()
del comm
dist.destroy_process_group()

"""Per-rank timings of the sharded smoother / residual microbenchmark kernels (torchrun)."""
import importlib
import os
import sys

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
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4097
sizes = [n * n]
while sizes[-1] > 200:
    sizes.append(amg.lib().amgb_n_H_dofs_from_n_h_dofs(sizes[-1]))
A, b = amg.Grid.laplacian(n), amg.Grid.rhs(n)
mg = amg.Multigrid(None, amg.DampedJacobi(2.0 / 3.0, 1), A, b, len(sizes), 1e-9, 1, 1, comm=comm, min_rows_per_rank=1 << 17)
mg.set_soln(0, 1e-3 * b + 1.0)
for rep in range(3):
    for kind in (9, 11):
        dist.barrier()
        ms = mg.time_kernel(0, kind, 3, 20)
        for r in range(world):
            if r == rank:
                print("rep %d kind %d rank %d: %.4f ms" % (rep, kind, rank, ms), flush=True)
            dist.barrier()
del mg
dist.barrier()
del comm
dist.destroy_process_group()

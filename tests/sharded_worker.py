"""Worker for the multi-GPU parity test (launched with torch.distributed.run, one rank per
GPU).  Each rank builds the row-block sharded hierarchy, runs V-cycles and compares the
gathered iterates with (a) the same hierarchy on one GPU -- damped Jacobi is independent of
the partition, so the bits must be identical -- and (b) the CPU oracle."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402

amg = importlib.import_module("algebraic-multigrid_b200")


def id_exchanger(rank):
    def ex(raw):
        box = [raw]
        dist.broadcast_object_list(box, src=0)
        return box[0]
    return ex


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    amg.lib().amgb_set_device(local)
    dist.init_process_group("gloo")
    comm = amg.Comm(rank, world, id_exchanger(rank))
    # (n, levels, eps, min rows per rank, sweeps (negative: multicolour Gauss-Seidel with that many
    #  symmetric sweeps), fuse option or None = default, arithmetic)
    cases = [(257, 10, 1.0, 1 << 10, 2, None, amg.ARITH_REFERENCE),
             (257, 10, 1.0, 1 << 10, -1, None, amg.ARITH_REFERENCE),    # multicolour GS: one exchange per colour pass
             (129, 9, 1e-3, 900, -2, None, amg.ARITH_REFERENCE),
             (513, 12, 1e-3, 1 << 12, 2, None, amg.ARITH_REFERENCE),
             (513, 12, 1.0, 1 << 12, 2, None, amg.ARITH_FAST),
             (257, 10, 1.0, 1 << 10, 1, None, amg.ARITH_REFERENCE),     # one sweep per smooth call
             (257, 10, 1.0, 3000, 3, None, amg.ARITH_REFERENCE),        # odd count, per-operator kernels, uneven blocks
             (257, 10, 1.0, 1 << 10, 2, 0, amg.ARITH_REFERENCE),        # nothing fused: one exchange per sweep
             (129, 9, 1.0, 700, 1, 0, amg.ARITH_REFERENCE),
             (513, 12, 1.0, 1 << 12, 2, 1 | 4 | 16 | 32 | 64 | 128, amg.ARITH_REFERENCE)]   # opt-in compressed operator formats
    if len(sys.argv) > 1 and sys.argv[1] == "big":
        cases = [(2049, 15, 1.0, 1 << 16, 2, None, amg.ARITH_REFERENCE), (2049, 15, 1.0, 1 << 16, 2, None, amg.ARITH_FAST)]
    want_mode = os.environ.get("AMGB_HALO", "peer")
    fused_push = want_mode == "peer" and os.environ.get("AMGB_FUSED_PUSH", "1") != "0"
    for n, L, eps, min_rows, nu, fuse, arith in cases:
        A, b = amg.Grid.laplacian(n, eps), amg.Grid.rhs(n)
        color = nu < 0
        nu = abs(nu)
        sm = amg.MulticolorGaussSeidel(nu) if color else amg.DampedJacobi(2.0 / 3.0, nu)
        for use_graph in (False, True):
            mg = amg.Multigrid(None, sm, A, b, L, 1e-9, 1, 1, comm=comm, min_rows_per_rank=min_rows,
                               use_graph=use_graph, fuse=fuse, arith=arith)
            ns = mg.n_sharded_levels()
            assert ns >= 1, ns
            legs = (fuse is None or fuse & 4) and nu in (1, 2) and not color
            if legs:  # the sharded levels run as fused legs on the rank's window (block + ghost rows)
                assert all(mg.fused_legs(l) for l in range(ns)), [mg.fused_legs(l) for l in range(L - 1)]
            assert mg.halo_mode() == want_mode, (mg.halo_mode(), want_mode)
            single = amg.Multigrid(None, sm, A, b, L, 1e-9, 1, 1, fuse=fuse, arith=arith)
            for _ in range(3):
                mg.vcycle()
                single.vcycle()
            for l in range(L):
                got, want = mg.get_soln(l), single.get_soln(l)
                if arith == amg.ARITH_REFERENCE:   # Jacobi is independent of the partition: same bits
                    assert got.tobytes() == want.tobytes(), (n, l, use_graph, nu, fuse, np.abs(got - want).max())
                else:   # fast arithmetic: a level may run in a different kernel (leg vs mid) than on one GPU
                    assert np.linalg.norm(got - want) <= 1e-13 * np.linalg.norm(want), (n, l, use_graph)
            assert not mg.halo_timed_out()
            if legs and fused_push:   # the legs push their boundary rows themselves: no exchange launches
                assert mg.halo_exchanges_per_vcycle() == 0, mg.halo_exchanges_per_vcycle()
            r_sh, r_one = mg.rss(), single.rss()
            assert abs(r_sh - r_one) <= 1e-13 * r_one, (r_sh, r_one)
            # row-block getters / setters: this rank's rows only, and a cycle from the state they set
            b0, b1 = mg.local_range(0)
            same = (lambda x, y: x.tobytes() == y.tobytes()) if arith == amg.ARITH_REFERENCE else (
                lambda x, y: np.linalg.norm(x - y) <= 1e-13 * np.linalg.norm(y))
            assert same(mg.get_soln_local(0), single.get_soln(0)[b0:b1])
            rng = np.random.default_rng(5)
            u_new, f_new = rng.standard_normal(n * n), rng.standard_normal(n * n)
            mg.set_soln_local(0, u_new[b0:b1]); mg.set_rhs_local(0, f_new[b0:b1])
            single.set_soln(0, u_new); single.set_rhs(0, f_new)
            mg.vcycle(); single.vcycle()
            assert same(mg.get_soln_local(0), single.get_soln(0)[b0:b1])
            assert mg.get_rhs_local(0).tobytes() == f_new[b0:b1].tobytes()
            mg.set_soln(0, np.zeros(n * n)); mg.set_rhs(0, b)
            if rank == 0 and n <= 600 and not use_graph and arith == amg.ARITH_REFERENCE:
                mo = O.Multigrid(O.laplacian(n, eps), b, L, 1e-9, 1, 1,
                                 O.SMOOTHER_COLOR_GS if color else O.SMOOTHER_JACOBI, nu, 2.0 / 3.0)
                for _ in range(3):
                    mo.vcycle()
                    mg.vcycle()
                assert np.linalg.norm(mg.get_soln(0) - mo.u(0)) <= 1e-12 * np.linalg.norm(mo.u(0))
            else:
                for _ in range(3):
                    mg.vcycle()
                mg.get_soln(0)  # collective: every rank takes part
            if rank == 0:
                print("ok %s n=%d levels=%d sharded_levels=%d nu=%d fuse=%s arith=%d graph=%s halo=%s exchanges/vcycle=%d rss=%.6e" % (
                    "color" if color else "jacobi", n, L, ns, nu, fuse, arith, use_graph, mg.halo_mode(),
                    mg.halo_exchanges_per_vcycle(), r_sh), flush=True)
            del mg, single
    dist.barrier()
    del comm
    dist.destroy_process_group()
    if rank == 0:
        print("SHARDED PARITY OK", flush=True)


if __name__ == "__main__":
    main()

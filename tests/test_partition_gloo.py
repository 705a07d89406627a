"""CPU, world_size 2 and 3 over gloo: the row-block partition plan libamgb.so uses for the
sharded V-cycle (block boundaries, halo widths, ghost rows, agglomeration boundary) is
exercised by an emulation in which every rank only ever touches its own block + halos
(everything else is NaN), exchanges halos with its neighbours through torch.distributed
and must reproduce the oracle's global damped-Jacobi V-cycle."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402

amg = importlib.import_module("algebraic-multigrid_b200")

N_GRID, LEVELS, MIN_ROWS, OMEGA, SWEEPS = 65, 9, 200, 2.0 / 3.0, 2


def half_bandwidths(mats):
    out = []
    for M in mats:
        coo = M.tocoo()
        nz = coo.data != 0.0
        out.append(int(np.abs(coo.row[nz] - coo.col[nz]).max()))
    return out


def sendrecv(rank, world, vec, lo_send, lo_recv, hi_send, hi_recv):
    """Exchange with rank-1 (lo) and rank+1 (hi); all arguments are slices of vec."""
    reqs, bufs = [], []
    if rank > 0:
        reqs.append(dist.isend(torch.from_numpy(vec[lo_send].copy()), rank - 1))
        t = torch.empty(lo_recv.stop - lo_recv.start, dtype=torch.float64)
        reqs.append(dist.irecv(t, rank - 1)); bufs.append((lo_recv, t))
    if rank + 1 < world:
        reqs.append(dist.isend(torch.from_numpy(vec[hi_send].copy()), rank + 1))
        t = torch.empty(hi_recv.stop - hi_recv.start, dtype=torch.float64)
        reqs.append(dist.irecv(t, rank + 1)); bufs.append((hi_recv, t))
    for r in reqs:
        r.wait()
    for sl, t in bufs:
        vec[sl] = t.numpy()


def worker(rank, world, port, result):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        Ao, b = O.laplacian(N_GRID), O.rhs(N_GRID)
        mo = O.Multigrid(Ao, b, LEVELS, 1e-9, 1, 1, O.SMOOTHER_JACOBI, SWEEPS, OMEGA)
        L = LEVELS
        mats = [mo.A(l).to_scipy().tocsr() for l in range(L)]
        Rs = [mo.R(l).to_scipy().tocsr() for l in range(L - 1)]
        Ps = [mo.P(l).to_scipy().tocsr() for l in range(L - 1)]
        sizes = [m.shape[0] for m in mats]
        ns, starts, hlo, hhi, ghost = amg.partition_plan(sizes, half_bandwidths(mats), world, MIN_ROWS)
        assert ns >= 2
        s = [int(starts[l][rank]) for l in range(ns)]
        e = [int(starts[l][rank + 1]) for l in range(ns)]
        for l in range(ns):   # structural properties of the plan
            assert s[l] % 2 == 0 and (l + 1 >= ns or s[l + 1] == s[l] // 2)
            assert e[l] - s[l] >= hlo[l] + hhi[l]

        def window(l):  # index range this rank may touch on level l
            return max(0, s[l] - hlo[l]), min(sizes[l], e[l] + hhi[l])

        u = [np.full(sizes[l], np.nan) if l < ns else np.zeros(sizes[l]) for l in range(L)]
        f = [np.full(sizes[l], np.nan) if l < ns else np.zeros(sizes[l]) for l in range(L)]
        lo, hi = window(0)
        u[0][lo:hi] = 0.0
        f[0][s[0]:min(sizes[0], e[0] + ghost[0])] = b[s[0]:min(sizes[0], e[0] + ghost[0])]

        def exchange(l, v):
            up, dn = hhi[l], hlo[l]
            sendrecv(rank, world, v,
                     slice(s[l], s[l] + up), slice(max(0, s[l] - hlo[l]), s[l]),
                     slice(e[l] - dn, e[l]), slice(e[l], min(sizes[l], e[l] + hhi[l])))

        def smooth(l):
            d = mats[l].diagonal()
            for _ in range(SWEEPS):
                if l < ns:
                    exchange(l, u[l])
                    rows = slice(s[l], e[l])
                    new = u[l].copy()
                    new[rows] = u[l][rows] + OMEGA * (f[l][rows] - mats[l][rows] @ np.nan_to_num(u[l], nan=np.inf)) / d[rows]
                    u[l] = new
                else:
                    u[l] = u[l] + OMEGA * (f[l] - mats[l] @ u[l]) / d

        for l in range(L - 1):
            smooth(l)
            if l < ns:
                exchange(l, u[l])
                top = min(sizes[l], e[l] + ghost[l])
                r = np.full(sizes[l], np.nan)
                r[s[l]:top] = f[l][s[l]:top] - mats[l][s[l]:top] @ np.nan_to_num(u[l], nan=np.inf)
                if l + 1 < ns:
                    c0, c1 = s[l + 1], min(sizes[l + 1], e[l + 1] + ghost[l + 1])
                else:
                    c0, c1 = s[l] // 2, (e[l] // 2 if rank + 1 < world else sizes[l + 1])
                fc = Rs[l][c0:c1] @ np.nan_to_num(r, nan=np.inf)
                assert np.isfinite(fc).all(), "restriction touched rows outside block+ghost"
                if l + 1 < ns:
                    f[l + 1][c0:c1] = fc
                    lo, hi = window(l + 1)
                    u[l + 1][lo:hi] = 0.0
                else:  # agglomeration boundary: gather the blocks, every rank keeps a replica
                    parts = [None] * world
                    dist.all_gather_object(parts, (c0, fc))
                    for c, blk in parts:
                        f[l + 1][c:c + len(blk)] = blk
                    u[l + 1][:] = 0.0
            else:
                f[l + 1] = Rs[l] @ (f[l] - mats[l] @ u[l])
                u[l + 1][:] = 0.0
        u[L - 1] = O.Ldlt(mo.A(L - 1)).solve(f[L - 1])
        for l in range(L - 2, -1, -1):
            if l + 1 < ns:
                exchange(l + 1, u[l + 1])
            if l < ns:
                corr = Ps[l][s[l]:e[l]] @ np.nan_to_num(u[l + 1], nan=np.inf)
                assert np.isfinite(corr).all(), "prolongation touched rows outside block+halo"
                u[l][s[l]:e[l]] += corr
            else:
                u[l] = u[l] + Ps[l] @ u[l + 1]
            smooth(l)
        mo.vcycle()
        for l in range(L):
            mine = u[l][s[l]:e[l]] if l < ns else u[l]
            ref = mo.u(l)[s[l]:e[l]] if l < ns else mo.u(l)
            assert np.isfinite(mine).all()
            err = np.linalg.norm(mine - ref) / max(np.linalg.norm(ref), 1e-300)
            assert err <= 1e-12, (l, err)
        result[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_partition_plan_emulated_vcycle(world):
    mgr = mp.Manager()
    result = mgr.dict()
    mp.spawn(worker, args=(world, 29620 + world, result), nprocs=world, join=True)
    assert sorted(result.keys()) == list(range(world))


def test_plan_shape_for_the_bench_config():
    sizes = O.level_sizes(4097 * 4097, 16)
    bw = [4097, 2049, 1025, 513, 257, 129, 65, 33, 17, 9, 5, 3, 2, 2, 2, 2]
    for world in (2, 4, 8):
        ns, starts, lo, hi, gh = amg.partition_plan(sizes, bw, world, 1 << 18)
        assert ns >= 3
        assert list(gh) == [2 ** (ns - l) - 1 for l in range(ns)]
        for l in range(ns):
            assert starts[l][0] == 0 and starts[l][world] == sizes[l]
            blocks = np.diff(starts[l])
            assert blocks.min() > 0 and blocks.max() - blocks.min() <= 2 ** ns + sizes[l] % world + 1
            assert (starts[l][:-1] % 2 == 0).all()
            # ghost rows for the fused legs: three chained stencil stages + the restriction's
            # neighbours (3 w + 4), on both sides, and every block at least as long as what it sends
            assert lo[l] >= 3 * bw[l] + 4 and hi[l] >= 3 * bw[l] + 4 + gh[l]
            assert blocks.min() >= max(lo[l], hi[l])
    ns1, *_ = amg.partition_plan(sizes, bw, 1, 1 << 18)
    assert ns1 == 0


def _window_leg_down(M, f, u, nu, lo, hi):
    """The fused down leg as a rank sees it: only rows [lo, hi) of the operator, f and u exist
    (everything else contributes zeros, like the kernel's out-of-window lanes); returns the
    smoothed iterate and the residual on the window."""
    n = M.shape[0]
    W = M[lo:hi, lo:hi].tocsr()
    d = W.diagonal()
    uw, fw = u[lo:hi].copy(), f[lo:hi]
    for _ in range(nu):
        r = fw - W @ uw
        uw = uw + OMEGA * (r / d)
    return uw, fw - W @ uw


@pytest.mark.parametrize("world", [2, 3, 4])
def test_ghost_rows_cover_the_fused_legs(world):
    """The plan's ghost rows are wide enough: a leg computed on a rank's window (block + ghost
    rows, nothing else) gives, on the OWNED rows, what the leg on the whole level gives -- two
    sweeps + residual for the down leg (and its restriction), prolongation + two sweeps for the
    up leg.  Same floating-point operation count per row, so the comparison is to rounding
    (the CUDA kernels are compared bit for bit on the GPU, tests/sharded_worker.py)."""
    n = 129
    Ao, b = O.laplacian(n), O.rhs(n)
    sizes = O.level_sizes(n * n, 8)
    mo = O.Multigrid(Ao, b, 8, 1e-9, 1, 1, O.SMOOTHER_JACOBI, SWEEPS, OMEGA)
    mats = [mo.A(l).to_scipy().tocsr() for l in range(8)]
    bw = half_bandwidths(mats)
    ns, starts, hlo, hhi, ghost = amg.partition_plan(sizes, bw, world, 500)
    assert ns >= 2
    rng = np.random.default_rng(5)
    for l in range(ns):
        M, N = mats[l], sizes[l]
        f, u = rng.standard_normal(N), rng.standard_normal(N)
        # whole-level reference
        uw, rw = _window_leg_down(M, f, u, SWEEPS, 0, N)
        for g in range(world):
            s, e = int(starts[l][g]), int(starts[l][g + 1])
            lo, hi = max(0, s - int(hlo[l])), min(N, e + int(hhi[l]))
            ug, rg = _window_leg_down(M, f, u, SWEEPS, lo, hi)
            np.testing.assert_allclose(ug[s - lo:e - lo], uw[s:e], rtol=0, atol=1e-13 * np.abs(uw).max())
            # residual on the owned rows and one row either side (the restriction reads r[k-1], r[k+1])
            a, z = max(s - 1, 0), min(e + 1, N)
            np.testing.assert_allclose(rg[a - lo:z - lo], rw[a:z], rtol=0, atol=1e-12 * np.abs(rw).max())

"""GPU parity tests: every kernel of the V-cycle path, called through the C ABI
(libamgb.so via ctypes), against the CPU oracle on the same deterministic
inputs.  Tolerance: the north star allows 1e-12 relative in fp64; kernels that
keep the reference's summation order are additionally required to be
bit-identical to the oracle.  Integer maps must be bit-exact."""
import importlib

import numpy as np
import pytest

import oracle as O

amg = importlib.import_module("algebraic-multigrid_b200")
pytestmark = pytest.mark.gpu

RTOL = 1e-12  # BASELINE.json north_star: "within 1e-12 relative in fp64"


def rel(a, b):
    d = np.linalg.norm(a - b)
    n = max(np.linalg.norm(b), 1e-300)
    return d / n


def problem(n, eps=1.0):
    A = amg.Grid.laplacian(n, eps)
    b = amg.Grid.rhs(n)
    return A, b, O.laplacian(n, eps)


def vec(n, seed):
    return np.random.default_rng(seed).standard_normal(n)


# ----------------------------------------------------------------- operators
@pytest.mark.parametrize("n,eps", [(2, 1.0), (35, 1.0), (64, 1.0), (33, 1e-3), (200, 1.0)])
def test_residual_bit_exact(n, eps):
    A, b, Ao = problem(n, eps)
    u = vec(n * n, 1)
    r = amg.DeviceMatrix(A).residual(u, b)
    assert r.tobytes() == O.residual(Ao, u, b).tobytes()


@pytest.mark.parametrize("n", [2, 35, 200])
def test_rss(n):
    A, b, Ao = problem(n)
    u = vec(n * n, 2) * 1e-3
    got = amg.rss(A, u, b)
    want = O.rss(Ao, u, b)
    assert abs(got - want) <= 1e-13 * want


@pytest.mark.parametrize("n,eps,sweeps", [(35, 1.0, 1), (35, 1.0, 2), (64, 1e-3, 3), (200, 1.0, 2)])
def test_jacobi_bit_exact(n, eps, sweeps):
    A, b, Ao = problem(n, eps)
    AT = Ao.transpose()
    u = vec(n * n, 3)
    want = u.copy()
    for _ in range(sweeps):
        want = O.jacobi_sweep(AT, want, b, 0.6)
    got = u.copy()
    amg.DampedJacobi(0.6, sweeps).smooth(A, got, b)
    assert got.tobytes() == want.tobytes()


@pytest.mark.parametrize("n,eps", [(35, 1.0), (64, 1.0), (33, 1e-3)])
def test_multicolor_gs_bit_exact(n, eps):
    A, b, Ao = problem(n, eps)
    AT = Ao.transpose()
    nc_o, color_o = O.greedy_coloring(Ao, AT)
    nc, color = amg.DeviceMatrix(A).coloring()
    assert nc == nc_o == 2
    np.testing.assert_array_equal(color, color_o)            # integer map: bit-exact
    u = vec(n * n, 4)
    want = u.copy()
    for _ in range(2):
        for c in list(range(nc_o)) + list(range(nc_o - 1, -1, -1)):
            O.color_gs_pass(AT, color_o, c, b, want)
    got = u.copy()
    amg.MulticolorGaussSeidel(2).smooth(A, got, b)
    assert got.tobytes() == want.tobytes()


@pytest.mark.parametrize("n,eps", [(2, 1.0), (35, 1.0), (64, 1.0), (33, 1e-3)])
def test_gauss_seidel_levelsched_bit_exact(n, eps):
    """Lexicographic forward+backward sweep == reference order (smoother.hpp:148-174),
    level-scheduled kernel: identical bits."""
    A, b, Ao = problem(n, eps)
    u = vec(n * n, 5)
    want = u.copy()
    O.gs_smooth(Ao, want, b, 1e-9, 0, 3)
    got = u.copy()
    sm = amg.SparseGaussSeidel(mode=amg.GS_LEVELSCHED)
    sm.n_iters = 3
    sm.smooth(A, got, b)
    assert got.tobytes() == want.tobytes()


@pytest.mark.parametrize("n,eps", [(2, 1.0), (3, 1.0), (35, 1.0), (64, 1.0), (33, 1e-3), (200, 1.0), (513, 1.0)])
def test_gauss_seidel_linescan(n, eps):
    """Banded line-scan kernel: same update order, the distance-1 chain is a parallel scan, so
    agreement is to rounding rather than bit for bit."""
    A, b, Ao = problem(n, eps)
    u = vec(n * n, 5)
    want = u.copy()
    O.gs_smooth(Ao, want, b, 1e-9, 0, 3)
    got = u.copy()
    sm = amg.SparseGaussSeidel(mode=amg.GS_LINESCAN)
    sm.n_iters = 3
    sm.smooth(A, got, b)
    assert rel(got, want) <= 1e-14
    assert np.abs(got - want).max() <= 1e-13 * np.abs(want).max()


@pytest.mark.parametrize("n,eps", [(3, 1.0), (4, 1.0), (31, 1.0), (35, 1.0), (64, 1.0), (33, 1e-3), (100, 1.0),
                                   (257, 1e-2), (513, 1.0)])
def test_gauss_seidel_wavefront_bit_exact(n, eps):
    """Default mode on a grid operator (level 0): the multi-SM wavefront kernel (gs_wave.cuh) -- one to 18
    blocks of 30 grid lines chained through the hand-over buffer -- reproduces the reference's sweeps
    (smoother.hpp:148-174) bit for bit, forward and backward, several iterations."""
    A, b, Ao = problem(n, eps)
    dm = amg.DeviceMatrix(A)
    assert dm.gs_kernel(amg.GS_AUTO) == amg.GS_KERNEL_WAVE
    assert dm.gs_kernel(amg.GS_LINESCAN) == amg.GS_KERNEL_LINESCAN
    assert dm.gs_kernel(amg.GS_LEVELSCHED) == amg.GS_KERNEL_FRONTS
    u = vec(n * n, 5)
    want = u.copy()
    O.gs_smooth(Ao, want, b, 1e-9, 0, 3)
    got = u.copy()
    sm = amg.SparseGaussSeidel()
    sm.n_iters = 3
    sm.smooth(dm, got, b)
    assert got.tobytes() == want.tobytes()


def test_wavefront_split_division_is_the_ieee_division():
    """The wavefront kernel evaluates (b - sigma) / a_ii in the split form of the compiler's division;
    the device self-test compares it with __ddiv_rn on 2^25 operand pairs (every 16th pair with
    exponents over the whole range, denormals and specials included): not one bit may differ."""
    assert amg.selftest_division(1 << 25, seed=12345) == 0
    assert amg.selftest_division(1 << 22, seed=7) == 0


@pytest.mark.parametrize("div,pd", [("split", "5"), ("exact", "5"), ("split", "2"), ("exact", "2")])
def test_gauss_seidel_wavefront_division_variants(div, pd, monkeypatch):
    """Both division variants and both look-ahead depths of the wavefront kernel give the oracle's bits
    (AMGB_GS_WAVE_DIV, AMGB_GS_WAVE_PD)."""
    monkeypatch.setenv("AMGB_GS_WAVE_DIV", div)
    monkeypatch.setenv("AMGB_GS_WAVE_PD", pd)
    n = 70
    A, b, Ao = problem(n, 1e-1)
    u = vec(n * n, 21)
    want = u.copy()
    O.gs_smooth(Ao, want, b, 1e-9, 0, 2)
    got = u.copy()
    sm = amg.SparseGaussSeidel()
    sm.n_iters = 2
    sm.smooth(amg.DeviceMatrix(A), got, b)
    assert got.tobytes() == want.tobytes()


def test_gauss_seidel_wavefront_unsymmetric_and_pruned():
    """The wavefront kernel on an operator that is not symmetric and has pruned entries and a zero
    diagonal entry (that row is left alone, smoother.hpp:136): still the reference's bits.  The sweep
    works on "row c = CSC column c", so the oracle gets the very same CSC arrays."""
    n = 48
    A, b, _ = problem(n)
    rng = np.random.default_rng(11)
    cols = np.repeat(np.arange(n * n), np.diff(A.colptr))
    off = A.rowidx != cols
    A.val[off] *= 0.5 + rng.random(int(off.sum()))           # unsymmetric
    A.val[off & (rng.random(A.val.size) < 0.05)] = 0.0       # stored zeros: pruned from the mirror
    A.val[np.flatnonzero(~off)[777]] = 0.0                   # one zero diagonal entry
    Ao = O.Csc.from_arrays(A.rows, A.cols, A.colptr, A.rowidx, A.val)
    dm = amg.DeviceMatrix(A)
    assert dm.gs_kernel(amg.GS_AUTO) == amg.GS_KERNEL_WAVE
    u = vec(n * n, 9)
    want = u.copy()
    O.gs_smooth(Ao, want, b, 1e-9, 0, 2)
    got = u.copy()
    sm = amg.SparseGaussSeidel()
    sm.n_iters = 2
    sm.smooth(dm, got, b)
    assert got.tobytes() == want.tobytes()


def test_galerkin_levels_keep_the_scan_kernel():
    """The Galerkin operators couple across the ends of grid lines (1-D interpolation of the flattened
    vector): AUTO must not pick the wavefront kernel for them."""
    mo = O.Multigrid(O.laplacian(65), O.rhs(65), 4, 1e-9, 1, 1)
    for l in (1, 2, 3):
        Al = mo.A(l)
        c, r, v = Al.arrays()
        dm = amg.DeviceMatrix(amg.CscMatrix(Al.rows, Al.cols, c, r, v))
        assert dm.gs_kernel(amg.GS_AUTO) == amg.GS_KERNEL_LINESCAN, l


def test_vcycle_gs_level0_wavefront_matches_oracle():
    """V-cycles with the default Gauss-Seidel mode (wavefront kernel on level 0, scan kernel below) stay
    within rounding of the oracle on every level."""
    n, L = 129, 10
    mg, mo, _ = make_pair(n, L, amg.SparseGaussSeidel())
    for _ in range(3):
        mg.vcycle(); mo.vcycle()
    for l in range(L):
        assert rel(mg.get_soln(l), mo.u(l)) <= 1e-12, l


@pytest.mark.parametrize("n,L,eps", [(65, 9, 1.0), (129, 11, 1e-3), (257, 10, 1.0)])
def test_gauss_seidel_linescan_on_every_coarse_operator(n, L, eps):
    """The Galerkin operators (7 / 9 diagonals, distance-1 chain on every row) through the
    stand-alone smoother entry point, level by level."""
    Ao = O.laplacian(n, eps)
    mo = O.Multigrid(Ao, O.rhs(n), L, 1e-9, 1, 1)
    for l in range(L):
        Al = mo.A(l)
        c, r, v = Al.arrays()
        A = amg.CscMatrix(Al.rows, Al.cols, c, r, v)
        f, u = vec(Al.rows, 20 + l), vec(Al.rows, 40 + l)
        want = u.copy()
        O.gs_smooth(Al, want, f, 1e-9, 0, 2)
        got = u.copy()
        sm = amg.SparseGaussSeidel()
        sm.n_iters = 2
        sm.smooth(A, got, f)
        assert rel(got, want) <= 1e-13, (l, rel(got, want))


def test_spgs_as_solver_golden(capsys):
    """testlib.cpp:188-196: SparseGaussSeidel(1e-9,100,1000) on the 35x35 system."""
    A, b, Ao = problem(35)
    u = np.zeros(1225)
    uo = np.zeros(1225)
    O.gs_smooth(Ao, uo, b, 1e-9, 100, 1000)
    for mode in (amg.GS_LEVELSCHED, amg.GS_AUTO):
        u = np.zeros(1225)
        sm = amg.SparseGaussSeidel(1e-9, 100, 1000, mode=mode)
        sm.smooth(A, u, b)
        assert sm.iters_done == 900
        assert "SPGS converged after 900 iterations." in capsys.readouterr().out
        err = amg.rss(A, u, b)
        assert err < sm.tolerance
        assert "%.6g" % err == "8.69692e-10"
        if mode == amg.GS_LEVELSCHED:
            assert u.tobytes() == uo.tobytes()
        else:
            assert rel(u, uo) <= RTOL


def test_four_dof_smoothers_reach_direct_solution():
    """testlib.cpp:17-107 on the 2x2 grid."""
    A, b, Ao = problem(2)
    assert b.size == 4
    exact = O.Ldlt(Ao).solve(b)
    for sm in (amg.SparseGaussSeidel(100), amg.MulticolorGaussSeidel(100), amg.DampedJacobi(1.0, 100)):
        u = np.zeros(4)
        sm.smooth(A, u, b)
        d = u - exact
        assert np.dot(d, d) <= 1e-18 * min(np.dot(u, u), np.dot(exact, exact))   # isApprox(.,1e-9)


# ----------------------------------------------------------------- hierarchy
def make_pair(n, n_levels, smoother, eps=1.0, every=5, n_iters=100, **kw):
    A, b, Ao = problem(n, eps)
    kind = {amg.SMOOTHER_GS: O.SMOOTHER_GS, amg.SMOOTHER_JACOBI: O.SMOOTHER_JACOBI,
            amg.SMOOTHER_COLOR_GS: O.SMOOTHER_COLOR_GS}[smoother.kind]
    mo = O.Multigrid(Ao, b, n_levels, 1e-9, every, n_iters, kind, smoother.n_iters,
                     getattr(smoother, "omega", 2.0 / 3.0))
    li = amg.LinearInterpolator(n_levels)
    mg = amg.Multigrid(li, smoother, A, b, n_levels, 1e-9, every, n_iters, **kw)
    return mg, mo, li


@pytest.mark.parametrize("n,L,eps", [(35, 8, 1.0), (64, 6, 1.0), (65, 9, 1e-3)])
def test_hierarchy_maps_and_operators_bit_exact(n, L, eps):
    mg, mo, li = make_pair(n, L, amg.SparseGaussSeidel(), eps)
    for l in range(L):
        assert mg.get_n_dofs(l) == mo.n_dofs(l)
        M = mg.get_coefficient_matrix(l)
        c, r, v = mo.A(l).arrays()
        np.testing.assert_array_equal(M.colptr, c)       # structural pattern incl. explicit zeros
        np.testing.assert_array_equal(M.rowidx, r)
        assert M.val.tobytes() == v.tobytes()
        assert mg.nnz(l) == mo.A(l).nnz
        assert mg.nnz_device(l) == mo.A(l).nnz_nonzero
    for l in range(L - 1):
        c, r, v = mo.P(l).arrays()
        np.testing.assert_array_equal(li.get_P(l).colptr, c)
        np.testing.assert_array_equal(li.get_P(l).rowidx, r)
        c, r, v = mo.R(l).arrays()
        np.testing.assert_array_equal(li.get_R(l).colptr, c)
        np.testing.assert_array_equal(li.get_R(l).rowidx, r)


def test_level_shapes_decrease():
    """testlib.cpp:167-181."""
    mg, _, _ = make_pair(35, 8, amg.SparseGaussSeidel())
    sizes = [mg.get_n_dofs(l) for l in range(8)]
    assert sizes == [1225, 612, 305, 152, 75, 37, 18, 8]
    for l in range(1, 8):
        assert mg.get_soln(l - 1).size > mg.get_soln(l).size
        assert mg.get_rhs(l - 1).size > mg.get_rhs(l).size


@pytest.mark.parametrize("nh", [7, 24, 1225, 612, 305])
def test_transfers_bit_exact(nh):
    """restriction / prolongation+add incl. the even-n_h case whose last fine row is
    in no column of P."""
    nH = O.n_H_from_n_h(nh)
    # a 2-level hierarchy whose fine level has nh rows: use a 1-D Laplacian-like CSC
    colptr = np.arange(nh + 1, dtype=np.int32)
    A = amg.CscMatrix(nh, nh, colptr, np.arange(nh, dtype=np.int32), -np.ones(nh))
    mg = amg.Multigrid(None, amg.DampedJacobi(), A, np.ones(nh), 2, 1e-9, 1, 1)
    P = O.make_P(nh, nH)
    R = P.transpose()
    r = vec(nh, 6)
    assert mg.restrict(0, r).tobytes() == O.spmv(R, r).tobytes()
    e, u = vec(nH, 7), vec(nh, 8)
    assert mg.prolong_add(0, e, u).tobytes() == (u + O.spmv(P, e)).tobytes()


@pytest.mark.parametrize("n,L,eps", [(35, 8, 1.0), (64, 5, 1.0), (33, 6, 1e-3)])
def test_fused_residual_restrict_and_coarse_solve(n, L, eps):
    mg, mo, _ = make_pair(n, L, amg.SparseGaussSeidel(), eps)
    u0 = vec(n * n, 9)
    mg.set_soln(0, u0)
    mo.u(0)[:] = u0
    # fused kernel == residual then restriction, and zeroes the coarse solution
    mg.set_soln(1, np.ones(mg.get_n_dofs(1)))
    mg.residual_restrict_level(0)
    r = O.residual(mo.A(0), mo.u(0), mo.f(0))
    assert mg.residual_level(0).tobytes() == r.tobytes()
    assert mg.get_rhs(1).tobytes() == O.spmv(mo.R(0), r).tobytes()
    assert not mg.get_soln(1).any()
    # coarsest direct solve
    fL = vec(mg.get_n_dofs(L - 1), 10)
    mg.set_rhs(L - 1, fL)
    mg.coarse_solve()
    want = O.Ldlt(mo.A(L - 1)).solve(fL)
    assert rel(mg.get_soln(L - 1), want) <= 1e-13
    dense = mo.A(L - 1).to_scipy().toarray()
    assert rel(dense @ mg.get_soln(L - 1), fL) <= 1e-10


@pytest.mark.parametrize("n,L", [(35, 8), (35, 6), (35, 5), (35, 4), (35, 3), (64, 6), (64, 4), (129, 12),
                                 (129, 9), (257, 13)])
def test_coarse_solve_every_bandwidth(n, L):
    """Coarsest direct solve (multigrid.hpp:287-288) for half-bandwidths 1..8 (serial-recurrence
    kernel), 9..31 (warp kernel) and wider (block kernel): the oracle's banded LDL^T, bit for bit."""
    mg, mo, _ = make_pair(n, L, amg.DampedJacobi(2.0 / 3.0, 2), 1.0)
    N = mg.get_n_dofs(L - 1)
    for seed in (3, 4):
        fL = vec(N, seed)
        mg.set_rhs(L - 1, fL)
        mg.coarse_solve()
        want = O.Ldlt(mo.A(L - 1)).solve(fL)
        assert np.array_equal(mg.get_soln(L - 1), want)


SMOOTHERS = [
    ("gs", lambda: amg.SparseGaussSeidel(mode=amg.GS_LEVELSCHED)),
    ("gs_linescan", lambda: amg.SparseGaussSeidel(mode=amg.GS_LINESCAN)),
    ("gs_auto", lambda: amg.SparseGaussSeidel()),
    ("jacobi", lambda: amg.DampedJacobi(2.0 / 3.0, 2)),
    ("color", lambda: amg.MulticolorGaussSeidel(1)),
]


@pytest.mark.parametrize("name,mk", SMOOTHERS)
@pytest.mark.parametrize("n,L,eps", [(35, 8, 1.0), (64, 6, 1.0), (65, 9, 1e-3)])
def test_vcycle_per_level_iterates(name, mk, n, L, eps):
    """Per-level smoothed iterates after 1..3 V-cycles agree with the oracle."""
    mg, mo, _ = make_pair(n, L, mk(), eps)
    for cycle in range(3):
        mg.vcycle()
        mo.vcycle()
        for l in range(L):
            got, want = mg.get_soln(l), mo.u(l)
            assert rel(got, want) <= RTOL, (name, cycle, l, rel(got, want))
            if l + 1 < L or True:
                assert rel(mg.get_rhs(l), mo.f(l)) <= RTOL or not mo.f(l).any()
    # these kernels keep the oracle's operation order: expect identical bits on the fine level
    if name not in ("gs_linescan", "gs_auto"):   # (gs_auto: the scan kernel still smooths the Galerkin levels)
        assert mg.get_soln(0).tobytes() == mo.u(0).tobytes()


def test_multicolor_maps_per_level_bit_exact():
    mg, mo, _ = make_pair(35, 8, amg.MulticolorGaussSeidel(1))
    for l in range(7):
        nc, color = mg.coloring(l)
        assert nc == mo.n_colors(l)
        np.testing.assert_array_equal(color, mo.color(l))


@pytest.mark.parametrize("mode", [amg.GS_AUTO, amg.GS_LEVELSCHED])
def test_amg_solve_golden(capsys, mode):
    """testlib.cpp:147-212 on the GPU: 35 cycles, error 7.19199e-11, AMG ~= SPGS."""
    mg, mo, _ = make_pair(35, 8, amg.SparseGaussSeidel(mode=mode))
    A, b, _ = problem(35)
    amg_u = mg.solve()
    assert mg.iters_done == 35
    assert "AMG converged after 35 iterations." in capsys.readouterr().out
    mo.solve()
    assert mo.iters_done == 35                                   # identical iteration counts
    hist, hist_o = mg.error_history(), mo.history()
    assert len(hist) == len(hist_o) == 7
    if mode == amg.GS_LEVELSCHED:
        np.testing.assert_allclose(hist, hist_o, rtol=RTOL)      # residual norms, bit-exact kernels
    else:
        # The scan re-associates the distance-1 chain: iterates agree to ~1e-16, but a residual
        # that has dropped 11 orders of magnitude is a difference of nearly equal numbers, so
        # its RELATIVE agreement is limited to (1e-16 * |A||u|)^2-level noise over sum r^2.
        np.testing.assert_allclose(hist, hist_o, rtol=1e-6, atol=RTOL * hist_o[0])
    amg_error = amg.rss(A, amg_u, b)
    assert amg_error < mg.get_tolerance()                        # testlib.cpp:206
    assert "%.6g" % amg_error == "7.19199e-11"                   # README screenshot
    assert rel(amg_u, mo.u(0)) <= RTOL
    spgs_u = np.zeros(1225)
    amg.SparseGaussSeidel(1e-9, 100, 1000).smooth(A, spgs_u, b)
    d = amg_u - spgs_u
    assert np.dot(d, d) <= 1e-12 * min(np.dot(amg_u, amg_u), np.dot(spgs_u, spgs_u))  # :212


@pytest.mark.parametrize("name,mk", SMOOTHERS[3:])
def test_solve_iteration_counts_match_oracle(name, mk):
    mg, mo, _ = make_pair(35, 8, mk(), every=5, n_iters=400)
    mg.solve()
    mo.solve()
    assert mg.iters_done == mo.iters_done
    np.testing.assert_allclose(mg.error_history(), mo.history(), rtol=RTOL)
    assert mg.last_error <= 1e-9


def test_solve_resumes_from_stored_solution():
    """solve() continues from level_to_soln[0] (multigrid.hpp:101,336)."""
    mg, mo, _ = make_pair(35, 8, amg.SparseGaussSeidel(mode=amg.GS_LEVELSCHED), every=5, n_iters=10)
    mg.solve(); mo.solve()
    assert mg.iters_done == mo.iters_done == 10
    mg.solve(); mo.solve()
    assert rel(mg.get_soln(0), mo.u(0)) <= RTOL
    np.testing.assert_allclose(mg.error_history(), mo.history(), rtol=RTOL)


def test_dead_coarse_smooth_is_unobservable_and_graph_equals_stream():
    """The reference pre-smooths the coarsest level and then overwrites it with the direct
    solve (multigrid.hpp:265-288); skipping that must not change any observable value.
    A replayed CUDA graph must equal plain stream launches."""
    gs = lambda: amg.SparseGaussSeidel(mode=amg.GS_LEVELSCHED)
    base, _, _ = make_pair(35, 8, gs())
    full, _, _ = make_pair(35, 8, gs(), skip_dead_coarse_smooth=False)
    nograph, _, _ = make_pair(35, 8, gs(), use_graph=False)
    for m in (base, full, nograph):
        m.vcycles(3)
    for l in range(8):
        assert base.get_soln(l).tobytes() == full.get_soln(l).tobytes()
        assert base.get_soln(l).tobytes() == nograph.get_soln(l).tobytes()
    assert base.launches_per_vcycle() > 0


@pytest.mark.parametrize("n,L,eps", [(35, 8, 1.0), (64, 6, 1.0), (129, 10, 1e-3)])
def test_fused_jacobi_cycle_is_bit_identical(n, L, eps):
    """The zero-guess first sweep and the prolongation fused into the first post-smoothing
    sweep only skip memory traffic; every bit of every level must be unchanged."""
    fused, mo, _ = make_pair(n, L, amg.DampedJacobi(2.0 / 3.0, 2), eps, fuse=3)
    default, _, _ = make_pair(n, L, amg.DampedJacobi(2.0 / 3.0, 2), eps)
    plain, _, _ = make_pair(n, L, amg.DampedJacobi(2.0 / 3.0, 2), eps, fuse=0)
    odd, mo3, _ = make_pair(n, L, amg.DampedJacobi(0.6, 3), eps, fuse=3)
    for _ in range(3):
        fused.vcycle(); default.vcycle(); plain.vcycle(); mo.vcycle(); odd.vcycle(); mo3.vcycle()
    for l in range(L):
        assert fused.get_soln(l).tobytes() == plain.get_soln(l).tobytes()
        assert default.get_soln(l).tobytes() == plain.get_soln(l).tobytes()
        assert fused.get_soln(l).tobytes() == mo.u(l).tobytes()
        assert odd.get_soln(l).tobytes() == mo3.u(l).tobytes()
    assert fused.launches_per_vcycle() < plain.launches_per_vcycle()


@pytest.mark.parametrize("fuse", [5, 9, 13])
@pytest.mark.parametrize("n,L,eps,nu", [(35, 8, 1.0, 2), (64, 6, 1.0, 2), (100, 9, 1.0, 2), (129, 10, 1e-3, 2),
                                        (129, 12, 1.0, 1), (257, 13, 1.0, 2), (200, 11, 1.0, 3),
                                        (513, 14, 1.0, 2)])
def test_fused_legs_cycle_is_bit_identical(n, L, eps, nu, fuse):
    """Fused legs (one kernel per level and leg, operator read once) against the per-operator
    kernels and the oracle: every level's iterate and right-hand side, bit for bit.
    fuse bit 2 = register-streaming legs (3 x 3 line stencils, one or two sweeps), bit 3 = TMA-ring legs."""
    sm = amg.DampedJacobi(2.0 / 3.0, nu)
    legs, mo, _ = make_pair(n, L, sm, eps, fuse=fuse)
    plain, _, _ = make_pair(n, L, sm, eps, fuse=0)
    n_fused = sum(legs.fused_legs(l) for l in range(L - 1))
    if (fuse & 8) or (nu in (1, 2) and n >= 100):
        assert n_fused > 0
    assert not any(plain.fused_legs(l) for l in range(L))
    for _ in range(3):
        legs.vcycle(); plain.vcycle(); mo.vcycle()
    for l in range(L):
        assert legs.get_soln(l).tobytes() == plain.get_soln(l).tobytes(), l
        assert legs.get_soln(l).tobytes() == mo.u(l).tobytes(), l
        assert legs.get_rhs(l).tobytes() == mo.f(l).tobytes(), l
    if n_fused:
        assert legs.launches_per_vcycle() < plain.launches_per_vcycle()


@pytest.mark.parametrize("n,L,eps,nu", [(35, 8, 1.0, 2), (35, 7, 1.0, 1), (64, 9, 1.0, 3), (100, 11, 1.0, 2),
                                        (129, 12, 1e-3, 2), (257, 14, 1.0, 2)])
def test_coarse_tail_is_bit_identical(n, L, eps, nu):
    """The small levels + coarsest solve in ONE launch (fuse bit 4) against the per-operator
    kernels and the oracle, every level, bit for bit; and on its own as well as combined with
    the streaming legs."""
    sm = amg.DampedJacobi(2.0 / 3.0, nu)
    tail, mo, _ = make_pair(n, L, sm, eps, fuse=16)
    both, _, _ = make_pair(n, L, sm, eps, fuse=1 | 4 | 16)
    plain, _, _ = make_pair(n, L, sm, eps, fuse=0)
    assert 1 <= tail.tail_first() < L - 1
    assert plain.tail_first() == -1
    for _ in range(3):
        tail.vcycle(); both.vcycle(); plain.vcycle(); mo.vcycle()
    for l in range(L):
        assert tail.get_soln(l).tobytes() == mo.u(l).tobytes(), l
        assert both.get_soln(l).tobytes() == mo.u(l).tobytes(), l
        assert plain.get_soln(l).tobytes() == mo.u(l).tobytes(), l
        assert tail.get_rhs(l).tobytes() == mo.f(l).tobytes(), l
    assert tail.launches_per_vcycle() < plain.launches_per_vcycle()


def test_fused_legs_small_tiles(monkeypatch):
    """Force narrow strips / short line chunks so one level is cut into many tiles."""
    monkeypatch.setenv("AMGB_LEG_W", "37")
    monkeypatch.setenv("AMGB_LEG_LJ", "11")
    sm = amg.DampedJacobi(2.0 / 3.0, 2)
    legs, mo, _ = make_pair(257, 12, sm, 1.0, fuse=9)
    assert legs.leg_plan(0)["W"] <= 37 and legs.leg_plan(0)["tiles"] > 100
    for _ in range(2):
        legs.vcycle(); mo.vcycle()
    for l in range(12):
        assert legs.get_soln(l).tobytes() == mo.u(l).tobytes(), l


def test_fused_legs_iteration_count_matches_oracle():
    sm = amg.DampedJacobi(2.0 / 3.0, 2)
    legs, mo, _ = make_pair(35, 8, sm, 1.0, every=5, n_iters=400, fuse=5)
    legs.solve(); mo.solve()
    assert legs.iters_done == mo.iters_done
    np.testing.assert_allclose(legs.error_history(), mo.history(), rtol=1e-12)


@pytest.mark.parametrize("n,L,eps", [(35, 8, 1.0), (100, 9, 1.0), (129, 10, 1e-3), (257, 12, 1.0)])
def test_device_galerkin_is_bit_identical(n, L, eps):
    """SURVEY 8f rank 1 building block: A_{l+1} = R (A_l P) computed on the device from the level's
    DIA mirror equals the host-built (oracle-pinned) coarse operator bit for bit, every level."""
    mg, _, _ = make_pair(n, L, amg.DampedJacobi(2.0 / 3.0, 2), eps)
    for l in range(L - 1):
        if mg.format(l) != "dia" or mg.format(l + 1) != "dia":
            continue
        ms, bad = mg.galerkin_device(l)
        assert bad == 0, (l, bad)
        assert ms >= 0.0


def oracle_pcg(mo, Ao, b, rel_tol, max_iters):
    """CG preconditioned by one oracle V-cycle from a zero guess (test infrastructure)."""
    x = mo.u(0).copy()
    r = O.residual(Ao, x, b)
    bn2 = float(b @ b)
    hist, p, rz_old = [], None, 0.0
    for it in range(max_iters):
        if np.sqrt(float(r @ r) / bn2) <= rel_tol:
            break
        mo.f(0)[:] = r
        mo.u(0)[:] = 0.0
        mo.vcycle()
        z = mo.u(0).copy()
        rz = float(r @ z)
        p = z.copy() if p is None else z + (rz / rz_old) * p
        q = O.residual(Ao, p, np.zeros_like(p))   # 0 - A p
        alpha = -rz / float(p @ q)
        x = x + alpha * p
        r = r + alpha * q
        rz_old = rz
        hist.append(np.sqrt(float(r @ r) / bn2))
    return x, np.array(hist)


@pytest.mark.parametrize("n,L,eps", [(65, 9, 1.0), (129, 11, 1.0), (129, 11, 1e-3)])
def test_pcg_with_vcycle_preconditioner(n, L, eps):
    """Beyond the reference: CG preconditioned by the V-cycle reaches 1e-8
    relative residual in a fraction of the plain cycles, with the oracle's iteration count."""
    sm = amg.DampedJacobi(2.0 / 3.0, 2)
    mg, mo, _ = make_pair(n, L, sm, eps)
    A, b, Ao = problem(n, eps)
    want_x, want_hist = oracle_pcg(mo, Ao, b, 1e-8, 500)
    got_x = mg.solve_pcg(1e-8, 500)
    assert mg.iters_done == len(want_hist)
    assert mg.last_error <= 1e-8
    np.testing.assert_allclose(mg.error_history(), want_hist, rtol=1e-6)
    assert rel(got_x, want_x) <= 1e-9
    assert np.sqrt(O.rss(Ao, got_x, b) / float(b @ b)) <= 2e-8     # true residual of the returned iterate
    # the hierarchy still holds b and the solution: plain cycles continue from it
    assert mg.get_rhs(0).tobytes() == b.tobytes()
    plain, _, _ = make_pair(n, L, sm, eps, every=1, n_iters=mg.iters_done)
    plain.solve_relative(0.0)
    assert mg.last_error < plain.last_error   # far ahead of the same number of plain V-cycles


def test_relative_residual_criterion():
    mg, mo, _ = make_pair(35, 8, amg.SparseGaussSeidel(mode=amg.GS_LEVELSCHED), every=1, n_iters=200)
    A, b, Ao = problem(35)
    mg.solve_relative(1e-8)
    k = mg.iters_done
    bn = np.linalg.norm(b)
    for _ in range(k - 1):
        mo.vcycle()
    assert np.sqrt(mo.rss()) / bn > 1e-8
    mo.vcycle()
    assert np.sqrt(mo.rss()) / bn <= 1e-8
    assert abs(mg.last_error - np.sqrt(mo.rss()) / bn) <= 1e-10 * mg.last_error


# ----------------------------------------------------------------- larger sizes
def test_jacobi_vcycle_1025():
    """One full-size V-cycle (1025^2, 14 levels) against the oracle."""
    n, L = 1025, 14
    mg, mo, _ = make_pair(n, L, amg.DampedJacobi(2.0 / 3.0, 2), every=1, n_iters=1)
    mg.vcycle()
    mo.vcycle()
    assert [mg.get_n_dofs(l) for l in range(L)] == O.level_sizes(n * n, L)
    for l in range(L):
        assert rel(mg.get_soln(l), mo.u(l)) <= RTOL
    assert abs(mg.rss() - mo.rss()) <= 1e-12 * mo.rss()


def test_gs_vcycle_1025_config2():
    """BASELINE config 2: 1025^2, symmetric Gauss-Seidel (line-scan kernel), 14 levels."""
    n, L = 1025, 14
    mg, mo, _ = make_pair(n, L, amg.SparseGaussSeidel(), every=1, n_iters=2)
    for _ in range(2):
        mg.vcycle()
        mo.vcycle()
    for l in range(L):
        assert rel(mg.get_soln(l), mo.u(l)) <= RTOL, (l, rel(mg.get_soln(l), mo.u(l)))
    assert abs(mg.rss() - mo.rss()) <= 1e-12 * mo.rss()


def test_properties_at_scale_4097():
    """Size-independent properties at the BASELINE size (no oracle run needed):
    linearity of the residual/restriction, R = P^T adjointness, Jacobi fixed point."""
    n, L = 4097, 16
    A = amg.Grid.laplacian(n)
    b = amg.Grid.rhs(n)
    with pytest.raises(amg.InvalidArgument, match="coarsest level too large"):
        amg.Multigrid(None, amg.DampedJacobi(2.0 / 3.0, 2), amg.Grid.laplacian(1025),
                      amg.Grid.rhs(1025), 2, 1e-9, 1, 1)
    mg = amg.Multigrid(None, amg.DampedJacobi(2.0 / 3.0, 2), A, b, L, 1e-9, 1, 1)
    N0, N1 = mg.get_n_dofs(0), mg.get_n_dofs(1)
    assert (N0, N1, mg.get_n_dofs(2)) == (16785409, 8392704, 4196351)
    assert mg.nnz_device(0) == 5 * n * n - 4 * n
    x, y = vec(N0, 11), vec(N0, 12)
    # adjointness <R x, e> == <x, P e>
    e = vec(N1, 13)
    Rx = mg.restrict(0, x)
    Pe = mg.prolong_add(0, e, np.zeros(N0))
    assert abs(np.dot(Rx, e) - np.dot(x, Pe)) <= 1e-12 * np.linalg.norm(Rx) * np.linalg.norm(e)
    # residual linearity: r(u1) - r(u2) == -A (u1 - u2) == r_{f=0}(u1 - u2)
    mg.set_soln(0, x); r1 = mg.residual_level(0)
    mg.set_soln(0, y); r2 = mg.residual_level(0)
    mg.set_rhs(0, np.zeros(N0)); mg.set_soln(0, x - y); r3 = mg.residual_level(0)
    assert rel(r1 - r2, r3) <= 1e-12
    # the 5-point stencil applied to a constant: interior rows sum to zero
    mg.set_soln(0, np.ones(N0)); r = mg.residual_level(0).reshape(n, n)
    assert np.abs(r[1:-1, 1:-1]).max() <= 1e-6 * abs(A.val[2])


def test_vcycle_4097_one_cycle_against_oracle():
    """The benchmarked configuration itself (BASELINE configs[2]: 4097^2, 18 levels, damped Jacobi
    2 + 2 sweeps): one V-cycle against the oracle, both arithmetic modes -- reference order must be
    bit-identical on level 0, fast within the 1e-12 contract on every level."""
    n, L = 4097, 18
    sm = amg.DampedJacobi(2.0 / 3.0, 2)
    A, b, Ao = problem(n)
    mo = O.Multigrid(Ao, b, L, 1e-9, 1, 1, O.SMOOTHER_JACOBI, 2, 2.0 / 3.0)
    mo.vcycle()
    for arith in (amg.ARITH_REFERENCE, amg.ARITH_FAST):
        mg = amg.Multigrid(None, sm, A, b, L, 1e-9, 1, 1, arith=arith)
        assert mg.fused_legs(0)
        mg.vcycle()
        for l in range(L):
            assert rel(mg.get_soln(l), mo.u(l)) <= RTOL, (arith, l, rel(mg.get_soln(l), mo.u(l)))
        if arith == amg.ARITH_REFERENCE:
            assert mg.get_soln(0).tobytes() == mo.u(0).tobytes()
        # sum r^2: over 1.7e7 terms the reference's sequential sum (common.hpp:22-25) carries ~1e-11
        # rounding of its own; compare with an accurate (pairwise) sum of the oracle's residual vector
        r = O.residual(Ao, mo.u(0), b)
        want = float(np.dot(r, r))
        assert abs(mg.rss() - want) <= RTOL * want
        assert abs(mo.rss() - want) <= 1e-10 * want
        del mg


# ----------------------------------------------------------------- fast arithmetic (AMGB_ARITH_FAST)
@pytest.mark.parametrize("n,L,eps,nu", [(100, 9, 1.0, 2), (129, 12, 1.0, 1), (129, 10, 1e-3, 2), (257, 13, 1.0, 2),
                                        (513, 14, 1.0, 2), (1025, 14, 1.0, 2)])
def test_fast_arithmetic_cycle_within_contract(n, L, eps, nu):
    """FMA + refined-reciprocal legs: every level's iterate and right-hand side within 1e-12 of the
    oracle after three cycles (the north star's tolerance), and not bit-identical by accident
    (i.e. the fast kernels really ran)."""
    sm = amg.DampedJacobi(2.0 / 3.0, nu)
    fast, mo, _ = make_pair(n, L, sm, eps, arith=amg.ARITH_FAST)
    assert sum(fast.fused_legs(l) for l in range(L - 1)) > 0
    for _ in range(3):
        fast.vcycle(); mo.vcycle()
    worst = 0.0
    for l in range(L):
        worst = max(worst, rel(fast.get_soln(l), mo.u(l)), rel(fast.get_rhs(l), mo.f(l)))
    assert worst <= RTOL, worst
    assert abs(fast.rss() - mo.rss()) <= RTOL * mo.rss()


def test_fast_arithmetic_iteration_count_matches_oracle():
    sm = amg.DampedJacobi(2.0 / 3.0, 2)
    for n, L in ((35, 8), (100, 9)):
        fast, mo, _ = make_pair(n, L, sm, 1.0, every=5, n_iters=400, arith=amg.ARITH_FAST)
        fast.solve(); mo.solve()
        assert fast.iters_done == mo.iters_done
        # sum r^2 near convergence is a difference of nearly equal numbers: compare loosely there
        np.testing.assert_allclose(fast.error_history(), mo.history(), rtol=1e-6)


# ----------------------------------------------------------------- general (non-banded) matrices: SELL-32
def permuted_problem(n, seed):
    """The five-point operator under a random symmetric permutation: same spectrum, but hundreds of
    distinct diagonals, so the device mirror must take the SELL-32 layout (no DIA)."""
    import scipy.sparse as sp
    A = amg.Grid.laplacian(n)
    N = n * n
    perm = np.random.default_rng(seed).permutation(N)
    Pm = sp.csc_matrix((np.ones(N), (np.arange(N), perm)), shape=(N, N))
    M = (Pm @ A.to_scipy() @ Pm.T).tocsc()
    M.sort_indices()
    Ap = amg.CscMatrix(N, N, M.indptr, M.indices, M.data)
    return Ap, O.Csc.from_arrays(N, N, M.indptr, M.indices, M.data), amg.Grid.rhs(n)[perm]


@pytest.mark.parametrize("n", [24, 61])
def test_sell_layout_operators_bit_exact(n):
    """Residual, Jacobi, multicolour GS, level-scheduled GS and rss on a matrix that is NOT banded."""
    Ap, Ao, b = permuted_problem(n, 5)
    N = n * n
    dm = amg.DeviceMatrix(Ap)
    u = vec(N, 6)
    assert dm.residual(u, b).tobytes() == O.residual(Ao, u, b).tobytes()
    assert abs(dm.rss(u, b) - O.rss(Ao, u, b)) <= 1e-13 * O.rss(Ao, u, b)
    AT = Ao.transpose()
    want = O.jacobi_sweep(AT, O.jacobi_sweep(AT, u, b, 0.6), b, 0.6)
    got = u.copy()
    amg.DampedJacobi(0.6, 2).smooth(dm, got, b)
    assert got.tobytes() == want.tobytes()
    nc, color = O.greedy_coloring(Ao, AT)
    nc_g, color_g = dm.coloring()
    assert nc == nc_g and np.array_equal(color, color_g)
    want = u.copy()
    for c in list(range(nc)) + list(range(nc - 1, -1, -1)):
        O.color_gs_pass(AT, color, c, b, want)
    got = u.copy()
    amg.MulticolorGaussSeidel(1).smooth(dm, got, b)
    assert got.tobytes() == want.tobytes()
    want = u.copy()
    O.gs_forward(Ao, b, want); O.gs_backward(Ao, b, want)
    got = u.copy()
    amg.SparseGaussSeidel(mode=amg.GS_LEVELSCHED).smooth(dm, got, b)
    assert got.tobytes() == want.tobytes()
    got = u.copy()
    amg.SparseGaussSeidel().smooth(dm, got, b)     # AUTO must fall back to the generic kernel
    assert got.tobytes() == want.tobytes()


@pytest.mark.parametrize("name,mk", SMOOTHERS)
def test_sell_layout_vcycle(name, mk):
    """A three-level V-cycle whose finest operator is in the SELL-32 layout, every level vs the oracle."""
    n, L = 40, 3
    Ap, Ao, b = permuted_problem(n, 7)
    sm = mk()
    kind = {amg.SMOOTHER_GS: O.SMOOTHER_GS, amg.SMOOTHER_JACOBI: O.SMOOTHER_JACOBI,
            amg.SMOOTHER_COLOR_GS: O.SMOOTHER_COLOR_GS}[sm.kind]
    mg = amg.Multigrid(amg.LinearInterpolator(L), sm, Ap, b, L, 1e-9, 1, 1)
    mo = O.Multigrid(Ao, b, L, 1e-9, 1, 1, kind, sm.n_iters, getattr(sm, "omega", 2.0 / 3.0))
    assert mg.format(0) == "sell"
    for _ in range(2):
        mg.vcycle(); mo.vcycle()
    for l in range(L):
        assert rel(mg.get_soln(l), mo.u(l)) <= RTOL, (name, l)
    assert abs(mg.rss() - mo.rss()) <= RTOL * mo.rss()


def test_stored_interpolation_operators_are_applied():
    """InterpolatorBase::prolongation / restriction multiply by the STORED operators
    (interpolator.hpp:52-68), whatever they are: generic device SpMV vs the oracle."""
    li = amg.LinearInterpolator(2)
    li.make_operators(25, 12, 0)
    P, R = li.get_P(0), li.get_R(0)
    e, r = vec(12, 8), vec(25, 9)
    Po = O.Csc.from_arrays(P.rows, P.cols, P.colptr, P.rowidx, P.val)
    Ro = O.Csc.from_arrays(R.rows, R.cols, R.colptr, R.rowidx, R.val)
    assert li.prolongation(e, 0).tobytes() == O.spmv(Po, e).tobytes()
    assert li.restriction(r, 0).tobytes() == O.spmv(Ro, r).tobytes()
    P2 = amg.CscMatrix(P.rows, P.cols, P.colptr, P.rowidx, P.val * np.linspace(1, 2, P.nnz))
    li.set_level_to_P(0, P2)
    P2o = O.Csc.from_arrays(P2.rows, P2.cols, P2.colptr, P2.rowidx, P2.val)
    assert li.prolongation(e, 0).tobytes() == O.spmv(P2o, e).tobytes()


def test_mirror_cache_sees_in_place_edits():
    A, b, Ao = problem(35)
    u = vec(35 * 35, 10)
    r0 = amg.rss(A, u, b)
    A.val *= 2.0
    A2o = O.Csc.from_arrays(A.rows, A.cols, A.colptr, A.rowidx, A.val)
    want = O.rss(A2o, u, b)
    got = amg.rss(A, u, b)
    assert abs(got - want) <= 1e-13 * want and abs(got - r0) > 1e-3 * r0


# ----------------------------------------------------------------- mid levels (fuse bit 5, mid_levels.cuh)
@pytest.mark.parametrize("n,L,eps,nu", [(35, 8, 1.0, 2), (64, 9, 1.0, 3), (100, 11, 1.0, 2), (129, 12, 1e-3, 2),
                                        (257, 14, 1.0, 2), (513, 15, 1.0, 1)])
@pytest.mark.parametrize("tail_rows,fuse", [("6000", 1 | 4 | 16 | 32), ("300", 1 | 4 | 16 | 32), ("6000", 32)])
def test_mid_levels_are_bit_identical(monkeypatch, n, L, eps, nu, tail_rows, fuse):
    """All down legs of the mid levels in one launch, all up legs in another (shared-memory tiles with
    recomputed halos), alone and together with the streaming legs / the coarse tail: every level's
    iterate and right-hand side bit-identical to the oracle."""
    monkeypatch.setenv("AMGB_TAIL_ROWS", tail_rows)
    sm = amg.DampedJacobi(2.0 / 3.0, nu)
    mid, mo, _ = make_pair(n, L, sm, eps, fuse=fuse)
    plain, _, _ = make_pair(n, L, sm, eps, fuse=0)
    first, end, tile, blocks = mid.mid_range()
    assert 0 <= first < end <= L - 1 and tile > 0 and blocks > 0
    assert plain.mid_range()[0] == -1
    for _ in range(3):
        mid.vcycle(); mo.vcycle()
    for l in range(L):
        assert mid.get_soln(l).tobytes() == mo.u(l).tobytes(), l
        assert mid.get_rhs(l).tobytes() == mo.f(l).tobytes(), l
    plain.vcycle()
    assert mid.launches_per_vcycle() < plain.launches_per_vcycle()


# ----------------------------------------------------------------- matrix-free five-point legs (fuse bit 6)
@pytest.mark.parametrize("n,L,eps,nu", [(100, 9, 1.0, 2), (129, 10, 1e-3, 2), (257, 13, 1.0, 1), (513, 14, 1.0, 2)])
def test_matrix_free_legs_are_bit_identical(n, L, eps, nu):
    """Level 0 is verified at setup to be a constant five-point stencil and runs legs that read no
    operator row; every level's iterate and right-hand side stay bit-identical to the oracle."""
    sm = amg.DampedJacobi(2.0 / 3.0, nu)
    mf, mo, _ = make_pair(n, L, sm, eps, fuse=1 | 4 | 64)
    plain, _, _ = make_pair(n, L, sm, eps, fuse=1 | 4)
    assert mf.fused_legs(0) and mf.matrix_free(0) and not mf.matrix_free(1)
    assert plain.fused_legs(0) and not plain.matrix_free(0)
    for _ in range(3):
        mf.vcycle(); mo.vcycle()
    for l in range(L):
        assert mf.get_soln(l).tobytes() == mo.u(l).tobytes(), l
        assert mf.get_rhs(l).tobytes() == mo.f(l).tobytes(), l


def test_matrix_free_needs_an_exactly_constant_stencil():
    """One perturbed coefficient and the verification refuses: the general legs run (and stay exact)."""
    n, L = 129, 10
    A, b, _ = problem(n)
    A.val[A.colptr[5000] + 2] *= 1.0 + 2.0 ** -40     # one diagonal entry
    sm = amg.DampedJacobi(2.0 / 3.0, 2)
    mg = amg.Multigrid(None, sm, A, b, L, 1e-9, 1, 1, fuse=1 | 4 | 64)
    assert mg.fused_legs(0) and not mg.matrix_free(0)
    mo = O.Multigrid(O.Csc.from_arrays(A.rows, A.cols, A.colptr, A.rowidx, A.val), b, L, 1e-9, 1, 1,
                     O.SMOOTHER_JACOBI, 2, 2.0 / 3.0)
    for _ in range(2):
        mg.vcycle(); mo.vcycle()
    assert rel(mg.get_soln(0), mo.u(0)) <= RTOL


# ----------------------------------------------------------------- row-type dictionary legs (fuse bit 7)
@pytest.mark.parametrize("n,L,eps,nu", [(100, 9, 1.0, 2), (129, 10, 1e-3, 2), (257, 13, 1.0, 1), (513, 14, 1.0, 2),
                                        (1025, 14, 1.0, 2)])
def test_dictionary_legs_are_bit_identical(n, L, eps, nu):
    """The fused legs read one byte per row and take the operator row from a table of the level's
    distinct rows (verified at setup); every level's iterate and right-hand side stay bit-identical."""
    sm = amg.DampedJacobi(2.0 / 3.0, nu)
    dic, mo, _ = make_pair(n, L, sm, eps, fuse=1 | 4 | 128)
    assert dic.fused_legs(0) and 1 <= dic.dictionary_types(0) <= 256 and not dic.matrix_free(0)
    assert sum(dic.dictionary_types(l) > 0 for l in range(L - 1)) >= 2
    for _ in range(3):
        dic.vcycle(); mo.vcycle()
    for l in range(L):
        assert dic.get_soln(l).tobytes() == mo.u(l).tobytes(), l
        assert dic.get_rhs(l).tobytes() == mo.f(l).tobytes(), l


def test_dictionary_needs_few_distinct_rows():
    """A level-0 operator with more than 256 distinct rows keeps the plain DIA legs."""
    n, L = 129, 10
    A, b, _ = problem(n)
    rng = np.random.default_rng(4)
    diag = np.flatnonzero(A.rowidx == np.repeat(np.arange(n * n), np.diff(A.colptr)))
    A.val[diag] *= 1.0 + 1e-3 * rng.random(n * n)            # every diagonal entry different
    sm = amg.DampedJacobi(2.0 / 3.0, 2)
    mg = amg.Multigrid(None, sm, A, b, L, 1e-9, 1, 1, fuse=1 | 4 | 64 | 128)
    assert mg.fused_legs(0) and mg.dictionary_types(0) == 0 and not mg.matrix_free(0)
    mo = O.Multigrid(O.Csc.from_arrays(A.rows, A.cols, A.colptr, A.rowidx, A.val), b, L, 1e-9, 1, 1,
                     O.SMOOTHER_JACOBI, 2, 2.0 / 3.0)
    for _ in range(2):
        mg.vcycle(); mo.vcycle()
    assert mg.get_soln(0).tobytes() == mo.u(0).tobytes()


def test_asymmetric_operator_device_setup():
    """An operator that is NOT symmetric: the device-side setup must build the rows of A (transposed
    mirror of the CSC columns) and every Galerkin level from them; damped-Jacobi cycles against the
    oracle, which reads rows of A for Jacobi and the residual."""
    n, L = 65, 9
    A, b, _ = problem(n)
    cols = np.repeat(np.arange(n * n), np.diff(A.colptr))
    upper = A.rowidx < cols                      # entries A(r, c) with r < c
    A.val[upper] *= 1.001                        # a small skew: the cycle still converges
    Ao = O.Csc.from_arrays(A.rows, A.cols, A.colptr, A.rowidx, A.val)
    sm = amg.DampedJacobi(2.0 / 3.0, 2)
    mg = amg.Multigrid(amg.LinearInterpolator(L), sm, A, b, L, 1e-9, 1, 1)
    mo = O.Multigrid(Ao, b, L, 1e-9, 1, 1, O.SMOOTHER_JACOBI, 2, 2.0 / 3.0)
    for _ in range(3):
        mg.vcycle(); mo.vcycle()
    for l in range(L):
        assert mg.get_soln(l).tobytes() == mo.u(l).tobytes(), l
        assert mg.get_rhs(l).tobytes() == mo.f(l).tobytes(), l
    # the getters hand back the structural matrices of the host Galerkin chain
    Ag, Aw = mg.get_coefficient_matrix(2), mo.A(2).arrays()
    assert np.array_equal(Ag.colptr, Aw[0]) and np.array_equal(Ag.rowidx, Aw[1]) and Ag.val.tobytes() == Aw[2].tobytes()

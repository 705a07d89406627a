"""Pins the CPU oracle against every golden value the reference holds for the
V-cycle path (README screenshot image/README/output.png, reproduced in
SURVEY.md section 6 / BASELINE.md section 1; configuration
/root/reference/test/testlib.cpp:147-212) and against the reference test's own
assertions (testlib.cpp:28-107, :131-144, :167-181).  CPU only."""
import numpy as np
import pytest

import oracle as O


def fmt6(x):
    """std::cout default formatting (6 significant digits)."""
    return "%g" % float("%.6g" % x)


@pytest.fixture(scope="module")
def amg35():
    A = O.laplacian(35)
    b = O.rhs(35)
    mg = O.Multigrid(A, b, 8, 1e-9, 5, 100)   # testlib.cpp:158-159
    return A, b, mg


def test_level_sizes_golden(amg35):
    _, _, mg = amg35
    assert [mg.n_dofs(l) for l in range(8)] == [1225, 612, 305, 152, 75, 37, 18, 8]
    # closed form, SURVEY appendix B
    assert O.level_sizes(1025 * 1025, 14) == [
        1050625, 525312, 262655, 131327, 65663, 32831, 16415, 8207, 4103, 2051,
        1025, 512, 255, 127]
    assert O.level_sizes(4097 * 4097, 16) == [
        16785409, 8392704, 4196351, 2098175, 1049087, 524543, 262271, 131135,
        65567, 32783, 16391, 8195, 4097, 2048, 1023, 511]


def test_structural_nnz_golden(amg35):
    _, _, mg = amg35
    # Eigen-style (explicit zeros kept) and pruned counts, SURVEY section 8c
    assert [mg.A(l).nnz for l in range(8)] == [5985, 4212, 2689, 1634, 655, 247, 84, 34]
    assert [mg.A(l).nnz_nonzero for l in range(8)] == [5985, 3058, 2689, 1634, 655, 247, 84, 34]


def test_level_shapes_decrease(amg35):
    # testlib.cpp:167-181
    _, _, mg = amg35
    for l in range(1, 8):
        assert mg.A(l - 1).rows * mg.A(l - 1).cols > mg.A(l).rows * mg.A(l).cols
        assert mg.u(l - 1).size > mg.u(l).size
        assert mg.f(l - 1).size > mg.f(l).size


def test_amg_solve_golden(amg35):
    A, b, mg = amg35
    iters = mg.solve()
    assert iters == 35                                   # "AMG converged after 35 iterations."
    err = O.rss(A, mg.u(0).copy(), b)
    assert fmt6(err) == "7.19199e-11"                    # "AMG error: 7.19199e-11"
    assert err < 1e-9                                    # testlib.cpp:206
    hist = mg.history()
    golden = [10.1281025645, 0.139287315378, 1.93293056496e-3, 2.68445752995e-5,
              3.72840112831e-7, 5.17830115457e-9, 7.19199383911e-11]
    assert len(hist) == 7
    np.testing.assert_allclose(hist, golden, rtol=2e-8)


def test_spgs_golden_and_amg_matches_spgs(amg35):
    A, b, mg = amg35
    u = np.zeros(1225)
    iters, err = O.gs_smooth(A, u, b, 1e-9, 100, 1000)   # testlib.cpp:188-193
    assert iters == 900                                  # "SPGS converged after 900 iterations."
    assert fmt6(O.rss(A, u, b)) == "8.69692e-10"         # "SPGS error: 8.69692e-10"
    if mg.iters_done == 0:
        mg.solve()
    v = mg.u(0)
    # Eigen isApprox(b, 1e-6): ||a-b||^2 <= p^2 * min(||a||^2, ||b||^2); testlib.cpp:212
    assert np.dot(v - u, v - u) <= 1e-12 * min(np.dot(u, u), np.dot(v, v))


def test_four_dof_system():
    # testlib.cpp:17-107
    A = O.laplacian(2)
    b = O.rhs(2)
    assert b.size == 4
    h = O.lib().orc_grid_spacing_h(2)
    assert O.lib().orc_points_n_from_grid_spacing_h(h) == 2   # :60-62
    dense = A.to_scipy().toarray()
    np.testing.assert_array_equal(dense, 2.25 * np.array(
        [[-4, 1, 1, 0], [1, -4, 0, 1], [1, 0, -4, 1], [0, 1, 1, -4.0]]))
    exact = O.Ldlt(A).solve(b)
    np.testing.assert_allclose(exact, np.linalg.solve(dense, b), rtol=1e-14)
    u = np.zeros(4)
    O.gs_smooth(A, u, b, 1e-9, 100, 100)                      # SparseGaussSeidel(100), :103-107
    assert np.dot(u - exact, u - exact) <= 1e-18 * min(np.dot(u, u), np.dot(exact, exact))


def test_interpolator_patterns_golden():
    # testlib.cpp:118-128 prints these; SURVEY appendix B spells them out
    P = O.make_P(7, 3)
    colptr, rowidx, val = P.arrays()
    assert colptr.tolist() == [0, 3, 6, 9]
    assert rowidx.tolist() == [0, 1, 2, 2, 3, 4, 4, 5, 6]
    assert val.tolist() == [0.5, 1.0, 0.5] * 3
    P = O.make_P(24, 11)
    colptr, rowidx, val = P.arrays()
    assert colptr.tolist() == list(range(0, 34, 3))
    assert rowidx[-3:].tolist() == [20, 21, 22]
    assert 23 not in rowidx.tolist()
    R = P.transpose()
    rc, rr, rv = R.arrays()
    assert (R.rows, R.cols) == (11, 24)
    counts = np.diff(rc).tolist()
    assert counts == [1] + [1, 2] * 10 + [1, 1, 0]
    assert rr[rc[22]:rc[23]].tolist() == [10] and rr[rc[0]:rc[1]].tolist() == [0]
    for k in range(1, 22, 2):
        assert rr[rc[k]:rc[k + 1]].tolist() == [(k - 1) // 2] and rv[rc[k]] == 1.0
    for k in range(2, 22, 2):
        assert rr[rc[k]:rc[k + 1]].tolist() == [k // 2 - 1, k // 2]


def test_ctor_validation_order():
    # multigrid.hpp:165-178, testlib.cpp:131-144
    A = O.laplacian(2)
    b = O.rhs(2)
    with pytest.raises(ValueError, match="compute_error_every_n_iters"):
        O.Multigrid(A, b, 2, 1e-9, 100, 10)
    with pytest.raises(ValueError, match="same number of degrees"):
        O.Multigrid(A, np.zeros(5), 2, 1e-9, 5, 10)
    with pytest.raises(ValueError, match="compute_error_every_n_iters"):
        O.Multigrid(A, np.zeros(5), 2, 1e-9, 100, 10)


def test_galerkin_matches_scipy():
    import scipy.sparse as sp
    A = O.laplacian(17)
    nh = 17 * 17
    nH = O.n_H_from_n_h(nh)
    P = O.make_P(nh, nH)
    R = P.transpose()
    AH = O.galerkin(R, A, P)
    ref = (R.to_scipy() @ (A.to_scipy() @ P.to_scipy())).toarray()
    np.testing.assert_allclose(AH.to_scipy().toarray(), ref, rtol=0, atol=1e-9)
    # columns sorted, explicit zeros kept
    colptr, rowidx, val = AH.arrays()
    for c in range(AH.cols):
        assert np.all(np.diff(rowidx[colptr[c]:colptr[c + 1]]) > 0)
    assert AH.nnz > AH.nnz_nonzero


def test_vcycle_equals_hand_composition():
    """vcycle() is exactly the composition of the per-operator oracle calls."""
    A = O.laplacian(9)
    b = O.rhs(9)
    mg = O.Multigrid(A, b, 3, 1e-9, 1, 1)
    mg.vcycle()
    sizes = O.level_sizes(81, 3)
    As = [A]
    Ps, Rs = [], []
    for l in range(1, 3):
        P = O.make_P(sizes[l - 1], sizes[l])
        R = P.transpose()
        Ps.append(P); Rs.append(R)
        As.append(O.galerkin(R, As[-1], P))
    u = [np.zeros(s) for s in sizes]
    f = [b.copy()] + [np.zeros(s) for s in sizes[1:]]
    for l in range(3):
        O.gs_smooth(As[l], u[l], f[l])
        r = O.residual(As[l], u[l], f[l])
        if l + 1 != 3:
            u[l + 1][:] = 0
            f[l + 1] = O.spmv(Rs[l], r)
    u[2] = O.Ldlt(As[2]).solve(f[2])
    for l in (1, 0):
        u[l] = u[l] + O.spmv(Ps[l], u[l + 1])
        O.gs_smooth(As[l], u[l], f[l])
    for l in range(3):
        np.testing.assert_array_equal(u[l], mg.u(l))


def test_new_smoothers_are_consistent():
    """Damped Jacobi / multicolour GS (oracle-defined): sanity against dense algebra."""
    A = O.laplacian(8)
    AT = A.transpose()
    b = O.rhs(8)
    dense = A.to_scipy().toarray()
    rng = np.random.default_rng(0)
    u = rng.standard_normal(64)
    un = O.jacobi_sweep(AT, u, b, 2.0 / 3.0)
    np.testing.assert_allclose(un, u + (2.0 / 3.0) * (b - dense @ u) / np.diag(dense), rtol=1e-13)
    nc, color = O.greedy_coloring(A, AT)
    assert nc == 2
    ii, jj = np.nonzero(dense - np.diag(np.diag(dense)))
    assert np.all(color[ii] != color[jj])
    v = u.copy()
    O.color_gs_pass(AT, color, 0, b, v)
    red = color == 0
    off = dense - np.diag(np.diag(dense))
    np.testing.assert_allclose(v[red], ((b - off @ u) / np.diag(dense))[red], rtol=1e-13)
    np.testing.assert_array_equal(v[~red], u[~red])


@pytest.mark.parametrize("n,eps", [(2, 1.0), (7, 1.0), (35, 1.0), (64, 1e-3)])
def test_numpy_five_point_restatement_matches_the_c_oracle(n, eps):
    """The vectorised full-size checker of the 8193^2 microbenchmark (oracle.five_point_*) is
    bit-identical to the C oracle's residual and damped-Jacobi sweep on the generated operator."""
    A = O.laplacian(n, eps)
    AT = A.transpose()
    u = np.random.default_rng(1).standard_normal(n * n)
    f = O.rhs(n)
    assert O.five_point_residual(n, eps, u, f).tobytes() == O.residual(A, u, f).tobytes()
    assert O.five_point_jacobi(n, eps, u, f, 0.6).tobytes() == O.jacobi_sweep(AT, u, f, 0.6).tobytes()

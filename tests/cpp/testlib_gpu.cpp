// C++ acceptance test of the include/amg mirror headers on the GPU.  It follows the
// reference's own test (jfdev001/algebraic-multigrid test/testlib.cpp:17-213) step by step,
// with its dense Jacobi / SOR smoothers (not on the V-cycle path) replaced by the smoothers
// this build adds.  Catch2 is not available here, so CHECK is a small macro.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>

#include <amg/common.hpp>
#include <amg/grid.hpp>
#include <amg/interpolator.hpp>
#include <amg/multigrid.hpp>
#include <amg/smoother.hpp>

static int g_checks = 0, g_failed = 0;
#define CHECK(cond)                                                          \
  do {                                                                       \
    ++g_checks;                                                              \
    if (!(cond)) {                                                           \
      ++g_failed;                                                            \
      std::printf("CHECK FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);    \
    }                                                                        \
  } while (0)
#define CHECK_THROWS_AS(expr, ex)                                            \
  do {                                                                       \
    ++g_checks;                                                              \
    bool thrown_ = false;                                                    \
    try {                                                                    \
      (void)(expr);                                                          \
    } catch (const ex&) {                                                    \
      thrown_ = true;                                                        \
    } catch (...) {                                                          \
    }                                                                        \
    if (!thrown_) {                                                          \
      ++g_failed;                                                            \
      std::printf("CHECK_THROWS FAILED %s:%d: %s\n", __FILE__, __LINE__, #expr); \
    }                                                                        \
  } while (0)

using Vec = AMG::VectorT<double>;
using Mat = AMG::SparseMatrixT<double>;

int main() {
  // testlib.cpp:18-28
  size_t n_interior_points = 2;
  Mat A = AMG::Grid<double>::laplacian(n_interior_points);
  Vec b = AMG::Grid<double>::rhs(n_interior_points);
  size_t ndofs = n_interior_points * n_interior_points;
  CHECK(b.size() == ndofs);

  // testlib.cpp:31-39: direct solution.  A one-level hierarchy is just the coarsest
  // direct solve (multigrid.hpp:287-288), the path's replacement for SimplicialLDLT.
  AMG::LinearInterpolator<double> one_level(1);
  AMG::SparseGaussSeidel<double> any_smoother;
  AMG::Multigrid<double> direct(&one_level, &any_smoother, A, b, 1, 1e-9, 1, 1);
  direct.vcycle();
  Vec exact_u = direct.get_soln(0);
  double rss = AMG::rss(A, exact_u, b);
  std::cout << "RSS: " << rss << std::endl;
  CHECK(rss < 1e-24);

  // testlib.cpp:58-62
  double h_from_n = AMG::Grid<double>::grid_spacing_h(n_interior_points);
  size_t n_from_h = AMG::Grid<double>::points_n_from_grid_spacing_h(h_from_n);
  CHECK(n_from_h == n_interior_points);

  // testlib.cpp:64-71 (relaxation-parameter validation), the reference's own lines
  double bad_omega_less_than_0 = -0.01;
  double bad_omega_greater_than_2 = 2.01;
  using bad_sor = AMG::SuccessiveOverRelaxation<double>;
  CHECK_THROWS_AS(bad_sor(bad_omega_less_than_0), std::invalid_argument);
  CHECK_THROWS_AS(bad_sor(bad_omega_greater_than_2), std::invalid_argument);
  CHECK_THROWS_AS(AMG::DampedJacobi<double>(-0.01), std::invalid_argument);
  CHECK_THROWS_AS(AMG::DampedJacobi<double>(2.01), std::invalid_argument);

  // testlib.cpp:73-115: the reference's host test smoothers (host-only classes of the mirror)
  size_t niters = 100;
  {
    Vec ref_jacobi_u(ndofs);
    ref_jacobi_u.setZero();
    AMG::Jacobi<double> ref_jacobi(niters);
    ref_jacobi.smooth(A, ref_jacobi_u, b);
    CHECK(ref_jacobi_u.isApprox(exact_u, ref_jacobi.tolerance));
    Vec sor_u(ndofs);
    sor_u.setZero();
    AMG::SuccessiveOverRelaxation<double> sor(niters);
    sor.smooth(A, sor_u, b);
    CHECK(sor_u.isApprox(exact_u, sor.tolerance));
    double tolerance = 1e-10;
    size_t compute_error_every_n_iters = 100;
    AMG::Jacobi<double> jacobi_base(tolerance, compute_error_every_n_iters, niters);
    AMG::SuccessiveOverRelaxation<double> sor_base(tolerance, compute_error_every_n_iters, niters);
    CHECK(jacobi_base.tolerance == tolerance && sor_base.n_iters == niters);
    // the GPU driver refuses host-only smoothers
    AMG::LinearInterpolator<double> two(2);
    CHECK_THROWS_AS(AMG::Multigrid<double>(&two, &sor, A, b, 2, 1e-9, 1, 1), std::invalid_argument);
  }

  // stale-mirror protection: values edited in place must not hit the cached device mirror
  {
    Mat A2 = AMG::Grid<double>::laplacian(n_interior_points);
    Vec zero_u(ndofs);
    zero_u.setZero();
    const double r_before = AMG::rss(A2, exact_u, b);
    for (long k = 0; k < A2.nonZeros(); ++k) A2.valuePtr()[k] *= 2.0;
    const double r_after = AMG::rss(A2, exact_u, b);  // (b - 2 A u)^2 = (b - 2 b)^2 = sum b^2
    const double bb = AMG::rss(A2, zero_u, b);
    CHECK(r_before < 1e-24);
    CHECK(std::fabs(r_after - bb) <= 1e-12 * bb);
  }

  // InterpolatorBase applies the STORED operators (interpolator.hpp:52-68): replace P by 2 P
  {
    AMG::LinearInterpolator<double> li(2);
    li.make_operators(7, 3, 0);
    Vec e(3);
    e[0] = 1.0; e[1] = -2.0; e[2] = 0.25;
    Vec lin = li.prolongation(e, 0);
    Mat P2 = li.get_P(0);
    for (long k = 0; k < P2.nonZeros(); ++k) P2.valuePtr()[k] *= 2.0;
    li.set_level_to_P(0, P2);
    Vec twice = li.prolongation(e, 0);
    bool ok = lin.size() == 7 && twice.size() == 7;
    for (size_t i = 0; ok && i < 7; ++i) ok = twice[i] == 2.0 * lin[i];
    CHECK(ok);
    CHECK(lin[0] == 0.5 && lin[1] == 1.0 && lin[2] == -0.5 && lin[6] == 0.125);
    Vec r(7);
    for (size_t i = 0; i < 7; ++i) r[i] = (double)(i + 1);
    Vec fr = li.restriction(r, 0);
    CHECK(fr.size() == 3 && fr[0] == 0.5 * 1 + 2 + 0.5 * 3 && fr[2] == 0.5 * 5 + 6 + 0.5 * 7);
  }

  // testlib.cpp:73-107 with the smoothers this build adds, as solvers on the 4-DOF system
  Vec jacobi_u(ndofs);
  jacobi_u.setZero();
  AMG::DampedJacobi<double> jacobi(1.0, niters);
  jacobi.smooth(A, jacobi_u, b);
  CHECK(jacobi_u.isApprox(exact_u, jacobi.tolerance));
  Vec color_u(ndofs);
  color_u.setZero();
  AMG::MulticolorGaussSeidel<double> color(niters);
  color.smooth(A, color_u, b);
  CHECK(color_u.isApprox(exact_u, color.tolerance));
  AMG::SparseGaussSeidel<double> spgs(niters);
  Vec spgs_u(ndofs);
  spgs_u.setZero();
  spgs.smooth(A, spgs_u, b);
  CHECK(spgs_u.isApprox(exact_u, spgs.tolerance));

  // testlib.cpp:117-128: interpolation operators (printed there; pinned here)
  size_t n_levels = 8;
  AMG::LinearInterpolator<double> linear_interpolator(n_levels);
  linear_interpolator.make_operators(7, 3, 0);
  {
    const Mat& P = linear_interpolator.get_P(0);
    CHECK(P.rows() == 7 && P.cols() == 3 && P.nonZeros() == 9);
    const int rows[9] = {0, 1, 2, 2, 3, 4, 4, 5, 6};
    const double vals[3] = {0.5, 1.0, 0.5};
    bool same = true;
    for (int k = 0; k < 9; ++k) same = same && P.innerIndexPtr()[k] == rows[k] && P.valuePtr()[k] == vals[k % 3];
    CHECK(same);
  }
  linear_interpolator.make_operators(24, 11, 0);
  {
    const Mat& P = linear_interpolator.get_P(0);
    CHECK(P.rows() == 24 && P.cols() == 11 && P.nonZeros() == 33);
    CHECK(P.innerIndexPtr()[32] == 22);  // row 23 is in no column
    const Mat& R = linear_interpolator.get_R(0);
    CHECK(R.rows() == 11 && R.cols() == 24 && R.outerIndexPtr()[24] - R.outerIndexPtr()[23] == 0);
  }

  // testlib.cpp:130-144: constructor validation
  using bad_amg = AMG::Multigrid<double>;
  size_t bad_compute_error_every_n_iters = 100;
  size_t bad_n_iters = 10;
  CHECK_THROWS_AS(bad_amg(&linear_interpolator, &spgs, A, b, n_levels, 1e-9, bad_compute_error_every_n_iters,
                          bad_n_iters),
                  std::invalid_argument);
  Mat bad_A(10, 10);
  Vec bad_b(11);
  CHECK_THROWS_AS(bad_amg(&linear_interpolator, &spgs, bad_A, bad_b, n_levels, 1e-9, 5, bad_n_iters),
                  std::invalid_argument);

  // testlib.cpp:146-181
  size_t n_fine_nodes = 35;
  Mat amg_A = AMG::Grid<double>::laplacian(n_fine_nodes);
  Vec amg_b = AMG::Grid<double>::rhs(n_fine_nodes);
  AMG::SparseGaussSeidel<double> amg_spgs;
  CHECK(amg_spgs.compute_error_every_n_iters == 0 && amg_spgs.n_iters == 1 && amg_spgs.tolerance == 1e-9);
  AMG::Multigrid<double> amg(&linear_interpolator, &amg_spgs, amg_A, amg_b, n_levels, 1e-9, 5, 100);
  std::cout << "Dofs at Levels in Multigrid:" << std::endl;
  std::cout << amg.get_coefficient_matrix(0).rows() << std::endl;
  const size_t golden_dofs[8] = {1225, 612, 305, 152, 75, 37, 18, 8};
  CHECK(amg.get_n_dofs(0) == golden_dofs[0]);
  for (size_t level = 1; level < n_levels; ++level) {
    const Mat& finer_A = amg.get_coefficient_matrix(level - 1);
    const Mat& coarser_A = amg.get_coefficient_matrix(level);
    std::cout << coarser_A.rows() << std::endl;
    CHECK((size_t)coarser_A.rows() == golden_dofs[level]);
    CHECK(finer_A.size() > coarser_A.size());
    CHECK(amg.get_n_dofs(level - 1) > amg.get_n_dofs(level));
    Vec finer_b = amg.get_rhs(level - 1);
    CHECK(finer_b.size() > amg.get_rhs(level).size());
  }

  // testlib.cpp:183-196
  AMG::SparseGaussSeidel<double> realistic_spgs(1e-9, 100, 1000);
  Mat A_h = AMG::Grid<double>::laplacian(n_fine_nodes);
  Vec rhs_h = AMG::Grid<double>::rhs(n_fine_nodes);
  Vec spgs_u_h(rhs_h.rows());
  spgs_u_h.setZero();
  realistic_spgs.smooth(A_h, spgs_u_h, rhs_h);
  double spgs_error = AMG::rss(A_h, spgs_u_h, rhs_h);
  std::cout << "SPGS error: " << spgs_error << std::endl;
  CHECK(spgs_error < realistic_spgs.tolerance);
  CHECK(realistic_spgs.iters_done == 900);  // README: "SPGS converged after 900 iterations."

  // testlib.cpp:198-212
  Vec amg_u = amg.solve();
  double amg_error = AMG::rss(A_h, amg_u, rhs_h);
  std::cout << "AMG error: " << amg_error << std::endl;
  CHECK(amg_error < amg.get_tolerance());
  CHECK(amg.iterations_done() == 35);  // README: "AMG converged after 35 iterations."
  CHECK(amg_u.isApprox(spgs_u_h, 1e-6));

  // grid transfers through the interpolator's non-virtual members (interpolator.hpp:52-68)
  {
    Vec r(1225), e(612);
    for (size_t i = 0; i < 1225; ++i) r[i] = 1.0;
    for (size_t i = 0; i < 612; ++i) e[i] = 1.0;
    Vec fc = linear_interpolator.restriction(r, 0);
    Vec pe = linear_interpolator.prolongation(e, 0);
    CHECK(fc.size() == 612 && fc[0] == 2.0 && fc[611] == 2.0);
    CHECK(pe.size() == 1225 && pe[0] == 0.5 && pe[1] == 1.0 && pe[2] == 1.0 && pe[1224] == 0.5);
  }

  std::printf("%d assertions, %d failed\n", g_checks, g_failed);
  return g_failed == 0 ? 0 : 1;
}

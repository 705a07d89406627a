// Mirror of include/amg/grid.hpp (AMG::Grid, :19-141): problem generators.  Inputs only;
// they run on the host inside libamgb.so and reproduce the reference's values bit for bit.
#pragma once
#include <functional>

#include "compat.hpp"

namespace AMG {

template <class EleType>
class Grid {
 public:
  static EleType grid_spacing_h(size_t n) { return amgb_grid_spacing_h((int64_t)n); }  // :31
  static size_t points_n_from_grid_spacing_h(EleType h = 1. / 50) {                    // :39-41
    return (size_t)amgb_points_n_from_grid_spacing_h(h);
  }
  // :88-98; eps_y scales the kron(D,I) coupling (1.0 = the reference operator)
  static SparseMatrixT<EleType> laplacian(size_t n, double eps_y = 1.0) {
    const int64_t N = (int64_t)n * (int64_t)n, nnz = amgb_grid_laplacian_nnz((int64_t)n);
    std::vector<int> outer(N + 1), inner(nnz);
    std::vector<double> val(nnz);
    detail::check(amgb_grid_laplacian((int64_t)n, eps_y, outer.data(), inner.data(), val.data()));
#if AMGB_HAVE_EIGEN
    SparseMatrixT<EleType> A = Eigen::Map<const Eigen::SparseMatrix<double>>(
        (int)N, (int)N, nnz, outer.data(), inner.data(), val.data());
    return A;
#else
    return SparseMatrixT<EleType>((int)N, (int)N, std::move(outer), std::move(inner), std::move(val));
#endif
  }
  // :108-140 with the default f; a custom f is evaluated on the host exactly like the reference
  static VectorT<EleType> rhs(size_t n) {
    VectorT<EleType> b(n * n);
    detail::check(amgb_grid_rhs((int64_t)n, b.data()));
    return b;
  }
  static VectorT<EleType> rhs(size_t n, std::function<EleType(EleType, EleType)> f) {
    VectorT<EleType> b(n * n);
    const EleType step = EleType(2.0) / EleType(n + 1);
    size_t dof = 0;
    for (size_t j = 1; j <= n; ++j) {
      const EleType xj = EleType(-1.0) + EleType(j) * step;
      for (size_t i = 1; i <= n; ++i) {
        const EleType xi = EleType(-1.0) + EleType(i) * step;
        b[dof++] = f(xj, xi);
      }
    }
    return b;
  }
};

}  // namespace AMG

// Mirror of include/amg/interpolator.hpp: InterpolatorBase (:15-87) keeps per-level P / R
// with the same getters / setters and a virtual make_operators; LinearInterpolator
// (:98-141) builds the reference's operators (bit-exact integer maps) through the C ABI.
// restriction() / prolongation() stay non-virtual members returning a new vector, as in the
// reference (:52-68); for LinearInterpolator they run matrix-free on the GPU.
#pragma once
#include <vector>

#include "compat.hpp"

namespace AMG {

template <class EleType>
class InterpolatorBase {
 private:
  std::vector<SparseMatrixT<EleType>> level_to_P;
  std::vector<SparseMatrixT<EleType>> level_to_R;

 public:
  InterpolatorBase(size_t n_levels) {  // interpolator.hpp:22-26
    level_to_P.resize(n_levels - 1);
    level_to_R.resize(n_levels - 1);
  }
  InterpolatorBase() {}
  virtual ~InterpolatorBase() {}

  virtual void make_operators(size_t n_h_dofs, size_t n_H_dofs, size_t level) = 0;
  // true when P is the reference's linear interpolation, which the device kernels apply
  // matrix-free; other interpolators are not supported by the GPU driver
  virtual bool is_linear_interpolation() const { return false; }

  VectorT<EleType> prolongation(const VectorT<EleType>& v, size_t level) {  // :52-56
    const auto& P = get_P(level);
    VectorT<EleType> result((size_t)P.rows());
    detail::check(amgb_linear_prolong((int64_t)P.rows(), (int64_t)P.cols(), v.data(), result.data()));
    return result;
  }
  VectorT<EleType> restriction(const VectorT<EleType>& v, size_t level) {  // :64-68
    const auto& R = get_R(level);
    VectorT<EleType> result((size_t)R.rows());
    detail::check(amgb_linear_restrict((int64_t)R.cols(), (int64_t)R.rows(), v.data(), result.data()));
    return result;
  }
  const SparseMatrixT<EleType>& get_P(size_t level) const { return level_to_P[level]; }
  const SparseMatrixT<EleType>& get_R(size_t level) const { return level_to_R[level]; }
  void set_level_to_P(size_t level, SparseMatrixT<EleType>& P) { level_to_P[level] = P; }
  void set_level_to_R(size_t level, SparseMatrixT<EleType>& R) { level_to_R[level] = R; }
};

template <class EleType>
class LinearInterpolator : public InterpolatorBase<EleType> {
 public:
  using InterpolatorBase<EleType>::InterpolatorBase;

  void make_operators(size_t n_h_dofs, size_t n_H_dofs, size_t level) override {  // :106-141
    const int64_t nnz = amgb_interp_nnz((int64_t)n_h_dofs, (int64_t)n_H_dofs);
    std::vector<int> Pc(n_H_dofs + 1), Pr(nnz), Rc(n_h_dofs + 1), Rr(nnz);
    std::vector<double> Pv(nnz), Rv(nnz);
    detail::check(amgb_interp_make_operators((int64_t)n_h_dofs, (int64_t)n_H_dofs, Pc.data(), Pr.data(),
                                             Pv.data(), Rc.data(), Rr.data(), Rv.data()));
#if AMGB_HAVE_EIGEN
    SparseMatrixT<EleType> P = Eigen::Map<const Eigen::SparseMatrix<double>>(
        (int)n_h_dofs, (int)n_H_dofs, nnz, Pc.data(), Pr.data(), Pv.data());
    SparseMatrixT<EleType> R = Eigen::Map<const Eigen::SparseMatrix<double>>(
        (int)n_H_dofs, (int)n_h_dofs, nnz, Rc.data(), Rr.data(), Rv.data());
#else
    SparseMatrixT<EleType> P((int)n_h_dofs, (int)n_H_dofs, std::move(Pc), std::move(Pr), std::move(Pv));
    SparseMatrixT<EleType> R((int)n_H_dofs, (int)n_h_dofs, std::move(Rc), std::move(Rr), std::move(Rv));
#endif
    this->set_level_to_P(level, P);
    this->set_level_to_R(level, R);
  }
  bool is_linear_interpolation() const override { return true; }
};

}  // namespace AMG

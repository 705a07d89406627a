// Mirror of include/amg/smoother.hpp: SmootherBase (:18-66) and SparseGaussSeidel
// (:86-216) with the same public fields, constructors and smooth() signature; smooth()
// forwards the CSC arrays and the vectors to the GPU through the C ABI (amgb.h).
// DampedJacobi and MulticolorGaussSeidel are the partitionable smoothers the B200 build
// adds; the reference's dense O(N^2) Jacobi / SuccessiveOverRelaxation test smoothers
// (:223-373) are not part of the V-cycle path and are not mirrored.
#pragma once
#include <iostream>
#include <string>

#include "common.hpp"

namespace AMG {

template <class EleType>
class SmootherBase {
 public:
  EleType tolerance{1e-9};
  size_t compute_error_every_n_iters{100};
  size_t n_iters{1};

  SmootherBase() {}
  SmootherBase(size_t n_iters_) : n_iters(n_iters_) {}
  SmootherBase(double tolerance_, size_t compute_error_every_n_iters_, size_t n_iters_)
      : tolerance(tolerance_), compute_error_every_n_iters(compute_error_every_n_iters_), n_iters(n_iters_) {}
  virtual ~SmootherBase() {}

  virtual void smooth(const SparseMatrixT<EleType>& A, VectorT<EleType>& u, const VectorT<EleType>& b) = 0;

  // Which device smoother the GPU Multigrid driver runs for this object
  // (AMGB_SMOOTHER_*), or -1 for a host-only user smoother the driver cannot use.
  virtual int device_kind() const { return -1; }
  virtual double device_omega() const { return 2.0 / 3.0; }
};

template <class EleType>
class SparseGaussSeidel : public SmootherBase<EleType> {
 public:
  using SmootherBase<EleType>::SmootherBase;
  size_t iters_done{0};     // beyond the reference, which only prints these
  EleType last_error{100};
  int mode{AMGB_GS_AUTO};

  SparseGaussSeidel() {  // smoother.hpp:183-187
    this->tolerance = 1e-9;
    this->compute_error_every_n_iters = 0;
    this->n_iters = 1;
  }
  void smooth(const SparseMatrixT<EleType>& A, VectorT<EleType>& u, const VectorT<EleType>& b) override {
    int64_t it = 0;
    double err = 100;
    detail::check(amgb_smooth_gs(detail::MirrorCache::instance().get(A), u.data(), b.data(), this->tolerance,
                                 (int64_t)this->compute_error_every_n_iters, (int64_t)this->n_iters, mode, &it,
                                 &err));
    iters_done = (size_t)it;
    last_error = err;
    if (this->compute_error_every_n_iters != 0) {  // smoother.hpp:205-212
      if (err <= this->tolerance)
        std::cout << "SPGS converged after " << it << " iterations." << std::endl;
      else
        std::cout << "SPGS did not converge after " << it << " iterations." << std::endl;
    }
  }
  int device_kind() const override { return AMGB_SMOOTHER_GS; }
};

// u <- u + omega D^-1 (f - A u); n_iters sweeps per smooth() call.
template <class EleType>
class DampedJacobi : public SmootherBase<EleType> {
  double omega_{2.0 / 3.0};

 public:
  DampedJacobi(double omega = 2.0 / 3.0, size_t n_sweeps = 2) : SmootherBase<EleType>(n_sweeps), omega_(omega) {
    if (omega <= 0 || omega > 2)
      throw std::invalid_argument("`omega` must be in (0, 2] but got omega=" + std::to_string(omega));
  }
  void smooth(const SparseMatrixT<EleType>& A, VectorT<EleType>& u, const VectorT<EleType>& b) override {
    detail::check(amgb_smooth_jacobi(detail::MirrorCache::instance().get(A), u.data(), b.data(), omega_,
                                     (int64_t)this->n_iters));
  }
  int device_kind() const override { return AMGB_SMOOTHER_JACOBI; }
  double device_omega() const override { return omega_; }
};

// greedy multicolour symmetric Gauss-Seidel (red-black on the five-point level)
template <class EleType>
class MulticolorGaussSeidel : public SmootherBase<EleType> {
 public:
  MulticolorGaussSeidel(size_t n_iters_ = 1) : SmootherBase<EleType>(n_iters_) {}
  void smooth(const SparseMatrixT<EleType>& A, VectorT<EleType>& u, const VectorT<EleType>& b) override {
    detail::check(amgb_smooth_color_gs(detail::MirrorCache::instance().get(A), u.data(), b.data(),
                                       (int64_t)this->n_iters));
  }
  int device_kind() const override { return AMGB_SMOOTHER_COLOR_GS; }
};

}  // namespace AMG

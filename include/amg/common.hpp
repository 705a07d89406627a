// Mirror of jfdev001/algebraic-multigrid include/amg/common.hpp (rss, :17-27): the sum of
// squared residuals is evaluated on the GPU through amgb_rss (one fused SpMV + reduction;
// the reference's accidental O(N*nnz) evaluation is not reproduced).
#pragma once
#include "compat.hpp"

namespace AMG {

template <class EleType>
EleType rss(const SparseMatrixT<EleType>& A, const VectorT<EleType>& u, const VectorT<EleType>& b) {
  static_assert(sizeof(EleType) == sizeof(double), "the device path is fp64 only");
  double out = 0.0;
  detail::check(amgb_rss(detail::MirrorCache::instance().get(A), u.data(), b.data(), &out));
  return out;
}

}  // namespace AMG

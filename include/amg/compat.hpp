// Matrix / vector types for the include/amg mirror headers.
//
// With Eigen on the include path these are exactly the reference's types
// (Eigen::SparseMatrix<T> ColMajor with int indices, Eigen::Matrix<T,-1,1>), so code
// written against jfdev001/algebraic-multigrid's headers compiles unchanged.  Without
// Eigen (this build image has none) a minimal stand-in with the same accessor names
// (rows, cols, nonZeros, outerIndexPtr, innerIndexPtr, valuePtr, isCompressed / size,
// data, operator[], setZero, isApprox) is used, enough for the V-cycle path and for
// tests/cpp/testlib_gpu.cpp.  Only these accessors are used by the other headers.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../amgb.h"

#if defined(AMGB_USE_EIGEN) || (__has_include(<Eigen/Sparse>) && !defined(AMGB_NO_EIGEN))
#include <Eigen/Core>
#include <Eigen/Sparse>
namespace AMG {
template <class T>
using SparseMatrixT = Eigen::SparseMatrix<T>;
template <class T>
using VectorT = Eigen::Matrix<T, -1, 1>;
}  // namespace AMG
#define AMGB_HAVE_EIGEN 1
#else
namespace amgb {
template <class T>
class Vector {
  std::vector<T> v_;

 public:
  Vector() {}
  explicit Vector(std::size_t n) : v_(n) {}
  std::size_t size() const { return v_.size(); }
  std::size_t rows() const { return v_.size(); }
  void resize(std::size_t n) { v_.resize(n); }
  void setZero() { v_.assign(v_.size(), T(0)); }
  T* data() { return v_.data(); }
  const T* data() const { return v_.data(); }
  T& operator[](std::size_t i) { return v_[i]; }
  const T& operator[](std::size_t i) const { return v_[i]; }
  // Eigen's DenseBase::isApprox: ||a-b||^2 <= p^2 * min(||a||^2, ||b||^2)
  bool isApprox(const Vector& o, T prec) const {
    if (o.size() != size()) return false;
    T d = 0, a = 0, b = 0;
    for (std::size_t i = 0; i < size(); ++i) {
      d += (v_[i] - o.v_[i]) * (v_[i] - o.v_[i]);
      a += v_[i] * v_[i];
      b += o.v_[i] * o.v_[i];
    }
    return d <= prec * prec * (a < b ? a : b);
  }
};

// Compressed ColMajor sparse matrix with int indices (Eigen::SparseMatrix<T> layout).
template <class T>
class SparseMatrix {
  int rows_ = 0, cols_ = 0;
  std::vector<int> outer_{0};
  std::vector<int> inner_;
  std::vector<T> val_;

 public:
  SparseMatrix() {}
  SparseMatrix(int rows, int cols) : rows_(rows), cols_(cols), outer_((std::size_t)cols + 1, 0) {}
  SparseMatrix(int rows, int cols, std::vector<int> outer, std::vector<int> inner, std::vector<T> val)
      : rows_(rows), cols_(cols), outer_(std::move(outer)), inner_(std::move(inner)), val_(std::move(val)) {}
  int rows() const { return rows_; }
  int cols() const { return cols_; }
  long size() const { return (long)rows_ * cols_; }
  long nonZeros() const { return (long)inner_.size(); }
  bool isCompressed() const { return true; }
  const int* outerIndexPtr() const { return outer_.data(); }
  const int* innerIndexPtr() const { return inner_.data(); }
  const T* valuePtr() const { return val_.data(); }
  int* outerIndexPtr() { return outer_.data(); }
  int* innerIndexPtr() { return inner_.data(); }
  T* valuePtr() { return val_.data(); }
  void resizeNonZeros(std::size_t nnz) {
    inner_.resize(nnz);
    val_.resize(nnz);
  }
};
}  // namespace amgb
namespace AMG {
template <class T>
using SparseMatrixT = amgb::SparseMatrix<T>;
template <class T>
using VectorT = amgb::Vector<T>;
}  // namespace AMG
#define AMGB_HAVE_EIGEN 0
#endif

namespace AMG {
namespace detail {
// C-ABI status -> the exceptions the reference throws (std::invalid_argument for the
// validated constructor conditions), std::runtime_error for CUDA / NCCL failures.
inline void check(int code) {
  if (code == AMGB_OK) return;
  const std::string msg = amgb_last_error();
  if (code == AMGB_EINVAL) throw std::invalid_argument(msg);
  throw std::runtime_error("amgb error " + std::to_string(code) + ": " + msg);
}

// Device mirror of a host matrix, cached by (value pointer, nnz, rows) AND a content
// fingerprint that is re-evaluated on every use, so repeated smooth()/rss() calls with the same
// matrix do not upload it again while values edited in place (coeffRef, A *= c) or a new
// matrix allocated at a recycled address are detected and uploaded afresh -- the reference
// always reads the live matrix.  The fingerprint hashes every index and value of matrices
// with at most 4 M entries; above that it hashes 65536 evenly spaced entries of each array
// (a cheap O(1) check next to the O(N) vector upload of the same call), so after editing
// single entries of a LARGE matrix in place call AMG::invalidate_device_mirror(A).
class MirrorCache {
  struct Entry {
    const void* key;
    long nnz;
    int rows;
    unsigned long long fingerprint;
    amgb_matrix* m;
  };
  std::vector<Entry> entries_;

  static unsigned long long mix(unsigned long long h, const void* p, std::size_t bytes) {
    const unsigned char* c = static_cast<const unsigned char*>(p);
    std::size_t i = 0;
    for (; i + 8 <= bytes; i += 8) {
      unsigned long long w;
      std::memcpy(&w, c + i, 8);
      h = (h ^ w) * 0x9E3779B97F4A7C15ull;
      h ^= h >> 29;
    }
    for (; i < bytes; ++i) h = (h ^ c[i]) * 0x100000001B3ull;
    return h;
  }
  template <class Mat>
  static unsigned long long fingerprint(const Mat& A) {
    const std::size_t nnz = (std::size_t)A.nonZeros(), nc = (std::size_t)A.cols() + 1;
    unsigned long long h = 0xCBF29CE484222325ull ^ (unsigned long long)nnz;
    if (nnz <= (std::size_t(1) << 22)) {
      h = mix(h, A.outerIndexPtr(), nc * sizeof(int));
      h = mix(h, A.innerIndexPtr(), nnz * sizeof(int));
      h = mix(h, A.valuePtr(), nnz * sizeof(*A.valuePtr()));
      return h;
    }
    const std::size_t step = nnz / 65536, cstep = nc / 65536 + 1;
    for (std::size_t i = 0; i < nc; i += cstep) h = mix(h, A.outerIndexPtr() + i, sizeof(int));
    for (std::size_t i = 0; i < nnz; i += step) {
      h = mix(h, A.innerIndexPtr() + i, sizeof(int));
      h = mix(h, A.valuePtr() + i, sizeof(*A.valuePtr()));
    }
    return h;
  }

 public:
  ~MirrorCache() {
    for (auto& e : entries_) amgb_matrix_destroy(e.m);
  }
  template <class Mat>
  amgb_matrix* get(const Mat& A) {
    if (!A.isCompressed()) throw std::invalid_argument("matrix must be compressed (makeCompressed())");
    const unsigned long long fp = fingerprint(A);
    for (std::size_t i = 0; i < entries_.size(); ++i) {
      Entry& e = entries_[i];
      if (e.key != A.valuePtr() || e.nnz != (long)A.nonZeros() || e.rows != (int)A.rows()) continue;
      if (e.fingerprint == fp) return e.m;
      amgb_matrix_destroy(e.m);  // same storage, different content: the mirror is stale
      entries_.erase(entries_.begin() + (long)i);
      break;
    }
    amgb_matrix* m = nullptr;
    check(amgb_matrix_create((int)A.rows(), (int)A.cols(), A.outerIndexPtr(), A.innerIndexPtr(),
                             A.valuePtr(), &m));
    if (entries_.size() >= 8) {
      amgb_matrix_destroy(entries_.front().m);
      entries_.erase(entries_.begin());
    }
    entries_.push_back({A.valuePtr(), (long)A.nonZeros(), (int)A.rows(), fp, m});
    return m;
  }
  // drop the mirror of A (or every mirror) explicitly
  template <class Mat>
  void invalidate(const Mat& A) {
    for (std::size_t i = 0; i < entries_.size(); ++i)
      if (entries_[i].key == A.valuePtr()) {
        amgb_matrix_destroy(entries_[i].m);
        entries_.erase(entries_.begin() + (long)i);
        return;
      }
  }
  void clear() {
    for (auto& e : entries_) amgb_matrix_destroy(e.m);
    entries_.clear();
  }
  static MirrorCache& instance() {
    static thread_local MirrorCache c;
    return c;
  }
};
}  // namespace detail

// Call after changing the values of a large matrix in place (see MirrorCache).
template <class Mat>
inline void invalidate_device_mirror(const Mat& A) {
  detail::MirrorCache::instance().invalidate(A);
}
}  // namespace AMG

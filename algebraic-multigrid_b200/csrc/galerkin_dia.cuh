// Galerkin coarse operator A_H = R (A P) for the reference's linear interpolation, evaluated
// row by row on the diagonal (DIA) layout -- the building block of a device-side setup
// (reference: include/amg/multigrid.hpp:219-223 with P, R from
// include/amg/interpolator.hpp:106-141).
//
// P (n_f x n_c) has, in column J, rows 2J, 2J+1, 2J+2 (those below n_f) with 0.5, 1, 0.5 and
// R = P^T, so coarse entry (I, J) only involves the fine entries A(i, k) with i in {2I, 2I+1, 2I+2}
// and k in {2J, 2J+1, 2J+2}: fine offsets 2 (J - I) - 2 .. 2 (J - I) + 2.  Eigen's conservative
// sparse product evaluates T = A P first -- T(i, J) accumulates A(i,k) P(k,J) over ascending k --
// and then R T -- A_H(I, J) accumulates R(I,i) T(i,J) over ascending i; coarse_entry() keeps
// exactly that order, so its values are bit-identical to host_setup.cpp's galerkin() (and to the
// oracle) for every entry.  Entries of A that are absent or explicit zeros contribute +0 in
// either formulation.  Every coarse entry is independent: one thread per coarse row on the GPU.
//
// Status: host-tested against the oracle's triple product (tests/test_galerkin_dia_host.py);
// not yet wired into amgb_hierarchy_create, which still builds the hierarchy on the host.
#pragma once

#if defined(__CUDACC__)
#define AMGB_GAL_FN __host__ __device__ __forceinline__
#else
#define AMGB_GAL_FN inline
#endif

namespace amgb {
namespace gal {

constexpr int kMaxDiag = 16;

// products and sums exactly as written (no FMA contraction on the device)
AMGB_GAL_FN double gmul(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
AMGB_GAL_FN double gadd(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}

// Fine operator in DIA: val[d * ld + row], offsets ascending.
struct FineDia {
  int n;   // rows
  int nd;  // diagonals
  int ld;
  int off[kMaxDiag];
  const double* val;
};

// A(i, k) of the fine operator (0 when outside the band / the matrix)
AMGB_GAL_FN double fine_entry(const FineDia& A, int i, int k) {
  if (i < 0 || i >= A.n || k < 0 || k >= A.n) return 0.0;
  const int o = k - i;
  for (int d = 0; d < A.nd; ++d)
    if (A.off[d] == o) return A.val[(long long)d * A.ld + i];
  return 0.0;
}

// Coarse offsets a fine offset set produces: c such that some fine offset lies in [2c-2, 2c+2].
// Returns their number (ascending in off_c); -1 if more than kMaxDiag.
inline int coarse_offsets(int nd_f, const int* off_f, int* off_c) {
  int n = 0;
  if (nd_f < 1) return 0;
  const int lo = (off_f[0] - 2) / 2 - 2, hi = (off_f[nd_f - 1] + 2) / 2 + 2;
  for (int c = lo; c <= hi; ++c) {
    bool hit = false;
    for (int d = 0; d < nd_f && !hit; ++d) hit = (off_f[d] >= 2 * c - 2 && off_f[d] <= 2 * c + 2);
    if (hit) {
      if (n == kMaxDiag) return -1;
      off_c[n++] = c;
    }
  }
  return n;
}

// A_H(I, J) in Eigen's evaluation order.
AMGB_GAL_FN double coarse_entry(const FineDia& A, int n_c, int I, int J) {
  if (I < 0 || I >= n_c || J < 0 || J >= n_c) return 0.0;
  double acc = 0.0;
  for (int di = 0; di < 3; ++di) {  // R(I, i) = P(i, I), ascending i
    const int i = 2 * I + di;
    if (i >= A.n) break;
    const double w = (di == 1) ? 1.0 : 0.5;
    double t = 0.0;                   // T(i, J) = sum over ascending k of A(i, k) P(k, J)
    for (int dk = 0; dk < 3; ++dk) {
      const int k = 2 * J + dk;
      if (k >= A.n) break;
      const double a = fine_entry(A, i, k);
      if (a != 0.0) t = gadd(t, gmul(a, (dk == 1) ? 1.0 : 0.5));
    }
    if (t != 0.0) acc = gadd(acc, gmul(w, t));
  }
  return acc;
}

// One coarse row: out[c] = A_H(I, I + off_c[c]).
AMGB_GAL_FN void coarse_row(const FineDia& A, int n_c, int nd_c, const int* off_c, int I, double* out,
                            int out_stride) {
  for (int c = 0; c < nd_c; ++c) out[(long long)c * out_stride] = coarse_entry(A, n_c, I, I + off_c[c]);
}

}  // namespace gal
}  // namespace amgb

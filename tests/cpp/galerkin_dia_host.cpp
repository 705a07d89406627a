// TEST INFRASTRUCTURE ONLY -- never linked into libamgb.so.
// Host build of algebraic-multigrid_b200/csrc/galerkin_dia.cuh (the row-wise Galerkin product on
// the DIA layout, building block of a device-side setup) for tests/test_galerkin_dia_host.py.
#include "../../algebraic-multigrid_b200/csrc/galerkin_dia.cuh"

extern "C" {

// Fine operator: n_f rows, nd_f ascending offsets, val_f[d * ld_f + row].  Writes the coarse
// offsets to off_c (capacity 16), their number to *nd_c and the coarse values to
// val_c[c * ld_c + I].  Returns 0, or 1 if the coarse operator has more than 16 diagonals.
int gal_host_coarse(int n_f, int nd_f, const int* off_f, int ld_f, const double* val_f, int n_c, int* off_c,
                    int* nd_c, int ld_c, double* val_c) {
  amgb::gal::FineDia A;
  A.n = n_f;
  A.nd = nd_f;
  A.ld = ld_f;
  for (int d = 0; d < nd_f; ++d) A.off[d] = off_f[d];
  A.val = val_f;
  const int n = amgb::gal::coarse_offsets(nd_f, off_f, off_c);
  if (n < 0) return 1;
  *nd_c = n;
  for (int I = 0; I < n_c; ++I) amgb::gal::coarse_row(A, n_c, n, off_c, I, val_c + I, ld_c);
  return 0;
}
}

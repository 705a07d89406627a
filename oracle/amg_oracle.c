/*
 * amg_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C, Eigen-free restatement of the arithmetic of the V-cycle hot path
 * of jfdev001/algebraic-multigrid.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library, and
 * only as the checker / reported CPU baseline.  The product (libamgb.so) never
 * links or calls it.
 *
 * Parity status: PINNED for the reference path (symmetric Gauss-Seidel
 * V-cycle) by the reference's golden stdout (README screenshot, SURVEY.md
 * section 6): level sizes 1225..8, 35 cycles / 7.19199e-11, 900 iterations /
 * 8.69692e-10 -- see tests/test_oracle_golden.py.  The reference itself cannot
 * be compiled here (needs Eigen 3.4.0 + Catch2, both absent, no network).
 * "parity unpinned" for the smoothers the reference does not have (damped
 * Jacobi, multicolour Gauss-Seidel) and for the anisotropic generator: their
 * definition below IS the specification.
 *
 * Every function cites the reference file:line it follows
 * (paths relative to /root/reference).  Floating point: no FMA contraction
 * (compile with -ffp-contract=off), sums in the exact order Eigen 3.4.0 uses.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int rows, cols;
  int64_t nnz;
  int* colptr; /* cols+1 */
  int* rowidx; /* nnz, ascending inside a column */
  double* val; /* nnz */
} orc_csc;

static void* xmalloc(size_t n) {
  void* p = malloc(n ? n : 1);
  if (!p) {
    fprintf(stderr, "oracle: out of memory (%zu bytes)\n", n);
    abort();
  }
  return p;
}

static orc_csc* csc_alloc(int rows, int cols, int64_t nnz) {
  orc_csc* m = (orc_csc*)xmalloc(sizeof(orc_csc));
  m->rows = rows;
  m->cols = cols;
  m->nnz = nnz;
  m->colptr = (int*)xmalloc(sizeof(int) * ((size_t)cols + 1));
  m->rowidx = (int*)xmalloc(sizeof(int) * (size_t)nnz);
  m->val = (double*)xmalloc(sizeof(double) * (size_t)nnz);
  return m;
}

void orc_csc_free(orc_csc* m) {
  if (!m) return;
  free(m->colptr);
  free(m->rowidx);
  free(m->val);
  free(m);
}
int orc_csc_rows(const orc_csc* m) { return m->rows; }
int orc_csc_cols(const orc_csc* m) { return m->cols; }
int64_t orc_csc_nnz(const orc_csc* m) { return m->nnz; }
void orc_csc_copy_out(const orc_csc* m, int* colptr, int* rowidx, double* val) {
  memcpy(colptr, m->colptr, sizeof(int) * ((size_t)m->cols + 1));
  memcpy(rowidx, m->rowidx, sizeof(int) * (size_t)m->nnz);
  memcpy(val, m->val, sizeof(double) * (size_t)m->nnz);
}
orc_csc* orc_csc_from_arrays(int rows, int cols, const int* colptr,
                             const int* rowidx, const double* val) {
  int64_t nnz = colptr[cols];
  orc_csc* m = csc_alloc(rows, cols, nnz);
  memcpy(m->colptr, colptr, sizeof(int) * ((size_t)cols + 1));
  memcpy(m->rowidx, rowidx, sizeof(int) * (size_t)nnz);
  memcpy(m->val, val, sizeof(double) * (size_t)nnz);
  return m;
}
/* number of stored entries whose value is not exactly 0.0 */
int64_t orc_csc_nnz_nonzero(const orc_csc* m) {
  int64_t c = 0;
  for (int64_t p = 0; p < m->nnz; ++p) c += (m->val[p] != 0.0);
  return c;
}

/* ------------------------------------------------------------------ */
/* Problem generators (inputs only)                                   */
/* ------------------------------------------------------------------ */

/* include/amg/grid.hpp:31  grid_spacing_h(n) = 2.0/(n+1) */
double orc_grid_spacing_h(int n) { return 2.0 / (double)((size_t)n + 1); }

/* include/amg/grid.hpp:39-41 */
int64_t orc_points_n_from_grid_spacing_h(double h) {
  return (int64_t)((2 / h) - 1);
}

/*
 * include/amg/grid.hpp:50-98.  D = tridiag(1,-2,1)/(h*h) (each stored value
 * divided by (h*h), :72), A = kron(I,D) + eps_y * kron(D,I).  DOF k = j*n+i.
 * kron(I,D) couples k+-1 inside a grid line, kron(D,I) couples k+-n.
 * eps_y == 1.0 gives exactly Grid::laplacian (1.0*x == x).  The anisotropic
 * variant (eps_y != 1) has no reference counterpart (oracle-defined,
 * eps multiplies kron(D,I)).
 */
orc_csc* orc_laplacian_aniso(int n, double eps_y) {
  double h = orc_grid_spacing_h(n);
  double off = 1.0 / (h * h);
  double dia = -2.0 / (h * h);
  int64_t N = (int64_t)n * n;
  int64_t nnz = 5 * N - 4 * (int64_t)n;
  orc_csc* A = csc_alloc((int)N, (int)N, nnz);
  int64_t p = 0;
  for (int j = 0; j < n; ++j) {
    for (int i = 0; i < n; ++i) {
      int64_t k = (int64_t)j * n + i;
      A->colptr[k] = (int)p;
      if (j > 0) { A->rowidx[p] = (int)(k - n); A->val[p++] = eps_y * off; }
      if (i > 0) { A->rowidx[p] = (int)(k - 1); A->val[p++] = off; }
      A->rowidx[p] = (int)k;
      A->val[p++] = dia + eps_y * dia; /* kron(I,D)(k,k) + kron(D,I)(k,k) */
      if (i < n - 1) { A->rowidx[p] = (int)(k + 1); A->val[p++] = off; }
      if (j < n - 1) { A->rowidx[p] = (int)(k + n); A->val[p++] = eps_y * off; }
    }
  }
  A->colptr[N] = (int)p;
  return A;
}
orc_csc* orc_laplacian(int n) { return orc_laplacian_aniso(n, 1.0); }

/*
 * include/amg/grid.hpp:108-140.  LinSpaced(n+2,-1,1)[i] = low + i*step,
 * step = (high-low)/(size-1) (Eigen 3.4.0 linspaced_op, |high| !< |low| so no
 * flip; the forced last element is a boundary point and never sampled).
 * b[dof] = 5*exp(-10*(xj*xj + xi*xi)), j outer, i inner (:129-137).
 */
void orc_rhs(int n, double* b) {
  double step = (1.0 - (-1.0)) / (double)((size_t)n + 2 - 1);
  int64_t dof = 0;
  for (int j = 1; j <= n; ++j) {
    double xj = -1.0 + (double)j * step;
    for (int i = 1; i <= n; ++i) {
      double xi = -1.0 + (double)i * step;
      b[dof++] = 5 * exp(-10 * (xj * xj + xi * xi));
    }
  }
}

/* ------------------------------------------------------------------ */
/* Setup: interpolation operators and Galerkin product                */
/* ------------------------------------------------------------------ */

/* include/amg/multigrid.hpp:127-130 */
int64_t orc_n_H_dofs_from_n_h_dofs(int64_t h_dofs) {
  return (int64_t)(((size_t)h_dofs + 1) / 2 - 1);
}

/*
 * include/amg/interpolator.hpp:106-130.  P is n_h x n_H, column j holds
 * rows i=2j (0.5), i+1 (1.0), i+2 (0.5), each only if inside [0,n_h).
 */
orc_csc* orc_make_P(int n_h, int n_H) {
  int64_t cnt = 0;
  for (int j = 0; j < n_H; ++j) {
    int64_t i = 2 * (int64_t)j;
    cnt += (i < n_h) + (i + 1 < n_h) + (i + 2 < n_h);
  }
  orc_csc* P = csc_alloc(n_h, n_H, cnt);
  int64_t p = 0;
  for (int j = 0; j < n_H; ++j) {
    int64_t i = 2 * (int64_t)j;
    P->colptr[j] = (int)p;
    if (i < n_h) { P->rowidx[p] = (int)i; P->val[p++] = 0.5; }
    if (i + 1 < n_h) { P->rowidx[p] = (int)(i + 1); P->val[p++] = 1.0; }
    if (i + 2 < n_h) { P->rowidx[p] = (int)(i + 2); P->val[p++] = 0.5; }
  }
  P->colptr[n_H] = (int)p;
  return P;
}

/* include/amg/interpolator.hpp:132-134  R = P.transpose() (sorted CSC). */
orc_csc* orc_transpose(const orc_csc* A) {
  orc_csc* T = csc_alloc(A->cols, A->rows, A->nnz);
  int* cnt = (int*)calloc((size_t)A->rows + 1, sizeof(int));
  for (int64_t p = 0; p < A->nnz; ++p) cnt[A->rowidx[p] + 1]++;
  for (int r = 0; r < A->rows; ++r) cnt[r + 1] += cnt[r];
  memcpy(T->colptr, cnt, sizeof(int) * ((size_t)A->rows + 1));
  for (int c = 0; c < A->cols; ++c)
    for (int p = A->colptr[c]; p < A->colptr[c + 1]; ++p) {
      int q = cnt[A->rowidx[p]]++;
      T->rowidx[q] = c;
      T->val[q] = A->val[p];
    }
  free(cnt);
  return T;
}

/*
 * Eigen 3.4.0 conservative sparse*sparse product (ColMajor x ColMajor), the
 * kernel behind include/amg/multigrid.hpp:219-223.  For result column j, walk
 * rhs(:,j) in ascending k; for every lhs(i,k): first hit assigns x*y, later
 * hits += x*y.  Numerical zeros are NOT pruned; result columns are sorted.
 */
orc_csc* orc_spgemm(const orc_csc* L, const orc_csc* Rr) {
  int rows = L->rows, cols = Rr->cols;
  char* mask = (char*)calloc((size_t)rows, 1);
  double* acc = (double*)xmalloc(sizeof(double) * (size_t)rows);
  int* idx = (int*)xmalloc(sizeof(int) * (size_t)rows);
  /* pass 1: count */
  int64_t total = 0;
  for (int j = 0; j < cols; ++j) {
    int n = 0;
    for (int q = Rr->colptr[j]; q < Rr->colptr[j + 1]; ++q) {
      int k = Rr->rowidx[q];
      for (int p = L->colptr[k]; p < L->colptr[k + 1]; ++p) {
        int i = L->rowidx[p];
        if (!mask[i]) { mask[i] = 1; idx[n++] = i; }
      }
    }
    for (int t = 0; t < n; ++t) mask[idx[t]] = 0;
    total += n;
  }
  if (total > 2147483647LL) {
    fprintf(stderr, "oracle: product nnz overflows int32\n");
    abort();
  }
  orc_csc* C = csc_alloc(rows, cols, total);
  int64_t out = 0;
  for (int j = 0; j < cols; ++j) {
    int n = 0;
    C->colptr[j] = (int)out;
    for (int q = Rr->colptr[j]; q < Rr->colptr[j + 1]; ++q) {
      int k = Rr->rowidx[q];
      double y = Rr->val[q];
      for (int p = L->colptr[k]; p < L->colptr[k + 1]; ++p) {
        int i = L->rowidx[p];
        double x = L->val[p];
        if (!mask[i]) {
          mask[i] = 1;
          acc[i] = x * y;
          idx[n++] = i;
        } else {
          acc[i] += x * y;
        }
      }
    }
    /* sort the (short) index list ascending */
    for (int a = 1; a < n; ++a) {
      int v = idx[a], b = a - 1;
      while (b >= 0 && idx[b] > v) { idx[b + 1] = idx[b]; --b; }
      idx[b + 1] = v;
    }
    for (int t = 0; t < n; ++t) {
      C->rowidx[out] = idx[t];
      C->val[out++] = acc[idx[t]];
      mask[idx[t]] = 0;
    }
  }
  C->colptr[cols] = (int)out;
  free(mask);
  free(acc);
  free(idx);
  return C;
}

/* include/amg/multigrid.hpp:219-223  A_H = R_h * (A_h * P_h) */
orc_csc* orc_galerkin(const orc_csc* R, const orc_csc* A, const orc_csc* P) {
  orc_csc* T = orc_spgemm(A, P);
  orc_csc* AH = orc_spgemm(R, T);
  orc_csc_free(T);
  return AH;
}

/* ------------------------------------------------------------------ */
/* Hot-path operators                                                 */
/* ------------------------------------------------------------------ */

/*
 * include/amg/smoother.hpp:101-138 (matvecprod + update): column `col` of the
 * CSC matrix is used as if it were row `col`.  rsum starts at +0.0, adds
 * val*u[row] in ascending row order, adds literal 0 for the diagonal entry;
 * u[col] = (b[col]-rsum)/diag unless diag == 0.
 */
static inline void gs_update(const orc_csc* A, const double* b, double* u,
                             int col) {
  double z = 0;
  double rsum = z, diag = z;
  for (int p = A->colptr[col]; p < A->colptr[col + 1]; ++p) {
    int row = A->rowidx[p];
    double val = A->val[p];
    diag = (col == row) ? val : diag;
    rsum += (col == row) ? z : val * u[row];
  }
  u[col] = (diag == z) ? u[col] : (b[col] - rsum) / diag;
}
/* include/amg/smoother.hpp:148-157 */
void orc_gs_forward(const orc_csc* A, const double* b, double* u) {
  for (int col = 0; col < A->cols; ++col) gs_update(A, b, u, col);
}
/* include/amg/smoother.hpp:167-174 */
void orc_gs_backward(const orc_csc* A, const double* b, double* u) {
  for (int col = A->cols - 1; col >= 0; --col) gs_update(A, b, u, col);
}

/*
 * include/amg/common.hpp:17-27.  bhat = A*u (Eigen ColMajor sparse*dense:
 * zero-init, then for j ascending: s = 1*u[j]; bhat[row] += val*s), then
 * error += (b[i]-bhat[i])*(b[i]-bhat[i]) in ascending i.  Linear time here;
 * the reference's accidental O(N*nnz) re-evaluation yields the same numbers.
 */
double orc_rss(const orc_csc* A, const double* u, const double* b) {
  double* bhat = (double*)calloc((size_t)A->rows, sizeof(double));
  for (int j = 0; j < A->cols; ++j) {
    double s = 1.0 * u[j];
    for (int p = A->colptr[j]; p < A->colptr[j + 1]; ++p)
      bhat[A->rowidx[p]] += A->val[p] * s;
  }
  double error = 0.0;
  for (int i = 0; i < A->rows; ++i) error += (b[i] - bhat[i]) * (b[i] - bhat[i]);
  free(bhat);
  return error;
}

/*
 * include/amg/smoother.hpp:189-215  SparseGaussSeidel::smooth.  Returns the
 * number of iterations done; *final_error receives the last rss (100 if it
 * was never evaluated).  every == 0 disables the check (default ctor,
 * :183-187).
 */
int64_t orc_gs_smooth(const orc_csc* A, double* u, const double* b,
                      double tolerance, int64_t every, int64_t n_iters,
                      double* final_error) {
  int64_t iter = 0;
  double error = 100;
  while (iter < n_iters && error > tolerance) {
    orc_gs_forward(A, b, u);
    orc_gs_backward(A, b, u);
    iter += 1;
    if (every != 0 && iter % every == 0) error = orc_rss(A, u, b);
  }
  if (final_error) *final_error = error;
  return iter;
}

/*
 * include/amg/multigrid.hpp:272-274 (also :204, :236).  Eigen evaluates
 * r = f - A*u as r = f; r -= A*u, i.e. scaleAndAddTo(r, A, u, -1):
 * for j ascending: s = (-1)*u[j]; r[row] += val*s.  Per row this is
 * ((f_i - a1*u1) - a2*u2) - ... in ascending column order.
 */
void orc_residual(const orc_csc* A, const double* u, const double* f,
                  double* r) {
  memcpy(r, f, sizeof(double) * (size_t)A->rows);
  for (int j = 0; j < A->cols; ++j) {
    double s = -1.0 * u[j];
    for (int p = A->colptr[j]; p < A->colptr[j + 1]; ++p)
      r[A->rowidx[p]] += A->val[p] * s;
  }
}

/* Eigen ColMajor sparse*dense, plain assignment y = M*x (zero-init scatter).
 * include/amg/interpolator.hpp:52-56 (prolongation), :64-68 (restriction). */
void orc_spmv(const orc_csc* M, const double* x, double* y) {
  memset(y, 0, sizeof(double) * (size_t)M->rows);
  for (int j = 0; j < M->cols; ++j) {
    double s = 1.0 * x[j];
    for (int p = M->colptr[j]; p < M->colptr[j + 1]; ++p)
      y[M->rowidx[p]] += M->val[p] * s;
  }
}

/* ------------------------------------------------------------------ */
/* Smoothers with no reference counterpart (oracle-defined; SURVEY a-10) */
/* ------------------------------------------------------------------ */

/*
 * Damped Jacobi sweep: r_k = ((f_k - a_k1 u_1) - a_k2 u_2) - ... over row k of
 * A in ascending column order (diagonal term included), then
 * u_new[k] = u[k] + omega * (r_k / a_kk); rows with a zero/absent diagonal
 * keep u[k].  Row k of A is taken from AT (= transpose(A) in CSC, i.e. CSR(A)).
 */
void orc_jacobi_sweep(const orc_csc* AT, const double* u, const double* f,
                      double omega, double* u_new) {
  for (int k = 0; k < AT->cols; ++k) {
    double r = f[k], diag = 0.0;
    for (int p = AT->colptr[k]; p < AT->colptr[k + 1]; ++p) {
      int j = AT->rowidx[p];
      double a = AT->val[p];
      if (j == k) diag = a;
      r = r - a * u[j];
    }
    u_new[k] = (diag == 0.0) ? u[k] : u[k] + omega * (r / diag);
  }
}

/*
 * Greedy first-fit colouring in natural order of the graph
 * { (k,j) : k != j, A(k,j) != 0 or A(j,k) != 0 } (explicit zeros ignored).
 * Returns the number of colours; color[k] in [0, n_colors).
 */
int orc_greedy_coloring(const orc_csc* A, const orc_csc* AT, int* color) {
  int N = A->cols, ncol = 0;
  int cap = 64;
  int* used = (int*)xmalloc(sizeof(int) * (size_t)cap);
  for (int c = 0; c < cap; ++c) used[c] = -1;
  for (int k = 0; k < N; ++k) {
    const orc_csc* M[2] = {A, AT};
    for (int s = 0; s < 2; ++s)
      for (int p = M[s]->colptr[k]; p < M[s]->colptr[k + 1]; ++p) {
        int j = M[s]->rowidx[p];
        if (j < k && M[s]->val[p] != 0.0) {
          int c = color[j];
          if (c >= cap) {
            int ncap = cap * 2;
            while (c >= ncap) ncap *= 2;
            used = (int*)realloc(used, sizeof(int) * (size_t)ncap);
            for (int t = cap; t < ncap; ++t) used[t] = -1;
            cap = ncap;
          }
          used[c] = k;
        }
      }
    int c = 0;
    while (c < cap && used[c] == k) ++c;
    color[k] = c;
    if (c + 1 > ncol) ncol = c + 1;
    if (ncol >= cap) {
      int ncap = cap * 2;
      used = (int*)realloc(used, sizeof(int) * (size_t)ncap);
      for (int t = cap; t < ncap; ++t) used[t] = -1;
      cap = ncap;
    }
  }
  free(used);
  return ncol;
}

/*
 * One multicolour Gauss-Seidel pass over colour c: every row k with
 * color[k]==c gets u[k] = (f[k] - sum_{j!=k, asc} a_kj u_j) / a_kk, the
 * reference's update formula (smoother.hpp:129-138) applied to row k of A
 * (taken from AT), rows of one colour being mutually independent.
 * A symmetric multicolour sweep = colours 0..C-1 then C-1..0.
 */
void orc_color_gs_pass(const orc_csc* AT, const int* color, int c,
                       const double* f, double* u) {
  for (int k = 0; k < AT->cols; ++k) {
    if (color[k] != c) continue;
    double rsum = 0.0, diag = 0.0;
    for (int p = AT->colptr[k]; p < AT->colptr[k + 1]; ++p) {
      int j = AT->rowidx[p];
      double a = AT->val[p];
      if (j == k) diag = a;
      else rsum += a * u[j];
    }
    if (diag != 0.0) u[k] = (f[k] - rsum) / diag;
  }
}

/* ------------------------------------------------------------------ */
/* Coarsest-level direct solve                                        */
/* ------------------------------------------------------------------ */

/*
 * include/amg/multigrid.hpp:33,240-243,287-288 use Eigen::SimplicialLDLT
 * (AMD ordering).  Any backward-stable LDL^T reproduces the reference's
 * printed digits; this one is banded, natural order,
 * no pivoting (A is symmetric negative definite).  Storage: L is unit lower
 * banded, Lb[i*(bw+1)+t] = L(i, i-bw+t) for t<bw; D in d[i].
 * Only the lower triangle of A (rows >= col in each CSC column) is read,
 * like SimplicialLDLT<.., Lower>.
 */
typedef struct {
  int n, bw;
  double* L; /* n*(bw) : L[i*bw + (j-(i-bw))], j in [i-bw, i) */
  double* d; /* n */
} orc_ldlt;

orc_ldlt* orc_ldlt_factor(const orc_csc* A) {
  int n = A->cols, bw = 0;
  for (int c = 0; c < n; ++c)
    for (int p = A->colptr[c]; p < A->colptr[c + 1]; ++p)
      if (A->rowidx[p] - c > bw) bw = A->rowidx[p] - c;
  orc_ldlt* F = (orc_ldlt*)xmalloc(sizeof(orc_ldlt));
  F->n = n;
  F->bw = bw;
  F->L = (double*)calloc((size_t)n * (size_t)(bw ? bw : 1), sizeof(double));
  F->d = (double*)calloc((size_t)n, sizeof(double));
  /* dense band of the lower triangle: W[i*(bw+1) + (j-(i-bw))] = A(i,j) */
  int w = bw + 1;
  double* W = (double*)calloc((size_t)n * (size_t)w, sizeof(double));
  for (int c = 0; c < n; ++c)
    for (int p = A->colptr[c]; p < A->colptr[c + 1]; ++p) {
      int r = A->rowidx[p];
      if (r >= c) W[(size_t)r * w + (c - (r - bw))] = A->val[p];
    }
#define LL(i, j) F->L[(size_t)(i) * (bw ? bw : 1) + ((j) - ((i) - bw))]
#define WW(i, j) W[(size_t)(i) * w + ((j) - ((i) - bw))]
  for (int i = 0; i < n; ++i) {
    int j0 = i - bw < 0 ? 0 : i - bw;
    for (int j = j0; j < i; ++j) {
      /* L(i,j) = (A(i,j) - sum_{k<j} L(i,k) d_k L(j,k)) / d_j */
      double s = WW(i, j);
      int k0 = j - bw < 0 ? 0 : j - bw;
      if (k0 < j0) k0 = j0;
      for (int k = k0; k < j; ++k) s -= LL(i, k) * F->d[k] * LL(j, k);
      LL(i, j) = s / F->d[j];
    }
    double s = WW(i, i);
    for (int k = j0; k < i; ++k) s -= LL(i, k) * F->d[k] * LL(i, k);
    F->d[i] = s;
  }
  free(W);
  return F;
}
int orc_ldlt_n(const orc_ldlt* F) { return F->n; }
int orc_ldlt_bw(const orc_ldlt* F) { return F->bw; }
void orc_ldlt_copy_out(const orc_ldlt* F, double* L, double* d) {
  memcpy(L, F->L, sizeof(double) * (size_t)F->n * (size_t)(F->bw ? F->bw : 1));
  memcpy(d, F->d, sizeof(double) * (size_t)F->n);
}
/*
 * x = A^{-1} f.  Right-looking substitutions: forward  y = f; for i asc:
 * for t in 1..bw: y[i+t] -= L(i+t,i)*y[i];  scale z = y/d;  backward for i
 * desc: for t in 1..bw: z[i-t] -= L(i,i-t)*z[i].  (Per entry the updates
 * arrive in a fixed order, which the GPU kernel reproduces bit for bit.)
 */
void orc_ldlt_solve(const orc_ldlt* F, const double* f, double* x) {
  int n = F->n, bw = F->bw;
  memcpy(x, f, sizeof(double) * (size_t)n);
  for (int i = 0; i < n; ++i)
    for (int t = 1; t <= bw && i + t < n; ++t) x[i + t] -= LL(i + t, i) * x[i];
  for (int i = 0; i < n; ++i) x[i] = x[i] / F->d[i];
  for (int i = n - 1; i >= 0; --i)
    for (int t = 1; t <= bw && i - t >= 0; ++t) x[i - t] -= LL(i, i - t) * x[i];
}
#undef LL
#undef WW
void orc_ldlt_free(orc_ldlt* F) {
  if (!F) return;
  free(F->L);
  free(F->d);
  free(F);
}

/* ------------------------------------------------------------------ */
/* Multigrid driver                                                   */
/* ------------------------------------------------------------------ */

enum { ORC_SMOOTHER_GS = 0, ORC_SMOOTHER_JACOBI = 1, ORC_SMOOTHER_COLOR_GS = 2 };

typedef struct {
  int n_levels;
  int smoother;        /* ORC_SMOOTHER_* */
  int smoother_iters;  /* SmootherBase::n_iters; GS: fwd+bwd pairs, Jacobi: sweeps */
  double omega;        /* damped Jacobi only */
  int64_t* n_dofs;
  orc_csc** A;  /* per level */
  orc_csc** AT; /* per level, transpose (rows of A) */
  orc_csc** P;  /* per level < n_levels-1 */
  orc_csc** R;
  double** u;
  double** f;
  double** r;
  double* tmp; /* size n_dofs[0] */
  int** color; /* per level (COLOR_GS only) */
  int* n_colors;
  orc_ldlt* coarse;
  double tolerance;
  int64_t every, n_iters;
  /* results of the last solve() */
  int64_t iters_done;
  double last_error;
  int n_hist;
  double hist[4096];
} orc_mg;

/*
 * include/amg/multigrid.hpp:151-244 (constructor).  Returns NULL and sets
 * *err to 1 / 2 for the two validated conditions (:165-178), in that order.
 */
orc_mg* orc_mg_create(const orc_csc* A, const double* b, int64_t b_rows,
                      int n_levels, double tolerance, int64_t every,
                      int64_t n_iters, int smoother, int smoother_iters,
                      double omega, int* err) {
  if (err) *err = 0;
  if (every > n_iters) { if (err) *err = 1; return NULL; }
  if (A->rows != b_rows) { if (err) *err = 2; return NULL; }
  orc_mg* g = (orc_mg*)calloc(1, sizeof(orc_mg));
  g->n_levels = n_levels;
  g->smoother = smoother;
  g->smoother_iters = smoother_iters;
  g->omega = omega;
  g->tolerance = tolerance;
  g->every = every;
  g->n_iters = n_iters;
  g->n_dofs = (int64_t*)calloc((size_t)n_levels, sizeof(int64_t));
  g->A = (orc_csc**)calloc((size_t)n_levels, sizeof(void*));
  g->AT = (orc_csc**)calloc((size_t)n_levels, sizeof(void*));
  g->P = (orc_csc**)calloc((size_t)n_levels, sizeof(void*));
  g->R = (orc_csc**)calloc((size_t)n_levels, sizeof(void*));
  g->u = (double**)calloc((size_t)n_levels, sizeof(void*));
  g->f = (double**)calloc((size_t)n_levels, sizeof(void*));
  g->r = (double**)calloc((size_t)n_levels, sizeof(void*));
  g->color = (int**)calloc((size_t)n_levels, sizeof(void*));
  g->n_colors = (int*)calloc((size_t)n_levels, sizeof(int));
  int64_t N0 = A->rows;
  g->n_dofs[0] = N0;
  g->A[0] = orc_csc_from_arrays(A->rows, A->cols, A->colptr, A->rowidx, A->val);
  g->u[0] = (double*)calloc((size_t)N0, sizeof(double));
  g->f[0] = (double*)xmalloc(sizeof(double) * (size_t)N0);
  memcpy(g->f[0], b, sizeof(double) * (size_t)N0);
  g->r[0] = (double*)calloc((size_t)N0, sizeof(double));
  g->tmp = (double*)calloc((size_t)N0, sizeof(double));
  orc_residual(g->A[0], g->u[0], g->f[0], g->r[0]); /* :204 */
  for (int l = 1; l < n_levels; ++l) {
    int64_t nh = g->n_dofs[l - 1];
    int64_t nH = orc_n_H_dofs_from_n_h_dofs(nh);
    g->n_dofs[l] = nH;
    g->P[l - 1] = orc_make_P((int)nh, (int)nH);
    g->R[l - 1] = orc_transpose(g->P[l - 1]);
    g->A[l] = orc_galerkin(g->R[l - 1], g->A[l - 1], g->P[l - 1]);
    g->u[l] = (double*)calloc((size_t)nH, sizeof(double));
    g->f[l] = (double*)calloc((size_t)nH, sizeof(double));
    g->r[l] = (double*)calloc((size_t)nH, sizeof(double));
  }
  for (int l = 0; l < n_levels; ++l) {
    g->AT[l] = orc_transpose(g->A[l]);
    if (smoother == ORC_SMOOTHER_COLOR_GS) {
      g->color[l] = (int*)xmalloc(sizeof(int) * (size_t)g->n_dofs[l]);
      g->n_colors[l] = orc_greedy_coloring(g->A[l], g->AT[l], g->color[l]);
    }
  }
  g->coarse = orc_ldlt_factor(g->A[n_levels - 1]); /* :240-243 */
  return g;
}

void orc_mg_free(orc_mg* g) {
  if (!g) return;
  for (int l = 0; l < g->n_levels; ++l) {
    orc_csc_free(g->A[l]);
    orc_csc_free(g->AT[l]);
    orc_csc_free(g->P[l]);
    orc_csc_free(g->R[l]);
    free(g->u[l]);
    free(g->f[l]);
    free(g->r[l]);
    free(g->color[l]);
  }
  free(g->A); free(g->AT); free(g->P); free(g->R);
  free(g->u); free(g->f); free(g->r); free(g->color); free(g->n_colors);
  free(g->n_dofs); free(g->tmp);
  orc_ldlt_free(g->coarse);
  free(g);
}

/* Switch the smoother of an existing hierarchy (the operators do not depend on it), so one
 * oracle hierarchy serves both the parity check and the reference-smoother timing in
 * bench.py.  Test infrastructure, no reference counterpart. */
void orc_mg_set_smoother(orc_mg* g, int smoother, int smoother_iters, double omega) {
  g->smoother = smoother;
  g->smoother_iters = smoother_iters;
  g->omega = omega;
  if (smoother == ORC_SMOOTHER_COLOR_GS)
    for (int l = 0; l < g->n_levels; ++l)
      if (!g->color[l]) {
        g->color[l] = (int*)xmalloc(sizeof(int) * (size_t)g->n_dofs[l]);
        g->n_colors[l] = orc_greedy_coloring(g->A[l], g->AT[l], g->color[l]);
      }
}
/* u_l = 0 on every level (restart from the zero guess) */
void orc_mg_reset(orc_mg* g) {
  for (int l = 0; l < g->n_levels; ++l) memset(g->u[l], 0, sizeof(double) * (size_t)g->n_dofs[l]);
}

int orc_mg_n_levels(const orc_mg* g) { return g->n_levels; }
int64_t orc_mg_n_dofs(const orc_mg* g, int l) { return g->n_dofs[l]; }
const orc_csc* orc_mg_A(const orc_mg* g, int l) { return g->A[l]; }
const orc_csc* orc_mg_P(const orc_mg* g, int l) { return g->P[l]; }
const orc_csc* orc_mg_R(const orc_mg* g, int l) { return g->R[l]; }
double* orc_mg_u(orc_mg* g, int l) { return g->u[l]; }
double* orc_mg_f(orc_mg* g, int l) { return g->f[l]; }
double* orc_mg_r(orc_mg* g, int l) { return g->r[l]; }
const int* orc_mg_color(const orc_mg* g, int l) { return g->color[l]; }
int orc_mg_n_colors(const orc_mg* g, int l) { return g->n_colors[l]; }
const orc_ldlt* orc_mg_coarse(const orc_mg* g) { return g->coarse; }
int64_t orc_mg_iters_done(const orc_mg* g) { return g->iters_done; }
double orc_mg_last_error(const orc_mg* g) { return g->last_error; }
int orc_mg_hist(const orc_mg* g, double* out, int cap) {
  int n = g->n_hist < cap ? g->n_hist : cap;
  memcpy(out, g->hist, sizeof(double) * (size_t)n);
  return g->n_hist;
}

/* one smoother->smooth(A_l, u_l, f_l) call (multigrid.hpp:268-269, :300-301) */
void orc_mg_smooth(orc_mg* g, int l) {
  int64_t N = g->n_dofs[l];
  if (g->smoother == ORC_SMOOTHER_GS) {
    orc_gs_smooth(g->A[l], g->u[l], g->f[l], 1e-9, 0, g->smoother_iters, NULL);
  } else if (g->smoother == ORC_SMOOTHER_JACOBI) {
    for (int s = 0; s < g->smoother_iters; ++s) {
      orc_jacobi_sweep(g->AT[l], g->u[l], g->f[l], g->omega, g->tmp);
      memcpy(g->u[l], g->tmp, sizeof(double) * (size_t)N);
    }
  } else {
    for (int s = 0; s < g->smoother_iters; ++s) {
      for (int c = 0; c < g->n_colors[l]; ++c)
        orc_color_gs_pass(g->AT[l], g->color[l], c, g->f[l], g->u[l]);
      for (int c = g->n_colors[l] - 1; c >= 0; --c)
        orc_color_gs_pass(g->AT[l], g->color[l], c, g->f[l], g->u[l]);
    }
  }
}

/* include/amg/multigrid.hpp:263-305 */
void orc_mg_vcycle(orc_mg* g) {
  int L = g->n_levels;
  for (int l = 0; l < L; ++l) {
    orc_mg_smooth(g, l);                                   /* :268-269 */
    orc_residual(g->A[l], g->u[l], g->f[l], g->r[l]);      /* :272-274 */
    if (l + 1 != L) {
      memset(g->u[l + 1], 0, sizeof(double) * (size_t)g->n_dofs[l + 1]); /* :278 */
      orc_spmv(g->R[l], g->r[l], g->f[l + 1]);             /* :281-282 */
    }
  }
  orc_ldlt_solve(g->coarse, g->f[L - 1], g->u[L - 1]);     /* :287-288 */
  for (int l = L - 2; l >= 0; --l) {
    orc_spmv(g->P[l], g->u[l + 1], g->tmp);                /* :294-296 */
    for (int64_t i = 0; i < g->n_dofs[l]; ++i) g->u[l][i] = g->u[l][i] + g->tmp[i];
    orc_mg_smooth(g, l);                                   /* :300-301 */
  }
}

double orc_mg_rss(const orc_mg* g) {
  return orc_rss(g->A[0], g->u[0], g->f[0]);
}

/*
 * include/amg/multigrid.hpp:311-337.  Returns iterations done; the error at
 * each check is appended to the history (the reference only prints it).
 */
int64_t orc_mg_solve(orc_mg* g) {
  int64_t iter = 0;
  double error = 100;
  g->n_hist = 0;
  while (iter < g->n_iters && error > g->tolerance) {
    orc_mg_vcycle(g);
    iter += 1;
    if (g->every != 0 && (iter % g->every) == 0) {
      error = orc_mg_rss(g);
      if (g->n_hist < 4096) g->hist[g->n_hist++] = error;
    }
  }
  g->iters_done = iter;
  g->last_error = error;
  return iter;
}

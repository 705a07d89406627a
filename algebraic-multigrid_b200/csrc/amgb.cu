// libamgb.so -- C ABI (include/amgb.h) over the sm_100a kernels of kernels.cuh.
// Holds the device mirrors (amgb_matrix, amgb_hierarchy), enqueues the V-cycle
// of include/amg/multigrid.hpp:263-305 on a stream and replays it as a CUDA
// graph.  There is no CPU fallback: every compute entry point needs a device.
#include "../../include/amgb.h"

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "host_setup.hpp"
#include "kernels.cuh"

namespace {

using namespace amgb;
using amgb::dev::DiaView;
using amgb::dev::SellView;

thread_local std::string g_err;
std::atomic<int64_t> g_launches{0};

struct ApiError : std::runtime_error {
  int code;
  ApiError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define CUDA_CHECK(expr)                                                                   \
  do {                                                                                     \
    cudaError_t e_ = (expr);                                                               \
    if (e_ != cudaSuccess)                                                                 \
      throw ApiError(AMGB_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e_) + " (" + \
                                     __FILE__ + ":" + std::to_string(__LINE__) + ")");     \
  } while (0)

#define LAUNCH(kernel, grid, block, smem, stream, ...)            \
  do {                                                            \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);   \
    g_launches.fetch_add(1, std::memory_order_relaxed);           \
    CUDA_CHECK(cudaGetLastError());                               \
  } while (0)

template <class Fn>
int guarded(Fn&& fn) {
  try {
    fn();
    return AMGB_OK;
  } catch (const ApiError& e) {
    g_err = e.what();
    return e.code;
  } catch (const std::invalid_argument& e) {
    g_err = e.what();
    return AMGB_EINVAL;
  } catch (const std::exception& e) {
    g_err = e.what();
    return AMGB_ECUDA;
  }
}

void require_device() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    throw ApiError(AMGB_ECUDA, "no CUDA device: libamgb has no CPU fallback");
  }
}

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  void alloc(size_t count) {
    release();
    n = count;
    CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)));
  }
  void zero(cudaStream_t s) { CUDA_CHECK(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
  void upload(const T* h, size_t count, cudaStream_t s) {
    if (count != n || !p) alloc(count);
    if (count) CUDA_CHECK(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void upload(const std::vector<T>& h, cudaStream_t s) { upload(h.data(), h.size(), s); }
  void download(T* h, cudaStream_t s) const {
    if (n) CUDA_CHECK(cudaMemcpyAsync(h, p, n * sizeof(T), cudaMemcpyDeviceToHost, s));
  }
};

struct DevSell {
  int n_rows = 0, n_slices = 0;
  int64_t nnz = 0;
  DevBuf<uint32_t> slice_ptr;
  DevBuf<int> col;
  DevBuf<double> val;
  DevBuf<int> rows;
  void upload(const Sell& S, cudaStream_t s) {
    n_rows = S.n_rows;
    n_slices = S.n_slices;
    nnz = S.nnz;
    slice_ptr.upload(S.slice_ptr, s);
    col.upload(S.col, s);
    val.upload(S.val, s);
    if (!S.rows.empty()) rows.upload(S.rows, s);
    CUDA_CHECK(cudaStreamSynchronize(s));  // host vectors may die after return
  }
  SellView view() const { return SellView{n_rows, n_slices, slice_ptr.p, col.p, val.p, rows.p}; }
};

struct DevDia {
  int n_rows = 0, n_cols = 0, ld = 0, n_diag = 0;
  int64_t nnz = 0;
  int off[dev::kMaxDiagDev] = {0};
  DevBuf<double> val;
  DevBuf<int> rows;
  void upload(const Dia& D, cudaStream_t s) {
    n_rows = D.n_rows;
    n_cols = D.n_cols;
    ld = D.ld;
    n_diag = D.n_diag;
    nnz = D.nnz;
    for (int d = 0; d < n_diag; ++d) off[d] = D.off[d];
    val.upload(D.val, s);
    if (!D.rows.empty()) rows.upload(D.rows, s);
    CUDA_CHECK(cudaStreamSynchronize(s));
  }
  DiaView view() const {
    DiaView v;
    v.n_rows = n_rows;
    v.n_cols = n_cols;
    v.ld = ld;
    v.n_diag = n_diag;
    for (int d = 0; d < dev::kMaxDiagDev; ++d) v.off[d] = off[d];
    v.val = val.p;
    v.rows = rows.p;
    return v;
  }
};

// One device matrix: DIA when the operator is banded enough, else SELL-32.
struct DevMat {
  bool is_dia = false;
  DevDia dia;
  DevSell sell;
  void upload(const Csc& M, const std::vector<int>* rows, cudaStream_t s, bool allow_dia = true) {
    Dia D;
    if (allow_dia) D = build_dia(M, rows);
    is_dia = D.ok;
    if (is_dia) dia.upload(D, s);
    else sell.upload(build_sell(M, rows), s);
  }
  int n_rows() const { return is_dia ? dia.n_rows : sell.n_rows; }
  int64_t nnz() const { return is_dia ? dia.nnz : sell.nnz; }
  // bytes of matrix data one pass streams
  int64_t stored_bytes() const {
    return is_dia ? (int64_t)dia.val.n * 8 + (int64_t)dia.rows.n * 4
                  : (int64_t)sell.val.n * 12 + (int64_t)sell.slice_ptr.n * 4 + (int64_t)sell.rows.n * 4;
  }
};
template <int ND>
dev::DiaViewT<ND> dia_view_t(const DevDia& d) {
  dev::DiaViewT<ND> v;
  static_cast<DiaView&>(v) = d.view();
  return v;
}
template <class Fn>
void with_view(const DevMat& m, Fn&& fn) {
  if (!m.is_dia) fn(m.sell.view());
  else if (m.dia.n_diag <= 6) fn(dia_view_t<6>(m.dia));
  else if (m.dia.n_diag <= 10) fn(dia_view_t<10>(m.dia));
  else fn(dia_view_t<16>(m.dia));
}

inline int blocks_for(int64_t n, int per_block) { return (int)((n + per_block - 1) / per_block); }

// Device mirror of one operator plus the lazily built smoother schedules.
struct Operator {
  Csc M;                  // structural CSC exactly as handed in (explicit zeros kept)
  Csc MT;                 // transpose (rows of A); built once
  bool symmetric = true;  // M == MT bitwise
  int n = 0;
  DevMat colrows;                   // row c = CSC column c (smoother.hpp:101-117)
  std::unique_ptr<DevMat> arows;    // rows of A when !symmetric
  // generic Gauss-Seidel fronts
  bool have_fronts = false;
  DevBuf<int> f_order, f_ptr, b_order, b_ptr;
  int n_ffronts = 0, n_bfronts = 0;
  // multicolour
  bool have_colors = false;
  int n_colors = 0;
  std::vector<int> color;
  std::vector<std::unique_ptr<DevMat>> color_sell;

  void build(Csc&& A, cudaStream_t s) {
    if (A.rows != A.cols) throw std::invalid_argument("operator must be square");
    M = std::move(A);
    n = M.cols;
    MT = transpose(M);
    symmetric = bitwise_equal(M, MT);
    colrows.upload(M, nullptr, s);
    if (!symmetric) {
      arows.reset(new DevMat());
      arows->upload(MT, nullptr, s);
    }
  }
  const DevMat& rows_of_A() const { return symmetric ? colrows : *arows; }
  const Csc& host_rows_of_A() const { return symmetric ? M : MT; }  // CSC whose column k = row k of A

  void ensure_fronts(cudaStream_t s) {
    if (have_fronts) return;
    Schedule F = gs_schedule(M, true), B = gs_schedule(M, false);
    n_ffronts = F.n_fronts();
    n_bfronts = B.n_fronts();
    f_order.upload(F.order, s);
    f_ptr.upload(F.front_ptr, s);
    b_order.upload(B.order, s);
    b_ptr.upload(B.front_ptr, s);
    CUDA_CHECK(cudaStreamSynchronize(s));
    have_fronts = true;
  }
  void ensure_colors(cudaStream_t s) {
    if (have_colors) return;
    n_colors = greedy_coloring(M, MT, color);
    std::vector<std::vector<int>> members(n_colors);
    for (int k = 0; k < n; ++k) members[color[k]].push_back(k);
    color_sell.clear();
    for (int c = 0; c < n_colors; ++c) {
      color_sell.emplace_back(new DevMat());
      color_sell.back()->upload(host_rows_of_A(), &members[c], s);
    }
    have_colors = true;
  }

  // ---- launches (all asynchronous on s) ----
  void residual(const double* u, const double* f, double* r, cudaStream_t s) const {
    if (!n) return;
    with_view(rows_of_A(), [&](auto V) {
      auto kern = dev::k_residual<decltype(V)>;
      LAUNCH(kern, blocks_for(n, 256), 256, 0, s, V, u, f, r);
    });
  }
  void jacobi(const double* u, const double* f, double omega, double* out, cudaStream_t s) const {
    if (!n) return;
    with_view(rows_of_A(), [&](auto V) {
      auto kern = dev::k_jacobi<decltype(V)>;
      LAUNCH(kern, blocks_for(n, 256), 256, 0, s, V, u, f, omega, out);
    });
  }
  void color_pass(int c, const double* f, double* u, cudaStream_t s) const {
    const DevMat& C = *color_sell[c];
    if (!C.n_rows()) return;
    with_view(C, [&](auto V) {
      auto kern = dev::k_color_gs<decltype(V)>;
      LAUNCH(kern, blocks_for(V.n_rows, 256), 256, 0, s, V, f, u);
    });
  }
  void gs_fronts(const int* order, const int* ptr, int n_fronts, const double* f, double* u,
                 cudaStream_t s) const {
    if (!n) return;
    with_view(colrows, [&](auto V) {
      auto kern = dev::k_gs_fronts<decltype(V)>;
      LAUNCH(kern, 1, 1024, 0, s, V, order, ptr, n_fronts, f, u);
    });
  }
  void gs_forward(const double* f, double* u, cudaStream_t s) const {
    gs_fronts(f_order.p, f_ptr.p, n_ffronts, f, u, s);
  }
  void gs_backward(const double* f, double* u, cudaStream_t s) const {
    gs_fronts(b_order.p, b_ptr.p, n_bfronts, f, u, s);
  }
  void residual_restrict(const double* u, const double* f, double* f_coarse, double* u_coarse,
                         int n_coarse, cudaStream_t s) const {
    with_view(rows_of_A(), [&](auto V) {
      auto kern = dev::k_residual_restrict<decltype(V)>;
      LAUNCH(kern, blocks_for(n, 256), 256, 0, s, V, u, f, f_coarse, u_coarse, n_coarse);
    });
  }
  int rss_blocks() const { return std::max(1, blocks_for(n, 256)); }
  // partial must hold rss_blocks() doubles; out one double
  void rss(const double* u, const double* b, double* partial, double* out, cudaStream_t s) const {
    const int nb = rss_blocks();
    with_view(rows_of_A(), [&](auto V) {
      auto kern = dev::k_rss_partial<decltype(V)>;
      LAUNCH(kern, nb, 256, 0, s, V, u, b, partial);
    });
    LAUNCH(dev::k_sum_partials, 1, 256, 0, s, partial, nb, out);
  }
  int64_t nnz_device() const { return rows_of_A().nnz(); }
};

void options_default(amgb_options* o) {
  o->n_levels = 2;
  o->tolerance = 1e-9;
  o->compute_error_every_n_iters = 10;
  o->n_iters = 100;
  o->smoother = AMGB_SMOOTHER_GS;
  o->smoother_iters = 1;
  o->omega = 2.0 / 3.0;
  o->gs_mode = AMGB_GS_AUTO;
  o->use_graph = 1;
  o->skip_dead_coarse_smooth = 1;
}

}  // namespace

// ============================================================================
// amgb_matrix
// ============================================================================
struct amgb_matrix {
  int device = 0;
  cudaStream_t stream = nullptr;
  Operator op;
  DevBuf<double> u, b, r, partial, scalar;
  ~amgb_matrix() {
    if (stream) cudaStreamDestroy(stream);
  }
};

// ============================================================================
// amgb_hierarchy
// ============================================================================
struct amgb_hierarchy {
  int device = 0;
  amgb_options opt{};
  int L = 0;
  std::vector<int64_t> n;
  std::vector<std::unique_ptr<Operator>> ops;
  std::vector<DevBuf<double>> u, f, tmp;
  DevBuf<double> partial, scalar;
  BandedLdlt factor;
  DevBuf<double> dL, dd, dwork;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  int64_t launches_per_vcycle = 0;
  bool ldlt_attr_set = false;
  int64_t iters_done = 0;
  std::vector<double> history;

  ~amgb_hierarchy() {
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
    if (own_stream) cudaStreamDestroy(own_stream);
  }

  void prepare_smoother(int l) {
    if (opt.smoother == AMGB_SMOOTHER_GS) ops[l]->ensure_fronts(stream);
    if (opt.smoother == AMGB_SMOOTHER_COLOR_GS) ops[l]->ensure_colors(stream);
  }

  // smoother->smooth(A_l, u_l, f_l)   (multigrid.hpp:268-269, :300-301)
  void smooth(int l, cudaStream_t s) {
    Operator& A = *ops[l];
    const int64_t iters = opt.smoother_iters;
    if (opt.smoother == AMGB_SMOOTHER_GS) {
      for (int64_t it = 0; it < iters; ++it) {  // smoother.hpp:195-198
        A.gs_forward(f[l].p, u[l].p, s);
        A.gs_backward(f[l].p, u[l].p, s);
      }
    } else if (opt.smoother == AMGB_SMOOTHER_JACOBI) {
      double* src = u[l].p;
      double* dst = tmp[l].p;
      for (int64_t it = 0; it < iters; ++it) {
        A.jacobi(src, f[l].p, opt.omega, dst, s);
        std::swap(src, dst);
      }
      if (src != u[l].p)
        CUDA_CHECK(cudaMemcpyAsync(u[l].p, src, sizeof(double) * n[l], cudaMemcpyDeviceToDevice, s));
    } else {
      for (int64_t it = 0; it < iters; ++it) {
        for (int c = 0; c < A.n_colors; ++c) A.color_pass(c, f[l].p, u[l].p, s);
        for (int c = A.n_colors - 1; c >= 0; --c) A.color_pass(c, f[l].p, u[l].p, s);
      }
    }
  }
  // f_{l+1} = R_l (f_l - A_l u_l), u_{l+1} = 0   (multigrid.hpp:272-282)
  void residual_restrict(int l, cudaStream_t s) {
    ops[l]->residual_restrict(u[l].p, f[l].p, f[l + 1].p, u[l + 1].p, (int)n[l + 1], s);
  }
  // u_l = u_l + P_l u_{l+1}   (multigrid.hpp:294-296)
  void prolong_add(int l, cudaStream_t s) {
    LAUNCH(dev::k_prolong_add, blocks_for(n[l], 256), 256, 0, s, u[l + 1].p, (int)n[l + 1], u[l].p,
           (int)n[l]);
  }
  void coarse_solve(cudaStream_t s) {  // multigrid.hpp:287-288
    const int nc = factor.n, bw = factor.bw;
    int threads = std::min(1024, std::max(32, ((bw + 31) / 32) * 32));
    const size_t xbytes = sizeof(double) * (size_t)nc;
    const size_t lbytes = sizeof(double) * (size_t)nc * std::max(bw, 1);
    const size_t cap = 200 * 1024;
    const int x_smem = xbytes <= cap;
    const int l_smem = x_smem && xbytes + lbytes <= cap;
    const size_t smem = (x_smem ? xbytes : 0) + (l_smem ? lbytes : 0);
    if (smem > 48 * 1024 && !ldlt_attr_set) {
      CUDA_CHECK(cudaFuncSetAttribute(dev::k_banded_ldlt_solve, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)cap));
      ldlt_attr_set = true;
    }
    LAUNCH(dev::k_banded_ldlt_solve, 1, threads, smem, s, dL.p, dd.p, nc, bw, f[L - 1].p, u[L - 1].p,
           dwork.p, x_smem, l_smem);
  }
  void enqueue_vcycle(cudaStream_t s) {
    for (int l = 0; l < L; ++l) {
      const bool coarsest = (l + 1 == L);
      if (coarsest && opt.skip_dead_coarse_smooth) break;
      smooth(l, s);
      if (!coarsest) residual_restrict(l, s);
      // on the coarsest level the reference also forms the residual (:272-274);
      // it is stored in a private member without a getter and never read.
    }
    coarse_solve(s);
    for (int l = L - 2; l >= 0; --l) {
      prolong_add(l, s);
      smooth(l, s);
    }
  }
  void build_graph() {
    if (exec) return;
    const int64_t before = g_launches.load();
    CUDA_CHECK(cudaStreamBeginCapture(own_stream, cudaStreamCaptureModeThreadLocal));
    try {
      enqueue_vcycle(own_stream);
    } catch (...) {
      cudaGraph_t g = nullptr;
      cudaStreamEndCapture(own_stream, &g);
      if (g) cudaGraphDestroy(g);
      throw;
    }
    CUDA_CHECK(cudaStreamEndCapture(own_stream, &graph));
    CUDA_CHECK(cudaGraphInstantiate(&exec, graph, 0));
    launches_per_vcycle = g_launches.load() - before;
    g_launches.store(before);  // capture itself launched nothing
  }
  void vcycle() {
    if (opt.use_graph) {
      build_graph();
      CUDA_CHECK(cudaGraphLaunch(exec, stream));
      g_launches.fetch_add(launches_per_vcycle, std::memory_order_relaxed);
    } else {
      const int64_t before = g_launches.load();
      enqueue_vcycle(stream);
      launches_per_vcycle = g_launches.load() - before;
    }
  }
  double rss() {
    ops[0]->rss(u[0].p, f[0].p, partial.p, scalar.p, stream);
    double out = 0.0;
    CUDA_CHECK(cudaMemcpyAsync(&out, scalar.p, sizeof(double), cudaMemcpyDeviceToHost, stream));
    CUDA_CHECK(cudaStreamSynchronize(stream));
    return out;
  }
  double sumsq(const double* x, int64_t count) {
    const int nb = std::max(1, blocks_for(count, 256));
    LAUNCH(dev::k_sumsq_partial, nb, 256, 0, stream, x, (int)count, partial.p);
    LAUNCH(dev::k_sum_partials, 1, 256, 0, stream, partial.p, nb, scalar.p);
    double out = 0.0;
    CUDA_CHECK(cudaMemcpyAsync(&out, scalar.p, sizeof(double), cudaMemcpyDeviceToHost, stream));
    CUDA_CHECK(cudaStreamSynchronize(stream));
    return out;
  }
  void check_level(int l, bool need_next = false) const {
    if (l < 0 || l >= L || (need_next && l + 1 >= L)) throw std::invalid_argument("level out of range");
  }
};

// ============================================================================
// C ABI
// ============================================================================
extern "C" {

const char* amgb_last_error(void) { return g_err.c_str(); }
int amgb_version(void) { return 100; }
int amgb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}
int amgb_set_device(int device) {
  return guarded([&] { CUDA_CHECK(cudaSetDevice(device)); });
}

// ---- generators / setup helpers (host only) ----
double amgb_grid_spacing_h(int64_t n) { return grid_spacing_h(n); }
int64_t amgb_points_n_from_grid_spacing_h(double h) { return (int64_t)((2 / h) - 1); }  // grid.hpp:39-41
int64_t amgb_grid_laplacian_nnz(int64_t n) { return 5 * n * n - 4 * n; }
int amgb_grid_laplacian(int64_t n, double eps_y, int* colptr, int* rowidx, double* val) {
  return guarded([&] {
    if (n < 1 || n * n > 2147483647LL / 5) throw std::invalid_argument("grid size out of range");
    Csc A = grid_laplacian(n, eps_y);
    std::memcpy(colptr, A.colptr.data(), A.colptr.size() * sizeof(int));
    std::memcpy(rowidx, A.rowidx.data(), A.rowidx.size() * sizeof(int));
    std::memcpy(val, A.val.data(), A.val.size() * sizeof(double));
  });
}
int amgb_grid_rhs(int64_t n, double* b) {
  return guarded([&] {
    if (n < 1) throw std::invalid_argument("grid size out of range");
    grid_rhs(n, b);
  });
}
int64_t amgb_n_H_dofs_from_n_h_dofs(int64_t n_h) { return coarse_dofs(n_h); }
int64_t amgb_interp_nnz(int64_t n_h, int64_t n_H) {
  int64_t c = 0;
  for (int64_t j = 0; j < n_H; ++j) c += (2 * j < n_h) + (2 * j + 1 < n_h) + (2 * j + 2 < n_h);
  return c;
}
int amgb_interp_make_operators(int64_t n_h, int64_t n_H, int* P_colptr, int* P_rowidx, double* P_val,
                               int* R_colptr, int* R_rowidx, double* R_val) {
  return guarded([&] {
    if (n_h < 0 || n_H < 0) throw std::invalid_argument("negative size");
    Csc P = make_prolongation(n_h, n_H);
    Csc R = transpose(P);
    std::memcpy(P_colptr, P.colptr.data(), P.colptr.size() * sizeof(int));
    std::memcpy(P_rowidx, P.rowidx.data(), P.rowidx.size() * sizeof(int));
    std::memcpy(P_val, P.val.data(), P.val.size() * sizeof(double));
    std::memcpy(R_colptr, R.colptr.data(), R.colptr.size() * sizeof(int));
    std::memcpy(R_rowidx, R.rowidx.data(), R.rowidx.size() * sizeof(int));
    std::memcpy(R_val, R.val.data(), R.val.size() * sizeof(double));
  });
}

// ---- amgb_matrix ----
int amgb_matrix_create(int n_rows, int n_cols, const int* colptr, const int* rowidx, const double* val,
                       amgb_matrix** out) {
  return guarded([&] {
    if (!out || !colptr || n_rows < 0 || n_cols < 0) throw std::invalid_argument("bad matrix arguments");
    require_device();
    std::unique_ptr<amgb_matrix> m(new amgb_matrix());
    CUDA_CHECK(cudaGetDevice(&m->device));
    CUDA_CHECK(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
    m->op.build(csc_from_arrays(n_rows, n_cols, colptr, rowidx, val), m->stream);
    m->u.alloc(n_cols);
    m->b.alloc(n_cols);
    m->r.alloc(n_cols);
    m->partial.alloc(m->op.rss_blocks());
    m->scalar.alloc(1);
    *out = m.release();
  });
}
int amgb_matrix_destroy(amgb_matrix* A) {
  return guarded([&] {
    if (!A) return;
    cudaSetDevice(A->device);
    delete A;
  });
}
int64_t amgb_matrix_nnz_device(const amgb_matrix* A) { return A ? A->op.nnz_device() : 0; }
int amgb_matrix_is_symmetric(const amgb_matrix* A) { return A && A->op.symmetric; }

static double matrix_rss(amgb_matrix* A) {
  A->op.rss(A->u.p, A->b.p, A->partial.p, A->scalar.p, A->stream);
  double out = 0.0;
  CUDA_CHECK(cudaMemcpyAsync(&out, A->scalar.p, sizeof(double), cudaMemcpyDeviceToHost, A->stream));
  CUDA_CHECK(cudaStreamSynchronize(A->stream));
  return out;
}

int amgb_smooth_gs(amgb_matrix* A, double* u, const double* b, double tolerance, int64_t every,
                   int64_t n_iters, int mode, int64_t* iters_done, double* final_error) {
  return guarded([&] {
    if (!A || !u || !b) throw std::invalid_argument("null argument");
    (void)mode;
    CUDA_CHECK(cudaSetDevice(A->device));
    cudaStream_t s = A->stream;
    A->op.ensure_fronts(s);
    A->u.upload(u, A->op.n, s);
    A->b.upload(b, A->op.n, s);
    int64_t iter = 0;
    double error = 100;
    while (iter < n_iters && error > tolerance) {  // smoother.hpp:195-203
      A->op.gs_forward(A->b.p, A->u.p, s);
      A->op.gs_backward(A->b.p, A->u.p, s);
      iter += 1;
      if (every != 0 && iter % every == 0) error = matrix_rss(A);
    }
    A->u.download(u, s);
    CUDA_CHECK(cudaStreamSynchronize(s));
    if (iters_done) *iters_done = iter;
    if (final_error) *final_error = error;
  });
}
int amgb_smooth_jacobi(amgb_matrix* A, double* u, const double* b, double omega, int64_t n_sweeps) {
  return guarded([&] {
    if (!A || !u || !b) throw std::invalid_argument("null argument");
    CUDA_CHECK(cudaSetDevice(A->device));
    cudaStream_t s = A->stream;
    A->u.upload(u, A->op.n, s);
    A->b.upload(b, A->op.n, s);
    double* src = A->u.p;
    double* dst = A->r.p;
    for (int64_t it = 0; it < n_sweeps; ++it) {
      A->op.jacobi(src, A->b.p, omega, dst, s);
      std::swap(src, dst);
    }
    if (A->op.n) CUDA_CHECK(cudaMemcpyAsync(u, src, sizeof(double) * A->op.n, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
  });
}
int amgb_smooth_color_gs(amgb_matrix* A, double* u, const double* b, int64_t n_iters) {
  return guarded([&] {
    if (!A || !u || !b) throw std::invalid_argument("null argument");
    CUDA_CHECK(cudaSetDevice(A->device));
    cudaStream_t s = A->stream;
    A->op.ensure_colors(s);
    A->u.upload(u, A->op.n, s);
    A->b.upload(b, A->op.n, s);
    for (int64_t it = 0; it < n_iters; ++it) {
      for (int c = 0; c < A->op.n_colors; ++c) A->op.color_pass(c, A->b.p, A->u.p, s);
      for (int c = A->op.n_colors - 1; c >= 0; --c) A->op.color_pass(c, A->b.p, A->u.p, s);
    }
    A->u.download(u, s);
    CUDA_CHECK(cudaStreamSynchronize(s));
  });
}
int amgb_matrix_coloring(amgb_matrix* A, int* n_colors, int* color) {
  return guarded([&] {
    if (!A) throw std::invalid_argument("null argument");
    CUDA_CHECK(cudaSetDevice(A->device));
    A->op.ensure_colors(A->stream);
    if (n_colors) *n_colors = A->op.n_colors;
    if (color) std::memcpy(color, A->op.color.data(), sizeof(int) * A->op.color.size());
  });
}
int amgb_residual(amgb_matrix* A, const double* u, const double* f, double* r) {
  return guarded([&] {
    if (!A || !u || !f || !r) throw std::invalid_argument("null argument");
    CUDA_CHECK(cudaSetDevice(A->device));
    cudaStream_t s = A->stream;
    A->u.upload(u, A->op.n, s);
    A->b.upload(f, A->op.n, s);
    A->op.residual(A->u.p, A->b.p, A->r.p, s);
    A->r.download(r, s);
    CUDA_CHECK(cudaStreamSynchronize(s));
  });
}
int amgb_rss(amgb_matrix* A, const double* u, const double* b, double* out) {
  return guarded([&] {
    if (!A || !u || !b || !out) throw std::invalid_argument("null argument");
    CUDA_CHECK(cudaSetDevice(A->device));
    A->u.upload(u, A->op.n, A->stream);
    A->b.upload(b, A->op.n, A->stream);
    *out = matrix_rss(A);
  });
}

// ---- hierarchy ----
void amgb_options_default(amgb_options* opt) {
  if (opt) options_default(opt);
}

int amgb_hierarchy_create(int n_rows, int n_cols, const int* colptr, const int* rowidx, const double* val,
                          const double* b, int64_t b_rows, const amgb_options* opt_in,
                          amgb_hierarchy** out) {
  return guarded([&] {
    if (!out || !opt_in || !colptr) throw std::invalid_argument("null argument");
    const amgb_options& o = *opt_in;
    // multigrid.hpp:165-178 -- same order, same messages
    if (o.compute_error_every_n_iters > o.n_iters)
      throw std::invalid_argument("`compute_error_every_n_iters` must be leq to `n_iters`, got " +
                                  std::to_string(o.compute_error_every_n_iters) + " and " +
                                  std::to_string(o.n_iters));
    if ((int64_t)n_rows != b_rows)
      throw std::invalid_argument("`A` and `b` must have the same number of degrees of freedom, got " +
                                  std::to_string(n_rows) + " and " + std::to_string(b_rows));
    if (o.n_levels < 1) throw std::invalid_argument("n_levels must be >= 1");
    if (n_rows != n_cols) throw std::invalid_argument("A must be square");
    if (o.smoother < 0 || o.smoother > 2) throw std::invalid_argument("unknown smoother kind");
    if (!b || !rowidx || !val) throw std::invalid_argument("null argument");
    require_device();

    std::unique_ptr<amgb_hierarchy> h(new amgb_hierarchy());
    h->opt = o;
    h->L = o.n_levels;
    CUDA_CHECK(cudaGetDevice(&h->device));
    CUDA_CHECK(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    h->stream = h->own_stream;
    cudaStream_t s = h->stream;

    h->n.resize(h->L);
    h->ops.resize(h->L);
    h->u.resize(h->L);
    h->f.resize(h->L);
    h->tmp.resize(h->L);

    // level 0 (multigrid.hpp:190-204)
    Csc A = csc_from_arrays(n_rows, n_cols, colptr, rowidx, val);
    for (int l = 0; l < h->L; ++l) {
      if (l > 0) {
        // multigrid.hpp:211-223
        const Csc& Ah = h->ops[l - 1]->M;
        const int64_t nh = h->n[l - 1];
        const int64_t nH = coarse_dofs(nh);
        if (nH < 1) throw std::invalid_argument("too many levels: level " + std::to_string(l) + " is empty");
        Csc P = make_prolongation(nh, nH);
        Csc R = transpose(P);
        A = galerkin(R, Ah, P);
      }
      h->n[l] = A.cols;
      h->ops[l].reset(new Operator());
      h->ops[l]->build(std::move(A), s);
      h->u[l].alloc(h->n[l]);
      h->u[l].zero(s);
      h->f[l].alloc(h->n[l]);
      if (l == 0) h->f[l].upload(b, h->n[0], s);
      else h->f[l].zero(s);
      if (o.smoother == AMGB_SMOOTHER_JACOBI) h->tmp[l].alloc(h->n[l]);
      if (!(l + 1 == h->L && o.skip_dead_coarse_smooth)) h->prepare_smoother(l);
    }
    h->partial.alloc(std::max(1, blocks_for(h->n[0], 256)));
    h->scalar.alloc(1);
    // coarsest factorisation (multigrid.hpp:240-243); the direct solve is a banded
    // substitution in one block, so refuse hierarchies whose coarsest level is not small
    {
      const Csc& Ac = h->ops[h->L - 1]->M;
      int64_t bw = 0;
      for (int c = 0; c < Ac.cols; ++c)
        for (int p = Ac.colptr[c]; p < Ac.colptr[c + 1]; ++p) bw = std::max<int64_t>(bw, Ac.rowidx[p] - c);
      if ((double)Ac.cols * (double)bw * (double)bw > 2e10 || (int64_t)Ac.cols * std::max<int64_t>(bw, 1) > (1ll << 27))
        throw std::invalid_argument("coarsest level too large for the direct solve (" +
                                    std::to_string(Ac.cols) + " DOF, half-bandwidth " +
                                    std::to_string(bw) + "): use more levels");
    }
    h->factor = factor_banded_ldlt(h->ops[h->L - 1]->M);
    h->dL.upload(h->factor.L, s);
    h->dd.upload(h->factor.d, s);
    h->dwork.alloc(h->factor.n);
    CUDA_CHECK(cudaStreamSynchronize(s));
    *out = h.release();
  });
}
int amgb_hierarchy_destroy(amgb_hierarchy* h) {
  return guarded([&] {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    delete h;
  });
}
int amgb_hierarchy_set_stream(amgb_hierarchy* h, void* cuda_stream) {
  return guarded([&] {
    if (!h) throw std::invalid_argument("null argument");
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
  });
}
int amgb_hierarchy_n_levels(const amgb_hierarchy* h) { return h ? h->L : 0; }
int64_t amgb_hierarchy_n_dofs(const amgb_hierarchy* h, int level) {
  return (h && level >= 0 && level < h->L) ? h->n[level] : -1;
}
int64_t amgb_hierarchy_nnz(const amgb_hierarchy* h, int level) {
  return (h && level >= 0 && level < h->L) ? h->ops[level]->M.nnz() : -1;
}
int64_t amgb_hierarchy_nnz_device(const amgb_hierarchy* h, int level) {
  return (h && level >= 0 && level < h->L) ? h->ops[level]->nnz_device() : -1;
}
double amgb_hierarchy_tolerance(const amgb_hierarchy* h) { return h ? h->opt.tolerance : 0.0; }
int amgb_hierarchy_get_matrix(const amgb_hierarchy* h, int level, int* colptr, int* rowidx, double* val) {
  return guarded([&] {
    if (!h) throw std::invalid_argument("null argument");
    h->check_level(level);
    const Csc& M = h->ops[level]->M;
    if (colptr) std::memcpy(colptr, M.colptr.data(), M.colptr.size() * sizeof(int));
    if (rowidx) std::memcpy(rowidx, M.rowidx.data(), M.rowidx.size() * sizeof(int));
    if (val) std::memcpy(val, M.val.data(), M.val.size() * sizeof(double));
  });
}
static int copy_level_vec(amgb_hierarchy* h, int level, std::vector<DevBuf<double>>& v, double* host,
                          const double* src) {
  return guarded([&] {
    if (!h || (!host && !src)) throw std::invalid_argument("null argument");
    h->check_level(level);
    CUDA_CHECK(cudaSetDevice(h->device));
    if (host) v[level].download(host, h->stream);
    else CUDA_CHECK(cudaMemcpyAsync(v[level].p, src, sizeof(double) * h->n[level], cudaMemcpyHostToDevice,
                                    h->stream));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
  });
}
int amgb_hierarchy_get_soln(amgb_hierarchy* h, int level, double* u) {
  return copy_level_vec(h, level, h->u, u, nullptr);
}
int amgb_hierarchy_get_rhs(amgb_hierarchy* h, int level, double* f) {
  return copy_level_vec(h, level, h->f, f, nullptr);
}
int amgb_hierarchy_set_soln(amgb_hierarchy* h, int level, const double* u) {
  return copy_level_vec(h, level, h->u, nullptr, u);
}
int amgb_hierarchy_set_rhs(amgb_hierarchy* h, int level, const double* f) {
  return copy_level_vec(h, level, h->f, nullptr, f);
}
int amgb_hierarchy_get_coloring(const amgb_hierarchy* h, int level, int* n_colors, int* color) {
  return guarded([&] {
    if (!h) throw std::invalid_argument("null argument");
    h->check_level(level);
    const Operator& A = *h->ops[level];
    if (!A.have_colors) throw ApiError(AMGB_ESTATE, "level has no colouring (smoother is not COLOR_GS)");
    if (n_colors) *n_colors = A.n_colors;
    if (color) std::memcpy(color, A.color.data(), sizeof(int) * A.color.size());
  });
}

int amgb_vcycle(amgb_hierarchy* h) {
  return guarded([&] {
    if (!h) throw std::invalid_argument("null argument");
    CUDA_CHECK(cudaSetDevice(h->device));
    h->vcycle();
  });
}
int amgb_vcycles(amgb_hierarchy* h, int64_t count) {
  return guarded([&] {
    if (!h) throw std::invalid_argument("null argument");
    CUDA_CHECK(cudaSetDevice(h->device));
    for (int64_t i = 0; i < count; ++i) h->vcycle();
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
  });
}
int amgb_hierarchy_rss(amgb_hierarchy* h, double* out) {
  return guarded([&] {
    if (!h || !out) throw std::invalid_argument("null argument");
    CUDA_CHECK(cudaSetDevice(h->device));
    *out = h->rss();
  });
}
int amgb_solve(amgb_hierarchy* h, int64_t* iters_done, double* last_error) {
  return guarded([&] {
    if (!h) throw std::invalid_argument("null argument");
    CUDA_CHECK(cudaSetDevice(h->device));
    // multigrid.hpp:311-337
    int64_t iter = 0;
    double error = 100;
    h->history.clear();
    const int64_t every = h->opt.compute_error_every_n_iters;
    while (iter < h->opt.n_iters && error > h->opt.tolerance) {
      h->vcycle();
      iter += 1;
      if (every != 0 && (iter % every) == 0) {
        error = h->rss();
        h->history.push_back(error);
      }
    }
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    h->iters_done = iter;
    if (iters_done) *iters_done = iter;
    if (last_error) *last_error = error;
  });
}
int amgb_solve_relative(amgb_hierarchy* h, double rel_tol, int64_t* iters_done, double* last_rel) {
  return guarded([&] {
    if (!h) throw std::invalid_argument("null argument");
    CUDA_CHECK(cudaSetDevice(h->device));
    const double bnorm2 = h->sumsq(h->f[0].p, h->n[0]);
    int64_t iter = 0;
    double rel = INFINITY;
    h->history.clear();
    const int64_t every = std::max<int64_t>(1, h->opt.compute_error_every_n_iters);
    while (iter < h->opt.n_iters && rel > rel_tol) {
      h->vcycle();
      iter += 1;
      if ((iter % every) == 0) {
        rel = std::sqrt(h->rss() / bnorm2);
        h->history.push_back(rel);
      }
    }
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    h->iters_done = iter;
    if (iters_done) *iters_done = iter;
    if (last_rel) *last_rel = rel;
  });
}
int64_t amgb_hierarchy_iters_done(const amgb_hierarchy* h) { return h ? h->iters_done : 0; }
int64_t amgb_hierarchy_error_history(const amgb_hierarchy* h, double* out, int64_t cap) {
  if (!h) return 0;
  const int64_t nh = (int64_t)h->history.size();
  if (out) std::memcpy(out, h->history.data(), sizeof(double) * std::min(cap, nh));
  return nh;
}
int amgb_synchronize(amgb_hierarchy* h) {
  return guarded([&] {
    if (!h) throw std::invalid_argument("null argument");
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
  });
}

int amgb_restrict(amgb_hierarchy* h, int level, const double* r_fine, double* f_coarse) {
  return guarded([&] {
    if (!h || !r_fine || !f_coarse) throw std::invalid_argument("null argument");
    h->check_level(level, true);
    CUDA_CHECK(cudaSetDevice(h->device));
    DevBuf<double> r, fc;
    r.upload(r_fine, h->n[level], h->stream);
    fc.alloc(h->n[level + 1]);
    LAUNCH(dev::k_restrict, blocks_for(h->n[level + 1], 256), 256, 0, h->stream, r.p, (int)h->n[level],
           fc.p, (int)h->n[level + 1]);
    fc.download(f_coarse, h->stream);
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
  });
}
int amgb_prolong_add(amgb_hierarchy* h, int level, const double* e_coarse, double* u_fine) {
  return guarded([&] {
    if (!h || !e_coarse || !u_fine) throw std::invalid_argument("null argument");
    h->check_level(level, true);
    CUDA_CHECK(cudaSetDevice(h->device));
    DevBuf<double> e, uf;
    e.upload(e_coarse, h->n[level + 1], h->stream);
    uf.upload(u_fine, h->n[level], h->stream);
    LAUNCH(dev::k_prolong_add, blocks_for(h->n[level], 256), 256, 0, h->stream, e.p, (int)h->n[level + 1],
           uf.p, (int)h->n[level]);
    uf.download(u_fine, h->stream);
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
  });
}
int amgb_smooth_level(amgb_hierarchy* h, int level) {
  return guarded([&] {
    if (!h) throw std::invalid_argument("null argument");
    h->check_level(level);
    CUDA_CHECK(cudaSetDevice(h->device));
    h->prepare_smoother(level);
    if (h->opt.smoother == AMGB_SMOOTHER_JACOBI && !h->tmp[level].p) h->tmp[level].alloc(h->n[level]);
    h->smooth(level, h->stream);
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
  });
}
int amgb_residual_level(amgb_hierarchy* h, int level, double* r) {
  return guarded([&] {
    if (!h || !r) throw std::invalid_argument("null argument");
    h->check_level(level);
    CUDA_CHECK(cudaSetDevice(h->device));
    DevBuf<double> rd;
    rd.alloc(h->n[level]);
    h->ops[level]->residual(h->u[level].p, h->f[level].p, rd.p, h->stream);
    rd.download(r, h->stream);
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
  });
}
int amgb_residual_restrict_level(amgb_hierarchy* h, int level) {
  return guarded([&] {
    if (!h) throw std::invalid_argument("null argument");
    h->check_level(level, true);
    CUDA_CHECK(cudaSetDevice(h->device));
    h->residual_restrict(level, h->stream);
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
  });
}
int amgb_coarse_solve(amgb_hierarchy* h) {
  return guarded([&] {
    if (!h) throw std::invalid_argument("null argument");
    CUDA_CHECK(cudaSetDevice(h->device));
    h->coarse_solve(h->stream);
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
  });
}

int64_t amgb_kernel_launches(void) { return g_launches.load(); }
int64_t amgb_hierarchy_launches_per_vcycle(const amgb_hierarchy* h) {
  return h ? h->launches_per_vcycle : 0;
}
int64_t amgb_hierarchy_pass_bytes(const amgb_hierarchy* h, int level) {
  if (!h || level < 0 || level >= h->L) return -1;
  return 12 * h->ops[level]->nnz_device() + 28 * h->n[level] + 4;
}
int64_t amgb_hierarchy_vcycle_bytes(const amgb_hierarchy* h) {
  // per non-coarsest level 4 smoother passes + 1 residual
  // (= 5 B_l) + restriction/prolongation vectors (24 N_l + 16 N_{l+1}); scaled by
  // the smoother passes actually configured.
  if (!h) return -1;
  int64_t total = 0;
  const int64_t passes_per_smooth =
      h->opt.smoother == AMGB_SMOOTHER_JACOBI ? h->opt.smoother_iters : 2 * h->opt.smoother_iters;
  for (int l = 0; l + 1 < h->L; ++l) {
    const int64_t B = amgb_hierarchy_pass_bytes(h, l);
    total += (2 * passes_per_smooth + 1) * B + 24 * h->n[l] + 16 * h->n[l + 1];
  }
  return total;
}

int amgb_hierarchy_format(const amgb_hierarchy* h, int level) {
  if (!h || level < 0 || level >= h->L) return -1;
  return h->ops[level]->rows_of_A().is_dia ? AMGB_FORMAT_DIA : AMGB_FORMAT_SELL;
}
int64_t amgb_hierarchy_matrix_bytes(const amgb_hierarchy* h, int level) {
  if (!h || level < 0 || level >= h->L) return -1;
  return h->ops[level]->rows_of_A().stored_bytes();
}

int amgb_time_kernel(amgb_hierarchy* h, int level, int kind, int warmup, int reps, double* ms_out) {
  return guarded([&] {
    if (!h || !ms_out || reps < 1) throw std::invalid_argument("bad argument");
    h->check_level(level, kind >= 2);
    CUDA_CHECK(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    h->prepare_smoother(level);
    if (!h->tmp[level].p) h->tmp[level].alloc(h->n[level]);
    Operator& A = *h->ops[level];
    DevBuf<double> scratch_u, scratch_c, scratch_c2;
    scratch_u.alloc(h->n[level]);
    CUDA_CHECK(cudaMemcpyAsync(scratch_u.p, h->u[level].p, sizeof(double) * h->n[level],
                               cudaMemcpyDeviceToDevice, s));
    if (kind >= 2) {
      scratch_c.alloc(h->n[level + 1]);
      scratch_c.zero(s);
      scratch_c2.alloc(h->n[level + 1]);
    }
    auto once = [&] {
      switch (kind) {
        case 0:
          if (h->opt.smoother == AMGB_SMOOTHER_JACOBI)
            A.jacobi(scratch_u.p, h->f[level].p, h->opt.omega, h->tmp[level].p, s);
          else if (h->opt.smoother == AMGB_SMOOTHER_COLOR_GS)
            for (int c = 0; c < A.n_colors; ++c) A.color_pass(c, h->f[level].p, scratch_u.p, s);
          else
            A.gs_forward(h->f[level].p, scratch_u.p, s);
          break;
        case 1:
          A.residual(scratch_u.p, h->f[level].p, h->tmp[level].p, s);
          break;
        case 2:
          A.residual_restrict(scratch_u.p, h->f[level].p, scratch_c2.p, scratch_c.p, (int)h->n[level + 1], s);
          break;
        case 3:
          LAUNCH(dev::k_prolong_add, blocks_for(h->n[level], 256), 256, 0, s, scratch_c.p,
                 (int)h->n[level + 1], scratch_u.p, (int)h->n[level]);
          break;
        default:
          throw std::invalid_argument("unknown kernel kind");
      }
    };
    for (int i = 0; i < warmup; ++i) once();
    cudaEvent_t e0, e1;
    CUDA_CHECK(cudaEventCreate(&e0));
    CUDA_CHECK(cudaEventCreate(&e1));
    CUDA_CHECK(cudaEventRecord(e0, s));
    for (int i = 0; i < reps; ++i) once();
    CUDA_CHECK(cudaEventRecord(e1, s));
    CUDA_CHECK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ms_out = (double)ms / reps;
  });
}

}  // extern "C"

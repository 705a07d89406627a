// TEST INFRASTRUCTURE ONLY -- never linked into libamgb.so.
// Runs the body of algebraic-multigrid_b200/csrc/mid_levels.cuh (multi-level fused legs of the mid
// levels) with a serial host Env: the threads of a phase run one after the other and every block
// gets freshly poisoned "shared memory", so the range logic (nested tiles, halos that shrink stage
// by stage, hand-over between levels) is checked against the oracle on a CPU-only box.  The GPU
// kernels k_mid_down / k_mid_up instantiate the very same body.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <vector>

#include "../../algebraic-multigrid_b200/csrc/mid_levels.cuh"

namespace {
struct HostEnv {
  std::vector<double> sm;
  int nthreads = 64;
  double* smem() { return sm.data(); }
  template <class F>
  void phase(F&& f) {
    for (int t = 0; t < nthreads; ++t) f(t, nthreads);
  }
  void sync() {}
  // bulk copies land when they are issued; alignment rules of the device TMA path are checked
  int bad = 0;
  void bulk_start(int) {}
  void bulk_copy(double* dst, const double* src, int doubles) {
    if ((reinterpret_cast<uintptr_t>(src) & 15) || ((dst - sm.data()) & 1) || (doubles & 1) || doubles <= 0) ++bad;
    std::memcpy(dst, src, sizeof(double) * (size_t)doubles);
  }
  void bulk_wait() {}
};
}  // namespace

extern "C" {
// levels: n_lv mid levels described by flat arrays; vectors are caller-owned, full length.
//   n[i], nd[i], off[i*10 + d], ld[i], val[i] (pointer table), f[i], u[i], tmp[i]
// returns 0, or -1 when no tile size fits; tile/blocks report the layout chosen.
int mid_host_run(int n_lv, int nu, int first_is_level0, double omega, const int* n, const int* nd, const int* off,
                 const int* ld, double** val, double** f, double** u, double** tmp, double* f_next,
                 const double* u_next, int n_next, int fast, int do_up, int cap_doubles, int force_T, int* tile,
                 int* blocks) {
  using namespace amgb::mid;
  Params P{};
  P.n_lv = n_lv;
  P.nu = nu;
  P.first_is_level0 = first_is_level0;
  P.omega = omega;
  for (int i = 0; i < n_lv; ++i) {
    Level& V = P.lv[i];
    V.n = n[i];
    V.n_coarse = (i + 1 < n_lv) ? n[i + 1] : n_next;
    V.nd = nd[i];
    V.w = 0;
    V.diag_d = -1;
    for (int d = 0; d < nd[i]; ++d) {
      V.off[d] = off[i * kMaxDiag + d];
      if (V.off[d] == 0) V.diag_d = d;
      V.w = std::max(V.w, std::abs(V.off[d]));
    }
    if (V.diag_d < 0) return -2;
    V.ld = ld[i];
    V.val = val[i];
    V.f = f[i];
    V.u = u[i];
    V.tmp = tmp[i];
  }
  P.f_next = f_next;
  P.u_next = u_next;
  P.n_next = n_next;
  if (force_T > 0) {  // exercise small tiles (no shared-memory cap on the host)
    if (!plan_layout(P, 1 << 28, force_T)) return -1;
  } else if (!plan_layout(P, cap_doubles)) {
    return -1;
  }
  if (tile) *tile = P.T;
  if (blocks) *blocks = P.n_blocks;
  HostEnv env;
  const int words = std::max(P.smem_doubles_down, P.smem_doubles_up) + 2;
  for (int b = 0; b < P.n_blocks; ++b) {
    env.sm.assign(words, std::numeric_limits<double>::quiet_NaN());  // stale reads show up as NaN
    if (reinterpret_cast<uintptr_t>(env.sm.data()) & 15) return -3;
    if (do_up) {
      if (fast) run_up<true>(P, b, env);
      else run_up<false>(P, b, env);
    } else {
      if (fast) run_down<true>(P, b, env);
      else run_down<false>(P, b, env);
    }
  }
  return env.bad ? -4 : 0;
}
}

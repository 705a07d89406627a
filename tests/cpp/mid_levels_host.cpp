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
  if (!plan_layout(P, cap_doubles)) return -1;
  if (force_T > 0) {  // exercise small tiles: recompute the layout for this tile size only
    Params Q = P;
    bool ok = false;
    for (int cap = cap_doubles; !ok && cap < (1 << 28); cap <<= 1) {
      Q = P;
      // plan_layout walks T downwards from 1024; emulate a fixed T by shrinking the first level's view
      Q.T = force_T;
      Q.n_blocks = (Q.lv[0].n + force_T - 1) / force_T;
      int len_d[kMaxLevels] = {0}, len_u[kMaxLevels] = {0};
      for (int b = 0; b < Q.n_blocks; ++b) {
        Plan pl;
        make_plan(Q, b, pl);
        for (int i = 0; i < n_lv; ++i) {
          len_d[i] = std::max(len_d[i], pl.in0[i].hi - pl.in0[i].lo + 1);
          len_u[i] = std::max(len_u[i], pl.inp[i].hi - pl.inp[i].lo + 1);
        }
      }
      int od = 0, ou = 0;
      for (int i = 0; i < n_lv; ++i) {
        Q.lv[i].len_down = (len_d[i] + 1) & ~1;
        Q.lv[i].off_down = od;
        od += 3 * Q.lv[i].len_down;
        Q.lv[i].len_up = (len_u[i] + 1) & ~1;
        Q.lv[i].off_up = ou;
        ou += 3 * Q.lv[i].len_up;
      }
      Q.smem_doubles_down = od;
      Q.smem_doubles_up = ou;
      ok = true;
    }
    P = Q;
  }
  if (tile) *tile = P.T;
  if (blocks) *blocks = P.n_blocks;
  HostEnv env;
  const int words = std::max(P.smem_doubles_down, P.smem_doubles_up) + 2;
  for (int b = 0; b < P.n_blocks; ++b) {
    env.sm.assign(words, std::numeric_limits<double>::quiet_NaN());  // stale reads show up as NaN
    if (do_up) {
      if (fast) run_up<true>(P, b, env);
      else run_up<false>(P, b, env);
    } else {
      if (fast) run_down<true>(P, b, env);
      else run_down<false>(P, b, env);
    }
  }
  return 0;
}
}

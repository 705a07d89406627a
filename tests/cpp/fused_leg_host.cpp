// TEST INFRASTRUCTURE ONLY -- never linked into libamgb.so.
// Runs the tile body of algebraic-multigrid_b200/csrc/fused_leg.cuh with a serial host Env
// (threads of a phase run one after the other, TMA copies are memcpy) so the tiling logic of
// the fused V-cycle legs -- line structure, rings, halos, TMA windows -- is checked against
// the oracle on a CPU-only box.  The GPU kernel instantiates the very same body.
//   tma_mode 0: copies land when issued (catches slots overwritten while still in use)
//   tma_mode 1: copies land when the mbarrier is waited on (catches reads before the wait)
#include <cmath>
#include <cstdio>
#include <limits>
#include <type_traits>
#include <vector>

#include "../../algebraic-multigrid_b200/csrc/fused_leg.cuh"

namespace {

struct Pending {
  void* dst;
  const void* src;
  uint32_t bytes;
};

struct HostEnv {
  std::vector<double> sm;
  int nthreads = 0;
  int tma_mode = 0;
  std::vector<std::vector<Pending>> pending;  // per barrier
  std::vector<long long> expected, issued;
  int bad = 0;

  double* smem() { return sm.data(); }
  template <class F>
  void phase(F&& f) {
    for (int t = 0; t < nthreads; ++t) f(t);
  }
  template <class F>
  void single(F&& f) {
    f();
  }
  void init_bars(int n) {
    pending.assign(n, {});
    expected.assign(n, 0);
    issued.assign(n, 0);
  }
  void expect(int q, uint32_t bytes) {
    if (!pending[q].empty() || expected[q] != issued[q]) ++bad;  // previous use not consumed
    expected[q] = bytes;
    issued[q] = 0;
  }
  void tma(void* dst, const void* src, uint32_t bytes, int q) {
    if ((reinterpret_cast<uintptr_t>(dst) & 15) || (reinterpret_cast<uintptr_t>(src) & 15) || (bytes & 15) ||
        bytes == 0)
      ++bad;  // TMA alignment rules
    issued[q] += bytes;
    if (tma_mode == 0) std::memcpy(dst, src, bytes);
    else pending[q].push_back(Pending{dst, src, bytes});
  }
  void wait(int q, int /*parity*/) {
    if (expected[q] != issued[q]) ++bad;
    for (const Pending& p : pending[q]) std::memcpy(p.dst, p.src, p.bytes);
    pending[q].clear();
  }
  void fence_async() {}
};

template <int ND, int NS>
int run_all(const amgb::leg::Plan& pl, int tma_mode) {
  HostEnv env;
  env.nthreads = pl.threads;
  env.tma_mode = tma_mode;
  int bad = 0;
  for (int tile = 0; tile < pl.tiles; ++tile) {
    // 16-byte aligned shared memory, poisoned so stale reads show up as NaN
    env.sm.assign(pl.smem_bytes / 8 + 2, std::numeric_limits<double>::quiet_NaN());
    amgb::leg::run_tile<ND, NS>(pl.P, tile, env);
    bad += env.bad;
    env.bad = 0;
  }
  return bad;
}

}  // namespace

extern "C" {

// info[0..9] = ok, m, rho, W, LJ, tiles, n_strips, PF, threads, smem_bytes
int leg_host_plan(int kind, int n_sweeps, int n, int nd, const int* off, int n_sm, int smem_cap, int W_ovr,
                  int LJ_ovr, int PF_ovr, int force_single, long long* info) {
  amgb::leg::Plan pl =
      amgb::leg::plan_leg(kind, n_sweeps, n, nd, off, n_sm, (size_t)smem_cap, W_ovr, LJ_ovr, PF_ovr, force_single);
  info[0] = pl.ok;
  info[1] = pl.P.m;
  info[2] = pl.P.rho;
  info[3] = pl.P.W;
  info[4] = pl.P.LJ;
  info[5] = pl.tiles;
  info[6] = pl.P.n_strips;
  info[7] = pl.P.PF;
  info[8] = pl.threads;
  info[9] = (long long)pl.smem_bytes;
  return pl.ok ? 0 : 1;
}

// All vectors must be padded by at least 2 entries past their logical end (the TMA windows are
// rounded to 16 bytes).  Returns 0 on success, 1 when no plan exists, 2 on a protocol violation.
int leg_host_run(int kind, int n_sweeps, int n, int nd, const int* off, int ld, const double* val,
                 const double* f, const double* uin, const double* e, int n_coarse, double omega, double* uout,
                 double* fc, int n_sm, int smem_cap, int W_ovr, int LJ_ovr, int PF_ovr, int force_single,
                 int tma_mode) {
  amgb::leg::Plan pl =
      amgb::leg::plan_leg(kind, n_sweeps, n, nd, off, n_sm, (size_t)smem_cap, W_ovr, LJ_ovr, PF_ovr, force_single);
  if (!pl.ok) return 1;
  amgb::leg::Params& P = pl.P;
  P.ld = ld;
  P.val = val;
  P.f = f;
  P.uin = uin;
  P.e = e;
  P.n_coarse = n_coarse;
  P.omega = omega;
  P.uout = uout;
  P.fc = fc;
  int bad;
  auto go = [&](auto nd_tag) {
    constexpr int ND = decltype(nd_tag)::value;
    switch (P.NS) {
      case 1: return run_all<ND, 1>(pl, tma_mode);
      case 2: return run_all<ND, 2>(pl, tma_mode);
      case 3: return run_all<ND, 3>(pl, tma_mode);
      default: return run_all<ND, 4>(pl, tma_mode);
    }
  };
  if (nd <= 6) bad = go(std::integral_constant<int, 6>());
  else if (nd <= 10) bad = go(std::integral_constant<int, 10>());
  else bad = go(std::integral_constant<int, 16>());
  return bad ? 2 : 0;
}
}

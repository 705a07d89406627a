// TEST INFRASTRUCTURE ONLY -- never linked into libamgb.so.
// Runs the lane logic of algebraic-multigrid_b200/csrc/gs_wave.cuh (multi-SM lexicographic
// Gauss-Seidel) on the CPU: every block is a 32-lane warp stepped in lockstep, the blocks are
// interleaved in a RANDOM order that respects only what the device guarantees (a block waits while
// the hand-over value its lane 0 needs is still the sentinel), and every lane reads its inputs PD
// steps ahead through the same register ring as the kernel.  The result must equal the plain
// row-by-row sweep bit for bit.  The GPU kernel k_gs_wave instantiates the very same step functions,
// packing function and plan.
#include <cstdint>
#include <cstring>
#include <random>
#include <vector>

#include "../../algebraic-multigrid_b200/csrc/gs_wave.cuh"

namespace {
using namespace amgb::gsw;

struct Geometry {
  int n, m, n_lines, n_blocks, T, Ts, S, dir, pd;
};

template <int S, int DIR, unsigned MASK>
struct Emu {
  Geometry G;
  const double* coef;
  const double* f;
  double* u;
  std::vector<double> hand;
  struct Block {
    int t = 0;
    bool loaded = false;  // the inputs of step t have been taken out of the ring
    std::vector<Lane<S>> L;
    std::vector<In> ring;  // [lane * pd + r]
    std::vector<In> cur;   // inputs of the step in flight
  };
  std::vector<Block> blocks;

  In load(int b, int j, int tt) const {
    In in{};
    for (int e = 0; e < kSlots; ++e)
      in.c[e] = ((MASK >> e) & 1u) ? coef[(((size_t)b * G.Ts + tt) * kPacked + e) * kLanes + j] : 0.0;
    const int Y = line_of(b, j), X = tt - S * j;
    const int Yc = Y < 0 ? 0 : (Y >= G.n_lines ? G.n_lines - 1 : Y);
    auto clampx = [&](int x) { return x < 0 ? 0 : (x >= G.m ? G.m - 1 : x); };
    in.f = f[mem_index(G.n, G.m, DIR, Yc, clampx(X))];
    in.uo = u[mem_index(G.n, G.m, DIR, Yc, clampx(X + S + 1))];
    in.sup = (j == 0 && b > 0) ? hand[(size_t)(b - 1) * G.m + clampx(X)] : 0.0;
    return in;
  }
  void init() {
    double nan_all;
    std::memset(&nan_all, 0xFF, 8);
    hand.assign((size_t)G.n_blocks * G.m, nan_all);
    blocks.resize(G.n_blocks);
    for (int b = 0; b < G.n_blocks; ++b) {
      Block& B = blocks[b];
      B.L.resize(kLanes);
      for (auto& l : B.L) l.clear();
      B.ring.resize((size_t)kLanes * G.pd);
      B.cur.resize(kLanes);
      for (int j = 0; j < kLanes; ++j)
        for (int r = 0; r < G.pd; ++r) B.ring[(size_t)j * G.pd + r] = load(b, j, r);
    }
  }
  int t_end() const { return (G.T + G.pd) / (G.pd + 1) * (G.pd + 1); }
  // one step of block b; false: the block has to wait (nothing happened)
  bool step(int b) {
    Block& B = blocks[b];
    const int t = B.t;
    if (!B.loaded) {
      for (int j = 0; j < kLanes; ++j) B.cur[j] = B.ring[(size_t)j * G.pd + t % G.pd];
      B.loaded = true;
    }
    if (b > 0 && t < G.m) {  // lane 0 is inside its line for t < m
      if (is_sentinel(B.cur[0].sup)) B.cur[0].sup = hand[(size_t)(b - 1) * G.m + t];
      if (is_sentinel(B.cur[0].sup)) return false;
    }
    double up[kLanes], down[kLanes], res[kLanes];
    for (int j = 0; j < kLanes; ++j) {
      down[j] = step_begin<S>(B.L[j], B.cur[j]);
      up[j] = B.L[j].out;
    }
    for (int j = 0; j < kLanes; ++j)
      res[j] = step_finish<S, DIR, MASK>(B.L[j], B.cur[j], j > 0 ? up[j - 1] : up[j], j + 1 < kLanes ? down[j + 1] : down[j]);
    for (int j = 0; j < kLanes; ++j) {
      const int Y = line_of(b, j), X = t - S * j;
      const bool line_ok = Y >= 0 && Y < G.n_lines;
      const bool inside = X >= 0 && X < G.m;
      double out = 0.0;
      if (j >= 1 && j <= kLinesPerBlock && line_ok && inside) {
        out = res[j];
        u[mem_index(G.n, G.m, DIR, Y, X)] = out;
        if (j == kLinesPerBlock && b + 1 < G.n_blocks) hand[(size_t)b * G.m + X] = out;
      } else if (j == 0 && b > 0 && inside) {
        out = B.cur[0].sup;
      }
      B.L[j].out = out;
    }
    // like the kernel: the look-ahead loads of step t + PD are issued at the END of step t
    for (int j = 0; j < kLanes; ++j) B.ring[(size_t)j * G.pd + t % G.pd] = load(b, j, t + G.pd);
    B.t += 1;
    B.loaded = false;
    return true;
  }
  // 0: done, -1: dead-lock (no block can move)
  int run(unsigned seed) {
    init();
    std::mt19937 rng(seed);
    const int end = t_end();
    for (;;) {
      std::vector<int> live;
      for (int b = 0; b < G.n_blocks; ++b)
        if (blocks[b].t < end) live.push_back(b);
      if (live.empty()) return 0;
      // favour the LAST live blocks half of the time: they then run as close behind their
      // predecessors as the hand-over allows, which is where a read-after-write slip would show
      bool moved = false;
      for (int attempt = 0; attempt < 4 * (int)live.size() && !moved; ++attempt) {
        int pickb;
        if (rng() & 1u) pickb = live[live.size() - 1 - (rng() % ((live.size() + 1) / 2))];
        else pickb = live[rng() % live.size()];
        moved = step(pickb);
      }
      if (!moved) {
        for (int b : live)
          if (step(b)) {
            moved = true;
            break;
          }
        if (!moved) return -1;
      }
    }
  }
};

template <int S, int DIR, unsigned MASK>
int run_emu(const Geometry& G, const double* coef, const double* f, double* u, unsigned seed) {
  Emu<S, DIR, MASK> E;
  E.G = G;
  E.coef = coef;
  E.f = f;
  E.u = u;
  return E.run(seed);
}
}  // namespace

extern "C" {
// One sweep (dir = +1 forward, -1 backward) of the operator given as DIA rows: val[d * ld + row],
// offsets off[].  Returns 0, -1 on dead-lock, -2 when the operator is not grid-structured.
int gsw_host_sweep(int n, int n_diag, const int* off, int ld, const double* val, const double* f, double* u, int dir,
                   int pd, unsigned seed, int* S_out, int* m_out) {
  int cand[3];
  plan_candidates(off, n_diag, cand);
  for (int ci = 0; ci < 3; ++ci) {
    Plan P = plan_for(off, n_diag, n, cand[ci]);
    if (!P.ok) continue;
    // the device check (k_gsw_check), restated
    unsigned mask = 0;
    bool bad = false;
    for (int k = 0; k < n && !bad; ++k)
      for (int d = 0; d < n_diag; ++d) {
        if (val[(size_t)d * ld + k] == 0.0) continue;
        const int e = P.e_of[d], a = e / 3 - 1, dl = e % 3 - 1, y = k / P.m, x = k % P.m;
        if (x + dl < 0 || x + dl >= P.m || y + a < 0 || y + a >= P.n_lines) bad = true;
        mask |= 1u << e;
      }
    if (bad) continue;
    Geometry G{};
    G.n = n;
    G.m = P.m;
    G.n_lines = P.n_lines;
    G.n_blocks = (P.n_lines + kLinesPerBlock - 1) / kLinesPerBlock;
    G.S = stride_for(mask, dir);
    G.T = P.m + G.S * 31;
    G.Ts = padded_steps(G.T);
    if (pd + 1 > kMaxLookAhead) return -3;
    G.dir = dir;
    G.pd = pd;
    Dia9 A{};
    A.val = val;
    A.ld = ld;
    for (int e = 0; e < kSlots; ++e) A.d_of[e] = P.d_of[e];
    std::vector<double> coef(packed_doubles(G.n_blocks, G.T), 0.0);
    for (int b = 0; b < G.n_blocks; ++b)
      for (int t = 0; t < G.T; ++t)
        for (int e = 0; e < kSlots; ++e)
          for (int j = 0; j < kLanes; ++j)
            coef[(((size_t)b * G.Ts + t) * kPacked + e) * kLanes + j] =
                packed_coef(A, n, P.m, P.n_lines, G.S, dir, b, t, j, e);
    if (S_out) *S_out = G.S;
    if (m_out) *m_out = P.m;
    // the same choice of instantiation as the device dispatch (Operator::launch_wave)
    if ((mask & ~kMaskFive) == 0)
      return dir > 0 ? run_emu<1, 1, kMaskFive>(G, coef.data(), f, u, seed)
                     : run_emu<1, -1, kMaskFive>(G, coef.data(), f, u, seed);
    if (G.S == 1)
      return dir > 0 ? run_emu<1, 1, kMaskAll>(G, coef.data(), f, u, seed)
                     : run_emu<1, -1, kMaskAll>(G, coef.data(), f, u, seed);
    return dir > 0 ? run_emu<2, 1, kMaskAll>(G, coef.data(), f, u, seed)
                   : run_emu<2, -1, kMaskAll>(G, coef.data(), f, u, seed);
  }
  return -2;
}

// The sweep as the reference defines it (smoother.hpp:119-176 on the row mirror): rows in order,
// sigma over the stored entries in ascending column order, zero diagonal = row left alone.
void gsw_host_reference(int n, int n_diag, const int* off, int ld, const double* val, const double* f, double* u,
                        int dir) {
  for (int i = 0; i < n; ++i) {
    const int k = dir > 0 ? i : n - 1 - i;
    double rsum = 0.0, diag = 0.0;
    for (int d = 0; d < n_diag; ++d) {  // offsets ascending
      const double a = val[(size_t)d * ld + k];
      if (a == 0.0) continue;
      const int c = k + off[d];
      if (c == k) diag = a;
      else rsum = rsum + a * u[c];
    }
    if (diag != 0.0) u[k] = (f[k] - rsum) / diag;
  }
}
}

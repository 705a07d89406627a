// Fused V-cycle legs for the damped-Jacobi cycle on banded (DIA) operators: each level's
// down leg  (pre-smoothing sweeps -> residual -> restriction, multigrid.hpp:268-282)  and
// up leg    (coarse-grid correction -> post-smoothing sweeps, multigrid.hpp:294-301)
// run as ONE kernel that reads the operator, f and the input vector from HBM exactly once.
//
// How: rows are grouped into "lines" of m consecutive rows so that every diagonal offset is
// a*m + delta with a in {-1,0,1} and |delta| <= rho (level 0 of an n x n grid: m = n,
// offsets {-n,-1,0,1,n}; Galerkin levels: m ~ n/2^l with far offsets m-1..m+1).  A CTA owns a
// tile of W elements x LJ lines and streams over its lines; a chain of NS dependent stencil
// stages runs one line behind the other (stage s on line jj-s+1 at step jj), every stage
// reading its predecessor's last three lines from a shared-memory ring.  The operator rows,
// f and the input vector of a line arrive PF lines ahead by TMA bulk copies (cp.async.bulk +
// mbarrier) into a ring of NS+1+PF line slots.  Stage s needs its predecessor on one more
// line / rho more elements on every side, which the tile computes redundantly (a few per
// cent), so tiles never talk to each other.  Narrow-band coarse levels use the same code
// with a single line (m = n): plain overlapped 1-D tiling with halo rho = half-bandwidth.
//
// The per-row arithmetic (operation order, no FMA contraction) is exactly that of
// k_jacobi / k_jacobi_zero / k_residual_restrict / k_prolong_add in kernels.cuh, so the
// fused cycle is bit-identical to the unfused one and to the oracle.
//
// The tile body is written once, as phases separated by block barriers, against an `Env`
// that supplies threads / barriers / TMA.  The CUDA kernel instantiates it with the device
// Env; tests/cpp/fused_leg_host.cpp instantiates it with a serial host Env so the tiling
// logic can be checked on a CPU-only box (test infrastructure, never a product path).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>

#if defined(__CUDACC__)
#define AMGB_LEG_FN __device__ __forceinline__
#else
#define AMGB_LEG_FN inline
#endif

namespace amgb {
namespace leg {

constexpr int kMaxDiag = 16;
enum Kind {
  DOWN_U = 0,     // input = u (level 0): nu Jacobi sweeps, residual, restriction
  DOWN_ZERO = 1,  // input = omega * f / d (first sweep from the zero guess of a coarse level,
                  // multigrid.hpp:278): nu - 1 more sweeps, residual, restriction
  UP = 2          // input = u + P e (multigrid.hpp:294-296): nu sweeps
};

struct Params {
  int kind, NS;  // NS dependent stencil stages: DOWN_U nu+1, DOWN_ZERO nu, UP nu
  int n;         // rows of the level
  int m;         // line length (m >= n: single line)
  int n_lines;   // ceil(n / m)
  int rho;       // element halo one stencil stage needs
  int x;         // 1 for down legs (the restriction reads r[k-1], r[k+1])
  int H;         // rho * NS + x: element halo of a tile
  int W, Wext, LS, LSe;  // strip width, W + 2H, line strides of the shared-memory arrays
  int n_strips, LJ, n_jchunks;
  int PF, NA;    // lines in flight; operator-ring slots = 2 NS + 1 + PF
  int nd, diag_d, ld;
  int line_a[kMaxDiag], delta[kMaxDiag];  // off[d] = line_a[d] * m + delta[d]
  int n_coarse;
  int o_A, o_stg, o_estg, o_vr, o_r;      // shared-memory layout, in doubles
  double omega;
  const double* val;  // DIA values, val[d * ld + row]
  const double* f;
  const double* uin;  // DOWN_U: u; UP: the down leg's result; DOWN_ZERO: unused
  const double* e;    // UP: coarse correction
  double* uout;       // smoothed iterate
  double* fc;         // down legs: coarse right-hand side
};

// ---- arithmetic in the reference's order, never contracted -------------------------------
AMGB_LEG_FN double fmul(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
AMGB_LEG_FN double fadd(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}
AMGB_LEG_FN double fsub(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dsub_rn(a, b);
#else
  return a - b;
#endif
}
AMGB_LEG_FN double fdiv(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __ddiv_rn(a, b);
#else
  return a / b;
#endif
}

// (P e)[i] with e[J - e_first] = coarse entry J (interpolator.hpp:52-56,118-125)
AMGB_LEG_FN double prolong_entry(const double* e, int e_first, int n_coarse, long long i) {
  double acc = 0.0;
  const int J = (int)(i >> 1);
  if (i & 1) {
    if (J >= 0 && J < n_coarse) acc = fadd(acc, fmul(1.0, e[J - e_first]));
  } else {
    if (J - 1 >= 0 && J - 1 < n_coarse) acc = fadd(acc, fmul(0.5, e[J - 1 - e_first]));
    if (J >= 0 && J < n_coarse) acc = fadd(acc, fmul(0.5, e[J - e_first]));
  }
  return acc;
}

// One tile.  env.phase(fn) runs fn(t) for all threads t and ends with a block barrier.
//
// Schedule: ONE phase per step.  At step jj the input stage handles line jj + 2, stencil stage
// s line jj - 2 (s - 1) and the restriction line jj - 2 (NS - 1) - 1, so everything a stage
// reads was written in an earlier step; a thread first evaluates all its stages into
// registers (independent dependency chains -> instruction-level parallelism) and stores last.
// Per-line bookkeeping (first row, ring slots, barrier parity) is carried incrementally from
// step to step -- no divisions in the loop.  Row numbers fit in int (the planner checks).
template <int ND, int NS, class Env>
AMGB_LEG_FN void run_tile(const Params& P, int tile, Env& env) {
  double* sm = env.smem();
  double* Aring = sm + P.o_A;    // [NA][nd + 1][LS]   (index nd = f)
  double* stg = sm + P.o_stg;    // [PF + 1][LS]       input vector of a line (DOWN_U, UP)
  double* estg = sm + P.o_estg;  // [PF + 1][LSe]      coarse entries of a line (UP)
  double* vr = sm + P.o_vr;      // [NS][4][LS]        ring s = output of stage s (0 = input)
  double* rbuf = sm + P.o_r;     // [2][LS]            residual lines (down legs)

  const int strip = tile % P.n_strips, chunk = tile / P.n_strips;
  const int i0 = strip * P.W, i1 = (i0 + P.W < P.m) ? i0 + P.W : P.m;
  const int j0 = chunk * P.LJ, j1 = (j0 + P.LJ < P.n_lines) ? j0 + P.LJ : P.n_lines;
  const int jA = j0 - NS, jB = j1 + NS;  // lines the input stage covers
  const int Wown = i1 - i0;
  const int nd = P.nd, LS = P.LS, Wext = P.Wext, n = P.n, m = P.m, NA = P.NA, NQ = P.PF + 1;
  const int kind = P.kind;
  const int s_out = (kind == UP) ? NS : NS - 1;  // stage whose result is the smoothed iterate
  const bool has_stg = (kind != DOWN_ZERO);
  const int e0 = i0 - P.H;                       // element 0 of line j is global row j * m + e0
  const int slotA = (nd + 1) * LS;               // doubles per operator-ring slot

  // first row of the 16-byte aligned window [alo, ahi) the TMA copies for a line starting at g
  auto win_lo = [&](int g) -> int { return (g > 0 ? g : 0) & ~1; };
  auto issue = [&](int g, int q, int sq) {  // one thread: start the copies of the line at row g
    const int lo = g > 0 ? g : 0, hi = (g + Wext < n) ? g + Wext : n;
    uint32_t bytes = 0, ebytes = 0;
    int alo = 0, calo = 0;
    if (lo < hi) {
      alo = lo & ~1;
      const int ahi = (hi + 1) & ~1;
      bytes = (uint32_t)(ahi - alo) * 8u;
      if (kind == UP) {  // coarse entries the prolongation reads: J in [(lo >> 1) - 1, (hi - 1) >> 1]
        int Jlo = (alo >> 1) - 1, Jhi = ((ahi - 1) >> 1) + 1;
        if (Jlo < 0) Jlo = 0;
        if (Jhi > P.n_coarse) Jhi = P.n_coarse;
        if (Jlo < Jhi) {
          calo = Jlo & ~1;
          ebytes = (uint32_t)(((Jhi + 1) & ~1) - calo) * 8u;
        }
      }
    }
    env.expect(q, bytes * (uint32_t)(nd + 1 + (has_stg ? 1 : 0)) + ebytes);
    if (bytes) {
      double* dst = Aring + (size_t)q * slotA;
      for (int d = 0; d < nd; ++d) env.tma(dst + d * LS, P.val + (size_t)d * P.ld + alo, bytes, q);
      env.tma(dst + nd * LS, P.f + alo, bytes, q);
      if (has_stg) env.tma(stg + (size_t)sq * LS, P.uin + alo, bytes, q);
    }
    if (ebytes) env.tma(estg + (size_t)sq * P.LSe, P.e + calo, ebytes, q);
  };
  auto e_lo = [&](int alo) -> int {  // first coarse entry staged for a line whose window starts at alo
    int Jlo = (alo >> 1) - 1;
    if (Jlo < 0) Jlo = 0;
    return Jlo & ~1;
  };

  env.init_bars(NA);
  // issue cursor: the next line to request
  int ji = jA, gi = jA * m + e0, qi = 0, sqi = 0;
  auto issue_next = [&] {  // advances the cursor on every thread, copies started by one
    if (ji < jB) env.single([&] { issue(gi, qi, sqi); });
    ++ji;
    gi += m;
    if (++qi == NA) qi = 0;
    if (++sqi == NQ) sqi = 0;
  };
  for (int k = 0; k < P.PF; ++k) issue_next();

  // input-stage cursor: line jj + 2, its ring slots and barrier parity
  int q0 = 0, sq0 = 0, par0 = 0;
  int g0 = jA * m + e0;  // first row of line jj + 2
  for (int jj = jA - 2; jj <= j1 + 2 * NS - 2; ++jj) {
    // The slot line jj + 2 + PF goes into was last READ (generic proxy) before the block barrier
    // that ended the previous step, so the copy may overwrite it: a write-after-read across
    // proxies needs only that ordering, no proxy fence.
    issue_next();
    const bool on0 = (jj + 2 < jB);
    if (on0) env.wait(q0, par0);
    const int jr0 = jj + 2 - jA;             // line index relative to jA (>= 0)
    const int alo0 = win_lo(g0);
    const double* Aq0 = Aring + (size_t)q0 * slotA;
    const int calo0 = e_lo(alo0);
    const int jR = jj - 2 * (NS - 1) - 1;    // line the restriction handles
    const bool r_on = (kind != UP) && jR >= j0 && jR < j1;
    const int gR = g0 - (2 * NS + 1) * m;

    env.phase([&](int t) {
      if (t >= Wext) return;
      const bool own_t = (t >= P.H && t < P.H + Wown);
      // ---- evaluate (loads + arithmetic only)
      double o0 = 0.0;
      bool v0 = false;
      if (on0) {
        const int k = g0 + t;
        v0 = (k >= 0 && k < n);
        if (v0) {
          const int pos = k - alo0;
          if (kind == DOWN_U) {
            o0 = stg[(size_t)sq0 * LS + pos];
          } else if (kind == DOWN_ZERO) {
            const double d = Aq0[P.diag_d * LS + pos];
            o0 = (d == 0.0) ? 0.0 : fmul(P.omega, fdiv(Aq0[nd * LS + pos], d));
          } else {
            o0 = fadd(stg[(size_t)sq0 * LS + pos], prolong_entry(estg + (size_t)sq0 * P.LSe, calo0, P.n_coarse, k));
          }
        }
      }
      double os[NS + 1];
      bool vs[NS + 1], as[NS + 1];
#pragma unroll
      for (int s = 1; s <= NS; ++s) {
        os[s] = 0.0;
        vs[s] = false;
        const int j = jj - 2 * (s - 1), rem = NS - s;
        const int margin = P.rho * s;
        as[s] = (j >= j0 - rem && j < j1 + rem) && t >= margin && t < Wext - margin;
        if (as[s]) {
          const int g = g0 - 2 * s * m;
          const int k = g + t;
          vs[s] = (k >= 0 && k < n);
          if (vs[s]) {
            const int pos = k - win_lo(g);
            int q = q0 - 2 * s;
            if (q < 0) q += NA;
            const double* Aq = Aring + (size_t)q * slotA;
            const double* X = vr + (size_t)(s - 1) * 4 * LS;
            const int jr = jr0 - 2 * s;  // >= 1
            double acc = Aq[nd * LS + pos], diag = 0.0;
#pragma unroll
            for (int d = 0; d < ND; ++d) {
              if (d < nd) {
                const double a = Aq[d * LS + pos];
                if (a != 0.0) {
                  const double xv = X[(size_t)((jr + P.line_a[d]) & 3) * LS + t + P.delta[d]];
                  acc = fsub(acc, fmul(a, xv));
                  if (d == P.diag_d) diag = a;
                }
              }
            }
            if (kind != UP && s == NS) {
              os[s] = acc;  // residual
            } else {
              const double xc = X[(size_t)(jr & 3) * LS + t];
              os[s] = (diag == 0.0) ? xc : fadd(xc, fmul(P.omega, fdiv(acc, diag)));
            }
          }
        }
      }
      double oc = 0.0;
      int Jc = -1;
      if (r_on && own_t) {
        const int k = gR + t;
        if (k >= 0 && k < n && (k & 1)) {
          const int J = (k - 1) >> 1;
          if (J < P.n_coarse) {
            const double* rb = rbuf + (size_t)(jR & 1) * LS;
            const double a = fmul(0.5, rb[t - 1]);
            const double b = fadd(a, rb[t]);
            oc = fadd(b, fmul(0.5, rb[t + 1]));
            Jc = J;
          }
        }
      }
      // ---- store
      if (on0) {
        vr[(size_t)(jr0 & 3) * LS + t] = o0;
        if (s_out == 0 && v0 && own_t && jr0 >= NS && jr0 < NS + (j1 - j0)) P.uout[g0 + t] = o0;
      }
#pragma unroll
      for (int s = 1; s <= NS; ++s) {
        if (as[s]) {
          const int jr = jr0 - 2 * s;
          if (kind != UP && s == NS) rbuf[(size_t)((jr + jA) & 1) * LS + t] = os[s];
          else if (s < NS) vr[((size_t)s * 4 + (jr & 3)) * LS + t] = os[s];
          if (s == s_out && vs[s] && own_t && jr >= NS && jr < NS + (j1 - j0)) P.uout[g0 - 2 * s * m + t] = os[s];
        }
      }
      if (Jc >= 0) P.fc[Jc] = oc;
    });
    // advance the input-stage cursor to line jj + 3
    g0 += m;
    if (++q0 == NA) {
      q0 = 0;
      par0 ^= 1;
    }
    if (++sq0 == NQ) sq0 = 0;
  }
}

// ------------------------------------------------------------------------------------------
// Host-side planning: line structure of the operator and the tiling for a given GPU.
struct Plan {
  bool ok = false;
  std::string why;
  Params P{};
  int threads = 0;
  int tiles = 0;
  size_t smem_bytes = 0;
};

inline int round_up(int v, int to) { return (v + to - 1) / to * to; }

inline size_t smem_doubles(Params& P) {
  const bool has_stg = P.kind != DOWN_ZERO, up = P.kind == UP;
  int o = 16;  // mbarriers
  P.o_A = o;
  o += P.NA * (P.nd + 1) * P.LS;
  P.o_stg = o;
  if (has_stg) o += (P.PF + 1) * P.LS;
  P.o_estg = o;
  if (up) o += (P.PF + 1) * P.LSe;
  P.o_vr = o;
  o += P.NS * 4 * P.LS;
  P.o_r = o;
  if (!up) o += 2 * P.LS;
  return (size_t)o;
}

// off[0..nd): ascending diagonal offsets (column - row).  n_sm / smem_cap describe the GPU.
// Overrides (0 = automatic) let the tests force small tiles.
inline Plan plan_leg(int kind, int n_sweeps, int n, int nd, const int* off, int n_sm, size_t smem_cap,
                     int W_override = 0, int LJ_override = 0, int PF_override = 0, int force_single = 0) {
  Plan R;
  Params& P = R.P;
  if (n < 1 || nd < 1 || nd > kMaxDiag || n_sweeps < 1) {
    R.why = "unsupported shape";
    return R;
  }
  P.kind = kind;
  P.NS = (kind == DOWN_U) ? n_sweeps + 1 : n_sweeps;
  if (P.NS > 4) {
    R.why = "too many sweeps to fuse";
    return R;
  }
  P.n = n;
  P.nd = nd;
  P.x = (kind == UP) ? 0 : 1;
  P.diag_d = -1;
  int w = 0;
  for (int d = 0; d < nd; ++d) {
    if (off[d] == 0) P.diag_d = d;
    w = std::max(w, std::abs(off[d]));
  }
  if (P.diag_d < 0) {
    R.why = "no diagonal";
    return R;
  }
  // line structure: the positive offsets split at their largest gap into a near cluster and a
  // far cluster centred on the line length m
  bool line_mode = false;
  int m = n, rho = w;
  if (!force_single) {
    int pos[kMaxDiag], np = 0;
    for (int d = 0; d < nd; ++d)
      if (off[d] > 0) pos[np++] = off[d];
    int neg_max = 0;
    for (int d = 0; d < nd; ++d) neg_max = std::max(neg_max, -off[d]);
    if (np >= 1) {
      int split = 0, gap = pos[0];  // gap before pos[0] counts: everything may be "far"
      for (int i = 1; i < np; ++i)
        if (pos[i] - pos[i - 1] > gap) {
          gap = pos[i] - pos[i - 1];
          split = i;
        }
      const int far_lo = pos[split], far_hi = pos[np - 1];
      const int mc = (far_lo + far_hi) / 2;
      int r = 0;
      bool fits = mc >= 1;
      for (int d = 0; d < nd && fits; ++d) {
        int best = std::abs(off[d]);
        for (int a = -1; a <= 1; a += 2) best = std::min(best, std::abs(off[d] - a * mc));
        r = std::max(r, best);
      }
      if (fits && r <= 8 && mc >= 32 && mc >= 12 * r && mc < n) {
        line_mode = true;
        m = mc;
        rho = r;
      }
    }
  }
  P.m = m;
  P.rho = rho;
  P.n_lines = (n + m - 1) / m;
  for (int d = 0; d < nd; ++d) {
    int a = 0;
    if (line_mode) {
      int best = std::abs(off[d]);
      for (int c = -1; c <= 1; c += 2)
        if (std::abs(off[d] - c * m) < best) {
          best = std::abs(off[d] - c * m);
          a = c;
        }
    }
    P.line_a[d] = a;
    P.delta[d] = off[d] - a * m;
  }
  for (int d = nd; d < kMaxDiag; ++d) P.line_a[d] = P.delta[d] = 0;
  P.H = P.rho * P.NS + P.x;
  if ((long long)n + (long long)(2 * P.NS + 4) * m + 4ll * P.H + 4096 > 2147483647ll) {
    R.why = "row numbers do not fit in int";
    return R;
  }
  if (2 * P.H + 32 > 992) {
    R.why = "band too wide for a one-line tile";
    return R;
  }
  // tiling: widest strip whose rings fit, preferring a deeper prefetch
  const int Wmax_threads = 992 - 2 * P.H;
  int best_W = 0, best_PF = 0;
  for (int PF = (PF_override ? PF_override : 3); PF >= (PF_override ? PF_override : 2) && !best_W; --PF) {
    P.PF = PF;
    P.NA = 2 * P.NS + 1 + PF;
    int lo = 1, hi = std::min(Wmax_threads, m);
    auto fits = [&](int W) {
      P.W = W;
      P.Wext = W + 2 * P.H;
      P.LS = round_up(P.Wext + 2, 2);
      P.LSe = round_up(P.Wext / 2 + 6, 2);
      return smem_doubles(P) * 8 <= smem_cap;
    };
    if (!fits(lo)) continue;
    while (lo < hi) {
      const int mid = (lo + hi + 1) / 2;
      if (fits(mid)) lo = mid;
      else hi = mid - 1;
    }
    if (lo >= std::min(m, 4 * P.H) || PF == 2 || PF_override) {
      best_W = lo;
      best_PF = PF;
    }
  }
  if (!best_W) {
    R.why = "rings do not fit in shared memory";
    return R;
  }
  P.PF = best_PF;
  P.NA = 2 * P.NS + 1 + P.PF;
  int Wcap = best_W;
  if (W_override) Wcap = std::min(Wcap, W_override);
  if (line_mode) {
    // strips x line chunks: strips as wide as the rings allow (a step costs a block barrier
    // whatever its width); among the next few strip counts take the one that minimises
    // (waves of CTAs) x (work of a tile incl. its redundant halo)
    const int ns_min = (m + Wcap - 1) / Wcap;
    double best_cost = 0.0;
    for (int ns = ns_min; ns <= ns_min + 3 && ns <= m; ++ns) {
      const int W = (m + ns - 1) / ns;
      const int per_strip = std::max(1, n_sm / ns);
      const int LJ = std::max((P.n_lines + per_strip - 1) / per_strip, std::min(P.n_lines, 4 * P.NS));
      const int chunks = (P.n_lines + LJ - 1) / LJ;
      const int waves = (ns * chunks + n_sm - 1) / n_sm;
      const double cost = (double)waves * (128 + W + 2 * P.H) * (LJ + 3 * P.NS + 1);  // 128: barrier + bookkeeping of a step
      if (ns == ns_min || cost < best_cost) {
        best_cost = cost;
        P.n_strips = ns;
        P.W = W;
        P.LJ = LJ;
      }
    }
  } else {
    // one line: strips only; aim at one strip per SM but keep the halo below ~1/3 of the strip
    int W = (n + n_sm - 1) / n_sm;
    W = std::max(W, std::min(6 * P.H, Wcap));
    W = std::min(W, Wcap);
    P.n_strips = (n + W - 1) / W;
    P.W = (n + P.n_strips - 1) / P.n_strips;
    P.LJ = 1;
  }
  if (LJ_override) P.LJ = std::min(P.LJ, LJ_override);
  P.LJ = std::max(P.LJ, 1);
  P.n_jchunks = (P.n_lines + P.LJ - 1) / P.LJ;
  P.Wext = P.W + 2 * P.H;
  P.LS = round_up(P.Wext + 2, 2);
  P.LSe = round_up(P.Wext / 2 + 6, 2);
  R.smem_bytes = smem_doubles(P) * 8;
  R.threads = round_up(P.Wext, 32) + 32;  // + the warp that issues the TMA copies
  R.tiles = P.n_strips * P.n_jchunks;
  R.ok = true;
  return R;
}


#if defined(__CUDACC__)
// ------------------------------------------------------------------------------------------
// Device Env: block barriers, mbarriers and TMA 1-D bulk copies (SASS: UBLKCP, SYNCS).
struct DevEnv {
  double* sm;
  __device__ __forceinline__ double* smem() const { return sm; }
  __device__ __forceinline__ uint64_t* bar(int q) const { return reinterpret_cast<uint64_t*>(sm) + q; }
  static __device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
  template <class F>
  __device__ __forceinline__ void phase(F&& f) {
    f((int)threadIdx.x);
    __syncthreads();
  }
  // the copies are issued by lane 0 of the block's last warp, which has no rows of its own
  template <class F>
  __device__ __forceinline__ void single(F&& f) {
    if (threadIdx.x == blockDim.x - 32) f();
  }
  __device__ __forceinline__ void init_bars(int n) {
    if (threadIdx.x == blockDim.x - 32) {
      for (int i = 0; i < n; ++i)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar(i))), "r"(1));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  }
  __device__ __forceinline__ void expect(int q, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar(q))), "r"(bytes)
                 : "memory");
  }
  __device__ __forceinline__ void tma(void* dst, const void* src, uint32_t bytes, int q) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)),
        "l"(src), "r"(bytes), "r"(s32(bar(q)))
        : "memory");
  }
  __device__ __forceinline__ void wait(int q, int parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "LEG_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra LEG_DONE_%=;\n\t"
        "bra LEG_WAIT_%=;\n\t"
        "LEG_DONE_%=:\n\t}" ::"r"(s32(bar(q))),
        "r"((uint32_t)parity)
        : "memory");
  }
  __device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
};

template <int ND, int NS>
__global__ void __launch_bounds__(1024, 1) k_fused_leg(const __grid_constant__ Params P) {
  extern __shared__ __align__(16) double leg_smem[];
  DevEnv env{leg_smem};
  run_tile<ND, NS>(P, (int)blockIdx.x, env);
}
#endif  // __CUDACC__

}  // namespace leg
}  // namespace amgb

// Register-streaming fused V-cycle legs for the damped-Jacobi cycle (sm_100a).
//
// Same job as fused_leg.cuh -- one kernel per level and leg,
//   down leg: pre-smoothing sweeps -> residual -> restriction   (multigrid.hpp:268-282)
//   up leg:   coarse-grid correction -> post-smoothing sweeps    (multigrid.hpp:294-301)
// with the operator, f and the input vector read from HBM once -- but for the operators whose
// rows are a 3 x 3 stencil in (line, element) space: every diagonal offset is a*m + delta with
// a, delta in {-1,0,1} for a line length m (level 0 of an n x n grid: m = n, five points;
// level 1: seven points; Galerkin levels >= 2: nine points).
//
// A WARP is the unit of work: its 32 lanes are 32 consecutive elements of a line and it
// streams down a chunk of lines.  The chained stencil stages run one line behind the other in
// program order (stage s on line jj - s + 1 at step jj); every lane keeps the last three
// lines of each stage's result for its own element in registers, and the delta = -1 / +1
// neighbours come from the adjacent lanes by warp shuffles -- no shared memory, no block
// barriers.  Each stage shrinks the correct lanes by one on both sides, so a warp owns the
// 32 - 2H middle lanes (H = stages [+ 1 for the restriction]) and neighbouring warps overlap.
// The operator rows, f and the input of the line PF steps ahead are loaded straight into a
// register ring (coalesced 256-byte requests per warp and array); the ring is indexed with
// compile-time slots by unrolling the line loop over its period.  The kernels use no shared
// memory and ask for the largest L1 carve-out: the halo columns neighbouring warps share are
// served from L1.  (A variant that prefetched deeper through a per-warp cp.async FIFO in shared
// memory measured slower at every level -- more load/store-unit work per row and a smaller L1;
// profiles/r1_stream_legs.md.)
//
// Per-row arithmetic (operation order, no FMA contraction) is that of k_jacobi /
// k_jacobi_zero / k_residual_restrict / k_prolong_add (kernels.cuh): bit-identical results.
#pragma once
#include <cstdint>

#include <cuda_runtime.h>

namespace amgb {
namespace sleg {

enum Kind { DOWN_U = 0, DOWN_ZERO = 1, UP = 2 };

// stencil slot sl = (a + 1) * 3 + (delta + 1), ascending in column order
constexpr unsigned kMask5 = 0x0BAu;   // (-1,0) (0,-1) (0,0) (0,1) (1,0): level 0 of an n x n grid, m = n
constexpr unsigned kMask7a = 0x1BBu;  // level 1 with m = its smaller far offset: (-1,-1) (-1,0) (0,*) (1,0) (1,1)
constexpr unsigned kMask7b = 0x0FEu;  // level 1 with m = its larger far offset:  (-1,0) (-1,1) (0,*) (1,-1) (1,0)
constexpr unsigned kMask9 = 0x1FFu;   // Galerkin levels >= 2

struct Params {
  // The kernel works on a WINDOW of the level: local row k is global row base + k.  A whole level
  // is the window [0, n) with base 0; a rank of the row-block sharded cycle passes its block plus
  // the ghost rows on both sides (whose inputs a halo exchange has filled) and stores results for
  // its own rows only -- the ghost rows are recomputed redundantly, like the lanes at a warp's edge.
  int base;       // global row of local row 0 (may be negative: rows before the level start)
  int n_global;   // rows of the level
  int own_begin;  // local rows [own_begin, own_end) are stored
  int own_end;
  int cbase;      // global coarse index of fc[0] and e[0]
  int n_e;        // entries of e / fc that exist locally
  int nu;         // Jacobi sweeps per smooth call (selects the kernel instantiation on the host)
  int n;         // rows of the window
  int m;         // line length
  int n_lines;   // ceil(n / m)
  int Wu;        // owned elements per warp (32 - 2H)
  int n_strips;  // ceil(m / Wu)
  int LJ;        // lines per chunk
  int n_chunks;
  int n_warps;   // n_strips * n_chunks
  int ld;
  int n_coarse;
  double omega;
  const double* val;
  const double* f;
  const double* uin;
  const double* e;
  double* uout;
  double* fc;
};

__host__ __device__ constexpr int popc9(unsigned v) {
  int c = 0;
  for (int i = 0; i < 9; ++i) c += (v >> i) & 1u;
  return c;
}

template <int KIND, unsigned MASK, int NU, int PF_ = 2>
struct Leg {
  static constexpr int ND = popc9(MASK);
  static constexpr int NS = (KIND == DOWN_U) ? NU + 1 : NU;  // chained stencil stages
  static constexpr int X = (KIND == UP) ? 0 : 1;
  static constexpr int H = NS + X;                           // lanes lost on each side
  static constexpr int PF = PF_;                             // lines in flight
  static constexpr int RS = NS + 1 + PF;                     // register-ring slots
  static constexpr int DC = popc9(MASK & 0xFu);              // rank of the centre slot (0,0)
  static constexpr int S_OUT = (KIND == UP) ? NS : NS - 1;   // stage whose result is the iterate
  static_assert((MASK >> 4) & 1u, "the diagonal must be present");

  struct Line {
    double a[ND];
    double f, u, e0, e1;
  };

  static __device__ __forceinline__ void load(Line& L, const Params& P, int g, int lane) {
    const int k = g + lane;
    const int kg = k + P.base;
    const bool ok = (k >= 0 && k < P.n && kg >= 0 && kg < P.n_global);
    const double* vp = P.val + k;
#pragma unroll
    for (int d = 0; d < ND; ++d) L.a[d] = ok ? __ldg(vp + (size_t)d * P.ld) : 0.0;
    L.f = ok ? __ldg(P.f + k) : 0.0;
    if (KIND != DOWN_ZERO) L.u = ok ? __ldg(P.uin + k) : 0.0;
    if (KIND == UP) {
      // (P e)[k]: odd k: 1 e[J]; even k: .5 e[J-1] + .5 e[J], J = k >> 1, terms outside
      // [0, n_coarse) absent (interpolator.hpp:118-125)
      const int J = kg >> 1, Jl = J - P.cbase;
      L.e0 = (ok && !(kg & 1) && J - 1 >= 0 && J - 1 < P.n_coarse && Jl - 1 >= 0 && Jl - 1 < P.n_e)
                 ? __ldg(P.e + Jl - 1) : 0.0;
      L.e1 = (ok && J < P.n_coarse && Jl >= 0 && Jl < P.n_e) ? __ldg(P.e + Jl) : 0.0;
    }
  }

  // input value of a row (stage 0)
  static __device__ __forceinline__ double input(const Line& L, const Params& P, int k) {
    if (KIND == DOWN_U) return L.u;
    if (KIND == DOWN_ZERO) {
      const double d = L.a[DC];
      return (d == 0.0) ? 0.0 : __dmul_rn(P.omega, __ddiv_rn(L.f, d));
    }
    double acc = 0.0;
    if (k & 1) {
      acc = __dadd_rn(acc, __dmul_rn(1.0, L.e1));
    } else {
      acc = __dadd_rn(acc, __dmul_rn(0.5, L.e0));
      acc = __dadd_rn(acc, __dmul_rn(0.5, L.e1));
    }
    return __dadd_rn(L.u, acc);
  }

  template <int SL>
  static __device__ __forceinline__ void slot(const Line& L, double xm, double x0, double xp, double& acc) {
    if constexpr ((MASK >> SL) & 1u) {
      constexpr int d = popc9(MASK & ((1u << SL) - 1u));
      constexpr int a = SL / 3 - 1, dl = SL % 3 - 1;
      const double xl = (a < 0) ? xm : (a == 0 ? x0 : xp);
      double xv = xl;
      if constexpr (dl < 0) xv = __shfl_up_sync(0xffffffffu, xl, 1);
      if constexpr (dl > 0) xv = __shfl_down_sync(0xffffffffu, xl, 1);
      // an absent entry is stored as 0.0 and every x a valid row can see is finite (invalid rows
      // carry zeros), so acc - 0 * x == acc bit for bit: no test, no select
      acc = __dsub_rn(acc, __dmul_rn(L.a[d], xv));
    }
  }
  // f - sum a x over the row, ascending column order
  static __device__ __forceinline__ double stencil(const Line& L, double xm, double x0, double xp) {
    double acc = L.f;
    slot<0>(L, xm, x0, xp, acc);
    slot<1>(L, xm, x0, xp, acc);
    slot<2>(L, xm, x0, xp, acc);
    slot<3>(L, xm, x0, xp, acc);
    slot<4>(L, xm, x0, xp, acc);
    slot<5>(L, xm, x0, xp, acc);
    slot<6>(L, xm, x0, xp, acc);
    slot<7>(L, xm, x0, xp, acc);
    slot<8>(L, xm, x0, xp, acc);
    return acc;
  }

  struct State {
    Line R[RS];
    double w[NS][3];  // w[s]: results of stage s (0 = input) on its last three lines
  };
  // The three-line windows rotate with the step.  When the ring period RS is a multiple of 3 the
  // slot of a line is a compile-time function of the step's position P_ in the unrolled period
  // (newest P_ % 3) and nothing moves; otherwise the window is shifted.
  static constexpr bool kStaticWindows = (RS % 3 == 0);
  static __device__ __forceinline__ constexpr int WP(int p) { return kStaticWindows ? p % 3 : 2; }        // line j + 1
  static __device__ __forceinline__ constexpr int W0(int p) { return kStaticWindows ? (p + 2) % 3 : 1; }  // line j
  static __device__ __forceinline__ constexpr int WM(int p) { return kStaticWindows ? (p + 1) % 3 : 0; }  // line j - 1
  static __device__ __forceinline__ void push(double (&w)[3], double v, int p) {
    if (kStaticWindows) {
      w[p % 3] = v;
    } else {
      w[0] = w[1];
      w[1] = w[2];
      w[2] = v;
    }
  }

  // One step: loads of line jj + 1 + PF, input stage on line jj + 1 (ring slot P_), stage s on
  // line jj - s + 1 (slot P_ - s), restriction of the residual line.  g1 = first row of line jj + 1.
  template <int P_>
  static __device__ __forceinline__ void step(State& S, const Params& P, int jj, int g1, int lane, int j0, int j1,
                                              int own_lo, int own_hi) {
    const int m = P.m;
    load(S.R[(P_ + PF) % RS], P, g1 + PF * m, lane);
    const bool own_lane = (lane >= own_lo && lane < own_hi);
    // ---- input stage, line jj + 1
    {
      const Line& L = S.R[P_ % RS];
      const int k = g1 + lane;
      const double in = input(L, P, k + P.base);
      push(S.w[0], in, P_);
      if (S_OUT == 0 && own_lane && jj + 1 >= j0 && jj + 1 < j1 && k >= P.own_begin && k < P.own_end) P.uout[k] = in;
    }
#pragma unroll
    for (int s = 1; s <= NS; ++s) {
      const Line& L = S.R[(P_ - s + 2 * RS) % RS];
      const int j = jj - s + 1;
      const int k = g1 - s * m + lane;
      const double acc = stencil(L, S.w[s - 1][WM(P_)], S.w[s - 1][W0(P_)], S.w[s - 1][WP(P_)]);
      const bool own = own_lane && j >= j0 && j < j1 && k >= P.own_begin && k < P.own_end;
      if (KIND != UP && s == NS) {
        // residual -> restriction: f_c[J] = (.5 r[2J] + r[2J+1]) + .5 r[2J+2]   (interpolator.hpp:64-68)
        const double rm = __shfl_up_sync(0xffffffffu, acc, 1);
        const double rp = __shfl_down_sync(0xffffffffu, acc, 1);
        const int kg = k + P.base;
        if (own && (kg & 1)) {
          const int J = (kg - 1) >> 1;
          if (J < P.n_coarse) P.fc[J - P.cbase] = __dadd_rn(__dadd_rn(__dmul_rn(0.5, rm), acc), __dmul_rn(0.5, rp));
        }
      } else {
        const double diag = L.a[DC];
        const double xc = S.w[s - 1][W0(P_)];
        const double out = (diag == 0.0) ? xc : __dadd_rn(xc, __dmul_rn(P.omega, __ddiv_rn(acc, diag)));
        if (s < NS) push(S.w[s], out, P_);
        if (s == S_OUT && own) P.uout[k] = out;
      }
    }
  }

  // RS consecutive steps (one period of the register ring); `left` counts the steps still to
  // do and is the same for every warp of the grid, so control flow stays convergent and the
  // shuffles need no re-convergence code.
  template <int P_>
  static __device__ __forceinline__ void steps(State& S, const Params& P, int& jj, int& g1, int& left, int lane, int j0,
                                               int j1, int own_lo, int own_hi) {
    if constexpr (P_ < RS) {
      if (left <= 0) return;
      step<P_>(S, P, jj, g1, lane, j0, j1, own_lo, own_hi);
      ++jj;
      --left;
      g1 += P.m;
      steps<P_ + 1>(S, P, jj, g1, left, lane, j0, j1, own_lo, own_hi);
    }
  }

  static __device__ __forceinline__ void run(const Params& P) {
    const int lane = threadIdx.x & 31;
    // warps past the end redo the last tile (same values to the same addresses) instead of
    // leaving early: no thread-dependent branch ahead of the shuffles
    int warp = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    warp = warp < P.n_warps ? warp : P.n_warps - 1;
    const int strip = warp % P.n_strips, chunk = warp / P.n_strips;
    const int i0 = strip * P.Wu, i1 = (i0 + P.Wu < P.m) ? i0 + P.Wu : P.m;
    const int j0 = chunk * P.LJ, j1 = (j0 + P.LJ < P.n_lines) ? j0 + P.LJ : P.n_lines;
    const int jA = j0 - NS;
    const int own_lo = H, own_hi = H + (i1 - i0);
    State S;
#pragma unroll
    for (int s = 0; s < NS; ++s) S.w[s][0] = S.w[s][1] = S.w[s][2] = 0.0;
    int g1 = jA * P.m + (i0 - H);  // first row of line jA
#pragma unroll
    for (int q = 0; q < PF; ++q) load(S.R[q], P, g1 + q * P.m, lane);
    int jj = jA - 1;
    int left = P.LJ + 2 * NS;  // steps jA - 1 .. j0 + LJ + NS - 2 (a short last chunk just runs past its end)
    while (left > 0) steps<0>(S, P, jj, g1, left, lane, j0, j1, own_lo, own_hi);
  }
};

template <int KIND, unsigned MASK, int NU, int PF_ = 2>
__global__ void __launch_bounds__(128, ((popc9(MASK) <= 5 && PF_ == 2) ? 4 : 3))
    k_stream_leg(const __grid_constant__ Params P) {
  Leg<KIND, MASK, NU, PF_>::run(P);
}

}  // namespace sleg
}  // namespace amgb

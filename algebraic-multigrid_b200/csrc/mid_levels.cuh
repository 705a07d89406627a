// Multi-level fused legs for the MID levels of the damped-Jacobi V-cycle (sm_100a).
//
// Below the levels that are worth streaming from HBM (stream_leg.cuh) and above the handful of
// tiny levels that run in one block (k_coarse_tail), the levels of the hierarchy hold a few
// thousand to a few ten thousand rows each: every kernel on them costs launch + dependency
// latency, not bandwidth, and the per-operator path launches six of them per level.  Here ONE
// kernel runs the down legs of ALL mid levels (multigrid.hpp:268-282 for each of them) and a
// second one all their up legs (:294-301): a block owns a tile of rows of the first mid level
// and the nested images of that tile on the deeper levels, keeps every level's vectors in shared
// memory and recomputes a halo of rows around its tile so that it never needs a neighbouring
// block's result (stage by stage the valid range shrinks by the level's half-bandwidth; the
// coarser the level the fewer rows the halo costs).  Block barriers stand where the per-operator
// path has launch boundaries; there is no grid-wide synchronisation.
//
// Per-row arithmetic is that of k_jacobi_zero / k_jacobi / k_residual / k_restrict /
// k_prolong_add (kernels.cuh) in the same operation order -- bit-identical to them and to the
// oracle in reference arithmetic; AMGB_ARITH_FAST uses FMAs and a refined reciprocal like the
// streaming legs.
//
// The body is written against an `Env` (phase = run a lambda for every thread, barrier between
// phases) so that tests/cpp/mid_levels_host.cpp runs the very same range logic serially on the
// CPU and checks it against the oracle.
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define AMGB_MID_FN __host__ __device__ __forceinline__
#else
#define AMGB_MID_FN inline
#endif

namespace amgb {
namespace mid {

constexpr int kMaxLevels = 8;
constexpr int kMaxDiag = 16;

struct Level {
  int n;         // rows
  int n_coarse;  // rows of the next level
  int nd;        // diagonals, offsets ascending
  int off[kMaxDiag];
  int diag_d;    // index of the main diagonal
  int w;         // half-bandwidth
  int ld;
  const double* val;  // rows of A in DIA: val[d * ld + row]
  double* f;          // right-hand side (read on the first mid level, written on the deeper ones)
  double* u;          // iterate at the end of the cycle
  double* tmp;        // iterate after pre-smoothing (down kernel writes, up kernel reads)
  // shared-memory layout (doubles): three buffers of `len` entries for the down kernel, three for the up kernel
  int len_down, off_down;
  int len_up, off_up;
};

struct Params {
  int n_lv;             // mid levels
  int nu;               // Jacobi sweeps per smooth call
  int first_is_level0;  // the first mid level is the finest level: pre-smoothing starts from its u
  int T;                // own rows per block on the first mid level (multiple of 2^n_lv)
  int n_blocks;
  double omega;
  Level lv[kMaxLevels];
  double* f_next;        // right-hand side of the level below the last mid level
  const double* u_next;  // its solution (read by the up kernel)
  int n_next;
  // shared memory (doubles): per-level vector buffers first, then ONE coefficient buffer (the DIA
  // rows of the level being processed, val[d * len + j]) that every level reuses
  int coef_off_down, coef_off_up;
  int smem_doubles_down, smem_doubles_up;
};

struct Range {  // inclusive; empty when hi < lo
  int lo, hi;
};
AMGB_MID_FN Range clip(Range r, int n) {
  if (r.lo < 0) r.lo = 0;
  if (r.hi > n - 1) r.hi = n - 1;
  return r;
}
AMGB_MID_FN Range grow(Range r, int by, int n) { return clip(Range{r.lo - by, r.hi + by}, n); }
AMGB_MID_FN Range hull(Range a, Range b) {
  if (a.hi < a.lo) return b;
  if (b.hi < b.lo) return a;
  return Range{a.lo < b.lo ? a.lo : b.lo, a.hi > b.hi ? a.hi : b.hi};
}
AMGB_MID_FN bool inside(Range r, int k) { return k >= r.lo && k <= r.hi; }

// stencil stages of a level's down leg, the residual included
AMGB_MID_FN int down_stages(const Params& P, int i) { return (i == 0 && P.first_is_level0) ? P.nu + 1 : P.nu; }

struct Plan {
  Range own[kMaxLevels + 1];  // rows a block stores on each mid level (+ the level below)
  Range res[kMaxLevels];      // down: rows whose residual is formed
  Range in0[kMaxLevels];      // down: rows of the input stage (= rows whose f is needed)
  Range fin[kMaxLevels];      // up: rows whose final iterate is needed (own rows + what the finer level interpolates from)
  Range inp[kMaxLevels];      // up: rows of the input stage (tmp + P e)
};

AMGB_MID_FN void make_plan(const Params& P, int b, Plan& Q) {
  // nested own ranges: coarse row J belongs to the owner of fine row 2J + 1
  int lo = b * P.T, hi = lo + P.T;  // [lo, hi)
  for (int i = 0; i <= P.n_lv; ++i) {
    const int n = (i < P.n_lv) ? P.lv[i].n : P.n_next;
    const bool last = (b == P.n_blocks - 1);
    Q.own[i] = Range{lo < n ? lo : n, (last || hi > n ? n : hi) - 1};
    lo >>= 1;
    hi >>= 1;
  }
  // down pass, from the deepest level up: what the level below needs decides what this one computes
  Range need_f = Q.own[P.n_lv];  // rows of the next level's f this block must produce
  for (int i = P.n_lv - 1; i >= 0; --i) {
    const Level& V = P.lv[i];
    Range r = clip(Range{2 * need_f.lo, 2 * need_f.hi + 2}, V.n);
    if (need_f.hi < need_f.lo) r = Range{0, -1};
    r = hull(r, Q.own[i]);
    Q.res[i] = r;
    Q.in0[i] = (r.hi < r.lo) ? r : grow(r, down_stages(P, i) * V.w, V.n);
    need_f = Q.in0[i];
  }
  // up pass, from the first level down: what the finer level interpolates from decides the final range
  Range need_e{0, -1};
  for (int i = 0; i < P.n_lv; ++i) {
    const Level& V = P.lv[i];
    Q.fin[i] = clip(hull(Q.own[i], need_e), V.n);
    Q.inp[i] = (Q.fin[i].hi < Q.fin[i].lo) ? Q.fin[i] : grow(Q.fin[i], P.nu * V.w, V.n);
    // fine row k reads coarse entries (k >> 1) - 1 (even k) and k >> 1
    need_e = (Q.inp[i].hi < Q.inp[i].lo) ? Q.inp[i] : clip(Range{(Q.inp[i].lo >> 1) - 1, Q.inp[i].hi >> 1}, V.n_coarse);
  }
}

// ---- arithmetic: reference order (bit-identical to kernels.cuh / the oracle) or fast ----
template <bool FAST>
struct Arith {
  static AMGB_MID_FN double mul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
  }
  static AMGB_MID_FN double add(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
  }
  static AMGB_MID_FN double sub(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
  }
  static AMGB_MID_FN double div(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
  }
  static AMGB_MID_FN double rcp(double d) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = __fma_rn(-d, r, 1.0);
    r = __fma_rn(r, e, r);
    e = __fma_rn(-d, r, 1.0);
    return __fma_rn(r, e, r);
#else
    return 1.0 / d;
#endif
  }
  static AMGB_MID_FN double mulsub(double acc, double a, double x) {
#if defined(__CUDA_ARCH__)
    if (FAST) return __fma_rn(-a, x, acc);
#endif
    return sub(acc, mul(a, x));
  }
  static AMGB_MID_FN double fma_(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return std::fma(a, b, c);
#endif
  }
  // the FAST formulas are those of stream_leg.cuh, so a level gives the same bits whichever kernel runs it
  static AMGB_MID_FN double relax(double x, double r, double d, double omega) {
    if (d == 0.0) return x;
    if (FAST) return fma_(omega * rcp(d), r, x);
    return add(x, mul(omega, div(r, d)));
  }
  // (.5 r0 + r1) + .5 r2 with absent terms passed as 0 (interpolator.hpp:64-68)
  static AMGB_MID_FN double restrict3(double r0, bool h0, double r1, bool h1, double r2, bool h2) {
    if (FAST) return fma_(0.5, (h0 ? r0 : 0.0) + (h2 ? r2 : 0.0), h1 ? r1 : 0.0);
    double acc = 0.0;
    if (h0) acc = add(acc, mul(0.5, r0));
    if (h1) acc = add(acc, mul(1.0, r1));
    if (h2) acc = add(acc, mul(0.5, r2));
    return acc;
  }
  // u + (P e)_k: odd k: e1; even k: .5 e0 + .5 e1, absent terms passed as 0 (interpolator.hpp:52-56)
  static AMGB_MID_FN double prolong_add(double u, bool odd, double e0, bool h0, double e1, bool h1) {
    if (FAST) return odd ? u + (h1 ? e1 : 0.0) : fma_(0.5, (h0 ? e0 : 0.0) + (h1 ? e1 : 0.0), u);
    double acc = 0.0;
    if (odd) {
      if (h1) acc = add(acc, mul(1.0, e1));
    } else {
      if (h0) acc = add(acc, mul(0.5, e0));
      if (h1) acc = add(acc, mul(0.5, e1));
    }
    return add(u, acc);
  }
  static AMGB_MID_FN double relax_zero(double f, double d, double omega) {
    if (d == 0.0) return 0.0;
    if (FAST) return f * (omega * rcp(d));
    return mul(omega, div(f, d));
  }
};

AMGB_MID_FN double ld_global(const double* p) {
#if defined(__CUDA_ARCH__)
  return __ldcg(p);  // vectors other kernels of the cycle wrote: L2, never the non-coherent path
#else
  return *p;
#endif
}
AMGB_MID_FN double ld_const(const double* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}

// The level's operator rows [base, base + cnt) into shared memory, Cf[d * len + j] = A(base + j, diagonal d),
// plus up to two vectors over the same rows: ONE round of bulk copies (TMA, cp.async.bulk + mbarrier on
// the device; memcpy in the host Env) instead of a latency chain per entry.  base is even and cnt a
// multiple of two, so every copy is 16-byte aligned (the DIA rows and the level vectors are padded).
template <class Env>
AMGB_MID_FN void load_level(const Level& V, int base, int cnt, int len, double* Cf, double* v0, const double* g0,
                            double* v1, const double* g1, Env& env) {
  env.bulk_start((V.nd + (g0 ? 1 : 0) + (g1 ? 1 : 0)) * cnt);
  for (int d = 0; d < V.nd; ++d) env.bulk_copy(Cf + (size_t)d * len, V.val + (size_t)d * V.ld + base, cnt);
  if (g0) env.bulk_copy(v0, g0 + base, cnt);
  if (g1) env.bulk_copy(v1, g1 + base, cnt);
  env.bulk_wait();
}
// f_k - sum_d a_d x[k + off_d], ascending column order; Cf and x are shared-memory buffers whose entry 0 is row `base`
template <bool FAST>
AMGB_MID_FN double stencil(const Level& V, int k, double fk, const double* Cf, int len, const double* x, int base) {
  double acc = fk;
  const int j = k - base;
  for (int d = 0; d < V.nd; ++d) {
    const double a = Cf[(size_t)d * len + j];
    if (a != 0.0) acc = Arith<FAST>::mulsub(acc, a, x[j + V.off[d]]);
  }
  return acc;
}

// ---- the down legs of all mid levels for block b
template <bool FAST, class Env>
AMGB_MID_FN void run_down(const Params& P, int b, Env& env) {
  Plan Q;
  make_plan(P, b, Q);
  double* sm = env.smem();
  for (int i = 0; i < P.n_lv; ++i) {
    const Level& V = P.lv[i];
    const Range in0 = Q.in0[i], res = Q.res[i], own = Q.own[i];
    if (in0.hi < in0.lo) continue;  // uniform over the block
    double* A = sm + V.off_down;
    double* B = A + V.len_down;
    double* F = B + V.len_down;
    double* Cf = sm + P.coef_off_down;
    const int len = V.len_down;
    const int base = in0.lo & ~1;
    const int cnt = (in0.hi - base + 2) & ~1;
    const bool from_u = (i == 0 && P.first_is_level0);
    const int NS = down_stages(P, i);
    // operator rows of this block's range into shared memory; the right-hand side of the first mid level
    // (and the stored iterate of the finest level) come from global memory, deeper right-hand sides were
    // restricted into F by the level above
    load_level(V, base, cnt, len, Cf, i == 0 ? F : nullptr, i == 0 ? V.f : nullptr, from_u ? A : nullptr,
               from_u ? V.u : nullptr, env);
    // input stage: the stored iterate (finest level) or the first sweep from the zero guess
    env.phase([&](int t, int nt) {
      for (int k = in0.lo + t; k <= in0.hi; k += nt) {
        double v;
        if (from_u) v = A[k - base];
        else v = Arith<FAST>::relax_zero(F[k - base], Cf[(size_t)V.diag_d * len + (k - base)], P.omega);
        A[k - base] = v;
        if (NS == 1 && inside(own, k)) V.tmp[k] = v;  // one sweep: this already is the smoothed iterate
      }
    });
    env.sync();
    double* src = A;
    double* dst = B;
    for (int s = 1; s <= NS; ++s) {
      const Range r = grow(res, (NS - s) * V.w, V.n);
      env.phase([&](int t, int nt) {
        for (int k = r.lo + t; k <= r.hi; k += nt) {
          const double acc = stencil<FAST>(V, k, F[k - base], Cf, len, src, base);
          if (s == NS) {
            dst[k - base] = acc;  // residual
          } else {
            const double v = Arith<FAST>::relax(src[k - base], acc, Cf[(size_t)V.diag_d * len + (k - base)], P.omega);
            dst[k - base] = v;
            if (s == NS - 1 && inside(own, k)) V.tmp[k] = v;
          }
        }
      });
      env.sync();
      double* t2 = src;
      src = dst;
      dst = t2;
    }
    // restriction of the residual (interpolator.hpp:64-68): into the next level's F and, for the rows
    // this block owns, into that level's right-hand side in global memory
    const bool deeper = (i + 1 < P.n_lv);
    const Range cr = deeper ? Q.in0[i + 1] : Q.own[i + 1];
    double* Fn = deeper ? sm + P.lv[i + 1].off_down + 2 * P.lv[i + 1].len_down : nullptr;
    double* fg = deeper ? P.lv[i + 1].f : P.f_next;
    const Range cown = Q.own[i + 1];
    env.phase([&](int t, int nt) {
      for (int J = cr.lo + t; J <= cr.hi; J += nt) {
        const int k = 2 * J;
        const bool h0 = k < V.n, h1 = k + 1 < V.n, h2 = k + 2 < V.n;
        const double acc = Arith<FAST>::restrict3(h0 ? src[k - base] : 0.0, h0, h1 ? src[k + 1 - base] : 0.0, h1,
                                                  h2 ? src[k + 2 - base] : 0.0, h2);
        if (Fn) Fn[J - (cr.lo & ~1)] = acc;
        if (inside(cown, J)) fg[J] = acc;
      }
    });
    env.sync();
  }
}

// ---- the up legs of all mid levels for block b
template <bool FAST, class Env>
AMGB_MID_FN void run_up(const Params& P, int b, Env& env) {
  Plan Q;
  make_plan(P, b, Q);
  double* sm = env.smem();
  const double* e = P.u_next;  // coarse correction of the level being processed: global at first, then shared
  int e_base = 0;
  bool e_global = true;
  for (int i = P.n_lv - 1; i >= 0; --i) {
    const Level& V = P.lv[i];
    const Range inp = Q.inp[i], fin = Q.fin[i], own = Q.own[i];
    if (inp.hi < inp.lo) continue;
    double* A = sm + V.off_up;
    double* B = A + V.len_up;
    double* F = B + V.len_up;
    double* Cf = sm + P.coef_off_up;
    const int len = V.len_up;
    const int base = inp.lo & ~1;
    const int cnt = (inp.hi - base + 2) & ~1;
    // operator rows, pre-smoothed iterate and right-hand side of this block's range into shared memory
    load_level(V, base, cnt, len, Cf, A, V.tmp, F, V.f, env);
    // input: tmp + P e (interpolator.hpp:52-56, multigrid.hpp:294-296)
    env.phase([&](int t, int nt) {
      for (int k = inp.lo + t; k <= inp.hi; k += nt) {
        const int J = k >> 1;
        auto ev = [&](int j) { return e_global ? ld_global(e + j) : e[j - e_base]; };
        const bool odd = k & 1;
        const bool h1 = J < V.n_coarse, h0 = !odd && J - 1 >= 0 && J - 1 < V.n_coarse;
        A[k - base] = Arith<FAST>::prolong_add(A[k - base], odd, h0 ? ev(J - 1) : 0.0, h0, h1 ? ev(J) : 0.0, h1);
      }
    });
    env.sync();
    double* src = A;
    double* dst = B;
    for (int s = 1; s <= P.nu; ++s) {
      const Range r = grow(fin, (P.nu - s) * V.w, V.n);
      env.phase([&](int t, int nt) {
        for (int k = r.lo + t; k <= r.hi; k += nt) {
          const double acc = stencil<FAST>(V, k, F[k - base], Cf, len, src, base);
          const double v = Arith<FAST>::relax(src[k - base], acc, Cf[(size_t)V.diag_d * len + (k - base)], P.omega);
          dst[k - base] = v;
          if (s == P.nu && inside(own, k)) V.u[k] = v;
        }
      });
      env.sync();
      double* t2 = src;
      src = dst;
      dst = t2;
    }
    e = src;
    e_base = base;
    e_global = false;
  }
}

// ---- shared-memory layout and tile size (host).  Returns false when no tile size fits `cap_doubles`.
inline bool plan_layout(Params& P, int cap_doubles, int force_T = 0) {
  const int gran = 1 << P.n_lv;
  for (int T = force_T > 0 ? force_T : 1024; T >= (force_T > 0 ? force_T : 64); T >>= 1) {
    if (T % gran) continue;
    P.T = T;
    P.n_blocks = (P.lv[0].n + T - 1) / T;
    int len_d[kMaxLevels] = {0}, len_u[kMaxLevels] = {0};
    for (int b = 0; b < P.n_blocks; ++b) {
      Plan Q;
      make_plan(P, b, Q);
      for (int i = 0; i < P.n_lv; ++i) {
        const int ld = Q.in0[i].hi - Q.in0[i].lo + 1, lu = Q.inp[i].hi - Q.inp[i].lo + 1;
        if (ld > len_d[i]) len_d[i] = ld;
        if (lu > len_u[i]) len_u[i] = lu;
      }
    }
    int od = 0, ou = 0;
    for (int i = 0; i < P.n_lv; ++i) {
      P.lv[i].len_down = (len_d[i] + 3) & ~1;  // base is rounded down to even, the count up
      P.lv[i].off_down = od;
      od += 3 * P.lv[i].len_down;
      P.lv[i].len_up = (len_u[i] + 3) & ~1;
      P.lv[i].off_up = ou;
      ou += 3 * P.lv[i].len_up;
    }
    int cd = 0, cu = 0;  // one coefficient buffer, sized for the level that needs the most
    for (int i = 0; i < P.n_lv; ++i) {
      if (P.lv[i].nd * P.lv[i].len_down > cd) cd = P.lv[i].nd * P.lv[i].len_down;
      if (P.lv[i].nd * P.lv[i].len_up > cu) cu = P.lv[i].nd * P.lv[i].len_up;
    }
    P.coef_off_down = od;
    P.coef_off_up = ou;
    P.smem_doubles_down = od + cd;
    P.smem_doubles_up = ou + cu;
    if (P.smem_doubles_down <= cap_doubles && P.smem_doubles_up <= cap_doubles) return true;
  }
  return false;
}

#if defined(__CUDACC__)
// thread 0 issues the bulk copies of a round; every thread waits on the block's mbarrier
struct DeviceEnv {
  double* sm;
  uint64_t* bar;
  uint32_t parity;
  __device__ __forceinline__ double* smem() { return sm; }
  template <class F>
  __device__ __forceinline__ void phase(F&& f) {
    f((int)threadIdx.x, (int)blockDim.x);
  }
  __device__ __forceinline__ void sync() { __syncthreads(); }
  __device__ __forceinline__ void bulk_start(int doubles) {
    if (threadIdx.x == 0) {
      // the buffers were last read before the block barrier that ended the previous phase
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      dev::mbar_expect_tx(bar, (uint32_t)doubles * 8u);
    }
  }
  __device__ __forceinline__ void bulk_copy(double* dst, const double* src, int doubles) {
    if (threadIdx.x == 0) dev::tma_load_1d(dst, src, (uint32_t)doubles * 8u, bar);
  }
  __device__ __forceinline__ void bulk_wait() {
    dev::mbar_wait(bar, parity);
    parity ^= 1u;
  }
};
template <bool FAST, bool UP>
__global__ void __launch_bounds__(1024) k_mid(const __grid_constant__ Params P) {
  extern __shared__ __align__(16) double mid_smem[];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    dev::mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  DeviceEnv env{mid_smem, &bar, 0u};
  if (UP) run_up<FAST>(P, (int)blockIdx.x, env);
  else run_down<FAST>(P, (int)blockIdx.x, env);
}
#endif

}  // namespace mid
}  // namespace amgb

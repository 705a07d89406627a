/*
 * amgb.h -- C ABI of the B200-native multigrid V-cycle path (libamgb.so).
 *
 * This is the drop-in boundary for the hot path of jfdev001/algebraic-multigrid
 * (reference paths below are relative to the reference repository root).
 * Host code (the C++ headers under include/amg/, or any FFI) hands the raw
 * CSC arrays of an Eigen::SparseMatrix<double> (outerIndexPtr / innerIndexPtr /
 * valuePtr of a compressed ColMajor matrix, int32 indices) and plain double
 * vectors to these entry points.  Every pointer argument is a HOST pointer
 * unless its name ends in _dev.  No C++ / torch types cross this boundary.
 *
 * Conventions
 *   - every function returns 0 on success or an AMGB_E* code; the message is
 *     available from amgb_last_error() (thread-local).
 *   - handles are opaque and own their device memory; host arrays are only
 *     borrowed for the duration of the call.
 *   - there is NO CPU fallback: without a CUDA device every compute entry
 *     point fails with AMGB_ECUDA.
 *   - the matrix mirror reads CSC column c as "row c" exactly like the
 *     reference's Gauss-Seidel (include/amg/smoother.hpp:101-117).  For
 *     operations defined on the rows of A (residual, Jacobi, multicolour GS)
 *     the mirror is shared when A is bitwise symmetric and a transposed mirror
 *     is built otherwise, so results never depend on the symmetry assumption.
 */
#ifndef AMGB_H_
#define AMGB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AMGB_OK 0
#define AMGB_EINVAL 1   /* invalid argument (maps to std::invalid_argument) */
#define AMGB_ECUDA 2    /* CUDA runtime / driver failure, or no device       */
#define AMGB_ENCCL 3    /* NCCL failure                                       */
#define AMGB_ESTATE 4   /* call not valid in the handle's current state       */

/* Smoother kinds.  GS is the reference's smoother; the other two are the
 * partitionable smoothers the north star adds (no reference counterpart). */
#define AMGB_SMOOTHER_GS 0        /* include/amg/smoother.hpp:86-216 */
#define AMGB_SMOOTHER_JACOBI 1    /* u <- u + omega D^-1 (f - A u)            */
#define AMGB_SMOOTHER_COLOR_GS 2  /* greedy multicolour (red-black on level 0) */

/* How amgb_smooth_gs orders its updates.  Every mode visits the rows in the reference's order
 * (include/amg/smoother.hpp:148-174); they differ in the kernel and in whether its arithmetic is
 * the reference's bit for bit.
 *   AUTO        grid-structured operators (every entry couples neighbours of an n_lines x m grid and
 *               none crosses the end of a line: level 0) run the multi-SM wavefront kernel, bit-exact;
 *               other banded operators (the Galerkin levels, whose 1-D interpolation couples across
 *               line ends) the line-scan kernel; anything else the level-scheduled kernel.
 *   LEVELSCHED  generic level-scheduled kernel, one SM, bit-exact.
 *   LINESCAN    single-SM affine-scan kernel (re-associates the distance-1 chain: <= 1e-14 relative),
 *               level-scheduled kernel where it does not apply. */
#define AMGB_GS_AUTO 0
#define AMGB_GS_LEVELSCHED 1
#define AMGB_GS_LINESCAN 2
/* kernels behind the modes (amgb_matrix_gs_kernel) */
#define AMGB_GS_KERNEL_FRONTS 0
#define AMGB_GS_KERNEL_LINESCAN 1
#define AMGB_GS_KERNEL_WAVE 2

typedef struct amgb_matrix amgb_matrix;       /* device mirror of one CSC matrix */
typedef struct amgb_hierarchy amgb_hierarchy; /* device mirror of AMG::Multigrid  */

const char* amgb_last_error(void);
int amgb_version(void);
/* number of visible CUDA devices (0 when there is none; never fails) */
int amgb_device_count(void);
/* selects the device this thread's subsequent calls use (cudaSetDevice) */
int amgb_set_device(int device);

/* ------------------------------------------------------------------------
 * Input generators -- host only, no device needed.
 * Replace AMG::Grid<double> (include/amg/grid.hpp:19-141).
 * ---------------------------------------------------------------------- */
/* grid.hpp:31 */
double amgb_grid_spacing_h(int64_t n);
/* grid.hpp:39-41 */
int64_t amgb_points_n_from_grid_spacing_h(double h);
/* nnz of the n*n x n*n five-point operator: 5n^2 - 4n */
int64_t amgb_grid_laplacian_nnz(int64_t n);
/* grid.hpp:88-98 with eps_y == 1.0; eps_y scales the kron(D,I) (+-n) coupling.
 * colptr has n*n+1 entries, rowidx/val amgb_grid_laplacian_nnz(n). */
int amgb_grid_laplacian(int64_t n, double eps_y, int* colptr, int* rowidx, double* val);
/* grid.hpp:108-140 with the default f(x,y) = 5 exp(-10 (x^2+y^2)); b has n*n entries */
int amgb_grid_rhs(int64_t n, double* b);

/* ------------------------------------------------------------------------
 * Setup helpers -- host only.  Replace LinearInterpolator::make_operators
 * (include/amg/interpolator.hpp:106-141) and Multigrid's level-size rule
 * (include/amg/multigrid.hpp:127-130).  Integer outputs are bit-exact.
 * ---------------------------------------------------------------------- */
int64_t amgb_n_H_dofs_from_n_h_dofs(int64_t n_h_dofs);
/* number of stored entries of P for (n_h, n_H) */
int64_t amgb_interp_nnz(int64_t n_h, int64_t n_H);
/* P is n_h x n_H (colptr: n_H+1), R = P^T is n_H x n_h (colptr: n_h+1) */
int amgb_interp_make_operators(int64_t n_h, int64_t n_H,
                               int* P_colptr, int* P_rowidx, double* P_val,
                               int* R_colptr, int* R_rowidx, double* R_val);

/* Matrix-free grid transfers of LinearInterpolator on host vectors:
 * out = R r (interpolator.hpp:64-68) and out = P e (interpolator.hpp:52-56). */
int amgb_linear_restrict(int64_t n_h, int64_t n_H, const double* r, double* out);
int amgb_linear_prolong(int64_t n_h, int64_t n_H, const double* e, double* out);
/* y = A x for ANY compressed CSC matrix (n_rows x n_cols, rectangular allowed), in Eigen's
 * evaluation order: what InterpolatorBase::restriction / prolongation compute with the stored
 * R / P (interpolator.hpp:52-68) when they are not LinearInterpolator's operators. */
int amgb_csc_spmv(int n_rows, int n_cols, const int* colptr, const int* rowidx, const double* val,
                  const double* x, double* y);

/* ------------------------------------------------------------------------
 * Matrix mirror + stand-alone operators (the SmootherBase::smooth boundary,
 * include/amg/smoother.hpp:63-65).  u is updated in place.
 * ---------------------------------------------------------------------- */
int amgb_matrix_create(int n_rows, int n_cols, const int* colptr, const int* rowidx,
                       const double* val, amgb_matrix** out);
int amgb_matrix_destroy(amgb_matrix* A);
/* stored entries after dropping explicit zeros (what the kernels stream) */
int64_t amgb_matrix_nnz_device(const amgb_matrix* A);
int amgb_matrix_is_symmetric(const amgb_matrix* A);

/* SparseGaussSeidel::smooth (smoother.hpp:189-215): n_iters x (forward then
 * backward lexicographic sweep); if every != 0 the residual sum of squares is
 * evaluated every `every` iterations and the loop stops once it is <= tolerance.
 * iters_done / final_error may be NULL (final_error is 100 if never evaluated). */
int amgb_smooth_gs(amgb_matrix* A, double* u, const double* b, double tolerance,
                   int64_t every, int64_t n_iters, int mode,
                   int64_t* iters_done, double* final_error);
/* n_sweeps damped-Jacobi sweeps */
int amgb_smooth_jacobi(amgb_matrix* A, double* u, const double* b, double omega,
                       int64_t n_sweeps);
/* n_iters x (colours 0..C-1 then C-1..0) */
int amgb_smooth_color_gs(amgb_matrix* A, double* u, const double* b, int64_t n_iters);
/* greedy colouring used by amgb_smooth_color_gs; color has n entries */
int amgb_matrix_coloring(amgb_matrix* A, int* n_colors, int* color);
/* r = f - A u, in Eigen's order (include/amg/multigrid.hpp:272-274) */
int amgb_residual(amgb_matrix* A, const double* u, const double* f, double* r);
/* sum_i (b_i - (A u)_i)^2 (include/amg/common.hpp:17-27) */
int amgb_rss(amgb_matrix* A, const double* u, const double* b, double* out);

/* Timing hook of the smoother + residual microbenchmark: mean milliseconds
 * per launch (CUDA events on the handle's stream) of one pass over the matrix on the vectors the
 * last amgb_residual / amgb_rss call uploaded.  kind 0: one damped-Jacobi sweep, 1: one
 * colour-complete multicolour Gauss-Seidel sweep, 2: one residual (3, 4: below).  amgb_matrix_stream_bytes: the
 * matrix bytes such a pass streams (kind 1 after the colouring exists). */
int amgb_matrix_time(amgb_matrix* A, int kind, double omega, int warmup, int reps, double* ms_per_launch);
int64_t amgb_matrix_stream_bytes(amgb_matrix* A, int kind);
/* kind 3 / 4 of amgb_matrix_time: one FORWARD lexicographic Gauss-Seidel sweep (in place on the uploaded
 * u) with AMGB_GS_AUTO / AMGB_GS_LINESCAN.  amgb_matrix_gs_kernel: the AMGB_GS_KERNEL_* that `mode`
 * selects for this matrix (builds the kernel's schedule if it does not exist yet); -1 on error. */
int amgb_matrix_gs_kernel(amgb_matrix* A, int mode);
/* Diagnostic: the wavefront kernel evaluates (b_i - sigma) / a_ii as the split form of the compiler's
 * IEEE division (reciprocal refinement at pack time, three dependent operations per row, the
 * compiler's own range guard and fallback).  This runs both forms on n_pairs pseudo-random operand
 * pairs on the current device and counts results whose bits differ (must be 0). */
int amgb_selftest_division(int64_t n_pairs, uint64_t seed, int64_t* mismatches);

/* ------------------------------------------------------------------------
 * Hierarchy = AMG::Multigrid<double> (include/amg/multigrid.hpp:22-365).
 * ---------------------------------------------------------------------- */
typedef struct amgb_options {
  int n_levels;             /* multigrid.hpp:155                                   */
  double tolerance;         /* :155  stop when sum r^2 <= tolerance                */
  int64_t compute_error_every_n_iters; /* :155                                     */
  int64_t n_iters;          /* :156  V-cycle cap                                   */
  int smoother;             /* AMGB_SMOOTHER_*                                     */
  int64_t smoother_iters;   /* SmootherBase::n_iters: GS / colour-GS symmetric
                               sweeps, or Jacobi sweeps, per smooth() call         */
  double omega;             /* Jacobi damping                                      */
  int gs_mode;              /* AMGB_GS_*                                           */
  int use_graph;            /* 1: replay the V-cycle as a CUDA graph               */
  int skip_dead_coarse_smooth; /* 1: drop the pre-smooth + residual the reference
                               does on the coarsest level and then overwrites
                               (multigrid.hpp:265-274 vs :287-288); unobservable   */
  int fuse;                 /* bit 0 (default on): damped-Jacobi cycles skip the operator read
                               of the first pre-smoothing sweep on coarse levels (u = 0
                               there); bit 1 (default off: measured slower, profiles/):
                               fold u += P e into the first post-smoothing sweep; bit 2
                               (default on): fused legs -- per level ONE kernel for
                               sweeps + residual + restriction (multigrid.hpp:268-282) and
                               ONE for prolongation + add + sweeps (:294-301), each reading
                               the operator from HBM once (register-streaming kernels for
                               operators that are 3 x 3 stencils over lines, one or two sweeps);
                               bit 3 (default off): TMA-ring fused legs for the other banded
                               levels; bit 4 (default on): the small coarse levels, the
                               coarsest solve included, run in ONE kernel launch; bit 5
                               (default on): the mid levels between the streamed ones and
                               that tail run all their down legs in one launch and all
                               their up legs in another (shared-memory tiles with
                               recomputed halos, mid_levels.cuh); bit 6 (opt-in, or
                               AMGB_COMPRESS=1): a level whose operator is VERIFIED at setup (bit for bit, on
                               the device) to be a constant five-point stencil runs
                               matrix-free legs: the coefficients travel as kernel
                               parameters and no operator row is read (28 instead of 68
                               bytes per row); bit 7 (opt-in, or
                               AMGB_COMPRESS=1): a level whose operator has at most 256
                               DISTINCT rows (the
                               Galerkin levels of a constant stencil have a few dozen) runs
                               dictionary legs: one byte per row selects a row of a table
                               held in shared memory, every row verified against it at
                               setup (1 instead of 8 x diagonals operator bytes per row).
                               The arithmetic, hence every bit of the result, is unchanged */
  int arith;                /* arithmetic of the damped-Jacobi cycle's kernels.
                               AMGB_ARITH_REFERENCE (default): the oracle's operation order,
                               separate multiply / subtract, IEEE division -- bit-identical
                               to the CPU oracle.  AMGB_ARITH_FAST: fused multiply-add and
                               u + (omega / d) * r with a Newton-refined reciprocal; results
                               agree with the oracle to ~1e-15 relative per sweep (the
                               north star's contract is 1e-12), iteration counts identical */
} amgb_options;
#define AMGB_ARITH_REFERENCE 0
#define AMGB_ARITH_FAST 1
void amgb_options_default(amgb_options* opt);

/* Multigrid constructor (multigrid.hpp:151-244).  Validation order and the two
 * AMGB_EINVAL conditions are the reference's (:165-178). */
int amgb_hierarchy_create(int n_rows, int n_cols, const int* colptr, const int* rowidx,
                          const double* val, const double* b, int64_t b_rows,
                          const amgb_options* opt, amgb_hierarchy** out);
int amgb_hierarchy_destroy(amgb_hierarchy* h);

/* ------------------------------------------------------------------------
 * Row-block sharded hierarchy over the GPUs of one NVSwitch node: one process per
 * GPU, one communicator per process (no reference counterpart -- the reference is
 * single-process).  Rank 0 creates the 128-byte id and hands
 * it to the other ranks by any means (bench.py: torch.distributed broadcast).
 * Levels with at least min_rows_per_rank rows per rank are split into contiguous row
 * blocks with NCCL send/recv halo exchange with ranks +-1 before every Jacobi sweep /
 * residual / prolongation; coarser levels are agglomerated: their right-hand side is
 * gathered once per cycle and every rank keeps a replica (instead of GPU 0 solving
 * and scattering back, which would add a second collective).  Every rank passes the
 * same full A and b; vectors are given / returned at full length (or per row block, the *_local
 * entry points).  Damped Jacobi (fused legs that push their halos themselves) and multicolour
 * Gauss-Seidel (one halo exchange per colour pass) are the partitioned smoothers; lexicographic
 * Gauss-Seidel is a single-GPU path.
 * ---------------------------------------------------------------------- */
typedef struct amgb_comm amgb_comm;
int amgb_comm_unique_id_bytes(void);
int amgb_comm_get_unique_id(void* id);
int amgb_comm_create(const void* id, int rank, int world, amgb_comm** out);
int amgb_comm_destroy(amgb_comm* comm);
int amgb_hierarchy_create_sharded(amgb_comm* comm, int64_t min_rows_per_rank, int n_rows, int n_cols,
                                  const int* colptr, const int* rowidx, const double* val,
                                  const double* b, int64_t b_rows, const amgb_options* opt,
                                  amgb_hierarchy** out);
int amgb_hierarchy_n_sharded_levels(const amgb_hierarchy* h);
/* rows [begin, end) of `level` this rank owns (the whole level when it is not sharded) */
int amgb_hierarchy_local_range(const amgb_hierarchy* h, int level, int64_t* begin, int64_t* end);
int64_t amgb_hierarchy_halo_exchanges_per_vcycle(const amgb_hierarchy* h);
/* How halos travel.  PEER: one kernel per exchange writes the boundary rows straight into
 * the neighbours' halo buffers (CUDA-IPC mapped peer memory over NVLink) and spins on an
 * epoch flag; NCCL: grouped ncclSend/ncclRecv (fallback when IPC mapping is unavailable, or
 * forced with the environment variable AMGB_HALO=nccl). */
#define AMGB_HALO_NONE 0
#define AMGB_HALO_NCCL 1
#define AMGB_HALO_PEER 2
int amgb_hierarchy_halo_mode(const amgb_hierarchy* h);
/* 1 if a peer-memory wait ever gave up on a neighbour (results are invalid).  amgb_vcycles,
 * amgb_solve*, amgb_hierarchy_rss, the getters and amgb_synchronize check the same flag after
 * their stream synchronise and fail with AMGB_ENCCL.  Timeout: ~4 s, AMGB_HALO_TIMEOUT_MS. */
int amgb_hierarchy_halo_timed_out(amgb_hierarchy* h);
/* host only: the plan a sharded hierarchy uses.  starts has n_levels*(world+1) slots
 * (row l holds the world+1 block boundaries of level l), the other arrays n_levels. */
int amgb_partition_plan(int n_levels, const int64_t* level_sizes, const int* half_bandwidth,
                        int world, int64_t min_rows_per_rank, int* n_sharded, int64_t* starts,
                        int* halo_lo, int* halo_hi, int* ghost);

/* make the handle launch on a caller-owned cudaStream_t (NULL = its own stream) */
int amgb_hierarchy_set_stream(amgb_hierarchy* h, void* cuda_stream);

/* getters (multigrid.hpp:339-354) */
int amgb_hierarchy_n_levels(const amgb_hierarchy* h);
int64_t amgb_hierarchy_n_dofs(const amgb_hierarchy* h, int level);
int64_t amgb_hierarchy_nnz(const amgb_hierarchy* h, int level);        /* structural (Eigen) */
int64_t amgb_hierarchy_nnz_device(const amgb_hierarchy* h, int level); /* streamed by kernels */
double amgb_hierarchy_tolerance(const amgb_hierarchy* h);
/* host copy of level matrix (structural CSC incl. explicit zeros) */
int amgb_hierarchy_get_matrix(const amgb_hierarchy* h, int level, int* colptr, int* rowidx,
                              double* val);
int amgb_hierarchy_get_soln(amgb_hierarchy* h, int level, double* u);
int amgb_hierarchy_get_rhs(amgb_hierarchy* h, int level, double* f);
int amgb_hierarchy_set_soln(amgb_hierarchy* h, int level, const double* u);
int amgb_hierarchy_set_rhs(amgb_hierarchy* h, int level, const double* f);
/* Row-block variants for sharded hierarchies: the caller passes / receives only the rows
 * [begin, end) of amgb_hierarchy_local_range (the whole vector when the level is not sharded).
 * No gather: each rank moves 1/world of the bytes.  The setters are collective (they refresh
 * the neighbours' ghost rows). */
int amgb_hierarchy_get_soln_local(amgb_hierarchy* h, int level, double* u_block);
int amgb_hierarchy_get_rhs_local(amgb_hierarchy* h, int level, double* f_block);
int amgb_hierarchy_set_soln_local(amgb_hierarchy* h, int level, const double* u_block);
int amgb_hierarchy_set_rhs_local(amgb_hierarchy* h, int level, const double* f_block);
/* colouring per level (COLOR_GS only) */
int amgb_hierarchy_get_coloring(const amgb_hierarchy* h, int level, int* n_colors, int* color);

/* one V-cycle on device-resident state (multigrid.hpp:263-305); asynchronous
 * on the handle's stream */
int amgb_vcycle(amgb_hierarchy* h);
/* n back-to-back V-cycles, then a stream synchronise */
int amgb_vcycles(amgb_hierarchy* h, int64_t n);
/* sum r^2 on the finest level (common.hpp:17-27); synchronises */
int amgb_hierarchy_rss(amgb_hierarchy* h, double* out);
/* Multigrid::solve (multigrid.hpp:311-337).  Beyond the reference (which only
 * prints them) the iteration count, last error and the error history are kept. */
int amgb_solve(amgb_hierarchy* h, int64_t* iters_done, double* last_error);
/* variant with the north star's criterion: stop when ||r||_2 / ||b||_2 <= rel_tol,
 * checked every compute_error_every_n_iters cycles, capped at n_iters */
int amgb_solve_relative(amgb_hierarchy* h, double rel_tol, int64_t* iters_done,
                        double* last_rel_residual);
/* Beyond the reference (README.md:127 of the reference names it as
 * the usual use): conjugate gradients on A u = b preconditioned by one V-cycle from a zero guess,
 * starting from the stored level-0 solution; stops when ||r||_2 / ||b||_2 <= rel_tol or after
 * max_iters iterations.  The residual history is kept like amgb_solve's.  Single GPU; the cycle
 * must be symmetric (damped Jacobi / multicolour GS / symmetric GS with equal pre- and
 * post-smoothing, which is how every cycle of this library is built). */
int amgb_solve_pcg(amgb_hierarchy* h, double rel_tol, int64_t max_iters, int64_t* iters_done,
                   double* last_rel_residual);
int64_t amgb_hierarchy_iters_done(const amgb_hierarchy* h);
/* copies min(cap, n) history entries, returns n */
int64_t amgb_hierarchy_error_history(const amgb_hierarchy* h, double* out, int64_t cap);
int amgb_synchronize(amgb_hierarchy* h);

/* per-operator entry points on one level, host vectors in/out (parity tests).
 * restriction: include/amg/interpolator.hpp:64-68; prolongation + add:
 * interpolator.hpp:52-56 with multigrid.hpp:294-296. */
int amgb_restrict(amgb_hierarchy* h, int level, const double* r_fine, double* f_coarse);
int amgb_prolong_add(amgb_hierarchy* h, int level, const double* e_coarse, double* u_fine);
/* smoother->smooth(A_l, u_l, f_l) on the device-resident level state */
int amgb_smooth_level(amgb_hierarchy* h, int level);
/* r_l = f_l - A_l u_l of the device-resident level state, copied to host */
int amgb_residual_level(amgb_hierarchy* h, int level, double* r);
/* fused residual + restriction as used inside the V-cycle: f_{l+1} = R_l (f_l - A_l u_l),
 * u_{l+1} = 0 (multigrid.hpp:272-282) */
int amgb_residual_restrict_level(amgb_hierarchy* h, int level);
/* coarsest direct solve u_L = A_L^{-1} f_L (multigrid.hpp:287-288) */
int amgb_coarse_solve(amgb_hierarchy* h);

/* 1 when `level` runs as fused legs (option fuse bit 2; whole banded levels of a damped-Jacobi
 * cycle), else 0.  amgb_hierarchy_leg_plan reports the tiling of its down (up = 0) or up leg:
 * info[0..9] = line length m, element halo per stage, strip width W, lines per tile, tiles
 * (= CTAs), strips, TMA lines in flight, threads per CTA, dynamic shared memory bytes, chained
 * stencil stages. */
int amgb_hierarchy_fused_legs(const amgb_hierarchy* h, int level);
/* 1 when the level's fused legs run matrix-free (option fuse bit 6, operator verified at setup) */
int amgb_hierarchy_matrix_free(const amgb_hierarchy* h, int level);
/* number of distinct operator rows when the level's fused legs run from a row-type dictionary
 * (option fuse bit 7), else 0 */
int amgb_hierarchy_dictionary_types(const amgb_hierarchy* h, int level);
/* first level of the coarse tail that runs in one launch (option fuse bit 4), -1 if none */
int amgb_hierarchy_tail_first(const amgb_hierarchy* h);
int amgb_hierarchy_leg_plan(const amgb_hierarchy* h, int level, int up, int64_t* info);
/* levels [first, end) that run inside the two mid-level kernels (option fuse bit 5; first = -1:
 * none), rows of the first of them a block owns, and the number of blocks */
int amgb_hierarchy_mid_range(const amgb_hierarchy* h, int* first, int* end, int* tile_rows, int* blocks);

/* Device-side Galerkin product of one level (measured, not yet used by
 * the setup): A_{level+1} = R (A_level P) from the level's device mirror, one thread per coarse row,
 * in the reference's evaluation order (multigrid.hpp:219-223).  Reports the kernel time and the
 * number of entries that differ bitwise from the host-built mirror of level + 1 (0 expected). */
int amgb_hierarchy_galerkin_device(amgb_hierarchy* h, int level, double* ms, int64_t* mismatches);

/* Mean milliseconds of the four phases of a V-cycle on this rank over `reps` un-captured cycles:
 * out[0] sharded down legs, out[1] gather of the first replicated right-hand side, out[2] the levels
 * below the sharded ones (on one GPU: the whole cycle), out[3] sharded up legs.  Collective on a
 * sharded hierarchy.  Advances the level state by reps + 1 cycles. */
int amgb_hierarchy_phase_times(amgb_hierarchy* h, int reps, double* out);

/* counters: kernels launched by this library in this process, and per V-cycle */
int64_t amgb_kernel_launches(void);
int64_t amgb_hierarchy_launches_per_vcycle(const amgb_hierarchy* h);

/* device layout of a level's operator and the matrix bytes one pass really streams */
#define AMGB_FORMAT_SELL 0 /* sliced ELL, 32 rows per slice, explicit column indices */
#define AMGB_FORMAT_DIA 1  /* diagonal storage (banded operators): no index array     */
int amgb_hierarchy_format(const amgb_hierarchy* h, int level);
int64_t amgb_hierarchy_matrix_bytes(const amgb_hierarchy* h, int level);
/* diagonals of the level's DIA mirror (0: SELL layout).  The fused legs stream all of them, 8 bytes per
 * row each (amgb_hierarchy_matrix_bytes excludes the 32-row slices the per-operator kernels skip). */
int amgb_hierarchy_n_diagonals(const amgb_hierarchy* h, int level);
/* AMGB_GS_KERNEL_* that smooths `level` of a Gauss-Seidel hierarchy (-1: not a Gauss-Seidel hierarchy) */
int amgb_hierarchy_gs_kernel(const amgb_hierarchy* h, int level);

/* algorithmic byte counts: B_l = 12 nnz_l + 28 N_l + 4 with
 * nnz_l the entries the kernels stream (explicit zeros pruned) */
int64_t amgb_hierarchy_pass_bytes(const amgb_hierarchy* h, int level);
int64_t amgb_hierarchy_vcycle_bytes(const amgb_hierarchy* h);

/* ------------------------------------------------------------------------
 * Per-kernel timing hooks for bench.py (CUDA events on the handle's stream).
 * kind: 0 smoother pass (one Jacobi sweep / one colour-complete pass / one GS
 * direction), 1 residual, 2 residual+restrict, 3 prolong+add, 4 fused down leg,
 * 5 fused up leg (levels with amgb_hierarchy_fused_legs), 6 / 7 the mid-level down / up
 * kernel, 8 the coarse tail (level ignored for 6-8); 9 one damped-Jacobi sweep and 11 one
 * residual of `level` INCLUDING the halo exchange with ranks +-1 on a sharded level
 * (collective).  Runs `reps`
 * launches after `warmup`, returns the mean milliseconds per launch.
 * ---------------------------------------------------------------------- */
int amgb_time_kernel(amgb_hierarchy* h, int level, int kind, int warmup, int reps,
                     double* ms_per_launch);

#ifdef __cplusplus
}
#endif
#endif /* AMGB_H_ */

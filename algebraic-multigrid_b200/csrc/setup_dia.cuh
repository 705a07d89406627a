// Device-side hierarchy setup on the diagonal (DIA) layout (reference:
// the Multigrid constructor, include/amg/multigrid.hpp:190-243).  The level-0 operator arrives as
// the raw CSC arrays of an Eigen::SparseMatrix; everything from there on -- conversion to DIA,
// symmetry check / transposition, the Galerkin products A_{l+1} = R (A_l P) (galerkin_dia.cuh),
// pruning of empty diagonals, slice masks, row-block and window mirrors of sharded levels -- runs
// on the GPU, so the host never touches the O(nnz) data.  Values are bit-identical to the host
// setup (host_setup.cpp) because k_galerkin_dia evaluates every entry in Eigen's order.
#pragma once
#include <cstdint>

#include <cuda_runtime.h>

namespace amgb {
namespace setup {

constexpr int kMaxOffsets = 16;
struct Offsets {
  int n;
  int v[kMaxOffsets];
};

// ---- distinct offsets rowidx[p] - column of a CSC matrix: bitmap over [-(n_rows-1), n_cols-1] ----
__global__ void __launch_bounds__(256) k_mark_offsets(const int* __restrict__ colptr, const int* __restrict__ rowidx,
                                                      int n_cols, int shift, unsigned* bitmap) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cols) return;
  for (int p = colptr[c]; p < colptr[c + 1]; ++p) {
    const unsigned o = (unsigned)(rowidx[p] - c + shift);
    const unsigned bit = 1u << (o & 31u);
    if (!(bitmap[o >> 5] & bit)) atomicOr(bitmap + (o >> 5), bit);  // a handful of words: mostly cached reads
  }
}
// list[0] = count (may exceed cap: too many diagonals), list[1 + i] = offsets in no particular order
__global__ void __launch_bounds__(256) k_collect_offsets(const unsigned* __restrict__ bitmap, int n_words, int shift,
                                                         int cap, int* list) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_words) return;
  unsigned bits = bitmap[w];
  while (bits) {
    const int b = __ffs(bits) - 1;
    bits &= bits - 1;
    const int slot = atomicAdd(list, 1);
    if (slot < cap) list[1 + slot] = w * 32 + b - shift;
  }
}

// ---- CSC -> DIA of "row c = CSC column c" (smoother.hpp:101-117); val pre-zeroed ----
__global__ void __launch_bounds__(256) k_csc_to_dia(const int* __restrict__ colptr, const int* __restrict__ rowidx,
                                                    const double* __restrict__ val, int n_cols, Offsets off,
                                                    double* __restrict__ dia, int ld) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cols) return;
  for (int p = colptr[c]; p < colptr[c + 1]; ++p) {
    const int o = rowidx[p] - c;
#pragma unroll
    for (int d = 0; d < kMaxOffsets; ++d)
      if (d < off.n && off.v[d] == o) dia[(size_t)d * ld + c] = val[p];
  }
}

// ---- bitwise symmetry of a square DIA operator: A(r, r+o) == A(r+o, r) ----
__global__ void __launch_bounds__(256) k_dia_asymmetric(const double* __restrict__ dia, int n, int ld, Offsets off,
                                                        int* flag) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  for (int d = 0; d < off.n; ++d) {
    const double a = dia[(size_t)d * ld + r];
    const int c = r + off.v[d];
    if (c < 0 || c >= n) {
      if (a != 0.0) *flag = 1;
      continue;
    }
    double b = 0.0;
    for (int e = 0; e < off.n; ++e)
      if (off.v[e] == -off.v[d]) b = dia[(size_t)e * ld + c];
    if (__double_as_longlong(a) != __double_as_longlong(b) && !(a == 0.0 && b == 0.0)) *flag = 1;
  }
}
// dst = transpose: dst(r + o, diagonal of -o) = src(r, diagonal of o); off_t lists the negated offsets ascending
__global__ void __launch_bounds__(256) k_dia_transpose(const double* __restrict__ src, int n, int ld, Offsets off,
                                                       Offsets off_t, double* __restrict__ dst) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  for (int d = 0; d < off.n; ++d) {
    const int c = r + off.v[d];
    if (c < 0 || c >= n) continue;
    for (int e = 0; e < off_t.n; ++e)
      if (off_t.v[e] == -off.v[d]) dst[(size_t)e * ld + c] = src[(size_t)d * ld + r];
  }
}

// ---- statistics: entries per diagonal, and the per-slice occupancy mask (bit d of mask[s] set when
// diagonal d has an entry among rows [32 s, 32 s + 32)), live = number of (slice, diagonal) pairs set
__global__ void __launch_bounds__(256) k_dia_stats(const double* __restrict__ dia, int n, int ld, int nd,
                                                   unsigned long long* count /* nd + 1 */, unsigned short* mask) {
  __shared__ unsigned int sh[kMaxOffsets + 1];
  if (threadIdx.x <= kMaxOffsets) sh[threadIdx.x] = 0;
  __syncthreads();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  unsigned m = 0;
  for (int d = 0; d < nd; ++d) {
    const bool nz = (r < n) && dia[(size_t)d * ld + r] != 0.0;
    const unsigned b = __ballot_sync(0xffffffffu, nz);
    if (b) m |= 1u << d;
    if (lane == 0 && b) {
      atomicAdd(&sh[d], __popc(b));
      atomicAdd(&sh[kMaxOffsets], 1u);
    }
  }
  if (lane == 0 && (r >> 5) < (ld >> 5)) mask[r >> 5] = (unsigned short)m;
  __syncthreads();
  if (threadIdx.x < nd && sh[threadIdx.x]) atomicAdd(count + threadIdx.x, (unsigned long long)sh[threadIdx.x]);
  if (threadIdx.x == kMaxOffsets && sh[kMaxOffsets]) atomicAdd(count + nd, (unsigned long long)sh[kMaxOffsets]);
}

// ---- rows [row_begin, row_begin + n_dst) of a DIA operator with local numbering; rows outside
// [0, n_src) stay empty (dst pre-zeroed) ----
__global__ void __launch_bounds__(256) k_dia_slice(const double* __restrict__ src, int n_src, int ld_src, int nd,
                                                   int row_begin, double* __restrict__ dst, int n_dst, int ld_dst) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_dst) return;
  const int r = row_begin + t;
  if (r < 0 || r >= n_src) return;
  for (int d = 0; d < nd; ++d) dst[(size_t)d * ld_dst + t] = src[(size_t)d * ld_src + r];
}

// ---- is the operator exactly a constant five-point stencil?  Row t of the mirror is global row
// base + t of a level with n_global rows and grid lines of m rows; the entries on the offsets
// -m, -1, 0, +1, +m must be c[0..4] wherever that neighbour exists (inside the level, and inside the
// grid line for -1 / +1) and zero elsewhere -- compared bit for bit.  *bad != 0: it is not.
struct Const5 {
  double c[5];
};
__global__ void __launch_bounds__(256) k_check_const5(const double* __restrict__ val, int n_local, int ld, int base,
                                                      int n_global, int m, Const5 C, int* bad) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_local) return;
  const int kg = base + t;
  const bool in = kg >= 0 && kg < n_global;
  const int pos = in ? kg % m : 0;
  const bool present[5] = {in && kg >= m, in && pos != 0, in, in && pos != m - 1, in && kg < n_global - m};
  for (int d = 0; d < 5; ++d) {
    const double want = present[d] ? C.c[d] : 0.0;
    if (__double_as_longlong(val[(size_t)d * ld + t]) != __double_as_longlong(want)) *bad = 1;
  }
}

// ---- row-type dictionary: the distinct rows of a DIA operator (the Galerkin levels of a constant
// stencil have a few dozen), found with a small open-addressing hash table of 64-bit row hashes.
// Every row is afterwards compared bit for bit with its table row (k_dict_assign), so a hash
// collision can only make the build fail, never produce a wrong operator.
constexpr int kDictSlots = 4096;
__device__ __forceinline__ unsigned long long dict_row_hash(const double* __restrict__ val, int ld, int nd, int t) {
  unsigned long long h = 0x9E3779B97F4A7C15ull;
  for (int d = 0; d < nd; ++d) {
    h ^= (unsigned long long)__double_as_longlong(val[(size_t)d * ld + t]);
    h *= 0xFF51AFD7ED558CCDull;
    h ^= h >> 32;
  }
  return h ? h : 1ull;  // 0 marks an empty slot
}
// slot of hash h (inserting it when insert is set); -1: table full
__device__ __forceinline__ int dict_find(unsigned long long* keys, unsigned long long h, bool insert) {
  int slot = (int)(h % kDictSlots);
  for (int probes = 0; probes < kDictSlots; ++probes) {
    const unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(keys + slot);
    if (cur == h) return slot;
    if (cur == 0ull) {
      if (!insert) return -1;
      const unsigned long long old = atomicCAS(keys + slot, 0ull, h);
      if (old == 0ull || old == h) return slot;
    }
    slot = (slot + 1 == kDictSlots) ? 0 : slot + 1;
  }
  return -1;
}
__global__ void __launch_bounds__(256) k_dict_insert(const double* __restrict__ val, int n, int ld, int nd,
                                                     unsigned long long* keys, int* rep, int* overflow) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int slot = dict_find(keys, dict_row_hash(val, ld, nd, t), true);
  if (slot < 0) {
    *overflow = 1;
    return;
  }
  if (*reinterpret_cast<volatile int*>(rep + slot) > t) atomicMin(rep + slot, t);  // representative = first row
}
__global__ void __launch_bounds__(256) k_dict_table(const double* __restrict__ val, int ld, int nd, const int* rep_of_id,
                                                    int n_types, double* table) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_types * nd) return;
  table[i] = val[(size_t)(i % nd) * ld + rep_of_id[i / nd]];
}
__global__ void __launch_bounds__(256) k_dict_assign(const double* __restrict__ val, int n, int ld, int nd,
                                                     unsigned long long* keys, const short* slot_id,
                                                     const double* __restrict__ table, unsigned char* tid, int* bad) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int slot = dict_find(keys, dict_row_hash(val, ld, nd, t), false);
  const int id = slot >= 0 ? (int)slot_id[slot] : -1;
  if (id < 0) {
    *bad = 1;
    return;
  }
  for (int d = 0; d < nd; ++d)
    if (__double_as_longlong(val[(size_t)d * ld + t]) != __double_as_longlong(table[id * nd + d])) *bad = 1;
  tid[t] = (unsigned char)id;
}

}  // namespace setup
}  // namespace amgb

// K7  interface Schwarz smoother: batched small dense patch solves.
//
// Replaces HAZmath smoother_dcsr_Schwarz_forward/backward (one UMFPACK solve per block,
// sequential over blocks; blocks factorised once at setup) configured by
// src/amg_parameters.py:82-87.  For every patch B:
//     x_B <- x_B + A_BB^{-1} (b - A x)_B
// Patches are coloured by conflict (schwarz_color in csrc/host/setup.cpp), so all patches of
// one colour are solved concurrently with a result identical to visiting them one by one;
// forward = colours ascending, backward = descending.
//
// Setup (once, on the device): every A_BB is gathered from the level's CSR, Cholesky-factorised
// and inverted in shared memory; the symmetric inverse is kept packed (s(s+1)/2 doubles per
// patch) -- the analogue of HAZmath keeping one UMFPACK factorisation per block.
//
// Apply, per patch (one warp for <= 32 dofs, else one CTA):
//   1. x on the patch neighbourhood N[B] (the union of the columns of the patch rows, ~130
//      dofs for a 30-dof 3-D patch) is gathered ONCE into shared memory through a stored list;
//      the row products then index shared memory through a stored 16-bit local column, so the
//      900 row entries of a patch cost 900 coalesced value loads instead of 900 scattered
//      32-byte sector gathers;
//   2. the packed inverse is staged through shared memory with coalesced loads;
//   3. residual rows (8 rows in flight per warp), symmetric packed mat-vec, update of x_B.
// Algorithmic bytes per patch: 8 B per row entry (values) + 2 B (local column) + 4 B per
// neighbour + 8 B per gathered x + 4 s(s+1) B of inverse + 28 B per patch dof.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cuda_runtime.h>
#include <cstring>
#include <functional>
#include <numeric>
#include <string>
#include <unordered_map>
#include <vector>

#include "../host/hierarchy.h"
#include "kernels.cuh"

namespace mamg {

struct SwPatch {      // 32 bytes per patch
  int q0;             // first entry in pdof arrays
  int s;              // dofs
  int n0;             // first entry in nbr
  int nn;             // neighbours
  long long e0;       // first entry in lcol
  long long i0;       // first entry in pinv
};

// Jagged layout of the patch row values on the fast path: the rows of a patch are ordered by
// decreasing number of off-patch entries, and entry e is stored only for the first cnt[e] lanes
// (cnt = the level-wide maximum over all patches), at offset off[e] inside the patch's slab.
struct SwProfile {
  int off[33];
  int cnt[32];
};

struct DSchwarz {
  int npatch = 0, ncolors = 0, max_size = 0, max_nbr = 0, srow = 1, warps = 1, ppc = 1;
  size_t smem_apply = 0, smem_setup = 0;
  SwPatch* pat = nullptr;      // patches sorted by colour
  int* pidx = nullptr;         // per patch dof: permuted row id (natural-ascending order inside a patch)
  int* prow = nullptr;         // per patch dof: start of the row in the level CSR (= ia[pidx])
  int* plen = nullptr;         // per patch dof: length of the row
  int* nbr = nullptr;          // neighbourhood lists (permuted ids, ascending)
  uint16_t* lcol = nullptr;    // per patch: s x srow (row-major, padded): position of the entry's column in nbr
  double* pinv = nullptr;      // packed lower triangles of A_BB^{-1}
  // fast path (every patch <= 32 dofs, <= 255 neighbours, rows <= 32 entries): per-patch blobs in
  // "lane = patch row" layout, uniform strides, so all addresses follow from the patch number
  bool fast = false;
  int nbq = 0, sq = 0, inv_stride = 0, sr_t = 32, vstride = 0;
  SwProfile prof;
  int* pidx32 = nullptr;       // [np][32] row id of patch dof k, -1 padding
  int* nbrp = nullptr;         // [np][nbq][32] neighbourhood list, padded with a valid index
  double* vt = nullptr;        // [np][srow][32] entry e of row k; 0 padding
  uint32_t* ct4 = nullptr;     // [np][sq][32] four 8-bit local columns per word; 255 = zero slot
  // de-duplication: patches whose blocks, off-patch values and local columns are bit-identical (most
  // patches of a uniform mesh with constant coefficients) share one stored blob; pinv / vt / ct4 (and
  // lcol on the general path) are indexed by the unique id, which keeps them L2-resident
  int nuniq = 0;
  int* uid = nullptr;          // [np] unique blob of every patch (fast path)
  long long* inv_off = nullptr;  // [nuniq] offset of the packed inverse of a unique blob
  // patches sorted by (conflict colour, block of the seed): cb_ptr[c*nb + b] .. [c*nb + b + 1]
  int nb = 1;
  std::vector<int> cb_ptr;     // host: size ncolors*nb + 1
  std::vector<int> qoff;       // host: first entry of every patch in pidx (size npatch + 1)
  // multi-GPU: after a conflict colour only the updated dofs that another part reads (or that sit in
  // another part's row block) are exchanged; xoff[c*nb + b] .. delimit them inside xidx
  std::vector<int> xoff;
  int* xidx = nullptr;
  // host, permuted dof ids (only with nb > 1): bit p of readers[j] = a patch seeded in part p gathers x[j]
  // (j outside that patch); bit p of members[j] = a patch seeded in part p contains j (reads b[j], writes x[j])
  std::vector<unsigned long long> readers, members;
  // host copies of the export lists: xidx and, per entry, the parts that need the value (gather it or own the row)
  std::vector<int> h_xidx;
  std::vector<unsigned long long> h_xmask;
  long long alg_bytes = 0;     // algorithmic bytes of one sweep over all patches (shared blobs once per colour)
  long long alg_bytes_stored = 0;  // the same with every patch owning its data (SURVEY 8d "stored factors")
};

__device__ __forceinline__ int tri(int i, int j) { return i * (i + 1) / 2 + j; }  // i >= j

// ---- setup: packed inverse of every A_BB -------------------------------------------------------
template <int T>
__global__ void __launch_bounds__(T)
schwarz_invert_kernel(int nuniq, const int* __restrict__ rep, const long long* __restrict__ inv_off,
                      const SwPatch* __restrict__ pat, const int* __restrict__ pidx,
                      const int* __restrict__ ia, const int* __restrict__ ja,
                      const double* __restrict__ a, double* __restrict__ pinv, int max_size) {
  extern __shared__ double smem[];
  if ((int)blockIdx.x >= nuniq) return;
  const int patch = rep[blockIdx.x];   // the representative patch of unique blob blockIdx.x
  const int q0 = pat[patch].q0, s = pat[patch].s;
  double* Lm = smem;                                                  // packed lower triangle
  double* col = Lm + (size_t)max_size * (max_size + 1) / 2;           // s scratch
  int* idx = reinterpret_cast<int*>(col + max_size);                  // s
  const int tid = threadIdx.x;
  for (int k = tid; k < s * (s + 1) / 2; k += T) Lm[k] = 0.0;
  for (int k = tid; k < s; k += T) idx[k] = pidx[q0 + k];
  __syncthreads();
  // gather the lower triangle of A_BB: linear search of the column among the patch dofs
  for (int k = 0; k < s; ++k) {
    const int i = idx[k];
    for (int p = ia[i] + tid; p < ia[i + 1]; p += T) {
      const int j = ja[p];
      for (int c = 0; c <= k; ++c)
        if (idx[c] == j) { Lm[tri(k, c)] = a[p]; break; }
    }
  }
  __syncthreads();
  // Cholesky A_BB = L L'
  for (int j = 0; j < s; ++j) {
    const double d = sqrt(Lm[tri(j, j)]);
    const double invd = 1.0 / d;
    __syncthreads();
    if (tid == 0) Lm[tri(j, j)] = d;
    for (int i = j + 1 + tid; i < s; i += T) Lm[tri(i, j)] *= invd;
    __syncthreads();
    for (int k = j + 1 + tid / 32; k < s; k += T / 32) {
      const double lkj = Lm[tri(k, j)];
      for (int i = k + (tid & 31); i < s; i += 32) Lm[tri(i, k)] -= Lm[tri(i, j)] * lkj;
    }
    __syncthreads();
  }
  // W = L^{-1} in place (column by column from the right; LAPACK dtrti2 'L')
  for (int j = s - 1; j >= 0; --j) {
    const double wjj = 1.0 / Lm[tri(j, j)];
    for (int i = j + 1 + tid; i < s; i += T) col[i] = Lm[tri(i, j)];
    __syncthreads();
    for (int i = j + 1 + tid; i < s; i += T) {
      double acc = 0.0;
      for (int k = j + 1; k <= i; ++k) acc += Lm[tri(i, k)] * col[k];   // W[i][k] (already inverted) * L[k][j]
      Lm[tri(i, j)] = -wjj * acc;
    }
    if (tid == 0) Lm[tri(j, j)] = wjj;
    __syncthreads();
  }
  // A_BB^{-1} = W' W in place, rows ascending (LAPACK dlauu2 'L')
  for (int i = 0; i < s; ++i) {
    for (int c = tid; c <= i; c += T) {
      double acc = 0.0;
      for (int k = i; k < s; ++k) acc += Lm[tri(k, i)] * Lm[tri(k, c)];
      col[c] = acc;
    }
    __syncthreads();
    for (int c = tid; c <= i; c += T) Lm[tri(i, c)] = col[c];
    __syncthreads();
  }
  double* out = pinv + inv_off[blockIdx.x];
  for (int k = tid; k < s * (s + 1) / 2; k += T) out[k] = Lm[k];
}

// ---- apply ---------------------------------------------------------------------------------------
// Shared-memory layout of one patch slot (sizes from the level maxima; 16-byte aligned blocks).
// Row values and local columns sit row-major with an ODD row stride `srow`, so that "thread k
// walks row k" is bank-conflict free.
struct SwLayout {
  int max_size, max_nbr, srow;
  __host__ __device__ size_t ent() const { return (size_t)max_size * srow; }
  __host__ __device__ size_t inv_d() const { return ((size_t)max_size * (max_size + 1) / 2 + 2) & ~(size_t)1; }
  __host__ __device__ size_t ls_d() const { return ((ent() + 8) * 2 + 15) / 16 * 2; }  // doubles holding the uint16 columns
  __host__ __device__ size_t as_d() const { return (ent() + 1) & ~(size_t)1; }
  __host__ __device__ size_t xs_d() const { return ((size_t)max_nbr + 2) & ~(size_t)1; }
  __host__ __device__ size_t rhs_d() const { return ((size_t)max_size + 1) & ~(size_t)1; }
  __host__ __device__ size_t int_d() const { return (3 * (size_t)max_size + 4) / 2 + 1; }
  __host__ __device__ size_t total_d() const { return (inv_d() + ls_d() + as_d() + xs_d() + rhs_d() + int_d() + 1) & ~(size_t)1; }
};

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
  const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// WPP warps per patch, PPC patches per CTA (PPC > 1 only with WPP == 1: warps are independent).
// Every byte a patch needs is requested up front with cp.async (two dependent waves: the
// descriptor-addressed arrays, then the row values / x gathers they point to), so one warp has a
// whole patch (~15 KB) in flight; the arithmetic then runs out of shared memory with one thread
// per patch row (no shuffles) and a packed symmetric mat-vec.
template <int WPP, int PPC, bool TAIL>
__global__ void __launch_bounds__(WPP * PPC * 32)
schwarz_apply_kernel(int p0, int p1, const SwPatch* __restrict__ pat, const int* __restrict__ pidx,
                     const int* __restrict__ prow, const int* __restrict__ plen,
                     const int* __restrict__ nbr, const uint16_t* __restrict__ lcol,
                     const double* __restrict__ pinv, const double* __restrict__ a,
                     const double* __restrict__ b, double* x, SwLayout lay, const HaloTail tail) {
  constexpr int T = WPP * 32;  // threads per patch
  extern __shared__ __align__(16) double smem[];
  const int slot = threadIdx.x / T;
  const int tid = threadIdx.x % T;
  const int patch = p0 + blockIdx.x * PPC + slot;
  const int S = lay.srow;
  double* Inv = smem + slot * lay.total_d();
  uint16_t* Ls = reinterpret_cast<uint16_t*>(Inv + lay.inv_d());
  double* As = Inv + lay.inv_d() + lay.ls_d();
  double* xs = As + lay.as_d();
  double* rhs = xs + lay.xs_d();
  int* idx = reinterpret_cast<int*>(rhs + lay.rhs_d());   // s
  int* rst = idx + lay.max_size;                           // s
  int* rln = rst + lay.max_size;                           // s
  const bool active = patch < p1;   // uniform per warp when PPC > 1 (WPP == 1)
  SwPatch P;
  P.s = 0;
  if (active) {
    P = pat[patch];
    // wave 1: packed inverse and local columns (16-byte chunks; both segments are 16-byte aligned)
    const double* src = pinv + P.i0;
    const int nch = (P.s * (P.s + 1) / 2 + 1) / 2;
    for (int k = tid; k < nch; k += T) cp_async16(Inv + 2 * k, src + 2 * k);
    const uint16_t* lsrc = lcol + P.e0;
    const int lch = (P.s * S + 7) / 8;
    for (int k = tid; k < lch; k += T) cp_async16(Ls + 8 * k, lsrc + 8 * k);
    for (int k = tid; k < P.s; k += T) {
      const int i = pidx[P.q0 + k];
      idx[k] = i;
      rst[k] = prow[P.q0 + k];
      rln[k] = plen[P.q0 + k];
      cp_async8(rhs + k, b + i);
    }
    for (int j = tid; j < P.nn; j += T) cp_async8(xs + j, x + nbr[P.n0 + j]);
    if (tid == 0) xs[P.nn] = 0.0;   // the slot every entry inside the patch points to (its coupling lives in the inverse)
  }
  if (WPP == 1) __syncwarp(); else __syncthreads();
  if (active) {
    // wave 2: the values of the patch rows (row k -> As[k*S ...]), one warp per row, coalesced
    const int lane = tid & 31, wrp = tid / 32;
    for (int k = wrp; k < P.s; k += WPP) {
      const int r0 = rst[k], len = rln[k];
      for (int e = lane; e < len; e += 32) cp_async8(As + k * S + e, a + r0 + e);
    }
  }
  cp_async_wait_all();
  if (WPP == 1) __syncwarp(); else __syncthreads();
  if (active) {
    // residual: thread k owns row k
    for (int k = tid; k < P.s; k += T) {
      const double* ar = As + k * S;
      const uint16_t* lr = Ls + k * S;
      const int len = rln[k];
      double acc = 0.0;
      for (int e = 0; e < len; ++e) acc += ar[e] * xs[lr[e]];
      rhs[k] -= acc;
    }
  }
  if (WPP == 1) __syncwarp(); else __syncthreads();
  if (active) {
    // delta = A_BB^{-1} rhs with the packed symmetric inverse: entry (k,c) lives at tri(max,min);
    // walk c with two running addresses (row part c <= k, column part c > k)
    for (int k = tid; k < P.s; k += T) {
      double d = 0.0;
      int a1 = k * (k + 1) / 2;   // (k, c) for c <= k
      int a2 = a1 + k;            // (c, k) for c >= k, advanced by c + 1
      for (int c = 0; c < P.s; ++c) {
        const int ad = c <= k ? a1 + c : a2;
        d += Inv[ad] * rhs[c];
        if (c >= k) a2 += c + 1;
      }
      x[idx[k]] = d;   // x_B = A_BB^{-1} (b_B - A_{B,out} x_out)
    }
  }
  if (TAIL) halo_tail(tail);   // multi-GPU halo mode: the last block sends the dofs this colour's patches updated to the neighbours
}

// ---- fast path -----------------------------------------------------------------------------------
// setup: build vt / ct4 of one patch from the level CSR (one warp per patch, lane = patch row)
__global__ void __launch_bounds__(256)
schwarz_index_kernel(int np, const SwPatch* __restrict__ pat, const int* __restrict__ pidx,
                     const int* __restrict__ nbr, int nbq, int* __restrict__ pidx32, int* __restrict__ nbrp) {
  // per-patch index data of the fast path: row ids (lane = patch row) and the padded neighbour list
  const int patch = blockIdx.x * 8 + threadIdx.x / 32, lane = threadIdx.x % 32;
  if (patch >= np) return;
  const SwPatch P = pat[patch];
  const size_t pp = (size_t)patch;
  pidx32[pp * 32 + lane] = lane < P.s ? pidx[P.q0 + lane] : -1;
  const int* nb = nbr + P.n0;
  for (int j = 0; j < nbq; ++j) {
    const int q = j * 32 + lane;
    nbrp[(pp * nbq + j) * 32 + lane] = q < P.nn ? nb[q] : 0;   // padding: any valid row id (never referenced)
  }
}

// 128-bit signature of everything the stored blob of a patch depends on: its size and, row by row in
// lane order, every matrix entry as (inside the patch: local row | outside: position in the neighbour
// list, value bits).  Equal signatures <=> bit-identical A_BB, off-patch values and local columns.
struct SwSig { unsigned long long a, b; };
__device__ __forceinline__ unsigned long long sw_mix(unsigned long long h, unsigned long long v, unsigned long long m) {
  h ^= v + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2);
  h *= m;
  h ^= h >> 29;
  return h;
}
__global__ void __launch_bounds__(256)
schwarz_sig_kernel(int np, const SwPatch* __restrict__ pat, const int* __restrict__ pidx,
                   const int* __restrict__ nbr, const int* __restrict__ ia, const int* __restrict__ ja,
                   const double* __restrict__ a, SwSig* __restrict__ sig) {
  const int patch = blockIdx.x * 8 + threadIdx.x / 32, lane = threadIdx.x % 32;
  if (patch >= np) return;
  const SwPatch P = pat[patch];
  const int row = lane < P.s ? pidx[P.q0 + lane] : -1;
  const int* nb = nbr + P.n0;
  const int r0 = row >= 0 ? ia[row] : 0, len = row >= 0 ? ia[row + 1] - r0 : 0;
  int maxlen = len;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o));
  unsigned long long h1 = 0x243f6a8885a308d3ULL + lane, h2 = 0x13198a2e03707344ULL ^ ((unsigned long long)lane << 32);
  for (int e = 0; e < maxlen; ++e) {
    const bool on = e < len;
    const int col = on ? ja[r0 + e] : -2;
    long long code = -1;
    if (on && P.nn > 0) {
      int lo = 0, hi = P.nn - 1;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (nb[mid] < col) lo = mid + 1; else hi = mid; }
      if (nb[lo] == col) code = (2LL << 32) | lo;
    }
    for (int t = 0; t < 32; ++t) {   // inside the patch: which lane owns that row
      const int rt = __shfl_sync(0xffffffffu, row, t);
      if (on && code < 0 && rt == col) code = (1LL << 32) | t;
    }
    if (on) {
      const unsigned long long vb = (unsigned long long)__double_as_longlong(a[r0 + e]);
      h1 = sw_mix(sw_mix(h1, (unsigned long long)code, 0xff51afd7ed558ccdULL), vb, 0xff51afd7ed558ccdULL);
      h2 = sw_mix(sw_mix(h2, vb, 0xc4ceb9fe1a85ec53ULL), (unsigned long long)code, 0xc4ceb9fe1a85ec53ULL);
    }
  }
  unsigned long long H1 = sw_mix(0x452821e638d01377ULL, (unsigned long long)P.s, 0xff51afd7ed558ccdULL);
  unsigned long long H2 = sw_mix(0xbe5466cf34e90c6cULL, (unsigned long long)P.s, 0xc4ceb9fe1a85ec53ULL);
  for (int t = 0; t < 32; ++t) {
    H1 = sw_mix(H1, __shfl_sync(0xffffffffu, h1, t), 0xff51afd7ed558ccdULL);
    H2 = sw_mix(H2, __shfl_sync(0xffffffffu, h2, t), 0xc4ceb9fe1a85ec53ULL);
  }
  if (lane == 0) { sig[patch].a = H1; sig[patch].b = H2; }
}

// setup: build vt / ct4 of unique blob u from the level CSR (one warp per blob, lane = row of its
// representative patch)
__global__ void __launch_bounds__(256)
schwarz_blob_kernel(int nuniq, const int* __restrict__ rep, const SwPatch* __restrict__ pat, const int* __restrict__ pidx,
                    const int* __restrict__ nbr, const int* __restrict__ ia, const int* __restrict__ ja,
                    const double* __restrict__ a, int srow, int sq, double* __restrict__ vt, uint32_t* __restrict__ ct4,
                    const SwProfile prof, int vstride) {
  const int u = blockIdx.x * 8 + threadIdx.x / 32, lane = threadIdx.x % 32;
  if (u >= nuniq) return;
  const SwPatch P = pat[rep[u]];
  const size_t pp = (size_t)u;
  const int row = lane < P.s ? pidx[P.q0 + lane] : -1;
  const int* nb = nbr + P.n0;
  // off-patch entries of this lane's row, compacted to the front (columns not found in the list of
  // outside neighbours belong to the patch itself and live in the stored inverse)
  const int r0 = row >= 0 ? ia[row] : 0, len = row >= 0 ? ia[row + 1] - r0 : 0;
  int src = 0;
  for (int q = 0; q < sq; ++q) {
    uint32_t word = 0;
    for (int u = 0; u < 4; ++u) {
      const int e = q * 4 + u;
      uint32_t loc = 255;
      double v = 0.0;
      while (src < len) {
        const int col = ja[r0 + src];
        int lo = 0, hi = P.nn - 1;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (nb[mid] < col) lo = mid + 1; else hi = mid; }
        if (P.nn > 0 && nb[lo] == col) { loc = (uint32_t)lo; v = a[r0 + src]; ++src; break; }
        ++src;
      }
      if (e < srow && lane < prof.cnt[e]) vt[pp * vstride + prof.off[e] + lane] = v;
      word |= loc << (8 * u);
    }
    ct4[(pp * sq + q) * 32 + lane] = word;
  }
}

#ifndef MAMG_SW_MINB24
#define MAMG_SW_MINB24 3   // CTAs per SM for the <24,4> specialisation (78 registers, no spills)
#endif
#ifndef MAMG_SW_MINB
#define MAMG_SW_MINB 2   // CTAs per SM the fast kernel is compiled for (2: ~100 registers, 3: 80 with spills)
#endif
constexpr int kSwFastWarps = 8;     // patches per CTA
constexpr int kSwFastSlot = 256 + 528 + 32;  // doubles per patch slot: xs[256], packed inverse (32*33/2), rhs[32]

// apply: one warp per patch, lane k = patch row k.  Wave 1 loads everything addressed by the patch
// number (row values, packed local columns, neighbour list, inverse via cp.async); wave 2 gathers
// x on the neighbourhood.  All loops have compile-time bounds; padding multiplies by the zero slot.
template <int SR, int NBQ>
__device__ __forceinline__ void
schwarz_fast_patch(int patch, int warp, int lane, double* smem, const int* __restrict__ pidx32, const int* __restrict__ nbrp,
                   const int* __restrict__ uid, const long long* __restrict__ inv_off,
                   const double* __restrict__ vt, const uint32_t* __restrict__ ct4,
                   const double* __restrict__ pinv, const double* __restrict__ b, double* x, int srow,
                   int sq, int nbq, const SwProfile& prof, int vstride) {
  double* xs = smem + warp * kSwFastSlot;
  double* Inv = xs + 256;
  double* rhs = Inv + 528;
  const size_t pp = (size_t)patch;
  const size_t uu = (size_t)uid[patch];   // the stored blob this patch shares with its look-alikes
  // ---- wave 1 ----
  const int my = pidx32[pp * 32 + lane];
  int nb[NBQ];
#pragma unroll
  for (int j = 0; j < NBQ; ++j) nb[j] = j < nbq ? ld_stream(nbrp + (pp * nbq + j) * 32 + lane) : 0;
  const int s = __popc(__ballot_sync(0xffffffffu, my >= 0));     // dofs of this patch
  {
    const double* src = pinv + inv_off[uu];
    const int nch = (s * (s + 1) / 2 + 1) / 2;                  // 16-byte chunks of its packed inverse
    for (int k = lane; k < nch; k += 32) cp_async16(Inv + 2 * k, src + 2 * k);
  }
  double v[SR];
  {
    const double* vp = vt + uu * vstride + lane;
#pragma unroll
    for (int e = 0; e < SR; ++e) v[e] = (e < srow && lane < prof.cnt[e]) ? ld_stream(vp + prof.off[e]) : 0.0;
  }
  uint32_t c4[SR / 4];
  {
    const uint32_t* cp = ct4 + (uu * sq) * 32 + lane;
#pragma unroll
    for (int q = 0; q < SR / 4; ++q) c4[q] = q < sq ? ld_stream(cp + q * 32) : 0xffffffffu;
  }
  const double bk = my >= 0 ? b[my] : 0.0;
  // ---- wave 2 ----
#pragma unroll
  for (int j = 0; j < NBQ; ++j)
    if (j < nbq) xs[j * 32 + lane] = x[nb[j]];
  if (lane == 31) xs[255] = 0.0;   // the zero slot every padded local column points to
  __syncwarp();
  double acc = 0.0;
#pragma unroll
  for (int e = 0; e < SR; ++e) acc += v[e] * xs[(c4[e / 4] >> (8 * (e % 4))) & 255u];
  rhs[lane] = my >= 0 ? bk - acc : 0.0;
  cp_async_wait_all();
  __syncwarp();
  double d = 0.0;
  const int base = lane * (lane + 1) / 2;
#pragma unroll
  for (int c = 0; c < 32; ++c) {
    if (c < s && lane < s) {   // only the s x s part of the staged inverse is defined
      const int ad = c <= lane ? base + c : c * (c + 1) / 2 + lane;
      d += Inv[ad] * rhs[c];
    }
  }
  if (my >= 0) x[my] = d;   // x_B = A_BB^{-1} (b_B - A_{B,out} x_out)
}

template <int SR, int NBQ, bool TAIL>
__global__ void __launch_bounds__(kSwFastWarps * 32, (SR <= 24 && NBQ <= 4) ? MAMG_SW_MINB24 : MAMG_SW_MINB)
schwarz_fast_kernel(int p0, int p1, const int* __restrict__ pidx32, const int* __restrict__ nbrp,
                    const int* __restrict__ uid, const long long* __restrict__ inv_off,
                    const double* __restrict__ vt, const uint32_t* __restrict__ ct4,
                    const double* __restrict__ pinv, const double* __restrict__ b, double* x, int srow,
                    int sq, int nbq, int smax, const SwProfile prof, int vstride, const HaloTail tail) {
  extern __shared__ __align__(16) double smem[];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int patch = p0 + blockIdx.x * kSwFastWarps + warp;
  if (patch < p1)
    schwarz_fast_patch<SR, NBQ>(patch, warp, lane, smem, pidx32, nbrp, uid, inv_off, vt, ct4, pinv, b, x, srow, sq, nbq, prof, vstride);
  if (TAIL) halo_tail(tail);   // multi-GPU halo mode: the last block sends the dofs this colour's patches updated to the neighbours
}

// Host side: reorder the patches by colour, translate to the permuted numbering, build the
// neighbourhood lists and local columns, upload, invert on the device.
// `alloc(bytes)` returns tracked device memory; pia/pja are the permuted CSR of the level.
inline void schwarz_upload(const Level& hl, int nb, const std::vector<int>& iperm, const int* d_ia,
                           const int* d_ja, const double* d_a, const std::vector<int>& pia,
                           const bigvec<int>& pja, const bigvec<double>& pa, DSchwarz& d,
                           const std::function<void*(size_t)>& alloc) {
  const SwPatch zero = {0, 0, 0, 0, 0, 0};
  const SchwarzPatches& sw = hl.sw;
  const int np = sw.npatch();
  const int n = hl.A.n;
  d.npatch = np;
  d.ncolors = sw.ncolors;
  d.max_size = sw.max_size;
  d.nb = nb;
  auto pkey = [&](int p) { return sw.color[p] * nb + (nb > 1 ? hl.part[sw.seed[p]] : 0); };
  d.cb_ptr.assign(sw.ncolors * nb + 1, 0);
  for (int p = 0; p < np; ++p) ++d.cb_ptr[pkey(p) + 1];
  for (int k = 0; k < sw.ncolors * nb; ++k) d.cb_ptr[k + 1] += d.cb_ptr[k];
  std::vector<int> order(np);
  {
    std::vector<int> fill(d.cb_ptr.begin(), d.cb_ptr.end() - 1);
    for (int p = 0; p < np; ++p) order[fill[pkey(p)]++] = p;
  }
  std::vector<SwPatch> pat(np, zero);
  std::vector<int> pidx(sw.dofs.size()), prow(sw.dofs.size()), plen(sw.dofs.size());
  // pass 1: sizes (rows, entries, neighbourhoods) per patch, in parallel
  std::vector<int> nn(np, 0), nn_out(np, 0), rlen_out(sw.dofs.size(), 0);
  std::vector<long long> ne(np, 0), ne_out(np, 0);
  int max_rowlen = 1, max_rowlen_out = 1;
#pragma omp parallel reduction(max : max_rowlen, max_rowlen_out)
  {
    std::vector<int> mark(n, -1), inb(n, -1);
#pragma omp for schedule(dynamic, 2048)
    for (int k = 0; k < np; ++k) {
      const int p = order[k];
      int cnt = 0, cnt_out = 0;
      long long ent = 0, ent_out = 0;
      for (int q = sw.ptr[p]; q < sw.ptr[p + 1]; ++q) inb[iperm[sw.dofs[q]]] = k;
      for (int q = sw.ptr[p]; q < sw.ptr[p + 1]; ++q) {
        const int i = iperm[sw.dofs[q]];
        ent += pia[i + 1] - pia[i];
        max_rowlen = std::max(max_rowlen, pia[i + 1] - pia[i]);
        int row_out = 0;
        for (int e = pia[i]; e < pia[i + 1]; ++e) {
          const bool outside = inb[pja[e]] != k;
          row_out += outside;
          if (mark[pja[e]] != k) { mark[pja[e]] = k; ++cnt; cnt_out += outside; }
        }
        ent_out += row_out;
        rlen_out[q] = row_out;
        max_rowlen_out = std::max(max_rowlen_out, row_out);
      }
      nn[k] = cnt;
      ne[k] = ent;
      nn_out[k] = cnt_out;
      ne_out[k] = ent_out;
    }
  }
  int max_nn = 0;   // on the fast path only the neighbours outside the patch are gathered
  for (int k = 0; k < np; ++k) max_nn = std::max(max_nn, nn_out[k]);
  const char* nofast = getenv("MAMG_SCHWARZ_GENERAL");
  const bool fast_shape = d.max_size <= 32 && max_nn <= 255 && max_rowlen_out <= 32 && !(nofast && atoi(nofast));
  // fast path: x_B + A_BB^{-1}(b - A x)_B = A_BB^{-1}(b_B - A_{B,out} x_out): the entries of the patch rows
  // that fall inside the patch are already in the stored inverse, so only the off-patch entries
  // (and the neighbours outside the patch) are kept -- about a quarter less traffic per patch
  const int srow = fast_shape ? max_rowlen_out : (max_rowlen | 1);   // general path: odd row stride (conflict-free)
  d.srow = srow;
  // every path uses x_B + A_BB^{-1}(b - A x)_B = A_BB^{-1}(b_B - A_{B,out} x_out): only the neighbours outside the
  // patch are gathered, the couplings inside the patch live in the stored inverse
  nn = nn_out;
  if (fast_shape) ne = ne_out;
  std::vector<int> dord(sw.dofs.size());   // sorted position -> position inside sw.dofs (identity on the general path)
  std::iota(dord.begin(), dord.end(), 0);
  SwProfile prof;
  std::memset(&prof, 0, sizeof(prof));
  if (fast_shape) {
#pragma omp parallel for schedule(static)
    for (int p = 0; p < np; ++p)
      std::stable_sort(dord.begin() + sw.ptr[p], dord.begin() + sw.ptr[p + 1],
                       [&](int u, int v) { return rlen_out[u] > rlen_out[v]; });
    for (int p = 0; p < np; ++p) {
      int covered = 0;   // rows are sorted by decreasing length: walk them from the shortest
      for (int q = sw.ptr[p + 1] - 1; q >= sw.ptr[p]; --q) {
        const int lane = q - sw.ptr[p], len = std::min(rlen_out[dord[q]], 32);
        for (int e = covered; e < len; ++e) prof.cnt[e] = std::max(prof.cnt[e], lane + 1);
        covered = std::max(covered, len);
      }
    }
    for (int e = 0; e < 32; ++e) prof.off[e + 1] = prof.off[e] + ((prof.cnt[e] + 3) & ~3);   // 32-byte aligned rows
  }
  d.prof = prof;
  d.vstride = prof.off[32];
  long long tot_val = 0, tot_n = 0, tot_q = 0, tot_tri = 0;
  int max_nbr = 0;
  for (int k = 0; k < np; ++k) {
    const int p = order[k];
    const int s = sw.ptr[p + 1] - sw.ptr[p];
    pat[k].q0 = (int)tot_q;
    pat[k].s = s;
    pat[k].n0 = (int)tot_n;
    pat[k].nn = nn[k];
    tot_q += s;
    tot_n += nn[k];
    tot_val += ne[k];
    tot_tri += (long long)s * (s + 1) / 2;
    max_nbr = std::max(max_nbr, nn[k]);
  }
  d.qoff.resize(np + 1);
  for (int k = 0; k < np; ++k) d.qoff[k] = pat[k].q0;
  d.qoff[np] = (int)tot_q;
  if (tot_n > 2000000000LL) throw std::runtime_error("Schwarz neighbourhood lists exceed int32 indexing");
  if (max_nbr > 65535) throw std::runtime_error("Schwarz patch neighbourhood larger than 65535 dofs");
  d.max_nbr = max_nbr;
  std::vector<int> nbr((size_t)tot_n);
  if (nb > 64) throw std::runtime_error("more than 64 parts are not supported by the Schwarz exchange lists");
  std::vector<unsigned long long>& readers = d.readers;
  std::vector<unsigned long long>& members = d.members;
  readers.assign(nb > 1 ? n : 0, 0ull);
  members.assign(nb > 1 ? n : 0, 0ull);
  // general path: host signature of what the stored inverse and the local columns depend on (size, row
  // lengths, local column of every entry, value of every entry inside the patch)
  std::vector<unsigned long long> sigA(np, 0), sigB(np, 0);
  auto mix = [](unsigned long long h, unsigned long long v, unsigned long long m) {
    h ^= v + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2);
    h *= m;
    h ^= h >> 29;
    return h;
  };
#pragma omp parallel
  {
    std::vector<int> mark(n, -1), pos(n, 0), lidx(n, -1), list;
#pragma omp for schedule(dynamic, 2048)
    for (int k = 0; k < np; ++k) {
      const int p = order[k];
      const int s = pat[k].s;
      list.clear();
      for (int q = 0; q < s; ++q) mark[iperm[sw.dofs[dord[sw.ptr[p] + q]]]] = k;   // the patch's own dofs are never gathered
      for (int q = 0; q < s; ++q) {
        const int i = iperm[sw.dofs[dord[sw.ptr[p] + q]]];
        pidx[pat[k].q0 + q] = i;
        prow[pat[k].q0 + q] = pia[i];
        plen[pat[k].q0 + q] = pia[i + 1] - pia[i];
        for (int e = pia[i]; e < pia[i + 1]; ++e)
          if (mark[pja[e]] != k) { mark[pja[e]] = k; list.push_back(pja[e]); }
      }
      std::sort(list.begin(), list.end());
      for (int j = 0; j < (int)list.size(); ++j) { pos[list[j]] = j; nbr[pat[k].n0 + j] = list[j]; }
      if (nb > 1) {   // which parts read which dof (permuted ids)
        const unsigned long long bit = 1ull << hl.part[sw.seed[p]];
        for (int j : list) {
#pragma omp atomic
          readers[j] |= bit;
        }
        for (int q = 0; q < s; ++q) {
#pragma omp atomic
          members[pidx[pat[k].q0 + q]] |= bit;
        }
      }
      if (!fast_shape) {
        for (int q = 0; q < s; ++q) lidx[pidx[pat[k].q0 + q]] = q;
        unsigned long long h1 = mix(0x452821e638d01377ULL, (unsigned long long)s, 0xff51afd7ed558ccdULL);
        unsigned long long h2 = mix(0xbe5466cf34e90c6cULL, (unsigned long long)s, 0xc4ceb9fe1a85ec53ULL);
        for (int q = 0; q < s; ++q) {
          const int i = pidx[pat[k].q0 + q];
          h1 = mix(h1, (unsigned long long)(pia[i + 1] - pia[i]), 0xff51afd7ed558ccdULL);
          for (int e = pia[i]; e < pia[i + 1]; ++e) {
            const int j = pja[e];
            const int lj = lidx[j];
            const bool inside = lj >= 0 && lj < s && pidx[pat[k].q0 + lj] == j;
            unsigned long long code = inside ? ((unsigned long long)(lj + 1) << 32) : (unsigned long long)pos[j];
            unsigned long long vb = 0;
            std::memcpy(&vb, &pa[e], 8);
            h1 = mix(mix(h1, code, 0xff51afd7ed558ccdULL), vb, 0xff51afd7ed558ccdULL);
            h2 = mix(mix(h2, vb, 0xc4ceb9fe1a85ec53ULL), code, 0xc4ceb9fe1a85ec53ULL);
          }
        }
        sigA[k] = h1;
        sigB[k] = h2;
      }
    }
  }
  const size_t tri_max = (size_t)d.max_size * (d.max_size + 1) / 2;
  d.smem_setup = (tri_max + d.max_size) * sizeof(double) + (size_t)d.max_size * sizeof(int);
  d.warps = d.max_size <= 32 ? 1 : (d.max_size <= 96 ? 2 : 4);
  const SwLayout lay = {d.max_size, d.max_nbr, d.srow};
  d.ppc = d.warps == 1 ? (lay.total_d() * 8 * 4 <= 100 * 1024 ? 4 : 2) : 1;
  d.smem_apply = (size_t)d.ppc * lay.total_d() * sizeof(double);
  if (d.smem_setup > 227 * 1024 || d.smem_apply > 227 * 1024)
    throw std::runtime_error("Schwarz patch of " + std::to_string(d.max_size) + " dofs does not fit in shared memory");
  auto up = [&](const void* src, size_t bytes) {
    void* p = alloc(bytes);
    if (bytes) cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice);
    return p;
  };
  auto sync_or_throw = [](const char* what) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + " failed: " + cudaGetErrorString(e));
  };
  if (nb > 1) {
    std::vector<int> xidx;
    d.xoff.assign(sw.ncolors * nb + 1, 0);
    for (int kb = 0; kb < sw.ncolors * nb; ++kb) {
      for (int k = d.cb_ptr[kb]; k < d.cb_ptr[kb + 1]; ++k) {
        const int p = order[k];
        const int pp = hl.part[sw.seed[p]];
        for (int q = 0; q < pat[k].s; ++q) {
          const int nat = sw.dofs[dord[sw.ptr[p] + q]], dof = pidx[pat[k].q0 + q];
          const unsigned long long need = (readers[dof] | (1ull << hl.part[nat])) & ~(1ull << pp);
          if (need) { xidx.push_back(dof); d.h_xmask.push_back(need); }
        }
      }
      d.xoff[kb + 1] = (int)xidx.size();
    }
    d.xidx = (int*)up(xidx.data(), xidx.size() * sizeof(int));
    d.h_xidx = xidx;
  }
  d.pidx = (int*)up(pidx.data(), pidx.size() * sizeof(int));
  d.nbr = (int*)up(nbr.data(), nbr.size() * sizeof(int));
  // ---- de-duplication: group the patches by signature (order of first appearance) -----------------
  const char* nodedup = getenv("MAMG_SW_DEDUP");
  const bool dedup = !(nodedup && atoi(nodedup) == 0);
  SwPatch* d_pat_tmp = nullptr;
  if (fast_shape) {
    // offsets e0 / i0 are not used by the fast path: upload the descriptors now, hash on the device
    d.pat = (SwPatch*)up(pat.data(), pat.size() * sizeof(SwPatch));
    if (dedup) {
      SwSig* d_sig = nullptr;
      if (cudaMalloc(&d_sig, sizeof(SwSig) * (size_t)np) != cudaSuccess) throw std::runtime_error("Schwarz signatures: out of device memory");
      schwarz_sig_kernel<<<(np + 7) / 8, 256>>>(np, d.pat, d.pidx, d.nbr, d_ia, d_ja, d_a, d_sig);
      sync_or_throw("Schwarz signature kernel");
      std::vector<SwSig> hs(np);
      cudaMemcpy(hs.data(), d_sig, sizeof(SwSig) * (size_t)np, cudaMemcpyDeviceToHost);
      cudaFree(d_sig);
      for (int k = 0; k < np; ++k) { sigA[k] = hs[k].a; sigB[k] = hs[k].b; }
    }
  }
  (void)d_pat_tmp;
  std::vector<int> uid(np), rep;
  if (dedup) {
    struct Key { unsigned long long a, b; bool operator==(const Key& o) const { return a == o.a && b == o.b; } };
    struct KeyHash { size_t operator()(const Key& k) const { return (size_t)(k.a ^ (k.b * 0x9e3779b97f4a7c15ULL)); } };
    std::unordered_map<Key, int, KeyHash> seen;
    seen.reserve(1 << 16);
    for (int k = 0; k < np; ++k) {
      auto it = seen.find(Key{sigA[k], sigB[k]});
      if (it == seen.end()) { seen.emplace(Key{sigA[k], sigB[k]}, (int)rep.size()); uid[k] = (int)rep.size(); rep.push_back(k); }
      else uid[k] = it->second;
    }
  } else {
    rep.resize(np);
    std::iota(rep.begin(), rep.end(), 0);
    uid = rep;
  }
  const int nu = (int)rep.size();
  d.nuniq = nu;
  // packed inverses (s(s+1)/2 doubles, 16-byte aligned) and, on the general path, the local columns of
  // the unique blobs
  std::vector<long long> inv_off(nu + 1, 0), e_off(nu + 1, 0);
  for (int u = 0; u < nu; ++u) {
    const long long s = pat[rep[u]].s;
    inv_off[u + 1] = inv_off[u] + (s * (s + 1) / 2 + 1) / 2 * 2;
    e_off[u + 1] = e_off[u] + (s * srow + 7) / 8 * 8;                 // 16-byte aligned uint16 segments
  }
  const long long tot_i = inv_off[nu], tot_e = fast_shape ? 0 : e_off[nu];
  long long uniq_blob_bytes = 8 * tot_i;
  if (!fast_shape) {
    for (int k = 0; k < np; ++k) { pat[k].i0 = inv_off[uid[k]]; pat[k].e0 = e_off[uid[k]]; }
    std::vector<uint16_t> lcol((size_t)tot_e + 8, 0);
#pragma omp parallel
    {
      std::vector<int> pos(n, 0);
#pragma omp for schedule(dynamic, 256)
      for (int u = 0; u < nu; ++u) {
        const int k = rep[u];
        for (int q = 0; q < pat[k].s; ++q) pos[pidx[pat[k].q0 + q]] = pat[k].nn;   // inside the patch: the zero slot
        for (int j = 0; j < pat[k].nn; ++j) pos[nbr[pat[k].n0 + j]] = j;
        uint16_t* out = &lcol[(size_t)e_off[u]];
        for (int q = 0; q < pat[k].s; ++q) {
          const int i = pidx[pat[k].q0 + q];
          for (int e = pia[i]; e < pia[i + 1]; ++e) out[(size_t)q * srow + (e - pia[i])] = (uint16_t)pos[pja[e]];
        }
      }
    }
    d.pat = (SwPatch*)up(pat.data(), pat.size() * sizeof(SwPatch));
    d.prow = (int*)up(prow.data(), prow.size() * sizeof(int));
    d.plen = (int*)up(plen.data(), plen.size() * sizeof(int));
    d.lcol = (uint16_t*)up(lcol.data(), lcol.size() * sizeof(uint16_t));
    uniq_blob_bytes += 2 * tot_e;
  }
  int* d_rep = (int*)up(rep.data(), rep.size() * sizeof(int));
  d.inv_off = (long long*)up(inv_off.data(), (size_t)nu * sizeof(long long));
  d.pinv = (double*)alloc(((size_t)tot_i + 2) * sizeof(double));
  cudaMemset(d.pinv, 0, ((size_t)tot_i + 2) * sizeof(double));   // padding must be 0, not NaN bits
  constexpr int TS = 128;
  if (d.smem_setup > 48 * 1024)
    cudaFuncSetAttribute(schwarz_invert_kernel<TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d.smem_setup);
  schwarz_invert_kernel<TS><<<nu, TS, d.smem_setup>>>(nu, d_rep, d.inv_off, d.pat, d.pidx, d_ia, d_ja, d_a, d.pinv, d.max_size);
  sync_or_throw("Schwarz setup kernel");
  if (fast_shape) {
    d.fast = true;
    d.nbq = (max_nn + 31) / 32;
    d.sq = (srow + 3) / 4;
    d.sr_t = (srow <= 12 && d.nbq <= 1) ? 12 : ((srow <= 24 && d.nbq <= 4) ? 24 : 32);
    d.uid = (int*)up(uid.data(), uid.size() * sizeof(int));
    d.pidx32 = (int*)alloc((size_t)np * 32 * sizeof(int));
    d.nbrp = (int*)alloc((size_t)np * d.nbq * 32 * sizeof(int));
    d.vt = (double*)alloc(((size_t)nu * d.vstride + 32) * sizeof(double));
    d.ct4 = (uint32_t*)alloc((size_t)nu * d.sq * 32 * sizeof(uint32_t));
    schwarz_index_kernel<<<(np + 7) / 8, 256>>>(np, d.pat, d.pidx, d.nbr, d.nbq, d.pidx32, d.nbrp);
    schwarz_blob_kernel<<<(nu + 7) / 8, 256>>>(nu, d_rep, d.pat, d.pidx, d.nbr, d_ia, d_ja, d_a, srow, d.sq,
                                              d.vt, d.ct4, d.prof, d.vstride);
    sync_or_throw("Schwarz blob kernel");
    uniq_blob_bytes += (long long)nu * (8LL * d.vstride + 4LL * d.sq * 32);
    const int fsm = kSwFastWarps * kSwFastSlot * (int)sizeof(double);
    cudaFuncSetAttribute(schwarz_fast_kernel<12, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, fsm);
    cudaFuncSetAttribute(schwarz_fast_kernel<24, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, fsm);
    cudaFuncSetAttribute(schwarz_fast_kernel<32, 8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, fsm);
    cudaFuncSetAttribute(schwarz_fast_kernel<12, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, fsm);
    cudaFuncSetAttribute(schwarz_fast_kernel<24, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, fsm);
    cudaFuncSetAttribute(schwarz_fast_kernel<32, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, fsm);
  }
  if (d.smem_apply > 48 * 1024) {
    // (the opt-in value is per function: keep the largest request of all hierarchies of this process)
    static size_t opted = 0;
    opted = std::max(opted, d.smem_apply);
    cudaFuncSetAttribute(schwarz_apply_kernel<1, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)opted);
    cudaFuncSetAttribute(schwarz_apply_kernel<1, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)opted);
    cudaFuncSetAttribute(schwarz_apply_kernel<2, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)opted);
    cudaFuncSetAttribute(schwarz_apply_kernel<4, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)opted);
    cudaFuncSetAttribute(schwarz_apply_kernel<1, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)opted);
    cudaFuncSetAttribute(schwarz_apply_kernel<1, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)opted);
    cudaFuncSetAttribute(schwarz_apply_kernel<2, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)opted);
    cudaFuncSetAttribute(schwarz_apply_kernel<4, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)opted);
  }
  // Algorithmic bytes of one sweep.
  //  stored-factor model (every patch owns its data): values + local columns of the row entries,
  //  neighbour list + gathered x, packed inverse, (idx, row start, row length, b, x update) per patch dof
  d.alg_bytes_stored = 10 * tot_val + 12 * tot_n + 8 * ((tot_tri + 1) / 2 * 2) + (12 + 8 + 16) * tot_q + 32LL * np;
  if (fast_shape) d.alg_bytes_stored = 9 * tot_val + 12 * tot_n + 8 * tot_tri + (4 + 8 + 8) * tot_q;
  //  with shared blobs: the per-patch index data, the gathered x, b and the x update stay per patch; the
  //  blobs are counted once per colour launch (they are what a launch has to bring in at least once);
  //  the general path still streams the row values of every patch from the level CSR
  d.alg_bytes = 12 * tot_n + (fast_shape ? (4 + 8 + 8) : (12 + 8 + 16)) * tot_q + (fast_shape ? 4LL : 32LL) * np +
                (fast_shape ? 0 : 8 * tot_val) + (long long)sw.ncolors * uniq_blob_bytes;
  if (nu == np) d.alg_bytes = d.alg_bytes_stored;
}

// patches [p0, p1) of one conflict colour (they commute, so they run concurrently)
inline void schwarz_range_launch(const DSchwarz& d, int p0, int p1, const double* a, const double* b, double* x,
                                 cudaStream_t stream, const HaloTail& tail) {
  if (d.fast) {
    const int g = (p1 - p0 + kSwFastWarps - 1) / kSwFastWarps;
    const size_t sm = (size_t)kSwFastWarps * kSwFastSlot * sizeof(double);
#define MAMG_SWF_ARGS p0, p1, d.pidx32, d.nbrp, d.uid, d.inv_off, d.vt, d.ct4, d.pinv, b, x, d.srow, d.sq, d.nbq, d.max_size, d.prof, d.vstride, tail
#define MAMG_SWF_LAUNCH(SR, NBQ)                                                                         \
    do {                                                                                                 \
      if (tail.nn > 0) schwarz_fast_kernel<SR, NBQ, true><<<g, kSwFastWarps * 32, sm, stream>>>(MAMG_SWF_ARGS); \
      else schwarz_fast_kernel<SR, NBQ, false><<<g, kSwFastWarps * 32, sm, stream>>>(MAMG_SWF_ARGS);      \
    } while (0)
    if (d.sr_t == 12) MAMG_SWF_LAUNCH(12, 1);
    else if (d.sr_t == 24) MAMG_SWF_LAUNCH(24, 4);
    else MAMG_SWF_LAUNCH(32, 8);
#undef MAMG_SWF_LAUNCH
#undef MAMG_SWF_ARGS
    return;
  }
  const int grid = (p1 - p0 + d.ppc - 1) / d.ppc;
  const SwLayout lay = {d.max_size, d.max_nbr, d.srow};
#define MAMG_SW_ARGS p0, p1, d.pat, d.pidx, d.prow, d.plen, d.nbr, d.lcol, d.pinv, a, b, x, lay, tail
#define MAMG_SW_LAUNCH(WPP, PPC, THREADS)                                                                  \
  do {                                                                                                     \
    if (tail.nn > 0) schwarz_apply_kernel<WPP, PPC, true><<<grid, THREADS, d.smem_apply, stream>>>(MAMG_SW_ARGS); \
    else schwarz_apply_kernel<WPP, PPC, false><<<grid, THREADS, d.smem_apply, stream>>>(MAMG_SW_ARGS);      \
  } while (0)
  if (d.warps == 1 && d.ppc == 4) MAMG_SW_LAUNCH(1, 4, 128);
  else if (d.warps == 1) MAMG_SW_LAUNCH(1, 2, 64);
  else if (d.warps == 2) MAMG_SW_LAUNCH(2, 1, 64);
  else MAMG_SW_LAUNCH(4, 1, 128);
#undef MAMG_SW_LAUNCH
#undef MAMG_SW_ARGS
}

}  // namespace mamg

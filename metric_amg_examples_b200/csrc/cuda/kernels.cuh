// Hand-written sm_100a kernels of the metric-AMG cycle and the Krylov loop (fp64, int32).
//
// Every kernel here is HBM-bound (SpMV: 2 flop per 12 B), so the rules that matter are
// coalescing, enough loads in flight and no re-reads; tensor cores are not used because no
// step is a dense contraction (BASELINE.json north_star).
//
// Matrix layout: CSR whose rows are permuted so that every Gauss-Seidel colour is one
// contiguous row range (stable inside a colour, so the banded x-locality of the
// lexicographic mesh ordering survives inside each colour block).  A sub-warp of LANES
// threads owns one row: its val/col loads are contiguous and the x-gathers of neighbouring
// rows fall in the same cache lines.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

namespace mamg {

constexpr int kBlock = 256;

// Sum over the LANES threads of one sub-warp (all of them hold the same row, so the whole
// sub-warp is converged here even when neighbouring sub-warps have exited); lane 0 gets the sum.
template <int LANES>
__device__ __forceinline__ double subwarp_sum(double v) {
  const unsigned int full = LANES == 32 ? 0xffffffffu : ((1u << LANES) - 1u);
  const unsigned int mask = full << ((threadIdx.x & 31) / LANES * LANES);
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_down_sync(mask, v, o, LANES);
  return v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// Loads of data that is read exactly once per kernel (matrix values / columns, patch index data) go
// through the read-only path.  L2 evict-first / no-L1-allocate cache hints were measured in round 1 and
// were slower than plain loads on B200 (the 126 MB L2 keeps the re-read vectors anyway); they are gone.
__device__ __forceinline__ double ld_stream(const double* p) { return __ldg(p); }
__device__ __forceinline__ int ld_stream(const int* p) { return __ldg(p); }
__device__ __forceinline__ unsigned int ld_stream(const unsigned int* p) { return __ldg(p); }

// ---- halo push folded into the producing kernel (multi-GPU halo mode) -----------------------------------
// A smoother kernel that is followed by a halo exchange carries this descriptor: when its last block
// retires (ticket), that block stores the listed boundary entries of the vector into the neighbour
// ranks' copies, raises this rank's flag there and waits for theirs -- the exchange costs no launch of
// its own.  nn = 0: no exchange.  See halo_push_kernel (device.cu) for the protocol.
// clock cycles a rank spins on a peer's flag before it gives up with a trap (a peer that never arrives would
// otherwise hang the GPU).  MAMG_PEER_TIMEOUT_S (default 60 s at ~2 GHz) sets it at mamg_dist_init; benign skew
// between ranks -- a rank that enters a solve seconds later because of host-side work -- stays far below it.
__device__ long long g_spin_limit = 120000000000LL;

struct HaloTail {
  int nn;
  int peer[8], beg[8], cnt[8];
  const int* send;
  long long voff;
  double* const* peers;
  int me;
  unsigned int* ticket;
  long long* phase_ctr;
};
__device__ __forceinline__ void halo_tail(const HaloTail& T) {
  if (T.nn == 0) return;   // kernel-uniform
  __shared__ bool ht_last;
  __shared__ long long ht_phase;
  __threadfence();          // this block's updates are visible device-wide before it takes a ticket
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int nblk = gridDim.x;
    ht_last = atomicInc(T.ticket, nblk - 1) == nblk - 1;
  }
  __syncthreads();
  if (!ht_last) return;
  __threadfence();
  const double* mine = T.peers[T.me] + T.voff;
  for (int k = 0; k < T.nn; ++k) {
    double* dst = T.peers[T.peer[k]] + T.voff;
    const int* idx = T.send + T.beg[k];
    for (int t0 = threadIdx.x; t0 < T.cnt[k]; t0 += 4 * blockDim.x) {   // 4 entries in flight per thread
      int i[4];
      double v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { const int t = t0 + u * blockDim.x; i[u] = t < T.cnt[k] ? idx[t] : -1; }
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = i[u] >= 0 ? __ldcg(mine + i[u]) : 0.0;   // written by other blocks: read past this SM's L1
#pragma unroll
      for (int u = 0; u < 4; ++u) if (i[u] >= 0) dst[i[u]] = v[u];
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    ht_phase = *reinterpret_cast<volatile long long*>(T.phase_ctr) + 1;
    for (int k = 0; k < T.nn; ++k) reinterpret_cast<volatile long long*>(T.peers[T.peer[k]])[T.me] = ht_phase;
  }
  __syncthreads();
  if ((int)threadIdx.x < T.nn) {
    const volatile long long* flag = reinterpret_cast<const volatile long long*>(T.peers[T.me]) + T.peer[threadIdx.x];
    const long long t0 = clock64();
    while (*flag < ht_phase) {
      if (clock64() - t0 > g_spin_limit) { printf("mamg: rank %d: neighbour %d never reached exchange %lld\n", T.me, T.peer[threadIdx.x], ht_phase); __trap(); }
    }
    __threadfence_system();
  }
  __syncthreads();
  if (threadIdx.x == 0) *T.phase_ctr = ht_phase;
}

// Row dot product a_i . x by one sub-warp; the full sum is valid in lane 0 of the sub-warp.
template <int LANES>
__device__ __forceinline__ double row_dot(const int* __restrict__ ja, const double* __restrict__ a,
                                          const double* x, int p0, int p1, int lane) {
  double s = 0.0;
  for (int p = p0 + lane; p < p1; p += LANES) s += ld_stream(a + p) * x[ld_stream(ja + p)];
  return subwarp_sum<LANES>(s);
}

// U adjacent rows [row0, row0+U) (clipped at rend) by one sub-warp, all U row streams in flight
// at once: every lane has U independent (value, column, x) load chains outstanding, which is
// what hides the DRAM latency of these 12-byte-per-flop kernels.  Sums valid in lane 0.
template <int LANES, int U>
__device__ __forceinline__ void rows_dot(const int* __restrict__ ia, const int* __restrict__ ja,
                                         const double* __restrict__ a, const double* x, int row0,
                                         int rend, int lane, double (&s)[U]) {
  int p[U], e[U];
  const int psafe = ia[row0];   // a valid entry index for lanes that have run out of work
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int r = row0 + u;
    p[u] = r < rend ? ia[r] + lane : 0;
    e[u] = r < rend ? ia[r + 1] : 0;
    s[u] = 0.0;
  }
  bool more = true;
  while (more) {
    more = false;
    // three unconditional load waves (columns, values, x): no branch sits between the U streams,
    // so all of them are in flight together; finished lanes re-read a valid entry and add 0
    int cj[U];
    double av[U], xv[U];
    bool on[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      on[u] = p[u] < e[u];
      const int q = on[u] ? p[u] : psafe;
      cj[u] = ld_stream(ja + q);
      av[u] = ld_stream(a + q);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) xv[u] = x[cj[u]];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      s[u] += on[u] ? av[u] * xv[u] : 0.0;
      p[u] += LANES;
      more |= p[u] < e[u];
    }
  }
#pragma unroll
  for (int u = 0; u < U; ++u) s[u] = subwarp_sum<LANES>(s[u]);
}

// ---------------------------------------------------------------------------------------
// K1  y = A x   |   y = b - A x          (HAZmath dcsr_mxv / dcsr_aAxpy; PETSc MatMult)
// ---------------------------------------------------------------------------------------
template <int LANES, int U, bool RESID>
__global__ void __launch_bounds__(kBlock)
spmv_kernel(int rbeg, int n, const int* __restrict__ ia, const int* __restrict__ ja,
            const double* __restrict__ a, const double* __restrict__ x,
            const double* __restrict__ b, double* __restrict__ y) {
  // rows [rbeg, n): the whole level, or the rows this rank owns on a row-distributed level
  const int lane = threadIdx.x % LANES;
  const int row0 = rbeg + ((blockIdx.x * kBlock + threadIdx.x) / LANES) * U;
  if (row0 >= n) return;
  double s[U];
  rows_dot<LANES, U>(ia, ja, a, x, row0, n, lane, s);
  if (lane == 0) {
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (row0 + u < n) y[row0 + u] = RESID ? b[row0 + u] - s[u] : s[u];
  }
}

// Deterministic two-stage reduction: every block writes its partial sums, the block that
// arrives last adds them in index order (fixed order => run-to-run and rank-count
// independent rounding) and stores the NV results in out[0..NV).
template <int NV>
__device__ __forceinline__ bool block_reduce_finish(double (&v)[NV], double* partial,
                                                    unsigned int* ticket, double* out) {
  __shared__ double sm[NV][kBlock / 32];
  __shared__ bool is_last;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double w = warp_sum(v[k]);
    if (lane == 0) sm[k][warp] = w;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double s = 0.0;
      for (int w = 0; w < kBlock / 32; ++w) s += sm[k][w];
      partial[(size_t)k * gridDim.x + blockIdx.x] = s;
    }
    __threadfence();
    unsigned int t = atomicInc(ticket, gridDim.x - 1);  // wraps to 0 after the last block
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return false;
  __threadfence();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double s = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += kBlock)
      s += ((volatile double*)partial)[(size_t)k * gridDim.x + i];
    // fixed-shape tree over the block: identical for every run
    double w = warp_sum(s);
    __syncthreads();
    if (lane == 0) sm[k][warp] = w;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int q = 0; q < kBlock / 32; ++q) t += sm[k][q];
      out[k] = t;
    }
  }
  return true;  // thread 0 of this block has written out[0..NV) and may post-process it
}

// PCG scalars live on the device: sc[0]=rz, sc[1]=d.q, sc[2]=alpha, sc[3]=rz_new, sc[4]=beta
// q = A d  with the dot product d.q in the epilogue; the finishing block sets
// sc[1] = d.q and alpha = sc[2] = rz / d.q                       (K1 + K10 fused)
template <int LANES, int U>
__global__ void __launch_bounds__(kBlock)
spmv_dot_kernel(int n, const int* __restrict__ ia, const int* __restrict__ ja,
                const double* __restrict__ a, const double* __restrict__ d, double* __restrict__ q,
                double* partial, unsigned int* ticket, double* sc) {
  const int lane = threadIdx.x % LANES;
  const int nsub = gridDim.x * (kBlock / LANES);  // fixed grid, row groups strided: few tickets, fixed partial count
  double v[1] = {0.0};
  for (int row0 = ((blockIdx.x * kBlock + threadIdx.x) / LANES) * U; row0 < n; row0 += nsub * U) {
    double s[U];
    rows_dot<LANES, U>(ia, ja, a, d, row0, n, lane, s);
    if (lane == 0) {
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (row0 + u < n) { q[row0 + u] = s[u]; v[0] += s[u] * d[row0 + u]; }
    }
  }
  if (block_reduce_finish<1>(v, partial, ticket, sc + 1) && threadIdx.x == 0) sc[2] = sc[0] / sc[1];
}

// ---------------------------------------------------------------------------------------
// K6  one colour of a Gauss-Seidel / SOR sweep on rows [r0, r1)   (HAZmath smoother_dcsr_gs /
//     _sgs / _sor in natural order; here in the multicolour order the oracle shares)
//       x_i <- x_i + w (b_i - sum_j a_ij x_j) / a_ii
//     Rows of one colour do not couple, so the in-place update equals the sequential sweep.
// ---------------------------------------------------------------------------------------
template <int LANES, int U>
__global__ void __launch_bounds__(kBlock)
gs_color_kernel(int r0, int r1, const int* __restrict__ ia, const int* __restrict__ ja,
                const double* __restrict__ a, const double* __restrict__ invd,
                const uint8_t* __restrict__ skip, const double* __restrict__ b, double* x,
                double omega) {
  const int lane = threadIdx.x % LANES;
  const int row0 = r0 + ((blockIdx.x * kBlock + threadIdx.x) / LANES) * U;
  if (row0 >= r1) return;
  // rows smoothed by Schwarz instead (skip mask) come in long runs inside a colour: test the group
  int rend = r1;
  if (skip != nullptr) {
    bool all = true;
#pragma unroll
    for (int u = 0; u < U; ++u) all &= (row0 + u >= r1) || skip[row0 + u];
    if (all) return;
  }
  double s[U];
  rows_dot<LANES, U>(ia, ja, a, x, row0, rend, lane, s);
  if (lane == 0) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int row = row0 + u;
      if (row < r1 && !(skip != nullptr && skip[row])) x[row] += omega * (b[row] - s[u]) * invd[row];
    }
  }
}

// damped Jacobi, out of place: xn = x + w D^-1 (b - A x)
template <int LANES>
__global__ void __launch_bounds__(kBlock)
jacobi_kernel(int rbeg, int n, const int* __restrict__ ia, const int* __restrict__ ja,
              const double* __restrict__ a, const double* __restrict__ invd,
              const uint8_t* __restrict__ skip, const double* __restrict__ b,
              const double* __restrict__ x, double* __restrict__ xn, double omega) {
  const int lane = threadIdx.x % LANES;
  const int row = rbeg + (blockIdx.x * kBlock + threadIdx.x) / LANES;
  if (row >= n) return;
  double s = row_dot<LANES>(ja, a, x, ia[row], ia[row + 1], lane);
  if (lane == 0) {
    bool sk = skip != nullptr && skip[row];
    xn[row] = sk ? x[row] : x[row] + omega * (b[row] - s) * invd[row];
  }
}

// ---------------------------------------------------------------------------------------
// K3  fused residual + unsmoothed-aggregation restriction (HAZmath dcsr_aAxpy + dcsr_mxv_agg):
//       bc[I] = sum_{i in aggregate I} (b_i - a_i . x),   xc[I] = 0
//     One sub-warp per coarse row walks the fine rows of its aggregate, so every fine row is
//     read exactly once and the fine residual is never stored.
// ---------------------------------------------------------------------------------------
template <int LANES>
__global__ void __launch_bounds__(kBlock)
resid_restrict_kernel(int cbeg, int nc, const int* __restrict__ cptr, const int* __restrict__ cidx,
                      const int* __restrict__ ia, const int* __restrict__ ja,
                      const double* __restrict__ a, const double* __restrict__ x,
                      const double* __restrict__ b, double* __restrict__ bc,
                      double* __restrict__ xc) {
  const int lane = threadIdx.x % LANES;
  const int I = cbeg + (blockIdx.x * kBlock + threadIdx.x) / LANES;
  if (I >= nc) return;
  double acc = 0.0;
  const int q1 = cptr[I + 1];
  for (int q = cptr[I]; q < q1; q += 2) {   // two member rows in flight (pairwise aggregates: one pass)
    const int i0 = cidx[q];
    const bool two = q + 1 < q1;
    const int i1 = two ? cidx[q + 1] : i0;
    int p0 = ia[i0] + lane, e0 = ia[i0 + 1];
    int p1 = two ? ia[i1] + lane : 0, e1 = two ? ia[i1 + 1] : 0;
    double s0 = 0.0, s1 = 0.0;
    const int psafe = ia[i0];
    while (p0 < e0 || p1 < e1) {
      const bool on0 = p0 < e0, on1 = p1 < e1;
      const int q0 = on0 ? p0 : psafe, q1 = on1 ? p1 : psafe;
      const int c0 = ld_stream(ja + q0), c1 = ld_stream(ja + q1);
      const double a0 = ld_stream(a + q0), a1 = ld_stream(a + q1);
      const double x0 = x[c0], x1 = x[c1];
      s0 += on0 ? a0 * x0 : 0.0;
      s1 += on1 ? a1 * x1 : 0.0;
      p0 += LANES;
      p1 += LANES;
    }
    s0 = subwarp_sum<LANES>(s0);
    s1 = subwarp_sum<LANES>(s1);
    acc += b[i0] - s0;
    if (two) acc += b[i1] - s1;
  }
  if (lane == 0) { bc[I] = acc; xc[I] = 0.0; }
}

// ---------------------------------------------------------------------------------------
// coarse scaling (src/amg_parameters.py:58 "coarse_scaling": ON):
//   alpha = (e . r) / (e . A e), clipped to <= 1 as in the FASP-lineage cycle; NaN -> 1.
//   out[0] = e.r, out[1] = e.Ae, out[2] = alpha
// ---------------------------------------------------------------------------------------
template <int LANES>
__global__ void __launch_bounds__(kBlock)
scale_dots_kernel(int n, const int* __restrict__ ia, const int* __restrict__ ja,
                  const double* __restrict__ a, const double* __restrict__ e,
                  const double* __restrict__ r, double* partial, unsigned int* ticket, double* out) {
  const int lane = threadIdx.x % LANES;
  const int nsub = gridDim.x * (kBlock / LANES);
  double v[2] = {0.0, 0.0};
  for (int row = (blockIdx.x * kBlock + threadIdx.x) / LANES; row < n; row += nsub) {
    double s = row_dot<LANES>(ja, a, e, ia[row], ia[row + 1], lane);
    if (lane == 0) { double ei = e[row]; v[0] += ei * r[row]; v[1] += ei * s; }
  }
  if (block_reduce_finish<2>(v, partial, ticket, out) && threadIdx.x == 0) {
    double al = out[0] / out[1];
    out[2] = (al < 1.0) ? al : 1.0;
  }
}

// distributed variant of the coarse scaling: t = A_c e was all-gathered; out[0] = e.r, out[1] = e.t, out[2] = alpha
__global__ void __launch_bounds__(kBlock)
scale_dots_vec_kernel(int n, const double* __restrict__ e, const double* __restrict__ r,
                      const double* __restrict__ t, double* partial, unsigned int* ticket, double* out) {
  double v[2] = {0.0, 0.0};
  for (int i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) {
    const double ei = e[i];
    v[0] += ei * r[i];
    v[1] += ei * t[i];
  }
  if (block_reduce_finish<2>(v, partial, ticket, out) && threadIdx.x == 0) {
    double al = out[0] / out[1];
    out[2] = (al < 1.0) ? al : 1.0;
  }
}
// sc[1] = d.q, sc[2] = alpha = rz / d.q      (distributed PCG: q was all-gathered)
__global__ void __launch_bounds__(kBlock)
pcg_dq_kernel(int n, const double* __restrict__ d, const double* __restrict__ q, double* partial,
              unsigned int* ticket, double* sc) {
  double v[1] = {0.0};
  for (int i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) v[0] += d[i] * q[i];
  if (block_reduce_finish<1>(v, partial, ticket, sc + 1) && threadIdx.x == 0) sc[2] = sc[0] / sc[1];
}
// buf[q] = x[idx[q0 + q]]  /  x[idx[q0 + q]] = buf[q]   (Schwarz patch-dof exchange)
__global__ void __launch_bounds__(kBlock)
pack_kernel(int cnt, const int* __restrict__ idx, const double* __restrict__ x, double* __restrict__ buf) {
  const int q = blockIdx.x * kBlock + threadIdx.x;
  if (q < cnt) buf[q] = x[idx[q]];
}
__global__ void __launch_bounds__(kBlock)
unpack_kernel(int cnt, const int* __restrict__ idx, const double* __restrict__ buf, double* __restrict__ x) {
  const int q = blockIdx.x * kBlock + threadIdx.x;
  if (q < cnt) x[idx[q]] = buf[q];
}

// K4  x_i += alpha * e[agg(i)]      (HAZmath dcsr_aAxpy_agg)
__global__ void __launch_bounds__(kBlock)
prolong_kernel(int rbeg, int n, const int* __restrict__ agg, const double* __restrict__ e,
               const double* __restrict__ alpha, double* __restrict__ x) {
  const int i = rbeg + blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  const int I = agg[i];
  if (I >= 0) x[i] += (alpha ? *alpha : 1.0) * e[I];
}

// K5  general CSR transfer operators of SA_AMG (smoothed prolongator P, restriction R = P'):
//     y = M x   or   y += alpha (M x)   (HAZmath dcsr_mxv / dcsr_aAxpy with P, R)
template <int LANES, bool ADD>
__global__ void __launch_bounds__(kBlock)
csr_apply_kernel(int rbeg, int n, const int* __restrict__ ia, const int* __restrict__ ja,
                 const double* __restrict__ a, const double* __restrict__ x,
                 const double* __restrict__ alpha, double* __restrict__ y, double* __restrict__ zero) {
  const int lane = threadIdx.x % LANES;
  const int row = rbeg + (blockIdx.x * kBlock + threadIdx.x) / LANES;
  if (row >= n) return;
  double s = row_dot<LANES>(ja, a, x, ia[row], ia[row + 1], lane);
  if (lane == 0) {
    if (ADD) y[row] += (alpha ? *alpha : 1.0) * s;
    else { y[row] = s; if (zero) zero[row] = 0.0; }
  }
}

// K8  coarsest solve x = Ainv b with the precomputed dense inverse, one warp per row
__global__ void __launch_bounds__(kBlock)
dense_gemv_kernel(int n, const double* __restrict__ M, const double* __restrict__ b,
                  double* __restrict__ x) {
  const int row = (blockIdx.x * kBlock + threadIdx.x) / 32, lane = threadIdx.x % 32;
  if (row >= n) return;
  double s = 0.0;
  for (int j = lane; j < n; j += 32) s += M[(size_t)row * n + j] * b[j];
  s = warp_sum(s);
  if (lane == 0) x[row] = s;
}

// ---------------------------------------------------------------------------------------
// vector kernels
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) fill_kernel(int n, double* x, double v) {
  const int i = blockIdx.x * kBlock + threadIdx.x;
  if (i < n) x[i] = v;
}
// out[i] = in[map[i]]
__global__ void __launch_bounds__(kBlock)
gather_kernel(int n, const int* __restrict__ map, const double* __restrict__ in, double* __restrict__ out) {
  const int i = blockIdx.x * kBlock + threadIdx.x;
  if (i < n) out[i] = in[map[i]];
}
__global__ void __launch_bounds__(kBlock)
copy_kernel(int n, const double* __restrict__ in, double* __restrict__ out) {
  const int i = blockIdx.x * kBlock + threadIdx.x;
  if (i < n) out[i] = in[i];
}

// block_vec <-> monolithic vector on the device (xii.ReductionOperator / ii_convert of vectors,
// src/utils.py:48-53): the blocks stay where the caller has them, the boundary gather / scatter
// finds the block of a monolithic index from the offsets -- no concatenated copy is ever made
struct BlockPtrs {
  int nb;
  int off[9];           // off[b] .. off[b+1]: monolithic range of block b
  double* p[8];
};
__device__ __forceinline__ int block_of(const BlockPtrs& B, int j) {
  int bk = 0;
#pragma unroll
  for (int q = 1; q < 8; ++q) bk += (q < B.nb && j >= B.off[q]) ? 1 : 0;
  return bk;
}
// out[i] = blocks[map[i]]   (map = perm: permuted <- natural)
__global__ void __launch_bounds__(kBlock)
gather_blocks_kernel(int n, const int* __restrict__ map, const BlockPtrs B, double* __restrict__ out) {
  const int i = blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  const int j = map[i];
  const int bk = block_of(B, j);
  out[i] = B.p[bk][j - B.off[bk]];
}
// blocks[j] = in[map[j]]    (map = iperm: natural -> permuted)
__global__ void __launch_bounds__(kBlock)
scatter_blocks_kernel(int n, const int* __restrict__ map, const double* __restrict__ in, const BlockPtrs B) {
  const int j = blockIdx.x * kBlock + threadIdx.x;
  if (j >= n) return;
  const int bk = block_of(B, j);
  B.p[bk][j - B.off[bk]] = in[map[j]];
}

// out[0] = u.v     (K10 dots, fixed-order reduction)
__global__ void __launch_bounds__(kBlock)
dot_kernel(int n, const double* __restrict__ u, const double* __restrict__ v, double* partial,
           unsigned int* ticket, double* out) {
  double acc[1] = {0.0};
  for (int i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) acc[0] += u[i] * v[i];
  block_reduce_finish<1>(acc, partial, ticket, out);
}
// rz_new = r.z; the finishing block sets sc[3] = rz_new, beta = sc[4] = rz_new / rz, sc[0] = rz_new
// (first == 1: only sc[0] = r.z, the initial residual)
__global__ void __launch_bounds__(kBlock)
pcg_rz_kernel(int n, const double* __restrict__ r, const double* __restrict__ z, double* partial,
              unsigned int* ticket, double* sc, int first) {
  double acc[2] = {0.0, 0.0};   // r.z and r.r (the latter for HAZmath's ||r||/||r0|| stopping rule)
  for (int i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) {
    const double ri = r[i];
    acc[0] += ri * z[i];
    acc[1] += ri * ri;
  }
  if (block_reduce_finish<2>(acc, partial, ticket, sc + 5) && threadIdx.x == 0) {
    sc[3] = sc[5];
    if (!first) sc[4] = sc[3] / sc[0];
    sc[0] = sc[3];
  }
}

// x += alpha d ; r -= alpha q
__global__ void __launch_bounds__(kBlock)
pcg_update_kernel(int n, const double* __restrict__ sc, const double* __restrict__ d,
                  const double* __restrict__ q, double* __restrict__ x, double* __restrict__ r) {
  const int i = blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  const double al = sc[2];
  x[i] += al * d[i];
  r[i] -= al * q[i];
}
// d = z + beta d
__global__ void __launch_bounds__(kBlock)
pcg_dir_kernel(int n, const double* __restrict__ sc, const double* __restrict__ z, double* __restrict__ d) {
  const int i = blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  d[i] = z[i] + sc[4] * d[i];
}
// y = a x + b y (host scalars), used by MINRES/GMRES
__global__ void __launch_bounds__(kBlock)
axpby_kernel(int n, double a, const double* __restrict__ x, double b, double* __restrict__ y) {
  const int i = blockIdx.x * kBlock + threadIdx.x;
  if (i < n) y[i] = a * x[i] + b * y[i];
}

// ---- MINRES / GMRES vector kernels with device-resident coefficients (no host round trip per dot) ------
// y += sign * coef[0] * scale * x
__global__ void __launch_bounds__(kBlock)
axpy_dev_kernel(int n, const double* __restrict__ coef, double scale, const double* __restrict__ x, double* __restrict__ y) {
  const int i = blockIdx.x * kBlock + threadIdx.x;
  if (i < n) y[i] += coef[0] * scale * x[i];
}
// MINRES direction and iterate in one pass: t = (v - oldeps*w1 - delta*w2) / gamma ; x += phi * t   (t becomes the new w)
__global__ void __launch_bounds__(kBlock)
minres_update_kernel(int n, double inv_gamma, double oldeps, double delta, double phi, const double* __restrict__ v,
                     const double* __restrict__ w1, const double* __restrict__ w2, double* __restrict__ t,
                     double* __restrict__ x) {
  const int i = blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  const double ti = ((v[i] * inv_gamma - (oldeps * inv_gamma) * w1[i]) - (delta * inv_gamma) * w2[i]);
  t[i] = ti;
  x[i] += phi * ti;
}

// ---- K-cycle scalars (nonlinear AMLI): s[0]=|b|^2 s[1]=rho1 s[2]=alpha1 s[3]=beta1 s[4]=|r~|^2 s[5]=gamma s[6]=alpha2
// s[7]=rho2 s[8]=beta3 s[9]=beta4.  One thread; keeps every coefficient of the two-step Krylov update on the device.
__global__ void kcycle_step1_kernel(double* s) {
  if (threadIdx.x == 0 && blockIdx.x == 0) s[3] = s[1] != 0.0 ? s[2] / s[1] : 0.0;
}
// x = beta3 c1 + beta4 c2; beta4 = 0 (and beta3 = beta1) when the first step already met the tolerance
__global__ void kcycle_step2_kernel(double* s, double tol2) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double nb2 = s[0], rho1 = s[1], alpha1 = s[2], beta1 = s[3], nr2 = s[4], gamma = s[5], alpha2 = s[6], rho2 = s[7];
  double beta3 = beta1, beta4 = 0.0;
  if (!(nr2 < tol2 * nb2) && nb2 != 0.0) {
    const double beta2 = alpha2 - gamma * gamma / rho1;
    if (beta2 != 0.0 && rho1 != 0.0) {
      beta3 = (alpha1 - gamma * rho2 / beta2) / rho1;
      beta4 = rho2 / beta2;
    }
  }
  s[8] = beta3;
  s[9] = beta4;
}
__global__ void __launch_bounds__(kBlock)
kcycle_combine_kernel(int n, const double* __restrict__ coef, const double* __restrict__ c1, double* __restrict__ x) {
  const int i = blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  const double b3 = coef[0], b4 = coef[1];
  x[i] = b4 == 0.0 ? b3 * c1[i] : b3 * c1[i] + b4 * x[i];
}

}  // namespace mamg

// Row kernels on a sliced-ELL layout: SpMV / residual, Gauss-Seidel colour sweeps, Jacobi, coarse
// scaling dots and the residual that feeds the aggregate restriction (K1, K3, K6 of SURVEY 2.4;
// HAZmath dcsr_mxv / dcsr_aAxpy / smoother_dcsr_gs / dcsr_mxv_agg).
//
// Layout ("SELL-32", built on the device from the colour-permuted CSR of a level): the rows are cut
// into slices of 32 consecutive rows; slice s is W_s = sp[s+1] - sp[s] entry slots wide (the longest
// of its rows; shorter rows are padded with value 0 and their own row as column) and its entries
// start at 32 * sp[s] in both arrays.  One warp owns one slice, LANE = ROW:
//   values   slot pairs   [p][lane][2] doubles  -> one 128-bit load per lane and pair   (LDG.E.128)
//            + one single [lane] column when W_s is odd
//   columns  slot quads   [q][lane][4] ints     -> one 128-bit load per lane and quad   (LDG.E.128)
//            + (W_s mod 4) single [lane] columns
// Every warp-wide load is one contiguous 512-byte (or 256/128-byte) run, so the matrix streams are
// perfectly coalesced whatever the row lengths.  What makes this layout the right one for these
// matrices is the x gather: on the (colour, natural)-ordered rows of a mesh operator, slot e of 32
// consecutive rows addresses 32 (nearly) consecutive entries of x -- the same stencil neighbour of
// neighbouring nodes -- so a warp-wide gather touches 8-9 sectors instead of the 32 that a
// sub-warp-per-row CSR kernel touches (its lanes read 8 different stencil bands).  The L2->SM
// traffic of the gathers drops from ~32 to ~9 bytes per entry, and the kernels stop being bound by
// L2 sector throughput (SELL-C-sigma idea, Kreutzer et al.; sigma = 1: no row sorting, the colour
// blocks keep their mesh order).  Per thread all loads of a row are independent (no shuffles, no
// reduction), so a thread keeps 2 slot quads = 6 wide loads + 8 gathers in flight.
//
// Arithmetic: each row is summed sequentially in ascending (permuted) column order, exactly one
// thread per row -- no cross-lane reduction, results independent of the launch geometry.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace mamg {

struct SellView {
  const int* sp = nullptr;       // [nslices + 1] prefix sum of the slice widths (entries start at 32 * sp[s])
  const double* val = nullptr;
  const int* col = nullptr;
  int n = 0;
  __host__ __device__ int nslices() const { return (n + 31) / 32; }
};

constexpr int kSellWarps = kBlock / 32;   // slices per CTA

// sum_j a_ij x_j of row (32 * slice + lane); padded slots multiply 0 by x[row]
__device__ __forceinline__ double sell_row_dot(const SellView& S, int slice, int lane,
                                               const double* __restrict__ x) {
  const int s0 = S.sp[slice], W = S.sp[slice + 1] - s0;
  const size_t base = (size_t)s0 * 32;
  const double2* vp = reinterpret_cast<const double2*>(S.val + base) + lane;
  const int4* cp = reinterpret_cast<const int4*>(S.col + base) + lane;
  const int nq = W >> 2;
  double s = 0.0;
#pragma unroll 2
  for (int q = 0; q < nq; ++q) {
    const int4 c = __ldg(cp + q * 32);
    const double2 v0 = __ldg(vp + (2 * q) * 32);
    const double2 v1 = __ldg(vp + (2 * q + 1) * 32);
    const double x0 = x[c.x], x1 = x[c.y], x2 = x[c.z], x3 = x[c.w];
    s += v0.x * x0;
    s += v0.y * x1;
    s += v1.x * x2;
    s += v1.y * x3;
  }
  const int t = W & 3;
  if (t) {
    const int* ct = S.col + base + (size_t)nq * 128 + lane;
    if (t >= 2) {
      const double2 v = __ldg(vp + (2 * nq) * 32);
      const int c0 = __ldg(ct), c1 = __ldg(ct + 32);
      const double x0 = x[c0], x1 = x[c1];
      s += v.x * x0;
      s += v.y * x1;
    }
    if (t & 1) {
      const double v = __ldg(S.val + base + (size_t)(W >> 1) * 64 + lane);
      const int c = __ldg(ct + (t - 1) * 32);
      s += v * x[c];
    }
  }
  return s;
}

// ---- build: fill slice arrays from the (permuted) CSR of the level; one warp per slice ---------
__global__ void __launch_bounds__(kBlock)
sell_fill_kernel(int n, int row_lo, int row_hi, const int* __restrict__ ia, const int* __restrict__ ja,
                 const double* __restrict__ a, const int* __restrict__ sp, double* __restrict__ val, int* __restrict__ col) {
  const int slice = blockIdx.x * kSellWarps + threadIdx.x / 32, lane = threadIdx.x % 32;
  if (slice * 32 >= n) return;
  const int row = slice * 32 + lane;
  const bool mine = row < n && row >= row_lo && row < row_hi;
  const int p0 = mine ? ia[row] : 0, len = mine ? ia[row + 1] - p0 : 0;
  const int s0 = sp[slice], W = sp[slice + 1] - s0;
  const size_t base = (size_t)s0 * 32;
  const int pad_col = row < n ? row : 0;
  for (int e = 0; e < W; ++e) {
    const double v = e < len ? a[p0 + e] : 0.0;
    const int c = e < len ? ja[p0 + e] : pad_col;
    const size_t pv = (e >> 1) < (W >> 1) ? base + (size_t)(e >> 1) * 64 + lane * 2 + (e & 1)
                                          : base + (size_t)(W >> 1) * 64 + lane;
    const size_t pc = (e >> 2) < (W >> 2) ? base + (size_t)(e >> 2) * 128 + lane * 4 + (e & 3)
                                          : base + (size_t)(W >> 2) * 128 + (size_t)(e - (W & ~3)) * 32 + lane;
    val[pv] = v;
    col[pc] = c;
  }
}

// ---- K1  y = A x | y = b - A x on rows [r0, r1) -----------------------------------------------------
template <bool RESID>
__global__ void __launch_bounds__(kBlock)
sell_spmv_kernel(int r0, int r1, const SellView S, const double* __restrict__ x,
                 const double* __restrict__ b, double* __restrict__ y) {
  const int lane = threadIdx.x % 32;
  const int slice = r0 / 32 + blockIdx.x * kSellWarps + threadIdx.x / 32;
  const int row = slice * 32 + lane;
  if (slice * 32 >= r1) return;
  const double s = sell_row_dot(S, slice, lane, x);
  if (row >= r0 && row < r1) y[row] = RESID ? b[row] - s : s;
}

// q = A d with d.q in the epilogue; sc[1] = d.q, sc[2] = alpha = rz / d.q   (K1 + K10 fused)
__global__ void __launch_bounds__(kBlock)
sell_spmv_dot_kernel(const SellView S, const double* __restrict__ d, double* __restrict__ q,
                     double* partial, unsigned int* ticket, double* sc) {
  const int lane = threadIdx.x % 32;
  const int ns = S.nslices();
  double v[1] = {0.0};
  for (int slice = blockIdx.x * kSellWarps + threadIdx.x / 32; slice < ns; slice += gridDim.x * kSellWarps) {
    const double s = sell_row_dot(S, slice, lane, d);
    const int row = slice * 32 + lane;
    if (row < S.n) { q[row] = s; v[0] += s * d[row]; }
  }
  if (block_reduce_finish<1>(v, partial, ticket, sc + 1) && threadIdx.x == 0) sc[2] = sc[0] / sc[1];
}

// ---- K6  one colour of Gauss-Seidel / SOR on rows [r0, r1): x_i += w (b_i - a_i . x) / a_ii -------
template <bool TAIL>
__global__ void __launch_bounds__(kBlock)
sell_gs_kernel(int r0, int r1, const SellView S, const double* __restrict__ invd,
               const uint8_t* __restrict__ skip, const double* __restrict__ b, double* x, double omega,
               const HaloTail tail) {
  const int lane = threadIdx.x % 32;
  const int slice = r0 / 32 + blockIdx.x * kSellWarps + threadIdx.x / 32;
  if (slice * 32 < r1) {
    const int row = slice * 32 + lane;
    const bool active = row >= r0 && row < r1 && !(skip != nullptr && skip[row]);
    if (__any_sync(0xffffffffu, active)) {
      // lanes of a slice that straddles the colour boundary (or Schwarz rows) compute and discard: the
      // rows of one colour do not couple, so what they read is never what an active lane writes
      const double s = sell_row_dot(S, slice, lane, x);
      if (active) x[row] += omega * (b[row] - s) * invd[row];
    }
  }
  if (TAIL) halo_tail(tail);   // multi-GPU halo mode: the last block sends this colour's boundary rows to the neighbours
}

// damped Jacobi, out of place
__global__ void __launch_bounds__(kBlock)
sell_jacobi_kernel(int r0, int r1, const SellView S, const double* __restrict__ invd,
                   const uint8_t* __restrict__ skip, const double* __restrict__ b,
                   const double* __restrict__ x, double* __restrict__ xn, double omega) {
  const int lane = threadIdx.x % 32;
  const int slice = r0 / 32 + blockIdx.x * kSellWarps + threadIdx.x / 32;
  if (slice * 32 >= r1) return;
  const int row = slice * 32 + lane;
  const double s = sell_row_dot(S, slice, lane, x);
  if (row >= r0 && row < r1) {
    const bool sk = skip != nullptr && skip[row];
    xn[row] = sk ? x[row] : x[row] + omega * (b[row] - s) * invd[row];
  }
}

// ---- coarse scaling: out[0] = e.r, out[1] = e.Ae, out[2] = min(e.r / e.Ae, 1) ----------------------
__global__ void __launch_bounds__(kBlock)
sell_scale_dots_kernel(const SellView S, const double* __restrict__ e, const double* __restrict__ r,
                       double* partial, unsigned int* ticket, double* out) {
  const int lane = threadIdx.x % 32;
  const int ns = S.nslices();
  double v[2] = {0.0, 0.0};
  for (int slice = blockIdx.x * kSellWarps + threadIdx.x / 32; slice < ns; slice += gridDim.x * kSellWarps) {
    const double s = sell_row_dot(S, slice, lane, e);
    const int row = slice * 32 + lane;
    if (row < S.n) { const double ei = e[row]; v[0] += ei * r[row]; v[1] += ei * s; }
  }
  if (block_reduce_finish<2>(v, partial, ticket, out) && threadIdx.x == 0) {
    const double al = out[0] / out[1];
    out[2] = (al < 1.0) ? al : 1.0;
  }
}

// ---- K3  restriction of a stored fine residual t = b - A x over the aggregates:
//          bc[I] = sum_{i in I} t_i (members in ascending fine row order), xc[I] = 0.
//     The residual itself comes from sell_spmv_kernel<true>, which streams every fine row exactly
//     once in layout order; the members of an aggregate sit in different colour blocks, so walking
//     the matrix aggregate by aggregate (round 1) read 1.8 x the algorithmic bytes.
__global__ void __launch_bounds__(kBlock)
agg_sum_kernel(int c0, int c1, const int* __restrict__ cptr, const int* __restrict__ cidx,
               const double* __restrict__ t, double* __restrict__ bc, double* __restrict__ xc) {
  const int I = c0 + blockIdx.x * kBlock + threadIdx.x;
  if (I >= c1) return;
  double acc = 0.0;
  for (int q = cptr[I]; q < cptr[I + 1]; ++q) acc += t[cidx[q]];
  bc[I] = acc;
  xc[I] = 0.0;
}

// the same sums for a list of coarse rows (first replicated level below a row-distributed one: every
// rank sums the aggregates it owns)
__global__ void __launch_bounds__(kBlock)
agg_sum_list_kernel(int cnt, const int* __restrict__ rows, const int* __restrict__ cptr, const int* __restrict__ cidx,
                    const double* __restrict__ t, double* __restrict__ bc) {
  const int q = blockIdx.x * kBlock + threadIdx.x;
  if (q >= cnt) return;
  const int I = rows[q];
  double acc = 0.0;
  for (int p = cptr[I]; p < cptr[I + 1]; ++p) acc += t[cidx[p]];
  bc[I] = acc;
}

// ---- row-range variants for a row-distributed level (halo mode): partial sums over the rows [r0, r1)
//      this rank owns; the all-reduce kernel combines them in rank order and post-processes the scalars
__global__ void __launch_bounds__(kBlock)
sell_spmv_dot_range_kernel(int r0, int r1, const SellView S, const double* __restrict__ d, double* __restrict__ q,
                           double* partial, unsigned int* ticket, double* out) {
  const int lane = threadIdx.x % 32;
  const int s0 = r0 / 32, s1 = (r1 + 31) / 32;
  double v[1] = {0.0};
  for (int slice = s0 + blockIdx.x * kSellWarps + threadIdx.x / 32; slice < s1; slice += gridDim.x * kSellWarps) {
    const double s = sell_row_dot(S, slice, lane, d);
    const int row = slice * 32 + lane;
    if (row >= r0 && row < r1) { q[row] = s; v[0] += s * d[row]; }
  }
  block_reduce_finish<1>(v, partial, ticket, out);
}
__global__ void __launch_bounds__(kBlock)
sell_scale_dots_range_kernel(int r0, int r1, const SellView S, const double* __restrict__ e, const double* __restrict__ r,
                             double* partial, unsigned int* ticket, double* out) {
  const int lane = threadIdx.x % 32;
  const int s0 = r0 / 32, s1 = (r1 + 31) / 32;
  double v[2] = {0.0, 0.0};
  for (int slice = s0 + blockIdx.x * kSellWarps + threadIdx.x / 32; slice < s1; slice += gridDim.x * kSellWarps) {
    const double s = sell_row_dot(S, slice, lane, e);
    const int row = slice * 32 + lane;
    if (row >= r0 && row < r1) { const double ei = e[row]; v[0] += ei * r[row]; v[1] += ei * s; }
  }
  block_reduce_finish<2>(v, partial, ticket, out);
}

}  // namespace mamg

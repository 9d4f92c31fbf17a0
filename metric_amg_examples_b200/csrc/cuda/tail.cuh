// K9  persistent coarse tail: the whole sub-cycle of the levels below a size threshold inside ONE
// kernel (one CTA of 1024 threads), level vectors in shared memory, matrices through L1/L2.
//
// Replaces the bottom of HAZmath's mgcycle: for a hierarchy of ~19 levels a V-cycle
// spends ~70 dependent colour phases on each of the ~8 levels that have fewer than a few thousand
// rows; as separate launches each phase costs microseconds of launch + ramp latency for
// nanoseconds of work, and a W-cycle visits those levels 2^l times.  Here a phase boundary is a
// __syncthreads().  The arithmetic is the one of the per-level kernels (same colour order, same
// row formula, same FASP visit counters), so the parity tests cover it whenever a test hierarchy
// has a level below the threshold.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace mamg {

constexpr int kTailThreads = 1024;
constexpr int kTailMaxLevels = 16;
constexpr int kTailLanes = 8;

struct TailLevel {
  int n, nc, ncolors;
  const int *ia, *ja, *color_ptr, *agg, *cptr, *cidx;
  const double *a, *invd;
  int off;   // offset of this level's x (and, after all x, b) inside the shared vectors
};

struct TailArgs {
  TailLevel lv[kTailMaxLevels];
  int nlev;            // tail levels (the last one is the coarsest of the hierarchy)
  int total;           // sum of n over the tail levels
  int cycle_type;      // V_CYCLE 1 | W_CYCLE 2
  int top_reps;        // visits of the first tail level per call (cycle_type, or 1 when it is level 0)
  int smoother;        // SMOOTHER_GS 2 | SGS 3 | SOR 5 | SSOR 6
  int pre, post, scaling;
  double omega;
  const double* coarse_inv;
  const double* b_in;  // right-hand side of the first tail level (global)
  double* x_io;        // its iterate (global): read when x_nonzero, always written
  int x_nonzero;
};

__device__ __forceinline__ void tail_gs_color(const TailLevel& L, int c, const double* b, double* x, double omega) {
  const int lane = threadIdx.x % kTailLanes;
  const int r0 = L.color_ptr[c], r1 = L.color_ptr[c + 1];
  for (int row = r0 + threadIdx.x / kTailLanes; row < r1; row += kTailThreads / kTailLanes) {
    double s = 0.0;
    for (int p = L.ia[row] + lane; p < L.ia[row + 1]; p += kTailLanes) s += __ldg(L.a + p) * x[__ldg(L.ja + p)];
    s = subwarp_sum<kTailLanes>(s);
    if (lane == 0) x[row] += omega * (b[row] - s) * L.invd[row];
  }
  __syncthreads();
}

__device__ __forceinline__ void tail_smooth(const TailArgs& A, const TailLevel& L, const double* b, double* x, bool post) {
  const int iters = post ? A.post : A.pre;
  for (int it = 0; it < iters; ++it) {
    switch (A.smoother) {
      case 2:  // GS: forward before, backward after
      case 5:  // SOR
        if (!post) for (int c = 0; c < L.ncolors; ++c) tail_gs_color(L, c, b, x, A.smoother == 2 ? 1.0 : A.omega);
        else for (int c = L.ncolors - 1; c >= 0; --c) tail_gs_color(L, c, b, x, A.smoother == 2 ? 1.0 : A.omega);
        break;
      case 3:  // SGS: forward, then backward without repeating the colour just finished
        for (int c = 0; c < L.ncolors; ++c) tail_gs_color(L, c, b, x, 1.0);
        for (int c = L.ncolors - 2; c >= 0; --c) tail_gs_color(L, c, b, x, 1.0);
        break;
      default:  // SSOR
        for (int c = 0; c < L.ncolors; ++c) tail_gs_color(L, c, b, x, A.omega);
        for (int c = L.ncolors - 1; c >= 0; --c) tail_gs_color(L, c, b, x, A.omega);
        break;
    }
  }
}

// b_c[I] = sum over the aggregate of (b_i - a_i . x), x_c = 0
__device__ __forceinline__ void tail_restrict(const TailLevel& F, const double* b, const double* x, double* bc, double* xc) {
  const int lane = threadIdx.x % kTailLanes;
  for (int I = threadIdx.x / kTailLanes; I < F.nc; I += kTailThreads / kTailLanes) {
    double acc = 0.0;
    for (int q = F.cptr[I]; q < F.cptr[I + 1]; ++q) {
      const int i = F.cidx[q];
      double s = 0.0;
      for (int p = F.ia[i] + lane; p < F.ia[i + 1]; p += kTailLanes) s += __ldg(F.a + p) * x[__ldg(F.ja + p)];
      s = subwarp_sum<kTailLanes>(s);
      acc += b[i] - s;
    }
    if (lane == 0) { bc[I] = acc; xc[I] = 0.0; }
  }
  __syncthreads();
}

// alpha = min((e.r)/(e.Ae), 1) on level C (fixed-order block reduction)
__device__ __forceinline__ double tail_alpha(const TailLevel& C, const double* e, const double* r, double* red) {
  const int lane = threadIdx.x % kTailLanes;
  double v0 = 0.0, v1 = 0.0;
  for (int row = threadIdx.x / kTailLanes; row < C.n; row += kTailThreads / kTailLanes) {
    double s = 0.0;
    for (int p = C.ia[row] + lane; p < C.ia[row + 1]; p += kTailLanes) s += __ldg(C.a + p) * e[__ldg(C.ja + p)];
    s = subwarp_sum<kTailLanes>(s);
    if (lane == 0) { v0 += e[row] * r[row]; v1 += e[row] * s; }
  }
  v0 = warp_sum(v0);
  v1 = warp_sum(v1);
  const int w = threadIdx.x / 32;
  if ((threadIdx.x & 31) == 0) { red[w] = v0; red[32 + w] = v1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a0 = 0.0, a1 = 0.0;
    for (int k = 0; k < kTailThreads / 32; ++k) { a0 += red[k]; a1 += red[32 + k]; }
    const double al = a0 / a1;
    red[64] = (al < 1.0) ? al : 1.0;
  }
  __syncthreads();
  const double alpha = red[64];
  __syncthreads();
  return alpha;
}

__global__ void __launch_bounds__(kTailThreads)
tail_cycle_kernel(const TailArgs A) {
  extern __shared__ double sm[];
  double* X = sm;               // x of every tail level, then b of every tail level
  double* B = sm + A.total;
  double* red = B + A.total;    // 65 doubles
  const int nl = A.nlev;
  {
    const TailLevel& T = A.lv[0];
    for (int i = threadIdx.x; i < T.n; i += kTailThreads) {
      B[T.off + i] = A.b_in[i];
      X[T.off + i] = A.x_nonzero ? A.x_io[i] : 0.0;
    }
  }
  __syncthreads();
  int cnt[kTailMaxLevels];
  for (int k = 0; k < kTailMaxLevels; ++k) cnt[k] = 0;
  int l = 0;
  while (true) {
    // forward sweep down to the coarsest level
    while (l < nl - 1) {
      ++cnt[l];
      const TailLevel& F = A.lv[l];
      const TailLevel& C = A.lv[l + 1];
      tail_smooth(A, F, B + F.off, X + F.off, false);
      tail_restrict(F, B + F.off, X + F.off, B + C.off, X + C.off);
      ++l;
    }
    {  // coarsest: x = Ainv b, one warp per row
      const TailLevel& C = A.lv[nl - 1];
      const int lane = threadIdx.x & 31;
      for (int row = threadIdx.x / 32; row < C.n; row += kTailThreads / 32) {
        double s = 0.0;
        for (int j = lane; j < C.n; j += 32) s += A.coarse_inv[(size_t)row * C.n + j] * B[C.off + j];
        s = warp_sum(s);
        if (lane == 0) X[C.off + row] = s;
      }
      __syncthreads();
    }
    // backward sweep; a level that has not had its cycle_type visits yet turns around again
    bool again = false;
    while (l > 0) {
      --l;
      const TailLevel& F = A.lv[l];
      const TailLevel& C = A.lv[l + 1];
      const double alpha = A.scaling ? tail_alpha(C, X + C.off, B + C.off, red) : 1.0;
      for (int i = threadIdx.x; i < F.n; i += kTailThreads) {
        const int I = F.agg[i];
        if (I >= 0) X[F.off + i] += alpha * X[C.off + I];
      }
      __syncthreads();
      tail_smooth(A, F, B + F.off, X + F.off, true);
      const int need = l == 0 ? A.top_reps : A.cycle_type;
      if (cnt[l] < need) { again = true; break; }
      cnt[l] = 0;
    }
    if (!again) break;
  }
  const TailLevel& T = A.lv[0];
  for (int i = threadIdx.x; i < T.n; i += kTailThreads) A.x_io[i] = X[T.off + i];
}

}  // namespace mamg

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
// Apply: a warp (or CTA for patches > 32 dofs) streams the patch's CSR rows once to form the
// residual (b - A x)_B, stages the packed inverse through shared memory and applies it as a
// small symmetric mat-vec.  The kernel is HBM-bound: per patch it reads 12 B per row entry plus
// 4 s(s+1) bytes of inverse.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <functional>
#include <vector>

#include "../host/hierarchy.h"

namespace mamg {

struct DSchwarz {
  int npatch = 0, ncolors = 0, max_size = 0, warps = 1, ppc = 1;
  size_t smem_apply = 0, smem_setup = 0;
  int* pptr = nullptr;         // npatch+1, patches sorted by colour
  int* pdofs = nullptr;        // patch dofs (permuted row ids) in natural-ascending order
  long long* ioff = nullptr;   // npatch+1: start of every packed inverse
  double* pinv = nullptr;      // packed lower triangles of A_BB^{-1}
  std::vector<int> color_ptr;  // host: patch range of every colour
  long long alg_bytes = 0;     // algorithmic bytes of one sweep over all patches
};

__device__ __forceinline__ int tri(int i, int j) { return i * (i + 1) / 2 + j; }  // i >= j

// ---- setup: packed inverse of every A_BB -------------------------------------------------------
template <int T>
__global__ void __launch_bounds__(T)
schwarz_invert_kernel(int npatch, const int* __restrict__ pptr, const int* __restrict__ pdofs,
                      const long long* __restrict__ ioff, const int* __restrict__ ia,
                      const int* __restrict__ ja, const double* __restrict__ a,
                      double* __restrict__ pinv, int max_size) {
  extern __shared__ double smem[];
  const int patch = blockIdx.x;
  if (patch >= npatch) return;
  const int q0 = pptr[patch], s = pptr[patch + 1] - q0;
  double* Lm = smem;                                                  // packed lower triangle
  double* col = Lm + (size_t)max_size * (max_size + 1) / 2;           // s scratch
  int* idx = reinterpret_cast<int*>(col + max_size);                  // s
  const int tid = threadIdx.x;
  for (int k = tid; k < s * (s + 1) / 2; k += T) Lm[k] = 0.0;
  for (int k = tid; k < s; k += T) idx[k] = pdofs[q0 + k];
  __syncthreads();
  // gather the lower triangle of A_BB: thread per (row k, entry) with a linear search of the column
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
  double* out = pinv + ioff[patch];
  for (int k = tid; k < s * (s + 1) / 2; k += T) out[k] = Lm[k];
}

// ---- apply ---------------------------------------------------------------------------------------
// WPP warps per patch, PPC patches per CTA (PPC > 1 only with WPP == 1: warps are independent)
template <int WPP, int PPC>
__global__ void __launch_bounds__(WPP * PPC * 32)
schwarz_apply_kernel(int p0, int p1, const int* __restrict__ pptr, const int* __restrict__ pdofs,
                     const long long* __restrict__ ioff, const double* __restrict__ pinv,
                     const int* __restrict__ ia, const int* __restrict__ ja,
                     const double* __restrict__ a, const double* __restrict__ b, double* x,
                     int max_size) {
  constexpr int T = WPP * 32;  // threads per patch
  extern __shared__ double smem[];
  const int slot = threadIdx.x / T;
  const int tid = threadIdx.x % T;
  const int patch = p0 + blockIdx.x * PPC + slot;
  const size_t per_patch = (size_t)max_size * (max_size + 1) / 2 + 2 * (size_t)max_size;
  double* Inv = smem + slot * per_patch;
  double* rhs = Inv + (size_t)max_size * (max_size + 1) / 2;
  int* idx = reinterpret_cast<int*>(rhs + max_size);
  const bool active = patch < p1;   // uniform per warp when PPC > 1 (WPP == 1)
  int s = 0, q0 = 0;
  if (active) {
    q0 = pptr[patch];
    s = pptr[patch + 1] - q0;
    for (int k = tid; k < s; k += T) idx[k] = pdofs[q0 + k];
    const double* src = pinv + ioff[patch];
    const int np = s * (s + 1) / 2;
    for (int k = tid; k < np; k += T) Inv[k] = src[k];
  }
  if (WPP == 1) __syncwarp(); else __syncthreads();
  if (active) {
    // residual of the patch rows: one warp per row, 4 rows in flight per warp
    const int lane = tid & 31, wrp = tid / 32;
    for (int k0 = wrp * 4; k0 < s; k0 += WPP * 4) {
      double acc[4] = {0.0, 0.0, 0.0, 0.0};
      int row[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = k0 + u;
        row[u] = k < s ? idx[k] : -1;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (row[u] < 0) continue;
        const int r0 = ia[row[u]], r1 = ia[row[u] + 1];
        for (int p = r0 + lane; p < r1; p += 32) acc[u] += a[p] * x[ja[p]];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[u] += __shfl_down_sync(0xffffffffu, acc[u], o);
        if (lane == 0 && row[u] >= 0) rhs[k0 + u] = b[row[u]] - acc[u];
      }
    }
  }
  if (WPP == 1) __syncwarp(); else __syncthreads();
  if (active) {
    // delta = A_BB^{-1} rhs with the packed symmetric inverse
    for (int k = tid; k < s; k += T) {
      double d = 0.0;
      const int base = k * (k + 1) / 2;
      for (int c = 0; c <= k; ++c) d += Inv[base + c] * rhs[c];
      for (int c = k + 1; c < s; ++c) d += Inv[c * (c + 1) / 2 + k] * rhs[c];
      x[idx[k]] += d;
    }
  }
}

// Host side: reorder the patches by colour, translate dofs to the permuted numbering, upload
// and invert on the device.  `alloc(bytes)` returns tracked device memory.
inline void schwarz_upload(const Level& hl, const std::vector<int>& iperm, const int* d_ia,
                           const int* d_ja, const double* d_a, const std::vector<int>& pia,
                           DSchwarz& d, const std::function<void*(size_t)>& alloc) {
  const SchwarzPatches& sw = hl.sw;
  const int np = sw.npatch();
  d.npatch = np;
  d.ncolors = sw.ncolors;
  d.max_size = sw.max_size;
  const size_t tri_max = (size_t)d.max_size * (d.max_size + 1) / 2;
  d.smem_setup = (tri_max + d.max_size) * sizeof(double) + (size_t)d.max_size * sizeof(int);
  if (d.smem_setup > 227 * 1024)
    throw std::runtime_error("Schwarz patch of " + std::to_string(d.max_size) + " dofs does not fit in shared memory");
  d.warps = d.max_size <= 32 ? 1 : (d.max_size <= 96 ? 2 : 4);
  d.ppc = d.warps == 1 ? 4 : 1;
  d.smem_apply = (size_t)d.ppc * (tri_max + 2 * (size_t)d.max_size) * sizeof(double);
  d.color_ptr.assign(sw.ncolors + 1, 0);
  for (int p = 0; p < np; ++p) ++d.color_ptr[sw.color[p] + 1];
  for (int c = 0; c < sw.ncolors; ++c) d.color_ptr[c + 1] += d.color_ptr[c];
  std::vector<int> order(np);
  {
    std::vector<int> fill(d.color_ptr.begin(), d.color_ptr.end() - 1);
    for (int p = 0; p < np; ++p) order[fill[sw.color[p]]++] = p;
  }
  std::vector<int> pptr(np + 1, 0), pdofs(sw.dofs.size());
  std::vector<long long> ioff(np + 1, 0);
  long long row_entries = 0;
  for (int k = 0; k < np; ++k) {
    const int p = order[k];
    const int s = sw.ptr[p + 1] - sw.ptr[p];
    pptr[k + 1] = pptr[k] + s;
    ioff[k + 1] = ioff[k] + (long long)s * (s + 1) / 2;
    for (int q = 0; q < s; ++q) {
      const int i = iperm[sw.dofs[sw.ptr[p] + q]];
      pdofs[pptr[k] + q] = i;
      row_entries += pia[i + 1] - pia[i];
    }
  }
  // per sweep: row entries (val + col), packed inverse, patch index, b and x of the patch dofs
  d.alg_bytes = 12 * row_entries + 8 * ioff[np] + (4 + 8 + 16) * (long long)pdofs.size() + 8 * (long long)np;
  auto up = [&](const void* src, size_t bytes) {
    void* p = alloc(bytes);
    if (bytes) cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice);
    return p;
  };
  d.pptr = (int*)up(pptr.data(), pptr.size() * sizeof(int));
  d.pdofs = (int*)up(pdofs.data(), pdofs.size() * sizeof(int));
  d.ioff = (long long*)up(ioff.data(), ioff.size() * sizeof(long long));
  d.pinv = (double*)alloc((size_t)ioff[np] * sizeof(double));
  constexpr int TS = 128;
  if (d.smem_setup > 48 * 1024)
    cudaFuncSetAttribute(schwarz_invert_kernel<TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d.smem_setup);
  schwarz_invert_kernel<TS><<<np, TS, d.smem_setup>>>(np, d.pptr, d.pdofs, d.ioff, d_ia, d_ja, d_a, d.pinv, d.max_size);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) throw std::runtime_error(std::string("Schwarz setup kernel failed: ") + cudaGetErrorString(e));
  if (d.smem_apply > 48 * 1024) {
    cudaFuncSetAttribute(schwarz_apply_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d.smem_apply);
    cudaFuncSetAttribute(schwarz_apply_kernel<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d.smem_apply);
  }
}

// one colour of a multiplicative sweep
inline void schwarz_color_launch(const DSchwarz& d, int c, const int* ia, const int* ja, const double* a,
                                 const double* b, double* x, cudaStream_t stream) {
  const int p0 = d.color_ptr[c], p1 = d.color_ptr[c + 1];
  const int grid = (p1 - p0 + d.ppc - 1) / d.ppc;
  switch (d.warps) {
    case 1: schwarz_apply_kernel<1, 4><<<grid, 128, d.smem_apply, stream>>>(p0, p1, d.pptr, d.pdofs, d.ioff, d.pinv, ia, ja, a, b, x, d.max_size); break;
    case 2: schwarz_apply_kernel<2, 1><<<grid, 64, d.smem_apply, stream>>>(p0, p1, d.pptr, d.pdofs, d.ioff, d.pinv, ia, ja, a, b, x, d.max_size); break;
    default: schwarz_apply_kernel<4, 1><<<grid, 128, d.smem_apply, stream>>>(p0, p1, d.pptr, d.pdofs, d.ioff, d.pinv, ia, ja, a, b, x, d.max_size); break;
  }
}

}  // namespace mamg

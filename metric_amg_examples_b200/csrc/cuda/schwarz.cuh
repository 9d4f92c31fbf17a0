// K7  interface Schwarz smoother: batched small dense patch solves in shared memory.
//
// Replaces HAZmath smoother_dcsr_Schwarz_forward/backward (one UMFPACK solve per block,
// sequential over blocks) configured by src/amg_parameters.py:82-87.  For every patch B:
//     x_B <- x_B + A_BB^{-1} (b - A x)_B
// Patches are coloured by conflict (schwarz_color in csrc/host/setup.cpp), so all patches of
// one colour are solved concurrently with a result identical to visiting them one by one;
// forward = colours ascending, backward = descending.
//
// One CTA (1, 2 or 4 warps) owns one patch.  The patch rows are streamed from the level's CSR
// exactly once: the same pass accumulates the residual (b - A x)_B and scatters the A_BB
// entries into a packed lower-triangular matrix in shared memory (a per-entry byte map gives
// the local column or 255).  A_BB is then Cholesky-factorised and solved in shared memory;
// storing the factors instead would cost more HBM traffic than re-reading the rows that the
// residual needs anyway (Schwarz).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <functional>
#include <vector>

#include "../host/hierarchy.h"

namespace mamg {

struct DSchwarz {
  int npatch = 0, ncolors = 0, max_size = 0, warps = 1;
  size_t smem = 0;
  int* pptr = nullptr;        // npatch+1, patches sorted by colour
  int* pdofs = nullptr;       // patch dofs (permuted row ids) in natural-ascending order
  int* poff = nullptr;        // per patch dof: offset of its row inside the patch's lmap segment
  long long* lbase = nullptr; // per patch: start of its lmap segment
  uint8_t* lmap = nullptr;    // per (patch row, row entry): local column inside the patch or 255
  std::vector<int> color_ptr; // host: patch range of every colour
};

constexpr int kSwSub = 8;  // lanes that share one matrix row while gathering

__device__ __forceinline__ int tri(int i, int j) { return i * (i + 1) / 2 + j; }  // i >= j

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
schwarz_patch_kernel(int p0, const int* __restrict__ pptr, const int* __restrict__ pdofs,
                     const int* __restrict__ poff, const long long* __restrict__ lbase,
                     const uint8_t* __restrict__ lmap, const int* __restrict__ ia,
                     const int* __restrict__ ja, const double* __restrict__ a,
                     const double* __restrict__ b, double* x, int max_size) {
  constexpr int T = WARPS * 32;
  extern __shared__ double smem[];
  const int patch = p0 + blockIdx.x;
  const int q0 = pptr[patch], s = pptr[patch + 1] - q0;
  double* Lm = smem;                                   // packed lower triangle, s(s+1)/2
  double* rhs = Lm + (size_t)max_size * (max_size + 1) / 2;  // s
  int* idx = reinterpret_cast<int*>(rhs + max_size);   // s
  const int tid = threadIdx.x;
  for (int k = tid; k < s * (s + 1) / 2; k += T) Lm[k] = 0.0;
  for (int k = tid; k < s; k += T) idx[k] = pdofs[q0 + k];
  __syncthreads();
  // ---- stream the patch rows once: residual and A_BB ----
  {
    const int sl = tid % kSwSub, grp = tid / kSwSub;
    const unsigned int mask = ((1u << kSwSub) - 1u) << ((tid & 31) / kSwSub * kSwSub);
    const uint8_t* lm = lmap + lbase[patch];
    for (int k = grp; k < s; k += T / kSwSub) {
      const int i = idx[k];
      const int r0 = ia[i], r1 = ia[i + 1];
      const uint8_t* lmk = lm + poff[q0 + k];
      double acc = 0.0;
      for (int p = r0 + sl; p < r1; p += kSwSub) {
        const double v = a[p];
        acc += v * x[ja[p]];
        const int c = lmk[p - r0];
        if (c <= k) Lm[tri(k, c)] = v;  // 255 (outside the patch) never passes: k < 255
      }
#pragma unroll
      for (int o = kSwSub / 2; o > 0; o >>= 1) acc += __shfl_down_sync(mask, acc, o, kSwSub);
      if (sl == 0) rhs[k] = b[i] - acc;
    }
  }
  __syncthreads();
  // ---- Cholesky A_BB = L L' (right-looking, in place) with the forward solve L y = rhs ----
  for (int j = 0; j < s; ++j) {
    const double d = sqrt(Lm[tri(j, j)]);
    const double invd = 1.0 / d;
    __syncthreads();
    if (tid == 0) { Lm[tri(j, j)] = d; rhs[j] *= invd; }
    for (int i = j + 1 + tid; i < s; i += T) Lm[tri(i, j)] *= invd;
    __syncthreads();
    // trailing update: rows i > j, columns j < k <= i, plus the rhs as an extra column
    // (warps stride over columns k, lanes over rows i >= k)
    for (int k = j + 1 + tid / 32; k < s; k += WARPS) {
      const double lkj = Lm[tri(k, j)];
      for (int i = k + (tid & 31); i < s; i += 32) Lm[tri(i, k)] -= Lm[tri(i, j)] * lkj;
    }
    for (int i = j + 1 + tid; i < s; i += T) rhs[i] -= Lm[tri(i, j)] * rhs[j];
    __syncthreads();
  }
  // ---- backward solve L' delta = y ----
  for (int j = s - 1; j >= 0; --j) {
    if (tid == 0) rhs[j] /= Lm[tri(j, j)];
    __syncthreads();
    const double dj = rhs[j];
    for (int i = tid; i < j; i += T) rhs[i] -= Lm[tri(j, i)] * dj;
    __syncthreads();
  }
  for (int k = tid; k < s; k += T) x[idx[k]] += rhs[k];
}

// Host side: reorder the patches by colour, translate dofs to the permuted numbering and
// build the per-entry local-column map.  `alloc(bytes)` returns tracked device memory.
inline void schwarz_upload(const Level& hl, const std::vector<int>& perm, const std::vector<int>& iperm,
                           const std::vector<int>& pia, const std::vector<int>& pja, DSchwarz& d,
                           const std::function<void*(size_t)>& alloc) {
  (void)perm;
  const SchwarzPatches& sw = hl.sw;
  const int np = sw.npatch();
  d.npatch = np;
  d.ncolors = sw.ncolors;
  d.max_size = sw.max_size;
  if (d.max_size > 254) throw std::runtime_error("Schwarz_mmsize > 254 is not supported by the device patch kernel");
  d.warps = d.max_size <= 32 ? 1 : (d.max_size <= 64 ? 2 : 4);
  d.smem = ((size_t)d.max_size * (d.max_size + 1) / 2 + d.max_size) * sizeof(double) + (size_t)d.max_size * sizeof(int);
  d.color_ptr.assign(sw.ncolors + 1, 0);
  for (int p = 0; p < np; ++p) ++d.color_ptr[sw.color[p] + 1];
  for (int c = 0; c < sw.ncolors; ++c) d.color_ptr[c + 1] += d.color_ptr[c];
  std::vector<int> order(np);
  {
    std::vector<int> fill(d.color_ptr.begin(), d.color_ptr.end() - 1);
    for (int p = 0; p < np; ++p) order[fill[sw.color[p]]++] = p;
  }
  std::vector<int> pptr(np + 1, 0), pdofs(sw.dofs.size()), poff(sw.dofs.size());
  std::vector<long long> lbase(np + 1, 0);
  for (int k = 0; k < np; ++k) {
    const int p = order[k];
    const int s = sw.ptr[p + 1] - sw.ptr[p];
    pptr[k + 1] = pptr[k] + s;
    long long len = 0;
    for (int q = 0; q < s; ++q) {
      const int i = iperm[sw.dofs[sw.ptr[p] + q]];
      pdofs[pptr[k] + q] = i;
      poff[pptr[k] + q] = (int)len;
      len += pia[i + 1] - pia[i];
    }
    lbase[k + 1] = lbase[k] + len;
  }
  std::vector<uint8_t> lmap((size_t)lbase[np]);
  const int n = hl.A.n;
#pragma omp parallel
  {
    std::vector<uint8_t> loc(n, 255);
#pragma omp for schedule(dynamic, 1024)
    for (int k = 0; k < np; ++k) {
      const int s = pptr[k + 1] - pptr[k];
      for (int q = 0; q < s; ++q) loc[pdofs[pptr[k] + q]] = (uint8_t)q;
      for (int q = 0; q < s; ++q) {
        const int i = pdofs[pptr[k] + q];
        uint8_t* out = &lmap[(size_t)lbase[k] + poff[pptr[k] + q]];
        for (int e = pia[i]; e < pia[i + 1]; ++e) out[e - pia[i]] = loc[pja[e]];
      }
      for (int q = 0; q < s; ++q) loc[pdofs[pptr[k] + q]] = 255;
    }
  }
  auto up = [&](const void* src, size_t bytes) {
    void* p = alloc(bytes);
    if (bytes) cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice);
    return p;
  };
  d.pptr = (int*)up(pptr.data(), pptr.size() * sizeof(int));
  d.pdofs = (int*)up(pdofs.data(), pdofs.size() * sizeof(int));
  d.poff = (int*)up(poff.data(), poff.size() * sizeof(int));
  d.lbase = (long long*)up(lbase.data(), lbase.size() * sizeof(long long));
  d.lmap = (uint8_t*)up(lmap.data(), lmap.size());
  if (d.smem > 48 * 1024) {
    cudaFuncSetAttribute(schwarz_patch_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d.smem);
    cudaFuncSetAttribute(schwarz_patch_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d.smem);
  }
}

// one multiplicative sweep over all patches; returns the number of kernel launches
inline int schwarz_sweep(const DSchwarz& d, const int* ia, const int* ja, const double* a,
                         const double* b, double* x, bool backward, cudaStream_t stream) {
  int launches = 0;
  for (int cc = 0; cc < d.ncolors; ++cc) {
    const int c = backward ? d.ncolors - 1 - cc : cc;
    const int p0 = d.color_ptr[c], cnt = d.color_ptr[c + 1] - p0;
    if (cnt <= 0) continue;
    switch (d.warps) {
      case 1: schwarz_patch_kernel<1><<<cnt, 32, d.smem, stream>>>(p0, d.pptr, d.pdofs, d.poff, d.lbase, d.lmap, ia, ja, a, b, x, d.max_size); break;
      case 2: schwarz_patch_kernel<2><<<cnt, 64, d.smem, stream>>>(p0, d.pptr, d.pdofs, d.poff, d.lbase, d.lmap, ia, ja, a, b, x, d.max_size); break;
      default: schwarz_patch_kernel<4><<<cnt, 128, d.smem, stream>>>(p0, d.pptr, d.pdofs, d.poff, d.lbase, d.lmap, ia, ja, a, b, x, d.max_size); break;
    }
    ++launches;
  }
  return launches;
}

}  // namespace mamg

// CPU setup of the unsmoothed-aggregation hierarchy that the cycle is applied on.
//
// The reference reaches this through metricAMG(A, W, idofs=, parameters=)
// (src/utils.py:86) / AMG(A, parameters=) (src/utils.py:40); the arithmetic lives
// in the un-vendored HAZmath C library, so the choices HAZmath's source would pin are
// frozen here explicitly (SURVEY 8c, items v-vii) and shared with the oracle through
// mamg_level_export:
//   * rows whose off-diagonal entries are all zero (Dirichlet identity rows after the
//     symmetric apply_bc, src/bidomain_2d.py:97) join no aggregate (empty P row);
//   * HEM (src/amg_parameters.py:60): heavy-edge matching in natural row order, the
//     partner is the unmatched neighbour with the largest |a_ij| (first one on ties);
//     rows left without a free partner join the aggregate of their heaviest neighbour;
//   * VMB (src/amg_parameters.py:16): Vanek-Mandel-Brezina greedy aggregation with the
//     symmetric strength test a_ij^2 >= theta^2 |a_ii a_jj| and the max_aggregation cap;
//   * tentative P is boolean, R = P', A_{l+1} = P' A_l P (Galerkin);
//   * coarsening stops at coarse_dof rows or max_levels (src/amg_parameters.py:50,56);
//   * the coarsest operator is inverted densely (coarse_solver 32 = direct).
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <numeric>
#include <queue>

#include "hierarchy.h"

namespace mamg {

static inline bool row_isolated(const Csr& A, int i) {
  for (int p = A.ia[i]; p < A.ia[i + 1]; ++p)
    if (A.ja[p] != i && A.a[p] != 0.0) return false;
  return true;
}

// `part` (may be NULL): owner part of every row; matches are made inside a part only, so that the
// aggregates -- and with them restriction and prolongation -- never cross a partition boundary
// ("Aggregates never cross partitions on distributed levels (setup constraint)").
void aggregate_hem(const Csr& A, const int* part, std::vector<int>& agg, int& nc) {
  const int n = A.n;
  agg.assign(n, -2);
  nc = 0;
  std::vector<int> pending;
  for (int i = 0; i < n; ++i) {
    if (agg[i] != -2) continue;
    if (row_isolated(A, i)) { agg[i] = -1; continue; }
    int best = -1;
    double bw = 0.0;
    for (int p = A.ia[i]; p < A.ia[i + 1]; ++p) {
      int j = A.ja[p];
      if (j == i || agg[j] != -2) continue;
      if (part && part[j] != part[i]) continue;
      double w = std::fabs(A.a[p]);
      if (w > bw) { bw = w; best = j; }
    }
    if (best >= 0 && !row_isolated(A, best)) {
      agg[i] = agg[best] = nc++;
    } else {
      pending.push_back(i);
    }
  }
  for (int i : pending) {
    int best = -1;
    double bw = 0.0;
    for (int p = A.ia[i]; p < A.ia[i + 1]; ++p) {
      int j = A.ja[p];
      if (j == i || agg[j] < 0) continue;
      if (part && part[j] != part[i]) continue;
      double w = std::fabs(A.a[p]);
      if (w > bw) { bw = w; best = j; }
    }
    agg[i] = best >= 0 ? agg[best] : nc++;
  }
}

// The remaining aggregation_type values of haznics (src/amg_parameters.py:16 lists VMB, MIS, MWM, HEC; HEM is
// what the metric dicts use).  HAZmath's source is not available here, so these are the textbook algorithms
// their names stand for, frozen like every other choice of this setup:
//   HEC  heavy-edge coarsening: in natural order, a free row joins its heaviest neighbour -- a new pair if
//        that neighbour is free, else the neighbour's aggregate while it has room (max_aggregation)
//   MWM  maximal weighted matching: greedy over all couplings sorted by decreasing |a_ij| (ties: lower row,
//        lower column first); leftovers join the aggregate of their heaviest matched neighbour
//   MIS  aggregation around a greedy maximal independent set of the strong-coupling graph: every root
//        takes its free strong neighbours, remaining rows join the root aggregate they are most strongly tied to
void aggregate_hec(const Csr& A, const int* part, int max_agg, std::vector<int>& agg, int& nc) {
  const int n = A.n;
  agg.assign(n, -2);
  nc = 0;
  if (max_agg < 2) max_agg = 2;
  std::vector<int> size;
  for (int i = 0; i < n; ++i) {
    if (agg[i] != -2) continue;
    if (row_isolated(A, i)) { agg[i] = -1; continue; }
    int best = -1;
    double bw = 0.0;
    for (int p = A.ia[i]; p < A.ia[i + 1]; ++p) {
      const int j = A.ja[p];
      if (j == i || agg[j] == -1 || (part && part[j] != part[i])) continue;
      if (agg[j] >= 0 && size[agg[j]] >= max_agg) continue;
      if (agg[j] == -2 && row_isolated(A, j)) continue;
      const double w = std::fabs(A.a[p]);
      if (w > bw) { bw = w; best = j; }
    }
    if (best < 0) { agg[i] = nc++; size.push_back(1); }
    else if (agg[best] == -2) { agg[i] = agg[best] = nc++; size.push_back(2); }
    else { agg[i] = agg[best]; ++size[agg[best]]; }
  }
}

void aggregate_mwm(const Csr& A, const int* part, std::vector<int>& agg, int& nc) {
  const int n = A.n;
  agg.assign(n, -2);
  nc = 0;
  struct Edge { double w; int i, j; };
  std::vector<Edge> edges;
  for (int i = 0; i < n; ++i) {
    if (row_isolated(A, i)) { agg[i] = -1; continue; }
    for (int p = A.ia[i]; p < A.ia[i + 1]; ++p) {
      const int j = A.ja[p];
      if (j <= i || A.a[p] == 0.0 || (part && part[j] != part[i])) continue;
      edges.push_back({std::fabs(A.a[p]), i, j});
    }
  }
  std::stable_sort(edges.begin(), edges.end(), [](const Edge& a, const Edge& b) { return a.w > b.w; });
  for (const Edge& e : edges)
    if (agg[e.i] == -2 && agg[e.j] == -2) agg[e.i] = agg[e.j] = nc++;
  std::vector<int> frozen(agg);
  for (int i = 0; i < n; ++i) {
    if (agg[i] != -2) continue;
    int best = -1;
    double bw = 0.0;
    for (int p = A.ia[i]; p < A.ia[i + 1]; ++p) {
      const int j = A.ja[p];
      if (j == i || frozen[j] < 0 || (part && part[j] != part[i])) continue;
      const double w = std::fabs(A.a[p]);
      if (w > bw) { bw = w; best = j; }
    }
    agg[i] = best >= 0 ? frozen[best] : nc++;
  }
}

void aggregate_mis(const Csr& A, const int* part, double strong, std::vector<int>& agg, int& nc) {
  const int n = A.n;
  std::vector<double> diag(n, 0.0);
  for (int i = 0; i < n; ++i)
    for (int p = A.ia[i]; p < A.ia[i + 1]; ++p)
      if (A.ja[p] == i) diag[i] = A.a[p];
  const double s2 = strong * strong;
  auto is_strong = [&](int i, int p) {
    const int j = A.ja[p];
    if (j == i || A.a[p] == 0.0 || (part && part[j] != part[i])) return false;
    return A.a[p] * A.a[p] >= s2 * std::fabs(diag[i] * diag[j]);
  };
  agg.assign(n, -2);
  nc = 0;
  std::vector<char> state(n, 0);   // 0 free, 1 root, 2 excluded
  for (int i = 0; i < n; ++i) {
    bool any = false;
    for (int p = A.ia[i]; p < A.ia[i + 1] && !any; ++p) any = is_strong(i, p);
    if (!any) { agg[i] = -1; state[i] = 2; }
  }
  for (int i = 0; i < n; ++i) {
    if (state[i]) continue;
    state[i] = 1;
    agg[i] = nc;
    for (int p = A.ia[i]; p < A.ia[i + 1]; ++p)
      if (is_strong(i, p)) {
        const int j = A.ja[p];
        if (state[j] == 0) state[j] = 2;
        if (agg[j] == -2) agg[j] = nc;
      }
    ++nc;
  }
  std::vector<int> frozen(agg);
  for (int i = 0; i < n; ++i) {
    if (agg[i] != -2) continue;
    int best = -1;
    double bw = 0.0;
    for (int p = A.ia[i]; p < A.ia[i + 1]; ++p) {
      if (!is_strong(i, p) || frozen[A.ja[p]] < 0) continue;
      const double w = std::fabs(A.a[p]);
      if (w > bw) { bw = w; best = A.ja[p]; }
    }
    agg[i] = best >= 0 ? frozen[best] : nc++;
  }
}

void aggregate_vmb(const Csr& A, const int* part, double strong, int max_agg, std::vector<int>& agg, int& nc) {
  const int n = A.n;
  std::vector<double> diag(n, 0.0);
  for (int i = 0; i < n; ++i)
    for (int p = A.ia[i]; p < A.ia[i + 1]; ++p)
      if (A.ja[p] == i) diag[i] = A.a[p];
  const double s2 = strong * strong;
  auto is_strong = [&](int i, int p) {
    int j = A.ja[p];
    if (j == i || A.a[p] == 0.0) return false;
    if (part && part[j] != part[i]) return false;
    return A.a[p] * A.a[p] >= s2 * std::fabs(diag[i] * diag[j]);
  };
  agg.assign(n, -2);
  nc = 0;
  if (max_agg < 2) max_agg = 2;
  // rows without strong neighbours are left out
  for (int i = 0; i < n; ++i) {
    bool any = false;
    for (int p = A.ia[i]; p < A.ia[i + 1] && !any; ++p) any = is_strong(i, p);
    if (!any) agg[i] = -1;
  }
  // pass 1: seed aggregates whose whole strong neighbourhood is free
  for (int i = 0; i < n; ++i) {
    if (agg[i] != -2) continue;
    bool free_nbhd = true;
    for (int p = A.ia[i]; p < A.ia[i + 1] && free_nbhd; ++p)
      if (is_strong(i, p) && agg[A.ja[p]] != -2 && agg[A.ja[p]] != -1) free_nbhd = false;
    if (!free_nbhd) continue;
    int cnt = 1;
    agg[i] = nc;
    for (int p = A.ia[i]; p < A.ia[i + 1] && cnt < max_agg; ++p)
      if (is_strong(i, p) && agg[A.ja[p]] == -2) { agg[A.ja[p]] = nc; ++cnt; }
    ++nc;
  }
  // pass 2: attach leftovers to the pass-1 aggregate they are most strongly tied to
  std::vector<int> frozen(agg);
  for (int i = 0; i < n; ++i) {
    if (agg[i] != -2) continue;
    int best = -1;
    double bw = 0.0;
    for (int p = A.ia[i]; p < A.ia[i + 1]; ++p) {
      if (!is_strong(i, p) || frozen[A.ja[p]] < 0) continue;
      double w = std::fabs(A.a[p]);
      if (w > bw) { bw = w; best = A.ja[p]; }
    }
    if (best >= 0) agg[i] = frozen[best];
  }
  // pass 3: what is still free forms new aggregates with its free strong neighbours
  for (int i = 0; i < n; ++i) {
    if (agg[i] != -2) continue;
    int cnt = 1;
    agg[i] = nc;
    for (int p = A.ia[i]; p < A.ia[i + 1] && cnt < max_agg; ++p)
      if (is_strong(i, p) && agg[A.ja[p]] == -2) { agg[A.ja[p]] = nc; ++cnt; }
    ++nc;
  }
}

// A_c = P' A P for boolean P (aggregate-specialised RAP): coarse row I is the sum of the
// fine rows of aggregate I with columns mapped through agg; columns sorted ascending.
void galerkin_ua(const Csr& A, const std::vector<int>& agg, int nc, Csr& Ac) {
  const int n = A.n;
  std::vector<int> cptr(nc + 1, 0), cidx;
  for (int i = 0; i < n; ++i)
    if (agg[i] >= 0) ++cptr[agg[i] + 1];
  for (int I = 0; I < nc; ++I) cptr[I + 1] += cptr[I];
  cidx.resize(cptr[nc]);
  {
    std::vector<int> fill(cptr.begin(), cptr.end() - 1);
    for (int i = 0; i < n; ++i)
      if (agg[i] >= 0) cidx[fill[agg[i]]++] = i;
  }
  Ac.n = Ac.m = nc;
  Ac.ia.assign(nc + 1, 0);
#pragma omp parallel
  {
    std::vector<int> mark(nc, -1);
#pragma omp for schedule(static)
    for (int I = 0; I < nc; ++I) {
      int cnt = 0;
      for (int q = cptr[I]; q < cptr[I + 1]; ++q) {
        int i = cidx[q];
        for (int p = A.ia[i]; p < A.ia[i + 1]; ++p) {
          int J = agg[A.ja[p]];
          if (J < 0) continue;
          if (mark[J] != I) { mark[J] = I; ++cnt; }
        }
      }
      Ac.ia[I + 1] = cnt;
    }
  }
  for (int I = 0; I < nc; ++I) Ac.ia[I + 1] += Ac.ia[I];
  Ac.ja.resize(Ac.ia[nc]);
  Ac.a.resize(Ac.ia[nc]);
#pragma omp parallel
  {
    std::vector<int> pos(nc, -1);
    std::vector<std::pair<int, double>> rowbuf;
#pragma omp for schedule(static)
    for (int I = 0; I < nc; ++I) {
      const int base = Ac.ia[I];
      int cnt = 0;
      for (int q = cptr[I]; q < cptr[I + 1]; ++q) {
        int i = cidx[q];
        for (int p = A.ia[i]; p < A.ia[i + 1]; ++p) {
          int J = agg[A.ja[p]];
          if (J < 0) continue;
          if (pos[J] < base) { pos[J] = base + cnt; Ac.ja[base + cnt] = J; Ac.a[base + cnt] = A.a[p]; ++cnt; }
          else Ac.a[pos[J]] += A.a[p];
        }
      }
      rowbuf.resize(cnt);
      for (int k = 0; k < cnt; ++k) rowbuf[k] = {Ac.ja[base + k], Ac.a[base + k]};
      std::sort(rowbuf.begin(), rowbuf.end(),
                [](const std::pair<int, double>& x, const std::pair<int, double>& y) { return x.first < y.first; });
      for (int k = 0; k < cnt; ++k) { Ac.ja[base + k] = rowbuf[k].first; Ac.a[base + k] = rowbuf[k].second; pos[rowbuf[k].first] = -1; }
    }
  }
}

// FEniCS/PETSc hand over the full P1 pattern; entries that are exactly zero (stiffness couplings along
// the diagonals of the Kuhn mesh, eliminated Dirichlet couplings) contribute nothing to any sum and
// the setup already ignores them.  Removing them is bit-neutral for every kernel.  The diagonal stays.
void csr_drop_zeros(Csr& A) {
  const int n = A.n;
  std::vector<int> ia2(n + 1, 0);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) {
    int cnt = 0;
    for (int p = A.ia[i]; p < A.ia[i + 1]; ++p) cnt += (A.a[p] != 0.0 || A.ja[p] == i);
    ia2[i + 1] = cnt;
  }
  for (int i = 0; i < n; ++i) ia2[i + 1] += ia2[i];
  if (ia2[n] == A.ia[n]) return;   // nothing to drop
  bigvec<int> ja2(ia2[n]);
  bigvec<double> a2(ia2[n]);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) {
    int k = ia2[i];
    for (int p = A.ia[i]; p < A.ia[i + 1]; ++p)
      if (A.a[p] != 0.0 || A.ja[p] == i) { ja2[k] = A.ja[p]; a2[k] = A.a[p]; ++k; }
  }
  A.ia.swap(ia2);
  A.ja.swap(ja2);
  A.a.swap(a2);
}

void csr_transpose(const Csr& A, Csr& At) {
  At.n = A.m;
  At.m = A.n;
  At.ia.assign(At.n + 1, 0);
  for (int p = 0; p < A.nnz(); ++p) ++At.ia[A.ja[p] + 1];
  for (int i = 0; i < At.n; ++i) At.ia[i + 1] += At.ia[i];
  At.ja.resize(A.nnz());
  At.a.resize(A.nnz());
  std::vector<int> fill(At.ia.begin(), At.ia.end() - 1);
  for (int i = 0; i < A.n; ++i)
    for (int p = A.ia[i]; p < A.ia[i + 1]; ++p) {
      const int q = fill[A.ja[p]]++;
      At.ja[q] = i;
      At.a[q] = A.a[p];
    }
}

// C = A B (Gustavson, row by row; columns of every row sorted ascending)
void csr_multiply(const Csr& A, const Csr& B, Csr& C) {
  C.n = A.n;
  C.m = B.m;
  C.ia.assign(A.n + 1, 0);
  std::vector<std::vector<std::pair<int, double>>> rows(A.n);
#pragma omp parallel
  {
    std::vector<int> pos(B.m, -1);
    std::vector<std::pair<int, double>> acc;
#pragma omp for schedule(dynamic, 256)
    for (int i = 0; i < A.n; ++i) {
      acc.clear();
      for (int p = A.ia[i]; p < A.ia[i + 1]; ++p) {
        const int k = A.ja[p];
        const double v = A.a[p];
        for (int q = B.ia[k]; q < B.ia[k + 1]; ++q) {
          const int j = B.ja[q];
          if (pos[j] < 0) { pos[j] = (int)acc.size(); acc.push_back({j, v * B.a[q]}); }
          else acc[pos[j]].second += v * B.a[q];
        }
      }
      for (auto& e : acc) pos[e.first] = -1;
      std::sort(acc.begin(), acc.end(),
                [](const std::pair<int, double>& x, const std::pair<int, double>& y) { return x.first < y.first; });
      rows[i] = acc;
      C.ia[i + 1] = (int)acc.size();
    }
  }
  {
    long long tot = 0;
    for (int i = 0; i < A.n; ++i) tot += C.ia[i + 1];
    if (tot > 0x7fffffffLL) throw std::runtime_error("sparse product has " + std::to_string(tot) + " entries: exceeds int32 indexing");
  }
  for (int i = 0; i < A.n; ++i) C.ia[i + 1] += C.ia[i];
  C.ja.resize(C.ia[A.n]);
  C.a.resize(C.ia[A.n]);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < A.n; ++i)
    for (size_t k = 0; k < rows[i].size(); ++k) {
      C.ja[C.ia[i] + k] = rows[i][k].first;
      C.a[C.ia[i] + k] = rows[i][k].second;
    }
}

// SA_AMG (src/input_metric.dat:68): P = (I - omega D^{-1} A) P_tent with the boolean tentative
// prolongator of the aggregates (near-kernel = constants) and the FASP-lineage default
// omega = tentative_smooth = 0.67; strong_coupled = 0.0 in that file means no filtering of A.
void smoothed_prolongator(const Csr& A, const std::vector<int>& agg, int nc, double omega, Csr& P) {
  const int n = A.n;
  P.n = n;
  P.m = nc;
  P.ia.assign(n + 1, 0);
  std::vector<std::vector<std::pair<int, double>>> rows(n);
#pragma omp parallel
  {
    std::vector<int> pos(nc, -1);
    std::vector<std::pair<int, double>> acc;
#pragma omp for schedule(dynamic, 1024)
    for (int i = 0; i < n; ++i) {
      acc.clear();
      double d = 1.0;
      for (int p = A.ia[i]; p < A.ia[i + 1]; ++p)
        if (A.ja[p] == i) d = A.a[p];
      auto add = [&](int J, double v) {
        if (J < 0) return;
        if (pos[J] < 0) { pos[J] = (int)acc.size(); acc.push_back({J, v}); }
        else acc[pos[J]].second += v;
      };
      add(agg[i], 1.0);
      if (agg[i] >= 0)   // rows outside every aggregate (Dirichlet) keep an empty P row
        for (int p = A.ia[i]; p < A.ia[i + 1]; ++p) add(agg[A.ja[p]], -omega * A.a[p] / d);
      for (auto& e : acc) pos[e.first] = -1;
      std::sort(acc.begin(), acc.end(),
                [](const std::pair<int, double>& x, const std::pair<int, double>& y) { return x.first < y.first; });
      rows[i] = acc;
      P.ia[i + 1] = (int)acc.size();
    }
  }
  for (int i = 0; i < n; ++i) P.ia[i + 1] += P.ia[i];
  P.ja.resize(P.ia[n]);
  P.a.resize(P.ia[n]);
  for (int i = 0; i < n; ++i)
    for (size_t k = 0; k < rows[i].size(); ++k) {
      P.ja[P.ia[i] + k] = rows[i][k].first;
      P.a[P.ia[i] + k] = rows[i][k].second;
    }
}

// Greedy multicolouring in natural row order over the nonzero off-diagonal couplings:
// the fixed ordering the Gauss-Seidel sweeps use on the device AND in the oracle
// (north_star: "Gauss-Seidel via a fixed multicolour ordering applied identically in
// the reference comparison").  Rows that the point smoother never touches (`skip`: the Schwarz
// seeds) are left out of the graph: they all get colour 0, which no sweep launches, and the rows
// that are smoothed share colours 1..k computed on their own subgraph.  On the device this keeps the
// Schwarz-only rows in one block in natural order (contiguous x-runs for the patch gathers) and
// gives the smoothed rows fewer, fully populated colour blocks.
void multicolor_greedy(const Csr& A, const std::vector<uint8_t>& skip, std::vector<int>& color, int& ncolors) {
  const int n = A.n;
  color.assign(n, -1);
  bool any_skip = false;
  for (int i = 0; i < n && !skip.empty(); ++i) any_skip = any_skip || skip[i];
  const int base = any_skip ? 1 : 0;
  ncolors = base;
  std::vector<int> forbid;  // forbid[c] == i  <=> colour c taken by a neighbour of row i
  for (int i = 0; i < n; ++i) {
    if (any_skip && skip[i]) continue;
    for (int p = A.ia[i]; p < A.ia[i + 1]; ++p) {
      int j = A.ja[p];
      if (j == i || A.a[p] == 0.0 || color[j] < 0) continue;
      if ((int)forbid.size() <= color[j]) forbid.resize(color[j] + 1, -1);
      forbid[color[j]] = i;
    }
    int c = base;
    while (c < (int)forbid.size() && forbid[c] == i) ++c;
    color[i] = c;
    if (c + 1 > ncolors) ncolors = c + 1;
    if ((int)forbid.size() < ncolors) forbid.resize(ncolors, -1);
  }
  if (any_skip)
    for (int i = 0; i < n; ++i)
      if (skip[i]) color[i] = 0;
}

// Schwarz blocks: seed + Schwarz_maxlvl graph rings (breadth first over nonzero couplings),
// at most Schwarz_mmsize dofs (src/amg_parameters.py:83-85); dofs sorted ascending.
void schwarz_patches(const Csr& A, const int* seeds, int nseeds, int maxlvl, int mmsize,
                     SchwarzPatches& out) {
  out.ptr.assign(nseeds + 1, 0);
  out.dofs.clear();
  out.seed.assign(seeds, seeds + nseeds);
  out.max_size = 0;
  if (mmsize < 1) mmsize = 1;
  // the breadth-first searches are independent: every thread takes one contiguous range of seeds
  // (static schedule) and the per-thread lists are concatenated in thread order
  int nthreads = 1;
#ifdef _OPENMP
  // every thread owns an n-sized marker array: only as many threads as the searches can pay for
  // (EMI: 0.4 % of the dofs are seeds -> one thread; bidomain: half of them -> all threads)
  nthreads = (int)std::max<long long>(1, std::min<long long>(omp_get_max_threads(), 32LL * nseeds / std::max(1, A.n)));
#endif
  std::vector<std::vector<int>> part(nthreads);
  int max_size = 0;
#pragma omp parallel num_threads(nthreads) reduction(max : max_size)
  {
    int tid = 0;
#ifdef _OPENMP
    tid = omp_get_thread_num();
#endif
    std::vector<int> mark(A.n, -1), cur, nxt, blk;
    std::vector<int>& mine = part[tid];
#pragma omp for schedule(static)
    for (int s = 0; s < nseeds; ++s) {
      const int seed = seeds[s];
      blk.assign(1, seed);
      mark[seed] = s;
      cur.assign(1, seed);
      for (int ring = 0; ring < maxlvl && (int)blk.size() < mmsize; ++ring) {
        nxt.clear();
        for (int i : cur) {
          for (int p = A.ia[i]; p < A.ia[i + 1]; ++p) {
            int j = A.ja[p];
            if (A.a[p] == 0.0 || mark[j] == s) continue;
            if ((int)blk.size() >= mmsize) break;
            mark[j] = s;
            blk.push_back(j);
            nxt.push_back(j);
          }
        }
        cur.swap(nxt);
      }
      std::sort(blk.begin(), blk.end());
      mine.insert(mine.end(), blk.begin(), blk.end());
      out.ptr[s + 1] = (int)blk.size();
      max_size = std::max(max_size, (int)blk.size());
    }
  }
  out.max_size = max_size;
  {
    long long tot = 0;
    for (int s = 0; s < nseeds; ++s) tot += out.ptr[s + 1];
    if (tot > 0x7fffffffLL) throw std::runtime_error("Schwarz patches: " + std::to_string(tot) + " patch dofs exceed int32 indexing");
  }
  for (int s = 0; s < nseeds; ++s) out.ptr[s + 1] += out.ptr[s];
  out.dofs.reserve(out.ptr[nseeds]);
  for (int t = 0; t < nthreads; ++t) out.dofs.insert(out.dofs.end(), part[t].begin(), part[t].end());
}

// Conflict colouring of the patches, greedy in patch order.  Two patches conflict when one
// touches the other's dofs or their matrix neighbours (then their multiplicative updates do
// not commute); patches of one colour can be solved concurrently with a result identical to
// visiting them one after another.
void schwarz_color(const Csr& A, SchwarzPatches& sw) {
  const int np = sw.npatch();
  sw.color.assign(np, 0);
  sw.ncolors = 0;
  if (np == 0) return;
  // reach[j] = colours of the patches that contain j or a dof coupled to j (the pattern of A is
  // symmetric), so the colours forbidden to a patch are the union of reach over its own dofs:
  // s word-vector ORs per patch to query, one single-word OR per row entry to record
  int W = 2;  // 64-bit words per dof mask, grown on demand
  std::vector<uint64_t> reach((size_t)A.n * W, 0), forb;
  for (int p = 0; p < np; ++p) {
    forb.assign(W, 0);
    for (int q = sw.ptr[p]; q < sw.ptr[p + 1]; ++q) {
      const uint64_t* mi = &reach[(size_t)sw.dofs[q] * W];
      for (int w = 0; w < W; ++w) forb[w] |= mi[w];
    }
    int c = -1;
    for (int w = 0; w < W && c < 0; ++w)
      if (~forb[w]) c = w * 64 + __builtin_ctzll(~forb[w]);
    if (c < 0) {  // all W*64 colours taken: widen the masks
      const int W2 = W * 2;
      std::vector<uint64_t> m2((size_t)A.n * W2, 0);
      for (size_t i = 0; i < (size_t)A.n; ++i)
        for (int w = 0; w < W; ++w) m2[i * W2 + w] = reach[i * W + w];
      reach.swap(m2);
      c = W * 64;
      W = W2;
    }
    sw.color[p] = c;
    sw.ncolors = std::max(sw.ncolors, c + 1);
    const size_t cw = c / 64;
    const uint64_t bit = 1ull << (c % 64);
    for (int q = sw.ptr[p]; q < sw.ptr[p + 1]; ++q) {
      const int i = sw.dofs[q];
      reach[(size_t)i * W + cw] |= bit;
      for (int e = A.ia[i]; e < A.ia[i + 1]; ++e)
        if (A.a[e] != 0.0) reach[(size_t)A.ja[e] * W + cw] |= bit;
    }
  }
}

// Dense inverse of the coarsest operator by Gauss-Jordan with partial pivoting in extended
// precision (stands in for the UMFPACK factorisation, coarse_solver 32).
bool dense_inverse(const Csr& A, std::vector<double>& inv) {
  const int n = A.n;
  std::vector<long double> M((size_t)n * 2 * n, 0.0L);
  for (int i = 0; i < n; ++i) {
    for (int p = A.ia[i]; p < A.ia[i + 1]; ++p) M[(size_t)i * 2 * n + A.ja[p]] += A.a[p];
    M[(size_t)i * 2 * n + n + i] = 1.0L;
  }
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int r = c + 1; r < n; ++r)
      if (fabsl(M[(size_t)r * 2 * n + c]) > fabsl(M[(size_t)piv * 2 * n + c])) piv = r;
    if (M[(size_t)piv * 2 * n + c] == 0.0L) return false;
    if (piv != c)
      for (int k = 0; k < 2 * n; ++k) std::swap(M[(size_t)piv * 2 * n + k], M[(size_t)c * 2 * n + k]);
    long double d = 1.0L / M[(size_t)c * 2 * n + c];
    for (int k = 0; k < 2 * n; ++k) M[(size_t)c * 2 * n + k] *= d;
    for (int r = 0; r < n; ++r) {
      if (r == c) continue;
      long double f = M[(size_t)r * 2 * n + c];
      if (f == 0.0L) continue;
      for (int k = c; k < 2 * n; ++k) M[(size_t)r * 2 * n + k] -= f * M[(size_t)c * 2 * n + k];
    }
  }
  inv.resize((size_t)n * n);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) inv[(size_t)i * n + j] = (double)M[(size_t)i * 2 * n + n + j];
  return true;
}

static void greedy_mis(const Csr& A, std::vector<int>& seeds) {
  std::vector<char> state(A.n, 0);  // 0 free, 1 in MIS, 2 excluded
  seeds.clear();
  for (int i = 0; i < A.n; ++i) {
    if (state[i]) continue;
    state[i] = 1;
    seeds.push_back(i);
    for (int p = A.ia[i]; p < A.ia[i + 1]; ++p)
      if (A.a[p] != 0.0 && state[A.ja[p]] == 0) state[A.ja[p]] = 2;
  }
}

// the option space of the parameter dict (src/amg_parameters.py) that setup, import and the device cycle accept
bool validate_params(const mamg_params& prm, std::string& err) {
  if (prm.AMG_type != MAMG_UA_AMG && prm.AMG_type != MAMG_SA_AMG) { err = "AMG_type: only UA_AMG and SA_AMG are implemented"; return false; }
  switch (prm.cycle_type) {
    case MAMG_V_CYCLE: case MAMG_W_CYCLE: case MAMG_AMLI_CYCLE: case MAMG_NL_AMLI_CYCLE: break;
    case MAMG_ADD_CYCLE:
      if (prm.maxit > 1) { err = "cycle_type ADD_CYCLE: the additive cycle is applied once per call (maxit 1)"; return false; }
      break;
    default: err = "cycle_type: unknown value"; return false;
  }
  if (prm.cycle_type == MAMG_AMLI_CYCLE && (prm.amli_degree < 0 || prm.amli_degree > 15)) {
    err = "amli_degree: 0..15";
    return false;
  }
  switch (prm.aggregation_type) {
    case MAMG_HEM: case MAMG_VMB: case MAMG_HEC: case MAMG_MWM: case MAMG_MIS: break;
    default: err = "aggregation_type: unknown value"; return false;
  }
  switch (prm.smoother) {
    case MAMG_SMOOTHER_JACOBI: case MAMG_SMOOTHER_GS: case MAMG_SMOOTHER_SGS:
    case MAMG_SMOOTHER_SOR: case MAMG_SMOOTHER_SSOR: case MAMG_SMOOTHER_L1DIAG: break;
    default: err = "smoother: only JACOBI, GS, SGS, SOR, SSOR, L1DIAG are implemented"; return false;
  }
  // coarse_solver / Schwarz_blksolver 0 ("iterative", src/amg_parameters.py:14,43): upstream these iterate to
  // tol*1e-4 on the coarsest operator / the patch blocks; the dense inverses used here are the limit of that
  // iteration, so both values are served by the same path
  if (prm.coarse_solver != MAMG_SOLVER_UMFPACK && prm.coarse_solver != MAMG_SOLVER_DEFAULT) {
    err = "coarse_solver: 32 (direct) or 0 (iterative)";
    return false;
  }
  if (prm.Schwarz_levels > 0 && prm.Schwarz_blksolver != MAMG_SOLVER_UMFPACK && prm.Schwarz_blksolver != MAMG_SOLVER_DEFAULT) {
    err = "Schwarz_blksolver: 32 (direct) or 0 (iterative)";
    return false;
  }
  return true;
}

bool build_hierarchy(const mamg_params& prm, Csr&& A0, const int* idofs, int n_idofs,
                     const int* part, int nparts, Hierarchy& H, std::string& err) {
  auto t0 = std::chrono::steady_clock::now();
  H.prm = prm;
  H.nparts = part ? std::max(1, nparts) : 1;
  H.lv.clear();
  H.lv.emplace_back();
  H.lv[0].A = std::move(A0);
  if (part) H.lv[0].part.assign(part, part + H.lv[0].A.n);
  if (!validate_params(prm, err)) return false;
  const int max_levels = std::max(1, prm.max_levels);
  // Store no explicit zeros on any level (MAMG_DROP_ZEROS=0 keeps the caller's full pattern): a +0*x
  // term never changes a sum, aggregation / colouring / patch search ignore zero couplings anyway, and
  // on the Kuhn-mesh P1 systems up to half of the stored entries are exact zeros.  nnz_structural
  // remembers the pattern that was handed in / that the Galerkin product produced.
  const bool drop_zeros = !(getenv("MAMG_DROP_ZEROS") && atoi(getenv("MAMG_DROP_ZEROS")) == 0);
  H.lv[0].nnz_structural = H.lv[0].A.nnz();
  if (drop_zeros) csr_drop_zeros(H.lv[0].A);
  const bool timing = getenv("MAMG_SETUP_TIMING") != nullptr;   // per-phase seconds on stderr
  auto tp = std::chrono::steady_clock::now();
  auto lap = [&](const char* what, int lev) {
    if (!timing) return;
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[mamg setup] level %d %-16s %.3f s\n", lev, what, std::chrono::duration<double>(now - tp).count());
    tp = now;
  };
  std::vector<int> seeds(idofs, idofs + n_idofs);
  const bool metric = n_idofs > 0;
  int l = 0;
  while (true) {
    Level& L = H.lv[l];
    const int n = L.A.n;
    const bool last = !(n > prm.coarse_dof && l < max_levels - 1);
    if (!last && l < prm.Schwarz_levels) {
      if (!metric) greedy_mis(L.A, seeds);
      schwarz_patches(L.A, seeds.data(), (int)seeds.size(), prm.Schwarz_maxlvl, prm.Schwarz_mmsize, L.sw);
      lap("schwarz_patches", l);
      schwarz_color(L.A, L.sw);
      lap("schwarz_color", l);
      L.gs_skip.assign(n, metric ? 0 : 1);
      if (metric) for (int s : seeds) L.gs_skip[s] = 1;
    }
    if (last) break;
    const int* lpart = L.part.empty() ? nullptr : L.part.data();
    switch (prm.aggregation_type) {
      case MAMG_HEM: aggregate_hem(L.A, lpart, L.agg, L.nc); break;
      case MAMG_HEC: aggregate_hec(L.A, lpart, prm.max_aggregation, L.agg, L.nc); break;
      case MAMG_MWM: aggregate_mwm(L.A, lpart, L.agg, L.nc); break;
      case MAMG_MIS: aggregate_mis(L.A, lpart, prm.strong_coupled, L.agg, L.nc); break;
      default: aggregate_vmb(L.A, lpart, prm.strong_coupled, prm.max_aggregation, L.agg, L.nc);
    }
    lap("aggregate", l);
    if (L.nc == 0 || L.nc >= n) {  // no coarsening possible: this level becomes the coarsest
      L.agg.clear(); L.nc = 0; L.sw = SchwarzPatches(); L.gs_skip.clear();
      break;
    }
    multicolor_greedy(L.A, L.gs_skip, L.color, L.ncolors);
    lap("multicolor", l);
    H.lv.emplace_back();
    if (prm.AMG_type == MAMG_SA_AMG) {
      Level& F = H.lv[l];
      smoothed_prolongator(F.A, F.agg, F.nc, 0.67, F.P);
      csr_transpose(F.P, F.R);
      Csr AP;
      csr_multiply(F.A, F.P, AP);
      csr_multiply(F.R, AP, H.lv[l + 1].A);
    } else {
      galerkin_ua(H.lv[l].A, H.lv[l].agg, H.lv[l].nc, H.lv[l + 1].A);
    }
    H.lv[l + 1].nnz_structural = H.lv[l + 1].A.nnz();
    if (drop_zeros) csr_drop_zeros(H.lv[l + 1].A);
    lap("galerkin", l);
    if (!H.lv[l].part.empty()) {   // a coarse row belongs to the part of its members
      std::vector<int>& cp = H.lv[l + 1].part;
      cp.assign(H.lv[l].nc, 0);
      for (int i = 0; i < H.lv[l].A.n; ++i)
        if (H.lv[l].agg[i] >= 0) cp[H.lv[l].agg[i]] = H.lv[l].part[i];
    }
    if (metric && l + 1 < prm.Schwarz_levels) {  // carry the interface seeds to the next level
      std::vector<int> nxt;
      std::vector<char> seen(H.lv[l].nc, 0);
      for (int s : seeds) {
        int I = H.lv[l].agg[s];
        if (I >= 0 && !seen[I]) { seen[I] = 1; nxt.push_back(I); }
      }
      std::sort(nxt.begin(), nxt.end());
      seeds.swap(nxt);
    }
    ++l;
  }
  const Csr& Ac = H.lv.back().A;
  if (Ac.n > 8192) {
    err = "coarsest level has " + std::to_string(Ac.n) + " rows (> 8192): direct coarse solve refused";
    return false;
  }
  if (!dense_inverse(Ac, H.coarse_inv)) { err = "coarsest operator is singular"; return false; }
  H.setup_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return true;
}

}  // namespace mamg

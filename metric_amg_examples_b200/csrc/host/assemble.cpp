// Structured P1 systems of BASELINE.json's configs, assembled row by row.
//
// Restates, for uniform meshes only, what the reference obtains from
// dolfin + FEniCS_ii and then flattens with ii_convert:
//   bidomain  src/bidomain_2d.py:64-68 (blocks), :73,:93-97 (Dirichlet tags 1,2,
//             symmetric apply_bc), meshes src/utils.py:149-182
//   EMI       src/emi_2d.py:83-94 (trace coupling), :104-108,:122-126 (Dirichlet
//             tags 3 and 6), meshes src/utils.py:187-260
// Mesh: UnitSquareMesh(n,n) "right" diagonal / UnitCubeMesh(n,n,n) 6-tet split =
// Kuhn triangulation (every simplex is a monotone path 0 -> e_a -> e_a+e_b -> 1).
// Dofs are lexicographic (x fastest); the monolithic order is [W0 dofs; W1 dofs]
// (src/bidomain_2d.py:210-211).  The sparsity is the structural P1 pattern
// (entries that are zero by geometry or zeroed by the boundary conditions are
// kept, as PETSc keeps them), which reproduces the nnz counts of SURVEY 8a.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <numeric>

#include "hierarchy.h"

namespace mamg {
namespace {

constexpr int MAXS = 27;

struct CornerEntry { int slot; double K, M; };

// Per-cube P1 matrices of the Kuhn split, organised by the corner role of the row vertex.
struct Stencil {
  int dim = 0;
  int ncell[3] = {1, 1, 1};
  std::vector<CornerEntry> role[8];  // role[c]: contributions of one cube where the row vertex is corner c

  Stencil(int d, const int* nc, const double* h) : dim(d) {
    for (int a = 0; a < d; ++a) ncell[a] = nc[a];
    const int ncorner = 1 << d;
    double Kc[8][8] = {{0}}, Mc[8][8] = {{0}};
    bool share[8][8] = {{false}};
    int perm[3] = {0, 1, 2};
    std::sort(perm, perm + d);
    double fact = 1;
    for (int k = 2; k <= d; ++k) fact *= k;
    do {
      int v[4];
      v[0] = 0;
      for (int k = 0; k < d; ++k) v[k + 1] = v[k] | (1 << perm[k]);
      // T columns = p_k - p_0 ; gradients of lambda_k (k>=1) = rows of T^{-1}
      double T[3][3] = {{0}}, Ti[3][3] = {{0}};
      for (int k = 0; k < d; ++k)
        for (int a = 0; a < d; ++a) T[a][k] = ((v[k + 1] >> a) & 1) * h[a];
      // Gauss-Jordan on [T | I]
      double W[3][6];
      for (int r = 0; r < d; ++r)
        for (int c = 0; c < 2 * d; ++c) W[r][c] = c < d ? T[r][c] : (c - d == r ? 1.0 : 0.0);
      double det = 1;
      for (int c = 0; c < d; ++c) {
        int p = c;
        for (int r = c + 1; r < d; ++r)
          if (std::fabs(W[r][c]) > std::fabs(W[p][c])) p = r;
        if (p != c) { for (int k = 0; k < 2 * d; ++k) std::swap(W[p][k], W[c][k]); det = -det; }
        det *= W[c][c];
        double piv = W[c][c];
        for (int k = 0; k < 2 * d; ++k) W[c][k] /= piv;
        for (int r = 0; r < d; ++r)
          if (r != c) {
            double f = W[r][c];
            for (int k = 0; k < 2 * d; ++k) W[r][k] -= f * W[c][k];
          }
      }
      for (int r = 0; r < d; ++r)
        for (int c = 0; c < d; ++c) Ti[r][c] = W[r][c + d];
      double vol = std::fabs(det) / fact;
      double g[4][3] = {{0}};
      for (int k = 1; k <= d; ++k)
        for (int a = 0; a < d; ++a) { g[k][a] = Ti[k - 1][a]; g[0][a] -= Ti[k - 1][a]; }
      double mscale = vol / ((d + 1) * (d + 2));
      for (int i = 0; i <= d; ++i)
        for (int j = 0; j <= d; ++j) {
          double dotg = 0;
          for (int a = 0; a < d; ++a) dotg += g[i][a] * g[j][a];
          Kc[v[i]][v[j]] += vol * dotg;
          Mc[v[i]][v[j]] += mscale * (i == j ? 2.0 : 1.0);
          share[v[i]][v[j]] = true;
        }
    } while (std::next_permutation(perm, perm + d));
    for (int c = 0; c < ncorner; ++c)
      for (int e = 0; e < ncorner; ++e)
        if (share[c][e]) {
          int slot = 0, mul = 1;
          for (int a = 0; a < d; ++a) {
            int off = ((e >> a) & 1) - ((c >> a) & 1);
            slot += (off + 1) * mul;
            mul *= 3;
          }
          role[c].push_back({slot, Kc[c][e], Mc[c][e]});
        }
  }

  int nslots() const { int s = 1; for (int a = 0; a < dim; ++a) s *= 3; return s; }
  int nvert() const { int s = 1; for (int a = 0; a < dim; ++a) s *= ncell[a] + 1; return s; }

  // stencil row of the vertex with multi-index idx; returns number of present slots
  int row(const int* idx, bool* present, double* K, double* M) const {
    const int ns = nslots();
    for (int s = 0; s < ns; ++s) { present[s] = false; K[s] = 0; M[s] = 0; }
    const int ncorner = 1 << dim;
    for (int c = 0; c < ncorner; ++c) {
      bool ok = true;
      for (int a = 0; a < dim; ++a) {
        int o = idx[a] - ((c >> a) & 1);
        if (o < 0 || o > ncell[a] - 1) { ok = false; break; }
      }
      if (!ok) continue;
      for (const CornerEntry& e : role[c]) { present[e.slot] = true; K[e.slot] += e.K; M[e.slot] += e.M; }
    }
    int cnt = 0;
    for (int s = 0; s < ns; ++s) cnt += present[s];
    return cnt;
  }

  // lexicographic index of idx + offset encoded by slot
  int neighbor(const int* idx, int slot) const {
    int id = 0, mul = 1;
    for (int a = 0; a < dim; ++a) {
      int off = slot % 3 - 1;
      slot /= 3;
      id += (idx[a] + off) * mul;
      mul *= ncell[a] + 1;
    }
    return id;
  }
  void unflatten(int v, int* idx) const {
    for (int a = 0; a < dim; ++a) { idx[a] = v % (ncell[a] + 1); v /= ncell[a] + 1; }
    for (int a = dim; a < 3; ++a) idx[a] = 0;
  }
  int slot_offset(int slot, int axis) const {
    for (int a = 0; a < axis; ++a) slot /= 3;
    return slot % 3 - 1;
  }
};

}  // namespace

void p1_scalar(int dim, const int* ncell, const double* h, double cK, double cM, Csr& out) {
  Stencil st(dim, ncell, h);
  const int nv = st.nvert(), ns = st.nslots();
  out.n = out.m = nv;
  out.ia.assign(nv + 1, 0);
#pragma omp parallel for schedule(static)
  for (int v = 0; v < nv; ++v) {
    int idx[3]; bool pr[MAXS]; double K[MAXS], M[MAXS];
    st.unflatten(v, idx);
    out.ia[v + 1] = st.row(idx, pr, K, M);
  }
  for (int v = 0; v < nv; ++v) out.ia[v + 1] += out.ia[v];
  out.ja.resize(out.ia[nv]);
  out.a.resize(out.ia[nv]);
#pragma omp parallel for schedule(static)
  for (int v = 0; v < nv; ++v) {
    int idx[3]; bool pr[MAXS]; double K[MAXS], M[MAXS];
    st.unflatten(v, idx);
    st.row(idx, pr, K, M);
    int p = out.ia[v];
    for (int s = 0; s < ns; ++s)
      if (pr[s]) { out.ja[p] = st.neighbor(idx, s); out.a[p] = cK * K[s] + cM * M[s]; ++p; }
  }
}

namespace {

// Shared two-pass driver: `emit(row, cols, vals)` returns the entries of one monolithic row.
template <class RowFn>
void build_rows(int n, RowFn rowfn, Csr& out) {
  out.n = out.m = n;
  out.ia.assign(n + 1, 0);
#pragma omp parallel for schedule(static)
  for (int r = 0; r < n; ++r) {
    int cols[2 * MAXS]; double vals[2 * MAXS];
    out.ia[r + 1] = rowfn(r, cols, vals);
  }
  for (int r = 0; r < n; ++r) out.ia[r + 1] += out.ia[r];
  out.ja.resize(out.ia[n]);
  out.a.resize(out.ia[n]);
#pragma omp parallel for schedule(static)
  for (int r = 0; r < n; ++r) {
    int cols[2 * MAXS]; double vals[2 * MAXS];
    int k = rowfn(r, cols, vals);
    std::copy(cols, cols + k, out.ja.begin() + out.ia[r]);
    std::copy(vals, vals + k, out.a.begin() + out.ia[r]);
  }
}

}  // namespace

// [[k1 K + g M, -g M], [-g M, k2 K + g M]], Dirichlet on x in {0,1} (2-D) / z in {0,1} (3-D)
// for both fields, applied symmetrically (rows and columns zeroed, unit diagonal).
void assemble_bidomain(int dim, int n, double k1, double k2, double g, Csr& out) {
  int ncell[3] = {n, n, n};
  double h[3] = {1.0 / n, 1.0 / n, 1.0 / n};
  Stencil st(dim, ncell, h);
  const int nv = st.nvert(), ns = st.nslots();
  const int daxis = dim == 2 ? 0 : 2;  // src/utils.py:159-160 (tags 1,2: x) / :177-178 (tags 1,2: z)
  auto is_dir = [&](const int* idx) { return idx[daxis] == 0 || idx[daxis] == n; };
  auto rowfn = [&](int r, int* cols, double* vals) {
    const int f = r / nv, v = r % nv;
    int idx[3]; bool pr[MAXS]; double K[MAXS], M[MAXS];
    st.unflatten(v, idx);
    st.row(idx, pr, K, M);
    const bool drow = is_dir(idx);
    const double kap = f == 0 ? k1 : k2;
    int k = 0;
    for (int blk = 0; blk < 2; ++blk)
      for (int s = 0; s < ns; ++s) {
        if (!pr[s]) continue;
        int w = st.neighbor(idx, s);
        int nidx[3] = {0, 0, 0};
        st.unflatten(w, nidx);
        double val;
        if (drow) val = (blk == f && w == v) ? 1.0 : 0.0;
        else if (is_dir(nidx)) val = 0.0;
        else val = blk == f ? kap * K[s] + g * M[s] : -g * M[s];
        cols[k] = blk * nv + w;
        vals[k] = val;
        ++k;
      }
    return k;
  };
  build_rows(2 * nv, rowfn, out);
}

// EMI: Omega_1 = upper half (last axis >= 1/2), Omega_2 = lower half; P1 on each,
// coupled through the interface mass matrix M_G (trace operators are vertex selections
// because both halves come from one conforming parent mesh).
//   A00 = k1 K1 + g T1' M_G T1, A01 = -g T1' M_G T2, A11 = k2 K2 + g T2' M_G T2.
// Dirichlet: top of Omega_1 (tag 3) and bottom of Omega_2 (tag 6).
void assemble_emi(int dim, int n, double k1, double k2, double g, Csr& out) {
  const int half = n / 2;
  int ncell[3] = {n, n, n};
  ncell[dim - 1] = half;
  double h[3] = {1.0 / n, 1.0 / n, 1.0 / n};
  Stencil st(dim, ncell, h);       // same box shape for both halves
  Stencil sg(dim - 1, ncell, h);   // interface mesh (first dim-1 axes)
  const int nv = st.nvert(), ns = st.nslots(), nsg = sg.nslots();
  const int la = dim - 1;
  auto rowfn = [&](int r, int* cols, double* vals) {
    const int f = r / nv, v = r % nv;
    int idx[3]; bool pr[MAXS]; double K[MAXS], M[MAXS];
    st.unflatten(v, idx);
    st.row(idx, pr, K, M);
    const int dir_plane = f == 0 ? half : 0;   // local last-axis index of the Dirichlet plane
    const int ifc_plane = f == 0 ? 0 : half;   // local last-axis index of the interface plane
    const bool drow = idx[la] == dir_plane;
    const bool irow = idx[la] == ifc_plane;
    const double kap = f == 0 ? k1 : k2;
    bool prg[MAXS]; double Kg[MAXS], Mg[MAXS];
    if (irow) sg.row(idx, prg, Kg, Mg);
    int k = 0;
    auto own_block = [&]() {
      for (int s = 0; s < ns; ++s) {
        if (!pr[s]) continue;
        int w = st.neighbor(idx, s);
        int nidx[3] = {0, 0, 0};
        st.unflatten(w, nidx);
        double val;
        if (drow) val = w == v ? 1.0 : 0.0;
        else if (nidx[la] == dir_plane) val = 0.0;
        else {
          val = kap * K[s];
          if (irow && st.slot_offset(s, la) == 0) {
            // in-plane neighbour: add g * M_G entry (slot of the (dim-1)-stencil)
            int sgslot = 0, mul = 1, ss = s;
            for (int a = 0; a < dim - 1; ++a) { sgslot += (ss % 3) * mul; ss /= 3; mul *= 3; }
            val += g * Mg[sgslot];
          }
        }
        cols[k] = f * nv + w;
        vals[k] = val;
        ++k;
      }
    };
    auto cross_block = [&]() {
      if (!irow) return;
      const int other_plane = f == 0 ? half : 0;  // interface plane index in the other half
      for (int s = 0; s < nsg; ++s) {
        if (!prg[s]) continue;
        int oidx[3] = {0, 0, 0};
        int ss = s;
        for (int a = 0; a < dim - 1; ++a) { oidx[a] = idx[a] + ss % 3 - 1; ss /= 3; }
        oidx[la] = other_plane;
        int w = 0, mul = 1;
        for (int a = 0; a < dim; ++a) { w += oidx[a] * mul; mul *= ncell[a] + 1; }
        cols[k] = (1 - f) * nv + w;
        vals[k] = -g * Mg[s];
        ++k;
      }
    };
    if (f == 0) { own_block(); cross_block(); } else { cross_block(); own_block(); }
    return k;
  };
  build_rows(2 * nv, rowfn, out);
}

}  // namespace mamg

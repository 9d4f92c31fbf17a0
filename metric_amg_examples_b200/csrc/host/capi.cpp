// Host half of the C-ABI declared in include/mamg.h (setup, export, synthetic systems).
// The device half (mamg_to_device, mamg_apply, mamg_pcg, ...) is in csrc/cuda/device.cu.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <string>

#include "handle.h"

namespace mamg {
static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
}  // namespace mamg

using namespace mamg;

#define MAMG_TRY try {
#define MAMG_CATCH                                                   \
  }                                                                  \
  catch (const std::exception& e) { set_error(e.what()); return -2; } \
  catch (...) { set_error("unknown C++ exception"); return -2; }

extern "C" {

const char* mamg_last_error(void) { return g_err.c_str(); }
const char* mamg_version(void) { return "mamg 0.1 sm_100a"; }

int mamg_params_default(mamg_params* p) {
  if (!p) { set_error("params: NULL"); return -1; }
  std::memset(p, 0, sizeof(*p));
  // src/utils.py:60-82
  p->AMG_type = MAMG_UA_AMG;
  p->cycle_type = MAMG_W_CYCLE;
  p->max_levels = 20;
  p->maxit = 1;
  p->smoother = MAMG_SMOOTHER_SGS;
  p->relaxation = 1.2;
  p->presmooth_iter = 1;
  p->postsmooth_iter = 1;
  p->coarse_dof = 100;
  p->coarse_solver = MAMG_SOLVER_UMFPACK;
  p->coarse_scaling = MAMG_ON;
  p->aggregation_type = MAMG_HEM;
  p->strong_coupled = 0.1;
  p->max_aggregation = 100;
  p->amli_degree = 3;
  p->Schwarz_levels = 1;
  p->Schwarz_mmsize = 100;
  p->Schwarz_maxlvl = 2;
  p->Schwarz_type = MAMG_SCHWARZ_SYMMETRIC;
  p->Schwarz_blksolver = MAMG_SOLVER_UMFPACK;
  p->print_level = 0;
  p->nl_amli_krylov_type = MAMG_SOLVER_VFGMRES;
  return 0;
}

int mamg_setup(const mamg_params* p, int32_t n, const int32_t* indptr, const int32_t* indices,
               const double* data, int32_t n_idofs, const int32_t* idofs, mamg_handle* out) {
  return mamg_setup_partitioned(p, n, indptr, indices, data, n_idofs, idofs, nullptr, 1, out);
}

int mamg_setup_partitioned(const mamg_params* p, int32_t n, const int32_t* indptr, const int32_t* indices,
                           const double* data, int32_t n_idofs, const int32_t* idofs, const int32_t* part,
                           int32_t nparts, mamg_handle* out) {
  MAMG_TRY
  if (!p || !indptr || !indices || !data || !out) { set_error("setup: NULL argument"); return -1; }
  if (n <= 0) { set_error("setup: matrix has no rows"); return -1; }
  if (indptr[0] != 0) { set_error("setup: indptr[0] != 0"); return -1; }
  for (int i = 0; i < n; ++i)
    if (indptr[i + 1] < indptr[i]) { set_error("setup: indptr not monotone"); return -1; }
  const int nnz = indptr[n];
  if (n_idofs < 0 || (n_idofs > 0 && !idofs)) { set_error("setup: n_idofs > 0 but idofs is NULL"); return -1; }
  for (int q = 0; q < n_idofs; ++q)
    if (idofs[q] < 0 || idofs[q] >= n) { set_error("setup: interface dof out of range"); return -1; }
  // a dof listed twice would seed the same patch twice (applied twice per sweep): keep the first occurrence
  std::vector<int> seeds;
  {
    std::vector<char> seen(n, 0);
    seeds.reserve(n_idofs);
    for (int q = 0; q < n_idofs; ++q)
      if (!seen[idofs[q]]) { seen[idofs[q]] = 1; seeds.push_back(idofs[q]); }
  }
  Csr A;
  A.n = A.m = n;
  A.ia.assign(indptr, indptr + n + 1);
  A.ja.resize(nnz);
  A.a.resize(nnz);
  // the copies and checks below walk 10^9 entries at the BASELINE sizes: all of them run on every core
  long long bad_col = 0, unsorted = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad_col, unsorted)
  for (int i = 0; i < n; ++i)
    for (int q = indptr[i]; q < indptr[i + 1]; ++q) {
      const int j = indices[q];
      A.ja[q] = j;
      A.a[q] = data[q];
      if (j < 0 || j >= n) ++bad_col;
      if (q > indptr[i] && j <= indices[q - 1]) ++unsorted;
    }
  if (bad_col) { set_error("setup: column index out of range"); return -1; }
  const bool canonical = unsorted == 0;   // columns strictly ascending inside every row (what PETSc / scipy hand over)
  if (!canonical) {
    // the aggregation and colouring break ties by position inside the row: sort the rows and add up
    // duplicate entries so that the hierarchy does not depend on the caller's storage order
    Csr B;
    B.n = B.m = n;
    B.ia.assign(n + 1, 0);
    B.ja.reserve(nnz);
    B.a.reserve(nnz);
    std::vector<std::pair<int, double>> row;
    for (int i = 0; i < n; ++i) {
      row.clear();
      for (int q = A.ia[i]; q < A.ia[i + 1]; ++q) row.emplace_back(A.ja[q], A.a[q]);
      std::stable_sort(row.begin(), row.end(),
                       [](const std::pair<int, double>& u, const std::pair<int, double>& v) { return u.first < v.first; });
      for (size_t k = 0; k < row.size(); ++k) {
        if (k > 0 && row[k].first == row[k - 1].first) B.a.back() += row[k].second;
        else { B.ja.push_back(row[k].first); B.a.push_back(row[k].second); }
      }
      B.ia[i + 1] = (int)B.ja.size();
    }
    A = std::move(B);
  }
  {
    int first_bad = n;
#pragma omp parallel for schedule(static) reduction(min : first_bad)
    for (int i = 0; i < n; ++i) {
      bool has_diag = false;
      for (int q = A.ia[i]; q < A.ia[i + 1]; ++q)
        if (A.ja[q] == i && A.a[q] != 0.0) has_diag = true;
      if (!has_diag) first_bad = std::min(first_bad, i);
    }
    if (first_bad < n) { set_error("setup: row " + std::to_string(first_bad) + " has no nonzero diagonal"); return -1; }
  }
  {
    // the Gauss-Seidel and patch colourings treat the nonzero pattern as an undirected graph: with a
    // nonsymmetric pattern two coupled rows could share a colour (a silent data race), so refuse it
    long long bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
    for (int i = 0; i < n; ++i)
      for (int q = A.ia[i]; q < A.ia[i + 1]; ++q) {
        const int j = A.ja[q];
        if (j == i || A.a[q] == 0.0) continue;
        const int* lo = A.ja.data() + A.ia[j];
        const int* hi = A.ja.data() + A.ia[j + 1];
        const int* it = std::lower_bound(lo, hi, i);
        if (it == hi || *it != i || A.a[it - A.ja.data()] == 0.0) ++bad;
      }
    if (bad) {
      set_error("setup: the nonzero pattern is not symmetric (" + std::to_string(bad) + " entries a_ij != 0 with a_ji == 0); "
                "the multicolour smoothers need a structurally symmetric matrix");
      return -1;
    }
  }
  if (part) {
    if (nparts < 1) { set_error("setup: nparts < 1"); return -1; }
    for (int i = 0; i < n; ++i)
      if (part[i] < 0 || part[i] >= nparts) { set_error("setup: part id out of range"); return -1; }
  }
  mamg_handle h = new mamg_handle_s();
  std::string err;
  if (!build_hierarchy(*p, std::move(A), seeds.data(), (int)seeds.size(), part, nparts, h->H, err)) {
    delete h;
    set_error("AMG levels failed to set up: " + err);
    return -3;
  }
  *out = h;
  return 0;
  MAMG_CATCH
}

int mamg_destroy(mamg_handle h) {
  if (!h) return 0;
  if (h->dev) device_state_free(h->dev);
  delete h;
  return 0;
}

int mamg_release_host(mamg_handle h) {
  if (!h) { set_error("NULL handle"); return -1; }
  if (!h->dev) { set_error("release_host: hierarchy is not on a device yet"); return -1; }
  for (Level& L : h->H.lv) {
    bigvec<int>().swap(L.A.ja);
    bigvec<double>().swap(L.A.a);
    std::vector<int>().swap(L.sw.dofs);
    Csr().ia.swap(L.P.ia); bigvec<int>().swap(L.P.ja); bigvec<double>().swap(L.P.a);
    bigvec<int>().swap(L.R.ja); bigvec<double>().swap(L.R.a);
  }
  h->H.released = true;
  return 0;
}

int mamg_num_levels(mamg_handle h, int32_t* nlevels) {
  if (!h || !nlevels) { set_error("num_levels: NULL"); return -1; }
  *nlevels = (int32_t)h->H.lv.size();
  return 0;
}

static const Level* get_level(mamg_handle h, int level) {
  if (!h) { set_error("NULL handle"); return nullptr; }
  if (h->H.released) { set_error("host hierarchy was released (mamg_release_host)"); return nullptr; }
  if (level < 0 || level >= (int)h->H.lv.size()) { set_error("level out of range"); return nullptr; }
  return &h->H.lv[level];
}

int mamg_level_info(mamg_handle h, int32_t level, int64_t info[12]) {
  const Level* L = get_level(h, level);
  if (!L) return -1;
  int64_t row_entries = 0, inv_entries = 0;
  for (int p = 0; p < L->sw.npatch(); ++p) {
    const int64_t s = L->sw.ptr[p + 1] - L->sw.ptr[p];
    inv_entries += s * (s + 1) / 2;
    for (int q = L->sw.ptr[p]; q < L->sw.ptr[p + 1]; ++q)
      row_entries += L->A.ia[L->sw.dofs[q] + 1] - L->A.ia[L->sw.dofs[q]];
  }
  info[8] = row_entries;
  info[9] = inv_entries;
  info[10] = L->P.nnz();
  info[11] = L->nnz_structural;
  info[0] = L->A.n;
  info[1] = L->A.nnz();
  info[2] = L->nc;
  info[3] = L->ncolors;
  info[4] = L->sw.npatch();
  info[5] = (int64_t)L->sw.dofs.size();
  info[6] = L->sw.ncolors;
  info[7] = L->sw.max_size;
  return 0;
}

int mamg_level_export(mamg_handle h, int32_t level, int32_t* indptr, int32_t* indices, double* data,
                      int32_t* agg, int32_t* color, uint8_t* gs_skip) {
  const Level* L = get_level(h, level);
  if (!L) return -1;
  const int n = L->A.n;
  if (indptr) std::memcpy(indptr, L->A.ia.data(), sizeof(int) * (n + 1));
  if (indices) std::memcpy(indices, L->A.ja.data(), sizeof(int) * L->A.ja.size());
  if (data) std::memcpy(data, L->A.a.data(), sizeof(double) * L->A.a.size());
  if (agg) {
    if (L->agg.empty()) for (int i = 0; i < n; ++i) agg[i] = -1;
    else std::memcpy(agg, L->agg.data(), sizeof(int) * n);
  }
  if (color) {
    if (L->color.empty()) for (int i = 0; i < n; ++i) color[i] = 0;
    else std::memcpy(color, L->color.data(), sizeof(int) * n);
  }
  if (gs_skip) {
    if (L->gs_skip.empty()) std::memset(gs_skip, 0, n);
    else std::memcpy(gs_skip, L->gs_skip.data(), n);
  }
  return 0;
}

int mamg_schwarz_export(mamg_handle h, int32_t level, int32_t* patch_ptr, int32_t* patch_dofs,
                        int32_t* patch_seed, int32_t* patch_color) {
  const Level* L = get_level(h, level);
  if (!L) return -1;
  const SchwarzPatches& s = L->sw;
  if (patch_ptr && !s.ptr.empty()) std::memcpy(patch_ptr, s.ptr.data(), sizeof(int) * s.ptr.size());
  if (patch_dofs && !s.dofs.empty()) std::memcpy(patch_dofs, s.dofs.data(), sizeof(int) * s.dofs.size());
  if (patch_seed && !s.seed.empty()) std::memcpy(patch_seed, s.seed.data(), sizeof(int) * s.seed.size());
  if (patch_color && !s.color.empty()) std::memcpy(patch_color, s.color.data(), sizeof(int) * s.color.size());
  return 0;
}

int mamg_part_export(mamg_handle h, int32_t level, int32_t* part) {
  const Level* L = get_level(h, level);
  if (!L || !part) return -1;
  for (int i = 0; i < L->A.n; ++i) part[i] = L->part.empty() ? 0 : L->part[i];
  return 0;
}

int mamg_prolongator_export(mamg_handle h, int32_t level, int32_t* indptr, int32_t* indices, double* data) {
  const Level* L = get_level(h, level);
  if (!L) return -1;
  if (L->P.n == 0) { set_error("level has no stored prolongator (UA_AMG uses the aggregate map)"); return -1; }
  if (indptr) std::memcpy(indptr, L->P.ia.data(), sizeof(int) * (L->P.n + 1));
  if (indices) std::memcpy(indices, L->P.ja.data(), sizeof(int) * L->P.ja.size());
  if (data) std::memcpy(data, L->P.a.data(), sizeof(double) * L->P.a.size());
  return 0;
}

int mamg_coarse_export(mamg_handle h, double* inv) {
  if (!h || !inv) { set_error("coarse_export: NULL"); return -1; }
  std::memcpy(inv, h->H.coarse_inv.data(), sizeof(double) * h->H.coarse_inv.size());
  return 0;
}

int mamg_setup_seconds(mamg_handle h, double* seconds) {
  if (!h || !seconds) { set_error("setup_seconds: NULL"); return -1; }
  *seconds = h->H.setup_seconds;
  return 0;
}

static int export_csr(const Csr& A, int64_t* nrows, int64_t* nnz, int32_t* indptr, int32_t* indices,
                      double* data) {
  const int64_t cap = (indptr && nnz) ? *nnz : 0;
  if (nrows) *nrows = A.n;
  if (nnz) *nnz = A.nnz();
  if (indptr) {
    if (cap < A.nnz()) { set_error("assemble: indices/data capacity (*nnz on entry) too small"); return -1; }
    std::memcpy(indptr, A.ia.data(), sizeof(int) * (A.n + 1));
    if (indices) std::memcpy(indices, A.ja.data(), sizeof(int) * A.ja.size());
    if (data) std::memcpy(data, A.a.data(), sizeof(double) * A.a.size());
  }
  return 0;
}

// indptr == NULL: sizes only.  Otherwise *nnz holds the capacity of indices/data on entry (an
// upper bound such as rows * 2 * 3^dim is enough) and the actual nnz on return.
static int check_problem(int dim, int n, bool emi) {
  if (dim != 2 && dim != 3) { set_error("assemble: dim must be 2 or 3"); return -1; }
  if (n < 2) { set_error("assemble: need at least 2 cells per direction"); return -1; }
  if (emi && (n < 4 || n % 2)) { set_error("assemble_emi: ncell must be even and >= 4 (src/utils.py:192)"); return -1; }
  double nv = 1;
  for (int a = 0; a < dim; ++a) nv *= (a == dim - 1 && emi) ? n / 2 + 1 : n + 1;
  double nnz_est = 2 * nv * (dim == 2 ? 7 : 15) * (emi ? 1.05 : 2);
  if (2 * nv > 2.0e9 || nnz_est > 2.1e9) { set_error("assemble: system exceeds int32 indexing"); return -1; }
  return 0;
}

int mamg_assemble_scalar(int32_t dim, const int32_t* ncell, const double* hh, double cK, double cM,
                         int64_t* nrows, int64_t* nnz, int32_t* indptr, int32_t* indices,
                         double* data) {
  MAMG_TRY
  if (dim < 1 || dim > 3 || !ncell || !hh) { set_error("assemble_scalar: bad arguments"); return -1; }
  Csr A;
  p1_scalar(dim, ncell, hh, cK, cM, A);
  return export_csr(A, nrows, nnz, indptr, indices, data);
  MAMG_CATCH
}

int mamg_assemble_bidomain(int32_t dim, int32_t ncell, double kappa1, double kappa2, double gamma,
                           int64_t* nrows, int64_t* nnz, int32_t* indptr, int32_t* indices,
                           double* data) {
  MAMG_TRY
  if (check_problem(dim, ncell, false)) return -1;
  Csr A;
  assemble_bidomain(dim, ncell, kappa1, kappa2, gamma, A);
  return export_csr(A, nrows, nnz, indptr, indices, data);
  MAMG_CATCH
}

int mamg_assemble_emi(int32_t dim, int32_t ncell, double kappa1, double kappa2, double gamma,
                      int64_t* nrows, int64_t* nnz, int32_t* indptr, int32_t* indices,
                      double* data) {
  MAMG_TRY
  if (check_problem(dim, ncell, true)) return -1;
  Csr A;
  assemble_emi(dim, ncell, kappa1, kappa2, gamma, A);
  return export_csr(A, nrows, nnz, indptr, indices, data);
  MAMG_CATCH
}

}  // extern "C"

// ---- import of an externally produced hierarchy -----------------------------------------------
namespace {
bool fail(const std::string& msg) { set_error("import_hierarchy: " + msg); return false; }

bool import_level(const mamg_level_arrays& in, int l, bool last, Level& L) {
  const std::string at = " (level " + std::to_string(l) + ")";
  // MAMG_IMPORT_NOCHECK=1 skips the colouring validation (tests of mamg_race_check need a racy layout)
  const bool check_colors = !(getenv("MAMG_IMPORT_NOCHECK") && atoi(getenv("MAMG_IMPORT_NOCHECK")) != 0);
  if (in.n <= 0 || !in.indptr || !in.indices || !in.data) return fail("missing matrix" + at);
  const int n = in.n;
  if (in.indptr[0] != 0) return fail("indptr[0] != 0" + at);
  for (int i = 0; i < n; ++i)
    if (in.indptr[i + 1] < in.indptr[i]) return fail("indptr not monotone" + at);
  const int nnz = in.indptr[n];
  L.A.n = L.A.m = n;
  L.A.ia.assign(in.indptr, in.indptr + n + 1);
  L.A.ja.assign(in.indices, in.indices + nnz);
  L.A.a.assign(in.data, in.data + nnz);
  L.nnz_structural = nnz;
  for (int i = 0; i < n; ++i) {
    bool diag = false;
    for (int q = L.A.ia[i]; q < L.A.ia[i + 1]; ++q) {
      const int j = L.A.ja[q];
      if (j < 0 || j >= n) return fail("column index out of range" + at);
      if (q > L.A.ia[i] && j <= L.A.ja[q - 1]) return fail("columns must be strictly ascending inside a row" + at);
      diag |= j == i && L.A.a[q] != 0.0;
    }
    if (!diag) return fail("row " + std::to_string(i) + " has no nonzero diagonal" + at);
  }
  if (!last) {
    if (!in.agg || in.n_aggregates <= 0) return fail("aggregates missing on a non-coarsest level" + at);
    L.agg.assign(in.agg, in.agg + n);
    L.nc = in.n_aggregates;
    std::vector<char> hit(L.nc, 0);
    for (int i = 0; i < n; ++i) {
      if (L.agg[i] < -1 || L.agg[i] >= L.nc) return fail("aggregate id out of range" + at);
      if (L.agg[i] >= 0) hit[L.agg[i]] = 1;
    }
    for (int I = 0; I < L.nc; ++I)
      if (!hit[I]) return fail("aggregate " + std::to_string(I) + " is empty" + at);
  }
  if (in.n_patches > 0) {
    if (!in.patch_ptr || !in.patch_dofs || !in.patch_color) return fail("patch arrays missing" + at);
    SchwarzPatches& s = L.sw;
    s.ptr.assign(in.patch_ptr, in.patch_ptr + in.n_patches + 1);
    if (s.ptr[0] != 0) return fail("patch_ptr[0] != 0" + at);
    for (int p = 0; p < in.n_patches; ++p)
      if (s.ptr[p + 1] <= s.ptr[p]) return fail("empty patch" + at);
    s.dofs.assign(in.patch_dofs, in.patch_dofs + s.ptr[in.n_patches]);
    s.color.assign(in.patch_color, in.patch_color + in.n_patches);
    s.seed.resize(in.n_patches);
    s.ncolors = in.n_patch_colors;
    s.max_size = 0;
    for (int p = 0; p < in.n_patches; ++p) {
      s.max_size = std::max(s.max_size, s.ptr[p + 1] - s.ptr[p]);
      if (s.color[p] < 0 || s.color[p] >= s.ncolors) return fail("patch colour out of range" + at);
      for (int q = s.ptr[p]; q < s.ptr[p + 1]; ++q) {
        if (s.dofs[q] < 0 || s.dofs[q] >= n) return fail("patch dof out of range" + at);
        if (q > s.ptr[p] && s.dofs[q] <= s.dofs[q - 1]) return fail("patch dofs must be strictly ascending" + at);
      }
      s.seed[p] = in.patch_seed ? in.patch_seed[p] : s.dofs[s.ptr[p]];
      if (s.seed[p] < 0 || s.seed[p] >= n) return fail("patch seed out of range" + at);
    }
    // patches of one colour are solved concurrently: none may share a dof with, or be coupled to, another
    // patch of its colour (owner[j] = the patch of the colour under test that has j among its dofs)
    std::vector<std::vector<int>> by_color(s.ncolors);
    for (int p = 0; p < in.n_patches; ++p) by_color[s.color[p]].push_back(p);
    std::vector<int> owner(n, -1);
    for (int c = 0; c < s.ncolors && check_colors; ++c) {
      for (int p : by_color[c])
        for (int q = s.ptr[p]; q < s.ptr[p + 1]; ++q) {
          if (owner[s.dofs[q]] >= 0) return fail("patches " + std::to_string(owner[s.dofs[q]]) + " and " + std::to_string(p) + " of one colour share a dof" + at);
          owner[s.dofs[q]] = p;
        }
      for (int p : by_color[c])
        for (int q = s.ptr[p]; q < s.ptr[p + 1]; ++q) {
          const int i = s.dofs[q];
          for (int e = L.A.ia[i]; e < L.A.ia[i + 1]; ++e) {
            if (L.A.a[e] == 0.0) continue;
            const int o = owner[L.A.ja[e]];
            if (o >= 0 && o != p) return fail("patches " + std::to_string(o) + " and " + std::to_string(p) + " of one colour are coupled" + at);
          }
        }
      for (int p : by_color[c])
        for (int q = s.ptr[p]; q < s.ptr[p + 1]; ++q) owner[s.dofs[q]] = -1;
    }
    if (!in.gs_skip) return fail("gs_skip missing on a Schwarz level" + at);
  }
  if (in.gs_skip && (in.n_patches > 0)) L.gs_skip.assign(in.gs_skip, in.gs_skip + n);
  if (!last) {
    if (in.color) {
      L.color.assign(in.color, in.color + n);
      L.ncolors = in.n_colors;
      for (int i = 0; i < n; ++i) {
        if (L.color[i] < 0 || L.color[i] >= L.ncolors) return fail("row colour out of range" + at);
        if (!check_colors || (!L.gs_skip.empty() && L.gs_skip[i])) continue;
        for (int q = L.A.ia[i]; q < L.A.ia[i + 1]; ++q) {
          const int j = L.A.ja[q];
          if (j == i || L.A.a[q] == 0.0 || (!L.gs_skip.empty() && L.gs_skip[j])) continue;
          if (L.color[j] == L.color[i]) return fail("rows " + std::to_string(i) + " and " + std::to_string(j) + " are coupled and share a colour" + at);
        }
      }
    } else {
      multicolor_greedy(L.A, L.gs_skip, L.color, L.ncolors);
    }
  }
  if (in.P_indptr && !last) {
    if (!in.P_indices || !in.P_data) return fail("prolongator arrays missing" + at);
    L.P.n = n;
    L.P.m = L.nc;
    L.P.ia.assign(in.P_indptr, in.P_indptr + n + 1);
    L.P.ja.assign(in.P_indices, in.P_indices + L.P.ia[n]);
    L.P.a.assign(in.P_data, in.P_data + L.P.ia[n]);
    for (int j : L.P.ja)
      if (j < 0 || j >= L.nc) return fail("prolongator column out of range" + at);
    csr_transpose(L.P, L.R);
  }
  if (in.part) L.part.assign(in.part, in.part + n);
  return true;
}
}  // namespace

extern "C" int mamg_import_hierarchy(const mamg_params* p, int32_t nlevels, const mamg_level_arrays* levels,
                                     const double* coarse_inv, int32_t nparts, mamg_handle* out) {
  MAMG_TRY
  if (!p || !levels || !out || nlevels < 1) { set_error("import_hierarchy: NULL argument or no levels"); return -1; }
  {
    std::string perr;
    if (!validate_params(*p, perr)) { set_error("import_hierarchy: " + perr); return -1; }
  }
  mamg_handle h = new mamg_handle_s();
  Hierarchy& H = h->H;
  H.prm = *p;
  H.nparts = std::max(1, nparts);
  H.lv.resize(nlevels);
  for (int l = 0; l < nlevels; ++l) {
    if (!import_level(levels[l], l, l == nlevels - 1, H.lv[l])) { delete h; return -1; }
    if (l > 0 && H.lv[l].A.n != H.lv[l - 1].nc) {
      delete h;
      set_error("import_hierarchy: level " + std::to_string(l) + " has " + std::to_string(levels[l].n) + " rows but level " +
                std::to_string(l - 1) + " has " + std::to_string(levels[l - 1].n_aggregates) + " aggregates");
      return -1;
    }
    if (H.nparts > 1 && H.lv[l].part.empty()) { delete h; set_error("import_hierarchy: nparts > 1 needs part[] on every level"); return -1; }
    for (int v : H.lv[l].part)
      if (v < 0 || v >= H.nparts) { delete h; set_error("import_hierarchy: part id out of range"); return -1; }
  }
  const int nc = H.lv.back().A.n;
  if (nc > 8192) { delete h; set_error("import_hierarchy: coarsest level has more than 8192 rows"); return -1; }
  if (coarse_inv) H.coarse_inv.assign(coarse_inv, coarse_inv + (size_t)nc * nc);
  else if (!dense_inverse(H.lv.back().A, H.coarse_inv)) { delete h; set_error("import_hierarchy: coarsest operator is singular"); return -1; }
  *out = h;
  return 0;
  MAMG_CATCH
}

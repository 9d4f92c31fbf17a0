/* mamg_oracle.c -- CPU restatement of the metric-AMG apply path.  TEST INFRASTRUCTURE ONLY:
 * it may be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs, never by the product package.
 *
 * PARITY UNPINNED.  The reference repository (anabudisa/metric-amg-examples) contains no
 * arithmetic of this path, no tests and no golden vectors; the algorithm lives in three
 * un-vendored, un-pinned dependencies that are absent from this machine:
 *   - HAZmath / haznics (github.com/HAZmathTeam/hazmath, branch unspecified, README.md:19-20):
 *     precond_amg -> mgcycle, smoother_dcsr_{gs,sgs,sor,jacobi}, smoother_dcsr_Schwarz_*,
 *     aggregate-specialised restriction/prolongation, UMFPACK coarse solve;
 *   - cbc.block (bitbucket.org/fenics-apps/cbc.block master, README.md:19):
 *     block.iterative.ConjGrad (precondconjgrad) and the hazmath Precond.matvec wrapper;
 *   - FEniCS_ii (github.com/MiroK/fenics_ii): ii_convert / ReductionOperator (block flattening).
 * What follows restates their published algorithms (FASP-lineage multigrid cycle, standard
 * preconditioned CG) and is anchored on the reference's own call sites:
 *   parameters      src/amg_parameters.py:47-89, src/utils.py:60-82
 *   operator        src/utils.py:45-90 (metricAMG(A, W, idofs=, parameters=))
 *   Krylov call     src/bidomain_2d.py:205-216, src/emi_2d.py:211, src/emi_3d.py:143
 * Choices HAZmath's source would pin are frozen here and in DESIGN.md ("Frozen choices").
 *
 * Two smoother orderings are implemented:
 *   ordering 0  natural row / patch order (the sequential order HAZmath uses)
 *   ordering 1  the multicolour order the device uses ("a fixed multicolour ordering that is
 *               applied identically in the reference comparison", BASELINE.json north_star)
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#else
static int omp_get_thread_num(void) { return 0; }
static int omp_get_max_threads(void) { return 1; }
#endif

enum { UA_AMG = 1, V_CYCLE = 1, W_CYCLE = 2, AMLI_CYCLE = 3, NL_AMLI_CYCLE = 4, ADD_CYCLE = 5, SOLVER_GCG = 5, SM_JACOBI = 1, SM_GS = 2, SM_SGS = 3, SM_SOR = 5, SM_SSOR = 6, SM_L1DIAG = 10,
       SW_FORWARD = 1, SW_BACKWARD = 2, SW_SYMMETRIC = 3 };

typedef struct {
  int n, nnz, nc, ncolors, npatch, npcolors, maxpatch;
  int *ia, *ja, *agg, *color, *pptr, *pdofs, *pcolor;
  /* rows / patches grouped by colour (built once) */
  int *crow_ptr, *crow, *cpat_ptr, *cpat;
  unsigned char* skip;
  double *a, *invd, *x, *b, *w;
  double* kw;       /* 4 n work vectors of the AMLI / K-cycle coarse corrections (allocated on first use) */
  /* Schwarz blocks are factorised once at setup, as HAZmath does with UMFPACK
   * (Schwarz_blksolver 32): packed lower Cholesky factors, pfoff[p] = start of patch p */
  double* pfac;
  long* pfoff;
  /* SA_AMG: smoothed prolongator P (n x nc CSR); NULL for unsmoothed aggregation */
  int *pia, *pja;
  double* pa;
} orc_level;

typedef struct {
  int cycle_type, maxit, smoother, presmooth, postsmooth, coarse_scaling, schwarz_type;
  int amli_degree, nl_amli_krylov_type;
  double amli_coef[16];
  double relaxation;
  int nlevels, cap;
  orc_level* lv;
  double* coarse_inv;
  double* rhs;      /* patch scratch, one slice per thread */
  int maxpatch;
  int threads;      /* > 1 only for the multicolour ordering (bench reference arm) */
  int ordering;
  long visits;      /* level visits of the last apply (for reporting) */
  double kmargin_all; /* the same since the hierarchy was created */
  double kmargin;   /* K-cycle: smallest |nr2 / (tol2 nb2) - 1| of the last apply (how close a stopping decision was) */
} orc_hier;

static double vdot(int n, const double* u, const double* v);
static void* xmalloc(size_t n) { void* p = malloc(n ? n : 1); if (!p) abort(); return p; }
static void* xdup(const void* src, size_t n) { void* p = xmalloc(n); if (n) memcpy(p, src, n); return p; }

#define TRI(i, j) ((size_t)(i) * ((i) + 1) / 2 + (j))

/* dense Cholesky A_BB = L L' of one patch (dofs in ascending order), packed lower storage */
static void patch_factor(orc_level* L, int p) {
  const int q0 = L->pptr[p], s = L->pptr[p + 1] - q0;
  const int* idx = L->pdofs + q0;
  double* F = L->pfac + L->pfoff[p];
  for (long k = 0; k < (long)s * (s + 1) / 2; ++k) F[k] = 0.0;
  for (int k = 0; k < s; ++k) {
    int i = idx[k];
    for (int e = L->ia[i]; e < L->ia[i + 1]; ++e) {
      int j = L->ja[e], lo = 0, hi = k;  /* idx is sorted ascending: binary search in [0, k] */
      while (lo <= hi) {
        int mid = (lo + hi) / 2;
        if (idx[mid] == j) { F[TRI(k, mid)] = L->a[e]; break; }
        if (idx[mid] < j) lo = mid + 1; else hi = mid - 1;
      }
    }
  }
  for (int j = 0; j < s; ++j) {
    double d = sqrt(F[TRI(j, j)]);
    F[TRI(j, j)] = d;
    for (int i = j + 1; i < s; ++i) F[TRI(i, j)] /= d;
    for (int k = j + 1; k < s; ++k) {
      double lkj = F[TRI(k, j)];
      for (int i = k; i < s; ++i) F[TRI(i, k)] -= F[TRI(i, j)] * lkj;
    }
  }
}

orc_hier* orc_create(int cycle_type, int maxit, int smoother, double relaxation, int presmooth,
                     int postsmooth, int coarse_scaling, int schwarz_type) {
  orc_hier* h = (orc_hier*)calloc(1, sizeof(orc_hier));
  h->cycle_type = cycle_type; h->maxit = maxit; h->smoother = smoother; h->relaxation = relaxation;
  h->presmooth = presmooth; h->postsmooth = postsmooth; h->coarse_scaling = coarse_scaling;
  h->schwarz_type = schwarz_type; h->ordering = 1; h->threads = 1;
  h->amli_degree = 3; h->nl_amli_krylov_type = 4; h->kmargin = h->kmargin_all = 1e300;
  return h;
}

int orc_add_level(orc_hier* h, int n, const int* ia, const int* ja, const double* a, const int* agg,
                  int nc, const int* color, int ncolors, const unsigned char* skip, int npatch,
                  const int* pptr, const int* pdofs, const int* pcolor, int npcolors) {
  if (h->nlevels == h->cap) {
    h->cap = h->cap ? 2 * h->cap : 8;
    h->lv = (orc_level*)realloc(h->lv, sizeof(orc_level) * h->cap);
  }
  orc_level* L = &h->lv[h->nlevels++];
  memset(L, 0, sizeof(*L));
  L->n = n; L->nnz = ia[n]; L->nc = nc; L->ncolors = ncolors > 0 ? ncolors : 1;
  L->ia = (int*)xdup(ia, sizeof(int) * (n + 1));
  L->ja = (int*)xdup(ja, sizeof(int) * L->nnz);
  L->a = (double*)xdup(a, sizeof(double) * L->nnz);
  L->agg = (int*)xdup(agg, sizeof(int) * n);
  L->color = (int*)xdup(color, sizeof(int) * n);
  L->skip = (unsigned char*)xdup(skip, n);
  L->invd = (double*)xmalloc(sizeof(double) * n);
  for (int i = 0; i < n; ++i) {
    L->invd[i] = 1.0;
    for (int p = ia[i]; p < ia[i + 1]; ++p) if (ja[p] == i) L->invd[i] = 1.0 / a[p];
  }
  L->x = (double*)calloc(n ? n : 1, sizeof(double));
  L->b = (double*)calloc(n ? n : 1, sizeof(double));
  L->w = (double*)calloc(n ? n : 1, sizeof(double));
  /* rows by colour, natural order inside a colour */
  L->crow_ptr = (int*)calloc(L->ncolors + 1, sizeof(int));
  L->crow = (int*)xmalloc(sizeof(int) * n);
  for (int i = 0; i < n; ++i) L->crow_ptr[L->color[i] + 1]++;
  for (int c = 0; c < L->ncolors; ++c) L->crow_ptr[c + 1] += L->crow_ptr[c];
  {
    int* fill = (int*)xdup(L->crow_ptr, sizeof(int) * (L->ncolors + 1));
    for (int i = 0; i < n; ++i) L->crow[fill[L->color[i]]++] = i;
    free(fill);
  }
  L->npatch = npatch; L->npcolors = npcolors;
  if (npatch > 0) {
    L->pptr = (int*)xdup(pptr, sizeof(int) * (npatch + 1));
    L->pdofs = (int*)xdup(pdofs, sizeof(int) * pptr[npatch]);
    L->pcolor = (int*)xdup(pcolor, sizeof(int) * npatch);
    L->cpat_ptr = (int*)calloc(npcolors + 1, sizeof(int));
    L->cpat = (int*)xmalloc(sizeof(int) * npatch);
    for (int p = 0; p < npatch; ++p) {
      L->cpat_ptr[pcolor[p] + 1]++;
      int s = pptr[p + 1] - pptr[p];
      if (s > L->maxpatch) L->maxpatch = s;
    }
    for (int c = 0; c < npcolors; ++c) L->cpat_ptr[c + 1] += L->cpat_ptr[c];
    int* fill = (int*)xdup(L->cpat_ptr, sizeof(int) * (npcolors + 1));
    for (int p = 0; p < npatch; ++p) L->cpat[fill[pcolor[p]]++] = p;
    free(fill);
    L->pfoff = (long*)xmalloc(sizeof(long) * (npatch + 1));
    L->pfoff[0] = 0;
    for (int p = 0; p < npatch; ++p) {
      long s = pptr[p + 1] - pptr[p];
      L->pfoff[p + 1] = L->pfoff[p] + s * (s + 1) / 2;
    }
    L->pfac = (double*)xmalloc(sizeof(double) * L->pfoff[npatch]);
#pragma omp parallel for schedule(dynamic, 64)
    for (int p = 0; p < npatch; ++p) patch_factor(L, p);
    if (L->maxpatch > h->maxpatch) h->maxpatch = L->maxpatch;
    size_t m = (size_t)h->maxpatch, nt = (size_t)omp_get_max_threads();
    h->rhs = (double*)realloc(h->rhs, sizeof(double) * m * nt);
  }
  return h->nlevels - 1;
}

/* smoothed prolongator of level lev (SA_AMG, src/input_metric.dat:68); restriction is its transpose */
void orc_set_prolongator(orc_hier* h, int lev, const int* pia, const int* pja, const double* pa) {
  orc_level* L = &h->lv[lev];
  L->pia = (int*)xdup(pia, sizeof(int) * (L->n + 1));
  L->pja = (int*)xdup(pja, sizeof(int) * pia[L->n]);
  L->pa = (double*)xdup(pa, sizeof(double) * pia[L->n]);
}

void orc_set_coarse(orc_hier* h, const double* inv) {
  int n = h->lv[h->nlevels - 1].n;
  h->coarse_inv = (double*)xdup(inv, sizeof(double) * (size_t)n * n);
}
void orc_set_ordering(orc_hier* h, int ordering) { h->ordering = ordering; if (ordering == 0) h->threads = 1; }
/* threads > 1 (multicolour ordering only): colour classes and vector loops run on several cores */
int orc_set_threads(orc_hier* h, int threads) {
  int mx = omp_get_max_threads();
  if (threads < 1 || threads > mx) threads = mx;
  h->threads = h->ordering == 1 ? threads : 1;
  return h->threads;
}
void orc_set_cycle(orc_hier* h, int cycle_type) { h->cycle_type = cycle_type; }
/* amli_degree (src/amg_parameters.py:62,82) and HAZmath's nl_amli_krylov_type (SOLVER_GCG = 5 -> GCG, else GCR) */
void orc_set_amli(orc_hier* h, int degree, int krylov_type) {
  h->amli_degree = degree < 0 ? 0 : (degree > 15 ? 15 : degree);
  h->nl_amli_krylov_type = krylov_type;
}
long orc_visits(orc_hier* h) { return h->visits; }
double orc_kcycle_margin(orc_hier* h, int all) { return all ? h->kmargin_all : h->kmargin; }

void orc_destroy(orc_hier* h) {
  if (!h) return;
  for (int l = 0; l < h->nlevels; ++l) {
    orc_level* L = &h->lv[l];
    free(L->ia); free(L->ja); free(L->a); free(L->agg); free(L->color); free(L->skip); free(L->invd);
    free(L->x); free(L->b); free(L->w); free(L->crow_ptr); free(L->crow);
    free(L->pptr); free(L->pdofs); free(L->pcolor); free(L->cpat_ptr); free(L->cpat);
    free(L->pfac); free(L->pfoff); free(L->pia); free(L->pja); free(L->pa); free(L->kw);
  }
  free(h->lv); free(h->coarse_inv); free(h->rhs); free(h);
}

/* ---- kernels ------------------------------------------------------------------------------ */
static double row_dot(const orc_level* L, int i, const double* x) {
  double s = 0.0;
  for (int p = L->ia[i]; p < L->ia[i + 1]; ++p) s += L->a[p] * x[L->ja[p]];
  return s;
}

void orc_spmv_level(const orc_level* L, const double* x, double* y) {
  for (int i = 0; i < L->n; ++i) y[i] = row_dot(L, i, x);
}

/* x_i <- x_i + w (b_i - a_i . x) / a_ii  for one row */
static void gs_row(const orc_level* L, int i, const double* b, double* x, double w) {
  if (L->skip[i]) return;
  x[i] += w * (b[i] - row_dot(L, i, x)) * L->invd[i];
}

/* one directional sweep. ordering 0: rows 0..n-1 (or reversed); ordering 1: colours ascending
 * (descending), natural order inside a colour; skip_first_color drops the colour the previous
 * sweep just finished (exact no-op for w = 1, mirrored from the device path). */
static void gs_sweep(const orc_hier* h, const orc_level* L, const double* b, double* x, double w,
                     int backward, int skip_first_color) {
  if (h->ordering == 0) {
    if (!backward) for (int i = 0; i < L->n; ++i) gs_row(L, i, b, x, w);
    else for (int i = L->n - 1; i >= 0; --i) gs_row(L, i, b, x, w);
    return;
  }
  for (int cc = skip_first_color; cc < L->ncolors; ++cc) {
    int c = backward ? L->ncolors - 1 - cc : cc;
    /* rows of one colour do not couple: the threaded loop gives the same numbers as the serial one */
#pragma omp parallel for schedule(static) num_threads(h->threads) if (h->threads > 1)
    for (int q = L->crow_ptr[c]; q < L->crow_ptr[c + 1]; ++q) gs_row(L, L->crow[q], b, x, w);
  }
}

static void jacobi(const orc_level* L, const double* b, double* x, double w) {
  for (int i = 0; i < L->n; ++i)
    L->w[i] = L->skip[i] ? x[i] : x[i] + w * (b[i] - row_dot(L, i, x)) * L->invd[i];
  memcpy(x, L->w, sizeof(double) * L->n);
}

/* l1-Jacobi (haznics SMOOTHER_L1DIAG): x += D_l1^{-1} (b - A x) with (D_l1)_ii = sum_j |a_ij|, all rows at once */
static void jacobi_l1(const orc_level* L, const double* b, double* x) {
  for (int i = 0; i < L->n; ++i) {
    double l1 = 0.0;
    for (int p = L->ia[i]; p < L->ia[i + 1]; ++p) l1 += fabs(L->a[p]);
    L->w[i] = L->skip[i] ? x[i] : x[i] + (b[i] - row_dot(L, i, x)) * (1.0 / l1);
  }
  memcpy(x, L->w, sizeof(double) * L->n);
}

/* exact solve on one patch: x_B += A_BB^{-1} (b - A x)_B with the stored Cholesky factor */
static void patch_solve(orc_hier* h, const orc_level* L, int p, const double* b, double* x) {
  const int q0 = L->pptr[p], s = L->pptr[p + 1] - q0;
  const int* idx = L->pdofs + q0;
  const double* F = L->pfac + L->pfoff[p];
  double* r = h->rhs + (size_t)omp_get_thread_num() * (size_t)h->maxpatch;
  for (int k = 0; k < s; ++k) r[k] = b[idx[k]] - row_dot(L, idx[k], x);
  for (int i = 0; i < s; ++i) {           /* L y = r */
    double acc = r[i];
    for (int k = 0; k < i; ++k) acc -= F[TRI(i, k)] * r[k];
    r[i] = acc / F[TRI(i, i)];
  }
  for (int i = s - 1; i >= 0; --i) {      /* L' d = y */
    double acc = r[i];
    for (int k = i + 1; k < s; ++k) acc -= F[TRI(k, i)] * r[k];
    r[i] = acc / F[TRI(i, i)];
  }
  for (int k = 0; k < s; ++k) x[idx[k]] += r[k];
}

static void schwarz_sweep(orc_hier* h, const orc_level* L, const double* b, double* x, int backward) {
  if (h->ordering == 0) {
    if (!backward) for (int p = 0; p < L->npatch; ++p) patch_solve(h, L, p, b, x);
    else for (int p = L->npatch - 1; p >= 0; --p) patch_solve(h, L, p, b, x);
    return;
  }
  for (int cc = 0; cc < L->npcolors; ++cc) {
    int c = backward ? L->npcolors - 1 - cc : cc;
#pragma omp parallel for schedule(dynamic, 16) num_threads(h->threads) if (h->threads > 1)
    for (int q = L->cpat_ptr[c]; q < L->cpat_ptr[c + 1]; ++q) patch_solve(h, L, L->cpat[q], b, x);
  }
}

/* pre-smoothing: Schwarz on the interface patches, then the point smoother on the other rows
 * (src/utils.py:84); post-smoothing: the adjoint order. */
static void smooth(orc_hier* h, int lev, const double* b, double* x, int post) {
  orc_level* L = &h->lv[lev];
  const int iters = post ? h->postsmooth : h->presmooth;
  const int mc = h->ordering == 1;
  for (int phase = 0; phase < 2; ++phase) {
    const int do_schwarz = post ? phase == 1 : phase == 0;
    if (do_schwarz) {
      if (L->npatch == 0) continue;
      int fwd, bwd;
      if (h->schwarz_type == SW_SYMMETRIC) fwd = bwd = 1;
      else if (h->schwarz_type == SW_FORWARD) { fwd = !post; bwd = post; }
      else { fwd = post; bwd = !post; }
      if (fwd) schwarz_sweep(h, L, b, x, 0);
      if (bwd) schwarz_sweep(h, L, b, x, 1);
    } else {
      for (int it = 0; it < iters; ++it) {
        switch (h->smoother) {
          case SM_JACOBI: jacobi(L, b, x, h->relaxation); break;
          case SM_L1DIAG: jacobi_l1(L, b, x); break;
          case SM_GS: gs_sweep(h, L, b, x, 1.0, post, 0); break;
          case SM_SOR: gs_sweep(h, L, b, x, h->relaxation, post, 0); break;
          case SM_SGS: gs_sweep(h, L, b, x, 1.0, 0, 0); gs_sweep(h, L, b, x, 1.0, 1, mc ? 1 : 0); break;
          case SM_SSOR: gs_sweep(h, L, b, x, h->relaxation, 0, 0); gs_sweep(h, L, b, x, h->relaxation, 1, 0); break;
        }
      }
    }
  }
}

static void coarse_solve(orc_hier* h) {
  orc_level* L = &h->lv[h->nlevels - 1];
  const int n = L->n;
  for (int i = 0; i < n; ++i) {
    double s = 0.0;
    for (int j = 0; j < n; ++j) s += h->coarse_inv[(size_t)i * n + j] * L->b[j];
    L->x[i] = s;
  }
}

/* b_c = R (b - A x) on level lev, x_c = 0 (aggregate sums for UA, P' for SA) */
static void restrict_residual(orc_hier* h, int lev) {
  orc_level* L = &h->lv[lev];
  orc_level* C = &h->lv[lev + 1];
  memset(C->b, 0, sizeof(double) * C->n);
  if (L->pia) {   /* SA_AMG: b_c = P' (b - A x) */
    for (int i = 0; i < L->n; ++i) {
      const double w = L->b[i] - row_dot(L, i, L->x);
      for (int p = L->pia[i]; p < L->pia[i + 1]; ++p) C->b[L->pja[p]] += L->pa[p] * w;
    }
  } else if (h->threads > 1) {
#pragma omp parallel for schedule(static) num_threads(h->threads)
    for (int i = 0; i < L->n; ++i) L->w[i] = L->b[i] - row_dot(L, i, L->x);
    for (int i = 0; i < L->n; ++i) { int I = L->agg[i]; if (I >= 0) C->b[I] += L->w[i]; }
  } else {
    for (int i = 0; i < L->n; ++i) {
      int I = L->agg[i];
      if (I >= 0) C->b[I] += L->b[i] - row_dot(L, i, L->x);
    }
  }
  memset(C->x, 0, sizeof(double) * C->n);
}

/* coarse scaling: alpha = min(1, e.rhs / e.A_c e) for the correction e = x_c */
static double coarse_alpha(orc_hier* h, const orc_level* C, const double* rhs) {
  if (!h->coarse_scaling) return 1.0;
  double num = 0.0, den = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : num, den) num_threads(h->threads) if (h->threads > 1)
  for (int i = 0; i < C->n; ++i) { num += C->x[i] * rhs[i]; den += C->x[i] * row_dot(C, i, C->x); }
  double alpha = num / den;
  return (alpha < 1.0) ? alpha : 1.0;  /* MIN(alpha, 1.0); NaN -> 1 */
}

/* x += alpha P x_c */
static void prolong_add(orc_hier* h, int lev, double alpha) {
  orc_level* L = &h->lv[lev];
  const orc_level* C = &h->lv[lev + 1];
  if (L->pia) {   /* SA_AMG: x += alpha P e */
    for (int i = 0; i < L->n; ++i) {
      double s = 0.0;
      for (int p = L->pia[i]; p < L->pia[i + 1]; ++p) s += L->pa[p] * C->x[L->pja[p]];
      L->x[i] += alpha * s;
    }
  } else {
    for (int i = 0; i < L->n; ++i) {
      int I = L->agg[i];
      if (I >= 0) L->x[i] += alpha * C->x[I];
    }
  }
}

/* FASP-lineage cycle: level l > 0 is visited cycle_type times per parent visit,
 * re-entering with its current iterate. */
static void cycle_level(orc_hier* h, int lev) {
  if (lev == h->nlevels - 1) { coarse_solve(h); return; }
  const int reps = (lev > 0 && h->cycle_type == W_CYCLE) ? 2 : 1;
  orc_level* L = &h->lv[lev];
  orc_level* C = &h->lv[lev + 1];
  for (int rep = 0; rep < reps; ++rep) {
    h->visits++;
    smooth(h, lev, L->b, L->x, 0);
    restrict_residual(h, lev);
    cycle_level(h, lev + 1);
    const double alpha = coarse_alpha(h, C, C->b);
    prolong_add(h, lev, alpha);
    smooth(h, lev, L->b, L->x, 1);
  }
}

static double* level_work(orc_level* L) {
  if (!L->kw) L->kw = (double*)calloc((size_t)4 * (L->n ? L->n : 1), sizeof(double));
  return L->kw;
}

/* Coefficients of the AMLI polynomial (FASP lineage, fasp_amg_amli_coef): q_k from the three-term recurrence
 * of the shifted Chebyshev polynomials on [lambda_min, lambda_max]; coef has degree + 1 entries. */
static void amli_coef(double lambda_max, double lambda_min, int degree, double* coef) {
  const double mu0 = 1.0 / lambda_max, mu1 = 1.0 / lambda_min;
  const double c = (sqrt(mu0) + sqrt(mu1)) * (sqrt(mu0) + sqrt(mu1));
  const double a = (4.0 * mu0 * mu1) / c;
  const double kappa = lambda_max / lambda_min;
  const double delta = (sqrt(kappa) - 1.0) / (sqrt(kappa) + 1.0);
  const double b = delta * delta;
  if (degree == 0) {
    coef[0] = 0.5 * (mu0 + mu1);
  } else if (degree == 1) {
    coef[0] = 0.5 * c;
    coef[1] = -1.0 * mu0 * mu1;
  } else {
    double ck[16] = {0}, ckm1[16] = {0};
    amli_coef(lambda_max, lambda_min, degree - 1, ck);
    amli_coef(lambda_max, lambda_min, degree - 2, ckm1);
    coef[0] = a - b * ckm1[0] + (1.0 + b) * ck[0];
    for (int i = 1; i < degree - 1; ++i) coef[i] = -b * ckm1[i] + (1.0 + b) * ck[i] - a * ck[i - 1];
    coef[degree - 1] = (1.0 + b) * ck[degree - 1] - a * ck[degree - 2];
    coef[degree] = -a * ck[degree - 1];
  }
}
void orc_amli_coef(int degree, double* coef) { amli_coef(2.0, 0.5, degree, coef); }

/* AMLI-cycle (FASP lineage, amli() of mgcycle.c): the coarse correction is the polynomial
 * q(B_c A_c) B_c r_c of degree amli_degree in the coarse cycle B_c, evaluated Horner-style:
 *   e = B_c r_c;  repeat degree times: rhs = A_c e + (coef[degree-i]/coef[degree]) r_c, e = B_c rhs;
 *   e *= coef[degree];  alpha = min(1, e.r_c / e.A_c e). */
static void cycle_amli(orc_hier* h, int lev) {
  if (lev == h->nlevels - 1) { coarse_solve(h); return; }
  orc_level* L = &h->lv[lev];
  orc_level* C = &h->lv[lev + 1];
  const int deg = h->amli_degree;
  const double* coef = h->amli_coef;
  h->visits++;
  smooth(h, lev, L->b, L->x, 0);
  restrict_residual(h, lev);
  double* r1 = level_work(C);
  memcpy(r1, C->b, sizeof(double) * C->n);
  for (int i = 1; i <= deg; ++i) {
    memset(C->x, 0, sizeof(double) * C->n);
    cycle_amli(h, lev + 1);
    const double s = coef[deg - i] / coef[deg];
    for (int k = 0; k < C->n; ++k) C->b[k] = row_dot(C, k, C->x);
    for (int k = 0; k < C->n; ++k) C->b[k] += s * r1[k];
  }
  memset(C->x, 0, sizeof(double) * C->n);
  cycle_amli(h, lev + 1);
  for (int k = 0; k < C->n; ++k) C->x[k] *= coef[deg];
  const double alpha = coarse_alpha(h, C, r1);
  prolong_add(h, lev, alpha);
  smooth(h, lev, L->b, L->x, 1);
}

static void cycle_nlamli(orc_hier* h, int lev);

/* K-cycle coarse correction (Notay & Vassilevski 2008; FASP lineage Kcycle_dcsr_pgcg / _pgcr): at most two steps of
 * a Krylov method on A_c x = b_c preconditioned by the nonlinear AMLI cycle of level lc itself; the second
 * step is skipped when the first reduced the residual below KCYCLE_TOL.  gcg: inner products with the
 * directions (flexible CG), otherwise with their images (GCR, HAZmath's default branch). */
#define KCYCLE_TOL2 0.04   /* relres < 0.2 */
static void kcycle(orc_hier* h, int lc) {
  orc_level* C = &h->lv[lc];
  const int n = C->n;
  const int gcg = h->nl_amli_krylov_type == SOLVER_GCG;
  double* bH = level_work(C);
  double *c1 = bH + n, *v1 = bH + 2 * (size_t)n, *v2 = bH + 3 * (size_t)n;
  double* r = C->b;   /* the Krylov residual lives in the level's right-hand side: it is what the cycle reads */
  memcpy(bH, C->b, sizeof(double) * n);
  const double nb2 = vdot(n, r, r);
  memset(C->x, 0, sizeof(double) * n);
  cycle_nlamli(h, lc);
  memcpy(c1, C->x, sizeof(double) * n);
  for (int i = 0; i < n; ++i) v1[i] = row_dot(C, i, c1);
  const double rho1 = vdot(n, gcg ? c1 : v1, v1);
  const double alpha1 = vdot(n, gcg ? c1 : v1, r);
  const double beta1 = rho1 != 0.0 ? alpha1 / rho1 : 0.0;
  for (int i = 0; i < n; ++i) r[i] += -beta1 * v1[i];
  const double nr2 = vdot(n, r, r);
  if (nb2 != 0.0) { double m = fabs(nr2 / (KCYCLE_TOL2 * nb2) - 1.0); if (m < h->kmargin) h->kmargin = m; if (m < h->kmargin_all) h->kmargin_all = m; }
  double beta3 = beta1, beta4 = 0.0;
  if (!(nr2 < KCYCLE_TOL2 * nb2) && nb2 != 0.0) {
    memset(C->x, 0, sizeof(double) * n);
    cycle_nlamli(h, lc);
    for (int i = 0; i < n; ++i) v2[i] = row_dot(C, i, C->x);
    const double* q = gcg ? C->x : v2;
    const double gamma = vdot(n, q, v1), alpha2 = vdot(n, q, v2), rho2 = vdot(n, q, r);
    const double beta2 = alpha2 - gamma * gamma / rho1;
    if (beta2 != 0.0 && rho1 != 0.0) {
      beta3 = (alpha1 - gamma * rho2 / beta2) / rho1;
      beta4 = rho2 / beta2;
    }
  }
  if (beta4 == 0.0) for (int i = 0; i < n; ++i) C->x[i] = beta3 * c1[i];
  else for (int i = 0; i < n; ++i) C->x[i] = beta3 * c1[i] + beta4 * C->x[i];
  memcpy(C->b, bH, sizeof(double) * n);
}

/* nonlinear AMLI-cycle (FASP lineage, nl_amli() of mgcycle.c): the coarse problem of every level but the
 * last is "solved" by the K-cycle above; the coarsest level is solved directly. */
static void cycle_nlamli(orc_hier* h, int lev) {
  if (lev == h->nlevels - 1) { coarse_solve(h); return; }
  orc_level* L = &h->lv[lev];
  orc_level* C = &h->lv[lev + 1];
  h->visits++;
  smooth(h, lev, L->b, L->x, 0);
  restrict_residual(h, lev);
  if (lev + 1 == h->nlevels - 1) coarse_solve(h);
  else kcycle(h, lev + 1);
  const double alpha = coarse_alpha(h, C, C->b);
  prolong_add(h, lev, alpha);
  smooth(h, lev, L->b, L->x, 1);
}

/* additive cycle: z = sum_l P_0..P_{l-1} S_l R_{l-1}..R_0 r with S_l = pre- then post-smoothing from a zero
 * iterate (a symmetric operator for the symmetric smoothers), the exact solve on the coarsest level and no
 * coarse scaling.  Frozen choice: HAZmath's additive cycle could not be consulted. */
static void cycle_add(orc_hier* h) {
  const int nl = h->nlevels;
  for (int lev = 0; lev + 1 < nl; ++lev) {
    orc_level* L = &h->lv[lev];
    h->visits++;
    memset(L->x, 0, sizeof(double) * L->n);
    restrict_residual(h, lev);          /* x = 0: the right-hand side itself is restricted */
    smooth(h, lev, L->b, L->x, 0);
    smooth(h, lev, L->b, L->x, 1);
  }
  coarse_solve(h);
  for (int lev = nl - 2; lev >= 0; --lev) prolong_add(h, lev, 1.0);
}

/* z = B r : haznics.apply_precond as called by cbc.block's Precond.matvec */
void orc_apply(orc_hier* h, const double* r, double* z) {
  orc_level* L0 = &h->lv[0];
  h->visits = 0;
  h->kmargin = 1e300;
  memcpy(L0->b, r, sizeof(double) * L0->n);
  if (h->nlevels == 1) { coarse_solve(h); memcpy(z, L0->x, sizeof(double) * L0->n); return; }
  memset(L0->x, 0, sizeof(double) * L0->n);
  if (h->cycle_type == ADD_CYCLE) { cycle_add(h); memcpy(z, L0->x, sizeof(double) * L0->n); return; }
  if (h->cycle_type == AMLI_CYCLE) amli_coef(2.0, 0.5, h->amli_degree, h->amli_coef);   /* lambda_max = 2, lambda_min = lambda_max / 4 */
  for (int it = 0; it < (h->maxit > 1 ? h->maxit : 1); ++it) {
    if (h->cycle_type == AMLI_CYCLE) cycle_amli(h, 0);
    else if (h->cycle_type == NL_AMLI_CYCLE) cycle_nlamli(h, 0);
    else cycle_level(h, 0);
  }
  memcpy(z, L0->x, sizeof(double) * L0->n);
}

void orc_spmv(orc_hier* h, int lev, const double* x, double* y) { orc_spmv_level(&h->lv[lev], x, y); }

void orc_smooth(orc_hier* h, int lev, const double* b, double* x, int post) { smooth(h, lev, b, x, post); }

/* cbc.block precondconjgrad.  Returns the iteration count;
 * residuals[0..niters], alphas/betas[0..niters-1]. */
int orc_pcg(orc_hier* h, const double* b, double* x, double tol, int relative, int maxiter,
            int use_guess, double* residuals, double* alphas, double* betas) {
  orc_level* L0 = &h->lv[0];
  const int n = L0->n;
  double* r = (double*)xmalloc(sizeof(double) * n);
  double* z = (double*)xmalloc(sizeof(double) * n);
  double* d = (double*)xmalloc(sizeof(double) * n);
  double* q = (double*)xmalloc(sizeof(double) * n);
  if (!use_guess) memset(x, 0, sizeof(double) * n);
  for (int i = 0; i < n; ++i) r[i] = b[i] - (use_guess ? row_dot(L0, i, x) : 0.0);
  orc_apply(h, r, z);
  memcpy(d, z, sizeof(double) * n);
  double rz = 0.0;
  for (int i = 0; i < n; ++i) rz += r[i] * z[i];
  double res = relative == 2 ? sqrt(vdot(n, r, r)) : sqrt(rz);
  residuals[0] = res;
  const double target = relative ? tol * res : tol;
  int it = 0;
  while (res > target && it < maxiter) {
    double dq = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : dq) num_threads(h->threads) if (h->threads > 1)
    for (int i = 0; i < n; ++i) { q[i] = row_dot(L0, i, d); dq += d[i] * q[i]; }
    double alpha = rz / dq;
    for (int i = 0; i < n; ++i) { x[i] += alpha * d[i]; r[i] -= alpha * q[i]; }
    orc_apply(h, r, z);
    double rz2 = 0.0;
    for (int i = 0; i < n; ++i) rz2 += r[i] * z[i];
    double beta = rz2 / rz;
    for (int i = 0; i < n; ++i) d[i] = z[i] + beta * d[i];
    rz = rz2;
    res = relative == 2 ? sqrt(vdot(n, r, r)) : sqrt(rz);
    if (alphas) alphas[it] = alpha;
    if (betas) betas[it] = beta;
    ++it;
    residuals[it] = res;
    if (!(rz >= 0.0) || !isfinite(alpha)) break;
  }
  free(r); free(z); free(d); free(q);
  return it;
}

/* ---- MINRES / GMRES: north_star names CG/MinRes/GMRES; block.iterative.MinRes and LGMRES share
 * ConjGrad's constructor upstream.  Standard algorithms (Paige-Saunders MINRES with an SPD
 * preconditioner, residual estimate in the B-norm; restarted right-preconditioned GMRES with
 * modified Gram-Schmidt), restated here so the device versions have an iteration-count oracle. */
static double vdot(int n, const double* u, const double* v) { double s = 0.0; for (int i = 0; i < n; ++i) s += u[i] * v[i]; return s; }

int orc_minres(orc_hier* h, const double* b, double* x, double tol, int relative, int maxiter, double* residuals) {
  orc_level* L0 = &h->lv[0];
  const int n = L0->n;
  double* buf = (double*)calloc((size_t)8 * n, sizeof(double));
  double *r1 = buf, *r2 = buf + n, *y = buf + 2 * n, *v = buf + 3 * n, *w = buf + 4 * n, *w1 = buf + 5 * n,
         *w2 = buf + 6 * n, *t = buf + 7 * n;
  memset(x, 0, sizeof(double) * n);
  memcpy(r1, b, sizeof(double) * n);
  memcpy(r2, b, sizeof(double) * n);
  orc_apply(h, r1, y);
  double beta1 = sqrt(vdot(n, r1, y));
  residuals[0] = beta1;
  const double target = relative ? tol * beta1 : tol;
  double oldb = 0.0, beta = beta1, dbar = 0.0, epsln = 0.0, phibar = beta1, cs = -1.0, sn = 0.0;
  int it = 0;
  while (phibar > target && it < maxiter) {
    ++it;
    for (int i = 0; i < n; ++i) v[i] = y[i] / beta;
    orc_spmv_level(L0, v, y);
    if (it >= 2) for (int i = 0; i < n; ++i) y[i] -= (beta / oldb) * r1[i];
    const double alfa = vdot(n, v, y);
    for (int i = 0; i < n; ++i) y[i] -= (alfa / beta) * r2[i];
    double* sw = r1; r1 = r2; r2 = sw;
    memcpy(r2, y, sizeof(double) * n);
    orc_apply(h, r2, y);
    oldb = beta;
    beta = sqrt(vdot(n, r2, y));
    const double oldeps = epsln, delta = cs * dbar + sn * alfa, gbar = sn * dbar - cs * alfa;
    epsln = sn * beta;
    dbar = -cs * beta;
    double gamma = sqrt(gbar * gbar + beta * beta);
    if (gamma < 1e-300) gamma = 1e-300;
    cs = gbar / gamma;
    sn = beta / gamma;
    const double phi = cs * phibar;
    phibar = sn * phibar;
    sw = w1; w1 = w2; w2 = w; w = sw;   /* w1 <- w2, w2 <- w, w <- scratch */
    for (int i = 0; i < n; ++i) t[i] = (v[i] - oldeps * w1[i] - delta * w2[i]) / gamma;
    sw = w; w = t; t = sw;
    for (int i = 0; i < n; ++i) x[i] += phi * w[i];
    residuals[it] = phibar;
  }
  free(buf);
  return it;
}

int orc_gmres(orc_hier* h, const double* b, double* x, double tol, int relative, int maxiter, int m, double* residuals) {
  orc_level* L0 = &h->lv[0];
  const int n = L0->n;
  if (m < 1) m = 30;
  double* V = (double*)calloc((size_t)(m + 1) * n, sizeof(double));
  double* buf = (double*)calloc((size_t)4 * n, sizeof(double));
  double *r = buf, *z = buf + n, *w = buf + 2 * n, *u = buf + 3 * n;
  double* Hm = (double*)calloc((size_t)(m + 1) * m, sizeof(double));
  double* cs = (double*)calloc(m, sizeof(double));
  double* sn = (double*)calloc(m, sizeof(double));
  double* g = (double*)calloc(m + 1, sizeof(double));
  double* yv = (double*)calloc(m, sizeof(double));
  memset(x, 0, sizeof(double) * n);
  memcpy(r, b, sizeof(double) * n);
  double rn = sqrt(vdot(n, r, r));
  residuals[0] = rn;
  const double target = relative ? tol * rn : tol;
  int it = 0;
  while (rn > target && it < maxiter) {
    for (int i = 0; i < n; ++i) V[i] = r[i] / rn;
    memset(g, 0, sizeof(double) * (m + 1));
    g[0] = rn;
    int j = 0;
    for (; j < m && it < maxiter && rn > target; ++j) {
      orc_apply(h, V + (size_t)j * n, z);
      orc_spmv_level(L0, z, w);
      for (int i = 0; i <= j; ++i) {
        const double hh = vdot(n, w, V + (size_t)i * n);
        Hm[(size_t)i * m + j] = hh;
        for (int k = 0; k < n; ++k) w[k] -= hh * V[(size_t)i * n + k];
      }
      const double hn = sqrt(vdot(n, w, w));
      Hm[(size_t)(j + 1) * m + j] = hn;
      if (hn > 0.0) for (int k = 0; k < n; ++k) V[(size_t)(j + 1) * n + k] = w[k] / hn;
      for (int i = 0; i < j; ++i) {
        const double t0 = cs[i] * Hm[(size_t)i * m + j] + sn[i] * Hm[(size_t)(i + 1) * m + j];
        Hm[(size_t)(i + 1) * m + j] = -sn[i] * Hm[(size_t)i * m + j] + cs[i] * Hm[(size_t)(i + 1) * m + j];
        Hm[(size_t)i * m + j] = t0;
      }
      const double den = hypot(Hm[(size_t)j * m + j], hn);
      cs[j] = den > 0 ? Hm[(size_t)j * m + j] / den : 1.0;
      sn[j] = den > 0 ? hn / den : 0.0;
      Hm[(size_t)j * m + j] = den;
      g[j + 1] = -sn[j] * g[j];
      g[j] = cs[j] * g[j];
      rn = fabs(g[j + 1]);
      ++it;
      residuals[it] = rn;
    }
    for (int i = j - 1; i >= 0; --i) {
      double sacc = g[i];
      for (int k = i + 1; k < j; ++k) sacc -= Hm[(size_t)i * m + k] * yv[k];
      yv[i] = sacc / Hm[(size_t)i * m + i];
    }
    memset(u, 0, sizeof(double) * n);
    for (int i = 0; i < j; ++i) for (int k = 0; k < n; ++k) u[k] += yv[i] * V[(size_t)i * n + k];
    orc_apply(h, u, z);
    for (int k = 0; k < n; ++k) x[k] += z[k];
    for (int k = 0; k < n; ++k) r[k] = b[k] - row_dot(L0, k, x);
    rn = sqrt(vdot(n, r, r));
    residuals[it] = rn;
  }
  free(V); free(buf); free(Hm); free(cs); free(sn); free(g); free(yv);
  return it;
}

double orc_now(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

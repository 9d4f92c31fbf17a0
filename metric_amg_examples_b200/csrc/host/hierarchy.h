// Host-side hierarchy of the metric-AMG preconditioner (natural ordering).
//
// What metricAMG(A, W, idofs=, parameters=) builds inside HAZmath when the
// reference calls it (reference src/utils.py:86): level operators A_l,
// unsmoothed-aggregation maps, Schwarz patches on the first Schwarz_levels
// levels and the factorised coarsest matrix.  The reference keeps this on the
// CPU; so do we (BASELINE.json north_star: "AMG setup ... may remain ... CPU
// setup exported once per problem").  Everything the device path and the
// oracle consume is exported from these structs, so both run on one hierarchy.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "../../../include/mamg.h"

namespace mamg {

// std::vector whose resize() leaves the new elements uninitialised: the entry arrays of the level matrices
// hold up to 1.5e9 entries and are always filled completely (in parallel) right after they are sized, so a
// value-initialising resize would be a single-threaded pass over 18 GB for nothing.
template <class T>
struct uninit_alloc : std::allocator<T> {
  template <class U> struct rebind { using other = uninit_alloc<U>; };
  template <class U, class... Args>
  void construct(U* p, Args&&... args) {
    if constexpr (sizeof...(Args) == 0) ::new (static_cast<void*>(p)) U;
    else ::new (static_cast<void*>(p)) U(std::forward<Args>(args)...);
  }
};
template <class T> using bigvec = std::vector<T, uninit_alloc<T>>;

struct Csr {
  int n = 0;     // rows
  int m = 0;     // columns
  std::vector<int> ia;
  bigvec<int> ja;
  bigvec<double> a;
  int nnz() const { return ia.empty() ? 0 : ia[n]; }
};

struct SchwarzPatches {
  // CSR-of-lists in natural numbering of the level (HAZmath iblock/jblock).
  std::vector<int> ptr;    // npatch+1
  std::vector<int> dofs;   // sorted ascending inside a patch
  std::vector<int> seed;   // seed dof of every patch
  std::vector<int> color;  // conflict colour of every patch
  int ncolors = 0;
  int max_size = 0;
  int npatch() const { return ptr.empty() ? 0 : (int)ptr.size() - 1; }
};

struct Level {
  Csr A;                       // natural ordering
  int64_t nnz_structural = 0;  // entries before explicit zeros were dropped (setup.cpp)
  std::vector<int> agg;        // size n; coarse index or -1 (row left out of every aggregate)
  int nc = 0;                  // number of aggregates == rows of next level
  std::vector<int> color;      // multicolour GS colour of every row
  int ncolors = 0;
  std::vector<uint8_t> gs_skip;  // 1: row is smoothed by Schwarz, not by GS (level < Schwarz_levels)
  SchwarzPatches sw;           // empty unless level < Schwarz_levels
  Csr P, R;                    // only for SA_AMG: smoothed prolongator (n x nc) and R = P' (nc x n)
  std::vector<int> part;       // owner part of every row (empty: not partitioned); aggregates never cross parts
};

struct Hierarchy {
  mamg_params prm;
  int nparts = 1;
  bool released = false;       // host matrices freed after the device upload
  std::vector<Level> lv;
  std::vector<double> coarse_inv;  // dense row-major inverse of the coarsest A (n_c x n_c)
  double setup_seconds = 0;
};

// ---- setup pieces (each file states the reference lines it restates) -------
void aggregate_hem(const Csr& A, const int* part, std::vector<int>& agg, int& nc);
void aggregate_vmb(const Csr& A, const int* part, double strong, int max_agg, std::vector<int>& agg, int& nc);
void aggregate_hec(const Csr& A, const int* part, int max_agg, std::vector<int>& agg, int& nc);
void aggregate_mwm(const Csr& A, const int* part, std::vector<int>& agg, int& nc);
void aggregate_mis(const Csr& A, const int* part, double strong, std::vector<int>& agg, int& nc);
void galerkin_ua(const Csr& A, const std::vector<int>& agg, int nc, Csr& Ac);
void csr_transpose(const Csr& A, Csr& At);
void csr_drop_zeros(Csr& A);                              // removes exact zeros off the diagonal
void csr_multiply(const Csr& A, const Csr& B, Csr& C);   // C = A B, columns sorted
void smoothed_prolongator(const Csr& A, const std::vector<int>& agg, int nc, double omega, Csr& P);
void multicolor_greedy(const Csr& A, const std::vector<uint8_t>& skip, std::vector<int>& color, int& ncolors);
void schwarz_patches(const Csr& A, const int* seeds, int nseeds, int maxlvl, int mmsize,
                     SchwarzPatches& out);
void schwarz_color(const Csr& A, SchwarzPatches& sw);
bool dense_inverse(const Csr& A, std::vector<double>& inv);
bool validate_params(const mamg_params& prm, std::string& err);
bool build_hierarchy(const mamg_params& prm, Csr&& A0, const int* idofs, int n_idofs,
                     const int* part, int nparts, Hierarchy& H, std::string& err);

// structured P1 problems (assemble.cpp)
void p1_scalar(int dim, const int* ncell, const double* h, double cK, double cM, Csr& out);
void assemble_bidomain(int dim, int n, double k1, double k2, double g, Csr& out);
void assemble_emi(int dim, int n, double k1, double k2, double g, Csr& out);

// error text shared by every C-ABI entry point (capi.cpp)
void set_error(const std::string& msg);

}  // namespace mamg

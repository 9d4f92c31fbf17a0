// Device half of the C-ABI: upload of the hierarchy (colour-permuted, sliced-ELL rows, shared Schwarz blobs),
// the multigrid cycle, the Krylov loops, the multi-GPU exchange.
//
// Replaces, on one B200, what the reference runs on the CPU for every B*r inside
// ConjGrad (src/bidomain_2d.py:205-206): haznics.apply_precond -> precond_amg -> mgcycle.
// The cycle is the FASP-lineage V/W recursion: pre-smooth, residual, restrict,
// recurse (twice per visit below level 0 for W), coarse scaling, prolong, post-smooth, dense
// solve on the coarsest level.  All scalars (alpha of the coarse scaling, alpha/beta of CG)
// stay on the device, so one apply is a fixed launch sequence without host synchronisation.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include <nccl.h>
#include <omp.h>
#include <nvtx3/nvToolsExt.h>

#include "../host/handle.h"
#include "kernels.cuh"
#include "schwarz.cuh"
#include "sell.cuh"
#include "tail.cuh"

namespace mamg {

#define CUDA_OK(call)                                                                     \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess)                                                                \
      throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " + \
                               __FILE__ + ":" + std::to_string(__LINE__));                \
  } while (0)

struct DCsr {
  int n = 0, m = 0, nnz = 0, lanes = 4;
  int *ia = nullptr, *ja = nullptr;
  double* a = nullptr;
};

struct DLevel {
  int n = 0, nnz = 0, nc = 0, ncolors = 0, lanes = 8, unroll = 4;
  int *ia = nullptr, *ja = nullptr;
  double *a = nullptr, *invd = nullptr;
  double* invl1 = nullptr;       // 1 / sum_j |a_ij| (SMOOTHER_L1DIAG only)
  uint8_t* skip = nullptr;
  int *agg = nullptr, *cptr = nullptr, *cidx = nullptr;
  double *x = nullptr, *b = nullptr, *t = nullptr;
  double *x_own = nullptr, *b_own = nullptr;
  int *perm = nullptr, *iperm = nullptr;  // device: new->old, old->new
  // Row layout: nb blocks (1, or the parts of a partitioned hierarchy on levels big enough to be
  // row-distributed), inside a block the colours, inside a colour the natural order.
  int nb = 1;
  std::vector<int> bc_ptr;                // host: row offset of (block b, colour c) at [b*ncolors + c]; size nb*ncolors+1
  std::vector<int> color_active;          // host: rows of (b, c) that the point smoother touches (empty: all)
  int row0(int b, int c) const { return bc_ptr[b * ncolors + c]; }
  int row1(int b, int c) const { return bc_ptr[b * ncolors + c + 1]; }
  DSchwarz sw;
  DCsr P, R;   // SA_AMG only
  int* d_color_ptr = nullptr;   // device copy of bc_ptr (tail kernel)
  // sliced-ELL copy of the level matrix (sell.cuh): the layout the row kernels stream; the CSR arrays
  // ja / a are released after the conversion unless something else reads them (tail kernel, Schwarz)
  SellView S;
  bool use_sell = false, has_csr = true;
  int64_t sell_slots = 0;       // stored entry slots (padding included)
  // multi-GPU halo lists of a row-distributed level (halo mode): the rows of MY blocks that a row of
  // another rank couples to, per (colour, neighbour rank); colour == ncolors holds all colours together
  std::vector<int> nbr_ranks;   // ranks that share a boundary with this rank on this level
  std::vector<int> send_off;    // [(ncolors + 1) * nnb + 1] ranges inside d_send
  int* d_send = nullptr;
  bool rows_owned_only = false; // matrix rows (CSR / sliced ELL) are stored for this rank's blocks only
  // transition to the replicated levels: the rows of the (replicated) coarse level below whose aggregates I own
  int* d_own_coarse = nullptr;
  int n_own_coarse = 0;
  // Schwarz colours in halo mode: the dofs my patches of colour c updated that neighbour k gathers or owns
  std::vector<int> sw_send_off;   // [ncolors_sw * nnb + 1]
  int* d_sw_send = nullptr;
};

enum KClass { K_SPMV = 0, K_GS, K_SCHWARZ, K_RESTRICT, K_SCALE, K_PROLONG, K_COARSE, K_VEC, K_DOT, K_EXCH, K_NCLS };

struct ProfEvent { cudaEvent_t start, stop; int cls; int lev; };

// pinned bounce buffers of the upload (see Stager::copy below)
struct Stager {
  void* buf[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  cudaStream_t stream = nullptr;
  size_t chunk = 0, min_bytes = 0;
  bool ok = false;
  void init();
  void release();
  void copy(void* dst, const void* src, size_t bytes);
};

struct DeviceState {
  Stager stager;
  bool prof_on = false;
  int cur_level = 0;          // level the cycle is working on (profiling breakdown only)
  std::vector<ProfEvent> prof_events;
  size_t prof_used = 0;
  int64_t cls_launches[K_NCLS] = {0};
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::vector<DLevel> lv;
  double* coarse_inv = nullptr;
  double* partial = nullptr;
  unsigned int* ticket = nullptr;
  double* scal = nullptr;     // device scalars: [0..7] PCG, [8..10] coarse scaling
  double* h_scal = nullptr;   // pinned mirror
  int red_blocks = 0;
  int64_t launches = 0;
  int64_t dev_bytes = 0;
  double* w[10] = {nullptr};  // Krylov work vectors (level-0 size)
  std::vector<double*> basis; // GMRES Krylov basis, allocated on first use
  double* hcoef = nullptr;    // GMRES: one column of the Hessenberg matrix on the device
  double *io_a = nullptr, *io_b = nullptr;  // staging for host-array calls
  std::vector<void*> allocs;
  mamg_params prm;
  // multi-GPU (one process per GPU): rank r executes the rows of its blocks on the row-distributed levels.
  // Halo mode (default): the matrices are stored per rank, vectors keep their global length and are valid on
  // the owned rows + halo, updates travel as halo index lists to the neighbour ranks (peer memory, CUDA IPC).
  // Round-1 mode (MAMG_HALO=0): hierarchy and vectors complete on every rank, updated ranges all-gathered.
  int rank = 0, world = 1;
  ncclComm_t comm = nullptr;
  int64_t collectives = 0;
  int64_t exch_bytes = 0;     // bytes this rank stored into peers (halo mode)
  double* xbuf = nullptr;     // staging for the Schwarz patch-dof exchange
  // Every vector that is ever exchanged between ranks lives in ONE allocation (the arena) that is
  // shared with the peer processes through CUDA IPC: the owner of a row range stores it straight
  // into the peers' copies over NVLink and raises a flag there (no NCCL call, no host hop).
  double* arena = nullptr;
  size_t arena_doubles = 0;
  std::vector<double*> peer_arena;   // arena base of every rank (own entry = arena)
  double** d_peer_arena = nullptr;   // device copy
  unsigned int* push_ticket = nullptr;
  long long* d_phase = nullptr;      // exchanges so far (device counter; the same sequence on every rank)
  bool use_p2p = false;
  // halo mode (default with several ranks and peer memory): matrices are stored for the owned rows only,
  // a kernel's updates travel to the neighbour ranks that read them (index lists, neighbour-only flags),
  // dots are per-rank partial sums combined in rank order.  MAMG_HALO=0: the round-1 scheme (every
  // updated range all-gathered to every rank).
  bool halo = false;
  size_t xcap = 0;
  long long xflip = 0;
  // L2 residency: the vector that a smoother / SpMV gathers from is marked "persisting" in a
  // set-aside part of the 126 MB L2, so that the matrix and patch streams cannot evict it between
  // (and inside) the per-colour launches; one access-policy window per stream, moved when the
  // gathered vector changes (captured into the graph nodes)
  size_t l2_persist_bytes = 0;
  const void* l2_window = nullptr;
  int tail_k0 = -1;           // first level executed by the single-CTA tail kernel (-1: none)
  // AMLI / nonlinear AMLI cycles: per level >= 1 three work vectors (saved right-hand side, first direction and
  // its image) and 16 scalars of the K-cycle, allocated when such a cycle is first applied
  std::vector<double*> kwork;
  double* kscal = nullptr;
  double amli_coef[16] = {0};
  TailArgs tail;
  size_t tail_smem = 0;
  // one apply is a fixed launch sequence: it is captured once per (input, output) pair into a
  // CUDA graph and replayed, which removes the host launch cost of the ~1000 small kernels of the
  // coarse levels (MAMG_GRAPH=0 disables; not used while profiling or for very long W sequences)
  struct GraphEntry { const double* r; double* z; cudaGraphExec_t exec; int64_t launches; int64_t cls[K_NCLS]; int64_t coll, xbytes; };
  std::vector<GraphEntry> graphs;
  bool use_graph = true;
  bool graph_dist = true;            // capture the cycle with several ranks too (MAMG_GRAPH_DIST=0 disables)
  bool capturing = false;
  bool nvtx = false;                 // MAMG_NVTX=1: NVTX ranges around applies, level visits and Krylov solves
};

// NVTX range on the host timeline (ncu --nvtx --nvtx-include "mamg:smooth L0/" selects the kernels launched inside;
// during graph capture the ranges bracket the capture, not the replay)
struct Nvtx {
  bool on;
  Nvtx(const DeviceState& D, const char* what, int level = -1) : on(D.nvtx) {
    if (!on) return;
    char buf[64];
    if (level >= 0) snprintf(buf, sizeof buf, "mamg:%s L%d", what, level);
    else snprintf(buf, sizeof buf, "mamg:%s", what);
    nvtxRangePushA(buf);
  }
  ~Nvtx() { if (on) nvtxRangePop(); }
  Nvtx(const Nvtx&) = delete;
  Nvtx& operator=(const Nvtx&) = delete;
};

// Brackets one kernel launch: counts it and, in profiling mode, times it with a CUDA event pair
// on the launching stream (bench.py's per-kernel durations and shares come from here).
struct KScope {
  DeviceState& D;
  cudaEvent_t stop = nullptr;
  KScope(DeviceState& d, int cls) : D(d) {
    ++D.launches;
    ++D.cls_launches[cls];
    if (!D.prof_on) return;
    if (D.prof_used == D.prof_events.size()) {
      ProfEvent e;
      cudaEventCreate(&e.start);
      cudaEventCreate(&e.stop);
      D.prof_events.push_back(e);
    }
    ProfEvent& e = D.prof_events[D.prof_used++];
    e.cls = cls;
    e.lev = D.cur_level;
    cudaEventRecord(e.start, D.stream);
    stop = e.stop;
  }
  ~KScope() { if (stop) cudaEventRecord(stop, D.stream); }
};

template <class T>
static T* dalloc(DeviceState& D, size_t count) {
  T* p = nullptr;
  if (count == 0) count = 1;
  CUDA_OK(cudaMalloc(&p, count * sizeof(T)));
  D.allocs.push_back(p);
  D.dev_bytes += (int64_t)(count * sizeof(T));
  return p;
}
template <class T>
static void dfree(DeviceState& D, T* p, size_t count) {
  if (!p) return;
  auto it = std::find(D.allocs.begin(), D.allocs.end(), (void*)p);
  if (it != D.allocs.end()) D.allocs.erase(it);
  cudaFree(p);
  D.dev_bytes -= (int64_t)(count * sizeof(T));
}
// Large host -> device copies of the upload (the level matrices: 17 GB at emi_3d n=464) go through two pinned
// bounce buffers that all cores fill while the previous chunk is on the bus: a plain cudaMemcpy from pageable
// memory is staged by one driver thread (measured 4.7 GB/s).  MAMG_STAGE_MIN_KB: smallest copy that is staged
// (default 256 MB; 0 stages everything), MAMG_STAGE_CHUNK_KB: chunk size (default 64 MB; 0 disables staging).
void Stager::init() {
  const char* ec = getenv("MAMG_STAGE_CHUNK_KB");
  const char* em = getenv("MAMG_STAGE_MIN_KB");
  chunk = (ec ? (size_t)atoll(ec) : (size_t)65536) << 10;
  min_bytes = (em ? (size_t)atoll(em) : (size_t)262144) << 10;
  if (chunk == 0) return;
  for (int k = 0; k < 2; ++k) {
    if (cudaMallocHost(&buf[k], chunk) != cudaSuccess) { cudaGetLastError(); release(); return; }
    if (cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); release(); return; }
  }
  if (cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); release(); return; }
  ok = true;
}
void Stager::release() {
  for (int k = 0; k < 2; ++k) {
    if (buf[k]) cudaFreeHost(buf[k]);
    if (ev[k]) cudaEventDestroy(ev[k]);
    buf[k] = nullptr;
    ev[k] = nullptr;
  }
  if (stream) cudaStreamDestroy(stream);
  stream = nullptr;
  ok = false;
}
void Stager::copy(void* dst, const void* src, size_t bytes) {
  size_t off = 0;
  int k = 0;
  while (off < bytes) {
    const size_t len = std::min(chunk, bytes - off);
    CUDA_OK(cudaEventSynchronize(ev[k]));   // the copy that last read this buffer is done (no-op before its first use)
    const char* from = (const char*)src + off;
    char* to = (char*)buf[k];
    const size_t piece = (size_t)1 << 20;
    const long long pieces = (long long)((len + piece - 1) / piece);
#pragma omp parallel for schedule(static)
    for (long long q = 0; q < pieces; ++q) {
      const size_t o = (size_t)q * piece;
      std::memcpy(to + o, from + o, std::min(piece, len - o));
    }
    CUDA_OK(cudaMemcpyAsync((char*)dst + off, to, len, cudaMemcpyHostToDevice, stream));
    CUDA_OK(cudaEventRecord(ev[k], stream));
    off += len;
    k ^= 1;
  }
  CUDA_OK(cudaStreamSynchronize(stream));
}

template <class V>
static typename V::value_type* upload(DeviceState& D, const V& v) {
  using T = typename V::value_type;
  T* p = dalloc<T>(D, v.size());
  const size_t bytes = v.size() * sizeof(T);
  if (bytes == 0) return p;
  if (D.stager.ok && bytes >= D.stager.min_bytes) D.stager.copy(p, v.data(), bytes);
  else CUDA_OK(cudaMemcpy(p, v.data(), bytes, cudaMemcpyHostToDevice));
  return p;
}

void device_state_free(DeviceState* D) {
  if (!D) return;
  cudaSetDevice(D->device);
  if (D->stream) cudaStreamSynchronize(D->stream);
  for (void* p : D->allocs) cudaFree(p);
  for (ProfEvent& e : D->prof_events) { cudaEventDestroy(e.start); cudaEventDestroy(e.stop); }
  for (auto& g : D->graphs) cudaGraphExecDestroy(g.exec);
  if (D->comm) ncclCommDestroy(D->comm);
  for (int q = 0; q < (int)D->peer_arena.size(); ++q)
    if (q != D->rank && D->peer_arena[q]) cudaIpcCloseMemHandle(D->peer_arena[q]);
  if (D->h_scal) cudaFreeHost(D->h_scal);
  if (D->own_stream && D->stream) cudaStreamDestroy(D->stream);
  delete D;
}

static int pick_lanes(double avg) {
  const char* env = getenv("MAMG_LANES");
  if (env && atoi(env) > 0) return atoi(env);
  int l = 2;
  while (l < 32 && l * 4 <= avg + 2) l *= 2;  // about 2-4 entries per lane: 30/row -> 8, 15 -> 4, 7 -> 2
  return l;
}

static int dist_min_rows(bool halo = false) {
  const char* env = getenv("MAMG_DIST_MIN_ROWS");
  // smaller levels are executed redundantly by every rank (no communication).  Round-1 scheme, measured on
  // 4 x B200 at 16 M DOFs: 1 M rows -> 994 ms per solve, 100 k rows -> 1076 ms (~35 us per exchange).  Halo mode:
  // a level visit costs ~24 exchanges of ~12 us (measured at C4 on 2 GPUs), which a level of ~420 B/row of
  // traffic per visit only wins back above ~5 M rows
  return env ? atoi(env) : (halo ? 6000000 : 1000000);
}

static int pick_unroll(int n) {
  const char* env = getenv("MAMG_UNROLL");
  if (env && atoi(env) > 0) return atoi(env) >= 4 ? 4 : atoi(env);
  // measured on B200 (bidomain_3d, 16 M DOFs): 2 rows in flight per sub-warp beat 4 (SpMV 4.16 vs 3.13 TB/s,
  // GS 2.09 vs 2.01) -- twice the CTAs per colour launch matter more than the extra loads in flight;
  // small levels need the parallelism more than the MLP
  return n >= (1 << 15) ? 2 : 1;
}

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

constexpr size_t kArenaHead = 64 + 2 * 64 * 4;

static bool rows_sell() {
  const char* env = getenv("MAMG_ROWS");   // "csr": the sub-warp-per-row CSR kernels of round 1 (kept for comparison)
  return !(env && std::string(env) == "csr");
}

// sliced-ELL copy of a level, filled on the device from the permuted CSR that was just uploaded
// (rows outside [row_lo, row_hi) -- the blocks of other ranks in halo mode -- get empty slices)
static void build_sell(DeviceState& D, DLevel& dl, const std::vector<int>& ia, int row_lo, int row_hi) {
  const int n = dl.n, ns = (n + 31) / 32;
  std::vector<int> sp(ns + 1, 0);
  for (int s = 0; s < ns; ++s) {
    int w = 0;
    for (int i = std::max(s * 32, row_lo); i < std::min(std::min(n, s * 32 + 32), row_hi); ++i) w = std::max(w, ia[i + 1] - ia[i]);
    sp[s + 1] = w;
  }
  int64_t tot = 0;
  for (int s = 0; s < ns; ++s) {
    tot += sp[s + 1];
    if (tot > 0x7fffffffLL) throw std::runtime_error("sliced-ELL slot count exceeds int32");
    sp[s + 1] = (int)tot;
  }
  dl.sell_slots = tot * 32;
  int* d_sp = upload(D, sp);
  double* val = dalloc<double>(D, (size_t)tot * 32 + 2);
  int* col = dalloc<int>(D, (size_t)tot * 32 + 4);
  if (ns > 0) {
    sell_fill_kernel<<<(ns + kSellWarps - 1) / kSellWarps, kBlock>>>(n, row_lo, row_hi, dl.ia, dl.ja, dl.a, d_sp, val, col);
    CUDA_OK(cudaGetLastError());
  }
  dl.S.sp = d_sp;
  dl.S.val = val;
  dl.S.col = col;
  dl.S.n = n;
  dl.use_sell = true;
}

// Halo lists of a row-distributed level: my rows that a row of rank q couples to (the pattern is
// symmetric -- validated at setup -- so these are my rows with a column owned by q), per colour.
// On a level with Schwarz patches the halo is deeper: a patch seeded in another rank's block gathers x up to
// Schwarz_maxlvl + 1 rings into my block (dl.sw.readers) and reads b on its own dofs there (dl.sw.members),
// so those rows travel with their colour as well; list ncolors + 1 holds the rows whose right-hand side the
// other ranks' patches need.
static void build_halo_lists(DeviceState& D, DLevel& dl, const std::vector<int>& ia, const bigvec<int>& ja) {
  const int per = dl.nb / D.world, nc = dl.ncolors;
  std::vector<int> rank_lo(D.world + 1);
  for (int q = 0; q <= D.world; ++q) rank_lo[q] = dl.bc_ptr[q * per * nc];
  const int lo = rank_lo[D.rank], hi = rank_lo[D.rank + 1];
  auto owner = [&](int j) { return (int)(std::upper_bound(rank_lo.begin(), rank_lo.end(), j) - rank_lo.begin()) - 1; };
  // colour of my row i: position inside its (block, colour) ranges
  std::vector<std::vector<std::vector<int>>> lists(D.world, std::vector<std::vector<int>>(nc + 2));
  std::vector<char> hit(D.world);
  const bool deep = !dl.sw.readers.empty();
  int kb = D.rank * per * nc;
  for (int i = lo; i < hi; ++i) {
    while (i >= dl.bc_ptr[kb + 1]) ++kb;
    const int c = kb % nc;
    std::fill(hit.begin(), hit.end(), 0);
    auto add = [&](int q) {
      if (q != D.rank && !hit[q]) { hit[q] = 1; lists[q][c].push_back(i); lists[q][nc].push_back(i); }
    };
    for (int p = ia[i]; p < ia[i + 1]; ++p) {
      const int j = ja[p];
      if (j >= lo && j < hi) continue;
      add(owner(j));
    }
    if (deep) {
      const unsigned long long rd = dl.sw.readers[i] | dl.sw.members[i], mb = dl.sw.members[i];
      for (int part = 0; part < dl.nb; ++part) {
        if ((rd >> part) & 1ull) add(part / per);
        if (((mb >> part) & 1ull) && part / per != D.rank) {
          std::vector<int>& v = lists[part / per][nc + 1];
          if (v.empty() || v.back() != i) v.push_back(i);
        }
      }
    }
  }
  dl.nbr_ranks.clear();
  for (int q = 0; q < D.world; ++q)
    if (q != D.rank && (!lists[q][nc].empty() || !lists[q][nc + 1].empty())) dl.nbr_ranks.push_back(q);
  // the neighbour relation must be symmetric across ranks; it is, because the pattern is symmetric
  const int nnb = (int)dl.nbr_ranks.size();
  if (nnb > 8) throw std::runtime_error("more than 8 neighbour ranks on a level: use MAMG_HALO=0");
  std::vector<int> flat;
  dl.send_off.assign((size_t)(nc + 2) * nnb + 1, 0);
  for (int c = 0; c <= nc + 1; ++c)
    for (int k = 0; k < nnb; ++k) {
      const std::vector<int>& v = lists[dl.nbr_ranks[k]][c];
      flat.insert(flat.end(), v.begin(), v.end());
      dl.send_off[(size_t)c * nnb + k + 1] = (int)flat.size();
    }
  dl.d_send = upload(D, flat);
  if (deep) {
    // per Schwarz colour and neighbour: the exported dofs (schwarz_upload's export lists) that neighbour needs
    const DSchwarz& sw = dl.sw;
    std::vector<int> sflat;
    dl.sw_send_off.assign((size_t)sw.ncolors * nnb + 1, 0);
    for (int c = 0; c < sw.ncolors; ++c)
      for (int k = 0; k < nnb; ++k) {
        unsigned long long rmask = 0;
        for (int part = dl.nbr_ranks[k] * per; part < (dl.nbr_ranks[k] + 1) * per; ++part) rmask |= 1ull << part;
        for (int q = sw.xoff[c * sw.nb + D.rank * per]; q < sw.xoff[c * sw.nb + (D.rank + 1) * per]; ++q)
          if (sw.h_xmask[q] & rmask) sflat.push_back(sw.h_xidx[q]);
        dl.sw_send_off[(size_t)c * nnb + k + 1] = (int)sflat.size();
      }
    dl.d_sw_send = upload(D, sflat);
  }
}

// ------------------------------------------------------------------------------------------
// upload: colour-permute every level (rows of one colour contiguous, stable inside a colour)
// ------------------------------------------------------------------------------------------
static void upload_hierarchy(const Hierarchy& H, DeviceState& D) {
  const int L = (int)H.lv.size();
  // MAMG_SETUP_TIMING: per-phase host seconds of the upload on stderr (as build_hierarchy does for the setup)
  const bool timing = getenv("MAMG_SETUP_TIMING") != nullptr;
  auto tp = std::chrono::steady_clock::now();
  auto lap = [&](const char* what, int lev) {
    if (!timing) return;
    cudaDeviceSynchronize();
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[mamg upload] level %d %-14s %.3f s\n", lev, what, std::chrono::duration<double>(now - tp).count());
    tp = now;
  };
  D.stager.init();
  struct StagerGuard { Stager& s; ~StagerGuard() { s.release(); } } stager_guard{D.stager};
  D.lv.resize(L);
  std::vector<std::vector<int>> perm(L), iperm(L);
  for (int l = 0; l < L; ++l) {
    const Level& hl = H.lv[l];
    const int n = hl.A.n;
    perm[l].resize(n);
    iperm[l].resize(n);
    DLevel& dl = D.lv[l];
    if (hl.color.empty()) {
      std::iota(perm[l].begin(), perm[l].end(), 0);
      dl.bc_ptr = {0, n};
      dl.ncolors = 1;
      dl.nb = 1;
    } else {
      dl.ncolors = hl.ncolors;
      dl.nb = (H.nparts > 1 && !hl.part.empty() && n >= dist_min_rows(D.halo && D.world > 1)) ? H.nparts : 1;
      const int nbc = dl.nb * dl.ncolors;
      auto key = [&](int i) { return (dl.nb > 1 ? hl.part[i] : 0) * dl.ncolors + hl.color[i]; };
      // stable counting sort by (part, colour) on all cores: every thread counts and later places one contiguous
      // chunk of rows, the chunks in thread order inside every bucket (the same permutation as a serial pass)
      dl.bc_ptr.assign(nbc + 1, 0);
      const int tmax = std::max(1, omp_get_max_threads());
      std::vector<int> cnt((size_t)tmax * nbc, 0);
#pragma omp parallel num_threads(tmax)
      {
        const int nt = omp_get_num_threads(), t = omp_get_thread_num();
        const int lo = (int)((long long)n * t / nt), hi = (int)((long long)n * (t + 1) / nt);
        int* c = &cnt[(size_t)t * nbc];
        for (int i = lo; i < hi; ++i) ++c[key(i)];
#pragma omp barrier
#pragma omp single
        {
          int run = 0;
          for (int k = 0; k < nbc; ++k) {
            dl.bc_ptr[k] = run;
            for (int u = 0; u < nt; ++u) { const int v = cnt[(size_t)u * nbc + k]; cnt[(size_t)u * nbc + k] = run; run += v; }
          }
          dl.bc_ptr[nbc] = run;
        }
        for (int i = lo; i < hi; ++i) perm[l][c[key(i)]++] = i;
      }
    }
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) iperm[l][perm[l][i]] = i;
  }
  lap("permutations", -1);
  size_t max_n = 0;
  {
    size_t need = kArenaHead;
    for (int l = 0; l < L; ++l) need += 3 * (size_t)H.lv[l].A.n + 6;
    need += 10 * ((size_t)H.lv[0].A.n + 2) + 2 * (size_t)H.lv[0].A.n + 8;
    D.arena = dalloc<double>(D, need);
    D.arena_doubles = need;
    CUDA_OK(cudaMemset(D.arena, 0, need * sizeof(double)));
  }
  size_t arena_used = kArenaHead;   // 64 arrival flags, then two banks of 64 x 4 all-reduce slots
  auto carve = [&](size_t count) {
    double* p = D.arena + arena_used;
    arena_used += (count + 1) & ~(size_t)1;
    if (arena_used > D.arena_doubles) throw std::runtime_error("vector arena overflow");
    return p;
  };
  for (int l = 0; l < L; ++l) {
    const Level& hl = H.lv[l];
    DLevel& dl = D.lv[l];
    const int n = hl.A.n;
    max_n = std::max(max_n, (size_t)n);
    dl.n = n;
    dl.nnz = hl.A.nnz();
    dl.nc = hl.nc;
    dl.lanes = pick_lanes(n ? (double)dl.nnz / n : 1.0);
    dl.unroll = pick_unroll(n);
    // halo mode: a rank stores the matrix rows of its own blocks only (levels with Schwarz patches keep all
    // rows: the patch setup reads the rows of every patch dof, which straddle the cuts)
    int row_lo = 0, row_hi = n;
    if (D.halo && D.world > 1 && dl.nb > 1 && hl.sw.npatch() == 0) {
      const int per = dl.nb / D.world;
      row_lo = dl.bc_ptr[D.rank * per * dl.ncolors];
      row_hi = dl.bc_ptr[(D.rank + 1) * per * dl.ncolors];
      dl.rows_owned_only = true;
    }
    std::vector<int> ia(n + 1, 0);
    for (int i = 0; i < n; ++i)
      ia[i + 1] = ia[i] + ((i >= row_lo && i < row_hi) ? hl.A.ia[perm[l][i] + 1] - hl.A.ia[perm[l][i]] : 0);
    dl.nnz = ia[n];
    bigvec<int> ja(dl.nnz);       // filled completely (in parallel) below: no value-initialising pass
    bigvec<double> a(dl.nnz);
    std::vector<double> invd(n, 1.0);
    std::vector<double> invl1(H.prm.smoother == MAMG_SMOOTHER_L1DIAG ? n : 0, 1.0);
#pragma omp parallel
    {
      std::vector<std::pair<int, double>> buf;
#pragma omp for schedule(static)
      for (int i = row_lo; i < row_hi; ++i) {
        const int o = perm[l][i];
        const int p0 = hl.A.ia[o], cnt = hl.A.ia[o + 1] - p0;
        buf.resize(cnt);
        for (int k = 0; k < cnt; ++k) buf[k] = {iperm[l][hl.A.ja[p0 + k]], hl.A.a[p0 + k]};
        std::sort(buf.begin(), buf.end(),
                  [](const std::pair<int, double>& u, const std::pair<int, double>& v) { return u.first < v.first; });
        double l1 = 0.0;
        for (int k = 0; k < cnt; ++k) {
          ja[ia[i] + k] = buf[k].first;
          a[ia[i] + k] = buf[k].second;
          l1 += std::fabs(buf[k].second);
          if (buf[k].first == i) invd[i] = 1.0 / buf[k].second;
        }
        if (!invl1.empty()) invl1[i] = 1.0 / l1;
      }
    }
    lap("permute+sort", l);
    dl.ia = upload(D, ia);
    dl.ja = upload(D, ja);
    dl.a = upload(D, a);
    dl.invd = upload(D, invd);
    if (!invl1.empty()) dl.invl1 = upload(D, invl1);
    dl.perm = upload(D, perm[l]);
    dl.iperm = upload(D, iperm[l]);
    dl.x_own = dl.x = carve(n);
    dl.b_own = dl.b = carve(n);
    dl.t = carve(n);
    lap("h2d csr", l);
    if (!hl.gs_skip.empty()) {
      std::vector<uint8_t> sk(n);
#pragma omp parallel for schedule(static)
      for (int i = 0; i < n; ++i) sk[i] = hl.gs_skip[perm[l][i]];
      dl.skip = upload(D, sk);
      dl.color_active.assign(dl.nb * dl.ncolors, 0);
      for (int k = 0; k < dl.nb * dl.ncolors; ++k) {
        int active = 0;
#pragma omp parallel for schedule(static) reduction(+ : active)
        for (int i = dl.bc_ptr[k]; i < dl.bc_ptr[k + 1]; ++i) active += sk[i] == 0;
        dl.color_active[k] = active;
      }
    }
    if (l + 1 < L && hl.P.n > 0) {
      // SA_AMG: P (fine' x coarse') and R = P' (coarse' x fine') in the permuted numberings
      auto permute = [&](const Csr& M, const std::vector<int>& rperm, const std::vector<int>& ciperm, DCsr& out) {
        std::vector<int> ia2(M.n + 1, 0), ja2(M.nnz());
        std::vector<double> a2(M.nnz());
        for (int i = 0; i < M.n; ++i) ia2[i + 1] = ia2[i] + (M.ia[rperm[i] + 1] - M.ia[rperm[i]]);
        for (int i = 0; i < M.n; ++i) {
          const int o = rperm[i];
          std::vector<std::pair<int, double>> buf;
          for (int p = M.ia[o]; p < M.ia[o + 1]; ++p) buf.push_back({ciperm[M.ja[p]], M.a[p]});
          std::sort(buf.begin(), buf.end(),
                    [](const std::pair<int, double>& u, const std::pair<int, double>& v) { return u.first < v.first; });
          for (size_t k = 0; k < buf.size(); ++k) { ja2[ia2[i] + k] = buf[k].first; a2[ia2[i] + k] = buf[k].second; }
        }
        out.n = M.n; out.m = M.m; out.nnz = M.nnz();
        out.lanes = pick_lanes(M.n ? (double)M.nnz() / M.n : 1.0);
        out.ia = upload(D, ia2); out.ja = upload(D, ja2); out.a = upload(D, a2);
      };
      permute(hl.P, perm[l], iperm[l + 1], dl.P);
      permute(hl.R, perm[l + 1], iperm[l], dl.R);
    }
    if (l + 1 < L) {
      // members of every aggregate in ascending (permuted) row order -- the order the restriction sums them in:
      // counted and placed with atomics on all cores, then every (short) member list is sorted
      std::vector<int> agg(n), cptr(hl.nc + 1, 0), cidx;
#pragma omp parallel for schedule(static)
      for (int i = 0; i < n; ++i) {
        const int I = hl.agg[perm[l][i]];
        agg[i] = I >= 0 ? iperm[l + 1][I] : -1;
        if (I >= 0) {
#pragma omp atomic
          ++cptr[agg[i] + 1];
        }
      }
      for (int I = 0; I < hl.nc; ++I) cptr[I + 1] += cptr[I];
      cidx.resize(cptr[hl.nc]);
      std::vector<int> fill(cptr.begin(), cptr.end() - 1);
#pragma omp parallel for schedule(static)
      for (int i = 0; i < n; ++i)
        if (agg[i] >= 0) {
          int slot;
#pragma omp atomic capture
          slot = fill[agg[i]]++;
          cidx[slot] = i;
        }
#pragma omp parallel for schedule(static, 4096)
      for (int I = 0; I < hl.nc; ++I) std::sort(cidx.begin() + cptr[I], cidx.begin() + cptr[I + 1]);
      dl.agg = upload(D, agg);
      dl.cptr = upload(D, cptr);
      dl.cidx = upload(D, cidx);
      if (D.halo && D.world > 1 && dl.nb > 1 && D.lv[l + 1].nb == 1) {
        // first replicated level below a distributed one: the coarse rows whose aggregates sit in my blocks
        const int per = dl.nb / D.world;
        const int lo = dl.bc_ptr[D.rank * per * dl.ncolors], hi = dl.bc_ptr[(D.rank + 1) * per * dl.ncolors];
        std::vector<int> mine;
        for (int i = lo; i < hi; ++i)
          if (agg[i] >= 0) mine.push_back(agg[i]);
        std::sort(mine.begin(), mine.end());
        mine.erase(std::unique(mine.begin(), mine.end()), mine.end());
        dl.n_own_coarse = (int)mine.size();
        dl.d_own_coarse = upload(D, mine);
      }
    }
    lap("agg+misc", l);
    if (hl.sw.npatch() > 0) schwarz_upload(hl, dl.nb, iperm[l], dl.ia, dl.ja, dl.a, ia, ja, a, dl.sw, [&](size_t bytes) {
      void* p = nullptr;
      CUDA_OK(cudaMalloc(&p, bytes ? bytes : 1));
      D.allocs.push_back(p);
      D.dev_bytes += (int64_t)bytes;
      return p;
    });
    lap("schwarz", l);
    if (D.halo && D.world > 1 && dl.nb > 1) build_halo_lists(D, dl, ia, ja);
    if (rows_sell()) {
      int slo = 0, shi = n;
      if (D.halo && D.world > 1 && dl.nb > 1) {   // the row kernels only ever stream this rank's blocks
        const int per = dl.nb / D.world;
        slo = dl.bc_ptr[D.rank * per * dl.ncolors];
        shi = dl.bc_ptr[(D.rank + 1) * per * dl.ncolors];
      }
      build_sell(D, dl, ia, slo, shi);
    }
    lap("halo+sell", l);
  }
  {
    size_t mx = 0;
    for (auto& dl : D.lv)
      if (dl.sw.nb > 1)
        for (int c = 0; c < dl.sw.ncolors; ++c)
          mx = std::max(mx, (size_t)(dl.sw.xoff[(c + 1) * dl.sw.nb] - dl.sw.xoff[c * dl.sw.nb]));
    if (mx) { D.xbuf = carve(2 * mx); D.xcap = mx; }   // two halves: a fast peer may already send the next colour
  }
  D.coarse_inv = upload(D, H.coarse_inv);
  {
    // persistent tail: the longest suffix of levels that are small, undistributed, plain UA levels
    const char* env = getenv("MAMG_TAIL_ROWS");
    const int tail_rows = env ? atoi(env) : 4096;
    const bool smoother_ok = H.prm.smoother != MAMG_SMOOTHER_JACOBI && H.prm.smoother != MAMG_SMOOTHER_L1DIAG;
    int k0 = L;
    while (k0 > 0) {
      const DLevel& dl = D.lv[k0 - 1];
      if (dl.n > tail_rows || dl.nb != 1 || dl.sw.npatch > 0 || dl.P.n > 0) break;
      --k0;
    }
    if (smoother_ok && tail_rows > 0 && L - k0 >= 2 && L - k0 <= kTailMaxLevels) {
      TailArgs& T = D.tail;
      T.nlev = L - k0;
      int off = 0;
      for (int l = k0; l < L; ++l) {
        DLevel& dl = D.lv[l];
        dl.d_color_ptr = upload(D, dl.bc_ptr);
        TailLevel& t = T.lv[l - k0];
        t.n = dl.n; t.nc = dl.nc; t.ncolors = dl.ncolors;
        t.ia = dl.ia; t.ja = dl.ja; t.a = dl.a; t.invd = dl.invd; t.color_ptr = dl.d_color_ptr;
        t.agg = dl.agg; t.cptr = dl.cptr; t.cidx = dl.cidx;
        t.off = off;
        off += dl.n;
      }
      T.total = off;
      T.cycle_type = H.prm.cycle_type;
      T.smoother = H.prm.smoother;
      T.pre = H.prm.presmooth_iter;
      T.post = H.prm.postsmooth_iter;
      T.scaling = H.prm.coarse_scaling == MAMG_ON;
      T.omega = H.prm.relaxation;
      T.coarse_inv = D.coarse_inv;
      D.tail_smem = ((size_t)2 * off + 80) * sizeof(double);
      if (D.tail_smem <= 220 * 1024) {
        CUDA_OK(cudaFuncSetAttribute(tail_cycle_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)D.tail_smem));
        D.tail_k0 = k0;
      }
    }
  }
  // the row kernels stream the sliced-ELL copy: drop the CSR entries wherever nothing else reads them
  // (the tail kernel walks CSR rows; the Schwarz kernels read the row values of their level)
  CUDA_OK(cudaDeviceSynchronize());
  for (int l = 0; l < L; ++l) {
    DLevel& dl = D.lv[l];
    const bool in_tail = D.tail_k0 >= 0 && l >= D.tail_k0;
    if (!dl.use_sell || in_tail || dl.sw.npatch > 0) continue;
    dfree(D, dl.ja, (size_t)dl.nnz);
    dfree(D, dl.a, (size_t)dl.nnz);
    dl.ja = nullptr;
    dl.a = nullptr;
    dl.has_csr = false;
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, D.device);
  D.red_blocks = sms * 8;
  D.partial = dalloc<double>(D, (size_t)D.red_blocks * 4);
  D.ticket = dalloc<unsigned int>(D, 4);
  CUDA_OK(cudaMemset(D.ticket, 0, 4 * sizeof(unsigned int)));
  D.scal = dalloc<double>(D, 32);
  CUDA_OK(cudaMemset(D.scal, 0, 32 * sizeof(double)));
  CUDA_OK(cudaMallocHost(&D.h_scal, 32 * sizeof(double)));
  for (int k = 0; k < 10; ++k) D.w[k] = carve(D.lv[0].n);
  D.push_ticket = dalloc<unsigned int>(D, 4);
  CUDA_OK(cudaMemset(D.push_ticket, 0, 4 * sizeof(unsigned int)));
  D.d_phase = dalloc<long long>(D, 2);
  CUDA_OK(cudaMemset(D.d_phase, 0, 2 * sizeof(long long)));
  D.io_a = dalloc<double>(D, max_n);
  D.io_b = dalloc<double>(D, max_n);
}

// ------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------
#define LANES_SWITCH(lanes, ...)                         \
  switch (lanes) {                                       \
    case 2: { constexpr int LN = 2; __VA_ARGS__; break; }   \
    case 4: { constexpr int LN = 4; __VA_ARGS__; break; }   \
    case 8: { constexpr int LN = 8; __VA_ARGS__; break; }   \
    case 16: { constexpr int LN = 16; __VA_ARGS__; break; } \
    default: { constexpr int LN = 32; __VA_ARGS__; break; } \
  }
// rows per sub-warp (independent row streams in flight per lane)
#define UNROLL_SWITCH(u, ...)                            \
  switch (u) {                                           \
    case 1: { constexpr int UN = 1; __VA_ARGS__; break; }   \
    case 2: { constexpr int UN = 2; __VA_ARGS__; break; }   \
    default: { constexpr int UN = 4; __VA_ARGS__; break; }  \
  }

#define NCCL_OK(call)                                                                      \
  do {                                                                                     \
    ncclResult_t r_ = (call);                                                              \
    if (r_ != ncclSuccess)                                                                 \
      throw std::runtime_error(std::string("NCCL error: ") + ncclGetErrorString(r_) + " at " + \
                               __FILE__ + ":" + std::to_string(__LINE__));                 \
  } while (0)

static void set_l2_window(DeviceState& D, const void* ptr, size_t bytes) {
  // only vectors that fit the set-aside completely: a partial window costs more normal L2 than it
  // saves (measured at 16 M DOFs: 2229 -> 2310 ms per solve with a 128 MB vector in a 64 MB window)
  if (bytes > D.l2_persist_bytes) { ptr = nullptr; bytes = 0; }
  if (D.l2_persist_bytes == 0 || ptr == D.l2_window) return;
  cudaStreamAttrValue v;
  std::memset(&v, 0, sizeof(v));
  v.accessPolicyWindow.base_ptr = const_cast<void*>(ptr);
  v.accessPolicyWindow.num_bytes = bytes;
  v.accessPolicyWindow.hitRatio = bytes ? (float)std::min(1.0, (double)D.l2_persist_bytes / (double)bytes) : 0.f;
  v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  cudaStreamSetAttribute(D.stream, cudaStreamAttributeAccessPolicyWindow, &v);
  D.l2_window = ptr;
}

// ---- row distribution ----------------------------------------------------------------------------
// A level with nb > 1 blocks is row-distributed: rank r executes blocks [b_lo, b_hi); every vector
// is complete on every rank, so after a kernel has updated the owned rows the updated ranges are
// all-gathered (one ncclBroadcast per block inside one group).  With world == 1 a partitioned
// hierarchy simply runs all blocks on the one GPU: same arithmetic, same order, no communication.
static bool is_dist(const DeviceState& D, const DLevel& l) { return l.nb > 1; }
static int blk_lo(const DeviceState& D, const DLevel& l) { return l.nb > 1 ? D.rank * (l.nb / D.world) : 0; }
static int blk_hi(const DeviceState& D, const DLevel& l) { return l.nb > 1 ? (D.rank + 1) * (l.nb / D.world) : 1; }
static int own_lo(const DeviceState& D, const DLevel& l) { return l.bc_ptr[blk_lo(D, l) * l.ncolors]; }
static int own_hi(const DeviceState& D, const DLevel& l) { return l.bc_ptr[blk_hi(D, l) * l.ncolors]; }
static void k_fill(DeviceState& D, int n, double* x, double v);

// all-gather of vector v on a distributed level: colour c of every block (c >= 0) or whole blocks (c < 0)
// ---- peer-memory exchange (CUDA IPC over NVLink) --------------------------------------------------
struct PushRanges { int n; int beg[8]; int len[8]; };

// Copies the listed ranges of the local vector (offset `voff` inside the arena) into every peer's
// arena, then the block that finishes last raises this rank's flag (= phase) in every peer.
__global__ void __launch_bounds__(kBlock)
push_kernel(PushRanges R, long long voff, double* const* __restrict__ peers, int me, int world,
            unsigned int* ticket, long long* phase_ctr) {
  int total = 0;
  for (int k = 0; k < R.n; ++k) total += R.len[k];
  const double* mine = peers[me] + voff;
  for (int i = blockIdx.x * kBlock + threadIdx.x; i < total; i += gridDim.x * kBlock) {
    int k = 0, j = i;
    while (j >= R.len[k]) { j -= R.len[k]; ++k; }
    const double val = mine[R.beg[k] + j];
    for (int q = 0; q < world; ++q)
      if (q != me) peers[q][voff + R.beg[k] + j] = val;
  }
  __threadfence_system();
  __syncthreads();
  // the exchange number lives in device memory (every rank runs the same sequence of exchanges), so
  // that a captured graph of the cycle can be replayed
  __shared__ bool last;
  __shared__ long long phase_s;
  if (threadIdx.x == 0) {
    const unsigned int t = atomicInc(ticket, gridDim.x - 1);
    last = t == gridDim.x - 1;
    if (last) {
      phase_s = *reinterpret_cast<volatile long long*>(phase_ctr) + 1;
      __threadfence_system();
      for (int q = 0; q < world; ++q)
        if (q != me) reinterpret_cast<volatile long long*>(peers[q])[me] = phase_s;
    }
  }
  __syncthreads();
  const long long phase = last ? phase_s : 0;
  // the block that finished last also waits for the peers' flags: when this kernel retires, every
  // peer's ranges of this exchange have landed here (one launch per exchange instead of two)
  if (last && threadIdx.x < world && (int)threadIdx.x != me) {
    const volatile long long* flag = reinterpret_cast<const volatile long long*>(peers[me]) + threadIdx.x;
    const long long t0 = clock64();
    while (*flag < phase) {
      if (clock64() - t0 > g_spin_limit) { printf("mamg: peer %d never reached exchange %lld\n", (int)threadIdx.x, phase); __trap(); }
    }
    __threadfence_system();
  }
  if (last) {   // block-uniform
    __syncthreads();
    if (threadIdx.x == 0) *phase_ctr = phase;
  }
}

static void push_ranges(DeviceState& D, const double* v, const PushRanges& R) {
  int total = 0;
  for (int k = 0; k < R.n; ++k) total += R.len[k];
  const long long voff = v - D.arena;
  const int grid = std::max(1, std::min(D.red_blocks, cdiv(std::max(total, 1), kBlock)));
  KScope ks(D, K_EXCH);
  push_kernel<<<grid, kBlock, 0, D.stream>>>(R, voff, D.d_peer_arena, D.rank, D.world, D.push_ticket, D.d_phase);
  ++D.collectives;
}

// ---- halo mode ---------------------------------------------------------------------------------------
// One exchange = the boundary rows of this rank (index list per neighbour) stored straight into the
// neighbours' vectors, then a flag handshake with THOSE neighbours only: ranks that share no boundary
// on the level neither send nor wait.  Every rank executes the same sequence of exchanges, so the
// exchange number (d_phase) is a global clock; a flag holds the number of the last exchange its owner
// finished sending, and flags only grow.
struct HaloPush { int nn; int peer[8]; int beg[8]; int cnt[8]; };

__global__ void __launch_bounds__(kBlock)
halo_push_kernel(HaloPush P, const int* __restrict__ send, long long voff, double* const* __restrict__ peers, int me,
                 unsigned int* ticket, long long* phase_ctr) {
  const double* mine = peers[me] + voff;
  for (int k = 0; k < P.nn; ++k) {
    double* dst = peers[P.peer[k]] + voff;
    const int* idx = send + P.beg[k];
    for (int t = blockIdx.x * kBlock + threadIdx.x; t < P.cnt[k]; t += gridDim.x * kBlock) {
      const int i = idx[t];
      dst[i] = mine[i];
    }
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  __shared__ long long phase_s;
  if (threadIdx.x == 0) {
    const unsigned int t = atomicInc(ticket, gridDim.x - 1);
    last = t == gridDim.x - 1;
    if (last) {
      phase_s = *reinterpret_cast<volatile long long*>(phase_ctr) + 1;
      __threadfence_system();
      for (int k = 0; k < P.nn; ++k) reinterpret_cast<volatile long long*>(peers[P.peer[k]])[me] = phase_s;
    }
  }
  __syncthreads();
  if (!last) return;
  const long long phase = phase_s;
  if ((int)threadIdx.x < P.nn) {
    const volatile long long* flag = reinterpret_cast<const volatile long long*>(peers[me]) + P.peer[threadIdx.x];
    const long long t0 = clock64();
    while (*flag < phase) {
      if (clock64() - t0 > g_spin_limit) { printf("mamg: rank %d: neighbour %d never reached exchange %lld\n", me, P.peer[threadIdx.x], phase); __trap(); }
    }
    __threadfence_system();
  }
  __syncthreads();
  if (threadIdx.x == 0) *phase_ctr = phase;
}

// listed entries of a vector to EVERY peer, all-rank handshake (the right-hand side of the first
// replicated level: each rank contributes the coarse rows whose aggregates it owns)
__global__ void __launch_bounds__(kBlock)
list_push_all_kernel(int cnt, const int* __restrict__ idx, long long voff, double* const* __restrict__ peers, int me, int world,
                     unsigned int* ticket, long long* phase_ctr) {
  const double* mine = peers[me] + voff;
  for (int t = blockIdx.x * kBlock + threadIdx.x; t < cnt; t += gridDim.x * kBlock) {
    const int i = idx[t];
    const double val = mine[i];
    for (int q = 0; q < world; ++q)
      if (q != me) peers[q][voff + i] = val;
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  __shared__ long long phase_s;
  if (threadIdx.x == 0) {
    const unsigned int t = atomicInc(ticket, gridDim.x - 1);
    last = t == gridDim.x - 1;
    if (last) {
      phase_s = *reinterpret_cast<volatile long long*>(phase_ctr) + 1;
      __threadfence_system();
      for (int q = 0; q < world; ++q)
        if (q != me) reinterpret_cast<volatile long long*>(peers[q])[me] = phase_s;
    }
  }
  __syncthreads();
  if (!last) return;
  const long long phase = phase_s;
  if ((int)threadIdx.x < world && (int)threadIdx.x != me) {
    const volatile long long* flag = reinterpret_cast<const volatile long long*>(peers[me]) + threadIdx.x;
    const long long t0 = clock64();
    while (*flag < phase) {
      if (clock64() - t0 > g_spin_limit) { printf("mamg: rank %d: peer %d never reached exchange %lld\n", me, (int)threadIdx.x, phase); __trap(); }
    }
    __threadfence_system();
  }
  __syncthreads();
  if (threadIdx.x == 0) *phase_ctr = phase;
}

// All-reduce of up to 4 scalars in fixed rank order (identical bits on every rank), then the scalar
// post-processing that the single-GPU reduction kernels do in their finishing block.
//   op 1  coarse scaling: v = sc+8:  v[0] = e.r, v[1] = e.Ae summed, v[2] = min(v[0]/v[1], 1)
//   op 2  CG step length: v = sc+1:  v[0] = d.q summed, sc[2] = sc[0] / sc[1]
//   op 3  CG r.z (v = sc+5: r.z, r.r): sc[3] = r.z; sc[4] = sc[3]/sc[0] unless first; sc[0] = sc[3]
//   op 0  plain sum of `count` values at v
// Slots: two banks (exchange number parity) of 64 ranks x 4 doubles behind the 64 flags of the arena.
__global__ void __launch_bounds__(64)
allreduce_kernel(int count, double* v, int op, int first, double* sc, double* const* __restrict__ peers, int me, int world,
                 long long* phase_ctr) {
  __shared__ long long phase_s;
  if (threadIdx.x == 0) phase_s = *reinterpret_cast<volatile long long*>(phase_ctr) + 1;
  __syncthreads();
  const long long phase = phase_s;
  const int bank = 64 + (int)(phase & 1) * 256;
  const int q = threadIdx.x;
  if (q < world && q != me) {
    for (int k = 0; k < count; ++k) reinterpret_cast<volatile double*>(peers[q])[bank + me * 4 + k] = v[k];
    __threadfence_system();
    reinterpret_cast<volatile long long*>(peers[q])[me] = phase;
    const volatile long long* flag = reinterpret_cast<const volatile long long*>(peers[me]) + q;
    const long long t0 = clock64();
    while (*flag < phase) {
      if (clock64() - t0 > g_spin_limit) { printf("mamg: rank %d: peer %d never reached all-reduce %lld\n", me, q, phase); __trap(); }
    }
    __threadfence_system();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const volatile double* slots = reinterpret_cast<const volatile double*>(peers[me]) + bank;
    for (int k = 0; k < count; ++k) {
      double acc = 0.0;
      for (int r = 0; r < world; ++r) acc += (r == me) ? v[k] : slots[r * 4 + k];
      v[k] = acc;
    }
    if (op == 1) { const double al = v[0] / v[1]; v[2] = (al < 1.0) ? al : 1.0; }
    else if (op == 2) sc[2] = sc[0] / sc[1];
    else if (op == 3) { sc[3] = sc[5]; if (!first) sc[4] = sc[3] / sc[0]; sc[0] = sc[3]; }
    *phase_ctr = phase;
  }
}

// partial sums over rows [lo, hi): out[0] = u.v, out[1] = u.u   (no post-processing; the all-reduce does it)
__global__ void __launch_bounds__(kBlock)
dot2_range_kernel(int lo, int hi, const double* __restrict__ u, const double* __restrict__ v, double* partial,
                  unsigned int* ticket, double* out) {
  double acc[2] = {0.0, 0.0};
  for (int i = lo + blockIdx.x * kBlock + threadIdx.x; i < hi; i += gridDim.x * kBlock) {
    const double ui = u[i];
    acc[0] += ui * v[i];
    acc[1] += ui * ui;
  }
  block_reduce_finish<2>(acc, partial, ticket, out);
}
// x += alpha d, r -= alpha q and d = z + beta d on a row range
__global__ void __launch_bounds__(kBlock)
pcg_update_range_kernel(int lo, int hi, const double* __restrict__ sc, const double* __restrict__ d,
                        const double* __restrict__ q, double* __restrict__ x, double* __restrict__ r) {
  const int i = lo + blockIdx.x * kBlock + threadIdx.x;
  if (i >= hi) return;
  const double al = sc[2];
  x[i] += al * d[i];
  r[i] -= al * q[i];
}
__global__ void __launch_bounds__(kBlock)
pcg_dir_range_kernel(int lo, int hi, const double* __restrict__ sc, const double* __restrict__ z, double* __restrict__ d) {
  const int i = lo + blockIdx.x * kBlock + threadIdx.x;
  if (i >= hi) return;
  d[i] = z[i] + sc[4] * d[i];
}

// descriptor of one halo exchange for a kernel that performs it in its last block (kernels.cuh: halo_tail)
static HaloTail halo_tail_desc(DeviceState& D, const DLevel& l, const double* v, int c, bool schwarz_color = false) {
  if (!(v >= D.arena && v < D.arena + D.arena_doubles)) throw std::runtime_error("halo exchange of a vector outside the peer arena");
  const int nnb = (int)l.nbr_ranks.size();
  HaloTail T;
  std::memset(&T, 0, sizeof(T));
  T.nn = nnb;
  const int cc = schwarz_color ? c : (c >= 0 ? c : (c == -1 ? l.ncolors : l.ncolors + 1));
  const std::vector<int>& off = schwarz_color ? l.sw_send_off : l.send_off;
  long long bytes = 0;
  for (int k = 0; k < nnb; ++k) {
    T.peer[k] = l.nbr_ranks[k];
    T.beg[k] = off[(size_t)cc * nnb + k];
    T.cnt[k] = off[(size_t)cc * nnb + k + 1] - T.beg[k];
    bytes += 8LL * T.cnt[k];
  }
  T.send = schwarz_color ? l.d_sw_send : l.d_send;
  T.voff = v - D.arena;
  T.peers = D.d_peer_arena;
  T.me = D.rank;
  T.ticket = D.push_ticket;
  T.phase_ctr = D.d_phase;
  ++D.collectives;
  D.exch_bytes += bytes;
  return T;
}
// entries one exchange of the level sends (largest neighbour list): the push is folded into the producing
// kernel only when ONE block can send it quickly; long lists (the finest levels) get the full-grid push kernel
static int halo_list_len(const DLevel& l, int c, bool schwarz_color = false) {
  const int nnb = (int)l.nbr_ranks.size();
  const int cc = schwarz_color ? c : (c >= 0 ? c : (c == -1 ? l.ncolors : l.ncolors + 1));
  const std::vector<int>& off = schwarz_color ? l.sw_send_off : l.send_off;
  int total = 0;
  for (int k = 0; k < nnb; ++k) total = std::max(total, off[(size_t)cc * nnb + k + 1] - off[(size_t)cc * nnb + k]);
  return total;
}
static int halo_fuse_max() {   // MAMG_HALO_FUSE_MAX: longest list a producing kernel sends itself (0: always a push kernel)
  static const int v = getenv("MAMG_HALO_FUSE_MAX") ? atoi(getenv("MAMG_HALO_FUSE_MAX")) : 2048;
  return v;
}

static HaloTail no_halo_tail() {
  HaloTail T;
  std::memset(&T, 0, sizeof(T));
  return T;
}

static void halo_exchange(DeviceState& D, const DLevel& l, const double* v, int c, bool schwarz_color = false) {
  // c in [0, ncolors): the rows of that colour; c = -1: all boundary rows; c = -2: right-hand-side rows of
  // foreign patches; schwarz_color: c is a patch colour, the list holds the dofs my patches of it exported
  if (!(v >= D.arena && v < D.arena + D.arena_doubles)) throw std::runtime_error("halo exchange of a vector outside the peer arena");
  const int nnb = (int)l.nbr_ranks.size();
  HaloPush P;
  P.nn = nnb;
  const int cc = schwarz_color ? c : (c >= 0 ? c : (c == -1 ? l.ncolors : l.ncolors + 1));
  const std::vector<int>& off = schwarz_color ? l.sw_send_off : l.send_off;
  int total = 0;
  for (int k = 0; k < nnb; ++k) {
    P.peer[k] = l.nbr_ranks[k];
    P.beg[k] = off[(size_t)cc * nnb + k];
    P.cnt[k] = off[(size_t)cc * nnb + k + 1] - P.beg[k];
    total = std::max(total, P.cnt[k]);
  }
  const int grid = std::max(1, std::min(D.red_blocks, cdiv(std::max(total, 1), kBlock)));
  KScope ks(D, K_EXCH);
  halo_push_kernel<<<grid, kBlock, 0, D.stream>>>(P, schwarz_color ? l.d_sw_send : l.d_send, v - D.arena, D.d_peer_arena, D.rank, D.push_ticket, D.d_phase);
  ++D.collectives;
  D.exch_bytes += 8LL * [&] { long long t = 0; for (int k = 0; k < nnb; ++k) t += P.cnt[k]; return t; }();
}

static void allreduce(DeviceState& D, int count, double* v, int op, int first) {
  KScope ks(D, K_EXCH);
  allreduce_kernel<<<1, 64, 0, D.stream>>>(count, v, op, first, D.scal, D.d_peer_arena, D.rank, D.world, D.d_phase);
  ++D.collectives;
  D.exch_bytes += 8LL * count * (D.world - 1);
}

// Peer stores of exchange k+1 may land while the receiver is still busy with work it queued after
// barrier k.  That is harmless along the colour / restrict / prolong chains (the ranges written
// are never the ranges read or written locally in between), but not after a replicated local
// write or read of the same vector (zero-fill, gather, Krylov vector updates): those call this
// flag-only barrier first, so that no peer starts the next push before everyone is past them.
static void barrier_only(DeviceState& D) {
  if (D.world == 1 || !D.use_p2p) return;
  PushRanges R;
  R.n = 0;
  push_ranges(D, D.arena + 64, R);
}

static bool halo_on(const DeviceState& D, const DLevel& l) { return D.halo && D.world > 1 && l.nb > 1; }

// every rank's own row blocks of v to every peer (the complete vector everywhere): API boundaries, and
// the whole exchange scheme of the round-1 mode
static void allgather_own(DeviceState& D, const DLevel& l, double* v) {
  const int per = l.nb / D.world;
  if (per > 8) throw std::runtime_error("more than 8 row blocks per rank");
  PushRanges R;
  R.n = 0;
  for (int b = D.rank * per; b < (D.rank + 1) * per; ++b) {
    const int r0 = l.bc_ptr[b * l.ncolors], r1 = l.bc_ptr[(b + 1) * l.ncolors];
    if (r1 > r0) { R.beg[R.n] = r0; R.len[R.n] = r1 - r0; ++R.n; }
  }
  push_ranges(D, v, R);
}

static void exchange(DeviceState& D, const DLevel& l, double* v, int c) {
  if (D.world == 1 || !is_dist(D, l)) return;
  if (D.halo) { halo_exchange(D, l, v, c); return; }
  const int per = l.nb / D.world;
  if (D.use_p2p && per <= 8 && v >= D.arena && v < D.arena + D.arena_doubles) {
    PushRanges R;
    R.n = 0;
    for (int b = D.rank * per; b < (D.rank + 1) * per; ++b) {
      const int r0 = c >= 0 ? l.row0(b, c) : l.bc_ptr[b * l.ncolors];
      const int r1 = c >= 0 ? l.row1(b, c) : l.bc_ptr[(b + 1) * l.ncolors];
      if (r1 > r0) { R.beg[R.n] = r0; R.len[R.n] = r1 - r0; ++R.n; }
    }
    push_ranges(D, v, R);
    return;
  }
  if (D.capturing) throw std::runtime_error("exchange of a vector outside the peer arena while capturing the cycle (set MAMG_GRAPH_DIST=0)");
  NCCL_OK(ncclGroupStart());
  for (int b = 0; b < l.nb; ++b) {
    const int r0 = c >= 0 ? l.row0(b, c) : l.bc_ptr[b * l.ncolors];
    const int r1 = c >= 0 ? l.row1(b, c) : l.bc_ptr[(b + 1) * l.ncolors];
    if (r1 > r0) NCCL_OK(ncclBroadcast(v + r0, v + r0, (size_t)(r1 - r0), ncclDouble, b / per, D.comm, D.stream));
  }
  NCCL_OK(ncclGroupEnd());
  ++D.collectives;
}

// CTAs that cover the slices overlapping rows [r0, r1)
static int sell_grid(int r0, int r1) { return cdiv((r1 + 31) / 32 - r0 / 32, kSellWarps); }

// complete = true: the result is made complete on every rank (all-gather of the owned rows); halo-aware
// callers that only need their own rows pass false
static void k_spmv(DeviceState& D, const DLevel& l, const double* x, const double* b, double* y, bool resid,
                   bool complete = true) {
  if (l.n == 0) return;
  const int r0 = own_lo(D, l), r1 = own_hi(D, l);
  if (is_dist(D, l) && complete) barrier_only(D);   // y may still be in local use on a peer (Krylov vector updates)
  if (r1 > r0 && l.use_sell) {
    const int grid = sell_grid(r0, r1);
    KScope ks(D, K_SPMV);
    if (resid) sell_spmv_kernel<true><<<grid, kBlock, 0, D.stream>>>(r0, r1, l.S, x, b, y);
    else sell_spmv_kernel<false><<<grid, kBlock, 0, D.stream>>>(r0, r1, l.S, x, b, y);
  } else if (r1 > r0) {
    const int grid = cdiv((long long)cdiv(r1 - r0, l.unroll) * l.lanes, kBlock);
    KScope ks(D, K_SPMV);
    LANES_SWITCH(l.lanes, UNROLL_SWITCH(l.unroll,
      if (resid) spmv_kernel<LN, UN, true><<<grid, kBlock, 0, D.stream>>>(r0, r1, l.ia, l.ja, l.a, x, b, y);
      else spmv_kernel<LN, UN, false><<<grid, kBlock, 0, D.stream>>>(r0, r1, l.ia, l.ja, l.a, x, b, y)));
  }
  if (D.world > 1 && is_dist(D, l) && complete) {
    if (D.halo) allgather_own(D, l, y); else exchange(D, l, y, -1);
  }
}

static void k_gs_color(DeviceState& D, const DLevel& l, int c, const double* b, double* x, double omega) {
  bool any = false;
  for (int blk = 0; blk < l.nb; ++blk)
    any |= l.row1(blk, c) > l.row0(blk, c) && (l.color_active.empty() || l.color_active[blk * l.ncolors + c] > 0);
  if (!any) return;   // every row of the colour belongs to Schwarz (same decision on every rank)
  // halo mode: the launch of this rank's last active block of the colour also sends the boundary rows
  int last_blk = -1;
  if (halo_on(D, l) && l.use_sell && halo_list_len(l, c) <= halo_fuse_max())
    for (int blk = blk_lo(D, l); blk < blk_hi(D, l); ++blk)
      if (l.row1(blk, c) > l.row0(blk, c) && (l.color_active.empty() || l.color_active[blk * l.ncolors + c] > 0)) last_blk = blk;
  for (int blk = blk_lo(D, l); blk < blk_hi(D, l); ++blk) {
    const int r0 = l.row0(blk, c), r1 = l.row1(blk, c);
    if (r1 <= r0) continue;
    if (!l.color_active.empty() && l.color_active[blk * l.ncolors + c] == 0) continue;
    KScope ks(D, K_GS);
    if (l.use_sell) {
      const HaloTail tail = blk == last_blk ? halo_tail_desc(D, l, x, c) : no_halo_tail();
      if (tail.nn > 0) sell_gs_kernel<true><<<sell_grid(r0, r1), kBlock, 0, D.stream>>>(r0, r1, l.S, l.invd, l.skip, b, x, omega, tail);
      else sell_gs_kernel<false><<<sell_grid(r0, r1), kBlock, 0, D.stream>>>(r0, r1, l.S, l.invd, l.skip, b, x, omega, tail);
      continue;
    }
    const int grid = cdiv((long long)cdiv(r1 - r0, l.unroll) * l.lanes, kBlock);
    LANES_SWITCH(l.lanes, UNROLL_SWITCH(l.unroll,
      gs_color_kernel<LN, UN><<<grid, kBlock, 0, D.stream>>>(r0, r1, l.ia, l.ja, l.a, l.invd, l.skip, b, x, omega)));
  }
  if (last_blk < 0) exchange(D, l, x, c);   // not folded into a kernel of this rank: a push kernel of its own
}

static void gs_forward(DeviceState& D, const DLevel& l, const double* b, double* x, double w, int first = 0) {
  for (int c = first; c < l.ncolors; ++c) k_gs_color(D, l, c, b, x, w);
}
static void gs_backward(DeviceState& D, const DLevel& l, const double* b, double* x, double w, int skip_last = 0) {
  for (int c = l.ncolors - 1 - skip_last; c >= 0; --c) k_gs_color(D, l, c, b, x, w);
}

// damped Jacobi (invd = 1 / a_ii, weight w) or l1-Jacobi (SMOOTHER_L1DIAG: invd = 1 / sum_j |a_ij|, w = 1)
static void k_jacobi(DeviceState& D, DLevel& l, const double* b, double* x, double w, const double* invd) {
  const int r0 = own_lo(D, l), r1 = own_hi(D, l);
  if (r1 > r0) {
    const int grid = cdiv((long long)(r1 - r0) * l.lanes, kBlock);
    KScope ks(D, K_GS);
    if (l.use_sell) sell_jacobi_kernel<<<sell_grid(r0, r1), kBlock, 0, D.stream>>>(r0, r1, l.S, invd, l.skip, b, x, l.t, w);
    else LANES_SWITCH(l.lanes,
      jacobi_kernel<LN><<<grid, kBlock, 0, D.stream>>>(r0, r1, l.ia, l.ja, l.a, invd, l.skip, b, x, l.t, w));
  }
  if (halo_on(D, l)) {   // new iterate on the owned rows, then its boundary rows to the neighbours
    if (r1 > r0) {
      KScope ks(D, K_VEC);
      copy_kernel<<<cdiv(r1 - r0, kBlock), kBlock, 0, D.stream>>>(r1 - r0, l.t + r0, x + r0);
    }
    halo_exchange(D, l, x, -1);
    return;
  }
  exchange(D, l, l.t, -1);
  {
    KScope ks(D, K_VEC);
    copy_kernel<<<cdiv(l.n, kBlock), kBlock, 0, D.stream>>>(l.n, l.t, x);
  }
  if (is_dist(D, l)) barrier_only(D);
}

// One smoothing application S(x, b) on a level.  Pre-smoothing = Schwarz on the interface
// patches (levels < Schwarz_levels) followed by the point smoother on the remaining rows
// (src/utils.py:84); post-smoothing is its adjoint (point smoother first, reverse directions
// for one-directional variants) so that the cycle stays symmetric.
static void smooth(DeviceState& D, int lev, const double* b, double* x, bool post) {
  Nvtx range(D, post ? "post-smooth" : "pre-smooth", lev);
  DLevel& l = D.lv[lev];
  D.cur_level = lev;
  set_l2_window(D, l.n >= (1 << 19) ? x : nullptr, l.n >= (1 << 19) ? sizeof(double) * (size_t)l.n : 0);
  const mamg_params& P = D.prm;
  const int iters = post ? P.postsmooth_iter : P.presmooth_iter;
  auto point = [&]() {
    for (int it = 0; it < iters; ++it) {
      switch (P.smoother) {
        case MAMG_SMOOTHER_JACOBI: k_jacobi(D, l, b, x, P.relaxation, l.invd); break;
        case MAMG_SMOOTHER_L1DIAG: k_jacobi(D, l, b, x, 1.0, l.invl1); break;
        case MAMG_SMOOTHER_GS:
          if (!post) gs_forward(D, l, b, x, 1.0); else gs_backward(D, l, b, x, 1.0);
          break;
        case MAMG_SMOOTHER_SOR:
          if (!post) gs_forward(D, l, b, x, P.relaxation); else gs_backward(D, l, b, x, P.relaxation);
          break;
        case MAMG_SMOOTHER_SGS:
          // the backward sweep restarts at the colour the forward sweep just finished: for w = 1
          // that update is a no-op in exact arithmetic and is skipped (the oracle does the same)
          gs_forward(D, l, b, x, 1.0);
          gs_backward(D, l, b, x, 1.0, 1);
          break;
        case MAMG_SMOOTHER_SSOR:
          gs_forward(D, l, b, x, P.relaxation);
          gs_backward(D, l, b, x, P.relaxation);
          break;
      }
    }
  };
  auto schwarz = [&]() {
    if (l.sw.npatch == 0) return;
    int type = P.Schwarz_type;
    bool fwd, bwd;
    if (type == MAMG_SCHWARZ_SYMMETRIC) { fwd = bwd = true; }
    else if (type == MAMG_SCHWARZ_FORWARD) { fwd = !post; bwd = post; }
    else { fwd = post; bwd = !post; }
    auto sweep = [&](bool backward) {
      const int snb = l.sw.nb, per = snb > 1 ? snb / D.world : 1;
      for (int cc = 0; cc < l.sw.ncolors; ++cc) {
        const int c = backward ? l.sw.ncolors - 1 - cc : cc;
        const int lo = snb > 1 ? D.rank * per : 0, hi = snb > 1 ? (D.rank + 1) * per : 1;
        // halo mode: the dofs this colour's patches updated go straight into the vectors of the neighbours
        // that gather or own them, sent by the last block of this rank's last patch launch of the colour
        const bool xch = D.world > 1 && snb > 1 && l.sw.xoff[(c + 1) * snb] > l.sw.xoff[c * snb];
        int last_blk = -1;
        if (xch && halo_on(D, l) && halo_list_len(l, c, true) <= halo_fuse_max())
          for (int blk = lo; blk < hi; ++blk)
            if (l.sw.cb_ptr[c * snb + blk + 1] > l.sw.cb_ptr[c * snb + blk]) last_blk = blk;
        for (int blk = lo; blk < hi; ++blk) {
          const int p0 = l.sw.cb_ptr[c * snb + blk], p1 = l.sw.cb_ptr[c * snb + blk + 1];
          if (p1 == p0) continue;
          KScope ks(D, K_SCHWARZ);
          schwarz_range_launch(l.sw, p0, p1, l.a, b, x, D.stream, blk == last_blk ? halo_tail_desc(D, l, x, c, true) : no_halo_tail());
        }
        if (D.world > 1 && snb > 1) {
          // the other ranks receive the just-updated dofs that they read in later colours (or that
          // sit in their row block); everything else travels once, at the end of the sweep
          const int qa = l.sw.xoff[c * snb], qb = l.sw.xoff[(c + 1) * snb];
          if (qb == qa) continue;
          if (halo_on(D, l)) {
            if (last_blk < 0) halo_exchange(D, l, x, c, true);   // no patch of mine in this colour: a push kernel of its own
            continue;
          }
          const int mq0 = l.sw.xoff[c * snb + lo], mq1 = l.sw.xoff[c * snb + hi];
          double* xb = D.xbuf + (D.xflip++ & 1) * D.xcap;
          if (mq1 > mq0) {
            KScope ks(D, K_VEC);
            pack_kernel<<<cdiv(mq1 - mq0, kBlock), kBlock, 0, D.stream>>>(mq1 - mq0, l.sw.xidx + mq0, x, xb + (mq0 - qa));
          }
          if (D.use_p2p) {
            PushRanges R;
            R.n = mq1 > mq0 ? 1 : 0;
            R.beg[0] = mq0 - qa;
            R.len[0] = mq1 - mq0;
            push_ranges(D, xb, R);
          } else {
            NCCL_OK(ncclGroupStart());
            for (int blk = 0; blk < snb; ++blk) {
              const int q0 = l.sw.xoff[c * snb + blk], q1 = l.sw.xoff[c * snb + blk + 1];
              if (q1 > q0) NCCL_OK(ncclBroadcast(xb + (q0 - qa), xb + (q0 - qa), (size_t)(q1 - q0), ncclDouble, blk / per, D.comm, D.stream));
            }
            NCCL_OK(ncclGroupEnd());
            ++D.collectives;
          }
          KScope ks(D, K_VEC);
          unpack_kernel<<<cdiv(qb - qa, kBlock), kBlock, 0, D.stream>>>(qb - qa, l.sw.xidx + qa, xb, x);
        }
      }
      if (D.world > 1 && snb > 1) exchange(D, l, x, -1);   // every rank's own row block is complete: all-gather it
    };
    if (fwd) sweep(false);
    if (bwd) sweep(true);
  };
  if (!post) { schwarz(); point(); } else { point(); schwarz(); }
}

static void k_csr_apply(DeviceState& D, const DCsr& M, int r0, int r1, const double* x, const double* alpha, double* y,
                        double* zero, bool add, int cls) {
  if (r1 <= r0) return;
  const int grid = cdiv((long long)(r1 - r0) * M.lanes, kBlock);
  KScope ks(D, cls);
  LANES_SWITCH(M.lanes,
    if (add) csr_apply_kernel<LN, true><<<grid, kBlock, 0, D.stream>>>(r0, r1, M.ia, M.ja, M.a, x, alpha, y, zero);
    else csr_apply_kernel<LN, false><<<grid, kBlock, 0, D.stream>>>(r0, r1, M.ia, M.ja, M.a, x, alpha, y, zero));
}

static void k_resid_restrict(DeviceState& D, int lev) {
  Nvtx range(D, "residual+restrict", lev);
  D.cur_level = lev;
  DLevel& f = D.lv[lev];
  DLevel& c = D.lv[lev + 1];
  // coarse rows this rank computes: its own blocks when the coarse level is distributed too, else all
  // of them (every rank then forms the small coarse right-hand side redundantly, without communication)
  const int c0 = own_lo(D, c), c1 = own_hi(D, c);
  if (f.R.n > 0) {   // SA_AMG: w = b - A x, b_c = R w, x_c = 0
    k_spmv(D, f, f.x, f.b, f.t, true);
    k_csr_apply(D, f.R, c0, c1, f.t, nullptr, c.b, c.x, false, K_RESTRICT);
  } else if (f.use_sell) {
    // t = b - A x on the fine rows this rank owns (aggregates never cross parts, so the members of the
    // coarse rows [c0, c1) are among them), streamed once in layout order; then the aggregate sums
    const int r0 = own_lo(D, f), r1 = own_hi(D, f);
    const bool all = !(is_dist(D, f) && D.world > 1) || (!is_dist(D, c) && !halo_on(D, f));
    const int a0 = all ? 0 : r0, a1 = all ? f.n : r1;
    if (a1 > a0) {
      KScope ks(D, K_RESTRICT);
      sell_spmv_kernel<true><<<sell_grid(a0, a1), kBlock, 0, D.stream>>>(a0, a1, f.S, f.x, f.b, f.t);
    }
    if (halo_on(D, f) && !is_dist(D, c)) {
      // first replicated level: every rank sums the aggregates it owns, then the entries travel to all peers
      k_fill(D, c.n, c.x, 0.0);
      if (f.n_own_coarse > 0) {
        KScope ks(D, K_RESTRICT);
        agg_sum_list_kernel<<<cdiv(f.n_own_coarse, kBlock), kBlock, 0, D.stream>>>(f.n_own_coarse, f.d_own_coarse, f.cptr, f.cidx, f.t, c.b);
      }
      {
        const int grid = std::max(1, std::min(D.red_blocks, cdiv(std::max(f.n_own_coarse, 1), kBlock)));
        KScope ks(D, K_EXCH);
        list_push_all_kernel<<<grid, kBlock, 0, D.stream>>>(f.n_own_coarse, f.d_own_coarse, c.b - D.arena, D.d_peer_arena, D.rank, D.world,
                                                           D.push_ticket, D.d_phase);
        ++D.collectives;
        D.exch_bytes += 8LL * f.n_own_coarse * (D.world - 1);
      }
      return;
    }
    if (c1 > c0) {
      KScope ks(D, K_RESTRICT);
      agg_sum_kernel<<<cdiv(c1 - c0, kBlock), kBlock, 0, D.stream>>>(c0, c1, f.cptr, f.cidx, f.t, c.b, c.x);
    }
  } else if (c1 > c0) {
    const int grid = cdiv((long long)(c1 - c0) * f.lanes, kBlock);
    KScope ks(D, K_RESTRICT);
    LANES_SWITCH(f.lanes,
      resid_restrict_kernel<LN><<<grid, kBlock, 0, D.stream>>>(c0, c1, f.cptr, f.cidx, f.ia, f.ja, f.a, f.x, f.b, c.b, c.x));
  }
  if (is_dist(D, c) && D.world > 1) {
    k_fill(D, c.n, c.x, 0.0);   // before the barrier of the exchange: afterwards peers push into c.x
    if (D.halo) barrier_only(D);   // the coarse right-hand side is only needed on the owned rows
    else exchange(D, c, c.b, -1);
  }
}

static int red_grid(const DeviceState& D, long long threads) {
  return std::max(1, std::min(D.red_blocks, cdiv(threads, kBlock)));
}

static void k_scale_dots(DeviceState& D, int lev) {
  Nvtx range(D, "coarse scaling", lev);
  DLevel& c = D.lv[lev];
  if (halo_on(D, c)) {   // partial e.r and e.A_c e over the owned rows, combined in rank order
    const int r0 = own_lo(D, c), r1 = own_hi(D, c);
    if (!c.use_sell) throw std::runtime_error("halo mode needs the sliced-ELL row kernels (MAMG_ROWS=sell)");
    {
      KScope ks(D, K_SCALE);
      sell_scale_dots_range_kernel<<<red_grid(D, std::max(r1 - r0, 1)), kBlock, 0, D.stream>>>(r0, r1, c.S, c.x, c.b, D.partial, D.ticket, D.scal + 8);
    }
    allreduce(D, 2, D.scal + 8, 1, 0);
    return;
  }
  if (is_dist(D, c) && D.world > 1) {   // t = A_c e on the owned rows, all-gather, then the two dots on complete vectors
    k_spmv(D, c, c.x, nullptr, c.t, false);
    KScope ks(D, K_SCALE);
    scale_dots_vec_kernel<<<red_grid(D, c.n), kBlock, 0, D.stream>>>(c.n, c.x, c.b, c.t, D.partial, D.ticket, D.scal + 8);
    return;
  }
  KScope ks(D, K_SCALE);
  if (c.use_sell) {
    sell_scale_dots_kernel<<<red_grid(D, c.n), kBlock, 0, D.stream>>>(c.S, c.x, c.b, D.partial, D.ticket, D.scal + 8);
    return;
  }
  const int grid = red_grid(D, (long long)c.n * c.lanes);
  LANES_SWITCH(c.lanes,
    scale_dots_kernel<LN><<<grid, kBlock, 0, D.stream>>>(c.n, c.ia, c.ja, c.a, c.x, c.b, D.partial, D.ticket, D.scal + 8));
}

static void k_prolong(DeviceState& D, int lev, bool scaled) {
  Nvtx range(D, "prolong", lev);
  DLevel& f = D.lv[lev];
  DLevel& c = D.lv[lev + 1];
  const int r0 = own_lo(D, f), r1 = own_hi(D, f);
  // a peer may still be reading f.x in its residual/restriction when nothing below was exchanged
  if (is_dist(D, f) && !is_dist(D, c)) barrier_only(D);
  if (f.P.n > 0) {   // SA_AMG: x += alpha P e
    k_csr_apply(D, f.P, r0, r1, c.x, scaled ? D.scal + 10 : nullptr, f.x, nullptr, true, K_PROLONG);
  } else if (r1 > r0) {
    KScope ks(D, K_PROLONG);
    prolong_kernel<<<cdiv(r1 - r0, kBlock), kBlock, 0, D.stream>>>(r0, r1, f.agg, c.x, scaled ? D.scal + 10 : nullptr, f.x);
  }
  exchange(D, f, f.x, -1);
}

static void k_coarse_solve(DeviceState& D) {
  Nvtx range(D, "coarse solve");
  DLevel& c = D.lv.back();
  KScope ks(D, K_COARSE);
  dense_gemv_kernel<<<cdiv((long long)c.n * 32, kBlock), kBlock, 0, D.stream>>>(c.n, D.coarse_inv, c.b, c.x);
}

static void k_fill(DeviceState& D, int n, double* x, double v);
static void k_copy(DeviceState& D, int n, const double* in, double* out);
static void k_axpby(DeviceState& D, int n, double a, const double* x, double b, double* y);
static void k_dot_dev(DeviceState& D, int n, const double* u, const double* v, double* out_dev);
static void k_axpy_dev(DeviceState& D, int n, const double* coef_dev, double scale, const double* x, double* y);

// ---- AMLI, nonlinear AMLI (K-cycle) and additive cycles -----------------------------------------------------
// The remaining values of the reference's cycle_type key (src/amg_parameters.py:6,26,49,69).  They run the
// per-level kernels down to the coarsest level (no persistent tail kernel) on one GPU and are not tuned:
// no reference configuration selects them.

// Coefficients q_0..q_degree of the AMLI polynomial: the polynomial of best uniform approximation to 1/t on
// [lambda_min, lambda_max] = [1/2, 2] (HAZmath / FASP take lambda_max = 2, lambda_min = lambda_max / 4), built by
// the three-term recurrence in the degree.
static void amli_coefficients(int degree, double* coef) {
  const double lmax = 2.0, lmin = 0.5;
  const double mu0 = 1.0 / lmax, mu1 = 1.0 / lmin;
  const double sq = std::sqrt(mu0) + std::sqrt(mu1);
  const double c = sq * sq, a = 4.0 * mu0 * mu1 / c;
  const double sk = std::sqrt(lmax / lmin), delta = (sk - 1.0) / (sk + 1.0), b = delta * delta;
  double q[16][16] = {{0.0}};
  q[0][0] = 0.5 * (mu0 + mu1);
  q[1][0] = 0.5 * c;
  q[1][1] = -mu0 * mu1;
  for (int k = 2; k <= degree; ++k) {
    q[k][0] = a - b * q[k - 2][0] + (1.0 + b) * q[k - 1][0];
    for (int i = 1; i <= k - 2; ++i) q[k][i] = -b * q[k - 2][i] + (1.0 + b) * q[k - 1][i] - a * q[k - 1][i - 1];
    q[k][k - 1] = (1.0 + b) * q[k - 1][k - 1] - a * q[k - 1][k - 2];
    q[k][k] = -a * q[k - 1][k - 1];
  }
  for (int i = 0; i <= degree; ++i) coef[i] = q[degree][i];
}

static bool recursive_poly_cycle(const DeviceState& D) {
  return D.prm.cycle_type == MAMG_AMLI_CYCLE || D.prm.cycle_type == MAMG_NL_AMLI_CYCLE;
}

// called outside stream capture (allocates)
static void ensure_cycle_work(DeviceState& D) {
  if (!recursive_poly_cycle(D)) return;
  if (D.prm.amli_degree < 0 || D.prm.amli_degree > 15) throw std::runtime_error("amli_degree: 0..15");
  amli_coefficients(D.prm.amli_degree, D.amli_coef);
  if (!D.kwork.empty()) return;
  const size_t L = D.lv.size();
  D.kwork.assign(L, nullptr);
  for (size_t l = 1; l < L; ++l) D.kwork[l] = dalloc<double>(D, 3 * (size_t)std::max(D.lv[l].n, 1));
  D.kscal = dalloc<double>(D, 16 * L);
  CUDA_OK(cudaMemset(D.kscal, 0, 16 * L * sizeof(double)));
}

static void coarse_correction_up(DeviceState& D, int lev) {
  const bool scaled = D.prm.coarse_scaling == MAMG_ON;
  if (scaled) k_scale_dots(D, lev + 1);
  k_prolong(D, lev, scaled);
  smooth(D, lev, D.lv[lev].b, D.lv[lev].x, true);
}

// AMLI-cycle (HAZmath / FASP amli() of mgcycle.c): the coarse correction is q(B_c A_c) B_c r_c with the coarse
// cycle B_c, in Horner form: e = B_c r_c; degree times: rhs = A_c e + (q_{degree-i} / q_degree) r_c, e = B_c rhs;
// then e *= q_degree and the usual scaling alpha = min(1, e.r_c / e.A_c e).
static void cycle_amli(DeviceState& D, int lev) {
  const int L = (int)D.lv.size();
  if (lev == L - 1) { k_coarse_solve(D); return; }
  DLevel& f = D.lv[lev];
  DLevel& c = D.lv[lev + 1];
  const int deg = D.prm.amli_degree;
  const double* q = D.amli_coef;
  double* rc = D.kwork[lev + 1];
  smooth(D, lev, f.b, f.x, false);
  k_resid_restrict(D, lev);   // c.b = R (b - A x), c.x = 0
  k_copy(D, c.n, c.b, rc);
  for (int i = 1; i <= deg; ++i) {
    cycle_amli(D, lev + 1);
    k_spmv(D, c, c.x, nullptr, c.b, false);
    k_axpby(D, c.n, q[deg - i] / q[deg], rc, 1.0, c.b);
    k_fill(D, c.n, c.x, 0.0);
  }
  cycle_amli(D, lev + 1);
  k_axpby(D, c.n, 0.0, rc, q[deg], c.x);   // e *= q_degree
  k_copy(D, c.n, rc, c.b);                 // the scaling reads the restricted residual
  coarse_correction_up(D, lev);
}

static void cycle_nlamli(DeviceState& D, int lev);

// K-cycle coarse correction (Notay & Vassilevski; HAZmath / FASP Kcycle_dcsr_pgcg / _pgcr): two steps of a Krylov
// method on A_c x = b_c preconditioned by the nonlinear AMLI cycle of that level.  The launch sequence is fixed
// (graph capture): the second step always runs, and kcycle_step2_kernel discards it on the device when the
// first step already reduced the residual by the K-cycle tolerance 0.2 (where upstream returns early).
static void kcycle(DeviceState& D, int lc) {
  DLevel& c = D.lv[lc];
  const int n = c.n;
  double* bsave = D.kwork[lc];
  double* c1 = bsave + n;
  double* v1 = bsave + 2 * (size_t)n;
  double* s = D.kscal + 16 * (size_t)lc;
  const bool gcg = D.prm.nl_amli_krylov_type == MAMG_SOLVER_GCG;
  k_copy(D, n, c.b, bsave);
  k_dot_dev(D, n, c.b, c.b, s + 0);
  cycle_nlamli(D, lc);                       // c.x (zero on entry) = B r
  k_copy(D, n, c.x, c1);
  k_spmv(D, c, c1, nullptr, v1, false);
  k_dot_dev(D, n, gcg ? c1 : v1, v1, s + 1);
  k_dot_dev(D, n, gcg ? c1 : v1, c.b, s + 2);
  { KScope ks(D, K_VEC); kcycle_step1_kernel<<<1, 32, 0, D.stream>>>(s); }
  k_axpy_dev(D, n, s + 3, -1.0, v1, c.b);    // r -= beta1 v1 (the residual lives in the level's right-hand side)
  k_dot_dev(D, n, c.b, c.b, s + 4);
  k_fill(D, n, c.x, 0.0);
  cycle_nlamli(D, lc);                       // c.x = B r~
  k_spmv(D, c, c.x, nullptr, c.t, false);
  const double* w = gcg ? c.x : c.t;
  k_dot_dev(D, n, w, v1, s + 5);
  k_dot_dev(D, n, w, c.t, s + 6);
  k_dot_dev(D, n, w, c.b, s + 7);
  { KScope ks(D, K_VEC); kcycle_step2_kernel<<<1, 32, 0, D.stream>>>(s, 0.04); }
  if (n > 0) { KScope ks(D, K_VEC); kcycle_combine_kernel<<<cdiv(n, kBlock), kBlock, 0, D.stream>>>(n, s + 8, c1, c.x); }
  k_copy(D, n, bsave, c.b);
}

// nonlinear AMLI-cycle (HAZmath / FASP nl_amli() of mgcycle.c): every coarse problem but the last is handed to
// the K-cycle, the coarsest one is solved directly.
static void cycle_nlamli(DeviceState& D, int lev) {
  const int L = (int)D.lv.size();
  if (lev == L - 1) { k_coarse_solve(D); return; }
  DLevel& f = D.lv[lev];
  smooth(D, lev, f.b, f.x, false);
  k_resid_restrict(D, lev);
  if (lev + 1 == L - 1) k_coarse_solve(D);
  else kcycle(D, lev + 1);
  coarse_correction_up(D, lev);
}

// additive cycle: z = sum_l P_0..P_{l-1} S_l R_{l-1}..R_0 r, S_l = pre- then post-smoothing from a zero iterate,
// exact solve on the coarsest level, no coarse scaling (the frozen choice of oracle/mamg_oracle.c cycle_add)
static void cycle_additive(DeviceState& D) {
  const int L = (int)D.lv.size();
  for (int lev = 0; lev + 1 < L; ++lev) {
    DLevel& f = D.lv[lev];
    k_resid_restrict(D, lev);   // x = 0: restricts the right-hand side itself and zeroes the coarse iterate
    smooth(D, lev, f.b, f.x, false);
    smooth(D, lev, f.b, f.x, true);
  }
  k_coarse_solve(D);
  for (int lev = L - 2; lev >= 0; --lev) k_prolong(D, lev, false);
}

static void cycle_level(DeviceState& D, int lev) {
  const int L = (int)D.lv.size();
  if (lev == D.tail_k0) {   // everything from here down runs inside one kernel
    Nvtx range(D, "tail cycle", lev);
    D.cur_level = lev;
    TailArgs T = D.tail;
    T.top_reps = (lev > 0 && D.prm.cycle_type == MAMG_W_CYCLE) ? 2 : 1;
    T.b_in = D.lv[lev].b;
    T.x_io = D.lv[lev].x;
    T.x_nonzero = lev == 0 ? 1 : 0;
    KScope ks(D, K_COARSE);
    tail_cycle_kernel<<<1, kTailThreads, D.tail_smem, D.stream>>>(T);
    return;
  }
  if (lev == L - 1) { k_coarse_solve(D); return; }
  const int reps = (lev > 0 && D.prm.cycle_type == MAMG_W_CYCLE) ? 2 : 1;
  for (int rep = 0; rep < reps; ++rep) {
    DLevel& l = D.lv[lev];
    smooth(D, lev, l.b, l.x, false);
    k_resid_restrict(D, lev);
    cycle_level(D, lev + 1);
    const bool scaled = D.prm.coarse_scaling == MAMG_ON;
    if (scaled) k_scale_dots(D, lev + 1);
    k_prolong(D, lev, scaled);
    smooth(D, lev, l.b, l.x, true);
  }
}

static void k_fill(DeviceState& D, int n, double* x, double v) {
  if (n == 0) return;
  KScope ks(D, K_VEC);
  fill_kernel<<<cdiv(n, kBlock), kBlock, 0, D.stream>>>(n, x, v);
}
static void k_gather(DeviceState& D, int n, const int* map, const double* in, double* out) {
  if (n == 0) return;
  KScope ks(D, K_VEC);
  gather_kernel<<<cdiv(n, kBlock), kBlock, 0, D.stream>>>(n, map, in, out);
}
// boundary gathers: a flat natural-order vector, or the blocks of a block_vec addressed by offsets
static void k_gather_in(DeviceState& D, int n, const int* perm, const double* flat, const BlockPtrs* blocks, double* out) {
  if (!blocks) { k_gather(D, n, perm, flat, out); return; }
  if (n == 0) return;
  KScope ks(D, K_VEC);
  gather_blocks_kernel<<<cdiv(n, kBlock), kBlock, 0, D.stream>>>(n, perm, *blocks, out);
}
static void k_scatter_out(DeviceState& D, int n, const int* iperm, const double* in, double* flat, const BlockPtrs* blocks) {
  if (!blocks) { k_gather(D, n, iperm, in, flat); return; }
  if (n == 0) return;
  KScope ks(D, K_VEC);
  scatter_blocks_kernel<<<cdiv(n, kBlock), kBlock, 0, D.stream>>>(n, iperm, in, *blocks);
}
static void k_copy(DeviceState& D, int n, const double* in, double* out) {
  if (n == 0) return;
  KScope ks(D, K_VEC);
  copy_kernel<<<cdiv(n, kBlock), kBlock, 0, D.stream>>>(n, in, out);
}

static void drop_graphs(DeviceState& D) {
  for (auto& g : D.graphs) cudaGraphExecDestroy(g.exec);
  D.graphs.clear();
}

static int64_t apply_launch_estimate(const DeviceState& D) {
  // launches of one apply: ~ (4 colours sweeps + 4) per visit, visits doubling per level for W
  double visits = 1, total = 0;
  const bool poly = recursive_poly_cycle(D);
  for (size_t l = 0; l + 1 < D.lv.size(); ++l) {
    if ((int)l == D.tail_k0 && !poly && D.prm.cycle_type != MAMG_ADD_CYCLE) { total += visits; break; }   // one kernel for everything below
    total += visits * (4.0 * D.lv[l].ncolors + 4.0 * D.lv[l].sw.ncolors + (poly ? 24.0 : 4.0));
    if (D.prm.cycle_type == MAMG_W_CYCLE || D.prm.cycle_type == MAMG_NL_AMLI_CYCLE) visits *= 2;
    if (D.prm.cycle_type == MAMG_AMLI_CYCLE) visits *= D.prm.amli_degree + 1;
  }
  return (int64_t)total;
}

static void apply_permuted_raw(DeviceState& D, const double* r, double* z);

// z' = B r' in the permuted ordering of level 0 (both device arrays of size n0)
static void apply_permuted(DeviceState& D, const double* r, double* z) {
  ensure_cycle_work(D);
  // several ranks: only the peer-memory exchange is captured (its exchange counter lives on the device)
  if (!D.use_graph || D.prof_on || (D.world > 1 && !(D.use_p2p && D.graph_dist)) || apply_launch_estimate(D) > 60000) {
    apply_permuted_raw(D, r, z);
    return;
  }
  for (auto& g : D.graphs)
    if (g.r == r && g.z == z) {
      CUDA_OK(cudaGraphLaunch(g.exec, D.stream));
      D.launches += g.launches;
      D.collectives += g.coll;
      D.exch_bytes += g.xbytes;
      for (int k = 0; k < K_NCLS; ++k) D.cls_launches[k] += g.cls[k];
      return;
    }
  // capture (nothing executes), instantiate, then launch through the cache on the next lookup
  const int64_t l0 = D.launches, coll0 = D.collectives, xb0 = D.exch_bytes;
  int64_t c0[K_NCLS];
  for (int k = 0; k < K_NCLS; ++k) c0[k] = D.cls_launches[k];
  cudaGraph_t graph = nullptr;
  CUDA_OK(cudaStreamBeginCapture(D.stream, cudaStreamCaptureModeThreadLocal));
  D.capturing = true;
  try {
    apply_permuted_raw(D, r, z);
  } catch (...) {
    D.capturing = false;
    cudaStreamEndCapture(D.stream, &graph);
    if (graph) cudaGraphDestroy(graph);
    throw;
  }
  D.capturing = false;
  CUDA_OK(cudaStreamEndCapture(D.stream, &graph));
  DeviceState::GraphEntry e;
  e.r = r;
  e.z = z;
  e.launches = D.launches - l0;
  e.coll = D.collectives - coll0;
  e.xbytes = D.exch_bytes - xb0;
  D.collectives = coll0;
  D.exch_bytes = xb0;
  for (int k = 0; k < K_NCLS; ++k) { e.cls[k] = D.cls_launches[k] - c0[k]; D.cls_launches[k] = c0[k]; }
  D.launches = l0;
  CUDA_OK(cudaGraphInstantiate(&e.exec, graph, 0));
  cudaGraphDestroy(graph);
  if (D.graphs.size() >= 4) { cudaGraphExecDestroy(D.graphs.front().exec); D.graphs.erase(D.graphs.begin()); }
  D.graphs.push_back(e);
  apply_permuted(D, r, z);
}

static void apply_permuted_raw(DeviceState& D, const double* r, double* z) {
  Nvtx range(D, "apply");
  DLevel& l0 = D.lv[0];
  l0.b = const_cast<double*>(r);
  l0.x = z;
  if (D.lv.size() == 1) {
    k_coarse_solve(D);
  } else {
    k_fill(D, l0.n, z, 0.0);
    if (is_dist(D, l0)) barrier_only(D);   // peers push into z from the first colour on
    // halo mode: the patches that straddle a cut read the right-hand side on dofs of the neighbour's block
    if (halo_on(D, l0) && l0.sw.npatch > 0) halo_exchange(D, l0, r, -2);
    switch (D.prm.cycle_type) {
      case MAMG_V_CYCLE: case MAMG_W_CYCLE:
        for (int it = 0; it < std::max(1, D.prm.maxit); ++it) cycle_level(D, 0);
        break;
      case MAMG_AMLI_CYCLE: case MAMG_NL_AMLI_CYCLE: case MAMG_ADD_CYCLE:
        if (D.world > 1) throw std::runtime_error("AMLI / NL_AMLI / ADD cycles run on one GPU (V_CYCLE and W_CYCLE are distributed)");
        if (D.prm.cycle_type == MAMG_ADD_CYCLE) { cycle_additive(D); break; }
        for (int it = 0; it < std::max(1, D.prm.maxit); ++it) {
          if (D.prm.cycle_type == MAMG_AMLI_CYCLE) cycle_amli(D, 0); else cycle_nlamli(D, 0);
        }
        break;
      default: throw std::runtime_error("cycle_type: unknown value");
    }
  }
  l0.b = l0.b_own;
  l0.x = l0.x_own;
  set_l2_window(D, nullptr, 0);
}

// z = B r with z complete on every rank (API boundary, MINRES / GMRES): in halo mode the cycle leaves z
// valid on the owned rows and their halo only
static void apply_complete(DeviceState& D, const double* r, double* z) {
  if (halo_on(D, D.lv[0]) && !(r >= D.arena && r < D.arena + D.arena_doubles)) {
    k_copy(D, D.lv[0].n, r, D.w[9]);   // a vector outside the peer arena (GMRES basis) cannot be exchanged
    r = D.w[9];
  }
  apply_permuted(D, r, z);
  if (halo_on(D, D.lv[0])) allgather_own(D, D.lv[0], z);
}

struct IoVec {  // natural-order vector handed over the ABI (host or device memory)
  DeviceState& D;
  bool on_device;
  IoVec(DeviceState& d, bool dev) : D(d), on_device(dev) {}
  const double* in(const double* p, int n, double* stage) {
    if (on_device) return p;
    CUDA_OK(cudaMemcpyAsync(stage, p, sizeof(double) * n, cudaMemcpyHostToDevice, D.stream));
    return stage;
  }
  double* out_ptr(double* p, double* stage) { return on_device ? p : stage; }
  void out(double* p, int n, double* stage) {
    if (on_device) return;
    CUDA_OK(cudaMemcpyAsync(p, stage, sizeof(double) * n, cudaMemcpyDeviceToHost, D.stream));
    CUDA_OK(cudaStreamSynchronize(D.stream));
  }
};

static DeviceState* get_dev(mamg_handle h) {
  if (!h) { set_error("NULL handle"); return nullptr; }
  if (!h->dev) { set_error("hierarchy is not on a device: call mamg_to_device first (there is no CPU fallback)"); return nullptr; }
  cudaSetDevice(h->dev->device);
  return h->dev;
}

static void read_scalars(DeviceState& D, int count) {
  CUDA_OK(cudaMemcpyAsync(D.h_scal, D.scal, sizeof(double) * count, cudaMemcpyDeviceToHost, D.stream));
  CUDA_OK(cudaStreamSynchronize(D.stream));
}

// cbc.block ConjGrad on the device, permuted ordering.
// stop: 0 sqrt(r.Br) <= tol, 1 sqrt(r.Br) <= tol*sqrt(r0.Br0) (cbc.block), 2 ||r||_2 <= tol*||r0||_2
// (HAZmath linear_stop_type 1, src/input_metric.dat:54)
static int pcg_device(DeviceState& D, const double* b_nat, double* x_nat, double tol, int stop,
                      int maxiter, bool use_guess, int* niters, double* residuals, double* alphas,
                      double* betas, const BlockPtrs* b_blocks = nullptr, const BlockPtrs* x_blocks = nullptr) {
  Nvtx range(D, "pcg");
  DLevel& l0 = D.lv[0];
  const int n = l0.n;
  double *b = D.w[0], *x = D.w[1], *r = D.w[2], *z = D.w[3], *d = D.w[4], *q = D.w[5];
  const int vgrid = cdiv(n, kBlock);
  const int rgrid = red_grid(D, n);
  k_gather_in(D, n, l0.perm, b_nat, b_blocks, b);
  if (use_guess) {
    k_gather_in(D, n, l0.perm, x_nat, x_blocks, x);
    k_spmv(D, l0, x, b, r, true);
  } else {
    k_fill(D, n, x, 0.0);
    k_copy(D, n, b, r);
  }
  const bool halo = halo_on(D, l0);
  const int lo = halo ? own_lo(D, l0) : 0, hi = halo ? own_hi(D, l0) : n;
  if (halo && !l0.use_sell) throw std::runtime_error("halo mode needs the sliced-ELL row kernels (MAMG_ROWS=sell)");
  auto rz_dots = [&](int first) {
    KScope ks(D, K_DOT);
    if (!halo) { pcg_rz_kernel<<<rgrid, kBlock, 0, D.stream>>>(n, r, z, D.partial, D.ticket, D.scal, first); return; }
    dot2_range_kernel<<<red_grid(D, std::max(hi - lo, 1)), kBlock, 0, D.stream>>>(lo, hi, r, z, D.partial, D.ticket, D.scal + 5);
  };
  apply_permuted(D, r, z);
  k_copy(D, n, z, d);   // owned rows and halo of z are valid, hence of d
  rz_dots(1);
  if (halo) allreduce(D, 2, D.scal + 5, 3, 1);
  read_scalars(D, 8);
  double rz = D.h_scal[0];
  int it = 0, status = 0;
  double res = stop == 2 ? std::sqrt(D.h_scal[6]) : std::sqrt(rz);
  if (residuals) residuals[0] = res;
  double target = stop != 0 ? tol * res : tol;
  while (res > target && it < maxiter) {
    if (halo) {   // q = A d and the partial d.q on the owned rows (d is current on the halo), combined in rank order
      {
        KScope ks(D, K_SPMV);
        sell_spmv_dot_range_kernel<<<red_grid(D, std::max(hi - lo, 1)), kBlock, 0, D.stream>>>(lo, hi, l0.S, d, q, D.partial, D.ticket, D.scal + 1);
      }
      allreduce(D, 1, D.scal + 1, 2, 0);
    } else if (is_dist(D, l0) && D.world > 1) {   // q = A d on the owned rows, all-gather, then d.q on complete vectors
      k_spmv(D, l0, d, nullptr, q, false);
      KScope ks(D, K_DOT);
      pcg_dq_kernel<<<rgrid, kBlock, 0, D.stream>>>(n, d, q, D.partial, D.ticket, D.scal);
    } else {
      KScope ks(D, K_SPMV);
      if (l0.use_sell) {
        sell_spmv_dot_kernel<<<red_grid(D, n), kBlock, 0, D.stream>>>(l0.S, d, q, D.partial, D.ticket, D.scal);
      } else {
        const int sgrid = red_grid(D, (long long)cdiv(n, l0.unroll) * l0.lanes);
        LANES_SWITCH(l0.lanes, UNROLL_SWITCH(l0.unroll,
          spmv_dot_kernel<LN, UN><<<sgrid, kBlock, 0, D.stream>>>(n, l0.ia, l0.ja, l0.a, d, q, D.partial, D.ticket, D.scal)));
      }
    }
    if (halo) {
      KScope ks(D, K_VEC);
      if (hi > lo) pcg_update_range_kernel<<<cdiv(hi - lo, kBlock), kBlock, 0, D.stream>>>(lo, hi, D.scal, d, q, x, r);
    } else {
      KScope ks(D, K_VEC);
      pcg_update_kernel<<<vgrid, kBlock, 0, D.stream>>>(n, D.scal, d, q, x, r);
    }
    apply_permuted(D, r, z);
    rz_dots(0);
    if (halo) {
      allreduce(D, 2, D.scal + 5, 3, 0);
      {
        KScope ks(D, K_VEC);
        if (hi > lo) pcg_dir_range_kernel<<<cdiv(hi - lo, kBlock), kBlock, 0, D.stream>>>(lo, hi, D.scal, z, d);
      }
      halo_exchange(D, l0, d, -1);   // the next SpMV reads d on the halo
    } else {
      KScope ks(D, K_VEC);
      pcg_dir_kernel<<<vgrid, kBlock, 0, D.stream>>>(n, D.scal, z, d);
    }
    read_scalars(D, 8);
    ++it;
    if (alphas) alphas[it - 1] = D.h_scal[2];
    if (betas) betas[it - 1] = D.h_scal[4];
    rz = D.h_scal[0];
    res = stop == 2 ? std::sqrt(D.h_scal[6]) : std::sqrt(rz);
    if (residuals) residuals[it] = res;
    if (!(rz >= 0.0) || !std::isfinite(D.h_scal[2])) { status = 1; break; }  // "ConjGrad breakdown"
  }
  if (halo) allgather_own(D, l0, x);   // the caller gets the complete solution on every rank
  k_scatter_out(D, n, l0.iperm, x, x_nat, x_blocks);
  *niters = it;
  return status;
}

// ---- host-scalar vector helpers for MINRES / GMRES ----------------------------------------------
static void k_axpby(DeviceState& D, int n, double a, const double* x, double b, double* y) {
  if (n == 0) return;
  KScope ks(D, K_VEC);
  axpby_kernel<<<cdiv(n, kBlock), kBlock, 0, D.stream>>>(n, a, x, b, y);
}
static double k_dot(DeviceState& D, int n, const double* u, const double* v) {
  {
    KScope ks(D, K_DOT);
    dot_kernel<<<red_grid(D, n), kBlock, 0, D.stream>>>(n, u, v, D.partial, D.ticket, D.scal + 16);
  }
  CUDA_OK(cudaMemcpyAsync(D.h_scal + 16, D.scal + 16, sizeof(double), cudaMemcpyDeviceToHost, D.stream));
  CUDA_OK(cudaStreamSynchronize(D.stream));
  return D.h_scal[16];
}

// dot product left on the device (no host round trip); read back later together with others
static void k_dot_dev(DeviceState& D, int n, const double* u, const double* v, double* out_dev) {
  KScope ks(D, K_DOT);
  dot_kernel<<<red_grid(D, n), kBlock, 0, D.stream>>>(n, u, v, D.partial, D.ticket, out_dev);
}
static void k_axpy_dev(DeviceState& D, int n, const double* coef_dev, double scale, const double* x, double* y) {
  if (n == 0) return;
  KScope ks(D, K_VEC);
  axpy_dev_kernel<<<cdiv(n, kBlock), kBlock, 0, D.stream>>>(n, coef_dev, scale, x, y);
}
static void read_dev(DeviceState& D, const double* dev, int count, double* host) {
  CUDA_OK(cudaMemcpyAsync(D.h_scal, dev, sizeof(double) * count, cudaMemcpyDeviceToHost, D.stream));
  CUDA_OK(cudaStreamSynchronize(D.stream));
  for (int k = 0; k < count; ++k) host[k] = D.h_scal[k];
}

// Preconditioned MINRES (Paige-Saunders with an SPD preconditioner; the residual estimate phibar
// is the B-norm of the residual, as in block.iterative.MinRes).  Permuted ordering.
static int minres_device(DeviceState& D, const double* b_nat, double* x_nat, double tol, bool relative,
                         int maxiter, int* niters, double* residuals) {
  DLevel& l0 = D.lv[0];
  const int n = l0.n;
  double *x = D.w[0], *r1 = D.w[1], *r2 = D.w[2], *y = D.w[3], *v = D.w[4], *w = D.w[5], *w1 = D.w[6],
         *w2 = D.w[7], *t = D.w[8];
  k_gather(D, n, l0.perm, b_nat, r1);
  k_copy(D, n, r1, r2);
  k_fill(D, n, x, 0.0);
  k_fill(D, n, w, 0.0);
  k_fill(D, n, w2, 0.0);
  apply_complete(D, r1, y);
  double beta1 = k_dot(D, n, r1, y);
  if (!(beta1 >= 0.0)) return 1;
  beta1 = std::sqrt(beta1);
  residuals[0] = beta1;
  const double target = relative ? tol * beta1 : tol;
  double oldb = 0.0, beta = beta1, dbar = 0.0, epsln = 0.0, phibar = beta1, cs = -1.0, sn = 0.0;
  int it = 0;
  while (phibar > target && it < maxiter) {
    ++it;
    k_axpby(D, n, 1.0 / beta, y, 0.0, v);               // v = y / beta
    k_spmv(D, l0, v, nullptr, y, false);                 // y = A v
    if (it >= 2) k_axpby(D, n, -beta / oldb, r1, 1.0, y);
    // alfa = v.y stays on the device: the update y -= (alfa/beta) r2 reads it there, and the host fetches it
    // together with beta^2 after the preconditioner apply -- one synchronisation per iteration
    k_dot_dev(D, n, v, y, D.scal + 16);
    k_axpy_dev(D, n, D.scal + 16, -1.0 / beta, r2, y);
    std::swap(r1, r2);                                   // r1 = r2
    k_copy(D, n, y, r2);                                 // r2 = y
    apply_complete(D, r2, y);                            // y = B r2
    oldb = beta;
    k_dot_dev(D, n, r2, y, D.scal + 17);
    double ab[2];
    read_dev(D, D.scal + 16, 2, ab);
    const double alfa = ab[0], b2 = ab[1];
    if (!(b2 >= 0.0)) { *niters = it; return 1; }
    beta = std::sqrt(b2);
    const double oldeps = epsln;
    const double delta = cs * dbar + sn * alfa;
    const double gbar = sn * dbar - cs * alfa;
    epsln = sn * beta;
    dbar = -cs * beta;
    const double gamma = std::max(std::sqrt(gbar * gbar + beta * beta), 1e-300);
    cs = gbar / gamma;
    sn = beta / gamma;
    const double phi = cs * phibar;
    phibar = sn * phibar;
    // w1 = w2; w2 = w; w = (v - oldeps*w1 - delta*w2) / gamma
    std::swap(w1, w2);
    std::swap(w2, w);
    {   // direction and iterate in one pass
      KScope ks(D, K_VEC);
      minres_update_kernel<<<cdiv(n, kBlock), kBlock, 0, D.stream>>>(n, 1.0 / gamma, oldeps, delta, phi, v, w1, w2, t, x);
    }
    std::swap(w, t);
    residuals[it] = phibar;
  }
  k_gather(D, n, l0.iperm, x, x_nat);
  *niters = it;
  return 0;
}

// Restarted GMRES(m), right-preconditioned: A B u = b, x = B u.  Modified Gram-Schmidt Arnoldi,
// Givens rotations on the host; the residual history is the 2-norm of b - A x.  Permuted ordering.
static int gmres_device(DeviceState& D, const double* b_nat, double* x_nat, double tol, bool relative,
                        int maxiter, int m, int* niters, double* residuals) {
  DLevel& l0 = D.lv[0];
  const int n = l0.n;
  if (m < 1 || m > 30) m = 30;
  while ((int)D.basis.size() < m + 1) D.basis.push_back(dalloc<double>(D, n));
  double *b = D.w[0], *x = D.w[1], *r = D.w[2], *z = D.w[3], *wv = D.w[4], *u = D.w[5];
  k_gather(D, n, l0.perm, b_nat, b);
  k_fill(D, n, x, 0.0);
  k_copy(D, n, b, r);
  double rn = std::sqrt(k_dot(D, n, r, r));
  residuals[0] = rn;
  const double target = relative ? tol * rn : tol;
  int it = 0;
  std::vector<double> Hm((size_t)(m + 1) * m), cs(m), sn(m), g(m + 1);
  while (rn > target && it < maxiter) {
    k_axpby(D, n, 1.0 / rn, r, 0.0, D.basis[0]);
    std::fill(g.begin(), g.end(), 0.0);
    g[0] = rn;
    int j = 0;
    for (; j < m && it < maxiter && rn > target; ++j) {
      apply_complete(D, D.basis[j], z);                   // z = B v_j
      k_spmv(D, l0, z, nullptr, wv, false);               // w = A z
      // modified Gram-Schmidt with the coefficients left on the device: every projection reads its h_ij there,
      // the host fetches the whole column (and ||w||^2) once per inner iteration
      if (!D.hcoef) D.hcoef = dalloc<double>(D, 32);
      if (j + 2 > 32) throw std::runtime_error("GMRES restart length above 30 is not supported");
      for (int i = 0; i <= j; ++i) {
        k_dot_dev(D, n, wv, D.basis[i], D.hcoef + i);
        k_axpy_dev(D, n, D.hcoef + i, -1.0, D.basis[i], wv);
      }
      k_dot_dev(D, n, wv, wv, D.hcoef + j + 1);
      double col[32];
      read_dev(D, D.hcoef, j + 2, col);
      for (int i = 0; i <= j; ++i) Hm[(size_t)i * m + j] = col[i];
      const double hn = std::sqrt(col[j + 1]);
      Hm[(size_t)(j + 1) * m + j] = hn;
      if (hn > 0.0) k_axpby(D, n, 1.0 / hn, wv, 0.0, D.basis[j + 1]);
      for (int i = 0; i < j; ++i) {
        const double t0 = cs[i] * Hm[(size_t)i * m + j] + sn[i] * Hm[(size_t)(i + 1) * m + j];
        Hm[(size_t)(i + 1) * m + j] = -sn[i] * Hm[(size_t)i * m + j] + cs[i] * Hm[(size_t)(i + 1) * m + j];
        Hm[(size_t)i * m + j] = t0;
      }
      const double den = std::hypot(Hm[(size_t)j * m + j], hn);
      cs[j] = den > 0 ? Hm[(size_t)j * m + j] / den : 1.0;
      sn[j] = den > 0 ? hn / den : 0.0;
      Hm[(size_t)j * m + j] = den;
      g[j + 1] = -sn[j] * g[j];
      g[j] = cs[j] * g[j];
      rn = std::fabs(g[j + 1]);
      ++it;
      residuals[it] = rn;
    }
    // back substitution, u = sum y_i v_i, x += B u
    std::vector<double> yv(j);
    for (int i = j - 1; i >= 0; --i) {
      double sacc = g[i];
      for (int k = i + 1; k < j; ++k) sacc -= Hm[(size_t)i * m + k] * yv[k];
      yv[i] = sacc / Hm[(size_t)i * m + i];
    }
    k_fill(D, n, u, 0.0);
    for (int i = 0; i < j; ++i) k_axpby(D, n, yv[i], D.basis[i], 1.0, u);
    apply_complete(D, u, z);
    k_axpby(D, n, 1.0, z, 1.0, x);
    k_spmv(D, l0, x, b, r, true);                         // true residual for the restart
    rn = std::sqrt(k_dot(D, n, r, r));
    residuals[it] = rn;
  }
  k_gather(D, n, l0.iperm, x, x_nat);
  *niters = it;
  return 0;
}

// ---- static race check of the device layout ------------------------------------------------------------
// compute-sanitizer is not available on the GPU pool this was developed on, and the correctness of the
// coloured smoothers IS a data-race property: two rows (patches) that run in one launch must not couple.
// These kernels verify exactly that on the arrays the kernels stream -- the permuted sliced-ELL / CSR rows
// of every colour block and the patch lists -- and count the violations.
__global__ void __launch_bounds__(kBlock)
check_gs_color_kernel(int r0, int r1, const SellView S, const uint8_t* __restrict__ skip, unsigned long long* bad) {
  const int lane = threadIdx.x % 32;
  const int slice = r0 / 32 + blockIdx.x * kSellWarps + threadIdx.x / 32;
  if (slice * 32 >= r1) return;
  const int row = slice * 32 + lane;
  if (row < r0 || row >= r1 || (skip != nullptr && skip[row])) return;
  const int s0 = S.sp[slice], W = S.sp[slice + 1] - s0;
  const size_t base = (size_t)s0 * 32;
  for (int e = 0; e < W; ++e) {
    const size_t pv = (e >> 1) < (W >> 1) ? base + (size_t)(e >> 1) * 64 + lane * 2 + (e & 1) : base + (size_t)(W >> 1) * 64 + lane;
    const size_t pc = (e >> 2) < (W >> 2) ? base + (size_t)(e >> 2) * 128 + lane * 4 + (e & 3)
                                          : base + (size_t)(W >> 2) * 128 + (size_t)(e - (W & ~3)) * 32 + lane;
    const int j = S.col[pc];
    // a row of the SAME launch (same colour range, smoothed) that this row reads with a nonzero weight
    if (j != row && j >= r0 && j < r1 && S.val[pv] != 0.0 && !(skip != nullptr && skip[j])) atomicAdd(bad, 1ull);
  }
}
// stamp[dof] = patch that writes it in this launch; then every patch checks its reads and writes
__global__ void __launch_bounds__(kBlock)
check_patch_stamp_kernel(int p0, int p1, const SwPatch* __restrict__ pat, const int* __restrict__ pidx, int* stamp,
                         unsigned long long* bad) {
  const int p = p0 + blockIdx.x * kBlock + threadIdx.x;
  if (p >= p1) return;
  for (int q = 0; q < pat[p].s; ++q)
    if (atomicExch(stamp + pidx[pat[p].q0 + q], p) != -1) atomicAdd(bad, 1ull);   // two patches write one dof
}
__global__ void __launch_bounds__(kBlock)
check_patch_reads_kernel(int p0, int p1, const SwPatch* __restrict__ pat, const int* __restrict__ pidx, const int* __restrict__ nbr,
                         int* stamp, unsigned long long* bad, int clear) {
  const int p = p0 + blockIdx.x * kBlock + threadIdx.x;
  if (p >= p1) return;
  if (clear) { for (int q = 0; q < pat[p].s; ++q) stamp[pidx[pat[p].q0 + q]] = -1; return; }
  for (int j = 0; j < pat[p].nn; ++j) {
    const int w = stamp[nbr[pat[p].n0 + j]];
    if (w != -1 && w != p) atomicAdd(bad, 1ull);   // reads a dof another patch of the launch writes
  }
}

}  // namespace mamg

using namespace mamg;

#define MAMG_TRY try {
#define MAMG_CATCH                                                   \
  }                                                                  \
  catch (const std::exception& e) { set_error(e.what()); return -2; } \
  catch (...) { set_error("unknown C++ exception"); return -2; }

extern "C" {

static int to_device_impl(mamg_handle h, int32_t device, void* stream, int32_t rank, int32_t world);

int mamg_to_device(mamg_handle h, int32_t device, void* stream) { return to_device_impl(h, device, stream, 0, 1); }

int mamg_to_device_dist(mamg_handle h, int32_t device, void* stream, int32_t rank, int32_t world) {
  if (world < 1 || rank < 0 || rank >= world) { set_error("to_device_dist: bad rank/world"); return -1; }
  return to_device_impl(h, device, stream, rank, world);
}

static int to_device_impl(mamg_handle h, int32_t device, void* stream, int32_t rank, int32_t world) {
  MAMG_TRY
  if (!h) { set_error("NULL handle"); return -1; }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    set_error(std::string("no CUDA device available (") + cudaGetErrorString(e) + "); this path has no CPU fallback");
    return -4;
  }
  if (device < 0 || device >= ndev) { set_error("device index out of range"); return -1; }
  if (h->H.released) { set_error("to_device: the host matrices were released (mamg_release_host); build a new handle to upload again"); return -1; }
  if (h->dev) { device_state_free(h->dev); h->dev = nullptr; }
  CUDA_OK(cudaSetDevice(device));
  DeviceState* D = new DeviceState();
  D->device = device;
  D->prm = h->H.prm;
  if (world > 1) {
    if (h->H.nparts % world != 0) {
      delete D;
      set_error("to_device_dist: the hierarchy has " + std::to_string(h->H.nparts) + " parts, not a multiple of world size " + std::to_string(world));
      return -1;
    }
    D->rank = rank;
    D->world = world;
    const char* eh = getenv("MAMG_HALO");
    const char* ep = getenv("MAMG_P2P");
    bool sa = false;
    for (const Level& L : h->H.lv) sa |= L.P.n > 0;
    // halo mode needs the peer-memory exchange and the sliced-ELL kernels; smoothed prolongators reach
    // across the parts, so SA hierarchies keep the all-gather scheme
    D->halo = !(eh && atoi(eh) == 0) && !(ep && atoi(ep) == 0) && rows_sell() && !sa;
  }
  { const char* g = getenv("MAMG_GRAPH"); if (g) D->use_graph = atoi(g) != 0; }
  { const char* g = getenv("MAMG_NVTX"); D->nvtx = g && atoi(g) != 0; }
  { const char* g = getenv("MAMG_GRAPH_DIST"); if (g) D->graph_dist = atoi(g) != 0; }
  {
    const char* e2 = getenv("MAMG_L2_PERSIST_MB");   // 0 disables; default: what the device allows
    int maxp = 0;
    cudaDeviceGetAttribute(&maxp, cudaDevAttrMaxPersistingL2CacheSize, device);
    size_t want = e2 ? (size_t)atoi(e2) << 20 : (size_t)64 << 20;
    want = std::min(want, (size_t)maxp);
    // only when the finest level's vector fits: with a larger problem the set-aside takes more L2 away
    // from the patch and matrix streams than the coarser levels win back (measured at 16 M DOFs)
    if ((size_t)h->H.lv[0].A.n * sizeof(double) > want) want = 0;
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) D->l2_persist_bytes = want;
    cudaGetLastError();
  }
  try {
    if (stream) { D->stream = (cudaStream_t)stream; D->own_stream = false; }
    else { CUDA_OK(cudaStreamCreateWithFlags(&D->stream, cudaStreamNonBlocking)); D->own_stream = true; }
    upload_hierarchy(h->H, *D);
    CUDA_OK(cudaDeviceSynchronize());
  } catch (...) { device_state_free(D); throw; }
  h->dev = D;
  return 0;
  MAMG_CATCH
}

int mamg_nccl_unique_id(void* out128) {
  MAMG_TRY
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  if (!out128) { set_error("nccl_unique_id: NULL"); return -1; }
  NCCL_OK(ncclGetUniqueId((ncclUniqueId*)out128));
  return 0;
  MAMG_CATCH
}

int mamg_ipc_handle(mamg_handle h, void* out64) {
  MAMG_TRY
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  DeviceState* D = get_dev(h);
  if (!D || !out64) return -1;
  CUDA_OK(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)out64, D->arena));
  return 0;
  MAMG_CATCH
}

int mamg_dist_peers(mamg_handle h, const void* handles64_per_rank) {
  MAMG_TRY
  DeviceState* D = get_dev(h);
  if (!D || !handles64_per_rank) return -1;
  if (D->world <= 1) return 0;
  D->peer_arena.assign(D->world, nullptr);
  for (int q = 0; q < D->world; ++q) {
    if (q == D->rank) { D->peer_arena[q] = D->arena; continue; }
    cudaIpcMemHandle_t hd;
    std::memcpy(&hd, (const char*)handles64_per_rank + 64 * q, 64);
    void* p = nullptr;
    CUDA_OK(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
    D->peer_arena[q] = (double*)p;
  }
  D->d_peer_arena = dalloc<double*>(*D, D->world);
  CUDA_OK(cudaMemcpy(D->d_peer_arena, D->peer_arena.data(), sizeof(double*) * D->world, cudaMemcpyHostToDevice));
  const char* env = getenv("MAMG_P2P");
  D->use_p2p = !(env && atoi(env) == 0);
  return 0;
  MAMG_CATCH
}

int mamg_dist_init(mamg_handle h, int32_t rank, int32_t world, const void* unique_id128) {
  MAMG_TRY
  DeviceState* D = get_dev(h);
  if (!D) return -1;
  if (world < 1 || rank < 0 || rank >= world) { set_error("dist_init: bad rank/world"); return -1; }
  if (D->world > 1 && (D->world != world || D->rank != rank)) {
    set_error("dist_init: rank/world differ from the ones the hierarchy was uploaded for (mamg_to_device_dist)");
    return -1;
  }
  if (world > 1 && D->world == 1 && !getenv("MAMG_HALO_ALLOW_LATE")) D->halo = false;   // uploaded without a rank: round-1 scheme
  if (world > 64) { set_error("dist_init: at most 64 ranks (the arrival flags occupy the first 64 slots of the vector arena; CUDA IPC peers are single-node anyway)"); return -1; }
  if (h->H.nparts % world != 0) {
    set_error("dist_init: the hierarchy has " + std::to_string(h->H.nparts) + " parts, not a multiple of world size " + std::to_string(world));
    return -1;
  }
  if (world > 1) {
    if (!unique_id128) { set_error("dist_init: NULL unique id"); return -1; }
    ncclUniqueId id;
    std::memcpy(&id, unique_id128, sizeof(id));
    NCCL_OK(ncclCommInitRank(&D->comm, world, id, rank));
  }
  D->rank = rank;
  D->world = world;
  {
    const char* env = getenv("MAMG_PEER_TIMEOUT_S");
    const long long limit = (long long)((env ? atof(env) : 60.0) * 2.0e9);
    CUDA_OK(cudaMemcpyToSymbol(g_spin_limit, &limit, sizeof(limit)));
  }
  drop_graphs(*D);
  return 0;
  MAMG_CATCH
}

int mamg_collective_count(mamg_handle h, int64_t* count, int32_t reset) {
  DeviceState* D = get_dev(h);
  if (!D || !count) return -1;
  *count = D->collectives;
  if (reset) D->collectives = 0;
  return 0;
}

int mamg_exchange_bytes(mamg_handle h, int64_t* bytes, int32_t reset) {
  DeviceState* D = get_dev(h);
  if (!D || !bytes) return -1;
  *bytes = D->exch_bytes;
  if (reset) D->exch_bytes = 0;
  return 0;
}

int mamg_set_stream(mamg_handle h, void* stream) {
  MAMG_TRY
  DeviceState* D = get_dev(h);
  if (!D) return -1;
  CUDA_OK(cudaStreamSynchronize(D->stream));
  drop_graphs(*D);
  if (D->own_stream) { cudaStreamDestroy(D->stream); D->own_stream = false; }
  if (stream) D->stream = (cudaStream_t)stream;
  else { CUDA_OK(cudaStreamCreateWithFlags(&D->stream, cudaStreamNonBlocking)); D->own_stream = true; }
  return 0;
  MAMG_CATCH
}

int mamg_set_cycle(mamg_handle h, int32_t cycle_type) {
  MAMG_TRY
  if (!h) { set_error("NULL handle"); return -1; }
  if (cycle_type < MAMG_V_CYCLE || cycle_type > MAMG_ADD_CYCLE) { set_error("set_cycle: unknown cycle_type"); return -1; }
  if (cycle_type == MAMG_ADD_CYCLE && h->H.prm.maxit > 1) { set_error("set_cycle: ADD_CYCLE is applied once per call (maxit 1)"); return -1; }
  if (cycle_type == MAMG_AMLI_CYCLE && (h->H.prm.amli_degree < 0 || h->H.prm.amli_degree > 15)) { set_error("set_cycle: amli_degree 0..15"); return -1; }
  if (cycle_type > MAMG_W_CYCLE && h->dev && h->dev->world > 1) { set_error("set_cycle: AMLI / NL_AMLI / ADD cycles run on one GPU"); return -1; }
  h->H.prm.cycle_type = cycle_type;
  if (h->dev) {
    cudaSetDevice(h->dev->device);
    CUDA_OK(cudaStreamSynchronize(h->dev->stream));
    drop_graphs(*h->dev);
    h->dev->prm.cycle_type = cycle_type;
    h->dev->tail.cycle_type = cycle_type;
  }
  return 0;
  MAMG_CATCH
}

int mamg_device_bytes(mamg_handle h, int64_t* bytes) {
  DeviceState* D = get_dev(h);
  if (!D || !bytes) return -1;
  *bytes = D->dev_bytes;
  return 0;
}

int mamg_sync(mamg_handle h) {
  MAMG_TRY
  DeviceState* D = get_dev(h);
  if (!D) return -1;
  CUDA_OK(cudaStreamSynchronize(D->stream));
  return 0;
  MAMG_CATCH
}

int mamg_apply(mamg_handle h, const double* r, double* z, int32_t on_device) {
  MAMG_TRY
  DeviceState* D = get_dev(h);
  if (!D) return -1;
  if (!r || !z) { set_error("apply: NULL vector"); return -1; }
  DLevel& l0 = D->lv[0];
  IoVec io(*D, on_device != 0);
  const double* rin = io.in(r, l0.n, D->io_a);
  double* zout = io.out_ptr(z, D->io_b);
  k_gather(*D, l0.n, l0.perm, rin, D->w[2]);
  apply_complete(*D, D->w[2], D->w[3]);
  k_gather(*D, l0.n, l0.iperm, D->w[3], zout);
  io.out(z, l0.n, D->io_b);
  CUDA_OK(cudaGetLastError());
  return 0;
  MAMG_CATCH
}

// block_vec entry points: sizes[] are the block lengths (sum = rows of level 0, at most 8 blocks)
static bool make_blocks(DeviceState& D, int32_t nblocks, const int32_t* sizes, const double* const* ptrs, BlockPtrs& B) {
  if (nblocks < 1 || nblocks > 8 || !sizes || !ptrs) { set_error("blocks: between 1 and 8 blocks"); return false; }
  B.nb = nblocks;
  long long off = 0;
  for (int q = 0; q < nblocks; ++q) {
    if (sizes[q] < 0 || !ptrs[q]) { set_error("blocks: negative size or NULL block"); return false; }
    B.off[q] = (int)off;
    B.p[q] = const_cast<double*>(ptrs[q]);
    off += sizes[q];
  }
  for (int q = nblocks; q < 9; ++q) B.off[q] = (int)off;
  for (int q = nblocks; q < 8; ++q) B.p[q] = nullptr;
  if (off != D.lv[0].n) { set_error("blocks: the block sizes do not add up to the matrix size"); return false; }
  return true;
}

int mamg_apply_blocks(mamg_handle h, int32_t nblocks, const int32_t* sizes, const double* const* r_blocks,
                      double* const* z_blocks, int32_t on_device) {
  MAMG_TRY
  DeviceState* D = get_dev(h);
  if (!D) return -1;
  BlockPtrs R, Z;
  if (!make_blocks(*D, nblocks, sizes, r_blocks, R) || !make_blocks(*D, nblocks, sizes, z_blocks, Z)) return -1;
  DLevel& l0 = D->lv[0];
  if (!on_device) {   // host blocks: stage them side by side (the concatenation happens inside the H2D copies)
    for (int q = 0; q < nblocks; ++q)
      CUDA_OK(cudaMemcpyAsync(D->io_a + R.off[q], r_blocks[q], sizeof(double) * sizes[q], cudaMemcpyHostToDevice, D->stream));
    k_gather(*D, l0.n, l0.perm, D->io_a, D->w[2]);
  } else {
    k_gather_in(*D, l0.n, l0.perm, nullptr, &R, D->w[2]);
  }
  apply_complete(*D, D->w[2], D->w[3]);
  if (!on_device) {
    k_gather(*D, l0.n, l0.iperm, D->w[3], D->io_b);
    for (int q = 0; q < nblocks; ++q)
      CUDA_OK(cudaMemcpyAsync(z_blocks[q], D->io_b + Z.off[q], sizeof(double) * sizes[q], cudaMemcpyDeviceToHost, D->stream));
    CUDA_OK(cudaStreamSynchronize(D->stream));
  } else {
    k_scatter_out(*D, l0.n, l0.iperm, D->w[3], nullptr, &Z);
  }
  CUDA_OK(cudaGetLastError());
  return 0;
  MAMG_CATCH
}

int mamg_pcg_blocks(mamg_handle h, int32_t nblocks, const int32_t* sizes, const double* const* b_blocks,
                    double* const* x_blocks, double tolerance, int32_t relative, int32_t maxiter,
                    int32_t use_initial_guess, int32_t on_device, int32_t* niters, double* residuals,
                    double* alphas, double* betas) {
  MAMG_TRY
  DeviceState* D = get_dev(h);
  if (!D) return -1;
  if (!niters) { set_error("pcg_blocks: NULL argument"); return -1; }
  BlockPtrs Bb, Xb;
  if (!make_blocks(*D, nblocks, sizes, b_blocks, Bb) || !make_blocks(*D, nblocks, sizes, x_blocks, Xb)) return -1;
  DLevel& l0 = D->lv[0];
  int it = 0, st = 0;
  if (!on_device) {
    for (int q = 0; q < nblocks; ++q) {
      CUDA_OK(cudaMemcpyAsync(D->io_a + Bb.off[q], b_blocks[q], sizeof(double) * sizes[q], cudaMemcpyHostToDevice, D->stream));
      if (use_initial_guess)
        CUDA_OK(cudaMemcpyAsync(D->io_b + Xb.off[q], x_blocks[q], sizeof(double) * sizes[q], cudaMemcpyHostToDevice, D->stream));
    }
    st = pcg_device(*D, D->io_a, D->io_b, tolerance, relative, maxiter, use_initial_guess != 0, &it, residuals, alphas, betas);
    for (int q = 0; q < nblocks; ++q)
      CUDA_OK(cudaMemcpyAsync(x_blocks[q], D->io_b + Xb.off[q], sizeof(double) * sizes[q], cudaMemcpyDeviceToHost, D->stream));
  } else {
    st = pcg_device(*D, nullptr, nullptr, tolerance, relative, maxiter, use_initial_guess != 0, &it, residuals, alphas, betas, &Bb, &Xb);
  }
  CUDA_OK(cudaStreamSynchronize(D->stream));
  CUDA_OK(cudaGetLastError());
  *niters = it;
  (void)l0;
  if (st) { set_error("ConjGrad breakdown (r.Br < 0 or d.Ad == 0)"); return 1; }
  return 0;
  MAMG_CATCH
}

int mamg_spmv(mamg_handle h, int32_t level, const double* x, double* y, int32_t on_device) {
  MAMG_TRY
  DeviceState* D = get_dev(h);
  if (!D) return -1;
  if (level < 0 || level >= (int)D->lv.size()) { set_error("level out of range"); return -1; }
  DLevel& l = D->lv[level];
  IoVec io(*D, on_device != 0);
  const double* xin = io.in(x, l.n, D->io_a);
  double* yout = io.out_ptr(y, D->io_b);
  k_gather(*D, l.n, l.perm, xin, l.x_own);
  k_spmv(*D, l, l.x_own, nullptr, l.t, false);
  k_gather(*D, l.n, l.iperm, l.t, yout);
  io.out(y, l.n, D->io_b);
  CUDA_OK(cudaGetLastError());
  return 0;
  MAMG_CATCH
}

int mamg_smooth(mamg_handle h, int32_t level, const double* b, double* x, int32_t post, int32_t on_device) {
  MAMG_TRY
  DeviceState* D = get_dev(h);
  if (!D) return -1;
  if (level < 0 || level >= (int)D->lv.size() - 1) { set_error("smooth: level out of range (the coarsest level is solved directly)"); return -1; }
  DLevel& l = D->lv[level];
  IoVec io(*D, on_device != 0);
  const double* bin = io.in(b, l.n, D->io_a);
  const double* xin = io.in(x, l.n, D->io_b);
  k_gather(*D, l.n, l.perm, bin, l.b_own);
  k_gather(*D, l.n, l.perm, xin, l.x_own);
  if (is_dist(*D, l)) barrier_only(*D);
  smooth(*D, level, l.b_own, l.x_own, post != 0);
  if (halo_on(*D, l)) allgather_own(*D, l, l.x_own);
  double* xout = io.out_ptr(x, D->io_b);
  k_gather(*D, l.n, l.iperm, l.x_own, xout);
  io.out(x, l.n, D->io_b);
  CUDA_OK(cudaGetLastError());
  return 0;
  MAMG_CATCH
}

int mamg_pcg(mamg_handle h, const double* b, double* x, double tolerance, int32_t relative,
             int32_t maxiter, int32_t use_initial_guess, int32_t on_device, int32_t* niters,
             double* residuals, double* alphas, double* betas) {
  MAMG_TRY
  DeviceState* D = get_dev(h);
  if (!D) return -1;
  if (!b || !x || !niters) { set_error("pcg: NULL argument"); return -1; }
  DLevel& l0 = D->lv[0];
  IoVec io(*D, on_device != 0);
  const double* bin = io.in(b, l0.n, D->io_a);
  double* xio = io.out_ptr(x, D->io_b);
  if (use_initial_guess && !on_device)
    CUDA_OK(cudaMemcpyAsync(D->io_b, x, sizeof(double) * l0.n, cudaMemcpyHostToDevice, D->stream));
  int it = 0;
  int st = pcg_device(*D, bin, xio, tolerance, relative, maxiter, use_initial_guess != 0, &it,
                      residuals, alphas, betas);
  io.out(x, l0.n, D->io_b);
  if (on_device) CUDA_OK(cudaStreamSynchronize(D->stream));
  CUDA_OK(cudaGetLastError());
  *niters = it;
  if (st) { set_error("ConjGrad breakdown (r.Br < 0 or d.Ad == 0)"); return 1; }
  return 0;
  MAMG_CATCH
}

int mamg_minres(mamg_handle h, const double* b, double* x, double tolerance, int32_t relative,
                int32_t maxiter, int32_t on_device, int32_t* niters, double* residuals) {
  MAMG_TRY
  DeviceState* D = get_dev(h);
  if (!D) return -1;
  if (!b || !x || !niters || !residuals) { set_error("minres: NULL argument"); return -1; }
  DLevel& l0 = D->lv[0];
  IoVec io(*D, on_device != 0);
  const double* bin = io.in(b, l0.n, D->io_a);
  double* xio = io.out_ptr(x, D->io_b);
  int it = 0;
  int st = minres_device(*D, bin, xio, tolerance, relative != 0, maxiter, &it, residuals);
  io.out(x, l0.n, D->io_b);
  if (on_device) CUDA_OK(cudaStreamSynchronize(D->stream));
  CUDA_OK(cudaGetLastError());
  *niters = it;
  if (st) { set_error("MinRes breakdown (r.Br < 0: preconditioner not positive definite)"); return 1; }
  return 0;
  MAMG_CATCH
}

int mamg_gmres(mamg_handle h, const double* b, double* x, double tolerance, int32_t relative,
               int32_t maxiter, int32_t restart, int32_t on_device, int32_t* niters, double* residuals) {
  MAMG_TRY
  DeviceState* D = get_dev(h);
  if (!D) return -1;
  if (!b || !x || !niters || !residuals) { set_error("gmres: NULL argument"); return -1; }
  DLevel& l0 = D->lv[0];
  IoVec io(*D, on_device != 0);
  const double* bin = io.in(b, l0.n, D->io_a);
  double* xio = io.out_ptr(x, D->io_b);
  int it = 0;
  gmres_device(*D, bin, xio, tolerance, relative != 0, maxiter, restart, &it, residuals);
  io.out(x, l0.n, D->io_b);
  if (on_device) CUDA_OK(cudaStreamSynchronize(D->stream));
  CUDA_OK(cudaGetLastError());
  *niters = it;
  return 0;
  MAMG_CATCH
}

int mamg_profile(mamg_handle h, int32_t on, double* ms_per_class, int64_t* launches_per_class) {
  MAMG_TRY
  DeviceState* D = get_dev(h);
  if (!D) return -1;
  CUDA_OK(cudaStreamSynchronize(D->stream));
  if (on) {
    D->prof_on = true;
    D->prof_used = 0;
    for (int k = 0; k < K_NCLS; ++k) D->cls_launches[k] = 0;
    return 0;
  }
  D->prof_on = false;
  if (ms_per_class) for (int k = 0; k < K_NCLS; ++k) ms_per_class[k] = 0.0;
  for (size_t i = 0; i < D->prof_used; ++i) {
    float ms = 0.f;
    CUDA_OK(cudaEventElapsedTime(&ms, D->prof_events[i].start, D->prof_events[i].stop));
    if (ms_per_class) ms_per_class[D->prof_events[i].cls] += ms;
  }
  if (launches_per_class) for (int k = 0; k < K_NCLS; ++k) launches_per_class[k] = D->cls_launches[k];
  D->prof_used = 0;
  return 0;
  MAMG_CATCH
}

int mamg_schwarz_sweep_bytes(mamg_handle h, int32_t level, int64_t* bytes) {
  DeviceState* D = get_dev(h);
  if (!D || !bytes) return -1;
  if (level < 0 || level >= (int)D->lv.size()) { set_error("level out of range"); return -1; }
  *bytes = D->lv[level].sw.alg_bytes;
  return 0;
}

int mamg_race_check(mamg_handle h, int64_t* gs_conflicts, int64_t* patch_conflicts, int64_t* launches_checked) {
  MAMG_TRY
  DeviceState* D = get_dev(h);
  if (!D || !gs_conflicts || !patch_conflicts) return -1;
  CUDA_OK(cudaStreamSynchronize(D->stream));
  unsigned long long* bad = nullptr;
  CUDA_OK(cudaMalloc(&bad, 2 * sizeof(unsigned long long)));
  CUDA_OK(cudaMemset(bad, 0, 2 * sizeof(unsigned long long)));
  int64_t checked = 0;
  for (DLevel& l : D->lv) {
    if (l.use_sell && l.ncolors > 0 && &l != &D->lv.back()) {
      for (int blk = 0; blk < l.nb; ++blk) {
        if (l.rows_owned_only && (blk < blk_lo(*D, l) || blk >= blk_hi(*D, l))) continue;
        for (int c = 0; c < l.ncolors; ++c) {
          const int r0 = l.row0(blk, c), r1 = l.row1(blk, c);
          if (r1 <= r0) continue;
          if (!l.color_active.empty() && l.color_active[blk * l.ncolors + c] == 0) continue;
          check_gs_color_kernel<<<sell_grid(r0, r1), kBlock, 0, D->stream>>>(r0, r1, l.S, l.skip, bad);
          ++checked;
        }
      }
    }
    if (l.sw.npatch > 0) {
      int* stamp = nullptr;
      CUDA_OK(cudaMalloc(&stamp, sizeof(int) * (size_t)l.n));
      CUDA_OK(cudaMemsetAsync(stamp, 0xff, sizeof(int) * (size_t)l.n, D->stream));
      for (int kb = 0; kb + 1 < (int)l.sw.cb_ptr.size(); ++kb) {
        // all blocks of one colour run concurrently on different ranks: check a colour as a whole
        if (kb % l.sw.nb != 0) continue;
        const int p0 = l.sw.cb_ptr[kb], p1 = l.sw.cb_ptr[kb + l.sw.nb];
        if (p1 <= p0) continue;
        const int g = cdiv(p1 - p0, kBlock);
        check_patch_stamp_kernel<<<g, kBlock, 0, D->stream>>>(p0, p1, l.sw.pat, l.sw.pidx, stamp, bad + 1);
        check_patch_reads_kernel<<<g, kBlock, 0, D->stream>>>(p0, p1, l.sw.pat, l.sw.pidx, l.sw.nbr, stamp, bad + 1, 0);
        check_patch_reads_kernel<<<g, kBlock, 0, D->stream>>>(p0, p1, l.sw.pat, l.sw.pidx, l.sw.nbr, stamp, bad + 1, 1);
        ++checked;
      }
      CUDA_OK(cudaStreamSynchronize(D->stream));
      cudaFree(stamp);
    }
  }
  unsigned long long hb[2] = {0, 0};
  CUDA_OK(cudaMemcpyAsync(hb, bad, sizeof(hb), cudaMemcpyDeviceToHost, D->stream));
  CUDA_OK(cudaStreamSynchronize(D->stream));
  cudaFree(bad);
  *gs_conflicts = (int64_t)hb[0];
  *patch_conflicts = (int64_t)hb[1];
  if (launches_checked) *launches_checked = checked;
  return 0;
  MAMG_CATCH
}

int mamg_stats(mamg_handle h, int32_t level, int64_t out[24]) {
  DeviceState* D = get_dev(h);
  if (!D || !out) return -1;
  if (level < 0 || level >= (int)D->lv.size()) { set_error("level out of range"); return -1; }
  const DLevel& l = D->lv[level];
  for (int k = 0; k < 24; ++k) out[k] = 0;
  out[0] = l.n;
  out[1] = l.nnz;
  out[2] = h->H.lv[level].nnz_structural;
  out[3] = l.sell_slots;
  out[4] = D->dev_bytes;
  out[5] = l.sw.npatch;
  out[6] = l.sw.nuniq;
  out[7] = l.sw.alg_bytes;
  out[8] = l.sw.alg_bytes_stored;
  out[9] = l.ncolors;
  out[10] = l.sw.ncolors;
  out[11] = (D->tail_k0 >= 0 && level >= D->tail_k0) ? 1 : 0;
  out[12] = l.use_sell ? 1 : 0;
  out[13] = l.has_csr ? 1 : 0;
  out[14] = l.nb;
  out[15] = l.sw.fast ? 1 : 0;
  out[16] = l.sw.max_size;
  out[17] = l.sw.max_nbr;
  return 0;
}

int mamg_profile_levels(mamg_handle h, double* ms_level_class, int32_t max_levels) {
  MAMG_TRY
  DeviceState* D = get_dev(h);
  if (!D || !ms_level_class) return -1;
  CUDA_OK(cudaStreamSynchronize(D->stream));
  for (int k = 0; k < max_levels * 16; ++k) ms_level_class[k] = 0.0;
  for (size_t i = 0; i < D->prof_used; ++i) {
    float ms = 0.f;
    CUDA_OK(cudaEventElapsedTime(&ms, D->prof_events[i].start, D->prof_events[i].stop));
    const int lev = std::min(D->prof_events[i].lev, max_levels - 1);
    ms_level_class[lev * 16 + D->prof_events[i].cls] += ms;
  }
  return 0;
  MAMG_CATCH
}

int mamg_launch_count(mamg_handle h, int64_t* launches, int32_t reset) {
  DeviceState* D = get_dev(h);
  if (!D || !launches) return -1;
  *launches = D->launches;
  if (reset) D->launches = 0;
  return 0;
}

// Algorithmic bytes of ONE cycle by the model of SURVEY 8(d) / BASELINE.md 4, with the level
// sizes of this hierarchy: fp64 values, int32 columns, vectors counted once per kernel.
int mamg_cycle_bytes(mamg_handle h, int64_t* bytes) {
  if (!h || !bytes) { set_error("cycle_bytes: NULL"); return -1; }
  const Hierarchy& H = h->H;
  const int L = (int)H.lv.size();
  double total = 0;
  double visits = 1;
  for (int l = 0; l < L - 1; ++l) {
    const double n = H.lv[l].A.n, nnz = H.lv[l].A.nnz(), nc = H.lv[l + 1].A.n, nnzc = H.lv[l + 1].A.nnz();
    const double b_gs = 12 * nnz + 4 * (n + 1) + 24 * n;
    const double b_spmv = 12 * nnz + 4 * (n + 1) + 16 * n;
    const double b_spmv_c = 12 * nnzc + 4 * (nc + 1) + 16 * nc;
    double sweeps = 0;
    switch (H.prm.smoother) {
      case MAMG_SMOOTHER_SGS: case MAMG_SMOOTHER_SSOR: sweeps = 2.0 * (H.prm.presmooth_iter + H.prm.postsmooth_iter); break;
      default: sweeps = 1.0 * (H.prm.presmooth_iter + H.prm.postsmooth_iter);
    }
    double visit = sweeps * b_gs + (b_spmv + 8 * n) + (32 * n + 16 * nc);
    if (H.prm.coarse_scaling == MAMG_ON) visit += b_spmv_c + 16 * nc;
    total += visits * visit;
    if (H.prm.cycle_type == MAMG_W_CYCLE) visits *= 2;
  }
  *bytes = (int64_t)total;
  return 0;
}

}  // extern "C"

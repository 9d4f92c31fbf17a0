/* mamg.h -- C-ABI of the B200-native metric-AMG apply path.
 *
 * Drop-in boundary for ONE path of anabudisa/metric-amg-examples: building and
 * applying the HAZmath metric-AMG preconditioner inside the cbc.block Krylov
 * solve.  The reference crosses Python -> C through the `haznics` SWIG module;
 * every entry point below names the reference-side call it replaces
 * (paths are relative to the reference checkout).
 *
 * Conventions
 *   - plain pointers and sizes only; the caller owns every array it passes in,
 *     the library copies what it keeps; outputs are caller-allocated.
 *   - every function returns 0 on success and a negative code on failure and
 *     never throws across the ABI; mamg_last_error() gives the message
 *     (mirrors cbc.block's RuntimeError when the C setup returns NULL).
 *   - a handle is not thread-safe; distinct handles are independent.
 *   - fp64 values, int32 indices (same widths as the reference's
 *     PETSc -> dCSRmat conversion, src/utils.py:108).
 */
#ifndef MAMG_H
#define MAMG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Integer values of the haznics macros used by src/amg_parameters.py:5-16,42
 * and src/input_metric.dat:68-99 (symbolic use only). */
enum {
  MAMG_UA_AMG = 1, MAMG_SA_AMG = 2,
  MAMG_V_CYCLE = 1, MAMG_W_CYCLE = 2, MAMG_AMLI_CYCLE = 3, MAMG_NL_AMLI_CYCLE = 4, MAMG_ADD_CYCLE = 5,
  MAMG_SMOOTHER_JACOBI = 1, MAMG_SMOOTHER_GS = 2, MAMG_SMOOTHER_SGS = 3,
  MAMG_SMOOTHER_SOR = 5, MAMG_SMOOTHER_SSOR = 6, MAMG_SMOOTHER_L1DIAG = 10,
  MAMG_VMB = 1, MAMG_MIS = 2, MAMG_MWM = 3, MAMG_HEC = 4, MAMG_HEM = 5,
  MAMG_SCHWARZ_FORWARD = 1, MAMG_SCHWARZ_BACKWARD = 2, MAMG_SCHWARZ_SYMMETRIC = 3,
  MAMG_SOLVER_DEFAULT = 0, MAMG_SOLVER_VFGMRES = 4, MAMG_SOLVER_GCG = 5, MAMG_SOLVER_GCR = 6, MAMG_SOLVER_UMFPACK = 32,
  MAMG_OFF = 0, MAMG_ON = 1
};

/* The parameter dict of src/amg_parameters.py:3-89 as a struct; field names are
 * the dict keys (what haznics.param_amg_set_dict consumes upstream). */
typedef struct mamg_params {
  int32_t AMG_type;          /* UA_AMG | SA_AMG */
  int32_t cycle_type;        /* V_CYCLE | W_CYCLE | AMLI_CYCLE | NL_AMLI_CYCLE | ADD_CYCLE */
  int32_t max_levels;
  int32_t maxit;             /* cycles per apply */
  int32_t smoother;          /* SMOOTHER_* */
  double  relaxation;        /* SOR/SSOR weight, Jacobi damping */
  int32_t presmooth_iter;
  int32_t postsmooth_iter;
  int32_t coarse_dof;
  int32_t coarse_solver;     /* 32 = direct; 0 = HAZmath's iterative solve to tol*1e-4, served by the same
                              * dense inverse (the exact solution the iteration converges to) */
  int32_t coarse_scaling;    /* ON | OFF */
  int32_t aggregation_type;  /* VMB | HEM | HEC */
  double  strong_coupled;
  int32_t max_aggregation;
  int32_t amli_degree;       /* degree of the AMLI polynomial (AMLI_CYCLE), at most 15 */
  int32_t Schwarz_levels;
  int32_t Schwarz_mmsize;
  int32_t Schwarz_maxlvl;
  int32_t Schwarz_type;      /* SCHWARZ_FORWARD | BACKWARD | SYMMETRIC */
  int32_t Schwarz_blksolver; /* 32 = direct; 0 = iterative upstream, served by the same block inverses */
  int32_t print_level;
  int32_t nl_amli_krylov_type; /* NL_AMLI_CYCLE: SOLVER_GCG (5) = flexible CG K-cycle, anything else GCR
                                * (HAZmath's default branch); haznics.AMG_param() default 4 */
  int32_t reserved[7];
} mamg_params;

typedef struct mamg_handle_s* mamg_handle;

/* error text of the last failing call in this thread */
const char* mamg_last_error(void);
/* library build tag, e.g. "mamg 0.1 sm_100a" */
const char* mamg_version(void);

/* haznics.AMG_param() defaults == src/utils.py:60-82 (the inline dict used when
 * the reference passes no parameters). */
int mamg_params_default(mamg_params* p);

/* ---- setup: replaces block.algebraic.hazmath.metricAMG(A, W, idofs=, parameters=)
 *      (src/utils.py:86,88) and AMG(A, parameters=) (src/utils.py:40).
 *      A is the monolithic CSR that PETSc_to_dCSRmat(A) hands to HAZmath
 *      (src/utils.py:108); idofs are 0-based monolithic indices
 *      (src/bidomain_2d.py:192, src/emi_2d.py:205, src/emi_3d.py:134-138);
 *      n_idofs = 0 means "no interface dofs" (plain AMG).  Host only: needs no GPU. */
int mamg_setup(const mamg_params* p, int32_t n, const int32_t* indptr, const int32_t* indices,
               const double* data, int32_t n_idofs, const int32_t* idofs, mamg_handle* out);
/* Same, with a row partition for multi-GPU execution: part[i] in [0, nparts) is the owner of row i
 * (NULL = one part).  Aggregates never cross parts, so restriction and prolongation stay local to
 * an owner on every level; the hierarchy (and hence the iteration count) depends on the partition,
 * not on how many GPUs later execute it. */
int mamg_setup_partitioned(const mamg_params* p, int32_t n, const int32_t* indptr, const int32_t* indices,
                           const double* data, int32_t n_idofs, const int32_t* idofs, const int32_t* part,
                           int32_t nparts, mamg_handle* out);
int mamg_part_export(mamg_handle h, int32_t level, int32_t* part);

/* ---- import of an externally built hierarchy (north_star: "the AMG setup ... may remain the
 *      reference's HAZmath CPU setup exported once per problem, so that apply is compared on an
 *      identical hierarchy").  One record per level, finest first, natural ordering, the same arrays
 *      mamg_level_export / mamg_schwarz_export / mamg_prolongator_export / mamg_part_export hand out
 *      (so export -> import is a round trip).  Pointers that do not apply may be NULL: agg on the
 *      coarsest level, gs_skip / patch_* without Schwarz patches, P_* for unsmoothed aggregation,
 *      part for an unpartitioned hierarchy, color (then the library colours the level itself).
 *      The colourings are validated (rows / patches of one colour must not couple); the caller's arrays
 *      are copied.  coarse_inv = dense row-major inverse of the coarsest operator or NULL (computed). */
typedef struct mamg_level_arrays {
  int32_t n;                    /* rows of A_l */
  int32_t n_aggregates;         /* rows of A_{l+1}; 0 on the coarsest level */
  int32_t n_colors;             /* Gauss-Seidel colours (0 with color == NULL) */
  int32_t n_patches;            /* Schwarz patches on this level */
  int32_t n_patch_colors;
  int32_t reserved;
  const int32_t* indptr;        /* CSR of A_l */
  const int32_t* indices;
  const double*  data;
  const int32_t* agg;           /* [n] aggregate of every row, -1 = none */
  const int32_t* color;         /* [n] */
  const uint8_t* gs_skip;       /* [n] 1 = smoothed by Schwarz only */
  const int32_t* patch_ptr;     /* [n_patches + 1] */
  const int32_t* patch_dofs;
  const int32_t* patch_seed;    /* [n_patches] */
  const int32_t* patch_color;   /* [n_patches] */
  const int32_t* P_indptr;      /* smoothed prolongator (SA_AMG), n x n_aggregates */
  const int32_t* P_indices;
  const double*  P_data;
  const int32_t* part;          /* [n] owner part */
} mamg_level_arrays;
int mamg_import_hierarchy(const mamg_params* p, int32_t nlevels, const mamg_level_arrays* levels,
                          const double* coarse_inv, int32_t nparts, mamg_handle* out);
int mamg_destroy(mamg_handle h);

/* ---- hierarchy introspection / export (natural ordering) so that the CPU oracle
 *      applies the cycle on the identical hierarchy (north_star) */
int mamg_num_levels(mamg_handle h, int32_t* nlevels);
/* info[0]=rows info[1]=nnz info[2]=n_aggregates info[3]=n_colors info[4]=n_patches
 * info[5]=n_patch_entries info[6]=n_patch_colors info[7]=max_patch_size
 * info[8]=matrix entries in all patch rows (sum over patches of the nnz of their rows)
 * info[9]=packed patch-inverse entries (sum of s(s+1)/2) info[10]=nnz of the smoothed
 * prolongator P (SA_AMG levels, else 0) info[11]=nnz of the pattern before explicit zeros were
 * dropped (info[1] counts the stored entries, which is what every kernel streams) */
int mamg_level_info(mamg_handle h, int32_t level, int64_t info[12]);
int mamg_level_export(mamg_handle h, int32_t level, int32_t* indptr, int32_t* indices, double* data,
                      int32_t* agg, int32_t* color, uint8_t* gs_skip);
int mamg_schwarz_export(mamg_handle h, int32_t level, int32_t* patch_ptr, int32_t* patch_dofs,
                        int32_t* patch_seed, int32_t* patch_color);
/* smoothed prolongator P (rows x n_aggregates CSR) of an SA_AMG level; R = P' */
int mamg_prolongator_export(mamg_handle h, int32_t level, int32_t* indptr, int32_t* indices, double* data);
/* dense row-major inverse of the coarsest operator, n_c*n_c doubles */
int mamg_coarse_export(mamg_handle h, double* inv);
int mamg_setup_seconds(mamg_handle h, double* seconds);

/* ---- device: upload the hierarchy (colour-permuted CSR per level) to one B200.
 *      stream = a cudaStream_t cast to void* the library launches on (NULL: the
 *      library creates its own non-blocking stream). Fails if no CUDA device. */
int mamg_to_device(mamg_handle h, int32_t device, void* stream);
/* Multi-GPU upload: the rank and world size are known when the hierarchy goes to the device, so that
 * rank r stores only the matrix rows of its own parts on the row-distributed levels ("halo mode":
 * device memory and upload time divide by the number of ranks; kernels' updates travel as halo index
 * lists to the neighbour ranks only).  mamg_dist_init / mamg_dist_peers follow as before.
 * MAMG_HALO=0 keeps the round-1 scheme (whole hierarchy on every rank, updated ranges all-gathered). */
int mamg_to_device_dist(mamg_handle h, int32_t device, void* stream, int32_t rank, int32_t world);
int mamg_set_stream(mamg_handle h, void* stream);

/* ---- multi-GPU, one process per GPU (torch.distributed / torchrun launches the ranks).
 *      Every rank calls mamg_setup_partitioned with the same matrix and partition and
 *      mamg_to_device_dist(rank, world) on its GPU; rank 0 creates an NCCL id (mamg_nccl_unique_id,
 *      128 bytes) that the host side broadcasts; then mamg_dist_init and mamg_dist_peers.  Afterwards
 *      apply / pcg run row-distributed on the levels with at least MAMG_DIST_MIN_ROWS rows: rank r
 *      executes the rows of parts [r*P/world, (r+1)*P/world); in halo mode (default) it stores only
 *      those rows and sends what its neighbours read (halo index lists over peer memory), dots are
 *      rank-ordered all-reduces; smaller levels are executed redundantly by every rank.  Vectors
 *      handed over the ABI are complete on every rank.  world == 1 is valid (a partitioned hierarchy
 *      on one GPU: same numbers, no communication). */
int mamg_nccl_unique_id(void* out128);
int mamg_dist_init(mamg_handle h, int32_t rank, int32_t world, const void* unique_id128);
int mamg_collective_count(mamg_handle h, int64_t* count, int32_t reset);
/* bytes this rank has stored into peers' memory by halo pushes / all-reduces since the last reset (halo mode) */
int mamg_exchange_bytes(mamg_handle h, int64_t* bytes, int32_t reset);
/* Peer-memory exchange (same node, NVLink): every rank publishes the CUDA IPC handle of its vector
 * arena (mamg_ipc_handle, 64 bytes), the host side all-gathers them, mamg_dist_peers maps the peers'
 * arenas.  From then on the owner of a row range stores it directly into the peers' vectors and
 * raises a flag there; NCCL is only the fallback (MAMG_P2P=0).  A host barrier must separate
 * mamg_dist_peers from the first apply. */
int mamg_ipc_handle(mamg_handle h, void* out64);
int mamg_dist_peers(mamg_handle h, const void* handles64_per_rank);
int mamg_device_bytes(mamg_handle h, int64_t* bytes);
/* switch between V_CYCLE and W_CYCLE on an existing hierarchy (the hierarchy does not depend on the
 * cycle type; src/amg_parameters.py configures W, BASELINE.json's metric names the V-cycle) */
int mamg_set_cycle(mamg_handle h, int32_t cycle_type);
/* free the host copy of the level matrices once they are on the device (exports fail afterwards;
 * sizes stay available): saves host memory when several ranks hold large hierarchies on one node */
int mamg_release_host(mamg_handle h);
/* block until everything queued on the handle's stream has finished */
int mamg_sync(mamg_handle h);

/* ---- apply: replaces haznics.apply_precond(b_np, x_np, precond) that cbc.block's
 *      Precond.matvec runs for every B*r inside ConjGrad (src/bidomain_2d.py:205-206).
 *      z = B r, one cycle (params.maxit cycles) from a zero initial guess.
 *      on_device = 0: r, z are host arrays (copied inside the call, synchronous);
 *      on_device = 1: r, z are device pointers in the caller's (natural) dof order,
 *      the call is asynchronous on the handle's stream. */
int mamg_apply(mamg_handle h, const double* r, double* z, int32_t on_device);
/* The block variant R.T * Minv * R of src/utils.py:45-53 (xii.ReductionOperator concatenates the
 * block_vec, .T splits the result) without the concatenated copies: the nblocks (<= 8) blocks of the
 * block_vec are handed over as they are; sizes[] are their lengths (sum = matrix size).  With
 * on_device = 1 the boundary gather / scatter kernels address the blocks by offsets. */
int mamg_apply_blocks(mamg_handle h, int32_t nblocks, const int32_t* sizes, const double* const* r_blocks,
                      double* const* z_blocks, int32_t on_device);

/* y = A_level x on the device copy of the hierarchy (natural order, host or device
 * arrays as for mamg_apply); what dolfin's A*x (PETSc MatMult) does in the Krylov
 * loop (src/bidomain_2d.py:206) for level 0. */
int mamg_spmv(mamg_handle h, int32_t level, const double* x, double* y, int32_t on_device);
/* one pre-smoothing application (params.presmooth_iter sweeps of params.smoother, and
 * Schwarz where configured) on level `level`: x <- S(x, b). For kernel parity tests. */
int mamg_smooth(mamg_handle h, int32_t level, const double* b, double* x, int32_t post,
                int32_t on_device);

/* ---- Krylov: replaces block.iterative.ConjGrad(A, precond=B, tolerance=, maxiter=,
 *      relativeconv=) followed by x = AAinv * b (src/bidomain_2d.py:205-206,
 *      src/emi_2d.py:211-212).  A is level 0 of the hierarchy.  Stopping rule of
 *      cbc.block: sqrt(r.Br) <= tolerance (relative == 0) or <= tolerance*sqrt(r0.Br0)
 *      (relative == 1); relative == 2 selects HAZmath's own rule ||r||_2 <= tolerance*||r0||_2
 *      (linear_stop_type 1 of src/input_metric.dat:54, used by fenics_metric_solver_xd_1d).  residuals has room for maxiter+1 entries, alphas/betas for
 *      maxiter (may be NULL).  x holds the initial guess on entry when
 *      use_initial_guess != 0, else it is ignored. */
int mamg_pcg(mamg_handle h, const double* b, double* x, double tolerance, int32_t relative,
             int32_t maxiter, int32_t use_initial_guess, int32_t on_device, int32_t* niters,
             double* residuals, double* alphas, double* betas);
/* the same solve on a block system (ConjGrad(AA, precond=R.T*Minv*R) * bb of src/emi_2d.py:207-212):
 * right-hand side and solution as block_vec blocks, see mamg_apply_blocks */
int mamg_pcg_blocks(mamg_handle h, int32_t nblocks, const int32_t* sizes, const double* const* b_blocks,
                    double* const* x_blocks, double tolerance, int32_t relative, int32_t maxiter,
                    int32_t use_initial_guess, int32_t on_device, int32_t* niters, double* residuals,
                    double* alphas, double* betas);
/* preconditioned MINRES / restarted right-preconditioned GMRES with the same surface
 * (block.iterative.MinRes / LGMRES share ConjGrad's constructor upstream). */
int mamg_minres(mamg_handle h, const double* b, double* x, double tolerance, int32_t relative,
                int32_t maxiter, int32_t on_device, int32_t* niters, double* residuals);
int mamg_gmres(mamg_handle h, const double* b, double* x, double tolerance, int32_t relative,
               int32_t maxiter, int32_t restart, int32_t on_device, int32_t* niters,
               double* residuals);

/* ---- measurement helpers (bench.py): kernel launches issued on the handle's stream
 *      since the last reset, and algorithmic bytes (SURVEY 8d model) of one cycle. */
int mamg_launch_count(mamg_handle h, int64_t* launches, int32_t reset);
/* while collecting (before mamg_profile(h, 0, ...)): the same times split by level,
 * ms_level_class[level*16 + class], max_levels rows */
int mamg_profile_levels(mamg_handle h, double* ms_level_class, int32_t max_levels);
/* algorithmic bytes of one Schwarz sweep over all patches of a level (device layout: row values,
 * 16-bit local columns, neighbourhood lists, packed inverses; see csrc/cuda/schwarz.cuh) */
int mamg_schwarz_sweep_bytes(mamg_handle h, int32_t level, int64_t* bytes);
/* per-kernel-class timing with CUDA events on the handle's stream.  on=1 starts collecting;
 * on=0 stops and returns, for the classes {0 spmv, 1 gs, 2 schwarz, 3 restrict, 4 scale,
 * 5 prolong, 6 coarse, 7 vector, 8 dot, 9 exchange (multi-GPU peer-memory push + flag barrier)}, the
 * summed kernel time in ms and the launch counts
 * (arrays of 16 entries).  Collecting adds two event records per launch; never leave it on
 * inside a timed region. */
int mamg_profile(mamg_handle h, int32_t on, double* ms_per_class, int64_t* launches_per_class);
int mamg_cycle_bytes(mamg_handle h, int64_t* bytes);
/* Static race check of the device layout (the colour correctness of SURVEY 5 is a data-race property):
 * counts, over every launch the smoothers would issue, the rows of one Gauss-Seidel colour launch that
 * couple to each other and the patches of one conflict colour that write or read a dof another patch of
 * that launch writes.  Both counts must be 0.  Runs on the arrays the kernels stream. */
int mamg_race_check(mamg_handle h, int64_t* gs_conflicts, int64_t* patch_conflicts, int64_t* launches_checked);
/* device-side statistics of one level (SURVEY 8b `mamg_stats`): out[0]=rows out[1]=stored nnz
 * out[2]=structural nnz (before zero dropping) out[3]=sliced-ELL entry slots (padding included)
 * out[4]=device bytes of the whole handle out[5]=Schwarz patches out[6]=unique stored patch blobs
 * out[7]=algorithmic bytes of one Schwarz sweep (shared blobs once per colour) out[8]=the same with
 * every patch owning its data (SURVEY 8d stored-factor model) out[9]=GS colours out[10]=patch colours
 * out[11]=1 if the level runs inside the persistent tail kernel out[12]=1 sliced-ELL row kernels
 * out[13]=1 CSR entries kept on the device out[14]=row blocks (parts) out[15]=1 Schwarz fast path
 * out[16]=largest patch out[17]=largest outside neighbourhood of a patch out[18..23] reserved */
int mamg_stats(mamg_handle h, int32_t level, int64_t out[24]);

/* ---- synthetic systems of BASELINE.json's configs: P1 on UnitSquare/UnitCube
 *      (right-diagonal / 6-tet Kuhn split, lexicographic dofs), the matrices the
 *      reference assembles with FEniCS_ii and flattens with ii_convert
 *      (src/bidomain_2d.py:64-68,96-97,178; src/emi_2d.py:90-94,125-126).
 *      indptr == NULL: only nrows/nnz are returned.  Otherwise *nnz is the capacity of
 *      indices/data on entry (rows * 2 * 3^dim always suffices) and the actual nnz on return. */
int mamg_assemble_scalar(int32_t dim, const int32_t* ncell, const double* h, double cK, double cM,
                         int64_t* nrows, int64_t* nnz, int32_t* indptr, int32_t* indices,
                         double* data);
int mamg_assemble_bidomain(int32_t dim, int32_t ncell, double kappa1, double kappa2, double gamma,
                           int64_t* nrows, int64_t* nnz, int32_t* indptr, int32_t* indices,
                           double* data);
int mamg_assemble_emi(int32_t dim, int32_t ncell, double kappa1, double kappa2, double gamma,
                      int64_t* nrows, int64_t* nnz, int32_t* indptr, int32_t* indices,
                      double* data);

#ifdef __cplusplus
}
#endif
#endif /* MAMG_H */

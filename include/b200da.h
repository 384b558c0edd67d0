/*
 * b200da.h — C ABI of the B200-native LETKF / ETKF analysis engine (libb200da.so).
 *
 * This is the drop-in boundary for the hot path of tobifinn/torch-assimilate (pytassim 0.2.1):
 * everything between "obs-space variables are ready" (interface/base.py:359-379) and "analysis is
 * assembled" (interface/base.py:257-278) for `LETKF.assimilate` / `ETKF.assimilate`
 * (interface/base.py:419-512 -> interface/filter.py:96-165 -> interface/letkf.py:104-148 /
 * interface/etkf.py:99-120).  Each entry point names the reference interface it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / xarray types;
 *   - every function returns 0 (B200DA_OK) or a negative b200da_status; b200da_strerror() explains it;
 *   - unless a name ends in `_host`, every data pointer is a DEVICE pointer owned by the caller;
 *     the library allocates only plan-owned scratch;
 *   - all work is enqueued on the given CUDA stream (a `cudaStream_t` passed as void*); set-up calls
 *     (`b200da_set_grid`, `b200da_bin_obs`) read a few scalars back and therefore synchronise the stream;
 *   - one plan per device and per thread of control; plans are independent of each other;
 *   - the CUDA device must be sm_100 (B200); there is no CPU fallback (B200DA_ERR_NO_DEVICE).
 *
 * Layouts (identical to the reference's in-memory layout)
 *   state      X   (n_slices, k, N)  grid index fastest; n_slices = n_var * n_time  (pytassim/state.py:114)
 *   obs perts  Yn  (k, M)            obs index fastest; already multiplied by R^{-1/2}
 *   innovation d   (M)               (y - mean_k Hx) R^{-1/2}                        (interface/base.py:367-372)
 *   coordinates    (n_coord, N|M)    struct-of-arrays float64: the columns after the time column of
 *                                    `_extract_state_information` / `_extract_obs_information`
 *                                    (interface/mixin_local.py:45-69)
 *   weights    W   (N, k, k)         dims (grid, ensemble, ensemble_new)             (interface/letkf.py:136)
 */
#ifndef B200DA_H
#define B200DA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* The library is built with -fvisibility=hidden: only the entry points declared here are exported. */
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

typedef struct b200da_plan b200da_plan;

typedef enum {
    B200DA_OK = 0,
    B200DA_ERR_INVALID = -1,      /* bad argument (null pointer, k < 2, ...)                                  */
    B200DA_ERR_SIZE = -2,         /* Yn / d observation sizes differ: the reference's ValueError
                                     (pytassim/core/base.py:28-38)                                            */
    B200DA_ERR_UNSUPPORTED = -3,  /* metric / taper / dtype / ensemble size outside the implemented set       */
    B200DA_ERR_NO_DEVICE = -4,    /* no CUDA device or not sm_100: there is no CPU fallback                   */
    B200DA_ERR_CUDA = -5,         /* a CUDA runtime call failed; see b200da_last_cuda_error()                 */
    B200DA_ERR_STATE = -6,        /* call order: set_grid and bin_obs must precede letkf / neighbour_*         */
    B200DA_ERR_NOMEM = -7,
    B200DA_ERR_OVERFLOW = -8      /* a grid-point block needed more candidate cell columns than the kernels hold (reported by
                                     b200da_pending_status): the analysis of that launch is invalid                        */
} b200da_status;

/* Distance functions.  The reference takes an arbitrary Python `dist_func` (localization/gaspari_cohn.py:60-69,125);
 * the engine implements this closed set.  metric_params: PERIODIC1D -> {period}; HAVERSINE -> {sphere radius}. */
typedef enum {
    B200DA_METRIC_ABS1D = 0,      /* |x_g - x_o|                        (examples/benchmark_letkf.py:85-87)  */
    B200DA_METRIC_PERIODIC1D = 1, /* min(|x_g - x_o|, L - |x_g - x_o|)                                        */
    B200DA_METRIC_EUCLID = 2,     /* sqrt(sum_c (x_gc - x_oc)^2), n_coord in 1..3                             */
    B200DA_METRIC_HAVERSINE = 3   /* great circle, coords = (lat, lon) in degrees                             */
} b200da_metric;

typedef enum {
    B200DA_TAPER_GC = 0,          /* GaspariCohn     (localization/gaspari_cohn.py:78-136)                    */
    B200DA_TAPER_GCINF = 1        /* GaspariCohnInf  (localization/gaspari_cohn.py:172-254)                   */
} b200da_taper;

typedef enum { B200DA_F64 = 0, B200DA_F32 = 1 } b200da_dtype;

/* How the k x k ensemble-space problem of every grid point is solved (both replace core/utils.py:26-93 +
 * core/etkf.py:57-77 and agree to rounding):
 *   NEWTON_SCHULZ  (default) (C + aI)^(-1/2) by a scaled coupled Newton-Schulz iteration: k x k x k products on the
 *                  FP64 tensor pipe, one warp per matrix up to k = 56;
 *   JACOBI         shared-memory parallel cyclic-Jacobi eigendecomposition (k <= 111). */
typedef enum { B200DA_SOLVER_NEWTON_SCHULZ = 0, B200DA_SOLVER_JACOBI = 1 } b200da_solver;

/* Operations of a kernel program (b200da_plan_set_kernel): the kernels of pytassim/kernels/*.py that are element-wise
 * functions of x.y, |x|^2, |y|^2, and the three compositions of kernels/base_kernels.py:40-161.  p0 / p1 per operation: */
typedef enum {
    B200DA_KOP_LINEAR = 0,    /* x.y                                      kernels/linear.py:62-63        (-, -)          */
    B200DA_KOP_GAUSS = 1,     /* exp(-|x/l - y/l|^2 / 2)                  kernels/rbf.py GaussKernel     (l, -); RBFKernel: l = sqrt(0.5 / gamma) */
    B200DA_KOP_POLY = 2,      /* (x.y + c)^p                              kernels/polynomial.py          (p, c)          */
    B200DA_KOP_TANH = 3,      /* tanh(alpha x.y + c)                      kernels/tanh.py                (alpha, c)      */
    B200DA_KOP_RATIONAL = 4,  /* (1 + |x/l - y/l|^2 / (2 w))^(-w)         kernels/rational.py            (l, w)          */
    B200DA_KOP_SCALE = 5,     /* c                                        kernels/scale.py               (c, -)          */
    B200DA_KOP_DIAG = 6,      /* c I on K(perts, perts), 0 on K(perts, obs)   kernels/diag.py            (c, -)          */
    B200DA_KOP_ADD = 7,       /* K1 + K2   (pops two)                     base_kernels.py:89-90                          */
    B200DA_KOP_MUL = 8,       /* K1 * K2                                  base_kernels.py:119-120                        */
    B200DA_KOP_POW = 9        /* K1 ^ K2                                  base_kernels.py:160-161                        */
} b200da_kernel_op;
#define B200DA_MAX_KERNEL_OPS 16

/* ---- plan ------------------------------------------------------------------------------------------------ */

/* Replaces the constructor state of LETKF(localization=GaspariCohn(length_scale, dist_func, epsilon),
 * inf_factor=rho) (interface/letkf.py:72-92, localization/gaspari_cohn.py:60-69).
 * k: ensemble size; n_slices: n_var * n_time rows of the state that share one weight matrix per grid point. */
int b200da_plan_create(b200da_plan** plan, int k, int n_slices, int n_coord, int metric,
                       const double* metric_params, int n_metric_params, const double* radius, int n_radius,
                       double epsilon, double inf_factor, int dtype, int taper);
void b200da_plan_destroy(b200da_plan* plan);

/* A `dist_func` that returns SEVERAL distance rows gets one Gaspari-Cohn factor per row, each row divided by its own entry of
 * `length_scale`, the factors multiplied (localization/gaspari_cohn.py:124-134).  The engine's form of it: the plan's metric
 * (row 0, radius[0] of b200da_plan_create) times n_extra rows |x_g - x_o| on further coordinate columns (vertical level,
 * time, ...; 0 <= n_extra <= 2) with radii extra_radius[0..n_extra).  Afterwards every coordinate array (b200da_set_grid,
 * b200da_bin_obs, b200da_letkf_host) carries n_coord + n_extra rows.  Call before b200da_set_grid.  B200DA_TAPER_GC only
 * (GaspariCohnInf evaluates a single distance, gaspari_cohn.py:216-254).  The neighbour search uses row 0 (every factor is
 * <= 1, so this is conservative); FP32 plans with extra rows use the DMMA Gram instead of the tcgen05 one. */
int b200da_plan_set_extra(b200da_plan* plan, int n_extra, const double* extra_radius);

/* Replaces the `kernel` argument of KETKF / LKETKF (interface/ketkf.py:70-90, interface/lketkf.py:84-110) and
 * KETKFModule._estimate_weights (core/ketkf.py:69-100): a postfix program of n_ops operations (b200da_kernel_op, parameters
 * p0[i], p1[i]; leaves push K(x, y), ADD / MUL / POW pop two values; must leave exactly one value).  With a program set,
 * b200da_letkf and b200da_etkf_weights[_from_gram] turn every augmented Gram into the double-centred kernel matrix and the
 * centred kernel column of the observations before the ensemble-space solve; the Gram always comes from the FP64 DMMA
 * kernel with the innovation row inside the tiles (it carries d.d, which the distance-based kernels need).  n_ops = 0
 * returns to the plain ETKF.  Kernel matrices that are not positive semi-definite (tanh, powers) need
 * B200DA_SOLVER_JACOBI, which clamps negative eigenvalues as core/utils.py:58 does; the Newton-Schulz solver assumes a
 * positive semi-definite centred kernel matrix.  Ensemble sizes that are a multiple of 8 are limited to k <= 120. */
int b200da_plan_set_kernel(b200da_plan* plan, int n_ops, const int* ops, const double* p0, const double* p1);

/* Replaces `_extract_state_information` + the dask chunking of the grid (interface/mixin_local.py:50-69,
 * interface/letkf.py:121): bins the N grid points into cells and forms blocks of neighbouring grid points
 * (one CTA each).  grid_coord: (n_coord, N) float64 device. */
int b200da_set_grid(b200da_plan* plan, const double* grid_coord, int64_t n_grid, void* stream);

/* Replaces `_extract_obs_information` + the per-grid-point boolean gather of wrapper_localization
 * (interface/mixin_local.py:45-47, interface/wrapper.py:91-97): bins the M observations into cells and builds
 * the cell-ordered, observation-major staging copy of [Yn; d].  obs_coord: (n_coord, M) float64 device;
 * Yn: (k, M); d: (M) of the plan's dtype. */
int b200da_bin_obs(b200da_plan* plan, const double* obs_coord, const void* Yn, const void* d, int64_t n_obs,
                   void* stream);

/* Replaces BaseAssimilation._get_obs_space_variables for observations with a diagonal R (interface/base.py:359-379,
 * observation.py:241-245,277-279; the step right before the hot path): from the ensemble of observation
 * equivalents HX (k, M) (dataset-, time-, obs_grid_1-stacked as _stack_obs does, interface/base.py:223-241), the
 * observations y (M) and their error variances (M):  Yn = (HX - mean_k HX) / sqrt(var),  d = (y - mean_k HX) / sqrt(var).
 * The ensemble mean is summed in member order, so FP64 results are bit-identical to the reference's numpy arithmetic.
 * Correlated R (Cholesky of an M x M matrix, observation.py:247-275) stays on the host. */
int b200da_obs_prep(b200da_plan* plan, const void* HX, const void* y, const void* variance, int64_t n_obs, void* Yn, void* d,
                    void* stream);

/* b200da_obs_prep with the observation operator fused in front, for operators that select grid columns of one state variable
 * (obs_ops/lorenz_96/identity.py:88-92 `sel(var_name='x').sel(grid=points)`, examples/benchmark_letkf.py:100-104
 * `sel(grid=obs_grid, method='nearest')`, time selection of obs_ops/base_ops.py:63-75):
 *   HX[i][j] = Xp[src_offset[j] + i * member_stride]
 * Xp: the pseudo state (n_var, n_time, k, N) on the device (member_stride = N); src_offset[j] (device int64) = element offset
 * of member 0 of observation j's (variable, time, grid column).  Gathered values are exact copies, the arithmetic is
 * b200da_obs_prep's: FP64 results are bit-identical to operator -> _get_obs_space_variables. */
int b200da_obs_gather_prep(b200da_plan* plan, const void* Xp, const int64_t* src_offset, int64_t member_stride, const void* y,
                           const void* variance, int64_t n_obs, void* Yn, void* d, void* stream);

int64_t b200da_num_blocks(const b200da_plan* plan);      /* grid-point blocks (unit of multi-GPU sharding)     */
int64_t b200da_num_grid(const b200da_plan* plan);
int64_t b200da_num_obs(const b200da_plan* plan);
/* first grid slot (position in the block-sorted order) of block b, b in [0, num_blocks]; host query */
int64_t b200da_block_offset(const b200da_plan* plan, int64_t block);
/* all num_blocks + 1 offsets at once into a HOST array */
int b200da_block_offsets(const b200da_plan* plan, int64_t* offsets_host);
/* copies the block-sorted grid order (N int32: slot -> original grid index) to a device buffer */
int b200da_grid_order(const b200da_plan* plan, int32_t* order_out, void* stream);

/* ---- the hot path ---------------------------------------------------------------------------------------- */

/* Replaces the whole per-grid-point loop of LETKF.estimate_weights (interface/letkf.py:127-143:
 * localize_obs -> sqrt(w) gather -> ETKFModule.forward, i.e. localization/gaspari_cohn.py:97-136,
 * interface/wrapper.py:54-98, core/etkf.py:57-103, core/utils.py:26-93) fused with _apply_weights
 * (interface/base.py:257-278) for the grid points of blocks [block_begin, block_end).
 * X, Xa: (n_slices, k, N) device, plan dtype; only the columns of the analysed grid points are written.
 * W_opt: (N, k, k) or NULL.  n_ambiguous_opt: device int64 counter (or NULL) incremented for every
 * (grid point, obs) pair whose taper value lies within 1e-13 of epsilon. */
int b200da_letkf(b200da_plan* plan, const void* X, void* Xa, void* W_opt, int64_t block_begin,
                 int64_t block_end, int64_t* n_ambiguous_opt, void* stream);

/* The first half of b200da_letkf alone: the localization-weighted augmented Gram matrix of every grid point of the
 * blocks, G_g = sum_j w_gj [y_j; d_j][y_j; d_j]^T (rows 0..k-1: C = Y~ Y~^T of core/etkf.py:68 after the sqrt(w) gather of
 * interface/wrapper.py:91-97; row k: b = Y~ d~^T of core/etkf.py:72), dense (N, k+1, k+1) FP64 in original grid order,
 * lower triangle filled, element (k, k) and the upper triangle zero.  Parity hook for the Gram kernels. */
int b200da_letkf_gram(b200da_plan* plan, double* gram_out, int64_t block_begin, int64_t block_end, void* stream);

/* Host-buffer convenience used for end-to-end timing: uploads (obs_coord, Yn, d, X), bins, analyses all
 * blocks, downloads Xa.  All pointers are HOST pointers (pinned memory makes the copies asynchronous).
 * The grid must have been set with b200da_set_grid. */
int b200da_letkf_host(b200da_plan* plan, const double* obs_coord_host, const void* Yn_host, const void* d_host,
                      int64_t n_obs, const void* X_host, void* Xa_host, void* stream);

/* After b200da_letkf_host: analyse blocks [block_begin, block_end) again from the inputs still staged on the device (with
 * the overrides set since) and download the analysis again — the repeat step of the ambiguity protocol below. */
int b200da_letkf_host_blocks(b200da_plan* plan, void* Xa_host, int64_t block_begin, int64_t block_end, void* stream);

/* Local-observation index lists = np.nonzero(use_obs)[0] of GaspariCohn.localize_obs
 * (localization/gaspari_cohn.py:135) for every grid point, as CSR in ORIGINAL grid order, ascending obs id.
 * Two passes: count -> caller scans -> fill.  w_opt receives the taper weights (before sqrt);
 * ambiguous_opt (same length as idx) is 1 where |w - epsilon| < 1e-13 (also set in entries that were
 * rejected: see b200da_neighbour_count's n_ambiguous). */
int b200da_neighbour_count(b200da_plan* plan, int64_t* counts, int64_t* n_ambiguous_opt, void* stream);
int b200da_neighbour_fill(b200da_plan* plan, const int64_t* offsets, int32_t* idx, double* w_opt,
                          uint8_t* ambiguous_opt, void* stream);
/* (grid points with offsets[g + 1] == offsets[g] are skipped by the fill pass: zero the counts of the grid points that are
 * not wanted before the scan to materialise the lists of a subset only — the full CSR of cfg3 has 5.7e10 entries) */
/* All pairs (grid index, obs index, w) inside the ambiguity band, accepted or not; capacity-limited. */
int b200da_neighbour_ambiguous(b200da_plan* plan, int64_t capacity, int64_t* grid_idx, int64_t* obs_idx,
                               double* w, int64_t* n_found, void* stream);

/* The ambiguity protocol for the mask `use_obs = weights > epsilon` of GaspariCohn.localize_obs (localization/
 * gaspari_cohn.py:135).  The reference evaluates the taper with numpy's `**`, whose last bits depend on the host's numpy
 * build; the device evaluates the same polynomial by Horner's rule.  The two values agree to ~1e-15, so the mask can differ
 * only for pairs with |w - epsilon| < 1e-13.  b200da_letkf records every such pair it meets (FP64 taper path; up to 4096):
 *   b200da_pending_status      synchronises the stream, copies the recorded pairs (original grid index, original obs index,
 *                              unmasked device taper value) to HOST arrays, resets the record, and returns
 *                              B200DA_ERR_OVERFLOW if a kernel since the last call had to drop candidate cells (see
 *                              b200da_status).  *n_found_host may exceed `capacity` / 4096: then only that many were kept;
 *   b200da_plan_set_overrides  hands the host's decisions back (HOST arrays): the weight of pair i becomes w_host[i]
 *                              (0 = not a local observation) in b200da_letkf*, b200da_letkf_gram and the neighbour lists
 *                              until the next b200da_bin_obs / b200da_set_grid; n = 0 clears the list;
 *   b200da_blocks_of_grid      the block that holds each original grid index (HOST arrays), so that only the blocks whose
 *                              mask changed are analysed again. */
int b200da_pending_status(b200da_plan* plan, int64_t capacity, int64_t* grid_idx_host, int64_t* obs_idx_host, double* w_host,
                          int64_t* n_found_host, void* stream);
int b200da_plan_set_overrides(b200da_plan* plan, int64_t n, const int64_t* grid_idx_host, const int64_t* obs_idx_host,
                              const double* w_host, void* stream);
int b200da_blocks_of_grid(b200da_plan* plan, int64_t n, const int64_t* grid_idx_host, int64_t* block_host, void* stream);

/* ---- global ETKF (no localization) ------------------------------------------------------------------------ */

/* Replaces ETKFModule.forward on the whole observation vector (interface/etkf.py:99-120, core/etkf.py:79-103):
 * W (k, k) = w_mean + w_perts.  M = 0 gives sqrt(inf_factor) * I (core/etkf.py:91-95). */
int b200da_etkf_weights(b200da_plan* plan, const void* Yn, const void* d, int64_t n_obs, void* W, void* stream);

/* Replaces BaseAssimilation._apply_weights (interface/base.py:257-278): Xa = mean + (X - mean) W with one
 * global W (k, k) (per_grid = 0) or W (N, k, k) (per_grid = 1).  X, Xa: (n_slices, k, N). */
int b200da_apply_weights(b200da_plan* plan, const void* X, const void* W, int per_grid, int64_t n_grid, void* Xa,
                         void* stream);

/* Observation- and state-sharded form of the two calls above (the reference hands the whole observation
 * vector to one torch call, interface/etkf.py:99-120; on G GPUs each rank owns a column range of Yn / d and of the state):
 *   b200da_etkf_gram              the augmented Gram  [Yn; d][Yn; d]^T  of an observation range (core/etkf.py:68,72 =
 *                                 core/utils.py:153-173 on that range): Yn points at the first column of the range inside a
 *                                 (k, ld_obs) array, d at its first element; gram_out: dense (k+1) x (k+1) FP64 row-major,
 *                                 lower triangle filled (row k = b), element (k, k) and the upper triangle zero.  The Grams of
 *                                 disjoint ranges add up (one all-reduce of (k+1)^2 doubles);
 *   b200da_etkf_weights_from_gram evd -> rev_evd x2 -> W = w_mean + w_perts (core/utils.py:26-93, core/etkf.py:57-77,102) from
 *                                 the summed Gram; n_obs_total = 0 gives sqrt(inf_factor) * I (core/etkf.py:91-95);
 *   b200da_apply_weights_cols     _apply_weights (interface/base.py:257-278) on grid columns [col_begin, col_end) of the
 *                                 (n_slices, k, n_grid) arrays X / Xa (other columns untouched); per_grid W is (n_grid, k, k). */
int b200da_etkf_gram(b200da_plan* plan, const void* Yn, const void* d, int64_t n_obs, int64_t ld_obs, double* gram_out,
                     void* stream);
int b200da_etkf_weights_from_gram(b200da_plan* plan, const double* gram, int64_t n_obs_total, void* W, void* stream);
int b200da_apply_weights_cols(b200da_plan* plan, const void* X, const void* W, int per_grid, int64_t col_begin,
                              int64_t col_end, int64_t n_grid, void* Xa, void* stream);
/* The same update of columns [col_begin, col_end) with one (k, k) W, fused with the all-gather of the state-sharded global
 * ETKF: every analysed tile is stored into Xa AND into the n_peers (<= 7) arrays Xa_peers[i] of the same shape that live in
 * the memory of other GPUs of the node (peer-accessible device pointers, e.g. from CUDA IPC / symmetric memory), straight from
 * the kernel's registers over NVLink.  Replaces "update, then gather the chunks" of the reference's dask graph
 * (interface/etkf.py:99-120, xr.apply_ufunc over grid chunks + compute()).  The caller orders the ranks (a barrier after the
 * call before anyone reads its Xa). */
/* Columns [col_begin, col_end) of a (rows, ld) row-major array from src to dst (either may be peer memory of another GPU of the
 * node): one strided copy on the copy engines (cudaMemcpy2DAsync), the transport of the overlapped update + all-gather in
 * pytassim_b200/parallel.py (ShardedETKF).  No plan needed. */
int b200da_peer_copy_cols(void* dst, const void* src, int64_t rows, int64_t ld, int64_t col_begin, int64_t col_end,
                          int elem_bytes, void* stream);
int b200da_apply_weights_cols_peers(b200da_plan* plan, const void* X, const void* W, int64_t col_begin, int64_t col_end,
                                    int64_t n_grid, void* Xa, int n_peers, void* const* Xa_peers, void* stream);

/* ---- iterative ensemble Kalman smoother (IEnKS), one iteration of the weight update -------------------------------- */

/* Replaces IEnKSTransformModule.forward / IEnKSBundleModule.forward (core/ienks.py:117-174) behind
 * LocalizedIEnKSTransform.inner_loop / LocalizedIEnKSBundle.inner_loop (interface/lienks.py:68-118: localized_module with the
 * incoming weights skipped by the localization, args_to_skip = (0,)) followed by _apply_weights (interface/base.py:257-278):
 * like b200da_letkf, but every grid point starts from its incoming weights W_in ((N, k, k) when w_per_grid, else one
 * (k, k) matrix for all grid points: the prior identity of the first iteration) and takes one step with learning rate
 * tau in (0, 1].  epsilon > 0 selects the bundle variant (dh_dw = Yn / epsilon, ienks.py:168-174), epsilon <= 0 the
 * transform variant (dh_dw = Wp^-1 Yn, ienks.py:74-75).  W_out (N, k, k) receives the updated weights, Xa the state
 * updated with them (either may not be null).  Grid points without local observations keep W_in (ienks.py:143); they are
 * recognised by an all-zero Gram diagonal and innovation row, so local observations whose localized perturbations and
 * innovations are all exactly zero count as none.  The plan's
 * inf_factor is not used (the IEnKS has none).  k is limited by shared memory (k <= 96). */
int b200da_letkf_ienks(b200da_plan* plan, const void* X, void* Xa, const void* W_in, int w_per_grid, void* W_out, double tau,
                       double epsilon, int64_t block_begin, int64_t block_end, void* stream);

/* Replaces IEnKSTransform.inner_loop / IEnKSBundle.inner_loop (interface/ienks.py:96-118): the global (unlocalized) update of
 * one (k, k) weight matrix from all observations.  m = 0 copies W_in to W_out. */
int b200da_etkf_ienks_weights(b200da_plan* plan, const void* Yn, const void* d, int64_t m, const void* W_in, double tau,
                              double epsilon, void* W_out, void* stream);

/* ---- multi-GPU helpers ------------------------------------------------------------------------------------ */

/* Pack / unpack the analysed columns of blocks [block_begin, block_end) between the (n_slices, k, N) layout and
 * a (n_slices * k, n_cols) buffer with row stride ld >= n_cols elements (0: dense) in block-sorted order: the all-gather
 * payload of the grid-sharded run, packed straight into / scattered straight out of a slot of the gather buffer. */
int b200da_pack_columns(b200da_plan* plan, const void* Xa, int64_t block_begin, int64_t block_end, void* packed,
                        int64_t ld, void* stream);
int b200da_unpack_columns(b200da_plan* plan, const void* packed, int64_t block_begin, int64_t block_end, void* Xa,
                          int64_t ld, void* stream);

/* ---- misc -------------------------------------------------------------------------------------------------- */

const char* b200da_strerror(int status);
const char* b200da_last_cuda_error(void);
int b200da_version(void);
/* number of kernel launches issued by this library since load (bench.py's gpu_launches) */
int64_t b200da_launch_count(void);
/* rows of the augmented [Yn; d] that the FP64 Gram kernel of this plan accumulates with DFMA next to its DMMA tiles */
int b200da_gram_extra_rows(const b200da_plan* plan);
/* name of the Gram kernel configuration chosen for the plan, e.g. "letkf_gram_f64_kt7_g8_w2" */
const char* b200da_kernel_name(const b200da_plan* plan);
/* device time (ms) of the last b200da_letkf main-kernel launch on this plan, measured with CUDA events on
 * the launch stream when timing was enabled with b200da_enable_timing(plan, 1); synchronises. */
int b200da_enable_timing(b200da_plan* plan, int on);
float b200da_last_kernel_ms(b200da_plan* plan);
/* after b200da_last_kernel_ms: summed device time of the Gram kernels (which = 0) / solve kernels (which = 1) */
float b200da_last_phase_ms(b200da_plan* plan, int which);

/* Per-launch phase statistics of the fused kernel (diagnostics; small atomics overhead when enabled):
 * out16 = {sum of Gram-phase cycles over CTAs, sum of EVD+transform+update cycles, Jacobi sweeps, EVDs,
 *         set-up cycles, staged tiles, 0, 0, Jacobi step profile x5, 0, 0, 0}.  b200da_get_stats synchronises the device. */
int b200da_collect_stats(b200da_plan* plan, int on);
/* choose the ensemble-space solver of b200da_letkf (b200da_solver) */
int b200da_set_solver(b200da_plan* plan, int solver);
int b200da_get_stats(b200da_plan* plan, int64_t* out8);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif

#ifdef __cplusplus
}
#endif
#endif /* B200DA_H */

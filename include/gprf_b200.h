/*
 * gprf_b200 - C-ABI of the B200-native GPRF objective-and-gradient hot path.
 *
 * The reference (davmre/gprf) has no FFI of its own: its boundary for this
 * path is the Python method surface of `GPRF` (gprf.py:85-331).  This header
 * is the plain-C boundary a maintainer binds from Python with ctypes (see
 * INTEGRATION.md); `gprf_b200/gprf.py` is that binding, mirroring the
 * reference class.  Each entry point names the reference code it replaces.
 *
 * All pointers are HOST pointers unless the parameter name ends in `_dev`.
 * All floating point is IEEE double; indices are int64 (numpy default) for
 * point indices and int32 for block ids.  Every call is synchronous with
 * respect to the host on return unless stated otherwise.
 *
 * theta = [noise_var, signal_var, l_0, ..., l_{nls-1}]   (gprf.py:160-167)
 *   nls = dx for dfn "euclidean", 2 for dfn "lld";  ncov = 2 + nls.
 */
#ifndef GPRF_B200_H
#define GPRF_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gprf_ctx* gprf_handle;

/* distance / weight functions of treegp's VectorTree used by the reference
 * (synthetic.py:149, run_seismic.py:299-301) */
enum { GPRF_DFN_EUCLIDEAN = 0, GPRF_DFN_LLD = 1 };
enum { GPRF_WFN_SE = 0, GPRF_WFN_MATERN32 = 1 };

/* return codes */
enum {
  GPRF_OK = 0,
  GPRF_ERR_NOT_PD = 1,        /* jitchol: "not positive definite, even with jitter." gpy_linalg.py:97 */
  GPRF_ERR_NONPOS_DIAG = 2,   /* jitchol: "not pd: non-positive diagonal elements"   gpy_linalg.py:85 */
  GPRF_ERR_ARG = 3,
  GPRF_ERR_CUDA = 4,
  GPRF_ERR_NO_STRUCTURE = 5
};

#define GPRF_MAX_NCOV 5

/* Replaces GPRF.__init__'s data capture (gprf.py:85-97): Y (n x dy, C order) is
 * uploaded once and stays resident in HBM; X changes every evaluation. */
int gprf_create(gprf_handle* out, int device, long long n, int dx, int dy,
                const double* Y, int dfn_id, int wfn_id);

int gprf_destroy(gprf_handle h);

/* Replaces the structure consumed by llgrad (gprf.py:206-216, 299-330):
 *   block_ptr[B+1], perm[block_ptr[B]] : concatenated GPRF.block_idxs
 *   edges[2E] : GPRF.neighbors as (i, j) pairs; pair unit e stacks block i's
 *               rows first, then block j's (gprf.py:310-330)
 *   unit_mask[B+E] or NULL : multi-GPU sharding - only units with mask != 0
 *               are evaluated on this device (replaces Pool.map_async,
 *               gprf.py:218-233).  Unit u < B is the unary of block u,
 *               unit B+e the pair of edge e.
 * Called whenever block membership changes (GPRF.update_X, gprf.py:169-174). */
int gprf_set_structure(gprf_handle h, int n_blocks, const long long* block_ptr,
                       const long long* perm, int n_edges, const int* edges,
                       const unsigned char* unit_mask);

/* The same structure in two halves: the edge list is fixed at construction
 * (gprf.py:111-116) while block membership changes with every update_X.
 * gprf_set_edges also selects the multi-GPU share: with shard_world > 1 this
 * device evaluates the units a longest-processing-time split on the work model
 * W(s) = s^3 + 4 s^2 dy assigns to shard_rank (same rule on every rank). */
int gprf_set_edges(gprf_handle h, int n_edges, const int* edges, int shard_rank,
                   int shard_world);
int gprf_set_blocks(gprf_handle h, int n_blocks, const long long* block_ptr,
                    const long long* perm);
/* Restrict the evaluations to the units with mask != 0 (mask[B+E]; NULL restores
 * all units / the shard selected by gprf_set_edges).  With raw_weights != 0 every
 * active unit counts once instead of with (1 - deg_i): a mask holding one unit
 * then returns that unit's own objective and gradients, which is how
 * GPRF.llgrad_unary / llgrad_joint (gprf.py:299-330) and subset evaluations run on
 * the live structure. */
int gprf_set_unit_mask(gprf_handle h, const unsigned char* mask, int n_units, int raw_weights);

/* Device-side partitioners: block membership is recomputed from X on the GPU,
 * bit-exact with the reference's numpy expressions.
 *   grid: Blocker.block_clusters (block_clustering.py:17-26).  centers B x dx,
 *         csq[b] = numpy's sum(c_b**2).
 *   tree: PDTree.recluster + pdtree_cluster's longitude wrap
 *         (pdtree_clustering.py:65-94).  Flattened nodes: center/direction
 *         (n_nodes x 2), cut, child (n_nodes x 2, negative = -(leaf)-1).
 *   dot_mode selects how the host BLAS rounds the length-dx dot product
 *         (0: k-ascending FMA, 1: k-descending FMA, 2: no FMA); the Python
 *         layer probes it against numpy on the actual data. */
int gprf_set_grid_partitioner(gprf_handle h, int n_blocks, const double* centers,
                              const double* csq, int dot_mode);
int gprf_set_tree_partitioner(gprf_handle h, int n_nodes, const double* center,
                              const double* direction, const double* cut,
                              const int* child, int root, int n_leaves,
                              double wrap_add, double wrap_mod, int dot_mode);

/* Replaces GPRF.update_X's `block_idxs = block_fn(new_X)` (gprf.py:169-174):
 * recompute block membership on the device from X (host) / X_dev. */
int gprf_reblock(gprf_handle h, const double* X);
int gprf_reblock_device(gprf_handle h, const double* X_dev, void* stream);
int gprf_block_count(gprf_handle h, int* n_blocks, long long* plen);
int gprf_get_blocks(gprf_handle h, long long* block_ptr, long long* perm);

/* update_X + llgrad in one call: H2D of X, device reblock, evaluation, D2H. */
int gprf_llgrad_reblock(gprf_handle h, const double* X, const double* theta,
                        int ncov, int grad_X, int grad_cov, double* ll,
                        double* gradX, double* gradTheta, int* failed_unit);

/* Replaces GPRF.llgrad (gprf.py:206-296) for host buffers:
 *   ll         <- sum_e ll_e + sum_i (1 - deg_i) ll_i
 *   gradX      <- n x dx (C order) or NULL   (grad_X=False)
 *   gradTheta  <- ncov or NULL               (grad_cov=False)
 *   failed_unit<- first unit that was not PD after jitter (or -1)
 * H2D of X and D2H of the results happen inside the call.  With a unit_mask
 * the outputs are this device's partial sums. */
int gprf_llgrad(gprf_handle h, const double* X, const double* theta, int ncov,
                int grad_X, int grad_cov, double* ll, double* gradX,
                double* gradTheta, int* failed_unit);

/* Same evaluation with X already in HBM and the result left in HBM:
 *   out_dev[0] = ll, out_dev[1..GPRF_MAX_NCOV] = gradTheta (zero padded),
 *   out_dev[1+GPRF_MAX_NCOV ...] = gradX (n*dx) when grad_X.
 * Work is enqueued on `stream` (a cudaStream_t); the call returns after the
 * stream has drained (the per-unit Cholesky status must be read to apply the
 * jitter rule). */
int gprf_llgrad_device(gprf_handle h, const double* X_dev, const double* theta,
                       int ncov, int grad_X, int grad_cov, double* out_dev,
                       void* stream, int* failed_unit);

/* Per-unit results of the last evaluation (llgrad_unary / llgrad_joint,
 * gprf.py:299-330): ll_units[B+E]; jitter_units[B+E] (0 = none); either may
 * be NULL. */
/* Optimiser glue (gprfopt.py:396-409; run_seismic.py:157-179).  gprf_set_x_prior uploads an
 * independent Gaussian prior on the locations (mean n x dx; inv_var and grad_scale per column,
 * grad_scale NULL = 1; mean NULL removes the prior).  gprf_neg_objective then returns what the
 * reference's L-BFGS callback hands to scipy,
 *     f = -(ll + x_prior(X)) up to the prior's constant,  g = -(gradX + d x_prior/dX) * grad_scale,
 * formed on the device in the combination epilogue, so that only (f, g) come back and the host makes
 * no pass over n x dx arrays.  gradTheta (when grad_cov) is the plain d ll / d theta for the host's
 * log-theta chain rule.  reblock != 0: block membership is recomputed from X on the device first. */
int gprf_set_x_prior(gprf_handle h, const double* mean, const double* inv_var, const double* grad_scale);
int gprf_neg_objective(gprf_handle h, const double* X, const double* theta, int ncov, int grad_cov,
                       int reblock, double* f, double* g, double* gradTheta, int* failed_unit);
/* gprf_llgrad_device without the host round trip (resident path): launches are enqueued on `stream`, the
 * status word of the evaluation is written as a double (0 = ok) to *status_dev by the combination kernel.
 * For multi-GPU use (replaces the Pool of gprf.py:218-233): the status travels inside the all-reduce of
 * the packed results; when the reduced status is non-zero every rank repeats the evaluation with
 * gprf_llgrad_device (jitter rule / oversized blocks).  *enqueued = 0: the structure is not of the
 * resident kind and the evaluation ran synchronously. */
int gprf_llgrad_device_nosync(gprf_handle h, const double* X_dev, const double* theta, int ncov,
                              int grad_X, int grad_cov, double* out_dev, double* status_dev, void* stream,
                              int* enqueued, int* failed_unit);
int gprf_unit_results(gprf_handle h, double* ll_units, double* jitter_units);

/* Replaces GPRF.compute_neighbors (gprf.py:119-150): for every block pair
 * (i, j), j < i, max |k(x_p, x_q)| / signal_var over p in i, q in j
 * (noise-free kernel).  maxk[B*B] symmetric, diagonal 1.  Uses the structure's
 * blocks and the given X / theta. */
int gprf_block_max_kernel(gprf_handle h, const double* X, const double* theta,
                          int ncov, double* maxk);

/* Replaces GPRF.kernel (gprf.py:333-343): K (n1 x n2).  X2 == NULL -> K(X1,X1)
 * + noise_var*I; else cross kernel without noise. */
int gprf_kernel_matrix(gprf_handle h, const double* X1, long long n1,
                       const double* X2, long long n2, const double* theta,
                       int ncov, double* K);

/* Debug / test access: copy unit u's working matrix ((sp+yr) x sp, row major)
 * and sizes after the last evaluation.  Any pointer may be NULL. */
/* GPRF.dKdx (gprf.py:345-355; mode 0: row p of d k(x_p, .)/d x_{p,which}, entry p zeroed, n doubles)
 * and the lengthscale part of GPRF.dKdi (gprf.py:372-374; mode 1: d K / d l_which, n x n doubles)
 * for one point set X (n x dx, host). */
int gprf_kernel_deriv(gprf_handle h, const double* X, long long n, const double* theta, int ncov,
                      int mode, int p, int which, double* out);

int gprf_debug_unit(gprf_handle h, int unit, int* s, int* sp, int* yr,
                    double* M, double* alpha, double* gx_unit);

/* Device time of the kernels of the last evaluation, in milliseconds
 * (CUDA events on the launching stream), and the number of kernel launches. */
/* out (s x dy, row major) = L Y_unit for a block unit, L the Cholesky factor the last evaluation
 * computed for it (jitter rule included) and Y_unit the unit's rows of the Y given to gprf_create.
 * The dense branch of the reference's sample_y (synthetic.py:106-114, y = jitchol(K) z): upload the
 * normal draw z as Y, evaluate the objective of the one-block structure, call this. */
int gprf_unit_lmul(gprf_handle h, int unit, double* out);
int gprf_last_timing(gprf_handle h, float* ms, int* launches);

/* Debug: timeline of the fused kernel, 512 (tag, %globaltimer ns) pairs per CTA.
 * out == NULL: allocate for the first n_ctas CTAs (0 frees); out != NULL: copy back
 * n_ctas * 512 * 2 values.  Tags: see scripts/trace_fused.py. */
int gprf_debug_trace(gprf_handle h, int n_ctas, unsigned long long* out);

/* K^-1 normally lives only in registers (it is turned into G = Alpha Alpha^T - dy K^-1
 * in place).  With keep != 0 every evaluation also writes it to the lower triangle
 * of the unit's working matrix, where gprf_debug_unit reads it (dpotri's result,
 * gpy_linalg.py:150-171; needed by tests and by a train_predictor port). */
int gprf_set_keep_kinv(gprf_handle h, int keep);

/* Scheduling knob (no effect on results: both schedules run the same tile tasks in the
 * same arithmetic order and agree bit for bit): units of up to `nt` 64-point tiles
 * are evaluated by the fused one-CTA-per-unit kernel, larger ones by the multi-launch
 * tile pipeline.  Default 0 = tile pipeline for every unit, the faster schedule on
 * B200 for every workload measured; environment override
 * GPRF_FUSED_NT.  When some unit exceeds `nt` (so the tile pipeline runs anyway), only
 * units of up to min(nt, 4) tiles stay fused (GPRF_FUSED_MIXED_NT). */
int gprf_set_fused_nt(gprf_handle h, int nt);

/* Edge factorisations reuse block i's Cholesky factor (default on).  The pair unit
 * of edge (i, j) stacks block i's rows first (gprf.py:310-330), so the leading
 * floor(n_i / 64) tiles of its factor L, of U = L^-T, of the forward solve L^-1 Y and
 * of the logdet are those of block i's own unit (gprf.py:299-308), bit for bit; a
 * pair evaluated by the tile pipeline reads them from there and factors only the
 * Schur complement (the tiles that involve block j).  Results are bit-identical
 * with on = 0, where the reference's independent pdinv per unit
 * (gpy_linalg.py:219-240) is executed literally.  Units that need jitter are
 * re-factored on their own, as jitchol does (gpy_linalg.py:77-97).
 * Pairs evaluated by the fused one-CTA-per-unit kernel reuse as well once there
 * are at least 8 per SM of them (GPRF_FUSED_SHARE_MIN), because their parent blocks
 * then need a launch of their own in front; on = 2 forces that for any count.
 * gprf_factor_reuse_stats: pair units that reuse, and tile tasks not executed. */
int gprf_set_factor_reuse(gprf_handle h, int on);
int gprf_factor_reuse_stats(gprf_handle h, int* n_units, long long* n_tiles);

/* Optional per-kernel-family timing: when on, every launch is bracketed by CUDA
 * events on the launching stream; gprf_family_timing returns, for the last
 * evaluation, the summed device time (ms) and launch count of each family
 * (gprf_family_name(i), i < GPRF_N_FAMILIES). */
#define GPRF_N_FAMILIES 13
int gprf_set_profiling(gprf_handle h, int on);
int gprf_family_timing(gprf_handle h, float* ms, int* launches);
const char* gprf_family_name(int fam);

/* ---- resident path (gprf_b200/csrc/resident.cuh) -------------------------------------------------
 * Units of small blocks (<= 160 points per block, dy <= 64) are evaluated by one CTA each, entirely
 * out of shared memory: block units export their factor, pair units factor only the Schur complement
 * of block j on top of block i's exported factor (the reference computes an independent pdinv per
 * unit, gprf.py:299-330; same values to rounding).  Tried automatically by gprf_llgrad*,
 * with ONE host synchronisation per evaluation; an evaluation in which a unit does not fit or a
 * Cholesky pivot fails (the jitter rule, gpy_linalg.py:77-97) is re-run by the tile pipeline.
 * gprf_set_resident(h, 0) switches it off (env GPRF_RESIDENT=0 likewise).
 * Environment switches read at gprf_create (diagnostics and A/B runs; the defaults are the measured best):
 *   GPRF_RES_EARLY=0      pairs wait for all of the parent block's factor exports instead of W alone
 *   GPRF_RES_DEFER=0      a pair whose parent is still running waits at its start instead of after its own K_ji
 *   GPRF_RES_SORTBLK=0    block units are dealt to the CTAs in id order instead of size order
 *   GPRF_BUCKET_SPLIT=k   placement CTAs of the re-blocking launch (0: one CTA does everything; default one per 4096 points)
 *   GPRF_RES_WATCHDOG_S=s limit of the kernel's spin waits (default ~2 s; raise it under compute-sanitizer) */
int gprf_set_resident(gprf_handle h, int on);
int gprf_resident_stats(gprf_handle h, long long* evals, long long* fallbacks, int* last_status);
/* Layout constants of the resident path's export records (tests): returns their number. */
int gprf_resident_layout(long long* out, int n);
/* Debug: dump the shared-memory matrices of `unit` after `phase` during the next evaluations. */
int gprf_set_resident_debug(gprf_handle h, int unit, int phase);
int gprf_get_resident_debug(gprf_handle h, double* out_dump, int block, double* out_export, int unit,
                            double* out_unit);

const char* gprf_strerror(int code);
const char* gprf_last_error(gprf_handle h);
int gprf_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* GPRF_B200_H */

/* g3b.h — C ABI of the B200-native exact-GP hot path (libg3b.so).
 *
 * Plain pointers and sizes only; no torch / numpy types.  All matrices that
 * cross this boundary are row-major, C-contiguous float64 HOST arrays owned by
 * the caller (NumPy / Theano `perform` storage); the library copies in/out and
 * keeps its own device buffers per context.
 *
 * Each entry point names the reference (griosd/g3py) interface it replaces,
 * path:line relative to the reference tree.
 *
 * Error convention: functions return 0 on success, <0 for a bad
 * argument or a CUDA error (text via g3_last_error).  Numerical trouble is NEVER
 * an error: it is reported per batch item in `status[]` / `info[]`, and the
 * Python Op maps it to the reference's conventions (L = 1e-10*I, logp = -1e30).
 */
#ifndef G3B_H
#define G3B_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct g3_ctx g3_ctx;

/* ---- kernel expression tree ------------------------------------------------------------
 * Post-order flattening of g3py's Kernel algebra:
 *   leaves  KernelStationary subclasses   g3py/processes/hypers/kernels.py:96-110,360-472
 *   nodes   KernelSum/Prod/Scale/Shift    g3py/processes/hypers/kernels.py:192-245
 *   `dim0,dim1` = Hypers.dims column slice g3py/processes/hypers/__init__.py:55-83
 * theta indices address the NATURAL-space (already exponentiated) hyper vector of one
 * batch item, in pymc3 creation order (var, then metric hypers; SIN: var, freq[], rate[]).
 */
enum {
  G3_K_SE = 1,    /* var*exp(-d),  d = sum_k 0.5*rate_k^2*(x_ik-x_jk)^2   kernels.py:434-436, metrics.py:100-102 */
  G3_K_OU = 2,    /* var*exp(-d),  d = sum_k rate_k*|x_ik-x_jk|           kernels.py:429-431, metrics.py:89-91  */
  G3_K_MAT32 = 3, /* var*(1+s)exp(-s), s = sqrt(3d)                        kernels.py:406-412 */
  G3_K_MAT52 = 4, /* var*(1+s+5d/3)exp(-s), s = sqrt(5d)                   kernels.py:415-421 */
  G3_K_RQ = 5,    /* var*(1+d/alpha)^-alpha                                kernels.py:388-403 */
  G3_K_SIN = 6,   /* var*exp(+2*sum_k rate_k*sin^2(pi*(x_ik-x_jk)*freq_k)) kernels.py:470-472 */
  G3_K_NOISE = 7, /* var*I when x1 is x2, zeros otherwise                  kernels.py:360-371 */
  G3_K_WN = 8,    /* var*I when x1 is x2, var*#equal coords otherwise      kernels.py:374-385 */
  G3_K_COS = 9,   /* var*prod_k cos(2 pi (x_ik-x_jk) freq_k)                kernels.py:462-467 */
  G3_K_SINC = 10, /* var*prod_k sinc_k, sinc = sin(2 pi^2 d f)/(2 pi^2 f d), 1 at d = 0   kernels.py:475-482 */
  G3_K_SM = 11,   /* var*exp(-2 pi^2 sum_k d_k^2 rate_k^2)*prod_k cos(2 pi d_k freq_k)    kernels.py:485-487 */
  /* non-stationary leaves: the diagonal of cov(x, x) depends on x */
  G3_K_DOT = 12,  /* var*(bias + sum_k rate_k^2 x_ik x_jk)^p: KernelDot/ARD_Dot (p1_idx < 0: bias = 0), LIN and POL with
                     ARD_DotBias (bias = theta[p1_idx]); integer p >= 1 in flags bits 8..15 (0 means 1)
                     kernels.py:82-96,319-336, metrics.py:110-137 */
  G3_K_BW = 13,   /* var*prod_k min(x_ik, x_jk)  (Brownian)                  kernels.py:291-293, metrics.py:49-51 */
  G3_K_VAR = 14,  /* var (constant kernel; NIL = var fixed to 0)              kernels.py:296-306, 309-320 */
  G3_K_EQ = 15,   /* sum_k [x_ik == e1][x_jk == e2] (+ [x_ik == e2][x_jk == e1] with G3_KF_EQ2): KernelEquals / KernelEquals2 over the
                   * DeltaEq / DeltaEq2 metrics.  e1 = `value`; e2 = e1, or with G3_KF_EQ2 the double whose low / high words are
                   * p0_idx / p1_idx.  No variance, no hypers (var_idx = -1).      kernels.py:262-288, metrics.py:38-51 */
  G3_K_SUM = 16, G3_K_PROD = 17, G3_K_SCALE = 18, G3_K_SHIFT = 19,
  G3_K_MAX = 20   /* elementwise max(k1, k2); ties send the gradient to both (Theano's maximum)  kernels.py:247-257 */
};
/* flags of G3_K_DOT: bits 8..15 the integer power p; G3_KF_NN = the NN kernel var*arcsin(2m/(1+2m)^2) on the same
 * metric m = bias + sum_k rate_k^2 x_ik x_jk, as its single-argument cov() is written (elementwise in m_ij; kernels.py:339-348).
 * flag of G3_K_EQ: G3_KF_EQ2. */
#define G3_KF_NN 0x10000
#define G3_KF_EQ2 0x20000
#define G3_KF_POWER(flags) ((((flags) >> 8) & 0xff) ? (((flags) >> 8) & 0xff) : 1)

typedef struct {
  int32_t op;
  int32_t dim0, dim1;  /* [dim0, dim1) columns of X used by the metric */
  int32_t var_idx;     /* theta index of `var`; -1 = use `value` (KernelProd fixes k2.var = 1.0, kernels.py:215-219) */
  int32_t p0_idx;      /* theta index of rate[dim1-dim0] (SE/OU/MAT32/MAT52/RQ/SIN/SM), else -1 */
  int32_t p1_idx;      /* theta index of RQ alpha, or of SIN/COS/SINC/SM freq[dim1-dim0], else -1 */
  int32_t flags;       /* G3_KF_PROCESS_NOISE on the KernelNoise leaf EllipticalProcess adds itself (elliptical.py:26-28) */
  double value;        /* fixed var (var_idx < 0) or the constant of SCALE / SHIFT */
} g3_knode;

enum { G3_KF_PROCESS_NOISE = 1 };

#define G3_MAX_NODES 32   /* up to 16 leaves and their 15 operators (+1); trees of <= 16 nodes and <= 32 slots keep the fast Gram paths */
#define G3_MAX_THETA 64
#define G3_MAX_DIM 16

typedef struct {
  int32_t n_nodes;
  int32_t n_theta;     /* natural-space hypers per batch item consumed by this tree */
  g3_knode nodes[G3_MAX_NODES];
} g3_kernel_desc;

/* ---- per-item status bits ------------------------------------------------------------- */
enum {
  G3_ST_NONFINITE_INPUT = 1,  /* NaN/inf scrubbed while building K (tt_to_num, libs/tensors.py:90-92) */
  G3_ST_DIAG_SHIFT = 2,       /* min diag <= 0: tt_to_cov shift applied (libs/tensors.py:95-98) */
  G3_ST_JITTER = 4,           /* jitter ladder used (libs/tensors.py:203-213); tries in bits 8..15 */
  G3_ST_POTRF_FAILED = 8,     /* ladder exhausted -> caller applies the 1e-10*I fallback (libs/tensors.py:218-222) */
  G3_ST_NONFINITE_RESULT = 16 /* NaN/inf in L, u or delta -> caller returns -1e30 (gaussian.py:234-241) */
};

enum { G3_KIND_GAUSS = 0, G3_KIND_STUDENT = 1 };

/* flags for g3_gp_posterior */
enum { G3_POST_NOISE = 1,   /* noise=True: K** gets the Noise variance (elliptical.py:70,86-88) */
       G3_POST_COV = 2 };   /* also return the full M x M posterior covariance */

/* ---- context --------------------------------------------------------------------------
 * One context = one device + one stream (+ group streams) + cached workspaces.  Created lazily per
 * process (fork-safe: no CUDA state at library load).  A context is NOT thread-safe: use one per
 * thread; different contexts (same or different devices) may be used concurrently.  Replaces nothing in the reference; it is
 * where theano's `perform` keeps device state between calls (libs/tensors.py:215-222). */
int g3_ctx_create(int device, g3_ctx** out);
int g3_ctx_destroy(g3_ctx* ctx);
/* Release every cached workspace of the context (they are re-created on demand): lets one process run a 60 GiB
 * single-matrix factorisation after a batched evaluation that held 16 GiB of workspaces. */
int g3_ctx_trim(g3_ctx* ctx);
const char* g3_last_error(g3_ctx* ctx);
int g3_sync(g3_ctx* ctx);
/* constants of the reference graph, supplied by the host so that "strict" (float32-rounded)
 * and exact modes agree bit-for-bit with the oracle: jitter = float32(1e-6) (tensors.py:98,204). */
int g3_set_jitter(g3_ctx* ctx, double jitter_rel, int max_tries);
/* Blocking of the factorisation: tile columns (of 128) per right-looking outer block; a value
 * >= N/128 makes it fully left-looking.  0 (default) chooses from the batch size and N: left-looking
 * when the batch supplies the parallelism, outer blocks of 8 tile columns for few large matrices. */
int g3_set_potrf_block(g3_ctx* ctx, int w_outer);
/* Look-ahead of the blocked (right-looking) factorisation: the next panel is updated and factored on a second,
 * high-priority stream while the rest of the trailing update runs (default on; results do not depend on it). */
int g3_set_lookahead(g3_ctx* ctx, int on);
/* Split-K for GEMM launches with few tiles and a deep contraction (single-matrix evaluations): up to 8 CTAs share one
 * output tile, partial tiles are added in a fixed order (bitwise reproducible).  on = 1 (default): at least 128 of the
 * contraction per share; 2: shares down to 32 and the triangular solves too (experiment, measured slower); 0: off. */
int g3_set_splitk(g3_ctx* ctx, int on);
/* Column split of GEMM tiles: a launch of a few 64 x 128 tiles (triangular solves and column updates of single-matrix
 * evaluations) is bound by one SM's fp64 rate per tile, so each tile is spread over 2 or 4 CTAs (halves / quarters of its
 * columns; same summation order, bitwise identical results) until one CTA per SM is reached.  Default on. */
int g3_set_tile_split(g3_ctx* ctx, int on);
/* Triangular solves u = L^-1 r and alpha = L^-T u (gaussian.py:219-231 solve_lower_triangular / the reverse sweep of the
 * analytic gradient) as ONE launch each: a CTA owns a 128-row block and waits on release / acquire flags for the blocks it
 * depends on, instead of T = N/128 dependent launches.  Fixed summation order (bitwise reproducible).  Default on. */
int g3_set_trsv_fused(g3_ctx* ctx, int on);
/* Value-then-gradient callers (the Theano Ops: GPLogpOp.perform followed by GPLogpGradOp.perform on the same inputs, what
 * NUTS / BFGS drive through libs/tensors.py:174-263): with on = 1 a logp-only evaluation of <= 8 matrices also computes
 * U = L^-T behind its factorisation (side stream, as the fused value-and-gradient call does), so that g3_gp_grad_resume
 * starts from it.  Wasted work if no gradient follows; default off. */
int g3_set_speculate_grad(g3_ctx* ctx, int on);
/* Gradient path of few large matrices: compute U = L^-T block by block on a third stream while the look-ahead
 * factorisation is still running (default on; results do not depend on it). */
int g3_set_trtri_pipeline(g3_ctx* ctx, int on);
/* Arithmetic of the deep panel updates of the batched (B > 8) Cholesky:
 *   G3_GEMM_DMMA  (default) fp64 tensor-core GEMM (mma.sync m8n8k4, gemm.cu) everywhere;
 *   G3_GEMM_OZAKI the update of a 256-wide block column with all earlier columns runs on the INT8 tensor cores
 *                 (tcgen05.mma kind::i8, accumulators in tensor memory) as exact products of 9 seven-bit slices of L - fp64-
 *                 equivalent (error of an fp64 dot product), past the DMMA peak for contractions >= min_k (0 keeps the current
 *                 threshold, default 1024).  Everything else (in-block work, solves, K^-1) stays on the DMMA GEMM.
 * Replaces nothing in the reference (the arithmetic lives inside LAPACK dpotrf there, g3py/libs/tensors.py:198). */
enum { G3_GEMM_DMMA = 0, G3_GEMM_OZAKI = 1 };
int g3_set_gemm_mode(g3_ctx* ctx, int mode, int min_k);
int64_t g3_ozaki_launch_count(g3_ctx* ctx);
/* CUDA-graph replay for evaluations of up to 8 matrices (one chain / one BFGS step): the second g3_gp_run with the same
 * (kernel tree, kind, B, N, gradient, schedule switches, workspaces) captures its launch sequence, later ones replay it
 * with one cudaGraphLaunch.  Default on; results are identical (same kernels, same order).  g3_graph_replays counts replays. */
int g3_set_graphs(g3_ctx* ctx, int on);
int64_t g3_graph_replays(g3_ctx* ctx);
/* Number of batch groups g3_gp_run processes concurrently on separate streams (default 4, max 8;
 * 1 = a single stream, which is what the per-kernel timers of g3_prof_* need). */
int g3_set_groups(g3_ctx* ctx, int n_groups);

/* Device timing on the context's stream (CUDA events; used by bench.py). */
int g3_timer_begin(g3_ctx* ctx);
int g3_timer_end(g3_ctx* ctx, float* ms);
/* Optional per-kernel-class device timing (CUDA event pairs around each launch on the context's
 * stream).  g3_prof_read synchronises, sums the elapsed ms and launch counts per class
 * {0 dgemm_nt, 1 potrf_diag, 2 gram_fwd, 3 gram_vjp, 4 trsv (whole sweep), 5 other} and resets. */
int g3_prof_enable(g3_ctx* ctx, int on);
int g3_prof_read(g3_ctx* ctx, double* ms6, int64_t* launches6);
/* Test accessor: copy the first `bytes` of a named device workspace (e.g. "gp_A", "gp_U", "gp_Dinv")
 * to the host after synchronising the stream. */
int g3_debug_read(g3_ctx* ctx, const char* name, void* host, size_t bytes);
/* Determinism stress test: factor the same batch `iters` times and compare tile by tile with the first
 * factor; out[it] = {#mismatching tiles, batch item, tile row, tile col of the first mismatch}. */
int g3_debug_potrf_stress(g3_ctx* ctx, const g3_kernel_desc* desc, const double* theta, int B, int iters, int* out4,
                          double* tiles2);
int g3_debug_gemm_stress(g3_ctx* ctx, int rows, int B, int launches, int inplace, int kdepth, long long* out4);
/* Roofline denominators measured in the run (bench.py): sustained fp64 rate of the DMMA.8x8x4 tensor pipe and of the
 * DFMA pipe in TFLOP/s, and a STREAM-style 1 GiB device copy in GB/s (read + written bytes), each timed for about
 * `seconds` with CUDA events on the context's stream.  Any output pointer may be NULL (that leg is skipped).
 * Replaces nothing in the reference. */
int g3_debug_fp64_peak(g3_ctx* ctx, double seconds, double* dmma_tflops, double* dfma_tflops, double* copy_gbs);
/* Which kernel factors and inverts the 128x128 diagonal tiles of the blocked Cholesky (the serial head of every tile
 * column; replaces the unblocked part of dpotrf behind CholeskyRobust.perform, libs/tensors.py:198):
 * 2 (default) = the low-latency kernel of csrc/diag.cu, 1 = the first, bulk-synchronous one (kept for A/B timing). */
int g3_set_diag_variant(g3_ctx* ctx, int variant);
/* Times that kernel alone: `reps` back-to-back launches of B CTAs on copies of one synthetic SPD tile
 * (M M^T / 128 + cond_shift I), microseconds per launch from CUDA events; checks tile 0 on the host:
 * err4 = {max |L L^T - A| / max |A|, max |Dinv L - I|, |logdet - host|, info}.  For variant 2, stamps32 (64 entries) receives
 * the clock64 marks of the kernel's phases (CTA 0); a negative cond_shift (|.| is used) runs that extra launch with nothing in
 * the shadow of the serial factorisation (timing experiment; its outputs are incomplete).  Replaces nothing in the reference. */
int g3_debug_diag_time(g3_ctx* ctx, int variant, int B, int reps, double cond_shift, float* us_per_launch,
                       long long* stamps32, double* err4);
/* Number of kernels launched by this context since creation (bench.py "gpu_launches"). */
int64_t g3_launch_count(g3_ctx* ctx);

/* ---- observations ---------------------------------------------------------------------
 * Keeps the training inputs resident on the device across calls; replaces the per-call
 * re-upload of th_inputs in makefn.__call__ (libs/tensors.py:60-69). */
int g3_set_data(g3_ctx* ctx, const double* X, int N, int D);

/* ---- Gram matrix ----------------------------------------------------------------------
 * Kernel.cov(x1, x2) for B hyper samples (kernels.py:106-110,225-241; metrics.py:11-13),
 * with tt_to_num scrubbing (tensors.py:90-92).  X2 == NULL means cov(x1) (x1 is x2: Noise/WN
 * contribute var*I).  K_out: B x n1 x n2.  status may be NULL. */
int g3_gram(g3_ctx* ctx, const g3_kernel_desc* desc, const double* X1, int n1,
            const double* X2_or_NULL, int n2, int D, const double* theta, int B,
            double* K_out, int* status);

/* dtheta[b][p] = sum_ij W[b][i][j] * dK[b][i][j]/dtheta_p — the contraction Theano's reverse
 * mode performs through Kernel.cov (gradient(), tensors.py:11-22); dK/dtheta never stored. */
int g3_gram_vjp(g3_ctx* ctx, const g3_kernel_desc* desc, const double* X1, int n1,
                const double* X2_or_NULL, int n2, int D, const double* theta, int B,
                const double* W, double* dtheta);

/* ---- robust Cholesky ------------------------------------------------------------------
 * CholeskyRobust.perform / _cholesky (libs/tensors.py:197-222): lower Cholesky of B
 * matrices (row-major, n x n, leading dimension lda), in place on the host array; upper
 * triangle zeroed.  On failure of item b the jitter ladder is run on the device copy:
 * dK = mean(diag)*jitter_rel, x10 per try, max_tries.  info[b]: 0 ok, k>0 succeeded at
 * ladder try k, -1 exhausted (A[b] left untouched; caller applies 1e-10*I).
 * jitter[b] (may be NULL) = the dK finally added. */
int g3_potrf_robust(g3_ctx* ctx, double* A, int n, int lda, int B, int* info, double* jitter);
/* Same for ONE matrix, followed by the forward substitution u = L^-1 rhs on the device (rhs, u_out: n doubles).
 * Replaces `tsl.solve_lower_triangular(cholesky_robust(K), v)` (g3py/processes/hypers/transports.py:227-232). */
int g3_potrf_robust_solve(g3_ctx* ctx, double* A, int n, int lda, const double* rhs, double* u_out, int* info,
                          double* jitter);

/* ---- fused marginal likelihood + gradient ---------------------------------------------
 * WarpedGaussianDistribution.logp_cho / WarpedStudentTDistribution.logp_cho core
 * (gaussian.py:208-224, studentT.py:116-129) and its gradient (which the reference obtains
 * by tt.grad through CholeskyRobust.grad, tensors.py:224-260), for B hyper samples on the
 * data given to g3_set_data:
 *   K_b  = tt_to_cov(cov(X; theta_b))            elliptical.py:71
 *   L_b  = cholesky_robust(K_b)                  tensors.py:197-222
 *   u_b  = L_b^-1 delta_b, beta_b = u'u, logdet_b = sum log diag L_b
 *   alpha_b = K_b^-1 delta_b
 *   dtheta[b][p] = 1/2 sum_ij (c_b alpha alpha' - K^-1)_ij dK_ij/dtheta_p      (natural space)
 *   ddelta[b]    = -c_b alpha_b
 *   c_b = 1 (gauss) or (nu_b + N)/(nu_b - 2 + beta_b) (student; d r1/d beta, studentT.py:126)
 * delta: B x N (delta_stride = N) or one shared vector (delta_stride = 0).
 * dtheta / ddelta may be NULL (logp only: no inverse is formed).  status: B ints. */
int g3_gp_logp_grad(g3_ctx* ctx, const g3_kernel_desc* desc, int kind,
                    const double* delta, int delta_stride, const double* theta, int B,
                    const double* nu_or_NULL, double* beta, double* logdet,
                    double* dtheta_or_NULL, double* ddelta_or_NULL, int* status);

/* Finish the gradient of the LAST logp-only evaluation (g3_gp_logp_grad with dtheta = ddelta = NULL, nothing else run on
 * the context since): the factor L, its block inverses and u = L^-1 delta are still resident, so only the K^-1 stages run
 * (2/3 N^3 instead of a second N^3/3 factorisation + 2/3 N^3).  This is what makes a Theano value-and-gradient pair
 * (th_logp followed by th_dlogp on the same theta, stochastic.py:300-313; the forward is recomputed inside dlogp in the
 * reference) cost one factorisation.  Fails (<0) if no matching factor is resident.  dtheta: B x n_theta, ddelta: B x N. */
int g3_gp_grad_resume(g3_ctx* ctx, double* dtheta_or_NULL, double* ddelta_or_NULL);

/* Split form of the same call for benchmarking with inputs resident in HBM:
 * upload copies theta/delta/nu host->device, run launches the device pipeline only,
 * download copies the results device->host. */
int g3_gp_upload(g3_ctx* ctx, const g3_kernel_desc* desc, int kind, const double* delta,
                 int delta_stride, const double* theta, int B, const double* nu_or_NULL, int want_grad);
int g3_gp_run(g3_ctx* ctx);
int g3_gp_download(g3_ctx* ctx, double* beta, double* logdet, double* dtheta_or_NULL,
                   double* ddelta_or_NULL, int* status);

/* ---- posterior moments ----------------------------------------------------------------
 * EllipticalProcess.th_define_process posterior block (elliptical.py:78-107) for ONE theta:
 *   mean_out = K* K^-1 delta                (caller adds m(X*), elliptical.py:81-84)
 *   var_out  = max(diag(K** - K* K^-1 K*'), 0)                    elliptical.py:86-97
 *   cov_out  = K** - K* K^-1 K*'  (M x M, only with G3_POST_COV)
 * K* uses the cross form of `desc` (Noise contributes 0, kernels.py:367-371); K** gets the
 * Noise variance only with G3_POST_NOISE.  beta_out (may be NULL) = delta' K^-1 delta, which
 * StudentTProcess.th_scaling needs (studentT.py:36-43). */
int g3_gp_posterior(g3_ctx* ctx, const g3_kernel_desc* desc, const double* Xs, int M,
                    const double* delta, const double* theta, int flags,
                    double* mean_out, double* var_out, double* cov_out_or_NULL,
                    double* beta_out_or_NULL, int* status);

/* ---- multi-GPU (SURVEY §8 b/e) ---------------------------------------------------------------------------------
 * One process per GPU; the library owns the NCCL communicator (libnccl.so.2 is dlopen'ed on first use; no torch).
 * Not in the reference: its only parallelism is multiprocessing.Pool.map over chain groups
 * (g3py/processes/stochastic.py:775-783).  The host distributes the 128-byte id of rank 0 to the other ranks
 * (g3py_b200/comm.py: file or TCP rendezvous from MASTER_ADDR / MASTER_PORT / RANK / WORLD_SIZE).
 *   g3_comm_allgather : host buffers, bytes_per_rank from every rank, rank order - the (logp, dlogp) rows of a sharded
 *                       theta batch / chains, 8 B (P + 1) bytes per rank
 *   g3_comm_allreduce : host doubles in place, op 0 sum / 1 max / 2 min (max-over-ranks device times)
 */
#define G3_COMM_ID_BYTES 128
int g3_comm_get_unique_id(char* id_out /* G3_COMM_ID_BYTES */);
int g3_comm_init(g3_ctx* ctx, int nranks, int rank, const char* id /* G3_COMM_ID_BYTES; may be NULL for nranks == 1 */);
int g3_comm_destroy(g3_ctx* ctx);
int g3_comm_size(g3_ctx* ctx);
int g3_comm_rank(g3_ctx* ctx);
int g3_comm_barrier(g3_ctx* ctx);
int g3_comm_allgather(g3_ctx* ctx, const void* send, void* recv, size_t bytes_per_rank);
int g3_comm_allreduce(g3_ctx* ctx, double* vals, int n, int op);

/* Exact GP too large / too slow for one GPU: 2-D block-cyclic Cholesky of K = tt_to_cov(cov(X; theta)) for the data of
 * g3_set_data (replicated on every rank) on a Pr x Pc process grid, Pr * Pc = communicator size (1 x 1 without a
 * communicator).  Block size nb (multiple of 128 dividing N); rank r = q Pr + p holds, for every panel J = q (mod Pc),
 * the blocks (I, J), I >= J, I = p (mod Pr).  Semantics of the factor: CholeskyRobust.perform without the jitter ladder
 * (g3py/libs/tensors.py:197-201); *info = 1-based index of the first non-positive pivot, 0 if none.  The pieces stay on
 * the devices for g3_dist_solve / g3_dist_residual until g3_dist_free or the next g3_dist_factor.
 * Times are device times (CUDA events), max over ranks.  All ranks must make the same calls in the same order. */
enum { G3_DIST_NO_LOOKAHEAD = 1,   /* factor panel J+1 only after the whole trailing update of panel J (for comparison) */
       G3_DIST_RING2 = 2,          /* (default now) two panel buffers */
       G3_DIST_RING3 = 4 };        /* three panel buffers: ranks may drift two steps apart (measured slower) */
int g3_dist_factor(g3_ctx* ctx, const g3_kernel_desc* desc, const double* theta, int nb, int Pr, int Pc, int flags,
                   double* logdet /* sum log diag L */, int* info, float* ms_gram, float* ms_potrf, double* local_gib);
/* u = L^-1 delta (delta: N host doubles, read on rank 0), beta = u'u - the quadratic form of logp_cho
 * (g3py/processes/gaussian.py:212-215); u_out_or_NULL receives u (N) on every rank. */
int g3_dist_solve(g3_ctx* ctx, const double* delta, double* beta, double* u_out_or_NULL, float* ms);
/* Posterior moments at M test points (Xs: M x D host doubles; flags: G3_POST_NOISE) from the distributed factor:
 * mean_out = K* K^-1 delta (caller adds m(X*)), var_out = max(diag(K** - K* K^-1 K*'), 0) - g3_gp_posterior's semantics
 * (elliptical.py:78-107) without gathering L.  Needs g3_dist_factor + g3_dist_solve on a 1 x G grid; results on every rank. */
int g3_dist_posterior(g3_ctx* ctx, const double* Xs, int M, int flags, double* mean_out, double* var_out);
/* Gradient from the distributed factor (consumes it: L is overwritten by L^-1; 1 x G grid):
 *   dtheta[p] = 1/2 sum_ij (cfac alpha_i alpha_j - K^-1_ij) dK_ij/dtheta_p (natural space), ddelta = -cfac alpha, alpha = K^-1 delta
 * i.e. the dtheta / ddelta of g3_gp_logp_grad; cfac = 1 (gauss) or (nu + N) / (nu - 2 + beta) (student).  K^-1 is formed block
 * pair by block pair and contracted at once, never stored.  ms3 = device times {alpha, inverse, contraction}, max over ranks. */
int g3_dist_grad(g3_ctx* ctx, double cfac, double* dtheta, double* ddelta_or_NULL, float* ms3);
/* Correctness probe on the hardware: nvec (<= 4) seeded +-1 vectors v, rel_err[k] = max|L (L^T v) - K v| / max|K v| with
 * K regenerated from X (never stored). */
int g3_dist_residual(g3_ctx* ctx, int nvec, unsigned seed, double* rel_err);
/* Test accessor: this rank's piece of panel J (count * nb x nb row-major) -> host (may be NULL to query *count). */
int g3_dist_read_piece(g3_ctx* ctx, int J, double* host_or_NULL, int* count);
int g3_dist_free(g3_ctx* ctx);
/* Layout arithmetic only (no device): out4 = {owner rank of block (I, J), first block row of piece (J, p),
 * blocks in piece (J, p), index of block (I, J) inside its piece}. */
int g3_dist_layout(int N, int nb, int Pr, int Pc, int I, int J, int p, int* out4);
/* factor + solve in one call (ms3 = {gram, potrf, solve}). */
int g3_potrf_2d(g3_ctx* ctx, const g3_kernel_desc* desc, const double* theta, int nb, int Pr, int Pc, int flags,
                const double* delta_or_NULL, double* logdet, double* beta_or_NULL, int* info, float* ms3);

/* Device-pointer building blocks (kept for callers that own device memory themselves; work is issued on the context's
 * stream, which g3_set_stream can point at the caller's stream).
 *   g3_dev_gram_block : K[row0:row0+rows, col0:col0+cols] of cov(X) for the resident X (tt_to_cov
 *                       shift `diag_shift` and Noise on the global diagonal), into out (leading dim ld)
 *   g3_dev_potrf_panel: P (rows x nb, ld = nb): Cholesky of the top nb x nb block, rows below solved;
 *                       sum(log diag) is ADDED to *logdet_dev, first bad pivot (+1) stored in *info_dev
 *                       (the 128x128 block inverses go to Dinv_dev if given: needed later by g3_dev_trsv_panel)
 *   g3_dev_syrk_panel : D[x][y] -= sum_k P[row_off+x][k] * P[row_off+y][k]   (D rowsD x nb, ld = nb) */
int g3_set_stream(g3_ctx* ctx, void* cuda_stream_or_NULL);
int g3_dev_gram_block(g3_ctx* ctx, const g3_kernel_desc* desc, const double* theta, int row0, int col0,
                      int rows, int cols, double diag_shift, double* out_dev, long long ld);
int g3_dev_potrf_panel(g3_ctx* ctx, double* P_dev, int rows, int nb, double* Dinv_dev_or_NULL, double* logdet_dev,
                       int* info_dev);
/* forward substitution with one factored panel: u = L_top^-1 r[0:nb]; r[nb:rows] -= L_below u; *beta_dev += |u|^2.
 * Dinv_dev = the (nb/128) x 128 x 128 block inverses g3_dev_potrf_panel wrote for this panel. */
int g3_dev_trsv_panel(g3_ctx* ctx, const double* P_dev, int rows, int nb, const double* Dinv_dev, double* r_dev,
                      double* u_dev, double* beta_dev);
int g3_dev_syrk_panel(g3_ctx* ctx, const double* P_dev, int rowsP, int nb, int row_off, double* D_dev, int rowsD);

/* ---- stand-alone Cholesky benchmark entry (BASELINE metric 2) --------------------------
 * Builds K = cov(X; theta) (+ tt_to_cov) for the resident data directly in device memory
 * (lower triangle only) and factors it in place; nothing N x N crosses the host boundary.
 * Returns logdet and info; ms (may be NULL) = device time of the factorisation alone. */
int g3_gram_potrf_device(g3_ctx* ctx, const g3_kernel_desc* desc, const double* theta,
                         double* logdet, int* info, float* ms_gram, float* ms_potrf);

#ifdef __cplusplus
}
#endif
#endif /* G3B_H */

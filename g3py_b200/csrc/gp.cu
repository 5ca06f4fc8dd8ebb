// C-ABI entry points of libg3b.so: Gram / VJP / robust Cholesky on host arrays, the fused
// marginal-likelihood + gradient pipeline, posterior moments, and the big-matrix Cholesky.
//
// Pipeline of g3_gp_run for B hyper samples (everything on ctx->stream, no host sync inside):
//   gram_diag  -> tt_to_cov shift            (g3py/libs/tensors.py:95-98)
//   gram_fwd   -> K_b lower tiles, padded    (kernels.py:106-110, elliptical.py:71)
//   potrf      -> L_b, Dinv, logdet, info    (tensors.py:197-201)
//   trsv_fwd   -> u = L^-1 delta, beta       (gaussian.py:212-215, studentT.py:118-119)
//   [grad] trsv_bwd -> alpha; trtri -> U = L^-T; lauum -> K^-1 (over L); gram_vjp with
//          W = c alpha alpha' - K^-1 -> dtheta; ddelta = -c alpha          
// The jitter ladder (tensors.py:203-213) runs in g3_gp_download only for items whose info != 0.
#include "g3b_internal.cuh"
#include <math.h>
#include <string.h>

namespace {

constexpr int TS = G3_TILE;

__global__ void gp_prep_kernel(const double* __restrict__ dmin, double jitter, double* __restrict__ shift,
                               double* __restrict__ beta, double* __restrict__ logdet, int* __restrict__ info,
                               int* __restrict__ status, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double m = dmin[b];
  int st = status[b];
  if (m > 0.0) {
    shift[b] = 0.0;
  } else {  // tt_to_cov: r + (1e-6 - m) * eye
    shift[b] = jitter - m;
    st |= G3_ST_DIAG_SHIFT;
  }
  status[b] = st;
  beta[b] = 0.0;
  logdet[b] = 0.0;
  info[b] = 0;
}

// r[b][i] = delta[b*stride + i] (i < N), 0 on the padding.  Non-finite delta -> status.
__global__ void gp_expand_delta_kernel(const double* __restrict__ delta, int stride, int N, int Np,
                                       double* __restrict__ r, int* __restrict__ status) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Np) return;
  double v = 0.0;
  if (i < N) {
    v = delta[(long long)b * stride + i];
    if (!isfinite(v)) atomicOr(status + b, G3_ST_NONFINITE_RESULT);
  }
  r[(long long)b * Np + i] = v;
}

__global__ void gp_zero_sub_kernel(double* __restrict__ beta, double* __restrict__ logdet, int* __restrict__ info,
                                   const int* __restrict__ bmap, int nb) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nb) return;
  const int b = bmap[k];
  if (beta) beta[b] = 0.0;
  logdet[b] = 0.0;
  info[b] = 0;
}

__global__ void gp_clear_bits_kernel(int* __restrict__ status, int mask, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) status[b] &= ~mask;
}

__global__ void gp_zero_beta_kernel(double* __restrict__ beta, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) beta[b] = 0.0;
}

// c_b = 1 (gauss) or (nu+N)/(nu-2+beta) (student).  Also the final non-finite check.
__global__ void gp_cfac_kernel(int kind, const double* __restrict__ nu, const double* __restrict__ beta,
                               const double* __restrict__ logdet, double n, double* __restrict__ cfac,
                               int* __restrict__ status, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  cfac[b] = (kind == G3_KIND_STUDENT) ? (nu[b] + n) / (nu[b] - 2.0 + beta[b]) : 1.0;
  if (!isfinite(beta[b]) || !isfinite(logdet[b])) status[b] |= G3_ST_NONFINITE_RESULT;
}

__global__ void gp_ddelta_kernel(const double* __restrict__ alpha, const double* __restrict__ cfac, int N, int Np,
                                 double* __restrict__ ddelta) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) ddelta[(long long)b * N + i] = -cfac[b] * alpha[(long long)b * Np + i];
}

// A (n x n, ld) += shift on the diagonal; padding of the Np x Np device matrix set to identity.
__global__ void mat_pad_shift_kernel(double* __restrict__ A, int n, int Np, long long strideA,
                                     const double* __restrict__ shift, const int* __restrict__ bmap) {
  const int b = bmap ? bmap[blockIdx.y] : (int)blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Np) return;
  double* d = A + (long long)b * strideA + (long long)i * Np + i;
  if (i < n) {
    if (shift) *d += shift[b];
  } else {
    *d = 1.0;
  }
}

// zero the strictly-upper 128x128 tiles (the factor leaves K's upper tiles untouched).
__global__ void __launch_bounds__(256)
zero_upper_tiles_kernel(double* __restrict__ A, int Np, long long strideA, int T) {
  const int b = blockIdx.y;
  int tile = blockIdx.x;  // enumerates (x, y), x < y
  int y = (int)((sqrt(8.0 * (double)tile + 1.0) + 1.0) * 0.5);
  while ((long long)y * (y - 1) / 2 > tile) --y;
  while ((long long)(y + 1) * y / 2 <= tile) ++y;
  const int x = tile - (int)((long long)y * (y - 1) / 2);
  double* At = A + (long long)b * strideA + (long long)x * TS * Np + (long long)y * TS;
  for (int idx = threadIdx.x; idx < TS * TS / 2; idx += 256) {
    const int r = idx >> 6, c = (idx & 63) * 2;
    *reinterpret_cast<double2*>(At + (long long)r * Np + c) = make_double2(0.0, 0.0);
  }
}

// kss[0] = min diag of cov(space); kss[2] = tt_to_cov shift (only the noisy selector applies tt_to_cov).
__global__ void post_kss_shift_kernel(double* kss, double jitter, int apply) {
  if (threadIdx.x == 0) kss[2] = (apply && !(kss[0] > 0.0)) ? jitter - kss[0] : 0.0;
}

// mean[m] = sum_n Vt[m][n] u[n];  var[m] = max(k(x*_m, x*_m) - sum_n Vt[m][n]^2, 0).  One warp per row.
__global__ void __launch_bounds__(256)
post_moments_kernel(const double* __restrict__ Vt, int Np, int M, const double* __restrict__ u,
                    const double* __restrict__ kss_vec, const double* __restrict__ kss_shift,
                    double* __restrict__ mean, double* __restrict__ var) {
  const int m = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (m >= M) return;
  const double* row = Vt + (long long)m * Np;
  double s1 = 0.0, s2 = 0.0;
  for (int n = lane * 2; n < Np; n += 64) {
    const double2 v = *reinterpret_cast<const double2*>(row + n);
    const double2 w = *reinterpret_cast<const double2*>(u + n);
    s1 += v.x * w.x + v.y * w.y;
    s2 += v.x * v.x + v.y * v.y;
  }
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if (lane == 0) {
    mean[m] = s1;
    const double v = (kss_vec[m] + kss_shift[0]) - s2;
    var[m] = v < 0.0 ? 0.0 : v;   // tt_to_bounded(extract_diag(..), 0)  elliptical.py:94-97
  }
}

// debug: count lower-triangle mismatches per 128x128 tile between two factor buffers
__global__ void __launch_bounds__(256)
dbg_tile_mismatch_kernel(const double* __restrict__ A, const double* __restrict__ R, int Np, int T, int* __restrict__ out) {
  const int b = blockIdx.z, ti = blockIdx.y, tj = blockIdx.x;
  if (tj > ti) return;
  __shared__ int cnt;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  const long long base = (long long)b * Np * Np + (long long)ti * TS * Np + (long long)tj * TS;
  int c = 0;
  for (int idx = threadIdx.x; idx < TS * TS; idx += 256) {
    const int r = idx >> 7, q = idx & 127;
    if (ti == tj && q > r) continue;
    const double x = A[base + (long long)r * Np + q], y = R[base + (long long)r * Np + q];
    if (!(x == y)) ++c;
  }
  if (c) atomicAdd(&cnt, c);
  __syncthreads();
  if (threadIdx.x == 0) out[((long long)b * T + ti) * T + tj] = cnt;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
static GemmArgs gz() {
  GemmArgs g;
  memset(&g, 0, sizeof g);
  return g;
}

struct GpBufs {
  double *theta, *delta, *nu, *A, *U, *Dinv, *r, *u, *s, *alpha, *beta, *logdet, *shift, *dmin, *dmean, *cfac, *dtheta,
      *ddelta, *shift2;
  int *info, *status, *bmap;
  double* partials;            // gram_vjp per-tile partial sums, partials_per_item doubles per batch item
  size_t partials_per_item;
};

static int gp_alloc(g3_ctx* ctx, GpBufs& w, int B, int P, int N, int want_grad, int delta_rows) {
  const int Np = g3_pad(N), T = Np / TS;
  const size_t mat = sizeof(double) * (size_t)B * Np * Np;
#define WS(field, name, bytes)                          \
  w.field = (decltype(w.field))g3_ws(ctx, name, bytes); \
  if (!w.field) return -2;
  WS(theta, "gp_theta", sizeof(double) * (size_t)B * (P > 0 ? P : 1));
  WS(delta, "gp_delta", sizeof(double) * (size_t)delta_rows * N);
  WS(nu, "gp_nu", sizeof(double) * B);
  WS(A, "gp_A", mat);
  WS(Dinv, "gp_Dinv", sizeof(double) * (size_t)B * T * TS * TS);
  WS(r, "gp_r", sizeof(double) * (size_t)B * Np);
  WS(u, "gp_u", sizeof(double) * (size_t)B * Np);
  WS(s, "gp_s", sizeof(double) * (size_t)B * Np);
  WS(alpha, "gp_alpha", sizeof(double) * (size_t)B * Np);
  WS(beta, "gp_beta", sizeof(double) * B);
  WS(logdet, "gp_logdet", sizeof(double) * B);
  WS(shift, "gp_shift", sizeof(double) * B);
  WS(shift2, "gp_shift2", sizeof(double) * B);
  WS(dmin, "gp_dmin", sizeof(double) * B);
  WS(dmean, "gp_dmean", sizeof(double) * B);
  WS(cfac, "gp_cfac", sizeof(double) * B);
  WS(dtheta, "gp_dtheta", sizeof(double) * (size_t)B * (P > 0 ? P : 1));
  WS(ddelta, "gp_ddelta", sizeof(double) * (size_t)B * N);
  WS(info, "gp_info", sizeof(int) * B);
  WS(status, "gp_status", sizeof(int) * B);
  WS(bmap, "gp_bmap", sizeof(int) * B);
  w.partials = nullptr;
  w.partials_per_item = 0;
  if (want_grad || (ctx->speculate_grad && B <= 8)) {
    WS(U, "gp_U", mat);
    w.partials_per_item = (size_t)T * (T + 1) / 2 * (P > 0 ? P : 1);
    WS(partials, "gp_vjp_partials", sizeof(double) * (size_t)B * w.partials_per_item);
  } else {
    w.U = nullptr;
  }
#undef WS
  return 0;
}

// Stages after the factorisation: solves, beta, and (optionally) the gradient.
// View of the workspaces for the batch items [b0, b0 + ...): every per-item pointer advanced by b0 items.
static GpBufs gp_view(const GpBufs& w, int b0, int N, int P, int delta_stride) {
  const int Np = g3_pad(N), T = Np / TS;
  GpBufs v = w;
  const size_t m = (size_t)Np * Np;
  v.theta += (size_t)b0 * P;
  if (delta_stride) v.delta += (size_t)b0 * N;
  v.nu += b0;
  v.A += m * b0;
  if (v.U) v.U += m * b0;
  v.Dinv += (size_t)b0 * T * TS * TS;
  v.r += (size_t)b0 * Np; v.u += (size_t)b0 * Np; v.s += (size_t)b0 * Np; v.alpha += (size_t)b0 * Np;
  v.beta += b0; v.logdet += b0; v.shift += b0; v.shift2 += b0; v.dmin += b0; v.dmean += b0; v.cfac += b0;
  v.dtheta += (size_t)b0 * P;
  v.ddelta += (size_t)b0 * N;
  v.info += b0; v.status += b0;
  if (v.partials) v.partials += (size_t)b0 * v.partials_per_item;
  return v;
}

static int gp_grad_stage(g3_ctx* ctx, GpBufs& w, int B);

static int gp_after_potrf(g3_ctx* ctx, GpBufs& w, int B) {
  G3_NVTX("g3:solve+beta");
  const g3_gp_state& st = ctx->gp;
  const int N = ctx->N, Np = g3_pad(N);
  int rc;
  gp_zero_beta_kernel<<<(B + 127) / 128, 128, 0, ctx->stream>>>(w.beta, B);
  G3_LAUNCH_CHECK(ctx);
  gp_expand_delta_kernel<<<dim3((Np + 255) / 256, B), 256, 0, ctx->stream>>>(w.delta, st.delta_stride, N, Np, w.r,
                                                                             w.status);
  G3_LAUNCH_CHECK(ctx);
  if ((rc = g3_trsv_fwd(ctx, w.A, w.Dinv, w.r, w.u, w.beta, Np, B))) return rc;
  gp_cfac_kernel<<<(B + 127) / 128, 128, 0, ctx->stream>>>(st.kind, w.nu, w.beta, w.logdet, (double)N, w.cfac,
                                                           w.status, B);
  G3_LAUNCH_CHECK(ctx);
  if (!st.want_grad) return 0;
  return gp_grad_stage(ctx, w, B);
}

// Gradient stages on a resident factor: alpha = L^-T u, U = L^-T, K^-1 = U U^T (over L), W-contraction, d/d delta.
static int gp_grad_stage(g3_ctx* ctx, GpBufs& w, int B) {
  G3_NVTX("g3:gradient(alpha,trtri,lauum,vjp)");
  const g3_gp_state& st = ctx->gp;
  const int N = ctx->N, Np = g3_pad(N), T = Np / TS, P = st.desc.n_theta;
  int rc;
  G3_CUDA(ctx, cudaMemcpyAsync(w.s, w.u, sizeof(double) * (size_t)B * Np, cudaMemcpyDeviceToDevice, ctx->stream));
  if ((rc = g3_trsv_bwd(ctx, w.A, w.Dinv, w.s, w.alpha, Np, B))) return rc;
  if (!ctx->trtri_done && (rc = g3_trtri_batched(ctx, w.A, w.U, Np, B, w.Dinv))) return rc;   // else: pipelined behind potrf
  ctx->trtri_done = 0;
  if ((rc = g3_lauum_batched(ctx, w.U, w.A, Np, B))) return rc;  // K^-1 (lower tiles) overwrites L
  VjpArgs v;
  memset(&v, 0, sizeof v);
  v.X1 = ctx->dX; v.X2 = ctx->dX; v.n1 = N; v.n2 = N; v.D = ctx->D; v.same = 1; v.lower_only = 1;
  v.theta = w.theta; v.P = P;
  v.W = w.A; v.ldw = Np; v.strideW = (long long)Np * Np;
  v.alpha = w.alpha; v.strideAlpha = Np; v.cfac = w.cfac;
  v.scale = 0.5;
  v.dtheta = w.dtheta;
  v.partials = w.partials;
  if ((rc = g3_gram_vjp_launch(ctx, st.desc, v, B))) return rc;
  gp_ddelta_kernel<<<dim3((N + 255) / 256, B), 256, 0, ctx->stream>>>(w.alpha, w.cfac, N, Np, w.ddelta);
  G3_LAUNCH_CHECK(ctx);
  (void)T;
  return 0;
}

static int gp_build_and_factor(g3_ctx* ctx, GpBufs& w, int B, const double* shift, const int* bmap, int nb) {
  G3_NVTX("g3:gram+potrf");
  const g3_gp_state& st = ctx->gp;
  const int N = ctx->N, Np = g3_pad(N);
  GramArgs a;
  memset(&a, 0, sizeof a);
  a.X1 = ctx->dX; a.X2 = ctx->dX; a.n1 = N; a.n2 = N; a.D = ctx->D;
  a.same = 1; a.lower_only = 1; a.pad_identity = 1;
  a.theta = w.theta; a.P = st.desc.n_theta;
  a.diag_shift = shift;
  a.K = w.A; a.ldk = Np; a.strideK = (long long)Np * Np;
  a.status = w.status; a.bmap = bmap;
  int rc;
  if ((rc = g3_gram_launch(ctx, st.desc, a, bmap ? nb : B))) return rc;
  // U = L^-T pipelined behind the factorisation: when the gradient follows in this call, or (g3_set_speculate_grad) when the
  // caller announced that it will ask for it right after a logp-only evaluation (g3_gp_grad_resume)
  const bool spec = ctx->speculate_grad && !st.want_grad && B <= 8;
  return g3_potrf_batched(ctx, w.A, Np, B, w.Dinv, w.logdet, w.info, bmap, nb, ctx->potrf_w,
                          ((st.want_grad || spec) && !bmap) ? w.U : nullptr);
}

extern "C" {

int g3_set_potrf_block(g3_ctx* ctx, int w_outer) {
  ctx->potrf_w = w_outer;
  return 0;
}

int g3_set_lookahead(g3_ctx* ctx, int on) {
  ctx->lookahead = on ? 1 : 0;
  return 0;
}

int g3_set_trtri_pipeline(g3_ctx* ctx, int on) {
  ctx->trtri_pipeline = on ? 1 : 0;
  return 0;
}

int g3_set_speculate_grad(g3_ctx* ctx, int on) {
  ctx->speculate_grad = on ? 1 : 0;        // part of the graph key: no need to drop the cached graph
  return 0;
}

int g3_set_trsv_fused(g3_ctx* ctx, int on) {
  if (ctx->trsv_fused != (on ? 1 : 0)) g3_graph_drop(ctx);
  ctx->trsv_fused = on ? 1 : 0;
  return 0;
}

int g3_set_tile_split(g3_ctx* ctx, int on) {
  if (ctx->tile_split != (on ? 1 : 0)) g3_graph_drop(ctx);
  ctx->tile_split = on ? 1 : 0;
  return 0;
}

int g3_set_splitk(g3_ctx* ctx, int on) {
  if (ctx->splitk != on) g3_graph_drop(ctx);
  ctx->splitk = on < 0 ? 0 : (on > 2 ? 2 : on);
  return 0;
}

int g3_set_gemm_mode(g3_ctx* ctx, int mode, int min_k) {
  if (mode != G3_GEMM_DMMA && mode != G3_GEMM_OZAKI) return g3_fail_msg(ctx, "g3_set_gemm_mode: unknown mode");
  ctx->gemm_mode = mode;
  if (min_k > 0) ctx->oz_min_k = (min_k + 127) / 128 * 128;
  return 0;
}

int64_t g3_ozaki_launch_count(g3_ctx* ctx) { return ctx->oz_launches; }

int g3_set_groups(g3_ctx* ctx, int n_groups) {
  ctx->n_groups = n_groups < 1 ? 1 : (n_groups > G3_MAX_GROUPS ? G3_MAX_GROUPS : n_groups);
  return 0;
}

int g3_gp_upload(g3_ctx* ctx, const g3_kernel_desc* desc, int kind, const double* delta, int delta_stride,
                 const double* theta, int B, const double* nu_or_NULL, int want_grad) {
  if (!ctx || !desc || !delta || B <= 0) return g3_fail_msg(ctx, "g3_gp_upload: bad arguments");
  if (!ctx->dX) return g3_fail_msg(ctx, "g3_gp_upload: call g3_set_data first");
  if (desc->n_theta > 0 && !theta) return g3_fail_msg(ctx, "g3_gp_upload: theta is NULL");
  if (kind == G3_KIND_STUDENT && !nu_or_NULL) return g3_fail_msg(ctx, "g3_gp_upload: student kind needs nu");
  if (delta_stride != 0 && delta_stride != ctx->N) return g3_fail_msg(ctx, "g3_gp_upload: delta_stride must be 0 or N");
  G3_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = g3_check_desc(ctx, *desc, ctx->D);
  if (rc) return rc;
  const int N = ctx->N, P = desc->n_theta;
  GpBufs w;
  const int drows = delta_stride ? B : 1;
  if ((rc = gp_alloc(ctx, w, B, P, N, want_grad, drows))) return rc;
  ctx->gp.factor_resident = 0;
  ctx->gp.trtri_spec = 0;
  ctx->gp.desc = *desc;
  ctx->gp.kind = kind;
  ctx->gp.B = B;
  ctx->gp.want_grad = want_grad;
  ctx->gp.delta_stride = delta_stride;
  // stage through page-locked memory so that the H2D copies are true async DMA (the caller's arrays are pageable)
  const size_t n_th = (size_t)B * P, n_dl = (size_t)drows * N, n_nu = nu_or_NULL ? (size_t)B : 0;
  double* stage = (double*)g3_pinned(ctx, "gp_h2d", sizeof(double) * (n_th + n_dl + n_nu + 1));
  if (!stage) return -2;
  G3_CUDA(ctx, cudaEventSynchronize(ctx->ev_h2d));          // previous upload has left the staging buffer
  if (n_th) memcpy(stage, theta, sizeof(double) * n_th);
  memcpy(stage + n_th, delta, sizeof(double) * n_dl);
  if (n_nu) memcpy(stage + n_th + n_dl, nu_or_NULL, sizeof(double) * n_nu);
  if (n_th) G3_CUDA(ctx, cudaMemcpyAsync(w.theta, stage, sizeof(double) * n_th, cudaMemcpyHostToDevice, ctx->stream));
  G3_CUDA(ctx, cudaMemcpyAsync(w.delta, stage + n_th, sizeof(double) * n_dl, cudaMemcpyHostToDevice, ctx->stream));
  if (n_nu)
    G3_CUDA(ctx, cudaMemcpyAsync(w.nu, stage + n_th + n_dl, sizeof(double) * n_nu, cudaMemcpyHostToDevice, ctx->stream));
  G3_CUDA(ctx, cudaEventRecord(ctx->ev_h2d, ctx->stream));
  ctx->gp.valid = 1;
  return 0;
}

}  // extern "C"

static int gp_run_body(g3_ctx* ctx);

// The launch sequence of one evaluation of up to 8 matrices (one MCMC chain / a BFGS step: ~150 dependent launches on
// three streams at N=2048) depends only on (kernel tree, kind, B, N, gradient or not, schedule switches, workspace
// addresses).  The second call with the same key captures it into a CUDA graph (stream capture: the look-ahead and
// pipelined-trtri streams fork from and join the context's stream through the events the schedule already uses);
// later calls replay the graph - one cudaGraphLaunch instead of ~150 launches, event records and tensor-map encodes.
static std::string gp_graph_key(const g3_ctx* ctx) {
  const g3_gp_state& st = ctx->gp;
  std::string k((const char*)&st.desc, sizeof st.desc);
  long long v[] = {st.kind, st.B, st.want_grad, st.delta_stride, ctx->N, ctx->D, ctx->potrf_w, ctx->lookahead, ctx->splitk,
                   ctx->speculate_grad, ctx->tile_split, ctx->trsv_fused, ctx->diag_variant, ctx->trtri_pipeline, ctx->gemm_mode, ctx->oz_min_k, ctx->n_groups, (long long)ctx->ws_gen,
                   (long long)(intptr_t)ctx->dX, (long long)(intptr_t)ctx->stream};
  k.append((const char*)v, sizeof v);
  k.append((const char*)&ctx->jitter_rel, sizeof(double));
  return k;
}

void g3_graph_drop(g3_ctx* ctx) {
  if (ctx->graph_exec) cudaGraphExecDestroy((cudaGraphExec_t)ctx->graph_exec);
  ctx->graph_exec = nullptr;
  ctx->graph_key.clear();
  ctx->graph_warm_key.clear();
}

extern "C" {

int64_t g3_graph_replays(g3_ctx* ctx) { return ctx->graph_replays; }

int g3_set_graphs(g3_ctx* ctx, int on) {
  ctx->graphs_on = on ? 1 : 0;
  if (!on) g3_graph_drop(ctx);
  return 0;
}

int g3_gp_run(g3_ctx* ctx) {
  G3_NVTX("g3_gp_run");
  if (!ctx || !ctx->gp.valid) return g3_fail_msg(ctx, "g3_gp_run: nothing uploaded");
  G3_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!ctx->graphs_on || ctx->gp.B > 8 || ctx->prof_on) return gp_run_body(ctx);
  const std::string key = gp_graph_key(ctx);
  if (ctx->graph_exec && key == ctx->graph_key) {               // replay
    G3_CUDA(ctx, cudaGraphLaunch((cudaGraphExec_t)ctx->graph_exec, ctx->stream));
    ctx->launches += ctx->graph_launches;
    ctx->graph_replays++;
    ctx->gp.trtri_spec = ctx->graph_trtri_spec;
    return 0;
  }
  if (key != ctx->graph_warm_key) {                             // first call with this key: plain run (allocations, attributes)
    const int rc = gp_run_body(ctx);
    ctx->graph_warm_key = rc ? std::string() : gp_graph_key(ctx);   // the run may have (re)allocated workspaces
    return rc;
  }
  g3_graph_drop(ctx);                                           // second call: capture, instantiate, launch
  const int64_t l0 = ctx->launches;
  cudaStream_t s = ctx->stream;
  G3_CUDA(ctx, cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
  int rc = gp_run_body(ctx);
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamEndCapture(s, &graph);
  ctx->stream = s;
  if (rc || e != cudaSuccess || !graph || gp_graph_key(ctx) != key) {   // could not capture (or buffers moved): run it plainly
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    ctx->launches = l0;
    ctx->graph_warm_key.clear();
    return rc ? rc : gp_run_body(ctx);
  }
  cudaGraphExec_t exec = nullptr;
  e = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess || !exec) {
    cudaGetLastError();
    ctx->launches = l0;
    ctx->graph_warm_key.clear();
    return gp_run_body(ctx);
  }
  ctx->graph_exec = exec;
  ctx->graph_key = key;
  ctx->graph_launches = ctx->launches - l0;
  ctx->graph_trtri_spec = ctx->gp.trtri_spec;
  G3_CUDA(ctx, cudaGraphLaunch(exec, s));
  return 0;
}

}  // extern "C"

static int gp_run_body(g3_ctx* ctx) {
  const g3_gp_state& st = ctx->gp;
  const int B = st.B, N = ctx->N, P = st.desc.n_theta;
  GpBufs w;
  int rc;
  if ((rc = gp_alloc(ctx, w, B, P, N, st.want_grad, st.delta_stride ? B : 1))) return rc;
  G3_CUDA(ctx, cudaMemsetAsync(w.status, 0, sizeof(int) * B, ctx->stream));
  if ((rc = g3_gram_diag_min(ctx, st.desc, ctx->dX, N, ctx->D, w.theta, P, B, w.dmin, w.dmean, w.status, 0))) return rc;
  gp_prep_kernel<<<(B + 127) / 128, 128, 0, ctx->stream>>>(w.dmin, ctx->jitter_rel, w.shift, w.beta, w.logdet, w.info,
                                                           w.status, B);
  G3_LAUNCH_CHECK(ctx);
  // Independent batch items are processed in groups on separate streams: the serial per-column chain of one
  // group (update GEMM -> diagonal kernel -> solve GEMM) leaves SMs idle in its kernel tails and during the
  // 16..64-CTA diagonal kernels; the other groups' kernels fill them.
  int groups = ctx->n_groups < 1 ? 1 : ctx->n_groups;
  if (groups > G3_MAX_GROUPS) groups = G3_MAX_GROUPS;
  while (groups > 1 && B / groups < 8) --groups;
  ctx->force_left = B > 8;               // schedule chosen on the whole batch, not per stream group
  if (groups == 1) {
    rc = gp_build_and_factor(ctx, w, B, w.shift, nullptr, 0);
    if (!rc) rc = gp_after_potrf(ctx, w, B);
    if (!st.want_grad) {                   // speculative U (if any) belongs to the resident factor, not to the next call
      ctx->gp.trtri_spec = rc ? 0 : ctx->trtri_done;
      ctx->trtri_done = 0;
    }
    ctx->force_left = 0;
    return rc;
  }
  cudaStream_t main_stream = ctx->stream;
  G3_CUDA(ctx, cudaEventRecord(ctx->gev_start, main_stream));
  rc = 0;
  for (int g = 0; g < groups && !rc; ++g) {
    const int b0 = (int)((long long)B * g / groups), b1 = (int)((long long)B * (g + 1) / groups);
    GpBufs v = gp_view(w, b0, N, P, st.delta_stride);
    ctx->stream = ctx->gstream[g];
    cudaStreamWaitEvent(ctx->stream, ctx->gev_start, 0);
    rc = gp_build_and_factor(ctx, v, b1 - b0, v.shift, nullptr, 0);
    if (!rc) rc = gp_after_potrf(ctx, v, b1 - b0);
    cudaEventRecord(ctx->gev_done[g], ctx->stream);
    cudaStreamWaitEvent(main_stream, ctx->gev_done[g], 0);
  }
  ctx->stream = main_stream;
  ctx->force_left = 0;
  return rc;
}

extern "C" {

int g3_gp_download(g3_ctx* ctx, double* beta, double* logdet, double* dtheta_or_NULL, double* ddelta_or_NULL,
                   int* status) {
  G3_NVTX("g3_gp_download");
  if (!ctx || !ctx->gp.valid) return g3_fail_msg(ctx, "g3_gp_download: nothing uploaded");
  G3_CUDA(ctx, cudaSetDevice(ctx->device));
  const g3_gp_state& st = ctx->gp;
  const int B = st.B, N = ctx->N, P = st.desc.n_theta;
  GpBufs w;
  int rc;
  if ((rc = gp_alloc(ctx, w, B, P, N, st.want_grad, st.delta_stride ? B : 1))) return rc;
  std::vector<int> info(B), stat(B);
  G3_CUDA(ctx, cudaMemcpyAsync(info.data(), w.info, sizeof(int) * B, cudaMemcpyDeviceToHost, ctx->stream));
  G3_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  std::vector<int> failed;
  for (int b = 0; b < B; ++b)
    if (info[b] != 0) failed.push_back(b);
  std::vector<int> tries(B, 0), exhausted(B, 0);
  if (!failed.empty()) {
    ctx->gp.trtri_spec = 0;                // the ladder refactors: a speculative U of the first pass is stale
    // CholeskyRobust._cholesky ladder (tensors.py:203-213): dK = mean(diag K) * jitter, x10 per try.
    std::vector<double> dmean(B), shift(B), dK(B), sh2(B);
    G3_CUDA(ctx, cudaMemcpyAsync(dmean.data(), w.dmean, sizeof(double) * B, cudaMemcpyDeviceToHost, ctx->stream));
    G3_CUDA(ctx, cudaMemcpyAsync(shift.data(), w.shift, sizeof(double) * B, cudaMemcpyDeviceToHost, ctx->stream));
    G3_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int b : failed) dK[b] = (dmean[b] + shift[b]) * ctx->jitter_rel;
    for (int t = 1; t <= ctx->max_tries && !failed.empty(); ++t) {
      for (int b = 0; b < B; ++b) sh2[b] = shift[b] + dK[b];
      const int nb = (int)failed.size();
      G3_CUDA(ctx, cudaMemcpyAsync(w.shift2, sh2.data(), sizeof(double) * B, cudaMemcpyHostToDevice, ctx->stream));
      G3_CUDA(ctx, cudaMemcpyAsync(w.bmap, failed.data(), sizeof(int) * nb, cudaMemcpyHostToDevice, ctx->stream));
      gp_zero_sub_kernel<<<(nb + 127) / 128, 128, 0, ctx->stream>>>(nullptr, w.logdet, w.info, w.bmap, nb);
      G3_LAUNCH_CHECK(ctx);
      if ((rc = gp_build_and_factor(ctx, w, B, w.shift2, w.bmap, nb))) return rc;
      G3_CUDA(ctx, cudaMemcpyAsync(info.data(), w.info, sizeof(int) * B, cudaMemcpyDeviceToHost, ctx->stream));
      G3_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      std::vector<int> still;
      for (int b : failed) {
        if (info[b] != 0) {
          still.push_back(b);
          dK[b] *= 10.0;
        } else {
          tries[b] = t;
        }
      }
      failed.swap(still);
    }
    for (int b : failed) exhausted[b] = 1;
    // The items repaired by the ladder need their solves / gradient again; the stages are idempotent
    // given (L, delta), so they are simply re-run for the whole batch (rare path).  K^-1 overwrote L for
    // the items that had succeeded, so with gradients on the whole batch is rebuilt first.
    if (st.want_grad) {
      for (int b = 0; b < B; ++b) sh2[b] = shift[b] + (tries[b] ? dK[b] : 0.0);
      G3_CUDA(ctx, cudaMemcpyAsync(w.shift2, sh2.data(), sizeof(double) * B, cudaMemcpyHostToDevice, ctx->stream));
      G3_CUDA(ctx, cudaMemsetAsync(w.logdet, 0, sizeof(double) * B, ctx->stream));
      G3_CUDA(ctx, cudaMemsetAsync(w.info, 0, sizeof(int) * B, ctx->stream));
      if ((rc = gp_build_and_factor(ctx, w, B, w.shift2, nullptr, 0))) return rc;
    }
    // the first pass flagged the failed items' NaN beta/logdet; the repaired values are re-checked
    gp_clear_bits_kernel<<<(B + 127) / 128, 128, 0, ctx->stream>>>(w.status, G3_ST_NONFINITE_RESULT, B);
    G3_LAUNCH_CHECK(ctx);
    if ((rc = gp_after_potrf(ctx, w, B))) return rc;
  }
  // results come back through page-locked staging (async DMA), then one host memcpy into the caller's arrays
  const bool get_th = st.want_grad && dtheta_or_NULL && P > 0, get_dl = st.want_grad && ddelta_or_NULL;
  const size_t o_beta = 0, o_ld = (size_t)B, o_th = 2 * (size_t)B, o_dl = o_th + (get_th ? (size_t)B * P : 0);
  const size_t o_st = o_dl + (get_dl ? (size_t)B * N : 0);
  double* out = (double*)g3_pinned(ctx, "gp_d2h", sizeof(double) * (o_st + (size_t)B + 1));
  if (!out) return -2;
  G3_CUDA(ctx, cudaMemcpyAsync(out + o_beta, w.beta, sizeof(double) * B, cudaMemcpyDeviceToHost, ctx->stream));
  G3_CUDA(ctx, cudaMemcpyAsync(out + o_ld, w.logdet, sizeof(double) * B, cudaMemcpyDeviceToHost, ctx->stream));
  G3_CUDA(ctx, cudaMemcpyAsync(out + o_st, w.status, sizeof(int) * B, cudaMemcpyDeviceToHost, ctx->stream));
  if (get_th)
    G3_CUDA(ctx, cudaMemcpyAsync(out + o_th, w.dtheta, sizeof(double) * (size_t)B * P, cudaMemcpyDeviceToHost, ctx->stream));
  if (get_dl)
    G3_CUDA(ctx, cudaMemcpyAsync(out + o_dl, w.ddelta, sizeof(double) * (size_t)B * N, cudaMemcpyDeviceToHost, ctx->stream));
  G3_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(beta, out + o_beta, sizeof(double) * B);
  memcpy(logdet, out + o_ld, sizeof(double) * B);
  memcpy(stat.data(), out + o_st, sizeof(int) * B);
  if (get_th) memcpy(dtheta_or_NULL, out + o_th, sizeof(double) * (size_t)B * P);
  if (get_dl) memcpy(ddelta_or_NULL, out + o_dl, sizeof(double) * (size_t)B * N);
  if (status) {
    for (int b = 0; b < B; ++b) {
      int s = stat[b];
      if (tries[b]) s |= G3_ST_JITTER | (tries[b] << 8);
      if (exhausted[b]) s |= G3_ST_POTRF_FAILED | G3_ST_JITTER | (ctx->max_tries << 8);
      status[b] = s;
    }
  }
  // logp-only evaluation: L, Dinv, u, c are final (ladder included) and stay in the workspaces, so the gradient of the
  // SAME evaluation can be finished later without refactoring (g3_gp_grad_resume)
  ctx->gp.factor_resident = st.want_grad ? 0 : 1;
  return 0;
}

int g3_gp_grad_resume(g3_ctx* ctx, double* dtheta, double* ddelta) {
  G3_NVTX("g3_gp_grad_resume");
  if (!ctx || !ctx->gp.valid || !ctx->gp.factor_resident || ctx->gp.want_grad)
    return g3_fail_msg(ctx, "g3_gp_grad_resume: no resident factor (call g3_gp_logp_grad without gradient outputs first)");
  G3_CUDA(ctx, cudaSetDevice(ctx->device));
  const g3_gp_state& st = ctx->gp;
  const int B = st.B, N = ctx->N, P = st.desc.n_theta;
  GpBufs w;
  int rc;
  if ((rc = gp_alloc(ctx, w, B, P, N, 1, st.delta_stride ? B : 1))) return rc;
  ctx->gp.factor_resident = 0;                     // K^-1 is about to overwrite L
  ctx->trtri_done = ctx->gp.trtri_spec;            // U already there (pipelined behind the logp-only factorisation)?
  ctx->gp.trtri_spec = 0;
  ctx->force_left = B > 8;
  rc = gp_grad_stage(ctx, w, B);
  ctx->force_left = 0;
  if (rc) return rc;
  const bool get_th = dtheta && P > 0, get_dl = ddelta != nullptr;
  double* out = (double*)g3_pinned(ctx, "gp_d2h", sizeof(double) * ((size_t)B * P + (size_t)B * N + 1));
  if (!out) return -2;
  if (get_th) G3_CUDA(ctx, cudaMemcpyAsync(out, w.dtheta, sizeof(double) * (size_t)B * P, cudaMemcpyDeviceToHost, ctx->stream));
  if (get_dl)
    G3_CUDA(ctx, cudaMemcpyAsync(out + (size_t)B * P, w.ddelta, sizeof(double) * (size_t)B * N, cudaMemcpyDeviceToHost, ctx->stream));
  G3_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (get_th) memcpy(dtheta, out, sizeof(double) * (size_t)B * P);
  if (get_dl) memcpy(ddelta, out + (size_t)B * P, sizeof(double) * (size_t)B * N);
  return 0;
}

int g3_gp_logp_grad(g3_ctx* ctx, const g3_kernel_desc* desc, int kind, const double* delta, int delta_stride,
                    const double* theta, int B, const double* nu_or_NULL, double* beta, double* logdet,
                    double* dtheta_or_NULL, double* ddelta_or_NULL, int* status) {
  if (!beta || !logdet) return g3_fail_msg(ctx, "g3_gp_logp_grad: beta/logdet outputs are required");
  const int want_grad = (dtheta_or_NULL || ddelta_or_NULL) ? 1 : 0;
  int rc;
  if ((rc = g3_gp_upload(ctx, desc, kind, delta, delta_stride, theta, B, nu_or_NULL, want_grad))) return rc;
  if ((rc = g3_gp_run(ctx))) return rc;
  return g3_gp_download(ctx, beta, logdet, dtheta_or_NULL, ddelta_or_NULL, status);
}

// Debug stress: build K once, then factor it `iters` times from a backup and compare every factor with the
// first one tile by tile.  out[it][0..3] = {#mismatching tiles, b, tile row, tile col of the first one}.
int g3_debug_potrf_stress(g3_ctx* ctx, const g3_kernel_desc* desc, const double* theta, int B, int iters, int* out,
                          double* tiles /* 2 x 128 x 128: bad tile, reference tile of the first mismatch */) {
  if (!ctx->dX) return g3_fail_msg(ctx, "stress: call g3_set_data first");
  G3_CUDA(ctx, cudaSetDevice(ctx->device));
  const int N = ctx->N, Np = g3_pad(N), T = Np / TS, P = desc->n_theta;
  GpBufs w;
  int rc;
  if ((rc = gp_alloc(ctx, w, B, P, N, 1, 1))) return rc;
  const size_t mat = sizeof(double) * (size_t)B * Np * Np;
  double* Kb = (double*)g3_ws(ctx, "dbg_K", mat);
  int* mm = (int*)g3_ws(ctx, "dbg_mm", sizeof(int) * (size_t)B * T * T);
  if (!Kb || !mm) return -2;
  ctx->gp.desc = *desc; ctx->gp.B = B; ctx->gp.kind = 0; ctx->gp.want_grad = 1; ctx->gp.delta_stride = 0; ctx->gp.valid = 0;
  G3_CUDA(ctx, cudaMemcpyAsync(w.theta, theta, sizeof(double) * (size_t)B * P, cudaMemcpyHostToDevice, ctx->stream));
  G3_CUDA(ctx, cudaMemsetAsync(w.status, 0, sizeof(int) * B, ctx->stream));
  G3_CUDA(ctx, cudaMemsetAsync(w.A, 0, mat, ctx->stream));
  GramArgs a;
  memset(&a, 0, sizeof a);
  a.X1 = ctx->dX; a.X2 = ctx->dX; a.n1 = N; a.n2 = N; a.D = ctx->D; a.same = 1; a.lower_only = 1; a.pad_identity = 1;
  a.theta = w.theta; a.P = P; a.K = w.A; a.ldk = Np; a.strideK = (long long)Np * Np; a.status = w.status;
  if ((rc = g3_gram_launch(ctx, *desc, a, B))) return rc;
  G3_CUDA(ctx, cudaMemcpyAsync(Kb, w.A, mat, cudaMemcpyDeviceToDevice, ctx->stream));
  std::vector<int> h((size_t)B * T * T);
  bool dumped = false;
  for (int it = 0; it < iters; ++it) {
    G3_CUDA(ctx, cudaMemcpyAsync(w.A, Kb, mat, cudaMemcpyDeviceToDevice, ctx->stream));
    G3_CUDA(ctx, cudaMemsetAsync(w.logdet, 0, sizeof(double) * B, ctx->stream));
    G3_CUDA(ctx, cudaMemsetAsync(w.info, 0, sizeof(int) * B, ctx->stream));
    if ((rc = g3_potrf_batched(ctx, w.A, Np, B, w.Dinv, w.logdet, w.info, nullptr, 0, ctx->potrf_w))) return rc;
    out[4 * it + 0] = out[4 * it + 1] = out[4 * it + 2] = out[4 * it + 3] = 0;
    if (it == 0) {
      G3_CUDA(ctx, cudaMemcpyAsync(w.U, w.A, mat, cudaMemcpyDeviceToDevice, ctx->stream));
      continue;
    }
    dbg_tile_mismatch_kernel<<<dim3(T, T, B), 256, 0, ctx->stream>>>(w.A, w.U, Np, T, mm);
    G3_LAUNCH_CHECK(ctx);
    G3_CUDA(ctx, cudaMemcpyAsync(h.data(), mm, sizeof(int) * h.size(), cudaMemcpyDeviceToHost, ctx->stream));
    G3_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int nbad = 0;
    for (int b = 0; b < B; ++b)
      for (int ti = 0; ti < T; ++ti)
        for (int tj = 0; tj <= ti; ++tj)
          if (h[((size_t)b * T + ti) * T + tj]) {
            if (!nbad) {
              out[4 * it + 1] = b; out[4 * it + 2] = ti; out[4 * it + 3] = tj;
              if (tiles && !dumped) {
                const size_t off = (size_t)b * Np * Np + (size_t)ti * TS * Np + (size_t)tj * TS;
                cudaMemcpy2D(tiles, sizeof(double) * TS, w.A + off, sizeof(double) * Np, sizeof(double) * TS, TS, cudaMemcpyDeviceToHost);
                cudaMemcpy2D(tiles + TS * TS, sizeof(double) * TS, w.U + off, sizeof(double) * Np, sizeof(double) * TS, TS, cudaMemcpyDeviceToHost);
                dumped = true;
              }
            }
            ++nbad;
          }
    out[4 * it + 0] = nbad;
  }
  return 0;
}

namespace {
__global__ void dbg_fill_kernel(double* p, size_t n, unsigned seed) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    unsigned long long z = (i + 1) * 0x9E3779B97F4A7C15ull + seed;
    z ^= z >> 31; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 29;
    p[i] = (double)(z >> 11) * (1.0 / 9007199254740992.0) - 0.5;
  }
}
__global__ void dbg_cmp_kernel(const double* a, const double* b, size_t n, unsigned long long* cnt, unsigned long long* first) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride)
    if (!(a[i] == b[i])) {
      atomicAdd(cnt, 1ull);
      atomicMin(first, (unsigned long long)i);
    }
}
}  // namespace

// GEMM determinism stress: D[rows x 128] = A[rows x 128] * Dm[128 x 128]^T for `B` batch items (the trsm-shaped
// launch of potrf), `launches` times; inplace=1 writes over A (restored from a backup before each launch).
// out[0] = launches with a mismatch vs the first launch, out[1] = mismatching elements in total,
// out[2] = first mismatching element index of the first bad launch, out[3] = index of the first bad launch.
int g3_debug_gemm_stress(g3_ctx* ctx, int rows, int B, int launches, int inplace, int kdepth, long long* out) {
  G3_CUDA(ctx, cudaSetDevice(ctx->device));
  const int Np = g3_pad(rows), T = Np / TS;
  const int KD = kdepth > 0 ? kdepth : TS;          // contraction depth (columns of A)
  const size_t nA = (size_t)B * Np * KD, nD = (size_t)B * TS * KD, nO = (size_t)B * Np * TS;
  double* A = (double*)g3_ws(ctx, "st_A", sizeof(double) * nA);
  double* A0 = (double*)g3_ws(ctx, "st_A0", sizeof(double) * nA);
  double* Dm = (double*)g3_ws(ctx, "st_D", sizeof(double) * nD);
  double* O = (double*)g3_ws(ctx, "st_O", sizeof(double) * nO);
  double* R = (double*)g3_ws(ctx, "st_R", sizeof(double) * nO);
  unsigned long long* cnt = (unsigned long long*)g3_ws(ctx, "st_cnt", 16);
  if (!A || !A0 || !Dm || !O || !R || !cnt) return -2;
  if (inplace && KD != TS) return g3_fail_msg(ctx, "stress: inplace needs kdepth 128");
  dbg_fill_kernel<<<1024, 256, 0, ctx->stream>>>(A0, nA, 1u);
  dbg_fill_kernel<<<1024, 256, 0, ctx->stream>>>(Dm, nD, 2u);
  CUtensorMap tmA, tmD;
  int rc;
  if ((rc = g3_make_tmap(ctx, &tmA, A, KD, Np, B, KD, (uint64_t)Np * KD, G3_BM))) return rc;
  if ((rc = g3_make_tmap(ctx, &tmD, Dm, KD, TS, B, KD, (uint64_t)TS * KD, G3_BN))) return rc;
  out[0] = out[1] = 0; out[2] = out[3] = -1;
  for (int it = 0; it < launches; ++it) {
    G3_CUDA(ctx, cudaMemcpyAsync(A, A0, sizeof(double) * nA, cudaMemcpyDeviceToDevice, ctx->stream));
    GemmArgs g = gz();
    g.D = inplace ? A : O; g.ldd = inplace ? KD : TS; g.strideD = inplace ? (long long)Np * KD : (long long)Np * TS;
    g.mode = 0; g.ntx = T; g.nty = 1;
    g.a_r0 = 0; g.a_rx = TS; g.ka0 = 0; g.b_r0 = 0; g.kb0 = 0; g.kl0 = KD;
    g.alpha = 1.0; g.beta = 0.0;
    if ((rc = g3_gemm_launch(ctx, tmA, tmD, g, B))) return rc;
    const double* res = inplace ? A : O;
    if (it == 0) {
      G3_CUDA(ctx, cudaMemcpyAsync(R, res, sizeof(double) * nO, cudaMemcpyDeviceToDevice, ctx->stream));
      continue;
    }
    G3_CUDA(ctx, cudaMemsetAsync(cnt, 0, 8, ctx->stream));
    G3_CUDA(ctx, cudaMemsetAsync(cnt + 1, 0xff, 8, ctx->stream));
    dbg_cmp_kernel<<<1024, 256, 0, ctx->stream>>>(res, R, nO, cnt, cnt + 1);
    unsigned long long h[2];
    G3_CUDA(ctx, cudaMemcpyAsync(h, cnt, 16, cudaMemcpyDeviceToHost, ctx->stream));
    G3_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (h[0]) {
      if (out[0] == 0) { out[2] = (long long)h[1]; out[3] = it; }
      out[0]++;
      out[1] += (long long)h[0];
    }
  }
  return 0;
}

// ---- Gram on host arrays -----------------------------------------------------------------
int g3_gram(g3_ctx* ctx, const g3_kernel_desc* desc, const double* X1, int n1, const double* X2, int n2, int D,
            const double* theta, int B, double* K_out, int* status) {
  if (!ctx || !desc || !X1 || !K_out || n1 <= 0 || B <= 0 || D <= 0 || D > G3_MAX_DIM)
    return g3_fail_msg(ctx, "g3_gram: bad arguments");
  G3_CUDA(ctx, cudaSetDevice(ctx->device));
  const int same = X2 == nullptr;
  if (same) n2 = n1;
  if (n2 <= 0) return g3_fail_msg(ctx, "g3_gram: n2 <= 0");
  const int P = desc->n_theta;
  const int Np1 = g3_pad(n1), Np2 = g3_pad(n2);
  double* dX1 = (double*)g3_ws(ctx, "gram_X1", sizeof(double) * (size_t)n1 * D);
  double* dX2 = same ? dX1 : (double*)g3_ws(ctx, "gram_X2", sizeof(double) * (size_t)n2 * D);
  double* dth = (double*)g3_ws(ctx, "gram_theta", sizeof(double) * (size_t)B * (P > 0 ? P : 1));
  double* dK = (double*)g3_ws(ctx, "gram_K", sizeof(double) * (size_t)B * Np1 * Np2);
  int* dst = (int*)g3_ws(ctx, "gram_status", sizeof(int) * B);
  if (!dX1 || !dX2 || !dth || !dK || !dst) return -2;
  G3_CUDA(ctx, cudaMemcpyAsync(dX1, X1, sizeof(double) * (size_t)n1 * D, cudaMemcpyHostToDevice, ctx->stream));
  if (!same) G3_CUDA(ctx, cudaMemcpyAsync(dX2, X2, sizeof(double) * (size_t)n2 * D, cudaMemcpyHostToDevice, ctx->stream));
  if (P > 0) G3_CUDA(ctx, cudaMemcpyAsync(dth, theta, sizeof(double) * (size_t)B * P, cudaMemcpyHostToDevice, ctx->stream));
  G3_CUDA(ctx, cudaMemsetAsync(dst, 0, sizeof(int) * B, ctx->stream));
  GramArgs a;
  memset(&a, 0, sizeof a);
  a.X1 = dX1; a.X2 = dX2; a.n1 = n1; a.n2 = n2; a.D = D; a.same = same;
  a.theta = dth; a.P = P;
  a.K = dK; a.ldk = Np2; a.strideK = (long long)Np1 * Np2; a.status = dst;
  int rc;
  if ((rc = g3_gram_launch(ctx, *desc, a, B))) return rc;
  G3_CUDA(ctx, cudaMemcpy2DAsync(K_out, sizeof(double) * n2, dK, sizeof(double) * Np2, sizeof(double) * n2,
                                 (size_t)n1, cudaMemcpyDeviceToHost, ctx->stream));
  for (int b = 1; b < B; ++b)
    G3_CUDA(ctx, cudaMemcpy2DAsync(K_out + (size_t)b * n1 * n2, sizeof(double) * n2, dK + (size_t)b * Np1 * Np2,
                                   sizeof(double) * Np2, sizeof(double) * n2, (size_t)n1, cudaMemcpyDeviceToHost,
                                   ctx->stream));
  if (status) G3_CUDA(ctx, cudaMemcpyAsync(status, dst, sizeof(int) * B, cudaMemcpyDeviceToHost, ctx->stream));
  G3_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

int g3_gram_vjp(g3_ctx* ctx, const g3_kernel_desc* desc, const double* X1, int n1, const double* X2, int n2, int D,
                const double* theta, int B, const double* W, double* dtheta) {
  if (!ctx || !desc || !X1 || !W || !dtheta || n1 <= 0 || B <= 0 || D <= 0 || D > G3_MAX_DIM)
    return g3_fail_msg(ctx, "g3_gram_vjp: bad arguments");
  G3_CUDA(ctx, cudaSetDevice(ctx->device));
  const int same = X2 == nullptr;
  if (same) n2 = n1;
  const int P = desc->n_theta;
  if (P <= 0) return 0;
  const int Np1 = g3_pad(n1), Np2 = g3_pad(n2);
  double* dX1 = (double*)g3_ws(ctx, "gram_X1", sizeof(double) * (size_t)n1 * D);
  double* dX2 = same ? dX1 : (double*)g3_ws(ctx, "gram_X2", sizeof(double) * (size_t)n2 * D);
  double* dth = (double*)g3_ws(ctx, "gram_theta", sizeof(double) * (size_t)B * P);
  double* dW = (double*)g3_ws(ctx, "gram_K", sizeof(double) * (size_t)B * Np1 * Np2);
  double* dg = (double*)g3_ws(ctx, "gram_dtheta", sizeof(double) * (size_t)B * P);
  if (!dX1 || !dX2 || !dth || !dW || !dg) return -2;
  G3_CUDA(ctx, cudaMemcpyAsync(dX1, X1, sizeof(double) * (size_t)n1 * D, cudaMemcpyHostToDevice, ctx->stream));
  if (!same) G3_CUDA(ctx, cudaMemcpyAsync(dX2, X2, sizeof(double) * (size_t)n2 * D, cudaMemcpyHostToDevice, ctx->stream));
  G3_CUDA(ctx, cudaMemcpyAsync(dth, theta, sizeof(double) * (size_t)B * P, cudaMemcpyHostToDevice, ctx->stream));
  G3_CUDA(ctx, cudaMemsetAsync(dW, 0, sizeof(double) * (size_t)B * Np1 * Np2, ctx->stream));
  for (int b = 0; b < B; ++b)
    G3_CUDA(ctx, cudaMemcpy2DAsync(dW + (size_t)b * Np1 * Np2, sizeof(double) * Np2, W + (size_t)b * n1 * n2,
                                   sizeof(double) * n2, sizeof(double) * n2, (size_t)n1, cudaMemcpyHostToDevice,
                                   ctx->stream));
  VjpArgs v;
  memset(&v, 0, sizeof v);
  v.X1 = dX1; v.X2 = dX2; v.n1 = n1; v.n2 = n2; v.D = D; v.same = same; v.lower_only = 0;
  v.theta = dth; v.P = P; v.W = dW; v.ldw = Np2; v.strideW = (long long)Np1 * Np2;
  v.scale = 1.0; v.dtheta = dg;
  int rc;
  if ((rc = g3_gram_vjp_launch(ctx, *desc, v, B))) return rc;
  G3_CUDA(ctx, cudaMemcpyAsync(dtheta, dg, sizeof(double) * (size_t)B * P, cudaMemcpyDeviceToHost, ctx->stream));
  G3_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

// ---- robust Cholesky on host matrices -------------------------------------------------------
int g3_potrf_robust(g3_ctx* ctx, double* A, int n, int lda, int B, int* info_out, double* jitter_out) {
  if (!ctx || !A || n <= 0 || lda < n || B <= 0 || !info_out) return g3_fail_msg(ctx, "g3_potrf_robust: bad arguments");
  G3_CUDA(ctx, cudaSetDevice(ctx->device));
  const int Np = g3_pad(n), T = Np / TS;
  const size_t mat = (size_t)Np * Np;
  double* dA = (double*)g3_ws(ctx, "pr_A", sizeof(double) * mat * B);
  double* dDinv = (double*)g3_ws(ctx, "pr_Dinv", sizeof(double) * (size_t)B * T * TS * TS);
  double* dld = (double*)g3_ws(ctx, "pr_logdet", sizeof(double) * B);
  double* dsh = (double*)g3_ws(ctx, "pr_shift", sizeof(double) * B);
  int* dinfo = (int*)g3_ws(ctx, "pr_info", sizeof(int) * B);
  int* dbmap = (int*)g3_ws(ctx, "pr_bmap", sizeof(int) * B);
  if (!dA || !dDinv || !dld || !dsh || !dinfo || !dbmap) return -2;
  int rc;
  auto upload = [&](int b) -> int {
    double* d = dA + mat * b;
    if (Np != n) G3_CUDA(ctx, cudaMemsetAsync(d, 0, sizeof(double) * mat, ctx->stream));
    G3_CUDA(ctx, cudaMemcpy2DAsync(d, sizeof(double) * Np, A + (size_t)b * n * lda, sizeof(double) * lda,
                                   sizeof(double) * n, (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    return 0;
  };
  for (int b = 0; b < B; ++b)
    if ((rc = upload(b))) return rc;
  mat_pad_shift_kernel<<<dim3((Np + 255) / 256, B), 256, 0, ctx->stream>>>(dA, n, Np, (long long)mat, nullptr, nullptr);
  G3_LAUNCH_CHECK(ctx);
  G3_CUDA(ctx, cudaMemsetAsync(dld, 0, sizeof(double) * B, ctx->stream));
  G3_CUDA(ctx, cudaMemsetAsync(dinfo, 0, sizeof(int) * B, ctx->stream));
  if ((rc = g3_potrf_batched(ctx, dA, Np, B, dDinv, dld, dinfo, nullptr, 0, ctx->potrf_w))) return rc;
  std::vector<int> info(B);
  G3_CUDA(ctx, cudaMemcpyAsync(info.data(), dinfo, sizeof(int) * B, cudaMemcpyDeviceToHost, ctx->stream));
  G3_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  std::vector<int> failed;
  for (int b = 0; b < B; ++b) {
    info_out[b] = 0;
    if (jitter_out) jitter_out[b] = 0.0;
    if (info[b] != 0) failed.push_back(b);
  }
  if (!failed.empty()) {
    std::vector<double> dK(B, 0.0), pre(B, 0.0), sh(B, 0.0);
    for (int b : failed) {  // tensors.py:203-207 (host O(n) scan of the caller's diagonal)
      const double* Ab = A + (size_t)b * n * lda;
      double sum = 0.0, mn = INFINITY;
      bool nonpos = false;
      for (int i = 0; i < n; ++i) {
        const double d = Ab[(size_t)i * lda + i];
        sum += d;
        if (d < mn) mn = d;
        if (d <= 0.0) nonpos = true;
      }
      const double mean = sum / n;
      dK[b] = mean * ctx->jitter_rel;
      if (nonpos) pre[b] = mean * ctx->jitter_rel - mn;
    }
    for (int t = 1; t <= ctx->max_tries && !failed.empty(); ++t) {
      const int nb = (int)failed.size();
      for (int b : failed) {
        sh[b] = pre[b] + dK[b];
        if ((rc = upload(b))) return rc;
      }
      G3_CUDA(ctx, cudaMemcpyAsync(dsh, sh.data(), sizeof(double) * B, cudaMemcpyHostToDevice, ctx->stream));
      G3_CUDA(ctx, cudaMemcpyAsync(dbmap, failed.data(), sizeof(int) * nb, cudaMemcpyHostToDevice, ctx->stream));
      mat_pad_shift_kernel<<<dim3((Np + 255) / 256, nb), 256, 0, ctx->stream>>>(dA, n, Np, (long long)mat, dsh, dbmap);
      G3_LAUNCH_CHECK(ctx);
      gp_zero_sub_kernel<<<(nb + 127) / 128, 128, 0, ctx->stream>>>(nullptr, dld, dinfo, dbmap, nb);
      G3_LAUNCH_CHECK(ctx);
      if ((rc = g3_potrf_batched(ctx, dA, Np, B, dDinv, dld, dinfo, dbmap, nb, ctx->potrf_w))) return rc;
      G3_CUDA(ctx, cudaMemcpyAsync(info.data(), dinfo, sizeof(int) * B, cudaMemcpyDeviceToHost, ctx->stream));
      G3_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      std::vector<int> still;
      for (int b : failed) {
        if (info[b] != 0) {
          still.push_back(b);
          dK[b] *= 10.0;
        } else {
          info_out[b] = t;
          if (jitter_out) jitter_out[b] = sh[b];
        }
      }
      failed.swap(still);
    }
    for (int b : failed) info_out[b] = -1;
  }
  if (T > 1) {
    zero_upper_tiles_kernel<<<dim3(T * (T - 1) / 2, B), 256, 0, ctx->stream>>>(dA, Np, (long long)mat, T);
    G3_LAUNCH_CHECK(ctx);
  }
  for (int b = 0; b < B; ++b) {
    if (info_out[b] < 0) continue;  // caller applies the 1e-10*I fallback (tensors.py:218-222)
    G3_CUDA(ctx, cudaMemcpy2DAsync(A + (size_t)b * n * lda, sizeof(double) * lda, dA + mat * b, sizeof(double) * Np,
                                   sizeof(double) * n, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  }
  G3_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

// robust Cholesky of ONE host matrix followed by the forward substitution u = L^-1 rhs on the device (L and the
// inverses of its diagonal blocks are still resident in the pr_* workspaces)
int g3_potrf_robust_solve(g3_ctx* ctx, double* A, int n, int lda, const double* rhs, double* u_out, int* info_out,
                          double* jitter_out) {
  if (!rhs || !u_out) return g3_fail_msg(ctx, "g3_potrf_robust_solve: bad arguments");
  int rc = g3_potrf_robust(ctx, A, n, lda, 1, info_out, jitter_out);
  if (rc) return rc;
  if (info_out[0] < 0) return 0;                  // exhausted ladder: the caller applies L = 1e-10*I itself
  const int Np = g3_pad(n);
  double* dA = (double*)g3_ws(ctx, "pr_A", sizeof(double) * (size_t)Np * Np);
  double* dDinv = (double*)g3_ws(ctx, "pr_Dinv", sizeof(double) * (size_t)(Np / TS) * TS * TS);
  double* dr = (double*)g3_ws(ctx, "pr_r", sizeof(double) * (2 * (size_t)Np + 1));
  if (!dA || !dDinv || !dr) return -2;
  double* du = dr + Np;
  G3_CUDA(ctx, cudaMemsetAsync(dr, 0, sizeof(double) * (2 * (size_t)Np + 1), ctx->stream));
  G3_CUDA(ctx, cudaMemcpyAsync(dr, rhs, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = g3_trsv_fwd(ctx, dA, dDinv, dr, du, du + Np, Np, 1))) return rc;
  G3_CUDA(ctx, cudaMemcpyAsync(u_out, du, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  G3_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

// ---- posterior moments for one theta -----------------------------------------------------------
int g3_gp_posterior(g3_ctx* ctx, const g3_kernel_desc* desc, const double* Xs, int M, const double* delta,
                    const double* theta, int flags, double* mean_out, double* var_out, double* cov_out,
                    double* beta_out, int* status) {
  if (!ctx || !desc || !Xs || M <= 0 || !delta || !mean_out || !var_out)
    return g3_fail_msg(ctx, "g3_gp_posterior: bad arguments");
  if (!ctx->dX) return g3_fail_msg(ctx, "g3_gp_posterior: call g3_set_data first");
  if ((flags & G3_POST_COV) && !cov_out) return g3_fail_msg(ctx, "g3_gp_posterior: G3_POST_COV needs cov_out");
  int rc;
  // factor K (and run the ladder if needed) through the logp path: leaves L, Dinv, u in the gp_* buffers
  double beta = 0.0, logdet = 0.0;
  int st = 0;
  if ((rc = g3_gp_upload(ctx, desc, G3_KIND_GAUSS, delta, 0, theta, 1, nullptr, 0))) return rc;
  if ((rc = g3_gp_run(ctx))) return rc;
  if ((rc = g3_gp_download(ctx, &beta, &logdet, nullptr, nullptr, &st))) return rc;
  if (beta_out) *beta_out = beta;
  if (status) *status = st;
  const int N = ctx->N, D = ctx->D, P = desc->n_theta, Np = g3_pad(N), Mp = g3_pad(M), T = Np / TS, TM = Mp / TS;
  GpBufs w;
  if ((rc = gp_alloc(ctx, w, 1, P, N, 0, 1))) return rc;
  const int skip_pn = (flags & G3_POST_NOISE) ? 0 : 1;
  double* dXs = (double*)g3_ws(ctx, "post_Xs", sizeof(double) * (size_t)M * D);
  double* Vt = (double*)g3_ws(ctx, "post_Vt", sizeof(double) * (size_t)Mp * Np);
  double* dmean = (double*)g3_ws(ctx, "post_mean", sizeof(double) * Mp);
  double* dvar = (double*)g3_ws(ctx, "post_var", sizeof(double) * Mp);
  double* kss = (double*)g3_ws(ctx, "post_kss", sizeof(double) * 4);
  double* kss_vec = (double*)g3_ws(ctx, "post_kss_vec", sizeof(double) * Mp);
  if (!dXs || !Vt || !dmean || !dvar || !kss || !kss_vec) return -2;
  G3_CUDA(ctx, cudaMemcpyAsync(dXs, Xs, sizeof(double) * (size_t)M * D, cudaMemcpyHostToDevice, ctx->stream));
  // K* = cov(space, inputs)  (cross form: Noise contributes zeros, kernels.py:367-371), tt_to_num scrubbed
  GramArgs a;
  memset(&a, 0, sizeof a);
  a.X1 = dXs; a.X2 = ctx->dX; a.n1 = M; a.n2 = N; a.D = D; a.same = 0; a.skip_process_noise = skip_pn;
  a.theta = w.theta; a.P = P; a.K = Vt; a.ldk = Np; a.strideK = (long long)Mp * Np; a.status = nullptr;
  if ((rc = g3_gram_launch(ctx, *desc, a, 1))) return rc;
  // Vt = K* L^-T by blocked forward substitution; all level-3 through the NT GEMM
  CUtensorMap tmV, tmL, tmD;
  if ((rc = g3_make_tmap(ctx, &tmV, Vt, Np, Mp, 1, Np, (uint64_t)Mp * Np, G3_BM))) return rc;
  if ((rc = g3_make_tmap(ctx, &tmL, w.A, Np, Np, 1, Np, (uint64_t)Np * Np, G3_BN))) return rc;
  if ((rc = g3_make_tmap(ctx, &tmD, w.Dinv, TS, (uint64_t)T * TS, 1, TS, (uint64_t)T * TS * TS, G3_BN))) return rc;
  for (int j = 0; j < T; ++j) {
    if (j > 0) {  // Vt[:, j] -= sum_{k<j} Vt[:, k] L[j][k]^T
      GemmArgs g = gz();
      g.D = Vt; g.ldd = Np; g.strideD = 0;
      g.mode = 0; g.ntx = TM; g.nty = 1;
      g.d_r0 = 0; g.d_c0 = j * TS;
      g.a_r0 = 0; g.a_rx = TS; g.ka0 = 0;
      g.b_r0 = j * TS; g.kb0 = 0;
      g.kl0 = j * TS;
      g.alpha = -1.0; g.beta = 1.0;
      if ((rc = g3_gemm_launch(ctx, tmV, tmL, g, 1))) return rc;
    }
    GemmArgs g = gz();  // Vt[:, j] = Vt[:, j] Linv_jj^T
    g.D = Vt; g.ldd = Np; g.strideD = 0;
    g.mode = 0; g.ntx = TM; g.nty = 1;
    g.d_r0 = 0; g.d_c0 = j * TS;
    g.a_r0 = 0; g.a_rx = TS; g.ka0 = j * TS;
    g.b_r0 = j * TS; g.kb0 = 0;
    g.kl0 = TS;
    g.alpha = 1.0; g.beta = 0.0; g.tri_b = 1;
    if ((rc = g3_gemm_launch(ctx, tmV, tmD, g, 1))) return rc;
  }
  // K** diagonal: the tree evaluated at (x*, x*) per test point (constant for stationary leaves), and its minimum
  if ((rc = g3_gram_diag_min(ctx, *desc, dXs, M, D, w.theta, P, 1, kss, kss + 1, nullptr, skip_pn, kss_vec))) return rc;
  // tt_to_cov on prior_kernel_space only for the noisy selector (elliptical.py:70 vs :73)
  post_kss_shift_kernel<<<1, 32, 0, ctx->stream>>>(kss, ctx->jitter_rel, (flags & G3_POST_NOISE) ? 1 : 0);
  G3_LAUNCH_CHECK(ctx);
  post_moments_kernel<<<(M + 7) / 8, 256, 0, ctx->stream>>>(Vt, Np, M, w.u, kss_vec, kss + 2, dmean, dvar);
  G3_LAUNCH_CHECK(ctx);
  G3_CUDA(ctx, cudaMemcpyAsync(mean_out, dmean, sizeof(double) * M, cudaMemcpyDeviceToHost, ctx->stream));
  G3_CUDA(ctx, cudaMemcpyAsync(var_out, dvar, sizeof(double) * M, cudaMemcpyDeviceToHost, ctx->stream));
  if (flags & G3_POST_COV) {
    double* C = (double*)g3_ws(ctx, "post_C", sizeof(double) * (size_t)Mp * Mp);
    if (!C) return -2;
    GramArgs c;
    memset(&c, 0, sizeof c);
    c.X1 = dXs; c.X2 = dXs; c.n1 = M; c.n2 = M; c.D = D; c.same = 1; c.skip_process_noise = skip_pn;
    c.theta = w.theta; c.P = P; c.K = C; c.ldk = Mp; c.strideK = (long long)Mp * Mp;
    c.diag_shift = (flags & G3_POST_NOISE) ? kss + 2 : nullptr;
    if ((rc = g3_gram_launch(ctx, *desc, c, 1))) return rc;
    CUtensorMap tmVb;
    if ((rc = g3_make_tmap(ctx, &tmVb, Vt, Np, Mp, 1, Np, (uint64_t)Mp * Np, G3_BN))) return rc;
    GemmArgs g = gz();  // C -= Vt Vt^T
    g.D = C; g.ldd = Mp; g.strideD = 0;
    g.mode = 0; g.ntx = TM; g.nty = TM;
    g.a_r0 = 0; g.a_rx = TS; g.b_r0 = 0; g.b_ry = TS;
    g.kl0 = Np;
    g.alpha = -1.0; g.beta = 1.0;
    if ((rc = g3_gemm_launch(ctx, tmV, tmVb, g, 1))) return rc;
    G3_CUDA(ctx, cudaMemcpy2DAsync(cov_out, sizeof(double) * M, C, sizeof(double) * Mp, sizeof(double) * M, (size_t)M,
                                   cudaMemcpyDeviceToHost, ctx->stream));
  }
  G3_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

// ---- device-pointer entry points for the multi-GPU block-cyclic Cholesky (g3py_b200/dist_potrf.py) ---------
int g3_dev_gram_block(g3_ctx* ctx, const g3_kernel_desc* desc, const double* theta, int row0, int col0, int rows,
                      int cols, double diag_shift, double* out, long long ld) {
  if (!ctx || !desc || !out || rows <= 0 || cols <= 0 || row0 < 0 || col0 < 0) return g3_fail_msg(ctx, "g3_dev_gram_block: bad arguments");
  if (!ctx->dX) return g3_fail_msg(ctx, "g3_dev_gram_block: call g3_set_data first");
  if (rows % TS || cols % TS) return g3_fail_msg(ctx, "g3_dev_gram_block: rows/cols must be multiples of 128");
  G3_CUDA(ctx, cudaSetDevice(ctx->device));
  const int N = ctx->N, P = desc->n_theta;
  int rc = g3_check_desc(ctx, *desc, ctx->D);
  if (rc) return rc;
  double* dth = (double*)g3_ws(ctx, "blk_theta", sizeof(double) * ((P > 0 ? P : 1) + 1));
  if (!dth) return -2;
  std::vector<double> h(P + 1);
  for (int p = 0; p < P; ++p) h[p] = theta[p];
  h[P] = diag_shift;
  G3_CUDA(ctx, cudaMemcpyAsync(dth, h.data(), sizeof(double) * (P + 1), cudaMemcpyHostToDevice, ctx->stream));
  if (N % TS) return g3_fail_msg(ctx, "g3_dev_gram_block: N must be a multiple of 128 on the multi-GPU path");
  if (row0 + rows > N || col0 + cols > N) return g3_fail_msg(ctx, "g3_dev_gram_block: block outside the matrix");
  GramArgs a;
  memset(&a, 0, sizeof a);
  a.X1 = ctx->dX + (size_t)row0 * ctx->D; a.X2 = ctx->dX + (size_t)col0 * ctx->D;
  a.n1 = rows; a.n2 = cols;
  a.D = ctx->D; a.same = 1; a.lower_only = 0; a.pad_identity = 0; a.diag_off = row0 - col0;
  a.theta = dth; a.P = P; a.diag_shift = dth + P;
  a.K = out; a.ldk = ld; a.strideK = 0;
  return g3_gram_launch(ctx, *desc, a, 1);
}

int g3_dev_potrf_panel(g3_ctx* ctx, double* P, int rows, int nb, double* Dinv_dev_or_NULL, double* logdet_dev,
                       int* info_dev) {
  if (!ctx || !P || !logdet_dev || !info_dev) return g3_fail_msg(ctx, "g3_dev_potrf_panel: bad arguments");
  G3_CUDA(ctx, cudaSetDevice(ctx->device));
  double* Dinv = Dinv_dev_or_NULL ? Dinv_dev_or_NULL
                                  : (double*)g3_ws(ctx, "blk_Dinv", sizeof(double) * (size_t)(nb / TS) * TS * TS);
  if (!Dinv) return -2;
  return g3_potrf_panel(ctx, P, rows, nb, Dinv, logdet_dev, info_dev);
}

int g3_dev_trsv_panel(g3_ctx* ctx, const double* P, int rows, int nb, const double* Dinv_dev, double* r_dev,
                      double* u_dev, double* beta_dev) {
  if (!ctx || !P || !Dinv_dev || !r_dev || !u_dev) return g3_fail_msg(ctx, "g3_dev_trsv_panel: bad arguments");
  G3_CUDA(ctx, cudaSetDevice(ctx->device));
  return g3_trsv_panel(ctx, P, rows, nb, Dinv_dev, r_dev, u_dev, beta_dev);
}

int g3_dev_syrk_panel(g3_ctx* ctx, const double* P, int rowsP, int nb, int row_off, double* D, int rowsD) {
  if (!ctx || !P || !D) return g3_fail_msg(ctx, "g3_dev_syrk_panel: bad arguments");
  G3_CUDA(ctx, cudaSetDevice(ctx->device));
  return g3_syrk_panel(ctx, P, rowsP, nb, row_off, D, rowsD);
}

// ---- big single-matrix Cholesky (BASELINE metric 2) -------------------------------------------
int g3_gram_potrf_device(g3_ctx* ctx, const g3_kernel_desc* desc, const double* theta, double* logdet, int* info,
                         float* ms_gram, float* ms_potrf) {
  if (!ctx || !desc || !logdet || !info) return g3_fail_msg(ctx, "g3_gram_potrf_device: bad arguments");
  if (!ctx->dX) return g3_fail_msg(ctx, "g3_gram_potrf_device: call g3_set_data first");
  G3_CUDA(ctx, cudaSetDevice(ctx->device));
  const int N = ctx->N, Np = g3_pad(N), T = Np / TS, P = desc->n_theta;
  int rc = g3_check_desc(ctx, *desc, ctx->D);
  if (rc) return rc;
  double* A = (double*)g3_ws(ctx, "gp_A", sizeof(double) * (size_t)Np * Np);
  double* Dinv = (double*)g3_ws(ctx, "gp_Dinv", sizeof(double) * (size_t)T * TS * TS);
  double* dth = (double*)g3_ws(ctx, "gp_theta", sizeof(double) * (P > 0 ? P : 1));
  double* sc = (double*)g3_ws(ctx, "big_scalars", sizeof(double) * 8);
  int* di = (int*)g3_ws(ctx, "big_info", sizeof(int) * 4);
  if (!A || !Dinv || !dth || !sc || !di) return -2;
  ctx->gp.valid = 0;
  if (P > 0) G3_CUDA(ctx, cudaMemcpyAsync(dth, theta, sizeof(double) * P, cudaMemcpyHostToDevice, ctx->stream));
  G3_CUDA(ctx, cudaMemsetAsync(di, 0, sizeof(int) * 4, ctx->stream));
  if ((rc = g3_gram_diag_min(ctx, *desc, ctx->dX, N, ctx->D, dth, P, 1, sc, sc + 1, di + 1, 0))) return rc;
  gp_prep_kernel<<<1, 32, 0, ctx->stream>>>(sc, ctx->jitter_rel, sc + 2, sc + 3, sc + 4, di, di + 1, 1);
  G3_LAUNCH_CHECK(ctx);
  cudaEvent_t e0, e1, e2;
  G3_CUDA(ctx, cudaEventCreate(&e0));
  G3_CUDA(ctx, cudaEventCreate(&e1));
  G3_CUDA(ctx, cudaEventCreate(&e2));
  G3_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
  GramArgs a;
  memset(&a, 0, sizeof a);
  a.X1 = ctx->dX; a.X2 = ctx->dX; a.n1 = N; a.n2 = N; a.D = ctx->D;
  a.same = 1; a.lower_only = 1; a.pad_identity = 1;
  a.theta = dth; a.P = P; a.diag_shift = sc + 2;
  a.K = A; a.ldk = Np; a.strideK = (long long)Np * Np; a.status = di + 1;
  if ((rc = g3_gram_launch(ctx, *desc, a, 1))) return rc;
  G3_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
  if ((rc = g3_potrf_batched(ctx, A, Np, 1, Dinv, sc + 4, di, nullptr, 0, ctx->potrf_w_big))) return rc;
  G3_CUDA(ctx, cudaEventRecord(e2, ctx->stream));
  G3_CUDA(ctx, cudaMemcpyAsync(logdet, sc + 4, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  G3_CUDA(ctx, cudaMemcpyAsync(info, di, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  G3_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  float t;
  if (ms_gram) { cudaEventElapsedTime(&t, e0, e1); *ms_gram = t; }
  if (ms_potrf) { cudaEventElapsedTime(&t, e1, e2); *ms_potrf = t; }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
  return 0;
}

}  // extern "C"

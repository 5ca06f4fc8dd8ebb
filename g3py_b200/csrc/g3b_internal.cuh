// Internal declarations shared by the translation units of libg3b.so (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include <map>
#include "../../include/g3b.h"

// NVTX ranges around the host-side stages (header-only nvtx3: no link dependency; a no-op without a profiler attached)
#include <nvtx3/nvToolsExt.h>
struct g3_nvtx_range {
  explicit g3_nvtx_range(const char* name) { nvtxRangePushA(name); }
  ~g3_nvtx_range() { nvtxRangePop(); }
};
#define G3_NVTX(name) g3_nvtx_range g3_nvtx_scope_##__LINE__(name)

#define G3_TILE 128          // factorisation block size: every device matrix is padded to a multiple
#define G3_BM 64             // GEMM CTA tile rows   (two CTAs per 128-row block)
#define G3_BN 128            // GEMM CTA tile cols
#define G3_BK 16             // doubles per k-tile = 128 bytes = one SWIZZLE_128B row
#define G3_STAGES 4
#define G3_MAX_GROUPS 8      // batch groups processed concurrently on their own streams

struct g3_buf {
  void* p = nullptr;
  size_t bytes = 0;
};

struct g3_gp_state {
  g3_kernel_desc desc;
  int kind = 0, B = 0, want_grad = 0, delta_stride = 0, valid = 0;
  int factor_resident = 0;     // L, Dinv, u of the last logp-only evaluation are still in the workspaces
  int trtri_spec = 0;          // ... and U = L^-T too: it was pipelined behind that factorisation (g3_set_speculate_grad)
};

struct g3_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;       // stream launches are issued to (switched to a group stream inside g3_gp_run)
  cudaStream_t own_stream = nullptr;   // the stream created by g3_ctx_create (g3_set_stream can substitute another)
  cudaStream_t gstream[G3_MAX_GROUPS] = {};
  cudaEvent_t gev_start = nullptr, gev_done[G3_MAX_GROUPS] = {};
  int n_groups = 4;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::string err;
  int64_t launches = 0;
  double jitter_rel = 9.999999974752427e-07;  // float32(1e-6), libs/tensors.py:204
  int max_tries = 20;
  // resident observations
  double* dX = nullptr;
  int N = 0, D = 0;
  std::map<std::string, g3_buf> bufs;  // grow-only named workspaces
  std::map<std::string, g3_buf> pinned;  // grow-only page-locked host staging buffers (hot-path H2D / D2H)
  cudaEvent_t ev_h2d = nullptr;        // last staged upload has left the pinned buffers
  g3_gp_state gp;
  void* encode_fn = nullptr;           // cuTensorMapEncodeTiled
  int sm_count = 148;
  bool gemm_ready = false, diag_ready = false, diag2_ready = false;
  int diag_variant = 2;                // 128x128 diagonal-tile kernel: 2 = diag.cu (low latency), 1 = the first kernel in potrf.cu
  // optional per-launch-class device timing (bench.py roofline): event pairs recorded around launches
  bool prof_on = false;
  std::vector<cudaEvent_t> prof_events;     // pairs
  std::vector<int> prof_class;
  size_t prof_used = 0;                     // pairs in use
  int potrf_w = 0;                     // tile columns per right-looking outer block in the gp path; 0 = choose from B and T
  int potrf_w_big = 8;                 // same, single big matrix (g3_gram_potrf_device)
  // look-ahead of the right-looking factorisation: the next panel is updated and factored on a high-priority
  // stream while the rest of the trailing update runs on the main stream
  cudaStream_t panel_stream = nullptr;
  cudaEvent_t ev_panel = nullptr, ev_main = nullptr;
  int lookahead = 1;
  cudaStream_t tri_stream = nullptr;   // U = L^-T pipelined behind the look-ahead factorisation
  cudaEvent_t ev_tri = nullptr;
  int trtri_pipeline = 1, trtri_done = 0;
  int force_left = 0;                  // set by g3_gp_run for batches of more than 8 items (their stream groups hold 8)
  int splitk = 1;                      // allow split-K for few-tile / deep-K GEMM launches (g3_set_splitk)
  int speculate_grad = 0;              // logp-only evaluations of <= 8 matrices also pipeline U = L^-T behind the factorisation
  int graph_trtri_spec = 0;            // whether the cached graph of such an evaluation contains that pipeline
  int trsv_fused = 1;                  // whole triangular solves in one launch (flags between CTAs) instead of T dependent launches
  int tile_split = 1;                  // spread each output tile of a few-tile GEMM launch over 2 / 4 CTAs by columns (g3_set_tile_split)
  struct g3_dist* dist = nullptr;      // multi-GPU state (dist.cu): NCCL communicator, block-cyclic panels
  // CUDA-graph replay of the launch sequence of small batches (gp.cu: g3_gp_run)
  int graphs_on = 1;
  uint64_t ws_gen = 0;                 // bumped whenever a workspace or the resident X is (re)allocated: cached graphs hold addresses
  void* graph_exec = nullptr;          // cudaGraphExec_t
  std::string graph_key, graph_warm_key;
  int64_t graph_launches = 0, graph_replays = 0;
  int gemm_mode = 0;                   // G3_GEMM_DMMA / G3_GEMM_OZAKI (g3_set_gemm_mode)
  int oz_min_k = 1024;                 // Ozaki updates only for contractions at least this deep (DMMA below)
  int64_t oz_launches = 0;             // int8 tensor-core update launches since creation
};
void g3_dist_destroy(g3_ctx* ctx);     // called by g3_ctx_destroy
void g3_graph_drop(g3_ctx* ctx);       // forget the cached CUDA graph (gp.cu)

enum { G3_PROF_GEMM = 0, G3_PROF_DIAG = 1, G3_PROF_GRAM = 2, G3_PROF_VJP = 3, G3_PROF_TRSV = 4, G3_PROF_OTHER = 5, G3_PROF_N = 6 };
void g3_prof_begin(g3_ctx* ctx, int cls);
void g3_prof_end(g3_ctx* ctx);

// ---- host helpers (ctx.cu) ----
int g3_fail(g3_ctx* ctx, const char* what, cudaError_t e, const char* file, int line);
int g3_fail_msg(g3_ctx* ctx, const std::string& msg);
void* g3_ws(g3_ctx* ctx, const char* name, size_t bytes);   // returns nullptr on failure (err set)
void* g3_pinned(g3_ctx* ctx, const char* name, size_t bytes);   // page-locked host staging, nullptr on failure
int g3_make_tmap(g3_ctx* ctx, CUtensorMap* out, const double* base, uint64_t cols, uint64_t rows,
                 uint64_t batch, uint64_t ld, uint64_t batch_stride, uint32_t box_rows);

#define G3_CUDA(ctx, call)                                                     \
  do {                                                                         \
    cudaError_t _e = (call);                                                   \
    if (_e != cudaSuccess) return g3_fail((ctx), #call, _e, __FILE__, __LINE__); \
  } while (0)

#define G3_LAUNCH_CHECK(ctx)                                                   \
  do {                                                                         \
    (ctx)->launches++;                                                         \
    cudaError_t _e = cudaGetLastError();                                       \
    if (_e != cudaSuccess) return g3_fail((ctx), "kernel launch", _e, __FILE__, __LINE__); \
  } while (0)

static inline int g3_pad(int n) { return (n + G3_TILE - 1) / G3_TILE * G3_TILE; }

// ---- GEMM (gemm.cu):  D[m][n] = beta*D[m][n] + alpha * sum_k A[m][k] * B[n][k] ----
// Both operands are K-contiguous ("NT"), fetched by TMA (3-D maps: k, row, batch).
// Tile (x,y) of a launch addresses operands through affine maps so that every blocked step of
// potrf / trtri / lauum / trsm is one launch of the same kernel.
struct GemmArgs {
  double* D;
  long long ldd, strideD;
  int mode;               // 0: RECT ntx x nty tiles, 1: TRI (x >= y), size ntx
  int ntx, nty;
  int d_r0, d_c0;         // D tile (x,y) starts at (d_r0 + 128x, d_c0 + 128y)
  int a_r0, a_rx, a_ry;   // A rows start at a_r0 + x*a_rx + y*a_ry
  int b_r0, b_rx, b_ry;
  int ka0, ka_x, ka_y;    // first k column in A
  int kb0, kb_x, kb_y;    // first k column in B
  int kl0, kl_x, kl_y;    // contraction length (multiple of 16)
  double alpha, beta;
  const int* bmap;        // optional: launch batch index -> matrix index
  int tri_b;              // B operand is a lower-triangular 128x128 block (B[n][k] = 0 for k > n, K = 128):
                          // a warp skips the k-tiles beyond its 32 output columns
  int splitk;             // > 1: the contraction is split over gridDim.z CTAs per tile; partial tiles go to sk_ws and
  double* sk_ws;          //      the CTA that arrives last sums them in split order (deterministic) and writes D.
  unsigned* sk_cnt;       //      Set by g3_gemm_launch for launches with few tiles and a deep contraction.
  int upper;              // bit0: tiles x == y are diagonal tiles of a symmetric result, bit1: tile x == 0 is,
                          // (only their lower triangle is needed: the upper-right 64x64 quarter is not computed)
                          // bit2: tiles with x < y are void (skipped entirely)
};
int g3_gemm_launch(g3_ctx* ctx, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& a, int B);

// ---- factorisation (potrf.cu) ----
// All matrices: Np x Np (Np multiple of 128), row-major, ld = Np, batch stride Np*Np.
// Dinv: [B][T][128][128] inverses of the diagonal blocks of L (lower, explicit zeros above).
// U_pipe (optional, blocked look-ahead schedule only): also compute U = L^-T, pipelined behind the factorisation;
// ctx->trtri_done tells the caller whether that happened.
int g3_potrf_batched(g3_ctx* ctx, double* A, int Np, int Btotal, double* Dinv, double* logdet, int* info,
                     const int* bmap, int nb, int w_outer, double* U_pipe = nullptr);
// Tall panel (rows x nb, ld = nb, rows/nb multiples of 128): factor the top nb x nb block and solve the rows below.
int g3_potrf_panel(g3_ctx* ctx, double* P, int rows, int nb, double* Dinv, double* logdet, int* info);
// diag.cu: factor + invert the diagonal tile j of B matrices (one CTA each); stamps (optional) receives clock64 phase marks of CTA 0
int g3_diag2_launch(g3_ctx* ctx, double* A, int Np, long long strideA, int j, double* Dinv, int T, double* logdet, int* info,
                    const int* bmap, int B, long long* stamps, int dbg = 0);
// around every set of diagonal tiles g3_diag2_launch factors, on the same stream: prepare before the first tile (zeros above
// the diagonal of the Dinv tiles), finish after the last GEMM update of the factorisation (zeros above the diagonal of A's tiles)
int g3_diag2_prepare(g3_ctx* ctx, int j0, int nj, double* Dinv, int T, const int* bmap, int B);
int g3_diag2_finish(g3_ctx* ctx, double* A, int Np, long long strideA, int j0, int nj, const int* bmap, int B);
// D[x][y] -= sum_k P[row_off + x][k] P[row_off + y][k]   (D: rowsD x nb, ld = nb; P: rowsP x nb)
int g3_syrk_panel(g3_ctx* ctx, const double* P, int rowsP, int nb, int row_off, double* D, int rowsD);
int g3_trsv_panel(g3_ctx* ctx, const double* P, int rows, int nb, const double* Dinv, double* r, double* u, double* beta);
int g3_panel_solve(g3_ctx* ctx, double* P, int rows, int nb, const double* Ld, const double* Dinv);
int g3_panel_update(g3_ctx* ctx, double* D, int rows, int nb, const double* A, const double* Bm, int has_diag);
int g3_panel_gemm(g3_ctx* ctx, double* D, long long ldd, int rows, int ncols, const double* A, long long lda, const double* Bm,
                  long long ldb, long long kdim, double alpha, double beta);
int g3_panel_rsolve(g3_ctx* ctx, double* Y, int rows, int nb, const double* LT, const double* DinvT, double s_upd, double s_mul);
int g3_trtri_batched(g3_ctx* ctx, const double* L, double* U, int Np, int B, const double* Dinv);
int g3_lauum_batched(g3_ctx* ctx, const double* U, double* Kinv, int Np, int B);
int g3_trsv_fwd(g3_ctx* ctx, const double* L, const double* Dinv, double* r, double* u, double* beta,
                int Np, int B);
int g3_trsv_bwd(g3_ctx* ctx, const double* L, const double* Dinv, double* s, double* alpha, int Np, int B);

// ---- int8 tensor-core (Ozaki) panel updates (ozaki.cu) ----
struct g3_oz_state {
  int Np = 0, B = 0;
  long long plane_stride = 0;          // bytes between significance planes: B * Np * Np
  int8_t* planes = nullptr;            // [9][B Np][Np] slices of L
  double* scale = nullptr;             // [B Np] power-of-two row scales
  CUtensorMap tmA, tmB;
};
int g3_oz_prepare(g3_ctx* ctx, const double* A, int Np, int B, g3_oz_state* st);
int g3_oz_slice(g3_ctx* ctx, const g3_oz_state* st, const double* A, int j0, int j1);
int g3_oz_update(g3_ctx* ctx, const g3_oz_state* st, double* A, int j0, int j1);

// ---- Gram (gram.cu) ----
struct GramArgs {
  const double* X1; const double* X2;  // row-major n x D
  int n1, n2, D;
  int same;                 // x1 is x2 (Noise/WN -> var*I)
  int lower_only;           // write only tiles with row-tile >= col-tile (same==1)
  int pad_identity;         // out is Np1 x Np2 padded; write identity on the padding diagonal
  int diag_off;             // global row - global col of element (0,0) of this block: diagonal is gi + diag_off == gj
  int skip_process_noise;   // noise=False selectors: the auto-added Noise leaf evaluates to 0 (elliptical.py:73-74)
  const double* theta; int P;       // B x P natural-space hypers
  const double* diag_shift;         // optional B: added on the diagonal (tt_to_cov / jitter), may be null
  double* K; long long ldk, strideK;
  int* status;              // optional B
  const int* bmap;
};
int g3_gram_launch(g3_ctx* ctx, const g3_kernel_desc& desc, const GramArgs& a, int B);
// min over the diagonal of cov(X) for each theta (tt_to_cov needs it): out[b]; diag_vec (optional, B x n) receives
// the diagonal itself (the posterior variance of non-stationary kernels needs k(x*, x*) per point)
int g3_gram_diag_min(g3_ctx* ctx, const g3_kernel_desc& desc, const double* X, int n, int D,
                     const double* theta, int P, int B, double* diag_min, double* diag_mean, int* status,
                     int skip_process_noise, double* diag_vec = nullptr);
int g3_check_desc(g3_ctx* ctx, const g3_kernel_desc& d, int D);
struct VjpArgs {
  const double* X1; const double* X2;
  int n1, n2, D, same, lower_only;
  const double* theta; int P;
  const double* W; long long ldw, strideW;   // weights; if alpha != null: W_ij = cfac*alpha_i*alpha_j - W_ij
  const double* alpha; long long strideAlpha; const double* cfac;
  const double* alpha2;     // optional: the column factor of the outer product (off-diagonal blocks: W_ij = c a_i a2_j - W_ij); null = alpha
  double scale;             // result multiplied by scale (0.5 for the GP gradient)
  double* dtheta;           // B x P
  double* partials;         // optional caller-provided scratch: B x ntiles x P
};
int g3_gram_vjp_launch(g3_ctx* ctx, const g3_kernel_desc& desc, const VjpArgs& a, int B);

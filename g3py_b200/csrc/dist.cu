// Multi-GPU exact GP behind the C ABI (SURVEY §8 b/e, K17): NCCL communicator owned by the library (no torch),
// 2-D block-cyclic Cholesky of K = cov(X) on a Pr x Pc process grid with look-ahead, distributed forward solve,
// log-det / beta, and a residual probe that checks the distributed factor against K regenerated from X.
// The reference has no counterpart (its only parallelism is multiprocessing.Pool.map over chains,
// g3py/processes/stochastic.py:775-783); the single-GPU semantics it must match are CholeskyRobust.perform
// (g3py/libs/tensors.py:197-222) and logp_cho (g3py/processes/gaussian.py:208-224).
//
// Layout.  Block size nb (multiple of 128), nP = N / nb block rows / panels.  Rank r = q * Pr + p sits at process row p,
// process column q.  Panel J (columns [J nb, (J+1) nb)) belongs to process column J mod Pc; block row I to process row
// I mod Pr.  Rank (p, q) stores, for each of its panels J, the PIECE {blocks (I, J): I >= J, I mod Pr == p}, contiguous in
// ascending I as a tall (count * nb) x nb matrix with leading dimension nb - already the K-contiguous operand of the NT
// GEMM, generated locally from the replicated X (K itself never crosses a link).
//
// Step J (right-looking between panels, look-ahead 1):
//   panel stream : column ranks update their piece of panel J with panel J-1; the owner of block (J, J) factors it
//                  (potrf_diag_kernel + GEMMs); L_JJ and its 128-block inverses go to the other Pr-1 ranks of the process
//                  column (ncclSend / ncclRecv); every column rank solves its rows, piece <- piece L_JJ^-T
//   comm stream  : the Pr pieces of panel J are broadcast to all ranks (one grouped ncclBroadcast per piece) into a
//                  ring of panel buffers, in piece-major order (= the senders' own layout, no packing)
//   main stream  : every rank applies panel J to its remaining pieces: piece(K, p) -= piece(J, p)[rows >= K] L_KJ^T,
//                  one GEMM launch per local panel K
// so the panel chain (update -> factor -> solve -> broadcast) of step J+1 runs while the trailing update of step J keeps
// the tensor pipes busy.  Dependencies are per panel (events), not per stream: the panel stream waits only for the
// main-stream update that touched the piece it needs, and a ring of `nslot` panel buffers lets ranks drift apart by
// nslot - 1 steps, which absorbs the +-1 panel imbalance of the cyclic layout.
#include "g3b_internal.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <algorithm>
#include <random>

namespace {

constexpr int TS = G3_TILE;

// ---- NCCL through dlopen: libg3b.so keeps no link-time dependency on it (single-GPU users never load it) -------------
struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};

NcclApi* nccl_api(std::string* err) {
  static NcclApi api;
  static bool tried = false, ok = false;
  if (!tried) {
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.h) break;
    }
    if (api.h) {
      bool all = true;
#define LOADSYM(field, sym)                                                   \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.h, sym));       \
  all = all && api.field != nullptr;
      LOADSYM(GetUniqueId, "ncclGetUniqueId")
      LOADSYM(CommInitRank, "ncclCommInitRank")
      LOADSYM(CommDestroy, "ncclCommDestroy")
      LOADSYM(Broadcast, "ncclBroadcast")
      LOADSYM(AllReduce, "ncclAllReduce")
      LOADSYM(AllGather, "ncclAllGather")
      LOADSYM(Reduce, "ncclReduce")
      LOADSYM(Send, "ncclSend")
      LOADSYM(Recv, "ncclRecv")
      LOADSYM(GroupStart, "ncclGroupStart")
      LOADSYM(GroupEnd, "ncclGroupEnd")
      LOADSYM(GetErrorString, "ncclGetErrorString")
      LOADSYM(GetVersion, "ncclGetVersion")
#undef LOADSYM
      ok = all;
    }
  }
  if (!ok && err) *err = api.h ? "libnccl.so.2 lacks a required symbol" : "libnccl.so.2 not found (dlopen)";
  return ok ? &api : nullptr;
}

}  // namespace

struct g3_dist {
  int nranks = 1, rank = 0;
  ncclComm_t comm = nullptr;
  cudaStream_t cs = nullptr;                     // collectives
  // ---- layout of the current factorisation
  int N = 0, nb = 0, nP = 0, Pr = 1, Pc = 1, p = 0, q = 0, nslot = 3, w = 0;
  std::vector<long long> off;                    // per panel: element offset of my piece in `store`, -1 if none
  std::vector<long long> dinv_off;               // per panel: offset of its block inverses in `dinv`, -1 if I do not own (J, J)
  double* store = nullptr; size_t store_elems = 0;
  double* ring = nullptr; size_t slot_elems = 0;
  double* diagbuf = nullptr; size_t diag_elems = 0;   // 2 x [L_JJ | Dinv_J] received from the owner of (J, J)
  double* dinv = nullptr;
  double* scal = nullptr;                        // device scalars: [0] logdet, [1] beta, [2] shift, [3] dmin, [4] dmean
  double* dtheta = nullptr;                      // device copy of theta (P) followed by the tt_to_cov shift
  int* info = nullptr;                           // per panel: first bad pivot (+1) inside its diagonal block
  int* st = nullptr;                             // status word (non-finite scrub)
  double* u = nullptr;                           // N: u = L^-1 delta (every rank, after g3_dist_solve)
  double* c = nullptr;                           // N: residual accumulator of the distributed substitution
  std::vector<cudaEvent_t> ev_ready, ev_free_main, ev_free_ps;
  cudaEvent_t ev_upd[4] = {}, ev_diag_fact = nullptr, ev_diag_arr = nullptr, ev_solved[2] = {}, ev_start = nullptr, ev_join = nullptr,
              ev_t[4] = {};
  g3_kernel_desc desc;
  int P = 0;
  bool factored = false, solved = false;
  int lookahead = 1;
};

namespace {

// restores ctx->stream on every exit path of a function that routes launches to other streams
struct stream_guard {
  g3_ctx* c;
  cudaStream_t s;
  explicit stream_guard(g3_ctx* ctx) : c(ctx), s(ctx->stream) {}
  ~stream_guard() { c->stream = s; }
};

inline int first_blk(int J, int p, int Pr) { return J + (((p - J % Pr) % Pr) + Pr) % Pr; }
inline int cnt_blk(int J, int p, int Pr, int nP) {
  const int f = first_blk(J, p, Pr);
  return f < nP ? (nP - 1 - f) / Pr + 1 : 0;
}
inline int rank_of(const g3_dist* d, int p, int q) { return q * d->Pr + p; }
inline size_t blk_elems(const g3_dist* d) { return (size_t)d->nb * d->nb; }

// offset (elements) of piece pp inside a full panel buffer (piece-major order)
inline size_t piece_off_in_panel(const g3_dist* d, int J, int pp) {
  size_t o = 0;
  for (int k = 0; k < pp; ++k) o += (size_t)cnt_blk(J, k, d->Pr, d->nP) * blk_elems(d);
  return o;
}
// where piece (J, pp) lives on THIS rank: its own storage, or the ring slot of panel J
inline double* piece_ptr(const g3_dist* d, int J, int pp) {
  if (J % d->Pc == d->q && pp == d->p) return d->store + d->off[J];
  return d->ring + (size_t)(J % d->nslot) * d->slot_elems + piece_off_in_panel(d, J, pp);
}
// block (K, J), K >= J
inline double* block_ptr(const g3_dist* d, int K, int J) {
  const int pp = K % d->Pr;
  const int idx = (K - first_blk(J, pp, d->Pr)) / d->Pr;
  return piece_ptr(d, J, pp) + (size_t)idx * blk_elems(d);
}

int nccl_fail(g3_ctx* ctx, NcclApi* api, ncclResult_t r, const char* what) {
  char buf[256];
  snprintf(buf, sizeof buf, "%s failed: %s", what, api && api->GetErrorString ? api->GetErrorString(r) : "NCCL error");
  ctx->err = buf;
  return -5;
}
#define G3_NCCL(ctx, api, call)                                     \
  do {                                                              \
    ncclResult_t _r = (call);                                       \
    if (_r != ncclSuccess) return nccl_fail((ctx), (api), _r, #call); \
  } while (0)

// ---- small kernels -----------------------------------------------------------------------------------------------
// Xg[(i nb + r) D + k] = X[((I0 + i Pr) nb + r) D + k]: the rows of one piece, contiguous
__global__ void gather_rows_kernel(const double* __restrict__ X, int D, double* __restrict__ Xg, int cnt, int nb, int I0, int Pr) {
  const long long n = (long long)cnt * nb * D;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / D;
    const int k = (int)(i - row * D);
    const int b = (int)(row / nb), r = (int)(row - (long long)b * nb);
    Xg[i] = X[((long long)(I0 + b * Pr) * nb + r) * D + k];
  }
}

__global__ void shift_kernel(double* scal, double jitter) {  // tt_to_cov: r + (1e-6 - m) eye when min diag <= 0 (tensors.py:95-98)
  if (threadIdx.x == 0) scal[2] = scal[3] > 0.0 ? 0.0 : jitter - scal[3];
}

// One warp per piece row: for NV right-hand columns at once,
//   y[grow][v] (+)= sign * sum_c P[r][c] x[c][v],   grow = (I0 + (r / nb) Pr) nb + r % nb.
// tri_first: the first block of the piece is the diagonal block (J, J): only its lower triangle is L (the tiles above
// hold stale K values).  P: rows x nb, ld = nb.  x: nb x NV (ld NV).  y: N x NV.
template <int NV>
__global__ void __launch_bounds__(256)
piece_gemv_n_kernel(const double* __restrict__ P, int rows, int nb, int I0, int Pr, int tri_first, const double* __restrict__ x,
                    double* __restrict__ y, double sign) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= rows) return;
  const int b = r / nb, rb = r - b * nb;
  const int cmax = (tri_first && b == 0) ? rb + 1 : nb;
  const double* row = P + (long long)r * nb;
  double acc[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) acc[v] = 0.0;
  for (int c = lane; c < cmax; c += 32) {
    const double a = row[c];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] += a * x[(long long)c * NV + v];
  }
#pragma unroll
  for (int v = 0; v < NV; ++v)
    for (int o = 16; o > 0; o >>= 1) acc[v] += __shfl_xor_sync(0xffffffffu, acc[v], o);
  if (lane == 0) {
    const long long g = ((long long)(I0 + b * Pr) * nb + rb) * NV;
#pragma unroll
    for (int v = 0; v < NV; ++v) y[g + v] += sign * acc[v];
  }
}

// z[c][v] += sum_r P[r][c] x[grow(r)][v]  (transposed product; rows split over CTAs, fp64 atomics: probe only)
template <int NV>
__global__ void __launch_bounds__(256)
piece_gemv_t_kernel(const double* __restrict__ P, int rows, int nb, int I0, int Pr, int tri_first, const double* __restrict__ x,
                    double* __restrict__ z) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  const int r0 = blockIdx.y * 128, r1 = min(rows, r0 + 128);
  if (c >= nb) return;
  double acc[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) acc[v] = 0.0;
  for (int r = r0; r < r1; ++r) {
    const int b = r / nb, rb = r - b * nb;
    if (tri_first && b == 0 && c > rb) continue;
    const double a = P[(long long)r * nb + c];
    const long long g = ((long long)(I0 + b * Pr) * nb + rb) * NV;
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] += a * x[g + v];
  }
#pragma unroll
  for (int v = 0; v < NV; ++v) atomicAdd(z + (long long)c * NV + v, acc[v]);
}

// y[row0 + r][v] = sum_c Kc[r][c] x[c][v]   (dense row chunk of K, ld = n)
template <int NV>
__global__ void __launch_bounds__(256)
dense_gemv_kernel(const double* __restrict__ Kc, int rows, long long n, const double* __restrict__ x, double* __restrict__ y) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= rows) return;
  const double* row = Kc + (long long)r * n;
  double acc[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) acc[v] = 0.0;
  for (long long c = lane; c < n; c += 32) {
    const double a = row[c];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] += a * x[c * NV + v];
  }
#pragma unroll
  for (int v = 0; v < NV; ++v)
    for (int o = 16; o > 0; o >>= 1) acc[v] += __shfl_xor_sync(0xffffffffu, acc[v], o);
  if (lane == 0) {
#pragma unroll
    for (int v = 0; v < NV; ++v) y[(long long)r * NV + v] = acc[v];
  }
}


// dst[c][r] = src[r][c] for an (rows x cols) block, 32 x 32 tiles through shared memory (rows, cols multiples of 32)
__global__ void __launch_bounds__(256)
transpose_kernel(const double* __restrict__ src, long long ld_src, double* __restrict__ dst, long long ld_dst, int rows, int cols) {
  __shared__ double tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;      // bx: column block of src, by: row block of src
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8)
    if (by + r < rows && bx + tx < cols) tile[r][tx] = src[(long long)(by + r) * ld_src + bx + tx];
  __syncthreads();
  for (int r = ty; r < 32; r += 8)
    if (bx + r < cols && by + tx < rows) dst[(long long)(bx + r) * ld_dst + by + tx] = tile[tx][r];
}

// deterministic column sums of a tall panel against a vector: part[chunk][c] = sum_{r in chunk} P[r][c] x[r]
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const double* __restrict__ P, int rows, int nb, const double* __restrict__ x, double* __restrict__ part, int chunk_rows) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= nb) return;
  const int r0 = blockIdx.y * chunk_rows, r1 = min(rows, r0 + chunk_rows);
  double acc = 0.0;
  for (int r = r0; r < r1; ++r) acc += P[(long long)r * nb + c] * x[r];
  part[(long long)blockIdx.y * nb + c] = acc;
}
// rhs[c] = u[c] - sum_chunks part[chunk][c]
__global__ void colsum_finish_kernel(const double* __restrict__ part, int nchunk, int nb, const double* __restrict__ u, double* __restrict__ rhs) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nb) return;
  double acc = 0.0;
  for (int k = 0; k < nchunk; ++k) acc += part[(long long)k * nb + c];
  rhs[c] = u[c] - acc;
}
__global__ void identity_kernel(double* __restrict__ A, int n) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < (long long)n * n) A[i] = (i / n == i % n) ? 1.0 : 0.0;
}
// mean[m] += sum_c V[m][c] u[c];  nrm[m] += sum_c V[m][c]^2   (one warp per test point, V: M x nb)
__global__ void __launch_bounds__(256)
post_accum_kernel(const double* __restrict__ V, int M, int nb, const double* __restrict__ u, double* __restrict__ mean, double* __restrict__ nrm) {
  const int m = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (m >= M) return;
  const double* row = V + (long long)m * nb;
  double s1 = 0.0, s2 = 0.0;
  for (int c = lane; c < nb; c += 32) {
    const double v = row[c];
    s1 += v * u[c];
    s2 += v * v;
  }
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if (lane == 0) {
    mean[m] += s1;
    nrm[m] += s2;
  }
}
// R[m][c] = Ks[m][c] - C[m][c]  (M x nb blocks with different leading dimensions)
__global__ void sub_block_kernel(const double* __restrict__ Ks, long long ldk, const double* __restrict__ Cc, long long ldc, double* __restrict__ R,
                                 int M, int nb) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)M * nb) return;
  const long long m = i / nb, c = i - m * nb;
  R[i] = Ks[m * ldk + c] - Cc[m * ldc + c];
}

void dist_free_matrix(g3_dist* d) {
  if (d->store) cudaFree(d->store);
  if (d->ring) cudaFree(d->ring);
  if (d->diagbuf) cudaFree(d->diagbuf);
  if (d->dinv) cudaFree(d->dinv);
  if (d->scal) cudaFree(d->scal);
  if (d->dtheta) cudaFree(d->dtheta);
  if (d->info) cudaFree(d->info);
  if (d->st) cudaFree(d->st);
  if (d->u) cudaFree(d->u);
  if (d->c) cudaFree(d->c);
  d->store = d->ring = d->diagbuf = d->dinv = d->scal = d->dtheta = d->u = d->c = nullptr;
  d->info = d->st = nullptr;
  d->factored = false;
}

g3_dist* dist_get(g3_ctx* ctx) {
  if (!ctx->dist) ctx->dist = new g3_dist();
  return ctx->dist;
}

int dist_events(g3_ctx* ctx, g3_dist* d) {
  auto mk = [&](cudaEvent_t* e, bool timing) -> int {
    if (*e) return 0;
    G3_CUDA(ctx, cudaEventCreateWithFlags(e, timing ? cudaEventDefault : cudaEventDisableTiming));
    return 0;
  };
  int rc;
  if ((int)d->ev_ready.size() < d->nslot) {
    d->ev_ready.resize(d->nslot, nullptr);
    d->ev_free_main.resize(d->nslot, nullptr);
    d->ev_free_ps.resize(d->nslot, nullptr);
  }
  for (int s = 0; s < d->nslot; ++s)
    if ((rc = mk(&d->ev_ready[s], false)) || (rc = mk(&d->ev_free_main[s], false)) || (rc = mk(&d->ev_free_ps[s], false))) return rc;
  for (int k = 0; k < 4; ++k)
    if ((rc = mk(&d->ev_upd[k], false)) || (rc = mk(&d->ev_t[k], true))) return rc;
  for (int k = 0; k < 2; ++k)
    if ((rc = mk(&d->ev_solved[k], false))) return rc;
  if ((rc = mk(&d->ev_diag_fact, false)) || (rc = mk(&d->ev_diag_arr, false)) || (rc = mk(&d->ev_start, false)) ||
      (rc = mk(&d->ev_join, false)))
    return rc;
  if (!d->cs) {
    int lo = 0, hi = 0;
    G3_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    G3_CUDA(ctx, cudaStreamCreateWithPriority(&d->cs, cudaStreamNonBlocking, hi));
  }
  if (!ctx->panel_stream) {
    int lo = 0, hi = 0;
    G3_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    G3_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->panel_stream, cudaStreamNonBlocking, hi));
    G3_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_panel, cudaEventDisableTiming));
    G3_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_main, cudaEventDisableTiming));
  }
  return 0;
}

// K pieces of this rank, generated from the replicated X
int dist_build(g3_ctx* ctx, g3_dist* d) {
  G3_NVTX("g3_dist:gram pieces");
  const int nb = d->nb, D = ctx->D;
  double* Xg = (double*)g3_ws(ctx, "dist_xg", sizeof(double) * (size_t)d->N * D);
  if (!Xg) return -2;
  int rc;
  for (int J = d->q; J < d->nP; J += d->Pc) {
    const int cnt = cnt_blk(J, d->p, d->Pr, d->nP);
    if (cnt == 0) continue;
    const int I0 = first_blk(J, d->p, d->Pr);
    const long long n = (long long)cnt * nb * D;
    gather_rows_kernel<<<(unsigned)std::min<long long>((n + 255) / 256, 4096), 256, 0, ctx->stream>>>(ctx->dX, D, Xg, cnt, nb, I0, d->Pr);
    G3_LAUNCH_CHECK(ctx);
    double* piece = d->store + d->off[J];
    int done = 0;
    if (I0 == J) {  // diagonal block: Noise / shift on its diagonal
      GramArgs a;
      memset(&a, 0, sizeof a);
      a.X1 = Xg; a.X2 = ctx->dX + (size_t)J * nb * D; a.n1 = nb; a.n2 = nb; a.D = D; a.same = 1;
      a.theta = d->dtheta; a.P = d->P; a.diag_shift = d->scal + 2;
      a.K = piece; a.ldk = nb; a.status = d->st;
      if ((rc = g3_gram_launch(ctx, d->desc, a, 1))) return rc;
      done = 1;
    }
    if (cnt - done > 0) {  // off-diagonal blocks: cov(x, x) form (Noise contributes var * I, i.e. nothing here)
      GramArgs a;
      memset(&a, 0, sizeof a);
      a.X1 = Xg + (size_t)done * nb * D; a.X2 = ctx->dX + (size_t)J * nb * D; a.n1 = (cnt - done) * nb; a.n2 = nb; a.D = D;
      a.same = 1; a.diag_off = 1 << 30;
      a.theta = d->dtheta; a.P = d->P;
      a.K = piece + (size_t)done * blk_elems(d); a.ldk = nb; a.status = d->st;
      if ((rc = g3_gram_launch(ctx, d->desc, a, 1))) return rc;
    }
  }
  return 0;
}

int dist_factor(g3_ctx* ctx, g3_dist* d) {
  G3_NVTX("g3_dist:block-cyclic potrf");
  NcclApi* api = d->nranks > 1 ? nccl_api(&ctx->err) : nullptr;
  if (d->nranks > 1 && !api) return -5;
  const int nP = d->nP, nb = d->nb, Pr = d->Pr, Pc = d->Pc, p = d->p, q = d->q, nslot = d->nslot, w = d->w;
  cudaStream_t MS = ctx->stream, PS = ctx->panel_stream, CS = d->cs;
  stream_guard guard(ctx);
  const size_t be = blk_elems(d), de = (size_t)w * TS * TS;
  int rc = 0;
  G3_CUDA(ctx, cudaEventRecord(d->ev_start, MS));
  G3_CUDA(ctx, cudaStreamWaitEvent(PS, d->ev_start, 0));
  G3_CUDA(ctx, cudaStreamWaitEvent(CS, d->ev_start, 0));
  const bool look = d->lookahead != 0;

  // piece(K, p) -= piece(J, p)[rows >= K] * L_KJ^T on the current ctx->stream
  auto update = [&](int K, int J) -> int {
    const int cnt = cnt_blk(K, p, Pr, nP);
    if (cnt == 0) return 0;
    const int IK = first_blk(K, p, Pr);
    const double* A = piece_ptr(d, J, p) + (size_t)((IK - first_blk(J, p, Pr)) / Pr) * be;
    const double* Bm = block_ptr(d, K, J);
    return g3_panel_update(ctx, d->store + d->off[K], cnt * nb, nb, A, Bm, IK == K ? 1 : 0);
  };

  // make panel J final and known to every rank (critical path)
  auto stage = [&](int J) -> int {
    const int qJ = J % Pc, pd = J % Pr, slot = J % nslot;
    const bool in_col = q == qJ;
    const int cnt = cnt_blk(J, p, Pr, nP);
    const bool diag_owner = in_col && p == pd;
    cudaStream_t S = look ? PS : MS;
    ctx->stream = S;
    if (look && in_col && cnt > 0 && J >= 1) {          // without look-ahead trailing(J-1) already applied panel J-1 to this piece
      if (J >= 2) cudaStreamWaitEvent(S, d->ev_upd[J % 4], 0);       // main-stream updates of this piece by panels < J-1
      cudaStreamWaitEvent(S, d->ev_ready[(J - 1) % nslot], 0);       // panel J-1 is here
      if ((rc = update(J, J - 1))) return rc;
      cudaEventRecord(d->ev_free_ps[(J - 1) % nslot], S);
    }
    double* piece = (in_col && cnt > 0) ? d->store + d->off[J] : nullptr;
    const double* Ld = nullptr;
    const double* Dv = nullptr;
    if (diag_owner) {
      double* Dj = d->dinv + d->dinv_off[J];
      if ((rc = g3_potrf_panel(ctx, piece, nb, nb, Dj, d->scal, d->info + J))) return rc;
      Ld = piece;
      Dv = Dj;
      if (Pr > 1) cudaEventRecord(d->ev_diag_fact, S);
    }
    if (Pr > 1 && in_col) {  // L_JJ and its block inverses to the other process rows of this column
      double* buf = d->diagbuf + (size_t)(J % 2) * d->diag_elems;
      if (diag_owner) {
        cudaStreamWaitEvent(CS, d->ev_diag_fact, 0);
      } else {
        cudaStreamWaitEvent(CS, d->ev_solved[J % 2], 0);             // the solve that last read this buffer (panel J - 2 Pc ...)
      }
      G3_NCCL(ctx, api, api->GroupStart());
      if (diag_owner) {
        for (int pp = 0; pp < Pr; ++pp) {
          if (pp == p) continue;
          if (cnt_blk(J, pp, Pr, nP) == 0) continue;
          G3_NCCL(ctx, api, api->Send(Ld, be, ncclDouble, rank_of(d, pp, q), d->comm, CS));
          G3_NCCL(ctx, api, api->Send(Dv, de, ncclDouble, rank_of(d, pp, q), d->comm, CS));
        }
      } else if (cnt > 0) {
        G3_NCCL(ctx, api, api->Recv(buf, be, ncclDouble, rank_of(d, pd, q), d->comm, CS));
        G3_NCCL(ctx, api, api->Recv(buf + be, de, ncclDouble, rank_of(d, pd, q), d->comm, CS));
      }
      G3_NCCL(ctx, api, api->GroupEnd());
      if (!diag_owner && cnt > 0) {
        cudaEventRecord(d->ev_diag_arr, CS);
        cudaStreamWaitEvent(S, d->ev_diag_arr, 0);
        Ld = buf;
        Dv = buf + be;
      }
    }
    if (in_col && cnt > 0) {
      const int skip = diag_owner ? 1 : 0;                           // the diagonal block itself is already L_JJ
      if (cnt - skip > 0 && (rc = g3_panel_solve(ctx, piece + (size_t)skip * be, (cnt - skip) * nb, nb, Ld, Dv))) return rc;
      cudaEventRecord(d->ev_solved[J % 2], S);
    }
    if (d->nranks > 1) {
      if (in_col && cnt > 0) cudaStreamWaitEvent(CS, d->ev_solved[J % 2], 0);
      const bool all_local = Pr == 1 && in_col;
      if (!all_local) {                                              // readers of the panel that used this ring slot before
        cudaStreamWaitEvent(CS, d->ev_free_main[slot], 0);
        cudaStreamWaitEvent(CS, d->ev_free_ps[slot], 0);
      }
      G3_NCCL(ctx, api, api->GroupStart());
      for (int pp = 0; pp < Pr; ++pp) {
        const int c2 = cnt_blk(J, pp, Pr, nP);
        if (c2 == 0) continue;
        double* ptr = piece_ptr(d, J, pp);
        G3_NCCL(ctx, api, api->Broadcast(ptr, ptr, (size_t)c2 * be, ncclDouble, rank_of(d, pp, qJ), d->comm, CS));
      }
      G3_NCCL(ctx, api, api->GroupEnd());
      if (all_local) cudaEventRecord(d->ev_ready[slot], S);          // the owner of a whole panel need not wait for its broadcast
      else cudaEventRecord(d->ev_ready[slot], CS);
    } else {
      cudaEventRecord(d->ev_ready[slot], S);
    }
    ctx->stream = MS;
    return 0;
  };

  // apply panel J to the pieces of this rank (panel J+1 was done in stage(J+1) with look-ahead)
  auto trailing = [&](int J) -> int {
    ctx->stream = MS;
    const int slot = J % nslot;
    bool waited = false;
    int K0 = J + 1;
    while (K0 % Pc != q) ++K0;
    for (int K = K0; K < nP; K += Pc) {
      if (look && K == J + 1) continue;
      if (cnt_blk(K, p, Pr, nP) == 0) continue;
      if (!waited) {
        cudaStreamWaitEvent(MS, d->ev_ready[slot], 0);
        waited = true;
      }
      if ((rc = update(K, J))) return rc;
      if (look && K == J + 2) cudaEventRecord(d->ev_upd[K % 4], MS);
    }
    cudaEventRecord(d->ev_free_main[slot], MS);
    return 0;
  };

  if (look) {
    if ((rc = stage(0))) return rc;
    for (int J = 0; J < nP && !rc; ++J) {
      if (J + 1 < nP) rc = stage(J + 1);
      if (!rc) rc = trailing(J);
    }
  } else {
    for (int J = 0; J < nP && !rc; ++J) {
      // without look-ahead stage(J) runs on the main stream after ALL updates by panel J-1 (update(J, J-1) included there)
      rc = stage(J);
      if (!rc) rc = trailing(J);
    }
  }
  ctx->stream = MS;
  if (rc) return rc;
  G3_CUDA(ctx, cudaEventRecord(d->ev_join, PS));
  G3_CUDA(ctx, cudaStreamWaitEvent(MS, d->ev_join, 0));
  G3_CUDA(ctx, cudaEventRecord(d->ev_join, CS));
  G3_CUDA(ctx, cudaStreamWaitEvent(MS, d->ev_join, 0));
  return 0;
}

}  // namespace

void g3_dist_destroy(g3_ctx* ctx) {
  g3_dist* d = ctx->dist;
  if (!d) return;
  dist_free_matrix(d);
  if (d->comm) {
    NcclApi* api = nccl_api(nullptr);
    if (api) api->CommDestroy(d->comm);
  }
  for (cudaEvent_t e : d->ev_ready) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : d->ev_free_main) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : d->ev_free_ps) if (e) cudaEventDestroy(e);
  for (int k = 0; k < 4; ++k) { if (d->ev_upd[k]) cudaEventDestroy(d->ev_upd[k]); if (d->ev_t[k]) cudaEventDestroy(d->ev_t[k]); }
  for (int k = 0; k < 2; ++k) if (d->ev_solved[k]) cudaEventDestroy(d->ev_solved[k]);
  if (d->ev_diag_fact) cudaEventDestroy(d->ev_diag_fact);
  if (d->ev_diag_arr) cudaEventDestroy(d->ev_diag_arr);
  if (d->ev_start) cudaEventDestroy(d->ev_start);
  if (d->ev_join) cudaEventDestroy(d->ev_join);
  if (d->cs) cudaStreamDestroy(d->cs);
  delete d;
  ctx->dist = nullptr;
}

extern "C" {

int g3_comm_get_unique_id(char* id_out) {
  if (!id_out) return -1;
  std::string err;
  NcclApi* api = nccl_api(&err);
  if (!api) return -5;
  ncclUniqueId id;
  if (api->GetUniqueId(&id) != ncclSuccess) return -5;
  static_assert(sizeof(ncclUniqueId) == G3_COMM_ID_BYTES, "unique id size");
  memcpy(id_out, &id, sizeof id);
  return 0;
}

int g3_comm_init(g3_ctx* ctx, int nranks, int rank, const char* id) {
  if (!ctx || nranks < 1 || rank < 0 || rank >= nranks || (nranks > 1 && !id)) return g3_fail_msg(ctx, "g3_comm_init: bad arguments");
  G3_CUDA(ctx, cudaSetDevice(ctx->device));
  g3_dist* d = dist_get(ctx);
  if (d->comm) return g3_fail_msg(ctx, "g3_comm_init: communicator already initialised on this context");
  d->nranks = nranks;
  d->rank = rank;
  if (nranks == 1) return 0;
  // The collectives here are panel broadcasts that only have to keep up with the trailing updates (tens of GB/s) and tiny
  // reductions; every NCCL channel is a CTA that spins on an SM the fp64 GEMMs could use.  8 channels measured 0.5 % faster
  // than the default on the N=131072 factorisation (profiles/r02g_*); a user setting wins.
  setenv("NCCL_MAX_NCHANNELS", "8", 0);
  NcclApi* api = nccl_api(&ctx->err);
  if (!api) return -5;
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof uid);
  G3_NCCL(ctx, api, api->CommInitRank(&d->comm, nranks, uid, rank));
  return 0;
}

int g3_comm_destroy(g3_ctx* ctx) {
  if (!ctx || !ctx->dist) return 0;
  G3_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStreamSynchronize(ctx->stream);
  g3_dist_destroy(ctx);
  return 0;
}

int g3_comm_size(g3_ctx* ctx) { return ctx && ctx->dist ? ctx->dist->nranks : 1; }
int g3_comm_rank(g3_ctx* ctx) { return ctx && ctx->dist ? ctx->dist->rank : 0; }

// Host-buffer collectives for the small result vectors of the sharded theta batch (8 B (P + 1) bytes per rank) and for
// max-over-ranks timing: staged through a device workspace, one NCCL call on the context's stream, synchronous.
int g3_comm_allgather(g3_ctx* ctx, const void* send, void* recv, size_t bytes_per_rank) {
  if (!ctx || !send || !recv) return g3_fail_msg(ctx, "g3_comm_allgather: bad arguments");
  g3_dist* d = ctx->dist;
  const int n = d ? d->nranks : 1;
  if (n == 1) {
    memcpy(recv, send, bytes_per_rank);
    return 0;
  }
  NcclApi* api = nccl_api(&ctx->err);
  if (!api || !d->comm) return g3_fail_msg(ctx, "g3_comm_allgather: call g3_comm_init first");
  G3_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t pad = (bytes_per_rank + 7) & ~size_t(7);
  char* dv = (char*)g3_ws(ctx, "comm_stage", pad * (n + 1));
  if (!dv) return -2;
  G3_CUDA(ctx, cudaMemcpyAsync(dv, send, bytes_per_rank, cudaMemcpyHostToDevice, ctx->stream));
  G3_NCCL(ctx, api, api->AllGather(dv, dv + pad, pad, ncclChar, d->comm, ctx->stream));
  G3_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int r = 0; r < n; ++r)
    G3_CUDA(ctx, cudaMemcpy((char*)recv + (size_t)r * bytes_per_rank, dv + pad * (r + 1), bytes_per_rank, cudaMemcpyDeviceToHost));
  return 0;
}

int g3_comm_allreduce(g3_ctx* ctx, double* vals, int n, int op /* 0 sum, 1 max, 2 min */) {
  if (!ctx || !vals || n <= 0) return g3_fail_msg(ctx, "g3_comm_allreduce: bad arguments");
  g3_dist* d = ctx->dist;
  if (!d || d->nranks == 1) return 0;
  NcclApi* api = nccl_api(&ctx->err);
  if (!api || !d->comm) return g3_fail_msg(ctx, "g3_comm_allreduce: call g3_comm_init first");
  G3_CUDA(ctx, cudaSetDevice(ctx->device));
  double* dv = (double*)g3_ws(ctx, "comm_red", sizeof(double) * n);
  if (!dv) return -2;
  G3_CUDA(ctx, cudaMemcpyAsync(dv, vals, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  G3_NCCL(ctx, api, api->AllReduce(dv, dv, n, ncclDouble, op == 1 ? ncclMax : (op == 2 ? ncclMin : ncclSum), d->comm, ctx->stream));
  G3_CUDA(ctx, cudaMemcpyAsync(vals, dv, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  G3_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

int g3_comm_barrier(g3_ctx* ctx) {
  if (!ctx) return -1;
  G3_CUDA(ctx, cudaSetDevice(ctx->device));
  G3_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  double one = 1.0;
  return g3_comm_allreduce(ctx, &one, 1, 0);
}

// Pure layout query (no device work): where block (I, J) lives and how large the pieces are.  out[0] = owner rank of
// block (I, J), out[1] = first block row of piece (J, p), out[2] = blocks in piece (J, p), out[3] = local block index of
// (I, J) inside its piece.  Used by the host side and by CPU tests of the index arithmetic.
int g3_dist_layout(int N, int nb, int Pr, int Pc, int I, int J, int p, int* out4) {
  if (!out4 || nb <= 0 || N % nb || Pr < 1 || Pc < 1 || I < J || J < 0 || I >= N / nb || p < 0 || p >= Pr) return -1;
  const int nP = N / nb;
  const int pp = I % Pr;
  out4[0] = (J % Pc) * Pr + pp;
  out4[1] = first_blk(J, p, Pr);
  out4[2] = cnt_blk(J, p, Pr, nP);
  out4[3] = (I - first_blk(J, pp, Pr)) / Pr;
  return 0;
}

int g3_dist_free(g3_ctx* ctx) {
  if (!ctx || !ctx->dist) return 0;
  G3_CUDA(ctx, cudaSetDevice(ctx->device));
  G3_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  dist_free_matrix(ctx->dist);
  return 0;
}

int g3_dist_factor(g3_ctx* ctx, const g3_kernel_desc* desc, const double* theta, int nb, int Pr, int Pc, int flags,
                   double* logdet, int* info, float* ms_gram, float* ms_potrf, double* local_gib) {
  if (!ctx || !desc || !logdet || !info) return g3_fail_msg(ctx, "g3_dist_factor: bad arguments");
  if (!ctx->dX) return g3_fail_msg(ctx, "g3_dist_factor: call g3_set_data first");
  G3_CUDA(ctx, cudaSetDevice(ctx->device));
  g3_dist* d = dist_get(ctx);
  if (Pr < 1 || Pc < 1 || Pr * Pc != d->nranks) return g3_fail_msg(ctx, "g3_dist_factor: Pr * Pc must equal the communicator size");
  if (nb < TS || nb % TS || ctx->N % nb) return g3_fail_msg(ctx, "g3_dist_factor: nb must be a multiple of 128 that divides N");
  int rc = g3_check_desc(ctx, *desc, ctx->D);
  if (rc) return rc;
  dist_free_matrix(d);
  ctx->gp.valid = 0;
  d->N = ctx->N; d->nb = nb; d->nP = ctx->N / nb; d->Pr = Pr; d->Pc = Pc; d->w = nb / TS;
  d->p = d->rank % Pr; d->q = d->rank / Pr;
  d->lookahead = (flags & G3_DIST_NO_LOOKAHEAD) ? 0 : 1;
  d->nslot = (flags & G3_DIST_RING3) ? 3 : 2;
  d->desc = *desc;
  d->P = desc->n_theta;
  if ((rc = dist_events(ctx, d))) return rc;
  // ---- storage
  const size_t be = blk_elems(d), de = (size_t)d->w * TS * TS;
  d->off.assign(d->nP, -1);
  d->dinv_off.assign(d->nP, -1);
  size_t tot = 0, ndiag = 0;
  for (int J = d->q; J < d->nP; J += Pc) {
    const int cnt = cnt_blk(J, d->p, Pr, d->nP);
    if (cnt == 0) continue;
    d->off[J] = (long long)tot;
    tot += (size_t)cnt * be;
    if (J % Pr == d->p) d->dinv_off[J] = (long long)(ndiag++ * de);
  }
  d->store_elems = tot;
  d->slot_elems = (size_t)d->nP * be;
  d->diag_elems = be + de;
  auto alloc = [&](void** p, size_t bytes, const char* what) -> int {
    cudaError_t e = cudaMalloc(p, bytes ? bytes : 256);
    if (e != cudaSuccess) {
      char buf[200];
      snprintf(buf, sizeof buf, "g3_dist_factor: cudaMalloc(%s, %zu bytes) failed: %s", what, bytes, cudaGetErrorString(e));
      ctx->err = buf;
      cudaGetLastError();
      return -2;
    }
    return 0;
  };
  if ((rc = alloc((void**)&d->store, sizeof(double) * tot, "panels"))) return rc;
  if (d->nranks > 1 && (rc = alloc((void**)&d->ring, sizeof(double) * d->slot_elems * d->nslot, "panel ring"))) return rc;
  if (Pr > 1 && (rc = alloc((void**)&d->diagbuf, sizeof(double) * 2 * d->diag_elems, "diag buffers"))) return rc;
  if ((rc = alloc((void**)&d->dinv, sizeof(double) * (ndiag ? ndiag : 1) * de, "block inverses"))) return rc;
  if ((rc = alloc((void**)&d->scal, sizeof(double) * 8, "scalars"))) return rc;
  if ((rc = alloc((void**)&d->dtheta, sizeof(double) * (d->P + 1), "theta"))) return rc;
  if ((rc = alloc((void**)&d->info, sizeof(int) * d->nP, "info"))) return rc;
  if ((rc = alloc((void**)&d->st, sizeof(int) * 4, "status"))) return rc;
  if ((rc = alloc((void**)&d->u, sizeof(double) * d->N, "u"))) return rc;
  if ((rc = alloc((void**)&d->c, sizeof(double) * d->N, "c"))) return rc;
  cudaStream_t MS = ctx->stream;
  G3_CUDA(ctx, cudaMemsetAsync(d->scal, 0, sizeof(double) * 8, MS));
  G3_CUDA(ctx, cudaMemsetAsync(d->info, 0, sizeof(int) * d->nP, MS));
  G3_CUDA(ctx, cudaMemsetAsync(d->st, 0, sizeof(int) * 4, MS));
  if (d->P > 0) G3_CUDA(ctx, cudaMemcpyAsync(d->dtheta, theta, sizeof(double) * d->P, cudaMemcpyHostToDevice, MS));
  // tt_to_cov: min of the diagonal decides the shift (libs/tensors.py:95-98); every rank computes it (O(N))
  if ((rc = g3_gram_diag_min(ctx, *desc, ctx->dX, d->N, ctx->D, d->dtheta, d->P, 1, d->scal + 3, d->scal + 4, d->st, 0))) return rc;
  shift_kernel<<<1, 32, 0, MS>>>(d->scal, ctx->jitter_rel);
  G3_LAUNCH_CHECK(ctx);
  if ((rc = g3_comm_barrier(ctx))) return rc;              // common start: times are comparable across ranks
  G3_CUDA(ctx, cudaEventRecord(d->ev_t[0], MS));
  if ((rc = dist_build(ctx, d))) return rc;
  G3_CUDA(ctx, cudaEventRecord(d->ev_t[1], MS));
  if ((rc = dist_factor(ctx, d))) return rc;
  G3_CUDA(ctx, cudaEventRecord(d->ev_t[2], MS));
  G3_CUDA(ctx, cudaStreamSynchronize(MS));
  float t_gram = 0.f, t_potrf = 0.f;
  G3_CUDA(ctx, cudaEventElapsedTime(&t_gram, d->ev_t[0], d->ev_t[1]));
  G3_CUDA(ctx, cudaEventElapsedTime(&t_potrf, d->ev_t[1], d->ev_t[2]));
  // results: log-det is the sum over the diagonal-block owners; times are the max over ranks
  std::vector<int> hinfo(d->nP);
  double hs[8];
  G3_CUDA(ctx, cudaMemcpy(hs, d->scal, sizeof hs, cudaMemcpyDeviceToHost));
  G3_CUDA(ctx, cudaMemcpy(hinfo.data(), d->info, sizeof(int) * d->nP, cudaMemcpyDeviceToHost));
  double first_bad = 0.0;
  for (int J = d->nP - 1; J >= 0; --J)
    if (hinfo[J]) first_bad = (double)J * nb + hinfo[J];
  double red[2] = {hs[0], 0.0}, mx[3] = {t_gram, t_potrf, first_bad > 0 ? -first_bad : -1e300};
  if ((rc = g3_comm_allreduce(ctx, red, 1, 0))) return rc;
  if ((rc = g3_comm_allreduce(ctx, mx, 3, 1))) return rc;   // max of -first_bad = the smallest failing index
  *logdet = red[0];
  *info = mx[2] > -1e299 ? (int)(-mx[2]) : 0;
  if (ms_gram) *ms_gram = (float)mx[0];
  if (ms_potrf) *ms_potrf = (float)mx[1];
  if (local_gib) *local_gib = (double)(sizeof(double) * (tot + (d->ring ? d->slot_elems * d->nslot : 0))) / (double)(1ull << 30);
  d->factored = true;
  d->solved = false;
  return 0;
}

// u = L^-1 delta across the ranks, beta = |u|^2.  Every rank keeps a length-N residual accumulator c (rank 0 starts with
// delta); the true residual of block row J is the sum over ranks of c[J], one nb-sized all-reduce; the owner of (J, J)
// solves with the diagonal block, u_J goes to every rank (nb doubles), and the column ranks push L_IJ u_J into their c.
// No panel data moves; L is read once.
int g3_dist_solve(g3_ctx* ctx, const double* delta, double* beta_out, double* u_out_or_NULL, float* ms) {
  G3_NVTX("g3_dist_solve");
  if (!ctx || !delta || !beta_out) return g3_fail_msg(ctx, "g3_dist_solve: bad arguments");
  g3_dist* d = ctx->dist;
  if (!d || !d->factored) return g3_fail_msg(ctx, "g3_dist_solve: call g3_dist_factor first");
  G3_CUDA(ctx, cudaSetDevice(ctx->device));
  NcclApi* api = d->nranks > 1 ? nccl_api(&ctx->err) : nullptr;
  if (d->nranks > 1 && !api) return -5;
  const int nP = d->nP, nb = d->nb, Pr = d->Pr, Pc = d->Pc, p = d->p, q = d->q;
  cudaStream_t MS = ctx->stream;
  int rc;
  if ((rc = g3_comm_barrier(ctx))) return rc;
  G3_CUDA(ctx, cudaEventRecord(d->ev_t[0], MS));
  if (d->rank == 0) G3_CUDA(ctx, cudaMemcpyAsync(d->c, delta, sizeof(double) * d->N, cudaMemcpyHostToDevice, MS));
  else G3_CUDA(ctx, cudaMemsetAsync(d->c, 0, sizeof(double) * d->N, MS));
  G3_CUDA(ctx, cudaMemsetAsync(d->scal + 1, 0, sizeof(double), MS));
  for (int J = 0; J < nP; ++J) {
    const int qJ = J % Pc, pd = J % Pr;
    double* seg = d->c + (size_t)J * nb;
    double* uJ = d->u + (size_t)J * nb;
    if (d->nranks > 1) G3_NCCL(ctx, api, api->AllReduce(seg, seg, nb, ncclDouble, ncclSum, d->comm, MS));
    const bool in_col = q == qJ;
    const int cnt = cnt_blk(J, p, Pr, nP);
    if (in_col && p == pd) {  // diagonal block on top of my piece: u_J = L_JJ^-1 r_J, beta += |u_J|^2
      if ((rc = g3_trsv_panel(ctx, d->store + d->off[J], nb, nb, d->dinv + d->dinv_off[J], seg, uJ, d->scal + 1))) return rc;
    }
    if (d->nranks > 1) G3_NCCL(ctx, api, api->Broadcast(uJ, uJ, nb, ncclDouble, rank_of(d, pd, qJ), d->comm, MS));
    if (in_col && cnt > 0) {
      const int skip = p == pd ? 1 : 0;
      const int rows = (cnt - skip) * nb;
      if (rows > 0) {
        const int I0 = first_blk(J, p, Pr) + skip * Pr;
        piece_gemv_n_kernel<1><<<(rows + 7) / 8, 256, 0, MS>>>(d->store + d->off[J] + (size_t)skip * blk_elems(d), rows, nb, I0, Pr, 0,
                                                               uJ, d->c, -1.0);
        G3_LAUNCH_CHECK(ctx);
      }
    }
  }
  G3_CUDA(ctx, cudaEventRecord(d->ev_t[1], MS));
  double beta = 0.0;
  G3_CUDA(ctx, cudaMemcpyAsync(&beta, d->scal + 1, sizeof(double), cudaMemcpyDeviceToHost, MS));
  if (u_out_or_NULL) G3_CUDA(ctx, cudaMemcpyAsync(u_out_or_NULL, d->u, sizeof(double) * d->N, cudaMemcpyDeviceToHost, MS));
  G3_CUDA(ctx, cudaStreamSynchronize(MS));
  float t = 0.f;
  G3_CUDA(ctx, cudaEventElapsedTime(&t, d->ev_t[0], d->ev_t[1]));
  double red[1] = {beta}, mx[1] = {t};
  if ((rc = g3_comm_allreduce(ctx, red, 1, 0))) return rc;
  if ((rc = g3_comm_allreduce(ctx, mx, 1, 1))) return rc;
  *beta_out = red[0];
  if (ms) *ms = (float)mx[0];
  d->solved = true;
  return 0;
}

// Residual probe of the distributed factor: nvec (<= 4) seeded +-1 vectors v, w = L (L^T v) from the distributed
// pieces against K v with K regenerated from X row chunk by row chunk; rel_err[v] = max|w - K v| / max|K v|.
int g3_dist_residual(g3_ctx* ctx, int nvec, unsigned seed, double* rel_err) {
  if (!ctx || !rel_err || nvec < 1 || nvec > 4) return g3_fail_msg(ctx, "g3_dist_residual: 1 <= nvec <= 4");
  g3_dist* d = ctx->dist;
  if (!d || !d->factored) return g3_fail_msg(ctx, "g3_dist_residual: call g3_dist_factor first");
  G3_CUDA(ctx, cudaSetDevice(ctx->device));
  NcclApi* api = d->nranks > 1 ? nccl_api(&ctx->err) : nullptr;
  if (d->nranks > 1 && !api) return -5;
  const int N = d->N, nb = d->nb, nP = d->nP, Pr = d->Pr, Pc = d->Pc, p = d->p, q = d->q;
  constexpr int NV = 4;
  cudaStream_t MS = ctx->stream;
  double* V = (double*)g3_ws(ctx, "dist_probe", sizeof(double) * (size_t)N * NV * 4);
  if (!V) return -2;
  double *Z = V + (size_t)N * NV, *W = Z + (size_t)N * NV, *Y = W + (size_t)N * NV;
  std::vector<double> hv((size_t)N * NV, 0.0);
  std::mt19937_64 gen(seed);
  for (int i = 0; i < N; ++i)
    for (int v = 0; v < nvec; ++v) hv[(size_t)i * NV + v] = (gen() & 1) ? 1.0 : -1.0;
  G3_CUDA(ctx, cudaMemcpyAsync(V, hv.data(), sizeof(double) * (size_t)N * NV, cudaMemcpyHostToDevice, MS));
  G3_CUDA(ctx, cudaMemsetAsync(Z, 0, sizeof(double) * (size_t)N * NV * 3, MS));
  int rc;
  // z = L^T v
  for (int J = q; J < nP; J += Pc) {
    const int cnt = cnt_blk(J, p, Pr, nP);
    if (cnt == 0) continue;
    const int I0 = first_blk(J, p, Pr), rows = cnt * nb;
    piece_gemv_t_kernel<NV><<<dim3((nb + 255) / 256, (rows + 127) / 128), 256, 0, MS>>>(d->store + d->off[J], rows, nb, I0, Pr, I0 == J,
                                                                                     V, Z + (size_t)J * nb * NV);
    G3_LAUNCH_CHECK(ctx);
  }
  if (d->nranks > 1) G3_NCCL(ctx, api, api->AllReduce(Z, Z, (size_t)N * NV, ncclDouble, ncclSum, d->comm, MS));
  // w = L z
  for (int J = q; J < nP; J += Pc) {
    const int cnt = cnt_blk(J, p, Pr, nP);
    if (cnt == 0) continue;
    const int I0 = first_blk(J, p, Pr), rows = cnt * nb;
    piece_gemv_n_kernel<NV><<<(rows + 7) / 8, 256, 0, MS>>>(d->store + d->off[J], rows, nb, I0, Pr, I0 == J, Z + (size_t)J * nb * NV, W, 1.0);
    G3_LAUNCH_CHECK(ctx);
  }
  if (d->nranks > 1) G3_NCCL(ctx, api, api->AllReduce(W, W, (size_t)N * NV, ncclDouble, ncclSum, d->comm, MS));
  // y = K v, block rows dealt round-robin to the ranks; K regenerated chunk by chunk (nb x N) from X
  double* Kc = d->ring ? d->ring : (double*)g3_ws(ctx, "dist_kchunk", sizeof(double) * (size_t)nb * N);
  if (!Kc) return -2;
  for (int I = d->rank; I < nP; I += d->nranks) {
    GramArgs a;
    memset(&a, 0, sizeof a);
    a.X1 = ctx->dX + (size_t)I * nb * ctx->D; a.X2 = ctx->dX; a.n1 = nb; a.n2 = N; a.D = ctx->D; a.same = 1; a.diag_off = I * nb;
    a.theta = d->dtheta; a.P = d->P; a.diag_shift = d->scal + 2;
    a.K = Kc; a.ldk = N;
    if ((rc = g3_gram_launch(ctx, d->desc, a, 1))) return rc;
    dense_gemv_kernel<NV><<<(nb + 7) / 8, 256, 0, MS>>>(Kc, nb, N, V, Y + (size_t)I * nb * NV);
    G3_LAUNCH_CHECK(ctx);
  }
  if (d->nranks > 1) G3_NCCL(ctx, api, api->AllReduce(Y, Y, (size_t)N * NV, ncclDouble, ncclSum, d->comm, MS));
  std::vector<double> hw((size_t)N * NV), hy((size_t)N * NV);
  G3_CUDA(ctx, cudaMemcpyAsync(hw.data(), W, sizeof(double) * (size_t)N * NV, cudaMemcpyDeviceToHost, MS));
  G3_CUDA(ctx, cudaMemcpyAsync(hy.data(), Y, sizeof(double) * (size_t)N * NV, cudaMemcpyDeviceToHost, MS));
  G3_CUDA(ctx, cudaStreamSynchronize(MS));
  for (int v = 0; v < nvec; ++v) {
    double num = 0.0, den = 0.0;
    for (int i = 0; i < N; ++i) {
      num = fmax(num, fabs(hw[(size_t)i * NV + v] - hy[(size_t)i * NV + v]));
      den = fmax(den, fabs(hy[(size_t)i * NV + v]));
    }
    rel_err[v] = (num == num && den > 0.0) ? num / den : INFINITY;
  }
  return 0;
}


// Posterior moments of the distributed exact GP at M test points (elliptical.py:78-107 semantics, Cholesky route):
//   mean = K* K^-1 delta = V u,  var = max(k** - |V_m|^2, 0),  V = K* L^-T
// by blocked forward substitution over the panels: the owner of panel J turns the true residual of its column block into
// V_J = R_J L_JJ^-T, folds V_J u_J and |V_J|^2 into the running moments and pushes V_J L_IJ^T (I > J) into its own
// accumulator; the residual of the next block is the sum of the accumulators over the ranks (one reduce of Mc x nb per
// step).  K* blocks are generated on the fly by the owner.  Needs g3_dist_factor + g3_dist_solve (for u) on a 1 x G grid.
int g3_dist_posterior(g3_ctx* ctx, const double* Xs, int M, int flags, double* mean_out, double* var_out) {
  G3_NVTX("g3_dist_posterior");
  if (!ctx || !Xs || M <= 0 || !mean_out || !var_out) return g3_fail_msg(ctx, "g3_dist_posterior: bad arguments");
  g3_dist* d = ctx->dist;
  if (!d || !d->factored || !d->solved) return g3_fail_msg(ctx, "g3_dist_posterior: call g3_dist_factor and g3_dist_solve first");
  if (d->Pr != 1) return g3_fail_msg(ctx, "g3_dist_posterior: needs a 1 x G process grid");
  G3_CUDA(ctx, cudaSetDevice(ctx->device));
  NcclApi* api = d->nranks > 1 ? nccl_api(&ctx->err) : nullptr;
  if (d->nranks > 1 && !api) return -5;
  const int N = d->N, nb = d->nb, nP = d->nP, G = d->nranks, D = ctx->D;
  const int Mc = std::min(g3_pad(M), 4096);
  cudaStream_t MS = ctx->stream;
  const int skip_pn = (flags & G3_POST_NOISE) ? 0 : 1;
  double* dXs = (double*)g3_ws(ctx, "dp_xs", sizeof(double) * (size_t)g3_pad(M) * D);
  double* C = (double*)g3_ws(ctx, "dp_c", sizeof(double) * (size_t)Mc * N);
  double* R = (double*)g3_ws(ctx, "dp_r", sizeof(double) * (size_t)Mc * nb * 4);      // 2 x [V_J | packed reduce block]
  const size_t be = blk_elems(d);
  double* mom = (double*)g3_ws(ctx, "dp_mom", sizeof(double) * (size_t)g3_pad(M) * 3);
  double* kss = (double*)g3_ws(ctx, "dp_kss", sizeof(double) * 4);
  if (!dXs || !C || !R || !mom || !kss) return -2;
  const int Mp = g3_pad(M);
  double *mean = mom, *nrm = mom + Mp, *kvec = mom + 2 * (size_t)Mp;
  G3_CUDA(ctx, cudaMemsetAsync(dXs, 0, sizeof(double) * (size_t)Mp * D, MS));
  G3_CUDA(ctx, cudaMemcpyAsync(dXs, Xs, sizeof(double) * (size_t)M * D, cudaMemcpyHostToDevice, MS));
  G3_CUDA(ctx, cudaMemsetAsync(mom, 0, sizeof(double) * (size_t)Mp * 3, MS));
  int rc;
  if ((rc = g3_gram_diag_min(ctx, d->desc, dXs, M, D, d->dtheta, d->P, 1, kss, kss + 1, nullptr, skip_pn, kvec))) return rc;
  // The owner's push V_J L_IJ^T into its accumulator is split into three launches on a background stream: the next column
  // block (needed by the very next step), the blocks up to this rank's next own panel, and the rest.  A step only waits for
  // the piece of this rank's latest panel that contains its column block (the stream is in order, so earlier panels are
  // done too), so the pushes of G - 1 ranks overlap with the reduce / solve chain of the current owner.
  // The chain runs on the HIGH-priority stream and the pushes on the context's own (lowest-priority) stream: pending CTAs of
  // the chain (NCCL reduce included) are scheduled ahead of the thousands of queued GEMM CTAs of a push.
  stream_guard guard(ctx);
  cudaStream_t BS = ctx->stream;
  cudaStream_t CH = ctx->panel_stream;
  G3_CUDA(ctx, cudaEventRecord(d->ev_start, BS));
  G3_CUDA(ctx, cudaStreamWaitEvent(CH, d->ev_start, 0));
  MS = CH;
  ctx->stream = CH;
  cudaEvent_t ev_piece[3] = {d->ev_upd[0], d->ev_upd[1], d->ev_upd[2]}, ev_main = d->ev_upd[3];
  for (int m0 = 0; m0 < M; m0 += Mc) {
    const int mc = std::min(Mc, Mp - m0);                // rows of this chunk (multiple of 128; padding rows are zero inputs)
    const int mreal = std::min(mc, M - m0);
    G3_CUDA(ctx, cudaEventRecord(d->ev_join, BS));       // pushes of the previous chunk still read / write C and R
    G3_CUDA(ctx, cudaStreamWaitEvent(MS, d->ev_join, 0));
    G3_CUDA(ctx, cudaMemsetAsync(C, 0, sizeof(double) * (size_t)mc * N, MS));
    int Jm = -1;                                         // my latest own panel
    for (int J = 0; J < nP; ++J) {
      const int owner = J % G;
      const bool mine = owner == d->rank;
      if (Jm >= 0) {                                     // my contributions to column block J are complete?
        const int piece = J == Jm + 1 ? 0 : (J < Jm + G ? 1 : 2);
        G3_CUDA(ctx, cudaStreamWaitEvent(MS, ev_piece[piece], 0));
      }
      double* Rj = R + (size_t)((J / G) & 1) * Mc * nb * 2;     // V_J buffer, alternating between this rank's own panels
      double* Sj = Rj + (size_t)Mc * nb;
      // pack my accumulator block J (strided in C) for the sum over the ranks onto the owner
      if (G > 1)
        G3_CUDA(ctx, cudaMemcpy2DAsync(Sj, sizeof(double) * nb, C + (size_t)J * nb, sizeof(double) * N, sizeof(double) * nb, (size_t)mc,
                                       cudaMemcpyDeviceToDevice, MS));
      if (mine) {
        GramArgs a;                                      // K*[chunk, J] (cross form: Noise contributes zeros, kernels.py:367-371)
        memset(&a, 0, sizeof a);
        a.X1 = dXs + (size_t)m0 * D; a.X2 = ctx->dX + (size_t)J * nb * D; a.n1 = mreal; a.n2 = nb; a.D = D; a.same = 0;
        a.skip_process_noise = skip_pn;
        a.theta = d->dtheta; a.P = d->P; a.K = Rj; a.ldk = nb;
        if (mreal < mc) G3_CUDA(ctx, cudaMemsetAsync(Rj, 0, sizeof(double) * (size_t)mc * nb, MS));
        if ((rc = g3_gram_launch(ctx, d->desc, a, 1))) return rc;
      }
      if (G > 1) G3_NCCL(ctx, api, api->Reduce(Sj, Sj, (size_t)mc * nb, ncclDouble, ncclSum, owner, d->comm, MS));
      if (!mine) continue;
      const double* acc = G > 1 ? Sj : C + (size_t)J * nb;
      sub_block_kernel<<<(unsigned)(((size_t)mc * nb + 255) / 256), 256, 0, MS>>>(Rj, nb, acc, G > 1 ? nb : N, Rj, mc, nb);
      G3_LAUNCH_CHECK(ctx);
      double* piece = d->store + d->off[J];
      if ((rc = g3_panel_solve(ctx, Rj, mc, nb, piece, d->dinv + d->dinv_off[J]))) return rc;     // V_J = R_J L_JJ^-T
      post_accum_kernel<<<(mreal + 7) / 8, 256, 0, MS>>>(Rj, mreal, nb, d->u + (size_t)J * nb, mean + m0, nrm + m0);
      G3_LAUNCH_CHECK(ctx);
      // pushes on the background stream, in column order
      G3_CUDA(ctx, cudaEventRecord(ev_main, MS));
      G3_CUDA(ctx, cudaStreamWaitEvent(BS, ev_main, 0));
      ctx->stream = BS;
      const int c0[3] = {J + 1, J + 2, std::max(J + G, J + 2)};
      const int c1[3] = {std::min(J + 2, nP), std::min(std::max(J + G, J + 2), nP), nP};
      for (int k = 0; k < 3 && !rc; ++k) {
        if (c1[k] > c0[k])
          rc = g3_panel_gemm(ctx, C + (size_t)c0[k] * nb, N, mc, (c1[k] - c0[k]) * nb, Rj, nb, piece + (size_t)(c0[k] - J) * be, nb, nb, 1.0, 1.0);
        cudaEventRecord(ev_piece[k], BS);
      }
      ctx->stream = MS;
      if (rc) { ctx->stream = BS; return rc; }
      Jm = J;
    }
  }
  ctx->stream = BS;                                      // back on the context's stream, after both are done
  G3_CUDA(ctx, cudaEventRecord(d->ev_join, CH));
  G3_CUDA(ctx, cudaStreamWaitEvent(BS, d->ev_join, 0));
  MS = BS;
  std::vector<double> h((size_t)Mp * 3);
  double hk[4];
  G3_CUDA(ctx, cudaMemcpyAsync(h.data(), mom, sizeof(double) * (size_t)Mp * 3, cudaMemcpyDeviceToHost, MS));
  G3_CUDA(ctx, cudaMemcpyAsync(hk, kss, sizeof hk, cudaMemcpyDeviceToHost, MS));
  G3_CUDA(ctx, cudaStreamSynchronize(MS));
  std::vector<double> red(2 * (size_t)M);
  for (int m = 0; m < M; ++m) { red[m] = h[m]; red[M + m] = h[Mp + m]; }
  if ((rc = g3_comm_allreduce(ctx, red.data(), 2 * M, 0))) return rc;
  const double shift = ((flags & G3_POST_NOISE) && !(hk[0] > 0.0)) ? ctx->jitter_rel - hk[0] : 0.0;    // tt_to_cov on K** (elliptical.py:70)
  for (int m = 0; m < M; ++m) {
    mean_out[m] = red[m];
    const double v = (h[2 * (size_t)Mp + m] + shift) - red[M + m];
    var_out[m] = v < 0.0 ? 0.0 : v;                      // tt_to_bounded(.., 0)  elliptical.py:94-97
  }
  return 0;
}

// Gradient of the distributed exact GP (SURVEY §8 a10 on the block-cyclic layout, 1 x G grid):
//   alpha = L^-T u (backward substitution over the panels),  X = L^-1 in place (right-looking, panel by panel from the last:
//   X[J+1:, J] = -(X[J+1:, J+1:] L[J+1:, J]) L_JJ^-1, the product summed over the ranks' own column panels and reduced onto
//   the owner),  K^-1[I][J] = sum_{M >= I} X[M, I]^T X[M, J] for every block pair I >= J (X panels broadcast one by one, the
//   ranks pair them with their own panels, both transposed to K-contiguous form), each block contracted at once with
//   dK[I][J]/dtheta regenerated from X (gram_vjp) - K^-1 is never stored:
//   dtheta[p] = 1/2 sum_ij (c alpha_i alpha_j - K^-1_ij) dK_ij/dtheta_p,  ddelta = -c alpha.
// The factor is consumed (L is overwritten by L^-1).  Device times (max over ranks) in ms3 = {alpha, inverse, contraction}.
int g3_dist_grad(g3_ctx* ctx, double cfac, double* dtheta_out, double* ddelta_out_or_NULL, float* ms3) {
  G3_NVTX("g3_dist_grad");
  if (!ctx || !dtheta_out) return g3_fail_msg(ctx, "g3_dist_grad: bad arguments");
  g3_dist* d = ctx->dist;
  if (!d || !d->factored || !d->solved) return g3_fail_msg(ctx, "g3_dist_grad: call g3_dist_factor and g3_dist_solve first");
  if (d->Pr != 1) return g3_fail_msg(ctx, "g3_dist_grad: needs a 1 x G process grid");
  G3_CUDA(ctx, cudaSetDevice(ctx->device));
  NcclApi* api = d->nranks > 1 ? nccl_api(&ctx->err) : nullptr;
  if (d->nranks > 1 && !api) return -5;
  const int N = d->N, nb = d->nb, nP = d->nP, G = d->nranks, w = d->w, P = d->P, D = ctx->D;
  const size_t be = blk_elems(d), de = (size_t)w * TS * TS;
  cudaStream_t MS = ctx->stream;
  int nloc = 0;
  for (int J = d->rank; J < nP; J += G) ++nloc;
  const int nchunk = 64;
  double* alpha = (double*)g3_ws(ctx, "dg_alpha", sizeof(double) * ((size_t)N + nb + 8));
  double* part = (double*)g3_ws(ctx, "dg_part", sizeof(double) * (size_t)nchunk * nb);
  double* BT = (double*)g3_ws(ctx, "dg_bt", sizeof(double) * (size_t)N * nb);
  double* Y = (double*)g3_ws(ctx, "dg_y", sizeof(double) * (size_t)N * nb);
  double* LT = (double*)g3_ws(ctx, "dg_lt", sizeof(double) * (2 * be + de));
  if (!alpha || !part || !BT || !Y || !LT) return -2;
  double* rhs = alpha + N;                               // nb scratch
  double* cf_dev = alpha + N + nb;
  double* EYE = LT + be;                                 // nb x nb
  double* DT = LT + 2 * be;                              // transposed 128-block inverses
  int rc;
  d->factored = false;                                   // L is consumed
  if ((rc = g3_comm_barrier(ctx))) return rc;
  G3_CUDA(ctx, cudaMemcpyAsync(cf_dev, &cfac, sizeof(double), cudaMemcpyHostToDevice, MS));
  G3_CUDA(ctx, cudaEventRecord(d->ev_t[0], MS));
  // ---- alpha = L^-T u
  for (int J = nP - 1; J >= 0; --J) {
    const int owner = J % G;
    double* aJ = alpha + (size_t)J * nb;
    if (owner == d->rank) {
      double* piece = d->store + d->off[J];
      const int below = (nP - 1 - J) * nb;
      int nch = 0;
      if (below > 0) {
        const int chunk_rows = (below + nchunk - 1) / nchunk;
        nch = (below + chunk_rows - 1) / chunk_rows;
        colsum_partial_kernel<<<dim3((nb + 255) / 256, nch), 256, 0, MS>>>(piece + be, below, nb, alpha + (size_t)(J + 1) * nb, part, chunk_rows);
        G3_LAUNCH_CHECK(ctx);
      }
      colsum_finish_kernel<<<(nb + 255) / 256, 256, 0, MS>>>(part, nch, nb, d->u + (size_t)J * nb, rhs);
      G3_LAUNCH_CHECK(ctx);
      if ((rc = g3_trsv_bwd(ctx, piece, d->dinv + d->dinv_off[J], rhs, aJ, nb, 1))) return rc;
    }
    if (G > 1) G3_NCCL(ctx, api, api->Broadcast(aJ, aJ, nb, ncclDouble, owner, d->comm, MS));
  }
  G3_CUDA(ctx, cudaEventRecord(d->ev_t[1], MS));
  // ---- X = L^-1 in place, panels from the last to the first
  for (int J = nP - 1; J >= 0; --J) {
    const int owner = J % G;
    const bool mine = owner == d->rank;
    const int below = (nP - 1 - J) * nb;
    double* piece = mine ? d->store + d->off[J] : nullptr;
    if (below > 0) {
      if (mine) {  // BT[K] = L[K, J]^T for K > J
        for (int k = 0; k < nP - 1 - J; ++k) {
          transpose_kernel<<<dim3(nb / 32, nb / 32), 256, 0, MS>>>(piece + (size_t)(k + 1) * be, nb, BT + (size_t)k * be, nb, nb, nb);
          G3_LAUNCH_CHECK(ctx);
        }
      }
      if (G > 1) G3_NCCL(ctx, api, api->Broadcast(BT, BT, (size_t)below * nb, ncclDouble, owner, d->comm, MS));
      G3_CUDA(ctx, cudaMemsetAsync(Y, 0, sizeof(double) * (size_t)below * nb, MS));
      int K0 = J + 1;
      while (K0 % G != d->rank) ++K0;
      for (int K = K0; K < nP; K += G) {                 // Y[I >= K] += X[I, K] L[K, J]
        const int rows = (nP - K) * nb;
        if ((rc = g3_panel_gemm(ctx, Y + (size_t)(K - J - 1) * be, nb, rows, nb, d->store + d->off[K], nb, BT + (size_t)(K - J - 1) * be, nb, nb,
                                1.0, 1.0)))
          return rc;
      }
      if (G > 1) G3_NCCL(ctx, api, api->Reduce(Y, Y, (size_t)below * nb, ncclDouble, ncclSum, owner, d->comm, MS));
    }
    if (mine) {
      const double* Dj = d->dinv + d->dinv_off[J];
      transpose_kernel<<<dim3(nb / 32, nb / 32), 256, 0, MS>>>(piece, nb, LT, nb, nb, nb);          // L_JJ^T (stale upper tiles of
      G3_LAUNCH_CHECK(ctx);                                                                          //  L_JJ land below: never read)
      for (int t = 0; t < w; ++t) {
        transpose_kernel<<<dim3(TS / 32, TS / 32), 256, 0, MS>>>(Dj + (size_t)t * TS * TS, TS, DT + (size_t)t * TS * TS, TS, TS, TS);
        G3_LAUNCH_CHECK(ctx);
      }
      if (below > 0) {                                   // X[J+1:, J] = -Y L_JJ^-1
        if ((rc = g3_panel_rsolve(ctx, Y, below, nb, LT, DT, 1.0, -1.0))) return rc;
        G3_CUDA(ctx, cudaMemcpyAsync(piece + be, Y, sizeof(double) * (size_t)below * nb, cudaMemcpyDeviceToDevice, MS));
      }
      identity_kernel<<<(unsigned)((be + 255) / 256), 256, 0, MS>>>(EYE, nb);                       // X_JJ = L_JJ^-1 (exact zeros above)
      G3_LAUNCH_CHECK(ctx);
      if ((rc = g3_panel_rsolve(ctx, EYE, nb, nb, LT, DT, -1.0, 1.0))) return rc;
      G3_CUDA(ctx, cudaMemcpyAsync(piece, EYE, sizeof(double) * be, cudaMemcpyDeviceToDevice, MS));
    }
  }
  G3_CUDA(ctx, cudaEventRecord(d->ev_t[2], MS));
  // ---- K^-1 blocks and their contraction with dK/dtheta
  double* XT = (double*)g3_ws(ctx, "dg_xt", sizeof(double) * (size_t)std::max(nloc, 1) * nb * N);
  double* XTb = (double*)g3_ws(ctx, "dg_xtb", sizeof(double) * (size_t)nb * N);
  double* Kb = (double*)g3_ws(ctx, "dg_kblk", sizeof(double) * (size_t)nb * std::max(nloc, 1) * nb);
  const size_t npairs_max = (size_t)nloc * nP;
  double* outp = (double*)g3_ws(ctx, "dg_out", sizeof(double) * std::max<size_t>(npairs_max, 1) * std::max(P, 1));
  if (!XT || !XTb || !Kb || !outp) return -2;
  G3_CUDA(ctx, cudaMemsetAsync(XT, 0, sizeof(double) * (size_t)std::max(nloc, 1) * nb * N, MS));
  for (int x = 0; x < nloc; ++x) {                       // XT[x][n][M] = X[M, J nb + n]
    const int J = d->rank + x * G, rows = (nP - J) * nb;
    transpose_kernel<<<dim3(nb / 32, rows / 32), 256, 0, MS>>>(d->store + d->off[J], nb, XT + (size_t)x * nb * N + (size_t)J * nb, N, rows, nb);
    G3_LAUNCH_CHECK(ctx);
  }
  size_t npairs = 0;
  const long long ldk = (long long)std::max(nloc, 1) * nb;
  for (int I = 0; I < nP; ++I) {
    const int owner = I % G, rows = (nP - I) * nb;
    double* Xi = owner == d->rank ? d->store + d->off[I] : BT;                                     // panel I (rows x nb)
    if (G > 1) G3_NCCL(ctx, api, api->Broadcast(Xi, Xi, (size_t)rows * nb, ncclDouble, owner, d->comm, MS));
    int nuse = 0;
    for (int J = d->rank; J <= I; J += G) ++nuse;        // my panels J <= I are my first `nuse` ones
    if (nuse == 0) continue;
    transpose_kernel<<<dim3(nb / 32, rows / 32), 256, 0, MS>>>(Xi, nb, XTb + (size_t)I * nb, N, rows, nb);
    G3_LAUNCH_CHECK(ctx);
    // Kb[m][x nb + n] = sum_{M >= I nb} XTb[m][M] XT[x][n][M]
    {
      CUtensorMap tmA, tmB;
      if ((rc = g3_make_tmap(ctx, &tmA, XTb, (uint64_t)N, nb, 1, (uint64_t)N, (uint64_t)nb * N, G3_BM))) return rc;
      if ((rc = g3_make_tmap(ctx, &tmB, XT, (uint64_t)N, (uint64_t)nuse * nb, 1, (uint64_t)N, (uint64_t)nuse * nb * N, G3_BN))) return rc;
      GemmArgs g;
      memset(&g, 0, sizeof g);
      g.D = Kb; g.ldd = ldk; g.strideD = 0;
      g.mode = 0; g.ntx = nb / TS; g.nty = nuse * (nb / TS);
      g.a_r0 = 0; g.a_rx = TS; g.b_r0 = 0; g.b_ry = TS;
      g.ka0 = I * nb; g.kb0 = I * nb; g.kl0 = rows;
      g.alpha = 1.0; g.beta = 0.0;
      if ((rc = g3_gemm_launch(ctx, tmA, tmB, g, 1))) return rc;
    }
    for (int x = 0; x < nuse; ++x) {
      const int J = d->rank + x * G;
      VjpArgs v;
      memset(&v, 0, sizeof v);
      v.X1 = ctx->dX + (size_t)I * nb * D; v.X2 = ctx->dX + (size_t)J * nb * D; v.n1 = nb; v.n2 = nb; v.D = D;
      v.same = I == J; v.lower_only = I == J;
      v.theta = d->dtheta; v.P = P;
      v.W = Kb + (size_t)x * nb; v.ldw = ldk; v.strideW = 0;
      v.alpha = alpha + (size_t)I * nb; v.alpha2 = alpha + (size_t)J * nb; v.strideAlpha = 0; v.cfac = cf_dev;
      v.scale = I == J ? 0.5 : 1.0;                      // off-diagonal blocks stand for (I, J) and (J, I)
      v.dtheta = outp + npairs * std::max(P, 1);
      if ((rc = g3_gram_vjp_launch(ctx, d->desc, v, 1))) return rc;
      ++npairs;
    }
  }
  G3_CUDA(ctx, cudaEventRecord(d->ev_t[3], MS));
  std::vector<double> ho(npairs * std::max(P, 1)), ha(N);
  if (npairs) G3_CUDA(ctx, cudaMemcpyAsync(ho.data(), outp, sizeof(double) * ho.size(), cudaMemcpyDeviceToHost, MS));
  G3_CUDA(ctx, cudaMemcpyAsync(ha.data(), alpha, sizeof(double) * N, cudaMemcpyDeviceToHost, MS));
  G3_CUDA(ctx, cudaStreamSynchronize(MS));
  std::vector<double> g(std::max(P, 1), 0.0);
  for (size_t k = 0; k < npairs; ++k)
    for (int p2 = 0; p2 < P; ++p2) g[p2] += ho[k * P + p2];
  if (P > 0 && (rc = g3_comm_allreduce(ctx, g.data(), P, 0))) return rc;
  for (int p2 = 0; p2 < P; ++p2) dtheta_out[p2] = g[p2];
  if (ddelta_out_or_NULL)
    for (int i = 0; i < N; ++i) ddelta_out_or_NULL[i] = -cfac * ha[i];
  float t[3];
  for (int k = 0; k < 3; ++k) G3_CUDA(ctx, cudaEventElapsedTime(&t[k], d->ev_t[k], d->ev_t[k + 1]));
  double mx[3] = {t[0], t[1], t[2]};
  if ((rc = g3_comm_allreduce(ctx, mx, 3, 1))) return rc;
  if (ms3) { ms3[0] = (float)mx[0]; ms3[1] = (float)mx[1]; ms3[2] = (float)mx[2]; }
  return 0;
}

// Test accessor: copy piece (J, p) of THIS rank (count * nb x nb, factored or not) to the host; returns the number of
// blocks through *count (0 if the rank holds nothing of panel J).
int g3_dist_read_piece(g3_ctx* ctx, int J, double* host, int* count) {
  g3_dist* d = ctx ? ctx->dist : nullptr;
  if (!d || !d->store || J < 0 || J >= d->nP || !count) return g3_fail_msg(ctx, "g3_dist_read_piece: bad arguments");
  G3_CUDA(ctx, cudaSetDevice(ctx->device));
  G3_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *count = 0;
  if (J % d->Pc != d->q || d->off[J] < 0) return 0;
  *count = cnt_blk(J, d->p, d->Pr, d->nP);
  if (host) G3_CUDA(ctx, cudaMemcpy(host, d->store + d->off[J], sizeof(double) * (size_t)*count * blk_elems(d), cudaMemcpyDeviceToHost));
  return 0;
}

// SURVEY §8b name: factor + solve in one call.
int g3_potrf_2d(g3_ctx* ctx, const g3_kernel_desc* desc, const double* theta, int nb, int Pr, int Pc, int flags,
                const double* delta_or_NULL, double* logdet, double* beta_or_NULL, int* info, float* ms3) {
  float mg = 0.f, mp = 0.f, msv = 0.f;
  int rc = g3_dist_factor(ctx, desc, theta, nb, Pr, Pc, flags, logdet, info, &mg, &mp, nullptr);
  if (rc) return rc;
  if (delta_or_NULL && beta_or_NULL && (rc = g3_dist_solve(ctx, delta_or_NULL, beta_or_NULL, nullptr, &msv))) return rc;
  if (ms3) { ms3[0] = mg; ms3[1] = mp; ms3[2] = msv; }
  return 0;
}

}  // extern "C"

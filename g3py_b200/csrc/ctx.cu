// Context, workspaces, TMA descriptor creation, timers.
#include "g3b_internal.cuh"
#include <cstdio>
#include <cstring>

int g3_fail(g3_ctx* ctx, const char* what, cudaError_t e, const char* file, int line) {
  char buf[512];
  snprintf(buf, sizeof buf, "%s failed: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
  if (ctx) ctx->err = buf;
  return -2;
}

int g3_fail_msg(g3_ctx* ctx, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return -1;
}

void* g3_ws(g3_ctx* ctx, const char* name, size_t bytes) {
  g3_buf& b = ctx->bufs[name];
  if (b.bytes >= bytes && b.p) return b.p;
  if (b.p) {
    cudaStreamSynchronize(ctx->stream);
    cudaFree(b.p);
    b.p = nullptr;
    b.bytes = 0;
  }
  size_t want = (bytes + 255) & ~size_t(255);
  ctx->ws_gen++;                                  // addresses change: cached CUDA graphs are stale
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess) {
    char buf[256];
    snprintf(buf, sizeof buf, "cudaMalloc(%s, %zu bytes) failed: %s", name, want, cudaGetErrorString(e));
    ctx->err = buf;
    b.p = nullptr;
    cudaGetLastError();
    return nullptr;
  }
  b.bytes = want;
  return b.p;
}

void* g3_pinned(g3_ctx* ctx, const char* name, size_t bytes) {
  g3_buf& b = ctx->pinned[name];
  if (b.bytes >= bytes && b.p) return b.p;
  if (b.p) {
    cudaStreamSynchronize(ctx->stream);
    cudaFreeHost(b.p);
    b.p = nullptr;
    b.bytes = 0;
  }
  size_t want = (bytes + 4095) & ~size_t(4095);
  cudaError_t e = cudaHostAlloc(&b.p, want, cudaHostAllocDefault);
  if (e != cudaSuccess) {
    char buf[256];
    snprintf(buf, sizeof buf, "cudaHostAlloc(%s, %zu bytes) failed: %s", name, want, cudaGetErrorString(e));
    ctx->err = buf;
    b.p = nullptr;
    cudaGetLastError();
    return nullptr;
  }
  b.bytes = want;
  return b.p;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int g3_make_tmap(g3_ctx* ctx, CUtensorMap* out, const double* base, uint64_t cols, uint64_t rows, uint64_t batch,
                 uint64_t ld, uint64_t batch_stride, uint32_t box_rows) {
  if (!ctx->encode_fn) return g3_fail_msg(ctx, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t gdim[3] = {cols, rows, batch};
  cuuint64_t gstr[2] = {ld * sizeof(double), batch_stride * sizeof(double)};
  cuuint32_t box[3] = {G3_BK, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = ((PFN_encodeTiled)ctx->encode_fn)(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)base, gdim, gstr, box,
                                                 estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed: CUresult %d (cols %llu rows %llu batch %llu ld %llu)", (int)r,
             (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)batch, (unsigned long long)ld);
    return g3_fail_msg(ctx, buf);
  }
  return 0;
}

void g3_prof_begin(g3_ctx* ctx, int cls) {
  if (!ctx->prof_on) return;
  if (ctx->prof_used * 2 >= ctx->prof_events.size()) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    ctx->prof_events.push_back(a);
    ctx->prof_events.push_back(b);
    ctx->prof_class.push_back(cls);
  }
  ctx->prof_class[ctx->prof_used] = cls;
  cudaEventRecord(ctx->prof_events[ctx->prof_used * 2], ctx->stream);
}

void g3_prof_end(g3_ctx* ctx) {
  if (!ctx->prof_on) return;
  cudaEventRecord(ctx->prof_events[ctx->prof_used * 2 + 1], ctx->stream);
  ctx->prof_used++;
}

extern "C" {

// Debug / test accessor: copy the first `bytes` of a named workspace to the host.
int g3_debug_read(g3_ctx* ctx, const char* name, void* host, size_t bytes) {
  auto it = ctx->bufs.find(name);
  if (it == ctx->bufs.end() || !it->second.p || it->second.bytes < bytes) return g3_fail_msg(ctx, "g3_debug_read: no such buffer / too small");
  G3_CUDA(ctx, cudaSetDevice(ctx->device));
  G3_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  G3_CUDA(ctx, cudaMemcpy(host, it->second.p, bytes, cudaMemcpyDeviceToHost));
  return 0;
}

int g3_prof_enable(g3_ctx* ctx, int on) {
  ctx->prof_on = on != 0;
  ctx->prof_used = 0;
  return 0;
}

int g3_prof_read(g3_ctx* ctx, double* ms, int64_t* launches) {
  G3_CUDA(ctx, cudaSetDevice(ctx->device));
  G3_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int c = 0; c < G3_PROF_N; ++c) { ms[c] = 0.0; launches[c] = 0; }
  for (size_t i = 0; i < ctx->prof_used; ++i) {
    float t = 0.f;
    G3_CUDA(ctx, cudaEventElapsedTime(&t, ctx->prof_events[2 * i], ctx->prof_events[2 * i + 1]));
    ms[ctx->prof_class[i]] += t;
    launches[ctx->prof_class[i]]++;
  }
  ctx->prof_used = 0;
  return 0;
}

int g3_ctx_create(int device, g3_ctx** out) {
  if (!out) return -1;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return -3;  // no CUDA device: fail loudly, no CPU fallback
  if (device < 0 || device >= ndev) return -1;
  g3_ctx* c = new g3_ctx();
  c->device = device;
  if (cudaSetDevice(device) != cudaSuccess) { delete c; return -2; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete c; return -2; }
  if (prop.major != 10) { delete c; return -4; }  // sm_100a only
  c->sm_count = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return -2; }
  c->own_stream = c->stream;
  cudaEventCreate(&c->ev0);
  cudaEventCreate(&c->ev1);
  cudaEventCreateWithFlags(&c->ev_h2d, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&c->gev_start, cudaEventDisableTiming);
  for (int g = 0; g < G3_MAX_GROUPS; ++g) {
    cudaStreamCreateWithFlags(&c->gstream[g], cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&c->gev_done[g], cudaEventDisableTiming);
  }
  cudaDriverEntryPointQueryResult q;
  void* fn = nullptr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
      q == cudaDriverEntryPointSuccess)
    c->encode_fn = fn;
  *out = c;
  return 0;
}

int g3_ctx_destroy(g3_ctx* ctx) {
  if (!ctx) return 0;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  g3_dist_destroy(ctx);
  g3_graph_drop(ctx);
  for (auto& kv : ctx->bufs)
    if (kv.second.p) cudaFree(kv.second.p);
  for (auto& kv : ctx->pinned)
    if (kv.second.p) cudaFreeHost(kv.second.p);
  cudaEventDestroy(ctx->ev_h2d);
  for (cudaEvent_t e : ctx->prof_events) cudaEventDestroy(e);
  if (ctx->dX) cudaFree(ctx->dX);
  cudaEventDestroy(ctx->ev0);
  cudaEventDestroy(ctx->ev1);
  cudaEventDestroy(ctx->gev_start);
  for (int g = 0; g < G3_MAX_GROUPS; ++g) {
    cudaStreamDestroy(ctx->gstream[g]);
    cudaEventDestroy(ctx->gev_done[g]);
  }
  if (ctx->tri_stream) {
    cudaStreamDestroy(ctx->tri_stream);
    cudaEventDestroy(ctx->ev_tri);
  }
  if (ctx->panel_stream) {
    cudaStreamDestroy(ctx->panel_stream);
    cudaEventDestroy(ctx->ev_panel);
    cudaEventDestroy(ctx->ev_main);
  }
  cudaStreamDestroy(ctx->own_stream);
  delete ctx;
  return 0;
}

const char* g3_last_error(g3_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

// Free every cached workspace (device and pinned).  The context stays usable: buffers are re-created on demand.
int g3_ctx_trim(g3_ctx* ctx) {
  if (!ctx) return -1;
  G3_CUDA(ctx, cudaSetDevice(ctx->device));
  G3_CUDA(ctx, cudaDeviceSynchronize());
  for (auto& kv : ctx->bufs)
    if (kv.second.p) cudaFree(kv.second.p);
  ctx->bufs.clear();
  ctx->ws_gen++;
  g3_graph_drop(ctx);
  for (auto& kv : ctx->pinned)
    if (kv.second.p) cudaFreeHost(kv.second.p);
  ctx->pinned.clear();
  ctx->gp.valid = 0;
  ctx->gp.factor_resident = 0;
  return 0;
}

int g3_sync(g3_ctx* ctx) {
  G3_CUDA(ctx, cudaSetDevice(ctx->device));
  G3_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

int g3_set_stream(g3_ctx* ctx, void* stream_or_NULL) {
  // no synchronisation: the caller orders the streams it hands in (events / torch stream semantics)
  ctx->stream = stream_or_NULL ? (cudaStream_t)stream_or_NULL : ctx->own_stream;
  return 0;
}

int g3_set_jitter(g3_ctx* ctx, double jitter_rel, int max_tries) {
  ctx->jitter_rel = jitter_rel;
  ctx->max_tries = max_tries;
  return 0;
}

int g3_timer_begin(g3_ctx* ctx) {
  G3_CUDA(ctx, cudaSetDevice(ctx->device));
  G3_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  return 0;
}

int g3_timer_end(g3_ctx* ctx, float* ms) {
  G3_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  G3_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
  G3_CUDA(ctx, cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
  return 0;
}

int64_t g3_launch_count(g3_ctx* ctx) { return ctx->launches; }

int g3_set_data(g3_ctx* ctx, const double* X, int N, int D) {
  if (!X || N <= 0 || D <= 0 || D > G3_MAX_DIM) return g3_fail_msg(ctx, "g3_set_data: bad arguments (1 <= D <= 16)");
  G3_CUDA(ctx, cudaSetDevice(ctx->device));
  G3_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->ws_gen++;
  if (ctx->dX) { cudaFree(ctx->dX); ctx->dX = nullptr; }
  G3_CUDA(ctx, cudaMalloc(&ctx->dX, sizeof(double) * (size_t)N * D));
  G3_CUDA(ctx, cudaMemcpyAsync(ctx->dX, X, sizeof(double) * (size_t)N * D, cudaMemcpyHostToDevice, ctx->stream));
  G3_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->N = N;
  ctx->D = D;
  ctx->gp.valid = 0;
  return 0;
}

}  // extern "C"

// In-run roofline denominators (bench.py): sustained fp64 rate of the DMMA.8x8x4 tensor pipe and of the
// DFMA pipe, and a STREAM-style device copy.  MEASURED_PEAKS.json (driver-written) carries HBM and bf16
// figures only, so the fp64 peak every `roofline.frac` divides by is measured here, in the same process
// and on the same GPU as the numbers it judges.
#include "g3b_internal.cuh"

namespace {

template <int NACC>
__global__ void __launch_bounds__(512) peak_dmma_kernel(double* out, int iters, double a0, double b0) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = 0; c[i][1] = 0; }
  const double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(512) peak_dfma_kernel(double* out, int iters, double a0, double b0) {
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) c[i] = i;
  const double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) peak_copy_kernel(const double4* __restrict__ src, double4* __restrict__ dst, size_t n4) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) dst[i] = src[i];
}

}  // namespace

extern "C" int g3_debug_fp64_peak(g3_ctx* ctx, double seconds, double* dmma_tflops, double* dfma_tflops, double* copy_gbs) {
  if (!ctx) return -1;
  G3_CUDA(ctx, cudaSetDevice(ctx->device));
  const int sms = ctx->sm_count, threads = 512, iters = 20000;
  double* out = (double*)g3_ws(ctx, "peak_out", sizeof(double) * (size_t)sms * threads);
  if (!out) return -2;
  cudaStream_t s = ctx->stream;
  if (seconds <= 0.0) seconds = 0.5;
  auto timed = [&](auto launch, double flop_per_launch, double* tflops) -> int {
    launch();                                                     // warm-up
    G3_CUDA(ctx, cudaStreamSynchronize(s));
    G3_CUDA(ctx, cudaEventRecord(ctx->ev0, s));
    launch();
    G3_CUDA(ctx, cudaEventRecord(ctx->ev1, s));
    G3_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
    float ms1 = 0.f;
    G3_CUDA(ctx, cudaEventElapsedTime(&ms1, ctx->ev0, ctx->ev1));
    int reps = (int)(seconds * 1e3 / (ms1 > 1e-3f ? ms1 : 1e-3f));
    reps = reps < 1 ? 1 : (reps > 2000 ? 2000 : reps);
    G3_CUDA(ctx, cudaEventRecord(ctx->ev0, s));
    for (int r = 0; r < reps; ++r) launch();
    G3_CUDA(ctx, cudaEventRecord(ctx->ev1, s));
    G3_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
    float ms = 0.f;
    G3_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->launches += reps + 2;
    *tflops = flop_per_launch * reps / (ms * 1e-3) / 1e12;
    return 0;
  };
  int rc;
  if (dmma_tflops) {
    const double fl = 2.0 * 256 * 8 * (double)iters * (threads / 32) * sms;       // m8n8k4 = 256 FMA per warp instruction
    if ((rc = timed([&] { peak_dmma_kernel<8><<<sms, threads, 0, s>>>(out, iters, 1.0, 1e-3); }, fl, dmma_tflops))) return rc;
  }
  if (dfma_tflops) {
    const double fl = 2.0 * 32 * 16 * (double)iters * (threads / 32) * sms;
    if ((rc = timed([&] { peak_dfma_kernel<16><<<sms, threads, 0, s>>>(out, iters, 1.0, 1e-3); }, fl, dfma_tflops))) return rc;
  }
  if (copy_gbs) {
    const size_t bytes = (size_t)1 << 30;                                          // 1 GiB read + 1 GiB written, >> L2
    double* a = (double*)g3_ws(ctx, "peak_src", bytes);
    double* b = (double*)g3_ws(ctx, "peak_dst", bytes);
    if (!a || !b) return -2;
    G3_CUDA(ctx, cudaMemsetAsync(a, 0, bytes, s));
    double tb = 0.0;
    const size_t n4 = bytes / sizeof(double4);
    if ((rc = timed([&] { peak_copy_kernel<<<sms * 16, 256, 0, s>>>((const double4*)a, (double4*)b, n4); }, 2.0 * bytes, &tb)))
      return rc;
    *copy_gbs = tb * 1e3;                                                          // "TFLOP/s" of bytes -> GB/s
  }
  G3_CUDA(ctx, cudaGetLastError());
  return 0;
}

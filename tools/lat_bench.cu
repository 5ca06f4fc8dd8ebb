// Dependent-issue latencies on sm_100a that the diagonal-tile kernel's critical path is made of (one warp, one CTA).
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
#include <mma.h>
#define N 4096
__global__ void k_dfma(double* out, long long* cyc, double x, double y) {
  double a = x;
  long long t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < N; ++i) a = fma(a, y, x);
  long long t1 = clock64();
  out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_dfma4(double* out, long long* cyc, double x, double y) {   // 4 independent chains
  double a = x, b = x + 1, c = x + 2, d = x + 3;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { a = fma(a, y, x); b = fma(b, y, x); c = fma(c, y, x); d = fma(d, y, x); }
  long long t1 = clock64();
  out[threadIdx.x] = a + b + c + d; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_dfma16(double* out, long long* cyc, double x, double y) {   // 16 independent chains
  double a[16];
  for (int q = 0; q < 16; ++q) a[q] = x + q;
  long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int q = 0; q < 16; ++q) a[q] = fma(a[q], y, x);
  long long t1 = clock64();
  double s = 0; for (int q = 0; q < 16; ++q) s += a[q];
  out[threadIdx.x] = s; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_dmul(double* out, long long* cyc, double x, double y) {
  double a = x;
  long long t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < N; ++i) a = a * y;
  long long t1 = clock64();
  out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_ffma(float* out, long long* cyc, float x, float y) {
  float a = x;
  long long t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < N; ++i) a = fmaf(a, y, x);
  long long t1 = clock64();
  out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_rcp(double* out, long long* cyc, double x) {
  double a = x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { double y; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a)); a = y; }
  long long t1 = clock64();
  out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_div(double* out, long long* cyc, double x) {
  double a = x;
  long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) a = 1.0 / a;
  long long t1 = clock64();
  out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_rsqrt(double* out, long long* cyc, double x) {
  double a = x;
  long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) a = rsqrt(a);
  long long t1 = clock64();
  out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_shfl(double* out, long long* cyc, double x) {
  double a = x + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) a = __shfl_sync(0xffffffffu, a, (threadIdx.x + 1) & 31);
  long long t1 = clock64();
  out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lds(double* out, long long* cyc) {
  __shared__ int nxt[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) nxt[i] = (i * 17 + 5) & 1023;
  __syncthreads();
  int p = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) p = nxt[p];
  long long t1 = clock64();
  out[threadIdx.x] = p; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_sts_lds(double* out, long long* cyc, double x) {   // store -> syncwarp -> load of another lane's value
  __shared__ double buf[64];
  double a = x + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) {
    buf[(i & 1) * 32 + threadIdx.x] = a;
    __syncwarp();
    a = buf[(i & 1) * 32 + ((threadIdx.x + 1) & 31)];
  }
  long long t1 = clock64();
  out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_dmma(double* out, long long* cyc, double x) {   // dependent DMMA.8x8x4 chain (accumulator dependence)
  using namespace nvcuda::wmma;
  fragment<matrix_a, 8, 8, 4, double, row_major> fa;
  fragment<matrix_b, 8, 8, 4, double, col_major> fb;
  fragment<accumulator, 8, 8, 4, double> fc;
  fill_fragment(fa, x); fill_fragment(fb, x * 0.5); fill_fragment(fc, 0.0);
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) mma_sync(fc, fa, fb, fc);
  long long t1 = clock64();
  out[threadIdx.x] = fc.x[0] + fc.x[1]; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_dmma4(double* out, long long* cyc, double x) {   // 4 independent DMMA chains
  using namespace nvcuda::wmma;
  fragment<matrix_a, 8, 8, 4, double, row_major> fa;
  fragment<matrix_b, 8, 8, 4, double, col_major> fb;
  fragment<accumulator, 8, 8, 4, double> fc[4];
  fill_fragment(fa, x); fill_fragment(fb, x * 0.5);
  for (int q = 0; q < 4; ++q) fill_fragment(fc[q], 0.0);
  long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int q = 0; q < 4; ++q) mma_sync(fc[q], fa, fb, fc[q]);
  long long t1 = clock64();
  out[threadIdx.x] = fc[0].x[0] + fc[1].x[1] + fc[2].x[0] + fc[3].x[1]; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_seed_err(double* out) {   // max |1 - d * rcp.approx(d)| and max |1 - d * rsqrt.approx(d)^2| over 2^22 mantissas x a few exponents
  double m1 = 0.0, m2 = 0.0;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < (1u << 22); i += gridDim.x * blockDim.x) {
    for (int e = -3; e <= 3; e += 3) {
      const double d = ldexp(1.0 + (double)i / (double)(1u << 22) + 1.1e-9, e);
      double y, z;
      asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
      asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(z) : "d"(d));
      m1 = fmax(m1, fabs(fma(-d, y, 1.0)));
      m2 = fmax(m2, fabs(fma(-d * z, z, 1.0)));
    }
  }
  for (int o = 16; o > 0; o >>= 1) { m1 = fmax(m1, __shfl_xor_sync(0xffffffffu, m1, o)); m2 = fmax(m2, __shfl_xor_sync(0xffffffffu, m2, o)); }
  if ((threadIdx.x & 31) == 0) { atomicMax((unsigned long long*)out, __double_as_longlong(m1)); atomicMax((unsigned long long*)out + 1, __double_as_longlong(m2)); }
}
int main() {
  double* out; long long* cyc; float* fout;
  cudaMalloc(&out, 8 * 1024); cudaMalloc(&fout, 4 * 1024); cudaMalloc(&cyc, 8);
  long long h;
#define RUN(name, per, launch)                                                          \
  for (int w = 0; w < 2; ++w) { launch; cudaDeviceSynchronize(); }                       \
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);                                        \
  printf("%-40s %8.2f cycles per %s   (%s)\n", name, (double)h / N, per, cudaGetErrorString(cudaGetLastError()));
  for (int nw = 1; nw <= 4; nw *= 4) {
    printf("--- %d warp(s) in one CTA ---\n", nw);
    RUN("DFMA dependent chain", "op", (k_dfma<<<1, 32 * nw>>>(out, cyc, 1.0, 0.999)));
    RUN("DFMA 4 independent chains", "4 ops", (k_dfma4<<<1, 32 * nw>>>(out, cyc, 1.0, 0.999)));
    RUN("DFMA 16 independent chains", "16 ops", (k_dfma16<<<1, 32 * nw>>>(out, cyc, 1.0, 0.999)));
    RUN("DMUL dependent chain", "op", (k_dmul<<<1, 32 * nw>>>(out, cyc, 1.0, 0.999)));
    RUN("FFMA dependent chain", "op", (k_ffma<<<1, 32 * nw>>>(fout, cyc, 1.0f, 0.999f)));
    RUN("rcp.approx.f64 (MUFU.RCP64H) chain", "op", (k_rcp<<<1, 32 * nw>>>(out, cyc, 1.3)));
    RUN("1.0 / x (IEEE) chain", "op", (k_div<<<1, 32 * nw>>>(out, cyc, 1.3)));
    RUN("rsqrt(double) chain", "op", (k_rsqrt<<<1, 32 * nw>>>(out, cyc, 1.3)));
    RUN("SHFL.64 dependent chain", "op", (k_shfl<<<1, 32 * nw>>>(out, cyc, 1.3)));
    RUN("LDS pointer chase", "op", (k_lds<<<1, 32 * nw>>>(out, cyc)));
    RUN("STS -> syncwarp -> LDS round trip", "op", (k_sts_lds<<<1, 32>>>(out, cyc, 1.3)));
    RUN("DMMA.8x8x4 dependent chain", "op", (k_dmma<<<1, 32 * nw>>>(out, cyc, 1.0)));
    RUN("DMMA.8x8x4 4 independent chains", "4 ops", (k_dmma4<<<1, 32 * nw>>>(out, cyc, 1.0)));
  }
  // 16 warps: aggregate DFMA throughput of one SM with 16 chains per thread
  RUN("DFMA 16 chains, 16 warps", "16 ops", (k_dfma16<<<1, 512>>>(out, cyc, 1.0, 0.999)));
  RUN("DMMA 4 chains, 16 warps", "4 ops", (k_dmma4<<<1, 512>>>(out, cyc, 1.0)));
  cudaMemset(out, 0, 16);
  k_seed_err<<<148, 256>>>(out);
  double he[2];
  cudaMemcpy(he, out, 16, cudaMemcpyDeviceToHost);
  printf("seed error: rcp.approx.f64 max |1 - d y| = %.3e = 2^%.1f;  rsqrt.approx.f64 max |1 - d z^2| = %.3e = 2^%.1f\n", he[0], log2(he[0]), he[1], log2(he[1]));
  return 0;
}

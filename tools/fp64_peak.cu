// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
// Microbenchmark: sustained FP64 rate of the DMMA.8x8x4 tensor pipe and of the DFMA pipe on one GPU.
// Used only to record the fp64 roofline denominator (MEASURED_PEAKS.json has no fp64 entry).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

template<int NACC>
__global__ void __launch_bounds__(1024) dmma_loop(double* out, int iters, double a0, double b0) {
  double c[NACC][2];
  #pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = 0; c[i][1] = 0; }
  double a = a0 + threadIdx.x * 1e-9, b = b0;
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

template<int NACC>
__global__ void __launch_bounds__(1024) dfma_loop(double* out, int iters, double a0, double b0) {
  double c[NACC];
  #pragma unroll
  for (int i = 0; i < NACC; i++) c[i] = i;
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; it++) {
    #pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0;
  #pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  printf("device %s sms %d clock %d kHz\n", p.name, sms, p.clockRate);
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  int iters = 20000;
  for (int warps = 4; warps <= 32; warps *= 2) {
    for (int rep = 0; rep < 2; rep++) {
      dmma_loop<8><<<sms, warps * 32>>>(out, iters, 1.0, 1e-3);
      CK(cudaDeviceSynchronize());
    }
    CK(cudaEventRecord(e0));
    dmma_loop<8><<<sms, warps * 32>>>(out, iters, 1.0, 1e-3);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double flops = 2.0 * 256 * 8 * (double)iters * warps * sms;
    printf("DMMA.8x8x4 warps/SM=%d: %.3f ms  %.2f TFLOP/s\n", warps, ms, flops / ms * 1e-9);
  }
  // longer sustained run (about 2 s) to see the power-capped rate
  {
    CK(cudaEventRecord(e0));
    for (int r = 0; r < 40; r++) dmma_loop<8><<<sms, 512>>>(out, iters * 4, 1.0, 1e-3);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double flops = 40.0 * 2.0 * 256 * 8 * (double)iters * 4 * 16 * sms;
    printf("DMMA.8x8x4 sustained (%.0f ms): %.2f TFLOP/s\n", ms, flops / ms * 1e-9);
  }
  for (int warps = 8; warps <= 32; warps *= 2) {
    dfma_loop<16><<<sms, warps * 32>>>(out, iters, 1.0, 1e-3);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    dfma_loop<16><<<sms, warps * 32>>>(out, iters, 1.0, 1e-3);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double flops = 2.0 * 32 * 16 * (double)iters * warps * sms;
    printf("DFMA warps/SM=%d: %.3f ms  %.2f TFLOP/s\n", warps, ms, flops / ms * 1e-9);
  }
  return 0;
}

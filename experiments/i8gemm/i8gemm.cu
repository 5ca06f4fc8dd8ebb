// EXPERIMENT, not part of libg3b.so: the building block of the Ozaki-type fp64-equivalent GEMM that DESIGN.md section 8
// names as the next step.  C (int32, M x N) = A (int8, M x K, K contiguous) * B^T (int8, N x K, K contiguous) on the
// 5th-generation tensor cores:
//   * operands: TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B, 128 int8 of K per row) into an mbarrier ring of shared memory,
//   * math: tcgen05.mma.cta_group::1.kind::i8, M=128 x N=256 x K=32 per instruction, issued by ONE thread, int32
//     accumulators in tensor memory (256 columns),
//   * completion: tcgen05.commit onto the ring's "empty" barriers and onto the accumulator-ready barrier,
//   * epilogue: 4 warps read the accumulator with tcgen05.ld (32 lanes x 32 columns each) and store rows to global.
// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = tensor-memory allocator, 4..7 = epilogue.  One tile per CTA.
//
// Stand-alone test program (main below): exact check against a CPU int32 GEMM, then a timed run.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o i8gemm i8gemm.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

constexpr int BM = 128, BN = 256, BK = 128;          // BK int8 = 128 bytes = one SWIZZLE_128B row
constexpr int UMMA_K = 32;                           // K per tcgen05.mma for 8-bit operands
constexpr int STAGE_A = BM * BK, STAGE_B = BN * BK;  // 16 KiB + 32 KiB
constexpr int TMEM_COLS = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded spin: a protocol error traps (the launch fails) instead of hanging the device
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t spin = 0; !ok; ++spin) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spin > (1u << 22)) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
// shared-memory matrix descriptor, K-major, SWIZZLE_128B (cute/arch/mma_sm100_desc.hpp SmemDescriptor): start address
// >> 4 in bits [0,14), leading byte offset (unused for swizzled K-major: 1) in [16,30), stride byte offset = 8 rows x
// 128 B = 1024 B >> 4 in [32,46), version 1 in [46,48), layout type 2 (SWIZZLE_128B) in [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor (InstrDescriptor): c_format S32 = 2 at [4,6), a/b format signed 8-bit = 1 at [7,10) / [10,13),
// K-major A and B (bits 15, 16 = 0), N >> 3 at [17,23), M >> 4 at [24,29)
constexpr uint32_t kIdesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(kIdesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
      "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}

template <int STAGES>
__global__ void __launch_bounds__(256, 1)
i8gemm_nt_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int32_t* __restrict__ C,
                 int ldc, int K) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sA = smem;                                  // STAGES x 16 KiB
  unsigned char* sB = smem + STAGES * STAGE_A;               // STAGES x 32 KiB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (STAGE_A + STAGE_B));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES), accum_bar = smem_u32(bars + 2 * STAGES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int nkb = K / BK;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {                                       // ===== TMA producer
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        mbar_wait(empty0 + 8 * s, ((kb / STAGES) & 1) ^ 1);
        mbar_expect_tx(full0 + 8 * s, STAGE_A + STAGE_B);
        tma_load_2d(smem_u32(sA + s * STAGE_A), &tmA, full0 + 8 * s, kb * BK, m0);
        tma_load_2d(smem_u32(sB + s * STAGE_B), &tmB, full0 + 8 * s, kb * BK, n0);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {                                       // ===== MMA issuer (one thread)
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        mbar_wait(full0 + 8 * s, (kb / STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t da = make_desc(smem_u32(sA + s * STAGE_A)), db = make_desc(smem_u32(sB + s * STAGE_B));
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k)               // +32 bytes of K inside the swizzle atom: +2 in address>>4
          umma_i8(tmem_base, da + 2 * k, db + 2 * k, (kb | k) != 0);
        umma_commit(empty0 + 8 * s);                         // frees the stage when these MMAs have read it
      }
      umma_commit(accum_bar);                                // accumulator complete
    }
  } else if (warp >= 4) {                                    // ===== epilogue: warp w reads tensor-memory lanes 32w..
    const int w = warp & 3;
    mbar_wait(accum_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    int32_t* crow = C + (size_t)(m0 + 32 * w + lane) * ldc + n0;
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(32 * w) << 16) + c, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<int4*>(crow + c + j) = make_int4((int)v[j], (int)v[j + 1], (int)v[j + 2], (int)v[j + 3]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
}

// ---------------------------------------------------------------------------------------------------------- host
#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess) {                                                                   \
      fprintf(stderr, "%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_));     \
      exit(2);                                                                                 \
    }                                                                                          \
  } while (0)

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(PFN_encodeTiled enc, const int8_t* base, uint64_t rows, uint64_t K, uint32_t box_rows) {
  CUtensorMap m;
  cuuint64_t gdim[2] = {K, rows};
  cuuint64_t gstr[1] = {K};
  cuuint32_t box[2] = {BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "cuTensorMapEncodeTiled failed: %d\n", (int)r);
    exit(2);
  }
  return m;
}

template <int STAGES>
static void launch(const CUtensorMap& a, const CUtensorMap& b, int32_t* C, int M, int N, int K, cudaStream_t st) {
  const int smem = STAGES * (STAGE_A + STAGE_B) + 8 * (2 * STAGES + 1) + 16;
  static bool once = false;
  if (!once) {
    CK(cudaFuncSetAttribute(i8gemm_nt_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    once = true;
  }
  i8gemm_nt_kernel<STAGES><<<dim3(N / BN, M / BM), 256, smem, st>>>(a, b, C, N, K);
}

static uint32_t rng_state = 12345u;
static inline int8_t rnd7() {                                // 7-bit slices: values in [-64, 64]
  rng_state = rng_state * 1664525u + 1013904223u;
  return (int8_t)((int)((rng_state >> 16) % 129u) - 64);
}

int main(int argc, char** argv) {
  int Mbig = argc > 1 ? atoi(argv[1]) : 8192, Kbig = argc > 2 ? atoi(argv[2]) : 4096, reps = argc > 3 ? atoi(argv[3]) : 20;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaFree(0));
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (q != cudaDriverEntryPointSuccess) {
    fprintf(stderr, "no cuTensorMapEncodeTiled\n");
    return 2;
  }
  PFN_encodeTiled enc = (PFN_encodeTiled)fn;
  cudaStream_t st;
  CK(cudaStreamCreate(&st));

  // ---- 1. exact check against the CPU: every entry of a 256 x 512 x K=640 product (5 k-blocks: ring wraps with 4 stages)
  {
    const int M = 256, N = 512, K = 640;
    std::vector<int8_t> hA((size_t)M * K), hB((size_t)N * K);
    for (auto& x : hA) x = rnd7();
    for (auto& x : hB) x = rnd7();
    int8_t *dA, *dB;
    int32_t* dC;
    CK(cudaMalloc(&dA, hA.size()));
    CK(cudaMalloc(&dB, hB.size()));
    CK(cudaMalloc(&dC, (size_t)M * N * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dC, 0xFF, (size_t)M * N * 4));
    CUtensorMap ta = make_map(enc, dA, M, K, BM), tb = make_map(enc, dB, N, K, BN);
    launch<4>(ta, tb, dC, M, N, K, st);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    std::vector<int32_t> hC((size_t)M * N);
    CK(cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost));
    long bad = 0;
    for (int i = 0; i < M; ++i)
      for (int j = 0; j < N; ++j) {
        int32_t s = 0;
        for (int k = 0; k < K; ++k) s += (int32_t)hA[(size_t)i * K + k] * (int32_t)hB[(size_t)j * K + k];
        if (s != hC[(size_t)i * N + j]) {
          if (bad < 8) printf("  mismatch C[%d,%d]: got %d want %d\n", i, j, hC[(size_t)i * N + j], s);
          ++bad;
        }
      }
    printf("{\"check\": \"exact\", \"M\": %d, \"N\": %d, \"K\": %d, \"mismatches\": %ld}\n", M, N, K, bad);
    cudaFree(dA);
    cudaFree(dB);
    cudaFree(dC);
    if (bad) return 1;
  }
  // ---- 2. timed: Mbig x Mbig x Kbig, sampled exact check
  {
    const int M = Mbig, N = Mbig, K = Kbig;
    std::vector<int8_t> hA((size_t)M * K), hB((size_t)N * K);
    for (auto& x : hA) x = rnd7();
    for (auto& x : hB) x = rnd7();
    int8_t *dA, *dB;
    int32_t* dC;
    CK(cudaMalloc(&dA, hA.size()));
    CK(cudaMalloc(&dB, hB.size()));
    CK(cudaMalloc(&dC, (size_t)M * N * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
    CUtensorMap ta = make_map(enc, dA, M, K, BM), tb = make_map(enc, dB, N, K, BN);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) launch<4>(ta, tb, dC, M, N, K, st);
    CK(cudaStreamSynchronize(st));
    CK(cudaEventRecord(e0, st));
    for (int i = 0; i < reps; ++i) launch<4>(ta, tb, dC, M, N, K, st);
    CK(cudaEventRecord(e1, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
    std::vector<int32_t> hC((size_t)M * N);
    CK(cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost));
    long bad = 0;
    for (int t = 0; t < 4096; ++t) {
      rng_state = rng_state * 1664525u + 1013904223u;
      int i = (rng_state >> 8) % M;
      rng_state = rng_state * 1664525u + 1013904223u;
      int j = (rng_state >> 8) % N;
      int32_t s = 0;
      for (int k = 0; k < K; ++k) s += (int32_t)hA[(size_t)i * K + k] * (int32_t)hB[(size_t)j * K + k];
      bad += s != hC[(size_t)i * N + j];
    }
    printf("{\"bench\": \"i8gemm_nt\", \"M\": %d, \"N\": %d, \"K\": %d, \"ms\": %.4f, \"TOPs\": %.1f, \"sampled_mismatches\": %ld, "
           "\"tile\": \"128x256x128\", \"stages\": 4}\n",
           M, N, K, ms, 2.0 * M * N * (double)K / (ms * 1e-3) / 1e12, bad);
    if (bad) return 1;
  }
  return 0;
}

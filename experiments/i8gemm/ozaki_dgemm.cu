// EXPERIMENT, not part of libg3b.so: fp64-equivalent  C -= A * B^T  (A: M x K, B: N x K, both K-contiguous fp64; the "NT"
// shape of every level-3 step of the blocked Cholesky) on the INT8 tensor cores, by error-free slicing (Ozaki scheme):
//
//   1. slice_rows: every row of A and B is scaled by a power of two to |x| < 1 and cut into S slices of 7 bits + sign,
//      x = 2^e * sum_t q_t 2^(-7(t+1)), q_t integer, |q_t| <= 127: exact (every step is a power-of-two scaling, a
//      truncation and an exact subtraction).  Stored as int8 tensors [S][rows][K].
//   2. ozaki_kernel: for d = S-1 .. 0 (least significant first) the products A_t B_u^T of all slice pairs with
//      t + u = d are accumulated EXACTLY in one int32 accumulator in tensor memory (|sum| <= (d+1) K 127^2 < 2^31 for
//      K <= 14 000): tcgen05.mma kind::i8 (128 x 256 x 32 per instruction) fed by a TMA / mbarrier ring, exactly the
//      kernel of i8gemm.cu with the K loop running over (pair, k-block).  The epilogue converts to fp64, applies the
//      scale 2^(ea_i + eb_j - 7(d+2)) (exact) and subtracts from C.  Pairs with t + u >= S are dropped (the truncation
//      error, ~2^(-7S) relative to the row scales).
//      Two accumulators (2 x 256 of the 512 tensor-memory columns) alternate, so the epilogue of pass d overlaps the
//      MMAs of pass d-1; the C tile (256 KiB per CTA) is re-read from L2 between passes.
//
// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = tensor-memory allocator, 4..11 = epilogue.  One C tile per CTA.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o ozaki_dgemm ozaki_dgemm.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

constexpr int BM = 128, BN = 256, BK = 128;
constexpr int UMMA_K = 32;
constexpr int STAGE_A = BM * BK, STAGE_B = BN * BK;
constexpr int STAGES = 4;
constexpr int TMEM_COLS = 512;
#ifndef OZAKI_BULK_EPILOGUE
constexpr int SCRATCH_LD = 33;                               // 32 x 33 words per epilogue warp: conflict-free transpose
constexpr int EPI_WARPS = 8;                                 // two per tensor-memory lane quarter (128 columns each)
constexpr int kScratchBytes = EPI_WARPS * 32 * SCRATCH_LD * 4;
#else
// NEXT STEP, compiled only with -DOZAKI_BULK_EPILOGUE (make next) and NOT YET RUN ON HARDWARE: the epilogue stages
// the scaled, negated values of a 32-row x 32-column chunk in shared memory (one 256-byte row segment per lane) and
// lets the bulk-copy engine add them to C in L2 (cp.reduce.async.bulk ... add.f64): no global loads in the epilogue.
constexpr int EPI_WARPS = 4;
constexpr int STAGE_LD = 272;                                // bytes per staged row: 256 + 16 (16-byte aligned, 4-way conflicts)
constexpr int kScratchBytes = EPI_WARPS * 32 * STAGE_LD;
#endif
constexpr int THREADS = 128 + 32 * EPI_WARPS;
constexpr int kSmemBytes = STAGES * (STAGE_A + STAGE_B) + 8 * (2 * STAGES + 4) + 16 + kScratchBytes;
static_assert(kSmemBytes <= 232448, "shared memory per CTA");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {   // bounded: a protocol error traps
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
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
// K-major SWIZZLE_128B shared-memory descriptor and the s8 x s8 -> s32 instruction descriptor: see i8gemm.cu
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
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

// ------------------------------------------------------------------------------------------------ 1. slicing
// One CTA per row.  out: [S][rows][K] int8, scale[row] = 2^e with |x| / 2^e < 1 on the row.
template <int S>
__global__ void __launch_bounds__(256) slice_rows_kernel(const double* __restrict__ X, int ld, int K, int rows,
                                                         int8_t* __restrict__ out, double* __restrict__ scale) {
  __shared__ double red[8];
  const int row = blockIdx.x;
  const double* x = X + (size_t)row * ld;
  double amax = 0.0;
  for (int k = threadIdx.x; k < K; k += 256) amax = fmax(amax, fabs(x[k]));
  for (int o = 16; o; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = amax;
  __syncthreads();
  amax = red[0];
  for (int w = 1; w < 8; ++w) amax = fmax(amax, red[w]);
  int e = 0;
  if (amax > 0.0) frexp(amax, &e);                           // amax = f 2^e, f in [0.5, 1)
  const double inv = ldexp(1.0, -e);
  if (threadIdx.x == 0) scale[row] = ldexp(1.0, e);
  for (int k4 = threadIdx.x * 4; k4 < K; k4 += 1024) {
    double r[4];
    for (int j = 0; j < 4; ++j) r[j] = x[k4 + j] * inv;      // exact
#pragma unroll
    for (int t = 0; t < S; ++t) {
      char4 q;
      signed char* qq = reinterpret_cast<signed char*>(&q);
      for (int j = 0; j < 4; ++j) {
        r[j] *= 128.0;
        const double f = trunc(r[j]);
        r[j] -= f;
        qq[j] = (signed char)(int)f;
      }
      *reinterpret_cast<char4*>(out + ((size_t)t * rows + row) * K + k4) = q;
    }
  }
}

// ------------------------------------------------------------------------------------------------ 2. slice products
__global__ void __launch_bounds__(THREADS, 1)
ozaki_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, double* __restrict__ C, int ldc,
             int K, int S, const double* __restrict__ sa, const double* __restrict__ sb, int lower) {
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  if (lower && n0 >= m0 + BM) return;                        // tile entirely above the diagonal (whole CTA leaves)
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sA = smem;
  unsigned char* sB = smem + STAGES * STAGE_A;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (STAGE_A + STAGE_B));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  uint32_t* scratch = reinterpret_cast<uint32_t*>(smem + STAGES * (STAGE_A + STAGE_B) + 8 * (2 * STAGES + 4) + 16);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES);
  const uint32_t tfull0 = smem_u32(bars + 2 * STAGES), tempty0 = smem_u32(bars + 2 * STAGES + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = K / BK;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull0 + 8 * b, 1);
      mbar_init(tempty0 + 8 * b, 32 * EPI_WARPS);             // every epilogue thread arrives
    }
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
      uint32_t it = 0;
      for (int d = S - 1; d >= 0; --d)
        for (int t = 0; t <= d; ++t)
          for (int kb = 0; kb < nkb; ++kb, ++it) {
            const uint32_t s = it % STAGES;
            mbar_wait(empty0 + 8 * s, ((it / STAGES) & 1) ^ 1);
            mbar_expect_tx(full0 + 8 * s, STAGE_A + STAGE_B);
            tma_load_3d(smem_u32(sA + s * STAGE_A), &tmA, full0 + 8 * s, kb * BK, m0, t);
            tma_load_3d(smem_u32(sB + s * STAGE_B), &tmB, full0 + 8 * s, kb * BK, n0, d - t);
          }
    }
  } else if (warp == 1) {
    if (elect_one()) {                                       // ===== MMA issuer
      uint32_t it = 0;
      for (int p = 0; p < S; ++p) {
        const int d = S - 1 - p, buf = p & 1;
        mbar_wait(tempty0 + 8 * buf, ((p >> 1) & 1) ^ 1);    // epilogue has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int n_it = (d + 1) * nkb;
        for (int i = 0; i < n_it; ++i, ++it) {
          const uint32_t s = it % STAGES;
          mbar_wait(full0 + 8 * s, (it / STAGES) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t da = make_desc(smem_u32(sA + s * STAGE_A)), db = make_desc(smem_u32(sB + s * STAGE_B));
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) umma_i8(tmem_base + buf * BN, da + 2 * k, db + 2 * k, (i | k) != 0);
          umma_commit(empty0 + 8 * s);
        }
        umma_commit(tfull0 + 8 * buf);
      }
    }
#ifndef OZAKI_BULK_EPILOGUE
  } else if (warp >= 4) {                                    // ===== epilogue: C -= double(acc) * 2^(ea + eb - 7(d+2))
    const int w = warp & 3, half = (warp - 4) >> 2;          // lanes 32w.. of tensor memory, columns 128*half..
    uint32_t* sc = scratch + (warp - 4) * 32 * SCRATCH_LD;
    const double rs_mine = sa[m0 + 32 * w + lane];           // lane r holds the scale of row 32w + r
    double* ctile = C + (size_t)(m0 + 32 * w) * ldc + n0 + 128 * half + lane;
    for (int p = 0; p < S; ++p) {
      const int d = S - 1 - p, buf = p & 1;
      const double common = __longlong_as_double((long long)(1023 - 7 * (d + 2)) << 52);   // 2^(-7(d+2))
      double cv[32];                                         // this lane's column of the 32 x 32 chunk: 32 loads in flight
#pragma unroll
      for (int r = 0; r < 32; ++r) cv[r] = ctile[(size_t)r * ldc];          // first chunk: before the accumulator is ready
      mbar_wait(tfull0 + 8 * buf, (p >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int q = 0; q < 4; ++q) {
        const int c = 128 * half + 32 * q;
        double* cp = ctile + 32 * q;
        if (q > 0) {
#pragma unroll
          for (int r = 0; r < 32; ++r) cv[r] = cp[(size_t)r * ldc];
        }
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(32 * w) << 16) + buf * BN + c, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) sc[lane * SCRATCH_LD + j] = v[j];
        __syncwarp();
        const double cs = sb[n0 + c + lane] * common;
#pragma unroll
        for (int r = 0; r < 32; ++r) {
          const double rs = __shfl_sync(0xffffffffu, rs_mine, r);
          cv[r] -= (double)(int)sc[r * SCRATCH_LD + lane] * (rs * cs);
        }
#pragma unroll
        for (int r = 0; r < 32; ++r) cp[(size_t)r * ldc] = cv[r];
        __syncwarp();
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(tempty0 + 8 * buf);
    }
  }
#else
  } else if (warp >= 4) {                                    // ===== epilogue: C += -(double)acc * 2^(ea + eb - 7(d+2)) in L2
    const int w = warp & 3;
    unsigned char* stage = reinterpret_cast<unsigned char*>(scratch) + w * 32 * STAGE_LD;
    double* myrow = reinterpret_cast<double*>(stage + lane * STAGE_LD);
    const uint32_t myrow_s = smem_u32(myrow);
    const double rs = -sa[m0 + 32 * w + lane];               // this lane's row scale, negated: the reduction adds
    double* crow = C + (size_t)(m0 + 32 * w + lane) * ldc + n0;
    for (int p = 0; p < S; ++p) {
      const int d = S - 1 - p, buf = p & 1;
      const double common = __longlong_as_double((long long)(1023 - 7 * (d + 2)) << 52);   // 2^(-7(d+2))
      // levels must reach C in order (fp64 addition is not associative): the previous level's reductions are complete
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      mbar_wait(tfull0 + 8 * buf, (p >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const double rsc = rs * common;
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(32 * w) << 16) + buf * BN + c, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the engine has read this lane's staged row
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const double2 cs = *reinterpret_cast<const double2*>(sb + n0 + c + j);          // same address in every lane
          *reinterpret_cast<double2*>(myrow + j) = make_double2((double)(int)v[j] * (rsc * cs.x), (double)(int)v[j + 1] * (rsc * cs.y));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> visible to the engine
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;"
                     ::"l"(crow + c), "r"(myrow_s), "r"(256)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(tempty0 + 8 * buf);
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
#endif
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
}

// ---------------------------------------------------------------------------------------------------------- host
#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      fprintf(stderr, "%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
      exit(2);                                                                             \
    }                                                                                      \
  } while (0)

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_enc;

static CUtensorMap make_map3(const int8_t* base, uint64_t rows, uint64_t K, uint64_t S, uint32_t box_rows) {
  CUtensorMap m;
  cuuint64_t gdim[3] = {K, rows, S};
  cuuint64_t gstr[2] = {K, rows * K};
  cuuint32_t box[3] = {BK, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "cuTensorMapEncodeTiled failed: %d\n", (int)r);
    exit(2);
  }
  return m;
}

constexpr int S_SLICES = 9;

struct Sliced {
  int8_t* q = nullptr;
  double* scale = nullptr;
  int rows = 0, K = 0;
  CUtensorMap map;
};
static Sliced make_sliced(int rows, int K, uint32_t box_rows) {
  Sliced s;
  s.rows = rows;
  s.K = K;
  CK(cudaMalloc(&s.q, (size_t)S_SLICES * rows * K));
  CK(cudaMalloc(&s.scale, (size_t)rows * 8));
  s.map = make_map3(s.q, rows, K, S_SLICES, box_rows);
  return s;
}
static void slice(const double* dX, Sliced& s, cudaStream_t st) {
  slice_rows_kernel<S_SLICES><<<s.rows, 256, 0, st>>>(dX, s.K, s.K, s.rows, s.q, s.scale);
}
static void ozaki(const Sliced& a, const Sliced& b, double* dC, int M, int N, int K, int lower, cudaStream_t st) {
  ozaki_kernel<<<dim3(N / BN, M / BM), THREADS, kSmemBytes, st>>>(a.map, b.map, dC, N, K, S_SLICES, a.scale, b.scale, lower);
}

static uint64_t rs64 = 88172645463325252ull;
static inline double urand() {                               // xorshift, uniform in (-1, 1)
  rs64 ^= rs64 << 13;
  rs64 ^= rs64 >> 7;
  rs64 ^= rs64 << 17;
  return ((double)(rs64 >> 11) / 9007199254740992.0) * 2.0 - 1.0;
}
// rows spanning three orders of magnitude, like the panels of a Cholesky factor
static void fill(std::vector<double>& X, int rows, int K) {
  for (int i = 0; i < rows; ++i) {
    const double rsc = pow(10.0, -3.0 * fabs(urand()));
    for (int k = 0; k < K; ++k) X[(size_t)i * K + k] = rsc * urand();
  }
}

// errors of sampled entries against a long double dot product, in units of |c0_ij| + sum_k |a_ik| |b_jk| (the backward
// error scale of C -= A B^T: the final subtraction alone rounds at half an ulp of that)
static void grade(const std::vector<double>& A, const std::vector<double>& B, const std::vector<double>& C0,
                  const std::vector<double>& C, int M, int N, int K, int lower, int samples, double* err_oz, double* err_f64) {
  *err_oz = *err_f64 = 0.0;
  for (int t = 0; t < samples; ++t) {
    int i, j;
    if (samples >= M * N) {
      i = t / N;
      j = t % N;
    } else {
      i = (int)((urand() * 0.5 + 0.5) * M) % M;
      j = (int)((urand() * 0.5 + 0.5) * N) % N;
    }
    if (lower && j > i) continue;
    long double s = 0.0L, mag = 0.0L;
    double f = 0.0;
    for (int k = 0; k < K; ++k) {
      const double a = A[(size_t)i * K + k], b = B[(size_t)j * K + k];
      s += (long double)a * (long double)b;
      mag += fabsl((long double)a * (long double)b);
      f += a * b;
    }
    const long double ref = (long double)C0[(size_t)i * N + j] - s;
    mag += fabsl((long double)C0[(size_t)i * N + j]);
    const double e1 = (double)(fabsl((long double)C[(size_t)i * N + j] - ref) / mag);
    const double e2 = (double)(fabsl((long double)(C0[(size_t)i * N + j] - f) - ref) / mag);
    if (e1 > *err_oz) *err_oz = e1;
    if (e2 > *err_f64) *err_f64 = e2;
  }
}

int main(int argc, char** argv) {
  const int Mbig = argc > 1 ? atoi(argv[1]) : 8192, reps = argc > 2 ? atoi(argv[2]) : 5;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaFree(0));
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (q != cudaDriverEntryPointSuccess) return 2;
  g_enc = (PFN_encodeTiled)fn;
  CK(cudaFuncSetAttribute(ozaki_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));

  // ---- 1. small case, every entry: slicing exactness and the full pipeline
  {
    const int M = 256, N = 512, K = 640;
    std::vector<double> A((size_t)M * K), B((size_t)N * K), C0((size_t)M * N), C((size_t)M * N);
    fill(A, M, K);
    fill(B, N, K);
    for (size_t i = 0; i < C0.size(); ++i) C0[i] = urand();
    // cancellation: make a quarter of the entries the fp64 product plus something tiny (a Schur complement)
    for (int i = 0; i < M; ++i)
      for (int j = 0; j < N; j += 4) {
        double f = 0.0;
        for (int k = 0; k < K; ++k) f += A[(size_t)i * K + k] * B[(size_t)j * K + k];
        C0[(size_t)i * N + j] = f + 1e-9 * urand();
      }
    double *dA, *dB, *dC;
    CK(cudaMalloc(&dA, A.size() * 8));
    CK(cudaMalloc(&dB, B.size() * 8));
    CK(cudaMalloc(&dC, C.size() * 8));
    CK(cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dC, C0.data(), C.size() * 8, cudaMemcpyHostToDevice));
    Sliced sa = make_sliced(M, K, BM), sb = make_sliced(N, K, BN);
    slice(dA, sa, st);
    slice(dB, sb, st);
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    {  // slicing is exact: x == 2^e sum_t q_t 2^(-7(t+1)) + remainder below 2^(e - 7S)
      std::vector<int8_t> hq((size_t)S_SLICES * M * K);
      std::vector<double> hs(M);
      CK(cudaMemcpy(hq.data(), sa.q, hq.size(), cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(hs.data(), sa.scale, M * 8, cudaMemcpyDeviceToHost));
      double worst = 0.0;
      for (int i = 0; i < M; ++i)
        for (int k = 0; k < K; ++k) {
          long double rec = 0.0L;
          for (int t = 0; t < S_SLICES; ++t)
            rec += (long double)hq[((size_t)t * M + i) * K + k] * ldexpl(1.0L, -7 * (t + 1));
          const double rel = (double)(fabsl(rec * (long double)hs[i] - (long double)A[(size_t)i * K + k]) / (long double)hs[i]);
          if (rel > worst) worst = rel;
        }
      printf("{\"check\": \"slicing\", \"S\": %d, \"max_residual_over_row_scale\": %.3e, \"bound_2^-7S\": %.3e}\n", S_SLICES, worst,
             ldexp(1.0, -7 * S_SLICES));
      if (!(worst <= ldexp(1.0, -7 * S_SLICES))) return 1;
    }
    ozaki(sa, sb, dC, M, N, K, 0, st);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    CK(cudaMemcpy(C.data(), dC, C.size() * 8, cudaMemcpyDeviceToHost));
    double eo, ef;
    grade(A, B, C0, C, M, N, K, 0, M * N, &eo, &ef);
    printf("{\"check\": \"ozaki_vs_long_double\", \"M\": %d, \"N\": %d, \"K\": %d, \"S\": %d, \"max_err_ozaki\": %.3e, "
           "\"max_err_fp64_dot\": %.3e, \"unit\": \"|c0|+sum_k|a||b|\"}\n", M, N, K, S_SLICES, eo, ef);
    cudaFree(dA);
    cudaFree(dB);
    cudaFree(dC);
    if (!(eo < 2e-15)) return 1;
  }
  // ---- 2. timed: Mbig x Mbig, K = 1024 (one panel of the blocked Cholesky) and 4096; full and lower-triangle-only
  for (int K : {1024, 4096}) {
    const int M = Mbig, N = Mbig;
    std::vector<double> A((size_t)M * K), B((size_t)N * K), C0((size_t)M * N), C((size_t)M * N);
    fill(A, M, K);
    fill(B, N, K);
    for (size_t i = 0; i < C0.size(); ++i) C0[i] = urand();
    double *dA, *dB, *dC;
    CK(cudaMalloc(&dA, A.size() * 8));
    CK(cudaMalloc(&dB, B.size() * 8));
    CK(cudaMalloc(&dC, C.size() * 8));
    CK(cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice));
    Sliced sa = make_sliced(M, K, BM), sb = make_sliced(N, K, BN);
    cudaEvent_t e0, e1, e2;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventCreate(&e2));
    for (int lower = 0; lower < 2; ++lower) {
      // accuracy on one application
      CK(cudaMemcpy(dC, C0.data(), C.size() * 8, cudaMemcpyHostToDevice));
      slice(dA, sa, st);
      slice(dB, sb, st);
      ozaki(sa, sb, dC, M, N, K, lower, st);
      CK(cudaStreamSynchronize(st));
      CK(cudaGetLastError());
      CK(cudaMemcpy(C.data(), dC, C.size() * 8, cudaMemcpyDeviceToHost));
      double eo, ef;
      grade(A, B, C0, C, M, N, K, lower, 4096, &eo, &ef);
      // timing (C keeps accumulating: the values do not matter)
      float ms_slice = 0.f, ms_gemm = 0.f;
      for (int r = 0; r < reps + 1; ++r) {
        CK(cudaEventRecord(e0, st));
        slice(dA, sa, st);
        slice(dB, sb, st);
        CK(cudaEventRecord(e1, st));
        ozaki(sa, sb, dC, M, N, K, lower, st);
        CK(cudaEventRecord(e2, st));
        CK(cudaStreamSynchronize(st));
        float a, b;
        CK(cudaEventElapsedTime(&a, e0, e1));
        CK(cudaEventElapsedTime(&b, e1, e2));
        if (r) {                                             // first repetition is warm-up
          ms_slice += a / reps;
          ms_gemm += b / reps;
        }
      }
      CK(cudaGetLastError());
      const double flops = 2.0 * M * (double)N * K * (lower ? 0.5 * (1.0 + (double)BM / M) : 1.0);   // fp64 flops replaced
      printf("{\"bench\": \"ozaki_dgemm_nt\", \"M\": %d, \"N\": %d, \"K\": %d, \"S\": %d, \"lower_only\": %d, \"ms_slice\": %.4f, "
             "\"ms_products\": %.4f, \"fp64_equiv_TFLOPs_products\": %.1f, \"fp64_equiv_TFLOPs_with_slicing\": %.1f, "
             "\"int8_TOPs\": %.0f, \"max_err_ozaki\": %.3e, \"max_err_fp64_dot\": %.3e}\n",
             M, N, K, S_SLICES, lower, ms_slice, ms_gemm, flops / (ms_gemm * 1e-3) / 1e12,
             flops / ((ms_gemm + ms_slice) * 1e-3) / 1e12, flops * 45.0 / (ms_gemm * 1e-3) / 1e12, eo, ef);
      fflush(stdout);
    }
    cudaFree(dA);
    cudaFree(dB);
    cudaFree(dC);
    cudaFree(sa.q);
    cudaFree(sb.q);
    cudaFree(sa.scale);
    cudaFree(sb.scale);
  }
  return 0;
}

// fp64-equivalent panel updates on the INT8 tensor cores (tcgen05.mma kind::i8, accumulators in tensor memory) by
// error-free slicing (Ozaki scheme) - the only way past the DMMA ceiling the fp64 GEMM of gemm.cu sits at.
// Selected with g3_set_gemm_mode(ctx, G3_GEMM_OZAKI); used by the batched left-looking Cholesky (potrf.cu) for the
// update of a 256-wide block column with ALL earlier columns, where the contraction is deep:
//
//     C[r][c] -= sum_{k < k1} L[r][k] L[c][k],     r >= row0 (128-row tiles), c in [col0, col0 + 256)
//
// The reference computes this inside LAPACK dpotrf (g3py/libs/tensors.py:198); here
//   1. every row r of L has ONE power-of-two scale for the whole factorisation, 2^e_r >= sqrt(K_rr) >= |L[r][k]| (the row of
//      L has norm sqrt(K_rr)), known before the factorisation starts (oz_row_scale_kernel on the diagonal of K);
//   2. when a block column of L is final it is cut into S = 9 slices of 7 bits + sign, L[r][k] = 2^e_r sum_t q_t 2^(-7(t+1))
//      (exact: power-of-two scalings, truncations, exact subtractions) - int8 planes [S][B Np][Np] (oz_slice_kernel);
//   3. oz_update_kernel: for significance d = S-1 .. 0 all slice pairs (t, u), t + u = d, are accumulated EXACTLY in one
//      int32 accumulator over the whole contraction (|sum| <= 9 * 127^2 * K < 2^31 for K <= 14 000), TMA -> 4-stage
//      mbarrier ring -> tcgen05.mma 128 x 256 x 32 issued by one thread, two 256-column accumulators alternating so the
//      fp64 epilogue of level d (convert, scale by 2^(e_r + e_c - 7(d+2)) - exact -, subtract from C) overlaps the MMAs
//      of level d-1.  Dropped pairs (t + u >= S) are below 2^-63 of the row scales.
// Measured on the stand-alone prototype (experiments/i8gemm, round 1): error 2.5-3.1e-16 of |c0| + sum|a||b|, the same as
// an fp64 dot product; 44.6 / 54.6 TFLOP/s fp64-equivalent at K = 1024 / 4096 against 37.1 for the DMMA pipe.
//
// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = tensor-memory allocator, 4..11 = epilogue.  One C tile per CTA,
// 226 KB of shared memory, so one CTA per SM (and never two tensor-memory allocations on one SM).
#include "g3b_internal.cuh"
#include <math.h>

namespace {

constexpr int TS = G3_TILE;
constexpr int BM = 128, BN = 256, BK = 128;
constexpr int UMMA_K = 32;
constexpr int STAGE_A = BM * BK, STAGE_B = BN * BK;
constexpr int STAGES = 4;
constexpr int TMEM_COLS = 512;
constexpr int SCRATCH_LD = 33;                               // 32 x 33 words per epilogue warp: conflict-free transpose
constexpr int EPI_WARPS = 8;                                 // two per tensor-memory lane quarter (128 columns each)
constexpr int kScratchBytes = EPI_WARPS * 32 * SCRATCH_LD * 4;
constexpr int THREADS = 128 + 32 * EPI_WARPS;
constexpr int kSmemBytes = STAGES * (STAGE_A + STAGE_B) + 8 * (2 * STAGES + 4) + 16 + kScratchBytes;
static_assert(kSmemBytes <= 232448, "shared memory per CTA");
constexpr int S_SLICES = 9;

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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {   // bounded: a protocol error traps, never hangs
  uint32_t ok = 0;
  for (uint32_t spin = 0; !ok; ++spin) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spin > (1u << 24)) __trap();
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
// K-major SWIZZLE_128B shared-memory descriptor (rows of 128 bytes written by TMA; SBO = 8 rows) and the
// s8 x s8 -> s32 instruction descriptor, M = 128, N = 256 (validated bit-exact in experiments/i8gemm/i8gemm.cu)
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

// ---- 1. row scales: 2^e_r > sqrt(K_rr) with one bit of head room; the diagonal of K as built by the Gram kernel
__global__ void oz_row_scale_kernel(const double* __restrict__ A, int Np, long long strideA, double* __restrict__ scale) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (r >= Np) return;
  const double d = A[(long long)b * strideA + (long long)r * Np + r];
  int e = 0;
  if (d > 0.0 && d < 1e300) frexp(sqrt(d), &e);              // sqrt(d) = f 2^e, f in [0.5, 1)
  scale[(long long)b * Np + r] = ldexp(1.0, e + 1);          // |L[r][k]| <= sqrt(K_rr) < 2^e: quotient < 1/2, never 127 + 1
}

// ---- 2. slicing of the finished tile columns [c0, c0 + ncols) of L, rows >= r0, for every matrix of the batch.
// One thread per 4 consecutive columns.  planes: [S][B Np][Np] int8.
__global__ void __launch_bounds__(256)
oz_slice_kernel(const double* __restrict__ A, int Np, long long strideA, int r0, int c0, int ncols, const double* __restrict__ scale,
                int8_t* __restrict__ planes, long long plane_stride) {
  const int per_row = ncols / 4;
  const int idx = blockIdx.x * 256 + threadIdx.x;
  const int r = r0 + idx / per_row, c = c0 + (idx % per_row) * 4, b = blockIdx.y;
  if (r >= Np) return;
  const double* x = A + (long long)b * strideA + (long long)r * Np + c;
  const double inv = 1.0 / scale[(long long)b * Np + r];     // power of two: exact
  double v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    double q = (c + j <= r) ? x[j] * inv : 0.0;              // strictly lower + diagonal of L; nothing above
    q = fmin(fmax(q, -0.99), 0.99);                          // a failed (NaN / huge) factor stays in range: flagged by info anyway
    v[j] = (q == q) ? q : 0.0;
  }
  int8_t* out = planes + ((long long)b * Np + r) * Np + c;
#pragma unroll
  for (int t = 0; t < S_SLICES; ++t) {
    char4 q4;
    signed char* qq = reinterpret_cast<signed char*>(&q4);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j] *= 128.0;
      const double f = trunc(v[j]);
      v[j] -= f;
      qq[j] = (signed char)(int)f;
    }
    *reinterpret_cast<char4*>(out + (long long)t * plane_stride) = q4;
  }
}

// ---- 3. the update.  grid (1, row tiles, B).  C tile rows [row0 + 128 y, +128), columns [col0, col0 + 256).
__global__ void __launch_bounds__(THREADS, 1)
oz_update_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, double* __restrict__ C, int ldc,
                 long long strideC, int Np, int row0, int col0, int K, int S, const double* __restrict__ scale, int ncols_valid) {
  const int b = blockIdx.z;
  const int m0 = row0 + blockIdx.y * BM, n0 = col0;
  const int arow = b * Np + m0, brow = b * Np + n0;          // rows of the batch-folded slice planes
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
    for (int q = 0; q < 2; ++q) {
      mbar_init(tfull0 + 8 * q, 1);
      mbar_init(tempty0 + 8 * q, 32 * EPI_WARPS);            // every epilogue thread arrives
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
            tma_load_3d(smem_u32(sA + s * STAGE_A), &tmA, full0 + 8 * s, kb * BK, arow, t);
            tma_load_3d(smem_u32(sB + s * STAGE_B), &tmB, full0 + 8 * s, kb * BK, brow, d - t);
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
  } else if (warp >= 4) {                                    // ===== epilogue: C -= double(acc) * 2^(e_r + e_c - 7(d+2))
    const int w = warp & 3, half = (warp - 4) >> 2;          // lanes 32w.. of tensor memory, columns 128*half..
    const bool live = 128 * half < ncols_valid;              // a 128-wide last block column: the second half does not exist
    uint32_t* sc = scratch + (warp - 4) * 32 * SCRATCH_LD;
    const double rs_mine = scale[(long long)b * Np + m0 + 32 * w + lane];          // lane r holds the scale of row 32w + r
    double* ctile = C + (long long)b * strideC + (long long)(m0 + 32 * w) * ldc + n0 + 128 * half + lane;
    for (int p = 0; p < S; ++p) {
      const int d = S - 1 - p, buf = p & 1;
      const double common = __longlong_as_double((long long)(1023 - 7 * (d + 2)) << 52);   // 2^(-7(d+2))
      double cv[32];                                         // this lane's column of the 32 x 32 chunk: 32 loads in flight
      if (live) {
#pragma unroll
        for (int r = 0; r < 32; ++r) cv[r] = ctile[(long long)r * ldc];     // first chunk: before the accumulator is ready
      }
      mbar_wait(tfull0 + 8 * buf, (p >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (live) {
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
          const int c = 128 * half + 32 * q;
          double* cp = ctile + 32 * q;
          if (q > 0) {
#pragma unroll
            for (int r = 0; r < 32; ++r) cv[r] = cp[(long long)r * ldc];
          }
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(32 * w) << 16) + buf * BN + c, v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 32; ++j) sc[lane * SCRATCH_LD + j] = v[j];
          __syncwarp();
          const double cs = scale[(long long)b * Np + n0 + c + lane] * common;
#pragma unroll
          for (int r = 0; r < 32; ++r) {
            const double rs = __shfl_sync(0xffffffffu, rs_mine, r);
            cv[r] -= (double)(int)sc[r * SCRATCH_LD + lane] * (rs * cs);
          }
#pragma unroll
          for (int r = 0; r < 32; ++r) cp[(long long)r * ldc] = cv[r];
          __syncwarp();
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(tempty0 + 8 * buf);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_plane_map(g3_ctx* ctx, CUtensorMap* out, const int8_t* base, uint64_t cols, uint64_t rows, uint32_t box_rows) {
  if (!ctx->encode_fn) return g3_fail_msg(ctx, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t gdim[3] = {cols, rows, (cuuint64_t)S_SLICES};
  cuuint64_t gstr[2] = {cols, rows * cols};
  cuuint32_t box[3] = {BK, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = ((PFN_encodeTiled)ctx->encode_fn)(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)base, gdim, gstr, box, estr,
                                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return g3_fail_msg(ctx, "cuTensorMapEncodeTiled (int8 planes) failed");
  return 0;
}

}  // namespace

// Workspace of the sliced factor for a batch of B matrices of order Np: planes + row scales.
int g3_oz_prepare(g3_ctx* ctx, const double* A, int Np, int B, g3_oz_state* st) {
  st->Np = Np;
  st->B = B;
  st->plane_stride = (long long)B * Np * Np;
  // per-stream buffers: the batch groups of g3_gp_run factor concurrently
  char name[64], name2[64];
  snprintf(name, sizeof name, "oz_planes_%p", (void*)ctx->stream);
  snprintf(name2, sizeof name2, "oz_scale_%p", (void*)ctx->stream);
  st->planes = (int8_t*)g3_ws(ctx, name, (size_t)S_SLICES * st->plane_stride);
  st->scale = (double*)g3_ws(ctx, name2, sizeof(double) * (size_t)B * Np);
  if (!st->planes || !st->scale) return -2;
  static bool attr = false;
  if (!attr) {
    G3_CUDA(ctx, cudaFuncSetAttribute(oz_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr = true;
  }
  int rc;
  if ((rc = make_plane_map(ctx, &st->tmA, st->planes, (uint64_t)Np, (uint64_t)B * Np, BM))) return rc;
  if ((rc = make_plane_map(ctx, &st->tmB, st->planes, (uint64_t)Np, (uint64_t)B * Np, BN))) return rc;
  oz_row_scale_kernel<<<dim3((Np + 255) / 256, B), 256, 0, ctx->stream>>>(A, Np, (long long)Np * Np, st->scale);
  G3_LAUNCH_CHECK(ctx);
  return 0;
}

// slices of the finished tile columns [j0, j1) of L (rows >= j0 * 128)
int g3_oz_slice(g3_ctx* ctx, const g3_oz_state* st, const double* A, int j0, int j1) {
  const int Np = st->Np, r0 = j0 * TS, ncols = (j1 - j0) * TS;
  const long long work = (long long)(Np - r0) * (ncols / 4);
  g3_prof_begin(ctx, G3_PROF_OTHER);
  oz_slice_kernel<<<dim3((unsigned)((work + 255) / 256), st->B), 256, 0, ctx->stream>>>(A, Np, (long long)Np * Np, r0, j0 * TS, ncols, st->scale,
                                                                                     st->planes, st->plane_stride);
  g3_prof_end(ctx);
  G3_LAUNCH_CHECK(ctx);
  return 0;
}

// A[r][c] -= sum_{k < j0*128} L[r][k] L[c][k] for r >= j0*128, c in the block columns [j0, j1) (j1 - j0 <= 2)
int g3_oz_update(g3_ctx* ctx, const g3_oz_state* st, double* A, int j0, int j1) {
  const int Np = st->Np, T = Np / TS;
  const int K = j0 * TS;
  if (K % BK || K <= 0 || j1 - j0 < 1 || j1 - j0 > 2) return g3_fail_msg(ctx, "g3_oz_update: bad geometry");
  dim3 grid(1, (unsigned)(T - j0), (unsigned)st->B);
  g3_prof_begin(ctx, G3_PROF_GEMM);
  oz_update_kernel<<<grid, THREADS, kSmemBytes, ctx->stream>>>(st->tmA, st->tmB, A, Np, (long long)Np * Np, Np, j0 * TS, j0 * TS, K, S_SLICES,
                                                               st->scale, (j1 - j0) * TS);
  g3_prof_end(ctx);
  G3_LAUNCH_CHECK(ctx);
  return 0;
}

// Batched fp64 "NT" GEMM for sm_100a:  D[m][n] = beta*D[m][n] + alpha * sum_k A[m][k]*B[n][k]
//
// * operands: TMA (cp.async.bulk.tensor.3d, SWIZZLE_128B) into a 4-stage shared-memory ring,
//   completion on mbarriers; one elected thread is the producer.
// * math: DMMA.8x8x4 (mma.sync.m8n8k4.f64), 8 warps as 2(m) x 4(n), 32x32 per warp.
//   A lane reads 16-byte chunks (two consecutive k) and the four k4-steps of a k-tile use the
//   permuted contraction sets {e, 4+e, 8+e, 12+e}; A and B use the same permutation so the
//   sum is unchanged while every fragment load is a conflict-free LDS.128.
// * CTA tile 64 x 128, 96 KiB of shared memory -> two CTAs per SM so that one CTA's epilogue
//   (read-modify-write of the 64x128 D tile) overlaps the other's main loop.
// This one kernel implements every level-3 step of potrf / trtri / lauum / trsm (see potrf.cu);
// the reference does these through LAPACK dpotrf and Theano's Murray reverse mode
// (g3py/libs/tensors.py:198,224-260).
#include "g3b_internal.cuh"
#include <stdio.h>
#include <string>

namespace {

constexpr int kStageBytesA = G3_BM * G3_BK * 8;  // 8 KiB
constexpr int kStageBytesB = G3_BN * G3_BK * 8;  // 16 KiB
constexpr int kStageBytes = kStageBytesA + kStageBytesB;
constexpr int kSmemBytes = G3_STAGES * kStageBytes + 64;  // stages (1024-B aligned base) + 8 mbarriers; NO static smem,
// so that one GEMM CTA (98,368 B + 1 KiB system reserve) leaves room for a 128x128 fp64 diagonal-block CTA on the same SM

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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void lds128(uint32_t addr, double& x, double& y) {
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x), "=d"(y) : "r"(addr));
}

// NJ = 8-column fragments per warp: 4 = the CTA computes the whole 64 x 128 tile; 2 / 1 = a half / a quarter of its columns
// (the other CTAs of the tile load the same operand boxes).  Launches of a few tiles are bound by ONE SM's fp64 rate --
// 8.3 us for a 64 x 128 x 128 tile -- so the column updates and triangular solves of single-matrix evaluations spread
// each tile over 2 or 4 SMs.
template <int NJ>
__global__ void __launch_bounds__(256, 2)
dgemm_nt_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs g) {
  constexpr int CS = 4 / NJ;                 // CTAs per 64-row half tile
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + G3_STAGES * kStageBytes);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int grp = lane >> 2, t4 = lane & 3;
  // warp -> (row group wm, column group wn); column groups are paired (0,3) and (1,2) on the same SM sub-partition
  // (warp % 4) so that the k-tile skipping of a triangular B operand stays balanced across the four DMMA pipes
  const int wm = warp & 1, wn = warp < 4 ? (warp >> 1) : 3 - ((warp - 4) >> 1);

  // ---- tile decode -------------------------------------------------------------------
  // blockIdx.x = (tile * 2 + h) * CS + cs: the CS CTAs that share one operand box are consecutive = one thread-block cluster
  const int cs = (int)blockIdx.x % CS;
  const int h = ((int)blockIdx.x / CS) & 1;  // which 64-row half of the 128-row block
  const int tile = (int)blockIdx.x / (2 * CS);
  const int col0 = cs * (32 * NJ) + wn * (8 * NJ);          // this warp's first column inside the 128-wide tile
  int x, y;
  if (g.mode == 0) {
    x = tile % g.ntx;
    y = tile / g.ntx;
  } else {
    x = (int)((sqrt(8.0 * (double)tile + 1.0) - 1.0) * 0.5);
    while ((long long)x * (x + 1) / 2 > tile) --x;
    while ((long long)(x + 1) * (x + 2) / 2 <= tile) ++x;
    y = tile - (int)((long long)x * (x + 1) / 2);
  }
  if ((g.upper & 4) && x < y) return;                       // void tile (strictly above the block diagonal)
  const bool diag_tile = ((g.upper & 1) && x == y) || ((g.upper & 2) && x == 0);
  // warps whose 32x32 block lies strictly above the diagonal of a diagonal tile have nothing to compute; they
  // only pace the ring (rows h*64 + wm*32 .. +31, columns wn*32 .. +31)
  const bool idle = diag_tile && col0 >= h * 64 + wm * 32 + 32;
  const int kt_lim = g.tri_b ? (col0 + 8 * NJ + 15) / 16 : 0x7fffffff;     // triangular B: k-tiles this warp's columns reach
  const int bidx = g.bmap ? g.bmap[blockIdx.y] : (int)blockIdx.y;
  const int a_row = g.a_r0 + x * g.a_rx + y * g.a_ry + h * G3_BM;
  const int b_row = g.b_r0 + x * g.b_rx + y * g.b_ry;
  int ka = g.ka0 + x * g.ka_x + y * g.ka_y;
  int kb = g.kb0 + x * g.kb_x + y * g.kb_y;
  int nk = (g.kl0 + x * g.kl_x + y * g.kl_y) / G3_BK;
  const int sk = g.splitk > 1 ? (int)blockIdx.z : 0;
  int kt0 = 0;         // first k-tile of this CTA's share (the triangular-B skipping counts absolute k-tiles)
  if (g.splitk > 1) {  // this CTA's share of the contraction: k-tiles [sk*chunk, (sk+1)*chunk)
    const int chunk = (nk + g.splitk - 1) / g.splitk;
    const int lo = sk * chunk;
    kt0 = lo;
    nk = nk - lo < chunk ? nk - lo : chunk;
    if (nk < 0) nk = 0;
    ka += lo * G3_BK;
    kb += lo * G3_BK;
  }
  double* Dt = g.D + (long long)bidx * g.strideD + (long long)(g.d_r0 + x * G3_TILE + h * G3_BM) * g.ldd +
               (g.d_c0 + y * G3_TILE);

  const uint32_t smem_base = smem_u32(smem_raw);
  if (smem_base & 1023u) __trap();  // SWIZZLE_128B needs 1024-B aligned stages
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (G3_STAGES + s); };

  if (tid == 0) {
    for (int s = 0; s < G3_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  auto issue = [&](int kt) {  // thread 0 only
    const int s = kt % G3_STAGES;
    const uint32_t dstA = smem_base + s * kStageBytes;
    const uint32_t dstB = dstA + kStageBytesA;
    mbar_expect_tx(full_bar(s), kStageBytes);
    tma_load_3d(dstA, &tmA, full_bar(s), ka + kt * G3_BK, a_row, bidx);
    tma_load_3d(dstB, &tmB, full_bar(s), kb + kt * G3_BK, b_row, bidx);
  };

  if (tid == 0) {
    const int pre = nk < G3_STAGES ? nk : G3_STAGES;
    for (int kt = 0; kt < pre; ++kt) issue(kt);
  }

  // beta != 0: the D tile enters through the accumulators (acc = (beta/alpha) * D, loaded while the TMA pipeline
  // fills), so the epilogue is a pure store and no global-load latency sits between the last DMMA and the write.
  double acc[4][NJ][2];
  if (g.beta != 0.0 && !idle && sk == 0) {
    const double sc = g.beta / g.alpha;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double* rowp = Dt + (long long)(wm * 32 + i * 8 + grp) * g.ldd + col0 + 2 * t4;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const double2 v = *reinterpret_cast<const double2*>(rowp + j * 8);
        acc[i][j][0] = sc * v.x;
        acc[i][j][1] = sc * v.y;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  }

  // per-thread fragment offsets inside a stage: row*128 + ((chunk ^ (row&7)) << 4), row&7 == grp
  const uint32_t offA = (uint32_t)((wm * 32 + grp) * 128);
  const uint32_t offB = (uint32_t)(kStageBytesA + (col0 + grp) * 128);
  const uint32_t sw0 = (uint32_t)(((2 * t4 + 0) ^ grp) << 4);
  const uint32_t sw1 = (uint32_t)(((2 * t4 + 1) ^ grp) << 4);

  for (int kt = 0; kt < nk; ++kt) {
    const int s = kt % G3_STAGES;
    const uint32_t ph = (uint32_t)((kt / G3_STAGES) & 1);
    // producer: refill the stage released one iteration ago
    if (tid == 0 && kt >= 1 && kt - 1 + G3_STAGES < nk) {
      const int sp = (kt - 1) % G3_STAGES;
      mbar_wait(empty_bar(sp), (uint32_t)(((kt - 1) / G3_STAGES) & 1));
      issue(kt - 1 + G3_STAGES);
    }
    mbar_wait(full_bar(s), ph);
    const uint32_t st = smem_base + s * kStageBytes;
    if (!idle && kt0 + kt < kt_lim) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const uint32_t sw = c ? sw1 : sw0;
      double a[4][2], b[NJ][2];
#pragma unroll
      for (int i = 0; i < 4; ++i) lds128(st + offA + i * 1024 + sw, a[i][0], a[i][1]);
#pragma unroll
      for (int j = 0; j < NJ; ++j) lds128(st + offB + j * 1024 + sw, b[j][0], b[j][1]);
#pragma unroll
      for (int e = 0; e < 2; ++e)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < NJ; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i][e], b[j][e]);
    }
    }
    // Release the stage.  The LDS above are asynchronous: ptxas hoists the arrive right behind the last LDS
    // *issue*, and an mbarrier arrive does not wait for this warp's outstanding shared-memory reads, so the
    // TMA refill (async proxy) could overwrite the stage under the last fragment loads (seen as rare, per-warp
    // garbage in the b[3] fragment).  The proxy fence drains this thread's loads and orders them before the
    // async-proxy write that the arrive enables.
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) mbar_arrive(empty_bar(s));
  }

  // ---- epilogue ------------------------------------------------------------------------
  if (CS > 1) {
    // In-place launches (L_ij = A_ij Linv^T, U_ji = -S_ji Linv^T) read the tile they overwrite: every CTA of the cluster
    // must have consumed its copy of the operand box before any of them stores its columns.
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (g.splitk > 1) {
    // split-K: park the partial tile, count arrivals; the last CTA of the tile adds the partials in split order
    __shared__ int is_last;
    const long long tile_lin = (long long)blockIdx.y * gridDim.x + blockIdx.x;
    double* part = g.sk_ws + (tile_lin * g.splitk + sk) * (long long)(G3_BM * G3_BN);
    if (!idle) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j)
          *reinterpret_cast<double2*>(part + (wm * 32 + i * 8 + grp) * G3_BN + col0 + j * 8 + 2 * t4) =
              make_double2(acc[i][j][0], acc[i][j][1]);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const unsigned prev = atomicAdd(g.sk_cnt + tile_lin, 1u);
      is_last = prev == (unsigned)(g.splitk - 1);
      if (is_last) g.sk_cnt[tile_lin] = 0;                    // ready for the next launch on this stream
    }
    __syncthreads();
    if (!is_last || idle) return;
    __threadfence();
    const double* base = g.sk_ws + tile_lin * g.splitk * (long long)(G3_BM * G3_BN);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double* rowp = Dt + (long long)(wm * 32 + i * 8 + grp) * g.ldd + col0 + 2 * t4;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const long long off = (wm * 32 + i * 8 + grp) * G3_BN + col0 + j * 8 + 2 * t4;
        double2 sum = make_double2(0.0, 0.0);
        for (int q = 0; q < g.splitk; ++q) {
          const double2 v = __ldcg(reinterpret_cast<const double2*>(base + (long long)q * (G3_BM * G3_BN) + off));
          sum.x += v.x;
          sum.y += v.y;
        }
        *reinterpret_cast<double2*>(rowp + j * 8) = make_double2(g.alpha * sum.x, g.alpha * sum.y);
      }
    }
    return;
  }
  if (idle) return;
  const double alpha = g.alpha;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double* rowp = Dt + (long long)(wm * 32 + i * 8 + grp) * g.ldd + col0 + 2 * t4;
#pragma unroll
    for (int j = 0; j < NJ; ++j)
      *reinterpret_cast<double2*>(rowp + j * 8) = make_double2(alpha * acc[i][j][0], alpha * acc[i][j][1]);
  }
}

}  // namespace

int g3_gemm_launch(g3_ctx* ctx, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& a, int B) {
  if (!ctx->gemm_ready) {
    G3_CUDA(ctx, cudaFuncSetAttribute(dgemm_nt_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    G3_CUDA(ctx, cudaFuncSetAttribute(dgemm_nt_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    G3_CUDA(ctx, cudaFuncSetAttribute(dgemm_nt_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    G3_CUDA(ctx, cudaFuncSetAttribute(dgemm_nt_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    G3_CUDA(ctx, cudaFuncSetAttribute(dgemm_nt_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    G3_CUDA(ctx, cudaFuncSetAttribute(dgemm_nt_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    ctx->gemm_ready = true;
  }
  long long ntiles = a.mode == 0 ? (long long)a.ntx * a.nty : (long long)a.ntx * (a.ntx + 1) / 2;
  if (ntiles <= 0 || B <= 0) return 0;
  GemmArgs g = a;
  g.splitk = 1;
  g.sk_ws = nullptr;
  g.sk_cnt = nullptr;
  // Few tiles with a deep contraction (column updates of a single matrix, rows of trtri) leave most SMs idle and each CTA
  // runs at one SM's fp64 rate: split the contraction over up to 8 CTAs per tile, at least one 128-block per share.
  // (g3_set_splitk(ctx, 2) goes down to 32 per share and also splits the triangular solves, whose 64 x 128 tile costs
  // 8.3 us on one SM even at depth 128 -- measured SLOWER on B200, profiles/r02t_splitk_ab.txt: the partial-tile round trip
  // through L2 and the extra CTAs cost more than the shorter main loop saves.)
  // column split of the tiles while one CTA per SM is not reached (g3_set_tile_split)
  int csplit = 1;
  if (ctx->tile_split) {
    const long long base = ntiles * 2 * B;
    csplit = base * 4 <= ctx->sm_count ? 4 : (base * 2 <= ctx->sm_count ? 2 : 1);
  }
  const long long ctas = ntiles * 2 * B * csplit;
  if (ctx->splitk && ctas * 2 <= 2 * ctx->sm_count) {
    int kmax = a.kl0;
    if (a.mode == 0) {
      const int kx = a.kl0 + a.kl_x * (a.ntx - 1), ky = a.kl0 + a.kl_y * (a.nty - 1);
      kmax = kmax > kx ? kmax : kx;
      kmax = kmax > ky ? kmax : ky;
    }
    int sk = (int)(2 * ctx->sm_count / ctas);
    const int min_share = ctx->splitk >= 2 ? 32 : 128;
    if (a.tri_b && ctx->splitk < 2) sk = 1;
    if (sk > kmax / min_share) sk = kmax / min_share;
    if (sk > 8) sk = 8;
    if (sk >= 2) {
      // per-stream scratch: launches on one stream are ordered, different streams (batch groups, look-ahead) are not
      char name[64];
      snprintf(name, sizeof name, "gemm_sk_%p", (void*)ctx->stream);
      const size_t ws_bytes = sizeof(double) * (size_t)ctas * sk * G3_BM * G3_BN;
      std::string key(name);
      const void* before = ctx->bufs.count(key + "_c") ? ctx->bufs[key + "_c"].p : nullptr;
      double* ws = (double*)g3_ws(ctx, name, ws_bytes);
      unsigned* cnt = (unsigned*)g3_ws(ctx, (key + "_c").c_str(), sizeof(unsigned) * 4096);
      if (ws && cnt && ctas <= 4096) {
        if (cnt != before) G3_CUDA(ctx, cudaMemsetAsync(cnt, 0, sizeof(unsigned) * 4096, ctx->stream));
        g.splitk = sk;
        g.sk_ws = ws;
        g.sk_cnt = cnt;
      }
    }
  }
  dim3 grid((unsigned)(ntiles * 2 * csplit), (unsigned)B, (unsigned)g.splitk);
  g3_prof_begin(ctx, G3_PROF_GEMM);
  if (csplit > 1) {   // the CTAs of one tile half form a cluster (barrier before the stores, see the kernel)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)csplit;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaError_t e = csplit == 4 ? cudaLaunchKernelEx(&cfg, dgemm_nt_kernel<1>, tmA, tmB, g)
                                : cudaLaunchKernelEx(&cfg, dgemm_nt_kernel<2>, tmA, tmB, g);
    if (e != cudaSuccess) return g3_fail(ctx, "cudaLaunchKernelEx(dgemm_nt, cluster)", e, __FILE__, __LINE__);
  } else {
    dgemm_nt_kernel<4><<<grid, 256, kSmemBytes, ctx->stream>>>(tmA, tmB, g);
  }
  g3_prof_end(ctx);
  G3_LAUNCH_CHECK(ctx);
  return 0;
}

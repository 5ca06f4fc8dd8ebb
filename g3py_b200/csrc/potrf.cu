// Blocked, batched fp64 Cholesky / triangular inverse / K^-1 and triangular solves for sm_100a.
//
// Replaces CholeskyRobust.perform -> scipy dpotrf (g3py/libs/tensors.py:198) and, for the
// gradient, CholeskyRobust.grad (Murray reverse mode, tensors.py:224-260) by the analytic route
// K^-1 = U U^T with U = L^-T.
//
// Layout: every matrix is Np x Np row-major (Np multiple of 128, identity on the padding),
// batch stride Np*Np.  Only the lower triangle of K / L is ever read.  U = L^-T is stored
// row-major upper, so that every level-3 step is the same "NT" GEMM (gemm.cu):
//   potrf :  A[i][j] -= sum_k L[i][k] L[j][k]^T            (left-looking inside an outer block of
//            L[i][j]  = A[i][j] Linv_jj^T                   `w` tile columns, right-looking between)
//   trtri :  U[j][i]  = -(sum_{k=j}^{i-1} U[j][k] L[i][k]^T) Linv_ii^T
//   lauum :  Kinv[i][j] = sum_{k>=i} U[i][k] U[j][k]^T      (i >= j)
// The 128x128 diagonal blocks are factored AND inverted by one CTA in shared memory
// (potrf_diag_kernel); their inverses (Dinv) turn every triangular solve into a GEMM/GEMV.
#include "g3b_internal.cuh"
#include <math.h>

namespace {

constexpr int TS = G3_TILE;       // 128
constexpr int LDS_ = TS + 1;      // padded row stride of the shared tile

// One CTA per matrix: factor the diagonal tile j in place and build Linv (-> Dinv).
// Blocked in shared memory with 32x32 sub-blocks; ~25 block-wide barriers in total (the first version eliminated
// one column per three barriers: 246 us per tile, all of it latency):
//   A  for kb = 0..3:  warp 0 factors the 32x32 diagonal sub-block in registers (row per lane, shuffles, rsqrt)
//                      and inverts it (column per lane); all warps then solve the sub-blocks below and update
//                      the rest with 4x4 register tiles (2 warps per 32x32 block)
//   B  off-diagonal sub-blocks of Linv by block forward substitution, X[I][J] = -Xd_I * sum_K L[I][K] X[K][J]
// Shared memory: S[128][129] holds L at S[i][c] (c <= i) and the off-diagonal blocks of X = Linv transposed in the
// upper blocks (X[i][c] = S[c][i]); XT[4][32][33] holds the inverses of the diagonal sub-blocks, dense and
// transposed (XT[b][c][k] = Xd_b[k][c], zero for k < c); Tm[3][32][33] is scratch.
constexpr int SB = 32;           // sub-block
constexpr int LDT = 33;          // scratch leading dimension

// Cholesky of a 32x32 sub-block by one warp, left-looking, lane = row: the lane keeps its row in registers
// (all indices static after unrolling) and publishes each finished entry to shared memory, from where the
// pivot row L[k][0..k) is read as a broadcast.  inv_out[k] = 1 / L[k][k].
__device__ __forceinline__ void warp_potrf32(double* blk, int lane, int& bad, double* inv_out) {
  double rw[SB];
#pragma unroll
  for (int c = 0; c < SB; ++c) rw[c] = blk[lane * LDS_ + c];     // entries above the diagonal are never used
#pragma unroll
  for (int k = 0; k < SB; ++k) {
    double p0 = rw[k], p1 = 0.0, p2 = 0.0, p3 = 0.0;
#pragma unroll
    for (int m = 0; m < SB; m += 4) {
      if (m + 0 < k) p0 -= rw[m + 0] * blk[k * LDS_ + m + 0];
      if (m + 1 < k) p1 -= rw[m + 1] * blk[k * LDS_ + m + 1];
      if (m + 2 < k) p2 -= rw[m + 2] * blk[k * LDS_ + m + 2];
      if (m + 3 < k) p3 -= rw[m + 3] * blk[k * LDS_ + m + 3];
    }
    const double v = (p0 + p1) + (p2 + p3);
    const double d = __shfl_sync(0xffffffffu, v, k);
    if (!(d > 0.0) && bad < 0) bad = k;
    const double inv = rsqrt(d);
    const double lv = (lane == k) ? d * inv : v * inv;   // sqrt(d) on the diagonal; NaN for d <= 0 (a failed pivot)
    rw[k] = lv;
    if (lane >= k) blk[lane * LDS_ + k] = lv;
    if (lane == k) inv_out[k] = inv;
    __syncwarp();
  }
}

// One thread per row: solve  y L^T = a  with the 32x32 lower-triangular block L (broadcast reads), y overwrites a.
__device__ __forceinline__ void row_trsolve32(double* prow, const double* Lblk, const double* inv) {
  double y[SB];
#pragma unroll
  for (int q = 0; q < SB; ++q) y[q] = prow[q];
#pragma unroll
  for (int q = 0; q < SB; ++q) {
    double p0 = y[q], p1 = 0.0, p2 = 0.0, p3 = 0.0;
#pragma unroll
    for (int c = 0; c < SB; c += 4) {
      if (c + 0 < q) p0 -= y[c + 0] * Lblk[q * LDS_ + c + 0];
      if (c + 1 < q) p1 -= y[c + 1] * Lblk[q * LDS_ + c + 1];
      if (c + 2 < q) p2 -= y[c + 2] * Lblk[q * LDS_ + c + 2];
      if (c + 3 < q) p3 -= y[c + 3] * Lblk[q * LDS_ + c + 3];
    }
    y[q] = ((p0 + p1) + (p2 + p3)) * inv[q];
  }
#pragma unroll
  for (int q = 0; q < SB; ++q) prow[q] = y[q];
}

// One thread per column c of Xd = L^-1 (32x32 lower): x[r] = Xd[r][c]; written to xt[r] (= XT[c][r]).
__device__ __forceinline__ void col_inverse32(const double* Lblk, const double* inv, int c, double* xt) {
  double x[SB];
#pragma unroll
  for (int r = 0; r < SB; ++r) {
    double p0 = (r == c) ? 1.0 : 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
#pragma unroll
    for (int k = 0; k < SB; k += 4) {
      if (k + 0 < r) p0 -= Lblk[r * LDS_ + k + 0] * x[k + 0];
      if (k + 1 < r) p1 -= Lblk[r * LDS_ + k + 1] * x[k + 1];
      if (k + 2 < r) p2 -= Lblk[r * LDS_ + k + 2] * x[k + 2];
      if (k + 3 < r) p3 -= Lblk[r * LDS_ + k + 3] * x[k + 3];
    }
    x[r] = (r >= c) ? ((p0 + p1) + (p2 + p3)) * inv[r] : 0.0;
  }
#pragma unroll
  for (int r = 0; r < SB; ++r) xt[r] = x[r];
}

__global__ void __launch_bounds__(256)
potrf_diag_kernel(double* __restrict__ A, int Np, long long strideA, int j, double* __restrict__ Dinv, int T,
                  double* __restrict__ U, double* __restrict__ logdet, int* __restrict__ info,
                  const int* __restrict__ bmap) {
  extern __shared__ double S[];                 // [128][129] | XT [4][32][33] | Tm [3][32][33]
  double* XT = S + TS * LDS_;
  double* Tm = XT + 4 * SB * LDT;
  __shared__ double invd[TS];
  __shared__ int first_bad;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = bmap ? bmap[blockIdx.x] : (int)blockIdx.x;
  double* At = A + (long long)b * strideA + (long long)j * TS * Np + (long long)j * TS;
  if (tid == 0) first_bad = -1;
  for (int idx = tid; idx < TS * TS; idx += 256) {
    const int r = idx >> 7, c = idx & 127;
    if (c <= r) S[r * LDS_ + c] = At[(long long)r * Np + c];
  }
  __syncthreads();

  // 4x4 register tile of a 32x32 block: 64 threads (slot g) per block
  const int g = tid >> 6, t64 = tid & 63, tr = (t64 >> 3) * 4, tc = (t64 & 7) * 4;

  // ---- phase A: blocked Cholesky ----------------------------------------------------------------
  for (int kb = 0; kb < TS / SB; ++kb) {
    const int o = kb * SB;
    if (warp == 0) {
      int bad = -1;
      warp_potrf32(S + o * LDS_ + o, lane, bad, invd + o);
      if (lane == 0 && bad >= 0 && first_bad < 0) first_bad = o + bad;
    }
    __syncthreads();
    const int nbelow = TS / SB - kb - 1;          // sub-blocks under the diagonal one
    if (nbelow > 0) {
      // A2: rows below: L[r][o..o+32) = A[r][o..o+32) L_kk^-T, one thread per row (triangular solve)
      if (tid < nbelow * SB) row_trsolve32(S + (o + SB + tid) * LDS_ + o, S + o * LDS_ + o, invd + o);
      __syncthreads();
      // A3: A[I][J] -= L[I][kb] L[J][kb]^T for kb < J <= I
      const int npair = nbelow * (nbelow + 1) / 2;
      for (int p0 = 0; p0 < npair; p0 += 4) {
        const int p = p0 + g;
        if (p < npair) {
          int I = 0, J = 0, cnt = 0;                       // enumerate pairs (I, J), J <= I, relative to kb + 1
          for (int ii = 0; ii < nbelow; ++ii)
            for (int jj = 0; jj <= ii; ++jj, ++cnt)
              if (cnt == p) { I = ii; J = jj; }
          const int rI = (kb + 1 + I) * SB + tr, rJ = (kb + 1 + J) * SB + tc;
          double c2[4][4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int q = 0; q < 4; ++q) c2[i][q] = 0.0;
          for (int k = 0; k < SB; ++k) {
            double pa[4], qb[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) pa[i] = S[(rI + i) * LDS_ + o + k];
#pragma unroll
            for (int q = 0; q < 4; ++q) qb[q] = S[(rJ + q) * LDS_ + o + k];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int q = 0; q < 4; ++q) c2[i][q] += pa[i] * qb[q];
          }
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (I != J || rJ + q <= rI + i) S[(rI + i) * LDS_ + rJ + q] -= c2[i][q];   // lower part only on diagonal blocks
        }
      }
      __syncthreads();
    }
  }

  // ---- phase B0: inverses of the four diagonal sub-blocks, one thread per column ------------------------
  if (tid < TS) {
    const int blk = tid >> 5, c = tid & 31;
    col_inverse32(S + (blk * SB) * LDS_ + blk * SB, invd + blk * SB, c, XT + (blk * SB + c) * LDT);
  }
  __syncthreads();

  // ---- phase B: off-diagonal sub-blocks of X = L^-1 --------------------------------------------------
  for (int dist = 1; dist < TS / SB; ++dist) {
    const int nblk = TS / SB - dist;
    const int J = g, I = g + dist;
    if (g < nblk) {
      // T[rr][cc] = sum_{k in [J*32, I*32)} L[I*32+rr][k] X[k][J*32+cc]
      double acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[i][q] = 0.0;
      for (int k = 0; k < SB; ++k) {                         // K == J: X[k][c] = XT[J][c][k]
        double pa[4], qb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) pa[i] = S[(I * SB + tr + i) * LDS_ + J * SB + k];
#pragma unroll
        for (int q = 0; q < 4; ++q) qb[q] = XT[(J * SB + tc + q) * LDT + k];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[i][q] += pa[i] * qb[q];
      }
      for (int k = (J + 1) * SB; k < I * SB; ++k) {          // K > J: X[k][c] = S[c][k]
        double pa[4], qb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) pa[i] = S[(I * SB + tr + i) * LDS_ + k];
#pragma unroll
        for (int q = 0; q < 4; ++q) qb[q] = S[(J * SB + tc + q) * LDS_ + k];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[i][q] += pa[i] * qb[q];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q) Tm[(g * SB + tr + i) * LDT + tc + q] = acc[i][q];
    }
    __syncthreads();
    if (g < nblk) {
      // X[I][J] = -Xd_I T :  out[rr][cc] = -sum_m XT[I][m][rr] * T[m][cc]   ->  S[c][r]
      double acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[i][q] = 0.0;
      for (int m = 0; m < SB; ++m) {
        double pa[4], qb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) pa[i] = XT[(I * SB + m) * LDT + tr + i];
#pragma unroll
        for (int q = 0; q < 4; ++q) qb[q] = Tm[(g * SB + m) * LDT + tc + q];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[i][q] += pa[i] * qb[q];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q) S[(J * SB + tc + q) * LDS_ + I * SB + tr + i] = -acc[i][q];
    }
    __syncthreads();
  }

  // write back: L tile (upper zeroed), Dinv (lower, zeros above)
  double* Dj = Dinv + ((long long)b * T + j) * TS * TS;
  for (int idx = tid; idx < TS * TS; idx += 256) {
    const int r = idx >> 7, c = idx & 127;
    At[(long long)r * Np + c] = (c <= r) ? S[r * LDS_ + c] : 0.0;
    double xv = 0.0;
    if (c <= r) xv = ((r >> 5) == (c >> 5)) ? XT[((r >> 5) * SB + (c & 31)) * LDT + (r & 31)] : S[c * LDS_ + r];
    Dj[idx] = xv;
  }
  (void)U;
  if (tid < 32) {
    double acc = 0.0;
    for (int k = tid; k < TS; k += 32) acc += log(S[k * LDS_ + k]);
    for (int o2 = 16; o2 > 0; o2 >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o2);
    if (tid == 0) {
      if (logdet) logdet[b] += acc;
      if (info && first_bad >= 0 && info[b] == 0) info[b] = j * TS + first_bad + 1;
    }
  }
}

// Forward substitution step j of u = L^-1 r (right-looking).  grid (T-j, B).
// CTA x=0 finalises u_j = Linv_jj r_j and accumulates beta += |u_j|^2;
// CTA x>0 recomputes u_j and updates r_{j+x} -= L[j+x][j] u_j.
__global__ void __launch_bounds__(256)
trsv_fwd_step_kernel(const double* __restrict__ L, const double* __restrict__ Dinv, double* __restrict__ r,
                     double* __restrict__ u, double* __restrict__ beta, int j, int Np, int T) {
  __shared__ double rj[TS], uj[TS];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y, x = blockIdx.x;
  const double* Dj = Dinv + ((long long)b * T + j) * TS * TS;
  double* rb = r + (long long)b * Np;
  if (tid < TS) rj[tid] = rb[j * TS + tid];
  __syncthreads();
  {  // warp per row, coalesced 1 KiB rows; all 16 rows of a warp are in flight before the first reduction
    double4 v[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = *reinterpret_cast<const double4*>(Dj + (warp + 8 * q) * TS + lane * 4);
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      double acc = v[q].x * rj[lane * 4] + v[q].y * rj[lane * 4 + 1] + v[q].z * rj[lane * 4 + 2] + v[q].w * rj[lane * 4 + 3];
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) uj[warp + 8 * q] = acc;
    }
  }
  __syncthreads();
  if (x == 0) {
    if (tid < TS) u[(long long)b * Np + j * TS + tid] = uj[tid];
    if (warp == 0 && beta) {
      double acc = 0.0;
      for (int k = lane; k < TS; k += 32) acc += uj[k] * uj[k];
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
      if (lane == 0) beta[b] += acc;
    }
    return;
  }
  const int i = j + x;
  const double* Lt = L + (long long)b * Np * Np + (long long)i * TS * Np + (long long)j * TS;
  double4 v[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) v[q] = *reinterpret_cast<const double4*>(Lt + (long long)(warp + 8 * q) * Np + lane * 4);
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    double acc = v[q].x * uj[lane * 4] + v[q].y * uj[lane * 4 + 1] + v[q].z * uj[lane * 4 + 2] + v[q].w * uj[lane * 4 + 3];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) rb[i * TS + warp + 8 * q] -= acc;
  }
}

// Backward substitution step j (j = T-1 .. 0) of alpha = L^-T s.  grid (j+1, B).
// CTA x=0 finalises alpha_j = Linv_jj^T s_j; CTA x>0 updates s_{i} -= L[j][i]^T alpha_j, i = x-1.
__global__ void __launch_bounds__(256)
trsv_bwd_step_kernel(const double* __restrict__ L, const double* __restrict__ Dinv, double* __restrict__ s,
                     double* __restrict__ alpha, int j, int Np, int T) {
  __shared__ double sj[TS], aj[TS], part[2][TS];
  const int tid = threadIdx.x;
  const int b = blockIdx.y, x = blockIdx.x;
  const double* Dj = Dinv + ((long long)b * T + j) * TS * TS;
  double* sb = s + (long long)b * Np;
  if (tid < TS) sj[tid] = sb[j * TS + tid];
  __syncthreads();
  const int c = tid & 127, half = tid >> 7;       // thread per column, two row halves
  {
    double acc = 0.0;
    for (int a0 = half * 64; a0 < half * 64 + 64; a0 += 16) {   // 16 loads in flight, same summation order
      double t[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) t[q] = Dj[(a0 + q) * TS + c];
#pragma unroll
      for (int q = 0; q < 16; ++q) acc += t[q] * sj[a0 + q];
    }
    part[half][c] = acc;
  }
  __syncthreads();
  if (tid < TS) aj[tid] = part[0][tid] + part[1][tid];
  __syncthreads();
  if (x == 0) {
    if (tid < TS) alpha[(long long)b * Np + j * TS + tid] = aj[tid];
    return;
  }
  const int i = x - 1;
  const double* Lt = L + (long long)b * Np * Np + (long long)j * TS * Np + (long long)i * TS;
  {
    double acc = 0.0;
    for (int a0 = half * 64; a0 < half * 64 + 64; a0 += 16) {
      double t[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) t[q] = Lt[(long long)(a0 + q) * Np + c];
#pragma unroll
      for (int q = 0; q < 16; ++q) acc += t[q] * aj[a0 + q];
    }
    part[half][c] = acc;
  }
  __syncthreads();
  if (tid < TS) sb[i * TS + tid] -= part[0][tid] + part[1][tid];
}


// ---- whole triangular solves in ONE launch ---------------------------------------------------------------------------
// The step kernels above make a substitution a chain of T dependent launches (10 us each: 0.32 ms of a 1.4 ms evaluation
// at N = 2048).  Here one CTA owns one 128-row block of one right-hand side for the whole solve: it subtracts
// L[i][j] u_j as the u_j appear (release / acquire flags in global memory), then applies the inverse of its diagonal tile
// and publishes u_i.  CTAs take their (item, block) from a ticket counter in dependency order, so every block a CTA waits
// for has already started: no co-residency requirement.  sync[0] = ticket counter, sync[1 + item*T + block] = flags.
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(256)
trsv_fwd_fused_kernel(const double* __restrict__ L, const double* __restrict__ Dinv, const double* __restrict__ r,
                      double* __restrict__ u, double* __restrict__ beta, int Np, int T, int* __restrict__ sync,
                      double* __restrict__ part) {
  __shared__ int s_t;
  __shared__ double rs[TS], us[TS];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) s_t = atomicAdd(&sync[0], 1);
  __syncthreads();
  const int b = s_t / T, i = s_t % T;
  int* flags = sync + 1 + (long long)b * T;
  const double* Lb = L + (long long)b * Np * Np;
  const double* ub = u + (long long)b * Np;
  if (tid < TS) rs[tid] = r[(long long)b * Np + i * TS + tid];
  // the inverse of the diagonal tile is needed last and depends on nothing: fetched first
  double4 dv[16];
  {
    const double* Dj = Dinv + ((long long)b * T + i) * TS * TS;
#pragma unroll
    for (int q = 0; q < 16; ++q) dv[q] = *reinterpret_cast<const double4*>(Dj + (warp + 8 * q) * TS + lane * 4);
  }
  for (int j = 0; j < i; ++j) {
    // the tile does not depend on the flag: its loads are in flight while this CTA waits
    const double* Lt = Lb + (long long)i * TS * Np + (long long)j * TS;
    double4 v[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = *reinterpret_cast<const double4*>(Lt + (long long)(warp + 8 * q) * Np + lane * 4);
    if (tid == 0)
      while (ld_acquire(flags + j) == 0) {}
    __syncthreads();
    if (tid < TS) us[tid] = __ldcg(ub + j * TS + tid);
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      double acc = v[q].x * us[lane * 4] + v[q].y * us[lane * 4 + 1] + v[q].z * us[lane * 4 + 2] + v[q].w * us[lane * 4 + 3];
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) rs[warp + 8 * q] -= acc;
    }
  }
  {
    __syncthreads();                       // rs complete; everybody is past its last read of us
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      double acc = dv[q].x * rs[lane * 4] + dv[q].y * rs[lane * 4 + 1] + dv[q].z * rs[lane * 4 + 2] + dv[q].w * rs[lane * 4 + 3];
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) us[warp + 8 * q] = acc;
    }
  }
  __syncthreads();
  if (tid < TS) u[(long long)b * Np + i * TS + tid] = us[tid];
  double mine = 0.0;
  if (warp == 0) {
    for (int k = lane; k < TS; k += 32) mine += us[k] * us[k];
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, o);
    if (lane == 0) part[(long long)b * T + i] = mine;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    st_release(flags + i, 1);
    if (i == T - 1 && beta) {              // every earlier flag of this item has been acquired: add the partial sums in block order
      double sum = 0.0;
      for (int k = 0; k < T - 1; ++k) sum += __ldcg(part + (long long)b * T + k);
      beta[b] += sum + mine;
    }
  }
}

// alpha = L^-T s, blocks from the last to the first: alpha_i = Linv_ii^T (s_i - sum_{j > i} L[j][i]^T alpha_j)
__global__ void __launch_bounds__(512)
trsv_bwd_fused_kernel(const double* __restrict__ L, const double* __restrict__ Dinv, const double* __restrict__ s,
                      double* __restrict__ alpha, int Np, int T, int* __restrict__ sync) {
  __shared__ int s_t;
  __shared__ double ss[TS], as[TS], pt[4][TS];
  const int tid = threadIdx.x;
  if (tid == 0) s_t = atomicAdd(&sync[0], 1);
  __syncthreads();
  const int b = s_t / T, i = T - 1 - s_t % T;
  int* flags = sync + 1 + (long long)b * T;
  const double* Lb = L + (long long)b * Np * Np;
  const double* ab = alpha + (long long)b * Np;
  const int c = tid & 127, qr = tid >> 7;  // thread = column c, rows qr*32 .. +31 of a tile
  double dv[32];                           // the inverse of the diagonal tile is needed last and depends on nothing: fetched first
  {
    const double* Dj = Dinv + ((long long)b * T + i) * TS * TS;
#pragma unroll
    for (int q = 0; q < 32; ++q) dv[q] = Dj[(qr * 32 + q) * TS + c];
  }
  double acc = 0.0;                        // this thread's quarter of sum_j (L[j][i]^T alpha_j)[c]
  for (int j = T - 1; j > i; --j) {
    // the tile does not depend on the flag: its loads are in flight while this CTA waits
    const double* Lt = Lb + (long long)j * TS * Np + (long long)i * TS;
    double t[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) t[q] = Lt[(long long)(qr * 32 + q) * Np + c];
    if (tid == 0)
      while (ld_acquire(flags + j) == 0) {}
    __syncthreads();
    if (tid < TS) as[tid] = __ldcg(ab + j * TS + tid);
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 32; ++q) acc += t[q] * as[qr * 32 + q];
  }
  pt[qr][c] = acc;
  if (tid < TS) ss[tid] = s[(long long)b * Np + i * TS + tid];
  __syncthreads();
  if (tid < TS) ss[tid] -= (pt[0][tid] + pt[1][tid]) + (pt[2][tid] + pt[3][tid]);
  __syncthreads();
  {
    double a2 = 0.0;
#pragma unroll
    for (int q = 0; q < 32; ++q) a2 += dv[q] * ss[qr * 32 + q];
    pt[qr][c] = a2;
  }
  __syncthreads();
  if (tid < TS) alpha[(long long)b * Np + i * TS + tid] = (pt[0][tid] + pt[1][tid]) + (pt[2][tid] + pt[3][tid]);
  __threadfence();
  __syncthreads();
  if (tid == 0) st_release(flags + i, 1);
}

}  // namespace

static constexpr int kDiagSmem = (TS * LDS_ + 7 * SB * LDT) * (int)sizeof(double);

// Factor + invert the diagonal tile j of B matrices: the low-latency kernel of diag.cu unless the first one is asked for.
static int g3_diag_launch(g3_ctx* ctx, double* A, int Np, long long strideA, int j, double* Dinv, int T, double* logdet,
                          int* info, const int* bmap, int B) {
  g3_prof_begin(ctx, G3_PROF_DIAG);
  if (ctx->diag_variant == 1) {
    if (!ctx->diag_ready) {
      G3_CUDA(ctx, cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDiagSmem));
      ctx->diag_ready = true;
    }
    potrf_diag_kernel<<<B, 256, kDiagSmem, ctx->stream>>>(A, Np, strideA, j, Dinv, T, nullptr, logdet, info, bmap);
  } else {
    const int rc = g3_diag2_launch(ctx, A, Np, strideA, j, Dinv, T, logdet, info, bmap, B, nullptr);
    if (rc) return rc;
  }
  g3_prof_end(ctx);
  G3_LAUNCH_CHECK(ctx);
  return 0;
}

static GemmArgs gemm_zero() {
  GemmArgs g;
  memset(&g, 0, sizeof g);
  return g;
}

// Outer-block width (tile columns) of the blocked schedules for few matrices, from sweeps on B200
// (tools/sweep_block.py, tools/sweep_batch.py): up to N=2048 every column is its own block (rank-128 trailing updates
// keep all SMs busy and the look-ahead / pipelined trtri overlap the per-column chains), at N=4096 too for one or two
// matrices, 4 for N=4096 with 3..8 matrices and at N=8192, 8 from N=16384 on.
static int g3_auto_block(const g3_ctx* ctx, int T, int B) {
  if (T <= 16) return 1;
  if (T <= 32) return B <= 2 ? 1 : 4;
  return T <= 64 ? 4 : ctx->potrf_w_big;
}
// Up to 8 matrices go by outer blocks (look-ahead, pipelined trtri); larger batches supply the parallelism
// themselves and stay left-looking (tools/sweep_batch.py).  g3_gp_run decides on the WHOLE batch and pins the choice
// for its stream groups (8 items each) through ctx->force_left.
static bool g3_use_blocked(const g3_ctx* ctx, int B) { return !ctx->force_left && B <= 8; }

static int trtri_block(g3_ctx* ctx, const CUtensorMap& tmU, const CUtensorMap& tmL, const CUtensorMap& tmD, double* U,
                       int Np, int B, int io, int ie);
static void launch_u_diag(g3_ctx* ctx, const double* Dinv, double* U, int Np, int T, int B, int j0, int nj);

static int potrf_batched_impl(g3_ctx* ctx, double* A, int Np, int Btotal, double* Dinv, double* logdet, int* info,
                              const int* bmap, int nb, int w_outer, double* U_pipe);

int g3_potrf_batched(g3_ctx* ctx, double* A, int Np, int Btotal, double* Dinv, double* logdet, int* info,
                     const int* bmap, int nb, int w_outer, double* U_pipe) {
  const int T = Np / TS, B = bmap ? nb : Btotal;
  int rc;
  if (ctx->diag_variant == 2 && (rc = g3_diag2_prepare(ctx, 0, T, Dinv, T, bmap, B))) return rc;
  if ((rc = potrf_batched_impl(ctx, A, Np, Btotal, Dinv, logdet, info, bmap, nb, w_outer, U_pipe))) return rc;
  if (ctx->diag_variant == 2 && (rc = g3_diag2_finish(ctx, A, Np, (long long)Np * Np, 0, T, bmap, B))) return rc;
  return 0;
}

static int potrf_batched_impl(g3_ctx* ctx, double* A, int Np, int Btotal, double* Dinv, double* logdet, int* info,
                              const int* bmap, int nb, int w_outer, double* U_pipe) {
  G3_NVTX("g3:potrf");
  ctx->trtri_done = 0;
  const int T = Np / TS;
  const int Blaunch = bmap ? nb : Btotal;
  if (w_outer < 1) {
    // auto: fully left-looking when the batch supplies the parallelism (a column update launches 2*B*(T-j) CTAs);
    // few matrices would leave most SMs idle in the column updates, so they go right-looking between outer blocks
    w_outer = g3_use_blocked(ctx, Blaunch) ? g3_auto_block(ctx, T, Blaunch) : (1 << 20);
  }
  // The tensor maps span the whole allocation (Btotal matrices); bmap (nb entries) picks the batch
  // coordinate of each launched CTA column, so the jitter ladder can refactor a subset in place.
  const int B = Blaunch;
  CUtensorMap tmA, tmB, tmD;
  const uint64_t batch_extent = (uint64_t)Btotal;
  int rc;
  if ((rc = g3_make_tmap(ctx, &tmA, A, Np, Np, batch_extent, Np, (uint64_t)Np * Np, G3_BM))) return rc;
  if ((rc = g3_make_tmap(ctx, &tmB, A, Np, Np, batch_extent, Np, (uint64_t)Np * Np, G3_BN))) return rc;
  if ((rc = g3_make_tmap(ctx, &tmD, Dinv, TS, (uint64_t)T * TS, batch_extent, TS, (uint64_t)T * TS * TS, G3_BN))) return rc;
  const long long strideA = (long long)Np * Np;

  // factor the tile columns [jo, je) (left-looking inside the block; earlier blocks were applied right-looking)
  auto factor_block = [&](int jo, int je) -> int {
    for (int j = jo; j < je; ++j) {
      if (j > jo) {  // left-looking update of tile column j with the columns of this outer block
        GemmArgs g = gemm_zero();
        g.D = A; g.ldd = Np; g.strideD = strideA;
        g.mode = 0; g.ntx = T - j; g.nty = 1;
        g.d_r0 = j * TS; g.d_c0 = j * TS;
        g.a_r0 = j * TS; g.a_rx = TS;
        g.b_r0 = j * TS;
        g.ka0 = jo * TS; g.kb0 = jo * TS; g.kl0 = (j - jo) * TS;
        g.alpha = -1.0; g.beta = 1.0; g.bmap = bmap; g.upper = 2;
        if ((rc = g3_gemm_launch(ctx, tmA, tmB, g, B))) return rc;
      }
      if ((rc = g3_diag_launch(ctx, A, Np, strideA, j, Dinv, T, logdet, info, bmap, B))) return rc;
      if (j < T - 1) {  // L[i][j] = A[i][j] Linv_jj^T for the tiles below the diagonal
        GemmArgs g = gemm_zero();
        g.D = A; g.ldd = Np; g.strideD = strideA;
        g.mode = 0; g.ntx = T - j - 1; g.nty = 1;
        g.d_r0 = (j + 1) * TS; g.d_c0 = j * TS;
        g.a_r0 = (j + 1) * TS; g.a_rx = TS; g.ka0 = j * TS;
        g.b_r0 = j * TS; g.kb0 = 0;
        g.kl0 = TS;
        g.alpha = 1.0; g.beta = 0.0; g.bmap = bmap; g.tri_b = 1;
        if ((rc = g3_gemm_launch(ctx, tmA, tmD, g, B))) return rc;
      }
    }
    return 0;
  };
  // right-looking update with the finished block [jo, je): tile columns [c0, c1) of the trailing matrix, rows >= c0
  auto trailing = [&](int jo, int je, int c0, int c1) -> int {
    GemmArgs g = gemm_zero();
    g.D = A; g.ldd = Np; g.strideD = strideA;
    g.a_rx = TS; g.b_ry = TS;
    g.ka0 = jo * TS; g.kb0 = jo * TS; g.kl0 = (je - jo) * TS;
    g.alpha = -1.0; g.beta = 1.0; g.bmap = bmap;
    g.d_r0 = c0 * TS; g.d_c0 = c0 * TS; g.a_r0 = c0 * TS; g.b_r0 = c0 * TS;
    if (c1 >= T) {          // everything to the right: lower triangle of tiles
      g.mode = 1; g.ntx = T - c0; g.upper = 1;
    } else {                // a block of columns: tiles above the block diagonal are void
      g.mode = 0; g.ntx = T - c0; g.nty = c1 - c0; g.upper = 1 | 4;
    }
    return g3_gemm_launch(ctx, tmA, tmB, g, B);
  };

  // ---- batched, fully left-looking, int8 tensor-core mode: 256-wide block columns; the update of a block column with ALL
  // earlier columns (contraction depth jo * 128) runs as exact int8 slice products (ozaki.cu) once it is deep enough, the
  // work inside the block column (depth 128) and the triangular solves stay on the DMMA GEMM.
  if (ctx->gemm_mode == G3_GEMM_OZAKI && w_outer >= T && bmap == nullptr && T >= 4) {
    g3_oz_state oz;
    if ((rc = g3_oz_prepare(ctx, A, Np, B, &oz))) return rc;
    for (int jo = 0; jo < T; jo += 2) {
      const int je = jo + 2 < T ? jo + 2 : T;
      if (jo > 0) {
        if (jo * TS >= ctx->oz_min_k) {
          if ((rc = g3_oz_update(ctx, &oz, A, jo, je))) return rc;
          ctx->oz_launches++;
        } else {  // shallow contraction: one DMMA launch for both tile columns, rows >= jo
          GemmArgs g = gemm_zero();
          g.D = A; g.ldd = Np; g.strideD = strideA;
          g.mode = 0; g.ntx = T - jo; g.nty = je - jo;
          g.d_r0 = jo * TS; g.d_c0 = jo * TS;
          g.a_r0 = jo * TS; g.a_rx = TS;
          g.b_r0 = jo * TS; g.b_ry = TS;
          g.ka0 = 0; g.kb0 = 0; g.kl0 = jo * TS;
          g.alpha = -1.0; g.beta = 1.0; g.upper = 1 | 4;
          if ((rc = g3_gemm_launch(ctx, tmA, tmB, g, B))) return rc;
        }
      }
      if ((rc = factor_block(jo, je))) return rc;
      if (je < T && (rc = g3_oz_slice(ctx, &oz, A, jo, je))) return rc;
    }
    return 0;
  }
  const bool look = ctx->lookahead && w_outer < T;              // at least two outer blocks
  if (!look) {
    for (int jo = 0; jo < T; jo += w_outer) {
      const int je = jo + w_outer < T ? jo + w_outer : T;
      if ((rc = factor_block(jo, je))) return rc;
      if (je < T && (rc = trailing(jo, je, je, T))) return rc;
    }
    return 0;
  }
  // ---- look-ahead: panel stream P factors block k+1 while the main stream finishes the trailing update of block k
  if (!ctx->panel_stream) {
    int lo = 0, hi = 0;
    G3_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    G3_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->panel_stream, cudaStreamNonBlocking, hi));
    G3_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_panel, cudaEventDisableTiming));
    G3_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_main, cudaEventDisableTiming));
  }
  cudaStream_t main_stream = ctx->stream, P = ctx->panel_stream;
  G3_CUDA(ctx, cudaEventRecord(ctx->ev_main, main_stream));      // the Gram matrix is ready
  G3_CUDA(ctx, cudaStreamWaitEvent(P, ctx->ev_main, 0));
  // U = L^-T pipelined behind the factorisation (gradient path, whole batch): rows [jo, je) of L^-1 only need L final
  // for rows < je, i.e. block [jo, je) factored, and read columns < je of A, which later trailing updates never touch.
  // They run on a third stream while the panel / main streams go on with the next blocks.
  const bool pipe = U_pipe != nullptr && bmap == nullptr && ctx->trtri_pipeline;
  CUtensorMap tmU;
  cudaStream_t Q = nullptr;
  if (pipe) {
    if (!ctx->tri_stream) {
      G3_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->tri_stream, cudaStreamNonBlocking));
      G3_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_tri, cudaEventDisableTiming));
    }
    Q = ctx->tri_stream;
    if ((rc = g3_make_tmap(ctx, &tmU, U_pipe, Np, Np, batch_extent, Np, (uint64_t)Np * Np, G3_BM))) return rc;
    G3_CUDA(ctx, cudaStreamWaitEvent(Q, ctx->ev_main, 0));        // earlier users of U (previous call) are done
  }
  rc = 0;
  for (int jo = 0; jo < T && !rc; jo += w_outer) {
    const int je = jo + w_outer < T ? jo + w_outer : T;
    const int je2 = je + w_outer < T ? je + w_outer : T;
    ctx->stream = P;
    rc = factor_block(jo, je);
    if (!rc) cudaEventRecord(ctx->ev_panel, P);
    if (!rc && pipe) {
      cudaEventRecord(ctx->ev_tri, P);
      cudaStreamWaitEvent(Q, ctx->ev_tri, 0);
      ctx->stream = Q;
      launch_u_diag(ctx, Dinv, U_pipe, Np, T, B, jo, je - jo);
      rc = trtri_block(ctx, tmU, tmB, tmD, U_pipe, Np, B, jo, je);
      ctx->stream = P;
    }
    if (!rc && je < T) {
      if (jo > 0) cudaStreamWaitEvent(P, ctx->ev_main, 0);         // the previous block's main-stream update wrote these columns
      rc = trailing(jo, je, je, je2);                              // next panel's columns first, on P
      ctx->stream = main_stream;
      cudaStreamWaitEvent(main_stream, ctx->ev_panel, 0);
      if (!rc && je2 < T) rc = trailing(jo, je, je2, T);           // the rest, on the main stream
      cudaEventRecord(ctx->ev_main, main_stream);
    }
  }
  ctx->stream = main_stream;
  cudaEventRecord(ctx->ev_panel, P);
  cudaStreamWaitEvent(main_stream, ctx->ev_panel, 0);
  if (pipe) {
    cudaEventRecord(ctx->ev_tri, Q);
    cudaStreamWaitEvent(main_stream, ctx->ev_tri, 0);
    if (!rc) ctx->trtri_done = 1;
  }
  return rc;
}

int g3_potrf_panel(g3_ctx* ctx, double* P, int rows, int nb, double* Dinv, double* logdet, int* info) {
  if (rows % TS || nb % TS || rows < nb) return g3_fail_msg(ctx, "potrf_panel: rows/nb must be multiples of 128, rows >= nb");
  const int Tr = rows / TS, w = nb / TS;
  CUtensorMap tmA, tmB, tmD;
  int rc;
  if ((rc = g3_make_tmap(ctx, &tmA, P, nb, rows, 1, nb, (uint64_t)rows * nb, G3_BM))) return rc;
  if ((rc = g3_make_tmap(ctx, &tmB, P, nb, rows, 1, nb, (uint64_t)rows * nb, G3_BN))) return rc;
  if ((rc = g3_make_tmap(ctx, &tmD, Dinv, TS, (uint64_t)w * TS, 1, TS, (uint64_t)w * TS * TS, G3_BN))) return rc;
  if (ctx->diag_variant == 2 && (rc = g3_diag2_prepare(ctx, 0, w, Dinv, w, nullptr, 1))) return rc;
  for (int j = 0; j < w; ++j) {
    if (j > 0) {  // left-looking inside the panel
      GemmArgs g = gemm_zero();
      g.D = P; g.ldd = nb; g.strideD = 0;
      g.mode = 0; g.ntx = Tr - j; g.nty = 1;
      g.d_r0 = j * TS; g.d_c0 = j * TS;
      g.a_r0 = j * TS; g.a_rx = TS;
      g.b_r0 = j * TS;
      g.ka0 = 0; g.kb0 = 0; g.kl0 = j * TS;
      g.alpha = -1.0; g.beta = 1.0; g.upper = 2;
      if ((rc = g3_gemm_launch(ctx, tmA, tmB, g, 1))) return rc;
    }
    if ((rc = g3_diag_launch(ctx, P, nb, 0, j, Dinv, w, logdet, info, nullptr, 1))) return rc;
    if (Tr - j - 1 > 0) {
      GemmArgs g = gemm_zero();
      g.D = P; g.ldd = nb; g.strideD = 0;
      g.mode = 0; g.ntx = Tr - j - 1; g.nty = 1;
      g.d_r0 = (j + 1) * TS; g.d_c0 = j * TS;
      g.a_r0 = (j + 1) * TS; g.a_rx = TS; g.ka0 = j * TS;
      g.b_r0 = j * TS; g.kb0 = 0;
      g.kl0 = TS;
      g.alpha = 1.0; g.beta = 0.0; g.tri_b = 1;
      if ((rc = g3_gemm_launch(ctx, tmA, tmD, g, 1))) return rc;
    }
  }
  if (ctx->diag_variant == 2 && (rc = g3_diag2_finish(ctx, P, nb, 0, 0, w, nullptr, 1))) return rc;
  return 0;
}

int g3_syrk_panel(g3_ctx* ctx, const double* P, int rowsP, int nb, int row_off, double* D, int rowsD) {
  if (rowsP % TS || nb % TS || rowsD % TS || row_off % TS || row_off + rowsD > rowsP)
    return g3_fail_msg(ctx, "syrk_panel: bad geometry");
  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = g3_make_tmap(ctx, &tmA, P, nb, rowsP, 1, nb, (uint64_t)rowsP * nb, G3_BM))) return rc;
  if ((rc = g3_make_tmap(ctx, &tmB, P, nb, rowsP, 1, nb, (uint64_t)rowsP * nb, G3_BN))) return rc;
  GemmArgs g = gemm_zero();
  g.D = D; g.ldd = nb; g.strideD = 0;
  g.mode = 0; g.ntx = rowsD / TS; g.nty = nb / TS;
  g.a_r0 = row_off; g.a_rx = TS;
  g.b_r0 = row_off; g.b_ry = TS;
  g.kl0 = nb;
  g.alpha = -1.0; g.beta = 1.0; g.upper = 1 | 4;
  return g3_gemm_launch(ctx, tmA, tmB, g, 1);
}

// Rows of a panel below (or beside) an already factored nb x nb diagonal block Ld (lower, ld = nb) whose 128x128 block
// inverses are Dinv:  P <- P Ld^-T, tile column by tile column (left-looking), all through the NT GEMM.  P: rows x nb.
// This is the part of g3_potrf_panel that does not need the diagonal kernel; in the 2-D block-cyclic layout the
// process rows that do not own the diagonal block run it on a received copy of Ld / Dinv.
int g3_panel_solve(g3_ctx* ctx, double* P, int rows, int nb, const double* Ld, const double* Dinv) {
  if (rows % TS || nb % TS || rows <= 0) return g3_fail_msg(ctx, "panel_solve: rows/nb must be positive multiples of 128");
  const int Tr = rows / TS, w = nb / TS;
  CUtensorMap tmA, tmL, tmD;
  int rc;
  if ((rc = g3_make_tmap(ctx, &tmA, P, nb, rows, 1, nb, (uint64_t)rows * nb, G3_BM))) return rc;
  if ((rc = g3_make_tmap(ctx, &tmL, Ld, nb, nb, 1, nb, (uint64_t)nb * nb, G3_BN))) return rc;
  if ((rc = g3_make_tmap(ctx, &tmD, Dinv, TS, (uint64_t)w * TS, 1, TS, (uint64_t)w * TS * TS, G3_BN))) return rc;
  for (int j = 0; j < w; ++j) {
    if (j > 0) {  // P[:, j] -= P[:, 0:j] Ld[j, 0:j]^T
      GemmArgs g = gemm_zero();
      g.D = P; g.ldd = nb; g.strideD = 0;
      g.mode = 0; g.ntx = Tr; g.nty = 1;
      g.d_r0 = 0; g.d_c0 = j * TS;
      g.a_r0 = 0; g.a_rx = TS; g.ka0 = 0;
      g.b_r0 = j * TS; g.kb0 = 0;
      g.kl0 = j * TS;
      g.alpha = -1.0; g.beta = 1.0;
      if ((rc = g3_gemm_launch(ctx, tmA, tmL, g, 1))) return rc;
    }
    GemmArgs g = gemm_zero();  // P[:, j] = P[:, j] Linv_jj^T
    g.D = P; g.ldd = nb; g.strideD = 0;
    g.mode = 0; g.ntx = Tr; g.nty = 1;
    g.d_r0 = 0; g.d_c0 = j * TS;
    g.a_r0 = 0; g.a_rx = TS; g.ka0 = j * TS;
    g.b_r0 = j * TS; g.kb0 = 0;
    g.kl0 = TS;
    g.alpha = 1.0; g.beta = 0.0; g.tri_b = 1;
    if ((rc = g3_gemm_launch(ctx, tmA, tmD, g, 1))) return rc;
  }
  return 0;
}

// D[x][y] -= sum_k A[x][k] Bm[y][k]   (D, A: rows x nb, ld = nb; Bm: nb x nb).  has_diag: the first nb rows of D are a
// diagonal block of the symmetric matrix (only its lower triangle is needed: tiles above it are skipped).
int g3_panel_update(g3_ctx* ctx, double* D, int rows, int nb, const double* A, const double* Bm, int has_diag) {
  if (rows % TS || nb % TS || rows <= 0) return g3_fail_msg(ctx, "panel_update: bad geometry");
  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = g3_make_tmap(ctx, &tmA, A, nb, rows, 1, nb, (uint64_t)rows * nb, G3_BM))) return rc;
  if ((rc = g3_make_tmap(ctx, &tmB, Bm, nb, nb, 1, nb, (uint64_t)nb * nb, G3_BN))) return rc;
  GemmArgs g = gemm_zero();
  g.D = D; g.ldd = nb; g.strideD = 0;
  g.mode = 0; g.ntx = rows / TS; g.nty = nb / TS;
  g.a_r0 = 0; g.a_rx = TS;
  g.b_r0 = 0; g.b_ry = TS;
  g.kl0 = nb;
  g.alpha = -1.0; g.beta = 1.0; g.upper = has_diag ? (1 | 4) : 0;
  return g3_gemm_launch(ctx, tmA, tmB, g, 1);
}

// General panel GEMM: D[x][y] = beta D[x][y] + alpha sum_k A[x][k] Bm[y][k]   (D: rows x ncols with leading dimension ldd,
// A: rows x kdim (lda), Bm: ncols x kdim (ldb)); everything a multiple of 128 (kdim of 16).
int g3_panel_gemm(g3_ctx* ctx, double* D, long long ldd, int rows, int ncols, const double* A, long long lda, const double* Bm,
                  long long ldb, long long kdim, double alpha, double beta) {
  if (rows % TS || ncols % TS || kdim % G3_BK || rows <= 0 || ncols <= 0 || kdim <= 0) return g3_fail_msg(ctx, "panel_gemm: bad geometry");
  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = g3_make_tmap(ctx, &tmA, A, (uint64_t)kdim, rows, 1, (uint64_t)lda, (uint64_t)rows * lda, G3_BM))) return rc;
  if ((rc = g3_make_tmap(ctx, &tmB, Bm, (uint64_t)kdim, ncols, 1, (uint64_t)ldb, (uint64_t)ncols * ldb, G3_BN))) return rc;
  GemmArgs g = gemm_zero();
  g.D = D; g.ldd = ldd; g.strideD = 0;
  g.mode = 0; g.ntx = rows / TS; g.nty = ncols / TS;
  g.a_r0 = 0; g.a_rx = TS;
  g.b_r0 = 0; g.b_ry = TS;
  g.kl0 = (int)kdim;
  g.alpha = alpha; g.beta = beta;
  return g3_gemm_launch(ctx, tmA, tmB, g, 1);
}

// Right-hand triangular solve with a NON-transposed lower factor: Y <- s_mul * (Y' L^-1) computed tile column by tile column from
// the last to the first,  T_j = Y[:, j] + s_upd * sum_{k > j} Z[:, k] L[k][j],  Z[:, j] = s_mul * T_j Linv_jj   (in place).
// LT = L^T (nb x nb row-major), DinvT = the w transposed 128-block inverses stacked (w*128 x 128).  Used by the distributed
// triangular inversion: (s_upd, s_mul) = (-1, +1) solves Z L = Y, (+1, -1) gives Z = -Y L^-1.
int g3_panel_rsolve(g3_ctx* ctx, double* Y, int rows, int nb, const double* LT, const double* DinvT, double s_upd, double s_mul) {
  if (rows % TS || nb % TS || rows <= 0) return g3_fail_msg(ctx, "panel_rsolve: bad geometry");
  const int Tr = rows / TS, w = nb / TS;
  CUtensorMap tmY, tmL, tmD;
  int rc;
  if ((rc = g3_make_tmap(ctx, &tmY, Y, nb, rows, 1, nb, (uint64_t)rows * nb, G3_BM))) return rc;
  if ((rc = g3_make_tmap(ctx, &tmL, LT, nb, nb, 1, nb, (uint64_t)nb * nb, G3_BN))) return rc;
  if ((rc = g3_make_tmap(ctx, &tmD, DinvT, TS, (uint64_t)w * TS, 1, TS, (uint64_t)w * TS * TS, G3_BN))) return rc;
  for (int j = w - 1; j >= 0; --j) {
    if (j < w - 1) {
      GemmArgs g = gemm_zero();
      g.D = Y; g.ldd = nb; g.strideD = 0;
      g.mode = 0; g.ntx = Tr; g.nty = 1;
      g.d_r0 = 0; g.d_c0 = j * TS;
      g.a_r0 = 0; g.a_rx = TS; g.ka0 = (j + 1) * TS;
      g.b_r0 = j * TS; g.kb0 = (j + 1) * TS;
      g.kl0 = (w - 1 - j) * TS;
      g.alpha = s_upd; g.beta = 1.0;
      if ((rc = g3_gemm_launch(ctx, tmY, tmL, g, 1))) return rc;
    }
    GemmArgs g = gemm_zero();
    g.D = Y; g.ldd = nb; g.strideD = 0;
    g.mode = 0; g.ntx = Tr; g.nty = 1;
    g.d_r0 = 0; g.d_c0 = j * TS;
    g.a_r0 = 0; g.a_rx = TS; g.ka0 = j * TS;
    g.b_r0 = j * TS; g.kb0 = 0;
    g.kl0 = TS;
    g.alpha = s_mul; g.beta = 0.0;
    if ((rc = g3_gemm_launch(ctx, tmY, tmD, g, 1))) return rc;
  }
  return 0;
}

// U = L^-T (row-major upper).  Diagonal tiles of U are Linv_jj^T, rebuilt here from Dinv.
namespace {
__global__ void __launch_bounds__(256)
u_diag_from_dinv_kernel(const double* __restrict__ Dinv, double* __restrict__ U, int Np, int T, int j0) {
  __shared__ double tile[32][33];
  const int b = blockIdx.z, j = j0 + blockIdx.y;
  const int bx = blockIdx.x & 3, by = blockIdx.x >> 2;  // 4x4 sub-tiles of 32x32
  const double* Dj = Dinv + ((long long)b * T + j) * TS * TS;
  double* Ut = U + (long long)b * Np * Np + (long long)j * TS * Np + (long long)j * TS;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) tile[r][tx] = Dj[(by * 32 + r) * TS + bx * 32 + tx];
  __syncthreads();
  for (int r = ty; r < 32; r += 8) Ut[(long long)(bx * 32 + r) * Np + by * 32 + tx] = tile[tx][r];
}
}  // namespace

// Rows [io, ie) of L^-1 (tile columns of U): S[j][i] = sum_{k=j}^{i-1} U[j][k] L[i][k]^T, then U[j][i] = -S Linv_ii^T.
// The part of the sum over finished rows k < io is ONE launch for the whole block, the part inside the block goes row
// by row.  Needs rows < io of U complete and L final for rows < ie.
static int trtri_block(g3_ctx* ctx, const CUtensorMap& tmU, const CUtensorMap& tmL, const CUtensorMap& tmD, double* U,
                       int Np, int B, int io, int ie) {
  const long long strideU = (long long)Np * Np;
  int rc;
  if (io > 0) {  // S[j][i] = sum_{k=j}^{io-1} U[j][k] L[i][k]^T,  j < io <= i < ie
    GemmArgs g = gemm_zero();
    g.D = U; g.ldd = Np; g.strideD = strideU;
    g.mode = 0; g.ntx = io; g.nty = ie - io;
    g.d_r0 = 0; g.d_c0 = io * TS;
    g.a_r0 = 0; g.a_rx = TS;
    g.b_r0 = io * TS; g.b_ry = TS;
    g.ka0 = 0; g.ka_x = TS; g.kb0 = 0; g.kb_x = TS;
    g.kl0 = io * TS; g.kl_x = -TS;
    g.alpha = 1.0; g.beta = 0.0;
    if ((rc = g3_gemm_launch(ctx, tmU, tmL, g, B))) return rc;
  }
  for (int i = io > 0 ? io : 1; i < ie; ++i) {
    if (i > io && io > 0) {  // rows above the block: S[j][i] += sum_{k=io}^{i-1} U[j][k] L[i][k]^T,  j < io
      GemmArgs g = gemm_zero();
      g.D = U; g.ldd = Np; g.strideD = strideU;
      g.mode = 0; g.ntx = io; g.nty = 1;
      g.d_r0 = 0; g.d_c0 = i * TS;
      g.a_r0 = 0; g.a_rx = TS;
      g.b_r0 = i * TS;
      g.ka0 = io * TS; g.kb0 = io * TS;
      g.kl0 = (i - io) * TS;
      g.alpha = 1.0; g.beta = 1.0;
      if ((rc = g3_gemm_launch(ctx, tmU, tmL, g, B))) return rc;
    }
    if (i > io) {  // rows inside the block: S[j][i] = sum_{k=j}^{i-1} U[j][k] L[i][k]^T,  io <= j < i
      GemmArgs g = gemm_zero();
      g.D = U; g.ldd = Np; g.strideD = strideU;
      g.mode = 0; g.ntx = i - io; g.nty = 1;
      g.d_r0 = io * TS; g.d_c0 = i * TS;
      g.a_r0 = io * TS; g.a_rx = TS;
      g.b_r0 = i * TS;
      g.ka0 = io * TS; g.ka_x = TS; g.kb0 = io * TS; g.kb_x = TS;
      g.kl0 = (i - io) * TS; g.kl_x = -TS;
      g.alpha = 1.0; g.beta = 0.0;
      if ((rc = g3_gemm_launch(ctx, tmU, tmL, g, B))) return rc;
    }
    {  // U[j][i] = -S[j][i] Linv_ii^T,  j < i
      GemmArgs g = gemm_zero();
      g.D = U; g.ldd = Np; g.strideD = strideU;
      g.mode = 0; g.ntx = i; g.nty = 1;
      g.d_r0 = 0; g.d_c0 = i * TS;
      g.a_r0 = 0; g.a_rx = TS; g.ka0 = i * TS;
      g.b_r0 = i * TS; g.kb0 = 0;
      g.kl0 = TS;
      g.alpha = -1.0; g.beta = 0.0; g.tri_b = 1;
      if ((rc = g3_gemm_launch(ctx, tmU, tmD, g, B))) return rc;
    }
  }
  return 0;
}

static void launch_u_diag(g3_ctx* ctx, const double* Dinv, double* U, int Np, int T, int B, int j0, int nj) {
  u_diag_from_dinv_kernel<<<dim3(16, nj, B), 256, 0, ctx->stream>>>(Dinv, U, Np, T, j0);
  ctx->launches++;
}

int g3_trtri_batched(g3_ctx* ctx, const double* L, double* U, int Np, int B, const double* Dinv) {
  G3_NVTX("g3:trtri");
  const int T = Np / TS;
  int rc;
  launch_u_diag(ctx, Dinv, U, Np, T, B, 0, T);
  G3_CUDA(ctx, cudaGetLastError());
  CUtensorMap tmU, tmL, tmD;
  if ((rc = g3_make_tmap(ctx, &tmU, U, Np, Np, B, Np, (uint64_t)Np * Np, G3_BM))) return rc;
  if ((rc = g3_make_tmap(ctx, &tmL, L, Np, Np, B, Np, (uint64_t)Np * Np, G3_BN))) return rc;
  if ((rc = g3_make_tmap(ctx, &tmD, Dinv, TS, (uint64_t)T * TS, B, TS, (uint64_t)T * TS * TS, G3_BN))) return rc;
  // With a batch the rows are done one at a time (2*B*i CTAs per launch); few large matrices go by outer blocks of
  // w rows -- the same split as the right-looking potrf.
  int w = ctx->potrf_w > 0 ? ctx->potrf_w : (g3_use_blocked(ctx, B) ? g3_auto_block(ctx, T, B) : T);
  if (w < 1) w = 1;
  for (int io = 0; io < T; io += w) {
    const int ie = io + w < T ? io + w : T;
    if ((rc = trtri_block(ctx, tmU, tmL, tmD, U, Np, B, io, ie))) return rc;
  }
  return 0;
}

int g3_lauum_batched(g3_ctx* ctx, const double* U, double* Kinv, int Np, int B) {
  G3_NVTX("g3:lauum");
  const int T = Np / TS;
  int rc;
  CUtensorMap tmA, tmB;
  if ((rc = g3_make_tmap(ctx, &tmA, U, Np, Np, B, Np, (uint64_t)Np * Np, G3_BM))) return rc;
  if ((rc = g3_make_tmap(ctx, &tmB, U, Np, Np, B, Np, (uint64_t)Np * Np, G3_BN))) return rc;
  GemmArgs g = gemm_zero();
  g.D = Kinv; g.ldd = Np; g.strideD = (long long)Np * Np;
  g.mode = 1; g.ntx = T;
  g.d_r0 = 0; g.d_c0 = 0;
  g.a_r0 = 0; g.a_rx = TS;
  g.b_r0 = 0; g.b_ry = TS;
  g.ka0 = 0; g.ka_x = TS; g.kb0 = 0; g.kb_x = TS;
  g.kl0 = Np; g.kl_x = -TS;
  g.alpha = 1.0; g.beta = 0.0; g.upper = 1;
  return g3_gemm_launch(ctx, tmA, tmB, g, B);
}

static int* trsv_sync(g3_ctx* ctx, int n, double** part) {
  char name[64];
  snprintf(name, sizeof name, "trsv_sync_%p", (void*)ctx->stream);     // per stream: batch groups solve concurrently
  int* sync = (int*)g3_ws(ctx, name, sizeof(int) * (size_t)(n + 1) + sizeof(double) * (size_t)n + 16);
  if (!sync) return nullptr;
  *part = reinterpret_cast<double*>(reinterpret_cast<char*>(sync) + ((sizeof(int) * (size_t)(n + 1) + 15) / 16) * 16);
  cudaMemsetAsync(sync, 0, sizeof(int) * (size_t)(n + 1), ctx->stream);
  return sync;
}

// one launch for few right-hand sides (the chain of launches is what costs there); a big batch keeps the step kernels, whose
// (T - j) x B CTAs per step are throughput-bound (measured: 2.84 ms against 3.46 ms per step of the N = 4096 x 64 bench)
static bool trsv_use_fused(const g3_ctx* ctx, int B) { return ctx->trsv_fused && B <= 8; }

int g3_trsv_fwd(g3_ctx* ctx, const double* L, const double* Dinv, double* r, double* u, double* beta, int Np, int B) {
  const int T = Np / TS;
  if (trsv_use_fused(ctx, B)) {
    double* part = nullptr;
    int* sync = trsv_sync(ctx, T * B, &part);
    if (!sync) return -2;
    g3_prof_begin(ctx, G3_PROF_TRSV);
    trsv_fwd_fused_kernel<<<T * B, 256, 0, ctx->stream>>>(L, Dinv, r, u, beta, Np, T, sync, part);
    g3_prof_end(ctx);
    G3_LAUNCH_CHECK(ctx);
    return 0;
  }
  g3_prof_begin(ctx, G3_PROF_TRSV);
  for (int j = 0; j < T; ++j) {
    trsv_fwd_step_kernel<<<dim3(T - j, B), 256, 0, ctx->stream>>>(L, Dinv, r, u, beta, j, Np, T);
    G3_LAUNCH_CHECK(ctx);
  }
  g3_prof_end(ctx);
  return 0;
}

// Forward substitution with one factored panel (rows x nb, ld = nb): u = L_top^-1 r[0:nb] and
// r[nb:rows] -= L_below u (r is updated in place, beta += |u|^2).
int g3_trsv_panel(g3_ctx* ctx, const double* P, int rows, int nb, const double* Dinv, double* r, double* u,
                  double* beta) {
  if (rows % TS || nb % TS || rows < nb) return g3_fail_msg(ctx, "trsv_panel: bad geometry");
  const int Tr = rows / TS, w = nb / TS;
  g3_prof_begin(ctx, G3_PROF_TRSV);
  for (int j = 0; j < w; ++j) {
    trsv_fwd_step_kernel<<<dim3(Tr - j, 1), 256, 0, ctx->stream>>>(P, Dinv, r, u, beta, j, nb, w);
    G3_LAUNCH_CHECK(ctx);
  }
  g3_prof_end(ctx);
  return 0;
}

int g3_trsv_bwd(g3_ctx* ctx, const double* L, const double* Dinv, double* s, double* alpha, int Np, int B) {
  const int T = Np / TS;
  if (trsv_use_fused(ctx, B)) {
    double* part = nullptr;
    int* sync = trsv_sync(ctx, T * B, &part);
    if (!sync) return -2;
    g3_prof_begin(ctx, G3_PROF_TRSV);
    trsv_bwd_fused_kernel<<<T * B, 512, 0, ctx->stream>>>(L, Dinv, s, alpha, Np, T, sync);
    g3_prof_end(ctx);
    G3_LAUNCH_CHECK(ctx);
    return 0;
  }
  g3_prof_begin(ctx, G3_PROF_TRSV);
  for (int j = T - 1; j >= 0; --j) {
    trsv_bwd_step_kernel<<<dim3(j + 1, B), 256, 0, ctx->stream>>>(L, Dinv, s, alpha, j, Np, T);
    G3_LAUNCH_CHECK(ctx);
  }
  g3_prof_end(ctx);
  return 0;
}

int g3_set_diag_variant(g3_ctx* ctx, int variant) {
  if (variant != 1 && variant != 2) return g3_fail_msg(ctx, "g3_set_diag_variant: 1 or 2");
  if (ctx->diag_variant != variant) g3_graph_drop(ctx);
  ctx->diag_variant = variant;
  return 0;
}

// Times the diagonal-tile kernel alone (reps back-to-back launches of B CTAs on fresh SPD tiles) and checks tile 0 on the
// host: err[0] = max |L L^T - A| / max |A|, err[1] = max |Dinv L - I|, err[2] = |logdet - host logdet|, err[3] = info.
int g3_debug_diag_time(g3_ctx* ctx, int variant, int B, int reps, double cond_shift, float* us_per_launch, long long* stamps32,
                       double* err4) {
  G3_CUDA(ctx, cudaSetDevice(ctx->device));
  if (B < 1 || reps < 1) return g3_fail_msg(ctx, "diag_time: B, reps >= 1");
  const size_t tile = (size_t)TS * TS;
  std::vector<double> M(tile), Ah(tile);
  unsigned long long z = 0x1234567ull;
  for (size_t i = 0; i < tile; ++i) {
    z = z * 6364136223846793005ull + 1442695040888963407ull;
    M[i] = (double)(z >> 11) * (1.0 / 9007199254740992.0) - 0.5;
  }
  for (int r = 0; r < TS; ++r)
    for (int c = 0; c < TS; ++c) {
      double acc = 0.0;
      for (int k = 0; k < TS; ++k) acc += M[r * TS + k] * M[c * TS + k];
      Ah[r * TS + c] = acc / TS + (r == c ? fabs(cond_shift) : 0.0);
    }
  double* dA0 = (double*)g3_ws(ctx, "dt_A0", sizeof(double) * tile);
  double* dA = (double*)g3_ws(ctx, "dt_A", sizeof(double) * tile * B * reps);
  double* dD = (double*)g3_ws(ctx, "dt_D", sizeof(double) * tile * B * reps);
  double* dld = (double*)g3_ws(ctx, "dt_ld", sizeof(double) * B * reps);
  int* dinfo = (int*)g3_ws(ctx, "dt_info", sizeof(int) * B * reps);
  long long* dst = (long long*)g3_ws(ctx, "dt_stamps", sizeof(long long) * 64);
  if (!dA0 || !dA || !dD || !dld || !dinfo || !dst) return -2;
  G3_CUDA(ctx, cudaMemcpyAsync(dA0, Ah.data(), sizeof(double) * tile, cudaMemcpyHostToDevice, ctx->stream));
  for (size_t q = 0; q < (size_t)B * reps; ++q)
    G3_CUDA(ctx, cudaMemcpyAsync(dA + q * tile, dA0, sizeof(double) * tile, cudaMemcpyDeviceToDevice, ctx->stream));
  G3_CUDA(ctx, cudaMemsetAsync(dld, 0, sizeof(double) * B * reps, ctx->stream));
  G3_CUDA(ctx, cudaMemsetAsync(dinfo, 0, sizeof(int) * B * reps, ctx->stream));
  G3_CUDA(ctx, cudaMemsetAsync(dst, 0, sizeof(long long) * 64, ctx->stream));
  const int keep = ctx->diag_variant;
  ctx->diag_variant = variant;
  int rc = 0;
  if (variant == 2) rc = g3_diag2_prepare(ctx, 0, 1, dD, 1, nullptr, B * reps);
  // warm-up launch on the last repetition's tiles (re-copied afterwards)
  if (!rc) rc = g3_diag_launch(ctx, dA + (size_t)(reps - 1) * B * tile, TS, (long long)tile, 0, dD, 1, dld, dinfo, nullptr, B);
  for (int q = 0; q < B && !rc; ++q)
    cudaMemcpyAsync(dA + ((size_t)(reps - 1) * B + q) * tile, dA0, sizeof(double) * tile, cudaMemcpyDeviceToDevice, ctx->stream);
  cudaMemsetAsync(dld, 0, sizeof(double) * B * reps, ctx->stream);
  if (!rc) rc = g3_timer_begin(ctx);
  for (int it = 0; it < reps && !rc; ++it)
    rc = g3_diag_launch(ctx, dA + (size_t)it * B * tile, TS, (long long)tile, 0, dD + (size_t)it * B * tile, 1, dld + (size_t)it * B,
                        dinfo + (size_t)it * B, nullptr, B);
  float ms = 0.f;
  if (!rc) rc = g3_timer_end(ctx, &ms);
  if (!rc && variant == 2 && stamps32) {  // one more launch with the phase clocks on
    cudaMemcpyAsync(dA, dA0, sizeof(double) * tile, cudaMemcpyDeviceToDevice, ctx->stream);
    cudaMemsetAsync(dld, 0, sizeof(double), ctx->stream);
    rc = g3_diag2_launch(ctx, dA, TS, (long long)tile, 0, dD, 1, dld, dinfo, nullptr, 1, dst, cond_shift < 0 ? 1 : 0);
  }
  if (!rc && variant == 2) rc = g3_diag2_finish(ctx, dA, TS, (long long)tile, 0, 1, nullptr, 1);
  ctx->diag_variant = keep;
  if (rc) return rc;
  std::vector<double> Lh(tile), Xh(tile);
  double ld = 0.0;
  int inf = 0;
  G3_CUDA(ctx, cudaMemcpyAsync(Lh.data(), dA, sizeof(double) * tile, cudaMemcpyDeviceToHost, ctx->stream));
  G3_CUDA(ctx, cudaMemcpyAsync(Xh.data(), dD, sizeof(double) * tile, cudaMemcpyDeviceToHost, ctx->stream));
  G3_CUDA(ctx, cudaMemcpyAsync(&ld, dld, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  G3_CUDA(ctx, cudaMemcpyAsync(&inf, dinfo, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  if (stamps32) G3_CUDA(ctx, cudaMemcpyAsync(stamps32, dst, sizeof(long long) * 64, cudaMemcpyDeviceToHost, ctx->stream));
  G3_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (us_per_launch) *us_per_launch = ms * 1000.f / reps;
  if (err4) {
    double amax = 0.0, e0 = 0.0, e1 = 0.0, hld = 0.0;
    for (size_t i = 0; i < tile; ++i) amax = fmax(amax, fabs(Ah[i]));
    for (int r = 0; r < TS; ++r)
      for (int c = 0; c < TS; ++c) {
        double acc = 0.0, acc2 = 0.0;
        for (int k = 0; k < TS; ++k) {
          acc += Lh[r * TS + k] * Lh[c * TS + k];
          acc2 += Xh[r * TS + k] * Lh[k * TS + c];
        }
        const double d0 = fabs(acc - Ah[r * TS + c]) / amax, d1 = fabs(acc2 - (r == c ? 1.0 : 0.0));
        e0 = (d0 > e0 || d0 != d0) ? d0 : e0;
        e1 = (d1 > e1 || d1 != d1) ? d1 : e1;
      }
    for (int k = 0; k < TS; ++k) hld += log(Lh[k * TS + k]);
    err4[0] = e0; err4[1] = e1; err4[2] = fabs(hld - ld); err4[3] = (double)inf;
  }
  return 0;
}

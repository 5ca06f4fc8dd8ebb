// Low-latency factor + inverse of one 128x128 diagonal tile by one CTA (sm_100a), second generation.
//
// This kernel sits at the head of every tile column of the blocked Cholesky (potrf.cu) that replaces
// CholeskyRobust.perform -> scipy dpotrf (g3py/libs/tensors.py:198): L_jj = chol(A_jj) and Dinv_j = L_jj^-1, which turn
// the triangular solves of the column into GEMMs.  For one or a few matrices (one NUTS / BFGS evaluation, the panel chain
// of the multi-GPU factorisation) it IS the critical path: T = N/128 of them run strictly one after the other, so what
// counts is its latency, not its throughput.  The first kernel (potrf_diag_kernel in potrf.cu, 60 us) is bulk-synchronous
// with dot-product (left-looking) inner solves and DFMA register tiles fed by broadcast shared loads.  Measured on B200
// (tools/lat_bench.cu): DFMA 8 cycles dependent / 2.16 per warp-instruction, LDS 29, STS->LDS 35, MUFU.RCP64H 18,
// DMMA.8x8x4 26 dependent / 4 per SM, and every LDS.64 costs 256 bytes of the 128 B/clk return path whether or not its
// addresses coincide -- so register-tile DFMA products out of shared memory are load-bound, not FMA-bound.  Hence:
//
//   * 512 threads, 32x32 sub-blocks, left-looking at the sub-block level:  U (update block column kb with the columns to
//     its left) -> F (warp 0 factors the diagonal sub-block) -> S (its inverse, then the rows below it).
//   * every product (U, the rows below a diagonal sub-block times its inverse, the blocks of X = L^-1) runs on the fp64
//     tensor pipe: mma.m8n8k4 fragments straight from shared memory (one 8-byte load per lane per fragment).  Row stride
//     131 doubles (= 3 mod 16): row walks are conflict-free and a fragment load takes 3 wavefronts instead of 2.
//   * F is a register-resident right-looking elimination, one lane per row, WITHOUT normalising the pivot column first:
//     a[i][m] -= (v_i r_k) v_m with v = the raw column, r_k = 1/d_k.  The broadcast values v_m do not depend on r_k, so
//     they are in flight while the reciprocal is refined, and every lane forms the NEXT pivot d_{k+1} = a_{k+1,k+1} -
//     v_{k+1}^2 r_k redundantly: the loop-carried chain is one DFMA + MUFU.RCP64H + two Newton steps.  rsqrt (for the
//     stored L = v rs_k) is off that chain.  No branch, call or generic address inside the loop (inline PTX on 32-bit
//     shared addresses), so each step is one basic block for ptxas to schedule.
//   * the 32x32 triangular inverse is a right-looking forward substitution on the identity (lane = one column), the
//     sub-block's transpose LT broadcast with 128-bit loads: chain per step DMUL + DFMA with 31-k independent DFMAs behind it.
//   * the off-diagonal blocks of X (X[I][J] = -X_II sum_{K=J}^{I-1} L[I][K] X[K][J]) are done by the 15 other warps WHILE
//     warp 0 factors the next diagonal sub-block; the sums T[I][J] accumulate in place, transposed, in the unused upper
//     blocks of the tile.
//
// Shared memory: S[128][131] (L below the diagonal, X^T above) | XDT[4][32][35] (inverses of the diagonal sub-blocks,
// transposed) | LT[32][32] | column / pivot mailboxes | 1/diag.
#include "g3b_internal.cuh"
#include <math.h>

namespace {

constexpr int TS = G3_TILE;      // 128
constexpr int NT = 512;          // threads
constexpr int LD = 131;          // row stride of S in doubles
constexpr int XS = 35;           // row stride of the 32x32 side blocks
constexpr int OFF_XDT = TS * LD;                 // XDT[b][c][r] = Xd_b[r][c]
constexpr int VS = 36;                           // row stride of VT / TT / TQ (= 4 mod 16: fragment loads in 2 wavefronts)
constexpr int OFF_TQ = OFF_XDT + 4 * 32 * XS;    // TQ[16][36]: scratch of the two-level inverse
constexpr int OFF_RV = OFF_TQ + 16 * VS;         // 1 / d_k of the current diagonal sub-block, 32
constexpr int OFF_VT = OFF_RV + 32;              // VT[k][m] = raw pivot column k of the current sub-block (v_m, m >= k): F's mailbox and history
constexpr int OFF_TT = OFF_VT + 32 * VS;         // TT[k][i] = -v_i / d_k for k < 16: multipliers of the first half
constexpr int OFF_DN = OFF_TT + 16 * VS;         // two "next column, one update behind" mailboxes of 32
constexpr int OFF_INVD = OFF_DN + 64;            // 1 / L[k][k], 128
constexpr int OFF_END = OFF_INVD + TS;
static_assert(OFF_VT % 2 == 0 && OFF_INVD % 2 == 0, "128-bit shared loads need even offsets");
constexpr uint32_t B8 = 8;       // bytes per double

// ---- shared memory through 32-bit shared-space addresses ------------------------------------------------------------
__device__ __forceinline__ double lds(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
  return v;
}
template <int OFF>
__device__ __forceinline__ double ldso(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(addr), "n"(OFF) : "memory");
  return v;
}
template <int OFF>
__device__ __forceinline__ void ldso2(uint32_t addr, double& x, double& y) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(x), "=d"(y) : "r"(addr), "n"(OFF) : "memory");
}
__device__ __forceinline__ void sts(uint32_t addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }
template <int OFF>
__device__ __forceinline__ void stso(uint32_t addr, double v) {
  asm volatile("st.shared.f64 [%0+%1], %2;" ::"r"(addr), "n"(OFF), "d"(v) : "memory");
}
__device__ __forceinline__ double rcp_seed(double d) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));      // MUFU.RCP64H, ~2^-20
  return y;
}
__device__ __forceinline__ double rsqrt_seed(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));    // MUFU.RSQ64H, ~2^-20
  return y;
}
// two Newton steps from the seed z ~ d^-1/2 (no special-case path: a non-positive pivot gives NaN, which is what a failed
// factorisation should leave behind)
__device__ __forceinline__ double rsqrt_refine(double d, double z) {
  double h = 0.5 * z, e = fma(-(d * z), z, 1.0);
  z = fma(h, e, z);
  h = 0.5 * z;
  e = fma(-(d * z), z, 1.0);
  return fma(h, e, z);
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ---- F: Cholesky of a 32x32 diagonal sub-block by one warp ------------------------------------------------------------
// a[m] -= t * p[m] for m in [M, N), p broadcast from byte address `pb` (16-byte aligned at even m)
template <int M, int N>
struct AxpyTail {
  static __device__ __forceinline__ void run(double (&a)[N], double t, uint32_t pb) {
    if constexpr ((M & 1) != 0) {
      a[M] = fma(-t, ldso<M * 8>(pb), a[M]);
      AxpyTail<M + 1, N>::run(a, t, pb);
    } else {
      double x, y;
      ldso2<M * 8>(pb, x, y);
      a[M] = fma(-t, x, a[M]);
      a[M + 1] = fma(-t, y, a[M + 1]);
      AxpyTail<M + 2, N>::run(a, t, pb);
    }
  }
};
template <int N>
struct AxpyTail<N, N> {
  static __device__ __forceinline__ void run(double (&)[N], double, uint32_t) {}
};

// One half (16 pivots) of the elimination.  H = 0: pivots 0..15, lane = row 0..31 of the sub-block (rows 16..31 ride along as a
// panel; their multipliers go to TT for the tensor-pipe update of the trailing 16x16).  H = 1: pivots 16..31, lane & 15 = row.
// Entering step K: a[K..15] = row `il` after updates 0..K-1; d = pivot, y ~ 1/d (seed); VT[K][i] = a_i[K]; DN[K&1][i] = a_i[K+1]
// before update K (so that every lane can form the next pivot the moment 1/d is known).
template <int K, int H>
struct FStep {
  static __device__ __forceinline__ void run(double (&a)[16], uint32_t vtb, uint32_t vtl, uint32_t ttl, uint32_t dnb, uint32_t dnl,
                                             int il, double d, double y, double& dmine, int& bad) {
    constexpr uint32_t PAR = (K & 1) * 256, NPAR = ((K + 1) & 1) * 256;
    // the two mailbox values the next pivot needs come first: everything shared below is ordered behind them
    double w1 = 0.0, ad = 0.0;
    if constexpr (K < 15) {
      w1 = ldso<(K * VS + K + 1) * 8>(vtb);
      ad = ldso<(K + 1) * 8 + PAR>(dnb);
    }
    // reciprocal: two Newton steps; the multiplier t = a[K] / d takes the second one as a correction
    const double e1 = fma(-d, y, 1.0);
    const double y1 = fma(y, e1, y);
    const double e2 = fma(-d, y1, 1.0);
    double t = a[K] * y1;
    t = fma(t, e2, t);
    dmine = (il == K) ? d : dmine;
    bad = (bad < 0 && !(d > 0.0)) ? (16 * H + K) : bad;
    double dn = 1.0, yn = 1.0;
    if constexpr (K < 15) {
      const double r = fma(y1, e2, y1);
      dn = fma(-(w1 * w1), r, ad);
      yn = rcp_seed(dn);
      AxpyTail<K + 1, 16>::run(a, t, vtb + K * VS * 8);
      stso<(K + 1) * VS * 8>(vtl, a[(K + 1) & 15]);
      stso<NPAR>(dnl, a[(K + 2) & 15]);
    }
    if constexpr (H == 0) stso<K * VS * 8>(ttl, -t);
    __syncwarp();
    FStep<K + 1, H>::run(a, vtb, vtl, ttl, dnb, dnl, il, dn, yn, dmine, bad);
  }
};
template <int H>
struct FStep<16, H> {
  static __device__ __forceinline__ void run(double (&)[16], uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, int, double, double,
                                             double&, int&) {}
};

template <int M, int N>
struct LoadRow {
  static __device__ __forceinline__ void run(double (&a)[N], uint32_t row, int il) {
    const double v = ldso<M * 8>(row);
    a[M] = (M <= il) ? v : 0.0;
    LoadRow<M + 1, N>::run(a, row, il);
  }
};
template <int N>
struct LoadRow<N, N> {
  static __device__ __forceinline__ void run(double (&)[N], uint32_t, int) {}
};

// Warp 0: Cholesky of the 32x32 diagonal sub-block at offset o, left as its raw pivot columns VT[k][m] = L[m][k] sqrt(d_k) with
// 1/sqrt(d_k) in invd[o + k] and 1/d_k in RV[k] (the scaled block is written into S by other warps afterwards).
// Returns the first failed pivot (or -1).  sb = shared byte address of S[0][0].
__device__ __forceinline__ int factor32(uint32_t sb, int o, int lane, long long* st) {
  const uint32_t vt = sb + OFF_VT * B8, tt = sb + OFF_TT * B8, dnb = sb + OFF_DN * B8;
  const uint32_t invd = sb + (OFF_INVD + o) * B8;
  const uint32_t row = sb + ((o + lane) * LD + o) * B8;
  int bad = -1;
  double dA = 1.0, dC = 1.0;
  // ---- first half: pivots 0..15, 32 rows ----
  double a[16];
  LoadRow<0, 16>::run(a, row, lane);
  stso<0>(vt + lane * B8, a[0]);
  stso<0>(dnb + lane * B8, a[1]);
  __syncwarp();
  {
    const double d = ldso<0>(vt);
    FStep<0, 0>::run(a, vt, vt + lane * B8, tt + lane * B8, dnb, dnb + lane * B8, lane, d, rcp_seed(d), dA, bad);
  }
  if (st) st[0] = clock64();
  // ---- trailing 16x16 (rows / columns 16..31 of the sub-block) += TT^T VT on the tensor pipe: tiles (0,0), (1,0), (1,1) ----
  {
    const int fr = lane >> 2, fk = lane & 3;
    const uint32_t cb = sb + ((o + 16 + fr) * LD + o + 16 + 2 * fk) * B8;
    double c00[2], c10[2], c11[2];
    c00[0] = lds(cb); c00[1] = lds(cb + 8);
    c10[0] = lds(cb + 8 * LD * B8); c10[1] = lds(cb + 8 * LD * B8 + 8);
    c11[0] = lds(cb + (8 * LD + 8) * B8); c11[1] = lds(cb + (8 * LD + 8) * B8 + 8);
    const uint32_t ap = tt + (fk * VS + 16 + fr) * B8, bp = vt + (fk * VS + 16 + fr) * B8;
#pragma unroll
    for (int k0 = 0; k0 < 16; k0 += 4) {
      const double a0 = lds(ap + k0 * VS * B8), a1 = lds(ap + (k0 * VS + 8) * B8);
      const double b0 = lds(bp + k0 * VS * B8), b1 = lds(bp + (k0 * VS + 8) * B8);
      dmma(c00[0], c00[1], a0, b0);
      dmma(c10[0], c10[1], a1, b0);
      dmma(c11[0], c11[1], a1, b1);
    }
    sts(cb, c00[0]); sts(cb + 8, c00[1]);
    sts(cb + 8 * LD * B8, c10[0]); sts(cb + 8 * LD * B8 + 8, c10[1]);
    sts(cb + (8 * LD + 8) * B8, c11[0]); sts(cb + (8 * LD + 8) * B8 + 8, c11[1]);
    __syncwarp();
  }
  if (st) st[1] = clock64();
  // ---- second half: pivots 16..31, rows 16..31 (lanes 16..31 mirror lanes 0..15) ----
  const int il = lane & 15;
  double c[16];
  LoadRow<0, 16>::run(c, sb + ((o + 16 + il) * LD + o + 16) * B8, il);
  {
    const uint32_t vtb2 = vt + (16 * VS + 16) * B8, vtl2 = vtb2 + il * B8;
    stso<0>(vtl2, c[0]);
    stso<0>(dnb + il * B8, c[1]);
    __syncwarp();
    const double d = ldso<0>(vtb2);
    FStep<0, 1>::run(c, vtb2, vtl2, 0u, dnb, dnb + il * B8, il, d, rcp_seed(d), dC, bad);
  }
  if (st) st[2] = clock64();
  // ---- 1/sqrt and 1/ of the pivots, one per lane ----
  const double dK = lane < 16 ? dA : dC;
  sts(invd + lane * B8, rsqrt_refine(dK, rsqrt_seed(dK)));
  {
    const double y0 = rcp_seed(dK), e1 = fma(-dK, y0, 1.0), y1 = fma(y0, e1, y0), e2 = fma(-dK, y1, 1.0);
    sts(sb + (OFF_RV + lane) * B8, fma(y1, e2, y1));
  }
  return bad;
}

// ---- inverse of the 32x32 diagonal sub-block by one warp, two levels -----------------------------------------------------
// Level 1: lanes 0..15 solve L11 x = e_c, lanes 16..31 solve L22 x = e_c (16 right-looking steps on the raw columns:
// y = b_k / d_k is the chain value, x_k = b_k / sqrt(d_k) the result, b_m -= y VT[k][m]).  x goes to XDT, y (= X11 with its
// rows scaled by 1/sqrt(d), i.e. what the raw columns of L21 must be multiplied with) to TT.
template <int K>
struct HalfSolve {
  // results stay in registers (xs, ys) until the end: a shared store inside the loop would order the next step's loads behind it
  static __device__ __forceinline__ void run(double (&b)[16], double (&xs)[16], double (&ys)[16], uint32_t vth, const double (&rv)[16],
                                             const double (&iv)[16]) {
    const double y = b[K] * rv[K];
    ys[K] = y;
    xs[K] = b[K] * iv[K];
    AxpyTail<K + 1, 16>::run(b, y, vth + K * VS * 8);
    HalfSolve<K + 1>::run(b, xs, ys, vth, rv, iv);
  }
};
template <>
struct HalfSolve<16> {
  static __device__ __forceinline__ void run(double (&)[16], double (&)[16], double (&)[16], uint32_t, const double (&)[16],
                                             const double (&)[16]) {}
};
// Level 2 on the tensor pipe: T = L21 X11 = V21 Z11 (TQ), then X21 = -X22 T.
__device__ __forceinline__ void inverse32(uint32_t sb, int kb, int o, int lane) {
  const uint32_t vt = sb + OFF_VT * B8, zs = sb + OFF_TT * B8, tq = sb + OFF_TQ * B8;
  const uint32_t xd = sb + (OFF_XDT + kb * 32 * XS) * B8;
  {
    const int h = lane >> 4, c = lane & 15;
    double bcol[16], xs[16], ys[16], rv[16], iv[16];
    const uint32_t rvp = sb + (OFF_RV + 16 * h) * B8, ivp = sb + (OFF_INVD + o + 16 * h) * B8;
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      bcol[m] = (m == c) ? 1.0 : 0.0;
      rv[m] = lds(rvp + m * B8);
      iv[m] = lds(ivp + m * B8);
    }
    HalfSolve<0>::run(bcol, xs, ys, vt + (16 * h * VS + 16 * h) * B8, rv, iv);
    const uint32_t zl = zs + lane * B8, xl = xd + ((16 * h + c) * XS + 16 * h) * B8;
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      sts(xl + m * B8, xs[m]);
      sts(zl + m * VS * B8, ys[m]);
    }
  }
  __syncwarp();
  const int fr = lane >> 2, fk = lane & 3;
  double t[2][2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) t[i][jj][0] = t[i][jj][1] = 0.0;
#pragma unroll
  for (int k0 = 0; k0 < 16; k0 += 4) {          // T(r, c) = sum_k VT[k][16 + r] Z[k][c];  Z[k][c] = 0 for k < c
    const double a0 = lds(vt + ((k0 + fk) * VS + 16 + fr) * B8), a1 = lds(vt + ((k0 + fk) * VS + 24 + fr) * B8);
    const double b0 = lds(zs + ((k0 + fk) * VS + fr) * B8);
    dmma(t[0][0][0], t[0][0][1], a0, b0);
    dmma(t[1][0][0], t[1][0][1], a1, b0);
    if (k0 >= 8) {
      const double b1 = lds(zs + ((k0 + fk) * VS + 8 + fr) * B8);
      dmma(t[0][1][0], t[0][1][1], a0, b1);
      dmma(t[1][1][0], t[1][1][1], a1, b1);
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const uint32_t tp = tq + ((8 * i + fr) * VS + 8 * jj + 2 * fk) * B8;      // TQ[m][c]
      sts(tp, t[i][jj][0]);
      sts(tp + 8, t[i][jj][1]);
    }
  __syncwarp();
  double x[2][2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) x[i][jj][0] = x[i][jj][1] = 0.0;
#pragma unroll
  for (int k0 = 0; k0 < 16; k0 += 4) {          // X21(r, c) = -sum_m X22[r][m] T[m][c];  X22[r][m] = XDT[16 + m][16 + r] = 0 for m > r
    const double b0 = lds(tq + ((k0 + fk) * VS + fr) * B8), b1 = lds(tq + ((k0 + fk) * VS + 8 + fr) * B8);
    const double a1 = lds(xd + ((16 + k0 + fk) * XS + 24 + fr) * B8);
    dmma(x[1][0][0], x[1][0][1], a1, b0);
    dmma(x[1][1][0], x[1][1][1], a1, b1);
    if (k0 < 8) {
      const double a0 = lds(xd + ((16 + k0 + fk) * XS + 16 + fr) * B8);
      dmma(x[0][0][0], x[0][0][1], a0, b0);
      dmma(x[0][1][0], x[0][1][1], a0, b1);
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const uint32_t xp = xd + ((8 * jj + 2 * fk) * XS + 16 + 8 * i + fr) * B8;    // XDT[c][16 + r]
      sts(xp, -x[i][jj][0]);
      sts(xp + XS * B8, -x[i][jj][1]);
    }
}

// ---- products on the fp64 tensor pipe -----------------------------------------------------------------------------------
// acc[i][j] += sum_{k in [K0, K1)} A(8i + fr, k) B(8j + fr', k) with X(r, k) at byte address X + (r*RS + k*KS)*8; lane holds
// A(8i + lane/4, k + lane%4), B(8j + lane/4, k + lane%4) and C(8i + lane/4, 8j + 2(lane%4) + {0,1}).
// TRI = 1: B(c, k) = 0 for k > c (skip the k-blocks above the diagonal);  TRI = 2: B(c, k) = 0 for k < c;  TRI = 3: A(r, k) = 0 for k > r.
template <int MT, int NTL, int ARS, int AKS, int BRS, int BKS, int TRI, int KN>
__device__ __forceinline__ void mma_acc(double (&acc)[MT][NTL][2], uint32_t A, uint32_t Bm, int lane) {
  const int fr = lane >> 2, fk = lane & 3;
  const uint32_t ap = A + (fr * ARS + fk * AKS) * B8, bp = Bm + (fr * BRS + fk * BKS) * B8;
#pragma unroll
  for (int k = 0; k < KN; k += 4) {
    {
      double af[MT], bf[NTL];
#pragma unroll
      for (int i = 0; i < MT; ++i) af[i] = lds(ap + (i * 8 * ARS + k * AKS) * B8);
#pragma unroll
      for (int j = 0; j < NTL; ++j) bf[j] = lds(bp + (j * 8 * BRS + k * BKS) * B8);
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NTL; ++j) {
          const bool live = TRI == 1 ? (k < 8 * (j + 1)) : TRI == 2 ? (k + 4 > 8 * j) : TRI == 3 ? (k < 8 * (i + 1)) : true;
          if (live) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
  }
}
template <int MT, int NTL>
__device__ __forceinline__ void zero_acc(double (&acc)[MT][NTL][2]) {
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NTL; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
}

// U: S[I-block rows r0.., kb-block cols c0..] -= sum_{k < kb*32} S[I rows][k] S[kb rows][k]     (tile 8 x 8*NTL)
template <int NTL, int KB>
__device__ __forceinline__ void update_item(uint32_t sb, int I, int r0, int c0, int lane) {
  constexpr int kb = KB;
  double acc[1][NTL][2];
  zero_acc(acc);
  mma_acc<1, NTL, LD, 1, LD, 1, 0, KB * 32>(acc, sb + ((I * 32 + r0) * LD) * B8, sb + ((kb * 32 + c0) * LD) * B8, lane);
  const uint32_t cp = sb + ((I * 32 + r0 + (lane >> 2)) * LD + kb * 32 + c0 + 2 * (lane & 3)) * B8;
#pragma unroll
  for (int j = 0; j < NTL; ++j) {
    sts(cp + j * 64, lds(cp + j * 64) - acc[0][j][0]);
    sts(cp + j * 64 + 8, lds(cp + j * 64 + 8) - acc[0][j][1]);
  }
}

// Rows r0..r0+8 below the diagonal sub-block kb:  L[r][kb cols] = A[r][kb cols] Xd_kb^T, in place (the warp owns whole rows)
__device__ __forceinline__ void below_item(uint32_t sb, int kb, int row0, int lane) {
  double acc[1][4][2];
  zero_acc(acc);
  const uint32_t ap = sb + (row0 * LD + kb * 32) * B8;
  mma_acc<1, 4, LD, 1, 1, XS, 1, 32>(acc, ap, sb + (OFF_XDT + kb * 32 * XS) * B8, lane);     // B(c, k) = Xd[c][k] = XDT[k][c]
  const uint32_t cp = ap + ((lane >> 2) * LD + 2 * (lane & 3)) * B8;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    sts(cp + j * 64, acc[0][j][0]);
    sts(cp + j * 64 + 8, acc[0][j][1]);
  }
}

// T[I][J] (stored transposed in the upper block (J, I) of S: T[r][c] at S[(J*32+c)*LD + I*32 + r]), rows r0..r0+8, all 32 columns:
//   first = 1:  T  = L[I][J] Xd_J                       (Xd_J[k][c] = XDT[J][c][k], zero for k < c)
//   first = 0:  T += L[I][K] X[K][J]                    (X[K][J][k][c] = S[(J*32+c)*LD + K*32 + k])
__device__ __forceinline__ void t_item(uint32_t sb, int I, int J, int K, int first, int r0, int lane) {
  double acc[1][4][2];
  zero_acc(acc);
  const uint32_t ap = sb + ((I * 32 + r0) * LD + K * 32) * B8;
  if (first)
    mma_acc<1, 4, LD, 1, XS, 1, 2, 32>(acc, ap, sb + (OFF_XDT + J * 32 * XS) * B8, lane);
  else
    mma_acc<1, 4, LD, 1, LD, 1, 0, 32>(acc, ap, sb + ((J * 32) * LD + K * 32) * B8, lane);
  const uint32_t tp = sb + ((J * 32 + 2 * (lane & 3)) * LD + I * 32 + r0 + (lane >> 2)) * B8;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t p0 = tp + j * 8 * LD * B8, p1 = p0 + LD * B8;
    if (first) {
      sts(p0, acc[0][j][0]);
      sts(p1, acc[0][j][1]);
    } else {
      sts(p0, lds(p0) + acc[0][j][0]);
      sts(p1, lds(p1) + acc[0][j][1]);
    }
  }
}

// X[I][J] = -Xd_I T[I][J], in place over T (transposed storage), columns c0..c0+8, all 32 rows by one warp (it has read its
// whole columns of T before the last tensor instruction completes, so the in-place store is safe).
__device__ __forceinline__ void x_item(uint32_t sb, int I, int J, int c0, int lane) {
  double acc[4][1][2];
  zero_acc(acc);
  // A(r, m) = Xd_I[r][m] = XDT[I][m][r] (zero for m > r);  B(c, m) = T[m][c] = S[(J*32 + c)*LD + I*32 + m]
  mma_acc<4, 1, 1, XS, LD, 1, 3, 32>(acc, sb + (OFF_XDT + I * 32 * XS) * B8, sb + ((J * 32 + c0) * LD + I * 32) * B8, lane);
  const uint32_t xp = sb + ((J * 32 + c0 + 2 * (lane & 3)) * LD + I * 32 + (lane >> 2)) * B8;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    sts(xp + i * 64, -acc[i][0][0]);
    sts(xp + i * 64 + LD * B8, -acc[i][0][1]);
  }
}

__device__ __forceinline__ void filler_barrier() { asm volatile("bar.sync 1, 480;" ::: "memory"); }

// Global stores of finished 8 x 32 slabs (rows I*32 + s*8.., columns J*32..) as 16-byte pairs; only pairs at or below the
// diagonal (the zeros above it are written once at the start).  kind 0: L from S;  kind 1: X = L^-1 (diagonal sub-blocks from
// XDT, the others transposed from the upper blocks of S).
__device__ __forceinline__ void store_slab(uint32_t sb, double* __restrict__ At, int Np, double* __restrict__ Dj, int kind, int I,
                                           int J, int s, int lane) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int p = lane + 32 * q;
    const int r = I * 32 + s * 8 + (p >> 4), c = J * 32 + 2 * (p & 15);
    if (c <= r) {
      double2 v;
      if (kind == 0) {
        const uint32_t lp = sb + (r * LD + c) * B8;
        v.x = lds(lp);
        v.y = (c + 1 <= r) ? lds(lp + 8) : 0.0;
        *reinterpret_cast<double2*>(At + (long long)r * Np + c) = v;
      } else {
        if (I == J) {
          const uint32_t xp = sb + (OFF_XDT + (I * 32 + (c & 31)) * XS + (r & 31)) * B8;
          v.x = lds(xp);
          v.y = (c + 1 <= r) ? lds(xp + XS * B8) : 0.0;
        } else {
          const uint32_t xp = sb + (c * LD + r) * B8;
          v.x = lds(xp);
          v.y = lds(xp + LD * B8);
        }
        *reinterpret_cast<double2*>(Dj + r * TS + c) = v;
      }
    }
  }
}
// task t of a list of blocks: block t / 4, slab t % 4; a block is packed as kind | I << 1 | J << 3 in 5 bits of `code`
__host__ __device__ constexpr unsigned blk(int kind, int I, int J) { return (unsigned)(kind | (I << 1) | (J << 3)); }
__host__ __device__ constexpr unsigned long long blocks(unsigned b0, unsigned b1 = 0, unsigned b2 = 0, unsigned b3 = 0, unsigned b4 = 0) {
  return (unsigned long long)b0 | ((unsigned long long)b1 << 5) | ((unsigned long long)b2 << 10) | ((unsigned long long)b3 << 15) |
         ((unsigned long long)b4 << 20);
}
__device__ __forceinline__ void store_tasks(uint32_t sb, double* __restrict__ At, int Np, double* __restrict__ Dj, int nblk,
                                            unsigned long long code, int w0, int nw, int lane) {
  if (w0 < 0) return;
  for (int t = w0; t < 4 * nblk; t += nw) {
    const unsigned e = (unsigned)(code >> (5 * (t >> 2))) & 31u;
    store_slab(sb, At, Np, Dj, e & 1, (e >> 1) & 3, (e >> 3) & 3, t & 3, lane);
  }
}

__global__ void __launch_bounds__(NT, 1)
potrf_diag2_kernel(double* __restrict__ A, int Np, long long strideA, int j, double* __restrict__ Dinv, int T,
                   double* __restrict__ logdet, int* __restrict__ info, const int* __restrict__ bmap,
                   long long* __restrict__ stamps, int dbg) {
  extern __shared__ __align__(16) double S[];
  __shared__ int first_bad;
  const uint32_t sb = (uint32_t)__cvta_generic_to_shared(S);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = bmap ? bmap[blockIdx.x] : (int)blockIdx.x;
  double* At = A + (long long)b * strideA + (long long)j * TS * Np + (long long)j * TS;
  double* Dj = Dinv + ((long long)b * T + j) * TS * TS;
  int nstamp = 0;
#define G3_STAMP()                                                     \
  do {                                                                 \
    if (stamps && tid == 0 && blockIdx.x == 0) {                       \
      const int dmy = *(volatile int*)&first_bad; /* a load behind the barrier: the clock is read after its release */ \
      if (dmy >= -1) stamps[nstamp] = clock64();                       \
    }                                                                  \
    ++nstamp;                                                          \
  } while (0)
  if (tid == 0) first_bad = -1;
  G3_STAMP();
  // zeros: the quadrant of each sub-block inverse that is never written (rows 0..15, columns 16..31)
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int i = tid + NT * q, blkq = i >> 8, c = 16 + ((i >> 4) & 15), r = i & 15;
    sts(sb + (OFF_XDT + (blkq * 32 + c) * XS + r) * B8, 0.0);
  }
  // ---- the first diagonal sub-block only (the rest of the tile is loaded while warp 0 factors it) ----
  {
    const int r = tid >> 4, c = (tid & 15) * 2;
    if (c <= r) {
      const double2 v = __ldg(reinterpret_cast<const double2*>(At + (long long)r * Np + c));
      const uint32_t sp = sb + (r * LD + c) * B8;
      sts(sp, v.x);
      sts(sp + 8, v.y);
    }
  }
  __syncthreads();
  G3_STAMP();

  // the 12 warps that do not share warp 0's scheduler work in the shadow of F; warps 4, 8, 12 stay idle there (anything
  // issued on warp 0's sub-partition slows the serial factorisation, which is issue-bound)
  const int f = (warp & 3) ? (warp >> 2) * 3 + (warp & 3) - 1 : -1;
  const int wi = f;            // index among those 12 warps (-1: not a shadow worker)
  double ldacc = 0.0;          // warp 15: sum of log L[k][k]
  for (int kb = 0; kb < 4; ++kb) {
    const int o = kb * 32;
    // ---- U: block column kb -= (columns to the left)^2, all warps ----
    if (kb == 1) {
      if (warp < 12) update_item<4, 1>(sb, 1 + (warp >> 2), (warp & 3) * 8, 0, lane);
    } else if (kb == 2) {
      update_item<2, 2>(sb, 2 + (warp >> 3), ((warp & 7) >> 1) * 8, (warp & 1) * 16, lane);
    } else if (kb == 3) {
      update_item<1, 3>(sb, 3, (warp >> 2) * 8, (warp & 3) * 8, lane);
    }
    if (kb > 0) __syncthreads();
    G3_STAMP();
    // ---- F: warp 0 factors the diagonal sub-block; the others load / store / work on the inverse ----
    if (warp == 0) {
      const int bad = factor32(sb, o, lane, (stamps && tid == 0 && blockIdx.x == 0) ? stamps + 32 + kb * 4 : nullptr);
      if (lane == 0 && bad >= 0 && first_bad < 0) first_bad = o + bad;
      G3_STAMP();
      --nstamp;
    } else if (dbg & 1) {
      if (kb == 2 || kb == 3) filler_barrier();   // timing experiment: F alone, nothing in its shadow (results are incomplete)
    } else if (kb == 0 && wi >= 0) {
      // rows 32..127 of the lower triangle (pairs of columns), 16 loads in flight per thread of the 12 shadow warps
      const int q0 = wi * 32 + lane;
      double2 v[16];
#pragma unroll
      for (int it = 0; it < 16; ++it) {
        const int p = q0 + 384 * it, r = 32 + (p >> 6), c = (p & 63) * 2;
        v[it] = make_double2(0.0, 0.0);
        if (p < 96 * 64 && c <= r) v[it] = __ldg(reinterpret_cast<const double2*>(At + (long long)r * Np + c));
      }
#pragma unroll
      for (int it = 0; it < 16; ++it) {
        const int p = q0 + 384 * it, r = 32 + (p >> 6), c = (p & 63) * 2;
        if (p < 96 * 64 && c <= r) {
          const uint32_t sp = sb + (r * LD + c) * B8;
          sts(sp, v[it].x);
          sts(sp + 8, v[it].y);
        }
      }
    } else if (kb == 0) {
      // warps 4, 8, 12: nothing
    } else if (kb == 1) {
      // T[I][0] = L[I][0] Xd_0, I = 1..3 (12 items of 8 x 32)
      if (f >= 0) t_item(sb, 1 + (f >> 2), 0, 0, 1, (f & 3) * 8, lane);
      store_tasks(sb, At, Np, Dj, 5, blocks(blk(0, 0, 0), blk(0, 1, 0), blk(0, 2, 0), blk(0, 3, 0), blk(1, 0, 0)), wi, 12, lane);
    } else if (kb == 2) {
      // X[1][0] in place (4 items of 32 x 8)  ||  T[I][1] = L[I][1] Xd_1, I = 2, 3 (8 items)
      if (f >= 0 && f < 4) x_item(sb, 1, 0, f * 8, lane);
      else if (f >= 4) t_item(sb, 2 + ((f - 4) >> 2), 1, 1, 1, ((f - 4) & 3) * 8, lane);
      store_tasks(sb, At, Np, Dj, 4, blocks(blk(0, 1, 1), blk(0, 2, 1), blk(0, 3, 1), blk(1, 1, 1)), wi, 12, lane);
      filler_barrier();
      // T[I][0] += L[I][1] X[1][0], I = 2, 3
      if (f >= 0 && f < 8) t_item(sb, 2 + (f >> 2), 0, 1, 0, (f & 3) * 8, lane);
      store_tasks(sb, At, Np, Dj, 1, blocks(blk(1, 1, 0)), wi, 12, lane);
    } else {
      // X[2][0], X[2][1] in place (8 items)  ||  T[3][2] = L[3][2] Xd_2 (4 items)
      if (f >= 0 && f < 8) x_item(sb, 2, f >> 2, (f & 3) * 8, lane);
      else if (f >= 8) t_item(sb, 3, 2, 2, 1, (f - 8) * 8, lane);
      store_tasks(sb, At, Np, Dj, 3, blocks(blk(0, 2, 2), blk(0, 3, 2), blk(1, 2, 2)), wi, 12, lane);
      filler_barrier();
      // T[3][J] += L[3][2] X[2][J], J = 0, 1
      if (f >= 0 && f < 8) t_item(sb, 3, f >> 2, 2, 0, (f & 3) * 8, lane);
      store_tasks(sb, At, Np, Dj, 2, blocks(blk(1, 2, 0), blk(1, 2, 1)), wi, 12, lane);
    }
    ++nstamp;
    __syncthreads();
    G3_STAMP();
    // ---- S1: inverse of the diagonal sub-block (warp 0, forward substitution on the identity); log-determinant (warp 15) ----
    if (warp == 0) {
      inverse32(sb, kb, o, lane);
    } else if (warp <= 2) {
      // the scaled diagonal sub-block: L[m][k] = VT[k][m] / sqrt(d_k), lane = row m, 16 columns per warp
      const int k0 = (warp - 1) * 16;
      const uint32_t vp = sb + (OFF_VT + k0 * VS + lane) * B8, ip = sb + (OFF_INVD + o + k0) * B8;
      const uint32_t lp = sb + ((o + lane) * LD + o + k0) * B8;
#pragma unroll
      for (int k = 0; k < 16; ++k) sts(lp + k * B8, lds(vp + k * VS * B8) * lds(ip + k * B8));
    } else if (warp == 15) {
      double v = -log(lds(sb + (OFF_INVD + o + lane) * B8));
#pragma unroll
      for (int o2 = 16; o2 > 0; o2 >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o2);
      ldacc += v;
    }
    __syncthreads();
    G3_STAMP();
    // ---- S2: rows below the diagonal sub-block times its inverse; for the last sub-block the last block row of X ----
    if (kb < 3) {
      if (warp < (3 - kb) * 4) below_item(sb, kb, o + 32 + warp * 8, lane);
    } else if (warp < 12) {
      x_item(sb, 3, warp >> 2, (warp & 3) * 8, lane);
    }
    __syncthreads();
    G3_STAMP();
  }
  // ---- what is left: the last diagonal sub-block of L and the last block row of X ----
  store_tasks(sb, At, Np, Dj, 5, blocks(blk(0, 3, 3), blk(1, 3, 0), blk(1, 3, 1), blk(1, 3, 2), blk(1, 3, 3)), warp, 16, lane);
  if (tid == NT - 32) {   // warp 15, lane 0
    if (logdet) logdet[b] += ldacc;
    if (info && first_bad >= 0 && info[b] == 0) info[b] = j * TS + first_bad + 1;
  }
  G3_STAMP();
#undef G3_STAMP
}

// Zeros above the diagonal of the diagonal tiles j0..j0+nj-1: of their inverses in Dinv (which = 2, BEFORE the factorisation:
// the tile kernel only ever writes Dinv at or below the diagonal and nothing else writes it) or of the tiles of A themselves
// (which = 1, AFTER it: the symmetric GEMM updates write whole diagonal tiles, and nothing on the device reads that part).
// grid (nj, B).
__global__ void __launch_bounds__(256)
diag2_zero_upper_kernel(double* __restrict__ A, int Np, long long strideA, int j0, double* __restrict__ Dinv, int T,
                        const int* __restrict__ bmap, int which) {
  const int j = j0 + blockIdx.x;
  const int b = bmap ? bmap[blockIdx.y] : (int)blockIdx.y;
  double* P = (which == 1) ? A + (long long)b * strideA + (long long)j * TS * Np + (long long)j * TS
                           : Dinv + ((long long)b * T + j) * TS * TS;
  const long long ld = (which == 1) ? Np : TS;
  const double2 z2 = make_double2(0.0, 0.0);
  for (int p = threadIdx.x; p < TS * 64; p += 256) {
    const int r = p >> 6, c = (p & 63) * 2;
    // both columns of the pair are above the diagonal (the pair that straddles it is written with the data)
    if (c > r) *reinterpret_cast<double2*>(P + r * ld + c) = z2;
  }
}

}  // namespace

static constexpr int kDiag2Smem = OFF_END * (int)sizeof(double);

int g3_diag2_launch(g3_ctx* ctx, double* A, int Np, long long strideA, int j, double* Dinv, int T, double* logdet, int* info,
                    const int* bmap, int B, long long* stamps, int dbg) {
  if (!ctx->diag2_ready) {
    G3_CUDA(ctx, cudaFuncSetAttribute(potrf_diag2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDiag2Smem));
    ctx->diag2_ready = true;
  }
  potrf_diag2_kernel<<<B, NT, kDiag2Smem, ctx->stream>>>(A, Np, strideA, j, Dinv, T, logdet, info, bmap, stamps, dbg);
  return 0;
}

int g3_diag2_prepare(g3_ctx* ctx, int j0, int nj, double* Dinv, int T, const int* bmap, int B) {
  diag2_zero_upper_kernel<<<dim3(nj, B), 256, 0, ctx->stream>>>(nullptr, 0, 0, j0, Dinv, T, bmap, 2);
  G3_LAUNCH_CHECK(ctx);
  return 0;
}
int g3_diag2_finish(g3_ctx* ctx, double* A, int Np, long long strideA, int j0, int nj, const int* bmap, int B) {
  diag2_zero_upper_kernel<<<dim3(nj, B), 256, 0, ctx->stream>>>(A, Np, strideA, j0, nullptr, 0, bmap, 1);
  G3_LAUNCH_CHECK(ctx);
  return 0;
}

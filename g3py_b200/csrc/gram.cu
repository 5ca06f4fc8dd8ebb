// Gram-matrix construction and its vector-Jacobian product for sm_100a.
//
// gram_fwd : Kernel.cov(x1, x2) for B hyper samples (g3py/processes/hypers/kernels.py:106-110,
//            192-245,360-472; metrics.py:11-13,89-102) fused with tt_to_num / tt_to_cov
//            (g3py/libs/tensors.py:90-98).  The reference materialises an N1 x N2 x D difference
//            tensor and one N1 x N2 temporary per tree node; here one CTA stages two 128-row X
//            tiles in shared memory, evaluates the whole expression tree in registers and writes
//            the 128x128 output tile with coalesced 16-byte stores.  No pairwise-distance tensor.
// gram_vjp : dtheta_p = scale * sum_ij W_ij dK_ij/dtheta_p, recomputing the K tile from X instead
//            of reading dK/dtheta (replaces Theano's reverse mode through Kernel.cov, which needs
//            several N^2 D passes, tensors.py:11-22).  Per-thread accumulators live in shared
//            memory; CTA partials are reduced in a fixed order (deterministic).
#include "g3b_internal.cuh"
#include <math.h>
#include <stdlib.h>

namespace {

constexpr int TS = G3_TILE;
constexpr double kInfRepl = 1e10;  // tt_to_num: +-inf -> float32(1e10) == 1e10 exactly

struct TileXY {
  int x, y;
};
__device__ __forceinline__ TileXY decode_tile(int tile, int lower_only, int ntx) {
  TileXY t;
  if (!lower_only) {
    t.x = tile % ntx;
    t.y = tile / ntx;
  } else {
    int x = (int)((sqrt(8.0 * (double)tile + 1.0) - 1.0) * 0.5);
    while ((long long)x * (x + 1) / 2 > tile) --x;
    while ((long long)(x + 1) * (x + 2) / 2 <= tile) ++x;
    t.x = x;
    t.y = tile - (int)((long long)x * (x + 1) / 2);
  }
  return t;
}

// Stage the two X tiles: x1s[r][d] (row-major), x2s[d][c] (transposed so lanes read consecutive c).
__device__ __forceinline__ void stage_x(const double* X1, const double* X2, int n1, int n2, int D, int r0, int c0,
                                        double* x1s, double* x2s) {
  for (int idx = threadIdx.x; idx < TS * D; idx += blockDim.x) {
    const int r = idx / D, d = idx - r * D;
    x1s[idx] = (r0 + r < n1) ? X1[(long long)(r0 + r) * D + d] : 0.0;
    x2s[d * TS + r] = (c0 + r < n2) ? X2[(long long)(c0 + r) * D + d] : 0.0;
  }
}

constexpr double kPi2 = 9.869604401089358;   // pi^2 (kernels.py:10)

// one factor of the product-form periodic kernels and its derivative with respect to freq
//   COS / SM: cos(2 pi d f)              kernels.py:466-467,486-487
//   SINC    : sin(2 pi^2 d f)/(2 pi^2 f d), 1 at d = 0   kernels.py:479-482
__device__ __forceinline__ double periodic_factor(int op, double df, double fq) {
  if (op == G3_K_SINC) {
    const double b = 2.0 * kPi2 * df * fq;
    return df != 0.0 ? sin(b) / b : 1.0;
  }
  return cospi(2.0 * df * fq);
}
__device__ __forceinline__ double periodic_dfactor(int op, double df, double fq, double fac) {
  if (op == G3_K_SINC) {
    const double b = 2.0 * kPi2 * df * fq;
    return df != 0.0 ? (cos(b) - fac) / fq : 0.0;
  }
  return -sinpi(2.0 * df * fq) * (2.0 * M_PI * df);
}

// m^p for a small integer p by repeated squaring (Theano specialises pow(x, int) the same way)
__device__ __forceinline__ double ipow(double m, int p) {
  double r = 1.0;
  while (p > 0) {
    if (p & 1) r *= m;
    m *= m;
    p >>= 1;
  }
  return r;
}

// the NN flavour of the dot leaf: value and derivative with respect to the metric m of arcsin(2m / (1 + 2m)^2)
__device__ __forceinline__ void nn_value(double m, double& kk, double& dk) {
  const double q = 1.0 + 2.0 * m, u = 2.0 * m / (q * q);
  kk = asin(u);
  dk = (2.0 * (1.0 - 2.0 * m) / (q * q * q)) * rsqrt(1.0 - u * u);
}
__device__ __forceinline__ double eq_second(const g3_knode& nd) {
  return (nd.flags & G3_KF_EQ2) ? __hiloint2double(nd.p1_idx, nd.p0_idx) : nd.value;
}
// DeltaEq / DeltaEq2 term of one coordinate pair
__device__ __forceinline__ double eq_term(const g3_knode& nd, double e2, double xi, double xj) {
  double t = (xi == nd.value && xj == e2) ? 1.0 : 0.0;
  if ((nd.flags & G3_KF_EQ2) && xi == e2 && xj == nd.value) t += 1.0;
  return t;
}

// value of one leaf on the diagonal of cov(x, x) (same = 1, i == j) for the row `x`
__device__ __forceinline__ double leaf_diag(const g3_knode& nd, const double* __restrict__ th,
                                            const double* __restrict__ x, int skip_pn) {
  const double var = nd.var_idx >= 0 ? th[nd.var_idx] : nd.value;
  switch (nd.op) {
    case G3_K_NOISE:
      return (skip_pn && (nd.flags & G3_KF_PROCESS_NOISE)) ? 0.0 : var;
    case G3_K_RQ:
      return var * pow(1.0, -th[nd.p1_idx]);
    case G3_K_DOT: {
      double m = nd.p1_idx >= 0 ? th[nd.p1_idx] : 0.0;
      for (int k = nd.dim0; k < nd.dim1; ++k) { const double r = th[nd.p0_idx + k - nd.dim0]; m += r * r * x[k] * x[k]; }
      if (nd.flags & G3_KF_NN) {
        double kk, dk;
        nn_value(m, kk, dk);
        return var * kk;
      }
      return var * ipow(m, G3_KF_POWER(nd.flags));
    }
    case G3_K_EQ: {
      const double e2 = eq_second(nd);
      double m = 0.0;
      for (int k = nd.dim0; k < nd.dim1; ++k) m += eq_term(nd, e2, x[k], x[k]);
      return m;
    }
    case G3_K_BW: {
      double m = 1.0;
      for (int k = nd.dim0; k < nd.dim1; ++k) m *= x[k];
      return var * m;
    }
    default:   // stationary leaves: k(0) = 1; WN / VAR: var
      return var;
  }
}

// value of one leaf for 4 columns at once; same_diag[e] = element lies on the diagonal of cov(x, x)
__device__ __forceinline__ void leaf_value4(const g3_knode& nd, const double* __restrict__ th,
                                            const double* __restrict__ x1row, const double* __restrict__ x2s,
                                            const int* cc, const bool* same_diag, int same, int skip_pn,
                                            double* out) {
  const double var = nd.var_idx >= 0 ? th[nd.var_idx] : nd.value;
  const int nd_ = nd.dim1 - nd.dim0;
  double d[4] = {0.0, 0.0, 0.0, 0.0};
  double pr[4] = {1.0, 1.0, 1.0, 1.0};
  switch (nd.op) {
    case G3_K_SE:
    case G3_K_MAT32:
    case G3_K_MAT52:
    case G3_K_RQ:
      for (int k = 0; k < nd_; ++k) {
        const double r = th[nd.p0_idx + k];
        const double hr2 = 0.5 * r * r;
        const double xi = x1row[nd.dim0 + k];
        const double* x2 = x2s + (nd.dim0 + k) * TS;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const double df = xi - x2[cc[e]];
          d[e] += df * df * hr2;
        }
      }
      break;
    case G3_K_OU:
      for (int k = 0; k < nd_; ++k) {
        const double r = th[nd.p0_idx + k];
        const double xi = x1row[nd.dim0 + k];
        const double* x2 = x2s + (nd.dim0 + k) * TS;
#pragma unroll
        for (int e = 0; e < 4; ++e) d[e] += fabs(xi - x2[cc[e]]) * r;
      }
      break;
    case G3_K_SIN:
      for (int k = 0; k < nd_; ++k) {
        const double fq = th[nd.p1_idx + k];
        const double r = th[nd.p0_idx + k];
        const double xi = x1row[nd.dim0 + k];
        const double* x2 = x2s + (nd.dim0 + k) * TS;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const double sn = sinpi((xi - x2[cc[e]]) * fq);
          d[e] += sn * sn * r;
        }
      }
      break;
    case G3_K_WN:
      if (!same)
        for (int k = 0; k < nd_; ++k) {
          const double xi = x1row[nd.dim0 + k];
          const double* x2 = x2s + (nd.dim0 + k) * TS;
#pragma unroll
          for (int e = 0; e < 4; ++e) d[e] += (xi - x2[cc[e]] == 0.0) ? 1.0 : 0.0;
        }
      break;
    case G3_K_COS:
    case G3_K_SINC:
    case G3_K_SM:
      for (int k = 0; k < nd_; ++k) {
        const double fq = th[nd.p1_idx + k];
        const double r = nd.op == G3_K_SM ? th[nd.p0_idx + k] : 0.0;
        const double xi = x1row[nd.dim0 + k];
        const double* x2 = x2s + (nd.dim0 + k) * TS;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const double df = xi - x2[cc[e]];
          pr[e] *= periodic_factor(nd.op, df, fq);
          d[e] += df * df * r * r;
        }
      }
      break;
    case G3_K_DOT:
      for (int k = 0; k < nd_; ++k) {
        const double r = th[nd.p0_idx + k];
        const double xr = x1row[nd.dim0 + k] * (r * r);
        const double* x2 = x2s + (nd.dim0 + k) * TS;
#pragma unroll
        for (int e = 0; e < 4; ++e) d[e] += xr * x2[cc[e]];
      }
      break;
    case G3_K_BW:
      for (int k = 0; k < nd_; ++k) {
        const double xi = x1row[nd.dim0 + k];
        const double* x2 = x2s + (nd.dim0 + k) * TS;
#pragma unroll
        for (int e = 0; e < 4; ++e) pr[e] *= fmin(xi, x2[cc[e]]);
      }
      break;
    case G3_K_EQ: {
      const double e2 = eq_second(nd);
      for (int k = 0; k < nd_; ++k) {
        const double xi = x1row[nd.dim0 + k];
        const double* x2 = x2s + (nd.dim0 + k) * TS;
#pragma unroll
        for (int e = 0; e < 4; ++e) d[e] += eq_term(nd, e2, xi, x2[cc[e]]);
      }
    } break;
    default:
      break;
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    double v;
    switch (nd.op) {
      case G3_K_COS:
      case G3_K_SINC:
        v = var * pr[e];
        break;
      case G3_K_SM:
        v = var * (exp(-2.0 * kPi2 * d[e]) * pr[e]);
        break;
      case G3_K_SE:
      case G3_K_OU:
        v = var * exp(-d[e]);
        break;
      case G3_K_MAT32: {
        const double s = sqrt(3.0 * d[e]);
        v = var * ((1.0 + s) * exp(-s));
      } break;
      case G3_K_MAT52: {
        const double s = sqrt(5.0 * d[e]);
        v = var * ((1.0 + s + 5.0 * d[e] / 3.0) * exp(-s));
      } break;
      case G3_K_RQ: {
        const double al = th[nd.p1_idx];
        v = var * pow(1.0 + d[e] / al, -al);
      } break;
      case G3_K_SIN:
        v = var * exp(2.0 * d[e]);
        break;
      case G3_K_NOISE:
        v = (same_diag[e] && !(skip_pn && (nd.flags & G3_KF_PROCESS_NOISE))) ? var : 0.0;
        break;
      case G3_K_WN:
        v = same ? (same_diag[e] ? var : 0.0) : var * d[e];
        break;
      case G3_K_DOT:
        if (nd.flags & G3_KF_NN) {
          double kk, dk;
          nn_value((nd.p1_idx >= 0 ? th[nd.p1_idx] : 0.0) + d[e], kk, dk);
          v = var * kk;
        } else {
          v = var * ipow((nd.p1_idx >= 0 ? th[nd.p1_idx] : 0.0) + d[e], G3_KF_POWER(nd.flags));
        }
        break;
      case G3_K_EQ:
        v = d[e];
        break;
      case G3_K_BW:
        v = var * pr[e];
        break;
      case G3_K_VAR:
        v = var;
        break;
      default:
        v = 0.0;
    }
    out[e] = v;
  }
}

__device__ __forceinline__ double scrub(double v, int& flag) {
  if (isnan(v)) { flag = 1; return 0.0; }
  if (isinf(v)) { flag = 1; return kInfRepl; }
  return v;
}

__global__ void __launch_bounds__(256)
gram_fwd_kernel(const __grid_constant__ g3_kernel_desc desc, const GramArgs a, int ntx) {
  extern __shared__ double sm[];
  double* x1s = sm;                     // [128][D]
  double* x2s = sm + TS * a.D;          // [D][128]
  double* th = x2s + TS * a.D;          // [P]
  const int b = a.bmap ? a.bmap[blockIdx.y] : (int)blockIdx.y;
  const TileXY t = decode_tile(blockIdx.x, a.lower_only, ntx);
  const int r0 = t.x * TS, c0 = t.y * TS;
  stage_x(a.X1, a.X2, a.n1, a.n2, a.D, r0, c0, x1s, x2s);
  for (int p = threadIdx.x; p < a.P; p += blockDim.x) th[p] = a.theta[(long long)b * a.P + p];
  __syncthreads();

  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int cc[4] = {2 * tx, 2 * tx + 1, 64 + 2 * tx, 64 + 2 * tx + 1};
  const double shift = (a.diag_shift && a.same) ? a.diag_shift[b] : 0.0;
  double* Kb = a.K + (long long)b * a.strideK;
  int flag = 0;
  for (int q = 0; q < 16; ++q) {
    const int rl = ty + 8 * q;
    const int gi = r0 + rl;
    bool sd[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) sd[e] = a.same && (gi + a.diag_off == c0 + cc[e]);
    // shift-register evaluation stack (depth 6) for 4 columns
    double s0[4], s1[4], s2[4], s3[4], s4[4], s5[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) s0[e] = s1[e] = s2[e] = s3[e] = s4[e] = s5[e] = 0.0;
    for (int n = 0; n < desc.n_nodes; ++n) {
      const g3_knode& nd = desc.nodes[n];
      if (nd.op < G3_K_SUM) {
        double v[4];
        leaf_value4(nd, th, x1s + rl * a.D, x2s, cc, sd, a.same, a.skip_process_noise, v);
#pragma unroll
        for (int e = 0; e < 4; ++e) { s5[e] = s4[e]; s4[e] = s3[e]; s3[e] = s2[e]; s2[e] = s1[e]; s1[e] = s0[e]; s0[e] = v[e]; }
      } else if (nd.op == G3_K_SUM || nd.op == G3_K_PROD || nd.op == G3_K_MAX) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          s0[e] = nd.op == G3_K_SUM ? s1[e] + s0[e] : (nd.op == G3_K_PROD ? s1[e] * s0[e] : fmax(s1[e], s0[e]));
          s1[e] = s2[e]; s2[e] = s3[e]; s3[e] = s4[e]; s4[e] = s5[e];
        }
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) s0[e] = nd.op == G3_K_SCALE ? nd.value * s0[e] : nd.value + s0[e];
      }
    }
    double v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int gj = c0 + cc[e];
      double val = scrub(s0[e], flag);
      if (sd[e]) val += shift;
      if (gi >= a.n1 || gj >= a.n2) val = (a.pad_identity && gi + a.diag_off == gj) ? 1.0 : 0.0;
      v[e] = val;
    }
    double* rowp = Kb + (long long)gi * a.ldk + c0;
    *reinterpret_cast<double2*>(rowp + cc[0]) = make_double2(v[0], v[1]);
    *reinterpret_cast<double2*>(rowp + cc[2]) = make_double2(v[2], v[3]);
  }
  if (a.status && flag) atomicOr(a.status + b, G3_ST_NONFINITE_INPUT);
}

// min and mean of diag(cov(X, X)) per theta (after tt_to_num scrubbing), optionally the diagonal itself.
// grid (B), 256 threads.
__global__ void __launch_bounds__(256)
gram_diag_kernel(const __grid_constant__ g3_kernel_desc desc, const double* __restrict__ X, int n, int D,
                 const double* __restrict__ theta, int P, double* __restrict__ dmin, double* __restrict__ dmean,
                 int* __restrict__ status, int skip_pn, double* __restrict__ dvec) {
  __shared__ double th[G3_MAX_THETA];
  __shared__ double rmin[8], rsum[8];
  const int b = blockIdx.x;
  for (int p = threadIdx.x; p < P; p += blockDim.x) th[p] = theta[(long long)b * P + p];
  __syncthreads();
  // diagonal element: all differences are zero -> every stationary leaf evaluates at d = 0; the dot-product /
  // Brownian leaves need the row itself
  double vmin = INFINITY, vsum = 0.0;
  int flag = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double st[6] = {0, 0, 0, 0, 0, 0};
    for (int k = 0; k < desc.n_nodes; ++k) {
      const g3_knode& nd = desc.nodes[k];
      if (nd.op < G3_K_SUM) {
        const double v = leaf_diag(nd, th, X + (long long)i * D, skip_pn);
        st[5] = st[4]; st[4] = st[3]; st[3] = st[2]; st[2] = st[1]; st[1] = st[0]; st[0] = v;
      } else if (nd.op == G3_K_SUM || nd.op == G3_K_PROD || nd.op == G3_K_MAX) {
        st[0] = nd.op == G3_K_SUM ? st[1] + st[0] : (nd.op == G3_K_PROD ? st[1] * st[0] : fmax(st[1], st[0]));
        st[1] = st[2]; st[2] = st[3]; st[3] = st[4]; st[4] = st[5];
      } else {
        st[0] = nd.op == G3_K_SCALE ? nd.value * st[0] : nd.value + st[0];
      }
    }
    const double v = scrub(st[0], flag);
    if (dvec) dvec[(long long)b * n + i] = v;
    vmin = fmin(vmin, v);
    vsum += v;
  }
  for (int o = 16; o > 0; o >>= 1) {
    vmin = fmin(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
    vsum += __shfl_xor_sync(0xffffffffu, vsum, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { rmin[warp] = vmin; rsum[warp] = vsum; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = rmin[0], s = rsum[0];
    for (int w = 1; w < 8; ++w) { m = fmin(m, rmin[w]); s += rsum[w]; }
    dmin[b] = m;
    dmean[b] = s / (double)n;
  }
  if (status && flag) atomicOr(status + b, G3_ST_NONFINITE_INPUT);
}

// ---- VJP -------------------------------------------------------------------------------
// One CTA per 128x128 tile; thread mapping as in gram_fwd.  acc[p][tid] in shared memory.
__global__ void __launch_bounds__(256)
gram_vjp_kernel(const __grid_constant__ g3_kernel_desc desc, const VjpArgs a, int ntx, double* __restrict__ partials,
                int ntiles) {
  extern __shared__ double sm[];
  double* x1s = sm;
  double* x2s = sm + TS * a.D;
  double* th = x2s + TS * a.D;
  double* al_r = th + G3_MAX_THETA;      // alpha rows [128]
  double* al_c = al_r + TS;              // alpha cols [128]
  double* acc = al_c + TS;               // [P][256]
  const int b = blockIdx.y;
  const int tid = threadIdx.x;
  const TileXY t = decode_tile(blockIdx.x, a.lower_only, ntx);
  const int r0 = t.x * TS, c0 = t.y * TS;
  stage_x(a.X1, a.X2, a.n1, a.n2, a.D, r0, c0, x1s, x2s);
  for (int p = tid; p < a.P; p += blockDim.x) th[p] = a.theta[(long long)b * a.P + p];
  if (a.alpha && tid < TS) {
    al_r[tid] = (r0 + tid < a.n1) ? a.alpha[(long long)b * a.strideAlpha + r0 + tid] : 0.0;
    al_c[tid] = (c0 + tid < a.n2) ? (a.alpha2 ? a.alpha2 : a.alpha)[(long long)b * a.strideAlpha + c0 + tid] : 0.0;
  }
  for (int p = 0; p < a.P; ++p) acc[p * 256 + tid] = 0.0;
  __syncthreads();
  const double cf = (a.alpha && a.cfac) ? a.cfac[b] : 1.0;

  const int tx = tid & 31, ty = tid >> 5;
  const int cc[4] = {2 * tx, 2 * tx + 1, 64 + 2 * tx, 64 + 2 * tx + 1};
  const double* Wb = a.W + (long long)b * a.strideW;
  for (int q = 0; q < 16; ++q) {
    const int rl = ty + 8 * q;
    const int gi = r0 + rl;
    const double* x1row = x1s + rl * a.D;
    bool sd[4];
    double w[4];
    {
      const double* wrow = Wb + (long long)gi * a.ldw + c0;
      const bool row_ok = gi < a.n1;
      double2 w01 = make_double2(0.0, 0.0), w23 = make_double2(0.0, 0.0);
      if (row_ok) {
        w01 = *reinterpret_cast<const double2*>(wrow + cc[0]);
        w23 = *reinterpret_cast<const double2*>(wrow + cc[2]);
      }
      w[0] = w01.x; w[1] = w01.y; w[2] = w23.x; w[3] = w23.y;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int gj = c0 + cc[e];
        sd[e] = a.same && (gi == gj);
        if (a.alpha) w[e] = cf * al_r[rl] * al_c[cc[e]] - w[e];
        double f = 1.0;
        if (a.lower_only) f = (gi > gj) ? 2.0 : (gi == gj ? 1.0 : 0.0);
        if (!row_ok || gj >= a.n2) f = 0.0;
        w[e] = f == 0.0 ? 0.0 : w[e] * f;   // also kills NaN garbage in masked positions
      }
    }
    // forward: node values, leaf auxiliaries
    double val[G3_MAX_NODES][4], aux[G3_MAX_NODES][4], dd[G3_MAX_NODES][4], adj[G3_MAX_NODES][4];
    for (int n = 0; n < desc.n_nodes; ++n) {
      const g3_knode& nd = desc.nodes[n];
      if (nd.op < G3_K_SUM) {
        const double var = nd.op == G3_K_EQ ? 1.0 : (nd.var_idx >= 0 ? th[nd.var_idx] : nd.value);
        const int nd_ = nd.dim1 - nd.dim0;
        double d[4] = {0.0, 0.0, 0.0, 0.0};
        if (nd.op == G3_K_EQ) {
          const double e2 = eq_second(nd);
          for (int k = 0; k < nd_; ++k) {
            const double xi = x1row[nd.dim0 + k];
            const double* x2 = x2s + (nd.dim0 + k) * TS;
#pragma unroll
            for (int e = 0; e < 4; ++e) d[e] += eq_term(nd, e2, xi, x2[cc[e]]);
          }
        } else if (nd.op == G3_K_SE || nd.op == G3_K_MAT32 || nd.op == G3_K_MAT52 || nd.op == G3_K_RQ) {
          for (int k = 0; k < nd_; ++k) {
            const double r = th[nd.p0_idx + k];
            const double hr2 = 0.5 * r * r, xi = x1row[nd.dim0 + k];
            const double* x2 = x2s + (nd.dim0 + k) * TS;
#pragma unroll
            for (int e = 0; e < 4; ++e) { const double df = xi - x2[cc[e]]; d[e] += df * df * hr2; }
          }
        } else if (nd.op == G3_K_OU) {
          for (int k = 0; k < nd_; ++k) {
            const double r = th[nd.p0_idx + k], xi = x1row[nd.dim0 + k];
            const double* x2 = x2s + (nd.dim0 + k) * TS;
#pragma unroll
            for (int e = 0; e < 4; ++e) d[e] += fabs(xi - x2[cc[e]]) * r;
          }
        } else if (nd.op == G3_K_SIN) {
          for (int k = 0; k < nd_; ++k) {
            const double fq = th[nd.p1_idx + k], r = th[nd.p0_idx + k], xi = x1row[nd.dim0 + k];
            const double* x2 = x2s + (nd.dim0 + k) * TS;
#pragma unroll
            for (int e = 0; e < 4; ++e) { const double sn = sinpi((xi - x2[cc[e]]) * fq); d[e] += sn * sn * r; }
          }
        } else if (nd.op == G3_K_WN && !a.same) {
          for (int k = 0; k < nd_; ++k) {
            const double xi = x1row[nd.dim0 + k];
            const double* x2 = x2s + (nd.dim0 + k) * TS;
#pragma unroll
            for (int e = 0; e < 4; ++e) d[e] += (xi - x2[cc[e]] == 0.0) ? 1.0 : 0.0;
          }
        }
        double pr[4] = {1.0, 1.0, 1.0, 1.0};
        if (nd.op == G3_K_DOT) {
          for (int k = 0; k < nd_; ++k) {
            const double r = th[nd.p0_idx + k];
            const double xr = x1row[nd.dim0 + k] * (r * r);
            const double* x2 = x2s + (nd.dim0 + k) * TS;
#pragma unroll
            for (int e = 0; e < 4; ++e) d[e] += xr * x2[cc[e]];
          }
        } else if (nd.op == G3_K_BW) {
          for (int k = 0; k < nd_; ++k) {
            const double xi = x1row[nd.dim0 + k];
            const double* x2 = x2s + (nd.dim0 + k) * TS;
#pragma unroll
            for (int e = 0; e < 4; ++e) pr[e] *= fmin(xi, x2[cc[e]]);
          }
        }
        if (nd.op == G3_K_COS || nd.op == G3_K_SINC || nd.op == G3_K_SM) {
          for (int k = 0; k < nd_; ++k) {
            const double fq = th[nd.p1_idx + k], r = nd.op == G3_K_SM ? th[nd.p0_idx + k] : 0.0;
            const double xi = x1row[nd.dim0 + k];
            const double* x2 = x2s + (nd.dim0 + k) * TS;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const double df = xi - x2[cc[e]];
              pr[e] *= periodic_factor(nd.op, df, fq);
              d[e] += df * df * r * r;
            }
          }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          double kk, dk = 0.0;  // kk = k(d) (unit variance), dk = d k / d d
          switch (nd.op) {
            case G3_K_SE: kk = exp(-d[e]); dk = -kk; break;
            case G3_K_OU: kk = exp(-d[e]); dk = -kk; break;
            case G3_K_MAT32: { const double s = sqrt(3.0 * d[e]); const double ex = exp(-s); kk = (1.0 + s) * ex; dk = -1.5 * ex; } break;
            case G3_K_MAT52: { const double s = sqrt(5.0 * d[e]); const double ex = exp(-s);
                               kk = (1.0 + s + 5.0 * d[e] / 3.0) * ex; dk = -(5.0 / 6.0) * (1.0 + s) * ex; } break;
            case G3_K_RQ: { const double al = th[nd.p1_idx]; const double base = 1.0 + d[e] / al;
                            kk = pow(base, -al); dk = -kk / base; } break;
            case G3_K_SIN: kk = exp(2.0 * d[e]); dk = 0.0; break;
            case G3_K_COS:
            case G3_K_SINC: kk = pr[e]; break;
            case G3_K_SM: kk = exp(-2.0 * kPi2 * d[e]) * pr[e]; break;
            case G3_K_NOISE: kk = sd[e] ? 1.0 : 0.0; break;
            case G3_K_WN: kk = a.same ? (sd[e] ? 1.0 : 0.0) : d[e]; break;
            case G3_K_DOT: { const int pw = G3_KF_POWER(nd.flags);
                             const double m = (nd.p1_idx >= 0 ? th[nd.p1_idx] : 0.0) + d[e];
                             if (nd.flags & G3_KF_NN) { nn_value(m, kk, dk); break; }
                             const double m1 = ipow(m, pw - 1); kk = m1 * m; dk = (double)pw * m1; } break;   // dk = d kk / d m
            case G3_K_EQ: kk = d[e]; break;
            case G3_K_BW: kk = pr[e]; break;
            case G3_K_VAR: kk = 1.0; break;
            default: kk = 0.0;
          }
          val[n][e] = var * kk;
          aux[n][e] = kk;
          dd[n][e] = (nd.op == G3_K_RQ) ? d[e] : var * dk;
        }
      } else if (nd.op == G3_K_SUM || nd.op == G3_K_PROD || nd.op == G3_K_MAX) {
        const int l = nd.dim0, r = nd.dim1;
#pragma unroll
        for (int e = 0; e < 4; ++e)
          val[n][e] = nd.op == G3_K_SUM ? val[l][e] + val[r][e]
                                        : (nd.op == G3_K_PROD ? val[l][e] * val[r][e] : fmax(val[l][e], val[r][e]));
      } else {
        const int c = nd.dim0;
#pragma unroll
        for (int e = 0; e < 4; ++e) val[n][e] = nd.op == G3_K_SCALE ? nd.value * val[c][e] : nd.value + val[c][e];
      }
    }
    // backward
    for (int n = 0; n < desc.n_nodes; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) adj[n][e] = 0.0;
#pragma unroll
    for (int e = 0; e < 4; ++e) adj[desc.n_nodes - 1][e] = w[e];
    for (int n = desc.n_nodes - 1; n >= 0; --n) {
      const g3_knode& nd = desc.nodes[n];
      if (nd.op == G3_K_SUM) {
#pragma unroll
        for (int e = 0; e < 4; ++e) { adj[nd.dim0][e] += adj[n][e]; adj[nd.dim1][e] += adj[n][e]; }
      } else if (nd.op == G3_K_PROD) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          adj[nd.dim0][e] += adj[n][e] * val[nd.dim1][e];
          adj[nd.dim1][e] += adj[n][e] * val[nd.dim0][e];
        }
      } else if (nd.op == G3_K_MAX) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {          // Theano: d max(x, y) = eq(out, x) * g, eq(out, y) * g
          if (val[n][e] == val[nd.dim0][e]) adj[nd.dim0][e] += adj[n][e];
          if (val[n][e] == val[nd.dim1][e]) adj[nd.dim1][e] += adj[n][e];
        }
      } else if (nd.op == G3_K_SCALE) {
#pragma unroll
        for (int e = 0; e < 4; ++e) adj[nd.dim0][e] += adj[n][e] * nd.value;
      } else if (nd.op == G3_K_SHIFT) {
#pragma unroll
        for (int e = 0; e < 4; ++e) adj[nd.dim0][e] += adj[n][e];
      } else {
        // leaf: scatter into the per-thread accumulators
        const int nd_ = nd.dim1 - nd.dim0;
        if (nd.var_idx >= 0) {
          double s = 0.0;
#pragma unroll
          for (int e = 0; e < 4; ++e) s += adj[n][e] * aux[n][e];
          acc[nd.var_idx * 256 + tid] += s;
        }
        if (nd.op == G3_K_SE || nd.op == G3_K_MAT32 || nd.op == G3_K_MAT52 || nd.op == G3_K_RQ) {
          const bool rq = nd.op == G3_K_RQ;
          double gk[4];  // adj * var * dk/dd
          if (rq) {
            const double al = th[nd.p1_idx];
            double s = 0.0;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const double base = 1.0 + dd[n][e] / al;
              gk[e] = -adj[n][e] * val[n][e] / base;
              s += adj[n][e] * val[n][e] * (-log1p(dd[n][e] / al) + dd[n][e] / (al + dd[n][e]));
            }
            acc[nd.p1_idx * 256 + tid] += s;
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) gk[e] = adj[n][e] * dd[n][e];
          }
          for (int k = 0; k < nd_; ++k) {
            const double r = th[nd.p0_idx + k], xi = x1row[nd.dim0 + k];
            const double* x2 = x2s + (nd.dim0 + k) * TS;
            double s = 0.0;
#pragma unroll
            for (int e = 0; e < 4; ++e) { const double df = xi - x2[cc[e]]; s += gk[e] * df * df; }
            acc[(nd.p0_idx + k) * 256 + tid] += s * r;
          }
        } else if (nd.op == G3_K_OU) {
          for (int k = 0; k < nd_; ++k) {
            const double xi = x1row[nd.dim0 + k];
            const double* x2 = x2s + (nd.dim0 + k) * TS;
            double s = 0.0;
#pragma unroll
            for (int e = 0; e < 4; ++e) s -= adj[n][e] * val[n][e] * fabs(xi - x2[cc[e]]);
            acc[(nd.p0_idx + k) * 256 + tid] += s;
          }
        } else if (nd.op == G3_K_DOT) {      // m = bias + sum_k r_k^2 x_ik x_jk;  dd = var * d(m^p)/dm
          if (nd.p1_idx >= 0) {
            double s = 0.0;
#pragma unroll
            for (int e = 0; e < 4; ++e) s += adj[n][e] * dd[n][e];
            acc[nd.p1_idx * 256 + tid] += s;
          }
          for (int k = 0; k < nd_; ++k) {
            const double r = th[nd.p0_idx + k], xi = x1row[nd.dim0 + k];
            const double* x2 = x2s + (nd.dim0 + k) * TS;
            double s = 0.0;
#pragma unroll
            for (int e = 0; e < 4; ++e) s += adj[n][e] * dd[n][e] * (xi * x2[cc[e]]);
            acc[(nd.p0_idx + k) * 256 + tid] += 2.0 * r * s;
          }
        } else if (nd.op == G3_K_COS || nd.op == G3_K_SINC || nd.op == G3_K_SM) {
          const double var = nd.var_idx >= 0 ? th[nd.var_idx] : nd.value;
          for (int k = 0; k < nd_; ++k) {
            const double fq = th[nd.p1_idx + k], xi = x1row[nd.dim0 + k];
            const double* x2 = x2s + (nd.dim0 + k) * TS;
            double sf = 0.0, sr = 0.0;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const double df = xi - x2[cc[e]];
              const double fac = periodic_factor(nd.op, df, fq);
              double others = 1.0, dsum = 0.0;                  // product of the other dims' factors (and SM envelope)
              for (int j2 = 0; j2 < nd_; ++j2) {
                const double df2 = x1row[nd.dim0 + j2] - x2s[(nd.dim0 + j2) * TS + cc[e]];
                if (j2 != k) others *= periodic_factor(nd.op, df2, th[nd.p1_idx + j2]);
                if (nd.op == G3_K_SM) { const double r2 = th[nd.p0_idx + j2]; dsum += df2 * df2 * r2 * r2; }
              }
              const double env = nd.op == G3_K_SM ? exp(-2.0 * kPi2 * dsum) : 1.0;
              sf += adj[n][e] * var * env * periodic_dfactor(nd.op, df, fq, fac) * others;
              if (nd.op == G3_K_SM) sr += adj[n][e] * val[n][e] * (-4.0 * kPi2 * df * df * th[nd.p0_idx + k]);
            }
            acc[(nd.p1_idx + k) * 256 + tid] += sf;
            if (nd.op == G3_K_SM) acc[(nd.p0_idx + k) * 256 + tid] += sr;
          }
        } else if (nd.op == G3_K_SIN) {
          for (int k = 0; k < nd_; ++k) {
            const double fq = th[nd.p1_idx + k], r = th[nd.p0_idx + k], xi = x1row[nd.dim0 + k];
            const double* x2 = x2s + (nd.dim0 + k) * TS;
            double sf = 0.0, sr = 0.0;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const double df = xi - x2[cc[e]];
              const double sn = sinpi(df * fq);
              const double gK = adj[n][e] * val[n][e];
              sr += gK * 2.0 * sn * sn;
              sf += gK * 2.0 * r * sinpi(2.0 * df * fq) * (M_PI * df);
            }
            acc[(nd.p1_idx + k) * 256 + tid] += sf;
            acc[(nd.p0_idx + k) * 256 + tid] += sr;
          }
        }
      }
    }
  }
  __syncthreads();
  // block reduction per parameter, fixed order
  const int warp = tid >> 5, lane = tid & 31;
  __shared__ double red[8];
  for (int p = 0; p < a.P; ++p) {
    double v = acc[p * 256 + tid];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
      for (int w = 0; w < 8; ++w) s += red[w];
      partials[((long long)b * ntiles + blockIdx.x) * a.P + p] = s;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
vjp_reduce_kernel(const double* __restrict__ partials, int ntiles, int P, double scale, double* __restrict__ dtheta) {
  __shared__ double red[8];
  const int b = blockIdx.y, p = blockIdx.x;
  double v = 0.0;
  for (int t = threadIdx.x; t < ntiles; t += blockDim.x) v += partials[((long long)b * ntiles + t) * P + p];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w];
    dtheta[(long long)b * P + p] = s * scale;
  }
}

}  // namespace

int g3_check_desc(g3_ctx* ctx, const g3_kernel_desc& d, int D) {
  if (d.n_nodes < 1 || d.n_nodes > G3_MAX_NODES) return g3_fail_msg(ctx, "kernel desc: 1 <= n_nodes <= 32");
  if (d.n_theta < 0 || d.n_theta > G3_MAX_THETA) return g3_fail_msg(ctx, "kernel desc: n_theta <= 64");
  int depth = 0;
  for (int n = 0; n < d.n_nodes; ++n) {
    const g3_knode& nd = d.nodes[n];
    if (nd.op < G3_K_SUM) {
      if (nd.op < G3_K_SE || nd.op > G3_K_EQ) return g3_fail_msg(ctx, "kernel desc: unknown leaf op");
      if (nd.op != G3_K_NOISE && (nd.dim0 < 0 || nd.dim1 > D || nd.dim1 <= nd.dim0))
        return g3_fail_msg(ctx, "kernel desc: leaf dims outside [0, D)");
      const int w = nd.dim1 - nd.dim0;
      if (nd.var_idx >= d.n_theta) return g3_fail_msg(ctx, "kernel desc: var_idx outside theta");
      if (nd.op == G3_K_EQ && nd.var_idx >= 0) return g3_fail_msg(ctx, "kernel desc: the equality leaf has no variance hyper");
      const bool has_rate = nd.op != G3_K_NOISE && nd.op != G3_K_WN && nd.op != G3_K_COS && nd.op != G3_K_SINC &&
                            nd.op != G3_K_BW && nd.op != G3_K_VAR && nd.op != G3_K_EQ;
      if (has_rate && (nd.p0_idx < 0 || nd.p0_idx + w > d.n_theta)) return g3_fail_msg(ctx, "kernel desc: rate index outside theta");
      if (nd.op == G3_K_RQ && (nd.p1_idx < 0 || nd.p1_idx >= d.n_theta)) return g3_fail_msg(ctx, "kernel desc: alpha index outside theta");
      const bool has_freq = nd.op == G3_K_SIN || nd.op == G3_K_COS || nd.op == G3_K_SINC || nd.op == G3_K_SM;
      if (has_freq && (nd.p1_idx < 0 || nd.p1_idx + w > d.n_theta)) return g3_fail_msg(ctx, "kernel desc: freq index outside theta");
      if (nd.op == G3_K_DOT && nd.p1_idx >= d.n_theta) return g3_fail_msg(ctx, "kernel desc: bias index outside theta");
      ++depth;
    } else if (nd.op == G3_K_SUM || nd.op == G3_K_PROD || nd.op == G3_K_MAX) {
      if (depth < 2) return g3_fail_msg(ctx, "kernel desc: malformed post-order tree");
      if (nd.dim0 < 0 || nd.dim0 >= n || nd.dim1 < 0 || nd.dim1 >= n) return g3_fail_msg(ctx, "kernel desc: bad child index");
      --depth;
    } else if (nd.op == G3_K_SCALE || nd.op == G3_K_SHIFT) {
      if (depth < 1 || nd.dim0 < 0 || nd.dim0 >= n) return g3_fail_msg(ctx, "kernel desc: malformed unary node");
    } else {
      return g3_fail_msg(ctx, "kernel desc: unknown op");
    }
    if (depth > 6) return g3_fail_msg(ctx, "kernel desc: expression stack deeper than 6");
  }
  if (depth != 1) return g3_fail_msg(ctx, "kernel desc: tree does not reduce to one value");
  return 0;
}

// fast paths for additive trees on <= 4 input columns (gram_add.cu)
bool g3_desc_is_additive(const g3_kernel_desc& d);
void g3_gram_fwd_add_launch(const g3_kernel_desc& desc, const GramArgs& a, int ntx, dim3 grid, cudaStream_t s);
int g3_gram_vjp_add_launch(g3_ctx* ctx, const g3_kernel_desc& desc, const VjpArgs& a, int ntx, double* partials,
                           int ntiles, dim3 grid, cudaStream_t s);
// second-generation fast path (gram_add.cu, "fast2"): metric leaves on <= 8 columns, leaf-outer loop nest
bool g3_desc_is_fast2(const g3_kernel_desc& d, int same);
int g3_gram_fwd_fast2_launch(g3_ctx* ctx, const g3_kernel_desc& desc, const GramArgs& a, int ntx, dim3 grid, cudaStream_t s);
int g3_gram_vjp_fast2_launch(g3_ctx* ctx, const g3_kernel_desc& desc, const VjpArgs& a, int ntx, double* partials, int ntiles,
                             dim3 grid, cudaStream_t s);
static int fast_mode() {           // G3_NO_FAST=1: generic interpreter only; G3_NO_FAST=2: first-generation fast path only
  static int mode = -1;
  if (mode < 0) { const char* e = getenv("G3_NO_FAST"); mode = e ? atoi(e) : 0; }
  return mode;
}
// the fast paths keep their shared-memory budget: small trees only (the usual case); larger ones take the generic interpreter
static bool small_tree(const g3_kernel_desc& desc) { return desc.n_nodes <= 16 && desc.n_theta <= 32; }
static bool use_fast2(const g3_kernel_desc& desc, int D, int same) {
  return fast_mode() == 0 && D <= 8 && small_tree(desc) && g3_desc_is_fast2(desc, same);
}
static bool use_fast_path(const g3_kernel_desc& desc, int D) {
  static int off = -1;
  if (off < 0) off = fast_mode() == 1 ? 1 : 0;
  return !off && D <= 4 && small_tree(desc) && g3_desc_is_additive(desc);
}

int g3_gram_launch(g3_ctx* ctx, const g3_kernel_desc& desc, const GramArgs& a, int B) {
  int rc = g3_check_desc(ctx, desc, a.D);
  if (rc) return rc;
  const int tr = (a.n1 + TS - 1) / TS, tc = (a.n2 + TS - 1) / TS;
  const long long ntiles = a.lower_only ? (long long)tr * (tr + 1) / 2 : (long long)tr * tc;
  const size_t smem = sizeof(double) * (2 * TS * a.D + G3_MAX_THETA);
  g3_prof_begin(ctx, G3_PROF_GRAM);
  if (use_fast2(desc, a.D, a.same)) {
    if ((rc = g3_gram_fwd_fast2_launch(ctx, desc, a, tr, dim3((unsigned)ntiles, B), ctx->stream))) return rc;
  } else if (use_fast_path(desc, a.D))
    g3_gram_fwd_add_launch(desc, a, tr, dim3((unsigned)ntiles, B), ctx->stream);
  else
    gram_fwd_kernel<<<dim3((unsigned)ntiles, B), 256, smem, ctx->stream>>>(desc, a, tr);
  g3_prof_end(ctx);
  G3_LAUNCH_CHECK(ctx);
  return 0;
}

int g3_gram_diag_min(g3_ctx* ctx, const g3_kernel_desc& desc, const double* X, int n, int D, const double* theta,
                     int P, int B, double* diag_min, double* diag_mean, int* status, int skip_process_noise,
                     double* diag_vec) {
  int rc = g3_check_desc(ctx, desc, D);
  if (rc) return rc;
  gram_diag_kernel<<<B, 256, 0, ctx->stream>>>(desc, X, n, D, theta, P, diag_min, diag_mean, status,
                                               skip_process_noise, diag_vec);
  G3_LAUNCH_CHECK(ctx);
  return 0;
}

int g3_gram_vjp_launch(g3_ctx* ctx, const g3_kernel_desc& desc, const VjpArgs& a, int B) {
  int rc = g3_check_desc(ctx, desc, a.D);
  if (rc) return rc;
  const int tr = (a.n1 + TS - 1) / TS, tc = (a.n2 + TS - 1) / TS;
  const long long ntiles = a.lower_only ? (long long)tr * (tr + 1) / 2 : (long long)tr * tc;
  double* partials = a.partials ? a.partials
                                : (double*)g3_ws(ctx, "vjp_partials", sizeof(double) * (size_t)B * ntiles * (a.P > 0 ? a.P : 1));
  if (!partials) return -2;
  const size_t smem = sizeof(double) * (2 * TS * a.D + G3_MAX_THETA + 2 * TS + (size_t)a.P * 256);
  static bool attr = false;
  if (!attr) {
    G3_CUDA(ctx, cudaFuncSetAttribute(gram_vjp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  if (a.P == 0) return 0;
  g3_prof_begin(ctx, G3_PROF_VJP);
  if (use_fast2(desc, a.D, a.same)) {
    if ((rc = g3_gram_vjp_fast2_launch(ctx, desc, a, tr, partials, (int)ntiles, dim3((unsigned)ntiles, B), ctx->stream))) return rc;
  } else if (use_fast_path(desc, a.D)) {
    if ((rc = g3_gram_vjp_add_launch(ctx, desc, a, tr, partials, (int)ntiles, dim3((unsigned)ntiles, B), ctx->stream))) return rc;
  } else {
    gram_vjp_kernel<<<dim3((unsigned)ntiles, B), 256, smem, ctx->stream>>>(desc, a, tr, partials, (int)ntiles);
  }
  g3_prof_end(ctx);
  G3_LAUNCH_CHECK(ctx);
  vjp_reduce_kernel<<<dim3(a.P, B), 256, 0, ctx->stream>>>(partials, (int)ntiles, a.P, a.scale, a.dtheta);
  G3_LAUNCH_CHECK(ctx);
  return 0;
}

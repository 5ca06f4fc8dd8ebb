// Fast paths of gram_fwd / gram_vjp for ADDITIVE kernel trees (every internal node is a KernelSum, which is
// what EllipticalProcess builds for "k1 + k2 + ... + Noise", g3py/processes/elliptical.py:26-28) on inputs with
// at most 4 columns.  Same arithmetic as the interpreter in gram.cu, but
//   * the thread's four X2 columns live in registers for the whole tile, the coordinate differences are formed
//     once per row and shared by all leaves (statically indexed, no local memory),
//   * no expression stack: leaf values are accumulated as they are produced,
//   * VJP: the adjoint of every leaf is W itself, so each leaf's contributions are reduced immediately.
// ncu on the interpreter (N=4096, B=64, SE+MAT52+Noise): fp64 pipe 23 % / issue slots 59 % busy for gram_fwd,
// 2 KiB of local memory per thread in gram_vjp - both instruction-bound, not HBM-bound.
#include "g3b_internal.cuh"
#include <math.h>

namespace {

constexpr int TS = G3_TILE;
constexpr double kInfRepl = 1e10;

__device__ __forceinline__ void decode_xy(int tile, int lower_only, int ntx, int& tx_, int& ty_) {
  if (!lower_only) {
    tx_ = tile % ntx;
    ty_ = tile / ntx;
  } else {
    int x = (int)((sqrt(8.0 * (double)tile + 1.0) - 1.0) * 0.5);
    while ((long long)x * (x + 1) / 2 > tile) --x;
    while ((long long)(x + 1) * (x + 2) / 2 <= tile) ++x;
    tx_ = x;
    ty_ = tile - (int)((long long)x * (x + 1) / 2);
  }
}

struct LeafTab {             // per-leaf constants staged in shared memory once per CTA
  int op, dim0, dim1, var_idx, p0_idx, p1_idx, flags, pad;
  double var;
  double c[4];               // SE/MAT/RQ: 0.5 r_k^2 ; OU: r_k ; SIN: r_k      (0 outside [dim0, dim1))
  double f[4];               // SIN: freq_k ; RQ: f[0] = alpha
  double r[4];               // raw rate_k (VJP chain rule)
};

template <int DT>
__device__ __forceinline__ void stage_tile(const double* X1, const double* X2, int n1, int n2, int r0, int c0,
                                           double* x1s, double* x2s) {
  for (int idx = threadIdx.x; idx < TS * DT; idx += blockDim.x) {
    const int r = idx / DT, d = idx - r * DT;
    x1s[idx] = (r0 + r < n1) ? X1[(long long)(r0 + r) * DT + d] : 0.0;
    x2s[d * TS + r] = (c0 + r < n2) ? X2[(long long)(c0 + r) * DT + d] : 0.0;
  }
}

__device__ __forceinline__ int build_leaf_table(const g3_kernel_desc& desc, const double* th, LeafTab* tab, int skip_pn) {
  // executed by thread 0; returns the number of leaves
  int nl = 0;
  for (int n = 0; n < desc.n_nodes; ++n) {
    const g3_knode& nd = desc.nodes[n];
    if (nd.op >= G3_K_SUM) continue;
    LeafTab& t = tab[nl++];
    t.op = nd.op; t.dim0 = nd.dim0; t.dim1 = nd.dim1; t.var_idx = nd.var_idx; t.p0_idx = nd.p0_idx; t.p1_idx = nd.p1_idx;
    t.flags = nd.flags;
    t.var = nd.var_idx >= 0 ? th[nd.var_idx] : nd.value;
    if (nd.op == G3_K_NOISE && skip_pn && (nd.flags & G3_KF_PROCESS_NOISE)) t.var = 0.0;
    const bool has_freq = nd.op == G3_K_SIN || nd.op == G3_K_COS || nd.op == G3_K_SINC || nd.op == G3_K_SM;
    for (int k = 0; k < 4; ++k) {
      t.c[k] = 0.0; t.f[k] = 0.0; t.r[k] = 0.0;
      if (k >= nd.dim0 && k < nd.dim1) {
        if (nd.p0_idx >= 0) {
          const double r = th[nd.p0_idx + (k - nd.dim0)];
          t.r[k] = r;
          t.c[k] = (nd.op == G3_K_OU || nd.op == G3_K_SIN) ? r : (nd.op == G3_K_SM ? r * r : 0.5 * r * r);
        }
        if (has_freq) t.f[k] = th[nd.p1_idx + (k - nd.dim0)];
      }
    }
    if (nd.op == G3_K_RQ) t.f[0] = th[nd.p1_idx];
  }
  return nl;
}

// metric value d for 4 columns from the shared differences (static indexing over k)
template <int DT>
__device__ __forceinline__ void leaf_metric(const LeafTab& t, const double (&df)[DT][4], int same, double (&d)[4]) {
#pragma unroll
  for (int e = 0; e < 4; ++e) d[e] = 0.0;
  if (t.op == G3_K_SE || t.op == G3_K_MAT32 || t.op == G3_K_MAT52 || t.op == G3_K_RQ || t.op == G3_K_SM) {
#pragma unroll
    for (int k = 0; k < DT; ++k) {
      const double c = t.c[k];
#pragma unroll
      for (int e = 0; e < 4; ++e) d[e] += df[k][e] * df[k][e] * c;
    }
  } else if (t.op == G3_K_OU) {
#pragma unroll
    for (int k = 0; k < DT; ++k) {
      const double c = t.c[k];
#pragma unroll
      for (int e = 0; e < 4; ++e) d[e] += fabs(df[k][e]) * c;
    }
  } else if (t.op == G3_K_SIN) {
#pragma unroll
    for (int k = 0; k < DT; ++k) {
      if (k >= t.dim0 && k < t.dim1) {
        const double c = t.c[k], f = t.f[k];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const double sn = sinpi(df[k][e] * f);
          d[e] += sn * sn * c;
        }
      }
    }
  } else if (t.op == G3_K_WN && !same) {
#pragma unroll
    for (int k = 0; k < DT; ++k) {
      if (k >= t.dim0 && k < t.dim1) {
#pragma unroll
        for (int e = 0; e < 4; ++e) d[e] += (df[k][e] == 0.0) ? 1.0 : 0.0;
      }
    }
  }
}

constexpr double kPi2 = 9.869604401089358;   // pi^2

__device__ __forceinline__ double pfactor(int op, double df, double fq) {      // see gram.cu periodic_factor
  if (op == G3_K_SINC) {
    const double b = 2.0 * kPi2 * df * fq;
    return df != 0.0 ? sin(b) / b : 1.0;
  }
  return cospi(2.0 * df * fq);
}
__device__ __forceinline__ double pdfactor(int op, double df, double fq, double fac) {
  if (op == G3_K_SINC) {
    const double b = 2.0 * kPi2 * df * fq;
    return df != 0.0 ? (cos(b) - fac) / fq : 0.0;
  }
  return -sinpi(2.0 * df * fq) * (2.0 * M_PI * df);
}
__device__ __forceinline__ bool is_product_leaf(int op) { return op == G3_K_COS || op == G3_K_SINC || op == G3_K_SM; }

// product-form periodic leaves: fac[k][e] (1 outside the leaf's dims), pr[e] = prod_k fac
template <int DT>
__device__ __forceinline__ void leaf_product(const LeafTab& t, const double (&df)[DT][4], double (&fac)[DT][4], double (&pr)[4]) {
#pragma unroll
  for (int e = 0; e < 4; ++e) pr[e] = 1.0;
#pragma unroll
  for (int k = 0; k < DT; ++k) {
    const bool in = k >= t.dim0 && k < t.dim1;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      fac[k][e] = in ? pfactor(t.op, df[k][e], t.f[k]) : 1.0;
      pr[e] *= fac[k][e];
    }
  }
}

// k(d) (unit variance) and dk/dd
__device__ __forceinline__ void leaf_k(const LeafTab& t, double d, bool on_diag, int same, double& kk, double& dk) {
  dk = 0.0;
  switch (t.op) {
    case G3_K_SE:
    case G3_K_OU: kk = exp(-d); dk = -kk; break;
    case G3_K_MAT32: { const double s = sqrt(3.0 * d), ex = exp(-s); kk = (1.0 + s) * ex; dk = -1.5 * ex; } break;
    case G3_K_MAT52: { const double s = sqrt(5.0 * d), ex = exp(-s); kk = (1.0 + s + 5.0 * d / 3.0) * ex;
                       dk = -(5.0 / 6.0) * (1.0 + s) * ex; } break;
    case G3_K_RQ: { const double al = t.f[0], base = 1.0 + d / al; kk = pow(base, -al); dk = -kk / base; } break;
    case G3_K_SIN: kk = exp(2.0 * d); break;
    case G3_K_COS:
    case G3_K_SINC: kk = 1.0; break;                       // times the product, applied by the caller
    case G3_K_SM: kk = exp(-2.0 * kPi2 * d); break;        // envelope; times the product, applied by the caller
    case G3_K_NOISE: kk = on_diag ? 1.0 : 0.0; break;
    case G3_K_WN: kk = same ? (on_diag ? 1.0 : 0.0) : d; break;
    default: kk = 0.0;
  }
}

template <int DT>
__global__ void __launch_bounds__(256, 3)
gram_fwd_add_kernel(const __grid_constant__ g3_kernel_desc desc, const GramArgs a, int ntx) {
  extern __shared__ double sm[];
  double* x1s = sm;                       // [128][DT]
  double* x2s = sm + TS * DT;             // [DT][128]
  double* th = x2s + TS * DT;             // [G3_MAX_THETA]
  LeafTab* tab = reinterpret_cast<LeafTab*>(th + G3_MAX_THETA);
  __shared__ int n_leaves;
  const int b = a.bmap ? a.bmap[blockIdx.y] : (int)blockIdx.y;
  int tx_, ty_;
  decode_xy(blockIdx.x, a.lower_only, ntx, tx_, ty_);
  const int r0 = tx_ * TS, c0 = ty_ * TS;
  stage_tile<DT>(a.X1, a.X2, a.n1, a.n2, r0, c0, x1s, x2s);
  for (int p = threadIdx.x; p < a.P; p += blockDim.x) th[p] = a.theta[(long long)b * a.P + p];
  __syncthreads();
  if (threadIdx.x == 0) n_leaves = build_leaf_table(desc, th, tab, a.skip_process_noise);
  __syncthreads();
  const int nl = n_leaves;

  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int cc[4] = {2 * tx, 2 * tx + 1, 64 + 2 * tx, 64 + 2 * tx + 1};
  double xr[DT][4];
#pragma unroll
  for (int k = 0; k < DT; ++k)
#pragma unroll
    for (int e = 0; e < 4; ++e) xr[k][e] = x2s[k * TS + cc[e]];
  const double shift = (a.diag_shift && a.same) ? a.diag_shift[b] : 0.0;
  double* Kb = a.K + (long long)b * a.strideK;
  int flag = 0;
  for (int q = 0; q < 16; ++q) {
    const int rl = ty + 8 * q;
    const int gi = r0 + rl;
    double df[DT][4];
#pragma unroll
    for (int k = 0; k < DT; ++k) {
      const double xi = x1s[rl * DT + k];
#pragma unroll
      for (int e = 0; e < 4; ++e) df[k][e] = xi - xr[k][e];
    }
    bool sd[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) sd[e] = a.same && (gi + a.diag_off == c0 + cc[e]);
    double sum[4] = {0.0, 0.0, 0.0, 0.0};
    for (int l = 0; l < nl; ++l) {
      const LeafTab& t = tab[l];
      double d[4];
      leaf_metric<DT>(t, df, a.same, d);
      double pr[4] = {1.0, 1.0, 1.0, 1.0};
      if (is_product_leaf(t.op)) {
        double fac[DT][4];
        leaf_product<DT>(t, df, fac, pr);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        double kk, dk;
        leaf_k(t, d[e], sd[e], a.same, kk, dk);
        sum[e] += t.var * kk * pr[e];
      }
    }
    double v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int gj = c0 + cc[e];
      double val = sum[e];
      if (isnan(val)) { flag = 1; val = 0.0; }
      else if (isinf(val)) { flag = 1; val = kInfRepl; }
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

template <int DT>
__global__ void __launch_bounds__(256, 2)
gram_vjp_add_kernel(const __grid_constant__ g3_kernel_desc desc, const VjpArgs a, int ntx, double* __restrict__ partials,
                    int ntiles) {
  extern __shared__ double sm[];
  double* x1s = sm;
  double* x2s = sm + TS * DT;
  double* th = x2s + TS * DT;
  double* al_r = th + G3_MAX_THETA;
  double* al_c = al_r + TS;
  LeafTab* tab = reinterpret_cast<LeafTab*>(al_c + TS);
  double* acc = reinterpret_cast<double*>(tab + G3_MAX_NODES);   // [P][256]
  __shared__ int n_leaves;
  __shared__ double red[8];
  const int b = blockIdx.y, tid = threadIdx.x;
  int tx_, ty_;
  decode_xy(blockIdx.x, a.lower_only, ntx, tx_, ty_);
  const int r0 = tx_ * TS, c0 = ty_ * TS;
  stage_tile<DT>(a.X1, a.X2, a.n1, a.n2, r0, c0, x1s, x2s);
  for (int p = tid; p < a.P; p += blockDim.x) th[p] = a.theta[(long long)b * a.P + p];
  if (a.alpha && tid < TS) {
    al_r[tid] = (r0 + tid < a.n1) ? a.alpha[(long long)b * a.strideAlpha + r0 + tid] : 0.0;
    al_c[tid] = (c0 + tid < a.n2) ? (a.alpha2 ? a.alpha2 : a.alpha)[(long long)b * a.strideAlpha + c0 + tid] : 0.0;
  }
  for (int p = 0; p < a.P; ++p) acc[p * 256 + tid] = 0.0;
  __syncthreads();
  if (tid == 0) n_leaves = build_leaf_table(desc, th, tab, 0);
  __syncthreads();
  const int nl = n_leaves;
  const double cf = (a.alpha && a.cfac) ? a.cfac[b] : 1.0;

  const int tx = tid & 31, ty = tid >> 5;
  const int cc[4] = {2 * tx, 2 * tx + 1, 64 + 2 * tx, 64 + 2 * tx + 1};
  double xr[DT][4];
#pragma unroll
  for (int k = 0; k < DT; ++k)
#pragma unroll
    for (int e = 0; e < 4; ++e) xr[k][e] = x2s[k * TS + cc[e]];
  double alc[4] = {0.0, 0.0, 0.0, 0.0};
  if (a.alpha) {
#pragma unroll
    for (int e = 0; e < 4; ++e) alc[e] = al_c[cc[e]];
  }
  const double* Wb = a.W + (long long)b * a.strideW;
  for (int q = 0; q < 16; ++q) {
    const int rl = ty + 8 * q;
    const int gi = r0 + rl;
    const bool row_ok = gi < a.n1;
    double w[4];
    bool sd[4];
    {
      const double* wrow = Wb + (long long)gi * a.ldw + c0;
      double2 w01 = make_double2(0.0, 0.0), w23 = make_double2(0.0, 0.0);
      if (row_ok) {
        w01 = *reinterpret_cast<const double2*>(wrow + cc[0]);
        w23 = *reinterpret_cast<const double2*>(wrow + cc[2]);
      }
      w[0] = w01.x; w[1] = w01.y; w[2] = w23.x; w[3] = w23.y;
      const double ar = a.alpha ? al_r[rl] : 0.0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int gj = c0 + cc[e];
        sd[e] = a.same && (gi == gj);
        if (a.alpha) w[e] = cf * ar * alc[e] - w[e];
        double f = 1.0;
        if (a.lower_only) f = (gi > gj) ? 2.0 : (gi == gj ? 1.0 : 0.0);
        if (!row_ok || gj >= a.n2) f = 0.0;
        w[e] = f == 0.0 ? 0.0 : w[e] * f;
      }
    }
    double df[DT][4];
#pragma unroll
    for (int k = 0; k < DT; ++k) {
      const double xi = x1s[rl * DT + k];
#pragma unroll
      for (int e = 0; e < 4; ++e) df[k][e] = xi - xr[k][e];
    }
    for (int l = 0; l < nl; ++l) {
      const LeafTab& t = tab[l];
      double d[4], kk[4], gk[4];            // gk = w * var * dk/dd
      leaf_metric<DT>(t, df, a.same, d);
      if (is_product_leaf(t.op)) {
        double fac[DT][4], pr[4], env[4];
        leaf_product<DT>(t, df, fac, pr);
        double svar2 = 0.0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          env[e] = t.op == G3_K_SM ? exp(-2.0 * kPi2 * d[e]) : 1.0;
          svar2 += w[e] * env[e] * pr[e];
        }
        if (t.var_idx >= 0) acc[t.var_idx * 256 + tid] += svar2;
#pragma unroll
        for (int k = 0; k < DT; ++k) {
          if (k >= t.dim0 && k < t.dim1) {
            double sf = 0.0, sr = 0.0;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              double others = 1.0;
#pragma unroll
              for (int j2 = 0; j2 < DT; ++j2)
                if (j2 != k) others *= fac[j2][e];
              sf += w[e] * t.var * env[e] * pdfactor(t.op, df[k][e], t.f[k], fac[k][e]) * others;
              sr += w[e] * t.var * env[e] * pr[e] * (-4.0 * kPi2 * df[k][e] * df[k][e] * t.r[k]);
            }
            acc[(t.p1_idx + k - t.dim0) * 256 + tid] += sf;
            if (t.op == G3_K_SM) acc[(t.p0_idx + k - t.dim0) * 256 + tid] += sr;
          }
        }
        continue;
      }
      double svar = 0.0, salpha = 0.0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        double dk;
        leaf_k(t, d[e], sd[e], a.same, kk[e], dk);
        svar += w[e] * kk[e];
        gk[e] = w[e] * t.var * dk;
        if (t.op == G3_K_RQ) {
          const double al = t.f[0];
          salpha += w[e] * t.var * kk[e] * (-log1p(d[e] / al) + d[e] / (al + d[e]));
        }
      }
      if (t.var_idx >= 0) acc[t.var_idx * 256 + tid] += svar;
      if (t.op == G3_K_RQ) acc[t.p1_idx * 256 + tid] += salpha;
      if (t.op == G3_K_SE || t.op == G3_K_MAT32 || t.op == G3_K_MAT52 || t.op == G3_K_RQ) {
#pragma unroll
        for (int k = 0; k < DT; ++k) {
          if (k >= t.dim0 && k < t.dim1) {
            double s = 0.0;
#pragma unroll
            for (int e = 0; e < 4; ++e) s += gk[e] * df[k][e] * df[k][e];
            acc[(t.p0_idx + k - t.dim0) * 256 + tid] += s * t.r[k];
          }
        }
      } else if (t.op == G3_K_OU) {
#pragma unroll
        for (int k = 0; k < DT; ++k) {
          if (k >= t.dim0 && k < t.dim1) {
            double s = 0.0;
#pragma unroll
            for (int e = 0; e < 4; ++e) s += gk[e] * fabs(df[k][e]);
            acc[(t.p0_idx + k - t.dim0) * 256 + tid] += s;
          }
        }
      } else if (t.op == G3_K_SIN) {
#pragma unroll
        for (int k = 0; k < DT; ++k) {
          if (k >= t.dim0 && k < t.dim1) {
            double sf = 0.0, sr = 0.0;
            const double fq = t.f[k], r = t.r[k];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const double sn = sinpi(df[k][e] * fq);
              const double gK = w[e] * t.var * kk[e];
              sr += gK * 2.0 * sn * sn;
              sf += gK * 2.0 * r * sinpi(2.0 * df[k][e] * fq) * (M_PI * df[k][e]);
            }
            acc[(t.p1_idx + k - t.dim0) * 256 + tid] += sf;
            acc[(t.p0_idx + k - t.dim0) * 256 + tid] += sr;
          }
        }
      }
    }
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  for (int p = 0; p < a.P; ++p) {
    double v = acc[p * 256 + tid];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
      for (int w2 = 0; w2 < 8; ++w2) s += red[w2];
      partials[((long long)b * ntiles + blockIdx.x) * a.P + p] = s;
    }
    __syncthreads();
  }
}

}  // namespace

bool g3_desc_is_additive(const g3_kernel_desc& d) {
  for (int n = 0; n < d.n_nodes; ++n) {
    if (d.nodes[n].op >= G3_K_SUM && d.nodes[n].op != G3_K_SUM) return false;
    if (d.nodes[n].op > G3_K_SM && d.nodes[n].op < G3_K_SUM) return false;   // DOT / BW / VAR: generic interpreter
  }
  return true;
}

size_t g3_gram_add_smem(int D) { return sizeof(double) * (2 * TS * D + G3_MAX_THETA) + sizeof(LeafTab) * G3_MAX_NODES; }
size_t g3_vjp_add_smem(int D, int P) {
  return sizeof(double) * (2 * TS * D + G3_MAX_THETA + 2 * TS + (size_t)P * 256) + sizeof(LeafTab) * G3_MAX_NODES;
}

void g3_gram_fwd_add_launch(const g3_kernel_desc& desc, const GramArgs& a, int ntx, dim3 grid, cudaStream_t s) {
  const size_t smem = g3_gram_add_smem(a.D);
  switch (a.D) {
    case 1: gram_fwd_add_kernel<1><<<grid, 256, smem, s>>>(desc, a, ntx); break;
    case 2: gram_fwd_add_kernel<2><<<grid, 256, smem, s>>>(desc, a, ntx); break;
    case 3: gram_fwd_add_kernel<3><<<grid, 256, smem, s>>>(desc, a, ntx); break;
    default: gram_fwd_add_kernel<4><<<grid, 256, smem, s>>>(desc, a, ntx); break;
  }
}

int g3_gram_vjp_add_launch(g3_ctx* ctx, const g3_kernel_desc& desc, const VjpArgs& a, int ntx, double* partials,
                           int ntiles, dim3 grid, cudaStream_t s) {
  const size_t smem = g3_vjp_add_smem(a.D, a.P);
  static bool attr = false;
  if (!attr) {
    G3_CUDA(ctx, cudaFuncSetAttribute(gram_vjp_add_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    G3_CUDA(ctx, cudaFuncSetAttribute(gram_vjp_add_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    G3_CUDA(ctx, cudaFuncSetAttribute(gram_vjp_add_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    G3_CUDA(ctx, cudaFuncSetAttribute(gram_vjp_add_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr = true;
  }
  switch (a.D) {
    case 1: gram_vjp_add_kernel<1><<<grid, 256, smem, s>>>(desc, a, ntx, partials, ntiles); break;
    case 2: gram_vjp_add_kernel<2><<<grid, 256, smem, s>>>(desc, a, ntx, partials, ntiles); break;
    case 3: gram_vjp_add_kernel<3><<<grid, 256, smem, s>>>(desc, a, ntx, partials, ntiles); break;
    default: gram_vjp_add_kernel<4><<<grid, 256, smem, s>>>(desc, a, ntx, partials, ntiles); break;
  }
  return 0;
}

// ====================================================================================================================
// Second-generation fast path ("fast2"): additive trees whose leaves are SE / OU / MAT32 / MAT52 / RQ / SIN / Noise / WN
// on up to 8 input columns.  ncu on the first fast path (r01c): fp64 pipe 40 % busy but issue slots 74 % - 2.7 non-fp64
// instructions per fp64 instruction (per-element leaf-table loads, a runtime switch per element, shared-memory
// accumulators in the VJP).  Here the loop nest is turned inside out:
//   for each leaf (ONE runtime switch per leaf and tile)  ->  templated body, leaf constants in registers
//     for 16 rows x 4 columns per thread                  ->  straight-line fp64 code
// and the coordinates are centred on X2[0] and pre-scaled by sqrt(0.5) rate_k per leaf, so the ARD metric costs one DADD
// and one DFMA per dimension and element:  d = sum_k (s_k (x_ik - o_k) - s_k (x_jk - o_k))^2.  The same origin for every
// tile keeps K[i][j] == K[j][i] to the bit.  Gradients with respect to rate_k use dK/drate_k = dk/dd * (2 / rate_k) * d_k.
// ====================================================================================================================
namespace {

struct Leaf2 {
  int op, dim0, dim1, var_idx, p0_idx, p1_idx;
  double var;
  double s[8];      // SE/MAT/RQ: sqrt(0.5) rate_k; OU: rate_k; SIN: freq_k   (0 outside [dim0, dim1))
  double r[8];      // rate_k (SIN: rate_k)
  double alpha;     // RQ
};

__device__ __forceinline__ int build_leaf2(const g3_kernel_desc& desc, const double* th, Leaf2* tab, int skip_pn, double* noise_var) {
  int nl = 0;
  double nv = 0.0;
  for (int n = 0; n < desc.n_nodes; ++n) {
    const g3_knode& nd = desc.nodes[n];
    if (nd.op >= G3_K_SUM) continue;
    const double var = nd.var_idx >= 0 ? th[nd.var_idx] : nd.value;
    if (nd.op == G3_K_NOISE || nd.op == G3_K_WN) {  // diagonal-only leaves (WN: cov(x) form only, see g3_desc_is_fast2) are folded into one scalar; their
      if (!(skip_pn && (nd.flags & G3_KF_PROCESS_NOISE))) nv += var;    // var gradient is handled by the caller
      Leaf2& t = tab[nl++];
      t.op = nd.op; t.var_idx = nd.var_idx; t.var = var; t.dim0 = t.dim1 = 0; t.p0_idx = t.p1_idx = -1;
      continue;
    }
    Leaf2& t = tab[nl++];
    t.op = nd.op; t.dim0 = nd.dim0; t.dim1 = nd.dim1; t.var_idx = nd.var_idx; t.p0_idx = nd.p0_idx; t.p1_idx = nd.p1_idx;
    t.var = var;
    t.alpha = nd.op == G3_K_RQ ? th[nd.p1_idx] : 0.0;
    for (int k = 0; k < 8; ++k) {
      t.s[k] = 0.0; t.r[k] = 0.0;
      if (k >= nd.dim0 && k < nd.dim1 && nd.p0_idx >= 0) {
        const double r = th[nd.p0_idx + (k - nd.dim0)];
        t.r[k] = r;
        t.s[k] = nd.op == G3_K_OU ? r : (nd.op == G3_K_SIN ? th[nd.p1_idx + (k - nd.dim0)] : 0.70710678118654752440 * r);
      }
    }
  }
  *noise_var = nv;
  return nl;
}

// exp(x) for x <= 0 with a 64-entry table of 2^(j/64) in shared memory and a degree-5 polynomial on |r| <= ln2/128:
//   x = (64 m + j) ln2/64 + r,  exp(x) = 2^m 2^(j/64) (1 + r + r^2/2 + ... + r^5/120),  truncation r^6/720 < 3.5e-17.
// 11 fp64-pipe instructions against libdevice's ~17 (degree-11 polynomial) and a handful of integer ones; measured
// max relative error 2.2e-16 against a 200-bit reference over [-700, 0].  Results below 2^-1021 flush to 0.
__device__ __forceinline__ double exp_neg(double x, const double* __restrict__ etab) {
  const double magic = 6755399441055744.0;                        // 1.5 * 2^52: the low word of x*64/ln2 + magic is round(x*64/ln2)
  const double t = fma(x, 92.33248261689366, magic);
  const int n = __double2loint(t);
  const double nd = t - magic;
  double r = fma(nd, -0.01083042469326756, x);                    // ln2/64 split: high part has 32 significant bits, nd * hi is exact
  r = fma(nd, -2.9815858269852933e-12, r);
  double p = fma(r, 1.0 / 120.0, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p *= r;
  const double tj = etab[n & 63];
  const double res = fma(tj, p, tj);
  const int hi = __double2hiint(res) + ((n >> 6) << 20);
  return x < -708.0 ? 0.0 : __hiloint2double(hi, __double2loint(res));
}

// k(d) and (optionally) dk/dd for the metric leaves; d >= 0
template <int OP, bool GRAD>
__device__ __forceinline__ void kfun(double d, double alpha, const double* __restrict__ etab, double& kk, double& dk) {
  if (OP == G3_K_SE || OP == G3_K_OU) {
    kk = exp_neg(-d, etab);
    if (GRAD) dk = -kk;
  } else if (OP == G3_K_MAT32) {
    const double a = 3.0 * d;
    const double s = a > 0.0 ? a * rsqrt(a) : 0.0;
    const double ex = exp_neg(-s, etab);
    kk = fma(s, ex, ex);
    if (GRAD) dk = -1.5 * ex;
  } else if (OP == G3_K_MAT52) {
    const double a = 5.0 * d;
    const double s = a > 0.0 ? a * rsqrt(a) : 0.0;
    const double ex = exp_neg(-s, etab);
    kk = (1.0 + s + a * (1.0 / 3.0)) * ex;
    if (GRAD) dk = -(5.0 / 6.0) * fma(s, ex, ex);
  } else if (OP == G3_K_RQ) {
    const double base = 1.0 + d / alpha;
    kk = pow(base, -alpha);
    if (GRAD) dk = -kk / base;
  } else {  // SIN: d = sum_k rate_k sin^2(pi df f), k = exp(+2 d)
    kk = exp(2.0 * d);
    if (GRAD) dk = 2.0 * kk;
  }
}

// metric pieces of one row against the thread's 4 columns.  xs: the thread's column coordinates already scaled for this
// leaf; xi: the row's scaled coordinates.  dk_[k][e] (per-dimension pieces) only when GRAD.
template <int OP, int DT, bool GRAD>
__device__ __forceinline__ void metric4(const double (&xi)[DT], const double (&xs)[DT][4], const double (&rr)[DT], double (&d)[4],
                                        double (&dk_)[DT][4]) {
#pragma unroll
  for (int e = 0; e < 4; ++e) d[e] = 0.0;
#pragma unroll
  for (int k = 0; k < DT; ++k) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const double df = xi[k] - xs[k][e];
      double piece;
      if (OP == G3_K_OU) piece = fabs(df);
      else if (OP == G3_K_SIN) { const double sn = sinpi(df); piece = rr[k] * sn * sn; }
      else piece = df * df;
      if (OP == G3_K_SE || OP == G3_K_MAT32 || OP == G3_K_MAT52 || OP == G3_K_RQ) d[e] = fma(df, df, d[e]);
      else d[e] += piece;
      if (GRAD) dk_[k][e] = (OP == G3_K_SIN) ? df : piece;       // SIN keeps the scaled difference (freq gradient)
    }
  }
}

constexpr int RC = 2;      // rows per chunk: the chunk's accumulators / weights stay in registers (static indexing)

template <int OP, int DT>
__device__ __forceinline__ void leaf_fwd(const Leaf2& t, const double* __restrict__ x1s, const double (&xc)[DT][4], int row0,
                                         const double* __restrict__ etab, double (&sum)[RC][4]) {
  double sc[DT], rr[DT], xs[DT][4];
#pragma unroll
  for (int k = 0; k < DT; ++k) {
    sc[k] = t.s[k];
    rr[k] = t.r[k];
#pragma unroll
    for (int e = 0; e < 4; ++e) xs[k][e] = __dmul_rn(xc[k][e], sc[k]);   // no FMA contraction with the subtraction below:
  }
  const double var = t.var, alpha = t.alpha;
#pragma unroll
  for (int q = 0; q < RC; ++q) {
    const int rl = row0 + 8 * q;
    double xi[DT], d[4], dummy[DT][4];
#pragma unroll
    for (int k = 0; k < DT; ++k) xi[k] = __dmul_rn(x1s[rl * DT + k], sc[k]);   // df_ij must equal -df_ji to the bit (K symmetric)
    metric4<OP, DT, false>(xi, xs, rr, d, dummy);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      double kk, dk;
      kfun<OP, false>(d[e], alpha, etab, kk, dk);
      sum[q][e] = fma(var, kk, sum[q][e]);
    }
  }
}

// COLD: also instantiate the RQ (pow) and SIN (sinpi) leaves, whose large libdevice bodies would otherwise set the register
// budget of the common exponential / Matern leaves
template <int DT, bool COLD>
__global__ void __launch_bounds__(256, 2)
gram_fwd_fast2_kernel(const __grid_constant__ g3_kernel_desc desc, const GramArgs a, int ntx) {
  extern __shared__ double sm[];
  double* x1s = sm;                       // [128][DT]   centred on X2[0]
  double* x2s = sm + TS * DT;             // [DT][128]
  double* th = x2s + TS * DT;             // [G3_MAX_THETA]
  Leaf2* tab = reinterpret_cast<Leaf2*>(th + G3_MAX_THETA);
  __shared__ int n_leaves;
  __shared__ double noise_var;
  __shared__ double etab[64];
  if (threadIdx.x < 64) etab[threadIdx.x] = exp2((double)threadIdx.x * (1.0 / 64.0));
  const int b = a.bmap ? a.bmap[blockIdx.y] : (int)blockIdx.y;
  int tx_, ty_;
  decode_xy(blockIdx.x, a.lower_only, ntx, tx_, ty_);
  const int r0 = tx_ * TS, c0 = ty_ * TS;
  for (int idx = threadIdx.x; idx < TS * DT; idx += blockDim.x) {
    const int r = idx / DT, k = idx - r * DT;
    const double o = a.X2[k];
    x1s[idx] = (r0 + r < a.n1) ? a.X1[(long long)(r0 + r) * DT + k] - o : 0.0;
    x2s[k * TS + r] = (c0 + r < a.n2) ? a.X2[(long long)(c0 + r) * DT + k] - o : 0.0;
  }
  for (int p = threadIdx.x; p < a.P; p += blockDim.x) th[p] = a.theta[(long long)b * a.P + p];
  __syncthreads();
  if (threadIdx.x == 0) n_leaves = build_leaf2(desc, th, tab, a.skip_process_noise, &noise_var);
  __syncthreads();
  const int nl = n_leaves;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int cc[4] = {2 * tx, 2 * tx + 1, 64 + 2 * tx, 64 + 2 * tx + 1};
  double xc[DT][4];
#pragma unroll
  for (int k = 0; k < DT; ++k)
#pragma unroll
    for (int e = 0; e < 4; ++e) xc[k][e] = x2s[k * TS + cc[e]];
  const double shift = (a.diag_shift && a.same) ? a.diag_shift[b] : 0.0;
  const double nvar = a.same ? noise_var : 0.0;
  double* Kb = a.K + (long long)b * a.strideK;
  int flag = 0;
  for (int ch = 0; ch < 16 / RC; ++ch) {
  const int row0 = ty + 8 * RC * ch;
  double sum[RC][4];
#pragma unroll
  for (int q = 0; q < RC; ++q)
#pragma unroll
    for (int e = 0; e < 4; ++e) sum[q][e] = 0.0;
  for (int l = 0; l < nl; ++l) {
    const Leaf2& t = tab[l];
    switch (t.op) {
      case G3_K_SE: leaf_fwd<G3_K_SE, DT>(t, x1s, xc, row0, etab, sum); break;
      case G3_K_OU: leaf_fwd<G3_K_OU, DT>(t, x1s, xc, row0, etab, sum); break;
      case G3_K_MAT32: leaf_fwd<G3_K_MAT32, DT>(t, x1s, xc, row0, etab, sum); break;
      case G3_K_MAT52: leaf_fwd<G3_K_MAT52, DT>(t, x1s, xc, row0, etab, sum); break;
      case G3_K_RQ: if (COLD) leaf_fwd<G3_K_RQ, DT>(t, x1s, xc, row0, etab, sum); break;
      case G3_K_SIN: if (COLD) leaf_fwd<G3_K_SIN, DT>(t, x1s, xc, row0, etab, sum); break;
      default: break;                      // Noise / WN(same): diagonal only, added below
    }
  }
#pragma unroll
  for (int q = 0; q < RC; ++q) {
    const int gi = r0 + row0 + 8 * q;
    double v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int gj = c0 + cc[e];
      double val = sum[q][e];
      const bool on_diag = a.same && (gi + a.diag_off == gj);
      if (on_diag) val += nvar;
      if (!(fabs(val) <= 1.7976931348623157e308)) { flag = 1; val = isnan(val) ? 0.0 : kInfRepl; }
      if (on_diag) val += shift;
      if (gi >= a.n1 || gj >= a.n2) val = (a.pad_identity && gi + a.diag_off == gj) ? 1.0 : 0.0;
      v[e] = val;
    }
    double* rowp = Kb + (long long)gi * a.ldk + c0;
    *reinterpret_cast<double2*>(rowp + cc[0]) = make_double2(v[0], v[1]);
    *reinterpret_cast<double2*>(rowp + cc[2]) = make_double2(v[2], v[3]);
  }
  }
  if (a.status && flag) atomicOr(a.status + b, G3_ST_NONFINITE_INPUT);
}

// One leaf of the VJP: accumulates sum_ij w_ij dK_ij/dtheta for the leaf's own hypers in registers over the thread's
// 16 x 4 elements and adds them to the thread's slots of the shared accumulator once.
template <int OP, int DT>
__device__ __forceinline__ void leaf_vjp(const Leaf2& t, const double* __restrict__ x1s, const double (&xc)[DT][4], int row0, int tid,
                                         const double* __restrict__ etab, const double (&w)[RC][4], double* __restrict__ acc) {
  double sc[DT], rr[DT], xs[DT][4];
#pragma unroll
  for (int k = 0; k < DT; ++k) {
    sc[k] = t.s[k];
    rr[k] = t.r[k];
#pragma unroll
    for (int e = 0; e < 4; ++e) xs[k][e] = __dmul_rn(xc[k][e], sc[k]);   // no FMA contraction with the subtraction below:
  }
  const double alpha = t.alpha;
  double g_var = 0.0, g_alpha = 0.0, g_r[DT], g_f[DT];
#pragma unroll
  for (int k = 0; k < DT; ++k) g_r[k] = g_f[k] = 0.0;
#pragma unroll
  for (int q = 0; q < RC; ++q) {
    const int rl = row0 + 8 * q;
    double xi[DT], d[4], pc[DT][4];
#pragma unroll
    for (int k = 0; k < DT; ++k) xi[k] = __dmul_rn(x1s[rl * DT + k], sc[k]);   // df_ij must equal -df_ji to the bit (K symmetric)
    metric4<OP, DT, true>(xi, xs, rr, d, pc);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      double kk, dk;
      kfun<OP, true>(d[e], alpha, etab, kk, dk);
      const double we = w[q][e];
      g_var = fma(we, kk, g_var);
      const double gk = we * dk;                       // times var at the end
      if (OP == G3_K_RQ) g_alpha = fma(we * kk, -log1p(d[e] / alpha) + d[e] / (alpha + d[e]), g_alpha);
#pragma unroll
      for (int k = 0; k < DT; ++k) {
        if (OP == G3_K_SIN) {                          // pc = df * f; d/drate_k = 2 K sin^2, d/dfreq_k = 2 K r sin(2 pi df f) pi df
          const double sn = sinpi(pc[k][e]);
          g_r[k] = fma(gk, sn * sn, g_r[k]);
          g_f[k] = fma(gk, sinpi(2.0 * pc[k][e]) * pc[k][e], g_f[k]);
        } else {
          g_r[k] = fma(gk, pc[k][e], g_r[k]);
        }
      }
    }
  }
  const double var = t.var;
  if (t.var_idx >= 0) acc[t.var_idx * 256 + tid] += g_var;
  if (OP == G3_K_RQ) acc[t.p1_idx * 256 + tid] += var * g_alpha;
#pragma unroll
  for (int k = 0; k < DT; ++k) {
    if (k >= t.dim0 && k < t.dim1) {
      if (OP == G3_K_SIN) {
        // gk = w dk = 2 w K/var;  d/drate_k = var * (w K 2 sn^2) = var * gk sn^2;  d/dfreq_k = var gk r_k sin(2 pi u) pi u / f_k, u = df f
        acc[(t.p0_idx + k - t.dim0) * 256 + tid] += var * g_r[k];
        acc[(t.p1_idx + k - t.dim0) * 256 + tid] += var * g_f[k] * rr[k] * M_PI / sc[k];
      } else if (OP == G3_K_OU) {
        acc[(t.p0_idx + k - t.dim0) * 256 + tid] += var * g_r[k] / rr[k];          // pc = r |df|:  d d/dr = |df|
      } else {
        acc[(t.p0_idx + k - t.dim0) * 256 + tid] += var * g_r[k] * (2.0 / rr[k]);  // pc = 0.5 r^2 df^2: d d/dr = 2 pc / r
      }
    }
  }
}

template <int DT, bool COLD>
__global__ void __launch_bounds__(256, 2)
gram_vjp_fast2_kernel(const __grid_constant__ g3_kernel_desc desc, const VjpArgs a, int ntx, double* __restrict__ partials,
                      int ntiles) {
  extern __shared__ double sm[];
  double* x1s = sm;
  double* x2s = sm + TS * DT;
  double* th = x2s + TS * DT;
  double* al_r = th + G3_MAX_THETA;
  double* al_c = al_r + TS;
  Leaf2* tab = reinterpret_cast<Leaf2*>(al_c + TS);
  double* acc = reinterpret_cast<double*>(tab + G3_MAX_NODES);   // [P][256]
  __shared__ int n_leaves;
  __shared__ double noise_var;
  __shared__ double red[8];
  __shared__ double etab[64];
  if (threadIdx.x < 64) etab[threadIdx.x] = exp2((double)threadIdx.x * (1.0 / 64.0));
  const int b = blockIdx.y, tid = threadIdx.x;
  int tx_, ty_;
  decode_xy(blockIdx.x, a.lower_only, ntx, tx_, ty_);
  const int r0 = tx_ * TS, c0 = ty_ * TS;
  for (int idx = tid; idx < TS * DT; idx += blockDim.x) {
    const int r = idx / DT, k = idx - r * DT;
    const double o = a.X2[k];
    x1s[idx] = (r0 + r < a.n1) ? a.X1[(long long)(r0 + r) * DT + k] - o : 0.0;
    x2s[k * TS + r] = (c0 + r < a.n2) ? a.X2[(long long)(c0 + r) * DT + k] - o : 0.0;
  }
  for (int p = tid; p < a.P; p += blockDim.x) th[p] = a.theta[(long long)b * a.P + p];
  if (a.alpha && tid < TS) {
    al_r[tid] = (r0 + tid < a.n1) ? a.alpha[(long long)b * a.strideAlpha + r0 + tid] : 0.0;
    al_c[tid] = (c0 + tid < a.n2) ? (a.alpha2 ? a.alpha2 : a.alpha)[(long long)b * a.strideAlpha + c0 + tid] : 0.0;
  }
  for (int p = 0; p < a.P; ++p) acc[p * 256 + tid] = 0.0;
  __syncthreads();
  if (tid == 0) n_leaves = build_leaf2(desc, th, tab, 0, &noise_var);
  __syncthreads();
  const int nl = n_leaves;
  const double cf = (a.alpha && a.cfac) ? a.cfac[b] : 1.0;
  const int tx = tid & 31, ty = tid >> 5;
  const int cc[4] = {2 * tx, 2 * tx + 1, 64 + 2 * tx, 64 + 2 * tx + 1};
  double xc[DT][4];
#pragma unroll
  for (int k = 0; k < DT; ++k)
#pragma unroll
    for (int e = 0; e < 4; ++e) xc[k][e] = x2s[k * TS + cc[e]];
  const double* Wb = a.W + (long long)b * a.strideW;
  double alc[4] = {0.0, 0.0, 0.0, 0.0};
  if (a.alpha) {
#pragma unroll
    for (int e = 0; e < 4; ++e) alc[e] = al_c[cc[e]];
  }
  for (int ch = 0; ch < 16 / RC; ++ch) {
    const int row0 = ty + 8 * RC * ch;
    // weights of the chunk's RC x 4 elements (all loads in flight before the first use)
    double w[RC][4];
    double wdiag = 0.0;                                   // sum of the weights on the diagonal (Noise / WN var gradient)
    {
      double2 w01[RC], w23[RC];
#pragma unroll
      for (int q = 0; q < RC; ++q) {
        const int gi = r0 + row0 + 8 * q;
        const double* wrow = Wb + (long long)gi * a.ldw + c0;
        w01[q] = make_double2(0.0, 0.0);
        w23[q] = make_double2(0.0, 0.0);
        if (gi < a.n1) {
          w01[q] = *reinterpret_cast<const double2*>(wrow + cc[0]);
          w23[q] = *reinterpret_cast<const double2*>(wrow + cc[2]);
        }
      }
#pragma unroll
      for (int q = 0; q < RC; ++q) {
        const int gi = r0 + row0 + 8 * q;
        const double ar = a.alpha ? cf * al_r[row0 + 8 * q] : 0.0;
        const double wv[4] = {w01[q].x, w01[q].y, w23[q].x, w23[q].y};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int gj = c0 + cc[e];
          double v = a.alpha ? fma(ar, alc[e], -wv[e]) : wv[e];
          double f = 1.0;
          if (a.lower_only) f = (gi > gj) ? 2.0 : (gi == gj ? 1.0 : 0.0);
          if (gi >= a.n1 || gj >= a.n2) f = 0.0;
          v = f == 0.0 ? 0.0 : v * f;
          w[q][e] = v;
          if (a.same && gi == gj) wdiag += v;
        }
      }
    }
    for (int l = 0; l < nl; ++l) {
      const Leaf2& t = tab[l];
      switch (t.op) {
        case G3_K_SE: leaf_vjp<G3_K_SE, DT>(t, x1s, xc, row0, tid, etab, w, acc); break;
        case G3_K_OU: leaf_vjp<G3_K_OU, DT>(t, x1s, xc, row0, tid, etab, w, acc); break;
        case G3_K_MAT32: leaf_vjp<G3_K_MAT32, DT>(t, x1s, xc, row0, tid, etab, w, acc); break;
        case G3_K_MAT52: leaf_vjp<G3_K_MAT52, DT>(t, x1s, xc, row0, tid, etab, w, acc); break;
        case G3_K_RQ: if (COLD) leaf_vjp<G3_K_RQ, DT>(t, x1s, xc, row0, tid, etab, w, acc); break;
        case G3_K_SIN: if (COLD) leaf_vjp<G3_K_SIN, DT>(t, x1s, xc, row0, tid, etab, w, acc); break;
        default:                                           // Noise / WN(same): dK/dvar = I
          if (t.var_idx >= 0) acc[t.var_idx * 256 + tid] += wdiag;
          break;
      }
    }
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  for (int p = 0; p < a.P; ++p) {
    double v = acc[p * 256 + tid];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
      for (int w2 = 0; w2 < 8; ++w2) s += red[w2];
      partials[((long long)b * ntiles + blockIdx.x) * a.P + p] = s;
    }
    __syncthreads();
  }
}

}  // namespace

// fast2 applies to additive trees of metric leaves; WN only in the cov(x) form (its cross form counts equal coordinates)
bool g3_desc_is_fast2(const g3_kernel_desc& d, int same) {
  for (int n = 0; n < d.n_nodes; ++n) {
    const int op = d.nodes[n].op;
    if (op >= G3_K_SUM) {
      if (op != G3_K_SUM) return false;
      continue;
    }
    if (op == G3_K_WN) {
      if (!same) return false;
      continue;
    }
    if (op > G3_K_NOISE) return false;                  // COS / SINC / SM / DOT / BW / VAR: first-generation paths
    if (d.nodes[n].dim1 > 8) return false;
  }
  return true;
}

static bool desc_has_cold(const g3_kernel_desc& d) {
  for (int n = 0; n < d.n_nodes; ++n)
    if (d.nodes[n].op == G3_K_RQ || d.nodes[n].op == G3_K_SIN) return true;
  return false;
}

size_t g3_fast2_fwd_smem(int D) { return sizeof(double) * (2 * TS * D + G3_MAX_THETA) + sizeof(Leaf2) * G3_MAX_NODES; }
size_t g3_fast2_vjp_smem(int D, int P) {
  return sizeof(double) * (2 * TS * D + G3_MAX_THETA + 2 * TS + (size_t)P * 256) + sizeof(Leaf2) * G3_MAX_NODES;
}

template <int DT>
static int fast2_fwd_launch(g3_ctx* ctx, const g3_kernel_desc& desc, const GramArgs& a, int ntx, dim3 grid, cudaStream_t s) {
  const size_t smem = g3_fast2_fwd_smem(DT);
  static bool attr = false;
  if (!attr) {
    G3_CUDA(ctx, cudaFuncSetAttribute(gram_fwd_fast2_kernel<DT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    G3_CUDA(ctx, cudaFuncSetAttribute(gram_fwd_fast2_kernel<DT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    attr = true;
  }
  if (desc_has_cold(desc)) gram_fwd_fast2_kernel<DT, true><<<grid, 256, smem, s>>>(desc, a, ntx);
  else gram_fwd_fast2_kernel<DT, false><<<grid, 256, smem, s>>>(desc, a, ntx);
  return 0;
}

template <int DT>
static int fast2_vjp_launch(g3_ctx* ctx, const g3_kernel_desc& desc, const VjpArgs& a, int ntx, double* partials, int ntiles,
                            dim3 grid, cudaStream_t s) {
  const size_t smem = g3_fast2_vjp_smem(DT, a.P);
  static bool attr = false;
  if (!attr) {
    G3_CUDA(ctx, cudaFuncSetAttribute(gram_vjp_fast2_kernel<DT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    G3_CUDA(ctx, cudaFuncSetAttribute(gram_vjp_fast2_kernel<DT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr = true;
  }
  if (desc_has_cold(desc)) gram_vjp_fast2_kernel<DT, true><<<grid, 256, smem, s>>>(desc, a, ntx, partials, ntiles);
  else gram_vjp_fast2_kernel<DT, false><<<grid, 256, smem, s>>>(desc, a, ntx, partials, ntiles);
  return 0;
}

int g3_gram_fwd_fast2_launch(g3_ctx* ctx, const g3_kernel_desc& desc, const GramArgs& a, int ntx, dim3 grid, cudaStream_t s) {
  switch (a.D) {
    case 1: return fast2_fwd_launch<1>(ctx, desc, a, ntx, grid, s);
    case 2: return fast2_fwd_launch<2>(ctx, desc, a, ntx, grid, s);
    case 3: return fast2_fwd_launch<3>(ctx, desc, a, ntx, grid, s);
    case 4: return fast2_fwd_launch<4>(ctx, desc, a, ntx, grid, s);
    case 5: return fast2_fwd_launch<5>(ctx, desc, a, ntx, grid, s);
    case 6: return fast2_fwd_launch<6>(ctx, desc, a, ntx, grid, s);
    case 7: return fast2_fwd_launch<7>(ctx, desc, a, ntx, grid, s);
    default: return fast2_fwd_launch<8>(ctx, desc, a, ntx, grid, s);
  }
}

int g3_gram_vjp_fast2_launch(g3_ctx* ctx, const g3_kernel_desc& desc, const VjpArgs& a, int ntx, double* partials, int ntiles,
                             dim3 grid, cudaStream_t s) {
  switch (a.D) {
    case 1: return fast2_vjp_launch<1>(ctx, desc, a, ntx, partials, ntiles, grid, s);
    case 2: return fast2_vjp_launch<2>(ctx, desc, a, ntx, partials, ntiles, grid, s);
    case 3: return fast2_vjp_launch<3>(ctx, desc, a, ntx, partials, ntiles, grid, s);
    case 4: return fast2_vjp_launch<4>(ctx, desc, a, ntx, partials, ntiles, grid, s);
    case 5: return fast2_vjp_launch<5>(ctx, desc, a, ntx, partials, ntiles, grid, s);
    case 6: return fast2_vjp_launch<6>(ctx, desc, a, ntx, partials, ntiles, grid, s);
    case 7: return fast2_vjp_launch<7>(ctx, desc, a, ntx, partials, ntiles, grid, s);
    default: return fast2_vjp_launch<8>(ctx, desc, a, ntx, partials, ntiles, grid, s);
  }
}

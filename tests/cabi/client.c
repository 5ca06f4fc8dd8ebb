/* Pure-C client of the drop-in boundary (include/g3b.h): no Python, no C++.  Builds the descriptor of
 * SE(ARD) + Noise (what EllipticalProcess compiles for `g3py.GP(x, Zero(), SE(x))`, elliptical.py:26-28), evaluates
 * beta = |L^-1 y|^2, log-det and the gradient for two hyper samples on a fixed synthetic data set, and prints them
 * with full precision.  tests/test_host.py compiles and links it on CPU; tests/test_gpu_parity.py runs it on the
 * GPU and compares the printed numbers with the ctypes path. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "g3b.h"

int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 300, D = 2, B = 2;
  double* X = (double*)malloc(sizeof(double) * N * D);
  double* y = (double*)malloc(sizeof(double) * N);
  for (int i = 0; i < N; ++i) {               /* deterministic data, no RNG: reproducible from Python */
    X[i * D + 0] = fmod(0.37 * i, 5.0);
    X[i * D + 1] = fmod(0.91 * i + 0.5, 3.0);
    y[i] = sin(X[i * D + 0]) + 0.3 * cos(2.0 * X[i * D + 1]);
  }
  g3_kernel_desc desc;
  memset(&desc, 0, sizeof desc);
  desc.n_nodes = 3;
  desc.n_theta = 4;                           /* natural space: [SE var, SE rate[2], Noise var] */
  desc.nodes[0].op = G3_K_SE;    desc.nodes[0].dim0 = 0; desc.nodes[0].dim1 = D;
  desc.nodes[0].var_idx = 0;     desc.nodes[0].p0_idx = 1; desc.nodes[0].p1_idx = -1;
  desc.nodes[1].op = G3_K_NOISE; desc.nodes[1].var_idx = 3; desc.nodes[1].p0_idx = -1; desc.nodes[1].p1_idx = -1;
  desc.nodes[1].flags = G3_KF_PROCESS_NOISE;
  desc.nodes[2].op = G3_K_SUM;   desc.nodes[2].dim0 = 0; desc.nodes[2].dim1 = 1;
  desc.nodes[2].var_idx = -1;    desc.nodes[2].p0_idx = -1; desc.nodes[2].p1_idx = -1;
  const double theta[2][4] = {{1.0, 0.8, 1.3, 0.05}, {0.6, 1.1, 0.7, 0.1}};
  double beta[2], logdet[2], dtheta[2][4];
  double* ddelta = (double*)malloc(sizeof(double) * B * N);
  int status[2];
  g3_ctx* ctx = NULL;
  int rc = g3_ctx_create(0, &ctx);
  if (rc) { fprintf(stderr, "g3_ctx_create failed: %d\n", rc); return 2; }
  if ((rc = g3_set_data(ctx, X, N, D)) ||
      (rc = g3_gp_logp_grad(ctx, &desc, G3_KIND_GAUSS, y, 0, &theta[0][0], B, NULL, beta, logdet, &dtheta[0][0], ddelta,
                            status))) {
    fprintf(stderr, "call failed: %d: %s\n", rc, g3_last_error(ctx));
    return 3;
  }
  for (int b = 0; b < B; ++b)
    printf("%d %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", b, status[b], beta[b], logdet[b], dtheta[b][0],
           dtheta[b][1], dtheta[b][2], dtheta[b][3], ddelta[b * N + N / 2]);
  g3_ctx_destroy(ctx);
  free(X); free(y); free(ddelta);
  return 0;
}

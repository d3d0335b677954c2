/* A plain-C caller of the ABI in include/nfm.h: cudaMalloc buffers, one
 * nfm_sym_solve + nfm_sym_matvec round trip, result checked on the host.
 * Built and run by tests/test_abi.py (no Python, no torch between this program
 * and libnfm_sm100a.so):
 *   gcc -std=c99 tests/abi_c_call.c -Iinclude -I$CUDA/include -Lnitorch_fastmath_b200 -lnfm_sm100a -L$CUDA/lib64 -lcudart
 * Prints "OK <max error>" and exits 0 on success. */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "nfm.h"

#define CHECK(x)                                                        \
  do {                                                                  \
    int rc_ = (int)(x);                                                 \
    if (rc_ != 0) {                                                     \
      fprintf(stderr, "%s failed: %d (%s)\n", #x, rc_, nfm_last_error_string()); \
      return 1;                                                         \
    }                                                                   \
  } while (0)

int main(void) {
  const int n = 3, nn = 6;
  const int64_t batch = 100003; /* odd: full tiles + a partial tile + a tail */
  float *h_mat = (float *)malloc(sizeof(float) * batch * nn), *h_vec = (float *)malloc(sizeof(float) * batch * n);
  float *h_back = (float *)malloc(sizeof(float) * batch * n);
  unsigned s = 12345u;
  for (int64_t b = 0; b < batch; ++b) {
    for (int k = 0; k < nn; ++k) {
      s = s * 1664525u + 1013904223u;
      float r = (float)(s >> 8) / 16777216.0f; /* [0,1) */
      h_mat[b * nn + k] = k < n ? 4.0f + r : 0.5f * r - 0.25f; /* diagonally dominant: SPD */
    }
    for (int k = 0; k < n; ++k) {
      s = s * 1664525u + 1013904223u;
      h_vec[b * n + k] = (float)(s >> 8) / 8388608.0f - 1.0f;
    }
  }
  float *d_mat, *d_vec, *d_x, *d_back;
  CHECK(cudaMalloc((void **)&d_mat, sizeof(float) * batch * nn));
  CHECK(cudaMalloc((void **)&d_vec, sizeof(float) * batch * n));
  CHECK(cudaMalloc((void **)&d_x, sizeof(float) * batch * n));
  CHECK(cudaMalloc((void **)&d_back, sizeof(float) * batch * n));
  CHECK(cudaMemcpy(d_mat, h_mat, sizeof(float) * batch * nn, cudaMemcpyHostToDevice));
  CHECK(cudaMemcpy(d_vec, h_vec, sizeof(float) * batch * n, cudaMemcpyHostToDevice));
  if (nfm_version() < 100) return 2;
  /* x = A^-1 v, then back = A x, both on the default stream */
  CHECK(nfm_sym_solve(NFM_F32, n, NFM_LAYOUT_SYM, NFM_ALGO_AUTO, batch, d_mat, nn, d_vec, n, NULL, 0, d_x, n, NULL));
  if (nfm_last_path_was_tma() != 1) {
    fprintf(stderr, "TMA fast path not taken\n");
    return 3;
  }
  CHECK(nfm_sym_matvec(NFM_F32, n, NFM_LAYOUT_SYM, batch, d_mat, nn, d_x, n, NULL, 0, 0, d_back, n, NULL));
  CHECK(cudaDeviceSynchronize());
  CHECK(cudaMemcpy(h_back, d_back, sizeof(float) * batch * n, cudaMemcpyDeviceToHost));
  double worst = 0;
  for (int64_t b = 0; b < batch; ++b) {
    double num = 0, den = 0;
    for (int k = 0; k < n; ++k) {
      double e = (double)h_back[b * n + k] - h_vec[b * n + k];
      num += e * e;
      den += (double)h_vec[b * n + k] * h_vec[b * n + k];
    }
    double err = sqrt(num / (den > 1e-30 ? den : 1e-30));
    if (err > worst) worst = err;
  }
  /* error paths are return codes, never exceptions */
  if (nfm_sym_solve(NFM_F32, 11, NFM_LAYOUT_SYM, 0, batch, d_mat, nn, d_vec, n, NULL, 0, d_x, n, NULL) != NFM_E_UNSUPPORTED) return 4;
  if (nfm_sym_solve(NFM_F32, n, NFM_LAYOUT_SYM, 0, batch, NULL, nn, d_vec, n, NULL, 0, d_x, n, NULL) != NFM_E_BADARG) return 5;
  printf("%s %.3e launches %llu\n", worst < 1e-5 ? "OK" : "FAIL", worst, (unsigned long long)nfm_launch_count());
  return worst < 1e-5 ? 0 : 6;
}

// tools/factor_probe.cu — phase timing of k_blk_factor (clock64 stamps) on a random SPD block.
#define FACT_DEBUG 1
#include "../ekf-monoslam_for_3d-reconstruction_b200/csrc/ekf_update.cu"
#include <cstdio>
#include <vector>
#include <cmath>
int main() {
  const int n = EKF_UB;
  std::vector<double> B(n * n), S(n * n, 0.0), nu(n);
  srand(1);
  for (auto& x : B) x = rand() / (double)RAND_MAX - 0.5;
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { double s = 0; for (int k = 0; k < n; ++k) s += B[i * n + k] * B[j * n + k]; S[i * n + j] = s + (i == j ? 4.0 : 0.0); }
  for (auto& x : nu) x = rand() / (double)RAND_MAX;
  double *dS, *dnu, *dL, *dy; DevCtl* ctl;
  cudaMalloc(&dS, n * n * 8); cudaMalloc(&dnu, n * 8); cudaMalloc(&dL, n * n * 8); cudaMalloc(&dy, n * 8); cudaMalloc(&ctl, sizeof(DevCtl));
  cudaMemset(ctl, 0, sizeof(DevCtl));
  cudaMemcpy(dS, S.data(), n * n * 8, cudaMemcpyHostToDevice); cudaMemcpy(dnu, nu.data(), n * 8, cudaMemcpyHostToDevice);
  update_kernels_init();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k_blk_factor<<<1, FACT_THREADS, kFactSmem>>>(dS, dnu, dL, dy, ctl);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long st[32]; cudaMemcpyFromSymbol(st, g_fact_stamp, sizeof st);
    printf("rep %d: %.1f us  err=%s\n", rep, ms * 1e3, cudaGetErrorString(cudaGetLastError()));
    const char* names[17] = {"load", "sync", "diag0", "trsm0", "upd0", "diag1", "trsm1", "upd1", "diag2", "trsm2", "upd2", "diag3", "-", "-", "dinv", "offdiag", "out+y"};
    long long prev = st[0];
    for (int i = 1; i <= 16; ++i) { if (i == 12 || i == 13) continue; printf("  %-8s %8lld cyc\n", names[i], st[i] - prev); prev = st[i]; }
  }
  // check Linv * S * Linv^T = I
  std::vector<double> L(n * n); cudaMemcpy(L.data(), dL, n * n * 8, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  std::vector<double> T(n * n);
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { double s = 0; for (int k = 0; k < n; ++k) s += L[i * n + k] * S[k * n + j]; T[i * n + j] = s; }
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { double s = 0; for (int k = 0; k < n; ++k) s += T[i * n + k] * L[j * n + k]; maxerr = fmax(maxerr, fabs(s - (i == j))); }
  printf("max |Linv S Linv^T - I| = %.3e\n", maxerr);
  return 0;
}

// tools/factor_probe.cu — phase timing of k_blk_factor (clock64 stamps of thread 0) on a random SPD block.
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
  double *dS, *dnu, *dL, *dD, *dy; DevCtl* ctl;
  cudaMalloc(&dS, n * n * 8); cudaMalloc(&dnu, n * 8); cudaMalloc(&dL, n * n * 8); cudaMalloc(&dD, n * 32 * 8); cudaMalloc(&dy, n * 8); cudaMalloc(&ctl, sizeof(DevCtl));
  cudaMemset(ctl, 0, sizeof(DevCtl));
  cudaMemcpy(dS, S.data(), n * n * 8, cudaMemcpyHostToDevice); cudaMemcpy(dnu, nu.data(), n * 8, cudaMemcpyHostToDevice);
  update_kernels_init();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* names[8] = {"load+piv0", "panel", "barrier1", "update/LA", "barrier2", "dinv", "out", "-"};
  for (int rep = 0; rep < 3; ++rep) {
    long long zero[16] = {0}; cudaMemcpyToSymbol(g_fact_acc, zero, sizeof zero);
    cudaEventRecord(e0);
    k_blk_factor<<<1, FACT_THREADS, kFactSmem>>>(dS, dnu, dL, dD, dy, ctl);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long st[16]; cudaMemcpyFromSymbol(st, g_fact_acc, sizeof st);
    printf("rep %d: %.1f us  err=%s\n", rep, ms * 1e3, cudaGetErrorString(cudaGetLastError()));
    for (int i = 0; i < 8; ++i) printf("  %-9s %8lld cyc\n", names[i], st[i]);
  }
  // check L L^T = S
  std::vector<double> L(n * n); cudaMemcpy(L.data(), dL, n * n * 8, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  for (int i = 0; i < n; ++i) for (int j = 0; j <= i; ++j) { double s = 0; for (int k = 0; k <= j; ++k) s += L[i * n + k] * L[j * n + k]; maxerr = fmax(maxerr, fabs(s - S[i * n + j])); }
  printf("max |L L^T - S| = %.3e\n", maxerr);
  return 0;
}

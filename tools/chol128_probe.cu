// tools/chol128_probe.cu — correctness and timing of cta_chol128 (register-resident 128 x 128 Cholesky) against the
// shared-memory kernel cta_chol_panel<128> and a host check, on a random SPD block.
#define CH_DEBUG 1
#include "../ekf-monoslam_for_3d-reconstruction_b200/csrc/ekf_chol128.cuh"
#include <cmath>
#include <cstdio>
#include <vector>
__global__ void __launch_bounds__(CH_THREADS) k_new(const double* S, const double* nu, double* L, double* D, double* y, int* fail) {
  extern __shared__ __align__(16) double sm[];
  cta_chol128(sm, S, 128, nu, L, 128, D, 32, y, fail);
}
__global__ void __launch_bounds__(FACT_THREADS) k_old(const double* S, const double* nu, double* L, double* D, double* y, int* fail) {
  extern __shared__ __align__(16) double sm[];
  cta_chol_panel<128>(sm, S, 128, nu, L, 128, D, 32, y, fail);
}
int main() {
  const int n = 128;
  std::vector<double> B(n * n), S(n * n, 0.0), nu(n);
  srand(1);
  for (auto& x : B) x = rand() / (double)RAND_MAX - 0.5;
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { double s = 0; for (int k = 0; k < n; ++k) s += B[i * n + k] * B[j * n + k]; S[i * n + j] = s + (i == j ? 4.0 : 0.0); }
  for (auto& x : nu) x = rand() / (double)RAND_MAX;
  double *dS, *dnu, *dL, *dD, *dy; int* fail;
  cudaMalloc(&dS, n * n * 8); cudaMalloc(&dnu, n * 8); cudaMalloc(&dL, n * n * 8); cudaMalloc(&dD, n * 32 * 8); cudaMalloc(&dy, n * 8); cudaMalloc(&fail, 4);
  cudaMemset(fail, 0, 4);
  cudaMemcpy(dS, S.data(), n * n * 8, cudaMemcpyHostToDevice); cudaMemcpy(dnu, nu.data(), n * 8, cudaMemcpyHostToDevice);
  const size_t smn = sizeof(Chol128Smem), smo = (size_t)cta_chol_panel_smem_doubles<128>() * 8;
  cudaFuncSetAttribute(k_new, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smn);
  cudaFuncSetAttribute(k_old, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smo);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  std::vector<double> Ln(n * n), Dn(n * 32), yn(n), Lo(n * n), Do(n * 32), yo(n);
  for (int which = 0; which < 2; ++which) {
    cudaMemset(dL, 0, n * n * 8);
    float best = 1e9;
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(e0);
      if (which == 0) k_new<<<1, CH_THREADS, smn>>>(dS, dnu, dL, dD, dy, fail); else k_old<<<1, FACT_THREADS, smo>>>(dS, dnu, dL, dD, dy, fail);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); best = fminf(best, ms);
    }
    printf("%s: %.1f us  (%s)\n", which == 0 ? "cta_chol128 (registers)" : "cta_chol_panel<128> (smem)", best * 1e3, cudaGetErrorString(cudaGetLastError()));
    cudaMemcpy(which == 0 ? Ln.data() : Lo.data(), dL, n * n * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(which == 0 ? Dn.data() : Do.data(), dD, n * 32 * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(which == 0 ? yn.data() : yo.data(), dy, n * 8, cudaMemcpyDeviceToHost);
  }
  {
    long long st[16]; cudaMemcpyFromSymbol(st, g_ch_acc, sizeof st);
    const char* n0[6] = {"load", "wait A", "pivot tile update", "factor+publish", "wait B", "Lout / block copy"};
    const char* n1[6] = {"load", "wait prologue", "dump panel tiles", "wait A (solvers)", "trailing update", "wait end bar"};
    printf("  warp0 inside factor: load+numeric %lld cyc, publish %lld cyc\n", st[6], st[7]);
    for (int i = 0; i < 6; ++i) printf("  warp0 %-28s %8lld cyc   | warp1 %-18s %8lld cyc\n", n0[i], st[i], n1[i], st[8 + i]);
  }
  int hf = 0; cudaMemcpy(&hf, fail, 4, cudaMemcpyDeviceToHost);
  double e_llt = 0, e_L = 0, e_D = 0, e_y = 0, e_res = 0;
  for (int i = 0; i < n; ++i) for (int j = 0; j <= i; ++j) {
    double s = 0; for (int k = 0; k <= j; ++k) s += Ln[i * n + k] * Ln[j * n + k];
    e_llt = fmax(e_llt, fabs(s - S[i * n + j]));
    e_L = fmax(e_L, fabs(Ln[i * n + j] - Lo[i * n + j]));
  }
  for (int i = 0; i < n * 32; ++i) e_D = fmax(e_D, fabs(Dn[i] - Do[i]));
  for (int i = 0; i < n; ++i) { e_y = fmax(e_y, fabs(yn[i] - yo[i])); double s = 0; for (int k = 0; k <= i; ++k) s += Ln[i * n + k] * yn[k]; e_res = fmax(e_res, fabs(s - nu[i])); }
  printf("fail flag %d  max|L L^T - S| %.3e  max|L - L_old| %.3e  max|Dinv - Dinv_old| %.3e  max|y - y_old| %.3e  max|L y - nu| %.3e\n", hf, e_llt, e_L, e_D, e_y, e_res);
  return (e_llt < 1e-11 && e_L < 1e-11 && e_D < 1e-11 && e_y < 1e-11 && !hf) ? 0 : 1;
}

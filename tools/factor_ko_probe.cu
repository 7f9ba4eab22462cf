// tools/factor_ko_probe.cu — attribution of k_blk_factor's time by knocking phases out (results are garbage when a
// phase is skipped; only the time matters).  KO bits: 1 pivot chain, 2 panel solve, 4 rank-4 update, 8 rank-32 update,
// 16 diagonal-block inverses, 32 barriers.
#include <cstdio>
#include <vector>
#include <cmath>
#include <cstdlib>
#include "../ekf-monoslam_for_3d-reconstruction_b200/csrc/ekf_factor.cuh"
#define NBK 128
template <int KO>
__global__ void __launch_bounds__(FACT_THREADS) k_probe(const double* Sb, const double* nu, double* L, double* D, double* y, int* fail) {
  extern __shared__ __align__(16) double fsm[];
  cta_chol_panel<NBK, KO>(fsm, Sb, NBK, nu, L, NBK, D, 32, y, fail);
}
// the same factorisation repeated inside one launch: enough PC samples for ncu's source view (one launch of the real
// kernel is 45 us on one SM, a handful of samples)
__global__ void __launch_bounds__(FACT_THREADS) k_probe_loop(const double* Sb, const double* nu, double* L, double* D, double* y, int* fail, int reps) {
  extern __shared__ __align__(16) double fsm[];
  for (int r = 0; r < reps; ++r) cta_chol_panel<NBK, 0>(fsm, Sb, NBK, nu, L, NBK, D, 32, y, fail);
}
template <int KO>
float run(const double* dS, const double* dnu, double* dL, double* dD, double* dy, int* df) {
  const size_t sm = (size_t)cta_chol_panel_smem_doubles<NBK>() * 8;
  cudaFuncSetAttribute(k_probe<KO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0);
    k_probe<KO><<<1, FACT_THREADS, sm>>>(dS, dnu, dL, dD, dy, df);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  printf("KO=%2d  %.2f us  (%s)\n", KO, best * 1e3, cudaGetErrorString(cudaGetLastError()));
  return best;
}
int main() {
  const int n = NBK;
  std::vector<double> B(n * n), S(n * n, 0.0), nu(n);
  srand(1);
  for (auto& x : B) x = rand() / (double)RAND_MAX - 0.5;
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { double s = 0; for (int k = 0; k < n; ++k) s += B[i * n + k] * B[j * n + k]; S[i * n + j] = s + (i == j ? 4.0 : 0.0); }
  for (auto& x : nu) x = rand() / (double)RAND_MAX;
  double *dS, *dnu, *dL, *dD, *dy; int* df;
  cudaMalloc(&dS, n * n * 8); cudaMalloc(&dnu, n * 8); cudaMalloc(&dL, n * n * 8); cudaMalloc(&dD, n * 32 * 8); cudaMalloc(&dy, n * 8); cudaMalloc(&df, 4);
  cudaMemset(df, 0, 4);
  cudaMemcpy(dS, S.data(), n * n * 8, cudaMemcpyHostToDevice); cudaMemcpy(dnu, nu.data(), n * 8, cudaMemcpyHostToDevice);
  run<0>(dS, dnu, dL, dD, dy, df);
  std::vector<double> L(n * n); cudaMemcpy(L.data(), dL, n * n * 8, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  for (int i = 0; i < n; ++i) for (int j = 0; j <= i; ++j) { double s = 0; for (int k = 0; k <= j; ++k) s += L[i * n + k] * L[j * n + k]; maxerr = fmax(maxerr, fabs(s - S[i * n + j])); }
  printf("max |L L^T - S| = %.3e\n", maxerr);
  run<1>(dS, dnu, dL, dD, dy, df);
  run<2>(dS, dnu, dL, dD, dy, df);
  run<4>(dS, dnu, dL, dD, dy, df);
  run<8>(dS, dnu, dL, dD, dy, df);
  run<16>(dS, dnu, dL, dD, dy, df);
  run<12>(dS, dnu, dL, dD, dy, df);
  run<13>(dS, dnu, dL, dD, dy, df);
  run<15>(dS, dnu, dL, dD, dy, df);
  run<31>(dS, dnu, dL, dD, dy, df);
  run<63>(dS, dnu, dL, dD, dy, df);
  run<32>(dS, dnu, dL, dD, dy, df);
  {
    const size_t sm = (size_t)cta_chol_panel_smem_doubles<NBK>() * 8;
    cudaFuncSetAttribute(k_probe_loop, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    k_probe_loop<<<1, FACT_THREADS, sm>>>(dS, dnu, dL, dD, dy, df, 200);
    printf("loop: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  }
  return 0;
}

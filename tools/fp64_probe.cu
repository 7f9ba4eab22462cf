// tools/fp64_probe.cu — measures the B200's fp64 issue peaks used as roofline denominators:
// DMMA.8x8x4 (mma.sync.m8n8k4.f64) and DFMA/DADD/DMUL throughput with independent chains.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dmma(double* out, int iters) {
  double c[8][2];
  for (int i = 0; i < 8; ++i) { c[i][0] = 0; c[i][1] = 0; }
  double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
__global__ void k_alu(double* out, int iters) {
  double x[8];
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3 + i;
  const double a = 1.0000001, b = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) x[i] = __fma_rn(x[i], a, b);
      if (MODE == 1) x[i] = __dadd_rn(x[i], b);
      if (MODE == 2) x[i] = __dmul_rn(x[i], a);
    }
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int blocks = p.multiProcessorCount * 4, threads = 512, iters = 20000;
  double* out; cudaMalloc(&out, sizeof(double) * blocks * threads);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0); k_dmma<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double fl = (double)blocks * (threads / 32) * iters * 8.0 * 512.0;
    if (rep) printf("{\"probe\":\"dmma884\",\"tflops\":%.2f,\"ms\":%.3f}\n", fl / ms / 1e9, ms);
  }
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0); k_alu<0><<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * threads * iters * 8.0;
    if (rep) printf("{\"probe\":\"dfma\",\"tflops\":%.2f,\"gops\":%.1f}\n", 2 * ops / ms / 1e9, ops / ms / 1e6);
  }
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0); k_alu<1><<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * threads * iters * 8.0;
    if (rep) printf("{\"probe\":\"dadd\",\"gops\":%.1f}\n", ops / ms / 1e6);
  }
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0); k_alu<2><<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * threads * iters * 8.0;
    if (rep) printf("{\"probe\":\"dmul\",\"gops\":%.1f}\n", ops / ms / 1e6);
  }
  printf("{\"sms\":%d,\"clock_khz\":%d,\"l2_mb\":%.1f}\n", p.multiProcessorCount, p.clockRate, p.l2CacheSize / 1048576.0);
  return 0;
}

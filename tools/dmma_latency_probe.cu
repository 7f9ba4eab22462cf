#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__global__ void k(double* out, long long* cyc, int chains) {
  double d0[8], d1[8]; double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
  for (int i = 0; i < 8; ++i) { d0[i] = i; d1[i] = -i; }
  long long t0 = clock64();
  for (int it = 0; it < 256; ++it) {
#pragma unroll
    for (int c = 0; c < 8; ++c) if (c < chains) dmma(d0[c], d1[c], a, b);
  }
  long long t1 = clock64();
  double s = 0; for (int i = 0; i < 8; ++i) s += d0[i] + d1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  double* o; long long* c; cudaMalloc(&o, 8 * 1024 * 1024); cudaMalloc(&c, 8192);
  for (int warps : {1, 4, 16}) for (int chains : {1, 2, 4, 8}) {
    k<<<1, 32 * warps>>>(o, c, chains); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("warps %2d chains %d: %.1f cycles per dependent DMMA step (%.2f cyc per DMMA per warp)\n", warps, chains, h / 256.0, h / 256.0 / chains);
  }
  return 0;
}

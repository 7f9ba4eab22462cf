// tools/idp_rate_probe.cu — issue rate of IDP.4A (DP4A), IMAD, IADD3 / LOP3 and two mixes on a B200 SM: the matcher's tile kernel is
// a stream of independent DP4As, so this is its roofline denominator.  8 independent chains per thread, 32 warps per SM.
// Prints warp-instructions per cycle per SM.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/idp_rate_probe tools/idp_rate_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(unsigned* out, int iters, unsigned a, unsigned b, long long* cyc) {
  unsigned acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x + i;
  unsigned x = a + threadIdx.x, y = b;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) acc[i] = __dp4a(x, y, acc[i]);
        if (MODE == 1) acc[i] = acc[i] * x + y;                       // IMAD
        if (MODE == 2) acc[i] = (acc[i] + x) ^ y;                     // IADD3 / LOP3
        if (MODE == 3) { if (i & 1) acc[i] = __dp4a(x, y, acc[i]); else acc[i] = (acc[i] + x) ^ y; }
        if (MODE == 4) { if (i & 1) acc[i] = __dp4a(x, y, acc[i]); else acc[i] = acc[i] * x + y; }
      }
    }
  }
  const long long t1 = clock64();
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE>
void run(const char* name, unsigned* out, long long* cyc, int per_iter_instr) {
  const int iters = 4096;
  k<MODE><<<148, 1024>>>(out, iters, 0x01020304u, 0x05060708u, cyc);
  cudaDeviceSynchronize();
  long long c;
  cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
  const double winstr = 32.0 * iters * per_iter_instr;   // warp-instructions per SM (32 warps)
  printf("%-28s %8.3f warp-instr / cycle / SM  (%lld cycles)\n", name, winstr / (double)c, c);
}
int main() {
  unsigned* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  run<0>("IDP.4A", out, cyc, 32);
  run<1>("IMAD", out, cyc, 32);
  run<2>("IADD3+LOP3 (2 per step)", out, cyc, 64);
  run<3>("IDP.4A : (IADD3+LOP3) 1:1", out, cyc, 48);
  run<4>("IDP.4A : IMAD 1:1", out, cyc, 32);
  return 0;
}

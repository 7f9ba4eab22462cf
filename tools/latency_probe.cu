// tools/latency_probe.cu — dependent-chain latencies (cycles) of the fp64 / shuffle / shared-memory
// operations the single-CTA factor kernel is built from.  One warp, clock64 around a chain of N ops.
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
template <int MODE>
__global__ void k(double* out, long long* cyc, double seed) {
  __shared__ double sm[64];
  sm[threadIdx.x] = seed + threadIdx.x * 1e-9; sm[threadIdx.x + 32] = 1.0;
  __syncthreads();
  double x = seed + threadIdx.x * 1e-7, y = 1.0000001;
  int idx = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) {
    if (MODE == 0) x = __fma_rn(x, y, 1e-9);
    if (MODE == 1) x = __dadd_rn(x, y);
    if (MODE == 2) x = __dmul_rn(x, y);
    if (MODE == 3) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
    if (MODE == 4) x = rsqrt(x) + 1.5;
    if (MODE == 5) x = sqrt(x) + 1.5;
    if (MODE == 6) x = 1.0 / x + 1.5;
    if (MODE == 7) x = __drcp_rn(x) + 1.5;
    if (MODE == 8) { idx = (int)sm[idx & 31 + 32 * 0] * 0 + ((idx + 1) & 31); x += sm[idx]; }   // LDS -> DADD chain
    if (MODE == 9) x = __fma_rn(sm[(i + threadIdx.x) & 63], y, x);  // independent LDS feeding a DFMA chain
    if (MODE == 10) x = (double)rsqrtf((float)x) + 1.5;
  }
  long long t1 = clock64();
  out[threadIdx.x] = x + idx;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int MODE>
void run(const char* name) {
  double* out; long long* cyc; cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 8);
  k<MODE><<<1, 32>>>(out, cyc, 1.37); k<MODE><<<1, 32>>>(out, cyc, 1.37);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("{\"op\":\"%s\",\"cycles_per_op\":%.1f}\n", name, (double)c / N);
}
int main() {
  run<0>("dfma_dependent"); run<1>("dadd_dependent"); run<2>("dmul_dependent"); run<3>("shfl64_dependent");
  run<4>("rsqrt_f64+dadd"); run<5>("sqrt_f64+dadd"); run<6>("div_f64+dadd"); run<7>("drcp_rn+dadd");
  run<8>("lds64+dadd"); run<9>("dfma_chain_with_lds_operand"); run<10>("rsqrtf_roundtrip+dadd");
  return 0;
}

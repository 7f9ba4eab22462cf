// tools/tma_probe2.cu — which tensor-map shapes does UTMALDG accept?  argv: elem_bytes box_w_bytes box_h rank l2promo
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void k_probe(const __grid_constant__ CUtensorMap tmap, int x0, int y0, int z, unsigned char* out, int bytes, int rank) {
  extern __shared__ __align__(1024) unsigned char sm[];
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(sm + ((bytes + 127) & ~127));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    if (rank == 2)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
                   ::"r"(smem_u32(sm)), "l"(&tmap), "r"(x0), "r"(y0), "r"(smem_u32(bar)) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n"
                   ::"r"(smem_u32(sm)), "l"(&tmap), "r"(x0), "r"(y0), "r"(z), "r"(smem_u32(bar)) : "memory");
  }
  unsigned done = 0;
  while (!done) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(0) : "memory");
  }
  __syncthreads();
  for (int e = threadIdx.x; e < bytes; e += blockDim.x) out[e] = sm[e];
}
typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
  const int es = atoi(argv[1]), bwb = atoi(argv[2]), bh = atoi(argv[3]), rank = atoi(argv[4]), l2 = atoi(argv[5]);
  const int Wb = 640, H = 480, F = 3;
  std::vector<unsigned char> img((size_t)Wb * H * F);
  for (size_t i = 0; i < img.size(); ++i) img[i] = (unsigned char)((i * 2654435761u) >> 13);
  unsigned char *d, *o;
  const int bytes = bwb * bh;
  cudaMalloc(&d, img.size()); cudaMalloc(&o, bytes);
  cudaMemcpy(d, img.data(), img.size(), cudaMemcpyHostToDevice);
  void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
  CUtensorMap tm;
  const CUtensorMapDataType dt = es == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : es == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT32;
  const cuuint64_t gdim[3] = {(cuuint64_t)Wb / es, (cuuint64_t)(rank == 2 ? H * F : H), (cuuint64_t)F};
  const cuuint64_t gstr[2] = {(cuuint64_t)Wb, (cuuint64_t)Wb * H};
  const cuuint32_t box[3] = {(cuuint32_t)(bwb / es), (cuuint32_t)bh, 1}; const cuuint32_t est[3] = {1, 1, 1};
  CUresult r = ((Enc)fn)(&tm, dt, rank, d, gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         (CUtensorMapL2promotion)l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  const int x0b = argc > 6 ? atoi(argv[6]) : 96, y0 = 50, z = rank == 2 ? 0 : 1;
  cudaMemset(o, 0xAA, bytes);
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
  k_probe<<<1, 256, ((bytes + 127) & ~127) + 16>>>(tm, x0b / es, y0, z, o, bytes, rank);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<unsigned char> h(bytes);
  cudaMemcpy(h.data(), o, bytes, cudaMemcpyDeviceToHost);
  int mism = 0;
  for (int y = 0; y < bh; ++y) for (int x = 0; x < bwb; ++x) {
    const int gx = x0b + x, gy = y0 + y;
    const unsigned char ref = (gx < Wb && gy < H) ? img[(size_t)z * Wb * H + (size_t)gy * Wb + gx] : 0;
    mism += h[y * bwb + x] != ref;
  }
  printf("elem %d B, box %d B x %d, rank %d, l2promo %d: encode %d, %s, %d mismatches\n", es, bwb, bh, rank, l2, (int)r, cudaGetErrorString(e), mism);
  return 0;
}

// csrc/ekf_export.cu — the caller-side rows of SURVEY.md §8(f) item 4 that touch filter state:
//   k_points_features   RosVSLAM::getPointsFeatures (RosVSLAMRansac.cpp:340-418): the (psize + 1) x 12
//                       matrix points.txt is written from, gathered on the device in one launch
//   k_rts_epoch         VSlamFilter::rts_epoch (vslamRansac.cpp:423-449): one backward step of the RTS
//                       smoother on the 13 camera states (a 13 x 13 problem: one CTA)
#include "ekf_kernels.h"
#include "ekf_math.cuh"

// One thread per live feature: XYZ features write position * map_scale and the 3 x 3 covariance block,
// row by row (Patch_sigma = block.transpose() copied in Eigen's column-major linear order, R:371-375);
// inverse-depth features leave their row zero (R:357-367).  out is zeroed by the caller.
__global__ void __launch_bounds__(128) k_points_features(const double* __restrict__ Sigma, int ld, const double* __restrict__ mu,
                                                         FeatTab ft, int N, double* __restrict__ out, int rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N || !ft.coding[i]) return;
  const int pos = ft.pos[i], radix = ft.real_index[i];
  if (radix < 0 || radix >= rows) return;
  const double map_scale = mu[13];
  double* o = out + (size_t)radix * 12;
  for (int c = 0; c < 3; ++c) o[c] = mu[pos + c] * map_scale;
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) o[3 + a * 3 + b] = Sigma[(size_t)(pos + a) * ld + pos + b];
}

void launch_points_features(cudaStream_t st, const double* Sigma, int ld, const double* mu, FeatTab ft, int N, double* out, int rows,
                            long long* launches) {
  if (N <= 0) return;
  k_points_features<<<(N + 127) / 128, 128, 0, st>>>(Sigma, ld, mu, ft, N, out, rows);
  *launches += 1;
}

// io: [0,13) MU in/out, [13,182) SIGMA in/out (row-major 13 x 13), [182,195) MU_S, [195,364) SIGMA_S,
//     [364,367) dTspeed, [367,370) dRspeed.  Operation order follows the reference expression by expression
// (products k-ascending; SIGMA_P.inverse() of a dynamic matrix = partial-pivot LU, as Eigen and the oracle do).
#define RTS_N 13
__global__ void __launch_bounds__(256) k_rts_epoch(double* __restrict__ io, DevCfg cfg, double dT, int* __restrict__ singular) {
  __shared__ double F[169], SG[169], T1[169], SP[169], INV[169], K[169], D[169], T2[169];
  __shared__ double MU[13], MUP[13], DM[13];
  const int tid = threadIdx.x;
  if (tid < 13) MU[tid] = io[tid];
  for (int e = tid; e < 169; e += blockDim.x) SG[e] = io[13 + e];
  __syncthreads();
  if (tid == 0) {
    const double drs[3] = {io[367], io[368], io[369]}, dts[3] = {io[364], io[365], io[366]};
    d_system_jacobian(MU, dT, drs, F);
    for (int i = 0; i < 13; ++i) MUP[i] = MU[i];
    d_predict_state(MUP, dts, drs, dT);                       // MU_P (V:440)
  }
  __syncthreads();
  for (int e = tid; e < 169; e += blockDim.x) {               // T1 = F * SIGMA
    const int a = e / 13, b = e % 13;
    double s = 0;
    for (int k = 0; k < 13; ++k) s += F[a * 13 + k] * SG[k * 13 + b];
    T1[e] = s;
  }
  __syncthreads();
  for (int e = tid; e < 169; e += blockDim.x) {               // SIGMA_P = T1 * F^T + Qtot (V:434-437)
    const int a = e / 13, b = e % 13;
    double s = 0, q = 0;
    for (int k = 0; k < 13; ++k) s += T1[a * 13 + k] * F[b * 13 + k];
    for (int k = 0; k < 6; ++k) q += (F[a * 13 + 7 + k] * ((cfg.Vmax[k] / dT) / dT)) * F[b * 13 + 7 + k];
    SP[e] = s + q;
    T2[e] = 0;
  }
  __syncthreads();
  if (tid == 0) {                                             // INV = SIGMA_P^-1 by LU with partial pivoting
    double A[169];
    int piv[13];
    for (int e = 0; e < 169; ++e) A[e] = SP[e];
    for (int i = 0; i < 13; ++i) piv[i] = i;
    bool bad = false;
    for (int k = 0; k < 13; ++k) {
      int p = k;
      double best = fabs(A[k * 13 + k]);
      for (int i = k + 1; i < 13; ++i)
        if (fabs(A[i * 13 + k]) > best) { best = fabs(A[i * 13 + k]); p = i; }
      if (best == 0.0) { bad = true; continue; }
      if (p != k) {
        for (int j = 0; j < 13; ++j) { const double t = A[k * 13 + j]; A[k * 13 + j] = A[p * 13 + j]; A[p * 13 + j] = t; }
        const int t = piv[k]; piv[k] = piv[p]; piv[p] = t;
      }
      for (int i = k + 1; i < 13; ++i) {
        const double l = A[i * 13 + k] / A[k * 13 + k];
        A[i * 13 + k] = l;
        for (int j = k + 1; j < 13; ++j) A[i * 13 + j] -= l * A[k * 13 + j];
      }
    }
    if (bad) *singular = 1;
    for (int c = 0; c < 13; ++c) {                            // solve A x = P e_c
      double x[13];
      for (int i = 0; i < 13; ++i) {
        double s = (piv[i] == c) ? 1.0 : 0.0;
        for (int j = 0; j < i; ++j) s -= A[i * 13 + j] * x[j];
        x[i] = s;
      }
      for (int i = 12; i >= 0; --i) {
        double s = x[i];
        for (int j = i + 1; j < 13; ++j) s -= A[i * 13 + j] * x[j];
        x[i] = s / A[i * 13 + i];
      }
      for (int i = 0; i < 13; ++i) INV[i * 13 + c] = x[i];
    }
  }
  for (int e = tid; e < 169; e += blockDim.x) {               // T1 = SIGMA * F^T ; D = SIGMA_S - SIGMA_P
    const int a = e / 13, b = e % 13;
    double s = 0;
    for (int k = 0; k < 13; ++k) s += SG[a * 13 + k] * F[b * 13 + k];
    T2[e] = s;
    D[e] = io[195 + e] - SP[e];
  }
  if (tid < 13) DM[tid] = io[182 + tid] - MUP[tid];
  __syncthreads();
  for (int e = tid; e < 169; e += blockDim.x) {               // K = (SIGMA F^T) SIGMA_P^-1 (V:442)
    const int a = e / 13, b = e % 13;
    double s = 0;
    for (int k = 0; k < 13; ++k) s += T2[a * 13 + k] * INV[k * 13 + b];
    K[e] = s;
  }
  __syncthreads();
  for (int e = tid; e < 169; e += blockDim.x) {               // T1 = K * D
    const int a = e / 13, b = e % 13;
    double s = 0;
    for (int k = 0; k < 13; ++k) s += K[a * 13 + k] * D[k * 13 + b];
    T1[e] = s;
  }
  __syncthreads();
  for (int e = tid; e < 169; e += blockDim.x) {               // SIGMA += (K D) K^T (V:447)
    const int a = e / 13, b = e % 13;
    double s = 0;
    for (int k = 0; k < 13; ++k) s += T1[a * 13 + k] * K[b * 13 + k];
    io[13 + e] = SG[e] + s;
  }
  if (tid < 13) {                                             // MU += K (MU_S - MU_P) (V:445)
    double s = 0;
    for (int k = 0; k < 13; ++k) s += K[tid * 13 + k] * DM[k];
    io[tid] = MU[tid] + s;
  }
}

void launch_rts_epoch(cudaStream_t st, double* io, const DevCfg& cfg, double dT, int* singular, long long* launches) {
  k_rts_epoch<<<1, 256, 0, st>>>(io, cfg, dT, singular);
  *launches += 1;
}

// csrc/ekf_math.cuh — per-thread fp64 math of the EKF path, written so the
// order of operations inside every expression is the reference's (files are compiled with
// --fmad=false, so nothing here is contracted into FMAs).  Citations: V: = vslamRansac.cpp,
// C: = camModel.cpp.
#pragma once
#include "ekf_common.cuh"

// V:1408-1421 — R(q), w-first, un-normalised form.  R row-major 3x3.
__device__ __forceinline__ void d_quat2rot(const double q[4], double R[9]) {
  const double qr = q[0], qi = q[1], qj = q[2], qk = q[3];
  R[0] = qr * qr + qi * qi - qj * qj - qk * qk; R[1] = -2 * qr * qk + 2 * qi * qj; R[2] = 2 * qr * qj + 2 * qi * qk;
  R[3] = 2 * qr * qk + 2 * qi * qj; R[4] = qr * qr - qi * qi + qj * qj - qk * qk; R[5] = -2 * qr * qi + 2 * qj * qk;
  R[6] = -2 * qr * qj + 2 * qi * qk; R[7] = 2 * qr * qi + 2 * qj * qk; R[8] = qr * qr - qi * qi - qj * qj + qk * qk;
}
// V:1388-1400
__device__ __forceinline__ void d_vec2quat(const double v[3], double q[4]) {
  const double alpha = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  if (alpha != 0) {
    q[0] = cos(alpha / 2);
    const double s = sin(alpha / 2);
    for (int i = 0; i < 3; ++i) q[1 + i] = v[i] * s / alpha;
  } else {
    q[0] = 1; q[1] = 0; q[2] = 0; q[3] = 0;
  }
}
// V:1423-1438 (Y) and V:1441-1455 (Ybar), row-major 4x4
__device__ __forceinline__ void d_yupsilon(const double q[4], double Y[16]) {
  const double r1 = q[0], x1 = q[1], y1 = q[2], z1 = q[3];
  Y[0] = r1; Y[1] = -x1; Y[2] = -y1; Y[3] = -z1;
  Y[4] = x1; Y[5] = r1; Y[6] = -z1; Y[7] = y1;
  Y[8] = y1; Y[9] = z1; Y[10] = r1; Y[11] = -x1;
  Y[12] = z1; Y[13] = -y1; Y[14] = x1; Y[15] = r1;
}
__device__ __forceinline__ void d_yupsilon_c(const double q[4], double Y[16]) {
  const double r1 = q[0], x1 = q[1], y1 = q[2], z1 = q[3];
  Y[0] = r1; Y[1] = -x1; Y[2] = -y1; Y[3] = -z1;
  Y[4] = x1; Y[5] = r1; Y[6] = z1; Y[7] = -y1;
  Y[8] = y1; Y[9] = -z1; Y[10] = r1; Y[11] = x1;
  Y[12] = z1; Y[13] = y1; Y[14] = -x1; Y[15] = r1;
}
// V:1457-1460
__device__ __forceinline__ void d_quat_mul(const double a[4], const double b[4], double o[4]) {
  double Y[16];
  d_yupsilon(a, Y);
  for (int i = 0; i < 4; ++i) {
    double s = 0;
    for (int k = 0; k < 4; ++k) s += Y[i * 4 + k] * b[k];
    o[i] = s;
  }
}
// V:1492-1535 — F (13x13 row-major) = d g / d x at mu13 with w + ctrl.
__device__ inline void d_system_jacobian(const double* mu13, double dT, const double ctrl[3], double F[169]) {
  for (int i = 0; i < 169; ++i) F[i] = 0;
  for (int i = 0; i < 13; ++i) F[i * 13 + i] = 1;
  const double* q = mu13 + 3;
  double wc[3], wdt[3], hq[4];
  for (int i = 0; i < 3; ++i) { wc[i] = mu13[10 + i] + ctrl[i]; wdt[i] = dT * wc[i]; }
  d_vec2quat(wdt, hq);
  double Yc[16];
  d_yupsilon_c(hq, Yc);
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) F[(3 + i) * 13 + 3 + j] = Yc[i * 4 + j];
  // Jacobian_qt_w(q, wc, dT) = Y(q) * t2
  const double n = sqrt(wc[0] * wc[0] + wc[1] * wc[1] + wc[2] * wc[2]);
  const double s = sin(dT * n / 2);
  const double c = cos(dT * n / 2);
  const double Sinc = (n == 0 ? 1.0 : 2 * sin(dT * n / 2) / (dT * n));
  double nw[3] = {0, 0, 0};  // reference leaves n_w uninitialised when n == 0 (V:1523); 0 here
  if (n > 0) for (int i = 0; i < 3; ++i) nw[i] = wc[i] / n;
  double t2[12];
  const double a0 = -dT * 0.5 * s;
  for (int j = 0; j < 3; ++j) t2[j] = a0 * nw[j];
  const double a1 = dT * 0.5;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) t2[(1 + i) * 3 + j] = a1 * (Sinc * (i == j ? 1.0 : 0.0) + ((c - Sinc) * nw[i]) * nw[j]);
  double Y[16];
  d_yupsilon(q, Y);
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 3; ++j) {
      double acc = 0;
      for (int k = 0; k < 4; ++k) acc += Y[i * 4 + k] * t2[k * 3 + j];
      F[(3 + i) * 13 + 10 + j] = acc;
    }
  for (int i = 0; i < 3; ++i) F[i * 13 + 7 + i] = dT * 1.0;
}
// V:1575-1589
__device__ inline void d_predict_state(double* X, const double dv[3], const double dw[3], double dT) {
  double v[3], w[3], wdt[3], hq[4], qn[4];
  for (int i = 0; i < 3; ++i) { v[i] = X[7 + i] + dv[i]; w[i] = X[10 + i] + dw[i]; }
  for (int i = 0; i < 3; ++i) X[i] += dT * v[i];
  for (int i = 0; i < 3; ++i) wdt[i] = dT * w[i];
  d_vec2quat(wdt, hq);
  d_quat_mul(X + 3, hq, qn);
  for (int i = 0; i < 4; ++i) X[3 + i] = qn[i];
  for (int i = 0; i < 3; ++i) { X[7 + i] = v[i]; X[10 + i] = w[i]; }
}
// C:18-47 — 2x2 distortion Jacobian, row-major
__device__ __forceinline__ void d_diff_distort(const CamParams& cm, double hx, double hy, double J[4]) {
  const double r_2 = hx * hx + hy * hy;
  const double Lrn = 1 + cm.k1 * r_2 + cm.k2 * r_2 * r_2 + cm.k3 * r_2 * r_2 * r_2;
  const double f = cm.k1 + 2 * cm.k2 * r_2 + 3 * cm.k3 * r_2 * r_2;
  const double hn[2] = {hx, hy}, hc[2] = {hy, hx}, pv[2] = {cm.p1, cm.p2}, pc[2] = {cm.p2, cm.p1};
  const double jm[4] = {cm.p2 * hx, 0.0, 0.0, cm.p1 * hy};
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) {
      double v = Lrn * (a == b ? 1.0 : 0.0);
      v = v + ((2 * f) * hn[a]) * hn[b];
      v = v + (2 * pv[a]) * hc[b];
      v = v + (2 * pc[a]) * hn[b];
      v = v + 4 * jm[a * 2 + b];
      J[a * 2 + b] = v;
    }
}
// C:113-138
__device__ __forceinline__ void d_cam_project(const CamParams& cm, const double h[3], double hd[2]) {
  const double x = h[0], y = h[1], z = h[2];
  const double x1 = x / z, y1 = y / z;
  const double r_2 = x1 * x1 + y1 * y1;
  const double l = 1 + cm.k1 * r_2 + cm.k2 * r_2 * r_2 + cm.k3 * r_2 * r_2 * r_2;
  const double x2 = x1 * l + 2 * cm.p1 * x1 * y1 + cm.p2 * (r_2 + 2 * x1 * x1);
  const double y2 = y1 * l + 2 * cm.p2 * x1 * y1 + cm.p1 * (r_2 + 2 * y1 * y1);
  hd[0] = cm.fx * x2 + cm.u0;
  hd[1] = cm.fy * y2 + cm.v0;
}
// C:68-111 — J is 2x3 row-major = (K * D) * Jn
__device__ __forceinline__ void d_cam_project_J(const CamParams& cm, const double h[3], double hd[2], double J[6]) {
  const double x = h[0], y = h[1], z = h[2];
  d_cam_project(cm, h, hd);
  const double x1 = x / z, y1 = y / z;
  double Jn[6];
  Jn[0] = 1 / z; Jn[1] = 0; Jn[2] = -x / z / z;
  Jn[3] = 0; Jn[4] = 1 / z; Jn[5] = -y / z / z;
  double D[4], KD[4];
  d_diff_distort(cm, x1, y1, D);
  const double Kp[4] = {cm.fx, 0.0, 0.0, cm.fy};
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) {
      double s = 0;
      for (int k = 0; k < 2; ++k) s += Kp[a * 2 + k] * D[k * 2 + b];
      KD[a * 2 + b] = s;
    }
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 3; ++b) {
      double s = 0;
      for (int k = 0; k < 2; ++k) s += KD[a * 2 + k] * Jn[k * 3 + b];
      J[a * 3 + b] = s;
    }
}
// C:140-192 — hC = (x1, y1, 1), J 3x2 row-major = (U * inv(D)) * diag(1/fx, 1/fy)
__device__ inline void d_cam_unproject_J(const CamParams& cm, const double hd[2], double hC[3], double J[6]) {
  const double x2 = (hd[0] - cm.u0) / cm.fx;
  const double y2 = (hd[1] - cm.v0) / cm.fy;
  double x1 = x2, y1 = y2;
  for (int i = 0; i < 50; ++i) {
    const double r_2 = x1 * x1 + y1 * y1;
    const double l = 1 + cm.k1 * r_2 + cm.k2 * r_2 * r_2 + cm.k3 * r_2 * r_2 * r_2;
    const double dx = 2 * cm.p1 * x1 * y1 + cm.p2 * (r_2 + 2 * x1 * x1);
    const double dy = 2 * cm.p2 * x1 * y1 + cm.p1 * (r_2 + 2 * y1 * y1);
    x1 = (x2 - dx) / l;
    y1 = (y2 - dy) / l;
  }
  hC[0] = x1; hC[1] = y1; hC[2] = 1;
  double D[4], Di[4];
  d_diff_distort(cm, x1, y1, D);
  const double det = D[0] * D[3] - D[2] * D[1];
  const double invdet = 1.0 / det;
  Di[0] = D[3] * invdet; Di[2] = -D[2] * invdet; Di[1] = -D[1] * invdet; Di[3] = D[0] * invdet;
  // U = [1 0; 0 1; 0 0]; (U*Di) rows 0,1 = Di (sums with exact zeros), row 2 = 0
  const double Jp[4] = {1 / cm.fx, 0.0, 0.0, 1 / cm.fy};
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 2; ++b) {
      double s = 0;
      for (int k = 0; k < 2; ++k) {
        const double ud = (a < 2) ? Di[a * 2 + k] : 0.0;
        s += ud * Jp[k * 2 + b];
      }
      J[a * 2 + b] = s;
    }
}
// V:1537-1566 + V:1654-1661 — d(R(q) d)/dq, 3x4 row-major
__device__ inline void d_jac_hW_q(const double q[4], const double d[3], double J[12]) {
  const double q0 = 2 * q[0], qx = 2 * q[1], qy = 2 * q[2], qz = 2 * q[3];
  const double dR[4][9] = {
      {q0, -qz, qy, qz, q0, -qx, -qy, qx, q0},
      {qx, qy, qz, qy, -qx, -q0, qz, q0, -qx},
      {-qy, qx, q0, qx, qy, qz, -q0, qz, -qy},
      {-qz, -q0, qx, q0, -qz, qy, qx, qy, qz}};
  for (int j = 0; j < 4; ++j)
    for (int i = 0; i < 3; ++i) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += dR[j][i * 3 + k] * d[k];
      J[i * 4 + j] = s;
    }
}
// V:1462-1489 — d = rho (x - r) + m(theta, phi); optional J (3x6 row-major)
__device__ inline void d_inverse2xyz(const double f[6], const double r[3], double d[3], double* J) {
  const double theta = f[3], phi = f[4], ro = f[5];
  double st, ct, sp, cp;
  st = sin(theta); ct = cos(theta); sp = sin(phi); cp = cos(phi);
  const double m[3] = {st * cp, -sp, ct * cp};
  if (J) {
    for (int i = 0; i < 18; ++i) J[i] = 0;
    for (int i = 0; i < 3; ++i) J[i * 6 + i] = ro * 1.0;
    J[0 * 6 + 3] = ct * cp; J[1 * 6 + 3] = 0; J[2 * 6 + 3] = -st * cp;
    J[0 * 6 + 4] = -st * sp; J[1 * 6 + 4] = -cp; J[2 * 6 + 4] = -ct * sp;
    for (int i = 0; i < 3; ++i) J[i * 6 + 5] = f[i] - r[i];
  }
  for (int i = 0; i < 3; ++i) d[i] = ro * (f[i] - r[i]) + m[i];
}
// V:508-578 / V:1080-1112 — h and compact H (2 x 13) of one feature.
//   fs: feature state (6 inverse depth / 3 XYZ), r: camera position, qc: conj(q), Rcw = R(qc).
// Compact columns: [0,3) d/dr, [3,7) d/dq, [7,13) d/dfeature (XYZ: [7,10), rest 0).
__device__ inline void d_feature_hH(const CamParams& cm, const double* fs, int coding, const double r[3],
                                    const double qc[4], const double Rcw[9], double hi[2], double Hc[26],
                                    double* hCz) {
  double d[3], hC[3], Jh[6], Jf[18];
  if (!coding) d_inverse2xyz(fs, r, d, Jf);
  else for (int i = 0; i < 3; ++i) d[i] = fs[i] - r[i];
  for (int i = 0; i < 3; ++i) {
    double s = 0;
    for (int k = 0; k < 3; ++k) s += Rcw[i * 3 + k] * d[k];
    hC[i] = s;
  }
  d_cam_project_J(cm, hC, hi, Jh);
  *hCz = hC[2];
  double Jq[12];
  d_jac_hW_q(qc, d, Jq);
  // * d_qbar_q() = diag(1,-1,-1,-1): column sign flips (exact; the reference sums x*1 and x*0 terms)
  for (int i = 0; i < 3; ++i) { Jq[i * 4 + 1] = -Jq[i * 4 + 1]; Jq[i * 4 + 2] = -Jq[i * 4 + 2]; Jq[i * 4 + 3] = -Jq[i * 4 + 3]; }
  for (int i = 0; i < 26; ++i) Hc[i] = 0;
  // JR = Jh * Rcw (2x3), with the leading scalar applied to Jh first where the reference does
  const double sc = coding ? -1.0 : -fs[5];
  double JRs[6], JR[6];
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 3; ++b) {
      double s0 = 0, s1 = 0;
      for (int k = 0; k < 3; ++k) { s0 += (sc * Jh[a * 3 + k]) * Rcw[k * 3 + b]; s1 += Jh[a * 3 + k] * Rcw[k * 3 + b]; }
      JRs[a * 3 + b] = s0; JR[a * 3 + b] = s1;
    }
  for (int a = 0; a < 2; ++a) {
    // XYZ: Hit.middleCols<3>(0) = -J*R is -(J*R) (unary minus on the product, V:568)
    for (int b = 0; b < 3; ++b) Hc[a * 13 + b] = coding ? -JR[a * 3 + b] : JRs[a * 3 + b];
    for (int b = 0; b < 4; ++b) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += Jh[a * 3 + k] * Jq[k * 4 + b];
      Hc[a * 13 + 3 + b] = s;
    }
    if (!coding) {
      for (int b = 0; b < 6; ++b) {
        double s = 0;
        for (int k = 0; k < 3; ++k) s += JR[a * 3 + k] * Jf[k * 6 + b];
        Hc[a * 13 + 7 + b] = s;
      }
    } else {
      for (int b = 0; b < 3; ++b) Hc[a * 13 + 7 + b] = JR[a * 3 + b];
    }
  }
}
// Reprojection without Jacobian (RANSAC consensus test, V:1010-1020).
__device__ inline void d_feature_h(const CamParams& cm, const double* fs, int coding, const double r[3],
                                   const double Rcw[9], double hi[2]) {
  double d[3], hC[3];
  if (!coding) d_inverse2xyz(fs, r, d, nullptr);
  else for (int i = 0; i < 3; ++i) d[i] = fs[i] - r[i];
  for (int i = 0; i < 3; ++i) {
    double s = 0;
    for (int k = 0; k < 3; ++k) s += Rcw[i * 3 + k] * d[k];
    hC[i] = s;
  }
  d_cam_project(cm, hC, hi);
}
// V:1599-1623 — rows 3 (theta) and 4 (phi) of the 6x3 d f / d hW; the other rows are zero.
__device__ inline void d_jac_f_hW(const double hW[3], double Jt[3], double Jp[3]) {
  const double hx = hW[0], hy = hW[1], hz = hW[2];
  const double normal = hx * hx + hz * hz;
  const double normal2 = hx * hx + hy * hy + hz * hz;
  Jt[0] = hz / normal; Jt[1] = 0; Jt[2] = -hx / normal;
  Jp[0] = hx * hy / sqrt(normal) / normal2;
  Jp[1] = -sqrt(normal) / normal2;
  Jp[2] = hz * hy / sqrt(normal) / normal2;
}
// Dynamic 2x2 inverse() = partial-pivot LU solve against I (see oracle inverse_pplu; V:995, P:223)
__host__ __device__ __forceinline__ void d_inv2_pplu(const double A[4], double X[4]) {
  double a00 = A[0], a01 = A[1], a10 = A[2], a11 = A[3];
  int swapped = 0;
  if (fabs(a10) > fabs(a00)) {
    double t = a00; a00 = a10; a10 = t;
    t = a01; a01 = a11; a11 = t;
    swapped = 1;
  }
  const double l = a10 / a00;
  const double u11 = a11 - l * a01;
  // rows of P*I: row0 = e_{perm0}, row1 = e_{perm1}
  double x0[2], x1[2];
  x0[0] = swapped ? 0.0 : 1.0; x0[1] = swapped ? 1.0 : 0.0;
  x1[0] = swapped ? 1.0 : 0.0; x1[1] = swapped ? 0.0 : 1.0;
  for (int j = 0; j < 2; ++j) if (l != 0.0) x1[j] -= l * x0[j];
  for (int j = 0; j < 2; ++j) x1[j] = x1[j] / u11;
  for (int j = 0; j < 2; ++j) { x0[j] -= a01 * x1[j]; x0[j] = x0[j] / a00; }
  X[0] = x0[0]; X[1] = x0[1]; X[2] = x1[0]; X[3] = x1[1];
}

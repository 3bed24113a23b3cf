// kalman.cuh — the 7-state constant-velocity box filter of the reference, per thread, FP64.
//
// Replaces KalmanBoxTracker.__init__/predict/update (tracking/sort/sort.py:88-178) and the
// filterpy KalmanFilter.predict/update they call (third-party, restated in SURVEY.md
// Appendix B).  The operation ORDER is part of the contract: it is the order NumPy/OpenBLAS
// executes for the reference's dot()/inv() calls (k-sequential FMA dgemm, (p0+p2)+(p1+p3)
// dgemv, left-looking LU with reciprocal pivots and FMA triangular solves behind
// numpy.linalg.inv), so filter states are bit-identical to the CPU path, not just close.
// All loops have compile-time bounds and indices so x, P and the temporaries stay in
// registers.  Requires -fmad=false (see common.cuh).
#pragma once

#include "common.cuh"

namespace w2t {

// sort.py:111-115 -> Q = diag(2,2,1,25,4,4,5);  sort.py:127 -> R = diag(1,1,10,10)
__device__ __forceinline__ double q_diag(int i) {
  return i == 0 ? 2. : i == 1 ? 2. : i == 2 ? 1. : i == 3 ? 25. : i == 6 ? 5. : 4.;
}
__device__ __forceinline__ double r_diag(int i) { return i < 2 ? 1. : 10.; }

// convert_bbox_to_z, sort.py:50-62, on a float32 row: every component stays float32 (NEP 50).
__device__ __forceinline__ void bbox_to_z(const float (&d)[4], float (&z)[4]) {
  const float w = d[2] - d[0];
  const float h = d[3] - d[1];
  z[0] = d[0] + w / 2.0f;
  z[1] = d[1] + h / 2.0f;
  z[2] = w * h;
  z[3] = w / h;
}

// convert_bbox_to_z (sort.py:50-62) of a float32 row under either NumPy promotion regime, widened to
// the float64 the filter then works in (np.array([x, y, s, r]), sort.py:62 -> kf.update / kf.x[:4]):
//   legacy (NumPy 1.x value-based casting, the reference's pinned environment: python 3.7,
//     environment.yml:7): w, h, s = w*h are float32 scalars, but float32-scalar (op) python-float is
//     float64, so x = bbox[0] + w/2. , y and r = w/float(h) are computed in float64;
//   NEP 50 (NumPy 2): the python floats adopt float32, all four components are float32.
__device__ __forceinline__ void bbox_to_z_d(const float d0, const float d1, const float d2, const float d3,
                                            const bool nep50, double (&z)[4]) {
  const float w = d2 - d0;
  const float h = d3 - d1;
  if (nep50) {
    z[0] = (double)(d0 + w / 2.0f);
    z[1] = (double)(d1 + h / 2.0f);
    z[3] = (double)(w / h);
  } else {
    z[0] = (double)d0 + (double)w / 2.0;
    z[1] = (double)d1 + (double)h / 2.0;
    z[3] = (double)w / (double)h;
  }
  z[2] = (double)(w * h);
}

// convert_x_to_bbox, sort.py:65-75.
__device__ __forceinline__ void x_to_bbox(const double (&x)[7], double (&b)[4]) {
  const double w = sqrt(x[2] * x[3]);
  const double h = x[2] / w;
  b[0] = x[0] - w / 2.;
  b[1] = x[1] - h / 2.;
  b[2] = x[0] + w / 2.;
  b[3] = x[1] + h / 2.;
}

// sort.py:97-137
__device__ __forceinline__ void kf_init(const float (&det)[4], double (&x)[7], double (&P)[49], const bool nep50 = true) {
  double z[4];
  bbox_to_z_d(det[0], det[1], det[2], det[3], nep50, z);
#pragma unroll
  for (int i = 0; i < 4; i++) x[i] = z[i];
  x[4] = x[5] = x[6] = 0.;
#pragma unroll
  for (int i = 0; i < 49; i++) P[i] = 0.;
#pragma unroll
  for (int i = 0; i < 4; i++) P[i * 7 + i] = 10.;
#pragma unroll
  for (int i = 4; i < 7; i++) P[i * 7 + i] = 10000.;
}

// sort.py:170-172: x = Fx, P = F P F' + Q with the two-stage association of dot(dot(F,P),F.T).
__device__ __forceinline__ void kf_predict(double (&x)[7], double (&P)[49]) {
  if (x[6] + x[2] <= 0) x[6] *= 0.0;
  x[0] = x[0] + x[4];
  x[1] = x[1] + x[5];
  x[2] = x[2] + x[6];
  // A = F P (rows 0..2 pick up rows 4..6), in place: rows 4..6 are not modified
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 7; j++) P[i * 7 + j] = P[i * 7 + j] + P[(i + 4) * 7 + j];
  // B = A F' (columns 0..2 pick up columns 4..6), in place; then + Q
#pragma unroll
  for (int i = 0; i < 7; i++) {
#pragma unroll
    for (int j = 0; j < 3; j++) P[i * 7 + j] = P[i * 7 + j] + P[i * 7 + j + 4];
#pragma unroll
    for (int j = 0; j < 7; j++) P[i * 7 + j] = P[i * 7 + j] + ((i == j) ? q_diag(i) : 0.0);
  }
}

// numpy.linalg.inv of a 4x4 in OpenBLAS dgesv order.  a is column-major: a[i + 4*j].
__device__ __forceinline__ void inv4_lapack(double (&a)[16], double (&b)[16]) {
#define A_(i, j) a[(i) + 4 * (j)]
#define B_(i, j) b[(i) + 4 * (j)]
  int piv[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    // apply earlier row interchanges to column j
#pragma unroll
    for (int i = 0; i < j; i++) {
#pragma unroll
      for (int r = i + 1; r < 4; r++)
        if (piv[i] == r) { const double t = A_(i, j); A_(i, j) = A_(r, j); A_(r, j) = t; }
    }
    // U part of the column: forward substitution with dot products (FMA from 0, then subtract)
#pragma unroll
    for (int i = 1; i < j; i++) {
      double t = 0.;
#pragma unroll
      for (int k = 0; k < i; k++) t = fma(A_(i, k), A_(k, j), t);
      A_(i, j) = A_(i, j) - t;
    }
    // L part: b[j:] -= A[j:, :j] b[:j]
    if (j > 0) {
#pragma unroll
      for (int r = j; r < 4; r++) {
        double t = 0.;
#pragma unroll
        for (int k = 0; k < j; k++) t = fma(A_(r, k), A_(k, j), t);
        A_(r, j) = A_(r, j) - t;
      }
    }
    // partial pivoting: first maximum of |.| on or below the diagonal
    int jp = j;
    double mx = fabs(A_(j, j));
#pragma unroll
    for (int r = j + 1; r < 4; r++)
      if (fabs(A_(r, j)) > mx) { mx = fabs(A_(r, j)); jp = r; }
    piv[j] = jp;
#pragma unroll
    for (int r = j + 1; r < 4; r++)
      if (jp == r) {
#pragma unroll
        for (int c = 0; c <= j; c++) { const double t = A_(j, c); A_(j, c) = A_(r, c); A_(r, c) = t; }
      }
    const double rp = 1.0 / A_(j, j);
#pragma unroll
    for (int r = j + 1; r < 4; r++) A_(r, j) = A_(r, j) * rp;
  }
  // B = P * I
#pragma unroll
  for (int i = 0; i < 16; i++) b[i] = 0.;
#pragma unroll
  for (int i = 0; i < 4; i++) B_(i, i) = 1.0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
#pragma unroll
    for (int r = i + 1; r < 4; r++)
      if (piv[i] == r) {
#pragma unroll
        for (int c = 0; c < 4; c++) { const double t = B_(i, c); B_(i, c) = B_(r, c); B_(r, c) = t; }
      }
  }
  // L y = b (unit lower), U x = y; right-looking FMA updates, reciprocal diagonal
#pragma unroll
  for (int c = 0; c < 4; c++) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const double bb = B_(i, c);
#pragma unroll
      for (int k = i + 1; k < 4; k++) B_(k, c) = fma(-bb, A_(k, i), B_(k, c));
    }
#pragma unroll
    for (int i = 3; i >= 0; i--) {
      const double bb = B_(i, c) * (1.0 / A_(i, i));
      B_(i, c) = bb;
#pragma unroll
      for (int k = 0; k < i; k++) B_(k, c) = fma(-bb, A_(k, i), B_(k, c));
    }
  }
#undef A_
#undef B_
}

// sort.py:164 -> filterpy update (Joseph form).
__device__ __forceinline__ void kf_update(double (&x)[7], double (&P)[49], const float (&det)[4], const bool nep50 = true) {
  double zf[4];
  bbox_to_z_d(det[0], det[1], det[2], det[3], nep50, zf);
  double y[4];
#pragma unroll
  for (int i = 0; i < 4; i++) y[i] = zf[i] - x[i];
  // S = P[:4,:4] + R, column-major for the LAPACK-order inverse
  double S[16], SI[16];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) S[i + 4 * j] = P[i * 7 + j] + ((i == j) ? r_diag(i) : 0.0);
  inv4_lapack(S, SI);  // SI column-major: SI[k + 4*j] = inv(S)[k][j]
  // K = P[:, :4] SI
  double K[28];
#pragma unroll
  for (int i = 0; i < 7; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      double acc = 0.;
#pragma unroll
      for (int k = 0; k < 4; k++) acc = fma(P[i * 7 + k], SI[k + 4 * j], acc);
      K[i * 4 + j] = acc;
    }
  // x = x + K y
#pragma unroll
  for (int i = 0; i < 7; i++) {
    const double p0 = K[i * 4 + 0] * y[0], p1 = K[i * 4 + 1] * y[1];
    const double p2 = K[i * 4 + 2] * y[2], p3 = K[i * 4 + 3] * y[3];
    x[i] = x[i] + ((p0 + p2) + (p1 + p3));
  }
  // A = (I - K H)[:, :4]
  double A[28];
#pragma unroll
  for (int i = 0; i < 7; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) A[i * 4 + j] = ((i == j) ? 1.0 : 0.0) - K[i * 4 + j];
  // M = (I-KH) P ; rows 4..6 only need their own old row, rows 0..3 need old rows 0..3
  double M[49];
#pragma unroll
  for (int i = 0; i < 7; i++)
#pragma unroll
    for (int j = 0; j < 7; j++) {
      double acc = 0.;
#pragma unroll
      for (int k = 0; k < 4; k++) acc = fma(A[i * 4 + k], P[k * 7 + j], acc);
      if (i >= 4) acc = acc + P[i * 7 + j];
      M[i * 7 + j] = acc;
    }
  // P = M (I-KH)' + (K R) K'
#pragma unroll
  for (int i = 0; i < 7; i++)
#pragma unroll
    for (int j = 0; j < 7; j++) {
      double n = 0.;
#pragma unroll
      for (int k = 0; k < 4; k++) n = fma(M[i * 7 + k], A[j * 4 + k], n);
      if (j >= 4) n = n + M[i * 7 + j];
      double acc = 0.;
#pragma unroll
      for (int k = 0; k < 4; k++) acc = fma(K[i * 4 + k] * r_diag(k), K[j * 4 + k], acc);
      P[i * 7 + j] = n + acc;
    }
}

// ---------------------------------------------------------------------------------------------
// Block form used by the tracker kernel.
//
// With the reference's constants — F couples only (x,vx), (y,vy), (s,vs) (sort.py:104-106), H
// selects x,y,s,r (sort.py:120-121), Q, R and P0 are diagonal (sort.py:111-134) — every entry
// of P outside the three 2x2 blocks {a, a+4} (a = 0,1,2) and P[3][3] is an exact zero, and it
// stays an exact zero through the FMA chains above (fma(0, v, acc) == acc, 0 + 0 == 0; checked
// against NumPy/OpenBLAS and the C oracle).  Dropping those terms from the sums leaves, for
// every remaining entry, exactly the operations listed above in exactly the same order, so
// the 13 live entries are bit-identical to the 49-entry computation at ~1/20 of the work:
//   S is diagonal -> LAPACK's LU does no pivoting and inv(S) = diag(fl(1/S_kk));
//   K[i][k] = fl(P[i][k] * SI_k);  x_i += fl(K[i][k] * y_k);  A = I - K;
//   M = A P (+P), N = M A' (+M), P = N + (K R) K'  restricted to each block.
// p[4a+0] = P[a][a], p[4a+1] = P[a][a+4], p[4a+2] = P[a+4][a], p[4a+3] = P[a+4][a+4], p[12] = P[3][3].
constexpr int kBlockP = 13;

__device__ __forceinline__ void kfb_init_z(const double (&z)[4], double (&x)[7], double (&p)[kBlockP]) {
#pragma unroll
  for (int i = 0; i < 4; i++) x[i] = z[i];
  x[4] = x[5] = x[6] = 0.;
#pragma unroll
  for (int a = 0; a < 3; a++) {
    p[4 * a + 0] = 10.;
    p[4 * a + 1] = 0.;
    p[4 * a + 2] = 0.;
    p[4 * a + 3] = 10000.;
  }
  p[12] = 10.;
}

__device__ __forceinline__ void kfb_init(const float (&det)[4], double (&x)[7], double (&p)[kBlockP], const bool nep50 = true) {
  double z[4];
  bbox_to_z_d(det[0], det[1], det[2], det[3], nep50, z);
  kfb_init_z(z, x, p);
}

__device__ __forceinline__ void kfb_predict(double (&x)[7], double (&p)[kBlockP]) {
  if (x[6] + x[2] <= 0) x[6] *= 0.0;
#pragma unroll
  for (int a = 0; a < 3; a++) {
    x[a] = x[a] + x[a + 4];
    // F P : row a += row a+4
    const double aa = p[4 * a + 0] + p[4 * a + 2];
    const double ab = p[4 * a + 1] + p[4 * a + 3];
    // (F P) F' : column a += column a+4, then + Q
    p[4 * a + 0] = (aa + ab) + q_diag(a);
    p[4 * a + 1] = ab + 0.0;
    p[4 * a + 2] = (p[4 * a + 2] + p[4 * a + 3]) + 0.0;
    p[4 * a + 3] = p[4 * a + 3] + q_diag(a + 4);
  }
  p[12] = p[12] + q_diag(3);
}

__device__ __forceinline__ void kfb_update_z(double (&x)[7], double (&p)[kBlockP], const double (&zf)[4]) {
#pragma unroll
  for (int a = 0; a < 3; a++) {
    const double y = zf[a] - x[a];
    const double P00 = p[4 * a + 0], P01 = p[4 * a + 1], P10 = p[4 * a + 2], P11 = p[4 * a + 3];
    const double si = 1.0 * (1.0 / (P00 + r_diag(a)));
    const double k0 = P00 * si, k1 = P10 * si;
    x[a] = x[a] + k0 * y;
    x[a + 4] = x[a + 4] + k1 * y;
    const double a0 = 1.0 - k0, a1 = 0.0 - k1;
    const double m00 = a0 * P00, m01 = a0 * P01;
    const double m10 = a1 * P00 + P10, m11 = a1 * P01 + P11;
    const double kr0 = k0 * r_diag(a), kr1 = k1 * r_diag(a);
    p[4 * a + 0] = (m00 * a0) + (kr0 * k0);
    p[4 * a + 1] = ((m00 * a1) + m01) + (kr0 * k1);
    p[4 * a + 2] = (m10 * a0) + (kr1 * k0);
    p[4 * a + 3] = ((m10 * a1) + m11) + (kr1 * k1);
  }
  {
    const double y = zf[3] - x[3];
    const double P33 = p[12];
    const double si = 1.0 * (1.0 / (P33 + r_diag(3)));
    const double k = P33 * si;
    x[3] = x[3] + k * y;
    const double a3 = 1.0 - k;
    const double m = a3 * P33;
    p[12] = (m * a3) + ((k * r_diag(3)) * k);
  }
}

__device__ __forceinline__ void kfb_update(double (&x)[7], double (&p)[kBlockP], const float (&det)[4], const bool nep50 = true) {
  double z[4];
  bbox_to_z_d(det[0], det[1], det[2], det[3], nep50, z);
  kfb_update_z(x, p, z);
}

// scatter / gather between the block form and the dense 7x7 (unit entry points, debug dumps)
__device__ __forceinline__ void kfb_to_dense(const double (&p)[kBlockP], double (&P)[49]) {
#pragma unroll
  for (int i = 0; i < 49; i++) P[i] = 0.;
#pragma unroll
  for (int a = 0; a < 3; a++) {
    P[a * 7 + a] = p[4 * a + 0];
    P[a * 7 + a + 4] = p[4 * a + 1];
    P[(a + 4) * 7 + a] = p[4 * a + 2];
    P[(a + 4) * 7 + a + 4] = p[4 * a + 3];
  }
  P[3 * 7 + 3] = p[12];
}

__device__ __forceinline__ bool kfb_from_dense(const double (&P)[49], double (&p)[kBlockP]) {
  bool structured = true;
#pragma unroll
  for (int i = 0; i < 7; i++)
#pragma unroll
    for (int j = 0; j < 7; j++) {
      const bool live = (i == j) || (i < 3 && j == i + 4) || (j < 3 && i == j + 4);
      if (!live && P[i * 7 + j] != 0.) structured = false;
    }
#pragma unroll
  for (int a = 0; a < 3; a++) {
    p[4 * a + 0] = P[a * 7 + a];
    p[4 * a + 1] = P[a * 7 + a + 4];
    p[4 * a + 2] = P[(a + 4) * 7 + a];
    p[4 * a + 3] = P[(a + 4) * 7 + a + 4];
  }
  p[12] = P[3 * 7 + 3];
  return structured;
}

}  // namespace w2t

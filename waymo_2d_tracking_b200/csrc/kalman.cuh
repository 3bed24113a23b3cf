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
__device__ __forceinline__ void kf_init(const float (&det)[4], double (&x)[7], double (&P)[49]) {
  float z[4];
  bbox_to_z(det, z);
#pragma unroll
  for (int i = 0; i < 4; i++) x[i] = (double)z[i];
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
__device__ __forceinline__ void kf_update(double (&x)[7], double (&P)[49], const float (&det)[4]) {
  float zf[4];
  bbox_to_z(det, zf);
  double y[4];
#pragma unroll
  for (int i = 0; i < 4; i++) y[i] = (double)zf[i] - x[i];
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

}  // namespace w2t

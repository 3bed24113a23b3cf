// units.cu — building-block entry points (unit parity against the oracle).
//
// The same device functions the persistent SORT kernel uses (kalman.cuh, munkres.cuh,
// iou_pair) exposed one operation at a time.
#include "sort_kernel.cuh"

using namespace w2t;

namespace {

__global__ void iou_matrix_kernel(const float4 *dets, int D, const double *trks, int T, float *out) {
  const size_t total = (size_t)D * T;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int d = (int)(i / T), t = (int)(i % T);
    const double *b = trks + 4 * (size_t)t;
    out[i] = iou_pair(dets[d], b[0], b[1], b[2], b[3]);
  }
}

struct LapLayout { size_t C, Z, rstar, cstar, rprime, ucols, crows, total; };

__host__ __device__ inline LapLayout lap_layout(int D, int T) {
  const size_t n = D < T ? D : T, m = D < T ? T : D;
  LapLayout L;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o = align_up(o + bytes, 16); return at; };
  L.C = take(4 * n * (size_t)munkres_pitch((int)m));
  L.Z = take(4 * n * (size_t)munkres_zstride((int)m));
  L.rstar = take(4 * n);
  L.cstar = take(4 * m);
  L.rprime = take(4 * n);
  L.ucols = take(4 * m);
  L.crows = take(4 * n);
  L.total = align_up(o + 16, 256);
  return L;
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) lap_kernel(const float *cost, int D, int T, int32_t *pairs, int32_t *n_pairs,
                                                   char *ws) {
  __shared__ MunkresShared ms;
  __shared__ __align__(16) float s_C[kSmemC];
  __shared__ __align__(16) uint32_t s_Z[kSmemZ];
  __shared__ int s_rstar[kSmemN], s_rprime[kSmemN], s_crows[kSmemN], s_cstar[kSmemM], s_ucols[kSmemM];
  const bool flipped = T < D;
  const int n = flipped ? T : D, m = flipped ? D : T;
  const LapLayout L = lap_layout(D, T);
  Munkres<BLOCK> mk;
  mk.s = &ms;
  mk.ph = nullptr;
  mk.t_last = 0;
  mk.rowwise = false;
  mk.n = n; mk.m = m; mk.mw = munkres_words(m); mk.zs = munkres_zstride(m); mk.ldc = munkres_pitch(m);
  mk.g.C = reinterpret_cast<float *>(ws + L.C);
  mk.g.Z = reinterpret_cast<uint32_t *>(ws + L.Z);
  mk.g.row_star = reinterpret_cast<int *>(ws + L.rstar);
  mk.g.col_star = reinterpret_cast<int *>(ws + L.cstar);
  mk.g.row_prime = reinterpret_cast<int *>(ws + L.rprime);
  mk.g.ucols = reinterpret_cast<int *>(ws + L.ucols);
  mk.g.crows = reinterpret_cast<int *>(ws + L.crows);
  // same dispatch as the tracker kernel: up to 128 x 128 the solver's masks, stars and index lists
  // live in shared memory, and so does the cost matrix if it is small enough (else it stays in the
  // global workspace and the same solver runs on it); anything larger takes the global path
  if ((n * mk.zs <= kSmemZ) && (n <= kSmemN) && (m <= 128)) {
    mk.rowwise = true;
    mk.g.Z = s_Z; mk.g.row_star = s_rstar; mk.g.col_star = s_cstar; mk.g.row_prime = s_rprime;
    mk.g.ucols = s_ucols; mk.g.crows = s_crows;
    if (n * mk.ldc <= kSmemC) mk.g.C = s_C;
  }
  for (size_t i = threadIdx.x; i < (size_t)n * m; i += BLOCK) {
    const int r = (int)(i / m), c = (int)(i % m);
    mk.g.C[(size_t)r * mk.ldc + c] = flipped ? cost[(size_t)c * T + r] : cost[(size_t)r * T + c];
  }
  __syncthreads();
  const int act = mk.solve();
  if (threadIdx.x == 0) {
    int k = 0;
    if (act == 0) {
      if (!flipped) {
        for (int r = 0; r < n; r++) { pairs[2 * k] = r; pairs[2 * k + 1] = mk.g.row_star[r]; k++; }
      } else {
        for (int c = 0; c < m; c++)
          if (mk.g.col_star[c] >= 0) { pairs[2 * k] = c; pairs[2 * k + 1] = mk.g.col_star[c]; k++; }
      }
    }
    *n_pairs = (act == 0) ? k : -1;
  }
}

__global__ void kf_init_kernel(double *x, double *P, const float4 *dets, int n, bool nep50) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 d4 = dets[i];
  const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
  double xv[7], Pv[49];
  kf_init(dd, xv, Pv, nep50);
  for (int k = 0; k < 7; k++) x[7 * (size_t)i + k] = xv[k];
  for (int k = 0; k < 49; k++) P[49 * (size_t)i + k] = Pv[k];
}

__global__ void kf_predict_kernel(double *x, double *P, double *boxes, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double xv[7], Pv[49];
  for (int k = 0; k < 7; k++) xv[k] = x[7 * (size_t)i + k];
  for (int k = 0; k < 49; k++) Pv[k] = P[49 * (size_t)i + k];
  double pb[kBlockP];
  if (kfb_from_dense(Pv, pb)) {  // the tracker's path: block form
    kfb_predict(xv, pb);
    kfb_to_dense(pb, Pv);
  } else {
    kf_predict(xv, Pv);
  }
  for (int k = 0; k < 7; k++) x[7 * (size_t)i + k] = xv[k];
  for (int k = 0; k < 49; k++) P[49 * (size_t)i + k] = Pv[k];
  if (boxes) {
    double b[4];
    x_to_bbox(xv, b);
    for (int k = 0; k < 4; k++) boxes[4 * (size_t)i + k] = b[k];
  }
}

__global__ void kf_update_kernel(double *x, double *P, const float4 *dets, double *boxes, int n, bool nep50) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 d4 = dets[i];
  const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
  double xv[7], Pv[49];
  for (int k = 0; k < 7; k++) xv[k] = x[7 * (size_t)i + k];
  for (int k = 0; k < 49; k++) Pv[k] = P[49 * (size_t)i + k];
  double pb[kBlockP];
  if (kfb_from_dense(Pv, pb)) {  // the tracker's path: block form
    kfb_update(xv, pb, dd, nep50);
    kfb_to_dense(pb, Pv);
  } else {
    kf_update(xv, Pv, dd, nep50);
  }
  for (int k = 0; k < 7; k++) x[7 * (size_t)i + k] = xv[k];
  for (int k = 0; k < 49; k++) P[49 * (size_t)i + k] = Pv[k];
  if (boxes) {
    double b[4];
    x_to_bbox(xv, b);
    for (int k = 0; k < 4; k++) boxes[4 * (size_t)i + k] = b[k];
  }
}

__global__ void bbox_to_z_kernel(const float4 *dets, double *z, int n, bool nep50) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 d4 = dets[i];
  double zz[4];
  bbox_to_z_d(d4.x, d4.y, d4.z, d4.w, nep50, zz);
  for (int k = 0; k < 4; k++) z[4 * (size_t)i + k] = zz[k];
}

__global__ void x_to_bbox_kernel(const double *x, int ldx, double *boxes, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double xv[7] = {0., 0., 0., 0., 0., 0., 0.}, b[4];
  for (int k = 0; k < 4; k++) xv[k] = x[(size_t)ldx * i + k];
  x_to_bbox(xv, b);
  for (int k = 0; k < 4; k++) boxes[4 * (size_t)i + k] = b[k];
}

// bbox_vote (detnet/utils/box_utils.py:401-430): one CTA per kept box; the boxes that overlap it by IoU >= thresh
// vote with their scores.  T = the tensors' dtype.  The reference's torch.sum order is not reproduced (tree
// reduction here), so results agree to rounding, not bit for bit.
template <typename T>
__global__ void __launch_bounds__(128) bbox_vote_kernel(const double *nms_boxes, int n, const double *all_boxes,
                                                        const double *all_scores, int m, double thresh, double *out) {
  __shared__ T red[5][4];
  const int i = blockIdx.x;
  const T b0 = (T)nms_boxes[4 * i], b1 = (T)nms_boxes[4 * i + 1], b2 = (T)nms_boxes[4 * i + 2], b3 = (T)nms_boxes[4 * i + 3];
  const T nms_area = (b2 - b0) * (b3 - b1);
  T acc[5] = {(T)0, (T)0, (T)0, (T)0, (T)0};
  for (int j = threadIdx.x; j < m; j += blockDim.x) {
    const T x1 = (T)all_boxes[4 * j], y1 = (T)all_boxes[4 * j + 1], x2 = (T)all_boxes[4 * j + 2], y2 = (T)all_boxes[4 * j + 3];
    const T area = (x2 - x1) * (y2 - y1);
    const T xx1 = x1 < b0 ? b0 : x1, yy1 = y1 < b1 ? b1 : y1;   // clamp(min=...)
    const T xx2 = x2 > b2 ? b2 : x2, yy2 = y2 > b3 ? b3 : y2;   // clamp(max=...)
    T w = xx2 - xx1, h = yy2 - yy1;
    if (w < (T)0) w = (T)0;
    if (h < (T)0) h = (T)0;
    const T inter = w * h;
    const T uni = (area + nms_area) - inter;
    const T iou = inter / uni;
    if (iou >= (T)thresh) {
      const T sc = (T)all_scores[j];
      acc[0] += x1 * sc; acc[1] += y1 * sc; acc[2] += x2 * sc; acc[3] += y2 * sc; acc[4] += sc;
    }
  }
#pragma unroll
  for (int k = 0; k < 5; k++) {
    T v = acc[k];
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    const T tot = (red[4][0] + red[4][1]) + (red[4][2] + red[4][3]);
    const T sum = (red[threadIdx.x][0] + red[threadIdx.x][1]) + (red[threadIdx.x][2] + red[threadIdx.x][3]);
    out[4 * i + threadIdx.x] = (double)(sum / tot);
  }
}

}  // namespace

extern "C" int w2t_bbox_vote(const double *nms_boxes, int32_t n, const double *all_boxes, const double *all_scores,
                             int32_t m, double thresh, int32_t compute_f32, double *out, w2t_stream_t stream) {
  if (n < 0 || m < 0) {
    set_last_error("w2t_bbox_vote: bad argument (negative size)");
    return W2T_ERR_ARG;
  }
  if (n == 0) return W2T_OK;
  if (!nms_boxes || !out || (m > 0 && (!all_boxes || !all_scores))) {
    set_last_error("w2t_bbox_vote: bad argument (null pointer)");
    return W2T_ERR_ARG;
  }
  if (compute_f32) bbox_vote_kernel<float><<<n, 128, 0, (cudaStream_t)stream>>>(nms_boxes, n, all_boxes, all_scores, m, thresh, out);
  else bbox_vote_kernel<double><<<n, 128, 0, (cudaStream_t)stream>>>(nms_boxes, n, all_boxes, all_scores, m, thresh, out);
  W2T_CUDA_TRY(cudaGetLastError());
  return W2T_OK;
}

extern "C" int w2t_bbox_to_z(const float *dets, double *z, int32_t n, int32_t promotion, w2t_stream_t stream) {
  if (n < 0) {
    set_last_error("w2t_bbox_to_z: bad argument (negative size)");
    return W2T_ERR_ARG;
  }
  if (n == 0) return W2T_OK;
  if (!dets || !z) {
    set_last_error("w2t_bbox_to_z: bad argument (null pointer)");
    return W2T_ERR_ARG;
  }
  bbox_to_z_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4 *>(dets), z, n,
                                                                    promotion == W2T_PROMOTION_NEP50);
  W2T_CUDA_TRY(cudaGetLastError());
  return W2T_OK;
}

extern "C" int w2t_x_to_bbox(const double *x, int32_t ldx, double *boxes, int32_t n, w2t_stream_t stream) {
  if (n < 0 || ldx < 4) {
    set_last_error("w2t_x_to_bbox: bad argument (negative size)");
    return W2T_ERR_ARG;
  }
  if (n == 0) return W2T_OK;
  if (!x || !boxes) {
    set_last_error("w2t_x_to_bbox: bad argument (null pointer)");
    return W2T_ERR_ARG;
  }
  x_to_bbox_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(x, ldx, boxes, n);
  W2T_CUDA_TRY(cudaGetLastError());
  return W2T_OK;
}

extern "C" int w2t_iou_matrix(const float *dets, int32_t D, const double *trks, int32_t T, float *out,
                              w2t_stream_t stream) {
  if (D < 0 || T < 0) {
    set_last_error("w2t_iou_matrix: bad argument (negative size)");
    return W2T_ERR_ARG;
  }
  if (D == 0 || T == 0) return W2T_OK;
  if (!dets || !trks || !out) {
    set_last_error("w2t_iou_matrix: bad argument (null pointer)");
    return W2T_ERR_ARG;
  }
  const size_t total = (size_t)D * T;
  const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 8);
  iou_matrix_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4 *>(dets), D, trks, T, out);
  W2T_CUDA_TRY(cudaGetLastError());
  return W2T_OK;
}

extern "C" size_t w2t_linear_assignment_workspace(int32_t D, int32_t T) {
  if (D <= 0 || T <= 0) return 256;
  return lap_layout(D, T).total;
}

extern "C" int w2t_linear_assignment(const float *cost, int32_t D, int32_t T, int32_t *pairs, int32_t *n_pairs,
                                     void *workspace, w2t_stream_t stream) {
  if (D < 0 || T < 0 || !n_pairs) {
    set_last_error("w2t_linear_assignment: bad argument (negative size)");
    return W2T_ERR_ARG;
  }
  if (D == 0 || T == 0) {
    W2T_CUDA_TRY(cudaMemsetAsync(n_pairs, 0, sizeof(int32_t), (cudaStream_t)stream));
    return W2T_OK;
  }
  if (!cost || !pairs || !workspace) {
    set_last_error("w2t_linear_assignment: bad argument (null pointer)");
    return W2T_ERR_ARG;
  }
  if (std::max(D, T) > kMunkresMaxDim) {
    set_last_error("w2t_linear_assignment: max(D,T)=%d exceeds %d", std::max(D, T), kMunkresMaxDim);
    return W2T_ERR_CAPACITY;
  }
  lap_kernel<256><<<1, 256, 0, (cudaStream_t)stream>>>(cost, D, T, pairs, n_pairs, static_cast<char *>(workspace));
  W2T_CUDA_TRY(cudaGetLastError());
  return W2T_OK;
}

extern "C" int w2t_kf_init(double *x, double *P, const float *dets, int32_t n, int32_t promotion, w2t_stream_t stream) {
  if (n < 0) {
    set_last_error("w2t_kf_init: bad argument (negative size)");
    return W2T_ERR_ARG;
  }
  if (n == 0) return W2T_OK;
  if (!x || !P || !dets) {
    set_last_error("w2t_kf_init: bad argument (null pointer)");
    return W2T_ERR_ARG;
  }
  kf_init_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(x, P, reinterpret_cast<const float4 *>(dets), n,
                                                                 promotion == W2T_PROMOTION_NEP50);
  W2T_CUDA_TRY(cudaGetLastError());
  return W2T_OK;
}

extern "C" int w2t_kf_predict(double *x, double *P, double *boxes, int32_t n, w2t_stream_t stream) {
  if (n < 0) {
    set_last_error("w2t_kf_predict: bad argument (negative size)");
    return W2T_ERR_ARG;
  }
  if (n == 0) return W2T_OK;
  if (!x || !P) {
    set_last_error("w2t_kf_predict: bad argument (null pointer)");
    return W2T_ERR_ARG;
  }
  kf_predict_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(x, P, boxes, n);
  W2T_CUDA_TRY(cudaGetLastError());
  return W2T_OK;
}

extern "C" int w2t_kf_update(double *x, double *P, const float *dets, double *boxes, int32_t n, int32_t promotion,
                             w2t_stream_t stream) {
  if (n < 0) {
    set_last_error("w2t_kf_update: bad argument (negative size)");
    return W2T_ERR_ARG;
  }
  if (n == 0) return W2T_OK;
  if (!x || !P || !dets) {
    set_last_error("w2t_kf_update: bad argument (null pointer)");
    return W2T_ERR_ARG;
  }
  kf_update_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(x, P, reinterpret_cast<const float4 *>(dets), boxes, n,
                                                                   promotion == W2T_PROMOTION_NEP50);
  W2T_CUDA_TRY(cudaGetLastError());
  return W2T_OK;
}

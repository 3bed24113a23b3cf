// softnms.cu — batched per-(image, category) soft-NMS merge of several submissions.
//
// Replaces, for every group at once, ensemble() (detnet/ensemble.py:50-64) ->
// nms_detections() (detnet/nn/tta.py:8-19) -> nms(soft=True)
// (detnet/utils/box_utils.py:307-395), and optionally the filters / box conversion that
// read_data_file() and track_sort() apply to the ensemble output before tracking
// (tracking/utils.py:79-87,32-35; tracker_sort.py:45).
//
// The reference's "soft-NMS" sorts the scores ONCE (box_utils.py:324) and never re-sorts
// after decaying, and with conf_thresh = 0 nothing is ever dropped, so the result is a
// fixed-order triangular product, all FP64:
//     s'_j = s_j * prod_{i ranked above j} clamp((cut - IoU_ij) / (cut - thr), 0, 1)
// multiplied in rank order.  One CTA per group; boxes are sorted into shared memory (point
// form + area), then one thread per box j walks i < j reading box i as a shared-memory
// broadcast.  Pairs that do not overlap have IoU = +0 and weight clamp(cut/(cut-thr)) — exactly
// 1.0 for every sensible setting — so they skip both FP64 divides.
//
// Tie rule: the reference's sort is unstable; the canonical order used here
// and by the oracle is (score descending, concatenation index descending).
#include <algorithm>

#include "common.cuh"

using namespace w2t;

namespace {

constexpr int kBytesPerBox = 8 /*raw score*/ + 6 * 8 /*x1 y1 x2 y2 area score*/ + 4 /*source index*/;
constexpr int kMaxSmem = 227 * 1024;

struct NmsParams {
  w2t_nms_problem_t p;
  w2t_nms_result_t r;
  double score_thr[W2T_MAX_CLASSES];
  int has_thr;
  int cap;  // boxes the shared-memory arrays hold
  int32_t *status;
};

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) softnms_kernel(const NmsParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_scan[2 * (BLOCK / 32)];
  const int cap = P.cap;
  double *raw = reinterpret_cast<double *>(smem_raw);  // [cap] scores in input order
  double *sx1 = raw + cap, *sy1 = sx1 + cap, *sx2 = sy1 + cap, *sy2 = sx2 + cap;
  double *sar = sy2 + cap, *ssc = sar + cap;
  int *src = reinterpret_cast<int *>(ssc + cap);

  const int tid = threadIdx.x;
  const int g = blockIdx.x;
  const int base = P.p.group_offsets[g];
  const int n = P.p.group_offsets[g + 1] - base;
  if (n > cap) {
    if (tid == 0) {
      if (P.status) atomicMax(P.status, W2T_ERR_CAPACITY);
      P.r.ens_count[g] = 0;
      if (P.r.trk_count) P.r.trk_count[g] = 0;
    }
    return;
  }
  const double *rows = P.p.rows + 5 * (size_t)base;

  // 1. scores to shared memory
  bool bad = false;
  for (int i = tid; i < n; i += BLOCK) {
    const double sc = rows[5 * i];
    raw[i] = sc;
    if (!(sc >= 0.)) bad = true;  // the reference would drop it at `ge(conf_thresh)` (box_utils.py:379)
  }
  __syncthreads();

  // 2. rank by counting, scatter the point-form box into rank order
  for (int i = tid; i < n; i += BLOCK) {
    const double si = raw[i];
    int rank = 0;
    for (int k = 0; k < n; k++) {
      const double sk = raw[k];
      rank += (sk > si || (sk == si && k > i)) ? 1 : 0;
    }
    const double *rw = rows + 5 * i;
    const double w = rw[3], h = rw[4];
    const double cx = rw[1] + w / 2, cy = rw[2] + h / 2;  // lxly2cxcy, ensemble.py:19-22
    const double hw = w * 0.5, hh = h * 0.5;              // point_form, box_utils.py:32-35
    const double x1 = cx - hw, y1 = cy - hh, x2 = cx + hw, y2 = cy + hh;
    const double area = (x2 - x1) * (y2 - y1);            // box_utils.py:342
    if (!(area > 0.)) bad = true;
    sx1[rank] = x1; sy1[rank] = y1; sx2[rank] = x2; sy2[rank] = y2;
    sar[rank] = area;
    ssc[rank] = si;
    src[rank] = i;
  }
  if (bad && P.status) atomicMax(P.status, W2T_ERR_ARG);
  __syncthreads();

  // 3. triangular decay + outputs, chunk by chunk in rank order
  const double cut = P.p.soft_nms_cut;
  const double denom = cut - P.p.iou_thresh;
  double wt0 = (cut - 0.0) / denom;  // weight of a pair with IoU = +0
  if (wt0 < 0.) wt0 = 0.;
  if (wt0 > 1.) wt0 = 1.;
  const bool skip_disjoint = (wt0 == 1.0);
  const int cls = (P.p.n_classes > 0) ? (g % P.p.n_classes) : 0;
  int n_ens = 0, n_trk = 0;
  for (int j0 = 0; j0 < n; j0 += BLOCK) {
    const int j = j0 + tid;
    bool f_ens = false, f_trk = false;
    int bx = 0, by = 0, bw = 0, bh = 0;
    double rs = 0.;
    if (j < n) {
      const double x1 = sx1[j], y1 = sy1[j], x2 = sx2[j], y2 = sy2[j], area = sar[j];
      double live = ssc[j];
      for (int i = 0; i < j; i++) {
        // box i is the kept (higher ranked) one; box_utils.py:349-370
        const double ix1 = sx1[i], iy1 = sy1[i], ix2 = sx2[i], iy2 = sy2[i];
        const double xx1 = x1 > ix1 ? x1 : ix1;
        const double yy1 = y1 > iy1 ? y1 : iy1;
        const double xx2 = x2 < ix2 ? x2 : ix2;
        const double yy2 = y2 < iy2 ? y2 : iy2;
        double w = xx2 - xx1;
        double h = yy2 - yy1;
        if (w < 0.) w = 0.;
        if (h < 0.) h = 0.;
        const double inter = w * h;
        if (inter == 0. && skip_disjoint) continue;  // IoU = +0 (areas are positive), weight 1.0
        const double uni = (area - inter) + sar[i];
        const double iou = inter / uni;
        double wt = (cut - iou) / denom;
        if (wt < 0.) wt = 0.;
        if (wt > 1.) wt = 1.;
        live = live * wt;
      }
      // center_size (box_utils.py:57-69) then cxcy2lxly (ensemble.py:25-28)
      const double cx = (x1 + x2) * 0.5, cy = (y1 + y2) * 0.5;
      const double w = x2 - x1, h = y2 - y1;
      const size_t o = (size_t)base + j;
      if (P.r.merged) {
        double *mr = P.r.merged + 5 * o;
        mr[0] = live; mr[1] = cx; mr[2] = cy; mr[3] = w; mr[4] = h;
      }
      if (P.r.src_index) P.r.src_index[o] = base + src[j];
      if (live > P.p.min_score) {  // ensemble.py:60
        f_ens = true;
        const double left = cx - w / 2, top = cy - h / 2;
        bx = (int)(long long)left; by = (int)(long long)top;  // astype(int), ensemble.py:62
        bw = (int)(long long)w; bh = (int)(long long)h;
        rs = rint(live * 1e5) / 1e5;  // round(np.float64, 5), ensemble.py:62
        if (P.has_thr && !(bw < 1 || bh < 1) && !(rs < P.score_thr[cls])) f_trk = true;  // utils.py:79-87
      }
    }
    int ee, et, te, tt;
    block_scan2<BLOCK>(f_ens, f_trk, s_scan, ee, et, te, tt);
    if (f_ens) {
      const size_t e = (size_t)base + n_ens + ee;
      int32_t *eb = P.r.ens_box + 4 * e;
      eb[0] = bx; eb[1] = by; eb[2] = bw; eb[3] = bh;
      P.r.ens_score[e] = rs;
    }
    if (f_trk) {
      const size_t t = (size_t)base + n_trk + et;  // utils.py:32-35, tracker_sort.py:45
      reinterpret_cast<float4 *>(P.r.trk_box)[t] =
          make_float4((float)bx, (float)by, (float)(bx + bw), (float)(by + bh));
    }
    n_ens += te;
    n_trk += tt;
  }
  if (tid == 0) {
    P.r.ens_count[g] = n_ens;
    if (P.r.trk_count) P.r.trk_count[g] = n_trk;
    if (P.r.img_exists && P.p.n_classes > 0 && n_ens > 0) P.r.img_exists[g / P.p.n_classes] = 1;
  }
}

template <int BLOCK>
int launch(const NmsParams &P, int n_groups, size_t smem, cudaStream_t stream) {
  if (smem > 48 * 1024)
    W2T_CUDA_TRY(cudaFuncSetAttribute(softnms_kernel<BLOCK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  softnms_kernel<BLOCK><<<n_groups, BLOCK, smem, stream>>>(P);
  W2T_CUDA_TRY(cudaGetLastError());
  return W2T_OK;
}

}  // namespace

extern "C" int w2t_softnms_max_group(void) { return (kMaxSmem - 1024) / kBytesPerBox; }

extern "C" int w2t_softnms_groups(const w2t_nms_problem_t *problem, w2t_nms_result_t *result, int max_group_size,
                                  int32_t *status, w2t_stream_t stream) {
  if (!problem || !result || problem->n_groups < 0 || problem->n_classes < 0 ||
      problem->n_classes > W2T_MAX_CLASSES || max_group_size < 0) {
    set_last_error("w2t_softnms_groups: bad argument");
    return W2T_ERR_ARG;
  }
  if (problem->n_groups == 0) return W2T_OK;
  if (!problem->group_offsets || !problem->rows || !result->ens_count || !result->ens_box || !result->ens_score) {
    set_last_error("w2t_softnms_groups: null buffer");
    return W2T_ERR_ARG;
  }
  if (problem->score_thr && (!result->trk_count || !result->trk_box || problem->n_classes < 1)) {
    set_last_error("w2t_softnms_groups: score_thr given without trk_count/trk_box/n_classes");
    return W2T_ERR_ARG;
  }
  if (max_group_size > w2t_softnms_max_group()) {
    set_last_error("w2t_softnms_groups: group of %d boxes exceeds the shared-memory limit of %d", max_group_size,
                   w2t_softnms_max_group());
    return W2T_ERR_CAPACITY;
  }
  NmsParams P;
  P.p = *problem;
  P.r = *result;
  P.has_thr = problem->score_thr != nullptr;
  for (int i = 0; i < W2T_MAX_CLASSES; i++)
    P.score_thr[i] = (P.has_thr && i < problem->n_classes) ? problem->score_thr[i] : 0.0;
  P.p.score_thr = nullptr;
  P.cap = (std::max(max_group_size, 1) + 1) & ~1;  // even: keeps the int array 8-byte aligned
  P.status = status;
  const size_t smem = (size_t)P.cap * kBytesPerBox;
  if (max_group_size <= 96) return launch<64>(P, problem->n_groups, smem, (cudaStream_t)stream);
  if (max_group_size <= 768) return launch<128>(P, problem->n_groups, smem, (cudaStream_t)stream);
  return launch<256>(P, problem->n_groups, smem, (cudaStream_t)stream);
}

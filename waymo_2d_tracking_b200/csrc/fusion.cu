// fusion.cu — batched per-(image, category) confidence-weighted box fusion.
//
// Replaces, for every group at once, ensemble() (detnet/ensemble.py:50-64) with
// merge_func = merge_detections (detnet/nn/tta.py:22-66), the reference CLI's default method
// (`-m weighted_fusion`, ensemble.py:94,138), including jaccard_bbox / iou_bbox / intersect
// (detnet/utils/box_utils.py:72-140) and, optionally, the hand-over filters of the tracker
// (tracking/utils.py:79-87,32-35; tracker_sort.py:45).
//
// The reference folds the submissions into a running result list one after the other:
//   * every row becomes (s', s'*cx, s'*cy, s'*w, s'*h) with s' = score / n_submissions;
//   * each box of the next submission looks for the result with the largest IoU (boxes are
//     de-weighted by a division first; first maximum wins); with IoU >= thresh its row is ADDED
//     to that result, with IoU < thresh it is appended as a new result;
//   * `results[idx] += rows` is a NumPy fancy-index update: when several boxes pick the same
//     result only the LAST one (largest index) is added, all others are lost;
//   * at the end coordinates are divided by the accumulated score.
// One CTA per group keeps the running results in shared memory: accumulators (5 doubles), the
// de-weighted point-form box and area of the current round (5 doubles) and the winning row
// (int).  Rounds are sequential (a barrier apart); inside a round one thread per incoming box
// scans the results (shared-memory broadcast reads), and appends are a stable block scan.
#include <algorithm>

#include "common.cuh"

using namespace w2t;

namespace {

constexpr int kBytesPerSlot = 10 * 8 + 4;
constexpr int kMaxSmem = 227 * 1024;

struct FusionParams {
  w2t_nms_problem_t p;
  w2t_nms_result_t r;
  const int32_t *sub_counts;  // [n_groups, n_sub]
  int n_sub;
  double score_thr[W2T_MAX_CLASSES];
  int has_thr;
  int cap;
  int32_t *status;
};

struct Row5 { double v[5]; };

// one input row -> the reference's weighted row (tta.py:34-35,39-40), after lxly2cxcy (ensemble.py:19-22)
__device__ __forceinline__ Row5 weighted_row(const double *rw, int fmt, double n_sub) {
  Row5 o;
  const double w = rw[3], h = rw[4];
  double cx = rw[1], cy = rw[2];
  if (fmt == W2T_BOX_LTWH) { cx = cx + w / 2; cy = cy + h / 2; }
  const double s = rw[0] / n_sub;
  o.v[0] = s; o.v[1] = cx * s; o.v[2] = cy * s; o.v[3] = w * s; o.v[4] = h * s;
  return o;
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) fusion_kernel(const FusionParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_scan[2 * (BLOCK / 32)];
  const int cap = P.cap;
  double *acc = reinterpret_cast<double *>(smem_raw);  // [5][cap] accumulators
  double *bx1 = acc + 5 * (size_t)cap, *by1 = bx1 + cap, *bx2 = by1 + cap, *by2 = bx2 + cap, *bar = by2 + cap;
  int *last = reinterpret_cast<int *>(bar + cap);

  const int tid = threadIdx.x;
  const int g = blockIdx.x;
  const int base = P.p.group_offsets[g];
  const int n = P.p.group_offsets[g + 1] - base;
  if (n > cap) {
    if (tid == 0) {
      if (P.status) atomicMax(P.status, W2T_ERR_CAPACITY);
      P.r.ens_count[g] = 0;
      if (P.r.trk_count) P.r.trk_count[g] = 0;
      if (P.r.kept_count) P.r.kept_count[g] = 0;
    }
    return;
  }
  const double *rows = P.p.rows + 5 * (size_t)base;
  const int fmt = P.p.box_format;
  const double nsub = (double)P.n_sub;
  const double thr = P.p.iou_thresh;
  const int32_t *cnt = P.sub_counts + (size_t)g * P.n_sub;

  int R = 0;      // running results
  int row0 = 0;   // first row of the current submission inside the group
  bool bad = false;
  for (int k = 0; k < P.n_sub; k++) {
    const int nk = cnt[k];
    if (nk == 0) continue;
    if (R == 0) {
      // the first non-empty submission becomes the result list as it is (tta.py:33,62-64)
      for (int o = tid; o < nk; o += BLOCK) {
        const double *rw = rows + 5 * (size_t)(row0 + o);
        if (!(rw[0] > 0.)) bad = true;  // 0/0 below: the reference silently loses such rows
        const Row5 w = weighted_row(rw, fmt, nsub);
#pragma unroll
        for (int i = 0; i < 5; i++) acc[(size_t)i * cap + o] = w.v[i];
      }
      R = nk;
      row0 += nk;
      __syncthreads();
      continue;
    }
    // a. de-weighted boxes of the current results (tta.py:43), point form + area (box_utils.py:123-139)
    for (int r = tid; r < R; r += BLOCK) {
      const double s = acc[r];
      const double cx = acc[(size_t)cap + r] / s, cy = acc[2 * (size_t)cap + r] / s;
      const double w = acc[3 * (size_t)cap + r] / s, h = acc[4 * (size_t)cap + r] / s;
      const double hw = w * 0.5, hh = h * 0.5;
      bx1[r] = cx - hw; by1[r] = cy - hh; bx2[r] = cx + hw; by2[r] = cy + hh;
      bar[r] = w * h;
      last[r] = -1;
    }
    __syncthreads();
    // b. best result for every incoming box; d. stable append of the unmatched ones
    int appended = 0;
    for (int o0 = 0; o0 < nk; o0 += BLOCK) {
      const int o = o0 + tid;
      bool unmatched = false;
      Row5 w;
      if (o < nk) {
        const double *rw = rows + 5 * (size_t)(row0 + o);
        if (!(rw[0] > 0.)) bad = true;
        w = weighted_row(rw, fmt, nsub);
        const double s = w.v[0];
        const double cx = w.v[1] / s, cy = w.v[2] / s, ow = w.v[3] / s, oh = w.v[4] / s;  // tta.py:44
        const double hw = ow * 0.5, hh = oh * 0.5;
        const double x1 = cx - hw, y1 = cy - hh, x2 = cx + hw, y2 = cy + hh;
        const double area = ow * oh;
        double best = -1.0;
        int arg = 0;
        for (int r = 0; r < R; r++) {
          const double mx2 = bx2[r] < x2 ? bx2[r] : x2, my2 = by2[r] < y2 ? by2[r] : y2;
          const double mx1 = bx1[r] > x1 ? bx1[r] : x1, my1 = by1[r] > y1 ? by1[r] : y1;
          double iw = mx2 - mx1, ih = my2 - my1;
          if (iw < 0.) iw = 0.;
          if (ih < 0.) ih = 0.;
          const double inter = iw * ih;
          double iou = 0.0;
          if (inter != 0.) iou = inter / ((bar[r] + area) - inter);
          else if (!((bar[r] + area) > 0.)) iou = inter / ((bar[r] + area) - inter);  // 0/0 or 0/negative
          if (iou > best) { best = iou; arg = r; }   // first maximum wins (torch.max)
          else if (iou != iou) bad = true;
        }
        if (best >= thr) atomicMax(&last[arg], o);    // NumPy fancy `+=`: the last writer wins
        else if (best < thr) unmatched = true;
      }
      int ea, eb, ta, tb;
      block_scan2<BLOCK>(unmatched, false, s_scan, ea, eb, ta, tb);
      if (unmatched) {
        const int slot = R + appended + ea;
#pragma unroll
        for (int i = 0; i < 5; i++) acc[(size_t)i * cap + slot] = w.v[i];
      }
      appended += ta;
    }
    __syncthreads();
    // c. matched results absorb their last matching row (tta.py:50-54)
    for (int r = tid; r < R; r += BLOCK) {
      const int o = last[r];
      if (o >= 0) {
        const Row5 w = weighted_row(rows + 5 * (size_t)(row0 + o), fmt, nsub);
#pragma unroll
        for (int i = 0; i < 5; i++) acc[(size_t)i * cap + r] = acc[(size_t)i * cap + r] + w.v[i];
      }
    }
    R += appended;
    row0 += nk;
    __syncthreads();
  }
  if (bad && P.status) atomicMax(P.status, W2T_ERR_ARG);

  // outputs in result order (tta.py:67-68, ensemble.py:57-63)
  const int cls = (P.p.n_classes > 0) ? (g % P.p.n_classes) : 0;
  int n_ens = 0, n_trk = 0;
  for (int j0 = 0; j0 < R; j0 += BLOCK) {
    const int j = j0 + tid;
    bool f_ens = false, f_trk = false;
    int bx = 0, by = 0, bw = 0, bh = 0;
    double rs = 0.;
    if (j < R) {
      const double s = acc[j];
      const double cx = acc[(size_t)cap + j] / s, cy = acc[2 * (size_t)cap + j] / s;
      const double w = acc[3 * (size_t)cap + j] / s, h = acc[4 * (size_t)cap + j] / s;
      if (P.r.merged) {
        double *mr = P.r.merged + 5 * ((size_t)base + j);
        mr[0] = s; mr[1] = cx; mr[2] = cy; mr[3] = w; mr[4] = h;
      }
      if (s > P.p.min_score) {
        f_ens = true;
        const double left = cx - w / 2, top = cy - h / 2;  // cxcy2lxly, ensemble.py:25-28
        bx = (int)(long long)left; by = (int)(long long)top;
        bw = (int)(long long)w; bh = (int)(long long)h;
        rs = rint(s * 1e5) / 1e5;
        if (P.has_thr && !(bw < 1 || bh < 1) && !(rs < P.score_thr[cls])) f_trk = true;
      }
    }
    int ee, et, te, tt;
    block_scan2<BLOCK>(f_ens, f_trk, s_scan, ee, et, te, tt);
    if (f_ens && P.r.ens_box) {
      const size_t e = (size_t)base + n_ens + ee;
      int32_t *eb = P.r.ens_box + 4 * e;
      eb[0] = bx; eb[1] = by; eb[2] = bw; eb[3] = bh;
      P.r.ens_score[e] = rs;
    }
    if (f_trk) {
      const size_t t = (size_t)base + n_trk + et;
      reinterpret_cast<float4 *>(P.r.trk_box)[t] =
          make_float4((float)bx, (float)by, (float)(bx + bw), (float)(by + bh));
    }
    n_ens += te;
    n_trk += tt;
  }
  if (tid == 0) {
    if (P.r.kept_count) P.r.kept_count[g] = R;
    P.r.ens_count[g] = n_ens;
    if (P.r.trk_count) P.r.trk_count[g] = n_trk;
    if (P.r.img_exists && P.p.n_classes > 0 && n_ens > 0) P.r.img_exists[g / P.p.n_classes] = 1;
  }
}

template <int BLOCK>
int launch(const FusionParams &P, int n_groups, size_t smem, cudaStream_t stream) {
  if (smem > 48 * 1024)
    W2T_CUDA_TRY(cudaFuncSetAttribute(fusion_kernel<BLOCK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  fusion_kernel<BLOCK><<<n_groups, BLOCK, smem, stream>>>(P);
  W2T_CUDA_TRY(cudaGetLastError());
  return W2T_OK;
}

}  // namespace

extern "C" int w2t_fusion_max_group(void) { return (kMaxSmem - 1024) / kBytesPerSlot; }

extern "C" int w2t_fusion_groups(const w2t_nms_problem_t *problem, const int32_t *sub_counts, int32_t n_sub,
                                 w2t_nms_result_t *result, int max_group_size, int32_t *status,
                                 w2t_stream_t stream) {
  if (!problem || !result || problem->n_groups < 0 || problem->n_classes < 0 ||
      problem->n_classes > W2T_MAX_CLASSES || max_group_size < 0 || n_sub < 1) {
    set_last_error("w2t_fusion_groups: bad argument");
    return W2T_ERR_ARG;
  }
  if (problem->n_groups == 0) return W2T_OK;
  if (!problem->group_offsets || !problem->rows || !sub_counts || !result->ens_count ||
      (result->ens_box && !result->ens_score) ||
      (problem->box_format != W2T_BOX_LTWH && problem->box_format != W2T_BOX_CXCYWH)) {
    set_last_error("w2t_fusion_groups: null buffer or unsupported box_format");
    return W2T_ERR_ARG;
  }
  if (problem->score_thr && (!result->trk_count || !result->trk_box || problem->n_classes < 1)) {
    set_last_error("w2t_fusion_groups: score_thr given without trk_count/trk_box/n_classes");
    return W2T_ERR_ARG;
  }
  if (max_group_size > w2t_fusion_max_group()) {
    set_last_error("w2t_fusion_groups: group of %d boxes exceeds the shared-memory limit of %d", max_group_size,
                   w2t_fusion_max_group());
    return W2T_ERR_CAPACITY;
  }
  FusionParams P;
  P.p = *problem;
  P.r = *result;
  P.sub_counts = sub_counts;
  P.n_sub = n_sub;
  P.has_thr = problem->score_thr != nullptr;
  for (int i = 0; i < W2T_MAX_CLASSES; i++)
    P.score_thr[i] = (P.has_thr && i < problem->n_classes) ? problem->score_thr[i] : 0.0;
  P.p.score_thr = nullptr;
  P.cap = (std::max(max_group_size, 1) + 1) & ~1;
  P.status = status;
  const size_t smem = (size_t)P.cap * kBytesPerSlot;
  if (max_group_size <= 96) return launch<64>(P, problem->n_groups, smem, (cudaStream_t)stream);
  if (max_group_size <= 768) return launch<128>(P, problem->n_groups, smem, (cudaStream_t)stream);
  return launch<256>(P, problem->n_groups, smem, (cudaStream_t)stream);
}

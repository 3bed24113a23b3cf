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
#include <climits>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

using namespace w2t;

namespace {

constexpr int kBytesPerBox = 8 /*raw score*/ + 6 * 8 /*x1 y1 x2 y2 area score*/ + 4 /*source index*/ + 4 /*strip mask*/;
constexpr int kMaxSmem = 227 * 1024;
constexpr int kMaskPad = 32;  // the mask array is read 32 entries at a time

struct NmsParams {
  w2t_nms_problem_t p;
  w2t_nms_result_t r;
  double score_thr[W2T_MAX_CLASSES];
  int has_thr;
  int cap;  // boxes the arrays hold
  int32_t *status;
  // oversized groups (more boxes than shared memory holds): a second launch of a few persistent CTAs keeps the
  // arrays in global memory (`work`, one slice of `work_stride` bytes per CTA) and serves only groups with more
  // than `small_cap` boxes; the regular launch skips those (`skip_big`)
  unsigned char *work;
  size_t work_stride;
  int small_cap;
  int skip_big;
  // size classes: a launch serves the groups with n_lo < boxes <= n_hi (n_lo = -1: from empty groups on), so that
  // small groups get small CTAs (one warp for up to 32 boxes) and waste neither lanes nor block-wide barriers
  int n_lo, n_hi;
};

// HARD = false: the soft branch (box_utils.py:335-391).  HARD = true: the hard branch
// (box_utils.py:329-333), i.e. torchvision.ops.nms on the ascending-sorted boxes: greedy
// suppression of every lower ranked box with IoU > overlap, scores untouched.
// Coarse occupancy mask of a box: bits 0-15 = the 128-unit column strips its x extent touches,
// bits 16-31 = the 128-unit row strips of its y extent (indices clamped to 0..15, so anything
// outside [0, 2048) lands in the edge strips).  The strip index is a monotone function of the
// coordinate: two boxes whose masks share no x strip (or no y strip) are strictly separated in
// x (or y), their intersection is exactly +0, and the pair needs no FP64 arithmetic at all.
// Irregular boxes (NaN, non-positive extent) get all ones = "always compute".
__device__ __forceinline__ uint32_t strip_mask(double x1, double y1, double x2, double y2) {
  if (!(x2 > x1 && y2 > y1)) return 0xffffffffu;
  const int a = min(max(__double2int_rd(x1 * (1.0 / 128.0)), 0), 15), b = min(max(__double2int_rd(x2 * (1.0 / 128.0)), 0), 15);
  const int c = min(max(__double2int_rd(y1 * (1.0 / 128.0)), 0), 15), d = min(max(__double2int_rd(y2 * (1.0 / 128.0)), 0), 15);
  const uint32_t mx = ((2u << b) - 1u) & ~((1u << a) - 1u);
  const uint32_t my = ((2u << d) - 1u) & ~((1u << c) - 1u);
  return (mx & 0xffffu) | (my << 16);
}

// T: the arithmetic type of the decay / suppression.  double: the ensemble path (ensemble.py:54, tta.py:11-12 are
// float64 end to end).  float: the detector head's call (detnet/nn/modules/detection.py:59-77 passes float32
// boxes and scores, and torch keeps every op of box_utils.py:335-391 in float32); inputs arrive as float64 rows
// holding float32 values, results are widened back.
template <int BLOCK, bool HARD, typename T>
__global__ void __launch_bounds__(BLOCK) softnms_kernel(const NmsParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_scan[2 * (BLOCK / 32)];
  const int cap = P.cap;
  unsigned char *mem = P.work ? P.work + (size_t)blockIdx.x * P.work_stride : smem_raw;
  double *raw = reinterpret_cast<double *>(mem);  // [cap] scores in input order
  T *sx1 = reinterpret_cast<T *>(raw + cap), *sy1 = sx1 + cap, *sx2 = sy1 + cap, *sy2 = sx2 + cap;
  T *sar = sy2 + cap, *ssc = sar + cap;
  int *src = reinterpret_cast<int *>(raw + 7 * (size_t)cap);  // same layout for both arithmetic types
  uint32_t *smk = reinterpret_cast<uint32_t *>(src + cap);  // coarse occupancy masks, see strip_mask()

  const int tid = threadIdx.x;
  // one group per CTA (regular launch: gridDim = groups); the oversized pass strides a few CTAs over all groups
#pragma unroll 1
  for (int g = blockIdx.x; g < P.p.n_groups; g += gridDim.x) {
  __syncthreads();  // the previous group's arrays are free
  const int base = P.p.group_offsets[g];
  const int n = P.p.group_offsets[g + 1] - base;
  if (P.work != nullptr && n <= P.small_cap) continue;  // the regular launch has it
  if (n <= P.n_lo || n > P.n_hi) continue;  // another size class has it
  if (n > cap) {
    if (P.skip_big) continue;  // the oversized pass has it
    if (tid == 0) {
      if (P.status) atomicMax(P.status, W2T_ERR_CAPACITY);
      P.r.ens_count[g] = 0;
      if (P.r.trk_count) P.r.trk_count[g] = 0;
      if (P.r.kept_count) P.r.kept_count[g] = 0;
    }
    continue;
  }
  const int fmt = P.p.box_format;
  // row i of the group: score and the four box columns, from 40-byte double rows or 16-byte compact rows
  const unsigned char *bytes = reinterpret_cast<const unsigned char *>(P.p.rows);
  auto row_score = [&](int i) -> double {
    if (fmt == W2T_BOX_LTWH_P64) {
      const unsigned long long q = *reinterpret_cast<const unsigned long long *>(bytes + 8 * ((size_t)base + i));
      return (double)(unsigned)(q & 0x1ffffu) / 100000.0;  // correctly rounded: the double Python parses "0.ddddd" to
    }
    return (fmt == W2T_BOX_LTWH_I16) ? *reinterpret_cast<const double *>(bytes + 16 * ((size_t)base + i))
                                     : P.p.rows[5 * ((size_t)base + i)];
  };
  auto row_box = [&](int i, double &b0, double &b1, double &b2, double &b3) {
    if (fmt == W2T_BOX_LTWH_P64) {
      const unsigned long long q = *reinterpret_cast<const unsigned long long *>(bytes + 8 * ((size_t)base + i));
      b0 = (double)((int)((q >> 17) & 0x1fffu) - 3072); b1 = (double)((int)((q >> 30) & 0xfffu) - 1536);
      b2 = (double)(unsigned)((q >> 42) & 0x7ffu); b3 = (double)(unsigned)(q >> 53);
    } else if (fmt == W2T_BOX_LTWH_I16) {
      const short4 q = *reinterpret_cast<const short4 *>(bytes + 16 * ((size_t)base + i) + 8);
      b0 = (double)q.x; b1 = (double)q.y; b2 = (double)q.z; b3 = (double)q.w;
    } else {
      const double *rw = P.p.rows + 5 * ((size_t)base + i);
      b0 = rw[1]; b1 = rw[2]; b2 = rw[3]; b3 = rw[4];
    }
  };

  // 1. scores to shared memory
  const double conf = P.p.conf_thresh;
  // removals make the result order-dependent (box_utils.py:379-381; always for hard NMS)
  const bool sequential = HARD || conf > 0.;
  bool bad = false;
  // Packed rows carry the score as a 17-bit integer (score * 1e5): (score, input position) then fits one 32-bit
  // key whose order IS the canonical order — score descending, later position first among equals — and the rank
  // of a box is a single count of larger keys.  (Soft branch without removals, groups of at most 1024 boxes.)
  const bool int_keys = !HARD && !sequential && fmt == W2T_BOX_LTWH_P64 && n <= 1024;
  uint32_t *ukey = reinterpret_cast<uint32_t *>(raw);  // instead of the input-order scores
  if (int_keys) {
    const int n4 = (n + 3) & ~3;  // the rank loop reads four keys at a time; 0 is never "larger"
    for (int i = tid; i < n4; i += BLOCK) {
      uint32_t key = 0u;
      if (i < n) {
        const unsigned long long q = *reinterpret_cast<const unsigned long long *>(bytes + 8 * ((size_t)base + i));
        key = ((uint32_t)(q & 0x1ffffu) << 10) | (uint32_t)i;
      }
      ukey[i] = key;
    }
    // scores are finite and >= 0 here; `ge(conf_thresh)` with conf <= 0 never fails
  } else {
    for (int i = tid; i < n; i += BLOCK) {
      const double sc = row_score(i);
      raw[i] = sc;
      // without removals the reference keeps every box only if no score ever fails `ge(conf_thresh)`
      if (!HARD && !sequential && !(sc >= (conf < 0. ? conf : 0.))) bad = true;
      if (sc != sc) bad = true;
    }
  }
  __syncthreads();

  // 2. rank by counting, scatter the point-form box into rank order; with top_k only the best
  //    top_k boxes take part (box_utils.py:325-327)
  const int m = (P.p.top_k > 0 && P.p.top_k < n) ? P.p.top_k : n;
  for (int i = tid; i < n; i += BLOCK) {
    double si;
    // G boxes score higher; of the equal ones L come earlier and H later in the input.  Ties are
    // rare, so the main loop only counts "greater" and "greater or equal" (two compares per box)
    // and the positions of the equal ones are looked at only when there are any.
    int G = 0, GE = 0, L = 0, H = 0;
    if (int_keys) {
      const uint32_t ui = ukey[i];
      const uint4 *k4 = reinterpret_cast<const uint4 *>(ukey);
      const int n4 = (n + 3) >> 2;
      for (int k = 0; k < n4; k++) {
        const uint4 v = k4[k];  // broadcast
        G += (v.x > ui) + (v.y > ui) + (v.z > ui) + (v.w > ui);
      }
      si = (double)(ui >> 10) / 100000.0;  // = row_score(i)
    } else {
      si = raw[i];
      for (int k = 0; k < n; k++) {
        const double sk = raw[k];
        G += (sk > si) ? 1 : 0;
        GE += (sk >= si) ? 1 : 0;
      }
      if (GE - G > 1) {
        for (int k = 0; k < n; k++) {
          const double sk = raw[k];
          L += (sk == si && k < i) ? 1 : 0;
          H += (sk == si && k > i) ? 1 : 0;
        }
      }
    }
    // The reference sorts ascending (stable, canonical rule) and consumes from the end, so among
    // equal scores the later box ranks higher; top_k cuts that order (box_utils.py:324-327).
    if (G + H >= m) continue;
    int rank = G + H;
    if (HARD) {
      // torchvision's own stable descending sort of the ascending array then puts the EARLIER of
      // two equal boxes first; equal boxes cut off by top_k are the earliest ones
      const int E = L + H + 1;
      const int sel = E < m - G ? E : m - G;
      rank = G + L - (E - sel);
    }
    double d0, d1, d2, d3;
    row_box(i, d0, d1, d2, d3);
    const T b0 = (T)d0, b1 = (T)d1, b2 = (T)d2, b3 = (T)d3;
    T x1, y1, x2, y2;
    if (fmt == W2T_BOX_XYXY) {
      x1 = b0; y1 = b1; x2 = b2; y2 = b3;
    } else {
      const T w = b2, h = b3;
      T cx = b0, cy = b1;
      if (fmt != W2T_BOX_CXCYWH) { cx = cx + w / 2; cy = cy + h / 2; }  // lxly2cxcy, ensemble.py:19-22
      const T hw = w * (T)0.5, hh = h * (T)0.5;                         // point_form, box_utils.py:32-35
      x1 = cx - hw; y1 = cy - hh; x2 = cx + hw; y2 = cy + hh;
    }
    const T area = (x2 - x1) * (y2 - y1);            // box_utils.py:342
    if (!HARD && !(area > (T)0)) bad = true;  // 0/0 weights: the reference would drop the box as NaN
    sx1[rank] = x1; sy1[rank] = y1; sx2[rank] = x2; sy2[rank] = y2;
    sar[rank] = area;
    ssc[rank] = (T)si;
    src[rank] = i;
    smk[rank] = strip_mask((double)x1, (double)y1, (double)x2, (double)y2);
  }
  if (bad && P.status) atomicMax(P.status, W2T_ERR_ARG);
  __syncthreads();

  // 3. decay
  // the python floats of box_utils.py:368 enter the tensor arithmetic in the tensor's dtype
  const T cut = (T)P.p.soft_nms_cut;
  const T denom = (T)(P.p.soft_nms_cut - P.p.iou_thresh);
  T wt0 = (cut - (T)0) / denom;  // weight of a pair with IoU = +0
  if (wt0 < (T)0) wt0 = (T)0;
  if (wt0 > (T)1) wt0 = (T)1;
  const bool skip_disjoint = (wt0 == (T)1);
  // weight box i (kept, higher ranked) applies to box j; box_utils.py:349-370
  auto pair_weight = [&](int i, T x1, T y1, T x2, T y2, T area, bool &one) -> T {
    const T ix1 = sx1[i], iy1 = sy1[i], ix2 = sx2[i], iy2 = sy2[i];
    // strictly separated boxes: the clamped width or height below is 0, so inter = +0 (sufficient,
    // not necessary: touching or degenerate cases fall through to the full computation)
    if (skip_disjoint && (x2 <= ix1 || ix2 <= x1 || y2 <= iy1 || iy2 <= y1)) { one = true; return (T)1; }
    const T xx1 = x1 > ix1 ? x1 : ix1;
    const T yy1 = y1 > iy1 ? y1 : iy1;
    const T xx2 = x2 < ix2 ? x2 : ix2;
    const T yy2 = y2 < iy2 ? y2 : iy2;
    T w = xx2 - xx1;
    T h = yy2 - yy1;
    if (w < (T)0) w = (T)0;
    if (h < (T)0) h = (T)0;
    const T inter = w * h;
    if (inter == (T)0 && skip_disjoint) { one = true; return (T)1; }  // IoU = +0 (areas are positive), weight 1.0
    one = false;
    const T uni = (area - inter) + sar[i];
    const T iou = inter / uni;
    T wt = (cut - iou) / denom;
    if (wt < (T)0) wt = (T)0;
    if (wt > (T)1) wt = (T)1;
    return wt;
  };
  // separated pairs leave box j untouched: weight exactly 1.0 (soft) / overlap 0 <= threshold (hard)
  const bool mask_skip = HARD ? (P.p.iou_thresh >= 0.) : skip_disjoint;
  int *alive = reinterpret_cast<int *>(raw);  // the input-order scores are no longer needed
  if (!sequential) {
    // Who can overlap whom: one bit set over the ranks per strip (32 strips x W words, in place of the scores) —
    // bit j of strip b: box j touches strip b.  Built 32 ranks at a time by a 32 x 32 bit transpose of the masks.
    uint32_t *strips = reinterpret_cast<uint32_t *>(raw);
    const int W = (m + 31) >> 5;  // 32 * W words <= cap + 31 words: inside the 2 * cap words of `raw` for cap >= 32
    const bool use_strips = mask_skip && 32 * W <= 2 * cap;
    if (use_strips) {
      const int lane = tid & 31;
      for (int j0 = (tid >> 5) * 32; j0 < m; j0 += BLOCK) {
        const uint32_t mj = (j0 + lane < m) ? smk[j0 + lane] : 0u;
        uint32_t mine = 0u;
#pragma unroll 8
        for (int b = 0; b < 32; b++) {
          const uint32_t w = __ballot_sync(0xffffffffu, (mj >> b) & 1u);
          if (lane == b) mine = w;
        }
        strips[lane * W + (j0 >> 5)] = mine;
      }
      __syncthreads();
    }
    // fixed-order triangular product: box j multiplies the weights of every i ranked above it — j units of work, so
    // a thread takes the boxes p and m-1-p together (m-1 units for every thread) instead of one box each
    const int n_pairs = (m + 1) >> 1;
#pragma unroll 1
    for (int u = tid; u < 2 * n_pairs; u += BLOCK) {
      // slots [0, n_pairs) = the light halves; after them the heavy halves in the same thread order (same tid for
      // BLOCK >= n_pairs; otherwise still m-1 units per consecutive pair of rounds)
      const int pq = u < n_pairs ? u : u - n_pairs;
      const int j = u < n_pairs ? pq : m - 1 - pq;
      if (u >= n_pairs && j == pq) continue;  // the middle box of an odd group was done in the first half
      const T x1 = sx1[j], y1 = sy1[j], x2 = sx2[j], y2 = sy2[j], area = sar[j];
      T live = ssc[j];
      const uint32_t mj = smk[j];
      // 32 higher ranked boxes at a time: the candidates are the boxes that share an x strip AND a y strip with
      // this one (a few percent of the pairs); only those go through the FP64 arithmetic, in rank order, so the
      // product is multiplied in the reference's order.
      for (int i0 = 0; i0 < j; i0 += 32) {
        uint32_t cand = 0xffffffffu;
        if (use_strips) {
          const uint32_t *col = strips + (i0 >> 5);
          uint32_t cx = 0u, cy = 0u;
          for (uint32_t w = mj & 0xffffu; w; w &= w - 1u) cx |= col[(__ffs(w) - 1) * W];
          for (uint32_t w = mj >> 16; w; w &= w - 1u) cy |= col[(16 + __ffs(w) - 1) * W];
          cand = cx & cy;
        } else if (mask_skip) {
          cand = 0u;
          const uint4 *mk4 = reinterpret_cast<const uint4 *>(smk + i0);  // broadcast loads; padded past m
#pragma unroll
          for (int q = 0; q < 8; q++) {
            const uint4 v = mk4[q];
            const uint32_t t0 = mj & v.x, t1 = mj & v.y, t2 = mj & v.z, t3 = mj & v.w;
            cand |= (uint32_t)((t0 & 0xffffu) != 0u && (t0 >> 16) != 0u) << (4 * q + 0);
            cand |= (uint32_t)((t1 & 0xffffu) != 0u && (t1 >> 16) != 0u) << (4 * q + 1);
            cand |= (uint32_t)((t2 & 0xffffu) != 0u && (t2 >> 16) != 0u) << (4 * q + 2);
            cand |= (uint32_t)((t3 & 0xffffu) != 0u && (t3 >> 16) != 0u) << (4 * q + 3);
          }
        }
        const int lim = j - i0;
        if (lim < 32) cand &= (1u << lim) - 1u;
        while (cand) {
          const int i = i0 + __ffs(cand) - 1;
          cand &= cand - 1u;
          bool one;
          const T wt = pair_weight(i, x1, y1, x2, y2, area, one);
          if (!one) live = live * wt;
        }
      }
      ssc[j] = live;
    }
  } else {
    // one block-wide pass per surviving rank: decay every lower ranked survivor, drop those that
    // fall below conf_thresh (also the ones the pass did not touch, like `sorted_scores.ge`)
    __syncthreads();
    for (int j = tid; j < m; j += BLOCK) alive[j] = 1;
    __syncthreads();
    for (int k = 0; k + 1 < m; k++) {
      if (!alive[k]) continue;  // uniform: shared memory, barrier-ordered
      const uint32_t mk = smk[k];
      for (int j = k + 1 + tid; j < m; j += BLOCK) {
        if (!alive[j]) continue;
        const uint32_t both = mk & smk[j];
        const bool apart = mask_skip && ((both & 0xffffu) == 0u || (both >> 16) == 0u);
        if (apart) {
          // soft branch: the score is unchanged but still has to pass `ge(conf_thresh)` (box_utils.py:379)
          if (!HARD && !(ssc[j] >= (T)conf)) alive[j] = 0;
          continue;
        }
        if (HARD) {
          // torchvision nms_kernel_impl: ovr = inter / (iarea + areas[j] - inter); suppress if > thr
          const T kx1 = sx1[k], ky1 = sy1[k], kx2 = sx2[k], ky2 = sy2[k];
          const T xx1 = kx1 > sx1[j] ? kx1 : sx1[j];
          const T yy1 = ky1 > sy1[j] ? ky1 : sy1[j];
          const T xx2 = kx2 < sx2[j] ? kx2 : sx2[j];
          const T yy2 = ky2 < sy2[j] ? ky2 : sy2[j];
          T w = xx2 - xx1, h = yy2 - yy1;
          if (!(w > (T)0)) w = (T)0;   // std::max(0, w): a NaN difference yields 0
          if (!(h > (T)0)) h = (T)0;
          const T inter = w * h;
          const T ovr = inter / ((sar[k] + sar[j]) - inter);
          if ((double)ovr > P.p.iou_thresh) alive[j] = 0;  // torchvision compares with its double threshold
        } else {
          bool one;
          const T wt = pair_weight(k, sx1[j], sy1[j], sx2[j], sy2[j], sar[j], one);
          const T live = one ? ssc[j] : ssc[j] * wt;
          ssc[j] = live;
          if (!(live >= (T)conf)) alive[j] = 0;
        }
      }
      __syncthreads();
    }
  }
  __syncthreads();

  // 4. outputs, chunk by chunk in rank order
  const int cls = (P.p.n_classes > 0) ? (g % P.p.n_classes) : 0;
  int n_kept = 0, n_ens = 0, n_trk = 0;
  for (int j0 = 0; j0 < m; j0 += BLOCK) {
    const int j = j0 + tid;
    const bool f_keep = (j < m) && (!sequential || alive[j] != 0);
    int pos = j;
    if (sequential) {
      int ek, eu, tk, tu;
      block_scan2<BLOCK>(f_keep, false, s_scan, ek, eu, tk, tu);
      pos = n_kept + ek;
      n_kept += tk;
    }
    bool f_ens = false, f_trk = false;
    int bx = 0, by = 0, bw = 0, bh = 0;
    double rs = 0.;
    if (f_keep) {
      const double x1 = (double)sx1[j], y1 = (double)sy1[j], x2 = (double)sx2[j], y2 = (double)sy2[j];
      const double live = (double)ssc[j];
      // center_size (box_utils.py:57-69) then cxcy2lxly (ensemble.py:25-28)
      const double cx = (x1 + x2) * 0.5, cy = (y1 + y2) * 0.5;
      const double w = x2 - x1, h = y2 - y1;
      const size_t o = (size_t)base + pos;
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
    if (f_ens && P.r.ens_box) {
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
  if (!sequential) n_kept = m;
  if (tid == 0 && P.r.kept_count) P.r.kept_count[g] = n_kept;
  if (tid == 0) {
    P.r.ens_count[g] = n_ens;
    if (P.r.trk_count) P.r.trk_count[g] = n_trk;
    if (P.r.img_exists && P.p.n_classes > 0 && n_ens > 0) P.r.img_exists[g / P.p.n_classes] = 1;
  }
  }  // groups
}

template <int BLOCK, bool HARD, typename T>
int launch_t(const NmsParams &P, int n_groups, size_t smem, cudaStream_t stream) {
  if (smem > 48 * 1024)
    W2T_CUDA_TRY(cudaFuncSetAttribute(softnms_kernel<BLOCK, HARD, T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
  softnms_kernel<BLOCK, HARD, T><<<n_groups, BLOCK, smem, stream>>>(P);
  W2T_CUDA_TRY(cudaGetLastError());
  return W2T_OK;
}

template <int BLOCK, bool HARD>
int launch(const NmsParams &P, int n_groups, size_t smem, cudaStream_t stream) {
  return P.p.compute_f32 ? launch_t<BLOCK, HARD, float>(P, n_groups, smem, stream)
                         : launch_t<BLOCK, HARD, double>(P, n_groups, smem, stream);
}

template <bool HARD>
int run_groups(const w2t_nms_problem_t *problem, w2t_nms_result_t *result, int max_group_size, int32_t *status,
               cudaStream_t stream, const char *who) {
  if (!problem || !result || problem->n_groups < 0 || problem->n_classes < 0 ||
      problem->n_classes > W2T_MAX_CLASSES || max_group_size < 0) {
    set_last_error("%s: bad argument", who);
    return W2T_ERR_ARG;
  }
  if (problem->n_groups == 0) return W2T_OK;
  if (!problem->group_offsets || !problem->rows || !result->ens_count || (result->ens_box && !result->ens_score) ||
      problem->box_format < W2T_BOX_LTWH || problem->box_format > W2T_BOX_LTWH_P64 || problem->top_k < 0) {
    set_last_error("%s: null buffer or bad box_format / top_k", who);
    return W2T_ERR_ARG;
  }
  if (problem->score_thr && (!result->trk_count || !result->trk_box || problem->n_classes < 1)) {
    set_last_error("%s: score_thr given without trk_count/trk_box/n_classes", who);
    return W2T_ERR_ARG;
  }
  const int smem_cap = w2t_softnms_max_group();
  NmsParams P;
  P.work = nullptr;
  P.work_stride = 0;
  P.small_cap = 0;
  P.n_lo = -1;
  P.n_hi = INT_MAX;
  P.skip_big = max_group_size > smem_cap ? 1 : 0;
  P.p = *problem;
  P.r = *result;
  P.has_thr = problem->score_thr != nullptr;
  for (int i = 0; i < W2T_MAX_CLASSES; i++)
    P.score_thr[i] = (P.has_thr && i < problem->n_classes) ? problem->score_thr[i] : 0.0;
  P.p.score_thr = nullptr;
  const int regular_max = std::min(max_group_size, smem_cap);
  P.cap = (std::max(regular_max, 1) + 3) & ~3;  // multiple of 4: keeps the int arrays 16-byte aligned
  P.status = status;
  const size_t smem = (size_t)P.cap * kBytesPerBox + 4 * kMaskPad;
  int rc;
  // many groups of mixed sizes (a Waymo-shaped job: vehicles ~200 boxes per group, pedestrians ~80, cyclists and
  // signs < 35): one launch per size class, each with CTAs and shared memory of its own size.  A CTA whose group
  // belongs to another class exits at once.  W2T_NMS_CLASSES="a,b" overrides the class bounds (debug aid; "" = one).
  int bounds[2] = {32, 96};
  // (a small job is bound by launch latency instead: one launch)
  int n_bounds = (regular_max > 64 && problem->n_groups >= 16384) ? 2 : 0;
  if (const char *e = getenv("W2T_NMS_CLASSES")) {
    n_bounds = 0;
    int a = 0, b2 = 0;
    const int got = sscanf(e, "%d,%d", &a, &b2);
    if (got >= 1 && a > 0) bounds[n_bounds++] = a;
    if (got >= 2 && b2 > a) bounds[n_bounds++] = b2;
  }
  while (n_bounds > 0 && bounds[n_bounds - 1] >= regular_max) n_bounds--;
  rc = W2T_OK;
  int lo = -1;
  for (int k = 0; k <= n_bounds && rc == W2T_OK; k++) {
    NmsParams Q = P;
    Q.n_lo = lo;
    Q.n_hi = (k < n_bounds) ? bounds[k] : INT_MAX;
    const int top = (k < n_bounds) ? bounds[k] : regular_max;   // largest group this launch serves
    Q.cap = (std::max(top, 1) + 3) & ~3;
    const size_t sm = (size_t)Q.cap * kBytesPerBox + 4 * kMaskPad;
    if (top <= 32) rc = launch<32, HARD>(Q, problem->n_groups, sm, stream);
    else if (top <= 96) rc = launch<64, HARD>(Q, problem->n_groups, sm, stream);
    else if (top <= 768) rc = launch<128, HARD>(Q, problem->n_groups, sm, stream);
    else rc = launch<256, HARD>(Q, problem->n_groups, sm, stream);
    lo = Q.n_hi;
  }
  if (rc != W2T_OK || max_group_size <= smem_cap) return rc;
  // oversized groups: the same kernel over arrays in global memory (stream-ordered scratch, freed behind the
  // launch), a few persistent CTAs striding over the groups
  const int ctas = std::min(problem->n_groups, 296);
  P.small_cap = P.cap;
  P.skip_big = 0;
  P.cap = (max_group_size + 3) & ~3;
  P.work_stride = (((size_t)P.cap * kBytesPerBox + 4 * kMaskPad) + 255) / 256 * 256;
  void *work = nullptr;
  W2T_CUDA_TRY(cudaMallocAsync(&work, P.work_stride * (size_t)ctas, stream));
  P.work = static_cast<unsigned char *>(work);
  rc = launch<256, HARD>(P, ctas, 0, stream);
  W2T_CUDA_TRY(cudaFreeAsync(work, stream));
  return rc;
}

}  // namespace

extern "C" int w2t_softnms_max_group(void) { return (kMaxSmem - 1024 - 4 * kMaskPad) / kBytesPerBox; }

extern "C" int w2t_softnms_groups(const w2t_nms_problem_t *problem, w2t_nms_result_t *result, int max_group_size,
                                  int32_t *status, w2t_stream_t stream) {
  return run_groups<false>(problem, result, max_group_size, status, (cudaStream_t)stream, "w2t_softnms_groups");
}

extern "C" int w2t_hardnms_groups(const w2t_nms_problem_t *problem, w2t_nms_result_t *result, int max_group_size,
                                  int32_t *status, w2t_stream_t stream) {
  return run_groups<true>(problem, result, max_group_size, status, (cudaStream_t)stream, "w2t_hardnms_groups");
}

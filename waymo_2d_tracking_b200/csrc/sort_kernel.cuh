// sort_kernel.cuh — the persistent SORT kernel: one CTA per (stream, category) sub-stream.
//
// Replaces, for every stream at once, the loop of tracking/track.py:42-47:
//   track_sort (tracking/utils.py:25-60) -> MultiClassTrackerSort.track
//   (tracking/sort/tracker_sort.py:22-51) -> Sort.update (tracking/sort/sort.py:244-296)
//   -> KalmanBoxTracker / associate_detections_to_trackers / iou (sort.py:33-230).
//
// A reference `Sort` object only ever sees the detections of one category of one stream, and
// the only thing sub-streams share is the global id counter (sort.py:86), which never feeds
// back into the dynamics.  So each sub-stream is an independent sequential recurrence over
// the stream's images and gets one CTA that walks them in order; ids are resolved afterwards
// from per-group creation counts (w2t_assign_ids).
//
// Per image the CTA runs:
//   A. association: -IoU cost matrix (warp per row, coalesced) -> Munkres (munkres.cuh) ->
//      matched / rejected / unassigned detections -> order of new trackers (sort.py:208-222);
//   B. one pass over the tracker list, one thread per tracker: Kalman update with the matched
//      detection or creation from an unmatched one, emission of the output row
//      (sort.py:280-289 + utils.py:37-58), age test (sort.py:292) and — for survivors — the
//      predict step of the NEXT image (sort.py:255-262), so a tracker's 56 doubles are read
//      and written once per image (20 of them: the filter is kept in block form, kalman.cuh);
//   C. stable compaction of the tracker list.
//
// Tracker state lives in a per-sub-stream slab of global memory (struct-of-arrays, one slot
// per tracker, L1/L2 resident while the CTA runs).  The tracker LIST is an array of slot
// numbers: its first T entries are the live trackers in creation order (the reference's list
// order), the rest are free slots, so removal moves 4-byte slot numbers, never filter state.
#pragma once

#include "kalman.cuh"
#include "munkres.cuh"

namespace w2t {

constexpr int kSortBlock = 128;
#ifndef W2T_MINB
#define W2T_MINB 4
#endif
constexpr int kSortMinBlocks = W2T_MINB;   // CTAs per SM the register budget is capped for
constexpr int kStateDoubles = 24;  // x[7], block-form P[13] (kalman.cuh), predicted box[4]
constexpr int kBoxAt = 20;         // first of the 4 box components

struct SlabLayout {
  size_t st, tsu, hs, bg, bk, list, tmp, flag, dstat, newdet, C, Z, rstar, cstar, rprime, ucols, crows, total;
};

// Shared-memory home of the per-image association problem.  Anything larger (crowded scenes)
// runs out of the global slab through the same pointers.
constexpr int kSmemC = 6144;     // floats: cost matrix incl. pitch (e.g. 76 x 80)
constexpr int kSmemZ = 512;      // words : zero bit matrix
constexpr int kSmemN = 128;      // rows
constexpr int kSmemM = 256;      // columns
constexpr int kSmemBox = 128;    // predicted tracker boxes staged per image
constexpr int kSmemDet = 128;    // detections of the next image prefetched while this one is tracked

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

__host__ __device__ inline SlabLayout slab_layout(int Tcap, int Dcap) {
  SlabLayout L;
  const size_t T = (size_t)Tcap, D = (size_t)Dcap;
  const size_t n = T < D ? T : D, m = T < D ? D : T;
  const size_t mcap = m < (size_t)kMunkresMaxDim ? m : (size_t)kMunkresMaxDim;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o = align_up(o + bytes, 16); return at; };
  L.st = take(sizeof(double) * kStateDoubles * T);
  L.tsu = take(4 * T);
  L.hs = take(4 * T);
  L.bg = take(4 * T);
  L.bk = take(4 * T);
  L.list = take(4 * T);
  L.tmp = take(4 * m);
  L.flag = take(4 * T);
  L.dstat = take(4 * D);
  L.newdet = take(4 * D);
  L.C = take(4 * n * (size_t)munkres_pitch((int)m));
  L.Z = take(4 * n * (size_t)munkres_zstride((int)mcap));
  L.rstar = take(4 * n);
  L.cstar = take(4 * m);
  L.rprime = take(4 * n);
  L.ucols = take(4 * m);
  L.crows = take(4 * n);
  L.total = align_up(o, 256);
  return L;
}

struct SortParams {
  w2t_sort_problem_t p;
  w2t_sort_result_t r;
  const int32_t *order, *track_cap, *det_cap;
  const int64_t *ws_offset;
  char *ws;
  int32_t *status;
  float *spill;       // warp kernel: per-team spill areas for cost matrices that outgrow tensor memory (sort_warp.cuh)
  long long *timers;  // optional [n_substreams,16] phase cycle counters (debug aid, may be NULL)
  const int32_t *chunk_of;  // optional completion tracking, see w2t_sort_plan_t
  int32_t *chunk_done;
  // frame-by-frame mode (w2t_sort_step, STEP instantiation only): [n_substreams,4] ints that carry a
  // sub-stream from one call to the next — live trackers, frame_count, "Sort object exists", flags
  // (bit 0: a predicted box is NaN, bit 1: slab initialised) — and the number the groups of this
  // call are offset by in the birth records (groups of different calls must not collide)
  int32_t *sub_state;
  int32_t group_base;
  // NumPy promotion regime of convert_bbox_to_z and of the IoU threshold test (w2t_sort_problem_t.promotion)
  int32_t nep50;
  // warp kernel (sort_warp.cuh): work queue counter, number of queue entries (P.order[0..n_items)), and the
  // per-sub-stream class flags (kCls*): classify_kernel marks the sub-streams that are too crowded for a warp
  // (kClsWide / kClsMid, by their ACTUAL detection counts - the plan may be built from upper bounds), the warp
  // kernel tracks those marked kClsWarp and flags the ones that outgrow it (kClsBailed).  A launch of the CTA
  // kernel with `bail` set serves one flag value: CTAs of every other sub-stream exit at once.
  int32_t *queue;
  int32_t n_items;
  int32_t *bail;
  int32_t bail_want;   // the flag value a CTA launch serves
};

// values of SortParams::bail[q] (aux area of the workspace): who tracks sub-stream q
enum {
  kClsWarp = 0,    // the warp kernel tracks it (sort_warp.cuh)
  kClsBailed = 1,  // outgrew the warp kernel: a 16-CTA cluster tracks it again (second pass)
  kClsWide = 2,    // very crowded (more than W2T_WIDE_DETS detections in some image): 16-CTA clusters (sort_crowd.cuh)
  kClsMid = 3,     // crowded (more than W2T_NARROW_DETS): 8-CTA clusters
  kClsHuge = 4,    // beyond the cluster kernel: CTAs with the cost matrix in global memory (sort_kernel.cuh, last pass)
  kClsDone = 5,    // tracked by a cluster
  kClsOver8 = 6    // outgrew an 8-CTA cluster: a 16-CTA cluster tracks it again (second pass)
};

// iou() of sort.py:33-47 as numba compiles it for (float32[:], float64[:]): the detection's
// own area is a float32 product, everything else float64; the result is stored as float32.
__device__ __forceinline__ float iou_pair(const float4 d, const double t0, const double t1, const double t2,
                                          const double t3) {
  const double xx1 = (double)d.x > t0 ? (double)d.x : t0;
  const double yy1 = (double)d.y > t1 ? (double)d.y : t1;
  const double xx2 = (double)d.z < t2 ? (double)d.z : t2;
  const double yy2 = (double)d.w < t3 ? (double)d.w : t3;
  double w = xx2 - xx1;
  if (!(w > 0.)) w = 0.;
  double h = yy2 - yy1;
  if (!(h > 0.)) h = 0.;
  const double wh = w * h;
  const float ad = (d.z - d.x) * (d.w - d.y);
  const double at = (t2 - t0) * (t3 - t1);
  const double den = ((double)ad + at) - wh;
  if (wh == 0. && den > 0.) return 0.0f;  // 0/den exactly; skips the FP64 divide for disjoint boxes
  return (float)(wh / den);
}

// Coarse occupancy mask of a box: bits 0-15 = the 120-pixel column strips its x extent touches,
// bits 16-31 = the 80-pixel row strips of its y extent (indices clamped to 0..15).  The strip index
// is a monotone function of the coordinate, so two boxes whose masks do not intersect in x (or
// in y) are strictly disjoint in x (or y): their IoU is exactly +0 and needs no FP64 arithmetic.
__device__ __forceinline__ uint32_t strip_mask(double x1, double y1, double x2, double y2) {
  const int a = min(max(__double2int_rd(x1 * (1.0 / 120.0)), 0), 15), b = min(max(__double2int_rd(x2 * (1.0 / 120.0)), 0), 15);
  const int c = min(max(__double2int_rd(y1 * (1.0 / 80.0)), 0), 15), d = min(max(__double2int_rd(y2 * (1.0 / 80.0)), 0), 15);
  const uint32_t mx = ((2u << b) - 1u) & ~((1u << a) - 1u);
  const uint32_t my = ((2u << d) - 1u) & ~((1u << c) - 1u);
  return (mx & 0xffffu) | (my << 16);
}

__device__ __forceinline__ double clipd(double v, double lo, double hi) {
  // numpy.clip: minimum(maximum(v, lo), hi), NaN propagates
  if (v < lo) v = lo;
  if (v > hi) v = hi;
  return v;
}

// Stable partition of the values val(i), i < n: those with fa(i) go to dst[0..na) and those
// with fb(i) to dst[na..na+nb), both in order.  dst may be the array val reads from.
template <int BLOCK, class VAL, class FA, class FB>
__device__ void partition3(int n, VAL val, FA fa, FB fb, int *dst, int *tmp, int *scratch, int &na, int &nb) {
  int ca = 0, cb = 0;
  for (int i0 = 0; i0 < n; i0 += BLOCK) {
    const int i = i0 + threadIdx.x;
    bool a = false, b = false;
    int v = 0;
    if (i < n) { v = val(i); a = fa(i); b = fb(i); }
    int ea, eb, ta, tb;
    block_scan2<BLOCK>(a, b, scratch, ea, eb, ta, tb);
    if (a) dst[ca + ea] = v;
    if (b) tmp[cb + eb] = v;
    ca += ta;
    cb += tb;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cb; i += BLOCK) dst[ca + i] = tmp[i];
  __syncthreads();
  na = ca;
  nb = cb;
}

// SMEMC: floats of the shared-memory cost matrix (larger matrices live in the slab, everything else
// of the solver stays in shared memory); with BLOCK and MINB (CTAs per SM the register budget is
// capped for) it sets how many sub-streams an SM holds at a time.
// STEP: frame-by-frame instantiation behind w2t_sort_step — the tracker state survives the launch in
// the slab and in P.sub_state, rows are emitted as Sort.update returns them (sort.py:280-289: corners
// and confidence unclipped, no size filter; utils.py:37-58 is then the caller's business).
template <int BLOCK, int MINB, bool TIMERS, int SMEMC = kSmemC, bool STEP = false>
__global__ void __launch_bounds__(BLOCK, MINB) sort_track_kernel(const SortParams P) {
  constexpr int NW = BLOCK / 32;
  __shared__ MunkresShared ms;
  __shared__ int s_scan[2 * NW];
  __shared__ int s_nan;
  __shared__ __align__(16) float s_C[SMEMC];
  __shared__ __align__(16) uint32_t s_Z[kSmemZ];
  __shared__ int s_rstar[kSmemN], s_rprime[kSmemN], s_crows[kSmemN];
  __shared__ int s_cstar[kSmemM], s_ucols[kSmemM];
  __shared__ double s_box[5][kSmemBox];  // x1, y1, x2, y2, area (NaN when the box is inverted)
  __shared__ __align__(16) float4 s_det[2][kSmemDet];  // double buffer: detections of this / the next image
  __shared__ uint32_t s_tmask[kSmemBox], s_dmask[kSmemDet];  // strip masks (all ones = "always test exactly")
  __shared__ __align__(16) uint32_t s_strip[32][4];   // per strip (0-15 x, 16-31 y): the columns whose box touches it
  __shared__ __align__(16) uint32_t s_cand[kSmemN][4];  // per row: columns that need the exact IoU
  __shared__ float s_rowmin[kSmemN];

  const int tid = threadIdx.x, lane = lane_id(), warp = warp_id();
  const int q = P.order[blockIdx.x];
  if (!STEP && P.bail != nullptr && P.bail[q] != P.bail_want) return;  // not this launch's class of sub-stream
  const int NC = P.p.n_classes;
  const int s = q / NC, c = q % NC;
  const int Tcap = P.track_cap[q], Dcap = P.det_cap[q];
  char *slab = P.ws + P.ws_offset[q];
  const SlabLayout L = slab_layout(Tcap, Dcap);
  double *st = reinterpret_cast<double *>(slab + L.st);
  int *tsuA = reinterpret_cast<int *>(slab + L.tsu);
  int *hsA = reinterpret_cast<int *>(slab + L.hs);
  int *bgA = reinterpret_cast<int *>(slab + L.bg);
  int *bkA = reinterpret_cast<int *>(slab + L.bk);
  int *list = reinterpret_cast<int *>(slab + L.list);
  int *tmp = reinterpret_cast<int *>(slab + L.tmp);
  int *flag = reinterpret_cast<int *>(slab + L.flag);   // per list position: matched det, then survivor flag
  int *dstat = reinterpret_cast<int *>(slab + L.dstat);
  int *newdet = reinterpret_cast<int *>(slab + L.newdet);

  Munkres<BLOCK, TIMERS> mk;
  mk.s = &ms;
  mk.ph = nullptr;
  mk.t_last = 0;
  MunkresGlobal g_glob, g_smem;
  g_glob.C = reinterpret_cast<float *>(slab + L.C);
  g_glob.Z = reinterpret_cast<uint32_t *>(slab + L.Z);
  g_glob.row_star = reinterpret_cast<int *>(slab + L.rstar);
  g_glob.col_star = reinterpret_cast<int *>(slab + L.cstar);
  g_glob.row_prime = reinterpret_cast<int *>(slab + L.rprime);
  g_glob.ucols = reinterpret_cast<int *>(slab + L.ucols);
  g_glob.crows = reinterpret_cast<int *>(slab + L.crows);
  g_smem.C = s_C;
  g_smem.Z = s_Z;
  g_smem.row_star = s_rstar;
  g_smem.col_star = s_cstar;
  g_smem.row_prime = s_rprime;
  g_smem.ucols = s_ucols;
  g_smem.crows = s_crows;

  const int img0 = P.p.stream_img_offsets[s], img1 = P.p.stream_img_offsets[s + 1];
  const double camW = P.p.cam_wh[2 * s], camH = P.p.cam_wh[2 * s + 1];
  // sort.py:220 `iou_matrix[m[0], m[1]] < iou_threshold`: NumPy 1.x compares the float32 entry with the python
  // float in float64, under NEP 50 the python float adopts float32
  const double thr_d = P.p.iou_thr[c];
  const float thr_f = (float)thr_d;
  const bool nep50 = P.nep50 != 0;
  const int max_age = P.p.max_age, min_hits = P.p.min_hits;
  const float4 *det_box = reinterpret_cast<const float4 *>(P.p.det_box);

  int T = 0, frame_count = 0, err = 0;
  bool started = false;
  if (STEP) {
    const int4 sv = reinterpret_cast<const int4 *>(P.sub_state)[q];
    if (!(sv.w & 2))
      for (int i = tid; i < Tcap; i += BLOCK) list[i] = i;
    else { T = sv.x; frame_count = sv.y; started = sv.z != 0; }
    if (tid == 0) s_nan = sv.w & 1;
  } else {
    for (int i = tid; i < Tcap; i += BLOCK) list[i] = i;
    if (tid == 0) { P.r.first_img[q] = -1; s_nan = 0; }
  }
  __syncthreads();

  // Software pipeline over the images: the (exists, count, start) triple of image i+1 is loaded
  // at the top of iteration i and its detections are copied to shared memory with cp.async in
  // the middle of iteration i, so that no DRAM round trip sits on the serial path.
  auto load_meta = [&](int img, int &exists, int &cnt, int &start) {
    exists = (P.p.img_exists == nullptr) ? 1 : (int)P.p.img_exists[img];
    cnt = P.p.det_count[img * NC + c];
    start = P.p.det_start[img * NC + c];
  };
  auto prefetch_dets = [&](int buf, int exists, int cnt, int start) {
    if (exists && cnt <= kSmemDet) {
      for (int i = tid; i < cnt; i += BLOCK) {
        const unsigned dst = (unsigned)__cvta_generic_to_shared(&s_det[buf][i]);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(det_box + start + i) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int cur_exists = 0, cur_cnt = 0, cur_start = 0, buf = 0;
  if (img0 < img1) {
    load_meta(img0, cur_exists, cur_cnt, cur_start);
    prefetch_dets(0, cur_exists, cur_cnt, cur_start);
  }
  long long ph[TIMERS ? 16 : 1];
  if (TIMERS) {
#pragma unroll
    for (int i = 0; i < (TIMERS ? 16 : 1); i++) ph[i] = 0;
    mk.ph = ph;
    mk.t_last = clock64();
  }
#define W2T_TICK(i) mk.tick(i)

  for (int img = img0; img < img1; ++img, buf ^= 1) {
    const int g = img * NC + c;
    const int this_exists = cur_exists, this_cnt = cur_cnt, this_start = cur_start;
    bool next_issued = true;
    if (img + 1 < img1) {
      load_meta(img + 1, cur_exists, cur_cnt, cur_start);  // in flight while this image is tracked
      next_issued = false;
    }
    // every path through this iteration ends up issuing the prefetch of the next image once
    auto issue_next = [&]() {
      if (!next_issued) {
        prefetch_dets(buf ^ 1, cur_exists, cur_cnt, cur_start);
        next_issued = true;
      }
    };
    bool skip = (this_exists == 0);
    const int D = skip ? 0 : this_cnt;
    if (!skip && !started) {
      if (D == 0) skip = true;  // no Sort object for this category yet (tracker_sort.py:32-33)
      else {
        started = true;
        if (!STEP && tid == 0) P.r.first_img[q] = img - img0;
      }
    }
    if (!skip && !err && (D > Dcap || D > kMunkresMaxDim || T > kMunkresMaxDim)) err = W2T_ERR_CAPACITY;
    if (skip || err) {
      if (tid == 0) { P.r.out_count[g] = 0; P.r.created[g] = 0; }
      issue_next();
      continue;
    }
    frame_count++;
    const int base = this_start;
    // this image's detections: prefetched into shared memory unless there are too many
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    const float4 *dets = (D <= kSmemDet) ? s_det[buf] : det_box + base;

    // ---- trackers whose predicted box is NaN are dropped before association (sort.py:261-265)
    if (s_nan) {
      int na, nb;
      partition3<BLOCK>(
          T, [&](int i) { return list[i]; },
          [&](int i) {
            const int sl = list[i];
            return !(isnan(st[(kBoxAt + 0) * Tcap + sl]) || isnan(st[(kBoxAt + 1) * Tcap + sl]) || isnan(st[(kBoxAt + 2) * Tcap + sl]) ||
                     isnan(st[(kBoxAt + 3) * Tcap + sl]));
          },
          [&](int i) {
            const int sl = list[i];
            return isnan(st[(kBoxAt + 0) * Tcap + sl]) || isnan(st[(kBoxAt + 1) * Tcap + sl]) || isnan(st[(kBoxAt + 2) * Tcap + sl]) ||
                   isnan(st[(kBoxAt + 3) * Tcap + sl]);
          },
          list, tmp, s_scan, na, nb);
      T = na;
      if (tid == 0) s_nan = 0;
      __syncthreads();
    }

    W2T_TICK(1);
    // ---- A. association ------------------------------------------------------------------
    for (int t = tid; t < T; t += BLOCK) flag[t] = -1;
    if (T > 0 && D > 0) {
      const bool flipped = D > T;  // the solver transposes when there are more rows than columns
      const int n = flipped ? T : D, m = flipped ? D : T;
      mk.n = n;
      mk.m = m;
      mk.mw = munkres_words(m);
      mk.zs = munkres_zstride(m);
      mk.ldc = munkres_pitch(m);
      // masks, stars and index lists of the solver live in shared memory whenever the problem is at
      // most 128 x 128; the cost matrix itself joins them if it is small enough, else it stays in the
      // slab (L1/L2) and the same solver runs on it
      const bool fits = (n * mk.zs <= kSmemZ) && (n <= kSmemN) && (m <= 128);
      mk.g = fits ? g_smem : g_glob;
      if (fits && n * mk.ldc > SMEMC) mk.g.C = g_glob.C;
      mk.rowwise = fits;
      // predicted boxes of the live trackers, by list position
      const bool boxes_staged = T <= kSmemBox;
      if (boxes_staged) {
        for (int t = tid; t < T; t += BLOCK) {
          const int sl = list[t];
          double b[4];
#pragma unroll
          for (int k = 0; k < 4; k++) { b[k] = st[(kBoxAt + k) * Tcap + sl]; s_box[k][t] = b[k]; }
          const double area = (b[2] > b[0] && b[3] > b[1]) ? (b[2] - b[0]) * (b[3] - b[1]) : NAN;
          s_box[4][t] = area;
          s_tmask[t] = (area > 0.) ? strip_mask(b[0], b[1], b[2], b[3]) : 0xffffffffu;
        }
      }
      const bool masked = boxes_staged && D <= kSmemDet;
      if (masked) {
        for (int d = tid; d < D; d += BLOCK) {
          const float4 b = dets[d];
          // regular detection: positive extent (its float32 area is then >= +0, sort.py:40-46)
          s_dmask[d] = (b.z > b.x && b.w > b.y) ? strip_mask((double)b.x, (double)b.y, (double)b.z, (double)b.w)
                                                : 0xffffffffu;
        }
      }
      if (boxes_staged) __syncthreads();
      auto tbox = [&](int t, double &t0, double &t1, double &t2, double &t3, double &at) {
        if (boxes_staged) {
          t0 = s_box[0][t]; t1 = s_box[1][t]; t2 = s_box[2][t]; t3 = s_box[3][t]; at = s_box[4][t];
        } else {
          const int sl = list[t];
          t0 = st[(kBoxAt + 0) * Tcap + sl]; t1 = st[(kBoxAt + 1) * Tcap + sl];
          t2 = st[(kBoxAt + 2) * Tcap + sl]; t3 = st[(kBoxAt + 3) * Tcap + sl];
          at = (t2 > t0 && t3 > t1) ? (t2 - t0) * (t3 - t1) : NAN;
        }
      };
      // -IoU.  Disjoint boxes (the vast majority) are settled by four compares: their IoU is
      // 0 / (area_det + area_trk - 0) = +0 whenever that denominator is positive; anything
      // unusual (NaN, inverted box, non-positive denominator) takes the full iou_pair path.
      auto neg_iou = [&](const float4 d, double t0, double t1, double t2, double t3, double at) -> float {
        const bool apart = !((double)d.z > t0 && t2 > (double)d.x && (double)d.w > t1 && t3 > (double)d.y);
        if (apart && d.z > d.x && d.w > d.y) {
          const float ad = (d.z - d.x) * (d.w - d.y);
          if ((double)ad + at > 0.) return -0.0f;
        }
        return -iou_pair(d, t0, t1, t2, t3);
      };
      const bool fused = mk.block_path() && masked;  // step 1 of the solver is folded into the construction of the matrix
      if (fused) {
        // Block path (everything but crowded scenes; boxes and detections are staged, n, m <= 128).
        // Almost every pair is strictly disjoint and costs -0.0f, so the matrix is built from the few
        // pairs that are not:
        //   1. every column enters the bit sets of the strips its box touches (s_strip);
        //   2. one thread per row: candidate columns = (union of the sets of the row's x strips) AND
        //      (union over its y strips) - the same predicate as "the strip masks meet in x and in y";
        //      exact cost of each candidate, row minimum (step 1 of the solver, munkres.cuh);
        //   3. one thread per (row, 32 columns): reduced costs and the zero bit word.
        const uint32_t *rmask = flipped ? s_tmask : s_dmask, *cmask = flipped ? s_dmask : s_tmask;
        for (int i = tid; i < 32 * 4; i += BLOCK) (&s_strip[0][0])[i] = 0u;
        __syncthreads();
        for (int cc = tid; cc < m; cc += BLOCK) {
          uint32_t w = cmask[cc];
          const uint32_t bit = 1u << (cc & 31);
          while (w) {
            const int b = __ffs(w) - 1;
            w &= w - 1u;
            atomicOr(&s_strip[b][cc >> 5], bit);
          }
        }
        __syncthreads();
        for (int r = tid; r < n; r += BLOCK) {
          uint4 cx = make_uint4(0u, 0u, 0u, 0u), cy = cx;
          const uint32_t rm = rmask[r];
          for (uint32_t w = rm & 0xffffu; w; w &= w - 1u) {
            const uint4 v = *reinterpret_cast<const uint4 *>(s_strip[__ffs(w) - 1]);
            cx.x |= v.x; cx.y |= v.y; cx.z |= v.z; cx.w |= v.w;
          }
          for (uint32_t w = rm >> 16; w; w &= w - 1u) {
            const uint4 v = *reinterpret_cast<const uint4 *>(s_strip[16 + __ffs(w) - 1]);
            cy.x |= v.x; cy.y |= v.y; cy.z |= v.z; cy.w |= v.w;
          }
          const uint4 cand = make_uint4(cx.x & cy.x, cx.y & cy.y, cx.z & cy.z, cx.w & cy.w);
          *reinterpret_cast<uint4 *>(s_cand[r]) = cand;
          float *row = mk.g.C + (size_t)r * mk.ldc;
          // minimum on the order-preserving integer image of the float (NaN sorts last, like fminf
          // ignores it); every column that is not a candidate contributes -0.0f
          const int ncand = __popc(cand.x) + __popc(cand.y) + __popc(cand.z) + __popc(cand.w);
          uint32_t mn_u = (ncand < m) ? Munkres<BLOCK, TIMERS>::ordered(-0.0f) : 0xffffffffu;
          int k = 0;
          uint32_t w = cand.x;
          for (;;) {
            while (w == 0u && ++k < 4) w = (k == 1) ? cand.y : (k == 2) ? cand.z : cand.w;
            if (w == 0u) break;
            const int cc = k * 32 + __ffs(w) - 1;
            w &= w - 1u;
            const int di = flipped ? cc : r, ti = flipped ? r : cc;
            double t0, t1, t2, t3, at;
            tbox(ti, t0, t1, t2, t3, at);
            const float v = neg_iou(dets[di], t0, t1, t2, t3, at);
            row[cc] = v;
            mn_u = min(mn_u, Munkres<BLOCK, TIMERS>::ordered(v));
          }
          s_rowmin[r] = Munkres<BLOCK, TIMERS>::unordered(mn_u);
          mk.g.row_star[r] = -1;
          mk.g.row_prime[r] = -1;
        }
        __syncthreads();
        const int mw = mk.mw;
        for (int it = tid; it < n * mw; it += BLOCK) {
          const int r = it / mw, k = it - r * mw;
          const uint32_t cw = s_cand[r][k];
          const float mn = s_rowmin[r];
          const int nb = min(32, m - k * 32);
          float *row = mk.g.C + (size_t)r * mk.ldc + k * 32;
          uint32_t z = 0u;
          if (cw == 0u) {
            const float bg = -0.0f - mn;
            for (int b = 0; b < nb; b++) row[b] = bg;
            if (bg == 0.0f) z = (nb == 32) ? 0xffffffffu : ((1u << nb) - 1u);
          } else {
#pragma unroll 4
            for (int b = 0; b < nb; b++) {
              const float raw = ((cw >> b) & 1u) ? row[b] : -0.0f;
              const float v = raw - mn;
              row[b] = v;
              z |= (v == 0.0f) ? (1u << b) : 0u;
            }
          }
          mk.g.Z[(size_t)r * mk.zs + k] = z;
        }
      } else {
        for (int r = warp; r < n; r += NW) {
          float *row = mk.g.C + (size_t)r * mk.ldc;
          if (!flipped) {
            const float4 d = dets[r];
            for (int cc = lane; cc < m; cc += 32) {
              double t0, t1, t2, t3, at;
              tbox(cc, t0, t1, t2, t3, at);
              row[cc] = neg_iou(d, t0, t1, t2, t3, at);
            }
          } else {
            double t0, t1, t2, t3, at;
            tbox(r, t0, t1, t2, t3, at);
            for (int cc = lane; cc < m; cc += 32) row[cc] = neg_iou(dets[cc], t0, t1, t2, t3, at);
          }
        }
      }
      __syncthreads();
      W2T_TICK(2);
      if (TIMERS && tid == 0) ph[13]++;
      if (mk.solve(fused) != 0) err = W2T_ERR_ARG;
      for (int d = tid; d < D; d += BLOCK) {
        const int t = flipped ? mk.g.col_star[d] : mk.g.row_star[d];
        int stt = 0;
        if (t >= 0) {
          double t0, t1, t2, t3, at;
          tbox(t, t0, t1, t2, t3, at);
          const float o = iou_pair(dets[d], t0, t1, t2, t3);
          if (nep50 ? (o < thr_f) : ((double)o < thr_d)) stt = 2;  // assigned but rejected: becomes a new tracker AFTER the unassigned ones
          else { stt = 1; flag[t] = d; }
        }
        dstat[d] = stt;
      }
    } else {
      for (int d = tid; d < D; d += BLOCK) dstat[d] = 0;
    }
    __syncthreads();
    W2T_TICK(7);
    issue_next();  // by now the next image's (count, start) have arrived
    int n_un, n_rej;
    partition3<BLOCK>(
        D, [&](int i) { return i; }, [&](int i) { return dstat[i] == 0; }, [&](int i) { return dstat[i] == 2; },
        newdet, tmp, s_scan, n_un, n_rej);
    W2T_TICK(8);
    const int n_new = n_un + n_rej;
    const int Ttot = T + n_new;
    if (Ttot > Tcap) err = W2T_ERR_CAPACITY;
    if (err) {
      if (tid == 0) { P.r.out_count[g] = 0; P.r.created[g] = 0; }
      continue;
    }

    // ---- B. one pass over the tracker list --------------------------------------------------
    int emitted = 0;
    for (int t0 = 0; t0 < Ttot; t0 += BLOCK) {
      const int t = t0 + tid;
      bool ok = false;
      double ob0 = 0, ob1 = 0, ob2 = 0, ob3 = 0, oconf = 0;
      int obg = 0, obk = 0;
      if (t < Ttot) {
        const int sl = list[t];
        double x[7], Pm[kBlockP];
        int tsu, hs;
        if (t >= T) {  // sort.py:276-278
          const float4 d4 = dets[newdet[t - T]];
          const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
          kfb_init(dd, x, Pm, nep50);
          tsu = 0;
          hs = 0;
          obg = STEP ? g + P.group_base : g;
          obk = t - T;
          bgA[sl] = obg;
          bkA[sl] = obk;
        } else {
#pragma unroll
          for (int k = 0; k < 7; k++) x[k] = st[k * Tcap + sl];
#pragma unroll
          for (int k = 0; k < kBlockP; k++) Pm[k] = st[(7 + k) * Tcap + sl];
          tsu = tsuA[sl];
          hs = hsA[sl];
          obg = bgA[sl];
          obk = bkA[sl];
          const int md = flag[t];
          if (md >= 0) {  // sort.py:270-273, :153-164
            const float4 d4 = dets[md];
            const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
            kfb_update(x, Pm, dd, nep50);
            tsu = 0;
            hs += 1;
          }
        }
        // sort.py:281-289 and utils.py:37-49
        if (tsu < 1 && (hs >= min_hits || frame_count <= min_hits)) {
          double b[4];
          x_to_bbox(x, b);
          const double e = ((Pm[0] + Pm[4]) + Pm[8]) / 3.0;  // mean(P00, P11, P22), sort.py:190
          const double conf = exp(-e * 0.1);
          if (STEP) {
            ok = true;
            ob0 = b[0]; ob1 = b[1]; ob2 = b[2]; ob3 = b[3];
            oconf = conf;
          } else {
            const double x1 = clipd(b[0], 0., camW), y1 = clipd(b[1], 0., camH);
            const double x2 = clipd(b[2], 0., camW), y2 = clipd(b[3], 0., camH);
            const double wd = x2 - x1, ht = y2 - y1;
            if (!(wd < 1 || ht < 1)) {
              ok = true;
              ob0 = x1; ob1 = y1; ob2 = wd; ob3 = ht;
              oconf = clipd(conf, 0.2, 1.0);
            }
          }
        }
        const bool surv = !(tsu > max_age);  // sort.py:292
        if (surv) {
          // predict of the next image (sort.py:166-178)
          kfb_predict(x, Pm);
          if (tsu > 0) hs = 0;
          tsu += 1;
          double b[4];
          x_to_bbox(x, b);
          if (isnan(b[0]) || isnan(b[1]) || isnan(b[2]) || isnan(b[3])) s_nan = 1;
          else if (isinf(b[0]) || isinf(b[1]) || isinf(b[2]) || isinf(b[3])) atomicMax(P.status, W2T_ERR_NONFINITE);
#pragma unroll
          for (int k = 0; k < 7; k++) st[k * Tcap + sl] = x[k];
#pragma unroll
          for (int k = 0; k < kBlockP; k++) st[(7 + k) * Tcap + sl] = Pm[k];
#pragma unroll
          for (int k = 0; k < 4; k++) st[(kBoxAt + k) * Tcap + sl] = b[k];
          tsuA[sl] = tsu;
          hsA[sl] = hs;
        }
        flag[t] = surv ? 1 : 0;
      }
      int ex, exb, tot, totb;
      block_scan2<BLOCK>(ok, false, s_scan, ex, exb, tot, totb);
      if (ok) {
        const size_t o = (size_t)base + emitted + ex;
        double *ob = P.r.out_box + 4 * o;
        ob[0] = ob0; ob[1] = ob1; ob[2] = ob2; ob[3] = ob3;
        P.r.out_score[o] = oconf;
        P.r.out_birth[2 * o + 0] = obg;
        P.r.out_birth[2 * o + 1] = obk;
      }
      emitted += tot;
    }
    if (tid == 0) { P.r.out_count[g] = emitted; P.r.created[g] = n_new; }
    __syncthreads();
    W2T_TICK(9);

    // ---- C. drop dead trackers, keep the list order (sort.py:292-293) --------------------------
    int n_live, n_dead;
    partition3<BLOCK>(
        Ttot, [&](int i) { return list[i]; }, [&](int i) { return flag[i] != 0; },
        [&](int i) { return flag[i] == 0; }, list, tmp, s_scan, n_live, n_dead);
    T = n_live;
    W2T_TICK(10);
  }
  if (TIMERS && P.timers != nullptr && tid == 0) {
    ph[0] = frame_count;
    for (int i = 0; i < (TIMERS ? 16 : 1); i++) P.timers[(size_t)q * 16 + i] = ph[i];
  }
#undef W2T_TICK

  if (err && tid == 0) atomicMax(P.status, err);
  if (STEP) {
    __syncthreads();
    if (tid == 0) reinterpret_cast<int4 *>(P.sub_state)[q] = make_int4(T, frame_count, started ? 1 : 0, (s_nan ? 1 : 0) | 2);
    if (P.r.first_img != nullptr) {
      // frame-by-frame mode: the oldest birth group any live tracker of the sub-stream refers to (INT_MAX: none),
      // so that the caller can forget the id bases of older calls
      int oldest = 0x7fffffff;
      for (int t = tid; t < T; t += BLOCK) oldest = min(oldest, bgA[list[t]]);
      oldest = __reduce_min_sync(0xffffffffu, oldest);
      __syncthreads();
      if ((tid & 31) == 0) s_scan[tid >> 5] = oldest;
      __syncthreads();
      if (tid == 0) {
        for (int w = 1; w < BLOCK / 32; w++) oldest = min(oldest, s_scan[w]);
        P.r.first_img[q] = oldest;
      }
      __syncthreads();
    }
  }

  // optional: filter state of every live tracker, already predicted one step past the last image
  if (P.r.final_count != nullptr) {
    if (tid == 0) P.r.final_count[q] = T;
    if (P.r.final_state != nullptr) {
      const int cap = P.r.final_cap;
      for (int t = tid; t < T && t < cap; t += BLOCK) {
        const int sl = list[t];
        double *dst = P.r.final_state + ((size_t)q * cap + t) * 56;
        double pb[kBlockP], Pd[49];
        for (int k = 0; k < 7; k++) dst[k] = st[k * Tcap + sl];
        for (int k = 0; k < kBlockP; k++) pb[k] = st[(7 + k) * Tcap + sl];
        kfb_to_dense(pb, Pd);
        for (int k = 0; k < 49; k++) dst[7 + k] = Pd[k];
      }
    }
  }

  // completion tracking: everything this CTA wrote is visible before its chunk's counter moves
  if (P.chunk_done != nullptr) {
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      atomicAdd(&P.chunk_done[P.chunk_of[q]], 1);
    }
  }
}

}  // namespace w2t

// sort_warp.cuh — the SORT tracker with one WARP (or a pair of warps) per (stream, category) sub-stream.
//
// Replaces, like sort_kernel.cuh, the loop of tracking/track.py:42-47 (track_sort, tracking/utils.py:25-60 ->
// MultiClassTrackerSort.track, tracking/sort/tracker_sort.py:22-51 -> Sort.update, tracking/sort/sort.py:244-296
// -> KalmanBoxTracker / associate_detections_to_trackers / iou, sort.py:33-230, with the Munkres solver of
// scikit-learn 0.22.2 behind sort.py:206) — for every sub-stream whose images hold at most kWarpDim
// detections and live trackers, which is everything but crowded scenes.
//
// Why warps.  A sub-stream is a serial recurrence over its images, and inside an image the solver is a serial
// state machine; the CTA-per-sub-stream kernel (sort_kernel.cuh) spent its time at CTA barriers (three of four
// warps waiting for the one that drives the solver) and, above all, on instruction fetch: its per-image path
// is ~50 KB of SASS against a 32 KB instruction cache per SM, shared by CTAs in different phases
// (profiles/r01h_*: i-cache hit rate 73 %, GPC instruction requests at 96 % of peak, issue slots 25 %).  Here:
//   * no CTA barrier on the path — phases are separated by __syncwarp(), or by a 64-thread named barrier
//     where two warps share a sub-stream;
//   * every cover / star mask of the solver is ONE register per lane (lane k holds bits 32k..32k+31), so a
//     mask operation is a predicated instruction, "first set bit" a ballot + shuffle, and the code has no
//     per-word unrolling: the whole per-image path fits the instruction cache (r02: hit rate 94-98 %);
//   * the cost matrix lives in TENSOR MEMORY (below), eight rows per access;
//   * trackers are kept in LIST ORDER (the reference's list, sort.py:238): the pass that updates / predicts
//     them writes each survivor back at its new position, so removal (sort.py:292-293) costs nothing and the
//     accesses are coalesced; the predicted boxes go straight to shared memory for the next image;
//   * warps are persistent and pull sub-streams from two queues in the plan's order (heaviest first).
// Twelve warps per SM.  The job is bounded by its longest chains (the most crowded streams: 200 images one
// after the other), so those get two warps: warps 0-3 ("leaders") serve the queue of crowded sub-streams,
// each with a helper (warps 4-7) that takes every other block of rows in the wide phases — construction
// of the cost matrix, the cost shifts of the solver (step 6), the Kalman pass — while the leader alone
// drives the solver's serial steps; warps 8-11 serve the queue of light sub-streams on their own.
// A sub-stream that outgrows the kernel (more than kWarpDim detections or live trackers, or a cost matrix
// beyond its share of tensor memory) raises its flag and is tracked again from its first image by the CTA
// kernel (sort.cu launches it behind this one; CTAs of all other sub-streams exit at once).
#pragma once

#include "sort_kernel.cuh"

namespace w2t {

constexpr int kWarpDim = 128;       // most detections / live trackers per image
constexpr int kWarpZS = 5;          // words per row of the zero bit matrix (4 + 1: odd stride)
constexpr int kWarpSmemBytes = 232448;  // 227 KB: the most dynamic shared memory a CTA can have
constexpr int kTeamsPerCta = 8;     // sub-streams in flight per SM: 4 pairs + 4 single warps
constexpr int kWarpsPerCta = 12;
constexpr int kBigCols = 384, kSmallCols = 128;  // TMEM columns (= cell words of the cost matrix) of a pair / a single warp
constexpr int kSpillFloats = 512 * 32;  // a team's spill area: the largest matrix (ceil8(128) rows x 4 column words)
constexpr int kSpillCtas = 160;          // CTAs the spill area is sized for (the grid is one CTA per SM, at most this)
constexpr int kSpillImages = 32;         // a sub-stream may spill at this many images and still be served by warps
constexpr int kSpillDets = 104;          // ... if it never holds more detections than this (its trackers: ~1.25x)
constexpr int kClassifyHuge = 896;  // = kCrowdM (sort_crowd.cuh): more detections than the cluster kernel takes

// The cost matrix of a team's current image lives in TENSOR MEMORY (256 KB per SM, 128 lanes x 512 columns x
// 32 bit, reached with tcgen05.ld / tcgen05.st; SASS: LDTM / STTM).  A warp can only touch the 32 TMEM lanes of
// its quarter (warp index % 4), which is exactly the shape this solver wants: TMEM lane = matrix column mod 32,
// and one TMEM column holds one "cell word" (32 consecutive matrix columns of one row), so a single
// 32x32b.x8 access moves the same 32 matrix columns of eight rows between TMEM and eight registers per lane —
// one instruction and ~12 cycles where shared memory needs eight instructions and ~30, and the 18-40 KB per
// sub-stream the matrix would take in shared memory are free for more sub-streams.  Cell word (row r, word k)
// of an n x m problem sits in column k * ceil8(n) + r of the team's share.  The 512 columns of a quarter are
// split between the teams whose warps can reach it: 384 for the pair (e.g. a 96 x 128 problem), 128 for the
// single warp (e.g. 40 x 64).
struct __align__(128) WarpShared {
  float4 det[kWarpDim];                 // this image's detections
  double box[4][kWarpDim];              // predicted boxes of the live trackers, by list position
  uint32_t Z[kWarpDim * kWarpZS];       // zero bit matrix (while the matrix is built: candidate columns per row)
  uint32_t strip[32][4];                // per strip (0-15 x, 16-31 y): the columns whose box touches it
  float rowmin[kWarpDim];
  int8_t row_star[kWarpDim], row_prime[kWarpDim], col_star[kWarpDim];
  int8_t match[kWarpDim];               // per tracker: matched detection or -1
  int8_t dstat[kWarpDim];               // per detection: 0 unassigned, 1 matched, 2 assigned but rejected
  int8_t newdet[kWarpDim];              // detections that become trackers, in the reference's order
  float raw[32 * kWarpDim];             // exact costs of the candidate pairs of the 32 rows being built
  // mailbox of a pair (leader -> helper, and partial results)
  uint32_t cov[2][4];                   // cost shift: row cover / column cover words
  uint32_t part[2];                     // cost shift: minimum found by each warp
  int32_t cmd;                          // inside the solver: 1 = shift the costs, 0 = solved, 9 = gave up
  int32_t cur_q;                        // sub-stream the team works on, -1 = none left
  int32_t n_new;                        // trackers created at this image
  int32_t cnt[2][2][2];                 // Kalman pass: [trip parity][warp of the pair][survivors, rows emitted]
};
static_assert(sizeof(WarpShared) * kTeamsPerCta <= kWarpSmemBytes, "shared memory of the warp kernel exceeds 227 KB");

// eight TMEM columns <-> eight registers per lane (lane l talks to TMEM lane l of the warp's quarter)
__device__ __forceinline__ void tm_ld8(const uint32_t taddr, float (&v)[8]) {
  uint32_t u[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 8; j++) v[j] = __uint_as_float(u[j]);
}
__device__ __forceinline__ void tm_st8(const uint32_t taddr, const float (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
               "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
               "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// The cost matrix of one image lives in the team's share of tensor memory — or, for the rare image whose matrix
// outgrows that share (ceil8(n) * ceil32(m) / 32 columns > kBigCols / kSmallCols), in the team's spill area in
// global memory (L2), laid out the same way: column c of lane l at spill[c * 32 + l].
// SPILL is a compile-time property of the code path: the common path carries no trace of the other one.
template <bool SPILL>
struct Cells {
  uint32_t tm;
  float *spill;
  __device__ __forceinline__ void ld8(const int col, float (&v)[8]) const {
    if (SPILL) {
      const float *p = spill + col * 32 + lane_id();
#pragma unroll
      for (int j = 0; j < 8; j++) v[j] = __ldcg(p + j * 32);
    } else {
      tm_ld8(tm + (uint32_t)col, v);
    }
  }
  __device__ __forceinline__ void st8(const int col, const float (&v)[8]) const {
    if (SPILL) {
      float *p = spill + col * 32 + lane_id();
#pragma unroll
      for (int j = 0; j < 8; j++) __stcg(p + j * 32, v[j]);
    } else {
      tm_st8(tm + (uint32_t)col, v);
    }
  }
  __device__ __forceinline__ void wait_st() const {
    if (!SPILL) tm_wait_st();  // global stores are ordered by the team's barrier
  }
};

// A team: one warp, or a pair (leader = half 0, helper = half 1) that meets at a 64-thread named barrier.
struct Team {
  int half, nh, bar;
  __device__ __forceinline__ void sync() const {
    if (nh == 2) {
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("bar.sync %0, 64;" ::"r"(bar) : "memory");
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    } else {
      __syncwarp();
    }
  }
};

// sqrt / divide of convert_x_to_bbox (sort.py:70-71), out of line: called twice per tracker and image
__device__ __noinline__ void box_wh(const double s, const double r, double &w, double &h) {
  w = sqrt(s * r);
  h = s / w;
}

__device__ __noinline__ float iou_pair_call(const float4 d, const double t0, const double t1, const double t2,
                                            const double t3) {
  return iou_pair(d, t0, t1, t2, t3);
}

// strip mask of a box, all ones ("always test exactly") unless the box has positive extent
__device__ __forceinline__ uint32_t strip_mask_checked(const double x1, const double y1, const double x2, const double y2) {
  const bool regular = (x2 > x1) && (y2 > y1) && ((x2 - x1) * (y2 - y1) > 0.);
  return regular ? strip_mask(x1, y1, x2, y2) : 0xffffffffu;
}

// ---- step 6 of the solver, by the whole team: min over uncovered rows x uncovered columns; covered rows += min,
// then uncovered columns -= min (float32, in that order).  The covers come from S.cov.  Lanes own columns (bit k
// of `uc`: column 32k + lane is uncovered, bit k of `ucw`: word k has an uncovered column at all); rows go eight
// at a time, one TMEM access per eight cell words, the blocks of eight dealt out between the warps of the team
// (the matrix holds ceil8(n) rows: the padding rows are computed on and never read).
template <bool SPILL>
__device__ __forceinline__ void cost_shift(WarpShared &S, const Cells<SPILL> tm, const int n, const int m, const Team &team) {
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = lane_id();
  const int mw = (m + 31) >> 5, NP = (n + 7) & ~7;
  uint32_t uc = 0u, ucw = 0u;
#pragma unroll 1
  for (int k = 0; k < mw; k++) {
    const uint32_t cw = S.cov[1][k];
    if (k * 32 + lane < m && !((cw >> lane) & 1u)) uc |= 1u << k;
    const uint32_t valid = (m - k * 32 >= 32) ? 0xffffffffu : ((1u << (m - k * 32)) - 1u);
    if (~cw & valid) ucw |= 1u << k;
  }
  // costs are >= +0 here (or NaN), so their bit patterns order like the values and NaN sorts above +inf
  uint32_t mn_u = 0x7f800000u;
#pragma unroll 1
  for (int r0 = team.half * 8; r0 < n; r0 += 8 * team.nh) {
    uint32_t ur = ~(S.cov[0][r0 >> 5] >> (r0 & 31));  // bit j: row r0 + j is uncovered
    if (n - r0 < 8) ur &= (1u << (n - r0)) - 1u;
    if ((ur & 0xffu) == 0u) continue;
#pragma unroll 1
    for (uint32_t w = ucw; w; w &= w - 1u) {
      const int k = __ffs(w) - 1;
      float v[8];
      tm.ld8(k * NP + r0, v);
      if ((uc >> k) & 1u) {
#pragma unroll
        for (int j = 0; j < 8; j++)
          if ((ur >> j) & 1u) mn_u = min(mn_u, __float_as_uint(v[j]));
      }
    }
  }
  mn_u = __reduce_min_sync(FULL, mn_u);
  if (team.nh == 2) {
    if (lane == 0) S.part[team.half] = mn_u;
    team.sync();
    mn_u = min(S.part[0], S.part[1]);
  }
  if (mn_u != 0x7f800000u) {  // nothing uncovered: the reference leaves the matrix alone
    const float mn = __uint_as_float(mn_u);
#pragma unroll 1
    for (int r0 = team.half * 8; r0 < n; r0 += 8 * team.nh) {
      const uint32_t cr = (S.cov[0][r0 >> 5] >> (r0 & 31)) & 0xffu;  // bit j: row r0 + j is covered
      uint32_t *zp = S.Z + r0 * kWarpZS;
#pragma unroll 1
      for (int k = 0; k < mw; k++) {
        if (cr == 0u && !((ucw >> k) & 1u)) continue;  // uncovered rows x covered columns: nothing changes
        const bool colv = k * 32 + lane < m;
        const bool u = (uc >> k) & 1u;
        float v[8];
        tm.ld8(k * NP + r0, v);
#pragma unroll
        for (int j = 0; j < 8; j++) {
          if ((cr >> j) & 1u) v[j] = v[j] + mn;
          if (u) v[j] = v[j] - mn;
          const uint32_t word = __ballot_sync(FULL, colv && v[j] == 0.0f);
          if (lane == 0) zp[j * kWarpZS + k] = word;
        }
        tm.st8(k * NP + r0, v);
      }
    }
    tm.wait_st();
  }
  team.sync();
}

// ---- scikit-learn 0.22.2 linear_assignment on the n x m (n <= m <= kWarpDim) matrix in tensor memory at `tm`
// (cell word (r, k) in column k * NP + r, NP = ceil8(n)), whose row minima are already subtracted and whose zeros
// are mirrored in S.Z (step 1 is folded into the construction of the matrix).  Same step machine, same decisions
// as munkres.cuh (see there for why each shortcut is exact).  Runs in the team's leading warp; lane k holds
// word k of every mask.  A helper warp follows in solver_helper().  Returns 0, or 9 when the iteration budget
// ran out (NaN costs).
template <bool TIMERS, bool SPILL>
__device__ __forceinline__ int warp_munkres(WarpShared &S, const Cells<SPILL> tm, const int n, const int m, const Team &team,
                                            long long *ph) {
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = lane_id();
  const unsigned lt = (1u << lane) - 1u;
  const int mw = (m + 31) >> 5, nwr = (n + 31) >> 5;
  uint32_t starcols = 0u, colcov = 0u, rowcov = 0u, rowhas = 0u;
  int stars = 0, act = 0;
  int budget = 4 * n * n + 64 * (n + m) + 1024;

  // ---- step 2: greedy stars in row-major order; the longest prefix of pending rows whose proposals are
  // pairwise distinct is exactly what the sequential loop would star
#pragma unroll 1
  for (int rb = 0; rb < n; rb += 32) {
    const int r = rb + lane;
    bool pending = r < n;
#pragma unroll 1
    for (;;) {
      int cand = -1;
#pragma unroll 1
      for (int k = 0; k < mw; k++) {
        const uint32_t sc = __shfl_sync(FULL, starcols, k);
        if (pending && cand < 0) {
          const uint32_t v = S.Z[r * kWarpZS + k] & ~sc;
          if (v) cand = k * 32 + __ffs(v) - 1;
        }
      }
      if (cand < 0) pending = false;  // every zero of the row is taken: no star, like the reference
      if (!__ballot_sync(FULL, pending)) break;
      const unsigned peers = __match_any_sync(FULL, pending ? cand : (-2 - lane));
      const unsigned clash = __ballot_sync(FULL, pending && (peers & lt) != 0u);
      const int first_clash = clash ? (__ffs(clash) - 1) : 32;
      const bool commit = pending && lane < first_clash;
      if (commit) {
        S.row_star[r] = (int8_t)cand;
        S.col_star[cand] = (int8_t)r;
        pending = false;
      }
      stars += __popc(__ballot_sync(FULL, commit));
#pragma unroll 1
      for (int k = 0; k < mw; k++) {
        const uint32_t add = __reduce_or_sync(FULL, (commit && (cand >> 5) == k) ? (1u << (cand & 31)) : 0u);
        if (lane == k) starcols |= add;
      }
    }
  }
  __syncwarp();
  colcov = starcols;

#pragma unroll 1
  for (;;) {
    if (stars >= n) break;
    // rows that own an uncovered zero, from scratch (after step 3 or after a cost shift)
    rowhas = 0u;
#pragma unroll 1
    for (int grp = 0; grp < nwr; grp++) {
      const int r = grp * 32 + lane;
      bool any = false;
#pragma unroll 1
      for (int k = 0; k < mw; k++) {
        const uint32_t cc = __shfl_sync(FULL, colcov, k);
        if (r < n) any = any || ((S.Z[r * kWarpZS + k] & ~cc) != 0u);
      }
      const uint32_t rc = __shfl_sync(FULL, rowcov, grp);
      const uint32_t b = __ballot_sync(FULL, any) & ~rc;
      if (lane == grp) rowhas = b;
    }
    bool augmented = false;
#pragma unroll 1
    for (;;) {
      // step 4: first uncovered zero in row-major order
      if (--budget < 0) { act = 9; break; }
      if (TIMERS && lane == 0) ph[12]++;
      const unsigned hb = __ballot_sync(FULL, rowhas != 0u);
      if (!hb) break;  // none left: step 6
      const int hsrc = __ffs(hb) - 1;
      const uint32_t hv = __shfl_sync(FULL, rowhas, hsrc);
      const int fr = hsrc * 32 + __ffs(hv) - 1;
      const int sc = S.row_star[fr];
      const uint32_t zv = (lane < mw) ? (S.Z[fr * kWarpZS + lane] & ~colcov) : 0u;
      const unsigned zb = __ballot_sync(FULL, zv != 0u);
      const int zsrc = __ffs(zb) - 1;
      const uint32_t zvv = __shfl_sync(FULL, zv, zsrc);
      const int fc = zsrc * 32 + __ffs(zvv) - 1;
      if (sc < 0) {
        // step 5: flip stars along the alternating path that starts at the primed zero (fr, fc)
        int endc = -1;
        __syncwarp();
        if (lane == 0) {
          int r = fr, c = fc;
#pragma unroll 1
          for (int hops = 0;; hops++) {
            const int rs = S.col_star[c];
            S.row_star[r] = (int8_t)c;
            S.col_star[c] = (int8_t)r;
            if (rs < 0) { endc = c; break; }
            r = rs;
            c = S.row_prime[r];
            if (c < 0 || hops > n + m) { endc = -2; break; }  // cannot happen in a valid state
          }
        }
        endc = __shfl_sync(FULL, endc, 0);
        if (endc < 0) { act = 9; break; }
        if (lane == (endc >> 5)) starcols |= 1u << (endc & 31);
        stars++;
        augmented = true;
        __syncwarp();
        break;
      }
      // the row has a star: prime the zero, cover the row, uncover the star's column
      if (lane == 0) S.row_prime[fr] = (int8_t)fc;
      if (lane == (fr >> 5)) {
        rowcov |= 1u << (fr & 31);
        rowhas &= ~(1u << (fr & 31));
      }
      if (lane == (sc >> 5)) colcov &= ~(1u << (sc & 31));
      // uncovered rows with a zero in the newly uncovered column now own an uncovered zero
      const int kw = sc >> 5;
      const uint32_t bit = 1u << (sc & 31);
#pragma unroll 1
      for (int grp = 0; grp < nwr; grp++) {
        const int r = grp * 32 + lane;
        const bool has = (r < n) && ((S.Z[r * kWarpZS + kw] & bit) != 0u);
        const uint32_t b = __ballot_sync(FULL, has);
        const uint32_t rc = __shfl_sync(FULL, rowcov, grp);
        if (lane == grp) rowhas |= b & ~rc;
      }
    }
    if (act != 0) break;
    if (augmented) {  // step 3: cover the starred columns, uncover all rows (stale primes are never read)
      colcov = starcols;
      rowcov = 0u;
      continue;
    }
    if (TIMERS) { if (lane == 0) { const long long now = clock64(); ph[5] += now - ph[15]; ph[15] = now; ph[11]++; } }
    // step 6, with the helper if there is one
    if (lane < 4) { S.cov[0][lane] = rowcov; S.cov[1][lane] = colcov; }
    if (lane == 0) S.cmd = 1;
    team.sync();
    cost_shift(S, tm, n, m, team);
    if (TIMERS) { if (lane == 0) { const long long now = clock64(); ph[6] += now - ph[15]; ph[15] = now; } }
  }
  if (team.nh == 2) {
    if (lane == 0) S.cmd = act;
    team.sync();
  }
  return act;
}

// the helper's side of warp_munkres: cost shifts until the leader reports the outcome
template <bool SPILL>
__device__ __forceinline__ int solver_helper(WarpShared &S, const Cells<SPILL> tm, const int n, const int m, const Team &team) {
#pragma unroll 1
  for (;;) {
    team.sync();
    const int cmd = S.cmd;
    if (cmd != 1) return cmd;
    cost_shift(S, tm, n, m, team);
  }
}

// aux area of the workspace (W2T_SORT_AUX_BYTES): 16 ints of header, then three arrays of n_substreams ints
struct WarpQueues {
  int32_t *hdr;         // [0] big items taken  [1] small items taken  [2] big items  [3] small items
  int32_t *cls;         // per sub-stream: kCls*
  int32_t *big, *small; // sub-streams for the warp kernel in launch order: those that need a big share of
                        // tensor memory, and the rest
};
__host__ __device__ inline WarpQueues warp_queues(char *aux, int nq) {
  WarpQueues Q;
  Q.hdr = reinterpret_cast<int32_t *>(aux);
  Q.cls = Q.hdr + 16;
  Q.big = Q.cls + nq;
  Q.small = Q.big + nq;
  return Q;
}

// One CTA per SM, twelve persistent warps: leaders 0-3 and their helpers 4-7 share kBigCols columns of their
// tensor-memory quarter and serve the queue of crowded sub-streams (then help with the other one); warps 8-11
// own kSmallCols columns and serve the queue of light sub-streams.
template <bool TIMERS>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 1) sort_warp_kernel(const SortParams P) {
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(128) unsigned char w2t_warp_smem[];
  __shared__ uint32_t s_tmem_base;
  const int lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const int wq = warp & 3;                  // quarter of tensor memory this warp can reach
  const bool bigw = warp < 8;
  Team team;
  team.half = (warp >> 2) == 1 ? 1 : 0;
  team.nh = bigw ? 2 : 1;
  team.bar = 1 + wq;
  const bool lead = team.half == 0;
  WarpShared &S = reinterpret_cast<WarpShared *>(w2t_warp_smem)[bigw ? wq : 4 + wq];
  const unsigned lt = (1u << lane) - 1u;
  // all of the SM's tensor memory: one CTA per SM (227 KB of shared memory see to that)
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
        (uint32_t)__cvta_generic_to_shared(&s_tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  // this team's share: the 32 lanes of its quarter; pairs take the low columns, single warps the rest
  const int kCols = bigw ? kBigCols : kSmallCols;
  const uint32_t tm_base = s_tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(bigw ? 0 : kBigCols);
  float *const spill = P.spill + ((size_t)blockIdx.x * kTeamsPerCta + (size_t)(bigw ? wq : 4 + wq)) * kSpillFloats;
  const int NC = P.p.n_classes;
  const float4 *det_box = reinterpret_cast<const float4 *>(P.p.det_box);
  const int max_age = P.p.max_age, min_hits = P.p.min_hits;
  const bool nep50 = P.nep50 != 0;
  const WarpQueues Q = warp_queues(reinterpret_cast<char *>(P.queue), P.p.n_streams * NC);
  const int n_big = Q.hdr[2], n_small = Q.hdr[3];

  // The first round of sub-streams is dealt out like cards — the k-th heaviest goes to SM k mod gridDim — so
  // that every SM starts with the same mix of long and short chains; after that teams pull from the queues.
  bool first = true;
#pragma unroll 1
  for (;;) {
    int q = -1;
    if (lead) {
      const int dealt = wq * (int)gridDim.x + (int)blockIdx.x, skip = 4 * (int)gridDim.x;
      int i = dealt;
      if (bigw) {
        if (!first) {
          if (lane == 0) i = atomicAdd(&Q.hdr[0], 1) + skip;
          i = __shfl_sync(FULL, i, 0);
        }
        if (i < n_big) q = Q.big[i];
        else {  // no crowded sub-stream left: take a light one
          if (lane == 0) i = atomicAdd(&Q.hdr[1], 1) + skip;
          i = __shfl_sync(FULL, i, 0);
          if (i < n_small) q = Q.small[i];
        }
      } else {
        if (!first) {
          if (lane == 0) i = atomicAdd(&Q.hdr[1], 1) + skip;
          i = __shfl_sync(FULL, i, 0);
        }
        if (i < n_small) q = Q.small[i];
      }
      first = false;
    }
    if (team.nh == 2) {
      if (lead && lane == 0) S.cur_q = q;
      team.sync();
      q = S.cur_q;
      team.sync();  // the helper has read it before the leader can post the next one
    }
    if (q < 0) break;
    const int s = q / NC, c = q - s * NC;
    const int Tcap = P.track_cap[q];
    char *slab = P.ws + P.ws_offset[q];
    const SlabLayout L = slab_layout(Tcap, P.det_cap[q]);
    double *st = reinterpret_cast<double *>(slab + L.st);   // [20][Tcap]: x[7], block-form P[13], by list position
    int *tsuA = reinterpret_cast<int *>(slab + L.tsu);
    int *hsA = reinterpret_cast<int *>(slab + L.hs);
    int *bgA = reinterpret_cast<int *>(slab + L.bg);
    int *bkA = reinterpret_cast<int *>(slab + L.bk);
    const int img0 = P.p.stream_img_offsets[s], img1 = P.p.stream_img_offsets[s + 1];
    const double camW = P.p.cam_wh[2 * s], camH = P.p.cam_wh[2 * s + 1];
    const double thr = P.p.iou_thr[c];
    const float thr_f = (float)thr;

    int T = 0, frame_count = 0, err = 0;
    bool started = false, bail = false;
    if (lead && lane == 0) P.r.first_img[q] = -1;
    long long ph[TIMERS ? 16 : 1];
    if (TIMERS) {
#pragma unroll
      for (int i = 0; i < (TIMERS ? 16 : 1); i++) ph[i] = 0;
      ph[TIMERS ? 15 : 0] = clock64();
      unsigned long long ns;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
      ph[TIMERS ? 14 : 0] = -(long long)ns;  // wall nanoseconds of the sub-stream (completed below)
    }
#define W2T_WTICK(i) do { if (TIMERS && lane == 0) { const long long now_ = clock64(); ph[i] += now_ - ph[TIMERS ? 15 : 0]; ph[TIMERS ? 15 : 0] = now_; } } while (0)

    // (exists, count, start) of image i+1 are loaded at the top of iteration i, and its detections are pulled
    // into L2 while image i is tracked, so no DRAM round trip sits on the serial path
    int cur_exists = 0, cur_cnt = 0, cur_start = 0;
    if (img0 < img1) {
      cur_exists = (P.p.img_exists == nullptr) ? 1 : (int)P.p.img_exists[img0];
      cur_cnt = P.p.det_count[img0 * NC + c];
      cur_start = P.p.det_start[img0 * NC + c];
    }
#pragma unroll 1
    for (int img = img0; img < img1; ++img) {
      const int g = img * NC + c;
      const int this_exists = cur_exists, D_in = cur_cnt, base = cur_start;
      if (img + 1 < img1) {
        cur_exists = (P.p.img_exists == nullptr) ? 1 : (int)P.p.img_exists[img + 1];
        cur_cnt = P.p.det_count[g + NC];
        cur_start = P.p.det_start[g + NC];
      }
      bool skip = (this_exists == 0);
      const int D = skip ? 0 : D_in;
      if (!skip && !started) {
        if (D == 0) skip = true;  // no Sort object for this category yet (tracker_sort.py:32-33)
        else {
          started = true;
          if (lead && lane == 0) P.r.first_img[q] = img - img0;
        }
      }
      if (skip || err) {
        if (lead && lane == 0) { P.r.out_count[g] = 0; P.r.created[g] = 0; }
        continue;
      }
      if (D > kWarpDim || T > kWarpDim) { bail = true; break; }
      frame_count++;
#pragma unroll 1
      for (int d = team.half * 32 + lane; d < D; d += 32 * team.nh) {
        S.det[d] = __ldg(det_box + base + d);
        S.dstat[d] = 0;
      }
#pragma unroll 1
      for (int t = team.half * 32 + lane; t < T; t += 32 * team.nh) S.match[t] = -1;
      team.sync();
      if (lead && img + 1 < img1 && cur_exists && lane * 8 < cur_cnt)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(det_box + cur_start + lane * 8));
      W2T_WTICK(1);

      // ---- A. association (sort.py:193-230) ----------------------------------------------------------
      if (T > 0 && D > 0) {
        const bool flipped = D > T;  // the solver transposes when there are more rows than columns
        const int n = flipped ? T : D, m = flipped ? D : T;
        const int mw = (m + 31) >> 5, NP = (n + 7) & ~7;
        auto mask_of = [&](const bool is_det, const int i) -> uint32_t {
          if (is_det) {
            const float4 b = S.det[i];
            return strip_mask_checked((double)b.x, (double)b.y, (double)b.z, (double)b.w);
          }
          return strip_mask_checked(S.box[0][i], S.box[1][i], S.box[2][i], S.box[3][i]);
        };
        if (lead) {
          // 1. every column enters the bit sets of the strips its box touches: a 32 x 32 bit transpose per
          //    column word (lane b collects bit b of every lane's mask)
          *reinterpret_cast<uint4 *>(S.strip[lane]) = make_uint4(0u, 0u, 0u, 0u);
          __syncwarp();
#pragma unroll 1
          for (int k = 0; k < mw; k++) {
            const int cc = k * 32 + lane;
            const uint32_t cm = (cc < m) ? mask_of(flipped, cc) : 0u;
            uint32_t mine = 0u;
#pragma unroll 4
            for (int b = 0; b < 32; b++) {
              const uint32_t w = __ballot_sync(FULL, (cm >> b) & 1u);
              if (lane == b) mine = w;
            }
            S.strip[lane][k] = mine;
          }
#pragma unroll 1
          for (int cc = lane; cc < m; cc += 32) S.col_star[cc] = -1;
          __syncwarp();
        }
        // 2. + 3., 32 rows at a time.
        //    The leader, one lane per row: candidate columns = (union of the sets of the row's x strips) AND
        //    (union over its y strips); exact cost of each candidate (parked in S.raw); row minimum (step 1 of
        //    the solver).  Every other pair is strictly disjoint and costs -0.0f.
        //    Then the team, lanes own columns: reduced costs -> tensor memory, zero bit words -> S.Z, eight rows
        //    at a time.
        // the matrix goes to the team's share of tensor memory, or — the rare image that outgrows it — to the
        // team's spill area (a second, cold instance of the same code)
        auto build_and_solve = [&](const auto tm) -> int {
#pragma unroll 1
          for (int rb = 0; rb < n; rb += 32) {
            const int r = rb + lane;
            if (lead && r < n) {
              const uint32_t rm = mask_of(!flipped, r);
              uint4 cx = make_uint4(0u, 0u, 0u, 0u), cy = cx;
#pragma unroll 1
              for (uint32_t w = rm & 0xffffu; w; w &= w - 1u) {
                const uint4 v = *reinterpret_cast<const uint4 *>(S.strip[__ffs(w) - 1]);
                cx.x |= v.x; cx.y |= v.y; cx.z |= v.z; cx.w |= v.w;
              }
#pragma unroll 1
              for (uint32_t w = rm >> 16; w; w &= w - 1u) {
                const uint4 v = *reinterpret_cast<const uint4 *>(S.strip[16 + __ffs(w) - 1]);
                cy.x |= v.x; cy.y |= v.y; cy.z |= v.z; cy.w |= v.w;
              }
              uint32_t *zr = S.Z + r * kWarpZS;
              zr[0] = cx.x & cy.x; zr[1] = cx.y & cy.y; zr[2] = cx.z & cy.z; zr[3] = cx.w & cy.w;
              const int ncand = __popc(zr[0]) + __popc(zr[1]) + __popc(zr[2]) + __popc(zr[3]);
              // minimum on the order-preserving integer image of the float (NaN sorts last, like fminf ignores it)
              uint32_t mn_u = (ncand < m) ? Munkres<32>::ordered(-0.0f) : 0xffffffffu;
              float *row = S.raw + lane * m;
#pragma unroll 1
              for (int k = 0; k < mw; k++) {
#pragma unroll 1
                for (uint32_t w = zr[k]; w; w &= w - 1u) {
                  const int cc = k * 32 + __ffs(w) - 1;
                  const int di = flipped ? cc : r, ti = flipped ? r : cc;
                  const float v = -iou_pair_call(S.det[di], S.box[0][ti], S.box[1][ti], S.box[2][ti], S.box[3][ti]);
                  row[cc] = v;
                  mn_u = min(mn_u, Munkres<32>::ordered(v));
                }
              }
              S.rowmin[r] = Munkres<32>::unordered(mn_u);
              S.row_star[r] = -1;
              S.row_prime[r] = -1;
            }
            team.sync();
            const int r_end = min(rb + 32, NP);
#pragma unroll 1
            for (int r0 = rb + team.half * 8; r0 < r_end; r0 += 8 * team.nh) {
              float mn[8];
#pragma unroll
              for (int j = 0; j < 8; j++) mn[j] = S.rowmin[r0 + j];
              const float *p = S.raw + (r0 - rb) * m + lane;
              uint32_t *zp = S.Z + r0 * kWarpZS;
#pragma unroll 1
              for (int k = 0; k < mw; k++) {
                const bool colv = k * 32 + lane < m;
                const float *pk = p + k * 32;
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                  const uint32_t cw = zp[j * kWarpZS + k];
                  v[j] = -0.0f;
                  if ((cw >> lane) & 1u) v[j] = pk[j * m];
                }
#pragma unroll
                for (int j = 0; j < 8; j++) {
                  v[j] = v[j] - mn[j];
                  const uint32_t word = __ballot_sync(FULL, colv && v[j] == 0.0f);
                  if (lane == 0) zp[j * kWarpZS + k] = word;
                }
                tm.st8(k * NP + r0, v);
              }
            }
            tm.wait_st();
            team.sync();
          }
          W2T_WTICK(2);
          if (TIMERS && lane == 0) ph[13]++;
            return lead ? warp_munkres<TIMERS>(S, tm, n, m, team, ph) : solver_helper(S, tm, n, m, team);
        };
        const int act = (mw * NP > kCols) ? build_and_solve(Cells<true>{tm_base, spill})
                                          : build_and_solve(Cells<false>{tm_base, spill});
        if (act != 0) err = W2T_ERR_ARG;
        W2T_WTICK(5);
        if (lead) {
          __syncwarp();
          // matched / rejected detections (sort.py:217-222)
#pragma unroll 1
          for (int d = lane; d < D; d += 32) {
            const int t = flipped ? S.col_star[d] : S.row_star[d];
            if (t >= 0) {
              const float o = iou_pair_call(S.det[d], S.box[0][t], S.box[1][t], S.box[2][t], S.box[3][t]);
              // NumPy 1.x compares the float32 entry with the python float in float64, NEP 50 in float32
              const bool rejected = nep50 ? (o < thr_f) : ((double)o < thr);
              if (rejected) S.dstat[d] = 2;  // becomes a new tracker AFTER the unassigned ones
              else { S.dstat[d] = 1; S.match[t] = (int8_t)d; }
            }
          }
          __syncwarp();
        }
      }
      W2T_WTICK(7);
      // new trackers: unassigned detections first, then the rejected ones (sort.py:208-222, :276-278)
      int n_new = 0;
      if (lead) {
#pragma unroll 1
        for (int pass = 0; pass < 2; pass++) {
#pragma unroll 1
          for (int d0 = 0; d0 < D; d0 += 32) {
            const int d = d0 + lane;
            const bool a = d < D && S.dstat[d] == (pass ? 2 : 0);
            const unsigned b = __ballot_sync(FULL, a);
            if (a) S.newdet[n_new + __popc(b & lt)] = (int8_t)d;
            n_new += __popc(b);
          }
        }
        if (team.nh == 2 && lane == 0) S.n_new = n_new;
      }
      team.sync();
      if (team.nh == 2) n_new = S.n_new;
      const int Ttot = T + n_new;
      if (Ttot > Tcap) err = W2T_ERR_CAPACITY;
      if (err) {
        if (lead && lane == 0) { P.r.out_count[g] = 0; P.r.created[g] = 0; }
        continue;
      }
      W2T_WTICK(8);

      // ---- B. one pass over the tracker list: update or create, emit, age test, predict of the NEXT image;
      // survivors are written back at their new list position.  32 trackers per warp and trip. -----------
      int n_live = 0, emitted = 0;
#pragma unroll 1
      for (int tp = 0, par = 0; tp < Ttot; tp += 32 * team.nh, par ^= 1) {
        const int t = tp + team.half * 32 + lane;
        bool ok = false, surv = false;
        double x[7], Pm[kBlockP], ob0 = 0., ob1 = 0., ob2 = 0., ob3 = 0., oconf = 0., nb[4];
        int tsu = 0, hs = 0, obg = 0, obk = 0;
        if (t < Ttot) {
          int md;
          if (t >= T) {  // sort.py:276-278
            md = S.newdet[t - T];
            obg = g;
            obk = t - T;
          } else {
#pragma unroll
            for (int k = 0; k < 7; k++) x[k] = st[k * Tcap + t];
#pragma unroll
            for (int k = 0; k < kBlockP; k++) Pm[k] = st[(7 + k) * Tcap + t];
            tsu = tsuA[t];
            hs = hsA[t];
            obg = bgA[t];
            obk = bkA[t];
            md = S.match[t];
          }
          if (md >= 0) {
            const float4 d4 = S.det[md];
            double z[4];
            bbox_to_z_d(d4.x, d4.y, d4.z, d4.w, nep50, z);
            if (t >= T) kfb_init_z(z, x, Pm);
            else {  // sort.py:270-273, :153-164
              kfb_update_z(x, Pm, z);
              tsu = 0;
              hs += 1;
            }
          }
          // sort.py:281-289 and utils.py:37-49
          if (tsu < 1 && (hs >= min_hits || frame_count <= min_hits)) {
            double w, h;
            box_wh(x[2], x[3], w, h);
            const double e = ((Pm[0] + Pm[4]) + Pm[8]) / 3.0;  // mean(P00, P11, P22), sort.py:190
            const double conf = exp(-e * 0.1);
            const double x1 = clipd(x[0] - w / 2., 0., camW), y1 = clipd(x[1] - h / 2., 0., camH);
            const double x2 = clipd(x[0] + w / 2., 0., camW), y2 = clipd(x[1] + h / 2., 0., camH);
            const double wd = x2 - x1, ht = y2 - y1;
            if (!(wd < 1 || ht < 1)) {
              ok = true;
              ob0 = x1; ob1 = y1; ob2 = wd; ob3 = ht;
              oconf = clipd(conf, 0.2, 1.0);
            }
          }
          surv = !(tsu > max_age);  // sort.py:292
          if (surv) {
            // predict of the next image (sort.py:166-178)
            kfb_predict(x, Pm);
            if (tsu > 0) hs = 0;
            tsu += 1;
            double w, h;
            box_wh(x[2], x[3], w, h);
            nb[0] = x[0] - w / 2.; nb[1] = x[1] - h / 2.; nb[2] = x[0] + w / 2.; nb[3] = x[1] + h / 2.;
            // a NaN box: the tracker is dropped before the next association (sort.py:261-265)
            if (isnan(nb[0]) || isnan(nb[1]) || isnan(nb[2]) || isnan(nb[3])) surv = false;
            else if (isinf(nb[0]) || isinf(nb[1]) || isinf(nb[2]) || isinf(nb[3])) atomicMax(P.status, W2T_ERR_NONFINITE);
          }
        }
        const unsigned bs = __ballot_sync(FULL, surv), be = __ballot_sync(FULL, ok);
        int live0 = n_live, emit0 = emitted;  // where this warp's survivors / rows start
        if (team.nh == 2) {
          if (lane == 0) { S.cnt[par][team.half][0] = __popc(bs); S.cnt[par][team.half][1] = __popc(be); }
          team.sync();  // also: every lane of the team has read its old position before any overwrites one
          const int l0 = S.cnt[par][0][0], e0 = S.cnt[par][0][1], l1 = S.cnt[par][1][0], e1 = S.cnt[par][1][1];
          if (team.half) { live0 += l0; emit0 += e0; }
          n_live += l0 + l1;
          emitted += e0 + e1;
        } else {
          __syncwarp();  // every lane has read its old position before any lane overwrites it
          n_live += __popc(bs);
          emitted += __popc(be);
        }
        if (surv) {
          const int pos = live0 + __popc(bs & lt);
#pragma unroll
          for (int k = 0; k < 7; k++) st[k * Tcap + pos] = x[k];
#pragma unroll
          for (int k = 0; k < kBlockP; k++) st[(7 + k) * Tcap + pos] = Pm[k];
          tsuA[pos] = tsu;
          hsA[pos] = hs;
          bgA[pos] = obg;
          bkA[pos] = obk;
          if (pos < kWarpDim) {
#pragma unroll
            for (int k = 0; k < 4; k++) S.box[k][pos] = nb[k];
          }
        }
        if (ok) {
          const size_t o = (size_t)base + emit0 + __popc(be & lt);
          double *ob = P.r.out_box + 4 * o;
          ob[0] = ob0; ob[1] = ob1; ob[2] = ob2; ob[3] = ob3;
          P.r.out_score[o] = oconf;
          P.r.out_birth[2 * o + 0] = obg;
          P.r.out_birth[2 * o + 1] = obk;
        }
      }
      if (lead && lane == 0) { P.r.out_count[g] = emitted; P.r.created[g] = n_new; }
      T = n_live;
      team.sync();
      W2T_WTICK(9);
    }
    if (TIMERS && P.timers != nullptr && lead && lane == 0) {
      ph[0] = frame_count;
      unsigned long long ns;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
      ph[TIMERS ? 14 : 0] += (long long)ns;
      for (int i = 0; i < (TIMERS ? 16 : 1); i++) P.timers[(size_t)q * 16 + i] = ph[i];
    }
#undef W2T_WTICK
    if (bail) {  // tracked again from its first image by the CTA kernel
      if (lead && lane == 0) Q.cls[q] = kClsBailed;
      continue;
    }
    if (!lead) continue;
    if (err && lane == 0) atomicMax(P.status, err);

    // optional: filter state of every live tracker, already predicted one step past the last image
    if (P.r.final_count != nullptr) {
      if (lane == 0) P.r.final_count[q] = T;
      if (P.r.final_state != nullptr) {
        const int cap = P.r.final_cap;
#pragma unroll 1
        for (int t = lane; t < T && t < cap; t += 32) {
          double *dst = P.r.final_state + ((size_t)q * cap + t) * 56;
          double pb[kBlockP], Pd[49];
          for (int k = 0; k < 7; k++) dst[k] = st[k * Tcap + t];
          for (int k = 0; k < kBlockP; k++) pb[k] = st[(7 + k) * Tcap + t];
          kfb_to_dense(pb, Pd);
          for (int k = 0; k < 49; k++) dst[7 + k] = Pd[k];
        }
      }
    }
    // completion tracking: everything this team wrote is visible before its chunk's counter moves (the
    // helper's writes reached the leader through the team barrier at the end of the last image)
    if (P.chunk_done != nullptr) {
      __syncwarp();
      if (lane == 0) {
        __threadfence();
        atomicAdd(&P.chunk_done[P.chunk_of[q]], 1);
      }
    }
  }
  // every warp of the CTA is done with its share before the tensor memory goes back
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(s_tmem_base));
}

// Actual crowding of every sub-stream (the launch plan may have been computed from upper bounds, e.g. the group
// sizes BEFORE the ensemble): most detections in one image (sort_dmax_kernel, one thread per sub-stream, left in
// Q.cls) -> class flag (kCls*) for the kernels above, and the two queues of the warp kernel, both in the plan's
// launch order (heaviest first; sort_classify_kernel, one block).
__global__ void sort_dmax_kernel(const w2t_sort_problem_t p, int32_t *dmax_out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  const int NC = p.n_classes;
  if (q >= p.n_streams * NC) return;
  const int s = q / NC, c = q - s * NC;
  int dmax = 0, over = 0;
  for (int img = p.stream_img_offsets[s]; img < p.stream_img_offsets[s + 1]; img++)
    if (p.img_exists == nullptr || p.img_exists[img]) {
      const int d = p.det_count[img * NC + c];
      dmax = max(dmax, d);
      over += d > W2T_NARROW_DETS;
    }
  dmax_out[q] = min(dmax, 0xffff) | (min(over, 0x7fff) << 16);  // images whose matrix may outgrow a pair's tensor memory
}

__global__ void __launch_bounds__(1024) sort_classify_kernel(const w2t_sort_problem_t p, const int32_t *order, WarpQueues Q) {
  __shared__ int s_tot[2][32];
  __shared__ int s_base[2];
  const int NC = p.n_classes, nq = p.n_streams * NC;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < 2) s_base[threadIdx.x] = 0;
  __syncthreads();
  for (int i0 = 0; i0 < nq; i0 += 1024) {
    const int i = i0 + threadIdx.x;
    bool big = false, small = false;
    int q = 0;
    if (i < nq) {
      q = order[i];
      const int dmax = Q.cls[q] & 0xffff, over = Q.cls[q] >> 16;  // left there by sort_dmax_kernel
      // warps take a sub-stream whose matrices fit a pair's tensor memory at all but a few images (those spill)
      const bool warps = dmax <= W2T_NARROW_DETS || (dmax <= kSpillDets && over <= kSpillImages);
      const int cls = dmax > kClassifyHuge ? kClsHuge : dmax > W2T_WIDE_DETS ? kClsWide : !warps ? kClsMid : kClsWarp;
      Q.cls[q] = cls;
      // a crowd of D detections meets about 1.3 D trackers (max_age 2): ceil8(D) rows x ceil32(1.3 D + 8) / 32 words
      const int words = (min(kWarpDim, dmax + dmax / 3 + 8) + 31) >> 5;
      big = cls == kClsWarp && ((dmax + 7) & ~7) * words > kSmallCols;
      small = cls == kClsWarp && !big;
    }
    const unsigned bb = __ballot_sync(0xffffffffu, big), bs = __ballot_sync(0xffffffffu, small);
    if (lane == 0) { s_tot[0][warp] = __popc(bb); s_tot[1][warp] = __popc(bs); }
    __syncthreads();
    int pb = s_base[0], ps = s_base[1], tb = 0, ts = 0;
    for (int w = 0; w < 32; w++) {
      if (w < warp) { pb += s_tot[0][w]; ps += s_tot[1][w]; }
      tb += s_tot[0][w];
      ts += s_tot[1][w];
    }
    const unsigned lt = (1u << lane) - 1u;
    if (big) Q.big[pb + __popc(bb & lt)] = q;
    if (small) Q.small[ps + __popc(bs & lt)] = q;
    __syncthreads();
    if (threadIdx.x == 0) { s_base[0] += tb; s_base[1] += ts; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { Q.hdr[2] = s_base[0]; Q.hdr[3] = s_base[1]; }
}

}  // namespace w2t

// sort_crowd.cuh — the SORT tracker for CROWDED sub-streams: one thread-block CLUSTER per (stream, category).
//
// Same path as sort_warp.cuh (tracking/sort/sort.py:193-296 with the scikit-learn 0.22.2 Munkres solver behind
// sort.py:206), for images with up to kCrowdN x kCrowdM assignment problems (BASELINE.json config C4: ~450
// detections meet ~520 live trackers, a 0.94 MB float32 cost matrix and 30-110 cost shifts per image).  Such a
// matrix fits neither shared memory nor tensor memory of one SM, and walking it in L2 through one CTA costs
// ~19 us per cost shift (sort_kernel.cuh, r01: 1.35 ms per image).  Here a CLUSTER of 8 or 16 CTAs (16 is the
// non-portable maximum; one CTA per SM) shares one sub-stream:
//   * the cost matrix is DISTRIBUTED over the shared memory of the cluster's worker CTAs (all but CTA 0): blocks of
//     eight rows are dealt out to their warps (56 or 120), and a warp keeps its blocks in its own CTA's shared memory for the whole image — construction
//     (dense IoU, lanes own columns), row reduction and every cost shift touch local shared memory only;
//   * what crosses CTAs is small and goes through distributed shared memory (st / red .shared::cluster): the
//     leader pushes the cover masks (36 words) and a command word to every CTA, every warp pushes its partial
//     minimum into every CTA's accumulator (red.min), and the zero-bit words of the updated rows go to the
//     leader's bit matrix; three cluster barriers per cost shift;
//   * the serial steps of the solver (2-5) run in one warp of CTA 0 with the lane-distributed masks of
//     sort_warp.cuh (28 column words, 20 row words); association read-out, the Kalman pass and the list
//     compaction run on CTA 0 (eight warps) — they are a few percent of a crowded image.
// Tracker state lives in the sub-stream's slab in list order, predicted boxes included (rows 20-23), so the
// other CTAs read them from global memory after the cluster barrier that ends an image.
// A sub-stream that outgrows even this (n > kCrowdN, m > kCrowdM, or a matrix beyond the cluster's shared
// memory) is flagged for the next bigger path: a 16-CTA cluster, then the global-memory CTA kernel (sort_kernel.cuh).
#pragma once

#include <cstdio>

#include "sort_warp.cuh"

namespace w2t {

constexpr int kCrowdWarps = 8;       // warps per CTA
constexpr int kCrowdN = 640;         // most rows (= min(D, T))
constexpr int kCrowdM = 896;         // most columns (= max(D, T)), also most detections / live trackers
constexpr int kCrowdMW = kCrowdM / 32;
constexpr int kCrowdZS = kCrowdN + 1;    // zero bit matrix, WORD-major: word k of row r at k * kCrowdZS + r (rows of one
                                         // word are contiguous, so a block's eight words go out as one 32-byte store)
constexpr int kCrowdNW = kCrowdN / 32;   // row words
constexpr int kCrowdMaxCtas = 16;

// at the same offset in every CTA of the cluster (targets of remote stores)
struct __align__(128) CrowdCommon {
  uint32_t cov[2][kCrowdMW];           // pushed by the leader: row cover / column cover words
  uint32_t minacc[2];                  // pushed by every worker CTA (red.min), by round parity
  uint32_t minloc;                     // this CTA's own minimum of the round
  int32_t cmd;                         // pushed by the leader: see kCmd*
  int32_t hdr_T;                       // pushed by the leader: live trackers of the image to associate
};
// CTA 0 (the leader) only; the other CTAs (workers) use this space and everything behind it for the matrix
struct __align__(128) CrowdLeader {
  float4 det[kCrowdM];                 // this image's detections
  double box[4][kCrowdM];              // predicted boxes of the live trackers, by list position
  uint32_t Z[kCrowdMW * kCrowdZS];     // zero bit matrix: word k of row r at k * kCrowdZS + r
  uint8_t ZT[(kCrowdN / 8) * kCrowdM]; // the same bits by column: byte b of column c (rows 8b..8b+7) at b * kCrowdM + c
  uint32_t rowflag[kCrowdNW];          // after a cost shift: rows that gained a zero in an uncovered column (pushed as bytes)
  int16_t row_star[kCrowdN], row_prime[kCrowdN], col_star[kCrowdM];
  int16_t match[kCrowdM];              // per tracker: matched detection or -1
  int16_t newdet[kCrowdM];             // detections that become trackers, in the reference's order
  int8_t dstat[kCrowdM];               // per detection: 0 unassigned, 1 matched, 2 assigned but rejected
  int32_t cnt[2][kCrowdWarps][2];      // Kalman pass: [trip parity][warp][survivors, rows emitted]
  int32_t bcast[4];
};
enum { kCmdExit = 0, kCmdAssociate = 1, kCmdSkip = 2, kCmdShift = 3, kCmdSolved = 4, kCmdGaveUp = 5 };

constexpr int kCrowdSmemBytes = 232448 - 1024;
static_assert(sizeof(CrowdCommon) + sizeof(CrowdLeader) <= kCrowdSmemBytes, "leader CTA out of shared memory");
// floats of cost matrix a worker warp can hold (whole cell-word blocks of 8 rows x 32 lanes)
constexpr int kCrowdMatrixFloats = ((kCrowdSmemBytes - (int)sizeof(CrowdCommon)) / 4 / kCrowdWarps / 256) * 256;

__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
// address of `p` (a shared-memory object of this CTA) in the shared memory of CTA `cta` of the cluster
__device__ __forceinline__ uint32_t dsmem_addr(const void *p, const uint32_t cta) {
  uint32_t la = (uint32_t)__cvta_generic_to_shared(p), ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(cta));
  return ra;
}
__device__ __forceinline__ void dsmem_st(const uint32_t ra, const uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(ra), "r"(v) : "memory");
}
__device__ __forceinline__ void dsmem_st8(const uint32_t ra, const uint32_t v) {
  asm volatile("st.shared::cluster.u8 [%0], %1;" ::"r"(ra), "r"(v) : "memory");
}
__device__ __forceinline__ void dsmem_red_min(const uint32_t ra, const uint32_t v) {
  asm volatile("red.shared::cluster.min.u32 [%0], %1;" ::"r"(ra), "r"(v) : "memory");
}

// The cluster's view of one association problem: which worker warp owns which block of eight rows, and where
// it keeps it.  Workers are the warps of CTAs 1..nctas-1, numbered CTA-fastest so that a small problem spreads
// over all SMs; block b belongs to worker b % nworkers.
struct Crowd {
  int n, m, mw, nblk;       // rows, columns, column words, row blocks
  int gw, nworkers;         // this warp's worker number (-1: a warp of the leader CTA), number of workers
  int nctas;
  float *mine;              // this warp's share of its CTA's shared memory
  // cell word (block b, word k): eight rows x 32 lanes
  __device__ __forceinline__ float *cell(int b, int k) const { return mine + ((size_t)((b / nworkers) * mw + k) * 8) * 32; }
};

// ---- step 6 by the whole cluster (see the header).  `par`: parity of the round (which accumulator to use).
// `Zl`: the leader's zero bit matrix (CTA 0), written remotely.
struct CrowdRemote { uint32_t z, zt, rowflag; };  // the leader's Z, ZT and rowflag arrays, as cluster addresses

__device__ __forceinline__ void crowd_shift(CrowdCommon &S, const CrowdRemote R, const Crowd &cw, const int par) {
  const uint32_t zbase = R.z;
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = lane_id();
  const int n = cw.n, m = cw.m, mw = cw.mw;
  // costs are >= +0 here (or NaN), so their bit patterns order like the values and NaN sorts above +inf
  uint32_t mn_u = 0x7f800000u;
  if (cw.gw >= 0) {
#pragma unroll 1
    for (int b = cw.gw; b < cw.nblk; b += cw.nworkers) {
      const int r0 = b * 8;
      uint32_t ur = ~(S.cov[0][r0 >> 5] >> (r0 & 31));  // bit j: row r0 + j is uncovered
      if (n - r0 < 8) ur &= (1u << (n - r0)) - 1u;
      if ((ur & 0xffu) == 0u) continue;
#pragma unroll 1
      for (int k = 0; k < mw; k++) {
        const uint32_t ccw = S.cov[1][k];
        const uint32_t valid = (m - k * 32 >= 32) ? 0xffffffffu : ((1u << (m - k * 32)) - 1u);
        if (!(((~ccw & valid) >> lane) & 1u)) continue;  // not an uncovered column of this lane
        const float *p = cw.cell(b, k) + lane;
#pragma unroll
        for (int j = 0; j < 8; j++)
          if ((ur >> j) & 1u) mn_u = min(mn_u, __float_as_uint(p[j * 32]));
      }
    }
    mn_u = __reduce_min_sync(FULL, mn_u);
    if (lane == 0) atomicMin(&S.minloc, mn_u);
    __syncthreads();  // worker CTAs only: all eight warps are here
    if (threadIdx.x < cw.nctas) dsmem_red_min(dsmem_addr(&S.minacc[par], threadIdx.x), S.minloc);
  }
  cluster_sync();
  mn_u = S.minacc[par];
  if (threadIdx.x == 0) { S.minacc[par ^ 1] = 0x7f800000u; S.minloc = 0x7f800000u; }  // ready for the next round
  if (mn_u != 0x7f800000u && cw.gw >= 0) {  // nothing uncovered: the reference leaves the matrix alone
    const float mn = __uint_as_float(mn_u);
#pragma unroll 1
    for (int b = cw.gw; b < cw.nblk; b += cw.nworkers) {
      const int r0 = b * 8;
      const uint32_t cr = (S.cov[0][r0 >> 5] >> (r0 & 31)) & 0xffu;  // bit j: row r0 + j is covered
      uint32_t rf = 0u;  // bit j: row r0 + j gained a zero in an uncovered column
#pragma unroll 1
      for (int k = 0; k < mw; k++) {
        const uint32_t ccw = S.cov[1][k];
        const uint32_t valid = (m - k * 32 >= 32) ? 0xffffffffu : ((1u << (m - k * 32)) - 1u);
        const uint32_t ucm = ~ccw & valid;
        if (cr == 0u && ucm == 0u) continue;  // uncovered rows x covered columns: nothing changes
        const bool colv = (valid >> lane) & 1u;
        const bool u = (ucm >> lane) & 1u;
        float *p = cw.cell(b, k) + lane;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = p[j * 32];
        uint32_t zbyte = 0u, myword = 0u;
#pragma unroll
        for (int j = 0; j < 8; j++) {
          if ((cr >> j) & 1u) v[j] = v[j] + mn;
          if (u) v[j] = v[j] - mn;
          p[j * 32] = v[j];
          const bool z = colv && r0 + j < n && v[j] == 0.0f;
          zbyte |= z ? (1u << j) : 0u;
          const uint32_t word = __ballot_sync(FULL, z);
          rf |= (word & ucm) ? (1u << j) : 0u;
          if (lane == j) myword = word;
        }
        if (lane < 8) dsmem_st(zbase + 4u * (uint32_t)(k * kCrowdZS + r0 + lane), myword);  // one 32-byte store
        dsmem_st8(R.zt + (uint32_t)(b * kCrowdM + k * 32 + lane), zbyte);                    // one 32-byte store
      }
      if (lane == 0) dsmem_st8(R.rowflag + (uint32_t)b, rf & ~cr);
    }
  }
  cluster_sync();
}

// ---- scikit-learn 0.22.2 linear_assignment, the serial steps (2-5), in warp 0 of CTA 0: warp_munkres of
// sort_warp.cuh with 16-bit index arrays, up to 28 column words, and the cluster's cost shift.  Every other
// warp of the cluster follows in crowd_helper().
__device__ __forceinline__ int crowd_munkres(CrowdCommon &S, CrowdLeader &Ld, const CrowdRemote R, const Crowd &cw, int &par) {
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = lane_id();
  const unsigned lt = (1u << lane) - 1u;
  const int n = cw.n, m = cw.m, mw = cw.mw, nwr = (n + 31) >> 5;
  uint32_t starcols = 0u, colcov = 0u, rowcov = 0u, rowhas = 0u;
  int stars = 0, act = 0;
  int budget = 4 * n * n + 64 * (n + m) + 1024;
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
          const uint32_t v = Ld.Z[k * kCrowdZS + r] & ~sc;
          if (v) cand = k * 32 + __ffs(v) - 1;
        }
      }
      if (cand < 0) pending = false;
      if (!__ballot_sync(FULL, pending)) break;
      const unsigned peers = __match_any_sync(FULL, pending ? cand : (-2 - lane));
      const unsigned clash = __ballot_sync(FULL, pending && (peers & lt) != 0u);
      const int first_clash = clash ? (__ffs(clash) - 1) : 32;
      const bool commit = pending && lane < first_clash;
      if (commit) {
        Ld.row_star[r] = (int16_t)cand;
        Ld.col_star[cand] = (int16_t)r;
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
  bool shifted = false;
  // rows of this lane's word that exist (bytes of ZT / rowflag beyond the last row block are stale)
  const uint32_t rv = (n - 32 * lane >= 32) ? 0xffffffffu : (n - 32 * lane > 0 ? ((1u << (n - 32 * lane)) - 1u) : 0u);
#pragma unroll 1
  for (;;) {
    if (stars >= n) break;
    // rows that own an uncovered zero: after a cost shift the workers have flagged them (new zeros appear only
    // in the cells the shift touched); after step 3 (covers reset) from scratch
    rowhas = 0u;
    if (shifted) {
      if (lane < nwr) rowhas = Ld.rowflag[lane] & ~rowcov & rv;
    } else {
#pragma unroll 1
      for (int grp = 0; grp < nwr; grp++) {
        const int r = grp * 32 + lane;
        bool any = false;
#pragma unroll 1
        for (int k = 0; k < mw; k++) {
          const uint32_t cc = __shfl_sync(FULL, colcov, k);
          if (r < n) any = any || ((Ld.Z[k * kCrowdZS + r] & ~cc) != 0u);
        }
        const uint32_t rc = __shfl_sync(FULL, rowcov, grp);
        const uint32_t b = __ballot_sync(FULL, any) & ~rc;
        if (lane == grp) rowhas = b;
      }
    }
    shifted = false;
    bool augmented = false;
#pragma unroll 1
    for (;;) {
      if (--budget < 0) { act = 9; break; }
      const unsigned hb = __ballot_sync(FULL, rowhas != 0u);
      if (!hb) break;
      const int hsrc = __ffs(hb) - 1;
      const uint32_t hv = __shfl_sync(FULL, rowhas, hsrc);
      const int fr = hsrc * 32 + __ffs(hv) - 1;
      const int sc = Ld.row_star[fr];
      const uint32_t zv = (lane < mw) ? (Ld.Z[lane * kCrowdZS + fr] & ~colcov) : 0u;
      const unsigned zb = __ballot_sync(FULL, zv != 0u);
      const int zsrc = __ffs(zb) - 1;
      const uint32_t zvv = __shfl_sync(FULL, zv, zsrc);
      const int fc = zsrc * 32 + __ffs(zvv) - 1;
      if (sc < 0) {
        int endc = -1;
        __syncwarp();
        if (lane == 0) {
          int r = fr, c = fc;
#pragma unroll 1
          for (int hops = 0;; hops++) {
            const int rs = Ld.col_star[c];
            Ld.row_star[r] = (int16_t)c;
            Ld.col_star[c] = (int16_t)r;
            if (rs < 0) { endc = c; break; }
            r = rs;
            c = Ld.row_prime[r];
            if (c < 0 || hops > n + m) { endc = -2; break; }
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
      if (lane == 0) Ld.row_prime[fr] = (int16_t)fc;
      if (lane == (fr >> 5)) {
        rowcov |= 1u << (fr & 31);
        rowhas &= ~(1u << (fr & 31));
      }
      if (lane == (sc >> 5)) colcov &= ~(1u << (sc & 31));
      // uncovered rows with a zero in the newly uncovered column now own an uncovered zero: one word of the
      // column-major bit matrix per lane
      if (lane < nwr) {
        const uint8_t *zt = Ld.ZT + (4 * lane) * kCrowdM + sc;
        const uint32_t w = (uint32_t)zt[0] | ((uint32_t)zt[kCrowdM] << 8) | ((uint32_t)zt[2 * kCrowdM] << 16) | ((uint32_t)zt[3 * kCrowdM] << 24);
        rowhas |= w & ~rowcov & rv;
      }
    }
    if (act != 0) break;
    if (augmented) {
      colcov = starcols;
      rowcov = 0u;
      continue;
    }
    // step 6: push the covers and the command to every CTA, then shift with the whole cluster
#pragma unroll 1
    for (int c = 0; c < cw.nctas; c++) {
      if (lane < kCrowdMW) {
        dsmem_st(dsmem_addr(&S.cov[0][lane], c), rowcov);
        dsmem_st(dsmem_addr(&S.cov[1][lane], c), colcov);
      }
      if (lane == 0) dsmem_st(dsmem_addr(&S.cmd, c), (uint32_t)kCmdShift);
    }
    cluster_sync();
    crowd_shift(S, R, cw, par);
    par ^= 1;
    shifted = true;
  }
  const int out = act == 0 ? kCmdSolved : kCmdGaveUp;
  if (lane < cw.nctas) dsmem_st(dsmem_addr(&S.cmd, lane), (uint32_t)out);
  cluster_sync();
  return act;
}

__device__ __forceinline__ int crowd_helper(CrowdCommon &S, const CrowdRemote R, const Crowd &cw, int &par) {
#pragma unroll 1
  for (;;) {
    cluster_sync();
    const int cmd = S.cmd;
    if (cmd != kCmdShift) return cmd == kCmdSolved ? 0 : 9;
    crowd_shift(S, R, cw, par);
    par ^= 1;
  }
}

// Launched with clusters of 8 or 16 CTAs (runtime cluster dimension); cluster c tracks sub-stream P.order[c]
// if its class flag is `want_a` or `want_b`.  A sub-stream that outgrows the cluster is flagged `overflow`.
__global__ void __launch_bounds__(kCrowdWarps * 32, 1)
sort_crowd_kernel(const SortParams P, const int want_a, const int want_b, const int overflow) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int BLOCK = kCrowdWarps * 32;
  extern __shared__ __align__(128) unsigned char w2t_crowd_smem[];
  CrowdCommon &S = *reinterpret_cast<CrowdCommon *>(w2t_crowd_smem);
  CrowdLeader &Ld = *reinterpret_cast<CrowdLeader *>(w2t_crowd_smem + sizeof(CrowdCommon));
  float *matrix = reinterpret_cast<float *>(w2t_crowd_smem + sizeof(CrowdCommon));  // workers only
  const int tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
  const unsigned lt = (1u << lane) - 1u;
  const int cta = (int)cluster_ctarank(), nctas = (int)cluster_nctarank();
  const bool lead_cta = cta == 0;
  const bool lead = lead_cta && warp == 0;
  const int q = P.order[blockIdx.x / nctas];
  const int cls0 = P.bail[q];
  if (cls0 != want_a && cls0 != want_b) return;  // the whole cluster takes the same exit
  const int NC = P.p.n_classes;
  const int s = q / NC, c = q - s * NC;
  const int Tcap = P.track_cap[q];
  char *slab = P.ws + P.ws_offset[q];
  const SlabLayout L = slab_layout(Tcap, P.det_cap[q]);
  double *st = reinterpret_cast<double *>(slab + L.st);   // [24][Tcap]: x[7], block-form P[13], predicted box[4], by list position
  int *tsuA = reinterpret_cast<int *>(slab + L.tsu);
  int *hsA = reinterpret_cast<int *>(slab + L.hs);
  int *bgA = reinterpret_cast<int *>(slab + L.bg);
  int *bkA = reinterpret_cast<int *>(slab + L.bk);
  const int img0 = P.p.stream_img_offsets[s], img1 = P.p.stream_img_offsets[s + 1];
  const double camW = P.p.cam_wh[2 * s], camH = P.p.cam_wh[2 * s + 1];
  const double thr = P.p.iou_thr[c];
  const float thr_f = (float)thr;
  const float4 *det_box = reinterpret_cast<const float4 *>(P.p.det_box);
  const int max_age = P.p.max_age, min_hits = P.p.min_hits;
  const bool nep50 = P.nep50 != 0;
  CrowdRemote R;
  R.z = dsmem_addr(Ld.Z, 0);
  R.zt = dsmem_addr(Ld.ZT, 0);
  R.rowflag = dsmem_addr(Ld.rowflag, 0);
  const uint32_t zbase = R.z;

  Crowd cw;
  cw.nctas = nctas;
  cw.nworkers = (nctas - 1) * kCrowdWarps;
  cw.gw = lead_cta ? -1 : warp * (nctas - 1) + (cta - 1);
  cw.mine = matrix + (size_t)warp * kCrowdMatrixFloats;

  int T = 0, frame_count = 0, err = 0, par = 0;
  bool started = false, huge = false;
  if (tid == 0) {
    S.minacc[0] = 0x7f800000u;
    S.minacc[1] = 0x7f800000u;
    S.minloc = 0x7f800000u;
    if (lead_cta) P.r.first_img[q] = -1;
  }
  cluster_sync();

#pragma unroll 1
  for (int img = img0; img <= img1; ++img) {
    // ---- the leader decides what this image needs and tells the cluster
    int cmd = kCmdExit, D = 0, base = 0, g = 0;
    if (img < img1) {
      g = img * NC + c;
      const bool exists = (P.p.img_exists == nullptr) ? true : (P.p.img_exists[img] != 0);
      D = exists ? P.p.det_count[g] : 0;
      base = P.p.det_start[g];
      if (lead_cta) {
        bool skip = !exists;
        if (!skip && !started) {
          if (D == 0) skip = true;  // no Sort object for this category yet (tracker_sort.py:32-33)
          else {
            started = true;
            if (tid == 0) P.r.first_img[q] = img - img0;
          }
        }
        if (skip || err || huge) {
          if (tid == 0) { P.r.out_count[g] = 0; P.r.created[g] = 0; }
          cmd = kCmdSkip;
        } else {
          const int n = min(D, T), m = max(D, T);
          const int nblk = (n + 7) >> 3, mw = (m + 31) >> 5;
          const int per_warp = (nblk + cw.nworkers - 1) / cw.nworkers;
          if (D > kCrowdM || T > kCrowdM || n > kCrowdN || (n > 0 && per_warp * mw * 256 > kCrowdMatrixFloats)) {
#ifdef W2T_CROWD_DEBUG
            if (tid == 0) printf("crowd overflow q=%d img=%d D=%d T=%d nctas=%d\n", q, img - img0, D, T, nctas);
#endif
            huge = true;
            cmd = kCmdSkip;
          } else {
            cmd = kCmdAssociate;
          }
        }
      }
    }
    if (lead && lane < nctas) {
      dsmem_st(dsmem_addr(&S.cmd, lane), (uint32_t)cmd);
      dsmem_st(dsmem_addr(&S.hdr_T, lane), (uint32_t)T);
    }
    cluster_sync();  // also: the predicted boxes the leader wrote for this image are visible to every CTA
    cmd = S.cmd;
    T = S.hdr_T;
    if (cmd == kCmdExit) break;
    if (cmd == kCmdSkip) continue;
    if (lead_cta) {
      frame_count++;
      // the leader CTA stages the detections and the predicted boxes (read-out, Kalman pass)
      for (int d = tid; d < D; d += BLOCK) { Ld.det[d] = __ldg(det_box + base + d); Ld.dstat[d] = 0; }
      for (int t = tid; t < T; t += BLOCK) {
#pragma unroll
        for (int k = 0; k < 4; k++) Ld.box[k][t] = st[(size_t)(kBoxAt + k) * Tcap + t];
        Ld.match[t] = -1;
      }
      __syncthreads();
    }

    // ---- A. association
    if (T > 0 && D > 0) {
      const bool flipped = D > T;
      cw.n = flipped ? T : D;
      cw.m = flipped ? D : T;
      cw.mw = (cw.m + 31) >> 5;
      cw.nblk = (cw.n + 7) >> 3;
      const int n = cw.n, m = cw.m, mw = cw.mw;
      if (lead_cta) {
        for (int i = tid; i < n; i += BLOCK) { Ld.row_star[i] = -1; Ld.row_prime[i] = -1; }
        for (int i = tid; i < m; i += BLOCK) Ld.col_star[i] = -1;
      } else {
        // workers: dense construction, lanes own columns: -IoU, row minimum, reduced costs, zero words -> the
        // leader's Z.  Detections come from global memory, predicted boxes from the slab (through L2: another
        // SM wrote them).
#pragma unroll 1
        for (int b = cw.gw; b < cw.nblk; b += cw.nworkers) {
          const int r0 = b * 8;
          uint32_t mn_u[8];
#pragma unroll
          for (int j = 0; j < 8; j++) mn_u[j] = 0xffffffffu;
#pragma unroll 1
          for (int k = 0; k < mw; k++) {
            const int cc = k * 32 + lane;
            // this lane's column: a detection (flipped) or a tracker box
            float4 cd = make_float4(0.f, 0.f, 0.f, 0.f);
            double c0 = 0., c1 = 0., c2 = 0., c3 = 0.;
            if (cc < m) {
              if (flipped) cd = __ldg(det_box + base + cc);
              else {
                c0 = __ldcg(st + (size_t)(kBoxAt + 0) * Tcap + cc); c1 = __ldcg(st + (size_t)(kBoxAt + 1) * Tcap + cc);
                c2 = __ldcg(st + (size_t)(kBoxAt + 2) * Tcap + cc); c3 = __ldcg(st + (size_t)(kBoxAt + 3) * Tcap + cc);
              }
            }
            float *p = cw.cell(b, k) + lane;
#pragma unroll 1
            for (int j = 0; j < 8; j++) {
              const int r = min(r0 + j, n - 1);  // padding rows repeat the last row; they are never read
              float v = 0.0f;
              if (cc < m) {
                if (flipped) {
                  v = -iou_pair_call(cd, __ldcg(st + (size_t)(kBoxAt + 0) * Tcap + r), __ldcg(st + (size_t)(kBoxAt + 1) * Tcap + r),
                                     __ldcg(st + (size_t)(kBoxAt + 2) * Tcap + r), __ldcg(st + (size_t)(kBoxAt + 3) * Tcap + r));
                } else {
                  v = -iou_pair_call(__ldg(det_box + base + r), c0, c1, c2, c3);
                }
                mn_u[j] = min(mn_u[j], Munkres<32>::ordered(v));
              }
              p[j * 32] = v;
            }
          }
#pragma unroll
          for (int j = 0; j < 8; j++) mn_u[j] = __reduce_min_sync(FULL, mn_u[j]);
#pragma unroll 1
          for (int k = 0; k < mw; k++) {
            const int cc = k * 32 + lane;
            float *p = cw.cell(b, k) + lane;
            uint32_t zbyte = 0u, myword = 0u;
#pragma unroll
            for (int j = 0; j < 8; j++) {
              const float v = p[j * 32] - Munkres<32>::unordered(mn_u[j]);
              p[j * 32] = v;
              const bool z = cc < m && r0 + j < n && v == 0.0f;
              zbyte |= z ? (1u << j) : 0u;
              const uint32_t word = __ballot_sync(FULL, z);
              if (lane == j) myword = word;
            }
            if (lane < 8) dsmem_st(zbase + 4u * (uint32_t)(k * kCrowdZS + r0 + lane), myword);
            dsmem_st8(R.zt + (uint32_t)(b * kCrowdM + cc), zbyte);
          }
        }
      }
      cluster_sync();
      const int act = lead ? crowd_munkres(S, Ld, R, cw, par) : crowd_helper(S, R, cw, par);
      if (act != 0) err = W2T_ERR_ARG;
      if (lead_cta) {
        __syncthreads();
        // matched / rejected detections (sort.py:217-222)
        for (int d = tid; d < D; d += BLOCK) {
          const int t = flipped ? Ld.col_star[d] : Ld.row_star[d];
          if (t >= 0) {
            const float o = iou_pair_call(Ld.det[d], Ld.box[0][t], Ld.box[1][t], Ld.box[2][t], Ld.box[3][t]);
            const bool rejected = nep50 ? (o < thr_f) : ((double)o < thr);
            if (rejected) Ld.dstat[d] = 2;
            else { Ld.dstat[d] = 1; Ld.match[t] = (int16_t)d; }
          }
        }
        __syncthreads();
      }
    }
    if (!lead_cta) continue;

    // ---- CTA 0: new trackers (unassigned detections first, then the rejected ones), then the Kalman pass
    if (warp == 0) {
      int n_new = 0;
#pragma unroll 1
      for (int pass = 0; pass < 2; pass++) {
#pragma unroll 1
        for (int d0 = 0; d0 < D; d0 += 32) {
          const int d = d0 + lane;
          const bool a = d < D && Ld.dstat[d] == (pass ? 2 : 0);
          const unsigned bb = __ballot_sync(FULL, a);
          if (a) Ld.newdet[n_new + __popc(bb & lt)] = (int16_t)d;
          n_new += __popc(bb);
        }
      }
      if (lane == 0) Ld.bcast[0] = n_new;
    }
    __syncthreads();
    const int n_new = Ld.bcast[0];
    const int Ttot = T + n_new;
    if (Ttot > Tcap) err = W2T_ERR_CAPACITY;
    if (err) {
      if (tid == 0) { P.r.out_count[g] = 0; P.r.created[g] = 0; }
      continue;
    }
    int n_live = 0, emitted = 0;
#pragma unroll 1
    for (int tp = 0, pr = 0; tp < Ttot; tp += BLOCK, pr ^= 1) {
      const int t = tp + tid;
      bool ok = false, surv = false;
      double x[7], Pm[kBlockP], ob0 = 0., ob1 = 0., ob2 = 0., ob3 = 0., oconf = 0., nb[4];
      int tsu = 0, hs = 0, obg = 0, obk = 0;
      if (t < Ttot) {
        int md;
        if (t >= T) {
          md = Ld.newdet[t - T];
          obg = g;
          obk = t - T;
        } else {
#pragma unroll
          for (int k = 0; k < 7; k++) x[k] = st[(size_t)k * Tcap + t];
#pragma unroll
          for (int k = 0; k < kBlockP; k++) Pm[k] = st[(size_t)(7 + k) * Tcap + t];
          tsu = tsuA[t];
          hs = hsA[t];
          obg = bgA[t];
          obk = bkA[t];
          md = Ld.match[t];
        }
        if (md >= 0) {
          const float4 d4 = Ld.det[md];
          double z[4];
          bbox_to_z_d(d4.x, d4.y, d4.z, d4.w, nep50, z);
          if (t >= T) kfb_init_z(z, x, Pm);
          else {
            kfb_update_z(x, Pm, z);
            tsu = 0;
            hs += 1;
          }
        }
        if (tsu < 1 && (hs >= min_hits || frame_count <= min_hits)) {
          double w, h;
          box_wh(x[2], x[3], w, h);
          const double e = ((Pm[0] + Pm[4]) + Pm[8]) / 3.0;
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
        surv = !(tsu > max_age);
        if (surv) {
          kfb_predict(x, Pm);
          if (tsu > 0) hs = 0;
          tsu += 1;
          double w, h;
          box_wh(x[2], x[3], w, h);
          nb[0] = x[0] - w / 2.; nb[1] = x[1] - h / 2.; nb[2] = x[0] + w / 2.; nb[3] = x[1] + h / 2.;
          if (isnan(nb[0]) || isnan(nb[1]) || isnan(nb[2]) || isnan(nb[3])) surv = false;
          else if (isinf(nb[0]) || isinf(nb[1]) || isinf(nb[2]) || isinf(nb[3])) atomicMax(P.status, W2T_ERR_NONFINITE);
        }
      }
      const unsigned bs = __ballot_sync(FULL, surv), be = __ballot_sync(FULL, ok);
      if (lane == 0) { Ld.cnt[pr][warp][0] = __popc(bs); Ld.cnt[pr][warp][1] = __popc(be); }
      __syncthreads();  // also: every thread has read its old position before any overwrites one
      int live0 = n_live, emit0 = emitted, lt_ = 0, et_ = 0;
#pragma unroll
      for (int w = 0; w < kCrowdWarps; w++) {
        const int l = Ld.cnt[pr][w][0], e = Ld.cnt[pr][w][1];
        if (w < warp) { live0 += l; emit0 += e; }
        lt_ += l;
        et_ += e;
      }
      n_live += lt_;
      emitted += et_;
      if (surv) {
        const int pos = live0 + __popc(bs & lt);
#pragma unroll
        for (int k = 0; k < 7; k++) st[(size_t)k * Tcap + pos] = x[k];
#pragma unroll
        for (int k = 0; k < kBlockP; k++) st[(size_t)(7 + k) * Tcap + pos] = Pm[k];
#pragma unroll
        for (int k = 0; k < 4; k++) st[(size_t)(kBoxAt + k) * Tcap + pos] = nb[k];
        tsuA[pos] = tsu;
        hsA[pos] = hs;
        bgA[pos] = obg;
        bkA[pos] = obk;
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
    if (tid == 0) { P.r.out_count[g] = emitted; P.r.created[g] = n_new; }
    T = n_live;
    __syncthreads();
  }

  if (!lead_cta) return;
  if (huge) {  // tracked again from its first image by a bigger cluster / the global-memory CTA kernel
    if (tid == 0) P.bail[q] = overflow;
    return;
  }
  if (tid == 0) P.bail[q] = kClsDone;
  if (err && tid == 0) atomicMax(P.status, err);
  if (P.r.final_count != nullptr) {
    if (tid == 0) P.r.final_count[q] = T;
    if (P.r.final_state != nullptr) {
      const int cap = P.r.final_cap;
      for (int t = tid; t < T && t < cap; t += BLOCK) {
        double *dst = P.r.final_state + ((size_t)q * cap + t) * 56;
        double pb[kBlockP], Pd[49];
        for (int k = 0; k < 7; k++) dst[k] = st[(size_t)k * Tcap + t];
        for (int k = 0; k < kBlockP; k++) pb[k] = st[(size_t)(7 + k) * Tcap + t];
        kfb_to_dense(pb, Pd);
        for (int k = 0; k < 49; k++) dst[7 + k] = Pd[k];
      }
    }
  }
  if (P.chunk_done != nullptr) {
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      atomicAdd(&P.chunk_done[P.chunk_of[q]], 1);
    }
  }
}

}  // namespace w2t

// munkres.cuh — exact emulation of scikit-learn 0.22.2 `linear_assignment` (Kuhn-Munkres,
// matrix form, float32) by one CTA.
//
// Replaces the call at tracking/sort/sort.py:206.  The solver is third-party code the
// reference imports (sklearn/utils/linear_assignment_.py, restated in SURVEY.md Appendix A);
// which optimal assignment it returns matters (zero-IoU ties with iou_threshold = 0), so the
// step machine is reproduced step for step rather than replaced by a dual-potential solver:
//   * float32 in-place arithmetic, including step 6's separate "+= min" / "-= min";
//   * every "first zero" search is row-major;
//   * covers, stars and primes evolve exactly as in the reference implementation.
//
// Representation.  The reference keeps an n x m `marked` matrix; a row holds at most one
// star and one prime and a column at most one star, so three index arrays carry the same
// state.  Zeros of the cost matrix are mirrored in a bit matrix Z (one 32-bit word per 32
// columns) that is updated only when the costs change (step 1 and step 6); with the column
// cover kept as a bit mask too, "first uncovered zero of a row" is an AND + find-first-set
// per word instead of a scan over floats.
//
// The serial parts are restructured so that their length is the number of CONFLICTS, not the
// number of rows, without changing any decision the reference makes:
//   * step 2 (greedy stars in row-major order): 32 rows at a time propose their first free
//     zero; the longest prefix of rows whose proposals are pairwise distinct is exactly what
//     the sequential loop would star, so it commits at once (__match_any_sync finds the
//     first clash), and only the clashing rows go around again;
//   * step 4 ("first uncovered zero"): a bit mask of rows that currently own an uncovered
//     zero is kept up to date — inside a step-4 run columns only get uncovered and rows only
//     get covered, so a row can only ENTER the set when a column holding one of its zeros is
//     uncovered — and the search is a find-first-set;
//   * step 6 only changes covered rows (+min) and uncovered columns (-min) and its minimum
//     only looks at uncovered rows x uncovered columns; warp 0 compacts both index lists
//     from the cover masks and the CTA touches just those cells.
// Steps 3-5 run in warp 0 while the other warps wait at the CTA barrier; step 1 and step 6
// use the whole CTA.
//
// Storage.  All arrays are reached through plain pointers: the caller points them at shared
// memory when the problem fits and at its global workspace otherwise.  `ldc` (the row pitch
// of C) and `zs` are odd so that "one thread per row" accesses are bank-conflict free.
#pragma once

#include "common.cuh"

namespace w2t {

constexpr int kMunkresMaxWords = 64;                      // mask words kept in shared memory
constexpr int kMunkresMaxDim = kMunkresMaxWords * 32;     // largest max(D, T)

struct MunkresShared {
  uint32_t colcov[kMunkresMaxWords];
  uint32_t rowcov[kMunkresMaxWords];
  uint32_t starcols[kMunkresMaxWords];
  uint32_t rowhas[kMunkresMaxWords];  // rows that own an uncovered zero (step 4)
  float red[32];
  int ctl[4];  // 0: next action (0 done, 6 shift, 9 error)  1: #uncovered cols  2: #covered rows (step 6)
};

struct MunkresGlobal {
  float *C;         // [n*ldc] cost, n <= m, row-major with pitch ldc
  uint32_t *Z;      // [n*zs] zero bit matrix
  int *row_star;    // [n] column of the row's star or -1
  int *col_star;    // [m] row of the column's star or -1
  int *row_prime;   // [n] column of the row's prime (valid for rows primed in the current phase)
  int *ucols;       // [m] scratch: uncovered columns (step 6)
  int *crows;       // [n] scratch: covered rows (step 6)
};

__host__ __device__ inline int munkres_pitch(int m) { return m | 1; }
__host__ __device__ inline int munkres_words(int m) { return (m + 31) >> 5; }
__host__ __device__ inline int munkres_zstride(int m) { return munkres_words(m) | 1; }

template <int BLOCK, bool TIMERS = false>
struct Munkres {
  static constexpr int NW = BLOCK / 32;
  int n, m, mw, zs, ldc;
  bool rowwise;  // arrays are in shared memory: one thread per row is the fast layout for step 1
  MunkresGlobal g;
  MunkresShared *s;
  long long *ph;  // phase cycle counters (debug aid, TIMERS only): [3] reduce [4] greedy [5] drive [6] shift
  long long t_last;

  __device__ __forceinline__ void tick(int i) {
    if (TIMERS && threadIdx.x == 0) {
      const long long now = clock64();
      ph[i] += now - t_last;
      t_last = now;
    }
  }

  __device__ __forceinline__ bool row_covered(int r) const { return (s->rowcov[r >> 5] >> (r & 31)) & 1u; }

  // ---- step 1: subtract the row minimum, build Z ------------------------------------------
  __device__ void reduce_rows() {
    if (rowwise) {
      for (int r = threadIdx.x; r < n; r += BLOCK) {
        float *row = g.C + (size_t)r * ldc;
        float mn = row[0];
        for (int c = 1; c < m; c++) mn = fminf(mn, row[c]);
        for (int k = 0; k < mw; k++) {
          uint32_t word = 0u;
          const int c1 = min(32, m - k * 32);
          for (int b = 0; b < c1; b++) {
            const float v = row[k * 32 + b] - mn;
            row[k * 32 + b] = v;
            word |= (v == 0.0f) ? (1u << b) : 0u;
          }
          g.Z[(size_t)r * zs + k] = word;
        }
      }
      return;
    }
    const int lane = lane_id();
    for (int r = warp_id(); r < n; r += NW) {
      float *row = g.C + (size_t)r * ldc;
      float mn = row[0];
      for (int c = lane; c < m; c += 32) mn = fminf(mn, row[c]);
#pragma unroll
      for (int o = 16; o; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      for (int k = 0; k < mw; k++) {
        const int c = k * 32 + lane;
        bool z = false;
        if (c < m) {
          const float v = row[c] - mn;
          row[c] = v;
          z = (v == 0.0f);
        }
        const unsigned word = __ballot_sync(0xffffffffu, z);
        if (lane == 0) g.Z[(size_t)r * zs + k] = word;
      }
    }
  }

  // first zero of row r that is not in the column cover, or -1 (one lane per row)
  __device__ __forceinline__ int first_free_zero_lane(int r) const {
    const uint32_t *zr = g.Z + (size_t)r * zs;
    for (int w = 0; w < mw; w++) {
      const uint32_t v = zr[w] & ~s->colcov[w];
      if (v) return w * 32 + (__ffs(v) - 1);
    }
    return -1;
  }

  // ---- step 2: star zeros greedily in row-major order (warp 0) ---------------------------------
  // s->colcov is used as the running column cover; on return it equals s->starcols.
  __device__ int greedy_stars() {
    const int lane = lane_id();
    const unsigned lt = (1u << lane) - 1u;
    for (int w = lane; w < kMunkresMaxWords; w += 32) s->colcov[w] = 0u;
    __syncwarp();
    int stars = 0;
    for (int rb = 0; rb < n; rb += 32) {
      const int r = rb + lane;
      bool pending = r < n;
      for (;;) {
        int cand = -1;
        if (pending) {
          cand = first_free_zero_lane(r);
          if (cand < 0) pending = false;  // every zero of the row is taken: no star, like the reference
        }
        if (!__ballot_sync(0xffffffffu, pending)) break;
        // rows whose proposal equals that of an earlier pending row must wait for the next round
        const unsigned peers = __match_any_sync(0xffffffffu, pending ? cand : (-2 - lane));
        const unsigned clash = __ballot_sync(0xffffffffu, pending && (peers & lt) != 0u);
        const int first_clash = clash ? (__ffs(clash) - 1) : 32;
        const bool commit = pending && lane < first_clash;
        if (commit) {
          g.row_star[r] = cand;
          g.col_star[cand] = r;
          atomicOr(&s->colcov[cand >> 5], 1u << (cand & 31));
          pending = false;
        }
        stars += __popc(__ballot_sync(0xffffffffu, commit));
        __syncwarp();
      }
    }
    for (int w = lane; w < kMunkresMaxWords; w += 32) s->starcols[w] = s->colcov[w];
    __syncwarp();
    return stars;
  }

  // first uncovered zero of row r: column index or -1 (warp 0, all lanes cooperate)
  __device__ __forceinline__ int first_open_zero(int r) {
    const int lane = lane_id();
    const uint32_t *zr = g.Z + (size_t)r * zs;
    for (int k0 = 0; k0 < mw; k0 += 32) {
      const int w = k0 + lane;
      const uint32_t v = (w < mw) ? (zr[w] & ~s->colcov[w]) : 0u;
      const unsigned b = __ballot_sync(0xffffffffu, v != 0u);
      if (b) {
        const int src = __ffs(b) - 1;
        const uint32_t vv = __shfl_sync(0xffffffffu, v, src);
        return (k0 + src) * 32 + (__ffs(vv) - 1);
      }
    }
    return -1;
  }

  // rowhas from scratch (warp 0): uncovered rows with a zero in an uncovered column
  __device__ void compute_rowhas() {
    const int lane = lane_id();
    for (int rb = 0; rb < n; rb += 32) {
      const int r = rb + lane;
      const bool any = (r < n) && !row_covered(r) && (first_free_zero_lane(r) >= 0);
      const unsigned b = __ballot_sync(0xffffffffu, any);
      if (lane == 0) s->rowhas[rb >> 5] = b;
    }
    __syncwarp();
  }

  // warp 0: compact the uncovered columns and the covered rows into index lists (step 6 prologue)
  __device__ void build_step6_lists() {
    const int lane = lane_id();
    const unsigned lt = (1u << lane) - 1u;
    int nu = 0, nc = 0;
    for (int w = 0; w < mw; w++) {
      const int c = w * 32 + lane;
      const bool open = (c < m) && !((s->colcov[w] >> lane) & 1u);
      const unsigned b = __ballot_sync(0xffffffffu, open);
      if (open) g.ucols[nu + __popc(b & lt)] = c;
      nu += __popc(b);
    }
    const int nwr = munkres_words(n);
    for (int w = 0; w < nwr; w++) {
      const int r = w * 32 + lane;
      const bool cov = (r < n) && ((s->rowcov[w] >> lane) & 1u);
      const unsigned b = __ballot_sync(0xffffffffu, cov);
      if (cov) g.crows[nc + __popc(b & lt)] = r;
      nc += __popc(b);
    }
    if (lane == 0) { s->ctl[1] = nu; s->ctl[2] = nc; }
  }

  // ---- steps 3, 4, 5 (warp 0): returns 0 when n stars exist, 6 when the costs must shift -----
  __device__ int drive(int step, int &stars, int &budget) {
    const int lane = lane_id();
    const int nwr = munkres_words(n);
    for (;;) {
      if (step == 3) {
        // cover every starred column, uncover all rows
        for (int w = lane; w < kMunkresMaxWords; w += 32) {
          s->rowcov[w] = 0u;
          s->colcov[w] = (w < mw) ? s->starcols[w] : 0u;
        }
        __syncwarp();
        if (stars >= n) return 0;
      }
      compute_rowhas();  // entering step 4: after a cover reset (step 3) or a cost shift (step 6)
      for (;;) {
        // step 4: first uncovered zero in row-major order
        if (--budget < 0) return 9;
        int fr = -1;
        for (int w0 = 0; w0 < nwr; w0 += 32) {
          const int w = w0 + lane;
          const uint32_t v = (w < nwr) ? s->rowhas[w] : 0u;
          const unsigned b = __ballot_sync(0xffffffffu, v != 0u);
          if (b) {
            const int src = __ffs(b) - 1;
            const uint32_t vv = __shfl_sync(0xffffffffu, v, src);
            fr = (w0 + src) * 32 + (__ffs(vv) - 1);
            break;
          }
        }
        if (fr < 0) {
          build_step6_lists();
          return 6;
        }
        const int fc = first_open_zero(fr);
        const int sc = g.row_star[fr];
        if (sc < 0) {
          // step 5: augment along the alternating path that starts at the primed zero (fr, fc)
          if (lane == 0) {
            int r = fr, c = fc;
            for (int hops = 0;; hops++) {
              const int rs = g.col_star[c];
              g.row_star[r] = c;
              g.col_star[c] = r;
              if (rs < 0) {
                s->starcols[c >> 5] |= (1u << (c & 31));
                break;
              }
              r = rs;
              c = g.row_prime[r];
              if (c < 0 || hops > n + m) { budget = -1; break; }  // cannot happen in a valid state
            }
          }
          __syncwarp();
          budget = __shfl_sync(0xffffffffu, budget, 0);
          if (budget < 0) return 9;
          stars++;
          step = 3;
          break;
        }
        // the row has a star: prime the zero, cover the row, uncover the star's column
        if (lane == 0) {
          g.row_prime[fr] = fc;
          s->rowcov[fr >> 5] |= (1u << (fr & 31));
          s->colcov[sc >> 5] &= ~(1u << (sc & 31));
          s->rowhas[fr >> 5] &= ~(1u << (fr & 31));
        }
        __syncwarp();
        // uncovered rows with a zero in the newly uncovered column now own an uncovered zero
        for (int rb = 0; rb < n; rb += 32) {
          const int r = rb + lane;
          const bool has = (r < n) && !row_covered(r) && ((g.Z[(size_t)r * zs + (sc >> 5)] >> (sc & 31)) & 1u);
          const unsigned b = __ballot_sync(0xffffffffu, has);
          if (lane == 0 && b) s->rowhas[rb >> 5] |= b;
        }
        __syncwarp();
      }
    }
  }

  // ---- step 6 (all threads): C[covered rows] += min ; C[:, uncovered cols] -= min ------------
  // min over uncovered rows x uncovered columns.  Only those cells are visited.
  __device__ void shift_by_min() {
    const int lane = lane_id(), warp = warp_id();
    const int nu = s->ctl[1], nc = s->ctl[2];
    float mn = INFINITY;
    for (int r = threadIdx.x; r < n; r += BLOCK) {
      if (row_covered(r)) continue;
      const float *row = g.C + (size_t)r * ldc;
      for (int u = 0; u < nu; u++) mn = fminf(mn, row[g.ucols[u]]);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    if (lane == 0) s->red[warp] = mn;
    __syncthreads();
    mn = s->red[0];
#pragma unroll
    for (int i = 1; i < NW; i++) mn = fminf(mn, s->red[i]);
    if (mn == INFINITY) return;  // nothing uncovered: the reference leaves the matrix alone
    // covered rows: every column gets +min, the uncovered ones then -min (two float32 roundings)
    for (int i = warp; i < nc; i += NW) {
      const int r = g.crows[i];
      float *row = g.C + (size_t)r * ldc;
      for (int k = 0; k < mw; k++) {
        const int c = k * 32 + lane;
        bool z = false;
        if (c < m) {
          float v = row[c] + mn;
          if (!((s->colcov[k] >> lane) & 1u)) v = v - mn;
          row[c] = v;
          z = (v == 0.0f);
        }
        const unsigned word = __ballot_sync(0xffffffffu, z);
        if (lane == 0) g.Z[(size_t)r * zs + k] = word;
      }
    }
    // uncovered rows: only the uncovered columns change (-min); one thread owns a row's Z words
    for (int r = threadIdx.x; r < n; r += BLOCK) {
      if (row_covered(r)) continue;
      float *row = g.C + (size_t)r * ldc;
      uint32_t *zr = g.Z + (size_t)r * zs;
      for (int u = 0; u < nu; u++) {
        const int c = g.ucols[u];
        const float v = row[c] - mn;
        row[c] = v;
        const uint32_t bit = 1u << (c & 31);
        const uint32_t zw = zr[c >> 5];
        zr[c >> 5] = (v == 0.0f) ? (zw | bit) : (zw & ~bit);
      }
    }
  }

  // =============================================================================================
  // Small problems (n <= m <= 64): every mask is one 64-bit word.
  //
  // Step 1 uses the CTA (one thread per row); everything else runs in warp 0 with NO
  // intra-warp communication on the serial path: covers, the "rows owning an uncovered zero"
  // set and the star count live in registers and all 32 lanes execute the same scalar
  // instructions on the same values, so one step-4 iteration is two dependent shared-memory
  // loads plus a dozen integer ops (no ballot, shuffle or barrier).  Zr[r] / Zc[c] are the
  // zero bit matrix by row and by column (g.Z reinterpreted: 64 + 64 words of 64 bits);
  // uncovering column sc adds Zc[sc] & ~rowcov to the row set.  Lanes spread over rows /
  // columns only where the data is wide: the initial row set and step 6.
  // =============================================================================================
  typedef unsigned long long u64;

  __device__ __forceinline__ static u64 low_mask(int k) { return k >= 64 ? ~0ull : ((1ull << k) - 1ull); }

  __device__ int solve_small() {
    u64 *Zr = reinterpret_cast<u64 *>(g.Z), *Zc = Zr + 64;
    const int lane = lane_id();
    // ---- step 1 (CTA): row minimum, subtract, zero masks by row (one thread per row) ...
    for (int c = threadIdx.x; c < m; c += BLOCK) g.col_star[c] = -1;
    for (int r = threadIdx.x; r < n; r += BLOCK) {
      float *row = g.C + (size_t)r * ldc;
      float mn = row[0];
      for (int c = 1; c < m; c++) mn = fminf(mn, row[c]);
      u64 zr = 0ull;
      for (int c = 0; c < m; c++) {
        const float v = row[c] - mn;
        row[c] = v;
        zr |= (v == 0.0f) ? (1ull << c) : 0ull;
      }
      Zr[r] = zr;
      g.row_star[r] = -1;
      g.row_prime[r] = -1;
    }
    __syncthreads();
    // ... and by column (one thread per column; consecutive threads read consecutive floats)
    for (int c = threadIdx.x; c < m; c += BLOCK) {
      u64 zc = 0ull;
      for (int r = 0; r < n; r++) zc |= (g.C[(size_t)r * ldc + c] == 0.0f) ? (1ull << r) : 0ull;
      Zc[c] = zc;
    }
    __syncthreads();
    tick(3);
    if (warp_id() == 0) {
      const u64 nmask = low_mask(n), mmask = low_mask(m);
      // ---- step 2: greedy stars in row-major order.  Lane L keeps rows L and L+32 in registers;
      // the serial loop gets row r by shuffle, so its only loop-carried chain is the cover word.
      u64 starcols = 0ull;
      int stars = 0;
      {
        const u64 z0 = (lane < n) ? Zr[lane] : 0ull, z1 = (lane + 32 < n) ? Zr[lane + 32] : 0ull;
        int rs0 = -1, rs1 = -1;  // star column of rows lane, lane+32
        int cs0 = -1, cs1 = -1;  // star row of columns lane, lane+32
        for (int r = 0; r < n; r++) {
          const u64 zr = __shfl_sync(0xffffffffu, (r < 32) ? z0 : z1, r & 31);
          const u64 v = zr & ~starcols;
          if (v) {
            const int c = __ffsll((long long)v) - 1;
            if (lane == (r & 31)) { if (r < 32) rs0 = c; else rs1 = c; }
            if (lane == (c & 31)) { if (c < 32) cs0 = r; else cs1 = r; }
            starcols |= 1ull << c;
            stars++;
          }
        }
        if (lane < n) g.row_star[lane] = rs0;
        if (lane + 32 < n) g.row_star[lane + 32] = rs1;
        if (lane < m) g.col_star[lane] = cs0;
        if (lane + 32 < m) g.col_star[lane + 32] = cs1;
      }
      // Every lane executes the scalar state machine redundantly on identical values; lanes are
      // re-aligned (and their shared-memory accesses ordered) wherever one lane could otherwise
      // overwrite a word another lane has yet to read.
      __syncwarp();
      tick(4);
      int act = 0;
      int budget = 4 * n * n + 64 * (n + m) + 1024;
      while (stars < n) {
        // ---- step 3: cover the starred columns, uncover all rows
        u64 rowcov = 0ull, colcov = starcols;
        for (;;) {
          // rows that own an uncovered zero, from scratch (entering step 4 / after a cost shift)
          u64 rowhas;
          {
            const int r0 = lane, r1 = lane + 32;
            const bool h0 = (r0 < n) && !((rowcov >> r0) & 1ull) && (Zr[r0] & ~colcov) != 0ull;
            const bool h1 = (r1 < n) && !((rowcov >> r1) & 1ull) && (Zr[r1] & ~colcov) != 0ull;
            const unsigned lo = __ballot_sync(0xffffffffu, h0), hi = __ballot_sync(0xffffffffu, h1);
            rowhas = (u64)lo | ((u64)hi << 32);
          }
          bool augmented = false;
          // ---- step 4: prime uncovered zeros in row-major order
          while (rowhas) {
            if (--budget < 0) { act = 9; break; }
            const int fr = __ffsll((long long)rowhas) - 1;
            const int sc = g.row_star[fr];
            const int fc = __ffsll((long long)(Zr[fr] & ~colcov)) - 1;
            if (sc < 0) {
              // ---- step 5: flip stars along the alternating path from the primed zero (fr, fc)
              int r = fr, c = fc;
              __syncwarp();
              for (int hops = 0;; hops++) {
                const int rs = g.col_star[c];
                __syncwarp();  // all lanes have read the old star before any lane replaces it
                g.row_star[r] = c;
                g.col_star[c] = r;
                if (rs < 0) {
                  starcols |= 1ull << c;
                  break;
                }
                r = rs;
                c = g.row_prime[r];
                if (c < 0 || hops > n + m) { act = 9; break; }  // cannot happen in a valid state
              }
              __syncwarp();
              stars++;
              augmented = true;
              break;
            }
            g.row_prime[fr] = fc;
            rowcov |= 1ull << fr;
            colcov &= ~(1ull << sc);
            rowhas = (rowhas & ~(1ull << fr)) | (Zc[sc] & ~rowcov & nmask);
          }
          if (augmented || act) break;
          __syncwarp();
          tick(5);
          // ---- step 6: min over uncovered rows x uncovered columns; covered rows += min,
          //      uncovered columns -= min (float32, in that order)
          if (--budget < 0) { act = 9; break; }
          const u64 ucols = ~colcov & mmask, urows = ~rowcov & nmask;
          float mn = INFINITY;
#pragma unroll
          for (int half = 0; half < 2; half++) {
            const int r = lane + 32 * half;
            if ((urows >> r) & 1ull) {
              const float *row = g.C + (size_t)r * ldc;
              for (u64 rem = ucols; rem; rem &= rem - 1ull) mn = fminf(mn, row[__ffsll((long long)rem) - 1]);
            }
          }
#pragma unroll
          for (int o = 16; o; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
          if (mn == INFINITY) { act = 9; break; }
          for (u64 rem = rowcov & nmask; rem; rem &= rem - 1ull) {
            const int r = __ffsll((long long)rem) - 1;
            float *row = g.C + (size_t)r * ldc;
            u64 zbits = 0ull;
#pragma unroll
            for (int half = 0; half < 2; half++) {
              const int c = lane + 32 * half;
              bool z = false;
              if (c < m) {
                float v = row[c] + mn;
                if ((ucols >> c) & 1ull) v = v - mn;
                row[c] = v;
                z = (v == 0.0f);
                Zc[c] = (Zc[c] & ~(1ull << r)) | ((u64)z << r);
              }
              zbits |= (u64)__ballot_sync(0xffffffffu, z) << (32 * half);
            }
            Zr[r] = zbits;
          }
          __syncwarp();  // Zc words written by their owning lanes above are read by every lane below
          {
            const int r0 = lane, r1 = lane + 32;
            const bool a0 = (urows >> r0) & 1ull, a1 = (urows >> r1) & 1ull;
            float *row0 = g.C + (size_t)r0 * ldc, *row1 = g.C + (size_t)r1 * ldc;
            u64 zr0 = a0 ? Zr[r0] : 0ull, zr1 = a1 ? Zr[r1] : 0ull;
            for (u64 rem = ucols; rem; rem &= rem - 1ull) {
              const int c = __ffsll((long long)rem) - 1;
              bool f0 = false, f1 = false;
              if (a0) {
                const float v = row0[c] - mn;
                row0[c] = v;
                f0 = (v == 0.0f);
                zr0 = (zr0 & ~(1ull << c)) | ((u64)f0 << c);
              }
              if (a1) {
                const float v = row1[c] - mn;
                row1[c] = v;
                f1 = (v == 0.0f);
                zr1 = (zr1 & ~(1ull << c)) | ((u64)f1 << c);
              }
              const u64 colbits = (u64)__ballot_sync(0xffffffffu, f0) | ((u64)__ballot_sync(0xffffffffu, f1) << 32);
              Zc[c] = (Zc[c] & ~urows) | colbits;
            }
            if (a0) Zr[r0] = zr0;
            if (a1) Zr[r1] = zr1;
          }
          __syncwarp();
          tick(6);
        }
        if (act) break;
      }
      tick(5);
      if (lane == 0) s->ctl[0] = act;
    }
    __syncthreads();
    return s->ctl[0];
  }

  // Whole solve.  Precondition: g.C holds the n x m costs (n <= m).  All threads call it.
  // Returns 0, or 9 if the iteration budget ran out (malformed input such as NaN costs).
  __device__ int solve() {
    if (rowwise && m <= 64) return solve_small();
    for (int i = threadIdx.x; i < n; i += BLOCK) { g.row_star[i] = -1; g.row_prime[i] = -1; }
    for (int i = threadIdx.x; i < m; i += BLOCK) g.col_star[i] = -1;
    reduce_rows();
    __syncthreads();
    tick(3);
    int stars = 0, step = 3;
    int budget = 4 * n * n + 64 * (n + m) + 1024;
    if (warp_id() == 0) stars = greedy_stars();
    tick(4);
    for (;;) {
      if (warp_id() == 0) {
        const int act = drive(step, stars, budget);
        if (lane_id() == 0) s->ctl[0] = act;
        step = 4;
      }
      __syncthreads();
      tick(5);
      const int act = s->ctl[0];
      if (act != 6) return act;
      shift_by_min();
      __syncthreads();
      tick(6);
    }
  }
};

}  // namespace w2t

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

#ifndef W2T_SOLVE_BLOCK_INLINE
#define W2T_SOLVE_BLOCK_INLINE __forceinline__
#endif

namespace w2t {

constexpr int kMunkresMaxWords = 128;                     // mask words kept in shared memory
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
          __syncwarp();  // every lane has read row_star[fr] before lane 0 rewrites the stars
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
  // Problems up to 128 x 128 whose arrays live in shared memory (the common case).
  //
  // Every cover / star mask is four 32-bit words held in REGISTERS of warp 0, identical in all
  // lanes, so the serial state machine (steps 3-5) spends its time on a handful of broadcast
  // shared-memory loads per iteration instead of mask traffic.  The wide parts use the data
  // layout that suits them: step 1 one thread per row, step 2 the clash-free-prefix trick of
  // greedy_stars(), step 6 the whole CTA with one warp per row and lanes across columns
  // (conflict-free, a few hundred cycles per round).  Z rows are kept in shared memory with an
  // odd word stride; "rows that gained an uncovered zero when column sc was uncovered" is one
  // conflict-free load + ballot per 32 rows, so no column-major copy of Z has to be maintained.
  // =============================================================================================
  // Four named words rather than an array: nothing can turn a bit update into a dynamically
  // indexed (= local memory) access.
  struct M4 {
    uint32_t a, b, c, d;
    __device__ __forceinline__ int first() const {
      if (a) return __ffs(a) - 1;
      if (b) return 32 + __ffs(b) - 1;
      if (c) return 64 + __ffs(c) - 1;
      if (d) return 96 + __ffs(d) - 1;
      return -1;
    }
    __device__ __forceinline__ uint32_t word(int k) const { return k == 0 ? a : k == 1 ? b : k == 2 ? c : d; }
    __device__ __forceinline__ bool test(int i) const { return (word(i >> 5) >> (i & 31)) & 1u; }
    __device__ __forceinline__ void set(int i) {
      const uint32_t bit = 1u << (i & 31);
      const int k = i >> 5;
      a |= (k == 0) ? bit : 0u;
      b |= (k == 1) ? bit : 0u;
      c |= (k == 2) ? bit : 0u;
      d |= (k == 3) ? bit : 0u;
    }
    __device__ __forceinline__ void clear(int i) {
      const uint32_t bit = 1u << (i & 31);
      const int k = i >> 5;
      a &= (k == 0) ? ~bit : ~0u;
      b &= (k == 1) ? ~bit : ~0u;
      c &= (k == 2) ? ~bit : ~0u;
      d &= (k == 3) ? ~bit : ~0u;
    }
  };

  // first zero of row r outside `cov`, or -1 (lane-private row, or the same row in every lane)
  static __device__ __forceinline__ int first_zero_outside_at(const uint32_t *Z, int zs, int mw, int r, const M4 &cov) {
    const uint32_t *zr = Z + (size_t)r * zs;
    M4 v = {0u, 0u, 0u, 0u};
    v.a = zr[0] & ~cov.a;
    if (mw > 1) v.b = zr[1] & ~cov.b;
    if (mw > 2) v.c = zr[2] & ~cov.c;
    if (mw > 3) v.d = zr[3] & ~cov.d;
    return v.first();
  }

  // uncovered rows (of the 32-row group `grp`, one per lane) for which `pred` holds, as a ballot
  template <class PRED>
  static __device__ __forceinline__ uint32_t rows_where_n(int n, int grp, uint32_t rowcov_word, PRED pred) {
    const int r = grp * 32 + lane_id();
    const bool ok = (r < n) && !((rowcov_word >> lane_id()) & 1u) && pred(r);
    return __ballot_sync(0xffffffffu, ok);
  }

  // order-preserving integer image of a float: a < b  <=>  ordered(a) < ordered(b), -0.0f < +0.0f,
  // NaN (positive payload) sorts last
  static __device__ __forceinline__ uint32_t ordered(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  }
  static __device__ __forceinline__ float unordered(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
  }

  // true when solve() will take the shared-memory path below (the producer of the cost matrix may then fold
  // step 1 into its own pass, sort_kernel.cuh, and call solve(true))
  __device__ __forceinline__ bool block_path() const { return rowwise && m <= 128; }

  // ---- warp 0 of the block solver: step 2 on first entry, then steps 3-5 until all rows are starred
  // (returns 0), the costs must shift (6: step 6, after which the caller comes back with the covers,
  // stars and primes untouched) or the iteration budget ran out (9).
  struct Drive {
    M4 starcols, colcov, rowcov;
    int stars, budget;
    bool fresh;  // first entry: step 2 still to do
    __device__ __forceinline__ void init(int n, int m) {
      starcols = M4{0u, 0u, 0u, 0u};
      colcov = starcols;
      rowcov = starcols;
      stars = 0;
      budget = 4 * n * n + 64 * (n + m) + 1024;
      fresh = true;
    }
  };

  __device__ __forceinline__ int drive_block(Drive &d, const int n, const int m, const int mw, const int zs,
                                             const MunkresGlobal &g) {
    auto first_zero_outside = [&](int r, const M4 &cov) { return first_zero_outside_at(g.Z, zs, mw, r, cov); };
    auto rows_where = [&](int grp, uint32_t rowcov_word, auto pred) { return rows_where_n(n, grp, rowcov_word, pred); };
    const int lane = lane_id();
    const int nwr = munkres_words(n);
    int act = 0;
    const unsigned lt = (1u << lane) - 1u;
    if (d.fresh) {
      // ---- step 2: greedy d.stars in row-major order; the longest prefix of pending rows whose
      // proposals are pairwise distinct is exactly what the sequential loop would star
      for (int rb = 0; rb < n; rb += 32) {
        const int r = rb + lane;
        bool pending = r < n;
        for (;;) {
          int cand = -1;
          if (pending) {
            cand = first_zero_outside(r, d.starcols);
            if (cand < 0) pending = false;  // every zero of the row is taken: no star
          }
          if (!__ballot_sync(0xffffffffu, pending)) break;
          const unsigned peers = __match_any_sync(0xffffffffu, pending ? cand : (-2 - lane));
          const unsigned clash = __ballot_sync(0xffffffffu, pending && (peers & lt) != 0u);
          const int first_clash = clash ? (__ffs(clash) - 1) : 32;
          const bool commit = pending && lane < first_clash;
          M4 add = {0u, 0u, 0u, 0u};
          if (commit) {
            g.row_star[r] = cand;
            g.col_star[cand] = r;
            add.set(cand);
            pending = false;
          }
          d.starcols.a |= __reduce_or_sync(0xffffffffu, add.a);
          if (mw > 1) d.starcols.b |= __reduce_or_sync(0xffffffffu, add.b);
          if (mw > 2) d.starcols.c |= __reduce_or_sync(0xffffffffu, add.c);
          if (mw > 3) d.starcols.d |= __reduce_or_sync(0xffffffffu, add.d);
          d.stars += __popc(__ballot_sync(0xffffffffu, commit));
        }
      }
      __syncwarp();
      tick(4);
      d.fresh = false;
      d.colcov = d.starcols;
    }
    // (otherwise: back from a cost shift — covers, d.stars and primes are untouched, step 6 -> step 4)
    // ---- steps 3-5
    for (;;) {
      if (d.stars >= n) { act = 0; break; }
      // rows that own an uncovered zero, from scratch
      M4 rowhas = {0u, 0u, 0u, 0u};
      auto open_zero = [&](int r) { return first_zero_outside(r, d.colcov) >= 0; };
      rowhas.a = rows_where(0, d.rowcov.a, open_zero);
      if (nwr > 1) rowhas.b = rows_where(1, d.rowcov.b, open_zero);
      if (nwr > 2) rowhas.c = rows_where(2, d.rowcov.c, open_zero);
      if (nwr > 3) rowhas.d = rows_where(3, d.rowcov.d, open_zero);
      bool augmented = false;
      for (;;) {
        // step 4: first uncovered zero in row-major order
        if (--d.budget < 0) { act = 9; break; }
        if (TIMERS && threadIdx.x == 0) ph[12]++;
        const int fr = rowhas.first();
        if (fr < 0) { act = 6; break; }
        const int sc = g.row_star[fr];
        const int fc = first_zero_outside(fr, d.colcov);
        if (sc < 0) {
          // step 5: flip d.stars along the alternating path that starts at the primed zero (fr, fc)
          int endc = -1;
          __syncwarp();  // every lane has read row_star[fr] before lane 0 rewrites the stars
          if (lane == 0) {
            int r = fr, c = fc;
            for (int hops = 0;; hops++) {
              const int rs = g.col_star[c];
              g.row_star[r] = c;
              g.col_star[c] = r;
              if (rs < 0) { endc = c; break; }
              r = rs;
              c = g.row_prime[r];
              if (c < 0 || hops > n + m) { endc = -2; break; }  // cannot happen in a valid state
            }
          }
          endc = __shfl_sync(0xffffffffu, endc, 0);
          if (endc < 0) { act = 9; break; }
          d.starcols.set(endc);
          d.stars++;
          augmented = true;
          break;
        }
        // the row has a star: prime the zero, cover the row, uncover the star's column
        if (lane == 0) g.row_prime[fr] = fc;
        d.rowcov.set(fr);
        d.colcov.clear(sc);
        rowhas.clear(fr);
        // uncovered rows with a zero in the newly uncovered column now own an uncovered zero
        const int kw = sc >> 5;
        const uint32_t bit = 1u << (sc & 31);
        auto zero_at_sc = [&](int r) { return (g.Z[(size_t)r * zs + kw] & bit) != 0u; };
        rowhas.a |= rows_where(0, d.rowcov.a, zero_at_sc);
        if (nwr > 1) rowhas.b |= rows_where(1, d.rowcov.b, zero_at_sc);
        if (nwr > 2) rowhas.c |= rows_where(2, d.rowcov.c, zero_at_sc);
        if (nwr > 3) rowhas.d |= rows_where(3, d.rowcov.d, zero_at_sc);
      }
      if (!augmented) break;  // act = 6 or 9
      // step 3: cover the starred columns, uncover all rows (stale primes are never read)
      __syncwarp();
      d.colcov = d.starcols;
      d.rowcov = M4{0u, 0u, 0u, 0u};
    }
    return act;
  }

  // Memory-resident matrix (g.C in shared memory, or in the slab when it is too big for it).
  // Inlined: measured 1.7 ms per C3 step faster than a real call (which turns every access to the
  // shared-memory arrays into a generic load and keeps the solver object in local memory).
  __device__ W2T_SOLVE_BLOCK_INLINE int solve_block(bool reduced = false) {
    const int n = this->n, m = this->m, mw = this->mw, zs = this->zs, ldc = this->ldc;
    const MunkresGlobal g = this->g;
    MunkresShared *const s = this->s;
    const int lane = lane_id(), warp = warp_id();
    const int nwr = munkres_words(n);
    // ---- step 1 (CTA, one thread per row): row minimum, subtract, zero bit words
    for (int c = threadIdx.x; c < m; c += BLOCK) g.col_star[c] = -1;
    for (int r = threadIdx.x; r < n && !reduced; r += BLOCK) {
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
      g.row_star[r] = -1;
      g.row_prime[r] = -1;
    }
    __syncthreads();
    tick(3);

    Drive d;
    d.init(n, m);
    int act = 0;
    for (;;) {
      if (warp == 0) {
        act = drive_block(d, n, m, mw, zs, g);
        const M4 &colcov = d.colcov, &rowcov = d.rowcov;
        if (act == 6) {
          // step 6 prologue: compact index lists, so that the CTA touches only the cells that change:
          // uncovered columns -> g.ucols[0..nu), uncovered rows -> g.ucols[128..128+nur), covered
          // rows -> g.crows[0..ncr)
          const unsigned ltm = (1u << lane) - 1u;
          int nu = 0, nur = 0, ncr = 0;
          for (int k = 0; k < mw; k++) {
            const int c = k * 32 + lane;
            const bool open = (c < m) && !((colcov.word(k) >> lane) & 1u);
            const unsigned b = __ballot_sync(0xffffffffu, open);
            if (open) g.ucols[nu + __popc(b & ltm)] = c;
            nu += __popc(b);
          }
          for (int k = 0; k < nwr; k++) {
            const int r = k * 32 + lane;
            const bool cov = (rowcov.word(k) >> lane) & 1u;
            const unsigned bu = __ballot_sync(0xffffffffu, (r < n) && !cov);
            const unsigned bc = __ballot_sync(0xffffffffu, (r < n) && cov);
            if (r < n && !cov) g.ucols[128 + nur + __popc(bu & ltm)] = r;
            if (r < n && cov) g.crows[ncr + __popc(bc & ltm)] = r;
            nur += __popc(bu);
            ncr += __popc(bc);
          }
          if (lane == 0) { s->ctl[1] = nu; s->ctl[2] = ncr; s->ctl[3] = nur; }
        }
        tick(5);
        if (lane == 0) {
          s->ctl[0] = act;
          s->colcov[0] = colcov.a; s->colcov[1] = colcov.b; s->colcov[2] = colcov.c; s->colcov[3] = colcov.d;
        }
      }
      __syncthreads();
      act = s->ctl[0];
      if (act != 6) return act;
      if (TIMERS && threadIdx.x == 0) ph[11]++;
      // ---- step 6 (CTA): min over uncovered rows x uncovered columns; covered rows += min, then
      // uncovered columns -= min (float32, in that order).  Only the cells that change are visited:
      // (uncovered rows x uncovered columns) and (covered rows x all columns).
      const int nu = s->ctl[1], ncr = s->ctl[2], nur = s->ctl[3];
      const int *ucols = g.ucols, *urows = g.ucols + 128, *crows = g.crows;
      // flat cell index i -> (uncovered row i / nu, uncovered column i % nu); four independent cells
      // per thread and trip, so their shared-memory round trips overlap
      const int ncell = nur * nu;
      const unsigned inv_nu = nu > 0 ? (0xffffffffu / (unsigned)nu + 1u) : 0u;  // exact for i < 2^16
      auto cell = [&](int i, int &r, int &c) {
        const int q = (nu == 1) ? i : (int)__umulhi((unsigned)i, inv_nu);  // 2^32 / 1 does not fit
        r = urows[q];
        c = ucols[i - q * nu];
      };
      float mn = INFINITY;
      for (int base = threadIdx.x; base < ncell; base += 4 * BLOCK) {
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int i = base + j * BLOCK;
          v[j] = INFINITY;
          if (i < ncell) {
            int r, c;
            cell(i, r, c);
            v[j] = g.C[(size_t)r * ldc + c];
          }
        }
        mn = fminf(fminf(mn, v[0]), fminf(fminf(v[1], v[2]), v[3]));
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      if (lane == 0) s->red[warp] = mn;
      __syncthreads();
      mn = s->red[0];
#pragma unroll
      for (int i = 1; i < NW; i++) mn = fminf(mn, s->red[i]);
      if (mn != INFINITY) {  // nothing uncovered: the reference leaves the matrix alone
        // covered rows: every column gets +min, the uncovered ones then -min (two roundings)
        const M4 cc = {s->colcov[0], s->colcov[1], s->colcov[2], s->colcov[3]};
        for (int i = warp; i < ncr; i += NW) {
          const int r = crows[i];
          float *row = g.C + (size_t)r * ldc;
          for (int k = 0; k < mw; k++) {
            const int c = k * 32 + lane;
            bool z = false;
            if (c < m) {
              float v = row[c] + mn;
              if (!((cc.word(k) >> lane) & 1u)) v = v - mn;
              row[c] = v;
              z = (v == 0.0f);
            }
            const unsigned word = __ballot_sync(0xffffffffu, z);
            if (lane == 0) g.Z[(size_t)r * zs + k] = word;
          }
        }
        // uncovered rows: only the uncovered columns change (-min).  None of these cells was zero
        // (step 4 ran out of uncovered zeros), so zero bits are only ever set here.
        for (int base = threadIdx.x; base < ncell; base += 4 * BLOCK) {
          float v[4];
          int rr[4], cc4[4];
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const int i = base + j * BLOCK;
            rr[j] = -1;
            if (i < ncell) {
              cell(i, rr[j], cc4[j]);
              v[j] = g.C[(size_t)rr[j] * ldc + cc4[j]];
            }
          }
#pragma unroll
          for (int j = 0; j < 4; j++) {
            if (rr[j] >= 0) {
              const float w = v[j] - mn;
              g.C[(size_t)rr[j] * ldc + cc4[j]] = w;
              if (w == 0.0f) atomicOr(&g.Z[(size_t)rr[j] * zs + (cc4[j] >> 5)], 1u << (cc4[j] & 31));
            }
          }
        }
      }
      __syncthreads();
      tick(6);
    }
  }

  // Whole solve.  Precondition: g.C holds the n x m costs (n <= m).  All threads call it.
  // Returns 0, or 9 if the iteration budget ran out (malformed input such as NaN costs).
  __device__ int solve(bool reduced = false) {
    if (block_path()) return solve_block(reduced);
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

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
// columns) that is rebuilt only when the costs change (step 1 and step 6); with the column
// cover kept as a bit mask too, "first uncovered zero of a row" is an AND + find-first-set
// per word instead of a scan over floats.
//
// Work split.  Steps 3-5 are a serial state machine: warp 0 runs them (lanes own words of
// the masks, rows are scanned 32 at a time) while the other warps wait at the CTA barrier.
// Step 1 and step 6 touch the whole matrix and are done by all warps, one warp per row,
// coalesced, with the Z words produced by __ballot_sync.
#pragma once

#include "common.cuh"

namespace w2t {

constexpr int kMunkresMaxWords = 64;                      // mask words kept in shared memory
constexpr int kMunkresMaxDim = kMunkresMaxWords * 32;     // largest max(D, T)

struct MunkresShared {
  uint32_t colcov[kMunkresMaxWords];
  uint32_t rowcov[kMunkresMaxWords];
  uint32_t starcols[kMunkresMaxWords];
  float red[32];
  int ctl[4];  // 0: next action (0 done, 6 shift, 9 error)  1: stars  2: resume step  3: unused
};

struct MunkresGlobal {
  float *C;         // [n*m] cost, n <= m, row-major
  uint32_t *Z;      // [n*zs] zero bit matrix
  int *row_star;    // [n] column of the row's star or -1
  int *col_star;    // [m] row of the column's star or -1
  int *row_prime;   // [n] column of the row's prime (valid for rows primed in the current phase)
};

__host__ __device__ inline int munkres_words(int m) { return (m + 31) >> 5; }
__host__ __device__ inline int munkres_zstride(int m) { return munkres_words(m) | 1; }

template <int BLOCK>
struct Munkres {
  static constexpr int NW = BLOCK / 32;
  int n, m, mw, zs;
  MunkresGlobal g;
  MunkresShared *s;

  __device__ __forceinline__ bool row_covered(int r) const { return (s->rowcov[r >> 5] >> (r & 31)) & 1u; }

  // ---- step 1: subtract the row minimum, build Z ------------------------------------------
  __device__ void reduce_rows() {
    const int lane = lane_id();
    for (int r = warp_id(); r < n; r += NW) {
      float *row = g.C + (size_t)r * m;
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

  // ---- step 2: star zeros greedily in row-major order (warp 0) ---------------------------------
  __device__ int greedy_stars() {
    const int lane = lane_id();
    uint32_t cov0 = 0, cov1 = 0;  // words lane and lane+32 of the column cover
    int stars = 0;
    for (int r = 0; r < n; r++) {
      const uint32_t *zr = g.Z + (size_t)r * zs;
      uint32_t v0 = (lane < mw) ? (zr[lane] & ~cov0) : 0u;
      unsigned b = __ballot_sync(0xffffffffu, v0 != 0u);
      int wsel = -1;
      uint32_t vsel = 0;
      if (b) {
        const int src = __ffs(b) - 1;
        vsel = __shfl_sync(0xffffffffu, v0, src);
        wsel = src;
        if (lane == src) cov0 |= (vsel & (0u - vsel));
      } else if (mw > 32) {
        uint32_t v1 = (lane + 32 < mw) ? (zr[lane + 32] & ~cov1) : 0u;
        b = __ballot_sync(0xffffffffu, v1 != 0u);
        if (b) {
          const int src = __ffs(b) - 1;
          vsel = __shfl_sync(0xffffffffu, v1, src);
          wsel = src + 32;
          if (lane == src) cov1 |= (vsel & (0u - vsel));
        }
      }
      if (wsel >= 0) {
        const int c = wsel * 32 + (__ffs(vsel) - 1);
        if (lane == 0) {
          g.row_star[r] = c;
          g.col_star[c] = r;
        }
        stars++;
      }
    }
    if (lane < mw) s->starcols[lane] = cov0;
    if (lane + 32 < mw) s->starcols[lane + 32] = cov1;
    __syncwarp();
    return stars;
  }

  // first uncovered zero of row r: column index or -1 (warp 0, all lanes)
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

  // ---- steps 3, 4, 5 (warp 0): returns 0 when n stars exist, 6 when the costs must shift -----
  __device__ int drive(int step, int &stars, int &budget) {
    const int lane = lane_id();
    for (;;) {
      if (step == 3) {
        for (int w = lane; w < kMunkresMaxWords; w += 32) {
          s->rowcov[w] = 0u;
          s->colcov[w] = (w < mw) ? s->starcols[w] : 0u;
        }
        __syncwarp();
        if (stars >= n) return 0;
        step = 4;
      }
      // step 4: first uncovered zero in row-major order
      if (--budget < 0) return 9;
      int fr = -1;
      for (int rb = 0; rb < n; rb += 32) {
        const int r = rb + lane;
        bool any = false;
        if (r < n && !row_covered(r)) {
          const uint32_t *zr = g.Z + (size_t)r * zs;
          for (int w = 0; w < mw; w++)
            if (zr[w] & ~s->colcov[w]) { any = true; break; }
        }
        const unsigned b = __ballot_sync(0xffffffffu, any);
        if (b) { fr = rb + __ffs(b) - 1; break; }
      }
      if (fr < 0) return 6;
      const int fc = first_open_zero(fr);
      const int sc = g.row_star[fr];
      if (sc >= 0) {
        // the row has a star: prime the zero, cover the row, uncover the star's column
        if (lane == 0) {
          g.row_prime[fr] = fc;
          s->rowcov[fr >> 5] |= (1u << (fr & 31));
          s->colcov[sc >> 5] &= ~(1u << (sc & 31));
        }
        __syncwarp();
        continue;
      }
      // step 5: augment along the alternating path that starts at the primed zero (fr, fc)
      if (lane == 0) {
        int r = fr, c = fc;
        for (;;) {
          const int rs = g.col_star[c];
          g.row_star[r] = c;
          g.col_star[c] = r;
          if (rs < 0) {
            s->starcols[c >> 5] |= (1u << (c & 31));
            break;
          }
          r = rs;
          c = g.row_prime[r];
        }
      }
      __syncwarp();
      stars++;
      step = 3;
    }
  }

  // ---- step 6 (all threads): shift by the smallest uncovered value, rebuild Z --------------
  __device__ void shift_by_min() {
    const int lane = lane_id(), warp = warp_id();
    float mn = INFINITY;
    for (int r = warp; r < n; r += NW) {
      if (row_covered(r)) continue;
      const float *row = g.C + (size_t)r * m;
      for (int c = lane; c < m; c += 32)
        if (!((s->colcov[c >> 5] >> (c & 31)) & 1u)) mn = fminf(mn, row[c]);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    if (lane == 0) s->red[warp] = mn;
    __syncthreads();
    mn = s->red[0];
#pragma unroll
    for (int i = 1; i < NW; i++) mn = fminf(mn, s->red[i]);
    if (mn == INFINITY) return;  // nothing uncovered: the reference leaves the matrix alone
    for (int r = warp; r < n; r += NW) {
      const bool rc = row_covered(r);
      float *row = g.C + (size_t)r * m;
      for (int k = 0; k < mw; k++) {
        const int c = k * 32 + lane;
        bool z = false;
        if (c < m) {
          const bool cu = !((s->colcov[k] >> lane) & 1u);
          float v = row[c];
          if (rc) v = v + mn;
          if (cu) v = v - mn;
          if (rc || cu) row[c] = v;
          z = (v == 0.0f);
        }
        const unsigned word = __ballot_sync(0xffffffffu, z);
        if (lane == 0) g.Z[(size_t)r * zs + k] = word;
      }
    }
  }

  // Whole solve.  Precondition: g.C holds the n x m costs (n <= m).  All threads call it.
  // Returns 0, or 9 if the iteration budget ran out (malformed input such as NaN costs).
  __device__ int solve() {
    for (int i = threadIdx.x; i < n; i += BLOCK) { g.row_star[i] = -1; g.row_prime[i] = -1; }
    for (int i = threadIdx.x; i < m; i += BLOCK) g.col_star[i] = -1;
    reduce_rows();
    __syncthreads();
    int stars = 0, step = 3;
    int budget = 4 * n * n + 64 * (n + m) + 1024;
    if (warp_id() == 0) stars = greedy_stars();
    for (;;) {
      if (warp_id() == 0) {
        const int act = drive(step, stars, budget);
        if (lane_id() == 0) s->ctl[0] = act;
        step = 4;
      }
      __syncthreads();
      const int act = s->ctl[0];
      if (act != 6) return act;
      shift_by_min();
      __syncthreads();
    }
  }
};

}  // namespace w2t

/*
 * w2t_oracle.c — CPU restatement (plain C) of the reference hot path.
 * TEST INFRASTRUCTURE; see w2t_oracle.h.  Build: oracle/Makefile
 * (gcc -O2 -ffp-contract=off -mfma).
 *
 * Each function cites the reference lines it follows (paths relative to
 * /root/reference).  The Hungarian solver and the Kalman filter restate
 * third-party code the reference imports (scikit-learn 0.22.2
 * sklearn/utils/linear_assignment_.py, filterpy kalman/kalman_filter.py);
 * see oracle/munkres.py and oracle/kalman.py for the NumPy restatements this
 * file is checked against.
 */
#include "w2t_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* Kalman box filter                                                   */
/* ------------------------------------------------------------------ */

/* sort.py:111-115 -> Q = diag(2,2,1,25,4,4,5); sort.py:127 -> R = diag(1,1,10,10) */
static const double Q_DIAG[7] = {2., 2., 1., 25., 4., 4., 5.};
static const double R_DIAG[4] = {1., 1., 10., 10.};

void w2t_oracle_bbox_to_z(const float d[4], float z[4]) {
  float w = d[2] - d[0];
  float h = d[3] - d[1];
  z[0] = d[0] + w / 2.0f;
  z[1] = d[1] + h / 2.0f;
  z[2] = w * h;
  z[3] = w / h;
}

/* sort.py:50-62 under either promotion regime (w2t_types.h, W2T_PROMOTION_*), widened to the float64 the
 * filter works in.  Legacy (NumPy 1.x): w, h and s = w*h are float32 scalars, but float32-scalar (op)
 * python-float is float64, so x = bbox[0] + w/2., y and r = w/float(h) are float64 computations. */
void w2t_oracle_bbox_to_z_d(const float d[4], int promotion, double z[4]) {
  float w = d[2] - d[0];
  float h = d[3] - d[1];
  if (promotion == W2T_PROMOTION_NEP50) {
    z[0] = (double)(float)(d[0] + w / 2.0f);
    z[1] = (double)(float)(d[1] + h / 2.0f);
    z[3] = (double)(float)(w / h);
  } else {
    z[0] = (double)d[0] + (double)w / 2.0;
    z[1] = (double)d[1] + (double)h / 2.0;
    z[3] = (double)w / (double)h;
  }
  z[2] = (double)(float)(w * h);
}

void w2t_oracle_x_to_bbox(const double x[7], double b[4]) {
  double w = sqrt(x[2] * x[3]);
  double h = x[2] / w;
  b[0] = x[0] - w / 2.;
  b[1] = x[1] - h / 2.;
  b[2] = x[0] + w / 2.;
  b[3] = x[1] + h / 2.;
}

void w2t_oracle_kf_init(const float det[4], double x[7], double P[49], int promotion) {
  double z[4];
  w2t_oracle_bbox_to_z_d(det, promotion, z);
  for (int i = 0; i < 4; i++) x[i] = z[i];
  x[4] = x[5] = x[6] = 0.;
  memset(P, 0, 49 * sizeof(double));
  for (int i = 0; i < 4; i++) P[i * 7 + i] = 10.;
  for (int i = 4; i < 7; i++) P[i * 7 + i] = 10000.;
}

void w2t_oracle_kf_predict(double x[7], double P[49]) {
  /* sort.py:170-171 */
  if (x[6] + x[2] <= 0) x[6] *= 0.0;
  /* x = F x : rows 0..2 pick up their velocity */
  x[0] = x[0] + x[4];
  x[1] = x[1] + x[5];
  x[2] = x[2] + x[6];
  /* A = F P ; B = A F' ; P = B + Q   (two-stage association) */
  double A[49];
  for (int i = 0; i < 7; i++)
    for (int j = 0; j < 7; j++) A[i * 7 + j] = (i < 3) ? P[i * 7 + j] + P[(i + 4) * 7 + j] : P[i * 7 + j];
  for (int i = 0; i < 7; i++)
    for (int j = 0; j < 7; j++) {
      double b = (j < 3) ? A[i * 7 + j] + A[i * 7 + j + 4] : A[i * 7 + j];
      P[i * 7 + j] = b + ((i == j) ? Q_DIAG[i] : 0.0);
    }
}

void w2t_oracle_inv4(const double S[16], double out[16]) {
  /* numpy.linalg.inv -> LAPACK dgesv(S, I) as OpenBLAS executes it for a 4x4:
   * left-looking getf2 (dot products accumulated by FMA from 0, then subtracted),
   * pivot rows scaled by the reciprocal, then two right-looking triangular solves
   * with FMA updates and a reciprocal diagonal.  Column-major like LAPACK. */
  double a[16], b[16];
  int piv[4];
#define A_(i, j) a[(i) + 4 * (j)]
#define B_(i, j) b[(i) + 4 * (j)]
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) A_(i, j) = S[i * 4 + j];
  for (int j = 0; j < 4; j++) {
    double *col = &A_(0, j);
    for (int i = 0; i < j; i++) {
      int ip = piv[i];
      if (ip != i) { double t = col[i]; col[i] = col[ip]; col[ip] = t; }
    }
    for (int i = 1; i < j; i++) {
      double t = 0.;
      for (int k = 0; k < i; k++) t = fma(A_(i, k), col[k], t);
      col[i] = col[i] - t;
    }
    if (j > 0)
      for (int r = j; r < 4; r++) {
        double t = 0.;
        for (int k = 0; k < j; k++) t = fma(A_(r, k), col[k], t);
        col[r] = col[r] - t;
      }
    int jp = j;
    double mx = fabs(col[j]);
    for (int r = j + 1; r < 4; r++)
      if (fabs(col[r]) > mx) { mx = fabs(col[r]); jp = r; }
    piv[j] = jp;
    if (jp != j)
      for (int c = 0; c <= j; c++) { double t = A_(j, c); A_(j, c) = A_(jp, c); A_(jp, c) = t; }
    double rp = 1.0 / col[j];
    for (int r = j + 1; r < 4; r++) col[r] = col[r] * rp;
  }
  memset(b, 0, sizeof b);
  for (int i = 0; i < 4; i++) B_(i, i) = 1.0;
  for (int i = 0; i < 4; i++) {
    int ip = piv[i];
    if (ip != i)
      for (int c = 0; c < 4; c++) { double t = B_(i, c); B_(i, c) = B_(ip, c); B_(ip, c) = t; }
  }
  for (int c = 0; c < 4; c++) {
    double *x = &B_(0, c);
    for (int i = 0; i < 4; i++) {
      double bb = x[i];
      for (int k = i + 1; k < 4; k++) x[k] = fma(-bb, A_(k, i), x[k]);
    }
    for (int i = 3; i >= 0; i--) {
      double bb = x[i] * (1.0 / A_(i, i));
      x[i] = bb;
      for (int k = 0; k < i; k++) x[k] = fma(-bb, A_(k, i), x[k]);
    }
  }
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) out[i * 4 + j] = B_(i, j);
#undef A_
#undef B_
}

void w2t_oracle_kf_update(double x[7], double P[49], const float det[4], int promotion) {
  double zf[4];
  double y[4], S[16], SI[16], K[28], A[28], M[49], N[49];
  w2t_oracle_bbox_to_z_d(det, promotion, zf);
  /* y = z - Hx */
  for (int i = 0; i < 4; i++) y[i] = zf[i] - x[i];
  /* S = H P H' + R  (H selects the leading 4x4 block) */
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) S[i * 4 + j] = P[i * 7 + j] + ((i == j) ? R_DIAG[i] : 0.0);
  w2t_oracle_inv4(S, SI);
  /* K = (P H') SI : dgemm, k-sequential FMA */
  for (int i = 0; i < 7; i++)
    for (int j = 0; j < 4; j++) {
      double acc = 0.;
      for (int k = 0; k < 4; k++) acc = fma(P[i * 7 + k], SI[k * 4 + j], acc);
      K[i * 4 + j] = acc;
    }
  /* x = x + K y : dgemv, (p0+p2)+(p1+p3) with separately rounded products */
  for (int i = 0; i < 7; i++) {
    double p0 = K[i * 4 + 0] * y[0], p1 = K[i * 4 + 1] * y[1];
    double p2 = K[i * 4 + 2] * y[2], p3 = K[i * 4 + 3] * y[3];
    x[i] = x[i] + ((p0 + p2) + (p1 + p3));
  }
  /* A = (I - K H)[:, :4] ; the remaining columns of I - KH are those of the identity */
  for (int i = 0; i < 7; i++)
    for (int j = 0; j < 4; j++) A[i * 4 + j] = ((i == j) ? 1.0 : 0.0) - K[i * 4 + j];
  /* M = (I-KH) P */
  for (int i = 0; i < 7; i++)
    for (int j = 0; j < 7; j++) {
      double acc = 0.;
      for (int k = 0; k < 4; k++) acc = fma(A[i * 4 + k], P[k * 7 + j], acc);
      if (i >= 4) acc = acc + P[i * 7 + j];
      M[i * 7 + j] = acc;
    }
  /* N = M (I-KH)' */
  for (int i = 0; i < 7; i++)
    for (int j = 0; j < 7; j++) {
      double acc = 0.;
      for (int k = 0; k < 4; k++) acc = fma(M[i * 7 + k], A[j * 4 + k], acc);
      if (j >= 4) acc = acc + M[i * 7 + j];
      N[i * 7 + j] = acc;
    }
  /* P = N + (K R) K' */
  for (int i = 0; i < 7; i++)
    for (int j = 0; j < 7; j++) {
      double acc = 0.;
      for (int k = 0; k < 4; k++) acc = fma(K[i * 4 + k] * R_DIAG[k], K[j * 4 + k], acc);
      P[i * 7 + j] = N[i * 7 + j] + acc;
    }
}

/* ------------------------------------------------------------------ */
/* IoU                                                                 */
/* ------------------------------------------------------------------ */

static inline float iou_pair(const float *d, const double *t) {
  /* sort.py:38-46 as numba types it: detection area in float32, the rest float64 */
  double xx1 = (double)d[0] > t[0] ? (double)d[0] : t[0];
  double yy1 = (double)d[1] > t[1] ? (double)d[1] : t[1];
  double xx2 = (double)d[2] < t[2] ? (double)d[2] : t[2];
  double yy2 = (double)d[3] < t[3] ? (double)d[3] : t[3];
  double w = xx2 - xx1; if (!(w > 0.)) w = 0.;
  double h = yy2 - yy1; if (!(h > 0.)) h = 0.;
  double wh = w * h;
  float ad = (d[2] - d[0]) * (d[3] - d[1]);
  double at = (t[2] - t[0]) * (t[3] - t[1]);
  return (float)(wh / (((double)ad + at) - wh));
}

void w2t_oracle_iou_matrix(const float *dets, int D, const double *trks, int T, float *out) {
  for (int d = 0; d < D; d++)
    for (int t = 0; t < T; t++) out[(size_t)d * T + t] = iou_pair(dets + 4 * d, trks + 4 * t);
}

/* ------------------------------------------------------------------ */
/* Munkres (scikit-learn 0.22.2 _hungarian)                            */
/* ------------------------------------------------------------------ */

typedef struct {
  int n, m;
  float *C;
  unsigned char *mark; /* 0 none, 1 star, 2 prime */
  unsigned char *row_unc, *col_unc;
  int z0r, z0c;
  int *path; /* (n+m) x 2 */
} hung_t;

/* debug counters (scripts/munkres_stats.py): per size bucket of m — calls, step-4 iterations,
 * augmentations, step-6 rounds, sum of n*m, stars after the greedy step */
static long long g_stats[8][6];
static int g_bucket = 0;
void w2t_oracle_munkres_stats(long long *out, int reset) {
  if (out) memcpy(out, g_stats, sizeof(g_stats));
  if (reset) memset(g_stats, 0, sizeof(g_stats));
}

static int hung_step3(hung_t *s) {
  int stars = 0;
  for (int r = 0; r < s->n; r++)
    for (int c = 0; c < s->m; c++)
      if (s->mark[(size_t)r * s->m + c] == 1) { s->col_unc[c] = 0; stars++; }
  return stars < s->n ? 4 : 0;
}

static int hung_step4(hung_t *s) {
  const int n = s->n, m = s->m;
  for (;;) {
    int fr = -1, fc = -1;
    for (int r = 0; r < n && fr < 0; r++) {
      if (!s->row_unc[r]) continue;
      const float *row = s->C + (size_t)r * m;
      for (int c = 0; c < m; c++)
        if (row[c] == 0.0f && s->col_unc[c]) { fr = r; fc = c; break; }
    }
    if (fr < 0) return 6;
    g_stats[g_bucket][1]++;
    s->mark[(size_t)fr * m + fc] = 2;
    int sc = -1;
    for (int c = 0; c < m; c++)
      if (s->mark[(size_t)fr * m + c] == 1) { sc = c; break; }
    if (sc < 0) { s->z0r = fr; s->z0c = fc; return 5; }
    s->row_unc[fr] = 0;
    s->col_unc[sc] = 1;
  }
}

static int hung_step5(hung_t *s) {
  const int n = s->n, m = s->m;
  int count = 0;
  int *path = s->path;
  path[0] = s->z0r; path[1] = s->z0c;
  for (;;) {
    int c = path[2 * count + 1], row = -1;
    for (int r = 0; r < n; r++)
      if (s->mark[(size_t)r * m + c] == 1) { row = r; break; }
    if (row < 0) break;
    count++;
    path[2 * count] = row; path[2 * count + 1] = c;
    int col = -1;
    for (int cc = 0; cc < m; cc++)
      if (s->mark[(size_t)row * m + cc] == 2) { col = cc; break; }
    count++;
    path[2 * count] = row; path[2 * count + 1] = col;
  }
  for (int i = 0; i <= count; i++) {
    unsigned char *cell = &s->mark[(size_t)path[2 * i] * m + path[2 * i + 1]];
    *cell = (*cell == 1) ? 0 : 1;
  }
  memset(s->row_unc, 1, n);
  memset(s->col_unc, 1, m);
  for (size_t i = 0; i < (size_t)n * m; i++)
    if (s->mark[i] == 2) s->mark[i] = 0;
  g_stats[g_bucket][2]++;
  return 3;
}

static int hung_step6(hung_t *s) {
  const int n = s->n, m = s->m;
  int any_r = 0, any_c = 0;
  for (int r = 0; r < n; r++) any_r |= s->row_unc[r];
  for (int c = 0; c < m; c++) any_c |= s->col_unc[c];
  g_stats[g_bucket][3]++;
  if (any_r && any_c) {
    float minval = INFINITY;
    for (int r = 0; r < n; r++) {
      if (!s->row_unc[r]) continue;
      const float *row = s->C + (size_t)r * m;
      for (int c = 0; c < m; c++)
        if (s->col_unc[c] && row[c] < minval) minval = row[c];
    }
    for (int r = 0; r < n; r++)
      if (!s->row_unc[r]) {
        float *row = s->C + (size_t)r * m;
        for (int c = 0; c < m; c++) row[c] = row[c] + minval;
      }
    for (int r = 0; r < n; r++) {
      float *row = s->C + (size_t)r * m;
      for (int c = 0; c < m; c++)
        if (s->col_unc[c]) row[c] = row[c] - minval;
    }
  }
  return 4;
}

int w2t_oracle_linear_assignment(const float *cost, int D, int T, int32_t *pairs) {
  if (D <= 0 || T <= 0) return 0;
  hung_t s;
  const int flipped = T < D;
  s.n = flipped ? T : D;
  s.m = flipped ? D : T;
  const int n = s.n, m = s.m;
  s.C = (float *)malloc(sizeof(float) * (size_t)n * m);
  s.mark = (unsigned char *)calloc((size_t)n * m, 1);
  s.row_unc = (unsigned char *)malloc(n);
  s.col_unc = (unsigned char *)malloc(m);
  s.path = (int *)malloc(sizeof(int) * 2 * (size_t)(n + m));
  if (!s.C || !s.mark || !s.row_unc || !s.col_unc || !s.path) return -1;
  for (int r = 0; r < n; r++)
    for (int c = 0; c < m; c++)
      s.C[(size_t)r * m + c] = flipped ? cost[(size_t)c * T + r] : cost[(size_t)r * T + c];
  memset(s.row_unc, 1, n);
  memset(s.col_unc, 1, m);
  /* step 1 (+2): row reduction, greedy stars over zeros in row-major order */
  for (int r = 0; r < n; r++) {
    float *row = s.C + (size_t)r * m;
    float mn = row[0];
    for (int c = 1; c < m; c++)
      if (row[c] < mn) mn = row[c];
    for (int c = 0; c < m; c++) row[c] = row[c] - mn;
  }
  for (int r = 0; r < n; r++)
    for (int c = 0; c < m; c++)
      if (s.C[(size_t)r * m + c] == 0.0f && s.col_unc[c] && s.row_unc[r]) {
        s.mark[(size_t)r * m + c] = 1;
        s.col_unc[c] = 0;
        s.row_unc[r] = 0;
      }
  memset(s.row_unc, 1, n);
  memset(s.col_unc, 1, m);
  g_bucket = m <= 16 ? 0 : m <= 32 ? 1 : m <= 64 ? 2 : m <= 128 ? 3 : m <= 256 ? 4 : m <= 512 ? 5 : 6;
  g_stats[g_bucket][0]++;
  g_stats[g_bucket][4] += (long long)n * m;
  for (size_t i = 0; i < (size_t)n * m; i++) g_stats[g_bucket][5] += (s.mark[i] == 1);
  int step = 3;
  while (step) {
    switch (step) {
      case 3: step = hung_step3(&s); break;
      case 4: step = hung_step4(&s); break;
      case 5: step = hung_step5(&s); break;
      case 6: step = hung_step6(&s); break;
    }
  }
  /* starred cells, columns swapped back, sorted by (row, col) */
  int k = 0;
  if (!flipped) {
    for (int r = 0; r < n; r++)
      for (int c = 0; c < m; c++)
        if (s.mark[(size_t)r * m + c] == 1) { pairs[2 * k] = r; pairs[2 * k + 1] = c; k++; }
  } else {
    for (int c = 0; c < m; c++)
      for (int r = 0; r < n; r++)
        if (s.mark[(size_t)r * m + c] == 1) { pairs[2 * k] = c; pairs[2 * k + 1] = r; k++; }
  }
  free(s.C); free(s.mark); free(s.row_unc); free(s.col_unc); free(s.path);
  return k;
}

/* ------------------------------------------------------------------ */
/* association (sort.py:193-230)                                       */
/* ------------------------------------------------------------------ */

int w2t_oracle_associate(const float *dets, int D, const double *trks, int T, double iou_threshold, int promotion,
                         int32_t *matched_det_of_trk, int32_t *new_order) {
  int n_new = 0;
  for (int t = 0; t < T; t++) matched_det_of_trk[t] = -1;
  if (T == 0) {
    for (int d = 0; d < D; d++) new_order[n_new++] = d;
    return n_new;
  }
  if (D == 0) return 0;
  float *M = (float *)malloc(sizeof(float) * (size_t)D * T);
  float *cost = (float *)malloc(sizeof(float) * (size_t)D * T);
  int kmax = D < T ? D : T;
  int32_t *pairs = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)kmax);
  unsigned char *assigned = (unsigned char *)calloc(D, 1);
  w2t_oracle_iou_matrix(dets, D, trks, T, M);
  for (size_t i = 0; i < (size_t)D * T; i++) cost[i] = -M[i];
  int k = w2t_oracle_linear_assignment(cost, D, T, pairs);
  for (int i = 0; i < k; i++) assigned[pairs[2 * i]] = 1;
  for (int d = 0; d < D; d++)
    if (!assigned[d]) new_order[n_new++] = d;
  /* sort.py:220: NumPy 1.x compares the float32 entry with the python float in float64; under NEP 50 the
   * python float adopts float32 */
  const float thr = (float)iou_threshold;
  for (int i = 0; i < k; i++) {
    int d = pairs[2 * i], t = pairs[2 * i + 1];
    const float o = M[(size_t)d * T + t];
    const int rejected = (promotion == W2T_PROMOTION_NEP50) ? (o < thr) : ((double)o < iou_threshold);
    if (rejected) new_order[n_new++] = d;
    else matched_det_of_trk[t] = d;
  }
  free(M); free(cost); free(pairs); free(assigned);
  return n_new;
}

/* ------------------------------------------------------------------ */
/* SORT over packed streams                                            */
/* ------------------------------------------------------------------ */

typedef struct {
  double x[7];
  double P[49];
  int tsu, hit_streak, birth_g, birth_k;
} trk_t;

static inline double clipd(double v, double lo, double hi) {
  /* numpy.clip = minimum(maximum(v, lo), hi); NaN propagates */
  if (v < lo) v = lo;
  if (v > hi) v = hi;
  return v;
}

static int track_substream(const w2t_sort_problem_t *p, w2t_sort_result_t *r, int s, int c) {
  const int NC = p->n_classes;
  const int q = s * NC + c;
  const int img0 = p->stream_img_offsets[s], img1 = p->stream_img_offsets[s + 1];
  const double W = p->cam_wh[2 * s], H = p->cam_wh[2 * s + 1];
  int cap = 64, T = 0, frame_count = 0, started = 0, status = W2T_OK;
  trk_t *trk = (trk_t *)malloc(sizeof(trk_t) * cap);
  double *boxes = NULL;
  int32_t *match = NULL, *new_order = NULL;
  int boxes_cap = 0, new_cap = 0;
  r->first_img[q] = -1;
  for (int img = img0; img < img1; img++) {
    const int g = img * NC + c;
    r->out_count[g] = 0;
    r->created[g] = 0;
    if (p->img_exists && !p->img_exists[img]) continue;
    const int D = p->det_count[g];
    if (!started) {
      if (D == 0) continue; /* no Sort object for this category yet, tracker_sort.py:32-33 */
      started = 1;
      r->first_img[q] = img - img0;
    }
    const int base = p->det_start[g];
    const float *dets = p->det_box + 4 * (size_t)base;
    frame_count++;
    /* sort.py:255-265 predict, drop NaN trackers */
    if (T > boxes_cap) {
      boxes_cap = T + 64;
      boxes = (double *)realloc(boxes, sizeof(double) * 4 * boxes_cap);
      match = (int32_t *)realloc(match, sizeof(int32_t) * boxes_cap);
    }
    int Tk = 0;
    for (int t = 0; t < T; t++) {
      trk_t *k = &trk[t];
      w2t_oracle_kf_predict(k->x, k->P);
      if (k->tsu > 0) k->hit_streak = 0;
      k->tsu += 1;
      double b[4];
      w2t_oracle_x_to_bbox(k->x, b);
      if (isnan(b[0]) || isnan(b[1]) || isnan(b[2]) || isnan(b[3])) continue;
      if (isinf(b[0]) || isinf(b[1]) || isinf(b[2]) || isinf(b[3])) status = W2T_ERR_NONFINITE;
      memcpy(boxes + 4 * Tk, b, sizeof b);
      if (Tk != t) trk[Tk] = *k;
      Tk++;
    }
    T = Tk;
    if (D > new_cap) {
      new_cap = D + 64;
      new_order = (int32_t *)realloc(new_order, sizeof(int32_t) * new_cap);
    }
    int n_new = w2t_oracle_associate(dets, D, boxes, T, p->iou_thr[c], p->promotion, match, new_order);
    /* sort.py:270-273 */
    for (int t = 0; t < T; t++)
      if (match[t] >= 0) {
        trk_t *k = &trk[t];
        k->tsu = 0;
        k->hit_streak += 1;
        w2t_oracle_kf_update(k->x, k->P, dets + 4 * match[t], p->promotion);
      }
    /* sort.py:276-278 */
    if (T + n_new > cap) {
      cap = (T + n_new) * 2;
      trk = (trk_t *)realloc(trk, sizeof(trk_t) * cap);
    }
    for (int i = 0; i < n_new; i++) {
      trk_t *k = &trk[T + i];
      w2t_oracle_kf_init(dets + 4 * new_order[i], k->x, k->P, p->promotion);
      k->tsu = 0;
      k->hit_streak = 0;
      k->birth_g = g;
      k->birth_k = i;
    }
    T += n_new;
    r->created[g] = n_new;
    /* sort.py:280-293 + utils.py:37-58.  The reference walks the list in REVERSE; rows are
     * stored here in list order (same convention as the CUDA library, include/w2t_types.h):
     * whether a tracker emits a row does not depend on the walk direction. */
    int emitted = 0;
    for (int t = 0; t < T; t++) {
      trk_t *k = &trk[t];
      if (k->tsu < 1 && (k->hit_streak >= p->min_hits || frame_count <= p->min_hits)) {
        double b[4];
        w2t_oracle_x_to_bbox(k->x, b);
        double err = ((k->P[0] + k->P[8]) + k->P[16]) / 3.0;
        double conf = exp(-err * 0.1);
        double x1 = clipd(b[0], 0., W), y1 = clipd(b[1], 0., H);
        double x2 = clipd(b[2], 0., W), y2 = clipd(b[3], 0., H);
        double width = x2 - x1, height = y2 - y1;
        if (!(width < 1 || height < 1)) {
          size_t o = (size_t)base + emitted;
          r->out_box[4 * o + 0] = x1;
          r->out_box[4 * o + 1] = y1;
          r->out_box[4 * o + 2] = width;
          r->out_box[4 * o + 3] = height;
          r->out_score[o] = clipd(conf, 0.2, 1.0);
          r->out_birth[2 * o + 0] = k->birth_g;
          r->out_birth[2 * o + 1] = k->birth_k;
          emitted++;
        }
      }
    }
    r->out_count[g] = emitted;
    int keep = 0;
    for (int t = 0; t < T; t++)
      if (!(trk[t].tsu > p->max_age)) {
        if (keep != t) trk[keep] = trk[t];
        keep++;
      }
    T = keep;
  }
  if (r->final_count) {
    /* state of every live tracker, predicted one step past the last image (the CUDA kernel
     * runs the next image's predict at the end of the current one) */
    r->final_count[q] = T;
    if (r->final_state)
      for (int t = 0; t < T && t < r->final_cap; t++) {
        double *dst = r->final_state + ((size_t)q * r->final_cap + t) * 56;
        w2t_oracle_kf_predict(trk[t].x, trk[t].P);
        memcpy(dst, trk[t].x, 7 * sizeof(double));
        memcpy(dst + 7, trk[t].P, 49 * sizeof(double));
      }
  }
  free(trk); free(boxes); free(match); free(new_order);
  return status;
}

int w2t_oracle_sort_track(const w2t_sort_problem_t *p, w2t_sort_result_t *r) {
  if (!p || !r || p->n_classes < 1 || p->n_classes > W2T_MAX_CLASSES || p->n_streams < 0) return W2T_ERR_ARG;
  int status = W2T_OK;
  for (int s = 0; s < p->n_streams; s++)
    for (int c = 0; c < p->n_classes; c++) {
      int st = track_substream(p, r, s, c);
      if (st != W2T_OK) status = st;
    }
  return status;
}

/* ------------------------------------------------------------------ */
/* soft-NMS ensemble                                                   */
/* ------------------------------------------------------------------ */

typedef struct { double s; int i; } sk_t;

static int sk_cmp(const void *a, const void *b) {
  /* stable ascending sort consumed from the end == descending score, larger index first on ties */
  const sk_t *x = (const sk_t *)a, *y = (const sk_t *)b;
  if (x->s > y->s) return -1;
  if (x->s < y->s) return 1;
  return (x->i > y->i) ? -1 : (x->i < y->i);
}

static inline double pair_weight(const double *bi, double area_i, const double *bj, double area_j,
                                 double cut, double denom) {
  /* box_utils.py:349-370: bi is the kept (higher ranked) box */
  double xx1 = bj[0] > bi[0] ? bj[0] : bi[0]; /* clamp(x1[idx], min=x1[i]) */
  double yy1 = bj[1] > bi[1] ? bj[1] : bi[1];
  double xx2 = bj[2] < bi[2] ? bj[2] : bi[2]; /* clamp(x2[idx], max=x2[i]) */
  double yy2 = bj[3] < bi[3] ? bj[3] : bi[3];
  double w = xx2 - xx1; if (w < 0.) w = 0.;
  double h = yy2 - yy1; if (h < 0.) h = 0.;
  double inter = w * h;
  double uni = (area_j - inter) + area_i;
  double iou = inter / uni;
  double wt = (cut - iou) / denom;
  if (wt < 0.) wt = 0.;
  if (wt > 1.) wt = 1.;
  return wt;
}

int w2t_oracle_soft_nms(const double *boxes, const double *scores, int n, double overlap, int top_k,
                        double conf_thresh, double soft_nms_cut, int32_t *keep, double *new_scores) {
  if (n <= 0) return 0;
  sk_t *order = (sk_t *)malloc(sizeof(sk_t) * n);
  double *area = (double *)malloc(sizeof(double) * n);
  double *live = (double *)malloc(sizeof(double) * n);
  unsigned char *alive = (unsigned char *)malloc(n);
  for (int i = 0; i < n; i++) { order[i].s = scores[i]; order[i].i = i; }
  qsort(order, n, sizeof(sk_t), sk_cmp);
  int m = n;
  if (top_k > 0 && top_k < n) m = top_k;
  for (int i = 0; i < n; i++)
    area[i] = (boxes[4 * i + 2] - boxes[4 * i]) * (boxes[4 * i + 3] - boxes[4 * i + 1]);
  for (int k = 0; k < m; k++) { live[k] = order[k].s; alive[k] = 1; }
  const double denom = soft_nms_cut - overlap;
  int kept = 0;
  /* the reference's loop runs while more than one candidate remains; the last one is appended
   * unconditionally (box_utils.py:344,386-388) */
  int remaining = m;
  for (int k = 0; k < m; k++) {
    if (!alive[k]) continue;
    keep[kept] = order[k].i;
    new_scores[kept] = live[k];
    kept++;
    remaining--;
    if (remaining == 0) break;
    const int bi = order[k].i;
    for (int j = k + 1; j < m; j++) {
      if (!alive[j]) continue;
      const int bj = order[j].i;
      double wt = pair_weight(boxes + 4 * bi, area[bi], boxes + 4 * bj, area[bj], soft_nms_cut, denom);
      live[j] = live[j] * wt;
      if (!(live[j] >= conf_thresh)) { alive[j] = 0; remaining--; }
    }
  }
  free(order); free(area); free(live); free(alive);
  return kept;
}

static int hk_cmp(const void *a, const void *b) {
  /* reference: stable ascending scores.sort(0) (box_utils.py:324, canonical tie rule), then
   * torchvision's stable DESCENDING sort of that array: descending score, smaller index first */
  const sk_t *x = (const sk_t *)a, *y = (const sk_t *)b;
  if (x->s > y->s) return -1;
  if (x->s < y->s) return 1;
  return (x->i < y->i) ? -1 : (x->i > y->i);
}

/* nms(soft=False), box_utils.py:329-333 -> torchvision.ops.nms (third-party, torchvision/csrc/ops/cpu/
 * nms_kernel.cpp nms_kernel_impl): greedy, suppress j when inter / (area_i + area_j - inter) > overlap. */
int w2t_oracle_hard_nms(const double *boxes, const double *scores, int n, double overlap, int top_k,
                        int32_t *keep) {
  if (n <= 0) return 0;
  sk_t *order = (sk_t *)malloc(sizeof(sk_t) * n);
  unsigned char *dead = (unsigned char *)calloc(n, 1);
  for (int i = 0; i < n; i++) { order[i].s = scores[i]; order[i].i = i; }
  qsort(order, n, sizeof(sk_t), hk_cmp);
  int m = n;
  if (top_k > 0 && top_k < n) {
    /* idx[-top_k:] of the stable ASCENDING sort keeps, among equal scores at the cut, the later boxes */
    qsort(order, n, sizeof(sk_t), sk_cmp);
    m = top_k;
    qsort(order, m, sizeof(sk_t), hk_cmp);
  }
  int kept = 0;
  for (int a = 0; a < m; a++) {
    if (dead[a]) continue;
    const double *bi = boxes + 4 * order[a].i;
    const double iarea = (bi[2] - bi[0]) * (bi[3] - bi[1]);
    keep[kept++] = order[a].i;
    for (int b = a + 1; b < m; b++) {
      if (dead[b]) continue;
      const double *bj = boxes + 4 * order[b].i;
      double xx1 = bi[0] > bj[0] ? bi[0] : bj[0];
      double yy1 = bi[1] > bj[1] ? bi[1] : bj[1];
      double xx2 = bi[2] < bj[2] ? bi[2] : bj[2];
      double yy2 = bi[3] < bj[3] ? bi[3] : bj[3];
      double w = xx2 - xx1; if (!(w > 0.)) w = 0.;
      double h = yy2 - yy1; if (!(h > 0.)) h = 0.;
      double inter = w * h;
      double jarea = (bj[2] - bj[0]) * (bj[3] - bj[1]);
      double ovr = inter / ((iarea + jarea) - inter);
      if (ovr > overlap) dead[b] = 1;
    }
  }
  free(order); free(dead);
  return kept;
}

static int nms_groups(const w2t_nms_problem_t *p, w2t_nms_result_t *r, int hard);

int w2t_oracle_softnms_groups(const w2t_nms_problem_t *p, w2t_nms_result_t *r) { return nms_groups(p, r, 0); }
int w2t_oracle_hardnms_groups(const w2t_nms_problem_t *p, w2t_nms_result_t *r) { return nms_groups(p, r, 1); }

static int nms_groups(const w2t_nms_problem_t *p, w2t_nms_result_t *r, int hard) {
  if (!p || !r || p->n_groups < 0) return W2T_ERR_ARG;
  const int NC = p->n_classes > 0 ? p->n_classes : 1;
  if (r->img_exists && p->n_classes > 0)
    memset(r->img_exists, 0, (size_t)(p->n_groups / NC));
  int maxn = 0;
  for (int g = 0; g < p->n_groups; g++) {
    int n = p->group_offsets[g + 1] - p->group_offsets[g];
    if (n > maxn) maxn = n;
  }
  double *pf = (double *)malloc(sizeof(double) * 4 * (maxn + 1));
  double *sc = (double *)malloc(sizeof(double) * (maxn + 1));
  double *ns = (double *)malloc(sizeof(double) * (maxn + 1));
  int32_t *keep = (int32_t *)malloc(sizeof(int32_t) * (maxn + 1));
  for (int g = 0; g < p->n_groups; g++) {
    const int base = p->group_offsets[g];
    const int n = p->group_offsets[g + 1] - base;
    const double *rows = p->rows + 5 * (size_t)base;
    for (int i = 0; i < n; i++) {
      double unpacked[5];
      const double *rw = rows + 5 * i;
      if (p->box_format == W2T_BOX_LTWH_I16) { /* 16-byte compact rows: double score + 4 x int16 */
        const unsigned char *q = (const unsigned char *)p->rows + 16 * ((size_t)base + i);
        short b[4];
        memcpy(&unpacked[0], q, 8);
        memcpy(b, q + 8, 8);
        for (int k = 0; k < 4; k++) unpacked[1 + k] = (double)b[k];
        rw = unpacked;
      }
      if (p->box_format == W2T_BOX_LTWH_P64) { /* 8-byte packed rows, layout in include/w2t_types.h */
        unsigned long long q;
        memcpy(&q, (const unsigned char *)p->rows + 8 * ((size_t)base + i), 8);
        unpacked[0] = (double)(q & 0x1ffffu) / 100000.0;
        unpacked[1] = (double)((int)((q >> 17) & 0x1fffu) - 3072);
        unpacked[2] = (double)((int)((q >> 30) & 0xfffu) - 1536);
        unpacked[3] = (double)((q >> 42) & 0x7ffu);
        unpacked[4] = (double)(q >> 53);
        rw = unpacked;
      }
      if (p->box_format == W2T_BOX_XYXY) {
        for (int k = 0; k < 4; k++) pf[4 * i + k] = rw[1 + k];
      } else {
        /* ensemble.py:19-22 (rows of convert_submission only) then box_utils.py:32-35 */
        double cx = rw[1], cy = rw[2];
        if (p->box_format != W2T_BOX_CXCYWH) { cx = cx + rw[3] / 2; cy = cy + rw[4] / 2; }
        double hw = rw[3] * 0.5, hh = rw[4] * 0.5;
        pf[4 * i + 0] = cx - hw;
        pf[4 * i + 1] = cy - hh;
        pf[4 * i + 2] = cx + hw;
        pf[4 * i + 3] = cy + hh;
      }
      sc[i] = rw[0];
    }
    int kept;
    if (hard) {
      kept = w2t_oracle_hard_nms(pf, sc, n, p->iou_thresh, p->top_k, keep);
      for (int k = 0; k < kept; k++) ns[k] = sc[keep[k]];
    } else {
      kept = w2t_oracle_soft_nms(pf, sc, n, p->iou_thresh, p->top_k, p->conf_thresh, p->soft_nms_cut, keep, ns);
    }
    if (r->kept_count) r->kept_count[g] = kept;
    int n_ens = 0, n_trk = 0;
    const int c = g % NC;
    for (int k = 0; k < kept; k++) {
      const double *b = pf + 4 * keep[k];
      /* box_utils.py:57-69 then ensemble.py:25-28 */
      double cx = (b[0] + b[2]) * 0.5, cy = (b[1] + b[3]) * 0.5;
      double w = b[2] - b[0], h = b[3] - b[1];
      size_t o = (size_t)base + k;
      if (r->merged) {
        double *mr = r->merged + 5 * o;
        mr[0] = ns[k]; mr[1] = cx; mr[2] = cy; mr[3] = w; mr[4] = h;
      }
      if (r->src_index) r->src_index[o] = base + keep[k];
      if (ns[k] > p->min_score) {
        double left = cx - w / 2, top = cy - h / 2;
        long bx = (long)left, by = (long)top, bw = (long)w, bh = (long)h; /* astype(int) */
        double rs = rint(ns[k] * 1e5) / 1e5;                              /* round(np.float64, 5) */
        size_t e = (size_t)base + n_ens;
        if (r->ens_box) {
          r->ens_box[4 * e + 0] = (int32_t)bx;
          r->ens_box[4 * e + 1] = (int32_t)by;
          r->ens_box[4 * e + 2] = (int32_t)bw;
          r->ens_box[4 * e + 3] = (int32_t)bh;
          r->ens_score[e] = rs;
        }
        n_ens++;
        if (p->score_thr && r->trk_box) {
          /* utils.py:79-87 then :32-35 and tracker_sort.py:45 */
          if (!(bw < 1 || bh < 1) && !(rs < p->score_thr[c])) {
            size_t t = (size_t)base + n_trk;
            r->trk_box[4 * t + 0] = (float)bx;
            r->trk_box[4 * t + 1] = (float)by;
            r->trk_box[4 * t + 2] = (float)(bx + bw);
            r->trk_box[4 * t + 3] = (float)(by + bh);
            n_trk++;
          }
        }
      }
    }
    r->ens_count[g] = n_ens;
    if (r->trk_count) r->trk_count[g] = n_trk;
    if (r->img_exists && p->n_classes > 0 && n_ens > 0) r->img_exists[g / NC] = 1;
  }
  free(pf); free(sc); free(ns); free(keep);
  return W2T_OK;
}

/* ------------------------------------------------------------------ */
/* weighted box fusion                                                 */
/* ------------------------------------------------------------------ */

/* merge_detections, detnet/nn/tta.py:22-66, with jaccard_bbox (detnet/utils/box_utils.py:72-140).
 * rows[n,5] = score, cx, cy, w, h of all submissions concatenated; counts[n_sub] rows per submission.
 * out[n,5] receives the result list; returns its length. */
int w2t_oracle_merge_detections(const double *rows, const int32_t *counts, int n_sub, double nms_thresh,
                                double *out) {
  int n = 0;
  for (int k = 0; k < n_sub; k++) n += counts[k];
  if (n == 0) return 0;
  double *acc = out;                                   /* the running results, weighted rows */
  double *box = (double *)malloc(sizeof(double) * 5 * n); /* de-weighted point form + area */
  int *last = (int *)malloc(sizeof(int) * n);
  int *fresh = (int *)malloc(sizeof(int) * n);
  const double N = (double)n_sub;
  int R = 0, row0 = 0;
  for (int k = 0; k < n_sub; k++) {
    const int nk = counts[k];
    if (nk == 0) continue;                             /* tta.py:38 (and an empty first submission) */
    const double *oth = rows + 5 * (size_t)row0;
    if (R == 0) {
      for (int o = 0; o < nk; o++) {
        const double s = oth[5 * o] / N;               /* tta.py:34-35 / :39-40 */
        acc[5 * o] = s;
        for (int i = 1; i < 5; i++) acc[5 * o + i] = oth[5 * o + i] * s;
      }
      R = nk;
      row0 += nk;
      continue;
    }
    for (int r = 0; r < R; r++) {                      /* tta.py:43, box_utils.py:32-35,123-128 */
      const double s = acc[5 * r];
      const double cx = acc[5 * r + 1] / s, cy = acc[5 * r + 2] / s, w = acc[5 * r + 3] / s, h = acc[5 * r + 4] / s;
      box[5 * r + 0] = cx - w * 0.5; box[5 * r + 1] = cy - h * 0.5;
      box[5 * r + 2] = cx + w * 0.5; box[5 * r + 3] = cy + h * 0.5;
      box[5 * r + 4] = w * h;
      last[r] = -1;
    }
    int n_fresh = 0;
    for (int o = 0; o < nk; o++) {
      const double s = oth[5 * o] / N;
      const double wc[4] = {oth[5 * o + 1] * s, oth[5 * o + 2] * s, oth[5 * o + 3] * s, oth[5 * o + 4] * s};
      const double cx = wc[0] / s, cy = wc[1] / s, w = wc[2] / s, h = wc[3] / s;   /* tta.py:44 */
      const double x1 = cx - w * 0.5, y1 = cy - h * 0.5, x2 = cx + w * 0.5, y2 = cy + h * 0.5;
      const double area = w * h;
      double best = -1.0;
      int arg = 0;
      for (int r = 0; r < R; r++) {                    /* box_utils.py:72-92,113-121 */
        double iw = (box[5 * r + 2] < x2 ? box[5 * r + 2] : x2) - (box[5 * r + 0] > x1 ? box[5 * r + 0] : x1);
        double ih = (box[5 * r + 3] < y2 ? box[5 * r + 3] : y2) - (box[5 * r + 1] > y1 ? box[5 * r + 1] : y1);
        if (iw < 0.) iw = 0.;
        if (ih < 0.) ih = 0.;
        const double inter = iw * ih;
        const double iou = inter / ((box[5 * r + 4] + area) - inter);
        if (iou > best) { best = iou; arg = r; }       /* overlaps.max(0): first maximum */
      }
      if (best >= nms_thresh) last[arg] = o;           /* fancy-index +=: the last duplicate wins */
      else if (best < nms_thresh) fresh[n_fresh++] = o;
    }
    for (int r = 0; r < R; r++) {
      const int o = last[r];
      if (o < 0) continue;
      const double s = oth[5 * o] / N;
      acc[5 * r] = acc[5 * r] + s;
      for (int i = 1; i < 5; i++) acc[5 * r + i] = acc[5 * r + i] + oth[5 * o + i] * s;
    }
    for (int f = 0; f < n_fresh; f++) {                /* tta.py:56-58 */
      const int o = fresh[f];
      const double s = oth[5 * o] / N;
      acc[5 * R] = s;
      for (int i = 1; i < 5; i++) acc[5 * R + i] = oth[5 * o + i] * s;
      R++;
    }
    row0 += nk;
  }
  for (int r = 0; r < R; r++)                          /* tta.py:67 */
    for (int i = 1; i < 5; i++) acc[5 * r + i] = acc[5 * r + i] / acc[5 * r];
  free(box); free(last); free(fresh);
  return R;
}

/* detnet/ensemble.py:50-64 with merge_func = merge_detections, for all groups (host pointers). */
int w2t_oracle_fusion_groups(const w2t_nms_problem_t *p, const int32_t *sub_counts, int n_sub,
                             w2t_nms_result_t *r) {
  if (!p || !r || !sub_counts || p->n_groups < 0 || n_sub < 1) return W2T_ERR_ARG;
  const int NC = p->n_classes > 0 ? p->n_classes : 1;
  if (r->img_exists && p->n_classes > 0) memset(r->img_exists, 0, (size_t)(p->n_groups / NC));
  int maxn = 0;
  for (int g = 0; g < p->n_groups; g++) {
    int n = p->group_offsets[g + 1] - p->group_offsets[g];
    if (n > maxn) maxn = n;
  }
  double *in = (double *)malloc(sizeof(double) * 5 * (maxn + 1));
  double *out = (double *)malloc(sizeof(double) * 5 * (maxn + 1));
  for (int g = 0; g < p->n_groups; g++) {
    const int base = p->group_offsets[g];
    const int n = p->group_offsets[g + 1] - base;
    for (int i = 0; i < n; i++) {
      const double *rw = p->rows + 5 * ((size_t)base + i);
      in[5 * i] = rw[0];
      in[5 * i + 1] = rw[1]; in[5 * i + 2] = rw[2]; in[5 * i + 3] = rw[3]; in[5 * i + 4] = rw[4];
      if (p->box_format == W2T_BOX_LTWH) {             /* lxly2cxcy, ensemble.py:19-22 */
        in[5 * i + 1] = rw[1] + rw[3] / 2;
        in[5 * i + 2] = rw[2] + rw[4] / 2;
      }
    }
    const int R = w2t_oracle_merge_detections(in, sub_counts + (size_t)g * n_sub, n_sub, p->iou_thresh, out);
    if (r->kept_count) r->kept_count[g] = R;
    int n_ens = 0, n_trk = 0;
    const int c = g % NC;
    for (int k = 0; k < R; k++) {
      const double s = out[5 * k], cx = out[5 * k + 1], cy = out[5 * k + 2], w = out[5 * k + 3], h = out[5 * k + 4];
      if (r->merged) {
        double *mr = r->merged + 5 * ((size_t)base + k);
        mr[0] = s; mr[1] = cx; mr[2] = cy; mr[3] = w; mr[4] = h;
      }
      if (s > p->min_score) {
        double left = cx - w / 2, top = cy - h / 2;
        long bx = (long)left, by = (long)top, bw = (long)w, bh = (long)h;
        double rs = rint(s * 1e5) / 1e5;
        size_t e = (size_t)base + n_ens;
        if (r->ens_box) {
          r->ens_box[4 * e + 0] = (int32_t)bx; r->ens_box[4 * e + 1] = (int32_t)by;
          r->ens_box[4 * e + 2] = (int32_t)bw; r->ens_box[4 * e + 3] = (int32_t)bh;
          r->ens_score[e] = rs;
        }
        n_ens++;
        if (p->score_thr && r->trk_box && !(bw < 1 || bh < 1) && !(rs < p->score_thr[c])) {
          size_t t = (size_t)base + n_trk;
          r->trk_box[4 * t + 0] = (float)bx; r->trk_box[4 * t + 1] = (float)by;
          r->trk_box[4 * t + 2] = (float)(bx + bw); r->trk_box[4 * t + 3] = (float)(by + bh);
          n_trk++;
        }
      }
    }
    r->ens_count[g] = n_ens;
    if (r->trk_count) r->trk_count[g] = n_trk;
    if (r->img_exists && p->n_classes > 0 && n_ens > 0) r->img_exists[g / NC] = 1;
  }
  free(in); free(out);
  return W2T_OK;
}

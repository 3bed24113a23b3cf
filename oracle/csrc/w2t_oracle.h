/*
 * w2t_oracle.h — CPU restatement (plain C) of the reference hot path.
 *
 * TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library, as the checker
 * or as the timed CPU baseline.  The product never links it.
 *
 * Same packed layout as the CUDA library (include/w2t_types.h), with HOST
 * pointers, so a test can hand the same buffers to both and compare bit for bit.
 *
 * Floating-point operation order is fixed and written out explicitly (fma()
 * where the reference's BLAS uses FMA, separate multiply/add where it does
 * not); compile with -ffp-contract=off.  The order reproduces what NumPy
 * 2.3 / OpenBLAS 0.3.30 compute for the reference's calls on an AVX-512 host
 * (k-sequential FMA dgemm, (p0+p2)+(p1+p3) dgemv, left-looking LU + FMA
 * triangular solves behind numpy.linalg.inv): see DESIGN.md "Kalman numerics".
 */
#ifndef W2T_ORACLE_H
#define W2T_ORACLE_H
#include "w2t_types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* sort.py:50-62 on a float32 row (NEP 50): z = x, y, s, r all float32 */
void w2t_oracle_bbox_to_z(const float det[4], float z[4]);
/* the same under either NumPy promotion regime (W2T_PROMOTION_*), as the float64 values the filter receives */
void w2t_oracle_bbox_to_z_d(const float det[4], int promotion, double z[4]);
/* sort.py:65-75 */
void w2t_oracle_x_to_bbox(const double x[7], double box[4]);
/* sort.py:97-137: x0 = [z,0,0,0], P0 = diag(10,10,10,10,1e4,1e4,1e4) */
void w2t_oracle_kf_init(const float det[4], double x[7], double P[49], int promotion);
/* sort.py:170-172 + filterpy predict */
void w2t_oracle_kf_predict(double x[7], double P[49]);
/* sort.py:164 + filterpy update (Joseph form, LAPACK-order 4x4 inverse) */
void w2t_oracle_kf_update(double x[7], double P[49], const float det[4], int promotion);
/* numpy.linalg.inv of a 4x4 (row-major) in OpenBLAS dgesv operation order */
void w2t_oracle_inv4(const double S[16], double out[16]);

/* sort.py:33-47,201-205: out[D,T] float32 */
void w2t_oracle_iou_matrix(const float *dets, int D, const double *trks, int T, float *out);

/* scikit-learn 0.22.2 linear_assignment on a float32 [D,T] matrix.
 * pairs[min(D,T),2] sorted by (row, col); returns the pair count, or -1 on allocation failure. */
int w2t_oracle_linear_assignment(const float *cost, int D, int T, int32_t *pairs);

/* sort.py:193-230.  matched_det_of_trk[T] = detection index or -1;
 * new_order[D] = detections that start a new tracker, in the reference's order; returns their count. */
int w2t_oracle_associate(const float *dets, int D, const double *trks, int T, double iou_threshold, int promotion,
                         int32_t *matched_det_of_trk, int32_t *new_order);

/* tracking/utils.py:25-60 for all streams (host pointers). */
int w2t_oracle_sort_track(const w2t_sort_problem_t *problem, w2t_sort_result_t *result);

/* detnet/ensemble.py:50-64 for all groups (host pointers). */
int w2t_oracle_softnms_groups(const w2t_nms_problem_t *problem, w2t_nms_result_t *result);

/* box_utils.py:307-395 soft branch with the general arguments (top_k, conf_thresh):
 * boxes[n,4] point form, returns the number of kept boxes; keep[n], new_scores[n]. */
int w2t_oracle_soft_nms(const double *boxes, const double *scores, int n, double overlap, int top_k,
                        double conf_thresh, double soft_nms_cut, int32_t *keep, double *new_scores);

/* box_utils.py:329-333 (hard branch): torchvision.ops.nms on the ascending-sorted boxes.
 * Returns the number of kept boxes; keep[n] in descending score order. */
int w2t_oracle_hard_nms(const double *boxes, const double *scores, int n, double overlap, int top_k,
                        int32_t *keep);
int w2t_oracle_hardnms_groups(const w2t_nms_problem_t *problem, w2t_nms_result_t *result);

/* detnet/nn/tta.py:22-66 merge_detections: rows[n,5] = score,cx,cy,w,h of all submissions
 * concatenated, counts[n_sub]; out[n,5]; returns the length of the result list. */
int w2t_oracle_merge_detections(const double *rows, const int32_t *counts, int n_sub, double nms_thresh,
                                double *out);
int w2t_oracle_fusion_groups(const w2t_nms_problem_t *problem, const int32_t *sub_counts, int n_sub,
                             w2t_nms_result_t *result);

#ifdef __cplusplus
}
#endif
#endif

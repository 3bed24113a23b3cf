/*
 * w2t.h — C-ABI of libw2t.so: the B200 (sm_100a) implementation of the
 * post-detection box pipeline of xuyuan/waymo_2d_tracking
 * (soft-NMS ensemble + per-class SORT).
 *
 * The reference has no FFI layer: the "operator API" of this path is its Python
 * call surface.  Each export below replaces the Python call it
 * cites; the Python drop-ins in waymo_2d_tracking_b200/ bind these symbols with
 * ctypes (INTEGRATION.md shows the stub).  Plain pointers and sizes only, no
 * torch types.  Unless stated otherwise pointers are DEVICE pointers, launches
 * are asynchronous on `stream`, nothing is allocated behind the caller's back
 * and there is no CPU fallback: without a CUDA device every compute entry
 * point returns W2T_ERR_CUDA.
 *
 * Paths are relative to /root/reference.
 */
#ifndef W2T_H
#define W2T_H

#include "w2t_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st *w2t_stream_t; /* == cudaStream_t */

/* ---- library ---------------------------------------------------------- */

const char *w2t_version(void);
/* message of the last failing call on this host thread ("" if none) */
const char *w2t_last_error(void);
/* forget it (callers that report a message call this afterwards, so that a later failure which sets no message of
 * its own is not reported with a stale one) */
void w2t_clear_error(void);
/* SM count and compute capability of the current device */
int w2t_device_info(int *sm_count, int *cc_major, int *cc_minor);

/* Stream-ordered wait: work queued on `stream` after this call starts only once
 * *addr >= value (device int32, e.g. a w2t_sort_plan_t.chunk_done counter).  Uses the driver's
 * cuStreamWaitValue32; returns W2T_ERR_CUDA if the device does not support it. */
int w2t_stream_wait_value32(w2t_stream_t stream, const int32_t *addr, int32_t value);

/* ---- soft-NMS ensemble ------------------------------------------------- */

/* Replaces the per-image loop of detnet/ensemble.py:145-157, i.e. ensemble()
 * (:50-64) -> nms_detections() (detnet/nn/tta.py:8-19) -> nms(soft=True)
 * (detnet/utils/box_utils.py:307-395) for every (image, category) group in one
 * launch, plus — when problem->score_thr is set — the filters and box
 * conversion read_data_file()/track_sort() apply to that output before
 * tracking (tracking/utils.py:79-87,32-35; tracker_sort.py:45).
 *
 * problem->box_format / top_k / conf_thresh open the general signature of nms()
 * (box_utils.py:307) to the drop-ins of nms() and nms_detections(); the ensemble CLI leaves
 * them 0.  With top_k or conf_thresh a group keeps fewer rows than it has: result->kept_count.
 *
 * max_group_size: largest group_offsets[g+1]-group_offsets[g] (host knows it
 *                 from the offsets it built).
 * status:         device int32, set to a W2T_ERR_* code by the kernel if an
 *                 input is outside what the soft-NMS path supports (negative
 *                 or NaN score, non-positive box area); must be zeroed by the
 *                 caller.  May be NULL.
 * result->img_exists, if given, must be zeroed by the caller. */
int w2t_softnms_groups(const w2t_nms_problem_t *problem, w2t_nms_result_t *result, int max_group_size,
                       int32_t *status, w2t_stream_t stream);

/* The same for the hard branch of nms() (box_utils.py:329-333: torchvision.ops.nms on the
 * ascending-sorted boxes) behind `python -m detnet.ensemble -m nms` (ensemble.py:139-140):
 * greedy suppression of lower ranked boxes with IoU > iou_thresh, scores unchanged.
 * soft_nms_cut and conf_thresh are ignored; everything else as above. */
int w2t_hardnms_groups(const w2t_nms_problem_t *problem, w2t_nms_result_t *result, int max_group_size,
                       int32_t *status, w2t_stream_t stream);

/* Confidence-weighted box fusion of every group: merge_detections() (detnet/nn/tta.py:22-66,
 * with jaccard_bbox, detnet/utils/box_utils.py:72-140), the reference CLI's default method
 * (`-m weighted_fusion`, ensemble.py:94,138).  The submissions of a group are folded into a
 * running result list in input-file order, so the kernel needs to know where each submission's
 * rows end: sub_counts[n_groups, n_sub] (device) = rows of group g that come from input file k
 * (rows of a group are concatenated in file order; n_sub = number of input files, the divisor
 * of tta.py:34).  problem->iou_thresh is nms_thresh; box_format W2T_BOX_LTWH or W2T_BOX_CXCYWH;
 * soft_nms_cut / top_k / conf_thresh are ignored.  result->merged rows and the ensemble rows are
 * in result-list order (first file's boxes, then the unmatched boxes of each later file);
 * result->kept_count[g] = length of the result list; src_index is not written. */
int w2t_fusion_groups(const w2t_nms_problem_t *problem, const int32_t *sub_counts, int32_t n_sub,
                      w2t_nms_result_t *result, int max_group_size, int32_t *status, w2t_stream_t stream);
int w2t_fusion_max_group(void);

/* Largest group the NMS kernels keep in shared memory (3401 boxes).  Larger groups are not an error: a second
 * launch of a few persistent CTAs runs the same kernel over arrays in global memory (stream-ordered scratch from
 * cudaMallocAsync, freed behind the launch) for just those groups. */
int w2t_softnms_max_group(void);

/* ---- SORT --------------------------------------------------------------- */

/* HOST function.  Sizes the per-sub-stream workspace slabs from the detection
 * counts (host copies of det_count / img_exists): a tracker alive at image f
 * was created or last updated by a distinct detection of the previous
 * max_age+1 images, so the window sum of counts bounds the tracker list
 * (tracking/sort/sort.py:276-278,292-293).  plan arrays are caller-allocated
 * host arrays of n_streams*n_classes entries. */
int w2t_sort_plan(int32_t n_streams, int32_t n_classes, const int32_t *stream_img_offsets,
                  const int32_t *det_count, const uint8_t *img_exists, int32_t max_age,
                  w2t_sort_plan_t *plan);

/* The same from the group offsets of the rows alone (group g holds rows group_offsets[g] ..
 * group_offsets[g+1]): capacities from the group SIZES, an image counted as present when it has any
 * row.  For callers that plan before the ensemble has run (sizes before soft-NMS are upper bounds of
 * what reaches the tracker), without materialising the count arrays. */
int w2t_sort_plan_offsets(int32_t n_streams, int32_t n_classes, const int32_t *stream_img_offsets,
                          const int32_t *group_offsets, int32_t max_age, w2t_sort_plan_t *plan);

/* Replaces the loop of tracking/track.py:42-47: track_sort()
 * (tracking/utils.py:25-60) -> MultiClassTrackerSort.track()
 * (tracking/sort/tracker_sort.py:22-51) -> Sort.update()
 * (tracking/sort/sort.py:244-296) for every stream.  One persistent CTA per
 * (stream, category) sub-stream walks the stream's images in order.
 * plan holds DEVICE copies of the arrays w2t_sort_plan filled.
 * Global object ids are not assigned here: every row carries the (group, k)
 * of its tracker's creation and `created` holds the per-group creation counts;
 * the id is an exclusive scan of `created` in the reference's processing order
 * (KalmanBoxTracker.count, sort.py:86,140-141) — see w2t_assign_ids. */
int w2t_sort_track(const w2t_sort_problem_t *problem, const w2t_sort_plan_t *plan,
                   w2t_sort_result_t *result, void *workspace, int32_t *status, w2t_stream_t stream);

/* Frame-by-frame tracking with device-resident state: the stateful API of the reference —
 * Sort.update (tracking/sort/sort.py:244-296) called once per image and category by
 * MultiClassTrackerSort.track (tracking/sort/tracker_sort.py:41-49).  The same kernel as
 * w2t_sort_track over a problem that holds ONE new image per stream (or a few); between calls the
 * trackers stay in the workspace slabs and `sub_state` (DEVICE int32 [n_streams*n_classes, 4],
 * zeroed by the caller before the first call) carries each sub-stream's live-tracker count,
 * frame_count, "Sort object exists" and flags.  The plan is the caller's: fixed capacities
 * track_cap / det_cap per sub-stream, slabs of w2t_sort_slab_bytes(track_cap, det_cap) bytes
 * at ws_offset (no window bound exists for an open-ended stream); exceeding a capacity reports
 * W2T_ERR_CAPACITY in *status.  Rows are those Sort.update returns (sort.py:280-289): out_box =
 * [x1, y1, x2, y2] and out_score = exp(-0.1 * error), both unclipped, in tracker-list order (the
 * reference returns the reversed list); clipping, the size filter and the confidence clip of
 * tracking/utils.py:37-58 are left to the caller.  out_birth = (group_base + group of the image
 * the tracker was created in, k): the caller, who knows KalmanBoxTracker.count at every
 * Sort.update call, turns it into the object id.  result->first_img is not used. */
int w2t_sort_step(const w2t_sort_problem_t *problem, const w2t_sort_plan_t *plan,
                  w2t_sort_result_t *result, void *workspace, int32_t *sub_state, int32_t group_base,
                  int32_t *status, w2t_stream_t stream);
size_t w2t_sort_slab_bytes(int32_t track_cap, int32_t det_cap);

/* Device-side id assignment + dense output list (finalize.cu): object ids by an exclusive
 * scan of `created` in the reference's processing order (KalmanBoxTracker.count,
 * sort.py:86,140-141), rows gathered into the order tracking/utils.py:37-58 appends them.
 * class_rank: DEVICE [n_streams*n_classes] position of each category in the stream's tracker
 * dict, or NULL to rank by (first_img, category id).  workspace:
 * w2t_sort_finalize_workspace() bytes.  n_groups = n_img * n_classes. */
size_t w2t_sort_finalize_workspace(int32_t n_streams, int32_t n_classes, int64_t n_groups);
int w2t_sort_finalize(const w2t_sort_problem_t *problem, const w2t_sort_result_t *result,
                      const int32_t *class_rank, int64_t id_base, int64_t n_groups, void *workspace,
                      w2t_rows_t *rows, w2t_stream_t stream);

/* HOST function.  Turns (birth group, k) into the reference's global
 * object_id (= KalmanBoxTracker.id + 1, sort.py:288).  All pointers are host
 * pointers.  class_rank[n_streams*n_classes] gives, per stream, the position
 * of each category in the reference's tracker dict (first-appearance order,
 * tracker_sort.py:32-33,41); pass NULL to rank by (first_img, category id).
 * id_base is the value of KalmanBoxTracker.count before the call; returns the
 * count after it through id_next. */
int w2t_assign_ids(int32_t n_streams, int32_t n_classes, const int32_t *stream_img_offsets,
                   const int32_t *det_start, const int32_t *out_count, const int32_t *created,
                   const int32_t *first_img, const int32_t *class_rank, const int32_t *out_birth,
                   int64_t id_base, int64_t *out_id, int64_t *id_next);

/* ---- JSON in / out (host code) ----------------------- */

/* Replaces json.load + the per-dict loops of detnet/ensemble.py:79 (load_input_submissions) and
 * tracking/utils.py:65-67 (read_data_file): parses a submission / annotation file — a list of
 * {"image_id", "category_id", "bbox": [x,y,w,h], "score"} objects, or {"annotations": [...]};
 * "score" is optional (ground truth), other keys are skipped — straight into flat arrays.
 * Image ids are interned in first-appearance order.  All functions here are HOST functions. */
typedef struct w2t_json_dets w2t_json_dets_t;
int w2t_json_load(const char *path, w2t_json_dets_t **out);
int64_t w2t_json_count(const w2t_json_dets_t *dets);     /* rows                    */
int64_t w2t_json_n_images(const w2t_json_dets_t *dets);  /* distinct image ids      */
/* any pointer may be NULL; bbox is [count,4]; has_score[i] = 0 where the row had no "score" (score 1.0) */
int w2t_json_copy(const w2t_json_dets_t *dets, int32_t *image_index, int32_t *category, double *bbox,
                  double *score, uint8_t *has_score);
/* the distinct image ids, each followed by '\n', in first-appearance order; valid until w2t_json_free */
const char *w2t_json_image_ids(w2t_json_dets_t *dets, int64_t *bytes);
void w2t_json_free(w2t_json_dets_t *dets);

/* load_input_submissions + the grouping of ensemble_detections (detnet/ensemble.py:31-47,78-95) on files: parses
 * the n_files submissions (concurrently), drops rows with width or height <= 0 or score * weight < min_score
 * (ensemble.py:37-42), and groups the rest by (image, category): images that keep a row, in sorted image_id order
 * (W2T_LAYOUT_ENSEMBLE, columns = the category ids that occur, ascending) or — for the tracker behind the ensemble —
 * as the streams read_data_file + track.py (tracking/utils.py:63-96, track.py:36-44) would form from the ensemble's
 * output file: segments and cameras in first-appearance order of the sorted list, frames in numeric order, columns =
 * categories 1..n_classes (W2T_LAYOUT_STREAMS).  Rows of a group: file order, then JSON order (tta.py:9-12).
 * weights may be NULL (all 1).  W2T_ERR_UNSUPPORTED: use the general (Python) packer.  HOST functions. */
typedef struct w2t_json_groups w2t_json_groups_t;
int w2t_json_group_files(const char *const *paths, int32_t n_files, const double *weights, double min_score,
                         int32_t layout, int32_t n_classes, w2t_json_groups_t **out);
/* info: [0] images, [1] columns per image, [2] rows, [3] largest group, [4] streams (W2T_LAYOUT_STREAMS),
 *       [5] 1 if every row fits the 8-byte W2T_BOX_LTWH_P64 format exactly, [6] category ids, [7] files */
int w2t_json_groups_info(const w2t_json_groups_t *groups, int64_t info[8]);
/* any pointer may be NULL.  category_ids [info 6]; image_order [images]: layout position -> index into the sorted
 * image ids; stream_img_offsets [streams+1] and frame_ids [images] (W2T_LAYOUT_STREAMS only); group_offsets
 * [images*columns+1]; sub_counts [images*columns, files]; rows [rows,5] = score*weight, left, top, width, height;
 * packed [rows] (only written when info[5] is 1) */
int w2t_json_groups_copy(const w2t_json_groups_t *groups, int32_t *category_ids, int32_t *image_order,
                         int32_t *stream_img_offsets, int64_t *frame_ids, int32_t *group_offsets, int32_t *sub_counts,
                         double *rows, uint64_t *packed);
/* the sorted image ids, each followed by '\n'; valid until w2t_json_groups_free */
const char *w2t_json_groups_image_ids(const w2t_json_groups_t *groups, int64_t *bytes);
void w2t_json_groups_free(w2t_json_groups_t *groups);

/* read_data_file (tracking/utils.py:63-96) and the packing of its result for w2t_sort_track, straight from the file:
 * parses `path` (a detection list or {"annotations": [...]}), drops rows with width or height < 1 and rows whose
 * score is below score_thr[category_id - 1] (utils.py:79-87, in that order), and lays the rest out as the tracker
 * wants them: streams = (segment, camera) pairs in first-appearance order of the file (segments, then cameras
 * inside a segment), every image of a stream in frame order whether or not a row of it survives, group g = image *
 * n_classes + category_id - 1 with rows in file order, boxes x, y, x+w, y+h rounded to float32
 * (tracker_sort.py:45), class_rank = position of the category in the stream's tracker dict (order of the first
 * surviving row).  segment_id (may be NULL): keep that segment only (track.py --segment-id); block_world > 1:
 * keep this rank's contiguous block of the remaining segments (one process per GPU).  W2T_ERR_UNSUPPORTED: use the
 * general (Python) packer, which reproduces the reference's behaviour or its exception.  HOST functions. */
typedef struct w2t_json_tracks w2t_json_tracks_t;
int w2t_json_pack_tracks(const char *path, const double *score_thr, int32_t n_thr, int32_t n_classes,
                         const char *segment_id, int32_t block_rank, int32_t block_world, w2t_json_tracks_t **out);
/* info: [0] streams, [1] images, [2] rows, [3] bytes of the stream names ("segment\tcamera\n" per stream) */
int w2t_json_tracks_info(const w2t_json_tracks_t *tracks, int64_t info[4]);
/* any pointer may be NULL: stream_img_offsets [streams+1], frame_ids [images], det_start / det_count
 * [images*n_classes], det_box [rows,4] float32, class_rank [streams*n_classes], stream_names [info 3] */
int w2t_json_tracks_copy(const w2t_json_tracks_t *tracks, int32_t *stream_img_offsets, int64_t *frame_ids,
                         int32_t *det_start, int32_t *det_count, float *det_box, int32_t *class_rank,
                         char *stream_names);
void w2t_json_tracks_free(w2t_json_tracks_t *tracks);

/* Replace json.dump of tracking/track.py:50 (rows of tracking/utils.py:52-58) and of
 * detnet/ensemble.py:159-160 (rows of :61-62).  Byte-identical to the reference's files: default
 * separators, ensure_ascii, floats as Python's float.__repr__.  image_ids[k] = NUL-terminated
 * id of image k; image[i] indexes it. */
int w2t_json_write_tracks(const char *path, int64_t n, const char *const *image_ids, const int32_t *image,
                          const double *bbox, const double *score, const int32_t *category,
                          const int64_t *object_id);
int w2t_json_write_detections(const char *path, int64_t n, const char *const *image_ids, const int32_t *image,
                              const int32_t *category, const int32_t *bbox, const double *score);

/* Replaces create_pb_submission / create_pd_objects / create_pd_object of coco_to_waymo.py:16-82, the
 * consumer of both output files (README.md:46,55 of the reference): writes the serialized
 * waymo_open_dataset Submission (or, with objects_only, just its metrics.Objects message) for rows in
 * the layout of the JSON writers above.  object_id NULL = detection rows (no Label.id).  Strings are NUL-terminated; a NULL
 * string is an unset field.  Schema: field numbers restated from the package's published protos
 * (csrc/waymo_pb.cpp; PARITY UNPINNED there — the package is not available to check against). */
int w2t_pb_write_submission(const char *path, int32_t objects_only, int32_t task, const char *account_name,
                            const char *unique_method_name, const char *const *authors, int32_t n_authors,
                            const char *affiliation, const char *description, const char *method_link,
                            int32_t sensor_type, int64_t n, const char *const *image_ids, int64_t n_images,
                            const int32_t *image, const double *bbox, const double *score,
                            const int32_t *category, const int64_t *object_id);

/* ---- building blocks (unit-parity entry points) ------------------------- */

/* iou() of sort.py:33-47 for every (detection, tracker) pair, as filled into
 * the float32 matrix of sort.py:201-205.  out[D,T]. */
int w2t_iou_matrix(const float *dets, int32_t D, const double *trks, int32_t T, float *out,
                   w2t_stream_t stream);

/* scikit-learn 0.22.2 linear_assignment (call site sort.py:206) on a float32
 * [D,T] cost matrix.  pairs[min(D,T),2] sorted by (row, col); *n_pairs =
 * min(D,T).  workspace: w2t_linear_assignment_workspace(D,T) bytes. */
size_t w2t_linear_assignment_workspace(int32_t D, int32_t T);
int w2t_linear_assignment(const float *cost, int32_t D, int32_t T, int32_t *pairs, int32_t *n_pairs,
                          void *workspace, w2t_stream_t stream);

/* KalmanBoxTracker.__init__/predict/update (sort.py:88-178) on n independent
 * filters: x[n,7], P[n,49] row-major, dets[n,4] float32 x1,y1,x2,y2;
 * boxes[n,4] (may be NULL) receives convert_x_to_bbox(x) (sort.py:65-75).
 * promotion: W2T_PROMOTION_* — how convert_bbox_to_z (sort.py:50-62) treats the float32 row. */
int w2t_kf_init(double *x, double *P, const float *dets, int32_t n, int32_t promotion, w2t_stream_t stream);
int w2t_kf_predict(double *x, double *P, double *boxes, int32_t n, w2t_stream_t stream);
int w2t_kf_update(double *x, double *P, const float *dets, double *boxes, int32_t n, int32_t promotion,
                  w2t_stream_t stream);

/* convert_bbox_to_z (sort.py:50-62) of n float32 rows x1,y1,x2,y2 -> z[n,4] = x, y, s, r as the float64 values
 * the filter receives: under W2T_PROMOTION_NEP50 every component is a float32 value, under W2T_PROMOTION_LEGACY
 * x, y and r are float64 computations and only s = w*h is a float32 product (w2t_types.h). */
int w2t_bbox_to_z(const float *dets, double *z, int32_t n, int32_t promotion, w2t_stream_t stream);

/* bbox_vote (detnet/utils/box_utils.py:401-430), the box-voting step of the detector head
 * (detnet/nn/modules/detection.py:69-71): out[i] = sum_j s_j * all_boxes[j] / sum_j s_j over the boxes j with
 * IoU(nms_boxes[i], all_boxes[j]) >= thresh.  Boxes are [.,4] x1,y1,x2,y2 in float64 storage; compute_f32 = 1
 * does the arithmetic in float32 (the head's tensors).  The summation order of torch.sum is not reproduced:
 * agreement is to rounding (1e-6 relative in float32, 1e-12 in float64), not bit-exact. */
int w2t_bbox_vote(const double *nms_boxes, int32_t n, const double *all_boxes, const double *all_scores, int32_t m,
                  double thresh, int32_t compute_f32, double *out, w2t_stream_t stream);

/* convert_x_to_bbox (sort.py:65-75) of n states: row i is x[i*ldx .. i*ldx+3] = x, y, s, r (ldx >= 4);
 * boxes[n,4] = x1, y1, x2, y2. */
int w2t_x_to_bbox(const double *x, int32_t ldx, double *boxes, int32_t n, w2t_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* W2T_H */

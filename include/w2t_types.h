/*
 * w2t_types.h — plain-C data layout shared by the CUDA library (include/w2t.h)
 * and the CPU oracle (oracle/csrc/w2t_oracle.h).  No torch types, no C++.
 *
 * Vocabulary follows the reference (xuyuan/waymo_2d_tracking):
 *   stream     one (segment, camera) sequence            tracking/utils.py:25-31
 *   image      one camera frame, image_id 'seg/frame/cam' tracking/utils.py:70
 *   group      one (image, category) set of boxes          detnet/ensemble.py:52-56
 *   sub-stream one (stream, category) = one reference `Sort` object
 *                                                          tracking/sort/tracker_sort.py:32-33
 *
 * Index spaces
 *   images of stream s are contiguous and sorted by frame id:
 *       img in [stream_img_offsets[s], stream_img_offsets[s+1])
 *   group  g = img * n_classes + (category_id - 1)
 *   sub-stream q = s * n_classes + (category_id - 1)
 *   rows (detections / emitted track rows) of group g live at
 *       [det_start[g], det_start[g] + det_count[g])   (capacity: up to det_start of the next group)
 */
#ifndef W2T_TYPES_H
#define W2T_TYPES_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define W2T_MAX_CLASSES 8

/* status codes returned by every entry point */
enum {
  W2T_OK = 0,
  W2T_ERR_ARG = 1,        /* bad argument (null pointer, negative size, too many classes) */
  W2T_ERR_CAPACITY = 2,   /* a plan capacity (tracks / detections per frame) was exceeded */
  W2T_ERR_CUDA = 3,       /* CUDA runtime error; see w2t_last_error() */
  W2T_ERR_NONFINITE = 4,  /* a tracker box became +-inf (unreachable with finite inputs; the
                             reference mis-indexes its tracker list there, sort.py:261-265) */
  W2T_ERR_UNSUPPORTED = 5 /* host packers only: an input this fast path does not cover (the caller uses the
                             general path, which reproduces the reference's behaviour or its exception) */
};

/* image layouts of w2t_json_group_files */
#define W2T_LAYOUT_ENSEMBLE 0 /* images in sorted image_id order, one column per category id that occurs  */
#define W2T_LAYOUT_STREAMS 1  /* images stream by stream in frame order, columns = categories 1..n_classes */

/* Per-sub-stream launch plan, computed on the host from the detection counts
 * (w2t_sort_plan).  All arrays have n_streams * n_classes entries. */
typedef struct w2t_sort_plan_t {
  int32_t *order;        /* launch order: sub-stream ids, heaviest first                     */
  int32_t *track_cap;    /* upper bound on simultaneously live trackers (window sum of dets) */
  int32_t *det_cap;      /* max detections of that category in any one image of the stream   */
  int64_t *ws_offset;    /* byte offset of the sub-stream's slab in the workspace            */
  int64_t  ws_bytes;     /* total workspace size                                             */
  /* optional completion tracking (both NULL = off): sub-stream q belongs to chunk chunk_of[q];  */
  /* when its CTA has written everything it adds 1 to chunk_done[chunk_of[q]] (device int32,     */
  /* zeroed by the caller).  Another stream can wait for a whole chunk with                      */
  /* w2t_stream_wait_value32 and post-process it while the rest of the launch is still running.  */
  const int32_t *chunk_of;
  int32_t *chunk_done;
  /* Launch classes, by the most detections a sub-stream MAY have in one image (det_cap can be an  */
  /* upper bound; the kernels classify again from the actual counts).  `order` is                   */
  /* [wide | mid | narrow]: the first n_wide entries have det_cap > W2T_WIDE_DETS, the next n_mid    */
  /* entries det_cap > W2T_NARROW_DETS.  Narrow sub-streams are tracked by warps (persistent warps   */
  /* that pull sub-streams from queues in this order); crowded ones (more than W2T_NARROW_DETS       */
  /* detections in some image, or live trackers beyond the warp kernel's reach) by clusters of 8     */
  /* CTAs with the cost matrix in distributed shared memory; anything beyond that by single CTAs     */
  /* with the matrix in global memory.                                                               */
  int32_t  n_wide;
  int32_t  n_mid;
  /* byte offset, inside the workspace, of the auxiliary area of the warp kernel (w2t_sort_plan     */
  /* sets it and includes its W2T_SORT_AUX_BYTES(n_substreams) in ws_bytes): work-queue counters,    */
  /* one class flag per sub-stream, the two queues, and the warp kernel's spill areas.               */
  /* < 0 = no aux area: CTAs track everything.                                                       */
  int64_t  aux_offset;
  /* most detections any NARROW sub-stream has in one image: picks how many warps (= sub-streams)   */
  /* share an SM's shared memory, i.e. how large a cost matrix each can hold (0 = unknown)          */
  int32_t  narrow_cap;
} w2t_sort_plan_t;

#define W2T_WIDE_DETS 320
#define W2T_NARROW_DETS 96
#define W2T_SORT_QUEUE_BYTES(n_substreams) (((64 + 12 * (int64_t)(n_substreams)) + 255) / 256 * 256)
/* + the spill areas of the warp kernel: 160 CTAs x 8 teams x 64 KB for the rare cost matrix that outgrows a team's */
/* share of tensor memory                                                                                           */
#define W2T_SORT_SPILL_BYTES (160 * 8 * (int64_t)65536)
#define W2T_SORT_AUX_BYTES(n_substreams) (W2T_SORT_QUEUE_BYTES(n_substreams) + W2T_SORT_SPILL_BYTES)

/* NumPy promotion regime the tracker reproduces (w2t_sort_problem_t.promotion):                     */
/*   LEGACY  NumPy 1.x value-based casting — the reference's pinned environment (python 3.7,         */
/*           environment.yml:7): in convert_bbox_to_z (sort.py:50-62) x, y and r are float64, only   */
/*           s = w*h stays float32; `iou_matrix[..] < iou_threshold` (sort.py:220) compares in       */
/*           float64.                                                                                */
/*   NEP50   NumPy 2: all four components of z are float32, the threshold is compared in float32.    */
enum { W2T_PROMOTION_LEGACY = 0, W2T_PROMOTION_NEP50 = 1 };

/* Inputs of the SORT stage (tracking/utils.py:25-60 for every stream at once).
 * Pointers are device pointers for the CUDA library and host pointers for the oracle. */
typedef struct w2t_sort_problem_t {
  int32_t n_streams;
  int32_t n_classes;                 /* category ids are 1..n_classes                         */
  const int32_t *stream_img_offsets; /* [n_streams+1]                                         */
  const int32_t *det_start;          /* [n_img*n_classes]                                     */
  const int32_t *det_count;          /* [n_img*n_classes] detections that survived            */
                                     /*   read_data_file's filters (utils.py:79-87)           */
  const float   *det_box;            /* [N,4] x1,y1,x2,y2 already rounded to float32          */
                                     /*   (tracker_sort.py:45)                                */
  const uint8_t *img_exists;         /* [n_img] or NULL (= every image is present in the      */
                                     /*   input JSON, utils.py:76-77)                         */
  const double  *cam_wh;             /* [n_streams,2] IMAGE_SIZES of the camera, utils.py:11  */
  double  iou_thr[W2T_MAX_CLASSES];  /* --iou-threshold, track.py:25-26                       */
  int32_t max_age;                   /* track.py:21                                           */
  int32_t min_hits;                  /* track.py:22                                           */
  int32_t promotion;                 /* W2T_PROMOTION_*                                       */
} w2t_sort_problem_t;

/* Outputs of the SORT stage.  Row k of group g is at det_start[g] + k, in tracker-list
 * (creation) order; the reference emits a group's rows in the REVERSE of that order
 * (sort.py:280), which the host layer restores when it builds the output list. */
typedef struct w2t_sort_result_t {
  double  *out_box;      /* [N,4] x1,y1,width,height after clip_xy (utils.py:40-44)           */
  double  *out_score;    /* [N]   clip(exp(-0.1*mean(P00,P11,P22)),0.2,1) (sort.py:286-287,   */
                         /*       utils.py:49)                                                */
  int32_t *out_birth;    /* [N,2] (group, k): the tracker was the k-th one created at that    */
                         /*       (image, category); see w2t_assign_ids                       */
  int32_t *out_count;    /* [n_img*n_classes] rows emitted (<= det_count)                     */
  int32_t *created;      /* [n_img*n_classes] trackers created at this (image, category)      */
  int32_t *first_img;    /* [n_streams*n_classes] stream-local index of the image where the   */
                         /*   category first appeared (tracker_sort.py:32-33) or -1.          */
                         /*   w2t_sort_step (optional there): the smallest `group` of          */
                         /*   out_birth any live tracker of the sub-stream still refers to     */
                         /*   (INT32_MAX: none) — id bases of older calls can be forgotten     */
  /* optional (may be NULL): final filter state of every sub-stream, for parity tests */
  int32_t *final_count;  /* [n_streams*n_classes] live trackers after the last image          */
  double  *final_state;  /* [n_streams*n_classes, final_cap, 56] x[7] then P[49] row-major,    */
                         /*   predicted one step past the last image                          */
  int32_t  final_cap;
} w2t_sort_result_t;

/* Dense output list of the SORT stage (w2t_sort_finalize): the rows the reference's
 * track.py writes (tracking/utils.py:52-58), in the reference's order for the streams as
 * given.  capacity = rows the arrays can hold (sum of det_count is always enough). */
typedef struct w2t_rows_t {
  double  *box;        /* [capacity,4] x1, y1, width, height                                  */
  double  *score;      /* [capacity]                                                          */
  int64_t *object_id;  /* [capacity] KalmanBoxTracker.id + 1 (sort.py:288)                    */
  int32_t *image;      /* [capacity] image index                                              */
  int32_t *category;   /* [capacity] category_id                                              */
  int64_t *totals;     /* [3] device: trackers created, rows written, next id base            */
                       /*   (= id_base + *id_base_device + created)                           */
  const int64_t *id_base_device; /* optional device int64 added to id_base: chains the ids of  */
                       /*   successive calls (pass the previous call's totals + 2) without a   */
                       /*   host round trip                                                    */
  int64_t  capacity;
  int32_t  image_base; /* added to every image index written: lets a caller finalize a shard   */
                       /*   (chunk of streams with shard-local image indices) of a larger job   */
  int64_t  birth_group_base; /* subtracted from out_birth's group index: set it to the first    */
                       /*   group of the shard when w2t_sort_track ran over the whole job but   */
                       /*   w2t_sort_finalize is called per shard with shard-local arrays       */
  int32_t *compact;    /* optional [capacity,2]: 8 bytes per row INSTEAD of object_id / image /  */
                       /*   category (those may then be NULL), for rows that travel to the host: */
                       /*   [0] object id minus the host-side id_base argument, [1] image * 8 +  */
                       /*   category_id - 1 (n_classes <= 8)                                     */
} w2t_rows_t;

/* layout of the four box columns of w2t_nms_problem_t.rows */
enum {
  W2T_BOX_LTWH = 0,    /* left, top, width, height: rows of convert_submission (ensemble.py:44);  */
                       /*   lxly2cxcy (ensemble.py:19-22) and point_form are applied on the device */
  W2T_BOX_CXCYWH = 1,  /* centre x, centre y, width, height: input of nms_detections (tta.py:8-13) */
  W2T_BOX_XYXY = 2,    /* x1, y1, x2, y2: input of nms (box_utils.py:307)                          */
  W2T_BOX_LTWH_I16 = 3,/* compact rows of 16 bytes: { double score*weight; int16 left, top, width,  */
                       /*   height } — the same values as W2T_BOX_LTWH when every box coordinate   */
                       /*   is an integer in int16 range, which is how detectors write them        */
                       /*   (detnet/data/coco.py:250); 2.5x less host->device traffic              */
  W2T_BOX_LTWH_P64 = 4 /* packed rows of 8 bytes (one uint64): bits 0-16 k = score*weight * 1e5,     */
                       /*   17-29 left + 3072, 30-41 top + 1536, 42-52 width, 53-63 height — the    */
                       /*   same values as W2T_BOX_LTWH when score*weight == (double)k / 1e5 exactly */
                       /*   for an integer k < 2^17 (detectors write scores rounded to 5 decimals,  */
                       /*   detnet/data/coco.py:249, and k / 1e5 is the double that literal parses   */
                       /*   to), left is an integer in [-3072, 5119], top in [-1536, 2559], width and */
                       /*   height in [0, 2047]; the packer checks all of it, bit for bit, and      */
                       /*   falls back to the wider rows otherwise; 5x less host->device traffic    */
};

/* Inputs of the soft-NMS ensemble stage (detnet/ensemble.py:50-64 for every image). */
typedef struct w2t_nms_problem_t {
  int32_t n_groups;
  const int32_t *group_offsets;  /* [n_groups+1] into rows                                    */
  const double  *rows;           /* [N,5] score*weight, left, top, width, height              */
                                 /*   (convert_submission, ensemble.py:44), concatenated over */
                                 /*   submissions in input-file order (tta.py:9-12)           */
  double iou_thresh;             /* --iou-thresh   ensemble.py:96                             */
  double soft_nms_cut;           /* --soft-nms-cut ensemble.py:97                             */
  double min_score;              /* --min-score    ensemble.py:98 (strict > on output, :60)   */
  /* optional hand-over to the SORT stage (NULL score_thr = skip) */
  int32_t n_classes;             /* group g has category (g % n_classes) + 1                  */
  const double *score_thr;       /* [n_classes] --score-threshold, track.py:23-24 (host ptr)  */
  /* general nms() API (detnet/utils/box_utils.py:307); all zero for the ensemble CLI          */
  int32_t box_format;            /* W2T_BOX_* : what rows[.,1:5] hold                         */
  int32_t top_k;                 /* > 0: only the top_k best scored boxes of a group take     */
                                 /*   part (box_utils.py:325-327)                             */
  double conf_thresh;            /* > 0: a box whose running score drops below it is removed  */
                                 /*   and neither output nor suppresses later ones            */
                                 /*   (box_utils.py:379-381)                                  */
  int32_t compute_f32;           /* 1: decay / suppression arithmetic in float32 — the detector */
                                 /*   head's call of nms() with float32 tensors                */
                                 /*   (detnet/nn/modules/detection.py:59-77); rows then hold     */
                                 /*   float32 values widened to float64.  0: float64 (ensemble) */
} w2t_nms_problem_t;

/* Outputs of the ensemble stage.  Row k of group g is at group_offsets[g] + k, in the
 * reference's output order (descending original score, box_utils.py:344-391). */
typedef struct w2t_nms_result_t {
  double  *merged;       /* [N,5] score', cx, cy, w, h = nms_detections() rows (tta.py:19); may be NULL */
  int32_t *src_index;    /* [N] row index (into rows) each output row came from; may be NULL  */
  int32_t *kept_count;   /* [n_groups] rows of merged/src_index that are valid (= group size  */
                         /*   unless top_k / conf_thresh removed boxes); may be NULL          */
  int32_t *ens_count;    /* [n_groups] rows with score' > min_score (ensemble.py:60)          */
  int32_t *ens_box;      /* [N,4] left,top,width,height truncated to int (ensemble.py:62)     */
  double  *ens_score;    /* [N] round(score',5) (ensemble.py:62)                              */
  int32_t *trk_count;    /* [n_groups] rows that also pass read_data_file (utils.py:79-87)    */
  float   *trk_box;      /* [N,4] x, y, x+w, y+h as float32 (utils.py:32-35, tracker_sort.py:45) */
  uint8_t *img_exists;   /* [n_groups/n_classes] image has >= 1 ensemble row; may be NULL     */
} w2t_nms_result_t;

#ifdef __cplusplus
}
#endif
#endif /* W2T_TYPES_H */

"""Host-side packing: the reference's nested dicts <-> the flat arrays of ``include/w2t_types.h``.

The reference keeps detections as ``segment -> camera -> frame -> [dict]``
(``tracking/utils.py:63-96``) and ``image_id -> category_id -> [[score,x,y,w,h]]``
(``detnet/ensemble.py:31-47``).  The CUDA library works on CSR-style arrays;
this module converts both ways and owns the bookkeeping that is order- rather
than arithmetic-related (stream order, category first-appearance order, global
object ids).
"""
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

# tracking/utils.py:11-17
IMAGE_SIZES = {
    'FRONT': [1920, 1280],
    'FRONT_LEFT': [1920, 1280],
    'FRONT_RIGHT': [1920, 1280],
    'SIDE_LEFT': [1920, 886],
    'SIDE_RIGHT': [1920, 886],
}


@dataclass
class PackedTracks:
    """Input of the SORT stage for a set of streams (host arrays)."""
    n_streams: int
    n_classes: int
    streams: List[Tuple[str, str]]            # (segment_id, camera_id) in processing order
    frame_ids: np.ndarray                     # [n_img] int64, sorted inside each stream
    stream_img_offsets: np.ndarray            # [S+1] int32
    det_start: np.ndarray                     # [n_img*NC] int32
    det_count: np.ndarray                     # [n_img*NC] int32
    det_box: np.ndarray                       # [N,4] float32 x1,y1,x2,y2
    cam_wh: np.ndarray                        # [S,2] float64
    img_exists: Optional[np.ndarray] = None   # [n_img] uint8 or None
    class_rank: Optional[np.ndarray] = None   # [S*NC] int32 position in the tracker dict, or None
    n_rows: int = 0
    extras: dict = field(default_factory=dict)

    @property
    def n_img(self):
        return int(self.stream_img_offsets[-1])


def camera_size(camera_id):
    """``IMAGE_SIZES[camera_id]`` — KeyError for an unknown camera like utils.py:21."""
    return IMAGE_SIZES[camera_id]


def pack_predictions(predictions, n_classes, streams=None):
    """``predictions`` = output of ``read_data_file``; one stream per (segment, camera).

    Group g = img * n_classes + (category_id - 1); rows of a group keep their
    order in the frame's list (the order ``MultiClassTrackerSort.track`` sees,
    tracker_sort.py:29-36).
    """
    if streams is None:
        streams = [(seg, cam) for seg in predictions.keys() for cam in predictions[seg]]
    NC = int(n_classes)
    frame_ids, offsets, cam_wh = [], [0], []
    counts, boxes, ranks = [], [], []
    for seg, cam in streams:
        frames = predictions[seg][cam]
        w, h = camera_size(cam)
        cam_wh.append((float(w), float(h)))
        rank = np.full(NC, -1, np.int64)
        seen = 0
        for fid in sorted(frames.keys()):
            per_class = [[] for _ in range(NC)]
            for e in frames[fid]:
                cat = e['category_id']
                if not (1 <= cat <= NC):
                    # the reference indexes iou_thresholds[category_id - 1] (tracker_sort.py:46)
                    raise IndexError("category_id %r outside 1..%d" % (cat, NC))
                b = e['bbox']
                per_class[cat - 1].append((b[0], b[1], b[0] + b[2], b[1] + b[3]))
                if rank[cat - 1] < 0:
                    rank[cat - 1] = seen
                    seen += 1
            frame_ids.append(fid)
            for c in range(NC):
                counts.append(len(per_class[c]))
                boxes.extend(per_class[c])
        for c in range(NC):
            if rank[c] < 0:
                rank[c] = seen
                seen += 1
        ranks.append(rank)
        offsets.append(len(frame_ids))
    det_count = np.asarray(counts, np.int32).reshape(-1)
    det_start = np.zeros_like(det_count)
    if len(det_count):
        det_start[1:] = np.cumsum(det_count[:-1], dtype=np.int64).astype(np.int32)
    det_box = np.asarray(boxes, dtype=np.float64).reshape(-1, 4).astype(np.float32)
    return PackedTracks(
        n_streams=len(streams), n_classes=NC, streams=list(streams),
        frame_ids=np.asarray(frame_ids, np.int64), stream_img_offsets=np.asarray(offsets, np.int32),
        det_start=det_start, det_count=det_count, det_box=det_box,
        cam_wh=np.asarray(cam_wh, np.float64).reshape(-1, 2),
        class_rank=np.asarray(ranks, np.int32).reshape(-1) if ranks else np.zeros(0, np.int32),
        n_rows=int(det_box.shape[0]))


def default_class_rank(first_img, n_streams, n_classes):
    """Rank categories of each stream by (first image they appear in, category id).

    Exact whenever no two categories first appear in the same image; with a tie
    the reference's order is the order inside that frame's detection list —
    ascending category for anything written by ``detnet.ensemble`` (it emits an
    image's rows category by category, ensemble.py:52-63).
    """
    fi = np.asarray(first_img, np.int64).reshape(n_streams, n_classes)
    key = np.where(fi < 0, np.iinfo(np.int64).max // 2, fi) * n_classes + np.arange(n_classes)[None, :]
    order = np.argsort(key, axis=1, kind='stable')
    rank = np.empty_like(order)
    np.put_along_axis(rank, order, np.arange(n_classes)[None, :].repeat(n_streams, 0), axis=1)
    return rank.astype(np.int32).reshape(-1)


def assign_ids(stream_img_offsets, n_classes, det_start, out_count, created, first_img, class_rank,
               out_birth, id_base=0):
    """NumPy statement of ``w2t_assign_ids`` (used to cross-check the C export).

    The reference numbers trackers with one process-global counter
    (``KalmanBoxTracker.count``, sort.py:86,140-141) in creation order: streams in
    order, frames in order, categories in tracker-dict order, new trackers in
    ``unmatched_dets`` order.  ``object_id = id + 1`` (sort.py:288).
    """
    offs = np.asarray(stream_img_offsets, np.int64)
    S, NC = len(offs) - 1, int(n_classes)
    n_img = int(offs[-1])
    created = np.asarray(created, np.int64).reshape(n_img, NC)
    if class_rank is None:
        class_rank = default_class_rank(first_img, S, NC)
    rank = np.asarray(class_rank, np.int64).reshape(S, NC)
    stream_of_img = np.repeat(np.arange(S), np.diff(offs))
    # column k of `ordered` = the category processed k-th in that stream
    cat_at_rank = np.argsort(rank, axis=1, kind='stable')
    ordered = np.take_along_axis(created, cat_at_rank[stream_of_img], axis=1)
    flat = ordered.reshape(-1)
    base_ordered = (np.cumsum(flat) - flat).reshape(n_img, NC) + id_base
    base = np.empty_like(base_ordered)
    np.put_along_axis(base, cat_at_rank[stream_of_img], base_ordered, axis=1)
    base = base.reshape(-1)
    birth = np.asarray(out_birth, np.int64).reshape(-1, 2)
    ids = np.zeros(len(birth), np.int64)        # rows outside [det_start, det_start+out_count) stay 0
    idx, _ = valid_row_index(det_start, out_count)
    ids[idx] = base[birth[idx, 0]] + birth[idx, 1] + 1
    return ids, int(id_base + flat.sum())


def _ranges(starts, counts):
    """Concatenate arange(s, s+c) for every (s, c) without a Python loop."""
    total = int(counts.sum())
    ends = np.cumsum(counts)
    out = np.ones(total, np.int64)
    out[0] = starts[0]
    out[ends[:-1]] = starts[1:] - (starts[:-1] + counts[:-1] - 1)
    return np.cumsum(out)


def valid_row_index(det_start, out_count):
    ds = np.asarray(det_start, np.int64)
    oc = np.asarray(out_count, np.int64)
    nz = np.nonzero(oc)[0]
    if len(nz) == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    return _ranges(ds[nz], oc[nz]), np.repeat(nz, oc[nz])


def unpack_tracks(packed, out_box, out_score, out_count, first_img, ids, class_rank=None):
    """Emit the reference's list of dicts (utils.py:52-58) in the reference's order:
    stream, frame, category in tracker-dict order, rows in ``Sort.update`` order."""
    S, NC = packed.n_streams, packed.n_classes
    rank = packed.class_rank if class_rank is None else class_rank
    if rank is None:
        rank = default_class_rank(first_img, S, NC)
    rank = np.asarray(rank).reshape(S, NC)
    first_img = np.asarray(first_img).reshape(S, NC)
    out = []
    offs = packed.stream_img_offsets
    for s, (seg, cam) in enumerate(packed.streams):
        cats = [int(c) for c in np.argsort(rank[s], kind='stable')]
        for img in range(int(offs[s]), int(offs[s + 1])):
            f_local = img - int(offs[s])
            image_id = '%s/%i/%s' % (seg, packed.frame_ids[img], cam)
            for c in cats:
                if first_img[s, c] < 0 or first_img[s, c] > f_local:
                    continue
                g = img * NC + c
                r0 = int(packed.det_start[g])
                for r in range(r0 + int(out_count[g]) - 1, r0 - 1, -1):   # reference walks the list reversed
                    out.append({
                        'image_id': image_id,
                        'bbox': [out_box[r, 0], out_box[r, 1], out_box[r, 2], out_box[r, 3]],
                        'score': out_score[r],
                        'category_id': c + 1,
                        'object_id': '%i' % int(ids[r]),
                    })
    return out


# ---------------------------------------------------------------------------
# ensemble side
# ---------------------------------------------------------------------------

@dataclass
class PackedGroups:
    """Input of the soft-NMS stage: one group per (image, category)."""
    image_ids: list                 # group g belongs to image_ids[g // len(category_ids)]
    category_ids: list              # iteration order of the reference's category set
    group_offsets: np.ndarray       # [G+1] int32
    rows: np.ndarray                # [N,5] float64: score*weight, left, top, width, height
    max_group: int = 0
    sub_counts: Optional[np.ndarray] = None   # [G,K] int32: rows of each group per input file (fusion only)
    packed: Optional[np.ndarray] = None       # [N] uint64: the same rows in 8 bytes each (:func:`packed_rows`) when exact


def pack_submissions(input_detections, image_ids, category_ids):
    """``input_detections``: one ``convert_submission`` result per input file (ensemble.py:82);
    rows of a group are concatenated in file order, then JSON order (tta.py:9-12)."""
    image_ids = list(image_ids)
    category_ids = list(category_ids)
    offsets = [0]
    chunks = []
    sub_counts = []
    total = 0
    for image_id in image_ids:
        per_file = [det.get(image_id) if hasattr(det, 'get') else det[image_id] for det in input_detections]
        for cat in category_ids:
            for det in per_file:
                rows = None if det is None else det.get(cat)
                sub_counts.append(len(rows) if rows else 0)
                if rows:
                    chunks.append(np.asarray(rows, dtype=np.float64).reshape(-1, 5))
                    total += len(rows)
            offsets.append(total)
    rows = np.concatenate(chunks, axis=0) if chunks else np.zeros((0, 5), np.float64)
    offsets = np.asarray(offsets, np.int64)
    max_group = int(np.diff(offsets).max()) if len(offsets) > 1 else 0
    return PackedGroups(image_ids, category_ids, offsets.astype(np.int32), np.ascontiguousarray(rows), max_group,
                        np.asarray(sub_counts, np.int32).reshape(len(offsets) - 1, max(len(input_detections), 1)))


def image_id_strings(packed):
    """``'%s/%i/%s' % (segment, frame, camera)`` for every image of ``packed`` (utils.py:51)."""
    out = []
    offs = packed.stream_img_offsets
    for s, (seg, cam) in enumerate(packed.streams):
        for img in range(int(offs[s]), int(offs[s + 1])):
            out.append('%s/%i/%s' % (seg, packed.frame_ids[img], cam))
    return out


def rows_to_dicts(packed, res):
    """Dense device-built output list (``w2t_rows_t``: rows_box/score/id/img/cat, already in the
    reference's order) -> the list of dicts ``track_sort`` returns (utils.py:52-58)."""
    ids = image_id_strings(packed)
    box = np.asarray(res["rows_box"]).tolist()
    score = np.asarray(res["rows_score"]).tolist()
    oid = np.asarray(res["rows_id"]).tolist()
    img = np.asarray(res["rows_img"]).tolist()
    cat = np.asarray(res["rows_cat"]).tolist()
    return [{'image_id': ids[i], 'bbox': b, 'score': s, 'category_id': c, 'object_id': '%i' % o}
            for i, b, s, c, o in zip(img, box, score, cat, oid)]


COMPACT_ROW = np.dtype([('score', np.float64), ('box', np.int16, 4)])   # 16 bytes, W2T_BOX_LTWH_I16


def compact_rows(rows):
    """[N,5] float64 rows (score*weight, left, top, width, height) -> 16-byte compact rows
    (``W2T_BOX_LTWH_I16``) when every box coordinate is an integer in int16 range (how detectors
    write them, detnet/data/coco.py:250), else None.  Same values, 2.5x fewer bytes to ship."""
    rows = np.asarray(rows, np.float64).reshape(-1, 5)
    box = rows[:, 1:]
    if len(rows) and not (np.all(box == np.rint(box)) and box.min() >= -32768 and box.max() <= 32767):
        return None
    out = np.empty(len(rows), COMPACT_ROW)
    out['score'] = rows[:, 0]
    out['box'] = box.astype(np.int16)
    return out


def packed_rows(rows):
    """[N,5] float64 rows -> 8-byte packed rows (``W2T_BOX_LTWH_P64``, one uint64 each: 17 bits of
    ``score*weight * 1e5``, 13 bits of left (biased by 3072), 12 bits of top (biased by 1536),
    11 + 11 bits of width / height) when that loses nothing: every ``score*weight`` is bit for bit
    ``k / 1e5`` for an integer ``k < 2**17`` (5-decimal detector scores with unit weights,
    detnet/data/coco.py:249), left is an integer in [-3072, 5119], top in [-1536, 2559], width and
    height in [0, 2047].
    Else None (use :func:`compact_rows` or the float64 rows).  5x fewer bytes to ship."""
    rows = np.asarray(rows, np.float64).reshape(-1, 5)
    box = rows[:, 1:]
    if len(rows) == 0:
        return np.zeros(0, np.uint64)
    if not (np.all(box == np.rint(box)) and box[:, 0].min() >= -3072 and box[:, 0].max() <= 5119
            and box[:, 1].min() >= -1536 and box[:, 1].max() <= 2559
            and box[:, 2:].min() >= 0 and box[:, 2:].max() <= 2047):
        return None
    k = np.rint(rows[:, 0] * 1e5)
    if not (k.min() >= 0 and k.max() < 2 ** 17 and np.array_equal(k / 1e5, rows[:, 0])):
        return None
    out = k.astype(np.uint64)
    out |= (box[:, 0] + 3072).astype(np.uint64) << np.uint64(17)
    out |= (box[:, 1] + 1536).astype(np.uint64) << np.uint64(30)
    out |= box[:, 2].astype(np.uint64) << np.uint64(42)
    out |= box[:, 3].astype(np.uint64) << np.uint64(53)
    return out


# ---------------------------------------------------------------------------
# array paths (native JSON reader -> packed arrays, no per-detection Python objects)
# ---------------------------------------------------------------------------

def pack_detections(dets, score_threshold, n_classes, segment_id=None, segment_block=None):
    """``read_data_file`` (tracking/utils.py:63-96) + :func:`pack_predictions` on the flat arrays of
    ``native_json.load``: same streams (segments, then cameras, in first-appearance order of the
    file), same frames (every frame that occurs in the file, filtered or not), same surviving rows
    in the same order, same category first-appearance ranks.  ``segment_id`` keeps one segment only
    (``track.py --segment-id``); ``segment_block=(rank, world)`` keeps this rank's contiguous block of the
    (remaining) segments in first-appearance order (``sharding.block``): one process per GPU."""
    NC = int(n_classes)
    thr = np.asarray(score_threshold, np.float64)
    # image id -> (segment, frame, camera); ValueError for anything but two slashes, like utils.py:70
    seg_of, cam_of, frame_of = [], [], np.zeros(len(dets.image_ids), np.int64)
    seg_index, cam_names = {}, {}
    for i, image_id in enumerate(dets.image_ids):
        seg, frame, cam = image_id.split('/')
        seg_of.append(seg_index.setdefault(seg, len(seg_index)))
        cam_of.append(cam_names.setdefault(cam, len(cam_names)))
        frame_of[i] = int(frame)
    seg_of, cam_of = np.asarray(seg_of, np.int64), np.asarray(cam_of, np.int64)
    segments, cameras = list(seg_index), list(cam_names)
    n_cam = max(len(cameras), 1)
    # streams in first-appearance order: segment dict order, camera dict order inside the segment.
    # image ids are interned in first-appearance order, so the first image of a (segment, camera)
    # pair has the smallest image index of that pair.
    pair = seg_of * n_cam + cam_of
    upair, first_img = np.unique(pair, return_index=True)
    seg_first = np.full(len(segments), np.iinfo(np.int64).max)
    np.minimum.at(seg_first, upair // n_cam, first_img)
    order = np.lexsort((first_img, seg_first[upair // n_cam]))
    stream_pairs = upair[order]
    if segment_id is not None:
        mask = np.asarray([segments[int(p) // n_cam] == segment_id for p in stream_pairs], bool)
        stream_pairs = stream_pairs[mask] if len(stream_pairs) else stream_pairs
    if segment_block is not None:
        seg_order = list(dict.fromkeys(int(p) // n_cam for p in stream_pairs))   # first-appearance order
        r, w = segment_block
        q, rem = divmod(len(seg_order), int(w))
        lo = r * q + min(r, rem)
        mine = set(seg_order[lo:lo + q + (1 if r < rem else 0)])
        mask = np.asarray([int(p) // n_cam in mine for p in stream_pairs], bool)
        stream_pairs = stream_pairs[mask] if len(stream_pairs) else stream_pairs
    stream_of_pair = {int(p): s for s, p in enumerate(stream_pairs)}
    streams = [(segments[int(p) // n_cam], cameras[int(p) % n_cam]) for p in stream_pairs]
    for _, cam in streams:
        camera_size(cam)                                  # KeyError for an unknown camera (utils.py:21)
    S = len(streams)
    # images of each stream sorted by frame id
    img_stream = np.asarray([stream_of_pair.get(int(p), -1) for p in pair], np.int64)
    used = np.nonzero(img_stream >= 0)[0]
    img_order = used[np.lexsort((frame_of[used], img_stream[used]))]
    new_img = np.full(len(dets.image_ids), -1, np.int64)
    new_img[img_order] = np.arange(len(img_order))
    offsets = np.zeros(S + 1, np.int64)
    np.cumsum(np.bincount(img_stream[used], minlength=S), out=offsets[1:])
    # filters of read_data_file, in its order: box validity first, then the category's threshold
    box = dets.bbox
    valid = ~((box[:, 2] < 1) | (box[:, 3] < 1)) & (new_img[dets.image_index] >= 0)
    cat = dets.category.astype(np.int64)
    bad_cat = valid & ((cat < 1) | (cat > min(NC, len(thr))))
    if bad_cat.any():
        raise IndexError("list index out of range: category_id %d with %d thresholds"
                         % (int(cat[np.nonzero(bad_cat)[0][0]]), min(NC, len(thr))))
    keep = valid.copy()
    keep[valid] = ~(dets.score[valid] < thr[cat[valid] - 1])
    rows = np.nonzero(keep)[0]
    img = new_img[dets.image_index[rows]]
    key = img * NC + (cat[rows] - 1)
    order = np.argsort(key, kind='stable')               # file order inside a group
    b = box[rows][order]
    det_box = np.stack([b[:, 0], b[:, 1], b[:, 0] + b[:, 2], b[:, 1] + b[:, 3]], 1).astype(np.float32)
    n_img = int(offsets[-1])
    counts = np.bincount(key, minlength=n_img * NC).astype(np.int32)
    start = (np.cumsum(counts, dtype=np.int64) - counts).astype(np.int32)
    # position of each category in the stream's tracker dict: first surviving row in (frame, file) order
    pos_first = np.full(S * NC, np.iinfo(np.int64).max, np.int64)
    stream_of_row = np.searchsorted(offsets, img, side='right') - 1
    np.minimum.at(pos_first, stream_of_row * NC + (cat[rows] - 1), img * (len(dets) + 1) + rows)
    rank = np.argsort(np.argsort(pos_first.reshape(S, NC), axis=1, kind='stable'), axis=1, kind='stable')
    cam_wh = np.asarray([camera_size(c) for _, c in streams], np.float64).reshape(-1, 2)
    return PackedTracks(
        n_streams=S, n_classes=NC, streams=streams, frame_ids=frame_of[img_order],
        stream_img_offsets=offsets.astype(np.int32), det_start=start, det_count=counts,
        det_box=np.ascontiguousarray(det_box), cam_wh=cam_wh, img_exists=None,
        class_rank=rank.astype(np.int32).reshape(-1), n_rows=int(det_box.shape[0]))


def pack_detection_files(files, weights, min_score):
    """``load_input_submissions`` (detnet/ensemble.py:78-84) + grouping on the flat arrays of
    ``native_json.load``: one ``Detections`` per input file -> :class:`PackedGroups` with images in
    sorted order and categories ascending, rows of a group in (file, JSON) order.  A row without "score" raises
    ``KeyError('score')`` like ``det['score']`` does there (ensemble.py:38)."""
    if any(len(d.has_score) and not d.has_score.all() for d in files):
        raise KeyError('score')
    category_ids = sorted(set(int(c) for d in files for c in np.unique(d.category)))
    cat_index = np.full((max(category_ids) + 1) if category_ids else 1, -1, np.int64)
    for i, c in enumerate(category_ids):
        cat_index[c] = i
    kept = []
    for d, w in zip(files, weights):
        score = d.score * np.float64(w)
        keep = (d.bbox[:, 2] > 0) & (d.bbox[:, 3] > 0) & (score >= min_score)
        kept.append((keep, score))
    image_ids = sorted(set(d.image_ids[i] for d, (keep, _) in zip(files, kept) for i in np.unique(d.image_index[keep])))
    img_index = {s: i for i, s in enumerate(image_ids)}
    ncat = max(len(category_ids), 1)
    G = len(image_ids) * ncat
    per_file_keys, per_file_rows = [], []
    for d, (keep, score) in zip(files, kept):
        local_to_global = np.asarray([img_index.get(s, -1) for s in d.image_ids], np.int64)
        idx = np.nonzero(keep)[0]
        per_file_keys.append(local_to_global[d.image_index[idx]] * ncat + cat_index[d.category[idx]])
        per_file_rows.append(np.concatenate([score[idx, None], d.bbox[idx]], axis=1))
    keys = np.concatenate(per_file_keys) if per_file_keys else np.zeros(0, np.int64)
    rows = np.concatenate(per_file_rows) if per_file_rows else np.zeros((0, 5))
    order = np.argsort(keys, kind='stable')
    counts = np.bincount(keys, minlength=G) if G else np.zeros(0, np.int64)
    offsets = np.zeros(G + 1, np.int64)
    np.cumsum(counts, out=offsets[1:])
    if G:
        sub_counts = np.stack([np.bincount(k, minlength=G) for k in per_file_keys], axis=1).astype(np.int32)
    else:
        sub_counts = np.zeros((0, max(len(files), 1)), np.int32)
    return PackedGroups(image_ids, category_ids, offsets.astype(np.int32), np.ascontiguousarray(rows[order]),
                        int(counts.max()) if G else 0, sub_counts)


def pack_files(paths, weights, min_score):
    """:func:`pack_detection_files` straight from the files: one native call parses and groups them
    (``w2t_json_group_files``); inputs it does not cover go through ``native_json.load`` + the array packer."""
    from . import native_json
    fast = native_json.group_files(paths, weights, min_score)
    if fast is None:
        return pack_detection_files([native_json.load(f) for f in paths], weights, min_score)
    return PackedGroups(fast.image_ids, fast.category_ids, fast.group_offsets, fast.rows, fast.max_group, fast.sub_counts,
                        fast.packed)


def pack_track_file(path, score_threshold, n_classes, segment_id=None, segment_block=None):
    """:func:`pack_detections` straight from the file: one native call parses, filters and lays the rows out
    (``w2t_json_pack_tracks``); inputs it does not cover go through ``native_json.load`` + :func:`pack_detections`,
    which reproduces the reference's behaviour or its exception."""
    from . import native_json
    fast = native_json.pack_tracks(path, score_threshold, n_classes, segment_id, segment_block)
    if fast is None:
        return pack_detections(native_json.load(path), score_threshold, n_classes, segment_id=segment_id,
                               segment_block=segment_block)
    streams = fast["streams"]
    cam_wh = np.asarray([camera_size(c) for _, c in streams], np.float64).reshape(-1, 2)   # KeyError: utils.py:21
    return PackedTracks(
        n_streams=len(streams), n_classes=int(n_classes), streams=streams, frame_ids=fast["frame_ids"],
        stream_img_offsets=fast["stream_img_offsets"], det_start=fast["det_start"], det_count=fast["det_count"],
        det_box=fast["det_box"], cam_wh=cam_wh, img_exists=None, class_rank=fast["class_rank"],
        n_rows=int(fast["det_box"].shape[0]))

"""CPU restatement ("port") of the reference SORT path.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Follows, function by
function, ``/root/reference/tracking/sort/sort.py:33-296``,
``tracking/sort/tracker_sort.py:10-51`` and ``tracking/utils.py:11-96`` with the
third-party solver / filter restated in ``munkres.py`` / ``kalman.py``.
Every dtype decision is an explicit cast so the port does not depend on the
installed NumPy's promotion rules; ``PROMOTION`` / ``promotion()`` select which
regime the casts restate: "legacy" (NumPy 1.x value-based casting, the reference's
pinned environment, ``/root/reference/environment.yml:7``) or "nep50" (NumPy 2,
what the reference's files do when executed in this container).  Checked
against the reference's own files executed through ``ref_shim`` by
``tests/golden/make_golden.py`` and ``tests/test_oracle.py``.

The per-object Python / small-NumPy-array structure of the reference is kept
on purpose: this port is also the timed CPU baseline (``bench.py``).
"""
import contextlib
import json
import math

import numpy as np

from waymo_2d_tracking_b200 import _abi

from .kalman import KalmanFilter
from .munkres import linear_assignment

f32 = np.float32
f64 = np.float64

PROMOTION = _abi.DEFAULT_PROMOTION     # "legacy" | "nep50"


@contextlib.contextmanager
def promotion(name):
    """Run the port under the given promotion regime (None = leave the current one)."""
    global PROMOTION
    prev = PROMOTION
    if name is not None:
        _abi.promotion_code(name)       # validates
        PROMOTION = name
    try:
        yield
    finally:
        PROMOTION = prev


def _legacy():
    return _abi.promotion_code(PROMOTION) == _abi.W2T_PROMOTION_LEGACY

# tracking/utils.py:11-17
IMAGE_SIZES = {
    'FRONT': [1920, 1280],
    'FRONT_LEFT': [1920, 1280],
    'FRONT_RIGHT': [1920, 1280],
    'SIDE_LEFT': [1920, 886],
    'SIDE_RIGHT': [1920, 886],
}


def iou(det, trk):
    """sort.py:33-47 as numba compiles it for (float32[:], float64[:]):
    the detection's own area is a float32 product, the rest is float64."""
    d0, d1, d2, d3 = f32(det[0]), f32(det[1]), f32(det[2]), f32(det[3])
    t0, t1, t2, t3 = f64(trk[0]), f64(trk[1]), f64(trk[2]), f64(trk[3])
    xx1 = max(f64(d0), t0)
    yy1 = max(f64(d1), t1)
    xx2 = min(f64(d2), t2)
    yy2 = min(f64(d3), t3)
    w = max(f64(0.), xx2 - xx1)
    h = max(f64(0.), yy2 - yy1)
    wh = w * h
    area_det = f32(f32(d2 - d0) * f32(d3 - d1))
    area_trk = (t2 - t0) * (t3 - t1)
    return wh / ((f64(area_det) + area_trk) - wh)


def iou_matrix(dets, trks):
    """sort.py:201-205 (float32 result matrix), vectorised but op-for-op equal to ``iou``."""
    dets = np.asarray(dets, dtype=f32).reshape(-1, dets.shape[-1] if np.ndim(dets) > 1 else 5)
    trks = np.asarray(trks, dtype=f64)
    D, T = len(dets), len(trks)
    out = np.zeros((D, T), dtype=f32)
    if D == 0 or T == 0:
        return out
    d = dets.astype(f64)
    xx1 = np.maximum(d[:, None, 0], trks[None, :, 0])
    yy1 = np.maximum(d[:, None, 1], trks[None, :, 1])
    xx2 = np.minimum(d[:, None, 2], trks[None, :, 2])
    yy2 = np.minimum(d[:, None, 3], trks[None, :, 3])
    w = np.maximum(0., xx2 - xx1)
    h = np.maximum(0., yy2 - yy1)
    wh = w * h
    area_det = ((dets[:, 2] - dets[:, 0]) * (dets[:, 3] - dets[:, 1])).astype(f64)  # f32 product
    area_trk = (trks[:, 2] - trks[:, 0]) * (trks[:, 3] - trks[:, 1])
    with np.errstate(divide='ignore', invalid='ignore'):
        out[:] = wh / ((area_det[:, None] + area_trk[None, :]) - wh)
    return out


def bbox_to_z(bbox):
    """sort.py:50-62 on a float32 row.  NEP 50: every component is float32.  Legacy (NumPy 1.x): ``w``, ``h``
    and ``s = w * h`` are float32 scalars, but float32-scalar (op) python-float is float64, so
    ``x = bbox[0] + w / 2.``, ``y`` and ``r = w / float(h)`` are float64 computations."""
    b0, b1, b2, b3 = f32(bbox[0]), f32(bbox[1]), f32(bbox[2]), f32(bbox[3])
    w = f32(b2 - b0)
    h = f32(b3 - b1)
    if _legacy():
        with np.errstate(divide='ignore', invalid='ignore'):
            x = f64(b0) + f64(w) / f64(2.)
            y = f64(b1) + f64(h) / f64(2.)
            s = f64(f32(w * h))
            r = f64(w) / f64(h)
        return np.array([x, y, s, r], dtype=f64).reshape((4, 1))
    x = f32(b0 + f32(w / f32(2.)))
    y = f32(b1 + f32(h / f32(2.)))
    s = f32(w * h)
    r = f32(w / h)
    return np.array([x, y, s, r], dtype=f32).reshape((4, 1))


def x_to_bbox(x):
    """sort.py:65-75 (float64)."""
    with np.errstate(invalid='ignore', divide='ignore'):
        w = np.sqrt(x[2] * x[3])
        h = x[2] / w
        return np.array([x[0] - w / 2., x[1] - h / 2., x[0] + w / 2., x[1] + h / 2.]).reshape((1, 4))


class BoxTracker:
    """sort.py:78-190 (KalmanBoxTracker)."""
    count = 0

    def __init__(self, bbox):
        kf = KalmanFilter(dim_x=7, dim_z=4)
        F = np.eye(7, dtype=np.int64)
        F[0, 4] = F[1, 5] = F[2, 6] = 1
        kf.F = F                                   # int64, as in the reference
        H = np.zeros((4, 7), dtype=np.int64)
        H[0, 0] = H[1, 1] = H[2, 2] = H[3, 3] = 1
        kf.H = H
        kf.Q[4:, 4:] *= 2 ** 2
        kf.Q[-1, -1] = 5
        kf.Q[0, 0] = 2
        kf.Q[1, 1] = 2
        kf.Q[3, 3] = 25
        kf.R[2:, 2:] *= 10.
        kf.P *= 10.
        kf.P[4:, 4:] *= 1000.
        kf.x[:4] = bbox_to_z(bbox)
        self.kf = kf
        self.id = BoxTracker.count
        BoxTracker.count += 1
        self.age = 0
        self.hits = 0
        self.hit_streak = 0
        self.time_since_update = 0

    def update(self, bbox):
        self.time_since_update = 0
        self.hits += 1
        self.hit_streak += 1
        self.kf.update(bbox_to_z(bbox))

    def predict(self):
        kf = self.kf
        if (kf.x[6] + kf.x[2]) <= 0:
            kf.x[6] *= 0.0
        kf.predict()
        self.age += 1
        if self.time_since_update > 0:
            self.hit_streak = 0
        self.time_since_update += 1
        return x_to_bbox(kf.x)

    def get_state(self):
        return x_to_bbox(self.kf.x)

    def get_error(self):
        P = self.kf.P
        return np.mean([P[0, 0], P[1, 1], P[2, 2]])


def associate(dets, trks, iou_threshold=0.3, return_iou=False):
    """sort.py:193-230."""
    if len(trks) == 0:
        res = (np.empty((0, 2), dtype=int), np.arange(len(dets)), np.empty((0, 5), dtype=int))
        return res + (np.zeros((len(dets), 0), f32),) if return_iou else res
    M = iou_matrix(dets, trks)
    pairs = linear_assignment(-M)
    used_d = set(pairs[:, 0].tolist())
    used_t = set(pairs[:, 1].tolist())
    free_d = [d for d in range(len(dets)) if d not in used_d]
    free_t = [t for t in range(len(trks)) if t not in used_t]
    # sort.py:220: NEP 50 - the python float adopts float32; legacy - float32 scalar vs python float compares in float64
    thr = f64(iou_threshold) if _legacy() else f32(iou_threshold)
    keep = []
    for d, t in pairs:
        if (f64(M[d, t]) if _legacy() else M[d, t]) < thr:
            free_d.append(int(d))
            free_t.append(int(t))
        else:
            keep.append((int(d), int(t)))
    matches = np.array(keep, dtype=int).reshape(-1, 2)
    res = (matches, np.array(free_d, dtype=int), np.array(free_t, dtype=int))
    return res + (M,) if return_iou else res


class Sort:
    """sort.py:233-296."""

    def __init__(self, max_age=1, min_hits=3):
        self.max_age = max_age
        self.min_hits = min_hits
        self.trackers = []
        self.frame_count = 0
        self.confidence_factor = 0.1

    def update(self, dets, iou_threshold):
        dets = np.asarray(dets, dtype=f32)
        if dets.ndim == 1:
            dets = dets.reshape(0, 5) if dets.size == 0 else dets.reshape(1, -1)
        self.frame_count += 1
        boxes = []
        alive = []
        for trk in self.trackers:
            pos = trk.predict()[0]
            if np.any(np.isnan(pos)):
                continue
            assert np.all(np.isfinite(pos)), "inf tracker box: unreachable with finite inputs (SURVEY §7 hard part 7)"
            boxes.append(pos)
            alive.append(trk)
        self.trackers = alive
        trks = np.array(boxes, dtype=f64).reshape(-1, 4)
        matched, free_d, free_t = associate(dets, trks, iou_threshold)
        free_t = set(int(t) for t in free_t)
        det_of = {int(t): int(d) for d, t in matched}
        for t, trk in enumerate(self.trackers):
            if t not in free_t:
                trk.update(dets[det_of[t], :])
        for d in free_d:
            self.trackers.append(BoxTracker(dets[int(d), :]))
        out = []
        i = len(self.trackers)
        for trk in reversed(self.trackers):
            box = trk.get_state()[0]
            if trk.time_since_update < 1 and (trk.hit_streak >= self.min_hits or self.frame_count <= self.min_hits):
                conf = np.exp(-trk.get_error() * self.confidence_factor)
                out.append(np.concatenate((box, [trk.id + 1, conf])).reshape(1, -1))
            i -= 1
            if trk.time_since_update > self.max_age:
                self.trackers.pop(i)
        if out:
            return np.concatenate(out)
        return np.empty((0, 6))


class MultiClassTracker:
    """tracker_sort.py:10-51."""

    def __init__(self, max_age=1, min_hits=0):
        self.max_age = max_age
        self.min_hits = min_hits
        self.trackers = {}

    def track(self, detected_objects, iou_thresholds):
        by_class = {}
        for obj in detected_objects:
            cls = obj[5]
            if cls not in self.trackers:
                self.trackers[cls] = Sort(max_age=self.max_age, min_hits=self.min_hits)
            by_class.setdefault(cls, []).append(obj[:5])
        result = {}
        for cls, trk in self.trackers.items():
            rows = by_class.get(cls)
            arr = np.array(rows, dtype=f32) if rows is not None else np.array([], dtype=f32)
            result[cls] = trk.update(arr, iou_threshold=iou_thresholds[cls - 1])
        return result


def clip_xy(camera_id, x, y):
    """utils.py:20-22."""
    w, h = IMAGE_SIZES[camera_id]
    return min(max(x, f64(0)), f64(w)), min(max(y, f64(0)), f64(h))


def track_sort(predictions, segment_id, camera_id, iou_thresholds, max_age, min_hits):
    """utils.py:25-60."""
    frames = predictions[segment_id][camera_id]
    out = []
    tracker = MultiClassTracker(max_age=max_age, min_hits=min_hits)
    for frame_id in sorted(frames.keys()):
        rows = [[e['bbox'][0], e['bbox'][1], e['bbox'][0] + e['bbox'][2], e['bbox'][1] + e['bbox'][3],
                 e['score'], e['category_id']] for e in frames[frame_id]]
        for category_id, tracked in tracker.track(rows, iou_thresholds).items():
            for obj in tracked:
                x1, y1 = clip_xy(camera_id, obj[0], obj[1])
                x2, y2 = clip_xy(camera_id, obj[2], obj[3])
                width, height = x2 - x1, y2 - y1
                if width < 1 or height < 1:
                    continue
                conf = min(max(obj[5], f64(0.2)), f64(1.0))
                out.append({
                    'image_id': '%s/%i/%s' % (segment_id, frame_id, camera_id),
                    'bbox': [x1, y1, width, height],
                    'score': conf,
                    'category_id': category_id,
                    'object_id': '%i' % int(obj[4]),
                })
    return out


def group_entries(raw_entries, score_threshold):
    """utils.py:63-96 minus the file read: segment -> camera -> int(frame) -> [entry]."""
    if isinstance(raw_entries, dict) and 'annotations' in raw_entries:
        raw_entries = raw_entries['annotations']
    entries = {}
    for entry in raw_entries:
        segment_id, frame_id, camera_id = entry['image_id'].split('/')
        frames = entries.setdefault(segment_id, {}).setdefault(camera_id, {})
        bucket = frames.setdefault(int(frame_id), [])
        bbox = entry['bbox']
        if bbox[2] < 1 or bbox[3] < 1:
            continue
        category_id = entry['category_id']
        score = entry.get('score', 1.0)
        if score < score_threshold[category_id - 1]:
            continue
        kept = {'bbox': bbox, 'score': score, 'category_id': category_id}
        if 'object_id' in entry:
            kept['object_id'] = entry['object_id']
        bucket.append(kept)
    return entries


def read_data_file(file_name, score_threshold):
    with open(file_name) as fp:
        return group_entries(json.load(fp), score_threshold)


def track_all(predictions, iou_thresholds, max_age, min_hits, reset_ids=True, promotion=None):
    """tracking/track.py:42-47 — the reference's own timed region."""
    if reset_ids:
        BoxTracker.count = 0
    out = []
    with globals()["promotion"](promotion):
        for segment_id in predictions.keys():
            for camera_id in predictions[segment_id]:
                out += track_sort(predictions, segment_id, camera_id, iou_thresholds, max_age, min_hits)
    return out

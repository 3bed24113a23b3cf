"""Drop-in for the tracker part of ``tracking/sort/sort.py`` (:33-296): ``iou``,
``convert_bbox_to_z``, ``convert_x_to_bbox``, ``KalmanBoxTracker``,
``associate_detections_to_trackers`` and ``Sort``.

Every arithmetic step runs on the device through the building-block exports of ``libw2t.so``
(``w2t_iou_matrix``, ``w2t_linear_assignment``, ``w2t_kf_init/predict/update``, ``w2t_bbox_to_z``,
``w2t_x_to_bbox``): the same device functions the persistent SORT kernel is made of.  The list
bookkeeping of ``Sort.update`` (who is matched, who is new, who is dropped) stays on the host
exactly as the reference orders it.  This frame-by-frame API exists for compatibility — each
call costs a handful of small launches; whole streams belong to ``tracking.utils.track_all``.
There is no CPU fallback.

NumPy promotion (``_abi.DEFAULT_PROMOTION``): by default the reference's pinned NumPy 1.x
environment — in ``convert_bbox_to_z`` of a float32 row x, y and r are float64 computations and only s is a
float32 product, and the IoU threshold is compared in float64; ``W2T_PROMOTION=nep50`` selects NumPy 2
semantics (all four components float32, threshold compared in float32).  The solver is the
scikit-learn 0.22.2 Munkres emulation (``csrc/munkres.cuh``), the filter the filterpy
``KalmanFilter`` restatement (``csrc/kalman.cuh``).
"""
from types import SimpleNamespace

import numpy as np

try:
    from ... import runtime
except ImportError:                     # imported as top-level `tracking` / `sort` (reference-style sys.path)
    import os as _os
    import sys as _sys
    _sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))))
    from waymo_2d_tracking_b200 import runtime


def iou(bb_test, bb_gt):
    """IoU of a detection row (rounded to float32, its area is a float32 product) and a float64
    tracker box, as numba compiles sort.py:33-47 for the call at sort.py:205; float32 result."""
    return float(runtime.iou_matrix(np.asarray(bb_test, np.float32)[:4], np.asarray(bb_gt, np.float64)[:4])[0, 0])


def iou_batch(bb_test, bb_gt):
    """[D,>=4] detections x [T,>=4] trackers -> float32 [D,T] (the matrix of sort.py:201-205).
    The name the upstream SORT project uses; same semantics as :func:`iou` pair by pair."""
    d = np.asarray(bb_test, np.float32).reshape(-1, np.shape(bb_test)[-1])[:, :4]
    t = np.asarray(bb_gt, np.float64).reshape(-1, np.shape(bb_gt)[-1])[:, :4]
    return runtime.iou_matrix(d, t)


def convert_bbox_to_z(bbox):
    """[x1,y1,x2,y2] -> [x,y,s,r] as a (4,1) array (sort.py:50-62 on a float32 row): float64 under legacy
    promotion, float32 under NEP 50."""
    return runtime.bbox_to_z(np.asarray(bbox, np.float32)[:4]).reshape(4, 1)


def convert_x_to_bbox(x, score=None):
    """[x,y,s,r,...] -> [x1,y1,x2,y2] as (1,4), or (1,5) with ``score`` appended (sort.py:65-75)."""
    box = runtime.x_to_bbox(np.asarray(x, np.float64).reshape(-1)[:4])
    if score is None:
        return box.reshape(1, 4)
    return np.concatenate([box.reshape(4), [score]]).reshape(1, 5)


class KalmanBoxTracker(object):
    """State of one tracked object (sort.py:78-190): 7-state constant-velocity filter on
    (x, y, s, r); ``kf.x`` is (7,1), ``kf.P`` (7,7) like filterpy's attributes."""

    count = 0   # process-global id counter (sort.py:86)

    def __init__(self, bbox):
        x, P = runtime.kf_init(np.asarray(bbox, np.float32)[:4])
        self.kf = SimpleNamespace(x=x.reshape(7, 1), P=P.reshape(7, 7))
        self.id = KalmanBoxTracker.count
        KalmanBoxTracker.count += 1
        self.age = 0
        self.hits = 0
        self.hit_streak = 0
        self.time_since_update = 0
        self.history = []

    def _after_update(self, x, P):
        self.time_since_update = 0
        self.history = []
        self.hits += 1
        self.hit_streak += 1
        self.kf.x, self.kf.P = x.reshape(7, 1), P.reshape(7, 7)

    def _after_predict(self, x, P, box):
        self.kf.x, self.kf.P = x.reshape(7, 1), P.reshape(7, 7)
        self.age += 1
        if self.time_since_update > 0:
            self.hit_streak = 0
        self.time_since_update += 1
        self.history.append(box.reshape(1, 4))
        return self.history[-1]

    def update(self, bbox):
        """Kalman update with an observed box (sort.py:153-164)."""
        x, P, _ = runtime.kf_update(self.kf.x, self.kf.P, np.asarray(bbox, np.float32)[:4])
        self._after_update(x[0], P[0])

    def predict(self):
        """Advance one frame and return the predicted box (sort.py:166-178); the ``s + ds <= 0``
        guard of :170-171 is applied on the device."""
        x, P, boxes = runtime.kf_predict(self.kf.x, self.kf.P)
        return self._after_predict(x[0], P[0], boxes[0])

    def get_state(self):
        return convert_x_to_bbox(self.kf.x)

    def get_error(self):
        """mean(P00, P11, P22), sort.py:186-190."""
        return np.mean([self.kf.P[i, i] for i in range(3)])


def associate_detections_to_trackers(detections, trackers, iou_threshold=0.3):
    """sort.py:193-230 -> (matches [k,2], unmatched detection indices, unmatched tracker indices).
    Assigned pairs whose IoU is below the threshold are un-matched again and appended AFTER the
    never-assigned ones, which fixes the order in which new trackers get their ids."""
    if len(trackers) == 0:
        return np.empty((0, 2), dtype=int), np.arange(len(detections)), np.empty((0, 5), dtype=int)
    iou_matrix = iou_batch(detections, trackers) if len(detections) else np.zeros((0, len(trackers)), np.float32)
    matched_indices = runtime.linear_assignment(-iou_matrix)
    unmatched_detections = [d for d in range(len(detections)) if d not in matched_indices[:, 0]]
    unmatched_trackers = [t for t in range(len(trackers)) if t not in matched_indices[:, 1]]
    # sort.py:220: NumPy 1.x compares the float32 entry with the python float in float64; under NEP 50 the python
    # float adopts the matrix dtype
    legacy = runtime.promotion_code() == 0
    thr = np.float64(iou_threshold) if legacy else np.float32(iou_threshold)
    matches = []
    for d, t in matched_indices:
        if (np.float64(iou_matrix[d, t]) if legacy else iou_matrix[d, t]) < thr:
            unmatched_detections.append(d)
            unmatched_trackers.append(t)
        else:
            matches.append((d, t))
    matches = np.asarray(matches, dtype=int).reshape(-1, 2)
    return matches, np.array(unmatched_detections), np.array(unmatched_trackers)


class Sort(object):
    """Single-category SORT (sort.py:233-296).  ``update`` must be called once per frame, with an
    empty array when there are no detections."""

    def __init__(self, max_age=1, min_hits=3):
        self.max_age = max_age
        self.min_hits = min_hits
        self.trackers = []
        self.frame_count = 0
        self.confidence_factor = 0.1

    def update(self, dets, iou_threshold):
        """``dets``: float32 [D,5] = x1,y1,x2,y2,score -> ndarray [m,6] = x1,y1,x2,y2,id+1,confidence."""
        dets = np.asarray(dets, dtype=np.float32).reshape(-1, 5) if np.size(dets) else np.zeros((0, 5), np.float32)
        self.frame_count += 1
        # predict every live tracker in one launch (sort.py:255-262)
        trks = np.zeros((len(self.trackers), 4))
        if self.trackers:
            x, P, boxes = runtime.kf_predict(np.stack([t.kf.x.reshape(7) for t in self.trackers]),
                                             np.stack([t.kf.P for t in self.trackers]))
            for i, trk in enumerate(self.trackers):
                trks[i] = trk._after_predict(x[i], P[i], boxes[i])[0]
        nan_rows = [i for i in range(len(trks)) if np.any(np.isnan(trks[i]))]
        if not np.all(np.isfinite(trks[[i for i in range(len(trks)) if i not in nan_rows]])):
            # an infinite box would be compressed out of `trks` but kept in the tracker list, shifting
            # every later index (sort.py:261-265); unreachable with finite inputs, refused here
            raise FloatingPointError("a tracker box became infinite")
        trks = np.delete(trks, nan_rows, axis=0)
        for i in reversed(nan_rows):
            self.trackers.pop(i)
        matched, unmatched_dets, unmatched_trks = associate_detections_to_trackers(dets, trks, iou_threshold)

        # update the matched trackers in one launch (sort.py:268-273)
        if len(matched):
            who = [self.trackers[t] for t in matched[:, 1]]
            x, P, _ = runtime.kf_update(np.stack([t.kf.x.reshape(7) for t in who]),
                                        np.stack([t.kf.P for t in who]), dets[matched[:, 0], :4])
            for k, trk in enumerate(who):
                trk._after_update(x[k], P[k])
        # new trackers in unmatched order (sort.py:276-278)
        for i in unmatched_dets:
            self.trackers.append(KalmanBoxTracker(dets[int(i), :]))

        ret = []
        if self.trackers:
            states = runtime.x_to_bbox(np.stack([t.kf.x.reshape(7) for t in self.trackers]))
        i = len(self.trackers)
        for trk in reversed(self.trackers):
            i -= 1
            if trk.time_since_update < 1 and (trk.hit_streak >= self.min_hits or self.frame_count <= self.min_hits):
                confidence = np.exp(-trk.get_error() * self.confidence_factor)
                ret.append(np.concatenate((states[i], [trk.id + 1, confidence])).reshape(1, -1))
            if trk.time_since_update > self.max_age:
                self.trackers.pop(i)
        if ret:
            return np.concatenate(ret)
        return np.empty((0, 6))

"""Drop-in for ``tracking/sort/tracker_sort.py``: one :class:`Sort` per category, created when the
category first appears; afterwards every category is stepped on every frame, with an empty
detection array when it has none (``tracker_sort.py:22-51``).

This is the stateful frame-by-frame API (kept for compatibility; each call launches the
building-block kernels of ``csrc/units.cu``).  The performance path is the batch one:
``tracking.utils.track_all`` / ``track_sort`` run whole streams in the persistent kernel.
"""
import numpy as np

from .sort import Sort


class MultiClassTrackerSort(object):
    def __init__(self, max_age=1, min_hits=0):
        """max_age: frames a track may go unmatched before it is removed;
        min_hits: consecutive matches needed before a track is reported."""
        self.max_age = max_age
        self.min_hits = min_hits
        self.trackers = {}

    def track(self, detected_objects, iou_thresholds):
        """``detected_objects``: ``[[x1, y1, x2, y2, confidence, class], ...]`` ->
        ``{class: ndarray[m, 6] = x1, y1, x2, y2, object id, confidence}``."""
        per_class = {}
        for row in detected_objects:
            name = row[5]
            if name not in self.trackers:
                self.trackers[name] = Sort(max_age=self.max_age, min_hits=self.min_hits)
            per_class.setdefault(name, []).append(row[:5])
        tracked = {}
        for name, tracker in self.trackers.items():
            dets = np.array(per_class.get(name, []), dtype=np.float32)
            tracked[name] = tracker.update(dets, iou_threshold=iou_thresholds[name - 1])
        return tracked


class DeviceMultiClassTrackerSort(object):
    """:class:`MultiClassTrackerSort` with device-resident state: same ``track()`` call and result, but the
    filters, tracker lists and counters of every category stay on the GPU between frames and a frame is
    ONE launch of the persistent tracker kernel over one image (``runtime.SortStepper`` /
    ``w2t_sort_step``) instead of a handful of building-block launches per category.  The IoU thresholds
    are part of the device problem, so they are fixed by the first ``track()`` call.  Object ids continue
    ``KalmanBoxTracker.count`` of :mod:`.sort` (read at construction, written back after every frame)."""

    def __init__(self, max_age=1, min_hits=0, track_cap=256, det_cap=256):
        self.max_age, self.min_hits = max_age, min_hits
        self.track_cap, self.det_cap = track_cap, det_cap
        self._stepper = None

    @property
    def trackers(self):
        """Categories that have a ``Sort`` object, in creation order (keys of the reference's dict)."""
        return [] if self._stepper is None else [c + 1 for c in self._stepper.class_order[0]]

    def track(self, detected_objects, iou_thresholds):
        from . import sort as _sort
        from ... import runtime
        if self._stepper is None:
            self._stepper = runtime.SortStepper(list(iou_thresholds), self.max_age, self.min_hits, 1,
                                                self.track_cap, self.det_cap, id_base=_sort.KalmanBoxTracker.count)
        elif list(iou_thresholds) != self._stepper.iou_thresholds:
            raise ValueError("iou_thresholds are fixed by the first call of a DeviceMultiClassTrackerSort")
        rows = np.asarray([r[:5] for r in detected_objects], np.float64).reshape(-1, 5)
        cats = np.asarray([int(r[5]) - 1 for r in detected_objects], np.int64)
        if len(cats) and cats.min() < 0:
            raise IndexError("category ids are 1-based")
        self._stepper.count = _sort.KalmanBoxTracker.count
        out = self._stepper.step([(rows, cats)])[0]
        _sort.KalmanBoxTracker.count = self._stepper.count
        return {c + 1: v for c, v in out.items()}

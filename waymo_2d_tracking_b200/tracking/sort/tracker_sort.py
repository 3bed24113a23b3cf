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

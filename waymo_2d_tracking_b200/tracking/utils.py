"""Drop-in for ``tracking/utils.py``: ``IMAGE_SIZES``, ``clip_xy``, ``read_data_file``, ``track_sort``.

``track_sort`` keeps the reference's signature (one (segment, camera) stream per call,
``utils.py:25-60``) and ``track_all`` is the batch form the CLI uses (every stream of the input in
one launch of the persistent SORT kernel, ``csrc/sort_kernel.cuh``).  Both return the reference's
list of dicts in the reference's order, with object ids drawn from the same process-global counter
(``KalmanBoxTracker.count``, ``sort.py:86``).  No CPU fallback.
"""
import json
import warnings

import numpy as np

try:                                    # package import (python -m waymo_2d_tracking_b200.tracking.track)
    from .. import native_json, packing, runtime, sharding
    from .sort import sort as _sort
except ImportError:                     # `python track.py` inside tracking/, or top-level `tracking` package
    import os as _os
    import sys as _sys
    _sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
    from waymo_2d_tracking_b200 import native_json, packing, runtime, sharding
    from waymo_2d_tracking_b200.tracking.sort import sort as _sort

warnings.simplefilter(action='ignore', category=FutureWarning)

IMAGE_SIZES = packing.IMAGE_SIZES


def clip_xy(camera_id, x, y):
    """Clip a point to the camera image (utils.py:20-22); ``KeyError`` for an unknown camera."""
    width, height = IMAGE_SIZES[camera_id]
    return np.clip(x, a_min=0, a_max=width), np.clip(y, a_min=0, a_max=height)


def read_data_file(file_name, score_threshold):
    """Submission / annotation JSON -> ``{segment: {camera: {frame: [entry, ...]}}}`` (utils.py:63-96).

    A frame's list is created as soon as the frame is seen, before the validity and score filters:
    a frame whose detections are all filtered out still steps the trackers (with no detections),
    a frame absent from the file does not."""
    with open(file_name) as fp:
        raw = json.load(fp)
    if 'annotations' in raw:
        raw = raw['annotations']
    entries = {}
    for item in raw:
        segment_id, frame_id, camera_id = item['image_id'].split('/')
        frame = entries.setdefault(segment_id, {}).setdefault(camera_id, {}).setdefault(int(frame_id), [])
        bbox = item['bbox']
        if bbox[2] < 1 or bbox[3] < 1:
            continue
        category_id = item['category_id']
        score = item.get('score', 1.0)          # ground-truth annotations carry no score
        if score < score_threshold[category_id - 1]:
            continue
        kept = {'bbox': bbox, 'score': score, 'category_id': category_id}
        if 'object_id' in item:
            kept['object_id'] = item['object_id']
        frame.append(kept)
    return entries


def _n_classes(predictions, streams, iou_thresholds):
    """Categories are 1..len(iou_thresholds); anything else fails like ``iou_thresholds[category-1]``
    does in the reference (tracker_sort.py:46) — before any kernel is launched."""
    n = len(iou_thresholds)
    for seg, cam in streams:
        for entries in predictions[seg][cam].values():
            for e in entries:
                if not (1 <= e['category_id'] <= n):
                    raise IndexError("list index out of range: category_id %r with %d IoU thresholds"
                                     % (e['category_id'], n))
    if n > runtime._abi.W2T_MAX_CLASSES:
        raise ValueError("at most %d categories are supported" % runtime._abi.W2T_MAX_CLASSES)
    return n


def track_streams(predictions, streams, iou_thresholds, max_age, min_hits):
    """Track the given (segment, camera) streams in one launch; ids continue the global counter."""
    streams = list(streams)
    for _, cam in streams:
        IMAGE_SIZES[cam]                        # KeyError like clip_xy (utils.py:21)
    if not streams:
        return []
    nc = _n_classes(predictions, streams, iou_thresholds)
    packed = packing.pack_predictions(predictions, nc, streams)
    res = runtime.sort_track(packed, list(iou_thresholds)[:nc], max_age, min_hits,
                             id_base=_sort.KalmanBoxTracker.count, raw=False)
    _sort.KalmanBoxTracker.count = res["id_next"]
    return packing.rows_to_dicts(packed, res)


def track_packed(packed, iou_thresholds, max_age, min_hits):
    """Track already packed streams (``packing.pack_detections`` / ``pack_predictions``); returns the
    dense result arrays (``rows_box/score/id/img/cat``) and advances the global id counter."""
    if packed.n_classes > runtime._abi.W2T_MAX_CLASSES:
        raise ValueError("at most %d categories are supported" % runtime._abi.W2T_MAX_CLASSES)
    res = runtime.sort_track(packed, list(iou_thresholds)[:packed.n_classes], max_age, min_hits,
                             id_base=_sort.KalmanBoxTracker.count, raw=False)
    _sort.KalmanBoxTracker.count = res["id_next"]
    return res


def track_sort(predictions, segment_id, camera_id, iou_thresholds, max_age, min_hits):
    """One stream (utils.py:25-60)."""
    return track_streams(predictions, [(segment_id, camera_id)], iou_thresholds, max_age, min_hits)


def track_all(predictions, iou_thresholds, max_age, min_hits):
    """Every stream in the reference's processing order (track.py:43-47): segments, then cameras,
    both in first-appearance order of the input file."""
    streams = [(seg, cam) for seg in predictions.keys() for cam in predictions[seg]]
    return track_streams(predictions, streams, iou_thresholds, max_age, min_hits)

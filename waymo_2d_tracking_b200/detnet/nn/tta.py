"""Drop-in for ``nms_detections`` of ``detnet/nn/tta.py:8-19`` (the merge function behind
``python -m detnet.ensemble -m soft_nms`` / ``-m nms``).

Input and output are what the reference's function takes and returns — a list of float64
``[n_k, 5]`` arrays ``[score, cx, cy, w, h]`` (one per submission) in, one ``[n, 5]`` array in
descending original-score order out — but point_form -> nms -> center_size run as ONE launch of
the CUDA kernel (``w2t_softnms_groups`` / ``w2t_hardnms_groups`` with ``W2T_BOX_CXCYWH`` rows).
The test-time-augmentation operator tree of the reference's ``tta.py`` (:69-346) wraps a detector
forward pass and is out of scope.
"""
import numpy as np

try:
    from ... import _abi, runtime
except ImportError:                     # imported as top-level `detnet` (PYTHONPATH=.../waymo_2d_tracking_b200)
    import os as _os
    import sys as _sys
    _sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))))
    from waymo_2d_tracking_b200 import _abi, runtime


def nms_detections(detections, iou_thresh=0.5, soft=False, soft_nms_cut=1):
    """tta.py:8-19.  Rows are stacked in list order, then row order (tta.py:9-12)."""
    stacked = [np.asarray(d, dtype=np.float64).reshape(-1, 5) for d in detections]
    rows = np.ascontiguousarray(np.vstack(stacked)) if stacked else np.zeros((0, 5))
    n = len(rows)
    if n == 0:
        return np.zeros((0, 5), np.float64)
    offsets = np.array([0, n], np.int32)
    res = runtime.softnms_groups(offsets, rows, float(iou_thresh), float(soft_nms_cut), -np.inf, max_group=n,
                                 box_format=_abi.W2T_BOX_CXCYWH, want_ensemble=False, hard=not soft)
    return res["merged"][:int(res["kept_count"][0])].copy()


def merge_detections(detections, nms_thresh=0.5):
    """tta.py:22-66 (confidence-weighted box fusion, the CLI's ``-m weighted_fusion``)."""
    from . import fusion
    return fusion.merge_detections(detections, nms_thresh)

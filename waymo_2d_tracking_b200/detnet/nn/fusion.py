"""Host side of the weighted-box-fusion merge (``merge_detections``, ``detnet/nn/tta.py:22-66``;
the default ``-m weighted_fusion`` of the ensemble CLI).  The arithmetic runs in
``csrc/fusion.cu`` (``w2t_fusion_groups``); no CPU fallback."""
import numpy as np

try:
    from ... import _abi, runtime
except ImportError:                     # imported as top-level `detnet` (PYTHONPATH=.../waymo_2d_tracking_b200)
    import os as _os
    import sys as _sys
    _sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))))
    from waymo_2d_tracking_b200 import _abi, runtime


def merge_detections(detections, nms_thresh=0.5):
    """tta.py:22-66: ``detections`` = one float64 ``[n_k, 5]`` array ``[score, cx, cy, w, h]`` per
    submission; returns the fused ``[m, 5]`` rows in result-list order.  (The reference also
    overwrites its input arrays with the weighted rows as a side effect; this one leaves them.)"""
    stacked = [np.asarray(d, dtype=np.float64).reshape(-1, 5) for d in detections]
    if not stacked:
        raise IndexError("list index out of range")     # detections[0] in the reference
    rows = np.ascontiguousarray(np.vstack(stacked))
    n = len(rows)
    if n == 0:
        return np.zeros((0, 5), np.float64)
    res = runtime.fusion_groups(np.array([0, n], np.int32), rows, np.array([[len(d) for d in stacked]], np.int32),
                                float(nms_thresh), -np.inf, max_group=n, box_format=_abi.W2T_BOX_CXCYWH)
    return res["merged"][:int(res["kept_count"][0])].copy()


def fuse_groups(groups, nms_thresh, min_score):
    """All (image, category) groups of a ``packing.PackedGroups`` in one launch."""
    if groups.sub_counts is None:
        raise ValueError("weighted fusion needs PackedGroups.sub_counts (rows per input file and group)")
    return runtime.fusion_groups(groups.group_offsets, groups.rows, groups.sub_counts, float(nms_thresh),
                                 float(min_score), max_group=groups.max_group, box_format=_abi.W2T_BOX_LTWH)

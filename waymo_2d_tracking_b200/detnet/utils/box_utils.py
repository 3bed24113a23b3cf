"""Drop-in for the three functions of ``detnet/utils/box_utils.py`` that lie on the ensemble path:
``point_form`` (:8-35), ``center_size`` (:57-69) and ``nms`` (:307-395).

``nms`` keeps the reference's signature and return convention — ``(keep, scores)`` with ``keep``
the indices of the surviving boxes in descending original-score order — but the work is one
launch of the CUDA soft-NMS kernel (``csrc/softnms.cu``) through ``w2t_softnms_groups``;
``top_k`` and ``conf_thresh`` are honoured on the device.  The hard-NMS branch (``soft=False``,
the reference delegates it to ``torchvision.ops.nms``) is the ``HARD`` instantiation of the same kernel
(``w2t_hardnms_groups``).  float64 tensors (the ensemble path) are computed in float64, float32 tensors (the
detector head's call, detnet/nn/modules/detection.py:59-77) in float32, like torch does.  ``bbox_vote``
(:401-430) runs ``w2t_bbox_vote``.
There is no CPU fallback: without a CUDA device ``nms`` raises.

Order of equal scores: the reference's ``scores.sort(0)`` is unstable, so its order among ties is
implementation-defined; this implementation uses the canonical rule of SURVEY.md §8c (a stable
ascending sort consumed from the end: of two equal scores the later box is processed first).
"""
import numpy as np
import torch

try:
    from ... import _abi, runtime
except ImportError:                     # imported as top-level `detnet` (PYTHONPATH=.../waymo_2d_tracking_b200)
    import os as _os
    import sys as _sys
    _sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))))
    from waymo_2d_tracking_b200 import _abi, runtime


def point_form(boxes):
    """(cx, cy, w, h) -> (xmin, ymin, xmax, ymax); box_utils.py:32-35 (axis-aligned boxes only)."""
    if boxes.size(-1) != 4:
        raise NotImplementedError("rotated boxes (last dimension 5) are outside the ensemble path")
    half = boxes[..., 2:4] * 0.5
    centre = boxes[..., :2]
    return torch.cat((centre - half, centre + half), -1)


def center_size(boxes):
    """(xmin, ymin, xmax, ymax) -> (cx, cy, w, h); box_utils.py:57-69."""
    lo, hi = boxes[:, :2], boxes[:, 2:4]
    return torch.cat(((lo + hi) * 0.5, hi - lo), dim=1)


def _rows(boxes, scores):
    if boxes.size(-1) == 8:
        raise NotImplementedError("rotated-box NMS (nms_rboxes) is outside the ensemble path")
    if boxes.dtype != scores.dtype or boxes.dtype not in (torch.float64, torch.float32):
        raise TypeError("nms: float64 (ensemble path) or float32 (detector head) boxes and scores expected, got %s / %s"
                        % (boxes.dtype, scores.dtype))
    n = int(scores.shape[0])
    rows = torch.empty((n, 5), dtype=torch.float64, device=boxes.device)   # float32 values widen exactly
    rows[:, 0] = scores
    rows[:, 1:] = boxes
    return rows, n, boxes.dtype == torch.float32


def nms(boxes, scores, overlap=0.5, top_k=0, soft=False, conf_thresh=0, soft_nms_cut=1):
    """box_utils.py:307-395.  ``boxes`` [n,4] point form, ``scores`` [n], both float64 or both float32; the
    arithmetic runs in that dtype."""
    rows, n, f32 = _rows(boxes, scores)
    if n == 0:
        return [], scores.new_zeros(0)
    offsets = np.array([0, n], np.int32)
    if not soft:
        res = runtime.hardnms_groups(offsets, rows, float(overlap), top_k=int(top_k), max_group=n, compute_f32=f32)
        keep = torch.from_numpy(res["keep"][:int(res["kept_count"][0])].astype(np.int64))
        return keep, scores[keep.to(scores.device)]
    res = runtime.softnms_groups(offsets, rows, float(overlap), float(soft_nms_cut), 0.0, max_group=n,
                                 box_format=_abi.W2T_BOX_XYXY, top_k=int(top_k), conf_thresh=float(conf_thresh),
                                 want_ensemble=False, compute_f32=f32)
    kept = int(res["kept_count"][0])
    keep = [int(i) for i in res["src_index"][:kept]]
    return keep, scores.new_tensor(res["merged"][:kept, 0].copy())


def bbox_vote(bbox_nms, score_nms, bbox_all, score_all, thresh):
    """box_utils.py:401-430: every kept box becomes the score-weighted mean of the boxes that overlap it by
    IoU >= ``thresh``.  Same signature and dtype behaviour (``score_nms`` is unused there as well); the sums are
    tree reductions on the device, so results agree with the reference to rounding, not bit for bit."""
    if int(bbox_nms.shape[0]) == 0:
        return torch.zeros_like(bbox_nms)
    f32 = bbox_all.dtype == torch.float32
    out = runtime.bbox_vote(bbox_nms.detach().double().cpu().numpy(), bbox_all.detach().double().cpu().numpy(),
                            score_all.detach().double().cpu().numpy(), float(thresh), compute_f32=f32)
    return torch.from_numpy(out).to(dtype=bbox_nms.dtype, device=bbox_nms.device)

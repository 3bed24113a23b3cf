"""ctypes binding of ``oracle/libw2t_oracle.so`` (the plain-C CPU oracle).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Takes the same packed NumPy
arrays the product's host layer builds (``waymo_2d_tracking_b200.packing``), so
a test hands identical buffers to the CUDA library and to this oracle.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from waymo_2d_tracking_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libw2t_oracle.so")
_lib = None

_p = C.c_void_p
_EXPORTS = {
    "w2t_oracle_bbox_to_z": (None, [_p, _p]),
    "w2t_oracle_x_to_bbox": (None, [_p, _p]),
    "w2t_oracle_bbox_to_z_d": (None, [_p, C.c_int, _p]),
    "w2t_oracle_kf_init": (None, [_p, _p, _p, C.c_int]),
    "w2t_oracle_kf_predict": (None, [_p, _p]),
    "w2t_oracle_kf_update": (None, [_p, _p, _p, C.c_int]),
    "w2t_oracle_inv4": (None, [_p, _p]),
    "w2t_oracle_iou_matrix": (None, [_p, C.c_int, _p, C.c_int, _p]),
    "w2t_oracle_linear_assignment": (C.c_int, [_p, C.c_int, C.c_int, _p]),
    "w2t_oracle_associate": (C.c_int, [_p, C.c_int, _p, C.c_int, C.c_double, C.c_int, _p, _p]),
    "w2t_oracle_sort_track": (C.c_int, [C.POINTER(_abi.SortProblem), C.POINTER(_abi.SortResult)]),
    "w2t_oracle_softnms_groups": (C.c_int, [C.POINTER(_abi.NmsProblem), C.POINTER(_abi.NmsResult)]),
    "w2t_oracle_merge_detections": (C.c_int, [_p, _p, C.c_int, C.c_double, _p]),
    "w2t_oracle_fusion_groups": (C.c_int, [C.POINTER(_abi.NmsProblem), _p, C.c_int, C.POINTER(_abi.NmsResult)]),
    "w2t_oracle_hardnms_groups": (C.c_int, [C.POINTER(_abi.NmsProblem), C.POINTER(_abi.NmsResult)]),
    "w2t_oracle_hard_nms": (C.c_int, [_p, _p, C.c_int, C.c_double, C.c_int, _p]),
    "w2t_oracle_soft_nms": (C.c_int, [_p, _p, C.c_int, C.c_double, C.c_int, C.c_double, C.c_double, _p, _p]),
}


def build(force=False):
    src = os.path.join(_HERE, "csrc", "w2t_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "libw2t_oracle.so"], check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = _abi.bind(C.CDLL(_SO), _EXPORTS)
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_p)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


# ---- building blocks ------------------------------------------------------

def bbox_to_z(det, promotion=None):
    """convert_bbox_to_z (sort.py:50-62) of one float32 row as the float64 values the filter receives."""
    det = _c(det, np.float32)
    z = np.zeros(4)
    lib().w2t_oracle_bbox_to_z_d(_ptr(det), _abi.promotion_code(promotion), _ptr(z))
    return z


def kf_init(det, promotion=None):
    det = _c(det, np.float32)
    x, P = np.zeros(7), np.zeros(49)
    lib().w2t_oracle_kf_init(_ptr(det), _ptr(x), _ptr(P), _abi.promotion_code(promotion))
    return x, P.reshape(7, 7)


def kf_predict(x, P):
    x, P = _c(x, np.float64).copy().reshape(7), _c(P, np.float64).copy().reshape(49)
    lib().w2t_oracle_kf_predict(_ptr(x), _ptr(P))
    return x, P.reshape(7, 7)


def kf_update(x, P, det, promotion=None):
    x, P = _c(x, np.float64).copy().reshape(7), _c(P, np.float64).copy().reshape(49)
    det = _c(det, np.float32)
    lib().w2t_oracle_kf_update(_ptr(x), _ptr(P), _ptr(det), _abi.promotion_code(promotion))
    return x, P.reshape(7, 7)


def x_to_bbox(x):
    x = _c(x, np.float64).reshape(7)
    b = np.zeros(4)
    lib().w2t_oracle_x_to_bbox(_ptr(x), _ptr(b))
    return b


def inv4(S):
    S = _c(S, np.float64).reshape(16)
    out = np.zeros(16)
    lib().w2t_oracle_inv4(_ptr(S), _ptr(out))
    return out.reshape(4, 4)


def iou_matrix(dets, trks):
    dets = _c(dets, np.float32).reshape(-1, 4)
    trks = _c(trks, np.float64).reshape(-1, 4)
    out = np.zeros((len(dets), len(trks)), np.float32)
    lib().w2t_oracle_iou_matrix(_ptr(dets), len(dets), _ptr(trks), len(trks), _ptr(out))
    return out


def linear_assignment(cost):
    cost = _c(cost, np.float32)
    D, T = cost.shape
    pairs = np.zeros((min(D, T), 2), np.int32)
    k = lib().w2t_oracle_linear_assignment(_ptr(cost), D, T, _ptr(pairs))
    assert k >= 0
    return pairs[:k].astype(int)


def soft_nms(boxes, scores, overlap=0.5, top_k=0, conf_thresh=0.0, soft_nms_cut=1.0):
    boxes = _c(boxes, np.float64).reshape(-1, 4)
    scores = _c(scores, np.float64).reshape(-1)
    n = len(scores)
    keep = np.zeros(n, np.int32)
    ns = np.zeros(n, np.float64)
    k = lib().w2t_oracle_soft_nms(_ptr(boxes), _ptr(scores), n, overlap, top_k, conf_thresh, soft_nms_cut,
                                  _ptr(keep), _ptr(ns))
    return keep[:k].tolist(), ns[:k].copy()


def hard_nms(boxes, scores, overlap=0.5, top_k=0):
    boxes = _c(boxes, np.float64).reshape(-1, 4)
    scores = _c(scores, np.float64).reshape(-1)
    n = len(scores)
    keep = np.zeros(n, np.int32)
    k = lib().w2t_oracle_hard_nms(_ptr(boxes), _ptr(scores), n, overlap, top_k, _ptr(keep))
    return keep[:k].tolist()


def merge_detections(detections, nms_thresh=0.5):
    """tta.py:22-66 on a list of [n_k,5] arrays (inputs are not modified)."""
    stacked = [_c(d, np.float64).reshape(-1, 5) for d in detections]
    rows = _c(np.vstack(stacked), np.float64)
    counts = np.asarray([len(d) for d in stacked], np.int32)
    out = np.zeros((max(len(rows), 1), 5))
    k = lib().w2t_oracle_merge_detections(_ptr(rows), _ptr(counts), len(counts), float(nms_thresh), _ptr(out))
    return out[:k].copy()


def fusion_groups(group_offsets, rows, sub_counts, iou_thresh, min_score, n_classes=0, score_thr=None, box_format=0):
    group_offsets = _c(group_offsets, np.int32)
    rows = _c(rows, np.float64).reshape(-1, 5)
    G, N = len(group_offsets) - 1, len(rows)
    sub_counts = _c(sub_counts, np.int32).reshape(G, -1)
    prob = _abi.NmsProblem()
    prob.n_groups = G
    prob.group_offsets, prob.rows = _ptr(group_offsets), _ptr(rows)
    prob.iou_thresh, prob.soft_nms_cut, prob.min_score = float(iou_thresh), 1.0, float(min_score)
    prob.n_classes = int(n_classes)
    thr = None if score_thr is None else _c(score_thr, np.float64)
    prob.score_thr = _ptr(thr)
    prob.box_format = int(box_format)
    out = dict(
        merged=np.zeros((N, 5)), kept_count=np.zeros(G, np.int32), ens_count=np.zeros(G, np.int32),
        ens_box=np.zeros((N, 4), np.int32), ens_score=np.zeros(N), trk_count=np.zeros(G, np.int32),
        trk_box=np.zeros((N, 4), np.float32), img_exists=np.zeros(G // n_classes, np.uint8) if n_classes else None)
    res = _abi.NmsResult()
    for k in ("merged", "kept_count", "ens_count", "ens_box", "ens_score", "trk_count", "trk_box", "img_exists"):
        setattr(res, k, _ptr(out[k]))
    out["status"] = lib().w2t_oracle_fusion_groups(C.byref(prob), _ptr(sub_counts), sub_counts.shape[1], C.byref(res))
    return out


# ---- packed stages ----------------------------------------------------------

def sort_track(packed, iou_thresholds, max_age, min_hits, final_cap=0, promotion=None):
    """``packed``: waymo_2d_tracking_b200.packing.PackedTracks (host arrays)."""
    S, NC = packed.n_streams, packed.n_classes
    n_img = int(packed.stream_img_offsets[-1])
    N = int(packed.n_rows)
    prob = _abi.SortProblem()
    prob.n_streams, prob.n_classes = S, NC
    keep = dict(
        off=_c(packed.stream_img_offsets, np.int32), start=_c(packed.det_start, np.int32),
        count=_c(packed.det_count, np.int32), box=_c(packed.det_box, np.float32),
        exists=None if packed.img_exists is None else _c(packed.img_exists, np.uint8),
        cam=_c(packed.cam_wh, np.float64))
    prob.stream_img_offsets = _ptr(keep["off"])
    prob.det_start = _ptr(keep["start"])
    prob.det_count = _ptr(keep["count"])
    prob.det_box = _ptr(keep["box"])
    prob.img_exists = _ptr(keep["exists"])
    prob.cam_wh = _ptr(keep["cam"])
    for i in range(NC):
        prob.iou_thr[i] = float(iou_thresholds[i])
    prob.max_age, prob.min_hits = int(max_age), int(min_hits)
    prob.promotion = _abi.promotion_code(promotion)
    out = dict(
        out_box=np.zeros((N, 4)), out_score=np.zeros(N), out_birth=np.zeros((N, 2), np.int32),
        out_count=np.zeros(n_img * NC, np.int32), created=np.zeros(n_img * NC, np.int32),
        first_img=np.zeros(S * NC, np.int32), final_count=np.zeros(S * NC, np.int32),
        final_state=np.zeros((S * NC, max(final_cap, 1), 56)) if final_cap else None)
    res = _abi.SortResult()
    for k in ("out_box", "out_score", "out_birth", "out_count", "created", "first_img", "final_count", "final_state"):
        setattr(res, k, _ptr(out[k]))
    res.final_cap = int(final_cap)
    status = lib().w2t_oracle_sort_track(C.byref(prob), C.byref(res))
    out["status"] = status
    return out


def softnms_groups(group_offsets, rows, iou_thresh, soft_nms_cut, min_score, n_classes=0, score_thr=None,
                   box_format=0, top_k=0, conf_thresh=0.0, hard=False):
    group_offsets = _c(group_offsets, np.int32)
    if isinstance(rows, np.ndarray) and rows.dtype.names is not None:   # packing.compact_rows
        rows = np.ascontiguousarray(rows)
        box_format = _abi.W2T_BOX_LTWH_I16
    elif isinstance(rows, np.ndarray) and rows.dtype == np.uint64:      # packing.packed_rows
        rows = np.ascontiguousarray(rows)
        box_format = _abi.W2T_BOX_LTWH_P64
    else:
        rows = _c(rows, np.float64).reshape(-1, 5)
    G, N = len(group_offsets) - 1, len(rows)
    prob = _abi.NmsProblem()
    prob.n_groups = G
    prob.group_offsets = _ptr(group_offsets)
    prob.rows = _ptr(rows)
    prob.iou_thresh, prob.soft_nms_cut, prob.min_score = float(iou_thresh), float(soft_nms_cut), float(min_score)
    prob.n_classes = int(n_classes)
    thr = None if score_thr is None else _c(score_thr, np.float64)
    prob.score_thr = _ptr(thr)
    prob.box_format, prob.top_k, prob.conf_thresh = int(box_format), int(top_k), float(conf_thresh)
    out = dict(
        merged=np.zeros((N, 5)), src_index=np.zeros(N, np.int32), kept_count=np.zeros(G, np.int32),
        ens_count=np.zeros(G, np.int32),
        ens_box=np.zeros((N, 4), np.int32), ens_score=np.zeros(N), trk_count=np.zeros(G, np.int32),
        trk_box=np.zeros((N, 4), np.float32),
        img_exists=np.zeros(G // n_classes, np.uint8) if n_classes else None)
    res = _abi.NmsResult()
    for k in ("merged", "src_index", "kept_count", "ens_count", "ens_box", "ens_score", "trk_count", "trk_box",
              "img_exists"):
        setattr(res, k, _ptr(out[k]))
    fn = lib().w2t_oracle_hardnms_groups if hard else lib().w2t_oracle_softnms_groups
    status = fn(C.byref(prob), C.byref(res))
    out["status"] = status
    return out

"""CPU restatement ("port") of the reference soft-NMS ensemble path.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Follows
``/root/reference/detnet/ensemble.py:19-64,78-84``,
``detnet/nn/tta.py:8-19`` and ``detnet/utils/box_utils.py:32-35,57-69,307-395``
(soft branch) in NumPy float64.  The reference runs the same element-wise
IEEE operations through torch CPU tensors, so results are bit-identical; this
is verified against the reference's own files by
``tests/golden/make_golden.py`` / ``tests/test_oracle.py``.

Tie rule: the reference's ``scores.sort(0)`` is unstable, so
the order of equal scores is implementation-defined there.  The canonical rule
used by this port and by the CUDA path is a *stable ascending* sort consumed
from the end: among equal scores the box with the larger concatenation index
is processed first.
"""
from collections import defaultdict

import numpy as np


def point_form(boxes):
    """box_utils.py:32-35."""
    half = boxes[..., 2:4] * 0.5
    return np.concatenate((boxes[..., :2] - half, boxes[..., :2] + half), axis=-1)


def center_size(boxes):
    """box_utils.py:57-69."""
    lo, hi = boxes[:, :2], boxes[:, 2:4]
    return np.concatenate(((lo + hi) * 0.5, hi - lo), axis=1)


def soft_nms(boxes, scores, overlap=0.5, top_k=0, conf_thresh=0, soft_nms_cut=1):
    """box_utils.py:307-395 with ``soft=True``: returns (keep indices, decayed scores)."""
    boxes = np.asarray(boxes, dtype=np.float64)
    scores = np.asarray(scores, dtype=np.float64)
    idx = np.argsort(scores, kind='stable')
    live = scores[idx]
    if top_k > 0:
        idx = idx[-top_k:]
        live = live[-top_k:]
    live = live.copy()
    x1, y1, x2, y2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    area = (x2 - x1) * (y2 - y1)
    keep, out = [], []
    denom = soft_nms_cut - overlap
    with np.errstate(divide='ignore', invalid='ignore'):
        while idx.shape[0] > 1:
            i = int(idx[-1])
            keep.append(i)
            idx = idx[:-1]
            w = np.maximum(np.minimum(x2[idx], x2[i]) - np.maximum(x1[idx], x1[i]), 0.0)
            h = np.maximum(np.minimum(y2[idx], y2[i]) - np.maximum(y1[idx], y1[i]), 0.0)
            inter = w * h
            union = (area[idx] - inter) + area[i]
            ratio = inter / union
            out.append(float(live[-1]))
            live = live[:-1]
            weight = np.minimum(np.maximum((soft_nms_cut - ratio) / denom, 0.0), 1.0)
            live = live * weight
            ok = live >= conf_thresh
            idx = idx[ok]
            live = live[ok]
    if idx.shape[0] > 0:
        keep.append(int(idx[-1]))
        out.append(float(live[-1]))
    return keep, np.array(out, dtype=np.float64)


def nms_detections(detections, iou_thresh=0.5, soft=True, soft_nms_cut=1):
    """tta.py:8-19 (soft branch only)."""
    assert soft, "the oracle restates the soft-NMS method only"
    scores = np.vstack([d[:, 0:1] for d in detections]).flatten()
    boxes = point_form(np.vstack([d[:, 1:5] for d in detections]))
    keep, new_scores = soft_nms(boxes, scores, overlap=iou_thresh, soft_nms_cut=soft_nms_cut)
    boxes = center_size(boxes[keep].reshape(-1, 4))
    return np.concatenate((new_scores[:, None], boxes), axis=1)


def convert_submission(det_list, weight, min_score=0):
    """ensemble.py:31-47."""
    grouped = defaultdict(lambda: defaultdict(list))
    for det in det_list:
        bbox = det['bbox']
        if bbox[2] > 0 and bbox[3] > 0:
            row = [det['score'] * weight] + list(bbox)
            if row[0] >= min_score:
                grouped[det['image_id']][det['category_id']].append(row)
    return grouped


def ensemble_image(image_id, detections, category_ids, min_score, iou_thresh, soft_nms_cut):
    """ensemble.py:50-64 for one image; ``detections`` = one {category: rows} per submission."""
    out = []
    for category_id in category_ids:
        per_sub = [np.asarray(det[category_id], dtype=np.float64).reshape(-1, 5) for det in detections]
        for b in per_sub:                          # lxly2cxcy, ensemble.py:19-22
            b[:, 1:3] += b[:, 3:5] / 2
        merged = nms_detections(per_sub, iou_thresh=iou_thresh, soft=True, soft_nms_cut=soft_nms_cut)
        merged[:, 1:3] -= merged[:, 3:5] / 2       # cxcy2lxly, ensemble.py:25-28
        for row in merged:
            if row[0] > min_score:
                out.append({'image_id': image_id, 'category_id': category_id,
                            'bbox': row[1:].astype(int).tolist(), 'score': round(row[0], 5)})
    return out


def ensemble_all(submissions, weights=None, min_score=0.0, iou_thresh=0.5, soft_nms_cut=1.0, image_order=None):
    """ensemble.py:78-84 + :145-149.  ``submissions`` = list of det-dict lists.

    ``image_order``: the reference iterates a ``set`` of image ids (hash order,
    different on every run); pass an explicit order to make the result
    reproducible (default: sorted).
    """
    if weights is None:
        weights = [1] * len(submissions)
    top = max(weights)
    weights = [w / top for w in weights]
    category_ids = set(sum([[d['category_id'] for d in det] for det in submissions], []))
    grouped = [convert_submission(d, w, min_score) for d, w in zip(submissions, weights)]
    if image_order is None:
        image_order = sorted(set(sum([list(g.keys()) for g in grouped], [])))
    out = []
    for image_id in image_order:
        out += ensemble_image(image_id, [g[image_id] for g in grouped], category_ids,
                              min_score, iou_thresh, soft_nms_cut)
    return out

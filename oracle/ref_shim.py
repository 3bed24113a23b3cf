"""Load the reference's own hot-path files from ``/root/reference`` (dev container only).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  ``/root/reference`` does not
exist on the GPU box, so nothing that runs there may call this module; it is
used by ``tests/golden/make_golden.py`` (to generate the committed fixtures) and
by the ``not gpu`` oracle tests, which skip when the directory is absent.

Recipe = SURVEY.md Appendix E: stub the imports the hot path does not need
(matplotlib, skimage, shapely, the detector packages) and inject the two
restated third-party pieces (``oracle.munkres`` for scikit-learn 0.22.2's
``linear_assignment``, ``oracle.kalman`` for filterpy's ``KalmanFilter``).
No reference source is copied: the files are executed where they lie.
"""
import contextlib
import importlib
import importlib.util
import os
import sys
import types
from argparse import Namespace
from functools import partial

REF_ROOT = os.environ.get("W2T_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "tracking", "sort", "sort.py"))


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


@contextlib.contextmanager
def _scoped_modules(extra, drop_after):
    saved = {k: sys.modules.get(k) for k in list(extra) + list(drop_after)}
    sys.modules.update(extra)
    try:
        yield
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


_tracking_cache = None
_ensemble_cache = None


def load_tracking():
    """Returns (utils_module, sort_module, tracker_sort_module) of the reference."""
    global _tracking_cache
    if _tracking_cache is not None:
        return _tracking_cache
    from . import kalman, munkres
    import sklearn.utils  # real package; only the removed sub-module is injected
    stubs = {
        "matplotlib": _module("matplotlib"),
        "matplotlib.pyplot": _module("matplotlib.pyplot"),
        "matplotlib.patches": _module("matplotlib.patches"),
        "skimage": _module("skimage", io=None),
        "skimage.io": _module("skimage.io"),
        "sklearn.utils.linear_assignment_": _module("sklearn.utils.linear_assignment_",
                                                    linear_assignment=munkres.linear_assignment),
        "filterpy": _module("filterpy"),
        "filterpy.kalman": _module("filterpy.kalman", KalmanFilter=kalman.KalmanFilter),
    }
    stubs["matplotlib"].pyplot = stubs["matplotlib.pyplot"]
    stubs["matplotlib"].patches = stubs["matplotlib.patches"]
    stubs["skimage"].io = stubs["skimage.io"]
    generic = ["utils", "sort", "sort.sort", "sort.tracker_sort"]
    tracking_dir = os.path.join(REF_ROOT, "tracking")
    with _scoped_modules(stubs, generic):
        for k in generic:
            sys.modules.pop(k, None)
        sys.path.insert(0, tracking_dir)
        try:
            ref_utils = importlib.import_module("utils")
            ref_sort = importlib.import_module("sort.sort")
            ref_tracker_sort = importlib.import_module("sort.tracker_sort")
        finally:
            sys.path.remove(tracking_dir)
    _tracking_cache = (ref_utils, ref_sort, ref_tracker_sort)
    return _tracking_cache


def load_ensemble():
    """Returns (ensemble_module, tta_module, box_utils_module) of the reference."""
    global _ensemble_cache
    if _ensemble_cache is not None:
        return _ensemble_cache

    def get_num_workers(jobs, device='cpu'):
        n = jobs if jobs > 0 else os.cpu_count() + jobs
        return n

    pk = {}
    for name in ("detnet", "detnet.nn", "detnet.utils", "detnet.trainer"):
        pk[name] = _module(name, __path__=[])
    pk["detnet.trainer.utils"] = _module("detnet.trainer.utils", get_num_workers=get_num_workers)
    pk["shapely"] = _module("shapely")
    pk["shapely.geometry"] = _module("shapely.geometry", asPolygon=None)
    loaded = ["detnet.utils.box_utils", "detnet.nn.tta", "detnet.ensemble"]
    files = {
        "detnet.utils.box_utils": "detnet/utils/box_utils.py",
        "detnet.nn.tta": "detnet/nn/tta.py",
        "detnet.ensemble": "detnet/ensemble.py",
    }
    mods = []
    with _scoped_modules(pk, loaded):
        for name in loaded:
            spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, files[name]))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[name] = mod
            spec.loader.exec_module(mod)
            mods.append(mod)
    box_utils, tta, ens = mods
    _ensemble_cache = (ens, tta, box_utils)
    return _ensemble_cache


@contextlib.contextmanager
def stable_torch_sort():
    """Force ``torch.Tensor.sort`` to be stable (canonical tie rule)."""
    import torch
    orig = torch.Tensor.sort

    def stable_sort(self, *args, **kwargs):
        kwargs.pop("stable", None)
        if args:
            kwargs.setdefault("dim", args[0])
            if len(args) > 1:
                kwargs.setdefault("descending", args[1])
        return orig(self, stable=True, **kwargs)

    torch.Tensor.sort = stable_sort
    try:
        yield
    finally:
        torch.Tensor.sort = orig


def _legacy_convert_bbox_to_z(bbox):
    """sort.py:50-62 statement by statement, with the one thing NumPy 2 no longer does spelled out: under NumPy
    1.x (the reference's pinned environment) a float32 SCALAR combined with a python float gives float64, so
    ``w / 2.``, ``bbox[0] + ...`` and ``w / float(h)`` are float64 operations; ``w``, ``h`` and ``s`` stay float32."""
    import numpy as np
    w = bbox[2] - bbox[0]
    h = bbox[3] - bbox[1]
    with np.errstate(divide='ignore', invalid='ignore'):
        x = np.float64(bbox[0]) + np.float64(w) / 2.
        y = np.float64(bbox[1]) + np.float64(h) / 2.
        s = w * h  # scale is just area
        r = np.float64(w) / float(h)
    return np.array([x, y, s, r]).reshape((4, 1))


def ref_track_all(predictions, iou_thresholds, max_age, min_hits, promotion="nep50"):
    """tracking/track.py:42-47 executed with the reference's own modules.

    ``promotion="nep50"``: the files exactly as they are, under this container's NumPy 2.
    ``promotion="legacy"``: EMULATION of the reference's pinned NumPy 1.x environment, which cannot be installed
    here — the reference's files still run, with two substitutions that restore value-based casting where the hot
    path depends on it: ``convert_bbox_to_z`` is replaced by ``_legacy_convert_bbox_to_z`` and the IoU thresholds
    are passed as ``numpy.float64`` scalars, so that ``iou_matrix[..] < iou_threshold`` (sort.py:220) compares in
    float64 as it did under NumPy 1.x."""
    import numpy as np
    ref_utils, ref_sort, _ = load_tracking()
    ref_sort.KalmanBoxTracker.count = 0
    legacy = str(promotion).lower() == "legacy"
    if not legacy and str(promotion).lower() != "nep50":
        raise ValueError("promotion must be 'legacy' or 'nep50'")
    saved = ref_sort.convert_bbox_to_z
    if legacy:
        ref_sort.convert_bbox_to_z = _legacy_convert_bbox_to_z
        iou_thresholds = [np.float64(t) for t in iou_thresholds]
    out = []
    try:
        for segment_id in predictions.keys():
            for camera_id in predictions[segment_id]:
                out += ref_utils.track_sort(predictions, segment_id, camera_id, iou_thresholds, max_age, min_hits)
    finally:
        ref_sort.convert_bbox_to_z = saved
    return out


def ref_ensemble_all(submissions, weights=None, min_score=0.0, iou_thresh=0.5, soft_nms_cut=1.0, image_order=None,
                     method="soft_nms"):
    """detnet/ensemble.py:123-149 executed with the reference's own modules (any of its three methods)."""
    ens, tta, _ = load_ensemble()
    if weights is None:
        weights = [1] * len(submissions)
    top = max(weights)
    weights = [w / top for w in weights]
    ens.args = Namespace(min_score=min_score)
    if method == "soft_nms":
        ens.merge_func = partial(tta.nms_detections, iou_thresh=iou_thresh, soft=True, soft_nms_cut=soft_nms_cut)
    elif method == "nms":
        ens.merge_func = partial(tta.nms_detections, iou_thresh=iou_thresh)
    else:
        ens.merge_func = partial(tta.merge_detections, nms_thresh=iou_thresh)
    category_ids = set(sum([[d['category_id'] for d in det] for det in submissions], []))
    grouped = [ens.convert_submission(d, w, min_score) for d, w in zip(submissions, weights)]
    if image_order is None:
        image_order = sorted(set(sum([list(g.keys()) for g in grouped], [])))
    out = []
    with stable_torch_sort():
        for image_id in image_order:
            out += ens.ensemble(image_id, [g[image_id] for g in grouped], category_ids)
    return out

"""Host runtime: torch owns device memory and streams, ``libw2t.so`` does the work.

Three entry points, all over the packed arrays of ``include/w2t_types.h``:

* :func:`softnms_groups` — the soft-NMS ensemble stage (``detnet/ensemble.py:145-157``);
* :func:`sort_track`     — the SORT stage (``tracking/track.py:42-47``);
* :func:`ensemble_and_track` — both, with the ensemble output handed to the tracker on the
  device (the int / 5-decimal rounding the reference applies when it writes the JSON in
  between is part of the data flow and is applied on the device).

Inputs may be NumPy arrays, CPU tensors (pinned ones are copied asynchronously) or CUDA
tensors.  There is no CPU fallback: without a CUDA device these functions raise.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _abi
from ._lib import W2TError, check, check_device_status, lib

_NP2T = {np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32,
         np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64, np.dtype(np.uint8): torch.uint8}


# When set to a list, every kernel launch through this module is bracketed by CUDA events on
# the launching stream: (kernel name, start, end).  bench.py uses it for the roofline numbers.
PROFILE = None


class _timed:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if PROFILE is not None:
            self.start = torch.cuda.Event(enable_timing=True)
            self.end = torch.cuda.Event(enable_timing=True)
            self.start.record()

    def __exit__(self, *exc):
        if PROFILE is not None:
            self.end.record()
            PROFILE.append((self.name, self.start, self.end))


def collect_profile():
    """Average device milliseconds per launch of each kernel recorded in PROFILE."""
    torch.cuda.synchronize()
    acc = {}
    for name, a, b in PROFILE or []:
        acc.setdefault(name, []).append(a.elapsed_time(b))
    return {k: sum(v) / len(v) for k, v in acc.items()}


# NumPy promotion regime the tracker reproduces (include/w2t_types.h, W2T_PROMOTION_*; default and meaning:
# _abi.DEFAULT_PROMOTION), overridable per call with ``promotion="legacy" | "nep50"``.
promotion_code = _abi.promotion_code


def require_cuda():
    if not torch.cuda.is_available():
        raise W2TError("no CUDA device visible: the box pipeline runs on sm_100a only (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _dev(a, dtype, device):
    """NumPy / CPU tensor / CUDA tensor -> contiguous CUDA tensor of ``dtype``."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        a = torch.from_numpy(np.ascontiguousarray(a, dtype=dtype))
    if a.dtype != _NP2T[np.dtype(dtype)]:
        a = a.to(_NP2T[np.dtype(dtype)])
    return a.to(device, non_blocking=True).contiguous()


def _host_rows(rows):
    """Ensemble input rows as a host tensor + their W2T_BOX_* layout: float64 [N,5] rows, or the
    16-byte compact rows of ``packing.compact_rows`` (structured array / uint8 [N,16] tensor)."""
    if isinstance(rows, np.ndarray) and rows.dtype == np.uint64:      # packing.packed_rows
        rows = torch.from_numpy(np.ascontiguousarray(rows).view(np.uint8).reshape(-1, 8))
    if isinstance(rows, np.ndarray) and rows.dtype.names is not None:
        rows = torch.from_numpy(np.ascontiguousarray(rows).view(np.uint8).reshape(-1, 16))
    if torch.is_tensor(rows) and rows.dtype == torch.uint8:
        if rows.dim() == 2 and rows.shape[1] == 8:
            return rows, _abi.W2T_BOX_LTWH_P64
        return rows.reshape(-1, 16), _abi.W2T_BOX_LTWH_I16
    if isinstance(rows, np.ndarray):
        rows = torch.from_numpy(np.ascontiguousarray(rows, np.float64))
    return rows.reshape(-1, 5), _abi.W2T_BOX_LTWH


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


# Pinned staging buffers are expensive to allocate (cudaHostAlloc), so they are kept and reused:
# one buffer per (tag, dtype), grown on demand.  A result returned to the caller is a view of such
# a buffer and is overwritten by the next call that uses the same tag.
_PINNED = {}


def _host(t, tag=None):
    """Device tensor -> pinned host tensor (async copy; the caller synchronises once)."""
    if t is None:
        return None
    n = t.numel()
    if tag is None:
        buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    else:
        key = (tag, t.dtype)
        pool = _PINNED.get(key)
        if pool is None or pool.numel() < n:
            pool = torch.empty(max(n, 1), dtype=t.dtype, pin_memory=True)
            _PINNED[key] = pool
        buf = pool[:n].view(t.shape)
    buf.copy_(t, non_blocking=True)
    return buf


def _np(t):
    return None if t is None else t.numpy()


# ---------------------------------------------------------------------------
# soft-NMS ensemble
# ---------------------------------------------------------------------------

def softnms_groups_device(d_offsets, d_rows, n_groups, max_group, iou_thresh, soft_nms_cut, min_score,
                          n_classes=0, score_thr=None, want_merged=True, box_format=_abi.W2T_BOX_LTWH, top_k=0,
                          conf_thresh=0.0, want_ensemble=True, hard=False, out=None, compute_f32=False):
    """Launch on device tensors; returns device tensors (no synchronisation).
    ``hard`` selects ``w2t_hardnms_groups`` (the ``-m nms`` method) instead of the soft branch.
    ``out``: preallocated result tensors to write into (all keys below; rows are indexed by the
    values of ``d_offsets``, so ``d_rows`` and the row outputs may be full arrays of a larger job
    while ``d_offsets`` and the per-group outputs are a chunk's slices)."""
    device = d_rows.device
    N = int(d_rows.shape[0])
    if out is not None:
        return _launch_nms(dict(out), d_offsets, d_rows, n_groups, max_group, iou_thresh, soft_nms_cut, min_score,
                           n_classes, score_thr, box_format, top_k, conf_thresh, hard, compute_f32)
    out = {
        "merged": torch.empty((N, 5), dtype=torch.float64, device=device) if want_merged else None,
        "src_index": torch.empty(N, dtype=torch.int32, device=device) if want_merged else None,
        "kept_count": torch.empty(n_groups, dtype=torch.int32, device=device) if want_merged else None,
        "ens_count": torch.empty(n_groups, dtype=torch.int32, device=device),
        "ens_box": torch.empty((N, 4), dtype=torch.int32, device=device) if want_ensemble else None,
        "ens_score": torch.empty(N, dtype=torch.float64, device=device) if want_ensemble else None,
        "trk_count": None, "trk_box": None, "img_exists": None,
        "status": torch.zeros(1, dtype=torch.int32, device=device),
    }
    if score_thr is not None:
        if n_classes < 1:
            raise W2TError("score_thr needs n_classes")
        out["trk_count"] = torch.empty(n_groups, dtype=torch.int32, device=device)
        out["trk_box"] = torch.empty((N, 4), dtype=torch.float32, device=device)
        out["img_exists"] = torch.zeros(max(n_groups // n_classes, 1), dtype=torch.uint8, device=device)
    return _launch_nms(out, d_offsets, d_rows, n_groups, max_group, iou_thresh, soft_nms_cut, min_score, n_classes,
                       score_thr, box_format, top_k, conf_thresh, hard, compute_f32)


def _launch_nms(out, d_offsets, d_rows, n_groups, max_group, iou_thresh, soft_nms_cut, min_score, n_classes,
                score_thr, box_format, top_k, conf_thresh, hard, compute_f32=False):
    thr = None
    if score_thr is not None:
        thr = (C.c_double * n_classes)(*[float(v) for v in score_thr[:n_classes]])
    prob = _abi.NmsProblem()
    prob.n_groups = int(n_groups)
    prob.group_offsets = _ptr(d_offsets)
    prob.rows = _ptr(d_rows)
    prob.iou_thresh, prob.soft_nms_cut, prob.min_score = float(iou_thresh), float(soft_nms_cut), float(min_score)
    prob.n_classes = int(n_classes)
    prob.score_thr = C.cast(thr, C.c_void_p) if thr is not None else None
    prob.box_format, prob.top_k, prob.conf_thresh = int(box_format), int(top_k), float(conf_thresh)
    prob.compute_f32 = 1 if compute_f32 else 0
    res = _abi.NmsResult()
    for k in ("merged", "src_index", "kept_count", "ens_count", "ens_box", "ens_score", "trk_count", "trk_box",
              "img_exists"):
        setattr(res, k, _ptr(out.get(k)))
    entry = lib().w2t_hardnms_groups if hard else lib().w2t_softnms_groups
    with _timed("softnms_kernel"):
        check(entry(C.byref(prob), C.byref(res), int(max_group), _ptr(out["status"]), _stream()),
              "w2t_hardnms_groups" if hard else "w2t_softnms_groups")
    return out


def softnms_groups(group_offsets, rows, iou_thresh=0.5, soft_nms_cut=1.0, min_score=0.0, n_classes=0,
                   score_thr=None, max_group=None, want_merged=True, box_format=_abi.W2T_BOX_LTWH, top_k=0,
                   conf_thresh=0.0, want_ensemble=True, hard=False, compute_f32=False):
    """Soft-NMS (or, with ``hard``, plain NMS) merge of every (image, category) group; NumPy in, NumPy out."""
    device = require_cuda()
    offs_np = group_offsets if isinstance(group_offsets, np.ndarray) else None
    n_groups = int(group_offsets.shape[0]) - 1
    if max_group is None:
        if offs_np is None:
            offs_np = torch.as_tensor(group_offsets).cpu().numpy()
        max_group = int(np.diff(offs_np).max()) if n_groups > 0 else 0
    d_offsets = _dev(group_offsets, np.int32, device)
    t_rows, fmt = _host_rows(rows)
    if fmt in (_abi.W2T_BOX_LTWH_I16, _abi.W2T_BOX_LTWH_P64):
        box_format = fmt
    d_rows = t_rows.to(device, non_blocking=True).contiguous()
    out = softnms_groups_device(d_offsets, d_rows, n_groups, max_group, iou_thresh, soft_nms_cut, min_score,
                                n_classes, score_thr, want_merged, box_format, top_k, conf_thresh, want_ensemble,
                                hard, compute_f32=compute_f32)
    host = {k: _host(v) for k, v in out.items()}
    torch.cuda.current_stream().synchronize()
    check_device_status(int(host["status"][0]), "soft-NMS")
    return {k: _np(v) for k, v in host.items()}


def fusion_groups_device(d_offsets, d_rows, d_sub_counts, n_sub, n_groups, max_group, iou_thresh, min_score,
                         n_classes=0, score_thr=None, want_merged=True, box_format=_abi.W2T_BOX_LTWH,
                         want_ensemble=True):
    """``w2t_fusion_groups`` on device tensors; returns device tensors (no synchronisation)."""
    device = d_rows.device
    N = int(d_rows.shape[0])
    out = {
        "merged": torch.empty((N, 5), dtype=torch.float64, device=device) if want_merged else None,
        "kept_count": torch.empty(n_groups, dtype=torch.int32, device=device),
        "ens_count": torch.empty(n_groups, dtype=torch.int32, device=device),
        "ens_box": torch.empty((N, 4), dtype=torch.int32, device=device) if want_ensemble else None,
        "ens_score": torch.empty(N, dtype=torch.float64, device=device) if want_ensemble else None,
        "trk_count": None, "trk_box": None, "img_exists": None,
        "status": torch.zeros(1, dtype=torch.int32, device=device),
    }
    thr = None
    if score_thr is not None:
        if n_classes < 1:
            raise W2TError("score_thr needs n_classes")
        thr = (C.c_double * n_classes)(*[float(v) for v in score_thr[:n_classes]])
        out["trk_count"] = torch.empty(n_groups, dtype=torch.int32, device=device)
        out["trk_box"] = torch.empty((N, 4), dtype=torch.float32, device=device)
        out["img_exists"] = torch.zeros(max(n_groups // n_classes, 1), dtype=torch.uint8, device=device)
    prob = _abi.NmsProblem()
    prob.n_groups = int(n_groups)
    prob.group_offsets, prob.rows = _ptr(d_offsets), _ptr(d_rows)
    prob.iou_thresh, prob.soft_nms_cut, prob.min_score = float(iou_thresh), 1.0, float(min_score)
    prob.n_classes = int(n_classes)
    prob.score_thr = C.cast(thr, C.c_void_p) if thr is not None else None
    prob.box_format = int(box_format)
    res = _abi.NmsResult()
    for k in ("merged", "kept_count", "ens_count", "ens_box", "ens_score", "trk_count", "trk_box", "img_exists"):
        setattr(res, k, _ptr(out[k]))
    with _timed("fusion_kernel"):
        check(lib().w2t_fusion_groups(C.byref(prob), _ptr(d_sub_counts), int(n_sub), C.byref(res), int(max_group),
                                      _ptr(out["status"]), _stream()), "w2t_fusion_groups")
    return out


def fusion_groups(group_offsets, rows, sub_counts, iou_thresh=0.5, min_score=0.0, n_classes=0, score_thr=None,
                  max_group=None, box_format=_abi.W2T_BOX_LTWH):
    """Weighted box fusion (``merge_detections``) of every group; NumPy in, NumPy out.
    ``sub_counts[G, K]``: rows of each group that come from each of the K input files."""
    device = require_cuda()
    n_groups = int(group_offsets.shape[0]) - 1
    sub_counts = np.ascontiguousarray(sub_counts, np.int32).reshape(n_groups, -1)
    if max_group is None:
        max_group = int(np.diff(np.asarray(group_offsets, np.int64)).max()) if n_groups > 0 else 0
    out = fusion_groups_device(_dev(group_offsets, np.int32, device), _dev(rows, np.float64, device).reshape(-1, 5),
                               _dev(sub_counts, np.int32, device), sub_counts.shape[1], n_groups, max_group,
                               iou_thresh, min_score, n_classes, score_thr, True, box_format)
    host = {k: _host(v) for k, v in out.items()}
    torch.cuda.current_stream().synchronize()
    check_device_status(int(host["status"][0]), "box fusion")
    return {k: _np(v) for k, v in host.items()}


def hardnms_groups(group_offsets, rows, iou_thresh=0.5, top_k=0, max_group=None, box_format=_abi.W2T_BOX_XYXY,
                   compute_f32=False):
    """``nms(soft=False)`` of every group: ``keep`` holds, per group from its first row on, the kept
    input rows (group-local indices) in descending score order; ``kept_count`` how many."""
    res = softnms_groups(group_offsets, rows, iou_thresh, 1.0, -np.inf, max_group=max_group, box_format=box_format,
                         top_k=top_k, want_ensemble=False, hard=True, compute_f32=compute_f32)
    offs = np.asarray(group_offsets, np.int64)
    local = res["src_index"].astype(np.int64) - np.repeat(offs[:-1], np.diff(offs))
    res["keep"] = local
    return res


# ---------------------------------------------------------------------------
# SORT
# ---------------------------------------------------------------------------

def make_plan(n_streams, n_classes, h_offsets, h_count, h_exists, max_age, group_offsets=None):
    """``w2t_sort_plan`` on host arrays.  With ``group_offsets`` (int32 [G+1]) instead of ``h_count`` /
    ``h_exists``: ``w2t_sort_plan_offsets`` — capacities from the group sizes (upper bounds before the ensemble)."""
    nq = n_streams * n_classes
    arrays = dict(order=np.zeros(nq, np.int32), track_cap=np.zeros(nq, np.int32),
                  det_cap=np.zeros(nq, np.int32), ws_offset=np.zeros(nq, np.int64))
    plan = _abi.SortPlan()
    for k, v in arrays.items():
        setattr(plan, k, v.ctypes.data_as(C.c_void_p))
    h_offsets = np.ascontiguousarray(h_offsets, np.int32)
    if group_offsets is not None:
        go = np.ascontiguousarray(group_offsets, np.int32)
        check(lib().w2t_sort_plan_offsets(int(n_streams), int(n_classes), h_offsets.ctypes.data_as(C.c_void_p),
                                          go.ctypes.data_as(C.c_void_p), int(max_age), C.byref(plan)),
              "w2t_sort_plan_offsets")
    else:
        h_count = np.ascontiguousarray(h_count, np.int32)
        h_exists = None if h_exists is None else np.ascontiguousarray(h_exists, np.uint8)
        check(lib().w2t_sort_plan(int(n_streams), int(n_classes), h_offsets.ctypes.data_as(C.c_void_p),
                                  h_count.ctypes.data_as(C.c_void_p),
                                  None if h_exists is None else h_exists.ctypes.data_as(C.c_void_p),
                                  int(max_age), C.byref(plan)), "w2t_sort_plan")
    arrays["ws_bytes"] = int(plan.ws_bytes)
    arrays["n_wide"] = int(plan.n_wide)
    arrays["n_mid"] = int(plan.n_mid)
    arrays["aux_offset"] = int(plan.aux_offset)
    arrays["narrow_cap"] = int(plan.narrow_cap)
    return arrays


FINALIZE_LAUNCHES = 5      # order, scan sums / offsets / apply, rows (csrc/finalize.cu)


def nms_launch_count(n_groups, max_group):
    """Kernels one ``w2t_softnms_groups`` call launches (``run_groups`` in csrc/softnms.cu): one launch per size
    class (up to 32 boxes, up to 96, the rest) when the job mixes sizes, plus the pass over global scratch for
    groups beyond shared memory."""
    cap = int(lib().w2t_softnms_max_group())
    regular = min(int(max_group), cap)
    classes = 1
    if regular > 64 and int(n_groups) >= 16384 and "W2T_NMS_CLASSES" not in os.environ:
        classes += int(regular > 32) + int(regular > 96)
    return classes + int(int(max_group) > cap)


def sort_launch_count(plan, n_substreams):
    """Kernels one ``w2t_sort_track`` call launches (``launch_sort`` in csrc/sort.cu, warp path): dmax, classify,
    the cluster kernel over the plan's wide / wide+mid classes, the warp kernel, the cluster kernel's second
    pass, the CTA kernel for what is left."""
    if plan.get("aux_offset", -1) < 0:
        return int(plan["n_wide"] > 0) + int(n_substreams > plan["n_wide"])
    return 2 + int(plan["n_wide"] > 0) + int(plan["n_wide"] + plan["n_mid"] > 0) + 3


def launch_classes(plan, order):
    """Reorder ``order`` into the launch classes of ``w2t_sort_plan_t`` — [wide | mid | narrow], each part
    keeping its relative order — and return (order, n_wide, n_mid)."""
    cap = plan["det_cap"][order]
    wide, mid = cap > _abi.W2T_WIDE_DETS, (cap > _abi.W2T_NARROW_DETS) & (cap <= _abi.W2T_WIDE_DETS)
    if wide.any() or mid.any():
        order = np.concatenate([order[wide], order[mid], order[~(wide | mid)]])
    return np.ascontiguousarray(order, np.int32), int(wide.sum()), int(mid.sum())


def sort_track_device(n_streams, n_classes, d_offsets, d_start, d_count, d_box, d_exists, d_cam,
                      iou_thresholds, max_age, min_hits, plan, final_cap=0, out=None, promotion=None):
    """Launch on device tensors; ``plan`` = host arrays from :func:`make_plan`.  No synchronisation.
    ``out``: preallocated result tensors (keys of ``w2t_sort_result_t`` + ``status``) to write into."""
    device = d_box.device
    NC = int(n_classes)
    if len(iou_thresholds) < NC:
        raise IndexError("iou_thresholds has %d entries for %d categories" % (len(iou_thresholds), NC))
    n_groups = int(d_start.shape[0])
    N = int(d_box.shape[0])
    nq = n_streams * NC
    out = dict(out) if out is not None else {
        "out_box": torch.empty((N, 4), dtype=torch.float64, device=device),
        "out_score": torch.empty(N, dtype=torch.float64, device=device),
        "out_birth": torch.empty((N, 2), dtype=torch.int32, device=device),
        "out_count": torch.empty(n_groups, dtype=torch.int32, device=device),
        "created": torch.empty(n_groups, dtype=torch.int32, device=device),
        "first_img": torch.empty(nq, dtype=torch.int32, device=device),
        "final_count": torch.empty(nq, dtype=torch.int32, device=device) if final_cap else None,
        "final_state": torch.zeros((nq, final_cap, 56), dtype=torch.float64, device=device) if final_cap else None,
        "status": torch.zeros(1, dtype=torch.int32, device=device),
    }
    d_plan = {k: _dev(plan[k], np.int64 if k == "ws_offset" else np.int32, device)
              for k in ("order", "track_cap", "det_cap", "ws_offset")}
    workspace = torch.empty(max(plan["ws_bytes"], 256), dtype=torch.uint8, device=device)
    prob = _abi.SortProblem()
    prob.n_streams, prob.n_classes = int(n_streams), NC
    prob.stream_img_offsets = _ptr(d_offsets)
    prob.det_start, prob.det_count, prob.det_box = _ptr(d_start), _ptr(d_count), _ptr(d_box)
    prob.img_exists = _ptr(d_exists)
    prob.cam_wh = _ptr(d_cam)
    for i in range(NC):
        prob.iou_thr[i] = float(iou_thresholds[i])
    prob.max_age, prob.min_hits = int(max_age), int(min_hits)
    prob.promotion = promotion_code(promotion)
    cplan = _abi.SortPlan()
    for k in ("order", "track_cap", "det_cap", "ws_offset"):
        setattr(cplan, k, _ptr(d_plan[k]))
    cplan.ws_bytes = plan["ws_bytes"]
    cplan.n_wide = int(plan.get("n_wide", 0))
    cplan.n_mid = int(plan.get("n_mid", 0))
    cplan.aux_offset = int(plan.get("aux_offset", -1))
    cplan.narrow_cap = int(plan.get("narrow_cap", 0))
    if plan.get("chunk_of") is not None:      # completion counters per chunk of sub-streams
        d_plan["chunk_of"] = _dev(plan["chunk_of"], np.int32, device)
        cplan.chunk_of, cplan.chunk_done = _ptr(d_plan["chunk_of"]), _ptr(plan["chunk_done"])
    res = _abi.SortResult()
    for k in ("out_box", "out_score", "out_birth", "out_count", "created", "first_img", "final_count", "final_state"):
        setattr(res, k, _ptr(out.get(k)))
    res.final_cap = int(final_cap)
    with _timed("sort_track_kernel"):
        check(lib().w2t_sort_track(C.byref(prob), C.byref(cplan), C.byref(res), _ptr(workspace), _ptr(out["status"]),
                                   _stream()), "w2t_sort_track")
    out["_keepalive"] = (d_plan, workspace)
    return out


def assign_ids(n_streams, n_classes, h_offsets, h_start, h_out_count, h_created, h_first_img, class_rank,
               h_out_birth, id_base=0):
    """``w2t_assign_ids`` on host arrays -> (ids[int64, N], next id base)."""
    def p(a, dt):
        return None if a is None else np.ascontiguousarray(a, dt)
    offs, start, cnt = p(h_offsets, np.int32), p(h_start, np.int32), p(h_out_count, np.int32)
    created, first, rank, birth = p(h_created, np.int32), p(h_first_img, np.int32), p(class_rank, np.int32), \
        p(h_out_birth, np.int32)
    ids = np.zeros(birth.shape[0], np.int64)
    nxt = C.c_int64(0)
    vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    check(lib().w2t_assign_ids(int(n_streams), int(n_classes), vp(offs), vp(start), vp(cnt), vp(created), vp(first),
                               vp(rank), vp(birth), int(id_base), vp(ids), C.byref(nxt)), "w2t_assign_ids")
    return ids, int(nxt.value)


def finalize_device(n_streams, n_classes, d_offsets, d_start, trk, d_class_rank, id_base, rows_cap, image_base=0,
                    rows=None, id_base_device=None, birth_group_base=0, compact=False):
    """Device-side ids + dense rows (``w2t_sort_finalize``) on the tensors ``sort_track_device`` returned.
    ``rows``: preallocated output tensors (same keys) to write into instead of allocating.  ``compact``: 8 bytes
    per row (``rows_compact`` [cap,2] int32: id - id_base, image * 8 + category - 1) instead of ``rows_id`` /
    ``rows_img`` / ``rows_cat`` (16 bytes): what the pipelined path sends home (:class:`HostRows` decodes it)."""
    device = trk["out_box"].device
    NC = int(n_classes)
    n_groups = int(d_start.shape[0])
    cap = max(int(rows_cap), 1)
    if rows is None:
        rows = {"rows_box": torch.empty((cap, 4), dtype=torch.float64, device=device),
                "rows_score": torch.empty(cap, dtype=torch.float64, device=device),
                "totals": torch.zeros(3, dtype=torch.int64, device=device)}
        if compact:
            rows["rows_compact"] = torch.empty((cap, 2), dtype=torch.int32, device=device)
        else:
            rows.update({"rows_id": torch.empty(cap, dtype=torch.int64, device=device),
                         "rows_img": torch.empty(cap, dtype=torch.int32, device=device),
                         "rows_cat": torch.empty(cap, dtype=torch.int32, device=device)})
    ws = torch.empty(int(lib().w2t_sort_finalize_workspace(int(n_streams), NC, n_groups)), dtype=torch.uint8,
                     device=device)
    prob = _abi.SortProblem()
    prob.n_streams, prob.n_classes = int(n_streams), NC
    prob.stream_img_offsets, prob.det_start = _ptr(d_offsets), _ptr(d_start)
    res = _abi.SortResult()
    for k in ("out_box", "out_score", "out_birth", "out_count", "created", "first_img"):
        setattr(res, k, _ptr(trk[k]))
    crows = _abi.Rows()
    crows.box, crows.score, crows.totals = _ptr(rows["rows_box"]), _ptr(rows["rows_score"]), _ptr(rows["totals"])
    crows.object_id, crows.image, crows.category = (_ptr(rows.get(k)) for k in ("rows_id", "rows_img", "rows_cat"))
    crows.compact = _ptr(rows.get("rows_compact"))
    crows.capacity = cap
    crows.image_base = int(image_base)
    crows.id_base_device = _ptr(id_base_device)
    crows.birth_group_base = int(birth_group_base)
    with _timed("finalize_kernels"):
        check(lib().w2t_sort_finalize(C.byref(prob), C.byref(res), _ptr(d_class_rank), int(id_base), n_groups,
                                      _ptr(ws), C.byref(crows), _stream()), "w2t_sort_finalize")
    rows["_keepalive"] = ws
    return rows


_ROW_KEYS = ("rows_box", "rows_score", "rows_id", "rows_img", "rows_cat", "totals")
_COMPACT_ROW_KEYS = ("rows_box", "rows_score", "rows_compact")


class HostRows(dict):
    """Result of the pipelined path.  The device sends 48 bytes per row home — box, confidence and 8 bytes holding
    (object id - ``id_base``, image * 8 + category - 1) — and ``rows_id`` (int64), ``rows_img`` and ``rows_cat``
    (int32) are decoded from them on first access."""

    def __missing__(self, key):
        pair = dict.__getitem__(self, "rows_compact")
        if key == "rows_id":
            value = pair[:, 0].astype(np.int64) + np.int64(dict.__getitem__(self, "id_base"))
        elif key == "rows_img":
            value = pair[:, 1] >> 3
        elif key == "rows_cat":
            value = (pair[:, 1] & 7) + 1
        else:
            raise KeyError(key)
        self[key] = value
        return value
_RAW_KEYS = ("out_box", "out_score", "out_birth", "out_count", "created", "first_img", "final_count", "final_state")


def _collect(trk, rows, raw, extra=None):
    """Copy the requested device results to pinned host buffers, sync once, check the status."""
    host = {k: _host(rows[k], k) for k in _ROW_KEYS}
    host["status"] = _host(trk["status"], "status")
    if raw:
        host.update({k: _host(trk[k], k) for k in _RAW_KEYS if trk.get(k) is not None})
    for k, v in (extra or {}).items():
        host[k] = _host(v, k)
    torch.cuda.current_stream().synchronize()
    check_device_status(int(host["status"][0]), "SORT")
    d2h = sum(v.numel() * v.element_size() for v in host.values() if v is not None)
    res = {k: _np(v) for k, v in host.items()}
    res["d2h_bytes"] = d2h
    n_rows = int(res["totals"][1])
    for k in _ROW_KEYS[:-1]:
        res[k] = res[k][:n_rows]
    res["n_rows"] = n_rows
    return res


def upload_tracks(packed):
    """Device copies of a ``packing.PackedTracks`` (for callers that keep the inputs resident in HBM)."""
    device = require_cuda()
    return dict(offsets=_dev(packed.stream_img_offsets, np.int32, device), start=_dev(packed.det_start, np.int32, device),
                count=_dev(packed.det_count, np.int32, device),
                box=_dev(packed.det_box, np.float32, device).reshape(-1, 4),
                exists=_dev(packed.img_exists, np.uint8, device), cam=_dev(packed.cam_wh, np.float64, device),
                rank=_dev(packed.class_rank, np.int32, device))


def sort_track(packed, iou_thresholds, max_age=1, min_hits=0, final_cap=0, id_base=0, raw=True, promotion=None,
               dev=None, to_host=True):
    """SORT over every stream of ``packed`` (``packing.PackedTracks``).

    Returns the dense output list (``rows_box/score/id/img/cat`` in the reference's order, ids
    assigned on the device) and, with ``raw``, the per-slot arrays of ``w2t_sort_result_t``.
    ``dev``: device copies from :func:`upload_tracks` (skips the host->device copies);
    ``to_host=False``: leave the results in HBM and return the device tensors (``trk``, ``rows``).
    """
    require_cuda()
    S, NC = packed.n_streams, packed.n_classes
    plan = make_plan(S, NC, packed.stream_img_offsets, packed.det_count, packed.img_exists, max_age)
    d = dev if dev is not None else upload_tracks(packed)
    d_offsets, d_start = d["offsets"], d["start"]
    trk = sort_track_device(S, NC, d_offsets, d_start, d["count"], d["box"], d["exists"], d["cam"], iou_thresholds,
                            max_age, min_hits, plan, final_cap, promotion=promotion)
    rows = finalize_device(S, NC, d_offsets, d_start, trk, d["rank"], id_base,
                           int(np.asarray(packed.det_count, np.int64).sum()))
    if not to_host:
        return {"trk": trk, "rows": rows, "launches": sort_launch_count(plan, S * NC) + FINALIZE_LAUNCHES}
    res = _collect(trk, rows, raw)
    res["id_next"] = int(id_base + res["totals"][0])
    if raw:
        res["ids"], nxt = assign_ids(S, NC, packed.stream_img_offsets, packed.det_start, res["out_count"],
                                     res["created"], res["first_img"], packed.class_rank, res["out_birth"], id_base)
        assert nxt == res["id_next"], "host and device id scans disagree"
    return res


# ---------------------------------------------------------------------------
# ensemble -> SORT without leaving the device
# ---------------------------------------------------------------------------

def ensemble_and_track(group_offsets, rows, stream_img_offsets, cam_wh, n_classes, iou_thresh, soft_nms_cut,
                       min_score, score_thr, iou_thresholds, max_age, min_hits, max_group=None,
                       want_ensemble=True, id_base=0, raw=True, to_host=True, host_group_offsets=None,
                       promotion=None):
    """Groups must be laid out as g = img * n_classes + (category - 1) with the images of a
    stream contiguous and in frame order (``synth.groups_from_scene`` / ``packing``).

    ``group_offsets`` / ``rows`` may be NumPy, CPU (ideally pinned) or CUDA tensors.
    ``host_group_offsets``: host copy of ``group_offsets`` when those live on the device; with it
    (``to_host=False`` only) the SORT launch plan is computed from the INPUT group sizes — an upper
    bound of what survives the ensemble — and the call never synchronises with the device."""
    device = require_cuda()
    NC = int(n_classes)
    h_offsets = np.ascontiguousarray(stream_img_offsets, np.int32)
    S = len(h_offsets) - 1
    n_groups = int(group_offsets.shape[0]) - 1
    if n_groups != int(h_offsets[-1]) * NC:
        raise W2TError("group layout does not match stream_img_offsets * n_classes")
    if max_group is None:
        go = group_offsets if isinstance(group_offsets, np.ndarray) else torch.as_tensor(group_offsets).cpu().numpy()
        max_group = int(np.diff(go).max()) if n_groups else 0
    d_goff = _dev(group_offsets, np.int32, device)
    t_rows, fmt = _host_rows(rows)
    d_rows = t_rows.to(device, non_blocking=True).contiguous()
    nms = softnms_groups_device(d_goff, d_rows, n_groups, max_group, iou_thresh, soft_nms_cut, min_score, NC,
                                score_thr, want_merged=False, want_ensemble=want_ensemble, box_format=fmt)
    d_offsets = _dev(h_offsets, np.int32, device)
    d_start = d_goff[:-1]
    if host_group_offsets is not None and not to_host and n_groups:
        # no round trip: capacities from the input sizes; W2T_ERR_CAPACITY (only possible when the
        # ensemble drops every box of an image) is left in trk["status"] for the caller to check
        plan = make_plan(S, NC, h_offsets, None, None, max_age, group_offsets=np.asarray(host_group_offsets))
        trk = sort_track_device(S, NC, d_offsets, d_start, nms["trk_count"], nms["trk_box"], nms["img_exists"],
                                _dev(cam_wh, np.float64, device), iou_thresholds, max_age, min_hits, plan,
                                promotion=promotion)
        out_rows = finalize_device(S, NC, d_offsets, d_start, trk, None, id_base, int(d_rows.shape[0]))
        return {"nms": nms, "trk": trk, "rows": out_rows, "n_trk": None,
                "launches": nms_launch_count(n_groups, max_group) + sort_launch_count(plan, S * NC) + FINALIZE_LAUNCHES}
    # the plan needs the surviving counts on the host: one small D2H between the stages
    h_cnt, h_exists, h_nms_status = _host(nms["trk_count"], "trk_count"), _host(nms["img_exists"], "img_exists"), \
        _host(nms["status"], "nms_status")
    torch.cuda.current_stream().synchronize()
    check_device_status(int(h_nms_status[0]), "soft-NMS")
    plan = make_plan(S, NC, h_offsets, h_cnt.numpy(), h_exists.numpy(), max_age)
    trk = sort_track_device(S, NC, d_offsets, d_start, nms["trk_count"], nms["trk_box"], nms["img_exists"],
                            _dev(cam_wh, np.float64, device), iou_thresholds, max_age, min_hits, plan,
                            promotion=promotion)
    out_rows = finalize_device(S, NC, d_offsets, d_start, trk, None, id_base, int(h_cnt.numpy().sum(dtype=np.int64)))
    n_trk = int(h_cnt.numpy().sum(dtype=np.int64))
    if not to_host:
        # results stay in HBM (bench.py's device-resident leg); 1 soft-NMS + 1 SORT + 4 finalize kernels
        dev = {"nms": nms, "trk": trk, "rows": out_rows, "n_trk": n_trk,
               "launches": nms_launch_count(n_groups, max_group) + sort_launch_count(plan, S * NC) + FINALIZE_LAUNCHES}
        return dev
    extra = {k: nms[k] for k in ("ens_count", "ens_box", "ens_score")} if want_ensemble else {}
    res = _collect(trk, out_rows, raw, extra)
    res["d2h_bytes"] += h_cnt.numel() * 4 + h_exists.numel() + 4
    res["n_trk"] = n_trk
    res["launches"] = nms_launch_count(n_groups, max_group) + sort_launch_count(plan, S * NC) + FINALIZE_LAUNCHES
    res["trk_count"], res["img_exists"] = h_cnt.numpy(), h_exists.numpy()
    res["id_next"] = int(id_base + res["totals"][0])
    if raw:
        go_np = group_offsets if isinstance(group_offsets, np.ndarray) else torch.as_tensor(group_offsets).cpu().numpy()
        res["det_start"] = np.ascontiguousarray(go_np[:-1], np.int32)
        res["ids"], _ = assign_ids(S, NC, h_offsets, res["det_start"], res["out_count"], res["created"],
                                   res["first_img"], None, res["out_birth"], id_base)
    return res


# ---------------------------------------------------------------------------
# ensemble -> SORT, chunked and pipelined: H2D, kernels and D2H of successive chunks overlap
# ---------------------------------------------------------------------------

def _pinned_pool(tag, dtype, n, keep=0, quiesce=None):
    """Persistent pinned buffer of at least ``n`` elements; the first ``keep`` survive a regrow
    (``quiesce`` is called first so that no copy into the old buffer is still in flight)."""
    key = (tag, dtype)
    pool = _PINNED.get(key)
    if pool is None or pool.numel() < n:
        grown = torch.empty(max(int(n * 1.25), 1), dtype=dtype, pin_memory=True)
        if pool is not None and keep:
            if quiesce is not None:
                quiesce()
            grown[:keep].copy_(pool[:keep])
        _PINNED[key] = pool = grown
    return pool


_STREAMS = {}
TRACE = None   # debug aid: list receiving (label, host seconds, cuda event or None) from the pipelined path


def _trace(label, stream=None):
    if TRACE is not None:
        import time
        ev = None
        if stream is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(stream)
        TRACE.append((label, time.perf_counter(), ev))


def _side_streams(device):
    key = (device.index, "io")
    if key not in _STREAMS:
        _STREAMS[key] = (torch.cuda.Stream(device), torch.cuda.Stream(device))
    return _STREAMS[key]


def _compute_streams(device, n=3):
    """n compute streams + one for the finalize kernels."""
    key = (device.index, "compute")
    if key not in _STREAMS:
        # the last one (finalize + hand-over kernels) gets the highest priority: its small CTAs must
        # slip in between the CTAs of a SORT launch that fills the machine
        _STREAMS[key] = [torch.cuda.Stream(device) for _ in range(n)] + [torch.cuda.Stream(device, priority=-1)]
    return list(_STREAMS[key])


def _chunk_bounds(h_offsets, group_offsets_np, NC, n_chunks):
    """Split the streams into at most ``n_chunks`` contiguous blocks with about equal input rows;
    ``n_chunks`` may also be a sequence of fractions of the input rows (one per block)."""
    S = len(h_offsets) - 1
    rows_at_stream = group_offsets_np[np.asarray(h_offsets, np.int64) * NC].astype(np.int64)
    total = int(rows_at_stream[-1])
    if isinstance(n_chunks, (list, tuple)):
        fr = np.cumsum(np.asarray(n_chunks, np.float64))
        fr = (fr / fr[-1])[:max(1, min(len(fr), S))]
        n_chunks = len(fr)
    else:
        n_chunks = max(1, min(int(n_chunks), S))
        fr = np.arange(1, n_chunks + 1) / n_chunks
    cuts = [0]
    for k in range(1, n_chunks):
        s = int(np.searchsorted(rows_at_stream, total * fr[k - 1], side="left"))
        s = min(max(s, cuts[-1] + 1), S - (n_chunks - k))
        cuts.append(s)
    cuts.append(S)
    return [(a, b) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]


def ensemble_and_track_pipelined(group_offsets, rows, stream_img_offsets, cam_wh, n_classes, iou_thresh,
                                 soft_nms_cut, min_score, score_thr, iou_thresholds, max_age, min_hits,
                                 max_group=None, id_base=0, n_chunks=8, hoist=1.0, hoist_by_work=False, light_first=0,
                                 promotion=None):
    """Host buffers in, dense host rows out (``rows_box/score/id/img/cat``), the same result as
    :func:`ensemble_and_track` — with the streams cut into chunks so that copies and kernels overlap
    (PCIe is full duplex): the host->device copies of all chunks are queued on a copy stream; the
    soft-NMS of a chunk starts when its rows have landed; the SORT stage is one launch over all
    sub-streams whose CTAs are ordered chunk by chunk — after the long chains (``hoist``: of that
    fraction of trailing chunks; 1.0 = of all chunks, measured best), which start first so that none
    of them is left running alone at the end — and count themselves into per-chunk completion
    counters; ids, dense rows and the device->host copy of chunk k start as soon as chunk k is tracked
    (stream wait on the counter), while the later chunks are still being tracked.

    No mid-pipeline round trip: the launch plan of the SORT stage (slab capacities) is computed on
    the host from the INPUT group sizes, an upper bound of what survives the ensemble.  (An image
    all of whose boxes the ensemble drops is treated by the tracker as absent, which can lengthen
    the window the bound assumes; the kernel then reports W2T_ERR_CAPACITY and the call falls back
    to the exact, unpipelined path.)"""
    device = require_cuda()
    NC = int(n_classes)
    h_offsets = np.ascontiguousarray(stream_img_offsets, np.int32)
    S = len(h_offsets) - 1
    t_goff = group_offsets if torch.is_tensor(group_offsets) else torch.from_numpy(np.ascontiguousarray(group_offsets, np.int32))
    t_rows, fmt = _host_rows(rows)
    go_np = t_goff.numpy()
    G = int(go_np.shape[0]) - 1
    N = int(t_rows.shape[0])
    if G != int(h_offsets[-1]) * NC:
        raise W2TError("group layout does not match stream_img_offsets * n_classes")
    if S == 0 or G == 0:
        return ensemble_and_track(group_offsets, rows, stream_img_offsets, cam_wh, NC, iou_thresh, soft_nms_cut,
                                  min_score, score_thr, iou_thresholds, max_age, min_hits, max_group, False, id_base,
                                  raw=False, promotion=promotion)
    # host work is ordered so that the copies and the soft-NMS launches are queued first; everything
    # only the SORT launch needs (group sizes, plans) is computed while they run
    sizes = None
    if max_group is None:
        sizes = np.diff(go_np).astype(np.int32)
        max_group = int(sizes.max())
    chunks = _chunk_bounds(h_offsets, go_np, NC, n_chunks)
    cam = np.ascontiguousarray(cam_wh, np.float64).reshape(S, 2)

    _trace("host prep done")
    main = torch.cuda.current_stream()
    s_in, s_out = _side_streams(device)
    i32, f64 = torch.int32, torch.float64
    n_img = int(h_offsets[-1])
    d_goff = torch.empty(G + 1, dtype=i32, device=device)
    d_rows = torch.empty(t_rows.shape, dtype=t_rows.dtype, device=device)
    nms_out = {"ens_count": torch.empty(G, dtype=i32, device=device),
               "trk_count": torch.empty(G, dtype=i32, device=device),
               "trk_box": torch.empty((N, 4), dtype=torch.float32, device=device),
               "img_exists": torch.zeros(n_img, dtype=torch.uint8, device=device),
               "status": torch.zeros(1, dtype=i32, device=device)}
    trk_out = {"out_box": torch.empty((N, 4), dtype=f64, device=device), "out_score": torch.empty(N, dtype=f64, device=device),
               "out_birth": torch.empty((N, 2), dtype=i32, device=device), "out_count": torch.empty(G, dtype=i32, device=device),
               "created": torch.empty(G, dtype=i32, device=device), "first_img": torch.empty(S * NC, dtype=i32, device=device),
               "status": torch.zeros(1, dtype=i32, device=device)}
    d_cam = _dev(cam, np.float64, device)
    for t in (d_goff, d_rows):
        t.record_stream(s_in)

    _trace("alloc done", main)
    # 1. all host->device copies are queued up front on the copy-in stream, chunk by chunk
    h2d_done = []
    with torch.cuda.stream(s_in):
        s_in.wait_stream(main)
        for (s0, s1) in chunks:
            g0, g1 = int(h_offsets[s0]) * NC, int(h_offsets[s1]) * NC
            r0, r1 = int(go_np[g0]), int(go_np[g1])
            d_goff[g0:g1 + 1].copy_(t_goff[g0:g1 + 1], non_blocking=True)
            d_rows[r0:r1].copy_(t_rows[r0:r1], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(s_in)
            h2d_done.append(ev)
            _trace("h2d queued %d" % len(h2d_done), s_in)

    # 2. kernels.  soft-NMS runs chunk by chunk behind the copies.  The SORT stage is ONE launch over
    #    all sub-streams (it is latency-bound: a launch costs a full chain of images however few streams
    #    it holds), ordered chunk by chunk, heaviest sub-stream first inside a chunk; every CTA counts
    #    itself into its chunk's completion counter.  The finalize kernels of chunk k wait for that
    #    counter on their own high-priority stream (w2t_stream_wait_value32), so ids, dense rows and the
    #    copy home of chunk k overlap with the tracking of the later chunks.  The id base is handed
    #    from chunk to chunk on the device (w2t_rows_t.id_base_device).
    comp = _compute_streams(device)
    fin = comp[-1]
    comp = comp[:-1]
    for cs in comp + [fin]:
        cs.wait_stream(main)
    nq = S * NC
    for k, (s0, s1) in enumerate(chunks):
        img0, img1 = int(h_offsets[s0]), int(h_offsets[s1])
        g0, g1 = img0 * NC, img1 * NC
        cs = comp[k % len(comp)]
        with torch.cuda.stream(cs):
            cs.wait_event(h2d_done[k])
            nms_k = {"ens_count": nms_out["ens_count"][g0:g1], "trk_count": nms_out["trk_count"][g0:g1],
                     "trk_box": nms_out["trk_box"], "img_exists": nms_out["img_exists"][img0:img1],
                     "status": nms_out["status"]}
            softnms_groups_device(d_goff[g0:g1 + 1], d_rows, g1 - g0, max_group, iou_thresh, soft_nms_cut, min_score,
                                  NC, score_thr, out=nms_k, box_format=fmt)
    # one plan for all streams (slab capacities and offsets, heaviest-first order) from the INPUT group sizes
    # (an image exists if any of its groups has a row); the launch order is then regrouped chunk by chunk,
    # heaviest first inside a chunk
    plan_all = make_plan(S, NC, h_offsets, None, None, max_age, group_offsets=go_np)
    chunk_of_stream = np.zeros(S, np.int32)
    for k, (s0, s1) in enumerate(chunks):
        chunk_of_stream[s0:s1] = k
    plan_all["chunk_of"] = np.repeat(chunk_of_stream, NC)
    plan_all["order"] = np.ascontiguousarray(
        plan_all["order"][np.argsort(plan_all["chunk_of"][plan_all["order"]], kind="stable")], np.int32)
    _trace("nms queued")
    # Launch order of the single SORT launch: chunk by chunk, so that the early chunks finish early and
    # go home while the rest is tracked — except that the LONG chains (the crowded category) of the
    # trailing chunks are hoisted to the front: started last, a 200-image chain of that category would
    # run on alone after everything else is done (longest-processing-time-first for the tail).
    if hoist and len(chunks) > 1:
        work = plan_all["det_cap"].astype(np.int64) * plan_all["track_cap"]   # size of the largest assignment problem
        order = plan_all["order"]
        late = plan_all["chunk_of"][order] >= len(chunks) - max(1, int(round(hoist * len(chunks))))
        front = late & (work[order] > 0.4 * np.percentile(work, 95))
        first = order[front]          # chunk by chunk as well: the last wave of long chains then belongs to the last chunks
        if hoist_by_work:
            first = first[np.argsort(-work[first], kind="stable")]
        rest = order[~front]
        if light_first > 0:
            # the short chains of the first chunk(s) start together with the long ones, so that those chunks are
            # complete — and on their way home — as soon as their own long chains are
            early = plan_all["chunk_of"][rest] < light_first
            first = np.concatenate([rest[early], first])
            rest = rest[~early]
        plan_all["order"] = np.ascontiguousarray(np.concatenate([first, rest]), np.int32)
    # launch classes of w2t_sort_track ([wide | mid | narrow], w2t_sort_plan_t)
    plan_all["order"], plan_all["n_wide"], plan_all["n_mid"] = launch_classes(plan_all, plan_all["order"])
    _trace("plans + nms queued")
    for cs in comp + [s_in]:
        main.wait_stream(cs)
    done = torch.zeros(len(chunks), dtype=i32, device=device)
    plan_all["chunk_done"] = done
    d_offsets = _dev(h_offsets, np.int32, device)
    counters_zeroed = torch.cuda.Event()     # recorded BEFORE the launch: the finalize stream must not wait for the kernel
    counters_zeroed.record(main)
    trk = sort_track_device(S, NC, d_offsets, d_goff[:-1], nms_out["trk_count"], nms_out["trk_box"],
                            nms_out["img_exists"], d_cam, iou_thresholds, max_age, min_hits, plan_all, out=trk_out,
                            promotion=promotion)
    _trace("sort queued", main)
    keep_alive, fin_done, rows_of, h_totals = [trk["_keepalive"]], [], [], _pinned_pool("pipe_totals", torch.int64, 3 * len(chunks))
    prev_totals = None
    with torch.cuda.stream(fin):
        fin.wait_event(counters_zeroed)
        for k, (s0, s1) in enumerate(chunks):
            img0, img1 = int(h_offsets[s0]), int(h_offsets[s1])
            g0, g1 = img0 * NC, img1 * NC
            ns = s1 - s0
            d_loc = _dev((h_offsets[s0:s1 + 1] - h_offsets[s0]).astype(np.int32), np.int32, device)
            check(lib().w2t_stream_wait_value32(C.c_void_p(fin.cuda_stream), _ptr(done[k:]), ns * NC),
                  "w2t_stream_wait_value32")
            trk_k = {"out_box": trk_out["out_box"], "out_score": trk_out["out_score"], "out_birth": trk_out["out_birth"],
                     "out_count": trk_out["out_count"][g0:g1], "created": trk_out["created"][g0:g1],
                     "first_img": trk_out["first_img"][s0 * NC:s1 * NC]}
            # ids relative to id_base all along the chain (the compact rows hold them as int32; HostRows adds id_base)
            rows_k = finalize_device(ns, NC, d_loc, d_goff[g0:g1], trk_k, None, 0,
                                     int(go_np[g1] - go_np[g0]), image_base=img0, birth_group_base=g0,
                                     id_base_device=None if prev_totals is None else prev_totals[2:], compact=True)
            prev_totals = rows_k["totals"]
            h_totals[3 * k:3 * k + 3].copy_(rows_k["totals"], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(fin)
            fin_done.append(ev)
            keep_alive += [rows_k["_keepalive"], d_loc]
            rows_of.append(rows_k)
            _trace("finalize queued %d" % k, fin)

    # 3. results go home chunk by chunk, each as soon as its finalize is done, while later chunks compute
    n_rows_total, d2h_bytes = 0, 0
    for k in range(len(chunks)):
        _trace("drain wait %d" % k)
        fin_done[k].synchronize()
        n_k = int(h_totals[3 * k + 1])
        with torch.cuda.stream(s_out):
            s_out.wait_event(fin_done[k])
            for key in _COMPACT_ROW_KEYS:
                src = rows_of[k][key][:n_k]
                width = src.shape[1] if src.dim() > 1 else 1
                pool = _pinned_pool("pipe_" + key, src.dtype, (n_rows_total + n_k) * width, n_rows_total * width,
                                    quiesce=s_out.synchronize)
                pool[n_rows_total * width:(n_rows_total + n_k) * width].view(src.shape).copy_(src, non_blocking=True)
                d2h_bytes += src.numel() * src.element_size()
            _trace("d2h queued %d" % k, s_out)
        n_rows_total += n_k
    created_total = int(h_totals[3 * (len(chunks) - 1) + 2])
    for cs in comp + [fin, s_in]:
        main.wait_stream(cs)
    h_status = _host(torch.stack([nms_out["status"][0], trk_out["status"][0]]), "pipe_status")
    main.wait_stream(s_out)
    main.synchronize()
    _trace("all done")
    if int(h_status[1]) == _abi.W2T_ERR_CAPACITY:
        return ensemble_and_track(group_offsets, rows, stream_img_offsets, cam_wh, NC, iou_thresh, soft_nms_cut,
                                  min_score, score_thr, iou_thresholds, max_age, min_hits, max_group, False, id_base,
                                  raw=False, promotion=promotion)
    check_device_status(int(h_status[0]), "soft-NMS")
    check_device_status(int(h_status[1]), "SORT")
    res = HostRows({"n_rows": n_rows_total, "id_next": int(id_base + created_total), "id_base": int(id_base),
           "d2h_bytes": d2h_bytes + 24 * len(chunks) + 8,
           "launches": len(chunks) * (nms_launch_count(max(G // len(chunks), 1), max_group) + FINALIZE_LAUNCHES)
           + sort_launch_count(plan_all, S * NC), "n_chunks": len(chunks)})
    for key in _COMPACT_ROW_KEYS:
        pool = _PINNED[("pipe_" + key, {"rows_box": f64, "rows_score": f64, "rows_compact": i32}[key])]
        width = {"rows_box": 4, "rows_compact": 2}.get(key, 1)
        res[key] = pool[:n_rows_total * width].view((n_rows_total, width) if width > 1 else (n_rows_total,)).numpy()
    return res


# ---------------------------------------------------------------------------
# frame-by-frame tracking with device-resident state (w2t_sort_step)
# ---------------------------------------------------------------------------

class SortStepper:
    """Device-resident state of ``n_streams`` multi-category SORT trackers, advanced one image per
    stream and call: what a ``MultiClassTrackerSort`` per stream holds in the reference
    (tracking/sort/tracker_sort.py:10-51), except that filters, lists and counters never leave the
    GPU between calls — a call is one small H2D copy, ONE launch of the tracker kernel
    (``w2t_sort_step``) and one small D2H copy.

    ``track_cap`` / ``det_cap``: most trackers alive / detections per image and category (fixed
    slab capacities; exceeding them raises ``W2TError``).  Object ids follow
    ``KalmanBoxTracker.count`` (sort.py:86,140-141) in the order the reference would create the
    trackers: stream by stream within a call, category by category in each stream's
    first-appearance order, unmatched detections in the order of sort.py:208-222.
    """

    def __init__(self, iou_thresholds, max_age=1, min_hits=0, n_streams=1, track_cap=256, det_cap=256, id_base=0,
                 promotion=None):
        self.device = require_cuda()
        self.promotion = promotion_code(promotion)
        self.S, self.NC = int(n_streams), len(iou_thresholds)
        if not 1 <= self.NC <= _abi.W2T_MAX_CLASSES:
            raise IndexError("iou_thresholds must hold 1..%d categories" % _abi.W2T_MAX_CLASSES)
        self.iou_thresholds = [float(t) for t in iou_thresholds]
        self.max_age, self.min_hits = int(max_age), int(min_hits)
        self.track_cap, self.det_cap = int(track_cap), int(det_cap)
        nq = self.S * self.NC
        slab = int(lib().w2t_sort_slab_bytes(self.track_cap, self.det_cap))
        self._plan = {
            "order": _dev(np.arange(nq, dtype=np.int32), np.int32, self.device),
            "track_cap": _dev(np.full(nq, self.track_cap, np.int32), np.int32, self.device),
            "det_cap": _dev(np.full(nq, self.det_cap, np.int32), np.int32, self.device),
            "ws_offset": _dev(np.arange(nq, dtype=np.int64) * slab, np.int64, self.device),
        }
        self._n_wide = nq if self.det_cap > _abi.W2T_WIDE_DETS else 0
        self._ws_bytes = nq * slab
        self._workspace = torch.empty(max(self._ws_bytes, 256), dtype=torch.uint8, device=self.device)
        self._sub_state = torch.zeros((nq, 4), dtype=torch.int32, device=self.device)
        self._cam = torch.zeros((self.S, 2), dtype=torch.float64, device=self.device)   # unused by the raw rows
        self.count = int(id_base)            # KalmanBoxTracker.count
        self.calls = 0
        # [call - _base_lo, group]: id of the first tracker the group created at that call; calls no live tracker
        # refers to any more are dropped (the kernel reports the oldest birth still alive), so the table stays as
        # long as the oldest live track, not as long as the run
        self._bases = np.zeros((64, nq), np.int64)
        self._base_lo = 0
        self.class_order = [[] for _ in range(self.S)]   # categories (0-based) in first-appearance order

    def step(self, boxes, exists=None):
        """``boxes[s][c]``: float [D,4] array (x1, y1, x2, y2; rounded to float32 like tracker_sort.py:45)
        of stream ``s``, category index ``c`` (0-based) — or ``None`` / empty; ``boxes[s]`` may also be a
        ``(rows[D,>=4], categories[D])`` pair in detection order, which also fixes the order in which
        categories that appear for the first time get their trackers (tracker_sort.py:28-36).
        ``exists[s]`` False = stream ``s`` has no image in this call (nothing is stepped, like a frame
        absent from the input).  Returns ``out[s] = {c: ndarray[m,6] = x1, y1, x2, y2, id, confidence}``
        for every category that has a ``Sort`` object, rows in the reference's (newest first) order."""
        S, NC = self.S, self.NC
        if len(boxes) != S:
            raise ValueError("expected detections for %d streams" % S)
        per = [[None] * NC for _ in range(S)]
        for s in range(S):
            b = boxes[s]
            if isinstance(b, tuple):
                cats = np.asarray(b[1], np.int64).reshape(-1)
                rows = np.asarray(b[0], np.float64)
                rows = rows.reshape(len(cats), rows.shape[-1] if rows.ndim > 1 and len(cats) else 4)
                for c in cats.tolist():
                    if not 0 <= c < NC:
                        raise IndexError("category index %d outside the threshold list" % c)
                    if c not in self.class_order[s] and (exists is None or exists[s]):
                        self.class_order[s].append(c)
                for c in range(NC):
                    per[s][c] = rows[cats == c, :4]
            else:
                for c in range(NC):
                    a = None if b is None or b[c] is None else np.asarray(b[c], np.float64).reshape(-1, np.shape(b[c])[-1] if np.ndim(b[c]) > 1 else 4)[:, :4]
                    per[s][c] = a
                    if a is not None and len(a) and c not in self.class_order[s] and (exists is None or exists[s]):
                        self.class_order[s].append(c)
        counts = np.array([[0 if per[s][c] is None else len(per[s][c]) for c in range(NC)] for s in range(S)], np.int32)
        if counts.max(initial=0) > self.det_cap:
            raise W2TError("%d detections of one category in an image exceed det_cap=%d" % (int(counts.max()), self.det_cap))
        N = int(counts.sum())
        start = np.concatenate([[0], np.cumsum(counts.reshape(-1))[:-1]]).astype(np.int32)
        det = np.zeros((max(N, 1), 4), np.float32)
        for s in range(S):
            for c in range(NC):
                if counts[s, c]:
                    o = start[s * NC + c]
                    det[o:o + counts[s, c]] = per[s][c].astype(np.float32)
        ex = np.ones(S, np.uint8) if exists is None else np.asarray(exists, np.uint8)
        dev = self.device
        d_offsets = _dev(np.arange(S + 1, dtype=np.int32), np.int32, dev)
        d_start, d_count = _dev(start, np.int32, dev), _dev(counts.reshape(-1), np.int32, dev)
        d_box, d_ex = _dev(det, np.float32, dev), _dev(ex, np.uint8, dev)
        G = S * NC
        out = {"out_box": torch.zeros((max(N, 1), 4), dtype=torch.float64, device=dev),
               "out_score": torch.zeros(max(N, 1), dtype=torch.float64, device=dev),
               "out_birth": torch.zeros((max(N, 1), 2), dtype=torch.int32, device=dev),
               "out_count": torch.zeros(G, dtype=torch.int32, device=dev),
               "created": torch.zeros(G, dtype=torch.int32, device=dev),
               "first_img": torch.zeros(G, dtype=torch.int32, device=dev),   # step mode: oldest live birth group
               "status": torch.zeros(1, dtype=torch.int32, device=dev)}
        if (self.calls + 1) * G >= 2 ** 31:
            raise W2TError("SortStepper: birth records are int32 (call * groups); %d calls of %d groups reached" % (self.calls, G))
        prob = _abi.SortProblem()
        prob.n_streams, prob.n_classes = S, NC
        prob.stream_img_offsets = _ptr(d_offsets)
        prob.det_start, prob.det_count, prob.det_box = _ptr(d_start), _ptr(d_count), _ptr(d_box)
        prob.img_exists, prob.cam_wh = _ptr(d_ex), _ptr(self._cam)
        for i in range(NC):
            prob.iou_thr[i] = self.iou_thresholds[i]
        prob.max_age, prob.min_hits = self.max_age, self.min_hits
        prob.promotion = self.promotion
        cplan = _abi.SortPlan()
        for k in ("order", "track_cap", "det_cap", "ws_offset"):
            setattr(cplan, k, _ptr(self._plan[k]))
        cplan.ws_bytes, cplan.n_wide = self._ws_bytes, self._n_wide
        cplan.n_mid, cplan.aux_offset, cplan.narrow_cap = 0, -1, 0
        res = _abi.SortResult()
        for k in ("out_box", "out_score", "out_birth", "out_count", "created", "first_img"):
            setattr(res, k, _ptr(out[k]))
        check(lib().w2t_sort_step(C.byref(prob), C.byref(cplan), C.byref(res), _ptr(self._workspace),
                                  _ptr(self._sub_state), int(self.calls * G), _ptr(out["status"]), _stream()),
              "w2t_sort_step")
        h = {k: v.cpu().numpy() for k, v in out.items()}
        check_device_status(int(h["status"][0]), "w2t_sort_step")
        # ids: KalmanBoxTracker.count advances stream by stream, category by category (dict order)
        created = h["created"].reshape(S, NC)
        base = np.zeros((S, NC), np.int64)
        for s in range(S):
            for c in self.class_order[s]:
                base[s, c] = self.count
                self.count += int(created[s, c])
        if self.calls - self._base_lo == len(self._bases):
            self._bases = np.concatenate([self._bases, np.zeros_like(self._bases)])
        self._bases[self.calls - self._base_lo] = base.reshape(-1)
        self.calls += 1
        result = []
        for s in range(S):
            tracked = {}
            for c in self.class_order[s]:
                if not ex[s]:
                    continue
                g = s * NC + c
                o, m = int(start[g]), int(h["out_count"][g])
                birth = h["out_birth"][o:o + m]
                ids = self._bases[birth[:, 0] // G - self._base_lo, birth[:, 0] % G] + birth[:, 1] + 1
                rows = np.concatenate([h["out_box"][o:o + m], ids[:, None].astype(np.float64),
                                       h["out_score"][o:o + m, None]], axis=1)
                tracked[c] = rows[::-1].copy() if m else np.empty((0, 6))
            result.append(tracked)
        # forget the id bases of calls no live tracker was born in
        oldest = min(int(h["first_img"].min()) // G if G else self.calls, self.calls)
        drop = oldest - self._base_lo
        if drop >= 64:
            keep = self.calls - oldest
            self._bases[:keep] = self._bases[drop:drop + keep].copy()
            self._base_lo = oldest
        return result


# ---------------------------------------------------------------------------
# building blocks
# ---------------------------------------------------------------------------

def iou_matrix(dets, trks):
    device = require_cuda()
    d = _dev(np.asarray(dets, np.float32).reshape(-1, 4), np.float32, device)
    t = _dev(np.asarray(trks, np.float64).reshape(-1, 4), np.float64, device)
    out = torch.zeros((d.shape[0], t.shape[0]), dtype=torch.float32, device=device)
    check(lib().w2t_iou_matrix(_ptr(d), int(d.shape[0]), _ptr(t), int(t.shape[0]), _ptr(out), _stream()),
          "w2t_iou_matrix")
    return out.cpu().numpy()


def linear_assignment(cost):
    device = require_cuda()
    cost = np.ascontiguousarray(cost, np.float32)
    D, T = cost.shape
    c = _dev(cost, np.float32, device)
    pairs = torch.zeros((max(min(D, T), 1), 2), dtype=torch.int32, device=device)
    n_pairs = torch.zeros(1, dtype=torch.int32, device=device)
    ws = torch.empty(int(lib().w2t_linear_assignment_workspace(D, T)), dtype=torch.uint8, device=device)
    check(lib().w2t_linear_assignment(_ptr(c), D, T, _ptr(pairs), _ptr(n_pairs), _ptr(ws), _stream()),
          "w2t_linear_assignment")
    k = int(n_pairs.cpu()[0])
    if k < 0:
        raise W2TError("w2t_linear_assignment: iteration budget exhausted (NaN costs?)")
    return pairs[:k].cpu().numpy().astype(int).reshape(-1, 2)


def kf_init(dets, promotion=None):
    device = require_cuda()
    d = _dev(np.asarray(dets, np.float32).reshape(-1, 4), np.float32, device)
    n = int(d.shape[0])
    x = torch.zeros((n, 7), dtype=torch.float64, device=device)
    P = torch.zeros((n, 49), dtype=torch.float64, device=device)
    check(lib().w2t_kf_init(_ptr(x), _ptr(P), _ptr(d), n, promotion_code(promotion), _stream()), "w2t_kf_init")
    return x.cpu().numpy(), P.cpu().numpy().reshape(n, 7, 7)


def kf_predict(x, P):
    device = require_cuda()
    xd = _dev(np.asarray(x, np.float64).reshape(-1, 7), np.float64, device).clone()
    n = int(xd.shape[0])
    Pd = _dev(np.asarray(P, np.float64).reshape(n, 49), np.float64, device).clone()
    boxes = torch.zeros((n, 4), dtype=torch.float64, device=device)
    check(lib().w2t_kf_predict(_ptr(xd), _ptr(Pd), _ptr(boxes), n, _stream()), "w2t_kf_predict")
    return xd.cpu().numpy(), Pd.cpu().numpy().reshape(n, 7, 7), boxes.cpu().numpy()


def kf_update(x, P, dets, promotion=None):
    device = require_cuda()
    xd = _dev(np.asarray(x, np.float64).reshape(-1, 7), np.float64, device).clone()
    n = int(xd.shape[0])
    Pd = _dev(np.asarray(P, np.float64).reshape(n, 49), np.float64, device).clone()
    d = _dev(np.asarray(dets, np.float32).reshape(n, 4), np.float32, device)
    boxes = torch.zeros((n, 4), dtype=torch.float64, device=device)
    check(lib().w2t_kf_update(_ptr(xd), _ptr(Pd), _ptr(d), _ptr(boxes), n, promotion_code(promotion), _stream()),
          "w2t_kf_update")
    return xd.cpu().numpy(), Pd.cpu().numpy().reshape(n, 7, 7), boxes.cpu().numpy()


def bbox_to_z(dets, promotion=None):
    """``convert_bbox_to_z`` (sort.py:50-62) of n float32 boxes -> [n,4]: float64 under legacy promotion (x, y
    and r are float64 computations, s a float32 product), float32 under NEP 50 — the dtypes the reference's
    ``np.array([x, y, s, r])`` has in either environment."""
    device = require_cuda()
    d = _dev(np.asarray(dets, np.float32).reshape(-1, 4), np.float32, device)
    z = torch.zeros(d.shape, dtype=torch.float64, device=device)
    code = promotion_code(promotion)
    check(lib().w2t_bbox_to_z(_ptr(d), _ptr(z), int(d.shape[0]), code, _stream()), "w2t_bbox_to_z")
    out = z.cpu().numpy()
    return out.astype(np.float32) if code == _abi.W2T_PROMOTION_NEP50 else out


def bbox_vote(nms_boxes, all_boxes, all_scores, thresh, compute_f32=False):
    """``bbox_vote`` (box_utils.py:401-430) on NumPy / torch inputs -> float64 [n,4] (``w2t_bbox_vote``)."""
    device = require_cuda()
    nb = _dev(np.asarray(nms_boxes, np.float64).reshape(-1, 4), np.float64, device)
    ab = _dev(np.asarray(all_boxes, np.float64).reshape(-1, 4), np.float64, device)
    sc = _dev(np.asarray(all_scores, np.float64).reshape(-1), np.float64, device)
    out = torch.zeros((nb.shape[0], 4), dtype=torch.float64, device=device)
    check(lib().w2t_bbox_vote(_ptr(nb), int(nb.shape[0]), _ptr(ab), _ptr(sc), int(ab.shape[0]), float(thresh),
                              1 if compute_f32 else 0, _ptr(out), _stream()), "w2t_bbox_vote")
    return out.cpu().numpy()


def x_to_bbox(x):
    """``convert_x_to_bbox`` (sort.py:65-75) of n states [n,>=4] -> float64 [n,4]."""
    device = require_cuda()
    x = np.ascontiguousarray(x, np.float64)
    x = x.reshape(-1, x.shape[-1])
    xd = _dev(x, np.float64, device)
    boxes = torch.zeros((xd.shape[0], 4), dtype=torch.float64, device=device)
    check(lib().w2t_x_to_bbox(_ptr(xd), int(xd.shape[1]), _ptr(boxes), int(xd.shape[0]), _stream()), "w2t_x_to_bbox")
    return boxes.cpu().numpy()

"""Multi-GPU sharding: one process per GPU, streams sharded by segment, no data-path collective.

(segment, camera) streams are independent (a fresh ``MultiClassTrackerSort`` per stream,
``tracking/utils.py:29``) and so are the (image, category) groups of the ensemble
(``detnet/ensemble.py:50-58``).  The only thing ranks share is the reference's process-global
track id counter (``KalmanBoxTracker.count``, ``tracking/sort/sort.py:86``), which never feeds back
into the dynamics: every rank tracks its contiguous block of segments with local ids and the ids
are rebased by an exclusive scan of the per-rank creation totals — one int64 per rank, the only
exchange.  Results are gathered to rank 0 on the host (north star: no NCCL on the data path;
``torch.distributed`` is used for the rendezvous, the scan and the gather, over gloo or NCCL).
"""
import os

import torch.distributed as dist

from ._bind import bind_rank_cores  # noqa: F401  (re-exported)


def block(n_items, rank, world):
    """Contiguous block [lo, hi) of ``n_items`` for ``rank``; sizes differ by at most one."""
    q, r = divmod(int(n_items), int(world))
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def shard_segments(predictions, rank, world):
    """This rank's segments of a ``read_data_file`` result, in the file's first-appearance order."""
    segments = list(predictions.keys())
    lo, hi = block(len(segments), rank, world)
    return {seg: predictions[seg] for seg in segments[lo:hi]}


def exclusive_scan_int(value, group=None):
    """Sum of ``value`` over the lower ranks (and the grand total): ``all_gather`` of one int per rank."""
    world = dist.get_world_size(group)
    gathered = [None] * world
    dist.all_gather_object(gathered, int(value), group=group)
    rank = dist.get_rank(group)
    return sum(gathered[:rank]), sum(gathered)


def init_from_env():
    """Join the process group ``torchrun`` describes (RANK / WORLD_SIZE / MASTER_*); returns (rank, world).
    NCCL when a GPU is visible (one rank per GPU, LOCAL_RANK selects it), gloo otherwise."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0, 1
    bind_rank_cores()
    import torch
    if not dist.is_initialized():
        if torch.cuda.is_available():
            local = int(os.environ.get("LOCAL_RANK", "0"))
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group("gloo")
    return dist.get_rank(), dist.get_world_size()


def _device_track(predictions, iou_thresholds, max_age, min_hits):
    """Track a shard on this rank's GPU with ids counted from 0; returns (rows, trackers created)."""
    from .tracking import utils as trk_utils
    from .tracking.sort import sort as sort_mod
    saved = sort_mod.KalmanBoxTracker.count
    sort_mod.KalmanBoxTracker.count = 0
    try:
        rows = trk_utils.track_all(predictions, iou_thresholds, max_age, min_hits)
        created = sort_mod.KalmanBoxTracker.count
    finally:
        sort_mod.KalmanBoxTracker.count = saved
    return rows, created


def track_all_sharded(predictions, iou_thresholds, max_age, min_hits, track_fn=None, group=None, id_base=0):
    """``tracking.utils.track_all`` over all ranks of ``group``.

    Every rank passes the same ``predictions``; rank r tracks its block of segments
    (``track_fn(shard, iou_thresholds, max_age, min_hits) -> (rows, n_created)`` with object ids
    counted from 1, default: the CUDA path), ids are rebased by the exclusive scan of ``n_created``
    and the rows are gathered in rank order.  Returns ``(rows, next_id_base)`` on rank 0 and
    ``(None, next_id_base)`` elsewhere; the rows are identical to a single-process run."""
    if track_fn is None:
        track_fn = _device_track
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    rows, created = track_fn(shard_segments(predictions, rank, world), iou_thresholds, max_age, min_hits)
    before, total = exclusive_scan_int(created, group)
    shift = id_base + before
    if shift:
        for row in rows:
            row['object_id'] = '%i' % (int(row['object_id']) + shift)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(rows, gathered, dst=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    if rank != 0:
        return None, id_base + total
    merged = []
    for part in gathered:
        merged += part
    return merged, id_base + total


def _device_track_packed(packed, iou_thresholds, max_age, min_hits):
    """Track packed streams on this rank's GPU with ids counted from 1; returns (dense rows, trackers created)."""
    from . import runtime
    res = runtime.sort_track(packed, list(iou_thresholds)[:packed.n_classes], max_age, min_hits, id_base=0, raw=False)
    return res, int(res["id_next"])


def track_arrays_sharded(dets, score_threshold, iou_thresholds, max_age, min_hits, segment_id=None, track_fn=None,
                         group=None, id_base=0):
    """The tracking CLI's work over all ranks of ``group``, on flat arrays from end to end.

    ``dets``: the input file's path (every rank parses it and packs its own block in one native call,
    ``packing.pack_track_file``) or the ``native_json.Detections`` of the whole file (``packing.pack_detections``).
    Rank r packs and tracks its block of segments
    (``track_fn(packed, iou_thresholds, max_age, min_hits) -> (rows, n_created)``: dense ``rows_*`` arrays with ids
    counted from 1; default: the CUDA path), ids are rebased by the exclusive scan of ``n_created`` and the
    ARRAYS — not lists of dicts — are gathered to rank 0 in rank order.  Returns, on rank 0,
    ``(image_ids, rows, next_id_base)`` with ``rows["rows_img"]`` indexing ``image_ids``; ``(None, None,
    next_id_base)`` elsewhere; ``rows["segments"]`` lists the segments in the order they were tracked.  The rows are
    identical to a single-process run."""
    import numpy as np
    from . import packing
    if track_fn is None:
        track_fn = _device_track_packed
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    pack = packing.pack_track_file if isinstance(dets, (str, os.PathLike)) else packing.pack_detections
    packed = pack(dets, score_threshold, len(iou_thresholds), segment_id=segment_id, segment_block=(rank, world))
    res, created = track_fn(packed, iou_thresholds, max_age, min_hits)
    before, total = exclusive_scan_int(created, group)
    n = int(res["n_rows"]) if "n_rows" in res else len(res["rows_id"])
    part = {"image_ids": packing.image_id_strings(packed), "segments": list(dict.fromkeys(seg for seg, _ in packed.streams)),
            "rows_img": np.asarray(res["rows_img"][:n], np.int32), "rows_cat": np.asarray(res["rows_cat"][:n], np.int32),
            "rows_box": np.asarray(res["rows_box"][:n], np.float64).reshape(-1, 4),
            "rows_score": np.asarray(res["rows_score"][:n], np.float64),
            "rows_id": np.asarray(res["rows_id"][:n], np.int64) + (id_base + before)}
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(part, gathered, dst=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    if rank != 0:
        return None, None, id_base + total
    image_ids, shift = [], 0
    for p in gathered:
        p["rows_img"] = p["rows_img"] + shift
        shift += len(p["image_ids"])
        image_ids += p["image_ids"]
    rows = {k: np.concatenate([p[k] for p in gathered]) for k in ("rows_img", "rows_cat", "rows_box", "rows_score", "rows_id")}
    rows["n_rows"] = len(rows["rows_id"])
    rows["segments"] = [seg for p in gathered for seg in p["segments"]]
    return image_ids, rows, id_base + total


def ensemble_arrays_sharded(files, weights, method, iou_thresh, soft_nms_cut, min_score, merge_fn=None, group=None):
    """The ensemble CLI's work over all ranks of ``group``, on flat arrays from end to end.

    ``files``: one ``native_json.Detections`` per input file (every rank parses them with the native reader).
    All ranks build the same grouping (``packing.pack_detection_files``: images sorted, categories ascending);
    rank r merges its contiguous block of images (``merge_fn(groups, method, iou_thresh, soft_nms_cut, min_score)``
    -> ``ens_count / ens_box / ens_score`` at the group offsets; default: the CUDA path) and the kept rows travel to
    rank 0 as ARRAYS.  Returns ``(image_ids, image_index, category, bbox, score)`` on rank 0 — the arguments of
    ``native_json.write_detections`` — and ``None`` elsewhere; identical to a single-process run."""
    import numpy as np
    from . import packing
    if merge_fn is None:
        from .detnet import ensemble as ens
        merge_fn = ens.merge_groups
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    groups = packing.pack_detection_files(files, weights, min_score)
    ncat = max(len(groups.category_ids), 1)
    lo, hi = block(len(groups.image_ids), rank, world)
    g0, g1 = lo * ncat, hi * ncat
    go = np.asarray(groups.group_offsets, np.int64)
    part = {"img": np.zeros(0, np.int32), "cat": np.zeros(0, np.int32), "box": np.zeros((0, 4), np.int32), "score": np.zeros(0)}
    if hi > lo:
        r0, r1 = int(go[g0]), int(go[g1])
        sizes = np.diff(go[g0:g1 + 1])
        mine = packing.PackedGroups(groups.image_ids[lo:hi], groups.category_ids, (go[g0:g1 + 1] - r0).astype(np.int32),
                                    groups.rows[r0:r1], int(sizes.max()) if len(sizes) else 0,
                                    None if groups.sub_counts is None else groups.sub_counts[g0:g1])
        res = merge_fn(mine, method, iou_thresh, soft_nms_cut, min_score)
        rows, grp = packing.valid_row_index(np.asarray(mine.group_offsets, np.int64)[:-1], res["ens_count"])
        part = {"img": (grp // ncat + lo).astype(np.int32), "cat": np.asarray(groups.category_ids, np.int32)[grp % ncat],
                "box": np.asarray(res["ens_box"])[rows].astype(np.int32), "score": np.asarray(res["ens_score"])[rows]}
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(part, gathered, dst=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    if rank != 0:
        return None
    cat = lambda k: np.concatenate([p[k] for p in gathered])
    return groups.image_ids, cat("img"), cat("cat"), cat("box"), cat("score")


def ensemble_sharded(submissions, weights, method, iou_thresh, soft_nms_cut, min_score, merge_fn=None, group=None):
    """``detnet.ensemble.ensemble_submissions`` over all ranks: images are sharded in sorted order
    (contiguous blocks, so a segment's images stay together), merged independently and gathered to
    rank 0 in rank order — the same list a single process produces."""
    if merge_fn is None:
        from .detnet import ensemble as ens
        merge_fn = ens.ensemble_submissions
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    image_ids = sorted(set(d['image_id'] for sub in submissions for d in sub))
    lo, hi = block(len(image_ids), rank, world)
    mine = set(image_ids[lo:hi])
    # the weight normalisation (ensemble.py:125-126) must see all weights; the category set only
    # matters for output order, which is ascending category either way
    part = [[d for d in sub if d['image_id'] in mine] for sub in submissions]
    rows = merge_fn(part, weights, method, iou_thresh, soft_nms_cut, min_score)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(rows, gathered, dst=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    if rank != 0:
        return None
    merged = []
    for part_rows in gathered:
        merged += part_rows
    return merged

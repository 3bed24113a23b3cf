"""Fused CLI (opt-in, not in the reference): submission JSONs -> tracker JSON in one process, without the
intermediate ensemble file.

    python -m waymo_2d_tracking_b200.pipeline SUB_A.json SUB_B.json SUB_C.json -o tracks.json \\
           --min-score=0.01 --soft-nms-cut=0.9 --max-age=2 --min-hits=0 [--ensemble-output ens.json]

It is the composition of the reference's two commands (README.md:41,54)

    python -m detnet.ensemble -m soft_nms ... -o ens.json && python tracking/track.py --input ens.json ...

and writes the same ``tracks.json`` byte for byte (and, with ``--ensemble-output``, the same ``ens.json``): the
ensemble rows are truncated to int / rounded to 5 decimals on the device exactly as ``ensemble.py:62`` writes
them, ``read_data_file``'s filters (``tracking/utils.py:79-87``) are applied there too, and the tracker sees the
streams in the order ``read_data_file`` would meet them in the ensemble file (images sorted by ``image_id``,
segments and cameras in first-appearance order, frames numerically).  Host side: native JSON reader / writer
(``csrc/json_io.cpp``), flat arrays throughout; device side: ``runtime.ensemble_and_track_pipelined`` (H2D, soft-NMS,
SORT, id scan and D2H overlapped chunk by chunk).  There is no CPU fallback.
"""
import argparse
import time
from pathlib import Path

import numpy as np

from . import _abi, native_json, packing, runtime
from .tracking.sort import sort as _sort


def _floats(text):
    return [float(item) for item in text.split(',')]


def build_parser():
    parser = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    parser.add_argument('inputs', type=str, nargs='+', help='input submission json files')
    parser.add_argument('-o', '--output', type=str, required=True, help='file to save the tracker predictions')
    parser.add_argument('--ensemble-output', type=str, help='also write the ensemble json (detnet.ensemble -o)')
    parser.add_argument('--weights', type=_floats, help='one weight per input file (default: all 1)')
    parser.add_argument('--iou-thresh', type=float, default=0.5, help='IOU threshold for merging bboxes')
    parser.add_argument('--soft-nms-cut', type=float, default=1.0, help='cutout IoU threshold for soft nms')
    parser.add_argument('--min-score', type=float, default=0, help='minimal score to keep')
    parser.add_argument("--max-age", type=int, default=1, help='SORT max-age')
    parser.add_argument("--min-hits", type=int, default=0, help='SORT min-hits')
    parser.add_argument("--score-threshold", type=_floats, default=[0.95, 0.6, 1.0, 0.9], help='score threshold to track')
    parser.add_argument("--iou-threshold", type=_floats, default=[0.01, 0.01, 1.0, 0.0], help='IOU threshold for tracking')
    return parser


def stream_layout(image_ids):
    """Sorted image ids -> the tracker's streams as ``read_data_file`` + ``track.py`` would build them from a file
    whose rows come image by image in that order: segments and, inside a segment, cameras in first-appearance
    order; frames of a stream sorted numerically.  Returns (streams, stream_img_offsets, image permutation,
    frame ids in the new order)."""
    seg_index, pair_index = {}, {}
    frames, pairs = np.zeros(len(image_ids), np.int64), np.zeros(len(image_ids), np.int64)
    pair_names = []
    for i, image_id in enumerate(image_ids):
        seg, frame, cam = image_id.split('/')
        packing.camera_size(cam)                  # KeyError for an unknown camera, like clip_xy (utils.py:21)
        s = seg_index.setdefault(seg, len(seg_index))
        key = (s, cam)
        if key not in pair_index:
            pair_index[key] = len(pair_index)
            pair_names.append((seg, cam))
        pairs[i] = pair_index[key]
        frames[i] = int(frame)
    # stream order: segment first appearance, then the pair's own first appearance (pairs are numbered that way
    # inside a segment already, segments may interleave in sorted order only if one is a prefix of another)
    seg_of_pair = np.asarray([seg_index[seg] for seg, _ in pair_names], np.int64)
    stream_order = np.lexsort((np.arange(len(pair_names)), seg_of_pair)) if len(pair_names) else np.zeros(0, np.int64)
    rank_of_pair = np.empty(len(pair_names), np.int64)
    rank_of_pair[stream_order] = np.arange(len(pair_names))
    perm = np.lexsort((frames, rank_of_pair[pairs])) if len(image_ids) else np.zeros(0, np.int64)
    counts = np.bincount(rank_of_pair[pairs], minlength=len(pair_names)) if len(image_ids) else np.zeros(0, np.int64)
    offsets = np.zeros(len(pair_names) + 1, np.int64)
    np.cumsum(counts, out=offsets[1:])
    streams = [pair_names[int(p)] for p in stream_order]
    return streams, offsets.astype(np.int32), perm, frames[perm]


def load_groups(files, weights, min_score, NC, want_rows=True):
    """The submissions grouped for the tracker's layout (group g = image * NC + category - 1, images stream by
    stream in frame order): one native call (``w2t_json_group_files``), or — for inputs that call does not cover —
    the general array packer.  Returns (sorted image ids, image permutation, stream offsets, frame ids, group
    offsets, rows, packed rows or None, largest group); ``want_rows=False``: rows may be None when packed rows exist."""
    fast = native_json.group_files(files, weights, min_score, _abi.W2T_LAYOUT_STREAMS, NC, want_rows=want_rows)
    if fast is not None:
        return (fast.image_ids, fast.image_order.astype(np.int64), fast.stream_img_offsets, fast.frame_ids,
                fast.group_offsets.astype(np.int64), fast.rows, fast.packed, fast.max_group)
    return groups_from_detections([native_json.load(f) for f in files], weights, min_score, NC)


def groups_from_detections(dets, weights, min_score, NC):
    """The general path of :func:`load_groups` on parsed files (``native_json.Detections``, one per submission):
    ``packing.pack_detection_files`` + the regrouping into the tracker's layout, in NumPy.  Same return value."""
    groups = packing.pack_detection_files(dets, weights, min_score)
    if any(not (1 <= c <= NC) for c in groups.category_ids):
        raise IndexError("list index out of range: category ids %r with %d IoU thresholds" % (groups.category_ids, NC))
    ncat = len(groups.category_ids)
    n_img = len(groups.image_ids)
    _, stream_offsets, perm, frame_ids = stream_layout(groups.image_ids)
    go = np.asarray(groups.group_offsets, np.int64)
    sizes = np.zeros((n_img, NC), np.int64)
    src_group = np.full((n_img, NC), -1, np.int64)
    for k, c in enumerate(groups.category_ids):
        sizes[:, c - 1] = np.diff(go)[perm * ncat + k]
        src_group[:, c - 1] = perm * ncat + k
    new_go = np.zeros(n_img * NC + 1, np.int64)
    np.cumsum(sizes.reshape(-1), out=new_go[1:])
    valid = src_group.reshape(-1) >= 0
    starts = go[np.where(valid, src_group.reshape(-1), 0)]
    lens = sizes.reshape(-1)
    idx = np.repeat(starts - new_go[:-1], lens) + np.arange(int(new_go[-1]))     # gather: new row -> old row
    rows = np.ascontiguousarray(groups.rows[idx]) if len(idx) else np.zeros((0, 5))
    return (groups.image_ids, perm, stream_offsets, frame_ids, new_go, rows, packing.packed_rows(rows),
            int(lens.max()) if len(lens) else 0)


def main(argv=None):
    args = build_parser().parse_args(argv)
    print(args)
    files = [Path(f) for f in args.inputs]
    weights = args.weights if args.weights else [1.0] * len(files)
    if len(weights) != len(files):
        raise ValueError("--weights needs one value per input file")
    top = max(weights)
    weights = [w / top for w in weights]                      # ensemble.py:125-126
    NC = len(args.iou_threshold)
    t0 = time.time()
    sorted_ids, perm, stream_offsets, frame_ids, new_go, rows, packed_rows, max_group = load_groups(
        files, weights, args.min_score, NC, want_rows=bool(args.ensemble_output))
    n_img = len(sorted_ids)
    starts = [sorted_ids[int(perm[int(o)])].split('/') for o in stream_offsets[:-1]]
    streams = [(seg, cam) for seg, _, cam in starts]
    cam_wh = np.asarray([packing.camera_size(c) for _, c in streams], np.float64).reshape(-1, 2)   # KeyError: utils.py:21
    for seg in dict.fromkeys(seg for seg, _ in streams):
        print(seg)
    t1 = time.time()
    common = dict(stream_img_offsets=stream_offsets, cam_wh=cam_wh, n_classes=NC, iou_thresh=args.iou_thresh,
                  soft_nms_cut=args.soft_nms_cut, min_score=args.min_score, score_thr=args.score_threshold,
                  iou_thresholds=args.iou_threshold, max_age=args.max_age, min_hits=args.min_hits, max_group=max_group,
                  id_base=_sort.KalmanBoxTracker.count)
    # utils.py:51 rebuilds the image id from the parsed frame number
    image_ids = ['%s/%i/%s' % (streams[s][0], f, streams[s][1])
                 for s in range(len(streams)) for f in frame_ids[stream_offsets[s]:stream_offsets[s + 1]].tolist()]
    if args.ensemble_output:
        res = runtime.ensemble_and_track(new_go.astype(np.int32), rows, raw=False, want_ensemble=True, **common)
        erows, grp = packing.valid_row_index(new_go[:-1], res["ens_count"])
        # the ensemble file lists images in sorted image_id order, categories ascending
        order = np.argsort(perm[grp // NC] * NC + grp % NC, kind='stable')
        native_json.write_detections(args.ensemble_output, sorted_ids, perm[grp // NC][order].astype(np.int32),
                                     (grp % NC + 1)[order].astype(np.int32),
                                     np.asarray(res["ens_box"])[erows][order].astype(np.int32),
                                     np.asarray(res["ens_score"])[erows][order])
    else:
        res = runtime.ensemble_and_track_pipelined(new_go.astype(np.int32), packed_rows if packed_rows is not None else rows,
                                                   n_chunks=min(8, max(len(streams), 1)), **common)
    _sort.KalmanBoxTracker.count = int(res["id_next"])
    print("duration: %.2fs (+ %.2fs reading and packing)" % (time.time() - t1, t1 - t0))
    n = int(res["n_rows"])
    native_json.write_tracks(args.output, image_ids, res["rows_img"][:n], res["rows_box"][:n], res["rows_score"][:n],
                             res["rows_cat"][:n], res["rows_id"][:n])
    return n


if __name__ == '__main__':
    main()

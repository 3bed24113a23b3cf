"""Drop-in for ``detnet/ensemble.py``: merge several COCO-style submission JSONs per (image, category).

    python -m waymo_2d_tracking_b200.detnet.ensemble S1.json S2.json S3.json -o out.json \\
           -m soft_nms --min-score=0.01 --soft-nms-cut=0.9 -j -1

(or, with ``waymo_2d_tracking_b200`` on ``PYTHONPATH``, literally ``python -m detnet.ensemble ...``).
Same flags, same input and output JSON, same exceptions as the reference (``ensemble.py:87-160``).
What differs is where the work happens: the reference loops over images (optionally in a process
pool) and calls torch CPU ops per (image, category); here every (image, category) group of every
image is merged by ONE launch of the CUDA kernel in ``csrc/softnms.cu`` (``w2t_softnms_groups`` /
``w2t_hardnms_groups``).  ``-j`` is validated like the reference does and otherwise unused.
There is no CPU fallback.

Deliberate, documented divergences from the reference:
* ``ensemble.py:124`` crashes with ``TypeError`` (``len(None)``) when no ``.yml`` weight file is
  given; here all weights default to 1 as the line intends;
* ``yaml.load`` without a Loader no longer exists; ``yaml.safe_load`` is used;
* the reference iterates a ``set`` of image-id strings (``ensemble.py:83,147``), so its output
  order changes from run to run (hash randomisation); here images are emitted in sorted order and
  categories in ascending order, which is one of the orders the reference can produce;
* ties between equal ``score*weight`` inside a group follow the canonical rule of SURVEY.md §8c
  (the reference's unstable sort leaves them implementation-defined).
``-m weighted_fusion`` (the reference's default) runs ``detnet.nn.fusion`` on the device as well.
"""
import argparse
import json
import numbers
from collections import defaultdict
from functools import partial
from pathlib import Path

import numpy as np

try:
    from .. import _abi, native_json, packing, runtime, sharding
except ImportError:                     # imported as top-level `detnet` (PYTHONPATH=.../waymo_2d_tracking_b200)
    import os as _os
    import sys as _sys
    _sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
    from waymo_2d_tracking_b200 import _abi, native_json, packing, runtime, sharding
from .nn.tta import merge_detections, nms_detections
from .trainer.utils import get_num_workers

# module globals the reference's ensemble() reads (ensemble.py:56,60); main() sets them
args = None
merge_func = None

METHODS = ("weighted_fusion", "nms", "soft_nms")


def lxly2cxcy(bbox):
    """[score, left, top, width, height] -> [score, cx, cy, width, height], in place (ensemble.py:19-22)."""
    bbox[:, 1:3] += bbox[:, 3:5] / 2
    return bbox


def cxcy2lxly(bbox):
    """Inverse of :func:`lxly2cxcy`, in place (ensemble.py:25-28)."""
    bbox[:, 1:3] -= bbox[:, 3:5] / 2
    return bbox


def convert_submission(det_list, weight, min_score=0):
    """Submission list -> ``{image_id: {category_id: [[score*weight, x, y, w, h], ...]}}``
    (ensemble.py:31-47): zero-size boxes and rows below ``min_score`` are dropped."""
    grouped = defaultdict(lambda: defaultdict(list))
    for det in det_list:
        x, y, w, h = det['bbox']
        if not (w > 0 and h > 0):
            continue
        weighted = det['score'] * weight
        if weighted >= min_score:
            grouped[det['image_id']][det['category_id']].append([weighted, x, y, w, h])
    return grouped


def _method_of(func):
    """(method, iou_thresh, soft_nms_cut) if ``func`` is a partial over this package's merge functions."""
    if isinstance(func, partial) and func.func is nms_detections and not func.args:
        kw = func.keywords
        return ("soft_nms" if kw.get('soft', False) else "nms"), kw.get('iou_thresh', 0.5), kw.get('soft_nms_cut', 1)
    if isinstance(func, partial) and func.func is merge_detections and not func.args:
        return "weighted_fusion", func.keywords.get('nms_thresh', 0.5), 1
    return None


def merge_groups(groups, method, iou_thresh, soft_nms_cut, min_score):
    """All groups of ``groups`` (``packing.PackedGroups``) in one launch -> host result arrays
    (``ens_count``, ``ens_box``, ``ens_score`` per ``include/w2t_types.h``)."""
    if method == "weighted_fusion":
        from .nn import fusion
        return fusion.fuse_groups(groups, iou_thresh, min_score)
    # soft-NMS takes the 8-byte rows when the packer proved them exact (5x less to ship, integer ranking keys)
    rows = groups.packed if (method == "soft_nms" and getattr(groups, "packed", None) is not None) else groups.rows
    return runtime.softnms_groups(groups.group_offsets, rows, iou_thresh, soft_nms_cut, min_score,
                                  max_group=groups.max_group, want_merged=False, box_format=_abi.W2T_BOX_LTWH,
                                  hard=(method == "nms"))


def rows_to_json(groups, res):
    """Result arrays -> the reference's output list (ensemble.py:59-63), image by image,
    category by category, rows in descending original-score order."""
    ncat = len(groups.category_ids)
    offs = np.asarray(groups.group_offsets, np.int64)
    rows, grp = packing.valid_row_index(offs[:-1], res["ens_count"])
    boxes = res["ens_box"][rows].tolist()
    scores = res["ens_score"][rows].tolist()
    image_of = (grp // ncat).tolist()
    cat_of = (grp % ncat).tolist()
    ids, cats = groups.image_ids, groups.category_ids
    return [{'image_id': ids[i], 'category_id': cats[c], 'bbox': b, 'score': s}
            for i, c, b, s in zip(image_of, cat_of, boxes, scores)]


def ensemble(image_id, detections, category_ids):
    """One image (ensemble.py:50-64): ``detections`` holds that image's ``{category: rows}`` dict
    of every submission.  Reads the module globals ``args.min_score`` and ``merge_func`` like the
    reference.  When ``merge_func`` is a partial over this package's merge functions all categories
    of the image go to the device in one launch; any other callable is applied per category."""
    category_ids = list(category_ids)
    known = _method_of(merge_func)
    if known is not None:
        method, iou_thresh, cut = known
        groups = packing.pack_submissions([{image_id: det} for det in detections], [image_id], category_ids)
        return rows_to_json(groups, merge_groups(groups, method, iou_thresh, cut, args.min_score))
    out = []
    for category_id in category_ids:
        per_file = [lxly2cxcy(np.asarray(det[category_id], dtype=np.float64).reshape(-1, 5)) for det in detections]
        for row in cxcy2lxly(merge_func(per_file)):
            if row[0] > args.min_score:
                out.append({'image_id': image_id, 'category_id': category_id,
                            'bbox': row[1:].astype(int).tolist(), 'score': round(row[0], 5)})
    return out


def load_yml_input_and_weight(input_files_with_weights, prefix=''):
    """Nested ``{dir: {file: weight}}`` mapping -> flat ``[(path, weight)]`` (ensemble.py:67-75)."""
    flat = []
    for name, value in input_files_with_weights.items():
        path = prefix + '/' + name if prefix else name
        if isinstance(value, numbers.Number):
            flat.append((path, value))
        else:
            flat.extend(load_yml_input_and_weight(value, path))
    return flat


def load_input_submissions(input_files, input_weights, min_score=None):
    """ensemble.py:78-84: returns (image ids, category ids, one ``convert_submission`` dict per file).
    ``min_score`` defaults to the module global ``args.min_score`` like the reference."""
    if min_score is None:
        min_score = args.min_score
    submissions = [json.load(Path(f).open()) for f in input_files]
    category_ids = set(d['category_id'] for sub in submissions for d in sub)
    converted = [convert_submission(sub, w, min_score) for sub, w in zip(submissions, input_weights)]
    image_ids = set(k for det in converted for k in det.keys())
    return image_ids, category_ids, converted


def pack_submission_lists(submissions, weights, min_score):
    """Vectorised ``convert_submission`` + grouping for the CLI: list-of-dict submissions ->
    ``packing.PackedGroups`` (images sorted, categories ascending), without nested dicts."""
    category_ids = sorted(set(d['category_id'] for sub in submissions for d in sub))
    cat_index = {c: i for i, c in enumerate(category_ids)}
    cols = []
    for sub, w in zip(submissions, weights):
        n = len(sub)
        box = np.asarray([d['bbox'] for d in sub], dtype=np.float64).reshape(n, 4)
        score = np.asarray([d['score'] for d in sub], dtype=np.float64).reshape(n) * np.float64(w)
        keep = (box[:, 2] > 0) & (box[:, 3] > 0) & (score >= min_score)
        idx = np.nonzero(keep)[0]
        cols.append(([sub[i]['image_id'] for i in idx],
                     np.asarray([cat_index[sub[i]['category_id']] for i in idx], np.int64),
                     np.concatenate([score[idx, None], box[idx]], axis=1)))
    image_ids = sorted(set(i for c in cols for i in c[0]))
    img_index = {s: i for i, s in enumerate(image_ids)}
    ncat = max(len(category_ids), 1)
    G = len(image_ids) * ncat
    per_file_keys = [np.asarray([img_index[s] for s in c[0]], np.int64) * ncat + c[1] for c in cols]
    keys = np.concatenate(per_file_keys) if cols else np.zeros(0, np.int64)
    rows = np.concatenate([c[2] for c in cols]) if cols else np.zeros((0, 5))
    order = np.argsort(keys, kind='stable')          # keeps (file, JSON) order inside a group (tta.py:9-12)
    counts = np.bincount(keys, minlength=G) if G else np.zeros(0, np.int64)
    offsets = np.zeros(G + 1, np.int64)
    np.cumsum(counts, out=offsets[1:])
    sub_counts = np.stack([np.bincount(k, minlength=G) for k in per_file_keys], axis=1).astype(np.int32) \
        if G else np.zeros((0, max(len(cols), 1)), np.int32)
    return packing.PackedGroups(image_ids, category_ids, offsets.astype(np.int32),
                                np.ascontiguousarray(rows[order]), int(counts.max()) if G else 0, sub_counts)


def ensemble_submissions(submissions, weights=None, method="soft_nms", iou_thresh=0.5, soft_nms_cut=1.0,
                         min_score=0.0):
    """Library form of the CLI: list-of-dict submissions in, merged list of dicts out."""
    if method not in METHODS:
        raise ValueError("method must be one of %s" % (METHODS,))
    if not weights:
        weights = [1] * len(submissions)
    top = max(weights)
    weights = [w / top for w in weights]
    groups = pack_submission_lists(submissions, weights, min_score)
    if len(groups.image_ids) == 0:
        return []
    return rows_to_json(groups, merge_groups(groups, method, iou_thresh, soft_nms_cut, min_score))


def build_parser():
    parser = argparse.ArgumentParser(description="ensemble submission together",
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter,
                                     fromfile_prefix_chars='@')
    parser.add_argument('inputs', type=str, nargs='+', help='input json files')
    parser.add_argument('-o', '--output', type=str, help='output json file')
    parser.add_argument('-m', '--method', choices=METHODS, default="weighted_fusion",
                        help='method to merge bbox detections')
    parser.add_argument('--iou-thresh', type=float, default=0.5, help='IOU threshold for merging bboxes')
    parser.add_argument('--soft-nms-cut', type=float, default=1.0, help='cutout IoU threshold for soft nms')
    parser.add_argument('--min-score', type=float, default=0, help='minimal score to keep')
    parser.add_argument('-j', '--jobs', type=int, default=1, help='number of workers')
    return parser


def main(argv=None):
    global args, merge_func
    import yaml
    args = build_parser().parse_args(argv)

    input_files = []
    for name in args.inputs:
        path = Path(name)
        if path.is_file():
            input_files.append(path)
        elif path.is_dir():
            input_files.extend(path.glob("**/*.json"))
        else:
            print(f"{path} is neither file nor dir?!")

    input_weights = None
    if len(input_files) == 1 and input_files[0].suffix == '.yml':
        listed = load_yml_input_and_weight(yaml.safe_load(input_files[0].open()))
        input_files, input_weights = zip(*listed)
        print(input_files, input_weights)

    assert len(input_files) > 1
    print('input files:', input_files)

    if not input_weights:
        input_weights = [1] * len(input_files)
    top = max(input_weights)
    input_weights = [w / top for w in input_weights]
    print('weights', input_weights)

    output_file = Path(args.output)
    output_file.parent.mkdir(parents=True, exist_ok=True)
    if output_file.exists():
        raise RuntimeError(f"output file {output_file} exists!")

    get_num_workers(args.jobs)   # same validation as the reference; the device path needs no workers
    merge_func = partial(merge_detections, nms_thresh=args.iou_thresh)
    if args.method == 'nms':
        merge_func = partial(nms_detections, iou_thresh=args.iou_thresh)
    elif args.method == 'soft_nms':
        merge_func = partial(nms_detections, iou_thresh=args.iou_thresh, soft=True, soft_nms_cut=args.soft_nms_cut)

    import os
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        # launched by torchrun: every rank parses the files natively and merges its block of images; the kept rows
        # travel to rank 0 as arrays, which writes the file
        sharding.init_from_env()
        out = sharding.ensemble_arrays_sharded([native_json.load(f) for f in input_files], list(input_weights),
                                               args.method, args.iou_thresh, args.soft_nms_cut, args.min_score)
        if out is None:
            return None
        image_ids, img, cat, box, score = out
        print('No. Images:', len(image_ids))
        native_json.write_detections(output_file, image_ids, img, cat, box, score)
        return len(score)
    # native reader / writer: files <-> flat arrays, no per-detection Python objects
    groups = packing.pack_files(input_files, input_weights, args.min_score)
    print('No. Images:', len(groups.image_ids))
    print('No. categories:', len(groups.category_ids))
    if not len(groups.image_ids):
        native_json.write_detections(output_file, [], [], [], np.zeros((0, 4), np.int32), [])
        return 0
    res = merge_groups(groups, args.method, args.iou_thresh, args.soft_nms_cut, args.min_score)
    ncat = len(groups.category_ids)
    rows, grp = packing.valid_row_index(np.asarray(groups.group_offsets, np.int64)[:-1], res["ens_count"])
    native_json.write_detections(output_file, groups.image_ids, grp // ncat,
                                 np.asarray(groups.category_ids, np.int32)[grp % ncat], res["ens_box"][rows],
                                 res["ens_score"][rows])
    return len(rows)


if __name__ == '__main__':
    main()

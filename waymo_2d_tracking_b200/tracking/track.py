"""Drop-in for ``tracking/track.py`` (the SORT CLI, README.md:54):

    python -m waymo_2d_tracking_b200.tracking.track --ground-truth GT/images.json --input ENS.json \\
           --output tracks.json --max-age=2 --min-hits=0 --score-threshold=0.95,0.6,1.0,0.9

Same flags, same input / output JSON, same prints (the parsed args, every segment id, the
duration of the tracking loop) as ``track.py:13-50``.  The reference tracks one (segment, camera)
stream after the other in Python; here all streams are tracked by one launch of the persistent
CUDA kernel, so the per-segment prints precede the single call.  Under ``torchrun`` (WORLD_SIZE > 1)
segments are sharded over the ranks / GPUs — every rank parses the file with the native reader, packs and tracks its own
block of segments, and rank 0 gathers the result ARRAYS and writes the output (``sharding.track_arrays_sharded``).  The ground-truth file is loaded
(and must exist) like in the reference although its content is not used (``track.py:32-35``).
"""
import argparse
import json
import os
import time
import warnings
from os.path import dirname, join

try:
    from .utils import native_json, packing, read_data_file, sharding, track_packed
except ImportError:                     # run as a script from inside tracking/, like the reference
    from utils import native_json, packing, read_data_file, sharding, track_packed

warnings.simplefilter(action='ignore', category=FutureWarning)


def _floats(text):
    return [float(item) for item in text.split(',')]


def build_parser():
    parser = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    parser.add_argument("--ground-truth", type=str, default='/data/waymo/det2d/validation/images.json',
                        help='ground-truth json')
    parser.add_argument("--input", type=str, default='submission_12545.json', help='submission.json')
    parser.add_argument("--output", type=str, default='tracker_predictions.json',
                        help='file to save the tracker predictions')
    parser.add_argument("--max-age", type=int, default=1, help='SORT max-age')
    parser.add_argument("--min-hits", type=int, default=0, help='SORT min-hits')
    parser.add_argument("--score-threshold", type=_floats, default=[0.95, 0.6, 1.0, 0.9],
                        help='score threshold to track')
    parser.add_argument("--iou-threshold", type=_floats, default=[0.01, 0.01, 1.0, 0.0],
                        help='IOU threshold for tracking')
    parser.add_argument("--segment-id", type=str, help='track only a single segment')
    return parser


def main(argv=None):
    args = build_parser().parse_args(argv)
    print(args)

    # native reader: the file goes straight into flat arrays (no per-detection dicts); the grouping, filters and
    # ordering of read_data_file (utils.py:63-96) are applied by packing.pack_track_file (one native call: parse +
    # pack) or, one process per GPU, by packing.pack_detections on each rank's block of segments
    multi = int(os.environ.get("WORLD_SIZE", "1")) > 1
    n_classes = len(args.iou_threshold)
    if not multi:
        packed = packing.pack_track_file(args.input, args.score_threshold, n_classes, segment_id=args.segment_id or None)
    image_id2path = {}
    ground_truth_dir = dirname(args.ground_truth)
    with open(args.ground_truth) as fp:
        for image in json.load(fp):
            image_id2path[image['id']] = join(ground_truth_dir, image['file_name'])

    if multi:
        # launched by torchrun: one rank per GPU, each packs and tracks its block of segments from the flat arrays,
        # rank 0 gathers ARRAYS and writes the file
        sharding.init_from_env()
        start_time = time.time()
        image_ids, rows, _ = sharding.track_arrays_sharded(args.input, args.score_threshold, args.iou_threshold,
                                                           args.max_age, args.min_hits, segment_id=args.segment_id or None)
        if rows is None:
            return None
        for segment_id in rows["segments"]:
            print(segment_id)
        print("duration: %.2fs" % (time.time() - start_time))
        native_json.write_tracks(args.output, image_ids, rows["rows_img"], rows["rows_box"], rows["rows_score"],
                                 rows["rows_cat"], rows["rows_id"])
        return int(rows["n_rows"])

    start_time = time.time()
    for segment_id in dict.fromkeys(seg for seg, _ in packed.streams):
        print(segment_id)
    res = track_packed(packed, args.iou_threshold, args.max_age, args.min_hits)
    print("duration: %.2fs" % (time.time() - start_time))
    native_json.write_tracks(args.output, packing.image_id_strings(packed), res["rows_img"], res["rows_box"],
                             res["rows_score"], res["rows_cat"], res["rows_id"])
    return int(res["n_rows"])


if __name__ == '__main__':
    main()

"""Drop-in for ``coco_to_waymo.py`` of the reference (:16-107): detection / tracking JSON -> serialized
``waymo_open_dataset`` ``Submission`` protobuf, the consumer of both output files of the path
(README.md:46,55 of the reference).  Same CLI flags, same file.

The reference builds the message with the generated classes of the ``waymo_open_dataset`` package;
that package is not available here, so the bytes are written by a hand-rolled encoder
(``csrc/waymo_pb.cpp`` behind ``w2t_pb_write_submission``) from the package's published ``.proto``
files.  PARITY UNPINNED at the schema (field numbers restated, not checkable here); given the
schema the bytes equal the protobuf runtime's (``tests/test_waymo_pb.py``).
"""
import argparse
import ctypes as C
import json
import os
from pathlib import Path

import numpy as np

from ._lib import check, lib

# submission.proto / dataset.proto / label.proto enum values used by the reference
TASK_DETECTION_2D, TASK_TRACKING_2D = 1, 3
SENSOR_CAMERA_ALL = 3
CAMERA_NAMES = {'UNKNOWN': 0, 'FRONT': 1, 'FRONT_LEFT': 2, 'FRONT_RIGHT': 3, 'SIDE_LEFT': 4, 'SIDE_RIGHT': 5}
LABEL_TYPES = (1, 2, 3, 4)     # TYPE_VEHICLE, TYPE_PEDESTRIAN, TYPE_SIGN, TYPE_CYCLIST; 0 = TYPE_UNKNOWN
AUTHORS = ('Yuan Xu', 'Erdene-Ochir Tuguldur')      # coco_to_waymo.py:69-70
AFFILIATION = 'DAInamite'


def _rows_to_arrays(detections):
    """List of row dicts (ensemble.py:61-62 / tracking/utils.py:52-58) -> flat arrays; the exceptions of
    create_pd_object (coco_to_waymo.py:16-50) are raised here, before anything is written."""
    image_ids, index = [], {}
    n = len(detections)
    image = np.empty(n, np.int32)
    bbox = np.empty((n, 4), np.float64)
    score = np.empty(n, np.float64)
    category = np.empty(n, np.int32)
    tracked = [('object_id' in d) for d in detections]
    object_id = np.zeros(n, np.int64) if any(tracked) else None
    if object_id is not None and not all(tracked):
        raise ValueError("either every row carries an object_id or none does")
    for i, d in enumerate(detections):
        iid = d['image_id']
        k = index.get(iid)
        if k is None:
            context_name, frame_timestamp_micros, camera_name = iid.split('/')   # ValueError unless two slashes
            int(frame_timestamp_micros)
            if camera_name not in CAMERA_NAMES:
                raise ValueError('Enum CameraName.Name has no value defined for name %r' % camera_name)
            k = index[iid] = len(image_ids)
            image_ids.append(iid)
        image[i] = k
        bbox[i] = d['bbox']
        score[i] = d['score']
        label = int(d['category_id'])
        if label == 0:
            raise AssertionError('TYPE_UNKNOWN')                 # coco_to_waymo.py:49
        if label not in LABEL_TYPES:
            raise ValueError('Unknown enum value: %d' % label)   # proto2 closed enum
        category[i] = label
        if object_id is not None:
            object_id[i] = int(d['object_id'])
            if '%i' % object_id[i] != str(d['object_id']):
                raise ValueError("object_id %r is not the decimal string tracking/utils.py:57 writes" % (d['object_id'],))
    return image_ids, image, bbox, score, category, object_id


def write_submission(path, image_ids, image, bbox, score, category, object_id=None, *, unique_method_name='',
                     description='', account_name='', tracking=False, objects_only=False):
    """Flat arrays (``image`` indexes ``image_ids``; the layout of ``native_json.write_*``) -> file.
    ``objects_only``: just the ``metrics.Objects`` message of the submission."""
    for name, v in (('unique_method_name', unique_method_name), ('description', description),
                    ('account_name', account_name)):
        if not objects_only and not isinstance(v, str):
            raise TypeError('%s must be a str, got %r' % (name, v))     # protobuf rejects None the same way
    enc = [s.encode('utf-8') for s in image_ids]
    table = (C.c_char_p * max(len(enc), 1))(*enc)
    authors = (C.c_char_p * len(AUTHORS))(*[a.encode('utf-8') for a in AUTHORS])
    image = np.ascontiguousarray(image, np.int32)
    bbox = np.ascontiguousarray(bbox, np.float64).reshape(-1, 4)
    score = np.ascontiguousarray(score, np.float64)
    category = np.ascontiguousarray(category, np.int32)
    oid = None if object_id is None else np.ascontiguousarray(object_id, np.int64)
    vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    check(lib().w2t_pb_write_submission(
        os.fsencode(str(path)), int(bool(objects_only)), TASK_TRACKING_2D if tracking else TASK_DETECTION_2D,
        (account_name or '').encode('utf-8'), (unique_method_name or '').encode('utf-8'), C.cast(authors, C.c_void_p),
        len(AUTHORS), AFFILIATION.encode('utf-8'), (description or '').encode('utf-8'), b'', SENSOR_CAMERA_ALL,
        len(image), C.cast(table, C.c_void_p), len(enc), vp(image), vp(bbox), vp(score), vp(category), vp(oid)),
        "w2t_pb_write_submission")


def create_pb_submission_file(detections, output, unique_method_name, description, account_name, tracking):
    """create_pb_submission (coco_to_waymo.py:64-82) + SerializeToString + write, for a list of row dicts."""
    arrays = _rows_to_arrays(detections)
    write_submission(output, *arrays, unique_method_name=unique_method_name, description=description,
                     account_name=account_name, tracking=tracking)


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('detection', type=str, nargs='+', help='detection result json file')
    parser.add_argument('--unique-method-name', type=str, required=True, help='unique method name. Max 25 chars.')
    parser.add_argument('--description', type=str, help='detailed description of method.')
    parser.add_argument('--account-name', type=str, required=True, help='email')
    parser.add_argument('--tracking', action='store_true', help='tracking submission')
    parser.add_argument('-o', '--output', type=str, help='output submission file')
    args = parser.parse_args(argv)

    detections = []
    for f in args.detection:
        detections += json.load(open(f))
    output = Path(args.output)
    output.parent.mkdir(parents=True, exist_ok=True)
    create_pb_submission_file(detections, args.output, args.unique_method_name, args.description, args.account_name,
                              args.tracking)


if __name__ == '__main__':
    main()

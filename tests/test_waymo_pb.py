"""Waymo Submission writer (csrc/waymo_pb.cpp, coco_to_waymo drop-in) against the protobuf runtime.

The `waymo_open_dataset` package is not available, so the schema is rebuilt here from the same restated
field numbers (PARITY UNPINNED at the schema, see csrc/waymo_pb.cpp); what these tests pin is everything
else: given that schema, the hand-rolled encoder writes byte for byte what `SerializeToString()` of the
messages the reference builds (coco_to_waymo.py:16-82) would write, and it parses back to the same values.
"""
import json

import numpy as np
import pytest

from waymo_2d_tracking_b200 import coco_to_waymo as c2w
from waymo_2d_tracking_b200 import generate_prediction_for_metrics as gpm

pb = pytest.importorskip("google.protobuf")
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory  # noqa: E402


def _schema():
    F = descriptor_pb2.FieldDescriptorProto
    fd = descriptor_pb2.FileDescriptorProto(name="w2t_waymo_restated.proto", package="w2t.waymo", syntax="proto2")

    def msg(parent, name, fields):
        m = parent.message_type.add() if isinstance(parent, descriptor_pb2.FileDescriptorProto) else parent.nested_type.add()
        m.name = name
        for fname, num, typ, label, type_name in fields:
            f = m.field.add(name=fname, number=num, type=typ, label=label)
            if type_name:
                f.type_name = type_name
        return m

    O, R = F.LABEL_OPTIONAL, F.LABEL_REPEATED
    label = msg(fd, "Label", [("box", 1, F.TYPE_MESSAGE, O, ".w2t.waymo.Label.Box"), ("type", 3, F.TYPE_INT32, O, None),
                              ("id", 4, F.TYPE_STRING, O, None), ("detection_difficulty_level", 5, F.TYPE_INT32, O, None),
                              ("tracking_difficulty_level", 6, F.TYPE_INT32, O, None),
                              ("num_lidar_points_in_box", 7, F.TYPE_INT32, O, None)])
    msg(label, "Box", [(n, i, F.TYPE_DOUBLE, O, None) for n, i in (("center_x", 1), ("center_y", 2), ("center_z", 3),
                                                                  ("width", 4), ("length", 5), ("height", 6), ("heading", 7))])
    msg(fd, "Object", [("object", 1, F.TYPE_MESSAGE, O, ".w2t.waymo.Label"), ("score", 2, F.TYPE_FLOAT, O, None),
                       ("overlap_with_nlz", 3, F.TYPE_BOOL, O, None), ("context_name", 4, F.TYPE_STRING, O, None),
                       ("frame_timestamp_micros", 5, F.TYPE_INT64, O, None), ("camera_name", 6, F.TYPE_INT32, O, None)])
    msg(fd, "Objects", [("objects", 1, F.TYPE_MESSAGE, R, ".w2t.waymo.Object")])
    msg(fd, "Submission", [("task", 1, F.TYPE_INT32, O, None), ("account_name", 2, F.TYPE_STRING, O, None),
                           ("unique_method_name", 3, F.TYPE_STRING, O, None), ("authors", 4, F.TYPE_STRING, R, None),
                           ("affiliation", 5, F.TYPE_STRING, O, None), ("description", 6, F.TYPE_STRING, O, None),
                           ("method_link", 7, F.TYPE_STRING, O, None), ("sensor_type", 8, F.TYPE_INT32, O, None),
                           ("number_past_frames_exclude_current", 9, F.TYPE_INT32, O, None),
                           ("number_future_frames_exclude_current", 10, F.TYPE_INT32, O, None),
                           ("inference_results", 11, F.TYPE_MESSAGE, O, ".w2t.waymo.Objects")])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName("w2t.waymo." + n))
    return get("Submission"), get("Objects"), get("Object")


def _reference_build(Submission, Objects, Object, detections, unique_method_name, description, account_name, tracking):
    """create_pb_submission of coco_to_waymo.py:64-82, statement for statement, on the restated schema
    (enums as their integer values)."""
    submission = Submission()
    submission.task = 3 if tracking else 1
    submission.account_name = account_name
    submission.authors.append('Yuan Xu')
    submission.authors.append('Erdene-Ochir Tuguldur')
    submission.affiliation = 'DAInamite'
    submission.unique_method_name = unique_method_name
    submission.description = description
    submission.method_link = ""
    submission.sensor_type = 3
    submission.number_past_frames_exclude_current = 0
    submission.number_future_frames_exclude_current = 0
    objects = Objects()
    for detection in detections:
        context_name, ts, camera_name = detection['image_id'].split('/')
        o = Object()
        o.context_name = context_name
        o.frame_timestamp_micros = int(ts)
        o.camera_name = c2w.CAMERA_NAMES[camera_name]
        bbox = detection['bbox']
        o.object.box.center_x = bbox[0] + bbox[2] * 0.5
        o.object.box.center_y = bbox[1] + bbox[3] * 0.5
        o.object.box.length = bbox[2]
        o.object.box.width = bbox[3]
        o.score = detection['score']
        if 'object_id' in detection:
            o.object.id = detection['object_id']
        o.object.type = detection['category_id']
        objects.objects.append(o)
    submission.inference_results.CopyFrom(objects)      # sets the field even when there are no objects
    return submission


def _rows(n, tracking, seed=0):
    rng = np.random.default_rng(seed)
    cams = list(c2w.CAMERA_NAMES)[1:]
    rows = []
    for i in range(n):
        seg = "%d_%d_000_%d_000" % (rng.integers(1 << 40, 1 << 62), rng.integers(0, 9999), rng.integers(0, 9999))
        row = {'image_id': "%s/%d/%s" % (seg, 1550000000000000 + 100000 * int(rng.integers(0, 200)), cams[i % 5]),
               'bbox': [float(v) for v in rng.uniform(-50, 1900, 4)] if tracking else [int(v) for v in rng.integers(0, 1900, 4)],
               'score': float(np.round(rng.uniform(0, 1), 5)) if not tracking else float(rng.uniform(0.2, 1)),
               'category_id': int(rng.integers(1, 5))}
        if tracking:
            row['object_id'] = '%i' % int(rng.integers(1, 3_000_000_000))
        rows.append(row)
    rows[0]['image_id'] = rows[1]['image_id']           # shared image ids are interned
    return rows


@pytest.mark.parametrize("tracking", [False, True])
def test_submission_bytes_equal_the_protobuf_runtime(tracking, tmp_path):
    Submission, Objects, Object = _schema()
    rows = _rows(300, tracking, seed=3 + tracking)
    rows[5]['image_id'] = "seg_with_negative_ts/-7/FRONT"      # int64 varint of a negative number: 10 bytes
    rows[6]['score'] = 1.0
    rows[7]['bbox'] = [0, 0, 0, 0] if not tracking else [0.0, 0.0, 0.0, 0.0]   # default values are still written (proto2)
    src = tmp_path / "rows.json"
    src.write_text(json.dumps(rows))
    out = tmp_path / "sub" / "submission.bin"
    argv = [str(src), "--unique-method-name", "w2t-b200", "--description", "ensemble + SORT, déjà vu", "--account-name",
            "someone@example.org", "-o", str(out)] + (["--tracking"] if tracking else [])
    c2w.main(argv)
    want = _reference_build(Submission, Objects, Object, rows, "w2t-b200", "ensemble + SORT, déjà vu", "someone@example.org", tracking)
    got = out.read_bytes()
    assert got == want.SerializeToString(deterministic=True)
    back = Submission()
    back.ParseFromString(got)
    assert len(back.inference_results.objects) == len(rows) and back.task == (3 if tracking else 1)
    o = back.inference_results.objects[11]
    assert o.context_name == rows[11]['image_id'].split('/')[0] and o.object.type == rows[11]['category_id']
    assert o.score == np.float32(rows[11]['score']) and o.object.box.length == rows[11]['bbox'][2]
    assert (o.object.id == rows[11]['object_id']) if tracking else (not o.object.HasField("id"))


def test_objects_only_and_array_entry_point(tmp_path):
    Submission, Objects, Object = _schema()
    rows = _rows(40, True, seed=9)
    image_ids, image, bbox, score, category, object_id = c2w._rows_to_arrays(rows)
    out = tmp_path / "objects.bin"
    c2w.write_submission(out, image_ids, image, bbox, score, category, object_id, objects_only=True)
    want = _reference_build(Submission, Objects, Object, rows, "", "", "", True).inference_results
    assert out.read_bytes() == want.SerializeToString(deterministic=True)
    empty = tmp_path / "empty.bin"
    c2w.write_submission(empty, [], [], np.zeros((0, 4)), [], [], None, unique_method_name="m", description="d",
                         account_name="a")
    assert empty.read_bytes() == _reference_build(Submission, Objects, Object, [], "m", "d", "a", False).SerializeToString()


def test_errors_of_the_reference(tmp_path):
    base = {'image_id': 'seg/1/FRONT', 'bbox': [1, 2, 3, 4], 'score': 0.5, 'category_id': 1}
    with pytest.raises(ValueError):
        c2w._rows_to_arrays([dict(base, image_id='seg/1')])                 # not segment/timestamp/camera
    with pytest.raises(ValueError):
        c2w._rows_to_arrays([dict(base, image_id='seg/1/REAR')])            # CameraName.Name.Value
    with pytest.raises(ValueError):
        c2w._rows_to_arrays([dict(base, image_id='seg/one/FRONT')])         # int(frame_timestamp_micros)
    with pytest.raises(AssertionError):
        c2w._rows_to_arrays([dict(base, category_id=0)])                    # TYPE_UNKNOWN
    with pytest.raises(TypeError):                                          # --description omitted: protobuf rejects None
        c2w.create_pb_submission_file([base], tmp_path / "x.bin", "m", None, "a", False)


def test_metrics_objects_file_equals_the_protobuf_runtime(tmp_path):
    """generate_prediction_for_metrics.py:44-80 statement for statement on the restated schema, predictions
    (tracker rows) and ground truth (annotation file with string ids and difficulty levels)."""
    Submission, Objects, Object = _schema()

    def reference(entries):
        objects = Objects()
        for e in entries:
            segment_id, frame_id, camera_id = e['image_id'].split('/')
            o = Object()
            o.context_name = segment_id
            o.frame_timestamp_micros = int(frame_id)
            o.camera_name = gpm.CAMERA_NAMES[camera_id]
            bbox = e['bbox']
            o.object.box.center_x = bbox[0] + bbox[2] / 2
            o.object.box.center_y = bbox[1] + bbox[3] / 2
            o.object.box.center_z = 0
            o.object.box.length = bbox[2]
            o.object.box.width = bbox[3]
            o.object.box.height = 0
            o.object.box.heading = 0
            o.object.type = e['category_id']
            if 'score' in e:
                o.score = e['score']
            if 'object_id' in e:
                o.object.id = e['object_id']
            if 'tracking_difficulty_level' in e:
                o.object.tracking_difficulty_level = e['tracking_difficulty_level']
            if 'detection_difficulty_level' in e:
                o.object.detection_difficulty_level = e['detection_difficulty_level']
            o.object.num_lidar_points_in_box = 100
            objects.objects.append(o)
        return objects.SerializeToString(deterministic=True)

    pred = _rows(120, True, seed=21)
    src = tmp_path / "pred.json"
    src.write_text(json.dumps(pred))
    out = tmp_path / "pred.bin"
    gpm.main(["--input", str(src), "--output", str(out)])
    assert out.read_bytes() == reference(pred)

    rng = np.random.default_rng(5)
    gt = [{'image_id': r['image_id'], 'bbox': [int(v) for v in rng.integers(0, 1900, 4)], 'category_id': r['category_id'],
           'object_id': "gt-%x_é" % int(rng.integers(1 << 40)), 'tracking_difficulty_level': int(rng.integers(1, 3)),
           'detection_difficulty_level': int(rng.integers(1, 3)), 'id': i} for i, r in enumerate(_rows(80, False, seed=22))]
    src = tmp_path / "annotations.json"
    src.write_text(json.dumps({'annotations': gt, 'images': []}))
    out = tmp_path / "gt.bin"
    gpm.main(["--type", "ground-truth", "--input", str(src), "--output", str(out)])
    assert out.read_bytes() == reference(gt)
    with pytest.raises(KeyError):
        gpm.encode_object(dict(pred[0], image_id="seg/1/REAR"))          # camera_names[camera_id]

"""Drop-in for ``generate_prediction_for_metrics.py`` of the reference (:8-81): submission / annotation JSON
-> serialized ``metrics.Objects`` for the Waymo metrics tools.  Same CLI flags, same file.

Like :mod:`.coco_to_waymo` it writes the protobuf wire format by hand because the ``waymo_open_dataset``
package is not available (PARITY UNPINNED at the schema: field numbers restated from the package's
published protos; ``Label.detection_difficulty_level = 5``, ``tracking_difficulty_level = 6``,
``num_lidar_points_in_box = 7`` in addition to those listed in ``csrc/waymo_pb.cpp``).  Pure Python: this
tool runs once per evaluation on a few hundred thousand rows and ground-truth ids are arbitrary strings.
"""
import argparse
import json
import struct

CAMERA_NAMES = {'FRONT': 1, 'FRONT_LEFT': 2, 'FRONT_RIGHT': 3, 'SIDE_LEFT': 4, 'SIDE_RIGHT': 5}   # :26-32
OBJECT_TYPES = {1: 1, 2: 2, 3: 3, 4: 4}       # TYPE_VEHICLE, TYPE_PEDESTRIAN, TYPE_SIGN, TYPE_CYCLIST (:33-38)
DIFFICULTY_LEVELS = {1: 1, 2: 2}              # LEVEL_1, LEVEL_2 (:39-42)


def _varint(v):
    v &= (1 << 64) - 1                        # negative int32 / int64 values travel as 64-bit two's complement
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7f) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _double(field, v):
    return bytes([(field << 3) | 1]) + struct.pack('<d', v)


def _bytes(field, b):
    return _varint((field << 3) | 2) + _varint(len(b)) + b


def encode_object(e):
    """One entry -> ``metrics.Object`` bytes, field for field what :46-76 of the reference assigns."""
    segment_id, frame_id, camera_id = e['image_id'].split('/')
    bbox = e['bbox']
    box = (_double(1, bbox[0] + bbox[2] / 2) + _double(2, bbox[1] + bbox[3] / 2) + _double(3, 0) +
           _double(4, bbox[3]) + _double(5, bbox[2]) + _double(6, 0) + _double(7, 0))
    label = _bytes(1, box) + b'\x18' + _varint(OBJECT_TYPES[e['category_id']])
    if 'object_id' in e:
        label += _bytes(4, e['object_id'].encode('utf-8'))
    if 'detection_difficulty_level' in e:
        label += b'\x28' + _varint(DIFFICULTY_LEVELS[e['detection_difficulty_level']])
    if 'tracking_difficulty_level' in e:
        label += b'\x30' + _varint(DIFFICULTY_LEVELS[e['tracking_difficulty_level']])
    label += b'\x38' + _varint(100)            # num_lidar_points_in_box: "work around for metrics computation" (:73)
    out = _bytes(1, label)
    if 'score' in e:
        out += b'\x15' + struct.pack('<f', e['score'])
    out += _bytes(4, segment_id.encode('utf-8')) + b'\x28' + _varint(int(frame_id)) + b'\x30' + _varint(CAMERA_NAMES[camera_id])
    return out


def encode_objects(entries):
    return b''.join(_bytes(1, encode_object(e)) for e in entries)


def main(argv=None):
    parser = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument("--type", choices=['prediction', 'ground-truth'],
                        default='prediction', help='to generate ground truth or prediction')
    parser.add_argument("--input", type=str, required=True,
                        help='either submission.json or annotations.json')
    parser.add_argument("--output", type=str,
                        default='out.bin', help='output file')
    args = parser.parse_args(argv)
    print(args)

    entries = json.load(open(args.input))
    if args.type == 'ground-truth':
        entries = entries['annotations']
    with open(args.output, 'wb') as f:
        f.write(encode_objects(entries))


if __name__ == '__main__':
    main()

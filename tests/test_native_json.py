"""CPU tests of the native JSON reader / writer (csrc/json_io.cpp) and of the array packers that sit
behind both CLIs: same arrays as Python's json + the reference-shaped dict loops, and files that are
byte-identical to json.dump of the reference's row dicts."""
import json

import numpy as np
import pytest

import helpers
from waymo_2d_tracking_b200 import native_json, packing, synth
from waymo_2d_tracking_b200._lib import W2TError
from waymo_2d_tracking_b200.detnet import ensemble as ens
from waymo_2d_tracking_b200.tracking import utils as trk_utils


def scene_and_lists(seed=9):
    cfg = synth.SynthConfig(n_segments=3, cameras=("SIDE_LEFT", "FRONT", "FRONT_RIGHT"), n_frames=12,
                            n_submissions=3, objects_per_frame=25.0, seed=seed)
    scene = synth.make_scene(cfg)
    return scene, [synth.to_json_list(scene, s) for s in scene.submissions]


def test_reader_matches_python_json(tmp_path):
    _, subs = scene_and_lists()
    dets = subs[0]
    dets[3]['extra'] = {"a": [1, 2, {"b": None}], "c": "x\"y\\z", "d": [True, False, 1.5e-3]}
    dets[5]['score'] = 1                              # an int where a float is usual
    dets[7]['image_id'] = 'se\u00e9g/12/FRONT'         # non-ASCII id (json.dump escapes it)
    del dets[9]['score']                              # ground-truth style row
    dets[11]['bbox'] = [1.5, -2.25, 3e2, 4.0]
    path = tmp_path / "s.json"
    path.write_text(json.dumps(dets, indent=1))       # whitespace everywhere
    got = native_json.load(path)
    ref = json.loads(path.read_text())
    assert [got.image_ids[i] for i in got.image_index] == [r['image_id'] for r in ref]
    assert got.image_ids == list(dict.fromkeys(r['image_id'] for r in ref))       # first-appearance order
    assert got.category.tolist() == [r['category_id'] for r in ref]
    assert got.bbox.tolist() == [[float(v) for v in r['bbox']] for r in ref]
    assert got.score.tolist() == [float(r.get('score', 1.0)) for r in ref]
    assert got.has_score.tolist() == [int('score' in r) for r in ref]
    path.write_text(json.dumps({'images': [{'id': 1}], 'annotations': dets[:10]}))
    assert len(native_json.load(path)) == 10
    path.write_text("[]")
    empty = native_json.load(path)
    assert len(empty) == 0 and empty.image_ids == []


@pytest.mark.parametrize("text", ['[{"image_id": "a/1/FRONT", "category_id": 1}]', '[{"image_id": "a", "bbox": [1,2,3]',
                                  '{"images": []}', '[1, 2]', '[] trailing'])
def test_reader_refuses_malformed_input(tmp_path, text):
    path = tmp_path / "bad.json"
    path.write_text(text)
    with pytest.raises(W2TError):
        native_json.load(path)
    with pytest.raises(W2TError):
        native_json.load(tmp_path / "missing.json")


def test_writers_are_byte_identical_to_json_dump(tmp_path):
    rng = np.random.default_rng(0)
    ids = ['seg_%d/%d/FRONT' % (i // 7, 1000 + i) for i in range(40)] + ['s\u00e9g "q"\\/1/SIDE_LEFT']
    n = 3000
    box = np.c_[rng.uniform(0, 1920, (n, 2)), rng.uniform(1, 300, (n, 2))]
    box[0] = [0.0, 1e-7, 123456789.0, 1e16]                          # repr switches to exponents at 1e16 ...
    box[1] = [5e-324, 1.7976931348623157e308, 0.1 + 0.2, 100.0]
    box[2] = [1e-5, 0.0001, 1e15, 123456.789e3]                       # ... and below 1e-4
    box[3] = [-0.0, -1.5, 2.0 ** 53, 1 / 3]
    score = np.clip(rng.uniform(0, 1.2, n), 0.2, 1.0)
    img, cat, oid = rng.integers(0, len(ids), n), rng.integers(1, 5, n), rng.integers(1, 10 ** 9, n)
    rows = [{'image_id': ids[i], 'bbox': [np.float64(v) for v in b], 'score': np.float64(s), 'category_id': int(c),
             'object_id': '%i' % o} for i, b, s, c, o in zip(img, box, score, cat, oid)]
    a, b = tmp_path / "py.json", tmp_path / "native.json"
    a.write_text(json.dumps(rows))
    native_json.write_tracks(b, ids, img, box, score, cat, oid)
    assert a.read_bytes() == b.read_bytes()
    ibox = rng.integers(-5, 2000, (n, 4))
    s5 = np.round(rng.uniform(0, 1, n), 5)
    rows = [{'image_id': ids[i], 'category_id': int(c), 'bbox': [int(v) for v in bb], 'score': np.float64(s)}
            for i, c, bb, s in zip(img, cat, ibox, s5)]
    a.write_text(json.dumps(rows))
    native_json.write_detections(b, ids, img, cat, ibox, s5)
    assert a.read_bytes() == b.read_bytes()
    native_json.write_tracks(b, [], [], np.zeros((0, 4)), [], [], [])
    assert b.read_text() == "[]"


def test_array_packers_equal_the_dict_packers(tmp_path):
    scene, subs = scene_and_lists()
    rng = np.random.default_rng(1)
    d = subs[0]
    rng.shuffle(d)                                   # streams, frames and categories first appear in any order
    d = [r for r in d if not (r['image_id'].endswith('FRONT') and r['image_id'].split('/')[1].endswith('500000'))]
    d[5]['bbox'][2] = 0                              # invalid box: the frame still exists
    d[6]['score'] = 0.0
    path = tmp_path / "a.json"
    path.write_text(json.dumps(d))
    pred = trk_utils.read_data_file(str(path), helpers.SCORE_THR)
    want = packing.pack_predictions(pred, 4)
    got = packing.pack_detections(native_json.load(path), helpers.SCORE_THR, 4)
    assert got.streams == want.streams
    for k in ("frame_ids", "stream_img_offsets", "det_start", "det_count", "det_box", "cam_wh", "class_rank"):
        np.testing.assert_array_equal(getattr(got, k), getattr(want, k), err_msg=k)
    seg = want.streams[2][0]
    one = packing.pack_detections(native_json.load(path), helpers.SCORE_THR, 4, segment_id=seg)
    w1 = packing.pack_predictions({seg: pred[seg]}, 4)
    assert one.streams == w1.streams
    np.testing.assert_array_equal(one.det_box, w1.det_box)
    np.testing.assert_array_equal(one.class_rank, w1.class_rank)
    # errors of the dict path, raised by the array path too
    d.append({'image_id': d[0]['image_id'], 'category_id': 9, 'bbox': [0, 0, 5, 5], 'score': 1.0})
    path.write_text(json.dumps(d))
    with pytest.raises(IndexError):
        packing.pack_detections(native_json.load(path), helpers.SCORE_THR, 4)
    path.write_text(json.dumps([{'image_id': 'no-slashes', 'category_id': 1, 'bbox': [0, 0, 5, 5], 'score': 1.0}]))
    with pytest.raises(ValueError):
        packing.pack_detections(native_json.load(path), helpers.SCORE_THR, 4)
    path.write_text(json.dumps([{'image_id': 'a/1/TOP', 'category_id': 1, 'bbox': [0, 0, 5, 5], 'score': 1.0}]))
    with pytest.raises(KeyError):
        packing.pack_detections(native_json.load(path), helpers.SCORE_THR, 4)
    # ensemble side
    files = []
    for k, s in enumerate(subs):
        pk = tmp_path / ("s%d.json" % k)
        pk.write_text(json.dumps(s))
        files.append(native_json.load(pk))
    a = packing.pack_detection_files(files, [1.0, 0.5, 0.25], 0.05)
    b = ens.pack_submission_lists(subs, [1.0, 0.5, 0.25], 0.05)
    assert a.image_ids == b.image_ids and a.category_ids == b.category_ids and a.max_group == b.max_group
    for k in ("group_offsets", "rows", "sub_counts"):
        np.testing.assert_array_equal(getattr(a, k), getattr(b, k), err_msg=k)


def test_threaded_reader_and_writers_equal_the_sequential_ones(tmp_path, monkeypatch):
    """Files large enough for the multi-threaded paths (csrc/json_io.cpp: the list is split at record boundaries,
    rows are formatted in ranges): same arrays, byte-identical files; a file whose strings contain the split
    pattern falls back to the sequential parse."""
    import dataclasses
    rng = np.random.default_rng(3)
    n = 70000
    rows = [{"image_id": "seg%d/%d/FRONT" % (i % 7, i % 997), "category_id": int(1 + i % 4),
             "bbox": [int(v) for v in rng.integers(0, 1900, 4)], "score": round(float(rng.uniform(0, 1)), 5)}
            for i in range(n)]
    path = tmp_path / "big.json"
    path.write_text(json.dumps(rows))
    assert path.stat().st_size > 6 << 20
    monkeypatch.setenv("W2T_JSON_THREADS", "1")
    seq = native_json.load(path)
    monkeypatch.setenv("W2T_JSON_THREADS", "3")
    par = native_json.load(path)
    for f in dataclasses.fields(seq):
        a, b = getattr(seq, f.name), getattr(par, f.name)
        assert np.array_equal(a, b) if isinstance(a, np.ndarray) else a == b, f.name
    assert len(par) == n and par.image_ids[:3] == ["seg0/0/FRONT", "seg1/1/FRONT", "seg2/2/FRONT"]
    box = rng.uniform(0, 1900, (n, 4))
    score = rng.uniform(0.2, 1, n)
    oid = rng.integers(1, 10 ** 6, n)
    for thr in ("1", "3"):
        monkeypatch.setenv("W2T_JSON_THREADS", thr)
        native_json.write_tracks(tmp_path / ("t%s.json" % thr), seq.image_ids, seq.image_index, box, score, seq.category, oid)
        native_json.write_detections(tmp_path / ("d%s.json" % thr), seq.image_ids, seq.image_index, seq.category,
                                     seq.bbox.astype(np.int32), seq.score)
    assert (tmp_path / "t1.json").read_bytes() == (tmp_path / "t3.json").read_bytes()
    assert (tmp_path / "d1.json").read_bytes() == (tmp_path / "d3.json").read_bytes()
    assert json.loads((tmp_path / "d3.json").read_text())[n - 1]["score"] == rows[n - 1]["score"]
    tricky = [{"image_id": 's}, {"x/%d/FRONT' % i, "category_id": 1, "bbox": [1, 2, 3, 4], "score": 0.5} for i in range(60000)]
    (tmp_path / "tricky.json").write_text(json.dumps(tricky))
    t = native_json.load(tmp_path / "tricky.json")
    assert len(t) == 60000 and t.image_ids[7] == 's}, {"x/7/FRONT'


def test_reader_is_as_strict_as_json_load(tmp_path):
    for bad in ('[{"image_id": "a", "category_id": 1, "bbox": [1-2, 2, 3, 4], "score": 0.5}]',
                '[{"image_id": "a", "category_id": 1, "bbox": [1e, 2, 3, 4], "score": 0.5}]',
                '[{"image_id": "a", "category_id": 1.5, "bbox": [1, 2, 3, 4], "score": 0.5}]'):
        p = tmp_path / "bad.json"
        p.write_text(bad)
        with pytest.raises(Exception):
            native_json.load(p)


def _write_files(scene, tmp_path, tweak=None):
    files = []
    for k, sub in enumerate(scene.submissions):
        rows = synth.to_json_list(scene, sub)
        if tweak:
            tweak(k, rows)
        p = tmp_path / ("sub%d.json" % k)
        p.write_text(json.dumps(rows))
        files.append(str(p))
    return files


@pytest.mark.parametrize("weights", [[1.0, 1.0, 1.0], [1.0, 0.5, 0.25]])
def test_native_grouping_equals_the_array_packer(tmp_path, weights):
    # w2t_json_group_files (parse + filter + group in one native call) against load + pack_detection_files
    from waymo_2d_tracking_b200 import _abi, pipeline
    cfg = synth.SynthConfig(n_segments=3, cameras=("FRONT", "FRONT_LEFT", "SIDE_RIGHT"), n_frames=12, n_submissions=3,
                            objects_per_frame=20.0, seed=31)
    scene = synth.make_scene(cfg)

    def tweak(k, rows):
        rows[3]['bbox'][2] = 0                      # dropped by the width filter
        rows[5]['score'] = 0.001                    # dropped by min_score
        if k == 1:                                  # an image that only this file holds
            rows.append({'image_id': 'zz_seg/17/FRONT', 'category_id': 4, 'bbox': [1, 2, 30, 40], 'score': 0.5})
    files = _write_files(scene, tmp_path, tweak)
    want = packing.pack_detection_files([native_json.load(f) for f in files], weights, 0.01)
    got = native_json.group_files(files, weights, 0.01)
    assert got.image_ids == want.image_ids and got.category_ids == want.category_ids
    np.testing.assert_array_equal(got.group_offsets, want.group_offsets)
    np.testing.assert_array_equal(got.rows, want.rows)
    np.testing.assert_array_equal(got.sub_counts, want.sub_counts)
    assert got.max_group == want.max_group and got.columns == len(want.category_ids)
    np.testing.assert_array_equal(got.image_order, np.arange(len(want.image_ids)))
    packed = packing.packed_rows(want.rows)
    if packed is None:
        assert got.packed is None
    else:
        np.testing.assert_array_equal(got.packed, packed)
    # the tracker's layout: the native call against the general path of pipeline.load_groups
    fast = pipeline.load_groups(files, weights, 0.01, 4)
    orig = native_json.group_files
    try:
        native_json.group_files = lambda *a, **k: None
        slow = pipeline.load_groups(files, weights, 0.01, 4)
    finally:
        native_json.group_files = orig
    assert fast[0] == slow[0] and fast[7] == slow[7]
    for a, b in zip(fast[1:6], slow[1:6]):
        np.testing.assert_array_equal(a, b)
    assert (fast[6] is None) == (slow[6] is None)
    if fast[6] is not None:
        np.testing.assert_array_equal(fast[6], slow[6])
    assert (weights[1] == 1.0) == (fast[6] is not None)       # weighted scores leave the 5-decimal grid


def test_native_grouping_leaves_unusual_inputs_to_the_general_path(tmp_path):
    from waymo_2d_tracking_b200 import _abi
    rows = [{'image_id': 'seg/1_0/FRONT', 'category_id': 1, 'bbox': [1, 2, 3, 4], 'score': 0.5}]
    p = tmp_path / "odd.json"
    p.write_text(json.dumps(rows))
    assert native_json.group_files([p], [1.0], 0.0, _abi.W2T_LAYOUT_STREAMS, 4) is None     # int('1_0') == 10
    assert native_json.group_files([p], [1.0], 0.0).image_ids == ['seg/1_0/FRONT']
    rows[0]['image_id'], rows[0]['category_id'] = 'seg/3/FRONT', 7
    p.write_text(json.dumps(rows))
    assert native_json.group_files([p], [1.0], 0.0, _abi.W2T_LAYOUT_STREAMS, 4) is None     # IndexError territory
    with pytest.raises(W2TError):
        native_json.group_files([tmp_path / "missing.json"], [1.0], 0.0)
    p.write_text("[]")
    empty = native_json.group_files([p], [1.0], 0.0, _abi.W2T_LAYOUT_STREAMS, 4)
    assert empty.image_ids == [] and len(empty.rows) == 0 and list(empty.group_offsets) == [0]


def _same_packed(a, b):
    assert a.streams == [tuple(s) for s in b.streams] and a.n_rows == b.n_rows and a.n_classes == b.n_classes
    for k in ("frame_ids", "stream_img_offsets", "det_start", "det_count", "det_box", "cam_wh", "class_rank"):
        x, y = np.asarray(getattr(a, k)), np.asarray(getattr(b, k))
        assert x.dtype == y.dtype or k in ("frame_ids",), k
        np.testing.assert_array_equal(x, y, err_msg=k)


def test_native_track_packer_equals_the_array_packer(tmp_path):
    # w2t_json_pack_tracks (parse + read_data_file's filters + stream layout in one native call) against
    # native_json.load + packing.pack_detections: every array, every option
    cfg = synth.SynthConfig(n_segments=4, cameras=("SIDE_LEFT", "FRONT", "FRONT_RIGHT"), n_frames=9, n_submissions=1,
                            objects_per_frame=30.0, seed=13)
    scene = synth.make_scene(cfg)
    rows = synth.to_json_list(scene, scene.submissions[0])
    rng = np.random.default_rng(3)
    rng.shuffle(rows)                                   # streams / frames / categories first appear in file order
    rows[7]['bbox'][2] = 0.5                            # dropped: width < 1
    rows[9]['score'] = 0.0                              # dropped: below every threshold
    del rows[11]['score']                               # ground-truth style row: score 1.0
    p = tmp_path / "dets.json"
    p.write_text(json.dumps(rows))
    dets = native_json.load(p)
    thr = [0.95, 0.6, 1.0, 0.9]
    _same_packed(packing.pack_track_file(p, thr, 4), packing.pack_detections(dets, thr, 4))
    seg = scene.segments[2]
    _same_packed(packing.pack_track_file(p, thr, 4, segment_id=seg), packing.pack_detections(dets, thr, 4, segment_id=seg))
    assert packing.pack_track_file(p, thr, 4, segment_id="no-such-segment").n_streams == 0
    for world in (2, 3, 5):
        for rank in range(world):
            _same_packed(packing.pack_track_file(p, thr, 4, segment_block=(rank, world)),
                         packing.pack_detections(dets, thr, 4, segment_block=(rank, world)))
    # {"annotations": [...]} wrapper, more categories than the file uses
    p2 = tmp_path / "gt.json"
    p2.write_text(json.dumps({"annotations": rows, "images": []}))
    _same_packed(packing.pack_track_file(p2, thr + [0.5, 0.5], 6), packing.pack_detections(native_json.load(p2), thr + [0.5, 0.5], 6))
    # empty file
    p3 = tmp_path / "empty.json"
    p3.write_text("[]")
    empty = packing.pack_track_file(p3, thr, 4)
    assert empty.n_streams == 0 and empty.n_rows == 0 and list(empty.stream_img_offsets) == [0]


def test_native_track_packer_leaves_exceptions_to_the_general_path(tmp_path):
    thr = [0.95, 0.6, 1.0, 0.9]
    row = {'image_id': 'seg/5/FRONT', 'category_id': 5, 'bbox': [1, 2, 30, 40], 'score': 0.99}
    p = tmp_path / "d.json"
    p.write_text(json.dumps([row]))
    assert native_json.pack_tracks(p, thr, 4) is None
    with pytest.raises(IndexError):                      # thresholds[4] in the reference (utils.py:85)
        packing.pack_track_file(p, thr, 4)
    row.update(category_id=1, image_id='seg/5')
    p.write_text(json.dumps([row]))
    with pytest.raises(ValueError):                      # image_id.split('/') unpacking (utils.py:70)
        packing.pack_track_file(p, thr, 4)
    row.update(image_id='seg/5/NOT_A_CAMERA')
    p.write_text(json.dumps([row]))
    with pytest.raises(KeyError):                        # IMAGE_SIZES[camera_id] (utils.py:21)
        packing.pack_track_file(p, thr, 4)
    row.update(image_id='seg/1_0/FRONT')                 # int('1_0') == 10: Python's spelling, general path
    p.write_text(json.dumps([row]))
    assert native_json.pack_tracks(p, thr, 4) is None
    assert list(packing.pack_track_file(p, thr, 4).frame_ids) == [10]
    with pytest.raises(W2TError):
        packing.pack_track_file(tmp_path / "missing.json", thr, 4)


def _random_rows(rng, n, segs, cams, frames, cats, ints=True):
    rows = []
    for _ in range(n):
        box = [float(rng.integers(-50, 1900)), float(rng.integers(-50, 1200)),
               float(rng.choice([0, 0.5, 1, 2, 30, 250])), float(rng.choice([-3, 0, 0.99, 1, 40, 300]))]
        if not ints:
            box = [v + float(rng.choice([0, 0.25])) for v in box]
        score = float(rng.choice([0.0, 0.005, 0.01, 0.59999, 0.6, 0.94999, 0.95, 0.9, 1.0, float("nan")]))
        rows.append({'image_id': '%s/%d/%s' % (rng.choice(segs), int(rng.choice(frames)), rng.choice(cams)),
                     'category_id': int(rng.choice(cats)), 'bbox': box, 'score': score})
    return rows


@pytest.mark.parametrize("seed", range(6))
def test_native_packers_equal_the_array_packers_on_awkward_inputs(tmp_path, seed):
    # random small files with the values the filters branch on: zero / negative / fractional sizes, scores on both
    # sides of every threshold, NaN scores (json.dump writes NaN), images that occur in one file only, unicode ids
    rng = np.random.default_rng(500 + seed)
    segs = ["s%d_%s" % (k, "é" if k == 1 else "x") for k in range(3)]
    cams = ["FRONT", "SIDE_LEFT", "FRONT_RIGHT"]
    frames = [3, 10, 7, 100, 20]
    thr = [0.95, 0.6, 1.0, 0.9]
    files = []
    for k in range(3):
        rows = _random_rows(rng, int(rng.integers(0, 60)), segs, cams, frames, [1, 2, 3, 4], ints=(seed % 2 == 0))
        p = tmp_path / ("f%d.json" % k)
        p.write_text(json.dumps(rows))
        files.append(p)
    weights = [1.0, 0.5, 1.0] if seed % 3 == 0 else [1.0, 1.0, 1.0]
    for min_score in (0.0, 0.01):
        want = packing.pack_detection_files([native_json.load(f) for f in files], weights, min_score)
        got = packing.pack_files(files, weights, min_score)
        assert got.image_ids == want.image_ids and got.category_ids == want.category_ids and got.max_group == want.max_group
        np.testing.assert_array_equal(got.group_offsets, want.group_offsets)
        np.testing.assert_array_equal(got.rows, want.rows)            # NaN == NaN positionally
        np.testing.assert_array_equal(got.sub_counts, want.sub_counts)
        packed = packing.packed_rows(want.rows)
        assert (got.packed is None) == (packed is None)
        if packed is not None:
            np.testing.assert_array_equal(got.packed, packed)
    for f in files:
        _same_packed(packing.pack_track_file(f, thr, 4), packing.pack_detections(native_json.load(f), thr, 4))


@pytest.mark.parametrize("seed", range(4))
def test_fused_cli_grouping_fast_and_general_paths_agree_on_awkward_inputs(tmp_path, seed):
    from waymo_2d_tracking_b200 import pipeline
    rng = np.random.default_rng(900 + seed)
    segs = ["b_seg", "a_seg", "a_seg_longer"]              # sorted order differs from first appearance
    cams = ["SIDE_RIGHT", "FRONT", "FRONT_LEFT"]
    files = []
    for k in range(2):
        rows = _random_rows(rng, int(rng.integers(5, 80)), segs, cams, [5, 40, 12, 7], [1, 2, 4])
        p = tmp_path / ("f%d.json" % k)
        p.write_text(json.dumps(rows))
        files.append(p)
    fast = pipeline.load_groups(files, [1.0, 1.0], 0.01, 4)
    orig = native_json.group_files
    try:
        native_json.group_files = lambda *a, **k: None
        slow = pipeline.load_groups(files, [1.0, 1.0], 0.01, 4)
    finally:
        native_json.group_files = orig
    assert fast[0] == slow[0] and fast[7] == slow[7]
    for a, b in zip(fast[1:6], slow[1:6]):
        np.testing.assert_array_equal(a, b)
    assert (fast[6] is None) == (slow[6] is None)
    if fast[6] is not None:
        np.testing.assert_array_equal(fast[6], slow[6])

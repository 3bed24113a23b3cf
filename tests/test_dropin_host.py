"""CPU tests of the host side of the drop-in layer: JSON readers, packers, argument parsers and the
error behaviour the reference has — everything that runs before a kernel launch."""
import json
import os

import numpy as np
import pytest

import golden_io
import helpers
from oracle import ensemble_port, ref_shim, sort_port
from waymo_2d_tracking_b200 import packing, synth
from waymo_2d_tracking_b200.detnet import ensemble as ens
from waymo_2d_tracking_b200.detnet.trainer.utils import get_num_workers
from waymo_2d_tracking_b200.tracking import track as track_cli
from waymo_2d_tracking_b200.tracking import utils as trk_utils


def small_scene(seed=3, n_sub=3):
    return synth.make_scene(synth.SynthConfig(n_segments=1, cameras=("FRONT", "SIDE_LEFT"), n_frames=5,
                                              n_submissions=n_sub, objects_per_frame=20.0, seed=seed))


def test_read_data_file_matches_port_and_keeps_filtered_frames(tmp_path):
    scene = small_scene(n_sub=1)
    dets = synth.to_json_list(scene, scene.submissions[0])
    dets.append({'image_id': dets[0]['image_id'].rsplit('/', 2)[0] + '/99/FRONT', 'category_id': 1,
                 'bbox': [1, 1, 0, 5], 'score': 1.0})                       # invalid box: frame exists, no entry
    path = tmp_path / "sub.json"
    path.write_text(json.dumps(dets))
    got = trk_utils.read_data_file(str(path), helpers.SCORE_THR)
    want = sort_port.group_entries(dets, helpers.SCORE_THR)
    assert got == want
    seg = next(iter(got))
    assert got[seg]['FRONT'][99] == []
    # the 'annotations' wrapper and missing scores (ground truth) are accepted like the reference does
    path.write_text(json.dumps({'annotations': [{'image_id': 'a/1/FRONT', 'category_id': 3, 'bbox': [0, 0, 5, 5]}]}))
    assert trk_utils.read_data_file(str(path), helpers.SCORE_THR)['a']['FRONT'][1][0]['score'] == 1.0


@pytest.mark.skipif(not ref_shim.available(), reason="reference checkout not mounted")
def test_read_data_file_matches_the_reference_itself(tmp_path):
    ref_utils, _, _ = ref_shim.load_tracking()
    scene = small_scene(seed=4, n_sub=1)
    path = tmp_path / "sub.json"
    path.write_text(json.dumps(synth.to_json_list(scene, scene.submissions[0])))
    assert trk_utils.read_data_file(str(path), helpers.SCORE_THR) == ref_utils.read_data_file(str(path), helpers.SCORE_THR)
    assert trk_utils.IMAGE_SIZES == ref_utils.IMAGE_SIZES
    np.testing.assert_array_equal(trk_utils.clip_xy('SIDE_LEFT', -3.0, 900.5), ref_utils.clip_xy('SIDE_LEFT', -3.0, 900.5))


def test_convert_submission_and_vectorised_packer_agree():
    scene = small_scene()
    subs = [synth.to_json_list(scene, s) for s in scene.submissions]
    subs[1].append({'image_id': subs[1][0]['image_id'], 'category_id': 2, 'bbox': [5, 5, 0, 9], 'score': 0.9})
    weights = [1.0, 0.5, 0.25]
    conv = [ens.convert_submission(s, w, 0.05) for s, w in zip(subs, weights)]
    port = [ensemble_port.convert_submission(s, w, 0.05) for s, w in zip(subs, weights)]
    assert conv == port
    image_ids = sorted(set(k for c in conv for k in c))
    a = packing.pack_submissions(conv, image_ids, [1, 2, 3, 4])
    b = ens.pack_submission_lists(subs, weights, 0.05)
    assert a.image_ids == b.image_ids and a.category_ids == b.category_ids
    np.testing.assert_array_equal(a.group_offsets, b.group_offsets)
    np.testing.assert_array_equal(a.rows, b.rows)
    np.testing.assert_array_equal(a.sub_counts, b.sub_counts)
    assert a.max_group == b.max_group
    # and the synthetic vectorised generator lays groups out the same way
    c = synth.groups_from_scene(scene, None, 0.01)
    d = ens.pack_submission_lists([synth.to_json_list(scene, s) for s in scene.submissions], [1, 1, 1], 0.01)
    order = helpers.sorted_image_order(scene.image_ids())
    assert [scene.image_ids()[i] for i in order if c.group_offsets[4 * i + 4] > c.group_offsets[4 * i]] == d.image_ids


def test_yml_weights_and_lxly_helpers():
    nested = {'runA': {'x.json': 2, 'y.json': 1}, 'z.json': 3}
    assert ens.load_yml_input_and_weight(nested) == [('runA/x.json', 2), ('runA/y.json', 1), ('z.json', 3)]
    b = np.array([[0.5, 10., 20., 4., 6.]])
    assert ens.lxly2cxcy(b.copy()).tolist() == [[0.5, 12., 23., 4., 6.]]
    assert ens.cxcy2lxly(ens.lxly2cxcy(b.copy())).tolist() == b.tolist()


def test_ensemble_cli_argument_surface_and_errors(tmp_path):
    p = ens.build_parser()
    a = p.parse_args(['a.json', 'b.json', '-o', 'out.json'])
    assert (a.method, a.iou_thresh, a.soft_nms_cut, a.min_score, a.jobs) == ("weighted_fusion", 0.5, 1.0, 0, 1)
    argfile = tmp_path / "args.txt"
    argfile.write_text("a.json\nb.json\n-m\nsoft_nms\n--min-score=0.01\n--soft-nms-cut=0.9\n-j\n-1\n")
    a = p.parse_args(['@' + str(argfile), '-o', 'o.json'])
    assert (a.method, a.min_score, a.soft_nms_cut, a.jobs) == ("soft_nms", 0.01, 0.9, -1)
    with pytest.raises(SystemExit):
        p.parse_args(['a.json', '-m', 'median'])
    # fewer than two inputs: AssertionError (ensemble.py:120); existing output: RuntimeError (:131-132)
    one = tmp_path / "one.json"
    one.write_text("[]")
    with pytest.raises(AssertionError):
        ens.main([str(one), '-o', str(tmp_path / "o.json")])
    out = tmp_path / "exists.json"
    out.write_text("[]")
    with pytest.raises(RuntimeError):
        ens.main([str(one), str(one), '-o', str(out)])


def test_get_num_workers_like_reference():
    n = os.cpu_count()
    assert get_num_workers(1) == 1 and get_num_workers(0) == n and get_num_workers(-1) == n - 1
    with pytest.raises(RuntimeError):
        get_num_workers(n + 1)
    with pytest.raises(RuntimeError):
        get_num_workers(1, device='tpu')


def test_track_cli_argument_surface():
    a = track_cli.build_parser().parse_args([])
    assert a.max_age == 1 and a.min_hits == 0 and a.segment_id is None
    assert a.score_threshold == [0.95, 0.6, 1.0, 0.9] and a.iou_threshold == [0.01, 0.01, 1.0, 0.0]
    a = track_cli.build_parser().parse_args(['--max-age=2', '--score-threshold=0.5,0.5,0.5,0.5', '--segment-id', 's'])
    assert a.max_age == 2 and a.score_threshold == [0.5] * 4 and a.segment_id == 's'


def test_tracking_errors_are_raised_before_any_launch(tmp_path):
    pred = {'seg': {'TOP': {1: [{'bbox': [0, 0, 5, 5], 'score': 1.0, 'category_id': 1}]}}}
    with pytest.raises(KeyError):                         # unknown camera (utils.py:21)
        trk_utils.track_sort(pred, 'seg', 'TOP', helpers.IOU_THR, 2, 0)
    pred = {'seg': {'FRONT': {1: [{'bbox': [0, 0, 5, 5], 'score': 1.0, 'category_id': 7}]}}}
    with pytest.raises(IndexError):                       # category beyond the threshold list (tracker_sort.py:46)
        trk_utils.track_sort(pred, 'seg', 'FRONT', helpers.IOU_THR, 2, 0)
    bad = tmp_path / "bad.json"
    bad.write_text(json.dumps([{'image_id': 'no-slashes', 'category_id': 1, 'bbox': [0, 0, 5, 5], 'score': 1.0}]))
    with pytest.raises(ValueError):                       # malformed image_id (utils.py:70)
        trk_utils.read_data_file(str(bad), helpers.SCORE_THR)
    ok = tmp_path / "ok.json"
    ok.write_text("[]")
    with pytest.raises(FileNotFoundError):                # --ground-truth must exist (track.py:34)
        track_cli.main(['--input', str(ok), '--ground-truth', str(tmp_path / "missing.json"),
                        '--output', str(tmp_path / "o.json")])


def test_rows_to_dicts_matches_unpack_tracks():
    g = golden_io.load("track_minhits")
    scene = helpers.golden_scene(g)
    packed = synth.tracks_from_submission(scene, scene.submissions[0], helpers.SCORE_THR)
    from oracle import c_oracle
    res = c_oracle.sort_track(packed, helpers.IOU_THR, int(g["max_age"]), int(g["min_hits"]))
    ids, _ = packing.assign_ids(packed.stream_img_offsets, 4, packed.det_start, res["out_count"], res["created"],
                                res["first_img"], packed.class_rank, res["out_birth"])
    want = packing.unpack_tracks(packed, res["out_box"], res["out_score"], res["out_count"], res["first_img"], ids)
    index = {iid: i for i, iid in enumerate(packing.image_id_strings(packed))}
    dense = dict(rows_box=np.array([r['bbox'] for r in want]), rows_score=np.array([r['score'] for r in want]),
                 rows_id=np.array([int(r['object_id']) for r in want]),
                 rows_img=np.array([index[r['image_id']] for r in want]),
                 rows_cat=np.array([r['category_id'] for r in want]))
    assert packing.rows_to_dicts(packed, dense) == want


def test_host_rows_decode_the_compact_pairs_lazily():
    # runtime.HostRows: (object id - id_base, image * 8 + category - 1) pairs -> int64 ids, int32 image / category
    from waymo_2d_tracking_b200 import runtime
    pair = np.array([[0, 5 * 8 + 0], [41, 5 * 8 + 3], [2 ** 31 - 1, (2 ** 28 - 1) * 8 + 7]], np.int32)
    rows = runtime.HostRows({"rows_compact": pair, "id_base": 2 ** 40, "n_rows": 3})
    assert "rows_id" not in rows
    np.testing.assert_array_equal(rows["rows_id"], np.array([2 ** 40, 2 ** 40 + 41, 2 ** 40 + 2 ** 31 - 1], np.int64))
    np.testing.assert_array_equal(rows["rows_img"], [5, 5, 2 ** 28 - 1])
    np.testing.assert_array_equal(rows["rows_cat"], [1, 4, 8])
    assert rows["rows_id"].dtype == np.int64 and rows["rows_img"].dtype == np.int32 and rows["rows_cat"].dtype == np.int32
    assert "rows_id" in rows                       # decoded once, then kept
    with pytest.raises(KeyError):
        rows["no_such_key"]


def test_rank_core_binding_splits_the_affinity_mask(monkeypatch):
    import os
    from waymo_2d_tracking_b200 import _bind
    if not hasattr(os, "sched_setaffinity"):
        pytest.skip("no sched_setaffinity on this platform")
    before = os.sched_getaffinity(0)
    cores = sorted(before)
    try:
        monkeypatch.setenv("W2T_BIND_CORES", "0")
        assert _bind.bind_rank_cores(0, 2) is None and os.sched_getaffinity(0) == before
        monkeypatch.delenv("W2T_BIND_CORES")
        assert _bind.bind_rank_cores(0, 1) is None                       # a single rank keeps every core
        if len(cores) >= 2:
            share = len(cores) // 2
            mine = _bind.bind_rank_cores(1, 2)
            assert mine == cores[share:2 * share] and os.sched_getaffinity(0) == set(mine)
    finally:
        os.sched_setaffinity(0, before)


def test_launch_counts_follow_the_library_rules():
    # runtime.nms_launch_count / sort_launch_count restate run_groups (csrc/softnms.cu) and launch_sort (csrc/sort.cu):
    # the bench line's gpu_launches is built from them
    from waymo_2d_tracking_b200 import runtime
    cap = int(runtime.lib().w2t_softnms_max_group())
    assert runtime.nms_launch_count(4000, 280) == 1              # a small job: one launch
    assert runtime.nms_launch_count(600000, 280) == 3            # size classes <= 32, <= 96, the rest
    assert runtime.nms_launch_count(600000, 90) == 2             # no group beyond 96 boxes
    assert runtime.nms_launch_count(600000, 60) == 1
    assert runtime.nms_launch_count(10, cap + 1) == 2            # + the pass over global scratch
    plan = {"aux_offset": 4096, "n_wide": 0, "n_mid": 0}
    assert runtime.sort_launch_count(plan, 3000) == 5            # dmax, classify, warps, clusters' second pass, CTAs
    plan = {"aux_offset": 4096, "n_wide": 2, "n_mid": 5}
    assert runtime.sort_launch_count(plan, 3000) == 7
    plan = {"aux_offset": -1, "n_wide": 2, "n_mid": 0}
    assert runtime.sort_launch_count(plan, 8) == 2 and runtime.sort_launch_count(plan, 2) == 1

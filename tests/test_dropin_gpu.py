"""GPU tests of the drop-in layer: the reference's Python call surface and both CLIs, run through
the CUDA library and compared with the committed golden fixtures (outputs of the reference's own
code) and with the oracle on seeded inputs."""
import json
from functools import partial

import numpy as np
import pytest
import torch

import golden_io
import helpers
from oracle import c_oracle, sort_port
from waymo_2d_tracking_b200 import runtime, synth
from waymo_2d_tracking_b200.detnet import ensemble as ens
from waymo_2d_tracking_b200.detnet.nn import tta
from waymo_2d_tracking_b200.detnet.utils import box_utils
from waymo_2d_tracking_b200.tracking import track as track_cli
from waymo_2d_tracking_b200.tracking import utils as trk_utils
from waymo_2d_tracking_b200.tracking.sort import sort as sort_mod
from waymo_2d_tracking_b200.tracking.sort.tracker_sort import DeviceMultiClassTrackerSort, MultiClassTrackerSort

pytestmark = pytest.mark.gpu

ALL_ENS = [("ensemble_c2_small", "soft_nms"), ("ensemble_weighted", "soft_nms"), ("ensemble_ties", "soft_nms"),
           ("ensemble_nms", "nms"), ("ensemble_fusion", "weighted_fusion"),
           ("ensemble_fusion_default", "weighted_fusion")]


def write_submissions(scene, tmp_path):
    files = []
    for k, sub in enumerate(scene.submissions):
        p = tmp_path / ("sub%d.json" % k)
        p.write_text(json.dumps(synth.to_json_list(scene, sub)))
        files.append(str(p))
    return files


# ---- detnet.ensemble ----------------------------------------------------------------------------

@pytest.mark.parametrize("name,method", ALL_ENS)
def test_ensemble_cli_matches_reference_golden(name, method, tmp_path):
    g = golden_io.load(name)
    scene = helpers.golden_scene(g)
    files = write_submissions(scene, tmp_path)
    weights = list(g["weights"])
    out = tmp_path / "out" / "ens.json"
    if len(set(weights)) > 1:                      # weights come from a .yml file (ensemble.py:113-118)
        yml = tmp_path / "inputs.yml"
        yml.write_text("".join("%s: %g\n" % (f, w) for f, w in zip(files, weights)))
        argv = [str(yml)]
        # a single .yml input is the one case with one positional argument
    else:
        argv = files
    argv += ['-o', str(out), '-m', method, '--iou-thresh', repr(float(g["iou_thresh"])),
             '--soft-nms-cut', repr(float(g["cut"])), '--min-score', repr(float(g["min_score"])), '-j', '-1']
    ens.main(argv)
    rows = json.loads(out.read_text())
    got = golden_io.dets_to_arrays(rows, scene.image_ids())
    for k in ("img", "cat", "bbox", "score"):
        np.testing.assert_array_equal(got[k], g["out_" + k])
    assert all(isinstance(v, int) for r in rows[:20] for v in r['bbox'])


def test_ensemble_function_and_library_form_agree_with_cli_path():
    g = golden_io.load("ensemble_c2_small")
    scene = helpers.golden_scene(g)
    subs = [synth.to_json_list(scene, s) for s in scene.submissions]
    whole = ens.ensemble_submissions(subs, None, "soft_nms", 0.5, 0.9, 0.01)
    want = golden_io.dets_to_arrays(whole, scene.image_ids())
    for k in ("img", "cat", "bbox", "score"):
        np.testing.assert_array_equal(want[k], g["out_" + k])
    # the reference's per-image function with its module globals (ensemble.py:50-64)
    from argparse import Namespace
    ens.args = Namespace(min_score=0.01)
    ens.merge_func = partial(tta.nms_detections, iou_thresh=0.5, soft=True, soft_nms_cut=0.9)
    conv = [ens.convert_submission(s, 1.0, 0.01) for s in subs]
    image_ids = sorted(set(k for c in conv for k in c))
    per_image = []
    for image_id in image_ids:
        per_image += ens.ensemble(image_id, [c[image_id] for c in conv], [1, 2, 3, 4])
    assert per_image == whole
    # an arbitrary callable takes the generic per-category route and gives the same rows
    ens.merge_func = lambda boxes: tta.nms_detections(boxes, iou_thresh=0.5, soft=True, soft_nms_cut=0.9)
    again = []
    for image_id in image_ids[:3]:
        again += ens.ensemble(image_id, [c[image_id] for c in conv], [1, 2, 3, 4])
    assert again == [r for r in whole if r['image_id'] in image_ids[:3]]


def test_nms_api_matches_reference_golden():
    g = golden_io.load("nms_api")
    for i in range(int(g["n_cases"])):
        p = "c%d_" % i
        overlap, top_k, soft, conf, cut = g[p + "args"]
        keep, sc = box_utils.nms(torch.from_numpy(g[p + "boxes"]), torch.from_numpy(g[p + "scores"]), overlap=overlap,
                                 top_k=int(top_k), soft=bool(soft), conf_thresh=conf, soft_nms_cut=cut)
        assert [int(k) for k in keep] == g[p + "keep"].tolist(), i
        np.testing.assert_array_equal(sc.numpy(), g[p + "out_scores"])     # decayed scores bit-exact


def test_detector_head_nms_float32_and_bbox_vote_match_reference_golden():
    """The Detect call pattern (detnet/nn/modules/detection.py:59-77): float32 boxes / scores through nms (arithmetic
    in float32 like torch: kept indices and decayed scores bit-exact) and bbox_vote (agreement to rounding: the
    reference's torch.sum order is not reproduced); the last cases are float64."""
    g = golden_io.load("nms_f32")
    for i in range(int(g["n_cases"])):
        p = "c%d_" % i
        overlap, top_k, soft, conf, cut = g[p + "args"]
        tb, ts = torch.from_numpy(g[p + "boxes"]), torch.from_numpy(g[p + "scores"])
        keep, sc = box_utils.nms(tb, ts, overlap=overlap, top_k=int(top_k), soft=bool(soft), conf_thresh=conf, soft_nms_cut=cut)
        assert [int(k) for k in keep] == g[p + "keep"].tolist(), i
        assert sc.dtype == ts.dtype
        np.testing.assert_array_equal(sc.numpy(), g[p + "out_scores"])
        keep_t = torch.as_tensor([int(k) for k in keep], dtype=torch.long)
        voted = box_utils.bbox_vote(tb[keep_t], sc, tb, ts, 0.6)
        assert voted.dtype == tb.dtype and tuple(voted.shape) == g[p + "voted"].shape
        np.testing.assert_allclose(voted.numpy(), g[p + "voted"], rtol=2e-6 if tb.dtype == torch.float32 else 1e-12, atol=0)


def test_nms_detections_and_merge_detections_vs_oracle():
    rng = np.random.default_rng(8)
    for trial in range(12):
        dets = []
        for k in range(3):
            n = int(rng.integers(0, 40))
            c = rng.uniform(50, 600, (n, 2))
            dets.append(np.c_[np.round(rng.uniform(0.02, 1, n), 5), c, rng.uniform(10, 150, (n, 2))])
        got = tta.nms_detections([d.copy() for d in dets], iou_thresh=0.45, soft=True, soft_nms_cut=0.9)
        rows = np.vstack(dets)
        want = c_oracle.softnms_groups(np.array([0, len(rows)], np.int32), rows, 0.45, 0.9, -np.inf, box_format=1)
        np.testing.assert_array_equal(got, want["merged"][:int(want["kept_count"][0])])
        got = tta.nms_detections([d.copy() for d in dets], iou_thresh=0.45)
        want = c_oracle.softnms_groups(np.array([0, len(rows)], np.int32), rows, 0.45, 1.0, -np.inf, box_format=1,
                                       hard=True)
        np.testing.assert_array_equal(got, want["merged"][:int(want["kept_count"][0])])
        if len(rows):
            np.testing.assert_array_equal(tta.merge_detections([d.copy() for d in dets], 0.5),
                                          c_oracle.merge_detections(dets, 0.5))
    assert tta.nms_detections([np.zeros((0, 5))], soft=True).shape == (0, 5)


def test_box_conversions_and_dtype_guard():
    b = torch.tensor([[10., 20., 4., 6.], [0.5, 0.25, 1., 3.]], dtype=torch.float64)
    pf = box_utils.point_form(b)
    assert pf.tolist() == [[8., 17., 12., 23.], [0., -1.25, 1., 1.75]]
    assert box_utils.center_size(pf).tolist() == b.tolist()
    with pytest.raises(TypeError):      # mixed dtypes: torch itself would refuse `sorted_scores *= weights`
        box_utils.nms(pf.float(), torch.tensor([0.5, 0.4], dtype=torch.float64), soft=True)
    with pytest.raises(TypeError):
        box_utils.nms(pf.half(), torch.tensor([0.5, 0.4]).half(), soft=True)
    keep, sc = box_utils.nms(pf[:0], torch.zeros(0, dtype=torch.float64), soft=True)
    assert keep == [] and sc.numel() == 0


def test_fusion_large_groups_vs_oracle():
    cfg = synth.preset("c5", cameras=("FRONT",), n_frames=2, seed=6)
    scene = synth.make_scene(cfg)
    groups = synth.groups_from_scene(scene, None, 0.01)
    got = runtime.fusion_groups(groups.group_offsets, groups.rows, groups.sub_counts, 0.5, 0.01, 4, helpers.SCORE_THR,
                                max_group=groups.max_group)
    want = c_oracle.fusion_groups(groups.group_offsets, groups.rows, groups.sub_counts, 0.5, 0.01, 4, helpers.SCORE_THR)
    np.testing.assert_array_equal(got["kept_count"], want["kept_count"])
    np.testing.assert_array_equal(got["ens_count"], want["ens_count"])
    np.testing.assert_array_equal(got["trk_count"], want["trk_count"])
    from waymo_2d_tracking_b200 import packing
    off = groups.group_offsets.astype(np.int64)
    rows, _ = packing.valid_row_index(off[:-1], want["kept_count"])
    np.testing.assert_array_equal(got["merged"][rows], want["merged"][rows])
    rows, _ = packing.valid_row_index(off[:-1], want["ens_count"])
    np.testing.assert_array_equal(got["ens_box"][rows], want["ens_box"][rows])
    np.testing.assert_array_equal(got["ens_score"][rows], want["ens_score"][rows])


# ---- tracking --------------------------------------------------------------------------------------

@pytest.mark.parametrize("name", ["track_c1_small", "track_minhits", "track_cyclist_ties"])
def test_track_cli_matches_reference_golden(name, tmp_path, capsys):
    g = golden_io.load(name)
    scene = helpers.golden_scene(g)
    sub = tmp_path / "sub.json"
    sub.write_text(json.dumps(synth.to_json_list(scene, scene.submissions[0])))
    gt = tmp_path / "images.json"
    gt.write_text(json.dumps([{'id': 'x', 'file_name': 'x.jpg'}]))
    out = tmp_path / "tracks.json"
    sort_mod.KalmanBoxTracker.count = 0
    track_cli.main(['--ground-truth', str(gt), '--input', str(sub), '--output', str(out),
                    '--max-age=%d' % int(g["max_age"]), '--min-hits=%d' % int(g["min_hits"]),
                    '--score-threshold=0.95,0.6,1.0,0.9'])
    printed = capsys.readouterr().out
    assert "duration:" in printed and scene.segments[0] in printed
    rows = json.loads(out.read_text())
    got = golden_io.tracks_to_arrays(rows, scene.image_ids())
    want = helpers.golden_tracks(g)      # the product's default promotion regime
    helpers.assert_tracks_equal(got, want, score_rtol=1e-9, box_exact=True)
    assert isinstance(rows[0]['object_id'], str)


def test_track_sort_per_stream_continues_the_global_id_counter():
    g = golden_io.load("track_c1_small")
    scene = helpers.golden_scene(g)
    pred = sort_port.group_entries(synth.to_json_list(scene, scene.submissions[0]), helpers.SCORE_THR)
    sort_mod.KalmanBoxTracker.count = 0
    rows = []
    for seg in pred:
        for cam in pred[seg]:
            rows += trk_utils.track_sort(pred, seg, cam, helpers.IOU_THR, 2, 0)
    got = golden_io.tracks_to_arrays(rows, scene.image_ids())
    want = helpers.golden_tracks(g)      # the product's default promotion regime
    helpers.assert_tracks_equal(got, want, score_rtol=1e-9, box_exact=True)
    assert sort_mod.KalmanBoxTracker.count == want["oid"].max()
    # --segment-id filter of the CLI == tracking that segment alone from a fresh counter
    sort_mod.KalmanBoxTracker.count = 0
    assert trk_utils.track_all({}, helpers.IOU_THR, 2, 0) == []


def test_stateful_sort_api_matches_port_frame_by_frame():
    cfg = synth.SynthConfig(n_segments=1, cameras=("FRONT",), n_frames=14, n_submissions=1, objects_per_frame=30.0,
                            seed=17)
    scene = synth.make_scene(cfg)
    pred = sort_port.group_entries(synth.to_json_list(scene, scene.submissions[0]), helpers.SCORE_THR)
    frames = pred[scene.segments[0]]['FRONT']
    sort_mod.KalmanBoxTracker.count = 0
    sort_port.BoxTracker.count = 0
    ours = MultiClassTrackerSort(max_age=2, min_hits=0)
    ref = sort_port.MultiClassTracker(max_age=2, min_hits=0)
    for fid in sorted(frames):
        rows = [[e['bbox'][0], e['bbox'][1], e['bbox'][0] + e['bbox'][2], e['bbox'][1] + e['bbox'][3], e['score'],
                 e['category_id']] for e in frames[fid]]
        a, b = ours.track(rows, helpers.IOU_THR), ref.track(rows, helpers.IOU_THR)
        assert list(a.keys()) == list(b.keys())
        for k in a:
            assert a[k].shape == b[k].shape
            np.testing.assert_array_equal(a[k][:, :5], b[k][:, :5])           # boxes and ids
            np.testing.assert_allclose(a[k][:, 5], b[k][:, 5], rtol=1e-9, atol=0)


@pytest.mark.parametrize("max_age,min_hits,seed", [(2, 0, 17), (1, 2, 18), (3, 1, 19)])
def test_device_resident_stateful_api_matches_port_frame_by_frame(max_age, min_hits, seed):
    """w2t_sort_step: one launch per frame, filters / lists / counters resident on the device between
    calls; frame by frame the same dict (key order), rows, ids and confidences as the reference's
    MultiClassTrackerSort (oracle port), including frames without detections of a category, an empty
    frame and a category that appears late."""
    cfg = synth.SynthConfig(n_segments=1, cameras=("FRONT",), n_frames=16, n_submissions=1, objects_per_frame=30.0,
                            seed=seed)
    scene = synth.make_scene(cfg)
    pred = sort_port.group_entries(synth.to_json_list(scene, scene.submissions[0]), helpers.SCORE_THR)
    frames = pred[scene.segments[0]]['FRONT']
    sort_mod.KalmanBoxTracker.count = 5
    sort_port.BoxTracker.count = 5
    ours = DeviceMultiClassTrackerSort(max_age=max_age, min_hits=min_hits, track_cap=128, det_cap=64)
    ref = sort_port.MultiClassTracker(max_age=max_age, min_hits=min_hits)
    n_rows = 0
    for i, fid in enumerate(sorted(frames)):
        rows = [[e['bbox'][0], e['bbox'][1], e['bbox'][0] + e['bbox'][2], e['bbox'][1] + e['bbox'][3], e['score'],
                 e['category_id']] for e in frames[fid]]
        if i < 3:
            rows = [r for r in rows if r[5] != 1]     # vehicles appear late: their Sort object is created last
        if i == 7:
            rows = []                                  # an image without any detection still steps every tracker
        if i in (9, 10):
            rows = [r for r in rows if r[5] != 2]     # a category without detections is stepped with an empty array
        a, b = ours.track(rows, helpers.IOU_THR), ref.track(rows, helpers.IOU_THR)
        assert list(a.keys()) == list(b.keys()) == ours.trackers
        for k in a:
            assert a[k].shape == b[k].shape, (i, k)
            np.testing.assert_array_equal(a[k][:, 4], b[k][:, 4])                      # ids
            np.testing.assert_allclose(a[k][:, :4], b[k][:, :4], rtol=1e-9, atol=0)    # boxes
            np.testing.assert_allclose(a[k][:, 5], b[k][:, 5], rtol=1e-9, atol=0)      # confidence
            n_rows += len(a[k])
        assert sort_mod.KalmanBoxTracker.count == sort_port.BoxTracker.count
    assert n_rows > 100
    with pytest.raises(runtime.W2TError):
        DeviceMultiClassTrackerSort(track_cap=8, det_cap=4).track([[0, 0, 10, 10, 1.0, 1]] * 5, helpers.IOU_THR)


def test_sort_stepper_two_streams_in_lockstep():
    """runtime.SortStepper over two streams advanced together (one launch per call for both); a stream may
    sit a call out (`exists`), which steps nothing — like a frame absent from the input.  Ids follow the
    global counter in call order: stream by stream, category by category."""
    scenes = [synth.make_scene(synth.SynthConfig(n_segments=1, cameras=(cam,), n_frames=12, n_submissions=1,
                                                 objects_per_frame=25.0, seed=sd)) for cam, sd in (("FRONT", 31), ("SIDE_LEFT", 32))]
    frames = []
    for sc in scenes:
        pred = sort_port.group_entries(synth.to_json_list(sc, sc.submissions[0]), helpers.SCORE_THR)
        f = pred[sc.segments[0]][sc.cameras[0]]
        frames.append([f[k] for k in sorted(f)])
    sort_port.BoxTracker.count = 0
    refs = [sort_port.MultiClassTracker(max_age=2, min_hits=0) for _ in scenes]
    stepper = runtime.SortStepper(helpers.IOU_THR, max_age=2, min_hits=0, n_streams=2, track_cap=128, det_cap=64)
    total = 0
    for i in range(12):
        exists = [True, i not in (4, 5)]
        batch, want = [], []
        for s in range(2):
            rows = [[e['bbox'][0], e['bbox'][1], e['bbox'][0] + e['bbox'][2], e['bbox'][1] + e['bbox'][3], e['score'],
                     e['category_id']] for e in frames[s][i]]
            batch.append((np.asarray([r[:5] for r in rows], np.float64).reshape(-1, 5),
                          np.asarray([r[5] - 1 for r in rows], np.int64)))
            want.append(refs[s].track(rows, helpers.IOU_THR) if exists[s] else {})
        got = stepper.step(batch, exists)
        for s in range(2):
            assert [c + 1 for c in got[s]] == list(want[s].keys()), (i, s)
            for c, v in got[s].items():
                w = want[s][c + 1]
                assert v.shape == w.shape
                np.testing.assert_array_equal(v[:, 4], w[:, 4])
                np.testing.assert_allclose(v[:, :4], w[:, :4], rtol=1e-9, atol=0)
                np.testing.assert_allclose(v[:, 5], w[:, 5], rtol=1e-9, atol=0)
                total += len(v)
    assert stepper.count == sort_port.BoxTracker.count and total > 100


def test_sort_stepper_long_run_forgets_old_id_bases():
    """400 calls with short-lived tracks: the per-call id-base table is trimmed to the oldest live track (the kernel
    reports it) and the ids still equal the port's."""
    sc = synth.make_scene(synth.SynthConfig(n_segments=1, cameras=("FRONT",), n_frames=400, n_submissions=1,
                                            objects_per_frame=12.0, mean_life=8.0, seed=77))
    pred = sort_port.group_entries(synth.to_json_list(sc, sc.submissions[0]), helpers.SCORE_THR)
    f = pred[sc.segments[0]][sc.cameras[0]]
    frames = [f[k] for k in sorted(f)]
    sort_port.BoxTracker.count = 0
    ref = sort_port.MultiClassTracker(max_age=2, min_hits=0)
    stepper = runtime.SortStepper(helpers.IOU_THR, max_age=2, min_hits=0, n_streams=1, track_cap=128, det_cap=64)
    for entries in frames:
        rows = [[e['bbox'][0], e['bbox'][1], e['bbox'][0] + e['bbox'][2], e['bbox'][1] + e['bbox'][3], e['score'],
                 e['category_id']] for e in entries]
        want = ref.track(rows, helpers.IOU_THR)
        got = stepper.step([(np.asarray([r[:5] for r in rows], np.float64).reshape(-1, 5),
                             np.asarray([r[5] - 1 for r in rows], np.int64))])[0]
        assert [c + 1 for c in got] == list(want.keys())
        for c, v in got.items():
            np.testing.assert_array_equal(v[:, 4], want[c + 1][:, 4])
    assert stepper.count == sort_port.BoxTracker.count > 300
    assert stepper._base_lo >= 256 and len(stepper._bases) <= 256


def test_sort_building_blocks_python_surface():
    d = np.array([10, 20, 50, 80, 0.9], np.float32)
    t = np.array([12.5, 18.0, 55.0, 77.0])
    assert sort_mod.iou(d, t) == float(np.float32(sort_port.iou(d, t)))
    np.testing.assert_array_equal(sort_mod.iou_batch(d[None], t[None]), sort_port.iou_matrix(d[None], t[None]))
    z = sort_mod.convert_bbox_to_z(d)
    assert z.shape == (4, 1) and z.dtype == (np.float64 if runtime.promotion_code() == 0 else np.float32)
    np.testing.assert_array_equal(z, sort_port.bbox_to_z(d).reshape(4, 1))
    x = np.array([30., 50., 2400., 0.66, 0, 0, 0])
    np.testing.assert_array_equal(sort_mod.convert_x_to_bbox(x), sort_port.x_to_bbox(x).reshape(1, 4))
    assert sort_mod.convert_x_to_bbox(x, 0.5).shape == (1, 5)
    m, ud, ut = sort_mod.associate_detections_to_trackers(np.zeros((3, 5), np.float32), np.zeros((0, 5)))
    assert m.shape == (0, 2) and ud.tolist() == [0, 1, 2] and ut.shape == (0, 5)
    trk = sort_mod.KalmanBoxTracker(d)
    ref = sort_port.BoxTracker(d)
    np.testing.assert_array_equal(trk.predict(), ref.predict())
    trk.update(d + 1)
    ref.update(d + 1)
    np.testing.assert_allclose(trk.kf.x, ref.kf.x, rtol=1e-9, atol=1e-9)
    assert trk.get_error() == pytest.approx(ref.get_error(), rel=1e-9)
    assert trk.hits == 1 and trk.time_since_update == 0 and trk.age == 1


# ---- fused CLI -------------------------------------------------------------------------------------

@pytest.mark.parametrize("with_ens", [False, True])
def test_fused_cli_writes_the_same_files_as_the_two_reference_commands(with_ens, tmp_path):
    # python -m detnet.ensemble ... -o ens.json && python tracking/track.py --input ens.json ... (README.md:41,54)
    # against the opt-in one-process pipeline: tracks.json (and ens.json) byte for byte
    from waymo_2d_tracking_b200 import pipeline
    cfg = synth.SynthConfig(n_segments=3, cameras=("FRONT", "FRONT_LEFT", "SIDE_RIGHT"), n_frames=25, n_submissions=3,
                            objects_per_frame=35.0, seed=77)
    scene = synth.make_scene(cfg)
    files = write_submissions(scene, tmp_path)
    gt = tmp_path / "images.json"
    gt.write_text(json.dumps([{'id': 'x', 'file_name': 'x.jpg'}]))
    ens_a, trk_a = tmp_path / "ens_a.json", tmp_path / "trk_a.json"
    ens.main(files + ['-o', str(ens_a), '-m', 'soft_nms', '--min-score=0.01', '--soft-nms-cut=0.9', '-j', '-1'])
    sort_mod.KalmanBoxTracker.count = 0
    track_cli.main(['--ground-truth', str(gt), '--input', str(ens_a), '--output', str(trk_a), '--max-age=2',
                    '--min-hits=0'])
    ens_b, trk_b = tmp_path / "ens_b.json", tmp_path / "trk_b.json"
    sort_mod.KalmanBoxTracker.count = 0
    argv = files + ['-o', str(trk_b), '--min-score=0.01', '--soft-nms-cut=0.9', '--max-age=2', '--min-hits=0']
    n = pipeline.main(argv + (['--ensemble-output', str(ens_b)] if with_ens else []))
    assert n == len(json.loads(trk_a.read_text())) and n > 1000
    assert trk_b.read_bytes() == trk_a.read_bytes()
    if with_ens:
        assert ens_b.read_bytes() == ens_a.read_bytes()

"""CPU tests: the oracle against the golden fixtures (outputs of the reference's own code)
and the three oracle layers (reference via shim, NumPy port, C restatement) against each other."""
import numpy as np
import pytest

import golden_io
import helpers
from oracle import c_oracle, ensemble_port, kalman, munkres, ref_shim, sort_port
from waymo_2d_tracking_b200 import packing, synth

TRACK_CASES = ["track_c1_small", "track_minhits", "track_cyclist_ties", "track_dense"]
ENS_CASES = ["ensemble_c2_small", "ensemble_weighted", "ensemble_ties"]


@pytest.mark.parametrize("promotion", helpers.PROMOTIONS)
@pytest.mark.parametrize("name", TRACK_CASES)
def test_c_oracle_tracking_matches_reference_golden(name, promotion):
    g = golden_io.load(name)
    scene = helpers.golden_scene(g)
    packed = synth.tracks_from_submission(scene, scene.submissions[0], helpers.SCORE_THR)
    res = c_oracle.sort_track(packed, helpers.IOU_THR, int(g["max_age"]), int(g["min_hits"]), promotion=promotion)
    assert res["status"] == 0
    got = helpers.track_rows_as_arrays(packed, res, scene.image_ids())
    want = helpers.golden_tracks(g, promotion)
    # boxes are bit-exact: the C oracle reproduces the BLAS/LAPACK operation order
    helpers.assert_tracks_equal(got, want, score_rtol=1e-12, box_exact=True)


@pytest.mark.parametrize("promotion", helpers.PROMOTIONS)
@pytest.mark.parametrize("name", TRACK_CASES[:2])
def test_numpy_port_tracking_matches_reference_golden(name, promotion):
    g = golden_io.load(name)
    scene = helpers.golden_scene(g)
    dets = synth.to_json_list(scene, scene.submissions[0])
    pred = sort_port.group_entries(dets, helpers.SCORE_THR)
    rows = sort_port.track_all(pred, helpers.IOU_THR, int(g["max_age"]), int(g["min_hits"]), promotion=promotion)
    got = golden_io.tracks_to_arrays(rows, scene.image_ids())
    want = helpers.golden_tracks(g, promotion)
    helpers.assert_tracks_equal(got, want, score_rtol=0, box_exact=True)


@pytest.mark.parametrize("promotion", helpers.PROMOTIONS)
def test_c_oracle_full_size_c1_matches_reference_golden(promotion):
    """Config C1 at full size (5 cameras x 200 frames): C oracle vs the reference's own output."""
    g = golden_io.load("big_c1")
    scene = synth.make_scene(synth.preset(str(g["preset"]), n_segments=1, seed=int(g["seed"])))
    packed = synth.tracks_from_submission(scene, scene.submissions[0], helpers.SCORE_THR)
    res = c_oracle.sort_track(packed, helpers.IOU_THR, int(g["max_age"]), int(g["min_hits"]), promotion=promotion)
    assert res["status"] == 0
    got = helpers.track_rows_as_arrays(packed, res, scene.image_ids())
    helpers.assert_big_tracks_equal(got, g, promotion)


def test_c_oracle_full_size_c2_matches_reference_golden():
    """Config C2 at full size (3 submissions, one segment): C oracle vs the reference's own output."""
    g = golden_io.load("big_c2")
    scene = synth.make_scene(synth.preset(str(g["preset"]), n_segments=1, seed=int(g["seed"])))
    groups = synth.groups_from_scene(scene, None, float(g["min_score"]))
    res = c_oracle.softnms_groups(groups.group_offsets, groups.rows, float(g["iou_thresh"]), float(g["cut"]),
                                  float(g["min_score"]), 4, helpers.SCORE_THR)
    got = helpers.ensemble_rows_as_arrays(groups.group_offsets, res, scene.n_img,
                                          image_order=helpers.sorted_image_order(scene.image_ids()))
    helpers.assert_big_ensemble_equal(got, g)


def test_dict_packer_equals_vectorised_packer():
    g = golden_io.load("track_c1_small")
    scene = helpers.golden_scene(g)
    dets = synth.to_json_list(scene, scene.submissions[0])
    pred = sort_port.group_entries(dets, helpers.SCORE_THR)
    a = packing.pack_predictions(pred, 4)
    b = synth.tracks_from_submission(scene, scene.submissions[0], helpers.SCORE_THR)
    np.testing.assert_array_equal(a.det_count, b.det_count)
    np.testing.assert_array_equal(a.det_start, b.det_start)
    np.testing.assert_array_equal(a.det_box, b.det_box)
    np.testing.assert_array_equal(a.class_rank, b.class_rank)
    np.testing.assert_array_equal(a.frame_ids, b.frame_ids)


@pytest.mark.parametrize("name", ENS_CASES)
def test_c_oracle_ensemble_matches_reference_golden(name):
    g = golden_io.load(name)
    scene = helpers.golden_scene(g)
    groups = synth.groups_from_scene(scene, list(g["weights"]), float(g["min_score"]))
    res = c_oracle.softnms_groups(groups.group_offsets, groups.rows, float(g["iou_thresh"]), float(g["cut"]),
                                  float(g["min_score"]), 4, helpers.SCORE_THR)
    got = helpers.ensemble_rows_as_arrays(groups.group_offsets, res, scene.n_img,
                                          image_order=helpers.sorted_image_order(scene.image_ids()))
    for k in ("img", "cat", "bbox", "score"):
        np.testing.assert_array_equal(got[k], g["out_" + k])     # kept sets, int boxes, rounded scores: exact


@pytest.mark.parametrize("name", ENS_CASES)
def test_numpy_port_ensemble_matches_reference_golden(name):
    g = golden_io.load(name)
    scene = helpers.golden_scene(g)
    subs = [synth.to_json_list(scene, s) for s in scene.submissions]
    rows = ensemble_port.ensemble_all(subs, list(g["weights"]), float(g["min_score"]), float(g["iou_thresh"]),
                                      float(g["cut"]))
    got = golden_io.dets_to_arrays(rows, scene.image_ids())
    for k in ("img", "cat", "bbox", "score"):
        np.testing.assert_array_equal(got[k], g["out_" + k])


@pytest.mark.parametrize("promotion", helpers.PROMOTIONS)
def test_c_oracle_pipeline_matches_reference_golden(promotion):
    """ensemble -> (int / 5-decimal rounding) -> tracker, against the reference run end to end."""
    g = golden_io.load("pipeline_small")
    scene = helpers.golden_scene(g)
    groups = synth.groups_from_scene(scene, None, 0.01)
    nms = c_oracle.softnms_groups(groups.group_offsets, groups.rows, 0.5, 0.9, 0.01, 4, helpers.SCORE_THR)
    ens = helpers.ensemble_rows_as_arrays(groups.group_offsets, nms, scene.n_img,
                                          image_order=helpers.sorted_image_order(scene.image_ids()))
    for k in ("img", "cat", "bbox", "score"):
        np.testing.assert_array_equal(ens[k], g["ens_" + k])
    packed = packing.PackedTracks(
        n_streams=scene.n_streams, n_classes=4, streams=scene.streams(), frame_ids=scene.frame_ids,
        stream_img_offsets=scene.stream_img_offsets, det_start=groups.group_offsets[:-1].copy(),
        det_count=nms["trk_count"], det_box=nms["trk_box"], cam_wh=scene.cam_wh(), img_exists=nms["img_exists"],
        class_rank=None, n_rows=len(groups.rows))
    res = c_oracle.sort_track(packed, helpers.IOU_THR, 2, 0, promotion=promotion)
    # the reference tracked streams in sorted-image_id order of the ensemble JSON
    order = stream_order_of_sorted_images(scene)
    got = rows_in_stream_order(packed, res, scene, order)
    want = helpers.golden_tracks(g, promotion)
    helpers.assert_tracks_equal(got, want, score_rtol=1e-12, box_exact=True)


def stream_order_of_sorted_images(scene):
    """Streams in the order read_data_file meets them when images come sorted by image_id."""
    ids = scene.image_ids()
    seen, order = set(), []
    nc = len(scene.cameras)
    # segment dict order = first appearance; cameras inside a segment likewise
    seg_first, cam_first = {}, {}
    for i in sorted(range(len(ids)), key=lambda k: ids[k]):
        s = int(np.searchsorted(scene.stream_img_offsets, i, side="right") - 1)
        seg = s // nc
        seg_first.setdefault(seg, len(seg_first))
        cam_first.setdefault((seg, s), len(cam_first))
    for seg in sorted(seg_first, key=seg_first.get):
        for (sg, s) in sorted((k for k in cam_first if k[0] == seg), key=cam_first.get):
            order.append(s)
    return order


def rows_in_stream_order(packed, res, scene, order):
    """Re-run the id scan with streams permuted to ``order`` and emit rows in that order."""
    NC = packed.n_classes
    offs = packed.stream_img_offsets
    img_perm = np.concatenate([np.arange(offs[s], offs[s + 1]) for s in order])
    inv = np.empty_like(img_perm)
    inv[img_perm] = np.arange(len(img_perm))
    gperm = (img_perm[:, None] * NC + np.arange(NC)[None, :]).reshape(-1)
    new_offs = np.r_[0, np.cumsum([offs[s + 1] - offs[s] for s in order])].astype(np.int32)
    birth = res["out_birth"].copy()
    valid, _ = packing.valid_row_index(packed.det_start, res["out_count"])
    birth[valid, 0] = inv[birth[valid, 0] // NC] * NC + birth[valid, 0] % NC
    p2 = packing.PackedTracks(
        n_streams=len(order), n_classes=NC, streams=[packed.streams[s] for s in order],
        frame_ids=packed.frame_ids[img_perm], stream_img_offsets=new_offs, det_start=packed.det_start[gperm],
        det_count=packed.det_count[gperm], det_box=packed.det_box, cam_wh=packed.cam_wh[order],
        class_rank=None, n_rows=packed.n_rows)
    first = res["first_img"].reshape(-1, NC)[order].reshape(-1)
    ids, _ = packing.assign_ids(new_offs, NC, p2.det_start, res["out_count"][gperm], res["created"][gperm], first,
                                None, birth)
    rows = packing.unpack_tracks(p2, res["out_box"], res["out_score"], res["out_count"][gperm], first, ids)
    return golden_io.tracks_to_arrays(rows, scene.image_ids())


# ---- layer against layer ------------------------------------------------------------------

def test_munkres_c_vs_numpy_restatement_and_scipy_cost():
    from scipy.optimize import linear_sum_assignment
    rng = np.random.default_rng(5)
    for trial in range(240):
        D, T = rng.integers(1, 48, 2)
        kind = trial % 6
        if kind == 0:
            M = rng.random((D, T))
        elif kind == 1:
            M = (rng.random((D, T)) < 0.12) * rng.random((D, T))          # sparse IoU-like
        elif kind == 2:
            M = np.round(rng.random((D, T)) * 4) / 4                      # heavy ties
        elif kind == 3:
            M = np.zeros((D, T))                                          # all zero: pure tie-breaking
        elif kind == 4:
            M = np.repeat(rng.random((D, 1)), T, 1)                       # identical columns
        else:
            M = (rng.random((D, T)) < 0.3).astype(float)                  # 0/1
        M = M.astype(np.float32)
        a = munkres.linear_assignment(-M)
        b = c_oracle.linear_assignment(-M)
        np.testing.assert_array_equal(a, b)
        assert len(a) == min(D, T) and len(set(a[:, 0])) == len(a) and len(set(a[:, 1])) == len(a)
        r, c = linear_sum_assignment(-M.astype(np.float64))
        assert abs(float(M[a[:, 0], a[:, 1]].astype(np.float64).sum()) - float(M[r, c].astype(np.float64).sum())) < 1e-4


def test_kalman_c_is_bit_identical_to_numpy_blas_path():
    rng = np.random.default_rng(7)
    mismatches = 0
    for trial in range(60):
        det = (np.array([0, 0, rng.uniform(10, 300), rng.uniform(10, 300)]) + rng.uniform(0, 900)).astype(np.float32)
        t = sort_port.BoxTracker(np.r_[det, 0.9].astype(np.float32))
        x, P = c_oracle.kf_init(det)
        np.testing.assert_array_equal(x, t.kf.x[:, 0])
        for step in range(15):
            box = t.predict()[0]
            x, P = c_oracle.kf_predict(x, P)
            np.testing.assert_array_equal(x, t.kf.x[:, 0])            # predict: always exact
            np.testing.assert_array_equal(P, t.kf.P)
            np.testing.assert_array_equal(c_oracle.x_to_bbox(x), box)
            if rng.random() < 0.8:
                det = (det + rng.normal(0, 3, 4)).astype(np.float32)
                t.update(np.r_[det, 0.9].astype(np.float32))
                x, P = c_oracle.kf_update(x, P, det)
                # BLAS kernels differ between CPUs: demand 1e-9, count bit mismatches
                np.testing.assert_allclose(x, t.kf.x[:, 0], rtol=1e-9, atol=1e-9)
                assert np.abs(P - t.kf.P).max() <= 1e-9 * np.abs(t.kf.P).max()
                mismatches += not (np.array_equal(x, t.kf.x[:, 0]) and np.array_equal(P, t.kf.P))
                x, P = t.kf.x[:, 0].copy(), t.kf.P.copy()
    print("kalman update bit mismatches vs this host's BLAS:", mismatches)


def test_kalman_nan_and_negative_scale_paths():
    # s + ds <= 0 zeroes the scale velocity (sort.py:170-171); a negative s*r gives a NaN box
    x = np.array([10., 10., 4., 1., 0., 0., -9.])
    P = np.eye(7)
    x2, _ = c_oracle.kf_predict(x, P)
    assert x2[6] == 0.0 and x2[2] == 4.0
    x = np.array([10., 10., -4., 1., 0., 0., 1.])
    b = c_oracle.x_to_bbox(x)
    assert np.isnan(b).all()


def test_iou_matrix_c_vs_numpy_port():
    rng = np.random.default_rng(9)
    dets = (rng.uniform(0, 500, (40, 2)))
    dets = np.c_[dets, dets + rng.uniform(1, 200, (40, 2))].astype(np.float32)
    trks = rng.uniform(0, 500, (33, 2))
    trks = np.c_[trks, trks + rng.uniform(1, 200, (33, 2))]
    np.testing.assert_array_equal(c_oracle.iou_matrix(dets, trks), sort_port.iou_matrix(dets, trks))
    one = np.float32(sort_port.iou(dets[3], trks[4]))
    assert c_oracle.iou_matrix(dets, trks)[3, 4] == one


def test_soft_nms_general_arguments_c_vs_numpy_port():
    rng = np.random.default_rng(3)
    for trial in range(40):
        n = int(rng.integers(0, 60))
        xy = rng.uniform(0, 200, (n, 2))
        boxes = np.c_[xy, xy + rng.uniform(5, 80, (n, 2))]
        scores = np.round(rng.uniform(0.01, 1, n), 2 if trial % 2 else 5)
        top_k = int(rng.integers(0, 20)) if trial % 3 == 0 else 0
        conf = 0.0 if trial % 4 else 0.3
        k1, s1 = ensemble_port.soft_nms(boxes, scores, 0.5, top_k, conf, 0.9)
        k2, s2 = c_oracle.soft_nms(boxes, scores, 0.5, top_k, conf, 0.9)
        assert k1 == k2
        np.testing.assert_array_equal(s1, s2)


def test_score_rounding_is_numpy_round_not_python_round():
    # ensemble.py:62 rounds an np.float64: rint(x*1e5)/1e5, which differs from round(float, 5)
    rows = np.array([[0.123455, 0, 0, 10, 10], [0.5000049999, 50, 50, 10, 10]])
    res = c_oracle.softnms_groups(np.array([0, 2], np.int32), rows, 0.5, 0.9, 0.0)
    got = sorted(res["ens_score"][:2].tolist())
    want = sorted(float(round(np.float64(v), 5)) for v in rows[:, 0])
    assert got == want


@pytest.mark.skipif(not ref_shim.available(), reason="reference checkout not mounted")
def test_reference_itself_agrees_with_ports_on_a_fresh_seed():
    cfg = synth.SynthConfig(n_segments=1, cameras=("FRONT",), n_frames=12, n_submissions=2, objects_per_frame=25.0,
                            seed=77)
    scene = synth.make_scene(cfg)
    subs = [synth.to_json_list(scene, s) for s in scene.submissions]
    ens_ref = ref_shim.ref_ensemble_all(subs, None, 0.01, 0.5, 0.9)
    ens_port = ensemble_port.ensemble_all(subs, None, 0.01, 0.5, 0.9)
    assert ens_ref == ens_port
    pred = sort_port.group_entries(ens_ref, helpers.SCORE_THR)
    for promotion in helpers.PROMOTIONS:
        a = ref_shim.ref_track_all(pred, helpers.IOU_THR, 2, 0, promotion=promotion)
        b = sort_port.track_all(pred, helpers.IOU_THR, 2, 0, promotion=promotion)
        assert len(a) == len(b)
        for x, y in zip(a, b):
            assert x["image_id"] == y["image_id"] and x["object_id"] == y["object_id"]
            assert [float(v) for v in x["bbox"]] == [float(v) for v in y["bbox"]] and float(x["score"]) == float(y["score"])


# ---- the CLI's other two methods and the general nms() signature -------------------------------

METHOD_CASES = ["ensemble_nms", "ensemble_fusion", "ensemble_fusion_default"]


@pytest.mark.parametrize("name", METHOD_CASES)
def test_c_oracle_other_methods_match_reference_golden(name):
    g = golden_io.load(name)
    scene = helpers.golden_scene(g)
    groups = synth.groups_from_scene(scene, list(g["weights"]), float(g["min_score"]))
    if str(g["method"]) == "nms":
        res = c_oracle.softnms_groups(groups.group_offsets, groups.rows, float(g["iou_thresh"]), 1.0,
                                      float(g["min_score"]), hard=True)
    else:
        res = c_oracle.fusion_groups(groups.group_offsets, groups.rows, groups.sub_counts, float(g["iou_thresh"]),
                                     float(g["min_score"]))
    got = helpers.ensemble_rows_as_arrays(groups.group_offsets, res, scene.n_img,
                                          image_order=helpers.sorted_image_order(scene.image_ids()))
    for k in ("img", "cat", "bbox", "score"):
        np.testing.assert_array_equal(got[k], g["out_" + k])


def test_c_oracle_nms_api_matches_reference_golden():
    g = golden_io.load("nms_api")
    for i in range(int(g["n_cases"])):
        p = "c%d_" % i
        overlap, top_k, soft, conf, cut = g[p + "args"]
        if soft:
            keep, sc = c_oracle.soft_nms(g[p + "boxes"], g[p + "scores"], overlap, int(top_k), conf, cut)
        else:
            keep = c_oracle.hard_nms(g[p + "boxes"], g[p + "scores"], overlap, int(top_k))
            sc = g[p + "scores"][keep]
        assert keep == g[p + "keep"].tolist(), i
        np.testing.assert_array_equal(sc, g[p + "out_scores"])


def test_compact_rows_give_the_same_ensemble():
    g = golden_io.load("ensemble_c2_small")
    scene = helpers.golden_scene(g)
    groups = synth.groups_from_scene(scene, None, 0.01)
    compact = packing.compact_rows(groups.rows)
    assert compact is not None and compact.dtype.itemsize == 16
    a = c_oracle.softnms_groups(groups.group_offsets, groups.rows, 0.5, 0.9, 0.01, 4, helpers.SCORE_THR)
    b = c_oracle.softnms_groups(groups.group_offsets, compact, 0.5, 0.9, 0.01, 4, helpers.SCORE_THR)
    for k in ("merged", "src_index", "ens_count", "ens_box", "ens_score", "trk_count", "trk_box"):
        np.testing.assert_array_equal(a[k], b[k])
    assert packing.compact_rows(np.array([[0.5, 1.5, 2, 3, 4]])) is None          # non-integer box
    assert packing.compact_rows(np.array([[0.5, 1, 2, 40000, 4]])) is None        # beyond int16


def test_packed_rows_give_the_same_ensemble():
    g = golden_io.load("ensemble_c2_small")
    scene = helpers.golden_scene(g)
    groups = synth.groups_from_scene(scene, None, 0.01)
    packed = packing.packed_rows(groups.rows)
    assert packed is not None and packed.dtype == np.uint64 and len(packed) == len(groups.rows)
    # lossless by construction: score and box decode to the very same doubles
    np.testing.assert_array_equal((packed & np.uint64(0x1ffff)).astype(np.float64) / 1e5, groups.rows[:, 0])
    for j, (shift, mask, bias) in enumerate(((17, 0x1fff, 3072), (30, 0xfff, 1536), (42, 0x7ff, 0), (53, 0x7ff, 0))):
        np.testing.assert_array_equal(((packed >> np.uint64(shift)) & np.uint64(mask)).astype(np.float64) - bias,
                                      groups.rows[:, 1 + j])
    a = c_oracle.softnms_groups(groups.group_offsets, groups.rows, 0.5, 0.9, 0.01, 4, helpers.SCORE_THR)
    b = c_oracle.softnms_groups(groups.group_offsets, packed, 0.5, 0.9, 0.01, 4, helpers.SCORE_THR)
    for k in ("merged", "src_index", "ens_count", "ens_box", "ens_score", "trk_count", "trk_box"):
        np.testing.assert_array_equal(a[k], b[k])
    # anything the 8 bytes cannot hold exactly is refused
    assert packing.packed_rows(np.array([[0.5, 1.5, 2, 3, 4]])) is None           # non-integer box
    assert packing.packed_rows(np.array([[0.5, 1, 2, 2048, 4]])) is None          # beyond 11 bits
    assert packing.packed_rows(np.array([[0.5, -1, 2, 3, 4]])) is not None        # negative left is fine ...
    assert packing.packed_rows(np.array([[0.5, 1, 2, -3, 4]])) is None            # ... a negative width is not
    assert packing.packed_rows(np.array([[0.5, -3073, 2, 3, 4]])) is None
    assert packing.packed_rows(np.array([[0.5, 1, 2560, 3, 4]])) is None
    assert packing.packed_rows(np.array([[0.123456, 1, 2, 3, 4]])) is None        # not a 5-decimal score
    assert packing.packed_rows(np.array([[0.12345 * 0.7, 1, 2, 3, 4]])) is None   # weighted score
    assert packing.packed_rows(np.array([[1.5, 1, 2, 3, 4]])) is None             # k >= 2**17
    rng = np.random.default_rng(3)
    k = rng.integers(0, 100001, 20000)
    exact = np.array([float("%.5f" % (v / 1e5)) for v in k])       # how Python parses the 5-decimal literal
    rows = np.column_stack([exact, rng.integers(-3072, 5120, 20000), rng.integers(-1536, 2560, 20000), rng.integers(0, 2048, (20000, 2))]).astype(np.float64)
    p = packing.packed_rows(rows)
    assert p is not None
    np.testing.assert_array_equal((p & np.uint64(0x1ffff)).astype(np.int64), k)


def _random_tracking_case(seed, n_frames=14):
    rng = np.random.default_rng(seed)
    mix = rng.dirichlet([1.0, 1.0, 0.3, 0.6])
    cfg = synth.SynthConfig(n_segments=1, cameras=("FRONT", "SIDE_RIGHT"), n_frames=n_frames, n_submissions=1,
                            objects_per_frame=float(rng.uniform(5, 70)), class_mix=tuple(float(v) for v in mix),
                            mean_life=float(rng.uniform(3, 50)), p_miss=float(rng.uniform(0.0, 0.5)),
                            jitter=float(rng.uniform(0.5, 6.0)), seed=seed)
    scene = synth.make_scene(cfg)
    score_thr = [float(rng.choice([0.0, 0.3, 0.6, 0.9, 0.95])) for _ in range(4)]
    iou_thr = [float(rng.choice([0.0, 0.01, 0.3, 0.5, 1.0])) for _ in range(4)]
    return scene, score_thr, iou_thr, int(rng.integers(0, 5)), int(rng.integers(0, 4))


@pytest.mark.parametrize("seed", range(300, 310))
def test_c_oracle_equals_numpy_port_on_random_configurations(seed):
    scene, score_thr, iou_thr, max_age, min_hits = _random_tracking_case(seed)
    promotion = helpers.PROMOTIONS[seed % 2]
    packed = synth.tracks_from_submission(scene, scene.submissions[0], score_thr)
    res = c_oracle.sort_track(packed, iou_thr, max_age, min_hits, promotion=promotion)
    got = helpers.track_rows_as_arrays(packed, res, scene.image_ids())
    pred = sort_port.group_entries(synth.to_json_list(scene, scene.submissions[0]), score_thr)
    want = golden_io.tracks_to_arrays(sort_port.track_all(pred, iou_thr, max_age, min_hits, promotion=promotion),
                                      scene.image_ids())
    helpers.assert_tracks_equal(got, want, score_rtol=1e-12, box_exact=False)


@pytest.mark.skipif(not ref_shim.available(), reason="reference checkout not mounted")
@pytest.mark.parametrize("seed", [400, 401, 402])
def test_reference_itself_equals_c_oracle_on_random_configurations(seed):
    scene, score_thr, iou_thr, max_age, min_hits = _random_tracking_case(seed, n_frames=10)
    promotion = helpers.PROMOTIONS[seed % 2]
    packed = synth.tracks_from_submission(scene, scene.submissions[0], score_thr)
    res = c_oracle.sort_track(packed, iou_thr, max_age, min_hits, promotion=promotion)
    got = helpers.track_rows_as_arrays(packed, res, scene.image_ids())
    pred = sort_port.group_entries(synth.to_json_list(scene, scene.submissions[0]), score_thr)
    want = golden_io.tracks_to_arrays(ref_shim.ref_track_all(pred, iou_thr, max_age, min_hits, promotion=promotion),
                                      scene.image_ids())
    helpers.assert_tracks_equal(got, want, score_rtol=1e-12, box_exact=False)

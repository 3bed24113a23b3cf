"""GPU parity tests: the CUDA path (through the C-ABI of libw2t.so) against the oracle on the
same inputs and against the committed golden fixtures (outputs of the reference's own code).

Bars (BASELINE.json north_star): track ids, assignments and kept-box sets bit-exact; Kalman
states and decayed scores within 1e-9 relative (they are in fact bit-identical to the C oracle,
which these tests demand wherever no transcendental function is involved)."""
import numpy as np
import pytest

import golden_io
import helpers
from oracle import c_oracle
from waymo_2d_tracking_b200 import packing, runtime, synth

pytestmark = pytest.mark.gpu

TRACK_CASES = ["track_c1_small", "track_minhits", "track_cyclist_ties", "track_dense"]
ENS_CASES = ["ensemble_c2_small", "ensemble_weighted", "ensemble_ties"]


def rand_boxes(rng, n, lo=0, hi=800, smin=2, smax=250):
    xy = rng.uniform(lo, hi, (n, 2))
    return np.c_[xy, xy + rng.uniform(smin, smax, (n, 2))]


# ---- building blocks ------------------------------------------------------------------------

@pytest.mark.parametrize("promotion", helpers.PROMOTIONS)
def test_kalman_bit_exact_vs_c_oracle(promotion):
    rng = np.random.default_rng(0)
    n = 257
    dets = rand_boxes(rng, n).astype(np.float32)
    x, P = runtime.kf_init(dets, promotion=promotion)
    ox = np.zeros_like(x)
    oP = np.zeros_like(P)
    for i in range(n):
        ox[i], oP[i] = c_oracle.kf_init(dets[i], promotion=promotion)
    z = runtime.bbox_to_z(dets, promotion=promotion)
    assert z.dtype == (np.float64 if promotion == "legacy" else np.float32)
    np.testing.assert_array_equal(z, np.stack([c_oracle.bbox_to_z(d, promotion=promotion) for d in dets]))
    np.testing.assert_array_equal(x, ox)
    np.testing.assert_array_equal(P, oP)
    for step in range(12):
        x, P, boxes = runtime.kf_predict(x, P)
        for i in range(n):
            ox[i], oP[i] = c_oracle.kf_predict(ox[i], oP[i])
        np.testing.assert_array_equal(x, ox)
        np.testing.assert_array_equal(P, oP)
        np.testing.assert_array_equal(boxes, np.stack([c_oracle.x_to_bbox(v) for v in ox]))
        dets = (dets + rng.normal(0, 3, dets.shape)).astype(np.float32)
        dets[:, 2:] = np.maximum(dets[:, 2:], dets[:, :2] + 1)
        x, P, _ = runtime.kf_update(x, P, dets, promotion=promotion)
        for i in range(n):
            ox[i], oP[i] = c_oracle.kf_update(ox[i], oP[i], dets[i], promotion=promotion)
        np.testing.assert_array_equal(x, ox)                     # bit-exact, far inside the 1e-9 bar
        np.testing.assert_array_equal(P, oP)


def test_kalman_negative_scale_gives_nan_box_like_reference():
    x = np.array([[10., 10., -4., 1., 0., 0., 1.], [10., 10., 4., 1., 0., 0., -9.]])
    P = np.tile(np.eye(7), (2, 1, 1))
    x2, _, boxes = runtime.kf_predict(x, P)
    assert np.isnan(boxes[0]).all()
    assert x2[1, 6] == 0.0 and x2[1, 2] == 4.0


def test_iou_matrix_bit_exact():
    rng = np.random.default_rng(1)
    for D, T in [(1, 1), (7, 3), (64, 50), (130, 257)]:
        dets = rand_boxes(rng, D).astype(np.float32)
        trks = rand_boxes(rng, T)
        trks[: T // 3] = dets[rng.integers(0, D, T // 3)] + rng.normal(0, 2, (T // 3, 4))
        np.testing.assert_array_equal(runtime.iou_matrix(dets, trks), c_oracle.iou_matrix(dets, trks))


def test_linear_assignment_bit_exact_including_ties():
    rng = np.random.default_rng(2)
    # (100, 128), (128, 128), (90, 120): the shared-memory solver with its cost matrix left in global memory
    shapes = [(1, 1), (1, 9), (9, 1), (5, 5), (17, 40), (40, 17), (33, 33), (64, 65), (70, 200), (150, 90),
              (100, 128), (128, 128), (120, 90), (65, 96), (128, 129)]
    for trial, (D, T) in enumerate(shapes * 4):
        kind = trial % 5
        if kind == 0:
            M = rng.random((D, T))
        elif kind == 1:
            M = (rng.random((D, T)) < 0.1) * rng.random((D, T))
        elif kind == 2:
            M = np.round(rng.random((D, T)) * 3) / 3
        elif kind == 3:
            M = np.zeros((D, T))
        else:
            M = (rng.random((D, T)) < 0.3).astype(float)
        cost = (-M).astype(np.float32)
        np.testing.assert_array_equal(runtime.linear_assignment(cost), c_oracle.linear_assignment(cost))


def test_linear_assignment_large_crowded():
    # IoU-like matrix of a crowded scene (C4 sizes): many step-6 rounds
    rng = np.random.default_rng(3)
    dets = rand_boxes(rng, 450, 0, 1500, 20, 120).astype(np.float32)
    trks = np.r_[dets[:380] + rng.normal(0, 6, (380, 4)), rand_boxes(rng, 140, 0, 1500, 20, 120)]
    cost = -c_oracle.iou_matrix(dets, trks)
    np.testing.assert_array_equal(runtime.linear_assignment(cost), c_oracle.linear_assignment(cost))
    np.testing.assert_array_equal(runtime.linear_assignment(cost.T.copy()), c_oracle.linear_assignment(cost.T.copy()))


def test_linear_assignment_beyond_2048():
    # past the old 2048 limit of the solver's mask words (now 4096): 300 detections against 2500 trackers, and the
    # transposed problem (a full 2500 x 2500 solve takes the C oracle two minutes, so the square case stays out)
    rng = np.random.default_rng(5)
    trks = rand_boxes(rng, 2500, 0, 6000, 20, 90)
    dets = np.r_[trks[:250] + rng.normal(0, 4, (250, 4)), rand_boxes(rng, 50, 0, 6000, 20, 90)].astype(np.float32)
    cost = -c_oracle.iou_matrix(dets, trks)
    np.testing.assert_array_equal(runtime.linear_assignment(cost), c_oracle.linear_assignment(cost))
    np.testing.assert_array_equal(runtime.linear_assignment(cost.T.copy()), c_oracle.linear_assignment(cost.T.copy()))


# ---- SORT stage -------------------------------------------------------------------------------

def compare_sort(packed, iou_thr, max_age, min_hits, final_cap=0, promotion=None):
    got = runtime.sort_track(packed, iou_thr, max_age, min_hits, final_cap=final_cap, promotion=promotion)
    want = c_oracle.sort_track(packed, iou_thr, max_age, min_hits, final_cap=final_cap, promotion=promotion)
    assert want["status"] == 0
    np.testing.assert_array_equal(got["out_count"], want["out_count"])
    np.testing.assert_array_equal(got["created"], want["created"])
    np.testing.assert_array_equal(got["first_img"], want["first_img"])
    rows, _ = packing.valid_row_index(packed.det_start, want["out_count"])
    np.testing.assert_array_equal(got["out_birth"][rows], want["out_birth"][rows])   # assignments
    np.testing.assert_array_equal(got["out_box"][rows], want["out_box"][rows])       # states -> boxes: bit-exact
    np.testing.assert_allclose(got["out_score"][rows], want["out_score"][rows], rtol=1e-9, atol=0)
    ids, nxt = packing.assign_ids(packed.stream_img_offsets, packed.n_classes, packed.det_start, want["out_count"],
                                  want["created"], want["first_img"], packed.class_rank, want["out_birth"])
    np.testing.assert_array_equal(got["ids"], ids)
    assert got["id_next"] == nxt
    # dense list built on the device == the host construction from the per-slot arrays
    dense = packing.unpack_tracks(packed, want["out_box"], want["out_score"], want["out_count"], want["first_img"],
                                  ids)
    assert got["n_rows"] == len(dense)
    np.testing.assert_array_equal(got["rows_id"], np.array([int(r["object_id"]) for r in dense], np.int64))
    np.testing.assert_array_equal(got["rows_cat"], np.array([r["category_id"] for r in dense], np.int32))
    np.testing.assert_array_equal(got["rows_box"], np.array([r["bbox"] for r in dense], np.float64).reshape(-1, 4))
    iid = {}
    for s_, (seg, cam) in enumerate(packed.streams):
        for img in range(int(packed.stream_img_offsets[s_]), int(packed.stream_img_offsets[s_ + 1])):
            iid['%s/%i/%s' % (seg, packed.frame_ids[img], cam)] = img
    np.testing.assert_array_equal(got["rows_img"], np.array([iid[r["image_id"]] for r in dense], np.int32))
    if final_cap:
        np.testing.assert_array_equal(got["final_count"], want["final_count"])
        for q, t in enumerate(want["final_count"]):
            t = min(int(t), final_cap)
            np.testing.assert_array_equal(got["final_state"][q, :t], want["final_state"][q, :t])
    return got


@pytest.mark.parametrize("promotion", helpers.PROMOTIONS)
@pytest.mark.parametrize("name", TRACK_CASES)
def test_sort_matches_reference_golden(name, promotion):
    g = golden_io.load(name)
    scene = helpers.golden_scene(g)
    packed = synth.tracks_from_submission(scene, scene.submissions[0], helpers.SCORE_THR)
    res = compare_sort(packed, helpers.IOU_THR, int(g["max_age"]), int(g["min_hits"]), final_cap=64, promotion=promotion)
    got = helpers.track_rows_as_arrays(packed, res, scene.image_ids(), ids=res["ids"])
    want = helpers.golden_tracks(g, promotion)
    helpers.assert_tracks_equal(got, want, score_rtol=1e-9, box_exact=True)


@pytest.mark.parametrize("promotion", helpers.PROMOTIONS)
def test_sort_full_size_c1_matches_reference_golden(promotion):
    """BASELINE.json config C1 at full size (one segment: 5 cameras x 200 frames, ~110 detections per frame in)
    against the reference's own output (tests/golden/big_c1.npz) and, row for row, against the C oracle."""
    g = golden_io.load("big_c1")
    scene = synth.make_scene(synth.preset(str(g["preset"]), n_segments=1, seed=int(g["seed"])))
    packed = synth.tracks_from_submission(scene, scene.submissions[0], helpers.SCORE_THR)
    res = compare_sort(packed, helpers.IOU_THR, int(g["max_age"]), int(g["min_hits"]), promotion=promotion)
    got = helpers.track_rows_as_arrays(packed, res, scene.image_ids(), ids=res["ids"])
    helpers.assert_big_tracks_equal(got, g, promotion)


def test_ensemble_full_size_c2_matches_reference_golden():
    """BASELINE.json config C2 at full size (3 submissions of one segment, 274 k boxes in) against the reference's own
    output (tests/golden/big_c2.npz)."""
    g = golden_io.load("big_c2")
    scene = synth.make_scene(synth.preset(str(g["preset"]), n_segments=1, seed=int(g["seed"])))
    groups = synth.groups_from_scene(scene, None, float(g["min_score"]))
    res = runtime.softnms_groups(groups.group_offsets, groups.rows, float(g["iou_thresh"]), float(g["cut"]),
                                 float(g["min_score"]), 4, helpers.SCORE_THR, max_group=groups.max_group)
    got = helpers.ensemble_rows_as_arrays(groups.group_offsets, res, scene.n_img,
                                          image_order=helpers.sorted_image_order(scene.image_ids()))
    helpers.assert_big_ensemble_equal(got, g)


@pytest.mark.parametrize("seed,max_age,min_hits", [(101, 2, 0), (102, 1, 3), (103, 0, 0), (104, 5, 1)])
def test_sort_seeded_vs_oracle(seed, max_age, min_hits):
    cfg = synth.SynthConfig(n_segments=2, n_frames=50, n_submissions=1, objects_per_frame=70.0, seed=seed)
    scene = synth.make_scene(cfg)
    packed = synth.tracks_from_submission(scene, scene.submissions[0], helpers.SCORE_THR)
    compare_sort(packed, helpers.IOU_THR, max_age, min_hits, final_cap=160)


@pytest.mark.parametrize("seed", range(200, 212))
def test_sort_random_configurations_vs_oracle(seed):
    # thresholds, lifecycle parameters, densities and class mixes drawn at random: IoU thresholds of 0
    # (zero-IoU matches are kept, Munkres tie-breaking decides), 1 (nothing ever matches), long max_age
    # (trackers outnumber detections), short lives and many misses (detections outnumber trackers)
    rng = np.random.default_rng(seed)
    mix = rng.dirichlet([1.0, 1.0, 0.3, 0.6])
    cfg = synth.SynthConfig(n_segments=1, cameras=("FRONT", "SIDE_LEFT"), n_frames=int(rng.integers(20, 45)),
                            n_submissions=1, objects_per_frame=float(rng.uniform(5, 130)),
                            class_mix=tuple(float(v) for v in mix), mean_life=float(rng.uniform(3, 50)),
                            p_miss=float(rng.uniform(0.0, 0.5)), jitter=float(rng.uniform(0.5, 6.0)),
                            fp_per_frame=float(rng.uniform(0, 15)), confident=float(rng.uniform(0.5, 0.98)),
                            size_range=(float(rng.uniform(8, 30)), float(rng.uniform(60, 250))), seed=seed)
    scene = synth.make_scene(cfg)
    score_thr = [float(rng.choice([0.0, 0.3, 0.6, 0.9, 0.95])) for _ in range(4)]
    iou_thr = [float(rng.choice([0.0, 0.01, 0.3, 0.5, 1.0])) for _ in range(4)]
    packed = synth.tracks_from_submission(scene, scene.submissions[0], score_thr)
    compare_sort(packed, iou_thr, int(rng.integers(0, 5)), int(rng.integers(0, 4)), final_cap=128,
                 promotion=helpers.PROMOTIONS[seed % 2])


def test_sort_ragged_and_empty_streams():
    # empty stream, a stream whose detections are all filtered, missing images, a late-starting category
    cfg = synth.SynthConfig(n_segments=1, cameras=("FRONT", "SIDE_LEFT", "SIDE_RIGHT"), n_frames=25,
                            n_submissions=1, objects_per_frame=20.0, seed=5)
    scene = synth.make_scene(cfg)
    sub = scene.submissions[0]
    keep = np.ones(len(sub.score), bool)
    F = cfg.n_frames
    keep[sub.image_index // F == 1] = False                        # stream 1: no detections at all
    keep[(sub.image_index % F) % 7 == 3] = False                   # images absent from the JSON
    keep[(sub.category == 2) & (sub.image_index % F < 12)] = False  # pedestrians appear late
    sub2 = synth.Submission(sub.image_index[keep], sub.category[keep], sub.bbox[keep], sub.score[keep].copy())
    sub2.score[(sub2.image_index // F == 2)] = 0.011               # stream 2: everything below threshold
    packed = synth.tracks_from_submission(scene, sub2, helpers.SCORE_THR)
    assert packed.img_exists.min() == 0
    res = compare_sort(packed, helpers.IOU_THR, 2, 0, final_cap=64)
    assert res["out_count"].reshape(-1, 4)[F:2 * F].sum() == 0
    # no streams at all
    empty = packing.PackedTracks(0, 4, [], np.zeros(0, np.int64), np.zeros(1, np.int32), np.zeros(0, np.int32),
                                 np.zeros(0, np.int32), np.zeros((0, 4), np.float32), np.zeros((0, 2)), None,
                                 np.zeros(0, np.int32), 0)
    out = runtime.sort_track(empty, helpers.IOU_THR, 2, 0)
    assert out["id_next"] == 0 and len(out["ids"]) == 0


def test_sort_medium_density_matches_oracle():
    # 80-128 detections / trackers per category: assignment problems that use the shared-memory solver
    # with the cost matrix in the global slab (n * pitch > 6144 floats, m <= 128)
    cfg = synth.SynthConfig(n_segments=1, cameras=("FRONT", "SIDE_RIGHT"), n_frames=14, n_submissions=1,
                            objects_per_frame=340.0, class_mix=(0.5, 0.45, 0.0, 0.05), size_range=(12.0, 90.0),
                            mean_life=25.0, seed=19)
    scene = synth.make_scene(cfg)
    packed = synth.tracks_from_submission(scene, scene.submissions[0], helpers.SCORE_THR)
    per_class = packed.det_count.reshape(-1, 4)
    assert 80 <= per_class[:, 0].max() <= 128 and per_class[:, 0].mean() > 90
    compare_sort(packed, helpers.IOU_THR, 2, 0)


@pytest.mark.parametrize("opf,seed,top", [(160.0, 2, 103), (170.0, 1, 104), (170.0, 6, 110)])
def test_sort_images_that_outgrow_tensor_memory_match_oracle(opf, seed, top):
    # vehicles peak just above W2T_NARROW_DETS at a few images: the warp kernel keeps the sub-stream and spills those
    # images' cost matrices to its global-memory area (top <= 104), or leaves it to the cluster kernel (110)
    cfg = synth.SynthConfig(n_segments=1, cameras=("FRONT", "SIDE_LEFT"), n_frames=80, n_submissions=1,
                            objects_per_frame=opf, class_mix=(0.7, 0.25, 0.0, 0.05), seed=seed)
    scene = synth.make_scene(cfg)
    packed = synth.tracks_from_submission(scene, scene.submissions[0], helpers.SCORE_THR)
    vehicles = packed.det_count.reshape(-1, 4)[:, 0]
    assert vehicles.max() == top and 0 < (vehicles > 96).sum() <= 32
    compare_sort(packed, helpers.IOU_THR, 2, 0)


def test_sort_crowded_matches_oracle():
    # C4-shaped miniature: ~1000 dets/frame, hundreds of live tracks per category
    cfg = synth.preset("c4", cameras=("FRONT",), n_frames=8, seed=9)
    scene = synth.make_scene(cfg)
    packed = synth.tracks_from_submission(scene, scene.submissions[0], helpers.SCORE_THR)
    assert packed.det_count.max() > 300
    compare_sort(packed, helpers.IOU_THR, 2, 0)


def test_sort_is_invariant_to_sharding():
    # tracking a subset of streams alone gives the same rows; ids differ only by the scan base
    cfg = synth.SynthConfig(n_segments=2, cameras=("FRONT", "SIDE_LEFT"), n_frames=30, n_submissions=1,
                            objects_per_frame=40.0, seed=6)
    scene = synth.make_scene(cfg)
    packed = synth.tracks_from_submission(scene, scene.submissions[0], helpers.SCORE_THR)
    full = runtime.sort_track(packed, helpers.IOU_THR, 2, 0)
    F, NC = cfg.n_frames, 4
    base = 0
    for s in range(packed.n_streams):
        g0, g1 = s * F * NC, (s + 1) * F * NC
        r0 = int(packed.det_start[g0])
        r1 = int(packed.det_start[g1]) if g1 < len(packed.det_start) else packed.n_rows
        part = packing.PackedTracks(1, NC, [packed.streams[s]], packed.frame_ids[s * F:(s + 1) * F],
                                    np.array([0, F], np.int32), packed.det_start[g0:g1] - r0,
                                    packed.det_count[g0:g1], packed.det_box[r0:r1], packed.cam_wh[s:s + 1],
                                    packed.img_exists[s * F:(s + 1) * F], packed.class_rank[s * NC:(s + 1) * NC],
                                    r1 - r0)
        one = runtime.sort_track(part, helpers.IOU_THR, 2, 0, id_base=base)
        base = one["id_next"]
        np.testing.assert_array_equal(one["out_count"], full["out_count"][g0:g1])
        rows, _ = packing.valid_row_index(part.det_start, one["out_count"])
        np.testing.assert_array_equal(one["out_box"][rows], full["out_box"][r0:r1][rows])
        np.testing.assert_array_equal(one["ids"][rows], full["ids"][r0:r1][rows])
    assert base == full["id_next"]


# ---- soft-NMS ensemble stage --------------------------------------------------------------------

def compare_nms(groups, iou_thresh, cut, min_score):
    got = runtime.softnms_groups(groups.group_offsets, groups.rows, iou_thresh, cut, min_score, 4,
                                 helpers.SCORE_THR, max_group=groups.max_group)
    want = c_oracle.softnms_groups(groups.group_offsets, groups.rows, iou_thresh, cut, min_score, 4,
                                   helpers.SCORE_THR)
    np.testing.assert_array_equal(got["merged"], want["merged"])          # decayed scores: bit-exact
    np.testing.assert_array_equal(got["src_index"], want["src_index"])    # kept-box order
    np.testing.assert_array_equal(got["ens_count"], want["ens_count"])
    np.testing.assert_array_equal(got["trk_count"], want["trk_count"])
    np.testing.assert_array_equal(got["img_exists"], want["img_exists"])
    off = groups.group_offsets.astype(np.int64)
    rows, _ = packing.valid_row_index(off[:-1], want["ens_count"])
    np.testing.assert_array_equal(got["ens_box"][rows], want["ens_box"][rows])
    np.testing.assert_array_equal(got["ens_score"][rows], want["ens_score"][rows])
    rows, _ = packing.valid_row_index(off[:-1], want["trk_count"])
    np.testing.assert_array_equal(got["trk_box"][rows], want["trk_box"][rows])
    return got


@pytest.mark.parametrize("name", ENS_CASES)
def test_softnms_matches_reference_golden(name):
    g = golden_io.load(name)
    scene = helpers.golden_scene(g)
    groups = synth.groups_from_scene(scene, list(g["weights"]), float(g["min_score"]))
    res = compare_nms(groups, float(g["iou_thresh"]), float(g["cut"]), float(g["min_score"]))
    got = helpers.ensemble_rows_as_arrays(groups.group_offsets, res, scene.n_img,
                                          image_order=helpers.sorted_image_order(scene.image_ids()))
    for k in ("img", "cat", "bbox", "score"):
        np.testing.assert_array_equal(got[k], g["out_" + k])


def test_softnms_edge_groups():
    # n = 0, 1, 2; identical boxes (IoU 1 -> weight 0); disjoint boxes; IoU >= cut
    rows = np.array([
        [0.9, 10, 10, 20, 20],                               # group 1: single box
        [0.9, 10, 10, 20, 20], [0.8, 10, 10, 20, 20],        # group 2: identical boxes
        [0.9, 0, 0, 10, 10], [0.8, 100, 100, 10, 10],        # group 3: disjoint
        [0.9, 0, 0, 100, 100], [0.8, 2, 2, 100, 100], [0.7, 50, 0, 100, 100], [0.3, 1, 1, 99, 99],
    ], np.float64)
    offsets = np.array([0, 0, 1, 3, 5, 9, 9, 9, 9], np.int32)   # 8 groups = 2 images x 4 categories
    groups = packing.PackedGroups(None, [1, 2, 3, 4], offsets, rows, 4)
    res = compare_nms(groups, 0.5, 0.9, 0.0)
    assert res["ens_count"].tolist() == [0, 1, 1, 2, 2, 0, 0, 0]   # IoU >= cut zeroes two boxes
    assert res["merged"][2, 0] == 0.0 and res["merged"][1, 0] == 0.9


def test_softnms_large_group_tta_shaped():
    # C5-shaped: 5 submissions, thousands of boxes per image
    cfg = synth.preset("c5", cameras=("FRONT",), n_frames=2, seed=4)
    scene = synth.make_scene(cfg)
    groups = synth.groups_from_scene(scene, None, 0.01)
    assert groups.max_group > 1000
    compare_nms(groups, 0.5, 0.9, 0.01)


def test_softnms_group_beyond_shared_memory():
    # a group of 5000 boxes (shared memory holds 3401): the global-memory pass takes it, its neighbours stay on the
    # regular launch; soft and hard branch against the C oracle, bit for bit
    rng = np.random.default_rng(12)
    sizes = [7, 5000, 0, 120, 3600]
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    n = int(offs[-1])
    c = rng.uniform(0, 1900, (n, 2))
    rows = np.c_[np.round(rng.uniform(0.02, 1, n), 5), c.astype(np.int64), rng.integers(8, 160, (n, 2))].astype(np.float64)
    got = runtime.softnms_groups(offs, rows, 0.5, 0.9, 0.01, 4, helpers.SCORE_THR, max_group=max(sizes))
    want = c_oracle.softnms_groups(offs, rows, 0.5, 0.9, 0.01, 4, helpers.SCORE_THR)
    for k in ("ens_count", "trk_count", "kept_count"):
        np.testing.assert_array_equal(got[k], want[k])
    for g in range(len(sizes)):
        o, cnt = int(offs[g]), int(want["ens_count"][g])
        np.testing.assert_array_equal(got["ens_box"][o:o + cnt], want["ens_box"][o:o + cnt])
        np.testing.assert_array_equal(got["ens_score"][o:o + cnt], want["ens_score"][o:o + cnt])
        np.testing.assert_array_equal(got["merged"][o:o + sizes[g]], want["merged"][o:o + sizes[g]])
    got = runtime.hardnms_groups(offs, rows, 0.5, max_group=max(sizes), box_format=0)
    want = c_oracle.softnms_groups(offs, rows, 0.5, 1.0, -np.inf, box_format=0, hard=True)
    np.testing.assert_array_equal(got["kept_count"], want["kept_count"])
    for g in range(len(sizes)):
        o, cnt = int(offs[g]), int(want["kept_count"][g])
        np.testing.assert_array_equal(got["merged"][o:o + cnt], want["merged"][o:o + cnt])


def test_softnms_size_classes_and_their_boundaries():
    # >= 16384 groups of mixed sizes: one launch per size class (<= 32 boxes: one-warp CTAs, <= 96, the rest).  Sizes
    # on both sides of every bound, empty groups, tied scores (the integer-key ranking of packed rows must break
    # ties like the canonical rule), every branch of the kernel: soft, soft with removals, top_k, hard.
    rng = np.random.default_rng(77)
    sizes = rng.choice([0, 1, 2, 31, 32, 33, 64, 95, 96, 97, 128, 129, 150], size=17000, p=[.3, .2, .2, .05, .05, .05, .03, .02, .02, .02, .02, .02, .02])
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    n = int(offs[-1])
    xy = rng.integers(0, 1700, (n, 2))
    wh = rng.integers(10, 220, (n, 2))
    score = np.round(rng.uniform(0.02, 1, n), 2)          # 2 decimals: plenty of ties inside a group
    rows = np.c_[score, xy, wh].astype(np.float64)
    packed = packing.packed_rows(rows)
    assert packed is not None and int(sizes.max()) > 96
    mg = int(sizes.max())
    want = c_oracle.softnms_groups(offs, rows, 0.5, 0.9, 0.01, 4, helpers.SCORE_THR)
    for data in (rows, packed):
        got = runtime.softnms_groups(offs, data, 0.5, 0.9, 0.01, 4, helpers.SCORE_THR, max_group=mg)
        for k in ("merged", "src_index", "ens_count", "trk_count", "kept_count", "img_exists"):
            np.testing.assert_array_equal(got[k], want[k])
        idx, _ = packing.valid_row_index(offs[:-1].astype(np.int64), want["ens_count"])
        np.testing.assert_array_equal(got["ens_box"][idx], want["ens_box"][idx])
        np.testing.assert_array_equal(got["ens_score"][idx], want["ens_score"][idx])
    # removals (conf_thresh > 0) and top_k: the order-dependent branch
    want = c_oracle.softnms_groups(offs, rows, 0.4, 0.8, 0.0, top_k=40, conf_thresh=0.3)
    got = runtime.softnms_groups(offs, rows, 0.4, 0.8, 0.0, max_group=mg, top_k=40, conf_thresh=0.3)
    np.testing.assert_array_equal(got["kept_count"], want["kept_count"])
    idx, _ = packing.valid_row_index(offs[:-1].astype(np.int64), want["kept_count"])
    np.testing.assert_array_equal(got["merged"][idx], want["merged"][idx])
    np.testing.assert_array_equal(got["src_index"][idx], want["src_index"][idx])
    # hard NMS
    want = c_oracle.softnms_groups(offs, rows, 0.5, 1.0, -np.inf, box_format=0, hard=True)
    got = runtime.hardnms_groups(offs, rows, 0.5, max_group=mg, box_format=0)
    np.testing.assert_array_equal(got["kept_count"], want["kept_count"])
    idx, _ = packing.valid_row_index(offs[:-1].astype(np.int64), want["kept_count"])
    np.testing.assert_array_equal(got["merged"][idx], want["merged"][idx])


def test_softnms_rejects_unsupported_scores_loudly():
    rows = np.array([[-0.5, 0, 0, 10, 10], [0.5, 0, 0, 10, 10]], np.float64)
    with pytest.raises(Exception):
        runtime.softnms_groups(np.array([0, 2], np.int32), rows, 0.5, 0.9, 0.0)


# ---- ensemble -> SORT on the device ----------------------------------------------------------------

@pytest.mark.parametrize("promotion", helpers.PROMOTIONS)
def test_pipeline_matches_reference_golden_end_to_end(promotion):
    import test_oracle as T
    g = golden_io.load("pipeline_small")
    scene = helpers.golden_scene(g)
    groups = synth.groups_from_scene(scene, None, 0.01)
    res = runtime.ensemble_and_track(groups.group_offsets, groups.rows, scene.stream_img_offsets, scene.cam_wh(), 4,
                                     0.5, 0.9, 0.01, helpers.SCORE_THR, helpers.IOU_THR, 2, 0,
                                     max_group=groups.max_group, promotion=promotion)
    ens = helpers.ensemble_rows_as_arrays(groups.group_offsets, res, scene.n_img,
                                          image_order=helpers.sorted_image_order(scene.image_ids()))
    for k in ("img", "cat", "bbox", "score"):
        np.testing.assert_array_equal(ens[k], g["ens_" + k])
    packed = packing.PackedTracks(
        n_streams=scene.n_streams, n_classes=4, streams=scene.streams(), frame_ids=scene.frame_ids,
        stream_img_offsets=scene.stream_img_offsets, det_start=res["det_start"], det_count=res["trk_count"],
        det_box=np.zeros((len(groups.rows), 4), np.float32), cam_wh=scene.cam_wh(), img_exists=res["img_exists"],
        class_rank=None, n_rows=len(groups.rows))
    order = T.stream_order_of_sorted_images(scene)
    got = T.rows_in_stream_order(packed, res, scene, order)
    want = helpers.golden_tracks(g, promotion)
    helpers.assert_tracks_equal(got, want, score_rtol=1e-9, box_exact=True)


def test_pipelined_chunks_equal_the_single_pass():
    # chunked H2D / kernels / D2H overlap (runtime.ensemble_and_track_pipelined) changes nothing in the result
    cfg = synth.SynthConfig(n_segments=3, cameras=("FRONT", "SIDE_LEFT", "SIDE_RIGHT"), n_frames=30, n_submissions=3,
                            objects_per_frame=40.0, seed=44)
    scene = synth.make_scene(cfg)
    groups = synth.groups_from_scene(scene, None, 0.01)
    args = (groups.group_offsets, groups.rows, scene.stream_img_offsets, scene.cam_wh(), 4, 0.5, 0.9, 0.01,
            helpers.SCORE_THR, helpers.IOU_THR, 2, 0)
    want = runtime.ensemble_and_track(*args, max_group=groups.max_group, want_ensemble=False, raw=False, id_base=7)
    want = {k: np.array(want[k]) for k in ("rows_box", "rows_score", "rows_id", "rows_img", "rows_cat")} | \
        {"n_rows": want["n_rows"], "id_next": want["id_next"]}
    for n_chunks in (1, 2, 4, 9, 50):
        got = runtime.ensemble_and_track_pipelined(*args, max_group=groups.max_group, id_base=7, n_chunks=n_chunks)
        assert got["n_rows"] == want["n_rows"] and got["id_next"] == want["id_next"]
        for k in ("rows_box", "rows_score", "rows_id", "rows_img", "rows_cat"):
            np.testing.assert_array_equal(got[k], want[k])


def test_pipelined_crowded_streams_and_launch_orders():
    """Crowded sub-streams (more than W2T_WIDE_DETS detections in an image: wide CTAs, assignment problems in
    global memory) through the pipelined path, with every launch-order policy: same rows as the single pass."""
    cfg = synth.preset("c4", n_segments=3, cameras=("FRONT", "SIDE_LEFT"), n_frames=5, n_submissions=2, seed=9)
    scene = synth.make_scene(cfg)
    groups = synth.groups_from_scene(scene, None, 0.01)
    assert int(np.diff(groups.group_offsets).max()) > 2 * 320
    args = (groups.group_offsets, groups.rows, scene.stream_img_offsets, scene.cam_wh(), 4, 0.5, 0.9, 0.01,
            helpers.SCORE_THR, helpers.IOU_THR, 2, 0)
    want = runtime.ensemble_and_track(*args, max_group=groups.max_group, want_ensemble=False, raw=False)
    want = {k: np.array(want[k]) for k in ("rows_box", "rows_score", "rows_id", "rows_img", "rows_cat")}
    assert len(want["rows_id"]) > 0
    for n_chunks, hoist, by_work in ((3, 1.0, False), (3, 0.5, True), (2, 0.0, False), ([0.2, 0.8], 1.0, True)):
        got = runtime.ensemble_and_track_pipelined(*args, max_group=groups.max_group, n_chunks=n_chunks, hoist=hoist,
                                                   hoist_by_work=by_work)
        for k, v in want.items():
            np.testing.assert_array_equal(got[k], v)


def test_compact_input_rows_equal_float64_rows():
    cfg = synth.SynthConfig(n_segments=2, cameras=("FRONT", "SIDE_LEFT"), n_frames=20, n_submissions=3,
                            objects_per_frame=40.0, seed=45)
    scene = synth.make_scene(cfg)
    groups = synth.groups_from_scene(scene, None, 0.01)
    compact = packing.compact_rows(groups.rows)
    a = compare_nms(groups, 0.5, 0.9, 0.01)
    b = runtime.softnms_groups(groups.group_offsets, compact, 0.5, 0.9, 0.01, 4, helpers.SCORE_THR,
                               max_group=groups.max_group)
    for k in ("merged", "src_index", "ens_count", "trk_count", "img_exists"):
        np.testing.assert_array_equal(a[k], b[k])
    args = (scene.stream_img_offsets, scene.cam_wh(), 4, 0.5, 0.9, 0.01, helpers.SCORE_THR, helpers.IOU_THR, 2, 0)
    want = runtime.ensemble_and_track(groups.group_offsets, groups.rows, *args, max_group=groups.max_group,
                                      want_ensemble=False, raw=False)
    want = {k: np.array(want[k]) for k in ("rows_box", "rows_score", "rows_id", "rows_img", "rows_cat")}
    got = runtime.ensemble_and_track_pipelined(groups.group_offsets, compact, *args, max_group=groups.max_group,
                                               n_chunks=3)
    for k, v in want.items():
        np.testing.assert_array_equal(got[k], v)


def test_packed_input_rows_equal_float64_rows():
    """8-byte packed rows (W2T_BOX_LTWH_P64) through soft-NMS, the fused ensemble -> SORT path and the
    pipelined host path: bit-identical to the float64 rows and to the oracle."""
    cfg = synth.SynthConfig(n_segments=2, cameras=("FRONT", "SIDE_LEFT"), n_frames=20, n_submissions=3,
                            objects_per_frame=40.0, seed=46)
    scene = synth.make_scene(cfg)
    groups = synth.groups_from_scene(scene, None, 0.01)
    packed = packing.packed_rows(groups.rows)
    assert packed is not None and packed.dtype == np.uint64
    a = compare_nms(groups, 0.5, 0.9, 0.01)        # float64 rows, checked against the oracle
    b = runtime.softnms_groups(groups.group_offsets, packed, 0.5, 0.9, 0.01, 4, helpers.SCORE_THR,
                               max_group=groups.max_group)
    for k in ("merged", "src_index", "ens_count", "trk_count", "img_exists"):
        np.testing.assert_array_equal(a[k], b[k])
    args = (scene.stream_img_offsets, scene.cam_wh(), 4, 0.5, 0.9, 0.01, helpers.SCORE_THR, helpers.IOU_THR, 2, 0)
    want = runtime.ensemble_and_track(groups.group_offsets, groups.rows, *args, max_group=groups.max_group,
                                      want_ensemble=False, raw=False)
    want = {k: np.array(want[k]) for k in ("rows_box", "rows_score", "rows_id", "rows_img", "rows_cat")}
    got = runtime.ensemble_and_track_pipelined(groups.group_offsets, packed, *args, max_group=groups.max_group,
                                               n_chunks=3)
    for k, v in want.items():
        np.testing.assert_array_equal(got[k], v)

"""Shared helpers: golden fixtures -> packed inputs, packed results -> comparable arrays."""
import json

import numpy as np

import golden_io
from waymo_2d_tracking_b200 import packing, synth

SCORE_THR = [0.95, 0.6, 1.0, 0.9]
IOU_THR = [0.01, 0.01, 1.0, 0.0]
PROMOTIONS = ("legacy", "nep50")


def golden_tracks(g, promotion=None):
    """Tracker outputs of a fixture: ``out_*`` = the reference's files under this container's NumPy 2 (NEP 50),
    ``leg_*`` = the shim's emulation of the reference's pinned NumPy 1.x environment (tests/golden/make_golden.py).
    ``promotion=None``: the regime the product runs by default."""
    from waymo_2d_tracking_b200 import _abi
    legacy = _abi.promotion_code(promotion) == _abi.W2T_PROMOTION_LEGACY
    prefix = "leg_" if legacy else "out_"
    return {k[4:]: v for k, v in g.items() if k.startswith(prefix)}


def golden_scene(g):
    """Rebuild the synthetic scene of a fixture and pin its submissions to the stored arrays."""
    cfg = json.loads(str(g["cfg"]))
    for k in ("cameras", "class_mix", "size_range"):
        if k in cfg:
            cfg[k] = tuple(cfg[k])
    scene = synth.make_scene(synth.SynthConfig(**cfg))
    assert list(g["image_ids"]) == scene.image_ids(), "synthetic generator drifted from the fixture"
    subs = []
    for k in range(int(g["n_sub"])):
        subs.append(synth.Submission(g["s%d_img" % k], g["s%d_cat" % k], g["s%d_bbox" % k], g["s%d_score" % k]))
    scene.submissions = subs
    return scene


def track_rows_as_arrays(packed, res, image_ids, ids=None):
    """Packed SORT result -> arrays in the reference's output order."""
    if ids is None:
        ids, _ = packing.assign_ids(packed.stream_img_offsets, packed.n_classes, packed.det_start, res["out_count"],
                                    res["created"], res["first_img"], packed.class_rank, res["out_birth"])
    rows = packing.unpack_tracks(packed, res["out_box"], res["out_score"], res["out_count"], res["first_img"], ids)
    return golden_io.tracks_to_arrays(rows, image_ids)


def assert_tracks_equal(got, want, score_rtol=1e-9, box_exact=True):
    assert len(got["img"]) == len(want["img"]), (len(got["img"]), len(want["img"]))
    np.testing.assert_array_equal(got["img"], want["img"])
    np.testing.assert_array_equal(got["cat"], want["cat"])
    np.testing.assert_array_equal(got["oid"], want["oid"])       # track ids: bit-exact
    if box_exact:
        np.testing.assert_array_equal(got["bbox"], want["bbox"])
    else:
        np.testing.assert_allclose(got["bbox"], want["bbox"], rtol=1e-9, atol=1e-9)
    # the confidence goes through exp(): libm / NumPy SIMD / CUDA differ in the last ulp
    np.testing.assert_allclose(got["score"], want["score"], rtol=score_rtol, atol=0)


def ensemble_rows_as_arrays(group_offsets, res, n_img, n_classes=4, image_order=None):
    """Packed soft-NMS result -> arrays (img, cat, bbox, score) image by image, category by category."""
    if image_order is None:
        image_order = range(n_img)
    img, cat, bbox, score = [], [], [], []
    for i in image_order:
        for c in range(n_classes):
            g = i * n_classes + c
            o, k = int(group_offsets[g]), int(res["ens_count"][g])
            img += [i] * k
            cat += [c + 1] * k
            bbox.append(res["ens_box"][o:o + k])
            score.append(res["ens_score"][o:o + k])
    return dict(img=np.asarray(img, np.int32), cat=np.asarray(cat, np.int32),
                bbox=np.concatenate(bbox).astype(np.int64) if bbox else np.zeros((0, 4), np.int64),
                score=np.concatenate(score) if score else np.zeros(0))


def sorted_image_order(image_ids):
    return sorted(range(len(image_ids)), key=lambda i: image_ids[i])


def assert_big_tracks_equal(got, g, promotion, box_exact=True):
    """Full-size tracking fixture (tests/golden/make_golden.py, BIG_TRACK): ids / images / categories of every
    row, boxes and confidences of the rows of every ``sample``-th image."""
    from waymo_2d_tracking_b200 import _abi
    pre = "leg_" if _abi.promotion_code(promotion) == _abi.W2T_PROMOTION_LEGACY else "out_"
    assert len(got["img"]) == len(g[pre + "img"])
    np.testing.assert_array_equal(got["img"], g[pre + "img"])
    np.testing.assert_array_equal(got["cat"], g[pre + "cat"])
    np.testing.assert_array_equal(got["oid"], g[pre + "oid"])          # track ids: bit-exact
    keep = (got["img"] % int(g["sample"])) == 0
    if box_exact:
        np.testing.assert_array_equal(got["bbox"][keep], g[pre + "bbox_s"])
    else:
        np.testing.assert_allclose(got["bbox"][keep], g[pre + "bbox_s"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(got["score"][keep], g[pre + "score_s"], rtol=1e-9, atol=0)


def assert_big_ensemble_equal(got, g):
    """Full-size ensemble fixture (BIG_ENSEMBLE): images / categories of every row, boxes and scores of the rows
    of every ``sample``-th image; all bit-exact."""
    np.testing.assert_array_equal(got["img"], g["out_img"])
    np.testing.assert_array_equal(got["cat"], g["out_cat"])
    keep = (got["img"] % int(g["sample"])) == 0
    np.testing.assert_array_equal(got["bbox"][keep], g["out_bbox_s"])
    np.testing.assert_array_equal(got["score"][keep], g["out_score_s"])

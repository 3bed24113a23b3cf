"""Generate the committed golden fixtures by EXECUTING THE REFERENCE'S OWN FILES.

Run in the dev container only (needs /root/reference):

    python tests/golden/make_golden.py

For every case the inputs are synthetic submissions (``waymo_2d_tracking_b200.synth``,
seeded) written to JSON exactly as the reference expects them; the outputs are what the
reference's own code returns for them through ``oracle.ref_shim``:

* tracking cases : ``tracking/utils.py`` read_data_file + track_sort (-> tracker_sort.py,
  sort.py) with the two absent third-party pieces restated (oracle/munkres.py,
  oracle/kalman.py — parity unpinned there, see oracle/__init__.py);
* ensemble cases : ``detnet/ensemble.py`` convert_submission + ensemble (-> tta.py,
  box_utils.py), 100 % reference code on torch CPU.  ``torch.Tensor.sort`` is forced stable
  (canonical tie rule); all cases but ``ensemble_ties`` have distinct scores,
  for which the unpatched reference gives the same result.

Inputs and outputs are stored as exact NumPy arrays (npz), a few hundred KB in total.
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_shim  # noqa: E402
from waymo_2d_tracking_b200 import synth  # noqa: E402
import golden_io  # noqa: E402

SCORE_THR = [0.95, 0.6, 1.0, 0.9]
IOU_THR = [0.01, 0.01, 1.0, 0.0]

TRACK_CASES = {
    # README recipe: --max-age=2 --min-hits=0
    "track_c1_small": dict(cfg=dict(n_segments=1, cameras=("FRONT", "SIDE_LEFT"), n_frames=40, n_submissions=1,
                                    objects_per_frame=60.0, seed=11), max_age=2, min_hits=0),
    # upstream SORT defaults exercise the min_hits / hit_streak gate
    "track_minhits": dict(cfg=dict(n_segments=2, cameras=("FRONT_RIGHT",), n_frames=30, n_submissions=1,
                                   objects_per_frame=30.0, seed=12), max_age=1, min_hits=3),
    # cyclist-heavy and sparse: iou_threshold 0.0 keeps zero-IoU assignments, so Munkres ties decide
    "track_cyclist_ties": dict(cfg=dict(n_segments=1, cameras=("SIDE_RIGHT", "FRONT_LEFT"), n_frames=40,
                                        n_submissions=1, objects_per_frame=25.0, class_mix=(0.1, 0.1, 0.1, 0.7),
                                        mean_life=6.0, p_miss=0.3, seed=13), max_age=2, min_hits=0),
    # dense: more detections than trackers and vice versa, larger assignment problems
    "track_dense": dict(cfg=dict(n_segments=1, cameras=("FRONT",), n_frames=16, n_submissions=1,
                                 objects_per_frame=260.0, size_range=(12.0, 80.0), mean_life=10.0, p_miss=0.25,
                                 seed=14), max_age=2, min_hits=0),
}

ENSEMBLE_CASES = {
    "ensemble_c2_small": dict(cfg=dict(n_segments=1, cameras=("FRONT", "SIDE_LEFT"), n_frames=8, n_submissions=3,
                                       objects_per_frame=45.0, seed=21),
                              min_score=0.01, iou_thresh=0.5, cut=0.9, weights=None),
    "ensemble_weighted": dict(cfg=dict(n_segments=1, cameras=("FRONT_LEFT",), n_frames=6, n_submissions=4,
                                       objects_per_frame=30.0, seed=22),
                              min_score=0.05, iou_thresh=0.4, cut=1.0, weights=[2, 1, 4, 3]),
    "ensemble_ties": dict(cfg=dict(n_segments=1, cameras=("FRONT",), n_frames=4, n_submissions=3,
                                   objects_per_frame=40.0, seed=23),
                          min_score=0.01, iou_thresh=0.5, cut=0.9, weights=None, round_scores=2),
}


# the CLI's other two methods (ensemble.py:94,138-140), same generator
METHOD_CASES = {
    "ensemble_nms": dict(cfg=dict(n_segments=1, cameras=("FRONT", "SIDE_RIGHT"), n_frames=6, n_submissions=3,
                                  objects_per_frame=45.0, seed=24),
                         method="nms", min_score=0.01, iou_thresh=0.5, cut=1.0, weights=None),
    "ensemble_fusion": dict(cfg=dict(n_segments=1, cameras=("FRONT", "SIDE_LEFT"), n_frames=6, n_submissions=4,
                                     objects_per_frame=45.0, seed=25),
                            method="weighted_fusion", min_score=0.0, iou_thresh=0.5, cut=1.0, weights=[3, 1, 2, 2]),
    "ensemble_fusion_default": dict(cfg=dict(n_segments=1, cameras=("FRONT_LEFT",), n_frames=5, n_submissions=2,
                                             objects_per_frame=60.0, seed=26),
                                    method="weighted_fusion", min_score=0.05, iou_thresh=0.6, cut=1.0, weights=None),
}


def nms_api_cases():
    """The general signature of box_utils.nms (top_k, conf_thresh, soft / hard) on random box sets,
    ties included; outputs of the reference's own function (torch + torchvision CPU)."""
    import torch
    _, _, bu = ref_shim.load_ensemble()
    rng = np.random.default_rng(77)
    out = {}
    n_cases = 0
    for trial in range(48):
        n = int(rng.integers(1, 70))
        xy = rng.uniform(0, 300, (n, 2))
        boxes = np.c_[xy, xy + rng.uniform(8, 150, (n, 2))]
        scores = np.round(rng.uniform(0.02, 1, n), 1 if trial % 3 == 0 else 5)
        kw = dict(overlap=[0.5, 0.3, 0.7][trial % 3], top_k=[0, 0, 7, 500][trial % 4], soft=bool(trial % 2),
                  conf_thresh=[0.0, 0.25, 0.05][(trial // 2) % 3], soft_nms_cut=[1.0, 0.9][(trial // 4) % 2])
        with ref_shim.stable_torch_sort():
            keep, sc = bu.nms(torch.from_numpy(boxes), torch.from_numpy(scores), **kw)
        p = "c%d_" % n_cases
        out[p + "boxes"], out[p + "scores"] = boxes, scores
        out[p + "args"] = np.asarray([kw["overlap"], kw["top_k"], float(kw["soft"]), kw["conf_thresh"], kw["soft_nms_cut"]])
        out[p + "keep"] = np.asarray([int(k) for k in keep], np.int64)
        out[p + "out_scores"] = np.asarray(sc, np.float64)
        n_cases += 1
    out["n_cases"] = np.int64(n_cases)
    return out


def nms_f32_cases():
    """The detector head's call pattern (detnet/nn/modules/detection.py:59-77): float32 point-form boxes and
    float32 scores through box_utils.nms (soft / hard, top_k, conf_thresh) and box_utils.bbox_vote; outputs of the
    reference's own functions (torch CPU, float32 arithmetic).  Also two float64 bbox_vote cases."""
    import torch
    _, _, bu = ref_shim.load_ensemble()
    rng = np.random.default_rng(91)
    out = {}
    n_cases = 0
    for trial in range(32):
        n = int(rng.integers(2, 90))
        xy = rng.uniform(0, 300, (n, 2))
        dt = np.float32 if trial < 28 else np.float64
        boxes = np.c_[xy, xy + rng.uniform(8, 150, (n, 2))].astype(dt)
        scores = rng.uniform(0.05, 1, n).astype(dt)          # distinct with probability 1
        kw = dict(overlap=[0.5, 0.3, 0.45][trial % 3], top_k=[0, 200, 9][trial % 3], soft=bool(trial % 2),
                  conf_thresh=[0.05, 0.2][(trial // 2) % 2], soft_nms_cut=[1.0, 0.9][(trial // 4) % 2])
        tb, ts = torch.from_numpy(boxes), torch.from_numpy(scores)
        with ref_shim.stable_torch_sort():
            keep, sc = bu.nms(tb, ts, **kw)
        keep_t = torch.as_tensor([int(k) for k in keep], dtype=torch.long)
        voted = bu.bbox_vote(tb[keep_t], sc, tb, ts, 0.6) if len(keep_t) else torch.zeros((0, 4), dtype=tb.dtype)
        p = "c%d_" % n_cases
        out[p + "boxes"], out[p + "scores"] = boxes, scores
        out[p + "args"] = np.asarray([kw["overlap"], kw["top_k"], float(kw["soft"]), kw["conf_thresh"], kw["soft_nms_cut"]])
        out[p + "keep"] = np.asarray([int(k) for k in keep], np.int64)
        out[p + "out_scores"] = sc.numpy().copy()
        out[p + "voted"] = voted.numpy().copy()
        n_cases += 1
    out["n_cases"] = np.int64(n_cases)
    return out


def scene_inputs(scene):
    d = {"image_ids": np.asarray(scene.image_ids()), "n_sub": np.int64(len(scene.submissions))}
    for k, sub in enumerate(scene.submissions):
        d["s%d_img" % k] = sub.image_index
        d["s%d_cat" % k] = sub.category
        d["s%d_bbox" % k] = sub.bbox
        d["s%d_score" % k] = sub.score
    return d


def run_reference_tracking(dets, max_age, min_hits, promotion="nep50"):
    """``promotion="nep50"``: the reference's files as they are under this container's NumPy 2 (keys ``out_*``);
    ``"legacy"``: the shim's emulation of the reference's pinned NumPy 1.x environment (keys ``leg_*``), see
    ``oracle.ref_shim.ref_track_all``."""
    ref_utils, ref_sort, _ = ref_shim.load_tracking()
    with tempfile.NamedTemporaryFile("wt", suffix=".json", delete=False) as fp:
        json.dump(dets, fp)
        path = fp.name
    try:
        predictions = ref_utils.read_data_file(path, SCORE_THR)
    finally:
        os.unlink(path)
    return ref_shim.ref_track_all(predictions, IOU_THR, max_age, min_hits, promotion=promotion)


# Full-size cases (BASELINE.json configs C1 / C2: one whole segment).  Inputs are regenerated from the seeded
# synthetic scene; of the outputs the fixture keeps ids / counts for every row and boxes / scores for the rows of
# every SAMPLE-th image, which keeps the files small.
BIG_SAMPLE = 8
BIG_TRACK = {"big_c1": dict(preset="c1", seed=4101, max_age=2, min_hits=0)}
BIG_ENSEMBLE = {"big_c2": dict(preset="c2", seed=4102, min_score=0.01, iou_thresh=0.5, cut=0.9)}


def big_track_arrays(rows, image_ids, prefix):
    a = golden_io.tracks_to_arrays(rows, image_ids)
    keep = (a["img"] % BIG_SAMPLE) == 0
    return {prefix + "img": a["img"], prefix + "cat": a["cat"].astype(np.int8), prefix + "oid": a["oid"].astype(np.int32),
            prefix + "bbox_s": a["bbox"][keep], prefix + "score_s": a["score"][keep]}


def main_big():
    for name, case in BIG_TRACK.items():
        path = os.path.join(HERE, name + ".npz")
        if os.path.exists(path):
            continue
        scene = synth.make_scene(synth.preset(case["preset"], n_segments=1, seed=case["seed"]))
        dets = synth.to_json_list(scene, scene.submissions[0])
        out = {"preset": np.asarray(case["preset"]), "seed": np.int64(case["seed"]), "sample": np.int64(BIG_SAMPLE),
               "max_age": np.int64(case["max_age"]), "min_hits": np.int64(case["min_hits"])}
        for promotion, prefix in (("nep50", "out_"), ("legacy", "leg_")):
            rows = run_reference_tracking(dets, case["max_age"], case["min_hits"], promotion)
            out.update(big_track_arrays(rows, scene.image_ids(), prefix))
            print(name, promotion, "dets", len(dets), "rows", len(rows))
        np.savez_compressed(path, **out)
    for name, case in BIG_ENSEMBLE.items():
        path = os.path.join(HERE, name + ".npz")
        if os.path.exists(path):
            continue
        scene = synth.make_scene(synth.preset(case["preset"], n_segments=1, seed=case["seed"]))
        subs = [synth.to_json_list(scene, s) for s in scene.submissions]
        rows = ref_shim.ref_ensemble_all(subs, None, case["min_score"], case["iou_thresh"], case["cut"])
        a = golden_io.dets_to_arrays(rows, scene.image_ids())
        keep = (a["img"] % BIG_SAMPLE) == 0
        out = {"preset": np.asarray(case["preset"]), "seed": np.int64(case["seed"]), "sample": np.int64(BIG_SAMPLE),
               "min_score": np.float64(case["min_score"]), "iou_thresh": np.float64(case["iou_thresh"]),
               "cut": np.float64(case["cut"]), "out_img": a["img"], "out_cat": a["cat"].astype(np.int8),
               "out_bbox_s": a["bbox"][keep].astype(np.int16), "out_score_s": a["score"][keep]}
        assert np.array_equal(out["out_bbox_s"], a["bbox"][keep])
        np.savez_compressed(path, **out)
        print(name, "in", sum(len(s) for s in subs), "out", len(rows))


def add_legacy():
    """Adds the ``leg_*`` outputs (legacy promotion) to the tracking fixtures that do not have them yet."""
    for name, case in TRACK_CASES.items():
        path = os.path.join(HERE, name + ".npz")
        g = golden_io.load(name)
        if "leg_oid" in g:
            continue
        scene = synth.make_scene(synth.SynthConfig(**case["cfg"]))
        dets = synth.to_json_list(scene, scene.submissions[0])
        rows = run_reference_tracking(dets, case["max_age"], case["min_hits"], "legacy")
        g.update({"leg_" + k: v for k, v in golden_io.tracks_to_arrays(rows, scene.image_ids()).items()})
        np.savez_compressed(path, **g)
        print(name, "legacy rows", len(rows), "ids equal to nep50:", np.array_equal(g["leg_oid"], g["out_oid"]))
    g = golden_io.load("pipeline_small")
    if "leg_oid" not in g:
        scene = synth.make_scene(synth.SynthConfig(**json.loads(str(g["cfg"]))))
        subs = [synth.to_json_list(scene, s) for s in scene.submissions]
        ens = ref_shim.ref_ensemble_all(subs, None, 0.01, 0.5, 0.9)
        rows = run_reference_tracking(ens, 2, 0, "legacy")
        g.update({"leg_" + k: v for k, v in golden_io.tracks_to_arrays(rows, scene.image_ids()).items()})
        np.savez_compressed(os.path.join(HERE, "pipeline_small.npz"), **g)
        print("pipeline_small legacy rows", len(rows))


def main_methods():
    """Fixtures added after the first batch: written only where the file does not exist yet."""
    for name, case in METHOD_CASES.items():
        path = os.path.join(HERE, name + ".npz")
        if os.path.exists(path):
            continue
        scene = synth.make_scene(synth.SynthConfig(**case["cfg"]))
        subs = [synth.to_json_list(scene, s) for s in scene.submissions]
        rows = ref_shim.ref_ensemble_all(subs, case["weights"], case["min_score"], case["iou_thresh"], case["cut"],
                                         method=case["method"])
        out = scene_inputs(scene)
        out.update({"out_" + k: v for k, v in golden_io.dets_to_arrays(rows, scene.image_ids()).items()})
        out["min_score"], out["iou_thresh"], out["cut"] = (np.float64(case["min_score"]),
                                                           np.float64(case["iou_thresh"]), np.float64(case["cut"]))
        out["weights"] = np.asarray(case["weights"] if case["weights"] else [1] * len(subs), np.float64)
        out["method"] = np.asarray(case["method"])
        out["cfg"] = np.asarray(json.dumps(case["cfg"]))
        np.savez_compressed(path, **out)
        print(name, "in", sum(len(s) for s in subs), "out", len(rows))
    path = os.path.join(HERE, "nms_api.npz")
    if not os.path.exists(path):
        out = nms_api_cases()
        np.savez_compressed(path, **out)
        print("nms_api", int(out["n_cases"]), "cases")
    path = os.path.join(HERE, "nms_f32.npz")
    if not os.path.exists(path):
        out = nms_f32_cases()
        np.savez_compressed(path, **out)
        print("nms_f32", int(out["n_cases"]), "cases")


def main():
    assert ref_shim.available(), "reference not mounted at %s" % ref_shim.REF_ROOT
    if "--new-only" in sys.argv:
        main_methods()
        add_legacy()
        return main_big()
    main_methods()
    main_big()
    for name, case in TRACK_CASES.items():
        scene = synth.make_scene(synth.SynthConfig(**case["cfg"]))
        dets = synth.to_json_list(scene, scene.submissions[0])
        rows = run_reference_tracking(dets, case["max_age"], case["min_hits"])
        out = scene_inputs(scene)
        out.update({"out_" + k: v for k, v in golden_io.tracks_to_arrays(rows, scene.image_ids()).items()})
        out["max_age"], out["min_hits"] = np.int64(case["max_age"]), np.int64(case["min_hits"])
        out["cfg"] = np.asarray(json.dumps(case["cfg"]))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "dets", len(dets), "rows", len(rows), "ids", out["out_oid"].max() if len(rows) else 0)

    for name, case in ENSEMBLE_CASES.items():
        scene = synth.make_scene(synth.SynthConfig(**case["cfg"]))
        if case.get("round_scores"):
            for sub in scene.submissions:
                sub.score[:] = np.maximum(np.round(sub.score, case["round_scores"]), 0.01)
        subs = [synth.to_json_list(scene, s) for s in scene.submissions]
        rows = ref_shim.ref_ensemble_all(subs, case["weights"], case["min_score"], case["iou_thresh"], case["cut"])
        out = scene_inputs(scene)
        out.update({"out_" + k: v for k, v in golden_io.dets_to_arrays(rows, scene.image_ids()).items()})
        out["min_score"], out["iou_thresh"], out["cut"] = (np.float64(case["min_score"]),
                                                           np.float64(case["iou_thresh"]), np.float64(case["cut"]))
        out["weights"] = np.asarray(case["weights"] if case["weights"] else [1] * len(subs), np.float64)
        out["cfg"] = np.asarray(json.dumps(case["cfg"]))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "in", sum(len(s) for s in subs), "out", len(rows))

    # end to end: ensemble JSON -> tracker, both through the reference's own code
    case = ENSEMBLE_CASES["ensemble_c2_small"]
    cfg = dict(case["cfg"], n_frames=24, seed=31)
    scene = synth.make_scene(synth.SynthConfig(**cfg))
    subs = [synth.to_json_list(scene, s) for s in scene.submissions]
    ens = ref_shim.ref_ensemble_all(subs, None, 0.01, 0.5, 0.9)   # images in sorted image_id order
    rows = run_reference_tracking(ens, 2, 0)
    out = scene_inputs(scene)
    out.update({"ens_" + k: v for k, v in golden_io.dets_to_arrays(ens, scene.image_ids()).items()})
    out.update({"out_" + k: v for k, v in golden_io.tracks_to_arrays(rows, scene.image_ids()).items()})
    out["cfg"] = np.asarray(json.dumps(cfg))
    np.savez_compressed(os.path.join(HERE, "pipeline_small.npz"), **out)
    print("pipeline_small ens", len(ens), "rows", len(rows))
    add_legacy()


if __name__ == "__main__":
    main()

"""JSON-to-JSON wall time of both CLIs on a few synthetic segments: this repo's CLIs (native JSON + CUDA)
against the Python JSON path of the reference shape (json.load + dict loops + json.dump) around the same
kernels, and against the oracle port end to end.  usage: python scripts/cli_bench.py [segments]"""
import json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from waymo_2d_tracking_b200 import native_json, packing, synth
from waymo_2d_tracking_b200.detnet import ensemble as ens
from waymo_2d_tracking_b200.tracking import track as track_cli, utils as trk_utils
from waymo_2d_tracking_b200.tracking.sort import sort as sort_mod
import bench

seg = int(sys.argv[1]) if len(sys.argv) > 1 else 4
scene = synth.make_scene(synth.preset("c3", n_segments=seg, seed=5))
with tempfile.TemporaryDirectory() as tmp:
    files = []
    for k, sub in enumerate(scene.submissions):
        p = os.path.join(tmp, "sub%d.json" % k)
        json.dump(synth.to_json_list(scene, sub), open(p, "w"))
        files.append(p)
    gt = os.path.join(tmp, "images.json"); json.dump([], open(gt, "w"))
    mb = sum(os.path.getsize(f) for f in files) / 1e6
    print("%d segments = %d frames, %d input detections, %.0f MB of JSON" % (seg, scene.n_img, sum(len(s.score) for s in scene.submissions), mb))
    # warm-up (CUDA context, library load)
    ens.main(files + ["-o", os.path.join(tmp, "warm.json"), "-m", "soft_nms", "--min-score=0.01", "--soft-nms-cut=0.9"])
    t0 = time.perf_counter()
    ens.main(files + ["-o", os.path.join(tmp, "ens.json"), "-m", "soft_nms", "--min-score=0.01", "--soft-nms-cut=0.9"])
    t1 = time.perf_counter()
    sort_mod.KalmanBoxTracker.count = 0
    track_cli.main(["--ground-truth", gt, "--input", os.path.join(tmp, "ens.json"), "--output", os.path.join(tmp, "trk.json"), "--max-age=2", "--min-hits=0"])
    t2 = time.perf_counter()
    print("native CLIs : ensemble %.2f s, track %.2f s -> %.0f frames/s JSON to JSON" % (t1 - t0, t2 - t1, scene.n_img / (t2 - t0)))
    # the opt-in fused command: no intermediate file
    from waymo_2d_tracking_b200 import pipeline
    for _ in range(2):
        sort_mod.KalmanBoxTracker.count = 0
        t0 = time.perf_counter()
        pipeline.main(files + ["-o", os.path.join(tmp, "fused.json"), "--min-score=0.01", "--soft-nms-cut=0.9", "--max-age=2", "--min-hits=0"])
        t1 = time.perf_counter()
    print("fused CLI   : %.2f s -> %.0f frames/s JSON to JSON; same tracks.json: %s" % (
        t1 - t0, scene.n_img / (t1 - t0), open(os.path.join(tmp, "fused.json"), "rb").read() == open(os.path.join(tmp, "trk.json"), "rb").read()))
    # the same kernels behind Python's json + dict loops (what a pure-Python host layer costs)
    t0 = time.perf_counter()
    subs = [json.load(open(f)) for f in files]
    rows = ens.ensemble_submissions(subs, None, "soft_nms", 0.5, 0.9, 0.01)
    json.dump(rows, open(os.path.join(tmp, "ens_py.json"), "w"))
    t1 = time.perf_counter()
    pred = trk_utils.read_data_file(os.path.join(tmp, "ens_py.json"), bench.SCORE_THR)
    sort_mod.KalmanBoxTracker.count = 0
    out = trk_utils.track_all(pred, bench.IOU_THR, 2, 0)
    json.dump(out, open(os.path.join(tmp, "trk_py.json"), "w"))
    t2 = time.perf_counter()
    print("python JSON : ensemble %.2f s, track %.2f s -> %.0f frames/s JSON to JSON" % (t1 - t0, t2 - t1, scene.n_img / (t2 - t0)))
    same = open(os.path.join(tmp, "ens.json"), "rb").read() == open(os.path.join(tmp, "ens_py.json"), "rb").read() and \
        open(os.path.join(tmp, "trk.json"), "rb").read() == open(os.path.join(tmp, "trk_py.json"), "rb").read()
    print("output files byte-identical between the two host paths:", same)

"""SORT kernel time per category mix at bench size (debug aid): does an SM that holds one category at a time
run faster than the mixed launch?  usage: python scripts/by_class_time.py <segments>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from waymo_2d_tracking_b200 import runtime, synth
import bench
seg = int(sys.argv[1]) if len(sys.argv) > 1 else 150
scene = synth.make_scene(synth.preset("c3", n_segments=seg, n_submissions=1, seed=1000))
sub0 = scene.submissions[0]
for only in (0, 1, 2, 3, 4):
    sub = sub0
    if only:
        keep = sub.category == only
        sub = synth.Submission(sub.image_index[keep], sub.category[keep], sub.bbox[keep], sub.score[keep])
    packed = synth.tracks_from_submission(scene, sub, bench.SCORE_THR)
    ts = []
    for it in range(4):
        runtime.PROFILE = []
        res = runtime.sort_track(packed, bench.IOU_THR, 2, 0, raw=False)
        ts.append(runtime.collect_profile()["sort_track_kernel"])
    runtime.PROFILE = None
    print("category", only, "sort_track_kernel ms", ["%.2f" % t for t in ts], flush=True)

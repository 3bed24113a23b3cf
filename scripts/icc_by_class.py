"""Instruction-cache behaviour of the SORT kernel per category mix (run under ncu; debug aid).
usage: python scripts/icc_by_class.py <segments> <category or 0 for all>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from waymo_2d_tracking_b200 import runtime, synth
import bench
seg, only = int(sys.argv[1]), int(sys.argv[2])
scene = synth.make_scene(synth.preset("c3", n_segments=seg, n_submissions=1, seed=1000))
sub = scene.submissions[0]
if only:
    keep = sub.category == only
    sub = synth.Submission(sub.image_index[keep], sub.category[keep], sub.bbox[keep], sub.score[keep])
packed = synth.tracks_from_submission(scene, sub, bench.SCORE_THR)
for it in range(2):
    res = runtime.sort_track(packed, bench.IOU_THR, 2, 0, raw=False)
print("category", only, "rows", res["n_rows"])

"""Crowded-scene probe (debug aid): SORT of a C4 / C5 scene, kernel time and class flags.  usage: crowd_probe.py <config> [segments]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from waymo_2d_tracking_b200 import runtime, synth
import bench
cfg = sys.argv[1] if len(sys.argv) > 1 else "c4"
seg = int(sys.argv[2]) if len(sys.argv) > 2 else 1
scene = synth.make_scene(synth.preset(cfg, n_segments=seg, seed=1000))
packed = synth.tracks_from_submission(scene, scene.submissions[0], bench.SCORE_THR)
cnt = packed.det_count.reshape(-1, 4)
print("dets per image per class: mean", cnt.mean(0).round(1), "max", cnt.max(0))
dev = runtime.upload_tracks(packed)
for it in range(3):
    runtime.PROFILE = []
    out = runtime.sort_track(packed, bench.IOU_THR, 2, 0, raw=False, dev=dev, to_host=False)
    print(runtime.collect_profile())

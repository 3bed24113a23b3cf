"""A/B timing of libw2t builds (debug aid): SORT kernel time at bench size for each W2T_LIB given.
usage: python scripts/ab_sort.py <segments> <lib.so> [<lib.so> ...]   (each build runs in its own process)"""
import os, subprocess, sys
seg = sys.argv[1]
code = r'''
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from waymo_2d_tracking_b200 import runtime, synth
import bench
scene = synth.make_scene(synth.preset("c3", n_segments=int(sys.argv[1]), seed=1000))
groups = synth.groups_from_scene(scene, None, 0.01)
d_rows = torch.from_numpy(groups.rows).cuda(); d_offs = torch.from_numpy(groups.group_offsets).cuda()
kw = dict(stream_img_offsets=scene.stream_img_offsets, cam_wh=scene.cam_wh(), n_classes=4, score_thr=bench.SCORE_THR,
          iou_thresholds=bench.IOU_THR, max_age=2, min_hits=0, max_group=groups.max_group, **bench.NMS)
ts = []
for it in range(6):
    runtime.PROFILE = []
    out = runtime.ensemble_and_track(d_offs, d_rows, to_host=False, want_ensemble=False, raw=False,
                                     host_group_offsets=groups.group_offsets, **kw)
    ts.append(runtime.collect_profile()["sort_track_kernel"])
print(os.environ.get("W2T_LIB", "default"), "sort_track_kernel ms:", " ".join("%.2f" % t for t in ts),
      "rows", int(out["rows"]["totals"][1].item()), flush=True)
'''
for lib in sys.argv[2:]:
    env = dict(os.environ)
    if lib != "default":
        env["W2T_LIB"] = os.path.abspath(lib)
    subprocess.run([sys.executable, "-c", code, seg], env=env)

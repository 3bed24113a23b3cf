"""SORT kernel time per category mix (debug aid): does an SM that holds one category at a time run faster
than the mixed launch?  Same ensemble -> SORT data as bench.py; with 118 segments every category alone is
exactly one wave of CTAs (590 <= 592 slots).  usage: python scripts/by_class_time.py <segments>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from waymo_2d_tracking_b200 import runtime, synth
import bench
seg = int(sys.argv[1]) if len(sys.argv) > 1 else 118
scene = synth.make_scene(synth.preset("c3", n_segments=seg, seed=1000))
groups = synth.groups_from_scene(scene, None, 0.01)
off = groups.group_offsets.astype(np.int64)
sizes = np.diff(off)
cls_of_group = np.arange(len(sizes)) % 4
kw = dict(stream_img_offsets=scene.stream_img_offsets, cam_wh=scene.cam_wh(), n_classes=4, score_thr=bench.SCORE_THR,
          iou_thresholds=bench.IOU_THR, max_age=2, min_hits=0, **bench.NMS)
for only in (0, 1, 2, 3, 4):
    sz = sizes if only == 0 else np.where(cls_of_group == only - 1, sizes, 0)
    keep = np.repeat(sz > 0, sizes)
    rows = np.ascontiguousarray(groups.rows[keep])
    goff = np.concatenate([[0], np.cumsum(sz)]).astype(np.int32)
    d_rows, d_offs = torch.from_numpy(rows).cuda(), torch.from_numpy(goff).cuda()
    ts = []
    for it in range(4):
        runtime.PROFILE = []
        runtime.ensemble_and_track(d_offs, d_rows, to_host=False, want_ensemble=False, raw=False, host_group_offsets=goff,
                                   max_group=int(sz.max()), **kw)
        ts.append(runtime.collect_profile()["sort_track_kernel"])
    runtime.PROFILE = None
    print("category", only, "boxes", len(rows), "sort_track_kernel ms", " ".join("%.2f" % t for t in ts), flush=True)

"""Host-side cost of one device-resident step (debug aid): wall time of each call vs GPU time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from waymo_2d_tracking_b200 import runtime, synth
import bench
seg = int(sys.argv[1]) if len(sys.argv) > 1 else 150
scene = synth.make_scene(synth.preset("c3", n_segments=seg, seed=1000))
groups = synth.groups_from_scene(scene, None, 0.01)
d_rows = torch.from_numpy(groups.rows).cuda(); d_offs = torch.from_numpy(groups.group_offsets).cuda()
kw = dict(stream_img_offsets=scene.stream_img_offsets, cam_wh=scene.cam_wh(), n_classes=4, score_thr=bench.SCORE_THR,
          iou_thresholds=bench.IOU_THR, max_age=2, min_hits=0, max_group=groups.max_group, **bench.NMS)
import cProfile, pstats
def step():
    return runtime.ensemble_and_track(d_offs, d_rows, to_host=False, want_ensemble=False, raw=False,
                                      host_group_offsets=groups.group_offsets, **kw)
for it in range(3): step()
torch.cuda.synchronize()
hs = []
for it in range(40):
    t0 = time.perf_counter(); out = step(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    hs.append(((t1 - t0) * 1e3, (t2 - t0) * 1e3))
print("host call ms:", " ".join("%.0f" % a for a, b in hs))
print("step ms     :", " ".join("%.0f" % b for a, b in hs))
pr = cProfile.Profile(); pr.enable()
for it in range(5): step()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)

"""Tiny ensemble -> SORT run for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from waymo_2d_tracking_b200 import runtime, synth
import bench
scene = synth.make_scene(synth.SynthConfig(n_segments=1, cameras=("FRONT", "SIDE_LEFT"), n_frames=int(sys.argv[1]) if len(sys.argv) > 1 else 24,
                                           n_submissions=3, objects_per_frame=float(sys.argv[2]) if len(sys.argv) > 2 else 60.0, seed=3))
groups = synth.groups_from_scene(scene, None, 0.01)
if os.environ.get("TIMERS"):
    timers = torch.zeros((scene.n_streams * 4, 16), dtype=torch.int64, device="cuda")
    os.environ["W2T_SORT_TIMERS"] = str(timers.data_ptr())
res = runtime.ensemble_and_track(groups.group_offsets, groups.rows, scene.stream_img_offsets, scene.cam_wh(), 4, 0.5, 0.9, 0.01,
                                 bench.SCORE_THR, bench.IOU_THR, 2, 0, max_group=groups.max_group, raw=False, want_ensemble=False)
print("rows", res["n_rows"], "ids", res["id_next"])

"""e2e time of the pipelined path vs number of chunks (debug aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from waymo_2d_tracking_b200 import runtime, synth, packing
import bench
scene = synth.make_scene(synth.preset("c3", n_segments=150, seed=1000))
groups = synth.groups_from_scene(scene, None, 0.01)
packed = packing.packed_rows(groups.rows)
h_rows = torch.from_numpy(packed.view(np.uint8).reshape(-1, 8)).pin_memory()
h_offs = torch.from_numpy(groups.group_offsets).pin_memory()
kw = dict(stream_img_offsets=scene.stream_img_offsets, cam_wh=scene.cam_wh(), n_classes=4, score_thr=bench.SCORE_THR,
          iou_thresholds=bench.IOU_THR, max_age=2, min_hits=0, max_group=groups.max_group, **bench.NMS)
hoists = [float(x) for x in os.environ.get("HOIST", "1.0").split(",")]
lights = [int(x) for x in os.environ.get("LIGHT", "0").split(",")]
def parse(a):
    return [float(x) for x in a.split(",")] if "," in a else int(a)
for chunks in [parse(c) for c in sys.argv[1:]] or [1, 2, 3, 4, 6, 8]:
  for hoist in hoists:
   for light in lights:
    ts = []
    for it in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        res = runtime.ensemble_and_track_pipelined(h_offs, h_rows, n_chunks=chunks, hoist=abs(hoist), hoist_by_work=hoist < 0, light_first=light, **kw)
        torch.cuda.synchronize(); ts.append(1e3 * (time.perf_counter() - t0))
    print("chunks %s hoist %.2f light_first %d: %s ms  (rows %d)" % (chunks, hoist, light, " ".join("%.1f" % t for t in ts), res["n_rows"]), flush=True)

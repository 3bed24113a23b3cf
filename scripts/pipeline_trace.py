"""Host/device timeline of one pipelined end-to-end step (debug aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from waymo_2d_tracking_b200 import runtime, synth, packing
import bench
seg = int(sys.argv[1]) if len(sys.argv) > 1 else 150
chunks = int(sys.argv[2]) if len(sys.argv) > 2 else 8
hoist = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
scene = synth.make_scene(synth.preset("c3", n_segments=seg, seed=1000))
groups = synth.groups_from_scene(scene, None, 0.01)
h_rows = torch.from_numpy(packing.packed_rows(groups.rows).view(np.uint8).reshape(-1, 8)).pin_memory()
h_offs = torch.from_numpy(groups.group_offsets).pin_memory()
kw = dict(stream_img_offsets=scene.stream_img_offsets, cam_wh=scene.cam_wh(), n_classes=4, score_thr=bench.SCORE_THR,
          iou_thresholds=bench.IOU_THR, max_age=2, min_hits=0, max_group=groups.max_group, **bench.NMS)
for it in range(4):
    runtime.TRACE = [] if it == 3 else None
    torch.cuda.synchronize(); t0 = time.perf_counter()
    start = torch.cuda.Event(enable_timing=True); start.record()
    res = runtime.ensemble_and_track_pipelined(h_offs, h_rows, n_chunks=chunks, hoist=hoist, **kw)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print("it%d total %.1f ms rows %d" % (it, 1e3 * (t1 - t0), res["n_rows"]))
for label, t, ev in runtime.TRACE:
    print("%-18s host %7.2f ms   device %s" % (label, 1e3 * (t - t0), "%7.2f ms" % start.elapsed_time(ev) if ev is not None else "-"))

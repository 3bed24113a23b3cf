"""Throughput of the SORT stage on the crowded configuration C4 (large assignment problems; debug aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from waymo_2d_tracking_b200 import runtime, synth
import bench
seg = int(sys.argv[1]) if len(sys.argv) > 1 else 8
scene = synth.make_scene(synth.preset("c4", n_segments=seg, seed=7))
packed = synth.tracks_from_submission(scene, scene.submissions[0], bench.SCORE_THR)
print("C4: %d streams x 200 frames, dets/frame by class %s" % (scene.n_streams, packed.det_count.reshape(-1, 4).mean(0).round(1)))
for it in range(int(os.environ.get("PROBE_ITERS", "3"))):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = runtime.sort_track(packed, bench.IOU_THR, 2, 0, raw=False)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("it%d %.1f ms -> %.0f frames/s, rows %d, ids %d" % (it, 1e3 * dt, scene.n_img / dt, res["n_rows"], res["id_next"]), flush=True)

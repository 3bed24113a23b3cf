"""Per-frame latency of the stateful API (debug aid): device-resident stepping (w2t_sort_step) vs the
building-block drop-in vs the NumPy port of the reference, one C1-like camera stream."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import sort_port
from waymo_2d_tracking_b200 import synth
from waymo_2d_tracking_b200.tracking.sort import sort as sort_mod
from waymo_2d_tracking_b200.tracking.sort.tracker_sort import DeviceMultiClassTrackerSort, MultiClassTrackerSort
import bench
n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 100
scene = synth.make_scene(synth.preset("c1", cameras=("FRONT",), n_frames=n_frames, seed=3))
pred = sort_port.group_entries(synth.to_json_list(scene, scene.submissions[0]), bench.SCORE_THR)
frames = pred[scene.segments[0]]['FRONT']
rows_of = [[[e['bbox'][0], e['bbox'][1], e['bbox'][0] + e['bbox'][2], e['bbox'][1] + e['bbox'][3], e['score'], e['category_id']]
            for e in frames[fid]] for fid in sorted(frames)]
print("frames", len(rows_of), "detections per frame %.1f" % np.mean([len(r) for r in rows_of]))
for name, make in (("device-resident (w2t_sort_step)", lambda: DeviceMultiClassTrackerSort(2, 0)),
                   ("building blocks (units.cu)", lambda: MultiClassTrackerSort(2, 0)),
                   ("NumPy port of the reference", lambda: sort_port.MultiClassTracker(2, 0))):
    sort_mod.KalmanBoxTracker.count = 0
    sort_port.BoxTracker.count = 0
    trk = make()
    n = 0
    for rows in rows_of[:5]:
        trk.track(rows, bench.IOU_THR)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for rows in rows_of[5:]:
        n += sum(len(v) for v in trk.track(rows, bench.IOU_THR).values())
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("%-34s %8.3f ms per frame  (%d rows)" % (name, 1e3 * dt / (len(rows_of) - 5), n), flush=True)

"""Which sub-streams outgrow the warp kernel (debug aid)?  usage: python scripts/bail_stats.py <segments> [seed] [config]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from waymo_2d_tracking_b200 import runtime, synth
import bench
seg = int(sys.argv[1]) if len(sys.argv) > 1 else 30
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
cfgname = sys.argv[3] if len(sys.argv) > 3 else "c3"
scene = synth.make_scene(synth.preset(cfgname, n_segments=seg, seed=seed))
groups = synth.groups_from_scene(scene, None, 0.01)
for it in range(2):
    runtime.PROFILE = []
    out = runtime.ensemble_and_track(torch.from_numpy(groups.group_offsets).cuda(), torch.from_numpy(groups.rows).cuda(),
                                     scene.stream_img_offsets, scene.cam_wh(), 4, 0.5, 0.9, 0.01, bench.SCORE_THR, bench.IOU_THR, 2, 0,
                                     max_group=groups.max_group, to_host=False, want_ensemble=False, raw=False,
                                     host_group_offsets=groups.group_offsets)
    prof = runtime.collect_profile()
print(prof)
d_plan, ws = out["trk"]["_keepalive"]
sizes = np.diff(groups.group_offsets).astype(np.int32)
exists = (sizes.reshape(-1, 4).sum(1) > 0).astype(np.uint8)
plan = runtime.make_plan(scene.n_streams, 4, scene.stream_img_offsets, sizes, exists, 2)
nq = scene.n_streams * 4
aux = ws[plan["aux_offset"]:plan["aux_offset"] + 64 + 12 * nq].cpu().numpy().view(np.int32)
cls = aux[16:16 + nq]
print("queues: big taken %d small taken %d n_big %d n_small %d" % tuple(aux[:4]), "classes (0 warp, 1 bailed, 2 wide, 3 mid):", np.bincount(cls, minlength=4))
cnt = out["nms"]["trk_count"].cpu().numpy().reshape(-1, 4)
dmax = np.array([[cnt[scene.stream_img_offsets[s]:scene.stream_img_offsets[s + 1], c].max() for c in range(4)] for s in range(scene.n_streams)])
q = np.nonzero(cls)[0]
print("non-warp sub-streams (q, class, category, Dmax):", [(int(i), int(cls[i]), int(i % 4), int(dmax[i // 4, i % 4])) for i in q][:24])
print("tracked dets per image, per class: mean", cnt.mean(0).round(1), "max", cnt.max(0), "p99", np.percentile(cnt, 99, axis=0))

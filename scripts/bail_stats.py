"""Which sub-streams outgrow the warp kernel (debug aid)?  usage: python scripts/bail_stats.py <segments>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from waymo_2d_tracking_b200 import runtime, synth
import bench
seg = int(sys.argv[1]) if len(sys.argv) > 1 else 30
scene = synth.make_scene(synth.preset("c3", n_segments=seg, seed=1000))
groups = synth.groups_from_scene(scene, None, 0.01)
runtime.PROFILE = []
out = runtime.ensemble_and_track(torch.from_numpy(groups.group_offsets).cuda(), torch.from_numpy(groups.rows).cuda(),
                                 scene.stream_img_offsets, scene.cam_wh(), 4, 0.5, 0.9, 0.01, bench.SCORE_THR, bench.IOU_THR, 2, 0,
                                 max_group=groups.max_group, to_host=False, want_ensemble=False, raw=False,
                                 host_group_offsets=groups.group_offsets)
print(runtime.collect_profile())
d_plan, ws = out["trk"]["_keepalive"]
sizes = np.diff(groups.group_offsets).astype(np.int32)
exists = (sizes.reshape(-1, 4).sum(1) > 0).astype(np.uint8)
plan = runtime.make_plan(scene.n_streams, 4, scene.stream_img_offsets, sizes, exists, 2)
nq = scene.n_streams * 4
aux = ws[plan["aux_offset"]:plan["aux_offset"] + 64 + 4 * nq].cpu().numpy().view(np.int32)
bail = aux[16:16 + nq]
print("queue counter", aux[0], "bailed", int(bail.sum()), "of", nq, "narrow_cap", plan["narrow_cap"], "n_mid", plan["n_mid"], "n_wide", plan["n_wide"])
q = np.nonzero(bail)[0]
print("bailed q%4:", np.bincount(q % 4, minlength=4), "det_cap of bailed:", plan["det_cap"][q][:20], "track_cap:", plan["track_cap"][q][:20])
cnt = out["nms"]["trk_count"].cpu().numpy().reshape(-1, 4)
print("tracked dets per image, per class: mean", cnt.mean(0), "max", cnt.max(0), "p99", np.percentile(cnt, 99, axis=0))

"""Per-phase cycle counters of the SORT kernel (debug aid): where does a sub-stream's time go?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from waymo_2d_tracking_b200 import runtime, synth
import bench
seg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
scene = synth.make_scene(synth.preset("c3", n_segments=seg, seed=1000))
groups = synth.groups_from_scene(scene, None, 0.01)
nq = scene.n_streams * 4
timers = torch.zeros((nq, 16), dtype=torch.int64, device="cuda")
os.environ["W2T_SORT_TIMERS"] = str(timers.data_ptr())
runtime.PROFILE = []
res = runtime.ensemble_and_track(groups.group_offsets, groups.rows, scene.stream_img_offsets, scene.cam_wh(), 4, 0.5, 0.9, 0.01,
                                 bench.SCORE_THR, bench.IOU_THR, 2, 0, max_group=groups.max_group, raw=False, want_ensemble=False)
print("kernel ms:", runtime.collect_profile())
t = timers.cpu().numpy().reshape(-1, 4, 16)
names = ["frames", "nan/setup", "stage+cost", "mk.reduce", "mk.greedy", "mk.drive", "mk.step6", "status", "det-part", "phaseB(KF)", "list-part"]
for c, cname in enumerate(["vehicle", "pedestrian", "sign", "cyclist"]):
    tc = t[:, c, :]
    frames = tc[:, 0].mean()
    tot = tc[:, 1:11].sum(1).mean()
    print("%-10s frames %.1f  total %.0f kcyc/substream = %.1f us/frame @1.9GHz; wall %.2f ms/substream (globaltimer) -> %.0f MHz" % (
        cname, frames, tot / 1e3, tot / max(frames, 1) / 1900, tc[:, 14].mean() / 1e6, tot / max(tc[:, 14].mean(), 1) * 1e3))
    print("   " + "  ".join("%s %.1f%%" % (names[i], 100 * tc[:, i].mean() / tot) for i in range(1, 11)))
    print("   per frame: step-6 rounds %.2f  step-4 iterations %.2f  solves %.2f" % (
        tc[:, 11].mean() / max(frames, 1), tc[:, 12].mean() / max(frames, 1), tc[:, 13].mean() / max(frames, 1)))

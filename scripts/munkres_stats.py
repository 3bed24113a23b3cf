"""How much serial work does the reference's Munkres do on a workload?  (debug aid; C oracle counters)"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import c_oracle
from waymo_2d_tracking_b200 import synth
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
kw = dict(n_segments=1)
if name in ("c4", "c5"):
    kw.update(cameras=("FRONT",), n_frames=12)
scene = synth.make_scene(synth.preset(name, seed=1000, **kw))
lib = c_oracle.lib()
lib.w2t_oracle_munkres_stats.argtypes = [C.c_void_p, C.c_int]
if scene.cfg.n_submissions > 1:
    groups = synth.groups_from_scene(scene, None, 0.01)
    nms = c_oracle.softnms_groups(groups.group_offsets, groups.rows, 0.5, 0.9, 0.01, 4, bench.SCORE_THR)
    from waymo_2d_tracking_b200 import packing
    packed = packing.PackedTracks(n_streams=scene.n_streams, n_classes=4, streams=scene.streams(), frame_ids=scene.frame_ids,
        stream_img_offsets=scene.stream_img_offsets, det_start=groups.group_offsets[:-1].copy(), det_count=nms["trk_count"],
        det_box=nms["trk_box"], cam_wh=scene.cam_wh(), img_exists=nms["img_exists"], class_rank=None, n_rows=len(groups.rows))
else:
    packed = synth.tracks_from_submission(scene, scene.submissions[0], bench.SCORE_THR)
lib.w2t_oracle_munkres_stats(None, 1)
c_oracle.sort_track(packed, bench.IOU_THR, 2, 0)
st = np.zeros((8, 6), np.int64)
lib.w2t_oracle_munkres_stats(st.ctypes.data_as(C.c_void_p), 1)
print("frames", scene.n_img, "dets/frame by class", packed.det_count.reshape(-1, 4).mean(0))
print("bucket(m<=)   calls  step4-iters/call  augment/call  step6-rounds/call  n*m/call  greedy-stars/call")
for b, lab in enumerate(["16", "32", "64", "128", "256", "512", ">512"]):
    c = st[b, 0]
    if c:
        print("%-10s %8d %12.1f %14.1f %16.1f %12.0f %12.1f" % (lab, c, st[b, 1] / c, st[b, 2] / c, st[b, 3] / c, st[b, 4] / c, st[b, 5] / c))

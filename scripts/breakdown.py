"""Wall-clock breakdown of one end-to-end step (debug aid; not a benchmark)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from waymo_2d_tracking_b200 import runtime, synth
import bench

seg = int(sys.argv[1]) if len(sys.argv) > 1 else 150
scene = synth.make_scene(synth.preset("c3", n_segments=seg, seed=1000))
groups = synth.groups_from_scene(scene, None, 0.01)
h_rows = torch.from_numpy(groups.rows).pin_memory()
h_offs = torch.from_numpy(groups.group_offsets).pin_memory()
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for it in range(3):
    t0 = T(); d_rows = h_rows.cuda(non_blocking=True); d_offs = h_offs.cuda(non_blocking=True); t1 = T()
    nms = runtime.softnms_groups_device(d_offs, d_rows, len(h_offs) - 1, groups.max_group, 0.5, 0.9, 0.01, 4, bench.SCORE_THR, want_merged=False); t2 = T()
    h_cnt, h_ex = runtime._host(nms["trk_count"]), runtime._host(nms["img_exists"]); t3 = T()
    plan = runtime.make_plan(scene.n_streams, 4, scene.stream_img_offsets, h_cnt.numpy(), h_ex.numpy(), 2); t4 = T()
    d_o = runtime._dev(scene.stream_img_offsets, np.int32, d_rows.device)
    trk = runtime.sort_track_device(scene.n_streams, 4, d_o, d_offs[:-1], nms["trk_count"], nms["trk_box"], nms["img_exists"], runtime._dev(scene.cam_wh(), np.float64, d_rows.device), bench.IOU_THR, 2, 0, plan); t5 = T()
    rows = runtime.finalize_device(scene.n_streams, 4, d_o, d_offs[:-1], trk, None, 0, int(h_cnt.numpy().sum())); t6 = T()
    res = runtime._collect(trk, rows, False); t7 = T()
    print("it%d h2d %.1f nms %.1f cnt_d2h %.1f plan %.1f sort(+alloc) %.1f finalize %.1f d2h %.1f total %.1f ms | ws %.0f MB rows %d" % (
        it, *(1e3 * (b - a) for a, b in [(t0, t1), (t1, t2), (t2, t3), (t3, t4), (t4, t5), (t5, t6), (t6, t7), (t0, t7)]),
        plan["ws_bytes"] / 1e6, res["n_rows"]))
print("plan track_cap max", plan["track_cap"].max(), "det_cap max", plan["det_cap"].max())

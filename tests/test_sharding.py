"""Multi-process tests of the N > 1 path on CPU (gloo, world_size 2): segment sharding, the id
rebase by exclusive scan and the host gather reproduce the single-process result exactly.  The
tracker itself is the oracle port here (no GPU in this container); the GPU variant of the same
check is ``test_sort_is_invariant_to_sharding`` in test_gpu_parity.py."""
import json
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

import helpers
from waymo_2d_tracking_b200 import sharding, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def oracle_track(pred, iou_thr, max_age, min_hits):
    from oracle import sort_port
    rows = sort_port.track_all(pred, iou_thr, max_age, min_hits, reset_ids=True)
    return rows, sort_port.BoxTracker.count


def oracle_merge(subs, weights, method, iou_thresh, cut, min_score):
    from oracle import ensemble_port
    return ensemble_port.ensemble_all(subs, weights, min_score, iou_thresh, cut) if any(len(s) for s in subs) else []


def make_inputs():
    cfg = synth.SynthConfig(n_segments=3, cameras=("FRONT", "SIDE_LEFT"), n_frames=10, n_submissions=2,
                            objects_per_frame=18.0, seed=41)
    scene = synth.make_scene(cfg)
    return scene, [synth.to_json_list(scene, s) for s in scene.submissions]


def oracle_track_packed(packed, iou_thr, max_age, min_hits):
    """C oracle on packed streams -> dense rows like runtime.sort_track(raw=False), ids counted from 1."""
    from oracle import c_oracle
    from waymo_2d_tracking_b200 import packing
    res = c_oracle.sort_track(packed, iou_thr, max_age, min_hits)
    ids, nxt = packing.assign_ids(packed.stream_img_offsets, packed.n_classes, packed.det_start, res["out_count"],
                                  res["created"], res["first_img"], packed.class_rank, res["out_birth"])
    dense = packing.unpack_tracks(packed, res["out_box"], res["out_score"], res["out_count"], res["first_img"], ids)
    index = {iid: i for i, iid in enumerate(packing.image_id_strings(packed))}
    rows = dict(rows_img=np.array([index[r['image_id']] for r in dense], np.int32),
                rows_cat=np.array([r['category_id'] for r in dense], np.int32),
                rows_box=np.array([r['bbox'] for r in dense], np.float64).reshape(-1, 4),
                rows_score=np.array([r['score'] for r in dense], np.float64),
                rows_id=np.array([int(r['object_id']) for r in dense], np.int64), n_rows=len(dense))
    return rows, int(nxt)


def oracle_merge_groups(groups, method, iou_thresh, cut, min_score):
    from oracle import c_oracle
    return c_oracle.softnms_groups(groups.group_offsets, groups.rows, iou_thresh, cut, min_score)


def ensemble_array_worker(rank, world, port, paths, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from waymo_2d_tracking_b200 import native_json
    out = sharding.ensemble_arrays_sharded([native_json.load(p) for p in paths], [1.0, 0.5], "soft_nms", 0.5, 0.9, 0.01,
                                           merge_fn=oracle_merge_groups)
    if rank == 0:
        native_json.write_detections(out_path, *out)
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def array_worker(rank, world, port, json_path, out_path, from_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from waymo_2d_tracking_b200 import native_json
    dets = json_path if from_path else native_json.load(json_path)      # path: the native parse + pack call
    image_ids, rows, nxt = sharding.track_arrays_sharded(dets, helpers.SCORE_THR, helpers.IOU_THR, 2, 0,
                                                         track_fn=oracle_track_packed)
    if rank == 0:
        assert len(rows["segments"]) == len(set(rows["segments"])) > 1
        native_json.write_tracks(out_path, image_ids, rows["rows_img"], rows["rows_box"], rows["rows_score"],
                                 rows["rows_cat"], rows["rows_id"])
    else:
        assert rows is None and image_ids is None
    dist.barrier()
    dist.destroy_process_group()


def worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import sort_port
    scene, subs = make_inputs()
    pred = sort_port.group_entries(subs[0], helpers.SCORE_THR)
    rows, nxt = sharding.track_all_sharded(pred, helpers.IOU_THR, 2, 0, track_fn=oracle_track)
    ens = sharding.ensemble_sharded(subs, [2, 1], "soft_nms", 0.5, 0.9, 0.01, merge_fn=oracle_merge)
    before, total = sharding.exclusive_scan_int(rank + 5)
    assert (before, total) == (sum(r + 5 for r in range(rank)), sum(r + 5 for r in range(world)))
    if rank == 0:
        with open(out_path, "w") as fp:
            json.dump({"rows": [[r['image_id'], r['object_id'], r['category_id'], [float(v) for v in r['bbox']],
                                 float(r['score'])] for r in rows], "next": nxt, "ens": ens}, fp)
    else:
        assert rows is None and ens is None
    dist.barrier()
    dist.destroy_process_group()


def test_block_partition_is_contiguous_and_balanced():
    for n in (0, 1, 5, 150, 151):
        for world in (1, 2, 4, 8):
            blocks = [sharding.block(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_world_size_2_sharded_tracking_and_ensemble_equal_single_process(tmp_path):
    out = tmp_path / "rank0.json"
    mp.spawn(worker, args=(2, free_port(), str(out)), nprocs=2, join=True)
    got = json.loads(out.read_text())
    from oracle import ensemble_port, sort_port
    scene, subs = make_inputs()
    pred = sort_port.group_entries(subs[0], helpers.SCORE_THR)
    want = sort_port.track_all(pred, helpers.IOU_THR, 2, 0, reset_ids=True)
    assert got["next"] == sort_port.BoxTracker.count
    assert len(got["rows"]) == len(want)
    for a, b in zip(got["rows"], want):
        assert a[0] == b['image_id'] and a[1] == b['object_id'] and a[2] == b['category_id']
        assert a[3] == [float(v) for v in b['bbox']] and a[4] == float(b['score'])
    assert got["ens"] == json.loads(json.dumps(ensemble_port.ensemble_all(subs, [2, 1], 0.01, 0.5, 0.9)))


@pytest.mark.timeout(300)
@pytest.mark.parametrize("from_path", [True, False])
def test_world_size_2_array_path_of_the_tracking_cli_equals_single_process(tmp_path, from_path):
    """``sharding.track_arrays_sharded`` (what ``tracking/track.py`` runs under torchrun): native reader -> each rank
    packs and tracks its block of segments -> arrays gathered to rank 0 -> native writer; same file as one process."""
    from oracle import sort_port
    scene, subs = make_inputs()
    src = tmp_path / "sub.json"
    src.write_text(json.dumps(subs[0]))
    out = tmp_path / "tracks.json"
    mp.spawn(array_worker, args=(2, free_port(), str(src), str(out), from_path), nprocs=2, join=True)
    got = json.loads(out.read_text())
    pred = sort_port.group_entries(subs[0], helpers.SCORE_THR)
    want = sort_port.track_all(pred, helpers.IOU_THR, 2, 0, reset_ids=True)
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert a['image_id'] == b['image_id'] and a['object_id'] == b['object_id'] and a['category_id'] == b['category_id']
        assert a['bbox'] == [float(v) for v in b['bbox']]
        assert a['score'] == pytest.approx(float(b['score']), rel=1e-12)


@pytest.mark.timeout(300)
def test_world_size_2_array_path_of_the_ensemble_cli_equals_single_process(tmp_path):
    """``sharding.ensemble_arrays_sharded`` (what ``detnet/ensemble.py`` runs under torchrun) against the oracle port."""
    from oracle import ensemble_port
    scene, subs = make_inputs()
    paths = []
    for k, sub in enumerate(subs):
        p = tmp_path / ("sub%d.json" % k)
        p.write_text(json.dumps(sub))
        paths.append(str(p))
    out = tmp_path / "ens.json"
    mp.spawn(ensemble_array_worker, args=(2, free_port(), paths, str(out)), nprocs=2, join=True)
    got = json.loads(out.read_text())
    want = json.loads(json.dumps(ensemble_port.ensemble_all(subs, [2, 1], 0.01, 0.5, 0.9)))
    assert got == want

#!/usr/bin/env python
"""Benchmark of the post-detection box pipeline (soft-NMS ensemble + SORT) on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port)

Metric (BASELINE.json): tracked frames/sec (ensemble+SORT); a frame is one camera image.
Workload at every N: per GPU, the full test-scale configuration C3 — 150 segments x 5 cameras x
200 frames, 3 submissions (~265 boxes/image in, ~88 tracked detections/frame), soft-NMS
(iou 0.5, cut 0.9, min-score 0.01) then SORT (max-age 2, min-hits 0, README thresholds).
Streams shard by segment with no collective, so per-GPU work is fixed: weak scaling.

A "step" is one pass of the hot path over that batch:
  value : inputs resident in HBM -> soft-NMS kernel -> (counts to host, launch plan) ->
          SORT kernel -> id scan + dense rows, all outputs left in HBM;
  e2e   : the same through the public API with HOST buffers: pinned host arrays -> H2D ->
          the same kernels -> D2H of the dense output rows and ids, every step; the streams are
          cut into --chunks blocks so that copies and kernels of successive blocks overlap.
One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SCORE_THR = [0.95, 0.6, 1.0, 0.9]
IOU_THR = [0.01, 0.01, 1.0, 0.0]
NMS = dict(iou_thresh=0.5, soft_nms_cut=0.9, min_score=0.01)
MAX_AGE, MIN_HITS = 2, 0
METRIC = "tracked frames/sec (ensemble+SORT)"
UNIT = "frames/s"


CONFIGS = {
    # BASELINE.json:configs, in order.  `stages`: what one step runs; `segments`: default size of the job
    "c1": dict(stages="sort", segments=1,
               name="C1 SORT on one segment: %d segment(s) x 5 cameras x 200 frames, ~110 dets/frame in, 4 classes, max-age 2 min-hits 0"),
    "c2": dict(stages="nms", segments=1,
               name="C2 soft-NMS ensemble of 3 submissions (min-score .01, soft-nms-cut .9) on %d segment(s) x 5 cameras x 200 frames"),
    "c3": dict(stages="both", segments=150,
               name="C3 full test-scale per GPU: %d segments x 5 cameras x 200 frames, 3 submissions, "
                    "soft-NMS(iou .5, cut .9, min-score .01) + SORT(max-age 2, min-hits 0)"),
    "c4": dict(stages="sort", segments=1,
               name="C4 crowded-scene stress: %d segment(s) x 5 cameras x 200 frames, ~1000 dets/frame, ~300 live tracks per class (large Hungarian solves)"),
    "c5": dict(stages="both", segments=1,
               name="C5 5-way TTA ensemble (~3000 boxes/frame) + SORT end to end on %d segment(s) x 5 cameras x 200 frames"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default: 20; 3 for --impl reference, whose steps take seconds)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=("ours", "reference"), default="ours")
    ap.add_argument("--config", choices=sorted(CONFIGS), default="c3",
                    help="configuration of BASELINE.json (default c3: the one the metric is quoted on)")
    ap.add_argument("--segments", type=int, default=None, help="segments per GPU (default: the configuration's own size; "
                    "with --scaling strong: segments of the whole job)")
    ap.add_argument("--scaling", choices=("weak", "strong"), default="weak",
                    help="weak: every GPU gets --segments segments; strong: ONE job of --segments segments split over the GPUs")
    ap.add_argument("--seed", type=int, default=1000)
    ap.add_argument("--cpu-sample-frames", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--parity-streams", type=int, default=16,
                    help="streams of the timed output checked against the C oracle after the timed region (0 = off)")
    ap.add_argument("--cli-segments", type=int, default=8,
                    help="segments of the JSON->JSON command-line record (c3, N=1 only; 0 = skip)")
    ap.add_argument("--chunks", type=int, default=8, help="pipeline depth of the end-to-end leg")
    ap.add_argument("--skip-e2e", action="store_true",
                    help="profiling aid: only the device-resident leg (the launch list then shows one step's kernels)")
    ap.add_argument("--wide-rows", action="store_true", help="ship 40-byte float64 input rows instead of the packed ones")
    ap.add_argument("--compact-rows", action="store_true", help="ship 16-byte rows (f64 score + 4 x int16) instead of 8-byte packed ones")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 3 if args.impl == "reference" else 20
    if args.segments is None:
        args.segments = CONFIGS[args.config]["segments"]
    if args.cpu_sample_frames is None:
        args.cpu_sample_frames = {"c1": 200, "c2": 100, "c3": 200, "c4": 12, "c5": 16}[args.config]
    return args


def workload_name(config, segments):
    return CONFIGS[config]["name"] % segments


# ---------------------------------------------------------------------------------------------
# CPU baseline = the oracle port (NumPy restatement of the reference's own Python path)
# ---------------------------------------------------------------------------------------------

def _cpu_stream_job(job):
    """One (segment, camera) stream, `frames` images, through the stages of the configuration."""
    config, seed, stream, frames = job
    from oracle import ensemble_port, sort_port
    from waymo_2d_tracking_b200 import synth
    stages = CONFIGS[config]["stages"]
    scene = synth.make_scene(synth.preset(config, n_segments=1, seed=seed))
    F = scene.cfg.n_frames
    lo, hi = stream * F, stream * F + frames
    ids = scene.image_ids()
    subs = []
    for sub in scene.submissions:
        m = (sub.image_index >= lo) & (sub.image_index < hi)
        subs.append([{'image_id': ids[int(i)], 'category_id': int(c), 'bbox': [int(v) for v in b], 'score': float(s)}
                     for i, c, b, s in zip(sub.image_index[m], sub.category[m], sub.bbox[m], sub.score[m])])
    t0 = time.perf_counter()
    if stages == "sort":
        ens = subs[0]
    else:
        ens = ensemble_port.ensemble_all(subs, None, NMS["min_score"], NMS["iou_thresh"], NMS["soft_nms_cut"])
    n_rows = len(ens)
    if stages != "nms":
        pred = sort_port.group_entries(ens, SCORE_THR)
        n_rows = len(sort_port.track_all(pred, IOU_THR, MAX_AGE, MIN_HITS))
    dt = time.perf_counter() - t0
    return frames, dt, n_rows


PORT_NOTE = ("oracle/ NumPy port of the reference path (ensemble_port + sort_port); the reference's own files, which "
             "cannot travel to the GPU box, measured 3-4x SLOWER than this port in the dev container (10 vs 38 frames/s "
             "per core on the same 400 frames), so ratios against it are conservative")


def cpu_baseline_single(config, seed, frames):
    n, dt, rows = _cpu_stream_job((config, seed, 0, frames))
    return dict(value=n / dt, unit=UNIT, cores=1, kind="port",
                sample="1 stream (segment seed %d, camera FRONT), first %d of 200 frames, single thread as "
                       "tracking/track.py runs it; %.1f s of CPU work; %s" % (seed, frames, dt, PORT_NOTE))


def run_reference_arm(args, rank, world):
    """The reference's CPU implementation of the path on all host cores (oracle port; the reference
    itself is Python that cannot travel to the GPU box and needs two uninstallable dependencies)."""
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    frames = args.cpu_sample_frames
    ctx = mp.get_context("fork")
    per_step = []
    with ctx.Pool(cores) as pool:
        for step in range(args.warmup + args.steps):
            # every worker gets its own stream: camera w % 5 of segment seed + step * ceil(cores / 5) + w // 5
            seg0 = args.seed + step * ((cores + 4) // 5)
            jobs = [(args.config, seg0 + w // 5, w % 5, frames) for w in range(cores)]
            res = pool.map(_cpu_stream_job, jobs)
            # the workers run concurrently: the step takes as long as the slowest one's timed region
            # (synthetic-data generation inside the worker is not part of the path and is excluded)
            if step >= args.warmup:
                per_step.append((sum(r[0] for r in res), max(r[1] for r in res)))
    n = sum(p[0] for p in per_step)
    t = sum(p[1] for p in per_step)
    value = n / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / max(args.steps, 1),
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.config, args.segments),
                   "sample": "each step: %d DISTINCT streams x first %d of 200 frames, one stream per host process" % (cores, frames)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d worker processes, each its own (segment, camera) stream x %d frames per step; %s"
                                   % (cores, frames, PORT_NOTE)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------

class ClockSampler(threading.Thread):
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def _nvml(self):
        """In-process NVML handle (nvidia_ml_py); querying it does not spawn a process or take the
        driver locks `nvidia-smi` takes at start-up, which can stall kernel launches for milliseconds."""
        try:
            import pynvml
            pynvml.nvmlInit()
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index)
        except Exception:
            return None, None

    def run(self):
        nv, h = self._nvml()
        self.source = "nvml (same counters as nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*)" if nv is not None else "nvidia-smi"
        while not self.stop_flag:
            try:
                if nv is not None:
                    sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                    mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                        else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    bits = (0x8, 0x40, 0x20, 0x4)   # hw_slowdown, hw_thermal_slowdown, sw_thermal_slowdown, sw_power_cap
                    self.samples.append([str(sm), str(mx)] + ["Active" if r & b else "Not Active" for b in bits])
                else:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                    parts = [p.strip() for p in out.strip().split(",")]
                    if len(parts) >= 6:
                        self.samples.append(parts)
            except Exception:
                pass
            time.sleep(0.05 if nv is not None else 0.1)

    def summary(self):
        self.stop_flag = True
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(self.samples), "source": getattr(self, "source", "nvidia-smi")}


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------

# ---------------------------------------------------------------------------------------------
# parity sample: streams of the timed output against the C oracle (test infrastructure; the checker only)
# ---------------------------------------------------------------------------------------------

def _rows_of_images(rows, img0, img1):
    """Dense output rows (dict of arrays in the reference's order) restricted to images [img0, img1)."""
    m = (rows["rows_img"] >= img0) & (rows["rows_img"] < img1)
    return {k: np.asarray(rows[k])[m] for k in ("rows_img", "rows_cat", "rows_box", "rows_score", "rows_id")}


def parity_sample(stages, scene, groups, packed, rows, ens, n_streams, seed):
    """Checks `n_streams` randomly chosen streams of the FULL-SIZE output against oracle/ (plain-C restatement):
    kept ensemble rows bit-exact; track images / categories / boxes bit-exact, ids exact relative to the stream's
    first id (the absolute base is the creation count of all earlier streams), confidences within 1e-9."""
    from oracle import c_oracle
    from waymo_2d_tracking_b200 import packing
    NC = 4
    S = scene.n_streams
    rng = np.random.default_rng(seed)
    pick = sorted(rng.choice(S, size=min(n_streams, S), replace=False).tolist())
    offs = scene.stream_img_offsets
    bad, n_rows = 0, 0
    for s_ in pick:
        img0, img1 = int(offs[s_]), int(offs[s_ + 1])
        g0, g1 = img0 * NC, img1 * NC
        if stages == "sort":
            sub = packing.PackedTracks(
                n_streams=1, n_classes=NC, streams=[packed.streams[s_]], frame_ids=packed.frame_ids[img0:img1],
                stream_img_offsets=np.asarray([0, img1 - img0], np.int32),
                det_start=(packed.det_start[g0:g1] - packed.det_start[g0]).astype(np.int32),
                det_count=packed.det_count[g0:g1], cam_wh=packed.cam_wh[s_:s_ + 1],
                det_box=packed.det_box[int(packed.det_start[g0]):int(packed.det_start[g1 - 1] + packed.det_count[g1 - 1])],
                img_exists=None if packed.img_exists is None else packed.img_exists[img0:img1],
                class_rank=None if packed.class_rank is None else packed.class_rank[s_ * NC:(s_ + 1) * NC],
                n_rows=int(packed.det_count[g0:g1].sum()))
        else:
            go = groups.group_offsets
            r0, r1 = int(go[g0]), int(go[g1])
            loc = (go[g0:g1 + 1] - r0).astype(np.int32)
            nms = c_oracle.softnms_groups(loc, groups.rows[r0:r1], NMS["iou_thresh"], NMS["soft_nms_cut"], NMS["min_score"],
                                          NC, SCORE_THR)
            if ens is not None:
                for k in ("ens_count",):
                    bad += int(np.count_nonzero(nms[k] != ens[k][g0:g1]))
                for g in range(g1 - g0):
                    c, o = int(nms["ens_count"][g]), int(loc[g])
                    bad += int(np.count_nonzero(nms["ens_box"][o:o + c] != ens["ens_box"][r0 + o:r0 + o + c]))
                    bad += int(np.count_nonzero(nms["ens_score"][o:o + c] != ens["ens_score"][r0 + o:r0 + o + c]))
                    n_rows += c
            if stages == "nms":
                continue
            sub = packing.PackedTracks(
                n_streams=1, n_classes=NC, streams=[scene.streams()[s_]], frame_ids=scene.frame_ids[img0:img1],
                stream_img_offsets=np.asarray([0, img1 - img0], np.int32), det_start=loc[:-1].copy(),
                det_count=nms["trk_count"], det_box=nms["trk_box"], cam_wh=scene.cam_wh()[s_:s_ + 1],
                img_exists=nms["img_exists"], class_rank=None, n_rows=r1 - r0)
        want = c_oracle.sort_track(sub, IOU_THR, MAX_AGE, MIN_HITS)
        ids, _ = packing.assign_ids(sub.stream_img_offsets, NC, sub.det_start, want["out_count"], want["created"],
                                    want["first_img"], sub.class_rank, want["out_birth"])
        dense = packing.unpack_tracks(sub, want["out_box"], want["out_score"], want["out_count"], want["first_img"], ids)
        got = _rows_of_images(rows, img0, img1)
        n_rows += len(dense)
        if len(dense) != len(got["rows_id"]):
            bad += abs(len(dense) - len(got["rows_id"])) + 1
            continue
        if not dense:
            continue
        index = {iid: i for i, iid in enumerate(packing.image_id_strings(sub))}
        w_img = np.array([index[r["image_id"]] for r in dense]) + img0
        w_cat = np.array([r["category_id"] for r in dense])
        w_box = np.array([r["bbox"] for r in dense], np.float64).reshape(-1, 4)
        w_score = np.array([r["score"] for r in dense], np.float64)
        w_id = np.array([int(r["object_id"]) for r in dense], np.int64)
        bad += int(np.count_nonzero(w_img != got["rows_img"])) + int(np.count_nonzero(w_cat != got["rows_cat"]))
        bad += int(np.count_nonzero(w_box != got["rows_box"]))
        bad += int(np.count_nonzero(~np.isclose(w_score, got["rows_score"], rtol=1e-9, atol=0)))
        bad += int(np.count_nonzero((w_id - w_id.min()) != (got["rows_id"] - got["rows_id"].min())))
    return {"streams": len(pick), "rows": int(n_rows), "mismatches": int(bad), "oracle": "oracle/csrc (plain-C restatement)",
            "checked": "full-size output of the timed configuration: " + (
                "kept ensemble rows bit-exact" if stages == "nms" else
                "track rows (image, category, box bit-exact; ids exact relative to the stream's first id; confidence rtol 1e-9)")}


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------

def cli_record(n_segments, seed):
    """JSON -> JSON through the command lines (README.md:41,54 of the reference): submission files on disk in,
    tracks.json out.  The fused CLI (``python -m waymo_2d_tracking_b200.pipeline``) and the two drop-in CLIs run on
    the same files; the outputs must be byte-identical.  A bounded sample of the C3 workload (file parsing is
    per-segment work, so frames/s does not depend on the sample size beyond launch overheads)."""
    import contextlib
    import tempfile
    from waymo_2d_tracking_b200 import native_json, pipeline, synth
    from waymo_2d_tracking_b200.detnet import ensemble as ens_cli
    from waymo_2d_tracking_b200.tracking import track as track_cli
    from waymo_2d_tracking_b200.tracking.sort import sort as sort_mod
    scene = synth.make_scene(synth.preset("c3", n_segments=n_segments, seed=seed + 77))
    ids = scene.image_ids()
    with tempfile.TemporaryDirectory() as tmp, contextlib.redirect_stdout(sys.stderr):
        files = []
        for k, sub in enumerate(scene.submissions):
            files.append(os.path.join(tmp, "sub%d.json" % k))
            native_json.write_detections(files[-1], ids, sub.image_index, sub.category, sub.bbox, sub.score)
        gt = os.path.join(tmp, "images.json")
        with open(gt, "w") as fp:
            fp.write("[]")
        in_bytes = sum(os.path.getsize(f) for f in files)
        nms = ["--min-score=%r" % NMS["min_score"], "--soft-nms-cut=%r" % NMS["soft_nms_cut"]]
        trk = ["--max-age=%d" % MAX_AGE, "--min-hits=%d" % MIN_HITS]
        fused_out, ens_out, two_out = (os.path.join(tmp, n) for n in ("fused.json", "ens.json", "two.json"))
        best_fused, best_two = float("inf"), float("inf")
        for _ in range(2):                                  # first pass warms the page cache and the pinned pools
            sort_mod.KalmanBoxTracker.count = 0
            t0 = time.perf_counter()
            pipeline.main(files + ["-o", fused_out] + nms + trk)
            best_fused = min(best_fused, time.perf_counter() - t0)
            sort_mod.KalmanBoxTracker.count = 0
            if os.path.exists(ens_out):
                os.remove(ens_out)                          # the reference refuses to overwrite (ensemble.py:134)
            t0 = time.perf_counter()
            ens_cli.main(files + ["-o", ens_out, "-m", "soft_nms"] + nms)
            track_cli.main(["--ground-truth", gt, "--input", ens_out, "--output", two_out] + trk)
            best_two = min(best_two, time.perf_counter() - t0)
        # level (iii) of SURVEY.md §8d: parsed arrays in host memory (what the JSON reader hands over) -> packing on
        # the host (filters, grouping, stream layout, 8-byte rows: NumPy) -> kernels -> dense rows on the host
        from waymo_2d_tracking_b200 import packing, runtime
        dets = [native_json.load(f) for f in files]
        best_arrays = float("inf")
        for _ in range(2):
            t0 = time.perf_counter()
            ids_s, perm, s_off, _frames, go, rows, packed, max_group = pipeline.groups_from_detections(dets, [1.0] * len(dets), NMS["min_score"], 4)
            cams = [ids_s[int(perm[int(o)])].split('/')[2] for o in s_off[:-1]]
            res = runtime.ensemble_and_track_pipelined(
                go.astype(np.int32), packed if packed is not None else rows, stream_img_offsets=s_off,
                cam_wh=np.asarray([packing.camera_size(c) for c in cams], np.float64).reshape(-1, 2), n_classes=4,
                score_thr=SCORE_THR, iou_thresholds=IOU_THR, max_age=MAX_AGE, min_hits=MIN_HITS, max_group=max_group,
                n_chunks=min(8, len(cams)), **NMS)
            best_arrays = min(best_arrays, time.perf_counter() - t0)
        arrays_rows = int(res["n_rows"])
        sort_mod.KalmanBoxTracker.count = 0
        out_bytes = os.path.getsize(fused_out)
        with open(fused_out, "rb") as a, open(two_out, "rb") as b:
            same = a.read() == b.read()
    if not same:
        raise SystemExit("bench.py: the fused CLI and the two-command path wrote different tracks.json files")
    return {"value": scene.n_img / best_fused, "unit": UNIT, "path": "python -m waymo_2d_tracking_b200.pipeline (fused, opt-in)",
            "two_command_value": scene.n_img / best_two,
            "two_command_path": "detnet.ensemble -m soft_nms && tracking/track.py (the reference's commands)",
            "sample": "%d segments = %d frames, %.0f MB of submission JSON in, %.0f MB of tracks JSON out; best of 2 "
                      "in-process runs, files on local disk" % (n_segments, scene.n_img, in_bytes / 1e6, out_bytes / 1e6),
            "json_in_mb_s": in_bytes / 1e6 / best_fused, "outputs_byte_identical": same,
            "arrays_in_rows_out_value": scene.n_img / best_arrays,
            "arrays_in_rows_out_path": "parsed arrays in host memory -> NumPy packing (filters, grouping, stream layout, "
                                       "8-byte rows) -> ensemble_and_track_pipelined -> %d dense rows on the host; the "
                                       "level the CPU arm is timed at (dicts in, dicts out)" % arrays_rows}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    # one process per GPU: keep each rank's host threads within its share of the cores (and, before torch starts
    # its thread pools, pin the process to that share)
    bound = None
    if world > 1:
        from waymo_2d_tracking_b200 import _bind
        bound = _bind.bind_rank_cores(local_rank, int(os.environ.get("LOCAL_WORLD_SIZE", world)))
    share = len(bound) if bound else max(1, (os.cpu_count() or 1) // max(world, 1))
    os.environ.setdefault("W2T_PLAN_THREADS", str(max(1, min(4, share - 1))))
    import torch
    from waymo_2d_tracking_b200 import packing, runtime, synth
    torch.set_num_threads(max(1, min(4, share)))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    stages = CONFIGS[args.config]["stages"]
    # ---- synthetic workload.  weak: every rank its own --segments segments; strong: the job's segments are split
    if args.scaling == "strong":
        lo, hi = args.segments * rank // world, args.segments * (rank + 1) // world
        n_seg = max(hi - lo, 1)
    else:
        n_seg = args.segments
    t0 = time.time()
    scene = synth.make_scene(synth.preset(args.config, n_segments=n_seg, seed=args.seed + rank))
    n_frames = scene.n_img
    groups = packed_trk = None
    if stages == "sort":
        packed_trk = synth.tracks_from_submission(scene, scene.submissions[0], SCORE_THR)
        n_in = int(packed_trk.det_box.shape[0])
    else:
        groups = synth.groups_from_scene(scene, None, NMS["min_score"])
        n_in = int(groups.rows.shape[0])
    gen_s = time.time() - t0

    input_rows = None
    if stages == "sort":
        dev_in = runtime.upload_tracks(packed_trk)
        h2d = int(packed_trk.det_box.nbytes + packed_trk.det_count.nbytes + packed_trk.det_start.nbytes)
        input_rows = "16 B (4 x float32 corner boxes, already filtered by read_data_file)"

        def step_device():
            return runtime.sort_track(packed_trk, IOU_THR, MAX_AGE, MIN_HITS, raw=False, dev=dev_in, to_host=False)

        def step_e2e():
            return runtime.sort_track(packed_trk, IOU_THR, MAX_AGE, MIN_HITS, raw=False)
    else:
        # the synthetic submissions carry integer pixel boxes and 5-decimal scores like real detector output
        # (detnet/data/coco.py:249-250), so 8-byte packed rows hold them exactly (packing.packed_rows verifies that
        # bit for bit and returns None otherwise); --compact-rows / --wide-rows ship 16 B / 40 B rows instead
        packed = None if (args.wide_rows or args.compact_rows) else packing.packed_rows(groups.rows)
        compact = None if (args.wide_rows or packed is not None) else packing.compact_rows(groups.rows)
        if packed is not None:
            h_rows = torch.from_numpy(packed.view(np.uint8).reshape(-1, 8)).pin_memory()
            input_rows = "8 B packed (17-bit score*1e5 + 13/12/11/11-bit box, verified lossless by the packer)"
        elif compact is not None:
            h_rows = torch.from_numpy(compact.view(np.uint8).reshape(-1, 16)).pin_memory()
            input_rows = "16 B compact (f64 score + 4 x int16 box)"
        else:
            h_rows = torch.from_numpy(groups.rows).pin_memory()
            input_rows = "40 B (5 x f64)"
        h_offs = torch.from_numpy(groups.group_offsets).pin_memory()
        d_rows, d_offs = h_rows.cuda(), h_offs.cuda()
        h2d = int(h_rows.numel() * h_rows.element_size() + h_offs.numel() * 4)
        fmt = runtime._host_rows(h_rows)[1]
        G = int(groups.group_offsets.shape[0]) - 1
        if stages == "nms":
            def step_device():
                return {"nms": runtime.softnms_groups_device(d_offs, d_rows, G, groups.max_group, NMS["iou_thresh"],
                                                             NMS["soft_nms_cut"], NMS["min_score"], 4, None,
                                                             want_merged=False, box_format=fmt),
                        "launches": runtime.nms_launch_count(G, groups.max_group)}

            def step_e2e():
                return runtime.softnms_groups(h_offs, h_rows, NMS["iou_thresh"], NMS["soft_nms_cut"], NMS["min_score"], 4,
                                              None, max_group=groups.max_group, want_merged=False)
        else:
            kw = dict(stream_img_offsets=scene.stream_img_offsets, cam_wh=scene.cam_wh(), n_classes=4, score_thr=SCORE_THR,
                      iou_thresholds=IOU_THR, max_age=MAX_AGE, min_hits=MIN_HITS, max_group=groups.max_group, **NMS)

            def step_device():
                # inputs resident in HBM, outputs left in HBM, no host round trip inside the step
                return runtime.ensemble_and_track(d_offs, d_rows, to_host=False, want_ensemble=False, raw=False,
                                                  host_group_offsets=groups.group_offsets, **kw)

            def step_e2e():
                # public API with HOST buffers: chunked so that H2D, kernels and D2H overlap (runtime.py)
                return runtime.ensemble_and_track_pipelined(h_offs, h_rows, n_chunks=min(args.chunks, scene.n_streams), **kw)

    def timed(fn, steps):
        barrier()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        out = None
        host_ms = []
        for _ in range(steps):
            t_host = time.perf_counter()
            out = fn()
            host_ms.append((time.perf_counter() - t_host) * 1e3)
        end.record()
        barrier()
        if os.environ.get("W2T_BENCH_STEP_TIMES"):      # debug aid: host time of every call, per rank
            print("rank %d %s host ms per call: %s" % (rank, fn.__name__, " ".join("%.1f" % v for v in host_ms)),
                  file=sys.stderr, flush=True)
        ms = start.elapsed_time(end)
        if dist is not None:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out

    # warm up exactly the code the timed loop runs (per-kernel CUDA events included), then keep Python's cyclic
    # garbage collector out of the timed regions: a generation-2 pass over the scene's objects is a host stall of
    # tens of milliseconds that has nothing to do with the path being measured
    import gc
    runtime.PROFILE = []
    held = None
    for _ in range(args.warmup):
        # like the timed loop: the previous step's outputs stay alive while the next step allocates its own, so both
        # sets of device buffers exist before the timing starts (a first-time cudaMalloc of ~2 GB inside the timed
        # region stalls the host for 50-80 ms: the "intermittent" stall of earlier bench lines)
        held = step_device()
    del held
    torch.cuda.synchronize()
    gc.collect()
    gc.freeze()
    gc.disable()
    sampler = ClockSampler(local_rank)
    if not os.environ.get("W2T_BENCH_NO_SAMPLER"):      # debug aid
        sampler.start()
    runtime.PROFILE = []
    ms_dev, out = timed(step_device, args.steps)
    kernel_ms = runtime.collect_profile()
    if os.environ.get("W2T_BENCH_STEP_TIMES") and runtime.PROFILE:   # debug aid: device timeline of the timed steps
        first = runtime.PROFILE[0][1]
        print("rank %d device timeline (ms from the first kernel): %s" % (rank, "  ".join(
            "%s@%.1f+%.1f" % (name.split("_")[0], first.elapsed_time(a), a.elapsed_time(b)) for name, a, b in runtime.PROFILE)),
            file=sys.stderr, flush=True)
    runtime.PROFILE = None
    out_e2e = None
    if args.skip_e2e:
        ms_e2e = float("nan")
    else:
        held = None
        for _ in range(max(2, min(args.warmup, 3))):
            held = step_e2e()
        del held
        ms_e2e, out_e2e = timed(step_e2e, args.steps)
    clocks = sampler.summary()
    gc.enable()

    # ---- bookkeeping -------------------------------------------------------------------------------
    K = args.steps
    frames_all = torch.tensor([float(n_frames)], device="cuda")
    if dist is not None:
        dist.all_reduce(frames_all)
    total_frames = int(frames_all.item())
    value = total_frames * K / (ms_dev / 1e3)
    e2e_value = total_frames * K / (ms_e2e / 1e3)
    torch.cuda.synchronize()
    for stage in ("nms", "trk"):
        if stage in out and int(out[stage]["status"].item()) != 0:
            raise SystemExit("bench.py: %s kernel reported status %d" % (stage, int(out[stage]["status"].item())))
    n_trk = n_in if stages == "sort" else (int(out["nms"]["trk_count"].sum().item()) if stages == "both" else 0)
    if stages == "nms":
        n_out = int(out["nms"]["ens_count"].sum().item())
    else:
        n_out = int(out["rows"]["totals"][1].item())
    consistency = None
    if out_e2e is not None and stages != "nms":
        # full-size consistency of the two legs (single launch vs chunked pipeline / host-buffer call): every row's
        # track id, image, category, box and confidence must be identical, bit for bit
        if n_out != int(out_e2e["n_rows"]):
            raise SystemExit("bench.py: device-resident and end-to-end legs disagree on the number of rows")
        dev_rows = out["rows"]
        same = True
        for key in ("rows_id", "rows_img", "rows_cat", "rows_box", "rows_score"):
            same = same and bool(torch.equal(dev_rows[key][:n_out].cpu(), torch.from_numpy(np.asarray(out_e2e[key]))))
        if not same:
            raise SystemExit("bench.py: the end-to-end leg and the device-resident leg produced different rows")
        consistency = "all %d rows of the e2e leg bit-identical to the device-resident leg" % n_out
    # ---- full-size parity sample against the oracle (after the timed region; rank 0's share of the job)
    parity = None
    if rank == 0 and args.parity_streams > 0:
        if stages == "nms":
            ens = {k: out["nms"][k].cpu().numpy() for k in ("ens_count", "ens_box", "ens_score")}
            rows_h = None
        else:
            ens = None
            rows_h = {k: out["rows"][k][:n_out].cpu().numpy() for k in ("rows_img", "rows_cat", "rows_box", "rows_score", "rows_id")}
        parity = parity_sample(stages, scene, groups, packed_trk, rows_h, ens, args.parity_streams, args.seed)
    d2h = 0 if out_e2e is None else int(out_e2e.get("d2h_bytes", 0) if stages != "nms" else
                                        sum(np.asarray(v).nbytes for v in out_e2e.values() if v is not None))
    # ---- what the copies of one end-to-end step cost on their own, all ranks at once: the same byte counts from / to
    # pinned memory on two streams, no kernels.  On a box whose host memory system caps the aggregate DMA rate this
    # floor grows with the number of ranks and bounds the e2e leg from below ("multi-GPU").
    copy_floor_ms = None
    if out_e2e is not None and h2d > 0 and d2h > 0:
        src, dst = torch.empty(h2d, dtype=torch.uint8).pin_memory(), torch.empty(d2h, dtype=torch.uint8).pin_memory()
        d_src, d_dst = torch.empty(h2d, dtype=torch.uint8, device="cuda"), torch.empty(d2h, dtype=torch.uint8, device="cuda")
        s_a, s_b = torch.cuda.Stream(), torch.cuda.Stream()

        def step_copies():
            cur = torch.cuda.current_stream()
            s_a.wait_stream(cur)
            s_b.wait_stream(cur)
            with torch.cuda.stream(s_a):
                d_src.copy_(src, non_blocking=True)
            with torch.cuda.stream(s_b):
                dst.copy_(d_dst, non_blocking=True)
            cur.wait_stream(s_a)
            cur.wait_stream(s_b)
            return None

        for _ in range(2):
            step_copies()
        ms_copy, _ = timed(step_copies, args.steps)
        copy_floor_ms = ms_copy / args.steps
        del src, dst, d_src, d_dst
    # algorithmic bytes: soft-NMS 88 B per input box; SORT 24 B per tracked detection + 60 B per row
    alg = {}
    if stages != "sort":
        alg["softnms_kernel"] = 88.0 * n_in
    if stages != "nms":
        alg["sort_track_kernel"] = 24.0 * n_trk + 60.0 * n_out
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    dominant = max(alg, key=lambda k: kernel_ms.get(k, 0.0)) if kernel_ms else None
    roofline = None
    if dominant:
        traffic, ncu = None, None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath) and args.config == "c3":
            tj = json.load(open(tpath))
            traffic, ncu = tj.get(dominant), tj.get("ncu", {}).get(dominant)
        achieved = alg[dominant] / (kernel_ms[dominant] / 1e3) / 1e9
        roofline = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "kernel_ms": kernel_ms, "algorithmic_bytes": alg, "ncu": ncu,
                    "note": "dependency/latency-bound path (sequential frames per stream, serial Munkres): "
                            "HBM fraction is expected to be small; see DESIGN.md"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
        "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.config, args.segments), "name": args.config, "stages": stages,
                   "frames_per_gpu": n_frames, "frames_total": total_frames,
                   "boxes_in_per_gpu": n_in, "tracked_dets_per_gpu": n_trk, "rows_out_per_gpu": n_out,
                   "input_rows": input_rows,
                   "l2": ("inputs (%.2f GB per step) are larger than the 126 MB L2; no flush needed" % (h2d / 1e9)) if h2d > 126e6
                   else "inputs (%.1f MB per step) fit the 126 MB L2: the step is bound by serial chain latency, not by HBM "
                        "(no flush: every kernel of a step writes its outputs, 2-7x the input size, in between)" % (h2d / 1e6),
                   "parallelism": "streams sharded by segment, %d rank(s), no collective" % world,
                   "host_cores_per_rank": share,
                   "generate_s": round(gen_s, 1), "full_size_check": consistency},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / K,
                "result": None if stages != "both" else
                "dense rows in pinned host memory, 48 B each: 4 x f64 box, f64 confidence, int32 object id - id_base, "
                "int32 image * 8 + category - 1 (runtime.HostRows decodes int64 ids / image / category arrays on first "
                "access, outside the timed step; the JSON writer consumes them)",
                "copies_alone_ms_per_step": copy_floor_ms,
                "copies_alone_note": "the step's host<->device copies by themselves (pinned memory, both directions at "
                                     "once, all ranks together, max over ranks): the floor the host's memory system sets"},
        "gpu_launches": int(out["launches"]) * K,
        "clocks": clocks,
        "roofline": roofline,
        "parity_sample": parity,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_single(args.config, args.seed, args.cpu_sample_frames)
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0 and world == 1 and args.config == "c3" and args.cli_segments > 0:
        line["cli"] = cli_record(args.cli_segments, args.seed)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    if parity is not None and parity["mismatches"] != 0:
        raise SystemExit("bench.py: %d mismatches against the oracle on the parity sample" % parity["mismatches"])


if __name__ == "__main__":
    main()

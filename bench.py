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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default: 20; 3 for --impl reference, whose steps take seconds)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=("ours", "reference"), default="ours")
    ap.add_argument("--segments", type=int, default=150, help="segments per GPU (150 = configuration C3)")
    ap.add_argument("--seed", type=int, default=1000)
    ap.add_argument("--cpu-sample-frames", type=int, default=200)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--chunks", type=int, default=8, help="pipeline depth of the end-to-end leg")
    ap.add_argument("--skip-e2e", action="store_true",
                    help="profiling aid: only the device-resident leg (the launch list then shows one step's kernels)")
    ap.add_argument("--wide-rows", action="store_true", help="ship 40-byte float64 input rows instead of the packed ones")
    ap.add_argument("--compact-rows", action="store_true", help="ship 16-byte rows (f64 score + 4 x int16) instead of 8-byte packed ones")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 3 if args.impl == "reference" else 20
    return args


def workload_name(segments):
    return ("C3 full test-scale per GPU: %d segments x 5 cameras x 200 frames, 3 submissions, "
            "soft-NMS(iou .5, cut .9, min-score .01) + SORT(max-age 2, min-hits 0)" % segments)


# ---------------------------------------------------------------------------------------------
# CPU baseline = the oracle port (NumPy restatement of the reference's own Python path)
# ---------------------------------------------------------------------------------------------

def _cpu_stream_job(job):
    """One (segment, camera) stream, `frames` images: ensemble of the submissions, then SORT."""
    seed, stream, frames = job
    from oracle import ensemble_port, sort_port
    from waymo_2d_tracking_b200 import synth
    scene = synth.make_scene(synth.preset("c3", n_segments=1, seed=seed))
    F = scene.cfg.n_frames
    lo, hi = stream * F, stream * F + frames
    ids = scene.image_ids()
    subs = []
    for sub in scene.submissions:
        m = (sub.image_index >= lo) & (sub.image_index < hi)
        subs.append([{'image_id': ids[int(i)], 'category_id': int(c), 'bbox': [int(v) for v in b], 'score': float(s)}
                     for i, c, b, s in zip(sub.image_index[m], sub.category[m], sub.bbox[m], sub.score[m])])
    t0 = time.perf_counter()
    ens = ensemble_port.ensemble_all(subs, None, NMS["min_score"], NMS["iou_thresh"], NMS["soft_nms_cut"])
    pred = sort_port.group_entries(ens, SCORE_THR)
    rows = sort_port.track_all(pred, IOU_THR, MAX_AGE, MIN_HITS)
    dt = time.perf_counter() - t0
    return frames, dt, len(rows)


def cpu_baseline_single(seed, frames):
    n, dt, rows = _cpu_stream_job((seed, 0, frames))
    return dict(value=n / dt, unit=UNIT, cores=1, kind="port",
                sample="1 stream (segment seed %d, camera FRONT), first %d of 200 frames, oracle/ NumPy port of the "
                       "reference path, single thread as tracking/track.py runs it; %.1f s of CPU work"
                       % (seed, frames, dt))


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
            jobs = [(args.seed + step, w % 5, frames) for w in range(cores)]
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
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.segments),
                   "sample": "each step: %d streams x first %d of 200 frames, one stream per host process" % (cores, frames)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d worker processes, each one (segment, camera) stream x %d frames per step, "
                                   "ensemble_port + sort_port (NumPy restatement of the reference)" % (cores, frames)},
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

def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    from waymo_2d_tracking_b200 import runtime, synth
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

    # ---- synthetic workload (per rank: its own 150 segments) ----------------------------------
    t0 = time.time()
    scene = synth.make_scene(synth.preset("c3", n_segments=args.segments, seed=args.seed + rank))
    groups = synth.groups_from_scene(scene, None, NMS["min_score"])
    gen_s = time.time() - t0
    n_frames = scene.n_img
    # the synthetic submissions carry integer pixel boxes like real detector output (detnet/data/coco.py:250),
    # so the packer emits 16-byte compact rows (double score + 4 x int16) instead of 5 doubles
    from waymo_2d_tracking_b200 import packing
    # ... and 5-decimal scores (coco.py:249), so 8-byte packed rows hold them exactly (packing.packed_rows
    # verifies that bit for bit and returns None otherwise)
    packed = None if (args.wide_rows or args.compact_rows) else packing.packed_rows(groups.rows)
    compact = None if (args.wide_rows or packed is not None) else packing.compact_rows(groups.rows)
    if packed is not None:
        h_rows = torch.from_numpy(packed.view(np.uint8).reshape(-1, 8)).pin_memory()
    elif compact is not None:
        h_rows = torch.from_numpy(compact.view(np.uint8).reshape(-1, 16)).pin_memory()
    else:
        h_rows = torch.from_numpy(groups.rows).pin_memory()
    h_offs = torch.from_numpy(groups.group_offsets).pin_memory()
    d_rows, d_offs = h_rows.cuda(), h_offs.cuda()
    cam_wh = scene.cam_wh()
    kw = dict(stream_img_offsets=scene.stream_img_offsets, cam_wh=cam_wh, n_classes=4, score_thr=SCORE_THR,
              iou_thresholds=IOU_THR, max_age=MAX_AGE, min_hits=MIN_HITS, max_group=groups.max_group, **NMS)

    def step_device():
        # inputs resident in HBM, outputs left in HBM, no host round trip inside the step
        return runtime.ensemble_and_track(d_offs, d_rows, to_host=False, want_ensemble=False, raw=False,
                                          host_group_offsets=groups.group_offsets, **kw)

    def step_e2e():
        # public API with HOST buffers: chunked so that H2D, kernels and D2H overlap (runtime.py)
        return runtime.ensemble_and_track_pipelined(h_offs, h_rows, n_chunks=args.chunks, **kw)

    def timed(fn, steps):
        barrier()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        out = None
        for _ in range(steps):
            out = fn()
        end.record()
        barrier()
        ms = start.elapsed_time(end)
        if dist is not None:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local_rank)
    sampler.start()
    runtime.PROFILE = []
    ms_dev, out = timed(step_device, args.steps)
    kernel_ms = runtime.collect_profile()
    runtime.PROFILE = None
    if args.skip_e2e:
        ms_e2e, out_e2e = float("nan"), {"n_rows": int(out["rows"]["totals"][1].item()), "d2h_bytes": 0}
    else:
        for _ in range(max(1, min(args.warmup, 2))):
            step_e2e()
        ms_e2e, out_e2e = timed(step_e2e, args.steps)
    clocks = sampler.summary()

    # ---- bookkeeping -------------------------------------------------------------------------------
    K = args.steps
    total_frames = n_frames * world
    value = total_frames * K / (ms_dev / 1e3)
    e2e_value = total_frames * K / (ms_e2e / 1e3)
    n_in = int(groups.rows.shape[0])
    torch.cuda.synchronize()
    for stage in ("nms", "trk"):
        if int(out[stage]["status"].item()) != 0:
            raise SystemExit("bench.py: %s kernel reported status %d" % (stage, int(out[stage]["status"].item())))
    n_trk = int(out["nms"]["trk_count"].sum().item())
    n_dev_rows = int(out["rows"]["totals"][1].item())
    if n_dev_rows != int(out_e2e["n_rows"]):
        raise SystemExit("bench.py: device-resident and end-to-end legs disagree on the number of rows")
    parity = None
    if not args.skip_e2e:
        # full-size consistency of the two legs (single launch vs chunked pipeline): every row's track id,
        # image, category and box must be identical, bit for bit
        dev_rows = out["rows"]
        same = True
        for key, host_arr in (("rows_id", out_e2e["rows_id"]), ("rows_img", out_e2e["rows_img"]),
                              ("rows_cat", out_e2e["rows_cat"]), ("rows_box", out_e2e["rows_box"]),
                              ("rows_score", out_e2e["rows_score"])):
            same = same and bool(torch.equal(dev_rows[key][:n_dev_rows].cpu(), torch.from_numpy(np.asarray(host_arr))))
        if not same:
            raise SystemExit("bench.py: the chunked end-to-end leg and the single-launch leg produced different rows")
        parity = "all %d rows of the e2e leg (%d chunks) bit-identical to the single-launch leg" % (n_dev_rows, args.chunks)
    n_out = int(out_e2e["n_rows"])
    h2d = int(h_rows.numel() * h_rows.element_size() + h_offs.numel() * 4)
    d2h = int(out_e2e["d2h_bytes"])
    # algorithmic bytes: soft-NMS 88 B per input box; SORT 24 B per tracked detection + 60 B per row
    alg = {"softnms_kernel": 88.0 * n_in, "sort_track_kernel": 24.0 * n_trk + 60.0 * n_out}
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    dominant = max(alg, key=lambda k: kernel_ms.get(k, 0.0)) if kernel_ms else None
    roofline = None
    if dominant:
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(dominant)
        achieved = alg[dominant] / (kernel_ms[dominant] / 1e3) / 1e9
        roofline = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "kernel_ms": kernel_ms, "algorithmic_bytes": alg,
                    "note": "dependency/latency-bound path (sequential frames per stream, serial Munkres): "
                            "HBM fraction is expected to be small; see DESIGN.md"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
        "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.segments), "frames_per_gpu": n_frames,
                   "boxes_in_per_gpu": n_in, "tracked_dets_per_gpu": n_trk, "track_rows_per_gpu": n_out,
                   "input_rows": ("8 B packed (17-bit score*1e5 + 13/12/11/11-bit box, verified lossless by the packer)" if packed is not None else
                                  "16 B compact (f64 score + 4 x int16 box)" if compact is not None else "40 B (5 x f64)"),
                   "l2": "inputs (%.2f GB per step) are larger than the 126 MB L2; no flush needed" % (h2d / 1e9),
                   "parallelism": "streams sharded by segment, %d rank(s), no collective" % world,
                   "generate_s": round(gen_s, 1), "full_size_check": parity},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / K},
        "gpu_launches": int(out["launches"]) * K,
        "clocks": clocks,
        "roofline": roofline,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_single(args.seed, args.cpu_sample_frames)
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

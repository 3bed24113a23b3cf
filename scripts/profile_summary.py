"""Turn the raw outputs of scripts/gpu_round.sh <tag> (gpurun_out/) into the committed summaries under profiles/."""
import collections, csv, os, shutil, subprocess, sys
tag = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go, pr = os.path.join(root, "gpurun_out"), os.path.join(root, "profiles")
for f in ("bench.json", "bench_ref.json", "launches.csv", "bench_c1.json", "bench_c2.json", "bench_c4.json", "bench_c5.json"):
    if os.path.exists(os.path.join(go, "%s_%s" % (tag, f))):
        shutil.copy(os.path.join(go, "%s_%s" % (tag, f)), os.path.join(pr, "%s_%s" % (tag, f)))
# launch list -> shares
rows = list(csv.DictReader(l for l in open(os.path.join(go, tag + "_launches.csv")) if l.startswith('"')))
acc = collections.OrderedDict()
for r in rows:
    acc.setdefault(r['Kernel Name'][:70], []).append(float(r['Metric Value'].replace(',', '')))
tot = sum(sum(v) for v in acc.values())
with open(os.path.join(pr, tag + "_launch_shares.txt"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none: device time per kernel (ns), shares of the step\n")
    for k, v in sorted(acc.items(), key=lambda kv: -sum(kv[1])):
        f.write("%-72s n=%3d total=%12.0f share=%5.1f%%\n" % (k, len(v), sum(v), 100 * sum(v) / tot))
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct', 'sm__warps_active.avg.pct',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit', 'launch__waves',
        'sm__throughput.avg.pct', 'smsp__inst_executed.sum', 'sm__pipe_fp64_cycles_active.avg', 'smsp__issue_active.avg.pct',
        'launch__shared_mem_per_block', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'smsp__average_warps_issue_stalled', 'sm__icc_request_hit_rate', 'sm__icc_requests.sum', 'gcc__cache_requests_type_instruction.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct']
keys += ['sm__inst_executed_pipe_tmem.avg', 'launch__cluster', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
for k, name in (("sort", "sort_warp_kernel (the SORT stage's dominant kernel)"), ("nms", "softnms_kernel")):
    rep = os.path.join(go, "%s_prof_%s.ncu-rep" % (tag, k))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    h, u, v = r[0], r[1], r[2]
    with open(os.path.join(pr, "%s_ncu_%s_metrics.txt" % (tag, k)), "w") as f:
        f.write("# ncu --set full --clock-control none, one launch of %s, bench.py at its default size (%s)\n" % (name, tag))
        for i, n in enumerate(h):
            if any(w in n for w in keys) and 'peak_sustained.' not in n and 'not_issued' not in n:
                f.write("%s | %s | %s\n" % (n, u[i], v[i]))
# DRAM traffic per launch of each kernel -> profiles/traffic.json (bench.py's roofline.traffic)
import json
traffic = {}
for k, name in (("sort", "sort_track_kernel"), ("nms", "softnms_kernel")):
    vals = {}
    for line in open(os.path.join(pr, "%s_ncu_%s_metrics.txt" % (tag, k))):
        parts = [x.strip() for x in line.split("|")]
        if len(parts) == 3 and parts[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[parts[1]]
            vals[parts[0]] = float(parts[2]) * scale
    if len(vals) == 2:
        traffic[name] = sum(vals.values())
# occupancy / issue figures of the same captures, quoted in the bench line (roofline.ncu)
ncu = {}
pick = {"sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct_of_peak", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slots_busy_pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct_of_peak", "sm__icc_request_hit_rate.pct": "icache_hit_pct",
        "smsp__thread_inst_executed_per_inst_executed.ratio": "active_lanes_per_instruction", "launch__registers_per_thread": "registers_per_thread",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_pct", "gpu__time_duration.sum": "duration_under_ncu"}
for k, name in (("sort", "sort_track_kernel"), ("nms", "softnms_kernel")):
    d = {"capture": "%s, ncu --set full of one full-size launch (numbers under a profiler: shares and rates, not timings)" % tag}
    for line in open(os.path.join(pr, "%s_ncu_%s_metrics.txt" % (tag, k))):
        parts = [x.strip() for x in line.split("|")]
        if len(parts) == 3 and parts[0] in pick:
            d[pick[parts[0]]] = float(parts[2].replace(",", "")) if parts[0] != "gpu__time_duration.sum" else "%s %s" % (parts[2], parts[1])
    ncu[name] = d
traffic["ncu"] = ncu
traffic["source"] = "ncu --set full, one launch, %s (dram__bytes_read.sum + dram__bytes_write.sum)" % tag
json.dump(traffic, open(os.path.join(pr, "traffic.json"), "w"), indent=1)
print("ok", traffic)

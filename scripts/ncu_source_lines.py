"""Per-source-line instruction and stall-sample shares of one kernel from an ncu report captured with
--import-source on.  usage: python scripts/ncu_source_lines.py report.ncu-rep [top N]"""
import csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur_file, agg = None, []
for r in csv.reader(out.splitlines()):
    if len(r) >= 2 and r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
        continue
    if len(r) < 8 or r[0] in ('Line No', 'Function Name', ''):
        continue
    try:
        ln, inst, samp = int(r[0]), int(r[7]), int(r[6])
    except ValueError:
        continue
    agg.append((inst, samp, cur_file, ln, r[1].strip()[:100]))
tot, tots = sum(a[0] for a in agg) or 1, sum(a[1] for a in agg) or 1
print("warp instructions %d, stall samples %d" % (tot, tots))
for a in sorted(agg, reverse=True)[:top]:
    print("%5.1f%% inst %5.1f%% samples  %s:%d  %s" % (100 * a[0] / tot, 100 * a[1] / tots, a[2], a[3], a[4]))
print("-- by stall samples")
for a in sorted(agg, key=lambda a: -a[1])[:12]:
    print("%5.1f%% inst %5.1f%% samples  %s:%d  %s" % (100 * a[0] / tot, 100 * a[1] / tots, a[2], a[3], a[4]))

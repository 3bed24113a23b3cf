"""Join an ncu SASS source page with nvdisasm line info: warp-stall samples per source line / region.

  ncu -i X.ncu-rep --page source --csv > src.csv
  cuobjdump -xelf all waymo_2d_tracking_b200/libw2t.so ; nvdisasm -g -c sort.sm_100a.cubin > sort.sass
  python scripts/ncu_lines.py src.csv sort.sass <kernel mangled-name substring> [top N]
"""
import csv, re, sys, collections

src_csv, sass, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# line info per instruction offset
line_of = {}
cur = None
inside = False
for ln in open(sass):
    if ln.startswith('.text.'):
        inside = kname in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, isamp, iinst = hdr.index('Address'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
base = None
per_line = collections.Counter()
per_line_inst = collections.Counter()
per_line_stall = collections.defaultdict(collections.Counter)
total = 0
for r in rows[2:]:
    if len(r) <= isamp or not r[ia].startswith('0x'):
        continue
    a = int(r[ia], 16)
    if base is None:
        base = a
    off = a - base
    loc = line_of.get(off, ((None, 0), ''))[0]
    s = int(r[isamp] or 0)
    per_line[loc] += s
    per_line_inst[loc] += int(r[iinst] or 0)
    for i, h in stall_cols:
        v = int(r[i] or 0)
        if v:
            per_line_stall[loc][h] += v
    total += s
print("total samples", total)
for loc, s in per_line.most_common(top):
    st = ", ".join("%s %d" % (k.replace('stall_', ''), v) for k, v in per_line_stall[loc].most_common(3))
    print("%-22s %5.1f%%  inst %10d  %s" % ("%s:%d" % loc if loc else "?", 100.0 * s / total, per_line_inst[loc], st))
# per file region summary (function-level buckets by line ranges given on the command line: file:lo-hi=name)

# ---- non-barrier samples (the warps doing work), by source line ---------------------------------
nb = collections.Counter()
for loc, c in per_line_stall.items():
    nb[loc] = sum(v for k, v in c.items() if k != 'stall_barrier')
tnb = sum(nb.values())
print("\nnon-barrier samples", tnb, "(%.1f%% of all)" % (100.0 * tnb / total))
byfile = collections.defaultdict(list)
for loc, s in nb.items():
    if loc and loc[0]:
        byfile[loc[0]].append((loc[1], s))
for f, lst in byfile.items():
    lst.sort()
    print("== %s: %.1f%%" % (f, 100.0 * sum(s for _, s in lst) / tnb))
    # buckets of 10 lines
    b = collections.Counter()
    for l, s in lst:
        b[l // 10 * 10] += s
    for l0 in sorted(b):
        if b[l0] * 200 > tnb:
            print("   lines %4d-%4d  %5.1f%%" % (l0, l0 + 9, 100.0 * b[l0] / tnb))

# ---- warp instructions executed, by region -----------------------------------------------------------
ti = sum(per_line_inst.values())
print("\nwarp instructions executed", ti)
bi = collections.Counter()
for loc, s in per_line_inst.items():
    if loc and loc[0]:
        bi[(loc[0], loc[1] // 10 * 10)] += s
for (f, l0), s in sorted(bi.items()):
    if s * 100 > ti:
        print("   %-22s lines %4d-%4d  %5.1f%%" % (f, l0, l0 + 9, 100.0 * s / ti))

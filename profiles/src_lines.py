#!/usr/bin/env python3
"""profiles/src_lines.py <source.csv> <kernel substring> [n] -- per source line (file:line) share of issued warp instructions, stall
samples and active lanes of one kernel, from `ncu -i rep --page source --csv --print-source sass,cuda`."""
import csv, sys, collections
path, pat = sys.argv[1], sys.argv[2]; ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(path)))
agg = collections.OrderedDict(); fn = None; fpath = None; cur = None; hdr = None; on = False
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": fn = r[1]; on = pat in fn.replace("(int)", "").replace(" ", ""); continue
    if r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; continue
    if not on or hdr is None: continue
    if r[0]:   # a source line header: its own totals
        cur = (fpath, int(r[0]), r[1].strip()[:90])
        a = agg.setdefault(cur, [0.0, 0.0, 0.0])
        for q, h in enumerate(("# Samples", "Instructions Executed", "Thread Instructions Executed")):
            try: a[q] += float(r[hdr[h]] or 0)
            except ValueError: pass
ts = sum(a[0] for a in agg.values()); ti = sum(a[1] for a in agg.values())
print("kernel ~", pat, "| warp instructions %.3e | samples %d" % (ti, ts))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:ntop]:
    print("%5.1f%% inst %5.1f%% samples lanes %4.1f  %s:%d  %s" % (100 * a[1] / ti, 100 * a[0] / ts, a[2] / max(a[1], 1), k[0], k[1], k[2]))

#!/usr/bin/env python3
"""profiles/sass_mix.py <rep> [index] -- opcode mix (stall samples, issued warp instructions, active lanes) of the
longest (or index-th longest) launch in an ncu report captured with --import-source on."""
import subprocess, csv, sys, collections
rep = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
blocks = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name": cur = {"name": r[1], "rows": []}; blocks.append(cur); continue
    if cur is not None: cur["rows"].append(r)
def tot(b):
    hdr = b["rows"][0]; i = hdr.index("# Samples")
    return sum(float(r[i] or 0) for r in b["rows"][1:] if len(r) > i)
blocks.sort(key=lambda b: -tot(b))
b = blocks[which]
hdr = b["rows"][0]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in b["rows"][1:] if len(r) > 10]
f = lambda r, k: float(r[idx[k]] or 0)
ts = sum(f(r, "# Samples") for r in data); ti = sum(f(r, "Instructions Executed") for r in data); tt = sum(f(r, "Thread Instructions Executed") for r in data)
print(b["name"][:80], "| samples %d warp-inst %.3e avg active lanes %.1f" % (ts, ti, tt / ti))
agg = collections.defaultdict(lambda: [0, 0, 0])
for r in data:
    toks = r[idx["Source"]].strip().split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0]
    a = agg[op]; a[0] += f(r, "# Samples"); a[1] += f(r, "Instructions Executed"); a[2] += f(r, "Thread Instructions Executed")
for op, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:16]:
    print("  %-8s samples %5.1f%%  inst %5.1f%%  lanes %4.1f" % (op, 100 * a[0] / ts, 100 * a[1] / ti, a[2] / max(a[1], 1)))
print("  hottest:")
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:12]:
    print("   %4.1f%% lanes %4s  %s" % (100 * f(r, "# Samples") / ts, r[idx["Avg. Threads Executed"]], r[idx["Source"]].strip()[:70]))

#!/usr/bin/env python3
"""profiles/lanes_ab.py <a.csv> <b.csv> -- per-class time, warp instructions and active lanes per instruction of two ncu
launch lists of the same build (metrics gpu__time_duration.sum, smsp__inst_executed.sum,
smsp__thread_inst_executed_per_inst_executed.ratio), e.g. option ket_windows = 1 against 0."""
import csv, re, collections, sys
def load(p):
    rows = [r for r in csv.reader(open(p)) if len(r) > 5 and r[0].isdigit()]
    d = collections.OrderedDict()
    for r in rows:
        d.setdefault(r[0], {'name': r[4]})[r[-3]] = float(r[-1].replace(',', ''))
    return list(d.values())
def key(x):
    m = re.search(r"<(\d), (\d), (\d), (\d)", x['name'].replace("(int)", "")); return ''.join(m.groups())
a = load(sys.argv[1]); b = load(sys.argv[2])
T, I, R = 'gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio'
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0, 0])
for x, y in zip(a, b):
    k = key(x); assert k == key(y)
    g = agg[k]
    g[0] += x[T] / 1e6; g[1] += y[T] / 1e6; g[2] += x[I]; g[3] += y[I]; g[4] += x[I] * x[R]; g[5] += y[I] * y[R]
for k, g in agg.items():
    print("(%s|%s)  ms %.1f vs %.1f | warp instructions %.3e vs %.3e | lanes per instruction %.1f vs %.1f" % (k[:2], k[2:], g[0], g[1], g[2], g[3], g[4] / g[2], g[5] / g[3]))
print("all tile launches: %.1f ms vs %.1f ms (ncu-serialised)" % (sum(g[0] for g in agg.values()), sum(g[1] for g in agg.values())))
for x, y in sorted(zip(a, b), key=lambda t: -t[1][T])[:10]:
    print("  (%s) %.2f ms %.3e inst %.1f lanes | %.2f ms %.3e inst %.1f lanes" % (key(x), x[T] / 1e6, x[I], x[R], y[T] / 1e6, y[I], y[R]))

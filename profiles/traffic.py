#!/usr/bin/env python3
"""profiles/traffic.py <metrics.csv> <workload> <out.json> -- per-class and total DRAM / L2 traffic, red sectors and FP64-pipe
activity of one Fock build from
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sectors_op_red.sum,\
sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio \
      --clock-control none --csv --log-file metrics.csv python profiles/prof_driver.py <workload> 1
(`roofline.traffic` / `l2_bytes` of bench.py read the total from the JSON)."""
import collections, csv, json, re, sys
path, wl, out = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [r for r in csv.reader(open(path)) if len(r) > 5 and r[0].isdigit()]
launch = collections.OrderedDict()
for r in rows:
    d = launch.setdefault(r[0], {"name": r[4]})
    d[r[-3]] = (float(r[-1].replace(",", "")), r[-2])
scale_t = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
scale_b = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
def val(d, k, sc=None):
    if k not in d: return 0.0
    v, u = d[k]
    return v * (sc.get(u, 1.0) if sc else 1.0)
agg = collections.OrderedDict()
for d in launch.values():
    m = re.search(r"(eri_\w+)<(\d), (\d), (\d), (\d)", d["name"])
    if not m: continue
    key = "%s<%s%s|%s%s>" % m.groups()
    a = agg.setdefault(key, collections.Counter())
    ms = val(d, "gpu__time_duration.sum", scale_t)
    a["ms"] += ms; a["n"] += 1
    a["dram_read"] += val(d, "dram__bytes_read.sum", scale_b); a["dram_write"] += val(d, "dram__bytes_write.sum", scale_b)
    a["l2"] += val(d, "lts__t_bytes.sum", scale_b); a["red_sectors"] += val(d, "lts__t_sectors_op_red.sum")
    a["fp64_ms"] += ms * val(d, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")
    a["lanes_ms"] += ms * val(d, "smsp__thread_inst_executed_per_inst_executed.ratio")
res = {"workload": wl, "per_class": {}, "total": {}}
tot = collections.Counter()
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
    if a["ms"] < 0.05: continue
    res["per_class"][k] = {"launches": int(a["n"]), "ms": round(a["ms"], 2), "dram_read_MB": round(a["dram_read"] / 1e6, 1),
                           "dram_write_MB": round(a["dram_write"] / 1e6, 1), "l2_GB": round(a["l2"] / 1e9, 1),
                           "l2_TBps": round(a["l2"] / 1e12 / (a["ms"] * 1e-3), 2), "red_sectors_M": round(a["red_sectors"] / 1e6, 1),
                           "fp64_pipe_pct_mean": round(a["fp64_ms"] / a["ms"], 1), "active_lanes_mean": round(a["lanes_ms"] / a["ms"], 1)}
    for f in ("ms", "dram_read", "dram_write", "l2", "red_sectors", "fp64_ms", "lanes_ms"): tot[f] += a[f]
res["total"] = {"ms_serialised": round(tot["ms"], 1), "dram_bytes": tot["dram_read"] + tot["dram_write"], "l2_bytes": tot["l2"],
                "red_sectors": tot["red_sectors"], "fp64_pipe_pct_mean": round(tot["fp64_ms"] / max(tot["ms"], 1e-9), 1),
                "active_lanes_mean": round(tot["lanes_ms"] / max(tot["ms"], 1e-9), 1)}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res["total"]))

#!/usr/bin/env python3
"""profiles/summarize.py -- turn ncu output brought back in gpurun_out/ into the short summaries kept here.
  launches <launches.csv>          per-kernel totals of gpu__time_duration.sum (shares of the step)
  raw <file.ncu-rep> [regex]       key metrics per captured launch from `ncu --page raw --csv`
"""
import collections, csv, re, subprocess, sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_inst_executed_op_global_red.sum",
        "lts__t_sectors_op_red.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        name, val, unit = r[4], float(r[-1].replace(",", "")), r[-2]
        m = re.search(r"(eri_\w+)<(\d), (\d), (\d), (\d)(?:, (\d))?>", name)
        key = ("%s<%s%s|%s%s>%s" % (m.group(1), m.group(2), m.group(3), m.group(4), m.group(5), (" mode" + m.group(6)) if m.group(6) else "")) if m else name.split("(")[0][:48]
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += val * scale
    tot = sum(v[1] for v in agg.values())
    print("%s: %d launches, total %.3f ms (ncu-serialised, cold cache: compare shares)" % (path, len(rows), tot))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("  %-36s n=%3d %11.3f ms %6.2f%%" % (k, v[0], v[1], 100 * v[1] / tot))


def raw(path, pat=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        if pat and not re.search(pat, name):
            continue
        print("----", name[:90])
        for k in KEYS:
            if k in idx:
                print("   %-72s %16s %s" % (k, r[idx[k]], units[idx[k]]))
        st = sorted(((float(r[idx[c]].replace(",", "") or 0), c) for c in stall), reverse=True)[:7]
        for v, c in st:
            print("   stall/issue %-58s %8.2f" % (c.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        raw(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)

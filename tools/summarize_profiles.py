#!/usr/bin/env python
"""Turn the scratch outputs of a gpurun call (gpurun_out/) into the tracked evidence under profiles/.

    python tools/summarize_profiles.py r01 [workload]

Writes profiles/<tag>_launches_<w>.md   per-kernel totals / shares of the ncu launch list (gpu__time_duration.sum)
       profiles/<tag>_launches_<w>.csv  the launch list itself (kernel, grid, block, ns) -- one line per launch
       profiles/<tag>_hist_full_<w>.md  selected metrics of the `ncu --set full` capture of the dominant kernel
       profiles/<tag>_bench_<w>.json    the bench.py JSON lines of the same call
"""
import csv
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

FULL_KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.sum", "smsp__inst_executed_op_shared_atom.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "smsp__cycles_active.avg",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
    "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio", "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio",
    "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warp_latency_issue_stalled_wait.ratio",
    "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "smsp__average_warp_latency_issue_stalled_not_selected.ratio",
]


def short(name):
    name = re.sub(r"\(.*$", "", name)
    return name.replace("void ", "").replace("gb::", "").strip()


def launches(tag, w):
    src = os.path.join(OUT, "launches_%s.csv" % w)
    if not os.path.exists(src):
        return
    rows = []
    with open(src) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, ig, ib, iv, im = (hdr.index(x) for x in ("Kernel Name", "Grid Size", "Block Size", "Metric Value", "Metric Name"))
    for r in rd:
        if r[im] != "gpu__time_duration.sum":
            continue
        rows.append((short(r[ik]), r[ig], r[ib], float(r[iv].replace(",", ""))))
    with open(os.path.join(PROF, "%s_launches_%s.csv" % (tag, w)), "w") as f:
        f.write("kernel,grid,block,ns\n")
        for k, g, b, ns in rows:
            f.write('%s,"%s","%s",%d\n' % (k, g, b, ns))
    agg = collections.OrderedDict()
    for k, g, b, ns in rows:
        a = agg.setdefault(k, [0, 0.0, 0.0])
        a[0] += 1; a[1] += ns; a[2] = max(a[2], ns)
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(PROF, "%s_launches_%s.md" % (tag, w)), "w") as f:
        f.write("# ncu launch list, workload %s (%s)\n\n" % (w, tag))
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gb:: python bench.py --workload %s --steps 2 --warmup 1 "
                "--no-e2e --no-cpu-baseline`\n\nPer-launch times are cold-cache and serialised (ncu replays every kernel): compare SHARES, "
                "not absolutes.  %d launches, %.3f ms total.\n\n" % (w, len(rows), tot / 1e6))
        f.write("| kernel | launches | total ms | share | avg us | max us |\n|---|---:|---:|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.3f | %.1f%% | %.1f | %.1f |\n" % (k, a[0], a[1] / 1e6, 100 * a[1] / tot, a[1] / a[0] / 1e3, a[2] / 1e3))
    print("launch list:", len(rows), "launches")


def full(tag, w, stem="prof_hist"):
    rep = os.path.join(OUT, "%s_%s.ncu-rep" % (stem, w))
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader([l for l in raw.splitlines() if l.startswith('"')]))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(os.path.join(PROF, "%s_%s_full_%s.md" % (tag, stem.replace("prof_", ""), w)), "w") as f:
        f.write("# `ncu --set full --clock-control none --import-source on` capture, workload %s (%s)\n\n" % (w, tag))
        f.write("Kernel: `%s`; %d launches captured (columns).  Values straight from `ncu -i ... --page raw --csv`.\n\n" % (
            short(data[0][hdr.index("Kernel Name")]), len(data)))
        f.write("| metric | unit | " + " | ".join("launch %d" % i for i in range(len(data))) + " |\n|---|---|" + "---:|" * len(data) + "\n")
        for k in FULL_KEYS:
            if k in hdr:
                i = hdr.index(k)
                f.write("| %s | %s | %s |\n" % (k, units[i], " | ".join(r[i] for r in data)))
        i_r, i_w = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        f.write("\nDRAM traffic per launch (read+write): " + ", ".join(
            "%.1f %s" % (float(r[i_r].replace(",", "")) + float(r[i_w].replace(",", "")), units[i_r]) for r in data) + "\n")
    print("full capture:", len(data), "launches")


def bench(tag, w):
    lines = []
    for name in ("bench_%s.json" % w, "bench_%s_noreplay.json" % w):
        p = os.path.join(OUT, name)
        if os.path.exists(p):
            for l in open(p):
                l = l.strip()
                if l.startswith("{"):
                    json.loads(l)
                    lines.append(l)
    if lines:
        with open(os.path.join(PROF, "%s_bench_%s.json" % (tag, w)), "w") as f:
            f.write("\n".join(lines) + "\n")
        print("bench lines:", len(lines))


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    w = sys.argv[2] if len(sys.argv) > 2 else "c2"
    os.makedirs(PROF, exist_ok=True)
    launches(tag, w)
    full(tag, w)
    bench(tag, w)

"""Turns the scratch ncu outputs under gpurun_out/ into the small tracked summaries under profiles/.

  python profiles/summarize.py <tag> [launches.csv] [report.ncu-rep]

  profiles/<tag>_launches.csv      per-kernel aggregate of the ncu launch list (gpu__time_duration.sum): count, total,
                                   average, share of the captured window  (cold-cache, serialised: compare SHARES)
  profiles/<tag>_launches_raw.csv.gz   the launch list itself
  profiles/<tag>_kernels.csv       one row per captured launch of the `--set full` report: duration, DRAM bytes read /
                                   written, DRAM throughput %, achieved occupancy, registers, L1/L2 hit rates, FP64 pipe %
"""
import collections
import csv
import gzip
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def launches(tag, path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1.0)
        k = row["Kernel Name"]
        a = agg.setdefault(k, [0, 0.0, row["Grid Size"], row["Block Size"]])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(HERE, tag + "_launches.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_us", "avg_us", "share_pct", "grid", "block"])
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            w.writerow([k, a[0], "%.1f" % a[1], "%.2f" % (a[1] / a[0]), "%.2f" % (100 * a[1] / tot), a[2], a[3]])
    with gzip.open(os.path.join(HERE, tag + "_launches_raw.csv.gz"), "wt") as f:
        f.writelines(lines)


WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def kernels(tag, rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    idx = [hdr.index(w) for w in WANT if w in hdr]
    with open(os.path.join(HERE, tag + "_kernels.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([rows[1][i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[i] for i in idx])


if __name__ == "__main__":
    tag = sys.argv[1]
    lp = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "launches.csv")
    rp = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", "prof.ncu-rep")
    if os.path.exists(lp):
        launches(tag, lp)
    if os.path.exists(rp):
        kernels(tag, rp)

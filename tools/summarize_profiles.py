"""Turn gpurun_out/ ncu artefacts into small tracked summaries under profiles/ (gpurun_out/ is scratch).

    python tools/summarize_profiles.py <round-tag> <launches.csv> [name=report.ncu-rep ...]
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__waves_per_multiprocessor", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]


def launches(path, out):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        agg[(name, row["Grid Size"], row["Block Size"])][0] += 1
        agg[(name, row["Grid Size"], row["Block Size"])][1] += float(row["Metric Value"].replace(",", "")) / 1e3
    tot = sum(v[1] for v in agg.values())
    byname = collections.defaultdict(lambda: [0, 0.0])
    for (n, g, b), v in agg.items():
        byname[n][0] += v[0]
        byname[n][1] += v[1]
    with open(out, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised): compare SHARES\n")
        f.write(f"# source: {path}; total {tot / 1e3:.3f} ms over {sum(v[0] for v in agg.values())} launches\n")
        f.write("## by kernel\nkernel,launches,total_us,share_pct\n")
        for n, v in sorted(byname.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{n}\",{v[0]},{v[1]:.1f},{100 * v[1] / tot:.2f}\n")
        f.write("## by kernel and grid\nkernel,grid,block,launches,total_us,avg_us\n")
        for (n, g, b), v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{n}\",\"{g}\",\"{b}\",{v[0]},{v[1]:.1f},{v[1] / v[0]:.1f}\n")


def full(name, rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    with open(out, "a") as f:
        f.write(f"## {name}  ({rep})\n")
        kn = hdr.index("Kernel Name") if "Kernel Name" in hdr else None
        if kn is not None:
            f.write(f"kernel: {vals[kn]}\n")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                f.write(f"{k:85s} {vals[i]:>18s} {units[i]}\n")
        f.write("\n")


if __name__ == "__main__":
    tag = sys.argv[1]
    launches(sys.argv[2], f"profiles/{tag}_launches.csv")
    out = f"profiles/{tag}_ncu_full.txt"
    open(out, "w").write("# ncu --set full --clock-control none --import-source on, one launch each (after 3 warm-ups); tools/ncu_full.sh\n\n")
    for spec in sys.argv[3:]:
        n, rep = spec.split("=")
        full(n, rep, out)

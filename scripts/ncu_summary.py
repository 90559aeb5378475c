"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.

  python scripts/ncu_summary.py launches gpurun_out/X_launches.csv profiles/rNN_launches.txt
  python scripts/ncu_summary.py full     gpurun_out/X.ncu-rep      profiles/rNN_X.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warp_latency_issue_stalled_barrier.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct"]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[row["Metric Unit"]]
        a = agg.setdefault(row["Kernel Name"].split("(")[0][-48:], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised): {src}\n")
        f.write(f"# {sum(a[0] for a in agg.values())} launches, {tot:.3f} ms total\n")
        f.write(f"{'kernel':50s} {'n':>5s} {'total_ms':>10s} {'avg_ms':>9s} {'share':>6s}\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k:50s} {a[0]:5d} {a[1]:10.3f} {a[1] / a[0]:9.4f} {a[1] / tot:6.3f}\n")
    print(open(dst).read())


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none: {src}\n")
        for row in rows[2:]:
            f.write(f"\n## {row[hdr.index('Kernel Name')][:100]}\n")
            for k in KEYS:
                if k in hdr:
                    f.write(f"{k:80s} {row[hdr.index(k)]:>16s} {units[hdr.index(k)]}\n")
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])

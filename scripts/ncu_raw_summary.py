import csv, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "lts__t_sector_hit_rate.pct"]
src, dst, title = sys.argv[1:4]
rows = list(csv.reader(open(src)))
hdr, units = rows[0], rows[1]
seen = {}
with open(dst, "w") as f:
    f.write(f"# {title}\n# ncu --set full --clock-control none, raw page; one block per kernel (its last captured launch)\n")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        seen[d["Kernel Name"].split("(")[0]] = d
    for name, d in seen.items():
        f.write(f"\n## {name}\n")
        for k in KEYS:
            if k in d:
                f.write("%-86s %16s %s\n" % (k, d[k], units[hdr.index(k)]))

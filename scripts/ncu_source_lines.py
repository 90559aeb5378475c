"""Per-source-line stall samples from an ncu report captured with --import-source on:
   python scripts/ncu_source_lines.py report.ncu-rep [top_n]"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
lines = []
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and r and r[0].strip().isdigit():
        d = dict(zip(hdr, r))
        num = lambda k: int(d[k]) if d.get(k, "").strip().isdigit() else 0
        lines.append((int(r[0]), r[1], num("# Samples"), num("Instructions Executed")))
tot = sum(l[2] for l in lines) or 1
print(f"total samples {tot}")
for ln, src, smp, ins in sorted(lines, key=lambda l: -l[2])[:top]:
    print(f"{ln:5d} {smp:8d} {smp / tot:6.3f} {ins:12d}  {src.strip()[:110]}")

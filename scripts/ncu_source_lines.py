"""Per-source-line stall samples from an ncu report captured with --import-source on, per
(kernel, file): python scripts/ncu_source_lines.py report.ncu-rep [top_n]"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(io.StringIO(out)))
num = lambda v: int(v) if v.strip().isdigit() else 0
groups, key, hdr, fpath = {}, None, None, ""
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        key = (r[1].split("(")[0][-40:], fpath); groups.setdefault(key, []); continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr and r[0].strip().isdigit() and key:
        d = dict(zip(hdr, r))
        groups[key].append((int(r[0]), r[1], num(d.get("# Samples", "")), num(d.get("Instructions Executed", ""))))
kern_tot = {}
for (k, f), lines in groups.items():
    kern_tot[k] = kern_tot.get(k, 0) + sum(l[2] for l in lines)
for (k, f), lines in groups.items():
    tot = kern_tot[k] or 1
    sub = sum(l[2] for l in lines)
    if sub < 0.02 * tot:
        continue
    print(f"===== {k} :: {f}  ({sub / tot:.2f} of {tot} samples)")
    for ln, src, smp, ins in sorted(lines, key=lambda l: -l[2])[:top]:
        if smp < 0.004 * tot:
            break
        print(f"{ln:5d} {smp:8d} {smp / tot:6.3f} {ins:12d}  {src.strip()[:110]}")

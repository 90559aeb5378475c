"""SASS instructions (with stall samples) attributed to a source-line range of an ncu report:
   python scripts/ncu_sass_region.py report.ncu-rep LINE_FROM LINE_TO"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
lo, hi = int(sys.argv[2]), int(sys.argv[3])
rows = list(csv.reader(io.StringIO(out)))
hdr = None; cur = None
num = lambda v: int(v) if v.strip().isdigit() else 0
for r in rows:
    if r and r[0] == "Line No":
        hdr = r; continue
    if not hdr or not r:
        continue
    if r[0].strip().isdigit():
        cur = int(r[0]); continue
    if r[0] == "" and len(r) > 6 and cur is not None and lo <= cur <= hi:
        d = dict(zip(hdr, r))
        st = {k[6:]: num(d[k]) for k in hdr if k.startswith("stall_") and "Not" not in k and num(d[k]) > 0}
        top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
        print(cur, f"{num(d['# Samples']):6d} {num(d['Instructions Executed']):9d}", r[3].strip()[:72], top)

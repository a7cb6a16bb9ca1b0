"""Executed-instruction shares of an ncu report, per source line in file order (to read the kernel region by region).
  python tools/ncu_regions.py gpurun_out/prof.ncu-rep [min_pct]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.15
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, last = None, None, None
agg = collections.defaultdict(lambda: [0, 0, ""])
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 4 and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        ie, it = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
        try:
            line = int(r[0]) if r[0] else None
        except ValueError:
            line = None
        if line is not None:
            last = (cur, line)
            agg[last][2] = r[1].strip()[:110]
        if r[ie].isdigit() and r[2] and last:
            agg[last][0] += int(r[ie])
            agg[last][1] += int(r[it])
tot = sum(v[0] for v in agg.values())
print(f"total warp-instructions {tot}")
byfile = collections.defaultdict(int)
for (f, l), v in agg.items():
    byfile[f] += v[0]
for f, c in sorted(byfile.items(), key=lambda x: -x[1]):
    print(f"== {f}: {100 * c / tot:.1f}%")
for (f, l), (c, t, src) in sorted(agg.items()):
    if 100 * c / tot >= min_pct:
        print(f"{100 * c / tot:5.2f}%  lanes={t / max(c, 1):5.1f}  {f}:{l:<4d} {src}")

"""Per-source-line executed-instruction shares from an ncu report (needs -lineinfo and --import-source on).
  python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top_n]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr = None, None
agg = collections.defaultdict(lambda: [0, 0, ""])
kernels = 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) >= 2 and r[0] == "Function Name":
        continue
    if len(r) > 4 and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        ie = hdr.index("Instructions Executed")
        it = hdr.index("Thread Instructions Executed")
        try:
            line = int(r[0]) if r[0] else None
        except ValueError:
            line = None
        if line is not None:
            last = (cur, line)
            agg[last][2] = r[1].strip()[:100]
        if r[ie].isdigit() and r[2]:           # a SASS row under the last source line
            agg[last][0] += int(r[ie])
            agg[last][1] += int(r[it])
tot = sum(v[0] for v in agg.values())
print(f"total warp-instructions {tot}")
for (f, l), (c, t, src) in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print(f"{100 * c / tot:5.1f}%  lanes={t / max(c, 1):5.1f}  {f}:{l:<4d} {src}")

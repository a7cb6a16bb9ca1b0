"""Exclusive per-source-line shares (executed warp instructions, stall samples) of the FIRST launch in an ncu report.
Each SASS instruction is listed once, under the innermost source line it was compiled from.
  python tools/ncu_lines.py rep.ncu-rep [min_pct] [file-substring]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]
min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
only = sys.argv[3] if len(sys.argv) > 3 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, last, seen = None, None, None, set()
skip = False
agg = collections.defaultdict(lambda: [0, 0, 0, ""])
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        skip = cur in seen
        seen.add(cur); continue
    if len(r) > 4 and r[0] == "Line No":
        hdr = r; ie, it, isamp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples"); continue
    if skip or not hdr or len(r) != len(hdr): continue
    if r[0]:
        last = (cur, int(r[0])); agg[last][3] = r[1].strip()[:100]
    elif r[ie].isdigit() and last is not None:
        agg[last][0] += int(r[ie]); agg[last][1] += int(r[it]); agg[last][2] += int(r[isamp] or 0)
tot = sum(v[0] for v in agg.values()); ts = sum(v[2] for v in agg.values())
print(f"warp-instructions {tot}, samples {ts}")
byfile = collections.defaultdict(lambda: [0, 0])
for (f, l), v in agg.items():
    byfile[f][0] += v[0]; byfile[f][1] += v[2]
for f, (c, s) in sorted(byfile.items(), key=lambda x: -x[1][0]):
    print(f"== {f}: inst {100*c/tot:.1f}%  samples {100*s/max(ts,1):.1f}%")
for (f, l), (c, t, s, src) in sorted(agg.items()):
    if only and only not in f: continue
    if 100 * c / tot >= min_pct or 100 * s / max(ts, 1) >= min_pct:
        print(f"inst {100*c/tot:5.2f}%  smp {100*s/max(ts,1):5.2f}%  lanes={t/max(c,1):5.1f}  {f}:{l:<4d} {src}")

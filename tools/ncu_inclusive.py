"""Inclusive executed-instruction / stall-sample shares per source line of ONE file of an ncu report: with inlining a
SASS instruction is listed under every level of its inline chain, so the kernel file's lines give a per-call-site
(per-phase) breakdown.   python tools/ncu_inclusive.py rep.ncu-rep rs_kernels.cu [min_pct]"""
import collections, csv, subprocess, sys
rep, fname = sys.argv[1], sys.argv[2]
min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, last = None, None, None
agg = collections.defaultdict(lambda: [0, 0, 0, ""])
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if len(r) > 4 and r[0] == "Line No":
        hdr = r; continue
    if hdr and len(r) == len(hdr) and cur == fname:
        ie, it, isamp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
        try: line = int(r[0]) if r[0] else None
        except ValueError: line = None
        if line is not None:
            last = line; agg[last][3] = r[1].strip()[:100]
        elif r[ie].isdigit() and last is not None:
            agg[last][0] += int(r[ie]); agg[last][1] += int(r[it]); agg[last][2] += int(r[isamp] or 0)
tot = sum(v[0] for v in agg.values()); ts = sum(v[2] for v in agg.values())
print(f"{fname}: warp-instructions {tot}, samples {ts}")
for l, (c, t, s, src) in sorted(agg.items()):
    if 100 * c / tot >= min_pct or 100 * s / max(ts, 1) >= min_pct:
        print(f"inst {100*c/tot:5.2f}%  samples {100*s/max(ts,1):5.2f}%  lanes={t/max(c,1):5.1f}  {fname}:{l:<4d} {src}")

"""Per-source-line stall samples of one launch in an ncu report, for a chosen stall reason.
  python tools/ncu_stalls.py rep.ncu-rep stall_long_sb [launch-index] [top]"""
import collections, csv, subprocess, sys
rep, reason = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, last = None, None, None
seen = collections.Counter()
agg = collections.defaultdict(lambda: [0, 0, ""])
sass = collections.defaultdict(list)
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        seen[cur] += 1
        continue
    if len(r) > 4 and r[0] == "Line No":
        hdr = r; ic = hdr.index(reason); isamp = hdr.index("# Samples"); continue
    if not hdr or len(r) != len(hdr) or seen[cur] - 1 != which: continue
    if r[0]:
        last = (cur, int(r[0])); agg[last][2] = r[1].strip()[:90]
    elif last is not None:
        v = int(r[ic]) if r[ic].isdigit() else 0
        agg[last][0] += v; agg[last][1] += int(r[isamp]) if r[isamp].isdigit() else 0
        if v: sass[last].append((v, r[3].strip()[:70]))
tot = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print(f"{reason}: {tot} of {ts} samples")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*v[0]/max(tot,1):5.1f}%  {k[0]}:{k[1]:<4d} {v[2]}")
    for c, ins in sorted(sass[k], reverse=True)[:3]:
        print(f"          {c:4d}  {ins}")

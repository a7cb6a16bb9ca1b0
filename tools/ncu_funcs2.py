"""Executed warp instructions / stall samples per device function for ONE launch of an ncu report.
  python tools/ncu_funcs2.py rep.ncu-rep [launch-index] [git-commit-of-the-profiled-tree]"""
import csv, collections, subprocess, sys, re
rep = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0; commit = sys.argv[3] if len(sys.argv) > 3 else "HEAD"
out = subprocess.run(["ncu", "-i", rep, "--launch-skip", str(which), "--launch-count", "1", "--page", "source", "--csv",
                      "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = hdr = last = None
seen = set(); skip = False
dyn = collections.Counter(); thr = collections.Counter(); smp = collections.Counter()
num = lambda x: int(x) if x.isdigit() else 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; skip = cur in seen; seen.add(cur); continue
    if len(r) > 4 and r[0] == "Line No":
        hdr = r; ie = hdr.index("Instructions Executed"); it = hdr.index("Thread Instructions Executed"); isamp = hdr.index("# Samples"); continue
    if skip or not hdr or len(r) != len(hdr): continue
    if r[0]: last = (cur, num(r[0]))
    elif last:
        dyn[last] += num(r[ie]); smp[last] += num(r[isamp]); thr[last] += num(r[it])
regs = {}
for f in {k[0] for k in dyn}:
    src = subprocess.run(["git", "show", f"{commit}:radiation_ppo_b200/csrc/{f}"], capture_output=True, text=True).stdout.split("\n")
    starts = []
    for i, l in enumerate(src, 1):
        m = re.match(r"^.*?\b([A-Za-z_0-9]+)\s*\(", l)
        if l.startswith(("__device__", "__global__")) and m: starts.append((i, m.group(1)))
    regs[f] = starts
def region(f, l):
    name = f
    for i, n in regs.get(f, []):
        if i <= l: name = n
        else: break
    return f.split('.')[0][3:] + ":" + name
D = collections.Counter(); T = collections.Counter(); M = collections.Counter()
for k in dyn:
    g = region(*k); D[g] += dyn[k]; T[g] += thr[k]; M[g] += smp[k]
td = sum(D.values()); tm = sum(M.values())
print(f"launch {which}: warp-instructions {td} ({td/4096:.0f} per 32-env tile of 131072 envs), samples {tm}")
for g, _ in D.most_common(45):
    if D[g] * 1000 < td and M[g] * 200 < tm: continue
    print(f"{g:36s} inst {100*D[g]/td:5.1f}% ({D[g]/4096:6.0f}/tile)  samples {100*M[g]/max(tm,1):5.1f}%  lanes {T[g]/max(D[g],1):5.1f}")

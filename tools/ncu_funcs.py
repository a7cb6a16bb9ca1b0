"""Static SASS size, executed warp instructions and stall samples of the first launch in an ncu report, grouped by
source-line ranges (REGIONS below = functions of the step kernel).  python tools/ncu_funcs.py rep.ncu-rep"""
import csv, collections, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur=None; seen=set(); skip=False; last=None
static=collections.Counter(); dyn=collections.Counter(); smp=collections.Counter(); thr=collections.Counter()
hdr=None
def num(x):
    try: return int(x)
    except ValueError: return 0
for r in rows:
    if len(r)>=2 and r[0]=="File Path":
        cur=r[1].split("/")[-1]; skip=cur in seen; seen.add(cur); continue
    if len(r)>4 and r[0]=="Line No":
        hdr=r; ie=hdr.index("Instructions Executed"); it=hdr.index("Thread Instructions Executed"); isamp=hdr.index("# Samples"); continue
    if skip or not hdr or len(r)!=len(hdr): continue
    if r[0]: last=(cur,num(r[0]))
    elif last:
        static[last]+=1; dyn[last]+=num(r[ie]); smp[last]+=num(r[isamp]); thr[last]+=num(r[it])
import re
def load_regions():
    """function name -> line range, parsed from '// @region name' ... markers is overkill: use the def lines"""
    regs = {}
    import os
    base = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "radiation_ppo_b200", "csrc")
    for f in os.listdir(base):
        lines = open(os.path.join(base, f), errors="replace").read().split("\n")
        starts = []
        for i, l in enumerate(lines, 1):
            m = re.match(r"^(?:template.*\n)?\s*(?:__device__|__global__|__host__|static|inline|int |void |size_t |bool ).*?([A-Za-z_0-9]+)\s*\(", l)
            if m and not l.startswith(" " * 8) and ("{" in l or l.rstrip().endswith(",") or l.rstrip().endswith("(")) and not l.strip().startswith("//"):
                if l.startswith("    ") and not ("__device__" in l or "__global__" in l): continue
                starts.append((i, m.group(1)))
        regs[f] = starts
    return regs
REGS = load_regions()
def region(f, l):
    st = REGS.get(f)
    if not st: return f
    name = f
    for i, n in st:
        if i <= l: name = n
        else: break
    return f"{f.split('.')[0][3:]}:{name}"
S=collections.Counter(); D=collections.Counter(); M=collections.Counter(); T=collections.Counter()
for k in static:
    g=region(*k); S[g]+=static[k]; D[g]+=dyn[k]; M[g]+=smp[k]; T[g]+=thr[k]
td=sum(D.values()); tm=sum(M.values())
print(f"static SASS {sum(S.values())}  executed warp-inst {td}  samples {tm}")
for g,_ in D.most_common():
    if D[g] == 0 and M[g] == 0: continue
    print(f"{g:34s} static {S[g]:6d}  inst {100*D[g]/td:5.1f}%  samples {100*M[g]/max(tm,1):5.1f}%  lanes {T[g]/max(D[g],1):5.1f}")

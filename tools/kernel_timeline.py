"""Kernel timeline of the headline loop (CUPTI through torch.profiler): which kernels run when, on which stream, for how long --
to see what the resets cost besides the step kernel.   python tools/kernel_timeline.py [--episode-steps 120] [--ring 4]"""
import argparse, collections, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--episode-steps", type=int, default=120)
ap.add_argument("--ring", type=int, default=4)
ap.add_argument("--blocks", type=int, default=16)
ap.add_argument("--period", type=int, default=0, help="steps per prefetch block (0: the library's)")
ap.add_argument("--rows", type=int, default=60)
a = ap.parse_args()

class A: pass
args = A(); args.__dict__.update(gpus=1)
cx = bench.Ctx(args)
envs = bench.make_ring(cx, 131072, a.ring, True, a.episode_steps, period=a.period)
R, P = a.ring, envs[0].PREFETCH_PERIOD
g = torch.Generator(device=cx.dev).manual_seed(7)
acts = torch.randint(0, 8, (8, P, 131072, 1), generator=g, device=cx.dev, dtype=torch.int32)
main = torch.cuda.current_stream(cx.dev)
streams = [torch.cuda.Stream(device=cx.dev) for _ in range(R)]
def run(nb, i0=0):
    for st in streams: st.wait_stream(main)
    for i in range(i0, i0 + nb):
        with torch.cuda.stream(streams[i % R]):
            envs[i % R].step_block(acts[i % 8])
    for st in streams: main.wait_stream(st)
run(40 * R)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    run(a.blocks, 40 * R)
    torch.cuda.synchronize()
kev = [e for e in prof.profiler.kineto_results.events() if "CUDA" in str(e.device_type()) and e.duration_ns() > 0]
kev.sort(key=lambda e: e.start_ns())
t0 = kev[0].start_ns()
tot = collections.defaultdict(lambda: [0, 0.0])
rows = []
import re
sid = {}
for e in kev:
    m = re.search(r"(step1_kernel|step_kernel|reset_kernel|maps_\w+|gae_\w+|Memcpy \w+|Memset)", e.name())
    name = m.group(1) if m else e.name()[:28]
    b, d = (e.start_ns() - t0) / 1e3, e.duration_ns() / 1e3
    if name == "reset_kernel":
        name = "prepare" if d > 50 else "reset(stragglers)"
    st = sid.setdefault(e.device_resource_id(), len(sid))
    tot[name][0] += 1; tot[name][1] += d
    rows.append((b, b + d, name, st))
span = max(r[1] for r in rows)
print(f"span {span:.1f} us for {a.blocks * P} steps = {span / (a.blocks * P):.2f} us/step  (period {P}, ring {R})")
for k, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:30s} n={c:4d}  total {t:9.1f} us  mean {t / c:7.2f} us")
print(f"first {a.rows} kernels: start, end, stream, name")
for r in rows[:a.rows]:
    print(f"{r[0]:9.1f} {r[1]:9.1f}  s{r[3]}  {r[2]}")

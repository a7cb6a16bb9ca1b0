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
a = ap.parse_args()

class A: pass
args = A(); args.__dict__.update(gpus=1)
cx = bench.Ctx(args)
envs = bench.make_ring(cx, 131072, a.ring, True, a.episode_steps)
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
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
tot = collections.defaultdict(lambda: [0, 0.0])
rows = []
for e in ev:
    import re
    m = re.search(r"(step1_kernel|step_kernel|reset_kernel|maps_\w+|gae_\w+|Memcpy \w+|Memset)", e.name)
    name = m.group(1) if m else e.name[:28]
    if name == "reset_kernel": name += f" grid{getattr(e, 'grid', '')}" 
    tot[name][0] += 1; tot[name][1] += e.time_range.end - e.time_range.start
    rows.append((e.time_range.start - t0, e.time_range.end - t0, name))
span = max(r[1] for r in rows)
print(f"span {span:.1f} us for {a.blocks * P} steps = {span / (a.blocks * P):.2f} us/step")
for k, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:30s} n={c:4d}  total {t:9.1f} us  mean {t / c:7.2f} us")
print("first 60 kernels: start, end, name")
for r in rows[:60]:
    print(f"{r[0]:9.1f} {r[1]:9.1f}  {r[2]}")

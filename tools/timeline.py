"""Where does a step's time go?  CPU submission time vs GPU span, and a CUPTI kernel timeline via torch.profiler."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import radiation_ppo_b200 as rp

dev = torch.device("cuda:0")
N = 131072
mode = sys.argv[1] if len(sys.argv) > 1 else "graph"
envs = []
for r in range(4):
    e = rp.RadSearch(obstruction_count=5, enforce_grid_boundaries=True, num_envs=N, seed=2, env_id_offset=r * N, auto_reset=True,
                     fast_poisson=True, prefetch=mode != "plain", use_cuda_graph=mode == "graph")
    g = torch.Generator(device=dev).manual_seed(r)
    e._meta.add_(torch.randint(0, 120, (N,), generator=g, device=dev, dtype=torch.int32) << 16)
    torch.cuda.synchronize(); e.capture_graphs(); envs.append(e)
acts = torch.randint(0, 8, (16, N, 1), device=dev, dtype=torch.int32)
def run(k0, K):
    for i in range(k0, k0 + K):
        envs[i % 4].step_batch(acts[i % 16])
run(0, 24); torch.cuda.synchronize()
t0 = time.perf_counter(); run(24, 240); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"mode={mode} cpu submit {1e6*(t1-t0)/240:.1f} us/step, total {1e6*(t2-t0)/240:.1f} us/step")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    run(264, 36); torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
import collections
agg = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    agg[e.name[:60]][0] += 1; agg[e.name[:60]][1] += e.time_range.elapsed_us()
span = ev[-1].time_range.end - ev[0].time_range.start
print(f"GPU span {span/36:.1f} us/step over 36 steps; busy sum {sum(v[1] for v in agg.values())/36:.1f} us/step")
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:10]:
    print(f"  {t/36:8.1f} us/step  {c/36:5.1f} launches/step  {t/c:8.1f} us each  {k}")
# first 14 events of a step as a timeline
base = ev[40].time_range.start
for e in ev[40:58]:
    print(f"   +{e.time_range.start-base:8.1f} .. +{e.time_range.end-base:8.1f}  {e.name[:70]}")

"""Ad-hoc timings on the GPU box (CUDA events): GAE variants, sparse / bulk reset latency.  Not part of the product."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import radiation_ppo_b200 as rp

dev = torch.device("cuda:0")
def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts)//2]

out = {}
for N in (65536, 131072):
    T = 480
    g = torch.Generator(device=dev).manual_seed(1)
    rew = -0.5 * torch.rand(T, N, generator=g, device=dev) * 1.5
    val = torch.randn(T, N, generator=g, device=dev)
    end = (torch.rand(T, N, generator=g, device=dev) < 0.01).to(torch.uint8); end[T-1] = 1
    boot = torch.randn(T, N, generator=g, device=dev) * end
    adv, ret = torch.empty_like(rew), torch.empty_like(rew)
    for v in (3, 7, 8, 10, 1):
        if v == 2 and N > 65536: continue
        ms = timeit(lambda: rp.gae_advantages(rew, val, end, boot, adv=adv, ret=ret, variant=v))
        out[f"gae_N{N}_v{v}"] = dict(ms=ms, gbs=17 * T * N / ms / 1e6)
    del rew, val, end, boot, adv, ret
print(json.dumps(out, indent=1)); sys.exit(0)
N = 131072
env = rp.RadSearch(obstruction_count=5, enforce_grid_boundaries=True, num_envs=N, seed=2, auto_reset=True, fast_poisson=True)
for frac, name in ((1/120, "sparse"), (1/16, "mid"), (1.0, "bulk")):
    mask = (torch.rand(N, device=dev) < frac)
    out[f"reset_{name}_{int(mask.sum())}"] = timeit(lambda: env.reset_batch(mask=mask), n=5, warm=2)
out["reset_bulk_newobs"] = timeit(lambda: env.reset_batch(new_obstacles=True), n=3, warm=1)
acts = torch.randint(0, 8, (N, 1), device=dev, dtype=torch.int32)
out["step_only_ms"] = timeit(lambda: env.step_batch(acts, auto_reset=False), n=20, warm=5)
print(json.dumps(out, indent=1))

"""Ad-hoc timings on the GPU box (CUDA events).  Not part of the product."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import radiation_ppo_b200 as rp

dev = torch.device("cuda:0")
def timeit(fn, n=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts)//2]

out = {}
for (N, A) in ((16384, 4), (16384, 1), (65536, 1), (131072, 1), (32768, 2)):
    env = rp.RadSearch(obstruction_count=5, enforce_grid_boundaries=True, number_agents=A, num_envs=N, seed=4, auto_reset=True, fast_poisson=True)
    g = torch.Generator(device=dev).manual_seed(1)
    env._meta.add_(torch.randint(0, 120, (N,), generator=g, device=dev, dtype=torch.int32) << 16)
    acts = torch.randint(0, 8, (N, A), generator=g, device=dev, dtype=torch.int32)
    for _ in range(30): env.step_batch(acts)
    out[f"step+reset_N{N}_A{A}_us"] = 1e3 * timeit(lambda: env.step_batch(acts))
    out[f"step_only_N{N}_A{A}_us"] = 1e3 * timeit(lambda: env.step_batch(acts, auto_reset=False))
    if A == 4:
        mb = rp.BatchedMapsBuffer(N, A, 120, environment_scale=env.scale)
        pred = torch.rand(N, A, 2, device=dev)
        for _ in range(60):
            mb.update(env.obs, pred); env.step_batch(acts); mb.reset(mask=(env.ended & 4) != 0)
        out["maps_update_us"] = 1e3 * timeit(lambda: mb.update(env.obs, pred))
        m = (env.ended & 4) != 0
        out["maps_reset_masked_us"] = 1e3 * timeit(lambda: mb.reset(mask=m))
        out["mask_op_us"] = 1e3 * timeit(lambda: (env.ended & 4) != 0)
    del env
print(json.dumps(out, indent=1))

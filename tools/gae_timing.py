"""GAE variant timings on the GPU box (CUDA events).  python tools/gae_timing.py"""
import json, os, sys
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
    ts.sort(); return ts[len(ts) // 2]

out = {}
for N in (1024, 16384, 32768, 65536, 131072):
    T = 480
    g = torch.Generator(device=dev).manual_seed(1)
    rew = -0.5 * torch.rand(T, N, generator=g, device=dev) * 1.5
    val = torch.randn(T, N, generator=g, device=dev)
    end = (torch.rand(T, N, generator=g, device=dev) < 0.01).to(torch.uint8); end[T - 1] = 1
    boot = torch.randn(T, N, generator=g, device=dev) * end
    adv, ret = torch.empty_like(rew), torch.empty_like(rew)
    for v in (0, 7, 8, 11, 12, 13, 14, 15) if N % 128 == 0 and N >= 16384 else (0, 2, 3):
        ms = timeit(lambda: rp.gae_advantages(rew, val, end, boot, adv=adv, ret=ret, variant=v))
        out[f"N{N}_v{v}"] = f"{ms * 1e3:.0f} us  {17 * T * N / ms / 1e6:.0f} GB/s"
    del rew, val, end, boot, adv, ret
print(json.dumps(out, indent=1))

"""Full-size parity soak (not collected by pytest: run by hand on a GPU box, ~2 min):
  python tests/full_size_soak.py [n_envs] [steps]
BASELINE configs[4] size: 131072 envs, 5 obstructions, T = 480 with auto-reset (prefetch + CUDA graph), every output of
every step against the oracle (exact sampler, so that the counts are compared too)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import c_oracle as co
from tests import parity_util as pu
import radiation_ppo_b200 as rp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
T = int(sys.argv[2]) if len(sys.argv) > 2 else 480
ML, A = 120, 1
env = rp.RadSearch(obstruction_count=5, enforce_grid_boundaries=True, num_envs=n, seed=777, steps_per_episode=ML,
                   auto_reset=True, prefetch=True, use_cuda_graph=True)
ob = co.OracleBatch(n, co.default_config(obstruction_count=5, enforce=1, max_ep_len=ML), seed=777)
ob.reset()
g = torch.Generator(device=env.device).manual_seed(1)
stagger = torch.randint(0, ML, (n,), generator=g, device=env.device, dtype=torch.int32)
env._meta.add_(stagger << 16)
ob.envs["ep_len"] = stagger.cpu().numpy()
rng = np.random.default_rng(1)
t0 = time.time()
resets = 0
for t in range(1, T + 1):
    acts = rng.integers(0, 8, size=(n, A))
    epoch_end = t == T
    env.step_batch(torch.as_tensor(acts, dtype=torch.int32, device=env.device), epoch_end=epoch_end)
    ob.step(acts, env._ctr)
    e, o = ob.envs, ob.outs
    terminal, timeout = e["done"] == 1, e["ep_len"] == ML
    want = terminal * 1 | timeout * 2 | ((terminal | timeout | epoch_end) * 4)
    mask = (want & 4) != 0
    final = o["obs"][:, :A].copy()
    rew = o["reward"][:, :A].astype(np.float32).copy()
    if mask.any():
        ob.reset(mask=mask, new_obstacles=np.full(n, epoch_end))
    v = pu.GpuView(env)
    np.testing.assert_array_equal(v.ended, want)
    np.testing.assert_array_equal(v.reward, rew)
    pu.compare_obs(v.final_obs, final, sel=np.where(mask)[0])
    pu.compare_obs(v.obs, np.where(mask[:, None, None], o["obs"][:, :A], final))
    pu.compare_state(v, ob, A)
    resets += int(mask.sum())
print(f"full-size soak ok: {n} envs x {T} steps = {n * T / 1e6:.1f} M env-steps, {resets} resets, bit-exact vs the oracle "
      f"({time.time() - t0:.0f} s)")

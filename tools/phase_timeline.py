"""Phase-level timeline of the step kernel (RS_TUNE=8: clock64 stamps by thread 0 of every CTA).
  RS_TUNE=8 python tools/phase_timeline.py [N]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import radiation_ppo_b200 as rp
from radiation_ppo_b200 import _lib as L
N = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
env = rp.RadSearch(obstruction_count=5, enforce_grid_boundaries=True, num_envs=N, seed=4, auto_reset=True, fast_poisson=True)
g = torch.Generator(device=env.device).manual_seed(1)
env._meta.add_(torch.randint(0, 120, (N,), generator=g, device=env.device, dtype=torch.int32) << 16)
acts = torch.randint(0, 8, (N, 1), generator=g, device=env.device, dtype=torch.int32)
for _ in range(40): env.step_batch(acts)
torch.cuda.synchronize()
env.step_batch(acts, auto_reset=False)
n_cta = min(2048, (N + 127) // 128)
buf = np.zeros((n_cta, 12), np.int64)
L.check(L.load().rs_debug_timeline(buf.ctypes.data_as(C.c_void_p), n_cta), "timeline")
names = ["load", "move(+push)", "seed", "pairs", "hint+sense+count", "commit", "store"]
d = np.diff(buf[:, :8], axis=1)
tot = buf[:, 7] - buf[:, 0]
print(f"N={N} CTAs={n_cta} cycles per CTA: total median {np.median(tot):.0f} (p10 {np.percentile(tot,10):.0f}, p90 {np.percentile(tot,90):.0f}) = {np.median(tot)/1965:.1f} us at 1965 MHz")
for i, nm in enumerate(names):
    print(f"  {nm:18s} median {np.median(d[:, i]):8.0f}  p90 {np.percentile(d[:, i], 90):8.0f}   {100*np.median(d[:, i])/np.median(tot):5.1f}%")
st, en = buf[:, 8] - buf[:, 8].min(), buf[:, 9] - buf[:, 8].min()
print(f"globaltimer (ns): CTA starts p0/p50/p90/max {st.min()}/{np.median(st):.0f}/{np.percentile(st,90):.0f}/{st.max()}  "
      f"ends p10/p50/p90/max {np.percentile(en,10):.0f}/{np.median(en):.0f}/{np.percentile(en,90):.0f}/{en.max()}")
dur = (buf[:, 9] - buf[:, 8])
print(f"CTA duration ns p10/p50/p90/max {np.percentile(dur,10):.0f}/{np.median(dur):.0f}/{np.percentile(dur,90):.0f}/{dur.max()}")
late = np.argsort(en)[-5:]
print("latest CTAs:", [(int(b), int(st[b]), int(en[b])) for b in late])

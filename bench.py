"""bench.py -- RadSearch env-steps/s (5 obstructions) on N B200s of one node, plus GAE GB/s, roofline and CPU baseline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--envs-per-gpu E]
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one batched environment step of every env on every GPU: rs_step (all agents: move, collision, shortest
path, LOS, expected counts, Philox+Poisson, 8 sensors, reward/terminal, caller rules) followed by rs_reset for the envs
that finished (auto-reset).  Workload = BASELINE.json configs[4]: 131,072 envs per GPU, 5 obstructions, enforced
boundaries, uniform random actions 0..7, episodes staggered so that ~1/120 of the envs reset at every step.
Weak scaling: envs shard independently over ranks, no data-path collective.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "RadSearch env-steps/sec (5 obstructions)"
UNIT = "env-steps/s"
K_OBS = 5
BYTES_PER_ENV_STEP = 122 + 32 * K_OBS        # algorithmic bytes per env-step, single agent
GAE_BYTES_PER_ELEM = 17                      # rew 4 + val 4 + path_end 1 + adv 4 + ret 4
T_EPOCH = 480


def measured_traffic(kernel: str, n_env: int):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture (profiles/)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[kernel]
        if t.get("n_env", t.get("N")) == n_env:
            return t["dram_bytes_read"] + t["dram_bytes_write"]
    except Exception:
        pass
    return None


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def cpu_baseline(n_envs: int, T: int, threads: int):
    """The oracle's scalar per-env loop (reference semantics, OpenMP over envs) on the host cores."""
    from oracle import c_oracle as co

    ob = co.OracleBatch(n_envs, co.default_config(obstruction_count=K_OBS, enforce=1), seed=2, threads=threads)
    ob.reset()
    t0 = time.perf_counter()
    n, chk = ob.rollout(T, 1, epoch_end_last=False)
    dt = time.perf_counter() - t0
    return n / dt, dt, chk


def run_reference(args):
    """--impl reference: the CPU path (oracle port of the reference's per-env loop) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import c_oracle as co

    cores = os.cpu_count() or 1
    n = 16384                    # per step: enough envs per thread for the OpenMP loop to run at its best rate
    ob = co.OracleBatch(n, co.default_config(obstruction_count=K_OBS, enforce=1), seed=2, threads=cores)
    ob.reset()
    # one reference step = 16384 envs x `inner` steps: 8 (= 131072 env-steps, the GPU arm's units per step) unless the
    # requested number of steps would then run for more than ~90 s on these cores
    t0 = time.perf_counter()
    ob.rollout(2, 1, epoch_end_last=False)
    rate = 2 * n / (time.perf_counter() - t0)
    inner = int(max(1, min(8, rate * 90.0 / max(args.steps + args.warmup, 1) / n)))
    ctr = 3
    for _ in range(args.warmup):
        ob.rollout(inner, ctr, epoch_end_last=False); ctr += inner
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ob.rollout(inner, ctr, epoch_end_last=False); ctr += inner
    dt = time.perf_counter() - t0
    v = n * inner * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32+f64", "data": "synthetic",
            "config": {"workload": "RadSearch env step, 5 obstructions, enforced boundaries, random actions, auto-reset; "
                                   f"one step = {n} envs x {inner} steps = {n * inner} env-steps (the GPU arm's units per step)"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{n} envs x {inner * args.steps} steps, oracle/radsearch_oracle.c (scalar per-env loop, OpenMP)"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12000)
    ap.add_argument("--warmup", type=int, default=240)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=131072)
    ap.add_argument("--ring", type=int, default=4, help="env batches cycled so that each step's state comes from HBM")
    ap.add_argument("--exact-poisson", action="store_true", help="fp64 numpy-exact PTRS acceptance instead of fp32")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="device-resident throughput and the step-kernel timing only "
                    "(what the ncu passes replay)")
    ap.add_argument("--no-prefetch", action="store_true", help="synchronous reset kernel after every step")
    ap.add_argument("--no-graph", action="store_true", help="plain stream launches instead of CUDA-graph replay")
    ap.add_argument("--episode-steps", type=int, default=120, help="steps_per_episode (120 = the reference's; a huge value "
                    "shows the throughput without resets)")
    ap.add_argument("--streams", type=int, default=0,
                    help="CUDA streams the ring's env batches are spread over (0 = one per batch; 1 = serialised)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import radiation_ppo_b200 as rp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    N, K, W, R = args.envs_per_gpu, args.steps, max(args.warmup, 3), max(args.ring, 1)
    fast = not args.exact_poisson

    # ---- R independent env batches of N envs (global env ids: rank-major, then ring slot) --------------------------
    envs = []
    for r in range(R):
        e = rp.RadSearch(obstruction_count=K_OBS, enforce_grid_boundaries=True, num_envs=N, seed=2, device=dev,
                         env_id_offset=(rank * R + r) * N, auto_reset=True, fast_poisson=fast,
                         steps_per_episode=args.episode_steps,
                         prefetch=not args.no_prefetch, use_cuda_graph=not (args.no_graph or args.no_prefetch))
        # stagger the episodes: steady state of a training run (about 1/120 of the envs finish at every step)
        g = torch.Generator(device=dev).manual_seed(1000 + rank * R + r)
        e._meta.add_(torch.randint(0, min(args.episode_steps, 120), (N,), generator=g, device=dev, dtype=torch.int32) << 16)
        torch.cuda.synchronize()
        e.capture_graphs()
        envs.append(e)
    g = torch.Generator(device=dev).manual_seed(7 + rank)
    n_act = 16
    actions = torch.randint(0, 8, (n_act, N, 1), generator=g, device=dev, dtype=torch.int32)   # resident in HBM
    state_mb = R * N * (BYTES_PER_ENV_STEP + 160 + 80) / 1e6

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # The env batches of the ring are independent vector envs: batch r is stepped on its own stream, so that the short
    # tail kernels of one batch (reset of finished envs, counter bump) and the drain of its step kernel's last CTAs
    # overlap the step kernel of the next batch instead of idling the GPU.
    n_streams = R if args.streams <= 0 else min(args.streams, R)
    main_stream = torch.cuda.current_stream(dev)
    streams = [torch.cuda.Stream(device=dev) for _ in range(n_streams)] if n_streams > 1 else [main_stream]

    def one_step(i):
        with torch.cuda.stream(streams[(i % R) % n_streams]):
            return envs[i % R].step_batch(actions[i % n_act], epoch_end=False)

    def fork():
        for st in streams:
            st.wait_stream(main_stream)

    def join():
        for st in streams:
            main_stream.wait_stream(st)

    # ---- device-resident throughput: `value` -----------------------------------------------------------------------
    fork()
    for i in range(W):
        one_step(i)
    join()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # nvidia-smi samples every 100 ms: when the K timed steps last less than that, the same steps keep running (untimed)
    # in front of the timed region so that the samples are taken under this very load
    t_pre = time.perf_counter()
    i_pre = 0
    while time.perf_counter() - t_pre < 0.6:
        fork()
        for _ in range(200):
            one_step(W + i_pre)
            i_pre += 1
        join()
        torch.cuda.synchronize()
    W0, W = W, W + i_pre
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    fork()
    for i in range(K):
        one_step(W + i)
    join()
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    value = N * world * K / (total_ms / 1e3)

    # ---- roofline pass: CUDA events around the step kernel alone (same stream), averaged over K launches ------------
    import ctypes as C
    from radiation_ppo_b200 import _lib as L

    lib = L.load()
    Kr = min(K, 480)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(Kr)]
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    for e in envs:
        e._quiesce_prefetch()
    torch.cuda.synchronize()
    sflags = L.F_AUTO_RESET | L.F_DEVICE_CTR | (L.F_FAST_POISSON if fast else 0)     # DEVICE_CTR: no memset inside
    for i in range(Kr):
        e = envs[i % R]
        a = actions[i % n_act]
        e._ctr += 1
        e._ctr_dev.fill_(e._ctr)
        e._reset_count.zero_()
        kev[i][0].record()
        L.check(lib.rs_step(C.byref(e._cfg), C.byref(e._st), C.c_void_p(a.data_ptr()), C.c_void_p(e.obs.data_ptr()),
                            C.c_void_p(e.reward.data_ptr()), C.c_void_p(e.team_reward.data_ptr()),
                            C.c_void_p(e.done_flags.data_ptr()), C.c_void_p(e.info_flags.data_ptr()),
                            C.c_void_p(e.ended.data_ptr()), C.c_void_p(e.final_obs.data_ptr()), N, e.env_id_offset,
                            e.seed, e._ctr, None, 0, sflags, stream), "rs_step")
        kev[i][1].record()
        L.check(lib.rs_reset(C.byref(e._cfg), C.byref(e._st), None, None, C.c_void_p(e.obs.data_ptr()), N,
                             e.env_id_offset, e.seed, e._ctr, None, 0,
                             L.F_RESET_LIST | (L.F_FAST_POISSON if fast else 0), stream), "rs_reset")
        e._ctr_dev_val = -1
    torch.cuda.synchronize()
    k_ms = sum(a.elapsed_time(b) for a, b in kev) / Kr
    peak, peak_src = measured_peak_gbs()
    achieved = BYTES_PER_ENV_STEP * N / (k_ms / 1e3) / 1e9

    if args.quick:
        if rank == 0:
            print(json.dumps({"quick": True, "value": value, "ms_per_step": total_ms / K, "kernel_ms": k_ms,
                              "frac": achieved / peak}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- GAE over the [T, N] rollout buffer ("GAE GB/s vs HBM peak") -------------------------------------------------
    T = T_EPOCH
    rew = -0.5 * torch.rand(T, N, generator=g, device=dev) * 1.5
    val = torch.randn(T, N, generator=g, device=dev)
    end = (torch.rand(T, N, generator=g, device=dev) < 1 / 100).to(torch.uint8)
    end[T - 1] = 1
    boot = torch.randn(T, N, generator=g, device=dev) * end
    adv, ret = torch.empty_like(rew), torch.empty_like(rew)
    for _ in range(3):
        rp.gae_advantages(rew, val, end, boot, adv=adv, ret=ret, variant=1)
    torch.cuda.synchronize()
    gev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    for a, b in gev:
        a.record()
        rp.gae_advantages(rew, val, end, boot, adv=adv, ret=ret, variant=1)
        b.record()
    torch.cuda.synchronize()
    gae_ms = sorted(a.elapsed_time(b) for a, b in gev)[len(gev) // 2]
    gae_gbs = GAE_BYTES_PER_ELEM * T * N / (gae_ms / 1e3) / 1e9
    del rew, val, end, boot, adv, ret

    # ---- RAD-TEAM pipeline of BASELINE configs[3]: 16,384 envs x 4 agents, env step -> shared map observation ---------
    maps_line = None
    if world == 1:
        Nm, Am, Tm = 16384, 4, 120
        menv = rp.RadSearch(obstruction_count=K_OBS, enforce_grid_boundaries=True, number_agents=Am, num_envs=Nm, seed=4,
                            device=dev, auto_reset=True, fast_poisson=fast)
        mb = rp.BatchedMapsBuffer(Nm, Am, 120, environment_scale=menv.scale, device=dev)
        macts = torch.randint(0, 8, (8, Nm, Am), generator=g, device=dev, dtype=torch.int32)
        mpred = torch.rand(Nm, Am, 2, generator=g, device=dev)

        def maps_step(i):
            mb.update(menv.obs, mpred)
            menv.step_batch(macts[i % 8])
            mb.reset(mask=menv.ended, mask_bits=4)

        for i in range(20):
            maps_step(i)
        torch.cuda.synchronize()
        mev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(Tm)]
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for i in range(Tm):
            mev[i][0].record()
            mb.update(menv.obs, mpred)
            mev[i][1].record()
            menv.step_batch(macts[i % 8])
            mb.reset(mask=menv.ended, mask_bits=4)
        p1.record()
        torch.cuda.synchronize()
        upd_ms = sorted(a.elapsed_time(b) for a, b in mev)[Tm // 2]
        maps_line = {"workload": f"{Nm} envs x {Am} agents, 27x27 maps, env step + rs_maps_update + rs_maps_reset (BASELINE configs[3])",
                     "update_ms": upd_ms, "agent_map_updates_per_s": Nm * Am / (upd_ms / 1e3),
                     "pipeline_env_steps_per_s": Nm * Tm / (p0.elapsed_time(p1) / 1e3),
                     "maps_status_flags": int(mb.status.sum().item())}
        del menv, mb

    # ---- end to end through the public API with HOST buffers (pinned), copies inside the timed region ----------------
    # RadSearch.step_host: pinned host actions -> device, step + auto-reset, ALL step outputs -> pinned host in one
    # transfer, on the env batch's own stream.  `e2e`: the R env batches of the ring are driven round-robin, the host
    # waiting for batch r's previous results before it sends batch r's next actions (an asynchronous vector-env loop:
    # copies of one batch overlap the kernels of the others).  `e2e_sync`: one batch at a time, host waits every step.
    h_act = torch.randint(0, 8, (n_act, N, 1), dtype=torch.int32).pin_memory()
    hbs = [e.host_buffers() for e in envs]
    torch.cuda.synchronize()
    Ke = min(K, 480)

    def e2e_run(k0, k1, depth_all):
        chk = 0
        for i in range(k0, k1):
            r = i % R
            hb = hbs[r]
            if depth_all:
                hb.wait()                                      # results of this batch's previous step are on the host
            envs[r].step_host(hb, actions=h_act[i % n_act])
            if not depth_all:
                hb.wait()
            chk += int(hb.ended[0])                            # the host touches the results
        for hb in hbs:
            hb.wait()
        return chk

    e2e_vals = {}
    for name, depth_all in (("sync", False), ("pipelined", True)):
        e2e_run(0, W, depth_all)
        barrier()
        t0 = time.perf_counter()
        e2e_run(W, W + Ke, depth_all)
        barrier()
        e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        e2e_vals[name] = N * world * Ke / float(e2e_s.item())
    # both drive RadSearch.step_host with host buffers; report the faster schedule (with many ranks on one host the
    # pipelined one can lose to the synchronous one: more copies in flight than the host side can absorb)
    e2e_mode = "pipelined" if e2e_vals["pipelined"] >= e2e_vals["sync"] else "sync"
    e2e_value = e2e_vals[e2e_mode]
    h2d, d2h = hbs[0].h2d_bytes, hbs[0].d2h_bytes

    status = int(sum(int((e.status & ~2).any()) for e in envs))

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W0, "preroll_steps": i_pre,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32+f64", "data": "synthetic",
            "config": {"workload": f"RadSearch env step + auto-reset, {N} envs/GPU (BASELINE configs[4]), 5 obstructions, "
                                   f"enforced boundaries, 1 agent, uniform random actions, staggered {args.episode_steps}-step episodes",
                       "envs_per_gpu": N, "obstructions": K_OBS, "poisson": "fp32-acceptance PTRS" if fast else "numpy-exact PTRS",
                       "l2": f"ring of {R} env batches ({state_mb:.0f} MB of state) cycled: every step reads its state from HBM",
                       "resets": ("next episodes prefetched by rs_prepare on a parallel graph branch / side stream"
                                  if not args.no_prefetch else "synchronous rs_reset after every step"),
                       "launch": ("CUDA graph replay" if not (args.no_graph or args.no_prefetch) else "stream launches") +
                                 f", ring batches on {n_streams} stream(s)",
                       "parallelism": f"env-sharded x{world}, no data-path collective"},
            "roofline": {"bound": "hbm", "kernel": "step_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": measured_traffic("step_kernel", N), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": BYTES_PER_ENV_STEP * N,
                         "algorithmic_bytes_per_env_step": BYTES_PER_ENV_STEP, "kernel_ms": k_ms,
                         "kernel_env_steps_per_s": N / (k_ms / 1e3)},
            "gae": {"T": T, "N": N, "ms": gae_ms, "achieved": gae_gbs, "peak": peak, "unit": "GB/s", "frac": gae_gbs / peak,
                    "bytes_per_element": GAE_BYTES_PER_ELEM, "kernel": "gae_tile_kernel<128,8,3> (bulk-async tiles, per-column fp64 recurrence)",
                    "traffic": measured_traffic("gae_cols_kernel", N) if T == 480 else None},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": Ke, "schedule": e2e_mode, "sync_value": e2e_vals["sync"],
                    "pipelined_value": e2e_vals["pipelined"],
                    "note": f"RadSearch.step_host: pinned host actions -> device, step+reset, all outputs (obs, reward, "
                            f"done/info/ended flags; 51 B per env) -> pinned host in one copy; pipelined_value = {R} env batches "
                            "round-robin on their own streams (host waits for a batch's previous results before sending its "
                            "next actions); sync_value = host waits after every step; value = the faster of the two"},
            "maps": maps_line,
            "gpu_launches": int((2 + 1 / rp.RadSearch.PREFETCH_PERIOD) * K) if not args.no_prefetch else 2 * K, "clocks": clocks, "status_flags_raised": status,
        }
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            # GAE on the host cores: the oracle's per-column float64 recurrence (the reference's lfilter arithmetic), OpenMP
            import numpy as np
            from oracle import c_oracle as co
            Tg, Ng = T_EPOCH, 16384
            rg = np.random.default_rng(0)
            c_rew = (-0.5 * rg.uniform(0, 1.5, (Tg, Ng))).astype(np.float32)
            c_val = rg.normal(size=(Tg, Ng)).astype(np.float32)
            c_end = (rg.random((Tg, Ng)) < 0.01).astype(np.uint8); c_end[-1] = 1
            c_boot = (rg.normal(size=(Tg, Ng)) * c_end).astype(np.float32)
            co.gae(c_rew, c_val, c_end, c_boot, threads=cores)
            t0 = time.perf_counter()
            for _ in range(3):
                co.gae(c_rew, c_val, c_end, c_boot, threads=cores)
            dtg = (time.perf_counter() - t0) / 3
            line["gae"]["cpu_baseline"] = {"value": GAE_BYTES_PER_ELEM * Tg * Ng / dtg / 1e9, "unit": "GB/s", "cores": cores,
                                           "kind": "port", "sample": f"[{Tg}, {Ng}] rollout, oracle orc_gae (OpenMP over columns)"}
            v1, dt1, _ = cpu_baseline(256, 60, cores)                 # calibrate
            n_s = max(256, min(16384, int(256 * 12.0 / max(dt1, 1e-3)) // 256 * 256))
            v, dt, _ = cpu_baseline(n_s, 60, cores)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{n_s} envs x 60 steps in {dt:.1f}s, oracle/radsearch_oracle.c "
                                              "(scalar per-env loop with per-step Dijkstra, OpenMP over envs)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

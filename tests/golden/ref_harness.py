"""TEST INFRASTRUCTURE ONLY -- drives the unmodified reference RadSearch (through oracle/shims) and records
trajectories: state before each call, actions, the uniforms numpy's poisson consumed, and every output.

Used by tests/golden/make_golden.py (writes tests/golden/*.npz) and by tests/test_oracle_vs_reference.py (container only).
"""
from __future__ import annotations

import numpy as np

from oracle.ref_env import RecordingGenerator, load_reference_env

N_UNIFORMS = 32


def rect_of(poly):
    xs = [p[0] for p in poly]
    ys = [p[1] for p in poly]
    return [int(min(xs)), int(min(ys)), int(max(xs)), int(max(ys))]


def snapshot(env, K=8):
    """Scenario state of a reference env as plain ints (what rs_load_scenarios takes)."""
    A = len(env.agents)
    rects = np.zeros((K, 4), np.int32)
    for k, poly in enumerate(env.poly[: env.num_obs] if env.num_obs > 0 else []):
        rects[k] = rect_of(poly)
    return dict(
        src=np.array(env.src_coords, np.int32),
        det=np.array([env.agents[i].det_coords for i in range(A)], np.int32),
        intensity=np.int32(env.intensity),
        bkg=np.int32(env.bkg_intensity),
        rects=rects,
        num_obs=np.int32(env.num_obs),
        best=np.array([env.agents[i].prev_det_dist for i in range(A)], np.float64),
    )


def record_episodes(seed, n_steps, obstruction_count=5, enforce=True, n_agents=1, action_seed=0, idle_prob=0.0,
                    max_ep_len=120, epoch_len=480):
    """Run the reference env for n_steps with random actions under the caller rules of train.py:394-548 and return a
    list of records: dict(kind='reset'|'step', pre=state, actions, uniforms[A,N_UNIFORMS], obs[A,11], reward[A],
    team_reward, done[A], oob[A], oob_count[A], blocked[A], det[A,2], sp[A], best[A])."""
    m = load_reference_env()
    rng = RecordingGenerator(np.random.default_rng(seed), N_UNIFORMS)
    env = m.RadSearch(obstruction_count=obstruction_count, np_random=rng, enforce_grid_boundaries=enforce,
                      number_agents=n_agents)
    arng = np.random.default_rng(action_seed)
    A = n_agents
    recs = []

    def pack(kind, pre, actions, ret, n_calls_before):
        obs, rew, done, info = ret
        calls = rng.poisson_log[n_calls_before:]
        assert len(calls) == A, (kind, len(calls))
        return dict(
            kind=kind, pre=pre, actions=np.array(actions if actions is not None else [-1] * A, np.int32),
            uniforms=np.stack([c[2] for c in calls]), lam=np.array([c[0] for c in calls]),
            obs=np.stack([np.asarray(obs[i], np.float64) for i in range(A)]),
            reward=np.array([rew["individual_reward"][i] for i in range(A)], np.float64),
            team_reward=np.float64(np.nan if rew["team_reward"] is None else rew["team_reward"]),
            done=np.array([done[i] for i in range(A)], np.int32),
            oob=np.array([info[i]["out_of_bounds"] for i in range(A)], np.int32),
            oob_count=np.array([info[i]["out_of_bounds_count"] for i in range(A)], np.int32),
            blocked=np.array([info[i]["blocked"] for i in range(A)], np.int32),
            det=np.array([env.agents[i].det_coords for i in range(A)], np.int32),
            sp=np.array([env.agents[i].sp_dist for i in range(A)], np.float64),
            best=np.array([env.agents[i].prev_det_dist for i in range(A)], np.float64),
            los=np.array([env.agents[i].intersect for i in range(A)], np.int32),
        )

    def do_reset():
        # the env was constructed (or epoch_end set) by the caller; reset() draws a scenario and takes step(None)
        n0 = len(rng.poisson_log)
        ret = env.reset()
        # a recursive "not valid" retry draws more than one probe step; keep the last A calls
        rng.poisson_log[n0:] = rng.poisson_log[-A:]
        pre = snapshot(env)
        recs.append(pack("reset", pre, None, ret, n0))

    do_reset()
    ep_len = 0
    for t in range(n_steps):
        acts = [int(8 if arng.random() < idle_prob else arng.integers(0, 8)) for _ in range(A)]
        pre = snapshot(env)
        pre["iter_count"] = np.int32(env.iter_count)
        pre["done"] = np.int32(env.done)
        pre["oob_count"] = np.array([env.agents[i].out_of_bounds_count for i in range(A)], np.int32)
        pre["blocked"] = np.array([env.agents[i].obstacle_blocking for i in range(A)], np.int32)
        pre["sp"] = np.array([env.agents[i].sp_dist for i in range(A)], np.float64)
        pre["euc"] = np.array([env.agents[i].euc_dist for i in range(A)], np.float64)
        n0 = len(rng.poisson_log)
        ret = env.step({i: acts[i] for i in range(A)})
        recs.append(pack("step", pre, acts, ret, n0))
        ep_len += 1
        timeout = ep_len == max_ep_len
        over = any(ret[2].values()) or timeout
        epoch_ended = (t % epoch_len) == epoch_len - 1
        if over or epoch_ended:
            if epoch_ended:
                env.epoch_end = True
            do_reset()
            ep_len = 0
    return recs


def _pre_full(env):
    A = len(env.agents)
    pre = snapshot(env)
    pre["iter_count"] = np.int32(env.iter_count)
    pre["done"] = np.int32(env.done)
    pre["oob_count"] = np.array([env.agents[i].out_of_bounds_count for i in range(A)], np.int32)
    pre["blocked"] = np.array([env.agents[i].obstacle_blocking for i in range(A)], np.int32)
    pre["sp"] = np.array([env.agents[i].sp_dist for i in range(A)], np.float64)
    pre["euc"] = np.array([env.agents[i].euc_dist for i in range(A)], np.float64)
    return pre


def record_probes(seed, n_scenarios, probes_per_scenario, obstruction_count=5, enforce=True, n_agents=1):
    """Adversarial single-step probes: the detector(s) are teleported (the way refresh_environment sets them,
    rad_search_env.py:821-824) onto / next to obstruction edges and corners, next to the walls and next to the
    source, then one step is taken.  Records have the same fields as record_episodes (kind='step')."""
    m = load_reference_env()
    import visilibity as vis  # the shim

    rng = RecordingGenerator(np.random.default_rng(seed), N_UNIFORMS)
    env = m.RadSearch(obstruction_count=obstruction_count, np_random=rng, enforce_grid_boundaries=enforce,
                      number_agents=n_agents)
    prng = np.random.default_rng(seed + 1000)
    A = n_agents
    recs = []
    offs = [0, 1, 29, 50, 70, 71, 72, 99, 100, 101, 109, 110, 111]

    def candidate():
        kind = prng.integers(0, 10)
        if env.num_obs > 0 and kind < 7:
            r = rect_of(env.poly[int(prng.integers(0, env.num_obs))])
            x0, y0, x1, y1 = r
            side = prng.integers(0, 8)
            o = int(prng.choice(offs))
            along_x = int(prng.integers(x0 - 120, x1 + 121))
            along_y = int(prng.integers(y0 - 120, y1 + 121))
            if side == 0:
                return (x0 - o, along_y)
            if side == 1:
                return (x1 + o, along_y)
            if side == 2:
                return (along_x, y0 - o)
            if side == 3:
                return (along_x, y1 + o)
            cx, cy = [(x0, y0), (x0, y1), (x1, y1), (x1, y0)][side - 4]
            sx, sy = [(-1, -1), (-1, 1), (1, 1), (1, -1)][side - 4]
            o2 = int(prng.choice(offs))
            return (cx + sx * o * int(prng.integers(0, 2)), cy + sy * o2 * int(prng.integers(0, 2)))
        if kind < 8:  # near a wall
            o = int(prng.choice(offs))
            w = prng.integers(0, 4)
            t = int(prng.integers(0, 2700))
            return [(o, t), (2699 - o, t), (t, o), (t, 2699 - o)][w]
        # near the source
        o = int(prng.choice([0, 50, 100, 109, 110, 111, 150, 200]))
        ang = prng.integers(0, 8)
        c = [(-1, 0), (-1, 1), (0, 1), (1, 1), (1, 0), (1, -1), (0, -1), (-1, -1)][ang]
        return (int(env.src_coords[0]) + c[0] * o, int(env.src_coords[1]) + c[1] * o)

    def legal(p):
        if not (0 <= p[0] < 2700 and 0 <= p[1] < 2700):
            return not enforce and -500 < p[0] < 3200 and -500 < p[1] < 3200
        for poly in env.poly[: env.num_obs] if env.num_obs > 0 else []:
            x0, y0, x1, y1 = rect_of(poly)
            if x0 < p[0] < x1 and y0 < p[1] < y1:
                return False
        return True

    for s in range(n_scenarios):
        env.epoch_end = True
        env.reset()
        for _ in range(probes_per_scenario):
            env.iter_count = int(prng.integers(0, 2))
            same = prng.random() < 0.3
            base = None
            for i in range(A):
                while True:
                    p = candidate() if not (same and base is not None) else base
                    if legal(p):
                        break
                    base = None
                base = p
                ag = env.agents[i]
                ag.det_coords = (float(p[0]), float(p[1]))
                ag.detector = vis.Point(float(p[0]), float(p[1]))
                # keep the agent self-consistent, as refresh_environment does (R:864-868): sp/euc belong to the new
                # position; the running minimum may be anything at or around it
                ag.sp_dist = env.world.shortest_path(env.source, ag.detector, env.vis_graph, m.EPSILON).length()
                ag.euc_dist = m.dist_p(ag.det_coords, env.src_coords)
                # (at iter_count == 0 the reference takes sp from prev_det_dist, R:551-553: they must agree there)
                ag.prev_det_dist = ag.sp_dist + (float(prng.choice([0.0, 0.0, -37.5, 60.25])) if env.iter_count else 0.0)
            env.done = False
            acts = [int(prng.integers(0, 9)) for _ in range(A)]
            pre = _pre_full(env)
            n0 = len(rng.poisson_log)
            try:
                ret = env.step({i: acts[i] for i in range(A)})
            except (ValueError, ZeroDivisionError, OverflowError) as ex:  # detector exactly on the source
                del rng.poisson_log[n0:]
                continue
            obs, rew, done, info = ret
            calls = rng.poisson_log[n0:]
            recs.append(dict(
                kind="step", pre=pre, actions=np.array(acts, np.int32),
                uniforms=np.stack([c[2] for c in calls]), lam=np.array([c[0] for c in calls]),
                obs=np.stack([np.asarray(obs[i], np.float64) for i in range(A)]),
                reward=np.array([rew["individual_reward"][i] for i in range(A)], np.float64),
                team_reward=np.float64(np.nan if rew["team_reward"] is None else rew["team_reward"]),
                done=np.array([done[i] for i in range(A)], np.int32),
                oob=np.array([info[i]["out_of_bounds"] for i in range(A)], np.int32),
                oob_count=np.array([info[i]["out_of_bounds_count"] for i in range(A)], np.int32),
                blocked=np.array([info[i]["blocked"] for i in range(A)], np.int32),
                det=np.array([env.agents[i].det_coords for i in range(A)], np.int32),
                sp=np.array([env.agents[i].sp_dist for i in range(A)], np.float64),
                best=np.array([env.agents[i].prev_det_dist for i in range(A)], np.float64),
                los=np.array([env.agents[i].intersect for i in range(A)], np.int32),
            ))
    return recs

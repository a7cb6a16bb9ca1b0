"""CPU tests: the oracle (oracle/radsearch_oracle.c) against the golden vectors recorded from the unmodified reference
(tests/golden/make_golden.py), numpy's own poisson, Python's round, the Philox known-answer vectors and scipy's lfilter."""
import copy

import numpy as np
import pytest

from oracle import c_oracle as co
from tests import parity_util as pu


@pytest.mark.parametrize("name", list(pu.STEP_FILES))
def test_oracle_reproduces_reference_records(name):
    g = pu.load_golden(name)
    kw = pu.STEP_FILES[name]
    envs, outs = pu.step_records_with_oracle(g, kw)
    assert not envs["status"].any()
    pu.assert_matches_golden(g, envs, outs, kw["n_agents"], obs_exact=True, label=name)
    np.testing.assert_array_equal(envs["sp"][:, : kw["n_agents"]], g["out_sp"])
    np.testing.assert_array_equal(envs["los_blocked"][:, : kw["n_agents"]], g["out_los"])
    np.testing.assert_array_equal(outs["lam"][:, : kw["n_agents"]], g["out_lam"])


def test_reset_records_shortest_path_matches_reference():
    # prev_det_dist recorded after reference resets == oracle shortest path from the same scenario
    for name, kw in pu.STEP_FILES.items():
        g = pu.load_golden(name)
        ob = pu.oracle_load_records(g, kw)
        idx = np.where(g["is_reset"])[0]
        for i in idx[:40]:
            assert ob.shortest_path(i, g["pre_det"][i, 0]) == g["pre_best"][i, 0]


def test_action_lut_matches_get_step():
    lut = pu.load_golden("action_lut")["step"]
    want = [(-100, 0), (-71, 71), (0, 100), (71, 71), (100, 0), (71, -71), (0, -100), (-71, -71), (0, 0)]
    np.testing.assert_array_equal(lut, np.array(want, np.float64))


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    assert co.philox([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert co.philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert co.philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_poisson_identical_to_numpy_under_injected_uniforms():
    rng = np.random.default_rng(123)
    lams = np.concatenate([np.random.default_rng(1).uniform(10, 100000, 20000),
                           np.random.default_rng(2).uniform(0.01, 10, 4000),
                           np.random.default_rng(3).uniform(10, 60, 16000), [0.0, 10.0, 9.999999]])
    for lam in lams:
        st = copy.deepcopy(rng.bit_generator.state)
        k = rng.poisson(lam)
        probe = np.random.Generator(np.random.PCG64())
        probe.bit_generator.state = st
        k2, status = co.poisson_injected(float(lam), probe.random(64))
        assert (k2, status) == (k, 0), lam


def test_round2_is_python_round():
    rng = np.random.default_rng(0)
    xs = list(rng.uniform(-2, 0.2, 20000)) + [-0.5 * s / 2000.0 for s in range(0, 6000)] + [
        i / 1000.0 for i in range(-300, 100)] + [0.125, -0.125, 0.375, 0.1, 2.675, -0.285, 1e-17]
    for x in xs:
        assert co.round2(float(x)) == round(float(x), 2), x


def test_gae_golden_reference_ppobuffer():
    g = pu.load_golden("ref_gae")
    adv, ret = co.gae(g["rew"], g["val"], g["end"], g["boot"])
    np.testing.assert_array_equal(adv, g["adv"])
    np.testing.assert_array_equal(ret, g["ret"])
    # the float32-ndarray bootstrap variant (RADA2C_core.py:549) agrees to the stated tolerance
    np.testing.assert_allclose(adv, g["adv32"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(ret, g["ret32"], rtol=1e-5, atol=2e-6)
    # known-answer vectors of unit_tests/test_PPO.py:259-286, 462-497
    T = 10
    rew = g["kat_rewards"].astype(np.float32).reshape(T, 1)
    val = g["kat_values"].astype(np.float32).reshape(T, 1)
    end = np.zeros((T, 1), np.uint8)
    boot = np.zeros((T, 1), np.float32)
    boot[-1] = g["kat_last_val"]
    a, r = co.gae(rew, val, end, boot)
    np.testing.assert_allclose(a[:, 0], g["kat_adv"], rtol=1e-5)
    np.testing.assert_allclose(r[:, 0], g["kat_ret"], rtol=1e-5)


def test_gae_matches_scipy_lfilter_restatement():
    rew, val, end, boot = pu.synthetic_rollout(200, 37, seed=5, max_ep=50)
    a0, r0 = pu.gae_numpy_reference(rew, val, end, boot)
    a1, r1 = co.gae(rew, val, end, boot)
    np.testing.assert_array_equal(a1, a0)
    np.testing.assert_array_equal(r1, r0)


def test_oracle_reset_distribution_matches_reference():
    """Philox-driven reset sampler vs 1500 resets of the reference (PCG64): same marginals (two-sample KS / chi2)."""
    from scipy import stats

    g = pu.load_golden("ref_reset_stats")
    rows = g["rows"]
    cfg = co.default_config(obstruction_count=-1, enforce=1)
    ob = co.OracleBatch(6000, cfg, seed=4242)
    ob.reset()
    e = ob.envs
    assert not e["status"].any()
    mine = dict(num_obs=e["num_obs"], src_x=e["src"][:, 0], src_y=e["src"][:, 1], det_x=e["det"][:, 0, 0],
                det_y=e["det"][:, 0, 1], intensity=e["intensity"], bkg=e["bkg"], sp=e["best"][:, 0],
                los_blocked=e["los_blocked"][:, 0])
    cols = list(g["cols"])
    for name in ("src_x", "src_y", "det_x", "det_y", "intensity", "bkg", "sp"):
        p = stats.ks_2samp(rows[:, cols.index(name)], mine[name]).pvalue
        assert p > 1e-3, (name, p)
    # num_obs is re-drawn only every third reference reset; compare the histogram loosely
    ref_hist = np.bincount(rows[:, 0].astype(int), minlength=6)[1:] / len(rows)
    my_hist = np.bincount(mine["num_obs"], minlength=6)[1:] / ob.n
    assert np.abs(ref_hist - my_hist).max() < 0.06
    assert abs(rows[:, cols.index("los_blocked")].mean() - mine["los_blocked"].mean()) < 0.05
    # structural invariants of a valid scenario
    for i in range(0, ob.n, 7):
        k = e["num_obs"][i]
        r = e["rect"][i, :k]
        for a in range(k):
            assert 200 <= r[a, 0] < 1980 and 200 <= r[a, 2] - r[a, 0] < 500 and 200 <= r[a, 3] - r[a, 1] < 500
            for b in range(a + 1, k):
                assert r[a, 2] < r[b, 0] or r[b, 2] < r[a, 0] or r[a, 3] < r[b, 1] or r[b, 3] < r[a, 1]
        d = e["det"][i, 0].astype(np.int64) - e["src"][i].astype(np.int64)
        assert d @ d >= 1000000


@pytest.mark.parametrize("mode", [1, 2])
def test_oracle_standardizer_reproduces_reference_classes(mode):
    """oracle/radsearch_oracle.c::orc_stat_update_standardize against StatisticStandardization (RADTEAM_core.py:188-277,
    mode 1) and StatBuff + np.clip(.., -8, 8) (algos/test_environment/core.py:55-79, ppo.py:502, mode 2), recorded from
    the reference classes by tests/golden/make_golden.py: z-scores, running mean, M2 and std bit for bit in float64."""
    g = pu.load_golden("ref_standardize")
    S = len(g["length"])
    st = co.Standardizer(S, mode)
    sfx = "1" if mode == 1 else "2"
    for t in range(int(g["length"].max())):
        live = g["length"] > t
        z = st.update_standardize(g["x"][:, t], mask=live)
        np.testing.assert_array_equal(z[live], g["z" + sfx][live, t])
        np.testing.assert_array_equal(st.mean[live], g["mean" + sfx][live, t])
        np.testing.assert_array_equal(st.m2[live], g["m2_" + sfx][live, t])
        np.testing.assert_array_equal(st.std[live], g["std" + sfx][live, t])


MAPS_FILES = {"ref_maps_a1": 1, "ref_maps_a4": 4, "ref_maps_a2_idle": 2}


@pytest.mark.parametrize("name", list(MAPS_FILES))
def test_oracle_maps_reproduce_reference_mapsbuffer(name):
    """oracle/maps_oracle.c against the reference's MapsBuffer.observation_to_map (RADTEAM_core.py:532-616) run on
    oracle-env rollouts by tests/golden/make_golden.py: all seven float32 maps of every agent's buffer after every call,
    bit for bit (location / others / combined / prediction counts, median-of-samples readings through the running
    standardiser, log-scale visit counts, obstacle detections), including the bootstrap call and reset at episode ends."""
    g = pu.load_golden(name)
    A = MAPS_FILES[name]
    bufs = [co.MapsOracle(A, 120, tuple(g["dims"]), float(g["ra"])) for _ in range(A)]
    seen_multi = 0
    for t in range(len(g["obs"])):
        for i in range(A):
            got = bufs[i].observation_to_map(g["obs"][t], i, g["pred"][t, i])
            np.testing.assert_array_equal(got, g["maps"][t, i], err_msg=f"{name} call {t} agent {i}")
            assert bufs[i].status == 0
        seen_multi += int((g["maps"][t, 0, 6] > 1).any())
        if g["reset_after"][t]:
            for b in bufs:
                b.reset()
    assert g["reset_after"].sum() >= 1 and (A == 1 or seen_multi > 0)


def test_shortest_path_through_a_grazed_corner_is_the_direct_segment():
    """Source, an obstruction corner and the detector on one line: the two points are mutually visible (a grazed corner
    does not block), so `shortest_path` is the direct segment (the exact-arithmetic visilibity restatement the reference
    runs on, oracle/shims/visilibity.py), not the 1-ulp different sum over the corner.  Found by the GPU rollouts
    (env 524 of a 2048-env, 3-agent run): C oracle, kernel logic (host emulation) and the shim must agree bit for bit."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "shims"))
    try:
        import visilibity as vis
    finally:
        sys.path.pop(0)
    from tests.emu.harness import EmuEnv, make_config

    src = np.array([[437, 1730]])
    rects = np.array([[[1924, 1843, 2273, 2051], [1803, 917, 2177, 1269], [1362, 201, 1851, 700], [865, 1612, 1136, 2034]]])
    rng = np.random.default_rng(0)
    dets = [[1721, 1376]] + [[int(x), int(y)] for x, y in rng.integers(200, 2200, size=(40, 2))]
    # more collinear probes: points on the ray source -> corner (865, 1612), beyond the corner
    for k in range(2, 6):
        dets.append([437 + k * 428 // 2 * 2 // 2, 1730 - k * 118 // 1])
    polys = [vis.Polygon([vis.Point(0, 0), vis.Point(2700, 0), vis.Point(2700, 2700), vis.Point(0, 2700)])]
    for r in rects[0]:
        x0, y0, x1, y1 = (float(v) for v in r)
        polys.append(vis.Polygon([vis.Point(x0, y0), vis.Point(x0, y1), vis.Point(x1, y1), vis.Point(x1, y0)]))
    world = vis.Environment(polys)
    em = EmuEnv(1, make_config(n_agents=1, obstruction_count=4, enforce=True), seed=1)
    ob = co.OracleBatch(1, co.default_config(n_agents=1, obstruction_count=4, enforce=1))
    checked = 0
    for det in dets:
        if any(r[0] <= det[0] <= r[2] and r[1] <= det[1] <= r[3] for r in rects[0]):
            continue
        em.load_scenarios(src, np.array([det]), np.array([5000000]), np.array([20]), rects, np.array([4]))
        ob.load_scenarios(src, np.array([det]), [5000000], [20], rects, [4])
        want = world.shortest_path(vis.Point(float(src[0, 0]), float(src[0, 1])), vis.Point(float(det[0]), float(det[1])),
                                   None, 1e-7).length()
        assert ob.shortest_path(0, det) == want, det
        assert em.query_sp(np.array([det]), 0)[0] == want and em.query_sp(np.array([det]), 1)[0] == want, det
        checked += 1
    assert checked > 30


def test_shortest_path_c_oracle_vs_exact_shim_random_and_collinear():
    """orc_shortest_path and the kernel's pruned search (host emulation) against the exact-arithmetic visilibity
    restatement on random obstruction sets, with detectors drawn at random AND placed on lines through the source and
    obstruction corners (grazing / collinear cases, where a 1-ulp slip is possible): bit-exact path lengths."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "shims"))
    try:
        import visilibity as vis
    finally:
        sys.path.pop(0)
    from tests.emu.harness import EmuEnv, make_config

    rng = np.random.default_rng(123)
    checked = collinear = 0
    for case in range(10):
        K = int(rng.integers(1, 6))
        # obstructions the way create_obs draws them (R:948-1011), rejected when they touch
        rects = []
        while len(rects) < K:
            x, y = int(rng.integers(200, 1980)), int(rng.integers(200, 1980))
            w, h = int(rng.integers(200, 500)), int(rng.integers(200, 500))
            r = [x, y, x + w, y + h]
            if all(r[2] < q[0] - 1 or q[2] < r[0] - 1 or r[3] < q[1] - 1 or q[3] < r[1] - 1 for q in rects):
                rects.append(r)
        rects_a = np.zeros((1, 5, 4), np.int64); rects_a[0, :K] = rects
        inside = lambda p: any(r[0] <= p[0] <= r[2] and r[1] <= p[1] <= r[3] for r in rects)      # noqa: E731
        while True:
            src = [int(v) for v in rng.integers(200, 2200, 2)]
            if not inside(src):
                break
        dets = [[int(v) for v in rng.integers(200, 2200, 2)] for _ in range(8)]
        for r in rects:                                   # points on the ray source -> corner, beyond the corner
            for cx, cy in ((r[0], r[1]), (r[0], r[3]), (r[2], r[3]), (r[2], r[1])):
                dx, dy = cx - src[0], cy - src[1]
                g = int(np.gcd(abs(dx), abs(dy))) or 1
                for mult in (1, 2, 5):
                    p = [cx + mult * dx // g, cy + mult * dy // g]
                    if 0 <= p[0] <= 2700 and 0 <= p[1] <= 2700:
                        dets.append(p); collinear += 1
        polys = [vis.Polygon([vis.Point(0, 0), vis.Point(2700, 0), vis.Point(2700, 2700), vis.Point(0, 2700)])]
        for r in rects:
            x0, y0, x1, y1 = (float(v) for v in r)
            polys.append(vis.Polygon([vis.Point(x0, y0), vis.Point(x0, y1), vis.Point(x1, y1), vis.Point(x1, y0)]))
        world = vis.Environment(polys)
        em = EmuEnv(1, make_config(n_agents=1, obstruction_count=K, enforce=True, k_max=5), seed=1)
        ob = co.OracleBatch(1, co.default_config(n_agents=1, obstruction_count=K, enforce=1))
        for det in dets:
            if inside(det) or det == src:
                continue
            em.load_scenarios(np.array([src]), np.array([det]), np.array([5000000]), np.array([20]), rects_a, np.array([K]))
            ob.load_scenarios(np.array([src]), np.array([det]), [5000000], [20], rects_a, [K])
            want = world.shortest_path(vis.Point(float(src[0]), float(src[1])), vis.Point(float(det[0]), float(det[1])),
                                       None, 1e-7).length()
            assert ob.shortest_path(0, det) == want, (case, src, det, rects)
            assert em.query_sp(np.array([det]), 0)[0] == want, (case, src, det, rects)
            assert em.query_sp(np.array([det]), 1)[0] == want, (case, src, det, rects)
            checked += 1
    assert checked > 150 and collinear > 50

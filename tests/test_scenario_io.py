"""CPU tests of the scenario I/O (radiation_ppo_b200/scenario_io.py): the reference's test-environment dict format
(algos/test_environment/eval/test_env_gen.py:13-24, read by rad_search_env.py:799-874) both ways, and the SNR classes of
create_envs_snr (test_env_gen.py:26-97)."""
import numpy as np

from radiation_ppo_b200 import scenario_io as sio
from tests import parity_util as pu


def test_env_dict_round_trip_on_the_references_saved_scenarios(tmp_path):
    sc = pu.load_golden("scenarios_v4")
    for k in (0, 3, 7):
        arr = {key: sc[f"obs{k}_{key}"][:60] for key in ("src", "det", "intensity", "bkg", "rects", "num_obs")}
        d = sio.to_env_dict(arr)
        assert len(d) == 60 and len(d["env_0"]) == (5 if k else 4)
        if k:
            v = d["env_5"][4][0][0]                                         # [4 x 2] vertices, create_obs order
            assert v.shape == (4, 2) and (v[0] == [arr["rects"][5, 0, 0], arr["rects"][5, 0, 1]]).all()
            assert (v[2] == [arr["rects"][5, 0, 2], arr["rects"][5, 0, 3]]).all()
        path = str(tmp_path / f"test_env_dict_obs{k}")
        sio.save_test_env_dict(path, d)
        back = sio.scenario_arrays(sio.load_test_env_dict(path), k_max=7, with_obstacles=k > 0)
        for key in arr:
            np.testing.assert_array_equal(back[key], arr[key], err_msg=key)


def test_snr_classes_follow_create_envs_snr():
    rng = np.random.default_rng(0)
    n = 20000
    arr = dict(src=rng.integers(200, 2200, (n, 2)), det=rng.integers(200, 2200, (n, 2)),
               intensity=rng.integers(1_000_000, 10_000_000, n), bkg=rng.integers(10, 51, n))
    snr = sio.snr_of(arr)
    for name, (lo, hi) in (("low", (1.0, 1.2)), ("med", (1.2, 1.6)), ("high", (1.6, 2.0))):
        idx = sio.select_by_snr(arr, 100, name)
        assert len(idx) == 100 and (np.diff(idx) > 0).all()
        s = np.round(snr[idx], 3)
        assert (s > lo).all() and (s <= hi).all()
        div = np.round((hi - lo) / 4, 2)
        # the reference's loop, scenario by scenario (classify_snr, test_env_gen.py:80-97)
        counts, picked = np.zeros(4), []
        for i, v in enumerate(np.round(snr, 3)):
            if lo < v <= hi:
                for b in range(4):
                    if counts[b] < 25 and (div * b + lo) < v <= (div * (b + 1) + lo):
                        counts[b] += 1; picked.append(i); break
            if len(picked) == 100:
                break
        np.testing.assert_array_equal(idx, picked)
        assert (counts == 25).all()
    assert len(sio.select_by_snr(arr, 50, "none")) == 50

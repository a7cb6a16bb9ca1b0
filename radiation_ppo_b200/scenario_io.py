"""Scenario I/O: the reference's saved test environments -> batched arrays for rs_load_scenarios.

Format (algos/multiagent/evaluation/test_environments/test_env_dict_obs*_v4, written by
algos/test_environment/eval/test_env_gen.py:140-172, read at evaluate.py:203 and rad_search_env.py:799-874):
a joblib dict ``env_<i> -> (src xy, det xy, intensity, background[, [ [4x2 vertex array] per obstruction ])``.
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional

import numpy as np


def load_test_env_dict(path: str) -> Dict:
    import joblib

    return joblib.load(path)


def scenario_arrays(env_dict: Dict, ids: Optional[Iterable[int]] = None, k_max: int = 7,
                    with_obstacles: bool = True) -> Dict[str, np.ndarray]:
    """Pack scenarios ``ids`` (default: all, in order) into int32 arrays: src[N,2], det[N,2], intensity[N], bkg[N],
    rects[N,k_max,4] (x0,y0,x1,y1) and num_obs[N].  Coordinates must be integer-valued (they always are)."""
    if ids is None:
        ids = range(len(env_dict))
    ids = list(ids)
    n = len(ids)
    src = np.zeros((n, 2), np.int32)
    det = np.zeros((n, 2), np.int32)
    intensity = np.zeros(n, np.int32)
    bkg = np.zeros(n, np.int32)
    rects = np.zeros((n, max(k_max, 1), 4), np.int32)
    num_obs = np.zeros(n, np.int32)
    for j, i in enumerate(ids):
        e = env_dict["env_" + str(i)]
        s, d = np.asarray(e[0], np.float64), np.asarray(e[1], np.float64)
        if not (np.all(s == np.rint(s)) and np.all(d == np.rint(d))):
            raise ValueError(f"env_{i}: non-integer coordinates are outside the lattice contract")
        src[j], det[j] = s, d
        intensity[j], bkg[j] = int(e[2]), int(e[3])
        if with_obstacles and len(e) > 4:
            obs = e[4]
            if len(obs) > k_max:
                raise ValueError(f"env_{i} has {len(obs)} obstructions, k_max is {k_max}")
            num_obs[j] = len(obs)
            for k, o in enumerate(obs):
                v = np.asarray(o[0], np.float64)
                if not np.all(v == np.rint(v)):
                    raise ValueError(f"env_{i}: non-integer obstruction vertex")
                rects[j, k] = (v[:, 0].min(), v[:, 1].min(), v[:, 0].max(), v[:, 1].max())
    return dict(src=src, det=det, intensity=intensity, bkg=bkg, rects=rects, num_obs=num_obs)


def save_npz(path: str, arrays: Dict[str, np.ndarray]) -> None:
    np.savez_compressed(path, **arrays)


def load_npz(path: str) -> Dict[str, np.ndarray]:
    with np.load(path) as z:
        return {k: z[k] for k in z.files}

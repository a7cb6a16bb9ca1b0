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


def to_env_dict(arrays: Dict[str, np.ndarray]) -> Dict:
    """The inverse of scenario_arrays: batched arrays -> the reference's dict ``env_<i> -> (src xy, det xy, intensity,
    background[, obstructions])`` as test_env_gen.py:13-24 writes it (float64 coordinate arrays, Python ints, one
    ``[4x2 vertex array]`` per obstruction in the vertex order of create_obs, rad_search_env.py:975-983)."""
    n = len(arrays["intensity"])
    with_obs = "rects" in arrays and "num_obs" in arrays and int(np.max(arrays["num_obs"], initial=0)) > 0
    out = {}
    for i in range(n):
        src = np.asarray(arrays["src"][i], np.float64)
        det = np.asarray(arrays["det"][i], np.float64)
        entry = [src, det, int(arrays["intensity"][i]), int(arrays["bkg"][i])]
        if with_obs:
            obs = []
            for k in range(int(arrays["num_obs"][i])):
                x0, y0, x1, y1 = (float(v) for v in arrays["rects"][i][k])
                obs.append([np.array([[x0, y0], [x0, y1], [x1, y1], [x1, y0]], np.float64)])
            entry.append(obs)
        out["env_" + str(i)] = tuple(entry)
    return out


def save_test_env_dict(path: str, env_dict: Dict) -> None:
    """joblib file the reference's evaluate.py:203 / rad_search_env.py:799-874 read."""
    import joblib

    joblib.dump(env_dict, path)


def snr_of(arrays: Dict[str, np.ndarray]) -> np.ndarray:
    """Expected signal-to-noise ratio at the start position as test_env_gen.py:36-47 defines it:
    (intensity / d^2 + background) / background with d the source-detector distance."""
    d = np.asarray(arrays["src"], np.float64) - np.asarray(arrays["det"], np.float64)
    det2 = (d ** 2).sum(axis=1)
    bkg = np.asarray(arrays["bkg"], np.float64)
    return (np.asarray(arrays["intensity"], np.float64) / det2 + bkg) / bkg


SNR_RANGES = {"none": (0.0, 0.0), "low": (1.0, 1.2), "med": (1.2, 1.6), "high": (1.6, 2.0)}     # test_env_gen.py:30


def select_by_snr(arrays: Dict[str, np.ndarray], num_envs: int, snr: str = "high", split: int = 4) -> np.ndarray:
    """Indices (in sampling order) of the first scenarios that fill the reference's SNR classes: the range of `snr` cut
    into `split` equal bins (upper edges inclusive, bin width rounded to 2 decimals) with round(num_envs / split)
    scenarios each (create_envs_snr + classify_snr, test_env_gen.py:26-97).  'none' takes the first num_envs."""
    if snr == "none":
        return np.arange(min(num_envs, len(arrays["intensity"])))
    lo, hi = SNR_RANGES[snr]
    div = np.round((hi - lo) / split, 2)
    per_bin = round(num_envs / split)
    s = np.round(snr_of(arrays), 3)
    counts = np.zeros(split, np.int64)
    picked = []
    for i, v in enumerate(s):
        if not (lo < v <= hi):
            continue
        for b in range(split):
            if counts[b] < per_bin and (lo + div * b) < v <= (lo + div * (b + 1)):
                counts[b] += 1
                picked.append(i)
                break
        if len(picked) >= num_envs:
            break
    return np.asarray(picked, np.int64)


def save_npz(path: str, arrays: Dict[str, np.ndarray]) -> None:
    np.savez_compressed(path, **arrays)


def load_npz(path: str) -> Dict[str, np.ndarray]:
    with np.load(path) as z:
        return {k: z[k] for k in z.files}

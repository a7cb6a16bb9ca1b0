"""Measure what "geometry parity unpinned" amounts to.

The oracle evaluates VisiLibity1's predicates in exact rationals (oracle/shims/visilibity.py); the real library evaluates
the same formulas in doubles (restated in oracle/shims/visilibity_f64.py).  This script draws RadSearch-shaped cases on
the integer lattice the environment lives on -- a rectangle from the reference's generator (R:960-982), a detector
position biased towards the degenerate places (on an edge, on a corner, on the extension of an edge, one unit off), a
source, a second rectangle -- and evaluates every call site of the dependency in both modes:

  in          Point::in(poly, 1e-7)                                        R:1061, 1108, 1155, 1291
  los         boundary_distance(Line_Segment(det, src), poly) < 0.001      R:1110, 1141
  sens_hit    intersect(edge, ray, 1e-7), 8 rays x 4 edges                 R:1205
  sens_dist   distance(det, edge) where the ray hits                       R:1207
  sens_vec    the 8 proximity values the env derives, (110-d)/110 and the max over edges   R:1196-1216
  sens_fire   the "more than three sensors == 1.0" trigger of correct_coords               R:1219
  rect_rect   isclose(boundary_distance(poly, poly), 0, abs_tol=1e-7)      R:988

and counts disagreements per class.  Output: one JSON object (profiles/r02_geometry_f64_vs_exact.json).

R: = /root/reference/gym_rad_search/gym_rad_search/envs/rad_search_env.py.  Runs on CPU only, no reference import.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
from multiprocessing import Pool

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.shims import visilibity as ex  # noqa: E402
from oracle.shims import visilibity_f64 as fl  # noqa: E402

EPS = 1e-7
DIST_TH = 110.0
STEPS = [(-100, 0), (-71, 71), (0, 100), (71, 71), (100, 0), (71, -71), (0, -100), (-71, -71)]  # R:178-224


def rect_poly(x0, y0, w, h):
    """Vertex order of R:975-982 and the edge list of R:997-1006."""
    v = [ex.Point(x0, y0), ex.Point(x0, y0 + h), ex.Point(x0 + w, y0 + h), ex.Point(x0 + w, y0)]
    poly = ex.Polygon(v)
    segs = [ex.Line_Segment(v[0], v[1]), ex.Line_Segment(v[0], v[3]), ex.Line_Segment(v[2], v[1]), ex.Line_Segment(v[2], v[3])]
    return poly, segs


def sensors(mod, det, segs):
    """R:1196-1216 for one obstruction; returns (values, hit flags, raw distances)."""
    dists = [0.0] * 8
    hits, raws = [], []
    for d, (dx, dy) in enumerate(STEPS):
        ray = ex.Line_Segment(det, ex.Point(det._x + dx, det._y + dy))
        inter = 0
        seg_dist = [0.0] * 4
        for k, edge in enumerate(segs):
            h = mod.intersect(edge, ray, EPS)
            hits.append(h)
            if inter < 2 and h:
                od = mod.distance(det, edge)
                raws.append((d, k, od))
                seg_dist[k] = (DIST_TH - od) / DIST_TH
                inter += 1
        if inter > 0:
            dists[d] = max(dists[d], max(seg_dist))
    return dists, hits, raws


def ulps(a: float, b: float) -> int:
    if a == b:
        return 0
    ia = np.float64(a).view(np.int64)
    ib = np.float64(b).view(np.int64)
    return int(abs(int(ia) - int(ib)))


def draw_point(rng, x0, y0, w, h):
    """Half the draws land on the degenerate places; the rest anywhere within sensor range of the rectangle."""
    r = rng.random()
    if r < 0.5:
        xs = [x0, x0 + w, x0 - 1, x0 + 1, x0 + w - 1, x0 + w + 1, int(rng.integers(x0 - 120, x0 + w + 121))]
        ys = [y0, y0 + h, y0 - 1, y0 + 1, y0 + h - 1, y0 + h + 1, int(rng.integers(y0 - 120, y0 + h + 121))]
        if rng.random() < 0.5:
            return xs[int(rng.integers(0, 6))], ys[int(rng.integers(0, 7))]
        return xs[int(rng.integers(0, 7))], ys[int(rng.integers(0, 6))]
    return int(rng.integers(x0 - 120, x0 + w + 121)), int(rng.integers(y0 - 120, y0 + h + 121))


def chunk(args):
    seed, n = args
    rng = np.random.default_rng(seed)
    c = dict(
        scenes=0, in_cases=0, in_diff=0, los_cases=0, los_diff=0, los_value_diff=0, sens_hit_cases=0, sens_hit_diff=0,
        sens_dist_cases=0, sens_dist_diff=0, sens_dist_max_ulps=0, sens_dist_zero_cases=0, sens_dist_zero_nonzero=0,
        sens_dist_max_abs=0.0, sens_vec_max_abs=0.0, sens_vec_cases=0, sens_vec_diff=0, sens_vec_diff_on_boundary=0, sens_fire_exact=0, sens_fire_f64=0,
        sens_fire_diff=0, on_boundary_scenes=0, rect_rect_cases=0, rect_rect_diff=0,
    )
    for _ in range(n):
        x0 = int(rng.integers(200, 2200))
        y0 = int(rng.integers(200, 2200))
        w = int(rng.integers(200, 500))
        h = int(rng.integers(200, 500))
        poly, segs = rect_poly(x0, y0, w, h)
        px, py = draw_point(rng, x0, y0, w, h)
        det = ex.Point(px, py)
        c["scenes"] += 1
        on_b = (px in (x0, x0 + w) and y0 <= py <= y0 + h) or (py in (y0, y0 + h) and x0 <= px <= x0 + w)
        c["on_boundary_scenes"] += on_b
        # in
        a = det._in(poly, EPS)
        b = fl.point_in(det, poly, EPS)
        c["in_cases"] += 1
        c["in_diff"] += a != b
        # line of sight: half the sources are placed so that the segment grazes a corner or runs along an edge
        if rng.random() < 0.5:
            cx, cy = [(x0, y0), (x0, y0 + h), (x0 + w, y0 + h), (x0 + w, y0)][int(rng.integers(0, 4))]
            k = int(rng.integers(1, 4))
            sx, sy = cx + k * (cx - px), cy + k * (cy - py)  # det, corner, src collinear
        else:
            sx, sy = int(rng.integers(0, 2700)), int(rng.integers(0, 2700))
        L = ex.Line_Segment(det, ex.Point(sx, sy))
        da, db = ex.boundary_distance(L, poly), fl.boundary_distance(L, poly)
        c["los_cases"] += 1
        c["los_diff"] += (da < 0.001) != (db < 0.001)
        c["los_value_diff"] += da != db
        # sensors
        va, ha, ra = sensors(ex, det, segs)
        vb, hb, rb = sensors(fl, det, segs)
        c["sens_hit_cases"] += len(ha)
        c["sens_hit_diff"] += sum(x != y for x, y in zip(ha, hb))
        if [r[:2] for r in ra] == [r[:2] for r in rb]:
            for (_, _, x), (_, _, y) in zip(ra, rb):
                c["sens_dist_cases"] += 1
                if x != y:
                    c["sens_dist_diff"] += 1
                    c["sens_dist_max_ulps"] = max(c["sens_dist_max_ulps"], ulps(x, y)) if x != 0.0 else c["sens_dist_max_ulps"]
                    c["sens_dist_max_abs"] = max(c["sens_dist_max_abs"], abs(x - y))
                if x == 0.0:
                    c["sens_dist_zero_cases"] += 1
                    c["sens_dist_zero_nonzero"] += y != 0.0
        c["sens_vec_cases"] += 1
        if va != vb:
            c["sens_vec_diff"] += 1
            c["sens_vec_diff_on_boundary"] += on_b
            c["sens_vec_max_abs"] = max(c["sens_vec_max_abs"], max(abs(x - y) for x, y in zip(va, vb)))
        fa = sum(1 for v in va if v == 1.0) > 3
        fb = sum(1 for v in vb if v == 1.0) > 3
        c["sens_fire_exact"] += fa
        c["sens_fire_f64"] += fb
        c["sens_fire_diff"] += fa != fb
        # a second rectangle, half the time sharing an edge line / a corner with the first
        w2, h2 = int(rng.integers(200, 500)), int(rng.integers(200, 500))
        if rng.random() < 0.5:
            x2 = [x0 + w, x0 - w2, x0 + w + 1, int(rng.integers(x0 - w2, x0 + w))][int(rng.integers(0, 4))]
            y2 = [y0 + h, y0 - h2, y0 + h + 1, int(rng.integers(y0 - h2, y0 + h))][int(rng.integers(0, 4))]
        else:
            x2, y2 = int(rng.integers(x0 - 600, x0 + 600)), int(rng.integers(y0 - 600, y0 + 600))
        poly2, _ = rect_poly(x2, y2, w2, h2)
        ga = math.isclose(ex.boundary_distance(poly, poly2), 0.0, abs_tol=EPS)
        gb = math.isclose(fl.boundary_distance(poly, poly2), 0.0, abs_tol=EPS)
        c["rect_rect_cases"] += 1
        c["rect_rect_diff"] += ga != gb
    return c


def run(scenes: int, procs: int, seed: int = 20260101):
    per = 2000
    jobs = [(seed + i, min(per, scenes - i * per)) for i in range((scenes + per - 1) // per)]
    if procs > 1:
        with Pool(procs) as pool:
            parts = pool.map(chunk, jobs)
    else:
        parts = [chunk(j) for j in jobs]
    tot = dict(parts[0])
    for p in parts[1:]:
        for k, v in p.items():
            tot[k] = max(tot[k], v) if "_max_" in k else tot[k] + v
    return {k: (int(v) if not isinstance(v, float) else v) for k, v in tot.items()}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenes", type=int, default=1_000_000)
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    res = run(a.scenes, a.procs)
    res["note"] = (
        "exact = oracle/shims/visilibity.py (rationals), f64 = oracle/shims/visilibity_f64.py (VisiLibity1's double formulas); "
        "integer-lattice cases, half of them on the degenerate places (detector on an edge / corner / edge extension, line of "
        "sight through a corner, rectangles sharing an edge line)"
    )
    s = json.dumps(res, indent=1)
    print(s)
    if a.out:
        open(a.out, "w").write(s + "\n")

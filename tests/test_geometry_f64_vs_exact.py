"""The exact-rational restatement of the un-vendored `visilibity` dependency against the double-precision restatement of the
same VisiLibity1 formulas (oracle/shims/visilibity_f64.py): every DECISION the environment takes from the library is the
same in both; only the proximity-sensor distance carries the library's projection noise (measured
on 1e6 cases in profiles/r02_geometry_f64_vs_exact.json)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))
import geometry_f64_vs_exact as g  # noqa: E402


def test_double_precision_formulas_take_the_same_decisions_as_the_exact_ones():
    r = g.run(scenes=600, procs=1, seed=7)
    assert r["scenes"] == 600 and r["sens_hit_cases"] == 600 * 32
    # containment, line of sight, ray/edge hits, rectangle separation: identical
    assert r["in_diff"] == 0 and r["los_diff"] == 0 and r["sens_hit_diff"] == 0 and r["rect_rect_diff"] == 0
    # the sensor distance differs only by rounding noise of the projection (<= 1e-12 cm on coordinates of a few thousand)
    assert r["sens_dist_max_abs"] < 1e-11 and r["sens_vec_max_abs"] < 1e-13
    # the one observable class: a detector exactly on an edge, where the noise decides whether four sensors read 1.0
    assert r["sens_fire_diff"] <= r["on_boundary_scenes"]
    assert r["sens_fire_f64"] <= r["sens_fire_exact"] == r["on_boundary_scenes"]

"""TEST INFRASTRUCTURE ONLY -- the same `visilibity` subset as oracle/shims/visilibity.py, evaluated the way VisiLibity1
itself evaluates it: in IEEE doubles, formula by formula.

oracle/shims/visilibity.py evaluates VisiLibity1's published predicates in exact rationals and is what the C oracle and the
CUDA path are pinned on.  The real library (peproctor/PyVisiLibity @ c7602007, un-vendored, not buildable here) runs the
same formulas in double precision.  This module restates those double-precision formulas so that the gap between "exact"
and "as the library rounds" can be MEASURED instead of argued (tools/geometry_f64_vs_exact.py):

  Point::projection_onto(Line_Segment)   theta = ((s.x-p.x)(s.x-f.x) + (s.y-p.y)(s.y-f.y)) / (pow(s.x-f.x,2)+pow(s.y-f.y,2));
                                         0 <= theta <= 1 -> theta*first + (1-theta)*second, else the closer endpoint
  distance(Point, Point)                 sqrt(pow(dx,2) + pow(dy,2))
  distance(Point, Line_Segment)          distance(p, p.projection_onto(seg))
  intersect_proper(seg, seg, eps=0)      min endpoint/segment distance <= eps -> false; else O'Rourke's two cross-product signs
  distance(seg, seg)                     0 if intersect_proper, else the min of the 4 endpoint/segment distances
  intersect(seg, seg, eps)               distance(seg, seg) <= eps
  boundary_distance(...)                 min over polygon edges
  Point::in(Polygon, eps)                boundary_distance(p, poly) <= eps, else the pnpoly crossing number

`shortest_path` / `is_valid` are inherited from the exact module: the library builds its visibility graph with its
eps-robust visibility-polygon sweep, which is not restated here (the one part of the dependency that stays argued, not
measured: DESIGN.md section 4).

Only tests/ and tools/geometry_f64_vs_exact.py import this file.
"""
from __future__ import annotations

import math

from . import visilibity as _exact
from .visilibity import Bounding_Box, Environment, Line_Segment, Point, Polygon, Polyline, Visibility_Graph  # noqa: F401


def _pp(ax: float, ay: float, bx: float, by: float) -> float:
    return math.sqrt(math.pow(ax - bx, 2) + math.pow(ay - by, 2))


def _projection(px: float, py: float, fx: float, fy: float, sx: float, sy: float):
    """Point::projection_onto(Line_Segment(first=f, second=s))."""
    if fx == sx and fy == sy:
        return fx, fy
    theta = ((sx - px) * (sx - fx) + (sy - py) * (sy - fy)) / (math.pow(sx - fx, 2) + math.pow(sy - fy, 2))
    if 0.0 <= theta <= 1.0:
        return theta * fx + (1.0 - theta) * sx, theta * fy + (1.0 - theta) * sy
    if _pp(px, py, fx, fy) < _pp(px, py, sx, sy):
        return fx, fy
    return sx, sy


def _pt_seg(p: Point, a: Point, b: Point) -> float:
    qx, qy = _projection(p._x, p._y, a._x, a._y, b._x, b._y)
    return _pp(p._x, p._y, qx, qy)


def _cross(ax, ay, bx, by):
    return ax * by - bx * ay


def _proper(a: Point, b: Point, c: Point, d: Point, eps: float = 0.0) -> bool:
    m = min(_pt_seg(a, c, d), _pt_seg(b, c, d), _pt_seg(c, a, b), _pt_seg(d, a, b))
    if m <= eps:
        return False
    s1 = _cross(b._x - a._x, b._y - a._y, c._x - b._x, c._y - b._y) * _cross(b._x - a._x, b._y - a._y, d._x - b._x, d._y - b._y)
    s2 = _cross(d._x - c._x, d._y - c._y, b._x - d._x, b._y - d._y) * _cross(d._x - c._x, d._y - c._y, a._x - d._x, a._y - d._y)
    return s1 < 0 and s2 < 0


def _seg_seg(a: Point, b: Point, c: Point, d: Point) -> float:
    if _proper(a, b, c, d):
        return 0.0
    return min(_pt_seg(a, c, d), _pt_seg(b, c, d), _pt_seg(c, a, b), _pt_seg(d, a, b))


def distance(p, s) -> float:
    if isinstance(s, Line_Segment):
        return _pt_seg(p, s.a, s.b)
    if isinstance(s, Point):
        return _pp(p._x, p._y, s._x, s._y)
    raise TypeError("unsupported distance() operands")


def intersect(s1: Line_Segment, s2: Line_Segment, epsilon: float = 0.0) -> bool:
    return _seg_seg(s1.a, s1.b, s2.a, s2.b) <= epsilon


def boundary_distance(a, b) -> float:
    if isinstance(a, Polygon) and isinstance(b, Polygon):
        return min(_seg_seg(p, q, r, s) for p, q in a.edges() for r, s in b.edges())
    if isinstance(a, Polygon):
        a, b = b, a
    if isinstance(a, Line_Segment) and isinstance(b, Polygon):
        return min(_seg_seg(a.a, a.b, r, s) for r, s in b.edges())
    if isinstance(a, Point) and isinstance(b, Polygon):
        return min(_pt_seg(a, r, s) for r, s in b.edges())
    raise TypeError("unsupported boundary_distance() operands")


def point_in(p: Point, poly: Polygon, epsilon: float = 0.0) -> bool:
    """Point::in(Polygon, eps) in doubles (Point._in of the exact module is the rational version)."""
    n = poly.n()
    if n < 3:
        return False
    if boundary_distance(p, poly) <= epsilon:
        return True
    c = False
    j = n - 1
    for i in range(n):
        pi, pj = poly[i], poly[j]
        if ((pi._y <= p._y < pj._y) or (pj._y <= p._y < pi._y)) and (
            p._x < (pj._x - pi._x) * (p._y - pi._y) / (pj._y - pi._y) + pi._x
        ):
            c = not c
        j = i
    return c


exact = _exact

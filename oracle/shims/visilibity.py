"""TEST INFRASTRUCTURE ONLY -- exact-arithmetic restatement of the `visilibity` subset RadSearch uses.

The reference pins `visilibity` = peproctor/PyVisiLibity @ c76020079110231f882f38f61b3ab25d01de21f0
(/root/reference/gym_rad_search/setup.py:11), a SWIG wrap of K. Obermeyer's VisiLibity1.  Its source is not in
/root/reference and cannot be installed here, so this module restates the *published algorithms* of the calls made
from /root/reference/gym_rad_search/gym_rad_search/envs/rad_search_env.py (call sites listed per function below)
with one deliberate difference: every predicate is evaluated in exact rational arithmetic on the exact values of the
float inputs, and every returned distance is the correctly rounded exact distance.  VisiLibity itself evaluates the
same formulas in double precision; on the integer lattice RadSearch lives on, the two agree except where the
library's float projection `theta*first + (1-theta)*second` leaves ~1e-13 of noise on a point that is exactly on an
edge (see DESIGN.md "parity unpinned").

PARITY UNPINNED: no reference test exercises these calls and the real library cannot be run here.

Only tests/, tools that generate tests/golden/, and bench.py's cpu_baseline leg may import this file.
"""
from __future__ import annotations

import heapq
import math
from fractions import Fraction as Fr
from typing import Iterable, List, Sequence


def _fr(v) -> Fr:
    return Fr(float(v))


def _sqrt_fr(q: Fr) -> float:
    """Correctly rounded float sqrt of a non-negative rational."""
    if q == 0:
        return 0.0
    # scale to >= 2*64 fractional bits, integer sqrt, then one correctly rounded int/int division
    s = 128
    n = (q.numerator << (2 * s)) // q.denominator
    r = math.isqrt(n)
    return r / (1 << s)


class Point:
    """VisiLibity::Point (x(), y(), in()).  Call sites: rad_search_env.py:139-150, 1061, 1108, 1155, 1291."""

    __slots__ = ("_x", "_y", "fx", "fy")

    def __init__(self, x=0.0, y=0.0):
        self._x = float(x)
        self._y = float(y)
        self.fx = Fr(self._x)
        self.fy = Fr(self._y)

    def x(self) -> float:
        return self._x

    def y(self) -> float:
        return self._y

    def _in(self, poly: "Polygon", epsilon: float = 0.0) -> bool:
        """Point::in(Polygon, eps): within eps of the boundary, or odd crossing number (closed containment)."""
        if poly.n() < 3:
            return False
        if _pt_poly_boundary_dist2(self, poly) < _fr(epsilon) ** 2:
            return True
        c = False
        n = poly.n()
        j = n - 1
        for i in range(n):
            pi, pj = poly[i], poly[j]
            if ((pi.fy <= self.fy < pj.fy) or (pj.fy <= self.fy < pi.fy)) and (
                self.fx < (pj.fx - pi.fx) * (self.fy - pi.fy) / (pj.fy - pi.fy) + pi.fx
            ):
                c = not c
            j = i
        return c

    def __repr__(self):
        return f"vis.Point({self._x}, {self._y})"


class Bounding_Box:
    __slots__ = ("x_min", "x_max", "y_min", "y_max")


class Polygon:
    """VisiLibity::Polygon (vertex list, bbox(), n()).  Call sites: rad_search_env.py:160-164, 1160."""

    def __init__(self, pts: Iterable[Point] = ()):
        self.v: List[Point] = list(pts)

    def n(self) -> int:
        return len(self.v)

    def __getitem__(self, i: int) -> Point:
        return self.v[i % len(self.v)]

    def __iter__(self):
        return iter(self.v)

    def __len__(self):
        return len(self.v)

    def bbox(self) -> Bounding_Box:
        b = Bounding_Box()
        b.x_min = min(p.x() for p in self.v)
        b.x_max = max(p.x() for p in self.v)
        b.y_min = min(p.y() for p in self.v)
        b.y_max = max(p.y() for p in self.v)
        return b

    def area2(self) -> Fr:
        """Twice the signed area (positive = counter-clockwise)."""
        a = Fr(0)
        for i in range(self.n()):
            p, q = self[i], self[i + 1]
            a += p.fx * q.fy - q.fx * p.fy
        return a

    def edges(self):
        for i in range(self.n()):
            yield self[i], self[i + 1]


class Line_Segment:
    """VisiLibity::Line_Segment (first(), second()).  Call sites: rad_search_env.py:997-1006, 1105, 1139, 1182."""

    __slots__ = ("a", "b")

    def __init__(self, a: Point, b: Point):
        self.a = a
        self.b = b

    def first(self) -> Point:
        return self.a

    def second(self) -> Point:
        return self.b


# --------------------------------------------------------------------------------------------------------------
# exact kernels
# --------------------------------------------------------------------------------------------------------------
def _cross(ax, ay, bx, by):
    return ax * by - ay * bx


def _pt_seg_dist2(p: Point, a: Point, b: Point) -> Fr:
    """Squared distance from p to the closed segment ab (distance to the clamped projection)."""
    dx, dy = b.fx - a.fx, b.fy - a.fy
    l2 = dx * dx + dy * dy
    wx, wy = p.fx - a.fx, p.fy - a.fy
    if l2 == 0:
        return wx * wx + wy * wy
    t = wx * dx + wy * dy
    if t <= 0:
        return wx * wx + wy * wy
    if t >= l2:
        ux, uy = p.fx - b.fx, p.fy - b.fy
        return ux * ux + uy * uy
    c = _cross(wx, wy, dx, dy)
    return c * c / l2


def _proper(a: Point, b: Point, c: Point, d: Point) -> bool:
    """intersect_proper(ab, cd): no endpoint touches the other segment, and the O'Rourke left/right-turn test holds."""
    if (
        _pt_seg_dist2(a, c, d) == 0
        or _pt_seg_dist2(b, c, d) == 0
        or _pt_seg_dist2(c, a, b) == 0
        or _pt_seg_dist2(d, a, b) == 0
    ):
        return False
    abx, aby = b.fx - a.fx, b.fy - a.fy
    cdx, cdy = d.fx - c.fx, d.fy - c.fy
    s1 = _cross(abx, aby, c.fx - b.fx, c.fy - b.fy) * _cross(abx, aby, d.fx - b.fx, d.fy - b.fy)
    s2 = _cross(cdx, cdy, b.fx - d.fx, b.fy - d.fy) * _cross(cdx, cdy, a.fx - d.fx, a.fy - d.fy)
    return s1 < 0 and s2 < 0


def _seg_seg_dist2(a: Point, b: Point, c: Point, d: Point) -> Fr:
    """distance(Line_Segment, Line_Segment)^2: 0 if properly crossing, else min of the 4 endpoint distances."""
    if _proper(a, b, c, d):
        return Fr(0)
    return min(_pt_seg_dist2(a, c, d), _pt_seg_dist2(b, c, d), _pt_seg_dist2(c, a, b), _pt_seg_dist2(d, a, b))


def _pt_poly_boundary_dist2(p: Point, poly: Polygon) -> Fr:
    return min(_pt_seg_dist2(p, a, b) for a, b in poly.edges())


# --------------------------------------------------------------------------------------------------------------
# module-level functions used by the env
# --------------------------------------------------------------------------------------------------------------
def distance(p, s) -> float:
    """vis.distance(Point, Line_Segment) (rad_search_env.py:1207) / (Point, Point)."""
    if isinstance(s, Line_Segment):
        return _sqrt_fr(_pt_seg_dist2(p, s.a, s.b))
    if isinstance(s, Point):
        return _sqrt_fr((p.fx - s.fx) ** 2 + (p.fy - s.fy) ** 2)
    raise TypeError("unsupported distance() operands")


def intersect(s1: Line_Segment, s2: Line_Segment, epsilon: float = 0.0) -> bool:
    """vis.intersect(seg, seg, eps) = distance(seg, seg) <= eps (rad_search_env.py:1205)."""
    return _seg_seg_dist2(s1.a, s1.b, s2.a, s2.b) <= _fr(epsilon) ** 2


def boundary_distance(a, b) -> float:
    """vis.boundary_distance(Line_Segment, Polygon) (rad_search_env.py:1110, 1141) and (Polygon, Polygon) (:988)."""
    if isinstance(a, Polygon) and isinstance(b, Polygon):
        return _sqrt_fr(min(_seg_seg_dist2(p, q, r, s) for p, q in a.edges() for r, s in b.edges()))
    if isinstance(a, Polygon):
        a, b = b, a
    if isinstance(a, Line_Segment) and isinstance(b, Polygon):
        return _sqrt_fr(min(_seg_seg_dist2(a.a, a.b, r, s) for r, s in b.edges()))
    if isinstance(a, Point) and isinstance(b, Polygon):
        return _sqrt_fr(_pt_poly_boundary_dist2(a, b))
    raise TypeError("unsupported boundary_distance() operands")


class Polyline:
    def __init__(self, pts: Sequence[Point]):
        self.pts = list(pts)

    def size(self) -> int:
        return len(self.pts)

    def length(self) -> float:
        """Polyline::length(): left-to-right double sum of consecutive vertex distances."""
        s = 0.0
        for p, q in zip(self.pts[:-1], self.pts[1:]):
            s += distance(p, q)
        return s


def _rect_of(poly: Polygon):
    xs = sorted({p.fx for p in poly})
    ys = sorted({p.fy for p in poly})
    if poly.n() != 4 or len(xs) != 2 or len(ys) != 2:
        raise NotImplementedError("shortest_path restatement supports axis-aligned rectangular holes only")
    return xs[0], ys[0], xs[1], ys[1]


def _seg_hits_open_rect(p: Point, q: Point, r) -> bool:
    """True iff the open segment pq meets the open rectangle r (grazing an edge or corner is not a hit)."""
    x0, y0, x1, y1 = r
    lo, hi = Fr(0), Fr(1)
    for (s, d, a, b) in ((p.fx, q.fx - p.fx, x0, x1), (p.fy, q.fy - p.fy, y0, y1)):
        if d == 0:
            if not (a < s < b):
                return False
        else:
            t0, t1 = (a - s) / d, (b - s) / d
            if t0 > t1:
                t0, t1 = t1, t0
            lo, hi = max(lo, t0), min(hi, t1)
    return lo < hi


class Environment:
    """VisiLibity::Environment: outer boundary + holes.  Call sites: rad_search_env.py:757, 788, 857, 491, 774, 866."""

    def __init__(self, polys: Sequence[Polygon]):
        polys = list(polys)
        self.outer = polys[0]
        self.holes = polys[1:]

    def h(self) -> int:
        return len(self.holes)

    def is_valid(self, epsilon: float = 0.0) -> bool:
        """Simple polygons, boundaries pairwise > eps apart, hole vertices inside the outer boundary and outside every
        other hole, outer CCW and holes CW."""
        eps2 = _fr(epsilon) ** 2
        if self.outer.n() + sum(h.n() for h in self.holes) <= 2:
            return False
        for poly in [self.outer, *self.holes]:
            if not _is_simple(poly, eps2):
                return False
        for hpoly in self.holes:
            if min(_seg_seg_dist2(p, q, r, s) for p, q in self.outer.edges() for r, s in hpoly.edges()) <= eps2:
                return False
        for i in range(self.h()):
            for j in range(i + 1, self.h()):
                if min(
                    _seg_seg_dist2(p, q, r, s) for p, q in self.holes[i].edges() for r, s in self.holes[j].edges()
                ) <= eps2:
                    return False
        for i, hpoly in enumerate(self.holes):
            for v in hpoly:
                if not v._in(self.outer, epsilon):
                    return False
                for k, other in enumerate(self.holes):
                    if k != i and v._in(other, epsilon):
                        return False
        if self.outer.area2() <= 0:
            return False
        for hpoly in self.holes:
            if hpoly.area2() >= 0:
                return False
        return True

    def shortest_path(self, start: Point, finish: Point, graph=None, epsilon: float = 0.0) -> Polyline:
        """Euclidean shortest path from start to finish around the holes: the direct segment when the two points are
        mutually visible, otherwise the optimum over the visibility graph on {start, finish, hole vertices}
        (VisiLibity runs A*; any exact search returns the same optimum).  The outer boundary is convex in RadSearch
        and is ignored."""
        if (start.fx - finish.fx) ** 2 + (start.fy - finish.fy) ** 2 <= _fr(epsilon) ** 2:
            return Polyline([start])
        rects = [_rect_of(h) for h in self.holes]

        def visible(p, q):
            return not any(_seg_hits_open_rect(p, q, r) for r in rects)

        if visible(start, finish):
            return Polyline([start, finish])
        nodes = [start] + [v for h in self.holes for v in h] + [finish]
        m = len(nodes)
        dist = [math.inf] * m
        prev = [-1] * m
        dist[0] = 0.0
        heap = [(0.0, 0)]
        done = [False] * m
        while heap:
            d, u = heapq.heappop(heap)
            if done[u]:
                continue
            done[u] = True
            if u == m - 1:
                break
            for w in range(1, m):
                if done[w] or w == u:
                    continue
                if visible(nodes[u], nodes[w]):
                    nd = d + distance(nodes[u], nodes[w])
                    if nd < dist[w]:
                        dist[w] = nd
                        prev[w] = u
                        heapq.heappush(heap, (nd, w))
        path = []
        u = m - 1
        while u != -1:
            path.append(nodes[u])
            u = prev[u]
        return Polyline(path[::-1])


def _is_simple(poly: Polygon, eps2: Fr) -> bool:
    n = poly.n()
    if n < 3:
        return False
    for i in range(n):
        for j in range(i + 2, n):
            if i == 0 and j == n - 1:
                continue
            if _seg_seg_dist2(poly[i], poly[i + 1], poly[j], poly[j + 1]) <= eps2:
                return False
    return True


class Visibility_Graph:
    """Placeholder: the restated shortest_path evaluates visibility on demand (rad_search_env.py:760, 858)."""

    def __init__(self, env: Environment = None, epsilon: float = 0.0):
        self.env = env

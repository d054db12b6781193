"""Restatement of Clipper 6.4.2's ClipperOffset for closed polygons -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference un-clips every candidate polygon with pyclipper (src/det/test.py:37-43):

    offset = pyclipper.PyclipperOffset()                       # MiterLimit 2.0, ArcTolerance 0.25
    offset.AddPath(box, pyclipper.JT_ROUND, pyclipper.ET_CLOSEDPOLYGON)
    expanded = np.array(offset.Execute(distance))

pyclipper (a Cython wrapper of Angus Johnson's Clipper 6.4.2, clipper.cpp) is a third-party dependency that is
neither vendored in /root/reference nor installable here, so its published algorithm is restated:
ClipperOffset::AddPath / FixOrientations / DoOffset / OffsetPoint / DoRound and the Round() helper, integer
coordinates, join type jtRound, end type etClosedPolygon.

NOT restated: the final `Clipper::Execute(ctUnion, pftPositive)` that removes the self-intersections a raw offset
path has at concave vertices (and re-orders / de-duplicates vertices).  The union keeps every extreme vertex of the
raw path, so `cv2.boundingRect` of the result -- the only thing the reference's consumer takes from a box
(crop_image, src/det/test.py:123-130) -- is unchanged; the vertex list of a concave input may differ.

`PyclipperOffset`, `JT_ROUND`, `ET_CLOSEDPOLYGON` mirror the pyclipper names so that the reference's own
`unclip()` runs unmodified on top of this module (tests/golden/make_postproc_reference.py).
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

JT_SQUARE, JT_ROUND, JT_MITER = 0, 1, 2
ET_CLOSEDPOLYGON, ET_CLOSEDLINE, ET_OPENBUTT, ET_OPENSQUARE, ET_OPENROUND = 0, 1, 2, 3, 4

_TWO_PI = 2.0 * math.pi
_DEF_ARC_TOLERANCE = 0.25


def _round(v: float) -> int:
    """clipper.cpp Round(): half away from zero, truncating cast."""
    return int(v - 0.5) if v < 0 else int(v + 0.5)


def _area(path: Sequence[Tuple[int, int]]) -> float:
    """clipper.cpp Area(const Path&): a = sum (x_j + x_i) * (y_j - y_i), returns -a / 2."""
    n = len(path)
    if n < 3:
        return 0.0
    a = 0.0
    j = n - 1
    for i in range(n):
        a += (float(path[j][0]) + path[i][0]) * (float(path[j][1]) - path[i][1])
        j = i
    return -a * 0.5


def _unit_normal(p1, p2) -> Tuple[float, float]:
    if p2[0] == p1[0] and p2[1] == p1[1]:
        return 0.0, 0.0
    dx, dy = float(p2[0] - p1[0]), float(p2[1] - p1[1])
    f = 1.0 / math.sqrt(dx * dx + dy * dy)
    return dy * f, -dx * f


class PyclipperOffset:
    def __init__(self, miter_limit: float = 2.0, arc_tolerance: float = 0.25):
        self.MiterLimit = miter_limit
        self.ArcTolerance = arc_tolerance
        self._paths: List[Tuple[List[Tuple[int, int]], int, int]] = []
        self._lowest = None  # (path index, vertex index) of the lowest (max y, then min x) closed-polygon vertex

    def AddPath(self, path, join_type: int, end_type: int) -> None:
        pts = [(int(p[0]), int(p[1])) for p in path]
        high = len(pts) - 1
        if high < 0:
            return
        if end_type in (ET_CLOSEDLINE, ET_CLOSEDPOLYGON):  # strip a duplicated closing vertex
            while high > 0 and pts[0] == pts[high]:
                high -= 1
        contour = [pts[0]]
        j, k = 0, 0
        for i in range(1, high + 1):  # drop consecutive duplicates, remember the lowest vertex
            if contour[j] != pts[i]:
                j += 1
                contour.append(pts[i])
                if pts[i][1] > contour[k][1] or (pts[i][1] == contour[k][1] and pts[i][0] < contour[k][0]):
                    k = j
        if end_type == ET_CLOSEDPOLYGON and j < 2:
            return
        self._paths.append((contour, join_type, end_type))
        if end_type != ET_CLOSEDPOLYGON:
            return
        idx = len(self._paths) - 1
        if self._lowest is None:
            self._lowest = (idx, k)
        else:
            ip = self._paths[self._lowest[0]][0][self._lowest[1]]
            if contour[k][1] > ip[1] or (contour[k][1] == ip[1] and contour[k][0] < ip[0]):
                self._lowest = (idx, k)

    def _fix_orientations(self) -> None:
        # the closed polygon holding the lowest vertex must have a positive Clipper area, else every closed polygon
        # is reversed (so that a positive delta always grows the outer polygon)
        if self._lowest is not None and not (_area(self._paths[self._lowest[0]][0]) >= 0):
            self._paths = [(c[::-1] if et == ET_CLOSEDPOLYGON else c, jt, et) for c, jt, et in self._paths]
        # (ET_CLOSEDLINE re-orientation is not needed on this path)

    def Execute(self, delta: float) -> List[List[List[int]]]:
        self._fix_orientations()
        out: List[List[List[int]]] = []
        if abs(delta) < 1e-20:
            return [[list(p) for p in c] for c, _, et in self._paths if et == ET_CLOSEDPOLYGON]
        y = self.ArcTolerance
        if y <= 0.0:
            y = _DEF_ARC_TOLERANCE
        elif y > abs(delta) * _DEF_ARC_TOLERANCE:
            y = abs(delta) * _DEF_ARC_TOLERANCE
        steps = math.pi / math.acos(1.0 - y / abs(delta))
        if steps > abs(delta) * math.pi:
            steps = abs(delta) * math.pi  # excessive precision check
        m_sin, m_cos = math.sin(_TWO_PI / steps), math.cos(_TWO_PI / steps)
        steps_per_rad = steps / _TWO_PI
        if delta < 0.0:
            m_sin = -m_sin
        for src, jt, et in self._paths:
            n = len(src)
            if n == 0 or (delta <= 0 and (n < 3 or et != ET_CLOSEDPOLYGON)):
                continue
            if et != ET_CLOSEDPOLYGON or jt != JT_ROUND or n < 3:
                raise NotImplementedError("restated for JT_ROUND / ET_CLOSEDPOLYGON polygons (src/det/test.py:41)")
            normals = [_unit_normal(src[j], src[j + 1]) for j in range(n - 1)]
            normals.append(_unit_normal(src[n - 1], src[0]))
            dest: List[List[int]] = []
            k = n - 1
            for j in range(n):
                nk, nj = normals[k], normals[j]
                sin_a = nk[0] * nj[1] - nj[0] * nk[1]
                done = False
                if abs(sin_a * delta) < 1.0:
                    cos_a = nk[0] * nj[0] + nj[1] * nk[1]
                    if cos_a > 0:  # angle ==> 0 degrees
                        dest.append([_round(src[j][0] + nk[0] * delta), _round(src[j][1] + nk[1] * delta)])
                        done = True
                    # else angle ==> 180 degrees
                elif sin_a > 1.0:
                    sin_a = 1.0
                elif sin_a < -1.0:
                    sin_a = -1.0
                if not done:
                    if sin_a * delta < 0:  # concave vertex: the union removes the resulting loop
                        dest.append([_round(src[j][0] + nk[0] * delta), _round(src[j][1] + nk[1] * delta)])
                        dest.append([src[j][0], src[j][1]])
                        dest.append([_round(src[j][0] + nj[0] * delta), _round(src[j][1] + nj[1] * delta)])
                    else:  # jtRound
                        a = math.atan2(sin_a, nk[0] * nj[0] + nk[1] * nj[1])
                        nsteps = max(_round(steps_per_rad * abs(a)), 1)
                        x, yy = nk
                        for _ in range(nsteps):
                            dest.append([_round(src[j][0] + x * delta), _round(src[j][1] + yy * delta)])
                            x2 = x
                            x = x * m_cos - m_sin * yy
                            yy = x2 * m_sin + yy * m_cos
                        dest.append([_round(src[j][0] + nj[0] * delta), _round(src[j][1] + nj[1] * delta)])
                k = j
            out.append(dest)
        return out


class Polygon:
    """shapely.geometry.Polygon as far as src/det/test.py:38-39 uses it: unsigned shoelace area and ring length."""

    def __init__(self, shell):
        self._pts = [(float(p[0]), float(p[1])) for p in shell]

    @property
    def area(self) -> float:
        p = self._pts
        return abs(sum(p[i][0] * p[(i + 1) % len(p)][1] - p[(i + 1) % len(p)][0] * p[i][1] for i in range(len(p)))) * 0.5

    @property
    def length(self) -> float:
        p = self._pts
        return sum(math.hypot(p[(i + 1) % len(p)][0] - p[i][0], p[(i + 1) % len(p)][1] - p[i][1]) for i in range(len(p)))

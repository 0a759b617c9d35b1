"""Restatement of the two geometry wheels DBPostProcess.unclip calls (TEST INFRASTRUCTURE — oracle side).

`onnxocr/db_postprocess.py:151-157`:

    poly = Polygon(box)                                  # shapely (GEOS)      requirements.txt:7, unpinned
    distance = poly.area * unclip_ratio / poly.length
    offset = pyclipper.PyclipperOffset()                 # Clipper 6.4.2       requirements.txt:8, unpinned
    offset.AddPath(box, pyclipper.JT_ROUND, pyclipper.ET_CLOSEDPOLYGON)
    expanded = np.array(offset.Execute(distance))

Neither wheel is installed here or in /opt/wheelhouse, and the reference holds no test or
golden vector for them: **parity unpinned** at this boundary.  What follows restates the
published algorithms:

  * Polygon.area / .length: shoelace formula (absolute value) and ring perimeter in float64.
  * ClipperOffset (clipper.cpp, v6.4.2: AddPath / FixOrientations / DoOffset / OffsetPoint /
    DoRound / GetUnitNormal / Round), MiterLimit 2.0, ArcTolerance 0.25 (pyclipper defaults),
    input coordinates truncated to integers as pyclipper's `<cInt>` cast does.
  * The final `Clipper.Execute(ctUnion, pftPositive)` clean-up pass of ClipperOffset::Execute is
    NOT restated: for a convex quad grown outwards it returns the same vertex set (possibly
    starting at a different vertex and without collinear/duplicate points), and the only consumer
    is `cv2.minAreaRect`, which works on the convex hull.  The vertex order can only matter through
    float tie-breaking inside rotatingCalipers.
"""
import math

import numpy as np

JT_ROUND = 1
ET_CLOSEDPOLYGON = 0


class Polygon(object):
    """Stand-in for shapely.geometry.Polygon (area and length only)."""

    def __init__(self, pts):
        self.pts = [(float(p[0]), float(p[1])) for p in pts]

    @property
    def area(self):
        s = 0.0
        n = len(self.pts)
        for i in range(n):
            xi, yi = self.pts[i]
            xj, yj = self.pts[(i + 1) % n]
            s += xi * yj - xj * yi
        return abs(s) * 0.5

    @property
    def length(self):
        s = 0.0
        n = len(self.pts)
        for i in range(n):
            xi, yi = self.pts[i]
            xj, yj = self.pts[(i + 1) % n]
            dx, dy = xj - xi, yj - yi
            s += math.sqrt(dx * dx + dy * dy)
        return s


def _round(v):
    return int(v - 0.5) if v < 0 else int(v + 0.5)


class PyclipperOffset(object):
    """Stand-in for pyclipper.PyclipperOffset (one closed polygon, JT_ROUND)."""

    def __init__(self, miter_limit=2.0, arc_tolerance=0.25):
        self.MiterLimit = miter_limit
        self.ArcTolerance = arc_tolerance
        self.paths = []

    def AddPath(self, path, join_type, end_type):
        assert join_type == JT_ROUND and end_type == ET_CLOSEDPOLYGON
        pts = [(int(p[0]), int(p[1])) for p in path]          # <cInt> cast: truncation toward zero
        high = len(pts) - 1
        if high < 0:
            return
        while high > 0 and pts[0] == pts[high]:
            high -= 1
        contour = [pts[0]]
        for i in range(1, high + 1):
            if contour[-1] != pts[i]:
                contour.append(pts[i])
        if len(contour) < 3:
            return
        self.paths.append(contour)

    @staticmethod
    def _area(poly):
        a = 0.0
        j = len(poly) - 1
        for i in range(len(poly)):
            a += (float(poly[j][0]) + float(poly[i][0])) * (float(poly[j][1]) - float(poly[i][1]))
            j = i
        return -a * 0.5

    def Execute(self, delta):
        out = []
        paths = [list(p) for p in self.paths]
        # FixOrientations: the lowest path decides; with a single path: reverse if its area < 0
        if paths and not (self._area(paths[0]) >= 0):
            paths = [p[::-1] for p in paths]
        for src in paths:
            out.append(self._offset(src, float(delta)))
        return [p for p in out if p]

    def _offset(self, src, delta):
        n = len(src)
        if abs(delta) < 1.0e-20:
            return [list(p) for p in src]
        pi = 3.141592653589793238
        two_pi = pi * 2
        if self.ArcTolerance <= 0.0:
            y = 0.25
        elif self.ArcTolerance > abs(delta) * 0.25:
            y = abs(delta) * 0.25
        else:
            y = self.ArcTolerance
        steps = pi / math.acos(1 - y / abs(delta))
        if steps > abs(delta) * pi:
            steps = abs(delta) * pi
        m_sin = math.sin(two_pi / steps)
        m_cos = math.cos(two_pi / steps)
        steps_per_rad = steps / two_pi
        if delta < 0.0:
            m_sin = -m_sin
        normals = []
        for j in range(n):
            p1, p2 = src[j], src[(j + 1) % n]
            if p1 == p2:
                normals.append((0.0, 0.0))
                continue
            dx = float(p2[0] - p1[0])
            dy = float(p2[1] - p1[1])
            f = 1 * 1.0 / math.sqrt(dx * dx + dy * dy)
            dx *= f
            dy *= f
            normals.append((dy, -dx))
        dest = []
        k = n - 1
        for j in range(n):
            sin_a = normals[k][0] * normals[j][1] - normals[j][0] * normals[k][1]
            done = False
            if abs(sin_a * delta) < 1.0:
                cos_a = normals[k][0] * normals[j][0] + normals[j][1] * normals[k][1]
                if cos_a > 0:
                    dest.append([_round(src[j][0] + normals[k][0] * delta), _round(src[j][1] + normals[k][1] * delta)])
                    done = True
            elif sin_a > 1.0:
                sin_a = 1.0
            elif sin_a < -1.0:
                sin_a = -1.0
            if not done:
                if sin_a * delta < 0:
                    dest.append([_round(src[j][0] + normals[k][0] * delta), _round(src[j][1] + normals[k][1] * delta)])
                    dest.append([src[j][0], src[j][1]])
                    dest.append([_round(src[j][0] + normals[j][0] * delta), _round(src[j][1] + normals[j][1] * delta)])
                else:
                    a = math.atan2(sin_a, normals[k][0] * normals[j][0] + normals[k][1] * normals[j][1])
                    isteps = max(_round(steps_per_rad * abs(a)), 1)
                    X, Y = normals[k]
                    for _ in range(isteps):
                        dest.append([_round(src[j][0] + X * delta), _round(src[j][1] + Y * delta)])
                        X2 = X
                        X = X * m_cos - m_sin * Y
                        Y = X2 * m_sin + Y * m_cos
                    dest.append([_round(src[j][0] + normals[j][0] * delta), _round(src[j][1] + normals[j][1] * delta)])
            k = j
        return dest

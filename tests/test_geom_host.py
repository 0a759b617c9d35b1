"""CPU fuzzing of the product's geometry header (csrc/dbgeom.h, host build) against the installed cv2.

The same functions run inside the CUDA DBPostProcess kernels; here they are pinned, bit for bit, against
cv2.findContours / minAreaRect / boxPoints / fillPoly / line, i.e. the third-party arithmetic
`onnxocr/db_postprocess.py` calls.
"""
import ctypes as C
import os
import sys

import cv2
import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "host"))
import build_host  # noqa: E402

from oracle import clipper_offset, restate_post  # noqa: E402


@pytest.fixture(scope="module")
def G():
    L = C.CDLL(build_host.build())
    L.geom_min_area_rect.restype = C.c_float
    L.geom_box_score.restype = C.c_double
    L.geom_box_score_slow.restype = C.c_double
    return L


def P(a):
    return a.ctypes.data_as(C.c_void_p)


def blob_bitmap(rng, h, w, n_blobs, holes=True, noise=0.0):
    bm = np.zeros((h, w), np.uint8)
    for _ in range(n_blobs):
        c = rng.uniform([5, 5], [w - 5, h - 5])
        ang = rng.uniform(-0.6, 0.6)
        bw, bh = rng.uniform(3, w / 3), rng.uniform(2, h / 6)
        R = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]])
        pts = (np.array([[-bw / 2, -bh / 2], [bw / 2, -bh / 2], [bw / 2, bh / 2], [-bw / 2, bh / 2]]) @ R.T + c).astype(np.int32)
        cv2.fillPoly(bm, [pts], 1)
    if holes:
        for _ in range(n_blobs):
            x, y = int(rng.integers(1, w - 4)), int(rng.integers(1, h - 4))
            bm[y:y + int(rng.integers(1, 4)), x:x + int(rng.integers(1, 5))] = 0
    if noise > 0:
        flip = rng.random((h, w)) < noise
        bm[flip] ^= 1
    return bm


def contour_starts(bm):
    """Start pixels in cv2 order: fg 8-components' raster-first pixel (outer) and, for
    bg 4-components that do not touch the border, the pixel left of their raster-first pixel (hole)."""
    h, w = bm.shape
    n_fg, lab_fg = cv2.connectedComponents(bm, connectivity=8)
    n_bg, lab_bg = cv2.connectedComponents((1 - bm).astype(np.uint8), connectivity=4)
    cands = []
    flat_fg, flat_bg = lab_fg.ravel(), lab_bg.ravel()
    first_fg = {}
    for i, l in enumerate(flat_fg):
        if l and l not in first_fg:
            first_fg[l] = i
    first_bg = {}
    for i, l in enumerate(flat_bg):
        if l and l not in first_bg:
            first_bg[l] = i
    border = set(lab_bg[0, :].tolist()) | set(lab_bg[-1, :].tolist()) | set(lab_bg[:, 0].tolist()) | set(lab_bg[:, -1].tolist())
    for l, i in first_fg.items():
        cands.append((i, 0))
    for l, i in first_bg.items():
        if l not in border:
            cands.append((i - 1, 1))
    cands.sort(key=lambda t: -t[0])
    return [((i % w, i // w), hole) for i, hole in cands]


@pytest.mark.parametrize("seed,noise", [(0, 0.0), (1, 0.0), (2, 0.02), (3, 0.2), (4, 0.5)])
def test_contours_match_cv2(G, seed, noise):
    rng = np.random.default_rng(seed)
    for _ in range(6):
        h, w = int(rng.integers(16, 120)), int(rng.integers(16, 160))
        bm = blob_bitmap(rng, h, w, int(rng.integers(1, 8)), noise=noise)
        ref, _ = cv2.findContours((bm * 255).astype(np.uint8), cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
        starts = contour_starts(bm)
        assert len(starts) == len(ref)
        buf = np.zeros((8192, 2), np.int32)
        for (sx, sy), hole in starts:
            pass
        for k, ((sx, sy), hole) in enumerate(starts):
            n = G.geom_trace(P(bm), h, w, sx, sy, hole, P(buf), 8192)
            got = buf[:n].copy()
            exp = ref[k].reshape(-1, 2)
            assert n == len(exp), (k, n, len(exp))
            assert np.array_equal(got, exp), (k, hole)


def _rect(G, pts, clockwise=0):
    pts = np.ascontiguousarray(pts, np.int32)
    rect = np.zeros(5, np.float32)
    bp = np.zeros(8, np.float32)
    mini = np.zeros(8, np.float32)
    hull = np.zeros(len(pts) + 2, np.int32)
    nh = C.c_int()
    ss = G.geom_min_area_rect(P(pts), len(pts), P(rect), P(bp), P(mini), P(hull), C.byref(nh), clockwise)
    return rect, bp.reshape(4, 2), mini.reshape(4, 2), ss, hull[:nh.value]


def test_min_area_rect_on_contours_bit_exact(G):
    rng = np.random.default_rng(7)
    n_checked = 0
    for it in range(40):
        h, w = int(rng.integers(30, 200)), int(rng.integers(30, 300))
        bm = blob_bitmap(rng, h, w, int(rng.integers(1, 10)), noise=0.02 if it % 3 == 0 else 0.0)
        ref, _ = cv2.findContours((bm * 255).astype(np.uint8), cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
        for c in ref:
            r = cv2.minAreaRect(c)
            exp_rect = np.array([r[0][0], r[0][1], r[1][0], r[1][1], r[2]], np.float32)
            exp_bp = cv2.boxPoints(r)
            exp_mini, exp_ss = restate_post.get_mini_boxes(c)
            rect, bp, mini, ss, hull = _rect(G, c.reshape(-1, 2))
            # hull: same vertices in the same cyclic order.  When a contour visits a pixel twice cv2 4.13 may
            # pick the other copy (and therefore another start vertex); minAreaRect is unaffected (asserted below).
            pts = c.reshape(-1, 2)
            exp_hull = cv2.convexHull(c, clockwise=False, returnPoints=False).ravel()
            a, b = [tuple(p) for p in pts[hull]], [tuple(p) for p in pts[exp_hull]]
            assert len(a) == len(b) and a[0] in b
            k = b.index(a[0])
            assert a == b[k:] + b[:k]
            if len(set(map(tuple, pts.tolist()))) == len(pts):
                assert np.array_equal(hull, exp_hull)
            assert np.array_equal(rect, exp_rect), (c.reshape(-1, 2).tolist(), rect, exp_rect)
            assert np.array_equal(bp, exp_bp)
            assert np.array_equal(mini, np.array(exp_mini))
            assert ss == np.float32(exp_ss)
            n_checked += 1
    assert n_checked > 300


def test_min_area_rect_random_point_sets_bit_exact(G):
    rng = np.random.default_rng(3)
    for t in range(20000):
        n = int(rng.integers(1, 14))
        pts = rng.integers(-60, int(rng.integers(4, 3000)), (n, 2)).astype(np.int32)
        if t % 3 == 0:
            c = rng.uniform(20, 140, 2)
            ang = rng.uniform(-np.pi, np.pi)
            bw, bh = rng.uniform(2, 130), rng.uniform(2, 30)
            R = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]])
            pts = (np.array([[-bw / 2, -bh / 2], [bw / 2, -bh / 2], [bw / 2, bh / 2], [-bw / 2, bh / 2]]) @ R.T + c).astype(np.int32)
        if t % 17 == 0:
            pts[:, 1] = pts[0, 1]          # collinear
        r = cv2.minAreaRect(pts.reshape(-1, 1, 2))
        exp_rect = np.array([r[0][0], r[0][1], r[1][0], r[1][1], r[2]], np.float32)
        rect, bp, mini, ss, hull = _rect(G, pts)
        assert np.array_equal(rect, exp_rect), (pts.tolist(), rect, exp_rect)
        assert np.array_equal(bp, cv2.boxPoints(r))


def test_min_area_rect_on_offset_polygons_bit_exact(G):
    rng = np.random.default_rng(8)
    for it in range(400):
        c = rng.uniform(40, 400, 2)
        ang = rng.uniform(-np.pi / 2, np.pi / 2) if it % 4 else 0.0
        bw, bh = rng.uniform(10, 300), rng.uniform(4, 40)
        R = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]])
        box = (np.array([[-bw / 2, -bh / 2], [bw / 2, -bh / 2], [bw / 2, bh / 2], [-bw / 2, bh / 2]]) @ R.T + c).astype(np.float32)
        exp = restate_post.unclip(box, 1.5)
        out = np.zeros((256, 2), np.int32)
        dist = C.c_double()
        n = G.geom_unclip(P(np.ascontiguousarray(box)), C.c_double(1.5), P(out), 256, C.byref(dist))
        assert n == exp.shape[1]
        assert np.array_equal(out[:n], exp[0])
        poly = clipper_offset.Polygon(box)
        assert dist.value == poly.area * 1.5 / poly.length
        pts = exp.reshape(-1, 1, 2).astype(np.int32)
        exp_mini, exp_ss = restate_post.get_mini_boxes(pts)
        rect, bp, mini, ss, hull = _rect(G, pts.reshape(-1, 2))
        assert np.array_equal(mini, np.array(exp_mini)), it
        assert ss == np.float32(exp_ss)


def test_fillpoly_and_line8_bit_exact(G):
    rng = np.random.default_rng(1)
    for t in range(3000):
        c = rng.uniform(-5, 70, 2)
        ang = rng.uniform(-np.pi, np.pi)
        bw, bh = rng.uniform(1, 60), rng.uniform(1, 25)
        R = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]])
        q = (np.array([[-bw / 2, -bh / 2], [bw / 2, -bh / 2], [bw / 2, bh / 2], [-bw / 2, bh / 2]]) @ R.T + c).astype(np.int32)
        if t % 7 == 0:
            q = rng.integers(-20, 90, (4, 2)).astype(np.int32)
        mh, mw = int(rng.integers(8, 64)), int(rng.integers(8, 64))
        ref = np.zeros((mh, mw), np.uint8)
        cv2.fillPoly(ref, q.reshape(1, -1, 2), 1)
        m = np.zeros((mh, mw), np.uint8)
        G.geom_fill_mask(P(np.ascontiguousarray(q[:, 0])), P(np.ascontiguousarray(q[:, 1])), mh, mw, P(m))
        assert np.array_equal(ref, m), q.tolist()
    for t in range(2000):
        a, b = rng.integers(0, 40, 2), rng.integers(0, 40, 2)
        ref = np.zeros((40, 40), np.uint8)
        cv2.line(ref, (int(a[0]), int(a[1])), (int(b[0]), int(b[1])), 1, 1, cv2.LINE_8)
        m = np.zeros((40, 40), np.uint8)
        G.geom_line_mask(int(a[0]), int(a[1]), int(b[0]), int(b[1]), 40, 40, P(m))
        assert np.array_equal(ref, m)


def test_box_score_matches_cv2_mean(G):
    rng = np.random.default_rng(5)
    h, w = 96, 128
    pred = rng.random((h, w)).astype(np.float32)
    for t in range(400):
        c = rng.uniform([-4, -4], [w + 4, h + 4])
        ang = rng.uniform(-1.0, 1.0)
        bw, bh = rng.uniform(3, 80), rng.uniform(3, 30)
        R = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]])
        box = (np.array([[-bw / 2, -bh / 2], [bw / 2, -bh / 2], [bw / 2, bh / 2], [-bw / 2, bh / 2]]) @ R.T + c).astype(np.float32)
        exp = restate_post.box_score_fast(pred, box)
        got = G.geom_box_score(P(pred), h, w, P(np.ascontiguousarray(box)))
        assert abs(got - exp) < 1e-9, (t, got, exp)


@pytest.mark.parametrize("seed,noise", [(0, 0.0), (1, 0.02), (2, 0.2), (3, 0.5)])
def test_box_score_slow_matches_cv2_fillpoly_of_the_contour(G, seed, noise):
    """det_db_score_mode='slow' (db_postprocess.py:199-218): the mask is cv2.fillPoly of the contour itself, outer and
    hole contours alike; the product's even-odd scan of the chain must give the same mean (same mask) as cv2."""
    rng = np.random.default_rng(100 + seed)
    n_checked = n_holes = 0
    for it in range(8):
        h, w = int(rng.integers(16, 120)), int(rng.integers(16, 200))
        if it % 4 == 3:
            bm = (cv2.GaussianBlur(rng.random((h, w)).astype(np.float32), (7, 7), 0) * 1.6 > 0.8).astype(np.uint8)
        else:
            bm = blob_bitmap(rng, h, w, int(rng.integers(1, 8)), noise=noise)
        # scores that expose any single-pixel difference of the mask: distinct powers of two would overflow, random is enough
        pred = rng.random((h, w)).astype(np.float32)
        ref, _ = cv2.findContours((bm * 255).astype(np.uint8), cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
        starts = contour_starts(bm)
        assert len(starts) == len(ref)
        for k, ((sx, sy), hole) in enumerate(starts):
            exp = restate_post.box_score_slow(pred, ref[k])
            got = G.geom_box_score_slow(P(pred), P(bm), h, w, sx, sy, hole)
            assert abs(got - exp) < 1e-12, (it, k, hole, got, exp)
            n_checked += 1
            n_holes += hole
    assert n_checked > 20 and (noise == 0.0 or n_holes > 0)


def test_scale_coord_matches_numpy(G):
    rng = np.random.default_rng(9)
    for t in range(2000):
        v = np.float32(rng.uniform(-5, 1000))
        width = int(rng.integers(32, 2560))
        dest = float(rng.integers(10, 4000))
        box = np.array([v], np.float32)
        exp = np.clip(np.round(box / width * np.float64(dest)), 0, dest).astype(np.float32).astype("int32")[0]
        assert G.geom_scale_coord(C.c_float(v), width, C.c_double(dest)) == exp

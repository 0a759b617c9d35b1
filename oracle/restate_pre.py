"""numpy restatement of the det / cls / rec pre-processing (TEST INFRASTRUCTURE — oracle side).

Each function cites the reference lines it follows (paths relative to /root/reference).
The integer pixel arithmetic of `cv2.resize(INTER_LINEAR)` and
`cv2.warpPerspective(INTER_CUBIC, BORDER_REPLICATE)` lives in opencv-python
(pinned 4.7.0.72 in requirements.txt:3-4; this image has 4.13.0), which is not under
/root/reference: it is restated here from OpenCV's published fixed-point algorithm and
pinned against the installed cv2 by tests/test_oracle_pre.py (0 mismatching bytes).
"""
import math

import numpy as np


# ---------------------------------------------------------------------------------------
# DetResizeForTest.resize_image_type0   onnxocr/operators.py:91-138 (+ :54-76)
# ---------------------------------------------------------------------------------------
def det_target_size(h, w, limit_side_len=960, limit_type="max"):
    """Returns (pad_h, pad_w, resize_h, resize_w, ratio_h, ratio_w) exactly as the reference computes them."""
    pad_h, pad_w = h, w
    if h + w < 64:                                   # operators.py:57-58, 72-76
        pad_h, pad_w = max(32, h), max(32, w)
    hh, ww = pad_h, pad_w
    if limit_type == "max":                          # operators.py:103-111
        if max(hh, ww) > limit_side_len:
            ratio = float(limit_side_len) / hh if hh > ww else float(limit_side_len) / ww
        else:
            ratio = 1.0
    elif limit_type == "min":                        # operators.py:112-119
        if min(hh, ww) < limit_side_len:
            ratio = float(limit_side_len) / hh if hh < ww else float(limit_side_len) / ww
        else:
            ratio = 1.0
    elif limit_type == "resize_long":                # operators.py:120-121
        ratio = float(limit_side_len) / max(hh, ww)
    else:
        raise Exception("not support limit type, image ")
    resize_h = int(hh * ratio)                       # operators.py:124-128 (round() is half-to-even)
    resize_w = int(ww * ratio)
    resize_h = max(int(round(resize_h / 32) * 32), 32)
    resize_w = max(int(round(resize_w / 32) * 32), 32)
    return pad_h, pad_w, resize_h, resize_w, resize_h / float(hh), resize_w / float(ww)


def _linear_coeffs(src, dst, horizontal):
    """cv2 INTER_LINEAR taps for one axis: (index0, index1, w0, w1) with 11-bit integer weights."""
    scale = 1.0 / (dst / src)
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if horizontal:
        lo = s < 0
        f[lo] = 0
        s[lo] = 0
        hi = s >= src - 1
        f[hi] = 0
        s[hi] = src - 1
        i0, i1 = s, np.minimum(s + 1, src - 1)
    else:
        i0 = np.clip(s, 0, src - 1)
        i1 = np.clip(s + 1, 0, src - 1)
    w0 = np.rint((np.float32(1.0) - f) * np.float32(2048)).astype(np.int64)
    w1 = np.rint(f * np.float32(2048)).astype(np.int64)
    return i0, i1, np.clip(w0, -32768, 32767), np.clip(w1, -32768, 32767)


def resize_linear_u8(img, dst_w, dst_h):
    """Bit-exact `cv2.resize(img, (dst_w, dst_h))` for HWC uint8 (operators.py:132, predict_rec.py:73,
    predict_cls.py:31)."""
    h, w = img.shape[:2]
    if (h, w) == (dst_h, dst_w):
        return img.copy()
    if w == 2 * dst_w and h == 2 * dst_h:            # cv2 silently switches to INTER_AREA
        a = img.astype(np.int32)
        return ((a[0::2, 0::2] + a[0::2, 1::2] + a[1::2, 0::2] + a[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    x0, x1, wx0, wx1 = _linear_coeffs(w, dst_w, True)
    y0, y1, wy0, wy1 = _linear_coeffs(h, dst_h, False)
    a = img.astype(np.int64)
    rows = a[:, x0] * wx0[None, :, None] + a[:, x1] * wx1[None, :, None]          # scale 2^11
    r0, r1 = rows[y0], rows[y1]
    out = (((wy0[:, None, None] * (r0 >> 4)) >> 16) + ((wy1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


_DET_MEAN = np.array([0.485, 0.456, 0.406], dtype=np.float32)
_DET_STD = np.array([0.229, 0.224, 0.225], dtype=np.float32)


def det_preprocess(img, limit_side_len=960, limit_type="max"):
    """DetResizeForTest + NormalizeImage + ToCHWImage (operators.py:54-70, 22-31, 170-176; predict_det.py:98-104).

    Returns (chw float32 [1,3,H,W], shape float64 [src_h, src_w, ratio_h, ratio_w])."""
    src_h, src_w = img.shape[:2]
    pad_h, pad_w, rh, rw, ratio_h, ratio_w = det_target_size(src_h, src_w, limit_side_len, limit_type)
    if (pad_h, pad_w) != (src_h, src_w):
        pad = np.zeros((pad_h, pad_w, 3), np.uint8)
        pad[:src_h, :src_w] = img
        img = pad
    res = resize_linear_u8(img, rw, rh)
    x = (res.astype(np.float32) * np.float32(1.0 / 255.0) - _DET_MEAN) / _DET_STD      # operators.py:29-30
    return x.transpose(2, 0, 1)[None].copy(), np.array([src_h, src_w, ratio_h, ratio_w])


# ---------------------------------------------------------------------------------------
# rec / cls resize_norm_img   predict_rec.py:54-80, predict_cls.py:22-42
# ---------------------------------------------------------------------------------------
def rec_resized_w(h, w, img_w, img_h=48):
    ratio = w / float(h)
    if math.ceil(img_h * ratio) > img_w:
        return img_w
    return int(math.ceil(img_h * ratio))


def resize_norm_img(img, img_w, img_h=48):
    """((resize(img)/255) - 0.5) / 0.5, CHW, zero padded on the right to img_w."""
    h, w = img.shape[:2]
    rw = rec_resized_w(h, w, img_w, img_h)
    res = resize_linear_u8(img, rw, img_h).astype(np.float32)
    res = res.transpose(2, 0, 1) / 255
    res -= 0.5
    res /= 0.5
    out = np.zeros((3, img_h, img_w), np.float32)
    out[:, :, :rw] = res
    return out


def rec_batches(whs, batch_num=6, img_h=48, img_w0=320):
    """Batch plan of TextRecognizer.__call__ (predict_rec.py:282-300): list of (indices, imgW)."""
    ratios = np.array([w / float(h) for (w, h) in whs])
    order = np.argsort(ratios)
    out = []
    for beg in range(0, len(whs), batch_num):
        idx = order[beg:beg + batch_num]
        max_wh = img_w0 / img_h
        for i in idx:
            w, h = whs[i]
            max_wh = max(max_wh, w * 1.0 / h)
        out.append((idx, int(img_h * max_wh)))
    return out


# ---------------------------------------------------------------------------------------
# get_rotate_crop_image   onnxocr/utils.py:12-53
# ---------------------------------------------------------------------------------------
def crop_size(points):
    """(crop_w, crop_h) = int(max of float32 edge norms)  utils.py:24-33."""
    p = points.astype(np.float32)

    def norm(a, b):
        return np.linalg.norm(a - b)

    cw = int(max(norm(p[0], p[1]), norm(p[2], p[3])))
    ch = int(max(norm(p[0], p[3]), norm(p[1], p[2])))
    return cw, ch


def perspective_matrix(src, dst):
    """cv2.getPerspectiveTransform(src f32[4,2], dst f32[4,2]) — 8x8 system solved by LU with partial pivoting
    in float64 (OpenCV `getPerspectiveTransform` -> `solve(A, B, X, DECOMP_LU)`)."""
    a = np.zeros((8, 8), np.float64)
    b = np.zeros(8, np.float64)
    for i in range(4):
        sx, sy = float(src[i][0]), float(src[i][1])
        dx, dy = float(dst[i][0]), float(dst[i][1])
        a[i] = [sx, sy, 1, 0, 0, 0, -sx * dx, -sy * dx]
        a[i + 4] = [0, 0, 0, sx, sy, 1, -sx * dy, -sy * dy]
        b[i] = dx
        b[i + 4] = dy
    x = _lu_solve(a, b)
    return np.append(x, 1.0).reshape(3, 3)


def _lu_solve(a, b):
    """Gaussian elimination with partial pivoting, the operation order of OpenCV's hal::LU."""
    a = a.copy()
    b = b.copy()
    n = a.shape[0]
    for i in range(n):
        k = i
        for j in range(i + 1, n):
            if abs(a[j, i]) > abs(a[k, i]):
                k = j
        if abs(a[k, i]) < np.finfo(np.float64).eps * 100:
            return np.zeros(n)
        if k != i:
            a[[i, k]] = a[[k, i]]
            b[[i, k]] = b[[k, i]]
        d = -1.0 / a[i, i]
        for j in range(i + 1, n):
            alpha = a[j, i] * d
            for c in range(i + 1, n):
                a[j, c] += alpha * a[i, c]
            b[j] += alpha * b[i]
    for i in range(n - 1, -1, -1):
        s = b[i]
        for k in range(i + 1, n):
            s -= a[i, k] * b[k]
        b[i] = s / a[i, i]
    return b


def invert3x3(m):
    """cv2 `invert(M, dst)` for a 3x3 double matrix (closed-form adjugate / determinant)."""
    d = (m[0, 0] * (m[1, 1] * m[2, 2] - m[1, 2] * m[2, 1]) - m[0, 1] * (m[1, 0] * m[2, 2] - m[1, 2] * m[2, 0]) +
         m[0, 2] * (m[1, 0] * m[2, 1] - m[1, 1] * m[2, 0]))
    if d == 0:
        return np.zeros((3, 3))
    d = 1.0 / d
    t = np.empty((3, 3), np.float64)
    t[0, 0] = (m[1, 1] * m[2, 2] - m[1, 2] * m[2, 1]) * d
    t[0, 1] = (m[0, 2] * m[2, 1] - m[0, 1] * m[2, 2]) * d
    t[0, 2] = (m[0, 1] * m[1, 2] - m[0, 2] * m[1, 1]) * d
    t[1, 0] = (m[1, 2] * m[2, 0] - m[1, 0] * m[2, 2]) * d
    t[1, 1] = (m[0, 0] * m[2, 2] - m[0, 2] * m[2, 0]) * d
    t[1, 2] = (m[0, 2] * m[1, 0] - m[0, 0] * m[1, 2]) * d
    t[2, 0] = (m[1, 0] * m[2, 1] - m[1, 1] * m[2, 0]) * d
    t[2, 1] = (m[0, 1] * m[2, 0] - m[0, 0] * m[2, 1]) * d
    t[2, 2] = (m[0, 0] * m[1, 1] - m[0, 1] * m[1, 0]) * d
    return t


def _cubic_tab():
    """OpenCV's 32x32 table of 4x4 int16 bicubic weights (A = -0.75), sum forced to 32768."""
    A = np.float32(-0.75)
    tab = np.zeros((32, 4), np.float32)
    for i in range(32):
        x = np.float32(i) * np.float32(1.0 / 32)
        c = np.zeros(4, np.float32)
        c[0] = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A
        c[1] = ((A + 2) * x - (A + 3)) * x * x + 1
        c[2] = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1
        c[3] = np.float32(1.0) - c[0] - c[1] - c[2]
        tab[i] = c
    out = np.zeros((32, 32, 4, 4), np.int16)
    for iy in range(32):
        for ix in range(32):
            w = np.zeros((4, 4), np.int32)
            for ky in range(4):
                for kx in range(4):
                    v = np.float32(tab[iy, ky] * tab[ix, kx])
                    w[ky, kx] = int(np.clip(np.rint(np.float32(v * np.float32(32768))), -32768, 32767))
            s = int(w.sum())
            if s != 32768:
                diff = s - 32768
                # OpenCV searches the entry to fix only among ky,kx in {2,3} (ksize2 = 2)
                sub = w[2:4, 2:4]
                if diff < 0:
                    k = np.unravel_index(np.argmax(sub), sub.shape)
                else:
                    k = np.unravel_index(np.argmin(sub), sub.shape)
                w[2 + k[0], 2 + k[1]] -= diff
            out[iy, ix] = w.astype(np.int16)
    return out


_CUBIC = None


def warp_perspective_cubic(img, M, cw, ch):
    """Bit-exact `cv2.warpPerspective(img, M, (cw, ch), flags=INTER_CUBIC, borderMode=BORDER_REPLICATE)` (utils.py:43-49)."""
    global _CUBIC
    if _CUBIC is None:
        _CUBIC = _cubic_tab()
    h, w = img.shape[:2]
    Mi = invert3x3(np.asarray(M, np.float64))
    xs = np.arange(cw, dtype=np.float64)[None, :]
    ys = np.arange(ch, dtype=np.float64)[:, None]
    X0 = Mi[0, 0] * xs + (Mi[0, 1] * ys + Mi[0, 2])
    Y0 = Mi[1, 0] * xs + (Mi[1, 1] * ys + Mi[1, 2])
    W0 = Mi[2, 0] * xs + (Mi[2, 1] * ys + Mi[2, 2])
    with np.errstate(divide="ignore", invalid="ignore"):
        Wn = np.where(W0 != 0, 32.0 / W0, 0.0)
    fX = np.clip(X0 * Wn, -2147483648.0, 2147483647.0)
    fY = np.clip(Y0 * Wn, -2147483648.0, 2147483647.0)
    X = np.rint(fX).astype(np.int64)
    Y = np.rint(fY).astype(np.int64)
    sx = (X >> 5) - 1
    sy = (Y >> 5) - 1
    ax = (X & 31).astype(np.int64)
    ay = (Y & 31).astype(np.int64)
    wt = _CUBIC[ay, ax].astype(np.int64)                              # [ch,cw,4,4]
    acc = np.zeros((ch, cw, img.shape[2]), np.int64)
    src = img.astype(np.int64)
    for ky in range(4):
        yy = np.clip(sy + ky, 0, h - 1)
        for kx in range(4):
            xx = np.clip(sx + kx, 0, w - 1)
            acc += src[yy, xx] * wt[:, :, ky, kx][:, :, None]
    return np.clip((acc + (1 << 14)) >> 15, 0, 255).astype(np.uint8)


def get_rotate_crop_image(img, points):
    """utils.py:12-53 with the two cv2 calls replaced by the restatements above."""
    points = np.asarray(points, np.float32)
    cw, ch = crop_size(points)
    pts_std = np.float32([[0, 0], [cw, 0], [cw, ch], [0, ch]])
    M = perspective_matrix(points, pts_std)
    dst = warp_perspective_cubic(img, M, cw, ch)
    if dst.shape[0] * 1.0 / dst.shape[1] >= 1.5:
        dst = np.rot90(dst)
    return dst

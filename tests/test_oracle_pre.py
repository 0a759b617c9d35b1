"""CPU: the numpy restatement of cv2's fixed-point resize / bicubic warp is bit-exact against the installed cv2
(the third-party arithmetic behind operators.py:132, predict_rec.py:73, predict_cls.py:31, utils.py:42-49)."""
import cv2
import numpy as np

from oracle import restate_pre as rp


def test_resize_linear_bit_exact():
    rng = np.random.default_rng(0)
    cases = [(200, 300, 100, 150), (64, 64, 64, 64), (48, 100, 48, 100), (96, 200, 48, 100)]
    for _ in range(40):
        cases.append((int(rng.integers(8, 300)), int(rng.integers(8, 300)), int(rng.integers(8, 400)), int(rng.integers(8, 400))))
    for h, w, dh, dw in cases:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(cv2.resize(img, (dw, dh)), rp.resize_linear_u8(img, dw, dh)), (h, w, dh, dw)


def test_det_target_size_rules():
    assert rp.det_target_size(1920, 2560)[2:4] == (704, 960)
    assert rp.det_target_size(720, 1150)[2:4] == (608, 960)
    assert rp.det_target_size(170, 296)[2:4] == (160, 288)
    assert rp.det_target_size(64, 64)[2:4] == (64, 64)
    assert rp.det_target_size(20, 30)[:4] == (32, 32, 32, 32)
    assert rp.det_target_size(3508, 2480, 2560)[2:4] == (2560, 1824)
    assert rp.det_target_size(48, 48)[2:4] == (64, 64)          # round(1.5) == 2 (half to even)
    assert rp.det_target_size(80, 80)[2:4] == (64, 64)          # round(2.5) == 2


def test_warp_perspective_cubic_bit_exact():
    rng = np.random.default_rng(2)
    img = rng.integers(0, 256, (240, 320, 3), dtype=np.uint8)
    total = 0
    for t in range(10):
        c = rng.uniform([40, 40], [280, 200]) if t % 4 else np.array([6.0, 9.0])
        ang = rng.uniform(-0.6, 0.6)
        bw, bh = rng.uniform(30, 160), rng.uniform(10, 50)
        R = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]])
        pts = (np.array([[-bw / 2, -bh / 2], [bw / 2, -bh / 2], [bw / 2, bh / 2], [-bw / 2, bh / 2]]) @ R.T + c).astype(np.float32)
        cw, ch = rp.crop_size(pts)
        std = np.float32([[0, 0], [cw, 0], [cw, ch], [0, ch]])
        M = cv2.getPerspectiveTransform(pts, std)
        ref = cv2.warpPerspective(img, M, (cw, ch), borderMode=cv2.BORDER_REPLICATE, flags=cv2.INTER_CUBIC)
        assert np.array_equal(ref, rp.warp_perspective_cubic(img, M, cw, ch))
        assert np.array_equal(ref, rp.warp_perspective_cubic(img, rp.perspective_matrix(pts, std), cw, ch))
        total += ref.size
    assert total > 50000

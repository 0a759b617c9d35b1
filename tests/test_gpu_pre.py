"""GPU parity of pixel pre-processing (bit-exact) against the numpy/cv2 restatement."""
import numpy as np
import pytest

from onnxocr_b200 import _lib, synth_pages
from oracle import restate_pre

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("hw", [(960, 960), (1280, 720), (300, 517), (170, 296), (20, 30), (2000, 1500), (1920, 1920)])
def test_det_preprocess_bit_exact(engine_fp32, hw):
    h, w = hw
    img = np.random.default_rng(h * 7 + w).integers(0, 256, (h, w, 3), dtype=np.uint8)
    ref, ref_shape = restate_pre.det_preprocess(img)
    out, shape = engine_fp32.det_preprocess(img)
    assert out.shape == ref.shape
    assert np.array_equal(out, ref), np.abs(out - ref).max()
    assert np.array_equal(shape, ref_shape)


def test_det_preprocess_limit_types(engine_fp32):
    img = np.random.default_rng(0).integers(0, 256, (500, 700, 3), dtype=np.uint8)
    for lt, lim in (("min", 736), ("resize_long", 640), ("max", 2560)):
        ref, ref_shape = restate_pre.det_preprocess(img, lim, lt)
        out, shape = engine_fp32.det_preprocess(img, _lib.det_params(limit_side_len=lim, limit_type=lt))
        assert np.array_equal(out, ref)
        assert np.array_equal(shape, ref_shape)


def test_crop_boxes_bit_exact(engine_fp32):
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (600, 800, 3), dtype=np.uint8)
    boxes = []
    for t in range(24):
        c = rng.uniform([60, 60], [740, 540])
        ang = rng.uniform(-0.5, 0.5) if t % 3 else 0.0
        bw, bh = rng.uniform(30, 300), rng.uniform(10, 60)
        if t % 5 == 0:
            bw, bh = bh, bw * 1.2          # tall boxes -> rot90 branch
        if t % 7 == 0:
            c = np.array([5.0, 8.0])       # clipped by the border -> BORDER_REPLICATE
        R = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]])
        boxes.append((np.array([[-bw / 2, -bh / 2], [bw / 2, -bh / 2], [bw / 2, bh / 2], [-bw / 2, bh / 2]]) @ R.T + c))
    boxes = np.array(boxes, np.float32)
    crops = engine_fp32.crop_boxes(img, boxes)
    assert len(crops) == len(boxes)
    for b, c in zip(boxes, crops):
        ref = restate_pre.get_rotate_crop_image(img, b)
        assert c.shape == ref.shape
        assert np.array_equal(c, ref)

"""GPU parity on real photographs (four of the reference's bundled test images, tests/golden/real_images/).

The det / rec weights are the synthetic stand-ins, so the probability maps of photographs are rich in speckle, holes
and ragged blobs — exactly the inputs on which the integer post-processing has to agree with cv2 bit for bit."""
import glob
import os

import cv2
import numpy as np
import pytest

from onnxocr_b200 import _lib
from oracle import restate_post, restate_pre
from oracle.restate_pipeline import OraclePipeline

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
IMAGES = sorted(glob.glob(os.path.join(HERE, "golden", "real_images", "*.png")))


def _read(path):
    return cv2.imdecode(np.fromfile(path, dtype=np.uint8), cv2.IMREAD_COLOR)


@pytest.mark.parametrize("path", IMAGES, ids=[os.path.basename(p) for p in IMAGES])
def test_det_postprocess_and_crops_bit_exact_on_photos(engine_fp32, path):
    img = _read(path)
    assert img is not None
    boxes, maps = engine_fp32.det_pages([img], want_prob=True)
    x, shape = restate_pre.det_preprocess(img)
    assert maps[0].shape == x.shape[2:]
    # integer post-processing given the identical map: bitmap, components, boxes, filter — bit exact
    exp = restate_post.filter_tag_det_res(restate_post.db_postprocess(maps[0][None, None], shape[None])[0][0], img.shape)
    assert boxes[0].shape == exp.shape
    assert np.array_equal(boxes[0], exp)
    # the fused resize / normalise kernel against the cv2 restatement on a real resize (non-identity ratios)
    pre, _ = engine_fp32.det_preprocess(img)
    assert np.array_equal(pre, x)
    if len(exp):
        crops = engine_fp32.crop_boxes(img, boxes[0])
        for b, c in zip(boxes[0], crops):
            assert np.array_equal(c, restate_pre.get_rotate_crop_image(img, b))


def test_ocr_fp32_matches_oracle_on_photos(engine_fp32, model_paths):
    oracle = OraclePipeline(model_paths, use_angle_cls=True)
    charset = restate_post.load_charset(model_paths["dict"])
    params = _lib.ocr_params(use_angle_cls=True)
    pages = [_read(p) for p in IMAGES]
    got = engine_fp32.ocr_pages(pages, params)
    total = 0
    for img, g in zip(pages, got):
        oracle.drop_score = 0.0                      # compare every line, also the ones the reference would drop
        exp = oracle.ocr(img.copy())
        exp = exp[0]
        assert len(g["boxes"]) == len(exp)
        for k, (eb, (et, es)) in enumerate(exp):
            assert g["boxes"][k].tolist() == eb
            assert "".join(charset[i] for i in g["ids"][k]) == et
            assert abs(float(g["scores"][k]) - es) < 2e-4
        total += len(exp)
    assert total >= 1

"""Golden vectors produced by the UNMODIFIED reference (tests/golden/make_goldens.py imports it from /root/reference in the
authoring container; the reference's third-party engines onnxruntime / shapely / pyclipper are stood in for by the oracle's
restatements, see oracle/ref_harness.py).  The inputs are regenerated from seeds (tests/golden/golden_inputs.py).

  * not-gpu tests: the oracle's restatements reproduce the reference's outputs bit for bit — this is what pins the oracle
    on the GPU box, where /root/reference does not exist;
  * gpu tests: the CUDA path through the C ABI reproduces the same vectors.
"""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import golden_inputs as gi  # noqa: E402

from onnxocr_b200 import _lib  # noqa: E402
from oracle import restate_post, restate_pre  # noqa: E402


@pytest.fixture(scope="module")
def gold():
    z = np.load(os.path.join(HERE, "golden", "stage_goldens.npz"))
    with open(os.path.join(HERE, "golden", "e2e_goldens.json")) as fh:
        e = json.load(fh)
    return z, e


# ------------------------------------------------------------------------------------------------ oracle vs reference
def test_oracle_det_preprocess_matches_reference(gold):
    z, _ = gold
    for k, img in enumerate(gi.det_images()):
        x, shape = restate_pre.det_preprocess(img)
        assert np.array_equal(x[0], z["det_pre_%d" % k])
        assert np.array_equal(shape, z["det_shape_%d" % k])


def test_oracle_db_postprocess_matches_reference(gold):
    z, _ = gold
    for k, (prob, src) in enumerate(gi.prob_maps()):
        h, w = prob.shape
        shape = np.array([[float(src[0]), float(src[1]), h / src[0], w / src[1]]])
        boxes, _ = restate_post.db_postprocess(prob[None, None], shape)[0]
        assert np.array_equal(boxes, z["db_boxes_%d" % k])
        boxes, _ = restate_post.db_postprocess(prob[None, None], shape, box_thresh=0.5, score_mode="slow")[0]
        assert np.array_equal(boxes, z["db_boxes_slow_%d" % k])


def test_oracle_crops_match_reference(gold):
    z, _ = gold
    img, quads = gi.crop_case()
    for k, q in enumerate(quads):
        assert np.array_equal(restate_pre.get_rotate_crop_image(img, q.copy()), z["crop_%d" % k])


def test_oracle_ctc_decode_matches_reference(gold, model_paths):
    _, e = gold
    charset = restate_post.load_charset(model_paths["dict"])
    ids, scores, texts = restate_post.ctc_decode(gi.ctc_probs(), charset)
    assert [[t, float(s)] for t, s in zip(texts, scores)] == e["ctc"]


def test_oracle_pipeline_matches_reference(gold, model_paths):
    from oracle.restate_pipeline import OraclePipeline
    _, e = gold
    oracle = OraclePipeline(model_paths, use_angle_cls=True)
    for page, exp in zip(gi.e2e_pages(), e["ocr"]):
        got = oracle.ocr(page.copy())[0]
        assert [[b, [t, float(s)]] for b, (t, s) in got] == exp
    img, quads = gi.crop_case()
    crops = [restate_pre.get_rotate_crop_image(img, q.copy()) for q in quads]
    assert [[t, float(s)] for t, s in oracle.recognize(crops)] == e["rec_only"]


# ------------------------------------------------------------------------------------------------ CUDA path vs reference
@pytest.mark.gpu
def test_gpu_det_preprocess_matches_reference(gold, engine_fp32):
    z, _ = gold
    for k, img in enumerate(gi.det_images()):
        x, shape = engine_fp32.det_preprocess(img)
        assert np.array_equal(x[0], z["det_pre_%d" % k])
        assert np.array_equal(shape, z["det_shape_%d" % k])


@pytest.mark.gpu
def test_gpu_db_postprocess_matches_reference(gold, engine_fp32):
    z, _ = gold
    for k, (prob, src) in enumerate(gi.prob_maps()):
        boxes, _ = engine_fp32.db_postprocess(prob, src[0], src[1])
        assert np.array_equal(boxes, z["db_boxes_%d" % k])
        boxes, _ = engine_fp32.db_postprocess(prob, src[0], src[1], _lib.det_params(box_thresh=0.5, score_mode="slow"))
        assert np.array_equal(boxes, z["db_boxes_slow_%d" % k])


@pytest.mark.gpu
def test_gpu_crops_and_cls_match_reference(gold, engine_fp32):
    z, _ = gold
    img, quads = gi.crop_case()
    crops = engine_fp32.crop_boxes(img, np.stack(quads))
    for k, c in enumerate(crops):
        assert np.array_equal(c, z["crop_%d" % k])
    labels, scores = engine_fp32.cls_crops(crops)
    assert np.array_equal(labels, z["cls_labels"])
    assert np.abs(scores - z["cls_scores"]).max() < 1e-4


@pytest.mark.gpu
def test_gpu_end_to_end_matches_reference(gold, lib_built, model_paths):
    from onnxocr_b200 import pipeline
    _, e = gold
    ocr = pipeline.ONNXPaddleOcr(use_angle_cls=True, det_model_dir=model_paths["det"], rec_model_dir=model_paths["rec"],
                                 cls_model_dir=model_paths["cls"], rec_char_dict_path=model_paths["dict"], precision="fp32")
    for page, exp in zip(gi.e2e_pages(), e["ocr"]):
        got = ocr.ocr(page.copy())[0]
        assert len(got) == len(exp)
        for (gb, (gt, gs)), (eb, (et, es)) in zip(got, exp):
            assert gb == eb and gt == et
            assert abs(gs - es) < 2e-4
    img, quads = gi.crop_case()
    crops = [z for z in (np.load(os.path.join(HERE, "golden", "stage_goldens.npz"))["crop_%d" % k] for k in range(len(quads)))]
    got = ocr.text_recognizer(crops)
    assert [t for t, _ in got] == [t for t, _ in e["rec_only"]]
    assert np.allclose([s for _, s in got], [s for _, s in e["rec_only"]], atol=2e-4)

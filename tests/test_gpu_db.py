"""GPU parity of the device DBPostProcess and the det stage against the numpy/cv2 restatement
(bit-exact boxes given identical probability maps — BASELINE.json's integer post-processing bar)."""
import cv2
import numpy as np
import pytest

from onnxocr_b200 import _lib, synth_pages
from oracle import restate_post, restate_pre

pytestmark = pytest.mark.gpu


def blob_prob(rng, h, w, n_blobs, holes=False, noise=0.05):
    prob = np.zeros((h, w), np.float32)
    for _ in range(n_blobs):
        c = rng.uniform([20, 15], [w - 20, h - 15])
        ang = rng.uniform(-0.5, 0.5) if rng.random() < 0.7 else 0.0
        bw, bh = rng.uniform(15, w / 2), rng.uniform(5, 30)
        R = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]])
        pts = (np.array([[-bw / 2, -bh / 2], [bw / 2, -bh / 2], [bw / 2, bh / 2], [-bw / 2, bh / 2]]) @ R.T + c).astype(np.int32)
        cv2.fillPoly(prob, [pts], float(rng.uniform(0.45, 1.0)))
    prob = cv2.GaussianBlur(prob, (5, 5), 0) + rng.uniform(0, noise, (h, w)).astype(np.float32)
    if holes:
        for _ in range(n_blobs):
            x, y = int(rng.integers(2, w - 8)), int(rng.integers(2, h - 8))
            prob[y:y + int(rng.integers(1, 5)), x:x + int(rng.integers(1, 6))] = 0
    return prob.astype(np.float32)


def check(engine, prob, src_h, src_w, params=None, **kw):
    h, w = prob.shape
    exp_boxes, exp_scores = restate_post.db_postprocess(prob[None, None], np.array([[float(src_h), float(src_w), h / src_h, w / src_w]]), **kw)[0]
    boxes, scores = engine.db_postprocess(prob, src_h, src_w, params)
    assert boxes.shape[0] == len(exp_boxes), (boxes.shape, len(exp_boxes))
    if len(exp_boxes):
        assert np.array_equal(boxes, exp_boxes)
        assert np.allclose(scores, np.array(exp_scores), rtol=0, atol=1e-9)
    return len(exp_boxes)


@pytest.mark.parametrize("seed,hw,src,holes", [(0, (320, 480), (700, 1000), False), (1, (160, 288), (170, 296), True),
                                               (2, (960, 960), (960, 960), True), (3, (64, 64), (64, 64), False),
                                               (4, (704, 960), (1920, 2560), True), (5, (33, 47), (33, 47), True)])
def test_db_postprocess_blobs(engine_fp32, seed, hw, src, holes):
    rng = np.random.default_rng(seed)
    total = 0
    for _ in range(3):
        prob = blob_prob(rng, hw[0], hw[1], int(rng.integers(1, 30)), holes)
        total += check(engine_fp32, prob, src[0], src[1])
    assert total > 0 or hw[0] < 64


def test_db_postprocess_empty_and_full(engine_fp32):
    check(engine_fp32, np.zeros((64, 64), np.float32), 64, 64)           # the service's 64x64 warm-up image
    check(engine_fp32, np.ones((96, 128), np.float32), 300, 400)
    one = np.zeros((40, 40), np.float32)
    one[20, 20] = 1.0
    check(engine_fp32, one, 40, 40)


def test_db_postprocess_noise_and_candidate_cap(engine_fp32):
    rng = np.random.default_rng(11)
    prob = rng.random((200, 300)).astype(np.float32)                      # ~10^4 speckle contours -> first 1000 only
    check(engine_fp32, prob, 200, 300)
    prob = cv2.GaussianBlur(rng.random((256, 256)).astype(np.float32), (7, 7), 0) * 1.6
    n = check(engine_fp32, prob, 512, 512)
    assert n >= 1
    # > 16384 component roots on one page: the candidate buffer overflows and the batch is redone through the scan path
    prob = rng.random((512, 768)).astype(np.float32)
    assert len(cv2.findContours((prob > 0.3).astype(np.uint8), cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)[0]) > 20000
    check(engine_fp32, prob, 512, 768)
    # the same speckle page next to ordinary pages in one det batch (only the speckle page overflows)


def test_db_postprocess_params(engine_fp32):
    rng = np.random.default_rng(12)
    prob = blob_prob(rng, 320, 320, 14, True)
    check(engine_fp32, prob, 640, 640, _lib.det_params(thresh=0.5, box_thresh=0.7, unclip_ratio=2.0),
          thresh=0.5, box_thresh=0.7, unclip_ratio=2.0)
    check(engine_fp32, prob, 640, 640, _lib.det_params(max_candidates=5), max_candidates=5)
    check(engine_fp32, prob, 640, 640, _lib.det_params(use_dilation=True), use_dilation=True)


@pytest.mark.parametrize("seed,hw,src,holes", [(20, (320, 480), (700, 1000), True), (21, (160, 288), (170, 296), True),
                                               (22, (960, 960), (960, 960), True), (23, (33, 47), (33, 47), True)])
def test_db_postprocess_score_mode_slow(engine_fp32, seed, hw, src, holes):
    """det_db_score_mode='slow' (db_postprocess.py:129-132, 199-218): mean over cv2.fillPoly of the contour itself."""
    rng = np.random.default_rng(seed)
    slow = _lib.det_params(score_mode="slow")
    total = differs = 0
    for _ in range(3):
        prob = blob_prob(rng, hw[0], hw[1], int(rng.integers(1, 30)), holes)
        total += check(engine_fp32, prob, src[0], src[1], slow, score_mode="slow")
        _, s_fast = engine_fp32.db_postprocess(prob, src[0], src[1])
        _, s_slow = engine_fp32.db_postprocess(prob, src[0], src[1], slow)
        differs += len(s_fast) != len(s_slow) or not np.array_equal(s_fast, s_slow)
    assert total > 0 or hw[0] < 64
    assert differs > 0 or hw[0] < 64                                       # the option really selects another score


def test_db_postprocess_score_mode_slow_hard_maps(engine_fp32):
    rng = np.random.default_rng(24)
    slow = _lib.det_params(score_mode="slow", box_thresh=0.3)
    # blurred noise: ragged contours with holes and islands; a low box threshold keeps many of them
    prob = cv2.GaussianBlur(rng.random((256, 256)).astype(np.float32), (7, 7), 0) * 1.6
    assert check(engine_fp32, prob, 512, 512, slow, score_mode="slow", box_thresh=0.3) >= 1
    # speckle: ~10^4 contours (first 1000 only), one percolating component with a contour far beyond the warp kernel's
    # shared-memory capacity, more than 16384 roots on the page (candidate fallback) - all before the slow scores
    prob = rng.random((200, 300)).astype(np.float32)
    check(engine_fp32, prob, 200, 300, slow, score_mode="slow", box_thresh=0.3)
    prob = rng.random((512, 768)).astype(np.float32)
    check(engine_fp32, prob, 512, 768, slow, score_mode="slow", box_thresh=0.3)
    check(engine_fp32, np.zeros((64, 64), np.float32), 64, 64, slow, score_mode="slow", box_thresh=0.3)
    check(engine_fp32, np.ones((96, 128), np.float32), 300, 400, slow, score_mode="slow", box_thresh=0.3)
    # dilation + slow scores, candidate cap
    prob = blob_prob(rng, 320, 320, 14, True)
    check(engine_fp32, prob, 640, 640, _lib.det_params(score_mode="slow", use_dilation=True), score_mode="slow", use_dilation=True)
    check(engine_fp32, prob, 640, 640, _lib.det_params(score_mode="slow", max_candidates=5), score_mode="slow", max_candidates=5)


def test_det_pages_score_mode_slow_through_text_detector(engine_fp32):
    """The option as a user sets it (TextDetector(args) with det_db_score_mode='slow'), batch of pages, streaming path."""
    pages = [synth_pages.render_page(640, 480, 14, seed=50)[0], synth_pages.render_page(640, 480, 10, seed=51)[0]]
    slow = _lib.det_params(score_mode="slow")
    boxes, maps = engine_fp32.det_pages(pages, slow, want_prob=True)
    n_total = 0
    for page, b, m in zip(pages, boxes, maps):
        h, w = m.shape
        exp, _ = restate_post.db_postprocess(m[None, None], np.array([[float(page.shape[0]), float(page.shape[1]), h / page.shape[0], w / page.shape[1]]]),
                                             score_mode="slow")[0]
        exp = restate_post.filter_tag_det_res(exp, page.shape)
        assert b.shape[0] == len(exp)
        if len(exp):
            assert np.array_equal(b, exp)
        n_total += len(exp)
    assert n_total > 10
    got = list(engine_fp32.det_pages_stream([pages, pages], slow))
    for bb in got:
        for b, e in zip(bb, boxes):
            assert np.array_equal(b, e)


def _oracle_det(prob, page_shape):
    h, w = prob.shape
    src_h, src_w = page_shape[:2]
    boxes, _ = restate_post.db_postprocess(prob[None, None], np.array([[float(src_h), float(src_w), h / src_h, w / src_w]]))[0]
    return restate_post.filter_tag_det_res(boxes, page_shape)


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_det_pages_boxes_bit_exact_given_maps(engine_fp32, engine_bf16, prec):
    eng = engine_fp32 if prec == "fp32" else engine_bf16
    pages = [synth_pages.render_page(640, 480, 14, seed=40)[0], synth_pages.render_page(640, 480, 10, seed=41)[0],
             synth_pages.render_page(1280, 720, 20, seed=42)[0], synth_pages.render_page(50, 60, 1, seed=43, min_px=10, max_px=12)[0]]
    boxes, maps = eng.det_pages(pages, want_prob=True)
    n_total = 0
    for page, b, m in zip(pages, boxes, maps):
        exp = _oracle_det(m, page.shape)
        assert b.shape[0] == len(exp)
        if len(exp):
            assert np.array_equal(b, exp)
        n_total += len(exp)
    assert n_total > 20


def test_det_pages_match_session_path(engine_fp32, torch_graphs):
    """Level B (fused) == Level A (reference pipeline around session.run) on the same page."""
    page = synth_pages.render_page(480, 640, 10, seed=44)[0]
    boxes, maps = engine_fp32.det_pages([page], want_prob=True)
    x, shape = restate_pre.det_preprocess(page)
    ref = torch_graphs("det").run(None, {"x": x})[0]
    assert np.abs(maps[0] - ref[0, 0]).max() < 1e-4
    exp = restate_post.filter_tag_det_res(restate_post.db_postprocess(ref, shape[None])[0][0], page.shape)
    # the maps agree to 1e-4, the boxes must agree unless a pixel sits within 1e-4 of the threshold
    if np.abs(ref[0, 0] - 0.3).min() > 2e-4:
        assert np.array_equal(boxes[0], exp)

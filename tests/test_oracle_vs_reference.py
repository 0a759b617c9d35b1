"""This container only: the restatements in oracle/ against the UNMODIFIED reference imported from /root/reference
(onnxruntime -> torch interpreter, shapely/pyclipper -> clipper_offset shims).  Skipped where the mount is absent."""
import cv2
import numpy as np
import pytest

from oracle import ref_harness, restate_post, restate_pre

pytestmark = [pytest.mark.reference, pytest.mark.skipif(not ref_harness.available(), reason="/root/reference not mounted")]


@pytest.fixture(scope="module")
def ref():
    return ref_harness.import_reference()


def _imgs():
    import glob
    files = sorted(glob.glob(ref_harness.REFERENCE + "/onnxocr/test_images/*.jpg"))[:6]
    return [cv2.imread(f) for f in files]


def test_det_preprocess_matches_reference(ref):
    from onnxocr.imaug import create_operators, transform
    ops = create_operators([{"DetResizeForTest": {"limit_side_len": 960, "limit_type": "max"}},
                            {"NormalizeImage": {"std": [0.229, 0.224, 0.225], "mean": [0.485, 0.456, 0.406], "scale": "1./255.", "order": "hwc"}},
                            {"ToCHWImage": None}, {"KeepKeys": {"keep_keys": ["image", "shape"]}}])
    rng = np.random.default_rng(0)
    imgs = _imgs() + [rng.integers(0, 256, (20, 30, 3), dtype=np.uint8), rng.integers(0, 256, (333, 777, 3), dtype=np.uint8)]
    for img in imgs:
        exp_img, exp_shape = transform({"image": img}, ops)
        got_img, got_shape = restate_pre.det_preprocess(img)
        assert np.array_equal(got_img[0], exp_img)
        assert np.array_equal(got_shape, exp_shape)


def test_crop_and_rec_cls_preprocess_match_reference(ref, model_paths):
    from onnxocr.utils import get_rotate_crop_image, infer_args
    from onnxocr.predict_rec import TextRecognizer
    from onnxocr.predict_cls import TextClassifier
    import argparse
    args = argparse.Namespace(**{a.dest: a.default for a in infer_args()._actions})
    args.rec_image_shape = "3, 48, 320"
    args.rec_model_dir, args.cls_model_dir, args.det_model_dir = model_paths["rec"], model_paths["cls"], model_paths["det"]
    args.rec_char_dict_path = model_paths["dict"]
    args.use_gpu = False
    rec, cls = TextRecognizer(args), TextClassifier(args)
    rng = np.random.default_rng(1)
    img = _imgs()[0]
    h, w = img.shape[:2]
    for t in range(12):
        c = rng.uniform([80, 60], [w - 80, h - 60])
        ang = rng.uniform(-0.4, 0.4)
        bw, bh = rng.uniform(30, 300), rng.uniform(10, 60)
        if t % 4 == 0:
            bw, bh = bh, bw * 1.5
        R = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]])
        pts = (np.array([[-bw / 2, -bh / 2], [bw / 2, -bh / 2], [bw / 2, bh / 2], [-bw / 2, bh / 2]]) @ R.T + c).astype(np.float32)
        exp = get_rotate_crop_image(img, pts.copy())
        got = restate_pre.get_rotate_crop_image(img, pts.copy())
        assert np.array_equal(exp, got)
        ratio = max(320 / 48.0, got.shape[1] / got.shape[0])
        assert np.array_equal(rec.resize_norm_img(got, ratio), restate_pre.resize_norm_img(got, int(48 * ratio)))
        assert np.array_equal(cls.resize_norm_img(got), restate_pre.resize_norm_img(got, 192))


@pytest.mark.parametrize("score_mode", ["fast", "slow"])
def test_db_postprocess_matches_reference(ref, score_mode):
    from onnxocr.db_postprocess import DBPostProcess
    from onnxocr.predict_system import sorted_boxes
    rng = np.random.default_rng(0)
    pp = DBPostProcess(thresh=0.3, box_thresh=0.6, max_candidates=1000, unclip_ratio=1.5, score_mode=score_mode)
    n = 0
    for t in range(8):
        H, W = 320, 480
        prob = np.zeros((H, W), np.float32)
        for k in range(14):
            c = rng.uniform([40, 30], [W - 40, H - 30])
            ang = rng.uniform(-0.4, 0.4)
            bw, bh = rng.uniform(20, 150), rng.uniform(6, 24)
            R = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]])
            pts = (np.array([[-bw / 2, -bh / 2], [bw / 2, -bh / 2], [bw / 2, bh / 2], [-bw / 2, bh / 2]]) @ R.T + c).astype(np.int32)
            cv2.fillPoly(prob, [pts], float(rng.uniform(0.5, 1.0)))
        prob = cv2.GaussianBlur(prob, (5, 5), 0) + rng.uniform(0, 0.05, (H, W)).astype(np.float32)
        if t % 2:
            for k in range(6):
                x, y = int(rng.integers(10, W - 10)), int(rng.integers(10, H - 10))
                prob[y:y + 3, x:x + 4] = 0
        shape_list = np.array([[700.0, 1000.0, H / 700.0, W / 1000.0]])
        exp = pp({"maps": prob[None, None]}, shape_list)[0]["points"]
        got, _ = restate_post.db_postprocess(prob[None, None], shape_list, score_mode=score_mode)[0]
        assert exp.shape == got.shape and np.array_equal(exp, got)
        if len(exp):
            a = sorted_boxes(exp.astype(np.float32))
            b = restate_post.sorted_boxes(got.astype(np.float32))
            assert all(np.array_equal(x, y) for x, y in zip(a, b))
        n += len(exp)
    assert n > 20


def test_ctc_decode_matches_reference(ref, model_paths):
    from onnxocr.rec_postprocess import CTCLabelDecode
    dec = CTCLabelDecode(character_dict_path=model_paths["dict"], use_space_char=True)
    assert len(dec.character) == 18385
    rng = np.random.default_rng(3)
    logits = rng.standard_normal((5, 40, 18385)).astype(np.float32) * 3
    logits[:, ::3, 0] += 30          # blanks
    logits[2, :, 0] += 100           # an all-blank line -> ("", 0.0)
    p = np.exp(logits - logits.max(axis=2, keepdims=True))
    p /= p.sum(axis=2, keepdims=True)
    exp = dec(p)
    ids, scores, texts = restate_post.ctc_decode(p, restate_post.load_charset(model_paths["dict"]))
    assert [e[0] for e in exp] == texts
    assert [e[1] for e in exp] == scores


def test_end_to_end_matches_reference(ref, model_paths):
    """ONNXPaddleOcr(use_angle_cls=True).ocr(img) of the reference (sessions = torch interpreter) == OraclePipeline.ocr."""
    from onnxocr_b200 import synth_pages
    from oracle.restate_pipeline import OraclePipeline
    model = ref_harness.make_ocr(model_paths)
    mine = OraclePipeline(model_paths, use_angle_cls=True)
    for seed in (60, 61):
        page = synth_pages.render_page(480, 640, 9, seed=seed, rot180_frac=0.3)[0]
        exp = model.ocr(page.copy())
        got = mine.ocr(page.copy())
        assert len(exp[0]) == len(got[0]) and len(exp[0]) > 0
        for (eb, (et, es)), (gb, (gt, gs)) in zip(exp[0], got[0]):
            assert eb == gb
            assert et == gt and es == gs

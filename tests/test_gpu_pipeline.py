"""GPU: the Level-B stage mirrors (onnxocr_b200/pipeline.py) against the oracle pipeline (predict_system.py:35-72
restated) on synthetic pages.  fp32 mode: boxes and strings identical; bf16 mode: >= 99.5 % identical text lines
is BASELINE.json's bar — asserted on the lines both pipelines detect."""
import numpy as np
import pytest

from onnxocr_b200 import _lib, pipeline, synth_pages
from oracle import restate_post, restate_pre
from oracle.restate_pipeline import OraclePipeline

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def oracle(model_paths):
    return OraclePipeline(model_paths, use_angle_cls=True)


@pytest.fixture(scope="module")
def system_fp32(lib_built, model_paths):
    return pipeline.ONNXPaddleOcr(use_angle_cls=True, det_model_dir=model_paths["det"], rec_model_dir=model_paths["rec"],
                                  cls_model_dir=model_paths["cls"], rec_char_dict_path=model_paths["dict"], precision="fp32")


@pytest.fixture(scope="module")
def system_bf16(lib_built, model_paths):
    return pipeline.ONNXPaddleOcr(use_angle_cls=True, det_model_dir=model_paths["det"], rec_model_dir=model_paths["rec"],
                                  cls_model_dir=model_paths["cls"], rec_char_dict_path=model_paths["dict"], precision="bf16")


def test_rec_crops_match_oracle(system_fp32, oracle):
    crops = synth_pages.rec_crops_c3(9, seed=21) + synth_pages.rec_crops_c3(4, seed=22, w=700) + synth_pages.rec_crops_c3(3, seed=23, w=130)
    exp = oracle.recognize(crops)
    got = system_fp32.text_recognizer(crops)
    assert [t for t, _ in got] == [t for t, _ in exp]
    assert np.allclose([s for _, s in got], [s for _, s in exp], atol=2e-4)


def test_cls_crops_match_oracle(system_fp32, oracle):
    import cv2
    crops = synth_pages.rec_crops_c3(8, seed=24, w=260)
    crops = [cv2.rotate(c, 1) if i % 2 else c for i, c in enumerate(crops)]
    exp_imgs, exp = oracle.classify(crops)
    got_imgs, got = system_fp32.text_classifier(crops)
    assert [r[0] for r in got] == [r[0] for r in exp]
    assert np.allclose([r[1] for r in got], [r[1] for r in exp], atol=1e-4)
    for a, b in zip(got_imgs, exp_imgs):
        assert np.array_equal(a, b)


def test_ocr_fp32_matches_oracle(system_fp32, oracle):
    n_lines = 0
    for seed in (70, 71, 72):
        page = synth_pages.render_page(480, 640, 9, seed=seed, rot180_frac=0.3)[0]
        exp = oracle.ocr(page.copy())[0]
        got = system_fp32.ocr(page.copy())[0]
        assert len(got) == len(exp)
        for (gb, (gt, gs)), (eb, (et, es)) in zip(got, exp):
            assert gb == eb
            assert gt == et
            assert abs(gs - es) < 2e-4
        n_lines += len(exp)
    assert n_lines >= 12


def test_ocr_staged_equals_fused(system_fp32, model_paths):
    """The stage-by-stage path (reference call sequence) and the fused device pipeline give identical results."""
    staged = pipeline.ONNXPaddleOcr(use_angle_cls=True, det_model_dir=model_paths["det"], rec_model_dir=model_paths["rec"],
                                    cls_model_dir=model_paths["cls"], rec_char_dict_path=model_paths["dict"], precision="fp32",
                                    fused=False)
    staged.engine = system_fp32.engine
    for seed in (73, 74):
        page = synth_pages.render_page(512, 704, 11, seed=seed, rot180_frac=0.3)[0]
        a = staged.ocr(page.copy())[0]
        b = system_fp32.ocr(page.copy())[0]
        assert len(a) == len(b) and len(a) >= 5
        for (ab, (at, asc)), (bb, (bt, bsc)) in zip(a, b):
            assert ab == bb and at == bt and abs(asc - bsc) < 1e-6


def test_ocr_batch_mixed_pages_matches_single(system_fp32):
    """Level C: many pages of different sizes in one call == one call per page (page order kept, empty page included)."""
    pages = [synth_pages.render_page(480, 640, 8, seed=90)[0], np.zeros((64, 64, 3), np.uint8),
             synth_pages.render_page(384, 800, 6, seed=91, rot180_frac=0.5)[0], synth_pages.render_page(480, 640, 7, seed=92)[0]]
    got = system_fp32.ocr_batch(pages)
    assert len(got) == len(pages) and got[1] == []
    for page, g in zip(pages, got):
        single = system_fp32.ocr(page)[0]
        assert [x[0] for x in g] == [x[0] for x in single]
        assert [x[1][0] for x in g] == [x[1][0] for x in single]
        assert np.allclose([x[1][1] for x in g], [x[1][1] for x in single], atol=1e-6)


def test_ocr_warmup_image_has_no_boxes(system_fp32):
    """app/engine.py:149-164 warms the model with a 64x64 zero image."""
    assert system_fp32.ocr(np.zeros((64, 64, 3), np.uint8)) == [[]]


def test_ocr_bf16_text_lines(system_bf16, oracle):
    same = total = 0
    for seed in (80, 81, 82, 83):
        page = synth_pages.render_page(640, 800, 14, seed=seed)[0]
        exp = oracle.ocr(page.copy())[0]
        got = system_bf16.ocr(page.copy())[0]
        exp_t = [t for _, (t, _) in exp]
        got_t = [t for _, (t, _) in got]
        total += len(exp_t)
        same += sum(1 for a, b in zip(got_t, exp_t) if a == b) if len(got_t) == len(exp_t) else len(set(got_t) & set(exp_t))
    assert total >= 30
    assert same / total >= 0.9, (same, total)


def test_c5_a4_300dpi_det_matches_oracle(lib_built, model_paths):
    """BASELINE config 5: A4 @ 300 dpi (3508 x 2480) with det_limit_side_len 2560 -> det input 2560 x 1824 (exercises
    the real resize path of the fused pre-processing kernel and the large-map post-processing)."""
    page = synth_pages.render_page(3508, 2480, 24, seed=300, min_px=40, max_px=64)[0]
    oracle = OraclePipeline(model_paths, use_angle_cls=False, det_limit_side_len=2560)
    exp, exp_map = oracle.detect(page, return_map=True)
    eng = _lib.Engine(0)
    eng.load_onnx(_lib.DET, model_paths["det"], _lib.FP32)
    params = _lib.det_params(limit_side_len=2560)
    boxes, maps = eng.det_pages([page], params, want_prob=True)
    assert maps[0].shape == (2560, 1824) == exp_map.shape
    assert np.abs(maps[0] - exp_map).max() < 1e-4
    assert len(exp) >= 10
    assert boxes[0].shape == exp.shape and np.array_equal(boxes[0], exp)
    eng.close()


def test_ocr_batch_bf16_c5_and_tiny_pages(system_bf16):
    """Level C on pages of very different sizes in one call: an A4 @ 300 dpi page, a 20 x 30 page (h + w < 64: the
    reference pads it to 32 x 32, operators.py:72-76) and a blank page."""
    big = synth_pages.render_page(3508, 2480, 12, seed=301, min_px=40, max_px=64)[0]
    tiny = np.full((20, 30, 3), 255, np.uint8)
    blank = np.full((700, 500, 3), 255, np.uint8)
    res = system_bf16.ocr_batch([big, tiny, blank])
    assert len(res) == 3 and res[1] == [] and res[2] == []
    assert len(res[0]) >= 8
    assert all(len(t) > 0 and 0.5 <= s <= 1.0 for _, (t, s) in res[0])


def test_ocr_pages_line_capacity_error(system_fp32):
    """More lines than the caller's buffers hold is reported as an error, never truncated silently."""
    page = synth_pages.render_page(480, 640, 9, seed=70)[0]
    with pytest.raises(RuntimeError, match="line capacity too small"):
        system_fp32.engine.ocr_pages([page], system_fp32._ocr_params(True), line_cap=2)


def test_concurrent_callers_share_one_engine(system_fp32):
    """The reference calls `model.ocr` on ONE model object from a ThreadPoolExecutor(max_workers=4)
    (ocr_images_pdfs.py:106,154) and from the service's executor (app/engine.py:111-120): concurrent calls must be
    safe and give the same results as sequential ones."""
    from concurrent.futures import ThreadPoolExecutor
    pages = [synth_pages.render_page(384, 512, 6, seed=400 + i, rot180_frac=0.2)[0] for i in range(8)]
    expected = [system_fp32.ocr(p)[0] for p in pages]
    with ThreadPoolExecutor(max_workers=4) as ex:
        got = list(ex.map(lambda p: system_fp32.ocr(p)[0], pages))
    assert got == expected
    assert sum(len(r) for r in expected) >= 24


def test_micro_batcher_equals_per_page_calls(system_fp32):
    """Cross-request batching: 8 threads calling `ocr(img)` through the batcher get exactly the
    per-page results, from fewer library calls."""
    from concurrent.futures import ThreadPoolExecutor
    from onnxocr_b200.batcher import MicroBatcher, decode_pages
    import cv2
    pages = [synth_pages.render_page(384 + 32 * (i % 3), 512, 5 + i % 4, seed=500 + i, rot180_frac=0.2)[0] for i in range(12)]
    expected = [system_fp32.ocr(p) for p in pages]
    mb = MicroBatcher(system_fp32, max_batch=8, max_wait_ms=20)
    with ThreadPoolExecutor(max_workers=8) as ex:
        got = list(ex.map(mb.ocr, pages))
    mb.close()
    assert got == expected
    assert mb.pages == 12 and mb.batches < 12
    # page feed: encoded bytes -> pinned buffers -> one batched call
    blobs = [cv2.imencode(".png", p)[1].tobytes() for p in pages[:4]]
    decoded = decode_pages(blobs, engine=system_fp32.engine, workers=4)
    assert all(np.array_equal(d, p) for d, p in zip(decoded, pages[:4]))
    assert system_fp32.ocr_batch(decoded) == [e[0] for e in expected[:4]]


def test_bf16_results_do_not_depend_on_call_composition(lib_built, model_paths):
    """The same page must give bit-identical boxes, ids and scores whatever else shares its call — including the very
    first use of every launch plan (a regression test: per-plan tables uploaded on the legacy stream were once still in
    flight at a plan's first launch, and narrow layers were pixel-packed depending on the batch total)."""
    ocr = pipeline.ONNXPaddleOcr(use_angle_cls=True, det_model_dir=model_paths["det"], rec_model_dir=model_paths["rec"],
                                 cls_model_dir=model_paths["cls"], rec_char_dict_path=model_paths["dict"], precision="bf16")
    ocr.args._b200_engine = None
    pages = synth_pages.e2e_pages_c4(20, seed=11)
    P = ocr._ocr_params(True)

    def raw(pgs):
        return [(r["boxes"].tolist(), [i.tolist() for i in r["ids"]], r["scores"].tolist(), r["cls_labels"].tolist())
                for r in ocr.engine.ocr_pages(pgs, P)]

    ref = raw(pages)
    assert sum(len(r[2]) for r in ref) > 200
    for k in (2, 3, 7):
        got = [None] * len(pages)
        for j in range(k):
            got[j::k] = raw(pages[j::k])
        assert got == ref, "results changed when the pages were split over %d calls" % k
    assert raw(pages) == ref


def test_plan_cache_eviction_keeps_results(system_bf16):
    """More distinct (batch, width) shapes than the plan cache holds (768): plans are evicted and rebuilt; results for a
    shape seen before the eviction must come back unchanged."""
    eng = system_bf16.engine
    base = synth_pages.rec_crops_c3(1, seed=31, w=2000)[0]
    first = {}
    for w in range(330, 1150):                      # 820 distinct padded widths, one crop each
        ids, scores = eng.rec_crops([base[:, :w].copy()])
        if w < 340:
            first[w] = (ids[0].tolist(), float(scores[0]))
    for w, (ids0, s0) in first.items():
        ids, scores = eng.rec_crops([base[:, :w].copy()])
        assert ids[0].tolist() == ids0 and float(scores[0]) == s0

"""CPU: the cross-request micro-batcher's host logic with a stand-in model (no GPU): batching, ordering, error fan-out."""
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from onnxocr_b200.batcher import MicroBatcher, decode_pages


class FakeModel(object):
    def __init__(self, delay=0.01, fail_on=None):
        self.calls = []
        self.delay = delay
        self.fail_on = fail_on

    def ocr_batch(self, imgs, cls=True):
        self.calls.append(len(imgs))
        time.sleep(self.delay)
        if self.fail_on is not None and any(int(im[0, 0, 0]) == self.fail_on for im in imgs):
            raise RuntimeError("boom")
        return [[[[0, 0], [1, 0], [1, 1], [0, 1]], ("page%d" % int(im[0, 0, 0]), 0.9)] for im in imgs]

    def ocr(self, img, det=True, rec=True, cls=True):
        return ["partial"]


def test_concurrent_callers_are_batched_and_get_their_own_result():
    model = FakeModel()
    mb = MicroBatcher(model, max_batch=8, max_wait_ms=20)
    imgs = [np.full((4, 4, 3), i, np.uint8) for i in range(24)]
    with ThreadPoolExecutor(max_workers=12) as ex:
        got = list(ex.map(mb.ocr, imgs))
    mb.close()
    assert [g[0][1][0] for g in got] == ["page%d" % i for i in range(24)]
    assert sum(model.calls) == 24 and max(model.calls) <= 8
    assert len(model.calls) < 24                     # at least some requests shared a call


def test_one_bad_page_fails_only_its_own_request_and_batcher_survives():
    """pages of unrelated requests share a call; a failing page must not fail its co-batched neighbours
    (the reference's per-image model.ocr fails only the offending request)."""
    model = FakeModel(fail_on=3)
    mb = MicroBatcher(model, max_batch=4, max_wait_ms=50)
    futs = [mb.submit(np.full((2, 2, 3), i, np.uint8)) for i in range(4)]
    for i, f in enumerate(futs):
        if i == 3:
            with pytest.raises(RuntimeError, match="boom"):
                f.result(timeout=10)
        else:
            assert f.result(timeout=10)[0][1][0] == "page%d" % i
    ok = mb.ocr(np.full((2, 2, 3), 7, np.uint8))
    assert ok[0][1][0] == "page7"
    assert mb.ocr(np.zeros((2, 2, 3), np.uint8), det=True, rec=False) == ["partial"]
    mb.close()


def test_invalid_inputs_are_rejected_before_they_are_enqueued():
    model = FakeModel()
    mb = MicroBatcher(model, max_batch=4, max_wait_ms=5)
    for bad, exc in ((None, ValueError), (np.zeros((4, 4), np.uint8), ValueError), (np.zeros((4, 4, 4), np.uint8), ValueError),
                     (np.zeros((4, 4, 3), np.float32), ValueError), ([[1, 2, 3]], TypeError), (np.zeros((0, 4, 3), np.uint8), ValueError)):
        with pytest.raises(exc):
            mb.ocr(bad)
        with pytest.raises(exc):
            mb.submit(bad)
    assert model.calls == []                          # nothing reached the model
    assert mb.ocr(np.full((2, 2, 3), 5, np.uint8))[0][1][0] == "page5"
    mb.close()


def test_decode_pages_order_and_failures(tmp_path):
    import cv2
    paths = []
    for i in range(5):
        p = tmp_path / ("im%d.png" % i)
        cv2.imwrite(str(p), np.full((8 + i, 6, 3), 10 * i, np.uint8))
        paths.append(str(p))
    blob = open(paths[2], "rb").read()
    out = decode_pages(paths + [blob, b"not an image"], workers=3)
    assert [o.shape for o in out[:5]] == [(8 + i, 6, 3) for i in range(5)]
    assert np.array_equal(out[5], out[2]) and out[6] is None

"""CPU: host-side logic of the streaming det generator (Engine.det_pages_stream) with the C calls replaced by a recorder:
batches are yielded in order, never more than `depth` (<= 4) are in flight, everything submitted is collected."""
import os

from onnxocr_b200 import _lib


class FakeEngine(object):
    det_pages_stream = _lib.Engine.det_pages_stream

    def __init__(self):
        self.in_flight = []
        self.max_in_flight = 0
        self.log = []

    def det_submit(self, pages, params=None):
        self.in_flight.append(pages)
        self.max_in_flight = max(self.max_in_flight, len(self.in_flight))
        self.log.append(("submit", pages))
        return (len(self.log), pages)

    def det_collect(self, ticket, max_boxes=1000):
        assert self.in_flight and self.in_flight[0] is ticket[1], "collected out of order"
        self.in_flight.pop(0)
        self.log.append(("collect", ticket[1]))
        return ["boxes of %s" % ticket[1]]


def test_stream_generator_order_and_depth(monkeypatch):
    batches = ["b%d" % i for i in range(7)]
    for depth, expect in ((1, 1), (2, 2), (4, 4), (9, 4), (None, 4)):
        monkeypatch.delenv("B200OCR_DET_DEPTH", raising=False)
        eng = FakeEngine()
        out = list(eng.det_pages_stream(iter(batches), depth=depth))
        assert out == [["boxes of %s" % b] for b in batches]
        assert eng.max_in_flight == expect and not eng.in_flight
    monkeypatch.setenv("B200OCR_DET_DEPTH", "3")
    eng = FakeEngine()
    assert len(list(eng.det_pages_stream(batches))) == 7 and eng.max_in_flight == 3
    eng = FakeEngine()
    assert list(eng.det_pages_stream([])) == [] and eng.max_in_flight == 0


def test_header_declares_the_streaming_entry_points():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "b200ocr.h")).read()
    for sym in ("b200ocr_det_submit", "b200ocr_det_collect", "b200ocr_det_resident_stream", "b200ocr_trace_dump"):
        assert sym in hdr
        assert hasattr(_lib.lib(), sym), sym + " is not exported by libb200ocr.so"

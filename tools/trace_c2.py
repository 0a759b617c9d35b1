"""Coarse device timeline of the streaming C2 benchmark (det network / DBPostProcess per lane).
    B200OCR_TRACE=/tmp/t.jsonl python tools/trace_c2.py [steps]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
path = os.environ.setdefault("B200OCR_TRACE", "/tmp/b200ocr_trace.jsonl")
from onnxocr_b200 import _lib, synth_graphs, synth_pages  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
eng = _lib.Engine(0)
eng.load_onnx(_lib.DET, synth_graphs.ensure_models()["det"], _lib.BF16)
pages = synth_pages.det_pages_c2(64, 960, 40, seed=0)
eng.det_resident_upload(pages)
eng.det_resident_stream(None, 8, 16)
_lib.lib().b200ocr_trace_dump()
open(path, "w").close()
ms, nb = eng.det_resident_stream(None, steps, 16)
_lib.lib().b200ocr_trace_dump()
recs = [json.loads(l) for l in open(path)]
print("steps %d: %.3f ms (%.3f per step), %d boxes" % (steps, ms, ms / steps, nb))
for r in sorted(recs, key=lambda r: r["t0"]):
    print("%8.3f %8.3f  (%6.3f)  s%d %-9s %d" % (r["t0"], r["t1"], r["t1"] - r["t0"], r["stream"], r["name"], r["tag"]))

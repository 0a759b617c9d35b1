"""Coarse device timeline of one C4 batch (det lanes, DBPostProcess, warp, cls, rec batches per stream).
    B200OCR_TRACE=/tmp/t.jsonl python tools/trace_c4.py [pages] [bucket] > summary"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
path = os.environ.setdefault("B200OCR_TRACE", "/tmp/b200ocr_trace.jsonl")
from onnxocr_b200 import _lib, pipeline, synth_pages  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ocr = pipeline.ONNXPaddleOcr(use_angle_cls=True, precision="bf16", rec_width_bucket=bucket)
pages = synth_pages.e2e_pages_c4(n, seed=2)
for _ in range(3):
    ocr.ocr_batch(pages)
_lib.lib().b200ocr_trace_dump()
open(path, "w").close()
ocr.ocr_batch(pages)
_lib.lib().b200ocr_trace_dump()
recs = [json.loads(l) for l in open(path)]
t_end = max(r["t1"] for r in recs)
print("scopes %d, span %.3f ms" % (len(recs), t_end))
by = {}
for r in recs:
    by.setdefault(r["name"], []).append(r)
for k, v in by.items():
    busy = sum(r["t1"] - r["t0"] for r in v)
    print("%-10s n=%3d  first start %.3f  last end %.3f  sum of durations %.3f ms  mean %.3f" % (
        k, len(v), min(r["t0"] for r in v), max(r["t1"] for r in v), busy, busy / len(v)))
print("per stream:")
st = {}
for r in recs:
    st.setdefault(r["stream"], []).append(r)
for k in sorted(st):
    v = st[k]
    print("  stream %d: %d scopes, %.3f .. %.3f, busy %.3f ms  %s" % (k, len(v), min(r["t0"] for r in v), max(r["t1"] for r in v),
                                                                  sum(r["t1"] - r["t0"] for r in v), sorted({r["name"] for r in v})))
for r in sorted(recs, key=lambda r: r["t0"]):
    print("%8.3f %8.3f  s%d %-9s %d" % (r["t0"], r["t1"], r["stream"], r["name"], r["tag"]))

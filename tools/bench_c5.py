"""C5: A4 @ 300 dpi pages (3508 x 2480), det_limit_side_len 2560, end to end: python tools/bench_c5.py [pages] [bf16|fp32]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onnxocr_b200 import pipeline, synth_pages  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
ocr = pipeline.ONNXPaddleOcr(use_angle_cls=True, precision=prec, det_limit_side_len=2560)
pages = synth_pages.pdf_pages_c5(n, seed=3)
for _ in range(2):
    res = ocr.ocr_batch(pages)
t0 = time.perf_counter()
reps = 3
for _ in range(reps):
    res = ocr.ocr_batch(pages)
dt = (time.perf_counter() - t0) / reps
lines = sum(len(r) for r in res)
ocr.text_detector.batch(pages)                     # builds the 8-page det plan (the waves above use 4-page plans)
t1 = time.perf_counter()
for _ in range(reps):
    det_only = ocr.text_detector.batch(pages)
dt_det = (time.perf_counter() - t1) / reps
print(json.dumps({"config": "C5 A4@300dpi (3508x2480) pages, det_limit_side_len 2560, det+cls+rec", "precision": prec, "pages": n,
                  "pages_per_s": n / dt, "lines_per_s": lines / dt, "lines": lines, "det_only_pages_per_s": n / dt_det,
                  "page_bytes": int(pages[0].nbytes), "stage_ms_last_batch": {k: round(v, 2) for k, v in ocr.engine.ocr_stage_times().items()}}))

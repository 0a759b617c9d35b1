"""Concurrent single-page callers with and without the micro-batcher: python tools/bench_batcher.py [threads] [pages]"""
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onnxocr_b200 import pipeline, synth_pages  # noqa: E402
from onnxocr_b200.batcher import MicroBatcher  # noqa: E402

threads = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
ocr = pipeline.ONNXPaddleOcr(use_angle_cls=True, precision="bf16")
pages = synth_pages.e2e_pages_c4(n, seed=2)
for p in pages[:8]:
    ocr.ocr(p)
ocr.ocr_batch(pages[:16])


def run(fn):
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        res = list(ex.map(fn, pages))
    return n / (time.perf_counter() - t0), res


direct, r0 = run(ocr.ocr)
mb = MicroBatcher(ocr, max_batch=16, max_wait_ms=2.0)
run(mb.ocr)
batched, r1 = run(mb.ocr)
mb.close()
print(json.dumps({"config": "C4 pages, %d caller threads, one page per ocr() call" % threads, "pages": n,
                  "direct_pages_per_s": direct, "micro_batched_pages_per_s": batched, "library_calls": mb.batches,
                  "identical": [[x[1][0] for x in a[0]] for a in r0] == [[x[1][0] for x in a[0]] for a in r1]}))

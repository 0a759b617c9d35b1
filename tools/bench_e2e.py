"""C4 end-to-end through the fused Level-C entry: python tools/bench_e2e.py [bf16|fp32] [pages] [bucket] [pinned]
   pinned = 1: the pages live in page-locked buffers (what batcher.decode_pages(engine=...) produces), so their H2D copies
   skip the host-side staging of pageable numpy arrays."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onnxocr_b200 import pipeline, synth_pages  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 0
ocr = pipeline.ONNXPaddleOcr(use_angle_cls=True, precision=prec, rec_width_bucket=bucket)
pages = synth_pages.e2e_pages_c4(n, seed=2)
pinned = len(sys.argv) > 4 and sys.argv[4] == "1"
if pinned:
    pp = []
    for p in pages:
        a = ocr.engine.pinned_array(p.shape, p.dtype)
        a[...] = p
        pp.append(a)
    pages = pp
for _ in range(2):
    res = ocr.ocr_batch(pages)
t0 = time.perf_counter()
reps = 3
for _ in range(reps):
    res = ocr.ocr_batch(pages)
dt = (time.perf_counter() - t0) / reps
stages = ocr.engine.ocr_stage_times()
lines = sum(len(r) for r in res)
t1 = time.perf_counter()
single, per_page_ms = [], []
for p in pages:
    tp0 = time.perf_counter()
    single.append(ocr.ocr(p)[0])
    per_page_ms.append(round((time.perf_counter() - tp0) * 1e3, 2))
dt1 = time.perf_counter() - t1
same = sum(1 for a, b in zip(res, single) if [x[1][0] for x in a] == [x[1][0] for x in b])
print(json.dumps({"config": "C4 end-to-end det+cls+rec, %d mixed-size pages" % n, "precision": prec, "rec_width_bucket": bucket, "pinned_pages": pinned,
                  "batched_pages_per_s": n / dt, "batched_lines_per_s": lines / dt, "lines": lines,
                  "one_page_per_call_pages_per_s": n / dt1, "pages_identical_to_single_calls": "%d/%d" % (same, n),
                  "one_page_ms": per_page_ms, "launches": ocr.engine.launch_count(), "stage_ms_last_batch": {k: round(v, 2) for k, v in stages.items()}}))

"""Determinism / batch-independence / device-independence of the fused path:

    python tools/determinism_check.py [bf16|fp32] [gpu_id] [dump.json]

Runs 128 C4 pages twice in one call and once as two interleaved shards; with a dump path the results are written so that
runs on different GPUs can be compared (`cmp`)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onnxocr_b200 import pipeline, synth_pages  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
gpu = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ocr = pipeline.ONNXPaddleOcr(use_angle_cls=True, precision=prec, gpu_id=gpu)
pages = synth_pages.e2e_pages_c4(128, seed=2)


def txt(res):
    return [[[x[0], x[1][0], round(x[1][1], 5)] for x in r] for r in res]


a = txt(ocr.ocr_batch(pages))
b = txt(ocr.ocr_batch(pages))
print("gpu %d repeat identical: %d / 128" % (gpu, sum(x == y for x, y in zip(a, b))))
c = [None] * 128
c[0::2] = txt(ocr.ocr_batch(pages[0::2]))
c[1::2] = txt(ocr.ocr_batch(pages[1::2]))
bad = [i for i in range(128) if a[i] != c[i]]
print("gpu %d shard identical: %d / 128 %s" % (gpu, 128 - len(bad), bad[:5]))
for i in bad[:2]:
    for k, (x, y) in enumerate(zip(a[i], c[i])):
        if x != y:
            print(i, k, x[1:], y[1:], x[0] == y[0])
if len(sys.argv) > 3:
    json.dump(a, open(sys.argv[3], "w"))

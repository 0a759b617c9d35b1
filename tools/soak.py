"""Stability soak: many calls with changing shapes / batch sizes through every entry point (run under `timeout`)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onnxocr_b200 import _lib, pipeline, synth_pages  # noqa: E402

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 60
ocr = pipeline.ONNXPaddleOcr(use_angle_cls=True, precision="bf16")
rng = np.random.default_rng(0)
pool = synth_pages.e2e_pages_c4(24, seed=5) + synth_pages.det_pages_c2(8, 960, 30, seed=9)
t0 = time.time()
calls = pages = lines = 0
while time.time() - t0 < secs:
    k = int(rng.integers(1, 20))
    idx = rng.integers(0, len(pool), k)
    batch = [pool[i] for i in idx]
    mode = int(rng.integers(0, 5))
    if mode == 0:
        res = ocr.ocr_batch(batch)
        lines += sum(len(r) for r in res)
    elif mode == 1:
        if rng.random() < 0.5:
            res = ocr.text_detector.batch(batch)
        else:                                    # det_db_score_mode='slow' (runs its scores from db_post_collect)
            res = ocr.engine.det_pages(batch, _lib.det_params(score_mode="slow", box_thresh=float(rng.uniform(0.3, 0.7))))
        lines += sum(len(r) for r in res)
    elif mode == 2:
        res = ocr.ocr(batch[0])
        lines += len(res[0])
    elif mode == 4:
        # streaming det with random batch sizes / shapes, checked against the blocking call on a sample
        batches = [[pool[i] for i in rng.integers(0, len(pool), int(rng.integers(1, 18)))] for _ in range(int(rng.integers(1, 7)))]
        sp = ocr.text_detector.params if rng.random() < 0.7 else _lib.det_params(score_mode="slow")
        outs = list(ocr.engine.det_pages_stream(batches, sp))
        lines += sum(len(b) for o in outs for b in o)
        j = int(rng.integers(0, len(batches)))
        ref = ocr.engine.det_pages(batches[j], sp)
        assert all(np.array_equal(a, b) for a, b in zip(outs[j], ref)), "streaming det differs from the blocking call"
        pages += sum(len(b) for b in batches)
    else:
        crops = synth_pages.rec_crops_c3(int(rng.integers(1, 40)), seed=int(rng.integers(0, 1000)), w=int(rng.integers(60, 900)))
        lines += len(ocr.text_recognizer(crops))
    calls += 1
    pages += k
import subprocess  # noqa: E402
try:
    mem = subprocess.run(["nvidia-smi", "--query-gpu=memory.used", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
except Exception:
    mem = "?"
print("device memory in use at the end: " + mem)
print("soak ok: %d calls, %d pages, %d lines in %.1f s, %d launches" % (calls, pages, lines, time.time() - t0, ocr.engine.launch_count()))

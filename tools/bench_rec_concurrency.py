"""How well do independent small rec batches overlap on the auxiliary streams?  K batches of 6 crops, every batch with
its own width (so nothing is merged), through Engine.rec_crops; prints ms per call and per batch.
    python tools/bench_rec_concurrency.py            (env: B200OCR_AUX, B200OCR_TC=stream, B200OCR_NO_TC=1, ...)"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onnxocr_b200 import _lib, synth_graphs, synth_pages  # noqa: E402

paths = synth_graphs.ensure_models()
eng = _lib.Engine(0)
eng.load_onnx(_lib.REC, paths["rec"], _lib.BF16)
base = synth_pages.rec_crops_c3(6, seed=1)
rng = np.random.default_rng(0)
for K in (1, 2, 4, 8, 16, 32, 64):
    crops = []
    for k in range(K):
        w = 320 + 32 * k
        for c in base:
            img = np.full((48, w, 3), 255, np.uint8)
            img[:, :320] = c
            crops.append(img)
    for _ in range(3):
        eng.rec_crops(crops)
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        eng.rec_crops(crops)
    ms = (time.perf_counter() - t0) / reps * 1e3
    print("K=%3d batches of 6: %7.3f ms per call, %6.3f ms per batch, %7.0f lines/s" % (K, ms, ms / K, 6 * K / ms * 1e3), flush=True)

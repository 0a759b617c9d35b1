import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from onnxocr_b200 import pipeline, synth_pages
ocr = pipeline.ONNXPaddleOcr(use_angle_cls=True, precision="bf16")
rec = ocr.text_recognizer
crops0 = synth_pages.rec_crops_c3(6, seed=1, w=400)
rec(crops0); rec(crops0); rec(crops0)
ts = []
for w in range(401, 461):
    crops = synth_pages.rec_crops_c3(6, seed=1, w=w)
    t = []
    for rep in range(5):
        t0 = time.perf_counter(); rec(crops); t.append((time.perf_counter() - t0) * 1e3)
    ts.append(t)
ts = np.array(ts)
print("ms per call (6 crops), median over 60 new widths: 1st %.2f  2nd %.2f  3rd %.2f  4th %.2f  5th %.2f" % tuple(np.median(ts, axis=0)))

import sys, time, json
sys.path.insert(0, "/root/repo")
import numpy as np
from onnxocr_b200 import pipeline, synth_pages
ocr = pipeline.ONNXPaddleOcr(use_angle_cls=True, precision="bf16")
pages = synth_pages.e2e_pages_c4(32, seed=2)
for rep in range(3):
    t0 = time.perf_counter()
    st = []
    for p in pages:
        ocr.ocr(p)
        st.append(ocr.engine.ocr_stage_times())
    dt = time.perf_counter() - t0
    agg = {k: round(float(np.mean([s[k] for s in st])), 2) for k in st[0]}
    print("pass %d: %.1f pages/s  (%.2f ms/page)  mean stage ms %s" % (rep, len(pages) / dt, 1e3 * dt / len(pages), agg))

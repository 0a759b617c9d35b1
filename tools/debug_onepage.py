import os, sys, time
sys.path.insert(0, os.getcwd())
import warnings; warnings.filterwarnings("ignore")
import numpy as np
from onnxocr_b200 import pipeline, synth_pages
ocr = pipeline.ONNXPaddleOcr(use_angle_cls=True, precision="bf16", rec_width_bucket=0)
pages = synth_pages.e2e_pages_c4(16, seed=2)
for _ in range(2): ocr.ocr_batch(pages)
for i, pg in enumerate(pages):
    t0 = time.perf_counter(); r = ocr.ocr(pg); t1 = time.perf_counter()
    st = ocr.engine.ocr_stage_times()
    print("page %2d %s: %.2f ms, %d lines" % (i, pg.shape, (t1 - t0) * 1e3, len(r[0])), {k: round(v, 2) for k, v in st.items()})

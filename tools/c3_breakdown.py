"""C3 (rec-only, 256 crops of 48x320) through the public API: where the time between the device's 3.9 ms and the
caller's wall clock goes.  `python tools/c3_breakdown.py`"""
import sys
import time

import numpy as np

sys.path.insert(0, '/root/repo')
from onnxocr_b200 import _lib, pipeline, synth_pages  # noqa: E402

ocr = pipeline.ONNXPaddleOcr(use_angle_cls=True, precision="bf16")
rec = ocr.text_recognizer
eng = rec.engine
crops = synth_pages.rec_crops_c3(256, seed=1)
for _ in range(5):
    rec(crops)


def timeit(fn, n=30):
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n * 1e3


print("TextRecognizer.__call__      %.3f ms" % timeit(lambda: rec(crops)))
print("Engine.rec_crops             %.3f ms" % timeit(lambda: eng.rec_crops(crops, 6)))
print("  _crop_args only            %.3f ms" % timeit(lambda: eng._crop_args(crops)))
ids, scores = eng.rec_crops(crops, 6)
print("  charset mapping only       %.3f ms" % timeit(lambda: [(rec.postprocess_op.text(i), float(s)) for i, s in zip(ids, scores)]))
print("stage times of the last call:", eng.stage_times())
print("time_plan(REC, 256, 48, 320, head) %.3f ms" % eng.time_plan(_lib.REC, 256, 48, 320, head=True))
one = np.ascontiguousarray(np.concatenate([c.reshape(-1) for c in crops]))
views = [one[i * 48 * 320 * 3:(i + 1) * 48 * 320 * 3].reshape(48, 320, 3) for i in range(256)]
print("Engine.rec_crops, crops as views of one buffer  %.3f ms" % timeit(lambda: eng.rec_crops(views, 6)))

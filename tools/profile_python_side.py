import sys, cProfile, pstats, io
sys.path.insert(0, "/root/repo")
from onnxocr_b200 import pipeline, synth_pages
ocr = pipeline.ONNXPaddleOcr(use_angle_cls=True, precision="bf16")
crops = synth_pages.rec_crops_c3(256, seed=1)
for _ in range(3): ocr.text_recognizer(crops)
pr = cProfile.Profile(); pr.enable()
for _ in range(20): ocr.text_recognizer(crops)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(14); print(s.getvalue()[:2500])

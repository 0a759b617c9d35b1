import sys
sys.path.insert(0, '/root/repo')
from onnxocr_b200 import _lib, synth_graphs, synth_pages
paths = synth_graphs.ensure_models()
eng = _lib.Engine(0)
eng.load_onnx(_lib.REC, paths["rec"], _lib.BF16)
crops = synth_pages.rec_crops_c3(256, seed=1)
for _ in range(3):
    ids, scores = eng.rec_crops(crops)
print(len(ids))

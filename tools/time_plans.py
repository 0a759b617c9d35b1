"""Graph-replay time of whole plans at the shapes the stages use: python tools/time_plans.py [bf16|fp32]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onnxocr_b200 import _lib, synth_graphs  # noqa: E402

prec = _lib.BF16 if (len(sys.argv) < 2 or sys.argv[1] == "bf16") else _lib.FP32
paths = synth_graphs.ensure_models()
paths["cls"] = os.path.join(synth_graphs.ASSET_DIR, "ppocrv5_cls.onnx")
eng = _lib.Engine(0)
for k, nm in ((_lib.DET, "det"), (_lib.CLS, "cls"), (_lib.REC, "rec")):
    eng.load_onnx(k, paths[nm], prec)
cases = [("det", _lib.DET, 16, 960, 960, False), ("det", _lib.DET, 1, 960, 960, False), ("det", _lib.DET, 1, 736, 1280, False),
         ("det", _lib.DET, 4, 2560, 1824, False),
         ("cls", _lib.CLS, 6, 48, 192, False), ("cls", _lib.CLS, 64, 48, 192, False), ("cls", _lib.CLS, 512, 48, 192, False),
         ("rec", _lib.REC, 6, 48, 320, True), ("rec", _lib.REC, 6, 48, 640, True), ("rec", _lib.REC, 6, 48, 1280, True),
         ("rec", _lib.REC, 32, 48, 640, True), ("rec", _lib.REC, 256, 48, 320, True), ("rec", _lib.REC, 256, 48, 320, False)]
for nm, k, n, h, w, head in cases:
    ms = eng.time_plan(k, n, h, w, head, iters=10)
    st = eng.plan_stats(k, n, h, w)
    fl, launches = st["flops"], st["launches"]
    print("%-4s n=%-4d %4dx%-5d head=%d  %8.3f ms  %6d launches  %.1f TF/s  %.3f ms/item" % (nm, n, h, w, head, ms, launches, fl / ms / 1e9, ms / n))

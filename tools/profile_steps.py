"""Per-launch profile of one plan (CUDA events between launches): python tools/profile_steps.py det|rec|cls [bf16|fp32] n h w"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onnxocr_b200 import _lib, synth_graphs  # noqa: E402

kind = {"det": _lib.DET, "cls": _lib.CLS, "rec": _lib.REC}[sys.argv[1]]
prec = _lib.BF16 if (len(sys.argv) < 3 or sys.argv[2] == "bf16") else _lib.FP32
n, h, w = [int(v) for v in sys.argv[3:6]] if len(sys.argv) >= 6 else (16, 960, 960)
paths = synth_graphs.ensure_models()
paths["cls"] = os.path.join(synth_graphs.ASSET_DIR, "ppocrv5_cls.onnx")
eng = _lib.Engine(0)
eng.load_onnx(kind, paths[sys.argv[1]], prec)
eng.profile_plan(kind, n, h, w, iters=2)
prof = eng.profile_plan(kind, n, h, w, iters=5)
tot = sum(s["ms"] for s in prof)
print("total %.3f ms over %d launches" % (tot, len(prof)))
rows = prof if os.environ.get("ORDER") == "1" else sorted(prof, key=lambda s: -s["ms"])[:int(os.environ.get("TOP", "40"))]
for s in rows:
    print("%8.4f ms %5.1f%%  %8.2f TF/s %8.1f GB/s  %s" % (s["ms"], 100 * s["ms"] / tot, s["flops"] / max(s["ms"], 1e-9) / 1e9,
                                                       s["bytes"] / max(s["ms"], 1e-9) / 1e6, s["name"]))

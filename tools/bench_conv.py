"""Single-layer conv micro-benchmark through the engine: python tools/bench_conv.py cin cout k n h w [iters]"""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onnxocr_b200 import _lib  # noqa: E402
from onnxocr_b200 import onnx_writer as ow  # noqa: E402

cin, cout, k, n, h, w = [int(v) for v in sys.argv[1:7]]
iters = int(sys.argv[7]) if len(sys.argv) > 7 else 5
rng = np.random.default_rng(0)
wt = (rng.standard_normal((cout, cin, k, k)) / np.sqrt(cin * k * k)).astype(np.float32)
b = (rng.standard_normal(cout) * 0.1).astype(np.float32)
nodes = [ow.enc_node("Constant", [], ["w"], value=wt), ow.enc_node("Constant", [], ["b"], value=b),
         ow.enc_node("Conv", ["x", "w", "b"], ["c"], dilations=[1, 1], kernel_shape=[k, k], strides=[1, 1], group=1, pads=[k // 2] * 4),
         ow.enc_node("Relu", ["c"], ["y"])]
blob = ow.enc_model(nodes, [ow.enc_value_info("x", ["n", cin, "h", "w"])], [ow.enc_value_info("y", ["n", cout, "h", "w"])])
path = os.path.join(tempfile.mkdtemp(), "conv.onnx")
open(path, "wb").write(blob)
eng = _lib.Engine(0)
eng.load_onnx(_lib.DET, path, _lib.BF16)
eng.profile_plan(_lib.DET, n, h, w, iters=2)
for s in eng.profile_plan(_lib.DET, n, h, w, iters=iters):
    print("%8.4f ms  %8.2f TF/s %8.1f GB/s  %s" % (s["ms"], s["flops"] / max(s["ms"], 1e-9) / 1e9, s["bytes"] / max(s["ms"], 1e-9) / 1e6, s["name"]))

"""How far is the bf16 recogniser from the fp32 one, and how much margin does the synthetic model leave?
Prints line agreement on the 256 C3 crops and, for 48 of them, per-time-step statistics of the class probabilities
(top-1 / top-2 margin of the fp32 run where the bf16 argmax differs, max |dp|).  `python tools/bf16_margin.py`"""
import sys

import numpy as np

sys.path.insert(0, '/root/repo')
from onnxocr_b200 import _lib, synth_graphs, synth_pages  # noqa: E402

paths = synth_graphs.ensure_models()
crops = synth_pages.rec_crops_c3(256, seed=1)
res = {}
for name, prec in (("fp32", _lib.FP32), ("bf16", _lib.BF16)):
    eng = _lib.Engine(0)
    eng.load_onnx(_lib.REC, paths["rec"], prec)
    ids, scores = eng.rec_crops(crops)
    x = np.stack([((c.astype(np.float32) / 255.0 - 0.5) / 0.5).transpose(2, 0, 1) for c in crops[:48]])
    probs = eng.session_run(_lib.REC, x)
    res[name] = (ids, scores, probs)
    eng.close()
same = sum(1 for a, b in zip(res["fp32"][0], res["bf16"][0]) if np.array_equal(a, b))
print("identical lines: %d / %d" % (same, len(crops)))
pf, pb = res["fp32"][2], res["bf16"][2]
print("probs shape", pf.shape, "max |dp| = %.4f" % np.abs(pf - pb).max())
af, ab = pf.argmax(2), pb.argmax(2)
srt = np.sort(pf, axis=2)
margin = srt[:, :, -1] - srt[:, :, -2]
diff = af != ab
print("time steps with another argmax: %d of %d" % (diff.sum(), diff.size))
print("fp32 margins at those steps:", np.round(np.sort(margin[diff]), 4).tolist()[:40])
print("fp32 margin quantiles over all steps (1%%, 5%%, 25%%, 50%%): %s" % np.round(np.quantile(margin, [0.01, 0.05, 0.25, 0.5]), 4).tolist())
print("|dp| of the top-1 class: mean %.5f, 99%% %.5f, max %.5f" % (
    np.abs(np.take_along_axis(pf - pb, af[:, :, None], 2)).mean(),
    np.quantile(np.abs(np.take_along_axis(pf - pb, af[:, :, None], 2)), 0.99),
    np.abs(np.take_along_axis(pf - pb, af[:, :, None], 2)).max()))
for i, (a, b) in enumerate(zip(res["fp32"][0], res["bf16"][0])):
    if not np.array_equal(a, b) and i < 48:
        t = np.nonzero(diff[i])[0]
        print("line %d: steps %s, margins %s" % (i, t.tolist(), np.round(margin[i, t], 4).tolist()))

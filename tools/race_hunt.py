"""Hunts for batch-composition / timing dependence of the fused path: the same pages in differently sized calls must
give bit-identical boxes, ids, scores.  python tools/race_hunt.py [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from onnxocr_b200 import pipeline, synth_pages  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 6
ocr = pipeline.ONNXPaddleOcr(use_angle_cls=True, precision="bf16")
pages = synth_pages.e2e_pages_c4(128, seed=2)
P = ocr._ocr_params(True)


def raw(pgs):
    return [(r["boxes"].tolist(), [i.tolist() for i in r["ids"]], r["scores"].tolist(), r["cls_labels"].tolist())
            for r in ocr.engine.ocr_pages(pgs, P)]


ref = raw(pages)
nb = 0
for it in range(iters):
    full = raw(pages)
    k = 2 + it % 5
    sh = [None] * 128
    for j in range(k):
        sh[j::k] = raw(pages[j::k])
    for name, x in (("full", full), ("shard%d" % k, sh)):
        bad = [i for i in range(128) if x[i] != ref[i]]
        if bad:
            nb += 1
            i = bad[0]
            what = [n for n, a, b in zip(("boxes", "ids", "scores", "cls"), x[i], ref[i]) if a != b]
            print(it, name, "bad pages", bad[:6], "first differs in", what)
            lines = [l for l in range(len(ref[i][2])) if x[i][2][l] != ref[i][2][l] or x[i][1][l] != ref[i][1][l]]
            for l in lines[:6]:
                b = np.array(ref[i][0][l])
                w = max(np.linalg.norm(b[0] - b[1]), np.linalg.norm(b[2] - b[3]))
                h = max(np.linalg.norm(b[0] - b[3]), np.linalg.norm(b[1] - b[2]))
                print("   line %d of %d  crop %dx%d  got score %.6f len %d | ref score %.6f len %d" % (
                    l, len(ref[i][2]), int(w), int(h), x[i][2][l], len(x[i][1][l]), ref[i][2][l], len(ref[i][1][l])))
print("differences:", nb)

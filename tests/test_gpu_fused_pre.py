"""GPU: the det plans read the u8 pages directly (DetResizeForTest + NormalizeImage + ToCHWImage fused into the stem
convolution).  Three variants are compared on identity-size pages, resized pages (up- and down-scaling, the exact-2x
INTER_AREA case) and the padded tiny page:

  two-kernel path   B200OCR_FUSED_PRE=0: resize_norm_kernel + conv_stem_kernel
  fused FFMA        pre_stem_kernel (fp32 plans; bf16 plans with B200OCR_STEM_MMA=0): must be BIT-IDENTICAL to the two-kernel path
  fused tensor core pre_stem_mma_kernel (bf16 plans, default): same bf16 inputs, weights as bf16 hi + lo, fp32 accumulation in
                    another order -> the two bf16 variants are each within the north star's bf16 tolerance (2e-2) of the fp32
                    maps, hence within 4e-2 of each other; same boxes up to the rare pixel that flips across the threshold."""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import sys
sys.path.insert(0, %(root)r)
import numpy as np
from onnxocr_b200 import _lib, synth_graphs, synth_pages
paths = synth_graphs.ensure_models()
out = {}
shapes = [(480, 640), (512, 704), (1200, 900), (1920, 1280), (300, 333), (20, 30), (960, 960)]
pages = [synth_pages.render_page(h, w, 6, seed=300 + i)[0] for i, (h, w) in enumerate(shapes)]
pages += [synth_pages.render_page(480, 640, 6, seed=400 + i)[0] for i in range(9)]      # a 10-page group: one lane, n = 10
for prec, name in ((_lib.FP32, "fp32"), (_lib.BF16, "bf16")):
    eng = _lib.Engine(0)
    eng.load_onnx(_lib.DET, paths["det"], prec)
    for rep in range(2):                      # the second pass replays the CUDA graphs
        boxes, maps = eng.det_pages(pages, want_prob=True)
    for i, (b, m) in enumerate(zip(boxes, maps)):
        out["%%s_box_%%d" %% (name, i)] = b
        out["%%s_map_%%d" %% (name, i)] = m
    out[name + "_launches"] = np.int64(eng.launch_count())
    eng.close()
out["n_pages"] = np.int64(len(pages))
np.savez(%(dst)r, **out)
"""


def _run(env_extra, dst):
    env = dict(os.environ)
    env.update(env_extra)
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT, "dst": dst}], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    return np.load(dst)


def test_fused_pre_stem_variants(lib_built):
    with tempfile.TemporaryDirectory() as d:
        fused = _run({}, os.path.join(d, "a.npz"))
        plain = _run({"B200OCR_FUSED_PRE": "0"}, os.path.join(d, "b.npz"))
        ffma = _run({"B200OCR_STEM_MMA": "0"}, os.path.join(d, "c.npz"))
        n = int(plain["n_pages"])
        assert sum(len(plain["fp32_box_%d" % i]) for i in range(n)) > 40
        for prec, other in (("fp32", fused), ("fp32", ffma), ("bf16", ffma)):
            for i in range(n):
                assert np.array_equal(other["%s_map_%d" % (prec, i)], plain["%s_map_%d" % (prec, i)]), (prec, i, "maps differ: the FFMA fused path must be bit-identical")
                assert np.array_equal(other["%s_box_%d" % (prec, i)], plain["%s_box_%d" % (prec, i)])
            assert int(other[prec + "_launches"]) < int(plain[prec + "_launches"]), "the fused path did not remove the resize/normalise launches"
        same_boxes = total_boxes = 0
        worst = 0.0
        for i in range(n):
            a, b = fused["bf16_map_%d" % i], plain["bf16_map_%d" % i]
            worst = max(worst, float(np.abs(a - b).max()))
            ba, bb = fused["bf16_box_%d" % i], plain["bf16_box_%d" % i]
            assert abs(len(ba) - len(bb)) <= 1
            total_boxes += len(bb)
            same_boxes += sum(1 for q in bb if any(np.array_equal(q, r) for r in ba))
        assert worst <= 4e-2, "tensor-core stem: probability maps differ by %g from the FFMA path" % worst
        # against the fp32 maps: the same bar tests/test_gpu_net.py applies to the bf16 mode (99.9th percentile within 2e-2)
        diff = np.concatenate([np.abs(fused["bf16_map_%d" % i] - plain["fp32_map_%d" % i]).ravel() for i in range(n)])
        assert float(np.percentile(diff, 99.9)) <= 2e-2, "tensor-core stem: bf16 maps too far from the fp32 maps"
        assert same_boxes >= 0.9 * total_boxes, (same_boxes, total_boxes)
        print("tensor-core stem vs FFMA stem (bf16): max |dprob| %.3g, %d / %d boxes identical" % (worst, same_boxes, total_boxes))

"""GPU: the det plans read the u8 pages directly (DetResizeForTest + NormalizeImage + ToCHWImage fused into the stem
convolution, pre_stem_kernel).  With B200OCR_FUSED_PRE=0 the engine runs resize_norm_kernel + conv_stem_kernel instead;
both paths must give bit-identical probability maps and boxes, in fp32 and in bf16, for identity-size pages, resized pages
(up- and down-scaling, the exact-2x INTER_AREA case) and the padded tiny page."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import hashlib, json, sys
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
    hb = hashlib.sha1(b"".join(np.ascontiguousarray(b).tobytes() for b in boxes)).hexdigest()
    hm = hashlib.sha1(b"".join(np.ascontiguousarray(m).tobytes() for m in maps)).hexdigest()
    out[name] = {"boxes": hb, "maps": hm, "n": [int(len(b)) for b in boxes], "launches": eng.launch_count()}
    eng.close()
print(json.dumps(out))
"""


def _run(env_extra):
    env = dict(os.environ)
    env.update(env_extra)
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_fused_pre_stem_is_bit_identical(lib_built):
    fused = _run({})
    plain = _run({"B200OCR_FUSED_PRE": "0"})
    for prec in ("fp32", "bf16"):
        assert sum(fused[prec]["n"]) > 40
        assert fused[prec]["n"] == plain[prec]["n"]
        assert fused[prec]["maps"] == plain[prec]["maps"], prec + ": probability maps differ between the fused and the two-kernel path"
        assert fused[prec]["boxes"] == plain[prec]["boxes"]
        assert fused[prec]["launches"] < plain[prec]["launches"], "the fused path did not remove the resize/normalise launches"

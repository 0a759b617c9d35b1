"""GPU: the register-sliding depthwise kernel (bf16 activations, fp32 weights/accumulation) against torch fp32 on
single-layer graphs, for every kernel size / stride / channel-width variant the det, cls and rec graphs use.

Tolerance: inputs are rounded to bf16 before the torch reference is computed, so the remaining difference is the
fp32 accumulation order plus the bf16 rounding of the output (<= 2^-8 relative)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from onnxocr_b200 import _lib
from onnxocr_b200 import onnx_writer as ow

pytestmark = pytest.mark.gpu


def bf16_round(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(torch.bfloat16).to(torch.float32).numpy()


def dw_model(path, c, k, sh, sw):
    rng = np.random.default_rng(c * 100 + k * 10 + sh * 2 + sw)
    w = (rng.standard_normal((c, 1, k, k)) / k).astype(np.float32)
    b = (rng.standard_normal(c) * 0.1).astype(np.float32)
    nodes = [ow.enc_node("Constant", [], ["w"], value=w), ow.enc_node("Constant", [], ["b"], value=b),
             ow.enc_node("Conv", ["x", "w", "b"], ["c"], dilations=[1, 1], kernel_shape=[k, k], strides=[sh, sw], group=c,
                         pads=[k // 2] * 4),
             ow.enc_node("Relu", ["c"], ["y"])]
    blob = ow.enc_model(nodes, [ow.enc_value_info("x", ["n", c, "h", "w"])], [ow.enc_value_info("y", ["n", c, "h", "w"])])
    with open(path, "wb") as fh:
        fh.write(blob)
    return w, b


CASES = [  # c, k, sh, sw, (h, w), n
    (16, 3, 1, 1, (70, 90), 2), (32, 3, 2, 2, (64, 96), 1), (48, 3, 1, 1, (33, 47), 1), (96, 3, 2, 2, (41, 59), 2),
    (192, 5, 1, 1, (30, 30), 1), (384, 5, 2, 2, (30, 31), 1), (240, 5, 1, 2, (12, 80), 3), (480, 5, 2, 1, (6, 80), 2),
    (64, 3, 2, 1, (24, 160), 1), (120, 3, 1, 1, (1, 40), 2), (24, 5, 1, 1, (48, 192), 1), (8, 3, 1, 2, (17, 19), 1),
    (200, 5, 2, 2, (7, 9), 1), (40, 5, 2, 1, (19, 66), 1),
]


@pytest.mark.parametrize("c,k,sh,sw,hw,n", CASES)
def test_dwconv_vs_torch(lib_built, tmp_path, c, k, sh, sw, hw, n):
    path = str(tmp_path / "dw.onnx")
    w, b = dw_model(path, c, k, sh, sw)
    eng = _lib.Engine(0)
    eng.load_onnx(_lib.DET, path, _lib.BF16)
    rng = np.random.default_rng(2)
    x = bf16_round(rng.standard_normal((n, c, hw[0], hw[1])).astype(np.float32))
    out = eng.session_run(_lib.DET, x)
    ref = F.relu(F.conv2d(torch.from_numpy(x), torch.from_numpy(w), torch.from_numpy(b), stride=(sh, sw), padding=k // 2,
                          groups=c)).numpy()
    assert out.shape == ref.shape
    err = np.abs(out - ref)
    tol = 2.0 ** -7 * np.maximum(np.abs(ref), 1.0)
    bad = err > tol
    assert not bad.any(), "max err %g at %s, %d bad of %d" % (err.max(), np.unravel_index(err.argmax(), err.shape), bad.sum(), bad.size)
    eng.close()


_AB_SCRIPT = r"""
import sys, hashlib
import numpy as np
sys.path.insert(0, %(root)r)
sys.path.insert(0, %(tests)r)
import test_gpu_dw as t
from onnxocr_b200 import _lib
h = hashlib.sha256()
for c, k, sh, sw, hw, n in t.CASES:
    path = %(tmp)r + "/dw_%%d_%%d_%%d_%%d.onnx" %% (c, k, sh, sw)
    t.dw_model(path, c, k, sh, sw)
    eng = _lib.Engine(0)
    eng.load_onnx(_lib.DET, path, _lib.BF16)
    x = t.bf16_round(np.random.default_rng(3).standard_normal((n, c, hw[0], hw[1])).astype(np.float32))
    h.update(eng.session_run(_lib.DET, x).tobytes())
    eng.close()
print("DIGEST", h.hexdigest())
"""


def test_dwconv_tma_tile_equals_cpasync_tile(lib_built, tmp_path):
    """The halo tile loaded as one TMA box (default) and by the cp.async loop (B200OCR_DW_TMA=0, read once per process)
    must give bit-identical outputs on every variant, incl. the stride-2 ones whose tile layout lost its padding."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = _AB_SCRIPT % {"root": root, "tests": os.path.join(root, "tests"), "tmp": str(tmp_path)}
    digests = []
    for flag in ("1", "0"):
        env = dict(os.environ, B200OCR_DW_TMA=flag)
        r = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        digests.append([l for l in r.stdout.splitlines() if l.startswith("DIGEST")][0])
    assert digests[0] == digests[1]

"""GPU: the tcgen05/TMA implicit-GEMM convolution (bf16 in, fp32 accumulate) against torch fp32 on single-layer graphs.

Tolerance: bf16 has 8 bits of mantissa; inputs and weights are rounded to bf16 before the torch reference is computed,
so the remaining difference is accumulation order + the bf16 rounding of the output (<= 2^-8 relative)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from onnxocr_b200 import _lib
from onnxocr_b200 import onnx_writer as ow

pytestmark = pytest.mark.gpu


def bf16_round(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(torch.bfloat16).to(torch.float32).numpy()


def single_conv_model(path, cin, cout, k, act):
    rng = np.random.default_rng(cin * 1000 + cout * 10 + k)
    w = (rng.standard_normal((cout, cin, k, k)) / np.sqrt(cin * k * k)).astype(np.float32)
    b = (rng.standard_normal(cout) * 0.1).astype(np.float32)
    nodes = [ow.enc_node("Constant", [], ["w"], value=w), ow.enc_node("Constant", [], ["b"], value=b),
             ow.enc_node("Conv", ["x", "w", "b"], ["c"], dilations=[1, 1], kernel_shape=[k, k], strides=[1, 1], group=1,
                         pads=[k // 2] * 4)]
    out = "c"
    if act == "relu":
        nodes.append(ow.enc_node("Relu", ["c"], ["y"]))
        out = "y"
    blob = ow.enc_model(nodes, [ow.enc_value_info("x", ["n", cin, "h", "w"])], [ow.enc_value_info(out, ["n", cout, "h", "w"])])
    with open(path, "wb") as fh:
        fh.write(blob)
    return w, b


@pytest.mark.parametrize("cin,cout,k,hw,n", [(96, 24, 1, (40, 56), 2), (96, 24, 3, (40, 56), 2), (16, 32, 1, (64, 64), 1),
                                               (384, 360, 1, (30, 30), 2), (24, 96, 3, (17, 23), 1), (96, 96, 3, (240, 240), 1),
                                               (128, 128, 1, (12, 160), 3), (64, 200, 3, (9, 9), 1)])
def test_conv_tc_vs_torch(lib_built, tmp_path, cin, cout, k, hw, n):
    path = str(tmp_path / "conv.onnx")
    act = "relu" if (cin + cout) % 3 == 0 else None
    w, b = single_conv_model(path, cin, cout, k, act)
    eng = _lib.Engine(0)
    eng.load_onnx(_lib.DET, path, _lib.BF16)
    rng = np.random.default_rng(1)
    x = bf16_round(rng.standard_normal((n, cin, hw[0], hw[1])).astype(np.float32))
    out = eng.session_run(_lib.DET, x)
    ref = F.conv2d(torch.from_numpy(x), torch.from_numpy(bf16_round(w)), torch.from_numpy(b), padding=k // 2)
    if act == "relu":
        ref = F.relu(ref)
    ref = ref.numpy()
    assert out.shape == ref.shape
    err = np.abs(out - ref)
    tol = 2.0 ** -7 * np.maximum(np.abs(ref), 1.0)
    bad = err > tol
    assert not bad.any(), "max err %g at %s (ref %g got %g), %d bad of %d" % (
        err.max(), np.unravel_index(err.argmax(), err.shape), ref.flat[err.argmax()], out.flat[err.argmax()], bad.sum(), bad.size)
    eng.close()

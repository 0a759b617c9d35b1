"""A module named `onnxruntime` that routes the reference's engine seam to libb200ocr.so.

Put `onnxocr_b200/ort_shim` first on `sys.path` (or call `onnxocr_b200.install_ort_shim()`) and the reference's
`onnxocr/predict_base.py` — `import onnxruntime` / `onnxruntime.InferenceSession(model_dir, None, providers=...)` —
runs unmodified on the B200 engine.  See INTEGRATION.md.
"""
from onnxocr_b200.session import B200Session as InferenceSession  # noqa: F401

__version__ = "1.14.1+b200ocr"


def get_device():
    return "GPU"


def get_available_providers():
    return ["B200ExecutionProvider"]


class SessionOptions(object):
    pass

"""b200ocr — B200-native PP-OCRv5 det/cls/rec engine behind OnnxOCR's onnxruntime seam."""
import os
import sys

__all__ = ["install_ort_shim"]


def install_ort_shim():
    """Makes `import onnxruntime` resolve to the libb200ocr-backed shim (onnxocr/predict_base.py:1 stays untouched)."""
    shim = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ort_shim")
    if shim not in sys.path:
        sys.path.insert(0, shim)
    sys.modules.pop("onnxruntime", None)
    import onnxruntime  # noqa: F401
    return onnxruntime

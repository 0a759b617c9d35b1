"""Imports the UNMODIFIED reference package from /root/reference with the three missing wheels
shimmed (TEST INFRASTRUCTURE — oracle side; usable only where /root/reference is mounted).

  * `onnxruntime`  -> oracle.torch_interp.TorchSession   (predict_base.py:1,14)
  * `shapely.geometry.Polygon`, `pyclipper` -> oracle.clipper_offset  (db_postprocess.py:25-26)

Used by the golden generator (tests/golden/make_goldens.py) and by tests marked `reference`, which
validate the numpy/cv2 restatements in oracle/restate_*.py against the reference's own code.
"""
import os
import sys
import types

REFERENCE = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REFERENCE, "onnxocr"))


def install_shims(session_factory=None):
    from . import clipper_offset, torch_interp
    ort = types.ModuleType("onnxruntime")
    ort.InferenceSession = session_factory or torch_interp.TorchSession
    ort.get_device = lambda: "CPU"
    sys.modules["onnxruntime"] = ort
    shp = types.ModuleType("shapely")
    geo = types.ModuleType("shapely.geometry")
    geo.Polygon = clipper_offset.Polygon
    shp.geometry = geo
    sys.modules["shapely"] = shp
    sys.modules["shapely.geometry"] = geo
    pc = types.ModuleType("pyclipper")
    pc.PyclipperOffset = clipper_offset.PyclipperOffset
    pc.JT_ROUND = clipper_offset.JT_ROUND
    pc.ET_CLOSEDPOLYGON = clipper_offset.ET_CLOSEDPOLYGON
    sys.modules["pyclipper"] = pc


def import_reference(session_factory=None):
    """Returns the reference's `onnxocr` package (imported from the read-only mount)."""
    if not available():
        raise RuntimeError("reference checkout not mounted at %s" % REFERENCE)
    install_shims(session_factory)
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    import onnxocr  # noqa: F401
    from onnxocr import db_postprocess, operators, predict_cls, predict_det, predict_rec, predict_system, rec_postprocess, utils  # noqa: F401
    return onnxocr


def make_ocr(model_paths, **kwargs):
    """ONNXPaddleOcr(...) of the reference wired to the given det/cls/rec model files."""
    import_reference()
    from onnxocr.onnx_paddleocr import ONNXPaddleOcr
    args = dict(use_angle_cls=True, use_gpu=False, det_model_dir=model_paths["det"], cls_model_dir=model_paths["cls"],
                rec_model_dir=model_paths["rec"], rec_char_dict_path=model_paths["dict"])
    args.update(kwargs)
    return ONNXPaddleOcr(**args)

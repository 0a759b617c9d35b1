"""Generates the committed golden vectors by running the UNMODIFIED reference (imported from /root/reference with the
oracle shims) on seeded inputs.  Run in the authoring container only:

    python tests/golden/make_goldens.py

Outputs (small, committed):  tests/golden/stage_goldens.npz, tests/golden/e2e_goldens.json
The inputs are regenerated from seeds by tests/test_goldens.py, so only outputs are stored.
"""
import argparse
import json
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import golden_inputs as gi  # noqa: E402
from oracle import ref_harness  # noqa: E402


def main():
    from onnxocr_b200 import synth_graphs
    ref_harness.import_reference()
    from onnxocr.db_postprocess import DBPostProcess
    from onnxocr.imaug import create_operators, transform
    from onnxocr.predict_cls import TextClassifier
    from onnxocr.predict_rec import TextRecognizer
    from onnxocr.rec_postprocess import CTCLabelDecode
    from onnxocr.utils import get_rotate_crop_image, infer_args
    paths = synth_graphs.ensure_models()
    paths["cls"] = os.path.join(synth_graphs.ASSET_DIR, "ppocrv5_cls.onnx")
    paths["dict"] = os.path.join(synth_graphs.ASSET_DIR, "ppocrv5_dict.txt")
    out = {}
    # 1. det pre-processing (operators.py)
    ops = create_operators([{"DetResizeForTest": {"limit_side_len": 960, "limit_type": "max"}},
                            {"NormalizeImage": {"std": [0.229, 0.224, 0.225], "mean": [0.485, 0.456, 0.406], "scale": "1./255.", "order": "hwc"}},
                            {"ToCHWImage": None}, {"KeepKeys": {"keep_keys": ["image", "shape"]}}])
    for k, img in enumerate(gi.det_images()):
        chw, shape = transform({"image": img}, ops)
        out["det_pre_%d" % k] = np.ascontiguousarray(chw)
        out["det_shape_%d" % k] = shape
    # 2. DBPostProcess on seeded maps (db_postprocess.py)
    pp = DBPostProcess(thresh=0.3, box_thresh=0.6, max_candidates=1000, unclip_ratio=1.5)
    for k, (prob, src) in enumerate(gi.prob_maps()):
        h, w = prob.shape
        shape_list = np.array([[float(src[0]), float(src[1]), h / src[0], w / src[1]]])
        out["db_boxes_%d" % k] = pp({"maps": prob[None, None]}, shape_list)[0]["points"]
    # 2b. the same maps with det_db_score_mode='slow' (box_score_slow, db_postprocess.py:199-218) and a lower box threshold
    pp_slow = DBPostProcess(thresh=0.3, box_thresh=0.5, max_candidates=1000, unclip_ratio=1.5, score_mode="slow")
    for k, (prob, src) in enumerate(gi.prob_maps()):
        h, w = prob.shape
        shape_list = np.array([[float(src[0]), float(src[1]), h / src[0], w / src[1]]])
        out["db_boxes_slow_%d" % k] = pp_slow({"maps": prob[None, None]}, shape_list)[0]["points"]
    # 3. crops (utils.get_rotate_crop_image) and rec/cls pre-processing
    args = argparse.Namespace(**{a.dest: a.default for a in infer_args()._actions})
    args.rec_image_shape = "3, 48, 320"
    args.use_gpu = False
    args.rec_model_dir, args.cls_model_dir, args.det_model_dir = paths["rec"], paths["cls"], paths["det"]
    args.rec_char_dict_path = paths["dict"]
    rec, cls = TextRecognizer(args), TextClassifier(args)
    img, quads = gi.crop_case()
    crops = [get_rotate_crop_image(img, q.copy()) for q in quads]
    for k, c in enumerate(crops):
        out["crop_%d" % k] = c
    # 4. cls on real weights: TextClassifier.__call__ (labels + scores)
    _, cls_res = cls(crops)
    out["cls_labels"] = np.array([int(r[0] == "180") for r in cls_res], np.int32)
    out["cls_scores"] = np.array([r[1] for r in cls_res], np.float32)
    # 5. CTC decode on seeded probabilities
    dec = CTCLabelDecode(character_dict_path=paths["dict"], use_space_char=True)
    res = dec(gi.ctc_probs())
    np.savez_compressed(os.path.join(HERE, "stage_goldens.npz"), **out)
    e2e = {"ctc": [[t, s] for t, s in res]}
    # 6. end to end with the committed synthetic det/rec weights + real cls
    model = ref_harness.make_ocr(paths)
    e2e["ocr"] = []
    for page in gi.e2e_pages():
        r = model.ocr(page.copy())
        e2e["ocr"].append([[b, [t, s]] for b, (t, s) in r[0]])
    e2e["rec_only"] = [[t, s] for t, s in rec(crops)]
    with open(os.path.join(HERE, "e2e_goldens.json"), "w") as fh:
        json.dump(e2e, fh, ensure_ascii=False, indent=0)
    print("wrote goldens:", {k: v.shape for k, v in out.items() if k.startswith("db_")}, len(e2e["ocr"][0]), "lines on page 0")


if __name__ == "__main__":
    main()

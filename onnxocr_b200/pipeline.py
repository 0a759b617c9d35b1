"""Level B host mirror of the reference's stage classes: same names, signatures and return shapes as
`onnxocr/predict_det.py`, `predict_cls.py`, `predict_rec.py`, `predict_system.py`, `onnx_paddleocr.py`, with each
`__call__` executing as fused device pipelines through libb200ocr.so instead of Python loops around session.run.
"""
import argparse
import copy
import os
import warnings

import numpy as np

from . import _lib, session, synth_graphs


def default_args(**kwargs):
    """The subset of `utils.infer_args()` defaults the hot path reads (utils.py:251-387, onnx_paddleocr.py:13-23)."""
    assets = synth_graphs.ASSET_DIR
    d = dict(use_gpu=True, gpu_id=0, det_algorithm="DB", det_model_dir=None, det_limit_side_len=960, det_limit_type="max",
             det_box_type="quad", det_db_thresh=0.3, det_db_box_thresh=0.6, det_db_unclip_ratio=1.5, use_dilation=False,
             det_db_score_mode="fast", rec_algorithm="SVTR_LCNet", rec_model_dir=None, rec_image_shape="3, 48, 320",
             rec_batch_num=6, rec_char_dict_path=os.path.join(assets, "ppocrv5_dict.txt"), use_space_char=True,
             drop_score=0.5, use_angle_cls=False, cls_model_dir=os.path.join(assets, "ppocrv5_cls.onnx"),
             cls_image_shape="3, 48, 192", label_list=["0", "180"], cls_batch_num=6, cls_thresh=0.9,
             save_crop_res=False, precision="fp32", fused=True, rec_width_bucket=None)
    d.update(kwargs)
    if d["rec_width_bucket"] is None:
        # fp32 = the exact mode: every crop padded exactly as the reference's batch plan pads it.  bf16 = the fast mode
        # (BASELINE.json: "width-bucketed by aspect ratio"): padded widths rounded up to a multiple of 32, which merges
        # the recogniser's batches across pages (+20-30 % end to end); parity there is statistical anyway.
        d["rec_width_bucket"] = 32 if str(d["precision"]).lower() == "bf16" else 0
    if d["det_model_dir"] is None or d["rec_model_dir"] is None:
        # The reference defaults to its shipped PP-OCRv5 det/rec graphs (utils.py:251-387); those blobs are not in this
        # checkout, so the stand-ins are architecture-faithful graphs whose weights were fitted on rendered ASCII pages.
        # They are meant for throughput runs and tests: say so loudly instead of returning nonsense on real images.
        missing = [k for k in ("det_model_dir", "rec_model_dir") if d[k] is None]
        warnings.warn("onnxocr_b200: %s not given - falling back to the SYNTHETIC PP-OCRv5 stand-in graph(s) "
                      "(weights fitted on rendered ASCII pages; text read from real images will be meaningless). "
                      "Pass the real det.onnx / rec.onnx paths for production use." % " and ".join(missing),
                      RuntimeWarning, stacklevel=3)
        paths = synth_graphs.ensure_models()
        d["det_model_dir"] = d["det_model_dir"] or paths["det"]
        d["rec_model_dir"] = d["rec_model_dir"] or paths["rec"]
    return argparse.Namespace(**d)


def _precision(args):
    return _lib.BF16 if str(getattr(args, "precision", "fp32")).lower() == "bf16" else _lib.FP32


def _engine_for(args):
    eng = getattr(args, "_b200_engine", None)
    if eng is None:
        eng = _lib.Engine(int(getattr(args, "gpu_id", 0) or 0))
        args._b200_engine = eng
    return eng


class TextDetector(object):
    """predict_det.py:6-120"""

    def __init__(self, args):
        if args.det_algorithm != "DB" or args.det_box_type != "quad":
            raise NotImplementedError("only DB / quad boxes are on the B200 path (SURVEY.md §8f)")
        self.args = args
        self.engine = _engine_for(args)
        self.engine.load_onnx(_lib.DET, args.det_model_dir, _precision(args))
        self.params = _lib.det_params(args.det_limit_side_len, args.det_limit_type, args.det_db_thresh,
                                      args.det_db_box_thresh, args.det_db_unclip_ratio, 1000, 3, args.use_dilation,
                                      args.det_db_score_mode)

    def __call__(self, img):
        return self.engine.det_pages([img], self.params)[0]

    def batch(self, imgs):
        return self.engine.det_pages(list(imgs), self.params)


class TextClassifier(object):
    """predict_cls.py:10-89"""

    def __init__(self, args):
        self.cls_batch_num = args.cls_batch_num
        self.cls_thresh = args.cls_thresh
        self.label_list = args.label_list
        self.engine = _engine_for(args)
        self.engine.load_onnx(_lib.CLS, args.cls_model_dir, _precision(args))

    def __call__(self, img_list):
        import cv2
        img_list = copy.deepcopy(img_list)
        labels, scores = self.engine.cls_crops(img_list, self.cls_batch_num)
        cls_res = []
        for i, (l, s) in enumerate(zip(labels, scores)):
            label = self.label_list[int(l)]
            cls_res.append([label, s])
            if "180" in label and s > self.cls_thresh:
                img_list[i] = cv2.rotate(img_list[i], 1)
        return img_list, cls_res


class CTCLabelDecode(object):
    """rec_postprocess.py:9-37,111-113: index -> character table (blank + dict lines + ' ')."""

    def __init__(self, character_dict_path, use_space_char=True):
        chars = []
        with open(character_dict_path, "rb") as fin:
            for line in fin.readlines():
                chars.append(line.decode("utf-8").strip("\n").strip("\r\n"))
        if use_space_char:
            chars.append(" ")
        self.character = ["blank"] + chars

    def text(self, ids):
        return "".join(self.character[i] for i in ids)


class TextRecognizer(object):
    """predict_rec.py:11-24, 280-326"""

    def __init__(self, args):
        self.rec_batch_num = args.rec_batch_num
        self.postprocess_op = CTCLabelDecode(args.rec_char_dict_path, args.use_space_char)
        self.engine = _engine_for(args)
        self.engine.load_onnx(_lib.REC, args.rec_model_dir, _precision(args))

    def __call__(self, img_list):
        ids, scores = self.engine.rec_crops(img_list, self.rec_batch_num)
        return [(self.postprocess_op.text(i), float(s)) for i, s in zip(ids, scores)]


def sorted_boxes(dt_boxes):
    """predict_system.py:75-97 (the exact, non-transitive bubble pass)."""
    num_boxes = dt_boxes.shape[0]
    _boxes = list(sorted(dt_boxes, key=lambda x: (x[0][1], x[0][0])))
    for i in range(num_boxes - 1):
        for j in range(i, -1, -1):
            if abs(_boxes[j + 1][0][1] - _boxes[j][0][1]) < 10 and (_boxes[j + 1][0][0] < _boxes[j][0][0]):
                _boxes[j], _boxes[j + 1] = _boxes[j + 1], _boxes[j]
            else:
                break
    return _boxes


class TextSystem(object):
    """predict_system.py:10-72.  `__call__` keeps the reference's signature; by default it runs as ONE fused device
    pipeline (`b200ocr_ocr_pages`: det -> crops -> cls -> rec without leaving the GPU).  `fused=False` in the args runs
    the stage objects one after the other exactly like the reference does (same results, more PCIe round trips);
    `batch(imgs)` is the Level-C throughput entry for many pages per call."""

    def __init__(self, args):
        self.text_detector = TextDetector(args)
        self.text_recognizer = TextRecognizer(args)
        self.use_angle_cls = args.use_angle_cls
        self.drop_score = args.drop_score
        if self.use_angle_cls:
            self.text_classifier = TextClassifier(args)
        self.args = args
        self.engine = _engine_for(args)
        self.fused = bool(getattr(args, "fused", True))
        self.rec_width_bucket = int(getattr(args, "rec_width_bucket", 0) or 0)

    def _ocr_params(self, cls):
        a = self.args
        return _lib.ocr_params(self.text_detector.params, self.use_angle_cls and cls, a.cls_thresh, a.label_list,
                               a.rec_batch_num, self.rec_width_bucket)

    def batch(self, imgs, cls=True):
        """[(filter_boxes, filter_rec_res), ...] for a list of pages; one library call."""
        decode = self.text_recognizer.postprocess_op.text
        out = []
        for page in self.engine.ocr_pages(list(imgs), self._ocr_params(cls)):
            fb, fr = [], []
            for box, ids, score in zip(page["boxes"], page["ids"], page["scores"]):
                if score >= self.drop_score:
                    fb.append(box)
                    fr.append((decode(ids), float(score)))
            out.append((fb, fr))
        return out

    def __call__(self, img, cls=True):
        if self.fused:
            return self.batch([img], cls)[0]
        dt_boxes = self.text_detector(img)
        if dt_boxes is None:
            return None, None
        dt_boxes = sorted_boxes(dt_boxes) if len(dt_boxes) else []
        img_crop_list = self.engine.crop_boxes(img, np.array(dt_boxes, np.float32)) if len(dt_boxes) else []
        if self.use_angle_cls and cls:
            img_crop_list, _ = self.text_classifier(img_crop_list)
        rec_res = self.text_recognizer(img_crop_list)
        filter_boxes, filter_rec_res = [], []
        for box, rec_result in zip(dt_boxes, rec_res):
            if rec_result[1] >= self.drop_score:
                filter_boxes.append(box)
                filter_rec_res.append(rec_result)
        return filter_boxes, filter_rec_res


class ONNXPaddleOcr(TextSystem):
    """onnx_paddleocr.py:10-61"""

    def __init__(self, **kwargs):
        super().__init__(default_args(**kwargs))

    def ocr(self, img, det=True, rec=True, cls=True):
        if det and rec:
            dt_boxes, rec_res = self.__call__(img, cls)
            return [[[box.tolist(), res] for box, res in zip(dt_boxes, rec_res)]]
        if det and not rec:
            return [[box.tolist() for box in self.text_detector(img)]]
        if not isinstance(img, list):
            img = [img]
        cls_res = []
        if self.use_angle_cls and cls:
            img, cls_res_tmp = self.text_classifier(img)
            if not rec:
                cls_res.append(cls_res_tmp)
        rec_res = self.text_recognizer(img)
        if not rec:
            return cls_res
        return [rec_res]

    def ocr_batch(self, imgs, cls=True):
        """`ocr(img)` for a list of pages in one fused call: [[[box, (text, score)], ...], ...] (one entry per page)."""
        return [[[box.tolist(), res] for box, res in zip(b, r)] for b, r in self.batch(imgs, cls)]

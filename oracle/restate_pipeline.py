"""End-to-end restatement of `TextSystem.__call__` / `ONNXPaddleOcr.ocr` (TEST INFRASTRUCTURE — oracle side).

predict_system.py:35-72 with the stages of predict_det.py:94-120, predict_cls.py:44-89, predict_rec.py:280-326 built from
the restatements in restate_pre.py / restate_post.py and a session factory (torch_interp.TorchGraph by default).
Validated against the imported reference in tests/test_oracle_vs_reference.py.
"""
import numpy as np
import cv2

from . import restate_post, restate_pre
from .torch_interp import TorchGraph


class OraclePipeline(object):
    def __init__(self, model_paths, use_angle_cls=True, drop_score=0.5, rec_batch_num=6, cls_batch_num=6, cls_thresh=0.9,
                 det_limit_side_len=960, det_limit_type="max"):
        self.det = TorchGraph(model_paths["det"])
        self.rec = TorchGraph(model_paths["rec"])
        self.cls = TorchGraph(model_paths["cls"]) if use_angle_cls else None
        self.charset = restate_post.load_charset(model_paths["dict"])
        self.use_angle_cls = use_angle_cls
        self.drop_score = drop_score
        self.rec_batch_num = rec_batch_num
        self.cls_batch_num = cls_batch_num
        self.cls_thresh = cls_thresh
        self.limit = (det_limit_side_len, det_limit_type)

    # predict_det.py:94-120
    def detect(self, img, return_map=False):
        x, shape = restate_pre.det_preprocess(img, *self.limit)
        prob = self.det.run(None, {"x": x})[0]
        boxes, _ = restate_post.db_postprocess(prob, shape[None])[0]
        out = restate_post.filter_tag_det_res(boxes, img.shape)
        return (out, prob[0, 0]) if return_map else out

    # predict_cls.py:44-89
    def classify(self, img_list):
        img_list = [im.copy() for im in img_list]
        ratios = np.array([im.shape[1] / float(im.shape[0]) for im in img_list])
        order = np.argsort(ratios)
        res = [["", 0.0]] * len(img_list)
        for beg in range(0, len(img_list), self.cls_batch_num):
            idx = order[beg:beg + self.cls_batch_num]
            x = np.stack([restate_pre.resize_norm_img(img_list[i], 192) for i in idx]).astype(np.float32)
            prob = self.cls.run(None, {"x": x})[0]
            for k, i in enumerate(idx):
                lab = int(prob[k].argmax())
                score = prob[k, lab]
                label = ["0", "180"][lab]
                res[i] = [label, score]
                if "180" in label and score > self.cls_thresh:
                    img_list[i] = cv2.rotate(img_list[i], 1)
        return img_list, res

    # predict_rec.py:280-326
    def recognize(self, img_list):
        res = [["", 0.0]] * len(img_list)
        plan = restate_pre.rec_batches([(im.shape[1], im.shape[0]) for im in img_list], self.rec_batch_num)
        for idx, img_w in plan:
            x = np.stack([restate_pre.resize_norm_img(img_list[i], img_w) for i in idx]).astype(np.float32)
            preds = self.rec.run(None, {"x": x})[0]
            ids, scores, texts = restate_post.ctc_decode(preds, self.charset)
            for k, i in enumerate(idx):
                res[i] = (texts[k], scores[k])
        return res

    # predict_system.py:35-72
    def __call__(self, img, cls=True):
        dt_boxes = self.detect(img)
        dt_boxes = restate_post.sorted_boxes(dt_boxes) if len(dt_boxes) else []
        crops = [restate_pre.get_rotate_crop_image(img, b.copy()) for b in dt_boxes]
        if self.use_angle_cls and cls:
            crops, _ = self.classify(crops)
        rec_res = self.recognize(crops)
        boxes, out = [], []
        for box, r in zip(dt_boxes, rec_res):
            if r[1] >= self.drop_score:
                boxes.append(box)
                out.append(r)
        return boxes, out

    # onnx_paddleocr.py:28-39
    def ocr(self, img, cls=True):
        boxes, res = self(img, cls)
        return [[[b.tolist(), r] for b, r in zip(boxes, res)]]

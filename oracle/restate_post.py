"""numpy/cv2 restatement of the post-processing half of the hot path (TEST INFRASTRUCTURE — oracle side).

Each function cites the reference lines it follows (paths relative to /root/reference).  cv2 (the
reference's own dependency, same build on the GPU box) performs findContours / minAreaRect / boxPoints /
fillPoly / mean exactly as in the reference; Clipper and shapely are restated in clipper_offset.py
(parity unpinned there).  Validated against the imported reference by tests/test_oracle_vs_reference.py.
"""
import numpy as np
import cv2

from . import clipper_offset


# ------------------------------------------------------------------ DBPostProcess (db_postprocess.py)
def get_mini_boxes(contour):
    """db_postprocess.py:159-180"""
    bounding_box = cv2.minAreaRect(contour)
    points = sorted(list(cv2.boxPoints(bounding_box)), key=lambda x: x[0])
    if points[1][1] > points[0][1]:
        index_1, index_4 = 0, 1
    else:
        index_1, index_4 = 1, 0
    if points[3][1] > points[2][1]:
        index_2, index_3 = 2, 3
    else:
        index_2, index_3 = 3, 2
    box = [points[index_1], points[index_2], points[index_3], points[index_4]]
    return box, min(bounding_box[1])


def box_score_fast(bitmap, _box):
    """db_postprocess.py:182-197"""
    h, w = bitmap.shape[:2]
    box = _box.copy()
    xmin = np.clip(np.floor(box[:, 0].min()).astype("int32"), 0, w - 1)
    xmax = np.clip(np.ceil(box[:, 0].max()).astype("int32"), 0, w - 1)
    ymin = np.clip(np.floor(box[:, 1].min()).astype("int32"), 0, h - 1)
    ymax = np.clip(np.ceil(box[:, 1].max()).astype("int32"), 0, h - 1)
    mask = np.zeros((ymax - ymin + 1, xmax - xmin + 1), dtype=np.uint8)
    box[:, 0] = box[:, 0] - xmin
    box[:, 1] = box[:, 1] - ymin
    cv2.fillPoly(mask, box.reshape(1, -1, 2).astype("int32"), 1)
    return cv2.mean(bitmap[ymin:ymax + 1, xmin:xmax + 1], mask)[0]


def box_score_slow(bitmap, contour):
    """db_postprocess.py:199-218"""
    h, w = bitmap.shape[:2]
    contour = contour.copy()
    contour = np.reshape(contour, (-1, 2))
    xmin = np.clip(np.min(contour[:, 0]), 0, w - 1)
    xmax = np.clip(np.max(contour[:, 0]), 0, w - 1)
    ymin = np.clip(np.min(contour[:, 1]), 0, h - 1)
    ymax = np.clip(np.max(contour[:, 1]), 0, h - 1)
    mask = np.zeros((ymax - ymin + 1, xmax - xmin + 1), dtype=np.uint8)
    contour[:, 0] = contour[:, 0] - xmin
    contour[:, 1] = contour[:, 1] - ymin
    cv2.fillPoly(mask, contour.reshape(1, -1, 2).astype("int32"), 1)
    return cv2.mean(bitmap[ymin:ymax + 1, xmin:xmax + 1], mask)[0]


def unclip(box, unclip_ratio):
    """db_postprocess.py:151-157"""
    poly = clipper_offset.Polygon(box)
    distance = poly.area * unclip_ratio / poly.length
    offset = clipper_offset.PyclipperOffset()
    offset.AddPath(box, clipper_offset.JT_ROUND, clipper_offset.ET_CLOSEDPOLYGON)
    return np.array(offset.Execute(distance))


def boxes_from_bitmap(pred, bitmap, dest_width, dest_height, box_thresh=0.6, unclip_ratio=1.5, max_candidates=1000,
                      min_size=3, score_mode="fast"):
    """db_postprocess.py:104-149 (box_type 'quad')"""
    height, width = bitmap.shape
    outs = cv2.findContours((bitmap * 255).astype(np.uint8), cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
    contours = outs[0] if len(outs) == 2 else outs[1]
    num_contours = min(len(contours), max_candidates)
    boxes, scores = [], []
    for index in range(num_contours):
        contour = contours[index]
        points, sside = get_mini_boxes(contour)
        if sside < min_size:
            continue
        points = np.array(points)
        if score_mode == "fast":
            score = box_score_fast(pred, points.reshape(-1, 2))
        else:
            score = box_score_slow(pred, contour)
        if box_thresh > score:
            continue
        box = unclip(points, unclip_ratio).reshape(-1, 1, 2)
        box, sside = get_mini_boxes(box)
        if sside < min_size + 2:
            continue
        box = np.array(box)
        box[:, 0] = np.clip(np.round(box[:, 0] / width * dest_width), 0, dest_width)
        box[:, 1] = np.clip(np.round(box[:, 1] / height * dest_height), 0, dest_height)
        boxes.append(box.astype("int32"))
        scores.append(score)
    return np.array(boxes, dtype="int32"), scores


def db_postprocess(maps, shape_list, thresh=0.3, box_thresh=0.6, unclip_ratio=1.5, max_candidates=1000,
                   use_dilation=False, score_mode="fast"):
    """DBPostProcess.__call__ db_postprocess.py:220-246.  maps: [N,1,H,W] float32; shape_list: [N,4] float64."""
    pred = maps[:, 0, :, :]
    segmentation = pred > thresh
    out = []
    for b in range(pred.shape[0]):
        src_h, src_w, _, _ = shape_list[b]
        if use_dilation:
            mask = cv2.dilate(np.array(segmentation[b]).astype(np.uint8), np.array([[1, 1], [1, 1]]))
        else:
            mask = segmentation[b]
        boxes, scores = boxes_from_bitmap(pred[b], mask, src_w, src_h, box_thresh, unclip_ratio, max_candidates,
                                          score_mode=score_mode)
        out.append((boxes, scores))
    return out


# ------------------------------------------------------------------ TextDetector tail (predict_det.py:50-81)
def order_points_clockwise(pts):
    rect = np.zeros((4, 2), dtype="float32")
    s = pts.sum(axis=1)
    rect[0] = pts[np.argmin(s)]
    rect[2] = pts[np.argmax(s)]
    tmp = np.delete(pts, (np.argmin(s), np.argmax(s)), axis=0)
    diff = np.diff(np.array(tmp), axis=1)
    rect[1] = tmp[np.argmin(diff)]
    rect[3] = tmp[np.argmax(diff)]
    return rect


def clip_det_res(points, img_height, img_width):
    for pno in range(points.shape[0]):
        points[pno, 0] = int(min(max(points[pno, 0], 0), img_width - 1))
        points[pno, 1] = int(min(max(points[pno, 1], 0), img_height - 1))
    return points


def filter_tag_det_res(dt_boxes, image_shape):
    img_height, img_width = image_shape[0:2]
    out = []
    for box in dt_boxes:
        box = order_points_clockwise(np.array(box))
        box = clip_det_res(box, img_height, img_width)
        rect_width = int(np.linalg.norm(box[0] - box[1]))
        rect_height = int(np.linalg.norm(box[0] - box[3]))
        if rect_width <= 3 or rect_height <= 3:
            continue
        out.append(box)
    return np.array(out)


# ------------------------------------------------------------------ TextSystem (predict_system.py:75-97)
def sorted_boxes(dt_boxes):
    num_boxes = dt_boxes.shape[0]
    _boxes = list(sorted(dt_boxes, key=lambda x: (x[0][1], x[0][0])))
    for i in range(num_boxes - 1):
        for j in range(i, -1, -1):
            if abs(_boxes[j + 1][0][1] - _boxes[j][0][1]) < 10 and (_boxes[j + 1][0][0] < _boxes[j][0][0]):
                _boxes[j], _boxes[j + 1] = _boxes[j + 1], _boxes[j]
            else:
                break
    return _boxes


# ------------------------------------------------------------------ CTCLabelDecode (rec_postprocess.py:58-113)
def load_charset(dict_path, use_space_char=True):
    chars = []
    with open(dict_path, "rb") as fin:
        for line in fin.readlines():
            chars.append(line.decode("utf-8").strip("\n").strip("\r\n"))
    if use_space_char:
        chars.append(" ")
    return ["blank"] + chars


def ctc_decode(preds, charset=None):
    """CTCLabelDecode.__call__: argmax/max over classes, drop repeats and blanks, mean prob of the kept steps.
    Returns (list of index arrays, list of float scores[, list of strings])."""
    idx = preds.argmax(axis=2)
    prob = preds.max(axis=2)
    ids, scores, texts = [], [], []
    for b in range(idx.shape[0]):
        sel = np.ones(idx.shape[1], dtype=bool)
        sel[1:] = idx[b][1:] != idx[b][:-1]
        sel &= idx[b] != 0
        ids.append(idx[b][sel])
        conf = prob[b][sel]
        if len(conf) == 0:
            conf = [0]
        scores.append(np.mean(conf).tolist())
        if charset is not None:
            texts.append("".join(charset[i] for i in idx[b][sel]))
    return (ids, scores, texts) if charset is not None else (ids, scores)

"""ctypes binding of libb200ocr.so (include/b200ocr.h).  numpy in / numpy out.

The library is the product: importing this module never falls back to a CPU path.
If the shared object is missing or no CUDA device is visible, the error is raised.
"""
import ctypes as C
import os
import threading
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200ocr.so")

DET, CLS, REC = 0, 1, 2
FP32, BF16 = 0, 1

_lib = None
_lock = threading.Lock()


class DetParams(C.Structure):
    _fields_ = [("limit_side_len", C.c_float), ("limit_type", C.c_int), ("thresh", C.c_float),
                ("max_candidates", C.c_int), ("min_size", C.c_int), ("use_dilation", C.c_int),
                ("box_thresh", C.c_double), ("unclip_ratio", C.c_double), ("score_mode", C.c_int)]


class OcrParams(C.Structure):
    _fields_ = [("det", DetParams), ("use_angle_cls", C.c_int), ("cls_thresh", C.c_float),
                ("cls_rot_label_mask", C.c_uint), ("rec_batch_num", C.c_int), ("rec_width_bucket", C.c_int)]


_LIMIT_TYPES = {"max": 0, "min": 1, "resize_long": 2}
_SCORE_MODES = {"fast": 0, "slow": 1}      # det_db_score_mode, utils.py:283


def det_params(limit_side_len=960, limit_type="max", thresh=0.3, box_thresh=0.6, unclip_ratio=1.5,
               max_candidates=1000, min_size=3, use_dilation=False, score_mode="fast"):
    if score_mode not in _SCORE_MODES:
        raise ValueError("score_mode can only be one of %s" % list(_SCORE_MODES))
    return DetParams(float(limit_side_len), _LIMIT_TYPES[limit_type], float(thresh), int(max_candidates),
                     int(min_size), int(bool(use_dilation)), float(box_thresh), float(unclip_ratio),
                     _SCORE_MODES[score_mode])


def ocr_params(det=None, use_angle_cls=False, cls_thresh=0.9, label_list=("0", "180"), rec_batch_num=6,
               rec_width_bucket=0):
    mask = 0
    for i, lab in enumerate(label_list):
        if "180" in lab:
            mask |= 1 << i
    return OcrParams(det or det_params(), int(bool(use_angle_cls)), float(cls_thresh), mask, int(rec_batch_num),
                     int(rec_width_bucket))


def lib():
    """Loads libb200ocr.so once (building is `python -m onnxocr_b200.build` / __graft_entry__.build())."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libb200ocr.so is not built (%s); run `python -m onnxocr_b200.build`. "
                               "There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        vp, i, f = C.c_void_p, C.c_int, C.c_float
        pi, pf, pd = C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_double)
        pi64 = C.POINTER(C.c_int64)
        sz = C.c_size_t
        L.b200ocr_version.restype = C.c_char_p
        L.b200ocr_create.argtypes = [i, C.POINTER(vp)]
        L.b200ocr_destroy.argtypes = [vp]
        L.b200ocr_destroy.restype = None
        L.b200ocr_last_error.argtypes = [vp]
        L.b200ocr_last_error.restype = C.c_char_p
        L.b200ocr_load_onnx.argtypes = [vp, i, C.c_char_p, i]
        L.b200ocr_io_info.argtypes = [vp, i, C.c_char_p, sz, C.c_char_p, sz, pi64, pi64, pi]
        L.b200ocr_session_run.argtypes = [vp, i, vp, pi64, vp, sz, pi64, pi]
        L.b200ocr_det_params_default.argtypes = [C.POINTER(DetParams)]
        L.b200ocr_det_params_default.restype = None
        L.b200ocr_det_preprocess.argtypes = [vp, vp, i, i, i, C.POINTER(DetParams), vp, sz, pi, pi, pd]
        L.b200ocr_db_postprocess.argtypes = [vp, vp, i, i, i, i, C.POINTER(DetParams), vp, vp, i, pi]
        L.b200ocr_det_pages.argtypes = [vp, C.POINTER(vp), pi, pi, pi, i, C.POINTER(DetParams), vp, i, pi, vp, sz, pi]
        L.b200ocr_crop_boxes.argtypes = [vp, vp, i, i, i, vp, i, vp, sz, pi]
        L.b200ocr_cls_crops.argtypes = [vp, C.POINTER(vp), pi, pi, i, i, vp, vp]
        L.b200ocr_rec_crops.argtypes = [vp, C.POINTER(vp), pi, pi, i, i, vp, vp, i, pi, vp]
        L.b200ocr_rec_argmax.argtypes = [vp, vp, pi64, vp, vp, i, pi]
        L.b200ocr_ocr_params_default.argtypes = [C.POINTER(OcrParams)]
        L.b200ocr_ocr_params_default.restype = None
        L.b200ocr_ocr_pages.argtypes = [vp, C.POINTER(vp), pi, pi, pi, i, C.POINTER(OcrParams), vp, i, pi, vp, i, vp, vp, vp, vp]
        L.b200ocr_describe.argtypes = [vp, i, C.c_char_p, sz]
        L.b200ocr_plan_stats.argtypes = [vp, i, i, i, i, pd, pi, pd]
        L.b200ocr_profile_plan.argtypes = [vp, i, i, i, i, i, pf, pd, pd, i, C.c_char_p, sz]
        L.b200ocr_time_plan.argtypes = [vp, i, i, i, i, i, i, pf]
        L.b200ocr_l2_flush.argtypes = [vp]
        L.b200ocr_host_alloc.argtypes = [sz]
        L.b200ocr_host_alloc.restype = vp
        L.b200ocr_host_free.argtypes = [vp]
        L.b200ocr_host_free.restype = None
        L.b200ocr_launch_count.argtypes = [vp]
        L.b200ocr_launch_count.restype = C.c_int64
        L.b200ocr_stage_times.argtypes = [vp, pf]
        L.b200ocr_ocr_stage_times.argtypes = [vp, pf]
        L.b200ocr_det_resident_upload.argtypes = [vp, C.POINTER(vp), pi, pi, pi, i]
        L.b200ocr_det_resident_run.argtypes = [vp, C.POINTER(DetParams), i, pf, pi, pf]
        L.b200ocr_trace_dump.argtypes = []
        L.b200ocr_trace_dump.restype = None
        L.b200ocr_det_resident_stream.argtypes = [vp, C.POINTER(DetParams), i, i, pf, pi]
        L.b200ocr_det_submit.argtypes = [vp, C.POINTER(vp), pi, pi, pi, i, C.POINTER(DetParams), pi]
        L.b200ocr_det_collect.argtypes = [vp, i, vp, i, pi]
        _lib = L
        return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _int_arr(v):
    return (C.c_int * len(v))(*[int(x) for x in v])


class Engine(object):
    """One engine per CUDA device (b200ocr_create / b200ocr_destroy)."""

    def __init__(self, device=0):
        self._L = lib()
        h = C.c_void_p()
        rc = self._L.b200ocr_create(int(device), C.byref(h))
        if rc != 0:
            raise RuntimeError("b200ocr_create failed (%d): %s" % (rc, self._L.b200ocr_last_error(None).decode()))
        self._h = h
        self.device = device
        self.loaded = {}

    def close(self):
        if getattr(self, "_h", None):
            self._L.b200ocr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise RuntimeError("libb200ocr error %d: %s" % (rc, self._L.b200ocr_last_error(self._h).decode()))
        return rc

    # ---- Level A -----------------------------------------------------------------
    def load_onnx(self, kind, path, precision=FP32):
        self._check(self._L.b200ocr_load_onnx(self._h, kind, os.fsencode(path), precision))
        self.loaded[kind] = (path, precision)

    def io_info(self, kind):
        a, b = C.create_string_buffer(256), C.create_string_buffer(256)
        di, do = (C.c_int64 * 4)(), (C.c_int64 * 4)()
        r = C.c_int()
        self._check(self._L.b200ocr_io_info(self._h, kind, a, 256, b, 256, di, do, C.byref(r)))
        return a.value.decode(), list(di), b.value.decode(), list(do)[:r.value]

    def session_run(self, kind, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 4:
            raise ValueError("input must be NCHW")
        dims = (C.c_int64 * 4)(*x.shape)
        od = (C.c_int64 * 4)()
        r = C.c_int()
        self._check(self._L.b200ocr_session_run(self._h, kind, _ptr(x), dims, None, 0, od, C.byref(r)))
        shape = list(od)[:r.value]
        out = np.empty(shape, dtype=np.float32)
        self._check(self._L.b200ocr_session_run(self._h, kind, _ptr(x), dims, _ptr(out), out.size, od, C.byref(r)))
        return out

    def rec_argmax(self, x):
        """rec session.run fused with argmax/max over the classes: float32 NCHW -> (int32 [n,T], float32 [n,T])."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        dims = (C.c_int64 * 4)(*x.shape)
        t = C.c_int()
        self._check(self._L.b200ocr_rec_argmax(self._h, _ptr(x), dims, None, None, 0, C.byref(t)))
        idx = np.zeros((x.shape[0], t.value), np.int32)
        prob = np.zeros((x.shape[0], t.value), np.float32)
        self._check(self._L.b200ocr_rec_argmax(self._h, _ptr(x), dims, _ptr(idx), _ptr(prob), idx.size, C.byref(t)))
        return idx, prob

    def describe(self, kind):
        buf = C.create_string_buffer(1 << 20)
        self._check(self._L.b200ocr_describe(self._h, kind, buf, len(buf)))
        return buf.value.decode()

    def plan_stats(self, kind, n, h, w):
        fl, ar = C.c_double(), C.c_double()
        la = C.c_int()
        self._check(self._L.b200ocr_plan_stats(self._h, kind, n, h, w, C.byref(fl), C.byref(la), C.byref(ar)))
        return {"flops": fl.value, "launches": la.value, "arena_bytes": ar.value}

    def profile_plan(self, kind, n, h, w, iters=3):
        """-> list of dict(name, ms, flops, bytes) per kernel launch of the plan for input [n,3,h,w]."""
        cap = 4096
        ms = (C.c_float * cap)()
        fl = (C.c_double * cap)()
        by = (C.c_double * cap)()
        names = C.create_string_buffer(1 << 18)
        ns = self._check(self._L.b200ocr_profile_plan(self._h, kind, n, h, w, iters, ms, fl, by, cap, names, len(names)))
        parts = names.raw.split(b"\0")
        return [{"name": parts[k].decode(), "ms": ms[k], "flops": fl[k], "bytes": by[k]} for k in range(ns)]

    def time_plan(self, kind, n, h, w, head=False, iters=10):
        ms = C.c_float()
        self._check(self._L.b200ocr_time_plan(self._h, kind, n, h, w, int(head), iters, C.byref(ms)))
        return ms.value

    def l2_flush(self):
        self._check(self._L.b200ocr_l2_flush(self._h))

    def pinned_array(self, shape, dtype=np.uint8):
        """numpy array over page-locked host memory (b200ocr_host_alloc).  The buffer belongs to the array: it is
        released (b200ocr_host_free) when the last view of it is garbage collected, independently of the engine, so a
        long-running page feed does not accumulate page-locked memory and no view dangles after Engine.close()."""
        nbytes = max(1, int(np.prod(shape)) * np.dtype(dtype).itemsize)
        ptr = self._L.b200ocr_host_alloc(nbytes)
        if not ptr:
            raise MemoryError("b200ocr_host_alloc(%d) failed" % nbytes)
        buf = (C.c_uint8 * nbytes).from_address(ptr)
        # every numpy view keeps `buf` alive through its .base chain; the finalizer runs once `buf` is unreachable
        weakref.finalize(buf, self._L.b200ocr_host_free, ptr)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def launch_count(self):
        return int(self._L.b200ocr_launch_count(self._h))

    def stage_times(self):
        ms = (C.c_float * 4)()
        self._check(self._L.b200ocr_stage_times(self._h, ms))
        return list(ms)

    # ---- Level B -----------------------------------------------------------------
    def ocr_stage_times(self):
        ms = (C.c_float * 6)()
        self._check(self._L.b200ocr_ocr_stage_times(self._h, ms))
        return dict(zip(["upload", "det", "host_plan", "warp_cls", "rec", "drain"], [float(v) for v in ms]))

    def det_preprocess(self, img, params=None):
        """-> (chw float32 [1,3,H,W], shape float64[4]) like transform(data, ops) in predict_det.py:98."""
        params = params or det_params()
        img = np.ascontiguousarray(img, dtype=np.uint8)
        h, w = img.shape[:2]
        oh, ow = C.c_int(), C.c_int()
        shape = (C.c_double * 4)()
        self._check(self._L.b200ocr_det_preprocess(self._h, _ptr(img), h, w, img.strides[0], C.byref(params), None, 0,
                                                   C.byref(oh), C.byref(ow), shape))
        out = np.empty((1, 3, oh.value, ow.value), np.float32)
        self._check(self._L.b200ocr_det_preprocess(self._h, _ptr(img), h, w, img.strides[0], C.byref(params), _ptr(out),
                                                   out.size, C.byref(oh), C.byref(ow), shape))
        return out, np.array(list(shape))

    def db_postprocess(self, prob, src_h, src_w, params=None):
        """-> (boxes int32 [n,4,2], scores float64 [n]) like DBPostProcess.boxes_from_bitmap."""
        params = params or det_params()
        prob = np.ascontiguousarray(prob, dtype=np.float32)
        h, w = prob.shape
        cap = max(1, params.max_candidates)
        boxes = np.zeros((cap, 4, 2), np.int32)
        scores = np.zeros((cap,), np.float64)
        n = C.c_int()
        self._check(self._L.b200ocr_db_postprocess(self._h, _ptr(prob), h, w, int(src_h), int(src_w), C.byref(params),
                                                   _ptr(boxes), _ptr(scores), cap, C.byref(n)))
        return boxes[:n.value].copy(), scores[:n.value].copy()

    def det_pages(self, pages, params=None, max_boxes=1000, want_prob=False):
        """-> list of float32 [n_i,4,2] (TextDetector.__call__ per page) [, list of prob maps]."""
        params = params or det_params()
        pages = [np.ascontiguousarray(p, dtype=np.uint8) for p in pages]
        n = len(pages)
        if n == 0:
            return ([], []) if want_prob else []
        ptrs = (C.c_void_p * n)(*[p.ctypes.data for p in pages])
        hs, ws = _int_arr([p.shape[0] for p in pages]), _int_arr([p.shape[1] for p in pages])
        st = _int_arr([p.strides[0] for p in pages])
        boxes = np.zeros((n, max_boxes, 4, 2), np.float32)
        counts = (C.c_int * n)()
        dims = (C.c_int * (2 * n))()
        prob = None
        cap = 0
        if want_prob:
            cap = sum(_det_hw(p.shape[0], p.shape[1], params) for p in pages)
            prob = np.empty((cap,), np.float32)
        self._check(self._L.b200ocr_det_pages(self._h, ptrs, hs, ws, st, n, C.byref(params), _ptr(boxes), max_boxes,
                                              counts, _ptr(prob) if want_prob else None, cap, dims))
        out = [boxes[i, :counts[i]].copy() for i in range(n)]
        if not want_prob:
            return out
        maps, off = [], 0
        for i in range(n):
            h, w = dims[2 * i], dims[2 * i + 1]
            maps.append(prob[off:off + h * w].reshape(h, w).copy())
            off += h * w
        return out, maps

    def det_submit(self, pages, params=None):
        """Enqueues TextDetector.__call__ for a batch of pages without waiting for the GPU -> ticket (b200ocr_det_submit).
        Up to two batches may be in flight; the page arrays are kept alive by the ticket until det_collect."""
        params = params or det_params()
        pages = [np.ascontiguousarray(p, dtype=np.uint8) for p in pages]
        n = len(pages)
        ptrs = (C.c_void_p * n)(*[p.ctypes.data for p in pages])
        t = C.c_int()
        self._check(self._L.b200ocr_det_submit(self._h, ptrs, _int_arr([p.shape[0] for p in pages]),
                                               _int_arr([p.shape[1] for p in pages]),
                                               _int_arr([p.strides[0] for p in pages]), n, C.byref(params), C.byref(t)))
        return (t.value, pages)

    def det_collect(self, ticket, max_boxes=1000):
        """Blocks until the batch of `ticket` is done -> list of float32 [n_i,4,2], as det_pages returns."""
        tid, pages = ticket
        n = len(pages)
        boxes = np.zeros((n, max_boxes, 4, 2), np.float32)
        counts = (C.c_int * n)()
        self._check(self._L.b200ocr_det_collect(self._h, tid, _ptr(boxes), max_boxes, counts))
        return [boxes[i, :counts[i]].copy() for i in range(n)]

    def det_pages_stream(self, batches, params=None, max_boxes=1000, depth=None):
        """Generator over an iterable of page batches: yields each batch's boxes in order while keeping `depth` (<= 4)
        batches in flight, so uploads and the network of the next batch overlap the post-processing of this one."""
        if depth is None:
            depth = int(os.environ.get("B200OCR_DET_DEPTH", "4"))
        pending = []
        for b in batches:
            if len(pending) >= max(1, min(depth, 4)):
                yield self.det_collect(pending.pop(0), max_boxes)
            pending.append(self.det_submit(b, params))
        while pending:
            yield self.det_collect(pending.pop(0), max_boxes)

    def crop_boxes(self, img, boxes):
        """-> list of HWC uint8 crops (get_rotate_crop_image per box, utils.py:12-53)."""
        img = np.ascontiguousarray(img, dtype=np.uint8)
        boxes = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 4, 2)
        n = boxes.shape[0]
        if n == 0:
            return []
        dims = (C.c_int * (2 * n))()
        h, w = img.shape[:2]
        self._check(self._L.b200ocr_crop_boxes(self._h, _ptr(img), h, w, img.strides[0], _ptr(boxes), n, None, 0, dims))
        total = sum(dims[2 * i] * dims[2 * i + 1] * 3 for i in range(n))
        pool = np.empty((max(total, 1),), np.uint8)
        self._check(self._L.b200ocr_crop_boxes(self._h, _ptr(img), h, w, img.strides[0], _ptr(boxes), n, _ptr(pool),
                                               pool.size, dims))
        out, off = [], 0
        for i in range(n):
            ch, cw = dims[2 * i], dims[2 * i + 1]
            out.append(pool[off:off + ch * cw * 3].reshape(ch, cw, 3).copy())
            off += ch * cw * 3
        return out

    def _crop_args(self, crops):
        crops = [np.ascontiguousarray(c, dtype=np.uint8) for c in crops]
        n = len(crops)
        ptrs = (C.c_void_p * max(n, 1))(*[c.ctypes.data for c in crops])
        return crops, n, ptrs, _int_arr([c.shape[0] for c in crops]), _int_arr([c.shape[1] for c in crops])

    def cls_crops(self, crops, batch_num=6):
        crops, n, ptrs, hs, ws = self._crop_args(crops)
        labels = np.zeros((max(n, 1),), np.int32)
        scores = np.zeros((max(n, 1),), np.float32)
        if n:
            self._check(self._L.b200ocr_cls_crops(self._h, ptrs, hs, ws, n, batch_num, _ptr(labels), _ptr(scores)))
        return labels[:n], scores[:n]

    def rec_crops(self, crops, batch_num=6, rotate180=None, max_len=None):
        crops, n, ptrs, hs, ws = self._crop_args(crops)
        if max_len is None:
            max_len = 8
            for c in crops:
                max_len = max(max_len, int(48 * max(320 / 48.0, c.shape[1] / float(c.shape[0]))) // 8 + 8)
        ids = np.zeros((max(n, 1), max_len), np.int32)
        lens = (C.c_int * max(n, 1))()
        scores = np.zeros((max(n, 1),), np.float32)
        rot = None
        if rotate180 is not None:
            rot = np.ascontiguousarray(rotate180, dtype=np.uint8)
        if n:
            self._check(self._L.b200ocr_rec_crops(self._h, ptrs, hs, ws, n, batch_num, _ptr(rot) if rot is not None else None,
                                                  _ptr(ids), max_len, lens, _ptr(scores)))
        return [ids[i, :lens[i]].copy() for i in range(n)], scores[:n]

    def ocr_pages(self, pages, params=None, line_cap=None, max_len=None):
        """TextSystem.__call__ for a batch of pages in one call (Level C).  -> per page a dict with
        boxes float32 [m,4,2] (sorted_boxes order), ids (list of int32 arrays), scores float32 [m],
        cls_labels int32 [m], cls_scores float32 [m] (zeros when the classifier is off)."""
        params = params or ocr_params()
        pages = [np.ascontiguousarray(p, dtype=np.uint8) for p in pages]
        n = len(pages)
        if n == 0:
            return []
        ptrs = (C.c_void_p * n)(*[p.ctypes.data for p in pages])
        hs, ws = _int_arr([p.shape[0] for p in pages]), _int_arr([p.shape[1] for p in pages])
        st = _int_arr([p.strides[0] for p in pages])
        cap = int(line_cap or max(1000, 300 * n))
        # id buffer width: sticky across calls — a line longer than the buffer costs one re-run, after which the engine
        # remembers the width (very wide lines decode to more than 160 characters)
        sticky = max_len is None
        if sticky:
            max_len = getattr(self, "_ocr_max_len", 160)
        while True:
            # only the first sum(counts) rows are written by the library and read below: no need to clear 12 MB per call
            boxes = np.empty((cap, 4, 2), np.float32)
            ids = np.empty((cap, max_len), np.int32)
            lens = np.zeros((cap,), np.int32)
            scores = np.empty((cap,), np.float32)
            cl = np.zeros((cap,), np.int32)
            cs = np.zeros((cap,), np.float32)
            counts = (C.c_int * n)()
            rc = self._L.b200ocr_ocr_pages(self._h, ptrs, hs, ws, st, n, C.byref(params), _ptr(boxes), cap, counts, _ptr(ids),
                                           max_len, _ptr(lens), _ptr(scores), _ptr(cl), _ptr(cs))
            if rc == -4 and line_cap is None and cap < 1000 * n:   # B200OCR_E_CAPACITY: more lines than assumed
                cap = 1000 * n
                continue
            self._check(rc)
            total = sum(counts[i] for i in range(n))
            longest = int(lens[:total].max()) if total else 0
            if longest > max_len:                       # a line longer than the id buffer: rerun with room for it
                max_len = (longest + 8 + 31) // 32 * 32
                if sticky:
                    self._ocr_max_len = max_len
                continue
            break
        out, off = [], 0
        for i in range(n):
            m = counts[i]
            sl = slice(off, off + m)
            out.append({"boxes": boxes[sl].copy(), "ids": [ids[off + k, :lens[off + k]].copy() for k in range(m)],
                        "scores": scores[sl].copy(), "cls_labels": cl[sl].copy(), "cls_scores": cs[sl].copy()})
            off += m
        return out

    def det_resident_upload(self, pages):
        pages = [np.ascontiguousarray(p, dtype=np.uint8) for p in pages]
        n = len(pages)
        ptrs = (C.c_void_p * n)(*[p.ctypes.data for p in pages])
        self._check(self._L.b200ocr_det_resident_upload(self._h, ptrs, _int_arr([p.shape[0] for p in pages]),
                                                        _int_arr([p.shape[1] for p in pages]),
                                                        _int_arr([p.strides[0] for p in pages]), n))

    def det_resident_run(self, params=None, iters=1):
        params = params or det_params()
        ms, net = C.c_float(), C.c_float()
        nb = C.c_int()
        self._check(self._L.b200ocr_det_resident_run(self._h, C.byref(params), iters, C.byref(ms), C.byref(nb), C.byref(net)))
        return ms.value, nb.value, net.value

    def det_resident_stream(self, params=None, iters=1, pages_per_step=16):
        """-> (device ms of all `iters` steps, boxes found): the resident benchmark through det_submit / det_collect."""
        params = params or det_params()
        ms = C.c_float()
        nb = C.c_int()
        self._check(self._L.b200ocr_det_resident_stream(self._h, C.byref(params), iters, pages_per_step, C.byref(ms), C.byref(nb)))
        return ms.value, nb.value


def _det_hw(h, w, params):
    """H*W of the det input for an h x w page (DetResizeForTest.resize_image_type0, operators.py:91-138)."""
    ph, pw = (max(32, h), max(32, w)) if h + w < 64 else (h, w)
    lim = float(params.limit_side_len)
    ratio = 1.0
    if params.limit_type == 0:
        if max(ph, pw) > lim:
            ratio = lim / ph if ph > pw else lim / pw
    elif params.limit_type == 1:
        if min(ph, pw) < lim:
            ratio = lim / ph if ph < pw else lim / pw
    else:
        ratio = lim / max(ph, pw)
    rh, rw = int(ph * ratio), int(pw * ratio)
    rh = max(int(round(rh / 32) * 32), 32)
    rw = max(int(round(rw / 32) * 32), 32)
    return rh * rw

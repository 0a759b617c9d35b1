"""Cross-request micro-batching and page feed.

The reference serves one image per `model.ocr(img)` call: `EngineManager.run_ocr` gates concurrent requests with a
semaphore and runs `_sync_ocr` in the default executor (app/engine.py:104-147), `ocr_images_pdfs.py:154` does the same
from a ThreadPoolExecutor(4).  On a B200 one page per call leaves most of the device idle (≈ 120–160 pages/s against
≈ 450 when 16 pages share a call), so `MicroBatcher` sits where those callers call `model.ocr`: it has the same
`ocr(img)` signature, blocks the calling thread until its page is done, and lets a single worker thread hand whatever
arrived within `max_wait_ms` (up to `max_batch` pages) to `ONNXPaddleOcr.ocr_batch` — one `b200ocr_ocr_pages` call.
Results are identical to per-page calls (each crop keeps the padded width its own page's batch plan gives it).

`decode_pages` is the feed side: encoded images (paths or bytes) are decoded on a thread pool straight into pinned host
buffers (`b200ocr_host_alloc`), the form in which the engine's uploads are asynchronous.
"""
import queue
import threading
from concurrent.futures import Future, ThreadPoolExecutor

import numpy as np


class MicroBatcher(object):
    def __init__(self, model, max_batch=16, max_wait_ms=2.0):
        self.model = model
        self.max_batch = int(max_batch)
        self.max_wait = float(max_wait_ms) / 1e3
        self._q = queue.Queue()
        self._stop = False
        self.batches = 0
        self.pages = 0
        self._worker = threading.Thread(target=self._run, name="b200ocr-microbatcher", daemon=True)
        self._worker.start()

    @staticmethod
    def _validate(img):
        """Rejects an input in the caller's own thread, before it can be merged with other requests' pages: the engine
        takes H x W x 3 uint8 BGR arrays (what cv2.imread / cv2.imdecode yield; None = an undecodable image)."""
        if img is None:
            raise ValueError("ocr(): image is None (undecodable input?)")
        if not isinstance(img, np.ndarray):
            raise TypeError("ocr(): expected a numpy array, got %s" % type(img).__name__)
        if img.ndim != 3 or img.shape[2] != 3 or img.dtype != np.uint8:
            raise ValueError("ocr(): expected an H x W x 3 uint8 BGR image, got shape %s dtype %s" % (img.shape, img.dtype))
        if img.shape[0] == 0 or img.shape[1] == 0:
            raise ValueError("ocr(): empty image")

    # ---- the reference-facing call (onnx_paddleocr.py:28): blocks until this page's result is ready
    def ocr(self, img, det=True, rec=True, cls=True):
        if not (det and rec):
            return self.model.ocr(img, det=det, rec=rec, cls=cls)      # partial modes are not batched
        return self.submit(img, cls=cls).result()

    def submit(self, img, cls=True):
        """Non-blocking variant: a concurrent.futures.Future resolving to `model.ocr(img)`'s result."""
        self._validate(img)
        fut = Future()
        self._q.put((img, cls, fut))
        return fut

    def close(self):
        self._stop = True
        self._q.put(None)
        self._worker.join(timeout=10)

    def _collect(self):
        first = self._q.get()
        if first is None:
            return None
        items = [first]
        try:
            while len(items) < self.max_batch:
                # the first page already waited for its predecessors' batch; later arrivals get a short grace period
                nxt = self._q.get(timeout=self.max_wait)
                if nxt is None:
                    self._stop = True
                    break
                items.append(nxt)
        except queue.Empty:
            pass
        return items

    def _run(self):
        while not self._stop:
            items = self._collect()
            if not items:
                break
            for cls_flag in (True, False):
                group = [it for it in items if bool(it[1]) == cls_flag]
                if not group:
                    continue
                try:
                    results = self.model.ocr_batch([g[0] for g in group], cls=cls_flag)
                    for (_, _, fut), res in zip(group, results):
                        fut.set_result([res])                          # ocr() returns one list per image
                except Exception:
                    # One bad page must fail only its own request (the reference's per-image model.ocr does): retry
                    # every page of the group on its own, so that only the offending page's future gets the exception.
                    for img, _, fut in group:
                        if fut.done():
                            continue
                        try:
                            fut.set_result([self.model.ocr_batch([img], cls=cls_flag)[0]])
                        except Exception as exc:
                            fut.set_exception(exc)
            self.batches += 1
            self.pages += len(items)


def decode_pages(sources, engine=None, workers=8, flags=None):
    """Decodes encoded images (file paths or bytes-like objects) in parallel.  With `engine` (an `_lib.Engine`) the pixels
    land in pinned host buffers, from which `ocr_batch` / `det_pages` upload asynchronously.  Returns HWC uint8 BGR arrays
    in input order (None where decoding failed, like `cv2.imread`)."""
    import cv2
    flags = cv2.IMREAD_COLOR if flags is None else flags

    def one(src):
        if isinstance(src, (bytes, bytearray, memoryview)):
            buf = np.frombuffer(src, dtype=np.uint8)
        else:
            buf = np.fromfile(src, dtype=np.uint8)        # also handles non-ASCII paths (test_images/QQ群.jpg)
        img = cv2.imdecode(buf, flags)
        if img is None or engine is None:
            return img
        pinned = engine.pinned_array(img.shape, np.uint8)   # page-locked; released when the array is garbage collected
        pinned[...] = img
        return pinned

    with ThreadPoolExecutor(max_workers=max(1, int(workers))) as ex:
        return list(ex.map(one, sources))

"""GPU: the streaming det entry points (b200ocr_det_submit / b200ocr_det_collect, up to four batches in flight on separate
lane pairs) return exactly what the synchronous b200ocr_det_pages returns, interleave correctly with synchronous calls,
and reject misuse (a fifth batch in flight, a stale ticket) with an error instead of corrupting a slot."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _batches():
    from onnxocr_b200 import synth_pages
    shapes = [(480, 640), (512, 704), (320, 480)]
    out = []
    for k in range(5):                       # 5 batches: every slot is reused, the last one drains alone
        n = [16, 3, 9, 12, 1][k]             # 16 -> two sub-batches of 8 (both lanes), the others one lane / mixed groups
        pages = []
        for i in range(n):
            h, w = shapes[(k + (i if k == 2 else 0)) % 3]
            pages.append(synth_pages.render_page(h, w, 6, seed=1000 + 37 * k + i)[0])
        out.append(pages)
    return out


def _same(a, b):
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert x.shape == y.shape and np.array_equal(x, y)


@pytest.mark.parametrize("which", ["engine_bf16", "engine_fp32"])
def test_stream_equals_sync(which, request):
    eng = request.getfixturevalue(which)
    batches = _batches()
    ref = [eng.det_pages(b) for b in batches]
    assert sum(len(x) for r in ref for x in r) > 100
    for rep in range(2):                     # second pass replays the CUDA graphs of the slot plans
        got = list(eng.det_pages_stream(batches))
        assert len(got) == len(ref)
        for g, r in zip(got, ref):
            _same(g, r)


def test_stream_interleaved_with_sync_calls(engine_bf16):
    eng = engine_bf16
    batches = _batches()
    ref = [eng.det_pages(b) for b in batches]
    t0 = eng.det_submit(batches[0])
    t1 = eng.det_submit(batches[1])
    mid = eng.det_pages(batches[2])          # a synchronous call drains the batches in flight; their results are kept
    _same(mid, ref[2])
    _same(eng.det_collect(t1), ref[1])       # collecting out of order is allowed
    _same(eng.det_collect(t0), ref[0])
    # the fused Level-C entry uses the same lanes
    t3 = eng.det_submit(batches[3])
    lines = eng.ocr_pages(batches[4])
    assert len(lines) == len(batches[4])
    _same(eng.det_collect(t3), ref[3])


def test_stream_misuse_is_an_error(engine_bf16):
    eng = engine_bf16
    batches = _batches()
    t0 = eng.det_submit(batches[1])
    t1 = eng.det_submit(batches[4])
    t2 = eng.det_submit(batches[4])
    t3 = eng.det_submit(batches[1])
    with pytest.raises(RuntimeError, match="in flight"):
        eng.det_submit(batches[1])            # four slots
    _same(eng.det_collect(t3), eng.det_collect(t0))
    t0 = eng.det_submit(batches[1])
    _same(eng.det_collect(t2), eng.det_collect(t1))
    t1 = eng.det_submit(batches[4])
    a = eng.det_collect(t0)
    with pytest.raises(RuntimeError, match="ticket"):
        eng.det_collect(t0)                  # already collected
    with pytest.raises(RuntimeError, match="ticket"):
        eng.det_collect((12345, t1[1]))
    with pytest.raises(RuntimeError, match="capacity"):
        eng.det_collect(t1, max_boxes=1)     # too small a box buffer: the ticket stays collectable
    b = eng.det_collect(t1)
    _same(a, eng.det_pages(batches[1]))
    _same(b, eng.det_pages(batches[4]))


def test_resident_stream_counts_the_same_boxes(engine_bf16):
    from onnxocr_b200 import synth_pages
    eng = engine_bf16
    pages = synth_pages.det_pages_c2(8, 480, 12, seed=5)
    eng.det_resident_upload(pages)
    _, n_sync, _ = eng.det_resident_run(iters=1)
    ms, n_stream = eng.det_resident_stream(iters=6, pages_per_step=4)     # 6 steps of 4 pages = the 8 pages three times
    assert n_sync > 20 and n_stream == 3 * n_sync and ms > 0


def test_failed_submit_leaves_the_slot_clean(engine_bf16):
    """A batch that fails while it is being enqueued (here: a null page in its second shape group, after the first group is
    already on a lane) must not leave a pending sub-batch behind: the next batches on the same slot return correct boxes."""
    import ctypes as C
    from onnxocr_b200 import _lib, synth_pages
    eng = engine_bf16
    a = [synth_pages.render_page(320, 480, 6, seed=2000 + i)[0] for i in range(3)]
    b = [synth_pages.render_page(480, 640, 6, seed=2100 + i)[0] for i in range(2)]
    pages = a + b
    ref = eng.det_pages(pages)
    n = len(pages)
    ptrs = (C.c_void_p * n)(*[p.ctypes.data for p in pages])
    ptrs[n - 1] = None                                   # a page of the second (larger) shape group
    hs, ws = _lib._int_arr([p.shape[0] for p in pages]), _lib._int_arr([p.shape[1] for p in pages])
    st = _lib._int_arr([p.strides[0] for p in pages])
    params = _lib.det_params()
    t = C.c_int()
    for _ in range(3):                                   # repeatedly: every slot sees a failed batch
        rc = eng._L.b200ocr_det_submit(eng._h, ptrs, hs, ws, st, n, C.byref(params), C.byref(t))
        assert rc != 0 and b"null page" in eng._L.b200ocr_last_error(eng._h)
    boxes = np.zeros((n, 1000, 4, 2), np.float32)
    counts = (C.c_int * n)()
    rc = eng._L.b200ocr_det_pages(eng._h, ptrs, hs, ws, st, n, C.byref(params), _lib._ptr(boxes), 1000, counts, None, 0,
                                  (C.c_int * (2 * n))())
    assert rc != 0                                       # the synchronous entry point fails the same way ...
    for _ in range(2):                                   # ... and both paths work afterwards
        for got in eng.det_pages_stream([pages, pages, pages]):
            _same(got, ref)
        _same(eng.det_pages(pages), ref)

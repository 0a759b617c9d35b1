"""Secondary measurements (not the headline bench): BASELINE.json configs C3 (rec-only) and C4 (end to end).

    python tools/bench_configs.py [bf16|fp32]

Prints one JSON line per config with GPU throughput through the public API (host buffers in, results out) and the
oracle-port CPU throughput on a bounded sample."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from onnxocr_b200 import pipeline, synth_graphs, synth_pages  # noqa: E402


def main():
    prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
    paths = synth_graphs.ensure_models()
    ocr = pipeline.ONNXPaddleOcr(use_angle_cls=True, precision=prec)
    from oracle.restate_pipeline import OraclePipeline
    mp = dict(paths, cls=os.path.join(synth_graphs.ASSET_DIR, "ppocrv5_cls.onnx"), dict=os.path.join(synth_graphs.ASSET_DIR, "ppocrv5_dict.txt"))
    oracle = OraclePipeline(mp, use_angle_cls=True)
    # ---- C3: rec-only, 256 crops 48x320
    crops = synth_pages.rec_crops_c3(256, seed=1)
    for _ in range(3):
        ocr.text_recognizer(crops)
    t0 = time.perf_counter()
    reps = 10
    for _ in range(reps):
        res = ocr.text_recognizer(crops)
    dt = (time.perf_counter() - t0) / reps
    t1 = time.perf_counter()
    ref = oracle.recognize(crops[:24])
    cpu = 24 / (time.perf_counter() - t1)
    same = sum(1 for a, b in zip(res[:24], ref) if a[0] == b[0])
    print(json.dumps({"config": "C3 rec-only 256 x 3x48x320, CTC greedy decode", "precision": prec, "lines_per_s": 256 / dt,
                      "ms_per_batch": dt * 1e3, "cpu_port_lines_per_s": cpu, "identical_lines_vs_oracle": "%d/24" % same}))
    # ---- C4: end to end, mixed-size pages
    pages = synth_pages.e2e_pages_c4(16, seed=2)
    for p in pages[:3]:
        ocr.ocr(p)
    t0 = time.perf_counter()
    n_lines = 0
    for p in pages:
        n_lines += len(ocr.ocr(p)[0])
    dt = time.perf_counter() - t0
    t1 = time.perf_counter()
    ref_lines = len(oracle.ocr(pages[0])[0])
    cpu_dt = time.perf_counter() - t1
    print(json.dumps({"config": "C4 end-to-end det+cls+rec, 16 mixed-size pages, one page per call", "precision": prec,
                      "pages_per_s": len(pages) / dt, "lines_per_s": n_lines / dt, "lines": n_lines,
                      "cpu_port_pages_per_s": 1.0 / cpu_dt, "cpu_port_lines_first_page": ref_lines}))


if __name__ == "__main__":
    main()

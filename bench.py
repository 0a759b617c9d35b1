#!/usr/bin/env python
"""Headline benchmark: PP-OCRv5 mobile det pages/s on B200 (BASELINE.json config 2).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--precision bf16|fp32]

One "step" = one pass of the det hot path (fused resize/normalise -> DBNet -> DBPostProcess -> boxes) over a
batch of 16 synthetic 960x960 pages per GPU.  Pages shard across GPUs as independent per-GPU pipelines
(weak scaling, no collective on the data path); torch.distributed is used for the barrier and the max-over-ranks
of the device time only.

  value     pages/s with the page bytes already resident in HBM (CUDA events on the engine stream), through the
            streaming entry points (b200ocr_det_submit / _collect: four batches in flight, so the DBPostProcess tail
            and box read-back of step k overlap the network of step k+1).  Steps rotate through 4 resident page sets
            (177 MB > the 126 MB L2).
  e2e       pages/s through the public API (`Engine.det_pages_stream`) with HOST page buffers: H2D of the page
            bytes from pinned memory and D2H of the boxes are inside the timed region
  sync      the same two numbers with one blocking `det_pages` call per step (nothing overlaps across steps; L2
            flushed between the device-timed steps)
  roofline  dominant kernel class of the step, timed live per launch with CUDA events
  cpu_baseline / --impl reference: the oracle port of the reference's CPU path (reference pre/post-processing
            restated with cv2/numpy, network through the torch-CPU fp32 interpreter: onnxruntime is not
            installed in this image) timed on the host cores of the same box.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PAGES_PER_GPU = 16
SIDE = 960
LINES = 40
SETS = 4            # page sets a rank rotates through: 4 x 16 x 2.76 MB = 177 MB of inputs > 126 MB L2
METRIC = "PP-OCRv5 mobile det pages/sec (960x960 pages, batch 16, prob map + DBPostProcess boxes)"
# the same dict is printed by both arms (the reference arm runs a bounded sample of this workload on the host cores)
CONFIG = {"workload": "C2 det-only: 16 synthetic 960x960 pages per GPU per step (synthetic-weight PP-OCRv5 mobile det graph, "
                      "fitted on rendered pages), prob map + DBPostProcess boxes",
          "pages_per_gpu": PAGES_PER_GPU,
          "l2": "inputs larger than L2: the steps rotate through %d page sets (%d pages, 177 MB > 126 MB L2) and every step "
                "rewrites ~2 GB of activations" % (SETS, SETS * PAGES_PER_GPU),
          "pipeline": "streaming det: four batches in flight (submit / collect), each a 16-page plan on its own det lane; "
                      "uploads, network and DBPostProcess of neighbouring steps overlap"}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        super().__init__(daemon=True)
        self.device = device
        self.samples = []
        self.reasons = set()
        self.stop_flag = False
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                parts = [x.strip() for x in line.split(",")]
                if len(parts) < 6:
                    continue
                try:
                    self.samples.append((float(parts[0]), float(parts[1])))
                except ValueError:
                    continue
                for nm, v in zip(names, parts[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
        except Exception:
            pass

    def stop(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": sorted(self.reasons)}
        return {"sm_mhz": statistics.median(s[0] for s in self.samples), "sm_max_mhz": self.samples[0][1],
                "reasons": sorted(self.reasons)}


def make_pages(rank, sets=SETS):
    from onnxocr_b200 import synth_pages
    return synth_pages.det_pages_c2(PAGES_PER_GPU * sets, SIDE, LINES, seed=rank)


def oracle_det_pages(pages, graph):
    """The reference's det path on the CPU: DetResizeForTest/NormalizeImage/ToCHWImage -> session.run -> DBPostProcess
    -> filter_tag_det_res, one page at a time like TextDetector.__call__ (predict_det.py:94-120)."""
    from oracle import restate_post, restate_pre
    n_boxes = 0
    for img in pages:
        x, shape = restate_pre.det_preprocess(img)
        prob = graph.run(None, {"x": x})[0]
        boxes, _ = restate_post.db_postprocess(prob, shape[None])[0]
        n_boxes += len(restate_post.filter_tag_det_res(boxes, img.shape))
    return n_boxes


def cpu_arm(pages, steps, warmup, sample_pages):
    import torch
    from onnxocr_b200 import synth_graphs
    from oracle.torch_interp import TorchGraph
    torch.set_num_threads(os.cpu_count() or 1)
    graph = TorchGraph(synth_graphs.ensure_models()["det"])
    sample = pages[:sample_pages]
    for _ in range(warmup):
        oracle_det_pages(sample[:1], graph)
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle_det_pages(sample, graph)
    dt = time.perf_counter() - t0
    return {"value": len(sample) * steps / dt, "ms_per_step": dt / steps * 1e3, "cores": torch.get_num_threads(),
            "sample": "%d of the %d 960x960 pages per step, %d step(s); torch-CPU restatement of the ORT-CPU path "
                      "(onnxruntime absent), cv2 %s pre/post-processing" % (len(sample), PAGES_PER_GPU, steps,
                                                                           __import__("cv2").__version__)}


def kernel_class(name):
    return name.split(" ")[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    warmup = max(args.warmup, 0)

    if args.impl == "reference":
        if rank != 0:
            return
        pages = make_pages(0, sets=1)
        r = cpu_arm(pages, max(1, args.steps), min(warmup, 1), sample_pages=4)
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "pages/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": CONFIG,
            "cpu_baseline": {"value": r["value"], "unit": "pages/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": "pages/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist
    from onnxocr_b200 import _lib, build, synth_graphs
    build.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libb200ocr has no CPU fallback")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    prec = _lib.BF16 if args.precision == "bf16" else _lib.FP32
    paths = synth_graphs.ensure_models()
    eng = _lib.Engine(local_rank)
    eng.load_onnx(_lib.DET, paths["det"], prec)
    pages = make_pages(rank)
    params = _lib.det_params()

    # ---- device-resident timing ("value"): K steps through the streaming path, one device-timed region
    eng.det_resident_upload(pages)
    # untimed warm-up: every streaming slot launches its plan twice (first launch eager, second captured as a CUDA graph)
    warm_steps = max(warmup, 3, 2 * 4)
    eng.det_resident_stream(params, warm_steps, PAGES_PER_GPU)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count()
    barrier()
    dev_ms, n_boxes = eng.det_resident_stream(params, args.steps, PAGES_PER_GPU)
    barrier()
    launches = eng.launch_count() - launches0
    dev_ms = max_over_ranks(dev_ms)
    value = PAGES_PER_GPU * world * args.steps / (dev_ms / 1e3)

    # ---- end to end through the public API with host buffers ("e2e")
    pinned = []
    for p in pages:
        a = eng.pinned_array(p.shape, np.uint8)
        a[...] = p
        pinned.append(a)
    sets = [pinned[k * PAGES_PER_GPU:(k + 1) * PAGES_PER_GPU] for k in range(SETS)]
    for _ in eng.det_pages_stream([sets[k % SETS] for k in range(warm_steps)], params):
        pass
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    for out in eng.det_pages_stream((sets[k % SETS] for k in range(args.steps)), params):
        d2h = sum(b.nbytes for b in out) + 4 * len(out)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e = PAGES_PER_GPU * world * args.steps / e2e_s
    h2d = sum(p.nbytes for p in sets[0])

    # ---- one blocking call per step (no overlap across steps), L2 flushed between the device-timed steps
    sync_steps = min(args.steps, 10)
    eng.det_resident_upload(pages[:PAGES_PER_GPU])
    for _ in range(3):
        eng.det_resident_run(params, 1)
    barrier()
    sync_ms = 0.0
    for _ in range(sync_steps):
        eng.l2_flush()
        sync_ms += eng.det_resident_run(params, 1)[0]
    sync_ms = max_over_ranks(sync_ms)
    for _ in range(3):
        eng.det_pages(sets[0], params)
    barrier()
    t0 = time.perf_counter()
    for k in range(sync_steps):
        eng.det_pages(sets[k % SETS], params)
    barrier()
    sync_e2e_s = max_over_ranks(time.perf_counter() - t0)
    sync = {"value": PAGES_PER_GPU * world * sync_steps / (sync_ms / 1e3), "ms_per_step": sync_ms / sync_steps,
            "e2e": PAGES_PER_GPU * world * sync_steps / sync_e2e_s, "steps": sync_steps, "unit": "pages/s"}
    if rank == 0:
        sampler.stop()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- the network alone: the 16-page plan of a step, replayed back to back as a CUDA graph
    net_ms = eng.time_plan(_lib.DET, PAGES_PER_GPU, SIDE, SIDE, iters=10)

    # ---- roofline of the dominant kernel (live CUDA-event timing per launch of the plan a step actually runs:
    #      one 16-page plan per streaming batch)
    peaks = load_peaks()
    sub = PAGES_PER_GPU
    prof = eng.profile_plan(_lib.DET, sub, SIDE, SIDE, iters=5)
    classes = {}
    for s in prof:
        c = classes.setdefault(kernel_class(s["name"]), {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
        c["ms"] += s["ms"]
        c["flops"] += s["flops"]
        c["bytes"] += s["bytes"]
        c["launches"] += 1
    total_ms = sum(c["ms"] for c in classes.values())
    top = max(prof, key=lambda s: s["ms"])
    # every layer of this network is far below the tensor ridge (<= 120 FLOP/B vs 1381e12 / 6543e9 = 211): HBM-bound
    ach = top["bytes"] / (top["ms"] / 1e3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r2_traffic.json")       # ncu --set full captures of the same 16-page launches
    if os.path.exists(tpath):
        with open(tpath) as fh:
            tj = json.load(fh)
        key = {"tcconv conv2d.out_2.tmp": "ws_out2_packed", "tcconvT2 p2o.ConvTranspose_0.tmp": "ws_convT2_tail",
               "eltwise p2o.Add_82.tmp": "fpn_add82", "pre+stem conv2d.out_0.tmp": "pre_stem", "tcconv conv2d.out_46.tmp": "ws_out46_sums",
               "tcconv conv2d.out_58.tmp": "ws_halo3x3_out58", "concat p2o.Concat_0.tmp": "concat",
               "dwconv conv2d.out_1.tmp": "dw3x3_first"}.get(top["name"])
        if key in tj:
            traffic = (tj[key]["dram_read_mb"] + tj[key]["dram_write_mb"]) * 1e6
    kern_of = {"tcconv": "tc::conv_tc_ws_kernel (tcgen05 + TMA implicit GEMM, weight stationary)", "conv": "conv_stem_kernel",
               "pre+stem": "pre_stem_mma_kernel (pre-processing fused into the stem)", "concat": "concat_multi_bf16_kernel",
               "dwconv": "dwconv_strip_kernel", "tcconvT2": "tc::conv_tc_ws_kernel with the fused DB-head tail epilogue",
               "eltwise": "eltwise_bf16x8_kernel"}.get(kernel_class(top["name"]), kernel_class(top["name"]))
    roof = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
            "traffic": traffic, "kernel": top["name"], "algorithmic_bytes_per_launch": top["bytes"],
            "launch_ms": top["ms"], "launches_per_step": 1, "share_of_network_time": top["ms"] / total_ms,
            "peak_source": peaks["source"],
            "tensor_tflops": top["flops"] / (top["ms"] / 1e3) / 1e12,
            "tensor_frac_of_sustained": top["flops"] / (top["ms"] / 1e3) / 1e12 / peaks["bf16_tflops_sustained"],
            "note": "kernel = %s for the named layer at 16 pages (%.0f FLOP per algorithmic byte, below the ridge of %.0f, so "
                    "HBM-bound); traffic = dram read+write of one launch from profiles/r2_traffic.json (ncu --set full); output "
                    "lines still resident in the 126 MB L2 at kernel end are not in it" % (
                        kern_of, top["flops"] / max(top["bytes"], 1.0), peaks["bf16_tflops_sustained"] * 1e3 / peaks["hbm_gbs"])}
    breakdown = {k: {"ms": round(v["ms"], 4), "share": round(v["ms"] / total_ms, 4),
                     "tflops": round(v["flops"] / max(v["ms"], 1e-9) / 1e9, 3),
                     "gbs": round(v["bytes"] / max(v["ms"], 1e-9) / 1e6, 1), "launches": v["launches"]}
                 for k, v in sorted(classes.items(), key=lambda kv: -kv[1]["ms"])}
    net_bytes = sum(s["bytes"] for s in prof)
    net_flops = sum(s["flops"] for s in prof)

    out = {
        "metric": METRIC, "value": value, "unit": "pages/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.precision, "data": "synthetic",
        "config": CONFIG,
        "boxes_per_step": n_boxes / max(args.steps, 1),
        "e2e": {"value": e2e, "unit": "pages/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "sync": sync,
        "gpu_launches": int(launches),
        "network_ms_per_step": net_ms,
        "network_algorithmic": {"gbytes_per_step": net_bytes / 1e9, "gflop_per_step": net_flops / 1e9,
                                "hbm_gbs": net_bytes / (net_ms / 1e3) / 1e9,
                                "hbm_frac_of_measured": net_bytes / (net_ms / 1e3) / 1e9 / peaks["hbm_gbs"]},
        "roofline": roof,
        "kernel_classes": breakdown,
        "clocks": sampler.summary(),
    }
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_arm(pages[:PAGES_PER_GPU], 8, 1, sample_pages=PAGES_PER_GPU)          # ~10-25 s of CPU work on the box's cores
        out["cpu_baseline"] = {"value": r["value"], "unit": "pages/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

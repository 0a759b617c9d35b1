"""Dev-time tool: fit the synthetic DBNet graph to rendered pages so its probability map
looks like a text detector's (line-shaped blobs) instead of white noise.

The real `det.onnx` is missing from the reference mount; an untrained graph thresholds
into ~10^5 speckle components per page, which is not the workload DBPostProcess sees in
production.  This script optimises the *same* graph (same nodes, same shapes) on
`synth_pages.render_page` crops against shrunk line masks with torch autograd on CPU and
stores the float constants as fp16 in `onnxocr_b200/assets/det_synth_weights.npz`.
`synth_graphs.build_det` overlays them on its seeded draws when the file exists.

    python tools/train_synth_det.py [steps]
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onnxocr_b200 import synth_graphs, synth_pages          # noqa: E402
from oracle.torch_interp import TorchGraph                  # noqa: E402  (dev tool, not product path)

MEAN = np.array([0.485, 0.456, 0.406], np.float32)
STD = np.array([0.229, 0.224, 0.225], np.float32)


def make_sample(seed, crop=320):
    rng = np.random.default_rng(seed)
    side = 960
    lines = int(rng.integers(8, 45))
    page, boxes = synth_pages.render_page(side, side, lines, seed, min_px=16, max_px=40)
    if rng.random() < 0.3:       # grey / tinted backgrounds
        tint = rng.integers(170, 256, size=3).astype(np.float32) / 255.0
        page = (page.astype(np.float32) * tint).astype(np.uint8)
    tgt = np.zeros((side, side), np.float32)
    for (x0, y0, x1, y1) in boxes:
        h = y1 - y0
        s = int(round(0.28 * h))
        tgt[y0 + s:y1 - s, x0 + s:x1 - s] = 1.0
    oy = int(rng.integers(0, side - crop + 1))
    ox = int(rng.integers(0, side - crop + 1))
    img = page[oy:oy + crop, ox:ox + crop]
    x = ((img.astype(np.float32) * np.float32(1 / 255.0) - MEAN) / STD).transpose(2, 0, 1)
    return x, tgt[oy:oy + crop, ox:ox + crop]


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 600
    torch.set_num_threads(int(os.environ.get("TRAIN_THREADS", "4")))
    torch.manual_seed(0)
    path = synth_graphs.ensure_models()["det"]
    g = TorchGraph(path)
    params = {}
    for k, v in g.consts.items():
        trainable = v.is_floating_point() and (k.endswith(".w_0") or k.endswith(".b_0") or k.startswith("lab_"))
        if trainable:
            params[k] = torch.nn.Parameter(v.clone())
    opt = torch.optim.Adam(list(params.values()), lr=2e-3)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=3e-3, total_steps=steps)
    out_name = g.output_names()[0]
    t0 = time.time()
    for step in range(steps):
        xs, ts = zip(*[make_sample(100000 + step * 4 + i) for i in range(4)])
        env = dict(g.consts)
        env.update(params)
        env["x"] = torch.from_numpy(np.stack(xs))
        for n in g.nodes:
            ins = [env[i] if i != "" else None for i in n["input"]]
            out = getattr(g, "op_" + n["op_type"])(n["attrs"], *ins)
            env[n["output"][0]] = out
        p = env[out_name][:, 0].clamp(1e-6, 1 - 1e-6)
        t = torch.from_numpy(np.stack(ts))
        bce = -(t * p.log() * 3.0 + (1 - t) * (1 - p).log()).mean()
        dice = 1 - (2 * (p * t).sum() + 1) / (p.sum() + t.sum() + 1)
        loss = bce + dice
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(list(params.values()), 5.0)
        opt.step()
        sched.step()
        if step % 20 == 0 or step == steps - 1:
            print("step %d loss %.4f bce %.4f dice %.4f  %.0fs" % (step, loss.item(), bce.item(), dice.item(),
                                                                  time.time() - t0), flush=True)
    out = os.path.join(synth_graphs.ASSET_DIR, "det_synth_weights.npz")
    np.savez_compressed(out, **{k: v.detach().numpy().astype(np.float16) for k, v in params.items()})
    print("wrote", out, os.path.getsize(out))


if __name__ == "__main__":
    main()

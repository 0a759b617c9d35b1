"""Dev-time tool: fit the synthetic SVTR_LCNet recogniser graph to rendered ASCII text lines.

The real `rec.onnx` is missing from the reference mount.  An untrained graph emits
low-confidence garbage (every line falls under `drop_score`) and its top-1/top-2 logit gaps
are noise-sized, so "identical text lines in bf16 mode" cannot be assessed on it.  This script
optimises the same graph with CTC loss on rendered lines (the glyph set of
`synth_pages`, printable ASCII) with torch autograd on CPU.  Only the classes it trains are
stored for the 120 -> 18385 FC (all other columns keep their seeded values and a -12 bias);
everything is saved as fp16 in `onnxocr_b200/assets/rec_synth_weights.npz`, which
`synth_graphs.build_rec` overlays on its seeded draws.

    python tools/train_synth_rec.py [steps]
"""
import math
import os
import string
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F
from PIL import Image, ImageDraw, ImageFont

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onnxocr_b200 import synth_graphs, synth_pages          # noqa: E402
from oracle.torch_interp import TorchGraph                  # noqa: E402  (dev tool, not product path)
from oracle import restate_pre                              # noqa: E402

CHARS = string.ascii_letters + string.digits + ".,-:/()%$#@&+=' "


def load_dict():
    lines = open(os.path.join(synth_graphs.ASSET_DIR, "ppocrv5_dict.txt"), "rb").read().decode("utf-8").split("\n")
    d = {c: i + 1 for i, c in enumerate(lines[:18383])}
    d[" "] = 18384
    return d


def make_line(rng):
    if rng.random() < 0.6:
        text = synth_pages._sentence(rng, int(rng.integers(4, 40)))
    else:
        text = "".join(CHARS[int(i)] for i in rng.integers(0, len(CHARS), int(rng.integers(3, 24)))).strip() or "a"
    text = " ".join(text.split())
    px = int(rng.integers(18, 44))
    font = ImageFont.load_default(size=px)
    tmp = Image.new("RGB", (8, 8))
    bb = ImageDraw.Draw(tmp).textbbox((0, 0), text, font=font)
    mx, my = int(rng.integers(1, 10)), int(rng.integers(1, 8))
    w, h = bb[2] - bb[0] + 2 * mx, bb[3] - bb[1] + 2 * my
    bg = int(rng.integers(200, 256))
    img = Image.new("RGB", (w, h), (bg, bg, bg))
    ImageDraw.Draw(img).text((mx - bb[0], my - bb[1]), text, font=font, fill=(int(rng.integers(0, 90)),) * 3)
    a = np.array(img)[:, :, ::-1].copy()
    if rng.random() < 0.3:
        a = np.clip(a.astype(np.int16) + rng.integers(-12, 13, a.shape), 0, 255).astype(np.uint8)
    return a, text


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    torch.set_num_threads(int(os.environ.get("TRAIN_THREADS", "4")))
    torch.manual_seed(0)
    cdict = load_dict()
    used = sorted(set(cdict[c] for c in CHARS) | {0})
    col_of = {cls: i for i, cls in enumerate(used)}
    blob, _ = synth_graphs.build_rec(fitted=False)
    path = "/tmp/rec_untrained.onnx"
    open(path, "wb").write(blob)
    g = TorchGraph(path)
    fc_w = [n["input"][1] for n in g.nodes if n["op_type"] == "MatMul" and g.consts.get(n["input"][1]) is not None
            and g.consts[n["input"][1]].shape[-1] == 18385][0]
    fc_name = fc_w[:-len(".w_0")]
    fc_b = fc_name + ".b_0"
    params = {}
    for k, v in g.consts.items():
        if v.is_floating_point() and (k.endswith(".w_0") or k.endswith(".b_0") or k.startswith("lab_")) and k not in (fc_w, fc_b):
            params[k] = torch.nn.Parameter(v.clone())
    used_t = torch.tensor(used)
    w_cols = torch.nn.Parameter(g.consts[fc_w][:, used_t].clone() * 0.3)
    b_cols = torch.nn.Parameter(torch.zeros(len(used)))
    plist = list(params.values()) + [w_cols, b_cols]
    opt = torch.optim.Adam(plist, lr=1e-3)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=2e-3, total_steps=steps, pct_start=0.15)
    # the sub-graph up to the FC input is run through the interpreter; the FC is applied on the used columns only
    fc_node = [n for n in g.nodes if n["op_type"] == "MatMul" and n["input"][1] == fc_w][0]
    fc_in = fc_node["input"][0]
    body = g.nodes[:g.nodes.index(fc_node)]
    rng = np.random.default_rng(1234)
    t0 = time.time()
    B = 12
    for step in range(steps):
        items = [make_line(rng) for _ in range(B)]
        max_wh = max(320 / 48.0, max(im.shape[1] / im.shape[0] for im, _ in items))
        W = int(48 * max_wh)
        x = np.stack([restate_pre.resize_norm_img(im, W) for im, _ in items]).astype(np.float32)
        env = dict(g.consts)
        env.update(params)
        env["x"] = torch.from_numpy(x)
        for n in body:
            ins = [env[i] if i != "" else None for i in n["input"]]
            env[n["output"][0]] = getattr(g, "op_" + n["op_type"])(n["attrs"], *ins)
        feat = env[fc_in]                                   # [B, T, 120]
        logits = feat @ w_cols + b_cols                     # [B, T, n_used]
        logp = F.log_softmax(logits, dim=2).permute(1, 0, 2)
        T = logp.shape[0]
        tgt = [torch.tensor([col_of[cdict[c]] for c in t]) for _, t in items]
        tl = torch.tensor([len(t) for t in tgt])
        loss = F.ctc_loss(logp, torch.cat(tgt), torch.full((B,), T, dtype=torch.long), tl, blank=0, zero_infinity=True)
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(plist, 5.0)
        opt.step()
        sched.step()
        if step % 20 == 0 or step == steps - 1:
            with torch.no_grad():
                pred = logits.argmax(2)
                ok = 0
                for b in range(B):
                    seq, prev = [], -1
                    for v in pred[b].tolist():
                        if v != prev and v != 0:
                            seq.append(v)
                        prev = v
                    ok += int(seq == tgt[b].tolist())
            print("step %d loss %.4f line-acc %d/%d W=%d  %.0fs" % (step, loss.item(), ok, B, W, time.time() - t0), flush=True)
        if (step % 500 == 499) or step == steps - 1:
            out = {k: v.detach().numpy().astype(np.float16) for k, v in params.items()}
            out[fc_name + ".cols"] = np.array(used, np.int32)
            out[fc_name + ".w_cols"] = w_cols.detach().numpy().astype(np.float16)
            out[fc_name + ".b_cols"] = b_cols.detach().numpy().astype(np.float16)
            dst = os.path.join(synth_graphs.ASSET_DIR, "rec_synth_weights.npz")
            np.savez_compressed(dst + ".tmp.npz", **out)
            os.replace(dst + ".tmp.npz", dst)
            print("wrote", dst, os.path.getsize(dst), flush=True)


if __name__ == "__main__":
    main()

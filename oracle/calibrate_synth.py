"""Dev-time tool: per-conv gain table for the synthetic det/rec graphs (TEST/DEV INFRASTRUCTURE).

Random fan-in-scaled weights let activations drift by >100x through ~60 layers, which
makes the final sigmoid saturate and a bf16-vs-fp32 comparison meaningless.  This script
walks the un-calibrated graph once with the torch interpreter, rescales every spatial
Conv/ConvTranspose so its output RMS is 1 on a rendered calibration input, and stores
only the scalar gains in `onnxocr_b200/synth_gains.json` (committed).  The graph builder
multiplies its seeded draws by those constants, so the .onnx bytes stay reproducible on
every box without running torch at build time.

    python -m oracle.calibrate_synth
"""
import json
import os
import tempfile

import numpy as np
import torch

from onnxocr_b200 import synth_graphs, synth_pages
from oracle.torch_interp import TorchGraph


def _calibrate(blob, x, final_rms=None):
    with tempfile.NamedTemporaryFile(suffix=".onnx", delete=False) as fh:
        fh.write(blob)
        path = fh.name
    g = TorchGraph(path)
    os.unlink(path)
    env = dict(g.consts)
    env[g.input_names()[0]] = torch.from_numpy(x)
    gains = {}
    convs = [n for n in g.nodes if n["op_type"] in ("Conv", "ConvTranspose")]
    last = convs[-1]
    with torch.no_grad():
        for n in g.nodes:
            fn = getattr(g, "op_" + n["op_type"])
            ins = [env[i] if i != "" else None for i in n["input"]]
            out = fn(n["attrs"], *ins)
            if n["op_type"] in ("Conv", "ConvTranspose") and out.shape[2] * out.shape[3] > 1:
                target = 1.0
                if n is last and final_rms is not None:
                    target = final_rms
                rms = float(out.pow(2).mean().sqrt())
                gain = target / max(rms, 1e-12)
                gains[n["input"][1]] = gain
                env[n["input"][1]] = ins[1] * gain
                if len(ins) > 2 and ins[2] is not None:
                    env[n["input"][2]] = ins[2] * gain
                ins = [env[i] if i != "" else None for i in n["input"]]
                out = fn(n["attrs"], *ins)
            outs = out if isinstance(out, (tuple, list)) else [out]
            for name, val in zip(n["output"], outs):
                env[name] = val
    return gains, env[g.output_names()[0]].numpy()


def main():
    mean = np.array([0.485, 0.456, 0.406], np.float32)
    std = np.array([0.229, 0.224, 0.225], np.float32)
    page = synth_pages.render_page(640, 640, 24, seed=7)[0]
    x = ((page.astype(np.float32) * np.float32(1 / 255.0) - mean) / std).transpose(2, 0, 1)[None].copy()
    det_blob, _ = synth_graphs.build_det(gains={})
    det_gains, prob = _calibrate(det_blob, x, final_rms=2.0)
    print("det prob: mean %.3f  frac>0.3 %.3f" % (prob.mean(), (prob > 0.3).mean()))
    crops = synth_pages.rec_crops_c3(8, seed=9)
    xr = np.stack([((c.astype(np.float32).transpose(2, 0, 1) / 255) - 0.5) / 0.5 for c in crops]).astype(np.float32)
    rec_blob, _ = synth_graphs.build_rec(gains={})
    rec_gains, probs = _calibrate(rec_blob, xr)
    print("rec max-prob mean %.3f, distinct argmax %d" % (probs.max(axis=2).mean(), len(np.unique(probs.argmax(axis=2)))))
    out = os.path.join(os.path.dirname(synth_graphs.__file__), "synth_gains.json")
    with open(out, "w") as fh:
        json.dump({"det": det_gains, "rec": rec_gains}, fh, indent=0, sort_keys=True)
    print("wrote", out, len(det_gains), len(rec_gains))


if __name__ == "__main__":
    main()

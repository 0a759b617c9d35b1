"""Architecture-faithful PP-OCRv5 mobile det / rec ONNX graphs with seeded weights.

The reference's `onnxocr/models/ppocrv5/det/det.onnx` and `rec/rec.onnx` are
missing from the mount (`/root/reference/.MISSING_LARGE_BLOBS`), so throughput
and parity are measured on graphs that reproduce the upstream architectures
 in the export style verified on the shipped `cls.onnx`
(paddle2onnx, opset 11: weights as `Constant` nodes, BatchNormalization kept
separate, hardswish decomposed as Add(3)/Clip(0,6)/Mul/Div(6),
`HardSigmoid(alpha=0.2, beta=0.5)`, SE = GAP/Conv/Relu/Conv/HardSigmoid/Mul).

  det = PP-LCNetV3(scale 0.75, det) + RSEFPN(96, shortcut) + DBHead      [1,3,H,W] -> [1,1,H,W]
  rec = PP-LCNetV3(scale 0.95, rec) + EncoderWithSVTR(120, depth 2, 8 heads) + CTC FC -> softmax
                                                                          [B,3,48,W] -> [B,W/8,18385]

Weights are `numpy.random.default_rng(seed)` draws with fan-in scaling times the
committed per-layer gains in `synth_gains.json`; for det, the committed
`assets/det_synth_weights.npz` (fp16, fitted on rendered pages by
`tools/train_synth_det.py`) is overlaid so that the probability map shows
text-line blobs rather than noise.  The .onnx bytes are therefore bit-reproducible
on every box (nothing is downloaded).
When the real blobs appear they load through the same engine entry point.
"""
import os
import numpy as np

from . import onnx_writer as ow

_HERE = os.path.dirname(os.path.abspath(__file__))
ASSET_DIR = os.path.join(_HERE, "assets")


def make_divisible(v, divisor=16, min_value=None):
    if min_value is None:
        min_value = divisor
    new_v = max(min_value, int(v + divisor / 2) // divisor * divisor)
    if new_v < 0.9 * v:
        new_v += divisor
    return new_v


class _G(object):
    """Tiny graph builder that names things the way paddle2onnx does."""

    def __init__(self, seed, gains=None, overrides=None):
        self.rng = np.random.default_rng(seed)
        self.gains = gains or {}
        self.overrides = overrides or {}
        self.nodes = []
        self.cnt = {}
        self.params = 0
        self.macs_per_px = 0.0   # filled by callers that track spatial scale

    def _n(self, prefix):
        i = self.cnt.get(prefix, 0)
        self.cnt[prefix] = i + 1
        return "%s_%d" % (prefix, i)

    def const(self, arr, name=None):
        name = name or self._n("p2o.helper.constant")
        arr = np.asarray(arr)
        if name in self.overrides:          # fitted weights (tools/train_synth_det.py), stored as fp16
            arr = np.asarray(self.overrides[name]).astype(arr.dtype).reshape(arr.shape)
        self.nodes.append(ow.enc_node("Constant", [], [name], value=arr))
        if arr.dtype == np.float32 and arr.size > 1:
            self.params += arr.size
        return name

    def op(self, op_type, inputs, prefix=None, n_out=1, **attrs):
        outs = [self._n(prefix or ("p2o." + op_type)) + ".tmp" for _ in range(n_out)]
        self.nodes.append(ow.enc_node(op_type, inputs, outs, **attrs))
        return outs[0] if n_out == 1 else outs

    # ---- layers -------------------------------------------------------
    def conv(self, x, cin, cout, k, stride=1, groups=1, bias=False, gain=1.0):
        kh, kw = (k, k) if isinstance(k, int) else k
        sh, sw = (stride, stride) if isinstance(stride, int) else stride
        fan_in = (cin // groups) * kh * kw
        w = self.rng.standard_normal((cout, cin // groups, kh, kw)).astype(np.float32)
        w *= np.float32(gain / np.sqrt(fan_in))
        name = self._n("conv2d")
        cal = np.float32(self.gains.get(name + ".w_0", 1.0))
        ins = [x, self.const(w * cal, name + ".w_0")]
        if bias:
            b = (self.rng.standard_normal(cout) * 0.05).astype(np.float32)
            ins.append(self.const(b * cal, name + ".b_0"))
        return self.op("Conv", ins, prefix="conv2d.out", dilations=[1, 1], kernel_shape=[kh, kw],
                       strides=[sh, sw], group=groups, pads=[kh // 2, kw // 2, kh // 2, kw // 2])

    def bn(self, x, c):
        name = self._n("batch_norm")
        scale = (1.0 + 0.1 * self.rng.standard_normal(c)).astype(np.float32)
        bias = (0.1 * self.rng.standard_normal(c)).astype(np.float32)
        mean = (0.1 * self.rng.standard_normal(c)).astype(np.float32)
        var = (1.0 + 0.2 * self.rng.random(c)).astype(np.float32)
        ins = [x, self.const(scale, name + ".w_0"), self.const(bias, name + ".b_0"),
               self.const(mean, name + ".w_1"), self.const(var, name + ".w_2")]
        return self.op("BatchNormalization", ins, prefix="batch_norm.out",
                       epsilon=9.999999747378752e-06, momentum=0.8999999761581421)

    def hardswish(self, x):
        six = self.const(np.float32(6.0))
        three = self.const(np.float32(3.0))
        a = self.op("Add", [x, three])
        c = self.op("Clip", [a, self.const(np.float32(0.0)), self.const(np.float32(6.0))])
        m = self.op("Mul", [x, c])
        return self.op("Div", [m, six], prefix="hardswish")

    def relu(self, x):
        return self.op("Relu", [x], prefix="relu")

    def swish(self, x):
        s = self.op("Sigmoid", [x])
        return self.op("Mul", [x, s], prefix="swish")

    def lab(self, x):
        """LearnableAffineBlock: scalar scale * x + bias."""
        s = self.const(np.array([1.0 + 0.1 * self.rng.standard_normal()], dtype=np.float32), self._n("lab_scale"))
        b = self.const(np.array([0.1 * self.rng.standard_normal()], dtype=np.float32), self._n("lab_bias"))
        return self.op("Add", [self.op("Mul", [x, s]), b])

    def se(self, x, c, reduction=4):
        g = self.op("GlobalAveragePool", [x])
        h = self.relu(self.conv(g, c, c // reduction, 1, bias=True))
        h = self.conv(h, c // reduction, c, 1, bias=True)
        h = self.op("HardSigmoid", [h], prefix="hardsigmoid", alpha=0.20000000298023224, beta=0.5)
        return self.op("Mul", [x, h])

    def rep_layer(self, x, cin, cout, k, stride, groups):
        """LearnableRepLayer after re-parameterisation: conv(+bias) -> lab -> [hardswish -> lab]."""
        y = self.conv(x, cin, cout, k, stride=stride, groups=groups, bias=True, gain=1.6)
        y = self.lab(y)
        if stride != 2:       # upstream quirk: tuple strides such as (2, 1) still activate
            y = self.lab(self.hardswish(y))
        return y

    def lcnet_block(self, x, cin, cout, k, stride, use_se):
        y = self.rep_layer(x, cin, cin, k, stride, cin)
        if use_se:
            y = self.se(y, cin)
        return self.rep_layer(y, cin, cout, 1, 1, 1)

    def conv_bn_act(self, x, cin, cout, k, act):
        y = self.bn(self.conv(x, cin, cout, k, gain=1.4), cout)
        if act == "swish":
            y = self.swish(y)
        elif act == "relu":
            y = self.relu(y)
        elif act == "hardswish":
            y = self.hardswish(y)
        return y

    def linear(self, x, cin, cout, gain=1.0, bias_shift=0.0):
        name = self._n("linear")
        w = (self.rng.standard_normal((cin, cout)) * gain / np.sqrt(cin)).astype(np.float32)
        b = (self.rng.standard_normal(cout) * 0.05 + bias_shift).astype(np.float32)
        # fitted column subset (tools/train_synth_rec.py stores only the classes it trained)
        cols = self.overrides.get(name + ".cols")
        if cols is not None:
            cols = np.asarray(cols).astype(np.int64)
            w[:, cols] = np.asarray(self.overrides[name + ".w_cols"]).astype(np.float32)
            b[cols] = np.asarray(self.overrides[name + ".b_cols"]).astype(np.float32)
        y = self.op("MatMul", [x, self.const(w, name + ".w_0")])
        return self.op("Add", [y, self.const(b, name + ".b_0")])

    def layer_norm(self, x, c, eps):
        name = self._n("layer_norm")
        mean = self.op("ReduceMean", [x], axes=[-1], keepdims=1)
        d = self.op("Sub", [x, mean])
        p = self.op("Pow", [d, self.const(np.float32(2.0))])
        var = self.op("ReduceMean", [p], axes=[-1], keepdims=1)
        ve = self.op("Add", [var, self.const(np.float32(eps))])
        sd = self.op("Sqrt", [ve])
        y = self.op("Div", [d, sd])
        scale = (1.0 + 0.1 * self.rng.standard_normal(c)).astype(np.float32)
        bias = (0.1 * self.rng.standard_normal(c)).astype(np.float32)
        y = self.op("Mul", [y, self.const(scale, name + ".w_0")])
        return self.op("Add", [y, self.const(bias, name + ".b_0")], prefix="layer_norm.out")


_DET_CFG = {
    "blocks2": [[3, 16, 32, 1, False]],
    "blocks3": [[3, 32, 64, 2, False], [3, 64, 64, 1, False]],
    "blocks4": [[3, 64, 128, 2, False], [3, 128, 128, 1, False]],
    "blocks5": [[3, 128, 256, 2, False], [5, 256, 256, 1, False], [5, 256, 256, 1, False],
                [5, 256, 256, 1, False], [5, 256, 256, 1, False]],
    "blocks6": [[5, 256, 512, 2, True], [5, 512, 512, 1, True], [5, 512, 512, 1, False],
                [5, 512, 512, 1, False]],
}
_REC_CFG = {
    "blocks2": [[3, 16, 32, 1, False]],
    "blocks3": [[3, 32, 64, 1, False], [3, 64, 64, 1, False]],
    "blocks4": [[3, 64, 128, (2, 1), False], [3, 128, 128, 1, False]],
    "blocks5": [[3, 128, 256, (1, 2), False], [5, 256, 256, 1, False], [5, 256, 256, 1, False],
                [5, 256, 256, 1, False], [5, 256, 256, 1, False]],
    "blocks6": [[5, 256, 512, (2, 1), True], [5, 512, 512, 1, True], [5, 512, 512, (2, 1), False],
                [5, 512, 512, 1, False]],
}


def _backbone(g, x, cfg, scale):
    c0 = make_divisible(16 * scale)
    y = g.bn(g.conv(x, 3, c0, 3, stride=2, gain=1.4), c0)     # conv1: ConvBNLayer, no activation
    taps = []
    for stage in ("blocks2", "blocks3", "blocks4", "blocks5", "blocks6"):
        for (k, cin, cout, s, se) in cfg[stage]:
            y = g.lcnet_block(y, make_divisible(cin * scale), make_divisible(cout * scale), k, s, se)
        taps.append((y, make_divisible(cfg[stage][-1][2] * scale)))
    return y, taps


def _load_gains(kind):
    import json
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "synth_gains.json")
    if not os.path.exists(p):
        return {}
    with open(p) as fh:
        return json.load(fh).get(kind, {})


def _load_overrides(fname):
    p = os.path.join(ASSET_DIR, fname)
    if not os.path.exists(p):
        return {}
    with np.load(p) as z:
        return {k: z[k] for k in z.files}


def build_det(seed=50, gains=None, fitted=True):
    """DBNet: returns serialized ONNX bytes; input 'x' [N,3,H,W], output 'sigmoid_0.tmp_0' [N,1,H,W]."""
    g = _G(seed, _load_gains("det") if gains is None else gains,
           _load_overrides("det_synth_weights.npz") if fitted else None)
    scale = 0.75
    _, taps = _backbone(g, "x", _DET_CFG, scale)
    mv_c = [16, 24, 56, 480]
    feats = []
    for (t, c), oc in zip(taps[1:], mv_c):
        oc = int(oc * scale)
        feats.append((g.conv(t, c, oc, 1, bias=True), oc))
    # RSEFPN(out=96, shortcut=True)
    oc = 96

    def rse(x, cin, cout, k):
        y = g.conv(x, cin, cout, k, gain=1.2)
        return g.op("Add", [y, g.se(y, cout)])

    def up(x, f):
        scales = g.const(np.array([1.0, 1.0, float(f), float(f)], dtype=np.float32))
        roi = g.const(np.zeros((0,), dtype=np.float32))
        return g.op("Resize", [x, roi, scales], coordinate_transformation_mode="asymmetric",
                    mode="nearest", nearest_mode="floor")

    in5 = rse(feats[3][0], feats[3][1], oc, 1)
    in4 = rse(feats[2][0], feats[2][1], oc, 1)
    in3 = rse(feats[1][0], feats[1][1], oc, 1)
    in2 = rse(feats[0][0], feats[0][1], oc, 1)
    out4 = g.op("Add", [in4, up(in5, 2)])
    out3 = g.op("Add", [in3, up(out4, 2)])
    out2 = g.op("Add", [in2, up(out3, 2)])
    p5 = rse(in5, oc, oc // 4, 3)
    p4 = rse(out4, oc, oc // 4, 3)
    p3 = rse(out3, oc, oc // 4, 3)
    p2 = rse(out2, oc, oc // 4, 3)
    fuse = g.op("Concat", [up(p5, 8), up(p4, 4), up(p3, 2), p2], axis=1)
    # DBHead.binarize
    y = g.relu(g.bn(g.conv(fuse, oc, oc // 4, 3, gain=1.4), oc // 4))

    def conv_t(x, cin, cout):
        name = g._n("conv2d_transpose")
        w = (g.rng.standard_normal((cin, cout, 2, 2)) * 1.4 / np.sqrt(cin)).astype(np.float32)
        b = (g.rng.standard_normal(cout) * 0.05).astype(np.float32)
        cal = np.float32(g.gains.get(name + ".w_0", 1.0))
        return g.op("ConvTranspose", [x, g.const(w * cal, name + ".w_0"), g.const(b * cal, name + ".b_0")],
                    dilations=[1, 1], kernel_shape=[2, 2], strides=[2, 2], group=1, pads=[0, 0, 0, 0])

    y = g.relu(g.bn(conv_t(y, oc // 4, oc // 4), oc // 4))
    y = conv_t(y, oc // 4, 1)
    g.nodes.append(ow.enc_node("Sigmoid", [y], ["sigmoid_0.tmp_0"]))
    n = "p2o.DynamicDimension."
    model = ow.enc_model(g.nodes, [ow.enc_value_info("x", [n + "0", 3, n + "1", n + "2"])],
                         [ow.enc_value_info("sigmoid_0.tmp_0", [n + "3", 1, n + "4", n + "5"])],
                         graph_name="ppocrv5_mobile_det_synth")
    return model, g.params


def build_rec(seed=51, n_classes=18385, gains=None, fitted=True):
    """SVTR_LCNet recogniser: input 'x' [B,3,48,W], output 'softmax_0.tmp_0' [B,W/8,n_classes]."""
    ov = _load_overrides("rec_synth_weights.npz") if fitted else {}
    g = _G(seed, _load_gains("rec") if gains is None else gains, ov)
    scale = 0.95
    y, taps = _backbone(g, "x", _REC_CFG, scale)
    c = taps[-1][1]                                           # 480
    y = g.op("AveragePool", [y], kernel_shape=[3, 2], strides=[3, 2], pads=[0, 0, 0, 0], ceil_mode=0)
    # EncoderWithSVTR(in=480, dims=120, depth=2, hidden=120, heads=8, kernel [1,3], use_guide)
    dims, hidden, heads = 120, 120, 8
    h = y
    z = g.conv_bn_act(y, c, c // 8, (1, 3), "swish")
    z = g.conv_bn_act(z, c // 8, hidden, 1, "swish")
    # z.flatten(2).transpose([0, 2, 1])  -> [B, HW, C]
    shp = g.op("Shape", [z])
    z = g.op("Reshape", [z, g.const(np.array([0, 0, -1], dtype=np.int64))])
    z = g.op("Transpose", [z], perm=[0, 2, 1])
    hd = hidden // heads
    for _ in range(2):
        # x = x + mixer(norm1(x))
        n1 = g.layer_norm(z, hidden, 1e-5)
        qkv = g.linear(n1, hidden, 3 * hidden)
        qkv = g.op("Reshape", [qkv, g.const(np.array([0, -1, 3, heads, hd], dtype=np.int64))])
        qkv = g.op("Transpose", [qkv], perm=[2, 0, 3, 1, 4])
        parts = []
        for i in range(3):
            s = g.op("Slice", [qkv, g.const(np.array([i], dtype=np.int64)), g.const(np.array([i + 1], dtype=np.int64)),
                               g.const(np.array([0], dtype=np.int64)), g.const(np.array([1], dtype=np.int64))])
            parts.append(g.op("Squeeze", [s], axes=[0]))
        q = g.op("Mul", [parts[0], g.const(np.array([hd ** -0.5], dtype=np.float32))])
        kt = g.op("Transpose", [parts[1]], perm=[0, 1, 3, 2])
        attn = g.op("Softmax", [g.op("MatMul", [q, kt])], axis=-1)
        o = g.op("MatMul", [attn, parts[2]])
        o = g.op("Transpose", [o], perm=[0, 2, 1, 3])
        o = g.op("Reshape", [o, g.const(np.array([0, -1, hidden], dtype=np.int64))])
        o = g.linear(o, hidden, hidden, gain=0.5)
        z = g.op("Add", [z, o])
        # x = x + mlp(norm2(x))
        n2 = g.layer_norm(z, hidden, 1e-5)
        m = g.swish(g.linear(n2, hidden, 2 * hidden))
        m = g.linear(m, 2 * hidden, hidden, gain=0.5)
        z = g.op("Add", [z, m])
    z = g.layer_norm(z, hidden, 1e-6)
    # z.reshape([0, H, W, C]).transpose([0, 3, 1, 2])
    hw = g.op("Slice", [shp, g.const(np.array([2], dtype=np.int64)), g.const(np.array([4], dtype=np.int64)),
                        g.const(np.array([0], dtype=np.int64)), g.const(np.array([1], dtype=np.int64))])
    tgt = g.op("Concat", [g.const(np.array([0], dtype=np.int64)), hw, g.const(np.array([hidden], dtype=np.int64))], axis=0)
    z = g.op("Reshape", [z, tgt])
    z = g.op("Transpose", [z], perm=[0, 3, 1, 2])
    z = g.conv_bn_act(z, hidden, c, 1, "swish")
    z = g.op("Concat", [h, z], axis=1)
    z = g.conv_bn_act(z, 2 * c, c // 8, (1, 3), "swish")
    z = g.conv_bn_act(z, c // 8, dims, 1, "swish")
    # Im2Seq: squeeze(axis=2) + transpose([0, 2, 1])
    z = g.op("Squeeze", [z], axes=[2])
    z = g.op("Transpose", [z], perm=[0, 2, 1])
    # CTCHead.fc + softmax(axis=2)
    # when fitted weights exist, classes outside the fitted subset keep small seeded weights and a
    # strongly negative bias, so they never win the argmax (the FC still computes all 18385 logits)
    z = g.linear(z, dims, n_classes, gain=0.5 if ov else 6.0, bias_shift=-12.0 if ov else 0.0)
    g.nodes.append(ow.enc_node("Softmax", [z], ["softmax_0.tmp_0"], axis=2))
    n = "p2o.DynamicDimension."
    model = ow.enc_model(g.nodes, [ow.enc_value_info("x", [n + "0", 3, 48, n + "1"])],
                         [ow.enc_value_info("softmax_0.tmp_0", [n + "2", n + "3", n_classes])],
                         graph_name="ppocrv5_mobile_rec_synth")
    return model, g.params


def ensure_models(asset_dir=None):
    """Write det/rec synthetic graphs next to the package (idempotent); return their paths."""
    asset_dir = asset_dir or ASSET_DIR
    os.makedirs(asset_dir, exist_ok=True)
    out = {}
    for kind, fn in (("det", build_det), ("rec", build_rec)):
        p = os.path.join(asset_dir, "ppocrv5_%s_synth.onnx" % kind)
        deps = [os.path.join(_HERE, "synth_gains.json"), os.path.join(ASSET_DIR, "%s_synth_weights.npz" % kind),
                os.path.abspath(__file__)]
        stale = os.path.exists(p) and any(os.path.exists(d) and os.path.getmtime(d) > os.path.getmtime(p) for d in deps)
        if stale or not os.path.exists(p):
            blob, _ = fn()
            tmp = p + ".tmp%d" % os.getpid()
            with open(tmp, "wb") as fh:
                fh.write(blob)
            os.replace(tmp, p)
        out[kind] = p
    return out


if __name__ == "__main__":
    for k, p in ensure_models().items():
        print(k, p, os.path.getsize(p))

"""torch-CPU fp32 interpreter for paddle2onnx graphs (TEST INFRASTRUCTURE — oracle side).

Stand-in for the reference's `onnxruntime.InferenceSession(...).run(...)` at
`/root/reference/onnxocr/predict_base.py:14`, `predict_det.py:107`,
`predict_cls.py:75`, `predict_rec.py:316`.  onnxruntime==1.14.1 is pinned in
`/root/reference/requirements.txt:5` but is absent from this image and its
wheelhouse, so the graph semantics (ONNX opset 11 operator definitions) are
restated here with torch CPU fp32 kernels.  Every number produced by this file
is "torch-CPU restatement of the ORT-CPU path": **parity unpinned** against the
real onnxruntime (no reference test holds a golden vector for it).

Only tests/, bench.py's reference/cpu_baseline legs and the golden generator may
import this module.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import onnx_wire

_NP2T = {np.dtype(np.float32): torch.float32, np.dtype(np.int64): torch.int64,
         np.dtype(np.int32): torch.int32, np.dtype(np.bool_): torch.bool,
         np.dtype(np.float64): torch.float64, np.dtype(np.uint8): torch.uint8}
_ONNX2T = {1: torch.float32, 6: torch.int32, 7: torch.int64, 9: torch.bool, 11: torch.float64}


def _pads(p):
    # ONNX pads = [t, l, b, r]; torch conv wants symmetric or explicit F.pad
    return p if p is not None else [0, 0, 0, 0]


class TorchGraph(object):
    """Executes an ONNX graph node by node (topological order as stored)."""

    def __init__(self, path_or_model):
        m = onnx_wire.load(path_or_model) if isinstance(path_or_model, str) else path_or_model
        self.model = m
        self.inputs = m["inputs"]
        self.outputs = m["outputs"]
        self.consts = {}
        for k, v in m["initializers"].items():
            self.consts[k] = torch.from_numpy(np.ascontiguousarray(v))
        self.nodes = []
        for n in m["nodes"]:
            if n["op_type"] == "Constant":
                self.consts[n["output"][0]] = torch.from_numpy(np.ascontiguousarray(n["attrs"]["value"]))
            else:
                self.nodes.append(n)

    # -- the same three-method surface PredictBase uses -------------------
    def input_names(self):
        return [i["name"] for i in self.inputs]

    def output_names(self):
        return [o["name"] for o in self.outputs]

    @torch.no_grad()
    def run(self, output_names, input_feed, keep=None):
        env = dict(self.consts)
        for k, v in input_feed.items():
            env[k] = torch.from_numpy(np.ascontiguousarray(v))
        for n in self.nodes:
            fn = getattr(self, "op_" + n["op_type"], None)
            if fn is None:
                raise NotImplementedError("torch_interp: op %s" % n["op_type"])
            ins = [env[i] if i != "" else None for i in n["input"]]
            outs = fn(n["attrs"], *ins)
            if not isinstance(outs, (tuple, list)):
                outs = [outs]
            for name, val in zip(n["output"], outs):
                env[name] = val
        if keep is not None:
            keep.update({k: v for k, v in env.items() if k not in self.consts})
        if output_names is None:
            output_names = self.output_names()
        return [env[o].numpy() for o in output_names]

    # -- operators ---------------------------------------------------------
    def op_Conv(self, a, x, w, b=None):
        pads = _pads(a.get("pads"))
        strides = a.get("strides", [1, 1])
        dil = a.get("dilations", [1, 1])
        g = a.get("group", 1)
        if pads[0] == pads[2] and pads[1] == pads[3]:
            return F.conv2d(x, w, b, stride=strides, padding=(pads[0], pads[1]), dilation=dil, groups=g)
        x = F.pad(x, (pads[1], pads[3], pads[0], pads[2]))
        return F.conv2d(x, w, b, stride=strides, dilation=dil, groups=g)

    def op_ConvTranspose(self, a, x, w, b=None):
        pads = _pads(a.get("pads"))
        return F.conv_transpose2d(x, w, b, stride=a.get("strides", [1, 1]), padding=(pads[0], pads[1]),
                                  output_padding=tuple(a.get("output_padding", [0, 0])),
                                  groups=a.get("group", 1), dilation=a.get("dilations", [1, 1]))

    def op_BatchNormalization(self, a, x, scale, bias, mean, var):
        return F.batch_norm(x, mean, var, scale, bias, training=False, eps=a.get("epsilon", 1e-5))

    def op_Add(self, a, x, y):
        return x + y

    def op_Sub(self, a, x, y):
        return x - y

    def op_Mul(self, a, x, y):
        return x * y

    def op_Div(self, a, x, y):
        if not x.is_floating_point() and not y.is_floating_point():
            return torch.div(x, y, rounding_mode="trunc")
        return x / y

    def op_Pow(self, a, x, y):
        return torch.pow(x, y)

    def op_Sqrt(self, a, x):
        return torch.sqrt(x)

    def op_Clip(self, a, x, lo=None, hi=None):
        lo = a.get("min", None) if lo is None else lo
        hi = a.get("max", None) if hi is None else hi
        return torch.clamp(x, min=None if lo is None else float(lo), max=None if hi is None else float(hi))

    def op_Relu(self, a, x):
        return F.relu(x)

    def op_Sigmoid(self, a, x):
        return torch.sigmoid(x)

    def op_HardSigmoid(self, a, x):
        return torch.clamp(x * a.get("alpha", 0.2) + a.get("beta", 0.5), 0.0, 1.0)

    def op_HardSwish(self, a, x):
        return x * torch.clamp(x / 6.0 + 0.5, 0.0, 1.0)

    def op_GlobalAveragePool(self, a, x):
        return x.mean(dim=(2, 3), keepdim=True)

    def op_MaxPool(self, a, x):
        pads = _pads(a.get("pads"))
        return F.max_pool2d(x, a["kernel_shape"], a.get("strides", [1, 1]), (pads[0], pads[1]),
                            ceil_mode=bool(a.get("ceil_mode", 0)))

    def op_AveragePool(self, a, x):
        pads = _pads(a.get("pads"))
        return F.avg_pool2d(x, a["kernel_shape"], a.get("strides", [1, 1]), (pads[0], pads[1]),
                            ceil_mode=bool(a.get("ceil_mode", 0)),
                            count_include_pad=bool(a.get("count_include_pad", 0)))

    def op_Shape(self, a, x):
        return torch.tensor(list(x.shape), dtype=torch.int64)

    def op_Cast(self, a, x):
        return x.to(_ONNX2T[a["to"]])

    def op_Slice(self, a, x, starts=None, ends=None, axes=None, steps=None):
        starts = a["starts"] if starts is None else starts.tolist()
        ends = a["ends"] if ends is None else ends.tolist()
        axes = (a.get("axes") if axes is None else axes.tolist()) or list(range(len(starts)))
        steps = [1] * len(starts) if steps is None else steps.tolist()
        idx = [slice(None)] * x.dim()
        for s, e, ax, st in zip(starts, ends, axes, steps):
            dim = x.shape[ax]
            s = max(min(s + dim if s < 0 else s, dim), 0)
            e = max(min(e + dim if e < 0 else e, dim), 0)
            idx[ax] = slice(s, e, st)
        return x[tuple(idx)]

    def op_Concat(self, a, *xs):
        return torch.cat(list(xs), dim=a["axis"])

    def op_Reshape(self, a, x, shape):
        shape = shape.tolist()
        shape = [x.shape[i] if s == 0 else s for i, s in enumerate(shape)]
        return x.reshape(shape)

    def op_Transpose(self, a, x):
        return x.permute(a["perm"]).contiguous()

    def op_Squeeze(self, a, x, axes=None):
        axes = a.get("axes") if axes is None else axes.tolist()
        for ax in sorted([ax % x.dim() for ax in axes], reverse=True):
            x = x.squeeze(ax)
        return x

    def op_Unsqueeze(self, a, x, axes=None):
        axes = a.get("axes") if axes is None else axes.tolist()
        for ax in sorted(axes):
            x = x.unsqueeze(ax)
        return x

    def op_Flatten(self, a, x):
        ax = a.get("axis", 1)
        return x.reshape(int(np.prod(x.shape[:ax])), -1)

    def op_MatMul(self, a, x, y):
        return torch.matmul(x, y)

    def op_Gemm(self, a, x, w, b=None):
        if a.get("transA", 0):
            x = x.t()
        if a.get("transB", 0):
            w = w.t()
        y = a.get("alpha", 1.0) * (x @ w)
        return y + a.get("beta", 1.0) * b if b is not None else y

    def op_Softmax(self, a, x):
        ax = a.get("axis", 1)
        # opset-11 semantics: flatten to 2-D at `axis`; identical to a plain
        # last-dim softmax for the graphs here (axis is always the last one).
        assert ax % x.dim() == x.dim() - 1, "softmax over a non-last axis"
        return torch.softmax(x, dim=-1)

    def op_ReduceMean(self, a, x):
        return x.mean(dim=a["axes"], keepdim=bool(a.get("keepdims", 1)))

    def op_Resize(self, a, x, roi=None, scales=None, sizes=None):
        assert a.get("mode", b"nearest") in (b"nearest", "nearest")
        if sizes is not None and sizes.numel():
            oh, ow = int(sizes[2]), int(sizes[3])
        else:
            oh = int(np.floor(x.shape[2] * float(scales[2])))
            ow = int(np.floor(x.shape[3] * float(scales[3])))
        # asymmetric + floor == integer replication for integer scale factors
        iy = (torch.arange(oh) * x.shape[2] // oh)
        ix = (torch.arange(ow) * x.shape[3] // ow)
        return x[:, :, iy][:, :, :, ix]

    def op_Gather(self, a, x, idx):
        return torch.index_select(x, a.get("axis", 0), idx.reshape(-1).long()).reshape(
            list(x.shape[:a.get("axis", 0)]) + list(idx.shape) + list(x.shape[a.get("axis", 0) + 1:]))

    def op_Identity(self, a, x):
        return x


class TorchSession(object):
    """Duck-typed `onnxruntime.InferenceSession` (the three methods PredictBase touches)."""

    class _Arg(object):
        def __init__(self, name, shape):
            self.name, self.shape = name, shape

    def __init__(self, path, sess_options=None, providers=None):
        self.graph = TorchGraph(path)
        self.path = path

    def get_inputs(self):
        return [self._Arg(i["name"], i["shape"]) for i in self.graph.inputs]

    def get_outputs(self):
        return [self._Arg(o["name"], o["shape"]) for o in self.graph.outputs]

    def run(self, output_names, input_feed=None):
        return self.graph.run(output_names, input_feed)

"""ONNX protobuf wire writer.

Used by `onnxocr_b200.synth_graphs` to emit architecture-faithful PP-OCRv5
det/rec graphs with seeded weights: the reference's real blobs
(`onnxocr/models/ppocrv5/{det,rec}/*.onnx`) are absent from the mount
(`/root/reference/.MISSING_LARGE_BLOBS`).  Field numbers follow onnx.proto3.
"""
import struct
import numpy as np

_DTYPE_CODES = {np.dtype(np.float32): 1, np.dtype(np.uint8): 2, np.dtype(np.int8): 3,
                np.dtype(np.int32): 6, np.dtype(np.int64): 7, np.dtype(np.bool_): 9,
                np.dtype(np.float16): 10, np.dtype(np.float64): 11}

# ---------------------------------------------------------------- wire encode

def _enc_varint(v):
    if v < 0:
        v += 1 << 64
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _key(fno, wt):
    return _enc_varint((fno << 3) | wt)


def _ld(fno, payload):
    return _key(fno, 2) + _enc_varint(len(payload)) + payload


def _vi(fno, v):
    return _key(fno, 0) + _enc_varint(int(v))


def enc_tensor(arr, name=""):
    arr = np.ascontiguousarray(arr)
    out = b""
    for d in arr.shape:
        out += _vi(1, d)
    out += _vi(2, _DTYPE_CODES[arr.dtype])
    if name:
        out += _ld(8, name.encode())
    out += _ld(9, arr.tobytes())
    return out


def enc_attr(name, val):
    out = _ld(1, name.encode())
    if isinstance(val, float):
        out += _key(2, 5) + struct.pack("<f", val) + _vi(20, 1)
    elif isinstance(val, (int, np.integer)):
        out += _vi(3, val) + _vi(20, 2)
    elif isinstance(val, (bytes, str)):
        out += _ld(4, val.encode() if isinstance(val, str) else val) + _vi(20, 3)
    elif isinstance(val, np.ndarray):
        out += _ld(5, enc_tensor(val)) + _vi(20, 4)
    elif isinstance(val, (list, tuple)) and len(val) and isinstance(val[0], float):
        for f in val:
            out += _key(7, 5) + struct.pack("<f", f)
        out += _vi(20, 6)
    elif isinstance(val, (list, tuple)):
        for i in val:
            out += _vi(8, i)
        out += _vi(20, 7)
    else:
        raise TypeError("attr %s: %r" % (name, type(val)))
    return out


def enc_node(op_type, inputs, outputs, name="", **attrs):
    out = b""
    for i in inputs:
        out += _ld(1, i.encode())
    for o in outputs:
        out += _ld(2, o.encode())
    if name:
        out += _ld(3, name.encode())
    out += _ld(4, op_type.encode())
    for k, v in attrs.items():
        out += _ld(5, enc_attr(k, v))
    return out


def enc_value_info(name, shape, elem_type=1):
    dims = b""
    for d in shape:
        if isinstance(d, str):
            dims += _ld(1, _ld(2, d.encode()))
        else:
            dims += _ld(1, _vi(1, d))
    tensor_type = _vi(1, elem_type) + _ld(2, dims)
    return _ld(1, name.encode()) + _ld(2, _ld(1, tensor_type))


def enc_model(nodes, inputs, outputs, graph_name="graph", opset=11, producer="b200ocr-synth",
              initializers=()):
    g = b""
    for n in nodes:
        g += _ld(1, n)
    g += _ld(2, graph_name.encode())
    for name, arr in initializers:
        g += _ld(5, enc_tensor(arr, name))
    for vi in inputs:
        g += _ld(11, vi)
    for vi in outputs:
        g += _ld(12, vi)
    m = _vi(1, 7) + _ld(2, producer.encode()) + _ld(7, g)
    m += _ld(8, _ld(1, b"") + _vi(2, opset))
    return m

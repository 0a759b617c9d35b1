"""ONNX protobuf wire reader (TEST INFRASTRUCTURE — oracle side).

The container has neither `onnx` nor `onnxruntime`, so the oracle reads
ONNX files at the protobuf wire level (the writer used to synthesise the missing
det/rec graphs lives in onnxocr_b200/onnx_writer.py and is checked against this reader).  Only the fields that paddle2onnx
graphs use are handled (see SURVEY.md §7.1 step 0b for the verified field list):

  ModelProto   ir_version=1 producer_name=2 graph=7 opset_import=8
  GraphProto   node=1 name=2 initializer=5 input=11 output=12
  NodeProto    input=1 output=2 name=3 op_type=4 attribute=5
  Attribute    name=1 f=2 i=3 s=4 t=5 floats=7 ints=8 type=20
  TensorProto  dims=1 data_type=2 float_data=4 int32_data=5 int64_data=7 name=8 raw_data=9
  ValueInfo    name=1 type=2{tensor_type=1{elem_type=1 shape=2{dim=1{dim_value=1|dim_param=2}}}}

This module is imported only by tests/, bench.py's reference leg and the golden
generator.  The product path has its own C++ reader (onnxocr_b200/csrc/onnx_reader.cc).
"""
import struct
import numpy as np

# ---------------------------------------------------------------- wire decode

def _varint(buf, pos):
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not (b & 0x80):
            break
        shift += 7
    return result, pos


def _signed64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def _fields(buf):
    """Yield (field_number, wire_type, value) over a message body."""
    pos = 0
    n = len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = bytes(buf[pos:pos + 8]); pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]; pos += ln
        elif wt == 5:
            v = bytes(buf[pos:pos + 4]); pos += 4
        else:
            raise ValueError("unsupported wire type %d" % wt)
        yield fno, wt, v


def _packed_varints(v, wt):
    if wt == 0:
        return [_signed64(v)]
    out = []
    pos = 0
    while pos < len(v):
        x, pos = _varint(v, pos)
        out.append(_signed64(x))
    return out


_DTYPES = {1: np.float32, 2: np.uint8, 3: np.int8, 6: np.int32, 7: np.int64,
           9: np.bool_, 10: np.float16, 11: np.float64}
_DTYPE_CODES = {np.dtype(v): k for k, v in _DTYPES.items()}


def _tensor(buf):
    dims, dt, name, raw = [], 1, "", None
    floats, i32, i64 = [], [], []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            dims += _packed_varints(v, wt)
        elif fno == 2:
            dt = v
        elif fno == 4:
            floats += list(struct.unpack("<%df" % (len(v) // 4), bytes(v))) if wt == 2 else [struct.unpack("<f", v)[0]]
        elif fno == 5:
            i32 += _packed_varints(v, wt)
        elif fno == 7:
            i64 += _packed_varints(v, wt)
        elif fno == 8:
            name = bytes(v).decode()
        elif fno == 9:
            raw = bytes(v)
    dtype = _DTYPES[dt]
    if raw is not None:
        arr = np.frombuffer(raw, dtype=dtype).copy()
    elif dt == 1:
        arr = np.array(floats, dtype=np.float32)
    elif dt == 7:
        arr = np.array(i64, dtype=np.int64)
    else:
        arr = np.array(i32).astype(dtype)
    return name, arr.reshape(dims)


def _attribute(buf):
    name, val, typ = "", None, 0
    floats, ints = [], []
    f = i = s = t = None
    for fno, wt, v in _fields(buf):
        if fno == 1:
            name = bytes(v).decode()
        elif fno == 2:
            f = struct.unpack("<f", v)[0]
        elif fno == 3:
            i = _signed64(v)
        elif fno == 4:
            s = bytes(v)
        elif fno == 5:
            t = _tensor(v)[1]
        elif fno == 7:
            floats += list(struct.unpack("<%df" % (len(v) // 4), bytes(v))) if wt == 2 else [struct.unpack("<f", v)[0]]
        elif fno == 8:
            ints += _packed_varints(v, wt)
        elif fno == 20:
            typ = v
    if typ == 1 or (typ == 0 and f is not None):
        val = f
    elif typ == 2 or (typ == 0 and i is not None):
        val = i
    elif typ == 3 or (typ == 0 and s is not None):
        val = s
    elif typ == 4 or (typ == 0 and t is not None):
        val = t
    elif typ == 6 or (typ == 0 and floats):
        val = floats
    elif typ == 7 or (typ == 0 and ints):
        val = ints
    else:
        val = ints if ints else floats
    return name, val


def _node(buf):
    n = {"input": [], "output": [], "name": "", "op_type": "", "attrs": {}}
    for fno, wt, v in _fields(buf):
        if fno == 1:
            n["input"].append(bytes(v).decode())
        elif fno == 2:
            n["output"].append(bytes(v).decode())
        elif fno == 3:
            n["name"] = bytes(v).decode()
        elif fno == 4:
            n["op_type"] = bytes(v).decode()
        elif fno == 5:
            k, a = _attribute(v)
            n["attrs"][k] = a
    return n


def _value_info(buf):
    name, shape, elem = "", [], 1
    for fno, wt, v in _fields(buf):
        if fno == 1:
            name = bytes(v).decode()
        elif fno == 2:
            for f2, _, v2 in _fields(v):
                if f2 != 1:
                    continue
                for f3, _, v3 in _fields(v2):
                    if f3 == 1:
                        elem = v3
                    elif f3 == 2:
                        for f4, _, v4 in _fields(v3):
                            if f4 != 1:
                                continue
                            d = None
                            for f5, w5, v5 in _fields(v4):
                                if f5 == 1:
                                    d = _signed64(v5)
                                elif f5 == 2:
                                    d = bytes(v5).decode()
                            shape.append(d)
    return {"name": name, "shape": shape, "elem_type": elem}


def load(path):
    """Parse an .onnx file → dict(nodes, initializers, inputs, outputs, opset, ir_version)."""
    with open(path, "rb") as fh:
        buf = memoryview(fh.read())
    model = {"nodes": [], "initializers": {}, "inputs": [], "outputs": [],
             "opset": {}, "ir_version": None, "producer": ""}
    for fno, wt, v in _fields(buf):
        if fno == 1:
            model["ir_version"] = v
        elif fno == 2:
            model["producer"] = bytes(v).decode()
        elif fno == 8:
            dom, ver = "", 0
            for f2, _, v2 in _fields(v):
                if f2 == 1:
                    dom = bytes(v2).decode()
                elif f2 == 2:
                    ver = v2
            model["opset"][dom] = ver
        elif fno == 7:
            for f2, _, v2 in _fields(v):
                if f2 == 1:
                    model["nodes"].append(_node(v2))
                elif f2 == 5:
                    k, a = _tensor(v2)
                    model["initializers"][k] = a
                elif f2 == 11:
                    model["inputs"].append(_value_info(v2))
                elif f2 == 12:
                    model["outputs"].append(_value_info(v2))
    init = model["initializers"]
    model["inputs"] = [i for i in model["inputs"] if i["name"] not in init]
    return model

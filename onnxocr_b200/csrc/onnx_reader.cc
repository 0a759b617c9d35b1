#include "onnx_reader.h"

#include <cstdio>
#include <cstring>
#include <stdexcept>

namespace b200ocr {
namespace {

struct Span {
    const uint8_t* p;
    const uint8_t* end;
};

uint64_t varint(Span& s) {
    uint64_t r = 0;
    int shift = 0;
    while (true) {
        if (s.p >= s.end) throw std::runtime_error("onnx: truncated varint");
        uint8_t b = *s.p++;
        r |= uint64_t(b & 0x7F) << shift;
        if (!(b & 0x80)) break;
        shift += 7;
        if (shift > 63) throw std::runtime_error("onnx: varint too long");
    }
    return r;
}

struct Field {
    uint32_t no;
    uint32_t wt;
    uint64_t v;      // wt 0
    Span bytes;      // wt 2 / 1 / 5
};

bool next(Span& s, Field& f) {
    if (s.p >= s.end) return false;
    uint64_t key = varint(s);
    f.no = uint32_t(key >> 3);
    f.wt = uint32_t(key & 7);
    if (f.wt == 0) {
        f.v = varint(s);
    } else if (f.wt == 2) {
        uint64_t n = varint(s);
        if (uint64_t(s.end - s.p) < n) throw std::runtime_error("onnx: truncated field");
        f.bytes = {s.p, s.p + n};
        s.p += n;
    } else if (f.wt == 5 || f.wt == 1) {
        size_t n = f.wt == 5 ? 4 : 8;
        if (size_t(s.end - s.p) < n) throw std::runtime_error("onnx: truncated fixed");
        f.bytes = {s.p, s.p + n};
        s.p += n;
    } else {
        throw std::runtime_error("onnx: unsupported wire type");
    }
    return true;
}

std::string str(const Span& b) { return std::string((const char*)b.p, b.end - b.p); }

void ints(const Field& f, std::vector<int64_t>& out) {
    if (f.wt == 0) {
        out.push_back((int64_t)f.v);
    } else {
        Span s = f.bytes;
        while (s.p < s.end) out.push_back((int64_t)varint(s));
    }
}

void floats(const Field& f, std::vector<float>& out) {
    size_t n = (f.bytes.end - f.bytes.p) / 4;
    size_t o = out.size();
    out.resize(o + n);
    memcpy(out.data() + o, f.bytes.p, n * 4);
}

HostTensor tensor(Span s, std::string* name) {
    HostTensor t;
    Span raw{nullptr, nullptr};
    std::vector<float> fd;
    std::vector<int64_t> i32, i64;
    Field f;
    while (next(s, f)) {
        switch (f.no) {
            case 1: ints(f, t.dims); break;
            case 2: t.dtype = (int)f.v; break;
            case 4: floats(f, fd); break;
            case 5: ints(f, i32); break;
            case 7: ints(f, i64); break;
            case 8: if (name) *name = str(f.bytes); break;
            case 9: raw = f.bytes; break;
            default: break;
        }
    }
    int64_t n = t.numel();
    if (t.dtype == 1) {
        if (raw.p) {
            if ((raw.end - raw.p) != n * 4) throw std::runtime_error("onnx: f32 raw_data size mismatch");
            t.f.resize(n);
            memcpy(t.f.data(), raw.p, n * 4);
        } else {
            t.f = fd;
        }
    } else if (t.dtype == 7) {
        if (raw.p) {
            t.i.resize(n);
            memcpy(t.i.data(), raw.p, n * 8);
        } else {
            t.i = i64;
        }
    } else if (t.dtype == 6) {
        if (raw.p) {
            t.i.resize(n);
            for (int64_t k = 0; k < n; ++k) {
                int32_t v;
                memcpy(&v, raw.p + 4 * k, 4);
                t.i[k] = v;
            }
        } else {
            t.i = i32;
        }
    } else if (t.dtype == 9 || t.dtype == 2 || t.dtype == 3) {
        t.i.resize(n);
        for (int64_t k = 0; k < n; ++k) t.i[k] = raw.p ? raw.p[k] : (k < (int64_t)i32.size() ? i32[k] : 0);
    } else {
        throw std::runtime_error("onnx: unsupported tensor data_type " + std::to_string(t.dtype));
    }
    return t;
}

Attr attribute(Span s, std::string* name) {
    Attr a;
    bool has_f = false, has_i = false, has_s = false, has_t = false;
    Field f;
    while (next(s, f)) {
        switch (f.no) {
            case 1: *name = str(f.bytes); break;
            case 2: memcpy(&a.f, f.bytes.p, 4); has_f = true; break;
            case 3: a.i = (int64_t)f.v; has_i = true; break;
            case 4: a.s = str(f.bytes); has_s = true; break;
            case 5: a.t = tensor(f.bytes, nullptr); has_t = true; break;
            case 7: if (f.wt == 2) floats(f, a.floats); else { float v; memcpy(&v, f.bytes.p, 4); a.floats.push_back(v); } break;
            case 8: ints(f, a.ints); break;
            case 20: a.type = (int)f.v; break;
            default: break;
        }
    }
    if (a.type == 0) a.type = has_f ? 1 : has_i ? 2 : has_s ? 3 : has_t ? 4 : !a.floats.empty() ? 6 : 7;
    return a;
}

OnnxNode node(Span s) {
    OnnxNode n;
    Field f;
    while (next(s, f)) {
        switch (f.no) {
            case 1: n.inputs.push_back(str(f.bytes)); break;
            case 2: n.outputs.push_back(str(f.bytes)); break;
            case 3: n.name = str(f.bytes); break;
            case 4: n.op_type = str(f.bytes); break;
            case 5: { std::string k; Attr a = attribute(f.bytes, &k); n.attrs[k] = a; break; }
            default: break;
        }
    }
    return n;
}

ValueInfo value_info(Span s) {
    ValueInfo v;
    Field f;
    while (next(s, f)) {
        if (f.no == 1) v.name = str(f.bytes);
        if (f.no != 2) continue;
        Span t = f.bytes;
        Field f2;
        while (next(t, f2)) {
            if (f2.no != 1) continue;          // TypeProto.tensor_type
            Span tt = f2.bytes;
            Field f3;
            while (next(tt, f3)) {
                if (f3.no != 2) continue;      // shape
                Span sh = f3.bytes;
                Field f4;
                while (next(sh, f4)) {
                    if (f4.no != 1) continue;  // dim
                    int64_t d = -1;
                    Span dm = f4.bytes;
                    Field f5;
                    while (next(dm, f5))
                        if (f5.no == 1) d = (int64_t)f5.v;
                    v.shape.push_back(d);
                }
            }
        }
    }
    return v;
}

}  // namespace

OnnxModel load_onnx(const std::string& path) {
    FILE* fh = fopen(path.c_str(), "rb");
    if (!fh) throw std::runtime_error("onnx: cannot open " + path);
    fseek(fh, 0, SEEK_END);
    long sz = ftell(fh);
    fseek(fh, 0, SEEK_SET);
    std::vector<uint8_t> buf(sz);
    if (sz > 0 && fread(buf.data(), 1, sz, fh) != (size_t)sz) {
        fclose(fh);
        throw std::runtime_error("onnx: short read " + path);
    }
    fclose(fh);
    OnnxModel m;
    Span s{buf.data(), buf.data() + buf.size()};
    Field f;
    bool have_graph = false;
    while (next(s, f)) {
        if (f.no == 1 && f.wt == 0) m.ir_version = (int64_t)f.v;
        if (f.no == 8 && f.wt == 2) {
            Span o = f.bytes;
            Field g;
            std::string dom;
            int64_t ver = 0;
            while (next(o, g)) {
                if (g.no == 1) dom = str(g.bytes);
                if (g.no == 2) ver = (int64_t)g.v;
            }
            if (dom.empty() || dom == "ai.onnx") m.opset = ver;
        }
        if (f.no == 7 && f.wt == 2) {
            have_graph = true;
            Span gs = f.bytes;
            Field g;
            while (next(gs, g)) {
                if (g.wt != 2) continue;
                if (g.no == 1) m.nodes.push_back(node(g.bytes));
                else if (g.no == 5) { std::string nm; HostTensor t = tensor(g.bytes, &nm); m.initializers[nm] = t; }
                else if (g.no == 11) m.inputs.push_back(value_info(g.bytes));
                else if (g.no == 12) m.outputs.push_back(value_info(g.bytes));
            }
        }
    }
    if (!have_graph) throw std::runtime_error("onnx: no graph in " + path);
    std::vector<ValueInfo> real_inputs;
    for (auto& vi : m.inputs)
        if (!m.initializers.count(vi.name)) real_inputs.push_back(vi);
    m.inputs = real_inputs;
    return m;
}

}  // namespace b200ocr

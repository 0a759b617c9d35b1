// ONNX protobuf wire reader (no protobuf / onnx dependency).
//
// Replaces what onnxruntime.InferenceSession(model_path, ...) does when it opens the
// file at /root/reference/onnxocr/predict_base.py:14.  Only the fields paddle2onnx
// emits are decoded (field numbers from onnx.proto3):
//   ModelProto  ir_version=1 graph=7 opset_import=8
//   GraphProto  node=1 initializer=5 input=11 output=12
//   NodeProto   input=1 output=2 name=3 op_type=4 attribute=5
//   Attribute   name=1 f=2 i=3 s=4 t=5 floats=7 ints=8 type=20
//   TensorProto dims=1 data_type=2 float_data=4 int32_data=5 int64_data=7 name=8 raw_data=9
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace b200ocr {

struct HostTensor {
    std::vector<int64_t> dims;
    int dtype = 1;                 // ONNX TensorProto.DataType: 1 f32, 6 i32, 7 i64
    std::vector<float> f;          // valid when dtype == 1
    std::vector<int64_t> i;        // valid when dtype is integral
    int64_t numel() const {
        int64_t n = 1;
        for (auto d : dims) n *= d;
        return n;
    }
    bool is_float() const { return dtype == 1; }
};

struct Attr {
    int type = 0;                  // 1 f, 2 i, 3 s, 4 t, 6 floats, 7 ints
    float f = 0.f;
    int64_t i = 0;
    std::string s;
    HostTensor t;
    std::vector<float> floats;
    std::vector<int64_t> ints;
};

struct OnnxNode {
    std::string name, op_type;
    std::vector<std::string> inputs, outputs;
    std::map<std::string, Attr> attrs;
    int64_t attr_i(const char* k, int64_t dflt) const {
        auto it = attrs.find(k);
        return it == attrs.end() ? dflt : it->second.i;
    }
    float attr_f(const char* k, float dflt) const {
        auto it = attrs.find(k);
        return it == attrs.end() ? dflt : it->second.f;
    }
    std::vector<int64_t> attr_ints(const char* k) const {
        auto it = attrs.find(k);
        return it == attrs.end() ? std::vector<int64_t>() : it->second.ints;
    }
    std::string attr_s(const char* k, const char* dflt) const {
        auto it = attrs.find(k);
        return it == attrs.end() ? std::string(dflt) : it->second.s;
    }
};

struct ValueInfo {
    std::string name;
    std::vector<int64_t> shape;    // -1 for dynamic dims
};

struct OnnxModel {
    int64_t ir_version = 0, opset = 0;
    std::vector<OnnxNode> nodes;
    std::map<std::string, HostTensor> initializers;
    std::vector<ValueInfo> inputs, outputs;
};

// Throws std::runtime_error with a description on malformed input.
OnnxModel load_onnx(const std::string& path);

}  // namespace b200ocr

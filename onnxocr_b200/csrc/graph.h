// Canonical graph: the paddle2onnx node soup rewritten into the operators the B200
// executor has kernels for.  Built once per model at load time (shape independent);
// engine.cu turns it into a launch plan per concrete input shape.
//
// Rewrites (patterns verified on the reference's shipped cls.onnx):
//   Constant nodes / initializers                      -> weight table
//   Conv + BatchNormalization                          -> Conv with folded scale/bias
//   Add(3) -> Clip(0,6) -> Mul(x) -> Div(6)            -> hardswish epilogue step
//   Sigmoid -> Mul(x)                                  -> swish epilogue step
//   Mul/Add by a 1-element constant (LearnableAffine)  -> affine epilogue step
//   ReduceMean/Sub/Pow/ReduceMean/Add/Sqrt/Div/Mul/Add -> LayerNorm
//   Reshape/Transpose/Slice/Squeeze/MatMul/Softmax/... -> Attention (SVTR global mixer)
//   GAP -> Conv -> Relu -> Conv -> HardSigmoid -> Mul  -> SE
//   MatMul(const) + Add(const)                         -> Linear (1x1 conv on [B,T,C])
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "onnx_reader.h"

namespace b200ocr {

enum EpiType : int {
    EPI_AFFINE = 1,     // v = v * a + b
    EPI_RELU = 2,
    EPI_HSWISH = 3,     // v * clamp(v + 3, 0, 6) / 6
    EPI_SIGMOID = 4,
    EPI_SWISH = 5,      // v * sigmoid(v)
    EPI_HSIGMOID = 6,   // clamp(v * a + b, 0, 1)
    EPI_CLIP = 7,       // clamp(v, a, b)
};

struct EpiStep {
    int type;
    float a, b;
};

constexpr int kMaxEpi = 8;
struct Epilogue {
    int n = 0;
    EpiStep s[kMaxEpi];
    // Compiled form: v = act(v * a0 + b0) * a1 + b1, valid when `fast` (covers every chain paddle2onnx emits for these
    // models: [LearnableAffine] -> activation -> [LearnableAffine]).  Kernels evaluate this instead of interpreting s[].
    int fast = 1, act = 0;
    float a0 = 1.f, b0 = 0.f, a1 = 1.f, b1 = 0.f, pa = 0.f, pb = 0.f;
    bool push(int type, float a = 0.f, float b = 0.f) {
        if (n >= kMaxEpi) return false;
        s[n++] = {type, a, b};
        compile();
        return true;
    }
    void compile() {
        fast = 1; act = 0; a0 = 1.f; b0 = 0.f; a1 = 1.f; b1 = 0.f; pa = 0.f; pb = 0.f;
        for (int i = 0; i < n; ++i) {
            const EpiStep& e = s[i];
            if (e.type == EPI_AFFINE) {
                if (act == 0) { a0 = a0 * e.a; b0 = b0 * e.a + e.b; }
                else { a1 = a1 * e.a; b1 = b1 * e.a + e.b; }
            } else if (act == 0) {
                act = e.type; pa = e.a; pb = e.b;
            } else {
                fast = 0;            // two non-affine steps: interpret
                return;
            }
        }
    }
};

enum CKind : int {
    C_CONV = 1,       // dense conv (groups == 1), any kernel/stride; also Linear (MatMul) as 1x1
    C_DWCONV,         // depthwise conv (groups == Cin == Cout)
    C_CONVT,          // ConvTranspose kernel 2x2 stride 2
    C_SE,             // y = x * hsigmoid(W2 relu(W1 gap(x) + b1) + b2)
    C_GAP,
    C_MAXPOOL,
    C_AVGPOOL,
    C_RESIZE,         // nearest, integer factors
    C_CONCAT,         // channel concat
    C_ADD,            // tensor + tensor (same shape or [N,C,1,1] broadcast)
    C_MUL,
    C_UNARY,          // stand-alone epilogue chain
    C_LAYERNORM,
    C_ATTENTION,
    C_SOFTMAX,        // over the channel (last logical) axis
    C_RESHAPE,
    C_TRANSPOSE,
    C_SQUEEZE,
    C_FLATTEN,
    C_CHANVEC,        // x (op) const per-channel vector, op in {add, mul}
    C_HOST,           // shape arithmetic evaluated on the host at plan time
};

struct CNode {
    int kind = 0;
    std::string name;                    // first ONNX node it came from (for error messages)
    std::vector<std::string> in;         // value names
    std::string out;
    Epilogue epi;
    // conv / dwconv / convT / linear
    int cin = 0, cout = 0, kh = 1, kw = 1, sh = 1, sw = 1, pt = 0, pl = 0, pb = 0, pr = 0, groups = 1;
    bool is_linear = false;              // came from MatMul: acts on the last logical axis
    int weight_id = -1;                  // index into CGraph::weights (packed, BN folded)
    // SE
    int se_w1 = -1, se_w2 = -1, se_mid = 0;
    // pool
    int ceil_mode = 0, count_include_pad = 0;
    // resize
    int fh = 1, fw = 1;
    // concat / softmax / squeeze / transpose
    int axis = 0;
    std::vector<int64_t> perm, axes;
    // layernorm
    float eps = 1e-5f;
    int ln_id = -1;
    // attention
    int heads = 0;
    float scale = 1.f;
    // chanvec
    int vec_id = -1;
    int vec_op = 0;                      // 0 add, 1 mul
    // host op: original ONNX node
    OnnxNode host;
};

// fp32 host copies; uploaded (and converted) by the engine.
struct ConvWeights {
    // layout [kh][kw][cin][cout]  (for depthwise: [kh][kw][c]; for convT: [dy][dx][cin][cout])
    std::vector<float> w;
    std::vector<float> bias;             // cout (zeros when absent)
    int cin = 0, cout = 0, kh = 1, kw = 1;
};

struct VecParam {
    std::vector<float> a, b;             // layernorm gamma/beta, or per-channel vector in `a`
};

struct CGraph {
    std::vector<CNode> nodes;
    std::vector<ConvWeights> weights;
    std::vector<VecParam> vecs;
    std::map<std::string, HostTensor> host_consts;   // small integer constants for shape math
    std::string input_name, output_name;
    std::vector<int64_t> input_shape, output_shape;  // -1 dynamic
    int64_t opset = 0;
    double param_count = 0;
};

// Throws std::runtime_error naming the offending op when a pattern is unsupported.
std::unique_ptr<CGraph> build_cgraph(const OnnxModel& m);

std::string dump_cgraph(const CGraph& g);

}  // namespace b200ocr

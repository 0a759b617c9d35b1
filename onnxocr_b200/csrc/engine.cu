#include "engine.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <type_traits>

#include "net_kernels.cuh"
#include "tc_conv.cuh"

namespace b200ocr {

std::vector<int64_t> PVal::logical() const {
    switch (form) {
        case F_NCHW: return {n, c, h, w};
        case F_NCS: return {n, c, (int64_t)h * w};
        case F_NSC: return {n, (int64_t)h * w, c};
        case F_NHWC: return {n, h, w, c};
        case F_NC: return {n, c};
        default: return host.empty() ? std::vector<int64_t>{} : std::vector<int64_t>{(int64_t)host.size()};
    }
}

Plan::~Plan() {
    if (own_arena) cudaFree(own_arena);
    if (out_dense) cudaFree(out_dense);
    if (graph_exec) cudaGraphExecDestroy(graph_exec);
}

static int grid_for(int64_t total, int block = 256) {
    int64_t g = (total + block - 1) / block;
    return (int)std::max<int64_t>(1, std::min<int64_t>(g, 148 * 32));
}

static float* upload(const std::vector<float>& v) {
    float* d = nullptr;
    B200_CUDA_CHECK(cudaMalloc(&d, std::max<size_t>(1, v.size()) * sizeof(float)));
    if (!v.empty()) B200_CUDA_CHECK(cudaMemcpy(d, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
    return d;
}

static int round_up(int a, int b) { return (a + b - 1) / b * b; }

// ------------------------------------------------------------------------------------ TMA descriptors
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        B200_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (!p || q != cudaDriverEntryPointSuccess) throw std::runtime_error("CUDA error: cuTensorMapEncodeTiled not available");
        fn = (EncodeTiledFn)p;
    }
    return fn;
}

// bf16 tensor map, 128-byte swizzle, zero fill out of bounds.  dims/box innermost first; strides in bytes for dims 1..rank-1.
static void make_tmap(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box) {
    cuuint64_t gd[4], gs[3];
    cuuint32_t bx[4], es[4];
    for (int i = 0; i < rank; ++i) {
        gd[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
        if (i) gs[i - 1] = strides_bytes[i - 1];
    }
    CUresult r = encode_tiled_fn()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw std::runtime_error("CUDA error: cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
}

// Depthwise strip kernel: NHWC activation as a 4-D map {CP, W, H, N}, box {box_c, box_w, box_h, 1}, no swizzle, zero fill.
void dw_encode_tmap(DwTmap* m, const void* base, int N, int H, int W, int CP, int box_c, int box_w, int box_h) {
    cuuint64_t gd[4] = {(cuuint64_t)CP, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t gs[3] = {(cuuint64_t)CP * 2, (cuuint64_t)W * CP * 2, (cuuint64_t)H * W * CP * 2};
    cuuint32_t bx[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = encode_tiled_fn()(&m->map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gd, gs, bx, es,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw std::runtime_error("CUDA error: cuTensorMapEncodeTiled (depthwise tile) failed with code " + std::to_string((int)r));
    m->base = base; m->box_c = box_c; m->box_w = box_w; m->box_h = box_h;
}

bool dw_tma_enabled() {                    // B200OCR_DW_TMA=0: cp.async tile load (A/B measurements)
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("B200OCR_DW_TMA");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

static bool tc_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("B200OCR_NO_TC");
        v = (e && e[0] == '1') ? 0 : 1;
    }
    return v == 1;
}

static bool tc_ws_enabled() {             // B200OCR_TC=stream selects the A+B streaming kernel (A/B measurements)
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("B200OCR_TC");
        v = (e && std::string(e) == "stream") ? 0 : 1;
    }
    return v == 1;
}

static bool gap_se_fused() {              // B200OCR_NO_GAP_SE=1: separate gap_partial + se_fc launches (A/B measurements)
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("B200OCR_NO_GAP_SE");
        v = (e && e[0] == '1') ? 0 : 1;
    }
    return v == 1;
}

static bool pixel_pack_enabled() {        // B200OCR_NO_PIXEL_PACK=1 disables the packed narrow-layer GEMM (A/B measurements)
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("B200OCR_NO_PIXEL_PACK");
        v = (e && e[0] == '1') ? 0 : 1;
    }
    return v == 1;
}

static bool se_lazy_enabled() {           // B200OCR_NO_SE_LAZY=1: always materialise x * gate (A/B measurements)
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("B200OCR_NO_SE_LAZY");
        v = (e && e[0] == '1') ? 0 : 1;
    }
    return v == 1;
}

static bool se_sums_enabled() {           // B200OCR_NO_SE_SUMS=1: separate GAP pass instead of producer-side channel sums
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("B200OCR_NO_SE_SUMS");
        v = (e && e[0] == '1') ? 0 : 1;
    }
    return v == 1;
}

static bool dw_strip_enabled() {          // B200OCR_DW=tile selects the previous depthwise kernel (A/B measurements)
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("B200OCR_DW");
        v = (e && std::string(e) == "tile") ? 0 : 1;
    }
    return v == 1;
}

struct TcState {            // per conv step: tensor maps, re-encoded when the activation slab moves
    CUtensorMap map_a, map_b;
    const void* a_ptr = nullptr;
};

// Two 1x1 convolutions in a row with nothing non-linear between them (and no other reader of the intermediate) are ONE
// linear map: W = W_b W_a, b = W_b b_a + b_b, merged in float64 at load time.  The PP-LCNetV3 det backbone ends every
// stage that way (a 1x1 "neck adapter" to 12..360 channels feeding RSEFPN's 1x1 ins_conv), so four launches and the
// narrow intermediates disappear.  Only the bf16 fast mode does this: the fp32 mode executes the graph node by node.
static int fuse_pointwise_chains(CGraph& g) {
    auto is_pw = [](const CNode& c) {
        return c.kind == C_CONV && !c.is_linear && c.kh == 1 && c.kw == 1 && c.sh == 1 && c.sw == 1 && c.pt == 0 && c.pl == 0 &&
               c.pb == 0 && c.pr == 0 && c.groups == 1;
    };
    int fused = 0;
    for (bool changed = true; changed;) {
        changed = false;
        std::map<std::string, int> nuse;
        for (auto& c : g.nodes)
            for (auto& in : c.in) nuse[in]++;
        nuse[g.output_name]++;
        for (size_t i = 0; i < g.nodes.size() && !changed; ++i) {
            const CNode& a = g.nodes[i];
            if (!is_pw(a) || a.epi.n != 0 || nuse[a.out] != 1) continue;
            for (size_t j = i + 1; j < g.nodes.size(); ++j) {
                CNode& b = g.nodes[j];
                if (b.in.empty() || b.in[0] != a.out) continue;
                if (!is_pw(b) || b.cin != a.cout) break;
                const ConvWeights &wa = g.weights[a.weight_id], &wb = g.weights[b.weight_id];
                ConvWeights wm;
                wm.cin = a.cin; wm.cout = b.cout; wm.kh = wm.kw = 1;
                wm.w.assign((size_t)a.cin * b.cout, 0.f);
                wm.bias.assign(b.cout, 0.f);
                for (int co = 0; co < b.cout; ++co) {
                    for (int ci = 0; ci < a.cin; ++ci) {
                        double acc = 0;
                        for (int m = 0; m < a.cout; ++m) acc += (double)wa.w[(size_t)ci * a.cout + m] * (double)wb.w[(size_t)m * b.cout + co];
                        wm.w[(size_t)ci * b.cout + co] = (float)acc;
                    }
                    double bb = wb.bias[co];
                    for (int m = 0; m < a.cout; ++m) bb += (double)wa.bias[m] * (double)wb.w[(size_t)m * b.cout + co];
                    wm.bias[co] = (float)bb;
                }
                g.weights.push_back(std::move(wm));
                b.weight_id = (int)g.weights.size() - 1;
                b.in[0] = a.in[0];
                b.cin = a.cin;
                g.nodes.erase(g.nodes.begin() + i);
                ++fused;
                changed = true;
                break;
            }
        }
    }
    return fused;
}

static bool pw_fuse_enabled() {           // B200OCR_NO_PW_FUSE=1: keep consecutive 1x1 convolutions separate (A/B measurements)
    const char* e = getenv("B200OCR_NO_PW_FUSE");
    return !(e && e[0] == '1');
}

// ------------------------------------------------------------------------------------ load
Model::Model(const std::string& path, int precision) : prec_(precision) {
    OnnxModel m = load_onnx(path);
    g_ = build_cgraph(m);
    if (prec_ == PREC_BF16 && pw_fuse_enabled()) fuse_pointwise_chains(*g_);
    if (g_->input_shape.size() == 4 && g_->input_shape[1] > 0) in_c_ = (int)g_->input_shape[1];
    dw_.resize(g_->weights.size());
    // pre = (a0, b0) of the node's compiled epilogue: folded into the tensor-core operand pack (W * a0, bias * a0 + b0)
    auto pack_dense = [&](int wid, bool convt, float pre_a = 1.f, float pre_b = 0.f) {
        const ConvWeights& cw = g_->weights[wid];
        DevWeights& d = dw_[wid];
        int cinp = cpad_of(cw.cin), coutp = cpad_of(cw.cout);
        int taps = cw.kh * cw.kw;
        int cols = convt ? 4 * coutp : coutp;
        d.K = convt ? cinp : taps * cinp;
        d.Kpad = round_up(d.K, IG_BK);
        d.ColsW = round_up(cols, 64);
        std::vector<float> w((size_t)d.Kpad * d.ColsW, 0.f), bias(d.ColsW, 0.f);
        for (int t = 0; t < taps; ++t)
            for (int ci = 0; ci < cw.cin; ++ci)
                for (int co = 0; co < cw.cout; ++co) {
                    float v = cw.w[((size_t)t * cw.cin + ci) * cw.cout + co];
                    if (convt) w[(size_t)ci * d.ColsW + t * coutp + co] = v;
                    else w[((size_t)t * cinp + ci) * d.ColsW + co] = v;
                }
        for (int co = 0; co < cw.cout; ++co) bias[co] = cw.bias[co];
        d.w = upload(w);
        d.bias = upload(bias);
        if (convt && prec_ == PREC_BF16 && cinp % 8 == 0 && coutp % 8 == 0) {
            // transposed conv as one GEMM: column = tap * coutp + co, K = cin (single tap)
            int gcols = 4 * coutp;
            int bn = gcols <= 32 ? 32 : (gcols <= 64 ? 64 : 128);
            int kchunks = (cinp + 63) / 64;
            int rows = round_up(gcols, bn);
            size_t K = (size_t)kchunks * 64;
            std::vector<bf16> wb((size_t)rows * K, __float2bfloat16(0.f));
            std::vector<float> tb(rows + 128, 0.f);
            for (int t = 0; t < taps; ++t)
                for (int co = 0; co < cw.cout; ++co) {
                    for (int ci = 0; ci < cw.cin; ++ci)
                        wb[(size_t)(t * coutp + co) * K + ci] = __float2bfloat16(pre_a * cw.w[((size_t)t * cw.cin + ci) * cw.cout + co]);
                    tb[t * coutp + co] = cw.bias[co] * pre_a + pre_b;
                }
            B200_CUDA_CHECK(cudaMalloc(&d.w_bf16, wb.size() * sizeof(bf16)));
            B200_CUDA_CHECK(cudaMemcpy(d.w_bf16, wb.data(), wb.size() * sizeof(bf16), cudaMemcpyHostToDevice));
            d.tc_bias = upload(tb);
            d.tc_bn = bn;
            d.tc_rows = rows;
            d.tc_k = (int)K;
        }
        if (!convt && prec_ == PREC_BF16 && cinp % 8 == 0) {
            // tcgen05 operand pack: bf16 [rows = round_up(CoutP, BN)][K = taps * kchunks * 64], K-major, zero padded
            int bn = coutp <= 32 ? 32 : (coutp <= 64 ? 64 : 128);
            int kchunks = (cinp + 63) / 64;
            int rows = round_up(coutp, bn);
            size_t K = (size_t)taps * kchunks * 64;
            std::vector<bf16> wb((size_t)rows * K, __float2bfloat16(0.f));
            for (int t = 0; t < taps; ++t)
                for (int ci = 0; ci < cw.cin; ++ci)
                    for (int co = 0; co < cw.cout; ++co)
                        wb[(size_t)co * K + (size_t)t * kchunks * 64 + ci] = __float2bfloat16(pre_a * cw.w[((size_t)t * cw.cin + ci) * cw.cout + co]);
            B200_CUDA_CHECK(cudaMalloc(&d.w_bf16, wb.size() * sizeof(bf16)));
            B200_CUDA_CHECK(cudaMemcpy(d.w_bf16, wb.data(), wb.size() * sizeof(bf16), cudaMemcpyHostToDevice));
            std::vector<float> tb(rows + 128, 0.f);
            for (int co = 0; co < cw.cout; ++co) tb[co] = cw.bias[co] * pre_a + pre_b;
            d.tc_bias = upload(tb);
            if (taps == 1 && bn == 128 && kchunks <= 4) {             // candidate CTC head (Linear with >= 65 outputs)
                std::vector<float> hb(rows + 128, -1.0e30f);
                for (int co = 0; co < cw.cout; ++co) hb[co] = cw.bias[co];
                d.head_bias = upload(hb);
            }
            d.tc_bn = bn;
            d.tc_rows = rows;
            d.tc_k = (int)K;
            if (taps == 1 && (cinp == 16 || cinp == 32) && cw.cout == coutp && pixel_pack_enabled()) {
                const int P = 64 / cinp, cols = P * coutp;
                const int prow = round_up(cols, 256);
                std::vector<bf16> pw((size_t)prow * 64, __float2bfloat16(0.f));
                std::vector<float> pb(prow + 256, 0.f);
                for (int j = 0; j < P; ++j)
                    for (int co = 0; co < cw.cout; ++co) {
                        for (int ci = 0; ci < cw.cin; ++ci)
                            pw[(size_t)(j * coutp + co) * 64 + j * cinp + ci] = __float2bfloat16(pre_a * cw.w[(size_t)ci * cw.cout + co]);
                        pb[j * coutp + co] = cw.bias[co] * pre_a + pre_b;
                    }
                B200_CUDA_CHECK(cudaMalloc(&d.pk_w, pw.size() * sizeof(bf16)));
                B200_CUDA_CHECK(cudaMemcpy(d.pk_w, pw.data(), pw.size() * sizeof(bf16), cudaMemcpyHostToDevice));
                d.pk_bias = upload(pb);
                d.pk_P = P;
                d.pk_rows = prow;
            }
        }
    };
    for (auto& c : g_->nodes) {
        const bool fold = prec_ == PREC_BF16 && c.epi.n > 0 && c.epi.fast && (c.epi.a0 != 1.f || c.epi.b0 != 0.f);
        if (c.kind == C_CONV || c.kind == C_CONVT) {
            pack_dense(c.weight_id, c.kind == C_CONVT, fold ? c.epi.a0 : 1.f, fold ? c.epi.b0 : 0.f);
            dw_[c.weight_id].tc_prefolded = fold;
        }
        else if (c.kind == C_SE) {
            pack_dense(c.se_w1, false);
            pack_dense(c.se_w2, false);
        } else if (c.kind == C_DWCONV) {
            const ConvWeights& cw = g_->weights[c.weight_id];
            DevWeights& d = dw_[c.weight_id];
            int cp = cpad_of(cw.cout), taps = cw.kh * cw.kw;
            std::vector<float> w((size_t)taps * cp, 0.f), bias(cp, 0.f);
            for (int t = 0; t < taps; ++t)
                for (int ch = 0; ch < cw.cout; ++ch) w[(size_t)t * cp + ch] = cw.w[(size_t)t * cw.cout + ch];
            for (int ch = 0; ch < cw.cout; ++ch) bias[ch] = cw.bias[ch];
            d.w = upload(w);
            d.bias = upload(bias);
            d.ColsW = cp;
            if (prec_ == PREC_BF16) {
                std::vector<bf16> wb(w.size());
                for (size_t i = 0; i < w.size(); ++i) wb[i] = __float2bfloat16(w[i]);
                B200_CUDA_CHECK(cudaMalloc(&d.w_bf16, wb.size() * sizeof(bf16)));
                B200_CUDA_CHECK(cudaMemcpy(d.w_bf16, wb.data(), wb.size() * sizeof(bf16), cudaMemcpyHostToDevice));
            }
        }
    }
    for (auto& v : g_->vecs) {
        std::vector<float> a = v.a, b = v.b;
        a.resize(round_up((int)a.size(), 8), 0.f);
        b.resize(a.size(), 0.f);
        dvec_.push_back({upload(a), upload(b)});
    }
    // the uploads above went through the legacy stream from pageable memory; the engine's streams are non-blocking
    B200_CUDA_CHECK(cudaDeviceSynchronize());
}

Model::~Model() {
    plans_.clear();
    for (auto& d : dw_) {
        if (d.w) cudaFree(d.w);
        if (d.bias) cudaFree(d.bias);
        if (d.w_bf16) cudaFree(d.w_bf16);
        if (d.head_bias) cudaFree(d.head_bias);
        if (d.pk_w) cudaFree(d.pk_w);
        if (d.pk_bias) cudaFree(d.pk_bias);
        if (d.tc_bias) cudaFree(d.tc_bias);
    }
    for (auto& v : dvec_) {
        cudaFree(v.first);
        cudaFree(v.second);
    }
    if (stage_in_) cudaFree(stage_in_);
    if (arena_) cudaFree(arena_);
}

std::string Model::describe() const { return dump_cgraph(*g_); }

// ------------------------------------------------------------------------------------ host shape math
static std::vector<int64_t> host_eval(const OnnxNode& n, const std::vector<std::vector<int64_t>>& in) {
    const std::string& op = n.op_type;
    if (op == "Cast" || op == "Unsqueeze" || op == "Squeeze") return in.at(0);
    if (op == "Concat") {
        std::vector<int64_t> r;
        for (auto& v : in) r.insert(r.end(), v.begin(), v.end());
        return r;
    }
    if (op == "Slice") {
        const auto& x = in.at(0);
        int64_t st, en, step = 1;
        if (in.size() >= 3) {
            st = in[1].at(0);
            en = in[2].at(0);
            if (in.size() >= 5 && !in[4].empty()) step = in[4][0];
        } else {
            st = n.attr_ints("starts").at(0);
            en = n.attr_ints("ends").at(0);
        }
        int64_t d = (int64_t)x.size();
        if (st < 0) st += d;
        if (en < 0) en += d;
        st = std::max<int64_t>(0, std::min(st, d));
        en = std::max<int64_t>(0, std::min(en, d));
        std::vector<int64_t> r;
        for (int64_t i = st; i < en; i += step) r.push_back(x[i]);
        return r;
    }
    if (op == "Gather") {
        std::vector<int64_t> r;
        for (auto i : in.at(1)) r.push_back(in[0].at(i < 0 ? i + in[0].size() : i));
        return r;
    }
    if (op == "Add" || op == "Sub" || op == "Mul" || op == "Div") {
        const auto &a = in.at(0), &b = in.at(1);
        size_t len = std::max(a.size(), b.size());
        std::vector<int64_t> r(len);
        for (size_t i = 0; i < len; ++i) {
            int64_t x = a[a.size() == 1 ? 0 : i], y = b[b.size() == 1 ? 0 : i];
            r[i] = op == "Add" ? x + y : op == "Sub" ? x - y : op == "Mul" ? x * y : (y ? x / y : 0);
        }
        return r;
    }
    throw std::runtime_error("host op not supported: " + op);
}

// ------------------------------------------------------------------------------------ plan
namespace {
struct Arena {
    struct Blk { size_t off, size; };
    std::vector<Blk> free_;
    size_t end = 0;
    size_t alloc(size_t bytes) {
        bytes = (bytes + 255) / 256 * 256;
        for (size_t i = 0; i < free_.size(); ++i) {
            if (free_[i].size >= bytes) {
                size_t off = free_[i].off;
                if (free_[i].size == bytes) free_.erase(free_.begin() + i);
                else { free_[i].off += bytes; free_[i].size -= bytes; }
                return off;
            }
        }
        // extend: reuse a trailing free block if it touches the end
        if (!free_.empty() && free_.back().off + free_.back().size == end) {
            size_t off = free_.back().off;
            end = off + bytes;
            free_.pop_back();
            return off;
        }
        size_t off = end;
        end += bytes;
        return off;
    }
    void release(size_t off, size_t bytes) {
        bytes = (bytes + 255) / 256 * 256;
        Blk b{off, bytes};
        auto it = std::lower_bound(free_.begin(), free_.end(), b, [](const Blk& x, const Blk& y) { return x.off < y.off; });
        it = free_.insert(it, b);
        size_t i = it - free_.begin();
        if (i + 1 < free_.size() && free_[i].off + free_[i].size == free_[i + 1].off) {
            free_[i].size += free_[i + 1].size;
            free_.erase(free_.begin() + i + 1);
        }
        if (i > 0 && free_[i - 1].off + free_[i - 1].size == free_[i].off) {
            free_[i - 1].size += free_[i].size;
            free_.erase(free_.begin() + i);
        }
    }
};
}  // namespace

Plan* Model::plan_for(int n, int h, int w, bool head, int own_arena, int slot, bool fused_pre) {
    std::vector<int> key{n, h, w, head ? 1 : 0, own_arena, slot, fused_pre ? 1 : 0};
    auto it = plans_.find(key);
    if (it != plans_.end()) return it->second.get();
    if (plans_.size() >= 768) {                         // bounded cache (dynamic rec widths x batch sizes)
        // Evicted plans are parked for one generation instead of being destroyed: a caller may still hold a pointer
        // into one (the det lanes keep a plan's output map pending until its boxes are collected).
        B200_CUDA_CHECK(cudaDeviceSynchronize());
        retired_.clear();
        for (auto& kv : plans_) retired_.push_back(std::move(kv.second));
        plans_.clear();
        private_bytes_ = 0;
    }
    std::unique_ptr<Plan> p = build_plan(n, h, w, head, own_arena, fused_pre);
    Plan* raw = p.get();
    plans_[key] = std::move(p);
    return raw;
}

std::unique_ptr<Plan> Model::build_plan(int n, int h, int w, bool head, int own_arena, bool fused_pre) {
    std::unique_ptr<Plan> p(new Plan());
    p->head_mode = head;
    p->fused_pre = fused_pre;
    for (int k = 0; k < 8; ++k) p->pre_norm[k] = pre_norm_[k];
    p->model_arena = &arena_;
    p->n = n;
    p->h = h;
    p->w = w;
    if (prec_ == PREC_BF16) build_steps<bf16>(*p);
    else build_steps<float>(*p);
    // build_steps uploads small per-plan tables (fused-tail weights, SE ticket counters) with cudaMemcpy / cudaMemset on
    // the legacy stream.  Those are NOT ordered with the engine's non-blocking streams and may still be in flight when
    // they return (device memset; pageable H2D), so a plan's first launch could see unset counters.  Plans are built
    // rarely: drain the device once.
    B200_CUDA_CHECK(cudaDeviceSynchronize());
    constexpr size_t kPrivateMax = (size_t)192 << 20;
    // B200OCR_PRIVATE_BUDGET_MB overrides the private-slab budget (tests use a tiny one to force the shared-slab fallback)
    static const size_t kPrivateBudget = [] {
        const char* e = getenv("B200OCR_PRIVATE_BUDGET_MB");
        return e ? (size_t)atoll(e) << 20 : (size_t)24 << 30;
    }();
    if ((own_arena == 2 || (own_arena == 1 && p->arena_bytes <= kPrivateMax)) && private_bytes_ + p->arena_bytes <= kPrivateBudget) {
        B200_CUDA_CHECK(cudaMalloc(&p->own_arena, p->arena_bytes));
        p->model_arena = &p->own_arena;
        private_bytes_ += p->arena_bytes;
        return p;
    }
    if (p->arena_bytes > arena_cap_) {
        B200_CUDA_CHECK(cudaDeviceSynchronize());
        if (arena_) cudaFree(arena_);
        arena_ = nullptr;
        arena_cap_ = p->arena_bytes + p->arena_bytes / 8;
        B200_CUDA_CHECK(cudaMalloc(&arena_, arena_cap_));
        // pad lanes of every buffer are written by the producing kernels; no memset needed
    }
    return p;
}

static int ws_min_tiles() {
    static const int v = [] {
        const char* e = getenv("B200OCR_WS_MIN_TILES");
        int k = e ? atoi(e) : 1;            // measured: 4 costs 35 % single-batch latency and buys nothing at 8 streams
        return k >= 1 && k <= 64 ? k : 1;
    }();
    return v;
}

template <typename T>
void Model::build_steps(Plan& p) {
    const CGraph& g = *g_;
    std::map<std::string, int> vid;
    std::vector<int> buf_first, buf_last;
    std::vector<size_t>& buf_bytes = p.buf_bytes;
    int cur = 0;   // node index being planned

    auto new_buf = [&](size_t bytes) {
        buf_bytes.push_back(std::max<size_t>(bytes, 16));
        buf_first.push_back(cur);
        buf_last.push_back(cur);
        return (int)buf_bytes.size() - 1;
    };
    auto new_val = [&](int form, int vn, int vh, int vw, int vc) {
        PVal v;
        v.form = form; v.n = vn; v.h = vh; v.w = vw; v.c = vc; v.cp = cpad_of(vc);
        v.root = new_buf((size_t)v.elems() * sizeof(T));
        p.vals.push_back(v);
        return (int)p.vals.size() - 1;
    };
    auto alias_val = [&](int src, int form, int vh, int vw) {
        PVal v = p.vals[src];
        v.form = form; v.h = vh; v.w = vw;
        p.vals.push_back(v);
        return (int)p.vals.size() - 1;
    };
    std::function<void(int)> use = [&](int v) {
        const PVal& pv = p.vals[v];
        if (pv.up_src >= 0) { use(pv.up_src); return; }
        if (pv.root >= 0) buf_last[pv.root] = std::max(buf_last[pv.root], cur);
        if (pv.gate >= 0) buf_last[pv.gate] = std::max(buf_last[pv.gate], cur);
    };
    auto get = [&](const std::string& name, const CNode& c) -> int {
        auto it = vid.find(name);
        if (it == vid.end()) throw std::runtime_error("plan: value '" + name + "' used before definition at " + c.name);
        return it->second;
    };
    Plan* pp = &p;
    double cost_flops = 0, cost_bytes = 0;      // set by a handler right before its emit()
    const double esz = (double)sizeof(T);
    auto emit = [&](const std::string& name, std::function<void(cudaStream_t)> fn) {
        p.steps.push_back(std::move(fn));
        p.step_names.push_back(name);
        p.step_flops.push_back(cost_flops);
        p.step_bytes.push_back(cost_bytes);
        cost_flops = cost_bytes = 0;
    };
    auto vbytes = [&](int v) { return (double)p.vals[v].n * p.vals[v].h * p.vals[v].w * p.vals[v].c * esz; };
    // value without a pending SE gate: runs se_apply once per gated value (cached) when a consumer cannot apply it itself
    std::map<int, int> ungated;
    std::function<int(int)> plain = [&](int v) -> int {
        if (p.vals[v].up_src >= 0) {
            int src = p.vals[v].up_src, ps = plain(src);
            if (ps == src) return v;
            PVal nv = p.vals[v];
            nv.up_src = ps;
            p.vals.push_back(nv);
            return (int)p.vals.size() - 1;
        }
        if (p.vals[v].gate < 0) return v;
        auto it = ungated.find(p.vals[v].root * 4096 + p.vals[v].gate);
        if (it != ungated.end()) return alias_val(it->second, p.vals[v].form, p.vals[v].h, p.vals[v].w);
        const PVal gv = p.vals[v];
        int out = new_val(gv.form, gv.n, gv.h, gv.w, gv.c);
        use(v);
        const int HW = gv.h * gv.w, CP = gv.cp, C = gv.c, residual = gv.gate_res, gbuf = gv.gate;
        const int64_t total4 = (int64_t)gv.n * HW * (gv.cp / 4), total8 = (int64_t)gv.n * HW * (gv.cp / 8);
        const bool x8 = std::is_same<T, bf16>::value && gv.cp % 8 == 0 && total8 < ((int64_t)1 << 31);
        cost_bytes = 2.0 * (double)gv.n * HW * gv.c * esz;
        Epilogue none;
        emit("se_apply(materialise)", [=](cudaStream_t st) {
            if (x8) {
                launch_pdl(se_apply_bf16x8_kernel, grid_for(total8), 256, 0, st, (const bf16*)pp->ptr(v), (const float*)pp->bufptr(gbuf),
                           (bf16*)pp->ptr(out), (uint32_t)HW, CP, C, (uint32_t)total8, residual, none);
                return;
            }
            launch_pdl(se_apply_kernel<T>, grid_for(total4), 256, 0, st, (const T*)pp->ptr(v), (const float*)pp->bufptr(gbuf),
                       (T*)pp->ptr(out), (int64_t)HW, CP, C, total4, residual, none);
        });
        ungated[gv.root * 4096 + gv.gate] = out;
        return out;
    };
    // materialise a virtual upsample (and any pending gate)
    auto materialize = [&](int v) -> int {
        v = plain(v);
        if (p.vals[v].up_src < 0) return v;
        PVal pv = p.vals[v];
        int src = pv.up_src;
        int out = new_val(pv.form, pv.n, pv.h, pv.w, pv.c);
        use(src);
        int cp = pv.cp, fh = pv.fh, fw = pv.fw, N = pv.n, H = pv.h, W = pv.w;
        emit("upsample", [=](cudaStream_t st) {
            launch_pdl(concat_slice_kernel<T>, grid_for((int64_t)N * H * W * cp), 256, 0, st, 
                (const T*)pp->ptr(src), (T*)pp->ptr(out), N, H, W, cp, cp, cp, 0, fh, fw);
        });
        return out;
    };
    std::map<std::string, int> nuse;
    for (auto& c : g.nodes)
        for (auto& in : c.in) nuse[in]++;
    nuse[g.output_name]++;

    // graph input
    cur = -1;
    p.in_val = new_val(F_NCHW, p.n, p.h, p.w, in_c_);
    vid[g.input_name] = p.in_val;
    std::vector<char> skip(g.nodes.size(), 0);

    for (size_t ni = 0; ni < g.nodes.size(); ++ni) {
        if (skip[ni]) continue;
        const CNode& c = g.nodes[ni];
        cur = (int)ni;
        switch (c.kind) {
            case C_HOST: {
                std::vector<std::vector<int64_t>> ins;
                if (c.host.op_type == "Shape") {
                    ins.push_back(p.vals[get(c.in[0], c)].logical());
                    PVal v;
                    v.form = F_HOST;
                    v.host = ins[0];
                    p.vals.push_back(v);
                } else {
                    for (auto& in : c.in) {
                        if (in.empty()) { ins.push_back({}); continue; }
                        auto hc = g.host_consts.find(in);
                        if (hc != g.host_consts.end()) {
                            std::vector<int64_t> t;
                            if (hc->second.is_float()) for (float f : hc->second.f) t.push_back((int64_t)f);
                            else t = hc->second.i;
                            ins.push_back(t);
                        } else {
                            ins.push_back(p.vals[get(in, c)].host);
                        }
                    }
                    PVal v;
                    v.form = F_HOST;
                    v.host = host_eval(c.host, ins);
                    p.vals.push_back(v);
                }
                vid[c.out] = (int)p.vals.size() - 1;
                break;
            }
            case C_CONV:
            case C_CONVT: {
                int in = materialize(get(c.in[0], c));
                const PVal iv = p.vals[in];
                bool convt = c.kind == C_CONVT;
                if (c.is_linear) {
                    if (iv.form != F_NSC && iv.form != F_NC && iv.form != F_NHWC)
                        throw std::runtime_error("MatMul on a tensor whose last axis is not the channel axis: " + c.name);
                } else if (iv.form != F_NCHW) {
                    throw std::runtime_error("Conv input is not NCHW-shaped: " + c.name);
                }
                if (iv.c != c.cin)
                    throw std::runtime_error("channel mismatch at " + c.name + ": got " + std::to_string(iv.c) +
                                             " expected " + std::to_string(c.cin));
                int Ho, Wo;
                if (convt) { Ho = iv.h; Wo = iv.w; }
                else {
                    Ho = (iv.h + c.pt + c.pb - c.kh) / c.sh + 1;
                    Wo = (iv.w + c.pl + c.pr - c.kw) / c.sw + 1;
                }
                if (Ho <= 0 || Wo <= 0) throw std::runtime_error("empty conv output at " + c.name);
                int out = new_val(iv.form, iv.n, convt ? 2 * Ho : Ho, convt ? 2 * Wo : Wo, c.cout);
                use(in);
                const DevWeights& dw = dw_[c.weight_id];
                ConvP cp;
                cp.N = iv.n; cp.H = iv.h; cp.W = iv.w; cp.CinP = iv.cp;
                cp.Ho = Ho; cp.Wo = Wo; cp.Cout = c.cout; cp.CoutP = cpad_of(c.cout);
                cp.kh = convt ? 1 : c.kh; cp.kw = convt ? 1 : c.kw;
                cp.sh = convt ? 1 : c.sh; cp.sw = convt ? 1 : c.sw;
                cp.pt = convt ? 0 : c.pt; cp.pl = convt ? 0 : c.pl;
                cp.K = dw.K; cp.ColsW = dw.ColsW;
                cp.Cols = convt ? 4 * cp.CoutP : cp.CoutP;
                cp.convT = convt ? 1 : 0;
                cp.M = iv.n * Ho * Wo;
                cp.epi = c.epi;
                p.flops += 2.0 * cp.M * (double)c.cout * c.cin * (convt ? 4 : c.kh * c.kw);
                cost_flops = 2.0 * cp.M * (double)c.cout * c.cin * (convt ? 4 : c.kh * c.kw);
                cost_bytes = vbytes(in) + vbytes(out) + (double)c.cout * c.cin * (convt ? 4 : c.kh * c.kw) * esz;
                const float* wptr = dw.w;
                const float* bptr = dw.bias;
                int cols = cp.Cols;
                // ---- tensor-core path (bf16 storage, stride-1 "same" convolutions and linear layers)
                bool same = !convt && c.sh == 1 && c.sw == 1 && Ho == iv.h && Wo == iv.w;
                bool pointwise = same && c.kh == 1 && c.kw == 1;
                bool tc_ct = convt && cp.CoutP % 8 == 0;
                // ---- fused CTC head: Linear -> Softmax (graph output) -> argmax / max, logits never materialised
                if (p.head_mode && std::is_same<T, bf16>::value && tc_enabled() && c.is_linear && c.epi.n == 0 && dw.head_bias &&
                    iv.cp % 8 == 0 && (iv.cp + tc::BK - 1) / tc::BK <= tc::HEAD_MAX_KB && ni + 1 < g.nodes.size() &&
                    g.nodes[ni + 1].kind == C_SOFTMAX && g.nodes[ni + 1].in[0] == c.out && nuse[c.out] == 1 &&
                    g.nodes[ni + 1].out == g.output_name && (iv.form == F_NSC || iv.form == F_NHWC)) {
                    int rows = iv.n * iv.h * iv.w;
                    int n_tiles = dw.tc_rows / tc::HEAD_BN;
                    int m_tiles = (int)ceil_div(rows, tc::BM);
                    // the class range is always cut the same way (independent of the batch size): the merged sum of
                    // exponentials is then grouped identically for every batch a row may find itself in
                    int splits = std::max(1, std::min(16, n_tiles));
                    int per = (int)ceil_div(n_tiles, splits);
                    splits = (int)ceil_div(n_tiles, per);
                    int kblocks = (iv.cp + tc::BK - 1) / tc::BK;
                    int b_pmax = new_buf((size_t)splits * rows * 4), b_pidx = new_buf((size_t)splits * rows * 4),
                        b_psum = new_buf((size_t)splits * rows * 4);
                    p.head_idx_buf = new_buf((size_t)rows * 4);
                    p.head_prob_buf = new_buf((size_t)rows * 4);
                    for (int b : {b_pmax, b_pidx, b_psum, p.head_idx_buf, p.head_prob_buf}) buf_last[b] = (int)g.nodes.size();
                    p.head_rows = rows;
                    p.head_T = iv.h * iv.w;
                    std::shared_ptr<TcState> ts(new TcState());
                    {
                        uint64_t dims[2] = {(uint64_t)dw.tc_k, (uint64_t)dw.tc_rows};
                        uint64_t str[1] = {(uint64_t)dw.tc_k * 2};
                        uint32_t box[2] = {(uint32_t)tc::BK, (uint32_t)tc::HEAD_BN};
                        make_tmap(&ts->map_b, dw.w_bf16, 2, dims, str, box);
                    }
                    tc::HeadParams hp;
                    hp.rows = rows; hp.n_tiles = n_tiles; hp.tiles_per_split = per; hp.kblocks = kblocks;
                    const float* hbias = dw.head_bias;
                    int icp = iv.cp, ib_idx = p.head_idx_buf, ib_prob = p.head_prob_buf;
                    size_t smem = tc::head_smem_bytes(kblocks);
                    cost_flops = 2.0 * rows * (double)c.cout * c.cin;
                    cost_bytes = vbytes(in) + (double)c.cout * c.cin * esz + (double)rows * 8;
                    p.flops += cost_flops;
                    use(in);
                    emit("ctc_head " + c.name, [=](cudaStream_t st) {
                        const void* x = pp->ptr(in);
                        if (ts->a_ptr != x) {
                            uint64_t dims[2] = {(uint64_t)icp, (uint64_t)rows};
                            uint64_t str[1] = {(uint64_t)icp * 2};
                            uint32_t box[2] = {(uint32_t)tc::BK, (uint32_t)tc::BM};
                            make_tmap(&ts->map_a, x, 2, dims, str, box);
                            ts->a_ptr = x;
                        }
                        cudaFuncSetAttribute(tc::ctc_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                        launch_pdl(tc::ctc_head_kernel, dim3(m_tiles, splits), tc::kThreads, smem, st, 
                            ts->map_a, ts->map_b, hbias, (float*)pp->bufptr(b_pmax), (int*)pp->bufptr(b_pidx),
                            (float*)pp->bufptr(b_psum), hp);
                    });
                    emit("ctc_head_merge " + c.name, [=](cudaStream_t st) {
                        launch_pdl(tc::ctc_head_merge_kernel, (unsigned)ceil_div(rows, 128), 128, 0, st, 
                            (const float*)pp->bufptr(b_pmax), (const int*)pp->bufptr(b_pidx), (const float*)pp->bufptr(b_psum),
                            splits, rows, (int*)pp->bufptr(ib_idx), (float*)pp->bufptr(ib_prob));
                    });
                    PVal ov = iv;                           // logical [n, T, classes]; lives only as idx / prob
                    ov.c = c.cout; ov.cp = cpad_of(c.cout); ov.root = -1; ov.form = iv.form;
                    p.vals.push_back(ov);
                    vid[g.nodes[ni + 1].out] = (int)p.vals.size() - 1;
                    p.out_written_dense = true;
                    skip[ni + 1] = 1;
                    break;
                }
                if (std::is_same<T, bf16>::value && tc_enabled() && (same || tc_ct) && iv.cp % 8 == 0 && cp.CoutP % 8 == 0 && dw.w_bf16) {
                    tc::Params tp;
                    memset(&tp, 0, sizeof(tp));
                    tp.mode = (pointwise || tc_ct) ? 0 : 1;
                    tp.M = cp.M;
                    tp.N = iv.n; tp.H = iv.h; tp.W = iv.w;
                    tp.taps = c.kh * c.kw; tp.kw = c.kw; tp.pad_t = c.pt; tp.pad_l = c.pl;
                    tp.kchunks = (iv.cp + tc::BK - 1) / tc::BK;
                    tp.Cout = c.cout; tp.CoutP = cp.CoutP;
                    if (tc_ct) {
                        tp.taps = 1; tp.kw = 1; tp.pad_t = tp.pad_l = 0;
                        tp.M = iv.n * iv.h * iv.w;
                        tp.convT = 1; tp.ctC = cp.CoutP; tp.ctCout = c.cout;
                        tp.Cout = tp.CoutP = 4 * cp.CoutP;
                    }
                    tp.epi = c.epi;
                    if (dw.tc_prefolded) { tp.epi.a0 = 1.f; tp.epi.b0 = 0.f; }   // already inside the packed weights / bias
                    // DB head tail: ConvT -> ConvT(-> <= 1 channel) [-> activation] == graph output, fused into this epilogue
                    bool ct2 = false;
                    std::shared_ptr<float> ct2_w;
                    if (tc_ct && ni + 1 < g.nodes.size() && cp.CoutP <= tc::CT2_MAX_C && c.epi.fast) {
                        const CNode& nx = g.nodes[ni + 1];
                        if (nx.kind == C_CONVT && nx.in[0] == c.out && nuse[c.out] == 1 && nx.out == g.output_name && nx.cout == 1 &&
                            nx.cin == c.cout) {
                            const ConvWeights& w2 = g.weights[nx.weight_id];
                            std::vector<float> hw((size_t)4 * cp.CoutP + 1, 0.f);
                            for (int t2 = 0; t2 < 4; ++t2)
                                for (int ci = 0; ci < nx.cin; ++ci) hw[(size_t)t2 * cp.CoutP + ci] = w2.w[((size_t)t2 * w2.cin + ci) * w2.cout];
                            hw[(size_t)4 * cp.CoutP] = w2.bias[0];
                            float* dptr = nullptr;
                            B200_CUDA_CHECK(cudaMalloc(&dptr, hw.size() * sizeof(float)));
                            B200_CUDA_CHECK(cudaMemcpy(dptr, hw.data(), hw.size() * sizeof(float), cudaMemcpyHostToDevice));
                            ct2_w.reset(dptr, [](float* q) { cudaFree(q); });
                            ct2 = true;
                            tp.ct2 = 1;
                            tp.ct2_w = dptr;
                            tp.epi2 = nx.epi;
                            p.flops += 2.0 * (4.0 * cp.M) * 4.0 * nx.cin;
                            cost_flops += 2.0 * (4.0 * cp.M) * 4.0 * nx.cin;
                            cost_bytes = vbytes(in) + 16.0 * cp.M * 4.0 + (double)c.cout * c.cin * 4 * esz;
                        }
                    }
                    tp.TW = 128; tp.TH = 1;
                    int64_t tap_cover = 0;                   // pixels the per-tap kernel's tiles cover (incl. the empty part)
                    if (tp.mode == 1) {
                        int64_t best = -1;
                        for (int tw = 128; tw >= 8; tw >>= 1) {
                            int th = 128 / tw;
                            int64_t cover = (int64_t)ceil_div(iv.w, tw) * tw * ceil_div(iv.h, th) * th;
                            if (best < 0 || cover < best) { best = cover; tp.TW = tw; tp.TH = th; }
                        }
                        tap_cover = best;
                    }
                    tp.tiles_x = (int)ceil_div(iv.w, tp.TW);
                    tp.tiles_y = (int)ceil_div(iv.h, tp.TH);
                    int bn = dw.tc_bn;
                    // narrow pointwise layer: P pixels per GEMM row against the block-diagonal weight
                    const void* wb_ptr = dw.w_bf16;
                    int wb_rows = dw.tc_rows, wb_k = dw.tc_k;
                    const float* tbias = dw.tc_bias;
                    int icp = iv.cp;
                    // (decided per sample — H*W, not the batch total — so that a crop's result never depends on what it is batched with)
                    if (pointwise && !c.is_linear && dw.pk_w && dw.pk_P * iv.cp == 64 && (iv.h * iv.w) % dw.pk_P == 0 && tc_ws_enabled()) {
                        tp.M /= dw.pk_P;
                        tp.Cout = tp.CoutP = dw.pk_P * cp.CoutP;
                        tp.kchunks = 1;
                        icp = 64;
                        wb_ptr = dw.pk_w; wb_rows = dw.pk_rows; wb_k = 64;
                        tbias = dw.pk_bias;
                    }
                    tp.cinp = icp;
                    // weight-stationary kernel whenever the layer's weight slab fits in shared memory
                    const int nkb = tp.taps * tp.kchunks;
                    // k x k convolutions re-read the activation tile once per tap, which dominates: they keep the streaming
                    // kernel (2 CTAs per SM); 1x1 / linear / transposed convolutions run weight-stationary
                    // k x k: halo mode — one (16+kh-1) x (8+kw-1) pixel tile per 64-channel chunk in smem, taps as shifted
                    // descriptors.  Correct (tests/test_gpu_tc.py passes with B200OCR_TC_HALO=1) and it cuts the L2->SM
                    // traffic 6x, but measured SLOWER than the per-tap re-load kernel on these layers (0.132 vs 0.096 ms
                    // for 3x3 96->24 @ 240^2 x 8): with N = 32 the 72 small UMMAs per tile are issue-bound on the single
                    // MMA thread and the streaming kernel runs two CTAs per SM.  Mode 2 (default since round 2) interleaves the
                    // MMAs of two tiles' accumulators and issues them with running descriptors: 0.067 vs 0.077 ms on the
                    // same layer, det plan 2.18 vs 2.29 ms at 16 pages.  B200OCR_TC_HALO=0 selects the per-tap kernel.
                    static const int halo_on = [] { const char* e = getenv("B200OCR_TC_HALO"); return e ? atoi(e) : 2; }();
                    // The halo tile is a fixed 8 x 16 pixel block: on a flat map (the recogniser's [B, 1, 40] sequence feeds two 3x3
                    // convolutions) 15 of its 16 rows are empty, five times the tiles of the per-tap kernel's 128 x 1 strip.
                    const int64_t halo_cover = (int64_t)ceil_div(iv.w, tc::WS_HALO_TW) * tc::WS_HALO_TW * ceil_div(iv.h, tc::WS_HALO_TH) * tc::WS_HALO_TH;
                    const bool halo = tp.mode == 1 && halo_on && c.kh <= 5 && c.kw <= 5 && halo_cover <= 2 * tap_cover;
                    size_t stage_bytes = tc::WS_A_BYTES;
                    if (halo) {
                        stage_bytes = tc::ws_halo_stage_bytes(c.kh, c.kw);
                        tp.TW = tc::WS_HALO_TW; tp.TH = tc::WS_HALO_TH;
                        tp.tiles_x = (int)ceil_div(iv.w, tp.TW);
                        tp.tiles_y = (int)ceil_div(iv.h, tp.TH);
                        tp.kh = c.kh;
                    }
                    const int ws_bn = (tc_ws_enabled() && (tp.mode == 0 || halo)) ? tc::ws_pick_bn(tp.CoutP, nkb, stage_bytes) : 0;
                    if (ws_bn) {
                        bn = ws_bn;
                        tp.ws_bn = ws_bn;
                        tp.ws_stages = tc::ws_pick_stages(ws_bn, nkb, stage_bytes);
                        tp.ws_halo = halo ? halo_on : 0;   // 2: two tiles' MMAs interleaved (needs 2 * kchunks + 1 stages)
                        tp.ws_stage_bytes = (int)stage_bytes;
                    }
                    // an SE gate follows: the epilogue accumulates the per-image channel sums (no separate GAP pass)
                    int gap_buf = -1;
                    if (ws_bn && !tc_ct && c.epi.fast && se_sums_enabled() && ni + 1 < g.nodes.size() && g.nodes[ni + 1].kind == C_SE &&
                        g.nodes[ni + 1].in[0] == c.out && cp.CoutP <= kSeSumsThreads && cp.CoutP % 8 == 0) {
                        const size_t bytes = (size_t)iv.n * tc::kGapReplicas * cp.CoutP * sizeof(unsigned long long);
                        gap_buf = new_buf(bytes);
                        // never shared with another buffer: the sums must still be zero (left so by se_fc_sums_kernel) when
                        // the next launch starts, see zero_sums below
                        buf_first[gap_buf] = -1;
                        buf_last[gap_buf] = (int)g.nodes.size();
                        p.vals[out].sums_replicas = tc::kGapReplicas;
                        tp.gap = reinterpret_cast<unsigned long long*>(1);   // resolved at launch (the slab may move)
                        tp.gap_mod = cp.CoutP;
                        tp.gap_rows = tp.mode == 0 ? tp.M / iv.n : 0;
                        p.vals[out].sums = gap_buf;
                        const int gb = gap_buf;
                        const double keep_f = cost_flops, keep_b = cost_bytes;     // the work belongs to the conv step below
                        cost_flops = 0; cost_bytes = (double)bytes;
                        emit("zero_sums " + c.name, [=](cudaStream_t st) {
                            // se_fc_sums_kernel re-zeroes the sums it consumed: in a private slab (nobody else writes
                            // there) only the first launch needs the memset, the captured graph carries none
                            if (pp->own_arena && pp->sums_clean) return;
                            cudaMemsetAsync(pp->bufptr(gb), 0, bytes, st);
                        });
                        cost_flops = keep_f; cost_bytes = keep_b;
                    }
                    std::shared_ptr<TcState> ts(new TcState());
                    {
                        uint64_t dims[2] = {(uint64_t)wb_k, (uint64_t)wb_rows};
                        uint64_t str[1] = {(uint64_t)wb_k * 2};
                        uint32_t box[2] = {(uint32_t)tc::BK, (uint32_t)bn};
                        make_tmap(&ts->map_b, wb_ptr, 2, dims, str, box);
                    }
                    int iN = iv.n, iH = iv.h, iW = iv.w;
                    const tc::Params tp0 = tp;
                    emit(std::string(ct2 ? "tcconvT2 " : tc_ct ? "tcconvT " : c.is_linear ? "tclinear " : "tcconv ") + c.name, [=](cudaStream_t st) {
                        (void)ct2_w;                                    // keeps the second-stage weights alive with the plan
                        tc::Params tp = tp0;
                        tp.ct2_out = pp->out_dense;
                        if (gap_buf >= 0) tp.gap = (unsigned long long*)pp->bufptr(gap_buf);
                        const void* x = pp->ptr(in);
                        if (ts->a_ptr != x) {
                            if (tp.mode == 0) {
                                uint64_t dims[2] = {(uint64_t)icp, (uint64_t)tp.M};
                                uint64_t str[1] = {(uint64_t)icp * 2};
                                uint32_t box[2] = {(uint32_t)tc::BK, (uint32_t)tc::BM};
                                make_tmap(&ts->map_a, x, 2, dims, str, box);
                            } else {
                                uint64_t dims[4] = {(uint64_t)icp, (uint64_t)iW, (uint64_t)iH, (uint64_t)iN};
                                uint64_t str[3] = {(uint64_t)icp * 2, (uint64_t)iW * icp * 2, (uint64_t)iH * iW * icp * 2};
                                uint32_t box[4] = {(uint32_t)tc::BK, (uint32_t)tp.TW, (uint32_t)tp.TH, 1};
                                if (tp.ws_halo) { box[1] = (uint32_t)(tp.TW + tp.kw - 1); box[2] = (uint32_t)(tp.TH + tp.kh - 1); }
                                make_tmap(&ts->map_a, x, 4, dims, str, box);
                            }
                            ts->a_ptr = x;
                        }
                        int m_tiles = tp.mode == 0 ? (tp.gap ? tp.N * (int)ceil_div(tp.gap_rows, tc::BM) : (int)ceil_div(tp.M, tc::BM))
                                                   : tp.tiles_x * tp.tiles_y * tp.N;
                        if (ws_bn) {
                            int n_tiles = (int)ceil_div(tp.CoutP, ws_bn);
                            // A CTA of this kernel owns a whole SM (its weight slab and the activation ring fill the shared
                            // memory), so a small layer that spreads one tile to every SM locks out the kernels of every
                            // other stream (rec batches on the auxiliary streams, the other det lanes) for its whole
                            // latency-bound lifetime.  Give every CTA at least ws_min_tiles() tiles: the layer itself takes
                            // barely longer (the slab load and pipeline fill dominate) and the other SMs stay free.
                            int m_groups = std::max(1, std::min(std::max(1, m_tiles / ws_min_tiles()), 148 / n_tiles));
                            size_t smem = tc::ws_smem_bytes(ws_bn, nkb, tp.ws_stages, tp.ws_halo ? (size_t)tp.ws_stage_bytes : (size_t)tc::WS_A_BYTES);
                            cudaFuncSetAttribute(tc::conv_tc_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::WS_SMEM_LIMIT);
                            launch_pdl(tc::conv_tc_ws_kernel, dim3((unsigned)(n_tiles * m_groups)), tc::kWsThreads, smem, st, 
                                ts->map_a, ts->map_b, tbias, (bf16*)pp->ptr(out), tp);
                            return;
                        }
                        int total_tiles = m_tiles * (int)ceil_div(tp.CoutP, bn);
                        int per_sm = bn == 128 ? 1 : 2;                     // shared-memory footprint per CTA
                        dim3 grid((unsigned)std::min(total_tiles, 148 * per_sm));
                        bf16* yo = (bf16*)pp->ptr(out);
                        if (bn == 32) {
                            cudaFuncSetAttribute(tc::conv_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::smem_bytes<32>());
                            launch_pdl(tc::conv_tc_kernel<32>, grid, tc::kThreads, tc::smem_bytes<32>(), st, ts->map_a, ts->map_b, tbias, yo, tp);
                        } else if (bn == 64) {
                            cudaFuncSetAttribute(tc::conv_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::smem_bytes<64>());
                            launch_pdl(tc::conv_tc_kernel<64>, grid, tc::kThreads, tc::smem_bytes<64>(), st, ts->map_a, ts->map_b, tbias, yo, tp);
                        } else {
                            cudaFuncSetAttribute(tc::conv_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::smem_bytes<128>());
                            launch_pdl(tc::conv_tc_kernel<128>, grid, tc::kThreads, tc::smem_bytes<128>(), st, ts->map_a, ts->map_b, tbias, yo, tp);
                        }
                    });
                    if (ct2) {                                          // the dense fp32 graph output was written directly
                        const CNode& nx = g.nodes[ni + 1];
                        PVal ov = p.vals[out];
                        ov.h = 4 * iv.h; ov.w = 4 * iv.w; ov.c = 1; ov.cp = 1; ov.root = -1;
                        p.vals.push_back(ov);
                        vid[nx.out] = (int)p.vals.size() - 1;
                        p.out_written_dense = true;
                        skip[ni + 1] = 1;
                        break;
                    }
                    vid[c.out] = out;
                    break;
                }
                int w_ld = dw.ColsW;
                // ---- pre-processing fused into the stem (plans built with fused_pre): u8 pages -> stem output
                if (p.fused_pre && in == p.in_val && !convt && !c.is_linear && iv.c == 3 && iv.cp == 4 && c.cout == 16 &&
                    cp.CoutP == 16 && c.kh == 3 && c.kw == 3 && c.sh == 2 && c.sw == 2 && c.pt == 1 && c.pl == 1 &&
                    nuse[g.input_name] == 1) {
                    std::vector<float> hw((size_t)36 * dw.ColsW), hb(16);
                    B200_CUDA_CHECK(cudaMemcpy(hw.data(), dw.w, hw.size() * sizeof(float), cudaMemcpyDeviceToHost));
                    B200_CUDA_CHECK(cudaMemcpy(hb.data(), dw.bias, 16 * sizeof(float), cudaMemcpyDeviceToHost));
                    StemW sw;
                    for (int tap = 0; tap < 9; ++tap)
                        for (int ci = 0; ci < 3; ++ci)
                            for (int co = 0; co < 16; ++co) sw.w[tap * 3 + ci][co] = hw[(size_t)(tap * 4 + ci) * dw.ColsW + co];
                    for (int co = 0; co < 16; ++co) sw.b[co] = hb[co];
                    // the job table the stages overwrite before every launch; until then every image reads a plan-owned
                    // grey page at identity scale (profile_plan / time_plan replay the plan without a caller)
                    uint8_t* dummy = nullptr;
                    ResizeJob* jobs = nullptr;
                    B200_CUDA_CHECK(cudaMalloc(&dummy, (size_t)iv.h * iv.w * 3));
                    p.owned.emplace_back(dummy, [](void* q) { cudaFree(q); });
                    B200_CUDA_CHECK(cudaMemset(dummy, 0x80, (size_t)iv.h * iv.w * 3));
                    B200_CUDA_CHECK(cudaMalloc(&jobs, sizeof(ResizeJob) * iv.n));
                    p.owned.emplace_back(jobs, [](void* q) { cudaFree(q); });
                    std::vector<ResizeJob> hj(iv.n);
                    for (auto& j : hj) {
                        j.src.p = dummy; j.src.h = iv.h; j.src.w = iv.w; j.src.stride = iv.w * 3; j.src.ph = iv.h; j.src.pw = iv.w;
                        j.src.rot180 = 0; j.dh = iv.h; j.dw = iv.w; j.scale_y = 1.0; j.scale_x = 1.0;
                    }
                    B200_CUDA_CHECK(cudaMemcpy(jobs, hj.data(), sizeof(ResizeJob) * iv.n, cudaMemcpyHostToDevice));
                    p.pre_jobs = jobs;
                    const int tiles_x = (int)ceil_div(Wo, kPsTW), tiles_y = (int)ceil_div(Ho, kPsTH), tiles = tiles_x * tiles_y * iv.n;
                    cost_bytes = (double)iv.n * iv.h * iv.w * 3 + vbytes(out) + 27.0 * 16 * 4;
                    const int ih = iv.h, iw = iv.w;
                    // bf16 plans: the stem as a tensor-core GEMM (B200OCR_STEM_MMA=0 keeps the FFMA kernel, which is
                    // bit-identical to conv_stem_kernel)
                    static const bool stem_mma = [] { const char* e = getenv("B200OCR_STEM_MMA"); return !(e && e[0] == '0'); }();
                    if (std::is_same<T, bf16>::value && stem_mma) {
                        std::unique_ptr<StemMmaTab> ht(new StemMmaTab());
                        auto wk = [&](int k, int n) -> float {          // K index = tap * 4 + channel (BGR0)
                            return (k < 36 && (k & 3) < 3) ? sw.w[(k >> 2) * 3 + (k & 3)][n] : 0.f;
                        };
                        auto split = [&](float v, int part) -> uint32_t {
                            const bf16 hi = __float2bfloat16(v);
                            const bf16 out = part == 0 ? hi : __float2bfloat16(v - __bfloat162float(hi));
                            uint16_t u;
                            memcpy(&u, &out, 2);
                            return u;
                        };
                        for (int part = 0; part < 2; ++part)
                            for (int ks = 0; ks < 3; ++ks)
                                for (int jn = 0; jn < 2; ++jn)
                                    for (int ln = 0; ln < 32; ++ln)
                                        for (int r = 0; r < 2; ++r) {
                                            const int gq = ln >> 2, tq = ln & 3, n = 8 * jn + gq, k = 16 * ks + 8 * r + 2 * tq;
                                            ht->b[part][ks][jn][ln][r] = split(wk(k, n), part) | (split(wk(k + 1, n), part) << 16);
                                        }
                        for (int co = 0; co < 16; ++co) ht->bias[co] = sw.b[co];
                        StemMmaTab* dtab = nullptr;
                        B200_CUDA_CHECK(cudaMalloc(&dtab, sizeof(StemMmaTab)));
                        p.owned.emplace_back(dtab, [](void* q) { cudaFree(q); });
                        B200_CUDA_CHECK(cudaMemcpy(dtab, ht.get(), sizeof(StemMmaTab), cudaMemcpyHostToDevice));
                        const bool hsw = cp.epi.n > 0 && cp.epi.fast && cp.epi.act == EPI_HSWISH;
                        const bool affine_only = cp.epi.fast && cp.epi.act == 0;        // BN folded, no activation (LCNetV3 stem)
                        emit("pre+stem " + c.name, [=](cudaStream_t st) {
                            NormP nrm;
                            nrm.mul = pp->pre_norm[0];
                            for (int k = 0; k < 3; ++k) { nrm.sub[k] = pp->pre_norm[1 + k]; nrm.div[k] = pp->pre_norm[4 + k]; }
                            nrm.rec_style = (int)pp->pre_norm[7];
                            const unsigned grid = (unsigned)std::min(tiles, 148 * 8);
                            if (hsw)
                                launch_pdl(pre_stem_mma_kernel<EPI_HSWISH>, grid, 256, 0, st, (const ResizeJob*)pp->pre_jobs, ih, iw, nrm,
                                           (const StemMmaTab*)dtab, (bf16*)pp->ptr(out), cp, tiles_x, tiles_y, tiles);
                            else if (affine_only)
                                launch_pdl(pre_stem_mma_kernel<0>, grid, 256, 0, st, (const ResizeJob*)pp->pre_jobs, ih, iw, nrm,
                                           (const StemMmaTab*)dtab, (bf16*)pp->ptr(out), cp, tiles_x, tiles_y, tiles);
                            else
                                launch_pdl(pre_stem_mma_kernel<-1>, grid, 256, 0, st, (const ResizeJob*)pp->pre_jobs, ih, iw, nrm,
                                           (const StemMmaTab*)dtab, (bf16*)pp->ptr(out), cp, tiles_x, tiles_y, tiles);
                        });
                        vid[c.out] = out;
                        break;
                    }
                    emit("pre+stem " + c.name, [=](cudaStream_t st) {
                        NormP nrm;
                        nrm.mul = pp->pre_norm[0];
                        for (int k = 0; k < 3; ++k) { nrm.sub[k] = pp->pre_norm[1 + k]; nrm.div[k] = pp->pre_norm[4 + k]; }
                        nrm.rec_style = (int)pp->pre_norm[7];
                        launch_pdl(pre_stem_kernel<T>, (unsigned)std::min(tiles, 148 * 8), 256, 0, st, (const ResizeJob*)pp->pre_jobs,
                                   ih, iw, nrm, sw, (T*)pp->ptr(out), cp, tiles_x, tiles_y, tiles);
                    });
                    vid[c.out] = out;
                    break;
                }
                emit(std::string(convt ? "convT " : c.is_linear ? "linear " : "conv ") + c.name, [=](cudaStream_t st) {
                    const T* x = (const T*)pp->ptr(in);
                    T* y = (T*)pp->ptr(out);
                    if (convt && cp.Cout <= 4 && (cp.CoutP == 1 || cp.CoutP == 4) && cp.CinP % 4 == 0) {
                        size_t smem = (size_t)4 * cp.Cout * cp.CinP * sizeof(float);
                        launch_pdl(convt_small_kernel<T>, grid_for(cp.M), 256, smem, st, x, wptr, w_ld, bptr, y, cp);
                        return;
                    }
                    if (!convt && cp.CinP == 4 && cp.Cout <= 32 && cp.CoutP % 4 == 0) {
                        int co = cp.Cout <= 8 ? 8 : (cp.Cout <= 16 ? 16 : 32);
                        size_t smem = (size_t)cp.kh * cp.kw * 4 * co * sizeof(float);
                        int g = grid_for(cp.M, 128);
                        if (co == 8) launch_pdl(conv_stem_kernel<T, 8>, g, 128, smem, st, x, wptr, w_ld, bptr, y, cp);
                        else if (co == 16) launch_pdl(conv_stem_kernel<T, 16>, g, 128, smem, st, x, wptr, w_ld, bptr, y, cp);
                        else launch_pdl(conv_stem_kernel<T, 32>, g, 128, smem, st, x, wptr, w_ld, bptr, y, cp);
                        return;
                    }
                    if (cols <= 16) {
                        dim3 grid((unsigned)ceil_div(cp.M, 256), (unsigned)ceil_div(cols, 16));
                        launch_pdl(conv_igemm_kernel<T, 256, 16>, grid, 256, 0, st, x, wptr, bptr, y, cp);
                    } else if (cols <= 32) {
                        dim3 grid((unsigned)ceil_div(cp.M, 128), (unsigned)ceil_div(cols, 32));
                        launch_pdl(conv_igemm_kernel<T, 128, 32>, grid, 256, 0, st, x, wptr, bptr, y, cp);
                    } else {
                        dim3 grid((unsigned)ceil_div(cp.M, 64), (unsigned)ceil_div(cols, 64));
                        launch_pdl(conv_igemm_kernel<T, 64, 64>, grid, 256, 0, st, x, wptr, bptr, y, cp);
                    }
                });
                vid[c.out] = out;
                break;
            }
            case C_DWCONV: {
                int in = materialize(get(c.in[0], c));
                const PVal iv = p.vals[in];
                if (iv.form != F_NCHW || iv.c != c.cin) throw std::runtime_error("bad depthwise input at " + c.name);
                int Ho = (iv.h + c.pt + c.pb - c.kh) / c.sh + 1;
                int Wo = (iv.w + c.pl + c.pr - c.kw) / c.sw + 1;
                int out = new_val(F_NCHW, iv.n, Ho, Wo, c.cout);
                use(in);
                DwP dp;
                dp.N = iv.n; dp.H = iv.h; dp.W = iv.w; dp.CP = iv.cp; dp.C = iv.c; dp.Ho = Ho; dp.Wo = Wo;
                dp.kh = c.kh; dp.kw = c.kw; dp.sh = c.sh; dp.sw = c.sw; dp.pt = c.pt; dp.pl = c.pl;
                dp.epi = c.epi;
                p.flops += 2.0 * iv.n * Ho * Wo * (double)c.cout * c.kh * c.kw;
                const float* wptr = dw_[c.weight_id].w;
                const float* bptr = dw_[c.weight_id].bias;
                const bf16* wb16 = (const bf16*)dw_[c.weight_id].w_bf16;
                int64_t total = (int64_t)iv.n * Ho * Wo * (iv.cp / 4);
                cost_flops = 2.0 * iv.n * Ho * Wo * (double)c.cout * c.kh * c.kw;
                cost_bytes = vbytes(in) + vbytes(out);
                const bool strip_ok = std::is_same<T, bf16>::value && wb16 && dp.CP % 8 == 0 && dp.kh == dp.kw && (dp.kw == 3 || dp.kw == 5) &&
                                      (dp.sw == 1 || dp.sw == 2) && (dp.sh == 1 || dp.sh == 2) && dw_strip_enabled();
                int dw_gap_buf = -1;
                if (strip_ok && se_sums_enabled() && ni + 1 < g.nodes.size() && g.nodes[ni + 1].kind == C_SE && g.nodes[ni + 1].in[0] == c.out &&
                    iv.cp <= kSeSumsThreads) {
                    const size_t bytes = (size_t)iv.n * iv.cp * sizeof(unsigned long long);
                    dw_gap_buf = new_buf(bytes);
                    buf_first[dw_gap_buf] = -1;                  // whole-plan lifetime, see the tensor-core case
                    buf_last[dw_gap_buf] = (int)g.nodes.size();
                    p.vals[out].sums = dw_gap_buf;
                    const int gb = dw_gap_buf;
                    emit("zero_sums " + c.name, [=](cudaStream_t st) {
                            // se_fc_sums_kernel re-zeroes the sums it consumed: in a private slab (nobody else writes
                            // there) only the first launch needs the memset, the captured graph carries none
                            if (pp->own_arena && pp->sums_clean) return;
                            cudaMemsetAsync(pp->bufptr(gb), 0, bytes, st);
                        });
                }
                const DwP dp0 = dp;
                std::shared_ptr<DwTmap> dtm(new DwTmap());
                emit("dwconv " + c.name, [=](cudaStream_t st) {
                    DwP dp = dp0;
                    DwTmap* tmq = dtm.get();
                    if (dw_gap_buf >= 0) dp.gap = (unsigned long long*)pp->bufptr(dw_gap_buf);
                    if (std::is_same<T, bf16>::value && wb16 && dp.CP % 8 == 0 && dp.kh == dp.kw && (dp.kw == 3 || dp.kw == 5) &&
                        (dp.sw == 1 || dp.sw == 2) && (dp.sh == 1 || dp.sh == 2)) {
                        const bf16* xi = (const bf16*)pp->ptr(in);
                        bf16* yo = (bf16*)pp->ptr(out);
                        int key = dp.kw * 100 + dp.sh * 10 + dp.sw;
                        if (dw_strip_enabled()) {
                            switch (key) {
                                case 311: launch_dwconv_strip<3, 1, 1>(xi, wptr, bptr, yo, dp, st, tmq); break;
                                case 312: launch_dwconv_strip<3, 1, 2>(xi, wptr, bptr, yo, dp, st, tmq); break;
                                case 321: launch_dwconv_strip<3, 2, 1>(xi, wptr, bptr, yo, dp, st, tmq); break;
                                case 322: launch_dwconv_strip<3, 2, 2>(xi, wptr, bptr, yo, dp, st, tmq); break;
                                case 511: launch_dwconv_strip<5, 1, 1>(xi, wptr, bptr, yo, dp, st, tmq); break;
                                case 512: launch_dwconv_strip<5, 1, 2>(xi, wptr, bptr, yo, dp, st, tmq); break;
                                case 521: launch_dwconv_strip<5, 2, 1>(xi, wptr, bptr, yo, dp, st, tmq); break;
                                default: launch_dwconv_strip<5, 2, 2>(xi, wptr, bptr, yo, dp, st, tmq); break;
                            }
                            return;
                        }
                        switch (key) {
                            case 311: launch_dwconv_tile<3, 1, 1>(xi, wb16, bptr, yo, dp, st); break;
                            case 312: launch_dwconv_tile<3, 1, 2>(xi, wb16, bptr, yo, dp, st); break;
                            case 321: launch_dwconv_tile<3, 2, 1>(xi, wb16, bptr, yo, dp, st); break;
                            case 322: launch_dwconv_tile<3, 2, 2>(xi, wb16, bptr, yo, dp, st); break;
                            case 511: launch_dwconv_tile<5, 1, 1>(xi, wb16, bptr, yo, dp, st); break;
                            case 512: launch_dwconv_tile<5, 1, 2>(xi, wb16, bptr, yo, dp, st); break;
                            case 521: launch_dwconv_tile<5, 2, 1>(xi, wb16, bptr, yo, dp, st); break;
                            default: launch_dwconv_tile<5, 2, 2>(xi, wb16, bptr, yo, dp, st); break;
                        }
                        return;
                    }
                    launch_pdl(dwconv_kernel<T>, grid_for(total), 256, 0, st, (const T*)pp->ptr(in), wptr, bptr, (T*)pp->ptr(out), dp);
                });
                vid[c.out] = out;
                break;
            }
            case C_SE:
            case C_GAP: {
                int in = materialize(get(c.in[0], c));
                const PVal iv = p.vals[in];
                if (iv.form != F_NCHW) throw std::runtime_error("GlobalAveragePool input is not NCHW-shaped: " + c.name);
                if (iv.cp > 1024 || iv.cp < 4) throw std::runtime_error("GlobalAveragePool channel count unsupported: " + c.name);
                int HW = iv.h * iv.w;
                int chunks = std::max(1, std::min(64, HW / 256));   // depends on HW only: per-sample sums do not depend on the batch
                int partial = new_buf((size_t)iv.n * chunks * iv.cp * sizeof(float));
                int rows = std::max(1, 256 / (iv.cp / 4));
                size_t smem = (size_t)rows * iv.cp * sizeof(float);
                int N = iv.n, CP = iv.cp, C = iv.c;
                use(in);
                cost_bytes = vbytes(in);
                // the producer already accumulated the channel sums in its epilogue: no pass over the tensor at all
                const bool from_sums = c.kind == C_SE && iv.sums >= 0 && c.se_mid <= kSeSumsThreads && iv.cp <= kSeSumsThreads;
                const bool fuse_gap_se = c.kind == C_SE && iv.cp <= kGapSeThreads && c.se_mid <= kGapSeThreads && gap_se_fused();
                if (!fuse_gap_se && !from_sums)
                    emit("gap_partial " + c.name, [=](cudaStream_t st) {
                        launch_pdl(gap_partial_kernel<T>, dim3(chunks, N), 256, smem, st, (const T*)pp->ptr(in),
                                   (float*)pp->bufptr(partial), HW, CP, chunks);
                    });
                if (c.kind == C_GAP) {
                    int out = new_val(F_NCHW, iv.n, 1, 1, iv.c);
                    emit("gap_finish " + c.name, [=](cudaStream_t st) {
                        launch_pdl(gap_finish_kernel<T>, N, 256, 0, st, (const float*)pp->bufptr(partial), (T*)pp->ptr(out), HW, C, CP, chunks);
                    });
                    vid[c.out] = out;
                    break;
                }
                if (c.cout != iv.c) throw std::runtime_error("SE channel mismatch at " + c.name);
                int scale = new_buf((size_t)iv.n * iv.cp * sizeof(float));
                // y = x + x*s when the very next node adds the SE input back (RSELayer shortcut)
                int residual = 0;
                std::string out_name = c.out;
                Epilogue epi = c.epi;
                if (ni + 1 < g.nodes.size() && c.epi.n == 0) {
                    const CNode& nx = g.nodes[ni + 1];
                    if (nx.kind == C_ADD && nx.epi.n == 0 && nuse[c.out] == 1 &&
                        ((nx.in[0] == c.in[0] && nx.in[1] == c.out) || (nx.in[1] == c.in[0] && nx.in[0] == c.out))) {
                        residual = 1;
                        out_name = nx.out;
                        skip[ni + 1] = 1;
                    }
                }
                // The gate stays pending on the value (PVal::gate) when nothing has to be applied after the multiplication:
                // consumers that can apply it on the fly read the un-gated tensor, the others materialise it (plain()).
                const bool lazy = std::is_same<T, bf16>::value && iv.cp % 8 == 0 && epi.n == 0 && se_lazy_enabled();
                int out = lazy ? alias_val(in, iv.form, iv.h, iv.w) : new_val(F_NCHW, iv.n, iv.h, iv.w, iv.c);
                if (lazy) { p.vals[out].gate = scale; p.vals[out].gate_res = residual; p.vals[out].sums = -1; p.vals[out].sums_replicas = 1; }
                const DevWeights &w1 = dw_[c.se_w1], &w2 = dw_[c.se_w2];
                int mid = c.se_mid;
                float alpha = c.scale, beta = c.eps;
                size_t smem2 = (size_t)(iv.cp + mid + kSeThreads) * sizeof(float);
                if (from_sums) {
                    const int sums_buf = iv.sums;
                    buf_last[sums_buf] = std::max(buf_last[sums_buf], cur);
                    const int reps = iv.sums_replicas;
                    const size_t smem4 = (size_t)(iv.cp + mid + kSeSumsThreads) * sizeof(float);
                    cost_bytes = (double)iv.n * iv.cp * 8 + 2.0 * (double)iv.c * mid * 4;
                    emit("se_fc_sums " + c.name, [=](cudaStream_t st) {
                        launch_pdl(se_fc_sums_kernel, N, kSeSumsThreads, smem4, st, (unsigned long long*)pp->bufptr(sums_buf), reps, HW, C, CP, mid, w1.w, w1.ColsW, w1.bias, w2.w,
                                   w2.ColsW, w2.bias, alpha, beta, (float*)pp->bufptr(scale));
                    });
                } else if (fuse_gap_se) {
                    int* dt = nullptr;                              // one ticket counter per image, self-resetting
                    B200_CUDA_CHECK(cudaMalloc(&dt, (size_t)N * sizeof(int)));
                    B200_CUDA_CHECK(cudaMemset(dt, 0, (size_t)N * sizeof(int)));
                    std::shared_ptr<int> tickets(dt, [](int* q) { cudaFree(q); });
                    const int grows = std::max(1, kGapSeThreads / (iv.cp / 4));
                    const size_t smem3 = std::max((size_t)grows * iv.cp, (size_t)(iv.cp + mid + kGapSeThreads)) * sizeof(float);
                    cost_bytes = vbytes(in);
                    emit("gap_se " + c.name, [=](cudaStream_t st) {
                        launch_pdl(gap_se_kernel<T>, dim3(chunks, N), kGapSeThreads, smem3, st, (const T*)pp->ptr(in),
                                   (float*)pp->bufptr(partial), HW, C, CP, chunks, mid, w1.w, w1.ColsW, w1.bias, w2.w, w2.ColsW, w2.bias,
                                   alpha, beta, (float*)pp->bufptr(scale), tickets.get());
                    });
                } else
                emit("se_fc " + c.name, [=](cudaStream_t st) {
                    launch_pdl(se_fc_kernel, N, kSeThreads, smem2, st, (const float*)pp->bufptr(partial), chunks, HW, C, CP, mid, w1.w, w1.ColsW,
                                                         w1.bias, w2.w, w2.ColsW, w2.bias, alpha, beta, (float*)pp->bufptr(scale));
                });
                int64_t total4 = (int64_t)iv.n * HW * (iv.cp / 4);
                cost_bytes = vbytes(in) + vbytes(out);
                const int64_t total8 = (int64_t)iv.n * HW * (iv.cp / 8);
                const bool se_x8 = std::is_same<T, bf16>::value && iv.cp % 8 == 0 && total8 < ((int64_t)1 << 31);
                if (!lazy)
                emit("se_apply " + c.name, [=](cudaStream_t st) {
                    if (se_x8) {
                        launch_pdl(se_apply_bf16x8_kernel, grid_for(total8), 256, 0, st, (const bf16*)pp->ptr(in), (const float*)pp->bufptr(scale),
                                                                                  (bf16*)pp->ptr(out), (uint32_t)HW, CP, C,
                                                                                  (uint32_t)total8, residual, epi);
                        return;
                    }
                    launch_pdl(se_apply_kernel<T>, grid_for(total4), 256, 0, st, (const T*)pp->ptr(in), (const float*)pp->bufptr(scale),
                                                                          (T*)pp->ptr(out), HW, CP, C, total4, residual, epi);
                });
                p.flops += 4.0 * iv.n * (double)iv.c * mid;
                vid[out_name] = out;
                break;
            }
            case C_MAXPOOL:
            case C_AVGPOOL: {
                int in = materialize(get(c.in[0], c));
                const PVal iv = p.vals[in];
                if (iv.form != F_NCHW) throw std::runtime_error("pool input is not NCHW-shaped: " + c.name);
                auto odim = [&](int d, int k, int s, int p0, int p1) {
                    int num = d + p0 + p1 - k;
                    return (c.ceil_mode ? (num + s - 1) / s : num / s) + 1;
                };
                PoolP pl;
                pl.N = iv.n; pl.H = iv.h; pl.W = iv.w; pl.CP = iv.cp; pl.C = iv.c;
                pl.Ho = odim(iv.h, c.kh, c.sh, c.pt, c.pb);
                pl.Wo = odim(iv.w, c.kw, c.sw, c.pl, c.pr);
                pl.kh = c.kh; pl.kw = c.kw; pl.sh = c.sh; pl.sw = c.sw; pl.pt = c.pt; pl.pl = c.pl;
                pl.is_max = c.kind == C_MAXPOOL;
                pl.count_include_pad = c.count_include_pad;
                int out = new_val(F_NCHW, iv.n, pl.Ho, pl.Wo, iv.c);
                use(in);
                int64_t total = (int64_t)iv.n * pl.Ho * pl.Wo * iv.cp;
                cost_bytes = vbytes(in) + vbytes(out);
                static const bool pool_x8_on = [] { const char* e = getenv("B200OCR_POOL_X8"); return !(e && e[0] == '0'); }();   // A/B switch
                const bool pool_x8 = pool_x8_on && std::is_same<T, bf16>::value && iv.cp % 8 == 0 && total < ((int64_t)1 << 31) &&
                                     (int64_t)iv.n * iv.h * iv.w * iv.cp < ((int64_t)1 << 31);
                emit("pool " + c.name, [=](cudaStream_t st) {
                    if (pool_x8) {
                        launch_pdl(pool_bf16x8_kernel, grid_for(total / 8), 256, 0, st, (const bf16*)pp->ptr(in), (bf16*)pp->ptr(out), pl);
                        return;
                    }
                    launch_pdl(pool_kernel<T>, grid_for(total), 256, 0, st, (const T*)pp->ptr(in), (T*)pp->ptr(out), pl);
                });
                vid[c.out] = out;
                break;
            }
            case C_RESIZE: {
                int in = get(c.in[0], c);
                if (p.vals[in].up_src >= 0) in = materialize(in);   // nested upsample: the inner one becomes real
                PVal v = p.vals[in];                                // a pending SE gate stays with the source (up_src)
                if (v.form != F_NCHW) throw std::runtime_error("Resize input is not NCHW-shaped: " + c.name);
                v.up_src = in;
                v.root = -1;
                v.gate = -1; v.gate_res = 0; v.sums = -1;
                v.fh = c.fh; v.fw = c.fw;
                v.h *= c.fh; v.w *= c.fw;
                p.vals.push_back(v);
                vid[c.out] = (int)p.vals.size() - 1;
                break;
            }
            case C_CONCAT: {
                std::vector<int> ins;
                int ctot = 0;
                for (auto& nm : c.in) ins.push_back(get(nm, c));
                const PVal v0 = p.vals[ins[0]];
                int axis = c.axis < 0 ? c.axis + (int)v0.logical().size() : c.axis;
                bool chan_axis = (v0.form == F_NCHW && axis == 1) || (v0.form == F_NSC && axis == 2) ||
                                 (v0.form == F_NC && axis == 1) || (v0.form == F_NHWC && axis == 3);
                if (!chan_axis) throw std::runtime_error("Concat along a non-channel axis: " + c.name);
                for (int v : ins) {
                    const PVal& pv = p.vals[v];
                    if (pv.n != v0.n || pv.h != v0.h || pv.w != v0.w) throw std::runtime_error("Concat shape mismatch: " + c.name);
                    ctot += pv.c;
                }
                int out = new_val(v0.form, v0.n, v0.h, v0.w, ctot);
                int cdp = p.vals[out].cp;
                {
                    // bf16, every input a multiple of 8 channels: all inputs in one launch that writes whole rows, SE gates
                    // applied on the fly
                    bool multi = std::is_same<T, bf16>::value && (int)ins.size() <= kCatMaxIn && cdp % 8 == 0 && cdp / 8 <= kCatMaxChunks &&
                                 v0.h <= 65535 && v0.n <= 65535 &&
                                 (int64_t)v0.n * v0.h * v0.w * (cdp / 8) < ((int64_t)1 << 31);
                    auto log2i = [](int f) { int l = 0; while ((1 << l) < f) ++l; return (1 << l) == f ? l : -1; };
                    for (int v : ins) {
                        const PVal& q = p.vals[v];
                        multi = multi && q.c % 8 == 0 && q.cp == q.c && (q.up_src < 0 || (log2i(q.fh) >= 0 && log2i(q.fw) >= 0));
                    }
                    if (multi) {
                        CatP cp0;
                        memset(&cp0, 0, sizeof(cp0));
                        memset(cp0.kmap, 255, sizeof(cp0.kmap));
                        cp0.N = v0.n; cp0.H = v0.h; cp0.W = v0.w; cp0.CdP = cdp; cp0.nin = (int)ins.size();
                        std::vector<int> srcs, gates;
                        int off8 = 0;
                        double bytes = (double)v0.n * v0.h * v0.w * ctot * esz;
                        for (size_t k = 0; k < ins.size(); ++k) {
                            const PVal pv = p.vals[ins[k]];
                            const int src = pv.up_src >= 0 ? pv.up_src : ins[k];
                            use(ins[k]);
                            srcs.push_back(src);
                            gates.push_back(p.vals[src].gate);
                            cp0.fh[k] = pv.up_src >= 0 ? log2i(pv.fh) : 0;
                            cp0.fw[k] = pv.up_src >= 0 ? log2i(pv.fw) : 0;
                            cp0.CsP[k] = pv.cp;
                            cp0.g0[k] = off8;
                            for (int q = 0; q < pv.c / 8; ++q) cp0.kmap[off8 + q] = (unsigned char)k;
                            cp0.res[k] = p.vals[src].gate_res;
                            off8 += pv.c / 8;
                            bytes += vbytes(src);
                        }
                        const uint32_t total8 = (uint32_t)((int64_t)v0.n * v0.h * v0.w * (cdp / 8));
                        cost_bytes = bytes;
                        emit("concat " + c.name, [=](cudaStream_t st) {
                            CatP cp = cp0;
                            for (int k = 0; k < cp.nin; ++k) {
                                cp.src[k] = (const bf16*)pp->ptr(srcs[k]);
                                cp.scale[k] = gates[k] >= 0 ? (const float*)pp->bufptr(gates[k]) : nullptr;
                            }
                            (void)total8;
                            dim3 grid((unsigned)ceil_div((int64_t)cp.W * (cp.CdP / 8), 256), (unsigned)cp.H, (unsigned)cp.N);
                            launch_pdl(concat_multi_bf16_kernel, grid, 256, 0, st, (bf16*)pp->ptr(out), cp);
                        });
                        vid[c.out] = out;
                        break;
                    }
                }
                for (int& v : ins) v = plain(v);
                if (cdp != ctot) {
                    int64_t tot = p.vals[out].elems();
                    emit("concat_zero " + c.name, [=](cudaStream_t st) {
                        launch_pdl(fill_zero_kernel<T>, grid_for(tot), 256, 0, st, (T*)pp->ptr(out), tot);
                    });
                }
                int off = 0;
                for (int v : ins) {
                    const PVal pv = p.vals[v];
                    int src = pv.up_src >= 0 ? pv.up_src : v;
                    use(v);
                    int N = pv.n, H = pv.h, W = pv.w, Cs = pv.c, CsP = pv.cp, fh = pv.fh, fw = pv.fw, coff = off;
                    cost_bytes = vbytes(src) + (double)N * H * W * Cs * esz;
                    emit("concat " + c.name, [=](cudaStream_t st) {
                        constexpr int epv = 16 / (int)sizeof(T);
                        if (Cs % epv == 0 && CsP % epv == 0 && cdp % epv == 0 && coff % epv == 0)
                            launch_pdl(concat_slice_vec_kernel<T>, grid_for((int64_t)N * H * W * (Cs / epv)), 256, 0, st, 
                                (const T*)pp->ptr(src), (T*)pp->ptr(out), N, H, W, Cs, CsP, cdp, coff, fh, fw);
                        else
                            launch_pdl(concat_slice_kernel<T>, grid_for((int64_t)N * H * W * Cs), 256, 0, st, 
                                (const T*)pp->ptr(src), (T*)pp->ptr(out), N, H, W, Cs, CsP, cdp, coff, fh, fw);
                    });
                    off += pv.c;
                }
                vid[c.out] = out;
                break;
            }
            case C_ADD:
            case C_MUL: {
                int a = get(c.in[0], c), b = get(c.in[1], c);
                auto is_bcast = [&](int big, int small) {
                    const PVal &x = p.vals[big], &y = p.vals[small];
                    return y.h == 1 && y.w == 1 && (x.h != 1 || x.w != 1) && x.n == y.n && x.c == y.c;
                };
                if (p.vals[a].up_src >= 0 && p.vals[b].up_src < 0) std::swap(a, b);
                if (is_bcast(b, a)) std::swap(a, b);
                // pending SE gates are applied by the bf16x8 kernel itself; every other kernel gets plain operands
                const bool gate_ok = std::is_same<T, bf16>::value && c.kind == C_ADD && p.vals[a].up_src < 0 && p.vals[a].cp % 8 == 0 &&
                                     !is_bcast(a, b) &&
                                     (int64_t)p.vals[a].n * p.vals[a].h * p.vals[a].w * (p.vals[a].cp / 8) < ((int64_t)1 << 31);
                if (!gate_ok) { a = plain(a); b = plain(b); }
                if (p.vals[a].up_src >= 0) a = materialize(a);
                if (p.vals[b].up_src >= 0 && is_bcast(a, b)) b = materialize(b);
                const PVal av = p.vals[a], bv = p.vals[b];
                bool bc = is_bcast(a, b);
                if (!bc && (av.n != bv.n || av.h != bv.h || av.w != bv.w || av.c != bv.c))
                    throw std::runtime_error("element-wise shape mismatch at " + c.name);
                int out = new_val(av.form, av.n, av.h, av.w, av.c);
                int bsrc = bv.up_src >= 0 ? bv.up_src : b;
                use(a);
                use(b);
                EltP ep;
                ep.N = av.n; ep.H = av.h; ep.W = av.w; ep.CP = av.cp; ep.C = av.c;
                ep.fh = bv.up_src >= 0 ? bv.fh : 1;
                ep.fw = bv.up_src >= 0 ? bv.fw : 1;
                ep.b_bcast = bc;
                ep.op = c.kind == C_MUL;
                ep.epi = c.epi;
                if (av.cp < 4) {
                    if (bc || bv.up_src >= 0) throw std::runtime_error("scalar-channel element-wise with broadcast: " + c.name);
                    int64_t tot = av.elems();
                    emit("eltwise " + c.name, [=](cudaStream_t st) {
                        launch_pdl(eltwise_scalar_kernel<T>, grid_for(tot), 256, 0, st, (const T*)pp->ptr(a), (const T*)pp->ptr(bsrc),
                                                                                 (T*)pp->ptr(out), tot, ep.op, ep.epi);
                    });
                } else {
                    int64_t tot = (int64_t)av.n * av.h * av.w * (av.cp / 4);
                    cost_bytes = vbytes(a) + vbytes(bsrc) + vbytes(out);
                    const int64_t tot8 = (int64_t)av.n * av.h * av.w * (av.cp / 8);
                    const bool x8 = std::is_same<T, bf16>::value && av.cp % 8 == 0 && !bc && tot8 < ((int64_t)1 << 31);
                    const int gate_a = av.gate, gate_b = p.vals[bsrc].gate;
                    ep.res_a = av.gate_res;
                    ep.res_b = p.vals[bsrc].gate_res;
                    if ((gate_a >= 0 || gate_b >= 0) && !x8) throw std::runtime_error("internal: pending SE gate on a non-vector element-wise at " + c.name);
                    auto log2i = [](int f) { int l = 0; while ((1 << l) < f) ++l; return (1 << l) == f ? l : -1; };
                    const bool fpn = x8 && c.kind == C_ADD && c.epi.n == 0 && av.c == av.cp && log2i(ep.fh) >= 0 && log2i(ep.fw) >= 0 &&
                                     av.h <= 65535 && av.n <= 65535 && (ep.fh > 1 || ep.fw > 1 || gate_a >= 0 || gate_b >= 0);
                    const int bh = av.h / ep.fh, bw = av.w / ep.fw;
                    emit("eltwise " + c.name, [=](cudaStream_t st) {
                        if (fpn) {
                            FpnAddP fp;
                            fp.W = ep.W; fp.Hb = bh; fp.Wb = bw; fp.cg = ep.CP / 8; fp.sh = log2i(ep.fh); fp.sw = log2i(ep.fw);
                            fp.sa = gate_a >= 0 ? (const float*)pp->bufptr(gate_a) : nullptr;
                            fp.sb = gate_b >= 0 ? (const float*)pp->bufptr(gate_b) : nullptr;
                            fp.res_a = ep.res_a; fp.res_b = ep.res_b;
                            dim3 grid((unsigned)ceil_div((int64_t)ep.W * fp.cg, 256), (unsigned)ep.H, (unsigned)ep.N);
                            launch_pdl(fpn_add_bf16_kernel, grid, 256, 0, st, (const bf16*)pp->ptr(a), (const bf16*)pp->ptr(bsrc),
                                       (bf16*)pp->ptr(out), fp);
                            return;
                        }
                        if (x8) {
                            EltP e2 = ep;
                            e2.sa = gate_a >= 0 ? (const float*)pp->bufptr(gate_a) : nullptr;
                            e2.sb = gate_b >= 0 ? (const float*)pp->bufptr(gate_b) : nullptr;
                            launch_pdl(eltwise_bf16x8_kernel, grid_for(tot8), 256, 0, st, (const bf16*)pp->ptr(a), (const bf16*)pp->ptr(bsrc),
                                                                                  (bf16*)pp->ptr(out), e2, (uint32_t)tot8);
                            return;
                        }
                        launch_pdl(eltwise_kernel<T>, grid_for(tot), 256, 0, st, (const T*)pp->ptr(a), (const T*)pp->ptr(bsrc),
                                                                          (T*)pp->ptr(out), ep);
                    });
                }
                vid[c.out] = out;
                break;
            }
            case C_UNARY:
            case C_CHANVEC: {
                int in = materialize(get(c.in[0], c));
                const PVal iv = p.vals[in];
                if (c.kind == C_CHANVEC && c.cout != iv.c) throw std::runtime_error("per-channel vector length mismatch at " + c.name);
                int out = new_val(iv.form, iv.n, iv.h, iv.w, iv.c);
                use(in);
                const float* vec = c.kind == C_CHANVEC ? dvec_[c.vec_id].first : nullptr;
                int vop = c.kind == C_CHANVEC ? c.vec_op : -1;
                int64_t tot = iv.elems();
                int CP = iv.cp, C = iv.c;
                Epilogue epi = c.epi;
                cost_bytes = vbytes(in) + vbytes(out);
                emit("unary " + c.name, [=](cudaStream_t st) {
                    launch_pdl(unary_kernel<T>, grid_for(tot), 256, 0, st, (const T*)pp->ptr(in), (T*)pp->ptr(out), tot, CP, C, vec, vop, epi);
                });
                vid[c.out] = out;
                break;
            }
            case C_LAYERNORM: {
                int in = materialize(get(c.in[0], c));
                const PVal iv = p.vals[in];
                if (iv.form != F_NSC && iv.form != F_NC && iv.form != F_NHWC)
                    throw std::runtime_error("LayerNorm over a non-channel axis: " + c.name);
                if (iv.c != c.cout) throw std::runtime_error("LayerNorm width mismatch at " + c.name);
                int out = new_val(iv.form, iv.n, iv.h, iv.w, iv.c);
                use(in);
                int64_t rows = (int64_t)iv.n * iv.h * iv.w;
                int C = iv.c, CP = iv.cp;
                float eps = c.eps;
                const float* ga = dvec_[c.ln_id].first;
                const float* be = dvec_[c.ln_id].second;
                cost_bytes = vbytes(in) + vbytes(out);
                emit("layernorm " + c.name, [=](cudaStream_t st) {
                    launch_pdl(layernorm_kernel<T>, (unsigned)ceil_div(rows * 32, 256), 256, 0, st, (const T*)pp->ptr(in), (T*)pp->ptr(out),
                                                                                             ga, be, rows, C, CP, eps);
                });
                vid[c.out] = out;
                break;
            }
            case C_ATTENTION: {
                int in = materialize(get(c.in[0], c));
                const PVal iv = p.vals[in];
                if (iv.form != F_NSC || iv.c % (3 * c.heads) != 0) throw std::runtime_error("attention input layout unsupported: " + c.name);
                int dim = iv.c / 3, hd = dim / c.heads, Tn = iv.h * iv.w;
                int out = new_val(F_NSC, iv.n, iv.h, iv.w, dim);
                use(in);
                size_t smem = (size_t)2 * Tn * hd * sizeof(float);
                if (smem > 200 * 1024) throw std::runtime_error("attention sequence too long for shared memory: " + c.name);
                int heads = c.heads, B = iv.n, CPin = iv.cp, CPout = p.vals[out].cp;
                float scale = c.scale;
                int qblocks = std::max(1, std::min(8, (Tn + 63) / 64));
                p.flops += 4.0 * B * heads * (double)Tn * Tn * hd;
                cost_flops = 4.0 * B * heads * (double)Tn * Tn * hd;
                cost_bytes = vbytes(in) + vbytes(out);
                emit("attention " + c.name, [=](cudaStream_t st) {
                    const T* q = (const T*)pp->ptr(in);
                    T* o = (T*)pp->ptr(out);
                    dim3 grid(B * heads, qblocks);
#define B200_ATTN(HD)                                                                                               \
    {                                                                                                               \
        if (smem > 48 * 1024)                                                                                       \
            cudaFuncSetAttribute(attention_kernel<T, HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        launch_pdl(attention_kernel<T, HD>, grid, 64, smem, st, q, o, Tn, heads, CPin, CPout, scale);                       \
    }
                    switch (hd) {
                        case 8: B200_ATTN(8) break;
                        case 15: B200_ATTN(15) break;
                        case 16: B200_ATTN(16) break;
                        case 30: B200_ATTN(30) break;
                        case 32: B200_ATTN(32) break;
                        default: break;
                    }
#undef B200_ATTN
                    if (CPout != dim) {
                        int64_t rows = (int64_t)B * Tn;
                        launch_pdl(zero_pad_lanes_kernel<T>, grid_for(rows * (CPout - dim)), 256, 0, st, o, rows, dim, CPout);
                    }
                });
                if (hd != 8 && hd != 15 && hd != 16 && hd != 30 && hd != 32)
                    throw std::runtime_error("attention head dim " + std::to_string(hd) + " unsupported: " + c.name);
                vid[c.out] = out;
                break;
            }
            case C_SOFTMAX: {
                int in = materialize(get(c.in[0], c));
                const PVal iv = p.vals[in];
                int rank = (int)iv.logical().size();
                int axis = c.axis < 0 ? c.axis + rank : c.axis;
                bool last_is_c = iv.form == F_NSC || iv.form == F_NC || iv.form == F_NHWC;
                if (!last_is_c || axis != rank - 1) throw std::runtime_error("Softmax over a non-channel axis: " + c.name);
                int64_t rows = (int64_t)iv.n * iv.h * iv.w;
                int C = iv.c, CP = iv.cp;
                use(in);
                if (c.out == g.output_name) {
                    // graph output: write the dense fp32 logical tensor directly
                    PVal ov = iv;
                    ov.root = -1;
                    p.vals.push_back(ov);
                    int out = (int)p.vals.size() - 1;
                    p.out_written_dense = true;
                    emit("softmax_out " + c.name, [=](cudaStream_t st) {
                        launch_pdl(softmax_kernel<T>, (unsigned)rows, 256, 0, st, (const T*)pp->ptr(in), nullptr, pp->out_dense, C, CP);
                    });
                    vid[c.out] = out;
                } else {
                    int out = new_val(iv.form, iv.n, iv.h, iv.w, iv.c);
                    emit("softmax " + c.name, [=](cudaStream_t st) {
                        launch_pdl(softmax_kernel<T>, (unsigned)rows, 256, 0, st, (const T*)pp->ptr(in), (T*)pp->ptr(out), nullptr, C, CP);
                    });
                    vid[c.out] = out;
                }
                break;
            }
            case C_RESHAPE: {
                int in = materialize(get(c.in[0], c));
                const PVal iv = p.vals[in];
                std::vector<int64_t> tgt;
                auto hc = g.host_consts.find(c.in[1]);
                if (hc != g.host_consts.end()) tgt = hc->second.i;
                else tgt = p.vals[get(c.in[1], c)].host;
                std::vector<int64_t> cur_shape = iv.logical();
                int64_t total = 1;
                for (auto d : cur_shape) total *= d;
                int64_t known = 1;
                int neg = -1;
                for (size_t i = 0; i < tgt.size(); ++i) {
                    if (tgt[i] == 0) tgt[i] = i < cur_shape.size() ? cur_shape[i] : 1;
                    if (tgt[i] == -1) neg = (int)i;
                    else known *= tgt[i];
                }
                if (neg >= 0) tgt[neg] = known ? total / known : 0;
                int out = -1;
                int64_t S = (int64_t)iv.h * iv.w;
                if (tgt == cur_shape) {
                    out = alias_val(in, iv.form, iv.h, iv.w);
                } else if (iv.form == F_NCHW && tgt.size() == 3 && tgt[0] == iv.n && tgt[1] == iv.c && tgt[2] == S) {
                    out = alias_val(in, F_NCS, iv.h, iv.w);
                } else if (iv.form == F_NCHW && tgt.size() == 2 && S == 1 && tgt[0] == iv.n && tgt[1] == iv.c) {
                    out = alias_val(in, F_NC, 1, 1);
                } else if (iv.form == F_NSC && tgt.size() == 4 && tgt[0] == iv.n && tgt[3] == iv.c && tgt[1] * tgt[2] == S) {
                    out = alias_val(in, F_NHWC, (int)tgt[1], (int)tgt[2]);
                } else if (iv.form == F_NHWC && tgt.size() == 3 && tgt[0] == iv.n && tgt[1] == S && tgt[2] == iv.c) {
                    out = alias_val(in, F_NSC, iv.h, iv.w);
                } else if (iv.form == F_NCS && tgt.size() == 4 && tgt[0] == iv.n && tgt[1] == iv.c && tgt[2] * tgt[3] == S) {
                    out = alias_val(in, F_NCHW, (int)tgt[2], (int)tgt[3]);
                } else if (iv.form == F_NC && tgt.size() == 4 && tgt[0] == iv.n && tgt[1] == iv.c && tgt[2] == 1 && tgt[3] == 1) {
                    out = alias_val(in, F_NCHW, 1, 1);
                } else {
                    std::ostringstream os;
                    os << "unsupported Reshape at " << c.name << ": [";
                    for (auto d : cur_shape) os << d << ",";
                    os << "] -> [";
                    for (auto d : tgt) os << d << ",";
                    os << "]";
                    throw std::runtime_error(os.str());
                }
                vid[c.out] = out;
                break;
            }
            case C_TRANSPOSE: {
                int in = materialize(get(c.in[0], c));
                const PVal iv = p.vals[in];
                const auto& pm = c.perm;
                int out = -1;
                auto is = [&](std::initializer_list<int64_t> l) { return pm == std::vector<int64_t>(l); };
                if (iv.form == F_NCS && is({0, 2, 1})) out = alias_val(in, F_NSC, iv.h, iv.w);
                else if (iv.form == F_NSC && is({0, 2, 1})) out = alias_val(in, F_NCS, iv.h, iv.w);
                else if (iv.form == F_NHWC && is({0, 3, 1, 2})) out = alias_val(in, F_NCHW, iv.h, iv.w);
                else if (iv.form == F_NCHW && is({0, 2, 3, 1})) out = alias_val(in, F_NHWC, iv.h, iv.w);
                else throw std::runtime_error("unsupported Transpose at " + c.name);
                vid[c.out] = out;
                break;
            }
            case C_SQUEEZE: {
                int in = materialize(get(c.in[0], c));
                const PVal iv = p.vals[in];
                int out = -1;
                std::vector<int64_t> ax = c.axes;
                for (auto& a : ax) if (a < 0) a += (int64_t)iv.logical().size();
                std::sort(ax.begin(), ax.end());
                if (iv.form == F_NCHW && ax == std::vector<int64_t>{2} && iv.h == 1) out = alias_val(in, F_NCS, 1, iv.w);
                else if (iv.form == F_NCHW && ax == std::vector<int64_t>{3} && iv.w == 1) out = alias_val(in, F_NCS, iv.h, 1);
                else if (iv.form == F_NCHW && ax == std::vector<int64_t>{2, 3} && iv.h == 1 && iv.w == 1) out = alias_val(in, F_NC, 1, 1);
                else throw std::runtime_error("unsupported Squeeze at " + c.name);
                vid[c.out] = out;
                break;
            }
            case C_FLATTEN: {
                int in = materialize(get(c.in[0], c));
                const PVal iv = p.vals[in];
                if (iv.form == F_NCHW && iv.h == 1 && iv.w == 1 && c.axis == 1) vid[c.out] = alias_val(in, F_NC, 1, 1);
                else throw std::runtime_error("unsupported Flatten at " + c.name);
                break;
            }
            default:
                throw std::runtime_error("plan: unhandled node kind at " + c.name);
        }
    }

    // graph output
    cur = (int)g.nodes.size();
    auto oit = vid.find(g.output_name);
    if (oit == vid.end()) throw std::runtime_error("graph output '" + g.output_name + "' was never produced");
    int ov = materialize(oit->second);
    p.out_val = ov;
    const PVal out = p.vals[ov];
    p.out_shape = out.logical();
    p.out_dense_elems = 1;
    for (auto d : p.out_shape) p.out_dense_elems *= (size_t)d;
    if (!p.out_written_dense) {
        use(ov);
        int to_nchw = out.form == F_NCHW || out.form == F_NCS;
        int N = out.n, C = out.c, H = out.h, W = out.w, CP = out.cp;
        int64_t tot = (int64_t)p.out_dense_elems;
        emit("to_dense", [=](cudaStream_t st) {
            launch_pdl(to_dense_f32_kernel<T>, grid_for(tot), 256, 0, st, (const T*)pp->ptr(ov), pp->out_dense, N, C, H, W, CP, to_nchw);
        });
    }
    if (p.head_rows == 0)
        B200_CUDA_CHECK(cudaMalloc(&p.out_dense, std::max<size_t>(1, p.out_dense_elems) * sizeof(float)));

    // ---- liveness-based offsets
    size_t nb = buf_bytes.size();
    p.buf_off.assign(nb, 0);
    std::vector<std::vector<int>> starts(g.nodes.size() + 2), ends(g.nodes.size() + 2);
    for (size_t b = 0; b < nb; ++b) {
        starts[buf_first[b] + 1].push_back((int)b);
        ends[buf_last[b] + 1].push_back((int)b);
    }
    Arena ar;
    for (size_t t = 0; t < starts.size(); ++t) {
        for (int b : starts[t]) p.buf_off[b] = ar.alloc(buf_bytes[b]);
        for (int b : ends[t]) ar.release(p.buf_off[b], buf_bytes[b]);
    }
    p.arena_bytes = ar.end;
    p.launches = (int)p.steps.size();
}

static bool graphs_enabled() {            // B200OCR_NO_GRAPH=1: launch every kernel individually (A/B measurements)
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("B200OCR_NO_GRAPH");
        v = (e && e[0] == '1') ? 0 : 1;
    }
    return v == 1;
}

// The first launch of a plan runs its kernels one by one (this also performs the lazy one-time work inside the step
// closures: tensor-map encoding, shared-memory attributes).  The second launch records the same sequence into a CUDA
// graph; from then on a plan is ONE cudaGraphLaunch, which removes the per-kernel launch latency that dominates the
// small rec / cls / single-page det batches.  A graph bakes in the activation-slab address, so it is re-recorded when
// the model's slab has been re-allocated (arena_version).
void Plan::launch(cudaStream_t st) {
    if (!graphs_enabled() || steps.size() < 8) {
        for (auto& s : steps) s(st);
        sums_clean = true;
        return;
    }
    if (graph_exec && graph_arena == *model_arena) {
        if (cudaGraphLaunch(graph_exec, st) == cudaSuccess) return;
        cudaGetLastError();
        cudaGraphExecDestroy(graph_exec);
        graph_exec = nullptr;
    }
    if (graph_exec) {
        cudaGraphExecDestroy(graph_exec);
        graph_exec = nullptr;
    }
    if (warm_launches < 1 || warm_arena != *model_arena) {
        for (auto& s : steps) s(st);
        warm_launches = 1;
        warm_arena = *model_arena;
        sums_clean = true;
        return;
    }
    cudaGraph_t g = nullptr;
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        for (auto& s : steps) s(st);
        return;
    }
    for (auto& s : steps) s(st);
    cudaError_t e = cudaStreamEndCapture(st, &g);
    if (e != cudaSuccess || !g) {
        cudaGetLastError();
        if (g) cudaGraphDestroy(g);
        for (auto& s : steps) s(st);              // capture was discarded: nothing has run yet
        return;
    }
    e = cudaGraphInstantiate(&graph_exec, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) {
        cudaGetLastError();
        graph_exec = nullptr;
        for (auto& s : steps) s(st);
        return;
    }
    graph_arena = *model_arena;
    B200_CUDA_CHECK(cudaGraphLaunch(graph_exec, st));
}

void Model::profile(int n, int h, int w, int iters, cudaStream_t st, std::vector<float>* ms) {
    std::lock_guard<std::mutex> lk(mu_);
    Plan* p = plan_for(n, h, w, false, 0, 0, fused_pre_default);
    size_t ns = p->steps.size();
    std::vector<cudaEvent_t> ev(ns + 1);
    for (auto& e : ev) B200_CUDA_CHECK(cudaEventCreate(&e));
    ms->assign(ns, 0.f);
    for (int it = 0; it < iters; ++it) {
        B200_CUDA_CHECK(cudaEventRecord(ev[0], st));
        for (size_t i = 0; i < ns; ++i) {
            p->steps[i](st);
            B200_CUDA_CHECK(cudaEventRecord(ev[i + 1], st));
        }
        B200_CUDA_CHECK(cudaStreamSynchronize(st));
        for (size_t i = 0; i < ns; ++i) {
            float t = 0;
            cudaEventElapsedTime(&t, ev[i], ev[i + 1]);
            (*ms)[i] += t / iters;
        }
    }
    for (auto& e : ev) cudaEventDestroy(e);
    B200_CUDA_CHECK(cudaGetLastError());
}

Plan* Model::launch_host_input(const float* nchw, int n, int c, int h, int w, cudaStream_t st, bool head) {
    if (c != in_c_) throw std::runtime_error("input has " + std::to_string(c) + " channels, model expects " + std::to_string(in_c_));
    Plan* p = plan_for(n, h, w, head);
    size_t in_bytes = (size_t)n * c * h * w * sizeof(float);
    if (in_bytes > stage_in_bytes_) {
        if (stage_in_) cudaFree(stage_in_);
        stage_in_ = nullptr;
        B200_CUDA_CHECK(cudaMalloc(&stage_in_, in_bytes));
        stage_in_bytes_ = in_bytes;
    }
    B200_CUDA_CHECK(cudaMemcpyAsync(stage_in_, nchw, in_bytes, cudaMemcpyHostToDevice, st));
    int cp = cpad_of(c);
    int64_t tot = (int64_t)n * h * w * cp;
    if (prec_ == PREC_BF16)
        launch_pdl(nchw_to_nhwc_kernel<bf16>, grid_for(tot), 256, 0, st, stage_in_, (bf16*)p->in_ptr(), n, c, h, w, cp);
    else
        launch_pdl(nchw_to_nhwc_kernel<float>, grid_for(tot), 256, 0, st, stage_in_, (float*)p->in_ptr(), n, c, h, w, cp);
    p->launch(st);
    B200_CUDA_CHECK(cudaGetLastError());
    return p;
}

void Model::run_host(const float* nchw, int n, int c, int h, int w, std::vector<int64_t>* out_shape, float* out,
                     size_t out_capacity, cudaStream_t st) {
    std::lock_guard<std::mutex> lk(mu_);
    Plan* p = launch_host_input(nchw, n, c, h, w, st, false);
    if (out_shape) *out_shape = p->out_shape;
    if (out) {
        if (p->out_dense_elems > out_capacity)
            throw std::runtime_error("output buffer too small: need " + std::to_string(p->out_dense_elems) + " floats");
        B200_CUDA_CHECK(cudaMemcpyAsync(out, p->out_dense, p->out_dense_elems * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    B200_CUDA_CHECK(cudaStreamSynchronize(st));
}

Plan* Model::run_device(const void* nhwc_in, int n, int h, int w, cudaStream_t st, bool want_dense) {
    (void)want_dense;
    Plan* p = plan_for(n, h, w);
    if (nhwc_in && nhwc_in != p->in_ptr())
        B200_CUDA_CHECK(cudaMemcpyAsync(p->in_ptr(), nhwc_in, (size_t)p->vals[p->in_val].elems() * elem_size(),
                                        cudaMemcpyDeviceToDevice, st));
    p->launch(st);
    B200_CUDA_CHECK(cudaGetLastError());
    return p;
}

}  // namespace b200ocr

#include "graph.h"

#include <algorithm>
#include <cmath>
#include <set>
#include <sstream>
#include <stdexcept>

namespace b200ocr {
namespace {

struct Builder {
    const OnnxModel& m;
    std::map<std::string, HostTensor> consts;
    std::vector<OnnxNode> nodes;           // non-Constant nodes, rewritten in place
    std::vector<char> dead;
    std::map<std::string, int> producer;
    std::map<std::string, std::vector<int>> consumers;
    std::set<std::string> graph_outputs;

    explicit Builder(const OnnxModel& model) : m(model) {
        for (auto& kv : m.initializers) consts[kv.first] = kv.second;
        for (auto& n : m.nodes) {
            if (n.op_type == "Constant") {
                auto it = n.attrs.find("value");
                if (it == n.attrs.end()) throw std::runtime_error("Constant node without tensor value: " + n.name);
                consts[n.outputs.at(0)] = it->second.t;
            } else {
                nodes.push_back(n);
            }
        }
        dead.assign(nodes.size(), 0);
        for (auto& o : m.outputs) graph_outputs.insert(o.name);
        reindex();
    }

    void reindex() {
        producer.clear();
        consumers.clear();
        for (size_t i = 0; i < nodes.size(); ++i) {
            if (dead[i]) continue;
            for (auto& o : nodes[i].outputs) producer[o] = (int)i;
            for (auto& in : nodes[i].inputs)
                if (!in.empty()) consumers[in].push_back((int)i);
        }
    }

    bool is_const(const std::string& v) const { return consts.count(v) != 0; }

    bool scalar(const std::string& v, float* out) const {
        auto it = consts.find(v);
        if (it == consts.end() || it->second.numel() != 1) return false;
        if (it->second.is_float()) *out = it->second.f[0];
        else *out = (float)it->second.i[0];
        return true;
    }

    int uses(const std::string& v) const {
        auto it = consumers.find(v);
        int n = it == consumers.end() ? 0 : (int)it->second.size();
        if (graph_outputs.count(v)) n += 1;
        return n;
    }

    // node that produces `v` if it is alive and has op_type `op`, else -1
    int prod(const std::string& v, const char* op) const {
        auto it = producer.find(v);
        if (it == producer.end() || dead[it->second]) return -1;
        return nodes[it->second].op_type == op ? it->second : -1;
    }

    // ---------------------------------------------------------------- rewrites
    void rewrite_hardswish() {
        for (size_t i = 0; i < nodes.size(); ++i) {
            if (dead[i] || nodes[i].op_type != "Div") continue;
            float six;
            if (!scalar(nodes[i].inputs[1], &six)) continue;
            int mul = prod(nodes[i].inputs[0], "Mul");
            if (mul < 0 || uses(nodes[mul].outputs[0]) != 1) continue;
            for (int side = 0; side < 2; ++side) {
                const std::string& x = nodes[mul].inputs[side];
                int clip = prod(nodes[mul].inputs[1 - side], "Clip");
                if (clip < 0 || uses(nodes[clip].outputs[0]) != 1) continue;
                float lo = 0.f, hi = 6.f;
                if (nodes[clip].inputs.size() >= 3) {
                    if (!scalar(nodes[clip].inputs[1], &lo) || !scalar(nodes[clip].inputs[2], &hi)) continue;
                } else {
                    lo = nodes[clip].attr_f("min", 0.f);
                    hi = nodes[clip].attr_f("max", 6.f);
                }
                int add = prod(nodes[clip].inputs[0], "Add");
                if (add < 0 || uses(nodes[add].outputs[0]) != 1) continue;
                float three;
                bool ok = (nodes[add].inputs[0] == x && scalar(nodes[add].inputs[1], &three)) ||
                          (nodes[add].inputs[1] == x && scalar(nodes[add].inputs[0], &three));
                if (!ok || three != 3.f || lo != 0.f || hi != 6.f || six != 6.f) continue;
                OnnxNode hs;
                hs.op_type = "HardSwish";
                hs.name = nodes[i].name.empty() ? nodes[i].outputs[0] : nodes[i].name;
                hs.inputs = {x};
                hs.outputs = nodes[i].outputs;
                nodes[i] = hs;
                dead[mul] = dead[clip] = dead[add] = 1;
                reindex();
                break;
            }
        }
    }

    void rewrite_swish() {
        for (size_t i = 0; i < nodes.size(); ++i) {
            if (dead[i] || nodes[i].op_type != "Mul") continue;
            for (int side = 0; side < 2; ++side) {
                int sg = prod(nodes[i].inputs[1 - side], "Sigmoid");
                if (sg < 0 || uses(nodes[sg].outputs[0]) != 1) continue;
                if (nodes[sg].inputs[0] != nodes[i].inputs[side]) continue;
                OnnxNode sw;
                sw.op_type = "Swish";
                sw.name = nodes[i].outputs[0];
                sw.inputs = {nodes[i].inputs[side]};
                sw.outputs = nodes[i].outputs;
                nodes[i] = sw;
                dead[sg] = 1;
                reindex();
                break;
            }
        }
    }

    void rewrite_layernorm() {
        for (size_t i = 0; i < nodes.size(); ++i) {
            if (dead[i] || nodes[i].op_type != "Add") continue;
            // Add(Mul(Div(d, Sqrt(Add(ReduceMean(Pow(d, 2)), eps))), gamma), beta),  d = Sub(x, ReduceMean(x))
            int beta_side = is_const(nodes[i].inputs[1]) ? 1 : is_const(nodes[i].inputs[0]) ? 0 : -1;
            if (beta_side < 0) continue;
            int mul = prod(nodes[i].inputs[1 - beta_side], "Mul");
            if (mul < 0 || uses(nodes[mul].outputs[0]) != 1) continue;
            int gamma_side = is_const(nodes[mul].inputs[1]) ? 1 : is_const(nodes[mul].inputs[0]) ? 0 : -1;
            if (gamma_side < 0) continue;
            int div = prod(nodes[mul].inputs[1 - gamma_side], "Div");
            if (div < 0 || uses(nodes[div].outputs[0]) != 1) continue;
            int sub = prod(nodes[div].inputs[0], "Sub");
            int sq = prod(nodes[div].inputs[1], "Sqrt");
            if (sub < 0 || sq < 0) continue;
            int addeps = prod(nodes[sq].inputs[0], "Add");
            if (addeps < 0) continue;
            float eps;
            int eps_side = scalar(nodes[addeps].inputs[1], &eps) ? 1 : scalar(nodes[addeps].inputs[0], &eps) ? 0 : -1;
            if (eps_side < 0) continue;
            int rm2 = prod(nodes[addeps].inputs[1 - eps_side], "ReduceMean");
            if (rm2 < 0) continue;
            int pw = prod(nodes[rm2].inputs[0], "Pow");
            float two;
            if (pw < 0 || !scalar(nodes[pw].inputs[1], &two) || two != 2.f) continue;
            if (nodes[pw].inputs[0] != nodes[sub].outputs[0]) continue;
            int rm1 = prod(nodes[sub].inputs[1], "ReduceMean");
            if (rm1 < 0 || nodes[rm1].inputs[0] != nodes[sub].inputs[0]) continue;
            auto ax1 = nodes[rm1].attr_ints("axes"), ax2 = nodes[rm2].attr_ints("axes");
            if (ax1.size() != 1 || ax2.size() != 1) continue;
            if (uses(nodes[sub].outputs[0]) != 2) continue;
            OnnxNode ln;
            ln.op_type = "LayerNormalization";
            ln.name = nodes[i].outputs[0];
            ln.inputs = {nodes[sub].inputs[0], nodes[mul].inputs[gamma_side], nodes[i].inputs[beta_side]};
            ln.outputs = nodes[i].outputs;
            Attr e;
            e.type = 1;
            e.f = eps;
            ln.attrs["epsilon"] = e;
            Attr ax;
            ax.type = 2;
            ax.i = ax1[0];
            ln.attrs["axis"] = ax;
            nodes[i] = ln;
            for (int d : {mul, div, sub, sq, addeps, rm2, pw, rm1}) dead[d] = 1;
            reindex();
        }
    }

    // x -> [Slice(i, i+1, axis 0) -> Squeeze(axis 0)]  or  Gather(idx i, axis 0); returns source + index
    bool head_select(const std::string& v, std::string* src, int* idx, std::vector<int>* kill) {
        int sqz = prod(v, "Squeeze");
        if (sqz >= 0) {
            int sl = prod(nodes[sqz].inputs[0], "Slice");
            if (sl < 0) return false;
            const OnnxNode& s = nodes[sl];
            int64_t st, en, ax = 0;
            if (s.inputs.size() >= 3) {
                auto a = consts.find(s.inputs[1]);
                auto b = consts.find(s.inputs[2]);
                if (a == consts.end() || b == consts.end()) return false;
                st = a->second.i.at(0);
                en = b->second.i.at(0);
                if (s.inputs.size() >= 4) {
                    auto c = consts.find(s.inputs[3]);
                    if (c == consts.end()) return false;
                    ax = c->second.i.at(0);
                }
            } else {
                st = s.attr_ints("starts").at(0);
                en = s.attr_ints("ends").at(0);
                auto axs = s.attr_ints("axes");
                ax = axs.empty() ? 0 : axs[0];
            }
            if (ax != 0 || en != st + 1) return false;
            *src = s.inputs[0];
            *idx = (int)st;
            kill->push_back(sqz);
            kill->push_back(sl);
            return true;
        }
        int ga = prod(v, "Gather");
        if (ga >= 0) {
            auto c = consts.find(nodes[ga].inputs[1]);
            if (c == consts.end() || c->second.numel() != 1 || nodes[ga].attr_i("axis", 0) != 0) return false;
            *src = nodes[ga].inputs[0];
            *idx = (int)c->second.i[0];
            kill->push_back(ga);
            return true;
        }
        return false;
    }

    void rewrite_attention() {
        for (size_t i = 0; i < nodes.size(); ++i) {
            if (dead[i] || nodes[i].op_type != "Softmax") continue;
            int qk = prod(nodes[i].inputs[0], "MatMul");
            if (qk < 0 || is_const(nodes[qk].inputs[1])) continue;
            std::vector<int> kill{(int)i, qk};
            // q side: optional Mul(scale)
            float scale = 1.f;
            std::string qv = nodes[qk].inputs[0];
            int qmul = prod(qv, "Mul");
            if (qmul >= 0) {
                float s;
                if (scalar(nodes[qmul].inputs[1], &s)) qv = nodes[qmul].inputs[0];
                else if (scalar(nodes[qmul].inputs[0], &s)) qv = nodes[qmul].inputs[1];
                else continue;
                scale = s;
                kill.push_back(qmul);
            }
            int kt = prod(nodes[qk].inputs[1], "Transpose");
            if (kt < 0) continue;
            auto p = nodes[kt].attr_ints("perm");
            if (p != std::vector<int64_t>({0, 1, 3, 2})) continue;
            kill.push_back(kt);
            // consumer chain: MatMul(attn, v) -> Transpose[0,2,1,3] -> Reshape
            auto& cs = consumers[nodes[i].outputs[0]];
            if (cs.size() != 1 || nodes[cs[0]].op_type != "MatMul") continue;
            int av = cs[0];
            auto& c2 = consumers[nodes[av].outputs[0]];
            if (c2.size() != 1 || nodes[c2[0]].op_type != "Transpose") continue;
            int tr = c2[0];
            if (nodes[tr].attr_ints("perm") != std::vector<int64_t>({0, 2, 1, 3})) continue;
            auto& c3 = consumers[nodes[tr].outputs[0]];
            if (c3.size() != 1 || nodes[c3[0]].op_type != "Reshape") continue;
            int rs = c3[0];
            std::string srcq, srck, srcv;
            int iq, ik, iv;
            if (!head_select(qv, &srcq, &iq, &kill)) continue;
            if (!head_select(nodes[kt].inputs[0], &srck, &ik, &kill)) continue;
            if (!head_select(nodes[av].inputs[1], &srcv, &iv, &kill)) continue;
            if (srcq != srck || srcq != srcv || iq != 0 || ik != 1 || iv != 2) continue;
            int t5 = prod(srcq, "Transpose");
            if (t5 < 0 || nodes[t5].attr_ints("perm") != std::vector<int64_t>({2, 0, 3, 1, 4})) continue;
            int r5 = prod(nodes[t5].inputs[0], "Reshape");
            if (r5 < 0) continue;
            auto shp = consts.find(nodes[r5].inputs[1]);
            if (shp == consts.end() || shp->second.numel() != 5 || shp->second.i[2] != 3) continue;
            int heads = (int)shp->second.i[3];
            OnnxNode at;
            at.op_type = "Attention";
            at.name = nodes[rs].outputs[0];
            at.inputs = {nodes[r5].inputs[0]};
            at.outputs = nodes[rs].outputs;
            Attr h;
            h.type = 2;
            h.i = heads;
            at.attrs["heads"] = h;
            Attr sc;
            sc.type = 1;
            sc.f = scale;
            at.attrs["scale"] = sc;
            nodes[rs] = at;
            kill.push_back(av);
            kill.push_back(tr);
            kill.push_back(t5);
            kill.push_back(r5);
            for (int d : kill) dead[d] = 1;
            reindex();
        }
    }

    void drop_unused() {
        bool changed = true;
        while (changed) {
            changed = false;
            for (size_t i = 0; i < nodes.size(); ++i) {
                if (dead[i]) continue;
                bool used = false;
                for (auto& o : nodes[i].outputs)
                    if (uses(o) > 0) used = true;
                if (!used) {
                    dead[i] = 1;
                    changed = true;
                }
            }
            if (changed) reindex();
        }
    }
};

const std::set<std::string> kHostOps = {"Shape", "Cast", "Slice", "Concat", "Gather", "Unsqueeze", "Squeeze",
                                        "Mul", "Add", "Sub", "Div"};

struct Lowering {
    Builder& b;
    CGraph& g;
    std::map<std::string, int> cprod;       // value -> index in g.nodes
    std::set<std::string> host_vals;

    Lowering(Builder& bb, CGraph& gg) : b(bb), g(gg) {}

    static bool epi_capable(int kind) {
        return kind == C_CONV || kind == C_DWCONV || kind == C_CONVT || kind == C_UNARY || kind == C_ADD ||
               kind == C_CHANVEC || kind == C_MUL;
    }

    // Try to append an epilogue step to the producer of `x`; returns false when a stand-alone node is needed.
    bool fuse_step(const std::string& x, const std::string& out, int type, float a, float bb) {
        auto it = cprod.find(x);
        if (it == cprod.end()) return false;
        CNode& p = g.nodes[it->second];
        if (!epi_capable(p.kind) || b.uses(x) != 1) return false;
        if (!p.epi.push(type, a, bb)) return false;
        int idx = it->second;
        cprod.erase(it);
        p.out = out;
        cprod[out] = idx;
        return true;
    }

    void unary(const OnnxNode& n, int type, float a = 0.f, float bb = 0.f, int in_idx = 0) {
        const std::string& x = n.inputs[in_idx];
        if (fuse_step(x, n.outputs[0], type, a, bb)) return;
        CNode c;
        c.kind = C_UNARY;
        c.name = n.name.empty() ? n.outputs[0] : n.name;
        c.in = {x};
        c.out = n.outputs[0];
        c.epi.push(type, a, bb);
        push(c);
    }

    void push(const CNode& c) {
        g.nodes.push_back(c);
        cprod[c.out] = (int)g.nodes.size() - 1;
    }

    bool all_host_inputs(const OnnxNode& n) {
        for (auto& in : n.inputs) {
            if (in.empty()) continue;
            if (b.is_const(in)) {
                continue;
            }
            if (!host_vals.count(in)) return false;
        }
        return true;
    }

    int add_conv_weights(const OnnxNode& n, bool transposed) {
        const HostTensor& w = b.consts.at(n.inputs[1]);
        if (!w.is_float() || w.dims.size() != 4) throw std::runtime_error("Conv weight must be 4-D float: " + n.inputs[1]);
        ConvWeights cw;
        int64_t d0 = w.dims[0], d1 = w.dims[1], kh = w.dims[2], kw = w.dims[3];
        cw.kh = (int)kh;
        cw.kw = (int)kw;
        if (!transposed) {            // OIHW -> [kh][kw][i][o]
            cw.cout = (int)d0;
            cw.cin = (int)d1;
            cw.w.resize(w.f.size());
            for (int64_t o = 0; o < d0; ++o)
                for (int64_t i = 0; i < d1; ++i)
                    for (int64_t y = 0; y < kh; ++y)
                        for (int64_t x = 0; x < kw; ++x)
                            cw.w[((y * kw + x) * d1 + i) * d0 + o] = w.f[((o * d1 + i) * kh + y) * kw + x];
        } else {                      // IOHW -> [kh][kw][i][o]
            cw.cin = (int)d0;
            cw.cout = (int)d1;
            cw.w.resize(w.f.size());
            for (int64_t i = 0; i < d0; ++i)
                for (int64_t o = 0; o < d1; ++o)
                    for (int64_t y = 0; y < kh; ++y)
                        for (int64_t x = 0; x < kw; ++x)
                            cw.w[((y * kw + x) * d0 + i) * d1 + o] = w.f[((i * d1 + o) * kh + y) * kw + x];
        }
        cw.bias.assign(cw.cout, 0.f);
        if (n.inputs.size() > 2 && !n.inputs[2].empty()) {
            const HostTensor& bt = b.consts.at(n.inputs[2]);
            for (int o = 0; o < cw.cout; ++o) cw.bias[o] = bt.f.at(o);
        }
        g.param_count += (double)w.f.size() + (n.inputs.size() > 2 ? cw.cout : 0);
        g.weights.push_back(cw);
        return (int)g.weights.size() - 1;
    }

    void fold_bn(ConvWeights& cw, const OnnxNode& bn, bool depthwise) {
        const HostTensor& gamma = b.consts.at(bn.inputs[1]);
        const HostTensor& beta = b.consts.at(bn.inputs[2]);
        const HostTensor& mean = b.consts.at(bn.inputs[3]);
        const HostTensor& var = b.consts.at(bn.inputs[4]);
        double eps = bn.attr_f("epsilon", 1e-5f);
        int co = cw.cout;
        size_t per = cw.w.size() / co;      // elements per output channel (co is the fastest axis)
        for (int o = 0; o < co; ++o) {
            double s = (double)gamma.f.at(o) / std::sqrt((double)var.f.at(o) + eps);
            for (size_t k = 0; k < per; ++k) cw.w[k * co + o] = (float)((double)cw.w[k * co + o] * s);
            cw.bias[o] = (float)(((double)cw.bias[o] - (double)mean.f.at(o)) * s + (double)beta.f.at(o));
        }
        (void)depthwise;
        g.param_count += 4.0 * co;
    }

    void lower_conv(const OnnxNode& n) {
        auto ks = n.attr_ints("kernel_shape");
        auto st = n.attr_ints("strides");
        auto pd = n.attr_ints("pads");
        auto dl = n.attr_ints("dilations");
        for (auto d : dl)
            if (d != 1) throw std::runtime_error("unsupported op: Conv with dilation != 1 (" + n.name + ")");
        if (!b.is_const(n.inputs[1])) throw std::runtime_error("unsupported op: Conv with dynamic weights (" + n.name + ")");
        CNode c;
        c.name = n.name.empty() ? n.outputs[0] : n.name;
        c.in = {n.inputs[0]};
        c.out = n.outputs[0];
        int groups = (int)n.attr_i("group", 1);
        c.weight_id = add_conv_weights(n, false);
        ConvWeights& cw = g.weights[c.weight_id];
        c.kh = cw.kh;
        c.kw = cw.kw;
        c.sh = st.size() > 0 ? (int)st[0] : 1;
        c.sw = st.size() > 1 ? (int)st[1] : 1;
        if (pd.size() == 4) {
            c.pt = (int)pd[0]; c.pl = (int)pd[1]; c.pb = (int)pd[2]; c.pr = (int)pd[3];
        }
        (void)ks;
        if (groups == 1) {
            c.kind = C_CONV;
            c.cin = cw.cin;
            c.cout = cw.cout;
        } else if (groups == cw.cout && cw.cin == 1) {
            c.kind = C_DWCONV;
            c.cin = c.cout = cw.cout;
            c.groups = groups;
            // [kh][kw][1][c] is already [kh][kw][c]
        } else {
            throw std::runtime_error("unsupported op: grouped Conv (group=" + std::to_string(groups) + ") at " + c.name);
        }
        push(c);
    }

    void run() {
        const auto& nodes = b.nodes;
        for (size_t i = 0; i < nodes.size(); ++i) {
            if (b.dead[i]) continue;
            const OnnxNode& n = nodes[i];
            const std::string& op = n.op_type;
            std::string nm = n.name.empty() ? n.outputs[0] : n.name;
            if (kHostOps.count(op) && (op == "Shape" || all_host_inputs(n))) {
                CNode c;
                c.kind = C_HOST;
                c.name = nm;
                c.in = n.inputs;
                c.out = n.outputs[0];
                c.host = n;
                push(c);
                host_vals.insert(n.outputs[0]);
                continue;
            }
            if (op == "Conv") {
                lower_conv(n);
            } else if (op == "ConvTranspose") {
                auto st = n.attr_ints("strides");
                auto pd = n.attr_ints("pads");
                const HostTensor& w = b.consts.at(n.inputs[1]);
                bool ok = w.dims.size() == 4 && w.dims[2] == 2 && w.dims[3] == 2 && st.size() == 2 && st[0] == 2 &&
                          st[1] == 2 && n.attr_i("group", 1) == 1;
                for (auto p : pd) ok = ok && p == 0;
                for (auto p : n.attr_ints("output_padding")) ok = ok && p == 0;
                if (!ok) throw std::runtime_error("unsupported op: ConvTranspose other than 2x2 stride 2 (" + nm + ")");
                CNode c;
                c.kind = C_CONVT;
                c.name = nm;
                c.in = {n.inputs[0]};
                c.out = n.outputs[0];
                c.weight_id = add_conv_weights(n, true);
                c.cin = g.weights[c.weight_id].cin;
                c.cout = g.weights[c.weight_id].cout;
                c.kh = c.kw = 2;
                c.sh = c.sw = 2;
                push(c);
            } else if (op == "BatchNormalization") {
                auto it = cprod.find(n.inputs[0]);
                bool ok = it != cprod.end() && b.uses(n.inputs[0]) == 1;
                if (ok) {
                    CNode& p = g.nodes[it->second];
                    ok = (p.kind == C_CONV || p.kind == C_DWCONV || p.kind == C_CONVT) && p.epi.n == 0;
                    if (ok) {
                        fold_bn(g.weights[p.weight_id], n, p.kind == C_DWCONV);
                        int idx = it->second;
                        cprod.erase(it);
                        p.out = n.outputs[0];
                        cprod[p.out] = idx;
                    }
                }
                if (!ok) throw std::runtime_error("unsupported op: BatchNormalization not directly after a Conv (" + nm + ")");
            } else if (op == "MatMul" || op == "Gemm") {
                if (!b.is_const(n.inputs[1]))
                    throw std::runtime_error("unsupported op: MatMul with two dynamic operands outside attention (" + nm + ")");
                const HostTensor& w = b.consts.at(n.inputs[1]);
                if (w.dims.size() != 2) throw std::runtime_error("unsupported op: MatMul weight rank != 2 (" + nm + ")");
                bool tb = op == "Gemm" && n.attr_i("transB", 0);
                ConvWeights cw;
                cw.cin = (int)(tb ? w.dims[1] : w.dims[0]);
                cw.cout = (int)(tb ? w.dims[0] : w.dims[1]);
                cw.w.resize(w.f.size());
                for (int i2 = 0; i2 < cw.cin; ++i2)
                    for (int o = 0; o < cw.cout; ++o)
                        cw.w[(size_t)i2 * cw.cout + o] = tb ? w.f[(size_t)o * cw.cin + i2] : w.f[(size_t)i2 * cw.cout + o];
                cw.bias.assign(cw.cout, 0.f);
                if (op == "Gemm" && n.inputs.size() > 2) {
                    const HostTensor& bt = b.consts.at(n.inputs[2]);
                    for (int o = 0; o < cw.cout; ++o) cw.bias[o] = bt.f.at(o);
                }
                g.param_count += (double)w.f.size();
                g.weights.push_back(cw);
                CNode c;
                c.kind = C_CONV;
                c.is_linear = true;
                c.name = nm;
                c.in = {n.inputs[0]};
                c.out = n.outputs[0];
                c.weight_id = (int)g.weights.size() - 1;
                c.cin = cw.cin;
                c.cout = cw.cout;
                push(c);
            } else if (op == "Add" || op == "Mul" || op == "Sub" || op == "Div") {
                int cs = b.is_const(n.inputs[1]) ? 1 : b.is_const(n.inputs[0]) ? 0 : -1;
                if (cs < 0) {
                    if (op != "Add" && op != "Mul") throw std::runtime_error("unsupported op: tensor " + op + " (" + nm + ")");
                    CNode c;
                    c.kind = op == "Add" ? C_ADD : C_MUL;
                    c.name = nm;
                    c.in = {n.inputs[0], n.inputs[1]};
                    c.out = n.outputs[0];
                    push(c);
                    continue;
                }
                const HostTensor& k = b.consts.at(n.inputs[cs]);
                const std::string& x = n.inputs[1 - cs];
                if (k.numel() == 1) {
                    float v = k.is_float() ? k.f[0] : (float)k.i[0];
                    if (op == "Add") unary(n, EPI_AFFINE, 1.f, v, 1 - cs);
                    else if (op == "Mul") unary(n, EPI_AFFINE, v, 0.f, 1 - cs);
                    else if (op == "Sub" && cs == 1) unary(n, EPI_AFFINE, 1.f, -v, 0);
                    else if (op == "Div" && cs == 1) unary(n, EPI_AFFINE, 1.f / v, 0.f, 0);
                    else throw std::runtime_error("unsupported op: const " + op + " tensor (" + nm + ")");
                    continue;
                }
                if (op != "Add" && op != "Mul") throw std::runtime_error("unsupported op: vector " + op + " (" + nm + ")");
                // per-channel vector: [C] on the last axis, or [1,C,1,1] / [C,1,1]
                auto it = cprod.find(x);
                if (op == "Add" && it != cprod.end() && b.uses(x) == 1) {
                    CNode& p = g.nodes[it->second];
                    if ((p.kind == C_CONV || p.kind == C_CONVT || p.kind == C_DWCONV) && p.epi.n == 0 &&
                        (int64_t)p.cout == k.numel()) {
                        auto& bias = g.weights[p.weight_id].bias;
                        for (int o = 0; o < p.cout; ++o) bias[o] += k.f[o];
                        g.param_count += p.cout;
                        int idx = it->second;
                        cprod.erase(it);
                        p.out = n.outputs[0];
                        cprod[p.out] = idx;
                        continue;
                    }
                }
                VecParam vp;
                vp.a = k.f;
                g.vecs.push_back(vp);
                CNode c;
                c.kind = C_CHANVEC;
                c.name = nm;
                c.in = {x};
                c.out = n.outputs[0];
                c.vec_id = (int)g.vecs.size() - 1;
                c.vec_op = op == "Mul";
                c.cout = (int)k.numel();
                push(c);
            } else if (op == "Relu") {
                unary(n, EPI_RELU);
            } else if (op == "Sigmoid") {
                unary(n, EPI_SIGMOID);
            } else if (op == "HardSwish") {
                unary(n, EPI_HSWISH);
            } else if (op == "Swish") {
                unary(n, EPI_SWISH);
            } else if (op == "HardSigmoid") {
                unary(n, EPI_HSIGMOID, n.attr_f("alpha", 0.2f), n.attr_f("beta", 0.5f));
            } else if (op == "Clip") {
                float lo = n.attr_f("min", -3.4e38f), hi = n.attr_f("max", 3.4e38f);
                if (n.inputs.size() > 1 && !n.inputs[1].empty() && !b.scalar(n.inputs[1], &lo))
                    throw std::runtime_error("unsupported op: Clip with dynamic min (" + nm + ")");
                if (n.inputs.size() > 2 && !n.inputs[2].empty() && !b.scalar(n.inputs[2], &hi))
                    throw std::runtime_error("unsupported op: Clip with dynamic max (" + nm + ")");
                unary(n, EPI_CLIP, lo, hi);
            } else if (op == "Identity" || op == "Dropout") {
                unary(n, EPI_AFFINE, 1.f, 0.f);
            } else if (op == "GlobalAveragePool") {
                CNode c;
                c.kind = C_GAP;
                c.name = nm;
                c.in = {n.inputs[0]};
                c.out = n.outputs[0];
                push(c);
            } else if (op == "MaxPool" || op == "AveragePool") {
                CNode c;
                c.kind = op == "MaxPool" ? C_MAXPOOL : C_AVGPOOL;
                c.name = nm;
                c.in = {n.inputs[0]};
                c.out = n.outputs[0];
                auto ks = n.attr_ints("kernel_shape");
                auto st = n.attr_ints("strides");
                auto pd = n.attr_ints("pads");
                c.kh = (int)ks.at(0);
                c.kw = (int)ks.at(1);
                c.sh = st.size() > 0 ? (int)st[0] : 1;
                c.sw = st.size() > 1 ? (int)st[1] : 1;
                if (pd.size() == 4) {
                    c.pt = (int)pd[0]; c.pl = (int)pd[1]; c.pb = (int)pd[2]; c.pr = (int)pd[3];
                }
                c.ceil_mode = (int)n.attr_i("ceil_mode", 0);
                c.count_include_pad = (int)n.attr_i("count_include_pad", 0);
                push(c);
            } else if (op == "Resize" || op == "Upsample") {
                std::string mode = n.attr_s("mode", "nearest");
                if (mode != "nearest") throw std::runtime_error("unsupported op: Resize mode " + mode + " (" + nm + ")");
                size_t si = op == "Resize" ? 2 : 1;
                if (n.inputs.size() <= si || !b.is_const(n.inputs[si]))
                    throw std::runtime_error("unsupported op: Resize without constant scales (" + nm + ")");
                const HostTensor& sc = b.consts.at(n.inputs[si]);
                if (sc.numel() != 4 || sc.f[0] != 1.f || sc.f[1] != 1.f || sc.f[2] != std::floor(sc.f[2]) ||
                    sc.f[3] != std::floor(sc.f[3]) || sc.f[2] < 1.f || sc.f[3] < 1.f)
                    throw std::runtime_error("unsupported op: Resize with non-integer scales (" + nm + ")");
                CNode c;
                c.kind = C_RESIZE;
                c.name = nm;
                c.in = {n.inputs[0]};
                c.out = n.outputs[0];
                c.fh = (int)sc.f[2];
                c.fw = (int)sc.f[3];
                push(c);
            } else if (op == "Concat") {
                CNode c;
                c.kind = C_CONCAT;
                c.name = nm;
                c.in = n.inputs;
                c.out = n.outputs[0];
                c.axis = (int)n.attr_i("axis", 1);
                push(c);
            } else if (op == "Softmax") {
                CNode c;
                c.kind = C_SOFTMAX;
                c.name = nm;
                c.in = {n.inputs[0]};
                c.out = n.outputs[0];
                c.axis = (int)n.attr_i("axis", b.m.opset >= 13 ? -1 : 1);
                push(c);
            } else if (op == "LayerNormalization") {
                VecParam vp;
                vp.a = b.consts.at(n.inputs[1]).f;
                vp.b = n.inputs.size() > 2 ? b.consts.at(n.inputs[2]).f : std::vector<float>(vp.a.size(), 0.f);
                g.param_count += 2.0 * vp.a.size();
                g.vecs.push_back(vp);
                CNode c;
                c.kind = C_LAYERNORM;
                c.name = nm;
                c.in = {n.inputs[0]};
                c.out = n.outputs[0];
                c.eps = n.attr_f("epsilon", 1e-5f);
                c.ln_id = (int)g.vecs.size() - 1;
                c.cout = (int)vp.a.size();
                push(c);
            } else if (op == "Attention") {
                CNode c;
                c.kind = C_ATTENTION;
                c.name = nm;
                c.in = {n.inputs[0]};
                c.out = n.outputs[0];
                c.heads = (int)n.attr_i("heads", 1);
                c.scale = n.attr_f("scale", 1.f);
                push(c);
            } else if (op == "Reshape") {
                CNode c;
                c.kind = C_RESHAPE;
                c.name = nm;
                c.in = {n.inputs[0], n.inputs[1]};
                c.out = n.outputs[0];
                push(c);
            } else if (op == "Transpose") {
                CNode c;
                c.kind = C_TRANSPOSE;
                c.name = nm;
                c.in = {n.inputs[0]};
                c.out = n.outputs[0];
                c.perm = n.attr_ints("perm");
                push(c);
            } else if (op == "Squeeze") {
                CNode c;
                c.kind = C_SQUEEZE;
                c.name = nm;
                c.in = {n.inputs[0]};
                c.out = n.outputs[0];
                c.axes = n.attr_ints("axes");
                if (c.axes.empty() && n.inputs.size() > 1 && b.is_const(n.inputs[1])) c.axes = b.consts.at(n.inputs[1]).i;
                push(c);
            } else if (op == "Flatten") {
                CNode c;
                c.kind = C_FLATTEN;
                c.name = nm;
                c.in = {n.inputs[0]};
                c.out = n.outputs[0];
                c.axis = (int)n.attr_i("axis", 1);
                push(c);
            } else {
                throw std::runtime_error("unsupported op: " + op + " (node " + nm + ")");
            }
        }
    }

    // GAP -> Conv1x1[relu] -> Conv1x1[hsigmoid] -> Mul(x, .)   =>  SE
    void fuse_se() {
        std::map<std::string, int> prod;
        std::map<std::string, int> nuse;
        for (size_t i = 0; i < g.nodes.size(); ++i) {
            prod[g.nodes[i].out] = (int)i;
            for (auto& in : g.nodes[i].in) nuse[in]++;
        }
        std::vector<char> kill(g.nodes.size(), 0);
        for (size_t i = 0; i < g.nodes.size(); ++i) {
            CNode& mul = g.nodes[i];
            if (mul.kind != C_MUL) continue;
            for (int side = 0; side < 2; ++side) {
                auto p2 = prod.find(mul.in[1 - side]);
                if (p2 == prod.end()) continue;
                CNode& fc2 = g.nodes[p2->second];
                if (fc2.kind != C_CONV || fc2.kh != 1 || fc2.kw != 1 || fc2.epi.n != 1 || fc2.epi.s[0].type != EPI_HSIGMOID ||
                    nuse[fc2.out] != 1)
                    continue;
                auto p1 = prod.find(fc2.in[0]);
                if (p1 == prod.end()) continue;
                CNode& fc1 = g.nodes[p1->second];
                if (fc1.kind != C_CONV || fc1.kh != 1 || fc1.kw != 1 || fc1.epi.n != 1 || fc1.epi.s[0].type != EPI_RELU ||
                    nuse[fc1.out] != 1)
                    continue;
                auto p0 = prod.find(fc1.in[0]);
                if (p0 == prod.end()) continue;
                CNode& gap = g.nodes[p0->second];
                if (gap.kind != C_GAP || gap.in[0] != mul.in[side] || nuse[gap.out] != 1) continue;
                CNode se;
                se.kind = C_SE;
                se.name = mul.name;
                se.in = {mul.in[side]};
                se.out = mul.out;
                se.epi = mul.epi;
                se.se_w1 = fc1.weight_id;
                se.se_w2 = fc2.weight_id;
                se.se_mid = fc1.cout;
                se.cin = se.cout = fc2.cout;
                se.scale = fc2.epi.s[0].a;      // hsigmoid alpha
                se.eps = fc2.epi.s[0].b;        // hsigmoid beta
                kill[p2->second] = kill[p1->second] = kill[p0->second] = 1;
                g.nodes[i] = se;
                break;
            }
        }
        std::vector<CNode> kept;
        for (size_t i = 0; i < g.nodes.size(); ++i)
            if (!kill[i]) kept.push_back(g.nodes[i]);
        g.nodes.swap(kept);
    }
};

}  // namespace

std::unique_ptr<CGraph> build_cgraph(const OnnxModel& m) {
    if (m.inputs.size() != 1) throw std::runtime_error("model must have exactly one input, has " + std::to_string(m.inputs.size()));
    if (m.outputs.size() != 1) throw std::runtime_error("model must have exactly one output, has " + std::to_string(m.outputs.size()));
    Builder b(m);
    b.rewrite_hardswish();
    b.rewrite_swish();
    b.rewrite_layernorm();
    b.rewrite_attention();
    b.drop_unused();
    std::unique_ptr<CGraph> g(new CGraph());
    g->input_name = m.inputs[0].name;
    g->output_name = m.outputs[0].name;
    g->input_shape = m.inputs[0].shape;
    g->output_shape = m.outputs[0].shape;
    g->opset = m.opset;
    for (auto& kv : b.consts)
        if (!kv.second.is_float() || kv.second.numel() <= 8) g->host_consts[kv.first] = kv.second;
    Lowering l(b, *g);
    l.run();
    l.fuse_se();
    return g;
}

static const char* kind_name(int k) {
    switch (k) {
        case C_CONV: return "conv";
        case C_DWCONV: return "dwconv";
        case C_CONVT: return "convT2x2";
        case C_SE: return "se";
        case C_GAP: return "gap";
        case C_MAXPOOL: return "maxpool";
        case C_AVGPOOL: return "avgpool";
        case C_RESIZE: return "resize";
        case C_CONCAT: return "concat";
        case C_ADD: return "add";
        case C_MUL: return "mul";
        case C_UNARY: return "unary";
        case C_LAYERNORM: return "layernorm";
        case C_ATTENTION: return "attention";
        case C_SOFTMAX: return "softmax";
        case C_RESHAPE: return "reshape";
        case C_TRANSPOSE: return "transpose";
        case C_SQUEEZE: return "squeeze";
        case C_FLATTEN: return "flatten";
        case C_CHANVEC: return "chanvec";
        case C_HOST: return "host";
    }
    return "?";
}

std::string dump_cgraph(const CGraph& g) {
    std::ostringstream os;
    os << "input " << g.input_name << " output " << g.output_name << " params " << (long long)g.param_count << "\n";
    for (auto& c : g.nodes) {
        os << kind_name(c.kind);
        if (c.kind == C_HOST) os << ":" << c.host.op_type;
        if (c.kind == C_CONV || c.kind == C_DWCONV || c.kind == C_CONVT)
            os << " " << c.cin << "->" << c.cout << " k" << c.kh << "x" << c.kw << " s" << c.sh << "x" << c.sw
               << (c.is_linear ? " linear" : "");
        if (c.kind == C_SE) os << " c" << c.cout << " mid" << c.se_mid;
        if (c.kind == C_RESIZE) os << " x" << c.fh;
        if (c.kind == C_ATTENTION) os << " heads" << c.heads << " scale" << c.scale;
        if (c.epi.n) {
            os << " epi[";
            for (int i = 0; i < c.epi.n; ++i) os << (i ? "," : "") << c.epi.s[i].type;
            os << "]";
        }
        os << "  (";
        for (size_t i = 0; i < c.in.size(); ++i) os << (i ? ", " : "") << c.in[i];
        os << ") -> " << c.out << "\n";
    }
    return os.str();
}

}  // namespace b200ocr

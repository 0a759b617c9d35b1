// B200 graph executor: canonical graph -> per-shape launch plan -> kernels on a stream.
#pragma once
#include <cuda_runtime.h>

#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "graph.h"

namespace b200ocr {

enum Precision : int { PREC_FP32 = 0, PREC_BF16 = 1 };

// How a logical ONNX tensor sits on the physical NHWC(cpad) buffer.
enum Form : int {
    F_NCHW = 0,   // logical [n,c,h,w]
    F_NCS,        // logical [n,c,h*w]
    F_NSC,        // logical [n,h*w,c]
    F_NHWC,       // logical [n,h,w,c]
    F_NC,         // logical [n,c]      (h = w = 1)
    F_HOST,       // small integer vector living on the host (shape arithmetic)
};

struct PVal {
    int form = F_NCHW;
    int n = 0, h = 0, w = 0, c = 0, cp = 0;
    int root = -1;                 // buffer id
    // virtual nearest-upsample of `up_src` (a C_RESIZE that was not materialised)
    int up_src = -1, fh = 1, fw = 1;
    // lazily applied SE gate: the logical value is x * gate[n][c] (+ x when gate_res) with x = this buffer; consumers that
    // can apply it on the fly (element-wise add, concat) read x and the gate, everyone else gets it materialised
    int gate = -1, gate_res = 0;
    // per (image, channel) sums of this tensor accumulated by its producer (fixed point, see common.cuh) for an SE gate
    int sums = -1;                 // buffer id (zeroed by a memset node right before the producer runs)
    int sums_replicas = 1;
    std::vector<int64_t> host;
    std::vector<int64_t> logical() const;
    int64_t elems() const { return (int64_t)n * h * w * cp; }
};

struct DevWeights {
    float* w = nullptr;            // [Kpad][ColsW] fp32 (dense) or [taps][CP] (depthwise)
    float* bias = nullptr;         // [ColsW]
    void* w_bf16 = nullptr;        // tcgen05 operand pack: bf16 [tc_rows][tc_k], K-major (bf16 models only)
    float* tc_bias = nullptr;
    // pixel-packed operand for narrow pointwise layers (CinP 16 / 32): P = 64 / CinP pixels form one 64-wide GEMM row and
    // the weight becomes block diagonal [P * CoutP rows][64] — same tensors in memory, 128-byte TMA rows, P x fewer tiles
    void* pk_w = nullptr;
    float* pk_bias = nullptr;
    int pk_P = 0, pk_rows = 0;
    bool tc_prefolded = false;     // the epilogue's leading affine (a0, b0) is folded into w_bf16 / tc_bias
    float* head_bias = nullptr;    // CTC head: bias with -1e30 beyond the last class (padding never wins the argmax)
    int tc_bn = 0, tc_rows = 0, tc_k = 0;
    int K = 0, Kpad = 0, ColsW = 0;
};

struct Plan {
    int n = 0, h = 0, w = 0;
    std::vector<PVal> vals;
    std::vector<size_t> buf_off, buf_bytes;
    size_t arena_bytes = 0;
    char** model_arena = nullptr;  // the owning Model's activation slab (shared by all of its plans) or &own_arena
    char* own_arena = nullptr;     // private slab: lets this plan run concurrently with other plans on another stream
    std::vector<std::shared_ptr<void>> owned;     // small device allocations that live as long as the plan
    std::vector<std::function<void(cudaStream_t)>> steps;
    std::vector<std::string> step_names;
    std::vector<double> step_flops, step_bytes;   // algorithmic work per launch (unique in+out+weight bytes)
    int in_val = -1, out_val = -1;
    // fused CTC head (bf16 rec plans built with head = true): per (b, t) row the argmax class and its softmax value
    bool head_mode = false;
    int head_rows = 0, head_T = 0, head_idx_buf = -1, head_prob_buf = -1;
    const int* head_idx() const { return (const int*)bufptr(head_idx_buf); }
    const float* head_prob() const { return (const float*)bufptr(head_prob_buf); }
    // Pre-processing fused into the stem (plan_for(..., fused_pre)): the first step reads u8 pages through `pre_jobs`
    // (device table of n ResizeJob, overwritten by the caller before every launch) instead of in_ptr(); pre_jobs stays
    // null when the graph's first layer is not an eligible stem (the caller then runs resize_norm_kernel as before).
    bool fused_pre = false;
    void* pre_jobs = nullptr;
    float pre_norm[8] = {0, 0, 0, 0, 1, 1, 1, 0};   // NormP: mul, sub[3], div[3], rec_style
    float* out_dense = nullptr;    // dense fp32 logical output (device)
    size_t out_dense_elems = 0;
    bool out_written_dense = false;
    std::vector<int64_t> out_shape;
    double flops = 0;              // 2*MACs of Conv/ConvT/MatMul at this shape
    int launches = 0;
    cudaGraphExec_t graph_exec = nullptr;
    const char* graph_arena = nullptr;   // slab address the graph was recorded against
    const char* warm_arena = nullptr;
    int warm_launches = 0;
    bool sums_clean = false;       // the SE channel-sum buffers were zeroed by a first launch (see zero_sums in build_steps)
    void* ptr(int v) const { return *model_arena + buf_off[vals[v].root]; }
    void* bufptr(int b) const { return *model_arena + buf_off[b]; }
    void* in_ptr() const { return ptr(in_val); }
    void launch(cudaStream_t st);
    ~Plan();
};

class Model {
public:
    Model(const std::string& path, int precision);
    ~Model();
    Model(const Model&) = delete;

    // Input already on the device in the executor's own layout: NHWC, storage type T, cpad(3)=4.
    // Returns the plan (its out_dense holds the logical fp32 output after the stream completes).
    Plan* run_device(const void* nhwc_in, int n, int h, int w, cudaStream_t st, bool want_dense = true);
    // Host NCHW fp32 in -> logical fp32 out (the onnxruntime-compatible seam).
    void run_host(const float* nchw, int n, int c, int h, int w, std::vector<int64_t>* out_shape,
                  float* out, size_t out_capacity, cudaStream_t st);
    // head = true: the final Linear + Softmax run as the fused CTC head (no dense output; see Plan::head_*)
    // own_arena = true: give the plan a private activation slab when it is small (so independent small batches can
    // overlap on several streams); falls back to the shared slab when the private budget is exhausted.
    //   own_arena: 0 shared slab, 1 private when small, 2 always private; slot distinguishes otherwise identical plans
    //   that must be able to run at the same time (the two det lanes).
    Plan* plan_for(int n, int h, int w, bool head = false, int own_arena = 0, int slot = 0, bool fused_pre = false);
    // normalisation constants of the fused pre-processing (NormP layout: mul, sub[3], div[3], rec_style)
    void set_pre_norm(const float v[8]) { for (int k = 0; k < 8; ++k) pre_norm_[k] = v[k]; }
    bool fused_pre_default = false;    // profile / time_plan build the plan the stages actually run
    // Host NCHW fp32 in -> plan launched on `st` (no synchronisation); head selects the fused CTC head variant.
    Plan* launch_host_input(const float* nchw, int n, int c, int h, int w, cudaStream_t st, bool head);
    // Per-step device time (ms, averaged over iters) measured with CUDA events between launches.
    void profile(int n, int h, int w, int iters, cudaStream_t st, std::vector<float>* ms);
    size_t elem_size() const { return prec_ == PREC_BF16 ? 2 : 4; }
    int precision() const { return prec_; }
    const CGraph& graph() const { return *g_; }
    int in_channels() const { return in_c_; }
    std::string describe() const;

private:
    std::unique_ptr<Plan> build_plan(int n, int h, int w, bool head, int own_arena, bool fused_pre);
    float pre_norm_[8] = {0, 0, 0, 0, 1, 1, 1, 0};
    size_t private_bytes_ = 0;
    template <typename T> void build_steps(Plan& p);
    std::unique_ptr<CGraph> g_;
    int prec_;
    int in_c_ = 3;
    std::vector<DevWeights> dw_;
    std::vector<std::pair<float*, float*>> dvec_;
    std::map<std::vector<int>, std::unique_ptr<Plan>> plans_;
    std::vector<std::unique_ptr<Plan>> retired_;   // previous cache generation (see plan_for)
    char* arena_ = nullptr;        // activation slab, grown to the largest plan
    size_t arena_cap_ = 0;
    float* stage_in_ = nullptr;    // device staging for host NCHW input
    size_t stage_in_bytes_ = 0;
    std::mutex mu_;
};

}  // namespace b200ocr

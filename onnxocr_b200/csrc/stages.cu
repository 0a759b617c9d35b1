#include "stages.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <numeric>
#include <thread>

#include "pre_kernels.cuh"

namespace b200ocr {

void* DevBuf::get(size_t bytes) {
    if (bytes > cap) {
        if (p) cudaFree(p);
        p = nullptr;
        cap = bytes + bytes / 4 + 256;
        B200_CUDA_CHECK(cudaMalloc(&p, cap));
    }
    return p;
}
DevBuf::~DevBuf() {
    if (p) cudaFree(p);
}
void* PinBuf::get(size_t bytes) {
    if (bytes > cap) {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = bytes + bytes / 4 + 256;
        B200_CUDA_CHECK(cudaMallocHost(&p, cap));
    }
    return p;
}
PinBuf::~PinBuf() {
    if (p) cudaFreeHost(p);
}

// DetResizeForTest.resize_image_type0 (operators.py:91-138) — Python float/int semantics in C++:
// int() truncates, round() is half-to-even (nearbyint under the default rounding mode).
// ------------------------------------------------------------------------------- trace
namespace {
struct TraceRec { std::string name; int tag; uintptr_t stream; cudaEvent_t a, b; };
struct TraceState {
    const char* path = getenv("B200OCR_TRACE");
    std::vector<TraceRec> recs;
    std::mutex mu;
};
TraceState& trace_state() { static TraceState t; return t; }
}  // namespace

TraceScope::TraceScope(const char* name, cudaStream_t s, int tag) : st(s) {
    TraceState& t = trace_state();
    if (!t.path || !t.path[0]) return;
    std::lock_guard<std::mutex> lk(t.mu);
    if (t.recs.size() >= 200000) return;
    TraceRec r{name, tag, (uintptr_t)s, nullptr, nullptr};
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    cudaEventRecord(r.a, s);
    idx = (int)t.recs.size();
    t.recs.push_back(r);
}
TraceScope::~TraceScope() {
    if (idx < 0) return;
    TraceState& t = trace_state();
    std::lock_guard<std::mutex> lk(t.mu);
    cudaEventRecord(t.recs[idx].b, st);
}
void trace_dump() {
    TraceState& t = trace_state();
    if (!t.path || !t.path[0] || t.recs.empty()) return;
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(t.mu);
    FILE* f = fopen(t.path, "a");
    if (!f) return;
    std::map<uintptr_t, int> sid;
    for (auto& r : t.recs) {
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, t.recs[0].a, r.a);
        cudaEventElapsedTime(&b, t.recs[0].a, r.b);
        if (!sid.count(r.stream)) { int k = (int)sid.size(); sid[r.stream] = k; }
        fprintf(f, "{\"name\": \"%s\", \"tag\": %d, \"stream\": %d, \"t0\": %.4f, \"t1\": %.4f}\n", r.name.c_str(), r.tag, sid[r.stream], a, b);
    }
    for (auto& r : t.recs) {
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    fclose(f);
    t.recs.clear();
}

DetGeom det_geometry(int h, int w, const b200ocr_det_params& p) {
    DetGeom g;
    g.src_h = h;
    g.src_w = w;
    g.pad_h = h;
    g.pad_w = w;
    if (h + w < 64) {
        g.pad_h = std::max(32, h);
        g.pad_w = std::max(32, w);
    }
    int hh = g.pad_h, ww = g.pad_w;
    double limit = (double)p.limit_side_len, ratio = 1.0;
    if (p.limit_type == 0) {
        if ((double)std::max(hh, ww) > limit) ratio = hh > ww ? limit / hh : limit / ww;
    } else if (p.limit_type == 1) {
        if ((double)std::min(hh, ww) < limit) ratio = hh < ww ? limit / hh : limit / ww;
    } else {
        ratio = limit / std::max(hh, ww);
    }
    int rh = (int)(hh * ratio), rw = (int)(ww * ratio);
    rh = std::max((int)(std::nearbyint(rh / 32.0) * 32), 32);
    rw = std::max((int)(std::nearbyint(rw / 32.0) * 32), 32);
    g.dst_h = rh;
    g.dst_w = rw;
    g.ratio_h = rh / (double)hh;
    g.ratio_w = rw / (double)ww;
    return g;
}

static NormP det_norm() {
    NormP n;
    n.mul = (float)(1.0 / 255.0);
    const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
    for (int c = 0; c < 3; ++c) {
        n.sub[c] = mean[c];
        n.div[c] = stdv[c];
    }
    n.rec_style = 0;
    return n;
}
void Stages::det_norm_values(float v[8]) {
    const NormP n = det_norm();
    v[0] = n.mul;
    for (int c = 0; c < 3; ++c) { v[1 + c] = n.sub[c]; v[4 + c] = n.div[c]; }
    v[7] = (float)n.rec_style;
}
static NormP rec_norm() {
    NormP n = det_norm();
    n.rec_style = 1;
    return n;
}

// OpenCV's bicubic weight table (A = -0.75, 5-bit sub-pixel, 15-bit weights)
static std::vector<short> cubic_table() {
    const float A = -0.75f;
    float tab[32][4];
    for (int i = 0; i < 32; ++i) {
        float x = (float)i * (1.f / 32);
        float c0 = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
        float c1 = ((A + 2) * x - (A + 3)) * x * x + 1;
        float c2 = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
        tab[i][0] = c0; tab[i][1] = c1; tab[i][2] = c2;
        tab[i][3] = 1.f - c0 - c1 - c2;
    }
    std::vector<short> out(32 * 32 * 16);
    for (int iy = 0; iy < 32; ++iy)
        for (int ix = 0; ix < 32; ++ix) {
            int w[4][4], sum = 0;
            for (int ky = 0; ky < 4; ++ky)
                for (int kx = 0; kx < 4; ++kx) {
                    volatile float v = tab[iy][ky] * tab[ix][kx];
                    volatile float s = v * 32768.f;
                    int r = (int)std::nearbyint((double)s);
                    r = std::min(std::max(r, -32768), 32767);
                    w[ky][kx] = r;
                    sum += r;
                }
            if (sum != 32768) {
                int diff = sum - 32768, by = 2, bx = 2;
                for (int ky = 2; ky < 4; ++ky)
                    for (int kx = 2; kx < 4; ++kx) {
                        if (diff < 0 ? w[ky][kx] > w[by][bx] : w[ky][kx] < w[by][bx]) { by = ky; bx = kx; }
                    }
                w[by][bx] -= diff;
            }
            for (int ky = 0; ky < 4; ++ky)
                for (int kx = 0; kx < 4; ++kx) out[((iy * 32 + ix) * 4 + ky) * 4 + kx] = (short)w[ky][kx];
        }
    return out;
}

// cv2.getPerspectiveTransform (8x8 LU with partial pivoting, float64) + 3x3 inverse.
static bool perspective_inverse(const float* src /*[4][2]*/, int cw, int ch, double mi[9]) {
    const double dst[4][2] = {{0, 0}, {(double)cw, 0}, {(double)cw, (double)ch}, {0, (double)ch}};
    double a[8][8], b[8];
    for (int i = 0; i < 4; ++i) {
        double sx = src[2 * i], sy = src[2 * i + 1], dx = dst[i][0], dy = dst[i][1];
        double r0[8] = {sx, sy, 1, 0, 0, 0, -sx * dx, -sy * dx};
        double r1[8] = {0, 0, 0, sx, sy, 1, -sx * dy, -sy * dy};
        memcpy(a[i], r0, sizeof(r0));
        memcpy(a[i + 4], r1, sizeof(r1));
        b[i] = dx;
        b[i + 4] = dy;
    }
    const int n = 8;
    bool ok = true;
    for (int i = 0; i < n && ok; ++i) {
        int k = i;
        for (int j = i + 1; j < n; ++j)
            if (std::fabs(a[j][i]) > std::fabs(a[k][i])) k = j;
        if (std::fabs(a[k][i]) < 2.220446049250313e-16 * 100) { ok = false; break; }
        if (k != i) {
            for (int c = 0; c < n; ++c) std::swap(a[i][c], a[k][c]);
            std::swap(b[i], b[k]);
        }
        double d = -1.0 / a[i][i];
        for (int j = i + 1; j < n; ++j) {
            double alpha = a[j][i] * d;
            for (int c = i + 1; c < n; ++c) a[j][c] += alpha * a[i][c];
            b[j] += alpha * b[i];
        }
    }
    double m[9];
    if (ok) {
        for (int i = n - 1; i >= 0; --i) {
            double s = b[i];
            for (int k = i + 1; k < n; ++k) s -= a[i][k] * b[k];
            b[i] = s / a[i][i];
        }
        for (int i = 0; i < 8; ++i) m[i] = b[i];
    } else {
        for (int i = 0; i < 8; ++i) m[i] = 0;
    }
    m[8] = 1.0;
    double d = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
    if (d == 0) {
        for (int i = 0; i < 9; ++i) mi[i] = 0;
        return false;
    }
    d = 1.0 / d;
    mi[0] = (m[4] * m[8] - m[5] * m[7]) * d;
    mi[1] = (m[2] * m[7] - m[1] * m[8]) * d;
    mi[2] = (m[1] * m[5] - m[2] * m[4]) * d;
    mi[3] = (m[5] * m[6] - m[3] * m[8]) * d;
    mi[4] = (m[0] * m[8] - m[2] * m[6]) * d;
    mi[5] = (m[2] * m[3] - m[0] * m[5]) * d;
    mi[6] = (m[3] * m[7] - m[4] * m[6]) * d;
    mi[7] = (m[1] * m[6] - m[0] * m[7]) * d;
    mi[8] = (m[0] * m[4] - m[1] * m[3]) * d;
    return true;
}

static short* g_cubic_tab[16] = {nullptr};

Stages::Stages(cudaStream_t st) : st_(st) {
    for (auto& e : ev_) B200_CUDA_CHECK(cudaEventCreate(&e));
    for (auto& a : aux_) B200_CUDA_CHECK(cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking));
    {
        // B200OCR_DET_PRIO=1 gives the lanes distinct stream priorities (experiment, off): the block scheduler then drains
        // the networks one after the other at their solo speed (2.3 ms per 16 pages), which loses more than the overlap of
        // their DBPostProcess tails gains (concurrent networks fill each other's gaps: 1.8 ms each).  The rolling window
        // in det_enqueue is what de-phases the lanes instead.
        int lo = 0, hi = 0;
        B200_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));      // lo = least urgent (numerically larger)
        const char* e = getenv("B200OCR_DET_PRIO");
        const bool prio = e && e[0] == '1';
        for (int l = 0; l < kDetLanes; ++l) {
            int pr = prio ? std::max(hi, lo - 1 - (l % std::max(1, lo - hi))) : lo;
            B200_CUDA_CHECK(cudaStreamCreateWithPriority(&det_st_[l], cudaStreamNonBlocking, pr));
        }
    }
    for (auto& e : ev_det_done_) B200_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : ev_net_done_) B200_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    B200_CUDA_CHECK(cudaEventCreateWithFlags(&ev_fork_, cudaEventDisableTiming));
    for (auto& e : ev_join_) B200_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    db_ = nullptr;
    if (const char* e = getenv("B200OCR_AUX")) n_aux_ = std::max(1, std::min(kAux, atoi(e)));
}

Stages::~Stages() {
    trace_dump();
    for (auto& a : aux_) { cudaStreamSynchronize(a); cudaStreamDestroy(a); }
    for (auto& a : det_st_) { cudaStreamSynchronize(a); cudaStreamDestroy(a); }
    for (auto& e : ev_det_done_) cudaEventDestroy(e);
    for (auto& e : ev_net_done_) cudaEventDestroy(e);
    for (auto& d : db_lane_) db_free(d);
    cudaEventDestroy(ev_fork_);
    for (auto& e : ev_join_) cudaEventDestroy(e);
    for (auto& e : ev_) cudaEventDestroy(e);
    db_free(db_);
}

static const short* cubic_tab_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 15;
    if (!g_cubic_tab[dev]) {
        std::vector<short> t = cubic_table();
        B200_CUDA_CHECK(cudaMalloc(&g_cubic_tab[dev], t.size() * sizeof(short)));
        B200_CUDA_CHECK(cudaMemcpy(g_cubic_tab[dev], t.data(), t.size() * sizeof(short), cudaMemcpyHostToDevice));
        B200_CUDA_CHECK(cudaDeviceSynchronize());          // legacy-stream copy vs the engine's non-blocking streams
    }
    return g_cubic_tab[dev];
}

static ResizeJob make_job(const uint8_t* d_src, int h, int w, int stride, int ph, int pw, int dh, int dw, int rot180) {
    ResizeJob j;
    j.src.p = d_src; j.src.h = h; j.src.w = w; j.src.stride = stride; j.src.ph = ph; j.src.pw = pw; j.src.rot180 = rot180;
    j.dh = dh; j.dw = dw;
    j.scale_y = 1.0 / ((double)dh / ph);
    j.scale_x = 1.0 / ((double)dw / pw);
    return j;
}

// ------------------------------------------------------------------------------- det preprocess (Level A helper)
void Stages::det_preprocess_host(const uint8_t* bgr, int h, int w, int row_stride, const b200ocr_det_params& p, float* chw,
                                 size_t chw_capacity, int* out_h, int* out_w, double shape[4]) {
    DetGeom g = det_geometry(h, w, p);
    if (out_h) *out_h = g.dst_h;
    if (out_w) *out_w = g.dst_w;
    if (shape) {
        shape[0] = g.src_h; shape[1] = g.src_w; shape[2] = g.ratio_h; shape[3] = g.ratio_w;
    }
    if (!chw) return;
    size_t need = (size_t)3 * g.dst_h * g.dst_w;
    if (chw_capacity < need) throw std::runtime_error("chw buffer too small");
    if (row_stride <= 0) row_stride = w * 3;
    size_t img_bytes = (size_t)h * row_stride;
    uint8_t* d_img = (uint8_t*)d_img_.get(img_bytes + sizeof(ResizeJob));
    B200_CUDA_CHECK(cudaMemcpyAsync(d_img, bgr, img_bytes, cudaMemcpyHostToDevice, st_));
    ResizeJob job = make_job(d_img, h, w, row_stride, g.pad_h, g.pad_w, g.dst_h, g.dst_w, 0);
    ResizeJob* d_job = (ResizeJob*)d_misc_.get(sizeof(ResizeJob));
    B200_CUDA_CHECK(cudaMemcpyAsync(d_job, &job, sizeof(job), cudaMemcpyHostToDevice, st_));
    float* d_out = (float*)d_chw_.get(need * sizeof(float));
    dim3 grid((unsigned)ceil_div(g.dst_w, 256), g.dst_h, 1);
    resize_norm_kernel<float><<<grid, 256, 0, st_>>>(d_job, g.dst_h, g.dst_w, d_out, 1, det_norm());
    launches += 1;
    B200_CUDA_CHECK(cudaGetLastError());
    B200_CUDA_CHECK(cudaMemcpyAsync(chw, d_out, need * sizeof(float), cudaMemcpyDeviceToHost, st_));
    B200_CUDA_CHECK(cudaStreamSynchronize(st_));
}

// ------------------------------------------------------------------------------- crops
void Stages::crop_boxes_host(const uint8_t* bgr, int h, int w, int row_stride, const float* boxes, int n, uint8_t* crops_out,
                             size_t crops_capacity, int* crop_dims_out) {
    if (row_stride <= 0) row_stride = w * 3;
    std::vector<WarpJob> jobs(n);
    size_t off = 0;
    int max_w = 1, max_h = 1;
    for (int i = 0; i < n; ++i) {
        const float* b = boxes + (size_t)i * 8;
        auto norm = [&](int a, int c) {
            float dx = b[2 * a] - b[2 * c], dy = b[2 * a + 1] - b[2 * c + 1];
            // np.linalg.norm on float32: sqrt(sum of squares) evaluated in float32
            float s = dx * dx;
            s += dy * dy;
            return std::sqrt(s);
        };
        int cw = (int)std::max(norm(0, 1), norm(2, 3));
        int ch = (int)std::max(norm(0, 3), norm(1, 2));
        WarpJob& j = jobs[i];
        j.cw = cw;
        j.ch = ch;
        j.rot90 = (cw > 0 && ch > 0 && (ch * 1.0 / cw) >= 1.5) ? 1 : 0;
        j.out_off = off;
        if (cw <= 0 || ch <= 0) throw std::runtime_error("bad argument: degenerate box (zero-size crop) at index " + std::to_string(i));
        perspective_inverse(b, cw, ch, j.mi);
        crop_dims_out[2 * i] = j.rot90 ? cw : ch;
        crop_dims_out[2 * i + 1] = j.rot90 ? ch : cw;
        off += (size_t)cw * ch * 3;
        max_w = std::max(max_w, cw);
        max_h = std::max(max_h, ch);
    }
    if (!crops_out) return;
    if (off > crops_capacity) throw std::runtime_error("crop buffer too small: need " + std::to_string(off) + " bytes");
    if (n == 0) return;
    size_t img_bytes = (size_t)h * row_stride;
    uint8_t* d_img = (uint8_t*)d_img_.get(img_bytes);
    B200_CUDA_CHECK(cudaMemcpyAsync(d_img, bgr, img_bytes, cudaMemcpyHostToDevice, st_));
    for (auto& j : jobs) { j.img = d_img; j.h = h; j.w = w; j.stride = row_stride; }
    WarpJob* d_jobs = (WarpJob*)d_misc_.get(sizeof(WarpJob) * n);
    B200_CUDA_CHECK(cudaMemcpyAsync(d_jobs, jobs.data(), sizeof(WarpJob) * n, cudaMemcpyHostToDevice, st_));
    uint8_t* d_pool = (uint8_t*)d_crops_.get(off);
    dim3 block(32, 8);
    dim3 grid((unsigned)ceil_div(max_w, 32), (unsigned)ceil_div(max_h, 8), n);
    warp_cubic_kernel<<<grid, block, 0, st_>>>(d_jobs, cubic_tab_device(), d_pool);
    launches += 1;
    B200_CUDA_CHECK(cudaGetLastError());
    B200_CUDA_CHECK(cudaMemcpyAsync(crops_out, d_pool, off, cudaMemcpyDeviceToHost, st_));
    B200_CUDA_CHECK(cudaStreamSynchronize(st_));
}

// ------------------------------------------------------------------------------- cls / rec
namespace {
struct BatchPlan {
    std::vector<int> order;                     // argsort of aspect ratios (stable, like np.argsort on distinct keys)
};

// np.argsort(np.array(width_list)) — quicksort is not stable in numpy, but for equal keys the
// permutation only reorders crops inside a batch or across batches with identical ratio; the result
// per crop does not depend on it in the reference (each crop's output row is written back by index).
std::vector<int> aspect_order(const int* hs, const int* ws, int n) {
    std::vector<int> idx(n);
    std::iota(idx.begin(), idx.end(), 0);
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) {
        return (ws[a] / (double)hs[a]) < (ws[b] / (double)hs[b]);
    });
    return idx;
}

int resized_w_rule(int h, int w, int img_h, int img_w) {      // predict_rec.py:66-71 / predict_cls.py:25-30
    double ratio = w / (double)h;
    int c = (int)std::ceil(img_h * ratio);
    return c > img_w ? img_w : c;
}
}  // namespace

// uploads the crops (host pointers) into one device pool; returns per-crop device offsets
static std::vector<size_t> upload_crops(DevBuf& pool, PinBuf& pin, const uint8_t* const* crops, const int* hs, const int* ws,
                                        int n, cudaStream_t st, uint8_t** d_base) {
    std::vector<size_t> off(n + 1, 0);
    for (int i = 0; i < n; ++i) off[i + 1] = off[i] + (((size_t)hs[i] * ws[i] * 3 + 15) & ~(size_t)15);
    uint8_t* h = (uint8_t*)pin.get(off[n] + 16);
    for (int i = 0; i < n; ++i) memcpy(h + off[i], crops[i], (size_t)hs[i] * ws[i] * 3);
    uint8_t* d = (uint8_t*)pool.get(off[n] + 16);
    B200_CUDA_CHECK(cudaMemcpyAsync(d, h, off[n], cudaMemcpyHostToDevice, st));
    *d_base = d;
    return off;
}

void Stages::cls_crops(Model* cls, const uint8_t* const* crops, const int* hs, const int* ws, int n, int batch_num,
                       int32_t* labels_out, float* scores_out) {
    if (n == 0) return;
    if (batch_num <= 0) batch_num = 6;
    const int img_h = 48, img_w = 192;          // args.cls_image_shape "3, 48, 192" (utils.py:349)
    uint8_t* d_base = nullptr;
    std::vector<size_t> off = upload_crops(d_crops_, h_pin_, crops, hs, ws, n, st_, &d_base);
    std::vector<int> order = aspect_order(hs, ws, n);
    // every cls batch has the same canvas (48 x 192), so the whole set runs as ONE batch on the device;
    // per-crop results do not depend on batch composition (no cross-sample op in the graph).
    std::vector<ResizeJob> jobs(n);
    for (int k = 0; k < n; ++k) {
        int i = order[k];
        int rw = resized_w_rule(hs[i], ws[i], img_h, img_w);
        jobs[k] = make_job(d_base + off[i], hs[i], ws[i], ws[i] * 3, hs[i], ws[i], img_h, rw, 0);
    }
    ResizeJob* d_jobs = (ResizeJob*)d_misc_.get(sizeof(ResizeJob) * n);
    B200_CUDA_CHECK(cudaMemcpyAsync(d_jobs, jobs.data(), sizeof(ResizeJob) * n, cudaMemcpyHostToDevice, st_));
    const int chunk = 512;
    std::vector<float> probs;
    for (int beg = 0; beg < n; beg += chunk) {
        int b = std::min(chunk, n - beg);
        Plan* p = cls->plan_for(b, img_h, img_w);
        dim3 grid((unsigned)ceil_div(img_w, 256), img_h, b);
        if (cls->precision() == PREC_BF16)
            resize_norm_kernel<bf16><<<grid, 256, 0, st_>>>(d_jobs + beg, img_h, img_w, p->in_ptr(), 0, rec_norm());
        else
            resize_norm_kernel<float><<<grid, 256, 0, st_>>>(d_jobs + beg, img_h, img_w, p->in_ptr(), 0, rec_norm());
        p->launch(st_);
        launches += p->launches + 1;
        B200_CUDA_CHECK(cudaGetLastError());
        probs.resize(p->out_dense_elems);
        B200_CUDA_CHECK(cudaMemcpyAsync(probs.data(), p->out_dense, p->out_dense_elems * sizeof(float), cudaMemcpyDeviceToHost, st_));
        B200_CUDA_CHECK(cudaStreamSynchronize(st_));
        int ncls = (int)(p->out_dense_elems / b);
        for (int k = 0; k < b; ++k) {            // ClsPostProcess: argmax over classes (cls_postprocess.py:24-26)
            const float* r = probs.data() + (size_t)k * ncls;
            int best = 0;
            for (int c = 1; c < ncls; ++c)
                if (r[c] > r[best]) best = c;
            labels_out[order[beg + k]] = best;
            scores_out[order[beg + k]] = r[best];
        }
    }
}

// One device batch of the recogniser: resize/normalise/pad -> SVTR_LCNet -> greedy CTC decode into the slots.
// `lane` selects the stream (0 = the engine stream, 1.. = auxiliary streams): small batches get a plan with a private
// activation slab, so independent batches overlap on the GPU instead of queueing behind each other.
void Stages::rec_batch(Model* rec, const ResizeJob* d_jobs, const int* d_slots, int b, int img_h, int W, int max_len,
                       int* d_ids, int* d_lens, float* d_scores, int lane) {
    cudaStream_t st = lane == 0 ? st_ : aux_[lane - 1];
    Plan* p = rec->plan_for(b, img_h, W, /*head=*/true, /*own_arena=*/lane != 0);
    if (lane != 0 && !p->own_arena) {            // private budget exhausted: this batch must run on the engine stream
        lane = 0;
        st = st_;
    }
    TraceScope tr_rec("rec_batch", st, b * 10000 + W);
    dim3 grid((unsigned)ceil_div(W, 256), img_h, b);
    if (rec->precision() == PREC_BF16)
        resize_norm_kernel<bf16><<<grid, 256, 0, st>>>(d_jobs, img_h, W, p->in_ptr(), 0, rec_norm());
    else
        resize_norm_kernel<float><<<grid, 256, 0, st>>>(d_jobs, img_h, W, p->in_ptr(), 0, rec_norm());
    p->launch(st);
    if (p->out_shape.size() != 3) throw std::runtime_error("rec model output must be [B, T, classes]");
    int T = (int)p->out_shape[1], C = (int)p->out_shape[2];
    int rows = b * T;
    int* d_idx = (int*)d_lane_[lane].get((size_t)rows * (sizeof(int) + 2 * sizeof(float)));
    float* d_mx = (float*)(d_idx + rows);
    float* d_scr = d_mx + rows;
    if (p->head_rows > 0) {                      // fused CTC head: argmax / max came out of the GEMM epilogue
        ctc_collapse_kernel<<<(unsigned)ceil_div(b, 64), 64, 0, st>>>(p->head_idx(), p->head_prob(), b, T, d_slots, max_len,
                                                                       d_ids, d_lens, d_scores, d_scr);
        launches += p->launches + 2;
        B200_CUDA_CHECK(cudaGetLastError());
        return;
    }
    ctc_argmax_kernel<<<rows, 256, 0, st>>>(p->out_dense, C, d_idx, d_mx);
    ctc_collapse_kernel<<<(unsigned)ceil_div(b, 64), 64, 0, st>>>(d_idx, d_mx, b, T, d_slots, max_len, d_ids, d_lens,
                                                                   d_scores, d_scr);
    launches += p->launches + 3;
    B200_CUDA_CHECK(cudaGetLastError());
}

// Runs the rec batches [beg, end) of equal imgW (given in `imgw_sorted`) and spreads small ones over the auxiliary
// streams.  Everything enqueued on the engine stream so far is visible to the auxiliary streams, and the engine stream
// continues only after all of them have finished.
void Stages::rec_batches(Model* rec, const ResizeJob* d_jobs, const int* d_slots, const std::vector<int>& imgw_sorted,
                         int img_h, int max_len, int* d_ids, int* d_lens, float* d_scores) {
    const int total = (int)imgw_sorted.size();
    const int max_batch = 256, small_batch = 48;
    std::vector<std::pair<int, int>> batches;
    for (int beg = 0; beg < total;) {
        int end = beg;
        while (end < total && imgw_sorted[end] == imgw_sorted[beg] && end - beg < max_batch) ++end;
        batches.push_back({beg, end});
        beg = end;
    }
    int n_small = 0;
    for (auto& be : batches) n_small += (be.second - be.first) <= small_batch;
    const bool fan_out = n_small >= 2 && n_aux_ > 0;
    if (fan_out) {
        B200_CUDA_CHECK(cudaEventRecord(ev_fork_, st_));
        for (int a = 0; a < n_aux_; ++a) B200_CUDA_CHECK(cudaStreamWaitEvent(aux_[a], ev_fork_, 0));
    }
    std::vector<char> used(n_aux_, 0);
    // Small batches are spread over the auxiliary streams longest-first onto the least loaded stream (they are latency
    // bound: a batch of 6 crops costs about as much as one of 24).  A plan (b, W) that already ran in this call stays
    // on its stream: its private slab may still be in use by the previous wave's batch of the same shape, and stream
    // order is what protects it (rec_lane_of_ is cleared when a call starts, i.e. after the previous call drained).
    std::vector<int> lane_of(batches.size(), 0);
    if (fan_out) {
        std::vector<double> load(n_aux_, 0.0);
        auto cost = [&](int i) { return 1.0 + (double)(batches[i].second - batches[i].first) * imgw_sorted[batches[i].first] / 8000.0; };
        std::vector<int> order;
        for (int i = 0; i < (int)batches.size(); ++i) {
            int b = batches[i].second - batches[i].first, W = imgw_sorted[batches[i].first];
            if (b > small_batch) continue;
            auto it = rec_lane_of_.find({b, W});
            if (it != rec_lane_of_.end()) {
                lane_of[i] = it->second;
                load[it->second - 1] += cost(i);
            } else {
                order.push_back(i);
            }
        }
        std::stable_sort(order.begin(), order.end(), [&](int a, int c) { return cost(a) > cost(c); });
        for (int i : order) {
            int best = (int)(std::min_element(load.begin(), load.end()) - load.begin());
            load[best] += cost(i);
            lane_of[i] = best + 1;
            rec_lane_of_[{batches[i].second - batches[i].first, imgw_sorted[batches[i].first]}] = best + 1;
        }
    }
    for (int i = 0; i < (int)batches.size(); ++i) {
        auto& be = batches[i];
        int b = be.second - be.first, W = imgw_sorted[be.first];
        int lane = lane_of[i];
        if (lane > 0) used[lane - 1] = 1;
        rec_batch(rec, d_jobs + be.first, d_slots + be.first, b, img_h, W, max_len, d_ids, d_lens, d_scores, lane);
    }
    if (fan_out)
        for (int a = 0; a < n_aux_; ++a)
            if (used[a]) {
                B200_CUDA_CHECK(cudaEventRecord(ev_join_[a], aux_[a]));
                B200_CUDA_CHECK(cudaStreamWaitEvent(st_, ev_join_[a], 0));
            }
}

// session.run + `preds.argmax(axis=2)` / `preds.max(axis=2)` (predict_rec.py:316, rec_postprocess.py:100-101) for a
// caller-provided NCHW tensor: the per-step class index and its probability, without the [n, T, classes] tensor.
void Stages::rec_argmax_host(Model* rec, const float* nchw, int n, int c, int h, int w, int32_t* idx_out, float* prob_out,
                             int capacity, int* T_out) {
    Plan* p = rec->launch_host_input(nchw, n, c, h, w, st_, /*head=*/true);
    if (p->out_shape.size() != 3) throw std::runtime_error("rec model output must be [B, T, classes]");
    int T = (int)p->out_shape[1], C = (int)p->out_shape[2];
    int rows = n * T;
    if (T_out) *T_out = T;
    launches += p->launches + 1;
    if (!idx_out || !prob_out) {
        B200_CUDA_CHECK(cudaStreamSynchronize(st_));
        return;
    }
    if (rows > capacity) throw std::runtime_error("output buffer too small: need " + std::to_string(rows) + " rows");
    const int* d_idx;
    const float* d_mx;
    if (p->head_rows > 0) {
        d_idx = p->head_idx();
        d_mx = p->head_prob();
    } else {
        int* di = (int*)d_batch_.get((size_t)rows * (sizeof(int) + sizeof(float)));
        float* dm = (float*)(di + rows);
        ctc_argmax_kernel<<<rows, 256, 0, st_>>>(p->out_dense, C, di, dm);
        launches += 1;
        d_idx = di;
        d_mx = dm;
    }
    B200_CUDA_CHECK(cudaMemcpyAsync(idx_out, d_idx, (size_t)rows * sizeof(int), cudaMemcpyDeviceToHost, st_));
    B200_CUDA_CHECK(cudaMemcpyAsync(prob_out, d_mx, (size_t)rows * sizeof(float), cudaMemcpyDeviceToHost, st_));
    B200_CUDA_CHECK(cudaStreamSynchronize(st_));
}

void Stages::rec_crops(Model* rec, const uint8_t* const* crops, const int* hs, const int* ws, int n, int batch_num,
                       const uint8_t* rotate180, int32_t* ids_out, int max_len, int* lens_out, float* scores_out) {
    if (n == 0) return;
    if (batch_num <= 0) batch_num = 6;
    rec_lane_of_.clear();                       // the previous call drained every stream
    const int img_h = 48, img_w0 = 320;         // rec_image_shape "3, 48, 320" (onnx_paddleocr.py:20)
    uint8_t* d_base = nullptr;
    std::vector<size_t> off = upload_crops(d_crops_, h_pin_, crops, hs, ws, n, st_, &d_base);
    std::vector<int> order = aspect_order(hs, ws, n);
    // reference batches: consecutive groups of batch_num in aspect order, imgW = int(48 * max ratio) per group
    // (predict_rec.py:290-300).  Groups with the same imgW are merged into one device batch.
    std::vector<int> imgw(n);
    for (int beg = 0; beg < n; beg += batch_num) {
        int end = std::min(n, beg + batch_num);
        double max_wh = img_w0 / (double)img_h;
        for (int k = beg; k < end; ++k) max_wh = std::max(max_wh, ws[order[k]] * 1.0 / hs[order[k]]);
        int W = (int)(img_h * max_wh);
        for (int k = beg; k < end; ++k) imgw[k] = W;
    }
    std::vector<ResizeJob> jobs(n);
    for (int k = 0; k < n; ++k) {
        int i = order[k];
        int rw = resized_w_rule(hs[i], ws[i], img_h, imgw[k]);
        jobs[k] = make_job(d_base + off[i], hs[i], ws[i], ws[i] * 3, hs[i], ws[i], img_h, rw, rotate180 ? rotate180[i] != 0 : 0);
    }
    ResizeJob* d_jobs = (ResizeJob*)d_misc_.get(sizeof(ResizeJob) * n + sizeof(int) * n);
    int* d_slots = (int*)((char*)d_jobs + sizeof(ResizeJob) * n);
    B200_CUDA_CHECK(cudaMemcpyAsync(d_jobs, jobs.data(), sizeof(ResizeJob) * n, cudaMemcpyHostToDevice, st_));
    B200_CUDA_CHECK(cudaMemcpyAsync(d_slots, order.data(), sizeof(int) * n, cudaMemcpyHostToDevice, st_));
    int* d_ids = (int*)d_boxes_.get(((size_t)n * max_len + 2 * (size_t)n) * sizeof(int));
    int* d_lens = d_ids + (size_t)n * max_len;
    float* d_scores = (float*)(d_lens + n);
    B200_CUDA_CHECK(cudaMemsetAsync(d_ids, 0, (size_t)n * max_len * sizeof(int), st_));
    rec_batches(rec, d_jobs, d_slots, imgw, img_h, max_len, d_ids, d_lens, d_scores);
    B200_CUDA_CHECK(cudaMemcpyAsync(ids_out, d_ids, (size_t)n * max_len * sizeof(int), cudaMemcpyDeviceToHost, st_));
    B200_CUDA_CHECK(cudaMemcpyAsync(lens_out, d_lens, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st_));
    B200_CUDA_CHECK(cudaMemcpyAsync(scores_out, d_scores, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, st_));
    B200_CUDA_CHECK(cudaStreamSynchronize(st_));
}

// ------------------------------------------------------------------------------- det pages
// filter_tag_det_res (predict_det.py:67-81) on int32 boxes in cv2 order -> float32 boxes appended to `out`
static int filter_boxes(const int32_t* in, int n, int img_h, int img_w, std::vector<float>* out) {
    int m = 0;
    for (int i = 0; i < n; ++i) {
        float pts[4][2];
        for (int k = 0; k < 4; ++k) {
            pts[k][0] = (float)in[i * 8 + 2 * k];
            pts[k][1] = (float)in[i * 8 + 2 * k + 1];
        }
        // order_points_clockwise (predict_det.py:50-59): np.argmin/argmax return the first extremum
        float s[4];
        for (int k = 0; k < 4; ++k) s[k] = pts[k][0] + pts[k][1];
        int imin = 0, imax = 0;
        for (int k = 1; k < 4; ++k) {
            if (s[k] < s[imin]) imin = k;
            if (s[k] > s[imax]) imax = k;
        }
        float rect[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
        rect[0][0] = pts[imin][0]; rect[0][1] = pts[imin][1];
        rect[2][0] = pts[imax][0]; rect[2][1] = pts[imax][1];
        int rest[4], nr = 0;
        for (int k = 0; k < 4; ++k)
            if (k != imin && k != imax) rest[nr++] = k;        // np.delete keeps 3 rows when imin == imax
        float d[4];
        for (int k = 0; k < nr; ++k) d[k] = pts[rest[k]][1] - pts[rest[k]][0];
        int dmin = 0, dmax = 0;
        for (int k = 1; k < nr; ++k) {
            if (d[k] < d[dmin]) dmin = k;
            if (d[k] > d[dmax]) dmax = k;
        }
        rect[1][0] = pts[rest[dmin]][0]; rect[1][1] = pts[rest[dmin]][1];
        rect[3][0] = pts[rest[dmax]][0]; rect[3][1] = pts[rest[dmax]][1];
        // clip_det_res (predict_det.py:61-65)
        for (int k = 0; k < 4; ++k) {
            rect[k][0] = (float)(int)std::min(std::max(rect[k][0], 0.f), (float)(img_w - 1));
            rect[k][1] = (float)(int)std::min(std::max(rect[k][1], 0.f), (float)(img_h - 1));
        }
        auto norm = [&](int a, int b) {
            float dx = rect[a][0] - rect[b][0], dy = rect[a][1] - rect[b][1];
            float q = dx * dx;
            q += dy * dy;
            return std::sqrt(q);
        };
        int rw = (int)norm(0, 1), rh = (int)norm(0, 3);
        if (rw <= 3 || rh <= 3) continue;
        out->insert(out->end(), &rect[0][0], &rect[0][0] + 8);
        ++m;
    }
    return m;
}

// Host -> device copy of one page.  Pageable sources (plain numpy arrays) are first copied into the engine's pinned
// staging area by several host threads (a single-threaded pageable cudaMemcpy tops out near 6 GB/s, far below PCIe 5),
// then DMA'd asynchronously, so the staging copy of the next page overlaps the DMA of this one.  Pinned sources
// (b200ocr_host_alloc) go straight to the DMA engine.
void Stages::upload_page(uint8_t* dst, const uint8_t* src, size_t bytes, cudaStream_t st, PageStore* store) {
    cudaPointerAttributes attr;
    bool pinned = cudaPointerGetAttributes(&attr, src) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (pinned || bytes < ((size_t)1 << 20) || store->used + bytes > store->cap) {
        B200_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
        return;
    }
    uint8_t* hp = (uint8_t*)store->h.p + store->used;
    store->used += (bytes + 4095) & ~(size_t)4095;
    static const int n_threads = [] {
        unsigned hc = std::thread::hardware_concurrency();
        return (int)std::max(1u, std::min(8u, hc / 2));
    }();
    const size_t slice = ((bytes + n_threads - 1) / n_threads + 4095) & ~(size_t)4095;
    std::vector<std::thread> workers;
    for (int t = 1; t < n_threads; ++t) {
        size_t b0 = (size_t)t * slice;
        if (b0 >= bytes) break;
        size_t len = std::min(slice, bytes - b0);
        workers.emplace_back([=] { memcpy(hp + b0, src + b0, len); });
    }
    memcpy(hp, src, std::min(slice, bytes));
    for (auto& w : workers) w.join();
    B200_CUDA_CHECK(cudaMemcpyAsync(dst, hp, bytes, cudaMemcpyHostToDevice, st));
}

// All pages of a call go to the device once and stay there until their crops are cut (the warp samples the
// full-resolution original, predict_system.py:36,51).  plan_pool lays the pool out; det_enqueue uploads each page right
// before the sub-batch that needs it is enqueued, so the copy of the next sub-batch overlaps the network of this one.
void Stages::plan_pool(const int* hs, const int* ws, const int* strides, int n, PagePool* pool, PageStore* store) {
    pool->off.assign(n, 0);
    pool->stride.assign(n, 0);
    size_t total = 0;
    for (int i = 0; i < n; ++i) {
        if (hs[i] <= 0 || ws[i] <= 0) throw std::runtime_error("bad argument: empty page at index " + std::to_string(i));
        pool->stride[i] = strides && strides[i] > 0 ? strides[i] : ws[i] * 3;
        if (pool->stride[i] < ws[i] * 3) throw std::runtime_error("bad argument: row stride smaller than a row at page " + std::to_string(i));
        pool->off[i] = total;
        total += ((size_t)hs[i] * pool->stride[i] + 255) & ~(size_t)255;
    }
    pool->base = (uint8_t*)store->d.get(total);
    // pinned staging for pageable sources: the previous user of this store has completed (every synchronous entry point
    // drains its streams, a streaming slot is reused only after its batch was collected), so the area is free; bounded
    // so that a huge batch falls back to direct copies for its tail
    store->cap = std::min<size_t>(total + (size_t)n * 4096, (size_t)2 << 30);
    store->h.get(store->cap);
    store->used = 0;
}

// One lane's sub-batch: boxes to the host (the only host synchronisation of the det path), reference tail on the host.
void Stages::det_collect_lane(DetTicket* t, int k) {
    if (!t->pend[k].active) return;
    std::vector<int32_t> boxes;
    std::vector<double> scores;
    std::vector<int> counts;
    db_post_collect(&t->pend[k], &boxes, &scores, &counts);
    const DetChunk& c = t->pend_chunk[k];
    size_t bo = 0;
    for (int j = 0; j < c.b; ++j) {
        int i = t->idx[c.g0 + j];
        filter_boxes(boxes.data() + bo * 8, counts[j], c.h, c.w, &t->boxes[i]);
        bo += counts[j];
    }
}

void Stages::det_finish(DetTicket* t) {
    for (int k = 0; k < kLanesPerSlot; ++k) det_collect_lane(t, (t->turn + k) % kLanesPerSlot);   // oldest first
}

// A batch failed while it was being enqueued (bad argument, CUDA error): wait for whatever it already put on its lanes and
// drop the pending DBPostProcess state, so that the next user of the slot does not collect a stale sub-batch.
void Stages::det_abort(DetTicket* t) {
    for (int k = 0; k < kLanesPerSlot; ++k) {
        if (!t->pend[k].active) continue;
        t->pend[k].active = false;
        cudaStreamSynchronize(det_st_[t->pend[k].lane]);
    }
    cudaGetLastError();
    t->boxes.clear();
}

void Stages::drain_tickets() {
    for (auto& t : tickets_)
        if (t.busy) det_finish(&t);
}

// TextDetector.__call__ for a batch of pages: equal-shaped pages run as sub-batches of up to 8 that alternate between
// the two lanes of the ticket's slot (stream + activation slab + DBPostProcess scratch each).  A lane's boxes are
// collected just before the lane is reused, i.e. while the other lane's network is executing; whatever is still pending
// when this returns is collected by det_finish.  t->boxes[i] receives page i's filtered float32 quads (cv2 contour
// order), 8 floats each.
void Stages::det_enqueue(DetTicket* t, Model* det, const PagePool& pool, const uint8_t* const* pages, const int* hs,
                         const int* ws, int n, const b200ocr_det_params& p, float* probmap_out, size_t probmap_capacity,
                         int* probmap_dims_out, cudaStream_t resident_src) {
    t->n = n;
    t->turn = 0;
    t->boxes.assign(n, {});
    std::vector<int>& idx = t->idx;
    idx.resize(n);
    std::iota(idx.begin(), idx.end(), 0);
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) {
        return hs[a] != hs[b] ? hs[a] < hs[b] : ws[a] < ws[b];
    });
    std::vector<size_t> prob_off(n + 1, 0);
    std::vector<DetGeom> geo(n);
    for (int i = 0; i < n; ++i) {
        geo[i] = det_geometry(hs[i], ws[i], p);
        prob_off[i + 1] = prob_off[i] + (size_t)geo[i].dst_h * geo[i].dst_w;
        if (probmap_dims_out) {
            probmap_dims_out[2 * i] = geo[i].dst_h;
            probmap_dims_out[2 * i + 1] = geo[i].dst_w;
        }
    }
    if (probmap_out && prob_off[n] > probmap_capacity) throw std::runtime_error("probability-map buffer too small");
    const int lane0 = t->slot * kLanesPerSlot;
    // resident pages were uploaded on another stream: make them visible to the lanes.  When the lanes upload the
    // pages themselves nothing on the engine stream concerns them, and NOT waiting lets this detection overlap the
    // recogniser work of the previous wave that is still queued there.
    if (!pages && resident_src) {
        B200_CUDA_CHECK(cudaEventRecord(ev_fork_, resident_src));
        for (int l = 0; l < kLanesPerSlot; ++l) B200_CUDA_CHECK(cudaStreamWaitEvent(det_st_[lane0 + l], ev_fork_, 0));
    }

    // A synchronous call splits 16 equal pages into 8 + 8 on its two lanes (DBPostProcess of one half overlaps the network
    // of the other); a streaming batch keeps them together (one 16-page plan is 13 % cheaper than two 8-page plans, and
    // the overlap comes from the other batches in flight).  B200OCR_DET_SUB overrides both.
    static const int sub_env = [] {
        const char* e = getenv("B200OCR_DET_SUB");
        int v = e ? atoi(e) : 0;
        return v >= 1 && v <= 16 ? v : 0;
    }();
    const int sub = sub_env ? sub_env : (t == &sync_ticket_ ? 8 : 16);
    const int max_group = 16;
    for (int g0 = 0; g0 < n;) {
        int g1 = g0;
        while (g1 < n && hs[idx[g1]] == hs[idx[g0]] && ws[idx[g1]] == ws[idx[g0]] && g1 - g0 < max_group) ++g1;
        const int group = g1 - g0;
        const int step = group > sub + sub / 2 ? sub : group;          // 16 pages -> 8 + 8; up to 12 stay together
        for (int c0 = g0; c0 < g1; c0 += step) {
            const int b = std::min(step, g1 - c0);
            const int k = t->turn++ % kLanesPerSlot;
            const int lane = lane0 + k;
            det_collect_lane(t, k);                                   // frees the lane's scratch, slab and job table
            cudaStream_t st = det_st_[lane];
            const DetGeom& g = geo[idx[c0]];
            const int h = hs[idx[c0]], w = ws[idx[c0]];
            std::vector<ResizeJob> jobs(b);
            for (int j = 0; j < b; ++j) {
                int i = idx[c0 + j];
                if (pages) {
                    if (!pages[i]) throw std::runtime_error("bad argument: null page at index " + std::to_string(i));
                    upload_page(pool.base + pool.off[i], pages[i], (size_t)h * pool.stride[i], st, &t->store);
                }
                jobs[j] = make_job(pool.base + pool.off[i], h, w, pool.stride[i], g.pad_h, g.pad_w, g.dst_h, g.dst_w, 0);
            }
            Plan* pl = det->plan_for(b, g.dst_h, g.dst_w, false, /*own_arena=*/2, /*slot=*/lane, det->fused_pre_default);
            // the lane's previous sub-batch was collected above, so the plan's job table / the lane's table is free
            ResizeJob* d_jobs = pl->pre_jobs ? (ResizeJob*)pl->pre_jobs : (ResizeJob*)d_lane_jobs_[lane].get(sizeof(ResizeJob) * max_group);
            B200_CUDA_CHECK(cudaMemcpyAsync(d_jobs, jobs.data(), sizeof(ResizeJob) * b, cudaMemcpyHostToDevice, st));
            if (!pl->own_arena) {
                // The private-slab budget is exhausted and this plan lives in the model's shared slab: it must not run
                // while anything else may be using that slab (another lane's network, or the DBPostProcess tail of
                // another sub-batch, which reads the probability map out of the slab).  Finish everything else first; a
                // later shared-slab chunk drains this one the same way, so shared-slab chunks run strictly one at a time.
                for (int l = 0; l < kLanesPerSlot; ++l)
                    if (l != k) det_collect_lane(t, l);
                for (auto& o : tickets_)
                    if (&o != t && o.busy) det_finish(&o);
                if (t != &sync_ticket_) det_finish(&sync_ticket_);
                for (int l = 0; l < kDetLanes; ++l)
                    if (l != lane) B200_CUDA_CHECK(cudaStreamSynchronize(det_st_[l]));
            }
            TraceScope tr_net("det_net", st, b);
            if (!pl->pre_jobs) {
                dim3 grid((unsigned)ceil_div(g.dst_w, 256), g.dst_h, b);
                if (det->precision() == PREC_BF16)
                    resize_norm_kernel<bf16><<<grid, 256, 0, st>>>(d_jobs, g.dst_h, g.dst_w, pl->in_ptr(), 0, det_norm());
                else
                    resize_norm_kernel<float><<<grid, 256, 0, st>>>(d_jobs, g.dst_h, g.dst_w, pl->in_ptr(), 0, det_norm());
                launches += 1;
            }
            // Rolling window: this network starts when the network enqueued `window` chunks earlier has finished, so at
            // most `window` networks run at a time and the lanes can never fall into lock step (all networks together,
            // then all the latency-bound DBPostProcess tails together on an idle GPU).  The uploads above are not held
            // back.  A chunk that reuses the lane of chunk c - window is ordered by the stream anyway.
            static const int window = [] {
                const char* e = getenv("B200OCR_DET_WINDOW");
                int v = e ? atoi(e) : 2;
                return v >= 0 && v < kDetLanes ? v : 2;
            }();
            if (window > 0 && net_seq_ >= window) {
                const int prev_lane = net_lane_ring_[(net_seq_ - window) % kDetLanes];
                if (prev_lane != lane) B200_CUDA_CHECK(cudaStreamWaitEvent(st, ev_net_done_[prev_lane], 0));
            }
            pl->launch(st);
            B200_CUDA_CHECK(cudaEventRecord(ev_net_done_[lane], st));
            net_lane_ring_[net_seq_ % kDetLanes] = lane;
            ++net_seq_;
            launches += pl->launches;
            B200_CUDA_CHECK(cudaGetLastError());
            // the dense output is [b,1,H,W] fp32 == the probability maps
            if (probmap_out)
                for (int j = 0; j < b; ++j)
                    B200_CUDA_CHECK(cudaMemcpyAsync(probmap_out + prob_off[idx[c0 + j]],
                                                    pl->out_dense + (size_t)j * g.dst_h * g.dst_w,
                                                    (size_t)g.dst_h * g.dst_w * sizeof(float), cudaMemcpyDeviceToHost, st));
            std::vector<int> sh(b, h), sw(b, w);
            db_post_enqueue(lane, pl->out_dense, b, g.dst_h, g.dst_w, sh.data(), sw.data(), p, p.max_candidates, &t->pend[k]);
            t->pend_chunk[k] = {c0, b, h, w};
        }
        g0 = g1;
    }
}

// The synchronous form used by det_pages / ocr_pages / resident_run: enqueue on the lanes of slot 0, collect everything.
void Stages::det_on_pool(Model* det, const PagePool& pool, const uint8_t* const* pages, const int* hs, const int* ws, int n,
                         const b200ocr_det_params& p, std::vector<std::vector<float>>* boxes_pp, float* probmap_out,
                         size_t probmap_capacity, int* probmap_dims_out) {
    drain_tickets();                                   // streaming batches may own the lanes of slot 0
    DetTicket* t = &sync_ticket_;
    t->slot = 0;
    try {
        det_enqueue(t, det, pool, pages, hs, ws, n, p, probmap_out, probmap_capacity, probmap_dims_out, st_);
    } catch (...) {
        det_abort(t);
        throw;
    }
    det_finish(t);
    boxes_pp->swap(t->boxes);
    // later work on the engine stream (crops, cls, rec) reads the pages the lanes uploaded
    for (int l = 0; l < kLanesPerSlot; ++l) {
        B200_CUDA_CHECK(cudaEventRecord(ev_det_done_[l], det_st_[l]));
        B200_CUDA_CHECK(cudaStreamWaitEvent(st_, ev_det_done_[l], 0));
    }
    for (float& v : stage_ms) v = 0.f;
}

void Stages::det_pages(Model* det, const uint8_t* const* pages, const int* hs, const int* ws, const int* strides, int n,
                       const b200ocr_det_params& p, float* boxes_out, int max_boxes, int* n_boxes_out, float* probmap_out,
                       size_t probmap_capacity, int* probmap_dims_out) {
    if (n == 0) return;
    drain_tickets();
    PagePool pool;
    plan_pool(hs, ws, strides, n, &pool, &sync_ticket_.store);
    std::vector<std::vector<float>> boxes;
    det_on_pool(det, pool, pages, hs, ws, n, p, &boxes, probmap_out, probmap_capacity, probmap_dims_out);
    B200_CUDA_CHECK(cudaStreamSynchronize(st_));
    for (int i = 0; i < n; ++i) {
        int m = (int)(boxes[i].size() / 8);
        if (m > max_boxes) throw std::runtime_error("box capacity too small: page " + std::to_string(i) + " has " + std::to_string(m));
        n_boxes_out[i] = m;
        if (m) memcpy(boxes_out + (size_t)i * max_boxes * 8, boxes[i].data(), boxes[i].size() * sizeof(float));
    }
}

// ---- streaming detection
int Stages::det_submit(Model* det, const uint8_t* const* pages, const int* hs, const int* ws, const int* strides, int n,
                       const b200ocr_det_params& p) {
    DetTicket* t = nullptr;
    for (auto& c : tickets_)
        if (!c.busy) { t = &c; break; }
    if (!t) throw std::runtime_error("capacity: " + std::to_string(kDetSlots) + " det batches are already in flight (collect one first)");
    if (sync_ticket_.busy) throw std::runtime_error("internal: synchronous det call in progress");
    t->slot = (int)(t - tickets_);
    t->id = ++ticket_seq_ * kDetSlots + t->slot;
    PagePool pool;
    plan_pool(hs, ws, strides, n, &pool, &t->store);
    try {
        det_enqueue(t, det, pool, pages, hs, ws, n, p, nullptr, 0, nullptr, nullptr);
    } catch (...) {
        det_abort(t);                              // the slot stays free; nothing of this batch is left on its lanes
        throw;
    }
    t->busy = true;
    return t->id;
}

void Stages::det_collect(int ticket, float* boxes_out, int max_boxes, int* n_boxes_out) {
    DetTicket* t = ticket >= 0 ? &tickets_[ticket % kDetSlots] : nullptr;
    if (!t || !t->busy || t->id != ticket) throw std::runtime_error("bad argument: unknown or already collected det ticket");
    det_finish(t);
    // a too small box buffer leaves the ticket collectable (its results are complete): the caller retries with more room
    for (int i = 0; i < t->n; ++i) {
        int m = (int)(t->boxes[i].size() / 8);
        if (m > max_boxes) throw std::runtime_error("box capacity too small: page " + std::to_string(i) + " has " + std::to_string(m));
    }
    t->busy = false;
    for (int i = 0; i < t->n; ++i) {
        int m = (int)(t->boxes[i].size() / 8);
        n_boxes_out[i] = m;
        if (m) memcpy(boxes_out + (size_t)i * max_boxes * 8, t->boxes[i].data(), t->boxes[i].size() * sizeof(float));
    }
}

// ------------------------------------------------------------------------------- fused end-to-end (Level C)
// sorted_boxes (predict_system.py:75-97): Python's stable sort on (y0, x0), then the non-transitive adjacent-swap pass.
static void sorted_boxes(std::vector<float>* boxes) {
    int n = (int)(boxes->size() / 8);
    if (n < 2) return;
    std::vector<int> order(n);
    std::iota(order.begin(), order.end(), 0);
    const float* b = boxes->data();
    std::stable_sort(order.begin(), order.end(), [&](int a, int c) {
        if (b[a * 8 + 1] != b[c * 8 + 1]) return b[a * 8 + 1] < b[c * 8 + 1];
        return b[a * 8] < b[c * 8];
    });
    std::vector<float> t(boxes->size());
    for (int i = 0; i < n; ++i) memcpy(&t[(size_t)i * 8], b + (size_t)order[i] * 8, 8 * sizeof(float));
    for (int i = 0; i < n - 1; ++i)
        for (int j = i; j >= 0; --j) {
            float* lo = &t[(size_t)j * 8];
            float* hi = &t[(size_t)(j + 1) * 8];
            if (std::fabs(hi[1] - lo[1]) < 10.f && hi[0] < lo[0]) {
                float tmp[8];
                memcpy(tmp, lo, sizeof(tmp)); memcpy(lo, hi, sizeof(tmp)); memcpy(hi, tmp, sizeof(tmp));
            } else {
                break;
            }
        }
    boxes->swap(t);
}

static bool make_warp_job(const float* b, WarpJob* j) {
    auto norm = [&](int a, int c) {
        float dx = b[2 * a] - b[2 * c], dy = b[2 * a + 1] - b[2 * c + 1];
        float s = dx * dx;                       // np.linalg.norm on float32
        s += dy * dy;
        return std::sqrt(s);
    };
    int cw = (int)std::max(norm(0, 1), norm(2, 3));
    int ch = (int)std::max(norm(0, 3), norm(1, 2));
    j->cw = cw;
    j->ch = ch;
    j->rot90 = (cw > 0 && ch > 0 && (ch * 1.0 / cw) >= 1.5) ? 1 : 0;
    if (cw <= 0 || ch <= 0) return false;
    perspective_inverse(b, cw, ch, j->mi);
    return true;
}

// TextSystem.__call__ (predict_system.py:35-72) for a batch of pages without leaving the device between stages:
// det -> boxes (host: filter + sorted_boxes, a few microseconds) -> one warp launch for every crop of every page ->
// cls over all crops -> rotation flags patched into the rec jobs on the device -> rec batches (the reference's per-page
// batch plan decides each crop's imgW; crops of all pages with equal imgW share a launch) -> CTC decode.
void Stages::ocr_pages(Model* det, Model* cls, Model* rec, const uint8_t* const* pages, const int* hs, const int* ws,
                       const int* strides, int n, const b200ocr_ocr_params& op, float* boxes_out, int line_cap,
                       int* n_boxes_out, int32_t* ids_out, int max_len, int* lens_out, float* scores_out,
                       int32_t* cls_labels_out, float* cls_scores_out) {
    if (n == 0) return;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms_since = [&](std::chrono::steady_clock::time_point t0) { return std::chrono::duration<float, std::milli>(now() - t0).count(); };
    for (float& v : ocr_ms) v = 0.f;
    auto t0 = now();
    PagePool pool;
    drain_tickets();
    rec_lane_of_.clear();                       // the previous call drained every stream
    plan_pool(hs, ws, strides, n, &pool, &sync_ticket_.store);
    ocr_ms[0] = ms_since(t0);
    const bool use_cls = op.use_angle_cls && cls != nullptr;
    const int img_h = 48, img_w0 = 320;
    const int batch_num = op.rec_batch_num > 0 ? op.rec_batch_num : 6;
    // results of all waves land in pinned memory (an asynchronous copy into the caller's pageable arrays would block the
    // host and with it the next wave's detection); copied out after the final drain
    const size_t res_ints = (size_t)line_cap * ((size_t)max_len + 4);
    int* h_res = (int*)h_pin2_.get(res_ints * sizeof(int));
    int* h_ids = h_res;
    int* h_lens = h_ids + (size_t)line_cap * max_len;
    float* h_scores = (float*)(h_lens + line_cap);
    int* h_clabel = (int*)(h_scores + line_cap);
    float* h_cscore = (float*)(h_clabel + line_cap);
    // Waves: while the recogniser of wave k runs on the engine / auxiliary streams, the host is already driving the
    // upload + detection of wave k+1 on the det lanes.  Measured: with ordinary pages the GPU is already saturated by
    // one wave (C4: 444 pages/s as one wave vs 386 as four — smaller det batches cost more than the overlap gains), so
    // waves are used only when the page uploads are the long pole (A4 @ 300 dpi: 26 MB per page, +4 %).
    size_t page_bytes = 0;
    for (int i = 0; i < n; ++i) page_bytes += (size_t)hs[i] * pool.stride[i];
    const int wave = (page_bytes / (size_t)n > ((size_t)16 << 20) && n >= 8) ? std::max(4, (n + 1) / 2) : n;   // A4 @ 300 dpi class pages only
    int line0 = 0;
    for (int p0 = 0; p0 < n; p0 += wave) {
        const int p1 = std::min(n, p0 + wave), np = p1 - p0;
        t0 = now();
        PagePool sub;
        sub.base = pool.base;
        sub.off.assign(pool.off.begin() + p0, pool.off.begin() + p1);
        sub.stride.assign(pool.stride.begin() + p0, pool.stride.begin() + p1);
        std::vector<std::vector<float>> boxes;
        det_on_pool(det, sub, pages + p0, hs + p0, ws + p0, np, op.det, &boxes, nullptr, 0, nullptr);
        ocr_ms[1] += ms_since(t0); t0 = now();
        int total = 0;
        for (int i = 0; i < np; ++i) {
            sorted_boxes(&boxes[i]);
            n_boxes_out[p0 + i] = (int)(boxes[i].size() / 8);
            total += n_boxes_out[p0 + i];
        }
        if (line0 + total > line_cap) {
            cudaStreamSynchronize(st_);
            throw std::runtime_error("line capacity too small: need more than " + std::to_string(line0 + total));
        }
        if (total == 0) continue;
        // ---- crops: get_rotate_crop_image for every box of every page of the wave in one launch
        std::vector<WarpJob> wj(total);
        std::vector<int> ch(total), cw(total), first_line(np + 1, 0);
        std::vector<size_t> coff(total + 1, 0);
        int max_w = 1, max_h = 1, li = 0;
        for (int i = 0; i < np; ++i) {
            first_line[i] = li;
            for (int k = 0; k < n_boxes_out[p0 + i]; ++k, ++li) {
                const float* b = &boxes[i][(size_t)k * 8];
                memcpy(boxes_out + (size_t)(line0 + li) * 8, b, 8 * sizeof(float));
                WarpJob& j = wj[li];
                if (!make_warp_job(b, &j)) {
                    cudaStreamSynchronize(st_);
                    throw std::runtime_error("degenerate box (zero-size crop) on page " + std::to_string(p0 + i));
                }
                j.img = pool.base + pool.off[p0 + i];
                j.h = hs[p0 + i]; j.w = ws[p0 + i]; j.stride = pool.stride[p0 + i];
                j.out_off = coff[li];
                coff[li + 1] = coff[li] + (((size_t)j.cw * j.ch * 3 + 15) & ~(size_t)15);
                ch[li] = j.rot90 ? j.cw : j.ch;
                cw[li] = j.rot90 ? j.ch : j.cw;
                max_w = std::max(max_w, j.cw);
                max_h = std::max(max_h, j.ch);
            }
        }
        first_line[np] = li;
        ocr_ms[2] += ms_since(t0); t0 = now();
        // one device block for every job table of this wave (reused by the next wave in stream order)
        auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
        size_t o_wj = 0, o_cj = o_wj + al(sizeof(WarpJob) * total), o_rj = o_cj + al(sizeof(ResizeJob) * total),
               o_cs = o_rj + al(sizeof(ResizeJob) * total), o_rs = o_cs + al(sizeof(int) * total),
               o_cl = o_rs + al(sizeof(int) * total), o_cc = o_cl + al(sizeof(int) * total), o_end = o_cc + al(sizeof(float) * total);
        char* d_tab = (char*)d_jobs_.get(o_end);
        WarpJob* d_wj = (WarpJob*)(d_tab + o_wj);
        ResizeJob* d_cj = (ResizeJob*)(d_tab + o_cj);
        ResizeJob* d_rj = (ResizeJob*)(d_tab + o_rj);
        int* d_cslot = (int*)(d_tab + o_cs);
        int* d_rslot = (int*)(d_tab + o_rs);
        int* d_clabel = (int*)(d_tab + o_cl);
        float* d_cscore = (float*)(d_tab + o_cc);
        uint8_t* d_pool = (uint8_t*)d_crops_.get(coff[total] + 16);
        B200_CUDA_CHECK(cudaMemcpyAsync(d_wj, wj.data(), sizeof(WarpJob) * total, cudaMemcpyHostToDevice, st_));
        {
            TraceScope tr_warp("warp", st_, total);
            dim3 block(32, 8);
            dim3 grid((unsigned)ceil_div(max_w, 32), (unsigned)ceil_div(max_h, 8), total);
            warp_cubic_kernel<<<grid, block, 0, st_>>>(d_wj, cubic_tab_device(), d_pool);
            launches += 1;
            B200_CUDA_CHECK(cudaGetLastError());
        }
        // ---- cls: every batch has the same 48 x 192 canvas, so all crops of the wave form one device batch
        if (use_cls) {
            const int cimg_w = 192;
            std::vector<int> order = aspect_order(ch.data(), cw.data(), total);
            std::vector<ResizeJob> jobs(total);
            for (int k = 0; k < total; ++k) {
                int i = order[k];
                int rw = resized_w_rule(ch[i], cw[i], img_h, cimg_w);
                jobs[k] = make_job(d_pool + coff[i], ch[i], cw[i], cw[i] * 3, ch[i], cw[i], img_h, rw, 0);
            }
            B200_CUDA_CHECK(cudaMemcpyAsync(d_cj, jobs.data(), sizeof(ResizeJob) * total, cudaMemcpyHostToDevice, st_));
            B200_CUDA_CHECK(cudaMemcpyAsync(d_cslot, order.data(), sizeof(int) * total, cudaMemcpyHostToDevice, st_));
            const int chunk = 512;
            for (int beg = 0; beg < total; beg += chunk) {
                int b = std::min(chunk, total - beg);
                TraceScope tr_cls("cls", st_, b);
                Plan* p = cls->plan_for(b, img_h, cimg_w);
                dim3 grid((unsigned)ceil_div(cimg_w, 256), img_h, b);
                if (cls->precision() == PREC_BF16)
                    resize_norm_kernel<bf16><<<grid, 256, 0, st_>>>(d_cj + beg, img_h, cimg_w, p->in_ptr(), 0, rec_norm());
                else
                    resize_norm_kernel<float><<<grid, 256, 0, st_>>>(d_cj + beg, img_h, cimg_w, p->in_ptr(), 0, rec_norm());
                p->launch(st_);
                int ncls = (int)(p->out_dense_elems / b);
                cls_argmax_kernel<<<(unsigned)ceil_div(b, 128), 128, 0, st_>>>(p->out_dense, b, ncls, d_cslot + beg, d_clabel, d_cscore);
                launches += p->launches + 2;
                B200_CUDA_CHECK(cudaGetLastError());
            }
        }
        ocr_ms[3] += ms_since(t0); t0 = now();
        // ---- rec: the reference's plan per page (aspect order, groups of rec_batch_num, imgW = int(48 * max ratio))
        std::vector<int> imgw(total);
        for (int pg = 0; pg < np; ++pg) {
            int l0 = first_line[pg], m = first_line[pg + 1] - l0;
            if (m == 0) continue;
            std::vector<int> order = aspect_order(ch.data() + l0, cw.data() + l0, m);
            for (int beg = 0; beg < m; beg += batch_num) {
                int end = std::min(m, beg + batch_num);
                double max_wh = img_w0 / (double)img_h;
                for (int k = beg; k < end; ++k) max_wh = std::max(max_wh, cw[l0 + order[k]] * 1.0 / ch[l0 + order[k]]);
                int W = (int)(img_h * max_wh);
                if (op.rec_width_bucket > 0) W = (int)ceil_div(W, op.rec_width_bucket) * op.rec_width_bucket;
                for (int k = beg; k < end; ++k) imgw[l0 + order[k]] = W;
            }
        }
        std::vector<int> rorder(total);
        std::iota(rorder.begin(), rorder.end(), 0);
        std::stable_sort(rorder.begin(), rorder.end(), [&](int a, int b) { return imgw[a] < imgw[b]; });
        {
            std::vector<ResizeJob> jobs(total);
            for (int k = 0; k < total; ++k) {
                int i = rorder[k];
                int rw = resized_w_rule(ch[i], cw[i], img_h, imgw[i]);
                jobs[k] = make_job(d_pool + coff[i], ch[i], cw[i], cw[i] * 3, ch[i], cw[i], img_h, rw, 0);
            }
            B200_CUDA_CHECK(cudaMemcpyAsync(d_rj, jobs.data(), sizeof(ResizeJob) * total, cudaMemcpyHostToDevice, st_));
            B200_CUDA_CHECK(cudaMemcpyAsync(d_rslot, rorder.data(), sizeof(int) * total, cudaMemcpyHostToDevice, st_));
        }
        if (use_cls) {                            // predict_cls.py:85-88: '180' in label and score > cls_thresh
            rot_patch_kernel<<<(unsigned)ceil_div(total, 128), 128, 0, st_>>>(d_rj, d_rslot, total, d_clabel, d_cscore,
                                                                             (unsigned)op.cls_rot_label_mask, op.cls_thresh);
            launches += 1;
        }
        int* d_ids = (int*)d_boxes_.get(((size_t)total * max_len + 2 * (size_t)total) * sizeof(int));
        int* d_lens = d_ids + (size_t)total * max_len;
        float* d_scores = (float*)(d_lens + total);
        B200_CUDA_CHECK(cudaMemsetAsync(d_ids, 0, (size_t)total * max_len * sizeof(int), st_));
        {
            std::vector<int> imgw_sorted(total);
            for (int k = 0; k < total; ++k) imgw_sorted[k] = imgw[rorder[k]];
            rec_batches(rec, d_rj, d_rslot, imgw_sorted, img_h, max_len, d_ids, d_lens, d_scores);
        }
        B200_CUDA_CHECK(cudaMemcpyAsync(h_ids + (size_t)line0 * max_len, d_ids, (size_t)total * max_len * sizeof(int), cudaMemcpyDeviceToHost, st_));
        B200_CUDA_CHECK(cudaMemcpyAsync(h_lens + line0, d_lens, (size_t)total * sizeof(int), cudaMemcpyDeviceToHost, st_));
        B200_CUDA_CHECK(cudaMemcpyAsync(h_scores + line0, d_scores, (size_t)total * sizeof(float), cudaMemcpyDeviceToHost, st_));
        if (use_cls) {
            B200_CUDA_CHECK(cudaMemcpyAsync(h_clabel + line0, d_clabel, (size_t)total * sizeof(int), cudaMemcpyDeviceToHost, st_));
            B200_CUDA_CHECK(cudaMemcpyAsync(h_cscore + line0, d_cscore, (size_t)total * sizeof(float), cudaMemcpyDeviceToHost, st_));
        }
        ocr_ms[4] += ms_since(t0);
        line0 += total;
    }
    t0 = now();
    B200_CUDA_CHECK(cudaStreamSynchronize(st_));
    if (line0 > 0) {
        memcpy(ids_out, h_ids, (size_t)line0 * max_len * sizeof(int));
        memcpy(lens_out, h_lens, (size_t)line0 * sizeof(int));
        memcpy(scores_out, h_scores, (size_t)line0 * sizeof(float));
        if (use_cls && cls_labels_out) memcpy(cls_labels_out, h_clabel, (size_t)line0 * sizeof(int));
        if (use_cls && cls_scores_out) memcpy(cls_scores_out, h_cscore, (size_t)line0 * sizeof(float));
    }
    ocr_ms[5] = ms_since(t0);
}

void Stages::db_postprocess_host(const float* prob, int h, int w, int src_h, int src_w, const b200ocr_det_params& p,
                                 int32_t* boxes_out, double* scores_out, int capacity, int* n_boxes) {
    float* d = (float*)d_prob_.get((size_t)h * w * sizeof(float));
    B200_CUDA_CHECK(cudaMemcpyAsync(d, prob, (size_t)h * w * sizeof(float), cudaMemcpyHostToDevice, st_));
    std::vector<int32_t> boxes;
    std::vector<double> scores;
    std::vector<int> counts;
    db_postprocess_device(d, 1, h, w, &src_h, &src_w, p, p.max_candidates, &boxes, &scores, &counts);
    int n = counts.empty() ? 0 : counts[0];
    if (n > capacity) throw std::runtime_error("box capacity too small: need " + std::to_string(n));
    memcpy(boxes_out, boxes.data(), (size_t)n * 8 * sizeof(int32_t));
    if (scores_out) memcpy(scores_out, scores.data(), (size_t)n * sizeof(double));
    *n_boxes = n;
}

void Stages::l2_flush() {
    const size_t bytes = (size_t)256 << 20;      // > 126 MB L2
    void* p = d_flush_.get(bytes);
    B200_CUDA_CHECK(cudaMemsetAsync(p, 0x5a, bytes, st_));
    B200_CUDA_CHECK(cudaStreamSynchronize(st_));
}

void Stages::resident_upload(const uint8_t* const* pages, const int* hs, const int* ws, const int* strides, int n) {
    res_off_.assign(n, 0);
    res_h_.assign(hs, hs + n);
    res_w_.assign(ws, ws + n);
    res_stride_.resize(n);
    size_t total = 0;
    for (int i = 0; i < n; ++i) {
        res_stride_[i] = strides && strides[i] > 0 ? strides[i] : ws[i] * 3;
        res_off_[i] = total;
        total += ((size_t)hs[i] * res_stride_[i] + 255) & ~(size_t)255;
    }
    uint8_t* d = (uint8_t*)d_res_.get(total);
    for (int i = 0; i < n; ++i)
        B200_CUDA_CHECK(cudaMemcpyAsync(d + res_off_[i], pages[i], (size_t)hs[i] * res_stride_[i], cudaMemcpyHostToDevice, st_));
    B200_CUDA_CHECK(cudaStreamSynchronize(st_));
}

void Stages::resident_run(Model* det, const b200ocr_det_params& p, int iters, float* ms_total, int* n_boxes_total,
                          float* net_ms_total) {
    int n = (int)res_h_.size();
    if (n == 0) throw std::runtime_error("bad argument: no resident pages uploaded");
    PagePool pool;
    pool.base = (uint8_t*)d_res_.p;
    pool.off = res_off_;
    pool.stride = res_stride_;
    float total = 0;
    int boxes_total = 0;
    for (int it = 0; it < iters; ++it) {
        std::vector<std::vector<float>> boxes;
        cudaEventRecord(ev_[0], st_);
        det_on_pool(det, pool, nullptr, res_h_.data(), res_w_.data(), n, p, &boxes, nullptr, 0, nullptr);
        cudaEventRecord(ev_[3], st_);                  // the engine stream has joined both lanes
        B200_CUDA_CHECK(cudaEventSynchronize(ev_[3]));
        float ms;
        cudaEventElapsedTime(&ms, ev_[0], ev_[3]);
        total += ms;
        for (auto& b : boxes) boxes_total += (int)(b.size() / 8);
    }
    if (ms_total) *ms_total = total;
    if (n_boxes_total) *n_boxes_total = boxes_total;
    if (net_ms_total) *net_ms_total = 0.f;             // the lanes overlap network and post-processing: see b200ocr_time_plan
}

// The same benchmark through the streaming path: two batches in flight, step k on pages [k*pps, (k+1)*pps) modulo the
// resident set (upload more page sets than the 126 MB L2 holds and no flush is needed between steps).
void Stages::resident_stream(Model* det, const b200ocr_det_params& p, int iters, int pps, float* ms_total, int* n_boxes_total) {
    const int n = (int)res_h_.size();
    if (n == 0) throw std::runtime_error("bad argument: no resident pages uploaded");
    if (pps <= 0 || pps > n || n % pps) throw std::runtime_error("bad argument: pages_per_step must divide the resident page count");
    drain_tickets();
    for (auto& t : tickets_)
        if (t.busy) throw std::runtime_error("bad argument: collect the det batches in flight first");
    int boxes_total = 0;
    auto finish = [&](DetTicket* t) {
        det_finish(t);
        t->busy = false;
        for (auto& b : t->boxes) boxes_total += (int)(b.size() / 8);
    };
    B200_CUDA_CHECK(cudaEventRecord(ev_[0], st_));
    static const int depth = [] {
        const char* e = getenv("B200OCR_DET_DEPTH");
        int v = e ? atoi(e) : kDetSlots;
        return v >= 1 && v <= kDetSlots ? v : kDetSlots;
    }();
    for (int it = 0; it < iters; ++it) {
        DetTicket* t = &tickets_[it % depth];
        if (t->busy) finish(t);
        t->slot = it % depth;
        const int first = (int)(((int64_t)it * pps) % n);
        PagePool pool;
        pool.base = (uint8_t*)d_res_.p;
        pool.off.assign(res_off_.begin() + first, res_off_.begin() + first + pps);
        pool.stride.assign(res_stride_.begin() + first, res_stride_.begin() + first + pps);
        try {
            det_enqueue(t, det, pool, nullptr, res_h_.data() + first, res_w_.data() + first, pps, p, nullptr, 0, nullptr,
                        it < depth ? st_ : nullptr);
        } catch (...) {
            det_abort(t);
            for (auto& o : tickets_)
                if (o.busy) { det_abort(&o); o.busy = false; }
            throw;
        }
        t->busy = true;
    }
    for (int k = 0; k < depth; ++k) {
        DetTicket* t = &tickets_[(iters + k) % depth];
        if (t->busy) finish(t);
    }
    for (int l = 0; l < kDetLanes; ++l) {
        B200_CUDA_CHECK(cudaEventRecord(ev_det_done_[l], det_st_[l]));
        B200_CUDA_CHECK(cudaStreamWaitEvent(st_, ev_det_done_[l], 0));
    }
    B200_CUDA_CHECK(cudaEventRecord(ev_[3], st_));
    B200_CUDA_CHECK(cudaEventSynchronize(ev_[3]));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev_[0], ev_[3]);
    if (ms_total) *ms_total = ms;
    if (n_boxes_total) *n_boxes_total = boxes_total;
}

}  // namespace b200ocr

// extern "C" surface of libb200ocr.so (declared in include/b200ocr.h).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>

#include "../../include/b200ocr.h"
#include "engine.h"
#include "stages.h"
#include "common.cuh"

using namespace b200ocr;

struct b200ocr_engine {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::unique_ptr<Model> models[3];
    std::unique_ptr<Stages> stages;
    std::string err;
    std::mutex mu;
};

static std::string g_create_err;

namespace {
template <typename F>
int guarded(b200ocr_engine* e, F&& fn) {
    if (!e) return B200OCR_E_BAD_ARG;
    try {
        cudaSetDevice(e->device);
        return fn();
    } catch (const std::exception& ex) {
        e->err = ex.what();
        if (e->err.find("CUDA error") != std::string::npos) return B200OCR_E_CUDA;
        if (e->err.find("unsupported") != std::string::npos) return B200OCR_E_UNSUPPORTED;
        if (e->err.find("too small") != std::string::npos || e->err.find("capacity") != std::string::npos)
            return B200OCR_E_CAPACITY;
        if (e->err.find("cannot open") != std::string::npos) return B200OCR_E_IO;
        return B200OCR_E_BAD_ARG;
    } catch (...) {
        e->err = "unknown C++ exception";
        return B200OCR_E_BAD_ARG;
    }
}
Model* need(b200ocr_engine* e, int kind) {
    if (kind < 0 || kind > 2) throw std::runtime_error("bad model kind");
    if (!e->models[kind]) throw std::runtime_error("model not loaded");
    return e->models[kind].get();
}
}  // namespace

#pragma GCC visibility push(default)
extern "C" {

const char* b200ocr_version(void) { return "b200ocr 0.1 (sm_100a)"; }

int b200ocr_create(int device, b200ocr_engine** out) {
    if (!out) return B200OCR_E_BAD_ARG;
    *out = nullptr;
    int n = 0;
    cudaError_t err = cudaGetDeviceCount(&n);
    if (err != cudaSuccess || n == 0) {
        g_create_err = std::string("no CUDA device available (") + cudaGetErrorString(err) +
                       "); libb200ocr has no CPU fallback";
        return B200OCR_E_CUDA;
    }
    if (device < 0 || device >= n) {
        g_create_err = "device index out of range";
        return B200OCR_E_BAD_ARG;
    }
    err = cudaSetDevice(device);
    if (err != cudaSuccess) {
        g_create_err = cudaGetErrorString(err);
        return B200OCR_E_CUDA;
    }
    std::unique_ptr<b200ocr_engine> e(new b200ocr_engine());
    e->device = device;
    err = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
    if (err != cudaSuccess) {
        g_create_err = cudaGetErrorString(err);
        return B200OCR_E_CUDA;
    }
    try {
        e->stages.reset(new Stages(e->stream));
    } catch (const std::exception& ex) {
        g_create_err = ex.what();
        return B200OCR_E_CUDA;
    }
    *out = e.release();
    return B200OCR_OK;
}

void b200ocr_destroy(b200ocr_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    e->stages.reset();
    for (auto& m : e->models) m.reset();
    cudaStreamDestroy(e->stream);
    delete e;
}

const char* b200ocr_last_error(const b200ocr_engine* e) { return e ? e->err.c_str() : g_create_err.c_str(); }

int b200ocr_load_onnx(b200ocr_engine* e, int kind, const char* path, int precision) {
    return guarded(e, [&]() {
        if (kind < 0 || kind > 2 || !path) throw std::runtime_error("bad argument to load_onnx");
        std::lock_guard<std::mutex> lk(e->mu);
        e->models[kind].reset(new Model(path, precision));
        if (kind == B200OCR_DET) {
            // det plans read the u8 pages directly (pre-processing fused into the stem) unless B200OCR_FUSED_PRE=0
            float nv[8];
            Stages::det_norm_values(nv);
            e->models[kind]->set_pre_norm(nv);
            const char* f = getenv("B200OCR_FUSED_PRE");
            e->models[kind]->fused_pre_default = !(f && f[0] == '0');
        }
        return B200OCR_OK;
    });
}

int b200ocr_io_info(b200ocr_engine* e, int kind, char* input_name, size_t input_cap, char* output_name,
                    size_t output_cap, int64_t in_dims[4], int64_t out_dims[4], int* out_rank) {
    return guarded(e, [&]() {
        Model* m = need(e, kind);
        const CGraph& g = m->graph();
        if (input_name && input_cap) snprintf(input_name, input_cap, "%s", g.input_name.c_str());
        if (output_name && output_cap) snprintf(output_name, output_cap, "%s", g.output_name.c_str());
        for (int i = 0; i < 4; ++i) {
            if (in_dims) in_dims[i] = i < (int)g.input_shape.size() ? g.input_shape[i] : 0;
            if (out_dims) out_dims[i] = i < (int)g.output_shape.size() ? g.output_shape[i] : 0;
        }
        if (out_rank) *out_rank = (int)g.output_shape.size();
        return B200OCR_OK;
    });
}

int b200ocr_session_run(b200ocr_engine* e, int kind, const float* nchw, const int64_t dims[4], float* out,
                        size_t out_capacity, int64_t out_dims[4], int* out_rank) {
    return guarded(e, [&]() {
        Model* m = need(e, kind);
        if (!dims || dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0 || dims[3] <= 0) throw std::runtime_error("bad input dims");
        std::lock_guard<std::mutex> lk(e->mu);
        std::vector<int64_t> shape;
        if (!out) {
            Plan* p = m->plan_for((int)dims[0], (int)dims[2], (int)dims[3]);
            shape = p->out_shape;
        } else {
            if (!nchw) throw std::runtime_error("null input");
            m->run_host(nchw, (int)dims[0], (int)dims[1], (int)dims[2], (int)dims[3], &shape, out, out_capacity, e->stream);
            e->stages->launches += m->plan_for((int)dims[0], (int)dims[2], (int)dims[3])->launches + 1;
        }
        if (out_rank) *out_rank = (int)shape.size();
        if (out_dims)
            for (int i = 0; i < 4; ++i) out_dims[i] = i < (int)shape.size() ? shape[i] : 0;
        return B200OCR_OK;
    });
}

void b200ocr_det_params_default(b200ocr_det_params* p) {
    if (!p) return;
    p->limit_side_len = 960.f;
    p->limit_type = 0;
    p->thresh = 0.3f;
    p->box_thresh = 0.6;
    p->unclip_ratio = 1.5;
    p->max_candidates = 1000;
    p->min_size = 3;
    p->use_dilation = 0;
    p->score_mode = 0;
}

int b200ocr_det_preprocess(b200ocr_engine* e, const uint8_t* bgr, int h, int w, int row_stride,
                           const b200ocr_det_params* p, float* chw, size_t chw_capacity, int* out_h, int* out_w,
                           double shape[4]) {
    return guarded(e, [&]() {
        if (!bgr || !p || h <= 0 || w <= 0) throw std::runtime_error("bad argument to det_preprocess");
        std::lock_guard<std::mutex> lk(e->mu);
        e->stages->det_preprocess_host(bgr, h, w, row_stride, *p, chw, chw_capacity, out_h, out_w, shape);
        return B200OCR_OK;
    });
}

int b200ocr_db_postprocess(b200ocr_engine* e, const float* prob, int h, int w, int src_h, int src_w,
                           const b200ocr_det_params* p, int32_t* boxes_out, double* scores_out, int capacity,
                           int* n_boxes) {
    return guarded(e, [&]() {
        if (!prob || !p || !boxes_out || !n_boxes || h <= 0 || w <= 0) throw std::runtime_error("bad argument to db_postprocess");
        std::lock_guard<std::mutex> lk(e->mu);
        e->stages->db_postprocess_host(prob, h, w, src_h, src_w, *p, boxes_out, scores_out, capacity, n_boxes);
        return B200OCR_OK;
    });
}

int b200ocr_det_pages(b200ocr_engine* e, const uint8_t* const* pages, const int* hs, const int* ws,
                      const int* strides, int n, const b200ocr_det_params* p, float* boxes_out, int max_boxes,
                      int* n_boxes_out, float* probmap_out, size_t probmap_capacity, int* probmap_dims_out) {
    return guarded(e, [&]() {
        if (!pages || !hs || !ws || !p || n < 0 || !boxes_out || !n_boxes_out) throw std::runtime_error("bad argument to det_pages");
        Model* m = need(e, B200OCR_DET);
        std::lock_guard<std::mutex> lk(e->mu);
        e->stages->det_pages(m, pages, hs, ws, strides, n, *p, boxes_out, max_boxes, n_boxes_out, probmap_out,
                             probmap_capacity, probmap_dims_out);
        return B200OCR_OK;
    });
}

int b200ocr_det_submit(b200ocr_engine* e, const uint8_t* const* pages, const int* hs, const int* ws,
                       const int* strides, int n, const b200ocr_det_params* p, int* ticket_out) {
    return guarded(e, [&]() {
        if (!pages || !hs || !ws || !p || n <= 0 || !ticket_out) throw std::runtime_error("bad argument to det_submit");
        Model* m = need(e, B200OCR_DET);
        std::lock_guard<std::mutex> lk(e->mu);
        *ticket_out = e->stages->det_submit(m, pages, hs, ws, strides, n, *p);
        return B200OCR_OK;
    });
}

int b200ocr_det_collect(b200ocr_engine* e, int ticket, float* boxes_out, int max_boxes, int* n_boxes_out) {
    return guarded(e, [&]() {
        if (!boxes_out || !n_boxes_out || max_boxes <= 0) throw std::runtime_error("bad argument to det_collect");
        std::lock_guard<std::mutex> lk(e->mu);
        e->stages->det_collect(ticket, boxes_out, max_boxes, n_boxes_out);
        return B200OCR_OK;
    });
}

int b200ocr_crop_boxes(b200ocr_engine* e, const uint8_t* bgr, int h, int w, int row_stride, const float* boxes,
                       int n, uint8_t* crops_out, size_t crops_capacity, int* crop_dims_out) {
    return guarded(e, [&]() {
        if (!bgr || (!boxes && n > 0) || n < 0 || !crop_dims_out) throw std::runtime_error("bad argument to crop_boxes");
        std::lock_guard<std::mutex> lk(e->mu);
        e->stages->crop_boxes_host(bgr, h, w, row_stride, boxes, n, crops_out, crops_capacity, crop_dims_out);
        return B200OCR_OK;
    });
}

int b200ocr_cls_crops(b200ocr_engine* e, const uint8_t* const* crops, const int* hs, const int* ws, int n,
                      int batch_num, int32_t* labels_out, float* scores_out) {
    return guarded(e, [&]() {
        if ((n > 0 && (!crops || !hs || !ws)) || n < 0 || !labels_out || !scores_out) throw std::runtime_error("bad argument to cls_crops");
        Model* m = need(e, B200OCR_CLS);
        std::lock_guard<std::mutex> lk(e->mu);
        e->stages->cls_crops(m, crops, hs, ws, n, batch_num, labels_out, scores_out);
        return B200OCR_OK;
    });
}

int b200ocr_rec_crops(b200ocr_engine* e, const uint8_t* const* crops, const int* hs, const int* ws, int n,
                      int batch_num, const uint8_t* rotate180, int32_t* ids_out, int max_len, int* lens_out,
                      float* scores_out) {
    return guarded(e, [&]() {
        if ((n > 0 && (!crops || !hs || !ws)) || n < 0 || !ids_out || !lens_out || !scores_out || max_len <= 0)
            throw std::runtime_error("bad argument to rec_crops");
        Model* m = need(e, B200OCR_REC);
        std::lock_guard<std::mutex> lk(e->mu);
        e->stages->rec_crops(m, crops, hs, ws, n, batch_num, rotate180, ids_out, max_len, lens_out, scores_out);
        return B200OCR_OK;
    });
}

int b200ocr_rec_argmax(b200ocr_engine* e, const float* nchw, const int64_t dims[4], int32_t* idx_out, float* prob_out,
                       int capacity, int* steps_out) {
    return guarded(e, [&]() {
        if (!nchw || !dims) throw std::runtime_error("bad argument to rec_argmax");
        Model* m = need(e, B200OCR_REC);
        std::lock_guard<std::mutex> lk(e->mu);
        e->stages->rec_argmax_host(m, nchw, (int)dims[0], (int)dims[1], (int)dims[2], (int)dims[3], idx_out, prob_out, capacity,
                                   steps_out);
        return B200OCR_OK;
    });
}

void b200ocr_ocr_params_default(b200ocr_ocr_params* p) {
    if (!p) return;
    b200ocr_det_params_default(&p->det);
    p->use_angle_cls = 0;
    p->cls_thresh = 0.9f;
    p->cls_rot_label_mask = 0x2u;
    p->rec_batch_num = 6;
    p->rec_width_bucket = 0;
}

int b200ocr_ocr_pages(b200ocr_engine* e, const uint8_t* const* pages, const int* hs, const int* ws, const int* strides,
                      int n, const b200ocr_ocr_params* p, float* boxes_out, int line_cap, int* n_boxes_out,
                      int32_t* ids_out, int max_len, int* lens_out, float* scores_out, int32_t* cls_labels_out,
                      float* cls_scores_out) {
    return guarded(e, [&]() {
        if (!pages || !hs || !ws || !p || n < 0 || !boxes_out || !n_boxes_out || !ids_out || !lens_out || !scores_out ||
            line_cap <= 0 || max_len <= 0)
            throw std::runtime_error("bad argument to ocr_pages");
        Model* det = need(e, B200OCR_DET);
        Model* rec = need(e, B200OCR_REC);
        Model* cls = p->use_angle_cls ? need(e, B200OCR_CLS) : nullptr;
        std::lock_guard<std::mutex> lk(e->mu);
        e->stages->ocr_pages(det, cls, rec, pages, hs, ws, strides, n, *p, boxes_out, line_cap, n_boxes_out, ids_out, max_len,
                             lens_out, scores_out, cls_labels_out, cls_scores_out);
        return B200OCR_OK;
    });
}

int b200ocr_describe(b200ocr_engine* e, int kind, char* buf, size_t cap) {
    return guarded(e, [&]() {
        Model* m = need(e, kind);
        std::string s = m->describe();
        if (buf && cap) snprintf(buf, cap, "%s", s.c_str());
        return (int)std::min<size_t>(s.size(), 0x7fffffff);
    });
}

int b200ocr_plan_stats(b200ocr_engine* e, int kind, int n, int h, int w, double* flops, int* launches,
                       double* arena_bytes) {
    return guarded(e, [&]() {
        Model* m = need(e, kind);
        std::lock_guard<std::mutex> lk(e->mu);
        Plan* p = m->plan_for(n, h, w, false, 0, 0, m->fused_pre_default);
        if (flops) *flops = p->flops;
        if (launches) *launches = p->launches;
        if (arena_bytes) *arena_bytes = (double)p->arena_bytes;
        return B200OCR_OK;
    });
}

int b200ocr_profile_plan(b200ocr_engine* e, int kind, int n, int h, int w, int iters, float* ms_out, double* flops_out,
                         double* bytes_out, int capacity, char* names_buf, size_t names_cap) {
    return guarded(e, [&]() {
        Model* m = need(e, kind);
        std::lock_guard<std::mutex> lk(e->mu);
        std::vector<float> ms;
        m->profile(n, h, w, iters < 1 ? 1 : iters, e->stream, &ms);
        Plan* p = m->plan_for(n, h, w, false, 0, 0, m->fused_pre_default);
        int ns = (int)ms.size();
        if (ns > capacity) throw std::runtime_error("profile buffers too small: need " + std::to_string(ns));
        size_t off = 0;
        for (int i = 0; i < ns; ++i) {
            ms_out[i] = ms[i];
            if (flops_out) flops_out[i] = p->step_flops[i];
            if (bytes_out) bytes_out[i] = p->step_bytes[i];
            if (names_buf) {
                const std::string& nm = p->step_names[i];
                if (off + nm.size() + 1 < names_cap) {
                    memcpy(names_buf + off, nm.c_str(), nm.size() + 1);
                    off += nm.size() + 1;
                }
            }
        }
        e->stages->launches += (int64_t)ns * (iters < 1 ? 1 : iters);
        return ns;
    });
}

int b200ocr_time_plan(b200ocr_engine* e, int kind, int n, int h, int w, int head, int iters, float* ms_avg) {
    return guarded(e, [&]() {
        Model* m = need(e, kind);
        if (!ms_avg || iters < 1) throw std::runtime_error("bad argument to time_plan");
        std::lock_guard<std::mutex> lk(e->mu);
        Plan* p = m->plan_for(n, h, w, head != 0, 0, 0, m->fused_pre_default);
        for (int i = 0; i < 3; ++i) p->launch(e->stream);          // warm: lazy init, graph capture
        cudaEvent_t a, b;
        B200_CUDA_CHECK(cudaEventCreate(&a));
        B200_CUDA_CHECK(cudaEventCreate(&b));
        B200_CUDA_CHECK(cudaEventRecord(a, e->stream));
        for (int i = 0; i < iters; ++i) p->launch(e->stream);
        B200_CUDA_CHECK(cudaEventRecord(b, e->stream));
        B200_CUDA_CHECK(cudaEventSynchronize(b));
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        cudaEventDestroy(a);
        cudaEventDestroy(b);
        *ms_avg = ms / iters;
        e->stages->launches += (int64_t)p->launches * (iters + 3);
        return B200OCR_OK;
    });
}

int b200ocr_l2_flush(b200ocr_engine* e) {
    return guarded(e, [&]() {
        std::lock_guard<std::mutex> lk(e->mu);
        e->stages->l2_flush();
        return B200OCR_OK;
    });
}

void* b200ocr_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    return p;
}

void b200ocr_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int64_t b200ocr_launch_count(const b200ocr_engine* e) { return e && e->stages ? e->stages->launches : 0; }

int b200ocr_stage_times(b200ocr_engine* e, float ms_out[4]) {
    return guarded(e, [&]() {
        for (int i = 0; i < 4; ++i) ms_out[i] = e->stages->stage_ms[i];
        return B200OCR_OK;
    });
}

int b200ocr_ocr_stage_times(b200ocr_engine* e, float ms_out[6]) {
    return guarded(e, [&]() {
        for (int i = 0; i < 6; ++i) ms_out[i] = e->stages->ocr_ms[i];
        return B200OCR_OK;
    });
}

void b200ocr_trace_dump(void) { trace_dump(); }

int b200ocr_det_resident_upload(b200ocr_engine* e, const uint8_t* const* pages, const int* hs, const int* ws,
                                const int* strides, int n) {
    return guarded(e, [&]() {
        if (!pages || !hs || !ws || n <= 0) throw std::runtime_error("bad argument to det_resident_upload");
        std::lock_guard<std::mutex> lk(e->mu);
        e->stages->resident_upload(pages, hs, ws, strides, n);
        return B200OCR_OK;
    });
}

int b200ocr_det_resident_run(b200ocr_engine* e, const b200ocr_det_params* p, int iters, float* ms_total,
                             int* n_boxes_total, float* net_ms_total) {
    return guarded(e, [&]() {
        Model* m = need(e, B200OCR_DET);
        std::lock_guard<std::mutex> lk(e->mu);
        e->stages->resident_run(m, *p, iters, ms_total, n_boxes_total, net_ms_total);
        return B200OCR_OK;
    });
}

int b200ocr_det_resident_stream(b200ocr_engine* e, const b200ocr_det_params* p, int iters, int pages_per_step,
                                float* ms_total, int* n_boxes_total) {
    return guarded(e, [&]() {
        if (!p || iters <= 0) throw std::runtime_error("bad argument to det_resident_stream");
        Model* m = need(e, B200OCR_DET);
        std::lock_guard<std::mutex> lk(e->mu);
        e->stages->resident_stream(m, *p, iters, pages_per_step, ms_total, n_boxes_total);
        return B200OCR_OK;
    });
}

}  // extern "C"
#pragma GCC visibility pop

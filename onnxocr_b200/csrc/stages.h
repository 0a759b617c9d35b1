// Fused pre/post-processing stages around the graph executor (Level B of the boundary).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <utility>
#include <vector>

#include "../../include/b200ocr.h"
#include "engine.h"

namespace b200ocr {

struct DetGeom {            // DetResizeForTest.resize_image_type0 result (operators.py:91-138)
    int src_h, src_w;       // original size (shape[0], shape[1])
    int pad_h, pad_w;       // after image_padding (operators.py:72-76)
    int dst_h, dst_w;       // network input size
    double ratio_h, ratio_w;
};

DetGeom det_geometry(int h, int w, const b200ocr_det_params& p);

// growable device / pinned buffers
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    void* get(size_t bytes);
    ~DevBuf();
};
struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    void* get(size_t bytes);
    ~PinBuf();
};

// Coarse device timeline for diagnostics (B200OCR_TRACE=<file>): a scope records a CUDA event on its stream when it is
// opened and when it is closed, i.e. around the work enqueued in between; trace_dump() (engine destruction) writes one JSON
// line per scope with start / end in ms relative to the first event.  Off (no events, no cost) unless the variable is set.
struct TraceScope {
    TraceScope(const char* name, cudaStream_t st, int tag = 0);
    ~TraceScope();
    int idx = -1;
    cudaStream_t st = nullptr;
};
void trace_dump();

struct DbWork;
struct ResizeJob;              // device scratch of DBPostProcess (dbpost.cu)
void db_free(DbWork* w);

class Stages {
public:
    explicit Stages(cudaStream_t st);
    ~Stages();

    void det_preprocess_host(const uint8_t* bgr, int h, int w, int row_stride, const b200ocr_det_params& p, float* chw,
                             size_t chw_capacity, int* out_h, int* out_w, double shape[4]);
    void db_postprocess_host(const float* prob, int h, int w, int src_h, int src_w, const b200ocr_det_params& p,
                             int32_t* boxes_out, double* scores_out, int capacity, int* n_boxes);
    void det_pages(Model* det, const uint8_t* const* pages, const int* hs, const int* ws, const int* strides, int n,
                   const b200ocr_det_params& p, float* boxes_out, int max_boxes, int* n_boxes_out, float* probmap_out,
                   size_t probmap_capacity, int* probmap_dims_out);
    void crop_boxes_host(const uint8_t* bgr, int h, int w, int row_stride, const float* boxes, int n, uint8_t* crops_out,
                         size_t crops_capacity, int* crop_dims_out);
    void cls_crops(Model* cls, const uint8_t* const* crops, const int* hs, const int* ws, int n, int batch_num,
                   int32_t* labels_out, float* scores_out);
    void rec_crops(Model* rec, const uint8_t* const* crops, const int* hs, const int* ws, int n, int batch_num,
                   const uint8_t* rotate180, int32_t* ids_out, int max_len, int* lens_out, float* scores_out);
    // Level C: TextSystem.__call__ for a batch of pages, device resident between stages (flat line-major outputs)
    void ocr_pages(Model* det, Model* cls, Model* rec, const uint8_t* const* pages, const int* hs, const int* ws,
                   const int* strides, int n, const b200ocr_ocr_params& op, float* boxes_out, int line_cap, int* n_boxes_out,
                   int32_t* ids_out, int max_len, int* lens_out, float* scores_out, int32_t* cls_labels_out,
                   float* cls_scores_out);
    void rec_argmax_host(Model* rec, const float* nchw, int n, int c, int h, int w, int32_t* idx_out, float* prob_out,
                         int capacity, int* T_out);
    void resident_upload(const uint8_t* const* pages, const int* hs, const int* ws, const int* strides, int n);
    void resident_run(Model* det, const b200ocr_det_params& p, int iters, float* ms_total, int* n_boxes_total,
                      float* net_ms_total);
    // Streaming detection: det_submit enqueues a batch (uploads, network, DBPostProcess) on a free slot (two det lanes
    // of its own) and returns without waiting for the GPU; det_collect blocks until that batch's boxes are on the host.
    // With two batches in flight the upload and network of batch k+1 overlap the DBPostProcess tail and the box
    // read-back of batch k.  Host page buffers must stay valid until the batch is collected.
    int det_submit(Model* det, const uint8_t* const* pages, const int* hs, const int* ws, const int* strides, int n,
                   const b200ocr_det_params& p);
    void det_collect(int ticket, float* boxes_out, int max_boxes, int* n_boxes_out);
    // the resident benchmark through the streaming path: step k runs pages [k*pages_per_step ...) (mod the resident set)
    void resident_stream(Model* det, const b200ocr_det_params& p, int iters, int pages_per_step, float* ms_total,
                         int* n_boxes_total);

    void l2_flush();
    static void det_norm_values(float v[8]);     // NormalizeImage constants of the det path (NormP layout)

    int64_t launches = 0;
    float stage_ms[4] = {0, 0, 0, 0};
    // host wall clock of the last ocr_pages call: upload enqueue, det (incl. box read-back), host box/crop planning,
    // warp + cls enqueue, rec enqueue, final drain
    float ocr_ms[6] = {0, 0, 0, 0, 0, 0};

private:
    struct PagePool {
        uint8_t* base = nullptr;
        std::vector<size_t> off;
        std::vector<int> stride;
    };
    // device page pool + pinned staging area of one call (the synchronous entry points share one, every streaming slot
    // has its own)
    struct PageStore {
        DevBuf d;
        PinBuf h;
        size_t cap = 0, used = 0;
    };
    void upload_page(uint8_t* dst, const uint8_t* src, size_t bytes, cudaStream_t st, PageStore* store);
    void plan_pool(const int* hs, const int* ws, const int* strides, int n, PagePool* pool, PageStore* store);
    // Det lanes (stream + DBPostProcess scratch + job table each): while the host waits for the boxes of one
    // sub-batch, the network of the next one is already running on the other lane.
    struct DbPending {
        int lane = 0, n = 0, h = 0, w = 0, maxc = 0, nt = 0;
        const float* d_prob = nullptr;
        b200ocr_det_params p;
        bool active = false;
    };
    static constexpr int kLanesPerSlot = 2;
    static constexpr int kDetSlots = 4;                      // batches in flight through det_submit
    static constexpr int kDetLanes = kLanesPerSlot * kDetSlots;
    struct DetChunk { int g0, b, h, w; };
    struct DetTicket {                                       // the host-side state of one enqueued batch
        bool busy = false;                                   // submitted, not yet collected
        int slot = 0, id = 0, n = 0, turn = 0;
        std::vector<int> idx;
        std::vector<std::vector<float>> boxes;
        DbPending pend[kLanesPerSlot];
        DetChunk pend_chunk[kLanesPerSlot];
        PageStore store;
    };
    DetTicket tickets_[kDetSlots];
    DetTicket sync_ticket_;                                  // det_on_pool (lanes of slot 0)
    int ticket_seq_ = 0;
    // pages == nullptr: the pool already holds the pixels (resident benchmark path)
    void det_enqueue(DetTicket* t, Model* det, const PagePool& pool, const uint8_t* const* pages, const int* hs, const int* ws,
                     int n, const b200ocr_det_params& p, float* probmap_out, size_t probmap_capacity, int* probmap_dims_out,
                     cudaStream_t resident_src);
    void det_collect_lane(DetTicket* t, int k);
    void det_finish(DetTicket* t);
    void det_abort(DetTicket* t);                            // enqueue failed: leave the slot's lanes clean
    void drain_tickets();                                    // finish every streaming batch (results stay in its ticket)
    void det_on_pool(Model* det, const PagePool& pool, const uint8_t* const* pages, const int* hs, const int* ws, int n,
                     const b200ocr_det_params& p, std::vector<std::vector<float>>* boxes_pp, float* probmap_out,
                     size_t probmap_capacity, int* probmap_dims_out);
    void db_post_enqueue(int lane, const float* d_prob, int n, int h, int w, const int* src_hs, const int* src_ws,
                         const b200ocr_det_params& p, int capacity, DbPending* pend);
    void db_post_collect(DbPending* pend, std::vector<int32_t>* boxes, std::vector<double>* scores, std::vector<int>* counts);
    cudaStream_t det_st_[kDetLanes];
    DbWork* db_lane_[kDetLanes] = {};
    DevBuf d_lane_jobs_[kDetLanes];
    cudaEvent_t ev_det_done_[kDetLanes];
    cudaEvent_t ev_net_done_[kDetLanes];        // recorded right after a lane's network (rolling window in det_enqueue)
    int net_lane_ring_[kDetLanes] = {};
    long long net_seq_ = 0;
    void rec_batch(Model* rec, const struct ResizeJob* d_jobs, const int* d_slots, int b, int img_h, int W, int max_len,
                   int* d_ids, int* d_lens, float* d_scores, int lane);
    void rec_batches(Model* rec, const struct ResizeJob* d_jobs, const int* d_slots, const std::vector<int>& imgw_sorted,
                     int img_h, int max_len, int* d_ids, int* d_lens, float* d_scores);
    static constexpr int kAux = 16;             // auxiliary streams for independent small rec batches (n_aux_ in use)
    int n_aux_ = 8;
    std::map<std::pair<int, int>, int> rec_lane_of_;   // (batch, width) -> auxiliary stream within the current call
    cudaStream_t aux_[kAux];
    cudaEvent_t ev_fork_, ev_join_[kAux];
    DevBuf d_lane_[kAux + 1];                   // per-stream CTC scratch
    // device-side DBPostProcess on a device-resident fp32 probability map [n][h][w]; results on the host
    void db_postprocess_device(const float* d_prob, int n, int h, int w, const int* src_hs, const int* src_ws,
                               const b200ocr_det_params& p, int capacity, std::vector<int32_t>* boxes,
                               std::vector<double>* scores, std::vector<int>* counts);
    cudaStream_t st_;
    cudaEvent_t ev_[4];
    DevBuf d_img_, d_chw_, d_prob_, d_boxes_, d_crops_, d_misc_, d_batch_, d_jobs_;
    PinBuf h_pin_, h_pin2_;
    DbWork* db_ = nullptr;
    // resident pages
    DevBuf d_res_, d_flush_;
    std::vector<size_t> res_off_;
    std::vector<int> res_h_, res_w_, res_stride_;
};

}  // namespace b200ocr

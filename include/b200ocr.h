/* libb200ocr.so — C ABI of the B200-native PP-OCRv5 engine.
 *
 * Drop-in boundary for the engine seam of ding113/OnnxOCR (all citations are
 * relative to the reference checkout):
 *
 *   - onnxocr/predict_base.py:7-17   PredictBase.get_onnx_session -> onnxruntime.InferenceSession
 *   - onnxocr/predict_det.py:94-120  TextDetector.__call__
 *   - onnxocr/predict_cls.py:44-89   TextClassifier.__call__
 *   - onnxocr/predict_rec.py:280-326 TextRecognizer.__call__
 *   - onnxocr/predict_system.py:35-72 TextSystem.__call__
 *
 * Conventions: every entry point is `extern "C"`, takes plain pointers and sizes
 * (numpy buffers via ctypes), returns 0 on success or a negative B200OCR_E_* code;
 * `b200ocr_last_error` returns the message of the last failure on that engine
 * (or of the last failed `b200ocr_create` when `engine` is NULL).  All buffers are
 * caller-allocated; the library never retains a host pointer past the call.  There
 * is no CPU fallback: without a CUDA device `b200ocr_create` fails.
 */
#ifndef B200OCR_H_
#define B200OCR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200ocr_engine b200ocr_engine;

enum {
    B200OCR_OK = 0,
    B200OCR_E_BAD_ARG = -1,
    B200OCR_E_UNSUPPORTED = -2,   /* ONNX op / pattern the engine has no kernel for (named in last_error) */
    B200OCR_E_CUDA = -3,
    B200OCR_E_CAPACITY = -4,      /* caller buffer too small */
    B200OCR_E_STATE = -5,         /* model of that kind not loaded */
    B200OCR_E_IO = -6
};

enum { B200OCR_DET = 0, B200OCR_CLS = 1, B200OCR_REC = 2 };
enum { B200OCR_FP32 = 0, B200OCR_BF16 = 1 };

/* ---- lifetime -------------------------------------------------------------------------- */
/* Creates one engine bound to CUDA device `device` (one engine per GPU / per process rank).
 * CUDA is initialised here, never at library load (fork safety: Dockerfile:35 `--preload`). */
int b200ocr_create(int device, b200ocr_engine** out);
void b200ocr_destroy(b200ocr_engine* e);
const char* b200ocr_last_error(const b200ocr_engine* e);
const char* b200ocr_version(void);

/* ---- the onnxruntime seam (Level A) --------------------------------------------------- */
/* Replaces onnxruntime.InferenceSession(model_path, None, providers=...)  predict_base.py:14.
 * Parses the .onnx with the built-in wire reader, folds BN, fuses activations, uploads weights. */
int b200ocr_load_onnx(b200ocr_engine* e, int kind, const char* path, int precision);

/* Replaces session.get_inputs()/get_outputs()  predict_base.py:26-40.  Names are copied into
 * caller buffers (NUL terminated); dims use -1 for dynamic axes. */
int b200ocr_io_info(b200ocr_engine* e, int kind, char* input_name, size_t input_cap, char* output_name,
                    size_t output_cap, int64_t in_dims[4], int64_t out_dims[4], int* out_rank);

/* Replaces session.run(output_names, {name: ndarray})  predict_det.py:107, predict_cls.py:75,
 * predict_rec.py:316.  `nchw` is C-contiguous float32 [dims[0],dims[1],dims[2],dims[3]] on the host;
 * `out` receives the logical ONNX output (float32, C order), `out_dims`/`out_rank` its shape.
 * Call with out == NULL to query the output shape only (nothing is executed). */
int b200ocr_session_run(b200ocr_engine* e, int kind, const float* nchw, const int64_t dims[4], float* out,
                        size_t out_capacity, int64_t out_dims[4], int* out_rank);

/* ---- fused stages (Level B) ------------------------------------------------------------ */
typedef struct b200ocr_det_params {
    float limit_side_len;     /* args.det_limit_side_len (960)        utils.py:271 */
    int limit_type;           /* 0 'max', 1 'min', 2 'resize_long'    utils.py:272 */
    float thresh;             /* det_db_thresh 0.3 (compared in float32, like `pred > self.thresh`)  utils.py:276 */
    int max_candidates;       /* 1000                                 predict_det.py:33 */
    int min_size;             /* 3                                    db_postprocess.py:47 */
    int use_dilation;         /* 0                                    utils.py:280 */
    double box_thresh;        /* det_db_box_thresh 0.6 (Python float, compared in float64)  utils.py:277 */
    double unclip_ratio;      /* det_db_unclip_ratio 1.5 (Python float)                      utils.py:278 */
    int score_mode;           /* det_db_score_mode: 0 'fast' (box_score_fast), 1 'slow' (box_score_slow:
                                 mean over cv2.fillPoly of the contour itself)  utils.py:283, db_postprocess.py:129-132,199-218 */
} b200ocr_det_params;

void b200ocr_det_params_default(b200ocr_det_params* p);

/* DetResizeForTest + NormalizeImage + ToCHWImage (operators.py:54-138, 22-31, 170-176) as one
 * kernel.  `bgr` is HWC uint8 with `row_stride` bytes per row.  Writes the normalised NCHW float32
 * tensor [1,3,H,W] (what the reference feeds to session.run) and shape = [src_h, src_w, ratio_h,
 * ratio_w] as float64.  `*out_h/out_w` receive H, W.  chw may be NULL to query H, W only. */
int b200ocr_det_preprocess(b200ocr_engine* e, const uint8_t* bgr, int h, int w, int row_stride,
                           const b200ocr_det_params* p, float* chw, size_t chw_capacity, int* out_h, int* out_w,
                           double shape[4]);

/* DBPostProcess.__call__ + boxes_from_bitmap (db_postprocess.py:220-246, 104-149) on a probability
 * map given by the caller: `prob` float32 [h,w].  boxes_out: int32 [capacity][4][2] in the order
 * cv2.findContours(RETR_LIST) yields them; scores_out: float64 [capacity].  */
int b200ocr_db_postprocess(b200ocr_engine* e, const float* prob, int h, int w, int src_h, int src_w,
                           const b200ocr_det_params* p, int32_t* boxes_out, double* scores_out, int capacity,
                           int* n_boxes);

/* TextDetector.__call__ (predict_det.py:94-120) for a batch of pages: preprocess -> DBNet ->
 * DBPostProcess -> filter_tag_det_res.  Page i is HWC uint8 BGR at pages[i] (hs[i] x ws[i], row
 * stride strides[i]).  boxes_out: float32 [n][max_boxes][4][2] (ordered clockwise, clipped),
 * n_boxes_out: int32 [n].  probmap_out (optional, may be NULL): float32, pages concatenated,
 * each [H_i, W_i]; probmap_dims_out: int32 [n][2]. */
int b200ocr_det_pages(b200ocr_engine* e, const uint8_t* const* pages, const int* hs, const int* ws,
                      const int* strides, int n, const b200ocr_det_params* p, float* boxes_out, int max_boxes,
                      int* n_boxes_out, float* probmap_out, size_t probmap_capacity, int* probmap_dims_out);

/* Streaming form of b200ocr_det_pages for callers that keep the GPU fed (the reference runs model.ocr from a
 * ThreadPoolExecutor(max_workers=4), ocr_images_pdfs.py:106,154, and under Semaphore(MODEL_CONCURRENCY), app/engine.py:111-120;
 * here the overlap is explicit and single-threaded).  b200ocr_det_submit enqueues the uploads, the network and
 * DBPostProcess of one batch on a free slot (two det lanes of its own) and returns a ticket without waiting for the
 * GPU; b200ocr_det_collect blocks until that batch's boxes are on the host and writes them exactly as
 * b200ocr_det_pages would.  Up to four batches may be in flight (a fifth submit returns B200OCR_E_CAPACITY): the upload
 * and network of batch k+1 then overlap the DBPostProcess tail and box read-back of batch k.  The page buffers of a
 * batch must stay valid and unchanged until it has been collected.  Results are identical to b200ocr_det_pages. */
int b200ocr_det_submit(b200ocr_engine* e, const uint8_t* const* pages, const int* hs, const int* ws,
                       const int* strides, int n, const b200ocr_det_params* p, int* ticket_out);
int b200ocr_det_collect(b200ocr_engine* e, int ticket, float* boxes_out, int max_boxes, int* n_boxes_out);

/* get_rotate_crop_image (utils.py:12-53) for a batch of quads on one page: bicubic perspective warp
 * with BORDER_REPLICATE, rot90 when h/w >= 1.5.  boxes: float32 [n][4][2].  Crops are written
 * back-to-back into crops_out (HWC uint8); crop_dims_out: int32 [n][2] = (h, w) after rotation. */
int b200ocr_crop_boxes(b200ocr_engine* e, const uint8_t* bgr, int h, int w, int row_stride, const float* boxes,
                       int n, uint8_t* crops_out, size_t crops_capacity, int* crop_dims_out);

/* TextClassifier.__call__ (predict_cls.py:44-89) without the image rotation side effect:
 * crops are HWC uint8 BGR, crop i at crops[i] with dims (hs[i], ws[i]) densely packed.
 * labels_out: int32 [n] (index into label_list), scores_out: float32 [n]. */
int b200ocr_cls_crops(b200ocr_engine* e, const uint8_t* const* crops, const int* hs, const int* ws, int n,
                      int batch_num, int32_t* labels_out, float* scores_out);

/* TextRecognizer.__call__ + CTCLabelDecode (predict_rec.py:280-326, rec_postprocess.py:98-109,
 * 58-86): aspect sort, batches of `batch_num` with the reference's imgW rule, resize/normalise/pad,
 * SVTR_LCNet forward, greedy CTC collapse on the device.  For crop i: ids_out[i*max_len ..] holds
 * `lens_out[i]` class indices (0 = blank never appears), scores_out[i] = mean prob of kept steps
 * (0 when empty).  rotate180[i] != 0 rotates crop i by 180 degrees first (cls result, predict_cls.py:85-88);
 * may be NULL. */
int b200ocr_rec_crops(b200ocr_engine* e, const uint8_t* const* crops, const int* hs, const int* ws, int n,
                      int batch_num, const uint8_t* rotate180, int32_t* ids_out, int max_len, int* lens_out,
                      float* scores_out);

/* session.run on the rec graph fused with `preds.argmax(axis=2)` / `preds.max(axis=2)` (predict_rec.py:316,
 * rec_postprocess.py:100-101): for the NCHW float32 batch `nchw` [dims] writes, per sample and time step, the class
 * index and its softmax probability (idx_out / prob_out: [dims[0]][T], T returned in *steps_out).  The
 * [n, T, classes] probability tensor is never materialised in bf16 mode (tcgen05 GEMM with an online
 * softmax / argmax epilogue).  Call with idx_out == NULL to query T. */
int b200ocr_rec_argmax(b200ocr_engine* e, const float* nchw, const int64_t dims[4], int32_t* idx_out, float* prob_out,
                       int capacity, int* steps_out);

/* ---- fused end to end (Level C) --------------------------------------------------------- */
typedef struct b200ocr_ocr_params {
    b200ocr_det_params det;
    int use_angle_cls;        /* args.use_angle_cls (predict_system.py:20,57): run the classifier and rotate flagged crops */
    float cls_thresh;         /* args.cls_thresh 0.9: rotate when '180' in label and score > cls_thresh (predict_cls.py:85) */
    unsigned cls_rot_label_mask; /* bit i set <=> '180' in label_list[i]  (label_list ['0','180'] -> 0x2)  utils.py:351 */
    int rec_batch_num;        /* args.rec_batch_num 6: fixes each crop's padded width imgW exactly as the reference's
                                 per-page batches do (predict_rec.py:290-300); device batches are formed across pages */
    int rec_width_bucket;     /* 0: imgW exactly as the reference; k > 0: imgW rounded up to a multiple of k (fewer,
                                 larger device batches; results then differ from the reference like a different
                                 rec_batch_num would) */
} b200ocr_ocr_params;

void b200ocr_ocr_params_default(b200ocr_ocr_params* p);

/* TextSystem.__call__ (predict_system.py:35-72) for a batch of pages, device resident between the stages:
 * det -> filter_tag_det_res -> sorted_boxes -> get_rotate_crop_image (every crop of every page in one launch) ->
 * [cls -> rotate 180] -> rec -> CTC decode.  Lines are returned page after page in sorted_boxes order:
 * n_boxes_out[i] lines for page i; boxes_out float32 [line_cap][4][2]; ids_out int32 [line_cap][max_len] with
 * lens_out[l] valid entries; scores_out float32 [line_cap] (the caller applies drop_score, predict_system.py:66-70);
 * cls_labels_out / cls_scores_out (optional) [line_cap].  Needs the det and rec models (and cls when use_angle_cls). */
int b200ocr_ocr_pages(b200ocr_engine* e, const uint8_t* const* pages, const int* hs, const int* ws, const int* strides,
                      int n, const b200ocr_ocr_params* p, float* boxes_out, int line_cap, int* n_boxes_out,
                      int32_t* ids_out, int max_len, int* lens_out, float* scores_out, int32_t* cls_labels_out,
                      float* cls_scores_out);

/* ---- diagnostics ------------------------------------------------------------------------ */
/* Text dump of the canonical graph of a loaded model (for tests / INTEGRATION.md). */
int b200ocr_describe(b200ocr_engine* e, int kind, char* buf, size_t cap);
/* 2*MACs and kernel-launch count of the plan for a given input shape. */
int b200ocr_plan_stats(b200ocr_engine* e, int kind, int n, int h, int w, double* flops, int* launches,
                       double* arena_bytes);
/* Per-kernel device time of one plan, measured live with CUDA events between launches on the engine
 * stream (the plan's input buffer keeps whatever the last run left there).  For step i: ms_out[i] (average
 * over iters), flops_out[i] / bytes_out[i] = algorithmic work of that launch (2*MACs; unique input + output +
 * weight bytes at the storage type); names are written NUL-separated into names_buf.  Returns the step count. */
int b200ocr_profile_plan(b200ocr_engine* e, int kind, int n, int h, int w, int iters, float* ms_out, double* flops_out,
                         double* bytes_out, int capacity, char* names_buf, size_t names_cap);
/* Writes a buffer larger than L2 (256 MiB) on the engine stream: an L2 flush between timed iterations. */
/* Average device time (ms, CUDA events on the engine stream) of `iters` back-to-back launches of one plan, the way
 * the stages launch it (CUDA-graph replay); head != 0 selects the rec plan with the fused CTC head. */
int b200ocr_time_plan(b200ocr_engine* e, int kind, int n, int h, int w, int head, int iters, float* ms_avg);
int b200ocr_l2_flush(b200ocr_engine* e);
/* Page-locked host memory for callers that want the H2D copies of their page buffers to be truly async. */
void* b200ocr_host_alloc(size_t bytes);
void b200ocr_host_free(void* p);
/* Kernels launched by this engine since creation (bench.py's gpu_launches). */
int64_t b200ocr_launch_count(const b200ocr_engine* e);
/* Device time (ms, CUDA events on the engine stream) of the stages of the last fused call:
 * [0] preprocess  [1] network  [2] post-process  [3] total. */
int b200ocr_stage_times(b200ocr_engine* e, float ms_out[4]);
/* Host wall-clock split of the last b200ocr_ocr_pages call (ms): page upload enqueue, det incl. box read-back,
 * host box sorting / crop planning, warp + cls enqueue, rec enqueue, final drain. */
int b200ocr_ocr_stage_times(b200ocr_engine* e, float ms_out[6]);

/* Diagnostics: with B200OCR_TRACE=<file> in the environment the stages record CUDA events around the work they enqueue
 * (det network, DBPostProcess, crop warp, cls, every rec batch) on each stream; this appends one JSON line per scope
 * (name, tag, stream, start / end ms relative to the first scope) to the file and clears the log.  No-op otherwise. */
void b200ocr_trace_dump(void);

/* ---- device-resident benchmarking hooks (bench.py `value`: inputs already in HBM) ------- */
/* Uploads pages once; b200ocr_det_resident_run then repeats preprocess -> DBNet -> DBPostProcess
 * on the resident copies and returns the device time in ms (CUDA events on the engine stream). */
int b200ocr_det_resident_upload(b200ocr_engine* e, const uint8_t* const* pages, const int* hs, const int* ws,
                                const int* strides, int n);
int b200ocr_det_resident_run(b200ocr_engine* e, const b200ocr_det_params* p, int iters, float* ms_total,
                             int* n_boxes_total, float* net_ms_total);
/* The same through the streaming path (submit / collect, B200OCR_DET_DEPTH = 4 batches in flight by default): step k processes resident pages
 * [k*pages_per_step, (k+1)*pages_per_step) modulo the resident set, so uploading more page sets than the L2 holds
 * makes every step read its pages from HBM.  ms_total = device time of all `iters` steps (CUDA events). */
int b200ocr_det_resident_stream(b200ocr_engine* e, const b200ocr_det_params* p, int iters, int pages_per_step,
                                float* ms_total, int* n_boxes_total);

#ifdef __cplusplus
}
#endif
#endif /* B200OCR_H_ */

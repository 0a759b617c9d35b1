// Pixel pre-processing kernels: cv2-exact bilinear resize fused with normalisation and layout
// (det: operators.py:54-138,22-31,170-176; rec/cls: predict_rec.py:54-80, predict_cls.py:22-42),
// cv2-exact bicubic perspective warp (utils.py:12-53) and the greedy CTC decode
// (rec_postprocess.py:58-109).  Integer recipes follow OpenCV's published fixed-point
// INTER_LINEAR / INTER_CUBIC algorithms and are pinned against the
// installed cv2 through oracle/restate_pre.py.
#pragma once
#include "common.cuh"

namespace b200ocr {

// One axis of cv2's INTER_LINEAR: tap indices and 11-bit weights for destination index d.
// `horizontal` selects the x-axis clamping rule (weights zeroed at the borders) vs the y-axis rule.
__device__ __forceinline__ void linear_taps(int d, int src, double scale, bool horizontal, int* i0, int* i1, int* w0,
                                            int* w1) {
    float f = (float)__dsub_rn(__dmul_rn(__dadd_rn((double)d, 0.5), scale), 0.5);
    int s = (int)floorf(f);
    f = __fsub_rn(f, (float)s);
    if (horizontal) {
        if (s < 0) { f = 0.f; s = 0; }
        if (s >= src - 1) { f = 0.f; s = src - 1; }
        *i0 = s;
        *i1 = min(s + 1, src - 1);
    } else {
        *i0 = min(max(s, 0), src - 1);
        *i1 = min(max(s + 1, 0), src - 1);
    }
    int a = (int)rintf(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
    int b = (int)rintf(__fmul_rn(f, 2048.f));
    *w0 = min(max(a, -32768), 32767);
    *w1 = min(max(b, -32768), 32767);
}

struct SrcView {             // HWC u8 image with optional virtual zero padding and 180-degree rotation
    const uint8_t* p;
    int h, w;                // real size
    int stride;              // bytes per row
    int ph, pw;              // padded (virtual) size >= real size; pixels outside the real image read 0
    int rot180;
};

__device__ __forceinline__ void load_px(const SrcView& s, int y, int x, int* v) {
    if (y >= s.h || x >= s.w) { v[0] = v[1] = v[2] = 0; return; }
    if (s.rot180) { y = s.h - 1 - y; x = s.w - 1 - x; }
    const uint8_t* q = s.p + (size_t)y * s.stride + (size_t)x * 3;
    v[0] = q[0]; v[1] = q[1]; v[2] = q[2];
}

// cv2.resize(src, (dw, dh)) value at destination (y, x): u8 x 3.
__device__ __forceinline__ void resize_px(const SrcView& s, int dh, int dw, double scale_y, double scale_x, int y, int x,
                                          int* out) {
    if (dh == s.ph && dw == s.pw) { load_px(s, y, x, out); return; }
    if (s.pw == 2 * dw && s.ph == 2 * dh) {          // cv2 switches to INTER_AREA for exact 2x decimation
        int a[3], b[3], c[3], d[3];
        load_px(s, 2 * y, 2 * x, a); load_px(s, 2 * y, 2 * x + 1, b);
        load_px(s, 2 * y + 1, 2 * x, c); load_px(s, 2 * y + 1, 2 * x + 1, d);
#pragma unroll
        for (int k = 0; k < 3; ++k) out[k] = (a[k] + b[k] + c[k] + d[k] + 2) >> 2;
        return;
    }
    int x0, x1, wx0, wx1, y0, y1, wy0, wy1;
    linear_taps(x, s.pw, scale_x, true, &x0, &x1, &wx0, &wx1);
    linear_taps(y, s.ph, scale_y, false, &y0, &y1, &wy0, &wy1);
    int p00[3], p01[3], p10[3], p11[3];
    load_px(s, y0, x0, p00); load_px(s, y0, x1, p01);
    load_px(s, y1, x0, p10); load_px(s, y1, x1, p11);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        int r0 = p00[k] * wx0 + p01[k] * wx1;
        int r1 = p10[k] * wx0 + p11[k] * wx1;
        int v = (((wy0 * (r0 >> 4)) >> 16) + ((wy1 * (r1 >> 4)) >> 16) + 2) >> 2;
        out[k] = min(max(v, 0), 255);
    }
}

struct NormP {               // out = ((u8 * mul) - sub[c]) / div[c]   (det)   or ((u8 / 255) - 0.5) / 0.5 (rec/cls)
    float mul;
    float sub[3], div[3];
    int rec_style;
};

__device__ __forceinline__ float norm_px(int v, int c, const NormP& n) {
    if (n.rec_style) return __fdiv_rn(__fsub_rn(__fdiv_rn((float)v, 255.f), 0.5f), 0.5f);
    return __fdiv_rn(__fsub_rn(__fmul_rn((float)v, n.mul), n.sub[c]), n.div[c]);
}

struct ResizeJob {           // one image of a batch
    SrcView src;
    int dh, dw;              // resize target
    double scale_y, scale_x; // 1 / (dst / src)
};

// Batch resize + normalise.  Output canvas per image is [OH][OW] (OW >= dw: zero padded on the right).
//   out_mode 0: NHWC storage T with cpad 4 (executor input)     1: NCHW float32 (onnxruntime-style tensor)
template <typename T>
__global__ void __launch_bounds__(256) resize_norm_kernel(const ResizeJob* __restrict__ jobs, int OH, int OW,
                                                           void* __restrict__ out, int out_mode, const NormP nrm) {
    const int b = blockIdx.z;
    const ResizeJob j = jobs[b];
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= OW || y >= OH) return;
    float v[3] = {0.f, 0.f, 0.f};
    if (x < j.dw && y < j.dh) {
        int px[3];
        resize_px(j.src, j.dh, j.dw, j.scale_y, j.scale_x, y, x, px);
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = norm_px(px[c], c, nrm);
    }
    if (out_mode == 0) {
        float o[4] = {v[0], v[1], v[2], 0.f};
        ST<T>::st4((T*)out + (((size_t)b * OH + y) * OW + x) * 4, o);
    } else {
        float* o = (float*)out + (size_t)b * 3 * OH * OW + (size_t)y * OW + x;
        o[0] = v[0];
        o[(size_t)OH * OW] = v[1];
        o[(size_t)2 * OH * OW] = v[2];
    }
}

// ---------------------------------------------------------------------------------------------
// get_rotate_crop_image: bicubic perspective warp, BORDER_REPLICATE, rot90 folded into the store
// ---------------------------------------------------------------------------------------------
struct WarpJob {
    double mi[9];            // inverse perspective matrix (dst -> src), row major
    const uint8_t* img;      // source page (HWC u8, device) — crops of several pages share one launch
    int h, w, stride;        // source size / bytes per row
    int cw, ch;              // warp output size before rotation
    int rot90;               // 1: store np.rot90(dst) (counter-clockwise), output is [cw][ch]
    size_t out_off;          // byte offset of this crop in the pool
};

static __global__ void __launch_bounds__(256) warp_cubic_kernel(const WarpJob* __restrict__ jobs, const short* __restrict__ tab,
                                                          uint8_t* __restrict__ pool) {
    const WarpJob& j = jobs[blockIdx.z];
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= j.cw || y >= j.ch) return;
    const uint8_t* __restrict__ img = j.img;
    const int h = j.h, w = j.w, stride = j.stride;
    double X0 = __dadd_rn(__dmul_rn(j.mi[0], (double)x), __dadd_rn(__dmul_rn(j.mi[1], (double)y), j.mi[2]));
    double Y0 = __dadd_rn(__dmul_rn(j.mi[3], (double)x), __dadd_rn(__dmul_rn(j.mi[4], (double)y), j.mi[5]));
    double W0 = __dadd_rn(__dmul_rn(j.mi[6], (double)x), __dadd_rn(__dmul_rn(j.mi[7], (double)y), j.mi[8]));
    double Wn = W0 != 0.0 ? __ddiv_rn(32.0, W0) : 0.0;
    double fX = fmin(fmax(__dmul_rn(X0, Wn), -2147483648.0), 2147483647.0);
    double fY = fmin(fmax(__dmul_rn(Y0, Wn), -2147483648.0), 2147483647.0);
    long long X = (long long)rint(fX), Y = (long long)rint(fY);
    int sx = (int)(X >> 5) - 1, sy = (int)(Y >> 5) - 1;
    const short* wt = tab + (((int)(Y & 31)) * 32 + (int)(X & 31)) * 16;
    int acc[3] = {0, 0, 0};
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
        int yy = min(max(sy + ky, 0), h - 1);
        const uint8_t* row = img + (size_t)yy * stride;
#pragma unroll
        for (int kx = 0; kx < 4; ++kx) {
            int xx = min(max(sx + kx, 0), w - 1);
            int wv = wt[ky * 4 + kx];
            const uint8_t* q = row + (size_t)xx * 3;
            acc[0] += q[0] * wv;
            acc[1] += q[1] * wv;
            acc[2] += q[2] * wv;
        }
    }
    uint8_t* o;
    if (j.rot90) o = pool + j.out_off + ((size_t)(j.cw - 1 - x) * j.ch + y) * 3;
    else o = pool + j.out_off + ((size_t)y * j.cw + x) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c] = (uint8_t)min(max((acc[c] + (1 << 14)) >> 15, 0), 255);
}

// ClsPostProcess (cls_postprocess.py:24-26): argmax over the class probabilities, first maximum wins.
static __global__ void cls_argmax_kernel(const float* __restrict__ probs, int b, int ncls, const int* __restrict__ slots,
                                  int* __restrict__ labels, float* __restrict__ scores) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= b) return;
    const float* r = probs + (size_t)k * ncls;
    int best = 0;
    for (int c = 1; c < ncls; ++c)
        if (r[c] > r[best]) best = c;
    labels[slots[k]] = best;
    scores[slots[k]] = r[best];
}

// predict_cls.py:85-88: crops whose label contains '180' with score > cls_thresh are rotated before recognition;
// the rotation is folded into the rec sampler (SrcView::rot180) instead of rewriting the crop.
static __global__ void rot_patch_kernel(ResizeJob* __restrict__ jobs, const int* __restrict__ slots, int n,
                                 const int* __restrict__ labels, const float* __restrict__ scores, unsigned rot_mask,
                                 float thresh) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int i = slots[k];
    jobs[k].src.rot180 = (((rot_mask >> labels[i]) & 1u) && scores[i] > thresh) ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// CTCLabelDecode: argmax/max over classes (first maximum wins, like numpy), then collapse.
// ---------------------------------------------------------------------------------------------
// probs: dense fp32 [rows][C]; one block per row.
static __global__ void __launch_bounds__(256) ctc_argmax_kernel(const float* __restrict__ probs, int C, int* __restrict__ idx,
                                                          float* __restrict__ mx) {
    __shared__ float sv[8];
    __shared__ int si[8];
    const float* r = probs + (size_t)blockIdx.x * C;
    float best = -1.f;
    int bi = 0x7fffffff;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float v = r[c];
        if (v > best) { best = v; bi = c; }     // strided scan keeps the lowest index per thread
    }
    for (int o = 16; o; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); ++k)
            if (sv[k] > best || (sv[k] == best && si[k] < bi)) { best = sv[k]; bi = si[k]; }
        idx[blockIdx.x] = bi;
        mx[blockIdx.x] = best;
    }
}

// numpy's float32 pairwise summation (np.mean -> np.add.reduce): blocks of 128 with 8 accumulators.
__device__ inline float np_pairwise_sum(const float* a, int n) {
    if (n < 8) {
        float r = 0.f;                                   // numpy starts from a[0]; -0.0 never occurs for probabilities
        if (n > 0) r = a[0];
        for (int i = 1; i < n; ++i) r = __fadd_rn(r, a[i]);
        return r;
    }
    if (n <= 128) {
        float r[8];
        for (int k = 0; k < 8; ++k) r[k] = a[k];
        int i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int k = 0; k < 8; ++k) r[k] = __fadd_rn(r[k], a[i + k]);
        float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                              __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
        for (; i < n; ++i) res = __fadd_rn(res, a[i]);
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __fadd_rn(np_pairwise_sum(a, n2), np_pairwise_sum(a + n2, n - n2));
}

// One thread per sequence: keep t where idx[t] != idx[t-1] and idx[t] != 0 (rec_postprocess.py:63-69),
// score = float32 mean of the kept probabilities (rec_postprocess.py:85), 0 when nothing is kept.
static __global__ void ctc_collapse_kernel(const int* __restrict__ idx, const float* __restrict__ mx, int B, int T,
                                    const int* __restrict__ dst_slot, int max_len, int* __restrict__ ids_out,
                                    int* __restrict__ lens_out, float* __restrict__ scores_out,
                                    float* __restrict__ scratch) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int* ii = idx + (size_t)b * T;
    const float* pp = mx + (size_t)b * T;
    float* keep = scratch + (size_t)b * T;
    int slot = dst_slot ? dst_slot[b] : b;
    int n = 0;
    for (int t = 0; t < T; ++t) {
        int v = ii[t];
        if (v != 0 && (t == 0 || v != ii[t - 1])) {
            if (n < max_len) ids_out[(size_t)slot * max_len + n] = v;
            keep[n] = pp[t];
            ++n;
        }
    }
    lens_out[slot] = n;                         // untruncated: callers detect n > max_len and retry with room
    scores_out[slot] = n ? __fdiv_rn(np_pairwise_sum(keep, n), (float)n) : 0.f;
}

}  // namespace b200ocr

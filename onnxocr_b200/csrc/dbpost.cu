// DBPostProcess on the device (onnxocr/db_postprocess.py:220-246, 104-149).
//
//   threshold (pred > thresh, float32)                         -> u8 bitmap           [HBM bound]
//   union-find labelling: fg 8-connected, bg 4-connected       -> roots = raster-first pixel of each component
//   candidates = fg roots (outer borders) + roots of bg components that do not touch the frame (hole borders),
//                ordered by descending start pixel == cv2.findContours(RETR_LIST) order, first max_candidates
//   per candidate: border following, convex hull, rotating calipers, box score, Clipper offset, calipers, scaling
//                (dbgeom.h — the literal restatement of the cv2 / Clipper arithmetic)
//
// Only the final [n_pages][max_candidates] box table crosses PCIe; the filter_tag_det_res tail runs on the host.
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <stdexcept>

#include "common.cuh"
#include "dbgeom.h"
#include "stages.h"

namespace b200ocr {

using namespace geom;

struct DbWork {
    DevBuf bitmap, labels, flags, pos, border, scan_tmp;
    DevBuf cand, counts, npts, offs, box1, valid, score, res_box, cbuf, found;
    DevBuf keys, ipool, fpool;
    PinBuf h_counts, h_res;
};

void db_free(DbWork* w) { delete w; }

namespace {

constexpr int kHoleBit = 1 << 30;
constexpr int kUnclipCap = 256;

// All per-pixel kernels index with 32 bits (db_post_enqueue rejects batches of 2^30 pixels or more): 64-bit divisions
// were the largest cost of the merge and compact kernels.
__global__ void threshold_kernel(const float* __restrict__ prob, uint8_t* __restrict__ bm, int64_t total, float thresh) {
    const unsigned quads = (unsigned)(total >> 2);
    for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(prob)[q];
        uchar4 o;
        o.x = v.x > thresh ? 1 : 0; o.y = v.y > thresh ? 1 : 0; o.z = v.z > thresh ? 1 : 0; o.w = v.w > thresh ? 1 : 0;
        reinterpret_cast<uchar4*>(bm)[q] = o;
    }
    if (blockIdx.x == 0 && threadIdx.x < (unsigned)(total & 3)) {
        const int64_t i = (total & ~(int64_t)3) + threadIdx.x;
        bm[i] = prob[i] > thresh ? 1 : 0;
    }
}

// cv2.dilate(mask, ones((2, 2))): anchor (1, 1) -> out(y, x) = max over in(y-1..y, x-1..x)
__global__ void dilate2x2_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int n, int h, int w) {
    int64_t total = (int64_t)n * h * w;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int x = (int)(i % w);
        int y = (int)((i / w) % h);
        uint8_t v = in[i];
        if (x > 0) v |= in[i - 1];
        if (y > 0) v |= in[i - w];
        if (x > 0 && y > 0) v |= in[i - w - 1];
        out[i] = v;
    }
}

__device__ __forceinline__ int find_root(const int* L, int i) {
    int p = L[i];
    while (p != i) {
        i = p;
        p = L[i];
    }
    return i;
}

// find with path halving (ECL-CC style): every visited node is re-pointed at its grandparent.  Parents only ever move
// to a smaller index inside the same tree, so the racing plain stores are benign (any value written is an ancestor) and
// the final roots — the raster-first pixel of each component — are unchanged; the chains the later unions and the
// flatten pass have to walk become short instead of O(rows).
__device__ __forceinline__ int find_root_halving(int* L, int i) {
    int curr = L[i];
    if (curr != i) {
        int prev = i, next;
        while (curr > (next = L[curr])) {
            L[prev] = next;
            prev = curr;
            curr = next;
        }
    }
    return curr;
}

__device__ __forceinline__ void unite(int* L, int a, int b) {
    for (;;) {
        a = find_root_halving(L, a);
        b = find_root_halving(L, b);
        if (a == b) return;
        if (a < b) {
            int old = atomicMin(&L[b], a);
            if (old == b) return;
            b = old;
        } else {
            int old = atomicMin(&L[a], b);
            if (old == a) return;
            a = old;
        }
    }
}

// label initialisation: every pixel points at the first pixel of its horizontal run inside its 32-pixel warp segment
// PROB != nullptr: the threshold is applied here as well (bm is written, not read) — one pass instead of two
template <bool FROM_PROB>
__global__ void ccl_init_kernel(const float* __restrict__ prob, float thresh, uint8_t* __restrict__ bm, int* __restrict__ L,
                                int64_t total, int w) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    bool in = i < (unsigned)total;
    int v = 2;
    if (in) {
        if (FROM_PROB) {
            v = prob[i] > thresh ? 1 : 0;
            bm[i] = (uint8_t)v;
        } else {
            v = bm[i];
        }
    }
    int x = in ? (int)(i % (unsigned)w) : 0;
    int lane = threadIdx.x & 31;
    // a run breaks where the class changes or a new image row starts
    int left = __shfl_up_sync(0xffffffffu, v, 1);
    bool brk = lane == 0 || left != v || x == 0;
    unsigned m = __ballot_sync(0xffffffffu, brk);
    unsigned upto = m & (0xffffffffu >> (31 - lane));
    int start_lane = 31 - __clz(upto);
    if (in) L[i] = (int)i - (lane - start_lane);
}

__global__ void __launch_bounds__(256) ccl_merge_kernel(const uint8_t* __restrict__ bm, int* __restrict__ L, int n, int h, int w) {
    const int total = n * h * w;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const unsigned row = (unsigned)i / (unsigned)w;
        const int x = i - (int)row * w, y = (int)(row % (unsigned)h);
        const uint8_t* q = bm + i;
        const int v = q[0];
        const bool hasN = y > 0, hasW = x > 0, hasE = x + 1 < w;
        const int vN = hasN ? q[-w] : -1, vW = hasW ? q[-1] : -1, vNW = hasN && hasW ? q[-w - 1] : -1;
        const bool N = vN == v, W = vW == v;
        // The link to the pixel above is implied when the left neighbour (same class, same 32-pixel segment, hence already
        // in this pixel's run) has the same class above it: i ~ i-1 ~ i-1-w ~ i-w.  Only the first pixel of such a chain
        // performs the union, which removes most of the atomic traffic.
        const bool implied = N && W && (i & 31) != 0 && vNW == v;
        if (v) {                                  // foreground: 8-connectivity
            if (implied) {
            } else if (N) {
                unite(L, i, i - w);
            } else {
                if (vNW > 0) unite(L, i, i - w - 1);
                if (hasN && hasE && q[-w + 1]) unite(L, i, i - w + 1);
            }
            // ccl_init_kernel already joined W inside a 32-pixel segment; join across segment boundaries here
            if (W && (i & 31) == 0) unite(L, i, i - 1);
        } else {                                  // background: 4-connectivity
            if (N && !implied) unite(L, i, i - w);
            if (W && (i & 31) == 0) unite(L, i, i - 1);
        }
    }
}

// Four pixels per thread (w % 4 == 0): two aligned 4-byte loads (this row, the row above) and three single bytes
// (left, up-left, up-right) instead of five byte loads per pixel; the links examined are exactly those of the scalar kernel.
__global__ void __launch_bounds__(256) ccl_merge4_kernel(const uint8_t* __restrict__ bm, int* __restrict__ L, int n, int h, int w) {
    const int quads = (n * h * w) >> 2;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += gridDim.x * blockDim.x) {
        const int i = q << 2;
        const unsigned row = (unsigned)i / (unsigned)w;
        const int x = i - (int)row * w, y = (int)(row % (unsigned)h);
        const uint32_t c = *reinterpret_cast<const uint32_t*>(bm + i);
        const bool hasN = y > 0;
        const uint32_t u = hasN ? *reinterpret_cast<const uint32_t*>(bm + i - w) : 0xffffffffu;     // 0xff never equals a class
        const int l = x > 0 ? bm[i - 1] : -1;
        const int ul = hasN && x > 0 ? bm[i - w - 1] : -1;
        const int ur = hasN && x + 4 < w ? bm[i - w + 4] : -1;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int v = (c >> (8 * k)) & 0xff;
            const int vN = hasN ? (int)((u >> (8 * k)) & 0xff) : -1;
            const int vW = k == 0 ? l : (int)((c >> (8 * (k - 1))) & 0xff);
            const int vNW = k == 0 ? ul : (hasN ? (int)((u >> (8 * (k - 1))) & 0xff) : -1);
            const int vNE = k == 3 ? ur : (hasN ? (int)((u >> (8 * (k + 1))) & 0xff) : -1);
            const bool N = vN == v, W = vW == v;
            const int p = i + k;
            const bool implied = N && W && (p & 31) != 0 && vNW == v;
            if (v) {                              // foreground: 8-connectivity
                if (implied) {
                } else if (N) {
                    unite(L, p, p - w);
                } else {
                    if (vNW > 0) unite(L, p, p - w - 1);
                    if (vNE > 0) unite(L, p, p - w + 1);
                }
                if (W && (p & 31) == 0) unite(L, p, p - 1);
            } else {                              // background: 4-connectivity
                if (N && !implied) unite(L, p, p - w);
                if (W && (p & 31) == 0) unite(L, p, p - 1);
            }
        }
    }
}

__global__ void ccl_flatten_kernel(int* __restrict__ L, int64_t total) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        L[i] = find_root(L, (int)i);
}

// background components touching the image frame are not holes
__global__ void border_kernel(const uint8_t* __restrict__ bm, const int* __restrict__ L, uint8_t* __restrict__ border, int n,
                              int h, int w) {
    int per = 2 * (h + w);
    int64_t total = (int64_t)n * per;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        int p = (int)(t / per), k = (int)(t % per);
        int x, y;
        if (k < w) { x = k; y = 0; }
        else if (k < 2 * w) { x = k - w; y = h - 1; }
        else if (k < 2 * w + h) { x = 0; y = k - 2 * w; }
        else { x = w - 1; y = k - 2 * w - h; }
        int64_t i = ((int64_t)p * h + y) * w + x;
        if (!bm[i]) border[find_root(L, (int)i)] = 1;     // the labels are not flattened: nothing else needs every pixel's root
    }
}

__global__ void cand_flag_kernel(const uint8_t* __restrict__ bm, const int* __restrict__ L, const uint8_t* __restrict__ border,
                                 int* __restrict__ flags, int64_t total) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        flags[i] = (L[i] == (int)i && (bm[i] || !border[i])) ? 1 : 0;
}

// cand[p][r]: r-th contour of page p in cv2 order (descending start pixel); counts[p] = min(total, maxc)
__global__ void compact_kernel(const uint8_t* __restrict__ bm, const int* __restrict__ flags, const int* __restrict__ pos,
                               int* __restrict__ cand, int* __restrict__ counts, int n, int hw, int maxc) {
    const unsigned total = (unsigned)n * (unsigned)hw;
    const unsigned gtid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gtid < (unsigned)n) {                                        // one thread per page: its candidate count
        const unsigned first = gtid * (unsigned)hw, last = first + hw - 1;
        const int page_total = pos[last] + flags[last] - pos[first];
        counts[gtid] = page_total < maxc ? page_total : maxc;
    }
    for (unsigned i = gtid; i < total; i += gridDim.x * blockDim.x) {
        if (!flags[i]) continue;                                     // almost every pixel stops here
        const unsigned p = i / (unsigned)hw, first = p * (unsigned)hw, last = first + hw - 1;
        const int page_total = pos[last] + flags[last] - pos[first];
        int r = page_total - 1 - (pos[i] - pos[first]);
        if (r >= maxc) continue;
        int local = (int)(i - first);
        cand[(size_t)p * maxc + r] = bm[i] ? local : ((local - 1) | kHoleBit);
    }
}

// Candidate list without the flag / scan / compact passes over every pixel: component roots (few hundred per page) are
// appended with an atomic counter and then sorted per page; the order cv2.findContours(RETR_LIST) returns is descending
// start pixel.  found[p] counts every qualifying root, so a page with more than kCandBuf of them (speckle maps) is
// detected and the batch is redone through the scan path at collection time.
constexpr int kCandBuf = 16384;
__global__ void cand_collect_kernel(const uint8_t* __restrict__ bm, const int* __restrict__ L, const uint8_t* __restrict__ border,
                                    int* __restrict__ buf, int* __restrict__ found, int n, int hw) {
    const unsigned total = (unsigned)n * (unsigned)hw;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        if (L[i] != (int)i) continue;
        if (!bm[i] && border[i]) continue;
        const unsigned p = i / (unsigned)hw;
        const int slot = atomicAdd(&found[p], 1);
        if (slot < kCandBuf) buf[(size_t)p * kCandBuf + slot] = (int)(i - p * (unsigned)hw);
    }
}

// one block per page: ascending bitonic sort of the page's roots, then cand[p][r] = r-th largest (cv2 order)
__global__ void __launch_bounds__(1024) cand_sort_kernel(const uint8_t* __restrict__ bm, int* __restrict__ buf,
                                                         const int* __restrict__ found, int* __restrict__ cand,
                                                         int* __restrict__ counts, int* __restrict__ overflow, int hw, int maxc) {
    __shared__ int sk[4096];
    const int p = blockIdx.x, tid = threadIdx.x;
    const int m_all = found[p];
    if (m_all > kCandBuf) {
        if (tid == 0) { atomicAdd(overflow, 1); counts[p] = 0; }
        return;
    }
    const int m = m_all;
    int* g = buf + (size_t)p * kCandBuf;
    int npow2 = 1;
    while (npow2 < m) npow2 <<= 1;
    int* a = npow2 <= 4096 ? sk : g;                       // large lists are sorted in place in global memory
    if (a == sk) {
        for (int i = tid; i < npow2; i += 1024) sk[i] = i < m ? g[i] : 0x7fffffff;
    } else {
        for (int i = m + tid; i < npow2; i += 1024) g[i] = 0x7fffffff;
    }
    __syncthreads();
    for (int k = 2; k <= npow2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < npow2; i += 1024) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const int x = a[i], y = a[ixj];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { a[i] = y; a[ixj] = x; }
                }
            }
            __syncthreads();
        }
    const int cnt = m < maxc ? m : maxc;
    if (tid == 0) counts[p] = cnt;
    const uint8_t* pbm = bm + (size_t)p * hw;
    for (int r = tid; r < cnt; r += 1024) {
        const int local = a[m - 1 - r];
        cand[(size_t)p * maxc + r] = pbm[local] ? local : ((local - 1) | kHoleBit);
    }
}

__global__ void trace_count_kernel(const uint8_t* __restrict__ bm, const int* __restrict__ cand, const int* __restrict__ counts,
                                   const uint8_t* __restrict__ valid, int* __restrict__ npts, int n, int h, int w, int maxc) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * maxc) return;
    int p = t / maxc, r = t % maxc;
    int cnt = 0;
    if (r < counts[p] && valid[t] == 2) {
        int c = cand[t];
        int local = c & (kHoleBit - 1);
        Bitmap img{bm + (size_t)p * h * w, h, w};
        cnt = trace_contour(img, local % w, local / w, (c & kHoleBit) != 0, nullptr, 0);
    }
    npts[t] = cnt;
}

// phase A: contour -> minAreaRect -> get_mini_boxes
__global__ void contour_box_kernel(const uint8_t* __restrict__ bm, const int* __restrict__ cand, const int* __restrict__ counts,
                                   const int* __restrict__ npts, const int* __restrict__ offs, uint64_t* __restrict__ keys,
                                   int* __restrict__ ipool, float* __restrict__ fpool, float* __restrict__ box1,
                                   uint8_t* __restrict__ valid, int n, int h, int w, int maxc, float min_size) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * maxc) return;
    int p = t / maxc, r = t % maxc;
    if (valid[t] != 2) return;                         // only contours the warp kernel could not hold in shared memory
    valid[t] = 0;
    if (r >= counts[p]) return;
    int np = npts[t];
    int64_t off = offs[t];
    uint64_t* k = keys + off;
    int* stack = ipool + 2 * off + 2 * (int64_t)t;
    int* hull = stack + np + 2;
    PtF* hp = (PtF*)(fpool + 5 * off);
    PtF* vect = hp + np;
    float* inv = (float*)(vect + np);
    int c = cand[t];
    int local = c & (kHoleBit - 1);
    Bitmap img{bm + (size_t)p * h * w, h, w};
    trace_contour(img, local % w, local / w, (c & kHoleBit) != 0, k, np);
    RotRect rr = min_area_rect_keys(k, np, stack, hull, hp, vect, inv);
    PtF box[4];
    float sside = mini_box(rr, box);
    if (sside < min_size) return;                     // db_postprocess.py:127
    for (int i = 0; i < 4; ++i) {
        box1[(size_t)t * 8 + 2 * i] = box[i].x;
        box1[(size_t)t * 8 + 2 * i + 1] = box[i].y;
    }
    valid[t] = 1;
}

// phase A, fast path: one warp per contour.  The 8 neighbours of the current border pixel are fetched by 8 lanes at
// once (one memory round trip per step instead of up to 8 dependent ones), contour points go to shared memory, the
// (x, y, index) sort is a warp bitonic sort, hull + calipers run on lane 0 out of shared memory.  Contours with more than
// kWarpCap vertices or kWarpHull hull points are flagged and redone by the generic kernels above.
constexpr int kWarpCap = 1024;
constexpr int kWarpHull = 256;
constexpr int kWarpsPerBlock = 4;
constexpr size_t kWarpSmem = (size_t)kWarpCap * 8 + (size_t)(kWarpCap + 4) * 4 + (size_t)kWarpCap * 4 + (size_t)kWarpHull * 20;
static_assert((size_t)kWarpHull * 20 >= (size_t)(kWarpCap + 4) * 4 + 16, "the lower Sklansky chains borrow the calipers' scratch");

__device__ __forceinline__ unsigned neighbour_mask(const Bitmap& img, int x, int y, int lane) {
    int v = 0;
    if (lane < 8) v = img.at(x + dir_dx(lane), y + dir_dy(lane));
    return __ballot_sync(0xffffffffu, v != 0) & 0xffu;
}

// A 32 x 32 pixel window of the bitmap held by the warp in registers, one 32-bit row mask per lane.  Border following is
// serial and every step needs the 8 neighbours of the current pixel: out of the window that is three shuffles instead of
// a dependent global-memory round trip (the longest contour of a batch IS the latency of DBPostProcess).  The window is
// re-loaded (one round trip: eight 4-byte loads per lane) when the walk comes within one pixel of its edge, placed ahead
// of the walk in its current direction.
struct TraceWindow {
    int x0 = 0, y0 = 0;        // origin; x0 is a multiple of 4
    unsigned bits = 0;         // this lane's row: bit c = pixel (x0 + c, y0 + lane), 0 outside the image
    bool valid = false, usable = false;
};

__device__ __forceinline__ void window_load(const Bitmap& img, TraceWindow& wd, int cx, int cy, int s, int lane) {
    const int x0 = (cx - 16 + 10 * dir_dx(s)) & ~3, y0 = cy - 16 + 10 * dir_dy(s);
    const int y = y0 + lane;
    unsigned bits = 0;
    if ((unsigned)y < (unsigned)img.h) {
        const uint8_t* row = img.p + (size_t)y * img.w;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int x = x0 + 4 * k;
            if (x >= 0 && x < img.w) {           // w % 4 == 0 and x % 4 == 0: the word is entirely inside the row
                const unsigned wv = *reinterpret_cast<const unsigned*>(row + x) & 0x01010101u;
                bits |= ((wv * 0x10204080u) >> 28) << (4 * k);
            }
        }
    }
    wd.x0 = x0; wd.y0 = y0; wd.bits = bits; wd.valid = true;
}

// bit d = pixel at direction d from (x, y) — the same value neighbour_mask returns
__device__ __forceinline__ unsigned window_mask(const Bitmap& img, TraceWindow& wd, int x, int y, int s, int lane) {
    if (!wd.usable) return neighbour_mask(img, x, y, lane);
    int cx = x - wd.x0, cy = y - wd.y0;
    if (!wd.valid || cx < 1 || cx > 30 || cy < 1 || cy > 30) {
        window_load(img, wd, x, y, s, lane);
        cx = x - wd.x0;
        cy = y - wd.y0;
    }
    const unsigned u = (__shfl_sync(0xffffffffu, wd.bits, cy - 1) >> (cx - 1)) & 7u;
    const unsigned m = (__shfl_sync(0xffffffffu, wd.bits, cy) >> (cx - 1)) & 7u;
    const unsigned d = (__shfl_sync(0xffffffffu, wd.bits, cy + 1) >> (cx - 1)) & 7u;
    // 0 E, 1 NE, 2 N, 3 NW, 4 W, 5 SW, 6 S, 7 SE (dir_dx / dir_dy)
    return ((m >> 2) & 1u) | (((u >> 2) & 1u) << 1) | (((u >> 1) & 1u) << 2) | ((u & 1u) << 3) | ((m & 1u) << 4) |
           ((d & 1u) << 5) | (((d >> 1) & 1u) << 6) | (((d >> 2) & 1u) << 7);
}

// Warp-cooperative twin of geom::trace_contour (identical point sequence); all lanes return the point count.
__device__ int warp_trace_contour(const Bitmap& img, int sx, int sy, bool is_hole, uint64_t* keys, int cap, int lane) {
    int s_end = is_hole ? 0 : 4, s = s_end;
    TraceWindow wd;
    wd.usable = (img.w & 3) == 0 && (reinterpret_cast<uintptr_t>(img.p) & 3) == 0;
    unsigned mask = window_mask(img, wd, sx, sy, s, lane);
    do {
        s = (s - 1) & 7;
    } while (!((mask >> s) & 1u) && s != s_end);
    int n = 0;
    if (s == s_end) {
        if (lane == 0 && cap > 0) keys[0] = pack_key(sx, sy, 0);
        return 1;
    }
    const int i1x = sx + dir_dx(s), i1y = sy + dir_dy(s);
    int i3x = sx, i3y = sy;
    int prev_s = s ^ 4;
    int ptx = sx, pty = sy;
    for (;;) {
        // first set direction among s+1, s+2, ... (cyclic); at least the pixel we came from is set
        unsigned rot = ((mask | (mask << 8)) >> ((s + 1) & 7)) & 0xffu;
        int step = __ffs(rot) - 1;
        s = (s + 1 + step) & 7;
        int i4x = i3x + dir_dx(s), i4y = i3y + dir_dy(s);
        const bool turned = s != prev_s;
        if (turned) {
            if (lane == 0 && n < cap) keys[n] = pack_key(ptx, pty, n);
            ++n;
            prev_s = s;
        }
        ptx += dir_dx(s);
        pty += dir_dy(s);
        if (i4x == sx && i4y == sy && i3x == i1x && i3y == i1y) break;
        i3x = i4x;
        i3y = i4y;
        // Straight horizontal runs, many pixels per round trip.  Border following is serial (one dependent memory round
        // trip per pixel) and the longest contour of a page IS the latency of DBPostProcess; text lines are long
        // horizontal blobs.  Moving east (s == 0) the walk is on a bottom edge: it keeps going east exactly while the
        // pixels SW, S, SE are empty and E is set (the scan from the reversed direction 4 visits 5, 6, 7, 0); moving west
        // (s == 4) it is on a top edge: NE, N, NW empty and W set.  32 lanes test the next 32 pixels of both rows, the
        // walk jumps over the whole stretch that satisfies the condition — no direction change, so no contour point is
        // lost (CHAIN_APPROX_SIMPLE) and the point sequence is identical to the pixel-by-pixel walk.
        if (!turned && (s == 0 || s == 4)) {
            const int dir = s == 0 ? 1 : -1, oy = s == 0 ? 1 : -1;
            const int a = img.at(i3x + dir * (1 + lane), i3y), b = img.at(i3x + dir * (lane - 1), i3y + oy);
            const unsigned A = __ballot_sync(0xffffffffu, a != 0), B = __ballot_sync(0xffffffffu, b != 0);
            const int ones = ~A ? __ffs(~A) - 1 : 32, zeros = B ? __ffs(B) - 1 : 32;
            int k = min(ones, zeros - 2);
            // never run over the start pixel or the first step's pixel: the termination test belongs to the ordinary step
            if (i3y == sy && (sx - i3x) * dir > 0) k = min(k, (sx - i3x) * dir - 1);
            if (i3y == i1y && (i1x - i3x) * dir > 0) k = min(k, (i1x - i3x) * dir - 1);
            if (k > 0) {
                i3x += dir * k;
                ptx += dir * k;
            }
        }
        mask = window_mask(img, wd, i3x, i3y, s, lane);
        s = (s + 4) & 7;
    }
    return n;
}

__device__ void warp_bitonic_sort(uint64_t* a, int npow2, int lane) {
    for (int k = 2; k <= npow2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < npow2; i += 32) {
                int ixj = i ^ j;
                if (ixj > i) {
                    uint64_t x = a[i], y = a[ixj];
                    bool up = (i & k) == 0;
                    if ((x > y) == up) { a[i] = y; a[ixj] = x; }
                }
            }
            __syncwarp();
        }
}

// cv::convexHull on the sorted keys with the four Sklansky chains (top-left, top-right, bottom-left, bottom-right) walked
// by four lanes at once instead of one after the other on lane 0: the chains are independent, each is as serial as
// border following, and together they were the longest phase of the kernel.  The chain stacks live where convex_hull puts
// them (top-right right behind top-left, ...) once their lengths are known, so the assembly below is geom::convex_hull's
// own code path: `stack` and `stack2` (n + 4 ints each) hold the upper and the lower pair.
__device__ int warp_convex_hull(uint64_t* keys, int n, int* stack, int* stack2, int* hull, int lane) {
    if (n == 0) return 0;
    // first index of the minimum / maximum y (the serial scan keeps the first one: strict comparisons)
    int miny = 0, maxy = 0;
    {
        int bmin = 0x7fffffff, imin = 0x7fffffff, bmax = -0x7fffffff, imax = 0x7fffffff;
        for (int i = lane; i < n; i += 32) {
            const int y = key_y(keys[i]);
            if (y < bmin) { bmin = y; imin = i; }
            if (y > bmax) { bmax = y; imax = i; }
        }
        for (int o = 16; o; o >>= 1) {
            const int y1 = __shfl_xor_sync(0xffffffffu, bmin, o), i1 = __shfl_xor_sync(0xffffffffu, imin, o);
            if (y1 < bmin || (y1 == bmin && i1 < imin)) { bmin = y1; imin = i1; }
            const int y2 = __shfl_xor_sync(0xffffffffu, bmax, o), i2 = __shfl_xor_sync(0xffffffffu, imax, o);
            if (y2 > bmax || (y2 == bmax && i2 < imax)) { bmax = y2; imax = i2; }
        }
        miny = imin;
        maxy = imax;
    }
    if (key_x(keys[0]) == key_x(keys[n - 1]) && key_y(keys[0]) == key_y(keys[n - 1])) {
        if (lane == 0) hull[0] = 0;
        return 1;
    }
    // the four chains, each into its own region: tl at stack[0..], tr at stack[maxy + 3 ..], bl / br likewise in stack2
    int cnt = 0;
    if (lane == 0) cnt = sklansky(keys, 0, maxy, stack, -1, 1);
    else if (lane == 1) cnt = sklansky(keys, n - 1, maxy, stack + maxy + 3, -1, -1);
    else if (lane == 2) cnt = sklansky(keys, 0, miny, stack2, 1, -1);
    else if (lane == 3) cnt = sklansky(keys, n - 1, miny, stack2 + miny + 3, 1, 1);
    __syncwarp();
    int tl_count = __shfl_sync(0xffffffffu, cnt, 0), tr_count = __shfl_sync(0xffffffffu, cnt, 1);
    int bl_count = __shfl_sync(0xffffffffu, cnt, 2), br_count = __shfl_sync(0xffffffffu, cnt, 3);
    int nout = 0;
    if (lane == 0) {
        // geom::convex_hull with clockwise = false from here on (the upper pair is swapped)
        const int* tl_stack = stack + maxy + 3;            // <- tr
        const int* tr_stack = stack;                       // <- tl
        { int c = tl_count; tl_count = tr_count; tr_count = c; }
        for (int i = 0; i < tl_count - 1; ++i) hull[nout++] = tl_stack[i];
        for (int i = tr_count - 1; i > 0; --i) hull[nout++] = tr_stack[i];
        const int stop_idx = tr_count > 2 ? tr_stack[1] : tl_count > 2 ? tl_stack[tl_count - 2] : -1;
        const int* bl_stack = stack2;
        const int* br_stack = stack2 + miny + 3;
        if (stop_idx >= 0) {
            // convex_hull reads br_stack[2 - bl_count] with br_stack == bl_stack + bl_count, i.e. entry 2 of the pair
            const int check_idx = bl_count > 2 ? bl_stack[1] : bl_count + br_count > 2 ? br_stack[2 - bl_count] : -1;
            if (check_idx == stop_idx ||
                (check_idx >= 0 && key_x(keys[check_idx]) == key_x(keys[stop_idx]) && key_y(keys[check_idx]) == key_y(keys[stop_idx]))) {
                bl_count = bl_count < 2 ? bl_count : 2;
                br_count = br_count < 2 ? br_count : 2;
            }
        }
        for (int i = 0; i < bl_count - 1; ++i) hull[nout++] = bl_stack[i];
        for (int i = br_count - 1; i > 0; --i) hull[nout++] = br_stack[i];
        nout = hull_index_shift(keys, hull, nout, stack);
    }
    return __shfl_sync(0xffffffffu, nout, 0);
}

// One warp per candidate contour; the grid is persistent (a few blocks per SM) and every warp walks the dense list of the
// batch's candidates (page p contributes counts[p] of its maxc slots), so no block is launched for the empty slots.
__global__ void __launch_bounds__(32 * kWarpsPerBlock) contour_warp_kernel(
    const uint8_t* __restrict__ bm, const int* __restrict__ cand, const int* __restrict__ counts, float* __restrict__ box1,
    uint8_t* __restrict__ valid, int* __restrict__ n_flagged, int n, int h, int w, int maxc, float min_size, int dbg_phase) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* base = smem + (size_t)warp * kWarpSmem;
    uint64_t* keys = (uint64_t*)base;
    int* stack = (int*)(base + (size_t)kWarpCap * 8);
    int* hull = stack + kWarpCap + 4;
    PtF* hp = (PtF*)(hull + kWarpCap);
    PtF* vect = hp + kWarpHull;
    float* inv = (float*)(vect + kWarpHull);
    int* stack2 = (int*)hp;                                // lower chains: the calipers' scratch is free until the hull exists
    // slots beyond a page's count are simply invalid
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n * maxc; t += gridDim.x * blockDim.x)
        if (t % maxc >= counts[t / maxc]) valid[t] = 0;
    int total = 0;
    for (int p = 0; p < n; ++p) total += counts[p];
    for (int j = blockIdx.x * kWarpsPerBlock + warp; j < total; j += gridDim.x * kWarpsPerBlock) {
        int p = 0, r = j;
        while (r >= counts[p]) { r -= counts[p]; ++p; }
        const int t = p * maxc + r;
        const int c = cand[t];
        const int local = c & (kHoleBit - 1);
        Bitmap img{bm + (size_t)p * h * w, h, w};
        __syncwarp();
        int np = warp_trace_contour(img, local % w, local / w, (c & kHoleBit) != 0, keys, kWarpCap, lane);
        if (np > kWarpCap) {
            if (lane == 0) { valid[t] = 2; atomicAdd(n_flagged, 1); }
            continue;
        }
        __syncwarp();
        if (dbg_phase == 1) { if (lane == 0) valid[t] = 0; continue; }      // timing experiments only (B200OCR_DBG_CONTOUR)
        int npow2 = 1;
        while (npow2 < np) npow2 <<= 1;
        for (int i = np + lane; i < npow2; i += 32) keys[i] = ~(uint64_t)0;
        __syncwarp();
        warp_bitonic_sort(keys, npow2, lane);
        if (dbg_phase == 2) { if (lane == 0) valid[t] = 0; continue; }
        const int nh = warp_convex_hull(keys, np, stack, stack2, hull, lane);
        if (dbg_phase == 3) { if (lane == 0) valid[t] = 0; continue; }
        if (lane == 0) {
            if (nh > kWarpHull) {
                valid[t] = 2;
                atomicAdd(n_flagged, 1);
            } else {
                for (int i = 0; i < nh; ++i) {
                    hp[i].x = (float)key_x(keys[hull[i]]);
                    hp[i].y = (float)key_y(keys[hull[i]]);
                }
                RotRect rr = min_area_rect(hp, nh, vect, inv);
                PtF box[4];
                float sside = mini_box(rr, box);
                if (sside < min_size) {
                    valid[t] = 0;
                } else {
                    for (int i = 0; i < 4; ++i) {
                        box1[(size_t)t * 8 + 2 * i] = box[i].x;
                        box1[(size_t)t * 8 + 2 * i + 1] = box[i].y;
                    }
                    valid[t] = 1;
                }
            }
        }
    }
}

// phase B: box_score_fast, one block per candidate (rows of the bounding box are spread over the threads;
// partial sums are combined in a fixed order, so the result is deterministic)
__global__ void __launch_bounds__(128) score_kernel(const float* __restrict__ prob, const float* __restrict__ box1,
                                                     uint8_t* __restrict__ valid, double* __restrict__ score, int h, int w,
                                                     int maxc, double box_thresh, const int* __restrict__ counts, int n) {
    // persistent grid over the dense candidate list (page p contributes counts[p] of its maxc slots)
    __shared__ double ssum[128];
    __shared__ long long scnt[128];
    int total = 0;
    for (int q = 0; q < n; ++q) total += counts[q];
    for (int j = blockIdx.x; j < total; j += gridDim.x) {
    int p = 0, r = j;
    while (r >= counts[p]) { r -= counts[p]; ++p; }
    const int t = p * maxc + r;
    if (valid[t] != 1) continue;
    const float* pred = prob + (size_t)p * h * w;
    PtF box[4];
    for (int i = 0; i < 4; ++i) {
        box[i].x = box1[(size_t)t * 8 + 2 * i];
        box[i].y = box1[(size_t)t * 8 + 2 * i + 1];
    }
    float fxmin = box[0].x, fxmax = box[0].x, fymin = box[0].y, fymax = box[0].y;
    for (int i = 1; i < 4; ++i) {
        fxmin = fminf(fxmin, box[i].x); fxmax = fmaxf(fxmax, box[i].x);
        fymin = fminf(fymin, box[i].y); fymax = fmaxf(fymax, box[i].y);
    }
    int xmin = min(max((int)floorf(fxmin), 0), w - 1), xmax = min(max((int)ceilf(fxmax), 0), w - 1);
    int ymin = min(max((int)floorf(fymin), 0), h - 1), ymax = min(max((int)ceilf(fymax), 0), h - 1);
    int qx[4], qy[4];
    for (int i = 0; i < 4; ++i) {
        qx[i] = (int)(box[i].x - (float)xmin);
        qy[i] = (int)(box[i].y - (float)ymin);
    }
    FillPoly4 fp;
    fp.init(qx, qy, xmax - xmin + 1, ymax - ymin + 1);
    double sum = 0;
    long long cnt = 0;
    int bw = xmax - xmin + 1, bh = ymax - ymin + 1;
    // thread -> (row, column stripe): keeps a warp on one row when the box is wide
    for (int y = threadIdx.x / 32; y < bh; y += blockDim.x / 32) {
        fp.set_row(y);
        for (int x = threadIdx.x % 32; x < bw; x += 32)
            if (fp.test(x, y)) {
                sum += (double)pred[(size_t)(y + ymin) * w + x + xmin];
                ++cnt;
            }
    }
    ssum[threadIdx.x] = sum;
    scnt[threadIdx.x] = cnt;
    __syncthreads();
    for (int s = 64; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            ssum[threadIdx.x] += ssum[threadIdx.x + s];
            scnt[threadIdx.x] += scnt[threadIdx.x + s];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double sc = scnt[0] ? ssum[0] / (double)scnt[0] : 0.0;
        score[t] = sc;
        if (box_thresh > sc) valid[t] = 0;            // db_postprocess.py:134
    }
    __syncthreads();                                  // ssum / scnt are reused by the next candidate
    }
}

// ---- det_db_score_mode = 'slow' (db_postprocess.py:129-132, 199-218): the score is the mean of pred over cv2.fillPoly of
// the contour itself.  Two kernels, run from db_post_collect (the contour points of all candidates go to one pool whose
// size the host must know): point counts, then trace + even-odd scan (geom::box_score_slow is the serial specification).
__global__ void __launch_bounds__(32 * kWarpsPerBlock) slow_count_kernel(const uint8_t* __restrict__ bm, const int* __restrict__ cand,
                                                                         const int* __restrict__ counts, const uint8_t* __restrict__ valid,
                                                                         int* __restrict__ npts, int n, int h, int w, int maxc) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int total = 0;
    for (int p = 0; p < n; ++p) total += counts[p];
    for (int j = blockIdx.x * kWarpsPerBlock + warp; j < total; j += gridDim.x * kWarpsPerBlock) {
        int p = 0, r = j;
        while (r >= counts[p]) { r -= counts[p]; ++p; }
        const int t = p * maxc + r;
        if (valid[t] != 1) continue;                  // npts was zeroed by the host
        const int c = cand[t];
        const int local = c & (kHoleBit - 1);
        Bitmap img{bm + (size_t)p * h * w, h, w};
        const int np = warp_trace_contour(img, local % w, local / w, (c & kHoleBit) != 0, nullptr, 0, lane);
        if (lane == 0) npts[t] = np;
    }
}

constexpr int kSlowWarps = 4;
// One block per candidate (persistent grid over the dense candidate list): warp 0 traces the contour into the pool, then
// each warp takes every fourth row of the bounding box: all edges are tested against the row (chain pixels are OR-ed,
// crossings XOR-ed into two bit rows in shared memory), a suffix parity over the crossing bits gives the interior, and
// the masked pixels are summed with coalesced loads.  Partial sums meet in a fixed order (deterministic).
__global__ void __launch_bounds__(32 * kSlowWarps) score_slow_kernel(
    const float* __restrict__ prob, const uint8_t* __restrict__ bm, const int* __restrict__ cand, const int* __restrict__ counts,
    const int* __restrict__ npts, const int* __restrict__ offs, uint64_t* __restrict__ pool, uint8_t* __restrict__ valid,
    double* __restrict__ score, int n, int h, int w, int maxc, double box_thresh, int row_words) {
    extern __shared__ uint32_t srows[];                // [kSlowWarps][2][row_words]
    __shared__ double ssum[32 * kSlowWarps];
    __shared__ long long scnt[32 * kSlowWarps];
    __shared__ int sbox[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* chain = srows + (size_t)warp * 2 * row_words;
    uint32_t* tog = chain + row_words;
    int total = 0;
    for (int q = 0; q < n; ++q) total += counts[q];
    for (int j = blockIdx.x; j < total; j += gridDim.x) {
        int p = 0, r = j;
        while (r >= counts[p]) { r -= counts[p]; ++p; }
        const int t = p * maxc + r;
        if (valid[t] != 1) continue;                  // block-uniform
        const int nv = npts[t];
        uint64_t* keys = pool + offs[t];
        if (warp == 0) {
            const int c = cand[t];
            const int local = c & (kHoleBit - 1);
            Bitmap img{bm + (size_t)p * h * w, h, w};
            warp_trace_contour(img, local % w, local / w, (c & kHoleBit) != 0, keys, nv, lane);
        }
        if (threadIdx.x == 0) { sbox[0] = 0x7fffffff; sbox[1] = -0x7fffffff; sbox[2] = 0x7fffffff; sbox[3] = -0x7fffffff; }
        __threadfence_block();
        __syncthreads();
        {
            int xmin = 0x7fffffff, xmax = -0x7fffffff, ymin = 0x7fffffff, ymax = -0x7fffffff;
            for (int i = threadIdx.x; i < nv; i += blockDim.x) {
                const uint64_t k = keys[i];
                const int x = key_x(k), y = key_y(k);
                xmin = min(xmin, x); xmax = max(xmax, x); ymin = min(ymin, y); ymax = max(ymax, y);
            }
            for (int o = 16; o; o >>= 1) {
                xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, o)); xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
                ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, o)); ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
            }
            if (lane == 0) { atomicMin(&sbox[0], xmin); atomicMax(&sbox[1], xmax); atomicMin(&sbox[2], ymin); atomicMax(&sbox[3], ymax); }
        }
        __syncthreads();
        const int xmin = sbox[0], xmax = sbox[1], ymin = sbox[2], ymax = sbox[3];
        const int nw = (xmax - xmin) / 32 + 1;
        const float* pred = prob + (size_t)p * h * w;
        double sum = 0;
        long long cnt = 0;
        for (int y = ymin + warp; y <= ymax; y += kSlowWarps) {
            for (int i = lane; i < nw; i += 32) chain[i] = tog[i] = 0u;
            __syncwarp();
            for (int i = lane; i < nv; i += 32) {
                const uint64_t k0 = keys[i], k1 = keys[i + 1 == nv ? 0 : i + 1];
                int xa, xb;
                const int f = edge_on_row(key_x(k0), key_y(k0), key_x(k1), key_y(k1), y, &xa, &xb);
                if (f & 1) {
                    xa -= xmin; xb -= xmin;
                    for (int wi = xa >> 5; wi <= (xb >> 5); ++wi) {
                        const int lo = max(xa - wi * 32, 0), hi = min(xb - wi * 32, 31);
                        atomicOr(&chain[wi], (0xffffffffu >> (31 - hi)) & (0xffffffffu << lo));
                    }
                    if (f & 2) atomicXor(&tog[xa >> 5], 1u << (xa & 31));
                }
            }
            __syncwarp();
            unsigned carry = 0;
            for (int base = ((nw - 1) >> 5) << 5; base >= 0; base -= 32) {
                const int wi = base + lane;
                const uint32_t tg = wi < nw ? tog[wi] : 0u;
                const unsigned odd = __ballot_sync(0xffffffffu, __popc(tg) & 1);
                const unsigned above = lane == 31 ? 0u : (unsigned)__popc(odd >> (lane + 1)) & 1u;
                uint32_t m = suffix_xor(tg) ^ tg;
                if (carry ^ above) m = ~m;
                m = wi < nw ? (m | chain[wi]) : 0u;
                carry ^= (unsigned)__popc(odd) & 1u;
                const unsigned any = __ballot_sync(0xffffffffu, m != 0u);
                for (unsigned rest = any; rest; rest &= rest - 1) {
                    const int jw = __ffs(rest) - 1;
                    const uint32_t mj = __shfl_sync(0xffffffffu, m, jw);
                    if ((mj >> lane) & 1u) {
                        sum += (double)pred[(size_t)y * w + xmin + (base + jw) * 32 + lane];
                        ++cnt;
                    }
                }
            }
            __syncwarp();
        }
        ssum[threadIdx.x] = sum;
        scnt[threadIdx.x] = cnt;
        __syncthreads();
        for (int s = 16 * kSlowWarps; s > 0; s >>= 1) {
            if (threadIdx.x < s) {
                ssum[threadIdx.x] += ssum[threadIdx.x + s];
                scnt[threadIdx.x] += scnt[threadIdx.x + s];
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            const double sc = scnt[0] ? ssum[0] / (double)scnt[0] : 0.0;
            score[t] = sc;
            if (box_thresh > sc) valid[t] = 0;        // db_postprocess.py:133
        }
        __syncthreads();
    }
}

// phase C: unclip -> minAreaRect -> get_mini_boxes -> scale/round/clip
__global__ void unclip_kernel(const float* __restrict__ box1, uint8_t* __restrict__ valid, int* __restrict__ res_box,
                              const int* __restrict__ src_hw, int n, int h, int w, int maxc, double unclip_ratio, float min_size2) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * maxc || valid[t] != 1) return;
    int p = t / maxc;
    PtF box[4];
    for (int i = 0; i < 4; ++i) {
        box[i].x = box1[(size_t)t * 8 + 2 * i];
        box[i].y = box1[(size_t)t * 8 + 2 * i + 1];
    }
    int ox[kUnclipCap], oy[kUnclipCap];
    int m = unclip_offset(box, unclip_ratio, ox, oy, kUnclipCap, nullptr);
    if (m <= 0 || m > kUnclipCap) { valid[t] = 0; return; }
    uint64_t k2[kUnclipCap];
    int stack[kUnclipCap + 2], hull[kUnclipCap];
    PtF hp[kUnclipCap], vect[kUnclipCap];
    float inv[kUnclipCap];
    for (int i = 0; i < m; ++i) k2[i] = pack_key(ox[i], oy[i], i);
    RotRect rr = min_area_rect_keys(k2, m, stack, hull, hp, vect, inv);
    PtF b2[4];
    float sside = mini_box(rr, b2);
    if (sside < min_size2) { valid[t] = 0; return; }  // db_postprocess.py:139
    double dest_h = (double)src_hw[2 * p], dest_w = (double)src_hw[2 * p + 1];
    for (int i = 0; i < 4; ++i) {
        res_box[(size_t)t * 8 + 2 * i] = scale_coord(b2[i].x, w, dest_w);
        res_box[(size_t)t * 8 + 2 * i + 1] = scale_coord(b2[i].y, h, dest_h);
    }
}

int grid_for(int64_t total, int block = 256) {
    int64_t g = (total + block - 1) / block;
    return (int)std::max<int64_t>(1, std::min<int64_t>(g, 148 * 16));
}

static void launch_contour_warp(const uint8_t* bm, const int* cand, const int* counts, float* box1, uint8_t* valid, int* d_flagged,
                                int n, int h, int w, int maxc, float min_size, int nt, cudaStream_t st) {
    static bool attr_set[16] = {false};             // per device: function attributes are per context
    int dev = 0;
    cudaGetDevice(&dev);
    size_t smem = kWarpSmem * kWarpsPerBlock;
    static const int dbg_phase = [] { const char* e = getenv("B200OCR_DBG_CONTOUR"); return e ? atoi(e) : 0; }();
    if (!attr_set[dev & 15]) {
        B200_CUDA_CHECK(cudaFuncSetAttribute(contour_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[dev & 15] = true;
    }
    contour_warp_kernel<<<(unsigned)std::min<int64_t>(ceil_div(nt, kWarpsPerBlock), 148 * 2), 32 * kWarpsPerBlock, smem, st>>>(
        bm, cand, counts, box1, valid, d_flagged, n, h, w, maxc, min_size, dbg_phase);
}

}  // namespace


void Stages::db_postprocess_device(const float* d_prob, int n, int h, int w, const int* src_hs, const int* src_ws,
                                   const b200ocr_det_params& p, int capacity, std::vector<int32_t>* boxes,
                                   std::vector<double>* scores, std::vector<int>* counts_out) {
    DbPending pend;
    // lane 0's stream must see what the caller enqueued on the engine stream (the probability map)
    B200_CUDA_CHECK(cudaEventRecord(ev_fork_, st_));
    B200_CUDA_CHECK(cudaStreamWaitEvent(det_st_[0], ev_fork_, 0));
    db_post_enqueue(0, d_prob, n, h, w, src_hs, src_ws, p, capacity, &pend);
    db_post_collect(&pend, boxes, scores, counts_out);
}

// Everything of DBPostProcess up to the result copies, enqueued on the lane's stream without any host synchronisation.
void Stages::db_post_enqueue(int lane, const float* d_prob, int n, int h, int w, const int* src_hs, const int* src_ws,
                             const b200ocr_det_params& p, int capacity, DbPending* pend) {
    if (!db_lane_[lane]) db_lane_[lane] = new DbWork();
    DbWork& W = *db_lane_[lane];
    if ((int64_t)n * h * w >= (int64_t)1 << 30) throw std::runtime_error("bad argument: probability-map batch too large for 30-bit pixel indices");
    if (h >= 60000 || w >= 60000) throw std::runtime_error("bad argument: probability map too large");
    const int64_t total = (int64_t)n * h * w;
    const int hw = h * w;
    const int maxc = std::max(1, std::min(capacity, p.max_candidates));
    cudaStream_t st = det_st_[lane];
    TraceScope tr_post("db_post", st, n);
    uint8_t* bm = (uint8_t*)W.bitmap.get((size_t)total * (p.use_dilation ? 2 : 1));
    int* L = (int*)W.labels.get((size_t)total * sizeof(int));
    uint8_t* border = (uint8_t*)W.border.get((size_t)total);
    if (p.use_dilation) {
        threshold_kernel<<<grid_for(total), 256, 0, st>>>(d_prob, bm, total, p.thresh);
        dilate2x2_kernel<<<grid_for(total), 256, 0, st>>>(bm, bm + total, n, h, w);
        bm += total;
        launches += 2;
        ccl_init_kernel<false><<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(nullptr, 0.f, bm, L, total, w);
    } else {
        ccl_init_kernel<true><<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(d_prob, p.thresh, bm, L, total, w);
    }
    if ((w & 3) == 0) ccl_merge4_kernel<<<grid_for(total >> 2), 256, 0, st>>>(bm, L, n, h, w);
    else ccl_merge_kernel<<<grid_for(total), 256, 0, st>>>(bm, L, n, h, w);
    B200_CUDA_CHECK(cudaMemsetAsync(border, 0, (size_t)total, st));
    border_kernel<<<grid_for((int64_t)n * 2 * (h + w)), 256, 0, st>>>(bm, L, border, n, h, w);
    const int nt = n * maxc;
    int* cand = (int*)W.cand.get((size_t)nt * sizeof(int));
    int* counts = (int*)W.counts.get((size_t)(n + 2 * n + 1) * sizeof(int));
    int* d_src = counts + n;
    int* npts = (int*)W.npts.get((size_t)(nt + 1) * sizeof(int));
    int* offs = (int*)W.offs.get((size_t)(nt + 1) * sizeof(int));
    // candidate roots: atomic append + per-page sort (found[n] is the overflow flag, see db_post_collect)
    int* cbuf = (int*)W.cbuf.get((size_t)n * kCandBuf * sizeof(int));
    int* found = (int*)W.found.get((size_t)(n + 1) * sizeof(int));
    B200_CUDA_CHECK(cudaMemsetAsync(found, 0, (size_t)(n + 1) * sizeof(int), st));
    cand_collect_kernel<<<grid_for(total), 256, 0, st>>>(bm, L, border, cbuf, found, n, hw);
    cand_sort_kernel<<<n, 1024, 0, st>>>(bm, cbuf, found, cand, counts, found + n, hw, maxc);
    float* box1 = (float*)W.box1.get((size_t)nt * 8 * sizeof(float));
    uint8_t* valid = (uint8_t*)W.valid.get((size_t)nt);
    double* score = (double*)W.score.get((size_t)nt * sizeof(double));
    int* res_box = (int*)W.res_box.get((size_t)nt * 8 * sizeof(int));
    int* d_flagged = npts + nt;                       // reuse the scan sentinel slot as the overflow counter
    B200_CUDA_CHECK(cudaMemsetAsync(d_flagged, 0, sizeof(int), st));
    int* h_counts = (int*)W.h_counts.get((size_t)(n + 3) * sizeof(int) + 2 * n * sizeof(int));
    int* h_src = h_counts + n + 3;                     // [0, n) counts, [n] flagged contours, [n + 1] scratch, [n + 2] candidate overflow
    for (int i = 0; i < n; ++i) {
        h_src[2 * i] = src_hs[i];
        h_src[2 * i + 1] = src_ws[i];
    }
    B200_CUDA_CHECK(cudaMemcpyAsync(d_src, h_src, 2 * n * sizeof(int), cudaMemcpyHostToDevice, st));
    launch_contour_warp(bm, cand, counts, box1, valid, d_flagged, n, h, w, maxc, (float)p.min_size, nt, st);
    B200_CUDA_CHECK(cudaMemcpyAsync(h_counts, counts, n * sizeof(int), cudaMemcpyDeviceToHost, st));
    B200_CUDA_CHECK(cudaMemcpyAsync(h_counts + n, d_flagged, sizeof(int), cudaMemcpyDeviceToHost, st));
    B200_CUDA_CHECK(cudaMemcpyAsync(h_counts + n + 2, found + n, sizeof(int), cudaMemcpyDeviceToHost, st));
    launches += 6;                                    // init, merge, border, collect, sort, contour (memsets and copies are not kernels)
    auto run_tail = [&]() {
        if (p.score_mode) return;                     // 'slow' scores need a host-sized point pool: db_post_collect
        score_kernel<<<(unsigned)std::min(nt, 148 * 8), 128, 0, st>>>(d_prob, box1, valid, score, h, w, maxc, p.box_thresh, counts, n);
        unclip_kernel<<<(unsigned)ceil_div(nt, 32), 32, 0, st>>>(box1, valid, res_box, d_src, n, h, w, maxc, p.unclip_ratio,
                                                                  (float)(p.min_size + 2));
        launches += 2;
    };
    run_tail();
    B200_CUDA_CHECK(cudaGetLastError());
    // results -> host
    size_t res_bytes = (size_t)nt * (8 * sizeof(int) + sizeof(double) + 1);
    char* h_res = (char*)W.h_res.get(res_bytes + 64);
    int* hb = (int*)h_res;
    double* hs = (double*)(h_res + (size_t)nt * 8 * sizeof(int));
    uint8_t* hv = (uint8_t*)(h_res + (size_t)nt * (8 * sizeof(int) + sizeof(double)));
    B200_CUDA_CHECK(cudaMemcpyAsync(hb, res_box, (size_t)nt * 8 * sizeof(int), cudaMemcpyDeviceToHost, st));
    B200_CUDA_CHECK(cudaMemcpyAsync(hs, score, (size_t)nt * sizeof(double), cudaMemcpyDeviceToHost, st));
    B200_CUDA_CHECK(cudaMemcpyAsync(hv, valid, (size_t)nt, cudaMemcpyDeviceToHost, st));
    pend->lane = lane; pend->n = n; pend->h = h; pend->w = w; pend->maxc = maxc; pend->nt = nt;
    pend->d_prob = d_prob; pend->p = p; pend->active = true;
}

// Waits for the lane, runs the rare large-contour fallback, and gathers the boxes of each page in cv2 order.
void Stages::db_post_collect(DbPending* pend, std::vector<int32_t>* boxes, std::vector<double>* scores,
                             std::vector<int>* counts_out) {
    if (!pend->active) throw std::runtime_error("internal: collecting a det lane that has nothing pending");
    pend->active = false;
    DbWork& W = *db_lane_[pend->lane];
    cudaStream_t st = det_st_[pend->lane];
    const int n = pend->n, h = pend->h, w = pend->w, maxc = pend->maxc, nt = pend->nt;
    const b200ocr_det_params& p = pend->p;
    const float* d_prob = pend->d_prob;
    const int64_t total = (int64_t)n * h * w;
    uint8_t* bm = (uint8_t*)W.bitmap.p + (p.use_dilation ? total : 0);
    int* cand = (int*)W.cand.p;
    int* counts = (int*)W.counts.p;
    int* d_src = counts + n;
    int* npts = (int*)W.npts.p;
    int* offs = (int*)W.offs.p;
    float* box1 = (float*)W.box1.p;
    uint8_t* valid = (uint8_t*)W.valid.p;
    double* score = (double*)W.score.p;
    int* res_box = (int*)W.res_box.p;
    int* h_counts = (int*)W.h_counts.p;
    char* h_res = (char*)W.h_res.p;
    int* hb = (int*)h_res;
    double* hs = (double*)(h_res + (size_t)nt * 8 * sizeof(int));
    uint8_t* hv = (uint8_t*)(h_res + (size_t)nt * (8 * sizeof(int) + sizeof(double)));
    size_t tmp_bytes = 0;
    auto run_tail = [&]() {
        if (p.score_mode) return;                     // 'slow' scores need a host-sized point pool: db_post_collect
        score_kernel<<<(unsigned)std::min(nt, 148 * 8), 128, 0, st>>>(d_prob, box1, valid, score, h, w, maxc, p.box_thresh, counts, n);
        unclip_kernel<<<(unsigned)ceil_div(nt, 32), 32, 0, st>>>(box1, valid, res_box, d_src, n, h, w, maxc, p.unclip_ratio,
                                                                  (float)(p.min_size + 2));
        launches += 2;
    };
    B200_CUDA_CHECK(cudaStreamSynchronize(st));
    if (h_counts[n + 2] > 0) {
        // rare: a page has more than kCandBuf component roots (speckle maps) -> candidates through flag / scan / compact over
        // every pixel (which has no capacity limit), then the per-candidate kernels again
        const int hw = h * w;
        int* L = (int*)W.labels.p;
        const uint8_t* border = (const uint8_t*)W.border.p;
        int* flags = (int*)W.flags.get((size_t)total * sizeof(int));
        int* pos = (int*)W.pos.get((size_t)total * sizeof(int));
        cand_flag_kernel<<<grid_for(total), 256, 0, st>>>(bm, L, border, flags, total);
        size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, flags, pos, (int)total, st);
        void* tmp = W.scan_tmp.get(tb + 256);
        cub::DeviceScan::ExclusiveSum(tmp, tb, flags, pos, (int)total, st);
        compact_kernel<<<grid_for(total), 256, 0, st>>>(bm, flags, pos, cand, counts, n, hw, maxc);
        int* d_flagged = npts + nt;
        B200_CUDA_CHECK(cudaMemsetAsync(d_flagged, 0, sizeof(int), st));
        launch_contour_warp(bm, cand, counts, box1, valid, d_flagged, n, h, w, maxc, (float)p.min_size, nt, st);
        B200_CUDA_CHECK(cudaMemcpyAsync(h_counts, counts, n * sizeof(int), cudaMemcpyDeviceToHost, st));
        B200_CUDA_CHECK(cudaMemcpyAsync(h_counts + n, d_flagged, sizeof(int), cudaMemcpyDeviceToHost, st));
        run_tail();
        B200_CUDA_CHECK(cudaMemcpyAsync(hb, res_box, (size_t)nt * 8 * sizeof(int), cudaMemcpyDeviceToHost, st));
        B200_CUDA_CHECK(cudaMemcpyAsync(hs, score, (size_t)nt * sizeof(double), cudaMemcpyDeviceToHost, st));
        B200_CUDA_CHECK(cudaMemcpyAsync(hv, valid, (size_t)nt, cudaMemcpyDeviceToHost, st));
        launches += 7;
        B200_CUDA_CHECK(cudaStreamSynchronize(st));
    }
    if (h_counts[n] > 0) {
        // rare: some contour exceeded the shared-memory capacity of the warp kernel -> generic two-pass path for those
        trace_count_kernel<<<(unsigned)ceil_div(nt, 64), 64, 0, st>>>(bm, cand, counts, valid, npts, n, h, w, maxc);
        B200_CUDA_CHECK(cudaMemsetAsync(npts + nt, 0, sizeof(int), st));
        size_t tmp2 = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp2, npts, offs, nt + 1, st);
        void* tmpb = W.scan_tmp.get(std::max(tmp_bytes, tmp2) + 256);
        cub::DeviceScan::ExclusiveSum(tmpb, tmp2, npts, offs, nt + 1, st);
        B200_CUDA_CHECK(cudaMemcpyAsync(h_counts + n + 1, offs + nt, sizeof(int), cudaMemcpyDeviceToHost, st));
        B200_CUDA_CHECK(cudaStreamSynchronize(st));
        int64_t total_pts = h_counts[n + 1];
        if (total_pts >= (1 << 20) * (int64_t)64) throw std::runtime_error("bad argument: contour point pool overflow");
        uint64_t* keys = (uint64_t*)W.keys.get((size_t)(total_pts + 1) * sizeof(uint64_t));
        int* ipool = (int*)W.ipool.get((size_t)(2 * total_pts + 2 * (int64_t)nt + 2) * sizeof(int));
        float* fpool = (float*)W.fpool.get((size_t)(5 * total_pts + 5) * sizeof(float));
        contour_box_kernel<<<(unsigned)ceil_div(nt, 64), 64, 0, st>>>(bm, cand, counts, npts, offs, keys, ipool, fpool, box1, valid,
                                                                       n, h, w, maxc, (float)p.min_size);
        launches += 4;
        run_tail();
        B200_CUDA_CHECK(cudaMemcpyAsync(hb, res_box, (size_t)nt * 8 * sizeof(int), cudaMemcpyDeviceToHost, st));
        B200_CUDA_CHECK(cudaMemcpyAsync(hs, score, (size_t)nt * sizeof(double), cudaMemcpyDeviceToHost, st));
        B200_CUDA_CHECK(cudaMemcpyAsync(hv, valid, (size_t)nt, cudaMemcpyDeviceToHost, st));
        B200_CUDA_CHECK(cudaStreamSynchronize(st));
    }
    if (p.score_mode) {
        // det_db_score_mode 'slow': box1 / valid hold the mini boxes of every candidate that passed the size test; score
        // them over the filled contour, then the common unclip tail.  Not the default, so it may cost a synchronisation.
        B200_CUDA_CHECK(cudaMemsetAsync(npts, 0, (size_t)(nt + 1) * sizeof(int), st));
        slow_count_kernel<<<(unsigned)std::min<int64_t>(ceil_div(nt, kWarpsPerBlock), 148 * 4), 32 * kWarpsPerBlock, 0, st>>>(
            bm, cand, counts, valid, npts, n, h, w, maxc);
        size_t tmp3 = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp3, npts, offs, nt + 1, st);
        void* tmpc = W.scan_tmp.get(tmp3 + 256);
        cub::DeviceScan::ExclusiveSum(tmpc, tmp3, npts, offs, nt + 1, st);
        B200_CUDA_CHECK(cudaMemcpyAsync(h_counts + n + 1, offs + nt, sizeof(int), cudaMemcpyDeviceToHost, st));
        B200_CUDA_CHECK(cudaStreamSynchronize(st));
        const int64_t total_pts = h_counts[n + 1];
        if (total_pts >= (1 << 20) * (int64_t)64) throw std::runtime_error("bad argument: contour point pool overflow");
        uint64_t* pool = (uint64_t*)W.keys.get((size_t)(total_pts + 1) * sizeof(uint64_t));
        const int row_words = w / 32 + 2;
        const size_t smem = (size_t)kSlowWarps * 2 * row_words * sizeof(uint32_t);
        static bool attr_set[16] = {false};
        int dev = 0;
        cudaGetDevice(&dev);
        if (!attr_set[dev & 15]) {                    // w < 60000: at most 60 KB
            B200_CUDA_CHECK(cudaFuncSetAttribute(score_slow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
            attr_set[dev & 15] = true;
        }
        score_slow_kernel<<<(unsigned)std::max(1, std::min(nt, 148 * 8)), 32 * kSlowWarps, smem, st>>>(
            d_prob, bm, cand, counts, npts, offs, pool, valid, score, n, h, w, maxc, p.box_thresh, row_words);
        unclip_kernel<<<(unsigned)ceil_div(nt, 32), 32, 0, st>>>(box1, valid, res_box, d_src, n, h, w, maxc, p.unclip_ratio,
                                                                  (float)(p.min_size + 2));
        launches += 4;
        B200_CUDA_CHECK(cudaGetLastError());
        B200_CUDA_CHECK(cudaMemcpyAsync(hb, res_box, (size_t)nt * 8 * sizeof(int), cudaMemcpyDeviceToHost, st));
        B200_CUDA_CHECK(cudaMemcpyAsync(hs, score, (size_t)nt * sizeof(double), cudaMemcpyDeviceToHost, st));
        B200_CUDA_CHECK(cudaMemcpyAsync(hv, valid, (size_t)nt, cudaMemcpyDeviceToHost, st));
        B200_CUDA_CHECK(cudaStreamSynchronize(st));
    }
    boxes->clear();
    scores->clear();
    counts_out->assign(n, 0);
    for (int pg = 0; pg < n; ++pg) {
        int c = h_counts[pg];
        for (int r = 0; r < c; ++r) {
            size_t t = (size_t)pg * maxc + r;
            if (hv[t] != 1) continue;
            boxes->insert(boxes->end(), hb + t * 8, hb + t * 8 + 8);
            scores->push_back(hs[t]);
            (*counts_out)[pg]++;
        }
    }
}

}  // namespace b200ocr

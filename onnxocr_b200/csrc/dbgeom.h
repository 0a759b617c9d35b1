// Per-contour geometry of DBPostProcess.boxes_from_bitmap (onnxocr/db_postprocess.py:104-197) as
// host+device code.  The third-party arithmetic the reference calls is restated literally, in the
// same operation order and precision, so that results are bit-identical to cv2 4.x on the host and
// on the device (this TU is compiled with -fmad=false / -ffp-contract=off):
//
//   cv2.findContours(RETR_LIST, CHAIN_APPROX_SIMPLE)   border following of one contour (Suzuki-Abe)
//   cv2.minAreaRect                                    convexHull (Sklansky) + rotatingCalipers
//   cv2.boxPoints                                      RotatedRect::points
//   cv2.fillPoly + cv2.mean                            closed convex interior + LINE_8 edge pixels
//   shapely Polygon.area / .length, pyclipper offset   shoelace / perimeter, ClipperOffset (JT_ROUND)
//
// The same header is compiled by g++ into tests' libb200geom_host.so, where every function is
// fuzzed against the installed cv2 (tests/test_geom_host.py).
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define B200_HD __host__ __device__ __forceinline__
#define B200_HDN __host__ __device__
#else
#define B200_HD inline
#define B200_HDN inline
#endif

namespace b200ocr {
namespace geom {

struct PtF {
    float x, y;
};

// ---------------------------------------------------------------------------------------------
// contour tracing
// ---------------------------------------------------------------------------------------------
struct Bitmap {
    const uint8_t* p;
    int h, w;
    B200_HD int at(int x, int y) const { return ((unsigned)x < (unsigned)w && (unsigned)y < (unsigned)h) ? p[(size_t)y * w + x] : 0; }
};

// key = x << 40 | y << 20 | index.  x, y are biased so that unclipped polygons (which may leave the image by
// their offset distance) stay non-negative; valid for coordinates in [-kKeyBias, 2^20 - kKeyBias).
constexpr int kKeyBias = 65536;
B200_HD uint64_t pack_key(int x, int y, int idx) {
    return ((uint64_t)(x + kKeyBias) << 40) | ((uint64_t)(y + kKeyBias) << 20) | (uint64_t)idx;
}
B200_HD int key_x(uint64_t k) { return (int)(k >> 40) - kKeyBias; }
B200_HD int key_y(uint64_t k) { return (int)((k >> 20) & 0xFFFFF) - kKeyBias; }
B200_HD int key_i(uint64_t k) { return (int)(k & 0xFFFFF); }

// Follows one border starting at (sx, sy) exactly like OpenCV's icvFetchContour with
// CV_CHAIN_APPROX_SIMPLE; writes keys[idx] = pack(x, y, idx) when `keys` is non-null.
// Returns the number of contour points (may exceed cap; only the first cap are stored).
// 8-neighbourhood step of direction s (0 = east, counter-clockwise on the image with y down: {1,1,0,-1,-1,-1,0,1} /
// {0,-1,-1,-1,0,1,1,1}) as arithmetic on packed 2-bit fields: a run-time indexed table would live in LOCAL memory on the
// device, i.e. several dependent L1 round trips in every step of the (inherently serial) border following.
B200_HD int dir_dx(int s) { return (int)((0x901Au >> (2 * (s & 7))) & 3u) - 1; }
B200_HD int dir_dy(int s) { return (int)((0xA901u >> (2 * (s & 7))) & 3u) - 1; }

B200_HDN int trace_contour(const Bitmap& img, int sx, int sy, bool is_hole, uint64_t* keys, int cap) {
    int s_end = is_hole ? 0 : 4, s = s_end;
    int i1x, i1y;
    do {
        s = (s - 1) & 7;
        i1x = sx + dir_dx(s);
        i1y = sy + dir_dy(s);
    } while (img.at(i1x, i1y) == 0 && s != s_end);
    int n = 0;
    if (s == s_end) {               // single pixel domain
        if (keys && n < cap) keys[n] = pack_key(sx, sy, n);
        return 1;
    }
    int i3x = sx, i3y = sy, i4x = sx, i4y = sy;
    int prev_s = s ^ 4;
    int ptx = sx, pty = sy;
    for (;;) {
        s_end = s;
        while (s < 15) {
            ++s;
            i4x = i3x + dir_dx(s & 7);
            i4y = i3y + dir_dy(s & 7);
            if (img.at(i4x, i4y) != 0) break;
        }
        s &= 7;
        if (s != prev_s) {
            if (keys && n < cap) keys[n] = pack_key(ptx, pty, n);
            ++n;
            prev_s = s;
        }
        ptx += dir_dx(s);
        pty += dir_dy(s);
        if (i4x == sx && i4y == sy && i3x == i1x && i3y == i1y) break;
        i3x = i4x;
        i3y = i4y;
        s = (s + 4) & 7;
    }
    return n;
}

// in-place heapsort of 64-bit keys (ascending): total order (x, y, index) == OpenCV's CHullCmpPoints
B200_HDN void sort_keys(uint64_t* a, int n) {
    for (int start = n / 2 - 1; start >= 0; --start) {
        int root = start;
        uint64_t v = a[root];
        for (;;) {
            int child = 2 * root + 1;
            if (child >= n) break;
            if (child + 1 < n && a[child] < a[child + 1]) ++child;
            if (v >= a[child]) break;
            a[root] = a[child];
            root = child;
        }
        a[root] = v;
    }
    for (int end = n - 1; end > 0; --end) {
        uint64_t v = a[end];
        a[end] = a[0];
        int root = 0;
        for (;;) {
            int child = 2 * root + 1;
            if (child >= end) break;
            if (child + 1 < end && a[child] < a[child + 1]) ++child;
            if (v >= a[child]) break;
            a[root] = a[child];
            root = child;
        }
        a[root] = v;
    }
}

B200_HD int sgn64(int64_t v) { return (v > 0) - (v < 0); }

// OpenCV convhull.cpp Sklansky_ on the sorted key array (positions index `k`).
B200_HDN int sklansky(const uint64_t* k, int start, int end, int* stack, int nsign, int sign2) {
    int incr = end > start ? 1 : -1;
    int pprev = start, pcur = pprev + incr, pnext = pcur + incr;
    int stacksize = 3;
    if (start == end || (key_x(k[start]) == key_x(k[end]) && key_y(k[start]) == key_y(k[end]))) {
        stack[0] = start;
        return 1;
    }
    stack[0] = pprev;
    stack[1] = pcur;
    stack[2] = pnext;
    end += incr;
    while (pnext != end) {
        int cury = key_y(k[pcur]);
        int nexty = key_y(k[pnext]);
        int by = nexty - cury;
        if (((by > 0) - (by < 0)) != nsign) {
            int ax = key_x(k[pcur]) - key_x(k[pprev]);
            int bx = key_x(k[pnext]) - key_x(k[pcur]);
            int ay = cury - key_y(k[pprev]);
            int64_t convexity = (int64_t)ay * bx - (int64_t)ax * by;
            if (sgn64(convexity) == sign2 && (ax != 0 || ay != 0)) {
                pprev = pcur;
                pcur = pnext;
                pnext += incr;
                stack[stacksize] = pnext;
                stacksize++;
            } else {
                if (pprev == start) {
                    pcur = pnext;
                    stack[1] = pcur;
                    pnext += incr;
                    stack[2] = pnext;
                } else {
                    stack[stacksize - 2] = pnext;
                    pcur = pprev;
                    pprev = stack[stacksize - 4];
                    stacksize--;
                }
            }
        } else {
            pnext += incr;
            stack[stacksize - 1] = pnext;
        }
    }
    return --stacksize;
}

// The tail of cv::convexHull: cyclic shift of the hull so that the original contour indices ascend / descend.
B200_HDN int hull_index_shift(const uint64_t* keys, int* hull, int nout, int* stack) {
    // cyclic shift so that the original indices ascend / descend
    if (nout >= 3) {
        int min_idx = 0, max_idx = 0, lt = 0;
        int i;
        for (i = 1; i < nout; ++i) {
            int idx = key_i(keys[hull[i]]);
            lt += key_i(keys[hull[i - 1]]) < idx;
            if (lt > 1 && lt <= i - 2) break;
            if (idx < key_i(keys[hull[min_idx]])) min_idx = i;
            if (idx > key_i(keys[hull[max_idx]])) max_idx = i;
        }
        int mmdist = max_idx - min_idx;
        if (mmdist < 0) mmdist = -mmdist;
        if ((mmdist == 1 || mmdist == nout - 1) && (lt <= 1 || lt >= nout - 2)) {
            int ascending = (max_idx + 1) % nout == min_idx;
            int i0 = ascending ? min_idx : max_idx, j = i0;
            if (i0 > 0) {
                for (i = 0; i < nout; ++i) {
                    int cur = hull[j];
                    stack[i] = cur;
                    int next_j = j + 1 < nout ? j + 1 : 0;
                    int curr_idx = key_i(keys[cur]);
                    int next_idx = key_i(keys[hull[next_j]]);
                    if (i < nout - 1 && (ascending != (curr_idx < next_idx))) break;
                    j = next_j;
                }
                if (i == nout)
                    for (i = 0; i < nout; ++i) hull[i] = stack[i];
            }
        }
    }
    return nout;
}

// cv::convexHull(points, hull, clockwise, returnPoints=true) for int points (cv::minAreaRect passes clockwise=false).
//   keys   : n contour points packed by pack_key, in contour order on entry (sorted in place)
//   stack  : scratch, n + 2 ints
//   hull   : out, positions into the sorted key array (<= n entries); returns the hull size
B200_HDN int convex_hull(uint64_t* keys, int n, int* stack, int* hull, bool clockwise = false, bool presorted = false) {
    if (n == 0) return 0;
    if (!presorted) sort_keys(keys, n);
    int miny = 0, maxy = 0;
    for (int i = 1; i < n; ++i) {
        int y = key_y(keys[i]);
        if (key_y(keys[miny]) > y) miny = i;
        if (key_y(keys[maxy]) < y) maxy = i;
    }
    int nout = 0;
    if (key_x(keys[0]) == key_x(keys[n - 1]) && key_y(keys[0]) == key_y(keys[n - 1])) {
        hull[nout++] = 0;
        return nout;
    }
    // upper half
    int* tl_stack = stack;
    int tl_count = sklansky(keys, 0, maxy, tl_stack, -1, 1);
    int* tr_stack = stack + tl_count;
    int tr_count = sklansky(keys, n - 1, maxy, tr_stack, -1, -1);
    if (!clockwise) {
        int* t = tl_stack; tl_stack = tr_stack; tr_stack = t;
        int c = tl_count; tl_count = tr_count; tr_count = c;
    }
    for (int i = 0; i < tl_count - 1; ++i) hull[nout++] = tl_stack[i];
    for (int i = tr_count - 1; i > 0; --i) hull[nout++] = tr_stack[i];
    int stop_idx = tr_count > 2 ? tr_stack[1] : tl_count > 2 ? tl_stack[tl_count - 2] : -1;
    // lower half (the stack area is reused, exactly like OpenCV does)
    int* bl_stack = stack;
    int bl_count = sklansky(keys, 0, miny, bl_stack, 1, -1);
    int* br_stack = stack + bl_count;
    int br_count = sklansky(keys, n - 1, miny, br_stack, 1, 1);
    if (clockwise) {
        int* t = bl_stack; bl_stack = br_stack; br_stack = t;
        int c = bl_count; bl_count = br_count; br_count = c;
    }
    if (stop_idx >= 0) {
        int check_idx = bl_count > 2 ? bl_stack[1] : bl_count + br_count > 2 ? br_stack[2 - bl_count] : -1;
        if (check_idx == stop_idx ||
            (check_idx >= 0 && key_x(keys[check_idx]) == key_x(keys[stop_idx]) && key_y(keys[check_idx]) == key_y(keys[stop_idx]))) {
            bl_count = bl_count < 2 ? bl_count : 2;
            br_count = br_count < 2 ? br_count : 2;
        }
    }
    for (int i = 0; i < bl_count - 1; ++i) hull[nout++] = bl_stack[i];
    for (int i = br_count - 1; i > 0; --i) hull[nout++] = br_stack[i];
    return hull_index_shift(keys, hull, nout, stack);
}

// ---------------------------------------------------------------------------------------------
// rotating calipers (OpenCV rotcalipers.cpp, CALIPERS_MINAREARECT) on float hull points
// ---------------------------------------------------------------------------------------------
struct RotRect {
    float cx, cy, w, h, angle;
};

// pts: hull points (float), n > 2.  vect: scratch n PtF, inv_len: scratch n floats.  out[6].
B200_HDN void rotating_calipers(const PtF* points, int n, PtF* vect, float* inv_vect_length, float* out, float* best_edge = nullptr) {
    float minarea = 3.402823466e+38f;
    int left = 0, bottom = 0, right = 0, top = 0;
    int seq[4] = {-1, -1, -1, -1};
    float orientation = 0;
    float base_a, base_b = 0;
    float left_x, right_x, top_y, bottom_y;
    PtF pt0 = points[0];
    left_x = right_x = pt0.x;
    top_y = bottom_y = pt0.y;
    for (int i = 0; i < n; ++i) {
        if (pt0.x < left_x) left_x = pt0.x, left = i;
        if (pt0.x > right_x) right_x = pt0.x, right = i;
        if (pt0.y > top_y) top_y = pt0.y, top = i;
        if (pt0.y < bottom_y) bottom_y = pt0.y, bottom = i;
        PtF pt = points[(i + 1) < n ? (i + 1) : 0];
        double dx = pt.x - pt0.x;
        double dy = pt.y - pt0.y;
        vect[i].x = (float)dx;
        vect[i].y = (float)dy;
        inv_vect_length[i] = (float)(1. / sqrt(dx * dx + dy * dy));
        pt0 = pt;
    }
    {
        double ax = vect[n - 1].x, ay = vect[n - 1].y;
        for (int i = 0; i < n; ++i) {
            double bx = vect[i].x, by = vect[i].y;
            double convexity = ax * by - ay * bx;
            if (convexity != 0) {
                orientation = (convexity > 0) ? 1.f : (-1.f);
                break;
            }
            ax = bx;
            ay = by;
        }
    }
    base_a = orientation;
    seq[0] = bottom;
    seq[1] = right;
    seq[2] = top;
    seq[3] = left;
    // best-so-far record (OpenCV keeps it in a char buffer)
    int best_left = 0, best_bottom = 0;
    float best_a = 0, best_w = 0, best_b = 0, best_h = 0;
    float best_ex = 0, best_ey = 0, cur_ex = 1, cur_ey = 0;
    for (int k = 0; k < n; ++k) {
        // choose the calipers side with the minimal angle to its polygon edge by dot-product sign
        // (OpenCV >= 4.5.2: rotate90CW / rotate180 / rotate90CCW + firstVecIsRight)
        PtF rot_vect[4];
        rot_vect[0] = vect[seq[0]];
        rot_vect[1].x = vect[seq[1]].y;  rot_vect[1].y = -vect[seq[1]].x;
        rot_vect[2].x = -vect[seq[2]].x; rot_vect[2].y = -vect[seq[2]].y;
        rot_vect[3].x = -vect[seq[3]].y; rot_vect[3].y = vect[seq[3]].x;
        int main_element = 0;
        for (int i = 1; i < 4; ++i) {
            // firstVecIsRight(rot_vect[i], rot_vect[main_element])
            float tx = rot_vect[i].y, ty = -rot_vect[i].x;
            if (tx * rot_vect[main_element].x + ty * rot_vect[main_element].y < 0) main_element = i;
        }
        {
            int pindex = seq[main_element];
            float lead_x = vect[pindex].x * inv_vect_length[pindex];
            float lead_y = vect[pindex].y * inv_vect_length[pindex];
            switch (main_element) {
                case 0: base_a = lead_x; base_b = lead_y; cur_ex = vect[pindex].x; cur_ey = vect[pindex].y; break;
                case 1: base_a = lead_y; base_b = -lead_x; cur_ex = vect[pindex].y; cur_ey = -vect[pindex].x; break;
                case 2: base_a = -lead_x; base_b = -lead_y; cur_ex = -vect[pindex].x; cur_ey = -vect[pindex].y; break;
                default: base_a = -lead_y; base_b = lead_x; cur_ex = -vect[pindex].y; cur_ey = vect[pindex].x; break;
            }
        }
        seq[main_element] += 1;
        seq[main_element] = (seq[main_element] == n) ? 0 : seq[main_element];
        {
            float dx = points[seq[1]].x - points[seq[3]].x;
            float dy = points[seq[1]].y - points[seq[3]].y;
            float width = dx * base_a + dy * base_b;
            dx = points[seq[2]].x - points[seq[0]].x;
            dy = points[seq[2]].y - points[seq[0]].y;
            float height = -dx * base_b + dy * base_a;
            float area = width * height;
            if (area <= minarea) {
                minarea = area;
                best_left = seq[3];
                best_a = base_a;
                best_w = width;
                best_b = base_b;
                best_h = height;
                best_bottom = seq[0];
                best_ex = cur_ex;
                best_ey = cur_ey;
            }
        }
    }
    float A1 = best_a, B1 = best_b;
    float A2 = -best_b, B2 = best_a;
    float C1 = A1 * points[best_left].x + points[best_left].y * B1;
    float C2 = A2 * points[best_bottom].x + points[best_bottom].y * B2;
    float idet = 1.f / (A1 * B2 - A2 * B1);
    float px = (C1 * B2 - C2 * B1) * idet;
    float py = (A1 * C2 - A2 * C1) * idet;
    out[0] = px;
    out[1] = py;
    out[2] = A1 * best_w;
    out[3] = B1 * best_w;
    out[4] = A2 * best_h;
    out[5] = B2 * best_h;
    if (best_edge) {
        best_edge[0] = best_ex;
        best_edge[1] = best_ey;
        best_edge[2] = best_a;
        best_edge[3] = best_b;
        best_edge[4] = best_w;
        best_edge[5] = best_h;
    }
}

// cv::minAreaRect on hull points (already float, counter-clockwise hull as cv::convexHull(clockwise=false) gives).
// The angle is produced in double precision and folded into [-90, 0) in 90-degree steps, swapping width and
// height on odd steps; this normalisation is pinned empirically against cv2 4.13 (tests/test_geom_host.py:
// 0 mismatching rectangles over 2e4 random point sets) since OpenCV's source is not available in this image.
B200_HDN RotRect min_area_rect(const PtF* hp, int n, PtF* vect, float* inv_len) {
    RotRect box = {0, 0, 0, 0, 0};
    double ang = 0;
    if (n > 2) {
        float out[6];
        rotating_calipers(hp, n, vect, inv_len, out);
        box.cx = out[0] + (out[2] + out[4]) * 0.5f;
        box.cy = out[1] + (out[3] + out[5]) * 0.5f;
        box.w = (float)sqrt((double)out[2] * out[2] + (double)out[3] * out[3]);
        box.h = (float)sqrt((double)out[4] * out[4] + (double)out[5] * out[5]);
        ang = atan2((double)out[3], (double)out[2]);
    } else if (n == 2) {
        box.cx = (hp[0].x + hp[1].x) * 0.5f;
        box.cy = (hp[0].y + hp[1].y) * 0.5f;
        double dx = hp[1].x - hp[0].x;
        double dy = hp[1].y - hp[0].y;
        box.w = (float)sqrt(dx * dx + dy * dy);
        box.h = 0;
        ang = atan2(dy, dx);
    } else if (n == 1) {
        box.cx = hp[0].x;
        box.cy = hp[0].y;
    }
    double deg = ang * 180 / 3.1415926535897932384626433832795;
    int k = (int)floor(deg / 90.0) + 1;
    box.angle = (float)(deg - 90.0 * k);
    if (k & 1) {
        float t = box.w;
        box.w = box.h;
        box.h = t;
    }
    return box;
}

// cv::boxPoints / RotatedRect::points
B200_HDN void box_points(const RotRect& r, PtF pt[4]) {
    double _angle = r.angle * 3.1415926535897932384626433832795 / 180.;
    float b = (float)cos(_angle) * 0.5f;
    float a = (float)sin(_angle) * 0.5f;
    pt[0].x = r.cx - a * r.h - b * r.w;
    pt[0].y = r.cy + b * r.h - a * r.w;
    pt[1].x = r.cx + a * r.h - b * r.w;
    pt[1].y = r.cy - b * r.h - a * r.w;
    pt[2].x = 2 * r.cx - pt[0].x;
    pt[2].y = 2 * r.cy - pt[0].y;
    pt[3].x = 2 * r.cx - pt[1].x;
    pt[3].y = 2 * r.cy - pt[1].y;
}

// DBPostProcess.get_mini_boxes (db_postprocess.py:159-180): stable sort by x, then the index dance.
B200_HDN float mini_box(const RotRect& r, PtF box[4]) {
    PtF p[4];
    box_points(r, p);
    // insertion sort == stable
    for (int i = 1; i < 4; ++i) {
        PtF v = p[i];
        int j = i - 1;
        while (j >= 0 && p[j].x > v.x) {
            p[j + 1] = p[j];
            --j;
        }
        p[j + 1] = v;
    }
    int i1, i2, i3, i4;
    if (p[1].y > p[0].y) { i1 = 0; i4 = 1; } else { i1 = 1; i4 = 0; }
    if (p[3].y > p[2].y) { i2 = 2; i3 = 3; } else { i2 = 3; i3 = 2; }
    box[0] = p[i1];
    box[1] = p[i2];
    box[2] = p[i3];
    box[3] = p[i4];
    return r.w < r.h ? r.w : r.h;
}

// Full cv2.minAreaRect(contour) from packed contour keys (sorted in place).
//   scratch ints:  stack[n+2], hull[n];  scratch floats: hp[n] (PtF), vect[n] (PtF), inv_len[n]
B200_HDN RotRect min_area_rect_keys(uint64_t* keys, int n, int* stack, int* hull, PtF* hp, PtF* vect, float* inv_len) {
    int nh = convex_hull(keys, n, stack, hull);
    for (int i = 0; i < nh; ++i) {
        hp[i].x = (float)key_x(keys[hull[i]]);
        hp[i].y = (float)key_y(keys[hull[i]]);
    }
    return min_area_rect(hp, nh, vect, inv_len);
}

// ---------------------------------------------------------------------------------------------
// box_score_fast mask (db_postprocess.py:182-197): cv2.fillPoly of an integer quad
// ---------------------------------------------------------------------------------------------
// cv2.line(LINE_8) membership of pixel (x, y) for the segment a -> b, in closed form.
B200_HD bool on_line8(int x, int y, int ax, int ay, int bx, int by) {
    if (bx < ax) {
        int t = ax; ax = bx; bx = t;
        t = ay; ay = by; by = t;
    }
    int dx = bx - ax, dy = by - ay;
    int sy = (dy > 0) - (dy < 0);
    int ady = dy < 0 ? -dy : dy;
    if (ady > dx) {                 // y-major
        int j = (y - ay) * sy;      // step index along the major axis
        if (j < 0 || j > ady) return false;
        long long m = ((long long)2 * dx * j + ady - 1) / ((long long)2 * ady);
        if (j == 0) m = 0;
        return x == ax + (int)m;
    } else {                        // x-major (also a single point)
        int j = x - ax;
        if (j < 0 || j > dx) return false;
        if (dx == 0) return y == ay;
        long long m = ((long long)2 * ady * j + dx - 1) / ((long long)2 * dx);
        if (j == 0) m = 0;
        return y == ay + sy * (int)m;
    }
}

// cv::clipLine(Size, Point2l&, Point2l&) — literal port (64-bit integers, truncating double arithmetic).
B200_HDN bool clip_line(long long width, long long height, long long& x1, long long& y1, long long& x2, long long& y2) {
    int c1, c2;
    long long right = width - 1, bottom = height - 1;
    if (width <= 0 || height <= 0) return false;
    c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        long long a;
        if (c1 & 12) {
            a = c1 < 8 ? 0 : bottom;
            x1 += (long long)((double)(a - y1) * (x2 - x1) / (y2 - y1));
            y1 = a;
            c1 = (x1 < 0) + (x1 > right) * 2;
        }
        if (c2 & 12) {
            a = c2 < 8 ? 0 : bottom;
            x2 += (long long)((double)(a - y2) * (x2 - x1) / (y2 - y1));
            y2 = a;
            c2 = (x2 < 0) + (x2 > right) * 2;
        }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) {
                a = c1 == 1 ? 0 : right;
                y1 += (long long)((double)(a - x1) * (y2 - y1) / (x2 - x1));
                x1 = a;
                c1 = 0;
            }
            if (c2) {
                a = c2 == 1 ? 0 : right;
                y2 += (long long)((double)(a - x2) * (y2 - y1) / (x2 - x1));
                x2 = a;
                c2 = 0;
            }
        }
    }
    return (c1 | c2) == 0;
}

// cv2.fillPoly(mask, quad, 1) for one integer quad on a mw x mh mask, restated from OpenCV's drawing.cpp
// (CollectPolyEdges + FillEdgeCollection + Line/LineIterator with clipping): a pixel is set when it lies on a
// (clipped) LINE_8 boundary line or inside a scan-line span [ceil(x_left), floor(x_right)] between two active
// edges (16.16 fixed point, dx truncated).
struct FillPoly4 {
    // boundary lines after clipping
    int lax[4], lay[4], lbx[4], lby[4];
    bool lvis[4];
    // polygon edges
    int ne;
    int ey0[4], ey1[4];
    long long ex[4], edx[4];
    int mw, mh;
    bool fill_ok;
    int row_y;
    int nspan;
    int sp0[2], sp1[2];

    B200_HDN void init(const int* qx, const int* qy, int mask_w, int mask_h) {
        mw = mask_w;
        mh = mask_h;
        ne = 0;
        row_y = -2147483647;
        nspan = 0;
        long long p0x = (long long)qx[3] << 16, p0y = qy[3];
        for (int i = 0; i < 4; ++i) {
            long long p1x = (long long)qx[i] << 16, p1y = qy[i];
            long long p0cx = p0x, p0cy = p0y, p1cx = p1x, p1cy = p1y;
            long long t0x = p0x >> 16, t0y = p0y, t1x = p1x >> 16, t1y = p1y;
            {   // Line(): LineIterator clips a copy of the endpoints
                long long ax = t0x, ay = t0y, bx = t1x, by = t1y;
                lvis[i] = clip_line(mw, mh, ax, ay, bx, by);
                lax[i] = (int)ax; lay[i] = (int)ay; lbx[i] = (int)bx; lby[i] = (int)by;
            }
            if ((unsigned long long)t0x >= (unsigned long long)mw || (unsigned long long)t1x >= (unsigned long long)mw ||
                (unsigned long long)t0y >= (unsigned long long)mh || (unsigned long long)t1y >= (unsigned long long)mh) {
                // edges that leave the mask follow the clipped boundary line (pinned empirically against
                // cv2 4.13: 0 mismatching masks over 4000 random quads, tests/test_geom_host.py)
                clip_line(mw, mh, t0x, t0y, t1x, t1y);
                if (t0y != t1y) {
                    p0cy = t0y;
                    p1cy = t1y;
                }
                p0cx = t0x << 16;
                p1cx = t1x << 16;
            }
            if (p0y != p1y) {
                long long dx = (p1cx - p0cx) / (p1cy - p0cy);
                edx[ne] = dx;
                if (p0y < p1y) {
                    ey0[ne] = (int)p0y;
                    ey1[ne] = (int)p1y;
                    ex[ne] = p0cx + (p0y - p0cy) * dx;
                } else {
                    ey0[ne] = (int)p1y;
                    ey1[ne] = (int)p0y;
                    ex[ne] = p1cx + (p1y - p1cy) * dx;
                }
                ++ne;
            }
            p0x = p1x;
            p0y = p1y;
        }
        fill_ok = ne >= 2;
        if (fill_ok) {
            int y_max = -2147483647 - 1, y_min = 2147483647;
            long long x_max = -1, x_min = 0x7FFFFFFFFFFFFFFFLL;    // OpenCV initialises x_max to 0xFFFF...F == -1
            for (int i = 0; i < ne; ++i) {
                long long x1 = ex[i] + (long long)(ey1[i] - ey0[i]) * edx[i];
                y_min = ey0[i] < y_min ? ey0[i] : y_min;
                y_max = ey1[i] > y_max ? ey1[i] : y_max;
                x_min = ex[i] < x_min ? ex[i] : x_min;
                x_max = ex[i] > x_max ? ex[i] : x_max;
                x_min = x1 < x_min ? x1 : x_min;
                x_max = x1 > x_max ? x1 : x_max;
            }
            if (y_max < 0 || y_min >= mh || x_max < 0 || x_min >= ((long long)mw << 16)) fill_ok = false;
        }
    }

    // scan-line spans of row y
    B200_HDN void set_row(int y) {
        row_y = y;
        nspan = 0;
        if (!fill_ok || y < 0 || y >= mh) return;
        long long xs[4];
        int na = 0;
        for (int i = 0; i < ne; ++i)
            if (ey0[i] <= y && y < ey1[i]) xs[na++] = ex[i] + (long long)(y - ey0[i]) * edx[i];
        for (int i = 1; i < na; ++i) {
            long long v = xs[i];
            int j = i - 1;
            while (j >= 0 && xs[j] > v) { xs[j + 1] = xs[j]; --j; }
            xs[j + 1] = v;
        }
        for (int i = 0; i + 1 < na; i += 2) {
            int x1 = (int)((xs[i] + 65535) >> 16), x2 = (int)(xs[i + 1] >> 16);   // ceil(left) .. floor(right)
            if (x1 < mw && x2 >= 0) {
                if (x1 < 0) x1 = 0;
                if (x2 >= mw) x2 = mw - 1;
                sp0[nspan] = x1;
                sp1[nspan] = x2;
                ++nspan;
            }
        }
    }

    B200_HDN bool test(int x, int y) {
        if (y != row_y) set_row(y);
        for (int i = 0; i < nspan; ++i)
            if (x >= sp0[i] && x <= sp1[i]) return true;
        for (int i = 0; i < 4; ++i)
            if (lvis[i] && on_line8(x, y, lax[i], lay[i], lbx[i], lby[i])) return true;
        return false;
    }
};

// box_score_fast: mean of pred over the fillPoly mask of the (shifted, truncated) box.
B200_HDN double box_score_fast(const float* pred, int h, int w, const PtF box[4]) {
    float fxmin = box[0].x, fxmax = box[0].x, fymin = box[0].y, fymax = box[0].y;
    for (int i = 1; i < 4; ++i) {
        fxmin = box[i].x < fxmin ? box[i].x : fxmin;
        fxmax = box[i].x > fxmax ? box[i].x : fxmax;
        fymin = box[i].y < fymin ? box[i].y : fymin;
        fymax = box[i].y > fymax ? box[i].y : fymax;
    }
    auto clampi = [](int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); };
    int xmin = clampi((int)floorf(fxmin), 0, w - 1);
    int xmax = clampi((int)ceilf(fxmax), 0, w - 1);
    int ymin = clampi((int)floorf(fymin), 0, h - 1);
    int ymax = clampi((int)ceilf(fymax), 0, h - 1);
    int qx[4], qy[4];
    for (int i = 0; i < 4; ++i) {
        qx[i] = (int)(box[i].x - (float)xmin);      // float32 subtract, then truncation toward zero
        qy[i] = (int)(box[i].y - (float)ymin);
    }
    FillPoly4 fp;
    fp.init(qx, qy, xmax - xmin + 1, ymax - ymin + 1);
    double sum = 0;
    long long cnt = 0;
    for (int y = ymin; y <= ymax; ++y) {
        fp.set_row(y - ymin);
        for (int x = xmin; x <= xmax; ++x)
            if (fp.test(x - xmin, y - ymin)) {
                sum += (double)pred[(size_t)y * w + x];
                ++cnt;
            }
    }
    return cnt ? sum / (double)cnt : 0.0;
}

// ---------------------------------------------------------------------------------------------
// box_score_slow (db_postprocess.py:199-218, det_db_score_mode='slow'): cv2.fillPoly of the CONTOUR itself + cv2.mean.
// Every edge of a CHAIN_APPROX_SIMPLE contour is horizontal, vertical or diagonal (a run of one chain direction), so on
// an integer row an edge is either a pixel range (horizontal) or exactly one pixel, and cv2's scan conversion reduces to:
// mask = pixels on the closed chain  ∪  pixels with an odd number of edge crossings strictly to their right, an edge
// counting on the rows [min y, max y) (even-odd rule; tests/test_geom_host.py: 0 mismatching pixels against cv2.fillPoly
// over outer and hole contours of blob, blurred-noise and speckle bitmaps).
// ---------------------------------------------------------------------------------------------
// What edge (x0, y0) -> (x1, y1) contributes to row y: bit 0 = chain pixels [*xa, *xb], bit 1 = a crossing at *xa.
B200_HD int edge_on_row(int x0, int y0, int x1, int y1, int y, int* xa, int* xb) {
    if (y0 == y1) {
        if (y != y0) return 0;
        *xa = x0 < x1 ? x0 : x1;
        *xb = x0 < x1 ? x1 : x0;
        return 1;
    }
    const int lo = y0 < y1 ? y0 : y1, hi = y0 < y1 ? y1 : y0;
    if (y < lo || y > hi) return 0;
    const int x = x0 + (y - y0) * (x1 - x0) / (y1 - y0);      // exact: |x1 - x0| is 0 or |y1 - y0|
    *xa = *xb = x;
    return y < hi ? 3 : 1;
}

// p[b] = t[b] ^ t[b+1] ^ ... ^ t[31]
B200_HD uint32_t suffix_xor(uint32_t t) {
    t ^= t >> 1; t ^= t >> 2; t ^= t >> 4; t ^= t >> 8; t ^= t >> 16;
    return t;
}

// Serial form (host fuzzing, and the specification of score_slow_kernel in dbpost.cu).  keys[0..nv) are the contour
// points in trace order; chain / tog are scratch rows of ceil(bbox width / 32) words each.
B200_HDN double box_score_slow(const float* pred, int h, int w, const uint64_t* keys, int nv, uint32_t* chain, uint32_t* tog) {
    (void)h;
    int xmin = key_x(keys[0]), xmax = xmin, ymin = key_y(keys[0]), ymax = ymin;
    for (int i = 1; i < nv; ++i) {
        const int x = key_x(keys[i]), y = key_y(keys[i]);
        xmin = x < xmin ? x : xmin; xmax = x > xmax ? x : xmax;
        ymin = y < ymin ? y : ymin; ymax = y > ymax ? y : ymax;
    }
    const int nw = (xmax - xmin) / 32 + 1;
    double sum = 0;
    long long cnt = 0;
    for (int y = ymin; y <= ymax; ++y) {
        for (int i = 0; i < nw; ++i) chain[i] = tog[i] = 0u;
        for (int i = 0; i < nv; ++i) {
            const uint64_t k0 = keys[i], k1 = keys[i + 1 == nv ? 0 : i + 1];
            int xa, xb;
            const int f = edge_on_row(key_x(k0), key_y(k0), key_x(k1), key_y(k1), y, &xa, &xb);
            if (f & 1)
                for (int x = xa - xmin; x <= xb - xmin; ++x) chain[x >> 5] |= 1u << (x & 31);
            if (f & 2) tog[(xa - xmin) >> 5] ^= 1u << ((xa - xmin) & 31);
        }
        uint32_t carry = 0;
        for (int wi = nw - 1; wi >= 0; --wi) {
            const uint32_t t = tog[wi];
            uint32_t m = suffix_xor(t) ^ t;                   // crossings strictly to the right, inside this word
            if (carry) m = ~m;
            uint32_t pc = t;
            pc ^= pc >> 16; pc ^= pc >> 8; pc ^= pc >> 4; pc ^= pc >> 2; pc ^= pc >> 1;
            carry ^= pc & 1u;
            m |= chain[wi];
            for (int b = 0; b < 32; ++b)
                if ((m >> b) & 1u) {
                    sum += (double)pred[(size_t)y * w + xmin + wi * 32 + b];
                    ++cnt;
                }
        }
    }
    return cnt ? sum / (double)cnt : 0.0;
}

// ---------------------------------------------------------------------------------------------
// unclip (db_postprocess.py:151-157): shapely area/length + ClipperOffset(JT_ROUND, ET_CLOSEDPOLYGON)
// ---------------------------------------------------------------------------------------------
B200_HD long long clipper_round(double v) { return v < 0 ? (long long)(v - 0.5) : (long long)(v + 0.5); }

// Returns the number of output points written to ox/oy (cap >= 4 * (steps + 2)); 0 when the path is degenerate.
B200_HDN int unclip_offset(const PtF box[4], double unclip_ratio, int* ox, int* oy, int cap, double* distance_out) {
    // shapely: Polygon(box).area (shoelace, absolute) and .length (ring perimeter), float64
    double area2 = 0, length = 0;
    for (int i = 0; i < 4; ++i) {
        int j = (i + 1) & 3;
        double xi = box[i].x, yi = box[i].y, xj = box[j].x, yj = box[j].y;
        area2 += xi * yj - xj * yi;
        double dx = xj - xi, dy = yj - yi;
        length += sqrt(dx * dx + dy * dy);
    }
    double area = fabs(area2) * 0.5;
    double delta = area * unclip_ratio / length;
    if (distance_out) *distance_out = delta;
    // pyclipper: coordinates truncated to integers on entry
    long long px[4], py[4];
    int len = 0;
    {
        long long tx[4], ty[4];
        for (int i = 0; i < 4; ++i) {
            tx[i] = (long long)box[i].x;
            ty[i] = (long long)box[i].y;
        }
        int high = 3;
        while (high > 0 && tx[0] == tx[high] && ty[0] == ty[high]) high--;
        px[0] = tx[0];
        py[0] = ty[0];
        len = 1;
        for (int i = 1; i <= high; ++i)
            if (px[len - 1] != tx[i] || py[len - 1] != ty[i]) {
                px[len] = tx[i];
                py[len] = ty[i];
                ++len;
            }
        if (len < 3) return 0;
    }
    // FixOrientations: reverse when the signed area is negative
    {
        double a = 0;
        for (int i = 0, j = len - 1; i < len; ++i) {
            a += ((double)px[j] + (double)px[i]) * ((double)py[j] - (double)py[i]);
            j = i;
        }
        if (!(-a * 0.5 >= 0)) {
            for (int i = 0; i < len / 2; ++i) {
                long long t = px[i]; px[i] = px[len - 1 - i]; px[len - 1 - i] = t;
                t = py[i]; py[i] = py[len - 1 - i]; py[len - 1 - i] = t;
            }
        }
    }
    if (fabs(delta) < 1.0E-20) {                 // NEAR_ZERO: closed polygons are copied
        int n = 0;
        for (int i = 0; i < len && n < cap; ++i) { ox[n] = (int)px[i]; oy[n] = (int)py[i]; ++n; }
        return n;
    }
    const double pi = 3.141592653589793238;
    const double two_pi = pi * 2;
    const double def_arc_tolerance = 0.25;
    double y;
    double arc_tolerance = 0.25;                 // pyclipper default ArcTolerance
    if (arc_tolerance > fabs(delta) * def_arc_tolerance) y = fabs(delta) * def_arc_tolerance;
    else y = arc_tolerance;
    double steps = pi / acos(1 - y / fabs(delta));
    if (steps > fabs(delta) * pi) steps = fabs(delta) * pi;
    double m_sin = sin(two_pi / steps);
    double m_cos = cos(two_pi / steps);
    double steps_per_rad = steps / two_pi;
    if (delta < 0.0) m_sin = -m_sin;
    double nx[4], ny[4];
    for (int j = 0; j < len; ++j) {
        int j2 = (j + 1) % len;
        if (px[j2] == px[j] && py[j2] == py[j]) { nx[j] = 0; ny[j] = 0; continue; }
        double Dx = (double)(px[j2] - px[j]);
        double dy = (double)(py[j2] - py[j]);
        double f = 1 * 1.0 / sqrt(Dx * Dx + dy * dy);
        Dx *= f;
        dy *= f;
        nx[j] = dy;
        ny[j] = -Dx;
    }
    int n = 0;
    auto push = [&](long long x, long long yy) {
        if (n < cap) { ox[n] = (int)x; oy[n] = (int)yy; }
        ++n;
    };
    int k = len - 1;
    for (int j = 0; j < len; ++j) {
        double sinA = nx[k] * ny[j] - nx[j] * ny[k];
        bool done = false;
        if (fabs(sinA * delta) < 1.0) {
            double cosA = nx[k] * nx[j] + ny[j] * ny[k];
            if (cosA > 0) {
                push(clipper_round((double)px[j] + nx[k] * delta), clipper_round((double)py[j] + ny[k] * delta));
                done = true;
            }
        } else if (sinA > 1.0) sinA = 1.0;
        else if (sinA < -1.0) sinA = -1.0;
        if (!done) {
            if (sinA * delta < 0) {
                push(clipper_round((double)px[j] + nx[k] * delta), clipper_round((double)py[j] + ny[k] * delta));
                push(px[j], py[j]);
                push(clipper_round((double)px[j] + nx[j] * delta), clipper_round((double)py[j] + ny[j] * delta));
            } else {            // jtRound
                double a = atan2(sinA, nx[k] * nx[j] + ny[k] * ny[j]);
                long long st = clipper_round(steps_per_rad * fabs(a));
                int isteps = (int)(st > 1 ? st : 1);
                double X = nx[k], Y = ny[k], X2;
                for (int i = 0; i < isteps; ++i) {
                    push(clipper_round((double)px[j] + X * delta), clipper_round((double)py[j] + Y * delta));
                    X2 = X;
                    X = X * m_cos - m_sin * Y;
                    Y = X2 * m_sin + Y * m_cos;
                }
                push(clipper_round((double)px[j] + nx[j] * delta), clipper_round((double)py[j] + ny[j] * delta));
            }
        }
        k = j;
    }
    return n;
}

// np.clip(np.round(v / width * dest), 0, dest) -> float32 -> int32   (db_postprocess.py:143-147, numpy 2 promotion:
// float32 / python int stays float32, times the np.float64 `dest` is float64)
B200_HD int scale_coord(float v, int extent, double dest) {
    float q = v / (float)extent;
    double t = (double)q * dest;
    double r = rint(t);
    r = r < 0 ? 0 : (r > dest ? dest : r);
    return (int)(float)r;
}

}  // namespace geom
}  // namespace b200ocr

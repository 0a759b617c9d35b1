// Host build of the product's geometry header (onnxocr_b200/csrc/dbgeom.h) for CPU fuzzing against cv2.
// TEST INFRASTRUCTURE: compiled by tests/host/build_host.py into tests/host/libb200geom_host.so.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../onnxocr_b200/csrc/dbgeom.h"

using namespace b200ocr::geom;

extern "C" {

int geom_trace(const uint8_t* bitmap, int h, int w, int sx, int sy, int is_hole, int32_t* out_xy, int cap) {
    Bitmap bm{bitmap, h, w};
    std::vector<uint64_t> keys(cap);
    int n = trace_contour(bm, sx, sy, is_hole != 0, keys.data(), cap);
    for (int i = 0; i < n && i < cap; ++i) {
        out_xy[2 * i] = key_x(keys[i]);
        out_xy[2 * i + 1] = key_y(keys[i]);
    }
    return n;
}

// cv2.minAreaRect + boxPoints + get_mini_boxes on integer points
float geom_min_area_rect(const int32_t* xy, int n, float rect[5], float boxpts[8], float mini[8], int32_t* hull_out, int* nh_out, int clockwise) {
    std::vector<uint64_t> keys(n);
    for (int i = 0; i < n; ++i) keys[i] = pack_key(xy[2 * i], xy[2 * i + 1], i);
    std::vector<int> stack(n + 2), hull(n);
    std::vector<PtF> hp(n), vect(n);
    std::vector<float> inv(n);
    int nh = convex_hull(keys.data(), n, stack.data(), hull.data(), clockwise != 0);
    for (int i = 0; i < nh; ++i) {
        hp[i].x = (float)key_x(keys[hull[i]]);
        hp[i].y = (float)key_y(keys[hull[i]]);
        if (hull_out) hull_out[i] = key_i(keys[hull[i]]);
    }
    if (nh_out) *nh_out = nh;
    RotRect r = min_area_rect(hp.data(), nh, vect.data(), inv.data());
    rect[0] = r.cx; rect[1] = r.cy; rect[2] = r.w; rect[3] = r.h; rect[4] = r.angle;
    PtF p[4];
    box_points(r, p);
    for (int i = 0; i < 4; ++i) { boxpts[2 * i] = p[i].x; boxpts[2 * i + 1] = p[i].y; }
    PtF m[4];
    float sside = mini_box(r, m);
    for (int i = 0; i < 4; ++i) { mini[2 * i] = m[i].x; mini[2 * i + 1] = m[i].y; }
    return sside;
}

// raw rotatingCalipers output (6 floats) for hull points of an integer point set
int geom_calipers_raw(const int32_t* xy, int n, float out6[6], int clockwise, float edge[2]) {
    std::vector<uint64_t> keys(n);
    for (int i = 0; i < n; ++i) keys[i] = pack_key(xy[2 * i], xy[2 * i + 1], i);
    std::vector<int> stack(n + 2), hull(n);
    std::vector<PtF> hp(n), vect(n);
    std::vector<float> inv(n);
    int nh = convex_hull(keys.data(), n, stack.data(), hull.data(), clockwise != 0);
    for (int i = 0; i < nh; ++i) { hp[i].x = (float)key_x(keys[hull[i]]); hp[i].y = (float)key_y(keys[hull[i]]); }
    if (nh > 2) rotating_calipers(hp.data(), nh, vect.data(), inv.data(), out6, edge);
    return nh;
}

void geom_fill_mask(const int32_t qx[4], const int32_t qy[4], int h, int w, uint8_t* mask) {
    int x4[4] = {qx[0], qx[1], qx[2], qx[3]}, y4[4] = {qy[0], qy[1], qy[2], qy[3]};
    FillPoly4 fp;
    fp.init(x4, y4, w, h);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) mask[(size_t)y * w + x] = fp.test(x, y) ? 1 : 0;
}

void geom_line_mask(int ax, int ay, int bx, int by, int h, int w, uint8_t* mask) {
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) mask[(size_t)y * w + x] = on_line8(x, y, ax, ay, bx, by) ? 1 : 0;
}

double geom_box_score(const float* pred, int h, int w, const float box[8]) {
    PtF b[4];
    for (int i = 0; i < 4; ++i) { b[i].x = box[2 * i]; b[i].y = box[2 * i + 1]; }
    return box_score_fast(pred, h, w, b);
}

// box_score_slow on the contour that starts at (sx, sy): trace + fillPoly(contour) mean
double geom_box_score_slow(const float* pred, const uint8_t* bitmap, int h, int w, int sx, int sy, int is_hole) {
    Bitmap bm{bitmap, h, w};
    int n = trace_contour(bm, sx, sy, is_hole != 0, nullptr, 0);
    std::vector<uint64_t> keys(n);
    trace_contour(bm, sx, sy, is_hole != 0, keys.data(), n);
    std::vector<uint32_t> chain(w / 32 + 2), tog(w / 32 + 2);
    return box_score_slow(pred, h, w, keys.data(), n, chain.data(), tog.data());
}

int geom_unclip(const float box[8], double ratio, int32_t* out_xy, int cap, double* dist) {
    PtF b[4];
    for (int i = 0; i < 4; ++i) { b[i].x = box[2 * i]; b[i].y = box[2 * i + 1]; }
    std::vector<int> ox(cap), oy(cap);
    int n = unclip_offset(b, ratio, ox.data(), oy.data(), cap, dist);
    for (int i = 0; i < n && i < cap; ++i) { out_xy[2 * i] = ox[i]; out_xy[2 * i + 1] = oy[i]; }
    return n;
}

int geom_scale_coord(float v, int extent, double dest) { return scale_coord(v, extent, dest); }

}  // extern "C"

// Standalone timing harness for the tcgen05 convolution kernels (no Python, no engine): builds the same tensor maps and
// Params as engine.cu for one layer, launches the kernel directly and prints the device time plus the per-role cycle
// counters (Params::dbg).  Data is random: correctness is covered by tests/test_gpu_tc.py, this tool is for timing only.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I onnxocr_b200/csrc -o tools/ubench/conv_bench tools/ubench/conv_bench.cu
//   conv_bench cin cout k n h w mode(0 stream | 1 halo | 2 halo two-tile | 3 weight-stationary per-tap) [iters]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "tc_conv.cuh"
using namespace b200ocr;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn enc() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    return (EncodeTiledFn)p;
}
static void make_tmap(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* str, const uint32_t* box) {
    cuuint64_t gd[4], gs[3];
    cuuint32_t bx[4], es[4];
    for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; if (i) gs[i - 1] = str[i - 1]; }
    CUresult r = enc()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("tensor map failed %d\n", (int)r); exit(1); }
}
static int rup(int a, int b) { return (a + b - 1) / b * b; }

int main(int argc, char** argv) {
    if (argc < 8) { printf("usage: conv_bench cin cout k n h w mode [iters]\n"); return 1; }
    int cin = atoi(argv[1]), cout = atoi(argv[2]), k = atoi(argv[3]), n = atoi(argv[4]), h = atoi(argv[5]), w = atoi(argv[6]);
    int mode = atoi(argv[7]), iters = argc > 8 ? atoi(argv[8]) : 10;
    int cinp = cpad_of(cin), coutp = cpad_of(cout), taps = k * k;
    int kchunks = (cinp + 63) / 64;
    size_t K = (size_t)taps * kchunks * 64;
    int bn0 = coutp <= 32 ? 32 : (coutp <= 64 ? 64 : 128);
    int rows = rup(coutp, 256);
    bf16 *x, *y, *wt;
    float* bias;
    size_t xe = (size_t)n * h * w * cinp, ye = (size_t)n * h * w * coutp;
    cudaMalloc(&x, xe * 2); cudaMalloc(&y, ye * 2); cudaMalloc(&wt, (size_t)rows * K * 2); cudaMalloc(&bias, (rows + 256) * 4);
    cudaMemset(x, 0x3c, xe * 2); cudaMemset(wt, 0x38, (size_t)rows * K * 2); cudaMemset(bias, 0, (rows + 256) * 4);
    unsigned long long* dbg;
    cudaMalloc(&dbg, 148 * 2 * 8 * 8);
    char* flush; cudaMalloc(&flush, 256 << 20);

    tc::Params tp;
    memset(&tp, 0, sizeof(tp));
    tp.mode = k == 1 ? 0 : 1;
    tp.M = n * h * w; tp.N = n; tp.H = h; tp.W = w;
    tp.taps = taps; tp.kw = k; tp.kh = k; tp.pad_t = tp.pad_l = k / 2;
    tp.kchunks = kchunks; tp.cinp = cinp; tp.Cout = cout; tp.CoutP = coutp;
    tp.epi.push(EPI_RELU);
    tp.TW = 128; tp.TH = 1;
    if (tp.mode == 1) {
        long best = -1;
        for (int tw = 128; tw >= 8; tw >>= 1) {
            int th = 128 / tw;
            long cover = (long)((w + tw - 1) / tw) * tw * ((h + th - 1) / th) * th;
            if (best < 0 || cover < best) { best = cover; tp.TW = tw; tp.TH = th; }
        }
    }
    const bool halo = tp.mode == 1 && (mode == 1 || mode == 2);
    size_t stage_bytes = tc::WS_A_BYTES;
    if (halo) { stage_bytes = tc::ws_halo_stage_bytes(k, k); tp.TW = tc::WS_HALO_TW; tp.TH = tc::WS_HALO_TH; }
    tp.tiles_x = (w + tp.TW - 1) / tp.TW; tp.tiles_y = (h + tp.TH - 1) / tp.TH;
    const int nkb = taps * kchunks;
    int ws_bn = (mode != 0) ? tc::ws_pick_bn(coutp, nkb, stage_bytes) : 0;
    int bn = ws_bn ? ws_bn : bn0;
    if (ws_bn) { tp.ws_bn = ws_bn; tp.ws_stages = tc::ws_pick_stages(ws_bn, nkb, stage_bytes); tp.ws_halo = halo ? mode : 0; tp.ws_stage_bytes = (int)stage_bytes; }
    if (getenv("STAGES")) tp.ws_stages = atoi(getenv("STAGES"));
    tp.dbg = dbg;
    if (getenv("EPI")) tp.epi_variant = atoi(getenv("EPI"));
    if (getenv("GAP")) {          // SE producer: channel sums in the epilogue; GAP = channels per pixel (gap_mod)
        unsigned long long* sums;
        cudaMalloc(&sums, (size_t)n * tc::kGapReplicas * 1024 * 8);
        cudaMemset(sums, 0, (size_t)n * tc::kGapReplicas * 1024 * 8);
        tp.gap = sums;
        tp.gap_mod = atoi(getenv("GAP"));
        tp.gap_rows = tp.mode == 0 ? tp.M / n : 0;
    }
    CUtensorMap map_a, map_b;
    { uint64_t dims[2] = {K, (uint64_t)rows}; uint64_t str[1] = {K * 2}; uint32_t box[2] = {64, (uint32_t)bn}; make_tmap(&map_b, wt, 2, dims, str, box); }
    if (tp.mode == 0) {
        uint64_t dims[2] = {(uint64_t)cinp, (uint64_t)tp.M}; uint64_t str[1] = {(uint64_t)cinp * 2}; uint32_t box[2] = {64, 128};
        make_tmap(&map_a, x, 2, dims, str, box);
    } else {
        uint64_t dims[4] = {(uint64_t)cinp, (uint64_t)w, (uint64_t)h, (uint64_t)n};
        uint64_t str[3] = {(uint64_t)cinp * 2, (uint64_t)w * cinp * 2, (uint64_t)h * w * cinp * 2};
        uint32_t box[4] = {64, (uint32_t)tp.TW, (uint32_t)tp.TH, 1};
        if (halo) { box[1] = tp.TW + k - 1; box[2] = tp.TH + k - 1; }
        make_tmap(&map_a, x, 4, dims, str, box);
    }
    int m_tiles = tp.mode == 0 ? (tp.gap ? n * ((tp.gap_rows + 127) / 128) : (tp.M + 127) / 128) : tp.tiles_x * tp.tiles_y * n;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f, sum = 0;
    printf("cin %d cout %d k %d n %d %dx%d mode %d: bn %d stages %d m_tiles %d tile %dx%d\n", cin, cout, k, n, h, w, mode, bn, tp.ws_stages, m_tiles, tp.TH, tp.TW);
    for (int it = 0; it < iters + 2; ++it) {
        cudaMemsetAsync(flush, it, 256 << 20);
        cudaMemsetAsync(dbg, 0, 148 * 2 * 8 * 8);
        cudaEventRecord(e0);
        if (ws_bn) {
            int n_tiles = (coutp + ws_bn - 1) / ws_bn;
            int m_groups = std::max(1, std::min(m_tiles, 148 / n_tiles));
            size_t smem = tc::ws_smem_bytes(ws_bn, nkb, tp.ws_stages, stage_bytes);
            cudaFuncSetAttribute(tc::conv_tc_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::WS_SMEM_LIMIT);
            tc::conv_tc_ws_kernel<<<n_tiles * m_groups, tc::kWsThreads, smem>>>(map_a, map_b, bias, y, tp);
        } else {
            int total = m_tiles * ((coutp + bn - 1) / bn);
            int grid = std::min(total, 148 * (bn == 128 ? 1 : 2));
            if (bn == 32) { cudaFuncSetAttribute(tc::conv_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::smem_bytes<32>()); tc::conv_tc_kernel<32><<<grid, tc::kThreads, tc::smem_bytes<32>()>>>(map_a, map_b, bias, y, tp); }
            else if (bn == 64) { cudaFuncSetAttribute(tc::conv_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::smem_bytes<64>()); tc::conv_tc_kernel<64><<<grid, tc::kThreads, tc::smem_bytes<64>()>>>(map_a, map_b, bias, y, tp); }
            else { cudaFuncSetAttribute(tc::conv_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::smem_bytes<128>()); tc::conv_tc_kernel<128><<<grid, tc::kThreads, tc::smem_bytes<128>()>>>(map_a, map_b, bias, y, tp); }
        }
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it >= 2) { best = std::min(best, ms); sum += ms; }
    }
    double bytes = (double)(xe + ye) * 2 + (double)cout * cin * taps * 2, flops = 2.0 * tp.M * cout * cin * taps;
    printf("  avg %.4f ms best %.4f ms   %.1f GB/s  %.1f TF/s (best)\n", sum / iters, best, bytes / best / 1e6, flops / best / 1e9);
    std::vector<unsigned long long> hd(148 * 8);
    cudaMemcpy(hd.data(), dbg, 148 * 8 * 8, cudaMemcpyDeviceToHost);
    const char* names[8] = {"mma total", "mma wait weights", "mma wait acc", "mma wait stages", "epi(w2) total", "epi wait acc", "tma wait empty", "-"};
    for (int j = 0; j < 7; ++j) {
        double s = 0, mx = 0; int cnt = 0;
        for (int b = 0; b < 148; ++b) { double v = (double)hd[b * 8 + j]; if (hd[b * 8] || hd[b * 8 + 4]) { s += v; mx = std::max(mx, v); ++cnt; } }
        if (cnt) printf("  %-18s avg %9.0f max %9.0f cycles\n", names[j], s / cnt, mx);
    }
    return 0;
}

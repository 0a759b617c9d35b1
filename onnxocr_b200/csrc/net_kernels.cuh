// CUDA-core kernels of the graph executor.  All activations are NHWC with the channel
// count padded to cpad (1, 4 or a multiple of 8); pad lanes always hold zeros.
// T is the storage type (float for the exact fp32 mode, bf16 for the fast mode);
// accumulation is always fp32.
#pragma once
#include <cuda.h>
#ifndef B200_DW_DENSE_S2
#define B200_DW_DENSE_S2 1
#endif

#include "common.cuh"
#include "pre_kernels.cuh"

namespace b200ocr {

// ----------------------------------------------------------------------------------------
// layout conversion at the session seam (reference feeds NCHW f32: predict_det.py:102-107)
// ----------------------------------------------------------------------------------------
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, T* __restrict__ out, int N, int C, int H, int W,
                                    int CP) {
    pdl_start();
    int64_t total = (int64_t)N * H * W * CP;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % CP);
        int64_t p = i / CP;
        int x = (int)(p % W);
        int64_t q = p / W;
        int y = (int)(q % H);
        int n = (int)(q / H);
        float v = c < C ? in[(((int64_t)n * C + c) * H + y) * W + x] : 0.f;
        ST<T>::st(out + i, v);
    }
}

// NHWC(cpad) T -> dense fp32.  to_nchw: output [N,C,H,W]; else [N,H,W,C] (pad stripped).
template <typename T>
__global__ void to_dense_f32_kernel(const T* __restrict__ in, float* __restrict__ out, int N, int C, int H, int W,
                                    int CP, int to_nchw) {
    pdl_start();
    int64_t total = (int64_t)N * C * H * W;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int n, c, y, x;
        if (to_nchw) {
            x = (int)(i % W);
            int64_t q = i / W;
            y = (int)(q % H);
            q /= H;
            c = (int)(q % C);
            n = (int)(q / C);
        } else {
            c = (int)(i % C);
            int64_t q = i / C;
            x = (int)(q % W);
            q /= W;
            y = (int)(q % H);
            n = (int)(q / H);
        }
        out[i] = ST<T>::ld(in + (((int64_t)n * H + y) * W + x) * CP + c);
    }
}

// ----------------------------------------------------------------------------------------
// dense convolution / linear / 2x2-stride-2 transposed conv as an implicit GEMM on CUDA cores
//   M = N*Ho*Wo output pixels, Ncol = Cout (x4 taps for convT), K = kh*kw*CinP
//   weights: fp32 [Kpad][ColsW] (K index = tap*CinP + ci), zero padded
// ----------------------------------------------------------------------------------------
struct ConvP {
    int N, H, W, CinP;
    int Ho, Wo, Cout, CoutP;
    int kh, kw, sh, sw, pt, pl;
    int K;          // kh*kw*CinP
    int ColsW;      // padded weight row length
    int Cols;       // GEMM columns that carry data (CoutP, or 4*CoutP for convT)
    int convT;      // 1: columns are (tap, co); store with pixel shuffle to (2y+dy, 2x+dx)
    int M;
    Epilogue epi;
};

constexpr int IG_BK = 16;

template <typename T, int BM, int BN>
__global__ void __launch_bounds__(256) conv_igemm_kernel(const T* __restrict__ x, const float* __restrict__ w,
                                                          const float* __restrict__ bias, T* __restrict__ y,
                                                          const ConvP p) {
    pdl_start();
    constexpr int TM = 4, TN = 4;
    constexpr int TX = BN / TN;            // threads along columns
    constexpr int TY = BM / TM;            // threads along rows
    static_assert(TX * TY == 256, "tile must map to 256 threads");
    constexpr int A_LD = BM / 64;          // float4 A loads per thread per chunk (BM*4 groups / 256)
    __shared__ __align__(16) float As[IG_BK][BM + 4];
    __shared__ __align__(16) float Bs[IG_BK][BN];

    const int tid = threadIdx.x;
    const int tx = tid % TX, ty = tid / TX;
    const int m0 = blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    // A-load assignment: slot s = tid + j*256 -> row = s / 4, kgroup = s % 4
    int a_row[A_LD], a_kg[A_LD], a_n[A_LD], a_oy[A_LD], a_ox[A_LD];
    bool a_valid[A_LD];
#pragma unroll
    for (int j = 0; j < A_LD; ++j) {
        int s = tid + j * 256;
        a_row[j] = s >> 2;
        a_kg[j] = s & 3;
        int m = m0 + a_row[j];
        a_valid[j] = m < p.M;
        int mm = a_valid[j] ? m : 0;
        a_ox[j] = mm % p.Wo;
        int q = mm / p.Wo;
        a_oy[j] = q % p.Ho;
        a_n[j] = q / p.Ho;
    }
    // B-load assignment: 16 x BN floats = 4*BN float4; thread loads float4 index tid (if < 4*BN)
    constexpr int B_F4 = IG_BK * BN / 4;
    const int b_k = tid / (BN / 4), b_c = (tid % (BN / 4)) * 4;

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    float a_reg[A_LD][4];
    float4 b_reg = make_float4(0.f, 0.f, 0.f, 0.f);

    auto load_chunk = [&](int k0) {
#pragma unroll
        for (int j = 0; j < A_LD; ++j) {
            int k = k0 + a_kg[j] * 4;
            a_reg[j][0] = a_reg[j][1] = a_reg[j][2] = a_reg[j][3] = 0.f;
            if (a_valid[j] && k < p.K) {
                int tap = k / p.CinP;
                int ci = k - tap * p.CinP;
                int ky = tap / p.kw, kx = tap - ky * p.kw;
                int iy = a_oy[j] * p.sh - p.pt + ky;
                int ix = a_ox[j] * p.sw - p.pl + kx;
                if ((unsigned)iy < (unsigned)p.H && (unsigned)ix < (unsigned)p.W)
                    ST<T>::ld4(x + (((int64_t)a_n[j] * p.H + iy) * p.W + ix) * p.CinP + ci, a_reg[j]);
            }
        }
        if (tid < B_F4) b_reg = *reinterpret_cast<const float4*>(w + (int64_t)(k0 + b_k) * p.ColsW + n0 + b_c);
    };
    auto store_chunk = [&]() {
#pragma unroll
        for (int j = 0; j < A_LD; ++j) {
#pragma unroll
            for (int e = 0; e < 4; ++e) As[a_kg[j] * 4 + e][a_row[j]] = a_reg[j][e];
        }
        if (tid < B_F4) *reinterpret_cast<float4*>(&Bs[b_k][b_c]) = b_reg;
    };

    const int nchunks = (p.K + IG_BK - 1) / IG_BK;
    load_chunk(0);
    for (int ch = 0; ch < nchunks; ++ch) {
        store_chunk();
        __syncthreads();
        if (ch + 1 < nchunks) load_chunk((ch + 1) * IG_BK);
#pragma unroll
        for (int k = 0; k < IG_BK; ++k) {
            float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * TM]);
            float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * TN]);
            float av[4] = {a4.x, a4.y, a4.z, a4.w};
            float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }

    // epilogue
    const int col0 = n0 + tx * TN;
    if (col0 >= p.Cols) return;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int m = m0 + ty * TM + i;
        if (m >= p.M) continue;
        float v[4];
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int col = col0 + j;
            int co = p.convT ? col % p.CoutP : col;
            float r = 0.f;
            if (col < p.Cols && co < p.Cout) r = epi_apply(acc[i][j] + bias[co], p.epi);
            v[j] = r;
        }
        if (!p.convT) {
            T* dst = y + (int64_t)m * p.CoutP + col0;
            if ((p.CoutP & 3) == 0) {
                ST<T>::st4(dst, v);
            } else {
#pragma unroll
                for (int j = 0; j < TN; ++j)
                    if (col0 + j < p.CoutP) ST<T>::st(dst + j, v[j]);
            }
        } else {
            int ox = m % p.Wo;
            int q = m / p.Wo;
            int oy = q % p.Ho;
            int n = q / p.Ho;
            if ((p.CoutP & 3) == 0) {
                int tap = col0 / p.CoutP, co = col0 - tap * p.CoutP;
                int dy = tap >> 1, dx = tap & 1;
                T* dst = y + ((((int64_t)n * (2 * p.Ho) + 2 * oy + dy) * (2 * p.Wo)) + 2 * ox + dx) * p.CoutP + co;
                ST<T>::st4(dst, v);
            } else {
#pragma unroll
                for (int j = 0; j < TN; ++j) {
                    int col = col0 + j;
                    if (col >= p.Cols) continue;
                    int tap = col / p.CoutP, co = col - tap * p.CoutP;
                    int dy = tap >> 1, dx = tap & 1;
                    ST<T>::st(y + ((((int64_t)n * (2 * p.Ho) + 2 * oy + dy) * (2 * p.Wo)) + 2 * ox + dx) * p.CoutP + co,
                              v[j]);
                }
            }
        }
    }
}

// ----------------------------------------------------------------------------------------
// depthwise k x k convolution, V channels per thread (V*sizeof(T) == 16 bytes or 4 channels)
// ----------------------------------------------------------------------------------------
struct DwP {
    int N, H, W, CP, C;
    int Ho, Wo;
    int kh, kw, sh, sw, pt, pl;
    Epilogue epi;
    unsigned long long* gap = nullptr;   // strip kernel: per (image, channel) sums of the stored outputs for a following SE gate
    int tma = 0;                         // strip kernel: the halo tile arrives as ONE 4-D TMA box (see dwconv_strip_kernel)
};

// Tensor map of an NHWC bf16 activation for the strip kernel's halo tile: box = {2 * WPP channels, IC columns, IR rows, 1},
// no swizzle (the tile lands exactly as the cp.async path lays it out), zero fill outside the tensor.  Defined in engine.cu.
struct DwTmap {
    CUtensorMap map;
    const void* base = nullptr;
    int box_c = 0, box_w = 0, box_h = 0;
};
void dw_encode_tmap(DwTmap* m, const void* base, int N, int H, int W, int CP, int box_c, int box_w, int box_h);
bool dw_tma_enabled();

__device__ __forceinline__ uint32_t dw_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <typename T>
__global__ void __launch_bounds__(256) dwconv_kernel(const T* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ bias, T* __restrict__ y, const DwP p) {
    pdl_start();
    const int cg = p.CP >> 2;                       // 4-channel groups
    int64_t total = (int64_t)p.N * p.Ho * p.Wo * cg;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int g = (int)(i % cg);
        int64_t q = i / cg;
        int ox = (int)(q % p.Wo);
        q /= p.Wo;
        int oy = (int)(q % p.Ho);
        int n = (int)(q / p.Ho);
        int c = g * 4;
        float acc[4];
        float4 b4 = *reinterpret_cast<const float4*>(bias + c);
        acc[0] = b4.x; acc[1] = b4.y; acc[2] = b4.z; acc[3] = b4.w;
        for (int ky = 0; ky < p.kh; ++ky) {
            int iy = oy * p.sh - p.pt + ky;
            if ((unsigned)iy >= (unsigned)p.H) continue;
            for (int kx = 0; kx < p.kw; ++kx) {
                int ix = ox * p.sw - p.pl + kx;
                if ((unsigned)ix >= (unsigned)p.W) continue;
                float v[4];
                ST<T>::ld4(x + (((int64_t)n * p.H + iy) * p.W + ix) * p.CP + c, v);
                float4 w4 = *reinterpret_cast<const float4*>(w + (int64_t)(ky * p.kw + kx) * p.CP + c);
                acc[0] = fmaf(v[0], w4.x, acc[0]);
                acc[1] = fmaf(v[1], w4.y, acc[1]);
                acc[2] = fmaf(v[2], w4.z, acc[2]);
                acc[3] = fmaf(v[3], w4.w, acc[3]);
            }
        }
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = (c + j < p.C) ? epi_apply(acc[j], p.epi) : 0.f;
        ST<T>::st4(y + (((int64_t)n * p.Ho + oy) * p.Wo + ox) * p.CP + c, o);
    }
}

// Depthwise convolution with the input tile (halo included) staged in shared memory by coalesced 16-byte loads; each
// thread owns one 8-channel group and walks the pixels of the 16 x 8 output tile assigned to it.  3x3 weights live in
// registers (packed bf16), 5x5 weights in shared memory.  HBM traffic = input + output once; the k^2 re-reads hit smem.
// CC = channels per block (16 / 32 / 64), chosen on the host to minimise channel padding.
template <int K, int SH, int SW, int CC>
__global__ void __launch_bounds__(256) dwconv_tile_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w,
                                                           const float* __restrict__ bias, bf16* __restrict__ y, const DwP p,
                                                           int tiles_x, int tiles_y) {
    pdl_start();
    constexpr int TW = 16, TH = 8, G = CC / 8;
    constexpr int IR = (TH - 1) * SH + K, IC = (TW - 1) * SW + K;
    constexpr int SLOTS = 256 / G;                // pixels handled per pass
    constexpr int PASSES = (TW * TH) / SLOTS;
    constexpr bool W_SMEM = K > 3;
    extern __shared__ uint4 tile[];               // [IR][IC][G] (+ [K*K][G] weights when W_SMEM)
    uint4* wsm = tile + IR * IC * G;
    int t = blockIdx.x;
    const int tx = t % tiles_x;
    t /= tiles_x;
    const int ty = t % tiles_y;
    const int n = t / tiles_y;
    const int x0 = tx * TW, y0 = ty * TH;
    const int c0 = blockIdx.y * CC;
    const int iy0 = y0 * SH - p.pt, ix0 = x0 * SW - p.pl;
    for (int i = threadIdx.x; i < IR * IC * G; i += 256) {
        int gi = i % G, pix = i / G;
        int ic = pix % IC, ir = pix / IC;
        int iy = iy0 + ir, ix = ix0 + ic, ch = c0 + gi * 8;
        uint4 v = make_uint4(0, 0, 0, 0);
        if ((unsigned)iy < (unsigned)p.H && (unsigned)ix < (unsigned)p.W && ch < p.CP)
            v = *reinterpret_cast<const uint4*>(x + (((int64_t)n * p.H + iy) * p.W + ix) * p.CP + ch);
        tile[i] = v;
    }
    const int g = threadIdx.x % G, slot = threadIdx.x / G;
    const int c = c0 + g * 8;
    const bool active = c < p.CP;
    uint4 wreg[W_SMEM ? 1 : K * K];
    if (W_SMEM) {
        for (int i = threadIdx.x; i < K * K * G; i += 256) {
            int gi = i % G, k = i / G;
            int ch = c0 + gi * 8;
            wsm[i] = ch < p.CP ? *reinterpret_cast<const uint4*>(w + (int64_t)k * p.CP + ch) : make_uint4(0, 0, 0, 0);
        }
    } else {
#pragma unroll
        for (int k = 0; k < K * K; ++k)
            wreg[k] = active ? *reinterpret_cast<const uint4*>(w + (int64_t)k * p.CP + c) : make_uint4(0, 0, 0, 0);
    }
    float bb[8];
    {
        float4 b0 = active ? *reinterpret_cast<const float4*>(bias + c) : make_float4(0, 0, 0, 0);
        float4 b1 = active ? *reinterpret_cast<const float4*>(bias + c + 4) : make_float4(0, 0, 0, 0);
        bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b0.w; bb[4] = b1.x; bb[5] = b1.y; bb[6] = b1.z; bb[7] = b1.w;
    }
    __syncthreads();
    if (!active) return;
#pragma unroll
    for (int it = 0; it < PASSES; ++it) {
        int pixel = slot + it * SLOTS;
        int py = pixel / TW, px = pixel % TW;
        int oy = y0 + py, ox = x0 + px;
        if (oy >= p.Ho || ox >= p.Wo) continue;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = bb[j];
#pragma unroll
        for (int ky = 0; ky < K; ++ky)
#pragma unroll
            for (int kx = 0; kx < K; ++kx) {
                uint4 v = tile[((py * SH + ky) * IC + (px * SW + kx)) * G + g];
                uint4 ww = W_SMEM ? wsm[(ky * K + kx) * G + g] : wreg[W_SMEM ? 0 : ky * K + kx];
                acc[0] = fmaf(__uint_as_float(v.x << 16), __uint_as_float(ww.x << 16), acc[0]);
                acc[1] = fmaf(__uint_as_float(v.x & 0xffff0000u), __uint_as_float(ww.x & 0xffff0000u), acc[1]);
                acc[2] = fmaf(__uint_as_float(v.y << 16), __uint_as_float(ww.y << 16), acc[2]);
                acc[3] = fmaf(__uint_as_float(v.y & 0xffff0000u), __uint_as_float(ww.y & 0xffff0000u), acc[3]);
                acc[4] = fmaf(__uint_as_float(v.z << 16), __uint_as_float(ww.z << 16), acc[4]);
                acc[5] = fmaf(__uint_as_float(v.z & 0xffff0000u), __uint_as_float(ww.z & 0xffff0000u), acc[5]);
                acc[6] = fmaf(__uint_as_float(v.w << 16), __uint_as_float(ww.w << 16), acc[6]);
                acc[7] = fmaf(__uint_as_float(v.w & 0xffff0000u), __uint_as_float(ww.w & 0xffff0000u), acc[7]);
            }
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (c + j < p.C) ? epi_apply(acc[j], p.epi) : 0.f;
        uint4 pk;
        __nv_bfloat162 t0 = __floats2bfloat162_rn(o[0], o[1]);
        __nv_bfloat162 t1 = __floats2bfloat162_rn(o[2], o[3]);
        __nv_bfloat162 t2 = __floats2bfloat162_rn(o[4], o[5]);
        __nv_bfloat162 t3 = __floats2bfloat162_rn(o[6], o[7]);
        pk.x = *reinterpret_cast<uint32_t*>(&t0);
        pk.y = *reinterpret_cast<uint32_t*>(&t1);
        pk.z = *reinterpret_cast<uint32_t*>(&t2);
        pk.w = *reinterpret_cast<uint32_t*>(&t3);
        *reinterpret_cast<uint4*>(y + (((int64_t)n * p.Ho + oy) * p.Wo + ox) * p.CP + c) = pk;
    }
}

template <int K, int SH, int SW, int CC>
inline void launch_dwconv_tile_cc(const bf16* x, const bf16* w, const float* bias, bf16* y, const DwP& p, cudaStream_t st) {
    constexpr int TW = 16, TH = 8, G = CC / 8;
    constexpr int IR = (TH - 1) * SH + K, IC = (TW - 1) * SW + K;
    constexpr size_t smem = ((size_t)IR * IC * G + (K > 3 ? K * K * G : 0)) * 16;
    static bool attr[16] = {false};                     // per device: function attributes are per context
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr[dev & 15]) {
        cudaFuncSetAttribute(dwconv_tile_kernel<K, SH, SW, CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr[dev & 15] = true;
    }
    int tiles_x = (p.Wo + TW - 1) / TW, tiles_y = (p.Ho + TH - 1) / TH;
    dim3 grid((unsigned)(tiles_x * tiles_y * p.N), (unsigned)((p.CP + CC - 1) / CC));
    launch_pdl(dwconv_tile_kernel<K, SH, SW, CC>, grid, 256, smem, st, x, w, bias, y, p, tiles_x, tiles_y);
}

template <int K, int SH, int SW>
inline void launch_dwconv_tile(const bf16* x, const bf16* w, const float* bias, bf16* y, const DwP& p, cudaStream_t st) {
    // channels per block: the candidate with the least channel padding, larger preferred; 16 only for narrow tensors
    auto padded = [&](int cc) { return (p.CP + cc - 1) / cc * cc; };
    int cc = 64;
    if (padded(32) < padded(cc)) cc = 32;
    if (p.CP <= 48 && padded(16) < padded(cc)) cc = 16;
    if (cc == 64) launch_dwconv_tile_cc<K, SH, SW, 64>(x, w, bias, y, p, st);
    else if (cc == 32) launch_dwconv_tile_cc<K, SH, SW, 32>(x, w, bias, y, p, st);
    else launch_dwconv_tile_cc<K, SH, SW, 16>(x, w, bias, y, p, st);
}

// Depthwise convolution, register-sliding form.  The input tile (halo included) is staged in shared memory with
// 16-byte cp.async copies (zero-filled outside the image); each thread owns ONE 32-bit word (two bf16 channels) of ONE
// output column and walks down the tile's rows: every input row costs K conflict-free LDS.32 and feeds the K (or
// ceil(K/SH)) output rows in flight, so the k x k window is read K times instead of K*K times.  The two channels of a
// word are kept as one packed fp32x2 register pair and accumulated with FFMA2 (fma.rn.f32x2, sm_100): K*K packed FMAs
// per output pair, weights unpacked once into registers (fp32 — they are never rounded to bf16).
//   WPP = words per pixel handled by a block (8 / 16 / 32 -> 16 / 32 / 64 channels); TX = 256 / WPP output columns.
//   PS  = shared-memory pixel stride in words; padded 1.5x for SW == 2 with WPP < 32 so that the lanes of a warp
//         (which then read every second pixel) still hit 32 distinct banks.
typedef unsigned long long f32x2_t;             // two fp32 lanes in one 64-bit register pair

__device__ __forceinline__ f32x2_t pack_f32x2(float lo, float hi) {
    f32x2_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack_f32x2(f32x2_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2_t bf16x2_to_f32x2(uint32_t u) {
    return pack_f32x2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
__device__ __forceinline__ void ffma2(f32x2_t& acc, f32x2_t a, f32x2_t b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, bool valid) {
    uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}

template <int K, int SH, int SW, int WPP, int TY, int ACT, bool GAP>
__device__ __forceinline__ void dw_strip_rows(const uint32_t* __restrict__ col, const f32x2_t (&wf)[K * K], f32x2_t bb,
                                              bf16* __restrict__ yo, int64_t ystride, int rows_in, bool live0, bool live1,
                                              const Epilogue& epi, float& g0, float& g1) {
    constexpr int TX = 256 / WPP;
    constexpr int IR = (TY - 1) * SH + K, IC = (TX - 1) * SW + K;
    constexpr int PS = (SW == 2 && WPP < 32 && !B200_DW_DENSE_S2) ? WPP * 3 / 2 : WPP;
    f32x2_t acc[TY];
#pragma unroll
    for (int r = 0; r < IR; ++r) {
        if (r >= rows_in) break;                  // block-uniform
        f32x2_t in[K];
#pragma unroll
        for (int kx = 0; kx < K; ++kx) in[kx] = bf16x2_to_f32x2(col[(r * IC + kx) * PS]);
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {
            if (r - ky < 0 || (r - ky) % SH != 0 || (r - ky) / SH >= TY) continue;
            const int oy = (r - ky) / SH;
            if (ky == 0) acc[oy] = bb;                               // first row that touches this output
#pragma unroll
            for (int kx = 0; kx < K; ++kx) ffma2(acc[oy], in[kx], wf[ky * K + kx]);
            if (ky == K - 1) {                                       // last row: output complete
                float a0, a1;
                unpack_f32x2(acc[oy], a0, a1);
                float o0 = live0 ? epi_fast<ACT>(a0, epi) : 0.f;
                float o1 = live1 ? epi_fast<ACT>(a1, epi) : 0.f;
                __nv_bfloat162 pk = __floats2bfloat162_rn(o0, o1);
                const uint32_t u = *reinterpret_cast<uint32_t*>(&pk);
                *reinterpret_cast<uint32_t*>(yo + (int64_t)oy * ystride) = u;
                if (GAP) {                                           // sums of the values as stored (rows in order)
                    g0 += __uint_as_float(u << 16);
                    g1 += __uint_as_float(u & 0xffff0000u);
                }
            }
        }
    }
}

template <int K, int SH, int SW, int WPP, int TY, bool GAP>
__global__ void __launch_bounds__(256) dwconv_strip_kernel(const bf16* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ bias, bf16* __restrict__ y, const DwP p,
                                                            int tiles_x, int tiles_y, const __grid_constant__ CUtensorMap tmap) {
    pdl_start();
    constexpr int TX = 256 / WPP;
    constexpr int IR = (TY - 1) * SH + K, IC = (TX - 1) * SW + K;
    constexpr int PS = (SW == 2 && WPP < 32 && !B200_DW_DENSE_S2) ? WPP * 3 / 2 : WPP;
    constexpr int G = WPP / 4;                    // 16-byte chunks per pixel
    extern __shared__ __align__(128) uint32_t dw_tile[];         // [IR][IC][PS]
    __shared__ __align__(8) unsigned long long dw_bar;
    int t = blockIdx.x;
    const int tx = t % tiles_x;
    t /= tiles_x;
    const int ty = t % tiles_y;
    const int n = t / tiles_y;
    const int x0 = tx * TX, y0 = ty * TY;
    const int c0 = blockIdx.y * (2 * WPP);
    const int iy0 = y0 * SH - p.pt, ix0 = x0 * SW - p.pl;
    const int rows_out = min(TY, p.Ho - y0);
    const int rows_in = (rows_out - 1) * SH + K;
    // Halo tile -> shared memory.  With a tensor map (PS == WPP: the tile is dense) ONE thread issues ONE 4-D TMA box
    // {2 * WPP channels, IC, IR, 1} at (c0, ix0, iy0, n): padding and the ragged last tiles are the TMA unit's out-of-bounds
    // zero fill, and the ~25 address / predicate instructions per 16-byte chunk of the cp.async loop below (a third of
    // the warp-issue samples of these kernels in ncu) disappear.
    const bool use_tma = PS == WPP && p.tma;
    if (use_tma) {
        const uint32_t bar = dw_smem_u32(&dw_bar);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(IR * IC * PS * 4)) : "memory");
            asm volatile(
                "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                ::"r"(dw_smem_u32(dw_tile)), "l"(&tmap), "r"(bar), "r"(c0), "r"(ix0), "r"(iy0), "r"(n)
                : "memory");
        }
    } else {
        // thread -> (pixel column block, 16-byte chunk); rows advance by a fixed step so the index math is hoisted
        const int gi = threadIdx.x % G;
        const int ch = c0 + gi * 8;
        const bool ch_ok = ch < p.CP;
        constexpr int STEP = 256 / G;             // pixels advanced per iteration
        int pix = threadIdx.x / G;
        int ir = pix / IC, ic = pix - ir * IC;
        for (; pix < rows_in * IC; pix += STEP) {
            int iy = iy0 + ir, ix = ix0 + ic;
            bool ok = ch_ok && (unsigned)iy < (unsigned)p.H && (unsigned)ix < (unsigned)p.W;
            const bf16* src = ok ? x + (((int64_t)n * p.H + iy) * p.W + ix) * p.CP + ch : x;
            cp_async16_zfill(dw_tile + pix * PS + gi * 4, src, ok);
            ic += STEP % IC;
            ir += STEP / IC;
            if (ic >= IC) { ic -= IC; ++ir; }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    const int wl = threadIdx.x % WPP, tc = threadIdx.x / WPP;
    const int c = c0 + 2 * wl;
    const bool active = c < p.CP && x0 + tc < p.Wo;
    f32x2_t wf[K * K];
    f32x2_t bb = 0;
    if (active) {
#pragma unroll
        for (int k = 0; k < K * K; ++k) wf[k] = *reinterpret_cast<const f32x2_t*>(w + (int64_t)k * p.CP + c);
        bb = *reinterpret_cast<const f32x2_t*>(bias + c);
    }
    if (use_tma) {
        __syncthreads();                          // the barrier's initialisation is visible to every waiter
        const uint32_t bar = dw_smem_u32(&dw_bar);
        uint32_t ok = 0;
        while (!ok) {
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n"
                "selp.u32 %0, 1, 0, p;\n"
                "}\n"
                : "=r"(ok)
                : "r"(bar)
                : "memory");
        }
    } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
    }
    if (!active && !GAP) return;
    float g0 = 0.f, g1 = 0.f;
    if (active) {
        const bool live0 = c < p.C, live1 = c + 1 < p.C;
        const uint32_t* col = dw_tile + (tc * SW) * PS + wl;
        bf16* yo = y + (((int64_t)n * p.Ho + y0) * p.Wo + (x0 + tc)) * p.CP + c;
        const int64_t ystride = (int64_t)p.Wo * p.CP;
        // the activation is a launch constant: branch once, run a body compiled for it
        switch (p.epi.fast ? p.epi.act : -1) {
            case 0: dw_strip_rows<K, SH, SW, WPP, TY, 0, GAP>(col, wf, bb, yo, ystride, rows_in, live0, live1, p.epi, g0, g1); break;
            case EPI_RELU: dw_strip_rows<K, SH, SW, WPP, TY, EPI_RELU, GAP>(col, wf, bb, yo, ystride, rows_in, live0, live1, p.epi, g0, g1); break;
            case EPI_HSWISH: dw_strip_rows<K, SH, SW, WPP, TY, EPI_HSWISH, GAP>(col, wf, bb, yo, ystride, rows_in, live0, live1, p.epi, g0, g1); break;
            default: dw_strip_rows<K, SH, SW, WPP, TY, -1, GAP>(col, wf, bb, yo, ystride, rows_in, live0, live1, p.epi, g0, g1); break;
        }
    }
    if (!GAP) return;
    // channel sums for the SE gate that follows: per-thread column sums -> shared memory -> one value per channel of this
    // tile (columns in fixed order) -> fixed-point atomics (common.cuh)
    __syncthreads();                              // every thread has finished reading the tile
    float* red = reinterpret_cast<float*>(dw_tile);          // [TX][2 * WPP]
    red[tc * (2 * WPP) + 2 * wl] = g0;
    red[tc * (2 * WPP) + 2 * wl + 1] = g1;
    __syncthreads();
    if (threadIdx.x < 2 * WPP) {
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < TX; ++q) s += red[q * (2 * WPP) + threadIdx.x];
        const int cc = c0 + (int)threadIdx.x;
        if (cc < p.C) gap_fix_add(p.gap + (size_t)n * p.CP + cc, s);
    }
}

template <int K, int SH, int SW, int WPP>
inline void launch_dwconv_strip_w(const bf16* x, const float* w, const float* bias, bf16* y, const DwP& p, cudaStream_t st, DwTmap* tm) {
    constexpr int TY = SH == 2 ? 8 : 16;
    constexpr int TX = 256 / WPP;
    constexpr int IR = (TY - 1) * SH + K, IC = (TX - 1) * SW + K;
    constexpr int PS = (SW == 2 && WPP < 32 && !B200_DW_DENSE_S2) ? WPP * 3 / 2 : WPP;
    constexpr size_t smem = (size_t)IR * IC * PS * 4;
    static bool attr[16] = {false};                     // per device: function attributes are per context
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr[dev & 15]) {
        cudaFuncSetAttribute(dwconv_strip_kernel<K, SH, SW, WPP, TY, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(dwconv_strip_kernel<K, SH, SW, WPP, TY, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr[dev & 15] = true;
    }
    int tiles_x = (p.Wo + TX - 1) / TX, tiles_y = (p.Ho + TY - 1) / TY;
    dim3 grid((unsigned)(tiles_x * tiles_y * p.N), (unsigned)((p.CP + 2 * WPP - 1) / (2 * WPP)));
    DwP pk = p;
    static const DwTmap no_map{};
    const DwTmap* tmu = &no_map;
    if (PS == WPP && tm && IC <= 256 && IR <= 256 && dw_tma_enabled()) {
        if (tm->base != (const void*)x || tm->box_c != 2 * WPP || tm->box_w != IC || tm->box_h != IR)
            dw_encode_tmap(tm, x, p.N, p.H, p.W, p.CP, 2 * WPP, IC, IR);
        pk.tma = 1;
        tmu = tm;
    }
    // the channel-sum variant is a separate instantiation: its extra registers would cost every other layer occupancy
    if (p.gap) launch_pdl(dwconv_strip_kernel<K, SH, SW, WPP, TY, true>, grid, 256, smem, st, x, w, bias, y, pk, tiles_x, tiles_y, tmu->map);
    else launch_pdl(dwconv_strip_kernel<K, SH, SW, WPP, TY, false>, grid, 256, smem, st, x, w, bias, y, pk, tiles_x, tiles_y, tmu->map);
}

template <int K, int SH, int SW>
inline void launch_dwconv_strip(const bf16* x, const float* w, const float* bias, bf16* y, const DwP& p, cudaStream_t st, DwTmap* tm = nullptr) {
    // channels per block: the candidate with the least channel padding, larger preferred
    auto padded = [&](int cc) { return (p.CP + cc - 1) / cc * cc; };
    int cc = 64;
    if (padded(32) < padded(cc)) cc = 32;
    if (padded(16) < padded(cc)) cc = 16;
    if (cc == 64) launch_dwconv_strip_w<K, SH, SW, 32>(x, w, bias, y, p, st, tm);
    else if (cc == 32) launch_dwconv_strip_w<K, SH, SW, 16>(x, w, bias, y, p, st, tm);
    else launch_dwconv_strip_w<K, SH, SW, 8>(x, w, bias, y, p, st, tm);
}

// ----------------------------------------------------------------------------------------
// global average pool (deterministic two-stage) + squeeze-excite
// ----------------------------------------------------------------------------------------
// partial[n][chunk][CP] = sum over the chunk's pixels.  grid (chunks, N), block 256.
template <typename T>
__global__ void __launch_bounds__(256) gap_partial_kernel(const T* __restrict__ x, float* __restrict__ partial, int HW,
                                                           int CP, int chunks) {
    pdl_start();
    extern __shared__ float sm[];                   // [rows][CP]
    const int cg = CP >> 2;
    const int rows = max(1, 256 / cg);
    const int n = blockIdx.y, chunk = blockIdx.x;
    const int per = (HW + chunks - 1) / chunks;
    const int p0 = chunk * per, p1 = min(HW, p0 + per);
    const int r = threadIdx.x / cg, g = threadIdx.x % cg;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (r < rows) {
        for (int px = p0 + r; px < p1; px += rows) {
            float v[4];
            ST<T>::ld4(x + ((int64_t)n * HW + px) * CP + g * 4, v);
            acc[0] += v[0]; acc[1] += v[1]; acc[2] += v[2]; acc[3] += v[3];
        }
        for (int j = 0; j < 4; ++j) sm[r * CP + g * 4 + j] = acc[j];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < CP; c += blockDim.x) {
        float s = 0.f;
        for (int rr = 0; rr < rows; ++rr) s += sm[rr * CP + c];
        partial[((int64_t)n * chunks + chunk) * CP + c] = s;
    }
}

// GAP as a tensor: out[n][c] = mean
template <typename T>
__global__ void gap_finish_kernel(const float* __restrict__ partial, T* __restrict__ out, int HW, int C, int CP,
                                  int chunks) {
    pdl_start();
    int n = blockIdx.x;
    for (int c = threadIdx.x; c < CP; c += blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < chunks; ++k) s += partial[((int64_t)n * chunks + k) * CP + c];
        ST<T>::st(out + (int64_t)n * CP + c, c < C ? s / (float)HW : 0.f);
    }
}

// scale[n][c] = hsigmoid(W2 relu(W1 mean + b1) + b2).  One block of 1024 threads per image; both small matrix-vector
// products are split over thread slices (coalesced weight reads along the output index, partial sums through shared
// memory) so the dependent-load chains stay short — the kernel is pure latency.
// w1: [C][mid] (row stride w1_ld), w2: [mid][C] (row stride w2_ld).
constexpr int kSeThreads = 1024;
// The three phases of the SE gate for image n, executed by one whole block of NT threads (shared memory: mean[CP] +
// hidden[mid] + part[NT]).  Partials are read with ld.global.cg: in the fused kernel they were written by other blocks.
__device__ __forceinline__ void se_fc_body(const float* __restrict__ partial, int n, int chunks, int HW, int C, int CP, int mid,
                                           const float* __restrict__ w1, int w1_ld, const float* __restrict__ b1,
                                           const float* __restrict__ w2, int w2_ld, const float* __restrict__ b2, float alpha,
                                           float beta, float* __restrict__ scale, float* sm) {
    float* mean = sm;
    float* hid = sm + CP;
    float* part = hid + mid;
    const int NT = blockDim.x, tid = threadIdx.x;
    {   // mean over the pixels: slices of the chunk list per channel
        const int S = max(1, min(chunks, NT / CP));
        const int c = tid % CP, sl = tid / CP;
        float s = 0.f;
        if (sl < S)
            for (int k = sl; k < chunks; k += S) s += __ldcg(partial + ((int64_t)n * chunks + k) * CP + c);
        if (sl < S) part[sl * CP + c] = s;
        __syncthreads();
        for (int cc = tid; cc < CP; cc += NT) {
            float t = 0.f;
            for (int q = 0; q < S; ++q) t += part[q * CP + cc];
            mean[cc] = cc < C ? t / (float)HW : 0.f;
        }
        __syncthreads();
    }
    {   // hidden = relu(W1^T mean + b1)
        const int S = max(1, NT / mid);
        const int j = tid % mid, sl = tid / mid;
        float s = 0.f;
        if (sl < S)
            for (int c = sl; c < C; c += S) s = fmaf(mean[c], w1[(int64_t)c * w1_ld + j], s);
        if (sl < S) part[sl * mid + j] = s;
        __syncthreads();
        for (int jj = tid; jj < mid; jj += NT) {
            float t = 0.f;
            for (int q = 0; q < S; ++q) t += part[q * mid + jj];
            hid[jj] = fmaxf(t + b1[jj], 0.f);
        }
        __syncthreads();
    }
    {   // scale = hsigmoid(W2^T hidden + b2)
        const int S = max(1, min(mid, NT / CP));
        const int c = tid % CP, sl = tid / CP;
        float s = 0.f;
        if (sl < S && c < C)
            for (int j = sl; j < mid; j += S) s = fmaf(hid[j], w2[(int64_t)j * w2_ld + c], s);
        if (sl < S) part[sl * CP + c] = s;
        __syncthreads();
        for (int cc = tid; cc < CP; cc += NT) {
            float t = 0.f;
            if (cc < C) {
                for (int q = 0; q < S; ++q) t += part[q * CP + cc];
                t = fminf(fmaxf(fmaf(t + b2[cc], alpha, beta), 0.f), 1.f);
            }
            scale[(int64_t)n * CP + cc] = t;
        }
    }
}

__global__ void __launch_bounds__(kSeThreads) se_fc_kernel(const float* __restrict__ partial, int chunks, int HW, int C, int CP,
                                                           int mid, const float* __restrict__ w1, int w1_ld,
                                                           const float* __restrict__ b1, const float* __restrict__ w2, int w2_ld,
                                                           const float* __restrict__ b2, float alpha, float beta,
                                                           float* __restrict__ scale) {
    pdl_start();
    extern __shared__ float sm[];                   // mean[CP] + hidden[mid] + part[blockDim.x]
    se_fc_body(partial, blockIdx.x, chunks, HW, C, CP, mid, w1, w1_ld, b1, w2, w2_ld, b2, alpha, beta, scale, sm);
}

// ---- SE gate from channel sums that the PRODUCING kernel accumulated (no GAP pass over the tensor at all).
// The producer (tensor-core conv epilogue, depthwise strip kernel) adds, per image and channel, the sum of the bf16 values
// it stored as 48.16 fixed point with 64-bit integer atomics: integer addition is associative, so the total does not
// depend on the order in which tiles finish (deterministic run to run) nor on where the image sits in the batch (every
// tile belongs to one image and its partial sum is formed in a fixed order).  This kernel turns the sums into means,
// re-zeroes them for the next launch, and runs the two small matrix-vector products with all loads of a slice issued up
// front (the gate is pure latency: one block per image).
constexpr int kSeSumsThreads = 1024;
__global__ void __launch_bounds__(kSeSumsThreads) se_fc_sums_kernel(unsigned long long* __restrict__ sums, int replicas, int HW, int C, int CP,
                                                                    int mid, const float* __restrict__ w1, int w1_ld,
                                                                    const float* __restrict__ b1, const float* __restrict__ w2,
                                                                    int w2_ld, const float* __restrict__ b2, float alpha, float beta,
                                                                    float* __restrict__ scale) {
    pdl_start();
    extern __shared__ float sm[];                   // mean[CP] + hidden[mid] + part[blockDim.x]
    float* mean = sm;
    float* hid = sm + CP;
    float* part = hid + mid;
    const int n = blockIdx.x, NT = blockDim.x, tid = threadIdx.x;
    for (int c = tid; c < CP; c += NT) {
        long long v = 0;
        for (int r = 0; r < replicas; ++r) {        // producers spread their atomics over `replicas` copies
            unsigned long long* q = sums + ((size_t)n * replicas + r) * CP + c;
            v += (long long)__ldcg(q);
            *q = 0ull;                              // ready for the next launch of the plan
        }
        mean[c] = c < C ? (float)((double)v * (1.0 / (double)kGapFixScale) / (double)HW) : 0.f;
    }
    __syncthreads();
    {   // hidden = relu(W1^T mean + b1): thread (j, slice) takes channels slice, slice + S, ...
        const int S = max(1, NT / mid);
        const int j = tid % mid, sl = tid / mid;
        float s = 0.f;
        if (sl < S) {
            int c = sl;
            for (; c + 3 * S < C; c += 4 * S) {     // four independent loads in flight
                const float a0 = w1[(int64_t)c * w1_ld + j], a1 = w1[(int64_t)(c + S) * w1_ld + j];
                const float a2 = w1[(int64_t)(c + 2 * S) * w1_ld + j], a3 = w1[(int64_t)(c + 3 * S) * w1_ld + j];
                s = fmaf(mean[c], a0, s);
                s = fmaf(mean[c + S], a1, s);
                s = fmaf(mean[c + 2 * S], a2, s);
                s = fmaf(mean[c + 3 * S], a3, s);
            }
            for (; c < C; c += S) s = fmaf(mean[c], w1[(int64_t)c * w1_ld + j], s);
            part[sl * mid + j] = s;
        }
        __syncthreads();
        for (int jj = tid; jj < mid; jj += NT) {
            float t = 0.f;
            for (int q = 0; q < S; ++q) t += part[q * mid + jj];
            hid[jj] = fmaxf(t + b1[jj], 0.f);
        }
        __syncthreads();
    }
    {   // scale = hsigmoid(W2^T hidden + b2)
        const int S = max(1, min(mid, NT / CP));
        const int c = tid % CP, sl = tid / CP;
        float s = 0.f;
        if (sl < S) {
            if (c < C) {
                int j = sl;
                for (; j + 3 * S < mid; j += 4 * S) {
                    const float a0 = w2[(int64_t)j * w2_ld + c], a1 = w2[(int64_t)(j + S) * w2_ld + c];
                    const float a2 = w2[(int64_t)(j + 2 * S) * w2_ld + c], a3 = w2[(int64_t)(j + 3 * S) * w2_ld + c];
                    s = fmaf(hid[j], a0, s);
                    s = fmaf(hid[j + S], a1, s);
                    s = fmaf(hid[j + 2 * S], a2, s);
                    s = fmaf(hid[j + 3 * S], a3, s);
                }
                for (; j < mid; j += S) s = fmaf(hid[j], w2[(int64_t)j * w2_ld + c], s);
            }
            part[sl * CP + c] = s;
        }
        __syncthreads();
        for (int cc = tid; cc < CP; cc += NT) {
            float t = 0.f;
            if (cc < C) {
                for (int q = 0; q < S; ++q) t += part[q * CP + cc];
                t = fminf(fmaxf(fmaf(t + b2[cc], alpha, beta), 0.f), 1.f);
            }
            scale[(int64_t)n * CP + cc] = t;
        }
    }
}

// GAP partial sums and the SE gate in ONE launch: every block reduces its pixel chunk; the block that finishes last for
// an image (ticket counter, self-resetting) sums the partials in chunk order — deterministic whichever block it is —
// and runs the two small matrix-vector products.  Saves one dependent launch per SE layer (10 of the det plan's 80).
// grid (chunks, N), block 512; needs CP <= 512 (the slicing of se_fc_body handles any CP <= blockDim.x) .
constexpr int kGapSeThreads = 512;
template <typename T>
__global__ void __launch_bounds__(kGapSeThreads) gap_se_kernel(const T* __restrict__ x, float* __restrict__ partial, int HW, int C,
                                                               int CP, int chunks, int mid, const float* __restrict__ w1, int w1_ld,
                                                               const float* __restrict__ b1, const float* __restrict__ w2,
                                                               int w2_ld, const float* __restrict__ b2, float alpha, float beta,
                                                               float* __restrict__ scale, int* __restrict__ tickets) {
    pdl_start();
    extern __shared__ float sm[];                   // max([rows][CP], CP + mid + blockDim.x)
    __shared__ int s_last;
    const int cg = CP >> 2;
    const int rows = max(1, kGapSeThreads / cg);
    const int n = blockIdx.y, chunk = blockIdx.x;
    const int per = (HW + chunks - 1) / chunks;
    const int p0 = chunk * per, p1 = min(HW, p0 + per);
    const int r = threadIdx.x / cg, g = threadIdx.x % cg;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (r < rows) {
        // four independent loads in flight per thread (the per-thread accumulation order px, px+rows, ... is unchanged)
        int px = p0 + r;
        const T* base = x + (int64_t)n * HW * CP + g * 4;
        for (; px + 3 * rows < p1; px += 4 * rows) {
            float v0[4], v1[4], v2[4], v3[4];
            ST<T>::ld4(base + (int64_t)px * CP, v0);
            ST<T>::ld4(base + (int64_t)(px + rows) * CP, v1);
            ST<T>::ld4(base + (int64_t)(px + 2 * rows) * CP, v2);
            ST<T>::ld4(base + (int64_t)(px + 3 * rows) * CP, v3);
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j] = (((acc[j] + v0[j]) + v1[j]) + v2[j]) + v3[j];
        }
        for (; px < p1; px += rows) {
            float v[4];
            ST<T>::ld4(base + (int64_t)px * CP, v);
            acc[0] += v[0]; acc[1] += v[1]; acc[2] += v[2]; acc[3] += v[3];
        }
        for (int j = 0; j < 4; ++j) sm[r * CP + g * 4 + j] = acc[j];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < CP; c += blockDim.x) {
        float s = 0.f;
        for (int rr = 0; rr < rows; ++rr) s += sm[rr * CP + c];
        partial[((int64_t)n * chunks + chunk) * CP + c] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = atomicAdd(&tickets[n], 1);
        s_last = (t == chunks - 1);
        if (s_last) tickets[n] = 0;                 // ready for the next launch
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    se_fc_body(partial, n, chunks, HW, C, CP, mid, w1, w1_ld, b1, w2, w2_ld, b2, alpha, beta, scale, sm);
}

// y = epi(x * scale[n][c]) (+ x when residual)
template <typename T>
__global__ void se_apply_kernel(const T* __restrict__ x, const float* __restrict__ scale, T* __restrict__ y, int64_t HW,
                                int CP, int C, int64_t total4, int residual, const Epilogue epi) {
    pdl_start();
    const int cg = CP >> 2;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
        int g = (int)(i % cg);
        int64_t px = i / cg;
        int n = (int)(px / HW);
        float v[4], o[4];
        ST<T>::ld4(x + i * 4, v);
        float4 s = *reinterpret_cast<const float4*>(scale + (int64_t)n * CP + g * 4);
        float sv[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float t = epi_apply(v[j] * sv[j], epi);
            if (residual) t = v[j] + t;
            o[j] = (g * 4 + j < C) ? t : 0.f;
        }
        ST<T>::st4(y + i * 4, o);
    }
}

// bf16 fast path of se_apply_kernel: 8 channels (16 bytes) per thread, 32-bit index arithmetic, packed fp32x2 math.
__global__ void __launch_bounds__(256) se_apply_bf16x8_kernel(const bf16* __restrict__ x, const float* __restrict__ scale,
                                                              bf16* __restrict__ y, uint32_t HW, int CP, int C, uint32_t total8,
                                                              int residual, const Epilogue epi) {
    pdl_start();
    const uint32_t cg = (uint32_t)CP >> 3;
    const uint32_t per_img = HW * cg;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += gridDim.x * blockDim.x) {
        const uint32_t n = i / per_img;
        const uint32_t g = i % cg;
        const uint4 v = reinterpret_cast<const uint4*>(x)[i];
        const float4 s0 = *reinterpret_cast<const float4*>(scale + (size_t)n * CP + g * 8);
        const float4 s1 = *reinterpret_cast<const float4*>(scale + (size_t)n * CP + g * 8 + 4);
        const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        const uint32_t* u = &v.x;
        uint4 o;
        uint32_t* uo = &o.x;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float a0 = __uint_as_float(u[j] << 16), a1 = __uint_as_float(u[j] & 0xffff0000u);
            float t0 = a0 * sv[2 * j], t1 = a1 * sv[2 * j + 1];
            if (epi.n) { t0 = epi_apply(t0, epi); t1 = epi_apply(t1, epi); }
            if (residual) { t0 += a0; t1 += a1; }
            int c = (int)g * 8 + 2 * j;
            __nv_bfloat162 r = __floats2bfloat162_rn(c < C ? t0 : 0.f, c + 1 < C ? t1 : 0.f);
            uo[j] = *reinterpret_cast<uint32_t*>(&r);
        }
        reinterpret_cast<uint4*>(y)[i] = o;
    }
}

// ----------------------------------------------------------------------------------------
// pooling
// ----------------------------------------------------------------------------------------
struct PoolP {
    int N, H, W, CP, C, Ho, Wo, kh, kw, sh, sw, pt, pl, is_max, count_include_pad;
};
template <typename T>
__global__ void pool_kernel(const T* __restrict__ x, T* __restrict__ y, const PoolP p) {
    pdl_start();
    int64_t total = (int64_t)p.N * p.Ho * p.Wo * p.CP;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % p.CP);
        int64_t q = i / p.CP;
        int ox = (int)(q % p.Wo);
        q /= p.Wo;
        int oy = (int)(q % p.Ho);
        int n = (int)(q / p.Ho);
        float acc = p.is_max ? -3.4e38f : 0.f;
        int cnt = 0;
        for (int ky = 0; ky < p.kh; ++ky) {
            int iy = oy * p.sh - p.pt + ky;
            if ((unsigned)iy >= (unsigned)p.H) continue;
            for (int kx = 0; kx < p.kw; ++kx) {
                int ix = ox * p.sw - p.pl + kx;
                if ((unsigned)ix >= (unsigned)p.W) continue;
                float v = ST<T>::ld(x + (((int64_t)n * p.H + iy) * p.W + ix) * p.CP + c);
                acc = p.is_max ? fmaxf(acc, v) : acc + v;
                ++cnt;
            }
        }
        if (!p.is_max) acc = acc / (float)(p.count_include_pad ? p.kh * p.kw : max(cnt, 1));
        ST<T>::st(y + i, c < p.C ? acc : 0.f);
    }
}

// bf16 activations with CP % 8 == 0: eight channels per thread (16-byte accesses, 32-bit index arithmetic); per channel the
// same operations in the same order as pool_kernel, so the results are bit-identical.
__global__ void __launch_bounds__(256) pool_bf16x8_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, const PoolP p) {
    pdl_start();
    const int cg = p.CP >> 3;
    const int total = p.N * p.Ho * p.Wo * cg;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int g = i % cg;
        int q = i / cg;
        const int ox = q % p.Wo;
        q /= p.Wo;
        const int oy = q % p.Ho;
        const int n = q / p.Ho;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = p.is_max ? -3.4e38f : 0.f;
        int cnt = 0;
        for (int ky = 0; ky < p.kh; ++ky) {
            const int iy = oy * p.sh - p.pt + ky;
            if ((unsigned)iy >= (unsigned)p.H) continue;
            for (int kx = 0; kx < p.kw; ++kx) {
                const int ix = ox * p.sw - p.pl + kx;
                if ((unsigned)ix >= (unsigned)p.W) continue;
                const uint4 t = *reinterpret_cast<const uint4*>(x + ((size_t)((n * p.H + iy) * p.W + ix) * p.CP + g * 8));
                const uint32_t u[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float v0 = __uint_as_float(u[j] << 16), v1 = __uint_as_float(u[j] & 0xffff0000u);
                    acc[2 * j] = p.is_max ? fmaxf(acc[2 * j], v0) : acc[2 * j] + v0;
                    acc[2 * j + 1] = p.is_max ? fmaxf(acc[2 * j + 1], v1) : acc[2 * j + 1] + v1;
                }
                ++cnt;
            }
        }
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float a0 = acc[2 * j], a1 = acc[2 * j + 1];
            if (!p.is_max) {
                const float d = (float)(p.count_include_pad ? p.kh * p.kw : max(cnt, 1));
                a0 = a0 / d;
                a1 = a1 / d;
            }
            const int c = g * 8 + 2 * j;
            __nv_bfloat162 r = __floats2bfloat162_rn(c < p.C ? a0 : 0.f, c + 1 < p.C ? a1 : 0.f);
            o[j] = *reinterpret_cast<uint32_t*>(&r);
        }
        *reinterpret_cast<uint4*>(y + (size_t)i * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// ----------------------------------------------------------------------------------------
// element-wise: out = epi(a (op) b'), b' = b nearest-upsampled by (fh, fw) and/or broadcast over pixels
// ----------------------------------------------------------------------------------------
struct EltP {
    int N, H, W, CP, C;
    int fh, fw;           // upsample factors applied to b
    int b_bcast;          // 1: b is [N,1,1,CP]
    int op;               // 0 add, 1 mul
    Epilogue epi;
    // lazily applied SE gates (bf16x8 kernel): a' = a * sa[n][c] (+ a when res_a), likewise b; nullptr = plain operand
    const float* sa = nullptr;
    const float* sb = nullptr;
    int res_a = 0, res_b = 0;
};
template <typename T>
__global__ void eltwise_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ y, const EltP p) {
    pdl_start();
    const int cg = p.CP >> 2;
    int64_t total = (int64_t)p.N * p.H * p.W * cg;
    const int Hb = p.H / p.fh, Wb = p.W / p.fw;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int g = (int)(i % cg);
        int64_t q = i / cg;
        int xw = (int)(q % p.W);
        q /= p.W;
        int yh = (int)(q % p.H);
        int n = (int)(q / p.H);
        float va[4], vb[4], o[4];
        ST<T>::ld4(a + i * 4, va);
        int64_t bi = p.b_bcast ? ((int64_t)n * p.CP + g * 4)
                               : ((((int64_t)n * Hb + yh / p.fh) * Wb + xw / p.fw) * p.CP + g * 4);
        ST<T>::ld4(b + bi, vb);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float t = p.op == 0 ? va[j] + vb[j] : va[j] * vb[j];
            o[j] = (g * 4 + j < p.C) ? epi_apply(t, p.epi) : 0.f;
        }
        ST<T>::st4(y + i * 4, o);
    }
}

// bf16 fast path of eltwise_kernel: 8 channels (16 bytes) per thread, 32-bit index arithmetic, a plain add runs as four
// FPN top-down add, y = a' + up(b'): b nearest-upsampled by powers of two, a' / b' optionally carrying a pending SE gate
// (x * s[n][c], + x when res).  One block row per image row: grid (row chunks, H, N), so the only division left per thread
// is the split of its in-row index into (pixel, 16-byte chunk); gates are applied in fp32, one rounding at the end.
struct FpnAddP {
    int W, Hb, Wb, cg, sh, sw;     // cg = CP / 8; b is [N][Hb][Wb][CP], sh / sw = log2 of the upsample factors
    const float* sa;
    const float* sb;
    int res_a, res_b;
};
__global__ void __launch_bounds__(256) fpn_add_bf16_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b,
                                                           bf16* __restrict__ y, const FpnAddP p) {
    pdl_start();
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= (uint32_t)p.W * (uint32_t)p.cg) return;
    const uint32_t xw = j / (uint32_t)p.cg, g = j - xw * (uint32_t)p.cg;
    const uint32_t yh = blockIdx.y, n = blockIdx.z, H = gridDim.y;
    const size_t ia = ((size_t)(n * H + yh) * p.W) * p.cg + j;
    const size_t ib = ((size_t)(n * p.Hb + (yh >> p.sh)) * p.Wb + (xw >> p.sw)) * p.cg + g;
    const uint4 va = reinterpret_cast<const uint4*>(a)[ia];
    const uint4 vb = reinterpret_cast<const uint4*>(b)[ib];
    uint4 o;
    if (!p.sa && !p.sb) {
        const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&va);
        const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&vb);
        __nv_bfloat162* po = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int q = 0; q < 4; ++q) po[q] = __hadd2(pa[q], pb[q]);
    } else {
        const int CP = p.cg * 8;
        float ga[8], gb[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) { ga[q] = 0.f; gb[q] = 0.f; }
        if (p.sa) {
            const float4 s0 = *reinterpret_cast<const float4*>(p.sa + (size_t)n * CP + g * 8);
            const float4 s1 = *reinterpret_cast<const float4*>(p.sa + (size_t)n * CP + g * 8 + 4);
            ga[0] = s0.x; ga[1] = s0.y; ga[2] = s0.z; ga[3] = s0.w; ga[4] = s1.x; ga[5] = s1.y; ga[6] = s1.z; ga[7] = s1.w;
        }
        if (p.sb) {
            const float4 s0 = *reinterpret_cast<const float4*>(p.sb + (size_t)n * CP + g * 8);
            const float4 s1 = *reinterpret_cast<const float4*>(p.sb + (size_t)n * CP + g * 8 + 4);
            gb[0] = s0.x; gb[1] = s0.y; gb[2] = s0.z; gb[3] = s0.w; gb[4] = s1.x; gb[5] = s1.y; gb[6] = s1.z; gb[7] = s1.w;
        }
        const uint32_t* ua = &va.x;
        const uint32_t* ub = &vb.x;
        uint32_t* uo = &o.x;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float a0 = __uint_as_float(ua[q] << 16), a1 = __uint_as_float(ua[q] & 0xffff0000u);
            float b0 = __uint_as_float(ub[q] << 16), b1 = __uint_as_float(ub[q] & 0xffff0000u);
            if (p.sa) {
                const float t0 = a0 * ga[2 * q], t1 = a1 * ga[2 * q + 1];
                a0 = p.res_a ? a0 + t0 : t0;
                a1 = p.res_a ? a1 + t1 : t1;
            }
            if (p.sb) {
                const float t0 = b0 * gb[2 * q], t1 = b1 * gb[2 * q + 1];
                b0 = p.res_b ? b0 + t0 : t0;
                b1 = p.res_b ? b1 + t1 : t1;
            }
            __nv_bfloat162 r = __floats2bfloat162_rn(a0 + b0, a1 + b1);
            uo[q] = *reinterpret_cast<uint32_t*>(&r);
        }
    }
    reinterpret_cast<uint4*>(y)[ia] = o;
}

// packed bf16x2 adds (fp32 inside, one rounding — identical to add-in-fp32-then-round).  Needs CP % 8 == 0, no
// channel broadcast, fewer than 2^31 16-byte chunks.
__global__ void __launch_bounds__(256) eltwise_bf16x8_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b,
                                                             bf16* __restrict__ y, const EltP p, uint32_t total8) {
    pdl_start();
    const uint32_t cg = (uint32_t)p.CP >> 3;
    const uint32_t Hb = p.H / p.fh, Wb = p.W / p.fw;
    const bool same = p.fh == 1 && p.fw == 1;
    const bool gated = p.sa != nullptr || p.sb != nullptr;
    const bool plain_add = p.op == 0 && p.epi.n == 0 && p.C == p.CP && !gated;
    const uint32_t per_img = (uint32_t)p.H * (uint32_t)p.W * cg;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += gridDim.x * blockDim.x) {
        uint32_t bi = i;
        uint32_t g = 0;
        if (!same || !plain_add) {
            g = i % cg;
            if (!same) {
                uint32_t q = i / cg;
                uint32_t xw = q % (uint32_t)p.W;
                q /= (uint32_t)p.W;
                uint32_t yh = q % (uint32_t)p.H;
                uint32_t n = q / (uint32_t)p.H;
                bi = ((n * Hb + yh / (uint32_t)p.fh) * Wb + xw / (uint32_t)p.fw) * cg + g;
            }
        }
        const uint4 va = reinterpret_cast<const uint4*>(a)[i];
        const uint4 vb = reinterpret_cast<const uint4*>(b)[bi];
        uint4 o;
        if (gated) {
            // SE gates applied on the fly: the gated tensors themselves never exist in memory
            const uint32_t n = i / per_img;
            float ga[8], gb[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { ga[j] = 1.f; gb[j] = 1.f; }
            if (p.sa) {
                const float4 s0 = *reinterpret_cast<const float4*>(p.sa + (size_t)n * p.CP + g * 8);
                const float4 s1 = *reinterpret_cast<const float4*>(p.sa + (size_t)n * p.CP + g * 8 + 4);
                const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) ga[j] = sv[j];
            }
            if (p.sb) {
                const float4 s0 = *reinterpret_cast<const float4*>(p.sb + (size_t)n * p.CP + g * 8);
                const float4 s1 = *reinterpret_cast<const float4*>(p.sb + (size_t)n * p.CP + g * 8 + 4);
                const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) gb[j] = sv[j];
            }
            const uint32_t* ua = &va.x;
            const uint32_t* ub = &vb.x;
            uint32_t* uo = &o.x;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float a0 = __uint_as_float(ua[j] << 16), a1 = __uint_as_float(ua[j] & 0xffff0000u);
                float b0 = __uint_as_float(ub[j] << 16), b1 = __uint_as_float(ub[j] & 0xffff0000u);
                if (p.sa) {
                    const float t0 = a0 * ga[2 * j], t1 = a1 * ga[2 * j + 1];
                    a0 = p.res_a ? a0 + t0 : t0;
                    a1 = p.res_a ? a1 + t1 : t1;
                }
                if (p.sb) {
                    const float t0 = b0 * gb[2 * j], t1 = b1 * gb[2 * j + 1];
                    b0 = p.res_b ? b0 + t0 : t0;
                    b1 = p.res_b ? b1 + t1 : t1;
                }
                float t0 = p.op == 0 ? a0 + b0 : a0 * b0, t1 = p.op == 0 ? a1 + b1 : a1 * b1;
                int c = (int)g * 8 + 2 * j;
                t0 = c < p.C ? epi_apply(t0, p.epi) : 0.f;
                t1 = c + 1 < p.C ? epi_apply(t1, p.epi) : 0.f;
                __nv_bfloat162 r = __floats2bfloat162_rn(t0, t1);
                uo[j] = *reinterpret_cast<uint32_t*>(&r);
            }
        } else if (plain_add) {
            const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&va);
            const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&vb);
            __nv_bfloat162* po = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
            for (int j = 0; j < 4; ++j) po[j] = __hadd2(pa[j], pb[j]);
        } else {
            const uint32_t* ua = &va.x;
            const uint32_t* ub = &vb.x;
            uint32_t* uo = &o.x;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float a0 = __uint_as_float(ua[j] << 16), a1 = __uint_as_float(ua[j] & 0xffff0000u);
                float b0 = __uint_as_float(ub[j] << 16), b1 = __uint_as_float(ub[j] & 0xffff0000u);
                float t0 = p.op == 0 ? a0 + b0 : a0 * b0, t1 = p.op == 0 ? a1 + b1 : a1 * b1;
                int c = (int)g * 8 + 2 * j;
                t0 = c < p.C ? epi_apply(t0, p.epi) : 0.f;
                t1 = c + 1 < p.C ? epi_apply(t1, p.epi) : 0.f;
                __nv_bfloat162 r = __floats2bfloat162_rn(t0, t1);
                uo[j] = *reinterpret_cast<uint32_t*>(&r);
            }
        }
        reinterpret_cast<uint4*>(y)[i] = o;
    }
}

// generic scalar version for cpad == 1 tensors
template <typename T>
__global__ void eltwise_scalar_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ y, int64_t total,
                                      int op, const Epilogue epi) {
    pdl_start();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        float va = ST<T>::ld(a + i), vb = ST<T>::ld(b + i);
        ST<T>::st(y + i, epi_apply(op == 0 ? va + vb : va * vb, epi));
    }
}

// y = epi(x) ; optional per-channel vector op first (vec_op: -1 none, 0 add, 1 mul)
template <typename T>
__global__ void unary_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t total, int CP, int C,
                             const float* __restrict__ vec, int vec_op, const Epilogue epi) {
    pdl_start();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % CP);
        float v = ST<T>::ld(x + i);
        if (vec_op == 0) v += vec[c];
        else if (vec_op == 1) v *= vec[c];
        ST<T>::st(y + i, c < C ? epi_apply(v, epi) : 0.f);
    }
}

// channel concat with per-input nearest upsampling: copies one input into its channel slice.
template <typename T>
__global__ void concat_slice_kernel(const T* __restrict__ src, T* __restrict__ dst, int N, int H, int W, int Cs, int CsP,
                                    int CdP, int c_off, int fh, int fw) {
    pdl_start();
    int64_t total = (int64_t)N * H * W * Cs;
    const int Hs = H / fh, Ws = W / fw;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % Cs);
        int64_t q = i / Cs;
        int xw = (int)(q % W);
        q /= W;
        int yh = (int)(q % H);
        int n = (int)(q / H);
        dst[(((int64_t)n * H + yh) * W + xw) * CdP + c_off + c] =
            src[(((int64_t)n * Hs + yh / fh) * Ws + xw / fw) * CsP + c];
    }
}

// All inputs of a channel concat in ONE launch (bf16, every input's channel count a multiple of 8): one thread per
// (pixel, 16-byte chunk of the destination row), so a warp writes whole destination lines instead of every input writing
// its own slice of each row in a separate pass.  Inputs may be nearest-upsampled (fh, fw) and carry a lazily applied SE
// gate (scale[n][c], + x when res); destination pad chunks beyond the last input are zeroed here as well.
constexpr int kCatMaxIn = 8;
constexpr int kCatMaxChunks = 128;               // destination row <= 1024 channels
struct CatP {
    int N, H, W, CdP, nin;
    const bf16* src[kCatMaxIn];
    const float* scale[kCatMaxIn];
    int fh[kCatMaxIn], fw[kCatMaxIn];            // log2 of the nearest-upsample factors (powers of two only)
    int CsP[kCatMaxIn], g0[kCatMaxIn], res[kCatMaxIn];   // g0: first destination chunk of the input
    unsigned char kmap[kCatMaxChunks];           // destination chunk -> input index (255: pad chunk, written as zeros)
};
__global__ void __launch_bounds__(256) concat_multi_bf16_kernel(bf16* __restrict__ dst, const CatP p) {
    pdl_start();
    // grid (row chunks, H, N): one division per thread (in-row index -> pixel, chunk); upsampling by shifts.  The input a
    // chunk comes from is looked up (kmap) and its parameters are read with a run-time index, so the lanes of a warp —
    // which span several inputs — all execute the same instruction stream instead of one branch per input.  The tables
    // are first copied to shared memory: indexed reads of kernel parameters go through the constant cache, which serves
    // one address per cycle — a warp spanning 12 chunks of 4 inputs replayed every one of those loads 4 to 12 times.
    __shared__ unsigned char s_kmap[kCatMaxChunks];
    __shared__ const bf16* s_src[kCatMaxIn];
    __shared__ const float* s_scale[kCatMaxIn];
    __shared__ int s_fh[kCatMaxIn], s_fw[kCatMaxIn], s_csp[kCatMaxIn], s_g0[kCatMaxIn], s_res[kCatMaxIn];
    if (threadIdx.x < kCatMaxChunks) s_kmap[threadIdx.x] = p.kmap[threadIdx.x];
    if (threadIdx.x < kCatMaxIn) {
        const int k = threadIdx.x;
        s_src[k] = p.src[k]; s_scale[k] = p.scale[k]; s_fh[k] = p.fh[k]; s_fw[k] = p.fw[k];
        s_csp[k] = p.CsP[k]; s_g0[k] = p.g0[k]; s_res[k] = p.res[k];
    }
    __syncthreads();
    const uint32_t cg = (uint32_t)p.CdP >> 3;
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= (uint32_t)p.W * cg) return;
    const uint32_t xw = j / cg, g = j - xw * cg;
    const uint32_t yh = blockIdx.y, n = blockIdx.z;
    const int k = s_kmap[g];
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (k != 255) {
        const int fh = s_fh[k], fw = s_fw[k], csp = s_csp[k];
        const uint32_t Hs = (uint32_t)p.H >> fh, Ws = (uint32_t)p.W >> fw;
        const uint32_t c = (g - (uint32_t)s_g0[k]) * 8u;
        v = *reinterpret_cast<const uint4*>(s_src[k] + ((size_t)(n * Hs + (yh >> fh)) * Ws + (xw >> fw)) * csp + c);
        const float* sc = s_scale[k];
        if (sc) {
            const float4 s0 = *reinterpret_cast<const float4*>(sc + (size_t)n * csp + c);
            const float4 s1 = *reinterpret_cast<const float4*>(sc + (size_t)n * csp + c + 4);
            const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
            const float keep = s_res[k] ? 1.f : 0.f;      // x * s + x  ==  x * s + keep * x  (same two roundings)
            uint32_t* u = &v.x;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float a0 = __uint_as_float(u[q] << 16), a1 = __uint_as_float(u[q] & 0xffff0000u);
                const float t0 = a0 * sv[2 * q] + keep * a0, t1 = a1 * sv[2 * q + 1] + keep * a1;
                __nv_bfloat162 rr = __floats2bfloat162_rn(t0, t1);
                u[q] = *reinterpret_cast<uint32_t*>(&rr);
            }
        }
    }
    reinterpret_cast<uint4*>(dst)[((size_t)(n * p.H + yh) * p.W) * cg + j] = v;
}

// 16-byte vector version (channel counts and offset multiples of 16 / sizeof(T))
template <typename T>
__global__ void concat_slice_vec_kernel(const T* __restrict__ src, T* __restrict__ dst, int N, int H, int W, int Cs, int CsP,
                                        int CdP, int c_off, int fh, int fw) {
    pdl_start();
    constexpr int EPV = 16 / (int)sizeof(T);
    const int vc = Cs / EPV;
    int64_t total = (int64_t)N * H * W * vc;
    const int Hs = H / fh, Ws = W / fw;
    if (total < ((int64_t)1 << 31)) {               // 32-bit index arithmetic (64-bit divisions dominate otherwise)
        const uint32_t t32 = (uint32_t)total, uvc = (uint32_t)vc, uW = (uint32_t)W, uH = (uint32_t)H;
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < t32; i += gridDim.x * blockDim.x) {
            uint32_t q = i / uvc;
            int c = (int)(i - q * uvc) * EPV;
            uint32_t r = q / uW;
            int xw = (int)(q - r * uW);
            uint32_t n = r / uH;
            int yh = (int)(r - n * uH);
            uint4 v = *reinterpret_cast<const uint4*>(src + (((int64_t)n * Hs + yh / fh) * Ws + xw / fw) * CsP + c);
            *reinterpret_cast<uint4*>(dst + ((int64_t)q) * CdP + c_off + c) = v;
        }
        return;
    }
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % vc) * EPV;
        int64_t q = i / vc;
        int xw = (int)(q % W);
        q /= W;
        int yh = (int)(q % H);
        int n = (int)(q / H);
        uint4 v = *reinterpret_cast<const uint4*>(src + (((int64_t)n * Hs + yh / fh) * Ws + xw / fw) * CsP + c);
        *reinterpret_cast<uint4*>(dst + (((int64_t)n * H + yh) * W + xw) * CdP + c_off + c) = v;
    }
}

// Dense k x k convolution for tiny input channel counts (the 3-channel stem): one thread per output pixel computes all
// COUT channels; weights fp32 [tap][CinP = 4][COUT] in shared memory.
template <typename T, int COUT>
__global__ void __launch_bounds__(128) conv_stem_kernel(const T* __restrict__ x, const float* __restrict__ w, int w_ld,
                                                        const float* __restrict__ bias, T* __restrict__ y, const ConvP p) {
    pdl_start();
    extern __shared__ float sw[];                 // [taps * 4][COUT]
    const int taps = p.kh * p.kw;
    for (int i = threadIdx.x; i < taps * 4 * COUT; i += blockDim.x) sw[i] = w[(int64_t)(i / COUT) * w_ld + (i % COUT)];
    __syncthreads();
    for (int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; m < p.M; m += (int64_t)gridDim.x * blockDim.x) {
        int ox = (int)(m % p.Wo);
        int64_t q = m / p.Wo;
        int oy = (int)(q % p.Ho);
        int n = (int)(q / p.Ho);
        // output channels in packed fp32x2 pairs (FFMA2), weights fetched as 16-byte broadcast loads
        f32x2_t acc2[COUT / 2];
#pragma unroll
        for (int j = 0; j < COUT / 2; ++j) acc2[j] = pack_f32x2(bias[2 * j], bias[2 * j + 1]);
        for (int ky = 0; ky < p.kh; ++ky) {
            int iy = oy * p.sh - p.pt + ky;
            if ((unsigned)iy >= (unsigned)p.H) continue;
            for (int kx = 0; kx < p.kw; ++kx) {
                int ix = ox * p.sw - p.pl + kx;
                if ((unsigned)ix >= (unsigned)p.W) continue;
                float v[4];
                ST<T>::ld4(x + (((int64_t)n * p.H + iy) * p.W + ix) * 4, v);
                const float* wt = sw + (ky * p.kw + kx) * 4 * COUT;
#pragma unroll
                for (int ci = 0; ci < 3; ++ci) {
                    const f32x2_t vv = pack_f32x2(v[ci], v[ci]);
#pragma unroll
                    for (int j = 0; j < COUT; j += 4) {
                        const float4 w4 = *reinterpret_cast<const float4*>(wt + ci * COUT + j);
                        ffma2(acc2[j / 2], vv, pack_f32x2(w4.x, w4.y));
                        ffma2(acc2[j / 2 + 1], vv, pack_f32x2(w4.z, w4.w));
                    }
                }
            }
        }
        float acc[COUT];
#pragma unroll
        for (int j = 0; j < COUT / 2; ++j) unpack_f32x2(acc2[j], acc[2 * j], acc[2 * j + 1]);
        T* dst = y + m * p.CoutP;
#pragma unroll
        for (int j = 0; j < COUT; j += 4) {
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] = (j + e < p.Cout) ? epi_apply(acc[j + e], p.epi) : 0.f;
            if (j < p.CoutP) ST<T>::st4(dst + j, o);
        }
    }
}

// Pre-processing fused into the stem: u8 BGR page -> (cv2-exact resize) -> normalise -> 3x3 stride-2 pad-1 conv 3 -> 16 ->
// epilogue -> NHWC activations, in one pass (operators.py:54-138,22-31,170-176 + the first Conv of the graph).  The
// normalised fp32 / bf16 input tensor (H*W*4 elements per page, written and re-read by the separate kernels) never exists.
//   * a block owns an 8 x 32 tile of stem outputs: its 17 x 65 input pixels are resized / normalised once into shared
//     memory (planar fp32, even and odd columns in separate planes so the stride-2 reads of a warp are conflict free);
//   * the three rounded fp32 operations of NormalizeImage depend only on (channel, u8 value): a 3 x 256 table per block
//     (the true division is done 768 times per block instead of three times per pixel); in bf16 mode the table holds the
//     value already rounded to bf16, exactly what resize_norm_kernel would have stored;
//   * the 432 weights and 16 biases are kernel parameters (constant bank): every FFMA takes its weight as a constant
//     operand, there is no weight load at all in the loop;
//   * accumulation order (bias, then ky, kx, ci) is conv_stem_kernel's, so both paths give bit-identical results (taps in
//     the zero padding add v = 0 here and are skipped there).
struct StemW {
    float w[27][16];              // [(ky*3 + kx)*3 + ci][co]
    float b[16];
};
constexpr int kPsTH = 8, kPsTW = 32, kPsRH = 2 * kPsTH + 1, kPsRW = 2 * kPsTW + 1, kPsPitch = 48;

template <typename T>
__global__ void __launch_bounds__(256) pre_stem_kernel(const ResizeJob* __restrict__ jobs, int IH, int IW, const NormP nrm,
                                                       const __grid_constant__ StemW W, T* __restrict__ y, const ConvP p,
                                                       int tiles_x, int tiles_y, int tiles) {
    pdl_start();
    __shared__ float lut[3][256];
    __shared__ float tile[3][kPsRH][2][kPsPitch];
    const int tid = threadIdx.x;
    for (int i = tid; i < 768; i += 256) {
        const int c = i >> 8, v = i & 255;
        float f = norm_px(v, c, nrm);
        if (sizeof(T) == 2) f = __bfloat162float(__float2bfloat16_rn(f));
        lut[c][v] = f;
    }
    const int lx = tid & 31, ly = tid >> 5;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int tx = t % tiles_x;
        const int r = t / tiles_x;
        const int ty = r % tiles_y, n = r / tiles_y;
        const ResizeJob j = jobs[n];
        const int iy0 = 2 * ty * kPsTH - 1, ix0 = 2 * tx * kPsTW - 1;
        __syncthreads();                                    // table ready / previous tile consumed
        for (int q = tid; q < kPsRH * kPsRW; q += 256) {
            const int ry = q / kPsRW, rx = q - ry * kPsRW;
            const int iy = iy0 + ry, ix = ix0 + rx;
            float v0 = 0.f, v1 = 0.f, v2 = 0.f;
            if ((unsigned)iy < (unsigned)IH && (unsigned)ix < (unsigned)IW && ix < j.dw && iy < j.dh) {
                int px[3];
                resize_px(j.src, j.dh, j.dw, j.scale_y, j.scale_x, iy, ix, px);
                v0 = lut[0][px[0]]; v1 = lut[1][px[1]]; v2 = lut[2][px[2]];
            }
            tile[0][ry][rx & 1][rx >> 1] = v0;
            tile[1][ry][rx & 1][rx >> 1] = v1;
            tile[2][ry][rx & 1][rx >> 1] = v2;
        }
        __syncthreads();
        const int oy = ty * kPsTH + ly, ox = tx * kPsTW + lx;
        float acc[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[c] = W.b[c];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                for (int ci = 0; ci < 3; ++ci) {
                    const float v = tile[ci][2 * ly + ky][kx & 1][lx + (kx >> 1)];
#pragma unroll
                    for (int c = 0; c < 16; ++c) acc[c] = fmaf(v, W.w[(ky * 3 + kx) * 3 + ci][c], acc[c]);
                }
        if (oy < p.Ho && ox < p.Wo) {
            T* dst = y + (((int64_t)n * p.Ho + oy) * p.Wo + ox) * p.CoutP;
            float o[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) o[c] = c < p.Cout ? epi_apply(acc[c], p.epi) : 0.f;
            if (sizeof(T) == 2 && p.CoutP == 16) {
                uint32_t w[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    __nv_bfloat162 h = __floats2bfloat162_rn(o[2 * c], o[2 * c + 1]);
                    w[c] = *reinterpret_cast<uint32_t*>(&h);
                }
                uint4* d4 = reinterpret_cast<uint4*>(dst);
                d4[0] = make_uint4(w[0], w[1], w[2], w[3]);
                d4[1] = make_uint4(w[4], w[5], w[6], w[7]);
            } else {
#pragma unroll
                for (int c = 0; c < 16; c += 4)
                    if (c < p.CoutP) ST<T>::st4(dst + c, o + c);
            }
        }
    }
}

// The same fused pre-processing + stem on the tensor cores (bf16 plans).  On CUDA cores the 432 FMAs per output pixel
// make the kernel issue-bound at 3x its memory time (ncu: 68 M warp instructions, FMA pipe 34 %, 0.165 ms for 16 pages);
// as a GEMM the stem is M = pixels, N = 16, K = 9 taps x 4 channels (BGR0) = 36 -> 48: three k-steps of
// mma.sync.m16n8k16 (bf16 x bf16 -> fp32), two n-tiles.  The normalised input is bf16 already (that is what the two-kernel
// path stores), the fp32 weights are split into bf16 hi + lo parts (two MMAs), so the products carry 16 mantissa bits —
// the result differs from the FFMA kernel only by fp32 summation order.  A warp owns one row of 32 output pixels (two
// m16 tiles); the A fragments are gathered straight from the staged tile: pixel = (2*oy + ky, 2*ox + kx), one 8-byte
// BGR0 group per pixel, the fragment's k pairs are the channel pairs of consecutive taps (conflict-free 4-byte loads).
// B fragments (per lane: 3 k-steps x 2 n-tiles x 2 registers, hi and lo) come from a table built when the plan is built.
struct StemMmaTab {
    uint32_t b[2][3][2][32][2];       // [hi, lo][k-step][n-tile][lane][b0, b1]
    float bias[16];
};

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                               uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

constexpr int kPmPitch = kPsRW * 2;   // 32-bit words per tile row (two per pixel)

template <int ACT>
__global__ void __launch_bounds__(256, 3) pre_stem_mma_kernel(const ResizeJob* __restrict__ jobs, int IH, int IW, const NormP nrm,
                                                           const StemMmaTab* __restrict__ tab, bf16* __restrict__ y,
                                                           const ConvP p, int tiles_x, int tiles_y, int tiles) {
    pdl_start();
    __shared__ uint16_t lut[3][256];
    __shared__ __align__(16) uint32_t tile[kPsRH][kPmPitch];
    const int tid = threadIdx.x, lane = tid & 31, ly = tid >> 5, g = lane >> 2, tig = lane & 3;
    for (int i = tid; i < 768; i += 256) {
        const int c = i >> 8, v = i & 255;
        lut[c][v] = __bfloat16_as_ushort(__float2bfloat16_rn(norm_px(v, c, nrm)));
    }
    uint32_t bh[3][2][2], bl[3][2][2];
#pragma unroll
    for (int s = 0; s < 3; ++s)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                bh[s][j][r] = tab->b[0][s][j][lane][r];
                bl[s][j][r] = tab->b[1][s][j][lane][r];
            }
    float bias[2][2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        bias[j][0] = tab->bias[8 * j + 2 * tig];
        bias[j][1] = tab->bias[8 * j + 2 * tig + 1];
    }
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int tx = t % tiles_x;
        const int r = t / tiles_x;
        const int ty = r % tiles_y, n = r / tiles_y;
        const ResizeJob& j = jobs[n];
        const int iy0 = 2 * ty * kPsTH - 1, ix0 = 2 * tx * kPsTW - 1;
        const int jdh = j.dh, jdw = j.dw;
        // the common case (C2, every page that already has its network size): no resampling, no virtual padding
        const bool ident = jdh == j.src.ph && jdw == j.src.pw && j.src.h == j.src.ph && j.src.w == j.src.pw && !j.src.rot180;
        __syncthreads();                                    // table ready / previous tile consumed
        if (ident) {
            const uint8_t* __restrict__ sp = j.src.p;
            const int stride = j.src.stride;
            const int ylim = min(IH, jdh), xlim = min(IW, jdw);
            // all byte loads of this thread's (up to 5) pixels are issued before any is used: the kernel is bound by the
            // bytes it keeps in flight, not by instructions
            constexpr int kIt = (kPsRH * kPsRW + 255) / 256;
            uint32_t b0[kIt], b1[kIt], b2[kIt];
#pragma unroll
            for (int k = 0; k < kIt; ++k) {
                const int q = tid + 256 * k;
                const int ry = q / kPsRW, rx = q - ry * kPsRW;
                const int iy = iy0 + ry, ix = ix0 + rx;
                b0[k] = 0xffffffffu;
                if (q < kPsRH * kPsRW && (unsigned)iy < (unsigned)ylim && (unsigned)ix < (unsigned)xlim) {
                    const uint8_t* s = sp + (size_t)iy * stride + ix * 3;
                    b0[k] = s[0]; b1[k] = s[1]; b2[k] = s[2];
                }
            }
#pragma unroll
            for (int k = 0; k < kIt; ++k) {
                const int q = tid + 256 * k;
                if (q < kPsRH * kPsRW) {
                    const int ry = q / kPsRW, rx = q - ry * kPsRW;
                    uint2 v = make_uint2(0u, 0u);
                    if (b0[k] != 0xffffffffu) {
                        v.x = (uint32_t)lut[0][b0[k]] | ((uint32_t)lut[1][b1[k]] << 16);
                        v.y = (uint32_t)lut[2][b2[k]];
                    }
                    *reinterpret_cast<uint2*>(&tile[ry][2 * rx]) = v;
                }
            }
        } else {
            const ResizeJob jj = j;
            for (int q = tid; q < kPsRH * kPsRW; q += 256) {
                const int ry = q / kPsRW, rx = q - ry * kPsRW;
                const int iy = iy0 + ry, ix = ix0 + rx;
                uint2 v = make_uint2(0u, 0u);
                if ((unsigned)iy < (unsigned)IH && (unsigned)ix < (unsigned)IW && ix < jdw && iy < jdh) {
                    int px[3];
                    resize_px(jj.src, jdh, jdw, jj.scale_y, jj.scale_x, iy, ix, px);
                    v.x = (uint32_t)lut[0][px[0]] | ((uint32_t)lut[1][px[1]] << 16);
                    v.y = (uint32_t)lut[2][px[2]];
                }
                *reinterpret_cast<uint2*>(&tile[ry][2 * rx]) = v;
            }
        }
        __syncthreads();
        const int oy = ty * kPsTH + ly;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            float acc[2][4];
#pragma unroll
            for (int jn = 0; jn < 2; ++jn) { acc[jn][0] = acc[jn][2] = bias[jn][0]; acc[jn][1] = acc[jn][3] = bias[jn][1]; }
            // word index of (tap, pixel row `g` of this m-tile): ((2 ly + ky) * 65 + 2 x + kx) * 2 + (tig & 1)
            const int xw = 2 * (mt * 16 + g);               // input column of kx = 0 for output pixel g
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                uint32_t a[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                for (int half = 0; half < 2; ++half) {      // k 0..7 (a0, a1) and k 8..15 (a2, a3) of this k-step
                    const int tap = 4 * s + 2 * half + (tig >> 1);
                    if (4 * s + 2 * half <= 8) {            // compile time: the last k-step has only tap 8
                        const int ky = tap / 3, kx = tap - 3 * ky;
                        if (tap <= 8) {
                            const uint32_t* row = &tile[2 * ly + ky][0];
                            const int w0 = (xw + kx) * 2 + (tig & 1);
                            a[2 * half] = row[w0];
                            a[2 * half + 1] = row[w0 + 32];   // output pixel g + 8: 16 input pixels = 32 words further
                        }
                    }
                }
#pragma unroll
                for (int jn = 0; jn < 2; ++jn) {
                    mma_bf16_16816(acc[jn], a[0], a[1], a[2], a[3], bh[s][jn][0], bh[s][jn][1]);
                    mma_bf16_16816(acc[jn], a[0], a[1], a[2], a[3], bl[s][jn][0], bl[s][jn][1]);
                }
            }
            // epilogue: acc[jn][0..1] = pixel g, channels 8 jn + 2 tig (+1); acc[jn][2..3] = pixel g + 8
            uint32_t item[4];                                // [pixel g: words tig, 4 + tig] [pixel g + 8: words tig, 4 + tig]
#pragma unroll
            for (int jn = 0; jn < 2; ++jn)
#pragma unroll
                for (int hrow = 0; hrow < 2; ++hrow) {
                    const int c0 = 8 * jn + 2 * tig;
                    const float v0 = c0 < p.Cout ? epi_fast<ACT>(acc[jn][2 * hrow], p.epi) : 0.f;
                    const float v1 = c0 + 1 < p.Cout ? epi_fast<ACT>(acc[jn][2 * hrow + 1], p.epi) : 0.f;
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(v0, v1);
                    item[2 * hrow + jn] = *reinterpret_cast<uint32_t*>(&h2);
                }
            // 4 x 4 transpose inside the quad: lane q ends up with words 4 (q & 1) .. 4 (q & 1) + 3 of pixel g + 8 (q >> 1)
            uint32_t out[4];
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                const int pub = (tig - rr) & 3;              // the item this lane publishes in round rr
                const uint32_t v = pub == 0 ? item[0] : pub == 1 ? item[1] : pub == 2 ? item[2] : item[3];
                const uint32_t got = __shfl_sync(0xffffffffu, v, (lane & ~3) | ((tig + rr) & 3));
                const int slot = (tig + rr) & 3;             // it came from lane `slot` of the quad = word index within the group
                if (slot == 0) out[0] = got; else if (slot == 1) out[1] = got; else if (slot == 2) out[2] = got; else out[3] = got;
            }
            const int ox = tx * kPsTW + mt * 16 + g + 8 * (tig >> 1);
            if (oy < p.Ho && ox < p.Wo) {
                bf16* dst = y + (((int64_t)n * p.Ho + oy) * p.Wo + ox) * 16 + 8 * (tig & 1);
                *reinterpret_cast<uint4*>(dst) = make_uint4(out[0], out[1], out[2], out[3]);
            }
        }
    }
}

// ConvTranspose 2x2 stride 2 with a tiny output channel count (the DB head's last layer, Cout = 1): one thread per
// input pixel, weights fp32 [tap][CinP] per output channel in shared memory, output cpad == Cout (1 or 4).
template <typename T>
__global__ void __launch_bounds__(256) convt_small_kernel(const T* __restrict__ x, const float* __restrict__ w, int w_ld,
                                                          const float* __restrict__ bias, T* __restrict__ y, const ConvP p) {
    pdl_start();
    extern __shared__ float sw[];                 // [tap][co][CinP]
    const int CoP = p.CoutP;                      // 1 or 4
    for (int i = threadIdx.x; i < 4 * p.Cout * p.CinP; i += blockDim.x) {
        int ci = i % p.CinP, r = i / p.CinP, co = r % p.Cout, tap = r / p.Cout;
        sw[i] = w[(int64_t)ci * w_ld + tap * CoP + co];
    }
    __syncthreads();
    for (int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; m < p.M; m += (int64_t)gridDim.x * blockDim.x) {
        int ox = (int)(m % p.Wo);
        int64_t q = m / p.Wo;
        int oy = (int)(q % p.Ho);
        int n = (int)(q / p.Ho);
        const T* src = x + m * p.CinP;
        float acc[4][4];
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int co = 0; co < 4; ++co) acc[t][co] = co < p.Cout ? bias[co] : 0.f;
        for (int c = 0; c < p.CinP; c += 4) {
            float v[4];
            ST<T>::ld4(src + c, v);
#pragma unroll
            for (int t = 0; t < 4; ++t)
                for (int co = 0; co < p.Cout; ++co) {
                    const float* wt = sw + ((t * p.Cout + co) * p.CinP + c);
                    acc[t][co] = fmaf(v[0], wt[0], fmaf(v[1], wt[1], fmaf(v[2], wt[2], fmaf(v[3], wt[3], acc[t][co]))));
                }
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            int dy = t >> 1, dx = t & 1;
            T* dst = y + ((((int64_t)n * (2 * p.Ho) + 2 * oy + dy) * (2 * p.Wo)) + 2 * ox + dx) * CoP;
            for (int co = 0; co < CoP; ++co) ST<T>::st(dst + co, co < p.Cout ? epi_apply(acc[t][co], p.epi) : 0.f);
        }
    }
}

template <typename T>
__global__ void fill_zero_kernel(T* __restrict__ p, int64_t total) {
    pdl_start();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        ST<T>::st(p + i, 0.f);
}

// ----------------------------------------------------------------------------------------
// LayerNorm over channels: one warp per row
// ----------------------------------------------------------------------------------------
template <typename T>
__global__ void layernorm_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, int64_t rows, int C, int CP, float eps) {
    pdl_start();
    int warp = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
    int lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const T* xr = x + (int64_t)warp * CP;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += ST<T>::ld(xr + c);
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    float mean = s / (float)C;
    float v = 0.f;
    for (int c = lane; c < C; c += 32) {
        float d = ST<T>::ld(xr + c) - mean;
        v = fmaf(d, d, v);
    }
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    float sd = sqrtf(v / (float)C + eps);
    T* yr = y + (int64_t)warp * CP;
    for (int c = lane; c < CP; c += 32) {
        float o = 0.f;
        if (c < C) o = (ST<T>::ld(xr + c) - mean) / sd * gamma[c] + beta[c];
        ST<T>::st(yr + c, o);
    }
}

// ----------------------------------------------------------------------------------------
// SVTR global-mixer attention.  qkv rows: [3][heads][hd]; one thread per query, K/V of the
// (batch, head) pair staged in shared memory.
// ----------------------------------------------------------------------------------------
template <typename T, int HD>
__global__ void attention_kernel(const T* __restrict__ qkv, T* __restrict__ out, int Tn, int heads, int CPin, int CPout,
                                 float scale) {
    pdl_start();
    extern __shared__ float sm[];                   // K[Tn][HD] V[Tn][HD]
    float* Ks = sm;
    float* Vs = sm + (size_t)Tn * HD;
    const int b = blockIdx.x / heads, h = blockIdx.x % heads;
    const int dim = heads * HD;
    const T* base = qkv + (int64_t)b * Tn * CPin;
    for (int i = threadIdx.x; i < Tn * HD; i += blockDim.x) {
        int t = i / HD, d = i % HD;
        Ks[i] = ST<T>::ld(base + (int64_t)t * CPin + dim + h * HD + d);
        Vs[i] = ST<T>::ld(base + (int64_t)t * CPin + 2 * dim + h * HD + d);
    }
    __syncthreads();
    for (int t = blockIdx.y * blockDim.x + threadIdx.x; t < Tn; t += gridDim.y * blockDim.x) {
        float q[HD];
#pragma unroll
        for (int d = 0; d < HD; ++d) q[d] = ST<T>::ld(base + (int64_t)t * CPin + h * HD + d) * scale;
        float mx = -3.4e38f;
        for (int j = 0; j < Tn; ++j) {
            float s = 0.f;
#pragma unroll
            for (int d = 0; d < HD; ++d) s = fmaf(q[d], Ks[j * HD + d], s);
            mx = fmaxf(mx, s);
        }
        float acc[HD];
#pragma unroll
        for (int d = 0; d < HD; ++d) acc[d] = 0.f;
        float den = 0.f;
        for (int j = 0; j < Tn; ++j) {
            float s = 0.f;
#pragma unroll
            for (int d = 0; d < HD; ++d) s = fmaf(q[d], Ks[j * HD + d], s);
            float e = expf(s - mx);
            den += e;
#pragma unroll
            for (int d = 0; d < HD; ++d) acc[d] = fmaf(e, Vs[j * HD + d], acc[d]);
        }
        float inv = 1.f / den;
        T* o = out + ((int64_t)b * Tn + t) * CPout + h * HD;
#pragma unroll
        for (int d = 0; d < HD; ++d) ST<T>::st(o + d, acc[d] * inv);
    }
}

// zero the pad lanes [C, CP) of every row (used after attention, whose threads write only real channels)
template <typename T>
__global__ void zero_pad_lanes_kernel(T* __restrict__ x, int64_t rows, int C, int CP) {
    pdl_start();
    int padn = CP - C;
    int64_t total = rows * padn;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        ST<T>::st(x + (i / padn) * CP + C + (i % padn), 0.f);
}

// ----------------------------------------------------------------------------------------
// softmax over channels, one block per row.  Writes either T NHWC(CP) or dense fp32 (ld = C).
// ----------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) softmax_kernel(const T* __restrict__ x, T* __restrict__ y,
                                                       float* __restrict__ y_dense, int C, int CP) {
    pdl_start();
    __shared__ float red[32];
    const int64_t row = blockIdx.x;
    const T* xr = x + row * CP;
    float mx = -3.4e38f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) mx = fmaxf(mx, ST<T>::ld(xr + c));
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : -3.4e38f;
        for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
        if (threadIdx.x == 0) red[0] = v;
    }
    __syncthreads();
    mx = red[0];
    __syncthreads();
    float s = 0.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) s += expf(ST<T>::ld(xr + c) - mx);
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) red[0] = v;
    }
    __syncthreads();
    float inv = 1.f / red[0];
    if (y_dense) {
        float* yr = y_dense + row * C;
        for (int c = threadIdx.x; c < C; c += blockDim.x) yr[c] = expf(ST<T>::ld(xr + c) - mx) * inv;
    } else {
        T* yr = y + row * CP;
        for (int c = threadIdx.x; c < CP; c += blockDim.x)
            ST<T>::st(yr + c, c < C ? expf(ST<T>::ld(xr + c) - mx) * inv : 0.f);
    }
}

}  // namespace b200ocr

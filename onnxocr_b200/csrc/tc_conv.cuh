// Dense convolution / linear layers on the 5th-generation tensor cores (sm_100a only).
//
// Implicit GEMM: D[M = 128 output pixels][N = BN output channels] += A[M][K] * B[N][K]^T, K = taps x Cin, bf16 in,
// fp32 accumulation in TMEM.
//   * A (activations, NHWC bf16) is never materialised as im2col: for every (tap, 64-channel chunk) the TMA engine
//     loads a {64 ch, TW, TH} box at the tap-shifted pixel coordinates; out-of-image pixels and channels beyond Cin are
//     zero-filled by the TMA unit, which is exactly the convolution's zero padding.
//   * B (weights, bf16, K-major [Cout][taps*chunks*64]) comes through a 2-D tensor map.
//   * both land in 128-byte-swizzled shared-memory tiles that tcgen05.mma reads through shared-memory descriptors.
//   * warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM allocation), warps 2..5 = epilogue
//     (tcgen05.ld -> bias -> activation chain -> bf16 -> NHWC store); 4-stage mbarrier ring between them.
#pragma once
#include <cuda.h>

#include <algorithm>

#include "common.cuh"

namespace b200ocr {
namespace tc {

constexpr int BM = 128;          // UMMA M
constexpr int BK = 64;           // bf16 elements per K block (= one 128-byte swizzle row)
constexpr int STAGES = 4;
constexpr int kThreads = 192;    // 6 warps

struct Params {
    int mode;                    // 0: rows are flattened pixels (1x1 conv / linear); 1: 4-D tiles (k x k conv, stride 1)
    int M;                       // total output pixels
    int N, H, W;                 // mode 1: batch and spatial size (output == input size, stride 1)
    int TW, TH, tiles_x, tiles_y;
    int taps, kw, pad_t, pad_l;
    int kchunks;                 // ceil(CinP / 64)
    int cinp;                    // input channels as stored (GEMM K per tap); K slices of 16 beyond it are all zero and skipped
    int Cout, CoutP;             // GEMM columns that carry data / padded (for convT: 4 * out channels padded)
    int convT;                   // 1: column = tap * ctC + co, stored pixel-shuffled to (2y + dy, 2x + dx)
    int ctC, ctCout;             // convT: padded / real output channels
    Epilogue epi;
    // convT -> convT(2x2, stride 2, <= 1 output channel) -> activation fused into the epilogue (the DB head's tail):
    // ct2_w = [4 taps][ctC] second-stage weights followed by its bias; ct2_out = dense fp32 [N][4H][4W]
    int ws_bn, ws_stages;        // weight-stationary kernel: output-channel tile width and activation ring depth
    int ws_halo, ws_stage_bytes; // halo mode (k x k): bytes per ring stage; kh for the halo height
    int kh;
    // SE producer: per (image, channel) sums of the stored outputs, accumulated by the epilogue (common.cuh gap_fix_add).
    // gap_rows = GEMM rows per image (mode 0 is then tiled per image so that no tile straddles two images),
    // gap_mod = channels per pixel (a pixel-packed GEMM row carries several pixels: column % gap_mod = channel)
    unsigned long long* gap;
    int gap_rows, gap_mod;
    int epi_variant;             // tools/ubench A/B: bit 0 = 16-column TMEM loads only, bit 1 = shuffle-based store loop
    unsigned long long* dbg;     // tools/ubench only: per-CTA cycle counters of the MMA / TMA threads (nullptr in the engine)
    int ct2;
    const float* ct2_w;
    float* ct2_out;
    Epilogue epi2;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Warp-role dispatch must be visibly warp-uniform and the single issuing lane must come from elect.sync: with a plain
// `if (lane == 0)` ptxas cannot prove that one thread is active and wraps EVERY uniform-datapath instruction (UTCHMMA,
// UTMALDG, UTCBAR) in an ELECT / PLOP3 / BRA.U.ANY loop over the active lanes — measured (tools/ubench/mma_rate.cu): the
// bare tensor pipe takes 40 cycles for an M128 N32 K16 MMA, the wrapped issue loop of round 1 ran at ~107.
__device__ __forceinline__ int warp_idx_uniform() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred = 0, laneid = 0;
    asm volatile(
        "{\n"
        ".reg .b32 %%rx;\n"
        ".reg .pred %%px;\n"
        "elect.sync %%rx|%%px, %2;\n"
        "@%%px mov.s32 %1, 1;\n"
        "mov.s32 %0, %%rx;\n"
        "}\n"
        : "+r"(laneid), "+r"(pred)
        : "r"(0xFFFFFFFFu));
    return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// K-major, 128-byte swizzle: 8-row groups are 1024 bytes apart (SBO), LBO unused; descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);          // start address
    // leading byte offset (bits 16..29) stays 0: a 64-element K block is a single swizzle atom along K
    d |= (uint64_t)(1024 >> 4) << 32;                      // stride byte offset
    d |= (uint64_t)1 << 46;                                // descriptor version
    d |= (uint64_t)2 << 61;                                // SWIZZLE_128B
    return d;
}

// General form: 8-row core groups `sbo_bytes` apart and a start address that is only 128-byte aligned — the tap-shifted
// views of a halo tile.  Measured on B200: the 128-byte swizzle is applied to the absolute shared-memory address bits
// (exactly as the TMA unit wrote the tile), so the descriptor's "matrix base offset" field stays 0 even though the start
// is not aligned to the 1024-byte swizzle repeat; setting it to (start >> 7) & 7 gives wrong operands.
__device__ __forceinline__ uint64_t make_smem_desc_sbo(uint32_t smem_addr, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t make_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}

constexpr int CT2_MAX_C = 32;     // fused second transposed conv: 4 * ctC columns must fit one 128-column tile
template <int BN>
constexpr size_t smem_bytes() {
    return (size_t)STAGES * (BM * BK * 2 + BN * BK * 2) + 1024 /*alignment slack*/ + 256 /*barriers*/ +
           (4 * CT2_MAX_C + 4) * 4 /*second-stage convT weights*/;
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// One accumulator row (this thread's pixel) -> bias, activation chain, bf16, 16-byte NHWC stores.
template <int BN, int ACT>
__device__ __forceinline__ void epilogue_row(uint32_t trow, const float* __restrict__ bias, bf16* __restrict__ dst, bool valid,
                                             int n0, const Params& p, int64_t ct_base = 0) {
#pragma unroll 1
    for (int c = 0; c < BN; c += 16) {
        uint32_t v[16];
        tmem_ld8(trow + (uint32_t)c, v);
        tmem_ld8(trow + (uint32_t)c + 8, v + 8);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int col = n0 + c + 8 * h;
            if (valid && col < p.CoutP) {
                float4 b0 = *reinterpret_cast<const float4*>(bias + col);
                float4 b1 = *reinterpret_cast<const float4*>(bias + col + 4);
                float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                float o[8];
                int64_t doff = col;
                int lim = p.Cout - col;
                if (p.convT) {                            // 8-column groups never straddle taps (ctC % 8 == 0)
                    int tap = col / p.ctC, co = col - tap * p.ctC;
                    doff = ct_base + ((int64_t)(tap >> 1) * (2 * p.W) + (tap & 1)) * p.ctC + co;
                    lim = p.ctCout - co;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float f = epi_fast<ACT>(__uint_as_float(v[8 * h + j]) + bb[j], p.epi);
                    o[j] = (j < lim) ? f : 0.f;
                }
                uint4 pk;
                __nv_bfloat162 t0 = __floats2bfloat162_rn(o[0], o[1]);
                __nv_bfloat162 t1 = __floats2bfloat162_rn(o[2], o[3]);
                __nv_bfloat162 t2 = __floats2bfloat162_rn(o[4], o[5]);
                __nv_bfloat162 t3 = __floats2bfloat162_rn(o[6], o[7]);
                pk.x = *reinterpret_cast<uint32_t*>(&t0);
                pk.y = *reinterpret_cast<uint32_t*>(&t1);
                pk.z = *reinterpret_cast<uint32_t*>(&t2);
                pk.w = *reinterpret_cast<uint32_t*>(&t3);
                *reinterpret_cast<uint4*>(dst + doff) = pk;
            }
        }
    }
}

// DB head tail: this thread's accumulator row holds the first transposed conv's 4 taps x ctC channels for input pixel
// (n, yy, xx).  Bias + activation give the 2x2 block of mid-resolution pixels; each of them goes through the second
// transposed conv (ctC -> 1 channel, 2x2) and its activation, yielding a 4x4 block of the final map, stored as four
// 16-byte rows of the dense fp32 output.  The mid-resolution tensor never reaches HBM.
template <int ACT>
__device__ __forceinline__ void epilogue_row_ct2(uint32_t trow, const float* __restrict__ bias, const float* __restrict__ w2s,
                                                 bool valid, int m, const Params& p) {
    float res[4][4];
    const float b2 = w2s[4 * p.ctC];
#pragma unroll
    for (int tap = 0; tap < 4; ++tap) {
        float a[4] = {b2, b2, b2, b2};
#pragma unroll 1
        for (int cc = 0; cc < p.ctC; cc += 8) {
            uint32_t v[8];
            tmem_ld8(trow + (uint32_t)(tap * p.ctC + cc), v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const float* bp = bias + tap * p.ctC + cc;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float f = (cc + j < p.ctCout) ? epi_fast<ACT>(__uint_as_float(v[j]) + bp[j], p.epi) : 0.f;
#pragma unroll
                for (int t2 = 0; t2 < 4; ++t2) a[t2] = fmaf(f, w2s[t2 * p.ctC + cc + j], a[t2]);
            }
        }
#pragma unroll
        for (int t2 = 0; t2 < 4; ++t2) res[tap][t2] = epi_apply(a[t2], p.epi2);
    }
    if (!valid) return;
    int xx = m % p.W;
    int r = m / p.W;
    int yy = r % p.H;
    int nn = r / p.H;
    const int64_t OW = 4 * (int64_t)p.W;
    float* dst = p.ct2_out + ((int64_t)nn * (4 * p.H) + 4 * yy) * OW + 4 * xx;
#pragma unroll
    for (int r4 = 0; r4 < 4; ++r4) {
        const int dy = r4 >> 1, ey = r4 & 1;
        float4 o = make_float4(res[dy * 2][ey * 2], res[dy * 2][ey * 2 + 1], res[dy * 2 + 1][ey * 2], res[dy * 2 + 1][ey * 2 + 1]);
        *reinterpret_cast<float4*>(dst + r4 * OW) = o;
    }
}

// Persistent kernel: each CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...; the TMA ring runs ahead across tile
// boundaries and two TMEM accumulators let the MMA of tile t+1 overlap the epilogue of tile t.
template <int BN>
__global__ void __launch_bounds__(kThreads) conv_tc_kernel(const __grid_constant__ CUtensorMap map_a,
                                                           const __grid_constant__ CUtensorMap map_b,
                                                           const float* __restrict__ bias, bf16* __restrict__ y, const Params p) {
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_base = base, b_base = base + STAGES * A_BYTES;
    const uint32_t bar_base = b_base + STAGES * B_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
    auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
    uint32_t* tmem_slot_ptr = (uint32_t*)(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    float* w2s = (float*)(smem_raw + (bar_base + 256u - smem_u32(smem_raw)));
    if (p.ct2)
        for (int i = threadIdx.x; i < 4 * p.ctC + 1; i += kThreads) w2s[i] = p.ct2_w[i];

    const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
    const int nkb = p.taps * p.kchunks;
    const int m_tiles = p.mode == 0 ? (p.M + BM - 1) / BM : p.tiles_x * p.tiles_y * p.N;
    const int n_tiles = (p.CoutP + BN - 1) / BN;
    const int total_tiles = m_tiles * n_tiles;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tmem_full_bar(a), 1);
            mbar_init(tmem_empty_bar(a), 4);                // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)(2 * BN)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_wait();                                             // activations of the previous kernel are complete and visible

    // tile index -> origin.  n tiles vary fastest so that CTAs working on the same pixels run together (A stays in L2).
    auto tile_origin = [&](int tile, int& n0, int& m0, int& img, int& y0, int& x0) {
        int nt = tile % n_tiles, mt = tile / n_tiles;
        n0 = nt * BN;
        m0 = mt * BM;
        img = 0; y0 = 0; x0 = 0;
        if (p.mode == 1) {
            int tx = mt % p.tiles_x;
            int r = mt / p.tiles_x;
            int ty = r % p.tiles_y;
            img = r / p.tiles_y;
            y0 = ty * p.TH;
            x0 = tx * p.TW;
        }
    };

    if (warp == 0) {
        // ---------------------------------------------------------------- TMA producer
        if (elect_one_sync()) {
            uint32_t it = 0;                                // running k-block counter across tiles
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                int n0, m0, img, y0, x0;
                tile_origin(tile, n0, m0, img, y0, x0);
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    int s = it % STAGES;
                    uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(empty_bar(s), ph ^ 1);
                    mbar_expect_tx(full_bar(s), A_BYTES + B_BYTES);
                    int tap = kb / p.kchunks, chunk = kb - tap * p.kchunks;
                    if (p.mode == 0) {
                        tma_load_2d(a_base + s * A_BYTES, &map_a, full_bar(s), chunk * BK, m0);
                    } else {
                        int ky = tap / p.kw, kx = tap - ky * p.kw;
                        tma_load_4d(a_base + s * A_BYTES, &map_a, full_bar(s), chunk * BK, x0 + kx - p.pad_l, y0 + ky - p.pad_t, img);
                    }
                    tma_load_2d(b_base + s * B_BYTES, &map_b, full_bar(s), kb * BK, n0);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ---------------------------------------------------------------- MMA issuer
        if (elect_one_sync()) {
            constexpr uint32_t idesc = make_idesc(BN);
            uint32_t it = 0, t = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++t) {
                int acc = t & 1;
                uint32_t aph = (t >> 1) & 1;
                mbar_wait(tmem_empty_bar(acc), aph ^ 1);   // epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    int s = it % STAGES;
                    uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(full_bar(s), ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    uint64_t adesc = make_smem_desc(a_base + s * A_BYTES);
                    uint64_t bdesc = make_smem_desc(b_base + s * B_BYTES);
                    const int chunk = kb % p.kchunks;
                    const int ks = min(BK / 16, (p.cinp - chunk * BK + 15) >> 4);   // the last chunk may be partly padding
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle row: +2 in the (addr >> 4) field
                        if (k < ks) umma_bf16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
                    }
                    umma_commit(empty_bar(s));              // frees the smem stage once these MMAs retire
                }
                umma_commit(tmem_full_bar(acc));            // accumulator complete
            }
        }
        __syncwarp();
    } else {
        // ---------------------------------------------------------------- epilogue (warps 2..5 -> TMEM lane quarters)
        const int q = warp & 3;                             // this warp may touch TMEM lanes [32q, 32q + 32)
        const int row = q * 32 + lane;
        const int act_code = p.epi.fast ? p.epi.act : -1;
        uint32_t t = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++t) {
            int n0, m0, img, y0, x0;
            tile_origin(tile, n0, m0, img, y0, x0);
            int acc = t & 1;
            uint32_t aph = (t >> 1) & 1;
            mbar_wait(tmem_full_bar(acc), aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            bool valid;
            int64_t off;
            if (p.mode == 0) {
                int m = m0 + row;
                valid = m < p.M;
                off = (int64_t)m * p.CoutP;
            } else {
                int ty = row / p.TW, tx = row - ty * p.TW;
                int yy = y0 + ty, xx = x0 + tx;
                valid = yy < p.H && xx < p.W;
                off = (((int64_t)img * p.H + yy) * p.W + xx) * p.CoutP;
            }
            const uint32_t trow = tmem_base + (uint32_t)(acc * BN) + ((uint32_t)(q * 32) << 16);
            int64_t ctb = 0;
            bf16* dst = y + off;
            if (p.convT) {                                  // mode 0 rows are input pixels (n, yy, xx)
                int m = m0 + row;
                int xx = m % p.W;
                int r = m / p.W;
                int yy = r % p.H;
                int nn = r / p.H;
                ctb = ((((int64_t)nn * (2 * p.H) + 2 * yy) * (2 * p.W)) + 2 * xx) * p.ctC;
                dst = y;
            }
            if (p.ct2) {
                const int m = m0 + row;
                switch (act_code) {
                    case EPI_RELU: epilogue_row_ct2<EPI_RELU>(trow, bias, w2s, valid, m, p); break;
                    case EPI_HSWISH: epilogue_row_ct2<EPI_HSWISH>(trow, bias, w2s, valid, m, p); break;
                    default: epilogue_row_ct2<-1>(trow, bias, w2s, valid, m, p); break;
                }
            } else
            switch (act_code) {
                case 0: epilogue_row<BN, 0>(trow, bias, dst, valid, n0, p, ctb); break;
                case EPI_RELU: epilogue_row<BN, EPI_RELU>(trow, bias, dst, valid, n0, p, ctb); break;
                case EPI_HSWISH: epilogue_row<BN, EPI_HSWISH>(trow, bias, dst, valid, n0, p, ctb); break;
                case EPI_SIGMOID: epilogue_row<BN, EPI_SIGMOID>(trow, bias, dst, valid, n0, p, ctb); break;
                case EPI_SWISH: epilogue_row<BN, EPI_SWISH>(trow, bias, dst, valid, n0, p, ctb); break;
                case EPI_HSIGMOID: epilogue_row<BN, EPI_HSIGMOID>(trow, bias, dst, valid, n0, p, ctb); break;
                default: epilogue_row<BN, -1>(trow, bias, dst, valid, n0, p, ctb); break;
            }
            // all TMEM reads of this warp are complete (wait::ld above): hand the accumulator back to the MMA warp
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * BN)) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Weight-stationary variant (the default): the layers of these mobile networks have small weight matrices, so a CTA
// fixes its output-channel tile, loads that tile's WHOLE weight slab (all taps x K chunks) into shared memory once and
// then streams only activation tiles through the TMA ring.  Compared with streaming A and B together this halves the
// L2 -> SM traffic per tile (the per-SM L2 bandwidth was the limiter) and frees ring space for more A stages in flight.
// BN (32..256) is a run-time parameter chosen so that the slab fits; with BN = 256 a 240-channel layer reads its
// activations exactly once.  Eight epilogue warps (two per TMEM lane quarter, each taking half of the columns) drain
// the double-buffered accumulators.
#ifndef B200_WS_EPI_WARPS
#define B200_WS_EPI_WARPS 8
#endif
constexpr int kWsEpiWarps = B200_WS_EPI_WARPS;           // two per TMEM lane quarter, each taking half of the tile's columns (more warps
                                         // measured slower: the waiting warps' barrier polling competes for issue slots)
constexpr int kWsThreads = 64 + 32 * kWsEpiWarps;   // warp 0 TMA, warp 1 MMA, warps 2..17 epilogue
constexpr int WS_MAX_STAGES = 12;
constexpr int WS_MAX_ACC = 8;
constexpr int WS_A_BYTES = BM * BK * 2;
constexpr size_t WS_SMEM_LIMIT = 225 * 1024;

constexpr int kGapReplicas = 16;          // channel-sum buffers are replicated (CTA % 16) so that the atomics spread over lines
inline size_t ws_tail_bytes() {
    return 1024 /*alignment*/ + 512 /*barriers*/ + (4 * CT2_MAX_C + 4) * 4 + 64 + 256 * 4 /*bias slice*/ +
           (size_t)kWsEpiWarps * 128 * 8 /*per-warp channel-sum accumulators*/;
}
// epilogue staging: 16 warps x 32 rows x (their column part in bf16 + 16 bytes of padding against bank conflicts)
inline int ws_parts(int bn) { return std::min(kWsEpiWarps / 4, bn / 16); }             // column parts (each a multiple of 16 columns)
inline size_t ws_stage_row_bytes(int bn) { return (size_t)(bn / ws_parts(bn)) * 2 + 16; }
inline size_t ws_staging_bytes(int bn) { return (size_t)kWsEpiWarps * 32 * ws_stage_row_bytes(bn); }
inline size_t ws_slab_bytes(int bn, int nkb) { return (size_t)nkb * bn * BK * 2 + ws_staging_bytes(bn); }
// largest BN whose slab leaves room for at least 3 activation stages; 0 when none fits
// halo mode (k x k convolutions): one stage holds the (TH + kh - 1) x (TW + kw - 1) pixel halo of a 16 x 8 output tile
// for one 64-channel chunk; the taps read it through shifted descriptors instead of re-loading it from L2 per tap
constexpr int WS_HALO_TW = 8, WS_HALO_TH = 16;
inline size_t ws_halo_stage_bytes(int kh, int kw) {
    size_t b = (size_t)(WS_HALO_TH + kh - 1) * (WS_HALO_TW + kw - 1) * 128;
    return (b + 1023) & ~(size_t)1023;
}
inline int ws_pick_bn(int coutp, int nkb, size_t stage_bytes = WS_A_BYTES) {
    int want = coutp <= 32 ? 32 : (coutp <= 64 ? 64 : (coutp <= 128 ? 128 : 256));
    for (int bn = want; bn >= 32; bn >>= 1)
        if (ws_slab_bytes(bn, nkb) + 3 * stage_bytes + ws_tail_bytes() <= WS_SMEM_LIMIT) return bn;
    return 0;
}
inline int ws_pick_stages(int bn, int nkb, size_t stage_bytes = WS_A_BYTES) {
    size_t left = WS_SMEM_LIMIT - ws_slab_bytes(bn, nkb) - ws_tail_bytes();
    return (int)std::min<size_t>(WS_MAX_STAGES, left / stage_bytes);
}
inline size_t ws_smem_bytes(int bn, int nkb, int stages, size_t stage_bytes = WS_A_BYTES) {
    return ws_slab_bytes(bn, nkb) + (size_t)stages * stage_bytes + ws_tail_bytes();
}


// columns [c_begin, c_end) of this thread's accumulator row -> bias, activation chain, bf16, 16-byte NHWC stores
template <int ACT>
__device__ __forceinline__ void epilogue_cols(uint32_t trow, const float* __restrict__ bias, bf16* __restrict__ dst, bool valid,
                                              int n0, int c_begin, int c_end, const Params& p, int64_t ct_base) {
#pragma unroll 1
    for (int c = c_begin; c < c_end; c += 16) {
        uint32_t v[16];
        tmem_ld8(trow + (uint32_t)c, v);
        tmem_ld8(trow + (uint32_t)c + 8, v + 8);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int col = n0 + c + 8 * h;
            if (valid && col < p.CoutP) {
                float4 b0 = *reinterpret_cast<const float4*>(bias + col);
                float4 b1 = *reinterpret_cast<const float4*>(bias + col + 4);
                float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                float o[8];
                int64_t doff = col;
                int lim = p.Cout - col;
                if (p.convT) {                            // 8-column groups never straddle taps (ctC % 8 == 0)
                    int tap = col / p.ctC, co = col - tap * p.ctC;
                    doff = ct_base + ((int64_t)(tap >> 1) * (2 * p.W) + (tap & 1)) * p.ctC + co;
                    lim = p.ctCout - co;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float f = epi_fast<ACT>(__uint_as_float(v[8 * h + j]) + bb[j], p.epi);
                    o[j] = (j < lim) ? f : 0.f;
                }
                uint4 pk;
                __nv_bfloat162 t0 = __floats2bfloat162_rn(o[0], o[1]);
                __nv_bfloat162 t1 = __floats2bfloat162_rn(o[2], o[3]);
                __nv_bfloat162 t2 = __floats2bfloat162_rn(o[4], o[5]);
                __nv_bfloat162 t3 = __floats2bfloat162_rn(o[6], o[7]);
                pk.x = *reinterpret_cast<uint32_t*>(&t0);
                pk.y = *reinterpret_cast<uint32_t*>(&t1);
                pk.z = *reinterpret_cast<uint32_t*>(&t2);
                pk.w = *reinterpret_cast<uint32_t*>(&t3);
                *reinterpret_cast<uint4*>(dst + doff) = pk;
            }
        }
    }
}

// ---- packed fp32x2 arithmetic (sm_100: FADD2 / FMUL2 / FFMA2 issue two fp32 lanes per instruction)
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t f2_pack(float lo, float hi) {
    f2_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ f2_t f2_pack_u(uint32_t lo, uint32_t hi) {
    f2_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}
__device__ __forceinline__ void f2_unpack(f2_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f2_t f2_add(f2_t a, f2_t b) {
    f2_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f2_t f2_mul(f2_t a, f2_t b) {
    f2_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f2_t f2_fma(f2_t a, f2_t b, f2_t c) {
    f2_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ float fma_sat(float a, float b, float c) {
    float r;
    asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}

// act(acc + bias) * a1 + b1 on two columns at once.  The leading affine of the chain is already folded into the
// weights and the bias (DevWeights::tc_prefolded), hardswish is v * sat(v / 6 + 0.5): one saturating FMA per lane.
template <int ACT>
__device__ __forceinline__ uint32_t epi_pair(uint32_t acc_lo, uint32_t acc_hi, f2_t bias2, f2_t a1, f2_t b1, bool post,
                                             const Epilogue& e) {
    f2_t v = f2_add(f2_pack_u(acc_lo, acc_hi), bias2);
    float x, y;
    if (ACT == EPI_HSWISH) {
        f2_unpack(v, x, y);
        v = f2_mul(v, f2_pack(fma_sat(x, 0.16666667f, 0.5f), fma_sat(y, 0.16666667f, 0.5f)));
    } else if (ACT == EPI_RELU) {
        f2_unpack(v, x, y);
        v = f2_pack(fmaxf(x, 0.f), fmaxf(y, 0.f));
    } else if (ACT != 0) {
        f2_unpack(v, x, y);
        x = epi_act(x, e.act, e.pa, e.pb);
        y = epi_act(y, e.act, e.pa, e.pb);
        v = f2_pack(x, y);
    }
    if (post) v = f2_fma(v, a1, b1);
    f2_unpack(v, x, y);
    __nv_bfloat162 t = __floats2bfloat162_rn(x, y);
    return *reinterpret_cast<uint32_t*>(&t);
}

// Coalesced epilogue: columns [c_begin, c_end) of the warp's 32 accumulator rows -> bias, activation, bf16 -> a padded
// shared-memory tile owned by this warp -> 16-byte global stores in which consecutive lanes write consecutive chunks of
// a row (full 128-byte lines) instead of every lane writing into its own row.  `bias_s` is the CTA's bias slice in
// shared memory (index = column inside the tile); off8 / valid describe this lane's row (offset in 8-element units).
// Channel sums of one staged tile part (see epilogue_cols_staged); deliberately not inlined: one copy of this code for all
// activation variants of the epilogue keeps the kernel's instruction footprint down (small layers are launch-latency bound
// and start with a cold instruction cache).
__device__ __noinline__ void gap_accumulate_tile(const uint8_t* stage, int row_bytes, int ncols, bool valid, int lane,
                                                 unsigned long long* gacc) {
        // Column sums of the staged (bf16-rounded) values over this warp's valid rows.  All 32 lanes work: a lane owns one
        // pair of columns and, when the part has fewer than 32 pairs, one group of rows (partial sums then meet through
        // xor shuffles — a fixed tree, so the result is reproducible).  Loads are unconditional (rows that do not exist
        // are masked to zero), fully unrolled, conflict-free (consecutive lanes read consecutive words of a row).  The
        // tile's sums go as fixed point into this warp's own accumulators in shared memory (integer adds are associative:
        // the total does not depend on which tiles a CTA happens to process); they are flushed per image.
        const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
        const int npairs = ncols >> 1;
        int lgp = 0;
        while ((1 << lgp) < npairs && lgp < 5) ++lgp;         // lanes per row group = min(32, next pow2 >= npairs)
        const int lanes_pp = 1 << lgp, groups = 32 >> lgp, rows_pg = 32 / groups;
        const int rg = lane >> lgp, pl = lane & (lanes_pp - 1);
        const uint32_t sbase = smem_u32(stage);
        const uint32_t gbase = smem_u32(gacc);
        for (int p0 = 0; p0 < npairs; p0 += lanes_pp) {
            const int cp = p0 + pl;
            const bool live = cp < npairs;
            float s0 = 0.f, s1 = 0.f;
            const uint32_t a0 = sbase + (uint32_t)(rg * rows_pg) * (uint32_t)row_bytes + (uint32_t)(live ? cp : 0) * 4u;
            const uint32_t vm = vmask >> (rg * rows_pg);
            if (vmask == 0xffffffffu) {                     // every row of the warp exists (all tiles but an image's last)
#pragma unroll 8
                for (int r = 0; r < rows_pg; ++r) {
                    uint32_t u;
                    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u) : "r"(a0 + (uint32_t)r * (uint32_t)row_bytes));
                    s0 += __uint_as_float(u << 16);
                    s1 += __uint_as_float(u & 0xffff0000u);
                }
            } else {
#pragma unroll 8
                for (int r = 0; r < rows_pg; ++r) {
                    uint32_t u;
                    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u) : "r"(a0 + (uint32_t)r * (uint32_t)row_bytes));
                    u &= 0u - ((vm >> r) & 1u);
                    s0 += __uint_as_float(u << 16);
                    s1 += __uint_as_float(u & 0xffff0000u);
                }
            }
            for (int d = lanes_pp; d < 32; d <<= 1) {
                s0 += __shfl_xor_sync(0xffffffffu, s0, d);
                s1 += __shfl_xor_sync(0xffffffffu, s1, d);
            }
            if (live && rg == 0) {
                // 48.16 fixed point from two 32-bit conversions: integer part + fraction (s - rint(s) is exact in fp32)
                const float h0 = rintf(s0), h1 = rintf(s1);
                const long long v0 = (long long)__float2int_rn(h0) * 65536ll + (long long)__float2int_rn((s0 - h0) * kGapFixScale);
                const long long v1 = (long long)__float2int_rn(h1) * 65536ll + (long long)__float2int_rn((s1 - h1) * kGapFixScale);
                unsigned long long c0, c1;
                const uint32_t ga = gbase + (uint32_t)cp * 16u;
                asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(c0), "=l"(c1) : "r"(ga));
                c0 += (unsigned long long)v0;
                c1 += (unsigned long long)v1;
                asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(ga), "l"(c0), "l"(c1) : "memory");
            }
        }
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
        "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}

// eight accumulator columns -> bias, activation, bf16 -> one 16-byte store into the lane's staged row
template <int ACT>
__device__ __forceinline__ void epi_group8(const uint32_t* v, const float* __restrict__ bias8, f2_t a1, f2_t b1, bool post,
                                           const Params& p, int lim, uint8_t* dst) {
    const float4 b0 = *reinterpret_cast<const float4*>(bias8);
    const float4 b1v = *reinterpret_cast<const float4*>(bias8 + 4);
    uint4 pk;
    pk.x = epi_pair<ACT>(v[0], v[1], f2_pack(b0.x, b0.y), a1, b1, post, p.epi);
    pk.y = epi_pair<ACT>(v[2], v[3], f2_pack(b0.z, b0.w), a1, b1, post, p.epi);
    pk.z = epi_pair<ACT>(v[4], v[5], f2_pack(b1v.x, b1v.y), a1, b1, post, p.epi);
    pk.w = epi_pair<ACT>(v[6], v[7], f2_pack(b1v.z, b1v.w), a1, b1, post, p.epi);
    if (lim < 8) {                                           // pad lanes of the last group stay zero
        uint32_t* w = &pk.x;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (2 * j >= lim) w[j] = 0u;
            else if (2 * j + 1 >= lim) w[j] &= 0x0000ffffu;
        }
    }
    *reinterpret_cast<uint4*>(dst) = pk;
}

// Coalesced epilogue: columns [c_begin, c_end) of the warp's 32 accumulator rows -> bias, activation, bf16 -> a padded
// shared-memory tile owned by this warp -> 16-byte global stores in which consecutive lanes write consecutive chunks of
// a row (full 128-byte lines) instead of every lane writing into its own row.  `bias_s` is the CTA's bias slice in
// shared memory (index = column inside the tile); off8 / valid describe this lane's row (offset in 8-element units).
// The accumulator is read 32 columns per TMEM load (16 independent packed chains in flight per lane: the epilogue is
// latency-, not throughput-bound); LINEAR: the rows of the warp are consecutive GEMM rows (mode 0), so a row's offset and
// validity follow from its index and the store loop needs no shuffles.
template <int ACT>
__device__ __forceinline__ void epilogue_cols_staged(uint32_t trow, const float* __restrict__ bias_s, bf16* __restrict__ y,
                                                     uint32_t off8, bool valid, int n0, int c_begin, int c_end, const Params& p,
                                                     uint8_t* stage, int row_bytes, int lane, unsigned long long* gacc, bool LINEAR) {
    const int c_lim = min(c_end, p.CoutP - n0);              // columns that exist in the output tensor
    uint8_t* my = stage + (size_t)lane * row_bytes;
    const bool post = p.epi.a1 != 1.f || p.epi.b1 != 0.f;
    const f2_t a1 = f2_pack(p.epi.a1, p.epi.a1), b1 = f2_pack(p.epi.b1, p.epi.b1);
    int c = c_begin;
#pragma unroll 1
    for (; c < c_lim; c += 16) {
        uint32_t v[16];
        tmem_ld16(trow + (uint32_t)c, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int cl = c + 8 * h;
            if (cl < c_lim) epi_group8<ACT>(v + 8 * h, bias_s + cl, a1, b1, post, p, p.Cout - (n0 + cl), my + (size_t)(cl - c_begin) * 2);
        }
    }
    __syncwarp();
    if (gacc) gap_accumulate_tile(stage, row_bytes, c_lim - c_begin, valid, lane, gacc);
    const int chunks = (c_lim - c_begin) >> 3;               // 16-byte chunks per row (CoutP % 8 == 0)
    if (chunks > 0) {
        int lg = 0;
        while ((1 << lg) < chunks) ++lg;                      // lanes per row = next power of two (<= 16 for BN <= 256)
        const int cc = lane & ((1 << lg) - 1), rsub = lane >> lg, rows_per_pass = 32 >> lg;
        const uint32_t col8 = (uint32_t)(n0 + c_begin) >> 3;
        if (LINEAR) {
            // consecutive GEMM rows: row rr sits rr * (CoutP / 8) chunks after the warp's first row and the valid rows are
            // a prefix of the warp's 32
            const uint32_t base8 = __shfl_sync(0xffffffffu, off8, 0) + col8 + (uint32_t)cc;
            const int nvalid = __popc(__ballot_sync(0xffffffffu, valid));
            const uint32_t rstride8 = (uint32_t)p.CoutP >> 3;
            const uint8_t* src = stage + (size_t)rsub * row_bytes + (size_t)cc * 16;
            if (cc < chunks) {
#pragma unroll 4
                for (int r0 = 0; r0 < 32; r0 += rows_per_pass) {
                    const int rr = r0 + rsub;
                    if (rr < nvalid) {
                        const uint4 val = *reinterpret_cast<const uint4*>(src + (size_t)r0 * row_bytes);
                        *reinterpret_cast<uint4*>(y + ((size_t)(base8 + (uint32_t)rr * rstride8) << 3)) = val;
                    }
                }
            }
        } else {
#pragma unroll 1
            for (int r0 = 0; r0 < 32; r0 += rows_per_pass) {
                const int rr = r0 + rsub;
                const uint32_t o_rr = __shfl_sync(0xffffffffu, off8, rr);
                const int v_rr = __shfl_sync(0xffffffffu, (int)valid, rr);
                if (cc < chunks && v_rr) {
                    uint4 val = *reinterpret_cast<const uint4*>(stage + (size_t)rr * row_bytes + (size_t)cc * 16);
                    *reinterpret_cast<uint4*>(y + ((size_t)(o_rr + col8 + cc) << 3)) = val;
                }
            }
        }
    }
    __syncwarp();
}

// DB head tail, one first-stage tap (dy, dx) per epilogue warp part: this thread's accumulator row holds the first
// transposed conv's 4 taps x ctC channels for input pixel (n, yy, xx).  Bias + activation give the mid-resolution
// pixel (2yy + dy, 2xx + dx); the second transposed conv (ctC -> 1 channel, 2x2) and its activation turn it into a 2x2
// block of the final map, written as two 8-byte pieces of the dense fp32 output (the four parts of a row quarter
// together cover the 4x4 block; neighbouring lanes are neighbouring xx, so a warp writes whole lines).  All TMEM
// columns of the tap are fetched with one wait so the four output chains run with full instruction-level parallelism.
// w2q holds the second convolution's weights as float4 per input channel (its 4 taps), then its bias.
template <int ACT>
__device__ __forceinline__ void epilogue_ct2_tap(uint32_t trow, const float* __restrict__ bias_s, const float* __restrict__ w2q,
                                                 bool valid, int m, int tap, const Params& p) {
    const float b2 = w2q[4 * p.ctC];
    float a[4] = {b2, b2, b2, b2};
    uint32_t v[CT2_MAX_C];
#pragma unroll
    for (int cc = 0; cc < CT2_MAX_C; cc += 8)
        if (cc < p.ctC) tmem_ld8(trow + (uint32_t)(tap * p.ctC + cc), v + cc);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const float* bp = bias_s + tap * p.ctC;
#pragma unroll
    for (int cc = 0; cc < CT2_MAX_C; cc += 4) {
        if (cc < p.ctC) {
            const float4 bq = *reinterpret_cast<const float4*>(bp + cc);
            const float bb[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float f = (cc + j < p.ctCout) ? epi_fast<ACT>(__uint_as_float(v[cc + j]) + bb[j], p.epi) : 0.f;
                const float4 w = *reinterpret_cast<const float4*>(w2q + 4 * (cc + j));
                a[0] = fmaf(f, w.x, a[0]);
                a[1] = fmaf(f, w.y, a[1]);
                a[2] = fmaf(f, w.z, a[2]);
                a[3] = fmaf(f, w.w, a[3]);
            }
        }
    }
    if (valid) {
#pragma unroll
        for (int t2 = 0; t2 < 4; ++t2) a[t2] = epi_apply(a[t2], p.epi2);
        const int dy = tap >> 1, dx = tap & 1;
        int xx = m % p.W;
        int r = m / p.W;
        int yy = r % p.H;
        int nn = r / p.H;
        const int64_t OW = 4 * (int64_t)p.W;
        float* dst = p.ct2_out + ((int64_t)nn * (4 * p.H) + 4 * yy + 2 * dy) * OW + 4 * xx + 2 * dx;
        *reinterpret_cast<float2*>(dst) = make_float2(a[0], a[1]);            // (ey = 0, ex = 0..1)
        *reinterpret_cast<float2*>(dst + OW) = make_float2(a[2], a[3]);       // (ey = 1, ex = 0..1)
    }
    __syncwarp();
}

// Same as epilogue_ct2_tap with the channel count fixed at compile time (the PP-OCR DB head has 24): no per-channel
// guards, the four second-stage accumulators advance as two packed FFMA2, sigmoid through the fast exponential.
template <int ACT, int C, bool PLAIN>
__device__ __forceinline__ void epilogue_ct2_tap_fixed(uint32_t trow, const float* __restrict__ bias_s,
                                                       const float* __restrict__ w2q, bool valid, int m, int tap,
                                                       const Params& p) {
    const float b2 = w2q[4 * C];
    f2_t a01 = f2_pack(b2, b2), a23 = a01;
    uint32_t v[C];
#pragma unroll
    for (int cc = 0; cc < C; cc += 8) tmem_ld8(trow + (uint32_t)(tap * C + cc), v + cc);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const float* bp = bias_s + tap * C;
#pragma unroll
    for (int cc = 0; cc < C; cc += 4) {
        const float4 bq = *reinterpret_cast<const float4*>(bp + cc);
        const float bb[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float f = __uint_as_float(v[cc + j]) + bb[j];
            if (PLAIN) {
                if (ACT == EPI_RELU) f = fmaxf(f, 0.f);
                else if (ACT == EPI_HSWISH) f = f * fma_sat(f, 0.16666667f, 0.5f);
            } else {
                f = epi_fast<ACT>(f, p.epi);
            }
            const float4 w = *reinterpret_cast<const float4*>(w2q + 4 * (cc + j));
            const f2_t ff = f2_pack(f, f);
            a01 = f2_fma(ff, f2_pack(w.x, w.y), a01);
            a23 = f2_fma(ff, f2_pack(w.z, w.w), a23);
        }
    }
    if (valid) {
        float a[4];
        f2_unpack(a01, a[0], a[1]);
        f2_unpack(a23, a[2], a[3]);
        if (p.epi2.fast && p.epi2.act == EPI_SIGMOID && p.epi2.a0 == 1.f && p.epi2.b0 == 0.f && p.epi2.a1 == 1.f && p.epi2.b1 == 0.f) {
#pragma unroll
            for (int t2 = 0; t2 < 4; ++t2) a[t2] = __fdividef(1.f, 1.f + __expf(-a[t2]));
        } else {
#pragma unroll
            for (int t2 = 0; t2 < 4; ++t2) a[t2] = epi_apply(a[t2], p.epi2);
        }
        const int dy = tap >> 1, dx = tap & 1;
        int xx = m % p.W;
        int r = m / p.W;
        int yy = r % p.H;
        int nn = r / p.H;
        const int64_t OW = 4 * (int64_t)p.W;
        float* dst = p.ct2_out + ((int64_t)nn * (4 * p.H) + 4 * yy + 2 * dy) * OW + 4 * xx + 2 * dx;
        *reinterpret_cast<float2*>(dst) = make_float2(a[0], a[1]);
        *reinterpret_cast<float2*>(dst + OW) = make_float2(a[2], a[3]);
    }
    __syncwarp();
}

// One 64-channel chunk of a 3x3 halo tile: 9 taps x KS K-slices of 16 channels, issued with RUNNING descriptors (every
// step is one 64-bit add on the descriptor that the previous MMA used) — the issuing thread is the bottleneck of the
// small-N layers (measured 65 cycles per MMA with descriptors rebuilt from a base each time against 40 for the tensor
// pipe), so the loop body must stay at "add, add, mma".  A: the taps of one halo row are 8 units (128 bytes = 64
// channels) apart = exactly 4 K-slices, a row step is (TW + 2) * 8 units.  B: the chunk's taps lie b_tap units apart.
template <int KS>
__device__ __forceinline__ void issue_halo3x3(uint32_t tmem_d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc_first,
                                              uint64_t b_tap) {
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
        for (int k2 = 0; k2 < KS; ++k2) {
            umma_bf16(tmem_d, ad, bd, idesc, (tap | k2) ? 1u : acc_first);
            if (k2 + 1 < KS) { ad += 2; bd += 2; }
        }
        if (tap < 8) {
            ad += (tap % 3 == 2) ? (uint64_t)((WS_HALO_TW + 2) * 8 - 16 - 2 * (KS - 1)) : (uint64_t)(8 - 2 * (KS - 1));
            bd += b_tap - (uint64_t)(2 * (KS - 1));
        }
    }
}

__global__ void __launch_bounds__(kWsThreads, 1) conv_tc_ws_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                const __grid_constant__ CUtensorMap map_b,
                                                                const float* __restrict__ bias, bf16* __restrict__ y,
                                                                const Params p) {
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    const int BN = p.ws_bn, STG = p.ws_stages;
    const int nkb = p.taps * p.kchunks;
    const int B_KB_BYTES = BN * BK * 2;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t b_base = base;
    const uint32_t a_base = b_base + (uint32_t)nkb * B_KB_BYTES;
    const uint32_t ASB = p.ws_halo ? (uint32_t)p.ws_stage_bytes : (uint32_t)WS_A_BYTES;   // bytes per activation stage
    const uint32_t stage_base = a_base + (uint32_t)STG * ASB;
    const int parts = min(kWsEpiWarps / 4, BN / 16);        // column parts of the tile shared out among the epilogue warps
    const int stage_row = (BN / parts) * 2 + 16;            // bytes per staged row (one part in bf16 + padding)
    const uint32_t bar_base = (stage_base + (uint32_t)kWsEpiWarps * 32u * (uint32_t)stage_row + 63u) & ~63u;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (WS_MAX_STAGES + s); };
    // accumulator ring in TMEM: as many BN-column buffers as fit (at most 8), so that the MMA warp runs several tiles
    // ahead of the epilogue and the per-tile barrier round trips of small layers overlap instead of adding up
    const int NACC = min(WS_MAX_ACC, 512 / BN);
    auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * WS_MAX_STAGES + a); };
    auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * WS_MAX_STAGES + WS_MAX_ACC + a); };
    const uint32_t b_bar = bar_base + 8u * (2 * WS_MAX_STAGES + 2 * WS_MAX_ACC);
    const uint32_t tmem_slot = bar_base + 8u * (2 * WS_MAX_STAGES + 2 * WS_MAX_ACC + 1);
    uint32_t* tmem_slot_ptr = (uint32_t*)(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    float* w2q = (float*)(smem_raw + (bar_base + 512u - smem_u32(smem_raw)));
    float* bias_s = w2q + (4 * CT2_MAX_C + 4);              // this CTA's bias slice [BN] (its output-channel tile is fixed)
    unsigned long long* gacc_all = reinterpret_cast<unsigned long long*>(bias_s + 256);   // [epilogue warp][128 columns]
    {
        const int nt0 = (blockIdx.x % ((p.CoutP + BN - 1) / BN)) * BN;
        for (int i = threadIdx.x; i < BN; i += kWsThreads) bias_s[i] = bias[nt0 + i];
    }
    if (p.ct2) {                                            // [ci][4 taps] then the bias
        for (int i = threadIdx.x; i < 4 * p.ctC; i += kWsThreads) w2q[i] = p.ct2_w[(i & 3) * p.ctC + (i >> 2)];
        if (threadIdx.x == 0) w2q[4 * p.ctC] = p.ct2_w[4 * p.ctC];
    }

    const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
    const int tpi = (p.gap_rows + BM - 1) / BM;            // mode 0 with channel sums: tiles per image
    const int m_tiles = p.mode == 0 ? (p.gap ? p.N * tpi : (p.M + BM - 1) / BM) : p.tiles_x * p.tiles_y * p.N;
    const int n_tiles = (p.CoutP + BN - 1) / BN;
    const int my_nt = blockIdx.x % n_tiles;
    const int n0 = my_nt * BN;
    const int m_first = blockIdx.x / n_tiles, m_step = gridDim.x / n_tiles;
    const uint32_t tmem_cols = (uint32_t)(NACC * BN);

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STG; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < NACC; ++a) {
            mbar_init(tmem_full_bar(a), 1);
            mbar_init(tmem_empty_bar(a), kWsEpiWarps);      // one arrival per epilogue warp
        }
        mbar_init(b_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot_ptr;
    // Programmatic dependent launch: everything up to here — and the producer's weight-slab loads below — overlaps the tail
    // of the previous kernel.  The producer waits for it right before its first ACTIVATION load, the other roles before
    // they touch anything (their first action depends on the producer's loads anyway).
    if (warp != 0) pdl_wait();

    auto tile_origin = [&](int mt, int& m0, int& img, int& y0, int& x0) {
        m0 = mt * BM;
        img = 0; y0 = 0; x0 = 0;
        if (p.mode == 0 && p.gap) {                         // y0 = rows of this tile that belong to the image
            img = mt / tpi;
            const int tl = mt - img * tpi;
            m0 = img * p.gap_rows + tl * BM;
            y0 = p.gap_rows - tl * BM;
        }
        if (p.mode == 1) {
            int tx = mt % p.tiles_x;
            int r = mt / p.tiles_x;
            int ty = r % p.tiles_y;
            img = r / p.tiles_y;
            y0 = ty * p.TH;
            x0 = tx * p.TW;
        }
    };

    if (warp == 0) {
        // ---------------------------------------------------------------- TMA producer
        if (elect_one_sync()) {
            mbar_expect_tx(b_bar, (uint32_t)nkb * B_KB_BYTES);
            // slab slot of K block kb (= tap * kchunks + chunk in the packed weights): halo mode walks chunk-major, so a
            // chunk's taps are made contiguous there
            for (int kb = 0; kb < nkb; ++kb) {
                const int slot = p.ws_halo ? (kb % p.kchunks) * p.taps + kb / p.kchunks : kb;
                tma_load_2d(b_base + slot * B_KB_BYTES, &map_b, b_bar, kb * BK, n0);
            }
            pdl_wait();                                     // activations of the previous kernel are complete and visible
            uint32_t it = 0;
            for (int mt = m_first; mt < m_tiles; mt += m_step) {
                int m0, img, y0, x0;
                tile_origin(mt, m0, img, y0, x0);
                if (p.ws_halo) {
                    const uint32_t halo_bytes = (uint32_t)(p.TH + p.kh - 1) * (uint32_t)(p.TW + p.kw - 1) * 128u;
                    for (int ch = 0; ch < p.kchunks; ++ch, ++it) {
                        int s = it % STG;
                        uint32_t ph = (it / STG) & 1;
                        const long long c3 = p.dbg ? clock64() : 0;
                        mbar_wait(empty_bar(s), ph ^ 1);
                        if (p.dbg) p.dbg[(size_t)blockIdx.x * 8 + 6] += (unsigned long long)(clock64() - c3);   // producer: waiting for a free stage
                        mbar_expect_tx(full_bar(s), halo_bytes);
                        tma_load_4d(a_base + s * ASB, &map_a, full_bar(s), ch * BK, x0 - p.pad_l, y0 - p.pad_t, img);
                    }
                    continue;
                }
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    int s = it % STG;
                    uint32_t ph = (it / STG) & 1;
                    mbar_wait(empty_bar(s), ph ^ 1);
                    mbar_expect_tx(full_bar(s), WS_A_BYTES);
                    int tap = kb / p.kchunks, chunk = kb - tap * p.kchunks;
                    if (p.mode == 0) {
                        tma_load_2d(a_base + s * WS_A_BYTES, &map_a, full_bar(s), chunk * BK, m0);
                    } else {
                        int ky = tap / p.kw, kx = tap - ky * p.kw;
                        tma_load_4d(a_base + s * WS_A_BYTES, &map_a, full_bar(s), chunk * BK, x0 + kx - p.pad_l, y0 + ky - p.pad_t, img);
                    }
                }
            }
        }
        else pdl_wait();                                    // the other 31 lanes of the producer warp
        __syncwarp();
    } else if (warp == 1) {
        // ---------------------------------------------------------------- MMA issuer
        if (elect_one_sync()) {
            const uint32_t idesc = make_idesc(BN);
            long long dbg_acc = 0, dbg_full = 0;
            const long long dbg_t0 = p.dbg ? clock64() : 0;
            mbar_wait(b_bar, 0);
            const long long dbg_t1 = p.dbg ? clock64() : 0;
            uint32_t it = 0, t = 0;
            if (p.ws_halo == 2 && p.taps == 9 && p.kw == 3 && STG >= 2 * p.kchunks + 1 && NACC >= 2) {
                // 3x3 halo mode, TWO tiles in flight: the 72 small MMAs of a tile all accumulate into the same TMEM
                // columns, so back-to-back they run at the pipeline latency, not its throughput (measured ~100 cycles per
                // N = 32 MMA).  Alternating between two tiles' accumulators gives the tensor pipe independent work.
                const uint32_t hw = (uint32_t)(p.TW + p.kw - 1);
                const uint64_t bstep = (uint64_t)((uint32_t)B_KB_BYTES >> 4);
                for (int mt = m_first; mt < m_tiles; mt += 2 * m_step, t += 2) {
                    const bool two = mt + m_step < m_tiles;
                    const int acc0 = t % NACC, acc1 = (t + 1) % NACC;
                    mbar_wait(tmem_empty_bar(acc0), ((t / NACC) & 1) ^ 1);
                    if (two) mbar_wait(tmem_empty_bar(acc1), (((t + 1) / NACC) & 1) ^ 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t d0 = tmem_base + (uint32_t)(acc0 * BN), d1 = tmem_base + (uint32_t)(acc1 * BN);
                    for (int ch = 0; ch < p.kchunks; ++ch) {
                        const uint32_t i0 = it + ch, i1 = it + p.kchunks + ch;
                        const int s0 = i0 % STG, s1 = i1 % STG;
                        mbar_wait(full_bar(s0), (i0 / STG) & 1);
                        if (two) mbar_wait(full_bar(s1), (i1 / STG) & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint64_t a0 = make_smem_desc_sbo(a_base + s0 * ASB, hw * 128u);
                        const uint64_t a1 = make_smem_desc_sbo(a_base + s1 * ASB, hw * 128u);
                        const uint64_t b0 = make_smem_desc(b_base + ch * 9 * B_KB_BYTES);
                        const int ks = min(BK / 16, (p.cinp - ch * BK + 15) >> 4);
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            const uint64_t ao = (uint64_t)(((tap / 3) * (WS_HALO_TW + 2) + (tap % 3)) * 8);
                            const uint64_t bd = b0 + bstep * (uint64_t)tap;
#pragma unroll
                            for (int k2 = 0; k2 < BK / 16; ++k2) {
                                if (k2 >= ks) continue;
                                umma_bf16(d0, a0 + ao + (uint64_t)(2 * k2), bd + (uint64_t)(2 * k2), idesc, (ch | tap | k2) != 0);
                                if (two) umma_bf16(d1, a1 + ao + (uint64_t)(2 * k2), bd + (uint64_t)(2 * k2), idesc, (ch | tap | k2) != 0);
                            }
                        }
                        umma_commit(empty_bar(s0));
                        if (two) umma_commit(empty_bar(s1));
                    }
                    it += (two ? 2 : 1) * p.kchunks;
                    umma_commit(tmem_full_bar(acc0));
                    if (two) umma_commit(tmem_full_bar(acc1));
                }
            } else
            for (int mt = m_first; mt < m_tiles; mt += m_step, ++t) {
                int acc = t % NACC;
                uint32_t aph = (t / NACC) & 1;
                long long c0 = p.dbg ? clock64() : 0;
                mbar_wait(tmem_empty_bar(acc), aph ^ 1);
                if (p.dbg) dbg_acc += clock64() - c0;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
                if (p.ws_halo) {
                    const uint32_t hw = (uint32_t)(p.TW + p.kw - 1);          // halo width in pixels
                    for (int ch = 0; ch < p.kchunks; ++ch, ++it) {
                        int s = it % STG;
                        uint32_t ph = (it / STG) & 1;
                        long long c1 = p.dbg ? clock64() : 0;
                        mbar_wait(full_bar(s), ph);
                        if (p.dbg) dbg_full += clock64() - c1;
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        if (p.taps == 9 && p.kw == 3) {
                            // 3x3: descriptor arithmetic hoisted — one base descriptor per stage / chunk, per-tap offsets are
                            // compile-time (A: (ky * 10 + kx) halo rows of 128 bytes) or one multiply (B); the single
                            // issuing thread would otherwise spend more time building descriptors than the MMAs take
                            const uint64_t a0 = make_smem_desc_sbo(a_base + s * ASB, hw * 128u);
                            const uint64_t b0 = make_smem_desc(b_base + ch * 9 * B_KB_BYTES);
                            const uint64_t b_tap = (uint64_t)((uint32_t)B_KB_BYTES >> 4);
                            const int ks = min(BK / 16, (p.cinp - ch * BK + 15) >> 4);
                            const uint32_t accf = ch != 0 ? 1u : 0u;
                            if (ks == 4) issue_halo3x3<4>(tmem_d, a0, b0, idesc, accf, b_tap);
                            else if (ks == 2) issue_halo3x3<2>(tmem_d, a0, b0, idesc, accf, b_tap);
                            else if (ks == 3) issue_halo3x3<3>(tmem_d, a0, b0, idesc, accf, b_tap);
                            else issue_halo3x3<1>(tmem_d, a0, b0, idesc, accf, b_tap);
                        } else
                        for (int tap = 0; tap < p.taps; ++tap) {
                            int ky = tap / p.kw, kx = tap - ky * p.kw;
                            // rows of the A operand = pixels (ty + ky, tx + kx) of the halo tile: 8-pixel groups, one halo row apart
                            uint64_t adesc = make_smem_desc_sbo(a_base + s * ASB + ((uint32_t)ky * hw + (uint32_t)kx) * 128u, hw * 128u);
                            uint64_t bdesc = make_smem_desc(b_base + (ch * p.taps + tap) * B_KB_BYTES);
                            const int ks = min(BK / 16, (p.cinp - ch * BK + 15) >> 4);
#pragma unroll
                            for (int k2 = 0; k2 < BK / 16; ++k2)
                                if (k2 < ks) umma_bf16(tmem_d, adesc + (uint64_t)(2 * k2), bdesc + (uint64_t)(2 * k2), idesc, (ch | tap | k2) != 0);
                        }
                        umma_commit(empty_bar(s));
                    }
                } else
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    int s = it % STG;
                    uint32_t ph = (it / STG) & 1;
                    long long c1 = p.dbg ? clock64() : 0;
                    mbar_wait(full_bar(s), ph);
                    if (p.dbg) dbg_full += clock64() - c1;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    uint64_t adesc = make_smem_desc(a_base + s * WS_A_BYTES);
                    uint64_t bdesc = make_smem_desc(b_base + kb * B_KB_BYTES);
                    const int ks = min(BK / 16, (p.cinp - (kb % p.kchunks) * BK + 15) >> 4);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        if (k < ks) umma_bf16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
                    umma_commit(empty_bar(s));
                }
                umma_commit(tmem_full_bar(acc));
            }
            if (p.dbg) {
                unsigned long long* d = p.dbg + (size_t)blockIdx.x * 8;
                d[0] = (unsigned long long)(clock64() - dbg_t0);   // MMA thread: total
                d[1] = (unsigned long long)(dbg_t1 - dbg_t0);      //   waiting for the weight slab
                d[2] = (unsigned long long)dbg_acc;                //   waiting for a free accumulator
                d[3] = (unsigned long long)dbg_full;               //   waiting for activation stages
            }
        }
        __syncwarp();
    } else {
        // ---------------------------------------------------------------- epilogue: warps 2..17
        const int q = warp & 3;                             // TMEM lane quarter this warp may touch
        const int part = (warp - 2) >> 2;                   // which part of the tile's columns (0..3)
        const int row = q * 32 + lane;
        const int c_begin = part * (BN / parts), c_end = part < parts ? c_begin + BN / parts : c_begin;
        const int act_code = p.epi.fast ? p.epi.act : -1;
        uint32_t t = 0;
        long long dbg_wait = 0;
        const long long dbg_e0 = p.dbg ? clock64() : 0;
        // channel sums (SE producer): this warp's accumulators for its column part, flushed whenever the image changes
        unsigned long long* gacc = p.gap ? gacc_all + (size_t)(warp - 2) * 128 : nullptr;
        int gap_cur = -1;
        auto gap_flush = [&](int img_done) {
            __syncwarp();
            unsigned long long* dst = p.gap + ((size_t)img_done * kGapReplicas + (blockIdx.x % kGapReplicas)) * p.gap_mod;
            for (int cl = lane; cl < c_end - c_begin; cl += 32) {
                const unsigned long long v = gacc[cl];
                const int col = n0 + c_begin + cl;
                if (v != 0ull && col < p.CoutP) atomicAdd(dst + col % p.gap_mod, v);
                gacc[cl] = 0ull;
            }
            __syncwarp();
        };
        if (gacc) {
            for (int cl = lane; cl < 128; cl += 32) gacc[cl] = 0ull;
            __syncwarp();
        }
        for (int mt = m_first; mt < m_tiles; mt += m_step, ++t) {
            int m0, img, y0, x0;
            tile_origin(mt, m0, img, y0, x0);
            if (gacc && img != gap_cur) {
                if (gap_cur >= 0) gap_flush(gap_cur);
                gap_cur = img;
            }
            int acc = t % NACC;
            uint32_t aph = (t / NACC) & 1;
            const long long c2 = p.dbg ? clock64() : 0;
            mbar_wait(tmem_full_bar(acc), aph);
            if (p.dbg) dbg_wait += clock64() - c2;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            bool valid;
            int64_t off;
            if (p.mode == 0) {
                int m = m0 + row;
                valid = p.gap ? row < y0 : m < p.M;
                off = (int64_t)m * p.CoutP;
            } else {
                int ty = row / p.TW, tx = row - ty * p.TW;
                int yy = y0 + ty, xx = x0 + tx;
                valid = yy < p.H && xx < p.W;
                off = (((int64_t)img * p.H + yy) * p.W + xx) * p.CoutP;
            }
            const uint32_t trow = tmem_base + (uint32_t)(acc * BN) + ((uint32_t)(q * 32) << 16);
            int64_t ctb = 0;
            bf16* dst = y + off;
            if (p.convT) {                                  // mode 0 rows are input pixels (n, yy, xx)
                int m = m0 + row;
                int xx = m % p.W;
                int r = m / p.W;
                int yy = r % p.H;
                int nn = r / p.H;
                ctb = ((((int64_t)nn * (2 * p.H) + 2 * yy) * (2 * p.W)) + 2 * xx) * p.ctC;
                dst = y;
            }
            if (p.ct2) {
                const int m = m0 + row;
                const bool plain = p.epi.a0 == 1.f && p.epi.b0 == 0.f && p.epi.a1 == 1.f && p.epi.b1 == 0.f;
                for (int tap = part * (4 / parts); tap < (part + 1) * (4 / parts); ++tap)   // first-stage taps of this warp
                    if (p.ctC == 24 && p.ctCout == 24 && act_code == EPI_RELU && plain)
                        epilogue_ct2_tap_fixed<EPI_RELU, 24, true>(trow, bias_s, w2q, valid, m, tap, p);
                    else if (p.ctC == 24 && p.ctCout == 24 && act_code == EPI_RELU)
                        epilogue_ct2_tap_fixed<EPI_RELU, 24, false>(trow, bias_s, w2q, valid, m, tap, p);
                    else
                    switch (act_code) {
                        case EPI_RELU: epilogue_ct2_tap<EPI_RELU>(trow, bias_s, w2q, valid, m, tap, p); break;
                        case EPI_HSWISH: epilogue_ct2_tap<EPI_HSWISH>(trow, bias_s, w2q, valid, m, tap, p); break;
                        default: epilogue_ct2_tap<-1>(trow, bias_s, w2q, valid, m, tap, p); break;
                    }
            } else {
                if (p.convT) {
                    switch (act_code) {
                        case 0: epilogue_cols<0>(trow, bias, dst, valid, n0, c_begin, c_end, p, ctb); break;
                        case EPI_RELU: epilogue_cols<EPI_RELU>(trow, bias, dst, valid, n0, c_begin, c_end, p, ctb); break;
                        case EPI_HSWISH: epilogue_cols<EPI_HSWISH>(trow, bias, dst, valid, n0, c_begin, c_end, p, ctb); break;
                        default: epilogue_cols<-1>(trow, bias, dst, valid, n0, c_begin, c_end, p, ctb); break;
                    }
                } else {
                    uint8_t* stg = smem_raw + (stage_base - smem_u32(smem_raw)) + (size_t)(warp - 2) * 32 * stage_row;
                    const bool linear_rows = p.mode == 0 && !(p.epi_variant & 2);
#define B200_EPI(A) epilogue_cols_staged<A>(trow, bias_s, y, (uint32_t)(off >> 3), valid, n0, c_begin, c_end, p, stg, stage_row, lane, gacc, linear_rows)
                    switch (act_code) {
                        case 0: B200_EPI(0); break;
                        case EPI_RELU: B200_EPI(EPI_RELU); break;
                        case EPI_HSWISH: B200_EPI(EPI_HSWISH); break;
                        case EPI_SIGMOID:
                        case EPI_SWISH:
                        case EPI_HSIGMOID: B200_EPI(99); break;           // rare on convolutions: one variant, activation chosen at run time
                        default: epilogue_cols<-1>(trow, bias, dst, valid, n0, c_begin, c_end, p, ctb); break;   // interpreted chain
                    }
#undef B200_EPI
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
        }
        if (gacc && gap_cur >= 0) gap_flush(gap_cur);
        if (p.dbg && warp == 2 && lane == 0) {
            p.dbg[(size_t)blockIdx.x * 8 + 4] = (unsigned long long)(clock64() - dbg_e0);   // epilogue warp 2: total
            p.dbg[(size_t)blockIdx.x * 8 + 5] = (unsigned long long)dbg_wait;               //   waiting for accumulators
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------------------------
// CTC head: logits = A[rows][K] * W[classes][K]^T + bias, softmax over the classes and
// CTCLabelDecode's argmax / max (rec_postprocess.py:98-104) in ONE kernel — the [rows][classes] logits (753 MB at
// 256 x 40 x 18385 fp32) never exist in HBM.  A CTA owns one 128-row tile and a contiguous range of 128-class tiles:
// the A tile is loaded once and stays in shared memory, weight tiles stream through a TMA ring, tcgen05.mma fills two
// TMEM accumulators alternately, and each epilogue thread (one row) keeps a running (max, first argmax, sum of
// exp(v - max)) in registers.  grid = (row tiles, class splits); partial results of the splits are merged by
// ctc_head_merge_kernel (max prob = 1 / sum exp(v - max)).
struct HeadParams {
    int rows;                    // B * T
    int n_tiles;                 // ceil(classes / 128)
    int tiles_per_split;
    int kblocks;                 // ceil(K / 64)  (<= 4)
};

constexpr int HEAD_BN = 128;
constexpr int HEAD_STAGES = 4;
constexpr int HEAD_MAX_KB = 4;

inline size_t head_smem_bytes(int kblocks) {
    return (size_t)kblocks * (BM * BK * 2) + (size_t)HEAD_STAGES * kblocks * (HEAD_BN * BK * 2) + 2 * HEAD_BN * 4 + 1024 + 256;
}

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

__global__ void __launch_bounds__(kThreads) ctc_head_kernel(const __grid_constant__ CUtensorMap map_a,
                                                            const __grid_constant__ CUtensorMap map_b,
                                                            const float* __restrict__ bias /* -1e30 beyond the last class */,
                                                            float* __restrict__ pmax, int* __restrict__ pidx,
                                                            float* __restrict__ psum, const HeadParams p) {
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    constexpr int A_KB_BYTES = BM * BK * 2, B_KB_BYTES = HEAD_BN * BK * 2;
    const int KB = p.kblocks;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_base = base;
    const uint32_t b_base = a_base + KB * A_KB_BYTES;
    const uint32_t bias_base = b_base + HEAD_STAGES * KB * B_KB_BYTES;
    const uint32_t bar_base = bias_base + 2 * HEAD_BN * 4;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (HEAD_STAGES + s); };
    auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * HEAD_STAGES + a); };
    auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * HEAD_STAGES + 2 + a); };
    const uint32_t a_bar = bar_base + 8u * (2 * HEAD_STAGES + 4);
    const uint32_t tmem_slot = bar_base + 8u * (2 * HEAD_STAGES + 5);
    uint32_t* tmem_slot_ptr = (uint32_t*)(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    float* bias_sm = (float*)(smem_raw + (bias_base - smem_u32(smem_raw)));

    const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM;
    const int t_begin = blockIdx.y * p.tiles_per_split;
    const int t_end = min(p.n_tiles, t_begin + p.tiles_per_split);

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < HEAD_STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tmem_full_bar(a), 1);
            mbar_init(tmem_empty_bar(a), 4);
        }
        mbar_init(a_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)(2 * HEAD_BN)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_wait();                                             // activations of the previous kernel are complete and visible

    if (warp == 0) {
        if (elect_one_sync()) {
            mbar_expect_tx(a_bar, KB * A_KB_BYTES);
            for (int kb = 0; kb < KB; ++kb) tma_load_2d(a_base + kb * A_KB_BYTES, &map_a, a_bar, kb * BK, m0);
            uint32_t it = 0;
            for (int tile = t_begin; tile < t_end; ++tile, ++it) {
                int s = it % HEAD_STAGES;
                uint32_t ph = (it / HEAD_STAGES) & 1;
                mbar_wait(empty_bar(s), ph ^ 1);
                mbar_expect_tx(full_bar(s), KB * B_KB_BYTES);
                for (int kb = 0; kb < KB; ++kb)
                    tma_load_2d(b_base + (s * KB + kb) * B_KB_BYTES, &map_b, full_bar(s), kb * BK, tile * HEAD_BN);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (elect_one_sync()) {
            constexpr uint32_t idesc = make_idesc(HEAD_BN);
            mbar_wait(a_bar, 0);
            uint32_t it = 0;
            for (int tile = t_begin; tile < t_end; ++tile, ++it) {
                int acc = it & 1;
                uint32_t aph = (it >> 1) & 1;
                mbar_wait(tmem_empty_bar(acc), aph ^ 1);
                int s = it % HEAD_STAGES;
                uint32_t ph = (it / HEAD_STAGES) & 1;
                mbar_wait(full_bar(s), ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t tmem_d = tmem_base + (uint32_t)(acc * HEAD_BN);
                for (int kb = 0; kb < KB; ++kb) {
                    uint64_t adesc = make_smem_desc(a_base + kb * A_KB_BYTES);
                    uint64_t bdesc = make_smem_desc(b_base + (s * KB + kb) * B_KB_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        umma_bf16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
                }
                umma_commit(empty_bar(s));
                umma_commit(tmem_full_bar(acc));
            }
        }
        __syncwarp();
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int et = threadIdx.x - 64;                 // 0..127 among the epilogue threads
        float run_max = -3.0e38f, run_sum = 0.f;
        int run_idx = 0;
        uint32_t it = 0;
        for (int tile = t_begin; tile < t_end; ++tile, ++it) {
            int acc = it & 1;
            uint32_t aph = (it >> 1) & 1;
            bias_sm[acc * HEAD_BN + et] = bias[tile * HEAD_BN + et];
            named_bar_sync(1, 128);                      // bias tile visible; also fences the reuse of the other buffer
            mbar_wait(tmem_full_bar(acc), aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t trow = tmem_base + (uint32_t)(acc * HEAD_BN) + ((uint32_t)(q * 32) << 16);
            const float* bs = bias_sm + acc * HEAD_BN;
#pragma unroll 1
            for (int c = 0; c < HEAD_BN; c += 32) {
                uint32_t v[32];
                tmem_ld8(trow + (uint32_t)c, v);
                tmem_ld8(trow + (uint32_t)c + 8, v + 8);
                tmem_ld8(trow + (uint32_t)c + 16, v + 16);
                tmem_ld8(trow + (uint32_t)c + 24, v + 24);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float f[32];
                float cmax = -3.0e38f;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 b4 = *reinterpret_cast<const float4*>(bs + c + j);
                    f[j] = __uint_as_float(v[j]) + b4.x;
                    f[j + 1] = __uint_as_float(v[j + 1]) + b4.y;
                    f[j + 2] = __uint_as_float(v[j + 2]) + b4.z;
                    f[j + 3] = __uint_as_float(v[j + 3]) + b4.w;
                    cmax = fmaxf(cmax, fmaxf(fmaxf(f[j], f[j + 1]), fmaxf(f[j + 2], f[j + 3])));
                }
                if (cmax > run_max) {                     // strictly greater: the earliest maximum keeps its index
                    int ci = 0;
#pragma unroll
                    for (int j = 31; j >= 0; --j)
                        if (f[j] == cmax) ci = j;
                    run_sum *= __expf(run_max - cmax);
                    run_max = cmax;
                    run_idx = tile * HEAD_BN + c + ci;
                }
                float ssum = 0.f;
#pragma unroll
                for (int j = 0; j < 32; ++j) ssum += __expf(f[j] - run_max);
                run_sum += ssum;
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
        }
        if (m0 + row < p.rows) {
            size_t o = (size_t)blockIdx.y * p.rows + m0 + row;
            pmax[o] = run_max;
            pidx[o] = run_idx;
            psum[o] = run_sum;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * HEAD_BN)) : "memory");
    }
}

// Merges the class splits of ctc_head_kernel: idx = first index of the global maximum, prob = softmax value there.
__global__ void ctc_head_merge_kernel(const float* __restrict__ pmax, const int* __restrict__ pidx,
                                      const float* __restrict__ psum, int splits, int rows, int* __restrict__ idx,
                                      float* __restrict__ prob) {
    pdl_start();
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    float m = pmax[r];
    int bi = pidx[r];
    for (int s = 1; s < splits; ++s) {
        float v = pmax[(size_t)s * rows + r];
        int i = pidx[(size_t)s * rows + r];
        if (v > m || (v == m && i < bi)) { m = v; bi = i; }
    }
    float sum = 0.f;
    for (int s = 0; s < splits; ++s) sum += psum[(size_t)s * rows + r] * __expf(pmax[(size_t)s * rows + r] - m);
    idx[r] = bi;
    prob[r] = 1.f / sum;
}

}  // namespace tc
}  // namespace b200ocr

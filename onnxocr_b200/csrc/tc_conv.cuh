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
    int Cout, CoutP;             // GEMM columns that carry data / padded (for convT: 4 * out channels padded)
    int convT;                   // 1: column = tap * ctC + co, stored pixel-shuffled to (2y + dy, 2x + dx)
    int ctC, ctCout;             // convT: padded / real output channels
    Epilogue epi;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

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

template <int BN>
constexpr size_t smem_bytes() {
    return (size_t)STAGES * (BM * BK * 2 + BN * BK * 2) + 1024 /*alignment slack*/ + 256 /*barriers*/;
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

// Persistent kernel: each CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...; the TMA ring runs ahead across tile
// boundaries and two TMEM accumulators let the MMA of tile t+1 overlap the epilogue of tile t.
template <int BN>
__global__ void __launch_bounds__(kThreads) conv_tc_kernel(const __grid_constant__ CUtensorMap map_a,
                                                           const __grid_constant__ CUtensorMap map_b,
                                                           const float* __restrict__ bias, bf16* __restrict__ y, const Params p) {
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

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
        if (lane == 0) {
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
        if (lane == 0) {
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
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle row: +2 in the (addr >> 4) field
                        umma_bf16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
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

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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

    if (warp == 0) {
        if (lane == 0) {
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
        if (lane == 0) {
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

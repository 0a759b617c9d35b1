// Microbenchmark: cycles per tcgen05.mma (M = 128, K = 16, bf16, SS operands) as a function of N, of the A descriptor's
// stride (canonical 1024-byte core-group pitch vs the 1280-byte halo pitch), and of accumulator interleaving.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I onnxocr_b200/csrc -o /tmp/mma_rate tools/ubench/mma_rate.cu
#include <cstdio>
#include <vector>
#include "tc_conv.cuh"
using namespace b200ocr;
using namespace b200ocr::tc;

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0, laneid = 0;
    asm volatile("{\n.reg .b32 %%rx;\n.reg .pred %%px;\nelect.sync %%rx|%%px, %2;\n@%%px mov.s32 %1, 1;\nmov.s32 %0, %%rx;\n}\n"
                 : "+r"(laneid), "+r"(pred) : "r"(0xFFFFFFFFu));
    return pred != 0;
}

__global__ void __launch_bounds__(512) mma_rate_kernel(int N, int nmma, int reps, int sbo, int nacc, int a_span, int spin, int commit_every, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint64_t bar2;
    __shared__ uint64_t bar3;
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)(smem_raw + (base - smem_u32(smem_raw))))[i] = 0;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) { mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&bar2), 1); mbar_init(smem_u32(&bar3), 1u << 19); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_slot;
    if (warp == 1 && elect_one()) {
        const uint32_t idesc = make_idesc(N);
        const uint32_t a_base = base, b_base = base + 96 * 1024;
        long long best = 1ll << 60;
        uint32_t ph = 0;
        for (int r = 0; r < reps; ++r) {
            // descriptors are loop invariant (4 K-slices of two A tiles / two B tiles): the loop body is 8 bare MMAs, so the
            // measurement is the tensor pipe's rate, not this thread's address arithmetic (first version: 216 cycles per
            // MMA for every N, all of it integer division in the issuing thread)
            uint64_t ad[8], bd[8];
            for (int j = 0; j < 8; ++j) {
                // a_span < 0: emulate the 3x3 halo taps — start addresses only 128-byte aligned ((ky * 10 + kx) * 128)
                ad[j] = a_span < 0 ? make_smem_desc_sbo(a_base + (uint32_t)((j >> 2) * 10 + 1 + (j >> 2)) * 128u, (uint32_t)sbo) + (uint64_t)(2 * (j & 3))
                                   : make_smem_desc_sbo(a_base + (uint32_t)(j >> 2) * (uint32_t)a_span, (uint32_t)sbo) + (uint64_t)(2 * (j & 3));
                bd[j] = make_smem_desc(b_base + (uint32_t)(j >> 2) * 32768u) + (uint64_t)(2 * (j & 3));
            }
            const uint32_t d0 = tmem_base, d1 = tmem_base + (uint32_t)((nacc - 1) * N);
            long long t0 = clock64();
            for (int i = 0; i < nmma; i += 8) {
#pragma unroll
                for (int j = 0; j < 8; ++j) umma_bf16((j & 1) ? d1 : d0, ad[j], bd[j], idesc, 1u);
                if (commit_every && ((i + 8) % commit_every) == 0) umma_commit(smem_u32(&bar3));   // as the stage-release commits do
            }
            umma_commit(smem_u32(&bar));
            mbar_wait(smem_u32(&bar), ph);
            ph ^= 1;
            long long t1 = clock64();
            if (t1 - t0 < best) best = t1 - t0;
        }
        out[blockIdx.x] = best;
        mbar_arrive(smem_u32(&bar2));
    } else if (warp >= 2 && warp < 2 + spin) {
        // waiting warps: poll a barrier that completes only when the MMA thread is done (the epilogue warps of the real
        // kernels wait for an accumulator this way)
        mbar_wait(smem_u32(&bar2), 0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

int main() {
    long long* d;
    cudaMalloc(&d, 148 * 2 * sizeof(long long));
    const int smem = 200 * 1024;
    cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int nmma = 512;
    printf("N  sbo  nacc a_span ctas/sm : cycles per MMA (min over CTAs / max over CTAs)\n");
    for (int N : {32, 64, 128})
        for (int sbo : {1024})
            for (int nacc : {1})
                for (int spin : {0, 8, 16, 32, 64}) { const int a_span = 16384;
                    if (nacc * N > 512) continue;
                    mma_rate_kernel<<<148, 512, smem>>>(N, nmma, 5, sbo, nacc, a_span, 0, spin, d);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    std::vector<long long> h(148);
                    cudaMemcpy(h.data(), d, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
                    long long mn = 1ll << 60, mx = 0;
                    for (auto v : h) { mn = v < mn ? v : mn; mx = v > mx ? v : mx; }
                    printf("%3d %4d %d commit_every %2d : %.1f / %.1f\n", N, sbo, nacc, spin, (double)mn / nmma, (double)mx / nmma);
                }
    return 0;
}

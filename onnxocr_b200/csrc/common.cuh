// Shared device helpers: error checks, storage-type traits (fp32 | bf16), epilogue evaluation.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <utility>
#include <stdexcept>
#include <string>

#include "graph.h"

namespace b200ocr {

#define B200_CUDA_CHECK(expr)                                                                        \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " + \
                                     __FILE__ + ":" + std::to_string(__LINE__) + " (" #expr ")");    \
    } while (0)

inline int cpad_of(int c) { return c == 1 ? 1 : (c <= 4 ? 4 : (c + 7) / 8 * 8); }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

typedef __nv_bfloat16 bf16;

template <typename T> struct ST;   // storage traits
template <> struct ST<float> {
    static __device__ __forceinline__ float ld(const float* p) { return *p; }
    static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
    static __device__ __forceinline__ void ld4(const float* p, float* v) {
        float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ void st4(float* p, const float* v) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <> struct ST<bf16> {
    static __device__ __forceinline__ float ld(const bf16* p) { return __bfloat162float(*p); }
    static __device__ __forceinline__ void st(bf16* p, float v) { *p = __float2bfloat16_rn(v); }
    static __device__ __forceinline__ void ld4(const bf16* p, float* v) {
        uint2 t = *reinterpret_cast<const uint2*>(p);
        v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
        v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
    }
    static __device__ __forceinline__ void st4(bf16* p, const float* v) {
        __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
        __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
        uint2 t;
        t.x = *reinterpret_cast<uint32_t*>(&a);
        t.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(p) = t;
    }
};

// Programmatic dependent launch (PDL): every network kernel lets its successor's CTAs be scheduled while it is still
// running (launch_dependents at entry) and itself waits for its predecessor to complete and flush before it touches
// activation memory (wait).  With ~80 short kernels per plan this hides the launch / drain latency between them.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_start() { pdl_trigger(); pdl_wait(); }

// Measured on B200 (tools/time_plans.py, CUDA-graph replay): no difference within noise for any plan (det 16 pages
// 2.565 vs 2.576 ms, rec 6 crops 0.596 vs 0.587 ms) — the ~9 us per small kernel is its own serial TMA -> MMA -> epilogue
// latency, not the gap between kernels.  Off unless B200OCR_PDL=1; without the launch attribute both instructions are
// no-ops.
inline bool pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("B200OCR_PDL");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// Channel sums for the SE gate, accumulated by the producing kernel (see se_fc_sums_kernel): 48.16 fixed point, 64-bit
// integer atomics — associative, hence independent of the order in which tiles finish.
constexpr float kGapFixScale = 65536.f;
__device__ __forceinline__ void gap_fix_add(unsigned long long* dst, float v) {
    if (v != 0.f) atomicAdd(dst, (unsigned long long)__float2ll_rn(v * kGapFixScale));
}

__device__ __forceinline__ float epi_act(float v, int type, float a, float b) {
    switch (type) {
        case EPI_RELU: return fmaxf(v, 0.f);
        case EPI_HSWISH: return v * fminf(fmaxf(v + 3.f, 0.f), 6.f) * 0.16666667f;
        case EPI_SIGMOID: return 1.f / (1.f + expf(-v));
        case EPI_SWISH: return v / (1.f + expf(-v));
        case EPI_HSIGMOID: return fminf(fmaxf(fmaf(v, a, b), 0.f), 1.f);
        case EPI_CLIP: return fminf(fmaxf(v, a), b);
        default: return v;
    }
}

__device__ __forceinline__ float epi_apply(float v, const Epilogue& e) {
    if (e.n == 0) return v;
    if (e.fast) {
        v = fmaf(v, e.a0, e.b0);
        v = epi_act(v, e.act, e.pa, e.pb);
        return fmaf(v, e.a1, e.b1);
    }
#pragma unroll 1
    for (int i = 0; i < e.n; ++i) {
        const EpiStep s = e.s[i];
        v = s.type == EPI_AFFINE ? fmaf(v, s.a, s.b) : epi_act(v, s.type, s.a, s.b);
    }
    return v;
}

// Compiled epilogue with the activation fixed at compile time (ACT < 0: interpret the chain).
template <int ACT>
__device__ __forceinline__ float epi_fast(float v, const Epilogue& e) {
    if (ACT < 0) return epi_apply(v, e);                 // generic (interpreted) chain
    v = fmaf(v, e.a0, e.b0);
    if (ACT == EPI_RELU) v = fmaxf(v, 0.f);
    else if (ACT == EPI_HSWISH) v = v * fminf(fmaxf(v + 3.f, 0.f), 6.f) * 0.16666667f;
    else if (ACT == EPI_SIGMOID) v = __fdividef(1.f, 1.f + __expf(-v));
    else if (ACT == EPI_SWISH) v = __fdividef(v, 1.f + __expf(-v));
    else if (ACT == EPI_HSIGMOID) v = fminf(fmaxf(fmaf(v, e.pa, e.pb), 0.f), 1.f);
    else if (ACT == EPI_CLIP) v = fminf(fmaxf(v, e.pa), e.pb);
    return fmaf(v, e.a1, e.b1);
}

}  // namespace b200ocr

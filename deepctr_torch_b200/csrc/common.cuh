// Shared helpers for the sm_100a kernels behind libctr_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "ctr_b200.h"

void ctr_set_error(const char* fmt, ...);
int  ctr_sm_count();

#define CTR_ARG(cond, ...)                          \
    do {                                            \
        if (!(cond)) {                              \
            ctr_set_error(__VA_ARGS__);             \
            return -1;                              \
        }                                           \
    } while (0)

#define CTR_CUDA(expr)                                                              \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) {                                                    \
            ctr_set_error("%s failed: %s", #expr, cudaGetErrorString(_e));          \
            return (int)_e;                                                         \
        }                                                                           \
    } while (0)

void ctr_count_launch();

#define CTR_LAUNCH_OK(name)                                                         \
    do {                                                                            \
        ctr_count_launch();                                                         \
        cudaError_t _e = cudaGetLastError();                                        \
        if (_e != cudaSuccess) {                                                    \
            ctr_set_error("launch of %s failed: %s", name, cudaGetErrorString(_e)); \
            return (int)_e;                                                         \
        }                                                                           \
    } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// streaming 128-bit load that does not allocate in L1 (rows are touched once per kernel)
__device__ __forceinline__ float4 ld_stream4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ void st_stream4(float* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
                 "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

// 128-bit vector reduction to global memory (sm_90+): one request adds four fp32 lanes
__device__ __forceinline__ void red_add4(float* p, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
}

// (if chains, not switches, in both helpers: see act_grad_from_y)
__device__ __forceinline__ float act_apply(int act, float z) {
    if (act == CTR_ACT_RELU) return z > 0.f ? z : 0.f;
    if (act == CTR_ACT_LINEAR) return z;
    if (act == CTR_ACT_TANH) return tanhf(z);
    if (act == CTR_ACT_SIGMOID) return 1.f / (1.f + expf(-z));
    return z;
}

// derivative of the activation expressed through its OUTPUT y
// (an if chain on purpose: the switch became a jump table — one indirect branch per ELEMENT inside the GEMM
// converters' unrolled loops, ~150 cycles each: profiles/r02_raw/tsw_tl_dw2.log before / after)
__device__ __forceinline__ float act_grad_from_y(int act, float y) {
    if (act == CTR_ACT_RELU) return y > 0.f ? 1.f : 0.f;
    if (act == CTR_ACT_TANH) return 1.f - y * y;
    if (act == CTR_ACT_SIGMOID) return y * (1.f - y);
    if (act == 100) return y;          // internal: plain multiply by the "mask" operand (GEMM prologue A(m,k) *= aux(m,k))
    return 1.f;
}

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

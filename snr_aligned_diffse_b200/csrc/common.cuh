// Shared device/host helpers for the sm_100a kernels of the enhancement hot path.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

typedef __nv_bfloat16 bf16;

#define SNRSE_OK 0
#define SNRSE_ERR_ARG 1
#define SNRSE_ERR_CUDA 2
#define SNRSE_ERR_STATE 3
#define SNRSE_ERR_UNSUPPORTED 4

// Thread-local error message, readable through snrse_last_error().
void snrse_set_error(const char* fmt, ...);

#define SNRSE_CHECK_ARG(cond, ...)            \
    do {                                      \
        if (!(cond)) {                        \
            snrse_set_error(__VA_ARGS__);     \
            return SNRSE_ERR_ARG;             \
        }                                     \
    } while (0)

#define SNRSE_CUDA(call)                                                                      \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            snrse_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return SNRSE_ERR_CUDA;                                                            \
        }                                                                                     \
    } while (0)

// every kernel launch in the library goes through this macro: it also counts launches (snrse_launch_count)
extern long long g_snrse_launches;
#define SNRSE_LAUNCH_CHECK()          \
    do {                              \
        ++g_snrse_launches;           \
        SNRSE_CUDA(cudaGetLastError()); \
    } while (0)

#define SNRSE_TRY(call)              \
    do {                             \
        int rc__ = (call);           \
        if (rc__ != SNRSE_OK) return rc__; \
    } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

#ifdef __CUDACC__
// SiLU through ONE special-function op: x*sigmoid(x) = h + h*tanh(h), h = x/2 (tanh.approx.f32, rel. error
// ~2^-11, below the bf16 rounding of the stored result).  The exp+rcp form needs two SFU ops per element and
// made the GroupNorm+SiLU pass SFU-bound instead of HBM-bound (16 SFU lanes/clk/SM).
__device__ __forceinline__ float silu_f(float x) {
    const float h = 0.5f * x;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
}

__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
    const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 t = __bfloat1622float2(p[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}

__device__ __forceinline__ uint4 pack8(const float* f) {
    uint4 v;
    __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
#endif

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

// Launch helper of every kernel in the library.  With g_snrse_pdl != 0 the launch carries the programmatic-stream-
// serialization attribute (programmatic dependent launch): the grid may become resident while its predecessor in the
// stream (or captured graph) is still draining, runs its prologue, and blocks in pdl_sync() / griddepcontrol.wait until
// the predecessor has completed and flushed.  EVERY kernel of the library executes griddepcontrol.wait before its first
// access to memory another kernel may have written and before its own first global write, so completion stays transitive.
extern int g_snrse_pdl;
#ifdef __CUDACC__
// g_snrse_pdl is a mask: bit0 = the memory-bound / small kernels, bit1 = the two tcgen05 convolution kernels
template <typename... KArgs, typename... Args>
static inline void snrse_launch_m(int mask, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (g_snrse_pdl & mask) ? 1 : 0;
    (void)cudaLaunchKernelEx(&cfg, kern, static_cast<Args&&>(args)...);   // errors surface in SNRSE_LAUNCH_CHECK (cudaGetLastError)
}
template <typename... KArgs, typename... Args>
static inline void snrse_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    snrse_launch_m(1, kern, grid, block, smem, s, static_cast<Args&&>(args)...);
}
#endif

#define SNRSE_TRY(call)              \
    do {                             \
        int rc__ = (call);           \
        if (rc__ != SNRSE_OK) return rc__; \
    } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

#ifdef __CUDACC__
// Programmatic dependent launch, device side.  pdl_trigger(): the next kernel of the stream may start becoming resident
// (it still blocks in its own pdl_wait()).  pdl_wait(): the predecessor grid has completed and its writes are visible;
// a no-op when the launch did not carry the attribute.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() { pdl_trigger(); pdl_wait(); }

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

// Correctness probe (measurement tool, not product code): can a K-major SWIZZLE_128B UMMA operand start at an
// arbitrary 128-byte row of a TMA-written tile, with an 8-row-group stride (SBO) that is not 1024 bytes?
// If yes, a 3x3 convolution needs ONE halo tile per 64-channel chunk ([rows+2][8+2] pixels, tap (r,s) = start
// row (r*10+s), SBO = 10 rows) instead of three column-shifted copies.
//
// A: R rows x 64 bf16 (row i, col k = small integers), loaded by TMA (128B swizzle) to a 1024-aligned tile.
// B: 64 x 64 identity  ->  D[m][n] = A[row(m)][n],  row(m) = shift + (m/8)*sbo_rows + m%8.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I snr_aligned_diffse_b200/csrc tools/umma_shift_probe.cu -o tools/umma_shift_probe -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "ptx.cuh"

#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int R = 384;   // rows of the A tile in shared memory (48 KB)

__global__ void __launch_bounds__(128, 1)
shift_probe(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int shift, int sbo_rows,
            int use_base_offset, float* out) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar, done_bar;
    __shared__ uint32_t tmem_base_smem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_base = base, b_base = base + R * 128;
    if (threadIdx.x == 0) {
        ptx::mbar_init(ptx::smem_u32(&full_bar), 1);
        ptx::mbar_init(ptx::smem_u32(&done_bar), 1);
        ptx::fence_barrier_init();
    }
    __syncthreads();
    if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(&tmem_base_smem), 64); ptx::tmem_relinquish(); }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    if (threadIdx.x == 0) {
        ptx::mbar_arrive_expect_tx(ptx::smem_u32(&full_bar), R * 128 + 64 * 128);
        // TMA boxes are limited to 256 rows: two loads for A
        ptx::tma_load_3d(a_base, &mapA, ptx::smem_u32(&full_bar), 0, 0, 0);
        ptx::tma_load_3d(a_base + 192 * 128, &mapA, ptx::smem_u32(&full_bar), 0, 192, 0);
        ptx::tma_load_3d(b_base, &mapB, ptx::smem_u32(&full_bar), 0, 0, 0);
        ptx::mbar_wait(ptx::smem_u32(&full_bar), 0);
        ptx::tc_fence_after();
        const uint32_t start = a_base + (uint32_t)shift * 128u;
        uint64_t da = 0;
        da |= (uint64_t)((start >> 4) & 0x3FFF);
        da |= (uint64_t)1 << 16;
        da |= (uint64_t)((sbo_rows * 128) >> 4) << 32;
        da |= (uint64_t)1 << 46;
        if (use_base_offset) da |= (uint64_t)((start >> 7) & 7) << 49;
        da |= (uint64_t)2 << 61;
        const uint64_t db = ptx::umma_desc_k_sw128(b_base);
        const uint32_t idesc = ptx::umma_idesc_bf16(128, 64);
        for (int k = 0; k < 4; ++k) ptx::mma_bf16_ss(tmem_base, da + 2 * k, db + 2 * k, idesc, k > 0 ? 1u : 0u);
        ptx::mma_commit(ptx::smem_u32(&done_bar));
    }
    __syncthreads();
    ptx::mbar_wait(ptx::smem_u32(&done_bar), 0);
    ptx::tc_fence_after();
    uint32_t v[32];
    for (int c0 = 0; c0 < 64; c0 += 32) {
        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
        ptx::tmem_ld_wait();
        for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c0 + j] = __uint_as_float(v[j]);
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc(tmem_base, 64);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static void make_map(EncodeTiledFn enc, CUtensorMap* m, void* ptr, int rows, int box_rows) {
    cuuint64_t dims[3] = {64, (cuuint64_t)rows, 1};
    cuuint64_t strides[2] = {128, (cuuint64_t)rows * 128};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
}

int main() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fn);
    std::vector<__nv_bfloat16> hA(R * 64), hB(64 * 64);
    for (int i = 0; i < R; ++i)
        for (int k = 0; k < 64; ++k) hA[i * 64 + k] = __float2bfloat16((float)((i * 7 + k * 3) % 251) - 125.f);
    for (int n = 0; n < 64; ++n)
        for (int k = 0; k < 64; ++k) hB[n * 64 + k] = __float2bfloat16(n == k ? 1.f : 0.f);
    __nv_bfloat16 *dA, *dB;
    float* dOut;
    CK(cudaMalloc(&dA, hA.size() * 2));
    CK(cudaMalloc(&dB, hB.size() * 2));
    CK(cudaMalloc(&dOut, 128 * 64 * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    CUtensorMap mA, mB;
    make_map(enc, &mA, dA, R, 192);
    make_map(enc, &mB, dB, 64, 64);
    const int smem = R * 128 + 64 * 128 + 1024;
    CK(cudaFuncSetAttribute(shift_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    std::vector<float> h(128 * 64);
    const int cases[][2] = {{0, 8}, {1, 8}, {3, 8}, {8, 8}, {0, 10}, {1, 10}, {2, 10}, {11, 10}, {21, 10}, {5, 18}, {0, 16}, {7, 9}};
    for (auto& c : cases) {
        for (int ubo = 0; ubo < 2; ++ubo) {
            CK(cudaMemset(dOut, 0, 128 * 64 * 4));
            shift_probe<<<1, 128, smem>>>(mA, mB, c[0], c[1], ubo, dOut);
            CK(cudaGetLastError());
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h.data(), dOut, h.size() * 4, cudaMemcpyDeviceToHost));
            int bad = 0, first_bad = -1;
            for (int m = 0; m < 128; ++m) {
                const int row = c[0] + (m / 8) * c[1] + m % 8;
                for (int n = 0; n < 64; ++n) {
                    const float want = __bfloat162float(hA[row * 64 + n]);
                    if (h[m * 64 + n] != want) { ++bad; if (first_bad < 0) first_bad = m * 64 + n; }
                }
            }
            printf("{\"shift_rows\": %d, \"sbo_rows\": %d, \"base_offset_field\": %d, \"mismatches\": %d, \"first_bad\": %d}\n", c[0],
                   c[1], ubo, bad, first_bad);
        }
    }
    return 0;
}

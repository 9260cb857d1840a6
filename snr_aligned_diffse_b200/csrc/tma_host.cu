#include "tma_host.h"

#include <cudaTypedefs.h>
#include <mutex>

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
std::once_flag g_once;

int get_encode(EncodeTiledFn* fn) {
    std::call_once(g_once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess) {
            g_encode = reinterpret_cast<EncodeTiledFn>(p);
        }
    });
    if (!g_encode) {
        snrse_set_error("cuTensorMapEncodeTiled is not available from this driver");
        return SNRSE_ERR_CUDA;
    }
    *fn = g_encode;
    return SNRSE_OK;
}
}  // namespace

int tma_make_act_map(CUtensorMap* map, const bf16* ptr, int C, int W, int H, int B, int ld, int box_c, int box_w,
                     int box_h) {
    EncodeTiledFn enc;
    SNRSE_TRY(get_encode(&enc));
    SNRSE_CHECK_ARG((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA: activation pointer must be 16-byte aligned");
    SNRSE_CHECK_ARG(box_c * 2 == 128 && box_w <= 256 && box_h <= 256, "TMA: bad activation box");
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * W, (cuuint64_t)ld * 2 * W * H};
    cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        snrse_set_error("cuTensorMapEncodeTiled(act C=%d W=%d H=%d B=%d ld=%d box %dx%dx%d) failed: %d", C, W, H, B, ld,
                        box_c, box_w, box_h, (int)r);
        return SNRSE_ERR_CUDA;
    }
    return SNRSE_OK;
}

int tma_make_wt_map(CUtensorMap* map, const bf16* ptr, int64_t K, int64_t rows, int64_t batch, int64_t batch_stride,
                    int box_k, int box_rows, int64_t row_pitch) {
    if (row_pitch <= 0) row_pitch = K;
    EncodeTiledFn enc;
    SNRSE_TRY(get_encode(&enc));
    SNRSE_CHECK_ARG((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA: weight pointer must be 16-byte aligned");
    SNRSE_CHECK_ARG(box_k * 2 == 128 && box_rows <= 256 && K % 8 == 0 && batch_stride % 8 == 0 && row_pitch % 8 == 0,
                    "TMA: bad weight box");
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)row_pitch * 2, (cuuint64_t)batch_stride * 2};
    cuuint32_t box[3] = {(cuuint32_t)box_k, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<bf16*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        snrse_set_error("cuTensorMapEncodeTiled(wt K=%lld rows=%lld batch=%lld) failed: %d", (long long)K,
                        (long long)rows, (long long)batch, (int)r);
        return SNRSE_ERR_CUDA;
    }
    return SNRSE_OK;
}

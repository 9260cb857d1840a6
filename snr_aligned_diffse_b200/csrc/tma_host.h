// Host-side construction of TMA tensor maps (cuTensorMapEncodeTiled resolved through the runtime,
// so the library does not link libcuda directly).
#pragma once
#include <cuda.h>

#include "common.cuh"

// NHWC bf16 activation view as a 4-D tensor {C, W, H, B} with pixel pitch `ld` elements;
// box {box_c, box_w, box_h, 1}, 128-byte swizzle, zero fill outside the image.
int tma_make_act_map(CUtensorMap* map, const bf16* ptr, int C, int W, int H, int B, int ld, int box_c, int box_w,
                     int box_h);
// K-major bf16 matrix batch {K, rows, batch}; box {box_k, box_rows, 1}, 128-byte swizzle.  row_pitch: elements
// between rows (0: K, i.e. dense rows).
int tma_make_wt_map(CUtensorMap* map, const bf16* ptr, int64_t K, int64_t rows, int64_t batch, int64_t batch_stride,
                    int box_k, int box_rows, int64_t row_pitch = 0);

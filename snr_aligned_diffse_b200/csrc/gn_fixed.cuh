// Fixed-point encoding of GroupNorm partial sums (shared by norm.cu and the convolution epilogue in conv_halo2.cu).
// A partial (<= a few thousand bf16-range elements, computed in fp32) is rounded to a 64-bit integer and accumulated
// with integer atomics: exact and order-independent, so statistics do not depend on scheduling or batch composition.
//   sum    : 2^-30 resolution, |total| < 8.6e9 per 4-channel unit
//   sum sq : 2^-20 resolution,  total  < 8.8e12 per 4-channel unit
// Supported activation range (stated): a 4-channel unit of a 60 s utterance at full resolution holds
// 256 x 7552 x 4 = 7.7e6 elements, so the sum of squares stays representable up to an rms of ~1000 per unit (the network's
// activations are O(1..100); r01's 2^-24 scale gave ~270, or ~130 after gn_finalize added four units in integers); a
// partial's rounding error (2^-21 per <= 128 elements) moves the variance by <= 4e-9, far below eps = 1e-6, also for
// tensors whose values are of the order of sqrt(eps).  Conversions saturate (cvt.rni.s64.f32 clamps), and gn_finalize treats a unit whose accumulated
// square sum has crossed 2^63 as saturated (variance -> huge, output -> shift only) instead of letting it wrap negative.
#pragma once
#include <stdint.h>

#define GN_FIX_SUM_SCALE 1073741824.0f      /* 2^30 */
#define GN_FIX_SQ_SCALE 1048576.0f          /* 2^20 */

__device__ __forceinline__ unsigned long long gn_fix_sum(float v) { return (unsigned long long)__float2ll_rn(v * GN_FIX_SUM_SCALE); }
__device__ __forceinline__ unsigned long long gn_fix_sq(float v) { return (unsigned long long)__float2ll_rn(v * GN_FIX_SQ_SCALE); }
__device__ __forceinline__ double gn_unfix_sum(long long v) { return (double)v * (1.0 / 1073741824.0); }
__device__ __forceinline__ double gn_unfix_sq(unsigned long long v) {
    if (v >> 63) v = 1ull << 63;            // wrapped past the signed range: saturate
    return (double)v * (1.0 / 1048576.0);
}

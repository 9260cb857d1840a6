// FIR x2 resampling with the [1,3,3,1] (x) [1,3,3,1] / 64 kernel, NHWC.
// Replaces `upsample_2d` / `downsample_2d` (sgmse-bbed/sgmse/backbones/ncsnpp_utils/up_or_down_sampling.py:195-257)
// and the reference's own CUDA op `upfirdn2d_kernel` modes 3 and 5
// (ncsnpp_utils/op/upfirdn2d_kernel.cu:107-207,264-283) for exactly the two configurations NCSN++ uses:
//   up  : zero-insert x2, pad (2,1), gain 4  ->  per axis  out[2i]   = (x[i-1] + 3 x[i]) / 4
//                                                           out[2i+1] = (3 x[i] + x[i+1]) / 4
//   down: pad (1,1), keep every 2nd sample   ->  per axis  out[j] = (x[2j-1] + 3 x[2j] + 3 x[2j+1] + x[2j+2]) / 8
// with zeros outside the image.  One thread produces 8 channels (bf16) or one 4-channel pixel (fp32).
#include "kernels.h"

namespace {

struct V8 {
    float f[8];
};
__device__ __forceinline__ V8 ld8(const bf16* p) {
    V8 r;
    unpack8(__ldg(reinterpret_cast<const uint4*>(p)), r.f);
    return r;
}

__global__ void __launch_bounds__(256)
fir_down2_kernel(const bf16* __restrict__ x, int ld, int C, int H, int W, bf16* __restrict__ out, int out_ld,
                 int64_t total) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int tpp = C >> 3;
    const int c0 = (int)(idx % tpp) * 8;
    int64_t pix = idx / tpp;
    const int Ho = H >> 1, Wo = W >> 1;
    const int wo = (int)(pix % Wo);
    pix /= Wo;
    const int ho = (int)(pix % Ho);
    const int b = (int)(pix / Ho);
    const float k[4] = {0.125f, 0.375f, 0.375f, 0.125f};
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int hh = 2 * ho - 1 + a;
        if (hh < 0 || hh >= H) continue;
#pragma unroll
        for (int bb = 0; bb < 4; ++bb) {
            const int ww = 2 * wo - 1 + bb;
            if (ww < 0 || ww >= W) continue;
            const V8 v = ld8(x + (((int64_t)b * H + hh) * W + ww) * ld + c0);
            const float kw = k[a] * k[bb];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(kw, v.f[j], acc[j]);
        }
    }
    *reinterpret_cast<uint4*>(out + (((int64_t)b * Ho + ho) * Wo + wo) * out_ld + c0) = pack8(acc);
}

__global__ void __launch_bounds__(256)
fir_up2_kernel(const bf16* __restrict__ x, int ld, int C, int H, int W, bf16* __restrict__ out, int out_ld,
               int64_t total) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int tpp = C >> 3;
    const int c0 = (int)(idx % tpp) * 8;
    int64_t pix = idx / tpp;
    const int Ho = H * 2, Wo = W * 2;
    const int wo = (int)(pix % Wo);
    pix /= Wo;
    const int ho = (int)(pix % Ho);
    const int b = (int)(pix / Ho);
    // two contributing source rows / columns with weights (1/4, 3/4) or (3/4, 1/4)
    const int hi = ho >> 1, wi = wo >> 1;
    const int ha = (ho & 1) ? hi : hi - 1, hb = ha + 1;
    const int wa = (wo & 1) ? wi : wi - 1, wb = wa + 1;
    const float kha = (ho & 1) ? 0.75f : 0.25f, khb = 1.0f - kha;
    const float kwa = (wo & 1) ? 0.75f : 0.25f, kwb = 1.0f - kwa;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int hs[2] = {ha, hb};
    const int ws[2] = {wa, wb};
    const float kh[2] = {kha, khb};
    const float kw[2] = {kwa, kwb};
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        if (hs[a] < 0 || hs[a] >= H) continue;
#pragma unroll
        for (int bb = 0; bb < 2; ++bb) {
            if (ws[bb] < 0 || ws[bb] >= W) continue;
            const V8 v = ld8(x + (((int64_t)b * H + hs[a]) * W + ws[bb]) * ld + c0);
            const float kk = kh[a] * kw[bb];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(kk, v.f[j], acc[j]);
        }
    }
    *reinterpret_cast<uint4*>(out + (((int64_t)b * Ho + ho) * Wo + wo) * out_ld + c0) = pack8(acc);
}

__global__ void __launch_bounds__(256)
fir_down2_f4_kernel(const float4* __restrict__ x, int H, int W, float4* __restrict__ out, int64_t total) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int Ho = H >> 1, Wo = W >> 1;
    int64_t pix = idx;
    const int wo = (int)(pix % Wo);
    pix /= Wo;
    const int ho = (int)(pix % Ho);
    const int b = (int)(pix / Ho);
    const float k[4] = {0.125f, 0.375f, 0.375f, 0.125f};
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int hh = 2 * ho - 1 + a;
        if (hh < 0 || hh >= H) continue;
#pragma unroll
        for (int bb = 0; bb < 4; ++bb) {
            const int ww = 2 * wo - 1 + bb;
            if (ww < 0 || ww >= W) continue;
            const float4 v = __ldg(x + ((int64_t)b * H + hh) * W + ww);
            const float kw = k[a] * k[bb];
            acc.x = fmaf(kw, v.x, acc.x);
            acc.y = fmaf(kw, v.y, acc.y);
            acc.z = fmaf(kw, v.z, acc.z);
            acc.w = fmaf(kw, v.w, acc.w);
        }
    }
    out[idx] = acc;
}

__global__ void __launch_bounds__(256)
fir_up2_f4_kernel(const float4* __restrict__ x, int H, int W, float4* __restrict__ out, int64_t total) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int Ho = H * 2, Wo = W * 2;
    int64_t pix = idx;
    const int wo = (int)(pix % Wo);
    pix /= Wo;
    const int ho = (int)(pix % Ho);
    const int b = (int)(pix / Ho);
    const int hi = ho >> 1, wi = wo >> 1;
    const int ha = (ho & 1) ? hi : hi - 1;
    const int wa = (wo & 1) ? wi : wi - 1;
    const float kha = (ho & 1) ? 0.75f : 0.25f;
    const float kwa = (wo & 1) ? 0.75f : 0.25f;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const int hh = ha + a;
        if (hh < 0 || hh >= H) continue;
        const float kh = a == 0 ? kha : 1.0f - kha;
#pragma unroll
        for (int bb = 0; bb < 2; ++bb) {
            const int ww = wa + bb;
            if (ww < 0 || ww >= W) continue;
            const float kk = kh * (bb == 0 ? kwa : 1.0f - kwa);
            const float4 v = __ldg(x + ((int64_t)b * H + hh) * W + ww);
            acc.x = fmaf(kk, v.x, acc.x);
            acc.y = fmaf(kk, v.y, acc.y);
            acc.z = fmaf(kk, v.z, acc.z);
            acc.w = fmaf(kk, v.w, acc.w);
        }
    }
    out[idx] = acc;
}

}  // namespace

int fir_down2_launch(const ActView* x, const ActView* out, cudaStream_t s) {
    SNRSE_CHECK_ARG(x->H % 2 == 0 && x->W % 2 == 0 && x->C % 8 == 0, "fir_down2: H, W must be even, C %% 8 == 0");
    const int64_t total = (int64_t)x->B * (x->H / 2) * (x->W / 2) * (x->C / 8);
    fir_down2_kernel<<<(unsigned)cdiv64(total, 256), 256, 0, s>>>(x->ptr, x->ld, x->C, x->H, x->W, out->ptr, out->ld, total);
    SNRSE_LAUNCH_CHECK();
    return SNRSE_OK;
}

int fir_up2_launch(const ActView* x, const ActView* out, cudaStream_t s) {
    SNRSE_CHECK_ARG(x->C % 8 == 0, "fir_up2: C %% 8 == 0");
    const int64_t total = (int64_t)x->B * (x->H * 2) * (x->W * 2) * (x->C / 8);
    fir_up2_kernel<<<(unsigned)cdiv64(total, 256), 256, 0, s>>>(x->ptr, x->ld, x->C, x->H, x->W, out->ptr, out->ld, total);
    SNRSE_LAUNCH_CHECK();
    return SNRSE_OK;
}

int fir_down2_f4_launch(const float* x, float* out, int B, int H, int W, cudaStream_t s) {
    SNRSE_CHECK_ARG(H % 2 == 0 && W % 2 == 0, "fir_down2_f4: H, W must be even");
    const int64_t total = (int64_t)B * (H / 2) * (W / 2);
    fir_down2_f4_kernel<<<(unsigned)cdiv64(total, 256), 256, 0, s>>>(reinterpret_cast<const float4*>(x), H, W,
                                                                     reinterpret_cast<float4*>(out), total);
    SNRSE_LAUNCH_CHECK();
    return SNRSE_OK;
}

int fir_up2_f4_launch(const float* x, float* out, int B, int H, int W, cudaStream_t s) {
    const int64_t total = (int64_t)B * (H * 2) * (W * 2);
    fir_up2_f4_kernel<<<(unsigned)cdiv64(total, 256), 256, 0, s>>>(reinterpret_cast<const float4*>(x), H, W,
                                                                   reinterpret_cast<float4*>(out), total);
    SNRSE_LAUNCH_CHECK();
    return SNRSE_OK;
}

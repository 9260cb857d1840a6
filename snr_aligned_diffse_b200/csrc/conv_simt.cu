// CUDA-core convolutions of the NCSN++ path:
//   * conv_in4   : the 4 -> nf input 3x3 convolution (ncsnpp.py:285; K = 36 is too thin for the tensor cores)
//   * conv_out4  : the C -> 4 pyramid-output 3x3 convolutions (ncsnpp.py:350,364), fused with the
//                  `pyramid + pyramid_h` add (:366)
//   * combine4   : `Combine(method='sum')` = 1x1 conv 4 -> C on the input pyramid + h (layerspp.py:46-61)
//   * conv_simt  : a slow direct convolution with the same contract as conv_gemm, kept as an on-GPU
//                  cross-check for the tcgen05 kernel (selected with conv_impl=1; never the default).
#include "kernels.h"

namespace {

// ------------------------------------------------------------------------------------------------
// conv_in4: thread = (pixel quad, 8 output channels); block = strip of PIX pixels of one image row.
// ------------------------------------------------------------------------------------------------
constexpr int CI_PIX = 64;  // pixels per block along W

template <int COUT>
__global__ void __launch_bounds__(COUT / 8 * (CI_PIX / 4))
conv_in4_kernel(const float4* __restrict__ x4, const float* __restrict__ w, const float* __restrict__ bias, bf16* out,
                int H, int W, int ld) {
    constexpr int TPP = COUT / 8;  // threads covering the channels of one pixel
    __shared__ float4 s_in[3][CI_PIX + 2];
    __shared__ __align__(16) float s_w[36][COUT];  // [tap*4 + cin][cout]
    const int b = blockIdx.z, h = blockIdx.y, w0 = blockIdx.x * CI_PIX;
    const int tid = threadIdx.x, nthr = blockDim.x;
    for (int i = tid; i < 36 * COUT; i += nthr) {
        const int co = i / 36, k = i % 36;  // global layout [cout][tap][cin]
        s_w[k][co] = w[i];
    }
    for (int i = tid; i < 3 * (CI_PIX + 2); i += nthr) {
        const int r = i / (CI_PIX + 2), c = i % (CI_PIX + 2);
        const int hh = h + r - 1, ww = w0 + c - 1;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = x4[((int64_t)b * H + hh) * W + ww];
        s_in[r][c] = v;
    }
    __syncthreads();
    const int cg = tid % TPP, pq = tid / TPP;  // channel group, pixel quad
    float acc[4][8];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[p][j] = bias[cg * 8 + j];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            float4 xin[4];
#pragma unroll
            for (int p = 0; p < 4; ++p) xin[p] = s_in[r][pq * 4 + p + s];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float4 wa = *reinterpret_cast<const float4*>(&s_w[(r * 3 + s) * 4 + c][cg * 8]);
                const float4 wb = *reinterpret_cast<const float4*>(&s_w[(r * 3 + s) * 4 + c][cg * 8 + 4]);
                const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const float xv = c == 0 ? xin[p].x : (c == 1 ? xin[p].y : (c == 2 ? xin[p].z : xin[p].w));
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[p][j] = fmaf(xv, wv[j], acc[p][j]);
                }
            }
        }
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int ww = w0 + pq * 4 + p;
        if (ww < W) {
            const int64_t pix = ((int64_t)b * H + h) * W + ww;
            *reinterpret_cast<uint4*>(out + pix * ld + cg * 8) = pack8(acc[p]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// conv_out4: lane = pixel (32 consecutive pixels along W per warp), weights broadcast from smem.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
conv_out4_kernel(const bf16* __restrict__ a, int ld, int C, const float* __restrict__ w, const float* __restrict__ bias,
                 const float4* __restrict__ addend, float4* __restrict__ out, int H, int W) {
    extern __shared__ __align__(16) float s_w4[];  // [tap][c][4]
    const int b = blockIdx.z;
    for (int i = threadIdx.x; i < 9 * C * 4; i += blockDim.x) {
        const int o = i / (9 * C), rem = i % (9 * C);  // global layout [4][tap][C]
        s_w4[rem * 4 + o] = w[i];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.y * 8 + warp;
    const int ww = blockIdx.x * 32 + lane;
    if (h >= H || ww >= W) return;
    float acc0 = bias[0], acc1 = bias[1], acc2 = bias[2], acc3 = bias[3];
    for (int r = 0; r < 3; ++r) {
        const int hh = h + r - 1;
        if (hh < 0 || hh >= H) continue;
        for (int s = 0; s < 3; ++s) {
            const int wx = ww + s - 1;
            if (wx < 0 || wx >= W) continue;
            const uint4* src = reinterpret_cast<const uint4*>(a + (((int64_t)b * H + hh) * W + wx) * ld);
            const float4* wt = reinterpret_cast<const float4*>(s_w4) + (r * 3 + s) * C;
            for (int c8 = 0; c8 < C / 8; ++c8) {
                float f[8];
                unpack8(__ldg(src + c8), f);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 wv = wt[c8 * 8 + j];
                    acc0 = fmaf(f[j], wv.x, acc0);
                    acc1 = fmaf(f[j], wv.y, acc1);
                    acc2 = fmaf(f[j], wv.z, acc2);
                    acc3 = fmaf(f[j], wv.w, acc3);
                }
            }
        }
    }
    const int64_t pix = ((int64_t)b * H + h) * W + ww;
    if (addend) {
        const float4 ad = addend[pix];
        acc0 += ad.x; acc1 += ad.y; acc2 += ad.z; acc3 += ad.w;
    }
    out[pix] = make_float4(acc0, acc1, acc2, acc3);
}

// ------------------------------------------------------------------------------------------------
// combine4: out = h + W(Cx4) p4 + bias, thread = (pixel, 8 channels)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
combine4_kernel(const float4* __restrict__ p4, const bf16* __restrict__ h, int h_ld, const float4* __restrict__ w,
                const float* __restrict__ bias, bf16* out, int out_ld, int C, int64_t npix) {
    const int tpp = C / 8;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= npix * tpp) return;
    const int64_t pix = idx / tpp;
    const int c0 = (int)(idx % tpp) * 8;
    const float4 p = p4[pix];
    float f[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(h + pix * h_ld + c0)), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 wv = __ldg(w + c0 + j);
        f[j] += bias[c0 + j] + wv.x * p.x + wv.y * p.y + wv.z * p.z + wv.w * p.w;
    }
    *reinterpret_cast<uint4*>(out + pix * out_ld + c0) = pack8(f);
}

// ------------------------------------------------------------------------------------------------
// conv_simt: cross-check path.  Block = 8 consecutive pixels of one row; thread n -> output column(s).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
conv_simt_kernel(const bf16* __restrict__ a0, int ld0, int C0, int taps0, const bf16* __restrict__ a1, int ld1, int C1,
                 const bf16* __restrict__ wt, int N, const float* __restrict__ bias, const float* __restrict__ tbias,
                 int tb_stride, const bf16* __restrict__ res, int res_ld, float scale, bf16* out, int out_ld, int H,
                 int W) {
    __shared__ float s_a[8][512];
    const int b = blockIdx.z, h = blockIdx.y, w0 = blockIdx.x * 8;
    const int ktot = taps0 * C0 + C1;
    float acc[2][8];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int p = 0; p < 8; ++p) acc[q][p] = 0.f;
    const int nseg = taps0 + (C1 > 0 ? 1 : 0);
    for (int seg = 0; seg < nseg; ++seg) {
        const bool second = seg >= taps0;
        const bf16* src = second ? a1 : a0;
        const int ld = second ? ld1 : ld0, C = second ? C1 : C0;
        int dh = 0, dw = 0;
        if (!second && taps0 == 9) {
            dh = seg / 3 - 1;
            dw = seg % 3 - 1;
        }
        const int koff = second ? taps0 * C0 : seg * C0;
        __syncthreads();
        for (int i = threadIdx.x; i < 8 * C; i += blockDim.x) {
            const int p = i / C, c = i % C;
            const int hh = h + dh, ww = w0 + p + dw;
            float v = 0.f;
            if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = __bfloat162float(src[(((int64_t)b * H + hh) * W + ww) * ld + c]);
            s_a[p][c] = v;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int n = threadIdx.x + q * 128;
            if (n < N) {
                const bf16* wr = wt + (int64_t)n * ktot + koff;
                for (int c = 0; c < C; ++c) {
                    const float wv = __bfloat162float(wr[c]);
#pragma unroll
                    for (int p = 0; p < 8; ++p) acc[q][p] = fmaf(s_a[p][c], wv, acc[q][p]);
                }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int n = threadIdx.x + q * 128;
        if (n >= N) continue;
        for (int p = 0; p < 8; ++p) {
            const int ww = w0 + p;
            if (ww >= W) continue;
            const int64_t pix = ((int64_t)b * H + h) * W + ww;
            float v = acc[q][p];
            if (bias) v += bias[n];
            if (tbias) v += tbias[(int64_t)b * tb_stride + n];
            if (res) v += __bfloat162float(res[pix * res_ld + n]);
            out[pix * out_ld + n] = __float2bfloat16(v * scale);
        }
    }
}

}  // namespace

int conv_in4_launch(const float* x4, const float* w, const float* bias, const ActView* out, cudaStream_t s) {
    dim3 grid(cdiv(out->W, CI_PIX), out->H, out->B);
    if (out->C == 128) {
        conv_in4_kernel<128><<<grid, 128 / 8 * (CI_PIX / 4), 0, s>>>(reinterpret_cast<const float4*>(x4), w, bias,
                                                                     out->ptr, out->H, out->W, out->ld);
    } else if (out->C == 64) {
        conv_in4_kernel<64><<<grid, 64 / 8 * (CI_PIX / 4), 0, s>>>(reinterpret_cast<const float4*>(x4), w, bias,
                                                                   out->ptr, out->H, out->W, out->ld);
    } else {
        snrse_set_error("conv_in4: nf must be 64 or 128 (got %d)", out->C);
        return SNRSE_ERR_UNSUPPORTED;
    }
    SNRSE_LAUNCH_CHECK();
    return SNRSE_OK;
}

int conv_out4_launch(const ActView* a, const float* w, const float* bias, const float* addend4, float* out4,
                     cudaStream_t s) {
    SNRSE_CHECK_ARG(a->C % 8 == 0 && a->C <= 512, "conv_out4: unsupported channel count %d", a->C);
    static bool attr_set = false;
    if (!attr_set) {
        SNRSE_CUDA(cudaFuncSetAttribute(conv_out4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 9 * 512 * 16));
        attr_set = true;
    }
    dim3 grid(cdiv(a->W, 32), cdiv(a->H, 8), a->B);
    conv_out4_kernel<<<grid, 256, 9 * a->C * 16, s>>>(a->ptr, a->ld, a->C, w, bias,
                                                      reinterpret_cast<const float4*>(addend4),
                                                      reinterpret_cast<float4*>(out4), a->H, a->W);
    SNRSE_LAUNCH_CHECK();
    return SNRSE_OK;
}

int combine4_launch(const float* p4, const ActView* h, const float* w, const float* bias, const ActView* out,
                    cudaStream_t s) {
    const int64_t npix = (int64_t)h->B * h->H * h->W;
    const int64_t total = npix * (h->C / 8);
    combine4_kernel<<<(unsigned)cdiv64(total, 256), 256, 0, s>>>(reinterpret_cast<const float4*>(p4), h->ptr, h->ld,
                                                                 reinterpret_cast<const float4*>(w), bias, out->ptr,
                                                                 out->ld, h->C, npix);
    SNRSE_LAUNCH_CHECK();
    return SNRSE_OK;
}

int conv_simt_launch(const ActView* a0, int taps0, const ActView* a1, const bf16* wt, int N, const float* bias,
                     const float* tbias, int tb_stride, const ActView* res, float scale, bf16* out, int out_ld,
                     cudaStream_t s) {
    SNRSE_CHECK_ARG(N <= 256 && a0->C <= 512 && (!a1 || a1->C <= 512), "conv_simt: shape out of range");
    dim3 grid(cdiv(a0->W, 8), a0->H, a0->B);
    conv_simt_kernel<<<grid, 128, 0, s>>>(a0->ptr, a0->ld, a0->C, taps0, a1 ? a1->ptr : nullptr, a1 ? a1->ld : 0,
                                          a1 ? a1->C : 0, wt, N, bias, tbias, tb_stride, res ? res->ptr : nullptr,
                                          res ? res->ld : 0, scale, out, out_ld, a0->H, a0->W);
    SNRSE_LAUNCH_CHECK();
    return SNRSE_OK;
}

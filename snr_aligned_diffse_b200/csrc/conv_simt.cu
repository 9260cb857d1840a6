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
// conv_in4: lane = pixel.  A thread keeps the 3x3x4 input patches of TWO pixels (rows r and r+4 of an
// 8 x 32 pixel tile) in registers and walks the output channels 16 at a time; the weights of the current
// (tap, cin) are broadcast from shared memory (4 LDS.128 per 32 FMA -> FMA-bound, not LDS-bound).
// ------------------------------------------------------------------------------------------------
template <int COUT>
__global__ void __launch_bounds__(128)
conv_in4_kernel(const float4* __restrict__ x4, const float* __restrict__ w, const float* __restrict__ bias, bf16* out,
                int H, int W, int ld) {
    pdl_sync();
    __shared__ __align__(16) float s_w[36][COUT];  // [tap*4 + cin][cout]
    __shared__ float s_b[COUT];
    const int b = blockIdx.z, h0 = blockIdx.y * 8, w0 = blockIdx.x * 32;
    for (int i = threadIdx.x; i < 36 * COUT; i += 128) {
        const int co = i / 36, k = i % 36;  // global layout [cout][tap][cin]
        s_w[k][co] = w[i];
    }
    for (int i = threadIdx.x; i < COUT; i += 128) s_b[i] = bias[i];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ww = w0 + lane;
    float xin[2][36];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const int h = h0 + warp + 4 * p;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int hh = h + r - 1, wx = ww + q - 1;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (hh >= 0 && hh < H && wx >= 0 && wx < W) v = __ldg(x4 + ((int64_t)b * H + hh) * W + wx);
                xin[p][(r * 3 + q) * 4 + 0] = v.x;
                xin[p][(r * 3 + q) * 4 + 1] = v.y;
                xin[p][(r * 3 + q) * 4 + 2] = v.z;
                xin[p][(r * 3 + q) * 4 + 3] = v.w;
            }
    }
    __syncthreads();
    for (int c0 = 0; c0 < COUT; c0 += 16) {
        float acc[2][16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[0][j] = acc[1][j] = s_b[c0 + j];
#pragma unroll
        for (int k = 0; k < 36; ++k) {
            float wv[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 t = *reinterpret_cast<const float4*>(&s_w[k][c0 + 4 * q]);
                wv[4 * q] = t.x; wv[4 * q + 1] = t.y; wv[4 * q + 2] = t.z; wv[4 * q + 3] = t.w;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                acc[0][j] = fmaf(xin[0][k], wv[j], acc[0][j]);
                acc[1][j] = fmaf(xin[1][k], wv[j], acc[1][j]);
            }
        }
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const int h = h0 + warp + 4 * p;
            if (h < H && ww < W) {
                uint4* dst = reinterpret_cast<uint4*>(out + (((int64_t)b * H + h) * W + ww) * ld + c0);
                dst[0] = pack8(acc[p]);
                dst[1] = pack8(acc[p] + 8);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// conv_out4: 8 x 32 pixel tile per block (128 threads, 2 pixels per thread).  The input patch is staged
// through shared memory in 32-channel chunks, stored channel-pair-major so that lane = pixel reads are
// bank-conflict free; weights of the chunk are broadcast.  Fused with the `pyramid + pyramid_h` add.
// ------------------------------------------------------------------------------------------------
constexpr int CO_TH = 8, CO_TW = 32, CO_CH = 32;
__global__ void __launch_bounds__(128)
conv_out4_kernel(const bf16* __restrict__ a, int ld, int C, const float* __restrict__ w, const float* __restrict__ bias,
                 const float4* __restrict__ addend, float4* __restrict__ out, int H, int W) {
    pdl_sync();
    __shared__ uint32_t s_x[CO_CH / 2][CO_TH + 2][CO_TW + 2];   // bf16 pairs, [c2][row][col]   21.8 KB
    __shared__ __align__(16) float s_w[9][CO_CH][4];             // [tap][c][out]                4.6 KB
    const int b = blockIdx.z, h0 = blockIdx.y * CO_TH, w0 = blockIdx.x * CO_TW;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc[2][4];
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int o = 0; o < 4; ++o) acc[p][o] = bias[o];
    for (int c0 = 0; c0 < C; c0 += CO_CH) {
        __syncthreads();
        // stage weights of this chunk: global [4][tap][C]
        for (int i = threadIdx.x; i < 9 * CO_CH * 4; i += 128) {
            const int o = i & 3, c = (i >> 2) % CO_CH, tap = i / (4 * CO_CH);
            s_w[tap][c][o] = w[((int64_t)o * 9 + tap) * C + c0 + c];
        }
        // stage the (8+2) x (32+2) pixel patch, 32 channels = 4 pieces of 16 B per pixel
        for (int i = threadIdx.x; i < (CO_TH + 2) * (CO_TW + 2) * 4; i += 128) {
            const int piece = i & 3, pix = i >> 2;
            const int r = pix / (CO_TW + 2), c = pix % (CO_TW + 2);
            const int hh = h0 + r - 1, wx = w0 + c - 1;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (hh >= 0 && hh < H && wx >= 0 && wx < W)
                v = __ldg(reinterpret_cast<const uint4*>(a + (((int64_t)b * H + hh) * W + wx) * ld + c0 + piece * 8));
            s_x[piece * 4 + 0][r][c] = v.x;
            s_x[piece * 4 + 1][r][c] = v.y;
            s_x[piece * 4 + 2][r][c] = v.z;
            s_x[piece * 4 + 3][r][c] = v.w;
        }
        __syncthreads();
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int dy = tap / 3, dx = tap % 3;
#pragma unroll 4
            for (int c2 = 0; c2 < CO_CH / 2; ++c2) {
                const float4 wa = *reinterpret_cast<const float4*>(&s_w[tap][2 * c2][0]);
                const float4 wb = *reinterpret_cast<const float4*>(&s_w[tap][2 * c2 + 1][0]);
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    const uint32_t xv = s_x[c2][warp + 4 * p + dy][lane + dx];
                    const float x0 = __uint_as_float(xv << 16), x1 = __uint_as_float(xv & 0xffff0000u);
                    acc[p][0] = fmaf(x0, wa.x, fmaf(x1, wb.x, acc[p][0]));
                    acc[p][1] = fmaf(x0, wa.y, fmaf(x1, wb.y, acc[p][1]));
                    acc[p][2] = fmaf(x0, wa.z, fmaf(x1, wb.z, acc[p][2]));
                    acc[p][3] = fmaf(x0, wa.w, fmaf(x1, wb.w, acc[p][3]));
                }
            }
        }
    }
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const int h = h0 + warp + 4 * p, ww = w0 + lane;
        if (h < H && ww < W) {
            const int64_t pix = ((int64_t)b * H + h) * W + ww;
            float4 r = make_float4(acc[p][0], acc[p][1], acc[p][2], acc[p][3]);
            if (addend) {
                const float4 ad = addend[pix];
                r.x += ad.x; r.y += ad.y; r.z += ad.z; r.w += ad.w;
            }
            out[pix] = r;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// combine4: out = h + W(Cx4) p4 + bias.  Thread = (8-channel chunk, pixel slot); its 32 weights + 8 biases stay in
// registers while it walks pixels, so the pass is bound by reading h / writing out (HBM), not by weight loads.
// ------------------------------------------------------------------------------------------------
constexpr int CB_PIX = 32;   // pixels per thread
__global__ void __launch_bounds__(256)
combine4_kernel(const float4* __restrict__ p4, const bf16* __restrict__ h, int h_ld, const float4* __restrict__ w,
                const float* __restrict__ bias, bf16* out, int out_ld, int C, int64_t npix) {
    pdl_sync();
    const int tpp = C / 8, slots = blockDim.x / tpp;
    const int c0 = (threadIdx.x % tpp) * 8, slot = threadIdx.x / tpp;
    float wr[8][4], bs[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 wv = __ldg(w + c0 + j);
        wr[j][0] = wv.x; wr[j][1] = wv.y; wr[j][2] = wv.z; wr[j][3] = wv.w;
        bs[j] = __ldg(bias + c0 + j);
    }
    const int64_t p0 = (int64_t)blockIdx.x * slots * CB_PIX + slot;
#pragma unroll 4
    for (int k = 0; k < CB_PIX; ++k) {
        const int64_t pix = p0 + (int64_t)k * slots;
        if (pix >= npix) break;
        const float4 p = __ldg(p4 + pix);
        float f[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(h + pix * h_ld + c0)), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] += bs[j] + wr[j][0] * p.x + wr[j][1] * p.y + wr[j][2] * p.z + wr[j][3] * p.w;
        *reinterpret_cast<uint4*>(out + pix * out_ld + c0) = pack8(f);
    }
}

// ------------------------------------------------------------------------------------------------
// conv_simt: cross-check path.  Block = 8 consecutive pixels of one row; thread n -> output column(s).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
conv_simt_kernel(const bf16* __restrict__ a0, int ld0, int C0, int taps0, const bf16* __restrict__ a1, int ld1, int C1,
                 const bf16* __restrict__ wt, int N, const float* __restrict__ bias, const float* __restrict__ tbias,
                 int tb_stride, const bf16* __restrict__ res, int res_ld, float scale, bf16* out, int out_ld, int H,
                 int W) {
    pdl_sync();
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
    SNRSE_CHECK_ARG(out->C == 128, "conv_in4: nf must be 128 (got %d)", out->C);
    dim3 grid(cdiv(out->W, 32), cdiv(out->H, 8), out->B);
    snrse_launch(conv_in4_kernel<128>, dim3(grid), dim3(128), 0, s, reinterpret_cast<const float4*>(x4), w, bias, out->ptr, out->H, out->W,
                                              out->ld);
    SNRSE_LAUNCH_CHECK();
    return SNRSE_OK;
}

int conv_out4_launch(const ActView* a, const float* w, const float* bias, const float* addend4, float* out4,
                     cudaStream_t s) {
    SNRSE_CHECK_ARG(a->C % CO_CH == 0 && a->ld % 8 == 0, "conv_out4: C must be a multiple of %d (got %d)", CO_CH, a->C);
    dim3 grid(cdiv(a->W, CO_TW), cdiv(a->H, CO_TH), a->B);
    snrse_launch(conv_out4_kernel, dim3(grid), dim3(128), 0, s, a->ptr, a->ld, a->C, w, bias, reinterpret_cast<const float4*>(addend4),
                                          reinterpret_cast<float4*>(out4), a->H, a->W);
    SNRSE_LAUNCH_CHECK();
    return SNRSE_OK;
}

int combine4_launch(const float* p4, const ActView* h, const float* w, const float* bias, const ActView* out,
                    cudaStream_t s) {
    const int64_t npix = (int64_t)h->B * h->H * h->W;
    const int tpp = h->C / 8, nthr = tpp * (256 / tpp), slots = nthr / tpp;
    snrse_launch(combine4_kernel, dim3((unsigned)cdiv64(npix, (int64_t)slots * CB_PIX)), dim3(nthr), 0, s, reinterpret_cast<const float4*>(p4), h->ptr, h->ld,
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
    snrse_launch(conv_simt_kernel, dim3(grid), dim3(128), 0, s, a0->ptr, a0->ld, a0->C, taps0, a1 ? a1->ptr : nullptr, a1 ? a1->ld : 0,
                                          a1 ? a1->C : 0, wt, N, bias, tbias, tb_stride, res ? res->ptr : nullptr,
                                          res ? res->ld : 0, scale, out, out_ld, a0->H, a0->W);
    SNRSE_LAUNCH_CHECK();
    return SNRSE_OK;
}

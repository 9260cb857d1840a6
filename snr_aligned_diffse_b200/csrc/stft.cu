// STFT / iSTFT front and back end fused with the "exponent" spectrogram transform.
// Replaces torch.stft / torch.istft + `spec_fwd` / `spec_back` + `pad_spec`
// (sgmse-bbed/sgmse/data_module.py:241-297, util/other.py:83-99, model.py:612-625,715-719,749-752,828-830):
//   n_fft = 510 (256 bins), hop 128, periodic Hann, center=True (reflect pad 255), one-sided, unnormalised.
//   forward : frame f of utterance b reads samples reflect(f*128 + n - 255), n in [0,510);
//             X[k] = sum_n w[n] x[n] e^{-2 pi i k n / 510};  Y = beta |X|^alpha e^{i arg X};
//             frames >= 1 + len/128 (the `pad_spec` region up to Tpad) are written as zeros.
//   inverse : Z = (Y/beta) |Y/beta|^{1/alpha - 1};  x_f[n] = w[n]/510 * irDFT(Z_f)[n];
//             out[m] = sum_f x_f[m + 255 - 128 f] / sum_f w^2[m + 255 - 128 f], m in [0, len).
// n_fft = 510 = 2*3*5*17 has no radix-2 structure; the transform is evaluated as a direct DFT from a
// shared-memory twiddle table with exact integer phase indices ((k*n) mod 510), 8 frames per block.
// 0.52 MFLOP per frame: 0.025 % of the network's work, so HBM traffic, not FLOPs, bounds these kernels.
#include "kernels.h"

namespace {

constexpr int NFFT = 510, HOP = 128, NBINS = 256, FT = 8, HALF = 255;

__device__ __forceinline__ void fill_twiddles(float2* tw) {
    for (int i = threadIdx.x; i < NFFT; i += blockDim.x) {
        float s, c;
        sincospif((float)(2 * i) / (float)NFFT, &s, &c);
        tw[i] = make_float2(c, s);
    }
}

__global__ void __launch_bounds__(256)
stft_kernel(const float* __restrict__ wave, const int* __restrict__ len, const float* __restrict__ scale,
            int scale_is_divisor, float* __restrict__ out, int lstride, int tpad, int transform, float alpha,
            float beta, int planar) {
    __shared__ float2 tw[NFFT];
    __shared__ __align__(16) float xs[NFFT][FT];  // windowed samples, [n][frame]
    const int b = blockIdx.y, f0 = blockIdx.x * FT;
    const int L = len ? len[b] : lstride;
    const int nframes = 1 + L / HOP;
    fill_twiddles(tw);
    __syncthreads();
    float sc = 1.0f;
    if (scale) sc = scale[b];
    const float* wb = wave + (int64_t)b * lstride;
    for (int e = threadIdx.x; e < NFFT * FT; e += blockDim.x) {
        const int f = e / NFFT, n = e % NFFT;
        float v = 0.f;
        if (f0 + f < nframes) {
            int i = (f0 + f) * HOP + n - HALF;
            if (i < 0) i = -i;
            if (i >= L) i = 2 * (L - 1) - i;
            i = max(0, min(i, L - 1));  // only reachable for len <= 255, which torch.stft rejects
            v = wb[i];
            if (scale) v = scale_is_divisor ? v / sc : v * sc;
            v *= 0.5f - 0.5f * tw[n].x;
        }
        xs[n][f] = v;
    }
    __syncthreads();
    const int k = threadIdx.x;
    float re[FT], im[FT];
#pragma unroll
    for (int f = 0; f < FT; ++f) re[f] = im[f] = 0.f;
    int ph = 0;
    for (int n = 0; n < NFFT; ++n) {
        const float2 w = tw[ph];
        ph += k;
        if (ph >= NFFT) ph -= NFFT;
        const float4 xa = *reinterpret_cast<const float4*>(&xs[n][0]);
        const float4 xb = *reinterpret_cast<const float4*>(&xs[n][4]);
        const float xv[FT] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
        for (int f = 0; f < FT; ++f) {
            re[f] = fmaf(xv[f], w.x, re[f]);
            im[f] = fmaf(-xv[f], w.y, im[f]);
        }
    }
#pragma unroll
    for (int f = 0; f < FT; ++f) {
        const int t = f0 + f;
        if (t >= tpad) break;
        float r = re[f], i = im[f];
        if (t >= nframes) {
            r = 0.f;
            i = 0.f;
        } else if (transform == 1) {
            // beta * |X|^alpha * e^{i arg X} = X * beta * |X|^(alpha-1), 0 -> 0  (data_module.py:241-247)
            const float mag = sqrtf(r * r + i * i);
            float g = 0.f;
            if (mag > 0.f) g = (alpha == 0.5f) ? beta / sqrtf(mag) : beta * powf(mag, alpha - 1.0f);
            r *= g;
            i *= g;
        }
        if (planar) {
            out[(((int64_t)b * 2 + 0) * NBINS + k) * tpad + t] = r;
            out[(((int64_t)b * 2 + 1) * NBINS + k) * tpad + t] = i;
        } else {
            reinterpret_cast<float2*>(out)[((int64_t)b * NBINS + k) * tpad + t] = make_float2(r, i);
        }
    }
}

__global__ void __launch_bounds__(256)
istft_frames_kernel(const float2* __restrict__ spec, float* __restrict__ frames, int tpad, int transform, float alpha,
                    float beta) {
    __shared__ float2 tw[NFFT];
    __shared__ __align__(16) float2 zs[NBINS][FT];
    const int b = blockIdx.y, f0 = blockIdx.x * FT;
    fill_twiddles(tw);
    for (int e = threadIdx.x; e < NBINS * FT; e += blockDim.x) {
        const int k = e / FT, f = e % FT;
        float2 z = make_float2(0.f, 0.f);
        if (f0 + f < tpad) z = spec[((int64_t)b * NBINS + k) * tpad + f0 + f];
        if (transform == 1) {
            // (Y/beta) -> |.|^(1/alpha) e^{i arg}  (data_module.py:256-262)
            z.x /= beta;
            z.y /= beta;
            const float mag = sqrtf(z.x * z.x + z.y * z.y);
            float g = 0.f;
            if (mag > 0.f) g = (alpha == 0.5f) ? mag : powf(mag, 1.0f / alpha - 1.0f);
            z.x *= g;
            z.y *= g;
        }
        // one-sided inverse: DC and Nyquist count once and their imaginary parts are ignored
        if (k == 0 || k == NBINS - 1) z.y = 0.f; else { z.x *= 2.f; z.y *= 2.f; }
        zs[k][f] = z;
    }
    __syncthreads();
    for (int n = threadIdx.x; n < NFFT; n += blockDim.x) {
        float acc[FT];
#pragma unroll
        for (int f = 0; f < FT; ++f) acc[f] = 0.f;
        int ph = 0;
        for (int k = 0; k < NBINS; ++k) {
            const float2 w = tw[ph];
            ph += n;
            if (ph >= NFFT) ph -= NFFT;
            const float4* zp = reinterpret_cast<const float4*>(&zs[k][0]);
#pragma unroll
            for (int q = 0; q < FT / 2; ++q) {
                const float4 z2 = zp[q];
                acc[2 * q] = fmaf(z2.x, w.x, fmaf(-z2.y, w.y, acc[2 * q]));
                acc[2 * q + 1] = fmaf(z2.z, w.x, fmaf(-z2.w, w.y, acc[2 * q + 1]));
            }
        }
        const float win = (0.5f - 0.5f * tw[n].x) * (1.0f / (float)NFFT);
#pragma unroll
        for (int f = 0; f < FT; ++f)
            if (f0 + f < tpad) frames[((int64_t)b * tpad + f0 + f) * 512 + n] = acc[f] * win;
    }
}

__global__ void __launch_bounds__(256)
istft_ola_kernel(const float* __restrict__ frames, const int* __restrict__ len, const float* __restrict__ scale,
                 float* __restrict__ wave, int lstride, int tpad) {
    const int b = blockIdx.y;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= lstride) return;
    const int L = len ? len[b] : lstride;
    float v = 0.f;
    if (n < L) {
        const int m = n + HALF;
        int fhi = m / HOP;
        if (fhi > tpad - 1) fhi = tpad - 1;
        int flo = (m - (NFFT - 1) + HOP - 1) / HOP;
        if (m - (NFFT - 1) <= 0) flo = 0;
        float acc = 0.f, env = 0.f;
        for (int f = flo; f <= fhi; ++f) {
            const int j = m - f * HOP;
            float s, c;
            sincospif((float)(2 * j) / (float)NFFT, &s, &c);
            const float w = 0.5f - 0.5f * c;
            env = fmaf(w, w, env);
            acc += frames[((int64_t)b * tpad + f) * 512 + j];
        }
        v = env > 1e-11f ? acc / env : 0.f;
        if (scale) v *= scale[b];
    }
    wave[(int64_t)b * lstride + n] = v;
}

__global__ void __launch_bounds__(256)
absmax_kernel(const float* __restrict__ wave, const int* __restrict__ len, int lstride, float* __restrict__ out) {
    __shared__ float sm[8];
    const int b = blockIdx.x;
    const int L = len ? len[b] : lstride;
    float m = 0.f;
    for (int i = threadIdx.x; i < L; i += blockDim.x) m = fmaxf(m, fabsf(wave[(int64_t)b * lstride + i]));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) m = fmaxf(m, sm[i]);
        out[b] = m;
    }
}

// stand-alone spec_fwd / spec_back (data_module.py:241-267) for callers that hold a raw STFT
__global__ void __launch_bounds__(256)
spec_transform_kernel(const float2* __restrict__ in, float2* __restrict__ out, int64_t n, int inverse, float alpha,
                      float beta) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float2 z = in[i];
    if (inverse) {
        z.x /= beta;
        z.y /= beta;
        const float mag = sqrtf(z.x * z.x + z.y * z.y);
        float g = 0.f;
        if (mag > 0.f) g = (alpha == 0.5f) ? mag : powf(mag, 1.0f / alpha - 1.0f);
        if (alpha == 1.0f) g = 1.0f;
        z.x *= g;
        z.y *= g;
    } else {
        const float mag = sqrtf(z.x * z.x + z.y * z.y);
        float g = 0.f;
        if (mag > 0.f) g = (alpha == 0.5f) ? beta / sqrtf(mag) : beta * powf(mag, alpha - 1.0f);
        if (alpha == 1.0f) g = beta;
        z.x *= g;
        z.y *= g;
    }
    out[i] = z;
}

}  // namespace

int spec_transform_launch(const float2* in, float2* out, int64_t n, int inverse, float alpha, float beta, cudaStream_t s) {
    spec_transform_kernel<<<(unsigned)cdiv64(n, 256), 256, 0, s>>>(in, out, n, inverse, alpha, beta);
    SNRSE_LAUNCH_CHECK();
    return SNRSE_OK;
}

int stft_launch(const float* wave, const int* len, const float* scale, int scale_is_divisor, float* out, int B,
                int lstride, int tpad, int transform, float alpha, float beta, int planar, cudaStream_t s) {
    SNRSE_CHECK_ARG(B > 0 && tpad > 0 && lstride > HALF, "stft: need B>0, Tpad>0 and more than 255 samples");
    SNRSE_CHECK_ARG(transform == 0 || transform == 1, "stft: transform must be 0 (none) or 1 (exponent)");
    dim3 grid(cdiv(tpad, FT), B);
    stft_kernel<<<grid, 256, 0, s>>>(wave, len, scale, scale_is_divisor, out, lstride, tpad, transform, alpha, beta, planar);
    SNRSE_LAUNCH_CHECK();
    return SNRSE_OK;
}

int istft_launch(const float2* spec, const int* len, const float* scale, float* wave, float* frames_ws, int B,
                 int lstride, int tpad, int transform, float alpha, float beta, cudaStream_t s) {
    SNRSE_CHECK_ARG(B > 0 && tpad > 0 && lstride > 0, "istft: bad shape");
    dim3 g1(cdiv(tpad, FT), B);
    istft_frames_kernel<<<g1, 256, 0, s>>>(spec, frames_ws, tpad, transform, alpha, beta);
    SNRSE_LAUNCH_CHECK();
    dim3 g2(cdiv(lstride, 256), B);
    istft_ola_kernel<<<g2, 256, 0, s>>>(frames_ws, len, scale, wave, lstride, tpad);
    SNRSE_LAUNCH_CHECK();
    return SNRSE_OK;
}

int absmax_launch(const float* wave, const int* len, int B, int lstride, float* out, cudaStream_t s) {
    absmax_kernel<<<B, 256, 0, s>>>(wave, len, lstride, out);
    SNRSE_LAUNCH_CHECK();
    return SNRSE_OK;
}

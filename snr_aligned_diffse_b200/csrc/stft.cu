// STFT / iSTFT front and back end fused with the "exponent" spectrogram transform.
// Replaces torch.stft / torch.istft + `spec_fwd` / `spec_back` + `pad_spec`
// (sgmse-bbed/sgmse/data_module.py:241-297, util/other.py:83-99, model.py:612-625,715-719,749-752,828-830):
//   n_fft = 510 (256 bins), hop 128, periodic Hann, center=True (reflect pad 255), one-sided, unnormalised.
//   forward : frame f of utterance b reads samples reflect(f*128 + n - 255), n in [0,510);
//             X[k] = sum_n w[n] x[n] e^{-2 pi i k n / 510};  Y = beta |X|^alpha e^{i arg X};
//             frames >= 1 + len/128 (the `pad_spec` region up to Tpad) are written as zeros.
//   inverse : Z = (Y/beta) |Y/beta|^{1/alpha - 1};  x_f[n] = w[n]/510 * irDFT(Z_f)[n];
//             out[m] = sum_f x_f[m + 255 - 128 f] / sum_f w^2[m + 255 - 128 f], m in [0, len).
// FFT in shared memory.  The real 510-point transform is a complex 255-point transform of the packed signal
// z[m] = x[2m] + i x[2m+1] plus a split/merge pass.  255 = 3 * 5 * 17 with pairwise coprime factors, so the
// 255-point DFT is a twiddle-free 3 x 5 x 17 DFT (Good-Thomas): input index (85 n1 + 51 n2 + 15 n3) mod 255,
// output index k with (k mod 3, k mod 5, k mod 17) = (k1, k2, k3), i.e. k = (85 k1 + 51 k2 + 120 k3) mod 255.
// 25 complex multiply-adds per point instead of 510: 0.05 MFLOP per frame.  Eight frames per pass: one thread
// per transform point, the eight frames ride in registers as four float4 (two complex frames each).
// The inverse kernel also does the overlap-add, the window-envelope division and the rescale: every block
// recomputes a 4-frame halo so that all frames touching its output samples are in its own shared memory; no
// frame workspace in HBM, no floating-point atomics.
#include "kernels.h"

namespace {

constexpr int NFFT = 510, HOP = 128, NBINS = 256, FT = 8, HALF = 255, M255 = 255;
constexpr int QS = 256;                       // float4 stride between frame pairs in a transform buffer
constexpr int SPAN = (FT - 1) * HOP + NFFT;   // samples touched by FT consecutive frames (1406)

struct Twiddles {
    float2 w17[17], w5[5], w3[3];             // e^{-2 pi i j / r}
};

__device__ __forceinline__ void fill_twiddles(Twiddles& tw) {
    const int i = threadIdx.x;
    float s, c;
    if (i < 17) {
        sincospif((float)(2 * i) / 17.0f, &s, &c);
        tw.w17[i] = make_float2(c, -s);
    } else if (i >= 32 && i < 37) {
        sincospif((float)(2 * (i - 32)) / 5.0f, &s, &c);
        tw.w5[i - 32] = make_float2(c, -s);
    } else if (i >= 64 && i < 67) {
        sincospif((float)(2 * (i - 64)) / 3.0f, &s, &c);
        tw.w3[i - 64] = make_float2(c, -s);
    }
}

__device__ __forceinline__ float hann(int n) { return 0.5f - 0.5f * cospif((float)(2 * n) / (float)NFFT); }

// acc += x * w for the two complex frames packed in one float4
__device__ __forceinline__ void cmac2(float4& acc, const float4 x, const float2 w) {
    acc.x = fmaf(x.x, w.x, fmaf(-x.y, w.y, acc.x));
    acc.y = fmaf(x.x, w.y, fmaf(x.y, w.x, acc.y));
    acc.z = fmaf(x.z, w.x, fmaf(-x.w, w.y, acc.z));
    acc.w = fmaf(x.z, w.y, fmaf(x.w, w.x, acc.w));
}

// position (in natural order) of the element a thread stores at linear index j = (n1*5 + n2)*17 + n3
__device__ __forceinline__ int pfa_input_pos(int j) {
    const int g = j / 17, n3 = j - g * 17, n1 = g / 5, n2 = g - n1 * 5;
    return (85 * n1 + 51 * n2 + 15 * n3) % M255;
}

// 255-point DFT of FT frames.  In: A[q*QS + j] (linear PFA order, see pfa_input_pos).  Out: Bf[q*QS + k], natural
// order.  INV selects e^{+...}.  The caller synchronises after filling A; the result is visible on return.
template <bool INV>
__device__ __forceinline__ void fft255(float4* __restrict__ A, float4* __restrict__ Bf, const Twiddles& tw) {
    const int j = threadIdx.x;
    const float sg = INV ? -1.0f : 1.0f;
    if (j < M255) {  // 17-point DFTs along n3: j = g*17 + k3
        const int g = j / 17, k3 = j - g * 17;
        float4 acc[FT / 2];
#pragma unroll
        for (int q = 0; q < FT / 2; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        int ph = 0;
#pragma unroll
        for (int n3 = 0; n3 < 17; ++n3) {
            float2 w = tw.w17[ph];
            w.y *= sg;
            ph += k3;
            if (ph >= 17) ph -= 17;
#pragma unroll
            for (int q = 0; q < FT / 2; ++q) cmac2(acc[q], A[q * QS + g * 17 + n3], w);
        }
#pragma unroll
        for (int q = 0; q < FT / 2; ++q) Bf[q * QS + j] = acc[q];
    }
    __syncthreads();
    if (j < M255) {  // 5-point DFTs along n2: j = (n1*5 + k2)*17 + k3
        const int n1 = j / 85, r = j - n1 * 85, k2 = r / 17, k3 = r - k2 * 17;
        float4 acc[FT / 2];
#pragma unroll
        for (int q = 0; q < FT / 2; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        int ph = 0;
#pragma unroll
        for (int n2 = 0; n2 < 5; ++n2) {
            float2 w = tw.w5[ph];
            w.y *= sg;
            ph += k2;
            if (ph >= 5) ph -= 5;
#pragma unroll
            for (int q = 0; q < FT / 2; ++q) cmac2(acc[q], Bf[q * QS + (n1 * 5 + n2) * 17 + k3], w);
        }
#pragma unroll
        for (int q = 0; q < FT / 2; ++q) A[q * QS + j] = acc[q];
    }
    __syncthreads();
    if (j < M255) {  // 3-point DFTs along n1: j = (k1*5 + k2)*17 + k3, scattered to natural order
        const int k1 = j / 85, r = j - k1 * 85, k2 = r / 17, k3 = r - k2 * 17;
        float4 acc[FT / 2];
#pragma unroll
        for (int q = 0; q < FT / 2; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        int ph = 0;
#pragma unroll
        for (int n1 = 0; n1 < 3; ++n1) {
            float2 w = tw.w3[ph];
            w.y *= sg;
            ph += k1;
            if (ph >= 3) ph -= 3;
#pragma unroll
            for (int q = 0; q < FT / 2; ++q) cmac2(acc[q], A[q * QS + n1 * 85 + r], w);
        }
        const int k = (85 * k1 + 51 * k2 + 120 * k3) % M255;
#pragma unroll
        for (int q = 0; q < FT / 2; ++q) Bf[q * QS + k] = acc[q];
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256)
stft_kernel(const float* __restrict__ wave, const int* __restrict__ len, const float* __restrict__ scale,
            int scale_is_divisor, float* __restrict__ out, int lstride, int tpad, int transform, float alpha,
            float beta, int planar) {
    __shared__ __align__(16) float4 A[(FT / 2) * QS];
    __shared__ __align__(16) float4 Bf[(FT / 2) * QS];
    __shared__ __align__(16) float xs[SPAN + 2];
    __shared__ Twiddles tw;
    const int b = blockIdx.y, f0 = blockIdx.x * FT, tid = threadIdx.x;
    const int L = len ? len[b] : lstride;
    const int nframes = 1 + L / HOP;
    fill_twiddles(tw);
    float sc = 1.0f;
    if (scale) sc = scale[b];
    const float* wb = wave + (int64_t)b * lstride;
    for (int e = tid; e < SPAN; e += 256) {
        int i = f0 * HOP + e - HALF;
        if (i < 0) i = -i;
        if (i >= L) i = 2 * (L - 1) - i;
        i = max(0, min(i, L - 1));  // beyond one reflection: only frames >= nframes (written as zeros) or len <= 255
        float v = wb[i];
        if (scale) v = scale_is_divisor ? v / sc : v * sc;
        xs[e] = v;
    }
    __syncthreads();
    if (tid < M255) {  // windowed, packed z[m] = x[2m] + i x[2m+1] in PFA input order
        const int m = pfa_input_pos(tid);
        const float w0 = hann(2 * m), w1 = hann(2 * m + 1);
#pragma unroll
        for (int q = 0; q < FT / 2; ++q) {
            const float2 x0 = *reinterpret_cast<const float2*>(&xs[(2 * q) * HOP + 2 * m]);
            const float2 x1 = *reinterpret_cast<const float2*>(&xs[(2 * q + 1) * HOP + 2 * m]);
            A[q * QS + tid] = make_float4(x0.x * w0, x0.y * w1, x1.x * w0, x1.y * w1);
        }
    }
    __syncthreads();
    fft255<false>(A, Bf, tw);
    // split: X[k] = (Z[k] + conj Z[255-k])/2 - (i/2) e^{-2 pi i k/510} (Z[k] - conj Z[255-k]),  k = 0..255
    const int k = tid;
    const int ka = (k == M255) ? 0 : k, kb = (M255 - k) % M255;
    float es, ec;
    sincospif((float)k / (float)M255, &es, &ec);  // e^{-i pi k/255} = (ec, -es)
    float re[FT], im[FT];
#pragma unroll
    for (int q = 0; q < FT / 2; ++q) {
        const float4 za = Bf[q * QS + ka], zb = Bf[q * QS + kb];
        const float zr[2] = {za.x, za.z}, zi[2] = {za.y, za.w}, cr[2] = {zb.x, zb.z}, ci[2] = {-zb.y, -zb.w};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const float sr = zr[h] + cr[h], si = zi[h] + ci[h], dr = zr[h] - cr[h], di = zi[h] - ci[h];
            // E*d with E = (ec, -es):  (ec dr + es di,  ec di - es dr);  -i (a + i b) = b - i a
            const float pr = fmaf(ec, dr, es * di), pi = fmaf(ec, di, -es * dr);
            re[2 * q + h] = 0.5f * (sr + pi);
            im[2 * q + h] = 0.5f * (si - pr);
        }
    }
#pragma unroll
    for (int f = 0; f < FT; ++f) {
        const int t = f0 + f;
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
        re[f] = r;
        im[f] = i;
    }
    if (planar) {
        float* o0 = out + (((int64_t)b * 2 + 0) * NBINS + k) * tpad + f0;
        float* o1 = out + (((int64_t)b * 2 + 1) * NBINS + k) * tpad + f0;
        if ((tpad & 3) == 0 && f0 + FT <= tpad) {
            reinterpret_cast<float4*>(o0)[0] = make_float4(re[0], re[1], re[2], re[3]);
            reinterpret_cast<float4*>(o0)[1] = make_float4(re[4], re[5], re[6], re[7]);
            reinterpret_cast<float4*>(o1)[0] = make_float4(im[0], im[1], im[2], im[3]);
            reinterpret_cast<float4*>(o1)[1] = make_float4(im[4], im[5], im[6], im[7]);
        } else {
#pragma unroll
            for (int f = 0; f < FT; ++f)
                if (f0 + f < tpad) {
                    o0[f] = re[f];
                    o1[f] = im[f];
                }
        }
    } else {
        float2* o = reinterpret_cast<float2*>(out) + ((int64_t)b * NBINS + k) * tpad + f0;
        if ((tpad & 1) == 0 && f0 + FT <= tpad) {
#pragma unroll
            for (int q = 0; q < FT / 2; ++q)
                reinterpret_cast<float4*>(o)[q] = make_float4(re[2 * q], im[2 * q], re[2 * q + 1], im[2 * q + 1]);
        } else {
#pragma unroll
            for (int f = 0; f < FT; ++f)
                if (f0 + f < tpad) o[f] = make_float2(re[f], im[f]);
        }
    }
}

// Inverse: block x owns frames F0 = x*NEWF .. F0+NEWF-1 and output samples p in [128 F0, 128 (F0+NEWF)) (p = sample
// index in the 255-padded signal); it transforms frames F0-HALO .. F0+NEWF-1 in PASSES passes of FT frames, overlap-adds
// them in shared memory in ascending frame order, divides by the window envelope and writes wave[p - 255].
constexpr int PASSES = 4, HALO = 4, NEWF = PASSES * FT - HALO;      // 28 new frames per block
constexpr int OLA_SPAN = (PASSES * FT - 1) * HOP + NFFT;             // 4478 samples
constexpr int ISTFT_SMEM = 2 * (FT / 2) * QS * 16 + OLA_SPAN * 4 + NFFT * 4 + (int)sizeof(Twiddles) + 64;

__global__ void __launch_bounds__(256)
istft_kernel(const float2* __restrict__ spec, const int* __restrict__ len, const float* __restrict__ scale,
             float* __restrict__ wave, int lstride, int tpad, int transform, float alpha, float beta) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* A = reinterpret_cast<float4*>(smem_raw);
    float4* Bf = A + (FT / 2) * QS;
    float* ola = reinterpret_cast<float*>(Bf + (FT / 2) * QS);
    float* hw = ola + OLA_SPAN;                                       // Hann window
    Twiddles& tw = *reinterpret_cast<Twiddles*>(hw + NFFT + 2);
    const int b = blockIdx.y, F0 = blockIdx.x * NEWF, tid = threadIdx.x;
    const int L = len ? len[b] : lstride;
    fill_twiddles(tw);
    for (int i = tid; i < NFFT; i += 256) hw[i] = hann(i);
    for (int i = tid; i < OLA_SPAN; i += 256) ola[i] = 0.f;
    const float2* sb = spec + (int64_t)b * NBINS * tpad;
    const float inv_beta = 1.0f / beta;
    const bool vec_ok = (tpad & 1) == 0;
    // merge factors of this thread's transform point
    const int kz = (tid < M255) ? pfa_input_pos(tid) : 0;
    float es, ec;
    sincospif((float)kz / (float)M255, &es, &ec);                    // e^{+i pi k/255} = (ec, es)
    for (int pass = 0; pass < PASSES; ++pass) {
        const int fb = F0 - HALO + pass * FT;                        // first frame of this pass (even)
        if (fb + FT <= 0 || fb >= tpad) continue;                     // block-uniform: nothing to add
        __syncthreads();                                              // previous pass done with A / Bf
        // spectrogram tile -> Bf as float2 [k][FT], spec_back applied (data_module.py:256-262)
        float2* T = reinterpret_cast<float2*>(Bf);
        for (int e = tid; e < NBINS * (FT / 2); e += 256) {
            const int k = e >> 2, c = e & 3, t = fb + 2 * c;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t >= 0) {
                const float2* p = sb + (int64_t)k * tpad + t;
                if (vec_ok && t + 1 < tpad) {
                    v = *reinterpret_cast<const float4*>(p);
                } else {
                    if (t < tpad) { v.x = p[0].x; v.y = p[0].y; }
                    if (t + 1 < tpad) { v.z = p[1].x; v.w = p[1].y; }
                }
            }
            if (transform == 1) {
                v.x *= inv_beta; v.y *= inv_beta; v.z *= inv_beta; v.w *= inv_beta;
                const float m0 = sqrtf(v.x * v.x + v.y * v.y), m1 = sqrtf(v.z * v.z + v.w * v.w);
                float g0 = 0.f, g1 = 0.f;
                if (m0 > 0.f) g0 = (alpha == 0.5f) ? m0 : powf(m0, 1.0f / alpha - 1.0f);
                if (m1 > 0.f) g1 = (alpha == 0.5f) ? m1 : powf(m1, 1.0f / alpha - 1.0f);
                v.x *= g0; v.y *= g0; v.z *= g1; v.w *= g1;
            }
            if (k == 0 || k == NBINS - 1) { v.y = 0.f; v.w = 0.f; }  // one-sided inverse ignores Im of DC / Nyquist
            reinterpret_cast<float4*>(T)[e] = v;
        }
        __syncthreads();
        // merge: Z'[k] = (X[k] + conj X[255-k]) + i e^{+2 pi i k/510} (X[k] - conj X[255-k]),  k = 0..254
        if (tid < M255) {
            const float4* Ta = reinterpret_cast<const float4*>(T) + kz * 4;
            const float4* Tb = reinterpret_cast<const float4*>(T) + (M255 - kz) * 4;
#pragma unroll
            for (int q = 0; q < FT / 2; ++q) {
                const float4 xa = Ta[q], xb = Tb[q];
                float4 z;
                {
                    const float sr = xa.x + xb.x, si = xa.y - xb.y, dr = xa.x - xb.x, di = xa.y + xb.y;
                    const float pr = fmaf(ec, dr, -es * di), pi = fmaf(ec, di, es * dr);  // E*d
                    z.x = sr - pi;
                    z.y = si + pr;
                }
                {
                    const float sr = xa.z + xb.z, si = xa.w - xb.w, dr = xa.z - xb.z, di = xa.w + xb.w;
                    const float pr = fmaf(ec, dr, -es * di), pi = fmaf(ec, di, es * dr);
                    z.z = sr - pi;
                    z.w = si + pr;
                }
                A[q * QS + tid] = z;
            }
        }
        __syncthreads();
        fft255<true>(A, Bf, tw);
        // overlap-add the FT frames of this pass (ascending frame order), windowed and scaled by 1/510
        const float* zf = reinterpret_cast<const float*>(Bf);
        for (int s = tid; s < SPAN; s += 256) {
            int fhi = s / HOP;
            if (fhi > FT - 1) fhi = FT - 1;
            int flo = (s - (NFFT - 1) + HOP - 1) / HOP;
            if (s - (NFFT - 1) <= 0) flo = 0;
            float acc = 0.f;
            for (int f = flo; f <= fhi; ++f) {
                const int n = s - f * HOP;
                const float v = zf[((f >> 1) * QS + (n >> 1)) * 4 + (f & 1) * 2 + (n & 1)];
                acc = fmaf(v, hw[n], acc);
            }
            ola[pass * FT * HOP + s] += acc * (1.0f / (float)NFFT);
        }
    }
    __syncthreads();
    // owned samples: p in [128 F0, 128 (F0+NEWF)); the last block also owns everything after that
    const bool last = blockIdx.x == gridDim.x - 1;
    const int p0 = HOP * (F0 - HALO);                                 // padded-signal index of ola[0]
    const int pbeg = HOP * F0, pend = last ? lstride + HALF : HOP * (F0 + NEWF);
    float sc = 1.0f;
    if (scale) sc = scale[b];
    float* wo = wave + (int64_t)b * lstride;
    for (int p = pbeg + tid; p < pend; p += 256) {
        const int n = p - HALF;
        if (n < 0 || n >= lstride) continue;
        float v = 0.f;
        if (n < L && p - p0 < OLA_SPAN) {
            int fhi = p / HOP;
            if (fhi > tpad - 1) fhi = tpad - 1;
            int flo = (p - (NFFT - 1) + HOP - 1) / HOP;
            if (p - (NFFT - 1) <= 0) flo = 0;
            float env = 0.f;
            for (int f = flo; f <= fhi; ++f) {
                const float w = hw[p - f * HOP];
                env = fmaf(w, w, env);
            }
            v = env > 1e-11f ? ola[p - p0] / env * sc : 0.f;
        }
        wo[n] = v;
    }
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

// SI-SDR of an enhanced batch against clean references on the device (util/other.py:71-75, evaluated per utterance
// in B/eval.py:140-144 on host numpy arrays): alpha = <s_hat, s> / |s|^2;  10 log10(|alpha s|^2 / |alpha s - s_hat|^2).
// One block per utterance, two passes (alpha, then the two energies) in double with a fixed reduction tree.
__device__ __forceinline__ double block_sum_256(double v, double* sm) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sm[w];
    return t;
}

__global__ void __launch_bounds__(256)
si_sdr_kernel(const float* __restrict__ ref, const float* __restrict__ est, const int* __restrict__ len, int lstride,
              double* __restrict__ out) {
    __shared__ double sm[8];
    const int b = blockIdx.x;
    const int L = len ? len[b] : lstride;
    const float* s = ref + (int64_t)b * lstride;
    const float* e = est + (int64_t)b * lstride;
    double dot = 0.0, ss = 0.0;
    for (int i = threadIdx.x; i < L; i += 256) {
        const double a = s[i], c = e[i];
        dot += a * c;
        ss += a * a;
    }
    dot = block_sum_256(dot, sm);
    ss = block_sum_256(ss, sm);
    const double alpha = dot / ss;
    double num = 0.0, den = 0.0;
    for (int i = threadIdx.x; i < L; i += 256) {
        const double a = alpha * (double)s[i], d = a - (double)e[i];
        num += a * a;
        den += d * d;
    }
    num = block_sum_256(num, sm);
    den = block_sum_256(den, sm);
    if (threadIdx.x == 0) out[b] = 10.0 * log10(num / den);
}

}  // namespace

int si_sdr_launch(const float* ref, const float* est, const int* len, int B, int lstride, double* out, cudaStream_t s) {
    SNRSE_CHECK_ARG(ref && est && out && B > 0 && lstride > 0, "si_sdr: bad arguments");
    si_sdr_kernel<<<B, 256, 0, s>>>(ref, est, len, lstride, out);
    SNRSE_LAUNCH_CHECK();
    return SNRSE_OK;
}


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
    (void)frames_ws;  // kept in the ABI; the overlap-add happens in shared memory
    static bool attr_set = false;
    if (!attr_set) {
        SNRSE_CUDA(cudaFuncSetAttribute(istft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ISTFT_SMEM));
        attr_set = true;
    }
    dim3 grid(cdiv(tpad, NEWF), B);
    istft_kernel<<<grid, 256, ISTFT_SMEM, s>>>(spec, len, scale, wave, lstride, tpad, transform, alpha, beta);
    SNRSE_LAUNCH_CHECK();
    return SNRSE_OK;
}

int absmax_launch(const float* wave, const int* len, int B, int lstride, float* out, cudaStream_t s) {
    absmax_kernel<<<B, 256, 0, s>>>(wave, len, lstride, out);
    SNRSE_LAUNCH_CHECK();
    return SNRSE_OK;
}

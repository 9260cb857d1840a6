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
// output index k with (k mod 3, k mod 5, k mod 17) = (k1, k2, k3).  Each 17-, 5- and 3-point DFT is one thread's
// register-resident butterfly (inputs loaded once, outputs stored in place) that uses the x_j +- x_{R-j} symmetry:
// ~8.3 k flops per 510-sample frame (a direct DFT needs 522 k), no shared-memory traffic inside the butterflies,
// twiddles are compile-time immediates.  Eight frames per pass.
// The inverse kernel also does the overlap-add, the window-envelope division and the rescale: every block
// recomputes a 4-frame halo so that all frames touching its output samples are in its own shared memory; no
// frame workspace in HBM, no floating-point atomics.
#include "kernels.h"

namespace {

constexpr int NFFT = 510, HOP = 128, NBINS = 256, FT = 8, HALF = 255, M255 = 255;
constexpr int FS = 256;                       // float2 stride between the frames of a transform buffer

// gain g with  spec_fwd(X) = g(|X|) X  /  spec_back(S) = g(|S / beta|) S / beta   (data_module.py:241-267):
//   transform 1 "exponent": beta |X|^(alpha-1)   /  |S|^(1/alpha - 1)
//   transform 2 "log"     : beta log(1+|X|)/|X|  /  (exp|S| - 1)/|S|
// 0 -> 0 in every case (abs 0, angle 0 in the reference).
// the log variant's log1pf / expm1f are kept out of line: inlined they cost the STFT kernels 16 registers (48 -> 64) and a
// block of occupancy on the path every checkpoint uses (exponent)
__device__ __noinline__ float spec_gain_log_fwd(float mag, float beta) { return beta * log1pf(mag) / mag; }
__device__ __noinline__ float spec_gain_log_back(float mag) { return expm1f(mag) / mag; }
__device__ __forceinline__ float spec_gain_fwd(float mag, int transform, float alpha, float beta) {
    if (!(mag > 0.f)) return 0.f;
    if (transform == 2) return spec_gain_log_fwd(mag, beta);
    if (alpha == 1.0f) return beta;
    return (alpha == 0.5f) ? beta / sqrtf(mag) : beta * powf(mag, alpha - 1.0f);
}
__device__ __forceinline__ float spec_gain_back(float mag, int transform, float alpha) {
    if (!(mag > 0.f)) return 0.f;
    if (transform == 2) return spec_gain_log_back(mag);
    if (alpha == 1.0f) return 1.0f;
    return (alpha == 0.5f) ? mag : powf(mag, 1.0f / alpha - 1.0f);
}
// The configuration every checkpoint uses (transform "exponent", alpha = 0.5) without IEEE square roots / divisions and
// without per-element branching on the transform parameters (the generic path above cost ~100 instructions per bin and
// frame: more than half of the forward kernel's instruction count).  m2 = |X|^2.
//   forward : g = beta |X|^-1/2 = beta * rsqrt(sqrt(m2));   inverse: g = |S/beta| = sqrt(m2).   0 -> 0.
// rsqrt.approx has a relative error of 2^-22.9; |X|^2 below 1e-30 (|X| < 1e-15) is treated as 0.
__device__ __forceinline__ float fast_rsqrt(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float gain_fwd_half(float m2, float beta) {
    const float mag = m2 * fast_rsqrt(m2);
    return m2 > 1e-30f ? beta * fast_rsqrt(mag) : 0.f;
}
__device__ __forceinline__ float gain_back_half(float m2) { return m2 > 1e-30f ? m2 * fast_rsqrt(m2) : 0.f; }
constexpr int SPAN = (FT - 1) * HOP + NFFT;   // samples touched by FT consecutive frames (1406)

__device__ constexpr float C3[3] = {1.f, -0.5f, -0.5f};
__device__ constexpr float S3[3] = {0.f, 0.866025404f, -0.866025404f};
__device__ constexpr float C5[5] = {1.f, 0.309016994f, -0.809016994f, -0.809016994f, 0.309016994f};
__device__ constexpr float S5[5] = {0.f, 0.951056516f, 0.587785252f, -0.587785252f, -0.951056516f};
__device__ constexpr float C17[17] = {1.f, 0.932472229f, 0.739008917f, 0.445738356f, 0.0922683595f, -0.27366299f,
                                      -0.602634636f, -0.850217136f, -0.9829731f, -0.9829731f, -0.850217136f,
                                      -0.602634636f, -0.27366299f, 0.0922683595f, 0.445738356f, 0.739008917f, 0.932472229f};
__device__ constexpr float S17[17] = {0.f, 0.361241666f, 0.673695644f, 0.895163291f, 0.995734176f, 0.961825643f,
                                      0.798017227f, 0.526432163f, 0.183749518f, -0.183749518f, -0.526432163f,
                                      -0.798017227f, -0.961825643f, -0.995734176f, -0.895163291f, -0.673695644f, -0.361241666f};

__device__ __forceinline__ float hann(int n) { return 0.5f - 0.5f * cospif((float)(2 * n) / (float)NFFT); }

// R-point DFT (R odd) of v in registers, e^{-...} (INV = false) or e^{+...} (INV = true), unnormalised.
//   a_j = v_j + v_{R-j}, b_j = v_j - v_{R-j};  P_k = v_0 + sum_j a_j cos(2 pi j k / R);  Q_k = sum_j b_j sin(2 pi j k / R)
//   forward: V_k = P_k - i Q_k, V_{R-k} = P_k + i Q_k.
template <int R, bool INV>
__device__ __forceinline__ void bfly(float2 (&v)[R], const float (&C)[R], const float (&S)[R]) {
    constexpr int H = R / 2;
    float2 a[H], b[H];
#pragma unroll
    for (int j = 1; j <= H; ++j) {
        a[j - 1] = make_float2(v[j].x + v[R - j].x, v[j].y + v[R - j].y);
        b[j - 1] = make_float2(v[j].x - v[R - j].x, v[j].y - v[R - j].y);
    }
    const float2 x0 = v[0];
    float2 s0 = x0;
#pragma unroll
    for (int j = 0; j < H; ++j) {
        s0.x += a[j].x;
        s0.y += a[j].y;
    }
    v[0] = s0;
#pragma unroll
    for (int k = 1; k <= H; ++k) {
        float pr = x0.x, pi = x0.y, qr = 0.f, qi = 0.f;
#pragma unroll
        for (int j = 1; j <= H; ++j) {
            const int m = (j * k) % R;
            pr = fmaf(a[j - 1].x, C[m], pr);
            pi = fmaf(a[j - 1].y, C[m], pi);
            qr = fmaf(b[j - 1].x, S[m], qr);
            qi = fmaf(b[j - 1].y, S[m], qi);
        }
        const float2 lo = make_float2(pr + qi, pi - qr), hi = make_float2(pr - qi, pi + qr);   // P - iQ, P + iQ
        v[k] = INV ? hi : lo;
        v[R - k] = INV ? lo : hi;
    }
}

// natural position of the element stored at linear index j = (n1*5 + n2)*17 + n3 of the transform input
__device__ __forceinline__ int pfa_input_pos(int j) {
    const int g = j / 17, n3 = j - g * 17, n1 = g / 5, n2 = g - n1 * 5;
    return (85 * n1 + 51 * n2 + 15 * n3) % M255;
}
// its inverse: linear input index of natural position p  (85 = 1 mod 3, 51 = 1 mod 5, 15 * 8 = 1 mod 17)
__device__ __forceinline__ int pfa_input_index(int p) { return ((p % 3) * 5 + (p % 5)) * 17 + (8 * p) % 17; }
// linear index (k1*5 + k2)*17 + k3 at which output bin k is found after the transform
__device__ __forceinline__ int pfa_output_index(int k) { return (k % 3) * 85 + (k % 5) * 17 + (k % 17); }

// In-place 255-point DFT of FT frames: X[f*FS + j], j in PFA input order (pfa_input_pos); on return bin k of frame f is
// at X[f*FS + pfa_output_index(k)].  The caller synchronises after filling X; the result is visible on return.
template <bool INV>
__device__ __forceinline__ void fft255(float2* __restrict__ X) {
    const int tid = threadIdx.x;
    if (tid < 15 * FT) {  // 17-point DFTs along n3: 15 (n1, n2) groups per frame
        const int f = tid / 15, g = tid - f * 15;
        float2* p = X + f * FS + g * 17;
        float2 v[17];
#pragma unroll
        for (int n = 0; n < 17; ++n) v[n] = p[n];
        bfly<17, INV>(v, C17, S17);
#pragma unroll
        for (int n = 0; n < 17; ++n) p[n] = v[n];
    }
    __syncthreads();
    for (int it = tid; it < 51 * FT; it += 256) {  // 5-point DFTs along n2: 51 (n1, k3) columns per frame
        const int f = it / 51, col = it - f * 51, n1 = col / 17, k3 = col - n1 * 17;
        float2* p = X + f * FS + n1 * 85 + k3;
        float2 v[5];
#pragma unroll
        for (int n = 0; n < 5; ++n) v[n] = p[n * 17];
        bfly<5, INV>(v, C5, S5);
#pragma unroll
        for (int n = 0; n < 5; ++n) p[n * 17] = v[n];
    }
    __syncthreads();
    for (int it = tid; it < 85 * FT; it += 256) {  // 3-point DFTs along n1: 85 (k2, k3) columns per frame
        const int f = it / 85, r = it - f * 85;
        float2* p = X + f * FS + r;
        float2 v[3];
#pragma unroll
        for (int n = 0; n < 3; ++n) v[n] = p[n * 85];
        bfly<3, INV>(v, C3, S3);
#pragma unroll
        for (int n = 0; n < 3; ++n) p[n * 85] = v[n];
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256)
stft_kernel(const float* __restrict__ wave, const int* __restrict__ len, const float* __restrict__ scale,
            int scale_is_divisor, float* __restrict__ out, int lstride, int tpad, int transform, float alpha,
            float beta, int planar, int npass) {
    pdl_sync();
    __shared__ __align__(16) float2 X[FT * FS];
    __shared__ __align__(16) float xs[SPAN + 2];
    const int b = blockIdx.y, tid = threadIdx.x;
    const int L = len ? len[b] : lstride;
    const int nframes = 1 + L / HOP;
    float sc = 1.0f;
    if (scale) sc = scale[b];
    const float* wb = wave + (int64_t)b * lstride;
    // per-thread constants, reused by every pass: packed-input position + window (thread = transform point),
    // split indices + e^{-i pi k/255} = (ec, -es) (thread = output bin k)
    const int m2 = 2 * pfa_input_pos(tid < M255 ? tid : 0);
    const float w0 = hann(m2), w1 = hann(m2 + 1);
    // thread <-> output bin: thread t < 255 takes the bin stored at position t after the transform (k = CRT of
    // (t/85, (t%85)/17, t%17)), so the split reads X[.. + t] without bank conflicts and the partner bin 255-k at a
    // descending stride of 1 (53 -> 17 and 53 -> 35 wavefronts per 256 threads); every thread writes its own row of
    // the output, so the permutation costs nothing on the global side.  Thread 255 takes the Nyquist bin.
    const int k = tid < M255 ? (85 * (tid / 85) + 51 * ((tid % 85) / 17) + 120 * (tid % 17)) % M255 : M255;
    const int ia = pfa_output_index(k == M255 ? 0 : k), ib = pfa_output_index((M255 - k) % M255);
    float es, ec;
    sincospif((float)k / (float)M255, &es, &ec);
    const bool fast_half = transform == 1 && alpha == 0.5f;
    const float rsc = (scale && scale_is_divisor) ? 1.0f / sc : sc;     // y / norm as y * (1 / norm): one rounding apart
    for (int pass = 0; pass < npass; ++pass) {
        const int f0 = (blockIdx.x * npass + pass) * FT;
        if (f0 >= tpad) break;                                   // block-uniform
        if (pass) __syncthreads();                               // the previous pass has read X
        const int s0 = f0 * HOP - HALF;
        if (s0 >= 0 && s0 + SPAN <= L) {                         // interior: no reflection
            for (int e = tid; e < SPAN; e += 256) xs[e] = wb[s0 + e] * rsc;
        } else {
            for (int e = tid; e < SPAN; e += 256) {
                int i = s0 + e;
                if (i < 0) i = -i;
                if (i >= L) i = 2 * (L - 1) - i;
                i = max(0, min(i, L - 1));  // beyond one reflection: only frames >= nframes (written as zeros) or len <= 255
                xs[e] = wb[i] * rsc;
            }
        }
        __syncthreads();
        if (tid < M255) {  // windowed, packed z[m] = x[2m] + i x[2m+1] in PFA input order
#pragma unroll
            for (int f = 0; f < FT; ++f) {
                const float2 x = *reinterpret_cast<const float2*>(&xs[f * HOP + m2]);
                X[f * FS + tid] = make_float2(x.x * w0, x.y * w1);
            }
        }
        __syncthreads();
        fft255<false>(X);
        // split: X[k] = (Z[k] + conj Z[255-k])/2 - (i/2) e^{-2 pi i k/510} (Z[k] - conj Z[255-k]),  k = 0..255
        float re[FT], im[FT];
#pragma unroll
        for (int f = 0; f < FT; ++f) {
            const float2 za = X[f * FS + ia], zb = X[f * FS + ib];
            const float sr = za.x + zb.x, si = za.y - zb.y, dr = za.x - zb.x, di = za.y + zb.y;   // Z[k] +- conj Z[255-k]
            // E*d with E = (ec, -es):  (ec dr + es di,  ec di - es dr);  -i (a + i b) = b - i a
            const float pr = fmaf(ec, dr, es * di), pi = fmaf(ec, di, -es * dr);
            float r = 0.5f * (sr + pi), i = 0.5f * (si - pr);
            if (f0 + f >= nframes) {
                r = 0.f;
                i = 0.f;
            } else if (fast_half) {
                const float g = gain_fwd_half(fmaf(r, r, i * i), beta);
                r *= g;
                i *= g;
            } else if (transform != 0) {
                // beta * |X|^alpha * e^{i arg X} = X * beta * |X|^(alpha-1), 0 -> 0  (data_module.py:241-251)
                const float g = spec_gain_fwd(sqrtf(r * r + i * i), transform, alpha, beta);
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
}

// Inverse: block x owns frames F0 = x*NEWF .. F0+NEWF-1 and output samples p in [128 F0, 128 (F0+NEWF)) (p = sample
// index in the 255-padded signal); it transforms frames F0-HALO .. F0+NEWF-1 in PASSES passes of FT frames, overlap-adds
// them in shared memory in ascending frame order, divides by the window envelope and writes wave[p - 255].
constexpr int PASSES = 4, HALO = 4, NEWF = PASSES * FT - HALO;      // 28 new frames per block
constexpr int OLA_SPAN = (PASSES * FT - 1) * HOP + NFFT;             // 4478 samples
constexpr int TS = 258;                                              // float2 stride between frames of the spectrogram tile
constexpr int ISTFT_SMEM = (FT * FS + FT * TS) * 8 + OLA_SPAN * 4 + NFFT * 4 + 64;

__global__ void __launch_bounds__(256)
istft_kernel(const float2* __restrict__ spec, const int* __restrict__ len, const float* __restrict__ scale,
             float* __restrict__ wave, int lstride, int tpad, int transform, float alpha, float beta) {
    pdl_sync();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* X = reinterpret_cast<float2*>(smem_raw);                 // transform buffer [FT][FS]
    float2* T = X + FT * FS;                                          // spectrogram tile [FT][TS], later frames [FT][512 floats]
    float* ola = reinterpret_cast<float*>(T + FT * TS);
    float* hw = ola + OLA_SPAN;                                       // Hann window
    const int b = blockIdx.y, F0 = blockIdx.x * NEWF, tid = threadIdx.x;
    const int L = len ? len[b] : lstride;
    for (int i = tid; i < NFFT; i += 256) hw[i] = hann(i);
    for (int i = tid; i < OLA_SPAN; i += 256) ola[i] = 0.f;
    const float2* sb = spec + (int64_t)b * NBINS * tpad;
    const float inv_beta = 1.0f / beta;
    const bool fast_half = transform == 1 && alpha == 0.5f;
    const bool vec_ok = (tpad & 1) == 0;
    // merge: thread t < 255 produces transform input element t, i.e. Z'[k] for k = pfa_input_pos(t), from the bins
    // k and 255-k (k = 0 pairs with the Nyquist row 255):  e^{+i pi k/255} = (ec, es)
    const int mk = pfa_input_pos(tid < M255 ? tid : 0), mk2 = M255 - mk;
    float es, ec;
    sincospif((float)mk / (float)M255, &es, &ec);
    for (int pass = 0; pass < PASSES; ++pass) {
        const int fb = F0 - HALO + pass * FT;                        // first frame of this pass (even)
        if (fb + FT <= 0 || fb >= tpad) continue;                     // block-uniform: nothing to add
        __syncthreads();                                              // previous pass done with X / T
        // spectrogram tile -> T as float2 [frame][TS] (bin-contiguous: the merge below reads bins at a stride of 15,
        // conflict-free; the former [bin][FT] layout made it an 8-way bank conflict), spec_back applied
        // (data_module.py:256-262)
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
            if (fast_half) {
                v.x *= inv_beta; v.y *= inv_beta; v.z *= inv_beta; v.w *= inv_beta;
                const float g0 = gain_back_half(fmaf(v.x, v.x, v.y * v.y)), g1 = gain_back_half(fmaf(v.z, v.z, v.w * v.w));
                v.x *= g0; v.y *= g0; v.z *= g1; v.w *= g1;
            } else if (transform != 0) {
                v.x *= inv_beta; v.y *= inv_beta; v.z *= inv_beta; v.w *= inv_beta;
                const float g0 = spec_gain_back(sqrtf(v.x * v.x + v.y * v.y), transform, alpha);
                const float g1 = spec_gain_back(sqrtf(v.z * v.z + v.w * v.w), transform, alpha);
                v.x *= g0; v.y *= g0; v.z *= g1; v.w *= g1;
            }
            if (k == 0 || k == NBINS - 1) { v.y = 0.f; v.w = 0.f; }  // one-sided inverse ignores Im of DC / Nyquist
            T[(2 * c) * TS + k] = make_float2(v.x, v.y);
            T[(2 * c + 1) * TS + k] = make_float2(v.z, v.w);
        }
        __syncthreads();
        // merge: Z'[k] = (S[k] + conj S[255-k]) + i e^{+2 pi i k/510} (S[k] - conj S[255-k]),  k = 0..254, one thread per
        // element, written in the transform's input order (X[.. + tid]: conflict-free)
        if (tid < M255) {
#pragma unroll
            for (int f = 0; f < FT; ++f) {
                const float2 xa = T[f * TS + mk], xb = T[f * TS + mk2];
                const float sr = xa.x + xb.x, si = xa.y - xb.y, dr = xa.x - xb.x, di = xa.y + xb.y;
                const float pr = fmaf(ec, dr, -es * di), pi = fmaf(ec, di, es * dr);   // E * d
                X[f * FS + tid] = make_float2(sr - pi, si + pr);
            }
        }
        __syncthreads();
        fft255<true>(X);
        // back to natural order: frame f becomes 510 consecutive floats (x[2m] = Re z[m], x[2m+1] = Im z[m])
        if (tid < M255) {
            const int j = pfa_output_index(tid);
#pragma unroll
            for (int f = 0; f < FT; ++f) T[f * FS + tid] = X[f * FS + j];
        }
        __syncthreads();
        // overlap-add the FT frames of this pass (ascending frame order), windowed and scaled by 1/510
        const float* zf = reinterpret_cast<const float*>(T);
        for (int s = tid; s < SPAN; s += 256) {
            int fhi = s / HOP;
            if (fhi > FT - 1) fhi = FT - 1;
            int flo = (s - (NFFT - 1) + HOP - 1) / HOP;
            if (s - (NFFT - 1) <= 0) flo = 0;
            float acc = 0.f;
            for (int f = flo; f <= fhi; ++f) {
                const int n = s - f * HOP;
                acc = fmaf(zf[f * 2 * FS + n], hw[n], acc);
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

// max |y| per utterance.  One block per 8192-sample chunk (a single block per utterance walked 4 s of audio in 250
// dependent steps: 116 us for 4 MB); block maxima are merged with an integer atomicMax on the float's bit pattern
// (non-negative floats order like unsigned integers), which is exact and order-independent.  out[] is zeroed first.
constexpr int ABSMAX_CHUNK = 8192;
__global__ void __launch_bounds__(256)
absmax_kernel(const float* __restrict__ wave, const int* __restrict__ len, int lstride, float* __restrict__ out) {
    pdl_sync();
    __shared__ float sm[8];
    const int b = blockIdx.y;
    const int L = len ? len[b] : lstride;
    const int i0 = blockIdx.x * ABSMAX_CHUNK, i1 = min(i0 + ABSMAX_CHUNK, L);
    if (i0 >= L) return;                                         // block-uniform
    const float* w = wave + (int64_t)b * lstride;
    float m = 0.f;
    for (int i = i0 + threadIdx.x; i < i1; i += 256) m = fmaxf(m, fabsf(__ldg(w + i)));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) m = fmaxf(m, sm[i]);
        atomicMax(reinterpret_cast<unsigned int*>(out + b), __float_as_uint(m));
    }
}

// stand-alone spec_fwd / spec_back (data_module.py:241-267) for callers that hold a raw STFT
__global__ void __launch_bounds__(256)
spec_transform_kernel(const float2* __restrict__ in, float2* __restrict__ out, int64_t n, int inverse, int transform,
                      float alpha, float beta) {
    pdl_sync();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float2 z = in[i];
    if (inverse) {
        z.x /= beta;
        z.y /= beta;
        const float g = spec_gain_back(sqrtf(z.x * z.x + z.y * z.y), transform, alpha);
        z.x *= g;
        z.y *= g;
    } else {
        const float g = spec_gain_fwd(sqrtf(z.x * z.x + z.y * z.y), transform, alpha, beta);
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
    pdl_sync();
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
    snrse_launch(si_sdr_kernel, dim3(B), dim3(256), 0, s, ref, est, len, lstride, out);
    SNRSE_LAUNCH_CHECK();
    return SNRSE_OK;
}


int spec_transform_launch(const float2* in, float2* out, int64_t n, int inverse, int transform, float alpha, float beta,
                          cudaStream_t s) {
    SNRSE_CHECK_ARG(transform == 1 || transform == 2, "spec_transform: transform must be 1 (exponent) or 2 (log)");
    snrse_launch(spec_transform_kernel, dim3((unsigned)cdiv64(n, 256)), dim3(256), 0, s, in, out, n, inverse, transform, alpha, beta);
    SNRSE_LAUNCH_CHECK();
    return SNRSE_OK;
}

int stft_launch(const float* wave, const int* len, const float* scale, int scale_is_divisor, float* out, int B,
                int lstride, int tpad, int transform, float alpha, float beta, int planar, cudaStream_t s) {
    SNRSE_CHECK_ARG(B > 0 && tpad > 0 && lstride > HALF, "stft: need B>0, Tpad>0 and more than 255 samples");
    SNRSE_CHECK_ARG(transform >= 0 && transform <= 2, "stft: transform must be 0 (none), 1 (exponent) or 2 (log)");
    // one pass of FT frames per block.  (2-4 passes per block, which amortise the per-thread window / twiddle / index
    // constants, measured the same time on 64 x 60 s -- 1.07 ms either way: the kernel is bound by issue slots and by
    // load / barrier latency, not by those constants -- and a longer critical path on small problems.)
    const int npass = 1;
    dim3 grid(cdiv(tpad, FT * npass), B);
    snrse_launch(stft_kernel, dim3(grid), dim3(256), 0, s, wave, len, scale, scale_is_divisor, out, lstride, tpad, transform, alpha, beta, planar,
                                     npass);
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
    snrse_launch(istft_kernel, dim3(grid), dim3(256), ISTFT_SMEM, s, spec, len, scale, wave, lstride, tpad, transform, alpha, beta);
    SNRSE_LAUNCH_CHECK();
    return SNRSE_OK;
}

int absmax_launch(const float* wave, const int* len, int B, int lstride, float* out, cudaStream_t s) {
    SNRSE_CUDA(cudaMemsetAsync(out, 0, (size_t)B * sizeof(float), s));
    snrse_launch(absmax_kernel, dim3(dim3((unsigned)cdiv(lstride, ABSMAX_CHUNK), (unsigned)B)), dim3(256), 0, s, wave, len, lstride, out);
    SNRSE_LAUNCH_CHECK();
    return SNRSE_OK;
}

// C-ABI entry points for the stand-alone operators (declared in include/snrse_b200.h).
// The NCSN++ executor entry points live in engine.cu, the SNR estimator's in snrnet.cu.
#include <stdarg.h>
#include <stdlib.h>

#include <vector>

#include "kernels.h"

static thread_local char g_err[512] = "";
long long g_snrse_launches = 0;
static int pdl_default() {
    const char* e = getenv("SNRSE_PDL");
    return e ? (atoi(e) & 7) : 2;
}
int g_snrse_pdl = pdl_default();

void snrse_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }

static ActView mk_view(const void* p, int B, int H, int W, int C, int ld) {
    ActView v;
    v.ptr = const_cast<bf16*>(static_cast<const bf16*>(p));
    v.B = B; v.H = H; v.W = W; v.C = C; v.ld = ld;
    return v;
}

extern "C" {

int snrse_version(void) { return 100; }
long long snrse_launch_count(void) { return g_snrse_launches; }
int snrse_set_pdl(int on) {
    const int prev = g_snrse_pdl;
    if (on >= 0) g_snrse_pdl = on & 7;
    return prev;
}
const char* snrse_last_error(void) { return g_err; }

// 0 when the current device is a Blackwell B200-class GPU (compute capability 10.x); the kernels are sm_100a-only.
int snrse_device_check(void) {
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
        snrse_set_error("no CUDA device available: the B200 library has no CPU fallback");
        return SNRSE_ERR_CUDA;
    }
    if (prop.major != 10) {
        snrse_set_error("device %s is sm_%d%d; this library is built for sm_100a only", prop.name, prop.major, prop.minor);
        return SNRSE_ERR_UNSUPPORTED;
    }
    return SNRSE_OK;
}

int snrse_stft(const float* wave, const int* len, const float* scale, int scale_is_divisor, void* out, int B,
               int lstride, int tpad, int transform, float alpha, float beta, int planar, void* stream) {
    SNRSE_CHECK_ARG(wave && out, "stft: null pointer");
    return stft_launch(wave, len, scale, scale_is_divisor, static_cast<float*>(out), B, lstride, tpad, transform, alpha,
                       beta, planar, S(stream));
}

// the overlap-add runs in shared memory since the FFT kernels; a token workspace keeps the ABI and its callers unchanged
int64_t snrse_istft_workspace_bytes(int B, int tpad) { (void)B; (void)tpad; return 256; }

int snrse_istft(const void* spec, const int* len, const float* scale, float* wave, void* workspace, int B, int lstride,
                int tpad, int transform, float alpha, float beta, void* stream) {
    SNRSE_CHECK_ARG(spec && wave && workspace, "istft: null pointer");
    return istft_launch(static_cast<const float2*>(spec), len, scale, wave, static_cast<float*>(workspace), B, lstride,
                        tpad, transform, alpha, beta, S(stream));
}

int snrse_spec_transform(const void* in, void* out, int64_t n, int inverse, int transform, float alpha, float beta,
                         void* stream) {
    SNRSE_CHECK_ARG(in && out, "spec_transform: null pointer");
    return spec_transform_launch(static_cast<const float2*>(in), static_cast<float2*>(out), n, inverse, transform, alpha,
                                 beta, S(stream));
}

int snrse_si_sdr(const float* ref, const float* est, const int* len, int B, int lstride, double* out, void* stream) {
    return si_sdr_launch(ref, est, len, B, lstride, out, S(stream));
}

int snrse_absmax(const float* wave, const int* len, int B, int lstride, float* out, void* stream) {
    SNRSE_CHECK_ARG(wave && out, "absmax: null pointer");
    return absmax_launch(wave, len, B, lstride, out, S(stream));
}

int snrse_v3_scalars(const float* ratio, const float* peak, double snr_scale, float nf_const, const double* t30,
                     float* t_out, float* nf_out, int* idx_out, int B, void* stream) {
    SNRSE_CHECK_ARG(ratio && peak && t30 && t_out && nf_out, "v3_scalars: null pointer");
    return v3_scalars_launch(ratio, peak, snr_scale, nf_const, t30, t_out, nf_out, idx_out, B, S(stream));
}

int snrse_snr_ratio(const float* g, float* ratio, int B, void* stream) { return snr_ratio_launch(g, ratio, B, S(stream)); }

int snrse_lincomb(const void* x, const void* y, const void* sc, const void* z, const float* a, const float* b,
                  const float* c, const float* d, void* out_mean, void* out_x, int B, int64_t n, void* stream) {
    return lincomb_launch(static_cast<const float2*>(x), static_cast<const float2*>(y), static_cast<const float2*>(sc),
                          static_cast<const float2*>(z), a, b, c, d, static_cast<float2*>(out_mean),
                          static_cast<float2*>(out_x), B, n, S(stream));
}

int snrse_rk_combine(const void* y, const void* K, int nk, int64_t n, float h, const float* coef, void* out, void* stream) {
    return rk_combine_launch(static_cast<const float2*>(y), static_cast<const float2*>(K), nk, n, h, coef,
                             static_cast<float2*>(out), S(stream));
}

int snrse_rk_partials(int64_t n) { return rk_partials(n); }

int snrse_rk_scaled_sqnorm(const void* K, int nk, int64_t n, float h, const float* coef, const void* y, const void* y2,
                           float atol, float rtol, double* partial, void* stream) {
    return rk_scaled_sqnorm_launch(static_cast<const float2*>(K), nk, n, h, coef, static_cast<const float2*>(y),
                                   static_cast<const float2*>(y2), atol, rtol, partial, S(stream));
}

// ---- single-operator entry points (NHWC bf16), used by the parity tests and usable as drop-in ops
int snrse_conv_nhwc(const void* x0, int c0, int taps0, const void* x1, int c1, const void* wt, int n, const float* bias,
                    const float* tbias, int tb_stride, const void* res, float scale, void* out, int B, int H, int W,
                    int impl, void* stream) {
    SNRSE_CHECK_ARG(x0 && wt && out, "conv_nhwc: null pointer");
    ActView a0 = mk_view(x0, B, H, W, c0, c0), a1, r;
    if (x1) a1 = mk_view(x1, B, H, W, c1, c1);
    if (res) r = mk_view(res, B, H, W, n, n);
    if (impl == 1) {
        SNRSE_CHECK_ARG(n <= 256, "conv_nhwc: CUDA-core cross-check path handles N <= 256");
        return conv_simt_launch(&a0, taps0, x1 ? &a1 : nullptr, static_cast<const bf16*>(wt), n, bias, tbias, tb_stride,
                                res ? &r : nullptr, scale, static_cast<bf16*>(out), n, S(stream));
    }
    if (impl == 0 && conv_halo2_eligible(&a0, taps0, n)) {
        ConvHaloPlan hp;
        SNRSE_TRY(conv_halo2_make_plan(&hp, &a0, x1 ? &a1 : nullptr, static_cast<const bf16*>(wt), n, bias, tbias, tb_stride,
                                       res ? &r : nullptr, scale, static_cast<bf16*>(out), n, nullptr, nullptr));
        return conv_halo2_launch(&hp, S(stream));
    }
    ConvGemmPlan p;
    SNRSE_TRY(conv_gemm_make_plan(&p, &a0, taps0, x1 ? &a1 : nullptr, static_cast<const bf16*>(wt), n, 0, 0, bias, tbias,
                                  tb_stride, res ? &r : nullptr, scale, out, n, 0));
    return conv_gemm_launch(&p, S(stream));
}

// conv3x3 on the 2-CTA kernel that also accumulates the GroupNorm sums of its result (what the NCSN++ executor uses so
// that GroupNorm needs no pass of its own): ustats [B][n/4][2] 64-bit fixed point (sum * 2^30, sum of squares * 2^24
// per 4-channel unit), zeroed here first.
int snrse_conv3x3_nhwc_stats(const void* x0, int c0, const void* x1, int c1, const void* wt, int n, const float* bias,
                             const float* tbias, int tb_stride, const void* res, float scale, void* out, int B, int H,
                             int W, void* ustats, void* stream) {
    SNRSE_CHECK_ARG(x0 && wt && out && ustats, "conv3x3_stats: null pointer");
    ActView a0 = mk_view(x0, B, H, W, c0, c0), a1, r;
    if (x1) a1 = mk_view(x1, B, H, W, c1, c1);
    if (res) r = mk_view(res, B, H, W, n, n);
    SNRSE_CHECK_ARG(conv_halo2_eligible(&a0, 9, n), "conv3x3_stats: needs W >= 8, H >= 8, N in {128, 256}");
    SNRSE_CUDA(cudaMemsetAsync(ustats, 0, (size_t)B * (n / 4) * 16, S(stream)));
    ConvHaloPlan hp;
    SNRSE_TRY(conv_halo2_make_plan(&hp, &a0, x1 ? &a1 : nullptr, static_cast<const bf16*>(wt), n, bias, tbias, tb_stride,
                                   res ? &r : nullptr, scale, static_cast<bf16*>(out), n, nullptr,
                                   static_cast<unsigned long long*>(ustats)));
    return conv_halo2_launch(&hp, S(stream));
}

int64_t snrse_groupnorm_workspace_bytes(int B) { return (int64_t)B * (128 * 16 + 1024 * 4); }

int snrse_groupnorm_nhwc(const void* x, const float* gamma, const float* beta, void* out, int B, int H, int W, int C,
                         int silu, float eps, void* workspace, void* stream) {
    SNRSE_CHECK_ARG(x && gamma && beta && out && workspace, "groupnorm: null pointer");
    const ActView vx = mk_view(x, B, H, W, C, C), vo = mk_view(out, B, H, W, C, C);
    unsigned long long* ust = static_cast<unsigned long long*>(workspace);
    float* scsh = reinterpret_cast<float*>(ust + (int64_t)B * 128 * 2);
    const int64_t hw = (int64_t)H * W;
    SNRSE_CUDA(cudaMemsetAsync(ust, 0, (size_t)B * (C / 4) * 16, S(stream)));
    SNRSE_TRY(gn_stats_launch(&vx, ust, S(stream)));
    SNRSE_TRY(gn_finalize_launch(ust, C / 4, nullptr, 0, B, hw * (C / 32), gamma, beta, eps, scsh, S(stream)));
    return gn_apply_launch(&vx, scsh, silu, &vo, S(stream));
}

// GroupNorm(32, eps) + SiLU + conv3x3 (+1x1 shortcut / bias / time-embedding bias / residual, * scale) with the
// normalisation applied to the operand inside the convolution kernel: the body of ResnetBlockBigGANpp
// (ncsnpp_utils/layerspp.py:245-271) without materialising the normalised activation.
int snrse_gn_silu_conv3x3_nhwc(const void* x0, int c0, const float* gamma, const float* beta, float eps, const void* x1,
                               int c1, const void* wt, int n, const float* bias, const float* tbias, int tb_stride,
                               const void* res, float scale, void* out, int B, int H, int W, void* workspace,
                               void* stream) {
    SNRSE_CHECK_ARG(x0 && gamma && beta && wt && out && workspace, "gn_silu_conv3x3: null pointer");
    ActView a0 = mk_view(x0, B, H, W, c0, c0), a1, r;
    if (x1) a1 = mk_view(x1, B, H, W, c1, c1);
    if (res) r = mk_view(res, B, H, W, n, n);
    SNRSE_CHECK_ARG(conv_halo2_eligible(&a0, 9, n), "gn_silu_conv3x3: needs W >= 8, H >= 8, N in {128, 256}");
    unsigned long long* ust = static_cast<unsigned long long*>(workspace);
    float* scsh = reinterpret_cast<float*>(ust + (int64_t)B * 128 * 2);
    const int64_t hw = (int64_t)H * W;
    SNRSE_CUDA(cudaMemsetAsync(ust, 0, (size_t)B * (c0 / 4) * 16, S(stream)));
    SNRSE_TRY(gn_stats_launch(&a0, ust, S(stream)));
    SNRSE_TRY(gn_finalize_launch(ust, c0 / 4, nullptr, 0, B, hw * (c0 / 32), gamma, beta, eps, scsh, S(stream)));
    ConvHaloPlan hp;
    SNRSE_TRY(conv_halo2_make_plan(&hp, &a0, x1 ? &a1 : nullptr, static_cast<const bf16*>(wt), n, bias, tbias, tb_stride,
                                   res ? &r : nullptr, scale, static_cast<bf16*>(out), n, scsh, nullptr));
    return conv_halo2_launch(&hp, S(stream));
}

int snrse_fir_nhwc(const void* x, void* out, int B, int H, int W, int C, int up, void* stream) {
    SNRSE_CHECK_ARG(x && out, "fir: null pointer");
    const ActView vx = mk_view(x, B, H, W, C, C);
    const ActView vo = up ? mk_view(out, B, 2 * H, 2 * W, C, C) : mk_view(out, B, H / 2, W / 2, C, C);
    return up ? fir_up2_launch(&vx, &vo, S(stream)) : fir_down2_launch(&vx, &vo, S(stream));
}

// FIR resampling of silu(GroupNorm(x)) with the normalisation applied on load (ResnetBlockBigGANpp up / down blocks,
// ncsnpp_utils/layerspp.py:245-257).  workspace: snrse_groupnorm_workspace_bytes(B).
int snrse_gn_silu_fir_nhwc(const void* x, const float* gamma, const float* beta, float eps, void* out, int B, int H, int W,
                           int C, int up, void* workspace, void* stream) {
    SNRSE_CHECK_ARG(x && gamma && beta && out && workspace, "gn_silu_fir: null pointer");
    const ActView vx = mk_view(x, B, H, W, C, C);
    const ActView vo = up ? mk_view(out, B, 2 * H, 2 * W, C, C) : mk_view(out, B, H / 2, W / 2, C, C);
    unsigned long long* ust = static_cast<unsigned long long*>(workspace);
    float* scsh = reinterpret_cast<float*>(ust + (int64_t)B * 128 * 2);
    SNRSE_CUDA(cudaMemsetAsync(ust, 0, (size_t)B * (C / 4) * 16, S(stream)));
    SNRSE_TRY(gn_stats_launch(&vx, ust, S(stream)));
    SNRSE_TRY(gn_finalize_launch(ust, C / 4, nullptr, 0, B, (int64_t)H * W * (C / 32), gamma, beta, eps, scsh, S(stream)));
    return up ? fir_up2_launch(&vx, &vo, S(stream), scsh) : fir_down2_launch(&vx, &vo, S(stream), scsh);
}

int snrse_upfirdn2d(const float* input, const float* kernel, float* out, int64_t major, int in_h, int in_w, int kernel_h,
                    int kernel_w, int up_x, int up_y, int down_x, int down_y, int pad_x0, int pad_x1, int pad_y0, int pad_y1,
                    void* stream) {
    SNRSE_CHECK_ARG(input && kernel && out, "upfirdn2d: null pointer");
    return upfirdn2d_launch(input, kernel, out, major, in_h, in_w, kernel_h, kernel_w, up_x, up_y, down_x, down_y, pad_x0,
                            pad_x1, pad_y0, pad_y1, S(stream));
}

int snrse_fir_f4(const float* x, float* out, int B, int H, int W, int up, void* stream) {
    SNRSE_CHECK_ARG(x && out, "fir_f4: null pointer");
    return up ? fir_up2_f4_launch(x, out, B, H, W, S(stream)) : fir_down2_f4_launch(x, out, B, H, W, S(stream));
}

int64_t snrse_attention_workspace_bytes(int B, int n, int C) { return attention_workspace_bytes(B, n, C); }

int snrse_attention_nhwc(const void* q, const void* k, const void* v, void* scores, void* out, int B, int n, int C,
                         void* stream) {
    SNRSE_CHECK_ARG(q && k && v && scores && out, "attention: null pointer");
    const ActView vq = mk_view(q, B, 1, n, C, C), vk = mk_view(k, B, 1, n, C, C), vv = mk_view(v, B, 1, n, C, C),
                  vo = mk_view(out, B, 1, n, C, C);
    return attention_launch(&vq, &vk, &vv, scores, &vo, S(stream));
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// measurement hooks (include/snrse_b200_debug.h): not part of the product ABI
long long* g_halo_dbg_shared = nullptr;
// device buffer of [grid][8] cycle counters filled by subsequent conv_halo2 / conv_gemm launches (null: off)
extern "C" void snrse_conv_halo_set_debug(long long* dev_counters) { g_halo_dbg_shared = dev_counters; }
// L2 prefetch of the next tile's boxes in the 2-CTA convolution kernel: 1 (default) on, 0 off (A/B measurements)
extern int g_halo2_prefetch;
extern "C" void snrse_conv_halo_set_prefetch(int on) { g_halo2_prefetch = on; }

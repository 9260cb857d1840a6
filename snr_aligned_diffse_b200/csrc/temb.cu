// Time-embedding path of NCSN++ (sgmse-bbed/sgmse/backbones/ncsnpp.py:256-275) and the per-block
// `Dense_0(act(temb))` projections (ncsnpp_utils/layerspp.py:264-265), evaluated once per forward
// for all 49 residual blocks:  tb[b, row] = dense_w[row, :] . silu(temb[b, :]) + dense_b[row],
// where rows of all blocks are concatenated.  fp32 throughout (tiny GEMVs).
#include "kernels.h"

namespace {

// Gaussian Fourier features of log(t): emb[b] = [sin(p), cos(p)], p = log(t_b) * W * 2*pi  (layerspp.py:32-43),
// evaluated in the reference's order.  grid = B, block = nf.
__global__ void temb_fourier_kernel(const float* __restrict__ t, int nf, const float* __restrict__ fw, float* __restrict__ emb) {
    pdl_sync();
    const int b = blockIdx.x, tid = threadIdx.x;
    if (tid >= nf) return;
    const float proj = logf(t[b]) * fw[tid] * 2.0f * 3.14159265358979323846f;
    emb[(int64_t)b * 2 * nf + tid] = sinf(proj);
    emb[(int64_t)b * 2 * nf + nf + tid] = cosf(proj);
}

// out[b, row] = f(W[row, :] . in[b, :] + bias[row]),  f = identity or SiLU.  One warp per output row: its weights are
// read once (coalesced) into registers and reused for every batch item.  d <= 1024, d % 32 == 0.
template <int SILU>
__global__ void __launch_bounds__(256)
dense_rows_kernel(const float* __restrict__ in, int B, int d, const float* __restrict__ w, const float* __restrict__ bias,
                  int rows, float* __restrict__ out) {
    pdl_sync();
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float wr[32];
    const int per = d >> 5;
    const float* wp = w + (int64_t)row * d;
#pragma unroll
    for (int i = 0; i < 32; ++i) wr[i] = i < per ? __ldg(wp + i * 32 + lane) : 0.f;
    const float bs = bias[row];
    for (int b = 0; b < B; ++b) {
        const float* x = in + (int64_t)b * d;
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (i < per) a = fmaf(wr[i], __ldg(x + i * 32 + lane), a);
        a = warp_sum(a) + bs;
        if (SILU) a = a / (1.0f + expf(-a));
        if (lane == 0) out[(int64_t)b * rows + row] = a;
    }
}

}  // namespace

int temb_launch(const float* t, int B, int nf, const float* fourier_w, const float* w1, const float* b1,
                const float* w2, const float* b2, const float* dense_w, const float* dense_b, int rows, float* scratch,
                float* tb_out, cudaStream_t s) {
    SNRSE_CHECK_ARG(4 * nf <= 1024 && nf % 16 == 0, "temb: nf must be a multiple of 16, <= 256");
    // scratch: act(temb) [B][4nf] | emb [B][2nf] | h1 [B][4nf]   (caller provides B*4nf floats + this tail: see engine)
    const int d = 4 * nf;
    float* act = scratch;
    float* emb = scratch + (int64_t)B * d;
    float* h1 = emb + (int64_t)B * 2 * nf;
    snrse_launch(temb_fourier_kernel, dim3(B), dim3(nf), 0, s, t, nf, fourier_w, emb);
    SNRSE_LAUNCH_CHECK();
    // Linear(2nf -> 4nf) -> SiLU -> Linear(4nf -> 4nf) -> SiLU (the activation every Dense_0 applies to temb, ncsnpp.py:256-275)
    snrse_launch(dense_rows_kernel<1>, dim3(cdiv(d, 8)), dim3(256), 0, s, emb, B, 2 * nf, w1, b1, d, h1);
    SNRSE_LAUNCH_CHECK();
    snrse_launch(dense_rows_kernel<1>, dim3(cdiv(d, 8)), dim3(256), 0, s, h1, B, d, w2, b2, d, act);
    SNRSE_LAUNCH_CHECK();
    snrse_launch(dense_rows_kernel<0>, dim3(cdiv(rows, 8)), dim3(256), 0, s, act, B, d, dense_w, dense_b, rows, tb_out);
    SNRSE_LAUNCH_CHECK();
    return SNRSE_OK;
}

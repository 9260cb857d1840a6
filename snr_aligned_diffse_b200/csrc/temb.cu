// Time-embedding path of NCSN++ (sgmse-bbed/sgmse/backbones/ncsnpp.py:256-275) and the per-block
// `Dense_0(act(temb))` projections (ncsnpp_utils/layerspp.py:264-265), evaluated once per forward
// for all 49 residual blocks:  tb[b, row] = dense_w[row, :] . silu(temb[b, :]) + dense_b[row],
// where rows of all blocks are concatenated.  fp32 throughout (tiny GEMVs).
#include "kernels.h"

namespace {

// one block per batch item; blockDim = 4*nf (<= 512)
__global__ void temb_mlp_kernel(const float* __restrict__ t, int nf, const float* __restrict__ fw,
                                const float* __restrict__ w1, const float* __restrict__ b1,
                                const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ act_out) {
    extern __shared__ float sm[];  // emb[2nf] | h1[4nf]
    float* emb = sm;
    float* h1 = sm + 2 * nf;
    const int b = blockIdx.x, tid = threadIdx.x;
    const float lt = logf(t[b]);
    if (tid < nf) {
        // x_proj = log(t) * W * 2 * pi  (layerspp.py:42), evaluated in the reference's order
        const float proj = lt * fw[tid] * 2.0f * 3.14159265358979323846f;
        emb[tid] = sinf(proj);
        emb[nf + tid] = cosf(proj);
    }
    __syncthreads();
    const int d = 4 * nf;
    {
        float a = b1[tid];
        const float* wr = w1 + (int64_t)tid * 2 * nf;
        for (int k = 0; k < 2 * nf; ++k) a = fmaf(wr[k], emb[k], a);
        h1[tid] = a / (1.0f + expf(-a));  // act(temb) before the second Linear (ncsnpp.py:274)
    }
    __syncthreads();
    {
        float a = b2[tid];
        const float* wr = w2 + (int64_t)tid * d;
        for (int k = 0; k < d; ++k) a = fmaf(wr[k], h1[k], a);
        act_out[(int64_t)b * d + tid] = a / (1.0f + expf(-a));  // act(temb) fed to every Dense_0
    }
}

// one warp per output row, all batch items (B <= 32 per pass handled by looping)
__global__ void __launch_bounds__(256)
temb_dense_kernel(const float* __restrict__ act, int B, int d, const float* __restrict__ dw,
                  const float* __restrict__ db, int rows, float* __restrict__ out) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* wr = dw + (int64_t)row * d;
    for (int b = 0; b < B; ++b) {
        float a = 0.f;
        for (int k = lane; k < d; k += 32) a = fmaf(wr[k], act[(int64_t)b * d + k], a);
        a = warp_sum(a);
        if (lane == 0) out[(int64_t)b * rows + row] = a + db[row];
    }
}

}  // namespace

int temb_launch(const float* t, int B, int nf, const float* fourier_w, const float* w1, const float* b1,
                const float* w2, const float* b2, const float* dense_w, const float* dense_b, int rows, float* scratch,
                float* tb_out, cudaStream_t s) {
    SNRSE_CHECK_ARG(4 * nf <= 1024, "temb: nf too large");
    temb_mlp_kernel<<<B, 4 * nf, 6 * nf * sizeof(float), s>>>(t, nf, fourier_w, w1, b1, w2, b2, scratch);
    SNRSE_LAUNCH_CHECK();
    temb_dense_kernel<<<cdiv(rows, 8), 256, 0, s>>>(scratch, B, 4 * nf, dense_w, dense_b, rows, tb_out);
    SNRSE_LAUNCH_CHECK();
    return SNRSE_OK;
}

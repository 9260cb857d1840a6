// GroupNorm(32 groups, eps 1e-6) + SiLU for NHWC bf16 activations.
// Reference: nn.GroupNorm(num_groups=min(C//4,32), eps=1e-6) followed by nn.SiLU
// (sgmse-bbed/sgmse/backbones/ncsnpp_utils/layerspp.py:221,233,245,266; ncsnpp.py:210,348,362);
// biased variance over (C/32)*H*W elements per (sample, group).
//
// Three kernels, all deterministic (no floating-point atomics):
//   gn_stats    : per-(sample, pixel-chunk) partial sum / sum-of-squares per group   -> partial[B][chunks][32][2]
//   gn_finalize : combine partials in double, fold gamma/beta into per-(sample,channel) scale/shift
//   gn_apply    : y = silu(x*scale + shift)   (vectorised 8 channels / thread, HBM-bound)
#include "kernels.h"

namespace {

constexpr int GN_GROUPS = 32;
constexpr int GN_MAX_CHUNKS = 128;

// thread t handles channel vector (t % tpp) of pixels (t / tpp) + k*ppb
__global__ void __launch_bounds__(256)
gn_stats_kernel(const bf16* __restrict__ x, int ld, int C, int64_t hw, int pix_per_chunk, float* __restrict__ partial,
                int chunks) {
    const int tpp = C >> 3;
    const int ppb = blockDim.x / tpp;
    const int vec = threadIdx.x % tpp, prow = threadIdx.x / tpp;
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int64_t p0 = (int64_t)chunk * pix_per_chunk;
    int64_t p1 = p0 + pix_per_chunk;
    if (p1 > hw) p1 = hw;
    const bf16* base = x + (int64_t)b * hw * ld + vec * 8;
    float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
    for (int64_t p = p0 + prow; p < p1; p += 4 * ppb) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t q = p + (int64_t)u * ppb;
            v[u] = (q < p1) ? __ldg(reinterpret_cast<const uint4*>(base + q * ld)) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float f[8];
            unpack8(v[u], f);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                s0 += f[j];
                q0 = fmaf(f[j], f[j], q0);
                s1 += f[4 + j];
                q1 = fmaf(f[4 + j], f[4 + j], q1);
            }
        }
    }
    // deterministic block reduction: [prow][half-vector] -> per half-vector -> per group
    __shared__ float sh_s[256 * 2], sh_q[256 * 2];
    sh_s[threadIdx.x * 2] = s0;
    sh_s[threadIdx.x * 2 + 1] = s1;
    sh_q[threadIdx.x * 2] = q0;
    sh_q[threadIdx.x * 2 + 1] = q1;
    __syncthreads();
    const int halves = tpp * 2;  // 4-channel units per pixel
    __shared__ float hv_s[128], hv_q[128];
    if ((int)threadIdx.x < halves) {
        const int v = threadIdx.x >> 1, hf = threadIdx.x & 1;
        float a = 0.f, q = 0.f;
        for (int r = 0; r < ppb; ++r) {
            a += sh_s[(r * tpp + v) * 2 + hf];
            q += sh_q[(r * tpp + v) * 2 + hf];
        }
        hv_s[threadIdx.x] = a;
        hv_q[threadIdx.x] = q;
    }
    __syncthreads();
    if (threadIdx.x < GN_GROUPS) {
        const int units = (C / GN_GROUPS) >> 2;  // 4-channel units per group (1..4)
        float a = 0.f, q = 0.f;
        for (int u = 0; u < units; ++u) {
            a += hv_s[threadIdx.x * units + u];
            q += hv_q[threadIdx.x * units + u];
        }
        float* dst = partial + (((int64_t)b * chunks + chunk) * GN_GROUPS + threadIdx.x) * 2;
        dst[0] = a;
        dst[1] = q;
    }
}

__global__ void gn_finalize_kernel(const float* __restrict__ partial, int chunks, int C, double inv_count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float* __restrict__ scsh) {
    __shared__ float s_mean[GN_GROUPS], s_rstd[GN_GROUPS];
    const int b = blockIdx.x;
    if (threadIdx.x < GN_GROUPS) {
        double s = 0.0, q = 0.0;
        const float* src = partial + ((int64_t)b * chunks * GN_GROUPS + threadIdx.x) * 2;
        for (int c = 0; c < chunks; ++c) {
            s += (double)src[(int64_t)c * GN_GROUPS * 2];
            q += (double)src[(int64_t)c * GN_GROUPS * 2 + 1];
        }
        const double mean = s * inv_count;
        double var = q * inv_count - mean * mean;
        if (var < 0.0) var = 0.0;
        s_mean[threadIdx.x] = (float)mean;
        s_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
    }
    __syncthreads();
    const int cpg = C / GN_GROUPS;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int g = c / cpg;
        const float sc = s_rstd[g] * gamma[c];
        scsh[((int64_t)b * 2) * C + c] = sc;
        scsh[((int64_t)b * 2 + 1) * C + c] = beta[c] - s_mean[g] * sc;
    }
}

__global__ void __launch_bounds__(256)
gn_apply_kernel(const bf16* __restrict__ x, int ld, int C, int64_t hw, const float* __restrict__ scsh, int silu,
                bf16* __restrict__ out, int out_ld, int pix_per_block) {
    const int tpp = C >> 3;
    const int ppb = blockDim.x / tpp;
    const int vec = threadIdx.x % tpp, prow = threadIdx.x / tpp;
    const int b = blockIdx.y;
    const int64_t p0 = (int64_t)blockIdx.x * pix_per_block;
    int64_t p1 = p0 + pix_per_block;
    if (p1 > hw) p1 = hw;
    float sc[8], sh[8];
    {
        const float4* a = reinterpret_cast<const float4*>(scsh + ((int64_t)b * 2) * C + vec * 8);
        const float4* c = reinterpret_cast<const float4*>(scsh + ((int64_t)b * 2 + 1) * C + vec * 8);
        const float4 a0 = a[0], a1 = a[1], c0 = c[0], c1 = c[1];
        sc[0] = a0.x; sc[1] = a0.y; sc[2] = a0.z; sc[3] = a0.w; sc[4] = a1.x; sc[5] = a1.y; sc[6] = a1.z; sc[7] = a1.w;
        sh[0] = c0.x; sh[1] = c0.y; sh[2] = c0.z; sh[3] = c0.w; sh[4] = c1.x; sh[5] = c1.y; sh[6] = c1.z; sh[7] = c1.w;
    }
    const bf16* src = x + (int64_t)b * hw * ld + vec * 8;
    bf16* dst = out + (int64_t)b * hw * out_ld + vec * 8;
    // 4 pixels per thread per trip: all loads issued before the first use (memory-level parallelism)
    for (int64_t p = p0 + prow; p < p1; p += 4 * ppb) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t q = p + (int64_t)u * ppb;
            if (q < p1) v[u] = __ldg(reinterpret_cast<const uint4*>(src + q * ld));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t q = p + (int64_t)u * ppb;
            if (q < p1) {
                float f[8];
                unpack8(v[u], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float t = fmaf(f[j], sc[j], sh[j]);
                    f[j] = silu ? silu_f(t) : t;
                }
                *reinterpret_cast<uint4*>(dst + q * out_ld) = pack8(f);
            }
        }
    }
}

int gn_threads(int C) {
    const int tpp = C / 8;
    return tpp * (256 / tpp);
}

}  // namespace

int gn_max_chunks() { return GN_MAX_CHUNKS; }

int gn_stats_launch(const ActView* x, float* partial, int chunks, cudaStream_t s) {
    SNRSE_CHECK_ARG(x->C % 128 == 0 && x->C <= 512, "GroupNorm: C must be 128/256/384/512 (got %d)", x->C);
    SNRSE_CHECK_ARG(chunks >= 1 && chunks <= GN_MAX_CHUNKS, "GroupNorm: bad chunk count %d", chunks);
    const int64_t hw = (int64_t)x->H * x->W;
    const int ppc = (int)cdiv64(hw, chunks);
    dim3 grid(chunks, x->B);
    gn_stats_kernel<<<grid, gn_threads(x->C), 0, s>>>(x->ptr, x->ld, x->C, hw, ppc, partial, chunks);
    SNRSE_LAUNCH_CHECK();
    return SNRSE_OK;
}

int gn_finalize_launch(const float* partial, int chunks, int B, int C, int64_t count_per_group, const float* gamma,
                       const float* beta, float eps, float* scsh, cudaStream_t s) {
    gn_finalize_kernel<<<B, 256, 0, s>>>(partial, chunks, C, 1.0 / (double)count_per_group, gamma, beta, eps, scsh);
    SNRSE_LAUNCH_CHECK();
    return SNRSE_OK;
}

int gn_apply_launch(const ActView* x, const float* scsh, int silu, const ActView* out, cudaStream_t s) {
    const int64_t hw = (int64_t)x->H * x->W;
    const int nthr = gn_threads(x->C);
    const int ppb = nthr / (x->C / 8);
    // ~16 pixels per thread-row (4 trips of 4), at least one block
    int64_t blocks_per_img = cdiv64(hw, (int64_t)ppb * 16);
    if (blocks_per_img < 1) blocks_per_img = 1;
    const int pix_per_block = (int)cdiv64(hw, blocks_per_img);
    dim3 grid((unsigned)blocks_per_img, x->B);
    gn_apply_kernel<<<grid, nthr, 0, s>>>(x->ptr, x->ld, x->C, hw, scsh, silu, out->ptr, out->ld, pix_per_block);
    SNRSE_LAUNCH_CHECK();
    return SNRSE_OK;
}

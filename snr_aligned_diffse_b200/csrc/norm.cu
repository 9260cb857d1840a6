// GroupNorm(32 groups, eps 1e-6) + SiLU for NHWC bf16 activations.
// Reference: nn.GroupNorm(num_groups=min(C//4,32), eps=1e-6) followed by nn.SiLU
// (sgmse-bbed/sgmse/backbones/ncsnpp_utils/layerspp.py:221,233,245,266; ncsnpp.py:210,348,362);
// biased variance over (C/32)*H*W elements per (sample, group).
//
// Statistics are kept per 4-channel UNIT (the group size at C=128; 2 / 3 / 4 units per group at C = 256 / 384 / 512)
// as [B][U][2] (sum, sum of squares; U = C/4) 64-bit FIXED-POINT integers (gn_fixed.cuh) accumulated with integer
// atomics: a tensor's statistics can be produced by whoever writes it (the convolution epilogue, conv_halo2.cu) and
// re-grouped later, e.g. for the GroupNorm over torch.cat([h, skip]) (ncsnpp.py:337) from the two producers' units.
// Integer addition commutes, so the result is bit-identical from run to run and independent of how pixels are
// distributed over CTAs (no floating-point atomics anywhere).
//   gn_stats    : stand-alone pass over a tensor (fp32 partial per block of <= 1024 pixels -> integer atomics)
//   gn_finalize : regroup the units of one or two sources, fold gamma/beta into per-(sample, channel) scale/shift
//   gn_apply    : y = silu(x*scale + shift)   (vectorised 8 channels / thread, HBM-bound)
#include "gn_fixed.cuh"
#include "kernels.h"

namespace {

constexpr int GN_GROUPS = 32;
constexpr int GN_MAX_CHUNKS = 128;

// thread t handles channel vector (t % tpp) of pixels (t / tpp) + k*ppb
__global__ void __launch_bounds__(256)
gn_stats_kernel(const bf16* __restrict__ x, int ld, int C, int64_t hw, int pix_per_chunk,
                unsigned long long* __restrict__ ustats) {
    pdl_sync();
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
    if ((int)threadIdx.x < halves) {   // one (sum, sumsq) pair per 4-channel unit, accumulated in fixed point
        unsigned long long* dst = ustats + ((int64_t)b * halves + threadIdx.x) * 2;
        atomicAdd(dst, gn_fix_sum(hv_s[threadIdx.x]));
        atomicAdd(dst + 1, gn_fix_sq(hv_q[threadIdx.x]));
    }
}

// block = one sample.  src: [B][U][2] fixed-point sums; two sources = GroupNorm over the channel concatenation.
__global__ void __launch_bounds__(256)
gn_finalize_kernel(const unsigned long long* __restrict__ src0, int U0, const unsigned long long* __restrict__ src1, int U1,
                   double inv_count, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                   float* __restrict__ scsh) {
    pdl_sync();
    __shared__ float s_mean[GN_GROUPS], s_rstd[GN_GROUPS];
    const int b = blockIdx.x;
    const int U = U0 + U1, C = 4 * U;
    if (threadIdx.x < GN_GROUPS) {
        const int upg = U / GN_GROUPS;   // units per group
        long long sm = 0;                // exact integer sum of the fixed-point unit sums
        double q = 0.0;                  // unit square sums added in a fixed order (<= 4 terms): deterministic, no wrap
        for (int k = 0; k < upg; ++k) {
            const int u = threadIdx.x * upg + k;
            const unsigned long long* p = u < U0 ? src0 + ((int64_t)b * U0 + u) * 2 : src1 + ((int64_t)b * U1 + (u - U0)) * 2;
            sm += (long long)p[0];
            q += gn_unfix_sq(p[1]);
        }
        const double mean = gn_unfix_sum(sm) * inv_count;
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
    pdl_sync();
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


int gn_stats_launch(const ActView* x, unsigned long long* ustats, cudaStream_t s) {
    SNRSE_CHECK_ARG(x->C % 128 == 0 && x->C <= 512, "GroupNorm: C must be 128/256/384/512 (got %d)", x->C);
    const int64_t hw = (int64_t)x->H * x->W;
    int chunks = (int)(hw / 256 < 1 ? 1 : (hw / 256 > GN_MAX_CHUNKS ? GN_MAX_CHUNKS : hw / 256));
    // the chunking depends on the image size only: a sample's statistics do not depend on its batch
    const int ppc = (int)cdiv64(hw, chunks);
    dim3 grid(chunks, x->B);
    snrse_launch(gn_stats_kernel, dim3(grid), dim3(gn_threads(x->C)), 0, s, x->ptr, x->ld, x->C, hw, ppc, ustats);
    SNRSE_LAUNCH_CHECK();
    return SNRSE_OK;
}

int gn_finalize_launch(const unsigned long long* src0, int U0, const unsigned long long* src1, int U1, int B,
                       int64_t count_per_group, const float* gamma, const float* beta, float eps, float* scsh,
                       cudaStream_t s) {
    const int C = 4 * (U0 + U1);
    SNRSE_CHECK_ARG(C % 128 == 0 && C <= 512 && src0 && (U1 == 0 || src1), "GroupNorm finalize: bad sources");
    snrse_launch_m(1 | 4, gn_finalize_kernel, dim3(B), dim3(256), 0, s, src0, U0, src1, U1, 1.0 / (double)count_per_group, gamma, beta, eps, scsh);
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
    snrse_launch(gn_apply_kernel, dim3(grid), dim3(nthr), 0, s, x->ptr, x->ld, x->C, hw, scsh, silu, out->ptr, out->ld, pix_per_block);
    SNRSE_LAUNCH_CHECK();
    return SNRSE_OK;
}

// Single-head self-attention core of `AttnBlockpp` (sgmse-bbed/sgmse/backbones/ncsnpp_utils/layerspp.py:84-88):
//   w = softmax_j( q_i . k_j / sqrt(C) ),  o_i = sum_j w_ij v_j     over all n = H*W positions of one image.
// q, k, v come from the NIN projections (tcgen05 GEMMs).  Two paths:
//  * tensor cores (n % 64 == 0): S = scale * Q K^T and O = P V are batched GEMMs on conv_gemm.cu (operand B = K
//    resp. V^T of the same image), fp32 scores, fp32 softmax, probabilities rounded to bf16 for the second GEMM.
//    3076*T^2 FLOP per NFE: 0.1 % of the network at 4 s, 1.1 % at 60 s.
//  * CUDA cores in fp32 (any n; the 4 x T/64 bottleneck of short inputs, and the cross-check):
//      scores  S = scale * Q K^T        (64x64 tiles, fp32 accumulate)  -> f32 [B, n, n] workspace
//      softmax rows in place            (one warp per row)
//      mix     O = S V                  (64x64 tiles)                    -> bf16 view
#include "kernels.h"

namespace {

constexpr int TS = 64, TK = 32;

__global__ void __launch_bounds__(256)
attn_scores_kernel(const bf16* __restrict__ q, int q_ld, const bf16* __restrict__ k, int k_ld, int n, int C, float scale,
                   float* __restrict__ S) {
    pdl_sync();
    __shared__ float qs[TS][TK + 1], ks[TS][TK + 1];
    const int b = blockIdx.z, i0 = blockIdx.y * TS, j0 = blockIdx.x * TS;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const bf16* qb = q + (int64_t)b * n * q_ld;
    const bf16* kb = k + (int64_t)b * n * k_ld;
    float acc[4][4] = {};
    for (int c0 = 0; c0 < C; c0 += TK) {
        for (int e = threadIdx.x; e < TS * TK; e += 256) {
            const int r = e / TK, c = e % TK;
            qs[r][c] = (i0 + r < n) ? __bfloat162float(qb[(int64_t)(i0 + r) * q_ld + c0 + c]) : 0.f;
            ks[r][c] = (j0 + r < n) ? __bfloat162float(kb[(int64_t)(j0 + r) * k_ld + c0 + c]) : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int c = 0; c < TK; ++c) {
            float a[4], bb[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                a[u] = qs[ty * 4 + u][c];
                bb[u] = ks[tx * 4 + u][c];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(a[u], bb[v], acc[u][v]);
        }
        __syncthreads();
    }
    float* Sb = S + (int64_t)b * n * n;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int i = i0 + ty * 4 + u;
        if (i >= n) continue;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const int j = j0 + tx * 4 + v;
            if (j < n) Sb[(int64_t)i * n + j] = acc[u][v] * scale;
        }
    }
}

__global__ void __launch_bounds__(256)
attn_softmax_kernel(float* __restrict__ S, int n, int64_t rows) {
    pdl_sync();
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float* p = S + row * n;
    float m = -INFINITY;
    for (int j = lane; j < n; j += 32) m = fmaxf(m, p[j]);
    m = warp_max(m);
    float sum = 0.f;
    for (int j = lane; j < n; j += 32) {
        const float e = __expf(p[j] - m);
        p[j] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int j = lane; j < n; j += 32) p[j] *= inv;
}

__global__ void __launch_bounds__(256)
attn_mix_kernel(const float* __restrict__ S, const bf16* __restrict__ v, int v_ld, int n, int C, bf16* __restrict__ o,
                int o_ld) {
    pdl_sync();
    __shared__ float ps[TS][TK + 1], vs[TK][TS + 1];
    const int b = blockIdx.z, i0 = blockIdx.y * TS, c0 = blockIdx.x * TS;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const float* Sb = S + (int64_t)b * n * n;
    const bf16* vb = v + (int64_t)b * n * v_ld;
    float acc[4][4] = {};
    for (int j0 = 0; j0 < n; j0 += TK) {
        for (int e = threadIdx.x; e < TS * TK; e += 256) {
            const int r = e / TK, c = e % TK;
            ps[r][c] = (i0 + r < n && j0 + c < n) ? Sb[(int64_t)(i0 + r) * n + j0 + c] : 0.f;
            const int jr = e / TS, cc = e % TS;
            vs[jr][cc] = (j0 + jr < n && c0 + cc < C) ? __bfloat162float(vb[(int64_t)(j0 + jr) * v_ld + c0 + cc]) : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int j = 0; j < TK; ++j) {
            float a[4], bb[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                a[u] = ps[ty * 4 + u][j];
                bb[u] = vs[j][tx * 4 + u];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int w = 0; w < 4; ++w) acc[u][w] = fmaf(a[u], bb[w], acc[u][w]);
        }
        __syncthreads();
    }
    bf16* ob = o + (int64_t)b * n * o_ld;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int i = i0 + ty * 4 + u;
        if (i >= n) continue;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const int c = c0 + tx * 4 + w;
            if (c < C) ob[(int64_t)i * o_ld + c] = __float2bfloat16(acc[u][w]);
        }
    }
}

// row softmax of fp32 scores -> bf16 probabilities (one warp per row)
__global__ void __launch_bounds__(256)
attn_softmax_bf16_kernel(const float* __restrict__ S, bf16* __restrict__ P, int n, int64_t rows) {
    pdl_sync();
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* p = S + row * n;
    bf16* o = P + row * n;
    float m = -INFINITY;
    for (int j = lane; j < n; j += 32) m = fmaxf(m, p[j]);
    m = warp_max(m);
    float sum = 0.f;
    for (int j = lane; j < n; j += 32) sum += __expf(p[j] - m);
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int j = lane; j < n; j += 32) o[j] = __float2bfloat16(__expf(p[j] - m) * inv);
}

// V [B][n][ld] (C channels) -> V^T [B][C][n], 32x32 tiles through shared memory
__global__ void __launch_bounds__(256)
attn_transpose_kernel(const bf16* __restrict__ v, int v_ld, int n, int C, bf16* __restrict__ vt) {
    pdl_sync();
    __shared__ bf16 t[32][33];
    const int b = blockIdx.z, j0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8)
        if (j0 + r < n && c0 + tx < C) t[r][tx] = v[((int64_t)b * n + j0 + r) * v_ld + c0 + tx];
    __syncthreads();
    for (int r = ty; r < 32; r += 8)
        if (c0 + r < C && j0 + tx < n) vt[((int64_t)b * C + c0 + r) * n + j0 + tx] = t[tx][r];
}

}  // namespace

// The reference materialises the full [B, n, n] score matrix (layerspp.py:84-88).  Here at most ATTN_SCORE_BUDGET bytes
// of scores (fp32) + probabilities (bf16) are live: when B*n*n*6 exceeds it, the queries are walked in blocks of `qc`
// rows per utterance (every block is a complete softmax over all n keys, so the result is unchanged); a 60 s utterance
// (n = 7552) then needs 190 MB instead of 342 MB per attention block, a 64 x 60 s batch 190 MB instead of 22 GB.
constexpr int64_t ATTN_SCORE_BUDGET = 192ll << 20;

static void attn_blocking(int B, int n, int* pass_b, int* qc) {
    if ((int64_t)B * n * n * 6 <= ATTN_SCORE_BUDGET || n % 64 != 0) {
        *pass_b = B;
        *qc = n;
        return;
    }
    int64_t rows = ATTN_SCORE_BUDGET / (6ll * n);
    rows = rows / 64 * 64;
    if (rows < 64) rows = 64;
    if (rows > n) rows = n;
    *pass_b = 1;
    *qc = (int)rows;
}

int64_t attention_workspace_bytes(int B, int n, int C) {
    int pb, qc;
    attn_blocking(B, n, &pb, &qc);
    const int64_t nn = (int64_t)pb * qc * n;
    return nn * 4 + ((nn * 2 + 255) / 256) * 256 + (int64_t)B * C * n * 2 + 512;
}

int attention_launch(const ActView* q, const ActView* k, const ActView* v, void* workspace, const ActView* o,
                     cudaStream_t s, int allow_tensor_cores) {
    const int n = q->H * q->W, C = q->C, B = q->B;
    SNRSE_CHECK_ARG(C % TK == 0, "attention: C must be a multiple of %d", TK);
    float* scores = static_cast<float*>(workspace);
    if (allow_tensor_cores && n % 64 == 0 && C % 64 == 0 && q->ld % 8 == 0 && k->ld % 8 == 0) {
        int pb, qc;
        attn_blocking(B, n, &pb, &qc);
        const int64_t nn = (int64_t)pb * qc * n;
        bf16* probs = reinterpret_cast<bf16*>(static_cast<uint8_t*>(workspace) + nn * 4);
        bf16* vt = reinterpret_cast<bf16*>(reinterpret_cast<uint8_t*>(probs) + ((nn * 2 + 255) / 256) * 256);
        dim3 gt(cdiv(C, 32), cdiv(n, 32), B);
        snrse_launch(attn_transpose_kernel, dim3(gt), dim3(256), 0, s, v->ptr, v->ld, n, C, vt);
        SNRSE_LAUNCH_CHECK();
        for (int b0 = 0; b0 < B; b0 += pb) {
            for (int q0 = 0; q0 < n; q0 += qc) {
                const int rows_q = (n - q0 < qc) ? n - q0 : qc;
                // S = scale * Q K^T : A = Q block (tokens x C), B = K (keys x C, row pitch k->ld)
                ActView qa = *q;
                qa.ptr = q->ptr + ((int64_t)b0 * n + q0) * q->ld;
                qa.B = pb; qa.H = 1; qa.W = rows_q;
                ConvGemmPlan g1;
                SNRSE_TRY(conv_gemm_make_plan_ex(&g1, &qa, 1, nullptr, k->ptr + (int64_t)b0 * n * k->ld, n, (int64_t)n * k->ld, 1,
                                                 nullptr, nullptr, 0, nullptr, rsqrtf((float)C), scores, n, 1, k->ld));
                SNRSE_TRY(conv_gemm_launch(&g1, s));
                const int64_t rows = (int64_t)pb * rows_q;
                snrse_launch(attn_softmax_bf16_kernel, dim3((unsigned)cdiv64(rows, 8)), dim3(256), 0, s, scores, probs, n, rows);
                SNRSE_LAUNCH_CHECK();
                // O = P V : A = P (tokens x keys), B = V^T (C x keys)
                ActView pa;
                pa.ptr = probs; pa.B = pb; pa.H = 1; pa.W = rows_q; pa.C = n; pa.ld = n;
                ConvGemmPlan g2;
                SNRSE_TRY(conv_gemm_make_plan_ex(&g2, &pa, 1, nullptr, vt + (int64_t)b0 * C * n, C, (int64_t)C * n, 1, nullptr,
                                                 nullptr, 0, nullptr, 1.0f, o->ptr + ((int64_t)b0 * n + q0) * o->ld, o->ld, 0, n));
                SNRSE_TRY(conv_gemm_launch(&g2, s));
            }
        }
        return SNRSE_OK;
    }
    dim3 g1(cdiv(n, TS), cdiv(n, TS), B);
    snrse_launch(attn_scores_kernel, dim3(g1), dim3(256), 0, s, q->ptr, q->ld, k->ptr, k->ld, n, C, rsqrtf((float)C), scores);
    SNRSE_LAUNCH_CHECK();
    const int64_t rows = (int64_t)B * n;
    snrse_launch(attn_softmax_kernel, dim3((unsigned)cdiv64(rows, 8)), dim3(256), 0, s, scores, n, rows);
    SNRSE_LAUNCH_CHECK();
    dim3 g3(cdiv(C, TS), cdiv(n, TS), B);
    snrse_launch(attn_mix_kernel, dim3(g3), dim3(256), 0, s, scores, v->ptr, v->ld, n, C, o->ptr, o->ld);
    SNRSE_LAUNCH_CHECK();
    return SNRSE_OK;
}

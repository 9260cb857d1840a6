// Persistent halo-reuse implicit-GEMM 3x3 convolution for sm_100a (tcgen05 + TMEM + TMA).
//
// Same contract as conv_gemm.cu (ddpm_conv3x3 of sgmse-bbed/sgmse/backbones/ncsnpp_utils/layers.py:118-124,
// with the fused 1x1 shortcut / bias / time-embedding bias / residual epilogue of layerspp.py:262-276), but
// organised around L2->SM traffic, which is what bounded the first kernel (profiles/r01_v1_conv_gemm_ncu.md:
// 30 KB fetched per MMAC, tensor pipe 34 % active):
//
//   * Super-tile = (8*SUB) x 16 output pixels of one image, SUB in {1,2}: SUB accumulators of 128 rows x N
//     columns live in TMEM and share every weight tile  -> weight traffic per pixel / SUB.
//   * Input patch: for each 64-channel chunk only THREE TMA boxes are fetched, the column-shifted halo copies
//     {64 ch, 16, 8*SUB+2 rows} at w0-1, w0, w0+1.  The three row taps of a copy are the SAME shared-memory
//     buffer addressed (r + 8u) * 16 rows further down: the UMMA descriptor start address moves in steps of
//     2 KB, a multiple of the 1 KB swizzle atom, so no data is duplicated  -> 9 -> 3*(8*SUB+2)/(8*SUB) loads.
//   * Two independent TMA rings (A: halo copies, B: per-tap weight tiles) fed by two producer warps.
//   * Persistent CTAs (one per SM) walk super-tiles round-robin; with SUB*N <= 256 the TMEM accumulator is
//     double-buffered so the epilogue of tile i overlaps the main loop of tile i+1.
//   Per MMAC this fetches ~13 KB (N=128) / ~11 KB (N=256) instead of 30 KB.
//
// Warp roles (8 warps): 0 = A producer, 1 = MMA issuer + TMEM owner, 2 = B producer, 3 = idle,
// 4..7 = epilogue (TMEM lane quarter = warp & 3).
#include <cuda.h>

#include "kernels.h"
#include "ptx.cuh"
#include "tma_host.h"

namespace {

constexpr int HALO_THREADS = 256;
constexpr int MAX_A = 4, MAX_B = 8;
constexpr int TW = 16, SUB_ROWS = 8;

struct HaloArgs {
    int c0_chunks, c1_chunks;
    int H, W, B;
    int sub;                       // sub-tiles per super-tile (1 or 2)
    int tiles_h, tiles_w, n_tiles;
    int N;                         // output channels == columns per accumulator (128 or 256)
    int na, nb;                    // ring depths
    int acc_bufs;                  // 1 or 2 TMEM accumulator sets
    const float* bias;
    const float* tbias;
    int tb_stride;
    const bf16* res;
    int res_ld;
    float scale;
    bf16* out;
    int out_ld;
    long long* dbg;                // optional per-CTA cycle counters [grid][8] (measurement builds only)
};

#define DBG_T0() long long t0__ = g.dbg ? clock64() : 0
#define DBG_ADD(acc) do { if (g.dbg) { const long long t1__ = clock64(); acc += t1__ - t0__; t0__ = t1__; } } while (0)

__global__ void __launch_bounds__(HALO_THREADS, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                 const __grid_constant__ CUtensorMap mapB, const HaloArgs g) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t a_full[MAX_A], a_empty[MAX_A], b_full[MAX_B], b_empty[MAX_B];
    __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_bytes = (uint32_t)(SUB_ROWS * g.sub + 2) * TW * 128;   // 20 KB / 36 KB, multiple of 1 KB
    const uint32_t b_bytes = (uint32_t)g.N * 128;
    const uint32_t a_base = smem_base, b_base = smem_base + (uint32_t)g.na * a_bytes;
    const int n_astage = 3 * g.c0_chunks + g.c1_chunks;   // halo copies consumed per tile
    const uint32_t acc_cols = (uint32_t)(g.sub * g.N);
    const uint32_t tmem_cols = acc_cols * (uint32_t)g.acc_bufs;  // 128/256/512: power of two

    if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < g.na; ++i) {
                ptx::mbar_init(ptx::smem_u32(&a_full[i]), 1);
                ptx::mbar_init(ptx::smem_u32(&a_empty[i]), 1);
            }
            for (int i = 0; i < g.nb; ++i) {
                ptx::mbar_init(ptx::smem_u32(&b_full[i]), 1);
                ptx::mbar_init(ptx::smem_u32(&b_empty[i]), 1);
            }
            for (int i = 0; i < 2; ++i) {
                ptx::mbar_init(ptx::smem_u32(&acc_full[i]), 1);
                ptx::mbar_init(ptx::smem_u32(&acc_empty[i]), 4);   // one arrive per epilogue warp
            }
            ptx::fence_barrier_init();
        }
        __syncwarp();
        ptx::tmem_alloc(ptx::smem_u32(&tmem_base_smem), tmem_cols);
        ptx::tmem_relinquish();
    } else if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&mapA0);
        if (g.c1_chunks > 0) ptx::prefetch_tensormap(&mapA1);
    } else if (warp == 2 && lane == 0) {
        ptx::prefetch_tensormap(&mapB);
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    const int tiles_per_img = g.tiles_h * g.tiles_w;

    if (warp == 0) {
        // =========================== A producer: halo copies ===========================
        if (lane == 0) {
            uint32_t it = 0;   // running stage counter across tiles
            long long w_a = 0;
            DBG_T0();
            for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
                const int b = tile / tiles_per_img, rem = tile % tiles_per_img;
                const int h0 = (rem / g.tiles_w) * SUB_ROWS * g.sub, w0 = (rem % g.tiles_w) * TW;
                for (int j = 0; j < n_astage; ++j, ++it) {
                    const uint32_t s = it % (uint32_t)g.na, ph = (it / (uint32_t)g.na) & 1u;
                    if (g.dbg) t0__ = clock64();
                    ptx::mbar_wait(ptx::smem_u32(&a_empty[s]), ph ^ 1u);
                    DBG_ADD(w_a);
                    const uint32_t fb = ptx::smem_u32(&a_full[s]);
                    ptx::mbar_arrive_expect_tx(fb, a_bytes);
                    const uint32_t dst = a_base + s * a_bytes;
                    if (j < 3 * g.c0_chunks) {
                        const int c = j / 3, sh = j % 3;      // chunk, column shift index (dw = sh - 1)
                        ptx::tma_load_4d(dst, &mapA0, fb, c * 64, w0 + sh - 1, h0 - 1, b);
                    } else {
                        ptx::tma_load_4d(dst, &mapA1, fb, (j - 3 * g.c0_chunks) * 64, w0, h0, b);
                    }
                }
            }
            if (g.dbg) g.dbg[blockIdx.x * 8 + 6] = w_a;
        }
    } else if (warp == 2) {
        // =========================== B producer: weight tiles ===========================
        if (lane == 0) {
            uint32_t it = 0;
            long long w_b = 0;
            DBG_T0();
            for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
                for (int j = 0; j < n_astage; ++j) {
                    const bool seg0 = j < 3 * g.c0_chunks;
                    const int c = j / 3, sh = j % 3;
                    const int ntap = seg0 ? 3 : 1;
                    for (int r = 0; r < ntap; ++r, ++it) {
                        const uint32_t s = it % (uint32_t)g.nb, ph = (it / (uint32_t)g.nb) & 1u;
                        if (g.dbg) t0__ = clock64();
                        ptx::mbar_wait(ptx::smem_u32(&b_empty[s]), ph ^ 1u);
                        DBG_ADD(w_b);
                        const uint32_t fb = ptx::smem_u32(&b_full[s]);
                        ptx::mbar_arrive_expect_tx(fb, b_bytes);
                        // K layout of the packed weights: [tap = r*3 + sh][cin], then the shortcut channels
                        const int kb = seg0 ? ((r * 3 + sh) * g.c0_chunks + c) : (9 * g.c0_chunks + (j - 3 * g.c0_chunks));
                        ptx::tma_load_3d(b_base + s * b_bytes, &mapB, fb, kb * 64, 0, 0);
                    }
                }
            }
            if (g.dbg) g.dbg[blockIdx.x * 8 + 7] = w_b;
        }
    } else if (warp == 1) {
        // =========================== MMA issuer ===========================
        if (lane == 0) {
            const uint32_t idesc = ptx::umma_idesc_bf16(128, (uint32_t)g.N);
            uint32_t ita = 0, itb = 0, itt = 0;
            long long w_a = 0, w_b = 0, w_acc = 0;
            const long long t_start = g.dbg ? clock64() : 0;
            DBG_T0();
            for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++itt) {
                const uint32_t buf = g.acc_bufs == 2 ? (itt & 1u) : 0u;
                const uint32_t use = g.acc_bufs == 2 ? (itt >> 1) : itt;   // how often this buffer was used before
                if (g.dbg) t0__ = clock64();
                ptx::mbar_wait(ptx::smem_u32(&acc_empty[buf]), (use & 1u) ^ 1u);
                DBG_ADD(w_acc);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * acc_cols;
                for (int j = 0; j < n_astage; ++j, ++ita) {
                    const uint32_t sa = ita % (uint32_t)g.na, pha = (ita / (uint32_t)g.na) & 1u;
                    if (g.dbg) t0__ = clock64();
                    ptx::mbar_wait(ptx::smem_u32(&a_full[sa]), pha);
                    DBG_ADD(w_a);
                    ptx::tc_fence_after();
                    const bool seg0 = j < 3 * g.c0_chunks;
                    const int ntap = seg0 ? 3 : 1;
                    for (int r = 0; r < ntap; ++r, ++itb) {
                        const uint32_t sb = itb % (uint32_t)g.nb, phb = (itb / (uint32_t)g.nb) & 1u;
                        if (g.dbg) t0__ = clock64();
                        ptx::mbar_wait(ptx::smem_u32(&b_full[sb]), phb);
                        DBG_ADD(w_b);
                        ptx::tc_fence_after();
                        const uint64_t db = ptx::umma_desc_k_sw128(b_base + sb * b_bytes);
                        for (int u = 0; u < g.sub; ++u) {
                            // rows (r + 8u) .. of the halo copy: 16 pixels x 128 B per row = 2 KB steps
                            const uint64_t da = ptx::umma_desc_k_sw128(a_base + sa * a_bytes + (uint32_t)(r + SUB_ROWS * u) * (TW * 128));
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                ptx::mma_bf16_ss(d_tmem + (uint32_t)(u * g.N), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                                                 (j > 0 || r > 0 || k > 0) ? 1u : 0u);
                        }
                        ptx::mma_commit(ptx::smem_u32(&b_empty[sb]));
                    }
                    ptx::mma_commit(ptx::smem_u32(&a_empty[sa]));
                }
                ptx::mma_commit(ptx::smem_u32(&acc_full[buf]));
            }
            if (g.dbg) {
                g.dbg[blockIdx.x * 8 + 0] = w_a;
                g.dbg[blockIdx.x * 8 + 1] = w_b;
                g.dbg[blockIdx.x * 8 + 2] = w_acc;
                g.dbg[blockIdx.x * 8 + 3] = clock64() - t_start;
            }
        }
    } else if (warp >= 4) {
        // =========================== epilogue ===========================
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;          // TMEM lane == row of the 128-pixel sub-tile
        const int hl = row >> 4, wl = row & 15;
        uint32_t itt = 0;
        long long w_full = 0, t_body = 0;
        DBG_T0();
        for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++itt) {
            const int b = tile / tiles_per_img, rem = tile % tiles_per_img;
            const int h0 = (rem / g.tiles_w) * SUB_ROWS * g.sub, w0 = (rem % g.tiles_w) * TW;
            const uint32_t buf = g.acc_bufs == 2 ? (itt & 1u) : 0u;
            const uint32_t use = g.acc_bufs == 2 ? (itt >> 1) : itt;
            if (g.dbg) t0__ = clock64();
            ptx::mbar_wait(ptx::smem_u32(&acc_full[buf]), use & 1u);
            DBG_ADD(w_full);
            ptx::tc_fence_after();
            const float* tb = g.tbias ? g.tbias + (int64_t)b * g.tb_stride : nullptr;
            for (int u = 0; u < g.sub; ++u) {
                const int h = h0 + SUB_ROWS * u + hl, w = w0 + wl;
                const bool valid = (h < g.H) && (w < g.W);
                const int64_t pix = ((int64_t)b * g.H + h) * g.W + w;
                for (int c0 = 0; c0 < g.N; c0 += 32) {
                    uint32_t v[32];
                    ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * acc_cols + (uint32_t)(u * g.N + c0), v);
                    ptx::tmem_ld_wait();
                    if (valid) {
                        float f[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                        if (g.bias) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) f[j] += __ldg(g.bias + c0 + j);
                        }
                        if (tb) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) f[j] += __ldg(tb + c0 + j);
                        }
                        if (g.res) {
                            const uint4* rp = reinterpret_cast<const uint4*>(g.res + pix * g.res_ld + c0);
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                float rr[8];
                                unpack8(__ldg(rp + q), rr);
#pragma unroll
                                for (int j = 0; j < 8; ++j) f[q * 8 + j] += rr[j];
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] *= g.scale;
                        uint4* op = reinterpret_cast<uint4*>(g.out + pix * g.out_ld + c0);
#pragma unroll
                        for (int q = 0; q < 4; ++q) op[q] = pack8(f + 8 * q);
                    }
                }
            }
            // all TMEM reads of this accumulator set are complete: hand it back to the MMA warp
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&acc_empty[buf]));
            DBG_ADD(t_body);
        }
        if (g.dbg && threadIdx.x == 128) {
            g.dbg[blockIdx.x * 8 + 4] = w_full;
            g.dbg[blockIdx.x * 8 + 5] = t_body;
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        ptx::tmem_dealloc(tmem_base, tmem_cols);
    }
}

int g_num_sms = 0;
long long* g_halo_dbg = nullptr;
}  // namespace
long long* g_halo_dbg_shared = nullptr;
namespace {

}  // namespace

// measurement hook: device buffer of [grid][8] cycle counters filled by subsequent launches (null: off)
extern "C" void snrse_conv_halo_set_debug(long long* dev_counters) {
    g_halo_dbg = dev_counters;
    g_halo_dbg_shared = dev_counters;
}

bool conv_halo_eligible(const ActView* a0, int taps0, int n_rows) {
    return taps0 == 9 && a0->W >= 16 && a0->H >= 8 && (n_rows == 128 || n_rows == 256);
}

int conv_halo_make_plan(ConvHaloPlan* p, const ActView* a0, const ActView* a1, const bf16* wt, int n_rows,
                        const float* bias, const float* tbias, int tb_stride, const ActView* res, float scale, bf16* out,
                        int out_ld) {
    SNRSE_CHECK_ARG(conv_halo_eligible(a0, 9, n_rows), "conv_halo: shape not eligible");
    SNRSE_CHECK_ARG(a0->C % 64 == 0 && a0->ld % 8 == 0, "conv_halo: Cin must be a multiple of 64");
    SNRSE_CHECK_ARG(!a1 || (a1->C % 64 == 0 && a1->ld % 8 == 0 && a1->H == a0->H && a1->W == a0->W && a1->B == a0->B),
                    "conv_halo: bad shortcut operand");
    SNRSE_CHECK_ARG(out_ld % 8 == 0 && (!res || res->ld % 8 == 0), "conv_halo: pitches must be multiples of 8");
    if (g_num_sms == 0) {
        int dev = 0;
        SNRSE_CUDA(cudaGetDevice(&dev));
        SNRSE_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    memset(p, 0, sizeof(*p));
    const int tiles_w = cdiv(a0->W, TW);
    // two sub-tiles per weight tile when that still leaves every SM a super-tile
    int sub = 2;
    if (a0->H < 16 || (int64_t)a0->B * cdiv(a0->H, 16) * tiles_w < g_num_sms) sub = 1;
    p->sub = sub;
    p->c0_chunks = a0->C / 64;
    p->c1_chunks = a1 ? a1->C / 64 : 0;
    p->B = a0->B; p->H = a0->H; p->W = a0->W;
    p->tiles_h = cdiv(a0->H, SUB_ROWS * sub);
    p->tiles_w = tiles_w;
    p->n_tiles = a0->B * p->tiles_h * p->tiles_w;
    p->N = n_rows;
    p->acc_bufs = (sub * n_rows <= 256) ? 2 : 1;
    const int a_bytes = (SUB_ROWS * sub + 2) * TW * 128, b_bytes = n_rows * 128;
    // ring depths within ~220 KB: at least 3 halo copies, the rest to weight tiles
    int na = 3, nb = (220 * 1024 - na * a_bytes) / b_bytes;
    if (nb > MAX_B) nb = MAX_B;
    if (nb >= 6 && na < MAX_A && (220 * 1024 - (na + 1) * a_bytes) / b_bytes >= 5) {
        na += 1;
        nb = (220 * 1024 - na * a_bytes) / b_bytes;
        if (nb > MAX_B) nb = MAX_B;
    }
    SNRSE_CHECK_ARG(nb >= 3, "conv_halo: shared memory budget");
    p->na = na; p->nb = nb;
    p->smem_bytes = na * a_bytes + nb * b_bytes + 1024;
    p->grid = p->n_tiles < g_num_sms ? p->n_tiles : g_num_sms;
    p->bias = bias; p->tbias = tbias; p->tb_stride = tb_stride;
    p->res = res ? res->ptr : nullptr; p->res_ld = res ? res->ld : 0;
    p->scale = scale; p->out = out; p->out_ld = out_ld;
    const int box_h = SUB_ROWS * sub + 2;
    SNRSE_TRY(tma_make_act_map(&p->mapA0, a0->ptr, a0->C, a0->W, a0->H, a0->B, a0->ld, 64, TW, box_h));
    if (a1) SNRSE_TRY(tma_make_act_map(&p->mapA1, a1->ptr, a1->C, a1->W, a1->H, a1->B, a1->ld, 64, TW, box_h));
    else p->mapA1 = p->mapA0;
    const int64_t ktot = 64 * (int64_t)(9 * p->c0_chunks + p->c1_chunks);
    SNRSE_TRY(tma_make_wt_map(&p->mapB, wt, ktot, n_rows, 1, ktot * n_rows, 64, n_rows));
    return SNRSE_OK;
}

int conv_halo_launch(const ConvHaloPlan* p, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        SNRSE_CUDA(cudaFuncSetAttribute(conv_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
        attr_set = true;
    }
    HaloArgs g;
    g.c0_chunks = p->c0_chunks; g.c1_chunks = p->c1_chunks;
    g.H = p->H; g.W = p->W; g.B = p->B;
    g.sub = p->sub; g.tiles_h = p->tiles_h; g.tiles_w = p->tiles_w; g.n_tiles = p->n_tiles;
    g.N = p->N; g.na = p->na; g.nb = p->nb; g.acc_bufs = p->acc_bufs;
    g.bias = p->bias; g.tbias = p->tbias; g.tb_stride = p->tb_stride;
    g.res = p->res; g.res_ld = p->res_ld; g.scale = p->scale; g.out = p->out; g.out_ld = p->out_ld;
    g.dbg = g_halo_dbg;
    conv_halo_kernel<<<p->grid, HALO_THREADS, p->smem_bytes, s>>>(p->mapA0, p->mapA1, p->mapB, g);
    SNRSE_LAUNCH_CHECK();
    return SNRSE_OK;
}

// tcgen05.mma issue-rate probe for sm_100a (measurement tool, not product code).
//
// Question it answers: how many SM clocks does one bf16 tcgen05.mma (K=16) take as a function of
// M, N, cta_group, operand source (shared memory descriptor vs TMEM for A) and of concurrent bulk-copy
// traffic into shared memory?  The implicit-GEMM convolution kernels are designed around the answer
// (profiles/r01_mma_probe.md).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I snr_aligned_diffse_b200/csrc tools/mma_probe.cu -o tools/mma_probe
//   tools/mma_probe            (prints one JSON line per configuration)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "ptx.cuh"

namespace {

struct ProbeArgs {
    int M;          // 128 (cta_group 1) or 256 (cta_group 2: 128 rows per CTA)
    int N;          // MMA N
    int n_iter;     // iterations of 4 K16 MMAs
    int a_tmem;     // 1: A operand from TMEM
    int traffic;    // 0 none, 1 bulk copies global->shared in a side warp, 2 = two side warps
    int same_addr;  // 1: every MMA reads the same operand addresses (no ring)
    int chunk;      // bulk copy size in bytes
    int issuers;    // 1 or 2 issuing warps (each into its own accumulator)
    int commit_every; // MMAs between commits (4 default)
    int random_data;  // 0: zero operands, 1: pseudo-random bf16 operands in (-1, 1)
    const uint8_t* gsrc;
    long long* out; // per CTA: cycles, mma count, traffic bytes
};

__device__ __forceinline__ void mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

template <int CG>
__device__ __forceinline__ void probe_body(const ProbeArgs g) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t done_bar, done_bar2, dummy_bar, dummy_bar2, tr_bar[2][4];
    __shared__ uint32_t tmem_base_smem;
    __shared__ volatile int stop_flag;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const uint32_t rank = CG == 2 ? (blockIdx.x & 1u) : 0u;
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    // layout: A ring 4 x 16 KB | B ring 4 x 32 KB | traffic scratch 2 x 16 KB  = 224 KB - slack
    const uint32_t a_base = smem_base, b_base = smem_base + 4 * 16384, t_base = b_base + 4 * 32768;
    if (threadIdx.x == 0) {
        ptx::mbar_init(ptx::smem_u32(&done_bar), 1);
        ptx::mbar_init(ptx::smem_u32(&dummy_bar), 1);
        ptx::mbar_init(ptx::smem_u32(&done_bar2), 1);
        ptx::mbar_init(ptx::smem_u32(&dummy_bar2), 1);
        for (int w = 0; w < 2; ++w)
            for (int i = 0; i < 4; ++i) ptx::mbar_init(ptx::smem_u32(&tr_bar[w][i]), 1);
        ptx::fence_barrier_init();
        stop_flag = 0;
    }
    // zero-fill operands so the accumulators stay finite
    for (uint32_t i = threadIdx.x * 16; i < 4 * 16384 + 4 * 32768; i += blockDim.x * 16)
    {
        uint32_t w = 0;
        if (g.random_data) {   // two bf16 per word: sign + exponent 0x3e/0x3f + random mantissa
            uint32_t h = (i * 2654435761u) ^ (blockIdx.x * 40503u);
            h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
            w = (h & 0x807f807fu) | 0x3e803e80u | ((h >> 3) & 0x01000100u);
        }
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_base + i), "r"(w), "r"(w * 3u | 0x3e003e00u & 0xbfffbfffu), "r"(w ^ 0x00550055u), "r"(w ^ 0x80008000u) : "memory");
    }
    ptx::fence_proxy_async();
    __syncthreads();
    if (warp == 0) {
        if (CG == 2) { ptx::tmem_alloc2(ptx::smem_u32(&tmem_base_smem), 512); ptx::tmem_relinquish2(); }
        else { ptx::tmem_alloc(ptx::smem_u32(&tmem_base_smem), 512); ptx::tmem_relinquish(); }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (CG == 2) ptx::cluster_sync();
    ptx::tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_smem, 0);

    if (warp == 1 && rank != 0) {
        // peer CTA of a pair: the leader's final commit is multicast to this CTA's barrier too
        if (lane == 0) {
            ptx::mbar_wait(ptx::smem_u32(&done_bar), 0);
            stop_flag = 1;
        }
    } else if (warp == 1) {
        // warp-uniform issue loop, one elected lane issues (descriptors stay in uniform registers)
        const uint32_t idesc = ptx::umma_idesc_bf16((uint32_t)g.M, (uint32_t)g.N);
        const bool elected = ptx::elect_one();
        const uint32_t d_alt = g.a_tmem ? (2 * g.N <= 480 ? (uint32_t)g.N : 0u) : (2 * g.N <= 512 ? (uint32_t)g.N : 0u);
        const long long t0 = clock64();
        for (int it = 0; it < g.n_iter; ++it) {
            const uint32_t st = g.same_addr ? 0u : (uint32_t)(it & 3);
            const uint64_t da = ptx::umma_desc_k_sw128(a_base + st * 16384);
            const uint64_t db = ptx::umma_desc_k_sw128(b_base + st * 32768);
            const uint32_t d = tmem_base + ((it & 1) ? d_alt : 0u);
            if (elected) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (CG == 2) ptx::mma_bf16_ss_2sm(d, da + 2 * k, db + 2 * k, idesc, 1u);
                    else if (g.a_tmem) mma_bf16_ts(d, tmem_base + 480 + 8 * (k & 3), db + 2 * k, idesc, 1u);   // A: 8 columns per K16
                    else ptx::mma_bf16_ss(d, da + 2 * k, db + 2 * k, idesc, 1u);
                }
                if ((it & ((g.commit_every >> 2) - 1)) == 0) {
                    if (CG == 2) ptx::mma_commit_2sm(ptx::smem_u32(&dummy_bar)); else ptx::mma_commit(ptx::smem_u32(&dummy_bar));
                }
            }
            __syncwarp();
        }
        if (elected) {
            if (CG == 2) ptx::mma_commit_2sm(ptx::smem_u32(&done_bar)); else ptx::mma_commit(ptx::smem_u32(&done_bar));
        }
        ptx::mbar_wait(ptx::smem_u32(&done_bar), 0);
        if (g.issuers == 2) ptx::mbar_wait(ptx::smem_u32(&done_bar2), 0);
        const long long t1 = clock64();
        if (elected) {
            stop_flag = 1;
            g.out[blockIdx.x * 4 + 0] = t1 - t0;
            g.out[blockIdx.x * 4 + 1] = 4LL * g.n_iter * g.issuers;
        }
    } else if (warp == 3 && g.issuers == 2 && rank == 0) {
        const uint32_t idesc = ptx::umma_idesc_bf16((uint32_t)g.M, (uint32_t)g.N);
        const bool elected = ptx::elect_one();
        for (int it = 0; it < g.n_iter; ++it) {
            const uint32_t st = (uint32_t)(it & 3);
            const uint64_t da = ptx::umma_desc_k_sw128(a_base + st * 16384);
            const uint64_t db = ptx::umma_desc_k_sw128(b_base + st * 32768);
            const uint32_t d = tmem_base + 256u;
            if (elected) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (CG == 2) ptx::mma_bf16_ss_2sm(d, da + 2 * k, db + 2 * k, idesc, 1u);
                    else ptx::mma_bf16_ss(d, da + 2 * k, db + 2 * k, idesc, 1u);
                }
                if (CG == 2) ptx::mma_commit_2sm(ptx::smem_u32(&dummy_bar2)); else ptx::mma_commit(ptx::smem_u32(&dummy_bar2));
            }
            __syncwarp();
        }
        if (elected) {
            if (CG == 2) ptx::mma_commit_2sm(ptx::smem_u32(&done_bar2)); else ptx::mma_commit(ptx::smem_u32(&done_bar2));
        }
        ptx::mbar_wait(ptx::smem_u32(&done_bar2), 0);
    } else if ((warp == 2 || warp == 3) && g.traffic == 3) {
        // SIMT shared-memory traffic: each lane reads 16 B, modifies, writes back (conflict-free rows), until stopped
        const int w = warp - 2;
        long long bytes = 0;
        uint32_t addr = t_base + (uint32_t)w * 16384 + (uint32_t)lane * 16;
        uint32_t k = 0;
        while (!stop_flag) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                uint4 v = ptx::lds128(addr + ((k + u) & 31u) * 512u);
                v.x += 1u;
                ptx::sts128(addr + ((k + u) & 31u) * 512u, v);
            }
            k += 8;
            bytes += 8 * 32 * 32;   // 8 x (16 B read + 16 B write) x 32 lanes
        }
        if (lane == 0) g.out[blockIdx.x * 4 + 2 + w] = bytes;
    } else if ((warp == 2 || warp == 3) && lane == 0 && g.traffic >= warp - 1 && !(warp == 3 && g.issuers == 2)) {
        // side traffic: bulk copies global -> shared, 4 in flight
        const int w = warp - 2;
        const uint8_t* src = g.gsrc + ((size_t)blockIdx.x * 2 + w) * 65536;
        long long bytes = 0;
        uint32_t it = 0;
        const long long t0 = clock64();
        while (!stop_flag) {
            const uint32_t s = it & 3, ph = (it >> 2) & 1;
            if (it >= 4) ptx::mbar_wait(ptx::smem_u32(&tr_bar[w][s]), ph ^ 1u);
            ptx::mbar_arrive_expect_tx(ptx::smem_u32(&tr_bar[w][s]), (uint32_t)g.chunk);
            bulk_g2s(t_base + (uint32_t)w * 16384 + (s * (uint32_t)g.chunk) % 16384, src + (s * g.chunk) % 65536, (uint32_t)g.chunk,
                     ptx::smem_u32(&tr_bar[w][s]));
            bytes += g.chunk;
            ++it;
        }
        // drain
        for (uint32_t j = (it >= 4 ? it - 4 : 0); j < it; ++j) ptx::mbar_wait(ptx::smem_u32(&tr_bar[w][j & 3]), (j >> 2) & 1);
        g.out[blockIdx.x * 4 + 2 + w] = bytes;
        (void)t0;
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (CG == 2) ptx::cluster_sync();
    if (warp == 0) {
        if (CG == 2) ptx::tmem_dealloc2(tmem_base, 512); else ptx::tmem_dealloc(tmem_base, 512);
    }
}

__global__ void __launch_bounds__(128, 1) probe1(const ProbeArgs g) { probe_body<1>(g); }
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe2(const ProbeArgs g) { probe_body<2>(g); }

#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); exit(1); } } while (0)

void run(int cg, int M, int N, int a_tmem, int traffic, int same_addr, int chunk, int grid, const uint8_t* gsrc, long long* dout,
         int issuers = 1, int commit_every = 4, int random_data = 0, int n_iter = 2048) {
    ProbeArgs g{M, N, n_iter, a_tmem, traffic, same_addr, chunk, issuers, commit_every, random_data, gsrc, dout};
    const int smem = 4 * 16384 + 4 * 32768 + 2 * 16384 + 1024;
    CK(cudaMemset(dout, 0, sizeof(long long) * 4 * 148));
    if (cg == 1) {
        CK(cudaFuncSetAttribute(probe1, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        probe1<<<grid, 128, smem>>>(g);
    } else {
        CK(cudaFuncSetAttribute(probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        probe2<<<grid, 128, smem>>>(g);
    }
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    static long long h[4 * 148];
    CK(cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost));
    double cyc = 0, n = 0, tb = 0, cmax = 0;
    int cnt = 0;
    for (int i = 0; i < grid; ++i) {
        if (h[i * 4 + 1] == 0) continue;
        cyc += (double)h[i * 4]; n += (double)h[i * 4 + 1]; ++cnt;
        if ((double)h[i * 4] > cmax) cmax = (double)h[i * 4];
    }
    for (int i = 0; i < grid; ++i) tb += (double)(h[i * 4 + 2] + h[i * 4 + 3]);
    const double per = cyc / n;
    const double macs = (double)M * N * 16;
    printf("{\"cg\": %d, \"M\": %d, \"N\": %d, \"a_tmem\": %d, \"traffic\": %d, \"same_addr\": %d, \"chunk\": %d, \"grid\": %d, "
           "\"issuers\": %d, \"commit_every\": %d, \"random_data\": %d, \"n_iter\": %d, \"cycles_per_mma\": %.1f, \"max_cta_cycles_per_mma\": %.1f, \"mac_per_clk_per_sm\": %.0f, \"traffic_B_per_clk_per_sm\": %.1f}\n",
           cg, M, N, a_tmem, traffic, same_addr, chunk, grid, issuers, commit_every, random_data, n_iter, per, cmax / (n / cnt), macs / per / cg,
           tb / (cyc / cnt) / grid);
    fflush(stdout);
}

}  // namespace

int main() {
    uint8_t* gsrc;
    long long* dout;
    CK(cudaMalloc(&gsrc, 148 * 2 * 65536));
    CK(cudaMemset(gsrc, 0, 148 * 2 * 65536));
    CK(cudaMalloc(&dout, sizeof(long long) * 4 * 148));
    // SIMT LDS/STS traffic (2 warps) next to the MMAs
    run(2, 256, 128, 0, 0, 0, 8192, 148, gsrc, dout, 1, 4, 1, 8192);
    run(2, 256, 128, 0, 3, 0, 8192, 148, gsrc, dout, 1, 4, 1, 8192);
    run(2, 256, 256, 0, 3, 0, 8192, 148, gsrc, dout, 1, 4, 1, 8192);
    run(1, 128, 256, 0, 3, 0, 8192, 148, gsrc, dout, 1, 4, 1, 8192);
    run(2, 256, 32, 0, 3, 0, 8192, 148, gsrc, dout, 1, 4, 1, 8192);
    return 0;
}

// GPU probe (diagnostic, not a test): how fast does ONE SM retire tcgen05.mma kind::tf32 instructions as a function of
// the operand source and the tile shape?  Answers the question behind the narrow-layer gap of the tap-GEMM:
//   * SS mode (A and B from shared memory), M = 128, N = 64 / 128 / 256
//   * TS mode (A from TMEM, placed there by tcgen05.cp.128x256b from the K-major SWIZZLE_128B tile TMA would write)
//   * M = 64 (the swapped formulation: channels as M, pixels as N = 256)
//   * tcgen05.cp alone, and cp + TS MMA interleaved (what a conv tap would issue)
// and checks that cp + TS gives bit-identical accumulators to SS on the same tile.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I rgb-proprioceptive-pose-estimator_b200/csrc
//             -o tests/probe_mma_rate.bin tests/probe_mma_rate.cu        Run on the GPU box: tests/probe_mma_rate.bin
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "pe_common.cuh"

using namespace pe;

__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_cp_128x256b(uint32_t tmem_dst, uint64_t sdesc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tmem_dst), "l"(sdesc) : "memory");
}

// K-major SWIZZLE_128B element offset (floats) of (row, k) in a [rows][32] fp32 tile: 16-byte chunk index XOR (row % 8)
__device__ __forceinline__ int sw128(int row, int k) {
    return row * 32 + ((((k >> 2) ^ (row & 7)) << 2) | (k & 3));
}

struct Result {
    long long clk[40];
    int mismatch_ts, mismatch_ref;
    int m64_lane_of_row[64];
};

__device__ __forceinline__ void tc_mma_tf32_ts_elect(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                                     uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_cp_128x256b_elect(uint32_t tmem_dst, uint64_t sdesc) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.cp.cta_group::1.128x256b [%0], %1;\n\t}" ::"r"(tmem_dst), "l"(sdesc) : "memory");
}

// Whole warp, convergent (as the tap-GEMM's MMA warp issues): 8 instructions per loop trip, straight-line.
// mode: 0 SS M128, 1 TS M128 (A preloaded), 2 SS M64, 3 cp only, 4 cp + TS per instruction, 5/6/7: 2/4/8 independent
// accumulators, 8: M64 with 2 accumulators, 9: 2 accumulators in runs of 4
template <int MODE>
__device__ long long run_case(int N, int reps, uint32_t tmem, uint32_t sA, uint32_t sB, uint32_t bar, uint32_t& phase) {
    const uint32_t idesc128 = make_idesc_tf32(128, N, 0, 0), idesc64 = make_idesc_tf32(64, N, 0, 0);
    const uint64_t adesc = make_smem_desc(sA, 16, 1024, 2), bdesc = make_smem_desc(sB, 16, 1024, 2);
    const uint32_t d = tmem, a_t = tmem + 448;
    __syncwarp();
    long long t0 = clock64();
    for (int i0 = 0; i0 < reps; i0 += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int slab = u & 3;
            const uint64_t ad = adesc + (uint64_t)(slab * 2), bd = bdesc + (uint64_t)(slab * 2);   // +32 B per K = 8 slab
            if (MODE == 0) tc_mma_tf32_elect(d, ad, bd, idesc128, 1);
            if (MODE == 1) tc_mma_tf32_ts_elect(d, a_t + slab * 8, bd, idesc128, 1);
            if (MODE == 2) tc_mma_tf32_elect(d, ad, bd, idesc64, 1);
            if (MODE == 3) tc_cp_128x256b_elect(a_t + u * 8, ad);
            if (MODE == 4) {
                tc_cp_128x256b_elect(a_t + u * 8, ad);
                tc_mma_tf32_ts_elect(d, a_t + u * 8, bd, idesc128, 1);
            }
            if (MODE == 5) tc_mma_tf32_elect(d + (uint32_t)((u & 1) * 256), ad, bd, idesc128, 1);
            if (MODE == 6) tc_mma_tf32_elect(d + (uint32_t)((u & 3) * 128), ad, bd, idesc128, 1);
            if (MODE == 7) tc_mma_tf32_elect(d + (uint32_t)(u * 64), ad, bd, idesc128, 1);
            if (MODE == 8) tc_mma_tf32_elect(d + (uint32_t)((u & 1) * 256), ad, bd, idesc64, 1);
            if (MODE == 9) tc_mma_tf32_elect(d + (uint32_t)(((u >> 2) & 1) * 256), ad, bd, idesc128, 1);
        }
    }
    tc_commit_elect(bar);
    mbar_wait_warp(bar, phase);
    phase ^= 1;
    long long t1 = clock64();
    return t1 - t0;
}

template <int MODE>
__device__ void time_case(Result* res, int& slot, uint32_t tmem, uint32_t sA, uint32_t sB, uint32_t bar, uint32_t& phase) {
    for (int N = 64; N <= 256; N *= 2) {
        if (MODE == 3 && N != 64) continue;
        if (MODE == 6 && N > 128) continue;
        if (MODE == 7 && N > 64) continue;
        run_case<MODE>(N, 64, tmem, sA, sB, bar, phase);
        const long long c = run_case<MODE>(N, 2048, tmem, sA, sB, bar, phase);
        if ((threadIdx.x & 31) == 0) res->clk[slot] = c;
        ++slot;
    }
}

__global__ void __launch_bounds__(128) probe_kernel(Result* res, float* dump_ss, float* dump_ts, float* dump_m64) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    float* A = (float*)base;                  // 128 x 32 (16 KB)
    float* B = (float*)(base + 16384);        // 256 x 32 (32 KB)
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bar_store;
    const int tid = threadIdx.x, warp = tid >> 5;
    // exactly representable small integers
    for (int i = tid; i < 128 * 32; i += 128) {
        int r = i >> 5, k = i & 31;
        A[sw128(r, k)] = (float)(((r * 37 + k * 11) % 127) - 63);
    }
    for (int i = tid; i < 256 * 32; i += 128) {
        int r = i >> 5, k = i & 31;
        B[sw128(r, k)] = (float)(((r * 5 + k * 11) % 9) - 4);
    }
    fence_proxy_async_smem();
    const uint32_t bar = smem_u32(&bar_store);
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        tmem_alloc(smem_u32(&tmem_slot), 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t sA = smem_u32(A), sB = smem_u32(B);
    uint32_t phase = 0;

    // ---- correctness: SS into columns [0,256), cp + TS into columns [256, 512-64)?  N = 128 keeps both in range
    const int Nc = 128;
    if (tid == 0) {
        const uint32_t idesc = make_idesc_tf32(128, Nc, 0, 0);
        const uint64_t adesc = make_smem_desc(sA, 16, 1024, 2), bdesc = make_smem_desc(sB, 16, 1024, 2);
        for (int s = 0; s < 4; ++s) tc_mma_tf32(tmem, adesc + 2 * s, bdesc + 2 * s, idesc, s > 0);
        for (int s = 0; s < 4; ++s) {
            tc_cp_128x256b(tmem + 448 + 8 * s, adesc + 2 * s);
            tc_mma_tf32_ts(tmem + 256, tmem + 448 + 8 * s, bdesc + 2 * s, idesc, s > 0);
        }
        tc_commit(bar);
        mbar_wait(bar, phase);
    }
    phase ^= 1;
    __syncthreads();
    tc_fence_after();
    {
        float v[32];
        for (int c = 0; c < Nc; c += 32) {
            tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
            for (int j = 0; j < 32; ++j) dump_ss[(size_t)tid * Nc + c + j] = v[j];
            tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + 256 + c, v);
            for (int j = 0; j < 32; ++j) dump_ts[(size_t)tid * Nc + c + j] = v[j];
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // ---- M = 64 layout: zero D, one K = 32 product with M = 64, N = 64, dump all 128 lanes
    if (tid == 0) {
        const uint32_t idesc = make_idesc_tf32(64, 64, 0, 0);
        const uint64_t adesc = make_smem_desc(sA, 16, 1024, 2), bdesc = make_smem_desc(sB, 16, 1024, 2);
        for (int s = 0; s < 4; ++s) tc_mma_tf32(tmem + 128, adesc + 2 * s, bdesc + 2 * s, idesc, s > 0);
        tc_commit(bar);
        mbar_wait(bar, phase);
    }
    phase ^= 1;
    __syncthreads();
    tc_fence_after();
    {
        float v[32];
        for (int c = 0; c < 64; c += 32) {
            tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + 128 + c, v);
            for (int j = 0; j < 32; ++j) dump_m64[(size_t)tid * 64 + c + j] = v[j];
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // ---- timing
    if (warp == 0) {
        int slot = 0;
        time_case<0>(res, slot, tmem, sA, sB, bar, phase);
        time_case<1>(res, slot, tmem, sA, sB, bar, phase);
        time_case<2>(res, slot, tmem, sA, sB, bar, phase);
        time_case<3>(res, slot, tmem, sA, sB, bar, phase);
        time_case<4>(res, slot, tmem, sA, sB, bar, phase);
        time_case<5>(res, slot, tmem, sA, sB, bar, phase);
        time_case<6>(res, slot, tmem, sA, sB, bar, phase);
        time_case<7>(res, slot, tmem, sA, sB, bar, phase);
        time_case<8>(res, slot, tmem, sA, sB, bar, phase);
        time_case<9>(res, slot, tmem, sA, sB, bar, phase);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
    Result* res;
    float *dss, *dts, *dm64;
    cudaMalloc(&res, sizeof(Result));
    cudaMalloc(&dss, 128 * 128 * 4);
    cudaMalloc(&dts, 128 * 128 * 4);
    cudaMalloc(&dm64, 128 * 64 * 4);
    cudaMemset(res, 0, sizeof(Result));
    cudaMemset(dm64, 0, 128 * 64 * 4);
    const int smem = 16384 + 32768 + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe_kernel<<<1, 128, smem>>>(res, dss, dts, dm64);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("CUDA error: %s\n", cudaGetErrorString(e));
        return 1;
    }
    Result h;
    std::vector<float> ss(128 * 128), ts(128 * 128), m64(128 * 64);
    cudaMemcpy(&h, res, sizeof(Result), cudaMemcpyDeviceToHost);
    cudaMemcpy(ss.data(), dss, ss.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(ts.data(), dts, ts.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(m64.data(), dm64, m64.size() * 4, cudaMemcpyDeviceToHost);
    auto Af = [](int r, int k) { return (float)(((r * 37 + k * 11) % 127) - 63); };
    auto Bf = [](int r, int k) { return (float)(((r * 5 + k * 11) % 9) - 4); };
    int bad_ss = 0, bad_ts = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 128; ++n) {
            float ref = 0;
            for (int k = 0; k < 32; ++k) ref += Af(m, k) * Bf(n, k);
            bad_ss += ss[m * 128 + n] != ref;
            bad_ts += ts[m * 128 + n] != ref;
        }
    printf("correctness: SS mismatches %d, cp+TS mismatches %d (of 16384)\n", bad_ss, bad_ts);
    // M = 64 layout: which TMEM lane holds row r?
    printf("M=64 layout (row -> lane): ");
    for (int r = 0; r < 64; ++r) {
        int found = -1;
        for (int lane = 0; lane < 128 && found < 0; ++lane) {
            bool ok = true;
            for (int n = 0; n < 64 && ok; ++n) {
                float ref = 0;
                for (int k = 0; k < 32; ++k) ref += Af(r, k) * Bf(n, k);
                ok = m64[lane * 64 + n] == ref;
            }
            if (ok) found = lane;
        }
        printf("%d:%d ", r, found);
    }
    printf("\n");
    const char* names[10] = {"SS  M128", "TS  M128 (A resident in TMEM)", "SS  M64", "cp 128x256b only", "cp + TS M128",
                             "SS  M128, 2 accumulators", "SS  M128, 4 accumulators", "SS  M128, 8 accumulators",
                             "SS  M64, 2 accumulators", "SS  M128, 2 acc, runs of 4"};
    int slot = 0;
    for (int mi = 0; mi < 10; ++mi)
        for (int N = 64; N <= 256; N *= 2) {
            if (mi == 3 && N != 64) continue;
            if (mi == 6 && N > 128) continue;
            if (mi == 7 && N > 64) continue;
            const double c = (double)h.clk[slot++] / 2048.0;
            const int M = (mi == 2 || mi == 8) ? 64 : 128;
            if (mi == 3)
                printf("%-32s        : %7.1f clk / instr\n", names[mi], c);
            else
                printf("%-32s N = %3d: %7.1f clk / instr   (%5.1f %% of 2048 MAC/clk, %d x %d x 8)\n", names[mi], N, c,
                       100.0 * M * N * 8 / c / 2048.0, M, N);
        }
    return 0;
}

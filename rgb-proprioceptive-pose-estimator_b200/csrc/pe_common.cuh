// Common device/host helpers for the pose-estimator sm_100a kernels.
// PTX wrappers for mbarrier / TMA / tcgen05 (Blackwell), error plumbing.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace pe {

// ---------------------------------------------------------------------------------------------
// error plumbing (host)
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
const char* get_error();

#define PE_CHECK_CUDA(expr)                                                                   \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            pe::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return 2;                                                                         \
        }                                                                                     \
    } while (0)

#define PE_REQUIRE(cond, ...)                                                                 \
    do {                                                                                      \
        if (!(cond)) {                                                                        \
            pe::set_error(__VA_ARGS__);                                                       \
            return 1;                                                                         \
        }                                                                                     \
    } while (0)

#define PE_LAUNCH_CHECK()                                                                     \
    do {                                                                                      \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess) {                                                              \
            pe::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return 2;                                                                         \
        }                                                                                     \
    } while (0)

int num_sms();
// the sticky device-side error flag (one int, allocated on first use; nullptr if that fails): bits 1-32 tap-GEMM
// pipeline timeouts, 64 grid-barrier timeout of the persistent LSTM kernels
int* device_error_flag();
void lstm_seq_reset();      // re-arm the LSTM kernels' grid barrier (after a reported timeout)
// programmatic dependent launch: bit 0 = tap-GEMM launches, bit 1 = streaming / elementwise launches carry the
// attribute (PE_B200_PDL=<mask> or pe_debug_pdl(mask); default in pe_tapgemm.cu)
bool pdl_enabled(int kind = 1);

#ifdef __CUDACC__
// Launch with the programmatic-stream-serialization attribute: the grid may be scheduled while the previous kernel
// of the stream is still draining, so launch latency and ramp-up hide behind that kernel's tail.  Every kernel
// launched this way calls pdl_sync() before it touches global memory.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled(2) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// kernel<<<grid, block, smem, stream>>>(args...) through launch_pdl, inside an entry point with a `stream` argument
#define PE_LAUNCH(kernel, grid, block, smem, ...) \
    PE_CHECK_CUDA(pe::launch_pdl(kernel, grid, block, smem, (cudaStream_t)stream, __VA_ARGS__))

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
// First statement of a kernel launched with launch_pdl: let the next kernel of the stream be scheduled, then wait
// until the previous grid has completed and its writes are visible (a no-op without the launch attribute).
__device__ __forceinline__ void pdl_sync() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ float round_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// ---- mbarrier -------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done;
}
// Bounded wait: returns false after PE_WAIT_LIMIT SM cycles (2^30, ~0.55 s at 1.9 GHz: long enough to ride out a
// time-slice / MPS preemption, short enough that a protocol bug cannot hang the GPU).  A timeout sets the sticky
// device flag (pe_device_error); pe_poison_on_error turns it into NaNs in the step's result so that a training
// or rollout loop cannot silently continue on unwritten tiles.
#define PE_WAIT_LIMIT (1ll << 30)
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return true;
    long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > PE_WAIT_LIMIT) return false;
    }
    return true;
}

// Warp-convergent variant: ALL 32 lanes call it and get the same (vote-derived, hence provably uniform)
// answer, so the code that follows stays on the uniform datapath -- descriptors, coordinates and loop
// counters live in uniform registers and the single-lane issue (elect.sync inside the *_elect wrappers
// below) costs a handful of instructions instead of a vector->uniform waterfall per TMA / MMA.
__device__ __forceinline__ uint32_t mbar_test_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done;
}
__device__ __forceinline__ bool mbar_wait_warp(uint32_t bar, uint32_t parity, int spin = 0) {
    if (spin) {
        // non-blocking poll (latency-critical producer <-> MMA hand-offs)
        if (__all_sync(0xffffffffu, mbar_test_wait(bar, parity))) return true;
        const long long t0 = clock64();
        for (;;) {
            if (__all_sync(0xffffffffu, mbar_test_wait(bar, parity))) return true;
            if (__any_sync(0xffffffffu, clock64() - t0 > PE_WAIT_LIMIT)) return false;
        }
    }
    if (__all_sync(0xffffffffu, mbar_try_wait(bar, parity))) return true;
    const long long t0 = clock64();
    for (;;) {
        if (__all_sync(0xffffffffu, mbar_try_wait(bar, parity))) return true;
        if (__any_sync(0xffffffffu, clock64() - t0 > PE_WAIT_LIMIT)) return false;
    }
}
__device__ __forceinline__ void mbar_arrive_expect_tx_elect(uint32_t bar, uint32_t bytes) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}"
        ::"r"(bar), "r"(bytes)
        : "memory");
}

// ---- TMA ------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar,
                                            int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d_elect(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar,
                                                  int c0, int c1, int c2, int c3) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4, %5, %6}], [%2];\n\t}"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar,
                                            int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1,
                                             int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
        ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_store_commit() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_all() {
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- tcgen05 / TMEM -------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_slot),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tc_commit_elect(uint32_t bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(bar)
        : "memory");
}
// Convergent-warp issue: every lane executes the block, one elected lane issues the MMA.
__device__ __forceinline__ void tc_mma_tf32_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                                  uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], TF32 operands, fp32 accumulate. Issued by ONE thread.
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                            uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets TMEM lane (base_lane + i).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- CTA pairs (cta_group::2): two CTAs of a cluster on the SMs of one TPC run ONE 256-row MMA -------------------
// The shared::cta address of a barrier carries the CTA's rank in bit 24 when it is used in the shared::cluster window;
// clearing it addresses the same barrier in the pair's even ("leader") CTA.
constexpr uint32_t PE_PEER_BIT_MASK = 0xFEFFFFFFu;

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// all threads of both CTAs
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load into THIS CTA's shared memory whose transaction bytes are counted on the LEADER CTA's barrier
__device__ __forceinline__ void tma_load_4d_pair_elect(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar,
                                                       int c0, int c1, int c2, int c3) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4, %5, %6}], [%2];\n\t}"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & PE_PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3)
        : "memory");
}
// arrive on the leader CTA's copy of a barrier (from either CTA of the pair)
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & PE_PEER_BIT_MASK) : "memory");
}
// one warp of EACH CTA of the pair, same warp index, same shared-memory slot offset
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_slot), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// completion of the pair's MMAs signalled on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void tc_commit_pair_elect(uint32_t bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t.reg .b16 m;\n\t"
        "mov.b16 m, 3;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}"
        ::"r"(bar)
        : "memory");
}
// D[tmem, 256 rows over the pair] (+)= A * B: issued by one elected lane of the LEADER CTA; A rows 0-127 / 128-255 and
// B rows (N columns) 0..N/2-1 / N/2..N-1 are read from the two CTAs' shared memory at the descriptors' offsets
__device__ __forceinline__ void tc_mma_tf32_pair_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                       uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// Shared-memory matrix descriptor (sm_100 "version 1").
//  K-major : rows of 128 B (32 tf32 along K), 8-row groups SBO bytes apart.
//  MN-major: rows of 128 B (32 tf32 along M/N), one MMA (K=8) reads 8 rows = two 4-row swizzle
//            atoms SBO (= 512) bytes apart; 32-wide M/N atoms are LBO bytes apart.
//  layout_type: 2 = SWIZZLE_128B (16 B atoms; K-major operands),
//               1 = SWIZZLE_128B_BASE32B (32 B atoms, 4-row period; the only layout tcgen05
//                   accepts for MN-major TF32 operands -- TMA mode SWIZZLE_128B_ATOM_32B).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes,
                                                   uint32_t sbo_bytes, uint32_t layout_type) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= 1ull << 46;  // descriptor version (Blackwell)
    d |= static_cast<uint64_t>(layout_type) << 61;
    return d;
}
// Instruction descriptor for kind::tf32, fp32 accumulate.
__host__ __device__ inline uint32_t make_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
    uint32_t d = 0;
    d |= 1u << 4;   // D format = F32
    d |= 2u << 7;   // A format = TF32
    d |= 2u << 10;  // B format = TF32
    d |= (a_mn_major ? 1u : 0u) << 15;
    d |= (b_mn_major ? 1u : 0u) << 16;
    d |= static_cast<uint32_t>(N >> 3) << 17;
    d |= static_cast<uint32_t>(M >> 4) << 24;
    return d;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif  // __CUDACC__

}  // namespace pe

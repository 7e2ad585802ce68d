// Persistent LSTM recurrence: ONE launch walks all S timesteps of a single-layer seq-major nn.LSTM (gates i,f,g,o;
// reference models/time_sensitive.py:126-131,418,501-510), forward and backward-through-time.
//
// The input projection x W_ih^T of all S*N rows is one tensor-core GEMM done up front (pe_linear_fwd); what is
// strictly sequential is  gates_t = gx_t + h_{t-1} W_hh^T + b,  cell,  h_t.  The per-step work is tiny (N x H x 4H
// MACs), so the old path -- one GEMM launch + one cell launch per timestep, 2 S launches forward and 2 S backward --
// was latency, not arithmetic.  Here W_hh is SLICED ACROSS THE GRID and stays resident in shared memory for the
// whole sequence: CTA b owns U hidden units (4 U gate rows of W_hh forward; the same U columns of W_hh backward),
// stages h_{t-1} (forward) / dgates_t (backward) through shared memory, finishes its dot products in fp32 FMA with
// the LSTM cell fused behind them, and meets the other CTAs at ONE grid-wide barrier per timestep (cooperative
// launch: all CTAs are co-resident; the barrier spin is bounded and reports through the sticky device flag).
//
// Numerics: the recurrent product runs in fp32 on the un-rounded checkpoint weights (better than TF32); h_t is
// rounded to TF32 where it is produced because it is also an operand of the tensor-core GEMMs that follow (W_hh
// wgrad, the dense head), and so are the gate gradients the backward kernel writes (W_ih / W_hh wgrad, input dgrad).
#include "../../include/pe_b200.h"
#include "pe_common.cuh"

namespace pe {
namespace {

constexpr int LS_THREADS = 256;
constexpr int LS_ROWS = 32;          // rows of h_{t-1} staged per forward chunk
constexpr int LS_BROWS = 16;         // rows of dgates_t staged per backward chunk (each of the 8 warps takes two)

__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + expf(-x)); }

struct LstmSeq {
    // forward
    const float* gx;        // [S*N, 4H] input projection (no bias)
    const float* w_hh;      // [4H, H] fp32, checkpoint layout
    const float* b_ih;
    const float* b_hh;
    const float* h0;        // [N, H] or null
    const float* c0;
    float* h_all;           // [S*N, H]
    float* c_all;           // [S*N, H]
    float* act;             // [S*N, 4H] activated gates (null when no backward will follow)
    // backward
    const float* dh_all;    // [S*N, H] gradient w.r.t. every hidden output
    float* dg;              // [S*N, 4H] gate pre-activation gradients (written TF32-rounded when round_out)
    int S, N, H, U;         // U = hidden units per CTA
    int round_out;
    unsigned* bar;          // [2]: arrival count, generation
    int* error_flag;
};

// Grid-wide barrier (sense by generation).  All CTAs are co-resident (cooperative launch).  Bounded spin.
__device__ __forceinline__ bool grid_barrier(unsigned* bar, unsigned nblocks, int* error_flag) {
    __shared__ int s_ok;
    __syncthreads();
    if (threadIdx.x == 0) {
        int ok = 1;
        volatile unsigned* gen_p = bar + 1;
        const unsigned gen = *gen_p;
        __threadfence();
        const unsigned arrived = atomicAdd(bar, 1u);
        if (arrived == nblocks - 1) {
            bar[0] = 0u;
            __threadfence();
            atomicAdd(bar + 1, 1u);
        } else {
            const long long t0 = clock64();
            while (*gen_p == gen) {
                if (clock64() - t0 > PE_WAIT_LIMIT) {
                    atomicOr(error_flag, 64);
                    ok = 0;
                    break;
                }
            }
        }
        __threadfence();
        s_ok = ok;
    }
    __syncthreads();
    return s_ok != 0;
}

// Contiguous rows global (L2) -> shared memory, n4 float4s, with four independent loads in flight per thread (a plain
// load-store loop serialises on the L2 latency: 16 round trips per chunk).
__device__ __forceinline__ void stage_rows(float* dst, const float* src, int n4) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int base = threadIdx.x; base < n4; base += 4 * LS_THREADS) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = base + u * LS_THREADS;
            if (idx < n4) v[u] = __ldcg(s4 + idx);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = base + u * LS_THREADS;
            if (idx < n4) d4[idx] = v[u];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LS_THREADS, 1) lstm_seq_fwd_kernel(const LstmSeq p) {
    extern __shared__ __align__(16) float ls_smem[];
    const int H = p.H, U = p.U, G4 = 4 * U;           // gate rows held by this CTA
    const int ldw = H + 4;                            // padded row: conflict-free 128-bit reads across gate rows
    float* s_w = ls_smem;                             // [G4][ldw]
    float* s_h = s_w + G4 * ldw;                      // [LS_ROWS][H]
    float* s_g = s_h + LS_ROWS * H;                   // [LS_ROWS][G4] recurrent pre-activations of a chunk
    const int j0 = blockIdx.x * U;                    // first hidden unit of this CTA
    const int nu = min(U, H - j0);                    // units really owned (last CTA may own fewer)

    // W_hh slice: row q = gate * U + u  <->  W_hh[gate * H + j0 + u, :]
    for (int idx = threadIdx.x; idx < G4 * (H / 4); idx += LS_THREADS) {
        const int q = idx / (H / 4), k4 = idx % (H / 4);
        const int gate = q / U, u = q % U;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (u < nu) v = *reinterpret_cast<const float4*>(p.w_hh + (long long)(gate * H + j0 + u) * H + 4 * k4);
        *reinterpret_cast<float4*>(s_w + q * ldw + 4 * k4) = v;
    }
    __syncthreads();

    const int q = threadIdx.x % G4;                   // this thread's gate row ...
    const int slot = threadIdx.x / G4;                // ... and row slot; rows slot, slot + nslots, ...
    const int nslots = LS_THREADS / G4;
    bool ok = true;
    for (int t = 0; t < p.S && ok; ++t) {
        const float* h_prev = t == 0 ? p.h0 : p.h_all + (long long)(t - 1) * p.N * H;
        const float* c_prev = t == 0 ? p.c0 : p.c_all + (long long)(t - 1) * p.N * H;
        for (int r0 = 0; r0 < p.N; r0 += LS_ROWS) {
            const int nr = min(LS_ROWS, p.N - r0);
            if (h_prev) {
                // written by other CTAs during the previous timestep: read through L2, four loads in flight per thread
                stage_rows(s_h, h_prev + (long long)r0 * H, nr * (H / 4));
                __syncthreads();
                if (slot < nslots) {
                    for (int r = slot; r < nr; r += 2 * nslots) {
                        const int r2 = r + nslots;
                        const bool two = r2 < nr;
                        float a0 = 0.f, a1 = 0.f;
                        const float* wq = s_w + q * ldw;
                        const float* ha = s_h + r * H;
                        const float* hb = s_h + (two ? r2 : r) * H;
#pragma unroll 4
                        for (int k = 0; k < H; k += 4) {
                            const float4 w4 = *reinterpret_cast<const float4*>(wq + k);
                            const float4 x4 = *reinterpret_cast<const float4*>(ha + k);
                            const float4 y4 = *reinterpret_cast<const float4*>(hb + k);
                            a0 = fmaf(w4.x, x4.x, a0); a0 = fmaf(w4.y, x4.y, a0);
                            a0 = fmaf(w4.z, x4.z, a0); a0 = fmaf(w4.w, x4.w, a0);
                            a1 = fmaf(w4.x, y4.x, a1); a1 = fmaf(w4.y, y4.y, a1);
                            a1 = fmaf(w4.z, y4.z, a1); a1 = fmaf(w4.w, y4.w, a1);
                        }
                        s_g[r * G4 + q] = a0;
                        if (two) s_g[r2 * G4 + q] = a1;
                    }
                }
                __syncthreads();
            }
            // fused cell for (row, unit) pairs of this chunk
            for (int idx = threadIdx.x; idx < nr * nu; idx += LS_THREADS) {
                const int r = idx / nu, u = idx % nu;
                const int j = j0 + u;
                const long long row = (long long)t * p.N + r0 + r;
                float g4[4];
#pragma unroll
                for (int gate = 0; gate < 4; ++gate) {
                    const int col = gate * H + j;
                    float v = p.gx[row * 4 * H + col] + p.b_ih[col] + p.b_hh[col];
                    if (h_prev) v += s_g[r * G4 + gate * U + u];
                    g4[gate] = v;
                }
                const float gi = sigm(g4[0]), gf = sigm(g4[1]), gg = tanhf(g4[2]), go = sigm(g4[3]);
                const float cp = c_prev ? c_prev[(long long)(r0 + r) * H + j] : 0.f;
                const float c = gf * cp + gi * gg;
                const float h = go * tanhf(c);
                p.c_all[row * H + j] = c;
                p.h_all[row * H + j] = p.round_out ? round_tf32(h) : h;
                if (p.act) {
                    float* a = p.act + row * 4 * H;
                    a[j] = gi;
                    a[H + j] = gf;
                    a[2 * H + j] = gg;
                    a[3 * H + j] = go;
                }
            }
            __syncthreads();
        }
        // h_t is complete on every CTA before anyone reads it for step t + 1
        if (t + 1 < p.S) ok = grid_barrier(p.bar, gridDim.x, p.error_flag);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// backward through time.  CTA b owns hidden units j0..j0+U: their cell gradients (all rows) and the same U COLUMNS of
// dh_rec = dgates W_hh.  Per timestep: cell backward for the own units -> dgates_t slice to global -> grid barrier ->
// dh_rec[:, own units] from the complete dgates_t and the resident column slice of W_hh.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LS_THREADS, 1) lstm_seq_bwd_kernel(const LstmSeq p) {
    extern __shared__ __align__(16) float ls_smem[];
    const int H = p.H, U = p.U, G = 4 * H;
    float* s_wt = ls_smem;                            // [4H][U]: W_hh[r, j0 + u]
    float* s_dg = s_wt + G * U;                       // [LS_BROWS][4H] staged dgates rows
    float* s_dhr = s_dg + LS_BROWS * G;               // [N][U] recurrent gradient of the own units
    float* s_dc = s_dhr + p.N * U;                    // [N][U] cell-state gradient carried to step t - 1
    const int j0 = blockIdx.x * U;
    const int nu = min(U, H - j0);
    for (int idx = threadIdx.x; idx < G * U; idx += LS_THREADS) {
        const int r = idx / U, u = idx % U;
        s_wt[idx] = u < nu ? p.w_hh[(long long)r * H + j0 + u] : 0.f;
    }
    for (int idx = threadIdx.x; idx < p.N * U; idx += LS_THREADS) {
        s_dhr[idx] = 0.f;
        s_dc[idx] = 0.f;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    bool ok = true;
    for (int t = p.S - 1; t >= 0 && ok; --t) {
        const float* c_prev = t == 0 ? p.c0 : p.c_all + (long long)(t - 1) * p.N * H;
        // ---- cell backward (own units, every row) ----------------------------------------------------------
        for (int idx = threadIdx.x; idx < p.N * nu; idx += LS_THREADS) {
            const int n = idx / nu, u = idx % nu;
            const int j = j0 + u;
            const long long row = (long long)t * p.N + n;
            const float* a = p.act + row * G;
            const float gi = a[j], gf = a[H + j], gg = a[2 * H + j], go = a[3 * H + j];
            const float dht = p.dh_all[row * H + j] + s_dhr[n * U + u];
            const float tc = tanhf(p.c_all[row * H + j]);
            const float dc = dht * go * (1.f - tc * tc) + s_dc[n * U + u];
            const float cp = c_prev ? c_prev[(long long)n * H + j] : 0.f;
            float d0 = dc * gg * gi * (1.f - gi);
            float d1 = dc * cp * gf * (1.f - gf);
            float d2 = dc * gi * (1.f - gg * gg);
            float d3 = dht * tc * go * (1.f - go);
            if (p.round_out) {
                d0 = round_tf32(d0); d1 = round_tf32(d1); d2 = round_tf32(d2); d3 = round_tf32(d3);
            }
            float* dgr = p.dg + row * G;
            dgr[j] = d0;
            dgr[H + j] = d1;
            dgr[2 * H + j] = d2;
            dgr[3 * H + j] = d3;
            s_dc[n * U + u] = dc * gf;
        }
        if (t == 0) break;                             // no recurrent gradient is needed before the first step
        ok = grid_barrier(p.bar, gridDim.x, p.error_flag);
        if (!ok) break;
        // ---- dh_rec[:, own units] = dgates_t W_hh[:, own units] ---------------------------------------------
        const float* dgt = p.dg + (long long)t * p.N * G;
        if (U == 4 && (G & 255) == 0) {
            // a warp per row, straight from L2 (every CTA reads all of dgates_t; it was written this very step and
            // sits in L2): lane l owns the gate rows l + 32 i -- coalesced 128-byte loads of the row, conflict-free
            // 128-bit reads of the resident W_hh column slice (one float4 = the four own units of gate row r)
            const float4* wt4 = reinterpret_cast<const float4*>(s_wt);
            for (int n = warp; n < p.N; n += LS_THREADS / 32) {
                const float* dr = dgt + (long long)n * G;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int r0 = 0; r0 < G; r0 += 256) {
                    float d[8];
#pragma unroll
                    for (int v = 0; v < 8; ++v) d[v] = __ldcg(dr + r0 + 32 * v + lane);
#pragma unroll
                    for (int v = 0; v < 8; ++v) {
                        const float4 w = wt4[r0 + 32 * v + lane];
                        acc.x = fmaf(d[v], w.x, acc.x);
                        acc.y = fmaf(d[v], w.y, acc.y);
                        acc.z = fmaf(d[v], w.z, acc.z);
                        acc.w = fmaf(d[v], w.w, acc.w);
                    }
                }
                acc.x = warp_sum(acc.x); acc.y = warp_sum(acc.y); acc.z = warp_sum(acc.z); acc.w = warp_sum(acc.w);
                if (lane == 0) *reinterpret_cast<float4*>(s_dhr + n * 4) = acc;
            }
            __syncthreads();
            continue;
        }
        for (int n0 = 0; n0 < p.N; n0 += LS_BROWS) {
            const int nr = min(LS_BROWS, p.N - n0);
            stage_rows(s_dg, dgt + (long long)n0 * G, nr * (G / 4));
            __syncthreads();
            for (int rw = warp; rw < nr; rw += LS_THREADS / 32) {      // a warp per staged row, lanes split the 4H sum
                float acc[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) acc[u] = 0.f;
                const float* dr = s_dg + rw * G;
                for (int r = lane; r < G; r += 32) {
                    const float d = dr[r];
                    const float* w = s_wt + r * U;
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        if (u < U) acc[u] = fmaf(d, w[u], acc[u]);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (u < U) {
                        const float v = warp_sum(acc[u]);
                        if (lane == 0) s_dhr[(n0 + rw) * U + u] = v;
                    }
                }
            }
            __syncthreads();
        }
    }
}

unsigned* g_lstm_bar = nullptr;

}  // namespace

void lstm_seq_reset() {
    if (g_lstm_bar) cudaMemset(g_lstm_bar, 0, 2 * sizeof(unsigned));
}

}  // namespace pe

using namespace pe;

extern "C" {

// Units per CTA so that the grid fits the SMs; <= 8 (the backward kernel's register tile)
static int lstm_seq_units(int H) {
    int U = 4;
    while ((H + U - 1) / U > num_sms() && U < 8) ++U;
    return U;
}

static size_t lstm_fwd_smem(int H, int U) { return sizeof(float) * ((size_t)4 * U * (H + 4) + (size_t)LS_ROWS * H + (size_t)LS_ROWS * 4 * U); }
static size_t lstm_bwd_smem(int H, int U, int N) { return sizeof(float) * ((size_t)4 * H * U + (size_t)LS_BROWS * 4 * H + (size_t)2 * N * U); }

int pe_lstm_seq_supported(int N, int Hd, int backward) {
    if (Hd % 4 != 0 || N < 1) return 0;
    const int U = lstm_seq_units(Hd);
    if ((Hd + U - 1) / U > num_sms()) return 0;
    const size_t need = backward ? lstm_bwd_smem(Hd, U, N) : lstm_fwd_smem(Hd, U);
    return need <= 200 * 1024 ? 1 : 0;
}

static int lstm_seq_prepare(int** flag) {
    if (!g_lstm_bar) {
        PE_CHECK_CUDA(cudaMalloc(&g_lstm_bar, 2 * sizeof(unsigned)));
        PE_CHECK_CUDA(cudaMemset(g_lstm_bar, 0, 2 * sizeof(unsigned)));
    }
    *flag = device_error_flag();
    PE_REQUIRE(*flag != nullptr, "lstm_seq: no device error flag");
    return 0;
}

int pe_lstm_seq_fwd(const float* gx, const float* w_hh, const float* b_ih, const float* b_hh, const float* h0,
                    const float* c0, float* h_all, float* c_all, float* act, int S, int N, int Hd, int round_tf32,
                    void* stream) {
    PE_REQUIRE(pe_lstm_seq_supported(N, Hd, 0), "lstm_seq_fwd: shape N=%d H=%d not supported", N, Hd);
    PE_REQUIRE((h0 == nullptr) == (c0 == nullptr), "lstm_seq_fwd: h0 and c0 come together");
    int* flag = nullptr;
    if (lstm_seq_prepare(&flag)) return 2;
    LstmSeq p = {};
    p.gx = gx; p.w_hh = w_hh; p.b_ih = b_ih; p.b_hh = b_hh; p.h0 = h0; p.c0 = c0;
    p.h_all = h_all; p.c_all = c_all; p.act = act;
    p.S = S; p.N = N; p.H = Hd; p.U = lstm_seq_units(Hd); p.round_out = round_tf32;
    p.bar = g_lstm_bar; p.error_flag = flag;
    const size_t smem = lstm_fwd_smem(Hd, p.U);
    static size_t configured = 0;
    if (smem > configured) {
        PE_CHECK_CUDA(cudaFuncSetAttribute(lstm_seq_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    void* args[] = {&p};
    const int grid = (Hd + p.U - 1) / p.U;
    PE_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)lstm_seq_fwd_kernel, dim3(grid), dim3(LS_THREADS), args, smem,
                                              (cudaStream_t)stream));
    PE_LAUNCH_CHECK();
    return 0;
}

int pe_lstm_seq_bwd(const float* dh_all, const float* w_hh, const float* act, const float* c_all, const float* c0,
                    float* dgates, int S, int N, int Hd, int round_tf32, void* stream) {
    PE_REQUIRE(pe_lstm_seq_supported(N, Hd, 1), "lstm_seq_bwd: shape N=%d H=%d not supported", N, Hd);
    int* flag = nullptr;
    if (lstm_seq_prepare(&flag)) return 2;
    LstmSeq p = {};
    p.dh_all = dh_all; p.w_hh = w_hh; p.act = const_cast<float*>(act); p.c_all = const_cast<float*>(c_all); p.c0 = c0;
    p.dg = dgates;
    p.S = S; p.N = N; p.H = Hd; p.U = lstm_seq_units(Hd); p.round_out = round_tf32;
    p.bar = g_lstm_bar; p.error_flag = flag;
    const size_t smem = lstm_bwd_smem(Hd, p.U, N);
    static size_t configured = 0;
    if (smem > configured) {
        PE_CHECK_CUDA(cudaFuncSetAttribute(lstm_seq_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    void* args[] = {&p};
    const int grid = (Hd + p.U - 1) / p.U;
    PE_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)lstm_seq_bwd_kernel, dim3(grid), dim3(LS_THREADS), args, smem,
                                              (cudaStream_t)stream));
    PE_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"

// Fused fusion-head step for rollout inference (north_star kernel 2): feature concat + first dense layer /
// LSTM gate projection over the WHOLE grid, then LSTM cell + the remaining small layers by the last CTA to
// arrive -- one launch, no grid-wide wait.
//
//   phase A (every CTA): the fusion rows [latent | aux | proprio] of the <= 8 frames are staged in shared memory
//       (the 7 proprioceptive values are injected there: the reference's torch.cat never materialises), together
//       with the previous hidden state; each warp owns output neurons j, streams W[j][:] once from HBM with
//       coalesced loads and produces all rows' dot products with warp-shuffle reductions.
//   tail (last CTA, found with a fence + atomic ticket): LSTM cell (gates i,f,g,o), then up to three small dense
//       layers with their activations, the optional measurement difference (pre_out - x0bar) for the two-headed
//       models, and the ticket reset for the next launch / graph replay.
//
// fp32 FMA throughout (weights are read in the checkpoint layout, no packing): the step is bound by streaming
// 15-35 MB of weights once, so tensor cores would not help a <= 8-row GEMV.
#include "../../include/pe_b200.h"
#include "pe_common.cuh"

namespace pe {
namespace {

constexpr int FH_THREADS = 1024;
constexpr int FH_WARPS = FH_THREADS / 32;
constexpr int FH_MAX_ROWS = 8;

__device__ __forceinline__ float fh_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

// acc[n] += sum_k w[k] * xs[n][k], k strided over the lanes; xs rows are `ld` floats apart in shared memory
template <int NR>
__device__ __forceinline__ void fh_dot(const float* __restrict__ w, int K, const float* xs, int ld, int lane,
                                       float (&acc)[NR]) {
    int k = lane;
    for (; k + 96 < K; k += 128) {           // four independent 128-byte warp loads in flight
        const float w0 = __ldg(w + k), w1 = __ldg(w + k + 32), w2 = __ldg(w + k + 64), w3 = __ldg(w + k + 96);
#pragma unroll
        for (int n = 0; n < NR; ++n) {
            const float* xr = xs + n * ld + k;
            acc[n] = fmaf(w0, xr[0], acc[n]);
            acc[n] = fmaf(w1, xr[32], acc[n]);
            acc[n] = fmaf(w2, xr[64], acc[n]);
            acc[n] = fmaf(w3, xr[96], acc[n]);
        }
    }
    for (; k < K; k += 32) {
        const float w0 = __ldg(w + k);
#pragma unroll
        for (int n = 0; n < NR; ++n) acc[n] = fmaf(w0, xs[n * ld + k], acc[n]);
    }
}

template <int NR>
__global__ void __launch_bounds__(FH_THREADS, 1) fused_head_kernel(const pe_head_desc d) {
    extern __shared__ __align__(16) float fh_smem[];
    __shared__ int s_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ldxs = d.k_x + d.k_h;                      // one staged row: [x | h_prev]
    float* xs = fh_smem;

    // ---- stage the fusion rows (+ proprio injection) and the previous hidden state ---------------------
    for (int n = 0; n < NR; ++n) {
        float* row = xs + n * ldxs;
        if (n < d.n_rows) {
            const float* src = d.x + (long long)n * d.ldx;
            for (int k = tid; k < d.k_x; k += FH_THREADS) {
                float v = src[k];
                if (d.inj && k >= d.inj_col && k < d.inj_col + 7) v = d.inj[n * d.ld_inj + (k - d.inj_col)];
                row[k] = v;
            }
            if (d.k_h > 0) {
                for (int k = tid; k < d.k_h; k += FH_THREADS)
                    row[d.k_x + k] = d.h_prev ? d.h_prev[n * d.k_h + k] : 0.f;
            }
        } else {
            for (int k = tid; k < ldxs; k += FH_THREADS) row[k] = 0.f;
        }
    }
    __syncthreads();

    // ---- phase A: one output neuron per warp task --------------------------------------------------------
    for (int j = blockIdx.x * FH_WARPS + warp; j < d.j_a; j += gridDim.x * FH_WARPS) {
        float acc[NR];
#pragma unroll
        for (int n = 0; n < NR; ++n) acc[n] = 0.f;
        fh_dot<NR>(d.w_x + (long long)j * d.k_x, d.k_x, xs, ldxs, lane, acc);
        if (d.k_h > 0 && d.h_prev) fh_dot<NR>(d.w_h + (long long)j * d.k_h, d.k_h, xs + d.k_x, ldxs, lane, acc);
#pragma unroll
        for (int n = 0; n < NR; ++n) acc[n] = warp_sum(acc[n]);
        if (lane == 0) {
            float b = 0.f;
            if (d.b1) b += d.b1[j];
            if (d.b2) b += d.b2[j];
#pragma unroll
            for (int n = 0; n < NR; ++n) {
                if (n < d.n_rows) {
                    float v = acc[n] + b;
                    if (d.relu_a) v = fmaxf(v, 0.f);
                    d.out_a[n * d.j_a + j] = v;
                }
            }
        }
    }

    // ---- ticket: the last CTA to finish phase A runs the tail -------------------------------------------
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned int t = atomicAdd(d.counter, 1u);
        s_last = (t == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();

    // ---- tail ------------------------------------------------------------------------------------------------
    // two ping-pong buffers of n_rows x width floats reuse the staging area (phase A is over for this CTA)
    int width = d.lstm_hidden > 0 ? d.lstm_hidden : d.j_a;
    for (int t = 0; t < d.n_tail; ++t) width = max(width, d.tail_j[t]);
    float* buf0 = fh_smem;
    float* buf1 = fh_smem + NR * width;
    int K = 0;
    if (d.lstm_hidden > 0) {
        const int Hd = d.lstm_hidden;
        for (int idx = tid; idx < d.n_rows * Hd; idx += FH_THREADS) {
            const int n = idx / Hd, u = idx - n * Hd;
            const float* g = d.out_a + n * d.j_a;                 // written by other SMs: read through L2
            const float gi = __ldcg(g + u), gf = __ldcg(g + Hd + u), gg = __ldcg(g + 2 * Hd + u),
                        go = __ldcg(g + 3 * Hd + u);
            const float i = fh_sigmoid(gi), f = fh_sigmoid(gf), gt = tanhf(gg), o = fh_sigmoid(go);
            const float cp = d.c_prev ? d.c_prev[idx] : 0.f;
            const float c = f * cp + i * gt;
            const float h = o * tanhf(c);
            d.c_out[idx] = c;
            d.h_out[idx] = h;
            buf0[n * width + u] = h;
        }
        K = Hd;
    } else {
        for (int idx = tid; idx < d.n_rows * d.j_a; idx += FH_THREADS) {
            const int n = idx / d.j_a, u = idx - n * d.j_a;
            buf0[n * width + u] = __ldcg(d.out_a + idx);
        }
        K = d.j_a;
    }
    for (int idx = tid; idx < (NR - d.n_rows) * width; idx += FH_THREADS) buf0[d.n_rows * width + idx] = 0.f;
    __syncthreads();
    float* in = buf0;
    float* outb = buf1;
    for (int t = 0; t < d.n_tail; ++t) {
        const int J = d.tail_j[t];
        const bool last = t == d.n_tail - 1;
        for (int j = warp; j < J; j += FH_WARPS) {
            float acc[NR];
#pragma unroll
            for (int n = 0; n < NR; ++n) acc[n] = 0.f;
            fh_dot<NR>(d.tail_w[t] + (long long)j * K, K, in, width, lane, acc);
#pragma unroll
            for (int n = 0; n < NR; ++n) acc[n] = warp_sum(acc[n]);
            if (lane == 0) {
                const float b = d.tail_b[t] ? d.tail_b[t][j] : 0.f;
#pragma unroll
                for (int n = 0; n < NR; ++n) {
                    float v = acc[n] + b;
                    if (d.tail_relu[t]) v = fmaxf(v, 0.f);
                    outb[n * width + j] = v;
                    if (last && n < d.n_rows) {
                        d.out[n * d.ld_out + j] = v;
                        if (d.diff) d.diff[n * d.ld_diff + d.diff_col + j] = v - d.meas[n * d.ld_meas + j];
                    }
                }
            }
        }
        __syncthreads();
        float* tmp = in;
        in = outb;
        outb = tmp;
        K = J;
    }
    if (d.n_tail == 0) {
        // no tail layers: phase A already produced the final rows
        for (int idx = tid; idx < d.n_rows * d.j_a; idx += FH_THREADS) {
            const int n = idx / d.j_a, u = idx - n * d.j_a;
            const float v = in[n * width + u];
            d.out[n * d.ld_out + u] = v;
            if (d.diff) d.diff[n * d.ld_diff + d.diff_col + u] = v - d.meas[n * d.ld_meas + u];
        }
    }
    if (tid == 0) *d.counter = 0u;      // ready for the next launch / graph replay
}

}  // namespace
}  // namespace pe

using namespace pe;

extern "C" int pe_head_desc_size(void) { return (int)sizeof(pe_head_desc); }

extern "C" int pe_fused_head(const pe_head_desc* desc, void* stream) {
    PE_REQUIRE(desc != nullptr, "fused_head: null descriptor");
    const pe_head_desc& d = *desc;
    PE_REQUIRE(d.n_rows >= 1 && d.n_rows <= FH_MAX_ROWS, "fused_head: 1..%d rows supported (got %d)", FH_MAX_ROWS,
               d.n_rows);
    PE_REQUIRE(d.x && d.w_x && d.out_a && d.counter && d.out, "fused_head: missing pointer");
    PE_REQUIRE(d.n_tail >= 0 && d.n_tail <= 3, "fused_head: at most three tail layers");
    PE_REQUIRE(d.lstm_hidden == 0 || (d.j_a == 4 * d.lstm_hidden && d.c_out && d.h_out),
               "fused_head: LSTM tail needs j_a == 4 * hidden and state outputs");
    PE_REQUIRE(!d.inj || (d.inj_col >= 0 && d.inj_col + 7 <= d.k_x), "fused_head: injected columns out of range");
    PE_REQUIRE(!d.diff || d.meas, "fused_head: diff output needs the measurement");
    const int nr = d.n_rows <= 1 ? 1 : (d.n_rows <= 2 ? 2 : (d.n_rows <= 4 ? 4 : 8));
    int width = d.lstm_hidden > 0 ? d.lstm_hidden : d.j_a;
    for (int t = 0; t < d.n_tail; ++t) width = width > d.tail_j[t] ? width : d.tail_j[t];
    size_t smem = sizeof(float) * (size_t)nr * (d.k_x + d.k_h);
    const size_t tail = sizeof(float) * 2 * (size_t)nr * width;
    if (tail > smem) smem = tail;
    PE_REQUIRE(smem <= 220 * 1024, "fused_head: %zu bytes of staging do not fit in shared memory", smem);
    // one CTA per SM at most; never more CTAs than there are warp tasks
    int grid = (d.j_a + FH_WARPS - 1) / FH_WARPS;
    if (grid > num_sms()) grid = num_sms();
    if (grid < 1) grid = 1;
    cudaStream_t st = (cudaStream_t)stream;
#define PE_FH_LAUNCH(NR)                                                                                          \
    do {                                                                                                          \
        static size_t configured = 0;                                                                             \
        if (smem > 48 * 1024 && smem > configured) {                                                              \
            PE_CHECK_CUDA(cudaFuncSetAttribute(fused_head_kernel<NR>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                               220 * 1024));                                                      \
            configured = 220 * 1024;                                                                              \
        }                                                                                                         \
        fused_head_kernel<NR><<<grid, FH_THREADS, smem, st>>>(d);                                                 \
    } while (0)
    switch (nr) {
        case 1: PE_FH_LAUNCH(1); break;
        case 2: PE_FH_LAUNCH(2); break;
        case 4: PE_FH_LAUNCH(4); break;
        default: PE_FH_LAUNCH(8); break;
    }
#undef PE_FH_LAUNCH
    PE_LAUNCH_CHECK();
    return 0;
}

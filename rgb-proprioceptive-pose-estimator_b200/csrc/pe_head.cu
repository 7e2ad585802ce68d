// Head-side kernels: LSTM cell (forward / backward pointwise part), pose loss with gradient,
// fused Adam / SGD over flat parameter arenas.
#include "../../include/pe_b200.h"
#include "pe_common.cuh"

namespace pe {
namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// ---------------------------------------------------------------------------------------------
// LSTM cell, gate order (i, f, g, o) along the 4H axis  (torch.nn.LSTM convention)
// ---------------------------------------------------------------------------------------------
__global__ void lstm_cell_fwd_kernel(const float* __restrict__ gx, int ldgx, const float* __restrict__ gh, int ldgh,
                                     const float* __restrict__ b_ih, const float* __restrict__ b_hh,
                                     const float* __restrict__ c_prev, float* __restrict__ c_out,
                                     float* __restrict__ h_out, int ldh, float* __restrict__ act, int N, int Hd,
                                     int round_out) {
    pdl_sync();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * Hd) return;
    const int j = idx % Hd, n = idx / Hd;
    float g4[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int col = k * Hd + j;
        float v = gx[(long long)n * ldgx + col] + b_ih[col] + b_hh[col];
        if (gh) v += gh[(long long)n * ldgh + col];
        g4[k] = v;
    }
    const float i = sigmoidf_(g4[0]), f = sigmoidf_(g4[1]), g = tanhf(g4[2]), o = sigmoidf_(g4[3]);
    const float cp = c_prev ? c_prev[idx] : 0.f;
    const float c = f * cp + i * g;
    const float h = o * tanhf(c);
    c_out[idx] = c;
    h_out[(long long)n * ldh + j] = round_out ? round_tf32(h) : h;
    if (act) {
        float* a = act + (long long)n * 4 * Hd;
        a[j] = i;
        a[Hd + j] = f;
        a[2 * Hd + j] = g;
        a[3 * Hd + j] = o;
    }
}

__global__ void lstm_cell_bwd_kernel(const float* __restrict__ dh, int lddh, const float* __restrict__ dh_rec,
                                     const float* __restrict__ dc_next, const float* __restrict__ act,
                                     const float* __restrict__ c_prev, const float* __restrict__ c_out,
                                     float* __restrict__ dgates, int lddg, float* __restrict__ dc_prev, int N,
                                     int Hd) {
    pdl_sync();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * Hd) return;
    const int j = idx % Hd, n = idx / Hd;
    const float* a = act + (long long)n * 4 * Hd;
    const float i = a[j], f = a[Hd + j], g = a[2 * Hd + j], o = a[3 * Hd + j];
    float dht = dh ? dh[(long long)n * lddh + j] : 0.f;
    if (dh_rec) dht += dh_rec[idx];
    const float tc = tanhf(c_out[idx]);
    float dc = dht * o * (1.f - tc * tc);
    if (dc_next) dc += dc_next[idx];
    const float cp = c_prev ? c_prev[idx] : 0.f;
    float* dg = dgates + (long long)n * lddg;
    dg[j] = dc * g * i * (1.f - i);
    dg[Hd + j] = dc * cp * f * (1.f - f);
    dg[2 * Hd + j] = dc * i * (1.f - g * g);
    dg[3 * Hd + j] = dht * tc * o * (1.f - o);
    dc_prev[idx] = dc * f;
}

// ---------------------------------------------------------------------------------------------
// pose loss: rows strided over the threads of ONE block -- or, from 4096 rows on, of a thread-block cluster of 8 CTAs
// whose partial sums CTA 0 collects through distributed shared memory in rank order; warp-shuffle + smem reduction, no
// atomics, so the result is deterministic for a given row count
// ---------------------------------------------------------------------------------------------
constexpr int POSE_LOSS_CLUSTER = 8;
constexpr long long POSE_LOSS_CLUSTER_ROWS = 4096;

__global__ void __launch_bounds__(256)
pose_loss_kernel(const float* __restrict__ pred, int ldp, const float* __restrict__ truth, int ldt, long long n,
                 int metric, int mode, float alpha, float epsilon, float scale, float* __restrict__ loss,
                 float* __restrict__ dpred, int lddp, float* __restrict__ val) {
    pdl_sync();
    __shared__ float red[3][8];
    __shared__ float part[4];                 // this CTA's partial sums (read by CTA 0 of the cluster)
    const unsigned nctas = gridDim.x;         // 1, or the cluster size (the grid is exactly one cluster)
    float acc_loss = 0.f, acc_pos = 0.f, acc_ang = 0.f;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)nctas * blockDim.x) {
        const float* p = pred + r * ldp;
        const float* t = truth + r * ldt;
        const float d0 = p[0] - t[0], d1 = p[1] - t[1], d2 = p[2] - t[2];
        const float a0 = fabsf(d0), a1 = fabsf(d1), a2 = fabsf(d2);
        float pos = 0.f, g0 = 0.f, g1 = 0.f, g2 = 0.f;
        const float l2 = sqrtf(d0 * d0 + d1 * d1 + d2 * d2 + epsilon);
        if (metric == 1 || metric == 3) {
            pos += l2;
            g0 += d0 / l2;
            g1 += d1 / l2;
            g2 += d2 / l2;
        }
        if (metric == 0 || metric == 3) {
            pos += a0 + a1 + a2;
            g0 += (d0 > 0.f) - (d0 < 0.f);
            g1 += (d1 > 0.f) - (d1 < 0.f);
            g2 += (d2 > 0.f) - (d2 < 0.f);
        }
        if (metric == 2 || metric == 3) {
            // first maximum wins the gradient (torch.max(dim) convention)
            int k = 0;
            float m = a0;
            if (a1 > m) { m = a1; k = 1; }
            if (a2 > m) { m = a2; k = 2; }
            pos += m;
            if (k == 0) g0 += (d0 > 0.f) - (d0 < 0.f);
            if (k == 1) g1 += (d1 > 0.f) - (d1 < 0.f);
            if (k == 2) g2 += (d2 > 0.f) - (d2 < 0.f);
        }
        const float q0 = p[3], q1 = p[4], q2 = p[5], q3 = p[6];
        const float mag = sqrtf(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
        const float h0 = q0 / mag, h1 = q1 / mag, h2 = q2 / mag, h3 = q3 / mag;
        const float ip = h0 * t[3] + h1 * t[4] + h2 * t[5] + h3 * t[6];
        float ori = 0.f, e0 = 0.f, e1 = 0.f, e2 = 0.f, e3 = 0.f;
        if (mode == 1) {
            const float pen = fmaxf(-h3, 0.f);
            ori = (1.f - ip * ip) + pen;
            // gradient w.r.t. the normalised quaternion ...
            float u0 = -2.f * ip * t[3], u1 = -2.f * ip * t[4], u2 = -2.f * ip * t[5], u3 = -2.f * ip * t[6];
            if (-h3 >= 0.f) u3 -= 1.f;
            // ... pulled back through q / |q|
            const float dot = u0 * h0 + u1 * h1 + u2 * h2 + u3 * h3;
            e0 = (u0 - h0 * dot) / mag;
            e1 = (u1 - h1 * dot) / mag;
            e2 = (u2 - h2 * dot) / mag;
            e3 = (u3 - h3 * dot) / mag;
        }
        acc_loss += pos + alpha * ori;
        if (dpred) {
            float* g = dpred + r * lddp;
            g[0] = scale * g0;
            g[1] = scale * g1;
            g[2] = scale * g2;
            g[3] = scale * alpha * e0;
            g[4] = scale * alpha * e1;
            g[5] = scale * alpha * e2;
            g[6] = scale * alpha * e3;
        }
        if (val) {
            acc_pos += pos;
            const float w = fminf(fmaxf(ip, -1.f), 1.f);
            float ang = 2.f * acosf(w);
            if (ang > 3.14159265358979323846f) ang -= 6.28318530717958647692f;
            // robosuite quat2axisangle returns angle 0 when sqrt(1 - w^2) is ~0
            if (sqrtf(fmaxf(1.f - w * w, 0.f)) < 1e-9f) ang = 0.f;
            acc_ang += fabsf(ang);
        }
    }
    acc_loss = warp_sum(acc_loss);
    acc_pos = warp_sum(acc_pos);
    acc_ang = warp_sum(acc_ang);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        red[0][warp] = acc_loss;
        red[1][warp] = acc_pos;
        red[2][warp] = acc_ang;
    }
    __syncthreads();
    float a = 0.f, b = 0.f, c = 0.f;
    if (threadIdx.x == 0) {
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
            a += red[0][w];
            b += red[1][w];
            c += red[2][w];
        }
        part[0] = a;
        part[1] = b;
        part[2] = c;
    }
    if (nctas > 1) {
        // partial sums of the other CTAs through distributed shared memory, summed in rank order
        cluster_sync_all();
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            for (unsigned rk = 1; rk < nctas; ++rk) {
                uint32_t remote;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(part)), "r"(rk));
                float pa, pb, pc;
                asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(pa) : "r"(remote) : "memory");
                asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(pb) : "r"(remote + 4) : "memory");
                asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(pc) : "r"(remote + 8) : "memory");
                a += pa;
                b += pb;
                c += pc;
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (loss) loss[0] = scale * a;
        if (val) {
            val[0] = b;
            val[1] = c;
        }
    }
    if (nctas > 1) cluster_sync_all();        // nobody's shared memory disappears while CTA 0 is still reading it
}

// ---------------------------------------------------------------------------------------------
// optimizers: 128-bit streaming over flat arenas (28 B/param Adam, 12-20 B/param SGD)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float lr_c, float b1, float b2,
                                         float eps, float wd, float inv_bc2_sqrt, float gs) {
    g *= gs;
    if (wd != 0.f) g = fmaf(wd, p, g);
    m = m + (g - m) * (1.f - b1);              // exp_avg.lerp_(grad, 1 - beta1)
    v = v * b2 + (1.f - b2) * g * g;           // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1 - beta2)
    const float denom = sqrtf(v) * inv_bc2_sqrt + eps;
    p = p - lr_c * (m / denom);                // param.addcdiv_(exp_avg, denom, -lr / bias_correction1)
}

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            long long n, float lr_c, float b1, float b2, float eps, float wd, float inv_bc2_sqrt, float gs) {
    pdl_sync();
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 pp = reinterpret_cast<float4*>(p)[i];
        const float4 gg = reinterpret_cast<const float4*>(g)[i];
        float4 mm = reinterpret_cast<float4*>(m)[i];
        float4 vv = reinterpret_cast<float4*>(v)[i];
        adam_one(pp.x, gg.x, mm.x, vv.x, lr_c, b1, b2, eps, wd, inv_bc2_sqrt, gs);
        adam_one(pp.y, gg.y, mm.y, vv.y, lr_c, b1, b2, eps, wd, inv_bc2_sqrt, gs);
        adam_one(pp.z, gg.z, mm.z, vv.z, lr_c, b1, b2, eps, wd, inv_bc2_sqrt, gs);
        adam_one(pp.w, gg.w, mm.w, vv.w, lr_c, b1, b2, eps, wd, inv_bc2_sqrt, gs);
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
    }
    // tail
    const long long t = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) adam_one(p[t], g[t], m[t], v[t], lr_c, b1, b2, eps, wd, inv_bc2_sqrt, gs);
}

__device__ __forceinline__ void sgd_one(float& p, float g, float* mom, float lr, float momentum, float wd, int first,
                                        float gs) {
    float gi = g * gs;
    if (wd != 0.f) gi = fmaf(wd, p, gi);
    if (mom) {
        const float b = first ? gi : fmaf(momentum, *mom, gi);
        *mom = b;
        gi = b;
    }
    p -= lr * gi;
}

// torch.optim.SGD semantics (momentum buffer = first gradient on the first step, dampening 0, no nesterov), streamed
// with 128-bit loads / stores like the Adam kernel: 12 B per parameter without momentum, 20 B with
__global__ void __launch_bounds__(256)
sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ mom, long long n, float lr,
           float momentum, float wd, int first, float gs) {
    pdl_sync();
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 pp = reinterpret_cast<float4*>(p)[i];
        const float4 gg = reinterpret_cast<const float4*>(g)[i];
        float4 mm = mom ? reinterpret_cast<float4*>(mom)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        sgd_one(pp.x, gg.x, mom ? &mm.x : nullptr, lr, momentum, wd, first, gs);
        sgd_one(pp.y, gg.y, mom ? &mm.y : nullptr, lr, momentum, wd, first, gs);
        sgd_one(pp.z, gg.z, mom ? &mm.z : nullptr, lr, momentum, wd, first, gs);
        sgd_one(pp.w, gg.w, mom ? &mm.w : nullptr, lr, momentum, wd, first, gs);
        reinterpret_cast<float4*>(p)[i] = pp;
        if (mom) reinterpret_cast<float4*>(mom)[i] = mm;
    }
    const long long t = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x;     // tail
    if (t < n) sgd_one(p[t], g[t], mom ? mom + t : nullptr, lr, momentum, wd, first, gs);
}

}  // namespace
}  // namespace pe

using namespace pe;

extern "C" {

int pe_lstm_cell_fwd(const float* gx, int ldgx, const float* gh, int ldgh, const float* b_ih, const float* b_hh,
                     const float* c_prev, float* c_out, float* h_out, int ldh, float* act, int N, int Hd,
                     int round_tf32, void* stream) {
    const int n = N * Hd;
    if (n == 0) return 0;
    PE_LAUNCH(lstm_cell_fwd_kernel, (n + 255) / 256, 256, 0, gx, ldgx, gh, ldgh, b_ih, b_hh, c_prev, c_out, h_out,
              ldh, act, N, Hd, round_tf32);
    PE_LAUNCH_CHECK();
    return 0;
}

int pe_lstm_cell_bwd(const float* dh, int lddh, const float* dh_rec, const float* dc_next, const float* act,
                     const float* c_prev, const float* c_out, float* dgates, int lddg, float* dc_prev, int N, int Hd,
                     void* stream) {
    const int n = N * Hd;
    if (n == 0) return 0;
    PE_LAUNCH(lstm_cell_bwd_kernel, (n + 255) / 256, 256, 0, dh, lddh, dh_rec, dc_next, act, c_prev, c_out, dgates,
              lddg, dc_prev, N, Hd);
    PE_LAUNCH_CHECK();
    return 0;
}

int pe_pose_loss(const float* pred, int ldp, const float* truth, int ldt, long long n, int metric, int mode,
                 float alpha, float epsilon, float scale, float* loss, float* dpred, int lddp, float* val,
                 void* stream) {
    PE_REQUIRE(metric >= 0 && metric <= 3, "pose_loss: metric %d invalid", metric);
    PE_REQUIRE(mode == 0 || mode == 1, "pose_loss: mode %d invalid", mode);
    if (n >= POSE_LOSS_CLUSTER_ROWS) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(POSE_LOSS_CLUSTER);
        cfg.blockDim = dim3(256);
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = POSE_LOSS_CLUSTER;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        PE_CHECK_CUDA(cudaLaunchKernelEx(&cfg, pose_loss_kernel, pred, ldp, truth, ldt, n, metric, mode, alpha, epsilon,
                                         scale, loss, dpred, lddp, val));
    } else {
        PE_LAUNCH(pose_loss_kernel, 1, 256, 0, pred, ldp, truth, ldt, n, metric, mode, alpha, epsilon, scale, loss, dpred,
                  lddp, val);
    }
    PE_LAUNCH_CHECK();
    return 0;
}

int pe_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                 float eps, float weight_decay, int step, float grad_scale, void* stream) {
    PE_REQUIRE(step >= 1, "adam: step must be >= 1");
    PE_REQUIRE(((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) % 16 == 0, "adam: arenas must be 16-byte aligned");
    if (n == 0) return 0;
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    const float lr_c = (float)((double)lr / bc1);
    const float inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
    long long blocks = (n / 4 + 255) / 256;
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    PE_LAUNCH(adam_kernel, (unsigned)blocks, 256, 0, p, g, m, v, n, lr_c, beta1, beta2, eps, weight_decay,
              inv_bc2_sqrt, grad_scale);
    PE_LAUNCH_CHECK();
    return 0;
}

int pe_sgd_step(float* p, const float* g, float* mom, long long n, float lr, float momentum, float weight_decay,
                int first_step, float grad_scale, void* stream) {
    if (n == 0) return 0;
    PE_REQUIRE(((uintptr_t)p | (uintptr_t)g | (uintptr_t)mom) % 16 == 0, "sgd: arenas must be 16-byte aligned");
    long long blocks = (n / 4 + 255) / 256;
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    PE_LAUNCH(sgd_kernel, (unsigned)blocks, 256, 0, p, g, mom, n, lr, momentum, weight_decay, first_step, grad_scale);
    PE_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"

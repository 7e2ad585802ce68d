// HBM-streaming kernels of the trunk: BatchNorm statistics / apply / backward, stem pooling,
// auxiliary BN1 branch, weight (un)packing, im2col for the 7x7 stem, small copy helpers.
// All activations are NHWC fp32; every kernel moves 128-bit words and sizes its grid from the SM count.
#include "../../include/pe_b200.h"
#include "pe_common.cuh"

namespace pe {
namespace {

constexpr int EW_THREADS = 256;

inline int grid_for(long long work_items, int per_block, int max_waves = 8) {
    long long blocks = (work_items + per_block - 1) / per_block;
    long long cap = (long long)num_sms() * max_waves;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

// Grid of a grid-stride streaming kernel: exactly one wave -- (SM count) x (CTAs of this kernel resident per SM,
// from the occupancy calculator for its register / shared-memory footprint) -- or fewer when the work is small.
// 1184 CTAs of a kernel that fits 6 per SM would run as 1.33 waves with a thin tail.
template <typename K>
inline int one_wave_grid(K kernel, size_t smem, long long work_items, int per_block) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, EW_THREADS, smem) != cudaSuccess || occ < 1) occ = 4;
    long long blocks = (work_items + per_block - 1) / per_block;
    const long long cap = (long long)num_sms() * occ;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 ld4_stream(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ float4 round4(float4 v) {
    return make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
}

// ---------------------------------------------------------------------------------------------
// ReLU / join masks as bits: float4 index i of a dense [P][C/4] tensor owns bit (i & 31) of the four words
// maskbits[(i >> 5) * 4 + comp] (one word per float4 component), so a warp that handles 32 consecutive
// float4s writes its mask with four ballots and reads it back with one broadcast 128-bit load.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ld_maskbits(const unsigned* __restrict__ bits, long long i4) {
    return __ldg(reinterpret_cast<const uint4*>(bits) + (i4 >> 5));
}
__device__ __forceinline__ float4 apply_maskbits(float4 d, uint4 w, long long i4) {
    const int sh = (int)(i4 & 31);
    d.x = ((w.x >> sh) & 1u) ? d.x : 0.f;
    d.y = ((w.y >> sh) & 1u) ? d.y : 0.f;
    d.z = ((w.z >> sh) & 1u) ? d.z : 0.f;
    d.w = ((w.w >> sh) & 1u) ? d.w : 0.f;
    return d;
}

// ---------------------------------------------------------------------------------------------
// per-channel reductions over [P][C]:  acc0 += f0(row, c), acc1 += f1(row, c)
// block = 256 threads = TR row lanes x GW channel groups (4 channels each); grid.y covers C > 1024.
// Rows are consumed four at a time so every thread keeps 8-16 independent 128-bit loads in flight.
// ---------------------------------------------------------------------------------------------
template <int KIND>  // 0: (y, y*y)   1: (g, g*xhat) with g = (dout [+ dout2]) * mask
__global__ void __launch_bounds__(EW_THREADS)
channel_reduce_kernel(const float* __restrict__ a, const float* __restrict__ a2, const float* __restrict__ b,
                      const float* __restrict__ c, const float* __restrict__ mean,
                      const float* __restrict__ invstd, const float* __restrict__ msc,
                      const float* __restrict__ msh, const unsigned* __restrict__ maskbits,
                      double* __restrict__ sums, long long P, int C, int relu) {
    pdl_sync();
    __shared__ float sm0[EW_THREADS * 4];
    __shared__ float sm1[EW_THREADS * 4];
    constexpr int U = 4;
    const int G = C >> 2;
    const int GW = G < EW_THREADS ? G : EW_THREADS;
    const int TR = EW_THREADS / GW;
    const int g = blockIdx.y * GW + (threadIdx.x % GW);
    const int r = threadIdx.x / GW;
    const long long rows_per_block = (P + gridDim.x - 1) / gridDim.x;
    const long long row0 = (long long)blockIdx.x * rows_per_block;
    long long row1 = row0 + rows_per_block;
    if (row1 > P) row1 = P;

    float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
    float4 mu = s0, is = s0, ksc = s0, ksh = s0;
    if (KIND == 1) {
        mu = ld4(mean + 4 * g);
        is = ld4(invstd + 4 * g);
        if (relu && !b) {   // ReLU mask recomputed from y (no residual): out > 0 <=> y*scale + shift > 0
            ksc = ld4(msc + 4 * g);
            ksh = ld4(msh + 4 * g);
        }
    }
    const bool relu_from_out = KIND == 1 && relu && b != nullptr;
    const bool relu_from_y = KIND == 1 && relu && b == nullptr;
    for (long long row = row0 + r; row < row1; row += (long long)U * TR) {
        float4 d[U], e[U], y[U], o[U];
        uint4 mb[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long rr = row + (long long)u * TR;
            if (rr < row1) {
                const long long off = rr * C + 4 * g;
                if (KIND == 0) {
                    y[u] = ld4_stream(a + off);
                } else {
                    d[u] = ld4_stream(a + off);
                    if (a2) e[u] = ld4_stream(a2 + off);
                    y[u] = ld4_stream(c + off);
                    if (relu_from_out) o[u] = ld4_stream(b + off);
                    if (maskbits) mb[u] = ld_maskbits(maskbits, rr * G + g);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long rr = row + (long long)u * TR;
            if (rr < row1) {
                if (KIND == 0) {
                    const float4 v = y[u];
                    s0.x += v.x; s0.y += v.y; s0.z += v.z; s0.w += v.w;
                    s1.x += v.x * v.x; s1.y += v.y * v.y; s1.z += v.z * v.z; s1.w += v.w * v.w;
                } else {
                    float4 dd = d[u];
                    const float4 yy = y[u];
                    if (a2) { dd.x += e[u].x; dd.y += e[u].y; dd.z += e[u].z; dd.w += e[u].w; }
                    if (maskbits) dd = apply_maskbits(dd, mb[u], rr * G + g);
                    if (relu_from_out || relu_from_y) {
                        float4 oo;
                        if (relu_from_out) {
                            oo = o[u];
                        } else {
                            oo.x = fmaf(yy.x, ksc.x, ksh.x); oo.y = fmaf(yy.y, ksc.y, ksh.y);
                            oo.z = fmaf(yy.z, ksc.z, ksh.z); oo.w = fmaf(yy.w, ksc.w, ksh.w);
                        }
                        dd.x = oo.x > 0.f ? dd.x : 0.f;
                        dd.y = oo.y > 0.f ? dd.y : 0.f;
                        dd.z = oo.z > 0.f ? dd.z : 0.f;
                        dd.w = oo.w > 0.f ? dd.w : 0.f;
                    }
                    s0.x += dd.x; s0.y += dd.y; s0.z += dd.z; s0.w += dd.w;
                    s1.x += dd.x * (yy.x - mu.x) * is.x;
                    s1.y += dd.y * (yy.y - mu.y) * is.y;
                    s1.z += dd.z * (yy.z - mu.z) * is.z;
                    s1.w += dd.w * (yy.w - mu.w) * is.w;
                }
            }
        }
    }
    st4(sm0 + 4 * threadIdx.x, s0);
    st4(sm1 + 4 * threadIdx.x, s1);
    __syncthreads();
    if (r == 0) {
        for (int k = 1; k < TR; ++k) {
            const float4 t0 = ld4(sm0 + 4 * (threadIdx.x + k * GW));
            const float4 t1 = ld4(sm1 + 4 * (threadIdx.x + k * GW));
            s0.x += t0.x; s0.y += t0.y; s0.z += t0.z; s0.w += t0.w;
            s1.x += t1.x; s1.y += t1.y; s1.z += t1.z; s1.w += t1.w;
        }
        double* p0 = sums + 4 * g;
        double* p1 = sums + C + 4 * g;
        atomicAdd(p0 + 0, (double)s0.x); atomicAdd(p0 + 1, (double)s0.y);
        atomicAdd(p0 + 2, (double)s0.z); atomicAdd(p0 + 3, (double)s0.w);
        atomicAdd(p1 + 0, (double)s1.x); atomicAdd(p1 + 1, (double)s1.y);
        atomicAdd(p1 + 2, (double)s1.z); atomicAdd(p1 + 3, (double)s1.w);
    }
}

template <typename K>
dim3 reduce_grid(K kernel, long long P, int C) {
    const int G = C / 4;
    const int gy = G > EW_THREADS ? G / EW_THREADS : 1;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, EW_THREADS, 0) != cudaSuccess || occ < 1) occ = 2;
    long long gx = (long long)num_sms() * occ / gy;          // one wave of resident CTAs
    const long long max_gx = (P + 31) / 32;  // at least ~32 rows per block
    if (gx > max_gx) gx = max_gx;
    if (gx < 1) gx = 1;
    return dim3((unsigned)gx, (unsigned)gy, 1);
}

// Batch mean / invstd of one channel from the fp64 sums: E[x^2] - E[x]^2 in double (cancellation), the reciprocal
// square root in float -- one IEEE sqrt + divide instead of their ~150-instruction double versions, which every CTA
// of the apply kernels runs for every channel before it can touch an activation.
__device__ __forceinline__ void bn_moments(const double* __restrict__ stats, int C, int c, double inv_count,
                                           float eps, float& mean, float& invstd, double& var) {
    const double m = stats[c] * inv_count;
    var = stats[C + c] * inv_count - m * m;
    if (var < 0.0) var = 0.0;
    mean = (float)m;
    invstd = 1.f / sqrtf((float)var + eps);
}

__global__ void bn_finalize_kernel(const double* __restrict__ stats, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float* __restrict__ scale,
                                   float* __restrict__ shift, float* __restrict__ mean_out,
                                   float* __restrict__ invstd_out, double count, float momentum, float eps, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float mean, invstd;
    if (stats) {
        double var;
        bn_moments(stats, C, c, 1.0 / count, eps, mean, invstd, var);
        if (running_mean) {
            const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
            running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
            running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
        }
    } else {
        mean = running_mean[c];
        invstd = 1.f / sqrtf(running_var[c] + eps);
    }
    const float sc = gamma[c] * invstd;
    scale[c] = sc;
    shift[c] = beta[c] - mean * sc;
    if (mean_out) mean_out[c] = mean;
    if (invstd_out) invstd_out[c] = invstd;
}

__global__ void __launch_bounds__(EW_THREADS)
bn_apply_kernel(const float* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                const float* __restrict__ residual, float* __restrict__ out, long long n4, int G, int relu,
                int round_out) {
    pdl_sync();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const int g = (int)(i % G);
        const float4 v = ld4_stream(y + 4 * i);
        const float4 sc = ld4(scale + 4 * g);
        const float4 sh = ld4(shift + 4 * g);
        float4 o;
        o.x = fmaf(v.x, sc.x, sh.x);
        o.y = fmaf(v.y, sc.y, sh.y);
        o.z = fmaf(v.z, sc.z, sh.z);
        o.w = fmaf(v.w, sc.w, sh.w);
        if (residual) {
            const float4 r = ld4_stream(residual + 4 * i);
            o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
        }
        if (relu) {
            o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
        }
        if (round_out) o = round4(o);
        st4(out + 4 * i, o);
    }
}

// Training-mode BatchNorm forward in one pass over the activations: every block rebuilds the per-channel
// scale / shift from the fp64 batch sums (cheap: C <= 2048 channels), block 0 also publishes scale / shift /
// mean / invstd for the backward pass, updates the running statistics and bumps num_batches_tracked.
// Optionally emits the (out > 0) mask as bits for the residual joins (see ld_maskbits).
__global__ void __launch_bounds__(EW_THREADS)
bn_train_apply_kernel(const float* __restrict__ y, const double* __restrict__ stats,
                      const float* __restrict__ gamma, const float* __restrict__ beta,
                      float* __restrict__ running_mean, float* __restrict__ running_var,
                      long long* __restrict__ num_batches_tracked, float* __restrict__ scale_out,
                      float* __restrict__ shift_out, float* __restrict__ mean_out, float* __restrict__ invstd_out,
                      const float* __restrict__ residual, float* __restrict__ out, unsigned* __restrict__ maskbits,
                      long long P, int C, float momentum, float eps, int relu, int round_out) {
    pdl_sync();
    extern __shared__ float coef[];  // [2][C]
    float* csc = coef;
    float* csh = coef + C;
    const double count = (double)P;
    const double inv_count = 1.0 / count;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float mean, invstd;
        double var;
        bn_moments(stats, C, c, inv_count, eps, mean, invstd, var);
        const float sc = gamma[c] * invstd;
        const float sh = beta[c] - mean * sc;
        csc[c] = sc;
        csh[c] = sh;
        if (blockIdx.x == 0) {
            scale_out[c] = sc;
            shift_out[c] = sh;
            mean_out[c] = mean;
            invstd_out[c] = invstd;
            if (running_mean) {
                const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
                running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
                running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
            }
            if (c == 0 && num_batches_tracked) *num_batches_tracked += 1;
        }
    }
    __syncthreads();
    const int G = C >> 2;
    const long long n4 = P * G;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    // warp-uniform loop (the ballots below need every lane); two chunks per trip so that every thread keeps two
    // (four with a residual) independent 128-bit loads in flight
    for (long long i0 = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31); i0 < n4; i0 += 2 * stride) {
        long long ii[2];
        bool valid[2];
        float4 v[2], r[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            ii[u] = i0 + u * stride + lane;
            valid[u] = ii[u] < n4;
            if (valid[u]) {
                v[u] = ld4_stream(y + 4 * ii[u]);
                if (residual) r[u] = ld4_stream(residual + 4 * ii[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid[u]) {
                const int g = (int)(ii[u] % G);
                const float4 sc = ld4(csc + 4 * g);
                const float4 sh = ld4(csh + 4 * g);
                o.x = fmaf(v[u].x, sc.x, sh.x);
                o.y = fmaf(v[u].y, sc.y, sh.y);
                o.z = fmaf(v[u].z, sc.z, sh.z);
                o.w = fmaf(v[u].w, sc.w, sh.w);
                if (residual) {
                    o.x += r[u].x; o.y += r[u].y; o.z += r[u].z; o.w += r[u].w;
                }
                if (relu) {
                    o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
                }
                if (round_out) o = round4(o);
                st4(out + 4 * ii[u], o);
            }
            // (a chunk that starts past the end has no mask words: i0 + u*stride < n4 is warp-uniform)
            if (maskbits && i0 + u * stride < n4) {
                const unsigned bx = __ballot_sync(0xffffffffu, o.x > 0.f);
                const unsigned by = __ballot_sync(0xffffffffu, o.y > 0.f);
                const unsigned bz = __ballot_sync(0xffffffffu, o.z > 0.f);
                const unsigned bw = __ballot_sync(0xffffffffu, o.w > 0.f);
                if (lane == 0)
                    *(reinterpret_cast<uint4*>(maskbits) + ((i0 + u * stride) >> 5)) = make_uint4(bx, by, bz, bw);
            }
        }
    }
}

// dy = a*g + b*y + c per channel, a = gamma*invstd, b = -a*k2*invstd, c = -a*k1 + a*k2*invstd*mean,
// k1 = sum_g/P, k2 = sum_gxhat/P.  Coefficients are rebuilt per block into shared memory.
__global__ void __launch_bounds__(EW_THREADS)
bn_bwd_apply_kernel(const float* __restrict__ dout, const float* __restrict__ dout2,
                    const float* __restrict__ out, const float* __restrict__ y, const float* __restrict__ mean,
                    const float* __restrict__ invstd, const float* __restrict__ gamma,
                    const float* __restrict__ msc, const float* __restrict__ msh,
                    const unsigned* __restrict__ maskbits, const double* __restrict__ sums,
                    float* __restrict__ dy, float* __restrict__ dres,
                    int dres_acc, float* __restrict__ dgamma, float* __restrict__ dbeta, int param_acc,
                    long long P, int C, int relu, int round_out) {
    pdl_sync();
    extern __shared__ float coef[];  // [3][C]
    float* ca = coef;
    float* cb = coef + C;
    float* cc = coef + 2 * C;
    const double invP = 1.0 / (double)P;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const double sg = sums[c], sgx = sums[C + c];
        const float is = invstd[c], mu = mean[c];
        const float a = gamma[c] * is;
        const float k1 = (float)(sg * invP), k2 = (float)(sgx * invP);
        ca[c] = a;
        cb[c] = -a * k2 * is;
        cc[c] = -a * k1 + a * k2 * is * mu;
        if (blockIdx.x == 0) {
            if (dgamma) dgamma[c] = (param_acc ? dgamma[c] : 0.f) + (float)sgx;
            if (dbeta) dbeta[c] = (param_acc ? dbeta[c] : 0.f) + (float)sg;
        }
    }
    __syncthreads();
    const int G = C >> 2;
    const long long n4 = P * G;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const int g = (int)(i % G);
        float4 d = ld4_stream(dout + 4 * i);
        if (dout2) {
            const float4 e = ld4_stream(dout2 + 4 * i);
            d.x += e.x; d.y += e.y; d.z += e.z; d.w += e.w;
        }
        const float4 yv = ld4_stream(y + 4 * i);
        if (maskbits) d = apply_maskbits(d, ld_maskbits(maskbits, i), i);
        if (relu) {
            float4 o;
            if (out) {
                o = ld4_stream(out + 4 * i);
            } else {
                const float4 ksc = ld4(msc + 4 * g), ksh = ld4(msh + 4 * g);
                o.x = fmaf(yv.x, ksc.x, ksh.x); o.y = fmaf(yv.y, ksc.y, ksh.y);
                o.z = fmaf(yv.z, ksc.z, ksh.z); o.w = fmaf(yv.w, ksc.w, ksh.w);
            }
            d.x = o.x > 0.f ? d.x : 0.f;
            d.y = o.y > 0.f ? d.y : 0.f;
            d.z = o.z > 0.f ? d.z : 0.f;
            d.w = o.w > 0.f ? d.w : 0.f;
        }
        if (dres) {
            if (dres_acc) {
                float4 r = ld4(dres + 4 * i);
                r.x += d.x; r.y += d.y; r.z += d.z; r.w += d.w;
                st4(dres + 4 * i, r);
            } else {
                st4(dres + 4 * i, d);
            }
        }
        const float4 a = ld4(ca + 4 * g), b = ld4(cb + 4 * g), c = ld4(cc + 4 * g);
        float4 r;
        r.x = fmaf(a.x, d.x, fmaf(b.x, yv.x, c.x));
        r.y = fmaf(a.y, d.y, fmaf(b.y, yv.y, c.y));
        r.z = fmaf(a.z, d.z, fmaf(b.z, yv.z, c.z));
        r.w = fmaf(a.w, d.w, fmaf(b.w, yv.w, c.w));
        if (round_out) r = round4(r);
        st4(dy + 4 * i, r);
    }
}

// ---------------------------------------------------------------------------------------------
// weight packing
// ---------------------------------------------------------------------------------------------
__global__ void pack_weight_kernel(const float* __restrict__ w, float* __restrict__ tck, float* __restrict__ tkc,
                                   int Cout, int Cin, int RS, int round_out) {
    const long long n = (long long)Cout * Cin * RS;
    const long long stride = (long long)gridDim.x * blockDim.x;
    // iterate in tck order (coalesced writes): idx = (t*Cout + co)*Cin + ci
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int ci = (int)(i % Cin);
        const long long r = i / Cin;
        const int co = (int)(r % Cout);
        const int t = (int)(r / Cout);
        float v = w[((long long)co * Cin + ci) * RS + t];
        if (round_out) v = round_tf32(v);
        if (tck) tck[i] = v;
        if (tkc) tkc[((long long)t * Cin + ci) * Cout + co] = v;
    }
}

// All conv layers of a trunk in ONE launch: table rows = {src OIHW, tck, tkc (or 0), Cout, Cin, RS, first block,
// elements}; every block handles PACK_PER_BLOCK consecutive tck-order elements of the layer it falls into.
constexpr int PACK_PER_BLOCK = 2048;
__global__ void __launch_bounds__(EW_THREADS)
pack_weights_batched_kernel(const long long* __restrict__ table, int n_layers, int round_out) {
    pdl_sync();
    __shared__ int s_layer;
    if (threadIdx.x == 0) {
        int lo = 0, hi = n_layers - 1;                     // last layer whose first block <= blockIdx.x
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (table[mid * 8 + 6] <= (long long)blockIdx.x) lo = mid; else hi = mid - 1;
        }
        s_layer = lo;
    }
    __syncthreads();
    const long long* row = table + s_layer * 8;
    const float* __restrict__ w = reinterpret_cast<const float*>(row[0]);
    float* __restrict__ tck = reinterpret_cast<float*>(row[1]);
    float* __restrict__ tkc = reinterpret_cast<float*>(row[2]);
    const int Cout = (int)row[3], Cin = (int)row[4], RS = (int)row[5];
    const long long n = row[7];
    const long long i0 = ((long long)blockIdx.x - row[6]) * PACK_PER_BLOCK;
    for (int k = threadIdx.x; k < PACK_PER_BLOCK; k += EW_THREADS) {
        const long long i = i0 + k;
        if (i >= n) break;
        const int ci = (int)(i % Cin);
        const long long r = i / Cin;
        const int co = (int)(r % Cout);
        const int t = (int)(r / Cout);
        float v = w[((long long)co * Cin + ci) * RS + t];
        if (round_out) v = round_tf32(v);
        tck[i] = v;
        if (tkc) tkc[((long long)t * Cin + ci) * Cout + co] = v;
    }
}

__global__ void unpack_wgrad_kernel(const float* __restrict__ tck, float* __restrict__ w, int Cout, int Cin, int RS,
                                    int accumulate) {
    pdl_sync();
    const long long n = (long long)Cout * Cin * RS;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        // idx in OIHW order: i = (co*Cin + ci)*RS + t
        const int t = (int)(i % RS);
        const long long r = i / RS;
        const int ci = (int)(r % Cin);
        const int co = (int)(r / Cin);
        const float v = tck[((long long)t * Cout + co) * Cin + ci];
        w[i] = accumulate ? w[i] + v : v;
    }
}

// ---------------------------------------------------------------------------------------------
// space-to-depth stem operand (pe_stem_conv_fwd / pe_stem_conv_wgrad, pe_gemm_api.cu)
// ---------------------------------------------------------------------------------------------
// img NCHW (3 channels) -> s2d [B][H/2 + 3][W/2 + 3][12], channel k = a * 6 + b * 3 + c holds img[c][2 I + a][2 J + b]
// at padded position (I + 2, J + 2); the border (2 top / left, 1 bottom / right) is written as zeros.  One thread per
// s2d pixel: six coalesced 64-bit loads, three 128-bit stores.
__global__ void __launch_bounds__(EW_THREADS)
stem_s2d_pack_kernel(const float* __restrict__ img, float* __restrict__ s2d, int B, int H, int W, int round_out) {
    pdl_sync();
    const int Hs = H / 2 + 3, Ws = W / 2 + 3;
    const long long n = (long long)B * Hs * Ws;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int jp = (int)(i % Ws);
        const long long r = i / Ws;
        const int ip = (int)(r % Hs);
        const int b = (int)(r / Hs);
        const int I = ip - 2, J = jp - 2;
        float v[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) v[k] = 0.f;
        if (I >= 0 && I < H / 2 && J >= 0 && J < W / 2) {
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int a = 0; a < 2; ++a) {
                    const float2 t = *reinterpret_cast<const float2*>(
                        img + (((long long)b * 3 + c) * H + (2 * I + a)) * W + 2 * J);
                    v[a * 6 + c] = round_out ? round_tf32(t.x) : t.x;
                    v[a * 6 + 3 + c] = round_out ? round_tf32(t.y) : t.y;
                }
        }
        float* o = s2d + i * 12;
        st4(o, make_float4(v[0], v[1], v[2], v[3]));
        st4(o + 4, make_float4(v[4], v[5], v[6], v[7]));
        st4(o + 8, make_float4(v[8], v[9], v[10], v[11]));
    }
}

// conv1 weight OIHW [Cout][3][7][7]  <->  [4 filter rows U][Cout][64]: column V * 12 + a * 6 + b * 3 + c of row U holds
// w[co][c][2 U + a - 1][2 V + b - 1] (zero where an index is -1, and in columns 48..63).  dir = 0 packs (rounding to
// TF32 on request), dir = 1 scatters a gradient in the packed layout back to OIHW.
__global__ void __launch_bounds__(EW_THREADS)
stem_weight_kernel(float* __restrict__ w_oihw, float* __restrict__ w_s2d, int Cout, int dir, int round_out) {
    pdl_sync();
    const int n = 4 * Cout * 64;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int k = i & 63;
        const int co = (i >> 6) % Cout;
        const int U = i / (64 * Cout);
        const int V = k / 12, rem = k % 12;
        const int a = rem / 6, bq = (rem % 6) / 3, c = rem % 3;
        const int r = 2 * U + a - 1, q = 2 * V + bq - 1;
        const bool valid = k < 48 && r >= 0 && q >= 0;
        if (dir == 0) {
            float v = valid ? w_oihw[((co * 3 + c) * 7 + r) * 7 + q] : 0.f;
            w_s2d[i] = round_out ? round_tf32(v) : v;
        } else if (valid) {
            w_oihw[((co * 3 + c) * 7 + r) * 7 + q] = w_s2d[i];
        }
    }
}

// One block per output row (b, ho): the R input rows of every channel it needs are staged (zero padded,
// TF32-rounded) in shared memory, then written out as Wo im2col rows with 128-bit coalesced stores.
// blockDim = (ldc/4, rows in flight); thread x owns the same four k columns for every row.
__global__ void __launch_bounds__(EW_THREADS, 6)
im2col_stem_kernel(const float* __restrict__ img, float* __restrict__ col, int B, int C, int H, int W, int R, int S,
                   int stride, int pad, int Ho, int Wo, int ldc, int round_out) {
    pdl_sync();
    extern __shared__ float rows_sm[];                // [C*R + 1][SW]
    const int SW = W + 2 * pad;
    const int K = C * R * S;
    const int b = blockIdx.x / Ho, ho = blockIdx.x % Ho;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
    const float* src = img + (size_t)b * C * H * W;
    const int h0 = ho * stride - pad;
    // groups of 64 threads walk whole input rows: (c, r) advance incrementally, the row pointer and the in-range test
    // are per row, and the body is three plain segments (left pad, 128-bit loads of the row, right pad) -- the
    // per-element div / mod version was issue-bound (85 % issue slots busy at 3.4 TB/s)
    const int gsz = nthr >= 64 ? 64 : nthr;
    const int grp = tid / gsz, gl = tid - grp * gsz, ngrp = nthr / gsz;
    if (grp < ngrp) {
        int c = grp / R, r = grp - c * R;
        const bool vec = (W & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
        for (int cr = grp; cr < C * R; cr += ngrp) {
            const int h = h0 + r;
            float* dstp = rows_sm + cr * SW;
            if (h < 0 || h >= H) {
                for (int x = gl; x < SW; x += gsz) dstp[x] = 0.f;
            } else {
                const float* rowp = src + ((size_t)c * H + h) * W;
                for (int x = gl; x < pad; x += gsz) {
                    dstp[x] = 0.f;
                    dstp[pad + W + x] = 0.f;
                }
                float* body = dstp + pad;
                if (vec) {
                    const int W4 = W >> 2;
                    for (int q = gl; q < W4; q += gsz) {
                        float4 v = __ldg(reinterpret_cast<const float4*>(rowp) + q);
                        if (round_out) v = round4(v);
                        body[4 * q + 0] = v.x; body[4 * q + 1] = v.y; body[4 * q + 2] = v.z; body[4 * q + 3] = v.w;
                    }
                } else {
                    for (int w = gl; w < W; w += gsz) {
                        const float v = __ldg(rowp + w);
                        body[w] = round_out ? round_tf32(v) : v;
                    }
                }
            }
            r += ngrp;
            while (r >= R) { r -= R; ++c; }
        }
    }
    // one extra, all-zero row: the padding columns k >= K read it, so the copy loop below has no predicates
    // (the predicated version re-derived four shared addresses per store: 47 instructions per 16 bytes)
    for (int x = tid; x < SW; x += nthr) rows_sm[C * R * SW + x] = 0.f;
    const float* src_k[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int k = 4 * threadIdx.x + j;
        src_k[j] = rows_sm + (k < K ? ((k / S) * SW + (k % S)) : C * R * SW) + threadIdx.y * stride;  // k = (c*R + r)*S + s
    }
    __syncthreads();
    float* dst = col + ((size_t)blockIdx.x * Wo + threadIdx.y) * ldc + 4 * threadIdx.x;
    const int xstep = blockDim.y * stride;
    const size_t dstep = (size_t)blockDim.y * ldc;
    for (int wo = threadIdx.y; wo < Wo; wo += blockDim.y) {
        float4 v;
        v.x = *src_k[0]; v.y = *src_k[1]; v.z = *src_k[2]; v.w = *src_k[3];
        st4(dst, v);
        src_k[0] += xstep; src_k[1] += xstep; src_k[2] += xstep; src_k[3] += xstep;
        dst += dstep;
    }
}

__global__ void transpose_kernel(const float* __restrict__ src, int lds, float* __restrict__ dst, int ldd, int rows,
                                 int cols, int round_out) {
    pdl_sync();
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int r = r0 + j, c = c0 + threadIdx.x;
        tile[j][threadIdx.x] = (r < rows && c < cols) ? src[(long long)r * lds + c] : 0.f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int c = c0 + j, r = r0 + threadIdx.x;
        if (r < rows && c < cols) {
            float v = tile[threadIdx.x][j];
            if (round_out) v = round_tf32(v);
            dst[(long long)c * ldd + r] = v;
        }
    }
}

__global__ void copy_cols_kernel(const float* __restrict__ src, int lds, float* __restrict__ dst, int ldd,
                                 long long rows, int cols, int round_out) {
    pdl_sync();
    const long long n = rows * cols;
    const long long gs = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) {
        const int c = (int)(i % cols);
        const long long r = i / cols;
        float v = src[r * lds + c];
        if (round_out) v = round_tf32(v);
        dst[r * ldd + c] = v;
    }
}

__global__ void axpby_cols_kernel(const float* __restrict__ a, int lda, const float* __restrict__ b, int ldb,
                                  float* __restrict__ out, int ldo, long long rows, int cols, float alpha,
                                  float beta, int round_out) {
    pdl_sync();
    const long long n = rows * cols;
    const long long gs = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) {
        const int c = (int)(i % cols);
        const long long r = i / cols;
        float v = alpha * a[r * lda + c] + beta * b[r * ldb + c];
        if (round_out) v = round_tf32(v);
        out[r * ldo + c] = v;
    }
}

__global__ void colsum_kernel(const float* __restrict__ x, int ldx, float* __restrict__ out, int rows, int cols,
                              int accumulate) {
    pdl_sync();
    // one warp per column chunk of 32; blockDim = (32, 8): 8 row lanes
    __shared__ float sm[8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    float s = 0.f;
    if (c < cols)
        for (int r = threadIdx.y; r < rows; r += 8) s += x[(long long)r * ldx + c];
    sm[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < cols) {
        for (int k = 1; k < 8; ++k) s += sm[k][threadIdx.x];
        out[c] = accumulate ? out[c] + s : s;
    }
}

__global__ void relu_bwd_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ y, int ldy,
                                float* __restrict__ dz, int lddz, long long rows, int cols) {
    pdl_sync();
    const long long n = rows * cols;
    const long long gs = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) {
        const int c = (int)(i % cols);
        const long long r = i / cols;
        dz[r * lddz + c] = y[r * ldy + c] > 0.f ? dy[r * lddy + c] : 0.f;
    }
}

__global__ void fill_kernel(float* __restrict__ p, long long n, float v) {
    const long long gs = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) p[i] = v;
}

__global__ void add_i64_kernel(long long* __restrict__ p, long long n, long long v) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] += v;
}

// ---------------------------------------------------------------------------------------------
// stem pooling
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(EW_THREADS)
maxpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, unsigned char* __restrict__ argmax, int B,
                   int H, int W, int C, int Ho, int Wo) {
    pdl_sync();
    const int G = C >> 2;
    const long long n = (long long)B * Ho * Wo * G;
    const long long gs = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) {
        const int g = (int)(i % G);
        long long r = i / G;
        const int wo = (int)(r % Wo);
        r /= Wo;
        const int ho = (int)(r % Ho);
        const int b = (int)(r / Ho);
        float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        uchar4 am = make_uchar4(255, 255, 255, 255);
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const int h = 2 * ho - 1 + kh;
            if (h < 0 || h >= H) continue;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int w = 2 * wo - 1 + kw;
                if (w < 0 || w >= W) continue;
                const float4 v = ld4(x + (((long long)b * H + h) * W + w) * C + 4 * g);
                const unsigned char k = (unsigned char)(kh * 3 + kw);
                if (v.x > m.x || am.x == 255) { m.x = v.x; am.x = k; }
                if (v.y > m.y || am.y == 255) { m.y = v.y; am.y = k; }
                if (v.z > m.z || am.z == 255) { m.z = v.z; am.z = k; }
                if (v.w > m.w || am.w == 255) { m.w = v.w; am.w = k; }
            }
        }
        st4(y + 4 * i, m);
        if (argmax) *reinterpret_cast<uchar4*>(argmax + 4 * i) = am;
    }
}

// 3x3 / stride 2 / pad 1 max-pool backward as a gather.  One thread owns a 2x2 block of input pixels (x 4
// channels): the block is touched by exactly the four windows (i, i+1) x (j, j+1), so each window's gradient and
// arg-max byte is read once per block instead of once per pixel (2.25x fewer loads).  The aux branch (1x1 conv to
// one channel + 2x2 max pool over the same activation) pools exactly these 2x2 blocks: its gradient goes to the
// block's arg-max pixel in the same pass.
__global__ void __launch_bounds__(EW_THREADS)
maxpool_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ dy2,
                   const unsigned char* __restrict__ argmax, float* __restrict__ dx, int accumulate, int B, int H,
                   int W, int C, int Ho, int Wo, const float* __restrict__ aux_dout, int aux_lddo,
                   const unsigned char* __restrict__ aux_argmax, const float* __restrict__ aux_w) {
    pdl_sync();
    const int G = C >> 2;
    const int Hb = (H + 1) >> 1, Wb = (W + 1) >> 1;          // 2x2 blocks
    const long long n = (long long)B * Hb * Wb * G;
    const long long gs = (long long)gridDim.x * blockDim.x;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gs) {
        const int g = (int)(idx % G);
        long long r = idx / G;
        const int j = (int)(r % Wb);
        r /= Wb;
        const int i = (int)(r % Hb);
        const int b = (int)(r / Hb);
        float4 acc[4];                                        // pixels (2i + (k>>1), 2j + (k&1))
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        // window ho covers input rows 2ho-1 .. 2ho+1: row 2i is tap row 1 of window i only, row 2i+1 is tap row 2 of
        // window i and tap row 0 of window i+1 (same along w); taps outside 0..2 are filtered below
#pragma unroll
        for (int dh = 0; dh < 2; ++dh) {
            const int ho = i + dh;
            if (ho >= Ho) continue;
#pragma unroll
            for (int dw = 0; dw < 2; ++dw) {
                const int wo = j + dw;
                if (wo >= Wo) continue;
                const long long o = (((long long)b * Ho + ho) * Wo + wo) * G + g;
                const uchar4 am = *reinterpret_cast<const uchar4*>(argmax + 4 * o);
                float4 d = ld4(dy + 4 * o);
                if (dy2) {
                    const float4 e = ld4(dy2 + 4 * o);
                    d.x += e.x; d.y += e.y; d.z += e.z; d.w += e.w;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int h = 2 * i + (k >> 1), w = 2 * j + (k & 1);
                    const int kh = h - (2 * ho - 1), kw = w - (2 * wo - 1);     // tap of this pixel in the window
                    if (kh < 0 || kh > 2 || kw < 0 || kw > 2) continue;
                    const unsigned char t = (unsigned char)(kh * 3 + kw);
                    if (am.x == t) acc[k].x += d.x;
                    if (am.y == t) acc[k].y += d.y;
                    if (am.z == t) acc[k].z += d.z;
                    if (am.w == t) acc[k].w += d.w;
                }
            }
        }
        if (aux_dout) {
            const long long win = ((long long)b * (H >> 1) + i) * (W >> 1) + j;
            const int ak = aux_argmax[win];
            const float d = aux_dout[(long long)b * aux_lddo + i * (W >> 1) + j];
            const float4 w4 = ld4(aux_w + 4 * g);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (k == ak) {
                    acc[k].x = fmaf(d, w4.x, acc[k].x); acc[k].y = fmaf(d, w4.y, acc[k].y);
                    acc[k].z = fmaf(d, w4.z, acc[k].z); acc[k].w = fmaf(d, w4.w, acc[k].w);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int h = 2 * i + (k >> 1), w = 2 * j + (k & 1);
            if (h >= H || w >= W) continue;
            float* p = dx + ((((long long)b * H + h) * W + w) * G + g) * 4;
            float4 v = acc[k];
            if (accumulate) {
                const float4 prev = ld4(p);
                v.x += prev.x; v.y += prev.y; v.z += prev.z; v.w += prev.w;
            }
            st4(p, v);
        }
    }
}

__global__ void __launch_bounds__(EW_THREADS)
avgpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int ldy, int B, int HW, int C, int round_out) {
    pdl_sync();
    const int G = C >> 2;
    const long long n = (long long)B * G;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int g = (int)(i % G);
    const int b = (int)(i / G);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* p = x + (long long)b * HW * C + 4 * g;
    for (int k = 0; k < HW; ++k) {
        const float4 v = ld4(p + (long long)k * C);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    const float inv = 1.f / (float)HW;
    s.x *= inv; s.y *= inv; s.z *= inv; s.w *= inv;
    if (round_out) s = round4(s);
    st4(y + (long long)b * ldy + 4 * g, s);
}

__global__ void __launch_bounds__(EW_THREADS)
avgpool_bwd_kernel(const float* __restrict__ dy, int lddy, float* __restrict__ dx, int B, int HW, int C) {
    pdl_sync();
    const int G = C >> 2;
    const long long n = (long long)B * HW * G;
    const long long gs = (long long)gridDim.x * blockDim.x;
    const float inv = 1.f / (float)HW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) {
        const int g = (int)(i % G);
        const int b = (int)(i / ((long long)HW * G));
        float4 d = ld4(dy + (long long)b * lddy + 4 * g);
        d.x *= inv; d.y *= inv; d.z *= inv; d.w *= inv;
        st4(dx + 4 * i, d);
    }
}

// ---------------------------------------------------------------------------------------------
// auxiliary branch: per-pixel <a1[pixel,:], w> + bias, 2x2 max pool, flatten.  One warp per window.
// ---------------------------------------------------------------------------------------------
// Aux branch: Conv2d(C,1,1) + MaxPool2d(2).  A half-warp owns one 2x2 window: lane (l & 15) holds channels
// 4*(l&15) + 64*j, so every pixel is one coalesced 256-byte read per 64 channels and the dot product closes with
// four shuffles inside the half-warp.
__global__ void __launch_bounds__(EW_THREADS)
aux_fwd_kernel(const float* __restrict__ a1, const float* __restrict__ w, const float* __restrict__ bias,
               float* __restrict__ out, int ldo, unsigned char* __restrict__ argmax, int B, int H, int W, int C,
               int round_out) {
    pdl_sync();
    const int Ho = H >> 1, Wo = W >> 1;
    const int lane = threadIdx.x & 31, hl = lane & 15, half = lane >> 4;
    const long long gwarp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long nwin = (long long)B * Ho * Wo;
    const long long npair = (nwin + 1) >> 1;
    const float b0 = bias[0];
    for (long long pair = gwarp; pair < npair; pair += nwarps) {       // warp-uniform trip count
        const long long win = pair * 2 + half;
        const bool valid = win < nwin;
        const long long wv = valid ? win : 0;
        const int wo = (int)(wv % Wo);
        const int ho = (int)((wv / Wo) % Ho);
        const int b = (int)(wv / ((long long)Wo * Ho));
        const float* p0 = a1 + (((long long)b * H + 2 * ho) * W + 2 * wo) * C;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        for (int c = 4 * hl; c < C; c += 64) {
            const float4 w4 = ld4(w + c);
            const float4 v0 = ld4_stream(p0 + c), v1 = ld4_stream(p0 + C + c);
            const float4 v2 = ld4_stream(p0 + (long long)W * C + c), v3 = ld4_stream(p0 + (long long)W * C + C + c);
            s0 += v0.x * w4.x + v0.y * w4.y + v0.z * w4.z + v0.w * w4.w;
            s1 += v1.x * w4.x + v1.y * w4.y + v1.z * w4.z + v1.w * w4.w;
            s2 += v2.x * w4.x + v2.y * w4.y + v2.z * w4.z + v2.w * w4.w;
            s3 += v3.x * w4.x + v3.y * w4.y + v3.z * w4.z + v3.w * w4.w;
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            s3 += __shfl_xor_sync(0xffffffffu, s3, o);
        }
        if (hl == 0 && valid) {
            // first maximum wins ties, like torch's max_pool2d
            float best = s0;
            int besti = 0;
            if (s1 > best) { best = s1; besti = 1; }
            if (s2 > best) { best = s2; besti = 2; }
            if (s3 > best) { best = s3; besti = 3; }
            best += b0;
            out[(long long)b * ldo + ho * Wo + wo] = round_out ? round_tf32(best) : best;
            if (argmax) argmax[win] = (unsigned char)besti;
        }
    }
}

// Aux-branch backward.  da1 (optional): gradient scattered to the arg-max pixel of every window (zeros elsewhere);
// dw / db (optional): the 1x1 conv's own gradients, which only need a1 at the arg-max pixels.
__global__ void __launch_bounds__(EW_THREADS)
aux_bwd_kernel(const float* __restrict__ dout, int lddo, const unsigned char* __restrict__ argmax,
               const float* __restrict__ a1, const float* __restrict__ w, float* __restrict__ da1, int accumulate,
               float* __restrict__ dw, float* __restrict__ db, int B, int H, int W, int C,
               const float* __restrict__ pre_scale, const float* __restrict__ pre_shift, int pre_round) {
    pdl_sync();
    __shared__ float s_dw[256];
    __shared__ float s_db;
    const int Ho = H >> 1, Wo = W >> 1;
    const int lane = threadIdx.x & 31, hl = lane & 15, half = lane >> 4;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_dw[i] = 0.f;
    if (threadIdx.x == 0) s_db = 0.f;
    __syncthreads();
    const long long gwarp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long nwin = (long long)B * Ho * Wo;
    float4 pdw[4];                                  // channels 4*hl + 64*j, j < 4  (C <= 256)
#pragma unroll
    for (int j = 0; j < 4; ++j) pdw[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    float pdb = 0.f;
    for (long long win = gwarp * 2 + half; win < nwin; win += nwarps * 2) {
        const int wo = (int)(win % Wo);
        const int ho = (int)((win / Wo) % Ho);
        const int b = (int)(win / ((long long)Wo * Ho));
        const float d = dout[(long long)b * lddo + ho * Wo + wo];
        const int am = argmax[win];
        const long long base = (((long long)b * H + 2 * ho) * W + 2 * wo) * C;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = 4 * hl + 64 * j;
            if (c < C) {
                if (da1) {
                    const float4 w4 = ld4(w + c);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const long long off = base + ((long long)(k >> 1) * W + (k & 1)) * C + c;
                        const float dk = (k == am) ? d : 0.f;
                        float4 r = accumulate ? ld4(da1 + off) : make_float4(0.f, 0.f, 0.f, 0.f);
                        r.x += dk * w4.x; r.y += dk * w4.y; r.z += dk * w4.z; r.w += dk * w4.w;
                        st4(da1 + off, r);
                    }
                }
                if (dw) {
                    float4 av = ld4_stream(a1 + base + ((long long)(am >> 1) * W + (am & 1)) * C + c);
                    if (pre_scale) {   // `a1` is the pre-BN tensor: the activation is rebuilt on the fly
                        const float4 sc = ld4(pre_scale + c), sh = ld4(pre_shift + c);
                        av.x = fmaxf(fmaf(av.x, sc.x, sh.x), 0.f); av.y = fmaxf(fmaf(av.y, sc.y, sh.y), 0.f);
                        av.z = fmaxf(fmaf(av.z, sc.z, sh.z), 0.f); av.w = fmaxf(fmaf(av.w, sc.w, sh.w), 0.f);
                        if (pre_round) av = round4(av);
                    }
                    pdw[j].x = fmaf(d, av.x, pdw[j].x); pdw[j].y = fmaf(d, av.y, pdw[j].y);
                    pdw[j].z = fmaf(d, av.z, pdw[j].z); pdw[j].w = fmaf(d, av.w, pdw[j].w);
                }
            }
        }
        if (hl == 0) pdb += d;
    }
    if (dw) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = 4 * hl + 64 * j;
            if (c < C) {
                atomicAdd(&s_dw[c], pdw[j].x); atomicAdd(&s_dw[c + 1], pdw[j].y);
                atomicAdd(&s_dw[c + 2], pdw[j].z); atomicAdd(&s_dw[c + 3], pdw[j].w);
            }
        }
        if (hl == 0) atomicAdd(&s_db, pdb);
        __syncthreads();
        for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(dw + c, s_dw[c]);
        if (threadIdx.x == 0 && db) atomicAdd(db, s_db);
    }
}

// Training-mode stem tail in ONE pass over the conv1 output y [B,H,W,64]: BatchNorm with batch statistics + ReLU,
// the 3x3 / stride 2 / pad 1 max pool and the aux branch (1x1 conv to one channel + 2x2 max pool, whose window is
// the lower-right 2x2 of the pooling window).  The normalised activation is never written: the backward pass needs
// the two arg-max maps, y and scale / shift only (bn_train_apply + maxpool_fwd + aux_fwd would write it once and
// read it twice: 2.5 GB per 256-frame step).  A half-warp owns one pooled pixel (16 lanes x 4 channels); the
// arithmetic is that of the three separate kernels, expression for expression.
__global__ void __launch_bounds__(EW_THREADS, 4)
stem_post_train_kernel(const float* __restrict__ y, const double* __restrict__ stats,
                       const float* __restrict__ gamma, const float* __restrict__ beta,
                       float* __restrict__ running_mean, float* __restrict__ running_var,
                       long long* __restrict__ num_batches_tracked, float* __restrict__ scale_out,
                       float* __restrict__ shift_out, float* __restrict__ mean_out, float* __restrict__ invstd_out,
                       float* __restrict__ pool, unsigned char* __restrict__ pool_argmax,
                       const float* __restrict__ aux_w, const float* __restrict__ aux_bias,
                       float* __restrict__ aux_out, int ld_aux, unsigned char* __restrict__ aux_argmax, int B, int H,
                       int W, float momentum, float eps, int round_out, int aux_round) {
    pdl_sync();
    constexpr int C = 64;
    __shared__ __align__(16) float csc[C];
    __shared__ __align__(16) float csh[C];
    const double count = (double)B * H * W;
    const double inv_count = 1.0 / count;
    if (threadIdx.x < C) {
        const int c = threadIdx.x;
        float mean, invstd;
        double var;
        bn_moments(stats, C, c, inv_count, eps, mean, invstd, var);
        const float sc = gamma[c] * invstd;
        const float sh = beta[c] - mean * sc;
        csc[c] = sc;
        csh[c] = sh;
        if (blockIdx.x == 0) {
            scale_out[c] = sc;
            shift_out[c] = sh;
            mean_out[c] = mean;
            invstd_out[c] = invstd;
            if (running_mean) {
                const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
                running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
                running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
            }
            if (c == 0 && num_batches_tracked) *num_batches_tracked += 1;
        }
    }
    __syncthreads();
    const int Ho = H >> 1, Wo = W >> 1;
    const int lane = threadIdx.x & 31, hl = lane & 15, half = lane >> 4;
    const float4 sc = ld4(csc + 4 * hl), sh = ld4(csh + 4 * hl);
    const float4 w4 = aux_w ? ld4(aux_w + 4 * hl) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float b0 = aux_w ? aux_bias[0] : 0.f;
    const long long gwarp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long nwin = (long long)B * Ho * Wo;
    const long long npair = (nwin + 1) >> 1;
    for (long long pair = gwarp; pair < npair; pair += nwarps) {       // warp-uniform trip count
        const long long win = pair * 2 + half;
        const bool valid = win < nwin;
        const long long wv = valid ? win : 0;
        const int wo = (int)(wv % Wo);
        const int ho = (int)((wv / Wo) % Ho);
        const int b = (int)(wv / ((long long)Wo * Ho));
        float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        uchar4 am = make_uchar4(255, 255, 255, 255);
        float s[4] = {0.f, 0.f, 0.f, 0.f};
        const bool win_w0 = wo > 0;                           // column 2wo-1 exists (h < H, w < W always hold)
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {                      // one window row at a time: three loads in flight
            const int h = 2 * ho - 1 + kh;
            if (h < 0) continue;
            const float* rowp = y + (((long long)b * H + h) * W + 2 * wo - 1) * C + 4 * hl;
            float4 v[3];
#pragma unroll
            for (int kw = 0; kw < 3; ++kw)
                if (kw > 0 || win_w0) v[kw] = ld4(rowp + kw * C);
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                if (kw == 0 && !win_w0) continue;
                float4 o;
                o.x = fmaxf(fmaf(v[kw].x, sc.x, sh.x), 0.f);
                o.y = fmaxf(fmaf(v[kw].y, sc.y, sh.y), 0.f);
                o.z = fmaxf(fmaf(v[kw].z, sc.z, sh.z), 0.f);
                o.w = fmaxf(fmaf(v[kw].w, sc.w, sh.w), 0.f);
                if (round_out) o = round4(o);
                const unsigned char kk = (unsigned char)(kh * 3 + kw);
                if (o.x > m.x || am.x == 255) { m.x = o.x; am.x = kk; }
                if (o.y > m.y || am.y == 255) { m.y = o.y; am.y = kk; }
                if (o.z > m.z || am.z == 255) { m.z = o.z; am.z = kk; }
                if (o.w > m.w || am.w == 255) { m.w = o.w; am.w = kk; }
                if (kh >= 1 && kw >= 1)                       // the aux window: rows 2ho, 2ho+1 x cols 2wo, 2wo+1
                    s[(kh - 1) * 2 + (kw - 1)] += o.x * w4.x + o.y * w4.y + o.z * w4.z + o.w * w4.w;
            }
        }
        if (valid) {
            st4(pool + win * C + 4 * hl, m);
            if (pool_argmax) *reinterpret_cast<uchar4*>(pool_argmax + win * C + 4 * hl) = am;
        }
        if (aux_w) {                                         // warp-uniform
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                s[0] += __shfl_xor_sync(0xffffffffu, s[0], o);
                s[1] += __shfl_xor_sync(0xffffffffu, s[1], o);
                s[2] += __shfl_xor_sync(0xffffffffu, s[2], o);
                s[3] += __shfl_xor_sync(0xffffffffu, s[3], o);
            }
            if (hl == 0 && valid) {
                float best = s[0];                           // first maximum wins ties, like torch's max_pool2d
                int besti = 0;
                if (s[1] > best) { best = s[1]; besti = 1; }
                if (s[2] > best) { best = s[2]; besti = 2; }
                if (s[3] > best) { best = s[3]; besti = 3; }
                best += b0;
                aux_out[(long long)b * ld_aux + ho * Wo + wo] = aux_round ? round_tf32(best) : best;
                if (aux_argmax) aux_argmax[win] = (unsigned char)besti;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// depth branch (use_depth): AvgPool2d(2)^n + InstanceNorm2d(1, affine) + Flatten, multiplied into the aux
// features (reference models/naive.py:233-240,324-330).  One block per frame; the pooled map lives in smem.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum_256(float v, float* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    if (threadIdx.x < 32) {
        t = warp_sum(t);
        if (threadIdx.x == 0) red[0] = t;
    }
    __syncthreads();
    return red[0];
}

__global__ void __launch_bounds__(EW_THREADS)
depth_features_fwd_kernel(const float* __restrict__ depth, int H, int W, int pool, const float* __restrict__ gamma,
                          const float* __restrict__ beta, float eps, float* __restrict__ xhat,
                          float* __restrict__ aux, int ld, float* __restrict__ aux_pre, int round_out) {
    extern __shared__ float pooled[];   // [F]
    __shared__ float red[32];
    const int Ho = H / pool, Wo = W / pool, F = Ho * Wo;
    const int b = blockIdx.x;
    const float* src = depth + (size_t)b * H * W;
    const float inv = 1.f / (float)(pool * pool);
    float part = 0.f;
    for (int k = threadIdx.x; k < F; k += blockDim.x) {
        const int ho = k / Wo, wo = k - ho * Wo;
        float s = 0.f;
        for (int r = 0; r < pool; ++r)
            for (int c = 0; c < pool; ++c) s += __ldg(src + (size_t)(ho * pool + r) * W + wo * pool + c);
        s *= inv;
        pooled[k] = s;
        part += s;
    }
    const float mean = block_sum_256(part, red) / (float)F;
    part = 0.f;
    for (int k = threadIdx.x; k < F; k += blockDim.x) {
        const float d = pooled[k] - mean;
        part = fmaf(d, d, part);
    }
    const float var = block_sum_256(part, red) / (float)F;      // biased, as instance_norm normalises
    const float invstd = rsqrtf(var + eps);
    const float g = gamma[0], bt = beta[0];
    for (int k = threadIdx.x; k < F; k += blockDim.x) {
        const float xh = (pooled[k] - mean) * invstd;
        xhat[(size_t)b * F + k] = xh;
        const float a = aux[(size_t)b * ld + k];
        if (aux_pre) aux_pre[(size_t)b * F + k] = a;
        const float v = a * fmaf(xh, g, bt);
        aux[(size_t)b * ld + k] = round_out ? round_tf32(v) : v;
    }
}

__global__ void __launch_bounds__(EW_THREADS)
depth_features_bwd_kernel(float* __restrict__ dprod, int ld, const float* __restrict__ aux_pre,
                          const float* __restrict__ xhat, const float* __restrict__ gamma,
                          const float* __restrict__ beta, float* __restrict__ dgamma, float* __restrict__ dbeta,
                          int F) {
    __shared__ float red[32];
    const int b = blockIdx.x;
    const float g = gamma[0], bt = beta[0];
    float pg = 0.f, pb = 0.f;
    for (int k = threadIdx.x; k < F; k += blockDim.x) {
        const float d = dprod[(size_t)b * ld + k];
        const float xh = xhat[(size_t)b * F + k];
        const float df = d * aux_pre[(size_t)b * F + k];        // gradient w.r.t. the depth feature
        pg = fmaf(df, xh, pg);
        pb += df;
        dprod[(size_t)b * ld + k] = d * fmaf(xh, g, bt);        // gradient w.r.t. the aux feature, in place
    }
    const float sg = block_sum_256(pg, red);
    const float sb = block_sum_256(pb, red);
    if (threadIdx.x == 0) {
        if (dgamma) atomicAdd(dgamma, sg);
        if (dbeta) atomicAdd(dbeta, sb);
    }
}

// One thread converts FOUR consecutive pixels of a cropped row: 12 contiguous source bytes (three aligned 32-bit loads
// when the row start allows it) -> one 128-bit store per colour plane.  ToTensor's x / 255 followed by Normalize's
// (x - mean) / std with true divisions, like the reference's transform (util/data_utils.py:48-54): bit-identical to
// torchvision on the same uint8 frame.
__global__ void preprocess_u8_kernel(const unsigned char* __restrict__ src, float* __restrict__ dst, int B, int Hs,
                                     int Ws, int crop, float m0, float m1, float m2, float s0, float s1, float s2) {
    const int q = crop >> 2;                                  // 4-pixel groups per row (crop % 4 == 0)
    const long long n = (long long)B * crop * q;
    const long long gs = (long long)gridDim.x * blockDim.x;
    const int oy = (Hs - crop) / 2, ox = (Ws - crop) / 2;
    const long long plane = (long long)crop * crop;
    const float mean[3] = {m0, m1, m2}, sd[3] = {s0, s1, s2};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) {
        const int xg = (int)(i % q);
        const int y = (int)((i / q) % crop);
        const int b = (int)(i / ((long long)q * crop));
        const unsigned char* p = src + (((long long)b * Hs + (y + oy)) * Ws + (4 * xg + ox)) * 3;
        unsigned char px[12];
        if ((reinterpret_cast<uintptr_t>(p) & 3) == 0) {
            const unsigned* w = reinterpret_cast<const unsigned*>(p);
            const unsigned w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                px[k] = (unsigned char)(w0 >> (8 * k));
                px[4 + k] = (unsigned char)(w1 >> (8 * k));
                px[8 + k] = (unsigned char)(w2 >> (8 * k));
            }
        } else {
#pragma unroll
            for (int k = 0; k < 12; ++k) px[k] = p[k];
        }
        float* o = dst + (long long)b * 3 * plane + (long long)y * crop + 4 * xg;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float4 v;
            v.x = __fdiv_rn(__fdiv_rn((float)px[c], 255.f) - mean[c], sd[c]);
            v.y = __fdiv_rn(__fdiv_rn((float)px[3 + c], 255.f) - mean[c], sd[c]);
            v.z = __fdiv_rn(__fdiv_rn((float)px[6 + c], 255.f) - mean[c], sd[c]);
            v.w = __fdiv_rn(__fdiv_rn((float)px[9 + c], 255.f) - mean[c], sd[c]);
            *reinterpret_cast<float4*>(o + c * plane) = v;
        }
    }
}

}  // namespace
}  // namespace pe

using namespace pe;

extern "C" {

int pe_bn_stats(const float* y, long long P, int C, double* stats, void* stream) {
    PE_REQUIRE(C % 4 == 0 && ((C / 4) <= EW_THREADS ? EW_THREADS % (C / 4) == 0 : (C / 4) % EW_THREADS == 0),
               "bn_stats: unsupported channel count %d", C);
    PE_LAUNCH(channel_reduce_kernel<0>, reduce_grid(channel_reduce_kernel<0>, P, C), EW_THREADS, 0, y, nullptr,
              nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, stats, P, C, 0);
    PE_LAUNCH_CHECK();
    return 0;
}

int pe_bn_finalize(double* stats, const float* gamma, const float* beta, float* running_mean, float* running_var,
                   float* scale, float* shift, float* mean, float* invstd, long long count, float momentum,
                   float eps, int C, void* stream) {
    PE_REQUIRE(stats || (running_mean && running_var), "bn_finalize: need stats or running statistics");
    bn_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        stats, gamma, beta, running_mean, running_var, scale, shift, mean, invstd, (double)count, momentum, eps, C);
    PE_LAUNCH_CHECK();
    return 0;
}

int pe_bn_apply(const float* y, const float* scale, const float* shift, const float* residual, float* out,
                long long P, int C, int relu, int round_tf32, void* stream) {
    PE_REQUIRE(C % 4 == 0, "bn_apply: C %% 4 != 0");
    const long long n4 = P * (C / 4);
    PE_LAUNCH(bn_apply_kernel, one_wave_grid(bn_apply_kernel, 0, n4, EW_THREADS * 4), EW_THREADS, 0, y, scale, shift,
              residual, out, n4, C / 4, relu, round_tf32);
    PE_LAUNCH_CHECK();
    return 0;
}

int pe_bn_train_apply(const float* y, const double* stats, const float* gamma, const float* beta,
                      float* running_mean, float* running_var, long long* num_batches_tracked, float* scale,
                      float* shift, float* mean, float* invstd, const float* residual, float* out,
                      unsigned* maskbits, long long P, int C, float momentum, float eps, int relu, int round_tf32,
                      void* stream) {
    PE_REQUIRE(C % 4 == 0 && C <= 8192, "bn_train_apply: unsupported channel count %d", C);
    PE_REQUIRE(stats && scale && shift && mean && invstd, "bn_train_apply: stats / scale / shift / mean / invstd required");
    PE_REQUIRE(!maskbits || (reinterpret_cast<uintptr_t>(maskbits) & 15) == 0,
               "bn_train_apply: mask bits need a 16-byte aligned pointer");
    const long long n4 = P * (C / 4);
    PE_LAUNCH(bn_train_apply_kernel, one_wave_grid(bn_train_apply_kernel, 2 * C * sizeof(float), n4, EW_THREADS * 4),
              EW_THREADS, 2 * C * sizeof(float), y, stats, gamma, beta, running_mean, running_var,
              num_batches_tracked, scale, shift, mean, invstd, residual, out, maskbits, P, C, momentum, eps, relu,
              round_tf32);
    PE_LAUNCH_CHECK();
    return 0;
}

int pe_bn_bwd_reduce(const float* dout, const float* dout2, const float* out, const float* y, const float* mean,
                     const float* invstd, const float* mask_scale, const float* mask_shift,
                     const unsigned* maskbits, double* sums, long long P, int C, int relu, void* stream) {
    PE_REQUIRE(!maskbits || (C % 32 == 0 && (reinterpret_cast<uintptr_t>(maskbits) & 15) == 0),
               "bn_bwd_reduce: mask bits need C %% 32 == 0 and a 16-byte aligned pointer");
    PE_REQUIRE(!relu || out || (mask_scale && mask_shift), "bn_bwd_reduce: ReLU mask needs `out` or scale/shift");
    PE_REQUIRE(C % 4 == 0 && ((C / 4) <= EW_THREADS ? EW_THREADS % (C / 4) == 0 : (C / 4) % EW_THREADS == 0),
               "bn_bwd_reduce: unsupported channel count %d", C);
    PE_LAUNCH(channel_reduce_kernel<1>, reduce_grid(channel_reduce_kernel<1>, P, C), EW_THREADS, 0, dout, dout2, out,
              y, mean, invstd, mask_scale, mask_shift, maskbits, sums, P, C, relu);
    PE_LAUNCH_CHECK();
    return 0;
}

int pe_bn_bwd_apply(const float* dout, const float* dout2, const float* out, const float* y, const float* mean,
                    const float* invstd, const float* gamma, const float* mask_scale, const float* mask_shift,
                    const unsigned* maskbits, double* sums, float* dy, float* dres,
                    int dres_accumulate, float* dgamma, float* dbeta, int param_accumulate, long long P, int C,
                    int relu, int round_tf32, void* stream) {
    PE_REQUIRE(!maskbits || (reinterpret_cast<uintptr_t>(maskbits) & 15) == 0,
               "bn_bwd_apply: mask bits need a 16-byte aligned pointer");
    PE_REQUIRE(C % 4 == 0 && C <= 4096, "bn_bwd_apply: unsupported channel count %d", C);
    PE_REQUIRE(!relu || out || (mask_scale && mask_shift), "bn_bwd_apply: ReLU mask needs `out` or scale/shift");
    const long long n4 = P * (C / 4);
    PE_LAUNCH(bn_bwd_apply_kernel, one_wave_grid(bn_bwd_apply_kernel, 3 * C * sizeof(float), n4, EW_THREADS * 4),
              EW_THREADS, 3 * C * sizeof(float), dout, dout2, out, y, mean, invstd, gamma, mask_scale, mask_shift,
              maskbits, sums, dy, dres, dres_accumulate, dgamma, dbeta, param_accumulate, P, C, relu, round_tf32);
    PE_LAUNCH_CHECK();
    return 0;
}

int pe_pack_conv_weight(const float* w_oihw, float* w_tck, float* w_tkc, int Cout, int Cin, int R, int S,
                        int round_tf32, void* stream) {
    const long long n = (long long)Cout * Cin * R * S;
    pack_weight_kernel<<<grid_for(n, EW_THREADS * 4), EW_THREADS, 0, (cudaStream_t)stream>>>(
        w_oihw, w_tck, w_tkc, Cout, Cin, R * S, round_tf32);
    PE_LAUNCH_CHECK();
    return 0;
}

int pe_pack_conv_weights_batched(const long long* table_dev, int n_layers, int total_blocks, int round_tf32,
                                 void* stream) {
    if (n_layers <= 0 || total_blocks <= 0) return 0;
    PE_LAUNCH(pack_weights_batched_kernel, total_blocks, EW_THREADS, 0, table_dev, n_layers, round_tf32);
    PE_LAUNCH_CHECK();
    return 0;
}

int pe_pack_block_elems(void) { return PACK_PER_BLOCK; }

int pe_unpack_conv_wgrad(const float* dw_tck, float* dw_oihw, int Cout, int Cin, int R, int S, int accumulate,
                         void* stream) {
    const long long n = (long long)Cout * Cin * R * S;
    PE_LAUNCH(unpack_wgrad_kernel, grid_for(n, EW_THREADS * 4), EW_THREADS, 0, dw_tck, dw_oihw, Cout, Cin, R * S,
              accumulate);
    PE_LAUNCH_CHECK();
    return 0;
}

int pe_stem_s2d_pack(const float* img_nchw, float* s2d, int B, int H, int W, int round_tf32, void* stream) {
    PE_REQUIRE(H % 2 == 0 && W % 2 == 0, "stem s2d: image %d x %d must have even sides", H, W);
    PE_REQUIRE((reinterpret_cast<uintptr_t>(img_nchw) & 7) == 0 && (reinterpret_cast<uintptr_t>(s2d) & 15) == 0,
               "stem s2d: image / output pointers must be 8- / 16-byte aligned");
    const long long n = (long long)B * (H / 2 + 3) * (W / 2 + 3);
    PE_LAUNCH(stem_s2d_pack_kernel, grid_for(n, EW_THREADS), EW_THREADS, 0, img_nchw, s2d, B, H, W, round_tf32);
    PE_LAUNCH_CHECK();
    return 0;
}

int pe_stem_pack_weight(const float* w_oihw, float* w_s2d, int Cout, int round_tf32, void* stream) {
    PE_LAUNCH(stem_weight_kernel, grid_for(4ll * Cout * 64, EW_THREADS), EW_THREADS, 0, const_cast<float*>(w_oihw), w_s2d,
              Cout, 0, round_tf32);
    PE_LAUNCH_CHECK();
    return 0;
}

int pe_stem_unpack_wgrad(const float* dw_s2d, float* dw_oihw, int Cout, void* stream) {
    PE_LAUNCH(stem_weight_kernel, grid_for(4ll * Cout * 64, EW_THREADS), EW_THREADS, 0, dw_oihw, const_cast<float*>(dw_s2d),
              Cout, 1, 0);
    PE_LAUNCH_CHECK();
    return 0;
}

int pe_im2col_stem(const float* img_nchw, float* col, int B, int C, int H, int W, int R, int S, int stride, int pad,
                   int ldc, int round_tf32, void* stream) {
    PE_REQUIRE(ldc >= C * R * S && ldc % 4 == 0, "im2col: ldc %d too small / unaligned", ldc);
    const int Ho = (H + 2 * pad - R) / stride + 1, Wo = (W + 2 * pad - S) / stride + 1;
    const int bx = ldc / 4;
    PE_REQUIRE(bx >= 1 && bx <= EW_THREADS, "im2col: ldc %d out of range", ldc);
    const size_t smem = sizeof(float) * ((size_t)C * R + 1) * (W + 2 * pad);      // + one all-zero row
    PE_REQUIRE(smem <= 160 * 1024, "im2col: %d x %d rows of %d floats do not fit in shared memory", C, R, W + 2 * pad);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        PE_CHECK_CUDA(cudaFuncSetAttribute(im2col_stem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    dim3 block(bx, EW_THREADS / bx);
    PE_LAUNCH(im2col_stem_kernel, B * Ho, block, smem, img_nchw, col, B, C, H, W, R, S, stride, pad, Ho, Wo, ldc,
              round_tf32);
    PE_LAUNCH_CHECK();
    return 0;
}

int pe_transpose(const float* src, int lds, float* dst, int ldd, int rows, int cols, int round_tf32, void* stream) {
    dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
    PE_LAUNCH(transpose_kernel, grid, block, 0, src, lds, dst, ldd, rows, cols, round_tf32);
    PE_LAUNCH_CHECK();
    return 0;
}

int pe_copy_cols(const float* src, int lds, float* dst, int ldd, int rows, int cols, int round_tf32, void* stream) {
    const long long n = (long long)rows * cols;
    if (n == 0) return 0;
    PE_LAUNCH(copy_cols_kernel, grid_for(n, EW_THREADS), EW_THREADS, 0, src, lds, dst, ldd, rows, cols, round_tf32);
    PE_LAUNCH_CHECK();
    return 0;
}

int pe_axpby_cols(const float* a, int lda, const float* b, int ldb, float* out, int ldo, int rows, int cols,
                  float alpha, float beta, int round_tf32, void* stream) {
    const long long n = (long long)rows * cols;
    if (n == 0) return 0;
    PE_LAUNCH(axpby_cols_kernel, grid_for(n, EW_THREADS), EW_THREADS, 0, a, lda, b, ldb, out, ldo, rows, cols, alpha,
              beta, round_tf32);
    PE_LAUNCH_CHECK();
    return 0;
}

int pe_colsum(const float* x, int ldx, float* out, int rows, int cols, int accumulate, void* stream) {
    dim3 grid((cols + 31) / 32), block(32, 8);
    PE_LAUNCH(colsum_kernel, grid, block, 0, x, ldx, out, rows, cols, accumulate);
    PE_LAUNCH_CHECK();
    return 0;
}

int pe_relu_bwd(const float* dy, int lddy, const float* y, int ldy, float* dz, int lddz, int rows, int cols,
                void* stream) {
    const long long n = (long long)rows * cols;
    if (n == 0) return 0;
    PE_LAUNCH(relu_bwd_kernel, grid_for(n, EW_THREADS), EW_THREADS, 0, dy, lddy, y, ldy, dz, lddz, rows, cols);
    PE_LAUNCH_CHECK();
    return 0;
}

int pe_fill(float* p, long long n, float value, void* stream) {
    if (n == 0) return 0;
    fill_kernel<<<grid_for(n, EW_THREADS * 4), EW_THREADS, 0, (cudaStream_t)stream>>>(p, n, value);
    PE_LAUNCH_CHECK();
    return 0;
}

int pe_add_i64(long long* p, long long n, long long value, void* stream) {
    if (n == 0) return 0;
    add_i64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p, n, value);
    PE_LAUNCH_CHECK();
    return 0;
}

int pe_maxpool3x3s2_fwd(const float* x, float* y, unsigned char* argmax, int B, int H, int W, int C, void* stream) {
    PE_REQUIRE(C % 4 == 0, "maxpool: C %% 4 != 0");
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const long long n = (long long)B * Ho * Wo * (C / 4);
    PE_LAUNCH(maxpool_fwd_kernel, one_wave_grid(maxpool_fwd_kernel, 0, n, EW_THREADS), EW_THREADS, 0, x, y, argmax, B,
              H, W, C, Ho, Wo);
    PE_LAUNCH_CHECK();
    return 0;
}

int pe_maxpool3x3s2_bwd(const float* dy, const float* dy2, const unsigned char* argmax, float* dx, int accumulate,
                        int B, int H, int W, int C, const float* aux_dout, int aux_lddo,
                        const unsigned char* aux_argmax, const float* aux_w, void* stream) {
    PE_REQUIRE(C % 4 == 0, "maxpool: C %% 4 != 0");
    PE_REQUIRE(!aux_dout || (aux_argmax && aux_w && H % 2 == 0 && W % 2 == 0),
               "maxpool_bwd: the aux term needs its arg-max map, weights and even H, W");
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const long long n = (long long)B * ((H + 1) / 2) * ((W + 1) / 2) * (C / 4);
    PE_LAUNCH(maxpool_bwd_kernel, one_wave_grid(maxpool_bwd_kernel, 0, n, EW_THREADS), EW_THREADS, 0, dy, dy2, argmax,
              dx, accumulate, B, H, W, C, Ho, Wo, aux_dout, aux_lddo, aux_argmax, aux_w);
    PE_LAUNCH_CHECK();
    return 0;
}

int pe_avgpool_fwd(const float* x, float* y, int ldy, int B, int HW, int C, int round_tf32, void* stream) {
    PE_REQUIRE(C % 4 == 0 && ldy % 4 == 0, "avgpool: C, ldy must be multiples of 4");
    const long long n = (long long)B * (C / 4);
    PE_LAUNCH(avgpool_fwd_kernel, (unsigned)((n + EW_THREADS - 1) / EW_THREADS), EW_THREADS, 0, x, y, ldy, B, HW, C,
              round_tf32);
    PE_LAUNCH_CHECK();
    return 0;
}

int pe_avgpool_bwd(const float* dy, int lddy, float* dx, int B, int HW, int C, void* stream) {
    PE_REQUIRE(C % 4 == 0 && lddy % 4 == 0, "avgpool: C, lddy must be multiples of 4");
    const long long n = (long long)B * HW * (C / 4);
    PE_LAUNCH(avgpool_bwd_kernel, one_wave_grid(avgpool_bwd_kernel, 0, n, EW_THREADS), EW_THREADS, 0, dy, lddy, dx, B,
              HW, C);
    PE_LAUNCH_CHECK();
    return 0;
}

int pe_aux_fwd(const float* a1, const float* w, const float* bias, float* out, int ldo, unsigned char* argmax, int B,
               int H, int W, int C, int round_tf32, void* stream) {
    PE_REQUIRE(H % 2 == 0 && W % 2 == 0 && C % 4 == 0, "aux: H, W must be even and C a multiple of 4");
    const long long nwin = (long long)B * (H / 2) * (W / 2);
    PE_LAUNCH(aux_fwd_kernel, one_wave_grid(aux_fwd_kernel, 0, nwin, EW_THREADS / 32 * 2), EW_THREADS, 0, a1, w, bias,
              out, ldo, argmax, B, H, W, C, round_tf32);
    PE_LAUNCH_CHECK();
    return 0;
}

int pe_aux_bwd(const float* dout, int lddo, const unsigned char* argmax, const float* a1, const float* w, float* da1,
               int accumulate, float* dw, float* db, int B, int H, int W, int C, void* stream) {
    PE_REQUIRE(C <= 256 && C % 4 == 0, "aux_bwd: C <= 256, C %% 4 == 0 required");
    const long long nwin = (long long)B * (H / 2) * (W / 2);
    PE_LAUNCH(aux_bwd_kernel, grid_for(nwin, EW_THREADS / 32 * 2 * 8, 8), EW_THREADS, 0, dout, lddo, argmax, a1, w,
              da1, accumulate, dw, db, B, H, W, C, nullptr, nullptr, 0);
    PE_LAUNCH_CHECK();
    return 0;
}

int pe_aux_bwd_params(const float* dout, int lddo, const unsigned char* argmax, const float* y, const float* scale,
                      const float* shift, int round_tf32, float* dw, float* db, int B, int H, int W, int C,
                      void* stream) {
    PE_REQUIRE(C <= 256 && C % 4 == 0, "aux_bwd_params: C <= 256, C %% 4 == 0 required");
    PE_REQUIRE(y && scale && shift && dw, "aux_bwd_params: y, scale, shift and dw are required");
    const long long nwin = (long long)B * (H / 2) * (W / 2);
    PE_LAUNCH(aux_bwd_kernel, grid_for(nwin, EW_THREADS / 32 * 2 * 8, 8), EW_THREADS, 0, dout, lddo, argmax, y,
              nullptr, nullptr, 0, dw, db, B, H, W, C, scale, shift, round_tf32);
    PE_LAUNCH_CHECK();
    return 0;
}

int pe_stem_post_train(const float* y, const double* stats, const float* gamma, const float* beta,
                       float* running_mean, float* running_var, long long* num_batches_tracked, float* scale,
                       float* shift, float* mean, float* invstd, float* pool, unsigned char* pool_argmax,
                       const float* aux_w, const float* aux_bias, float* aux_out, int ld_aux,
                       unsigned char* aux_argmax, int B, int H, int W, int C, float momentum, float eps,
                       int round_tf32, int aux_round_tf32, void* stream) {
    PE_REQUIRE(C == 64, "stem_post_train: the fused stem tail is written for 64 channels (got %d)", C);
    PE_REQUIRE(H % 2 == 0 && W % 2 == 0, "stem_post_train: H, W must be even");
    PE_REQUIRE(stats && scale && shift && mean && invstd && pool, "stem_post_train: stats / scale / shift / mean / "
               "invstd / pool required");
    PE_REQUIRE(!aux_w || (aux_bias && aux_out), "stem_post_train: aux_w needs aux_bias and aux_out");
    const long long nwin = (long long)B * (H / 2) * (W / 2);
    if (nwin == 0) return 0;
    PE_LAUNCH(stem_post_train_kernel, one_wave_grid(stem_post_train_kernel, 0, nwin, EW_THREADS / 32 * 2), EW_THREADS,
              0, y, stats, gamma, beta, running_mean, running_var, num_batches_tracked, scale, shift, mean, invstd,
              pool, pool_argmax, aux_w, aux_bias, aux_out, ld_aux, aux_argmax, B, H, W, momentum, eps, round_tf32,
              aux_round_tf32);
    PE_LAUNCH_CHECK();
    return 0;
}

int pe_depth_features_fwd(const float* depth, int B, int H, int W, int pool, const float* gamma, const float* beta,
                          float eps, float* xhat, float* aux, int ld, float* aux_pre, int round_tf32, void* stream) {
    PE_REQUIRE(pool >= 1 && H % pool == 0 && W % pool == 0, "depth_features: %dx%d not divisible by pool %d", H, W,
               pool);
    const int F = (H / pool) * (W / pool);
    PE_REQUIRE(F * sizeof(float) <= 48 * 1024, "depth_features: pooled map of %d values does not fit", F);
    if (B == 0) return 0;
    depth_features_fwd_kernel<<<B, EW_THREADS, F * sizeof(float), (cudaStream_t)stream>>>(
        depth, H, W, pool, gamma, beta, eps, xhat, aux, ld, aux_pre, round_tf32);
    PE_LAUNCH_CHECK();
    return 0;
}

int pe_depth_features_bwd(float* dprod, int ld, const float* aux_pre, const float* xhat, const float* gamma,
                          const float* beta, float* dgamma, float* dbeta, int B, int F, void* stream) {
    if (B == 0) return 0;
    depth_features_bwd_kernel<<<B, EW_THREADS, 0, (cudaStream_t)stream>>>(dprod, ld, aux_pre, xhat, gamma, beta,
                                                                          dgamma, dbeta, F);
    PE_LAUNCH_CHECK();
    return 0;
}

int pe_preprocess_u8(const unsigned char* src, float* dst, int B, int Hs, int Ws, int crop, const float* mean3,
                     const float* std3, void* stream) {
    PE_REQUIRE(crop <= Hs && crop <= Ws, "preprocess: crop larger than frame");
    PE_REQUIRE(crop % 4 == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
               "preprocess: crop must be a multiple of 4 and the output 16-byte aligned");
    const long long n = (long long)B * crop * (crop / 4);
    preprocess_u8_kernel<<<grid_for(n, EW_THREADS), EW_THREADS, 0, (cudaStream_t)stream>>>(
        src, dst, B, Hs, Ws, crop, mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2]);
    PE_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"

// tcgen05 / TMEM / TMA tap-GEMM kernel (see pe_tapgemm.cuh for the contract).
//
// Persistent, warp-specialised: one CTA per SM walks the (n-tile, m-tile, z) work list round-robin.
//   warp 0      TMA producer     : runs ahead through a TG_STAGES-deep smem ring, across tile boundaries
//   warp 1      MMA issuer       : one thread issues tcgen05.mma into one of TWO 128-column TMEM
//                                  accumulators, so tile i+1 accumulates while tile i is drained
//   warps 2..5  epilogue         : TMEM -> registers -> (affine / residual / ReLU / TF32 round / BN stats)
//                                  -> swizzled smem staging (double buffered) -> TMA store
#include "pe_tapgemm.cuh"

namespace pe {

namespace {

struct TileOrigin {
    int w0, h0, n0;
};

__device__ __forceinline__ TileOrigin tile_origin(const TapParams& p, int t) {
    TileOrigin o;
    int tw = t % p.tiles_w;
    int r = t / p.tiles_w;
    int th = r % p.tiles_h;
    int tn = r / p.tiles_h;
    o.w0 = tw * p.box_w;
    o.h0 = th * p.box_h;
    o.n0 = tn * p.box_n;
    return o;
}

struct Work {
    int nt, mt, z;       // n tile, m tile, z (conv: k split; wgrad: tap * ksplit + split)
    int k_begin, nk;     // k-step range
    int tap;             // wgrad: the filter tap of this work item
};

__device__ __forceinline__ Work decode_work(const TapParams& p, int w) {
    Work k;
    k.nt = w % p.work_n;
    const int r = w / p.work_n;
    k.mt = r % p.work_m;
    k.z = r / p.work_m;
    k.tap = 0;
    if (p.mode == 0) {
        const int ktotal = p.n_taps * p.chunks;
        const int per = (ktotal + p.ksplit - 1) / p.ksplit;
        k.k_begin = k.z * per;
        k.nk = max(0, min(ktotal, k.k_begin + per) - k.k_begin);
    } else {
        k.tap = k.z / p.ksplit;
        const int split = k.z - k.tap * p.ksplit;
        const int per = (p.pt_total + p.ksplit - 1) / p.ksplit;
        k.k_begin = split * per;
        k.nk = max(0, min(p.pt_total, k.k_begin + per) - k.k_begin);
    }
    return k;
}

__global__ void __launch_bounds__(TG_THREADS, 1)
tapgemm_kernel(const __grid_constant__ TapMaps maps, const __grid_constant__ TapParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_full[TG_STAGES];
    __shared__ __align__(8) uint64_t s_empty[TG_STAGES];
    __shared__ __align__(8) uint64_t s_tmem_full[2];
    __shared__ __align__(8) uint64_t s_tmem_empty[2];
    __shared__ uint32_t s_tmem_base;
    __shared__ float s_scale[TG_MAX_BN];
    __shared__ float s_shift[TG_MAX_BN];
    __shared__ float s_sum[TG_MAX_BN];
    __shared__ float s_sq[TG_MAX_BN];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // 1024-byte aligned tile storage (SWIZZLE_128B atoms are 1024 B)
    const uint32_t smem_base = smem_u32(smem_raw);
    const uint32_t stage_base = smem_base + p.stages * p.stage_bytes;   // TG_NSTAGE_OUT x 16 KB epilogue staging

    // ---- one-time setup --------------------------------------------------------------------
    if (threadIdx.x == 0) {
        for (int s = 0; s < TG_STAGES; ++s) {
            mbar_init(smem_u32(&s_full[s]), 1);
            mbar_init(smem_u32(&s_empty[s]), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(smem_u32(&s_tmem_full[a]), 1);
            mbar_init(smem_u32(&s_tmem_empty[a]), 128);
        }
        fence_mbar_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&maps.a[0]);
        tma_prefetch_desc(&maps.b[0]);
        if (p.store_mode == TG_STORE_TMA) tma_prefetch_desc(&maps.d);
    }
    if (warp == 1) {
        tmem_alloc(smem_u32(&s_tmem_base), 2 * TG_MAX_BN);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem_base;

    if (warp == 0) {
        // =============================== TMA producer ========================================
        if (lane == 0) {
            uint32_t it = 0;
            bool ok = true;
            for (int w = blockIdx.x; w < p.work_total && ok; w += gridDim.x) {
                const Work wk = decode_work(p, w);
                const int n_off = wk.nt * p.bn;
                if (p.mode == 0) {
                    const TileOrigin o = tile_origin(p, wk.mt);
                    const uint32_t bytes = static_cast<uint32_t>(p.m_rows + p.bn) * 128u;
                    for (int j = 0; j < wk.nk; ++j, ++it) {
                        const int s = it % p.stages;
                        const uint32_t ph = (it / p.stages) & 1;
                        if (!mbar_wait(smem_u32(&s_empty[s]), ph ^ 1)) {
                            atomicOr(p.error_flag, 1);
                            ok = false;
                            break;
                        }
                        const int k = wk.k_begin + j;
                        const int tap = k / p.chunks;
                        const int ch = k - tap * p.chunks;
                        const uint32_t full = smem_u32(&s_full[s]);
                        const uint32_t sa = smem_base + s * p.stage_bytes;
                        const uint32_t sb = sa + TG_A_BYTES;
                        uint32_t nbytes = bytes;
                        if (p.dbg_flags & 4) nbytes -= static_cast<uint32_t>(p.m_rows) * 128u;
                        if (p.dbg_flags & 8) nbytes -= static_cast<uint32_t>(p.bn) * 128u;
                        mbar_arrive_expect_tx(full, nbytes);
                        if (!(p.dbg_flags & 4))
                            tma_load_4d(sa, &maps.a[p.tap_map[tap]], full, ch * TG_BK, o.w0 + p.tap_dw[tap],
                                        o.h0 + p.tap_dh[tap], o.n0);
                        if (!(p.dbg_flags & 8)) tma_load_4d(sb, &maps.b[0], full, ch * TG_BK, n_off, p.tap_b[tap], 0);
                    }
                } else {
                    const int m_off = wk.mt * TG_BM;
                    const int nb = (p.bn + 31) >> 5;
                    const uint32_t bytes = static_cast<uint32_t>(4 + nb) * 4096u;
                    const int dw = p.tap_dw[wk.tap], dh = p.tap_dh[wk.tap];
                    const CUtensorMap* mb = &maps.b[p.tap_map[wk.tap]];
                    for (int j = 0; j < wk.nk; ++j, ++it) {
                        const int s = it % p.stages;
                        const uint32_t ph = (it / p.stages) & 1;
                        if (!mbar_wait(smem_u32(&s_empty[s]), ph ^ 1)) {
                            atomicOr(p.error_flag, 2);
                            ok = false;
                            break;
                        }
                        const TileOrigin o = tile_origin(p, wk.k_begin + j);
                        const uint32_t full = smem_u32(&s_full[s]);
                        const uint32_t sa = smem_base + s * p.stage_bytes;
                        const uint32_t sb = sa + TG_A_BYTES;
                        mbar_arrive_expect_tx(full, bytes);
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            tma_load_4d(sa + q * 4096, &maps.a[0], full, m_off + 32 * q, o.w0, o.h0, o.n0);
                        for (int q = 0; q < nb; ++q)
                            tma_load_4d(sb + q * 4096, mb, full, n_off + 32 * q, o.w0 + dw, o.h0 + dh, o.n0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // =============================== MMA issuer ==========================================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(TG_BM, p.bn, p.mode, p.mode);
            // K-major: LBO unused (16 B), SBO = 8 rows * 128 B.  MN-major: LBO = one 32-wide
            // atom column (32 k-rows * 128 B), SBO = 4 k-rows * 128 B (128B_BASE32B atoms).
            const uint32_t a_lbo = p.dbg_a_lbo >= 0 ? p.dbg_a_lbo : (p.mode ? 4096 : 16);
            const uint32_t a_sbo = p.dbg_a_sbo >= 0 ? p.dbg_a_sbo : (p.mode ? 512 : 1024);
            const uint32_t b_lbo = p.dbg_b_lbo >= 0 ? p.dbg_b_lbo : (p.mode ? 4096 : 16);
            const uint32_t b_sbo = p.dbg_b_sbo >= 0 ? p.dbg_b_sbo : (p.mode ? 512 : 1024);
            const uint32_t ltype = p.mode ? 1u : 2u;
            const uint32_t kstep_bytes = p.mode ? 1024u : 32u;  // 8 tf32 along K
            uint32_t it = 0, tile_i = 0;   // tile_i counts tiles that really use an accumulator
            bool ok = true;
            for (int w = blockIdx.x; w < p.work_total && ok; w += gridDim.x) {
                const Work wk = decode_work(p, w);
                if (wk.nk == 0) continue;
                const uint32_t acc = tile_i & 1, aph = (tile_i >> 1) & 1;
                ++tile_i;
                if (!mbar_wait(smem_u32(&s_tmem_empty[acc]), aph ^ 1)) {
                    atomicOr(p.error_flag, 16);
                    break;
                }
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * TG_MAX_BN;
                for (int j = 0; j < wk.nk; ++j, ++it) {
                    const int s = it % p.stages;
                    const uint32_t ph = (it / p.stages) & 1;
                    if (!mbar_wait(smem_u32(&s_full[s]), ph)) {
                        atomicOr(p.error_flag, 4);
                        ok = false;
                        break;
                    }
                    tc_fence_after();
                    const uint32_t sa = smem_base + s * p.stage_bytes;
                    const uint32_t sb = sa + TG_A_BYTES;
#pragma unroll
                    for (int kk = 0; kk < TG_BK / 8; ++kk) {
                        const uint64_t ad = make_smem_desc(sa + kk * kstep_bytes, a_lbo, a_sbo, ltype);
                        const uint64_t bd = make_smem_desc(sb + kk * kstep_bytes, b_lbo, b_sbo, ltype);
                        tc_mma_tf32(tmem_d, ad, bd, idesc, (j > 0 || kk > 0) ? 1u : 0u);
                    }
                    tc_commit(smem_u32(&s_empty[s]));  // frees the stage when these MMAs retire
                }
                if (ok) tc_commit(smem_u32(&s_tmem_full[acc]));
            }
        }
        __syncwarp();
    } else {
        // =============================== epilogue ============================================
        const int q = warp & 3;            // TMEM lane quarter this warp may read
        const int row = q * 32 + lane;     // tile row == TMEM lane
        const int et = threadIdx.x - 64;   // 0..127
        uint32_t tile_i = 0, cc = 0;
        bool ok = true;
        for (int w = blockIdx.x; w < p.work_total && ok; w += gridDim.x) {
            const Work wk = decode_work(p, w);
            const int n_off = wk.nt * p.bn;
            const uint32_t acc = tile_i & 1, aph = (tile_i >> 1) & 1;
            if (wk.nk > 0) ++tile_i;

            // per-tile column constants (everyone is past the previous tile's reads after this barrier)
            asm volatile("bar.sync 1, 128;" ::: "memory");
            for (int cidx = et; cidx < p.bn; cidx += 128) {
                const int col = n_off + cidx;
                float sc = 1.f, sh = 0.f;
                if (col < p.n_total) {
                    if (p.scale) {
                        sc = p.scale[col];
                        sh = p.shift[col];
                    } else if (p.bias) {
                        sh = p.bias[col];
                    }
                }
                s_scale[cidx] = sc;
                s_shift[cidx] = sh;
                s_sum[cidx] = 0.f;
                s_sq[cidx] = 0.f;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");

            if (wk.nk > 0) {
                ok = mbar_wait(smem_u32(&s_tmem_full[acc]), aph);
                if (!ok) {
                    atomicOr(p.error_flag, 8);
                    break;
                }
            }
            tc_fence_after();

            // row -> output coordinates
            bool row_valid;
            long long row_lin;  // dense row index of the output (pixel index, or M row for wgrad)
            TileOrigin o = {0, 0, 0};
            if (p.mode == 0) {
                o = tile_origin(p, wk.mt);
                const int dw = row % p.box_w;
                const int r2 = row / p.box_w;
                const int dh = r2 % p.box_h;
                const int dn = r2 / p.box_h;
                const int wq = o.w0 + dw, hq = o.h0 + dh, nq = o.n0 + dn;
                row_valid = (row < p.m_rows) && (wq < p.out_w) && (hq < p.out_h) && (nq < p.out_n);
                row_lin = (static_cast<long long>(nq) * p.out_h + hq) * p.out_w + wq;
            } else {
                const int m = wk.mt * TG_BM + row;
                row_valid = m < p.m_total;
                row_lin = m;
            }
            float* out_row = nullptr;
            if (p.store_mode != TG_STORE_TMA)
                out_row = p.out + (p.mode ? wk.tap * p.out_tap_stride : 0ll) + row_lin * p.ldo;
            const float* res_row = p.residual ? p.residual + row_lin * p.ld_res : nullptr;

            const int nchunks = (p.bn + 31) >> 5;
            for (int c = 0; c < nchunks; ++c) {
                float v[32];
                if (wk.nk > 0) {
                    tmem_ld_32x32(tmem_base + acc * TG_MAX_BN + (static_cast<uint32_t>(q * 32) << 16) + c * 32, v);
                    if (c == nchunks - 1) {   // accumulator fully read: hand it back to the MMA warp
                        tc_fence_before();
                        mbar_arrive(smem_u32(&s_tmem_empty[acc]));
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = 0.f;
                }
                const int col0 = n_off + c * 32;
                if (!row_valid) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = 0.f;
                }
                // ---- per-channel batch statistics of the raw accumulator -----------------------
                if (p.stats) {
                    float a[32], b[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        a[i] = v[i];
                        b[i] = v[i] * v[i];
                    }
#pragma unroll
                    for (int off = 16; off >= 1; off >>= 1) {
                        const bool up = (lane & off) != 0;
#pragma unroll
                        for (int i = 0; i < off; ++i) {
                            const float sa_ = up ? a[i] : a[i + off];
                            const float ka_ = up ? a[i + off] : a[i];
                            a[i] = ka_ + __shfl_xor_sync(0xffffffffu, sa_, off);
                            const float sb_ = up ? b[i] : b[i + off];
                            const float kb_ = up ? b[i + off] : b[i];
                            b[i] = kb_ + __shfl_xor_sync(0xffffffffu, sb_, off);
                        }
                    }
                    atomicAdd(&s_sum[c * 32 + lane], a[0]);
                    atomicAdd(&s_sq[c * 32 + lane], b[0]);
                }
                // ---- affine / residual / activation --------------------------------------------
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = fmaf(v[i], s_scale[c * 32 + i], s_shift[c * 32 + i]);
                if (res_row && row_valid) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        if (col0 + i < p.n_total) {
                            const float4 r4 = *reinterpret_cast<const float4*>(res_row + col0 + i);
                            v[i] += r4.x;
                            v[i + 1] += r4.y;
                            v[i + 2] += r4.z;
                            v[i + 3] += r4.w;
                        }
                    }
                }
                if (p.relu) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
                }
                if (p.round_out) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = round_tf32(v[i]);
                }
                // ---- store ---------------------------------------------------------------------
                if (p.store_mode == TG_STORE_TMA) {
                    const uint32_t region = stage_base + (cc % p.nout) * TG_A_BYTES;
                    // the store issued from this buffer TG_NSTAGE_OUT chunks ago must have finished reading it
                    if (et == 0) {
                        if (p.nout >= 4) asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
                        else if (p.nout == 3) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
                        else if (p.nout == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                        else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    }
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    if (!(p.dbg_flags & 2)) {
                    const uint32_t rbase = region + row * 128;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint32_t addr = rbase + ((j ^ (row & 7)) << 4);
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v[4 * j]),
                                     "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3])
                                     : "memory");
                    }
                    }
                    fence_proxy_async_smem();
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    if (et == 0 && !(p.dbg_flags & 3)) {
                        tma_store_4d(&maps.d, region, col0, o.w0, o.h0, o.n0);
                        tma_store_commit();
                    }
                    ++cc;
                } else if (row_valid) {
                    if (p.store_mode == TG_STORE_DIRECT) {
                        if (col0 + 32 <= p.n_total && (p.ldo & 3) == 0) {
#pragma unroll
                            for (int i = 0; i < 32; i += 4)
                                *reinterpret_cast<float4*>(out_row + col0 + i) =
                                    make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                if (col0 + i < p.n_total) out_row[col0 + i] = v[i];
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (col0 + i < p.n_total) atomicAdd(out_row + col0 + i, v[i]);
                    }
                }
            }
            if (p.stats) {
                asm volatile("bar.sync 1, 128;" ::: "memory");
                for (int cidx = et; cidx < p.bn; cidx += 128) {
                    const int col = n_off + cidx;
                    if (col < p.n_total) {
                        atomicAdd(p.stats + col, static_cast<double>(s_sum[cidx]));
                        atomicAdd(p.stats + p.n_total + col, static_cast<double>(s_sq[cidx]));
                    }
                }
            }
        }
        if (p.store_mode == TG_STORE_TMA && et == 0) tma_store_wait_all();
    }

    // ---- teardown ---------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * TG_MAX_BN);
    }
}

}  // namespace

int launch_tapgemm(const TapMaps& maps, TapParams& p, dim3 work, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        PE_CHECK_CUDA(cudaFuncSetAttribute(tapgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           TG_SMEM_BYTES));
        configured = true;
    }
    p.work_n = work.x;
    p.work_m = work.y;
    p.work_total = static_cast<int>(work.x * work.y * work.z);
    int grid = p.work_total < num_sms() ? p.work_total : num_sms();
    if (grid < 1) return 0;
    tapgemm_kernel<<<grid, TG_THREADS, TG_SMEM_BYTES, stream>>>(maps, p);
    PE_LAUNCH_CHECK();
    return 0;
}

}  // namespace pe

// tcgen05 / TMEM / TMA tap-GEMM kernel (see pe_tapgemm.cuh for the contract).
//
// Persistent, warp-specialised: one CTA per SM walks the (n-tile, m-tile, z) work list round-robin.
//   warp 0        TMA producer   : runs ahead through a 2-8 stage smem ring, across tile boundaries (the ring
//                                  depth is what bounds narrow tiles: DESIGN 3.1)
//   warp 1        MMA issuer     : tcgen05.mma into one of TWO 256-column TMEM accumulators, so tile i+1
//                                  accumulates while tile i is drained (mode 2: all 512 columns hold one set of
//                                  per-tap accumulators)
//   warps 2..     epilogue       : G groups of four warps (G = 2, or 4 for store-bound launches); group g drains the
//                                  32-column chunks c with c % G == g:  TMEM -> registers -> (affine / TMA-prefetched
//                                  residual / ReLU / TF32 round) -> swizzled smem staging -> TMA store, plus the
//                                  per-channel batch statistics as a column pass over the staged tile
// Warps 0 and 1 run warp-convergent (role dispatch on a shfl-broadcast warp index, vote-derived barrier results,
// elect.sync inside the issuing asm) so that descriptors, coordinates and counters stay on the uniform datapath.
#include <stdlib.h>

#include "pe_tapgemm.cuh"

namespace pe {

namespace {

struct TileOrigin {
    int w0, h0, n0;
};

// n / d for a divisor prepared by launch_tapgemm (TapParams::fd_*), 0 <= n < 2^31
__device__ __forceinline__ int fast_div(int n, const unsigned (&fd)[2]) {
    return fd[0] ? static_cast<int>(__umulhi(static_cast<unsigned>(n), fd[0]) >> fd[1]) : n;
}

__device__ __forceinline__ TileOrigin tile_origin(const TapParams& p, int t) {
    TileOrigin o;
    const int r = fast_div(t, p.fd_tiles_w);
    const int tw = t - r * p.tiles_w;
    const int tn = fast_div(r, p.fd_tiles_h);
    const int th = r - tn * p.tiles_h;
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
    const int r = fast_div(w, p.fd_work_n);
    k.nt = w - r * p.work_n;
    k.z = fast_div(r, p.fd_work_m);
    k.mt = r - k.z * p.work_m;
    k.tap = 0;
    if (p.mode == 0) {
        const int ktotal = p.n_taps * p.chunks;
        const int per = p.k_per;
        k.k_begin = k.z * per;
        k.nk = max(0, min(ktotal, k.k_begin + per) - k.k_begin);
    } else if (p.mode == 1) {
        // taps vary fastest: CTAs running side by side reduce the same pixel range for different taps,
        // so the dY / X tiles they share are served from L2 instead of being re-read from HBM per tap
        k.tap = k.z % p.n_taps;
        const int split = k.z / p.n_taps;
        const int per = p.k_per;
        k.k_begin = split * per;
        k.nk = max(0, min(p.pt_total, k.k_begin + per) - k.k_begin);
    } else {
        // haloed wgrad: z = split * n_groups + tap group; `tap` is the first tap of the group
        k.tap = (k.z % p.n_groups) * p.tg_taps;
        const int split = k.z / p.n_groups;
        const int per = p.k_per;
        k.k_begin = split * per;
        k.nk = max(0, min(p.pt_total, k.k_begin + per) - k.k_begin);
    }
    return k;
}

// One stage of the haloed wgrad: BH tile rows x NTG taps, one MMA each (K = 8 pixels of one tile row).  Rows outer,
// taps inner: consecutive MMAs target different accumulators, and every descriptor is base + constant.
// ntg / bh are the runtime bounds (equal to NTG / BH in the straight-line specialisations).
template <int NTG, int BH>
__device__ __forceinline__ void issue_halo_stage(uint32_t tmem_d, uint32_t bn, uint32_t a_hi, uint32_t a_lo,
                                                 uint32_t b_hi, uint32_t b_lo, uint32_t row16,
                                                 const uint32_t (&tap_off)[8], uint32_t idesc, uint32_t first,
                                                 int ntg, int bh) {
#pragma unroll
    for (int hh = 0; hh < BH; ++hh) {
        if (hh < bh) {
            const uint64_t ad = (static_cast<uint64_t>(a_hi) << 32) | (a_lo + hh * 64u);
            const uint32_t b_row = b_lo + hh * row16;
#pragma unroll
            for (int tl = 0; tl < NTG; ++tl) {
                if (tl < ntg) {
                    const uint64_t bd = (static_cast<uint64_t>(b_hi) << 32) | (b_row + tap_off[tl]);
                    tc_mma_tf32_elect(tmem_d + tl * bn, ad, bd, idesc, (first | (hh > 0 ? 1u : 0u)));
                }
            }
        }
    }
}

// Work item of this CTA: with CTA pairs the list counts PAIRS of consecutive 128-pixel tiles, CTA rank r owns tile
// 2 * pair + r (a phantom tile past the end loads zero-filled boxes and stores nothing).
template <int CG>
__device__ __forceinline__ Work decode_work_cg(const TapParams& p, int w, uint32_t rank) {
    Work k = decode_work(p, w);
    if (CG == 2) k.mt = k.mt * 2 + static_cast<int>(rank);
    return k;
}

template <int G, int CG>      // epilogue groups (2 or 4), CTAs per MMA (1, or 2 = tcgen05 cta_group::2 pairs)
__global__ void __launch_bounds__(64 + 128 * G, 1)
tapgemm_kernel(const __grid_constant__ TapMaps maps, const __grid_constant__ TapParams p) {
    constexpr int EPI_THREADS = 128 * G;
    const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0u;
    // persistent walk over the work list: one CTA (or one pair) per list position
    const int w_first = CG == 2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    const int w_step = CG == 2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_full[TG_STAGES];
    __shared__ __align__(8) uint64_t s_empty[TG_STAGES];
    __shared__ __align__(8) uint64_t s_bfull[8];      // weight ring of the haloed conv (conv_halo)
    __shared__ __align__(8) uint64_t s_bempty[8];
    __shared__ __align__(8) uint64_t s_tmem_full[2];
    __shared__ __align__(8) uint64_t s_tmem_empty[2];
    __shared__ __align__(8) uint64_t s_res_full[2 * G];
    __shared__ uint32_t s_tmem_base;
    __shared__ __align__(16) float s_scale[TG_MAX_BN];
    __shared__ __align__(16) float s_shift[TG_MAX_BN];

    // broadcast from lane 0: lets ptxas prove the role dispatch below is warp-uniform, which is what allows the
    // producer / MMA warps to keep their addresses and descriptors on the uniform datapath
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;

    // 1024-byte aligned tile storage (SWIZZLE_128B atoms are 1024 B)
    const uint32_t smem_base = smem_u32(smem_raw);
    const uint32_t bring_base = smem_base + p.stages * p.stage_bytes;   // weight ring of the haloed conv (or empty)
    const uint32_t stage_base = bring_base + p.b_ring_bytes;            // nout x 16 KB epilogue staging
    // CTA-wide per-channel statistics (one global atomic per channel per CTA instead of per tile)
    const uint32_t res_base = stage_base + p.nout * TG_A_BYTES;          // nres x 16 KB residual tiles
    float* s_sum = reinterpret_cast<float*>(smem_raw + p.stages * p.stage_bytes + p.b_ring_bytes +
                                            (p.nout + p.nres) * TG_A_BYTES);
    float* s_sq = s_sum + p.stats_cols;
    // per-tile scratch of the statistics: [group][chunk slot][warp 4][sum 32 | sq 32] (8 KB, only with stats)
    float* s_part = s_sum + 2 * p.stats_cols;
    // bn_bwd mode: batch mean / 1/std of the tile's output columns (2 x TG_MAX_BN floats behind the 8 KB scratch)
    float* s_mean = s_part + 2048;
    float* s_istd = s_mean + TG_MAX_BN;

    // ---- one-time setup --------------------------------------------------------------------
    if (threadIdx.x == 0) {
        for (int s = 0; s < TG_STAGES; ++s) {
            mbar_init(smem_u32(&s_full[s]), 1);
            mbar_init(smem_u32(&s_empty[s]), 1);
        }
        for (int s = 0; s < 8; ++s) {
            mbar_init(smem_u32(&s_bfull[s]), 1);
            mbar_init(smem_u32(&s_bempty[s]), 1);
        }
        for (int a = 0; a < 2 * G; ++a) mbar_init(smem_u32(&s_res_full[a]), 1);
        for (int a = 0; a < 2; ++a) {
            mbar_init(smem_u32(&s_tmem_full[a]), 1);
            // pairs: the leader's MMA warp reuses an accumulator once BOTH CTAs' epilogues have drained it
            mbar_init(smem_u32(&s_tmem_empty[a]), EPI_THREADS * CG);
        }
        fence_mbar_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&maps.a[0]);
        tma_prefetch_desc(&maps.b[0]);
        if (p.store_mode == TG_STORE_TMA) tma_prefetch_desc(&maps.d);
        if (p.nres) tma_prefetch_desc(&maps.r);
    }
    if (warp == 1) {
        if (CG == 2) {
            tmem_alloc_pair(smem_u32(&s_tmem_base), 2 * TG_MAX_BN);
            tmem_relinquish_pair();
        } else {
            tmem_alloc(smem_u32(&s_tmem_base), 2 * TG_MAX_BN);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();      // the peer's barriers are initialised before anything is signalled on them
    tc_fence_after();
    const uint32_t tmem_base = s_tmem_base;
    // Programmatic dependent launch: everything above (barrier init, TMEM allocation, tensor-map prefetch) touched no
    // global memory, so it may overlap the tail of the previous kernel in the stream.  Let the NEXT kernel start its
    // own prologue now, then wait until the previous grid has completed and flushed before reading its outputs.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int spin = (p.dbg_flags & 128) ? 1 : 0;   // probe: non-blocking barrier polls in the producer / MMA warps

    if (warp == 0) {
        // =============================== TMA producer ========================================
        // The whole warp runs this loop convergently; every barrier result is vote-derived, so coordinates,
        // stage counters and addresses stay in uniform registers and one elected lane issues each TMA.
        uint32_t s = 0, ph = 0, sb = 0, phb = 0;
        bool ok = true;
        for (int w = w_first; w < p.work_total && ok; w += w_step) {
            const Work wk = decode_work_cg<CG>(p, w, cta_rank);
            const int n_off = wk.nt * p.bn;
            if (CG == 2) {
                // CTA pair: this CTA's own 128-pixel A box + its half of the weight tile; the bytes of both CTAs are
                // counted on the leader's barrier, which the leader arms for the pair
                const TileOrigin o = tile_origin(p, wk.mt);
                const int bhalf = p.bn >> 1;
                const uint32_t nbytes = static_cast<uint32_t>(p.m_rows + bhalf) * 128u;
                const int n_mine = n_off + static_cast<int>(cta_rank) * bhalf;
                int tap = wk.k_begin / p.chunks;
                int ch = wk.k_begin - tap * p.chunks;
                for (int j = 0; j < wk.nk; ++j) {
                    if (!mbar_wait_warp(smem_u32(&s_empty[s]), ph ^ 1, spin)) {
                        atomicOr(p.error_flag, 1);
                        ok = false;
                        break;
                    }
                    const uint32_t full = smem_u32(&s_full[s]);
                    const uint32_t sa = smem_base + s * p.stage_bytes;
                    if (leader) mbar_arrive_expect_tx_elect(full, 2u * nbytes);
                    tma_load_4d_pair_elect(sa, &maps.a[p.tap_map[tap]], full, ch * TG_BK, o.w0 + p.tap_dw[tap],
                                           o.h0 + p.tap_dh[tap], o.n0);
                    tma_load_4d_pair_elect(sa + TG_A_BYTES, &maps.b[0], full, ch * TG_BK, n_mine, p.tap_b[tap], 0);
                    if (++ch == p.chunks) {
                        ch = 0;
                        ++tap;
                    }
                    if (++s == static_cast<uint32_t>(p.stages)) {
                        s = 0;
                        ph ^= 1;
                    }
                }
            } else if (p.mode == 0 && p.conv_halo) {
                // haloed stride-1 conv: one haloed A tile per channel chunk, then one weight tile per tap
                const TileOrigin o = tile_origin(p, wk.mt);
                const uint32_t a_bytes = static_cast<uint32_t>(p.halo_w * p.halo_h) * 128u;
                const uint32_t b_bytes = static_cast<uint32_t>(p.bn * p.b_taps) * 128u;
                for (int ch = 0; ch < p.chunks && ok; ++ch) {
                    if (!mbar_wait_warp(smem_u32(&s_empty[s]), ph ^ 1, spin)) {
                        atomicOr(p.error_flag, 1);
                        ok = false;
                        break;
                    }
                    // (probe flags 4 / 8: the A / B loads are skipped -- timing experiments on garbage operands)
                    mbar_arrive_expect_tx_elect(smem_u32(&s_full[s]), (p.dbg_flags & 4) ? 0u : a_bytes);
                    if (!(p.dbg_flags & 4))
                        tma_load_4d_elect(smem_base + s * p.stage_bytes, &maps.a[0], smem_u32(&s_full[s]), ch * TG_BK,
                                          o.w0 + p.halo_dw, o.h0 + p.halo_dh, o.n0);
                    if (++s == static_cast<uint32_t>(p.stages)) {
                        s = 0;
                        ph ^= 1;
                    }
                    for (int t = 0; t < p.n_taps && !(p.dbg_flags & 8); t += p.b_taps) {
                        if (!mbar_wait_warp(smem_u32(&s_bempty[sb]), phb ^ 1, spin)) {
                            atomicOr(p.error_flag, 1);
                            ok = false;
                            break;
                        }
                        mbar_arrive_expect_tx_elect(smem_u32(&s_bfull[sb]), b_bytes);
                        // one box {32 ch, bn rows, b_taps taps}: the weight tiles of b_taps consecutive taps
                        tma_load_4d_elect(bring_base + sb * p.b_bytes, &maps.b[1], smem_u32(&s_bfull[sb]), ch * TG_BK,
                                          n_off, t, 0);
                        if (++sb == static_cast<uint32_t>(p.b_stages)) {
                            sb = 0;
                            phb ^= 1;
                        }
                    }
                }
            } else if (p.mode == 0) {
                const TileOrigin o = tile_origin(p, wk.mt);
                uint32_t nbytes = static_cast<uint32_t>(p.m_rows + p.bn) * 128u;
                if (p.dbg_flags & 4) nbytes -= static_cast<uint32_t>(p.m_rows) * 128u;
                if (p.dbg_flags & 8) nbytes -= static_cast<uint32_t>(p.bn) * 128u;
                int tap = wk.k_begin / p.chunks;
                int ch = wk.k_begin - tap * p.chunks;
                for (int j = 0; j < wk.nk; ++j) {
                    if (!mbar_wait_warp(smem_u32(&s_empty[s]), ph ^ 1, spin)) {
                        atomicOr(p.error_flag, 1);
                        ok = false;
                        break;
                    }
                    const uint32_t full = smem_u32(&s_full[s]);
                    const uint32_t sa = smem_base + s * p.stage_bytes;
                    const uint32_t sb = sa + TG_A_BYTES;
                    mbar_arrive_expect_tx_elect(full, nbytes);
                    if (!(p.dbg_flags & 4))
                        tma_load_4d_elect(sa, &maps.a[p.tap_map[tap]], full, ch * TG_BK, o.w0 + p.tap_dw[tap],
                                          o.h0 + p.tap_dh[tap], o.n0);
                    if (!(p.dbg_flags & 8))
                        tma_load_4d_elect(sb, &maps.b[0], full, ch * TG_BK, n_off, p.tap_b[tap], 0);
                    if (++ch == p.chunks) {
                        ch = 0;
                        ++tap;
                    }
                    if (++s == static_cast<uint32_t>(p.stages)) {
                        s = 0;
                        ph ^= 1;
                    }
                }
            } else if (p.mode == 2) {
                // haloed wgrad: one dY pixel tile + one X tile with an (R-1, S-1) halo per stage; every tap of
                // the group reads its shifted window of the SAME X tile through its own UMMA descriptor
                const int m_off = wk.mt * TG_BM;
                const int na = min(4, (p.m_total - m_off + 31) >> 5);
                const int nb = min((p.bn + 31) >> 5, (p.n_total - n_off + 31) >> 5);
                const uint32_t a_bytes = static_cast<uint32_t>(p.box_w * p.box_h * p.box_n) * 128u;
                const uint32_t b_bytes = static_cast<uint32_t>(p.halo_w * p.halo_h * p.box_n) * 128u;
                const uint32_t bytes = na * a_bytes + nb * b_bytes;
                // running pixel-tile coordinates (w fastest)
                int tw = wk.k_begin % p.tiles_w;
                const int r0 = wk.k_begin / p.tiles_w;
                int th = r0 % p.tiles_h;
                int tn = r0 / p.tiles_h;
                for (int j = 0; j < wk.nk; ++j) {
                    if (!mbar_wait_warp(smem_u32(&s_empty[s]), ph ^ 1, spin)) {
                        atomicOr(p.error_flag, 2);
                        ok = false;
                        break;
                    }
                    const int w0 = tw * p.box_w, h0 = th * p.box_h, n0 = tn * p.box_n;
                    const uint32_t full = smem_u32(&s_full[s]);
                    const uint32_t sa = smem_base + s * p.stage_bytes;
                    const uint32_t sb = sa + p.a_region_bytes;
                    uint32_t nbytes = bytes;
                    if (p.dbg_flags & 4) nbytes -= na * a_bytes;
                    if (p.dbg_flags & 8) nbytes -= nb * b_bytes;
                    mbar_arrive_expect_tx_elect(full, nbytes);
                    if (!(p.dbg_flags & 4))
                        for (int q = 0; q < na; ++q)
                            tma_load_4d_elect(sa + q * p.a_atom_bytes, &maps.a[0], full, m_off + 32 * q, w0, h0, n0);
                    if (!(p.dbg_flags & 8))
                        for (int q = 0; q < nb; ++q)
                            tma_load_4d_elect(sb + q * p.b_atom_bytes, &maps.b[0], full, n_off + 32 * q,
                                              w0 + p.halo_dw, h0 + p.halo_dh, n0);
                    if (++tw == p.tiles_w) {
                        tw = 0;
                        if (++th == p.tiles_h) {
                            th = 0;
                            ++tn;
                        }
                    }
                    if (++s == static_cast<uint32_t>(p.stages)) {
                        s = 0;
                        ph ^= 1;
                    }
                }
            } else {
                const int m_off = wk.mt * TG_BM;
                // only the 32-wide atoms that hold real channels are fetched; the MMA reads the rest of the
                // stage as don't-care rows / columns of the accumulator
                const int na = min(4, (p.m_total - m_off + 31) >> 5);
                const int nb = min((p.bn + 31) >> 5, (p.n_total - n_off + 31) >> 5);
                const uint32_t bytes = static_cast<uint32_t>(na + nb) * 4096u;
                const int dw = p.tap_dw[wk.tap], dh = p.tap_dh[wk.tap];
                const CUtensorMap* mb = &maps.b[p.tap_map[wk.tap]];
                int tw = wk.k_begin % p.tiles_w;
                const int r0 = wk.k_begin / p.tiles_w;
                int th = r0 % p.tiles_h;
                int tn = r0 / p.tiles_h;
                for (int j = 0; j < wk.nk; ++j) {
                    if (!mbar_wait_warp(smem_u32(&s_empty[s]), ph ^ 1, spin)) {
                        atomicOr(p.error_flag, 2);
                        ok = false;
                        break;
                    }
                    const int w0 = tw * p.box_w, h0 = th * p.box_h, n0 = tn * p.box_n;
                    const uint32_t full = smem_u32(&s_full[s]);
                    const uint32_t sa = smem_base + s * p.stage_bytes;
                    const uint32_t sb = sa + TG_A_BYTES;
                    mbar_arrive_expect_tx_elect(full, bytes);
                    for (int q = 0; q < na; ++q)
                        tma_load_4d_elect(sa + q * 4096, &maps.a[0], full, m_off + 32 * q, w0, h0, n0);
                    for (int q = 0; q < nb; ++q)
                        tma_load_4d_elect(sb + q * 4096, mb, full, n_off + 32 * q, w0 + dw, h0 + dh, n0);
                    if (++tw == p.tiles_w) {
                        tw = 0;
                        if (++th == p.tiles_h) {
                            th = 0;
                            ++tn;
                        }
                    }
                    if (++s == static_cast<uint32_t>(p.stages)) {
                        s = 0;
                        ph ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // =============================== MMA issuer ==========================================
        // Convergent warp, uniform descriptors, one elected lane per tcgen05.mma / tcgen05.commit.
        // (m64: a weight gradient with at most 64 output channels issues M = 64 instructions -- they retire in 32 clk at 64
        //  columns where the zero-padded M = 128 form takes 48, and read half the A rows; tests/probe_mma_rate.cu)
        const uint32_t idesc = make_idesc_tf32(p.m64 ? 64 : TG_BM * CG, p.bn, p.mode != 0, p.mode != 0);
        // K-major: LBO unused (16 B), SBO = 8 rows * 128 B.  MN-major: LBO = one 32-wide
        // atom column (32 k-rows * 128 B), SBO = 4 k-rows * 128 B (128B_BASE32B atoms).
        const uint32_t a_lbo = p.dbg_a_lbo >= 0 ? p.dbg_a_lbo : (p.mode ? 4096 : 16);
        const uint32_t a_sbo = p.dbg_a_sbo >= 0 ? p.dbg_a_sbo : (p.mode ? 512 : 1024);
        const uint32_t b_lbo = p.dbg_b_lbo >= 0 ? p.dbg_b_lbo : (p.mode ? 4096 : 16);
        const uint32_t b_sbo = p.dbg_b_sbo >= 0 ? p.dbg_b_sbo : (p.mode ? 512 : 1024);
        const uint32_t ltype = p.mode ? 1u : 2u;
        // descriptors are built once for smem address 0 and advanced by (bytes >> 4): the 14-bit address field
        // cannot carry because shared memory ends below 256 KB
        const uint64_t a_desc0 = p.mode == 2 ? make_smem_desc(0, p.a_atom_bytes, 512, 1u)
                                             : make_smem_desc(0, a_lbo, a_sbo, ltype);
        const uint64_t b_desc0 = p.mode == 2 ? make_smem_desc(0, p.b_atom_bytes, 512, 1u)
                                             : make_smem_desc(0, b_lbo, b_sbo, ltype);
        const uint32_t kstep16 = (p.mode ? 1024u : 32u) >> 4;  // 8 tf32 along K, in 16-byte units
        uint32_t s = 0, ph = 0, sb = 0, phb = 0, tile_i = 0;   // tile_i counts tiles that really use an accumulator
        bool ok = true;
        // pairs: only the leader CTA issues; its MMAs read both CTAs' stages and write both CTAs' accumulators
        for (int w = w_first; w < p.work_total && ok && (CG == 1 || leader); w += w_step) {
            const Work wk = decode_work_cg<CG>(p, w, cta_rank);
            if (wk.nk == 0) continue;
            // mode 2 owns all 512 TMEM columns as ONE set of per-tap accumulators
            const uint32_t acc = p.mode == 2 ? 0u : (tile_i & 1);
            const uint32_t aph = p.mode == 2 ? (tile_i & 1) : ((tile_i >> 1) & 1);
            ++tile_i;
            if (!mbar_wait_warp(smem_u32(&s_tmem_empty[acc]), aph ^ 1)) {
                atomicOr(p.error_flag, 16);
                break;
            }
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * TG_MAX_BN;
            if (p.mode == 2) {
                const int ntg = min(p.tg_taps, p.n_taps - wk.tap);
                // Issue loop kept free of loop-carried address arithmetic: every descriptor is base + table entry,
                // rows outer / taps inner so consecutive MMAs hit different accumulators.  (box_w == 8, box_n == 1:
                // one 8-pixel K segment per tile row.)
                uint32_t tap_off[8];
#pragma unroll
                for (int tl = 0; tl < 8; ++tl) {
                    const int tap = min(wk.tap + tl, p.n_taps - 1);
                    tap_off[tl] = static_cast<uint32_t>(p.tap_dh[tap] * p.halo_w + p.tap_dw[tap]) * 8u;
                }
                const uint32_t row16 = static_cast<uint32_t>(p.halo_w) * 8u;      // one halo row, 16-byte units
                const uint32_t a_hi = static_cast<uint32_t>(a_desc0 >> 32), b_hi = static_cast<uint32_t>(b_desc0 >> 32);
                const uint32_t a_lo0 = static_cast<uint32_t>(a_desc0), b_lo0 = static_cast<uint32_t>(b_desc0);
                for (int j = 0; j < wk.nk; ++j) {
                    if (!mbar_wait_warp(smem_u32(&s_full[s]), ph, spin)) {
                        atomicOr(p.error_flag, 4);
                        ok = false;
                        break;
                    }
                    tc_fence_after();
                    const uint32_t sa = smem_base + s * p.stage_bytes;
                    const uint32_t a_lo = a_lo0 + (sa >> 4);
                    const uint32_t b_lo = b_lo0 + ((sa + p.a_region_bytes) >> 4);
                    if (!(p.dbg_flags & 16)) {
                        const uint32_t first = j > 0 ? 1u : 0u;
                        const int key = ntg * 16 + p.box_h;
                        // straight-line specialisations for the shapes ResNet-50 produces (3x3: groups of 5 / 4 / 3
                        // taps; 8- or 7-row tiles); anything else takes the predicated generic loop
                        switch (key) {
                            case 5 * 16 + 8: issue_halo_stage<5, 8>(tmem_d, p.bn, a_hi, a_lo, b_hi, b_lo, row16, tap_off, idesc, first, 5, 8); break;
                            case 4 * 16 + 8: issue_halo_stage<4, 8>(tmem_d, p.bn, a_hi, a_lo, b_hi, b_lo, row16, tap_off, idesc, first, 4, 8); break;
                            case 3 * 16 + 8: issue_halo_stage<3, 8>(tmem_d, p.bn, a_hi, a_lo, b_hi, b_lo, row16, tap_off, idesc, first, 3, 8); break;
                            case 5 * 16 + 7: issue_halo_stage<5, 7>(tmem_d, p.bn, a_hi, a_lo, b_hi, b_lo, row16, tap_off, idesc, first, 5, 7); break;
                            case 4 * 16 + 7: issue_halo_stage<4, 7>(tmem_d, p.bn, a_hi, a_lo, b_hi, b_lo, row16, tap_off, idesc, first, 4, 7); break;
                            case 3 * 16 + 7: issue_halo_stage<3, 7>(tmem_d, p.bn, a_hi, a_lo, b_hi, b_lo, row16, tap_off, idesc, first, 3, 7); break;
                            default: issue_halo_stage<8, 8>(tmem_d, p.bn, a_hi, a_lo, b_hi, b_lo, row16, tap_off, idesc, first, ntg, p.box_h); break;
                        }
                    }
                    tc_commit_elect(smem_u32(&s_empty[s]));
                    if (++s == static_cast<uint32_t>(p.stages)) {
                        s = 0;
                        ph ^= 1;
                    }
                }
                if (ok) tc_commit_elect(smem_u32(&s_tmem_full[acc]));
                continue;
            }
            if (p.conv_halo) {
                // A descriptors: 8-pixel row groups are halo_w rows apart; tap (r, s) starts r * halo_w + s rows in
                const uint64_t ah_desc0 = make_smem_desc(0, 16, static_cast<uint32_t>(p.halo_w) * 128u, 2u);
                uint32_t accf = 0u;
                for (int ch = 0; ch < p.chunks && ok; ++ch) {
                    if (!mbar_wait_warp(smem_u32(&s_full[s]), ph, spin)) {
                        atomicOr(p.error_flag, 4);
                        ok = false;
                        break;
                    }
                    tc_fence_after();
                    const uint64_t ad0 = ah_desc0 + ((smem_base + s * p.stage_bytes) >> 4);
                    for (int t = 0; t < p.n_taps; t += p.b_taps) {
                        if (!(p.dbg_flags & 8) && !mbar_wait_warp(smem_u32(&s_bfull[sb]), phb, spin)) {
                            atomicOr(p.error_flag, 4);
                            ok = false;
                            break;
                        }
                        tc_fence_after();
                        const uint64_t bd0 = b_desc0 + ((bring_base + sb * p.b_bytes) >> 4);
                        const uint32_t tap16 = static_cast<uint32_t>(p.bn) * 8u;      // one weight tile, 16-byte units
                        for (int tl = 0; tl < p.b_taps; ++tl) {
                            const uint64_t ad =
                                ad0 + static_cast<uint32_t>(p.tap_dh[t + tl] * p.halo_w + p.tap_dw[t + tl]) * 8u;
                            const uint64_t bd = bd0 + tl * tap16;
#pragma unroll
                            for (int kk = 0; kk < TG_BK / 8; ++kk) {
                                tc_mma_tf32_elect(tmem_d, ad + kk * 2u, bd + kk * 2u, idesc, accf);
                                accf = 1u;
                            }
                        }
                        if (!(p.dbg_flags & 8)) tc_commit_elect(smem_u32(&s_bempty[sb]));
                        if (++sb == static_cast<uint32_t>(p.b_stages)) {
                            sb = 0;
                            phb ^= 1;
                        }
                    }
                    tc_commit_elect(smem_u32(&s_empty[s]));       // the haloed tile is free once its taps retire
                    if (++s == static_cast<uint32_t>(p.stages)) {
                        s = 0;
                        ph ^= 1;
                    }
                }
                if (ok) tc_commit_elect(smem_u32(&s_tmem_full[acc]));
                continue;
            }
            // descriptor arithmetic on the low words only (the 14-bit address field cannot carry): half the uniform ALU
            // work per MMA of a 64-bit add
            const uint32_t a_hi0 = static_cast<uint32_t>(a_desc0 >> 32), b_hi0 = static_cast<uint32_t>(b_desc0 >> 32);
            const uint32_t a_lo00 = static_cast<uint32_t>(a_desc0), b_lo00 = static_cast<uint32_t>(b_desc0);
            for (int j = 0; j < wk.nk; ++j) {
                if (!mbar_wait_warp(smem_u32(&s_full[s]), ph, spin)) {
                    atomicOr(p.error_flag, 4);
                    ok = false;
                    break;
                }
                tc_fence_after();
                const uint32_t sa16 = (smem_base + s * p.stage_bytes) >> 4;
                const uint32_t a_lo = a_lo00 + sa16;
                const uint32_t b_lo = b_lo00 + sa16 + (static_cast<uint32_t>(TG_A_BYTES) >> 4);
#pragma unroll
                for (int kk = 0; kk < TG_BK / 8; ++kk) {
                    const uint64_t ad = (static_cast<uint64_t>(a_hi0) << 32) | (a_lo + kk * kstep16);
                    const uint64_t bd = (static_cast<uint64_t>(b_hi0) << 32) | (b_lo + kk * kstep16);
                    if (CG == 2) tc_mma_tf32_pair_elect(tmem_d, ad, bd, idesc, (j > 0 || kk > 0) ? 1u : 0u);
                    else tc_mma_tf32_elect(tmem_d, ad, bd, idesc, (j > 0 || kk > 0) ? 1u : 0u);
                }
                // frees the stage (in both CTAs of a pair) when these MMAs retire
                if (CG == 2) tc_commit_pair_elect(smem_u32(&s_empty[s]));
                else tc_commit_elect(smem_u32(&s_empty[s]));
                if (++s == static_cast<uint32_t>(p.stages)) {
                    s = 0;
                    ph ^= 1;
                }
            }
            if (ok) {
                if (CG == 2) tc_commit_pair_elect(smem_u32(&s_tmem_full[acc]));
                else tc_commit_elect(smem_u32(&s_tmem_full[acc]));
            }
        }
        __syncwarp();
    } else {
        // =============================== epilogue ============================================
        // Two groups of four warps; group g drains the 32-column chunks with (chunk & 1) == g.  Each group
        // owns its named barrier, its share of the staging buffers and its TMA-store issuing thread.
        const int ew = warp - 2;           // 0..7
        const int grp = ew >> 2;           // epilogue group
        const int q = warp & 3;            // TMEM lane quarter this warp may read
        // tile row == TMEM lane; an M = 64 accumulator keeps rows 16 q .. 16 q + 15 in the first 16 lanes of each 32-lane
        // quarter (the other lanes get a row index past every bound and store nothing)
        const int row = p.m64 ? (lane < 16 ? q * 16 + lane : TG_BM) : q * 32 + lane;
        const int et = (ew & 3) * 32 + lane;      // 0..127 within the group
        const int eall = ew * 32 + lane;          // 0..EPI_THREADS-1 over all groups
        const int bar_id = 1 + grp;               // named barriers 1..G: one per group; G+1: all epilogue threads
        constexpr int BAR_ALL = G + 1;
        constexpr int SLOTS = (TG_MAX_BN / 32 + G - 1) / G;      // chunks of one tile a group can own
        // this thread's pixel inside a conv tile (row -> (dw, dh, dn)): the same for every tile
        const int row_dw = p.mode == 0 ? row % p.box_w : 0;
        const int row_dh = p.mode == 0 ? (row / p.box_w) % p.box_h : 0;
        const int row_dn = p.mode == 0 ? (row / p.box_w) / p.box_h : 0;
        const int nbuf = p.nout / G > 0 ? p.nout / G : 1;        // staging buffers per group
        const uint32_t my_stage = stage_base + (p.nout >= G ? grp * nbuf : 0) * TG_A_BYTES;
        // bn_bwd mode (dgrad whose output is the gradient of a BatchNorm + ReLU activation): the TMA-prefetched
        // "residual" tile holds that BatchNorm's INPUT y and is not added; the statistics pass accumulates the
        // BatchNorm backward sums  sum g, sum g * xhat  with  g = dx * (y * scale + shift > 0)  instead of sum / sum^2
        const bool bnb = p.bn_bwd != 0;
        const bool affine = !bnb && ((p.scale != nullptr) || (p.bias != nullptr));
        uint32_t tile_i = 0, cc = 0;
        bool ok = true;
        // ---- residual tiles by TMA: this group's chunk sequence is prefetched two chunks ahead into its two
        //      16 KB buffers (swizzled like the staging tile, so each thread reads back its own row conflict-free)
        const bool res_tma = p.nres > 0;
        const int res_chunks = (p.bn + 31) >> 5;
        const uint32_t rbufs = static_cast<uint32_t>(p.nres / G > 0 ? p.nres / G : 1);   // prefetch buffers per group
        const uint32_t my_res = res_base + grp * rbufs * TG_A_BYTES;
        uint32_t rq = 0;                      // residual chunks consumed by this group
        int pw = w_first, pc = grp;           // next chunk to prefetch (work item, chunk)
        auto res_issue = [&](uint32_t buf) {  // called by one thread
            if (pw >= p.work_total || pc >= res_chunks) return;
            const Work w2 = decode_work_cg<CG>(p, pw, cta_rank);
            const TileOrigin o2 = tile_origin(p, w2.mt);
            const uint32_t bar = smem_u32(&s_res_full[grp * 2 + buf]);
            mbar_arrive_expect_tx(bar, static_cast<uint32_t>(p.m_rows) * 128u);
            tma_load_4d(my_res + buf * TG_A_BYTES, &maps.r, bar, w2.nt * p.bn + pc * 32, o2.w0, o2.h0, o2.n0);
            pc += G;
            if (pc >= res_chunks) {
                pc = grp;
                pw += w_step;
            }
        };
        if (res_tma && et == 0) {
            for (uint32_t b = 0; b < rbufs; ++b) res_issue(b);
        }
        if (p.stats) {
            for (int cidx = eall; cidx < 2 * p.stats_cols; cidx += EPI_THREADS) s_sum[cidx] = 0.f;
            asm volatile("bar.sync %0, %1;" ::"n"(BAR_ALL), "n"(EPI_THREADS) : "memory");
        }
        for (int w = w_first; w < p.work_total && ok; w += w_step) {
            const Work wk = decode_work_cg<CG>(p, w, cta_rank);
            const int n_off = wk.nt * p.bn;
            const uint32_t acc = p.mode == 2 ? 0u : (tile_i & 1);
            const uint32_t aph = p.mode == 2 ? (tile_i & 1) : ((tile_i >> 1) & 1);
            if (wk.nk > 0) ++tile_i;

            // per-tile column constants (everyone is past the previous tile's reads after this barrier); with a single
            // column tile (work_n == 1: every tile has the same columns) they are loaded for the first tile only
            if ((affine || bnb) && (p.work_n > 1 || w == w_first)) {
                asm volatile("bar.sync %0, %1;" ::"n"(BAR_ALL), "n"(EPI_THREADS) : "memory");
                for (int cidx = eall; cidx < p.bn; cidx += EPI_THREADS) {
                    const int col = n_off + cidx;
                    float sc = 1.f, sh = 0.f;
                    if (col < p.n_total) {
                        if (p.scale) {
                            sc = p.scale[col];
                            sh = p.shift[col];
                        } else {
                            sh = p.bias[col];
                        }
                        if (bnb) {
                            s_mean[cidx] = p.bn_mean[col];
                            s_istd[cidx] = p.bn_invstd[col];
                        }
                    } else if (bnb) {
                        s_mean[cidx] = 0.f;
                        s_istd[cidx] = 0.f;
                    }
                    s_scale[cidx] = sc;
                    s_shift[cidx] = sh;
                }
                asm volatile("bar.sync %0, %1;" ::"n"(BAR_ALL), "n"(EPI_THREADS) : "memory");
            }

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
                const int wq = o.w0 + row_dw, hq = o.h0 + row_dh, nq = o.n0 + row_dn;
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

            // mode 2: the accumulators of the group's taps sit side by side, bn columns each
            const int chunks_per_tap = (p.bn + 31) >> 5;
            const int nchunks = p.mode == 2 ? chunks_per_tap * min(p.tg_taps, p.n_taps - wk.tap) : chunks_per_tap;
            float* const out_row0 = out_row;
            int last_c = -1;                       // last chunk this group reads from TMEM
            for (int c = grp; c < nchunks; c += G) last_c = c;
            if (last_c < 0 && wk.nk > 0) {
                if (CG == 2) mbar_arrive_leader(smem_u32(&s_tmem_empty[acc]));
                else mbar_arrive(smem_u32(&s_tmem_empty[acc]));
            }
            for (int c = grp; c < nchunks; c += G) {
                float v[32];
                if (wk.nk > 0) {
                    tmem_ld_32x32(tmem_base + acc * TG_MAX_BN + (static_cast<uint32_t>(q * 32) << 16) + c * 32, v);
                    if (c == last_c) {   // this thread is done with the accumulator: hand it back (to the leader's MMA warp)
                        tc_fence_before();
                        if (CG == 2) mbar_arrive_leader(smem_u32(&s_tmem_empty[acc]));
                        else mbar_arrive(smem_u32(&s_tmem_empty[acc]));
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = 0.f;
                }
                int col0 = n_off + c * 32;
                if (p.mode == 2) {
                    const int tl = c / chunks_per_tap;
                    col0 = n_off + (c - tl * chunks_per_tap) * 32;
                    out_row = out_row0 + tl * p.out_tap_stride;
                }
                if (!row_valid) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = 0.f;
                }
                // ---- affine / residual / activation --------------------------------------------
                if (affine) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 sc4 = *reinterpret_cast<const float4*>(&s_scale[c * 32 + i]);
                        const float4 sh4 = *reinterpret_cast<const float4*>(&s_shift[c * 32 + i]);
                        v[i] = fmaf(v[i], sc4.x, sh4.x);
                        v[i + 1] = fmaf(v[i + 1], sc4.y, sh4.y);
                        v[i + 2] = fmaf(v[i + 2], sc4.z, sh4.z);
                        v[i + 3] = fmaf(v[i + 3], sc4.w, sh4.w);
                    }
                }
                if (res_tma && !bnb) {
                    const uint32_t buf = rq % rbufs;
                    if (!mbar_wait(smem_u32(&s_res_full[grp * 2 + buf]), (rq / rbufs) & 1)) {
                        atomicOr(p.error_flag, 32);
                        ok = false;
                        break;
                    }
                    if (row_valid) {
                        uint4 mb = make_uint4(~0u, ~0u, ~0u, ~0u);
                        int sh = 0;
                        if (p.res_mask) {
                            // the eight float4s of this chunk sit in one 32-group of the bit mask
                            const long long i4 = (row_lin * p.ld_res + col0) >> 2;
                            mb = __ldg(reinterpret_cast<const uint4*>(p.res_mask) + (i4 >> 5));
                            sh = static_cast<int>(i4 & 31);
                        }
                        const uint32_t rrow = my_res + buf * TG_A_BYTES + row * 128;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float4 r4;
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                         : "=f"(r4.x), "=f"(r4.y), "=f"(r4.z), "=f"(r4.w)
                                         : "r"(rrow + ((j ^ (row & 7)) << 4)));
                            const int bit = sh + j;
                            v[4 * j] += ((mb.x >> bit) & 1u) ? r4.x : 0.f;
                            v[4 * j + 1] += ((mb.y >> bit) & 1u) ? r4.y : 0.f;
                            v[4 * j + 2] += ((mb.z >> bit) & 1u) ? r4.z : 0.f;
                            v[4 * j + 3] += ((mb.w >> bit) & 1u) ? r4.w : 0.f;
                        }
                    }
                } else if (res_row && row_valid && !bnb) {
                    if (p.res_mask) {
                        // masked residual (identity branch of a residual join in backward): the eight float4s of
                        // this chunk sit in one 32-group of the bit mask (ld_res % 32 == 0, col0 % 32 == 0)
                        const long long i4 = (row_lin * p.ld_res + col0) >> 2;
                        const uint4 mb = __ldg(reinterpret_cast<const uint4*>(p.res_mask) + (i4 >> 5));
                        const int sh = static_cast<int>(i4 & 31);
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const float4 r4 = *reinterpret_cast<const float4*>(res_row + col0 + i);
                            const int bit = sh + (i >> 2);
                            v[i] += ((mb.x >> bit) & 1u) ? r4.x : 0.f;
                            v[i + 1] += ((mb.y >> bit) & 1u) ? r4.y : 0.f;
                            v[i + 2] += ((mb.z >> bit) & 1u) ? r4.z : 0.f;
                            v[i + 3] += ((mb.w >> bit) & 1u) ? r4.w : 0.f;
                        }
                    } else {
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
                    const uint32_t region = my_stage + (nbuf == 2 ? (cc & 1u) : (cc % nbuf)) * TG_A_BYTES;
                    // the store issued from this buffer `nbuf` chunks ago must have finished reading it
                    if (et == 0) {
                        if (nbuf >= 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                        else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    }
                    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                    if (res_tma && !bnb) {
                        // every thread of the group has read residual buffer rq % rbufs: refill it `rbufs` chunks ahead
                        if (et == 0) res_issue(rq % rbufs);
                        ++rq;
                    }
                    const uint32_t rbase = region + row * 128;
                    if (!(p.dbg_flags & 2)) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const uint32_t addr = rbase + ((j ^ (row & 7)) << 4);
                            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v[4 * j]),
                                         "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3])
                                         : "memory");
                        }
                    }
                    fence_proxy_async_smem();
                    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                    if (et == 0 && !(p.dbg_flags & 3)) {
                        tma_store_4d(&maps.d, region, col0, o.w0, o.h0, o.n0);
                        tma_store_commit();
                    }
                    ++cc;
                    // ---- per-channel batch statistics: column pass over the staged (raw) tile ------
                    // thread = (4-column group cg, 8-row group rg); conflict-free 128-bit reads
                    if (p.stats) {
                        const int cg = et & 7, rg = et >> 3;
                        float4 sm = make_float4(0.f, 0.f, 0.f, 0.f), sq = sm;
                        if (bnb) {
                            // BatchNorm backward sums over the staged dx tile and the prefetched y tile (same swizzled
                            // layout): g = dx where the activation was positive, xhat = (y - mean) / std
                            const float4 sc4 = *reinterpret_cast<const float4*>(&s_scale[c * 32 + cg * 4]);
                            const float4 sh4 = *reinterpret_cast<const float4*>(&s_shift[c * 32 + cg * 4]);
                            const float4 mu4 = *reinterpret_cast<const float4*>(&s_mean[c * 32 + cg * 4]);
                            const float4 is4 = *reinterpret_cast<const float4*>(&s_istd[c * 32 + cg * 4]);
                            // the y tile of this chunk was requested one chunk ago (ONE buffer per group is enough:
                            // it is consumed only here, at the end of the chunk's work)
                            const uint32_t ybuf = my_res + (rq % rbufs) * TG_A_BYTES;
                            if (!mbar_wait(smem_u32(&s_res_full[grp * 2 + (rq % rbufs)]), (rq / rbufs) & 1)) {
                                atomicOr(p.error_flag, 32);
                                ok = false;
                            }
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                if (rg * 8 + i < p.m_rows) {          // rows past the box are not written by TMA
                                    const uint32_t off = (rg * 8 + i) * 128 + ((cg ^ i) << 4);
                                    float4 t, y;
                                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                                 : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w)
                                                 : "r"(region + off));
                                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                                 : "=f"(y.x), "=f"(y.y), "=f"(y.z), "=f"(y.w)
                                                 : "r"(ybuf + off));
                                    const float gx = fmaf(y.x, sc4.x, sh4.x) > 0.f ? t.x : 0.f;
                                    const float gy = fmaf(y.y, sc4.y, sh4.y) > 0.f ? t.y : 0.f;
                                    const float gz = fmaf(y.z, sc4.z, sh4.z) > 0.f ? t.z : 0.f;
                                    const float gw = fmaf(y.w, sc4.w, sh4.w) > 0.f ? t.w : 0.f;
                                    sm.x += gx; sm.y += gy; sm.z += gz; sm.w += gw;
                                    sq.x += gx * (y.x - mu4.x) * is4.x;
                                    sq.y += gy * (y.y - mu4.y) * is4.y;
                                    sq.z += gz * (y.z - mu4.z) * is4.z;
                                    sq.w += gw * (y.w - mu4.w) * is4.w;
                                }
                            }
                            // every thread of the group is done with the y buffer: request the next chunk's tile
                            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                            if (et == 0) res_issue(rq % rbufs);
                            ++rq;
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const uint32_t addr = region + (rg * 8 + i) * 128 + ((cg ^ i) << 4);
                                float4 t;
                                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                             : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w)
                                             : "r"(addr));
                                sm.x += t.x; sm.y += t.y; sm.z += t.z; sm.w += t.w;
                                sq.x = fmaf(t.x, t.x, sq.x); sq.y = fmaf(t.y, t.y, sq.y);
                                sq.z = fmaf(t.z, t.z, sq.z); sq.w = fmaf(t.w, t.w, sq.w);
                            }
                        }
                        // lanes {l, l^8, l^16, l^24} share cg: fold the four row groups of the warp
#pragma unroll
                        for (int off = 8; off <= 16; off <<= 1) {
                            sm.x += __shfl_xor_sync(0xffffffffu, sm.x, off);
                            sm.y += __shfl_xor_sync(0xffffffffu, sm.y, off);
                            sm.z += __shfl_xor_sync(0xffffffffu, sm.z, off);
                            sm.w += __shfl_xor_sync(0xffffffffu, sm.w, off);
                            sq.x += __shfl_xor_sync(0xffffffffu, sq.x, off);
                            sq.y += __shfl_xor_sync(0xffffffffu, sq.y, off);
                            sq.z += __shfl_xor_sync(0xffffffffu, sq.z, off);
                            sq.w += __shfl_xor_sync(0xffffffffu, sq.w, off);
                        }
                        // per-warp partials go to the tile scratch; they are folded once per tile (below) by threads
                        // that own their columns exclusively -- no shared-memory atomics on the per-chunk path
                        if (lane < 8) {
                            float* dst = s_part + (((grp * SLOTS + c / G) * 4 + (ew & 3)) << 6) + cg * 4;
                            *reinterpret_cast<float4*>(dst) = sm;
                            *reinterpret_cast<float4*>(dst + 32) = sq;
                        }
                    }
                } else if (row_valid && !(p.dbg_flags & 64)) {
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
                    } else if (col0 + 32 <= p.n_total && (p.ldo & 3) == 0) {
                        // split-K partial sums: 128-bit vector reductions (4x fewer RED instructions)
#pragma unroll
                        for (int i = 0; i < 32; i += 4)
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out_row + col0 + i),
                                         "f"(v[i]), "f"(v[i + 1]), "f"(v[i + 2]), "f"(v[i + 3])
                                         : "memory");
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (col0 + i < p.n_total) atomicAdd(out_row + col0 + i, v[i]);
                    }
                }
            }
            if (p.stats && p.store_mode == TG_STORE_TMA) {
                // fold this tile's per-warp partials into the CTA-wide sums: (group, chunk, column) -> unique owner
                asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
#pragma unroll
                for (int h = 0; h < (SLOTS * 64 + 127) / 128; ++h) {
                    const int idx = et + h * 128;
                    const int ci = idx >> 6, k = idx & 63;
                    const int c = G * ci + grp;
                    const int col = n_off + c * 32 + (k & 31);
                    if (ci < SLOTS && c < nchunks && col < p.n_total) {
                        const float* src = s_part + ((grp * SLOTS + ci) << 8) + k;
                        const float t = (src[0] + src[64]) + (src[128] + src[192]);
                        float* dst = (k < 32 ? s_sum : s_sq) + col;
                        *dst += t;
                    }
                }
            }
        }
        if (p.stats) {
            asm volatile("bar.sync %0, %1;" ::"n"(BAR_ALL), "n"(EPI_THREADS) : "memory");
            for (int col = eall; col < p.n_total; col += EPI_THREADS) {
                atomicAdd(p.stats + col, static_cast<double>(s_sum[col]));
                atomicAdd(p.stats + p.n_total + col, static_cast<double>(s_sq[col]));
            }
        }
        if (p.store_mode == TG_STORE_TMA && et == 0) tma_store_wait_all();
    }

    // ---- teardown ---------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();      // nothing of the pair is in flight any more: barriers, stages, accumulators
    if (warp == 1) {
        tc_fence_after();
        if (CG == 2) tmem_dealloc_pair(tmem_base, 2 * TG_MAX_BN);
        else tmem_dealloc(tmem_base, 2 * TG_MAX_BN);
    }
}

}  // namespace

int g_pdl = -1;     // programmatic dependent launch (pe_debug_pdl); -1 = not read from the environment yet
int g_sm_reserve = 0;   // SMs left free by the persistent tap-GEMM grids (pe_set_sm_reserve)

bool pdl_enabled(int kind) {
    if (g_pdl < 0) {
        const char* e = getenv("PE_B200_PDL");
        g_pdl = e ? (atoi(e) & 3) : 1;
    }
    return (g_pdl & kind) != 0;
}

// Multiply-shift division by a launch constant (TapParams::fd_*, fast_div() in the kernel): m = ceil(2^(31 + l) / d) with
// l = ceil(log2 d), q = umulhi(n, m) >> (l - 1), exact for 0 <= n < 2^31; {0, 0} encodes d <= 1.
void fast_div_of(unsigned (&fd)[2], int d) {
    fd[0] = fd[1] = 0;
    if (d <= 1) return;
    unsigned l = 0;
    while ((1u << l) < static_cast<unsigned>(d)) ++l;
    fd[0] = static_cast<unsigned>(((1ull << (31 + l)) + d - 1) / d);
    fd[1] = l - 1;
}

int launch_tapgemm(const TapMaps& maps, TapParams& p, dim3 work, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        PE_CHECK_CUDA(cudaFuncSetAttribute(tapgemm_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           TG_SMEM_BYTES));
        PE_CHECK_CUDA(cudaFuncSetAttribute(tapgemm_kernel<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           TG_SMEM_BYTES));
        PE_CHECK_CUDA(cudaFuncSetAttribute(tapgemm_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           TG_SMEM_BYTES));
        PE_CHECK_CUDA(cudaFuncSetAttribute(tapgemm_kernel<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           TG_SMEM_BYTES));
        configured = true;
    }
    if (p.epi_groups != 4) p.epi_groups = 2;
    if (p.cta_group != 2) p.cta_group = 1;
    PE_REQUIRE(p.cta_group == 1 || (p.mode == 0 && !p.conv_halo && p.ksplit == 1 && p.bn % 32 == 0),
               "tap-GEMM: CTA pairs are implemented for the plain conv / GEMM mode only");
    {
        // the carve-up of dynamic shared memory must fit: ring + weight ring + store staging + residual tiles +
        // statistics scratch (a violation here would be an out-of-bounds shared-memory access on the device)
        const long long need = (long long)p.stages * p.stage_bytes + p.b_ring_bytes +
                               (long long)(p.store_mode == TG_STORE_TMA ? p.nout + p.nres : 0) * TG_A_BYTES +
                               (p.stats ? 2ll * p.stats_cols * (long long)sizeof(float) + 8192 : 0) +
                               (p.bn_bwd ? 2ll * TG_MAX_BN * (long long)sizeof(float) : 0);
        PE_REQUIRE(p.stages >= 1 && need <= TG_SMEM_BYTES, "tap-GEMM shared-memory plan of %lld bytes does not fit in %d",
                   need, TG_SMEM_BYTES);
    }
    p.work_n = work.x;
    p.work_m = work.y;
    p.work_total = static_cast<int>(work.x * work.y * work.z);
    {
        fast_div_of(p.fd_work_n, p.work_n);
        fast_div_of(p.fd_work_m, p.work_m);
        fast_div_of(p.fd_tiles_w, p.tiles_w);
        fast_div_of(p.fd_tiles_h, p.tiles_h);
        const int ks = p.ksplit > 0 ? p.ksplit : 1;
        p.k_per = ((p.mode == 0 ? p.n_taps * p.chunks : p.pt_total) + ks - 1) / ks;
    }
    // one CTA per SM, or one CTA pair per TPC (the work list then counts pairs of tiles)
    // (minus the SMs set aside for a concurrent collective: a persistent grid that needs EVERY SM would otherwise wait
    // for the NCCL kernel's CTAs and run its last tiles as a second wave)
    const int sms = num_sms() - g_sm_reserve > 8 ? num_sms() - g_sm_reserve : num_sms();
    const int slots = p.cta_group == 2 ? sms / 2 : sms;
    int grid = p.work_total < slots ? p.work_total : slots;
    if (grid < 1) return 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid * p.cta_group, 1, 1);
    cfg.blockDim = dim3(64 + 128 * p.epi_groups, 1, 1);
    cfg.dynamicSmemBytes = TG_SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (p.cta_group == 2) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 2;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = 1;
        ++na;
    }
    if (pdl_enabled(1)) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    if (p.cta_group == 2) {
        if (p.epi_groups == 4) PE_CHECK_CUDA(cudaLaunchKernelEx(&cfg, tapgemm_kernel<4, 2>, maps, p));
        else PE_CHECK_CUDA(cudaLaunchKernelEx(&cfg, tapgemm_kernel<2, 2>, maps, p));
    } else {
        if (p.epi_groups == 4) PE_CHECK_CUDA(cudaLaunchKernelEx(&cfg, tapgemm_kernel<4, 1>, maps, p));
        else PE_CHECK_CUDA(cudaLaunchKernelEx(&cfg, tapgemm_kernel<2, 1>, maps, p));
    }
    PE_LAUNCH_CHECK();
    return 0;
}

}  // namespace pe

// Tap-GEMM: the one tcgen05/TMEM/TMA kernel behind every dense contraction on the path.
//
//   mode 0 ("conv"):  D[pixels, N] = sum_taps sum_k  A_tap[pixels, k] * B_tap[N, k]
//                     A, B K-major in shared memory (rows of 128 B = 32 tf32 along K).
//                     A rows are a 4-D TMA box of output pixels (w, h, n) shifted per filter tap;
//                     zero padding comes from TMA out-of-bounds fill.  Plain GEMMs (1x1 convs,
//                     linear layers, LSTM projections) are the 1-tap, 1-D-box special case.
//   mode 1 ("wgrad"): D[M, N] = sum_pixels A[pixel, M]^T * B[pixel+tap, N]      (one tap per work item)
//                     A (dY) and B (X) MN-major in shared memory: the reduction runs over pixels.
//   mode 2 ("haloed wgrad", stride 1): one dY pixel tile and ONE X tile with an (R-1, S-1) halo per stage; each
//                     tap of a group reads its shifted window of that X tile through its own shared-memory
//                     descriptor and accumulates into its own TMEM columns (3-5 taps per work item).
//
// Persistent CTAs (one per SM) walk a list of 128 x bn output tiles; warp 0 = TMA producer,
// warp 1 = TMEM owner + MMA issuer (two accumulators, so the next tile overlaps the epilogue),
// warps 2.. = 2 or 4 epilogue groups (TMEM -> registers -> swizzled smem -> TMA store, with optional
// bias / scale-shift / TMA-prefetched residual (+bit mask) / ReLU / TF32 rounding / per-channel batch statistics).
#pragma once
#include "pe_common.cuh"

namespace pe {

constexpr int TG_BM = 128;                     // tile rows  (TMEM lanes)
constexpr int TG_BK = 32;                      // tf32 elements per k-step (128 B swizzle span)
constexpr int TG_MAX_BN = 256;                 // tile columns (TMEM columns per accumulator)
constexpr int TG_STAGES = 8;                    // maximum ring depth (runtime: TapParams::stages)
constexpr int TG_A_BYTES = TG_BM * 128;        // 16 KB
constexpr int TG_B_BYTES = TG_MAX_BN * 128;    // 32 KB (16 KB when bn <= 128: TapParams::stage_bytes)
constexpr int TG_SMEM_BYTES = 223 * 1024;       // ring (stages x 32|48 KB) + store staging (nout x 16 KB)
constexpr int TG_THREADS = 320;                // TMA warp + MMA warp + 2 epilogue groups of 4 warps (default)
constexpr int TG_MAX_EPI_GROUPS = 4;           // store-bound launches use 4 groups (576 threads), see TapParams::epi_groups
constexpr int TG_MAX_TAPS = 16;

enum TgStore { TG_STORE_TMA = 0, TG_STORE_DIRECT = 1, TG_STORE_ATOMIC = 2 };

struct alignas(64) TapMaps {
    CUtensorMap a[4];
    CUtensorMap b[4];
    CUtensorMap d;
    CUtensorMap r;            // residual tensor, same geometry as d (only when TapParams::nres > 0)
};

struct TapParams {
    int mode;                 // 0 conv / gemm, 1 wgrad (one tap per work item), 2 wgrad with a haloed X tile
                              //   (several taps per work item, each tap in its own TMEM accumulator)
    int bn;                   // MMA N (multiple of 16; multiple of 32 in wgrad mode)
    int m_rows;               // rows the A box really fills (<= 128)
    int box_w, box_h, box_n;  // pixel box (conv: product = m_rows; wgrad: product = 32)
    int tiles_w, tiles_h, tiles_n;
    int out_w, out_h, out_n;  // valid extents of the pixel grid the boxes tile
    int n_taps, chunks;       // conv: k-steps = n_taps * chunks
    int ksplit;               // K splits (conv: gridDim.z; wgrad: per tap)
    int pt_total;             // wgrad: number of 32-pixel tiles
    int work_n, work_m, work_total;  // persistent work list: (n tile, m tile, z), n fastest
    // Division by the run-time constants of the work list without the ~25-instruction integer-divide sequence (every
    // warp of the CTA decodes every work item: for short-K tiles the divisions were a fifth of all stall samples):
    // q = umulhi(n, mul) >> shr for 0 <= n < 2^31 (mul == 0: divisor 1).  Filled by launch_tapgemm.
    unsigned fd_work_n[2], fd_work_m[2], fd_tiles_w[2], fd_tiles_h[2];
    int k_per;                // k-steps (mode 0) / pixel tiles (modes 1, 2) per K split
    int m_total, n_total;     // logical output extents (wgrad rows; columns in both modes)
    signed char tap_dw[TG_MAX_TAPS], tap_dh[TG_MAX_TAPS], tap_map[TG_MAX_TAPS], tap_b[TG_MAX_TAPS];
    int store_mode;
    float* out;               // direct / atomic destination
    long long out_tap_stride; // wgrad: elements between taps in `out`
    int ldo;                  // row stride of `out` in elements
    const float* bias;        // [n_total] or null
    const float* scale;       // [n_total] or null  (v = acc*scale + shift)
    const float* shift;
    const float* residual;    // same indexing as a dense NHWC output, row stride ld_res
    const unsigned* res_mask; // optional bit mask applied to `residual` (layout: pe_elementwise.cu ld_maskbits)
    int ld_res;
    int relu;
    int round_out;            // round result to tf32 (rna) so the consumer's truncation is exact
    double* stats;            // [2][n_total]: sum, sum of squares of the raw accumulator
    int* error_flag;
    // debug overrides for the smem descriptors (bytes, <0 = default)
    int dbg_a_lbo, dbg_a_sbo, dbg_b_lbo, dbg_b_sbo;
    int stages, nout;         // smem split: ring depth and number of 16 KB store-staging buffers
    int epi_groups;           // 2 or 4 epilogue groups of 4 warps; group g drains the 32-column chunks c with c % G == g
    int nres;                 // 16 KB residual tiles prefetched by TMA for the epilogue (0 or 4: two per group)
    int stage_bytes;          // 16 KB (A) + the B rows of one stage in whole 4 KB atoms
    int stats_cols;           // columns of the CTA-wide statistics scratch (n_total rounded up to 32; 0 = none)
    int dbg_flags;            // 1: skip TMA store issue, 2: skip staging write + store, 4: skip A loads, 8: skip B loads
    // mode 2 (haloed wgrad): pixel tile box_w x box_h x box_n (box_w % 8 == 0), halo tile halo_w x halo_h x box_n
    int tg_taps, n_groups;    // taps per work item (tg_taps * bn <= 512 TMEM columns), number of tap groups
    int m64;                  // modes 1 / 2 with m_total <= 64: tcgen05.mma M = 64 (rows 16 q + i in lane 32 q + i)
    int halo_w, halo_h;       // box_w + S - 1, box_h + R - 1
    int halo_dw, halo_dh;     // halo origin relative to the pixel tile origin (-pad)
    int a_atom_bytes, b_atom_bytes;   // bytes between 32-channel atoms of the dY / X tiles in a stage
    int a_region_bytes;       // offset of the B (X) region inside a stage
    int base_offset_mode;     // debug: 1 = put (start >> 7) & 7 into the descriptors' base-offset field
    // haloed stride-1 conv (mode 0 with conv_halo = 1): the A ring holds one haloed pixel tile per channel chunk
    // (halo_w x halo_h x 1 rows of 128 B, stage_bytes each); every tap reads its shifted window of it through its own
    // descriptor (8-pixel row groups SBO = halo_w * 128 B apart); weights stream through a separate B ring.
    int conv_halo;
    int b_stages, b_bytes;    // B ring: stages and bytes per stage (b_taps * bn * 128)
    int b_taps;               // filter taps per B stage (one TMA box {32, bn, b_taps}); divides n_taps
    int b_ring_bytes;         // b_stages * b_bytes (0 when conv_halo == 0): offset of the store staging after the A ring
    // CTA pairs (mode 0 without conv_halo): cta_group = 2 runs the work list on clusters of two CTAs; a pair owns two
    // consecutive 128-pixel tiles (work_m counts PAIRS), each CTA loads its own A box and HALF of the B tile
    // (bn / 2 weight rows), the leader issues tcgen05.mma.cta_group::2 with M = 256 for both.
    int cta_group;
    // bn_bwd (dgrad only): `residual` / maps.r hold the input y of the BatchNorm + ReLU whose OUTPUT gradient this
    // launch produces; it is not added.  With scale / shift (the forward's folded coefficients), bn_mean / bn_invstd
    // the statistics pass accumulates that BatchNorm's backward sums into stats[2][n_total] = (sum g, sum g * xhat),
    // g = dx * (y * scale + shift > 0): bn_bwd_reduce without a second read of dx and without its own launch.
    int bn_bwd;
    const float* bn_mean;
    const float* bn_invstd;
};

void fast_div_of(unsigned (&fd)[2], int d);   // multiply-shift constants of TapParams::fd_* for the divisor d
int launch_tapgemm(const TapMaps& maps, TapParams& p, dim3 work, cudaStream_t stream);

}  // namespace pe

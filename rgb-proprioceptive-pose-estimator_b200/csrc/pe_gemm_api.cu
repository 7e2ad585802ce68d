// Host side of the tap-GEMM: TMA tensor-map construction, pixel-box selection, and the C-ABI
// entry points for convolution forward / dgrad / wgrad and linear layers (see include/pe_b200.h).
#include <stdarg.h>
#include <string.h>

#include <map>
#include <mutex>
#include <tuple>

#include "../../include/pe_b200.h"
#include "pe_tapgemm.cuh"

namespace pe {

extern int g_pdl;      // pe_tapgemm.cu: programmatic dependent launch switch
extern int g_sm_reserve;   // pe_tapgemm.cu: SMs the persistent grids leave to a concurrent collective

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
            n = 148;
    }
    return n;
}

static int* g_error_flag = nullptr;  // device int, sticky; read by pe_device_error()
static int ensure_error_flag() {
    if (!g_error_flag) {
        PE_CHECK_CUDA(cudaMalloc(&g_error_flag, sizeof(int)));
        PE_CHECK_CUDA(cudaMemset(g_error_flag, 0, sizeof(int)));
    }
    return 0;
}

int* device_error_flag() { return ensure_error_flag() ? nullptr : g_error_flag; }

// Turns a pipeline timeout into something the caller cannot miss without a host synchronisation: when the sticky
// flag is set, the given result buffer (the step's loss, the rollout step's pose) is overwritten with NaNs.
__global__ void poison_on_error_kernel(const int* flag, float* buf, long long n) {
    if (*flag == 0) return;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) buf[i] = __int_as_float(0x7fc00000);
}

// ---------------------------------------------------------------------------------------------
// tensor maps
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static int ensure_encode() {
    if (g_encode) return 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    PE_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    PE_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess,
               "cuTensorMapEncodeTiled not available from the driver");
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    return 0;
}

// fp32 tensor, 4 dims (d0 contiguous), strides in ELEMENTS for dims 1..3, SWIZZLE_128B.
static int make_map(CUtensorMap* m, const void* base, const long long dims[4], const long long strides[3],
                    const int box[4], bool mn_major = false) {
    if (ensure_encode()) return 1;
    cuuint64_t gd[4], gs[3];
    cuuint32_t bx[4], es[4] = {1, 1, 1, 1};
    for (int i = 0; i < 4; ++i) {
        gd[i] = static_cast<cuuint64_t>(dims[i] > 0 ? dims[i] : 1);
        bx[i] = static_cast<cuuint32_t>(box[i]);
        PE_REQUIRE(box[i] >= 1 && box[i] <= 256, "TMA box dim %d = %d out of range", i, box[i]);
    }
    for (int i = 0; i < 3; ++i) {
        gs[i] = static_cast<cuuint64_t>(strides[i]) * 4ull;
        PE_REQUIRE(gs[i] % 16 == 0, "TMA stride %d = %llu B is not a multiple of 16", i,
                   (unsigned long long)gs[i]);
    }
    PE_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base %p not 16-byte aligned", base);
    PE_REQUIRE(box[0] * 4 <= 128 && (box[0] * 4) % 16 == 0, "TMA inner box %d elems invalid", box[0]);
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), gd, gs, bx, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE,
                          mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PE_REQUIRE(r == CUDA_SUCCESS,
               "cuTensorMapEncodeTiled failed (%d): dims %lld %lld %lld %lld strides %lld %lld %lld box %d %d %d %d",
               (int)r, dims[0], dims[1], dims[2], dims[3], strides[0], strides[1], strides[2], box[0], box[1],
               box[2], box[3]);
    return 0;
}

// NHWC activation view (optionally one stride-2 parity plane), channels innermost.
static int make_nhwc_map(CUtensorMap* m, const float* base, int N, int H, int W, int C, int ph, int pw,
                         int step, const int box[4], bool mn_major = false) {
    const long long dims[4] = {C, (W - pw + step - 1) / step, (H - ph + step - 1) / step, N};
    const long long strides[3] = {(long long)step * C, (long long)step * W * C, (long long)H * W * C};
    return make_map(m, base + ((long long)ph * W + pw) * C, dims, strides, box, mn_major);
}

// Activation view with explicit pixel strides (in elements): the space-to-depth stem reads its operand through a map
// whose pixel stride (12 floats) is SMALLER than the channel extent (48 floats) -- every "pixel row" of the map is the
// sliding window of four consecutive 12-channel pixels; channel coordinates >= c_extent are zero-filled by TMA.
struct XGeom {
    int c_extent;
    long long sw, sh, sn;
};
static int make_xgeom_map(CUtensorMap* m, const float* base, int N, int H, int W, const XGeom& g, const int box[4],
                          bool mn_major = false) {
    const long long dims[4] = {g.c_extent, W, H, N};
    const long long strides[3] = {g.sw, g.sh, g.sn};
    return make_map(m, base, dims, strides, box, mn_major);
}

// Pick a (bw, bh, bn) pixel box with bw*bh*bn <= target that tiles (W, H, N) with the fewest boxes.
// exact=true: bw*bh*bn == target exactly (boxes may overhang the tensor; TMA zero-fills), needed when
// the box rows are the reduction dimension (wgrad) and every shared-memory row must be defined.
static void choose_box(int W, int H, int N, int target, int* obw, int* obh, int* obn, bool exact = false) {
    static std::map<std::tuple<int, int, int, int>, std::tuple<int, int, int>> cache;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    auto key = std::make_tuple(W, H, N, exact ? -target : target);
    auto it = cache.find(key);
    if (it == cache.end()) {
        long long best = -1;
        int bbw = 1, bbh = 1, bbn = 1;
        for (int bw = 1; bw <= target && (exact || bw <= W); ++bw) {
            for (int bh = 1; bw * bh <= target && (exact || bh <= H); ++bh) {
                int bn = target / (bw * bh);
                if (exact) {
                    if (bw * bh * bn != target) continue;
                } else {
                    if (bn > N) bn = N;
                    if (bn < 1) bn = 1;
                }
                if (bn > 256 || bw > 256 || bh > 256) continue;
                long long tiles = (long long)((W + bw - 1) / bw) * ((H + bh - 1) / bh) * ((N + bn - 1) / bn);
                // fewest tiles; ties -> widest rows (better DRAM locality)
                if (best < 0 || tiles < best || (tiles == best && bw > bbw)) {
                    best = tiles;
                    bbw = bw;
                    bbh = bh;
                    bbn = bn;
                }
            }
        }
        it = cache.emplace(key, std::make_tuple(bbw, bbh, bbn)).first;
    }
    *obw = std::get<0>(it->second);
    *obh = std::get<1>(it->second);
    *obn = std::get<2>(it->second);
}

static int g_dbg_desc[4] = {-1, -1, -1, -1};
static int g_dbg_flags = 0;
static int g_dbg_stages = 0, g_dbg_nout = 0;

static void init_params(TapParams& p) {
    memset(&p, 0, sizeof(p));
    p.ksplit = 1;
    p.dbg_a_lbo = g_dbg_desc[0];
    p.dbg_a_sbo = g_dbg_desc[1];
    p.dbg_b_lbo = g_dbg_desc[2];
    p.dbg_b_sbo = g_dbg_desc[3];
    p.dbg_flags = g_dbg_flags;

    p.error_flag = g_error_flag;
}

static int g_dbg_max_bn = 256;
static int g_dbg_res_tma = 1;
static int g_dbg_min_bn = 32;         // narrowest tile the small-launch heuristic of pick_bn may choose
static int g_dbg_epi_groups = 0;      // 0 = automatic, 2 / 4 = forced, 6 = automatic without the residual-launch rule
// Tile width.  Wide tiles halve the operand traffic per MMA, but a launch with only a handful of tiles (rollout at
// batch 1: 49-3136 pixels per image) would leave most SMs idle behind one long serial K loop: there the width is
// halved (down to 64) until the tile count reaches half the SM count.
static int pick_bn(int n_total, long long m_tiles = 1 << 20) {
    int bn;
    if (n_total >= 256 && g_dbg_max_bn >= 256) bn = 256;
    else bn = n_total >= 128 ? 128 : ((n_total + 15) / 16) * 16;
    while (bn > g_dbg_min_bn && m_tiles * ((n_total + bn - 1) / bn) * 2 <= num_sms()) bn >>= 1;
    return bn;
}

static int g_dbg_cta_group = 0;       // 0: automatic, 1: never pair CTAs, 2: pair whenever the launch allows it,
                                      // 3: automatic, 256-column tiles only (the rule before the ring was deepened)
// CTA pairs (tcgen05 cta_group::2, M = 256): each CTA keeps its own 128-pixel A box and HALF of the weight tile, and ONE
// MMA warp issues for both SMs.  The tensor pipe itself is not the limit of a lone CTA (tests/probe_mma_rate.cu: 128 /
// 64 / 48 clk per 128 x {256, 128, 64} x 8 TF32 instruction straight from shared memory); the operand ring is: a stage's
// MMAs retire long before its slot has made the round trip commit -> producer -> TMA -> full barrier, so the rate is
// (MMAs held by the ring) / (round trip).  A pair holds the same MMAs in 2/3 .. 5/6 of the bytes -> a deeper ring out of
// the same shared memory, half the L2 -> SM weight traffic and half the issue / barrier rounds per SM.  Used whenever the
// launch has at least one tile per SM (measured on B200, 256 frames: 14x14 256->1024 forward 97 -> 69 us, 3x3 256->256
// @14 91 -> 82 us; with B regions sized by the rows really loaded also 56x56 64->64 3x3 196 -> 167 us, 28x28 128->128
// 3x3 114 -> 100 us: profiles/layers_r02b.txt).
static bool want_pair(const TapParams& p, long long m_tiles, int n_tiles) {
    if (p.mode != 0 || p.conv_halo || g_dbg_flags || g_dbg_cta_group == 1 || (p.bn % 32) || m_tiles < 2) return false;
    if (g_dbg_cta_group == 2) return true;
    return (p.bn == 256 || g_dbg_cta_group != 3) && m_tiles * n_tiles >= (long long)num_sms();
}

// Split the 216 KB of dynamic shared memory between the operand ring and the store staging buffers.
// Long reductions want ring depth, short ones (1x1 convs) are bound by the epilogue's store pipeline.
static void pick_pipeline(TapParams& p, int ksteps) {
    // B region: the weight (mode 1: X) rows this CTA loads per stage, in whole 32-row swizzle atoms.  Narrow tiles get
    // a DEEPER ring out of the same shared memory: their per-stage MMA time (4 x 48 clk at 64 columns) is far below the
    // slot round trip (commit -> producer -> TMA -> full barrier, ~1000 clk), so throughput ~ stages / round trip
    // (measured, 56x56 64->64 3x3: 308 / 239 / 209 us at 2 / 3 / 4 stages; tests/probe_narrow3x3.py).
    const int bn_cta = p.bn / (p.cta_group == 2 ? 2 : 1);
    p.stage_bytes = TG_A_BYTES + (g_dbg_flags & 2048 ? (bn_cta > 128 ? 32768 : 16384) : (bn_cta + 31) / 32 * 4096);
    p.stats_cols = p.stats ? (p.n_total + 31) / 32 * 32 : 0;
    // residual tiles are prefetched by TMA (two 16 KB buffers per epilogue group) when the output goes out by TMA
    // (bn_bwd mode reads its y tile only at the end of a chunk's work: one buffer per group, requested a chunk ahead)
    p.nres = (p.residual && p.store_mode == TG_STORE_TMA && (g_dbg_res_tma || p.bn_bwd)) ? (p.bn_bwd ? 2 : 4) : 0;
    const int budget = TG_SMEM_BYTES - (p.stats ? 2 * p.stats_cols * (int)sizeof(float) + 8192 : 0) -
                       (p.bn_bwd ? 2 * TG_MAX_BN * (int)sizeof(float) : 0) - p.nres * TG_A_BYTES;
    // Store-bound launches (training-mode conv with batch statistics, wide tile, short K loop) are limited by the
    // epilogue's instruction latency, not by HBM: they run four epilogue groups (16 warps) instead of two.
    p.epi_groups = (!p.residual && (g_dbg_epi_groups == 4 ||
                                    (g_dbg_epi_groups == 0 && p.stats && p.store_mode == TG_STORE_TMA &&
                                     p.bn >= 128 && ksteps < 24))) ? 4 : 2;
    // The dgrads that add a TMA-prefetched residual tile (conv1 of every block: short K loop, 128 KB read + 128 KB
    // written per tile) are epilogue-bound as well: on CTA pairs (whose 32 KB stages leave the room) they run four groups
    // with ONE residual and ONE staging buffer each instead of two groups with two -- when the K loop is at most four
    // steps, because the extra staging buffers leave the ring two stages (measured, 256 frames: 56x56 256->64 377 -> 337
    // us, 56x56 256->128 399 -> 341, 28x28 512->128 203 -> 175; with 8 or 16 k-steps 28x28 512->256 205 -> 220, 7x7
    // 2048->512 73 -> 94: those keep two groups; pe_debug_epilogue_groups(6) = the old rule everywhere)
    if (p.residual && p.nres == 4 && !p.bn_bwd && p.cta_group == 2 && ksteps <= 4 &&
        (g_dbg_epi_groups == 0 || g_dbg_epi_groups == 4))
        p.epi_groups = 4;
    // (the staging buffers are split evenly between the epilogue groups)
    int nout = (p.store_mode == TG_STORE_TMA) ? (p.epi_groups == 4 ? 4 : (ksteps >= 24 ? 2 : 4)) : 0;
    if (p.nres && p.bn > 128 && p.epi_groups != 4) nout = 2;
    if (g_dbg_nout > 0 && p.store_mode == TG_STORE_TMA && !(p.epi_groups == 4 && g_dbg_nout < 4)) nout = g_dbg_nout;
    int stages = (budget - nout * TG_A_BYTES) / p.stage_bytes;
    if (stages > TG_STAGES) stages = TG_STAGES;
    if (g_dbg_stages > 0 && g_dbg_stages < stages) stages = g_dbg_stages;
    p.stages = stages;
    p.nout = nout > 0 ? nout : 2;
}

static int g_dbg_halo_btaps = 0;      // taps per weight stage of the haloed conv (0 = default 3)
static int g_dbg_conv_halo = 0;       // 0: off, 1: on for <= 128 input channels (3x3 stride-1 forward / dgrad)

// Geometry of the haloed stride-1 conv (TapParams::conv_halo): 8-pixel-wide tiles of bh rows inside one image, the
// A ring holds (bh + R - 1) x (8 + S - 1) haloed tiles, weights stream through their own ring.  Returns false when
// the shape is not handled (the caller then uses the per-tap box path).
static bool setup_conv_halo(TapParams& p, int B, int H, int W, int Cin, int R, int S, int pad) {
    if (!g_dbg_conv_halo || R != 3 || S != 3 || pad != 1 || Cin > 128 || Cin % 32 != 0 || H < 14) return false;
    p.conv_halo = 1;
    p.box_w = 8;
    p.box_h = (H % 14 == 0) ? 14 : 16;
    p.box_n = 1;
    p.m_rows = p.box_w * p.box_h;
    p.tiles_w = (W + p.box_w - 1) / p.box_w;
    p.tiles_h = (H + p.box_h - 1) / p.box_h;
    p.tiles_n = B;
    p.halo_w = p.box_w + S - 1;
    p.halo_h = p.box_h + R - 1;
    p.halo_dw = -pad;
    p.halo_dh = -pad;
    return true;
}

static void pick_pipeline_halo_taps(TapParams& p) {
    p.b_taps = g_dbg_halo_btaps > 0 ? g_dbg_halo_btaps : 3;
    if (p.n_taps % p.b_taps) p.b_taps = 1;
}

// shared-memory split of the haloed conv: A ring (haloed tiles), B ring (one weight tile per tap), store staging
static void pick_pipeline_halo(TapParams& p) {
    p.stage_bytes = (p.halo_w * p.halo_h * 128 + 1023) / 1024 * 1024;
    // weights of several taps per B stage: one barrier round then covers 4 * b_taps MMAs (the per-round overhead of
    // the producer / MMA warps, ~400 clk, is what limits narrow tiles otherwise)
    pick_pipeline_halo_taps(p);
    p.b_bytes = p.bn * p.b_taps * 128;
    p.stats_cols = p.stats ? (p.n_total + 31) / 32 * 32 : 0;
    p.nres = 0;
    p.epi_groups = 2;
    p.nout = 2;
    const int budget = TG_SMEM_BYTES - (p.stats ? 2 * p.stats_cols * (int)sizeof(float) + 8192 : 0) - p.nout * TG_A_BYTES;
    p.b_stages = p.b_bytes <= 32768 ? 3 : 2;
    if (budget - p.b_stages * p.b_bytes < 2 * p.stage_bytes) {       // does not fit: fewer taps per weight stage
        p.b_taps = (p.n_taps % 3 == 0 && p.bn * 3 * 128 * 2 + 2 * p.stage_bytes <= budget) ? 3 : 1;
        p.b_bytes = p.bn * p.b_taps * 128;
        p.b_stages = 2;
    }
    p.stages = (budget - p.b_stages * p.b_bytes) / p.stage_bytes;
    if (p.stages > 4) p.stages = 4;
    p.b_ring_bytes = p.b_stages * p.b_bytes;
}

// BatchNorm backward sums fused into a dgrad (TapParams::bn_bwd)
struct BnBwd {
    const float* y = nullptr;          // input of the BatchNorm + ReLU whose output gradient the dgrad produces
    const float* scale = nullptr;      // forward's folded coefficients (ReLU mask: y * scale + shift > 0)
    const float* shift = nullptr;
    const float* mean = nullptr;
    const float* invstd = nullptr;
    double* sums = nullptr;            // [2][Cin] += (sum g, sum g * xhat)
};

struct Epilogue {
    const float* bias = nullptr;
    const float* scale = nullptr;
    const float* shift = nullptr;
    const float* residual = nullptr;
    int relu = 0, round_out = 0;
    double* stats = nullptr;
};

// Filter taps of a convolution as (plane, dh, dw) shifts on stride-`stride` parity planes.
// fwd / wgrad:  in[ho*stride + r - pad]  ->  plane ph = (r-pad) mod stride, shift (r-pad-ph)/stride
static int fill_taps_fwd(TapParams& p, int R, int S, int stride, int pad) {
    PE_REQUIRE(R * S <= TG_MAX_TAPS, "too many filter taps (%d)", R * S);
    PE_REQUIRE(stride == 1 || stride == 2, "stride %d unsupported", stride);
    int t = 0;
    for (int r = 0; r < R; ++r)
        for (int s = 0; s < S; ++s, ++t) {
            const int orr = r - pad, oss = s - pad;
            const int ph = ((orr % stride) + stride) % stride, pw = ((oss % stride) + stride) % stride;
            p.tap_dh[t] = (signed char)((orr - ph) / stride);
            p.tap_dw[t] = (signed char)((oss - pw) / stride);
            p.tap_map[t] = (signed char)(stride == 1 ? 0 : ph * 2 + pw);
            p.tap_b[t] = (signed char)t;
        }
    p.n_taps = t;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// convolution forward:  y[n,ho,wo,co] = sum x[n, ho*s + r - pad, wo*s + q - pad, ci] * w[t][co][ci]
// ---------------------------------------------------------------------------------------------
static int conv_fwd_impl(const float* x, const float* w_tck, float* y, int B, int H, int W, int Cin,
                         int Cout, int R, int S, int stride, int pad, const Epilogue& ep,
                         cudaStream_t stream, const XGeom* xg = nullptr) {
    if (ensure_error_flag()) return 2;
    PE_REQUIRE(Cin % 4 == 0 && Cout % 4 == 0, "conv channels must be multiples of 4 (Cin=%d Cout=%d)", Cin, Cout);
    PE_REQUIRE(!xg || (stride == 1 && pad == 0 && !ep.residual), "conv_fwd: strided views need stride 1, no padding");
    const int Ho = (H + 2 * pad - R) / stride + 1, Wo = (W + 2 * pad - S) / stride + 1;
    TapMaps maps;
    TapParams p;
    init_params(p);
    p.mode = 0;
    p.bn = pick_bn(Cout, ((long long)B * Ho * Wo + TG_BM - 1) / TG_BM);
    const bool halo = stride == 1 && !ep.residual && !xg && setup_conv_halo(p, B, H, W, Cin, R, S, pad);
    if (!halo) {
        choose_box(Wo, Ho, B, TG_BM, &p.box_w, &p.box_h, &p.box_n);
        p.m_rows = p.box_w * p.box_h * p.box_n;
        p.tiles_w = (Wo + p.box_w - 1) / p.box_w;
        p.tiles_h = (Ho + p.box_h - 1) / p.box_h;
        p.tiles_n = (B + p.box_n - 1) / p.box_n;
    }
    p.out_w = Wo;
    p.out_h = Ho;
    p.out_n = B;
    p.chunks = (Cin + TG_BK - 1) / TG_BK;
    if (fill_taps_fwd(p, R, S, stride, pad)) return 1;
    if (halo) {
        // taps as (row, column) offsets inside the haloed tile: tap (r, s) starts r rows / s pixels in
        for (int r = 0; r < R; ++r)
            for (int q = 0; q < S; ++q) {
                p.tap_dh[r * S + q] = (signed char)r;
                p.tap_dw[r * S + q] = (signed char)q;
            }
    }
    const int box[4] = {TG_BK, p.box_w, p.box_h, p.box_n};
    if (halo) {
        const int hbox[4] = {TG_BK, p.halo_w, p.halo_h, 1};
        if (make_nhwc_map(&maps.a[0], x, B, H, W, Cin, 0, 0, 1, hbox)) return 1;
    } else if (xg) {
        if (make_xgeom_map(&maps.a[0], x, B, H, W, *xg, box)) return 1;
    } else if (stride == 1) {
        if (make_nhwc_map(&maps.a[0], x, B, H, W, Cin, 0, 0, 1, box)) return 1;
    } else {
        for (int ph = 0; ph < 2; ++ph)
            for (int pw = 0; pw < 2; ++pw)
                if (make_nhwc_map(&maps.a[ph * 2 + pw], x, B, H, W, Cin, ph, pw, 2, box)) return 1;
    }
    const long long m_tiles = (long long)p.tiles_w * p.tiles_h * p.tiles_n;
    const int n_tiles = (Cout + p.bn - 1) / p.bn;
    p.cta_group = want_pair(p, m_tiles, n_tiles) ? 2 : 1;
    {
        const long long dims[4] = {Cin, Cout, (long long)R * S, 1};
        const long long strides[3] = {Cin, (long long)Cin * Cout, (long long)Cin * Cout * R * S};
        const int bbox[4] = {TG_BK, p.bn / p.cta_group, 1, 1};
        if (make_map(&maps.b[0], w_tck, dims, strides, bbox)) return 1;
    }
    if (make_nhwc_map(&maps.d, y, B, Ho, Wo, Cout, 0, 0, 1, box)) return 1;
    if (ep.residual && make_nhwc_map(&maps.r, ep.residual, B, Ho, Wo, Cout, 0, 0, 1, box)) return 1;
    p.n_total = Cout;
    p.store_mode = TG_STORE_TMA;
    p.bias = ep.bias;
    p.scale = ep.scale;
    p.shift = ep.shift;
    p.residual = ep.residual;
    p.ld_res = Cout;
    p.relu = ep.relu;
    p.round_out = ep.round_out;
    p.stats = ep.stats;
    PE_REQUIRE(!ep.stats || !(ep.scale || ep.bias || ep.residual || ep.relu || ep.round_out),
               "conv_fwd: batch statistics are taken from the raw output (no affine / residual / ReLU / rounding)");
    if (halo) {
        pick_pipeline_halo(p);
        // weight tiles of b_taps consecutive taps per TMA box
        const long long dims[4] = {Cin, Cout, (long long)R * S, 1};
        const long long strides[3] = {Cin, (long long)Cin * Cout, (long long)Cin * Cout * R * S};
        const int bbox[4] = {TG_BK, p.bn, p.b_taps, 1};
        if (make_map(&maps.b[1], w_tck, dims, strides, bbox)) return 1;
    } else {
        pick_pipeline(p, p.n_taps * p.chunks);
    }
    dim3 grid(n_tiles, (unsigned)((m_tiles + p.cta_group - 1) / p.cta_group), 1);
    return launch_tapgemm(maps, p, grid, stream);
}

// ---------------------------------------------------------------------------------------------
// convolution dgrad: dx[n,hi,wi,ci] = sum dy[n,ho,wo,co] * w[t][co][ci], hi = ho*s + r - pad
// `w_tkc` is the transposed pack [tap][Cin][Cout] so both operands stay K-major.
// ---------------------------------------------------------------------------------------------
static int conv_dgrad_impl(const float* dy, const float* w_tkc, float* dx, int B, int H, int W, int Cin,
                           int Cout, int R, int S, int stride, int pad, const float* residual,
                           const unsigned* res_mask, cudaStream_t stream, const BnBwd* bnb = nullptr) {
    if (ensure_error_flag()) return 2;
    PE_REQUIRE(!residual || stride == 1, "conv_dgrad: the residual epilogue is implemented for stride-1 convolutions");
    PE_REQUIRE(!bnb || (!residual && bnb->y && bnb->scale && bnb->shift && bnb->mean && bnb->invstd && bnb->sums &&
                        Cin % 32 == 0),
               "conv_dgrad: fused BatchNorm backward sums need y, scale, shift, mean, invstd, sums, Cin %% 32 == 0 and "
               "no residual");
    PE_REQUIRE(!res_mask || (residual && Cin % 32 == 0 && (reinterpret_cast<uintptr_t>(res_mask) & 15) == 0),
               "conv_dgrad: residual mask needs a residual, Cin %% 32 == 0 and a 16-byte aligned mask");
    const int Ho = (H + 2 * pad - R) / stride + 1, Wo = (W + 2 * pad - S) / stride + 1;
    PE_REQUIRE(stride == 1 || stride == 2, "stride %d unsupported", stride);
    PE_REQUIRE(R * S <= TG_MAX_TAPS, "too many filter taps");
    bool need_zero = false;
    for (int ph = 0; ph < stride; ++ph)
        for (int pw = 0; pw < stride; ++pw) {
            int cnt = 0;
            for (int r = 0; r < R; ++r)
                for (int s = 0; s < S; ++s)
                    if ((ph + pad - r) % stride == 0 && (pw + pad - s) % stride == 0) ++cnt;
            if (cnt == 0) need_zero = true;
        }
    if (need_zero) PE_CHECK_CUDA(cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)B * H * W * Cin, stream));

    for (int ph = 0; ph < stride; ++ph)
        for (int pw = 0; pw < stride; ++pw) {
            TapMaps maps;
            TapParams p;
            init_params(p);
            p.mode = 0;
            const int Hp = (H - ph + stride - 1) / stride, Wp = (W - pw + stride - 1) / stride;
            p.bn = pick_bn(Cin, ((long long)B * Hp * Wp + TG_BM - 1) / TG_BM);
            int t = 0;
            for (int r = 0; r < R; ++r)
                for (int s = 0; s < S; ++s) {
                    const int ah = ph + pad - r, aw = pw + pad - s;
                    if (ah % stride != 0 || aw % stride != 0) continue;
                    p.tap_dh[t] = (signed char)(ah / stride);
                    p.tap_dw[t] = (signed char)(aw / stride);
                    p.tap_map[t] = 0;
                    p.tap_b[t] = (signed char)(r * S + s);
                    ++t;
                }
            if (t == 0) continue;
            p.n_taps = t;
            // stride-1 3x3: dy has the shape of dx, tap (r, s) reads dy at (h + pad - r, w + pad - s), i.e. at
            // row 2*pad - r / column 2*pad - s of the haloed tile
            const bool halo = stride == 1 && !residual && !bnb && setup_conv_halo(p, B, Ho, Wo, Cout, R, S, pad);
            if (halo) {
                for (int k = 0; k < t; ++k) {
                    p.tap_dh[k] = (signed char)(p.tap_dh[k] + pad);
                    p.tap_dw[k] = (signed char)(p.tap_dw[k] + pad);
                }
            } else {
                choose_box(Wp, Hp, B, TG_BM, &p.box_w, &p.box_h, &p.box_n);
                p.m_rows = p.box_w * p.box_h * p.box_n;
                p.tiles_w = (Wp + p.box_w - 1) / p.box_w;
                p.tiles_h = (Hp + p.box_h - 1) / p.box_h;
                p.tiles_n = (B + p.box_n - 1) / p.box_n;
            }
            p.out_w = Wp;
            p.out_h = Hp;
            p.out_n = B;
            p.chunks = (Cout + TG_BK - 1) / TG_BK;
            const int box[4] = {TG_BK, p.box_w, p.box_h, p.box_n};
            const int hbox[4] = {TG_BK, p.halo_w, p.halo_h, 1};
            if (make_nhwc_map(&maps.a[0], dy, B, Ho, Wo, Cout, 0, 0, 1, halo ? hbox : box)) return 1;
            const long long m_tiles = (long long)p.tiles_w * p.tiles_h * p.tiles_n;
            const int n_tiles = (Cin + p.bn - 1) / p.bn;
            p.cta_group = want_pair(p, m_tiles, n_tiles) ? 2 : 1;
            const long long dims[4] = {Cout, Cin, (long long)R * S, 1};
            const long long strides[3] = {Cout, (long long)Cin * Cout, (long long)Cin * Cout * R * S};
            const int bbox[4] = {TG_BK, p.bn / p.cta_group, 1, 1};
            if (make_map(&maps.b[0], w_tkc, dims, strides, bbox)) return 1;
            if (make_nhwc_map(&maps.d, dx, B, H, W, Cin, ph, pw, stride, box)) return 1;
            if (residual && make_nhwc_map(&maps.r, residual, B, H, W, Cin, ph, pw, stride, box)) return 1;
            if (bnb && make_nhwc_map(&maps.r, bnb->y, B, H, W, Cin, ph, pw, stride, box)) return 1;
            p.n_total = Cin;
            p.store_mode = TG_STORE_TMA;
            p.residual = bnb ? bnb->y : residual;
            p.res_mask = res_mask;
            p.ld_res = Cin;
            if (bnb) {
                p.bn_bwd = 1;
                p.scale = bnb->scale;
                p.shift = bnb->shift;
                p.bn_mean = bnb->mean;
                p.bn_invstd = bnb->invstd;
                p.stats = bnb->sums;
            }
            if (halo) {
                pick_pipeline_halo(p);
                for (int k = 0; k < t; ++k) PE_REQUIRE(p.tap_b[k] == k, "conv_dgrad halo: taps out of order");
                const int tbox[4] = {TG_BK, p.bn, p.b_taps, 1};
                if (make_map(&maps.b[1], w_tkc, dims, strides, tbox)) return 1;
            } else {
                pick_pipeline(p, p.n_taps * p.chunks);
            }
            dim3 grid(n_tiles, (unsigned)((m_tiles + p.cta_group - 1) / p.cta_group), 1);
            if (launch_tapgemm(maps, p, grid, stream)) return 2;
        }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// convolution wgrad: dw[t][co][ci] = sum_pixels dy[n,ho,wo,co] * x[n, ho*s + r - pad, wo*s + q - pad, ci]
// Reduction over pixels is split across CTAs and combined with fp32 atomics into a zeroed dw.
// ---------------------------------------------------------------------------------------------
// Split-K factor for a persistent grid: base_items * ks work items should fill whole waves of SMs (an item count
// just above a multiple of the SM count costs a full extra wave).  Prefers the fewest waves that reach 85 %
// occupancy of the last wave, because every extra split adds a tile of atomic reductions.
static int pick_ksplit(int base_items, int k_tiles, int max_waves) {
    const int sms = num_sms();
    int best_ks = 1;
    double best_eff = 0.0;
    for (int waves = 1; waves <= max_waves; ++waves) {
        int ks = (waves * sms) / base_items;
        if (ks < 1) ks = 1;
        if (ks > k_tiles) ks = k_tiles;
        // every split must own at least one k tile
        while (ks > 1 && (long long)(ks - 1) * ((k_tiles + ks - 1) / ks) >= k_tiles) --ks;
        const int items = base_items * ks;
        const int w = (items + sms - 1) / sms;
        const double eff = (double)items / ((double)w * sms);
        if (eff > best_eff + 1e-9) {
            best_eff = eff;
            best_ks = ks;
        }
        if (best_eff >= 0.85) break;
    }
    return best_ks;
}

static int g_dbg_wgrad_halo = 1;      // 0: one tap per work item (mode 1) everywhere; 2: halo + base-offset descriptors

// Stride-1 multi-tap wgrad with a haloed X tile (tap-GEMM mode 2): the dY pixel tile and the X tile (+halo) are
// fetched ONCE per tap group instead of once per tap; every tap accumulates into its own TMEM columns.
static int conv_wgrad_halo_impl(const float* x, const float* dy, float* dw_tck, int B, int H, int W, int Cin,
                                int Cout, int R, int S, int pad, cudaStream_t stream, const XGeom* xg = nullptr) {
    const int Ho = H + 2 * pad - R + 1, Wo = W + 2 * pad - S + 1;
    TapMaps maps;
    TapParams p;
    init_params(p);
    p.mode = 2;
    p.base_offset_mode = g_dbg_wgrad_halo == 2;
    p.bn = Cin >= 128 ? 128 : ((Cin + 31) / 32) * 32;
    p.m_rows = TG_BM;
    p.box_w = 8;
    p.box_h = (Ho % 8 == 0) ? 8 : ((Ho % 7 == 0) ? 7 : 8);
    p.box_n = 1;
    p.halo_w = p.box_w + S - 1;
    p.halo_h = p.box_h + R - 1;
    p.halo_dw = -pad;
    p.halo_dh = -pad;
    p.tiles_w = (Wo + p.box_w - 1) / p.box_w;
    p.tiles_h = (Ho + p.box_h - 1) / p.box_h;
    p.tiles_n = (B + p.box_n - 1) / p.box_n;
    p.pt_total = p.tiles_w * p.tiles_h * p.tiles_n;
    p.out_w = Wo;
    p.out_h = Ho;
    p.out_n = B;
    PE_REQUIRE(R * S <= TG_MAX_TAPS, "too many filter taps (%d)", R * S);
    p.n_taps = R * S;
    for (int r = 0; r < R; ++r)
        for (int q = 0; q < S; ++q) {
            p.tap_dh[r * S + q] = (signed char)r;      // row / column of the tap inside the halo tile
            p.tap_dw[r * S + q] = (signed char)q;
        }
    int max_tg = 2 * TG_MAX_BN / p.bn;                       // accumulators that fit in 512 TMEM columns
    if (max_tg > 8) max_tg = 8;                              // the issue loop is unrolled for <= 8 taps per group
    p.n_groups = (p.n_taps + max_tg - 1) / max_tg;
    p.tg_taps = (p.n_taps + p.n_groups - 1) / p.n_groups;
    p.n_groups = (p.n_taps + p.tg_taps - 1) / p.tg_taps;
    const int pt = p.box_w * p.box_h * p.box_n, hp = p.halo_w * p.halo_h * p.box_n;
    p.a_atom_bytes = (pt * 128 + 1023) / 1024 * 1024;
    p.b_atom_bytes = (hp * 128 + 1023) / 1024 * 1024;
    const int na_max = Cout >= TG_BM ? 4 : (Cout + 31) / 32;
    p.a_region_bytes = na_max * p.a_atom_bytes;
    p.stage_bytes = p.a_region_bytes + (p.bn / 32) * p.b_atom_bytes;
    p.stages = TG_SMEM_BYTES / p.stage_bytes;
    if (p.stages > TG_STAGES) p.stages = TG_STAGES;
    PE_REQUIRE(p.stages >= 2, "wgrad halo: stage of %d bytes does not fit twice", p.stage_bytes);
    PE_REQUIRE(p.box_w == 8 && p.box_n == 1 && p.box_h <= 8 && p.tg_taps <= 8,
               "wgrad halo: the issue loop is written for 8 x (<=8) x 1 pixel tiles and <= 8 taps per group");
    p.nout = 2;
    const int abox[4] = {32, p.box_w, p.box_h, p.box_n};
    const int bbox[4] = {32, p.halo_w, p.halo_h, p.box_n};
    if (make_nhwc_map(&maps.a[0], dy, B, Ho, Wo, Cout, 0, 0, 1, abox, true)) return 1;
    if (xg) {
        if (make_xgeom_map(&maps.b[0], x, B, H, W, *xg, bbox, true)) return 1;
    } else if (make_nhwc_map(&maps.b[0], x, B, H, W, Cin, 0, 0, 1, bbox, true)) {
        return 1;
    }
    p.m_total = Cout;
    p.n_total = Cin;
    p.out = dw_tck;
    p.out_tap_stride = (long long)Cout * Cin;
    p.ldo = Cin;
    p.m64 = (Cout <= 64 && !(g_dbg_flags & 4096)) ? 1 : 0;
    const int mt = (Cout + TG_BM - 1) / TG_BM, nt = (Cin + p.bn - 1) / p.bn;
    const int ks = pick_ksplit(mt * nt * p.n_groups, p.pt_total, 2);
    p.ksplit = ks;
    p.store_mode = ks > 1 ? TG_STORE_ATOMIC : TG_STORE_DIRECT;
    if (ks > 1)
        PE_CHECK_CUDA(cudaMemsetAsync(dw_tck, 0, sizeof(float) * (size_t)R * S * Cout * Cin, stream));
    dim3 grid(nt, mt, p.n_groups * ks);
    return launch_tapgemm(maps, p, grid, stream);
}

static int conv_wgrad_impl(const float* x, const float* dy, float* dw_tck, int B, int H, int W, int Cin,
                           int Cout, int R, int S, int stride, int pad, cudaStream_t stream, const XGeom* xg = nullptr) {
    if (ensure_error_flag()) return 2;
    PE_REQUIRE(Cin % 4 == 0 && Cout % 4 == 0, "conv channels must be multiples of 4");
    PE_REQUIRE(!xg || (stride == 1 && pad == 0 && R * S > 1 && Cin % 32 == 0 && Cout % 32 == 0 &&
                       (H - R + 1) % 8 == 0 && (W - S + 1) % 8 == 0),
               "conv_wgrad: strided views go through the haloed multi-tap path (stride 1, no padding, 8 x 8 tiles)");
    if ((g_dbg_wgrad_halo || xg) && stride == 1 && R * S > 1 && Cin % 32 == 0 && Cout % 32 == 0)
        return conv_wgrad_halo_impl(x, dy, dw_tck, B, H, W, Cin, Cout, R, S, pad, stream, xg);
    const int Ho = (H + 2 * pad - R) / stride + 1, Wo = (W + 2 * pad - S) / stride + 1;
    TapMaps maps;
    TapParams p;
    init_params(p);
    p.mode = 1;
    p.bn = (Cin >= 256 && g_dbg_max_bn >= 256) ? 256 : (Cin >= 128 ? 128 : ((Cin + 31) / 32) * 32);
    p.m_rows = TG_BM;
    choose_box(Wo, Ho, B, TG_BK, &p.box_w, &p.box_h, &p.box_n, true);
    p.tiles_w = (Wo + p.box_w - 1) / p.box_w;
    p.tiles_h = (Ho + p.box_h - 1) / p.box_h;
    p.tiles_n = (B + p.box_n - 1) / p.box_n;
    p.pt_total = p.tiles_w * p.tiles_h * p.tiles_n;
    p.out_w = Wo;
    p.out_h = Ho;
    p.out_n = B;
    if (fill_taps_fwd(p, R, S, stride, pad)) return 1;
    const int box[4] = {32, p.box_w, p.box_h, p.box_n};
    if (make_nhwc_map(&maps.a[0], dy, B, Ho, Wo, Cout, 0, 0, 1, box, true)) return 1;
    if (stride == 1) {
        if (make_nhwc_map(&maps.b[0], x, B, H, W, Cin, 0, 0, 1, box, true)) return 1;
    } else {
        for (int ph = 0; ph < 2; ++ph)
            for (int pw = 0; pw < 2; ++pw)
                if (make_nhwc_map(&maps.b[ph * 2 + pw], x, B, H, W, Cin, ph, pw, 2, box, true)) return 1;
    }
    p.m_total = Cout;
    p.n_total = Cin;
    p.out = dw_tck;
    p.out_tap_stride = (long long)Cout * Cin;
    p.ldo = Cin;
    const int mt = (Cout + TG_BM - 1) / TG_BM, nt = (Cin + p.bn - 1) / p.bn;
    const int ks = pick_ksplit(mt * nt * p.n_taps, p.pt_total, 3);
    p.ksplit = ks;
    p.store_mode = ks > 1 ? TG_STORE_ATOMIC : TG_STORE_DIRECT;
    if (ks > 1)
        PE_CHECK_CUDA(cudaMemsetAsync(dw_tck, 0, sizeof(float) * (size_t)R * S * Cout * Cin, stream));
    pick_pipeline(p, (p.pt_total + ks - 1) / ks);
    dim3 grid(nt, mt, p.n_taps * ks);
    return launch_tapgemm(maps, p, grid, stream);
}

// ---------------------------------------------------------------------------------------------
// dense layers:  y[M,N] = x[M,K] * w[N,K]^T (+bias)(relu) ;  dw[N,K] = dy[M,N]^T x[M,K]
// ---------------------------------------------------------------------------------------------
static int linear_fwd_impl(const float* x, int ldx, const float* w, int ldw, float* y, int ldy, int M, int N,
                           int K, const Epilogue& ep, int accumulate_into_y, cudaStream_t stream) {
    if (ensure_error_flag()) return 2;
    PE_REQUIRE(ldx % 4 == 0 && ldw % 4 == 0, "linear: ldx/ldw must be multiples of 4 (got %d, %d)", ldx, ldw);
    PE_REQUIRE(N <= ldy && K <= ldx && K <= ldw, "linear: N = %d / K = %d exceed the row strides (ldy %d, ldx %d, ldw %d)",
               N, K, ldy, ldx, ldw);
    TapMaps maps;
    TapParams p;
    init_params(p);
    p.mode = 0;
    p.bn = pick_bn(N, (M + TG_BM - 1) / TG_BM);
    p.box_w = TG_BM;
    p.box_h = p.box_n = 1;
    p.m_rows = TG_BM;
    p.tiles_w = (M + TG_BM - 1) / TG_BM;
    p.tiles_h = p.tiles_n = 1;
    p.out_w = M;
    p.out_h = p.out_n = 1;
    p.n_taps = 1;
    p.chunks = (K + TG_BK - 1) / TG_BK;
    {
        const long long dims[4] = {K, M, 1, 1};
        const long long strides[3] = {ldx, (long long)ldx * M, (long long)ldx * M};
        const int box[4] = {TG_BK, TG_BM, 1, 1};
        if (make_map(&maps.a[0], x, dims, strides, box)) return 1;
    }
    const int n_tiles = (N + p.bn - 1) / p.bn;
    p.cta_group = want_pair(p, p.tiles_w, n_tiles) ? 2 : 1;
    {
        const long long dims[4] = {K, N, 1, 1};
        const long long strides[3] = {ldw, (long long)ldw * N, (long long)ldw * N};
        const int box[4] = {TG_BK, p.bn / p.cta_group, 1, 1};
        if (make_map(&maps.b[0], w, dims, strides, box)) return 1;
    }
    p.n_total = N;
    p.bias = ep.bias;
    p.scale = ep.scale;
    p.shift = ep.shift;
    p.relu = ep.relu;
    p.round_out = ep.round_out;
    p.stats = ep.stats;
    p.out = y;
    p.ldo = ldy;
    if (accumulate_into_y) {
        p.residual = y;
        p.ld_res = ldy;
    } else if (ep.residual) {
        p.residual = ep.residual;
        p.ld_res = ldy;
    }
    const bool tma_ok = (ldy % 4 == 0) && (N % 4 == 0) && !accumulate_into_y && !(p.residual && (ldy % 4));
    if (tma_ok) {
        const long long dims[4] = {N, M, 1, 1};
        const long long strides[3] = {ldy, (long long)ldy * M, (long long)ldy * M};
        const int box[4] = {32, TG_BM, 1, 1};
        if (make_map(&maps.d, y, dims, strides, box)) return 1;
        if (p.residual && make_map(&maps.r, p.residual, dims, strides, box)) return 1;
        p.store_mode = TG_STORE_TMA;
    } else {
        PE_REQUIRE(!p.residual || (ldy % 4 == 0 && N % 4 == 0),
                   "linear: residual/accumulate needs N and ldy multiples of 4");
        p.store_mode = TG_STORE_DIRECT;
    }
    PE_REQUIRE(!ep.stats || (p.store_mode == TG_STORE_TMA && !(ep.scale || ep.bias || ep.relu || ep.round_out || p.residual)),
               "linear_fwd: statistics need the TMA store path and a raw (un-activated) output");
    pick_pipeline(p, p.chunks);
    dim3 grid(n_tiles, (p.tiles_w + p.cta_group - 1) / p.cta_group, 1);
    return launch_tapgemm(maps, p, grid, stream);
}

static int linear_wgrad_impl(const float* x, int ldx, const float* dy, int lddy, float* dw, int lddw, int M,
                             int N, int K, cudaStream_t stream) {
    if (ensure_error_flag()) return 2;
    PE_REQUIRE(ldx % 4 == 0 && lddy % 4 == 0, "linear wgrad: ldx/lddy must be multiples of 4");
    TapMaps maps;
    TapParams p;
    init_params(p);
    p.mode = 1;
    p.bn = (K >= 256 && g_dbg_max_bn >= 256) ? 256 : ((K + 31) / 32) * 32;
    p.m_rows = TG_BM;
    p.box_w = TG_BK;
    p.box_h = p.box_n = 1;
    p.tiles_w = (M + TG_BK - 1) / TG_BK;
    p.tiles_h = p.tiles_n = 1;
    p.pt_total = p.tiles_w;
    p.out_w = M;
    p.out_h = p.out_n = 1;
    p.n_taps = 1;
    {
        const long long dims[4] = {N, M, 1, 1};
        const long long strides[3] = {lddy, (long long)lddy * M, (long long)lddy * M};
        const int box[4] = {32, TG_BK, 1, 1};
        if (make_map(&maps.a[0], dy, dims, strides, box, true)) return 1;
    }
    {
        const long long dims[4] = {K, M, 1, 1};
        const long long strides[3] = {ldx, (long long)ldx * M, (long long)ldx * M};
        const int box[4] = {32, TG_BK, 1, 1};
        if (make_map(&maps.b[0], x, dims, strides, box, true)) return 1;
    }
    p.m_total = N;
    p.n_total = K;
    p.out = dw;
    p.out_tap_stride = 0;
    p.ldo = lddw;
    const int mt = (N + TG_BM - 1) / TG_BM, nt = (K + p.bn - 1) / p.bn;
    const int ks = pick_ksplit(mt * nt, p.pt_total, 3);
    p.ksplit = ks;
    p.store_mode = ks > 1 ? TG_STORE_ATOMIC : TG_STORE_DIRECT;
    if (ks > 1) PE_CHECK_CUDA(cudaMemset2DAsync(dw, sizeof(float) * lddw, 0, sizeof(float) * K, N, stream));
    pick_pipeline(p, (p.pt_total + ks - 1) / ks);
    dim3 grid(nt, mt, ks);
    return launch_tapgemm(maps, p, grid, stream);
}

}  // namespace pe

// =============================================================================================
// C ABI
// =============================================================================================
using namespace pe;

extern "C" {

const char* pe_last_error(void) { return pe::get_error(); }

int pe_version(void) { return 100; }

void pe_debug_flags(int flags) { g_dbg_flags = flags; }

void pe_debug_pipeline(int stages, int nout) {
    g_dbg_stages = stages;
    g_dbg_nout = nout;
}

void pe_debug_max_bn(int bn) { g_dbg_max_bn = bn > 0 ? bn : 256; }

void pe_debug_min_bn(int bn) { g_dbg_min_bn = bn >= 16 ? bn : 32; }

void pe_debug_wgrad_halo(int mode) { g_dbg_wgrad_halo = mode; }

void pe_debug_residual_tma(int on) { g_dbg_res_tma = on; }

void pe_debug_epilogue_groups(int groups) { g_dbg_epi_groups = groups; }

void pe_debug_pdl(int mask) { pe::g_pdl = mask & 3; }

void pe_debug_cta_group(int mode) { g_dbg_cta_group = mode; }

int pe_debug_fast_div(int n, int d) {
    // host-side replay of the kernel's divide-free work-list decode (fast_div in pe_tapgemm.cu) for the CPU tests
    unsigned fd[2];
    pe::fast_div_of(fd, d);
    if (!fd[0]) return n;
    return static_cast<int>(((static_cast<unsigned long long>(static_cast<unsigned>(n)) * fd[0]) >> 32) >> fd[1]);
}

void pe_set_sm_reserve(int sms) { pe::g_sm_reserve = sms > 0 ? sms : 0; }

void pe_debug_conv_halo(int on) {
    g_dbg_conv_halo = on & 1;
    g_dbg_halo_btaps = on >> 4;      // optional: taps per weight stage in bits 4..
}

void pe_debug_desc_override(int a_lbo, int a_sbo, int b_lbo, int b_sbo) {
    g_dbg_desc[0] = a_lbo;
    g_dbg_desc[1] = a_sbo;
    g_dbg_desc[2] = b_lbo;
    g_dbg_desc[3] = b_sbo;
}

int pe_device_error(void) {
    if (!g_error_flag) return 0;
    int v = 0;
    if (cudaMemcpy(&v, g_error_flag, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return v;
}

int pe_poison_on_error(float* buf, long long n, void* stream) {
    if (ensure_error_flag()) return 2;
    if (n <= 0) return 0;
    poison_on_error_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(g_error_flag, buf, n);
    PE_LAUNCH_CHECK();
    return 0;
}

void pe_device_error_clear(void) {
    if (g_error_flag) cudaMemset(g_error_flag, 0, sizeof(int));
    pe::lstm_seq_reset();
}

int pe_conv2d_fwd(const float* x, const float* w_tck, float* y, int B, int H, int W, int Cin, int Cout, int R,
                  int S, int stride, int pad, const float* scale, const float* shift, const float* residual,
                  int relu, int round_out, double* stats, void* stream) {
    Epilogue ep;
    ep.scale = scale;
    ep.shift = shift;
    ep.residual = residual;
    ep.relu = relu;
    ep.round_out = round_out;
    ep.stats = stats;
    return conv_fwd_impl(x, w_tck, y, B, H, W, Cin, Cout, R, S, stride, pad, ep, (cudaStream_t)stream);
}

int pe_conv2d_dgrad(const float* dy, const float* w_tkc, float* dx, int B, int H, int W, int Cin, int Cout,
                    int R, int S, int stride, int pad, const float* residual, const unsigned* res_maskbits,
                    void* stream) {
    return conv_dgrad_impl(dy, w_tkc, dx, B, H, W, Cin, Cout, R, S, stride, pad, residual, res_maskbits,
                           (cudaStream_t)stream);
}

int pe_conv2d_dgrad_bn(const float* dy, const float* w_tkc, float* dx, int B, int H, int W, int Cin, int Cout, int R,
                       int S, int stride, int pad, const float* bn_y, const float* bn_scale, const float* bn_shift,
                       const float* bn_mean, const float* bn_invstd, double* bn_sums, void* stream) {
    BnBwd b;
    b.y = bn_y;
    b.scale = bn_scale;
    b.shift = bn_shift;
    b.mean = bn_mean;
    b.invstd = bn_invstd;
    b.sums = bn_sums;
    return conv_dgrad_impl(dy, w_tkc, dx, B, H, W, Cin, Cout, R, S, stride, pad, nullptr, nullptr, (cudaStream_t)stream,
                           &b);
}

int pe_conv2d_wgrad(const float* x, const float* dy, float* dw_tck, int B, int H, int W, int Cin, int Cout,
                    int R, int S, int stride, int pad, void* stream) {
    return conv_wgrad_impl(x, dy, dw_tck, B, H, W, Cin, Cout, R, S, stride, pad, (cudaStream_t)stream);
}

// ---- space-to-depth stem (torchvision resnet.py:197 conv1 = Conv2d(3, 64, 7, stride 2, pad 3)) ---------------------
// out(ho, wo) = sum_{u, v in 0..7} in(2 ho + u - 4, 2 wo + v - 4) w8[u][v]   (w8 = the 7x7 filter behind a zero row and
// column) = a 4 x 4 stride-1 convolution over the 2 x 2 space-to-depth image (12 channels).  Four horizontally adjacent
// taps are 48 CONTIGUOUS floats of that NHWC tensor, so the A operand of filter row U is a TMA box over a view whose
// pixel stride is 12 floats and whose channel extent is 48 (64 with TMA's zero fill): conv1 becomes the tap-GEMM's
// ordinary conv mode with 4 taps x 64 channels -- no im2col matrix (2 GB per 256-frame step) in HBM.
static XGeom stem_geom(int H, int W) {
    XGeom g;
    g.c_extent = 48;
    g.sw = 12;
    g.sh = 12ll * (W / 2 + 3);
    g.sn = g.sh * (H / 2 + 3);
    return g;
}

int pe_stem_conv_fwd(const float* s2d, const float* w_s2d, float* y, int B, int H, int W, int Cout, const float* scale,
                     const float* shift, int relu, int round_out, double* stats, void* stream) {
    PE_REQUIRE(H % 2 == 0 && W % 2 == 0 && H >= 16 && W >= 16, "stem: image %d x %d must have even sides", H, W);
    Epilogue ep;
    ep.scale = scale;
    ep.shift = shift;
    ep.relu = relu;
    ep.round_out = round_out;
    ep.stats = stats;
    const XGeom g = stem_geom(H, W);
    return conv_fwd_impl(s2d, w_s2d, y, B, H / 2 + 3, W / 2, 64, Cout, 4, 1, 1, 0, ep, (cudaStream_t)stream, &g);
}

int pe_stem_conv_wgrad(const float* s2d, const float* dy, float* dw_s2d, int B, int H, int W, int Cout, void* stream) {
    PE_REQUIRE(H % 16 == 0 && W % 16 == 0, "stem wgrad: image %d x %d must have sides that are multiples of 16", H, W);
    const XGeom g = stem_geom(H, W);
    return conv_wgrad_impl(s2d, dy, dw_s2d, B, H / 2 + 3, W / 2, 64, Cout, 4, 1, 1, 0, (cudaStream_t)stream, &g);
}

int pe_linear_fwd(const float* x, int ldx, const float* w, int ldw, const float* bias, const float* scale, float* y,
                  int ldy, int M, int N, int K, int relu, int accumulate, int round_out, double* stats,
                  void* stream) {
    Epilogue ep;
    if (scale) {
        PE_REQUIRE(bias != nullptr, "linear_fwd: scale needs a shift vector in `bias`");
        ep.scale = scale;
        ep.shift = bias;
    } else {
        ep.bias = bias;
    }
    ep.relu = relu;
    ep.round_out = round_out;
    ep.stats = stats;
    return linear_fwd_impl(x, ldx, w, ldw, y, ldy, M, N, K, ep, accumulate, (cudaStream_t)stream);
}

int pe_linear_wgrad(const float* x, int ldx, const float* dy, int lddy, float* dw, int lddw, int M, int N, int K,
                    void* stream) {
    return linear_wgrad_impl(x, ldx, dy, lddy, dw, lddw, M, N, K, (cudaStream_t)stream);
}

}  // extern "C"

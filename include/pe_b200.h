/*
 * pe_b200.h -- C ABI of the B200-native pose-estimator hot path (libpe_b200.so).
 *
 * The reference (cremebrule/rgb-proprioceptive-pose-estimator) has no FFI: its hot path is a chain
 * of PyTorch library calls.  Each entry point below replaces one such call site; the host mirror in
 * rgb-proprioceptive-pose-estimator_b200/{models,util} binds them with ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to fp32 (double where stated) owned by the caller;
 *   - activations are NHWC ("pixels x channels"), channels innermost; `ld*` are row strides in elements;
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous, allocate nothing on the
 *     device (except one 4-byte sticky error flag on first use) and are CUDA-graph capturable;
 *   - return 0 on success; non-zero on failure with a message in pe_last_error();
 *   - device-side pipeline timeouts set a sticky flag readable with pe_device_error() (synchronises).
 *   - all dense contractions run on tcgen05 tensor cores with TF32 operands and fp32 accumulation.
 */
#ifndef PE_B200_H_
#define PE_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

const char* pe_last_error(void);
int pe_device_error(void);      /* bitmask: 1 conv producer, 2 wgrad producer, 4 MMA issuer, 8 epilogue */
void pe_device_error_clear(void);
/* If the sticky flag is set, overwrite buf[0..n) with NaN (asynchronous, graph capturable): a training step's loss
 * or a rollout step's pose then shows the fault even when the caller never polls pe_device_error().  Guards the
 * results handed back at util/learn_utils.py:182 (loss.item()) and :448 (pose read-back). */
int pe_poison_on_error(float* buf, long long n, void* stream);
int pe_version(void);
/* debug: override UMMA shared-memory descriptor strides (bytes; <0 restores the default) */
void pe_debug_desc_override(int a_lbo, int a_sbo, int b_lbo, int b_sbo);
/* debug: tap-GEMM pipeline ablations (1 no TMA store, 2 no staging write, 4 no A loads, 8 no B loads); 0 = normal */
void pe_debug_flags(int flags);
/* CTA pairs (tcgen05 cta_group::2, 256-row MMAs over two SMs): 0 automatic (wide tiles of large launches), 1 never,
 * 2 whenever the launch allows it */
void pe_debug_cta_group(int mode);
/* test hook (host only): n / d as the tap-GEMM's divide-free work-list decode computes it (multiply-shift by constants
 * prepared per launch); exact for 0 <= n < 2^31, d >= 1 */
int pe_debug_fast_div(int n, int d);
/* Number of SMs the persistent tap-GEMM grids leave unused (0 = none).  The data-parallel trainer sets it during the
 * backward pass, while NCCL gradient all-reduce kernels share the GPU: a one-CTA-per-SM grid that needs every SM
 * would wait for the collective's CTAs and finish with a straggler wave. */
void pe_set_sm_reserve(int sms);
/* debug: force the smem split of the tap-GEMM (ring stages x 32 KB + nout x 16 KB <= 223 KB); 0 = default */
void pe_debug_pipeline(int stages, int nout);
/* debug: cap the tile width (128 or 256 columns) */
void pe_debug_max_bn(int bn);
/* debug: narrowest tile (32 by default) the few-tiles heuristic may pick for launches that cannot fill the SMs */
void pe_debug_min_bn(int bn);
/* debug: stride-1 multi-tap wgrad path: 0 one tap per work item, 1 haloed tile (default), 2 haloed tile with the
 * descriptors' base-offset field set */
void pe_debug_wgrad_halo(int mode);
/* debug: 0 = epilogue reads residual rows with plain global loads instead of TMA-prefetched tiles */
void pe_debug_residual_tma(int on);
/* debug: force the number of epilogue warp groups of the tap-GEMM (2 or 4); 0 = automatic; 6 = automatic without the
 * four-group rule for residual-prefetch dgrads on CTA pairs */
void pe_debug_epilogue_groups(int groups);
/* debug: haloed-tile path for 3x3 stride-1 forward / dgrad with <= 128 input channels (0 = per-tap boxes) */
void pe_debug_conv_halo(int on);
/* debug: programmatic dependent launch mask -- bit 0: the tap-GEMM launches carry the attribute (default: its
 * prologue overlaps the tail of the previous kernel; it waits for that kernel before touching global memory),
 * bit 1: the streaming / elementwise kernels too.  Same as the environment variable PE_B200_PDL.              */
void pe_debug_pdl(int mask);

/* ---- convolutions: torchvision resnet.py:143-163,266-282 Conv2d calls reached from
 *      models/naive.py:316 and models/time_sensitive.py:185,472 (bias-free, NHWC here) ------------
 * w_tck : weights packed [R*S][Cout][Cin]   (pe_pack_conv_weight)
 * w_tkc : weights packed [R*S][Cin][Cout]
 * fwd epilogue: y = act(acc*scale[c] + shift[c] + residual) when scale != NULL (eval-mode folded BN),
 *               y = acc otherwise; stats (double[2*Cout]: sum, sum of squares of acc) accumulated
 *               when non-NULL (training-mode BatchNorm statistics); round_out rounds y to TF32.   */
int pe_conv2d_fwd(const float* x, const float* w_tck, float* y, int B, int H, int W, int Cin, int Cout, int R,
                  int S, int stride, int pad, const float* scale, const float* shift, const float* residual,
                  int relu, int round_out, double* stats, void* stream);
/* dgrad epilogue (1x1 stride-1 only): dx += residual * (res_maskbits ? mask : 1) -- the identity branch of a
 * residual join is added (and ReLU-masked from the join's bit mask) while dx is still in registers.   */
int pe_conv2d_dgrad(const float* dy, const float* w_tkc, float* dx, int B, int H, int W, int Cin, int Cout,
                    int R, int S, int stride, int pad, const float* residual, const unsigned* res_maskbits,
                    void* stream);
/* dgrad whose output dx is the gradient of a BatchNorm + ReLU activation a = relu(bn(y)) (bn1 -> conv2, bn2 -> conv3 of
 * a Bottleneck, torchvision resnet.py:143-163): the epilogue also accumulates that BatchNorm's backward sums
 * bn_sums[2*Cin] += (sum g, sum g * xhat), g = dx * (y * bn_scale + bn_shift > 0), xhat = (y - mean) * invstd, from the
 * y tile it prefetches by TMA -- pass 1 of the BatchNorm backward (pe_bn_bwd_reduce) without its own launch and without
 * a second read of dx.  bn_scale / bn_shift are the forward's folded coefficients; the caller zeroes bn_sums. */
int pe_conv2d_dgrad_bn(const float* dy, const float* w_tkc, float* dx, int B, int H, int W, int Cin, int Cout, int R,
                       int S, int stride, int pad, const float* bn_y, const float* bn_scale, const float* bn_shift,
                       const float* bn_mean, const float* bn_invstd, double* bn_sums, void* stream);
int pe_conv2d_wgrad(const float* x, const float* dy, float* dw_tck, int B, int H, int W, int Cin, int Cout,
                    int R, int S, int stride, int pad, void* stream);
/* OIHW (checkpoint layout, util/model_utils.py:136-141) <-> packed tap-major layouts */
int pe_pack_conv_weight(const float* w_oihw, float* w_tck, float* w_tkc, int Cout, int Cin, int R, int S,
                        int round_tf32, void* stream);
int pe_unpack_conv_wgrad(const float* dw_tck, float* dw_oihw, int Cout, int Cin, int R, int S, int accumulate,
                         void* stream);
/* every conv layer of a model in one launch.  table_dev: DEVICE array of n_layers rows of 8 int64:
 * {w_oihw pointer, w_tck pointer, w_tkc pointer or 0, Cout, Cin, R*S, first block index, Cout*Cin*R*S};
 * block b packs pe_pack_block_elems() consecutive elements of the layer with first block <= b            */
int pe_pack_conv_weights_batched(const long long* table_dev, int n_layers, int total_blocks, int round_tf32,
                                 void* stream);
int pe_pack_block_elems(void);
/* 7x7/2 stem (Cin = 3): NCHW image -> im2col rows [B*Ho*Wo][ldc], columns ordered (c, r, s) like OIHW */
int pe_im2col_stem(const float* img_nchw, float* col, int B, int C, int H, int W, int R, int S, int stride,
                   int pad, int ldc, int round_tf32, void* stream);
/* The same stem (torchvision resnet.py:197 conv1 = Conv2d(3, 64, 7, stride 2, padding 3), reached from
 * models/naive.py:316 / models/time_sensitive.py:185,472) WITHOUT the im2col matrix: the 7x7/2 convolution is a 4x4/1
 * convolution over the 2x2 space-to-depth image (12 channels); four horizontally adjacent taps are 48 contiguous floats
 * of that NHWC tensor, so the tap-GEMM reads its operand through a TMA view with overlapping pixel rows.
 *   pe_stem_s2d_pack     img NCHW (3,H,W) -> s2d [B][H/2+3][W/2+3][12] (zero border written), optionally TF32-rounded
 *   pe_stem_pack_weight  OIHW [Cout][3][7][7] -> [4][Cout][64];  pe_stem_unpack_wgrad: the inverse for the gradient
 *   pe_stem_conv_fwd     y [B*(H/2)*(W/2)][Cout] = conv1(img) with the conv epilogues of pe_conv2d_fwd
 *   pe_stem_conv_wgrad   dw [4][Cout][64] from s2d and dy                                                       */
int pe_stem_s2d_pack(const float* img_nchw, float* s2d, int B, int H, int W, int round_tf32, void* stream);
int pe_stem_pack_weight(const float* w_oihw, float* w_s2d, int Cout, int round_tf32, void* stream);
int pe_stem_unpack_wgrad(const float* dw_s2d, float* dw_oihw, int Cout, void* stream);
int pe_stem_conv_fwd(const float* s2d, const float* w_s2d, float* y, int B, int H, int W, int Cout, const float* scale,
                     const float* shift, int relu, int round_out, double* stats, void* stream);
int pe_stem_conv_wgrad(const float* s2d, const float* dy, float* dw_s2d, int B, int H, int W, int Cout, void* stream);

/* ---- dense layers: nn.Linear / nn.LSTM projections (models/naive.py:274,343-345,
 *      models/time_sensitive.py:126-131,418-423) -------------------------------------------------
 * y[M,N] = act(x[M,K] w[N,K]^T * scale[n] + bias[n]); scale NULL -> 1; accumulate!=0 adds into the
 * existing y; stats (double[2*N]) accumulates column sums / sums of squares of the raw product
 * (the 7x7 stem runs through this entry as im2col + GEMM and needs BatchNorm statistics).          */
int pe_linear_fwd(const float* x, int ldx, const float* w, int ldw, const float* bias, const float* scale, float* y,
                  int ldy, int M, int N, int K, int relu, int accumulate, int round_out, double* stats,
                  void* stream);
/* dw[N,K] = dy[M,N]^T x[M,K] */
int pe_linear_wgrad(const float* x, int ldx, const float* dy, int lddy, float* dw, int lddw, int M, int N, int K,
                    void* stream);
/* dst[c][r] = src[r][c] (rows x cols), optional TF32 rounding; dst rows beyond `cols` untouched */
int pe_transpose(const float* src, int lds, float* dst, int ldd, int rows, int cols, int round_tf32, void* stream);
/* dst[r][0:cols] = src[r][0:cols] with independent row strides (concat / slicing), optional rounding */
int pe_copy_cols(const float* src, int lds, float* dst, int ldd, int rows, int cols, int round_tf32, void* stream);
/* out[r][c] = alpha*a[r][c] + beta*b[r][c] (out may alias a or b): measurement_diff = pre_out - x0bar
 * (models/naive.py:95, models/time_sensitive.py:224) and gradient joins of the two-headed models */
int pe_axpby_cols(const float* a, int lda, const float* b, int ldb, float* out, int ldo, int rows, int cols,
                  float alpha, float beta, int round_tf32, void* stream);
/* out[c] (+)= sum_r x[r][c] : bias gradients */
int pe_colsum(const float* x, int ldx, float* out, int rows, int cols, int accumulate, void* stream);
/* dz = dy * (y > 0) */
int pe_relu_bwd(const float* dy, int lddy, const float* y, int ldy, float* dz, int lddz, int rows, int cols,
                void* stream);

/* ---- BatchNorm2d (torchvision resnet.py:134-138,198; torch semantics: biased var to normalise,
 *      unbiased into running_var, momentum 0.1, eps 1e-5) ----------------------------------------- */
/* standalone statistics (same result as the conv epilogue): stats = double[2*C], zeroed by the caller */
int pe_bn_stats(const float* y, long long P, int C, double* stats, void* stream);
/* train: from stats -> mean, invstd, scale=gamma*invstd, shift=beta-mean*scale; updates running stats.
 * eval (stats == NULL): scale/shift from running stats.  The caller zeroes `stats` before each step.   */
int pe_bn_finalize(double* stats, const float* gamma, const float* beta, float* running_mean,
                   float* running_var, float* scale, float* shift, float* mean, float* invstd, long long count,
                   float momentum, float eps, int C, void* stream);
/* out = act(y*scale[c] + shift[c] (+ residual)) */
int pe_bn_apply(const float* y, const float* scale, const float* shift, const float* residual, float* out,
                long long P, int C, int relu, int round_tf32, void* stream);
/* train-mode forward in ONE pass: scale / shift rebuilt from `stats` by every block (published with mean /
 * invstd for the backward pass), running statistics and num_batches_tracked updated, out = act(y*scale +
 * shift (+ residual)).  maskbits (optional, ceil(P*C/128)*4 words, 16-byte aligned): the (out > 0) mask,
 * one bit per element -- float4 index i owns bit (i & 31) of words [(i >> 5)*4 + component].           */
int pe_bn_train_apply(const float* y, const double* stats, const float* gamma, const float* beta,
                      float* running_mean, float* running_var, long long* num_batches_tracked, float* scale,
                      float* shift, float* mean, float* invstd, const float* residual, float* out,
                      unsigned* maskbits, long long P, int C, float momentum, float eps, int relu, int round_tf32,
                      void* stream);
/* backward pass 1: g = (dout + dout2)*(out>0 if relu)*(maskbits if given); sums = double[2*C] += (sum g, sum g*xhat).
 * `out` may be NULL for a BN without residual input: the ReLU mask is then recomputed from
 * y*mask_scale + mask_shift (the forward's folded scale / shift), saving one full read of the activations.
 * dout2 (optional) is the second gradient branch of a residual join, summed on the fly.              */
int pe_bn_bwd_reduce(const float* dout, const float* dout2, const float* out, const float* y, const float* mean,
                     const float* invstd, const float* mask_scale, const float* mask_shift,
                     const unsigned* maskbits, double* sums, long long P, int C, int relu, void* stream);
/* backward pass 2: dy = gamma*invstd*(g - sum_g/P - xhat*sum_gx/P); dres = g (optional);
 * dgamma = sum_gx, dbeta = sum_g (written or accumulated).  The caller zeroes `sums` before pass 1.   */
int pe_bn_bwd_apply(const float* dout, const float* dout2, const float* out, const float* y, const float* mean,
                    const float* invstd, const float* gamma, const float* mask_scale, const float* mask_shift,
                    const unsigned* maskbits, double* sums, float* dy, float* dres,
                    int dres_accumulate, float* dgamma, float* dbeta, int param_accumulate, long long P, int C,
                    int relu, int round_tf32, void* stream);

/* ---- stem pooling + auxiliary BN1 branch (torchvision resnet.py:268-272; models/naive.py:223-231,
 *      models/time_sensitive.py:377-385: Conv2d(64,1,1) + MaxPool2d(2) + Flatten on post-ReLU bn1) -- */
int pe_maxpool3x3s2_fwd(const float* x, float* y, unsigned char* argmax, int B, int H, int W, int C, void* stream);
/* aux_* (optional, all or none): the gradient of the aux branch, which reads the same activation, is added in the
 * same pass -- da1 += [aux_argmax(window) == pixel] * aux_dout[b][window] * aux_w[c] -- so dx is written once  */
int pe_maxpool3x3s2_bwd(const float* dy, const float* dy2, const unsigned char* argmax, float* dx, int accumulate,
                        int B, int H, int W, int C, const float* aux_dout, int aux_lddo,
                        const unsigned char* aux_argmax, const float* aux_w, void* stream);
int pe_avgpool_fwd(const float* x, float* y, int ldy, int B, int HW, int C, int round_tf32, void* stream);
int pe_avgpool_bwd(const float* dy, int lddy, float* dx, int B, int HW, int C, void* stream);
int pe_aux_fwd(const float* a1, const float* w, const float* bias, float* out, int ldo, unsigned char* argmax,
               int B, int H, int W, int C, int round_tf32, void* stream);
/* da1 (+)= scatter(dout) * w (da1 may be NULL: the scatter is then left to pe_maxpool3x3s2_bwd's aux term);
 * dw[C] += ..., db[1] += ... (dw/db may be NULL: frozen aux conv, td model) */
int pe_aux_bwd(const float* dout, int lddo, const unsigned char* argmax, const float* a1, const float* w,
               float* da1, int accumulate, float* dw, float* db, int B, int H, int W, int C, void* stream);
/* The aux conv's own gradients when bn1's output was never materialised (pe_stem_post_train): the activation at
 * every arg-max pixel is rebuilt as relu(y*scale + shift) (rounded to TF32 when round_tf32) from conv1's output y. */
int pe_aux_bwd_params(const float* dout, int lddo, const unsigned char* argmax, const float* y, const float* scale,
                      const float* shift, int round_tf32, float* dw, float* db, int B, int H, int W, int C,
                      void* stream);
/* Training-mode stem tail in one pass over conv1's output y [B,H,W,64] (torchvision resnet.py:268-272 bn1, relu,
 * maxpool + the aux branch above): pe_bn_train_apply (batch statistics from `stats`, running-stat update, ReLU),
 * pe_maxpool3x3s2_fwd and pe_aux_fwd without ever writing the normalised activation.  pool [B,H/2,W/2,64] and
 * pool_argmax as pe_maxpool3x3s2_fwd; aux_* as pe_aux_fwd (aux_w NULL: no aux branch); argmax maps may be NULL.  */
int pe_stem_post_train(const float* y, const double* stats, const float* gamma, const float* beta,
                       float* running_mean, float* running_var, long long* num_batches_tracked, float* scale,
                       float* shift, float* mean, float* invstd, float* pool, unsigned char* pool_argmax,
                       const float* aux_w, const float* aux_bias, float* aux_out, int ld_aux,
                       unsigned char* aux_argmax, int B, int H, int W, int C, float momentum, float eps,
                       int round_tf32, int aux_round_tf32, void* stream);

/* ---- depth branch (use_depth=True; models/naive.py:233-240,324-330, models/time_sensitive.py:387-394,481-487):
 *      depth (B,1,H,W) -> AvgPool2d(2) x log2(pool) -> InstanceNorm2d(1, affine, biased variance, eps) -> Flatten,
 *      multiplied element-wise into the aux features aux[b][0:F] (row stride ld) in place.  xhat [B][F] and
 *      aux_pre [B][F] (the aux features before the product) are kept for the backward pass.              */
int pe_depth_features_fwd(const float* depth, int B, int H, int W, int pool, const float* gamma, const float* beta,
                          float eps, float* xhat, float* aux, int ld, float* aux_pre, int round_tf32, void* stream);
/* dprod (gradient w.r.t. the product, row stride ld) becomes the gradient w.r.t. the aux features in place;
 * dgamma / dbeta (1 element each, zeroed by the caller, may be NULL) accumulate the InstanceNorm gradients */
int pe_depth_features_bwd(float* dprod, int ld, const float* aux_pre, const float* xhat, const float* gamma,
                          const float* beta, float* dgamma, float* dbeta, int B, int F, void* stream);

/* ---- LSTM cell (nn.LSTM single layer, gate order i,f,g,o; models/time_sensitive.py:126-131,418) -----
 * gates = gx + gh + b_ih + b_hh  ([N][4H]); act = (sig i, sig f, tanh g, sig o); c = f*c_prev + i*g;
 * h = o*tanh(c).  `act` keeps the activated gates for backward.                                       */
int pe_lstm_cell_fwd(const float* gx, int ldgx, const float* gh, int ldgh, const float* b_ih, const float* b_hh,
                     const float* c_prev, float* c_out, float* h_out, int ldh, float* act, int N, int Hd,
                     int round_tf32, void* stream);
/* dgates (pre-activation) and dc_prev from dh (= dh_out + dh_next) and dc_next */
int pe_lstm_cell_bwd(const float* dh, int lddh, const float* dh_rec, const float* dc_next, const float* act,
                     const float* c_prev, const float* c_out, float* dgates, int lddg, float* dc_prev, int N,
                     int Hd, void* stream);

/* ---- persistent LSTM recurrence: ONE launch for all S timesteps (models/time_sensitive.py:501-510 nn.LSTM call;
 *      gates i,f,g,o).  gx = x W_ih^T for all S*N rows (pe_linear_fwd, no bias); W_hh stays in the checkpoint layout
 *      [4H, H] (fp32, un-rounded) and is sliced across the grid, resident in shared memory for the whole sequence; one
 *      grid-wide barrier per timestep (cooperative launch), LSTM cell fused behind the recurrent dot products.
 *      h0 / c0: carried state [N, H] or NULL (zeros).  act: activated gates [S*N, 4H] for the backward pass or NULL.
 *      Backward (h0 = NULL sequences): dgates [S*N, 4H] (TF32-rounded when round_tf32) from dh_all [S*N, H]; the weight
 *      and input gradients then follow as plain GEMMs on dgates. ------------------------------------------------- */
int pe_lstm_seq_supported(int N, int Hd, int backward);      /* 1 if the shape fits (grid <= SMs, smem <= 200 KB) */
int pe_lstm_seq_fwd(const float* gx, const float* w_hh, const float* b_ih, const float* b_hh, const float* h0,
                    const float* c0, float* h_all, float* c_all, float* act, int S, int N, int Hd, int round_tf32,
                    void* stream);
int pe_lstm_seq_bwd(const float* dh_all, const float* w_hh, const float* act, const float* c_all, const float* c0,
                    float* dgates, int S, int N, int Hd, int round_tf32, void* stream);

/* ---- fused fusion-head step for rollout inference (<= 8 frames): replaces, in ONE launch, the torch.cat of
 *      image / aux features with the proprioceptive vector, the first dense layer or LSTM gate projection, the
 *      LSTM cell and the remaining small layers (models/naive.py:333-345, :92-104; models/time_sensitive.py:
 *      213-246, :491-513).  Phase A runs on every SM (each warp streams rows of the first weight matrix once);
 *      the last CTA to finish runs the tail.  Weights are read in their checkpoint (nn.Linear / nn.LSTM) layout. */
typedef struct pe_head_desc {
    /* phase A: out_a[n][j] = act(w_x[j][:] . x[n][:] + w_h[j][:] . h_prev[n][:] + b1[j] + b2[j]),  j < j_a */
    const float* x;          /* [n_rows][ldx] fusion rows (latent | aux | 7 slots), first k_x columns are used */
    int ldx, n_rows;
    const float* inj;        /* optional [n_rows][ld_inj]: 7 values injected at columns inj_col.. (proprio vector) */
    int ld_inj, inj_col;
    int k_x;
    const float* w_x;        /* [j_a][k_x] */
    int k_h;                 /* 0 = no recurrent term */
    const float* w_h;        /* [j_a][k_h] */
    const float* h_prev;     /* [n_rows][k_h] or NULL (zero state) */
    const float* b1;
    const float* b2;         /* biases, either may be NULL */
    int j_a, relu_a;
    float* out_a;            /* [n_rows][j_a] scratch (gates / first-layer activations) */
    unsigned int* counter;   /* zero-initialised ticket, reset by the kernel */
    /* tail */
    int lstm_hidden;         /* > 0: out_a holds LSTM gates (i,f,g,o) and the cell runs first */
    const float* c_prev;     /* [n_rows][lstm_hidden] or NULL */
    float* c_out;
    float* h_out;            /* may alias c_prev / h_prev (in-place state update) */
    int n_tail;              /* 0..3 further dense layers */
    const float* tail_w[3];  /* [tail_j[t]][previous width] */
    const float* tail_b[3];
    int tail_j[3], tail_relu[3];
    float* out;              /* [n_rows][ld_out] output of the last layer */
    int ld_out;
    const float* meas;       /* optional: diff[n][diff_col + j] = out[n][j] - meas[n][j] (pre_out - x0bar) */
    int ld_meas;
    float* diff;
    int ld_diff, diff_col;
} pe_head_desc;
/* `desc` is a HOST pointer (copied into the launch); every pointer inside it is a device pointer */
int pe_fused_head(const pe_head_desc* desc, void* stream);
int pe_head_desc_size(void);   /* sizeof(pe_head_desc), for language bindings that mirror the struct */

/* ---- pose loss (models/losses.py:47-128) ------------------------------------------------------------
 * metric: 0 l1, 1 l2, 2 linf, 3 combined.  mode: 0 position, 1 pose.  loss[0] = scale*(pos+alpha*ori)
 * summed over the n rows; dpred = d loss / d pred (pass NULL to skip).  val-mode metrics:
 * val[0] = sum l2 position error, val[1] = sum |angle| (rad) between normalised quaternions.          */
int pe_pose_loss(const float* pred, int ldp, const float* truth, int ldt, long long n, int metric, int mode,
                 float alpha, float epsilon, float scale, float* loss, float* dpred, int lddp, float* val,
                 void* stream);

/* ---- optimizers (torch.optim.Adam at scripts/train_model.py:228; SGD for completeness) --------------
 * flat arrays; grad_scale multiplies the gradient first (1.0 normally); step is the 1-based step count. */
int pe_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                 float eps, float weight_decay, int step, float grad_scale, void* stream);
int pe_sgd_step(float* p, const float* g, float* mom, long long n, float lr, float momentum, float weight_decay,
                int first_step, float grad_scale, void* stream);

/* ---- misc -------------------------------------------------------------------------------------------*/
int pe_fill(float* p, long long n, float value, void* stream);
/* p[i] += value over an int64 array (BatchNorm num_batches_tracked counters) */
int pe_add_i64(long long* p, long long n, long long value, void* stream);
/* uint8 HWC frame batch -> normalised fp32 NCHW (Resize is the caller's; crop + /255 + mean/std here):
 * util/data_utils.py:48-54, util/learn_utils.py:299-305 */
/* mean3 / std3 are HOST arrays of 3 floats */
int pe_preprocess_u8(const unsigned char* src, float* dst, int B, int Hs, int Ws, int crop, const float* mean3,
                     const float* std3, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PE_B200_H_ */

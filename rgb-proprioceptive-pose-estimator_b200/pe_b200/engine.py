"""Host-side sequencing of the sm_100a kernels for the pose-estimator hot path.

`TrunkEngine` runs the ResNet-50 trunk + auxiliary BN1 branch (forward in train / eval mode, and the
full backward) on NHWC fp32 activations; `MLPHead`, `LSTMLayer` and friends run the fusion heads.
Parameters stay in the reference's own module tree / checkpoint layout (OIHW conv weights etc.);
the engine keeps TF32-rounded packed shadows that are refreshed whenever a parameter changes.

Everything here only *sequences* C-ABI calls (pe_b200.native); there is no torch compute on the
path and no fallback when the library is missing.
"""
import torch

from . import native

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
# training forward: bn1 + ReLU + max pool + aux branch as one kernel (pe_stem_post_train); False = the three separate
# kernels (kept for the kernel tests and as the reference the fused one is checked against)
FUSED_STEM_TAIL = [True]
# BatchNorm backward pass 1 (sum g, sum g * xhat) of bn1 / bn2 of every Bottleneck inside the epilogue of the dgrad that
# produces their output gradient (conv2 / conv3 dgrad); False = the stand-alone pe_bn_bwd_reduce launch
FUSE_BN_REDUCE = [True]
# 7x7/2 stem as a 4x4/1 convolution over the 2x2 space-to-depth image (pe_stem_conv_fwd / pe_stem_conv_wgrad: the
# tap-GEMM reads a TMA view with overlapping pixel rows, no im2col matrix in HBM); False = im2col + GEMM (the path for
# image sides that are not multiples of 16, kept under test)
STEM_S2D = [True]
# test hook: when set to a list, every training-mode convolution appends its raw output (an Act, execution order:
# stem, then conv1 / conv2 / conv3 / downsample of each block) -- read by the parity tests (tests/model_checks.py)
CAPTURE_CONV_OUTPUTS = [None]


def _dev_check(t):
    if not t.is_cuda:
        raise native.PeError("the B200 pose-estimator path needs CUDA tensors (got %s); there is no CPU fallback"
                             % t.device)


class Act:
    """An NHWC activation: tensor [B*H*W, C] (contiguous) plus its geometry.  `bn` = (BatchNorm index, its input
    Act) when this activation is relu(bn(y)) without a residual input: the dgrad of its single consumer can then
    accumulate that BatchNorm's backward sums in its epilogue (pe_conv2d_dgrad_bn)."""
    __slots__ = ("t", "B", "H", "W", "C", "bn")

    def __init__(self, t, B, H, W, C):
        self.t, self.B, self.H, self.W, self.C = t, B, H, W, C
        self.bn = None

    @property
    def P(self):
        return self.B * self.H * self.W


class GradSlots:
    """Pending gradient contributions per activation: (tensor, maskbits) pairs.  `maskbits` marks a lazy
    ReLU-masked view (the identity branch of a residual join hands its consumer the un-masked join gradient
    plus the join's bit mask instead of materialising the masked copy)."""

    def __init__(self):
        self.d = {}

    def add(self, act, g, maskbits=None):
        self.d.setdefault(id(act), []).append((g, maskbits))

    def pending(self, act):
        return len(self.d.get(id(act), ()))

    def pop(self, act):
        gs = self.d.pop(id(act), [])
        if len(gs) > 2:
            raise native.PeError("internal: more than two gradient branches for one activation")
        return gs

    def pop_plain(self, act, n_max=2):
        """Entries without masks, padded with None to n_max tensors."""
        gs = self.pop(act)
        if len(gs) > n_max or any(m is not None for _, m in gs):
            raise native.PeError("internal: unexpected gradient branches (%d entries) for this consumer" % len(gs))
        return ([g for g, _ in gs] + [None] * n_max)[:n_max]


def _params_version(params):
    return tuple((p.data_ptr(), p._version) for p in params)


class TrunkEngine:
    """ResNet-50 trunk (torchvision layout: conv1, bn1, layer1..4, fc) + optional aux 1x1 conv branch."""

    def __init__(self, net, aux_conv=None, aux_trainable=True):
        self.net = net
        self.aux_conv = aux_conv
        self.aux_trainable = aux_trainable
        # execution-ordered (conv, bn) pairs
        self.convs = [(net.conv1, net.bn1)]
        self.blocks = []
        for layer in (net.layer1, net.layer2, net.layer3, net.layer4):
            for blk in layer:
                ids = {}
                for k in (1, 2, 3):          # Bottleneck: conv1..conv3; BasicBlock (ResNet-18): conv1, conv2
                    if not hasattr(blk, "conv%d" % k):
                        break
                    ids[k] = len(self.convs)
                    self.convs.append((getattr(blk, "conv%d" % k), getattr(blk, "bn%d" % k)))
                ids["n"] = max(k for k in ids if isinstance(k, int))
                if blk.downsample is not None:
                    ids["d"] = len(self.convs)
                    self.convs.append((blk.downsample[0], blk.downsample[1]))
                self.blocks.append((blk, ids))
        self.bn_off = []
        off = 0
        for _, bn in self.convs:
            self.bn_off.append(off)
            off += bn.num_features
        self.bn_total = off
        self._dev = None
        self._packed_version = None
        self._eval_version = None
        self._pack_key = None
        self.round_tf32 = 1
        # The BN coefficient arenas (scale / shift / mean / invstd) are overwritten by every forward.  Under autograd
        # torch allows several forwards before a backward (loss(model(a)) + loss(model(b))), so a taped forward keeps
        # its own copy of the four vectors (26 560 floats each); the fused trainer runs strictly forward -> backward
        # and switches the copy off.
        self.snapshot_bn = True

    # ------------------------------------------------------------------------------------------
    def _ensure_device(self, dev):
        if self._dev == dev:
            return
        self._dev = dev
        f32 = dict(device=dev, dtype=torch.float32)
        self.scale = torch.empty(self.bn_total, **f32)
        self.shift = torch.empty(self.bn_total, **f32)
        self.mean = torch.empty(self.bn_total, **f32)
        self.invstd = torch.empty(self.bn_total, **f32)
        self.stats = torch.zeros(2 * self.bn_total, device=dev, dtype=torch.float64)
        self.sums = torch.zeros(2 * self.bn_total, device=dev, dtype=torch.float64)
        self.w_tck, self.w_tkc = [], []
        for i, (conv, _) in enumerate(self.convs):
            co, ci, r, s = conv.weight.shape
            if i == 0:
                self.w_tck.append(torch.zeros(co, 160, **f32))   # stem: [64][160] im2col-ordered, zero padded
                self.w_tkc.append(None)
                self.w_stem_s2d = torch.zeros(4, co, 64, **f32)  # stem, space-to-depth form: [filter row][Cout][64]
            else:
                self.w_tck.append(torch.empty(r * s, co, ci, **f32))
                self.w_tkc.append(torch.empty(r * s, ci, co, **f32))
        fc = self.net.fc
        self.fc_w = torch.empty(fc.out_features, fc.in_features, **f32)
        # transposed shadow for the fc dgrad: rows padded to a multiple of 4 floats (any latent_dim, e.g. the
        # constructors' default 50, keeps 16-byte aligned rows for TMA); the padding stays zero
        self.fc_wt = torch.zeros(fc.in_features, (fc.out_features + 3) // 4 * 4, **f32)
        self._packed_version = None
        self._eval_version = None
        self._pack_key = None

    def _weights(self):
        return [c.weight for c, _ in self.convs] + [self.net.fc.weight]

    def invalidate(self):
        """Call after parameters were modified behind torch's back (e.g. by the fused optimizer kernel)."""
        self._packed_version = None
        self._eval_version = None

    def _stem_s2d_ok(self, H=224, W=224):
        c1 = self.net.conv1
        co, ci, r, s = c1.weight.shape
        return (STEM_S2D[0] and (ci, r, s) == (3, 7, 7) and c1.stride[0] == 2 and c1.padding[0] == 3 and co % 32 == 0
                and H % 16 == 0 and W % 16 == 0)

    def pack_weights(self, need_dgrad, force=False):
        """Refresh the TF32-rounded packed shadows if any weight changed since the last call."""
        ver = (_params_version(self._weights()), need_dgrad)
        if ver == self._packed_version and not force:
            return
        L, st, P = native.lib(), native.stream_ptr(), native.ptr
        w0 = self.convs[0][0].weight
        _dev_check(w0)
        co, ci, r, s = w0.shape
        if self._stem_s2d_ok():
            L.pe_stem_pack_weight(P(w0), P(self.w_stem_s2d), co, self.round_tf32, st)
        self._stem_col_packed = False      # the im2col-ordered shadow is refreshed on demand (fallback path only)
        # all bottleneck convs in one launch; the pointer table is rebuilt only when a parameter moved
        key = (tuple(c.weight.data_ptr() for c, _ in self.convs[1:]), need_dgrad)
        if self._pack_key != key:
            per = L.pe_pack_block_elems()
            rows, blk = [], 0
            for i, (conv, _) in enumerate(self.convs):
                if i == 0:
                    continue
                w = conv.weight
                _dev_check(w)
                co, ci, r, s = w.shape
                n = co * ci * r * s
                rows.append([P(w), P(self.w_tck[i]), P(self.w_tkc[i]) if need_dgrad else 0, co, ci, r * s, blk, n])
                blk += (n + per - 1) // per
            self._pack_table = torch.tensor(rows, dtype=torch.int64).to(w0.device)
            self._pack_blocks = blk
            self._pack_key = key
        L.pe_pack_conv_weights_batched(P(self._pack_table), self._pack_table.shape[0], self._pack_blocks,
                                       self.round_tf32, st)
        fc = self.net.fc
        L.pe_copy_cols(P(fc.weight), fc.in_features, P(self.fc_w), fc.in_features, fc.out_features, fc.in_features,
                       self.round_tf32, st)
        if need_dgrad:
            L.pe_transpose(P(fc.weight), fc.in_features, P(self.fc_wt), self.fc_wt.shape[1], fc.out_features,
                           fc.in_features, self.round_tf32, st)
        self._packed_version = ver

    def _bn_views(self, i, arena=None):
        C = self.convs[i][1].num_features
        o = self.bn_off[i]
        sc, sh, mean, invstd = arena if arena is not None else (self.scale, self.shift, self.mean, self.invstd)
        return (sc[o:o + C], sh[o:o + C], mean[o:o + C], invstd[o:o + C],
                self.stats[2 * o:2 * o + 2 * C], self.sums[2 * o:2 * o + 2 * C])

    def prepare_eval(self):
        """Fold running statistics into per-channel scale / shift (cached until a BN tensor changes)."""
        tensors = []
        for _, bn in self.convs:
            tensors += [bn.weight, bn.bias, bn.running_mean, bn.running_var]
        ver = _params_version(tensors)
        if ver == self._eval_version:
            return
        L, st, P = native.lib(), native.stream_ptr(), native.ptr
        for i, (_, bn) in enumerate(self.convs):
            sc, sh, _, _, _, _ = self._bn_views(i)
            L.pe_bn_finalize(None, P(bn.weight), P(bn.bias), P(bn.running_mean), P(bn.running_var), P(sc), P(sh),
                             None, None, 1, BN_MOMENTUM, bn.eps, bn.num_features, st)
        self._eval_version = ver

    # ------------------------------------------------------------------------------------------
    def _conv_train(self, x, i, tape):
        L, st, P = native.lib(), native.stream_ptr(), native.ptr
        conv, bn = self.convs[i]
        co, ci, r, s = conv.weight.shape
        stride, pad = conv.stride[0], conv.padding[0]
        Ho = (x.H + 2 * pad - r) // stride + 1
        Wo = (x.W + 2 * pad - s) // stride + 1
        y = Act(torch.empty(x.B * Ho * Wo, co, device=x.t.device, dtype=torch.float32), x.B, Ho, Wo, co)
        stats = self._bn_views(i)[4]
        L.pe_conv2d_fwd(P(x.t), P(self.w_tck[i]), P(y.t), x.B, x.H, x.W, ci, co, r, s, stride, pad, None, None, None,
                        0, 0, P(stats), st)
        native.account("pe_conv2d_fwd", 4 * (x.P * ci + y.P * co + conv.weight.numel()))
        if CAPTURE_CONV_OUTPUTS[0] is not None:
            CAPTURE_CONV_OUTPUTS[0].append(y)
        if tape is not None:
            tape.append(("conv", i, x, y))
        return y

    def _bn_train(self, y, i, relu, residual, tape, update_running=True):
        """Training-mode BatchNorm (+residual, +ReLU) in one pass: statistics finalisation, running-stat update,
        num_batches_tracked and the normalisation share a kernel; residual joins also emit their ReLU mask as
        bits for the backward pass."""
        L, st, P = native.lib(), native.stream_ptr(), native.ptr
        _, bn = self.convs[i]
        sc, sh, mean, invstd, stats, _ = self._bn_views(i)
        out = Act(torch.empty_like(y.t), y.B, y.H, y.W, y.C)
        maskbits = None
        if residual is not None and relu and tape is not None:
            n4 = y.P * (y.C // 4)
            maskbits = torch.empty((n4 + 31) // 32 * 4, device=y.t.device, dtype=torch.int32)
        L.pe_bn_train_apply(P(y.t), P(stats), P(bn.weight), P(bn.bias),
                            P(bn.running_mean) if update_running else None,
                            P(bn.running_var) if update_running else None,
                            P(bn.num_batches_tracked) if update_running else None, P(sc), P(sh), P(mean), P(invstd),
                            P(residual.t) if residual is not None else None, P(out.t), P(maskbits), y.P, y.C,
                            bn.momentum if bn.momentum is not None else BN_MOMENTUM, bn.eps, int(relu),
                            self.round_tf32, st)
        native.account("pe_bn_train_apply", 4 * y.t.numel() * (2 + (residual is not None)) +
                       (4 * maskbits.numel() if maskbits is not None else 0))
        if tape is not None:
            if relu and residual is None:
                out.bn = (i, y)
            tape.append(("bn", i, y, out, relu, residual, maskbits))
        return out

    def _conv_eval(self, x, i, relu, residual):
        L, st, P = native.lib(), native.stream_ptr(), native.ptr
        conv, bn = self.convs[i]
        co, ci, r, s = conv.weight.shape
        stride, pad = conv.stride[0], conv.padding[0]
        Ho = (x.H + 2 * pad - r) // stride + 1
        Wo = (x.W + 2 * pad - s) // stride + 1
        y = Act(torch.empty(x.B * Ho * Wo, co, device=x.t.device, dtype=torch.float32), x.B, Ho, Wo, co)
        sc, sh = self._bn_views(i)[:2]
        L.pe_conv2d_fwd(P(x.t), P(self.w_tck[i]), P(y.t), x.B, x.H, x.W, ci, co, r, s, stride, pad, P(sc), P(sh),
                        P(residual.t) if residual is not None else None, int(relu), self.round_tf32, None, st)
        return y

    # ------------------------------------------------------------------------------------------
    def forward(self, img, training, need_grad, feat_out, ld_feat, aux_out=None, ld_aux=0, aux_round=True):
        """img: (B,3,224,224) NCHW fp32 CUDA.  Writes latent features into feat_out[:, :latent] (row stride
        ld_feat) and, when the aux branch exists, its 3136-vector into aux_out (row stride ld_aux).
        Returns a context for backward (None unless need_grad)."""
        _dev_check(img)
        if img.dtype != torch.float32:
            raise native.PeError("image tensor must be float32")
        img = img.contiguous()
        L, st, P = native.lib(), native.stream_ptr(), native.ptr
        self._ensure_device(img.device)
        self.pack_weights(need_grad, force=training)
        if training:
            self._eval_version = None   # scale/shift arenas are about to hold batch statistics
        B, Cimg, H, W = img.shape
        conv1 = self.net.conv1
        r, s = conv1.weight.shape[2:]
        stride, pad = conv1.stride[0], conv1.padding[0]
        Ho, Wo = (H + 2 * pad - r) // stride + 1, (W + 2 * pad - s) // stride + 1
        tape = [] if need_grad else None
        # feature-extraction mode (util/model_utils.py:110-113,137: every trunk parameter frozen, only the new fc
        # trains): nothing upstream of the average pool needs a gradient, so the convolutions are not taped and
        # backward reduces to the fc / aux-conv parameter gradients
        frozen = need_grad and not any(p.requires_grad for c, b in self.convs for p in (c.weight, b.weight, b.bias))
        head_tape = tape
        if frozen:
            tape = None
        dev = img.device

        # ---- stem: conv1 (+BN1, ReLU), 3x3/2 max pool, aux branch ---------------------------------
        # conv1's operand: the zero-bordered space-to-depth image (162 MB at 256 frames) read through an overlapping TMA
        # view, or -- for image sides that are not multiples of 16 -- the im2col matrix (2 GB) + a plain GEMM
        s2d = self._stem_s2d_ok(H, W)
        if s2d:
            col = torch.empty(B, H // 2 + 3, W // 2 + 3, 12, device=dev, dtype=torch.float32)
            L.pe_stem_s2d_pack(P(img), P(col), B, H, W, self.round_tf32, st)

            def stem_gemm(scale, shift, relu, rnd, stats):
                L.pe_stem_conv_fwd(P(col), P(self.w_stem_s2d), P(y0.t), B, H, W, 64, P(scale), P(shift), relu, rnd,
                                   P(stats), st)
        else:
            if not self._stem_col_packed:
                w0 = conv1.weight
                L.pe_copy_cols(P(w0), w0[0].numel(), P(self.w_tck[0]), 160, w0.shape[0], w0[0].numel(), self.round_tf32,
                               st)
                self._stem_col_packed = True
            col = torch.empty(B * Ho * Wo, 160, device=dev, dtype=torch.float32)
            L.pe_im2col_stem(P(img), P(col), B, Cimg, H, W, r, s, stride, pad, 160, self.round_tf32, st)

            def stem_gemm(scale, shift, relu, rnd, stats):
                L.pe_linear_fwd(P(col), 160, P(self.w_tck[0]), 160, P(shift), P(scale), P(y0.t), 64, y0.P, 64, 160, relu,
                                0, rnd, P(stats), st)
        y0 = Act(torch.empty(B * Ho * Wo, 64, device=dev, dtype=torch.float32), B, Ho, Wo, 64)
        if training:
            self.stats.zero_()
            stem_gemm(None, None, 0, 0, self._bn_views(0)[4])
            native.account("pe_stem_conv_fwd" if s2d else "pe_linear_fwd", 4 * (col.numel() + y0.t.numel()))
            if CAPTURE_CONV_OUTPUTS[0] is not None:
                CAPTURE_CONV_OUTPUTS[0].append(y0)
            if tape is not None:
                tape.append(("stem", col, y0))
            fused_tail = FUSED_STEM_TAIL[0] and Ho % 2 == 0 and Wo % 2 == 0
            a1 = None if fused_tail else self._bn_train(y0, 0, True, None, tape)
        else:
            self.prepare_eval()
            sc, sh = self._bn_views(0)[:2]
            stem_gemm(sc, sh, 1, self.round_tf32, None)
            a1 = y0
            del col
        Hp, Wp = (Ho + 2 - 3) // 2 + 1, (Wo + 2 - 3) // 2 + 1
        x = Act(torch.empty(B * Hp * Wp, 64, device=dev, dtype=torch.float32), B, Hp, Wp, 64)
        argmax = torch.empty(B * Hp * Wp * 64, device=dev, dtype=torch.uint8) if tape is not None else None
        aux_am = None
        if self.aux_conv is not None and need_grad:
            aux_am = torch.empty(B * (Ho // 2) * (Wo // 2), device=dev, dtype=torch.uint8)
        # aux_round=False: the depth branch multiplies these features next and rounds the product itself
        aux_rt = self.round_tf32 if aux_round else 0
        if a1 is None:
            # training: bn1 + ReLU + max pool + aux branch in one pass over y0; the normalised activation is never
            # written (`a1` is a geometry-only handle the gradient slots are keyed by)
            _, bn = self.convs[0]
            sc, sh, mean, invstd, stats, _ = self._bn_views(0)
            ac = self.aux_conv
            L.pe_stem_post_train(P(y0.t), P(stats), P(bn.weight), P(bn.bias), P(bn.running_mean), P(bn.running_var),
                                 P(bn.num_batches_tracked), P(sc), P(sh), P(mean), P(invstd), P(x.t), P(argmax),
                                 P(ac.weight) if ac is not None else None, P(ac.bias) if ac is not None else None,
                                 P(aux_out) if ac is not None else None, ld_aux, P(aux_am), B, Ho, Wo, 64,
                                 bn.momentum if bn.momentum is not None else BN_MOMENTUM, bn.eps, self.round_tf32,
                                 aux_rt, st)
            a1 = Act(None, B, Ho, Wo, 64)
            if tape is not None:
                tape.append(("bn", 0, y0, a1, True, None, None))
                tape.append(("maxpool", a1, x, argmax))
            if self.aux_conv is not None and head_tape is not None:
                head_tape.append(("aux", a1, aux_am, y0))
        else:
            L.pe_maxpool3x3s2_fwd(P(a1.t), P(x.t), P(argmax), B, Ho, Wo, 64, st)
            if tape is not None:
                tape.append(("maxpool", a1, x, argmax))
            if self.aux_conv is not None:
                L.pe_aux_fwd(P(a1.t), P(self.aux_conv.weight), P(self.aux_conv.bias), P(aux_out), ld_aux, P(aux_am),
                             B, Ho, Wo, 64, aux_rt, st)
                if head_tape is not None:
                    head_tape.append(("aux", a1, aux_am, None))

        # ---- bottleneck stages ------------------------------------------------------------------
        for blk, ids in self.blocks:
            last = ids["n"]                   # the conv whose BatchNorm output joins the identity branch
            o = x
            if training:
                for k in range(1, last):
                    o = self._bn_train(self._conv_train(o, ids[k], tape), ids[k], True, None, tape)
                y_last = self._conv_train(o, ids[last], tape)
                if "d" in ids:
                    idn = self._bn_train(self._conv_train(x, ids["d"], tape), ids["d"], False, None, tape)
                else:
                    idn = x
                x = self._bn_train(y_last, ids[last], True, idn, tape)
            else:
                for k in range(1, last):
                    o = self._conv_eval(o, ids[k], True, None)
                idn = self._conv_eval(x, ids["d"], False, None) if "d" in ids else x
                x = self._conv_eval(o, ids[last], True, idn)

        # ---- global average pool + fc -----------------------------------------------------------
        fc = self.net.fc
        pool = torch.empty(B, x.C, device=dev, dtype=torch.float32)
        L.pe_avgpool_fwd(P(x.t), P(pool), x.C, B, x.H * x.W, x.C, self.round_tf32, st)
        L.pe_linear_fwd(P(pool), x.C, P(self.fc_w), x.C, P(fc.bias), None, P(feat_out), ld_feat, B, fc.out_features,
                        x.C, 0, 0, self.round_tf32, None, st)
        if head_tape is None:
            return None
        head_tape.append(("tail", x, pool))
        arena = None
        if self.snapshot_bn and not frozen:
            arena = (self.scale.clone(), self.shift.clone(), self.mean.clone(), self.invstd.clone())
        return {"tape": head_tape, "B": B, "frozen": frozen, "arena": arena}

    # ------------------------------------------------------------------------------------------
    def backward(self, ctx, d_feat, ld_dfeat, d_aux, ld_daux, grad_of, on_ready=None):
        """d_feat: gradient w.r.t. the latent features (row stride ld_dfeat); d_aux likewise for the aux
        vector.  `grad_of(param)` returns the tensor that receives that parameter's gradient;
        `on_ready(params)` is told whenever a group of parameters has its final gradient (all-reduce
        buckets can start while the rest of backward is still running)."""
        L, st, P = native.lib(), native.stream_ptr(), native.ptr
        tape = ctx["tape"]
        frozen = ctx.get("frozen", False)
        arena = ctx.get("arena")
        done_after_conv = {}
        if on_ready is not None:
            for blk, ids in self.blocks:
                done_after_conv[ids[1]] = [p for p in blk.parameters()]
        B = ctx["B"]
        slots = GradSlots()
        reduced = {}          # BatchNorm index -> the dx tensor whose dgrad already accumulated its backward sums
        pending_aux = None
        dev = self.scale.device
        self.sums.zero_()
        rt = self.round_tf32
        for rec in reversed(tape):
            kind = rec[0]
            if kind == "tail":
                _, x, pool = rec
                fc = self.net.fc
                nin, nout = fc.in_features, fc.out_features
                # rounded copy of d_feat so that the TF32 truncation of the MMA operand is unbiased
                # (row stride padded to a multiple of 4 floats: latent_dim need not be one -- the default is 50)
                ldf = (nout + 3) // 4 * 4
                dfr = torch.zeros(B, ldf, device=dev, dtype=torch.float32) if ldf != nout else \
                    torch.empty(B, ldf, device=dev, dtype=torch.float32)
                L.pe_copy_cols(P(d_feat), ld_dfeat, P(dfr), ldf, B, nout, rt, st)
                L.pe_linear_wgrad(P(pool), nin, P(dfr), ldf, P(grad_of(fc.weight)), nin, B, nout, nin, st)
                L.pe_colsum(P(dfr), ldf, P(grad_of(fc.bias)), B, nout, 0, st)
                if frozen:
                    if on_ready is not None:
                        on_ready([fc.weight, fc.bias])
                    continue
                dpool = torch.empty(B, nin, device=dev, dtype=torch.float32)
                L.pe_linear_fwd(P(dfr), ldf, P(self.fc_wt), ldf, None, None, P(dpool), nin, B, nin, nout, 0, 0, 0,
                                None, st)
                dx = torch.empty_like(x.t)
                L.pe_avgpool_bwd(P(dpool), nin, P(dx), B, x.H * x.W, x.C, st)
                slots.add(x, dx)
                if on_ready is not None:
                    on_ready([fc.weight, fc.bias])
            elif kind == "bn":
                _, i, y, out, relu, residual, maskbits = rec
                _, bn = self.convs[i]
                sc, sh, mean, invstd, _, sums = self._bn_views(i, arena)
                entries = slots.pop(out)
                dy = torch.empty_like(y.t)
                if residual is not None and maskbits is not None:
                    # residual join: g = D * mask(out > 0) with the mask read as bits; the identity branch gets
                    # (D, bits) instead of a materialised masked copy
                    if len(entries) == 2 and entries[0][1] is None and entries[1][1] is None:
                        # two plain branches that no dgrad epilogue could merge (stride-2 3x3 conv1 + downsample of a
                        # BasicBlock stage transition): one elementwise add
                        g0, g1 = entries[0][0], entries[1][0]
                        L.pe_axpby_cols(P(g0), y.C, P(g1), y.C, P(g0), y.C, y.P, y.C, 1.0, 1.0, 0, st)
                        entries = [(g0, None)]
                    if len(entries) != 1 or entries[0][1] is not None:
                        raise native.PeError("internal: a residual join expects one complete gradient")
                    d1, d2, mb, mask_src, relu_k = entries[0][0], None, maskbits, None, 0
                else:
                    if len(entries) == 1 and entries[0][1] is not None:
                        d1, d2, mb = entries[0][0], None, entries[0][1]     # gradient of a join, masked lazily
                    else:
                        if any(m is not None for _, m in entries):
                            raise native.PeError("internal: masked gradient next to a second branch")
                        d1, d2 = ([g for g, _ in entries] + [None, None])[:2]
                        mb = None
                    # without a residual input the ReLU mask is recomputed from y (one activation read less)
                    mask_src = out.t if (relu and residual is not None) else None
                    relu_k = int(relu)
                fused_here = i in reduced
                if fused_here:
                    # pass 1 already happened in the epilogue of the dgrad that produced d1
                    if d1 is not reduced.pop(i) or d2 is not None or mb is not None:
                        raise native.PeError("internal: fused BatchNorm sums do not match the gradient at hand")
                else:
                    L.pe_bn_bwd_reduce(P(d1), P(d2), P(mask_src), P(y.t), P(mean), P(invstd), P(sc), P(sh), P(mb),
                                       P(sums), y.P, y.C, relu_k, st)
                dres = None
                if residual is not None and maskbits is None:
                    dres = torch.empty_like(y.t)
                L.pe_bn_bwd_apply(P(d1), P(d2), P(mask_src), P(y.t), P(mean), P(invstd), P(bn.weight), P(sc), P(sh),
                                  P(mb), P(sums), P(dy), P(dres), 0, P(grad_of(bn.weight)), P(grad_of(bn.bias)), 0,
                                  y.P, y.C, relu_k, rt, st)
                E = 4 * y.t.numel()
                reads = E * (2 + (d2 is not None) + (mask_src is not None)) + (4 * mb.numel() if mb is not None else 0)
                if not fused_here:
                    native.account("pe_bn_bwd_reduce", reads)
                native.account("pe_bn_bwd_apply", reads + E * (1 + (dres is not None)))
                slots.add(y, dy)
                if residual is not None:
                    if maskbits is not None:
                        slots.add(residual, d1, maskbits)
                    else:
                        slots.add(residual, dres)
            elif kind == "conv":
                _, i, x, y = rec
                conv, _ = self.convs[i]
                co, ci, r, s = conv.weight.shape
                stride, pad = conv.stride[0], conv.padding[0]
                dy, = slots.pop_plain(y, 1)
                gw = grad_of(conv.weight)
                if r == 1 and s == 1:
                    L.pe_conv2d_wgrad(P(x.t), P(dy), P(gw), x.B, x.H, x.W, ci, co, r, s, stride, pad, st)
                else:
                    tmp = torch.empty(r * s, co, ci, device=dev, dtype=torch.float32)
                    L.pe_conv2d_wgrad(P(x.t), P(dy), P(tmp), x.B, x.H, x.W, ci, co, r, s, stride, pad, st)
                    L.pe_unpack_conv_wgrad(P(tmp), P(gw), co, ci, r, s, 0, st)
                dx = torch.empty_like(x.t)
                res = res_mask = None
                if stride == 1 and slots.pending(x) == 1 and ci % 32 == 0:
                    # the other gradient branch of x (identity / downsample path) is added in the dgrad epilogue
                    (res, res_mask), = slots.pop(x)
                bn_src = x.bn if (FUSE_BN_REDUCE[0] and res is None and ci % 32 == 0 and
                                  slots.pending(x) == 0) else None
                if bn_src is not None:
                    # x = relu(bn(y)) feeds this conv only: its BatchNorm's backward sums ride in this epilogue
                    ib, yb = bn_src
                    bsc, bsh, bmean, bistd, _, bsums = self._bn_views(ib, arena)
                    L.pe_conv2d_dgrad_bn(P(dy), P(self.w_tkc[i]), P(dx), x.B, x.H, x.W, ci, co, r, s, stride, pad,
                                         P(yb.t), P(bsc), P(bsh), P(bmean), P(bistd), P(bsums), st)
                    reduced[ib] = dx
                    native.account("pe_conv2d_dgrad", 4 * x.P * ci)          # the extra read of y
                else:
                    L.pe_conv2d_dgrad(P(dy), P(self.w_tkc[i]), P(dx), x.B, x.H, x.W, ci, co, r, s, stride, pad, P(res),
                                      P(res_mask), st)
                native.account("pe_conv2d_wgrad", 4 * (x.P * ci + y.P * co + conv.weight.numel()))
                native.account("pe_conv2d_dgrad", 4 * (x.P * ci * (2 if res is not None else 1) + y.P * co +
                                                       conv.weight.numel()) +
                               (4 * res_mask.numel() if res_mask is not None else 0))
                slots.add(x, dx)
                if i in done_after_conv:
                    on_ready(done_after_conv[i])
            elif kind == "maxpool":
                _, a1, x, argmax = rec
                d1, d2 = slots.pop_plain(x, 2)
                da1 = torch.empty(a1.P, a1.C, device=dev, dtype=torch.float32)
                # the aux branch reads the same activation: its gradient is scattered in the same pass
                ad, ald, aam = pending_aux if pending_aux is not None else (None, 0, None)
                L.pe_maxpool3x3s2_bwd(P(d1), P(d2), P(argmax), P(da1), 0, a1.B, a1.H, a1.W, a1.C, P(ad), ald, P(aam),
                                      P(self.aux_conv.weight) if ad is not None else None, st)
                pending_aux = None
                slots.add(a1, da1)
            elif kind == "aux":
                _, a1, aux_am, y0 = rec
                # taped after "maxpool", hence visited first in the reversed walk
                if self.aux_trainable:
                    gw, gb = grad_of(self.aux_conv.weight), grad_of(self.aux_conv.bias)
                    gw.zero_()
                    gb.zero_()
                    # the 1x1 conv's own gradients (reads a1 at the arg-max pixels only); the scatter into da1 is
                    # fused into the max-pool backward below
                    if y0 is not None:      # fused stem tail: a1 = relu(bn1(y0)) is rebuilt at those pixels
                        sc, sh = self._bn_views(0, arena)[:2]
                        L.pe_aux_bwd_params(P(d_aux), ld_daux, P(aux_am), P(y0.t), P(sc), P(sh), self.round_tf32,
                                            P(gw), P(gb), a1.B, a1.H, a1.W, a1.C, st)
                    else:
                        L.pe_aux_bwd(P(d_aux), ld_daux, P(aux_am), P(a1.t), P(self.aux_conv.weight), None, 0, P(gw),
                                     P(gb), a1.B, a1.H, a1.W, a1.C, st)
                if frozen:
                    if on_ready is not None and self.aux_trainable:
                        on_ready([self.aux_conv.weight, self.aux_conv.bias])
                    continue
                pending_aux = (d_aux, ld_daux, aux_am)
            elif kind == "stem":
                _, col, y0 = rec
                dy, = slots.pop_plain(y0, 1)
                conv1 = self.net.conv1
                co, ci, r, s = conv1.weight.shape
                k = ci * r * s
                if col.dim() == 4:      # space-to-depth operand [B][H/2+3][W/2+3][12]
                    Hi, Wi = 2 * (col.shape[1] - 3), 2 * (col.shape[2] - 3)
                    tmp = torch.empty(4, co, 64, device=dev, dtype=torch.float32)
                    L.pe_stem_conv_wgrad(P(col), P(dy), P(tmp), col.shape[0], Hi, Wi, co, st)
                    L.pe_stem_unpack_wgrad(P(tmp), P(grad_of(conv1.weight)), co, st)
                    native.account("pe_stem_conv_wgrad", 4 * (col.numel() + y0.t.numel()))
                else:
                    tmp = torch.empty(co, 160, device=dev, dtype=torch.float32)
                    L.pe_linear_wgrad(P(col), 160, P(dy), co, P(tmp), 160, y0.P, co, 160, st)
                    L.pe_copy_cols(P(tmp), 160, P(grad_of(conv1.weight)), k, co, k, 0, st)
                if on_ready is not None:
                    ps = [conv1.weight, self.net.bn1.weight, self.net.bn1.bias]
                    if self.aux_conv is not None and self.aux_trainable:
                        ps += [self.aux_conv.weight, self.aux_conv.bias]
                    on_ready(ps)
            else:
                raise native.PeError("internal: unknown tape record %r" % (kind,))

    def parameters_in_backward_order(self):
        """Parameters grouped in the order their gradients become final (for bucketed all-reduce)."""
        groups = [[self.net.fc.weight, self.net.fc.bias]]
        for blk, ids in reversed(self.blocks):
            g = []
            for key in (3, "d", 2, 1):
                if key in ids:
                    conv, bn = self.convs[ids[key]]
                    g += [conv.weight, bn.weight, bn.bias]
            groups.append(g)
        stem = [self.net.conv1.weight, self.net.bn1.weight, self.net.bn1.bias]
        if self.aux_conv is not None and self.aux_trainable:
            stem += [self.aux_conv.weight, self.aux_conv.bias]
        groups.append(stem)
        return groups


class DepthOp:
    """use_depth branch (reference models/naive.py:233-240,324-330): depth -> AvgPool2d(2)^n -> InstanceNorm2d(1,
    affine) -> Flatten, multiplied into the aux features in place (one block per frame, pe_depth_features_*)."""

    def __init__(self, depth_seq, trainable=True):
        import torch.nn as nn
        mods = list(depth_seq)
        self.norm = [m for m in mods if isinstance(m, nn.InstanceNorm2d)][0]
        self.pool = 2 ** sum(isinstance(m, nn.AvgPool2d) for m in mods)
        self.trainable = trainable

    def forward(self, depth, aux_view, ld, M, need_grad):
        L, st, P = native.lib(), native.stream_ptr(), native.ptr
        _dev_check(depth)
        if self.norm.weight.device != depth.device:
            self.norm.to(depth.device)          # unregistered depth nets (td model) are not moved by .cuda()
        depth = depth.reshape(M, *depth.shape[-2:]).to(torch.float32).contiguous()
        H, W = depth.shape[-2:]
        F_ = (H // self.pool) * (W // self.pool)
        xhat = torch.empty(M, F_, device=depth.device, dtype=torch.float32)
        aux_pre = torch.empty(M, F_, device=depth.device, dtype=torch.float32) if need_grad else None
        L.pe_depth_features_fwd(P(depth), M, H, W, self.pool, P(self.norm.weight), P(self.norm.bias), self.norm.eps,
                                P(xhat), P(aux_view), ld, P(aux_pre), 1, st)
        return xhat, aux_pre, F_

    def backward(self, ctx, daux_view, ld, M, grad_of):
        L, st, P = native.lib(), native.stream_ptr(), native.ptr
        xhat, aux_pre, F_ = ctx
        gw = gb = None
        if self.trainable:
            gw, gb = grad_of(self.norm.weight), grad_of(self.norm.bias)
            gw.zero_()
            gb.zero_()
        L.pe_depth_features_bwd(P(daux_view), ld, P(aux_pre), P(xhat), P(self.norm.weight), P(self.norm.bias), P(gw),
                                P(gb), M, F_, st)


# ================================================================================================
# heads
# ================================================================================================
def _pad4(n):
    return (n + 3) // 4 * 4


def _pad32(n):
    return (n + 31) // 32 * 32


class LinearOp:
    """y = act(x W^T + b) through the tap-GEMM; keeps TF32-rounded W (K padded to ld_in) and W^T shadows."""

    def __init__(self, linear, ld_in=None):
        self.lin = linear
        self.nin, self.nout = linear.in_features, linear.out_features
        self.ld_in = ld_in or _pad4(self.nin)
        self.ld_out = _pad4(self.nout)
        self._ver = None
        self.w = self.wt = None

    def pack(self, need_dgrad, round_tf32=1):
        lin = self.lin
        _dev_check(lin.weight)
        ver = (_params_version([lin.weight]), need_dgrad)
        if ver == self._ver:
            return
        L, st, P = native.lib(), native.stream_ptr(), native.ptr
        dev = lin.weight.device
        if self.w is None or self.w.device != dev:
            self.w = torch.zeros(self.nout, self.ld_in, device=dev, dtype=torch.float32)
            # rows padded so a dgrad GEMM may ask for a few columns past nin (they come out as zeros)
            self.wt = torch.zeros(_pad32(self.nin), self.ld_out, device=dev, dtype=torch.float32)
        L.pe_copy_cols(P(lin.weight), self.nin, P(self.w), self.ld_in, self.nout, self.nin, round_tf32, st)
        if need_dgrad:
            L.pe_transpose(P(lin.weight), self.nin, P(self.wt), self.ld_out, self.nout, self.nin, round_tf32, st)
        self._ver = ver

    def forward(self, x, ldx, M, y, ldy, relu=False, accumulate=False, round_out=1, bias=True):
        L, st, P = native.lib(), native.stream_ptr(), native.ptr
        L.pe_linear_fwd(P(x), ldx, P(self.w), self.ld_in, P(self.lin.bias) if bias else None, None, P(y), ldy, M,
                        self.nout, self.nin, int(relu), int(accumulate), round_out, None, st)

    def backward(self, x, ldx, M, dy, lddy, grad_of, dx=None, lddx=0, dx_cols=None, bias_grad=True,
                 accumulate_w=False):
        """dy must be TF32-rounded, zero in its padding columns (lddy = ld_out).  dx_cols limits how many
        input columns get a gradient (the rest of the concat buffer does not need one)."""
        L, st, P = native.lib(), native.stream_ptr(), native.ptr
        gw = grad_of(self.lin.weight)
        if accumulate_w:
            tmp = torch.empty_like(gw)
            L.pe_linear_wgrad(P(x), ldx, P(dy), lddy, P(tmp), self.nin, M, self.nout, self.nin, st)
            gw.add_(tmp)
        else:
            L.pe_linear_wgrad(P(x), ldx, P(dy), lddy, P(gw), self.nin, M, self.nout, self.nin, st)
        if bias_grad:
            L.pe_colsum(P(dy), lddy, P(grad_of(self.lin.bias)), M, self.nout, 0, st)
        if dx is not None:
            n = dx_cols if dx_cols is not None else self.nin
            L.pe_linear_fwd(P(dy), lddy, P(self.wt), self.ld_out, None, None, P(dx), lddx, M, n, self.nout, 0, 0, 0,
                            None, st)


PERSISTENT_LSTM = [True]      # sequences (S > 1): the whole recurrence in one persistent launch (pe_lstm_seq_*)


class LSTMOp:
    """Single-layer seq-major nn.LSTM (gates i,f,g,o): one big input-projection GEMM, then the recurrence.  For
    sequences the recurrence is ONE persistent launch forward and one backward (W_hh sliced across the grid and
    resident in shared memory, a grid barrier per timestep, cell fused; fp32 on the un-rounded W_hh); single steps
    (streaming rollout beyond the fused head's 8 rows) and carried-state backward keep the per-step recurrent GEMM +
    cell kernel pair."""

    def __init__(self, lstm, ld_in):
        self.lstm = lstm
        self.nin, self.H = lstm.input_size, lstm.hidden_size
        self.ld_in = ld_in
        self._ver = None
        self.w_ih = self.w_ih_t = self.w_hh = self.w_hh_t = None

    def pack(self, need_dgrad, round_tf32=1):
        lstm = self.lstm
        _dev_check(lstm.weight_ih_l0)
        ver = (_params_version([lstm.weight_ih_l0, lstm.weight_hh_l0]), need_dgrad)
        if ver == self._ver:
            return
        L, st, P = native.lib(), native.stream_ptr(), native.ptr
        dev = lstm.weight_ih_l0.device
        G, H = 4 * self.H, self.H
        if self.w_ih is None or self.w_ih.device != dev:
            self.w_ih = torch.zeros(G, self.ld_in, device=dev, dtype=torch.float32)
            self.w_ih_t = torch.zeros(_pad32(self.nin), G, device=dev, dtype=torch.float32)
            self.w_hh = torch.empty(G, H, device=dev, dtype=torch.float32)
            self.w_hh_t = torch.empty(H, G, device=dev, dtype=torch.float32)
        L.pe_copy_cols(P(lstm.weight_ih_l0), self.nin, P(self.w_ih), self.ld_in, G, self.nin, round_tf32, st)
        L.pe_copy_cols(P(lstm.weight_hh_l0), H, P(self.w_hh), H, G, H, round_tf32, st)
        if need_dgrad:
            L.pe_transpose(P(lstm.weight_ih_l0), self.nin, P(self.w_ih_t), G, G, self.nin, round_tf32, st)
            L.pe_transpose(P(lstm.weight_hh_l0), H, P(self.w_hh_t), G, G, H, round_tf32, st)
        self._ver = ver

    def forward(self, x, S, N, h0=None, c0=None, need_grad=False):
        """x: [S*N, ld_in] (TF32-rounded, zero padded).  Returns (h_all [S*N, H], h_last, c_last, ctx)."""
        L, st, P = native.lib(), native.stream_ptr(), native.ptr
        lstm = self.lstm
        dev = x.device
        G, H = 4 * self.H, self.H
        gx = torch.empty(S * N, G, device=dev, dtype=torch.float32)
        L.pe_linear_fwd(P(x), self.ld_in, P(self.w_ih), self.ld_in, None, None, P(gx), G, S * N, G, self.nin, 0, 0, 0,
                        None, st)
        h_all = torch.empty(S * N, H, device=dev, dtype=torch.float32)
        c_all = torch.empty(S * N, H, device=dev, dtype=torch.float32)
        act = torch.empty(S * N, G, device=dev, dtype=torch.float32) if need_grad else None
        if PERSISTENT_LSTM[0] and S > 1 and L.pe_lstm_seq_supported(N, H, 0):
            _dev_check(lstm.weight_hh_l0)
            L.pe_lstm_seq_fwd(P(gx), P(lstm.weight_hh_l0), P(lstm.bias_ih_l0), P(lstm.bias_hh_l0), P(h0), P(c0),
                              P(h_all), P(c_all), P(act), S, N, H, 1, st)
            ctx = dict(x=x, S=S, N=N, h_all=h_all, c_all=c_all, act=act, h0=h0, c0=c0) if need_grad else None
            return h_all, h_all[(S - 1) * N:], c_all[(S - 1) * N:], ctx
        gh = torch.empty(N, G, device=dev, dtype=torch.float32)
        h_prev, c_prev = h0, c0
        for t in range(S):
            gh_t = None
            if h_prev is not None:
                L.pe_linear_fwd(P(h_prev), H, P(self.w_hh), H, None, None, P(gh), G, N, G, H, 0, 0, 0, None, st)
                gh_t = gh
            L.pe_lstm_cell_fwd(P(gx[t * N:]), G, P(gh_t), G, P(lstm.bias_ih_l0), P(lstm.bias_hh_l0), P(c_prev),
                               P(c_all[t * N:]), P(h_all[t * N:]), H, P(act[t * N:]) if act is not None else None, N,
                               H, 1, st)
            h_prev, c_prev = h_all[t * N:(t + 1) * N], c_all[t * N:(t + 1) * N]
        ctx = None
        if need_grad:
            ctx = dict(x=x, S=S, N=N, h_all=h_all, c_all=c_all, act=act, h0=h0, c0=c0)
        return h_all, h_prev, c_prev, ctx

    def backward(self, ctx, dh_all, grad_of, dx=None, lddx=0, dx_cols=None):
        """dh_all: [S*N, H] gradient w.r.t. every hidden output.  Writes parameter grads; optionally
        dx[:, :dx_cols] = gradient w.r.t. the LSTM input."""
        L, st, P = native.lib(), native.stream_ptr(), native.ptr
        lstm = self.lstm
        S, N, H = ctx["S"], ctx["N"], self.H
        G = 4 * H
        dev = dh_all.device
        h_all, c_all, act = ctx["h_all"], ctx["c_all"], ctx["act"]
        dg = torch.empty(S * N, G, device=dev, dtype=torch.float32)
        if (PERSISTENT_LSTM[0] and S > 1 and ctx["h0"] is None and dh_all.stride(0) == H and
                L.pe_lstm_seq_supported(N, H, 1)):
            # whole BPTT recurrence in one launch; dg comes out TF32-rounded for the GEMMs below
            L.pe_lstm_seq_bwd(P(dh_all), P(lstm.weight_hh_l0), P(act), P(c_all), P(ctx["c0"]), P(dg), S, N, H, 1, st)
        else:
            dh_rec = None
            dc = None
            dhr_buf = torch.empty(N, H, device=dev, dtype=torch.float32)
            dc_bufs = [torch.empty(N, H, device=dev, dtype=torch.float32) for _ in range(2)]
            for t in range(S - 1, -1, -1):
                c_prev = c_all[(t - 1) * N:] if t > 0 else ctx["c0"]
                dc_new = dc_bufs[t & 1]
                L.pe_lstm_cell_bwd(P(dh_all[t * N:]), H, P(dh_rec), P(dc), P(act[t * N:]), P(c_prev),
                                   P(c_all[t * N:]), P(dg[t * N:]), G, P(dc_new), N, H, st)
                dc = dc_new
                if t > 0 or ctx["h0"] is not None:
                    L.pe_linear_fwd(P(dg[t * N:]), G, P(self.w_hh_t), G, None, None, P(dhr_buf), H, N, H, G, 0, 0, 0,
                                    None, st)
                    dh_rec = dhr_buf
            # round dg once for the three big GEMMs that consume it as a TF32 operand
            L.pe_copy_cols(P(dg), G, P(dg), G, S * N, G, 1, st)
        L.pe_linear_wgrad(P(ctx["x"]), self.ld_in, P(dg), G, P(grad_of(lstm.weight_ih_l0)), self.nin, S * N, G,
                          self.nin, st)
        g_hh = grad_of(lstm.weight_hh_l0)
        if S > 1:
            L.pe_linear_wgrad(P(h_all), H, P(dg[N:]), G, P(g_hh), H, (S - 1) * N, G, H, st)
        else:
            g_hh.zero_()
        if ctx["h0"] is not None:
            tmp = torch.empty_like(g_hh)
            L.pe_linear_wgrad(P(ctx["h0"]), H, P(dg), G, P(tmp), H, N, G, H, st)
            g_hh.add_(tmp)
        g_bih = grad_of(lstm.bias_ih_l0)
        L.pe_colsum(P(dg), G, P(g_bih), S * N, G, 0, st)
        L.pe_copy_cols(P(g_bih), G, P(grad_of(lstm.bias_hh_l0)), G, 1, G, 0, st)
        if dx is not None:
            n = dx_cols if dx_cols is not None else self.nin
            L.pe_linear_fwd(P(dg), G, P(self.w_ih_t), G, None, None, P(dx), lddx, S * N, n, G, 0, 0, 0, None, st)

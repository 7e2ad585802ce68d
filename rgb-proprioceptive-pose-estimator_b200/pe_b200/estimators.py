"""Forward / backward drivers ("cores") of the four estimators, expressed as sequences of C-ABI calls.

A core owns no parameters: it reads them from the nn.Module that mirrors the reference class and
writes gradients wherever `grad_of(param)` points (fresh tensors for autograd, slices of a flat arena
for the fused trainer).  Layout of the fusion buffer ("cat"), one row per frame:

    [ latent features | 3136 aux features | 7 proprio (or pre_out - x0bar) | zero padding to 32 ]

The trunk's fc epilogue, the aux kernel and a 7-column copy write straight into it, so the reference's
torch.cat calls (models/naive.py:333-340, models/time_sensitive.py:491-498) never materialise.
"""
import torch

from . import native
from .engine import DepthOp, LinearOp, LSTMOp, _pad32

AUX_DIM = 3136
OUT_LD = 8          # 7-D poses live in 8-float rows so every row stays 16-byte aligned for TMA


# test hook: when set to a list, the naive estimators' MLPs append every layer's (post-ReLU) output buffer, in
# execution order (read by the parity tests, tests/model_checks.py)
CAPTURE_HEAD_OUTPUTS = [None]
FUSED_HEAD = [True]       # rollout-sized inference (<= 8 frames, no gradients) goes through pe_fused_head
FUSED_HEAD_MAX_ROWS = 8


def _fused_ok(need_grad, rows):
    return FUSED_HEAD[0] and not need_grad and rows <= FUSED_HEAD_MAX_ROWS


def _ticket(core, dev):
    """Persistent zeroed arrival counter of the fused head kernel (it resets itself after every launch)."""
    t = getattr(core, "_fh_ticket", None)
    if t is None or t.device != dev:
        t = torch.zeros(1, device=dev, dtype=torch.int32)
        core._fh_ticket = t
    return t


def _f32(x, name):
    if x.dtype != torch.float32:
        raise native.PeError("%s must be float32 (got %s)" % (name, x.dtype))
    if not x.is_cuda:
        raise native.PeError("%s must be a CUDA tensor: the B200 path has no CPU fallback" % name)
    return x.contiguous()


def _pad_grad(d, M, dev, round_tf32=1):
    """Copy an (M,7) gradient (any strides) into a zeroed, TF32-rounded (M, 8) buffer."""
    L, st, P = native.lib(), native.stream_ptr(), native.ptr
    buf = torch.zeros(M, OUT_LD, device=dev, dtype=torch.float32)
    if d is not None:
        d = d.reshape(M, 7)
        if d.stride(1) != 1:
            d = d.contiguous()
        L.pe_copy_cols(P(d), d.stride(0), P(buf), OUT_LD, M, 7, round_tf32, st)
    return buf


class TrunkCore:
    """Stand-alone trunk: (B,3,H,W) -> (B, latent)."""

    def __init__(self, net):
        self.net = net

    def params(self):
        return list(self.net.parameters())

    def forward(self, inputs, training, need_grad, state=None):
        img = _f32(inputs[0], "img")
        eng = self.net.pe_engine()
        B = img.shape[0]
        n = self.net.fc.out_features
        out = torch.empty(B, n, device=img.device, dtype=torch.float32)
        ctx = eng.forward(img, training, need_grad, out, n)
        return (out,), (eng, ctx, B, n), None

    def backward(self, saved, douts, grad_of, on_ready=None):
        eng, ctx, B, n = saved
        d = douts[0].contiguous()
        eng.backward(ctx, d, n, None, 0, grad_of, on_ready)


class NaiveObjectCore:
    """NaiveObjectStateEstimator (reference models/naive.py:130-367)."""

    def __init__(self, m):
        self.m = m
        self.latent = m.feature_net.module.fc.out_features
        self.use_aux = m.early_features is not None
        self.n_in = self.latent + (AUX_DIM if self.use_aux else 0) + (7 if m.use_proprioception else 0)
        self.ld_cat = _pad32(self.n_in)
        self.fcs = []
        for i in range(m.n_fc):
            lin = getattr(m, "fc%d" % i).module
            self.fcs.append(LinearOp(lin, ld_in=self.ld_cat if i == 0 else None))
        self.depth = DepthOp(m.depth_nets[0].module) if (self.use_aux and m.use_depth) else None

    def _engine(self):
        aux = self.m.aux_nets[0].module[0] if self.use_aux else None
        return self.m.feature_net.module.pe_engine(aux, True)

    def params(self):
        return list(self.m.parameters())

    def forward(self, inputs, training, need_grad, state=None):
        L, st, P = native.lib(), native.stream_ptr(), native.ptr
        img = _f32(inputs[0], "img")
        B = img.shape[0]
        dev = img.device
        eng = self._engine()
        cat = torch.zeros(B, self.ld_cat, device=dev, dtype=torch.float32)
        aux_view = cat[:, self.latent:] if self.use_aux else None
        tctx = eng.forward(img, training, need_grad, cat, self.ld_cat, aux_view, self.ld_cat,
                           aux_round=self.depth is None)
        dctx = self.depth.forward(inputs[2], aux_view, self.ld_cat, B, need_grad) if self.depth else None
        col = self.latent + (AUX_DIM if self.use_aux else 0)
        if _fused_ok(need_grad, B):
            # one launch: proprio injection + fc0 over the whole grid, remaining layers by the last CTA
            x0 = _f32(inputs[1], "self_measurement").reshape(B, 7) if self.m.use_proprioception else None
            lin0 = self.fcs[0].lin
            scratch = torch.empty(B, lin0.out_features, device=dev, dtype=torch.float32)
            out = torch.zeros(B, OUT_LD, device=dev, dtype=torch.float32)
            tail = [(op.lin.weight, op.lin.bias, op.nout, True) for op in self.fcs[1:]]
            native.fused_head(cat, self.ld_cat, B, self.n_in, lin0.weight, lin0.out_features, scratch,
                              _ticket(self, dev), out, OUT_LD, inj=x0, ld_inj=7, inj_col=col, b1=lin0.bias,
                              relu_a=True, tail=tail)
            return (out[:, :7],), None, None
        if self.m.use_proprioception:
            x0 = _f32(inputs[1], "self_measurement").reshape(B, 7)
            L.pe_copy_cols(P(x0), 7, P(cat[:, col:]), self.ld_cat, B, 7, 1, st)
        hs = [cat]
        x, ldx = cat, self.ld_cat
        for i, op in enumerate(self.fcs):
            op.pack(need_grad)
            y = torch.zeros(B, op.ld_out if i < len(self.fcs) - 1 else OUT_LD, device=dev, dtype=torch.float32)
            # ReLU after every layer including the last one (reference quirk, models/naive.py:343-345)
            op.forward(x, ldx, B, y, y.shape[1], relu=True, round_out=1 if i < len(self.fcs) - 1 else 0)
            hs.append(y)
            if CAPTURE_HEAD_OUTPUTS[0] is not None:
                CAPTURE_HEAD_OUTPUTS[0].append(y[:, :op.nout])
            x, ldx = y, y.shape[1]
        return (x[:, :7],), (eng, tctx, hs, B, dctx), None

    def backward(self, saved, douts, grad_of, on_ready=None):
        L, st, P = native.lib(), native.stream_ptr(), native.ptr
        eng, tctx, hs, B, dctx = saved
        dev = hs[0].device
        d = _pad_grad(douts[0], B, dev, 0)
        for i in range(len(self.fcs) - 1, -1, -1):
            op = self.fcs[i]
            y, x = hs[i + 1], hs[i]
            ldy = y.shape[1]
            dz = torch.zeros(B, ldy, device=dev, dtype=torch.float32)
            L.pe_relu_bwd(P(d), d.shape[1], P(y), ldy, P(dz), ldy, B, op.nout, st)
            L.pe_copy_cols(P(dz), ldy, P(dz), ldy, B, op.nout, 1, st)          # TF32-round the MMA operand
            if i > 0:
                dx = torch.empty(B, x.shape[1], device=dev, dtype=torch.float32)
                op.backward(x, x.shape[1], B, dz, ldy, grad_of, dx, x.shape[1])
            else:
                ncols = self.latent + (AUX_DIM if self.use_aux else 0)
                dx = torch.empty(B, self.ld_cat, device=dev, dtype=torch.float32)
                op.backward(x, self.ld_cat, B, dz, ldy, grad_of, dx, self.ld_cat, dx_cols=ncols)
            d = dx
        if on_ready is not None:
            on_ready([p for op in self.fcs for p in op.lin.parameters()])
        if self.depth is not None:
            self.depth.backward(dctx, d[:, self.latent:], self.ld_cat, B, grad_of)
            if on_ready is not None:
                on_ready(list(self.depth.norm.parameters()))
        eng.backward(tctx, d, self.ld_cat, d[:, self.latent:] if self.use_aux else None, self.ld_cat, grad_of, on_ready)


class NaiveEefCore:
    """NaiveEndEffectorStateEstimator (reference models/naive.py:8-127): no aux branch, two MLPs."""

    def __init__(self, m):
        self.m = m
        self.latent = m.feature_net.fc.out_features
        self.ld_cat = _pad32(self.latent + 7)
        self.pre = [LinearOp(getattr(m, "pre_fc%d" % i)) for i in range(m.n_pre_hidden)]
        self.post = [LinearOp(getattr(m, "post_fc%d" % i), ld_in=self.ld_cat if i == 0 else None)
                     for i in range(m.n_post_hidden)]

    def params(self):
        return list(self.m.parameters())

    def _mlp(self, ops, x, ldx, B, need_grad, dev):
        hs = [x]
        for i, op in enumerate(ops):
            op.pack(need_grad)
            last = i == len(ops) - 1
            y = torch.zeros(B, OUT_LD if last else op.ld_out, device=dev, dtype=torch.float32)
            op.forward(x, ldx, B, y, y.shape[1], relu=True, round_out=0 if last else 1)
            hs.append(y)
            if CAPTURE_HEAD_OUTPUTS[0] is not None:
                CAPTURE_HEAD_OUTPUTS[0].append(y[:, :op.nout])
            x, ldx = y, y.shape[1]
        return hs

    def forward(self, inputs, training, need_grad, state=None):
        L, st, P = native.lib(), native.stream_ptr(), native.ptr
        img = _f32(inputs[0], "img")
        x0 = _f32(inputs[1], "self_measurement").reshape(-1, 7)
        B = img.shape[0]
        dev = img.device
        eng = self.m.feature_net.pe_engine()
        cat = torch.zeros(B, self.ld_cat, device=dev, dtype=torch.float32)
        tctx = eng.forward(img, training, need_grad, cat, self.ld_cat)
        if _fused_ok(need_grad, B):
            # two launches: pre-measurement MLP (+ measurement difference into the fusion rows), post MLP
            outs = []
            for ops, k_x, with_diff in ((self.pre, self.latent, True), (self.post, self.latent + 7, False)):
                lin0 = ops[0].lin
                scratch = torch.empty(B, lin0.out_features, device=dev, dtype=torch.float32)
                out = torch.zeros(B, OUT_LD, device=dev, dtype=torch.float32)
                tail = [(op.lin.weight, op.lin.bias, op.nout, True) for op in ops[1:]]
                native.fused_head(cat, self.ld_cat, B, k_x, lin0.weight, lin0.out_features, scratch,
                                  _ticket(self, dev), out, OUT_LD, b1=lin0.bias, relu_a=True, tail=tail,
                                  meas=x0 if with_diff else None, ld_meas=7, diff=cat if with_diff else None,
                                  ld_diff=self.ld_cat, diff_col=self.latent)
                outs.append(out)
            return (outs[0][:, :7], outs[1][:, :7]), None, None
        hs_pre = self._mlp(self.pre, cat, self.ld_cat, B, need_grad, dev)
        pre_out = hs_pre[-1]
        # measurement_diff = pre_out - x0bar  -> columns [latent, latent+7) of the fusion buffer
        diff = cat[:, self.latent:]
        L.pe_axpby_cols(P(pre_out), OUT_LD, P(x0), 7, P(diff), self.ld_cat, B, 7, 1.0, -1.0, 1, st)
        hs_post = self._mlp(self.post, cat, self.ld_cat, B, need_grad, dev)
        return (pre_out[:, :7], hs_post[-1][:, :7]), (eng, tctx, hs_pre, hs_post, B), None

    def _mlp_bwd(self, ops, hs, d, B, grad_of, dev, first_cols, first_ld):
        L, st, P = native.lib(), native.stream_ptr(), native.ptr
        for i in range(len(ops) - 1, -1, -1):
            op = ops[i]
            y, x = hs[i + 1], hs[i]
            ldy = y.shape[1]
            dz = torch.zeros(B, ldy, device=dev, dtype=torch.float32)
            L.pe_relu_bwd(P(d), d.shape[1], P(y), ldy, P(dz), ldy, B, op.nout, st)
            L.pe_copy_cols(P(dz), ldy, P(dz), ldy, B, op.nout, 1, st)
            if i > 0:
                dx = torch.empty(B, x.shape[1], device=dev, dtype=torch.float32)
                op.backward(x, x.shape[1], B, dz, ldy, grad_of, dx, x.shape[1])
            else:
                dx = torch.zeros(B, first_ld, device=dev, dtype=torch.float32)
                op.backward(x, first_ld, B, dz, ldy, grad_of, dx, first_ld, dx_cols=first_cols)
            d = dx
        return d

    def backward(self, saved, douts, grad_of, on_ready=None):
        L, st, P = native.lib(), native.stream_ptr(), native.ptr
        eng, tctx, hs_pre, hs_post, B = saved
        dev = hs_pre[0].device
        d_post = _pad_grad(douts[1], B, dev, 0)
        dcat_post = self._mlp_bwd(self.post, hs_post, d_post, B, grad_of, dev, min(self.latent + 8, self.ld_cat),
                                  self.ld_cat)
        # gradient reaching pre_out: direct (loss on pre_out) + through measurement_diff
        d_pre = _pad_grad(douts[0], B, dev, 0)
        L.pe_axpby_cols(P(d_pre), OUT_LD, P(dcat_post[:, self.latent:]), self.ld_cat, P(d_pre), OUT_LD, B, 7, 1.0,
                        1.0, 0, st)
        dcat_pre = self._mlp_bwd(self.pre, hs_pre, d_pre, B, grad_of, dev, self.latent, self.ld_cat)
        # features feed both MLPs
        L.pe_axpby_cols(P(dcat_pre), self.ld_cat, P(dcat_post), self.ld_cat, P(dcat_pre), self.ld_cat, B,
                        self.latent, 1.0, 1.0, 0, st)
        if on_ready is not None:
            on_ready([p for op in self.pre + self.post for p in op.lin.parameters()])
        eng.backward(tctx, dcat_pre, self.ld_cat, None, 0, grad_of, on_ready)


class TDOCore:
    """TemporallyDependentObjectStateEstimator (reference models/time_sensitive.py:277-533)."""

    def __init__(self, m):
        self.m = m
        self.latent = m.feature_net.module.fc.out_features
        self.use_aux = m.early_features is not None
        self.n_in = self.latent + (AUX_DIM if self.use_aux else 0) + (7 if m.use_proprioception else 0)
        self.ld_cat = _pad32(self.n_in)
        self.rnn = LSTMOp(m.rnn.module, self.ld_cat)
        self.fc0 = LinearOp(m.fc.module[0])
        self.fc1 = LinearOp(m.fc.module[1])
        self.depth = DepthOp(m.depth_nets[0].module) if (self.use_aux and m.use_depth) else None

    def _engine(self):
        aux = self.m.aux_nets[0].module[0] if self.use_aux else None
        return self.m.feature_net.module.pe_engine(aux, True)

    def params(self):
        return list(self.m.parameters())

    def forward(self, inputs, training, need_grad, state=None):
        """inputs: img (S,N,3,H,W), x0bar (S,N,7); state: (h, c) each (N, H) for rollout.  Returns the
        outputs, the saved context and the new state."""
        L, st, P = native.lib(), native.stream_ptr(), native.ptr
        img = _f32(inputs[0], "img")
        S, N = img.shape[0], img.shape[1]
        M = S * N
        dev = img.device
        eng = self._engine()
        cat = torch.zeros(M, self.ld_cat, device=dev, dtype=torch.float32)
        aux_view = cat[:, self.latent:] if self.use_aux else None
        tctx = eng.forward(img.reshape(M, *img.shape[2:]), training, need_grad, cat, self.ld_cat, aux_view,
                           self.ld_cat, aux_round=self.depth is None)
        dctx = self.depth.forward(inputs[2], aux_view, self.ld_cat, M, need_grad) if self.depth else None
        col = self.latent + (AUX_DIM if self.use_aux else 0)
        if S == 1 and _fused_ok(need_grad, N):
            # streaming step: proprio injection + LSTM gate projections over the whole grid, cell + both dense
            # layers by the last CTA; the carried state is updated in place
            x0 = _f32(inputs[1], "self_measurement").reshape(M, 7) if self.m.use_proprioception else None
            lstm = self.m.rnn.module
            Hd = self.rnn.H
            h0, c0 = state if state is not None else (None, None)
            h_out = h0 if h0 is not None else torch.empty(N, Hd, device=dev, dtype=torch.float32)
            c_out = c0 if c0 is not None else torch.empty(N, Hd, device=dev, dtype=torch.float32)
            gates = torch.empty(N, 4 * Hd, device=dev, dtype=torch.float32)
            out = torch.zeros(N, OUT_LD, device=dev, dtype=torch.float32)
            fc0, fc1 = self.fc0.lin, self.fc1.lin
            native.fused_head(cat, self.ld_cat, N, self.n_in, lstm.weight_ih_l0, 4 * Hd, gates, _ticket(self, dev),
                              out, OUT_LD, inj=x0, ld_inj=7, inj_col=col, k_h=Hd, w_h=lstm.weight_hh_l0, h_prev=h0,
                              b1=lstm.bias_ih_l0, b2=lstm.bias_hh_l0, lstm_hidden=Hd, c_prev=c0, c_out=c_out,
                              h_out=h_out, tail=[(fc0.weight, fc0.bias, fc0.out_features, False),
                                                 (fc1.weight, fc1.bias, fc1.out_features, False)])
            return (out[:, :7].reshape(S, N, 7),), None, (h_out, c_out)
        if self.m.use_proprioception:
            x0 = _f32(inputs[1], "self_measurement").reshape(M, 7)
            L.pe_copy_cols(P(x0), 7, P(cat[:, col:]), self.ld_cat, M, 7, 1, st)
        self.rnn.pack(need_grad)
        self.fc0.pack(need_grad)
        self.fc1.pack(need_grad)
        h0, c0 = state if state is not None else (None, None)
        h_all, h_last, c_last, rctx = self.rnn.forward(cat, S, N, h0, c0, need_grad)
        z = torch.empty(M, self.fc0.ld_out, device=dev, dtype=torch.float32)
        self.fc0.forward(h_all, self.rnn.H, M, z, self.fc0.ld_out, relu=False, round_out=1)
        out = torch.zeros(M, OUT_LD, device=dev, dtype=torch.float32)
        self.fc1.forward(z, self.fc0.ld_out, M, out, OUT_LD, relu=False, round_out=0)
        saved = (eng, tctx, rctx, cat, h_all, z, S, N, dctx)
        return (out[:, :7].reshape(S, N, 7),), saved, (h_last, c_last)

    def backward(self, saved, douts, grad_of, on_ready=None):
        eng, tctx, rctx, cat, h_all, z, S, N, dctx = saved
        M = S * N
        dev = cat.device
        d = _pad_grad(douts[0], M, dev, 1)
        dz = torch.empty(M, self.fc0.ld_out, device=dev, dtype=torch.float32)
        self.fc1.backward(z, self.fc0.ld_out, M, d, OUT_LD, grad_of, dz, self.fc0.ld_out)
        L, st, P = native.lib(), native.stream_ptr(), native.ptr
        L.pe_copy_cols(P(dz), self.fc0.ld_out, P(dz), self.fc0.ld_out, M, self.fc0.nout, 1, st)
        dh = torch.empty(M, self.rnn.H, device=dev, dtype=torch.float32)
        self.fc0.backward(h_all, self.rnn.H, M, dz, self.fc0.ld_out, grad_of, dh, self.rnn.H)
        ncols = self.latent + (AUX_DIM if self.use_aux else 0)
        dcat = torch.empty(M, self.ld_cat, device=dev, dtype=torch.float32)
        self.rnn.backward(rctx, dh, grad_of, dcat, self.ld_cat, dx_cols=ncols)
        if on_ready is not None:
            on_ready(list(self.m.rnn.parameters()) + list(self.m.fc.parameters()))
        if self.depth is not None:
            self.depth.backward(dctx, dcat[:, self.latent:], self.ld_cat, M, grad_of)
            if on_ready is not None:
                on_ready(list(self.depth.norm.parameters()))
        eng.backward(tctx, dcat, self.ld_cat, dcat[:, self.latent:] if self.use_aux else None, self.ld_cat, grad_of,
                     on_ready)


class TDOV2Core:
    """TemporallyDependentObjectStateEstimatorV2 (reference models/time_sensitive.py:536-804): one LSTM over the
    image features (latent + aux), one over the 7-D proprioceptive measurement, hidden states concatenated into
    Linear(H, H//4) -> Linear(H//4, 7)."""

    def __init__(self, m):
        self.m = m
        self.latent = m.feature_net.module.fc.out_features
        self.use_aux = m.early_features is not None
        self.n_in = self.latent + (AUX_DIM if self.use_aux else 0)
        self.ld_cat = _pad32(self.n_in)
        self.img_rnn = LSTMOp(m.img_rnn.module, self.ld_cat)
        self.pro_rnn = LSTMOp(m.proprio_rnn.module, OUT_LD)
        self.fc0 = LinearOp(m.fc.module[0])
        self.fc1 = LinearOp(m.fc.module[1])
        self.H1, self.H2 = self.img_rnn.H, self.pro_rnn.H
        self.depth = DepthOp(m.depth_nets[0].module) if (self.use_aux and m.use_depth) else None

    def _engine(self):
        aux = self.m.aux_nets[0].module[0] if self.use_aux else None
        return self.m.feature_net.module.pe_engine(aux, True)

    def params(self):
        return list(self.m.parameters())

    def forward(self, inputs, training, need_grad, state=None):
        """inputs: img (S,N,3,H,W), x0bar (S,N,7); state: ((h_img, c_img), (h_pro, c_pro)), each (N, H)."""
        L, st, P = native.lib(), native.stream_ptr(), native.ptr
        img = _f32(inputs[0], "img")
        S, N = img.shape[0], img.shape[1]
        M = S * N
        dev = img.device
        eng = self._engine()
        cat = torch.zeros(M, self.ld_cat, device=dev, dtype=torch.float32)
        aux_view = cat[:, self.latent:] if self.use_aux else None
        tctx = eng.forward(img.reshape(M, *img.shape[2:]), training, need_grad, cat, self.ld_cat, aux_view,
                           self.ld_cat, aux_round=self.depth is None)
        dctx = self.depth.forward(inputs[2], aux_view, self.ld_cat, M, need_grad) if self.depth else None
        x0 = _f32(inputs[1], "self_measurement").reshape(M, 7)
        x0p = torch.zeros(M, OUT_LD, device=dev, dtype=torch.float32)      # TF32-rounded, zero padded to 8
        L.pe_copy_cols(P(x0), 7, P(x0p), OUT_LD, M, 7, 1, st)
        for op in (self.img_rnn, self.pro_rnn, self.fc0, self.fc1):
            op.pack(need_grad)
        s_img, s_pro = state if state is not None else ((None, None), (None, None))
        h1, h1_last, c1_last, rctx1 = self.img_rnn.forward(cat, S, N, s_img[0], s_img[1], need_grad)
        h2, h2_last, c2_last, rctx2 = self.pro_rnn.forward(x0p, S, N, s_pro[0], s_pro[1], need_grad)
        Hc = self.H1 + self.H2
        hcat = torch.empty(M, Hc, device=dev, dtype=torch.float32)          # torch.cat((img_h, proprio_h), -1)
        L.pe_copy_cols(P(h1), self.H1, P(hcat), Hc, M, self.H1, 0, st)
        L.pe_copy_cols(P(h2), self.H2, P(hcat[:, self.H1:]), Hc, M, self.H2, 0, st)
        z = torch.empty(M, self.fc0.ld_out, device=dev, dtype=torch.float32)
        self.fc0.forward(hcat, Hc, M, z, self.fc0.ld_out, relu=False, round_out=1)
        out = torch.zeros(M, OUT_LD, device=dev, dtype=torch.float32)
        self.fc1.forward(z, self.fc0.ld_out, M, out, OUT_LD, relu=False, round_out=0)
        saved = (eng, tctx, rctx1, rctx2, cat, hcat, z, S, N, dctx)
        return (out[:, :7].reshape(S, N, 7),), saved, ((h1_last, c1_last), (h2_last, c2_last))

    def backward(self, saved, douts, grad_of, on_ready=None):
        L, st, P = native.lib(), native.stream_ptr(), native.ptr
        eng, tctx, rctx1, rctx2, cat, hcat, z, S, N, dctx = saved
        M = S * N
        dev = cat.device
        Hc = self.H1 + self.H2
        d = _pad_grad(douts[0], M, dev, 1)
        dz = torch.empty(M, self.fc0.ld_out, device=dev, dtype=torch.float32)
        self.fc1.backward(z, self.fc0.ld_out, M, d, OUT_LD, grad_of, dz, self.fc0.ld_out)
        L.pe_copy_cols(P(dz), self.fc0.ld_out, P(dz), self.fc0.ld_out, M, self.fc0.nout, 1, st)
        dhcat = torch.empty(M, Hc, device=dev, dtype=torch.float32)
        self.fc0.backward(hcat, Hc, M, dz, self.fc0.ld_out, grad_of, dhcat, Hc)
        dh1 = torch.empty(M, self.H1, device=dev, dtype=torch.float32)
        dh2 = torch.empty(M, self.H2, device=dev, dtype=torch.float32)
        L.pe_copy_cols(P(dhcat), Hc, P(dh1), self.H1, M, self.H1, 0, st)
        L.pe_copy_cols(P(dhcat[:, self.H1:]), Hc, P(dh2), self.H2, M, self.H2, 0, st)
        self.pro_rnn.backward(rctx2, dh2, grad_of)
        dcat = torch.empty(M, self.ld_cat, device=dev, dtype=torch.float32)
        self.img_rnn.backward(rctx1, dh1, grad_of, dcat, self.ld_cat, dx_cols=self.n_in)
        if on_ready is not None:
            m = self.m
            on_ready(list(m.img_rnn.parameters()) + list(m.proprio_rnn.parameters()) + list(m.fc.parameters()))
        if self.depth is not None:
            self.depth.backward(dctx, dcat[:, self.latent:], self.ld_cat, M, grad_of)
            if on_ready is not None:
                on_ready(list(self.depth.norm.parameters()))
        eng.backward(tctx, dcat, self.ld_cat, dcat[:, self.latent:] if self.use_aux else None, self.ld_cat, grad_of,
                     on_ready)


class TDCore:
    """TemporallyDependentStateEstimator (reference models/time_sensitive.py:8-274): two LSTMs in series;
    the aux conv is an unregistered, frozen module (reference quirk, :77-78,102-115)."""

    def __init__(self, m):
        self.m = m
        self.latent = m.feature_net.fc.out_features
        self.use_aux = m.early_features is not None
        self.n_feat = self.latent + (AUX_DIM if self.use_aux else 0)
        self.ld_cat = _pad32(self.n_feat + 7)
        self.pre_rnn = LSTMOp(m.pre_measurement_rnn, self.ld_cat)
        self.post_rnn = LSTMOp(m.post_measurement_rnn, self.ld_cat)
        self.pre_fc = LinearOp(m.pre_measurement_fc)
        self.post_fc = LinearOp(m.post_measurement_fc)
        self._aux_dev = None
        # like the aux conv, the depth net sits in a plain python list in the reference: unregistered and frozen
        self.depth = DepthOp(m.depth_nets[0], trainable=False) if (self.use_aux and m.use_depth) else None

    def _engine(self, dev):
        aux = None
        if self.use_aux:
            aux = self.m.aux_nets[0][0]
            if aux.weight.device != dev:
                # the reference never moves this unregistered conv (it only runs on CPU there); the
                # accelerated path keeps its values and places them next to the rest of the model
                aux.to(dev)
            for p in aux.parameters():
                p.requires_grad_(False)
        return self.m.feature_net.pe_engine(aux, False)

    def params(self):
        return list(self.m.parameters())

    def forward(self, inputs, training, need_grad, state=None):
        L, st, P = native.lib(), native.stream_ptr(), native.ptr
        img = _f32(inputs[0], "img")
        S, N = img.shape[0], img.shape[1]
        M = S * N
        dev = img.device
        x0 = _f32(inputs[1], "self_measurement").reshape(M, 7)
        eng = self._engine(dev)
        cat = torch.zeros(M, self.ld_cat, device=dev, dtype=torch.float32)
        aux_view = cat[:, self.latent:] if self.use_aux else None
        tctx = eng.forward(img.reshape(M, *img.shape[2:]), training, need_grad, cat, self.ld_cat, aux_view,
                           self.ld_cat, aux_round=self.depth is None)
        dctx = self.depth.forward(inputs[2], aux_view, self.ld_cat, M, need_grad) if self.depth else None
        s_pre, s_post = state if state is not None else ((None, None), (None, None))
        if S == 1 and _fused_ok(need_grad, N):
            # streaming step in two launches: pre-measurement LSTM + fc (+ measurement difference written into the
            # fusion rows), then the post-measurement LSTM + fc
            outs, new_state = [], []
            m = self.m
            for lstm, fc, k_x, (h0, c0), with_diff in (
                    (m.pre_measurement_rnn, m.pre_measurement_fc, self.n_feat, s_pre, True),
                    (m.post_measurement_rnn, m.post_measurement_fc, self.n_feat + 7, s_post, False)):
                Hd = lstm.hidden_size
                h_out = h0 if h0 is not None else torch.empty(N, Hd, device=dev, dtype=torch.float32)
                c_out = c0 if c0 is not None else torch.empty(N, Hd, device=dev, dtype=torch.float32)
                gates = torch.empty(N, 4 * Hd, device=dev, dtype=torch.float32)
                out = torch.zeros(N, OUT_LD, device=dev, dtype=torch.float32)
                native.fused_head(cat, self.ld_cat, N, k_x, lstm.weight_ih_l0, 4 * Hd, gates, _ticket(self, dev), out,
                                  OUT_LD, k_h=Hd, w_h=lstm.weight_hh_l0, h_prev=h0, b1=lstm.bias_ih_l0,
                                  b2=lstm.bias_hh_l0, lstm_hidden=Hd, c_prev=c0, c_out=c_out, h_out=h_out,
                                  tail=[(fc.weight, fc.bias, fc.out_features, False)],
                                  meas=x0 if with_diff else None, ld_meas=7, diff=cat if with_diff else None,
                                  ld_diff=self.ld_cat, diff_col=self.n_feat)
                outs.append(out[:, :7].reshape(S, N, 7))
                new_state.append((h_out, c_out))
            return tuple(outs), None, tuple(new_state)
        for op in (self.pre_rnn, self.post_rnn, self.pre_fc, self.post_fc):
            op.pack(need_grad)
        h1, h1_last, c1_last, rctx1 = self.pre_rnn.forward(cat, S, N, s_pre[0], s_pre[1], need_grad)
        pre_out = torch.zeros(M, OUT_LD, device=dev, dtype=torch.float32)
        self.pre_fc.forward(h1, self.pre_rnn.H, M, pre_out, OUT_LD, relu=False, round_out=0)
        L.pe_axpby_cols(P(pre_out), OUT_LD, P(x0), 7, P(cat[:, self.n_feat:]), self.ld_cat, M, 7, 1.0, -1.0, 1, st)
        h2, h2_last, c2_last, rctx2 = self.post_rnn.forward(cat, S, N, s_post[0], s_post[1], need_grad)
        post_out = torch.zeros(M, OUT_LD, device=dev, dtype=torch.float32)
        self.post_fc.forward(h2, self.post_rnn.H, M, post_out, OUT_LD, relu=False, round_out=0)
        saved = (eng, tctx, rctx1, rctx2, cat, h1, h2, S, N, dctx)
        outs = (pre_out[:, :7].reshape(S, N, 7), post_out[:, :7].reshape(S, N, 7))
        return outs, saved, ((h1_last, c1_last), (h2_last, c2_last))

    def backward(self, saved, douts, grad_of, on_ready=None):
        L, st, P = native.lib(), native.stream_ptr(), native.ptr
        eng, tctx, rctx1, rctx2, cat, h1, h2, S, N, dctx = saved
        M = S * N
        dev = cat.device
        H1, H2 = self.pre_rnn.H, self.post_rnn.H
        d_post = _pad_grad(douts[1], M, dev, 1)
        dh2 = torch.empty(M, H2, device=dev, dtype=torch.float32)
        self.post_fc.backward(h2, H2, M, d_post, OUT_LD, grad_of, dh2, H2)
        dcat2 = torch.zeros(M, self.ld_cat, device=dev, dtype=torch.float32)
        self.post_rnn.backward(rctx2, dh2, grad_of, dcat2, self.ld_cat, dx_cols=min(self.n_feat + 8, self.ld_cat))
        d_pre = _pad_grad(douts[0], M, dev, 0)
        L.pe_axpby_cols(P(d_pre), OUT_LD, P(dcat2[:, self.n_feat:]), self.ld_cat, P(d_pre), OUT_LD, M, 7, 1.0, 1.0,
                        1, st)
        dh1 = torch.empty(M, H1, device=dev, dtype=torch.float32)
        self.pre_fc.backward(h1, H1, M, d_pre, OUT_LD, grad_of, dh1, H1)
        dcat1 = torch.empty(M, self.ld_cat, device=dev, dtype=torch.float32)
        self.pre_rnn.backward(rctx1, dh1, grad_of, dcat1, self.ld_cat, dx_cols=self.n_feat)
        L.pe_axpby_cols(P(dcat1), self.ld_cat, P(dcat2), self.ld_cat, P(dcat1), self.ld_cat, M, self.n_feat, 1.0,
                        1.0, 0, st)
        if on_ready is not None:
            m = self.m
            on_ready(list(m.pre_measurement_rnn.parameters()) + list(m.pre_measurement_fc.parameters()) +
                     list(m.post_measurement_rnn.parameters()) + list(m.post_measurement_fc.parameters()))
        if self.depth is not None:
            self.depth.backward(dctx, dcat1[:, self.latent:], self.ld_cat, M, grad_of)
        eng.backward(tctx, dcat1, self.ld_cat, dcat1[:, self.latent:] if self.use_aux else None, self.ld_cat,
                     grad_of, on_ready)

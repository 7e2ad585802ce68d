"""CPU oracle for the pose-estimator hot path.  TEST INFRASTRUCTURE ONLY.

A plain-PyTorch (CPU, fp32) restatement of the reference's algorithm for the one accelerated path:
ResNet-50 trunk -> auxiliary BN1 branch -> concat proprioception -> MLP / LSTM head -> pose loss ->
Adam.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this file; the product path (rgb-proprioceptive-pose-estimator_b200/) never does.

Parity pinning: the reference ships no tests or golden vectors for this path, so
this restatement is pinned against (a) the reference's own modules imported from /root/reference in
the build container or from the shipped copy oracle/_ref (tests/test_oracle_cpu.py, skipped where neither exists) and
(b) the fixtures under tests/golden/ that oracle/make_golden.py generated from those modules.

Third-party arithmetic that is not under /root/reference (requirements.txt leaves torch/torchvision
unpinned): torchvision.models.resnet50 (ResNet v1.5, stride on the 3x3) and torch.nn.{Conv2d,
BatchNorm2d,LSTM,Linear}; restated here functionally from their published definitions and checked
against torch 2.11.0 / torchvision 0.26.0 as installed.

Each function cites the reference lines it follows (paths relative to /root/reference).
"""
import contextlib
import math

import torch
import torch.nn.functional as F

# torchvision resnet50: (planes, blocks, stride) per stage -- torchvision/models/resnet.py:266-282
RESNET50_STAGES = ((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2))


# --------------------------------------------------------------------------------------------
# operand precision
# --------------------------------------------------------------------------------------------
# The CUDA path computes every dense contraction with TF32 OPERANDS (10-bit mantissa) and fp32 accumulation, as
# north_star's "fp32/TF32 tolerance" allows.  Against the plain fp32 restatement below that costs ~4e-3 on the
# outputs of a 50-layer network -- and far more on its gradients, because a ReLU network's gradient is
# discontinuous: an activation whose pre-activation moves across zero flips its mask, so a forward perturbation of
# relative size e changes a fraction ~e of the masks and the gradient by ~sqrt(e) (fp32 vs fp64 on the CPU already
# differ by 4e-3 in every trunk gradient; measured).  To check the backward LOGIC of the
# full-depth network to a tight tolerance the oracle can therefore be switched to the same operand precision:
# inside `tf32_operands()` every conv / linear operand is rounded to TF32 exactly where the CUDA path rounds it
# (cvt.rna.tf32: weights when they are packed, activations when they are produced, output gradients when BatchNorm /
# ReLU backward hands them to dgrad and wgrad), everything else (accumulation, BatchNorm, LSTM cell, loss) stays in the
# working precision -- run it in float64 and what remains between the two is accumulation order.
_TF32 = [False]
_FUSED_HEAD_ROWS = 8          # rollout-sized inference (<= 8 rows, no gradients) runs its head in plain fp32 FMA


@contextlib.contextmanager
def tf32_operands(on=True):
    prev = _TF32[0]
    _TF32[0] = bool(on)
    try:
        yield
    finally:
        _TF32[0] = prev


def round_tf32(x):
    """cvt.rna.tf32.f32: round the 24-bit significand to 11 bits, nearest, ties away from zero."""
    x32 = x.detach().to(torch.float32).contiguous()
    bits = x32.view(torch.int32)
    mag = bits & 0x7FFFFFFF
    rounded = (mag + 0x1000) & 0x7FFFE000
    out = (rounded | (bits & -0x80000000)).view(torch.float32)
    out = torch.where(torch.isfinite(x32), out, x32)
    return out.to(x.dtype)


class _RoundSTE(torch.autograd.Function):
    """Value rounded to TF32; gradient passed through (the CUDA path does not differentiate its roundings)."""

    @staticmethod
    def forward(ctx, x):
        return round_tf32(x)

    @staticmethod
    def backward(ctx, g):
        return g


def _rnd(x):
    return _RoundSTE.apply(x) if _TF32[0] else x


class _ConvTF32(torch.autograd.Function):
    """conv2d whose three GEMMs see TF32 operands: forward x (already rounded by its producer) * round(w);
    backward with r = round(dy): dx = r (*) round(w), dw = x (*) r."""

    @staticmethod
    def forward(ctx, x, w, stride, padding):
        wr = round_tf32(w)
        ctx.save_for_backward(x, wr)
        ctx.cfg = (stride, padding, w.shape)
        return F.conv2d(x, wr, stride=stride, padding=padding)

    @staticmethod
    def backward(ctx, g):
        x, wr = ctx.saved_tensors
        stride, padding, wshape = ctx.cfg
        r = round_tf32(g)
        dx = torch.nn.grad.conv2d_input(x.shape, wr, r, stride=stride, padding=padding) if ctx.needs_input_grad[0] else None
        dw = torch.nn.grad.conv2d_weight(x, wshape, r, stride=stride, padding=padding)
        return dx, dw, None, None


class _LinearTF32(torch.autograd.Function):
    """linear with TF32 operands: y = x round(w)^T + b; backward with r = round(dy): dx = r round(w), dw = r^T x,
    db = sum r (the CUDA path takes the bias gradient from the same rounded buffer)."""

    @staticmethod
    def forward(ctx, x, w, b):
        wr = round_tf32(w)
        ctx.save_for_backward(x, wr)
        ctx.has_bias = b is not None
        y = x.matmul(wr.t())
        return y + b if b is not None else y

    @staticmethod
    def backward(ctx, g):
        x, wr = ctx.saved_tensors
        r = round_tf32(g)
        dx = r.matmul(wr) if ctx.needs_input_grad[0] else None
        dw = r.reshape(-1, r.shape[-1]).t().matmul(x.reshape(-1, x.shape[-1]))
        db = r.reshape(-1, r.shape[-1]).sum(0) if ctx.has_bias else None
        return dx, dw, db


# Teacher forcing of the convolution outputs.  Even at identical operand precision two implementations of a 50-layer
# ReLU network disagree on ~1e-5 of their activations' signs (accumulation order), and every flipped ReLU mask is an
# O(1) change of that element's gradient: fp32 vs fp64 accumulation on the CPU, same TF32 operands, already differ by
# 0.09 (median) in the trunk gradients.  A tight check of the backward pass therefore has to run on the SAME masks:
# inside `forced_conv_outputs(ys)` every convolution still computes its own output (its relative error against the
# supplied tensor is recorded -- a per-layer forward check), but the VALUE that flows on is the supplied one while
# autograd still differentiates the oracle's own op.  With `ys` = the raw conv outputs the CUDA path saved for its
# backward pass, BatchNorm statistics, ReLU masks and x-hat are then common to both sides and the remaining gradient
# difference is arithmetic, not chaos.
_FORCE = [None]
FORCE_ERRORS = []


@contextlib.contextmanager
def forced_conv_outputs(ys):
    _FORCE[0] = iter(ys)
    del FORCE_ERRORS[:]
    try:
        yield FORCE_ERRORS
    finally:
        _FORCE[0] = None


RECORD = [None]          # when a list: every convolution appends its output (self-test of the forcing machinery)


def _force(y):
    if RECORD[0] is not None:
        RECORD[0].append(y.detach().clone())
    if _FORCE[0] is None:
        return y
    ext = next(_FORCE[0]).to(y.dtype)
    assert ext.shape == y.shape, (ext.shape, y.shape)
    FORCE_ERRORS.append(float((y.detach() - ext).abs().max() / ext.abs().max().clamp_min(1e-30)))
    return y + (ext - y).detach()


# The naive estimators' MLPs apply ReLU after every layer (models/naive.py:343-345): a handful of hidden units, so a
# single unit whose pre-activation sits within rounding of zero flips a visible share of the gradient.  Their layer
# outputs can be teacher-forced like the conv outputs: value AND mask come from the supplied post-ReLU tensor.
_FORCE_HEAD = [None]


@contextlib.contextmanager
def forced_head_outputs(ys):
    _FORCE_HEAD[0] = iter(ys) if ys is not None else None
    try:
        yield
    finally:
        _FORCE_HEAD[0] = None


def _relu_head(z):
    if _FORCE_HEAD[0] is None:
        return F.relu(z)
    ext = next(_FORCE_HEAD[0]).to(z.dtype)
    assert ext.shape == z.shape, (ext.shape, z.shape)
    own = F.relu(z)
    FORCE_ERRORS.append(float((own.detach() - ext).abs().max() / ext.abs().max().clamp_min(1e-30)))
    o = z * (ext > 0).to(z.dtype)
    return o + (ext - o).detach()


def _head_ops(fused):
    """(linear, round) used by a fusion head.  `fused`: the CUDA path runs rollout-sized inference heads (<= 8 rows,
    no gradients) as ONE fp32-FMA kernel on the un-rounded checkpoint weights (pe_fused_head) -- plain fp32 here too."""
    if _TF32[0] and not fused:
        return _linear, _rnd
    return F.linear, (lambda t: t)


class _RecurrentFP32(torch.autograd.Function):
    """h W_hh^T + b of a SEQUENCE's recurrence as the persistent LSTM kernel computes it: fp32 FMA on the un-rounded
    checkpoint weights (h itself was rounded where it was produced); backward on the TF32-rounded gate gradient the
    kernel writes out, r = round(dg):  dh = r W_hh (un-rounded weights),  dW_hh = r^T h and db = sum r (tensor-core
    GEMM / column sum over the same rounded buffer)."""

    @staticmethod
    def forward(ctx, h, w, b):
        ctx.save_for_backward(h, w)
        return h.matmul(w.t()) + b

    @staticmethod
    def backward(ctx, g):
        h, w = ctx.saved_tensors
        r = round_tf32(g)
        return r.matmul(w), r.t().matmul(h), r.sum(0)


def _conv(x, w, stride=1, padding=0):
    if _TF32[0]:
        return _force(_ConvTF32.apply(x, w, stride, padding))
    return _force(F.conv2d(x, w, stride=stride, padding=padding))


def _linear(x, w, b=None):
    if _TF32[0]:
        return _LinearTF32.apply(x, w, b)
    return F.linear(x, w, b)


# --------------------------------------------------------------------------------------------
# trunk
# --------------------------------------------------------------------------------------------
def _bn(x, sd, prefix, training, momentum=0.1, eps=1e-5):
    """nn.BatchNorm2d forward incl. running-stat update (torchvision resnet.py:134-138,198)."""
    rm, rv = sd[prefix + "running_mean"], sd[prefix + "running_var"]
    if training:
        sd[prefix + "num_batches_tracked"] += 1
    return F.batch_norm(x, rm, rv, sd[prefix + "weight"], sd[prefix + "bias"], training, momentum, eps)


def _bottleneck(x, sd, prefix, stride, training):
    """torchvision Bottleneck.forward (resnet.py:143-163), stride on conv2 (v1.5)."""
    identity = x
    out = _conv(x, sd[prefix + "conv1.weight"])
    out = _rnd(F.relu(_bn(out, sd, prefix + "bn1.", training)))
    out = _conv(out, sd[prefix + "conv2.weight"], stride=stride, padding=1)
    out = _rnd(F.relu(_bn(out, sd, prefix + "bn2.", training)))
    out = _conv(out, sd[prefix + "conv3.weight"])
    out = _bn(out, sd, prefix + "bn3.", training)
    if prefix + "downsample.0.weight" in sd:
        identity = _conv(x, sd[prefix + "downsample.0.weight"], stride=stride)
        identity = _rnd(_bn(identity, sd, prefix + "downsample.1.", training))
    return _rnd(F.relu(out + identity))


def _basic_block(x, sd, prefix, stride, training):
    """torchvision BasicBlock.forward (resnet.py:59-106; ResNet-18/34, reached through import_resnet(18, ...),
    util/model_utils.py:130-136): two 3x3 convolutions, stride on the first."""
    identity = x
    out = _conv(x, sd[prefix + "conv1.weight"], stride=stride, padding=1)
    out = _rnd(F.relu(_bn(out, sd, prefix + "bn1.", training)))
    out = _conv(out, sd[prefix + "conv2.weight"], padding=1)
    out = _bn(out, sd, prefix + "bn2.", training)
    if prefix + "downsample.0.weight" in sd:
        identity = _conv(x, sd[prefix + "downsample.0.weight"], stride=stride)
        identity = _rnd(_bn(identity, sd, prefix + "downsample.1.", training))
    return _rnd(F.relu(out + identity))


def resnet50_forward(sd, prefix, img, training):
    """ResNet._forward_impl (torchvision resnet.py:266-282) with fc = Linear(2048, latent)
    (util/model_utils.py:139-141).  Returns (latent features, post-ReLU bn1 map).

    The second output is what the reference's bn1 forward hook ends up holding: the hooked tensor
    is overwritten by relu(inplace=True) (models/naive.py:211,282-283)."""
    x = _conv(_rnd(img), sd[prefix + "conv1.weight"], stride=2, padding=3)
    # (TF32-operand mode) the stem activation is rounded where it is produced -- by the folded-BN GEMM epilogue in
    # eval mode, on the fly inside the fused stem tail in training mode (which never writes it) -- so the max pool
    # and the aux branch both see rounded values; ties between values that round to the same TF32 number go to the
    # first element of the window in both implementations
    x = _rnd(F.relu(_bn(x, sd, prefix + "bn1.", training)))
    early = x
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    for li, (_, blocks, stride) in enumerate(RESNET50_STAGES, start=1):
        # block count read from the checkpoint so that shallower Bottleneck stacks (used by the
        # well-conditioned gradient tests) run through the same code; 3/4/6/3 for ResNet-50
        blocks = len({k[len(prefix):].split(".")[1] for k in sd if k.startswith("%slayer%d." % (prefix, li))})
        block = _bottleneck if ("%slayer%d.0.conv3.weight" % (prefix, li)) in sd else _basic_block
        for b in range(blocks):
            x = block(x, sd, "%slayer%d.%d." % (prefix, li, b), stride if b == 0 else 1, training)
    x = _rnd(torch.flatten(F.adaptive_avg_pool2d(x, 1), 1))
    return _rnd(_linear(x, sd[prefix + "fc.weight"], sd[prefix + "fc.bias"])), early


def aux_forward(early, w, b, rounded=True):
    """Conv2d(64,1,1) + MaxPool2d(2) + Flatten (models/naive.py:225-229, time_sensitive.py:379-383)."""
    # fp32 FMA in the CUDA path as well (a 64-term dot product per pixel, no tensor cores); the result is rounded
    # because it becomes a GEMM operand of the head (with use_depth the product with the depth features is)
    a = torch.flatten(F.max_pool2d(F.conv2d(early, w, b), 2), 1)
    return _rnd(a) if rounded else a


def depth_forward(depth, weight, bias, n_pool=2):
    """AvgPool2d(2) x n_pool + InstanceNorm2d(1, affine=True) + Flatten (models/naive.py:233-240: for the bn1 hook,
    n_pool = log4(224^2 / 3136) = 2).  depth: (B,1,H,W)."""
    d = depth
    for _ in range(n_pool):
        d = F.avg_pool2d(d, 2)
    return torch.flatten(F.instance_norm(d, weight=weight, bias=bias, eps=1e-5), 1)


def lstm_forward(x, sd, prefix, state=None, ops=None):
    """Single-layer nn.LSTM, seq-major input (S, N, F), gates (i, f, g, o)
    (models/time_sensitive.py:126-131,418; torch.nn.LSTM definition)."""
    w_ih, w_hh = sd[prefix + "weight_ih_l0"], sd[prefix + "weight_hh_l0"]
    b_ih, b_hh = sd[prefix + "bias_ih_l0"], sd[prefix + "bias_hh_l0"]
    S, N, _ = x.shape
    H = w_hh.shape[1]
    lin, rnd = ops if ops is not None else (_linear, _rnd)
    # sequences run their recurrence in the persistent kernel (fp32 recurrent product); single steps outside the fused
    # head keep the TF32 recurrent GEMM
    rec = _RecurrentFP32.apply if (lin is _linear and _TF32[0] and S > 1) else lin
    if state is None:
        h = x.new_zeros(N, H)
        c = x.new_zeros(N, H)
    else:
        h, c = state[0][0], state[1][0]
    outs = []
    for t in range(S):
        gates = lin(x[t], w_ih, b_ih) + rec(h, w_hh, b_hh)
        i, f, g, o = gates.chunk(4, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h = rnd(torch.sigmoid(o) * torch.tanh(c))
        outs.append(h)
    return torch.stack(outs), (h.unsqueeze(0), c.unsqueeze(0))


# --------------------------------------------------------------------------------------------
# the four estimators
# --------------------------------------------------------------------------------------------
def naive_object_forward(sd, img, x0bar, training, n_fc, use_proprio=True, depth=None):
    """NaiveObjectStateEstimator.forward (models/naive.py:298-352); ReLU after EVERY fc incl. the
    last one (quirk Q2, models/naive.py:343-345)."""
    feats, early = resnet50_forward(sd, "feature_net.module.", img, training)
    lin, rnd = _head_ops(fused=(not training) and img.shape[0] <= _FUSED_HEAD_ROWS)
    aux = aux_forward(early, sd["aux_nets.0.module.0.weight"], sd["aux_nets.0.module.0.bias"], rounded=depth is None)
    if depth is not None:      # use_depth: aux features gated by the normalised, pooled depth map (:324-330)
        aux = _rnd(aux * depth_forward(depth, sd["depth_nets.0.module.2.weight"], sd["depth_nets.0.module.2.bias"]))
    out = torch.cat((feats, aux), dim=-1).view(img.shape[0], -1)
    if use_proprio:
        out = torch.cat((out, rnd(x0bar)), dim=-1)
    for i in range(n_fc):
        out = _relu_head(lin(out, sd["fc%d.module.weight" % i], sd["fc%d.module.bias" % i]))
        if i < n_fc - 1:
            out = rnd(out)
    return out


def naive_eef_forward(sd, img, x0bar, training, n_pre, n_post):
    """NaiveEndEffectorStateEstimator.forward (models/naive.py:68-112): no aux branch."""
    feats, _ = resnet50_forward(sd, "feature_net.", img, training)
    lin, rnd = _head_ops(fused=(not training) and img.shape[0] <= _FUSED_HEAD_ROWS)
    pre = feats
    for i in range(n_pre):
        pre = _relu_head(lin(pre, sd["pre_fc%d.weight" % i], sd["pre_fc%d.bias" % i]))
        if i < n_pre - 1:
            pre = rnd(pre)
    post = torch.cat([feats, rnd(pre - x0bar)], dim=1)
    for i in range(n_post):
        post = _relu_head(lin(post, sd["post_fc%d.weight" % i], sd["post_fc%d.bias" % i]))
        if i < n_post - 1:
            post = rnd(post)
    return pre, post


def tdo_forward(sd, img, x0bar, training, state=None, use_proprio=True, depth=None):
    """TemporallyDependentObjectStateEstimator.forward (models/time_sensitive.py:453-517); head is
    Linear(H, H//4) -> Linear(H//4, 7) with no nonlinearity (quirk Q6, :420-423).
    `state` = (h, c) each (1, N, H) for rollout mode; returns (out, new_state)."""
    S, N = img.shape[0], img.shape[1]
    feats, early = resnet50_forward(sd, "feature_net.module.", img.reshape(S * N, *img.shape[2:]), training)
    ops = _head_ops(fused=(not training) and S == 1 and N <= _FUSED_HEAD_ROWS)
    lin, rnd = ops
    aux = aux_forward(early, sd["aux_nets.0.module.0.weight"], sd["aux_nets.0.module.0.bias"], rounded=depth is None)
    if depth is not None:
        aux = _rnd(aux * depth_forward(depth.reshape(S * N, *depth.shape[2:]), sd["depth_nets.0.module.2.weight"],
                                       sd["depth_nets.0.module.2.bias"]))
    f = torch.cat((feats, aux), dim=-1).view(S, N, -1)
    if use_proprio:
        f = torch.cat((f, rnd(x0bar)), dim=-1)
    h, new_state = lstm_forward(f, sd, "rnn.module.", state, ops)
    out = lin(rnd(lin(h, sd["fc.module.0.weight"], sd["fc.module.0.bias"])),
              sd["fc.module.1.weight"], sd["fc.module.1.bias"])
    return out, new_state


def tdo_v2_forward(sd, img, x0bar, training, state=None):
    """TemporallyDependentObjectStateEstimatorV2.forward (models/time_sensitive.py:714-786): one LSTM per sensor
    modality -- image features (latent + aux) and the 7-D proprioceptive measurement -- whose hidden states are
    concatenated (:776) and fed to Linear(H, H//4) -> Linear(H//4, 7) (:683-686, no nonlinearity).
    `state` = ((h_img, c_img), (h_pro, c_pro)) for rollout mode; returns (out, new_state)."""
    S, N = img.shape[0], img.shape[1]
    feats, early = resnet50_forward(sd, "feature_net.module.", img.reshape(S * N, *img.shape[2:]), training)
    aux = aux_forward(early, sd["aux_nets.0.module.0.weight"], sd["aux_nets.0.module.0.bias"])
    f = torch.cat((feats, aux), dim=-1).view(S, N, -1)
    h_img, st_img = lstm_forward(f, sd, "img_rnn.module.", None if state is None else state[0])
    h_pro, st_pro = lstm_forward(_rnd(x0bar), sd, "proprio_rnn.module.", None if state is None else state[1])
    h = torch.cat((h_img, h_pro), dim=-1)
    out = _linear(_rnd(_linear(h, sd["fc.module.0.weight"], sd["fc.module.0.bias"])),
                  sd["fc.module.1.weight"], sd["fc.module.1.bias"])
    return out, (st_img, st_pro)


def td_forward(sd, img, x0bar, training, aux_w, aux_b, state=None):
    """TemporallyDependentStateEstimator.forward (models/time_sensitive.py:165-254).  The aux conv is
    NOT part of the state_dict (plain python list, quirk Q4, :77-78,102-115) so it is passed in.
    `state` = ((h_pre, c_pre), (h_post, c_post)) for rollout mode."""
    S, N = img.shape[0], img.shape[1]
    feats, early = resnet50_forward(sd, "feature_net.", img.reshape(S * N, *img.shape[2:]), training)
    aux = aux_forward(early, aux_w, aux_b)
    ops = _head_ops(fused=(not training) and S == 1 and N <= _FUSED_HEAD_ROWS)
    lin, rnd = ops
    f = torch.cat((feats, aux), dim=-1).view(S, N, -1)
    h_pre, st_pre = lstm_forward(f, sd, "pre_measurement_rnn.", None if state is None else state[0], ops)
    pre_out = lin(h_pre, sd["pre_measurement_fc.weight"], sd["pre_measurement_fc.bias"])
    post_in = torch.cat([f, rnd(pre_out - x0bar)], dim=-1)
    h_post, st_post = lstm_forward(post_in, sd, "post_measurement_rnn.", None if state is None else state[1], ops)
    post_out = lin(h_post, sd["post_measurement_fc.weight"], sd["post_measurement_fc.bias"])
    return pre_out, post_out, (st_pre, st_post)


# --------------------------------------------------------------------------------------------
# loss (models/losses.py:47-128)
# --------------------------------------------------------------------------------------------
def pose_loss(prediction, truth, distance_metric="l2", scale_factor=1.0, alpha=1.0, epsilon=1e-4, mode="pose"):
    """PoseDistanceLoss.forward for modes 'position' / 'pose': a SUM over samples (quirk Q7)."""
    pp, po = torch.split(prediction, (3, 4), dim=-1)
    tp, to = torch.split(truth, (3, 4), dim=-1)
    po = po / torch.sqrt(torch.sum(po.pow(2), dim=-1, keepdim=True))          # losses.py:68-69
    d = pp - tp
    l2 = torch.sum(torch.sqrt(torch.sum(d.pow(2), dim=-1) + epsilon))           # losses.py:74-75
    l1 = torch.sum(torch.abs(d))                                                # losses.py:80
    linf = torch.sum(torch.max(torch.abs(d), dim=-1)[0])                        # losses.py:82
    if distance_metric == "l2":
        pos = l2
    elif distance_metric == "l1":
        pos = l1
    elif distance_metric == "linf":
        pos = linf
    elif distance_metric == "combined":
        pos = l2 + l1 + linf                                                    # losses.py:85-92
    else:
        raise ValueError(distance_metric)
    if mode == "pose":
        ip = torch.sum(po * to, dim=-1)
        ori = torch.sum(1 - ip.pow(2)) + torch.sum(torch.clamp(-po[..., -1], min=0))   # losses.py:117-122
    elif mode == "position":
        ori = 0
    else:
        raise ValueError(mode)
    return scale_factor * (pos + alpha * ori)                                   # losses.py:128


def pose_val_metrics(prediction, truth, distance_metric="l2", epsilon=1e-4):
    """'val' mode (models/losses.py:95-113): (position distance, summed |angle| in radians).
    robosuite v1.0 semantics: angle = 2*acos(w) of q_pred * conj(q_true), 0 when sqrt(1-w^2) ~ 0,
    wrapped into [-pi, pi] before abs().  For unit quaternions w = <q_pred_normalised, q_true>."""
    pos = pose_loss(prediction, truth, distance_metric, 1.0, 0.0, epsilon, "position")
    po = prediction[..., 3:] / torch.sqrt(torch.sum(prediction[..., 3:].pow(2), dim=-1, keepdim=True))
    to = truth[..., 3:]
    total = 0.0
    for p, t in zip(po.reshape(-1, 4).tolist(), to.reshape(-1, 4).tolist()):
        w = max(-1.0, min(1.0, sum(a * b for a, b in zip(p, t)) / max(sum(b * b for b in t), 1e-300)))
        den = math.sqrt(max(1.0 - w * w, 0.0))
        angle = 0.0 if math.isclose(den, 0.0) else 2.0 * math.acos(w)
        if angle > math.pi:
            angle -= 2 * math.pi
        total += abs(angle)
    return float(pos), total


# --------------------------------------------------------------------------------------------
# Adam (torch.optim.Adam defaults; scripts/train_model.py:228, util/learn_utils.py:179)
# --------------------------------------------------------------------------------------------
def adam_step(params, grads, exp_avg, exp_avg_sq, step, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8):
    """In-place Adam update of lists of tensors; `step` is the 1-based step count."""
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    for p, g, m, v in zip(params, grads, exp_avg, exp_avg_sq):
        if g is None:
            continue
        m.lerp_(g, 1 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
        p.addcdiv_(m, denom, value=-lr / bc1)


# --------------------------------------------------------------------------------------------
# convenience: one training step of any estimator on CPU
# --------------------------------------------------------------------------------------------
class OracleEstimator:
    """Holds a reference-layout state_dict and runs forward / loss / backward / Adam on CPU."""

    def __init__(self, kind, state_dict, extra=None):
        assert kind in ("no", "n", "td", "tdo", "tdo_v2")
        self.kind = kind
        self.sd = {k: v.detach().clone() for k, v in state_dict.items()}
        self.extra = {k: v.detach().clone() for k, v in (extra or {}).items()}  # td: aux conv (frozen)
        self.param_names = [k for k, v in self.sd.items()
                            if v.dtype.is_floating_point and "running_" not in k]
        self.state = None
        self.adam_m = None
        self.adam_v = None
        self.adam_t = 0

    def _count(self, pattern):
        return len([k for k in self.sd if k.startswith(pattern) and k.endswith("weight")])

    def forward(self, img, x0bar, training=True, rollout=False, depth=None):
        sd = self.sd
        if self.kind == "no":
            return naive_object_forward(sd, img, x0bar, training, self._count("fc"), depth=depth)
        if self.kind == "n":
            return naive_eef_forward(sd, img, x0bar, training, self._count("pre_fc"), self._count("post_fc"))
        if self.kind == "tdo":
            out, st = tdo_forward(sd, img, x0bar, training, self.state if rollout else None, depth=depth)
            if rollout:
                self.state = st
            return out
        if self.kind == "tdo_v2":
            out, st = tdo_v2_forward(sd, img, x0bar, training, self.state if rollout else None)
            if rollout:
                self.state = st
            return out
        pre, post, st = td_forward(sd, img, x0bar, training, self.extra["aux_w"], self.extra["aux_b"],
                                   self.state if rollout else None)
        if rollout:
            self.state = st
        return pre, post

    def reset_state(self, n):
        h = self.sd["rnn.module.weight_hh_l0"].shape[1] if self.kind == "tdo" else None
        if self.kind == "tdo":
            self.state = (torch.zeros(1, n, h), torch.zeros(1, n, h))
        elif self.kind == "tdo_v2":
            hi = self.sd["img_rnn.module.weight_hh_l0"].shape[1]
            hp = self.sd["proprio_rnn.module.weight_hh_l0"].shape[1]
            self.state = ((torch.zeros(1, n, hi), torch.zeros(1, n, hi)), (torch.zeros(1, n, hp), torch.zeros(1, n, hp)))
        elif self.kind == "td":
            hp = self.sd["pre_measurement_rnn.weight_hh_l0"].shape[1]
            hq = self.sd["post_measurement_rnn.weight_hh_l0"].shape[1]
            self.state = ((torch.zeros(1, n, hp), torch.zeros(1, n, hp)),
                          (torch.zeros(1, n, hq), torch.zeros(1, n, hq)))

    def loss_and_grads(self, img, x0bar, target, loss_kwargs, which=-1):
        """Returns (outputs, loss, {param name: grad}).  `which` picks the output the loss applies to
        for two-headed models (util/learn_utils.py:160-176 trains on the object / post output)."""
        for k in self.param_names:
            self.sd[k].requires_grad_(True)
            self.sd[k].grad = None
        out = self.forward(img, x0bar, training=True)
        pred = out[which] if isinstance(out, tuple) else out
        loss = pose_loss(pred, target, **loss_kwargs)
        loss.backward()
        grads = {k: (None if self.sd[k].grad is None else self.sd[k].grad.detach().clone())
                 for k in self.param_names}
        for k in self.param_names:
            self.sd[k].requires_grad_(False)
        outs = tuple(o.detach() for o in out) if isinstance(out, tuple) else out.detach()
        return outs, loss.detach(), grads

    def train_step(self, img, x0bar, target, loss_kwargs, lr=1e-3, which=-1):
        outs, loss, grads = self.loss_and_grads(img, x0bar, target, loss_kwargs, which)
        if self.adam_m is None:
            self.adam_m = {k: torch.zeros_like(self.sd[k]) for k in self.param_names}
            self.adam_v = {k: torch.zeros_like(self.sd[k]) for k in self.param_names}
        self.adam_t += 1
        names = [k for k in self.param_names if grads[k] is not None]
        with torch.no_grad():
            adam_step([self.sd[k] for k in names], [grads[k] for k in names], [self.adam_m[k] for k in names],
                      [self.adam_v[k] for k in names], self.adam_t, lr=lr)
        return outs, loss


def synthetic_batch(kind, n, s=None, seed=1, hw=224):
    """Synthetic inputs of SURVEY 8(d): img ~ N(0,1); positions U(-0.5,0.5)^3; unit quaternions with
    w >= 0 (util/data_utils.py:207-211).  Returns (img, x0bar, target) as CPU fp32 tensors."""
    g = torch.Generator().manual_seed(seed)
    lead = (n,) if kind in ("no", "n") else (s, n)
    img = torch.randn(*lead, 3, hw, hw, generator=g)

    def pose():
        pos = torch.rand(*lead, 3, generator=g) - 0.5
        q = torch.randn(*lead, 4, generator=g)
        q = q / q.norm(dim=-1, keepdim=True)
        q[..., 3] = q[..., 3].abs()
        return torch.cat([pos, q], dim=-1)

    return img, pose(), pose()

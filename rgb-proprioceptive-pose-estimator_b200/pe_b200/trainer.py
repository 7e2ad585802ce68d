"""Fused training step: forward -> pose loss (+gradient) -> backward -> [NCCL all-reduce] -> fused Adam.

This is the body of the reference's hot loop (util/learn_utils.py:151-184: zero_grad / forward / loss /
backward / optimizer.step) with autograd and the per-parameter optimizer loop taken out of the way:
parameters, gradients and Adam moments live in flat 128-byte-aligned arenas, gradients are written in
place by the backward kernels, and one 128-bit streaming kernel applies the update.  Semantics are those
of `torch.optim.Adam(model.parameters(), lr)` (scripts/train_model.py:228).

Multi-GPU: one process per GPU, batch (naive) or episodes (sequence models) sharded across ranks,
gradients summed (not averaged: the loss is a sum over samples, models/losses.py:75,128) with NCCL in
backward-ordered buckets on a side stream; BatchNorm statistics stay per GPU.
"""
import os

import torch

from . import native
from .engine import LinearOp, LSTMOp

_METRIC_ID = {"l1": 0, "l2": 1, "linf": 2, "combined": 3}
_ALIGN = 32  # floats (128 B)


def get_core(model):
    """The estimator core of a mirrored model (built lazily, shared with model.forward)."""
    from . import estimators as est
    core = getattr(model, "_core", None)
    if core is None:
        name = type(model).__name__
        cls = {"NaiveObjectStateEstimator": est.NaiveObjectCore,
               "NaiveEndEffectorStateEstimator": est.NaiveEefCore,
               "TemporallyDependentObjectStateEstimator": est.TDOCore,
               "TemporallyDependentObjectStateEstimatorV2": est.TDOV2Core,
               "TemporallyDependentStateEstimator": est.TDCore}[name]
        core = cls(model)
        object.__setattr__(model, "_core", core)
    return core


def engine_of(core, dev):
    """The TrunkEngine a core drives (created on first use by the core's own accessor)."""
    if hasattr(core, "_engine"):
        try:
            return core._engine()
        except TypeError:
            return core._engine(dev)          # TDCore places its unregistered aux conv on the device first
    net = core.net if hasattr(core, "net") else core.m.feature_net
    return getattr(net, "module", net).pe_engine()


def invalidate_core(core):
    """Drop every packed-weight cache (parameters were updated by a raw kernel, not a torch op)."""
    for v in vars(core).values():
        items = v if isinstance(v, (list, tuple)) else [v]
        for it in items:
            if isinstance(it, (LinearOp, LSTMOp)):
                it._ver = None
    m = core.m if hasattr(core, "m") else None
    net = getattr(core, "net", None)
    if m is not None:
        fn = m.feature_net
        net = getattr(fn, "module", fn)
    if net is not None and getattr(net, "_pe_engine", None) is not None:
        net._pe_engine.invalidate()


class FusedTrainer:
    POLL_EVERY = 64       # steps between host reads of the device error flag (a 4-byte D2H copy + sync)

    def __init__(self, model, distance_metric="l2", alpha=1.0, epsilon=1e-4, scale_factor=1.0, mode="pose",
                 lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, optimizer="adam", momentum=0.0,
                 process_group=None, bucket_mb=32, sm_reserve=None):
        if mode not in ("pose", "position"):
            raise ValueError("training loss mode must be 'pose' or 'position'")
        self.model = model
        self.core = get_core(model)
        self.loss_cfg = (_METRIC_ID[distance_metric], 1 if mode == "pose" else 0, float(alpha), float(epsilon),
                         float(scale_factor))
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.optimizer, self.momentum = optimizer, momentum
        self.pg = process_group
        # All-reduce bucket size.  Measured on 8 x B200, config 2 (profiles/ddp_overlap_r02c.txt): 8 MB buckets 34.92
        # ms/step, 32 MB 33.55, one all-reduce after the backward pass 33.41 (no collective at all: 33.92 on another box).
        # Every NCCL kernel in flight takes SMs from the persistent one-CTA-per-SM GEMM grids, whose last tiles then run as a
        # second wave: a few large buckets overlap almost as much and disturb far less than many small ones.
        self.bucket_bytes = int(os.environ.get("PE_B200_BUCKET_MB", bucket_mb)) << 20
        # SMs the persistent GEMM grids leave to the NCCL kernels during the backward pass (multi-GPU only)
        self.sm_reserve = int(os.environ.get("PE_B200_SM_RESERVE", "0")) if sm_reserve is None else sm_reserve
        self.t = 0
        self._flat = None
        self.comm_stream = None
        self.reducer = None
        self._fused_update = False

    # ------------------------------------------------------------------------------------------
    def _flatten(self):
        params = [p for p in self.model.parameters()]
        if not params or not params[0].is_cuda:
            raise native.PeError("FusedTrainer: move the model to a CUDA device first (no CPU fallback)")
        dev = params[0].device
        offs, total = [], 0
        for p in params:
            offs.append(total)
            total += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        pf = torch.zeros(total, device=dev, dtype=torch.float32)
        gf = torch.zeros(total, device=dev, dtype=torch.float32)
        self.views = {}
        with torch.no_grad():
            for p, o in zip(params, offs):
                n = p.numel()
                pf[o:o + n].copy_(p.detach().reshape(-1))
                p.data = pf[o:o + n].view(p.shape)
                gv = gf[o:o + n].view(p.shape)
                self.views[id(p)] = gv
                p.grad = gv
        self.p_flat, self.g_flat = pf, gf
        self.m_flat = torch.zeros_like(pf)
        self.v_flat = torch.zeros_like(pf) if self.optimizer == "adam" else None
        self.param_offsets = {id(p): (o, p.numel()) for p, o in zip(params, offs)}
        # BatchNorm counters -> one int64 arena, one kernel launch per step
        nbts = [b for n, b in self.model.named_buffers() if n.endswith("num_batches_tracked")]
        self._flat = True
        invalidate_core(self.core)
        # strictly forward -> backward here: no per-forward copy of the BN coefficient arenas
        engine_of(self.core, dev).snapshot_bn = False
        self.touched = set()
        self.reducer = None
        if self.pg is not None:
            import torch.distributed as dist
            from .ddp import BucketedAllReduce
            if dist.get_world_size(self.pg) > 1:
                self.comm_stream = torch.cuda.Stream(device=dev)
                self.reducer = BucketedAllReduce(gf, self.bucket_bytes // 4, group=self.pg, stream=self.comm_stream,
                                                 after=self._bucket_update)
        self._fused_update = False

    def grad_of(self, p):
        self.touched.add(id(p))
        return self.views[id(p)]

    def _on_ready(self, params):
        for p in params:
            o, n = self.param_offsets[id(p)]
            self.reducer.ready(o, o + (n + _ALIGN - 1) // _ALIGN * _ALIGN)

    # ------------------------------------------------------------------------------------------
    def step(self, img, self_measurement, targets, depth=None):
        """One optimisation step.  `targets`: a tensor (object-pose models) or a (x0, x1) pair for the
        two-headed models, matching util/learn_utils.py:160-172.  Returns the loss as a 1-element device
        tensor (sum over the local samples).

        Multi-GPU: from the second step on, every all-reduce bucket is followed on the communication stream by the
        optimizer update of exactly that arena range, so reducing and updating the head / layer4 / layer3 buckets
        overlaps the backward pass of the earlier layers; the last (stem-side) bucket's reduce + update is exposed."""
        if self._flat is None:
            self._flatten()
        self._fused_update = self.reducer is not None and self.t > 0
        try:
            loss = self.forward_backward(img, self_measurement, targets, depth)
        finally:
            fused, self._fused_update = self._fused_update, False
        self.apply_update(already_applied=fused)
        return loss

    def _bucket_update(self, lo, hi):
        """Runs on the communication stream behind the all-reduce of g_flat[lo:hi] (BucketedAllReduce.after)."""
        if self._fused_update and hi > lo:
            self._update_range(lo, hi, self.t + 1)

    def _update_range(self, lo, hi, t):
        L, st, P = native.lib(), native.stream_ptr(), native.ptr
        n = hi - lo
        if self.optimizer == "adam":
            native.account("pe_adam_step", 28 * n)          # read p, g, m, v; write p, m, v
            L.pe_adam_step(P(self.p_flat[lo:]), P(self.g_flat[lo:]), P(self.m_flat[lo:]), P(self.v_flat[lo:]), n,
                           self.lr, self.betas[0], self.betas[1], self.eps, self.wd, t, 1.0, st)
        else:
            L.pe_sgd_step(P(self.p_flat[lo:]), P(self.g_flat[lo:]), P(self.m_flat[lo:]) if self.momentum else None, n,
                          self.lr, self.momentum, self.wd, int(t == 1), 1.0, st)

    def forward_backward(self, img, self_measurement, targets, depth=None):
        """forward + loss + backward (+ gradient all-reduce): leaves the summed gradient in self.g_flat."""
        L, st, P = native.lib(), native.stream_ptr(), native.ptr
        if self._flat is None:
            self._flatten()
        model, core = self.model, self.core
        state = None
        if img.dtype == torch.uint8:
            # raw renderer frames (..., 256, 256, 3) uint8 HWC: CenterCrop(224) + /255 + Normalize on the device
            # (util/data_utils.py:48-54) -- the frames cross PCIe at a third of the bytes of preprocessed fp32 tensors
            if getattr(self, "_pre", None) is None:
                from .preprocess import FramePreprocessor
                self._pre = FramePreprocessor(crop=224)
            img = self._pre(img)
        inputs = (img, self_measurement) if depth is None else (img, self_measurement, depth)
        outs, saved, _ = core.forward(inputs, True, True, state)
        if not isinstance(targets, (tuple, list)):
            targets = (targets,)
        if len(targets) != len(outs):
            raise ValueError("expected %d target tensor(s), got %d" % (len(outs), len(targets)))
        dev = img.device
        metric, mode, alpha, epsilon, scale = self.loss_cfg
        losses = torch.empty(len(outs), device=dev, dtype=torch.float32)
        douts = []
        for i, (o, t) in enumerate(zip(outs, targets)):
            n = o.numel() // 7
            o2 = o.reshape(n, 7)
            t2 = t.reshape(n, 7)
            if t2.stride(1) != 1:
                t2 = t2.contiguous()
            d = torch.empty(n, 7, device=dev, dtype=torch.float32)
            L.pe_pose_loss(P(o2), o2.stride(0), P(t2), t2.stride(0), n, metric, mode, alpha, epsilon, scale,
                           P(losses[i:]), P(d), 7, None, st)
            douts.append(d)
        red = self.reducer
        if red is not None:
            red.reset()
            if self.t > 0:
                # parameters that never receive a gradient (unused depth nets) are final from the start
                for pid, (o, n) in self.param_offsets.items():
                    if pid not in self.touched:
                        red.ready(o, o + (n + _ALIGN - 1) // _ALIGN * _ALIGN)
            if self.sm_reserve and self.t > 0:
                L.pe_set_sm_reserve(self.sm_reserve)
            try:
                core.backward(saved, tuple(douts), self.grad_of, self._on_ready if self.t > 0 else None)
            finally:
                if self.sm_reserve:
                    L.pe_set_sm_reserve(0)
            red.wait()
        else:
            core.backward(saved, tuple(douts), self.grad_of)
        self._pending_loss = losses.sum() if len(outs) > 1 else losses
        # a tcgen05 / TMA pipeline timeout anywhere in this step turns the loss into NaN (asynchronously, on the
        # device); every POLL_EVERY steps the sticky flag is also read back and raised as an exception
        L.pe_poison_on_error(P(self._pending_loss), 1, st)
        if self.t % self.POLL_EVERY == self.POLL_EVERY - 1:
            L.check_device()
        return self._pending_loss

    def apply_update(self, already_applied=False):
        """Optimizer update over the whole arena (one streaming kernel), unless the buckets already carried it."""
        self.t += 1
        if not already_applied:
            self._update_range(0, self.p_flat.numel(), self.t)
        invalidate_core(self.core)

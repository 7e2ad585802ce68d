"""Streaming rollout inference (the per-step forward of util/learn_utils.py:366-455) under CUDA Graph capture.

The reference runs one batch-1 forward per simulator step on the CPU with autograd enabled and the LSTM
state carried in python attributes.  Here the eval-mode forward -- folded BatchNorm,
fused conv epilogues, LSTM cell, head -- is captured ONCE into a CUDA graph over static input / state /
output buffers; every step is then two small H2D copies, one graph launch and one 28-byte D2H read.
"""
import torch

from . import native
from .trainer import get_core


class StreamingEstimator:
    def __init__(self, model, batch_size=1, use_graph=True, device=None, raw_hw=None):
        """raw_hw: side length of raw uint8 HWC frames (256 for robosuite's default render) when the caller wants
        to hand over un-preprocessed frames with `step_raw`; crop / scale / normalise then run inside the graph."""
        self.model = model
        self.core = get_core(model)
        self.N = batch_size
        self.dev = device or next(model.parameters()).device
        if self.dev.type != "cuda":
            raise native.PeError("StreamingEstimator needs the model on a CUDA device (no CPU fallback)")
        self.seq = bool(getattr(model, "requires_sequence", False))
        self.use_graph = use_graph
        lead = (1, batch_size) if self.seq else (batch_size,)
        self.img = torch.zeros(*lead, 3, 224, 224, device=self.dev)
        self.x0 = torch.zeros(*lead, 7, device=self.dev)
        self.state = self._zero_state()
        self.raw = None
        if raw_hw is not None:
            from .preprocess import FramePreprocessor
            self.raw = torch.zeros(*lead, raw_hw, raw_hw, 3, device=self.dev, dtype=torch.uint8)
            self.pre = FramePreprocessor(crop=224)
        self.graph = None
        self.outs = None
        model.eval()

    def _zero_state(self):
        m, N, dev = self.model, self.N, self.dev
        name = type(m).__name__
        if name == "TemporallyDependentObjectStateEstimator":
            return (torch.zeros(N, m.hidden_dim, device=dev), torch.zeros(N, m.hidden_dim, device=dev))
        if name == "TemporallyDependentObjectStateEstimatorV2":
            return ((torch.zeros(N, m.img_hidden_dim, device=dev), torch.zeros(N, m.img_hidden_dim, device=dev)),
                    (torch.zeros(N, m.proprio_hidden_dim, device=dev), torch.zeros(N, m.proprio_hidden_dim, device=dev)))
        if name == "TemporallyDependentStateEstimator":
            return ((torch.zeros(N, m.pre_measurement_hidden_dim, device=dev),
                     torch.zeros(N, m.pre_measurement_hidden_dim, device=dev)),
                    (torch.zeros(N, m.post_measurement_hidden_dim, device=dev),
                     torch.zeros(N, m.post_measurement_hidden_dim, device=dev)))
        return None

    def reset(self):
        """model.reset_initial_state(batch_size) of the rollout loop (util/learn_utils.py:342)."""
        def z(t):
            if isinstance(t, tuple):
                for u in t:
                    z(u)
            elif t is not None:
                t.zero_()
        z(self.state)

    def _forward(self):
        L, st, P = native.lib(), native.stream_ptr(), native.ptr
        if self.raw is not None:
            self.pre(self.raw, out=self.img)
        outs, _, new_state = self.core.forward((self.img, self.x0), False, False, self.state)

        def carry(dst, src):
            if isinstance(dst, tuple):
                for d, s in zip(dst, src):
                    carry(d, s)
            elif dst is not None and src.data_ptr() != dst.data_ptr():   # the fused head updates the state in place
                L.pe_copy_cols(P(src), src.stride(0), P(dst), dst.stride(0), dst.shape[0], dst.shape[1], 0, st)
        carry(self.state, new_state)
        for o in outs:
            # a pipeline timeout shows up as NaN poses (captured into the graph, so every replay checks)
            base = o._base if o._base is not None else o
            L.pe_poison_on_error(P(base), base.numel(), st)
        return outs

    def _capture(self):
        # warm every lazily initialised piece (packed weights, folded BN, tensor-map entry point) eagerly
        saved = self._clone_state()
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                self._forward()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self._restore_state(saved)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.outs = self._forward()
        self._restore_state(saved)

    def _clone_state(self):
        def c(t):
            return tuple(c(u) for u in t) if isinstance(t, tuple) else (None if t is None else t.clone())
        return c(self.state)

    def _restore_state(self, saved):
        def r(dst, src):
            if isinstance(dst, tuple):
                for d, s in zip(dst, src):
                    r(d, s)
            elif dst is not None:
                dst.copy_(src)
        r(self.state, saved)

    @torch.no_grad()
    def step_raw(self, frames_u8, self_measurement):
        """Raw uint8 HWC frames (N, raw_hw, raw_hw, 3) straight from the renderer + measurement (N,7)."""
        if self.raw is None:
            raise native.PeError("construct StreamingEstimator(..., raw_hw=256) to feed raw frames")
        self.raw.copy_(frames_u8.reshape(self.raw.shape), non_blocking=True)
        self.x0.copy_(self_measurement.reshape(self.x0.shape), non_blocking=True)
        return self._run()

    @torch.no_grad()
    def step(self, img, self_measurement):
        """img (N,3,224,224) [or (1,N,...)], self_measurement (N,7): host or device tensors.  Returns the
        pose estimate(s) as device tensors that stay valid until the next step."""
        if self.raw is not None:
            raise native.PeError("this estimator was built for raw frames: call step_raw")
        self.img.copy_(img.reshape(self.img.shape), non_blocking=True)
        self.x0.copy_(self_measurement.reshape(self.x0.shape), non_blocking=True)
        return self._run()

    def _run(self):
        if not self.use_graph:
            outs = self._forward()
        else:
            if self.graph is None:
                self._capture()
            self.graph.replay()
            outs = self.outs
        return outs if len(outs) > 1 else outs[0]


class PipelinedEstimator:
    """Large-batch rollout step with the host -> device copy of the frames overlapped with the trunk: the batch is cut
    into chunks of `chunk` episodes, each a CUDA-graph `StreamingEstimator` with its own input / LSTM-state / output
    buffers; chunk i + 1 is copied on a side stream while chunk i computes.  Episodes are independent (SURVEY 8e:
    rollout = replicas), so the result is the one-shot step's, row for row.  Pays from ~512 frames per step on, where
    the 150 KB .. 600 KB per frame of PCIe traffic is a third of the step (bench.py rollout sweep)."""

    def __init__(self, model, batch_size, chunk=256, device=None, raw_hw=None):
        if batch_size % chunk:
            raise native.PeError("PipelinedEstimator: batch_size %d is not a multiple of chunk %d" % (batch_size, chunk))
        self.N, self.chunk = batch_size, chunk
        self.parts = [StreamingEstimator(model, chunk, use_graph=True, device=device, raw_hw=raw_hw)
                      for _ in range(batch_size // chunk)]
        self.dev = self.parts[0].dev
        self.seq = self.parts[0].seq
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.landed = [torch.cuda.Event() for _ in self.parts]
        self.consumed = [torch.cuda.Event() for _ in self.parts]
        self.outs = None

    def reset(self):
        for p in self.parts:
            p.reset()

    def _step(self, frames, self_measurement, raw):
        c = self.chunk
        cur = torch.cuda.current_stream(self.dev)
        frames = frames.reshape(self.N, *frames.shape[-3:])
        x0 = self_measurement.reshape(self.N, 7)
        self.copy_stream.wait_stream(cur)          # device-resident inputs may still be in flight on the caller's stream
        with torch.cuda.stream(self.copy_stream):
            for i, p in enumerate(self.parts):
                self.copy_stream.wait_event(self.consumed[i])      # the previous step has read this chunk's buffers
                dst = p.raw if raw else p.img
                dst.copy_(frames[i * c:(i + 1) * c].reshape(dst.shape), non_blocking=True)
                p.x0.copy_(x0[i * c:(i + 1) * c].reshape(p.x0.shape), non_blocking=True)
                self.landed[i].record(self.copy_stream)
        L, st, P = native.lib(), native.stream_ptr(), native.ptr
        for i, p in enumerate(self.parts):
            cur.wait_event(self.landed[i])
            o = p._run()
            self.consumed[i].record(cur)
            o = o if isinstance(o, tuple) else (o,)
            if self.outs is None:
                lead = (1, self.N) if self.seq else (self.N,)
                self.outs = tuple(torch.empty(*lead, 8, device=self.dev) for _ in o)
            for dst, src in zip(self.outs, o):
                d2 = dst.view(self.N, 8)[i * c:(i + 1) * c]
                s2 = src.reshape(c, src.shape[-1])
                L.pe_copy_cols(P(s2), s2.stride(0), P(d2), 8, c, 7, 0, st)
        outs = tuple(t[..., :7] for t in self.outs)
        return outs if len(outs) > 1 else outs[0]

    @torch.no_grad()
    def step(self, img, self_measurement):
        if self.parts[0].raw is not None:
            raise native.PeError("this estimator was built for raw frames: call step_raw")
        return self._step(img, self_measurement, False)

    @torch.no_grad()
    def step_raw(self, frames_u8, self_measurement):
        if self.parts[0].raw is None:
            raise native.PeError("construct PipelinedEstimator(..., raw_hw=256) to feed raw frames")
        return self._step(frames_u8, self_measurement, True)

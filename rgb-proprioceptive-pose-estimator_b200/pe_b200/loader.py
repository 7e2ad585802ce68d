"""Host -> device input pipeline for the training loop: the next batch is copied from pinned host memory on a side
stream while the current step computes (the reference's loop copies nothing: it trains on the CPU; on a GPU the copy of
154 MB of frames per 256-frame step would otherwise sit on the critical path, util/learn_utils.py:146-160)."""
import torch

from . import native


class DevicePrefetcher:
    """Wraps an iterable of batches (tuples / lists of CPU tensors, ideally pinned) and yields the same structure on
    the device, always one batch ahead.  Two device buffers alternate, so a batch stays valid until the one after
    the next is requested."""

    def __init__(self, batches, device, depth=2):
        if device.type != "cuda":
            raise native.PeError("DevicePrefetcher needs a CUDA device (no CPU fallback)")
        self.it = iter(batches)
        self.dev = device
        self.stream = torch.cuda.Stream(device=device)
        self.slots = [None] * depth
        self.events = [None] * depth
        self.k = 0
        self._next = self._issue()

    def _ensure(self, obj, slot, path):
        """Device buffers of a slot, allocated on the CONSUMER's stream: they come from (and return to) the caching
        allocator's main pool, so a new prefetcher reuses the previous one's memory instead of paying a cudaMalloc
        per side stream."""
        if isinstance(obj, (tuple, list)):
            for i, o in enumerate(obj):
                self._ensure(o, slot, path + (i,))
        elif torch.is_tensor(obj):
            bufs = self.slots[slot]
            buf = bufs.get(path)
            if buf is None or buf.shape != obj.shape or buf.dtype != obj.dtype:
                bufs[path] = torch.empty(obj.shape, dtype=obj.dtype, device=self.dev)

    def _copy(self, obj, slot, path):
        if isinstance(obj, (tuple, list)):
            return type(obj)(self._copy(o, slot, path + (i,)) for i, o in enumerate(obj))
        if not torch.is_tensor(obj):
            return obj
        buf = self.slots[slot][path]
        buf.copy_(obj, non_blocking=True)
        buf.record_stream(self.stream)      # not handed back to the allocator while the copy may still be running
        return buf

    def _issue(self):
        try:
            batch = next(self.it)
        except StopIteration:
            return None
        slot = self.k % len(self.slots)
        self.k += 1
        if self.slots[slot] is None:
            self.slots[slot] = {}
        cur = torch.cuda.current_stream(self.dev)
        self._ensure(batch, slot, ())
        with torch.cuda.stream(self.stream):
            # the buffers of this slot may still be read by the step that used them two batches ago
            self.stream.wait_stream(cur)
            out = self._copy(batch, slot, ())
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return out, ev

    def __iter__(self):
        return self

    def __next__(self):
        if self._next is None:
            raise StopIteration
        out, ev = self._next
        torch.cuda.current_stream(self.dev).wait_event(ev)
        self._next = self._issue()          # start the copy of the following batch before this one is consumed
        return out

"""Multi-GPU plumbing: one process per GPU, samples (naive models) or episodes (sequence models) sharded
across ranks, gradients SUMMED with NCCL in backward-ordered buckets on a side stream.

The reference's own nn.DataParallel wrappers are single-process pass-throughs on <= 1 device and wrong
on more; the only exchange step of the path is the gradient all-reduce below.
"""
import os

import torch
import torch.distributed as dist

# measurement probe (profiles/ddp_overlap_*): PE_B200_SKIP_ALLREDUCE=1 drops the collective itself (the per-bucket
# optimizer update still runs on the communication stream), so step time with / without it = the EXPOSED share of the
# gradient all-reduce.  The gradients are then per-rank: never set it outside that measurement.
_SKIP_ALLREDUCE = os.environ.get("PE_B200_SKIP_ALLREDUCE", "0") == "1"


class _Done:
    def wait(self):
        return True


def shard_range(n, rank, world):
    """Contiguous [lo, hi) share of n independent units (frames / episodes) for this rank."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class BucketedAllReduce:
    """All-reduce(sum) of a flat gradient arena in buckets that are launched as soon as a contiguous
    tail of the arena is final (backward produces gradients from the end of the arena to its start)."""

    def __init__(self, flat, bucket_elems, group=None, async_op=True, stream=None, after=None):
        """after(lo, hi): optional callback enqueued on `stream` right behind each bucket's all-reduce (the fused
        trainer applies the optimizer update of that arena range there, so the update of the early buckets overlaps
        the rest of the backward pass and only the last, small bucket's reduce + update is exposed)."""
        self.flat = flat
        self.bucket = int(bucket_elems)
        self.group = group
        self.async_op = async_op
        self.stream = stream
        self.after = after
        self.reset()

    def reset(self):
        self.hi = self.flat.numel()     # everything in [hi, end) has been launched
        self.lo = self.hi               # everything in [lo, hi) is final but not launched yet
        self.final = []                 # announced ranges not yet contiguous with [lo, hi)
        self.works = []
        self.launched = 0

    def _launch(self, lo, hi):
        if hi <= lo:
            return
        view = self.flat[lo:hi]
        if self.stream is not None:
            ev = torch.cuda.current_stream().record_event()
            self.stream.wait_event(ev)
            with torch.cuda.stream(self.stream):
                w = _Done() if _SKIP_ALLREDUCE else dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group,
                                                                   async_op=True)
                if self.after is not None:
                    w.wait()                  # stream-level dependency on the collective, the host does not block
                    self.after(lo, hi)
            self.works.append(w)
        else:
            w = dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=self.async_op)
            if self.async_op:
                self.works.append(w)
        self.launched += 1

    def ready(self, lo, hi):
        """Gradients in [lo, hi) are final.  Ranges may arrive in any order; a bucket is launched only once
        a gap-free run of final ranges reaches back from the already-launched tail of the arena."""
        self.final.append((lo, hi))
        grew = True
        while grew:
            grew = False
            for iv in list(self.final):
                if iv[1] >= self.lo:          # touches (or overlaps) the pending region
                    self.lo = min(self.lo, iv[0])
                    self.final.remove(iv)
                    grew = True
        while self.hi - self.lo >= self.bucket:
            cut = self.hi - self.bucket
            self._launch(cut, self.hi)
            self.hi = cut

    def wait(self):
        self._launch(0, self.hi)        # whatever is left, including never-announced ranges
        self.hi = self.lo = 0
        self.final = []
        for w in self.works:
            w.wait()
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)
        self.works = []

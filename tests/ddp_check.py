"""Multi-GPU check (run under torchrun, NCCL): the bucketed, overlapped gradient all-reduce of the
FusedTrainer equals the plain sum of the ranks' local gradients, and every rank ends the step with
identical parameters.  Each rank trains on its own shard of a global synthetic batch."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rgb-proprioceptive-pose-estimator_b200"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import model_checks as mc  # noqa: E402
from oracle import pose_oracle as po  # noqa: E402
from pe_b200.ddp import shard_range  # noqa: E402
from pe_b200.trainer import FusedTrainer  # noqa: E402


def oracle_check(kind, trainer, model, img, x0, tgt, lk, rank, dev):
    """SURVEY 8e's multi-GPU oracle: the reference algorithm run ONCE PER SHARD with the gradients summed (BatchNorm
    statistics stay per GPU, the loss is a sum over samples).  Every rank runs the CPU oracle on its own shard -- at the
    CUDA path's operand precision and teacher-forced with the conv outputs of this very step, like
    model_checks.check_forced -- the per-shard reference gradients are summed across ranks, and the result is compared
    per parameter with the gradient arena the bucketed NCCL all-reduce left behind.  Step 0 reduces in one call,
    step 1 in backward-ordered buckets overlapped with the backward pass."""
    from pe_b200 import engine
    good = True
    for step in range(2):
        orc = mc.oracle_for(kind, model)
        orc.sd = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in orc.sd.items()}
        orc.extra = {k: v.double() for k, v in orc.extra.items()}
        engine.CAPTURE_CONV_OUTPUTS[0] = []
        try:
            trainer.forward_backward(img, x0, tgt)
            acts = engine.CAPTURE_CONV_OUTPUTS[0]
        finally:
            engine.CAPTURE_CONV_OUTPUTS[0] = None
        ys = [t.t.view(t.B, t.H, t.W, t.C).permute(0, 3, 1, 2).double().cpu() for t in acts]
        del acts
        with po.tf32_operands(True), po.forced_conv_outputs(ys):
            _, _, grads_ref = mc.oracle_loss_and_grads(orc, kind, img.double().cpu(), x0.double().cpu(),
                                                       tgt.double().cpu(), lk)
        ref = torch.zeros(trainer.g_flat.numel(), device=dev, dtype=torch.float64)
        named = dict(model.named_parameters())
        spans = {}
        for k, g in grads_ref.items():
            o, n = trainer.param_offsets[id(named[k])]
            spans[k] = (o, n)
            if g is not None:
                ref[o:o + n] = g.reshape(-1).to(dev)
        dist.all_reduce(ref)                                  # sum of the per-shard oracle gradients
        errs = []
        for k, (o, n) in spans.items():
            if grads_ref[k] is None:
                continue
            r = ref[o:o + n]
            errs.append((float((trainer.g_flat[o:o + n].double() - r).norm() / r.norm().clamp_min(1e-30)), k))
        errs.sort(reverse=True)
        buckets = trainer.reducer.launched if trainer.reducer is not None else 0
        if rank == 0:
            print("oracle step %d: worst per-parameter gradient error vs the summed per-shard oracle %.3e (%s), median "
                  "%.3e, %d all-reduce launches" % (step, errs[0][0], errs[0][1], errs[len(errs) // 2][0], buckets),
                  flush=True)
        good = good and errs[0][0] <= 1e-2
        trainer.apply_update()
    return good


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    kind = sys.argv[1] if len(sys.argv) > 1 else "tdo"
    n_global = 2 * world
    img, x0, tgt = po.synthetic_batch(kind, n_global, s=2, seed=1) if kind in ("td", "tdo") else po.synthetic_batch(kind, n_global, seed=1)
    lo, hi = shard_range(n_global, rank, world)
    sl = (slice(None), slice(lo, hi)) if kind in ("td", "tdo") else (slice(lo, hi),)
    img, x0, tgt = img[sl].contiguous().to(dev), x0[sl].contiguous().to(dev), tgt[sl].contiguous().to(dev)
    lk = mc.CONFIGS[kind]["loss"]
    # 4-block trunk: the full-depth random-init net amplifies summation-order noise of the split-K atomics
    # by ~1e5, which would drown the comparison; the plumbing under test is depth independent
    mc.SHALLOW[0] = "--full" not in sys.argv
    a = mc.build_model(kind).to(dev).train()
    b = mc.build_model(kind).to(dev).train()
    c = mc.build_model(kind).to(dev).train()
    tc = FusedTrainer(c, lr=1e-4, **lk)
    ta = FusedTrainer(a, lr=1e-4, process_group=dist.group.WORLD, bucket_mb=8, **lk)
    tb = FusedTrainer(b, lr=1e-4, **lk)
    ok = True
    from pe_b200.trainer import invalidate_core
    if "--oracle" in sys.argv:
        ok = oracle_check(kind, ta, a, img, x0, tgt, lk, rank, dev)
        dist.barrier()
        dist.destroy_process_group()
        return 0 if ok else 1
    for step in range(3):                       # step 0: single all-reduce; steps 1-2: overlapped buckets
        if step > 0:                            # same parameters everywhere: only the reduction is under test
            for t in (tb, tc):
                t.p_flat.copy_(ta.p_flat)
                invalidate_core(t.core)
        ta.forward_backward(img, x0, tgt)
        tb.forward_backward(img, x0, tgt)
        tc.forward_backward(img, x0, tgt)
        noise = float((tb.g_flat - tc.g_flat).abs().max() / tb.g_flat.abs().max())
        want = tb.g_flat.clone()
        dist.all_reduce(want)
        err = float((ta.g_flat - want).abs().max() / want.abs().max())
        buckets = ta.reducer.launched if ta.reducer is not None else 0
        ta.apply_update()
        tb.g_flat.copy_(want)
        tb.apply_update()
        tc.g_flat.copy_(want)
        tc.apply_update()
        perr = float((ta.p_flat - tb.p_flat).abs().max())
        ref = ta.p_flat.clone()
        dist.broadcast(ref, 0)
        same = bool(torch.equal(ref, ta.p_flat))
        if rank == 0:
            print("step %d: grad err vs plain sum %.2e (run-to-run noise of two identical local runs %.2e), "
                  "buckets %d, param diff %.2e, ranks identical %s" % (step, err, noise, buckets, perr, same),
                  flush=True)
        ok = ok and err < max(1e-4, 10 * noise) and same
    if "--step" in sys.argv:
        # FusedTrainer.step(): every all-reduce bucket is followed by the optimizer update of its arena range on the
        # communication stream.  Against the same steps done as forward_backward -> plain all-reduce -> one update:
        # identical parameters on every rank, and equal to the plain path up to the atomic summation order of the
        # split-K wgrad (Adam moves a weight by ~lr per step whatever the gradient's size: a few lr of slack)
        lr = 1e-4
        for t in (tb,):
            t.p_flat.copy_(ta.p_flat)
            t.m_flat.copy_(ta.m_flat)
            t.v_flat.copy_(ta.v_flat)
            t.t = ta.t
            invalidate_core(t.core)
        for step in range(3):
            ta.step(img, x0, tgt)
            tb.forward_backward(img, x0, tgt)
            dist.all_reduce(tb.g_flat)
            tb.apply_update()
            ref = ta.p_flat.clone()
            dist.broadcast(ref, 0)
            same = bool(torch.equal(ref, ta.p_flat))
            perr = float((ta.p_flat - tb.p_flat).abs().max())
            if rank == 0:
                print("fused step %d: ranks identical %s, max parameter difference to the plain path %.2e (lr %.0e), "
                      "%d all-reduce launches" % (step, same, perr, lr, ta.reducer.launched), flush=True)
            ok = ok and same and perr <= 20 * lr
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())

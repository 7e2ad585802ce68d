"""Loss-curve parity: N fused-trainer steps on the GPU vs the reference's Adam curve (golden fixture,
generated from the reference modules + torch.optim.Adam on CPU).  Prints per-step relative deviation."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rgb-proprioceptive-pose-estimator_b200"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import model_checks as mc  # noqa: E402
from oracle import pose_oracle as po  # noqa: E402


def run_curve(kind, steps=None, use_autograd=False, fixture=None):
    from pe_b200.trainer import FusedTrainer
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", fixture or "curve_%s.json" % kind)))
    steps = steps or len(fx["losses"])
    model = mc.build_model(kind).cuda().train()
    img, x0, tgt = po.synthetic_batch(kind, seed=1, **fx["shapes"])
    img, x0, tgt = img.cuda(), x0.cuda(), tgt.cuda()
    losses = []
    if use_autograd:
        from models.losses import PoseDistanceLoss
        crit = PoseDistanceLoss(**fx["loss_cfg"])
        opt = torch.optim.Adam(model.parameters(), lr=fx["lr"])
        for _ in range(steps):
            opt.zero_grad()
            if kind in ("td", "tdo"):
                model.reset_initial_state(img.shape[1])
            loss = crit(model(img, None, x0), tgt)
            loss.backward()
            opt.step()
            losses.append(float(loss))
    else:
        tr = FusedTrainer(model, lr=fx["lr"], **fx["loss_cfg"])
        for _ in range(steps):
            losses.append(float(tr.step(img, x0, tgt)))
    ref = fx["losses"][:steps]
    dev = [abs(a - b) / abs(b) for a, b in zip(losses, ref)]
    return losses, ref, dev


if __name__ == "__main__":
    kind = sys.argv[1] if len(sys.argv) > 1 else "no"
    fixture = sys.argv[2] if len(sys.argv) > 2 else None
    for mode in (False, True):
        losses, ref, dev = run_curve(kind, use_autograd=mode, fixture=fixture)
        print("== %s %s: max rel dev %.3e, mean %.3e, final ours %.4f ref %.4f" %
              (kind, "autograd+torch.optim.Adam" if mode else "FusedTrainer", max(dev), sum(dev) / len(dev), losses[-1], ref[-1]))
        for i in range(0, len(losses), max(1, len(losses) // 20)):
            print("  step %3d ours %.5f ref %.5f dev %.2e" % (i, losses[i], ref[i], dev[i]))

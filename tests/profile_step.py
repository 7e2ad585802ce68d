"""One or two fused training steps of the bench workload, for use under ncu (profiles/ recipes)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

if __name__ == "__main__":
    kind = sys.argv[1] if len(sys.argv) > 1 else "no"
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    from pe_b200.trainer import FusedTrainer
    model = bench.build(kind).cuda().train()
    tr = FusedTrainer(model, lr=1e-3, **bench.LOSS)
    seq = {"tdo": 20, "td": 10}.get(kind, 1)
    img, x0, tgt = bench.synth(kind, batch, seq, 1)
    img, x0, tgt = img.cuda(), x0.cuda(), tgt.cuda()
    tg = (x0, tgt) if kind in ("td", "n") else tgt
    for _ in range(steps):
        loss = tr.step(img, x0, tg)
    torch.cuda.synchronize()
    print("loss", float(loss))

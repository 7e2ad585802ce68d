"""Loss curve of the fused trainer at the reference's default learning rate 1e-3 (scripts/train_model.py:26,228) on
the data of tests/golden/curve_tdo.json, next to the reference's own curve (three CPU realisations that differ only in
thread count).  usage: python tests/curve_lr1e3.py [runs]"""
import json
import os
import sys

import torch

import model_checks as mc
from oracle import pose_oracle as po
from pe_b200.trainer import FusedTrainer

G = os.path.join(os.path.dirname(__file__), "golden")
fx = json.load(open(os.path.join(G, "curve_tdo.json")))
refs = [fx["losses"]] + [json.load(open(os.path.join(G, "curve_tdo_lr1e-3_threads%d.json" % t)))["losses"] for t in (1, 3)]
img, x0, tgt = po.synthetic_batch("tdo", seed=1, **fx["shapes"])
img, x0, tgt = img.cuda(), x0.cuda(), tgt.cuda()
for run in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    model = mc.build_model("tdo").cuda().train()
    tr = FusedTrainer(model, lr=fx["lr"], **fx["loss_cfg"])
    ours = [float(tr.step(img, x0, tgt)) for _ in range(30)]
    print("run", run, [round(v, 3) for v in ours])
    print("  dev vs ref[0]", [round(abs(a - b) / b, 3) for a, b in zip(ours, refs[0])])
    lo = [min(r[i] for r in refs) for i in range(30)]
    hi = [max(r[i] for r in refs) for i in range(30)]
    print("  outside reference band [lo/1.5, hi*1.5] at steps", [i for i in range(30) if not lo[i] / 1.5 <= ours[i] <= hi[i] * 1.5])

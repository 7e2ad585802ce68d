"""GPU probe: loss trajectory of the bench's config-2 training step under different tap-GEMM pairing rules (0 automatic,
3 = 256-column tiles only, 1 = never).  The trajectories must agree to TF32 noise.  Diagnostic only."""
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from pe_b200 import native  # noqa: E402
from pe_b200.trainer import FusedTrainer  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    lr = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-4
    modes = [int(m) for m in sys.argv[3].split(',')] if len(sys.argv) > 3 else [0, 3, 1]
    L = native.lib()
    dev = torch.device("cuda")
    out = {}
    for mode in modes:
        L.pe_debug_cta_group(mode)
        torch.manual_seed(0)
        model = bench.build("no").to(dev).train()
        tr = FusedTrainer(model, lr=lr, **bench.LOSS)
        img, x0, tgt = bench.synth("no", 256, None, 1)
        img, x0, tgt = img.to(dev), x0.to(dev), tgt.to(dev)
        losses = []
        for _ in range(steps):
            losses.append(float(tr.step(img, x0, tgt).item()))
        out[mode] = losses
        print("mode", mode, "device flag", L.pe_device_error(), flush=True)
        del model, tr
        torch.cuda.empty_cache()
    L.pe_debug_cta_group(0)
    for i in range(steps):
        print("%3d  %s" % (i, "  ".join("%12.4f" % out[m][i] for m in modes)))


if __name__ == "__main__":
    main()

"""A few eager (no CUDA graph) batch-1 rollout steps of the TDO estimator, for an ncu launch list."""
import sys
import torch
import model_checks as mc
from oracle import pose_oracle as po
from pe_b200.rollout import StreamingEstimator

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    model = mc.build_model("tdo").cuda().eval()
    est = StreamingEstimator(model, batch_size=n, use_graph=False)
    est.reset()
    img, x0, _ = po.synthetic_batch("tdo", n, s=1, seed=3)
    img, x0 = img.cuda(), x0.cuda()
    for _ in range(3):
        out = est.step(img, x0)
    torch.cuda.synchronize()
    print("ok", out.shape)

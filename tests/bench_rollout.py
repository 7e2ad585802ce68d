"""BASELINE config 5: batch-1 streaming rollout latency (CUDA graph) and a batch sweep, TDO estimator.
usage: python tests/bench_rollout.py [kind] [batches...]  -> JSON lines on stdout"""
import json
import statistics
import sys
import time

import torch

import model_checks as mc
from oracle import pose_oracle as po
from pe_b200.rollout import StreamingEstimator


RAW = False


def main():
    global RAW
    argv = [a for a in sys.argv[1:] if a != "--raw"]
    RAW = "--raw" in sys.argv
    for a in list(argv):
        if a == "--no-pdl":
            from pe_b200 import native
            native.lib().pe_debug_pdl(0)
            argv.remove(a)
        elif a.startswith("--epi-groups="):
            from pe_b200 import native
            native.lib().pe_debug_epilogue_groups(int(a.split("=")[1]))
            argv.remove(a)
        elif a.startswith("--cta-group="):
            from pe_b200 import native
            native.lib().pe_debug_cta_group(int(a.split("=")[1]))
            argv.remove(a)
        elif a.startswith("--min-bn="):
            from pe_b200 import native
            native.lib().pe_debug_min_bn(int(a.split("=")[1]))
            argv.remove(a)
    kind = argv[0] if argv else "tdo"
    batches = [int(a) for a in argv[1:]] or [1, 8, 64, 256, 1024]
    model = mc.build_model(kind).cuda().eval()
    for use_graph in (True, False):
        for N in batches:
            if not use_graph and N not in (1, batches[-1]):
                continue
            est = StreamingEstimator(model, batch_size=N, use_graph=use_graph, raw_hw=256 if RAW else None)
            est.reset()
            img, x0, _ = po.synthetic_batch(kind, N, s=1, seed=3) if kind in ("td", "tdo", "tdo_v2") else po.synthetic_batch(kind, N, seed=3)
            if RAW:   # uint8 HWC 256x256 frames as the renderer produces them; crop / scale / normalise on the GPU
                img = torch.randint(0, 256, (N, 256, 256, 3), dtype=torch.uint8)
            img_h, x0_h = img.pin_memory(), x0.pin_memory()
            step = est.step_raw if RAW else est.step
            for _ in range(5):
                out = step(img_h, x0_h)
            torch.cuda.synchronize()
            lat = []
            iters = 100 if N <= 64 else 20
            for _ in range(iters):
                t0 = time.perf_counter()
                out = step(img_h, x0_h)
                o = out[-1] if isinstance(out, tuple) else out
                _ = o.cpu()                                  # the rollout loop reads the pose every step
                lat.append((time.perf_counter() - t0) * 1e3)
            lat.sort()
            print(json.dumps({"workload": "rollout step (%s, eval, state carried%s)" % (kind, ", raw uint8 frames" if RAW else ""), "batch": N,
                              "cuda_graph": use_graph, "p50_ms": statistics.median(lat),
                              "p99_ms": lat[min(len(lat) - 1, int(0.99 * len(lat)))],
                              "frames_per_s": N / (statistics.median(lat) / 1e3),
                              "h2d_bytes_per_step": img_h.numel() * img_h.element_size() + x0_h.numel() * 4,
                              "d2h_bytes_per_step": N * 28}),
                  flush=True)


if __name__ == "__main__":
    main()

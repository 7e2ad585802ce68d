"""GPU probe: what bounds a 64-column output tile?  1x1 64->64 convolution over a 112x112 map (the stem's output
geometry) with the epilogue's staging stores / TMA stores / statistics switched off one at a time.  Diagnostic only."""
import sys

import torch

import kernel_checks as kc
from pe_b200 import native

P, S = kc.P, kc.S


def timeit(fn, iters=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    L = native.lib()
    for (H, ci, co) in ((112, 64, 64), (56, 64, 256)):
        x = torch.randn(B, H, H, ci, device="cuda")
        w = torch.randn(1, co, ci, device="cuda")
        y = torch.empty(B, H, H, co, device="cuda")
        stats = torch.zeros(2 * co, device="cuda", dtype=torch.float64)
        print("== 1x1 %d->%d over %dx%d, %d frames: output %.0f MB" % (ci, co, H, H, B, y.numel() * 4 / 1e6))
        for name, setup in (("pairs (default)", lambda: None), ("single CTA", lambda: L.pe_debug_cta_group(1)),
                            ("flags 1: no TMA store", lambda: L.pe_debug_flags(1)),
                            ("flags 2: no staging, no store", lambda: L.pe_debug_flags(2)),
                            ("flags 12: no operand loads", lambda: L.pe_debug_flags(12)),
                            ("flags 14: no loads, no staging / store", lambda: L.pe_debug_flags(14)),
                            ("4 epilogue groups", lambda: L.pe_debug_epilogue_groups(4)),
                            ("nout 2", lambda: L.pe_debug_pipeline(0, 2))):
            setup()
            t1 = timeit(lambda: L.pe_conv2d_fwd(P(x), P(w), P(y), B, H, H, ci, co, 1, 1, 1, 0, None, None, None, 0, 0, P(stats), S()))
            t2 = timeit(lambda: L.pe_conv2d_fwd(P(x), P(w), P(y), B, H, H, ci, co, 1, 1, 1, 0, None, None, None, 0, 0, None, S()))
            L.pe_debug_flags(0); L.pe_debug_cta_group(0); L.pe_debug_epilogue_groups(0); L.pe_debug_pipeline(0, 0)
            print("%-40s with stats %7.1f us   without %7.1f us" % (name, t1, t2), flush=True)
            L.pe_device_error_clear()


if __name__ == "__main__":
    main()

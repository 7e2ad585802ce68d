"""GPU probe: timing of the space-to-depth stem (pe_stem_conv_fwd / pe_stem_conv_wgrad) at the bench batch with the
operand loads switched off one at a time (garbage results, timing only).  Diagnostic only."""
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
    img = torch.randn(B, 3, 224, 224, device="cuda")
    s2d = torch.empty(B, 115, 115, 12, device="cuda")
    w = torch.randn(4, 64, 64, device="cuda")
    y = torch.empty(B * 112 * 112, 64, device="cuda")
    dw = torch.empty(4, 64, 64, device="cuda")
    stats = torch.zeros(128, device="cuda", dtype=torch.float64)
    L.pe_stem_s2d_pack(P(img), P(s2d), B, 224, 224, 1, S())
    fwd = lambda: L.pe_stem_conv_fwd(P(s2d), P(w), P(y), B, 224, 224, 64, None, None, 0, 0, P(stats), S())
    wg = lambda: L.pe_stem_conv_wgrad(P(s2d), P(y), P(dw), B, 224, 224, 64, S())
    print("s2d pack %.1f us" % timeit(lambda: L.pe_stem_s2d_pack(P(img), P(s2d), B, 224, 224, 1, S())))
    for name, setup in (("default", lambda: None), ("no pairs", lambda: L.pe_debug_cta_group(1)),
                        ("flags 4 (no A)", lambda: L.pe_debug_flags(4)), ("flags 8 (no B)", lambda: L.pe_debug_flags(8)),
                        ("flags 12 (no loads)", lambda: L.pe_debug_flags(12)),
                        ("nout 2", lambda: L.pe_debug_pipeline(0, 2)), ("nout 4", lambda: L.pe_debug_pipeline(0, 4)),
                        ("4 epilogue groups", lambda: L.pe_debug_epilogue_groups(4))):
        setup()
        print("%-22s fwd %7.1f us   wgrad %7.1f us  (flag %d)" % (name, timeit(fwd), timeit(wg), L.pe_device_error()), flush=True)
        L.pe_debug_flags(0); L.pe_debug_cta_group(0); L.pe_debug_pipeline(0, 0); L.pe_debug_epilogue_groups(0)
        L.pe_device_error_clear()
    # for scale: a plain 1x1 64->64 convolution over the same 112x112 map (same output bytes, K = 64 instead of 256)
    x = torch.randn(B, 112, 112, 64, device="cuda")
    wt = torch.randn(1, 64, 64, device="cuda")
    print("1x1 64->64 @112 fwd %.1f us" % timeit(lambda: L.pe_conv2d_fwd(P(x), P(wt), P(y), B, 112, 112, 64, 64, 1, 1, 1, 0, None, None, None, 0, 0, P(stats), S())))


if __name__ == "__main__":
    main()

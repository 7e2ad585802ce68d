"""GPU probe: what bounds the 56x56 64->64 3x3 convolution (40 % of its tensor bound)?  Times forward / dgrad with the
operand loads switched off one at a time (debug flags 4 = no A loads, 8 = no B loads: results are garbage, only the
time matters), with different ring depths, and with the haloed-tile path.  Diagnostic only, not a test."""
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
    for (H, ci, co) in ((56, 64, 64), (28, 128, 128), (14, 256, 256)):
        x = torch.randn(B, H, H, ci, device="cuda")
        w = torch.randn(co, ci, 3, 3, device="cuda")
        tck, tkc = kc.pack(w)
        y = torch.empty(B, H, H, co, device="cuda")
        dx = torch.empty_like(x)
        stats = torch.zeros(2 * co, device="cuda", dtype=torch.float64)

        def fwd(st=True):
            L.pe_conv2d_fwd(P(x), P(tck), P(y), B, H, H, ci, co, 3, 3, 1, 1, None, None, None, 0, 0,
                            P(stats) if st else None, S())

        def dgrad():
            L.pe_conv2d_dgrad(P(y), P(tkc), P(dx), B, H, H, ci, co, 3, 3, 1, 1, None, None, S())

        print("== %dx%d %d->%d 3x3, %d frames" % (H, H, ci, co, B))
        for name, setup in (("baseline", lambda: None),
                            ("no stats", None),
                            ("flags 2048 (old 16/32 KB B regions)", lambda: L.pe_debug_flags(2048)),
                            ("flags 4 (no A loads)", lambda: L.pe_debug_flags(4)),
                            ("flags 8 (no B loads)", lambda: L.pe_debug_flags(8)),
                            ("flags 12 (no loads)", lambda: L.pe_debug_flags(12)),
                            ("pipeline 2 stages", lambda: L.pe_debug_pipeline(2, 0)),
                            ("pipeline 4 stages", lambda: L.pe_debug_pipeline(4, 0)),
                            ("pipeline 6 stages", lambda: L.pe_debug_pipeline(6, 0)),
                            ("pipeline nout 2", lambda: L.pe_debug_pipeline(0, 2)),
                            ("pipeline nout 4", lambda: L.pe_debug_pipeline(0, 4)),
                            ("CTA pairs", lambda: L.pe_debug_cta_group(2)),
                            ("CTA pairs, nout 2", lambda: (L.pe_debug_cta_group(2), L.pe_debug_pipeline(0, 2))),
                            ("CTA pairs, nout 4", lambda: (L.pe_debug_cta_group(2), L.pe_debug_pipeline(0, 4))),
                            ("halo path", lambda: L.pe_debug_conv_halo(1)),
                            ("halo path, no B loads", lambda: (L.pe_debug_conv_halo(1), L.pe_debug_flags(8))),
                            ):
            if setup is None:
                t1, t2 = timeit(lambda: fwd(False)), float("nan")
            else:
                setup()
                t1, t2 = timeit(fwd), timeit(dgrad)
            L.pe_debug_flags(0)
            L.pe_debug_pipeline(0, 0)
            L.pe_debug_conv_halo(0)
            L.pe_debug_cta_group(0)
            print("%-32s fwd %7.1f us   dgrad %7.1f us   (device flag %d)" % (name, t1, t2, L.pe_device_error()), flush=True)
            L.pe_device_error_clear()


if __name__ == "__main__":
    main()

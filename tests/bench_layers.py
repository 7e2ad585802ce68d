"""Per-layer timing of the tap-GEMM entry points at the bench batch size, with roofline percentages.
Diagnostic (GPU): python tests/bench_layers.py [batch] [dbg_flags...]"""
import sys

import torch

import kernel_checks as kc
from pe_b200 import native

P, S = kc.P, kc.S
HBM, TF32 = 6455.6e9, 709e12

# (H, Cin, Cout, k, stride)  -- distinct ResNet-50 conv shapes
SHAPES = [(56, 64, 64, 1, 1), (56, 64, 64, 3, 1), (56, 64, 256, 1, 1), (56, 256, 64, 1, 1), (56, 256, 128, 1, 1),
          (56, 128, 128, 3, 2), (28, 128, 512, 1, 1), (56, 256, 512, 1, 2), (28, 512, 128, 1, 1), (28, 128, 128, 3, 1),
          (28, 512, 256, 1, 1), (28, 256, 256, 3, 2), (14, 256, 1024, 1, 1), (14, 1024, 256, 1, 1), (14, 256, 256, 3, 1),
          (14, 512, 512, 3, 2), (7, 512, 2048, 1, 1), (7, 2048, 512, 1, 1), (7, 512, 512, 3, 1)]


def timeit(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3   # us


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    flags = [int(a) for a in sys.argv[2:]] or [0]   # flags >= 100 encode a pipeline split: 100*stages + nout
    which = "fwd"
    L = native.lib()
    print("%-28s %8s %8s | %s" % ("shape", "mem_us", "flop_us", "  ".join("f=%d: fwd dgrad wgrad" % f for f in flags)))
    tot = {f: [0.0, 0.0, 0.0] for f in flags}
    for (H, ci, co, k, st) in SHAPES:
        pad = (k - 1) // 2
        Ho = (H + 2 * pad - k) // st + 1
        x = torch.randn(B, H, H, ci, device="cuda")
        w = torch.randn(co, ci, k, k, device="cuda")
        tck, tkc = kc.pack(w)
        y = torch.empty(B, Ho, Ho, co, device="cuda")
        dx = torch.empty_like(x)
        dw = torch.empty(k * k, co, ci, device="cuda")
        stats = torch.zeros(2 * co, device="cuda", dtype=torch.float64)
        mem = (x.numel() + y.numel()) * 4 / HBM * 1e6
        flop = 2.0 * B * Ho * Ho * co * ci * k * k / TF32 * 1e6
        cols = []
        for f in flags:
            if f == 5001:
                L.pe_debug_cta_group(1)
            elif f == 5002:
                L.pe_debug_cta_group(2)
            elif f == 4000:
                L.pe_debug_epilogue_groups(4)
            elif f == 4006:
                L.pe_debug_epilogue_groups(6)
            elif f == 2002:
                L.pe_debug_epilogue_groups(2)
            elif f == 2000:
                L.pe_debug_wgrad_halo(0)
            elif f == 1128:
                L.pe_debug_max_bn(128)
            elif f >= 100:
                L.pe_debug_pipeline(f // 100, f % 100)
            else:
                L.pe_debug_flags(f)
            t1 = timeit(lambda: L.pe_conv2d_fwd(P(x), P(tck), P(y), B, H, H, ci, co, k, k, st, pad, None, None, None, 0, 0, P(stats), S()))
            t2 = timeit(lambda: L.pe_conv2d_dgrad(P(y), P(tkc), P(dx), B, H, H, ci, co, k, k, st, pad, None, None, S()))
            t3 = timeit(lambda: L.pe_conv2d_wgrad(P(x), P(y), P(dw), B, H, H, ci, co, k, k, st, pad, S()))
            extra = ""
            if k == 1 and st == 1 and ci % 32 == 0:
                # dgrad with the masked residual epilogue (identity branch of a residual join)
                bits = torch.randint(-2 ** 31, 2 ** 31 - 1, ((x.numel() // 4 + 31) // 32 * 4,), device="cuda",
                                     dtype=torch.int32)
                res = torch.randn_like(x)
                t4 = timeit(lambda: L.pe_conv2d_dgrad(P(y), P(tkc), P(dx), B, H, H, ci, co, k, k, st, pad, P(res),
                                                      P(bits), S()))
                extra = " (+res %4.0f)" % t4
            cols.append("%6.0f %6.0f %6.0f%s" % (t1, t2, t3, extra))
            for i, t in enumerate((t1, t2, t3)):
                tot[f][i] += t
            L.pe_debug_max_bn(256)
            L.pe_debug_pipeline(0, 0)
            L.pe_debug_flags(0)
            L.pe_debug_wgrad_halo(1)
            L.pe_debug_epilogue_groups(0)
            L.pe_debug_cta_group(0)
        L.pe_debug_flags(0)
        L.pe_debug_pipeline(0, 0)
        L.pe_debug_max_bn(256)
        print("%-28s %8.0f %8.0f | %s" % ("%dx%d %d->%d k%d s%d" % (H, H, ci, co, k, st), mem, flop, "   ".join(cols)), flush=True)
    for f in flags:
        print("flags %d totals (distinct shapes, us): fwd %.0f dgrad %.0f wgrad %.0f" % (f, *tot[f]))
    print("device error flag:", L.pe_device_error())


if __name__ == "__main__":
    main()

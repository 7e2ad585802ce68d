"""Timing probe (GPU): store-bound 1x1 forward convs (with BN statistics) under different tile widths / smem splits."""
import torch
import kernel_checks as kc
from pe_b200 import native
from bench_layers import timeit

P, S = kc.P, kc.S
L = native.lib()
B = 256
for (H, ci, co) in ((56, 64, 256), (28, 128, 512), (14, 256, 1024), (7, 512, 2048), (14, 1024, 256)):
    x = torch.randn(B, H, H, ci, device="cuda")
    y = torch.empty(B, H, H, co, device="cuda")
    w = torch.randn(co, ci, 1, 1, device="cuda")
    tck, tkc = kc.pack(w)
    stats = torch.zeros(2 * co, device="cuda", dtype=torch.float64)
    row = []
    for (g, st, nout, maxbn) in ((0, 0, 0, 256), (2, 0, 0, 256), (4, 0, 0, 256), (4, 2, 8, 128), (4, 3, 4, 128), (2, 2, 4, 128), (2, 3, 4, 256),
                                 (4, 2, 4, 256), (4, 1, 8, 256)):
        L.pe_debug_epilogue_groups(g)
        L.pe_debug_pipeline(st, nout)
        L.pe_debug_max_bn(maxbn)
        t = timeit(lambda: L.pe_conv2d_fwd(P(x), P(tck), P(y), B, H, H, ci, co, 1, 1, 1, 0, None, None, None, 0, 0, P(stats), S()))
        row.append("[G%d st%d no%d bn%d] %.0f" % (g, st, nout, maxbn, t))
    L.pe_debug_epilogue_groups(0); L.pe_debug_pipeline(0, 0); L.pe_debug_max_bn(256)
    print("H%d %d->%d: %s" % (H, ci, co, "  ".join(row)), flush=True)
print("flag", L.pe_device_error())

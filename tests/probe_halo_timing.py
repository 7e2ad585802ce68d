"""Timing probe (GPU): where the haloed wgrad spends its time (pe_debug_flags ablations)."""
import torch
import kernel_checks as kc
from pe_b200 import native
from bench_layers import timeit

P, S = kc.P, kc.S
L = native.lib()
import sys
for (H, c) in ((56, 64), (28, 128), (14, 256)):
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    x = torch.randn(B, H, H, c, device="cuda")
    y = torch.randn(B, H, H, c, device="cuda")
    dw = torch.empty(9, c, c, device="cuda")
    for halo in (1, 0):
        L.pe_debug_wgrad_halo(halo)
        row = []
        for f in (0, 12, 16, 92):
            L.pe_debug_flags(f)
            t = timeit(lambda: L.pe_conv2d_wgrad(P(x), P(y), P(dw), B, H, H, c, c, 3, 3, 1, 1, S()))
            row.append("f%d=%.0f" % (f, t))
        L.pe_debug_flags(0)
        print("H%d C%d halo=%d: %s" % (H, c, halo, "  ".join(row)), flush=True)
L.pe_debug_wgrad_halo(1)
print("flag", L.pe_device_error())

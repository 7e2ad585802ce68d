"""Timing probe (GPU): dgrad with the masked-residual epilogue under different smem splits."""
import torch
import kernel_checks as kc
from pe_b200 import native
from bench_layers import timeit

P, S = kc.P, kc.S
L = native.lib()
B = 256
for (H, ci, co) in ((56, 256, 64), (28, 512, 128), (14, 1024, 256)):
    x = torch.randn(B, H, H, ci, device="cuda")
    y = torch.randn(B, H, H, co, device="cuda")
    dx = torch.empty_like(x)
    w = torch.randn(co, ci, 1, 1, device="cuda")
    tck, tkc = kc.pack(w)
    res = torch.randn_like(x)
    bits = torch.randint(-2 ** 31, 2 ** 31 - 1, ((x.numel() // 4 + 31) // 32 * 4,), device="cuda", dtype=torch.int32)
    row = []
    for (tma, st, nout, maxbn, flags) in ((-1, 0, 0, 256, 0), (1, 0, 0, 256, 0), (1, 2, 8, 128, 0), (1, 3, 4, 128, 0),
                                         (1, 2, 4, 128, 0), (1, 0, 0, 256, 1)):
        L.pe_debug_residual_tma(max(tma, 0))
        L.pe_debug_pipeline(st, nout)
        L.pe_debug_max_bn(maxbn)
        L.pe_debug_flags(flags)
        r, m = (None, None) if tma < 0 else (P(res), P(bits))
        t = timeit(lambda: L.pe_conv2d_dgrad(P(y), P(tkc), P(dx), B, H, H, ci, co, 1, 1, 1, 0, r, m, S()))
        row.append("[tma%d st%d no%d bn%d f%d] %.0f" % (tma, st, nout, maxbn, flags, t))
    L.pe_debug_residual_tma(1); L.pe_debug_pipeline(0, 0); L.pe_debug_max_bn(256); L.pe_debug_flags(0)
    print("H%d %d<-%d: %s" % (H, ci, co, "  ".join(row)), flush=True)
print("flag", L.pe_device_error())

"""A few launches of one tap-GEMM configuration for `ncu --set full` (profiles/ recipes).
usage: profile_conv.py B H Cin Cout k stride [mode: fwd|dgrad|wgrad] [stats 0|1] [wgrad halo mode 0|1]"""
import sys

import torch

import kernel_checks as kc
from pe_b200 import native

if __name__ == "__main__":
    a = sys.argv[1:]
    B, H, ci, co, k, st = (int(v) for v in a[:6])
    mode = a[6] if len(a) > 6 else "fwd"
    use_stats = int(a[7]) if len(a) > 7 else 1
    L, P, S = native.lib(), kc.P, kc.S
    if len(a) > 8:
        L.pe_debug_wgrad_halo(int(a[8]))
    pad = (k - 1) // 2
    Ho = (H + 2 * pad - k) // st + 1
    x = torch.randn(B, H, H, ci, device="cuda")
    w = torch.randn(co, ci, k, k, device="cuda")
    tck, tkc = kc.pack(w)
    y = torch.randn(B, Ho, Ho, co, device="cuda")
    dx = torch.empty_like(x)
    dw = torch.empty(k * k, co, ci, device="cuda")
    stats = torch.zeros(2 * co, device="cuda", dtype=torch.float64)
    for _ in range(3):
        if mode == "fwd":
            L.pe_conv2d_fwd(P(x), P(tck), P(y), B, H, H, ci, co, k, k, st, pad, None, None, None, 0, 0,
                            P(stats) if use_stats else None, S())
        elif mode == "dgrad":
            L.pe_conv2d_dgrad(P(y), P(tkc), P(dx), B, H, H, ci, co, k, k, st, pad, None, None, S())
        else:
            L.pe_conv2d_wgrad(P(x), P(y), P(dw), B, H, H, ci, co, k, k, st, pad, S())
    torch.cuda.synchronize()
    print("done", L.pe_device_error())

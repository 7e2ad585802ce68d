"""A few launches of the space-to-depth stem for `ncu --set full` (profiles/ recipes).
usage: profile_stem.py B [fwd|wgrad]"""
import sys

import torch

import kernel_checks as kc
from pe_b200 import native

if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    mode = sys.argv[2] if len(sys.argv) > 2 else "fwd"
    L, P, S = native.lib(), kc.P, kc.S
    img = torch.randn(B, 3, 224, 224, device="cuda")
    s2d = torch.empty(B, 115, 115, 12, device="cuda")
    w = torch.randn(4, 64, 64, device="cuda")
    y = torch.randn(B * 112 * 112, 64, device="cuda")
    dw = torch.empty(4, 64, 64, device="cuda")
    stats = torch.zeros(128, device="cuda", dtype=torch.float64)
    L.pe_stem_s2d_pack(P(img), P(s2d), B, 224, 224, 1, S())
    for _ in range(3):
        if mode == "fwd":
            L.pe_stem_conv_fwd(P(s2d), P(w), P(y), B, 224, 224, 64, None, None, 0, 0, P(stats), S())
        else:
            L.pe_stem_conv_wgrad(P(s2d), P(y), P(dw), B, 224, 224, 64, S())
    torch.cuda.synchronize()
    print("done", L.pe_device_error())

"""Probe (GPU): haloed-tile 3x3 stride-1 forward / dgrad (pe_debug_conv_halo) -- correctness against the fp64 reference
and timing against the per-tap box path."""
import torch
import kernel_checks as kc
from pe_b200 import native
from bench_layers import timeit

P, S = kc.P, kc.S
L = native.lib()
torch.backends.cudnn.allow_tf32 = False
for on in (1, 0):
    L.pe_debug_conv_halo(on)
    for args in ((2, 56, 56, 64, 64, 3, 1), (3, 28, 28, 128, 128, 3, 1), (2, 16, 16, 32, 32, 3, 1), (2, 24, 24, 64, 96, 3, 1),
                 (2, 14, 14, 128, 64, 3, 1)):
        rows = kc.check_conv(*args)
        torch.cuda.synchronize()
        for name, err, tol in rows:
            if ("conv_fwd" in name or "conv_dgrad" in name) and "+" not in name:
                print("halo=%d %-44s err %.3e %s" % (on, name, err, "ok" if err <= tol else "FAIL"), flush=True)
    print("device error flag:", L.pe_device_error())
    L.pe_device_error_clear()
B = 256
for (H, c) in ((56, 64), (28, 128)):
    x = torch.randn(B, H, H, c, device="cuda")
    y = torch.empty(B, H, H, c, device="cuda")
    dx = torch.empty_like(x)
    w = torch.randn(c, c, 3, 3, device="cuda")
    tck, tkc = kc.pack(w)
    stats = torch.zeros(2 * c, device="cuda", dtype=torch.float64)
    row = []
    for on in (0, 1 + (1 << 4), 1 + (3 << 4), 1 + (9 << 4)):
        L.pe_debug_conv_halo(on)
        t1 = timeit(lambda: L.pe_conv2d_fwd(P(x), P(tck), P(y), B, H, H, c, c, 3, 3, 1, 1, None, None, None, 0, 0, P(stats), S()))
        t2 = timeit(lambda: L.pe_conv2d_dgrad(P(y), P(tkc), P(dx), B, H, H, c, c, 3, 3, 1, 1, None, None, S()))
        row.append("halo=%d btaps=%d fwd %.0f dgrad %.0f" % (on & 1, on >> 4, t1, t2))
    print("H%d C%d: %s" % (H, c, "   ".join(row)), flush=True)
L.pe_debug_conv_halo(0)
print("flag", L.pe_device_error())

"""CTA-pair (tcgen05 cta_group::2) path of the tap-GEMM forced on for conv / dense shapes, checked against fp64 torch.
Diagnostic (GPU): python tests/probe_cta_pair.py"""
import sys

import torch

import kernel_checks as kc
from pe_b200 import native


def main():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    L = native.lib()
    L.pe_debug_cta_group(2)
    nfail = 0
    cases = [lambda: kc.check_conv(2, 56, 56, 64, 256, 1, 1), lambda: kc.check_conv(4, 14, 14, 256, 256, 3, 1),
             lambda: kc.check_conv(3, 28, 28, 256, 512, 1, 2), lambda: kc.check_conv(2, 14, 14, 1024, 256, 1, 1),
             lambda: kc.check_conv(5, 7, 7, 512, 2048, 1, 1), lambda: kc.check_conv(3, 7, 7, 512, 512, 3, 1),
             lambda: kc.check_conv(2, 56, 56, 64, 64, 3, 1), lambda: kc.check_conv(2, 28, 28, 128, 128, 3, 1),
             lambda: kc.check_conv_fused_eval(3, 14, 14, 256, 1024, 1), lambda: kc.check_conv_fused_eval(2, 28, 28, 128, 512, 1),
             lambda: kc.check_linear(300, 512, 256), lambda: kc.check_linear(1000, 2048, 3680, relu=True),
             lambda: kc.check_linear(129, 256, 64)]
    for fn in cases:
        try:
            rows = fn()
            torch.cuda.synchronize()
        except Exception as e:
            rows = [("EXCEPTION %r" % (e,), float("inf"), 0.0)]
        for name, err, tol in rows:
            ok = err <= tol
            nfail += (not ok)
            print("%-4s %-52s err %.3e tol %.1e" % ("ok" if ok else "FAIL", name, err, tol), flush=True)
        code = L.pe_device_error()
        if code:
            print("device error flag:", code, flush=True)
            L.pe_device_error_clear()
    L.pe_debug_cta_group(0)
    print("FAILED: %d" % nfail)
    return nfail


if __name__ == "__main__":
    sys.exit(1 if main() else 0)

"""GPU probe: sweep UMMA shared-memory descriptor strides (debug override) to confirm the canonical
K-major / MN-major SWIZZLE_128B encodings used by pe_tapgemm.cu.  Diagnostic only, not a test."""
import sys

import torch

import kernel_checks as kc
from pe_b200 import native


def main():
    L = native.lib()
    print("== default descriptors ==")
    for fn in (lambda: kc.check_linear(256, 128, 128), lambda: kc.check_linear_wgrad(256, 128, 128)):
        for name, err, tol in fn():
            print("%-40s err %.3e" % (name, err))
    print("== MN-major sweep (wgrad) ==")
    for lbo in (4096, 512, 1024, 128):
        for sbo in (512, 1024, 256, 4096):
            L.pe_debug_desc_override(lbo, sbo, lbo, sbo)
            L.pe_device_error_clear()
            try:
                (name, err, tol), = kc.check_linear_wgrad(256, 128, 128)
            except Exception as e:
                err = float("nan")
            torch.cuda.synchronize()
            print("wgrad lbo %5d sbo %5d -> err %.3e flag %d" % (lbo, sbo, err, L.pe_device_error()), flush=True)
    print("== K-major sweep (fwd) ==")
    for lbo in (0, 16, 1024):
        for sbo in (1024,):
            L.pe_debug_desc_override(lbo, sbo, lbo, sbo)
            (name, err, tol), = kc.check_linear(256, 128, 128)
            torch.cuda.synchronize()
            print("fwd   lbo %5d sbo %5d -> err %.3e" % (lbo, sbo, err), flush=True)
    L.pe_debug_desc_override(-1, -1, -1, -1)
    print("device error flag:", L.pe_device_error())


if __name__ == "__main__":
    sys.exit(main())

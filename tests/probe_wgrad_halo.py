"""Probe (GPU): multi-tap haloed wgrad (tap-GEMM mode 2) against the one-tap-per-item path and an fp64 reference.
usage: python tests/probe_wgrad_halo.py"""
import torch

import kernel_checks as kc
from pe_b200 import native

if __name__ == "__main__":
    L = native.lib()
    torch.backends.cudnn.allow_tf32 = False
    for mode in (1, 2, 0):
        L.pe_debug_wgrad_halo(mode)
        for args in ((2, 56, 56, 64, 64, 3, 1), (3, 28, 28, 128, 128, 3, 1), (5, 14, 14, 256, 256, 3, 1),
                     (3, 7, 7, 512, 512, 3, 1), (2, 16, 16, 32, 32, 3, 1), (2, 24, 24, 64, 96, 3, 1)):
            rows = kc.check_conv(*args)
            torch.cuda.synchronize()
            for name, err, tol in rows:
                if "wgrad" in name and "unpack" not in name:
                    print("halo=%d %-44s err %.3e %s" % (mode, name, err, "ok" if err <= tol else "FAIL"), flush=True)
        print("device error flag:", L.pe_device_error())
        L.pe_device_error_clear()
    L.pe_debug_wgrad_halo(1)
